// tools/membw.cu — read-bandwidth microbenchmark (run under gpurun): how does achieved HBM read bandwidth depend on
// the size of the contiguous chunk each warp streams before jumping to another region?  Mirrors the FDL access
// pattern of the MAC phase (rows of B*8 bytes per (stream, speaker, slot)).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/membw tools/membw.cu && /tmp/membw
#include <cstdio>
#include <cuda_runtime.h>

// Each warp owns a private region of `region` bytes and reads it in chunks of 512 B (32 lanes x 16 B) with `ilp` loads in flight.
// warps of a CTA own consecutive... `pattern`: 0 = fully linear grid-stride; 1 = per-warp private sequential regions.
template <int ILP>
__global__ void k_read(const float4 *__restrict__ p, size_t n4, size_t region4, int pattern, float *out)
{
    const size_t warp = (size_t)(blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    if (pattern == 0) {
        const size_t stride = (size_t)gridDim.x * blockDim.x;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + (ILP - 1) * stride < n4; i += ILP * stride) {
            float4 v[ILP];
#pragma unroll
            for (int k = 0; k < ILP; ++k) v[k] = __ldcs(p + i + k * stride);
#pragma unroll
            for (int k = 0; k < ILP; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
        }
    } else {
        // regions are dealt round-robin to warps; inside a region the warp reads sequentially
        const size_t nregions = n4 / region4;
        for (size_t r = warp; r < nregions; r += nwarps) {
            const float4 *q = p + r * region4 + lane;
            for (size_t i = 0; i + (ILP - 1) * 32 < region4; i += ILP * 32) {
                float4 v[ILP];
#pragma unroll
                for (int k = 0; k < ILP; ++k) v[k] = __ldcs(q + i + k * 32);
#pragma unroll
                for (int k = 0; k < ILP; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

int main()
{
    const size_t bytes = (size_t)2 << 30;   // 2 GiB >> L2
    float4 *p; float *out;
    cudaMalloc(&p, bytes); cudaMalloc(&out, 4);
    cudaMemset(p, 0, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const size_t n4 = bytes / 16;
    auto run = [&](const char *name, int pattern, size_t region_bytes, int blocks, int threads) {
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(a);
            k_read<4><<<blocks, threads>>>(p, n4, region_bytes / 16, pattern, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        printf("%-28s region %8zu B  grid %5d x %4d : %8.1f GB/s\n", name, region_bytes, blocks, threads, bytes / best / 1e6);
    };
    for (int occ : {4, 8, 16}) run("linear grid-stride", 0, 0, 148 * occ, 256);
    for (size_t region : {(size_t)2048, (size_t)8192, (size_t)34816, (size_t)131072, (size_t)1 << 20})
        for (int occ : {8, 16}) run("per-warp sequential regions", 1, region, 148 * occ, 128);
    // write+read copy for reference
    float4 *q; cudaMalloc(&q, bytes / 2);
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) { cudaEventRecord(a); cudaMemcpyAsync(q, p, bytes / 2, cudaMemcpyDeviceToDevice); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    printf("cudaMemcpy D2D 1 GiB (read+write bytes): %8.1f GB/s\n", bytes / best / 1e6);
    return 0;
}
