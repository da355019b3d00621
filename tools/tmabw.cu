// tools/tmabw.cu — bulk-copy (cp.async.bulk, "TMA-class") supply-rate microbenchmark (run under gpurun).
// Question it answers for the persistent kernel's ring: how many GB/s can ONE SM pull through cp.async.bulk into shared
// memory as a function of (copy size, copies per stage, ring depth, number of issuing warps), when HBM is not the limit
// (few CTAs) and when it is (148 CTAs)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tmabw tools/tmabw.cu && /tmp/tmabw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// stage = ncopies copies of copy_bytes each (source addresses scattered: each copy reads its own 64 KiB-strided row)
__global__ void __launch_bounds__(32 * 16, 1) k_ring(const char *src, size_t src_bytes, int copy_bytes, int ncopies, int stages, int producers,
                                                   int consumers, int iters, int read_smem, float *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int stage_bytes = copy_bytes * ncopies;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)stages * stage_bytes);
    uint64_t *empty = full + stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], consumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    if (warp < producers) {
        if (lane == 0) {
            // producer w owns ring slots w, w + producers, ...
            size_t off = ((size_t)blockIdx.x * 7919 + warp * 104729) * 65536 % src_bytes;
            for (int it = 0; it < iters; ++it) {
                for (int st = warp; st < stages; st += producers) {
                    mbar_wait(&empty[st], (unsigned)((it & 1) ^ 1));
                    mbar_expect_tx(&full[st], (unsigned)stage_bytes);
                    for (int c = 0; c < ncopies; ++c) {
                        bulk_g2s(smem + (size_t)st * stage_bytes + (size_t)c * copy_bytes, src + off, (unsigned)copy_bytes, &full[st]);
                        off += 65536 + 4096;
                        if (off + copy_bytes > src_bytes) off = (off + copy_bytes) % 65536;
                    }
                }
            }
        }
    } else if (warp < producers + consumers) {
        for (int it = 0; it < iters; ++it) {
            for (int st = 0; st < stages; ++st) {
                mbar_wait(&full[st], (unsigned)(it & 1));
                if (read_smem) {
                    const float4 *p = reinterpret_cast<const float4 *>(smem + (size_t)st * stage_bytes);
                    for (int i = (warp - producers) * 32 + lane; i < stage_bytes / 16; i += consumers * 32) { const float4 v = p[i]; acc += v.x + v.w; }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

int main()
{
    const size_t bytes = (size_t)4 << 30;
    char *p; float *out;
    cudaMalloc(&p, bytes); cudaMalloc(&out, 4);
    cudaMemset(p, 0, bytes);
    cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](int ctas, int copy_bytes, int ncopies, int ring_kb, int producers, int consumers, int read_smem) {
        int stages = ring_kb * 1024 / (copy_bytes * ncopies);
        stages = stages / producers * producers;
        if (stages < producers) return;
        const size_t smem = (size_t)stages * copy_bytes * ncopies + 2 * stages * 8 + 64;
        const int iters = (int)(((size_t)24 << 20) / ((size_t)stages * copy_bytes * ncopies));   // ~24 MiB per CTA
        float best = 1e9f;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(a);
            k_ring<<<ctas, 32 * (producers + consumers), smem>>>(p, bytes, copy_bytes, ncopies, stages, producers, consumers, iters, read_smem, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        const double per_cta = (double)iters * stages * copy_bytes * ncopies;
        printf("ctas %3d copy %5d B x %2d/stage ring %3d KB (%2d stages) prod %d cons %d read %d : %6.1f GB/s per SM, %7.1f GB/s total, %5.0f cyc/copy%s\n",
               ctas, copy_bytes, ncopies, ring_kb, stages, producers, consumers, read_smem, per_cta / best / 1e6, per_cta * ctas / best / 1e6,
               best * 1e-3 * 1.965e9 / ((double)iters * stages * ncopies), e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    for (int ctas : {32, 108, 148}) {
        for (int copy : {512, 1024, 2048, 4096, 8192})
            for (int prod : {1, 4}) run(ctas, copy, 6, 144, prod, 4, 0);
        run(ctas, 2048, 6, 96, 4, 4, 0);
        run(ctas, 2048, 6, 192, 4, 4, 0);
        run(ctas, 4096, 6, 192, 4, 4, 0);
        run(ctas, 4096, 3, 144, 2, 4, 0);
        run(ctas, 2048, 6, 144, 8, 4, 0);
        run(ctas, 2048, 6, 144, 4, 4, 1);
        run(ctas, 4096, 6, 144, 4, 8, 1);
    }
    return 0;
}
