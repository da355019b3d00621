"""GPU-side: what does this box's PCIe give for the e2e step's transfer sizes (134 MB in, 34 MB out, pinned)?"""
import time, torch
n_in, n_out = 4096 * 8 * 1024, 4096 * 2 * 1024
hin = [torch.empty(n_in, dtype=torch.float32).pin_memory() for _ in range(2)]
hout = [torch.empty(n_out, dtype=torch.float32).pin_memory() for _ in range(2)]
din = [torch.empty(n_in, dtype=torch.float32, device="cuda") for _ in range(2)]
dout = [torch.empty(n_out, dtype=torch.float32, device="cuda") for _ in range(2)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, iters=30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(iters):
        if h2d:
            with torch.cuda.stream(s1): din[i % 2].copy_(hin[i % 2], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): hout[i % 2].copy_(dout[i % 2], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / iters
    return dt
for name, a, b in [("H2D alone", True, False), ("D2H alone", False, True), ("H2D + D2H concurrent", True, True)]:
    run(a, b, 5); dt = run(a, b)
    print(f"{name:24s} {dt*1e3:6.3f} ms/step  H2D {n_in*4/dt/1e9 if a else 0:5.1f} GB/s  D2H {n_out*4/dt/1e9 if b else 0:5.1f} GB/s")
