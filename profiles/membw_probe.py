"""Write-only / copy / read-only HBM bandwidth probe (torch ops, CUDA events) -- context for the roofline of a
write-dominated kernel: MEASURED_PEAKS.json's hbm_gbs is a copy (read + write bytes)."""
import json
import torch

N = 1 << 30
a = torch.empty(N, dtype=torch.uint8, device="cuda")
b = torch.empty(N, dtype=torch.uint8, device="cuda")
af = a.view(torch.float32)


def timeit(fn, reps=10):
    best = 1e9
    for _ in range(3):
        fn()
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best * 1e-3


out = {}
t = timeit(lambda: a.fill_(1))
out["write_only_gbs_fill_u8"] = N / t / 1e9
t = timeit(lambda: af.fill_(1.0))
out["write_only_gbs_fill_f32"] = N / t / 1e9
t = timeit(lambda: b.copy_(a))
out["copy_gbs_read_plus_write"] = 2 * N / t / 1e9
t = timeit(lambda: af.sum())
out["read_only_gbs_sum_f32"] = N / t / 1e9
# 80 MB working set (fits L2): write rate into a warm L2
c = torch.empty(80 << 20, dtype=torch.uint8, device="cuda").view(torch.float32)
t = timeit(lambda: c.fill_(2.0), reps=30)
out["write_80MB_l2_resident_gbs"] = (80 << 20) / t / 1e9
print(json.dumps(out))
