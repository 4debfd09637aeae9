"""DRAM microbenchmark (torch ops, not product code): read-only, write-only and copy bandwidth."""
import torch, json
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
r = {}
r["copy_gbs"] = 2 * n * 2 / t(lambda: b.copy_(a)) / 1e6
r["write_gbs"] = n * 2 / t(lambda: b.zero_()) / 1e6
r["read_gbs"] = n * 2 / t(lambda: a.view(torch.int16).max()) / 1e6
r["read2_write1_gbs"] = 3 * n * 2 / t(lambda: torch.add(a, a.flip(0) if False else a, out=b)) / 1e6
print(json.dumps(r))
