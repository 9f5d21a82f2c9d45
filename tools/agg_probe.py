"""A2J anchor aggregation at large batch against read-only floors of the same bytes (torch.sum over the three head tensors)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from hn_b200 import ops
from oracle import a2j_oracle

flush = torch.empty(136 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(3_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


anc = a2j_oracle.all_anchors().cuda()
for n in (296, 512, 1024):
    cls = torch.randn(n, 1936, 21, device="cuda"); reg = torch.randn(n, 1936, 21, 2, device="cuda"); dep = torch.randn(n, 1936, 21, device="cuda")
    nbytes = n * 1936 * 21 * 4 * 4
    t = timeit(lambda: ops.a2j_aggregate(cls, reg, dep, anc))
    flat = torch.cat((cls.reshape(-1), reg.reshape(-1), dep.reshape(-1)))
    t_sum = timeit(lambda: flat.sum())
    print(f"n={n:5d}  {nbytes / 1e6:7.1f} MB  aggregate {t * 1e6:7.1f} us = {nbytes / t / 1e9:6.0f} GB/s   torch.sum of the same bytes {t_sum * 1e6:7.1f} us = {nbytes / t_sum / 1e9:6.0f} GB/s", flush=True)
