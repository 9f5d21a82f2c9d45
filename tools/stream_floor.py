"""What a small streaming kernel can reach under the timing protocol of tools/membound_roofline.py (L2 flushed, launches
enqueued behind a spin kernel): event-pair overhead, a memset node, and library reductions over 73 / 217 / 333 / 682 MB."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "handnet-pipeline_b200"))
flush = torch.empty(136 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_(); torch.cuda._sleep(3_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2] * 1e3
small = torch.zeros(4608, dtype=torch.int32, device="cuda")
print(f"event pair only            {timeit(lambda: None):8.1f} us")
print(f"memset 18 KB               {timeit(lambda: small.zero_()):8.1f} us")
for mb in (73, 217, 333, 682):
    x = torch.randn(mb * 250_000, device="cuda")
    y = torch.empty_like(x[: x.numel() // 2])
    t = timeit(lambda: x.sum())
    print(f"torch.sum over {mb:4d} MB      {t:8.1f} us  {mb / t * 1e3:7.0f} GB/s")
    t = timeit(lambda: torch.max(x, dim=0))
    print(f"torch.max over {mb:4d} MB      {t:8.1f} us  {mb / t * 1e3:7.0f} GB/s")
    h = x.numel() // 2
    t = timeit(lambda: y.copy_(x[:h]))
    print(f"copy {mb // 2:4d} -> {mb // 2:4d} MB        {t:8.1f} us  {mb / t * 1e3:7.0f} GB/s")
