"""Device time of the phases of one eager step (CUDA events between phases, the step queued behind a spin kernel so
that the host is out of the picture; graph branches run as real side streams)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
import bench
from hn_b200 import runtime
from hn_b200.runtime import GraphedHandNet

dev = torch.device("cuda", 0)
net = bench.build_net(dev)
rgb, depth = bench.synthetic_frames(1000, bench.FRAMES_PER_GPU)
step = GraphedHandNet(net, bench.FRAMES_PER_GPU, bench.IMG_H, bench.IMG_W, use_graph=False)
step.load_inputs(rgb.pin_memory(), depth.pin_memory())
flush = torch.empty(136 << 20, dtype=torch.uint8, device=dev)
with torch.inference_mode():
    for _ in range(3): step._eager()
    torch.cuda.synchronize()
    best = None
    for rep in range(5):
        flush.zero_()
        runtime.PHASES = []
        torch.cuda._sleep(200_000_000)
        runtime.mark("start")
        step._eager()
        torch.cuda.synchronize()
        ph, runtime.PHASES = runtime.PHASES, None
        t = [(b[0], a[1].elapsed_time(b[1]) * 1e3) for a, b in zip(ph[:-1], ph[1:])]
        tot = sum(v for _, v in t)
        if best is None or tot < best[0]:
            best = (tot, t)
    print(f"total {best[0]:.0f} us")
    for name, us in best[1]:
        print(f"  {name:16s} {us:8.1f} us  {100 * us / best[0]:5.1f}%")
