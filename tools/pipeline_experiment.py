"""Throughput with P steps in flight (one CUDA graph + buffer set + stream each) vs one step at a time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
import bench
from hn_b200.runtime import GraphedHandNet

dev = torch.device("cuda", 0)
net = bench.build_net(dev)
B = bench.FRAMES_PER_GPU
rgb, depth = bench.synthetic_frames(1000, B)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
K = 60
with torch.inference_mode():
    ref = None
    for P in (1, 2, 3):
        steps = [GraphedHandNet(net, B, bench.IMG_H, bench.IMG_W, slot=i) for i in range(P)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(P)]
        for st, s in zip(steps, streams):
            st.load_inputs(rgb.pin_memory(), depth.pin_memory())
            with torch.cuda.stream(s):
                for _ in range(3): st.run()
        torch.cuda.synchronize()
        if ref is None: ref = steps[0].records().clone()
        for st in steps: assert torch.equal(st.records(), ref), "slots disagree"
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for s in streams: s.wait_event(e0)
        for i in range(K):
            with torch.cuda.stream(streams[i % P]):
                flush.zero_()
                steps[i % P].run()
        for s in streams: main.wait_stream(s)
        e1.record(main)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        for st in steps: assert torch.equal(st.records(), ref), "results changed under concurrency"
        print(f"P={P}: {ms / K:.3f} ms/step  {B * K / ms * 1e3:.0f} frames/s", flush=True)
