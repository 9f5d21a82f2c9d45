"""A2J pose net on 8 crops: one cooperative multi-convolution launch vs one launch per convolution (both replayed
from CUDA graphs, so the host launch path is out of the picture)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from a2j.a2j import A2JModel
from hn_b200 import runtime, synth

sd = synth.a2j_state_dict(seed=1)
x = (torch.rand(8, 1, 176, 176) * 1.5).cuda()
for multi in (True, False, True, False):
    runtime.A2J_MULTI = multi
    m = A2JModel(21, 176, 176).eval(); m.load_state_dict(sd); m.cuda()
    with torch.inference_mode():
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): m.forward_device(x)
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = m.forward_device(x)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"A2J_MULTI={multi}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per forward (8 crops)", flush=True)
