"""Per-group durations of the A2J multi-convolution launch (clock64 of CTA 0 after every grid barrier)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from a2j.a2j import A2JModel
from hn_b200 import runtime, synth, ops, _lib
runtime.A2J_MULTI = True

sd = synth.a2j_state_dict(seed=1)
x = (torch.rand(8, 1, 176, 176) * 1.5).cuda()
m = A2JModel(21, 176, 176).eval(); m.load_state_dict(sd); m.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.inference_mode():
    for _ in range(3): m.forward_device(x)
    torch.cuda.synchronize()
    ex = m._exec if hasattr(m, "_exec") else None
    buf = torch.zeros(256, dtype=torch.int64, device="cuda")
    _lib.load().hn_conv_multi_set_trace(buf.data_ptr())
    for cold in (False, True):
        if cold: flush.zero_()
        m.forward_device(x); torch.cuda.synchronize()
        t = buf.cpu().tolist()
        t = [v for v in t if v > 0]
        d = [b - a for a, b in zip(t[:-1], t[1:])]
        print(("cold L2" if cold else "hot L2"), "groups:", len(t), "total cycles", t[-1] - t[0], "mean/group", sum(d) / len(d))
        print(" ".join(str(v) for v in d))
    _lib.load().hn_conv_multi_set_trace(None)
