import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from a2j.a2j import A2JModel
from hn_b200 import runtime, synth
sd = synth.a2j_state_dict(seed=1)
x = (torch.rand(int(os.environ.get("N", "2")), 1, 176, 176) * 1.5).cuda()
outs = {}
for multi in (False, True):
    runtime.A2J_MULTI = multi
    m = A2JModel(21, 176, 176).eval(); m.load_state_dict(sd); m.cuda()
    with torch.inference_mode():
        for rep in range(3):
            out = m.forward_device(x); torch.cuda.synchronize()
    outs[multi] = out.clone()
    print("multi", multi, "ok", out.shape, float(out.abs().max()), flush=True)
print("max diff multi vs per-layer:", float((outs[True] - outs[False]).abs().max()))
