import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from a2j.a2j import A2JModel
from hn_b200 import runtime, synth
runtime.A2J_MULTI = os.environ.get("MULTI", "1") == "1"
sd = synth.a2j_state_dict(seed=1)
x = (torch.rand(2, 1, 176, 176) * 1.5).cuda()
m = A2JModel(21, 176, 176).eval(); m.load_state_dict(sd); m.cuda()
with torch.inference_mode():
    out = m.forward_device(x); torch.cuda.synchronize()
print("ok", out.shape, float(out.abs().max()))
