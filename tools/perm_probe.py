"""Where does a frame's result depend on its position in the batch?  Runs the detector on 8 VGA frames and on the reversed
batch and compares the canvas, the pyramid levels (no GroupNorm before them) and the head outputs frame by frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
sys.path.insert(0, ROOT)
import torch
from hn_b200 import synth
from fcos_utils.fcos import FCOS
from oracle.golden_inputs import inputs_images

m = FCOS(3, ext=False).eval()
m.load_state_dict(synth.fcos_state_dict(3, False, seed=0))
m = m.cuda()
imgs = [i.cuda() for i in inputs_images(131, 8, 480, 640)]


def run(lst):
    with torch.inference_mode():
        ho = {k: v.clone() for k, v in m.head_outputs(lst).items()}
    pl = list(m._executor.plans.values())[0]
    return ho, pl.frame.canvas().clone(), [p.to_nchw().clone() for p in pl.p]


ho_a, cv_a, p_a = run(imgs)
ho_b, cv_b, p_b = run(imgs[::-1])
print("canvas equal:", torch.equal(cv_a, cv_b.flip(0)))
for i in range(len(p_a)):
    d = (p_a[i].float() - p_b[i].flip(0).float()).abs()
    print(f"P{i + 3}: equal {torch.equal(p_a[i], p_b[i].flip(0))}  max |d| {d.max().item():.3e}  differing {int((d > 0).sum())} of {d.numel()}")
for k in ho_a:
    d = (ho_a[k].float() - ho_b[k].flip(0).float()).abs()
    per = d.reshape(8, -1).amax(dim=1).tolist()
    print(f"{k}: max |d| {d.max().item():.3e}  differing {int((d > 0).sum())} of {d.numel()}  per frame {[f'{x:.1e}' for x in per]}")
