"""Is a frame's result independent of the batch it travels in?  HandNet on 8 VGA frames vs the same frames as 2 x 4, 4 x 2
and 8 x 1 (what hn_b200.parallel does over 2 / 4 / 8 ranks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
sys.path.insert(0, ROOT)
import torch
from hn_b200 import synth
from handnet_pipeline.handnet_pipeline import HandNet
from oracle.golden_inputs import inputs_images


class Args:
    pretrained_fcos = ""
    pretrained_a2j = ""


net = HandNet(Args(), num_classes=3).eval()
net.detector.load_state_dict(synth.fcos_state_dict(3, False, seed=0))
net.a2j.load_state_dict(synth.a2j_state_dict(seed=1))
net = net.cuda()
imgs = [i.cuda() for i in inputs_images(131, 8, 480, 640)]
depth = (torch.rand(8, 1, 480, 640, generator=torch.Generator().manual_seed(132)) * 1.5).cuda()
with torch.inference_mode():
    full = net(imgs, depth_images=depth)
    dfull = net.detector(imgs)
    for per in (4, 2, 1):
        j, c, nb, ns = [], [], 0, 0
        for s in range(0, 8, per):
            f, db, cr = net(imgs[s:s + per], depth_images=depth[s:s + per].contiguous())
            j.append(f); c.append(cr.cpu())
            for a, b in zip(net.detector(imgs[s:s + per]), dfull[s:s + per]):
                same = a["boxes"].shape == b["boxes"].shape and torch.equal(a["boxes"], b["boxes"]) and torch.equal(a["scores"], b["scores"])
                nb += int(same)
        j, c = torch.cat(j), torch.cat(c)
        print(f"shards of {per}: joints equal {torch.equal(j, full[0])} (max |d| {(j - full[0]).abs().max().item():.2e})  "
              f"crops equal {torch.equal(c, full[2].cpu())}  detections identical in {nb} of 8 frames")
