"""Does a UMMA smem descriptor that starts j rows (j*128 B) into a SWIZZLE_128B tile address the right data, and does
it need base_offset = j?  The A box is loaded j rows early and the descriptor is offset back; rows whose data
falls outside the 128-row box are excluded."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch, torch.nn.functional as F
from hn_b200 import ops
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator().manual_seed(0)
n, h, w, cin, cout = 1, 30, 30, 128, 64
x = (torch.randn(n, cin, h, w, generator=g)).to(torch.bfloat16).float().cuda()
wt = (torch.randn(cout, cin, 3, 3, generator=g) / 30).to(torch.bfloat16).float().cuda()
ref = F.conv2d(x, wt, padding=1)
xin = ops.Act.from_nchw(x, 1)
rows = torch.arange(n * 32 * 32).reshape(n, 32, 32)[:, 1:31, 1:31].cuda()      # flat padded row index of every pixel
for shift in (0, 1, 2, 3, 5):
    for bo in (0, 1):
        out = ops.Act(n, h, w, cout, 1, "cuda")
        ops.conv2d(xin, ops.pack_conv_weight(wt), cout=cout, ksize=3, out=out, block_n=64, debug=shift | (bo << 3))
        torch.cuda.synchronize()
        got = out.to_nchw()
        ok_rows = (rows % 128) < (128 - shift)
        err = (got - ref).abs().amax(dim=1)
        print(f"shift {shift} base_offset_field {bo}: max err on valid rows {err[ok_rows].max().item():.4f}  "
              f"(ref absmax {ref.abs().max().item():.2f}), frac bad {(err[ok_rows] > 0.05).float().mean().item():.3f}")
