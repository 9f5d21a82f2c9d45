"""Bring-up diagnostics for the tcgen05 conv kernel (run on the GPU box; prints, never asserts)."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
import torch.nn.functional as F

from hn_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"


def q(x):
    return x.to(torch.bfloat16).float()


def run(name, n, h, w, cin, cout, k, halo=1, block_n=0, identity=False, stride=1, dil=1, relu=False):
    try:
        g = torch.Generator(device="cpu").manual_seed(1)
        x = q(torch.randn(n, cin, h, w, generator=g)).to(dev)
        if identity:
            wt = torch.zeros(cout, cin, k, k)
            for i in range(min(cin, cout)):
                wt[i, i, k // 2, k // 2] = 1.0
        else:
            wt = q(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
        wt = wt.to(dev)
        ref = F.conv2d(x, wt, None, stride=stride, padding=(k // 2) * dil, dilation=dil)
        if relu:
            ref = F.relu(ref)
        if stride == 1:
            xin = ops.Act.from_nchw(x, halo)
        else:
            xin = ops.PhaseAct.from_nchw(x, halo)
        oh, ow = ref.shape[-2:]
        out = ops.Act(n, oh, ow, cout, halo, dev)
        ops.conv2d(xin, ops.pack_conv_weight(wt), cout=cout, ksize=k, stride=stride, dilation=dil, relu=relu, out=out,
                   block_n=block_n)
        torch.cuda.synchronize()
        got = out.to_nchw()
        err = (got - ref).abs()
        halo_sum = out.t.float().abs().sum().item() - out.interior().float().abs().sum().item()
        print(f"{name:40s} max_err {err.max().item():.4e} mean_err {err.mean().item():.4e} ref_absmax "
              f"{ref.abs().max().item():.3f} halo_abs_sum {halo_sum:.3e}", flush=True)
        if err.max().item() > 0.05 * max(1.0, ref.abs().max().item()):
            bad = (err > 0.05).float()
            print("   bad fraction per channel (first 16):", bad.mean(dim=(0, 2, 3))[:16].tolist())
            print("   bad fraction per row (first 8):", bad.mean(dim=(0, 1, 3))[:8].tolist())
            print("   got[0,:8,0,0]", got[0, :8, 0, 0].tolist())
            print("   ref[0,:8,0,0]", ref[0, :8, 0, 0].tolist())
            print("   got[0,:8,1,1]", got[0, :8, 1, 1].tolist())
            print("   ref[0,:8,1,1]", ref[0, :8, 1, 1].tolist())
    except Exception:
        print(f"{name}: EXCEPTION")
        traceback.print_exc()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), ops.device_info())
    run("identity 1x1 64->64 bn64", 1, 8, 14, 64, 64, 1, halo=1, block_n=64, identity=True)
    run("rand 1x1 64->64 bn64", 1, 8, 14, 64, 64, 1, halo=1, block_n=64)
    run("rand 1x1 128->64 bn64 (2 k-blocks)", 1, 8, 14, 128, 64, 1, halo=1, block_n=64)
    run("identity 3x3 64->64 bn64", 1, 8, 14, 64, 64, 3, halo=1, block_n=64, identity=True)
    run("rand 3x3 64->64 bn64", 2, 20, 28, 64, 64, 3, halo=1, block_n=64)
    run("rand 3x3 64->128 bn128", 2, 20, 28, 64, 128, 3, halo=1, block_n=128)
    run("rand 3x3 256->256 bn256", 2, 20, 28, 256, 256, 3, halo=1, block_n=256, relu=True)
    run("rand 3x3 64->32 bn32", 2, 20, 28, 64, 32, 3, halo=1, block_n=32)
    run("rand 3x3 64->16 bn16", 2, 20, 28, 64, 16, 3, halo=1, block_n=16)
    run("rand 3x3 s2 64->128", 2, 20, 28, 64, 128, 3, halo=1, stride=2)
    run("rand 1x1 s2 64->128", 2, 21, 27, 64, 128, 1, halo=1, stride=2)
    run("rand 3x3 dil2 128->128", 2, 11, 11, 128, 128, 3, halo=2, dil=2)
    run("rand 3x3 256->256 auto big", 4, 100, 136, 256, 256, 3, halo=1)
    print("launches", ops.launch_count())
