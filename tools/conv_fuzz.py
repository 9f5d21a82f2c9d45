"""Randomised conv shapes through hn_conv2d_bf16 against F.conv2d (bf16-rounded inputs, fp32 accumulate):
python tools/conv_fuzz.py [n_cases] [seed]"""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch, torch.nn.functional as F
from hn_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for i in range(cases):
    k = rng.choice([1, 3, 3])
    stride = rng.choice([1, 1, 1, 2])
    dil = rng.choice([1, 1, 2]) if (k == 3 and stride == 1) else 1
    cin = 64 * rng.choice([1, 1, 2, 3, 4, 8])
    cout = rng.choice([5, 13, 21, 48, 64, 96, 128, 192, 256, 336, 512])
    n = rng.choice([1, 2, 3, 8])
    h, w = rng.randint(3, 70), rng.randint(3, 90)
    if rng.random() < 0.15:
        h, w = rng.randint(100, 210), rng.randint(100, 280)
        cin, cout = rng.choice([64, 128]), rng.choice([5, 64, 128])
    halo = max(1, dil) if k == 3 else rng.choice([0, 1])
    if stride == 2:
        halo = 1
    bn = rng.choice([0, 0, 0, 16, 32, 64, 128, 256])
    cout_pad = ((cout + 15) // 16) * 16
    if bn and cout_pad % bn:
        bn = 0
    use_res = rng.random() < 0.3 and stride == 1
    relu = rng.random() < 0.7
    g = torch.Generator().manual_seed(i)
    x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16).float().cuda()
    wt = (torch.randn(cout, cin, k, k, generator=g) * (cin * k * k) ** -0.5).to(torch.bfloat16).float().cuda()
    scale = (0.5 + torch.rand(cout, generator=g)).cuda() if rng.random() < 0.6 else None
    shift = (0.3 * torch.randn(cout, generator=g)).cuda()
    ref = F.conv2d(x, wt, None, stride=stride, padding=(k // 2) * dil, dilation=dil)
    if scale is not None:
        ref = ref * scale[None, :, None, None]
    ref = ref + shift[None, :, None, None]
    oh, ow = ref.shape[2], ref.shape[3]
    idn = torch.randn(n, cout, oh, ow, generator=g).to(torch.bfloat16).float().cuda() if use_res else None
    if use_res:
        ref = ref + idn
    if relu:
        ref = F.relu(ref)
    xin = ops.Act.from_nchw(x, halo) if stride == 1 else ops.PhaseAct.from_nchw(x, halo)
    out = ops.Act(n, oh, ow, cout, rng.choice([0, 1, 2]), "cuda")
    res = ops.Act.from_nchw(idn, rng.choice([0, 1])) if use_res else None
    desc = f"#{i} n{n} {h}x{w} {cin}->{cout} k{k} s{stride} d{dil} halo{halo} bn{bn} res{int(use_res)} relu{int(relu)} scale{int(scale is not None)}"
    try:
        ops.conv2d(xin, ops.pack_conv_weight(wt), cout=cout, ksize=k, stride=stride, dilation=dil, scale=scale, shift=shift,
                   relu=relu, res=res, res_mode=1 if use_res else 0, out=out, block_n=bn)
        torch.cuda.synchronize()
    except RuntimeError as e:
        print("ERR ", desc, str(e)[:160]); bad += 1; continue
    got = out.to_nchw()
    tol = ref.abs() * 2 ** -7 + 2e-2
    nbad = int(((got - ref).abs() > tol).sum())
    halo_ok = bool(torch.isclose(out.t.float().abs().sum(), out.interior().float().abs().sum()))
    if nbad or not halo_ok:
        bad += 1
        print("FAIL", desc, f"{nbad} of {ref.numel()} off, max err {float((got - ref).abs().max()):.3f}, halo_ok {halo_ok}")
print(f"{cases - bad} of {cases} random conv cases ok")
