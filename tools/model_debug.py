"""Model-level bring-up on the GPU box: error statistics against the CPU oracle and first timings."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
sys.path.insert(0, ROOT)
import torch

from a2j.a2j import A2JModel
from fcos_utils.fcos import FCOS
from handnet_pipeline.handnet_pipeline import HandNet
from hn_b200 import ops, synth
from oracle import a2j_oracle, fcos_oracle
from oracle.golden_inputs import inputs_images

torch.set_num_threads(os.cpu_count())


def stats(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    print(f"  {name:22s} max_abs {err.max().item():.4e} mean_abs {err.mean().item():.4e} ref_absmax {ref.abs().max().item():.3f} "
          f"rel_to_max {err.max().item() / max(ref.abs().max().item(), 1e-9):.3e}", flush=True)


def fcos_small():
    print("== FCOS small (256/448 canvas) vs oracle")
    sd = synth.fcos_state_dict(3, False, seed=0)
    m = FCOS(3, ext=False, min_size=256, max_size=448).eval()
    m.load_state_dict(sd)
    m.cuda()
    imgs = inputs_images(5, 2, 120, 160)
    with torch.inference_mode():
        ho = m.head_outputs([i.cuda() for i in imgs])
        torch.cuda.synchronize()
        d_emu, t_emu = fcos_oracle.fcos_forward(sd, imgs, 3, False, 256, 448, emulate_bf16=True, return_taps=True)
        d_f32, t_f32 = fcos_oracle.fcos_forward(sd, imgs, 3, False, 256, 448, emulate_bf16=False, return_taps=True)
    pl = list(m._executor.plans.values())[0]
    stats("canvas vs emu", pl.canvas[..., :3].permute(0, 3, 1, 2), t_emu["canvas"].to(torch.bfloat16))
    for i in range(3):
        stats(f"P{i} vs emu", pl.p[i].to_nchw(), t_emu["p"][i])
        stats(f"P{i} emu vs f32", t_emu["p"][i], t_f32["p"][i])
    for k in ("cls_logits", "bbox_regression", "bbox_ctrness", "hand_lr"):
        stats(f"{k} vs emu", ho[k], t_emu["head"][k])
        stats(f"{k} emu vs f32", t_emu["head"][k], t_f32["head"][k])
    with torch.inference_mode():
        dets = m([i.cuda() for i in imgs])
    print("  kept: gpu", [len(d["boxes"]) for d in dets], "emu", [len(d["boxes"]) for d in d_emu], "f32",
          [len(d["boxes"]) for d in d_f32])
    print("  top box gpu", dets[0]["boxes"][0].tolist(), "emu", d_emu[0]["boxes"][0].tolist())


def a2j_small():
    print("== A2J vs oracle")
    sd = synth.a2j_state_dict(seed=1)
    m = A2JModel(21, 176, 176).eval()
    m.load_state_dict(sd)
    m.cuda()
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 1, 176, 176, generator=g) * 1.5
    with torch.inference_mode():
        cls, reg, dep = m.head_outputs(x.cuda())
        joints = m(x.cuda())
        j_emu, t_emu = a2j_oracle.a2j_forward(sd, x, emulate_bf16=True, return_taps=True)
        j_f32, t_f32 = a2j_oracle.a2j_forward(sd, x, emulate_bf16=False, return_taps=True)
    for k, t in (("cls", cls), ("reg", reg), ("dep", dep)):
        stats(f"{k} vs emu", t, t_emu[k])
        stats(f"{k} emu vs f32", t_emu[k], t_f32[k])
    stats("joints vs emu", joints, j_emu)
    stats("joints vs f32", joints, j_f32)
    stats("joints emu vs f32", j_emu, j_f32)


def timing():
    print("== timing, VGA batch 8")
    class Args:
        pretrained_fcos = ""
        pretrained_a2j = ""
    net = HandNet(Args(), num_classes=3).eval()
    net.detector.load_state_dict(synth.fcos_state_dict(3, False, seed=0))
    net.a2j.load_state_dict(synth.a2j_state_dict(seed=1))
    net.cuda()
    B = 8
    imgs = [i.cuda() for i in inputs_images(11, B, 480, 640)]
    depth = (torch.rand(B, 1, 480, 640) * 1.5).cuda()
    with torch.inference_mode():
        for it in range(3):
            l0 = ops.launch_count()
            torch.cuda.synchronize()
            t = time.perf_counter()
            out = net.forward_device(imgs, depth)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            print(f"  e2e forward_device: {dt * 1e3:.2f} ms for {B} frames -> {B / dt:.1f} frames/s, launches {ops.launch_count() - l0}",
                  flush=True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            out = net.forward_device(imgs, depth)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        print(f"  steady: {ms:.2f} ms per batch of {B} -> {B / ms * 1e3:.1f} frames/s")
        print("  keep_count", out["det"]["keep_count"].tolist(), "cand", out["det"]["cand_count"].tolist(),
              "has_hand", out["has_hand"].tolist())
        print("  crops", out["crops"][:2].tolist())
        # detector only
        ev[0].record()
        for _ in range(5):
            net.detector.forward_device(imgs)
        ev[1].record()
        torch.cuda.synchronize()
        print(f"  detector only: {ev[0].elapsed_time(ev[1]) / 5:.2f} ms per batch of {B}")
        x = out["depth_batch"]
        ev[0].record()
        for _ in range(5):
            net.a2j.forward_device(x)
        ev[1].record()
        torch.cuda.synchronize()
        print(f"  a2j only: {ev[0].elapsed_time(ev[1]) / 5:.2f} ms per batch of {B}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["fcos", "a2j", "timing"]
    if "fcos" in which:
        fcos_small()
    if "a2j" in which:
        a2j_small()
    if "timing" in which:
        timing()
