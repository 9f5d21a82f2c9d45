"""Generate tests/golden/*.pt by running the UNMODIFIED reference on the CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Weights come from hn_b200.synth (seeded, same keys/shapes as the reference modules; loaded
with strict=True), inputs from seeded torch generators, so a fixture stores only seeds, the
configuration and the reference's outputs (or strided samples of large tensors).
tests/test_oracle_golden.py replays oracle/ against these files; the GPU parity tests then
compare the CUDA path with the oracle.
"""
from __future__ import annotations

import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
from hn_b200 import synth  # noqa: E402  (pure-python weight tables, no CUDA involved)

from oracle import _refshim  # noqa: E402
from oracle.golden_inputs import inputs_images, pad_crop_inputs, sample, stress_head_tensors  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    warnings.filterwarnings("ignore")
    _refshim.install()
    import numpy as np
    import torch
    import torchvision
    from a2j.a2j import A2JModel, convert_joints
    from fcos_utils.fcos import FCOS
    from handnet_pipeline.handnet_pipeline import HandNet

    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    meta = {"torch": torch.__version__, "torchvision": torchvision.__version__}

    # ------------------------------------------------------------------ FCOS, small canvas
    cfg = dict(num_classes=3, ext=False, min_size=256, max_size=448, seed_w=0, seed_x=5, n=2, h=120, w=160)
    fsd = synth.fcos_state_dict(cfg["num_classes"], cfg["ext"], seed=cfg["seed_w"])
    m = FCOS(num_classes=3, ext=False, nms_thresh=0.5, min_size=256, max_size=448).eval()
    m.load_state_dict(fsd, strict=True)
    imgs = inputs_images(cfg["seed_x"], cfg["n"], cfg["h"], cfg["w"])
    with torch.inference_mode():
        il, _ = m.transform(imgs, None)
        feats = m.backbone(il.tensors)
        plist = [feats["0"], feats["1"], feats["2"]]
        ho = m.head(plist)
        dets = m(imgs)
    fx = {"cfg": cfg, "meta": meta, "canvas_shape": torch.tensor(il.tensors.shape),
          "image_sizes": torch.tensor(il.image_sizes), "dets": dets}
    fx["canvas_s"], fx["canvas_n"] = sample(il.tensors)
    for i, p in enumerate(plist):
        fx[f"p{i}_s"], fx[f"p{i}_n"] = sample(p)
    for k in ("cls_logits", "bbox_regression", "bbox_ctrness", "hand_lr"):
        fx[f"head_{k}_s"], fx[f"head_{k}_n"] = sample(ho[k])
    torch.save(fx, os.path.join(OUT, "fcos_small.pt"))
    print("fcos_small: kept", [len(d["boxes"]) for d in dets])

    # ------------------------------------------------------------------ FCOS ext=True heads
    cfg = dict(num_classes=2, ext=True, min_size=192, max_size=320, seed_w=3, seed_x=6, n=1, h=96, w=128)
    esd = synth.fcos_state_dict(2, True, seed=3)
    me = FCOS(num_classes=2, ext=True, nms_thresh=0.5, min_size=192, max_size=320).eval()
    me.load_state_dict(esd, strict=True)
    imgs = inputs_images(cfg["seed_x"], cfg["n"], cfg["h"], cfg["w"])
    with torch.inference_mode():
        dets = me(imgs)
    torch.save({"cfg": cfg, "meta": meta, "dets": dets}, os.path.join(OUT, "fcos_ext_small.pt"))
    print("fcos_ext_small: kept", [len(d["boxes"]) for d in dets])

    # ------------------------------------------------------------------ A2J
    cfg = dict(seed_w=1, seed_x=7, n=2)
    asd = synth.a2j_state_dict(seed=1)
    a = A2JModel(21, 176, 176).eval()
    a.load_state_dict(asd, strict=True)
    g = torch.Generator().manual_seed(cfg["seed_x"])
    x = torch.rand(cfg["n"], 1, 176, 176, generator=g) * 1.5
    with torch.inference_mode():
        c4, c5 = a.Backbone(x)
        cls, reg, dep = a.classificationModel(c4), a.regressionModel(c5), a.DepthRegressionModel(c5)
        joints = a(x)
    fx = {"cfg": cfg, "meta": meta, "joints": joints}
    for k, t in (("c4", c4), ("c5", c5), ("cls", cls), ("reg", reg), ("dep", dep)):
        fx[k + "_s"], fx[k + "_n"] = sample(t)
    box = np.array([100.0, 80.0, 260.0, 250.0], dtype=np.float32)
    paras = np.array([613.0, 614.0, 321.5, 239.1], dtype=np.float32)
    fx["convert_uv"] = torch.from_numpy(convert_joints(joints[0].numpy(), None, box, None, 176, 176))
    fx["convert_xyz"] = torch.from_numpy(convert_joints(joints[0].numpy(), None, box, paras, 176, 176))
    fx["convert_box"], fx["convert_paras"] = torch.from_numpy(box), torch.from_numpy(paras)
    torch.save(fx, os.path.join(OUT, "a2j_small.pt"))
    print("a2j_small: joints[0,:2]", joints[0, :2].tolist())

    # ------------------------------------------------------------------ HandNet end to end (VGA)
    cfg = dict(num_classes=3, seed_fcos=0, seed_a2j=1, seed_x=11, n=2, h=480, w=640)

    class Args:
        pretrained_fcos = ""
        pretrained_a2j = ""

    net = HandNet(Args(), reload_detector=False, num_classes=3, reload_a2j=False).eval()
    net.detector.load_state_dict(fsd, strict=True)
    net.a2j.load_state_dict(asd, strict=True)
    imgs = inputs_images(cfg["seed_x"], cfg["n"], cfg["h"], cfg["w"])
    g = torch.Generator().manual_seed(cfg["seed_x"] + 1)
    depth = torch.rand(cfg["n"], 1, cfg["h"], cfg["w"], generator=g) * 1.5
    with torch.inference_mode():
        dets = net.detector(imgs, None)
        final, depth_batch, crops = net(imgs, depth_images=depth)
    fx = {"cfg": cfg, "meta": meta, "final": final, "crops": crops,
          "n_kept": torch.tensor([len(d["boxes"]) for d in dets]),
          "top_boxes": torch.stack([d["boxes"][:8] for d in dets]),
          "top_scores": torch.stack([d["scores"][:8] for d in dets]),
          "top_labels": torch.stack([d["labels"][:8] for d in dets])}
    fx["depth_batch_s"], fx["depth_batch_n"] = sample(depth_batch)
    torch.save(fx, os.path.join(OUT, "handnet_vga.pt"))
    print("handnet_vga: kept", fx["n_kept"].tolist(), "crops", crops.tolist())

    # ------------------------------------------------------------------ S1/S2 pad + crop cases
    class FakeDet(torch.nn.Module):
        def __init__(self, boxes):
            super().__init__()
            self.boxes = boxes

        def forward(self, images, targets=None):
            return [{"boxes": b[None], "labels": torch.tensor([2])} for b in self.boxes]

    class FakeA2J(torch.nn.Module):
        def forward(self, xb):
            return torch.zeros(xb.shape[0], 21, 3)

    hh, ww, nb = 97, 131, 48
    boxes, depth = pad_crop_inputs(21, nb, hh, ww)
    net.detector, net.a2j = FakeDet(boxes), FakeA2J()
    imgs = [torch.zeros(3, hh, ww) for _ in range(nb)]
    with torch.inference_mode():
        _, depth_batch, crops = net(imgs, depth_images=depth)
    torch.save({"meta": meta, "hw": torch.tensor([hh, ww]), "seed": 21, "nb": nb, "crops": crops,
                "depth_batch_s4": depth_batch[:, :, ::4, ::4].clone(),
                "depth_batch_sum": depth_batch.double().sum(dim=(1, 2, 3))},
               os.path.join(OUT, "pad_crop_cases.pt"))
    print("pad_crop_cases:", crops[:4].tolist())

    # ------------------------------------------------------------------ post-process stress (config 4)
    cfg = dict(seed=31, batch=2, canvas=(800, 1088), grids=[(100, 136), (50, 68), (25, 34)], mu=-0.35)
    locs = sum(gh * gw for gh, gw in cfg["grids"])
    ho = stress_head_tensors(cfg["seed"], cfg["batch"], locs, 3, cfg["mu"])
    det = FCOS(num_classes=3, ext=False).eval()

    class IL:
        pass

    il = IL()
    il.tensors = torch.zeros(cfg["batch"], 3, *cfg["canvas"])
    il.image_sizes = [(800, 1066)] * cfg["batch"]
    featmaps = [torch.zeros(cfg["batch"], 1, gh, gw) for gh, gw in cfg["grids"]]
    with torch.inference_mode():
        anchors = det.anchor_generator(il, featmaps)
        dd = det.postprocess_detections(dict(ho, feature_idx=None), anchors, [gh * gw for gh, gw in cfg["grids"]])
        dd = det.postprocess(dd, il.image_sizes, [(480, 640)] * cfg["batch"])
        sc = torch.sqrt(torch.sigmoid(ho["cls_logits"]) * torch.sigmoid(ho["bbox_ctrness"])).max(-1)[0]
    fx = {"cfg": cfg, "meta": meta, "n_candidates": (sc > 0.7).sum(1), "dets": dd,
          "anchors_first_last": torch.stack((anchors[0][0], anchors[0][-1]))}
    torch.save(fx, os.path.join(OUT, "postprocess_stress.pt"))
    print("postprocess_stress: candidates", fx["n_candidates"].tolist(), "kept", [len(d["boxes"]) for d in dd])

    # ------------------------------------------------------------------ NMS known-answer cases
    from torchvision.ops import batched_nms, nms
    cases = []
    g = torch.Generator().manual_seed(41)
    for n, ncls, spread in ((0, 1, 50), (1, 1, 50), (7, 2, 20), (64, 3, 60), (65, 1, 30), (300, 3, 120),
                            (999, 3, 300), (1000, 3, 300), (1001, 3, 300), (2500, 4, 400)):
        xy = torch.rand(n, 2, generator=g) * spread - 5.0
        wh = torch.rand(n, 2, generator=g) * 40 + 1
        b = torch.cat((xy, xy + wh), dim=1)
        s = torch.rand(n, generator=g)
        if n >= 64:
            s[n // 2: n // 2 + 8] = s[3]                         # score ties
        lab = torch.randint(0, ncls, (n,), generator=g)
        cases.append({"boxes": b, "scores": s, "labels": lab, "thr": 0.3,
                      "keep_nms": nms(b, s, 0.3), "keep_batched": batched_nms(b, s, lab, 0.3)})
    # exact-threshold IoU (3/10 in float32 is 0.30000001 > 0.3) and integer grids with many equal IoUs
    b = torch.tensor([[0, 0, 10, 1], [7, 0, 10, 1]], dtype=torch.float32)
    s = torch.tensor([0.9, 0.8])
    lab = torch.zeros(2, dtype=torch.int64)
    cases.append({"boxes": b, "scores": s, "labels": lab, "thr": 0.3, "keep_nms": nms(b, s, 0.3),
                  "keep_batched": batched_nms(b, s, lab, 0.3)})
    xy = torch.randint(0, 30, (400, 2), generator=g).float()
    wh = torch.randint(1, 12, (400, 2), generator=g).float()
    b = torch.cat((xy, xy + wh), dim=1)
    s = torch.randint(0, 50, (400,), generator=g).float() / 50
    lab = torch.randint(0, 3, (400,), generator=g)
    cases.append({"boxes": b, "scores": s, "labels": lab, "thr": 0.3, "keep_nms": nms(b, s, 0.3),
                  "keep_batched": batched_nms(b, s, lab, 0.3)})
    torch.save({"meta": meta, "cases": cases}, os.path.join(OUT, "nms_cases.pt"))
    print("nms_cases:", [(len(c["scores"]), len(c["keep_batched"])) for c in cases])


if __name__ == "__main__":
    main()
