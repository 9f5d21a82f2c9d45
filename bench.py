#!/usr/bin/env python
"""Benchmark of the detect -> crop -> A2J-pose path (BASELINE.json metric: E2E frames/s, 640x480).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = the whole HandNet path over one batch of FRAMES_PER_GPU synthetic 640x480 RGB + depth frames per
GPU (BASELINE.json configs[2]: batch 64 over 8 GPUs = 8 frames per GPU; weak scaling).  Random-init weights of
the reference architectures (hn_b200.synth), head biases set so that a realistic handful of boxes passes the
hard-coded 0.7 cut and every frame yields a hand crop (candidate / kept counts are reported in `config`).

  value      frames/s, inputs resident in HBM, the step replayed as one CUDA graph, timed with CUDA events
  e2e        frames/s through HandNet's public API path with HOST (pinned) inputs: H2D of the frames and depth
             maps and D2H of joints / crops / hit mask inside the timed region
  roofline   the dominant kernel (tcgen05 shifted-GEMM conv): algorithmic conv FLOPs of one step / the summed
             device time of its launches in one step (CUDA events on the launch stream, GPU kept saturated)
  cpu_baseline  the oracle (a torch-CPU restatement of the reference, oracle/) timed on this box's host cores

`--impl reference` times that CPU restatement alone (the reference itself is Python that needs packages which
are not installed on the box) on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FRAMES_PER_GPU = 8
IMG_H, IMG_W = 480, 640
CLS_BIAS = [-6.0, -6.0, -1.5]          # ~50 candidates, ~30 kept boxes per frame
METRIC = "e2e_frames_per_s_640x480_detect_plus_a2j_pose"
UNIT = "frames/s"
# algorithmic conv FLOPs (2*MAC over the padded canvas), SURVEY.md 8d / BASELINE.md section 3
FLOP_PER_FRAME = 318.99e9 + 12.55e9


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained", 1400.0), "bf16_burst": d.get("bf16_tflops", 1590.0),
                "hbm": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, window=None):
        """Summary of the samples taken inside `window` = (t0, t1) wall-clock seconds (all samples if None)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if window is None or (window[0] <= t <= window[1] + 0.06)]
        sm = sorted(int(float(r[0])) for r in rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # median over the upper half of the samples = clocks under load (idle samples sit far below)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_net(device):
    from handnet_pipeline.handnet_pipeline import HandNet
    from hn_b200 import synth

    class Args:
        pretrained_fcos = ""
        pretrained_a2j = ""

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        net = HandNet(Args(), reload_detector=False, num_classes=3, reload_a2j=False, RGBD=False).eval()
    net.detector.load_state_dict(synth.fcos_state_dict(3, False, seed=0, cls_bias=CLS_BIAS))
    net.a2j.load_state_dict(synth.a2j_state_dict(seed=1))
    return net.to(device)


def synthetic_frames(seed: int, n: int):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(n, 3, IMG_H, IMG_W, generator=g)
    depth = torch.rand(n, 1, IMG_H, IMG_W, generator=g) * 1.5
    return rgb, depth


# --------------------------------------------------------------------------------------------------
def cpu_oracle_frames_per_s(frames: int, repeats: int, threads: int):
    """The reference's CPU path, restated (oracle/), on `frames` synthetic VGA frames per repeat."""
    from hn_b200 import synth
    from oracle import handnet_oracle
    torch.set_num_threads(threads)
    fsd = synth.fcos_state_dict(3, False, seed=0, cls_bias=CLS_BIAS)
    asd = synth.a2j_state_dict(seed=1)
    rgb, depth = synthetic_frames(100, frames)
    imgs = list(rgb)
    times = []
    with torch.inference_mode():
        for r in range(repeats + 1):
            t = time.perf_counter()
            handnet_oracle.handnet_forward(fsd, asd, imgs, depth, num_classes=3)
            dt = time.perf_counter() - t
            if r > 0:                      # first pass is the warm-up
                times.append(dt)
    times.sort()
    return frames / times[len(times) // 2]


def run_reference(args):
    """--impl reference: CPU restatement of the reference path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from hn_b200 import synth
    from oracle import handnet_oracle
    torch.set_num_threads(threads)
    fsd = synth.fcos_state_dict(3, False, seed=0, cls_bias=CLS_BIAS)
    asd = synth.a2j_state_dict(seed=1)
    sample = 2                                           # frames per step (bounded sample of the 8-frame batch)
    rgb, depth = synthetic_frames(100, sample)
    imgs = list(rgb)
    with torch.inference_mode():
        for _ in range(args.warmup):
            handnet_oracle.handnet_forward(fsd, asd, imgs, depth, num_classes=3)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            handnet_oracle.handnet_forward(fsd, asd, imgs, depth, num_classes=3)
        dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"e2e_handnet_{IMG_W}x{IMG_H}_b{FRAMES_PER_GPU}_per_gpu", "frames_per_step": sample,
                   "sample": f"{sample} of the {FRAMES_PER_GPU} frames of a step", "canvas": "800x1088",
                   "note": "reference is Python with uninstalled deps on the box; timed its CPU restatement (oracle/)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} VGA frames per step, {args.steps} steps"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"

    from hn_b200 import ops
    from hn_b200.runtime import GraphedHandNet, conv_flops_per_step, conv_profile

    net = build_net(dev)
    B = FRAMES_PER_GPU
    rgb_h, depth_h = synthetic_frames(1000 + rank, B)
    rgb_pin, depth_pin = rgb_h.pin_memory(), depth_h.pin_memory()
    step = GraphedHandNet(net, B, IMG_H, IMG_W, use_graph=not args.no_graph)
    step.load_inputs(rgb_pin, depth_pin)
    from hn_b200 import parallel
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def one_step():
        out = step.run()
        if world > 1:     # per-frame records to every rank (rank 0 consumes them): the path's only collective
            parallel.gather_records(step.records(), B)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.inference_mode():
        sampler = ClockSampler(local_rank)            # started early: nvidia-smi needs a moment before its first sample
        sampler.start()
        for _ in range(args.warmup):
            one_step()
        barrier()
        # ---------------- device-resident throughput (`value`) ----------------
        launches0 = ops.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        t_wall0 = time.time()
        torch.cuda.nvtx.range_push("hn_timed")
        for s, e in evs:
            l2_flush.zero_()                 # flush L2 between timed iterations (outside the events)
            s.record()
            one_step()
            e.record()
        barrier()
        torch.cuda.nvtx.range_pop()
        t_wall1 = time.time()
        launches = (ops.launch_count() - launches0) // args.steps if args.no_graph else step.launches_per_step
        dev_ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        value = world * B * args.steps / (total_ms * 1e-3)

        # ---------------- end to end through the public API with host buffers (`e2e`) ----------------
        # what a caller of the reference does (ros_demo.py:266-273): host frames -> .cuda() -> HandNet.forward ->
        # joints on the host.  H2D of rgb + depth and the D2H read-back of the result records are inside.
        # The caller double-buffers its uploads: the H2D copy of step i+1 is issued on a copy stream before step i is
        # submitted, so the PCIe transfer overlaps the kernels.  Every step still uploads its own frames and reads
        # its own results back.
        copy_stream = torch.cuda.Stream(device=dev)

        def upload():
            with torch.cuda.stream(copy_stream):
                imgs = rgb_pin.to(dev, non_blocking=True)
                dpt = depth_pin.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return imgs, dpt, ev

        def api_step(cur):
            imgs, dpt, ev = cur
            main = torch.cuda.current_stream()
            main.wait_event(ev)
            imgs.record_stream(main)
            dpt.record_stream(main)
            nxt = upload()
            final, depth_batch, crops = net(list(imgs.unbind(0)), depth_images=dpt)
            if world > 1:
                parallel.gather_records(net._steps[next(iter(net._steps))].records(), B)
            return final, nxt

        cur = upload()
        for _ in range(3):
            res, cur = api_step(cur)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res, cur = api_step(cur)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3          # host wall clock: the call returns host results
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * B * args.steps / (float(t.item()) * 1e-3)
        h2d = rgb_pin.numel() * 4 + depth_pin.numel() * 4
        d2h = step.d2h_bytes

        # ---------------- the same with the frames as the camera delivers them ----------------
        # uint8 BGR + uint16 millimetres on the host -> HandNet.forward_frames: H2D of 5 bytes per pixel, the
        # conversion ros_demo.py does with numpy on the host (x/255, BGR->RGB, mm/1000) runs on the device.
        bgr_pin = (rgb_h.flip(1).permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().pin_memory()
        mm_pin = (depth_h[:, 0] * 1000).round().clamp(0, 32767).to(torch.int16).contiguous().pin_memory()
        def upload_u8():                       # same double-buffered upload as the fp32 loop above
            with torch.cuda.stream(copy_stream):
                b8 = bgr_pin.to(dev, non_blocking=True)
                m16 = mm_pin.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return b8, m16, ev

        def api_step_u8(cur):
            b8, m16, ev = cur
            main = torch.cuda.current_stream()
            main.wait_event(ev)
            b8.record_stream(main)
            m16.record_stream(main)
            nxt = upload_u8()
            net.forward_frames(b8, m16)
            return nxt

        cur8 = upload_u8()
        for _ in range(3):
            cur8 = api_step_u8(cur8)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cur8 = api_step_u8(cur8)
        barrier()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_u8_value = world * B * args.steps / (float(t.item()) * 1e-3)
        h2d_u8 = bgr_pin.numel() + mm_pin.numel() * 2
        t_wall2 = time.time()
        # clocks / throttle reasons sampled DURING the device-timed region; if that region was too short for two
        # samples (it lasts steps x ~4 ms), the window is extended over the e2e loops that follow it (same load)
        clocks = sampler.stop((t_wall0, t_wall1))
        clocks["window"] = "timed region"
        if (clocks.get("samples") or 0) < 2:
            clocks = sampler.stop((t_wall0, t_wall2))
            clocks["window"] = "timed region + e2e loops"

        # ---------------- several steps in flight (extra; single GPU only) ----------------
        # Three independent batches at a time, each with its own CUDA graph, buffer set (runtime.PLAN_SLOT) and stream:
        # the latency-bound tail of one step (the A2J pose net keeps < 1/3 of the SMs busy) overlaps the detector of the
        # next.  Same work per step, same results (checked); `value` above stays the one-step-at-a-time number.
        pipelined = None
        if world == 1 and not args.no_graph:
            P = 3
            psteps = [step] + [GraphedHandNet(net, B, IMG_H, IMG_W, slot=i) for i in range(1, P)]
            pstreams = [torch.cuda.Stream(device=dev) for _ in range(P)]
            for st_, s_ in zip(psteps, pstreams):
                st_.load_inputs(rgb_pin, depth_pin)
                with torch.cuda.stream(s_):
                    for _ in range(3):
                        st_.run()
            torch.cuda.synchronize()
            same = all(torch.equal(st_.records(), step.records()) for st_ in psteps)
            main_s = torch.cuda.current_stream()
            pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pe0.record(main_s)
            for s_ in pstreams:
                s_.wait_event(pe0)
            for i in range(args.steps):
                with torch.cuda.stream(pstreams[i % P]):
                    l2_flush.zero_()
                    psteps[i % P].run()
            for s_ in pstreams:
                main_s.wait_stream(s_)
            pe1.record(main_s)
            torch.cuda.synchronize()
            pms = pe0.elapsed_time(pe1)
            pipelined = {"steps_in_flight": P, "value": B * args.steps / (pms * 1e-3), "unit": UNIT,
                         "ms_per_step": pms / args.steps, "results_identical": bool(same)}

        # ---------------- roofline of the dominant kernel (rank 0) ----------------
        roof = None
        counts = step.counts()
        if rank == 0:
            peaks = load_peaks()
            conv_ms, n_conv = conv_profile(step, repeats=3)
            flops = conv_flops_per_step(step)
            if os.environ.get("HN_CONV_TABLE"):
                json.dump(step.last_conv_table, open(os.environ["HN_CONV_TABLE"], "w"))
            # dominant launch shape: the 256->256 3x3 tower / FPN convs on the P3 level (16 + 1 launches per step)
            tbl = step.last_conv_table
            dom = [r for r in tbl if (r["cin"], r["cout"], r["k"], r["stride"]) == (256, 256, 3, 1)
                   and r["h"] * r["w"] == max(x["h"] * x["w"] for x in tbl if (x["cin"], x["cout"], x["k"]) == (256, 256, 3))]
            dom_ms = sum(r["ms"] for r in dom) / max(1, len(dom))
            dom_flop = dom[0]["gflop"] * 1e9 if dom else 0.0
            ach = dom_flop / (dom_ms * 1e-3) / 1e12 if dom else 0.0
            ach_all = flops / (conv_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "conv_igemm_kernel<256> (tcgen05 shifted GEMM), 256->256 3x3 @100x136 x8 frames",
                    "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                    # dram__bytes_read.sum + dram__bytes_write.sum of this launch, profiles/r01b_conv_igemm_full.txt
                    "traffic": 77.4e6, "algorithmic_flops_per_launch": dom_flop, "launch_ms": dom_ms,
                    "launches_per_step": len(dom), "step_share": dom_ms * len(dom) / (total_ms / args.steps),
                    "peak_source": peaks["source"] + " sustained",
                    "all_conv_launches": {"achieved": ach_all, "frac": ach_all / peaks["bf16_sustained"],
                                          "launches_per_step": n_conv, "ms_per_step_serial": conv_ms,
                                          "flops_per_step": flops}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        fps = cpu_oracle_frames_per_s(frames=2, repeats=3, threads=threads)
        cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2 VGA frames per pass, median of 3 passes after 1 warm-up (oracle/: torch-CPU restatement)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"e2e_handnet_{IMG_W}x{IMG_H}_b{B}_per_gpu", "frames_per_gpu": B,
                       "global_batch": world * B, "canvas": "800x1088", "parallelism": f"dp{world}",
                       "l2": "256 MiB buffer written between timed steps", "cuda_graph": not args.no_graph,
                       "candidates_per_frame": counts["cand"], "kept_per_frame": counts["kept"],
                       "frames_with_hand": counts["hands"], "weights": "random-init (hn_b200.synth), head bias " + str(CLS_BIAS)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_u8_ingest": {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h,
                              "api": "HandNet.forward_frames(uint8 BGR, uint16 mm)"},
            "pipelined": pipelined,
            "gpu_launches": int(launches) * args.steps,
            "gpu_launches_per_step": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
