#!/usr/bin/env python
"""Benchmark of the detect -> crop -> A2J-pose path (BASELINE.json metric: E2E frames/s, 640x480).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME]

One "step" = the whole HandNet path over one batch of FRAMES_PER_GPU synthetic 640x480 RGB + depth frames per
GPU (BASELINE.json configs[2]: batch 64 over 8 GPUs = 8 frames per GPU; weak scaling).  Random-init weights of
the reference architectures (hn_b200.synth), head biases set so that a realistic handful of boxes passes the
hard-coded 0.7 cut and every frame yields a hand crop (candidate / kept counts are reported in `config`).

  value      frames/s, inputs resident in HBM, timed with CUDA events around EXACTLY K steps submitted through the
             product's two-stage pipeline (runtime.GraphedHandNet: the pose stage of step i runs under the detect stage
             of step i+1; both stages are CUDA-graph replays); a 136 MiB buffer (> the 126 MB L2) is rewritten between
             steps INSIDE the timed region.  `sequential` (extra key) is the same with one step at a time.
  e2e        frames/s through HandNet's public API (submit / result) with HOST (pinned) inputs: H2D of the frames and
             depth maps and D2H of joints / crops / hit mask inside the timed region, host wall clock
  roofline   the dominant kernel (tcgen05 shifted-GEMM conv, ~87 % of a step): algorithmic conv FLOPs of ALL its launches
             in one step / their summed device time (CUDA events on the launch stream, launches back to back behind a
             spin kernel), against the SUSTAINED measured bf16 peak; the best launch shape is a sub-key.  `frac` is the
             product's configuration (tower launches capped to 80 % of the SMs: slower alone, as fast in the step);
             `frac_with_all_sms_per_launch` the same launches uncapped, `frac_of_step_time` conv FLOPs / the timed step;
             `event_pair_overhead_us` = what an empty CUDA-event pair reads under the same protocol, which every one of
             the ~120 timed launches carries: `frac` keeps it, `frac_net_of_event_overhead` subtracts it
  cpu_baseline  the oracle (a torch-CPU restatement of the reference, oracle/) timed on this box's host cores
  extra_configs  BASELINE.json configs 1, 2, 4 and 5 (A2J on the CPU, FCOS alone, post-process stress, 1080p strong
             scaling with up to 4 hands per frame = HandNet(max_hands=4)), the pose2mesh lifting network
             and `latency_b1` (one VGA frame per HandNet.forward call, the ros_demo.py loop: median host latency in ms):
             `--config NAME` runs one of them alone

`--impl reference` times that CPU restatement alone (the reference itself is Python that needs packages which
are not installed on the box) on the same 8-frame batch per step.
"""
from __future__ import annotations

import argparse
import glob
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
HD_H, HD_W = 1080, 1920
HD_FRAMES = 256                         # BASELINE.json configs[4]: global batch of the 1080p sweep (strong scaling)
CLS_BIAS = [-6.0, -6.0, -1.5]          # ~50 candidates, ~30 kept boxes per frame
METRIC = "e2e_frames_per_s_640x480_detect_plus_a2j_pose"
UNIT = "frames/s"
L2_FLUSH_BYTES = 136 << 20             # > 126 MB L2


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained", 1400.0), "bf16_burst": d.get("bf16_tflops", 1590.0),
                "hbm": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


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

    def summary(self, window=None):
        """Summary of the samples taken inside `window` = (t0, t1) wall-clock seconds (all samples if None)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [r for t, r in list(self.rows) if window is None or (window[0] <= t <= window[1] + 0.06)]
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

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()


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


def synthetic_frames(seed: int, n: int, h: int = IMG_H, w: int = IMG_W):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(n, 3, h, w, generator=g)
    depth = torch.rand(n, 1, h, w, generator=g) * 1.5
    return rgb, depth


# --------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/)
# --------------------------------------------------------------------------------------------------
def cpu_oracle(frames: int, repeats: int, threads: int):
    """The reference's CPU path, restated (oracle/): (frames/s of the whole path on `frames` synthetic VGA frames per pass,
    A2J batch-1 forwards/s = BASELINE.json configs[0])."""
    from hn_b200 import synth
    from oracle import a2j_oracle, handnet_oracle
    torch.set_num_threads(threads)
    fsd = synth.fcos_state_dict(3, False, seed=0, cls_bias=CLS_BIAS)
    asd = synth.a2j_state_dict(seed=1)
    rgb, depth = synthetic_frames(100, frames)
    imgs = list(rgb)
    times, a2j_times = [], []
    with torch.inference_mode():
        for r in range(repeats + 1):
            t = time.perf_counter()
            handnet_oracle.handnet_forward(fsd, asd, imgs, depth, num_classes=3)
            dt = time.perf_counter() - t
            if r > 0:                      # first pass is the warm-up
                times.append(dt)
        x = torch.rand(1, 1, 176, 176, generator=torch.Generator().manual_seed(0)) * 1.5
        for r in range(6):
            t = time.perf_counter()
            a2j_oracle.a2j_forward(asd, x)
            if r > 0:
                a2j_times.append(time.perf_counter() - t)
    times.sort()
    a2j_times.sort()
    return frames / times[len(times) // 2], 1.0 / a2j_times[len(a2j_times) // 2]


def run_reference(args):
    """--impl reference: CPU restatement of the reference path, all host threads, the same 8-frame batch per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from hn_b200 import synth
    from oracle import handnet_oracle
    torch.set_num_threads(threads)
    fsd = synth.fcos_state_dict(3, False, seed=0, cls_bias=CLS_BIAS)
    asd = synth.a2j_state_dict(seed=1)
    sample = FRAMES_PER_GPU
    rgb, depth = synthetic_frames(1000, sample)
    imgs = list(rgb)
    with torch.inference_mode():
        for _ in range(min(args.warmup, 2)):             # a CPU pass is ~1.5 s: two warm-up passes are plenty
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
        "config": {"workload": f"e2e_handnet_{IMG_W}x{IMG_H}_b{FRAMES_PER_GPU}_per_gpu", "frames_per_gpu": FRAMES_PER_GPU,
                   "frames_per_step": sample, "canvas": "800x1088",
                   "note": "reference is Python with uninstalled deps on the box; timed its CPU restatement (oracle/) on "
                           "rank 0's host cores, one 8-frame batch per step"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} VGA frames per step, {args.steps} steps"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU legs
# --------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of a GPU run."""

    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist_
            self.dist = dist_
            dist_.init_process_group("nccl", device_id=self.dev)
        self.l2_flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def pipelined_steps(ctx: Ctx, step, n: int, flush: bool, post=None, sequential: bool = False) -> float:
    """Device time (ms, CUDA events) of EXACTLY n steps through the two-stage pipeline, barrier + synchronize on both
    sides.  The L2 flush runs on the detect stream between steps, inside the timed region."""
    cur = torch.cuda.current_stream()
    step.drain()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    step.det_stream.wait_event(e0)
    for i in range(n):
        if flush:
            with torch.cuda.stream(step.det_stream):
                ctx.l2_flush.zero_()
        t = step.submit(post=post)
        if sequential:
            step.result(t)
        elif step.n_submitted - step.n_collected >= step.RING - 1:
            step.result(step.n_collected)          # the host stays at most RING-1 steps ahead
    step.drain()
    cur.wait_stream(step.det_stream)
    cur.wait_stream(step.pose_stream)
    e1.record(cur)
    ctx.barrier()
    return e0.elapsed_time(e1)


def run_main(ctx: Ctx):
    args = ctx.args
    from hn_b200 import ops, parallel
    from hn_b200.runtime import GraphedHandNet, conv_flops_per_step, conv_profile
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    net = build_net(dev)
    B = FRAMES_PER_GPU
    rgb_h, depth_h = synthetic_frames(1000 + rank, B)
    rgb_pin, depth_pin = rgb_h.pin_memory(), depth_h.pin_memory()
    step = GraphedHandNet(net, B, IMG_H, IMG_W, use_graph=not args.no_graph)
    step.load_inputs(rgb_pin.to(dev), depth_pin.to(dev))
    gathered = torch.empty((world * B, 68), dtype=torch.float32, device=dev) if world > 1 else None

    def post(rec):        # per-frame records to every rank (rank 0 consumes them): the path's only collective
        parallel.gather_records(rec, B, out=gathered)
    post_fn = post if world > 1 else None

    with torch.inference_mode():
        sampler = ClockSampler(ctx.local_rank)            # started early: nvidia-smi needs a moment before its first sample
        sampler.start()
        pipelined_steps(ctx, step, args.warmup, True, post_fn)
        # ---------------- device-resident throughput (`value`) ----------------
        t_wall0 = time.time()
        torch.cuda.nvtx.range_push("hn_timed")
        total_ms = ctx.max_over_ranks(pipelined_steps(ctx, step, args.steps, True, post_fn))
        torch.cuda.nvtx.range_pop()
        t_wall1 = time.time()
        launches = step.launches_per_step
        value = world * B * args.steps / (total_ms * 1e-3)
        # one step at a time (round-1 definition of `value`): the pose stage is exposed
        seq_ms = ctx.max_over_ranks(pipelined_steps(ctx, step, args.steps, True, post_fn, sequential=True))
        sequential = {"value": world * B * args.steps / (seq_ms * 1e-3), "unit": UNIT, "ms_per_step": seq_ms / args.steps,
                      "note": "same steps, each collected before the next is submitted (no overlap of pose and detect stages)"}
        counts = step.counts()
        if args.value_only:
            if rank == 0:
                print(json.dumps({"value": value, "ms_per_step": total_ms / args.steps, "sequential": sequential["value"],
                                  "n_gpus": world, "launches_per_step": int(launches)}), flush=True)
            sampler.stop()
            return

        # ---------------- end to end through the public API with host buffers (`e2e`) ----------------
        # what a caller of the reference does (ros_demo.py:266-273): host frames -> HandNet -> joints on the host.  The
        # asynchronous form of the same call: submit() uploads the frames (H2D on a copy stream, double-buffered) and
        # enqueues the step, result() waits for the records' D2H; two steps are kept in flight.
        imgs_pin = list(rgb_pin.unbind(0))

        def e2e_loop(submit, collect, n):
            tickets = []
            for _ in range(n):
                tickets.append(submit())
                if len(tickets) > 2:
                    collect(tickets.pop(0))
            while tickets:
                collect(tickets.pop(0))

        def timed_e2e(submit, collect):
            e2e_loop(submit, collect, 3)
            ctx.barrier()
            t0 = time.perf_counter()
            e2e_loop(submit, collect, args.steps)
            ctx.barrier()
            ms = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)          # host wall clock
            return world * B * args.steps / (ms * 1e-3)

        if world > 1:
            # multi-GPU: the records of every step are all-gathered as well (parallel.submit_sharded does this for a global
            # batch; here each rank keeps its own 8 frames, so the all-gather hook is attached directly)
            e2e_value = timed_e2e(lambda: net.submit_records(imgs_pin, depth_pin, post), net.result_records)
        else:
            e2e_value = timed_e2e(lambda: net.submit(imgs_pin, depth_pin), net.result)
        h2d = rgb_pin.numel() * 4 + depth_pin.numel() * 4
        d2h = step.d2h_bytes

        # ---------------- the same with the frames as the camera delivers them ----------------
        # uint8 BGR + uint16 millimetres on the host -> HandNet.submit_frames: H2D of 5 bytes per pixel, the
        # conversion ros_demo.py does with numpy on the host (x/255, BGR->RGB, mm/1000) runs on the device.
        bgr_pin = (rgb_h.flip(1).permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().pin_memory()
        mm_pin = (depth_h[:, 0] * 1000).round().clamp(0, 32767).to(torch.int16).contiguous().pin_memory()
        e2e_u8_value = timed_e2e(lambda: net.submit_frames(bgr_pin, mm_pin), net.result) if world == 1 else None
        h2d_u8 = bgr_pin.numel() + mm_pin.numel() * 2
        # synchronous public call (HandNet.forward, one step at a time): what an unmodified caller of the reference gets
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            net(imgs_pin, depth_images=depth_pin)
        ctx.barrier()
        e2e_sync = world * B * args.steps / (ctx.max_over_ranks((time.perf_counter() - t0) * 1e3) * 1e-3)
        t_wall2 = time.time()
        # clocks / throttle reasons sampled DURING the device-timed region; if that region was too short for two
        # samples (it lasts steps x ~3 ms), the window is extended over the e2e loops that follow it (same load)
        clocks = sampler.summary((t_wall0, t_wall1))
        clocks["window"] = "timed region"
        if (clocks.get("samples") or 0) < 2:
            clocks = sampler.summary((t_wall0, t_wall2))
            clocks["window"] = "timed region + sequential / e2e loops"
        sampler.stop()

        # ---------------- roofline of the dominant kernel (rank 0) ----------------
        roof = None
        if rank == 0:
            roof = conv_roofline(step, total_ms / args.steps, conv_profile, conv_flops_per_step)

        extras = None
        if rank == 0 and world == 1 and not args.no_extras:
            extras = {"fcos_b8": run_fcos_b8(ctx, step), "post_stress": run_post_stress(ctx, net),
                      "pose2mesh": run_pose2mesh(ctx, with_cpu=not args.no_cpu_baseline),
                      "latency_b1": run_latency_b1(ctx, net)}
        hd = None
        if not args.no_extras:
            hd = run_hd1080(ctx, net, quick=True)
            if extras is not None:
                extras["hd1080"] = hd
            elif rank == 0:
                extras = {"hd1080": hd}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        fps, a2j_fps = cpu_oracle(frames=FRAMES_PER_GPU, repeats=3, threads=threads)
        cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{FRAMES_PER_GPU} VGA frames per pass (one step), median of 3 passes after 1 warm-up "
                         "(oracle/: torch-CPU restatement)"}
        if extras is not None:
            extras["a2j_cpu"] = {"value": a2j_fps, "unit": "crops/s", "cores": threads,
                                 "workload": "BASELINE.json configs[0]: A2J ResNet-50 forward, batch 1, 176x176, CPU oracle, "
                                             "median of 5"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"e2e_handnet_{IMG_W}x{IMG_H}_b{B}_per_gpu", "frames_per_gpu": B,
                       "global_batch": world * B, "canvas": "800x1088", "parallelism": f"dp{world}",
                       "l2": f"{L2_FLUSH_BYTES >> 20} MiB buffer rewritten between steps, inside the timed region",
                       "cuda_graph": not args.no_graph,
                       "pipeline": "two stages (detect | pose) on two streams, pose of step i under detect of step i+1",
                       "candidates_per_frame": counts["cand"], "kept_per_frame": counts["kept"],
                       "frames_with_hand": counts["hands"], "weights": "random-init (hn_b200.synth), head bias " + str(CLS_BIAS)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "HandNet.submit / result (2 steps in flight), pinned host fp32 frames"},
            "e2e_sync": {"value": e2e_sync, "unit": UNIT, "api": "HandNet.forward, one step at a time"},
            "e2e_u8_ingest": None if e2e_u8_value is None else
                             {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h,
                              "api": "HandNet.submit_frames(uint8 BGR, uint16 mm) / result"},
            "sequential": sequential,
            "gpu_launches": int(launches) * args.steps,
            "gpu_launches_per_step": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "extra_configs": extras,
        }
        print(json.dumps(line), flush=True)


def conv_roofline(step, step_ms, conv_profile, conv_flops_per_step):
    """Whole-kernel roofline: every conv launch of one step.  The launches are timed back to back inside a step (behind a
    spin kernel, at the clocks a long step settles at), so the denominator is the SUSTAINED measured bf16 peak."""
    peaks = load_peaks()
    # the same launches with every SM allowed to the tower convolutions (the product caps them to ~80 % of the SMs: alone they
    # are then slower, in the step they are as fast and leave SMs to the kernels that run next to them)
    from hn_b200 import runtime as _rt
    cap_saved, _rt.TOWER_CTA_CAP = _rt.TOWER_CTA_CAP, 0
    try:
        uncapped_ms, _ = conv_profile(step, repeats=2)
    finally:
        _rt.TOWER_CTA_CAP = cap_saved
    conv_ms, n_conv = conv_profile(step, repeats=3)
    flops = conv_flops_per_step(step)
    # what the protocol itself adds to every timed launch: an event pair with NOTHING between its two records, enqueued like
    # the launches (behind a spin kernel, same stream), reads ~2.6 us.  `frac` keeps it (conservative); the net figure is reported
    # next to it
    torch.cuda._sleep(30_000_000)
    pairs = []
    for _ in range(101):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    ev_ms = sorted(a.elapsed_time(b) for a, b in pairs)[len(pairs) // 2]
    net_ms = conv_ms - n_conv * ev_ms
    tbl = step.last_conv_table
    if os.environ.get("HN_CONV_TABLE"):
        json.dump(tbl, open(os.environ["HN_CONV_TABLE"], "w"))
    ach_all = flops / (conv_ms * 1e-3) / 1e12
    # best launch shape: the 256->256 3x3 tower layers (one launch over P3+P4+P5)
    dom = [r for r in tbl if (r["cin"], r["cout"], r["k"], r["stride"]) == (256, 256, 3, 1) and r.get("levels", 1) > 1]
    if not dom:
        big = max((x["h"] * x["w"] for x in tbl if (x["cin"], x["cout"], x["k"]) == (256, 256, 3)), default=0)
        dom = [r for r in tbl if (r["cin"], r["cout"], r["k"], r["stride"]) == (256, 256, 3, 1) and r["h"] * r["w"] == big]
    best = None
    if dom:
        dom_ms = sum(r["ms"] for r in dom) / len(dom)
        dom_flop = dom[0]["gflop"] * 1e9
        ach = dom_flop / (dom_ms * 1e-3) / 1e12
        best = {"shape": "256->256 3x3, levels=%d, %dx%d x%d frames" % (dom[0].get("levels", 1), dom[0]["h"], dom[0]["w"], dom[0]["n"]),
                "achieved": ach, "frac_of_sustained": ach / peaks["bf16_sustained"], "frac_of_burst": ach / peaks["bf16_burst"],
                "algorithmic_flops_per_launch": dom_flop, "launch_ms": dom_ms, "launches_per_step": len(dom),
                "step_share": dom_ms * len(dom) / step_ms}
    # DRAM traffic of the dominant launch from a committed ncu capture, when there is one
    traffic, traffic_src = None, None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02*_conv_traffic.json"))):
        try:
            d = json.load(open(f))
            traffic, traffic_src = float(d["dram_bytes_per_launch"]), os.path.relpath(f, ROOT)
        except Exception:
            pass
    return {"bound": "tensor", "kernel": "conv_igemm_kernel (tcgen05 shifted GEMM): all %d conv launches of a step" % n_conv,
            "achieved": ach_all, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach_all / peaks["bf16_sustained"],
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peaks["source"] + ", sustained figure: launches timed back to back inside a step",
            "launches_per_step": n_conv, "ms_per_step_serial": conv_ms, "flops_per_step": flops,
            "tower_cta_cap": cap_saved if cap_saved >= 0 else "80 % of the SMs",
            "frac_with_all_sms_per_launch": flops / (uncapped_ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
            "frac_of_step_time": flops / (step_ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
            "frac_of_burst": ach_all / peaks["bf16_burst"],
            "event_pair_overhead_us": ev_ms * 1e3, "ms_per_step_serial_net_of_event_overhead": net_ms,
            "frac_net_of_event_overhead": flops / (net_ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
            "best_shape": best}


def run_fcos_b8(ctx: Ctx, step):
    """BASELINE.json configs[1]: FCOS forward + post-process (+ the crop kernel) alone, batch 8 x 640x480, one B200 = the
    detect stage's graph replayed back to back."""
    n = max(10, ctx.args.steps)
    step.drain()
    if step.g_det is None:
        return None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(step.det_stream):
        for _ in range(3):
            step.g_det.replay()
        e0.record()
        for _ in range(n):
            ctx.l2_flush.zero_()
            step.g_det.replay()
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return {"value": FRAMES_PER_GPU / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "workload": "BASELINE.json configs[1]: FCOS forward + post-process, batch 8 x 640x480, 1 GPU (L2 flush inside)"}


def run_latency_b1(ctx: Ctx, net):
    """The reference's deployment case (ros_demo.py:264-289: one camera frame per call): ``HandNet.forward`` on ONE VGA frame
    in pinned host memory, host wall clock from the call to the returned CPU joints (H2D, both stages, D2H, one host wait)."""
    rgb, depth = synthetic_frames(3000, 1)
    img = [rgb[0].pin_memory()]
    dpt = depth.pin_memory()
    for _ in range(5):
        net(img, depth_images=dpt)
    ts = []
    for _ in range(40):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        net(img, depth_images=dpt)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return {"value": ts[len(ts) // 2], "unit": "ms", "p95_ms": ts[int(len(ts) * 0.95)], "min_ms": ts[0], "higher_is_better": False,
            "frames_per_s": 1e3 / ts[len(ts) // 2], "calls": len(ts),
            "workload": "one 640x480 frame per HandNet.forward call from pinned host memory (ros_demo.py's loop), median host latency"}


def run_post_stress(ctx: Ctx, net):
    """BASELINE.json configs[3]: ~10 000 candidates per frame through decode + score + select, batched NMS and gather
    (fcos_utils/fcos.py:572-669), 8 frames.  Synthetic head tensors (hn_b200.synth.stress_head_tensors)."""
    from hn_b200 import ops, synth
    dev = ctx.dev
    peaks = load_peaks()
    lv = ops.Levels([(100, 136), (50, 68), (25, 34)], (800, 1088), (8, 16, 32))
    B = FRAMES_PER_GPU
    # channel planes [B][channel][locs], the layout the detector's output convolutions write
    ho = {k: ops.head_planes(v.to(dev)) for k, v in synth.stress_head_tensors(31, B, lv.locs, 3, -0.35).items()}
    m = net.detector
    ws_sel = torch.empty(int(ops._lib.load().hn_fcos_select_workspace_bytes(B, lv.locs)), dtype=torch.uint8, device=dev)
    ws_nms = ops.nms_workspace(B, lv.locs, dev)
    rh, rw = [480 / 800] * B, [640 / 1066] * B

    def decode():
        return ops.fcos_decode_select(ho["cls_logits"], ho["bbox_ctrness"], ho["bbox_regression"], 3, lv, m.score_cut, ws=ws_sel)

    def nms(c):
        return ops.nms_batched(c["box"], c["score"], c["label"], c["count"], m.nms_iou, m.nms_coord_trick_numel, ws=ws_nms)

    def timeit(fn, reps=10):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(reps):
            ctx.l2_flush.zero_()
            torch.cuda._sleep(3_000_000)       # the host enqueues the launches while the GPU spins: no launch gaps are timed
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    cand = decode()
    keep, kc = nms(cand)
    t_dec = timeit(decode)
    t_nms = timeit(lambda: nms(cand))

    def whole():
        c = decode()
        k, n_ = nms(c)
        ops.fcos_gather(k, n_, c, ho["hand_lr"], lv, rh, rw)
    t_all = timeit(whole)
    n8 = cand["count"].tolist()
    pairs = sum(n * n / 2 for n in n8)
    mask_bytes = sum(n * ((n + 63) // 64) * 8 for n in n8)
    dec_bytes = B * lv.locs * (3 + 1 + 4) * 4 + sum(n8) * 28
    return {"value": B / (t_all * 1e-3), "unit": UNIT, "ms_per_step": t_all,
            "workload": "BASELINE.json configs[3]: decode + NMS + gather on synthetic head tensors, 8 frames",
            "candidates_per_frame": n8, "kept_per_frame": kc.tolist(),
            "decode_select": {"ms": t_dec, "algorithmic_bytes": dec_bytes, "gb_per_s": dec_bytes / (t_dec * 1e-3) / 1e9,
                              "hbm_frac": dec_bytes / (t_dec * 1e-3) / 1e9 / peaks["hbm"]},
            "nms": {"ms": t_nms, "pair_tests": pairs, "pair_tests_per_s": pairs / (t_nms * 1e-3),
                    "bitmask_bytes_written_plus_read": 2 * mask_bytes,
                    "hbm_frac_bitmask_definition": 2 * mask_bytes / (t_nms * 1e-3) / 1e9 / peaks["hbm"],
                    "note": "SURVEY 8d: NMS is not bandwidth-bound at its algorithmic bytes; the fraction uses the bitmask "
                            "traffic n*ceil(n/64)*8 B each way as the stated definition"}}


def run_pose2mesh(ctx: Ctx, with_cpu: bool):
    """SURVEY.md 8f, last row: pose2mesh lifting (21 joints -> 1024-vertex mesh) behind models.pose2mesh_net.get_model, on the
    demo-shaped synthetic case of tests/golden/pose2mesh_case.pt (graph Laplacians from the reference's own coarsening of an
    icosphere, seeded weights): one CUDA graph per batch of 8 hands.  The forward streams 298 MB of fp32 weights whatever the
    batch (the 4096-wide PoseNet), so the roofline is HBM: weight bytes / time."""
    import scipy.sparse as sp
    sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200", "pose2mesh", "lib"))
    import models.pose2mesh_net as p2m
    from hn_b200 import synth
    case = torch.load(os.path.join(ROOT, "tests", "golden", "pose2mesh_case.pt"), weights_only=False)
    graph_L = [sp.coo_matrix((c["val"].numpy(), (c["row"].numpy(), c["col"].numpy())), shape=c["shape"]).tocsr()
               for c in case["graph_L"]]
    sd = synth.fill_state_dict(case["shapes"], seed=case["seed"])
    model = p2m.get_model(21, graph_L)
    model.load_state_dict(sd)
    model = model.to(ctx.dev).eval()
    hands = 8
    g = torch.Generator().manual_seed(3)
    joints = torch.randn(hands, 21, 2, generator=g).to(ctx.dev)
    peaks = load_peaks()
    out = {}
    with torch.no_grad():
        for _ in range(3):
            model(joints)
        torch.cuda.synchronize()
        side = torch.cuda.Stream(device=ctx.dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            model(joints)
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=side):
                mesh, pose3d = model(joints)
        torch.cuda.synchronize()
        reps = max(10, ctx.args.steps // 2)
        ts = []
        for _ in range(reps):
            ctx.l2_flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
    wbytes = sum(int(v.numel()) * 4 for k, v in sd.items() if v.dtype == torch.float32 and v.dim() == 2)
    out = {"value": hands / (ms * 1e-3), "unit": "hands/s", "ms_per_batch": ms, "hands_per_batch": hands,
           "workload": "SURVEY 8f: pose2mesh lifting, 21 joints -> 1024-vertex mesh, 8 hands per CUDA-graph replay (L2 flushed)",
           "weight_bytes": wbytes, "gb_per_s": wbytes / (ms * 1e-3) / 1e9, "hbm_frac": wbytes / (ms * 1e-3) / 1e9 / peaks["hbm"]}
    if with_cpu:
        from oracle import pose2mesh_oracle                        # CPU baseline leg only
        dense = []
        for c in case["graph_L"]:
            L = torch.zeros(c["shape"])
            L.index_put_((c["row"], c["col"]), c["val"], accumulate=True)
            dense.append(L)
        jc = joints.cpu()
        pose2mesh_oracle.flat_pose2mesh(sd, dense, jc)
        t0 = time.perf_counter()
        for _ in range(3):
            pose2mesh_oracle.flat_pose2mesh(sd, dense, jc)
        out["cpu_baseline"] = {"value": 3 * hands / (time.perf_counter() - t0), "unit": "hands/s", "cores": os.cpu_count() or 1,
                               "kind": "port", "sample": "3 batches of 8 hands (oracle/pose2mesh_oracle.py, dense torch)"}
    return out


HD_HANDS = 4          # BASELINE.json configs[4]: "up to 4 hands/frame" (HandNet(max_hands=4): the pose path on 4 hand slots per frame)


def run_hd1080(ctx: Ctx, net, quick: bool):
    """BASELINE.json configs[4]: 1920x1080 frames, GLOBAL batch 256 split over the ranks (strong scaling: 256 / 128 / 64 /
    32 frames per GPU at 1 / 2 / 4 / 8 GPUs), each rank streaming its slice through the pipeline in steps of 8 frames, up to
    4 hands per frame (`max_hands` = 4: the reference's per-box path, handnet_pipeline.py:88-102 + A2J, on the first four hand
    boxes of every frame = 32 crops per step; the reference itself keeps the first, :84-85).  `quick` (the default run's extra
    key) times one pass over the global batch after warm-up steps on every rank."""
    from hn_b200 import parallel
    from hn_b200.runtime import GraphedHandNet
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    B = FRAMES_PER_GPU
    begin, end = parallel.shard_range(HD_FRAMES, world, rank)
    n_steps = (end - begin + B - 1) // B
    rgb, depth = synthetic_frames(2000 + rank, B, HD_H, HD_W)
    saved_hands, net.max_hands = net.max_hands, HD_HANDS
    try:
        step = GraphedHandNet(net, B, HD_H, HD_W, use_graph=not ctx.args.no_graph, slot=1)
        step.load_inputs(rgb.to(dev), depth.to(dev))
        rows = B * HD_HANDS
        gathered = torch.empty((world * rows, 68), dtype=torch.float32, device=dev) if world > 1 else None
        post = (lambda rec: parallel.gather_records(rec, rows, out=gathered)) if world > 1 else None
        pipelined_steps(ctx, step, 3, True, post)
        passes = 1 if quick else 3
        ms = min(ctx.max_over_ranks(pipelined_steps(ctx, step, n_steps, True, post)) for _ in range(passes))
        counts = step.counts()
    finally:
        net.max_hands = saved_hands
    del step
    torch.cuda.empty_cache()
    return {"value": HD_FRAMES / (ms * 1e-3), "unit": UNIT, "ms_per_global_batch": ms, "scaling": "strong",
            "workload": "BASELINE.json configs[4]: 1920x1080, global batch 256 over %d GPU(s) = %d frames per GPU in steps of 8, "
                        "canvas 768x1344, up to %d hands per frame (%d crops per step)" % (world, end - begin, HD_HANDS, B * HD_HANDS),
            "hands_per_frame_max": HD_HANDS, "hand_slots_filled_last_step": counts["hands"], "hand_slots_per_step": B * HD_HANDS,
            "kept_per_frame": counts["kept"][:4]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="e2e", choices=["e2e", "a2j_cpu", "fcos_b8", "post_stress", "hd1080", "pose2mesh", "latency_b1"],
                    help="e2e (default, BASELINE.json configs[2] per-GPU slice, with the others as extra keys) or one of the "
                         "other BASELINE.json configs alone")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra_configs legs")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--value-only", action="store_true", help="experiments: only the device-timed value / sequential legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "a2j_cpu":
        threads = os.cpu_count() or 1
        _, a2j_fps = cpu_oracle(frames=1, repeats=1, threads=threads)
        print(json.dumps({"config": "a2j_cpu", "value": a2j_fps, "unit": "crops/s", "cores": threads}), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    ctx = Ctx(args)
    try:
        if args.config == "e2e":
            run_main(ctx)
        else:
            with torch.inference_mode():
                net = build_net(ctx.dev)
                if args.config == "hd1080":
                    out = run_hd1080(ctx, net, quick=False)
                elif args.config == "post_stress":
                    out = run_post_stress(ctx, net)
                elif args.config == "latency_b1":
                    out = run_latency_b1(ctx, net)
                elif args.config == "pose2mesh":
                    out = run_pose2mesh(ctx, with_cpu=not args.no_cpu_baseline)
                else:
                    from hn_b200.runtime import GraphedHandNet
                    step = GraphedHandNet(net, FRAMES_PER_GPU, IMG_H, IMG_W)
                    rgb, depth = synthetic_frames(1000, FRAMES_PER_GPU)
                    step.load_inputs(rgb.to(ctx.dev), depth.to(ctx.dev))
                    pipelined_steps(ctx, step, 3, False)
                    out = run_fcos_b8(ctx, step)
            if ctx.rank == 0:
                print(json.dumps({"config": args.config, "n_gpus": ctx.world, **out}), flush=True)
    finally:
        if ctx.world > 1:
            ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
