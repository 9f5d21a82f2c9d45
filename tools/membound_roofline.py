"""Achieved HBM bandwidth of the memory-bound kernels at batch sizes where the question is meaningful
(algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json hbm_gbs)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from hn_b200 import ops
from oracle import a2j_oracle
from oracle.golden_inputs import stress_head_tensors

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(136 << 20, dtype=torch.uint8, device="cuda")

REPS = int(os.environ.get("HN_MEMBOUND_REPS", "0"))     # 1 under ncu (each profiled launch is replayed ~40 times)

def timeit(fn, reps=10):
    if REPS: reps = REPS
    for _ in range(1 if REPS else 3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(3_000_000)     # ~1.5 ms spin: the host enqueues everything before the GPU reaches e0 (no launch gaps inside)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3

def report(name, nbytes, sec, note=""):
    gbs = nbytes / sec / 1e9
    print(f"{name:34s} {nbytes / 1e6:9.1f} MB  {sec * 1e6:9.1f} us  {gbs:8.0f} GB/s  {100 * gbs / PEAK:5.1f}% of {PEAK:.0f}  {note}", flush=True)

B = 64
# T1: 64 VGA frames -> 800x1088x4 bf16 canvas
imgs = [torch.rand(3, 480, 640, device="cuda") for _ in range(B)]
canvas = torch.empty((B, 800, 1088, 4), dtype=torch.bfloat16, device="cuda")
report("preprocess (T1) x64", B * (3 * 480 * 640 * 4 + 800 * 1088 * 4 * 2),
       timeit(lambda: ops.preprocess(imgs, [(800, 1066)] * B, (800, 1088), (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), canvas=canvas)))
# stem patches
a = torch.empty((B * 400 * 544, 256), dtype=torch.bfloat16, device="cuda")
report("im2col 7x7/2 (stem) x64", B * (800 * 1088 * 8 + 400 * 544 * 512), timeit(lambda: ops.im2col_7x7s2(canvas, 256, out=a)))
# maxpool
x = torch.randn((B, 400, 544, 64), device="cuda").to(torch.bfloat16)
pooled = ops.Act(B, 200, 272, 64, 1, "cuda")
report("maxpool 3x3/2 x64", B * (400 * 544 * 64 * 2 + 200 * 272 * 64 * 2), timeit(lambda: ops.maxpool3x3s2(x, pooled)))
del x, a
# groupnorm apply on P3 x64
act = ops.Act(B, 100, 136, 256, 1, "cuda"); act.interior().normal_()
stats = torch.zeros(B, 32, 2, dtype=torch.int64, device="cuda"); stats[..., 1] = int(100 * 136 * 8 * ops.GN_FIX_SCALE)
g = torch.ones(256, device="cuda"); bta = torch.zeros(256, device="cuda")
report("groupnorm+relu P3 x64", B * 100 * 136 * 256 * 2 * 2, timeit(lambda: ops.groupnorm_relu(act, stats, 32, g, bta)))
# decode + select, batch 256, ~55 % survivors
Bp = 256
lv = ops.Levels([(100, 136), (50, 68), (25, 34)], (800, 1088), (8, 16, 32))
ho = {k: v.cuda() for k, v in stress_head_tensors(31, 4, lv.locs, 3, -0.35).items()}
ho = {k: ops.head_planes(v.repeat(Bp // 4, 1, 1)) for k, v in ho.items()}   # channel planes, as the detector writes them
cand = ops.fcos_decode_select(ho["cls_logits"], ho["bbox_ctrness"], ho["bbox_regression"], 3, lv, 0.7)
ncand = int(cand["count"].sum())
report("decode+score+select x256 (stress)", Bp * lv.locs * (3 + 1 + 4) * 4 + ncand * 28,
       timeit(lambda: ops.fcos_decode_select(ho["cls_logits"], ho["bbox_ctrness"], ho["bbox_regression"], 3, lv, 0.7)),
       f"{ncand // Bp} candidates/frame")
lo = dict(ho); lo["cls_logits"] = ops.head_planes(ho["cls_logits"] - 3.0)
c2 = ops.fcos_decode_select(lo["cls_logits"], lo["bbox_ctrness"], lo["bbox_regression"], 3, lv, 0.7)
report("decode+score+select x256 (sparse)", Bp * lv.locs * (3 + 1) * 4 + int(c2["count"].sum()) * 44,
       timeit(lambda: ops.fcos_decode_select(lo["cls_logits"], lo["bbox_ctrness"], lo["bbox_regression"], 3, lv, 0.7)),
       f"{int(c2['count'].sum()) // Bp} candidates/frame")
# NMS stress: 8 frames x ~10k candidates
c8 = {k: v[:8].contiguous() for k, v in cand.items()}
ws = ops.nms_workspace(8, lv.locs, "cuda")
n8 = c8["count"].tolist()
t = timeit(lambda: ops.nms_batched(c8["box"], c8["score"], c8["label"], c8["count"], 0.3, 4000, ws=ws), reps=5)
pairs = sum(n * n / 2 for n in n8)
mask_bytes = sum(n * ((n + 63) // 64) * 8 for n in n8)
report("NMS x8 (~10k cand/frame) bitmask", 2 * mask_bytes / 2, t, f"{pairs / t / 1e9:.0f} G pair-tests/s (upper triangle only)")
# A2J aggregation, 512 crops
n = 512
cls = torch.randn(n, 1936, 21, device="cuda"); reg = torch.randn(n, 1936, 21, 2, device="cuda"); dep = torch.randn(n, 1936, 21, device="cuda")
anc = a2j_oracle.all_anchors().cuda()
report("A2J aggregate x512 crops", n * 1936 * 21 * 4 * 4, timeit(lambda: ops.a2j_aggregate(cls, reg, dep, anc)))
# crop + resize, 256 frames
boxes = torch.tensor([[100.0, 80.0, 400.0, 380.0]]).repeat(256, 1).reshape(256, 1, 4).cuda()
labels = torch.full((256, 1), 2, dtype=torch.int64, device="cuda"); kc = torch.ones(256, dtype=torch.int32, device="cuda")
depth = torch.rand(256, 1, 480, 640, device="cuda")
report("select+crop+resize x256", 256 * (421 * 421 * 4 + 176 * 176 * 4), timeit(lambda: ops.select_crop_resize(boxes, labels, kc, 2, depth)),
       "reads only the pixels the nearest rule samples; bytes counted as whole crop")
