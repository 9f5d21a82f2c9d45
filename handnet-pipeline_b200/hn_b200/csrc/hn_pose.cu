// Pipeline glue between the detector and the pose net, and the A2J anchor post-process:
//   S1 + S2  hand select, 40% box pad in the reference's int64/float32 arithmetic, depth crop, legacy-nearest
//            resize to 176x176                                   (handnet_pipeline/handnet_pipeline.py:74-102)
//   J4       softmax over the 1936 anchors per joint + weighted (anchor + offset) / depth sums
//                                                                 (a2j/anchor.py:57-82)
// Both are HBM-bound: coalesced streaming reads, warp-shuffle reductions, no tensor cores.
#include "hn_common.cuh"

namespace {

// ------------------------------------------------------------------------------- S1 + S2
__device__ __forceinline__ long long trunc_ll(float v) { return (long long)v; }   // toward zero, as tensor.to(int64)

__global__ void __launch_bounds__(256)
select_crop_resize_kernel(const float4* __restrict__ boxes, const long long* __restrict__ labels,
                          const int* __restrict__ keep_count, int cap, int hand_label,
                          const float* __restrict__ depth, int depth_c, int img_h, int img_w, int out_size,
                          long long* __restrict__ crops, int* __restrict__ has_hand, float* __restrict__ depth_batch) {
  __shared__ int found_s;
  const int b = blockIdx.y;
  const int n = min(keep_count[b], cap);
  // first kept detection (score-descending order) whose label is the hand class
  if (threadIdx.x == 0) found_s = 0x7fffffff;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int k = base + threadIdx.x;
    if (k < n && labels[(size_t)b * cap + k] == (long long)hand_label) atomicMin(&found_s, k);
    __syncthreads();
    const int f = found_s;
    __syncthreads();
    if (f != 0x7fffffff) break;
  }
  const int found = found_s;
  const int out_pixels = out_size * out_size;
  const size_t out_base = (size_t)b * depth_c * out_pixels;
  if (found == 0x7fffffff) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < depth_c * out_pixels; i += gridDim.x * blockDim.x)
      depth_batch[out_base + i] = 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      has_hand[b] = 0;
      for (int j = 0; j < 4; ++j) crops[(size_t)b * 4 + j] = 0;
    }
    return;
  }
  // handnet_pipeline.py:88-97.  box -> int64 (truncate); percent*w is a python float times a 0-dim int64 tensor:
  // float32 arithmetic; max(0, t) / min(W, t) keep t unless the bound wins; stores truncate toward zero.
  const float4 bx = boxes[(size_t)b * cap + found];
  const long long b0 = trunc_ll(bx.x), b1 = trunc_ll(bx.y), b2 = trunc_ll(bx.z), b3 = trunc_ll(bx.w);
  const float pw = __fmul_rn(0.4f, (float)(b2 - b0));
  const float ph = __fmul_rn(0.4f, (float)(b3 - b1));
  float v;
  v = __fsub_rn((float)b0, pw);
  const long long x1 = (v > 0.f) ? trunc_ll(v) : 0;
  v = __fsub_rn((float)b1, ph);
  const long long y1 = (v > 0.f) ? trunc_ll(v) : 0;
  v = __fadd_rn((float)b2, pw);
  const long long x2 = (v < (float)img_w) ? trunc_ll(v) : img_w;
  v = __fadd_rn((float)b3, ph);
  const long long y2 = (v < (float)img_h) ? trunc_ll(v) : img_h;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    has_hand[b] = 1;
    crops[(size_t)b * 4 + 0] = x1; crops[(size_t)b * 4 + 1] = y1;
    crops[(size_t)b * 4 + 2] = x2; crops[(size_t)b * 4 + 3] = y2;
  }
  // python slice [y1 : y2+1, x1 : x2+1] clamps to the image
  const long long ys = min(max(y1, 0ll), (long long)img_h), ye = min(max(y2 + 1, 0ll), (long long)img_h);
  const long long xs = min(max(x1, 0ll), (long long)img_w), xe = min(max(x2 + 1, 0ll), (long long)img_w);
  const int ih = (int)(ye - ys), iw = (int)(xe - xs);
  const float sy = (float)ih / (float)out_size, sx = (float)iw / (float)out_size;   // ATen: scale = in / out (fp32)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < depth_c * out_pixels; i += gridDim.x * blockDim.x) {
    const int c = i / out_pixels;
    const int pix = i - c * out_pixels;
    const int oy = pix / out_size, ox = pix - oy * out_size;
    float val = 0.f;
    if (ih > 0 && iw > 0) {
      // upsample_nearest (legacy): src = min(floor(dst * scale), in - 1)
      const int syi = min((int)floorf(__fmul_rn((float)oy, sy)), ih - 1);
      const int sxi = min((int)floorf(__fmul_rn((float)ox, sx)), iw - 1);
      val = __ldg(depth + (((size_t)b * depth_c + c) * img_h + (ys + syi)) * img_w + (xs + sxi));
    }
    depth_batch[out_base + i] = val;
  }
}

// ------------------------------------------------------------------------------------- J4
// grid (splits, n).  Thread t owns joint t % J and anchors (t / J) + k * (T / J) of its split, so the block reads
// contiguous rows of cls / reg / depth.  Online softmax per thread, then a shared-memory combine per joint.
struct Partial {
  float m, s, x, y, d;
};

__device__ __forceinline__ void merge(Partial& a, const Partial& b) {
  const float m = fmaxf(a.m, b.m);
  const float fa = (a.m == -INFINITY) ? 0.f : expf(a.m - m);
  const float fb = (b.m == -INFINITY) ? 0.f : expf(b.m - m);
  a.s = a.s * fa + b.s * fb;
  a.x = a.x * fa + b.x * fb;
  a.y = a.y * fa + b.y * fb;
  a.d = a.d * fa + b.d * fb;
  a.m = m;
}

constexpr int AGG_MAX_THREADS = 512;

__global__ void __launch_bounds__(AGG_MAX_THREADS)
a2j_partial_kernel(const float* __restrict__ cls, const float2* __restrict__ reg, const float* __restrict__ dep,
                   const float2* __restrict__ anchor_xy, int anchors, int joints, int rows_per_iter,
                   Partial* __restrict__ part) {
  extern __shared__ Partial sh[];
  const int n = blockIdx.y, split = blockIdx.x, splits = gridDim.x;
  const int per = (anchors + splits - 1) / splits;
  const int a_begin = split * per, a_end = min(anchors, a_begin + per);
  const int t = threadIdx.x;
  const int j = t % joints, r = t / joints;
  Partial p = {-INFINITY, 0.f, 0.f, 0.f, 0.f};
  if (r < rows_per_iter) {
    const size_t base = (size_t)n * anchors * joints;
    for (int a = a_begin + r; a < a_end; a += rows_per_iter) {
      const size_t e = base + (size_t)a * joints + j;
      const float c = __ldg(cls + e);
      const float2 rg = __ldg(reg + e);
      const float d = __ldg(dep + e);
      const float2 an = __ldg(anchor_xy + a);
      const float m = fmaxf(p.m, c);
      const float f = (p.m == -INFINITY) ? 0.f : expf(p.m - m);
      const float w = expf(c - m);
      p.s = p.s * f + w;
      p.x = p.x * f + w * (an.x + rg.x);
      p.y = p.y * f + w * (an.y + rg.y);
      p.d = p.d * f + w * d;
      p.m = m;
    }
  }
  sh[t] = p;
  __syncthreads();
  if (t < joints) {
    Partial acc = sh[t];
    for (int rr = 1; rr < rows_per_iter; ++rr) merge(acc, sh[rr * joints + t]);
    part[((size_t)n * splits + split) * joints + t] = acc;
  }
}

__global__ void a2j_combine_kernel(const Partial* __restrict__ part, int splits, int joints, int n_total,
                                   float* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_total * joints) return;
  const int n = g / joints, j = g - n * joints;
  Partial acc = part[((size_t)n * splits) * joints + j];
  for (int s = 1; s < splits; ++s) merge(acc, part[((size_t)n * splits + s) * joints + j]);
  out[(size_t)g * 3 + 0] = acc.x / acc.s;
  out[(size_t)g * 3 + 1] = acc.y / acc.s;
  out[(size_t)g * 3 + 2] = acc.d / acc.s;
}

constexpr int AGG_SPLITS = 16;

}  // namespace

extern "C" int hn_select_crop_resize(const float* boxes, const int64_t* labels, const int* keep_count, int batch,
                                     int cap, int hand_label, const float* depth, int depth_c, int img_h, int img_w,
                                     int out_size, int64_t* crops, int* has_hand, float* depth_batch, void* stream) {
  HN_REQUIRE(boxes && labels && keep_count && depth && crops && has_hand && depth_batch,
             "hn_select_crop_resize: null pointer");
  HN_REQUIRE(batch > 0 && cap > 0 && depth_c > 0 && img_h > 0 && img_w > 0 && out_size > 0,
             "hn_select_crop_resize: bad sizes");
  const int per_image = depth_c * out_size * out_size;
  dim3 grid(hn_div_up(per_image, 256 * 4), batch);
  select_crop_resize_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(boxes), reinterpret_cast<const long long*>(labels), keep_count, cap, hand_label,
      depth, depth_c, img_h, img_w, out_size, reinterpret_cast<long long*>(crops), has_hand, depth_batch);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int64_t hn_a2j_workspace_bytes(int n, int joints) {
  return (int64_t)n * AGG_SPLITS * joints * (int64_t)sizeof(Partial);
}

extern "C" int hn_a2j_aggregate(const float* cls, const float* reg, const float* depth, const float* anchor_xy, int n,
                                int anchors, int joints, float* out, void* workspace, int64_t workspace_bytes,
                                void* stream) {
  HN_REQUIRE(cls && reg && depth && anchor_xy && out && workspace, "hn_a2j_aggregate: null pointer");
  HN_REQUIRE(n > 0 && anchors > 0 && joints > 0 && joints <= AGG_MAX_THREADS, "hn_a2j_aggregate: bad sizes");
  HN_REQUIRE(workspace_bytes >= hn_a2j_workspace_bytes(n, joints), "hn_a2j_aggregate: workspace too small");
  const int rows = AGG_MAX_THREADS / joints > 12 ? 12 : AGG_MAX_THREADS / joints;
  const int threads = ((rows * joints + 31) / 32) * 32;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(AGG_SPLITS, n);
  a2j_partial_kernel<<<grid, threads, threads * sizeof(Partial), st>>>(
      cls, reinterpret_cast<const float2*>(reg), depth, reinterpret_cast<const float2*>(anchor_xy), anchors, joints, rows,
      reinterpret_cast<Partial*>(workspace));
  hn_count_launch();
  HN_LAUNCH_CHECK();
  a2j_combine_kernel<<<hn_div_up(n * joints, 128), 128, 0, st>>>(reinterpret_cast<const Partial*>(workspace), AGG_SPLITS,
                                                                 joints, n, out);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}
