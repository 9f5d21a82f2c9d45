// Pipeline glue between the detector and the pose net, and the A2J anchor post-process:
//   S1 + S2  hand select, 40% box pad in the reference's int64/float32 arithmetic, depth crop, legacy-nearest
//            resize to 176x176                                   (handnet_pipeline/handnet_pipeline.py:74-102)
//   J4       softmax over the 1936 anchors per joint + weighted (anchor + offset) / depth sums
//                                                                 (a2j/anchor.py:57-82)
// Both are HBM-bound: coalesced streaming reads, warp-shuffle reductions, no tensor cores.
#include "hn_common.cuh"

#include <stdlib.h>

namespace {

// ------------------------------------------------------------------------------- S1 + S2
__device__ __forceinline__ long long trunc_ll(float v) { return (long long)v; }   // toward zero, as tensor.to(int64)

__global__ void __launch_bounds__(256)
select_crop_resize_kernel(const float4* __restrict__ boxes, const long long* __restrict__ labels,
                          const int* __restrict__ keep_count, int cap, int hand_label, int hands,
                          const float* __restrict__ depth, int depth_c, int img_h, int img_w, int out_size,
                          long long* __restrict__ crops, int* __restrict__ has_hand, float* __restrict__ depth_batch) {
  __shared__ int found_s;
  // blockIdx.y = output slot = frame * hands + h: the h-th kept detection (score-descending order) whose label is the hand
  // class.  hands == 1 is the reference (handnet_pipeline.py:84-85 keeps boxes[:1]); hands > 1 applies the same per-box path
  // to the next hand boxes of the frame.
  const int slot = blockIdx.y;
  const int b = slot / hands, h = slot - b * hands;
  const int n = min(keep_count[b], cap);
  if (threadIdx.x < 32) {                        // warp 0 walks the kept list 32 entries at a time, counting hand labels
    int seen = 0, found = 0x7fffffff;
    for (int base = 0; base < n; base += 32) {
      const int k = base + (int)threadIdx.x;
      const bool is_hand = k < n && labels[(size_t)b * cap + k] == (long long)hand_label;
      const unsigned m = __ballot_sync(0xffffffffu, is_hand);
      const int c = __popc(m);
      if (seen + c > h) {
        found = base + (int)__fns(m, 0, h - seen + 1);
        break;
      }
      seen += c;
    }
    if (threadIdx.x == 0) found_s = found;
  }
  __syncthreads();
  const int found = found_s;
  const int out_pixels = out_size * out_size;
  const size_t out_base = (size_t)slot * depth_c * out_pixels;
  if (found == 0x7fffffff) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < depth_c * out_pixels; i += gridDim.x * blockDim.x)
      depth_batch[out_base + i] = 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      has_hand[slot] = 0;
      for (int j = 0; j < 4; ++j) crops[(size_t)slot * 4 + j] = 0;
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
    has_hand[slot] = 1;
    crops[(size_t)slot * 4 + 0] = x1; crops[(size_t)slot * 4 + 1] = y1;
    crops[(size_t)slot * 4 + 2] = x2; crops[(size_t)slot * 4 + 3] = y2;
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
      val = __ldcs(depth + (((size_t)b * depth_c + c) * img_h + (ys + syi)) * img_w + (xs + sxi));   // streamed
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
constexpr int AGG_SPLITS = 16;          // anchor ranges per crop (workspace is sized for this many)

__global__ void __launch_bounds__(AGG_MAX_THREADS)
a2j_partial_kernel(const float* __restrict__ cls, const float2* __restrict__ reg, const float* __restrict__ dep,
                   const float2* __restrict__ anchor_xy, int anchors, int joints, int rows_per_iter,
                   Partial* __restrict__ part, float* __restrict__ out) {
  extern __shared__ Partial sh[];
  const int n = blockIdx.y, split = blockIdx.x, splits = gridDim.x;
  const int per = (anchors + splits - 1) / splits;
  const int a_begin = split * per, a_end = min(anchors, a_begin + per);
  const int t = threadIdx.x;
  const int j = t % joints, r = t / joints;
  Partial p = {-INFINITY, 0.f, 0.f, 0.f, 0.f};
  if (r < rows_per_iter) {
    // four anchors per round: their 16 loads are in flight together, one running-max rescale per round
    constexpr int U = 4;
    const size_t base = (size_t)n * anchors * joints;
    for (int a0 = a_begin + r; a0 < a_end; a0 += U * rows_per_iter) {
      float c[U], d[U];
      float2 rg[U], an[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int a = a0 + u * rows_per_iter;
        const bool ok = a < a_end;
        const size_t e = base + (size_t)(ok ? a : a0) * joints + j;
        c[u] = __ldcs(cls + e);                    // streamed once (as in the vector kernel below)
        rg[u] = __ldcs(reg + e);
        d[u] = __ldcs(dep + e);
        an[u] = __ldg(anchor_xy + (ok ? a : a0));
        if (!ok) c[u] = -INFINITY;
      }
      float m = p.m;
#pragma unroll
      for (int u = 0; u < U; ++u) m = fmaxf(m, c[u]);
      if (m == -INFINITY) continue;               // only -inf logits so far: nothing to add
      const float f = (p.m == -INFINITY) ? 0.f : expf(p.m - m);
      p.s *= f; p.x *= f; p.y *= f; p.d *= f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float w = expf(c[u] - m);           // 0 for the padded slots
        p.s += w;
        p.x += w * (an[u].x + rg[u].x);
        p.y += w * (an[u].y + rg[u].y);
        p.d += w * d[u];
      }
      p.m = m;
    }
  }
  sh[t] = p;
  __syncthreads();
  if (t < joints) {
    Partial acc = sh[t];
    for (int rr = 1; rr < rows_per_iter; ++rr) merge(acc, sh[rr * joints + t]);
    if (splits == 1) {                             // the crop's only block: finish here (a2j/anchor.py:73-82), no combine launch
      float* o = out + ((size_t)n * joints + t) * 3;
      o[0] = acc.x / acc.s; o[1] = acc.y / acc.s; o[2] = acc.d / acc.s;
    } else {
      part[((size_t)n * splits + split) * joints + t] = acc;
    }
  }
}

// Vectorised variant for the shipped head shape (anchors * joints a multiple of 4, 16-byte aligned tensors): the block
// streams the flat [anchors * joints] rows with 16-byte loads.  With T = rows_per_round * joints / 4 threads a round
// covers rows_per_round whole anchors (rows_per_round a multiple of 4), so thread t keeps the same four joints
// (4t + k) % joints in every round and carries four online-softmax states.  grid (splits, n); a split = a range of rounds.
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB)
a2j_partial_vec_kernel(const float4* __restrict__ cls, const float4* __restrict__ reg, const float4* __restrict__ dep,
                       const float2* __restrict__ anchor_xy, int anchors, int joints, int rows_per_round,
                       Partial* __restrict__ part, float* __restrict__ out) {
  extern __shared__ Partial sh[];                // [4 * T] then reused for the merge tree
  const int n = blockIdx.y, split = blockIdx.x, splits = gridDim.x;
  const int T = blockDim.x, t = threadIdx.x;
  const int rounds = (anchors + rows_per_round - 1) / rows_per_round;
  const int r_begin = (int)((long long)rounds * split / splits), r_end = (int)((long long)rounds * (split + 1) / splits);
  const int row_vecs = anchors * joints / 4;     // float4 per crop in cls / dep (reg has twice as many)
  const size_t base = (size_t)n * row_vecs;
  int arow[4];                                   // anchor (within a round) of each of this thread's four elements
#pragma unroll
  for (int k = 0; k < 4; ++k) arow[k] = (4 * t + k) / joints;
  Partial p[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] = Partial{-INFINITY, 0.f, 0.f, 0.f, 0.f};
  for (int r0 = r_begin; r0 < r_end; r0 += U) {
    float4 c[U], d[U], g0[U], g1[U];
    float2 an[U][4];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = (r0 + u) * T + t;              // float4 index within the crop
      ok[u] = (r0 + u) < r_end && v < row_vecs;
      const size_t e = base + (ok[u] ? v : 0);
      c[u] = __ldcs(cls + e);                      // streamed once: evict-first, no reuse
      d[u] = __ldcs(dep + e);
      g0[u] = __ldcs(reg + 2 * e);
      g1[u] = __ldcs(reg + 2 * e + 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int a = (r0 + u) * rows_per_round + arow[k];
        an[u][k] = __ldg(anchor_xy + (ok[u] && a < anchors ? a : 0));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float cc[U], xx[U], yy[U], dd[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float ck = k == 0 ? c[u].x : k == 1 ? c[u].y : k == 2 ? c[u].z : c[u].w;
        cc[u] = ok[u] ? ck : -INFINITY;
        dd[u] = k == 0 ? d[u].x : k == 1 ? d[u].y : k == 2 ? d[u].z : d[u].w;
        const float4 g = k < 2 ? g0[u] : g1[u];
        xx[u] = an[u][k].x + ((k & 1) ? g.z : g.x);
        yy[u] = an[u][k].y + ((k & 1) ? g.w : g.y);
      }
      float m = p[k].m;
#pragma unroll
      for (int u = 0; u < U; ++u) m = fmaxf(m, cc[u]);
      if (m != -INFINITY) {
        const float f = (p[k].m == -INFINITY) ? 0.f : expf(p[k].m - m);
        p[k].s *= f; p[k].x *= f; p[k].y *= f; p[k].d *= f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float w = expf(cc[u] - m);
          p[k].s += w;
          p[k].x += w * xx[u];
          p[k].y += w * yy[u];
          p[k].d += w * dd[u];
        }
        p[k].m = m;
      }
    }
  }
  // merge: entry q = 4t + k belongs to joint q % joints, anchor row q / joints
#pragma unroll
  for (int k = 0; k < 4; ++k) sh[4 * t + k] = p[k];
  __syncthreads();
  // stage 1: thread (rg, j) folds rows 4*rg .. 4*rg+3 of joint j (T = rows_per_round / 4 * joints threads exactly)
  const int j = t % joints, rg = t / joints;
  Partial acc = sh[(4 * rg) * joints + j];
#pragma unroll
  for (int i = 1; i < 4; ++i) merge(acc, sh[(4 * rg + i) * joints + j]);
  __syncthreads();
  sh[t] = acc;                                   // [rows_per_round / 4][joints]
  __syncthreads();
  if (t < joints) {
    Partial a2 = sh[t];
    for (int i = 1; i < rows_per_round / 4; ++i) merge(a2, sh[i * joints + t]);
    if (splits == 1) {                             // the crop's only block: finish here, no combine launch
      float* o = out + ((size_t)n * joints + t) * 3;
      o[0] = a2.x / a2.s; o[1] = a2.y / a2.s; o[2] = a2.d / a2.s;
    } else {
      part[((size_t)n * splits + split) * joints + t] = a2;
    }
  }
}

__global__ void a2j_combine_kernel(const Partial* __restrict__ part, int splits, int joints, int n_total,
                                   float* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_total * joints) return;
  const int n = g / joints, j = g - n * joints;
  Partial ps[AGG_SPLITS];                        // all partials are fetched before the (serial) merge chain
#pragma unroll
  for (int s = 0; s < AGG_SPLITS; ++s)
    if (s < splits) ps[s] = part[((size_t)n * splits + s) * joints + j];
  Partial acc = ps[0];
#pragma unroll
  for (int s = 1; s < AGG_SPLITS; ++s)
    if (s < splits) merge(acc, ps[s]);
  out[(size_t)g * 3 + 0] = acc.x / acc.s;
  out[(size_t)g * 3 + 1] = acc.y / acc.s;
  out[(size_t)g * 3 + 2] = acc.d / acc.s;
}

}  // namespace

extern "C" int hn_select_crop_resize_multi(const float* boxes, const int64_t* labels, const int* keep_count, int batch,
                                           int cap, int hand_label, int max_hands, const float* depth, int depth_c,
                                           int img_h, int img_w, int out_size, int64_t* crops, int* has_hand,
                                           float* depth_batch, void* stream) {
  HN_REQUIRE(boxes && labels && keep_count && depth && crops && has_hand && depth_batch,
             "hn_select_crop_resize: null pointer");
  HN_REQUIRE(batch > 0 && cap > 0 && depth_c > 0 && img_h > 0 && img_w > 0 && out_size > 0,
             "hn_select_crop_resize: bad sizes");
  HN_REQUIRE(max_hands >= 1 && (long long)batch * max_hands <= 65535, "hn_select_crop_resize: bad max_hands");
  const int per_image = depth_c * out_size * out_size;
  dim3 grid(hn_div_up(per_image, 256 * 4), batch * max_hands);
  select_crop_resize_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(boxes), reinterpret_cast<const long long*>(labels), keep_count, cap, hand_label,
      max_hands, depth, depth_c, img_h, img_w, out_size, reinterpret_cast<long long*>(crops), has_hand, depth_batch);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int hn_select_crop_resize(const float* boxes, const int64_t* labels, const int* keep_count, int batch,
                                     int cap, int hand_label, const float* depth, int depth_c, int img_h, int img_w,
                                     int out_size, int64_t* crops, int* has_hand, float* depth_batch, void* stream) {
  return hn_select_crop_resize_multi(boxes, labels, keep_count, batch, cap, hand_label, 1, depth, depth_c, img_h, img_w,
                                     out_size, crops, has_hand, depth_batch, stream);
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
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // few crops: many anchor ranges per crop for parallelism; many crops: long ranges, so that the per-block
  // shared-memory merge tail is amortised over more streamed rows
  // (512 crops: 4 ranges + combine 89 us, 2 ranges 85 us, 1 range per crop finished in place 80 us)
  int splits = n >= 64 ? AGG_SPLITS / 4 : AGG_SPLITS;
  {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        n >= 2 * sms)
      splits = 1;                                  // enough crops to fill every SM with whole-crop blocks
  }
  dim3 grid(splits, n);
  // vector path: T = rows_per_round * joints / 4 threads, rows_per_round a multiple of 4 -> T a multiple of joints
  int rows_per_round = (4 * 256 / joints) & ~3;
  const int vec_threads = rows_per_round * joints / 4;
  const bool aligned = ((reinterpret_cast<uintptr_t>(cls) | reinterpret_cast<uintptr_t>(reg) |
                         reinterpret_cast<uintptr_t>(depth)) & 15) == 0;
  const bool vec = aligned && rows_per_round >= 4 && (anchors * (long long)joints) % 4 == 0;
  if (vec) {
    const size_t sh_bytes = (size_t)4 * vec_threads * sizeof(Partial);
    // one round in flight per thread at >= 4 CTAs per SM measured best (512 crops: 85 us; two rounds in flight at 2
    // CTAs per SM 93 us; 6 CTAs per SM spills: 140 us)
    auto* kern = a2j_partial_vec_kernel<1, 4>;
    kern<<<grid, vec_threads, sh_bytes, st>>>(reinterpret_cast<const float4*>(cls), reinterpret_cast<const float4*>(reg),
                                              reinterpret_cast<const float4*>(depth),
                                              reinterpret_cast<const float2*>(anchor_xy), anchors, joints, rows_per_round,
                                              reinterpret_cast<Partial*>(workspace), out);
  } else {
    const int rows = AGG_MAX_THREADS / joints > 12 ? 12 : AGG_MAX_THREADS / joints;
    const int threads = ((rows * joints + 31) / 32) * 32;
    a2j_partial_kernel<<<grid, threads, threads * sizeof(Partial), st>>>(
        cls, reinterpret_cast<const float2*>(reg), depth, reinterpret_cast<const float2*>(anchor_xy), anchors, joints, rows,
        reinterpret_cast<Partial*>(workspace), out);
  }
  hn_count_launch();
  HN_LAUNCH_CHECK();
  if (splits > 1) {
    a2j_combine_kernel<<<hn_div_up(n * joints, 128), 128, 0, st>>>(reinterpret_cast<const Partial*>(workspace), splits,
                                                                   joints, n, out);
    hn_count_launch();
    HN_LAUNCH_CHECK();
  }
  return HN_OK;
}
