// FCOS post-processing (fcos_utils/fcos.py:572-669) as batch-wide kernels:
//   decode + score + threshold with ordered compaction, ONE pass (look-back over per-chunk survivor counts)
//                                                              (P1-P4, anchors generated on the fly: A1)
//   stable sort by score (rank by counting up to 1024 candidates, radix beyond) + 64x64 IoU bitmask + serial scan
//                                                              (P5, torchvision CPU NMS semantics, bit-exact)
//   gather of the kept detections + resize_boxes                (P6)
// Integer/fp32 work on tiny data: coalesced vector loads, warp ballots/matches, no tensor cores.
#include "hn_common.cuh"

#include <math.h>

namespace {

// x > (double)t for a float x  <=>  x >= smallest float whose value exceeds t.  This is how a python-float
// threshold meets float32 data in the reference (torchvision's CPU NMS compares against a double; `scores > 0.7`
// promotes to float32, which gives the same set for 0.7).
float float_gt_as_ge(double t) {
  float f = (float)t;
  if ((double)f <= t) f = nextafterf(f, INFINITY);
  while ((double)nextafterf(f, -INFINITY) > t) f = nextafterf(f, -INFINITY);
  return f;
}

constexpr int MAX_LEVELS = 8;
struct Levels {
  int n;
  int start[MAX_LEVELS + 1];
  int h[MAX_LEVELS], w[MAX_LEVELS], sh[MAX_LEVELS], sw[MAX_LEVELS], anchor[MAX_LEVELS];
};

constexpr int SEL_BLOCK = 256;

__device__ __forceinline__ float sigmoid_rn(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// A head tensor [batch][locs][channels] of any layout: element (b, loc, c) at p[b * img + loc * loc_stride + c * chan_stride].
// The detector writes its heads channel-PLANAR ([batch][channel][locs]: loc_stride 1, chan_stride locs), so that a warp's
// loads of one channel are one contiguous 128-byte run and no padding bytes travel; row layouts work as well.
struct HeadView {
  const float* p;
  long long img;
  int loc, chan;
};

// ------------------------------------------------------------------------ decode + score + select (P1-P4), one pass
// score = max_c sqrt(sigmoid(cls_c) * sigmoid(ctr)), label = first arg max (fcos_utils/fcos.py:598-599).  Every rounding of
// the chain (IEEE division, multiply, square root) is monotone and CUDA's expf is within 2 ulp, so the class with the largest
// LOGIT has the largest score; an EARLIER class can only tie with it if its logit is within a hair of the largest one (or both
// sit in the saturated tail, where sigmoid rounds to 1.0).  So the kernel tracks the two largest logits of a location while it
// streams the class planes, evaluates ONE sigmoid chain for the arg-max logit, and falls back to the reference's loop over all
// classes (score_label_exact, re-reading the logits) only when the runner-up is inside that neighbourhood: first-max label
// semantics bit for bit at one chain per location instead of one per class.
__device__ __noinline__ void score_label_exact(const float* __restrict__ cls, size_t chan, int nc, float sc, float& best,
                                               int& label) {
  best = -1.f;
  label = 0;
  for (int c = 0; c < nc; ++c) {
    const float s = __fsqrt_rn(__fmul_rn(sigmoid_rn(__ldg(cls + c * chan)), sc));
    if (c == 0 || s > best) { best = s; label = c; }
  }
}

// neighbourhood of the largest logit in which another class's score may round to the same float: tiny below 4, a full unit
// up to 15 (the sigmoid's slope there is ~1e-6 per ulp of its value), everything beyond (saturation).  Far in the negative
// tail (exp overflows, scores underflow to equal values) the exact loop runs as well.
__device__ __forceinline__ bool needs_exact(float mx, float second, float ctr) {
  const float margin = mx <= 4.f ? 1e-4f * (1.f + fabsf(mx)) : (mx <= 15.f ? 1.f : INFINITY);
  return second >= mx - margin || mx < -80.f || ctr < -80.f;
}

constexpr int SEL_PER = 4;                        // consecutive locations of a thread: one 16-byte load per plane
constexpr int SEL_WARPS = SEL_BLOCK / 32;
constexpr int SEL_ROUND = 32 * SEL_PER;           // locations a warp takes per round
constexpr int SEL_CHUNK = SEL_BLOCK * SEL_PER;    // locations of a block per round

__device__ __forceinline__ void top2(float v, int c, float& mx, float& second, int& cm) {
  if (v > mx) { second = mx; mx = v; cm = c; }
  else second = fmaxf(second, v);
}

// One block = one chunk of G * SEL_CHUNK locations of one image (grid = batch * chunks, image-major; G = rounds per block, picked
// by the host so that the grid fills whole waves of resident blocks); a warp owns G * 128 consecutive locations and takes them
// in G rounds.  All rounds' loads are issued up front as 16-byte cp.async copies into the warp's own shared-memory staging
// (one commit group per round): the bytes in flight per SM are bounded by shared memory, not by registers, and a lane reads back
// exactly what it copied (no block barrier).  Survivors are parked, in location order, in the staging of the round they came
// from (consumed by then); the block counts them, publishes the count in state[] and sums the counts of the chunks before it in
// the same image (they have lower block indices: scheduled no later than this block, so the wait cannot deadlock; a chunk
// without survivors does not look back at all).  Then every survivor decodes its box (det_utils.py:266-294, anchors generated
// on the fly: anchor_utils.py:56-112) and goes to its place in the image's candidate list, ascending in location, with
// coalesced stores.
//
// A location whose centre-ness logit or whose largest class logit is below `skip_below` cannot reach the threshold
// (sqrt(sigmoid(c) * sigmoid(t)) <= sqrt(sigmoid(min(c, t)))), so in the sparse regime a detector normally runs in the
// kernel is a pure stream over the cls / ctr planes.  NC > 0: class count known at compile time (staged planes); NC = 0: any
// class count up to 256 (direct loads).
// state word: 0 = not published yet (the host clears the array before the launch), else survivors + 1.
constexpr int SEL_MAX_ROUNDS = 4;
template <int NC>
struct SelStage {
  static constexpr int PLANES = NC > 0 ? NC + 1 : 2;            // staged planes per round (>= 2: room for 128 parked survivors)
  static constexpr int ROUND_FLOATS = PLANES * SEL_ROUND;       // per warp and round
};

__device__ __forceinline__ void sel_cp16(float* dst, const float* src) {
  // the head planes are streamed once: L2 evict-first, so that they replace each other instead of lines somebody will read
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(hn_smem_u32(dst)), "l"(src), "l"(pol) : "memory");
}

template <int NC>
__global__ void __launch_bounds__(SEL_BLOCK)
select_decode_kernel(const HeadView cls, const HeadView ctr, const HeadView reg, int locs, int nc_rt, int chunks, int G,
                     float thresh_ge, float skip_below, const Levels lv, unsigned* __restrict__ state,
                     int* __restrict__ cand_count, int* __restrict__ cand_loc, float* __restrict__ cand_score,
                     int* __restrict__ cand_label, float4* __restrict__ cand_box) {
  extern __shared__ __align__(16) float stage[];          // [SEL_WARPS][G][PLANES][SEL_ROUND]
  __shared__ int warp_sums[SEL_WARPS];
  __shared__ int round_count[SEL_WARPS][SEL_MAX_ROUNDS];
  __shared__ int base_s;
  constexpr int RF = SelStage<NC>::ROUND_FLOATS;
  const int nc = NC > 0 ? NC : nc_rt;
  const int b = blockIdx.x / chunks, chunk = blockIdx.x - b * chunks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warp_loc0 = (chunk * SEL_WARPS + warp) * (SEL_ROUND * G);
  const float* cbase = cls.p + (size_t)b * cls.img;
  const float* tbase = ctr.p + (size_t)b * ctr.img;
  const size_t chan = (size_t)cls.chan;
  // channel planes whose rows are 16-byte aligned: vector loads (every loc0 is a multiple of 4)
  const bool planes = cls.loc == 1 && ctr.loc == 1 &&
                      ((reinterpret_cast<uintptr_t>(cbase) | reinterpret_cast<uintptr_t>(tbase) | (chan * 4)) & 15) == 0;
  float* my_stage = stage + (size_t)warp * G * RF;
  if (NC > 0 && planes) {
    for (int g = 0; g < G; ++g) {
      const int loc0 = warp_loc0 + g * SEL_ROUND + lane * SEL_PER;
      if (loc0 + SEL_PER <= locs) {
        float* d = my_stage + g * RF + lane * SEL_PER;
        sel_cp16(d, tbase + loc0);
#pragma unroll
        for (int c = 0; c < NC; ++c) sel_cp16(d + (c + 1) * SEL_ROUND, cbase + c * chan + loc0);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  int nsurv = 0;                                          // survivors of this warp so far (warp-uniform)
#pragma unroll 1
  for (int g = 0; g < G; ++g) {
    const int loc0 = warp_loc0 + g * SEL_ROUND + lane * SEL_PER;
    const bool vec = planes && loc0 + SEL_PER <= locs;
    float tt[SEL_PER], mx[SEL_PER], second[SEL_PER];
    int cm[SEL_PER];
#pragma unroll
    for (int i = 0; i < SEL_PER; ++i) { mx[i] = -INFINITY; second[i] = -INFINITY; cm[i] = 0; tt[i] = -INFINITY; }
    if (NC > 0 && planes) {
      switch (G - 1 - g) {                                // this round's copies have landed (later rounds may still fly)
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
      }
    }
    if (vec) {
      if (NC > 0) {
        const float* d = my_stage + g * RF + lane * SEL_PER;
        const float4 t4 = *reinterpret_cast<const float4*>(d);
        tt[0] = t4.x; tt[1] = t4.y; tt[2] = t4.z; tt[3] = t4.w;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(d + (c + 1) * SEL_ROUND);
          top2(v.x, c, mx[0], second[0], cm[0]);
          top2(v.y, c, mx[1], second[1], cm[1]);
          top2(v.z, c, mx[2], second[2], cm[2]);
          top2(v.w, c, mx[3], second[3], cm[3]);
        }
      } else {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(tbase + loc0));
        tt[0] = t4.x; tt[1] = t4.y; tt[2] = t4.z; tt[3] = t4.w;
#pragma unroll 4
        for (int c = 0; c < nc; ++c) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(cbase + c * chan + loc0));
          top2(v.x, c, mx[0], second[0], cm[0]);
          top2(v.y, c, mx[1], second[1], cm[1]);
          top2(v.z, c, mx[2], second[2], cm[2]);
          top2(v.w, c, mx[3], second[3], cm[3]);
        }
      }
    } else if (loc0 < locs) {
      for (int i = 0; i < SEL_PER; ++i) {                 // the ragged end of an image, or a row layout
        const int loc = loc0 + i;
        if (loc < locs) {
          tt[i] = __ldg(tbase + (size_t)loc * ctr.loc);
          for (int c = 0; c < nc; ++c) top2(__ldg(cbase + (size_t)loc * cls.loc + c * chan), c, mx[i], second[i], cm[i]);
        }
      }
    }
    float s[SEL_PER];
    int mine = 0;
    unsigned pmask = 0;
#pragma unroll
    for (int i = 0; i < SEL_PER; ++i) {
      s[i] = 0.f;
      if (loc0 + i < locs && tt[i] >= skip_below && mx[i] >= skip_below) {
        const float sc = sigmoid_rn(tt[i]);
        if (needs_exact(mx[i], second[i], tt[i]))
          score_label_exact(cbase + (size_t)(loc0 + i) * cls.loc, chan, nc, sc, s[i], cm[i]);
        else
          s[i] = __fsqrt_rn(__fmul_rn(sigmoid_rn(mx[i]), sc));
        if (s[i] >= thresh_ge) { pmask |= 1u << i; ++mine; }
      }
    }
    int cnt = 0;
    if (__any_sync(0xffffffffu, mine != 0)) {
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      // every lane has read its staged logits (the shuffles above are warp-wide): the round's staging is free for parking
      uint2* parked = reinterpret_cast<uint2*>(my_stage + g * RF);
      int q = incl - mine;
#pragma unroll
      for (int i = 0; i < SEL_PER; ++i)
        if (pmask & (1u << i)) parked[q++] = make_uint2((unsigned)(loc0 + i) | ((unsigned)cm[i] << 24), __float_as_uint(s[i]));
      cnt = __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) round_count[warp][g] = cnt;
    nsurv += cnt;
  }
  if (lane == 0) warp_sums[warp] = nsurv;
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int i = 0; i < SEL_WARPS; ++i) {
    const int c = warp_sums[i];
    if (i < warp) before += c;
    total += c;
  }
  unsigned* st = state + (size_t)b * chunks;
  const bool last = chunk == chunks - 1;
  if (warp == 0) {
    if (lane == 0) atomicExch(st + chunk, (unsigned)total + 1u);           // publish
    if (total != 0 || last) {
      int acc = 0;
      const long long t0 = clock64();
      for (int i = lane; i < chunk; i += 32) {                              // look back
        unsigned v;
        while ((v = *reinterpret_cast<volatile unsigned*>(st + i)) == 0u) {
          __nanosleep(40);
          if (clock64() - t0 > 4000000000LL) asm volatile("trap;");         // ~2 s: a predecessor never ran
        }
        acc += (int)(v - 1u);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        base_s = acc;
        if (last) cand_count[b] = acc + total;                              // the last chunk knows the image's total
      }
    }
  }
  if (total == 0) return;
  __syncthreads();
  const float* rbase = reg.p + (size_t)b * reg.img;
  size_t out0 = (size_t)b * locs + base_s + before;
  for (int g = 0; g < G; ++g) {
    const int cnt = round_count[warp][g];
    const uint2* parked = reinterpret_cast<const uint2*>(my_stage + g * RF);
    for (int q = lane; q < cnt; q += 32) {
      const uint2 e = parked[q];
      const int loc = (int)(e.x & 0xffffffu);
      int l = 0;
      while (l + 1 < lv.n && loc >= lv.start[l + 1]) ++l;
      const int cell = loc - lv.start[l];
      const int y = cell / lv.w[l], x = cell - y * lv.w[l];
      // anchor = (x*sw, y*sh, x*sw, y*sh) + round([-s,-s,s,s]/2)   (anchor_utils.py:56-112)
      const float half = rintf((float)lv.anchor[l] * 0.5f);
      const float a0 = (float)(x * lv.sw[l]) - half, a1 = (float)(y * lv.sh[l]) - half;
      const float a2 = (float)(x * lv.sw[l]) + half, a3 = (float)(y * lv.sh[l]) + half;
      const float cx = __fmul_rn(0.5f, __fadd_rn(a0, a2)), cy = __fmul_rn(0.5f, __fadd_rn(a1, a3));
      const float bw = __fsub_rn(a2, a0), bh = __fsub_rn(a3, a1);
      const float* rp = rbase + (size_t)loc * reg.loc;
      const float r0 = __ldg(rp), r1 = __ldg(rp + (size_t)reg.chan), r2 = __ldg(rp + 2 * (size_t)reg.chan),
                  r3 = __ldg(rp + 3 * (size_t)reg.chan);
      float4 box;
      box.x = __fsub_rn(cx, __fmul_rn(r0, bw));
      box.y = __fsub_rn(cy, __fmul_rn(r1, bh));
      box.z = __fadd_rn(cx, __fmul_rn(r2, bw));
      box.w = __fadd_rn(cy, __fmul_rn(r3, bh));
      const size_t o = out0 + q;
      cand_loc[o] = loc;
      cand_score[o] = __uint_as_float(e.y);
      cand_label[o] = (int)(e.x >> 24);
      cand_box[o] = box;
    }
    out0 += cnt;
  }
}

// ------------------------------------------------------------------------------------- NMS
constexpr int SORT_THREADS = 1024;
constexpr int SORT_WARPS = SORT_THREADS / 32;

struct NmsWs {
  uint32_t* keys[2];      // [batch][cap]
  uint32_t* idx[2];       // [batch][cap]
  float4* sbox;           // [batch][cap] boxes in sorted order (coordinate trick applied when active)
  int* slabel;            // [batch][cap] labels in sorted order (0 when the coordinate trick is active)
  unsigned long long* mask;  // [batch][cap][words]
  int words;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

__host__ NmsWs carve(void* base, int batch, int cap) {
  NmsWs w;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  const size_t n = (size_t)batch * cap;
  w.words = (cap + 63) / 64;
  for (int i = 0; i < 2; ++i) { w.keys[i] = reinterpret_cast<uint32_t*>(p); p += align256(n * 4); }
  for (int i = 0; i < 2; ++i) { w.idx[i] = reinterpret_cast<uint32_t*>(p); p += align256(n * 4); }
  w.sbox = reinterpret_cast<float4*>(p); p += align256(n * 16);
  w.slabel = reinterpret_cast<int*>(p); p += align256(n * 4);
  w.mask = reinterpret_cast<unsigned long long*>(p);
  return w;
}

__device__ __forceinline__ uint32_t desc_key(float s) {
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending-orderable
  return ~u;                                         // descending
}

// Lists long enough for the radix path that take torchvision's per-class branch are processed in class-major order (see
// nms_sort_kernel); every kernel of the chain derives the mode from (n, trick_max) alone.
__host__ __device__ inline bool class_major(int n, int trick_max) {
  return n > SORT_THREADS && 4ll * n > (long long)trick_max;
}

// One CTA per image: stable LSD radix sort of (score desc) with candidate index payload, then gather the boxes
// in sorted order, applying torchvision's coordinate trick when 4*n <= trick_max (ops/boxes.py:80-104).
__global__ void __launch_bounds__(SORT_THREADS)
nms_sort_kernel(const float4* __restrict__ cand_box, const float* __restrict__ cand_score,
                const int* __restrict__ cand_label, const int* __restrict__ cand_count, int cap, int trick_max,
                NmsWs ws) {
  __shared__ int hist[256 * SORT_WARPS];
  __shared__ int warp_tot[SORT_WARPS];
  __shared__ float red[SORT_WARPS];
  __shared__ int skip_pass;
  const int b = blockIdx.x;
  int n = cand_count[b];
  if (n > cap) n = cap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t off = (size_t)b * cap;
  uint32_t* keys[2] = {ws.keys[0] + off, ws.keys[1] + off};
  uint32_t* idx[2] = {ws.idx[0] + off, ws.idx[1] + off};
  for (int i = tid; i < n; i += SORT_THREADS) {
    keys[0][i] = desc_key(cand_score[off + i]);
    idx[0][i] = i;
  }
  __syncthreads();
  int cur = 0;
  if (n <= SORT_THREADS) {
    // At most one candidate per thread (what a detector normally produces: tens per frame): rank by counting -- the number
    // of candidates with a smaller key, or the same key and a smaller index, IS the position in the stable order -- instead
    // of four radix passes with their 8192-entry histograms.  Keys are broadcast from shared memory.
    uint32_t* skey = reinterpret_cast<uint32_t*>(hist);
    const uint32_t mine = tid < n ? keys[0][tid] : 0xffffffffu;
    if (tid < n) skey[tid] = mine;
    __syncthreads();
    if (tid < n) {
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const uint32_t k = skey[j];
        rank += (k < mine) || (k == mine && j < tid);
      }
      idx[1][rank] = tid;
    }
    __syncthreads();
    cur = 1;
  } else {
  // contiguous, 32-aligned segment per warp keeps the scatter stable
  const int per_warp = ((n + SORT_WARPS - 1) / SORT_WARPS + 31) & ~31;
  const int seg0 = min(warp * per_warp, n), seg1 = min(seg0 + per_warp, n);
  // Class-major mode (per-class branch of batched_nms, ops/boxes.py:107-121, on a long list): a FIFTH stable pass on the label
  // leaves the order (label asc, score desc, index asc).  Greedy suppression only acts inside a class and the order inside a
  // class is unchanged, so the kept SET is the same; the mask kernel can skip every tile whose row and column blocks hold
  // different classes (2/3 of the upper triangle for three balanced classes), and the scan kernel restores the score order of
  // the kept list from the score ranks saved in keys[0].
  const int n_pass = class_major(n, trick_max) ? 5 : 4;
  for (int pass = 0; pass < n_pass; ++pass) {
    const int shift = pass * 8;
    const bool by_label = pass == 4;
    for (int i = tid; i < 256 * SORT_WARPS; i += SORT_THREADS) hist[i] = 0;
    if (tid == 0) skip_pass = 0;
    __syncthreads();
    for (int base = seg0; base < seg1; base += 32) {
      const int i = base + lane;
      const bool act = i < seg1;
      const unsigned amask = __ballot_sync(0xffffffffu, act);
      if (act) {
        const int d = by_label ? (cand_label[off + idx[cur][i]] & 255) : (int)((keys[cur][i] >> shift) & 255);
        const unsigned peers = __match_any_sync(amask, d);
        if (lane == __ffs(peers) - 1) hist[d * SORT_WARPS + warp] += __popc(peers);
      }
      __syncwarp();
    }
    __syncthreads();
    // exclusive scan over (digit major, warp minor); 8 consecutive entries per thread
    int v[8], tsum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = hist[tid * 8 + j]; tsum += v[j]; }
    // a digit owning every element means this pass is the identity
    {
      const int dsum_part = tsum;   // entries tid*8..+7 belong to digit (tid*8)/32 = tid/4
      int dsum = dsum_part;
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
      if (dsum == n && n > 0) skip_pass = 1;
    }
    int inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (skip_pass) {
      if (by_label)                                 // one class only: the order stays the score order; ranks all the same
        for (int i = tid; i < n; i += SORT_THREADS) keys[0][idx[cur][i]] = (uint32_t)i;
      __syncthreads();
      continue;
    }
    int wbase = 0;
    for (int i = 0; i < warp; ++i) wbase += warp_tot[i];
    int run = wbase + inc - tsum;
#pragma unroll
    for (int j = 0; j < 8; ++j) { hist[tid * 8 + j] = run; run += v[j]; }
    __syncthreads();
    for (int base = seg0; base < seg1; base += 32) {
      const int i = base + lane;
      const bool act = i < seg1;
      const unsigned amask = __ballot_sync(0xffffffffu, act);
      if (act) {
        const uint32_t k = by_label ? 0u : keys[cur][i];
        const uint32_t ci = idx[cur][i];
        const int d = by_label ? (cand_label[off + ci] & 255) : (int)((k >> shift) & 255);
        const unsigned peers = __match_any_sync(amask, d);
        const int pos = hist[d * SORT_WARPS + warp] + __popc(peers & ((1u << lane) - 1u));
        if (!by_label) keys[cur ^ 1][pos] = k;
        else keys[0][ci] = (uint32_t)i;             // score rank of the candidate (position in the stable score order)
        idx[cur ^ 1][pos] = ci;
        __syncwarp(amask);
        if (lane == __ffs(peers) - 1) hist[d * SORT_WARPS + warp] += __popc(peers);
      }
      __syncwarp();
    }
    cur ^= 1;
    __syncthreads();
  }
  }
  // sorted candidate indices live in idx[cur]; publish them in idx[0] for the scan kernel
  if (cur != 0) {
    for (int i = tid; i < n; i += SORT_THREADS) idx[0][i] = idx[1][i];
  }
  // coordinate trick (needs the max coordinate over all candidates of this image)
  const bool trick = (4 * n <= trick_max);
  float off_scale = 0.f;
  if (trick && n > 0) {
    float m = -INFINITY;
    for (int i = tid; i < n; i += SORT_THREADS) {
      const float4 bx = cand_box[off + i];
      m = fmaxf(m, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < SORT_WARPS; ++i) m = fmaxf(m, red[i]);
    off_scale = __fadd_rn(m, 1.0f);
  }
  __syncthreads();
  for (int i = tid; i < n; i += SORT_THREADS) {
    const int ci = idx[cur][i];
    float4 bx = cand_box[off + ci];
    int lab = cand_label[off + ci];
    if (trick) {
      const float o = __fmul_rn((float)lab, off_scale);
      bx.x = __fadd_rn(bx.x, o); bx.y = __fadd_rn(bx.y, o); bx.z = __fadd_rn(bx.z, o); bx.w = __fadd_rn(bx.w, o);
      lab = 0;
    }
    ws.sbox[off + i] = bx;
    ws.slabel[off + i] = lab;
  }
  __syncthreads();
  if (tid == 0) keys[1][0] = 0u;       // arrival counter of the scan kernel's segment CTAs (keys[1] is free from here on)
}

// 64x64 tiles of the upper triangle: bit j of mask[i][cb] says "sorted box i suppresses sorted box cb*64+j".
// Arithmetic follows torchvision/csrc/ops/cpu/nms_kernel.cpp in fp32 with no contraction.
__global__ void __launch_bounds__(64)
nms_mask_kernel(const int* __restrict__ cand_count, int batch, int cap, float thr_ge, int trick_max, NmsWs ws) {
  __shared__ float4 cb_box[64];
  __shared__ int cb_lab[64];
  __shared__ float cb_area[64];
  // persistent over the list of (image, row block, col block >= row block)
  long long tile = blockIdx.x;
  int b = 0;
  long long first = 0;
  while (true) {
    // advance to the image that owns `tile`
    long long tiles_b = 0;
    int n = 0;
    for (; b < batch; ++b) {
      n = min(cand_count[b], cap);
      const long long nb = (n + 63) / 64;
      tiles_b = nb * (nb + 1) / 2;
      if (tile < first + tiles_b) break;
      first += tiles_b;
    }
    if (b >= batch) return;
    const int nb = (n + 63) / 64;
    // decode the triangular index: row block rb, column block cb >= rb
    long long t = tile - first;
    int rb = 0;
    {
      // rows have nb, nb-1, ... tiles; solve by a short loop bounded by nb (<= 331) using a closed-form start
      // (fp32: both squares are integers below 2^24 for nb <= 2047, so the radicand is exact; the two loops absorb
      // the rounding of the root.  The fp64 version of this line cost more than the 64 pair tests of the tile.)
      const float a = 2.0f * (float)nb + 1.0f;
      const float x = (a - sqrtf(fmaxf(a * a - 8.0f * (float)t, 0.f))) * 0.5f;
      rb = (int)x;
      if (rb < 0) rb = 0;
      if (rb > nb - 1) rb = nb - 1;
      while (rb > 0 && (long long)rb * nb - (long long)rb * (rb - 1) / 2 > t) --rb;
      while ((long long)(rb + 1) * nb - (long long)(rb + 1) * rb / 2 <= t) ++rb;
    }
    const int cb = rb + (int)(t - ((long long)rb * nb - (long long)rb * (rb - 1) / 2));
    const size_t off = (size_t)b * cap;
    const int j0 = cb * 64;
    if (cb > rb && class_major(n, trick_max) && ws.slabel[off + rb * 64 + 63] < ws.slabel[off + j0]) {
      // class-major order (labels ascending): every row of this block belongs to an earlier class than every column
      const int i = rb * 64 + threadIdx.x;
      if (i < n) ws.mask[((size_t)b * cap + i) * ws.words + cb] = 0ull;
      tile += gridDim.x;
      continue;
    }
    __syncthreads();
    if (j0 + (int)threadIdx.x < n) {
      const float4 bx = ws.sbox[off + j0 + threadIdx.x];
      cb_box[threadIdx.x] = bx;
      cb_lab[threadIdx.x] = ws.slabel[off + j0 + threadIdx.x];
      cb_area[threadIdx.x] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
    } else {
      cb_lab[threadIdx.x] = -0x40000000;           // columns beyond n: a label no box has
    }
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i < n) {
      const float4 bi = ws.sbox[off + i];
      const int li = ws.slabel[off + i];
      const float iarea = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
      unsigned long long bits = 0ull;
      // Uniform, branch-free loop over all 64 columns (lanes hold different rows, so a per-lane `continue` on the label
      // or a per-lane start column only diverges); the diagonal tile clears the bits j <= own column afterwards.
      // The decision is fl(inter / uni) >= thr_ge, exactly as torchvision divides and compares.  The IEEE division
      // (~15 instructions) is only executed inside a band of +-1e-4 (relative) around the threshold: outside it the
      // exact quotient is at least 1e-4 away from thr_ge, far more than the 6e-8 the rounding can move it.  (Only for
      // unions and thresholds well inside the normal range, so that the products neither underflow nor overflow;
      // everything else -- empty or inverted boxes, NaN -- takes the division.)
      const bool thr_ok = thr_ge > 1e-6f;
#pragma unroll 1
      for (int jb = 0; jb < 64; jb += 8) {
        uint32_t g = 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = jb + u;
          const float4 bj = cb_box[j];
          const float jarea = cb_area[j];
          const bool same = cb_lab[j] == li;
          const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
          const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
          const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
          const float inter = __fmul_rn(w, h);
          const float uni = __fsub_rn(__fadd_rn(iarea, jarea), inter);
          const float q = __fmul_rn(thr_ge, uni);
          const bool band_ok = thr_ok && uni > 1e-30f && uni < 1e30f;
          const bool sure_miss = band_ok && inter < __fmul_rn(q, 0.9999f);
          const bool sure_hit = band_ok && inter > __fmul_rn(q, 1.0001f);
          bool hit = same && sure_hit;
          if (same && !sure_miss && !sure_hit)
            hit = __fdiv_rn(inter, uni) >= thr_ge;    // == ((double)ovr > (double)iou_threshold); NaN (0/0) compares false
          g |= hit ? (1u << u) : 0u;
        }
        bits |= (unsigned long long)g << jb;
      }
      if (cb == rb) bits &= ~((2ull << threadIdx.x) - 1ull);     // upper triangle only: j > own column
      ws.mask[((size_t)b * cap + i) * ws.words + cb] = bits;
    }
    tile += gridDim.x;
  }
}

// One CTA per image: greedy scan over the bitmask in 64-box chunks, software-pipelined so that the only serial work
// per chunk is the 64-step resolve itself.  Iteration c, phase 1 (all concurrent):
//   thread 0      resolves chunk c against its diagonal words (shared memory / registers only) -> kept bits;
//   warps 1-2     fetch, for every row of chunk c, its candidate index and its mask word c+1 (the only word the NEXT
//                 resolve needs) -- before it is known which rows are kept;
//   warps 3-4     fetch the diagonal words of chunk c+1;
//   warps 5-31    OR the mask rows of the boxes kept in chunk c-1 into removed[c+1 ..] (one warp per row, lanes
//                 along the words: coalesced, all loads of a warp in flight together, 32-bit shared atomics).
// Phase 2: warps 1-2 write the kept indices (rank = popcount of the kept bits below) and OR the prefetched words of
// the kept rows into removed[c+1].  No global-memory latency is exposed on the c -> c+1 dependency.
// (First version: one dependent global load per kept row per thread and the keep list written from the serial loop,
// 19 us per chunk at 10 k candidates; this one ~1.5.)
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_OR_WARPS = SCAN_THREADS / 32 - 5;
// Class-major mode: the classes are independent chains, so grid.y CTAs scan one label segment each (labels 0, 1, 2 and >= 3:
// the last CTA takes whatever is left, which the chain handles like any mixed list) and the CTA that finishes last restores the
// score order of the image's kept list.  A chunk that straddles two segments is resolved by both CTAs, each counting the other
// segment's rows as removed (rows of different labels never suppress each other).  seg_info (in keys[1], free after the sort):
// [0] arrival counter (zeroed by the sort kernel), [1 + 2s] first row of segment s, [2 + 2s] its kept count; the segment's kept
// candidates are parked in idx[1] from its first row on.
constexpr int SCAN_SEGS = 4;

__device__ __forceinline__ int first_row_with_label_ge(const int* __restrict__ slabel, int n, int label) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (slabel[mid] < label) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(SCAN_THREADS)
nms_scan_kernel(const int* __restrict__ cand_count, int cap, int trick_max, NmsWs ws, int* __restrict__ keep,
                int* __restrict__ keep_count) {
  extern __shared__ unsigned long long removed[];   // [words], then [words] ints (class-major mode: word prefix counts)
  __shared__ unsigned long long diag[2][64];
  __shared__ unsigned long long kept_bits_s;
  __shared__ int seg_rows[2];
  __shared__ int last_s;
  const int b = blockIdx.x, seg = blockIdx.y;
  const int n = min(cand_count[b], cap);
  const size_t off = (size_t)b * cap;
  const int words = ws.words;
  const uint32_t* order = ws.idx[0] + off;
  const unsigned long long* mask = ws.mask + off * words;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* rem32 = reinterpret_cast<uint32_t*>(removed);
  const bool cm = class_major(n, trick_max);
  if (!cm && seg != 0) return;
  int row_lo = 0, row_hi = n;
  if (cm) {
    if (tid == 0) {
      const int* sl = ws.slabel + off;
      seg_rows[0] = seg == 0 ? 0 : first_row_with_label_ge(sl, n, seg);
      seg_rows[1] = seg == SCAN_SEGS - 1 ? n : first_row_with_label_ge(sl, n, seg + 1);
    }
    __syncthreads();
    row_lo = seg_rows[0];
    row_hi = seg_rows[1];
  }
  uint32_t* seg_info = ws.keys[1] + off;
  int* parked = reinterpret_cast<int*>(ws.idx[1] + off);
  int* out = cm ? parked + row_lo : keep + off;
  const int c0 = row_lo >> 6, c_end = row_hi > row_lo ? (row_hi + 63) >> 6 : c0;
  for (int i = c0 + tid; i < c_end; i += SCAN_THREADS) removed[i] = 0ull;
  if (tid < 64 && c0 < c_end && c0 * 64 + tid < n) diag[c0 & 1][tid] = mask[(size_t)(c0 * 64 + tid) * words + c0];
  int kept_total = 0;                                // uniform: every thread adds the popcount of each chunk
  unsigned long long kept_prev = 0ull;               // kept bits of chunk c-1 (uniform)
  __syncthreads();
  for (int c = c0; c < c_end; ++c) {
    const int cnt = min(64, row_hi - c * 64);
    unsigned long long next_word = 0ull;
    uint32_t ord = 0u;
    if (tid == 0) {
      const unsigned long long* dg = diag[c & 1];
      unsigned long long cur = removed[c], kept = 0ull;
      if (cnt < 64) cur |= ~0ull << cnt;             // rows beyond the segment (or beyond n) count as removed
      if (row_lo > c * 64) cur |= (1ull << (row_lo - c * 64)) - 1ull;   // and so do the rows of the segment before (first chunk)
#pragma unroll 1
      for (int r0 = 0; r0 < 64; r0 += 8) {
        unsigned long long d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = dg[r0 + i];     // (rows >= cnt hold stale words: never selected)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool alive = !((cur >> (r0 + i)) & 1ull);
          kept |= alive ? (1ull << (r0 + i)) : 0ull;
          cur |= alive ? d[i] : 0ull;
        }
      }
      kept_bits_s = kept;
    } else if (warp == 1 || warp == 2) {
      const int r = tid - 32;
      if (r < cnt) {
        ord = order[c * 64 + r];
        if (c + 1 < c_end) next_word = mask[(size_t)(c * 64 + r) * words + (c + 1)];
      }
    } else if (warp == 3 || warp == 4) {
      const int r = tid - 96, i = (c + 1) * 64 + r;
      if (c + 1 < c_end && i < n) diag[(c + 1) & 1][r] = mask[(size_t)i * words + (c + 1)];
    } else if (warp >= 5 && kept_prev != 0ull && c + 1 < c_end) {
      // warp w owns the rows w-5, w-5+27, w-5+54 of the chunk (no enumeration of the kept bits: ncu showed the
      // kernel issue-bound on exactly that loop, run by all 27 warps)
      for (int r = warp - 5; r < 64; r += SCAN_OR_WARPS) {
        if (!((kept_prev >> r) & 1ull)) continue;
        const unsigned long long* row = mask + (size_t)((c - 1) * 64 + r) * words;
        for (int w0 = c + 1 + lane; w0 < c_end; w0 += 256) {
          unsigned long long v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = (w0 + 32 * i < c_end) ? row[w0 + 32 * i] : 0ull;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t lo = (uint32_t)v[i], hi = (uint32_t)(v[i] >> 32);
            if (lo) atomicOr(rem32 + 2 * (w0 + 32 * i), lo);
            if (hi) atomicOr(rem32 + 2 * (w0 + 32 * i) + 1, hi);
          }
        }
      }
    }
    __syncthreads();
    const unsigned long long kept = kept_bits_s;
    if (warp == 1 || warp == 2) {
      const int r = tid - 32;
      const bool mine = r < cnt && ((kept >> r) & 1ull);
      if (mine) out[kept_total + __popcll(kept & ((1ull << r) - 1ull))] = (int)ord;
      const unsigned long long v = mine ? next_word : 0ull;
      const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)v);
      const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(v >> 32));
      if (lane == 0 && c + 1 < c_end) {
        if (lo) atomicOr(rem32 + 2 * (c + 1), lo);
        if (hi) atomicOr(rem32 + 2 * (c + 1) + 1, hi);
      }
    }
    kept_total += __popcll(kept);
    kept_prev = kept;
    __syncthreads();
  }
  if (!cm) {
    if (tid == 0) keep_count[b] = kept_total;
    return;
  }
  // ---- class-major mode: publish this segment, and let the CTA that arrives last restore the score order
  if (tid == 0) {
    seg_info[1 + 2 * seg] = (uint32_t)row_lo;
    seg_info[2 + 2 * seg] = (uint32_t)kept_total;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned old = atomicAdd(seg_info, 1u);
    last_s = old == (unsigned)SCAN_SEGS - 1u;
    if (last_s) seg_info[0] = 0u;                   // everyone has arrived: reset for the next launch
  }
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  // The kept candidates are in (label, score) order; batched_nms returns them by descending score (ops/boxes.py:120-121; ties:
  // the stable order, ascending candidate index, as everywhere in this file).  Every candidate's rank in the stable score order
  // was saved by the sort kernel: set bit `rank` for every kept candidate, count the bits below -> its output position.
  const uint32_t* rank = ws.keys[0] + off;
  const int nb = (n + 63) / 64;
  int* wprefix = reinterpret_cast<int*>(removed + words);
  int lo_s[SCAN_SEGS], kept_s[SCAN_SEGS], total = 0;
#pragma unroll
  for (int s_ = 0; s_ < SCAN_SEGS; ++s_) {
    lo_s[s_] = (int)__ldcg(seg_info + 1 + 2 * s_);
    kept_s[s_] = (int)__ldcg(seg_info + 2 + 2 * s_);
    total += kept_s[s_];
  }
  for (int i = tid; i < nb; i += SCAN_THREADS) removed[i] = 0ull;
  __syncthreads();
#pragma unroll
  for (int s_ = 0; s_ < SCAN_SEGS; ++s_)
    for (int k = tid; k < kept_s[s_]; k += SCAN_THREADS) {
      const uint32_t r = rank[__ldcg(parked + lo_s[s_] + k)];
      atomicOr(rem32 + (r >> 5), 1u << (r & 31));
    }
  __syncthreads();
  if (warp == 0) {                                                // exclusive prefix of the per-word popcounts
    int run = 0;
    for (int w0 = 0; w0 < nb; w0 += 32) {
      const int w = w0 + lane;
      const int cw = w < nb ? __popcll(removed[w]) : 0;
      int inc = cw;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (w < nb) wprefix[w] = run + inc - cw;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
#pragma unroll
  for (int s_ = 0; s_ < SCAN_SEGS; ++s_)
    for (int k = tid; k < kept_s[s_]; k += SCAN_THREADS) {
      const int cand = __ldcg(parked + lo_s[s_] + k);
      const uint32_t r = rank[cand];
      keep[off + wprefix[r >> 6] + __popcll(removed[r >> 6] & ((1ull << (r & 63)) - 1ull))] = cand;
    }
  if (tid == 0) keep_count[b] = total;
}

// ---------------------------------------------------------------------------------- gather
struct GatherParams {
  int num_levels;
  int level_start[MAX_LEVELS + 1];
};
constexpr int GATHER_MAX_BATCH = 64;
struct Ratios {
  float rh[GATHER_MAX_BATCH], rw[GATHER_MAX_BATCH];
};

__global__ void __launch_bounds__(256)
gather_kernel(const int* __restrict__ keep, const int* __restrict__ keep_count, const int* __restrict__ cand_loc,
              const float* __restrict__ cand_score, const int* __restrict__ cand_label,
              const float4* __restrict__ cand_box, const HeadView hand_lr, const HeadView contact_logits, const HeadView dxdy,
              int cap, int locs,
              int batch_offset, GatherParams gp, Ratios rt, float4* __restrict__ boxes, float* __restrict__ scores,
              long long* __restrict__ labels, long long* __restrict__ sides, float* __restrict__ level,
              long long* __restrict__ contacts, float* __restrict__ dxdymags) {
  const int bl = blockIdx.y;
  const int b = batch_offset + bl;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= min(keep_count[b], cap)) return;
  const size_t off = (size_t)b * cap;
  const int ci = keep[off + k];
  const int loc = cand_loc[off + ci];
  float4 bx = cand_box[off + ci];
  // resize_boxes (fcos_utils/fcos.py:770-783): x * ratio_w, y * ratio_h in fp32
  bx.x = __fmul_rn(bx.x, rt.rw[bl]); bx.z = __fmul_rn(bx.z, rt.rw[bl]);
  bx.y = __fmul_rn(bx.y, rt.rh[bl]); bx.w = __fmul_rn(bx.w, rt.rh[bl]);
  boxes[off + k] = bx;
  scores[off + k] = cand_score[off + ci];
  labels[off + k] = cand_label[off + ci];
  const float* lr = hand_lr.p + (size_t)b * hand_lr.img + (size_t)loc * hand_lr.loc;
  sides[off + k] = (sigmoid_rn(__ldg(lr + hand_lr.chan)) > sigmoid_rn(__ldg(lr))) ? 1 : 0;
  int l = 0;
  while (l + 1 < gp.num_levels && loc >= gp.level_start[l + 1]) ++l;
  level[off + k] = (float)l;
  if (contacts) {
    const float* cl = contact_logits.p + (size_t)b * contact_logits.img + (size_t)loc * contact_logits.loc;
    float best = sigmoid_rn(__ldg(cl));
    int bi = 0;
    for (int c = 1; c < 5; ++c) {
      const float s = sigmoid_rn(__ldg(cl + (size_t)c * contact_logits.chan));
      if (s > best) { best = s; bi = c; }
    }
    contacts[off + k] = bi;
  }
  if (dxdymags) {
    // (d0, 0.1 * normalize((d1, d2), p=2, eps=1e-12))   fcos_utils/fcos.py:299-303
    const float* dp = dxdy.p + (size_t)b * dxdy.img + (size_t)loc * dxdy.loc;
    float* o = dxdymags + (off + k) * 3;
    const float d1 = __ldg(dp + dxdy.chan), d2 = __ldg(dp + 2 * (size_t)dxdy.chan);
    const float nrm = fmaxf(__fsqrt_rn(__fadd_rn(__fmul_rn(d1, d1), __fmul_rn(d2, d2))), 1e-12f);
    o[0] = __ldg(dp);
    o[1] = __fmul_rn(0.1f, __fdiv_rn(d1, nrm));
    o[2] = __fmul_rn(0.1f, __fdiv_rn(d2, nrm));
  }
}

}  // namespace

extern "C" int64_t hn_fcos_select_workspace_bytes(int batch, int locs) {
  return (int64_t)align256((size_t)batch * hn_div_up(locs, SEL_CHUNK) * 4);      // one state word per chunk (of the finest split)
}

extern "C" int hn_fcos_decode_select(const float* cls_logits, int64_t cls_img_stride, int cls_loc_stride, int cls_chan_stride,
                                     const float* bbox_ctrness, int64_t ctr_img_stride, int ctr_loc_stride,
                                     const float* bbox_regression, int64_t reg_img_stride, int reg_loc_stride,
                                     int reg_chan_stride, int batch, int locs, int num_classes, int num_levels,
                                     const int* level_h_host, const int* level_w_host, const int* level_stride_h_host,
                                     const int* level_stride_w_host, const int* level_anchor_host, double score_thresh,
                                     int* cand_count, int* cand_loc, float* cand_score, int* cand_label,
                                     float* cand_box, void* workspace, int64_t workspace_bytes, void* stream) {
  HN_REQUIRE(cls_logits && bbox_ctrness && bbox_regression && cand_count && cand_loc && cand_score && cand_label &&
                 cand_box && workspace, "hn_fcos_decode_select: null pointer");
  HN_REQUIRE(batch > 0 && locs > 0 && num_classes > 0 && num_levels > 0 && num_levels <= MAX_LEVELS,
             "hn_fcos_decode_select: bad sizes");
  HN_REQUIRE(cls_loc_stride >= 1 && cls_chan_stride >= 1 && ctr_loc_stride >= 1 && reg_loc_stride >= 1 && reg_chan_stride >= 1 &&
                 (reinterpret_cast<uintptr_t>(cand_box) & 15) == 0,
             "hn_fcos_decode_select: strides must be positive and cand_box 16-byte aligned");
  HN_REQUIRE(workspace_bytes >= hn_fcos_select_workspace_bytes(batch, locs), "hn_fcos_decode_select: workspace too small");
  Levels lv;
  memset(&lv, 0, sizeof(lv));
  lv.n = num_levels;
  int acc = 0;
  for (int l = 0; l < num_levels; ++l) {
    lv.start[l] = acc;
    lv.h[l] = level_h_host[l];
    lv.w[l] = level_w_host[l];
    lv.sh[l] = level_stride_h_host[l];
    lv.sw[l] = level_stride_w_host[l];
    lv.anchor[l] = level_anchor_host[l];
    acc += lv.h[l] * lv.w[l];
  }
  lv.start[num_levels] = acc;
  HN_REQUIRE(acc == locs, "hn_fcos_decode_select: levels cover %d locations, locs=%d", acc, locs);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HN_REQUIRE(num_classes <= 256 && locs < (1 << 24), "hn_fcos_decode_select: at most 256 classes and 2^24 locations per image");
  // Rounds per block.  Measured on one B200 (256 frames, 17 850 locations, sparse / stress regime, tools/membound_roofline.py):
  // 1 round 35.8 / 66.7 us, 2 rounds 35.6 / 64.5, 3 rounds 37.9 / 74.5, 4 rounds 39.6 / 80.9 -- two rounds halve the block
  // tails (scan, look-back, exit) without emptying the last wave of resident blocks; small grids take one round.
  const int stage_planes = num_classes <= 4 ? num_classes + 1 : 2;
  const int rounds = (long long)batch * hn_div_up(locs, SEL_CHUNK * 2) >= 2 * 5 * 148 ? 2 : 1;
  const int chunks = hn_div_up(locs, SEL_CHUNK * rounds);
  HN_REQUIRE((long long)batch * chunks < (1ll << 31), "hn_fcos_decode_select: grid too large");
  const float thresh_ge = float_gt_as_ge(score_thresh);
  // score >= t needs sigmoid(cls) * sigmoid(ctr) >= t^2, hence each logit >= logit(t^2); 0.01 of margin covers the rounding
  // of the fp32 evaluation many times over (at t = 0.7: skip below -0.05, where the score can reach 0.698 at most)
  float skip_below = -INFINITY;
  if (score_thresh > 0.0 && score_thresh < 1.0) {
    const double t2 = score_thresh * score_thresh;
    skip_below = (float)(log(t2 / (1.0 - t2)) - 0.01);
  }
  const HeadView vcls = {cls_logits, (long long)cls_img_stride, cls_loc_stride, cls_chan_stride};
  const HeadView vctr = {bbox_ctrness, (long long)ctr_img_stride, ctr_loc_stride, 1};
  const HeadView vreg = {bbox_regression, (long long)reg_img_stride, reg_loc_stride, reg_chan_stride};
  unsigned* state = reinterpret_cast<unsigned*>(workspace);
  HN_CHECK_CUDA(cudaMemsetAsync(state, 0, (size_t)batch * chunks * sizeof(unsigned), st));
  float4* box4 = reinterpret_cast<float4*>(cand_box);
  const size_t parked_bytes = (size_t)SEL_WARPS * rounds * stage_planes * SEL_ROUND * sizeof(float);
#define HN_SELECT_LAUNCH(NC)                                                                                                 \
  {                                                                                                                            \
    static bool attr = false;                                                                                                  \
    if (!attr) {                                                                                                               \
      HN_CHECK_CUDA(cudaFuncSetAttribute(select_decode_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));   \
      attr = true;                                                                                                             \
    }                                                                                                                          \
  }                                                                                                                            \
  select_decode_kernel<NC><<<batch * chunks, SEL_BLOCK, parked_bytes, st>>>(vcls, vctr, vreg, locs, num_classes, chunks, rounds, \
                                                                            thresh_ge, skip_below, lv, state, cand_count,    \
                                                                            cand_loc, cand_score, cand_label, box4)
  switch (num_classes) {
    case 1: HN_SELECT_LAUNCH(1); break;
    case 2: HN_SELECT_LAUNCH(2); break;
    case 3: HN_SELECT_LAUNCH(3); break;
    case 4: HN_SELECT_LAUNCH(4); break;
    default: HN_SELECT_LAUNCH(0); break;
  }
#undef HN_SELECT_LAUNCH
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int64_t hn_nms_workspace_bytes(int batch, int cap) {
  const size_t n = (size_t)batch * cap;
  const size_t words = (cap + 63) / 64;
  return (int64_t)(align256(n * 4) * 4 + align256(n * 16) + align256(n * 4) + align256(n * words * 8));
}

extern "C" int hn_nms_batched(const float* cand_box, const float* cand_score, const int* cand_label,
                              const int* cand_count, int batch, int cap, double iou_thresh, int coord_trick_max_numel,
                              int* keep, int* keep_count, void* workspace, int64_t workspace_bytes, void* stream) {
  HN_REQUIRE(cand_box && cand_score && cand_label && cand_count && keep && keep_count && workspace,
             "hn_nms_batched: null pointer");
  HN_REQUIRE(batch > 0 && cap > 0, "hn_nms_batched: bad sizes");
  HN_REQUIRE(workspace_bytes >= hn_nms_workspace_bytes(batch, cap), "hn_nms_batched: workspace too small");
  NmsWs ws = carve(workspace, batch, cap);
  HN_REQUIRE((size_t)ws.words * 12 <= 200 * 1024, "hn_nms_batched: cap too large for the scan kernel");
  const float thr_ge = float_gt_as_ge(iou_thresh);   // (double)iou > thr, as torchvision's CPU kernel compares
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  nms_sort_kernel<<<batch, SORT_THREADS, 0, st>>>(reinterpret_cast<const float4*>(cand_box), cand_score, cand_label,
                                                  cand_count, cap, coord_trick_max_numel, ws);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  nms_mask_kernel<<<hn_num_sms() * 32, 64, 0, st>>>(cand_count, batch, cap, thr_ge, coord_trick_max_numel, ws);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  const size_t scan_smem = (size_t)ws.words * 12;     // removed[] + the word prefix counts of the class-major reorder
  static bool attr = false;
  if (!attr && scan_smem > 40 * 1024) {
    HN_CHECK_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  nms_scan_kernel<<<dim3(batch, SCAN_SEGS), SCAN_THREADS, scan_smem, st>>>(cand_count, cap, coord_trick_max_numel, ws, keep,
                                                                            keep_count);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int hn_fcos_gather(const int* keep, const int* keep_count, const int* cand_loc, const float* cand_score,
                              const int* cand_label, const float* cand_box, const float* hand_lr, int64_t lr_img_stride,
                              int lr_loc_stride, int lr_chan_stride, const float* contact_logits, int64_t contact_img_stride,
                              int contact_loc_stride, int contact_chan_stride, const float* dxdy, int64_t dxdy_img_stride,
                              int dxdy_loc_stride, int dxdy_chan_stride, int batch, int cap, int locs,
                              int num_levels, const int* level_start_host, const float* ratio_h_host,
                              const float* ratio_w_host, float* boxes, float* scores, int64_t* labels, int64_t* sides,
                              float* level, int64_t* contacts, float* dxdymags, void* stream) {
  HN_REQUIRE(keep && keep_count && cand_loc && cand_score && cand_label && cand_box && hand_lr && boxes && scores &&
                 labels && sides && level, "hn_fcos_gather: null pointer");
  HN_REQUIRE((contacts == nullptr) == (contact_logits == nullptr) && (dxdymags == nullptr) == (dxdy == nullptr),
             "hn_fcos_gather: ext outputs need their inputs");
  HN_REQUIRE(batch > 0 && cap > 0 && num_levels > 0 && num_levels <= MAX_LEVELS, "hn_fcos_gather: bad sizes");
  GatherParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.num_levels = num_levels;
  for (int l = 0; l <= num_levels; ++l) gp.level_start[l] = level_start_host[l];
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int b0 = 0; b0 < batch; b0 += GATHER_MAX_BATCH) {
    const int nb = (batch - b0 < GATHER_MAX_BATCH) ? batch - b0 : GATHER_MAX_BATCH;
    Ratios rt;
    for (int i = 0; i < nb; ++i) { rt.rh[i] = ratio_h_host[b0 + i]; rt.rw[i] = ratio_w_host[b0 + i]; }
    dim3 grid(hn_div_up(cap, 256), nb);
    gather_kernel<<<grid, 256, 0, st>>>(keep, keep_count, cand_loc, cand_score, cand_label,
                                        reinterpret_cast<const float4*>(cand_box),
                                        HeadView{hand_lr, (long long)lr_img_stride, lr_loc_stride, lr_chan_stride},
                                        HeadView{contact_logits, (long long)contact_img_stride, contact_loc_stride, contact_chan_stride},
                                        HeadView{dxdy, (long long)dxdy_img_stride, dxdy_loc_stride, dxdy_chan_stride},
                                        cap, locs, b0, gp, rt, reinterpret_cast<float4*>(boxes), scores,
                                        reinterpret_cast<long long*>(labels), reinterpret_cast<long long*>(sides), level,
                                        reinterpret_cast<long long*>(contacts), dxdymags);
    hn_count_launch();
    HN_LAUNCH_CHECK();
  }
  return HN_OK;
}
