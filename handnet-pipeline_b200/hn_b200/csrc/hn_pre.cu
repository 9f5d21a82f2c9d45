// Memory-bound glue kernels around the tensor-core convolutions:
//   T1  normalize + bilinear resize + zero-pad into the detector canvas (torchvision transform.py:119-255)
//   7x7/2 stem patches as GEMM rows (feeds hn_conv2d_bf16 for backbone.body.conv1 / Backbone.model.conv1)
//   3x3/2 max pool (resnet maxpool)
//   GroupNorm + ReLU of the FCOS towers (fcos_utils/fcos.py:232-240)
// All are HBM/L2-bound streaming kernels: 16-byte vector accesses, one pass.  (Two of them were instruction-bound until the end of
// round 2 -- per-element index arithmetic and per-block coefficient set-up -- see preprocess_pairs_kernel, groupnorm_relu_kernel.)
#include "hn_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------ T1
constexpr int PRE_MAX_IMAGES = 64;
struct PreParams {
  const float* img[PRE_MAX_IMAGES];
  int in_h[PRE_MAX_IMAGES], in_w[PRE_MAX_IMAGES], out_h[PRE_MAX_IMAGES], out_w[PRE_MAX_IMAGES];
  float mean[3], inv_std[3], stdv[3];             // inv_std = 1 / std (fp32, rounded once on the host)
  float scale_y[PRE_MAX_IMAGES], scale_x[PRE_MAX_IMAGES];   // in / out as fp32, per image
  int canvas_h, canvas_w, batch_offset;
  int pad_top, pad_left, pitch_h, pitch_w;   // the canvas sits at (pad_top, pad_left) of a [pitch_h][pitch_w] pixel frame
  int paired;                                // frame rows stored in pairs: [pitch_h/2][pitch_w][2 rows][4 ch] (direct stem)
};

__device__ __forceinline__ void bilinear_axis(int o, int in, float scale, int& i0, int& i1, float& l0, float& l1) {
  // ATen area_pixel_compute_source_index(align_corners=false): scale = in/out in fp32 (computed once per image on the
  // host, same IEEE division), src clamped at 0
  float src = scale * ((float)o + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// One thread = four consecutive canvas pixels of one row (the row interpolation is shared, one 32-byte store).
__global__ void __launch_bounds__(256) preprocess_kernel(const PreParams p, uint2* __restrict__ canvas) {
  const int b = blockIdx.z;
  const int y = blockIdx.y;
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x4 >= p.canvas_w) return;
  const int oh = p.out_h[b], ow = p.out_w[b];
  uint2 outv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) outv[j] = make_uint2(0u, 0u);
  if (y < oh && x4 < ow) {
    const int ih = p.in_h[b], iw = p.in_w[b];
    int y0, y1;
    float ly0, ly1;
    bilinear_axis(y, ih, p.scale_y[b], y0, y1, ly0, ly1);
    const float* r0 = p.img[b] + (size_t)y0 * iw;
    const float* r1 = p.img[b] + (size_t)y1 * iw;
    const size_t plane = (size_t)ih * iw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x4 + j;
      if (x < ow) {
        int x0, x1;
        float lx0, lx1;
        bilinear_axis(x, iw, p.scale_x[b], x0, x1, lx0, lx1);
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* q0 = r0 + c * plane;
          const float* q1 = r1 + c * plane;
          // the bilinear weights sum to one, so normalising after the interpolation equals torchvision's
          // normalise-then-resize up to fp32 rounding (the result is rounded to bf16 anyway): 1 division instead of 4
          const float t = ly0 * (lx0 * __ldg(q0 + x0) + lx1 * __ldg(q0 + x1)) + ly1 * (lx0 * __ldg(q1 + x0) + lx1 * __ldg(q1 + x1));
          v[c] = (t - p.mean[c]) * p.inv_std[c];   // (an IEEE division is ~20 instructions: the kernel was issue-bound)
        }
        outv[j].x = hn_pack_bf16(v[0], v[1]);
        outv[j].y = hn_pack_bf16(v[2], 0.f);
      }
    }
  }
  uint2* dst = canvas + ((size_t)(p.batch_offset + b) * p.pitch_h + y + p.pad_top) * p.pitch_w + x4 + p.pad_left;
  if (x4 + 4 <= p.canvas_w && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
    const uint32_t w8[8] = {outv[0].x, outv[0].y, outv[1].x, outv[1].y, outv[2].x, outv[2].y, outv[3].x, outv[3].y};
    hn_stg256(dst, w8);
  } else {
    for (int j = 0; j < 4 && x4 + j < p.canvas_w; ++j) dst[j] = outv[j];
  }
}

// Row-staged variant (the default): one CTA = one canvas row of one image.  The two source rows of each colour plane
// are read with coalesced 16-byte loads, blended vertically, and the blended row (3 * in_w floats) is kept in shared
// memory; then consecutive lanes interpolate consecutive canvas pixels horizontally from there (their source columns
// are <= one word apart: conflict-free broadcasts).  preprocess_kernel's 48 scattered 4-byte global loads and three
// IEEE divisions per thread made it issue-bound (ncu: 80 % issue-active at 24 % of the HBM roofline); this one
// spends ~55 instructions per pixel.  The bilinear weights are applied vertical-first here (ATen: horizontal-first):
// the same four products, summed in another order -- a few fp32 ulp before the rounding to bf16.
__global__ void __launch_bounds__(512) preprocess_rows_kernel(const PreParams p, uint2* __restrict__ canvas, int pitch) {
  extern __shared__ __align__(16) float staged[];   // [3 planes][pitch]
  const int b = blockIdx.z;
  const int y = blockIdx.y;
  const int oh = p.out_h[b], ow = p.out_w[b];
  const int ih = p.in_h[b], iw = p.in_w[b];
  const bool live = y < oh;
  if (live) {
    int y0, y1;
    float ly0, ly1;
    bilinear_axis(y, ih, p.scale_y[b], y0, y1, ly0, ly1);
    const float* base = p.img[b];
    const size_t plane = (size_t)ih * iw;
    if ((iw & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
      const int q = iw >> 2;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4* r0 = reinterpret_cast<const float4*>(base + c * plane + (size_t)y0 * iw);
        const float4* r1 = reinterpret_cast<const float4*>(base + c * plane + (size_t)y1 * iw);
        float4* dst = reinterpret_cast<float4*>(staged + c * pitch);
        for (int col = threadIdx.x; col < q; col += blockDim.x) {
          const float4 u = __ldg(r0 + col), v = __ldg(r1 + col);
          dst[col] = make_float4(ly0 * u.x + ly1 * v.x, ly0 * u.y + ly1 * v.y, ly0 * u.z + ly1 * v.z, ly0 * u.w + ly1 * v.w);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* r0 = base + c * plane + (size_t)y0 * iw;
        const float* r1 = base + c * plane + (size_t)y1 * iw;
        for (int col = threadIdx.x; col < iw; col += blockDim.x) staged[c * pitch + col] = ly0 * __ldg(r0 + col) + ly1 * __ldg(r1 + col);
      }
    }
  }
  __syncthreads();
  const float m0 = p.mean[0], m1 = p.mean[1], m2 = p.mean[2];
  const float s0 = p.inv_std[0], s1 = p.inv_std[1], s2 = p.inv_std[2];
  const float scale_x = p.scale_x[b];
  const float* q1 = staged + pitch;
  const float* q2 = staged + 2 * pitch;
  // row-major frame: pixel x of frame row fy at [fy][x]; paired frame (direct stem): at [fy / 2][x][fy & 1]
  const int fy = y + p.pad_top, xs = p.paired ? 2 : 1;
  uint2* dst_row = p.paired ? canvas + (((size_t)(p.batch_offset + b) * (p.pitch_h >> 1) + (fy >> 1)) * p.pitch_w + p.pad_left) * 2 + (fy & 1)
                            : canvas + ((size_t)(p.batch_offset + b) * p.pitch_h + fy) * p.pitch_w + p.pad_left;
  const int x_live = live ? ow : 0;
#pragma unroll 2
  for (int x = threadIdx.x; x < p.canvas_w; x += blockDim.x) {
    uint2 o = make_uint2(0u, 0u);
    if (x < x_live) {
      int x0, x1;
      float lx0, lx1;
      bilinear_axis(x, iw, scale_x, x0, x1, lx0, lx1);
      const float t0 = lx0 * staged[x0] + lx1 * staged[x1];
      const float t1 = lx0 * q1[x0] + lx1 * q1[x1];
      const float t2 = lx0 * q2[x0] + lx1 * q2[x1];
      o.x = hn_pack_bf16((t0 - m0) * s0, (t1 - m1) * s1);
      o.y = hn_pack_bf16((t2 - m2) * s2, 0.f);
    }
    dst_row[x * xs] = o;
  }
}

// Row-pair variant (the default when the source rows are 16-byte aligned): one CTA walks `ppc` consecutive PAIRS of frame rows
// of one image.  Per pair: (1) the four source rows the two canvas rows blend (3 colour planes each) arrive in shared memory
// by cp.async, issued while the previous pair is being interpolated; (2) a vertical blend pass in shared memory that leaves,
// per source column, the six blended values (row a: c0 c1 c2, row b: c0 c1 c2) side by side; (3) consecutive lanes interpolate
// consecutive canvas pixels horizontally for BOTH rows of the pair from 2 x 3 eight-byte reads, in packed fp32x2 arithmetic,
// and, in the paired frame layout of the direct stem ([rows/2][width][2 rows][4 ch]), write the pixel's two rows as one
// 16-byte store (the one-row kernel wrote 8-byte halves of every 16 bytes, twice).  The horizontal source index and weight
// of every canvas column are tabulated once per CTA.  Same arithmetic as preprocess_rows_kernel (vertical blend first, then
// the horizontal one, then (t - mean) * (1 / std)).
constexpr int PRE_T = 256;       // threads at most (the host picks a multiple of 32 that tiles the canvas width tightly)

__device__ __forceinline__ void pre_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(hn_smem_u32(dst)), "l"(src) : "memory");
}

template <bool PAIRED>
__global__ void __launch_bounds__(PRE_T, 4)
preprocess_pairs_kernel(const PreParams p, uint2* __restrict__ canvas, int pitch, int ppc, int first_pair, int npairs) {
  extern __shared__ __align__(16) float sm[];
  float* raw = sm;                    // [2 canvas rows][2 source rows][3 planes][pitch]
  float* bl = sm + 12 * pitch;        // [pitch columns][row a: c0 c1 c2, row b: c0 c1 c2]: vertically blended
  float2* xtab = reinterpret_cast<float2*>(sm + 18 * pitch);   // [canvas_w]: (source column x0, weight of x0 + 1)
  const int b = blockIdx.z;
  const int oh = p.out_h[b], ow = p.out_w[b];
  const int ih = p.in_h[b], iw = p.in_w[b];
  const int q = iw >> 2;
  const int plane = ih * iw;
  const float* base = p.img[b];
  const int tid = threadIdx.x, T = blockDim.x;
  const int pr_begin = first_pair + blockIdx.x * ppc;
  const int pr_end = min(pr_begin + ppc, first_pair + npairs);
  if (pr_begin >= pr_end) return;
  for (int x = tid; x < p.canvas_w; x += T) {
    int x0 = 0, x1;
    float l0, l1 = 0.f;
    if (x < ow) bilinear_axis(x, iw, p.scale_x[b], x0, x1, l0, l1);
    xtab[x] = make_float2(__int_as_float(x0), l1);
  }
  const float2 nm01 = make_float2(-p.mean[0], -p.mean[1]), nm20 = make_float2(-p.mean[2], -p.mean[0]), nm12 = make_float2(-p.mean[1], -p.mean[2]);
  const float2 s01 = make_float2(p.inv_std[0], p.inv_std[1]), s20 = make_float2(p.inv_std[2], p.inv_std[0]), s12 = make_float2(p.inv_std[1], p.inv_std[2]);
  const int iw1 = iw - 1;

  // rows of a pair: liveness, vertical weights; the cp.async copies of its four source rows
  bool live_a, live_b;
  float wa, wb;                       // weight of the lower source row (y1) for canvas row a / b
  auto issue = [&](int pr) {
    int ya0 = 0, ya1 = 0, yb0 = 0, yb1 = 0;
    float l0;
    const int y = 2 * pr - p.pad_top;
    live_a = y >= 0 && y < oh;
    live_b = y + 1 >= 0 && y + 1 < oh;
    wa = wb = 0.f;
    if (live_a) bilinear_axis(y, ih, p.scale_y[b], ya0, ya1, l0, wa);
    if (live_b) bilinear_axis(y + 1, ih, p.scale_y[b], yb0, yb1, l0, wb);
    const float* ra0 = base + (size_t)ya0 * iw;
    const float* ra1 = base + (size_t)ya1 * iw;
    const float* rb0 = base + (size_t)yb0 * iw;
    const float* rb1 = base + (size_t)yb1 * iw;
    for (int it = tid; it < 3 * q; it += T) {
      const int c = (it >= q) + (it >= 2 * q);
      const int col = (it - c * q) * 4;
      const int go = c * plane + col;
      float* d = raw + c * pitch + col;
      if (live_a) {
        pre_cp16(d, ra0 + go);
        pre_cp16(d + 3 * pitch, ra1 + go);
      }
      if (live_b) {
        pre_cp16(d + 6 * pitch, rb0 + go);
        pre_cp16(d + 9 * pitch, rb1 + go);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  issue(pr_begin);
  for (int pr = pr_begin; pr < pr_end; ++pr) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                     // the pair's source rows are in; the previous pair has been written out
    {
      // vertical blend: a thread takes 4 source columns, all planes, both rows -> 24 consecutive floats of bl
      const float2 wa1 = make_float2(wa, wa), wa0 = make_float2(1.f - wa, 1.f - wa);
      const float2 wb1 = make_float2(wb, wb), wb0 = make_float2(1.f - wb, 1.f - wb);
      for (int g = tid; g < q; g += T) {
        float v[2][3][4];                                // [row a / b][plane][column]
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* r = raw + c * pitch + g * 4;
          float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0, u2 = u0, u3 = u0;
          if (live_a) { u0 = *reinterpret_cast<const float4*>(r); u1 = *reinterpret_cast<const float4*>(r + 3 * pitch); }
          if (live_b) { u2 = *reinterpret_cast<const float4*>(r + 6 * pitch); u3 = *reinterpret_cast<const float4*>(r + 9 * pitch); }
          const float2 alo = __ffma2_rn(wa1, make_float2(u1.x, u1.y), __fmul2_rn(wa0, make_float2(u0.x, u0.y)));
          const float2 ahi = __ffma2_rn(wa1, make_float2(u1.z, u1.w), __fmul2_rn(wa0, make_float2(u0.z, u0.w)));
          const float2 blo = __ffma2_rn(wb1, make_float2(u3.x, u3.y), __fmul2_rn(wb0, make_float2(u2.x, u2.y)));
          const float2 bhi = __ffma2_rn(wb1, make_float2(u3.z, u3.w), __fmul2_rn(wb0, make_float2(u2.z, u2.w)));
          v[0][c][0] = alo.x; v[0][c][1] = alo.y; v[0][c][2] = ahi.x; v[0][c][3] = ahi.y;
          v[1][c][0] = blo.x; v[1][c][1] = blo.y; v[1][c][2] = bhi.x; v[1][c][3] = bhi.y;
        }
        float4* d = reinterpret_cast<float4*>(bl + g * 24);
#pragma unroll
        for (int h = 0; h < 2; ++h) {                    // two columns = 12 floats = three 16-byte stores
          const int i = 2 * h;
          d[3 * h + 0] = make_float4(v[0][0][i], v[0][1][i], v[0][2][i], v[1][0][i]);
          d[3 * h + 1] = make_float4(v[1][1][i], v[1][2][i], v[0][0][i + 1], v[0][1][i + 1]);
          d[3 * h + 2] = make_float4(v[0][2][i + 1], v[1][0][i + 1], v[1][1][i + 1], v[1][2][i + 1]);
        }
      }
    }
    __syncthreads();                                     // blended rows complete; the raw rows are free again
    const bool la = live_a, lb = live_b;
    if (pr + 1 < pr_end) issue(pr + 1);
    const int fy = 2 * pr;
    uint4* dst16 = nullptr;
    uint2* dst8 = nullptr;
    bool row_ok0 = true, row_ok1 = true;
    if (PAIRED) {
      dst16 = reinterpret_cast<uint4*>(canvas) + ((size_t)(p.batch_offset + b) * (p.pitch_h >> 1) + pr) * p.pitch_w + p.pad_left;
    } else {
      dst8 = canvas + ((size_t)(p.batch_offset + b) * p.pitch_h + fy) * p.pitch_w + p.pad_left;
      row_ok0 = fy - p.pad_top >= 0 && fy - p.pad_top < p.canvas_h;
      row_ok1 = fy + 1 - p.pad_top >= 0 && fy + 1 - p.pad_top < p.canvas_h;
    }
#pragma unroll 2
    for (int x = tid; x < p.canvas_w; x += T) {
      const float2 xc = xtab[x];
      const float2 l11 = make_float2(xc.y, xc.y), l00 = make_float2(1.f - xc.y, 1.f - xc.y);
      const int i0 = __float_as_int(xc.x);
      const float2* p0 = reinterpret_cast<const float2*>(bl + i0 * 6);
      const float2* p1 = reinterpret_cast<const float2*>(bl + min(i0 + 1, iw1) * 6);
      float2 t01 = __ffma2_rn(l11, p1[0], __fmul2_rn(l00, p0[0]));     // row a: c0, c1
      float2 t20 = __ffma2_rn(l11, p1[1], __fmul2_rn(l00, p0[1]));     // row a: c2; row b: c0
      float2 t12 = __ffma2_rn(l11, p1[2], __fmul2_rn(l00, p0[2]));     // row b: c1, c2
      t01 = __fmul2_rn(__fadd2_rn(t01, nm01), s01);
      t20 = __fmul2_rn(__fadd2_rn(t20, nm20), s20);
      t12 = __fmul2_rn(__fadd2_rn(t12, nm12), s12);
      const bool in = x < ow;
      uint2 oa = make_uint2(0u, 0u), ob = make_uint2(0u, 0u);
      if (in && la) oa = make_uint2(hn_pack_bf16(t01.x, t01.y), hn_pack_bf16(t20.x, 0.f));
      if (in && lb) ob = make_uint2(hn_pack_bf16(t20.y, t12.x), hn_pack_bf16(t12.y, 0.f));
      if (PAIRED) {
        dst16[x] = make_uint4(oa.x, oa.y, ob.x, ob.y);
      } else {
        if (row_ok0) dst8[x] = oa;
        if (row_ok1) dst8[p.pitch_w + x] = ob;
      }
    }
  }
}

// ------------------------------------------------------------------------------------- im2col
// Stem patches as GEMM rows.  K is laid out as 8 kernel rows x 8 pixels x C channels: row r (r < 7) of the 7x7
// window, pixels x = 2*ox - 4 .. 2*ox + 3 (the first one is outside the 7-wide window and meets a zero weight),
// so that every 16-byte output chunk is a contiguous, aligned piece of one input row.  r = 7 is zero padding.
// bf16 canvas (4 channels/pixel): k = j*64 + px*8 + rr*4 + ch with r = 2*j + rr (the direct stem's order: a k-block is
// two kernel rows interleaved per pixel), K = 256; one thread = one pixel of two rows = 16 bytes.
__global__ void __launch_bounds__(256)
im2col_rgb_kernel(const uint2* __restrict__ in, int h, int w, int oh, int ow, uint4* __restrict__ out) {
  // thread -> (output pixel, j, px): 32 threads per output pixel, consecutive threads write consecutive 16 B.  K order of
  // the direct stem: k = j*64 + px*8 + rr*4 + ch, kernel row r = 2*j + rr (r = 7 is zero padding)
  const int lane32 = threadIdx.x & 31;
  const int pix_in_block = threadIdx.x >> 5;
  const int ox = blockIdx.x * 8 + pix_in_block;
  const int oy = blockIdx.y;
  const int img = blockIdx.z;
  if (ox >= ow) return;
  const int j = lane32 >> 3, px = lane32 & 7;
  const int iy0 = 2 * oy - 3 + 2 * j, ix = 2 * ox - 4 + px;
  uint2 a = make_uint2(0u, 0u), b = make_uint2(0u, 0u);
  if (ix >= 0 && ix < w) {
    if (iy0 >= 0 && iy0 < h) a = __ldg(in + ((size_t)img * h + iy0) * w + ix);
    if (j < 3 && iy0 + 1 >= 0 && iy0 + 1 < h) b = __ldg(in + ((size_t)img * h + iy0 + 1) * w + ix);
  }
  out[(((size_t)img * oh + oy) * ow + ox) * 32 + lane32] = make_uint4(a.x, a.y, b.x, b.y);
}

// fp32 single-channel depth: k = r*8 + px, K = 64; one thread = one kernel row = 8 pixels = 16 bytes of bf16.
__global__ void __launch_bounds__(256)
im2col_depth_kernel(const float* __restrict__ in, int h, int w, int oh, int ow, uint4* __restrict__ out) {
  const int r = threadIdx.x & 7;
  const int ox = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int oy = blockIdx.y;
  const int img = blockIdx.z;
  if (ox >= ow) return;
  const int iy = 2 * oy - 3 + r;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (r < 7 && iy >= 0 && iy < h) {
    const float* row = in + ((size_t)img * h + iy) * w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ix = 2 * ox - 4 + j;
      if (ix >= 0 && ix < w) v[j] = __ldg(row + ix);
    }
  }
  out[(((size_t)img * oh + oy) * ow + ox) * 8 + r] =
      make_uint4(hn_pack_bf16(v[0], v[1]), hn_pack_bf16(v[2], v[3]), hn_pack_bf16(v[4], v[5]), hn_pack_bf16(v[6], v[7]));
}

// ------------------------------------------------------------------------------------ maxpool
__device__ __forceinline__ uint32_t bf162_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const uint4* __restrict__ in, int n, int h, int w, int c8, int oh, int ow, int halo,
                    uint4* __restrict__ out) {
  // grid (ceil(ow*c8 / 256), oh, n): consecutive threads = consecutive channel groups of consecutive pixels
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ow * c8) return;
  const int ox = t / c8, cg = t - ox * c8;
  const int oy = blockIdx.y, img = blockIdx.z;
  const uint32_t NEG = 0xff80ff80u;   // (-inf, -inf) in bf16
  uint4 m = make_uint4(NEG, NEG, NEG, NEG);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = 2 * oy - 1 + r;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ix = 2 * ox - 1 + s;
      if (ix < 0 || ix >= w) continue;
      const uint4 v = __ldg(in + (((size_t)img * h + iy) * w + ix) * c8 + cg);
      m.x = bf162_max(m.x, v.x); m.y = bf162_max(m.y, v.y); m.z = bf162_max(m.z, v.z); m.w = bf162_max(m.w, v.w);
    }
  }
  const int ohp = oh + 2 * halo, owp = ow + 2 * halo;
  out[(((size_t)img * ohp + oy + halo) * owp + ox + halo) * c8 + cg] = m;
}

// ---------------------------------------------------------------------------------- GroupNorm
// One block = a run of pixels of one image of one pyramid level (blocks are numbered level by level, image by image).
// Each block first turns the fp64 (sum, sumsq) accumulators of its image into fp32 per-channel a = rstd*gamma,
// b = beta - mean*a in shared memory, then streams pixels: y = relu(x*a + b).
constexpr int GN_MAX_C = 512;
constexpr int GN_MAX_LEVELS = 3;
struct GnLevels {
  uint4* x[GN_MAX_LEVELS];
  const long long* stats[GN_MAX_LEVELS];        // 40.24 fixed-point (sum, sumsq), see hn_conv_desc.gn_stats
  int h[GN_MAX_LEVELS], w[GN_MAX_LEVELS];
  int blocks_per_image[GN_MAX_LEVELS];
  int block_begin[GN_MAX_LEVELS + 1];
  int n_levels;
};
// (Round 2: one thread per 16 bytes, (a, b) of all channels rebuilt in double arithmetic by every block of 32 pixels and read
// from shared memory per element, row / column by division per element: ~120 warp instructions per 32 elements, 3.5 TB/s,
// instruction-bound.  Now: statistics once per group and block, coefficients in registers, one pixel per warp step: 4.2 TB/s,
// what a copy of the same 153 MB reaches under the same protocol.)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 4)
groupnorm_relu_kernel(const __grid_constant__ GnLevels lv, int c, int halo, int groups, int group_size,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int pixels_per_block) {
  __shared__ float sa[GN_MAX_C], sb[GN_MAX_C];
  __shared__ float s_mean[GN_MAX_C], s_rstd[GN_MAX_C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int blk = blockIdx.x; blk < lv.block_begin[lv.n_levels]; blk += gridDim.x) {
    const int l = (lv.n_levels > 2 && blk >= lv.block_begin[2]) ? 2 : ((lv.n_levels > 1 && blk >= lv.block_begin[1]) ? 1 : 0);
    const int rem_b = blk - lv.block_begin[l];
    const int img = rem_b / lv.blocks_per_image[l], pb = rem_b - img * lv.blocks_per_image[l];
    const int h = lv.h[l], w = lv.w[l];
    uint4* __restrict__ x = lv.x[l];
    const long long* __restrict__ stats = lv.stats[l];
    // mean / rstd of the image's groups (double arithmetic on the fixed-point sums: one thread per group) ...
    for (int g = threadIdx.x; g < groups; g += THREADS) {
      const double cnt = (double)h * (double)w * (double)group_size;
      const double s = (double)stats[((size_t)img * groups + g) * 2] * (1.0 / 16777216.0);
      const double q = (double)stats[((size_t)img * groups + g) * 2 + 1] * (1.0 / 16777216.0);
      const double mean = s / cnt;
      double var = q / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
      s_mean[g] = (float)mean;
    }
    __syncthreads();
    const int c8 = c >> 3;
    const int hp = h + 2 * halo, wp = w + 2 * halo;
    const int p0 = pb * pixels_per_block;
    const int p1 = min(h * w, p0 + pixels_per_block);
    auto relu_affine = [](uint4 v, const float* a, const float* b) {
      v.x = hn_pack_bf16(fmaxf(hn_bf16_lo(v.x) * a[0] + b[0], 0.f), fmaxf(hn_bf16_hi(v.x) * a[1] + b[1], 0.f));
      v.y = hn_pack_bf16(fmaxf(hn_bf16_lo(v.y) * a[2] + b[2], 0.f), fmaxf(hn_bf16_hi(v.y) * a[3] + b[3], 0.f));
      v.z = hn_pack_bf16(fmaxf(hn_bf16_lo(v.z) * a[4] + b[4], 0.f), fmaxf(hn_bf16_hi(v.z) * a[5] + b[5], 0.f));
      v.w = hn_pack_bf16(fmaxf(hn_bf16_lo(v.w) * a[6] + b[6], 0.f), fmaxf(hn_bf16_hi(v.w) * a[7] + b[7], 0.f));
      return v;
    };
    if (c8 == 32) {
      // ... and y = relu(x * a + b) with a = rstd * gamma, b = beta - mean * a.  256 channels: a warp takes one pixel per
      // step (32 lanes x 16 bytes = the pixel's 512 contiguous bytes), so a lane keeps the (a, b) of ITS eight channels in
      // registers for the whole pixel block, the pixel's row / column is computed once per warp, and U pixels are in flight.
      float a[8], b[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ch = lane * 8 + k, g = ch / group_size;
        a[k] = s_rstd[g] * __ldg(gamma + ch);
        b[k] = __ldg(beta + ch) - s_mean[g] * a[k];
      }
      constexpr int U = 4, NW = THREADS / 32;
      for (int pix0 = p0 + warp; pix0 < p1; pix0 += U * NW) {
        uint4* ptr[U];
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int pix = min(pix0 + u * NW, p1 - 1);
          const int yy = pix / w, xx = pix - yy * w;
          ptr[u] = x + (((size_t)img * hp + yy + halo) * wp + xx + halo) * 32 + lane;
          v[u] = *ptr[u];
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (pix0 + u * NW < p1) *ptr[u] = relu_affine(v[u], a, b);
      }
    } else {
      for (int ch = threadIdx.x; ch < c; ch += THREADS) {
        const int g = ch / group_size;
        const float av = s_rstd[g] * __ldg(gamma + ch);
        sa[ch] = av;
        sb[ch] = __ldg(beta + ch) - s_mean[g] * av;
      }
      __syncthreads();
      const int items = (p1 - p0) * c8;
      for (int i = threadIdx.x; i < items; i += THREADS) {
        const int pix = p0 + i / c8;
        const int cg = i - (i / c8) * c8;
        const int yy = pix / w, xx = pix - yy * w;
        uint4* ptr = x + (((size_t)img * hp + yy + halo) * wp + xx + halo) * c8 + cg;
        *ptr = relu_affine(*ptr, sa + cg * 8, sb + cg * 8);
      }
    }
    __syncthreads();
  }
}

}  // namespace

static int preprocess_impl(const float* const* images_host, const int* in_h_host, const int* in_w_host,
                           const int* out_h_host, const int* out_w_host, int batch, const float* mean3_host,
                           const float* std3_host, void* canvas_bf16, int canvas_h, int canvas_w, int pad_top, int pad_left,
                           int pitch_h, int pitch_w, int paired, void* stream);

extern "C" int hn_preprocess_resize_pad(const float* const* images_host, const int* in_h_host, const int* in_w_host,
                                        const int* out_h_host, const int* out_w_host, int batch,
                                        const float* mean3_host, const float* std3_host, void* canvas_bf16,
                                        int canvas_h, int canvas_w, void* stream) {
  return preprocess_impl(images_host, in_h_host, in_w_host, out_h_host, out_w_host, batch, mean3_host, std3_host, canvas_bf16,
                         canvas_h, canvas_w, 0, 0, canvas_h, canvas_w, 0, stream);
}

extern "C" int hn_preprocess_resize_pad_framed(const float* const* images_host, const int* in_h_host,
                                               const int* in_w_host, const int* out_h_host, const int* out_w_host,
                                               int batch, const float* mean3_host, const float* std3_host,
                                               void* canvas_bf16, int canvas_h, int canvas_w, int pad_top, int pad_left,
                                               int pitch_h, int pitch_w, void* stream) {
  return preprocess_impl(images_host, in_h_host, in_w_host, out_h_host, out_w_host, batch, mean3_host, std3_host, canvas_bf16,
                         canvas_h, canvas_w, pad_top, pad_left, pitch_h, pitch_w, 1, stream);
}

static int preprocess_impl(const float* const* images_host, const int* in_h_host, const int* in_w_host,
                           const int* out_h_host, const int* out_w_host, int batch, const float* mean3_host,
                           const float* std3_host, void* canvas_bf16, int canvas_h, int canvas_w, int pad_top, int pad_left,
                           int pitch_h, int pitch_w, int paired, void* stream) {
  HN_REQUIRE(!paired || pitch_h % 2 == 0, "hn_preprocess_resize_pad_framed: the frame height must be even (rows are stored in pairs)");
  HN_REQUIRE(images_host && in_h_host && in_w_host && out_h_host && out_w_host && canvas_bf16 && mean3_host && std3_host,
             "hn_preprocess_resize_pad: null pointer");
  HN_REQUIRE(batch > 0 && canvas_h > 0 && canvas_w > 0, "hn_preprocess_resize_pad: empty batch or canvas");
  HN_REQUIRE(pad_top >= 0 && pad_left >= 0 && pitch_h >= canvas_h + pad_top && pitch_w >= canvas_w + pad_left,
             "hn_preprocess_resize_pad: the canvas does not fit its frame");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int b0 = 0; b0 < batch; b0 += PRE_MAX_IMAGES) {
    const int nb = (batch - b0 < PRE_MAX_IMAGES) ? batch - b0 : PRE_MAX_IMAGES;
    PreParams p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < nb; ++i) {
      HN_REQUIRE(images_host[b0 + i], "hn_preprocess_resize_pad: image %d is null", b0 + i);
      HN_REQUIRE(out_h_host[b0 + i] <= canvas_h && out_w_host[b0 + i] <= canvas_w && out_h_host[b0 + i] > 0 &&
                     out_w_host[b0 + i] > 0 && in_h_host[b0 + i] > 0 && in_w_host[b0 + i] > 0,
                 "hn_preprocess_resize_pad: image %d does not fit the canvas", b0 + i);
      p.img[i] = images_host[b0 + i];
      p.in_h[i] = in_h_host[b0 + i];
      p.in_w[i] = in_w_host[b0 + i];
      p.out_h[i] = out_h_host[b0 + i];
      p.out_w[i] = out_w_host[b0 + i];
      p.scale_y[i] = (float)p.in_h[i] / (float)p.out_h[i];
      p.scale_x[i] = (float)p.in_w[i] / (float)p.out_w[i];
    }
    for (int c = 0; c < 3; ++c) {
      p.mean[c] = mean3_host[c];
      p.stdv[c] = std3_host[c];
      p.inv_std[c] = 1.0f / std3_host[c];
    }
    p.canvas_h = canvas_h;
    p.canvas_w = canvas_w;
    p.pad_top = pad_top;
    p.pad_left = pad_left;
    p.pitch_h = pitch_h;
    p.pitch_w = pitch_w;
    p.paired = paired;
    p.batch_offset = b0;
    int max_w = 0;
    bool aligned = true;                             // every source row starts on a 16-byte boundary
    for (int i = 0; i < nb; ++i) {
      max_w = p.in_w[i] > max_w ? p.in_w[i] : max_w;
      aligned = aligned && (p.in_w[i] & 3) == 0 && (reinterpret_cast<uintptr_t>(p.img[i]) & 15) == 0;
    }
    const int pitch = (max_w + 3) & ~3;
    const size_t staged_bytes = (size_t)3 * pitch * sizeof(float);
    const size_t pair_bytes = (size_t)18 * pitch * sizeof(float) + (size_t)canvas_w * sizeof(float2);
    static const bool use_pairs = !(getenv("HN_PRE_PAIRS") && atoi(getenv("HN_PRE_PAIRS")) == 0);
    if (use_pairs && aligned && pair_bytes <= 200 * 1024) {
      // frame rows 2k, 2k+1 that hold canvas rows: pairs first_pair .. first_pair + npairs - 1
      const int first_pair = pad_top >> 1, npairs = ((pad_top + canvas_h - 1) >> 1) - first_pair + 1;
      // pairs per CTA: long walks amortise the per-CTA set-up, but keep at least ~4 CTAs per SM slot in the grid
      const long long total = (long long)npairs * nb;
      const int ppc = total >= 8 * 1776 ? 8 : (total >= 4 * 1776 ? 4 : 2);          // (64 VGA frames: 2 / 4 / 8 / 16 pairs -> 180 / 174 / 172 / 176 us)
      dim3 grid(hn_div_up(npairs, ppc), 1, nb);
      const int per_thread = hn_div_up(canvas_w, PRE_T);                              // pixels of a row per thread
      const int threads = (hn_div_up(canvas_w, per_thread) + 31) / 32 * 32;     // (64 VGA frames: 192 / 224 / 256 threads within 2 %)
      static bool attr_set = false;
      if (!attr_set) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(preprocess_pairs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        HN_CHECK_CUDA(cudaFuncSetAttribute(preprocess_pairs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
      }
      uint2* cv = reinterpret_cast<uint2*>(canvas_bf16);
      if (paired) preprocess_pairs_kernel<true><<<grid, threads, pair_bytes, st>>>(p, cv, pitch, ppc, first_pair, npairs);
      else preprocess_pairs_kernel<false><<<grid, threads, pair_bytes, st>>>(p, cv, pitch, ppc, first_pair, npairs);
    } else if (staged_bytes <= 48 * 1024) {
      int threads = ((hn_div_up(canvas_w, 4) + 31) / 32) * 32;   // ~4 canvas pixels per thread
      threads = threads > 512 ? 512 : threads;
      preprocess_rows_kernel<<<dim3(1, canvas_h, nb), threads, staged_bytes, st>>>(p, reinterpret_cast<uint2*>(canvas_bf16),
                                                                                  pitch);
    } else {                                         // source rows too wide to stage: gather from global memory
      HN_REQUIRE(!paired, "hn_preprocess_resize_pad_framed: source rows wider than %d pixels are not supported", 48 * 1024 / 12);
      dim3 grid(hn_div_up(hn_div_up(canvas_w, 4), 256), canvas_h, nb);
      preprocess_kernel<<<grid, 256, 0, st>>>(p, reinterpret_cast<uint2*>(canvas_bf16));
    }
    hn_count_launch();
    HN_LAUNCH_CHECK();
  }
  return HN_OK;
}

extern "C" int hn_im2col_7x7s2(const void* in, int in_is_f32, int n, int h, int w, int c, void* out_bf16, int k_pad,
                               void* stream) {
  HN_REQUIRE(in && out_bf16, "hn_im2col_7x7s2: null pointer");
  HN_REQUIRE(n > 0 && h > 0 && w > 0, "hn_im2col_7x7s2: bad shape");
  HN_REQUIRE((in_is_f32 && c == 1 && k_pad == 64) || (!in_is_f32 && c == 4 && k_pad == 256),
             "hn_im2col_7x7s2: supported layouts are bf16 [n,h,w,4] -> K=256 and fp32 [n,h,w] -> K=64");
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;   // floor((h + 6 - 7)/2) + 1
  HN_REQUIRE(n <= 65535 && oh <= 65535, "hn_im2col_7x7s2: too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (in_is_f32) {
    dim3 grid(hn_div_up(ow, 32), oh, n);
    im2col_depth_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(in), h, w, oh, ow,
                                              reinterpret_cast<uint4*>(out_bf16));
  } else {
    dim3 grid(hn_div_up(ow, 8), oh, n);
    im2col_rgb_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint2*>(in), h, w, oh, ow,
                                            reinterpret_cast<uint4*>(out_bf16));
  }
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int hn_maxpool3x3s2(const void* in, int n, int h, int w, int c, void* out, int out_halo, void* stream) {
  HN_REQUIRE(in && out, "hn_maxpool3x3s2: null pointer");
  HN_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && out_halo >= 0, "hn_maxpool3x3s2: bad shape (c %% 8 == 0)");
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;   // floor((h + 2 - 3)/2) + 1
  HN_REQUIRE(oh <= 65535 && n <= 65535, "hn_maxpool3x3s2: too large");
  dim3 grid(hn_div_up(ow * (c / 8), 256), oh, n);
  maxpool3x3s2_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(in), n, h, w, c / 8, oh, ow, out_halo, reinterpret_cast<uint4*>(out));
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int hn_groupnorm_relu_levels(void* const* x_host, const int* n_host, const int* h_host, const int* w_host, int n_levels,
                                        int c, int halo, const int64_t* const* stats_host, int groups, const float* gamma,
                                        const float* beta, float eps, void* stream) {
  HN_REQUIRE(x_host && n_host && h_host && w_host && stats_host && gamma && beta, "hn_groupnorm_relu: null pointer");
  HN_REQUIRE(n_levels >= 1 && n_levels <= GN_MAX_LEVELS, "hn_groupnorm_relu: 1..%d levels (got %d)", GN_MAX_LEVELS, n_levels);
  HN_REQUIRE(groups > 0 && c % groups == 0 && c % 8 == 0 && c <= GN_MAX_C, "hn_groupnorm_relu: unsupported shape (c=%d groups=%d)",
             c, groups);
  long long pixels = 0;
  for (int l = 0; l < n_levels; ++l) {
    HN_REQUIRE(x_host[l] && stats_host[l] && n_host[l] > 0 && h_host[l] > 0 && w_host[l] > 0, "hn_groupnorm_relu: level %d", l);
    pixels += (long long)n_host[l] * h_host[l] * w_host[l];
  }
  // four resident 256-thread blocks per SM, all at once; each block amortises its image's statistics over its pixels
  // (8 frames, P3+P4+P5: 32 / 64 / 128 / 256 / 512 pixels per block -> 45.0 / 36.9 / 34.8 / 34.8 / 36.9 us)
  const long long want = (long long)hn_num_sms() * 4;
  int ppb = (int)((pixels + want - 1) / want);
  if (ppb < 32) ppb = 32;
  GnLevels lv;
  memset(&lv, 0, sizeof(lv));
  lv.n_levels = n_levels;
  int blocks = 0;
  for (int l = 0; l < n_levels; ++l) {
    lv.x[l] = reinterpret_cast<uint4*>(x_host[l]);
    lv.stats[l] = reinterpret_cast<const long long*>(stats_host[l]);
    lv.h[l] = h_host[l];
    lv.w[l] = w_host[l];
    lv.blocks_per_image[l] = hn_div_up(h_host[l] * w_host[l], ppb);
    lv.block_begin[l] = blocks;
    blocks += lv.blocks_per_image[l] * n_host[l];
  }
  lv.block_begin[n_levels] = blocks;
  groupnorm_relu_kernel<256><<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(lv, c, halo, groups, c / groups, gamma,
                                                                                       beta, eps, ppb);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int hn_groupnorm_relu(void* x, int n, int h, int w, int c, int halo, const int64_t* stats, int groups,
                                 const float* gamma, const float* beta, float eps, void* stream) {
  HN_REQUIRE(x && stats, "hn_groupnorm_relu: null pointer");
  return hn_groupnorm_relu_levels(&x, &n, &h, &w, 1, c, halo, &stats, groups, gamma, beta, eps, stream);
}
