// pose2mesh lifting (last "next" row): the three kernels behind models.pose2mesh_net.FlatPose2Mesh
//   Chebyshev graph convolution   pose2mesh/lib/models/backbones/cheby_graph_conv.py:5-47  (torch.sparse.mm + nn.Linear + BN1d)
//   PoseNet residual MLP          pose2mesh/lib/models/posenet.py:11-84                    (BN1d -> ReLU -> Linear, twice, + x)
//   mesh block glue               pose2mesh/lib/models/meshnet.py:86-117                   (linear interpolate of the block input
//                                                                                           along the FEATURE axis + add + x2 vertex upsample)
// One hand is 21 joints -> 1024 (778) vertices x <= 256 features: ~0.2 GFLOP and 300 MB of fp32 weights (the 4096-wide PoseNet).
// Everything is fp32 SIMT on purpose: the work is weight streaming (GEMV) and 5-nonzero-per-row sparse products, not tensor-core
// shaped, and fp32 keeps the result within 1e-5 of the reference.
#include "hn_common.cuh"

namespace {

// ---------------------------------------------------------------------------------- Chebyshev basis
// One warp per (batch, vertex) row of the CSR Laplacian; lanes stride over the features.
//   out[b][v][f] = alpha * sum_j L[v][j] * x[b][col_j][f] + beta * z[b][v][f]
// T1 = L T0 (alpha 1, beta 0), T2 = 2 L T1 - T0 (alpha 2, beta -1)   (cheby_graph_conv.py:26-31)
__global__ void __launch_bounds__(256)
cheby_spmm_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const float* __restrict__ val,
                  const float* __restrict__ x, const float* __restrict__ z, float alpha, float beta, int batch, int verts,
                  int feats, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= batch * verts) return;
  const int lane = threadIdx.x & 31;
  const int b = row / verts, v = row - b * verts;
  const int j0 = row_ptr[v], j1 = row_ptr[v + 1];
  const float* xb = x + (size_t)b * verts * feats;
  for (int f = lane; f < feats; f += 32) {
    float acc = 0.f;
    for (int j = j0; j < j1; ++j) acc = fmaf(__ldg(val + j), __ldg(xb + (size_t)__ldg(col + j) * feats + f), acc);
    float r = alpha * acc;
    if (z) r = fmaf(beta, z[(size_t)row * feats + f], r);
    out[(size_t)row * feats + f] = r;
  }
}

// ---------------------------------------------------------------------------------- linear + affine + ReLU
// y[m][n] = post( sum_kk pre(A[m][kk]) * W[n][kk] + bias[n] ) (+ res[m][n])
//   A is given as `planes` matrices [M][fin] and kk = f * planes + p reads plane p, feature f -- the order in which
//   graph_conv_cheby flattens its K Chebyshev terms (x.permute(3, 1, 2, 0).view(B*V, Fin*K), cheby_graph_conv.py:33-35);
//   planes = 1 is an ordinary matrix.
//   pre(a)  = relu(a * in_scale[kk] + in_shift[kk])   when in_scale != nullptr  (eval BatchNorm1d + ReLU in front: posenet.py:25-28)
//   post(t) = t * out_scale[n] + out_shift[n], ReLU when relu_out              (eval BatchNorm1d behind: cheby_graph_conv.py:40-41)
// A block = 8 warps = 8 output columns x a chunk of LIN_ROWS rows; the chunk's inputs are staged in shared memory K tile by
// K tile (pre applied once), every warp streams its weight row with coalesced loads and keeps LIN_ROWS accumulators.
constexpr int LIN_ROWS = 8, LIN_KT = 1024, LIN_COLS = 8;
struct LinArgs {
  const float* a[3];
  const float* w;
  const float* bias;
  const float* in_scale;
  const float* in_shift;
  const float* out_scale;
  const float* out_shift;
  const float* res;
  float* y;
  int m, n, fin, planes, relu_out;
};

__global__ void __launch_bounds__(32 * LIN_COLS)
linear_kernel(const LinArgs p) {
  __shared__ __align__(16) float xs[LIN_ROWS][LIN_KT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * LIN_COLS + warp;
  const int m0 = blockIdx.y * LIN_ROWS;
  const int rows = min(LIN_ROWS, p.m - m0);
  const int kd = p.fin * p.planes;
  float acc[LIN_ROWS];
#pragma unroll
  for (int r = 0; r < LIN_ROWS; ++r) acc[r] = 0.f;
  for (int i = threadIdx.x; i < LIN_ROWS * LIN_KT; i += blockDim.x) xs[i / LIN_KT][i % LIN_KT] = 0.f;
  for (int k0 = 0; k0 < kd; k0 += LIN_KT) {
    const int kt = min(LIN_KT, kd - k0);
    __syncthreads();
    // (eight independent loads per thread in flight: staged one element at a time this loop was a chain of load latencies)
    for (int i0 = threadIdx.x; i0 < rows * kt; i0 += 8 * blockDim.x) {
      float a[8], sc[8], sh[8];
      int rr[8], cc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        const bool ok = i < rows * kt;
        const int r = ok ? i / kt : 0, c = ok ? i - r * kt : 0, kk = k0 + c;
        const int f = p.planes == 1 ? kk : (p.planes == 3 ? kk / 3 : kk >> 1), pl = kk - f * p.planes;
        rr[u] = ok ? r : -1;
        cc[u] = c;
        a[u] = __ldg(p.a[pl] + (size_t)(m0 + r) * p.fin + f);
        sc[u] = p.in_scale ? __ldg(p.in_scale + kk) : 1.f;
        sh[u] = p.in_scale ? __ldg(p.in_shift + kk) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (rr[u] >= 0) xs[rr[u]][cc[u]] = p.in_scale ? fmaxf(fmaf(a[u], sc[u], sh[u]), 0.f) : a[u];
    }
    __syncthreads();
    if (n < p.n) {
      const float* wr = p.w + (size_t)n * kd + k0;
      if ((kd & 3) == 0 && (kt & 127) == 0 && (reinterpret_cast<uintptr_t>(p.w) & 15) == 0) {
        // weight streaming: 16 bytes per lane and load, four loads in flight per lane (a warp has 2 KB on the way)
#pragma unroll 1
        for (int k = lane * 4; k < kt; k += 512) {
          float4 wv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            wv[u] = (k + u * 128 < kt) ? __ldg(reinterpret_cast<const float4*>(wr + k + u * 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (k + u * 128 < kt) {
#pragma unroll
              for (int r = 0; r < LIN_ROWS; ++r) {
                const float4 xv = *reinterpret_cast<const float4*>(&xs[r][k + u * 128]);
                acc[r] = fmaf(wv[u].x, xv.x, fmaf(wv[u].y, xv.y, fmaf(wv[u].z, xv.z, fmaf(wv[u].w, xv.w, acc[r]))));
              }
            }
          }
        }
      } else {
        for (int k = lane; k < kt; k += 32) {
          const float wv = __ldg(wr + k);
#pragma unroll
          for (int r = 0; r < LIN_ROWS; ++r) acc[r] = fmaf(wv, xs[r][k], acc[r]);    // (rows beyond `rows` hold stale finite data)
        }
      }
    }
  }
  if (n >= p.n) return;
#pragma unroll
  for (int r = 0; r < LIN_ROWS; ++r) {
    float v = acc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && r < rows) {
      v += p.bias ? __ldg(p.bias + n) : 0.f;
      if (p.out_scale) v = fmaf(v, __ldg(p.out_scale + n), __ldg(p.out_shift + n));
      if (p.relu_out) v = fmaxf(v, 0.f);
      if (p.res) v += p.res[(size_t)(m0 + r) * p.n + n];
      p.y[(size_t)(m0 + r) * p.n + n] = v;
    }
  }
}

// Few rows, long K (the PoseNet layers: <= 8 hands x 4096 -> 4096, the joints -> mesh lift): weight streaming.  The whole
// (transformed) input sits in shared memory, staged once per block; a block is 16 warps, a warp owns one output column at a
// time and streams its weight row with 16-byte loads, eight in flight per lane, no barrier inside the column loop.  One block
// per SM walks the columns.
constexpr int GV_ROWS = 8, GV_WARPS = 16;
__global__ void __launch_bounds__(32 * GV_WARPS)
linear_gemv_kernel(const LinArgs p) {
  extern __shared__ __align__(16) float gx[];                      // [rows][kd]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kd = p.fin * p.planes;
  for (int i0 = threadIdx.x; i0 < p.m * kd; i0 += 8 * blockDim.x) {
    float a[8], sc[8], sh[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      const bool ok = i < p.m * kd;
      const int r = ok ? i / kd : 0, kk = ok ? i - r * kd : 0;
      const int f = p.planes == 1 ? kk : (p.planes == 3 ? kk / 3 : kk >> 1), pl = kk - f * p.planes;
      a[u] = __ldg(p.a[pl] + (size_t)r * p.fin + f);
      sc[u] = p.in_scale ? __ldg(p.in_scale + kk) : 1.f;
      sh[u] = p.in_scale ? __ldg(p.in_shift + kk) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < p.m * kd) gx[i] = p.in_scale ? fmaxf(fmaf(a[u], sc[u], sh[u]), 0.f) : a[u];
    }
  }
  __syncthreads();
  for (int n = blockIdx.x * GV_WARPS + warp; n < p.n; n += gridDim.x * GV_WARPS) {
    const float* wr = p.w + (size_t)n * kd;
    float acc[GV_ROWS];
#pragma unroll
    for (int r = 0; r < GV_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 1
    for (int k = lane * 4; k < kd; k += 8 * 128) {
      float4 wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        wv[u] = (k + u * 128 < kd) ? __ldg(reinterpret_cast<const float4*>(wr + k + u * 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (k + u * 128 < kd) {
#pragma unroll
          for (int r = 0; r < GV_ROWS; ++r) {
            if (r < p.m) {
              const float4 xv = *reinterpret_cast<const float4*>(gx + (size_t)r * kd + k + u * 128);
              acc[r] = fmaf(wv[u].x, xv.x, fmaf(wv[u].y, xv.y, fmaf(wv[u].z, xv.z, fmaf(wv[u].w, xv.w, acc[r]))));
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < GV_ROWS; ++r) {
      float v = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && r < p.m) {
        v += p.bias ? __ldg(p.bias + n) : 0.f;
        if (p.out_scale) v = fmaf(v, __ldg(p.out_scale + n), __ldg(p.out_shift + n));
        if (p.relu_out) v = fmaxf(v, 0.f);
        if (p.res) v += p.res[(size_t)r * p.n + n];
        p.y[(size_t)r * p.n + n] = v;
      }
    }
  }
}

// Tall inputs (the mesh levels: M = hands x vertices >= 64 rows, K <= 768, N <= 256) are ordinary fp32 GEMMs: GM x 64 block
// tile (GM = 128 when there are enough rows to fill the SMs, else 64), (GM / 16) x 4 accumulators per thread, K tiles of 32
// staged (transposed) in shared memory with the next tile prefetched into registers, same pre / post as linear_kernel.
constexpr int GN_ = 64;
template <int GM, int GK>
__global__ void __launch_bounds__(256)
linear_tiled_kernel(const LinArgs p) {
  constexpr int RA = GM / 16;                                       // rows per thread
  constexpr int EA = GM * GK / 256, EW = GN_ * GK / 256;            // elements of A / W a thread moves per K tile
  __shared__ __align__(16) float as[2][GK][GM + 4], ws[2][GK][GN_ + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;           // thread -> columns 4*tx.., rows RA*ty..
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN_;
  const int kd = p.fin * p.planes;
  float acc[RA][4];
#pragma unroll
  for (int i = 0; i < RA; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // k fastest across threads (coalesced along K).  The next tile's elements are fetched into registers while the current one is
  // multiplied: a block's K loop is a chain of global-load latencies otherwise.
  float ra[EA], rw[EW];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < EA; ++e) {
      const int idx = threadIdx.x + e * 256;
      const int r = idx / GK, kk = k0 + (idx % GK);
      float a = 0.f;
      if (kk < kd && m0 + r < p.m) {
        const int f = p.planes == 1 ? kk : (p.planes == 3 ? kk / 3 : kk >> 1), pl = kk - f * p.planes;
        a = __ldg(p.a[pl] + (size_t)(m0 + r) * p.fin + f);
        if (p.in_scale) a = fmaxf(fmaf(a, __ldg(p.in_scale + kk), __ldg(p.in_shift + kk)), 0.f);
      }
      ra[e] = a;
    }
#pragma unroll
    for (int e = 0; e < EW; ++e) {
      const int idx = threadIdx.x + e * 256;
      const int r = idx / GK, kk = k0 + (idx % GK);
      rw[e] = (kk < kd && n0 + r < p.n) ? __ldg(p.w + (size_t)(n0 + r) * kd + kk) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int e = 0; e < EA; ++e) {
      const int idx = threadIdx.x + e * 256;
      as[buf][idx % GK][idx / GK] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < EW; ++e) {
      const int idx = threadIdx.x + e * 256;
      ws[buf][idx % GK][idx / GK] = rw[e];
    }
  };
  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < kd; k0 += GK) {
    const bool more = k0 + GK < kd;
    if (more) fetch(k0 + GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a4[RA];
#pragma unroll
      for (int i = 0; i < RA; i += 4) {
        const float4 av = *reinterpret_cast<const float4*>(&as[buf][k][RA * ty + i]);
        a4[i] = av.x; a4[i + 1] = av.y; a4[i + 2] = av.z; a4[i + 3] = av.w;
      }
      const float4 wv = *reinterpret_cast<const float4*>(&ws[buf][k][4 * tx]);
      const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], w4[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);                                         // (the other buffer: nobody reads it in this iteration)
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < RA; ++i) {
    const int m = m0 + RA * ty + i;
    if (m >= p.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + 4 * tx + j;
      if (n >= p.n) continue;
      float v = acc[i][j] + (p.bias ? __ldg(p.bias + n) : 0.f);
      if (p.out_scale) v = fmaf(v, __ldg(p.out_scale + n), __ldg(p.out_shift + n));
      if (p.relu_out) v = fmaxf(v, 0.f);
      if (p.res) v += p.res[(size_t)m * p.n + n];
      p.y[(size_t)m * p.n + n] = v;
    }
  }
}

// ---------------------------------------------------------------------------------- block residual + vertex upsample
// out[b][v * up + j][f] = x[b][v][f] + interp(skip[b][v][:], f)      j < up
// interp = F.interpolate(skip, size=fout, mode='linear', align_corners=False) along the FEATURE axis (meshnet.py:107-114 treats
// [B, V, F] as (batch, channels, length)); up = 2 is nn.Upsample(scale_factor=2) (nearest) along the vertices (meshnet.py:69-76).
__global__ void __launch_bounds__(256)
mesh_residual_kernel(const float* __restrict__ x, const float* __restrict__ skip, int rows, int fout, int fskip, int up,
                     float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * fout) return;
  const int r = i / fout, f = i - r * fout;
  // ATen area_pixel_compute_source_index(scale = in / out, align_corners = false), clamped at 0
  const float scale = (float)fskip / (float)fout;
  float src = scale * ((float)f + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  int i0 = (int)src;
  if (i0 > fskip - 1) i0 = fskip - 1;
  const int i1 = i0 + (i0 < fskip - 1 ? 1 : 0);
  const float l1 = src - (float)i0, l0 = 1.f - l1;
  const float* s = skip + (size_t)r * fskip;
  const float v = x[i] + (l0 * s[i0] + l1 * s[i1]);
  for (int j = 0; j < up; ++j) out[((size_t)r * up + j) * fout + f] = v;
}

}  // namespace

extern "C" int hn_cheby_spmm(const int* row_ptr, const int* col, const float* val, const float* x, const float* z, float alpha,
                             float beta, int batch, int verts, int feats, float* out, void* stream) {
  HN_REQUIRE(row_ptr && col && val && x && out, "hn_cheby_spmm: null pointer");
  HN_REQUIRE(batch > 0 && verts > 0 && feats > 0, "hn_cheby_spmm: bad sizes");
  const int rows = batch * verts;
  cheby_spmm_kernel<<<hn_div_up(rows, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(row_ptr, col, val, x, z, alpha, beta,
                                                                                           batch, verts, feats, out);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int hn_linear_f32(const float* a0, const float* a1, const float* a2, int planes, int m, int fin, const float* weight,
                             const float* bias, int n, const float* in_scale, const float* in_shift, const float* out_scale,
                             const float* out_shift, int relu_out, const float* res, float* y, void* stream) {
  HN_REQUIRE(a0 && weight && y, "hn_linear_f32: null pointer");
  HN_REQUIRE(planes >= 1 && planes <= 3 && (planes < 2 || a1) && (planes < 3 || a2), "hn_linear_f32: 1..3 input planes");
  HN_REQUIRE(m > 0 && fin > 0 && n > 0, "hn_linear_f32: bad sizes");
  HN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr) && (out_scale == nullptr) == (out_shift == nullptr),
             "hn_linear_f32: scale and shift come in pairs");
  LinArgs p;
  p.a[0] = a0; p.a[1] = a1; p.a[2] = a2;
  p.w = weight; p.bias = bias;
  p.in_scale = in_scale; p.in_shift = in_shift; p.out_scale = out_scale; p.out_shift = out_shift;
  p.res = res; p.y = y;
  p.m = m; p.n = n; p.fin = fin; p.planes = planes; p.relu_out = relu_out;
  const long long kd = (long long)fin * planes;
  if (m <= GV_ROWS && kd >= 512 && kd % 4 == 0 && (size_t)m * kd * 4 <= 200 * 1024 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0) {
    static bool attr = false;
    if (!attr) {
      HN_CHECK_CUDA(cudaFuncSetAttribute(linear_gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
    const int blocks = hn_div_up(n, GV_WARPS) < hn_num_sms() ? hn_div_up(n, GV_WARPS) : hn_num_sms();
    linear_gemv_kernel<<<blocks, 32 * GV_WARPS, (size_t)m * kd * 4, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  } else if (m >= 64) {                                             // mesh levels: a tiled GEMM; else: row chunks x 8 columns per block
    if ((long long)hn_div_up(m, 128) * hn_div_up(n, GN_) >= hn_num_sms()) {
      dim3 grid(hn_div_up(n, GN_), hn_div_up(m, 128));
      linear_tiled_kernel<128, 16><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    } else {
      dim3 grid(hn_div_up(n, GN_), hn_div_up(m, 64));
      linear_tiled_kernel<64, 32><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    }
  } else {
    dim3 grid(hn_div_up(n, LIN_COLS), hn_div_up(m, LIN_ROWS));
    linear_kernel<<<grid, 32 * LIN_COLS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  }
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int hn_mesh_residual_upsample(const float* x, const float* skip, int rows, int fout, int fskip, int up, float* out,
                                         void* stream) {
  HN_REQUIRE(x && skip && out, "hn_mesh_residual_upsample: null pointer");
  HN_REQUIRE(rows > 0 && fout > 0 && fskip > 0 && up >= 1, "hn_mesh_residual_upsample: bad sizes");
  mesh_residual_kernel<<<hn_div_up(rows * fout, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, skip, rows, fout, fskip,
                                                                                                       up, out);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}
