// Memory-bound glue kernels around the tensor-core convolutions:
//   T1  normalize + bilinear resize + zero-pad into the detector canvas (torchvision transform.py:119-255)
//   7x7/2 stem patches as GEMM rows (feeds hn_conv2d_bf16 for backbone.body.conv1 / Backbone.model.conv1)
//   3x3/2 max pool (resnet maxpool)
//   GroupNorm + ReLU of the FCOS towers (fcos_utils/fcos.py:232-240)
// All are HBM/L2-bound streaming kernels: 16-byte vector accesses, one pass, no data reuse to stage.
#include "hn_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------ T1
constexpr int PRE_MAX_IMAGES = 32;
struct PreParams {
  const float* img[PRE_MAX_IMAGES];
  int in_h[PRE_MAX_IMAGES], in_w[PRE_MAX_IMAGES], out_h[PRE_MAX_IMAGES], out_w[PRE_MAX_IMAGES];
  float mean[3], inv_unused[3], stdv[3];
  int canvas_h, canvas_w, batch_offset;
};

__device__ __forceinline__ void bilinear_axis(int o, int in, int out, int& i0, int& i1, float& l0, float& l1) {
  // ATen area_pixel_compute_source_index(align_corners=false): scale = in/out in fp32, src clamped at 0
  const float scale = (float)in / (float)out;
  float src = scale * ((float)o + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) preprocess_kernel(const PreParams p, uint2* __restrict__ canvas) {
  const int b = blockIdx.z;
  const int y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= p.canvas_w) return;
  uint2 outv = make_uint2(0u, 0u);
  const int oh = p.out_h[b], ow = p.out_w[b];
  if (y < oh && x < ow) {
    const int ih = p.in_h[b], iw = p.in_w[b];
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_axis(y, ih, oh, y0, y1, ly0, ly1);
    bilinear_axis(x, iw, ow, x0, x1, lx0, lx1);
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = p.img[b] + (size_t)c * ih * iw;
      const float m = p.mean[c], s = p.stdv[c];
      const float p00 = (__ldg(pl + (size_t)y0 * iw + x0) - m) / s;
      const float p01 = (__ldg(pl + (size_t)y0 * iw + x1) - m) / s;
      const float p10 = (__ldg(pl + (size_t)y1 * iw + x0) - m) / s;
      const float p11 = (__ldg(pl + (size_t)y1 * iw + x1) - m) / s;
      v[c] = ly0 * (lx0 * p00 + lx1 * p01) + ly1 * (lx0 * p10 + lx1 * p11);
    }
    outv.x = hn_pack_bf16(v[0], v[1]);
    outv.y = hn_pack_bf16(v[2], 0.f);
  }
  canvas[((size_t)(p.batch_offset + b) * p.canvas_h + y) * p.canvas_w + x] = outv;
}

// ------------------------------------------------------------------------------------- im2col
template <bool IN_F32>
__global__ void __launch_bounds__(256)
im2col_7x7s2_kernel(const void* __restrict__ in, int n, int h, int w, int c, int cs, int oh, int ow, int k_pad,
                    uint4* __restrict__ out) {
  const int groups = k_pad >> 3;
  const long long total = (long long)n * oh * ow * groups;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int g = (int)(gid % groups);
  const long long row = gid / groups;
  const int ox = (int)(row % ow);
  const int oy = (int)((row / ow) % oh);
  const int img = (int)(row / ((long long)ow * oh));
  const int kmax = 49 * c;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = g * 8 + j;
    float x = 0.f;
    if (k < kmax) {
      const int tap = k / c, ch = k - tap * c;
      const int r = tap / 7, s = tap - r * 7;
      const int iy = 2 * oy - 3 + r, ix = 2 * ox - 3 + s;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
        const size_t off = (((size_t)img * h + iy) * w + ix) * cs + ch;
        if constexpr (IN_F32) {
          x = __ldg(reinterpret_cast<const float*>(in) + off);
        } else {
          x = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in)[off]);
        }
      }
    }
    v[j] = x;
  }
  out[gid] = make_uint4(hn_pack_bf16(v[0], v[1]), hn_pack_bf16(v[2], v[3]), hn_pack_bf16(v[4], v[5]),
                        hn_pack_bf16(v[6], v[7]));
}

// ------------------------------------------------------------------------------------ maxpool
__device__ __forceinline__ uint32_t bf162_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const uint4* __restrict__ in, int n, int h, int w, int c8, int oh, int ow, int halo,
                    uint4* __restrict__ out) {
  const long long total = (long long)n * oh * ow * c8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cg = (int)(gid % c8);
  const long long pix = gid / c8;
  const int ox = (int)(pix % ow);
  const int oy = (int)((pix / ow) % oh);
  const int img = (int)(pix / ((long long)ow * oh));
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
__global__ void __launch_bounds__(256)
groupnorm_relu_kernel(uint4* __restrict__ x, int n, int h, int w, int c8, int halo, const double* __restrict__ stats,
                      int groups, int group_size, const float* __restrict__ gamma, const float* __restrict__ beta,
                      float eps) {
  const long long total = (long long)n * h * w * c8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cg = (int)(gid % c8);
  const long long pix = gid / c8;
  const int xx = (int)(pix % w);
  const int yy = (int)((pix / w) % h);
  const int img = (int)(pix / ((long long)w * h));
  const int c0 = cg * 8;
  const int group = c0 / group_size;               // group_size is a multiple of 8
  const double cnt = (double)h * (double)w * (double)group_size;
  const double s = stats[((size_t)img * groups + group) * 2], q = stats[((size_t)img * groups + group) * 2 + 1];
  const double mean_d = s / cnt;
  double var_d = q / cnt - mean_d * mean_d;
  if (var_d < 0.0) var_d = 0.0;
  const float mean = (float)mean_d;
  const float rstd = (float)(1.0 / sqrt(var_d + (double)eps));
  const int hp = h + 2 * halo, wp = w + 2 * halo;
  uint4* ptr = x + (((size_t)img * hp + yy + halo) * wp + xx + halo) * c8 + cg;
  uint4 v = *ptr;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
  auto f = [&](float xv, float g, float b) { return fmaxf((xv - mean) * rstd * g + b, 0.f); };
  v.x = hn_pack_bf16(f(hn_bf16_lo(v.x), g0.x, b0.x), f(hn_bf16_hi(v.x), g0.y, b0.y));
  v.y = hn_pack_bf16(f(hn_bf16_lo(v.y), g0.z, b0.z), f(hn_bf16_hi(v.y), g0.w, b0.w));
  v.z = hn_pack_bf16(f(hn_bf16_lo(v.z), g1.x, b1.x), f(hn_bf16_hi(v.z), g1.y, b1.y));
  v.w = hn_pack_bf16(f(hn_bf16_lo(v.w), g1.z, b1.z), f(hn_bf16_hi(v.w), g1.w, b1.w));
  *ptr = v;
}

}  // namespace

extern "C" int hn_preprocess_resize_pad(const float* const* images_host, const int* in_h_host, const int* in_w_host,
                                        const int* out_h_host, const int* out_w_host, int batch,
                                        const float* mean3_host, const float* std3_host, void* canvas_bf16,
                                        int canvas_h, int canvas_w, void* stream) {
  HN_REQUIRE(images_host && in_h_host && in_w_host && out_h_host && out_w_host && canvas_bf16 && mean3_host && std3_host,
             "hn_preprocess_resize_pad: null pointer");
  HN_REQUIRE(batch > 0 && canvas_h > 0 && canvas_w > 0, "hn_preprocess_resize_pad: empty batch or canvas");
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
    }
    for (int c = 0; c < 3; ++c) {
      p.mean[c] = mean3_host[c];
      p.stdv[c] = std3_host[c];
    }
    p.canvas_h = canvas_h;
    p.canvas_w = canvas_w;
    p.batch_offset = b0;
    dim3 grid(hn_div_up(canvas_w, 256), canvas_h, nb);
    preprocess_kernel<<<grid, 256, 0, st>>>(p, reinterpret_cast<uint2*>(canvas_bf16));
    hn_count_launch();
    HN_LAUNCH_CHECK();
  }
  return HN_OK;
}

extern "C" int hn_im2col_7x7s2(const void* in, int in_is_f32, int n, int h, int w, int c, void* out_bf16, int k_pad,
                               void* stream) {
  HN_REQUIRE(in && out_bf16, "hn_im2col_7x7s2: null pointer");
  HN_REQUIRE(n > 0 && h > 0 && w > 0 && c >= 1 && c <= 4, "hn_im2col_7x7s2: bad shape");
  HN_REQUIRE(k_pad % 64 == 0 && k_pad >= 49 * c, "hn_im2col_7x7s2: k_pad=%d must be a multiple of 64 >= %d", k_pad, 49 * c);
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;   // floor((h + 6 - 7)/2) + 1
  const int cs = in_is_f32 ? c : 4;               // bf16 canvases are stored with 4 channels per pixel
  const long long total = (long long)n * oh * ow * (k_pad / 8);
  const long long blocks = (total + 255) / 256;
  HN_REQUIRE(blocks < (1ll << 31), "hn_im2col_7x7s2: too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (in_is_f32)
    im2col_7x7s2_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(in, n, h, w, c, cs, oh, ow, k_pad, reinterpret_cast<uint4*>(out_bf16));
  else
    im2col_7x7s2_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(in, n, h, w, c, cs, oh, ow, k_pad, reinterpret_cast<uint4*>(out_bf16));
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int hn_maxpool3x3s2(const void* in, int n, int h, int w, int c, void* out, int out_halo, void* stream) {
  HN_REQUIRE(in && out, "hn_maxpool3x3s2: null pointer");
  HN_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && out_halo >= 0, "hn_maxpool3x3s2: bad shape (c %% 8 == 0)");
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;   // floor((h + 2 - 3)/2) + 1
  const long long total = (long long)n * oh * ow * (c / 8);
  const long long blocks = (total + 255) / 256;
  HN_REQUIRE(blocks < (1ll << 31), "hn_maxpool3x3s2: too large");
  maxpool3x3s2_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(in), n, h, w, c / 8, oh, ow, out_halo, reinterpret_cast<uint4*>(out));
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int hn_groupnorm_relu(void* x, int n, int h, int w, int c, int halo, const double* stats, int groups,
                                 const float* gamma, const float* beta, float eps, void* stream) {
  HN_REQUIRE(x && stats && gamma && beta, "hn_groupnorm_relu: null pointer");
  HN_REQUIRE(n > 0 && h > 0 && w > 0 && groups > 0 && c % groups == 0 && (c / groups) % 8 == 0,
             "hn_groupnorm_relu: channels per group must be a multiple of 8 (c=%d groups=%d)", c, groups);
  const long long total = (long long)n * h * w * (c / 8);
  const long long blocks = (total + 255) / 256;
  HN_REQUIRE(blocks < (1ll << 31), "hn_groupnorm_relu: too large");
  groupnorm_relu_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint4*>(x), n, h, w, c / 8, halo, stats, groups, c / groups, gamma, beta, eps);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}
