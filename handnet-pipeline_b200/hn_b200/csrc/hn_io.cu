// The steps immediately before and after the detect -> pose path, all HBM-bound byte work:
//   hn_ingest_frames     camera frames as they arrive in ros_demo.py (uint8 BGR HWC, uint16 millimetres) -> the fp32
//                        RGB CHW 0..1 / fp32 metres tensors HandNet.forward takes (ros_demo.py:227-238, 266-267);
//                        host -> device then moves 1 + 2 bytes per pixel instead of 12 + 4
//   hn_pack_nhwc4_frame  fp32 NCHW crops -> the zero-framed 4-channel bf16 canvas of the direct 7x7 stem, with a channel
//                        permutation (the RGBD A2J variant reorders [2,1,0,3], handnet_pipeline.py:102)
//   hn_convert_joints    a2j.convert_joints + uvd2xyz, batched (a2j/a2j.py:17-43, datasets3d/a2jdataset.py:31-38)
#include "hn_common.cuh"

namespace {

// One thread = four consecutive pixels of one image row: 12 bytes of BGR in (three 32-bit loads when aligned), one
// float4 store per colour plane.  x / 255 and d / 1000 are IEEE float32 divisions, as numpy computes them.
__global__ void __launch_bounds__(256)
ingest_bgr_kernel(const uint8_t* __restrict__ bgr, int n, int h, int w, float* __restrict__ rgb) {
  const int quads = (w + 3) >> 2;
  const long long total = (long long)n * h * quads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % quads);
    const long long row = i / quads;                 // image * h + y
    const int img = (int)(row / h), y = (int)(row - (long long)img * h);
    const int x0 = q * 4;
    const uint8_t* src = bgr + (row * w + x0) * 3;
    uint8_t px[12];
    const int valid = min(4, w - x0);
    if (valid == 4 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
      const uint32_t a = __ldg(s32), b = __ldg(s32 + 1), c = __ldg(s32 + 2);
      *reinterpret_cast<uint32_t*>(px) = a;
      *reinterpret_cast<uint32_t*>(px + 4) = b;
      *reinterpret_cast<uint32_t*>(px + 8) = c;
    } else {
      for (int j = 0; j < 12; ++j) px[j] = (j < valid * 3) ? __ldg(src + j) : (uint8_t)0;
    }
    float* plane = rgb + (size_t)img * 3 * h * w + (size_t)y * w + x0;
    const size_t hw = (size_t)h * w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {                    // output channel c = R, G, B = input byte 2 - c
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __fdiv_rn((float)px[j * 3 + (2 - c)], 255.0f);
      float* dst = plane + c * hw;
      if (valid == 4 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        for (int j = 0; j < valid; ++j) dst[j] = v[j];
      }
    }
  }
}

__global__ void __launch_bounds__(256)
ingest_depth_kernel(const uint16_t* __restrict__ mm, long long count, float* __restrict__ metres) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    metres[i] = __fdiv_rn((float)__ldg(mm + i), 1000.0f);
}

struct ChanMap { int c[4]; };

// one thread = one pixel: up to four strided fp32 reads (coalesced across the warp), one 8-byte bf16x4 store
__global__ void __launch_bounds__(256)
pack_nhwc4_kernel(const float* __restrict__ src, int n, int c, int h, int w, ChanMap map, uint2* __restrict__ frame,
                  int pad_top, int pad_left, int pitch_h, int pitch_w) {
  const long long total = (long long)n * h * w;
  const size_t hw = (size_t)h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / hw);
    const int rem = (int)(i - (long long)img * hw);
    const int y = rem / w, x = rem - y * w;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (map.c[j] >= 0) ? __ldg(src + ((size_t)img * c + map.c[j]) * hw + rem) : 0.f;
    const int fy = y + pad_top;               // frame rows are stored in pairs: [pitch_h / 2][pitch_w][2 rows][4 ch]
    frame[(((size_t)img * (pitch_h >> 1) + (fy >> 1)) * pitch_w + x + pad_left) * 2 + (fy & 1)] =
        make_uint2(hn_pack_bf16(v[0], v[1]), hn_pack_bf16(v[2], v[3]));
  }
}

// one thread = one joint.  numpy (>= 2) evaluates u * (x_max - x_min) / crop + x_min in float64 (float32 array times
// int64 scalar) and stores float32; uvd2xyz likewise ((uv - c) * z / f in float64, stored as float32), then * 1000 in
// float32.
__global__ void __launch_bounds__(128)
convert_joints_kernel(const float* __restrict__ uvd, const long long* __restrict__ crops, const int* __restrict__ has_hand,
                      const void* __restrict__ paras, int paras_kind, int n, int joints, double crop_w, double crop_h,
                      float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * joints) return;
  const int b = i / joints;
  float* o = out + (size_t)i * 3;
  if (has_hand != nullptr && has_hand[b] == 0) {
    o[0] = o[1] = o[2] = 0.f;
    return;
  }
  const long long x_min = crops[b * 4 + 0], y_min = crops[b * 4 + 1], x_max = crops[b * 4 + 2], y_max = crops[b * 4 + 3];
  const float u = uvd[(size_t)i * 3 + 0], v = uvd[(size_t)i * 3 + 1], d = uvd[(size_t)i * 3 + 2];
  float px = (float)(__dadd_rn(__ddiv_rn(__dmul_rn((double)u, (double)(x_max - x_min)), crop_w), (double)x_min));
  float py = (float)(__dadd_rn(__ddiv_rn(__dmul_rn((double)v, (double)(y_max - y_min)), crop_h), (double)y_min));
  float pz = d;
  if (paras_kind != 0) {
    float xx, yy;
    if (paras_kind == 2) {                     // float64 intrinsics: numpy promotes the expression to float64
      const double* pd = reinterpret_cast<const double*>(paras);
      xx = (float)__ddiv_rn(__dmul_rn(__dsub_rn((double)px, pd[2]), (double)pz), pd[0]);
      yy = (float)__ddiv_rn(__dmul_rn(__dsub_rn((double)py, pd[3]), (double)pz), pd[1]);
    } else {                                   // float32 intrinsics: float32 throughout
      const float* pf = reinterpret_cast<const float*>(paras);
      xx = __fdiv_rn(__fmul_rn(__fsub_rn(px, pf[2]), pz), pf[0]);
      yy = __fdiv_rn(__fmul_rn(__fsub_rn(py, pf[3]), pz), pf[1]);
    }
    px = __fmul_rn(xx, 1000.0f);
    py = __fmul_rn(yy, 1000.0f);
    pz = __fmul_rn(pz, 1000.0f);
  }
  o[0] = px;
  o[1] = py;
  o[2] = pz;
}

int grid_for(long long work, int block) {
  const long long blocks = (work + block - 1) / block;
  const long long cap = (long long)hn_num_sms() * 16;      // a few waves of full SMs, grid-stride beyond that
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace

extern "C" int hn_ingest_frames(const void* bgr_u8, const void* depth_u16, int n, int h, int w, float* rgb_out,
                                float* depth_out, void* stream) {
  HN_REQUIRE(n > 0 && h > 0 && w > 0, "hn_ingest_frames: empty batch");
  HN_REQUIRE((bgr_u8 == nullptr) == (rgb_out == nullptr) && (depth_u16 == nullptr) == (depth_out == nullptr) &&
                 (bgr_u8 || depth_u16),
             "hn_ingest_frames: every input needs its output (and at least one pair)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bgr_u8) {
    const long long work = (long long)n * h * ((w + 3) / 4);
    ingest_bgr_kernel<<<grid_for(work, 256), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(bgr_u8), n, h, w, rgb_out);
    hn_count_launch();
    HN_LAUNCH_CHECK();
  }
  if (depth_u16) {
    HN_REQUIRE((reinterpret_cast<uintptr_t>(depth_u16) & 1) == 0, "hn_ingest_frames: depth must be 2-byte aligned");
    const long long count = (long long)n * h * w;
    ingest_depth_kernel<<<grid_for(count, 256), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(depth_u16), count, depth_out);
    hn_count_launch();
    HN_LAUNCH_CHECK();
  }
  return HN_OK;
}

extern "C" int hn_pack_nhwc4_frame(const float* src, int n, int c, int h, int w, const int* chan_map4_host, void* frame_bf16,
                                   int pad_top, int pad_left, int pitch_h, int pitch_w, void* stream) {
  HN_REQUIRE(src && chan_map4_host && frame_bf16 && n > 0 && c > 0 && h > 0 && w > 0, "hn_pack_nhwc4_frame: bad arguments");
  HN_REQUIRE(pad_top >= 0 && pad_left >= 0 && pitch_h >= h + pad_top && pitch_w >= w + pad_left,
             "hn_pack_nhwc4_frame: the image does not fit its frame");
  ChanMap m;
  for (int j = 0; j < 4; ++j) {
    HN_REQUIRE(chan_map4_host[j] < c, "hn_pack_nhwc4_frame: channel %d out of range", chan_map4_host[j]);
    m.c[j] = chan_map4_host[j];
  }
  const long long work = (long long)n * h * w;
  pack_nhwc4_kernel<<<grid_for(work, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, n, c, h, w, m, reinterpret_cast<uint2*>(frame_bf16), pad_top, pad_left, pitch_h, pitch_w);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}

extern "C" int hn_convert_joints(const float* uvd, const int64_t* crops, const int* has_hand, const void* paras4_dev,
                                 int paras_is_f64, int n, int joints, int crop_w, int crop_h, float* out, void* stream) {
  HN_REQUIRE(uvd && crops && out && n > 0 && joints > 0 && crop_w > 0 && crop_h > 0, "hn_convert_joints: bad arguments");
  const int total = n * joints;
  convert_joints_kernel<<<hn_div_up(total, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      uvd, reinterpret_cast<const long long*>(crops), has_hand, paras4_dev,
      paras4_dev == nullptr ? 0 : (paras_is_f64 ? 2 : 1), n, joints,
      (double)crop_w, (double)crop_h, out);
  hn_count_launch();
  HN_LAUNCH_CHECK();
  return HN_OK;
}
