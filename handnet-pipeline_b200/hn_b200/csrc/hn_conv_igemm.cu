// Convolution as a shifted GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces every 3x3 / 1x1 nn.Conv2d (+ FrozenBN / BN / bias / residual / ReLU) the reference reaches through
// cuDNN: torchvision resnet34 + FPN (fcos_utils/fcos.py:476), the FCOS towers and output convs
// (fcos_utils/fcos.py:232-264, 352-371), a2j/resnet.py and the A2J towers (a2j/a2j.py:44-181).
//
// Data layout: activations are haloed NHWC bf16, so for tap (r,s) the A operand of the implicit GEMM is the
// activation matrix [rows = N*Hp*Wp][Cin] shifted by (r*Wp + s) rows: one 2-D TMA box per (tap, 64-channel
// chunk), out-of-range rows zero-filled by TMA.  Weights are [K/64][Cout_pad][64] bf16 (k-block major).
//
// One persistent CTA per SM, 320 threads:
//   warp 0      TMA producer
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16; fp32 accumulators in TMEM, up to 8
//                               buffers, so the epilogue of a tile overlaps the main loops of the following ones)
//   warps 2..9  epilogue       (tcgen05.ld -> scale/shift -> +residual -> ReLU -> bf16/fp32 store, GroupNorm partial
//                               sums, optional phase-split copy for a following stride-2 conv)
//
// The kernel is a template over <BN, PIPE, FAST, SEG>; every instantiation contains ONE operand pipeline and ONE
// epilogue body (round 1 compiled all of them, the bring-up instrumentation, a multi-convolution launch and two unused
// experiments into every instantiation: 128 KB of SASS each, and ncu showed the narrow layers stalled on instruction
// fetch as often as they issued):
//   PIPE_RING  256-wide tiles: an A ring (one 136-row box per kernel row and chunk) and a B ring (one weight tile per tap)
//   PIPE_UNI   tiles <= 128 columns: unified stages -- one A box + one B box with the (up to three) weight tiles of the
//              k-step, two TMA operations and one barrier pair per k-step
//   PIPE_RB    resident weights (one narrow N tile, many M tiles per SM): only A boxes stream; for 3x3 stride 1 ONE box
//              brings the rows of all three kernel rows of a chunk (36 MMAs per barrier round trip)
//   FAST       bf16 output, complete 32-column chunks, 32-byte accesses, no split-K: a predicate-free epilogue body
//   SEG        the launch covers up to three SEGMENTS (pyramid levels) that share weights, scale and shift but have their
//              own geometry, input, output and GroupNorm accumulators: the FCOS towers and output convolutions run ONCE
//              over P3+P4+P5 (fcos_utils/fcos.py:278-289, 378-380 loop over the levels with the same modules) instead of
//              three launches of which the small ones cannot fill 148 SMs.
//
// A-box sharing: the three horizontal taps of a 3x3 kernel row read rows m0+shift-1 .. m0+shift+128 of the same
// matrix, so ONE 136-row box serves all three: their UMMA descriptors start 0, 1 and 2 rows (128 B each) into the
// box (a SWIZZLE_128B descriptor may start at any 128-byte row with base_offset 0).
#include "hn_common.cuh"

#include <stdlib.h>

#include <vector>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_BOX_ROWS_MAX = 136;                   // 128 + up to 8 rows of horizontal-tap slack
constexpr int A_SLOT_BYTES = A_BOX_ROWS_MAX * BLOCK_K * 2;   // 17 KiB (a multiple of 1024: slots keep swizzle alignment)
constexpr int EPI_WARPS = 8;                       // two per TMEM lane quarter; they split the column chunks
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int NUM_THREADS = 64 + EPI_THREADS;
constexpr int MAX_TAPS = 9;
constexpr int GN_SMEM_FLOATS = 4096;   // bytes / 4 of the CTA's GroupNorm accumulators (16 KiB = 2048 int64 sums)
constexpr int GN_SMEM_SUMS = GN_SMEM_FLOATS / 2;   // [segments][images][groups][2] fixed-point partial sums kept per CTA
// GroupNorm statistics are accumulated as 40.24 FIXED-POINT integers (hn_conv_desc.gn_stats): integer addition is
// associative, so the sums -- and everything downstream of them -- do not depend on the order in which warps and CTAs
// arrive (fp32 / fp64 atomics made the tower outputs differ in the last bits from run to run).  One pixel's partial
// (8 channels, summed in fp32 in a fixed order) is rounded to 2^-24 and every addition across pixels is an integer one, so
// the sums do not depend on the tile alignment of a frame (its position in the batch) either; |sum| < 5.5e11 fits.
constexpr float GN_FIX_SCALE = 16777216.0f;
constexpr int MAX_SEGS = 3;

enum { PIPE_RING = 0, PIPE_UNI = 1, PIPE_RB = 2 };

// Geometry and buffers of one segment of a multi-segment launch (one pyramid level).
struct SegGeo {
  int tile_begin;               // first M tile of the segment (tiles never straddle segments)
  int n_img, hp, wp, rows;      // padded height/width of one image, n_img * hp * wp
  uint32_t div_img_mul, div_wp_mul;
  int div_img_sh, div_wp_sh;
  int shift[3];                 // grp_shift of the (up to three) A-box groups in this geometry
  void* out;
  int out_hp, out_wp, out_row_offset;
  unsigned long long* gn_stats;
  int gn_off;                   // offset of the segment's accumulators in the CTA's shared-memory array
};

struct ConvParams {
  // compute geometry (== input geometry); with segments: those of segment 0
  int n_img, hp, wp, halo;     // padded height/width of one image and the halo size
  int rows;                    // n_img * hp * wp
  int num_taps, cin_chunks;
  // A-box groups: taps that read the same box of activation rows at different row offsets
  int n_groups;
  int a_box_bytes;                          // bytes one A box brings (128 or 136 rows)
  int grp_shift[MAX_TAPS];                  // first row of the group's box relative to the tile's first row
  // packed: ntaps | off_step << 4 | tap0 << 8 | tap_step << 16 | phase << 24.  Member t of the group is weight tap
  // tap0 + t * tap_step (r * kw + s) and reads the box from row t * off_step on; phase = phase image of the input
  int grp_info[MAX_TAPS];
  // unified stages (tiles of <= 128 columns, streamed weights): one stage = one A box (1-3 planes) + one B box (1-3
  // weight tiles), two TMA operations and one barrier pair per k-step.  Unit u of group g multiplies A plane
  // (units >> 8u) & 3 from row ((units >> (8u+2)) & 15) on with weight tile (units >> (8u+6)) & 3.
  int grp_units[MAX_TAPS];
  int k_steps;                              // k-steps per tile
  int uni_a_bytes, uni_b_bytes;             // bytes the two boxes of a stage bring
  int uni_plane_bytes;                      // distance between A planes inside a stage
  int uni_stages, uni_chunk_step;           // ring depth; 64-channel chunks per k-step (1x1 convolutions: up to 2)
  int uni_stride;                           // bytes between stages (>= uni_a_bytes + uni_b_bytes)
  int uni_a_rank4;                          // the A tensor map is 4-D (channels, rows, chunks, phases)
  // direct 7x7/2 stem: an M tile is 128 consecutive output columns of one output row; the A tensor map is the
  // overlapping-stride patch view of the canvas (build_conv)
  int stem_tpr, stem_h;                     // tiles per output row (0 = ordinary convolution), output rows per image
  int stem_win, stem_tile_w;                // window stem (see build_conv): output columns per tile (125; row-pair stem: 128)
  int duo;                                  // unified stages, 3x3 stride 1, one N tile: a work item is TWO consecutive M tiles that
                                            // share every weight box (stage = A box of tile 0 | A box of tile 1 | weight tiles).
                                            // The <= 128-column layers run at the L2 throughput cap and 73 % of their L2 -> SM
                                            // traffic is weights re-streamed for every M tile
  int epi_alt;                              // epilogue: the two warps of a lane quarter take alternate TILES (all chunks
                                            // of their tile) instead of alternate chunks of the same tile
  int m_tiles, n_tiles;
  int cout, cout_pad;
  const float* scale;
  const float* shift;
  int relu_lo, relu_hi;
  // residual
  const __nv_bfloat16* res;
  int res_mode, res_hp, res_wp, res_halo;
  // output
  void* out;
  int out_kind, out_hp, out_wp, out_halo;
  int out_rows_per_image, out_row_offset, out_ld, out_transpose_hw;
  __nv_bfloat16* out_phase;
  int ph_hp, ph_wp, ph_halo;
  long long ph_stride;         // elements between phase images
  unsigned long long* gn_stats;  // fixed-point (sum, sum of squares), see GN_FIX_SCALE
  int gn_groups, gn_group_size;
  int splits;                       // split-K factor (>= 1)
  float* sk_ws;                     // fp32 partial tiles [splits][m_tiles*128][sk_ld] (plain stores, summed in split order)
  long long sk_slice;               // floats per split slice
  unsigned* sk_cnt;                 // arrival counter per (m, n) tile, zero between uses
  int sk_ld;
  uint32_t div_img_mul, div_wp_mul; // x / d == (umulhi(x, mul) + x) >> sh for x < 2^31 (fastdiv())
  int div_img_sh, div_wp_sh;
  int na_stages, nb_stages;         // ring depths
  int rb_b_bytes;                   // resident-weights mode: bytes of the weight slice
  int rb3;                          // resident weights, 3x3 stride 1: ONE box brings the A rows of all three kernel rows of a
                                    // chunk ([3][136 rows][128 B], the third box dimension steps by one image row), so a
                                    // k-step is a whole 64-channel chunk: 36 MMAs per barrier round trip
  int vec32;                        // epilogue may use 32-byte global accesses (cout % 16 == 0, 32-byte aligned bases)
  int pipe_bytes;                   // bytes of the operand area this configuration really uses (host: dynamic smem size)
  int n_seg;                        // segments of the launch (SEG kernels; 1 otherwise)
  SegGeo seg[MAX_SEGS];
  int dbg_flags;                    // HN_CONV_DEBUG builds only: timing experiments (bit0 no stores, bit1 no epilogue work,
                                    // bit2 no MMA, bit3 no TMA, bit5 no epilogue role)
  long long* trace;                 // HN_CONV_DEBUG builds only: CTA 0 logs (clock64, tag) pairs per role, TRACE_EVENTS each
};

// Bring-up instrumentation (the timing-experiment flags of hn_conv_desc.debug and the clock64 trace) is compiled in only
// with -DHN_CONV_DEBUG (python -m hn_b200.build --debug); the production role loops carry none of it.
#ifdef HN_CONV_DEBUG
constexpr bool kDebug = true;
#else
constexpr bool kDebug = false;
#endif
#define HN_DBG(bits) (kDebug && (dbg_flags & (bits)))

constexpr int TRACE_EVENTS = 2048;
// role 0 producer, 1 MMA issuer, 2 first epilogue warp; written by lane 0 of CTA 0 only
__device__ __forceinline__ void hn_trace(long long* tr, int role, int& idx, int tag) {
  if constexpr (kDebug) {
    if (tr != nullptr && blockIdx.x == 0) {
      if ((threadIdx.x & 31) == 0 && idx < TRACE_EVENTS) {
        tr[(role * TRACE_EVENTS + idx) * 2] = clock64();
        tr[(role * TRACE_EVENTS + idx) * 2 + 1] = tag;
      }
      ++idx;
    }
  }
}

__device__ __forceinline__ void hn_epi_bar_sync() {   // named barrier 1: the epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
}

// Shared memory: [0, HDR_BYTES) barriers, GroupNorm accumulators, scale/shift of the N tile; then (1024-aligned) the
// operand area: [resident weight slice (RB mode only)] [A ring: na x 17 KiB] [B ring: nb x BN*128 B].
constexpr int MAX_STAGES = 12;
constexpr int HDR_BARS = 640;
constexpr int HDR_BYTES = HDR_BARS + GN_SMEM_FLOATS * 4 + 2 * 2 * 256 * 4;   // 20992, padded to 1024 below
constexpr int HDR_PAD = ((HDR_BYTES + 1023) / 1024) * 1024;
constexpr int PIPE_BYTES_MAX = 204800;                                        // 200 KiB for the operand rings
constexpr int SMEM_BYTES_ALL = 1024 /*alignment slack*/ + HDR_PAD + PIPE_BYTES_MAX;
static_assert(SMEM_BYTES_ALL <= 232448, "exceeds the 227 KiB of shared memory per CTA");

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
  // ring depths when A and B both stream (a 3x3 group consumes one A slot and three B slots)
  static constexpr int NA = 4;   // (only the 256-wide tiles use two rings; narrower ones use unified stages)
  static constexpr int NB_FIT = (PIPE_BYTES_MAX - NA * A_SLOT_BYTES) / B_STAGE_BYTES;
  static constexpr int NB = NB_FIT > MAX_STAGES ? MAX_STAGES : NB_FIT;
  static_assert(NB >= 4, "B ring too shallow");
  // accumulator buffers in TMEM: as many as its 512 columns hold (up to 8), so that the MMA warp can run several
  // tiles ahead of the epilogue and the hand-off latencies between the two never stall the tensor pipe (narrow tiles
  // are short: a 64-column tile of a 64-channel layer is ~2000 cycles of MMAs)
  static constexpr int NBUF = (512 / BN) > 8 ? 8 : (512 / BN);
  static constexpr int NBUF_LOG = NBUF == 8 ? 3 : (NBUF == 4 ? 2 : 1);
  static constexpr int TMEM_COLS = (NBUF * BN < 32) ? 32 : NBUF * BN;
};

// segment of M tile `mt` (tiles are ordered segment by segment)
__device__ __forceinline__ int seg_of(const ConvParams& p, int mt) {
  return (p.n_seg > 2 && mt >= p.seg[2].tile_begin) ? 2 : ((p.n_seg > 1 && mt >= p.seg[1].tile_begin) ? 1 : 0);
}

// ---------------------------------------------------------------------------------------------------------------
// The whole kernel: prologue (barriers, TMEM), the three role loops, teardown.
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int PIPE, bool FAST, bool SEG, bool PAIR = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a1,
                  const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_b,
                  const __grid_constant__ ConvParams p) {
  using C = Cfg<BN>;
  constexpr bool RB = PIPE == PIPE_RB;
  constexpr bool UNI = PIPE == PIPE_UNI;
  static_assert(PIPE != PIPE_RING || BN == 256, "two rings: 256-wide tiles only");
  static_assert(PIPE != PIPE_UNI || BN <= 128, "unified stages: tiles of at most 128 columns");
  static_assert(PIPE != PIPE_RB || BN <= 64, "resident weights: narrow tiles only");
  static_assert(!SEG || PIPE != PIPE_UNI, "segments: two-ring and resident-weights pipelines only");
  static_assert(!PAIR || (PIPE == PIPE_RING && FAST), "CTA pairs: the 256-wide two-ring pipeline with the FAST epilogue");
  // PAIR: the two CTAs of a cluster work on two consecutive M tiles of the same N tile as ONE tcgen05.mma.cta_group::2
  // (M = 256): each CTA loads its own A boxes and HALF of every weight tile, the leader (cluster rank 0) issues the MMAs
  // for both, every CTA drains its own TMEM.  Halves the weight traffic L2 -> shared memory and the B operand reads per SM.
  constexpr int B_ROWS = PAIR ? BN / 2 : BN;                        // weight rows this CTA holds per tap
  constexpr int B_STAGE = B_ROWS * BLOCK_K * 2;
  const int cta_rank = PAIR ? (int)hn_cluster_ctarank() : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_hdr = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_hdr);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + MAX_STAGES;
  uint64_t* b_full_ring = bars + 2 * MAX_STAGES;
  uint64_t* b_empty = bars + 3 * MAX_STAGES;
  uint64_t* tmem_full = bars + 4 * MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 8;
  uint64_t* b_full = tmem_empty + 8;                     // resident weights have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

  // ---- prologue: barrier init, TMEM allocation ----
  if (threadIdx.x == 0) {
    hn_tma_prefetch_desc(&tm_a);
    hn_tma_prefetch_desc(&tm_b);
    if constexpr (SEG) {
      hn_tma_prefetch_desc(&tm_a1);
      hn_tma_prefetch_desc(&tm_a2);
    }
    for (int s = 0; s < MAX_STAGES; ++s) {
      hn_mbar_init(&a_full[s], 1);
      hn_mbar_init(&a_empty[s], 1);
      hn_mbar_init(&b_full_ring[s], 1);
      hn_mbar_init(&b_empty[s], 1);
    }
    hn_mbar_init(b_full, 1);
    for (int b = 0; b < 8; ++b) {
      hn_mbar_init(&tmem_full[b], 1);
      hn_mbar_init(&tmem_empty[b], PAIR ? 2 * EPI_WARPS : EPI_WARPS);   // one arrive per epilogue warp (of both CTAs)
    }
    hn_mbar_init_fence();
  }
  if constexpr (PAIR) {
    __syncthreads();
    hn_cluster_sync();                       // the peer's barriers exist before anything can signal them
    if ((threadIdx.x >> 5) == 1) hn_tmem_alloc_pair<C::TMEM_COLS>(tmem_slot);
  } else {
    if ((threadIdx.x >> 5) == 1) hn_tmem_alloc<C::TMEM_COLS>(tmem_slot);
  }
  hn_tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) hn_cluster_sync();     // both CTAs hold their TMEM before the leader issues an MMA into it
  hn_tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may run
  // while the previous kernel of the stream is still draining; global memory is touched only after each role's own
  // griddepcontrol.wait, as late as it can: the producer first fetches what does not depend on the previous kernel
  // (the WEIGHTS of its first k-steps, or the whole resident slice), the MMA warp never touches global memory and
  // does not wait at all.  The early trigger lets the next kernel do the same under this one.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // everything the role loops need from the parameter block, read once (the asm statements in the loops clobber
  // memory, so anything left in `p` would be re-read from the constant bank every iteration)
  const int cin_chunks = p.cin_chunks;
  const int n_tiles = p.n_tiles;
  const int splits = p.splits;
  const int na = p.na_stages, nb = p.nb_stages;
  const int a_box_bytes = p.a_box_bytes;
  const int cout_pad = p.cout_pad;
  [[maybe_unused]] const int dbg_flags = p.dbg_flags;
  [[maybe_unused]] long long* const trace = p.trace;
  [[maybe_unused]] int tri = 0;
  const int k_steps = p.k_steps;                         // one k-step = one A box
  uint8_t* pipe = smem_hdr + HDR_PAD;                    // 1024-aligned operand area
  uint8_t* a_ring = pipe + (RB ? p.rb_b_bytes : 0);
  const int a_slot_bytes = (RB && p.rb3) ? 3 * A_SLOT_BYTES : A_SLOT_BYTES;
  uint8_t* b_ring = a_ring + na * a_slot_bytes;

  // warp index through a shuffle: tells the compiler it is warp-uniform, so the producer / MMA role loops below run
  // converged and their addresses and descriptors live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // tile schedule: work item = (M tile, N tile, K split), N fastest, strided over the CTAs
  const int first_tile = PAIR ? blockIdx.x >> 1 : blockIdx.x, tile_stride = PAIR ? gridDim.x >> 1 : gridDim.x;
  const bool duo = UNI && p.duo != 0;                    // (launch constant)
  const int num_items = (PAIR || duo) ? ((p.m_tiles + 1) >> 1) * n_tiles : p.m_tiles * n_tiles * splits;   // pairs of M tiles
  const int s_base = k_steps / splits, s_rem = k_steps - s_base * splits;

  // The producer and MMA loops are single-instruction-stream code on the kernel's critical path (a 64-wide tile has
  // only ~50 cycles of tensor work per MMA): no divisions or indexed parameter loads inside the k loop, descriptors
  // advance by additions.
  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // All 32 lanes walk the loop and poll the barriers; one elected lane issues the copies.
    int a_stage = 0, b_stage = 0;
    uint32_t a_phase = 0, b_phase = 0;
    if constexpr (RB) {
      // the layer's whole weight slice (n_tiles == 1), once per CTA
      if (hn_elect_one()) {
        hn_mbar_expect_tx(b_full, (uint32_t)p.rb_b_bytes);
        const int k_blocks = p.num_taps * cin_chunks;
        for (int kb = 0; kb < k_blocks; ++kb)
          hn_tma_load_2d(pipe + kb * C::B_STAGE_BYTES, &tm_b, b_full, 0, kb * p.cout_pad);
      }
      __syncwarp();
    }
    const int uni_stage_bytes = p.uni_a_bytes + p.uni_b_bytes, uni_chunk_step = p.uni_chunk_step;
    const int uni_stride = p.uni_stride;
    int early_b = 0;                                       // leading k-steps of the first item whose weights are on the way
    if constexpr (UNI) {
      if (first_tile < num_items && splits == 1 && !HN_DBG(8)) {
        const int nt0 = n_tiles > 1 ? first_tile % n_tiles : 0;
        early_b = k_steps < na ? k_steps : na;
        if (hn_elect_one()) {
          int g0 = 0, cc0 = 0;
          for (int e = 0; e < early_b; ++e) {
            hn_mbar_expect_tx(&a_full[e], (uint32_t)(uni_stage_bytes - ((duo && 2 * first_tile + 1 >= p.m_tiles) ? p.uni_plane_bytes : 0)));
            hn_tma_load_4d(a_ring + e * uni_stride + p.uni_a_bytes, &tm_b, &a_full[e], 0, nt0 * BN, cc0,
                           (p.grp_info[g0] >> 8) & 255);
            cc0 += uni_chunk_step;
            if (cc0 >= cin_chunks) { cc0 = 0; ++g0; }
          }
        }
        __syncwarp();
      }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int w_ = first_tile; w_ < num_items; w_ += tile_stride) {
      int st = w_, s_begin = 0, s_end = k_steps, g = 0, cc = 0;
      if (splits > 1) {                                    // tile and K split of this work item
        st = w_ / splits;
        const int ks = w_ - st * splits;
        s_begin = ks * s_base + (ks < s_rem ? ks : s_rem);
        s_end = s_begin + s_base + (ks < s_rem ? 1 : 0);
        if (UNI && uni_chunk_step > 1) {                   // 1x1: one group, k-step = chunk pair
          g = 0;
          cc = s_begin * uni_chunk_step;
        } else {
          g = s_begin / cin_chunks;                        // k-step = g * cin_chunks + cc
          cc = s_begin - g * cin_chunks;
        }
      }
      int mt = st, nt = 0;
      if (n_tiles > 1) { mt = st / n_tiles; nt = st - mt * n_tiles; }
      if constexpr (PAIR) mt = mt * 2 + cta_rank;
      if (duo) mt *= 2;                                    // first tile of the pair; the second one is mt + 1
      int m0 = mt * BLOCK_M;
      const int n0 = nt * BN;
      // segments: the tile's rows, box shifts and tensor map are those of its pyramid level
      const CUtensorMap* tma = &tm_a;
      const int* gshift = p.grp_shift;
      if constexpr (SEG) {
        const int si = seg_of(p, mt);
        m0 = (mt - p.seg[si].tile_begin) * BLOCK_M;
        gshift = p.seg[si].shift;
        tma = si == 0 ? &tm_a : (si == 1 ? &tm_a1 : &tm_a2);
      }
      int info = p.grp_info[g], shift = gshift[g];
      if constexpr (RB) {
        if (p.stem_win) {
          // window stem: ONE box per tile -- 8 canvas rows x 256 pixels from (2*ox0 - 4, 2*oy - 3) of the tile's image
          const int row = mt / p.stem_tpr;
          const int ox0 = (mt - row * p.stem_tpr) * p.stem_tile_w;
          const int sn = row / p.stem_h, oy = row - sn * p.stem_h;
          hn_mbar_wait(&a_empty[a_stage], a_phase ^ 1);
          if (hn_elect_one()) {
            hn_mbar_expect_tx(&a_full[a_stage], (uint32_t)a_box_bytes);
            hn_tma_load_3d(a_ring + a_stage * a_slot_bytes, &tm_a, &a_full[a_stage], 2 * ox0 - 4, 2 * oy - 3, sn);
          }
          if (++a_stage == na) { a_stage = 0; a_phase ^= 1; }
          continue;
        }
      }
      if constexpr (UNI) {
        int st_n = 0, st_oy = 0, st_ox = 0;                  // stem: image, output row, first output column of the tile
        if (p.stem_tpr > 0) {
          const int row = mt / p.stem_tpr;
          st_ox = (mt - row * p.stem_tpr) * BLOCK_M;
          st_n = row / p.stem_h;
          st_oy = row - st_n * p.stem_h;
        }
        for (int step = s_begin; step < s_end; ++step) {
          hn_mbar_wait(&a_empty[a_stage], a_phase ^ 1);
          hn_trace(trace, 0, tri, 1);
          if (hn_elect_one()) {
            if (HN_DBG(8)) {                                 // timing experiment: no loads at all
              hn_mbar_arrive(&a_full[a_stage]);
            } else {
              uint8_t* sa = a_ring + a_stage * uni_stride;
              const bool b_pending = w_ == first_tile && step - s_begin < early_b;   // armed and B issued before the wait
              const bool second = duo && mt + 1 < p.m_tiles;   // (the last pair of an odd tile count is a single tile)
              if (!b_pending)
                hn_mbar_expect_tx(&a_full[a_stage], (uint32_t)(uni_stage_bytes - ((duo && !second) ? p.uni_plane_bytes : 0)));
              if (p.stem_tpr > 0) hn_tma_load_4d(sa, &tm_a, &a_full[a_stage], 0, st_ox, st_oy + cc, st_n);
              else if (p.uni_a_rank4) hn_tma_load_4d(sa, &tm_a, &a_full[a_stage], 0, m0 + shift, cc, info >> 24);
              else hn_tma_load_3d(sa, &tm_a, &a_full[a_stage], cc * BLOCK_K, m0 + shift, info >> 24);
              if (second)
                hn_tma_load_3d(sa + p.uni_plane_bytes, &tm_a, &a_full[a_stage], cc * BLOCK_K, m0 + BLOCK_M + shift, info >> 24);
              if (!b_pending) hn_tma_load_4d(sa + p.uni_a_bytes, &tm_b, &a_full[a_stage], 0, n0, cc, (info >> 8) & 255);
            }
          }
          if (++a_stage == na) { a_stage = 0; a_phase ^= 1; }
          cc += uni_chunk_step;
          if (cc >= cin_chunks) {
            cc = 0;
            ++g;
            if (step + 1 < s_end) { info = p.grp_info[g]; shift = p.grp_shift[g]; }
          }
        }
      } else {
        for (int step = s_begin; step < s_end; ++step) {
          const int ntaps = info & 15;
          // ---- the A box of this (group, chunk)
          hn_mbar_wait(&a_empty[a_stage], a_phase ^ 1);
          hn_trace(trace, 0, tri, 1);
          if (hn_elect_one()) {
            if (HN_DBG(8)) {                                 // timing experiment: no loads at all
              hn_mbar_arrive(&a_full[a_stage]);
            } else {
              if constexpr (PAIR) {
                // both CTAs' boxes complete on the LEADER's barrier, which the leader arms for both
                if (cta_rank == 0) hn_mbar_expect_tx(&a_full[a_stage], 2u * (uint32_t)a_box_bytes);
                hn_tma_load_3d_pair(a_ring + a_stage * a_slot_bytes, tma, &a_full[a_stage], cc * BLOCK_K, m0 + shift, info >> 24);
              } else {
                hn_mbar_expect_tx(&a_full[a_stage], (uint32_t)a_box_bytes);
                hn_tma_load_3d(a_ring + a_stage * a_slot_bytes, tma, &a_full[a_stage], cc * BLOCK_K, m0 + shift, info >> 24);
              }
            }
          }
          if (++a_stage == na) { a_stage = 0; a_phase ^= 1; }
          // ---- one weight tile per member tap
          if constexpr (!RB) {
            // weights are stored k-block major ([k_block][cout_pad][64]): the tile of (k-block, n0) is one contiguous
            // run of BN * 128 bytes starting at row k_block * cout_pad + n0
            int krow = (((info >> 8) & 255) * cin_chunks + cc) * cout_pad + n0;
            const int krow_step = ((info >> 16) & 255) * cin_chunks * cout_pad;
            for (int t = 0; t < ntaps; ++t) {
              hn_mbar_wait(&b_empty[b_stage], b_phase ^ 1);
              hn_trace(trace, 0, tri, 2);
              if (hn_elect_one()) {
                if (HN_DBG(8)) {
                  hn_mbar_arrive(&b_full_ring[b_stage]);
                } else {
                  if constexpr (PAIR) {
                    // this CTA's half of the weight tile: rows [rank * BN/2, +BN/2)
                    if (cta_rank == 0) hn_mbar_expect_tx(&b_full_ring[b_stage], 2u * (uint32_t)B_STAGE);
                    hn_tma_load_2d_pair(b_ring + b_stage * B_STAGE, &tm_b, &b_full_ring[b_stage], 0, krow + cta_rank * B_ROWS);
                  } else {
                    hn_mbar_expect_tx(&b_full_ring[b_stage], (uint32_t)C::B_STAGE_BYTES);
                    hn_tma_load_2d(b_ring + b_stage * C::B_STAGE_BYTES, &tm_b, &b_full_ring[b_stage], 0, krow);
                  }
                }
              }
              krow += krow_step;
              if (++b_stage == nb) { b_stage = 0; b_phase ^= 1; }
            }
          }
          if (++cc == cin_chunks) {
            cc = 0;
            ++g;
            if (step + 1 < s_end) { info = p.grp_info[g]; shift = gshift[g]; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // (CTA pairs: the leader's warp issues for both CTAs; the follower's warp 1 only allocates / frees its TMEM)
    if (!PAIR || cta_rank == 0) {
    // Converged warp; tcgen05.mma / commit are issued by one elected lane.  (A second issuing warp taking alternate
    // tiles was tried twice for the narrow tiles and bought nothing once the epilogue ran, so there is one issuer.)
    constexpr uint32_t idesc = PAIR ? hn_umma_idesc_bf16_pair(BN) : hn_umma_idesc_bf16(BN);
    constexpr uint32_t A_SLOT_D = A_SLOT_BYTES >> 4, B_SLOT_D = B_STAGE >> 4, ROW_D = (BLOCK_K * 2) >> 4;
    int a_stage = 0, b_stage = 0, it = 0;
    uint32_t a_phase = 0, b_phase = 0;
    // low words of the shared-memory descriptors of slot 0 of each ring (start address >> 4 in bits 0..13); the high
    // word is the same for every operand tile
    const uint32_t a_desc0 = (uint32_t)hn_umma_smem_desc(hn_smem_u32(a_ring));
    const uint32_t b_desc0 = (uint32_t)hn_umma_smem_desc(hn_smem_u32(RB ? pipe : b_ring));
    const uint64_t desc_hi = hn_umma_smem_desc(0) & 0xffffffff00000000ull;
    if constexpr (RB) hn_mbar_wait(b_full, 0);
    // The tensor pipe queues only a few MMAs, so whatever the issuing thread does between two bursts must take less
    // than the burst it has just issued needs to execute.  Wide tiles (256 columns: 512 cycles per tap) poll and issue
    // tap by tap, so that a tap starts as soon as its weight tile has landed.  Narrow tiles have only ~50-64 cycles of
    // tensor work per MMA: they poll all barriers of a k-step first and issue its (up to) 36 MMAs and their commits as
    // one straight-line burst.
    const uint32_t uni_stage_d = (uint32_t)p.uni_stride >> 4, uni_a_d = (uint32_t)p.uni_a_bytes >> 4;
    const uint32_t uni_plane_d = (uint32_t)p.uni_plane_bytes >> 4;
    const int uni_chunk_step = p.uni_chunk_step;
    for (int w_ = first_tile; w_ < num_items; w_ += tile_stride, ++it) {
      int s_begin = 0, s_end = k_steps, g = 0, cc = 0;
      if (splits > 1) {
        const int ks = w_ % splits;                          // K split of this work item
        s_begin = ks * s_base + (ks < s_rem ? ks : s_rem);
        s_end = s_begin + s_base + (ks < s_rem ? 1 : 0);
        if (UNI && uni_chunk_step > 1) {
          g = 0;
          cc = s_begin * uni_chunk_step;
        } else {
          g = s_begin / cin_chunks;
          cc = s_begin - g * cin_chunks;
        }
      }
      int buf = it & (C::NBUF - 1);
      uint32_t d_tmem = tmem_base + buf * BN, empty_parity = ((it >> C::NBUF_LOG) & 1) ^ 1;
      bool duo_second = false;
      if constexpr (UNI) {
        if (duo) {                                           // an item = two M tiles = two adjacent accumulators
          constexpr int GROUPS = C::NBUF / 2;
          buf = it & (GROUPS - 1);
          d_tmem = tmem_base + buf * 2 * BN;
          empty_parity = ((it / GROUPS) & 1) ^ 1;
          duo_second = 2 * w_ + 1 < p.m_tiles;
        }
      }
      int info = p.grp_info[g];
      if (!HN_DBG(32)) hn_mbar_wait(&tmem_empty[buf], empty_parity);   // the epilogue has drained this buffer
      hn_trace(trace, 1, tri, 4);
      uint32_t accumulate = 0;
      if constexpr (UNI) {
        int units = p.grp_units[g];
        for (int step = s_begin; step < s_end; ++step) {
          int nu = info & 15;                                  // units of this k-step
          if (uni_chunk_step > 1 && cin_chunks - cc < nu) nu = cin_chunks - cc;   // ragged last chunk group of a 1x1
          const bool last = step == s_end - 1;
          const uint32_t sa = a_desc0 + a_stage * uni_stage_d, sb = sa + uni_a_d;
          const uint32_t ea = hn_smem_u32(&a_empty[a_stage]), tf = hn_smem_u32(&tmem_full[buf]);
          hn_mbar_wait(&a_full[a_stage], a_phase);
          hn_trace(trace, 1, tri, 1);
          hn_tc_fence_after();
          if (hn_elect_one()) {
            if (p.stem_tpr > 0) {
              // direct stem: two [128 x 128 B] tiles (row pairs oy + 2*step, +1 = kernel rows 4*step .. +3) against the two
              // weight k-blocks of the step
              if (!HN_DBG(4)) {
                hn_umma_bf16_x4(d_tmem, desc_hi | sa, desc_hi | sb, idesc, accumulate);
                hn_umma_bf16_x4(d_tmem, desc_hi | (sa + (uint32_t)((BLOCK_M * BLOCK_K * 2) >> 4)), desc_hi | (sb + B_SLOT_D), idesc, 1u);
              }
            } else if (!HN_DBG(4)) {                           // (timing experiment: bit 2 skips the MMAs)
              hn_umma_bf16_x4(d_tmem, desc_hi | (sa + (units & 3) * uni_plane_d + ((units >> 2) & 15) * ROW_D),
                              desc_hi | (sb + ((units >> 6) & 3) * B_SLOT_D), idesc, accumulate);
              if (nu > 1)
                hn_umma_bf16_x4(d_tmem, desc_hi | (sa + ((units >> 8) & 3) * uni_plane_d + ((units >> 10) & 15) * ROW_D),
                                desc_hi | (sb + ((units >> 14) & 3) * B_SLOT_D), idesc, 1u);
              if (nu > 2)
                hn_umma_bf16_x4(d_tmem, desc_hi | (sa + ((units >> 16) & 3) * uni_plane_d + ((units >> 18) & 15) * ROW_D),
                                desc_hi | (sb + ((units >> 22) & 3) * B_SLOT_D), idesc, 1u);
              if (duo_second) {       // the pair's second tile: its A box one plane further, the next accumulator, the same weights
                const uint32_t s2 = sa + uni_plane_d, d2 = d_tmem + BN;
                hn_umma_bf16_x4(d2, desc_hi | (s2 + ((units >> 2) & 15) * ROW_D), desc_hi | (sb + ((units >> 6) & 3) * B_SLOT_D), idesc,
                                accumulate);
                if (nu > 1)
                  hn_umma_bf16_x4(d2, desc_hi | (s2 + ((units >> 10) & 15) * ROW_D), desc_hi | (sb + ((units >> 14) & 3) * B_SLOT_D),
                                  idesc, 1u);
                if (nu > 2)
                  hn_umma_bf16_x4(d2, desc_hi | (s2 + ((units >> 18) & 15) * ROW_D), desc_hi | (sb + ((units >> 22) & 3) * B_SLOT_D),
                                  idesc, 1u);
              }
            }
            hn_umma_commit_addr<1>(ea);                        // stage free once these MMAs have read it
            if (last) hn_umma_commit_addr<1>(tf);              // accumulator complete -> epilogue
          }
          accumulate = 1;
          hn_trace(trace, 1, tri, 3);
          if (++a_stage == na) { a_stage = 0; a_phase ^= 1; }
          cc += uni_chunk_step;
          if (cc >= cin_chunks) {
            cc = 0;
            ++g;
            if (!last) { info = p.grp_info[g]; units = p.grp_units[g]; }
          }
        }
      } else {
        for (int step = s_begin; step < s_end; ++step) {
          const int ntaps = info & 15;
          const uint32_t off_step = ((info >> 4) & 15) * ROW_D;
          const uint32_t da_lo = a_desc0 + a_stage * A_SLOT_D;
          // resident weights: tile of tap (tap0 + t * tap_step), chunk cc
          const uint32_t db_rb = b_desc0 + (((info >> 8) & 255) * cin_chunks + cc) * B_SLOT_D;
          const uint32_t db_rb_step = ((info >> 16) & 255) * cin_chunks * B_SLOT_D;
          const bool last = step == s_end - 1;
          hn_mbar_wait(&a_full[a_stage], a_phase);
          hn_trace(trace, 1, tri, 1);
          if (RB && p.rb3) {
            // all nine taps of this chunk from one stage: A tile of kernel row t at slot + t * 17 KiB, tap (t, s) reads it
            // from row s * dil on; weight tile of tap t*3 + s, chunk cc
            const uint32_t sa = a_desc0 + a_stage * (3 * A_SLOT_D);
            const uint32_t sb = b_desc0 + cc * B_SLOT_D, tap_d = cin_chunks * B_SLOT_D;
            const uint32_t ea = hn_smem_u32(&a_empty[a_stage]), tf = hn_smem_u32(&tmem_full[buf]);
            hn_tc_fence_after();
            if (hn_elect_one()) {
              if (!HN_DBG(4)) {
#pragma unroll
                for (int t = 0; t < 3; ++t) {
#pragma unroll
                  for (int sx = 0; sx < 3; ++sx) {
                    hn_umma_bf16_x4(d_tmem, desc_hi | (sa + t * A_SLOT_D + sx * off_step), desc_hi | (sb + (t * 3 + sx) * tap_d),
                                    idesc, (t | sx) ? 1u : accumulate);
                  }
                }
              }
              hn_umma_commit_addr<1>(ea);
              if (last) hn_umma_commit_addr<1>(tf);
            }
          } else if (RB && p.stem_win) {
            // window stem: kernel row ky of the box is an un-swizzled K-major operand of 32 K values (8 pixels x 4 channels)
            // whose rows (output pixels) start 16 bytes apart: LBO = 16 (the descriptor's low word already says so), SBO =
            // 128.  Weights: k = ky * 32 + px * 4 + ch, i.e. k-block ky / 2, bytes (ky & 1) * 64 + 32 * s of its 128-byte rows.
            const uint32_t sa = a_desc0 + a_stage * A_SLOT_D;
            const uint64_t a_hi = (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46);
            const uint32_t ea = hn_smem_u32(&a_empty[a_stage]), tf = hn_smem_u32(&tmem_full[buf]);
            hn_tc_fence_after();
            if (hn_elect_one()) {
#pragma unroll
              for (int ky = 0; ky < 8; ++ky) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                  hn_umma_bf16(d_tmem, a_hi | (sa + ky * (2048 >> 4) + ks * 2),
                               desc_hi | (b_desc0 + (ky >> 1) * B_SLOT_D + (ky & 1) * 4 + ks * 2), idesc, (ky | ks) ? 1u : accumulate);
              }
              hn_umma_commit_addr<1>(ea);
              if (last) hn_umma_commit_addr<1>(tf);
            }
          } else if constexpr (RB) {
            // resident weights, one A box per group: poll once, issue the group's (up to three) taps as one burst
            const uint32_t ea = hn_smem_u32(&a_empty[a_stage]), tf = hn_smem_u32(&tmem_full[buf]);
            hn_tc_fence_after();
            if (hn_elect_one()) {
              if (!HN_DBG(4)) {
                hn_umma_bf16_x4(d_tmem, desc_hi | da_lo, desc_hi | db_rb, idesc, accumulate);
                if (ntaps > 1) hn_umma_bf16_x4(d_tmem, desc_hi | (da_lo + off_step), desc_hi | (db_rb + db_rb_step), idesc, 1u);
                if (ntaps > 2) hn_umma_bf16_x4(d_tmem, desc_hi | (da_lo + 2 * off_step), desc_hi | (db_rb + 2 * db_rb_step), idesc, 1u);
              }
              hn_umma_commit_addr<1>(ea);                       // A box free
              if (last) hn_umma_commit_addr<1>(tf);             // accumulator complete -> epilogue
            }
          } else {
            // two rings, 256-wide tiles: a tap starts as soon as its weight tile has landed
            uint32_t da_t = da_lo;
            for (int t = 0; t < ntaps; ++t) {
              hn_mbar_wait(&b_full_ring[b_stage], b_phase);
              hn_trace(trace, 1, tri, 2);
              hn_tc_fence_after();
              const uint32_t db_lo = b_desc0 + b_stage * B_SLOT_D;
              const uint32_t eb = hn_smem_u32(&b_empty[b_stage]);
              if (hn_elect_one()) {
                if constexpr (PAIR) {
                  hn_umma_bf16_x4_pair(d_tmem, desc_hi | da_t, desc_hi | db_lo, idesc, accumulate);
                  hn_umma_commit_pair(eb);                      // both CTAs' weight slots
                } else {
                  if (!HN_DBG(4)) hn_umma_bf16_x4(d_tmem, desc_hi | da_t, desc_hi | db_lo, idesc, accumulate);
                  hn_umma_commit_addr<1>(eb);                   // weight slot free once these MMAs have read it
                }
              }
              accumulate = 1;
              da_t += off_step;
              if (++b_stage == nb) { b_stage = 0; b_phase ^= 1; }
            }
            const uint32_t ea = hn_smem_u32(&a_empty[a_stage]), tf = hn_smem_u32(&tmem_full[buf]);
            if (hn_elect_one()) {
              if constexpr (PAIR) {
                hn_umma_commit_pair(ea);
                if (last) hn_umma_commit_pair(tf);                // both CTAs' epilogues
              } else {
                hn_umma_commit_addr<1>(ea);                       // A box free
                if (last) hn_umma_commit_addr<1>(tf);             // accumulator complete -> epilogue
              }
            }
          }
          accumulate = 1;
          hn_trace(trace, 1, tri, 3);
          if (++a_stage == na) { a_stage = 0; a_phase ^= 1; }
          if (++cc == cin_chunks) {
            cc = 0;
            ++g;
            if (!last) info = p.grp_info[g];
          }
        }
      }
    }
    }
  } else if (!HN_DBG(32)) {     // (experiment bit 5: no epilogue role at all -- the MMA warp does not wait for it)
    // ===================================== epilogue ==========================================
    const int quarter = warp & 3;              // TMEM lanes this warp may touch: 32*quarter .. +31
    const int half = (warp - 2) >> 2;          // which of the two warps of this quarter: takes every other chunk
    // scale/shift of the current N tile live in shared memory (the L1 left next to ~210 KiB of smem is too small
    // to keep them, and an L2 round trip per chunk was the epilogue's critical path)
    float* ss_base = reinterpret_cast<float*>(smem_hdr + HDR_BARS + GN_SMEM_FLOATS * 4);
    // per-CTA GroupNorm accumulator [segment][image][group][2] in shared memory (when it fits)
    unsigned long long* gn_acc = reinterpret_cast<unsigned long long*>(smem_hdr + HDR_BARS);
    int gn_vals = 0;
    if (p.gn_stats) {
      if constexpr (SEG) {
        for (int s = 0; s < p.n_seg; ++s) gn_vals += p.seg[s].n_img * p.gn_groups * 2;
      } else {
        gn_vals = p.n_img * p.gn_groups * 2;
      }
    }
    const bool gn_smem = gn_vals > 0 && gn_vals <= GN_SMEM_SUMS;      // (segments: checked by the host)
    if (gn_smem) {
      for (int i = threadIdx.x - 64; i < gn_vals; i += EPI_THREADS) gn_acc[i] = 0ull;
      hn_epi_bar_sync();
    }
    uint32_t* sk_flag = tmem_slot + 1;                                 // "this CTA finalises the tile"
    asm volatile("griddepcontrol.wait;" ::: "memory");                // before the first global read / write of this role
    const int halo = p.halo;
    int ss_n0[2] = {-1, -1};                    // N tile whose scale/shift each staging buffer holds
    const bool has_scale = p.scale != nullptr;
    // Alternate-tile mode (single N tile, no split-K): the epilogue of a tile is a latency chain (accumulator wait,
    // tcgen05.ld, residual / store round trips); with both warps of a quarter on the SAME tile nothing overlaps it.  Here
    // warps 2-5 drain the even tiles and warps 6-9 the odd ones, so two tiles are in flight.  Each warp arrives twice on
    // tmem_empty (the barrier counts eight arrivals per tile).  scale/shift are staged once for both buffers.
    // Duo items (two M tiles, two accumulators): warps 2-5 drain the first tile and warps 6-9 the second one, every chunk each.
    const bool alt = p.epi_alt != 0;
    const bool own_tile = alt || duo;          // this warp takes every chunk of its tile
    if (own_tile) {
      for (int i = threadIdx.x - 64; i < BN; i += EPI_THREADS) {
        const float sc = (p.scale && i < p.cout) ? __ldg(p.scale + i) : 1.0f;
        const float sh = (p.shift && i < p.cout) ? __ldg(p.shift + i) : 0.0f;
        ss_base[i] = sc; ss_base[256 + i] = sh; ss_base[512 + i] = sc; ss_base[768 + i] = sh;
      }
      ss_n0[0] = ss_n0[1] = 0;
      hn_epi_bar_sync();
    }
    int it = 0;
    for (int w_ = first_tile; w_ < num_items; w_ += tile_stride, ++it) {
      if (alt && (it & 1) != half) continue;
      const int st = splits > 1 ? w_ / splits : w_;
      if (warp == 2) hn_trace(trace, 2, tri, 1);
      int buf = it & (C::NBUF - 1);
      uint32_t acc_phase = (it >> C::NBUF_LOG) & 1;
      int tcol = buf * BN;                               // first TMEM column of this tile's accumulator
      int mt = st, nt = 0;
      if (n_tiles > 1) { mt = st / n_tiles; nt = st - mt * n_tiles; }
      if constexpr (PAIR) mt = mt * 2 + cta_rank;
      if constexpr (UNI) {
        if (duo) {
          constexpr int GROUPS = C::NBUF / 2;
          buf = it & (GROUPS - 1);
          acc_phase = (it / GROUPS) & 1;
          tcol = (buf * 2 + half) * BN;
          mt = mt * 2 + half;
        }
      }
      const bool absent = duo && mt >= p.m_tiles;        // second half of the single last pair: wait and hand back only
      const int n0 = nt * BN;
      // geometry of the tile's segment
      int g_rows = p.rows, g_hp = p.hp, g_wp = p.wp, g_nimg = p.n_img, g_tile0 = 0, g_gn_off = 0;
      uint32_t div_img_mul = p.div_img_mul, div_wp_mul = p.div_wp_mul;
      int div_img_sh = p.div_img_sh, div_wp_sh = p.div_wp_sh;
      void* g_out = p.out;
      int g_out_hp = p.out_hp, g_out_wp = p.out_wp, g_out_row_offset = p.out_row_offset;
      unsigned long long* g_gn_stats = p.gn_stats;
      if constexpr (SEG) {
        const SegGeo& sg = p.seg[seg_of(p, mt)];
        g_rows = sg.rows; g_hp = sg.hp; g_wp = sg.wp; g_nimg = sg.n_img; g_tile0 = sg.tile_begin; g_gn_off = sg.gn_off;
        div_img_mul = sg.div_img_mul; div_wp_mul = sg.div_wp_mul; div_img_sh = sg.div_img_sh; div_wp_sh = sg.div_wp_sh;
        g_out = sg.out; g_out_hp = sg.out_hp; g_out_wp = sg.out_wp; g_out_row_offset = sg.out_row_offset;
        g_gn_stats = sg.gn_stats;
      }
      const int m = (mt - g_tile0) * BLOCK_M + quarter * 32 + lane;
      // decode the padded pixel this accumulator row belongs to (divisions by multiply-high, see fastdiv())
      int img = 0, h = 0, w = 0;
      bool interior = false;
      if (p.stem_tpr > 0) {                              // stem tile: 128 output columns of one output row
        const int row = mt / p.stem_tpr;
        img = row / p.stem_h;
        h = row - img * p.stem_h;
        w = (mt - row * p.stem_tpr) * p.stem_tile_w + quarter * 32 + lane;
        interior = quarter * 32 + lane < p.stem_tile_w && w < g_wp && img < g_nimg;
      } else if (m < g_rows) {
        img = (int)((__umulhi((uint32_t)m, div_img_mul) + (uint32_t)m) >> div_img_sh);
        const int rem = m - img * (g_hp * g_wp);
        const int hh = (int)((__umulhi((uint32_t)rem, div_wp_mul) + (uint32_t)rem) >> div_wp_sh);
        const int ww = rem - hh * g_wp;
        h = hh - halo;
        w = ww - halo;
        interior = (h >= 0) && (w >= 0) && (h < g_hp - 2 * halo) && (w < g_wp - 2 * halo);
      }
      const int H = g_hp - 2 * halo, W = g_wp - 2 * halo;
      // output / residual element offsets of channel 0 of this row
      long long out_off = 0, res_off = 0, ph_off = 0;
      if (interior) {
        if (FAST || p.out_kind == 0) {
          out_off = ((long long)(img * g_out_hp + h + p.out_halo) * g_out_wp + (w + p.out_halo)) * p.cout;
        } else if (p.out_kind == 2) {
          // channel-planar fp32 [image][out_ld planes][out_rows_per_image]: channel c of this pixel at out_off + c * rows
          out_off = (long long)img * p.out_ld * p.out_rows_per_image + g_out_row_offset + (h * W + w);
        } else {
          const int pix = p.out_transpose_hw ? (w * H + h) : (h * W + w);
          out_off = ((long long)img * p.out_rows_per_image + g_out_row_offset + pix) * p.out_ld;
        }
        if (p.res_mode == 1) {
          res_off = ((long long)(img * p.res_hp + h + p.res_halo) * p.res_wp + (w + p.res_halo)) * p.cout;
        } else if (p.res_mode == 2) {
          res_off = ((long long)(img * p.res_hp + (h >> 1) + p.res_halo) * p.res_wp + ((w >> 1) + p.res_halo)) * p.cout;
        }
        if (p.out_phase) {
          const int ph = (h & 1) * 2 + (w & 1);
          ph_off = ph * p.ph_stride +
                   ((long long)(img * p.ph_hp + (h >> 1) + p.ph_halo) * p.ph_wp + ((w >> 1) + p.ph_halo)) * p.cout;
        }
      }
      // GroupNorm partial sums are reduced per warp when all its interior rows sit in one image
      const unsigned interior_mask = __ballot_sync(0xffffffffu, interior);
      const int warp_img = __shfl_sync(0xffffffffu, img, interior_mask ? (__ffs(interior_mask) - 1) : 0);
      const bool warp_uniform_img = __all_sync(0xffffffffu, (!interior) || (img == warp_img));

      constexpr int CHUNK = (BN >= 32) ? 32 : 16;
      float* ss = ss_base + (it & 1) * 512;   // [0,256) scale, [256,512) shift for columns n0 .. n0+BN
      if (ss_n0[it & 1] != n0) {              // (uniform over the epilogue warps) staged once per N tile, not per tile
        hn_epi_bar_sync();                      // a slower warp may still read this buffer for the tile before last
        for (int i = threadIdx.x - 64; i < BN; i += EPI_THREADS) {
          const int c = n0 + i;
          ss[i] = (p.scale && c < p.cout) ? __ldg(p.scale + c) : 1.0f;
          ss[256 + i] = (p.shift && c < p.cout) ? __ldg(p.shift + c) : 0.0f;
        }
        ss_n0[it & 1] = n0;
        hn_epi_bar_sync();
      }
      // residual rows are fetched one chunk ahead (the first one before the accumulator wait) so that their
      // global-load latency hides behind the wait / the previous chunk's work
      const bool vec32 = FAST || p.vec32 != 0;  // 32-byte (full-sector) global accesses: cout % 16 == 0, aligned bases
      const bool res_vec = p.res_mode != 0 && interior && vec32;
      uint32_t res_next[CHUNK / 2];
      constexpr int STEP = (BN / CHUNK >= 2) ? 2 * CHUNK : CHUNK;   // two warps interleave chunks when there are >= 2
      const int step_rt = own_tile ? CHUNK : STEP;                  // (alternate-tile / duo mode: this warp takes every chunk)
      const int c_first = (!own_tile && BN / CHUNK >= 2) ? half * CHUNK : 0;
      const bool idle_half = !own_tile && (BN / CHUNK < 2) && half == 1; // a single chunk: the second warp only arrives
      const bool res_first = res_vec && !idle_half && n0 + c_first + CHUNK <= p.cout;
      if (res_first) {
#pragma unroll
        for (int j = 0; j < CHUNK / 16; ++j) hn_ldg256(p.res + res_off + n0 + c_first + 16 * j, &res_next[8 * j]);
      }

      hn_mbar_wait(&tmem_full[buf], acc_phase);
      hn_tc_fence_after();
      if (warp == 2) hn_trace(trace, 2, tri, 2);
      const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + tcol;

      // Split-K: every work item stores its partial accumulator into its own slice of an fp32 scratch; the item that
      // arrives last (per-tile counter) adds the slices in split order -- a fixed summation order, so results do not
      // depend on which CTA finishes first -- and runs the epilogue.
      const bool split = !FAST && splits > 1;
      bool finalize = true;
      float* sk_row = nullptr;
      if constexpr (!FAST) {
        if (split) {
          sk_row = p.sk_ws + (size_t)m * p.sk_ld + n0;
          const int ks_item = w_ - st * splits;
          if constexpr (CHUNK == 32) {
#pragma unroll 1
            for (int c0 = c_first; c0 < (idle_half ? 0 : BN); c0 += STEP) {
              if (n0 + c0 >= p.cout) continue;
              uint32_t acc[32];
              hn_tmem_ld32(t_row + c0, acc);
              hn_tmem_ld_wait();
              if (interior) {
                float* dst = sk_row + (size_t)ks_item * p.sk_slice + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  __stcg(reinterpret_cast<float4*>(dst + j),
                         make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                                     __uint_as_float(acc[j + 3])));
              }
            }
          }
          hn_tc_fence_before();
          __syncwarp();
          if (lane == 0) hn_mbar_arrive(&tmem_empty[buf]);       // TMEM drained already
          __threadfence();
          hn_epi_bar_sync();
          if (threadIdx.x == 64) {
            const unsigned old = atomicAdd(p.sk_cnt + st, 1u);
            const bool is_last = old == (unsigned)splits - 1u;
            if (is_last) p.sk_cnt[st] = 0u;                      // everyone has arrived: reset for the next launch
            *sk_flag = is_last ? 1u : 0u;
          }
          hn_epi_bar_sync();
          finalize = *sk_flag != 0u;
          if (finalize) __threadfence();
        }
      }

      auto chunk_body = [&](const int c0) {
        uint32_t acc[CHUNK];
        uint32_t res_cur[CHUNK / 2];
#pragma unroll
        for (int j = 0; j < CHUNK / 2; ++j) res_cur[j] = res_next[j];
        if (res_vec && c0 + step_rt < BN && n0 + c0 + step_rt + CHUNK <= p.cout) {
#pragma unroll
          for (int j = 0; j < CHUNK / 16; ++j) hn_ldg256(p.res + res_off + n0 + c0 + step_rt + 16 * j, &res_next[8 * j]);
        }
        const int cbase = n0 + c0;
        if (!FAST && split) {
          if (cbase >= p.cout) return;
          if constexpr (CHUNK == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
              if (interior) {
                for (int sl = 0; sl < splits; ++sl) {          // fixed order: split 0, 1, 2, ...
                  const float4 u = __ldcg(reinterpret_cast<const float4*>(sk_row + (size_t)sl * p.sk_slice + c0 + j));
                  t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                }
              }
              acc[j] = __float_as_uint(t.x); acc[j + 1] = __float_as_uint(t.y);
              acc[j + 2] = __float_as_uint(t.z); acc[j + 3] = __float_as_uint(t.w);
            }
          }
        } else {
          if constexpr (CHUNK == 32) {
            hn_tmem_ld32(t_row + c0, acc);
          } else {
            hn_tmem_ld16(t_row + c0, acc);
          }
          hn_tmem_ld_wait();
          if (warp == 2) hn_trace(trace, 2, tri, 4);
          if (!FAST && cbase >= p.cout) return;        // padded output channels (warp-uniform)
        }
        float v[CHUNK];
        {
          // scale / shift of these columns: broadcast ld.shared (explicit state space: through the generic pointer the
          // compiler emitted generic loads); layers without a scale vector (bias only) skip the multiply
          const uint32_t ss_addr = hn_smem_u32(ss + c0);
          if (has_scale) {
#pragma unroll
            for (int j = 0; j < CHUNK; j += 4) {
              const float4 sc = hn_lds128(ss_addr + j * 4), sh = hn_lds128(ss_addr + 1024 + j * 4);
              v[j + 0] = __uint_as_float(acc[j + 0]) * sc.x + sh.x;
              v[j + 1] = __uint_as_float(acc[j + 1]) * sc.y + sh.y;
              v[j + 2] = __uint_as_float(acc[j + 2]) * sc.z + sh.z;
              v[j + 3] = __uint_as_float(acc[j + 3]) * sc.w + sh.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CHUNK; j += 4) {
              const float4 sh = hn_lds128(ss_addr + 1024 + j * 4);
              v[j + 0] = __uint_as_float(acc[j + 0]) + sh.x;
              v[j + 1] = __uint_as_float(acc[j + 1]) + sh.y;
              v[j + 2] = __uint_as_float(acc[j + 2]) + sh.z;
              v[j + 3] = __uint_as_float(acc[j + 3]) + sh.w;
            }
          }
        }
        if (p.res_mode != 0 && interior) {
          if (FAST || (cbase + CHUNK <= p.cout && vec32)) {
#pragma unroll
            for (int j = 0; j < CHUNK; j += 2) {
              v[j] += hn_bf16_lo(res_cur[j / 2]);
              v[j + 1] += hn_bf16_hi(res_cur[j / 2]);
            }
          } else if constexpr (!FAST) {
            const __nv_bfloat16* rp = p.res + res_off + cbase;
#pragma unroll
            for (int j = 0; j < CHUNK; ++j)
              if (cbase + j < p.cout) v[j] += __bfloat162float(rp[j]);
          }
        }
        if constexpr (CHUNK == 32) {
          // tiles of >= 32 columns: the ReLU range is chunk-aligned (checked by build_conv): one decision per chunk
          if (cbase >= p.relu_lo && cbase < p.relu_hi) {
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
        } else {
          if (p.relu_hi > p.relu_lo) {
#pragma unroll
            for (int j = 0; j < CHUNK; ++j)
              if (cbase + j >= p.relu_lo && cbase + j < p.relu_hi) v[j] = fmaxf(v[j], 0.0f);
          }
        }
        if constexpr (!FAST) {
          if (p.out_kind == 1) {
            if (interior) {
              float* op = reinterpret_cast<float*>(g_out) + out_off + cbase;
#pragma unroll
              for (int j = 0; j < CHUNK; ++j)
                if (cbase + j < p.cout) op[j] = v[j];
            }
            return;
          }
          if (p.out_kind == 2) {
            // consecutive lanes = consecutive pixels: every channel's store of a warp is one contiguous run
            if (interior) {
              float* op = reinterpret_cast<float*>(g_out) + out_off + (long long)cbase * p.out_rows_per_image;
#pragma unroll
              for (int j = 0; j < CHUNK; ++j)
                if (cbase + j < p.cout) op[(long long)j * p.out_rows_per_image] = v[j];
            }
            return;
          }
        }
        if (warp == 2) hn_trace(trace, 2, tri, 6);
        // bf16 outputs
        uint32_t packed[CHUNK / 2];
#pragma unroll
        for (int j = 0; j < CHUNK; j += 2) packed[j / 2] = hn_pack_bf16(v[j], v[j + 1]);
        if (warp == 2) hn_trace(trace, 2, tri, 7);
        if (interior && !HN_DBG(1)) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(g_out) + out_off + cbase;
          if (FAST || (cbase + CHUNK <= p.cout && vec32)) {
            // 32 bytes per lane and instruction: every store fills whole 32-byte sectors (16-byte stores at a 2*cout
            // byte lane stride left every sector half written and doubled the L2 requests)
#pragma unroll
            for (int j = 0; j < CHUNK / 2; j += 8) hn_stg256(op + 2 * j, &packed[j]);
            if (p.out_phase) {
              __nv_bfloat16* pp = p.out_phase + ph_off + cbase;
#pragma unroll
              for (int j = 0; j < CHUNK / 2; j += 8) hn_stg256(pp + 2 * j, &packed[j]);
            }
          } else if constexpr (!FAST) {
#pragma unroll
            for (int j = 0; j < CHUNK; ++j)
              if (cbase + j < p.cout) {
                const __nv_bfloat16 b = __float2bfloat16_rn(v[j]);
                op[j] = b;
                if (p.out_phase) p.out_phase[ph_off + cbase + j] = b;
              }
          }
        }
        if (warp == 2) hn_trace(trace, 2, tri, 5);
        if (p.gn_stats) {
          // GroupNorm partial sums over the bf16-rounded values.  Per lane (= one output pixel): (sum, sumsq) of the 4
          // channel octets of this chunk = 8 values, summed in a fixed channel order and converted to FIXED POINT before
          // anything is added across pixels: every cross-pixel addition is an integer addition, so the statistics -- and
          // everything downstream of them -- depend neither on the arrival order of warps and CTAs nor on which 32 rows
          // share a warp, i.e. not on where a frame sits in the batch (tile alignment) or how the batch is sharded.  A
          // transposing butterfly (4+2+1+1+1 64-bit shuffles) leaves total k in lane 4*k, which adds it to the CTA's
          // shared-memory accumulator of its (image, group); flushed once at kernel end.
          if constexpr (CHUNK == 32) {
            long long v8[8];
#pragma unroll
            for (int o8 = 0; o8 < 4; ++o8) {
              float s = 0.f, q = 0.f;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t pk = packed[o8 * 4 + j];
                const float x0 = hn_bf16_lo(pk), x1 = hn_bf16_hi(pk);
                s += x0 + x1;
                q += x0 * x0 + x1 * x1;
              }
              v8[2 * o8] = interior ? __float2ll_rn(s * GN_FIX_SCALE) : 0ll;
              v8[2 * o8 + 1] = interior ? __float2ll_rn(q * GN_FIX_SCALE) : 0ll;
            }
            if (warp_uniform_img) {
              {
                const bool hi = lane & 16;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const long long send = hi ? v8[i] : v8[i + 4];
                  const long long keep = hi ? v8[i + 4] : v8[i];
                  v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
              }
              {
                const bool hi = lane & 8;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const long long send = hi ? v8[i] : v8[i + 2];
                  const long long keep = hi ? v8[i + 2] : v8[i];
                  v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
              }
              {
                const bool hi = lane & 4;
                const long long send = hi ? v8[0] : v8[1];
                const long long keep = hi ? v8[1] : v8[0];
                v8[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
              v8[0] += __shfl_xor_sync(0xffffffffu, v8[0], 2);
              v8[0] += __shfl_xor_sync(0xffffffffu, v8[0], 1);
              if ((lane & 3) == 0 && interior_mask != 0u) {
                const int k = lane >> 2;                         // value index: octet k/2, stat k&1
                const int group = (cbase + (k >> 1) * 8) / p.gn_group_size;
                if (group < p.gn_groups) {
                  const unsigned long long fx = (unsigned long long)v8[0];
                  if (gn_smem) atomicAdd(&gn_acc[g_gn_off + (warp_img * p.gn_groups + group) * 2 + (k & 1)], fx);
                  else atomicAdd(g_gn_stats + ((long long)warp_img * p.gn_groups + group) * 2 + (k & 1), fx);
                }
              }
            } else if (interior) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int group = (cbase + (k >> 1) * 8) / p.gn_group_size;
                if (group < p.gn_groups) {
                  const unsigned long long fx = (unsigned long long)v8[k];
                  if (gn_smem) atomicAdd(&gn_acc[g_gn_off + (img * p.gn_groups + group) * 2 + (k & 1)], fx);
                  else atomicAdd(g_gn_stats + ((long long)img * p.gn_groups + group) * 2 + (k & 1), fx);
                }
              }
            }
          }
        }
      };
      const int c_end = (idle_half || absent || HN_DBG(2) || !finalize) ? 0 : BN;
      // (TMEM loads one chunk ahead -- tcgen05.ld of chunk i+1 in flight while chunk i is scaled, packed and stored -- were
      // measured in round 1: isolated 256-wide layers +5 %, the whole step -2 % with 48 bytes of spills at the 168-register
      // cap; not built.)
#pragma unroll 1
      for (int c0 = c_first; c0 < c_end; c0 += step_rt) chunk_body(c0);
      // accumulator buffer drained -> hand it back to the MMA warp (split-K items did so after their reduction)
      if (!split) {
        hn_tc_fence_before();
        __syncwarp();
        if constexpr (PAIR) {
          if (lane == 0) hn_mbar_arrive_leader(&tmem_empty[buf]);     // the leader's MMA warp waits for both CTAs' drains
        } else {
          if (lane < (alt ? 2 : 1)) hn_mbar_arrive(&tmem_empty[buf]);
        }
      }
      if (warp == 2) hn_trace(trace, 2, tri, 3);
    }
    if (gn_smem) {
      hn_epi_bar_sync();
      if constexpr (SEG) {
        for (int s = 0; s < p.n_seg; ++s) {
          const int cnt = p.seg[s].n_img * p.gn_groups * 2, off = p.seg[s].gn_off;
          for (int i = threadIdx.x - 64; i < cnt; i += EPI_THREADS) {
            const unsigned long long v = gn_acc[off + i];
            if (v != 0ull) atomicAdd(p.seg[s].gn_stats + i, v);
          }
        }
      } else {
        for (int i = threadIdx.x - 64; i < gn_vals; i += EPI_THREADS) {
          const unsigned long long v = gn_acc[i];
          if (v != 0ull) atomicAdd(p.gn_stats + i, v);
        }
      }
    }
  }

  // ---- teardown ----
  hn_tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) hn_cluster_sync();       // no CTA may retire (or free its TMEM) while its peer can still signal / use it
  if ((threadIdx.x >> 5) == 1) {
    hn_tc_fence_after();
    if constexpr (PAIR) hn_tmem_dealloc_pair<C::TMEM_COLS>(tmem_base);
    else hn_tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B,
             CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    hn_set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return HN_ERR_CUDA;
  }
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, dtype, rank, const_cast<void*>(ptr), dims, strides_bytes, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    hn_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu)", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0));
    return HN_ERR_CUDA;
  }
  return HN_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HN_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

bool duo_enabled() {            // HN_CONV_DUO=0: one M tile per work item everywhere
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HN_CONV_DUO");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int g_cta_cap = 0;      // hn_conv_set_cta_cap: upper bound on the CTAs of the following launches (0 = all SMs)
int g_pdl_off = 0;      // hn_conv_set_pdl(0): the following launches do not use programmatic dependent launch

template <int BN, int PIPE, bool FAST, bool SEG, bool PAIR = false>
int launch(const CUtensorMap* ta, const CUtensorMap& tb, const ConvParams& p, cudaStream_t st) {
  constexpr int SMEM = SMEM_BYTES_ALL;
  static bool attr_set = false;
  if (!attr_set) {
    HN_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, PIPE, FAST, SEG, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SMEM));
    attr_set = true;
  }
  // CTA pairs: a work item is a pair of M tiles and occupies a cluster of two CTAs (two SMs of one TPC)
  const int items = (PAIR || p.duo) ? ((p.m_tiles + 1) / 2) * p.n_tiles : p.m_tiles * p.n_tiles * p.splits;
  // equal work per CTA: with w = ceil(tiles / SMs) waves, ceil(tiles / w) CTAs finish at the same time as a full grid
  // would and leave the other SMs to kernels of concurrent streams (graph branches, the pose net of the previous step)
  int sms = (g_cta_cap > 0 && g_cta_cap < hn_num_sms()) ? g_cta_cap : hn_num_sms();
  if (PAIR) sms /= 2;
  const int waves = hn_div_up(items, sms);
  const int ctas = hn_div_up(items, waves) * (PAIR ? 2 : 1);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(NUM_THREADS);
  // only what this configuration uses: the 256-wide tower layers leave ~9 KB of the SM's shared memory (and 11 K registers)
  // free, enough for blocks of the memory-bound GroupNorm kernel of the OTHER tower to run on the same SMs underneath
  cfg.dynamicSmemBytes = 1024 + HDR_PAD + p.pipe_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled() && !g_pdl_off) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  HN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BN, PIPE, FAST, SEG, PAIR>, ta[0], ta[1], ta[2], tb, p));
  hn_count_launch();
  return HN_OK;
}

int pick_block_n(int cout_pad, int m_tiles, int k_blocks, int min_bn) {
  // Cost model from measurements of the current kernel (profiles/r01b_*): cycles per 64-channel k-block of one tile.
  // 256 columns run at the tensor pipe's rate (4 x 128 cycles); narrower tiles are bound by MMA issue / shared-memory
  // bandwidth (48-64 cycles per MMA) plus ~400 cycles of barrier round trip per k-step of three k-blocks.  Tiles
  // narrower than 64 columns re-read the activations once per N tile and measured clearly slower on the small A2J
  // layers (8 crops: 590 us with 64, 731 us with 32, 1141 us with 16), so they are used only when cout_pad < 64.
  // cost = waves * (k_blocks * kb_cycles + epilogue).
  const int sms = hn_num_sms();
  static int small_bn = -1;                       // experiment: HN_SMALL_BN pins the tile width of layers with < 148 M tiles
  if (small_bn < 0) {
    const char* e = getenv("HN_SMALL_BN");
    small_bn = e ? atoi(e) : 0;
  }
  if (small_bn > 0 && m_tiles < sms && cout_pad % small_bn == 0 && small_bn >= min_bn) return small_bn;
  const int cands[5] = {256, 128, 64, 32, 16};
  const double kb_cycles[5] = {530.0, 400.0, 330.0, 420.0, 560.0};
  int best = 0;
  double best_cost = 1e30;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    if (cout_pad % bn || bn < min_bn) continue;
    const long long tiles = (long long)m_tiles * (cout_pad / bn);
    const double waves = (double)((tiles + sms - 1) / sms);
    const double cost = waves * (k_blocks * kb_cycles[i] + 400.0 + 10.0 * bn);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best ? best : 16;
}

// Division by an invariant d >= 1 as (umulhi(x, mul) + x) >> sh, exact for 0 <= x < 2^31 (Granlund-Montgomery round-up
// method with a 33-bit multiplier 2^32 + mul).
void fastdiv(uint32_t d, uint32_t* mul, int* sh) {
  int l = 0;
  while ((1ull << l) < d) ++l;                                        // l = ceil(log2 d)
  const unsigned long long m = ((1ull << 32) * ((1ull << l) - d)) / d + 1;   // floor(2^32 * (2^l - d) / d) + 1
  *mul = (uint32_t)m;
  *sh = l;
}

int epi_alt_max_bn() {          // HN_EPI_ALT=<bn>: widest single-N-tile layer whose epilogue warps take alternate tiles
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HN_EPI_ALT");
    v = e ? atoi(e) : 64;      // measured: layer1 66.6 -> 53.2 us, whole step +2.4 %; wider tiles are MMA-bound (no change)
  }
  return v;
}

int split_min_kb() {            // experiment knob: HN_SPLIT_MIN_KB (k-blocks a layer needs before split-K is considered)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HN_SPLIT_MIN_KB");
    v = e ? atoi(e) : 48;
  }
  return v;
}

struct BuiltConv {
  ConvParams p;
  CUtensorMap ta, tb;
  int bn;
  bool rb;
  bool pair;      // CTA pairs (tcgen05 cta_group::2): 256-wide FAST layers
};

bool pair_enabled() {           // HN_CONV_PAIR=0 switches the CTA-pair kernels off (A/B timing)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HN_CONV_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// Validate a descriptor and derive kernel parameters + tensor maps.  total_m_tiles > 0: the descriptor is one segment of
// a multi-segment launch -- tile width and pipeline are chosen for the launch's total number of M tiles, force_bn pins the
// width the first segment chose.
int build_conv(const hn_conv_desc* d, int force_bn, int total_m_tiles, BuiltConv* out) {
  HN_REQUIRE(d && d->in && d->weight && d->out, "hn_conv2d_bf16: null pointer");
  HN_REQUIRE(d->cin > 0 && d->cin % BLOCK_K == 0, "hn_conv2d_bf16: cin=%d must be a multiple of 64", d->cin);
  HN_REQUIRE(d->kh == d->kw && (d->kh == 1 || d->kh == 3), "hn_conv2d_bf16: only 1x1 and 3x3 kernels (got %dx%d)",
             d->kh, d->kw);
  HN_REQUIRE(d->stride == 1 || d->stride == 2, "hn_conv2d_bf16: stride %d", d->stride);
  HN_REQUIRE((d->stride == 2) == (d->in_phases == 4) && (d->in_phases == 1 || d->in_phases == 4),
             "hn_conv2d_bf16: stride 2 needs a phase-split input (in_phases=4), stride 1 a plain one");
  HN_REQUIRE(d->dilation >= 1 && (d->stride == 1 || d->dilation == 1), "hn_conv2d_bf16: dilation %d", d->dilation);
  const int pad = (d->kh / 2) * d->dilation;
  HN_REQUIRE(d->stride == 2 || d->halo_in >= pad, "hn_conv2d_bf16: halo_in=%d < conv padding %d", d->halo_in, pad);
  HN_REQUIRE(d->stride == 1 || d->halo_in >= 1 || d->kh == 1, "hn_conv2d_bf16: stride-2 3x3 needs halo_in >= 1");
  HN_REQUIRE(d->cout > 0 && d->cout_pad >= d->cout && d->cout_pad % 16 == 0, "hn_conv2d_bf16: cout/cout_pad");
  HN_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, "hn_conv2d_bf16: empty input");

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = d->n;
  p.halo = d->halo_in;
  p.hp = d->h + 2 * d->halo_in;
  p.wp = d->w + 2 * d->halo_in;
  const long long rows_ll = (long long)d->n * p.hp * p.wp;
  HN_REQUIRE(rows_ll < (1ll << 31) - 4096, "hn_conv2d_bf16: too many rows");
  p.rows = (int)rows_ll;
  fastdiv((uint32_t)(p.hp * p.wp), &p.div_img_mul, &p.div_img_sh);
  fastdiv((uint32_t)p.wp, &p.div_wp_mul, &p.div_wp_sh);
  p.cin_chunks = d->cin / BLOCK_K;
  p.num_taps = d->kh * d->kw;
  // A-box groups.  Pixel (oh*stride + dr, ow*stride + ds) of tap (r, s), dr = (r - kh/2)*dil, ds = (s - kw/2)*dil, is row
  // m + shift of the (phase) matrix; taps of one kernel row whose shifts differ by a few rows share one box.
  HN_REQUIRE(2 * d->dilation <= A_BOX_ROWS_MAX - BLOCK_M, "hn_conv2d_bf16: dilation %d too large for a shared A box",
             d->dilation);
  p.m_tiles = hn_div_up(p.rows, BLOCK_M);
  const bool stem = d->stem_pitch_w > 0;
  const bool stem_win = stem && d->stem_window != 0;
  if (stem) {
    HN_REQUIRE(!total_m_tiles && d->kh == 1 && d->stride == 1 && d->cin == 256 && d->halo_in == 0 && d->in_phases == 1 &&
                   d->cout_pad <= 128 && !d->res && !d->gn_stats && !d->splitk_ws,
               "hn_conv2d_bf16: a direct stem is a plain 1x1-over-patches convolution with cin = 256, cout_pad <= 128");
    if (stem_win) {
      // Window stem: `in` is the plain row-major canvas [n][H][W][4] bf16 (8 bytes per pixel), no frame.  For output pixel
      // (oy, ox) kernel row ky reads the 8 pixels 2*ox - 4 .. 2*ox + 3 of canvas row 2*oy - 3 + ky: 64 bytes that start 16
      // bytes after those of ox - 1.  In the UN-SWIZZLED K-major operand layout element (row r, 16-byte K chunk j) lives at
      // start + 16 * (r % 8) + SBO * (r / 8) + LBO * j; with SBO = 128 and LBO = 16 that is start + 16 * r + 16 * j, i.e. the
      // im2col matrix of a kernel row IS the canvas row (tools/umma_overlap_test.cu: tcgen05.mma reads overlapping core
      // matrices correctly).  One TMA box {256 pixels, 8 rows} = 8 row requests of 2 KB brings everything a tile of 125 output
      // pixels needs (the row-pair stem: 512 requests of 128 B and 4x the bytes; it was bound by exactly that); borders are
      // TMA zero fill.  Accumulator rows 125..127 of a tile read into the next kernel row and are not stored.
      HN_REQUIRE(d->cout_pad == 64 && d->cout == 64, "hn_conv2d_bf16: the window stem has 64 output channels");
      HN_REQUIRE((d->stem_pitch_h + 1) / 2 == d->h && (d->stem_pitch_w + 1) / 2 == d->w && d->stem_pitch_w % 2 == 0 &&
                     reinterpret_cast<uintptr_t>(d->in) % 16 == 0,
                 "hn_conv2d_bf16: window stem: canvas %dx%d (even width, 16-byte aligned) does not give a %dx%d output",
                 d->stem_pitch_h, d->stem_pitch_w, d->h, d->w);
      p.stem_win = 1;
      p.stem_tile_w = 125;
    } else {
      HN_REQUIRE(d->stem_pitch_h >= 2 * d->h + 6 && d->stem_pitch_h % 2 == 0 && d->stem_pitch_w >= 2 * d->w + 8,
                 "hn_conv2d_bf16: stem frame %dx%d too small for a %dx%d output (needs >= %dx%d, even height)",
                 d->stem_pitch_h, d->stem_pitch_w, d->h, d->w, 2 * d->h + 6, 2 * d->w + 8);
      p.stem_tile_w = BLOCK_M;
    }
    p.stem_tpr = hn_div_up(d->w, p.stem_tile_w);
    p.stem_h = d->h;
    p.m_tiles = d->n * d->h * p.stem_tpr;
  }
  const int sched_m_tiles = total_m_tiles > 0 ? total_m_tiles : p.m_tiles;     // tiles the launch spreads over the SMs
  int bn = stem_win ? 64 : force_bn ? force_bn
                    : (d->block_n ? d->block_n
                                  : pick_block_n(d->cout_pad, sched_m_tiles, p.num_taps * p.cin_chunks, d->gn_stats ? 32 : 16));
  HN_REQUIRE((bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256) && d->cout_pad % bn == 0,
             "hn_conv2d_bf16: block_n=%d does not divide cout_pad=%d", bn, d->cout_pad);
  p.n_tiles = d->cout_pad / bn;
  // (CTA pairs that multicast the weight tile were measured in round 1: +-1 %; not built any more)
  HN_REQUIRE(d->cluster == 0 || d->cluster == 1, "hn_conv2d_bf16: cluster=%d is not supported (CTA-pair multicast was removed)",
             d->cluster);
  // resident weights: the layer's whole weight slice stays in shared memory and only A boxes stream, when it is one
  // narrow N tile whose weights fit next to >= 4 A slots and there are enough tiles per CTA to amortise the load
  const int k_blocks_total = p.num_taps * p.cin_chunks;
  const long long b_bytes = (long long)k_blocks_total * bn * BLOCK_K * 2;
  bool rb = (p.n_tiles == 1 && bn <= 64 && b_bytes <= PIPE_BYTES_MAX - 4 * A_SLOT_BYTES &&
             sched_m_tiles >= 2 * hn_num_sms() && !(d->debug & 16) && !stem) || stem_win;
  const bool uni = bn <= 128 && !rb;      // unified stages (must match the kernel's constexpr UNI)
  const bool rb3 = rb && d->kh == 3 && d->stride == 1 && !(d->debug & 32) &&
                   PIPE_BYTES_MAX - b_bytes >= 2 * 3 * A_SLOT_BYTES && p.rows > 2 * d->dilation * p.wp;
  // A-box groups.  Pixel (oh*stride + dr, ow*stride + ds) of tap (r, s), dr = (r - kh/2)*dil, ds = (s - kw/2)*dil, is row
  // m + shift of the (phase) matrix; taps of one kernel row whose shifts differ by a few rows share one box.
  int box_rows = BLOCK_M, a_planes = 1, b_tiles = 1;
  p.n_groups = 0;
  p.uni_chunk_step = 1;
  auto add_group = [&](int phase, int shift, int ntaps, const int* taps, const int* offs) {
    const int g = p.n_groups++;
    const int off_step = ntaps > 1 ? offs[1] - offs[0] : 0, tap_step = ntaps > 1 ? taps[1] - taps[0] : 0;
    p.grp_shift[g] = shift;
    p.grp_info[g] = ntaps | (off_step << 4) | (taps[0] << 8) | (tap_step << 16) | (phase << 24);
    if (off_step > 0) box_rows = A_BOX_ROWS_MAX;
  };
  auto unit = [](int plane, int rowoff, int tile) { return plane | (rowoff << 2) | (tile << 6); };
  if (d->kh == 1) {
    const int taps[1] = {0}, offs[1] = {0};
    add_group(0, 0, 1, taps, offs);
    if (uni && p.cin_chunks >= 2) {
      // 1x1: a k-step covers two 64-channel chunks (two A planes, two weight tiles): four MMAs per TMA operation
      p.uni_chunk_step = 2;
      a_planes = b_tiles = 2;
      p.grp_info[0] = 2;
      p.grp_units[0] = unit(0, 0, 0) | (unit(1, 0, 1) << 8);
    } else {
      p.grp_units[0] = unit(0, 0, 0);
    }
  } else {
    for (int r = 0; r < 3; ++r) {
      const int dr = (r - 1) * d->dilation;
      if (rb3) {
        if (r == 0) {                               // one group: the box spans the three kernel rows
          const int taps[3] = {0, 1, 2};
          const int offs[3] = {0, d->dilation, 2 * d->dilation};
          add_group(0, dr * p.wp - d->dilation, 3, taps, offs);
        }
      } else if (d->stride == 1) {
        const int taps[3] = {r * 3, r * 3 + 1, r * 3 + 2};
        const int offs[3] = {0, d->dilation, 2 * d->dilation};
        add_group(0, dr * p.wp - d->dilation, 3, taps, offs);
        p.grp_units[p.n_groups - 1] = unit(0, 0, 0) | (unit(0, d->dilation, 1) << 8) | (unit(0, 2 * d->dilation, 2) << 16);
        b_tiles = 3;
      } else if (uni) {
        // input pixel (2*oh + dr, 2*ow + ds) lives in phase (dr&1, ds&1) at (oh + floor(dr/2), ow + floor(ds/2)).  One
        // box over both column phases of row phase pr, starting one row early: ds = -1 -> plane 1 row 0, ds = 0 ->
        // plane 0 row 1, ds = +1 -> plane 1 row 1
        const int pr = dr & 1, fr = (dr - pr) / 2;
        const int taps[3] = {r * 3, r * 3 + 1, r * 3 + 2};
        const int offs[3] = {0, 1, 1};
        add_group(pr * 2, fr * p.wp - 1, 3, taps, offs);
        box_rows = A_BOX_ROWS_MAX;
        p.grp_units[p.n_groups - 1] = unit(1, 0, 0) | (unit(0, 1, 1) << 8) | (unit(1, 1, 2) << 16);
        a_planes = 2;
        b_tiles = 3;
      } else {
        // ds = -1 -> column phase 1 at ow-1, ds = +1 -> column phase 1 at ow (one box), ds = 0 -> column phase 0
        const int pr = dr & 1, fr = (dr - pr) / 2;
        const int taps_odd[2] = {r * 3, r * 3 + 2}, offs_odd[2] = {0, 1};
        add_group(pr * 2 + 1, fr * p.wp - 1, 2, taps_odd, offs_odd);
        const int taps_even[1] = {r * 3 + 1}, offs_even[1] = {0};
        add_group(pr * 2, fr * p.wp, 1, taps_even, offs_even);
      }
    }
  }
  p.a_box_bytes = stem_win ? 8 * 2048 : box_rows * BLOCK_K * 2;
  p.k_steps = stem_win ? 1 : (uni && p.uni_chunk_step > 1) ? hn_div_up(p.cin_chunks, p.uni_chunk_step) : p.n_groups * p.cin_chunks;
  if (uni) {
    // duo: two M tiles per work item behind one weight box (see ConvParams::duo)
    p.duo = (duo_enabled() && d->kh == 3 && d->stride == 1 && p.n_tiles == 1 && total_m_tiles == 0 && a_planes == 1 &&
             Cfg<128>::NBUF >= 2 && bn >= 64 && p.m_tiles >= 2 * hn_num_sms() && !(d->debug & 262144)) ? 1 : 0;
    if (p.duo) a_planes = 2;
    p.uni_plane_bytes = box_rows * BLOCK_K * 2;
    p.uni_a_bytes = a_planes * p.uni_plane_bytes;
    p.uni_b_bytes = b_tiles * bn * BLOCK_K * 2;
    p.uni_stride = p.uni_a_bytes + p.uni_b_bytes;
    int stages = PIPE_BYTES_MAX / p.uni_stride;
    p.uni_stages = stages > MAX_STAGES ? MAX_STAGES : stages;
    HN_REQUIRE(p.uni_stages >= 2 && p.uni_a_bytes + p.uni_b_bytes <= p.uni_stride,
               "hn_conv2d_bf16: internal: unified stage too large");
    p.na_stages = p.uni_stages;
    p.nb_stages = 1;
    p.uni_a_rank4 = p.uni_chunk_step > 1 ? 1 : 0;
  } else if (rb) {
    p.rb_b_bytes = (int)b_bytes;
    p.rb3 = rb3 ? 1 : 0;
    if (rb3) p.a_box_bytes = 3 * A_SLOT_BYTES;
    const int slots = (PIPE_BYTES_MAX - p.rb_b_bytes) / (rb3 ? 3 * A_SLOT_BYTES : A_SLOT_BYTES);
    p.na_stages = slots > MAX_STAGES ? MAX_STAGES : slots;
    p.nb_stages = 1;
  } else {
    p.na_stages = Cfg<256>::NA;
    p.nb_stages = Cfg<256>::NB;
  }
  p.pipe_bytes = uni ? p.uni_stages * p.uni_stride
                     : (rb ? p.rb_b_bytes + p.na_stages * (rb3 ? 3 * A_SLOT_BYTES : A_SLOT_BYTES)
                           : p.na_stages * A_SLOT_BYTES + p.nb_stages * bn * BLOCK_K * 2);
  HN_REQUIRE(p.pipe_bytes <= PIPE_BYTES_MAX, "hn_conv2d_bf16: internal: operand area %d > %d", p.pipe_bytes, PIPE_BYTES_MAX);
  p.cout = d->cout;
  p.cout_pad = d->cout_pad;
  p.scale = d->scale;
  p.shift = d->shift;
  p.relu_lo = d->relu_lo;
  p.relu_hi = d->relu_hi;
  HN_REQUIRE(bn == 16 || d->relu_hi <= d->relu_lo || (d->relu_lo % 32 == 0 && (d->relu_hi % 32 == 0 || d->relu_hi >= d->cout)),
             "hn_conv2d_bf16: a ReLU range [%d, %d) that cuts a 32-column chunk needs block_n = 16", d->relu_lo, d->relu_hi);
  p.res = reinterpret_cast<const __nv_bfloat16*>(d->res);
  p.res_mode = d->res ? d->res_mode : 0;
  if (p.res_mode) {
    HN_REQUIRE(p.res_mode == 1 || p.res_mode == 2, "hn_conv2d_bf16: res_mode %d", p.res_mode);
    p.res_halo = d->res_halo;
    p.res_hp = d->res_h + 2 * d->res_halo;
    p.res_wp = d->res_w + 2 * d->res_halo;
    if (p.res_mode == 1)
      HN_REQUIRE(d->res_h == d->h && d->res_w == d->w, "hn_conv2d_bf16: residual size mismatch");
    else
      HN_REQUIRE(d->res_h == (d->h + 1) / 2 && d->res_w == (d->w + 1) / 2, "hn_conv2d_bf16: upsample residual size");
  }
  p.out = d->out;
  p.out_kind = d->out_kind;
  p.out_halo = d->out_halo;
  p.out_hp = d->h + 2 * d->out_halo;
  p.out_wp = d->w + 2 * d->out_halo;
  p.out_rows_per_image = d->out_rows_per_image;
  p.out_row_offset = d->out_row_offset;
  p.out_ld = d->out_ld;
  p.out_transpose_hw = d->out_transpose_hw;
  HN_REQUIRE(d->out_kind >= 0 && d->out_kind <= 2, "hn_conv2d_bf16: out_kind %d", d->out_kind);
  if (d->out_kind == 1) HN_REQUIRE(d->out_ld >= d->cout, "hn_conv2d_bf16: out_ld < cout");
  if (d->out_kind == 2)
    HN_REQUIRE(d->out_ld >= d->cout && d->out_rows_per_image >= d->out_row_offset + d->h * d->w && !d->out_transpose_hw,
               "hn_conv2d_bf16: planar output needs out_ld (planes per image) >= cout and rows_per_image >= offset + h*w");
  p.out_phase = reinterpret_cast<__nv_bfloat16*>(d->out_phase);
  if (p.out_phase) {
    HN_REQUIRE(d->out_kind == 0, "hn_conv2d_bf16: phase-split copy only for bf16 outputs");
    p.ph_halo = d->out_phase_halo;
    p.ph_hp = (d->h + 1) / 2 + 2 * d->out_phase_halo;
    p.ph_wp = (d->w + 1) / 2 + 2 * d->out_phase_halo;
    p.ph_stride = (long long)d->n * p.ph_hp * p.ph_wp * d->cout;
  }
  p.vec32 = (d->cout % 16 == 0) && (reinterpret_cast<uintptr_t>(d->out) % 32 == 0) &&
            (reinterpret_cast<uintptr_t>(d->res) % 32 == 0) && (reinterpret_cast<uintptr_t>(d->out_phase) % 32 == 0);
  p.trace = reinterpret_cast<long long*>(d->trace);
  p.dbg_flags = (d->debug >> 6) & 255;     // bit0: no epilogue stores, bit1: no epilogue work at all (timing experiments)
  p.gn_stats = reinterpret_cast<unsigned long long*>(d->gn_stats);
  if (p.gn_stats) {
    HN_REQUIRE(d->gn_groups > 0 && d->cout % d->gn_groups == 0, "hn_conv2d_bf16: gn_groups");
    p.gn_groups = d->gn_groups;
    p.gn_group_size = d->cout / d->gn_groups;
    HN_REQUIRE(bn >= 32 && (p.gn_group_size == 8 || p.gn_group_size == 16 || p.gn_group_size == 32),
               "hn_conv2d_bf16: GroupNorm group size %d unsupported (8, 16 or 32 channels per group)",
               p.gn_group_size);
  }

  // split-K (needs caller-provided scratch): short, deep layers whose tiles cannot fill the GPU
  p.splits = 1;
  if (d->splitk_ws && d->splitk_counters && bn >= 32 && !total_m_tiles) {
    const int k_steps = p.k_steps;
    const int tiles = p.m_tiles * p.n_tiles;
    int sp = d->splits;
    if (sp <= 0) {
      sp = 1;
      // measured (tools/a2j_timing.py): the reduction + counter round trip + read-back cost ~3-4 us per layer, a
      // k-block ~0.2 us, so only deep layers (>= 48 k-blocks) gain, with at least 12 k-blocks per split
      if (tiles * 2 <= hn_num_sms() && k_blocks_total >= split_min_kb()) {
        sp = hn_num_sms() / tiles;
        if (sp > k_blocks_total / 12) sp = k_blocks_total / 12;
        if (sp > 16) sp = 16;
        if (sp < 1) sp = 1;
      }
    }
    if (sp > k_steps) sp = k_steps;
    // one fp32 slice of the whole (padded) output per split: use as many splits as the caller's scratch holds
    const long long slice_bytes = (long long)p.m_tiles * BLOCK_M * d->cout_pad * 4;
    if (d->splits > 0) {
      HN_REQUIRE(d->splitk_ws_bytes >= sp * slice_bytes, "hn_conv2d_bf16: split-K scratch too small (%lld < %lld bytes)",
                 (long long)d->splitk_ws_bytes, sp * slice_bytes);
    } else if (sp * slice_bytes > d->splitk_ws_bytes) {
      sp = (int)(d->splitk_ws_bytes / slice_bytes);
    }
    if (sp > 1) {
      HN_REQUIRE(d->splitk_counters_len >= tiles, "hn_conv2d_bf16: split-K counter array too small");
      p.splits = sp;
      p.sk_ws = reinterpret_cast<float*>(d->splitk_ws);
      p.sk_slice = slice_bytes / 4;
      p.sk_cnt = reinterpret_cast<unsigned*>(d->splitk_counters);
      p.sk_ld = d->cout_pad;
    }
  }
  HN_REQUIRE(!(p.duo && p.splits > 1), "hn_conv2d_bf16: internal: duo items with split-K");
  p.epi_alt = (p.n_tiles == 1 && p.splits == 1 && !p.duo && !(d->debug & 32768) &&
               ((d->debug & 65536) || bn <= epi_alt_max_bn())) ? 1 : 0;
  // CTA pairs for the 256-wide layers that take the FAST epilogue: every CTA then holds half of each weight tile, so the
  // B ring is twice as deep in the same shared memory
  const bool fast = p.out_kind == 0 && p.vec32 != 0 && (p.cout % 32) == 0 && p.cout_pad == p.cout && p.splits == 1 &&
                    !(p.dbg_flags & 64);
  const bool pair = pair_enabled() && bn == 256 && !rb && !uni && fast && !(d->debug & 131072) &&
                    (total_m_tiles > 0 ? total_m_tiles : p.m_tiles) >= 2;
  if (pair) {
    const int nb = (PIPE_BYTES_MAX - p.na_stages * A_SLOT_BYTES) / ((bn / 2) * BLOCK_K * 2);
    p.nb_stages = nb > MAX_STAGES ? MAX_STAGES : nb;
    p.pipe_bytes = p.na_stages * A_SLOT_BYTES + p.nb_stages * (bn / 2) * BLOCK_K * 2;
  }
  out->pair = pair;
  CUtensorMap& ta = out->ta;
  CUtensorMap& tb = out->tb;
  if (stem_win) {
    // one element = one pixel (4 bf16 channels = 8 bytes): dims (W, H, n), box {256 pixels, 8 rows, 1}, dense in shared memory
    const cuuint64_t pw = (cuuint64_t)d->stem_pitch_w, ph = (cuuint64_t)d->stem_pitch_h;
    const cuuint64_t dims[3] = {pw, ph, (cuuint64_t)d->n};
    const cuuint64_t strides[2] = {pw * 8, ph * pw * 8};
    const cuuint32_t box[3] = {256, 8, 1};
    int rc = make_map(&ta, d->in, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_DATA_TYPE_UINT64);
    if (rc) return rc;
  } else if (stem) {
    // Patch view of the zero-framed canvas.  The frame stores its rows in PAIRS, [n][ph/2][pw][2 rows][4 ch] (16 bytes per
    // column), so that kernel rows 2j and 2j+1 of output pixel (oy, ox) -- 8 pixels x 2 rows x 4 channels = 64 elements
    // -- are ONE contiguous 128-byte run starting at row pair oy + j, column 2*ox:
    //   dim0 64 elements | dim1 ox (two columns = 32 bytes: the runs of consecutive ox overlap) | dim2 row pair | dim3 image
    // box {64, 128, 2, 1} = two ordinary [128 output pixels][128 bytes] SWIZZLE_128B tiles = two k-blocks of the GEMM per
    // k-step.  (Until round 1 session 3 the frame was row-major and a kernel row a 64-byte run: SWIZZLE_64B tiles, twice
    // the TMA row requests and operand reads at about half the rate -- tools/stem_ablation.py.)
    const cuuint64_t pw = (cuuint64_t)d->stem_pitch_w, ph2 = (cuuint64_t)d->stem_pitch_h / 2;
    const cuuint64_t dims[4] = {64, (cuuint64_t)d->w, ph2, (cuuint64_t)d->n};
    const cuuint64_t strides[3] = {32, pw * 16, ph2 * pw * 16};
    const cuuint32_t box[4] = {64, BLOCK_M, 2, 1};
    int rc = make_map(&ta, d->in, 4, dims, strides, box);
    if (rc) return rc;
  } else if (uni && p.uni_a_rank4) {
    // 1x1 with two chunks per k-step: dims (64 channels, rows, chunks, phases); the box puts the two chunk planes one
    // after the other in shared memory.  (The chunk stride is smaller than the row stride; if a driver refuses that,
    // fall back to one chunk per k-step.)
    const cuuint64_t dims[4] = {(cuuint64_t)BLOCK_K, (cuuint64_t)p.rows, (cuuint64_t)p.cin_chunks, (cuuint64_t)d->in_phases};
    const cuuint64_t strides[3] = {(cuuint64_t)d->cin * 2, (cuuint64_t)BLOCK_K * 2, (cuuint64_t)p.rows * d->cin * 2};
    const cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)box_rows, (cuuint32_t)a_planes, 1};
    if (make_map(&ta, d->in, 4, dims, strides, box) != HN_OK) {
      p.uni_chunk_step = 1;
      p.uni_a_rank4 = 0;
      a_planes = b_tiles = 1;
      p.grp_info[0] = 1;
      p.grp_units[0] = 0;
      p.k_steps = p.cin_chunks;
      p.uni_a_bytes = p.uni_plane_bytes;
      p.uni_b_bytes = bn * BLOCK_K * 2;
      p.uni_stride = p.uni_a_bytes + p.uni_b_bytes;
      const int stages = PIPE_BYTES_MAX / p.uni_stride;
      p.uni_stages = p.na_stages = stages > MAX_STAGES ? MAX_STAGES : stages;
      if (p.splits > p.k_steps) p.splits = p.k_steps;
    }
  }
  if (!stem && !(uni && p.uni_a_rank4)) {
    // dims (channels, rows, phases); unified stride-2 3x3: the box spans both column phases of a row phase.
    // rb3: the third dimension steps by one (dilated) image row instead -- an overlapping view of the same matrix;
    // its row count is cut by two image rows so that row + 2 * wp stays inside the allocation (the rows cut off are the
    // last image's bottom halo, which only halo outputs of kernel row 0 would read)
    const cuuint64_t dims[3] = {(cuuint64_t)d->cin, (cuuint64_t)(rb3 ? p.rows - 2 * d->dilation * p.wp : p.rows),
                                (cuuint64_t)(rb3 ? 3 : d->in_phases)};
    const cuuint64_t strides[2] = {(cuuint64_t)d->cin * 2,
                                   rb3 ? (cuuint64_t)d->dilation * p.wp * d->cin * 2 : (cuuint64_t)p.rows * d->cin * 2};
    const cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)box_rows, (cuuint32_t)(rb3 ? 3 : ((uni && !p.duo) ? a_planes : 1))};   // (duo: the second plane is the next M tile's box)
    int rc = make_map(&ta, d->in, 3, dims, strides, box);
    if (rc) return rc;
  }
  if (uni) {
    // [tap][chunk][cout_pad][64]: the box brings the weight tiles of one k-step (three taps of a chunk, or two chunks
    // of a 1x1) as consecutive BN x 128-byte tiles
    const cuuint64_t dims[4] = {(cuuint64_t)BLOCK_K, (cuuint64_t)d->cout_pad, (cuuint64_t)p.cin_chunks, (cuuint64_t)p.num_taps};
    const cuuint64_t strides[3] = {(cuuint64_t)BLOCK_K * 2, (cuuint64_t)d->cout_pad * BLOCK_K * 2,
                                   (cuuint64_t)p.cin_chunks * d->cout_pad * BLOCK_K * 2};
    const cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)bn, (cuuint32_t)(p.uni_chunk_step > 1 ? b_tiles : 1),
                               (cuuint32_t)(p.uni_chunk_step > 1 ? 1 : b_tiles)};
    int rc = make_map(&tb, d->weight, 4, dims, strides, box);
    if (rc) return rc;
  } else {
    // [k_block][cout_pad][64] bf16: a 2-D matrix of 128-byte rows, row = k_block * cout_pad + n
    const cuuint64_t dims[2] = {(cuuint64_t)BLOCK_K, (cuuint64_t)k_blocks_total * d->cout_pad};
    const cuuint64_t strides[1] = {(cuuint64_t)BLOCK_K * 2};
    const cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)(pair ? bn / 2 : bn)};    // pairs: each CTA fetches half of the tile
    int rc = make_map(&tb, d->weight, 2, dims, strides, box);
    if (rc) return rc;
  }
  out->p = p;
  out->bn = bn;
  out->rb = rb;
  return HN_OK;
}

// Pick the instantiation: operand pipeline by tile width / resident weights, FAST epilogue when the layer qualifies.
bool epi_fast_ok(const ConvParams& p) {
  // bf16 output, cout a multiple of the chunk width (no ragged chunks) and no padded N tile (the FAST body does not
  // skip chunks beyond cout: with cout_pad > cout they would land on the next pixel), 32-byte aligned rows, no split-K
  return p.out_kind == 0 && p.vec32 != 0 && (p.cout % 32) == 0 && p.cout_pad == p.cout && p.splits == 1 &&
         !(p.dbg_flags & 64);
}

template <bool SEG>
int dispatch(const BuiltConv& b, const CUtensorMap* ta, cudaStream_t st) {
  const ConvParams& p = b.p;
  const CUtensorMap& tb = b.tb;
  const bool fast = epi_fast_ok(p);
  if constexpr (SEG) {
    // segment launches: the 256-wide tower layers (FAST) and the 16-wide output convolutions (fp32 rows)
    if (b.bn == 256 && fast && b.pair) return launch<256, PIPE_RING, true, true, true>(ta, tb, p, st);
    if (b.bn == 256 && fast) return launch<256, PIPE_RING, true, true>(ta, tb, p, st);
    if (b.bn == 16 && b.rb && !fast) return launch<16, PIPE_RB, false, true>(ta, tb, p, st);
    hn_set_error("hn_conv2d_bf16_levels: unsupported combination (block_n=%d resident=%d fast=%d); supported: 256-wide bf16 "
                 "layers with cout %% 32 == 0 and 16-wide fp32-row outputs", b.bn, (int)b.rb, (int)fast);
    return HN_ERR_ARG;
  } else {
    if (b.rb) {
      switch (b.bn) {
        case 64: return fast ? launch<64, PIPE_RB, true, false>(ta, tb, p, st) : launch<64, PIPE_RB, false, false>(ta, tb, p, st);
        case 32: return fast ? launch<32, PIPE_RB, true, false>(ta, tb, p, st) : launch<32, PIPE_RB, false, false>(ta, tb, p, st);
        default: return launch<16, PIPE_RB, false, false>(ta, tb, p, st);
      }
    }
    switch (b.bn) {
      case 256:
        if (fast && b.pair) return launch<256, PIPE_RING, true, false, true>(ta, tb, p, st);
        return fast ? launch<256, PIPE_RING, true, false>(ta, tb, p, st) : launch<256, PIPE_RING, false, false>(ta, tb, p, st);
      case 128: return fast ? launch<128, PIPE_UNI, true, false>(ta, tb, p, st) : launch<128, PIPE_UNI, false, false>(ta, tb, p, st);
      case 64: return fast ? launch<64, PIPE_UNI, true, false>(ta, tb, p, st) : launch<64, PIPE_UNI, false, false>(ta, tb, p, st);
      case 32: return fast ? launch<32, PIPE_UNI, true, false>(ta, tb, p, st) : launch<32, PIPE_UNI, false, false>(ta, tb, p, st);
      default: return launch<16, PIPE_UNI, false, false>(ta, tb, p, st);
    }
  }
}

}  // namespace

extern "C" int hn_conv2d_bf16(const hn_conv_desc* d, void* stream) {
  BuiltConv b;
  int rc = build_conv(d, 0, 0, &b);
  if (rc) return rc;
  b.p.n_seg = 1;
  const CUtensorMap ta[3] = {b.ta, b.ta, b.ta};
  return dispatch<false>(b, ta, reinterpret_cast<cudaStream_t>(stream));
}

// One launch over several pyramid levels that share the convolution's weights (see the header).
extern "C" int hn_conv2d_bf16_levels(const hn_conv_desc* descs, int n_levels, void* stream) {
  HN_REQUIRE(descs && n_levels >= 1 && n_levels <= MAX_SEGS, "hn_conv2d_bf16_levels: 1..%d levels (got %d)", MAX_SEGS, n_levels);
  if (n_levels == 1) return hn_conv2d_bf16(descs, stream);
  int total_tiles = 0;
  for (int i = 0; i < n_levels; ++i) {
    const hn_conv_desc& d = descs[i];
    const hn_conv_desc& d0 = descs[0];
    HN_REQUIRE(d.in && d.out, "hn_conv2d_bf16_levels: level %d: null pointer", i);
    HN_REQUIRE(d.weight == d0.weight && d.scale == d0.scale && d.shift == d0.shift && d.cin == d0.cin && d.cout == d0.cout &&
                   d.cout_pad == d0.cout_pad && d.kh == d0.kh && d.kw == d0.kw && d.dilation == d0.dilation &&
                   d.relu_lo == d0.relu_lo && d.relu_hi == d0.relu_hi && d.halo_in == d0.halo_in && d.out_kind == d0.out_kind &&
                   d.out_halo == d0.out_halo && d.out_ld == d0.out_ld && d.out_rows_per_image == d0.out_rows_per_image &&
                   d.out_transpose_hw == d0.out_transpose_hw && d.gn_groups == d0.gn_groups &&
                   (d.gn_stats != nullptr) == (d0.gn_stats != nullptr) && d.block_n == d0.block_n && d.debug == d0.debug,
               "hn_conv2d_bf16_levels: level %d differs from level 0 in more than its geometry and buffers", i);
    HN_REQUIRE(d.stride == 1 && d.in_phases == 1 && d.kh == 3 && !d.res && !d.out_phase && !d.splitk_ws && !d.stem_pitch_w,
               "hn_conv2d_bf16_levels: level %d: only plain 3x3 stride-1 convolutions without residual / phase copy / split-K", i);
    HN_REQUIRE(d.out_kind == 0 || d.out == d0.out, "hn_conv2d_bf16_levels: fp32-row outputs of all levels share one buffer");
    const long long rows = (long long)d.n * (d.h + 2 * d.halo_in) * (d.w + 2 * d.halo_in);
    total_tiles += (int)((rows + BLOCK_M - 1) / BLOCK_M);
  }
  // the multi-level instantiations: 256-wide bf16 layers (FAST epilogue) and 16-wide fp32-row outputs with resident weights;
  // anything else (narrow bf16 layers, too few tiles for resident weights) runs level by level -- same results
  const hn_conv_desc& d0 = descs[0];
  const int want_bn = d0.block_n ? d0.block_n : ((d0.out_kind == 0 && d0.cout_pad % 256 == 0) ? 256 : (d0.cout_pad == 16 ? 16 : 0));
  BuiltConv b0;
  int rc = want_bn ? build_conv(&descs[0], want_bn, total_tiles, &b0) : HN_OK;
  if (rc) return rc;
  const bool fused_ok = want_bn != 0 && ((b0.bn == 256 && epi_fast_ok(b0.p)) || (b0.bn == 16 && b0.rb && !epi_fast_ok(b0.p)));
  if (!fused_ok) {
    for (int i = 0; i < n_levels; ++i) {
      rc = hn_conv2d_bf16(&descs[i], stream);
      if (rc) return rc;
    }
    return HN_OK;
  }
  CUtensorMap ta[3] = {b0.ta, b0.ta, b0.ta};
  ConvParams& p = b0.p;
  p.n_seg = n_levels;
  int tile_begin = 0, gn_off = 0;
  for (int i = 0; i < n_levels; ++i) {
    BuiltConv bi;
    if (i > 0) {
      rc = build_conv(&descs[i], b0.bn, total_tiles, &bi);
      if (rc) return rc;
      HN_REQUIRE(bi.bn == b0.bn && bi.rb == b0.rb && bi.pair == b0.pair && bi.p.rb3 == p.rb3 && bi.p.n_groups == p.n_groups &&
                     bi.p.na_stages == p.na_stages && bi.p.a_box_bytes == p.a_box_bytes && bi.p.epi_alt == p.epi_alt &&
                     bi.p.vec32 == p.vec32,
                 "hn_conv2d_bf16_levels: level %d needs a different kernel configuration than level 0", i);
      ta[i] = bi.ta;
    }
    const ConvParams& q = i == 0 ? p : bi.p;
    SegGeo& sg = p.seg[i];
    sg.tile_begin = tile_begin;
    sg.n_img = q.n_img; sg.hp = q.hp; sg.wp = q.wp; sg.rows = q.rows;
    sg.div_img_mul = q.div_img_mul; sg.div_wp_mul = q.div_wp_mul; sg.div_img_sh = q.div_img_sh; sg.div_wp_sh = q.div_wp_sh;
    for (int g = 0; g < 3; ++g) sg.shift[g] = g < q.n_groups ? q.grp_shift[g] : 0;
    sg.out = q.out; sg.out_hp = q.out_hp; sg.out_wp = q.out_wp; sg.out_row_offset = q.out_row_offset;
    sg.gn_stats = q.gn_stats;
    sg.gn_off = gn_off;
    tile_begin += q.m_tiles;
    gn_off += q.gn_stats ? q.n_img * q.gn_groups * 2 : 0;
  }
  HN_REQUIRE(gn_off <= GN_SMEM_SUMS, "hn_conv2d_bf16_levels: GroupNorm accumulators of all levels (%d sums) exceed %d", gn_off,
             GN_SMEM_SUMS);
  p.m_tiles = tile_begin;
  return dispatch<true>(b0, ta, reinterpret_cast<cudaStream_t>(stream));
}

// Upper bound on the number of CTAs (= SMs) the following convolution launches use; 0 = all.  The runtime caps the
// latency-bound pose-net launches so that they fit next to the detector kernels of the next step (GraphedHandNet).
extern "C" int hn_conv_set_cta_cap(int max_ctas) {
  HN_REQUIRE(max_ctas >= 0, "hn_conv_set_cta_cap: negative cap");
  g_cta_cap = max_ctas;
  return HN_OK;
}

// Programmatic dependent launch for the following convolution launches: 1 (default) = a kernel's CTAs may become resident
// while its predecessor in the stream is still running (they wait inside the kernel); 0 = plain stream order.
extern "C" int hn_conv_set_pdl(int enabled) {
  g_pdl_off = enabled ? 0 : 1;
  return HN_OK;
}
