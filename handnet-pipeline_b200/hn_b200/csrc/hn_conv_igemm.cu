// Convolution as a shifted GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces every 3x3 / 1x1 nn.Conv2d (+ FrozenBN / BN / bias / residual / ReLU) the reference reaches through
// cuDNN: torchvision resnet34 + FPN (fcos_utils/fcos.py:476), the FCOS towers and output convs
// (fcos_utils/fcos.py:232-264, 352-371), a2j/resnet.py and the A2J towers (a2j/a2j.py:44-181).
//
// Data layout: activations are haloed NHWC bf16, so for tap (r,s) the A operand of the implicit GEMM is the
// activation matrix [rows = N*Hp*Wp][Cin] shifted by (r*Wp + s) rows: one 2-D TMA box per (tap, 64-channel
// chunk), out-of-range rows zero-filled by TMA.  Weights are [Cout_pad][taps*Cin] bf16 (K-major).
//
// One persistent CTA per SM, 320 threads:
//   warp 0      TMA producer   (two rings: A boxes of 128 or 136 rows x 64 ch, B tiles of BN rows x 64 k, SWIZZLE_128B)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16; fp32 accumulators in TMEM,
//                               double-buffered so the epilogue of tile i overlaps the main loop of tile i+1)
//   warps 2..9  epilogue       (tcgen05.ld -> scale/shift -> +residual -> ReLU -> bf16/fp32 store, GroupNorm partial
//                               sums, optional phase-split copy for a following stride-2 conv)
//
// A-box sharing: the three horizontal taps of a 3x3 kernel row read rows m0+shift-1 .. m0+shift+128 of the same
// matrix, so ONE 136-row box serves all three: their UMMA descriptors start 0, 1 and 2 rows (128 B each) into the
// box (a SWIZZLE_128B descriptor may start at any 128-byte row with base_offset 0, tools/desc_offset_experiment.py).
// That cuts the activation traffic L2 -> shared memory of a 3x3 convolution by 3x, which is what bounds the narrow
// layers: a tcgen05.mma with N <= 128 is limited by the 128 B/clk of shared-memory bandwidth that its operand reads
// share with the TMA writes (tools/sync_cost_bench.cu, tools/mma_issue_bench2.cu).
#include "hn_common.cuh"

#include <stdlib.h>

#include <vector>

namespace {

__device__ __forceinline__ void hn_epi_bar_sync();

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_BOX_ROWS_MAX = 136;                   // 128 + up to 8 rows of horizontal-tap slack
constexpr int A_SLOT_BYTES = A_BOX_ROWS_MAX * BLOCK_K * 2;   // 17 KiB (a multiple of 1024: slots keep swizzle alignment)
// Patch tiles (resident weights, 3x3 stride 1 on large maps): an M tile is a 16-row x 8-pixel patch of ONE image instead
// of 128 consecutive rows of the flattened matrix.  One 4-D TMA box {64 ch, 10 px, 18 rows} (22.5 KiB) then holds the
// activations of all nine taps -- 1.4x the tile's own pixels instead of the 3.2x of three 136-row boxes (the layer1
// convolutions are bound by L2 -> SM traffic).  In shared memory the box is 180 rows of 128 B; tap (t, s) reads it from
// row t*10 + s on with a stride of 10 rows (1280 B) between the 8-row groups of the UMMA descriptor.
constexpr int PATCH_H = 16, PATCH_W = 8;
constexpr int PATCH_BOX_BYTES = (PATCH_H + 2) * (PATCH_W + 2) * BLOCK_K * 2;     // 23 040
constexpr int PATCH_SLOT_BYTES = (PATCH_BOX_BYTES + 1023) / 1024 * 1024;         // 23 552
constexpr int EPI_WARPS = 8;                       // two per TMEM lane quarter; they split the column chunks
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int NUM_THREADS = 64 + EPI_THREADS;
constexpr int MAX_TAPS = 9;
constexpr int MAX_GROUP_TAPS = 3;
constexpr int GN_SMEM_FLOATS = 4096;   // [images][groups][2] fp32 partial sums kept per CTA (16 KiB)

struct ConvParams {
  // compute geometry (== input geometry)
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
  // direct 7x7/2 stem: an M tile is 128 consecutive output columns of one output row; the A tensor map is the 5-D
  // overlapping-stride patch view of the canvas (build_conv)
  int stem_tpr, stem_h;                     // tiles per output row (0 = ordinary convolution), output rows per image
  int patch_tx, patch_ty;                   // patch tiles per image row / column (0 = flattened M tiles)
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
  double* gn_stats;
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
  int dbg_flags;                    // bring-up / timing experiments (bit0 no stores, bit1 no epilogue, bit2 no MMA, bit3 no TMA)
  long long* trace;                 // bring-up: CTA 0 logs (clock64, tag) pairs per role, TRACE_EVENTS each
  const struct ConvDeps* deps;      // multi-convolution launches: tile-level dataflow dependencies (NULL otherwise)
};

// Dataflow synchronisation between the convolutions of one multi-convolution launch.  Every finished (m, n) output tile
// of a convolution adds 1 to done[m]; a consumer tile waits until the producer M tiles it reads (its own rows +- one
// image row for a 3x3, the same rows for a 1x1 or a residual) have all their N tiles.  No grid-wide barrier.
constexpr int MAX_DEPS = 3;
struct ConvDeps {
  unsigned* done;                       // this convolution's counters, one per M tile
  int n_deps;
  const unsigned* dep_done[MAX_DEPS];   // producers' counters
  int dep_need[MAX_DEPS];               // value of a producer counter that means "M tile complete" (its n_tiles)
  const short* dep_first[MAX_DEPS];     // per M tile of THIS convolution: first / last producer M tile read
  const short* dep_last[MAX_DEPS];
};

__device__ __forceinline__ void dep_wait(const ConvDeps* dp, int mt) {
  if ((threadIdx.x & 31) == 0) {
    const int nd = dp->n_deps;
    for (int d = 0; d < nd; ++d) {
      const int t0 = dp->dep_first[d][mt], t1 = dp->dep_last[d][mt];
      const unsigned need = (unsigned)dp->dep_need[d];
      const unsigned* cnt = dp->dep_done[d];
      for (int t = t0; t <= t1; ++t) {
        unsigned spins = 0;
        while (true) {
          unsigned v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt + t) : "memory");
          if (v >= need) break;
          if (++spins > (1u << 24)) {
            printf("hn: dependency wait timed out (block %d, producer tile %d: %u of %u)\n", (int)blockIdx.x, t, v, need);
            __trap();
          }
        }
      }
    }
  }
  __syncwarp();
  // the TMA (async proxy) reads that follow must observe what other CTAs wrote with ordinary stores
  asm volatile("fence.proxy.async;" ::: "memory");
}

constexpr int TRACE_EVENTS = 2048;
// role 0 producer, 1 MMA issuer, 2 first epilogue warp; written by lane 0 of CTA 0 only
__device__ __forceinline__ void hn_trace(long long* tr, int role, int& idx, int tag) {
  if (tr != nullptr && blockIdx.x == 0) {
    if ((threadIdx.x & 31) == 0 && idx < TRACE_EVENTS) {
      tr[(role * TRACE_EVENTS + idx) * 2] = clock64();
      tr[(role * TRACE_EVENTS + idx) * 2 + 1] = tag;
    }
    ++idx;
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

// CS = thread-block cluster size along M: the CS CTAs of a cluster work on CS consecutive M tiles of the same N
// tile in lock step; each loads 1/CS of the B (weight) tile and multicasts it to all of them, so the weights
// cross the L2 -> SM fabric once per cluster instead of once per CTA.
// ---------------------------------------------------------------------------------------------------------------
// One convolution's worth of work for the calling warp (role by warp index).  `first_tile` / `tile_stride` select this
// CTA's (super) tiles; `stage`, `phase`, `it` are the calling thread's pipeline state and persist across calls, so a
// persistent multi-layer kernel can chain convolutions through the same barriers and TMEM buffers.
// ---------------------------------------------------------------------------------------------------------------
template <bool V>
struct FastTag { static constexpr bool value = V; };

template <int N>
__device__ __forceinline__ void hn_tmem_ld_chunk(uint32_t taddr, uint32_t (&r)[N]) {
  if constexpr (N == 32) hn_tmem_ld32(taddr, r);
  else hn_tmem_ld16(taddr, r);
}

struct PipeState {      // per-thread pipeline state; persists across convolutions of a multi-convolution launch
  int a_stage, b_stage, it;
  uint32_t a_phase, b_phase;
};

template <int BN, int CS, bool RB>
__device__ __forceinline__ void conv_roles(const CUtensorMap* tm_a_ptr, const CUtensorMap* tm_b_ptr, const ConvParams& p,
                                           uint8_t* smem_hdr, const uint32_t tmem_base, const int first_tile,
                                           const int tile_stride, PipeState& ps, const bool pdl = false) {
  // pdl: launched with programmatic stream serialisation.  Each role executes griddepcontrol.wait itself, as late as
  // it can: the producer first fetches what does not depend on the previous kernel (the WEIGHTS of its first k-steps,
  // or the whole resident slice), the MMA warp never touches global memory and does not wait at all.
  using C = Cfg<BN>;
  const CUtensorMap& tm_a = *tm_a_ptr;
  const CUtensorMap& tm_b = *tm_b_ptr;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_hdr);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + MAX_STAGES;
  uint64_t* b_full_ring = bars + 2 * MAX_STAGES;
  uint64_t* b_empty = bars + 3 * MAX_STAGES;
  uint64_t* tmem_full = bars + 4 * MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 8;
  uint64_t* b_full = tmem_empty + 8;                     // resident weights have landed
  // everything the role loops need from the parameter block, read once (the asm statements in the loops clobber
  // memory, so anything left in `p` would be re-read from the constant bank / shared memory every iteration)
  const int cin_chunks = p.cin_chunks;
  const int n_tiles = p.n_tiles;
  const int splits = p.splits;
  const int na = p.na_stages, nb = p.nb_stages;
  const int a_box_bytes = p.a_box_bytes;
  const int cout_pad = p.cout_pad;
  const int dbg_flags = p.dbg_flags;
  long long* const trace = p.trace;
  const ConvDeps* const deps = p.deps;
  int tri = 0;
  const int k_steps = p.k_steps;                         // one k-step = one A box
  uint8_t* pipe = smem_hdr + HDR_PAD;                    // 1024-aligned operand area
  uint8_t* a_ring = pipe + (RB ? p.rb_b_bytes : 0);
  const int a_slot_bytes = (RB && p.patch_tx > 0) ? PATCH_SLOT_BYTES : ((RB && p.rb3) ? 3 * A_SLOT_BYTES : A_SLOT_BYTES);
  uint8_t* b_ring = a_ring + na * a_slot_bytes;
  int& it = ps.it;

  // warp index through a shuffle: tells the compiler it is warp-uniform, so the producer / MMA role loops below run
  // converged and their addresses and descriptors live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // tile schedule: "super tiles" of CS consecutive M tiles x one N tile, N fastest, strided over the clusters
  const int cta_rank = (CS > 1) ? (int)hn_cluster_ctarank() : 0;
  const int num_super = ((p.m_tiles + CS - 1) / CS) * n_tiles;
  const int num_items = num_super * splits;
  const int s_base = k_steps / splits, s_rem = k_steps - s_base * splits;

  // The producer and MMA loops are single-instruction-stream code on the kernel's critical path (a 64-wide tile has
  // only ~50 cycles of tensor work per MMA): no divisions or indexed parameter loads inside the k loop, descriptors
  // advance by additions.
  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // All 32 lanes walk the loop and poll the barriers; one elected lane issues the copies.
    int a_stage = ps.a_stage, b_stage = ps.b_stage;
    uint32_t a_phase = ps.a_phase, b_phase = ps.b_phase;
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
    constexpr bool UNI = (BN <= 128) && !RB;               // unified stages (see ConvParams)
    const int uni_stage_bytes = p.uni_a_bytes + p.uni_b_bytes, uni_chunk_step = p.uni_chunk_step;
    const int uni_stride = p.uni_stride;
    int early_b = 0;                                       // leading k-steps of an item whose weights are on the way
    int early_item = first_tile;                           // ... and which item that is
    if (pdl) {
      if constexpr (UNI) {
        if (first_tile < num_items && splits == 1 && !(dbg_flags & 8) && a_stage == 0) {
          const int nt0 = n_tiles > 1 ? first_tile % n_tiles : 0;
          early_b = k_steps < na ? k_steps : na;
          if (hn_elect_one()) {
            int g0 = 0, cc0 = 0;
            for (int e = 0; e < early_b; ++e) {
              hn_mbar_expect_tx(&a_full[e], (uint32_t)uni_stage_bytes);
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
    }
    for (int w_ = first_tile; w_ < num_items; w_ += tile_stride) {
      int st = w_, s_begin = 0, s_end = k_steps, g = 0, cc = 0;
      if (splits > 1) {                                    // (super) tile and K split of this work item
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
      const int m0 = (mt * CS + cta_rank) * BLOCK_M;
      const int n0 = nt * BN;
      int info = p.grp_info[g], shift = p.grp_shift[g];
      if constexpr (UNI) {
        if (deps != nullptr && splits == 1 && !(dbg_flags & 8)) {
          // multi-convolution launch: the weights do not depend on the producer tiles -- arm the first stages and fetch
          // their weight tiles BEFORE waiting for the dependencies, the activations after
          early_b = k_steps < na ? k_steps : na;
          early_item = w_;
          int es = a_stage, g0 = 0, cc0 = 0;
          uint32_t eph = a_phase;
          for (int e = 0; e < early_b; ++e) {
            hn_mbar_wait(&a_empty[es], eph ^ 1);
            if (hn_elect_one()) {
              hn_mbar_expect_tx(&a_full[es], (uint32_t)uni_stage_bytes);
              hn_tma_load_4d(a_ring + es * uni_stride + p.uni_a_bytes, &tm_b, &a_full[es], 0, n0, cc0,
                             (p.grp_info[g0] >> 8) & 255);
            }
            cc0 += uni_chunk_step;
            if (cc0 >= cin_chunks) { cc0 = 0; ++g0; }
            if (++es == na) { es = 0; eph ^= 1; }
          }
          __syncwarp();
        }
      }
      if (deps != nullptr) dep_wait(deps, mt);             // multi-convolution launch: the tiles this one reads are done
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
            if (dbg_flags & 8) {                             // timing experiment: no loads at all
              hn_mbar_arrive(&a_full[a_stage]);
            } else {
              uint8_t* sa = a_ring + a_stage * uni_stride;
              const bool b_pending = w_ == early_item && step - s_begin < early_b;   // armed and B issued before the wait
              if (!b_pending) hn_mbar_expect_tx(&a_full[a_stage], (uint32_t)uni_stage_bytes);
              if (p.stem_tpr > 0) hn_tma_load_4d(sa, &tm_a, &a_full[a_stage], 0, st_ox, st_oy + cc, st_n);
              else if (p.uni_a_rank4) hn_tma_load_4d(sa, &tm_a, &a_full[a_stage], 0, m0 + shift, cc, info >> 24);
              else hn_tma_load_3d(sa, &tm_a, &a_full[a_stage], cc * BLOCK_K, m0 + shift, info >> 24);
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
        continue;
      }
      int pt_n = 0, pt_y = 0, pt_x = 0;                      // patch tiles: image, first box row / column (padded coords)
      if (RB && p.patch_tx > 0) {
        const int per_img = p.patch_tx * p.patch_ty;
        pt_n = mt / per_img;
        const int rem = mt - pt_n * per_img, ty = rem / p.patch_tx;
        pt_y = ty * PATCH_H + p.halo - 1;
        pt_x = (rem - ty * p.patch_tx) * PATCH_W + p.halo - 1;
      }
      for (int step = s_begin; step < s_end; ++step) {
        const int ntaps = info & 15;
        // ---- the A box of this (group, chunk)
        hn_mbar_wait(&a_empty[a_stage], a_phase ^ 1);
        hn_trace(trace, 0, tri, 1);
        if (hn_elect_one()) {
          if (dbg_flags & 8) {                               // timing experiment: no loads at all
            hn_mbar_arrive(&a_full[a_stage]);
          } else {
            hn_mbar_expect_tx(&a_full[a_stage], (uint32_t)a_box_bytes);
            if (RB && p.patch_tx > 0)
              hn_tma_load_4d(a_ring + a_stage * a_slot_bytes, &tm_a, &a_full[a_stage], cc * BLOCK_K, pt_x, pt_y, pt_n);
            else
              hn_tma_load_3d(a_ring + a_stage * a_slot_bytes, &tm_a, &a_full[a_stage], cc * BLOCK_K, m0 + shift, info >> 24);
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
              if (dbg_flags & 8) {
                hn_mbar_arrive(&b_full_ring[b_stage]);
              } else {
                uint8_t* sb = b_ring + b_stage * C::B_STAGE_BYTES;
                hn_mbar_expect_tx(&b_full_ring[b_stage], (uint32_t)C::B_STAGE_BYTES);
                if constexpr (CS == 1) {
                  hn_tma_load_2d(sb, &tm_b, &b_full_ring[b_stage], 0, krow);
                } else {
                  constexpr int SLICE = BN / CS;   // weight rows this CTA fetches for the whole cluster
                  hn_tma_load_2d_mcast(sb + cta_rank * SLICE * BLOCK_K * 2, &tm_b, &b_full_ring[b_stage], 0,
                                       krow + cta_rank * SLICE, (uint16_t)((1u << CS) - 1u));
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
          if (step + 1 < s_end) { info = p.grp_info[g]; shift = p.grp_shift[g]; }
        }
      }
    }
    ps.a_stage = a_stage; ps.b_stage = b_stage; ps.a_phase = a_phase; ps.b_phase = b_phase;
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // Converged warp; tcgen05.mma / commit are issued by one elected lane.  (A second issuing warp taking alternate
    // tiles was tried for the narrow tiles: it needs a ring of two whole tiles to keep the parity waits unambiguous and
    // bought nothing once the epilogue ran -- layer1 52.2 -> 52.9 us -- so there is one issuer.)
    long long* const trace_m = trace;
    constexpr uint32_t idesc = hn_umma_idesc_bf16(BN);
    constexpr uint32_t A_SLOT_D = A_SLOT_BYTES >> 4, B_SLOT_D = C::B_STAGE_BYTES >> 4, ROW_D = (BLOCK_K * 2) >> 4;
    int a_stage = ps.a_stage, b_stage = ps.b_stage;
    uint32_t a_phase = ps.a_phase, b_phase = ps.b_phase;
    // low words of the shared-memory descriptors of slot 0 of each ring (start address >> 4 in bits 0..13); the high
    // word is the same for every operand tile
    const uint32_t a_desc0 = (uint32_t)hn_umma_smem_desc(hn_smem_u32(a_ring));
    const uint32_t b_desc0 = (uint32_t)hn_umma_smem_desc(hn_smem_u32(RB ? pipe : b_ring));
    const uint64_t desc_hi = hn_umma_smem_desc(0) & 0xffffffff00000000ull;
    if constexpr (RB) hn_mbar_wait(b_full, 0);
    // The tensor pipe queues only a few MMAs, so whatever the issuing thread does between two bursts must take less
    // than the burst it has just issued needs to execute.  Wide tiles (256 columns: 512 cycles per tap) poll and issue
    // tap by tap, so that a tap starts as soon as its weight tile has landed.  Narrow tiles have only ~50-64 cycles of
    // tensor work per MMA: they poll all barriers of a k-step first and issue its (up to) 12 MMAs and their commits as
    // one straight-line burst.
    constexpr bool STEP_ISSUE = BN <= 128;
    constexpr bool UNI = (BN <= 128) && !RB;               // unified stages: one barrier pair and one burst per k-step
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
      const int buf = it & (C::NBUF - 1);
      const uint32_t d_tmem = tmem_base + buf * BN;
      int info = p.grp_info[g];
      if (!(dbg_flags & 32)) hn_mbar_wait(&tmem_empty[buf], ((it >> C::NBUF_LOG) & 1) ^ 1);   // the epilogue has drained this buffer
      hn_trace(trace_m, 1, tri, 4);
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
          hn_trace(trace_m, 1, tri, 1);
          hn_tc_fence_after();
          if (hn_elect_one()) {
            if (p.stem_tpr > 0) {
              // direct stem: two [128 x 128 B] tiles (row pairs oy + 2*step, +1 = kernel rows 4*step .. +3) against the two
              // weight k-blocks of the step
              if (!(dbg_flags & 4)) {
                hn_umma_bf16_x4(d_tmem, desc_hi | sa, desc_hi | sb, idesc, accumulate);
                hn_umma_bf16_x4(d_tmem, desc_hi | (sa + (uint32_t)((BLOCK_M * BLOCK_K * 2) >> 4)), desc_hi | (sb + B_SLOT_D), idesc, 1u);
              }
            } else if (!(dbg_flags & 4)) {                     // (timing experiment: bit 2 skips the MMAs)
              hn_umma_bf16_x4(d_tmem, desc_hi | (sa + (units & 3) * uni_plane_d + ((units >> 2) & 15) * ROW_D),
                              desc_hi | (sb + ((units >> 6) & 3) * B_SLOT_D), idesc, accumulate);
              if (nu > 1)
                hn_umma_bf16_x4(d_tmem, desc_hi | (sa + ((units >> 8) & 3) * uni_plane_d + ((units >> 10) & 15) * ROW_D),
                                desc_hi | (sb + ((units >> 14) & 3) * B_SLOT_D), idesc, 1u);
              if (nu > 2)
                hn_umma_bf16_x4(d_tmem, desc_hi | (sa + ((units >> 16) & 3) * uni_plane_d + ((units >> 18) & 15) * ROW_D),
                                desc_hi | (sb + ((units >> 22) & 3) * B_SLOT_D), idesc, 1u);
            }
            hn_umma_commit_addr<1>(ea);                        // stage free once these MMAs have read it
            if (last) hn_umma_commit_addr<1>(tf);              // accumulator complete -> epilogue
          }
          accumulate = 1;
          hn_trace(trace_m, 1, tri, 3);
          if (++a_stage == na) { a_stage = 0; a_phase ^= 1; }
          cc += uni_chunk_step;
          if (cc >= cin_chunks) {
            cc = 0;
            ++g;
            if (!last) { info = p.grp_info[g]; units = p.grp_units[g]; }
          }
        }
        continue;
      }
      for (int step = s_begin; step < s_end; ++step) {
        const int ntaps = info & 15;
        const uint32_t off_step = ((info >> 4) & 15) * ROW_D;
        const uint32_t da_lo = a_desc0 + a_stage * A_SLOT_D;
        // resident weights: tile of tap (tap0 + t * tap_step), chunk cc
        const uint32_t db_rb = b_desc0 + (((info >> 8) & 255) * cin_chunks + cc) * B_SLOT_D;
        const uint32_t db_rb_step = ((info >> 16) & 255) * cin_chunks * B_SLOT_D;
        const bool last = step == s_end - 1;
        hn_mbar_wait(&a_full[a_stage], a_phase);
        hn_trace(trace_m, 1, tri, 1);
        if (RB && p.rb3) {
          // all nine taps of this chunk from one stage: A tile of kernel row t at slot + t * 17 KiB, tap (t, s) reads it
          // from row s * dil on; weight tile of tap t*3 + s, chunk cc
          const bool patch = p.patch_tx > 0;
          const uint32_t sa = a_desc0 + a_stage * (patch ? (uint32_t)(PATCH_SLOT_BYTES >> 4) : 3 * A_SLOT_D);
          const uint32_t sb = b_desc0 + cc * B_SLOT_D, tap_d = cin_chunks * B_SLOT_D;
          const uint32_t ea = hn_smem_u32(&a_empty[a_stage]), tf = hn_smem_u32(&tmem_full[buf]);
          // patch tiles: kernel row t starts t box rows (10 pixels) further, the 8-pixel groups are 10 rows apart
          const uint32_t t_step = patch ? (PATCH_W + 2) * ROW_D : A_SLOT_D, s_step = patch ? ROW_D : off_step;
          const uint64_t a_hi = patch ? ((desc_hi & ~(uint64_t(0x3FFF) << 32)) | (uint64_t(((PATCH_W + 2) * BLOCK_K * 2) >> 4) << 32))
                                      : desc_hi;
          hn_tc_fence_after();
          if (hn_elect_one()) {
            if (!(dbg_flags & 4)) {
#pragma unroll
              for (int t = 0; t < 3; ++t) {
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                  hn_umma_bf16_x4(d_tmem, a_hi | (sa + t * t_step + sx * s_step), desc_hi | (sb + (t * 3 + sx) * tap_d),
                                  idesc, (t | sx) ? 1u : accumulate);
                }
              }
            }
            hn_umma_commit_addr<1>(ea);
            if (last) hn_umma_commit_addr<1>(tf);
          }
        } else if constexpr (STEP_ISSUE) {
          // ring slots of the (up to) three weight tiles of this step
          int bs1 = b_stage + 1, bs2 = b_stage + 2;
          uint32_t bp1 = b_phase, bp2 = b_phase;
          if (bs1 >= nb) { bs1 -= nb; bp1 ^= 1; }
          if (bs2 >= nb) { bs2 -= nb; bp2 ^= 1; }
          uint32_t db0 = db_rb, db1 = db_rb + db_rb_step, db2 = db_rb + 2 * db_rb_step;
          if constexpr (!RB) {
            hn_mbar_wait(&b_full_ring[b_stage], b_phase);
            if (ntaps > 1) hn_mbar_wait(&b_full_ring[bs1], bp1);
            if (ntaps > 2) hn_mbar_wait(&b_full_ring[bs2], bp2);
            db0 = b_desc0 + b_stage * B_SLOT_D;
            db1 = b_desc0 + bs1 * B_SLOT_D;
            db2 = b_desc0 + bs2 * B_SLOT_D;
          }
          const uint32_t eb0 = hn_smem_u32(&b_empty[b_stage]), eb1 = hn_smem_u32(&b_empty[bs1]),
                         eb2 = hn_smem_u32(&b_empty[bs2]);
          const uint32_t ea = hn_smem_u32(&a_empty[a_stage]), tf = hn_smem_u32(&tmem_full[buf]);
          hn_tc_fence_after();
          if (hn_elect_one()) {
            const bool mma = !(dbg_flags & 4);                // (timing experiment: bit 2 skips the MMAs)
            if (mma) hn_umma_bf16_x4(d_tmem, desc_hi | da_lo, desc_hi | db0, idesc, accumulate);
            if constexpr (!RB) hn_umma_commit_addr<CS>(eb0);
            if (ntaps > 1) {
              if (mma) hn_umma_bf16_x4(d_tmem, desc_hi | (da_lo + off_step), desc_hi | db1, idesc, 1u);
              if constexpr (!RB) hn_umma_commit_addr<CS>(eb1);
            }
            if (ntaps > 2) {
              if (mma) hn_umma_bf16_x4(d_tmem, desc_hi | (da_lo + 2 * off_step), desc_hi | db2, idesc, 1u);
              if constexpr (!RB) hn_umma_commit_addr<CS>(eb2);
            }
            hn_umma_commit_addr<1>(ea);                       // A box free
            if (last) hn_umma_commit_addr<1>(tf);             // accumulator complete -> epilogue
          }
          if constexpr (!RB) {
            b_stage += ntaps;
            if (b_stage >= nb) { b_stage -= nb; b_phase ^= 1; }
          }
        } else {
          uint32_t da_t = da_lo, db_t = db_rb;
          for (int t = 0; t < ntaps; ++t) {
            if constexpr (!RB) {
              hn_mbar_wait(&b_full_ring[b_stage], b_phase);
              hn_trace(trace_m, 1, tri, 2);
            }
            hn_tc_fence_after();
            const uint32_t db_lo = RB ? db_t : b_desc0 + b_stage * B_SLOT_D;
            const uint32_t eb = hn_smem_u32(&b_empty[b_stage]);
            if (hn_elect_one()) {
              if (!(dbg_flags & 4)) hn_umma_bf16_x4(d_tmem, desc_hi | da_t, desc_hi | db_lo, idesc, accumulate);
              if constexpr (!RB) hn_umma_commit_addr<CS>(eb);   // weight slot free once these MMAs have read it
            }
            accumulate = 1;
            da_t += off_step;
            db_t += db_rb_step;
            if constexpr (!RB) {
              if (++b_stage == nb) { b_stage = 0; b_phase ^= 1; }
            }
          }
          const uint32_t ea = hn_smem_u32(&a_empty[a_stage]), tf = hn_smem_u32(&tmem_full[buf]);
          if (hn_elect_one()) {
            hn_umma_commit_addr<1>(ea);                         // A box free
            if (last) hn_umma_commit_addr<1>(tf);               // accumulator complete -> epilogue
          }
        }
        accumulate = 1;
        hn_trace(trace_m, 1, tri, 3);
        if (++a_stage == na) { a_stage = 0; a_phase ^= 1; }
        if (++cc == cin_chunks) {
          cc = 0;
          ++g;
          if (!last) info = p.grp_info[g];
        }
      }
    }
    ps.a_stage = a_stage; ps.b_stage = b_stage; ps.a_phase = a_phase; ps.b_phase = b_phase;
  } else {
    // ===================================== epilogue ==========================================
    const int quarter = warp & 3;              // TMEM lanes this warp may touch: 32*quarter .. +31
    const int half = (warp - 2) >> 2;          // which of the two warps of this quarter: takes every other chunk
    const int img_rows = p.hp * p.wp;
    // scale/shift of the current N tile live in shared memory (the L1 left next to ~210 KiB of smem is too small
    // to keep them, and an L2 round trip per chunk was the epilogue's critical path)
    float* ss_base = reinterpret_cast<float*>(smem_hdr + HDR_BARS + GN_SMEM_FLOATS * 4);
    // per-CTA GroupNorm accumulator [image][group][2] in shared memory (when it fits)
    float* gn_acc = reinterpret_cast<float*>(smem_hdr + HDR_BARS);
    const int gn_vals = p.gn_stats ? p.n_img * p.gn_groups * 2 : 0;
    const bool gn_smem = gn_vals > 0 && gn_vals <= GN_SMEM_FLOATS;
    if (gn_smem) {
      for (int i = threadIdx.x - 64; i < gn_vals; i += EPI_THREADS) gn_acc[i] = 0.f;
      hn_epi_bar_sync();
    }
    uint32_t* sk_flag = reinterpret_cast<uint32_t*>(b_full + 1) + 1;   // "this CTA finalises the tile" (after tmem_slot)
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");       // before the first global read / write of this role
    if (dbg_flags & 32) return;                   // experiment: no epilogue role at all (the MMA warp does not wait for it)
    const uint32_t div_img_mul = p.div_img_mul, div_wp_mul = p.div_wp_mul;
    const int div_img_sh = p.div_img_sh, div_wp_sh = p.div_wp_sh;
    const int rows = p.rows, wp = p.wp, halo = p.halo;
    int ss_n0[2] = {-1, -1};                    // N tile whose scale/shift each staging buffer holds
    const bool has_scale = p.scale != nullptr;
    // bf16 output, cout a multiple of the chunk width (no ragged chunks) and no padded N tile (the FAST body does not
    // skip chunks beyond cout: with cout_pad > cout they would land on the next pixel), 32-byte aligned rows, no split-K
    const bool epi_fast = p.out_kind == 0 && p.vec32 != 0 && (p.cout % 32) == 0 && p.cout_pad == p.cout && splits == 1 &&
                          !(p.dbg_flags & 64);
    // Alternate-tile mode (single N tile, no split-K): the epilogue of a tile is a latency chain (accumulator wait,
    // tcgen05.ld, residual / store round trips); with both warps of a quarter on the SAME tile nothing overlaps it.  Here
    // warps 2-5 drain the even tiles and warps 6-9 the odd ones, so two tiles are in flight.  Each warp arrives twice on
    // tmem_empty (the barrier counts eight arrivals per tile).  scale/shift are staged once for both buffers.
    const bool alt = p.epi_alt != 0;
    if (alt) {
      for (int i = threadIdx.x - 64; i < BN; i += EPI_THREADS) {
        const float sc = (p.scale && i < p.cout) ? __ldg(p.scale + i) : 1.0f;
        const float sh = (p.shift && i < p.cout) ? __ldg(p.shift + i) : 0.0f;
        ss_base[i] = sc; ss_base[256 + i] = sh; ss_base[512 + i] = sc; ss_base[768 + i] = sh;
      }
      ss_n0[0] = ss_n0[1] = 0;
      hn_epi_bar_sync();
    }
    for (int w_ = first_tile; w_ < num_items; w_ += tile_stride, ++it) {
      if (alt && (it & 1) != half) continue;
      const int st = splits > 1 ? w_ / splits : w_;
      if (warp == 2) hn_trace(trace, 2, tri, 1);
      const int buf = it & (C::NBUF - 1);
      const uint32_t acc_phase = (it >> C::NBUF_LOG) & 1;
      int mt = st, nt = 0;
      if (n_tiles > 1) { mt = st / n_tiles; nt = st - mt * n_tiles; }
      const int m0 = (mt * CS + cta_rank) * BLOCK_M;
      const int n0 = nt * BN;
      const int m = m0 + quarter * 32 + lane;
      // decode the padded pixel this accumulator row belongs to (divisions by multiply-high, see fastdiv())
      int img = 0, h = 0, w = 0;
      bool interior = false;
      if (p.patch_tx > 0) {                              // patch tile: accumulator row r = pixel (r / 8, r % 8) of the patch
        const int per_img = p.patch_tx * p.patch_ty;
        img = mt / per_img;
        const int rem = mt - img * per_img, ty = rem / p.patch_tx;
        const int r = quarter * 32 + lane;
        h = ty * PATCH_H + (r >> 3);
        w = (rem - ty * p.patch_tx) * PATCH_W + (r & 7);
        interior = img < p.n_img && h < p.hp - 2 * halo && w < wp - 2 * halo;
      } else if (p.stem_tpr > 0) {                       // stem tile: 128 output columns of one output row
        const int row = mt / p.stem_tpr;
        img = row / p.stem_h;
        h = row - img * p.stem_h;
        w = (mt - row * p.stem_tpr) * BLOCK_M + quarter * 32 + lane;
        interior = w < wp && img < p.n_img;
      } else if (m < rows) {
        img = (int)((__umulhi((uint32_t)m, div_img_mul) + (uint32_t)m) >> div_img_sh);
        const int rem = m - img * img_rows;
        const int hh = (int)((__umulhi((uint32_t)rem, div_wp_mul) + (uint32_t)rem) >> div_wp_sh);
        const int ww = rem - hh * wp;
        h = hh - halo;
        w = ww - halo;
        interior = (h >= 0) && (w >= 0) && (h < p.hp - 2 * halo) && (w < wp - 2 * halo);
      }
      const int H = p.hp - 2 * p.halo, W = p.wp - 2 * p.halo;
      // output / residual element offsets of channel 0 of this row
      long long out_off = 0, res_off = 0, ph_off = 0;
      if (interior) {
        if (p.out_kind == 0) {
          out_off = ((long long)(img * p.out_hp + h + p.out_halo) * p.out_wp + (w + p.out_halo)) * p.cout;
        } else {
          const int pix = p.out_transpose_hw ? (w * H + h) : (h * W + w);
          out_off = ((long long)img * p.out_rows_per_image + p.out_row_offset + pix) * p.out_ld;
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
      const bool vec32 = p.vec32 != 0;          // 32-byte (full-sector) global accesses: cout % 16 == 0, aligned bases
      const bool res_vec = p.res_mode != 0 && interior && vec32;
      uint32_t res_next[CHUNK / 2];
      constexpr int STEP = (BN / CHUNK >= 2) ? 2 * CHUNK : CHUNK;   // two warps interleave chunks when there are >= 2
      const int step_rt = alt ? CHUNK : STEP;                       // (alternate-tile mode: this warp takes every chunk)
      const int c_first = (!alt && BN / CHUNK >= 2) ? half * CHUNK : 0;
      const bool idle_half = !alt && (BN / CHUNK < 2) && half == 1; // a single chunk: the second warp only arrives
      // (in a multi-convolution launch the residual may still be in the making: its dependency is awaited by the
      // producer warp, which the accumulator wait below orders before us -- so no early fetch there)
      const bool res_first = res_vec && !idle_half && n0 + c_first + CHUNK <= p.cout;
      if (res_first && deps == nullptr) {
#pragma unroll
        for (int j = 0; j < CHUNK / 16; ++j) hn_ldg256(p.res + res_off + n0 + c_first + 16 * j, &res_next[8 * j]);
      }

      if (dbg_flags & 16) {                      // experiment: one polling lane per warp
        if (lane == 0) hn_mbar_wait(&tmem_full[buf], acc_phase);
        __syncwarp();
      } else {
        hn_mbar_wait(&tmem_full[buf], acc_phase);
      }
      hn_tc_fence_after();
      if (res_first && deps != nullptr) {
#pragma unroll
        for (int j = 0; j < CHUNK / 16; ++j) hn_ldg256(p.res + res_off + n0 + c_first + 16 * j, &res_next[8 * j]);
      }
      if (warp == 2) hn_trace(trace, 2, tri, 2);
      const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + buf * BN;

      // Split-K: every work item stores its partial accumulator into its own slice of an fp32 scratch; the item that
      // arrives last (per-tile counter) adds the slices in split order -- a fixed summation order, so results do not
      // depend on which CTA finishes first -- and runs the epilogue.
      const bool split = p.splits > 1;
      bool finalize = true;
      float* sk_row = split ? p.sk_ws + (size_t)m * p.sk_ld + n0 : nullptr;
      const int ks_item = split ? w_ - st * splits : 0;
      if (split) {
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
          const bool is_last = old == (unsigned)p.splits - 1u;
          if (is_last) p.sk_cnt[st] = 0u;                      // everyone has arrived: reset for the next launch
          *sk_flag = is_last ? 1u : 0u;
        }
        hn_epi_bar_sync();
        finalize = *sk_flag != 0u;
        if (finalize) __threadfence();
      }

      // The chunk loop exists twice: FAST for the common case (bf16 output, every chunk complete, 32-byte accesses, no
      // split-K) carries no per-element predicates; the general copy keeps the ragged / fp32-row / split-K paths.
      // (Inside one body the compiler if-converts the slow paths into ~400 predicated-off instructions per chunk.)
      auto run_chunks = [&](auto fast_tag) {
        constexpr bool FAST = decltype(fast_tag)::value;
      auto chunk_body = [&](const int c0, uint32_t (&acc)[CHUNK], const bool preloaded) {
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
          if (!preloaded) {
            if constexpr (CHUNK == 32) {
              hn_tmem_ld32(t_row + c0, acc);
            } else {
              hn_tmem_ld16(t_row + c0, acc);
            }
            hn_tmem_ld_wait();
          }
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
        if (!FAST && p.out_kind == 1) {
          if (interior) {
            float* op = reinterpret_cast<float*>(p.out) + out_off + cbase;
#pragma unroll
            for (int j = 0; j < CHUNK; ++j)
              if (cbase + j < p.cout) op[j] = v[j];
          }
          return;
        }
        if (warp == 2) hn_trace(trace, 2, tri, 6);
        // bf16 outputs
        uint32_t packed[CHUNK / 2];
#pragma unroll
        for (int j = 0; j < CHUNK; j += 2) packed[j / 2] = hn_pack_bf16(v[j], v[j + 1]);
        if (warp == 2) hn_trace(trace, 2, tri, 7);
        if (interior && !(p.dbg_flags & 1)) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + out_off + cbase;
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
          // GroupNorm partial sums over the bf16-rounded values.  Per lane: (sum, sumsq) of the 4 channel octets of
          // this chunk = 8 values; a transposing butterfly (4+2+1+1+1 shuffles) leaves total k in lane 4*k, which
          // adds it to the CTA's shared-memory accumulator of its (image, group); flushed once at kernel end.
          if constexpr (CHUNK == 32) {
            float v8[8];
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
              v8[2 * o8] = interior ? s : 0.f;
              v8[2 * o8 + 1] = interior ? q : 0.f;
            }
            if (warp_uniform_img) {
              {
                const bool hi = lane & 16;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float send = hi ? v8[i] : v8[i + 4];
                  const float keep = hi ? v8[i + 4] : v8[i];
                  v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
              }
              {
                const bool hi = lane & 8;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const float send = hi ? v8[i] : v8[i + 2];
                  const float keep = hi ? v8[i + 2] : v8[i];
                  v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
              }
              {
                const bool hi = lane & 4;
                const float send = hi ? v8[0] : v8[1];
                const float keep = hi ? v8[1] : v8[0];
                v8[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
              v8[0] += __shfl_xor_sync(0xffffffffu, v8[0], 2);
              v8[0] += __shfl_xor_sync(0xffffffffu, v8[0], 1);
              if ((lane & 3) == 0 && interior_mask != 0u) {
                const int k = lane >> 2;                         // value index: octet k/2, stat k&1
                const int group = (cbase + (k >> 1) * 8) / p.gn_group_size;
                if (group < p.gn_groups) {
                  if (gn_smem) atomicAdd(&gn_acc[(warp_img * p.gn_groups + group) * 2 + (k & 1)], v8[0]);
                  else atomicAdd(p.gn_stats + ((long long)warp_img * p.gn_groups + group) * 2 + (k & 1), (double)v8[0]);
                }
              }
            } else if (interior) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int group = (cbase + (k >> 1) * 8) / p.gn_group_size;
                if (group < p.gn_groups) {
                  if (gn_smem) atomicAdd(&gn_acc[(img * p.gn_groups + group) * 2 + (k & 1)], v8[k]);
                  else atomicAdd(p.gn_stats + ((long long)img * p.gn_groups + group) * 2 + (k & 1), (double)v8[k]);
                }
              }
            }
          }
        }
            };
      const int c_end = (idle_half || (p.dbg_flags & 2) || !finalize) ? 0 : BN;
      // TMEM loads one chunk ahead (tcgen05.ld of chunk i+1 in flight while chunk i is scaled, packed and stored; two
      // register sets, loop fully unrolled).  Measured A/B on one box: isolated 256-wide layers gain 5 % (layer3 34.2 ->
      // 32.4 us) but the whole step loses 2 % (2108 -> 2066 frames/s; 48 bytes of spills at the 168-register cap), so
      // it is compiled out.
      constexpr bool TMEM_PREFETCH = false;
      if constexpr (TMEM_PREFETCH && FAST && CHUNK == 32 && (BN / STEP) >= 2) {
        if (c_first < c_end) {
          uint32_t acc_a[CHUNK], acc_b[CHUNK];
          hn_tmem_ld_chunk<CHUNK>(t_row + c_first, acc_a);
#pragma unroll
          for (int i = 0; i < BN / STEP; ++i) {
            const int c0 = c_first + i * STEP;
            hn_tmem_ld_wait();
            if (i & 1) {
              if (i + 1 < BN / STEP) hn_tmem_ld_chunk<CHUNK>(t_row + c0 + STEP, acc_a);
              chunk_body(c0, acc_b, true);
            } else {
              if (i + 1 < BN / STEP) hn_tmem_ld_chunk<CHUNK>(t_row + c0 + STEP, acc_b);
              chunk_body(c0, acc_a, true);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c0 = c_first; c0 < c_end; c0 += step_rt) {
          uint32_t acc[CHUNK];
          chunk_body(c0, acc, false);
        }
      }
      };
      if (epi_fast) run_chunks(FastTag<true>{});
      else run_chunks(FastTag<false>{});
      // accumulator buffer drained -> hand it back to the MMA warp (split-K items did so after their reduction)
      if (!split) {
        hn_tc_fence_before();
        __syncwarp();
        if (lane < (alt ? 2 : 1)) hn_mbar_arrive(&tmem_empty[buf]);
      }
      if (deps != nullptr && finalize) {                  // this (m, n) output tile is in global memory: tell the consumers
        __threadfence();
        hn_epi_bar_sync();
        if (threadIdx.x == 64) atomicAdd(deps->done + mt, 1u);
      }
      if (warp == 2) hn_trace(trace, 2, tri, 3);
    }
    if (gn_smem) {
      hn_epi_bar_sync();
      for (int i = threadIdx.x - 64; i < gn_vals; i += EPI_THREADS) {
        const float v = gn_acc[i];
        if (v != 0.f) atomicAdd(p.gn_stats + i, (double)v);
      }
    }
  }

}

// Barrier init, TMEM allocation.  Returns the TMEM base address.
template <int BN, int CS>
__device__ __forceinline__ uint32_t conv_prologue(uint8_t* smem_hdr, const CUtensorMap* pf_a, const CUtensorMap* pf_b) {
  using C = Cfg<BN>;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_hdr);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + MAX_STAGES;
  uint64_t* b_full_ring = bars + 2 * MAX_STAGES;
  uint64_t* b_empty = bars + 3 * MAX_STAGES;
  uint64_t* tmem_full = bars + 4 * MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 8;
  uint64_t* b_full = tmem_empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
  if (threadIdx.x == 0) {
    if (pf_a) hn_tma_prefetch_desc(pf_a);
    if (pf_b) hn_tma_prefetch_desc(pf_b);
    for (int s = 0; s < MAX_STAGES; ++s) {
      hn_mbar_init(&a_full[s], 1);
      hn_mbar_init(&a_empty[s], 1);      // A boxes are private to the CTA
      hn_mbar_init(&b_full_ring[s], 1);
      hn_mbar_init(&b_empty[s], CS);     // every CTA of the cluster releases the slot (its peers write into it)
    }
    hn_mbar_init(b_full, 1);
    for (int b = 0; b < 8; ++b) {
      hn_mbar_init(&tmem_full[b], 1);
      hn_mbar_init(&tmem_empty[b], EPI_WARPS);   // one arrive per epilogue warp
    }
    hn_mbar_init_fence();
  }
  if ((threadIdx.x >> 5) == 1) hn_tmem_alloc<C::TMEM_COLS>(tmem_slot);
  hn_tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) hn_cluster_sync();   // peers' barriers must exist before anyone signals them
  hn_tc_fence_after();
  return *tmem_slot;
}

template <int BN, int CS>
__device__ __forceinline__ void conv_teardown(uint32_t tmem_base) {
  using C = Cfg<BN>;
  hn_tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) hn_cluster_sync();   // no CTA may retire while a peer can still signal its barriers
  if ((threadIdx.x >> 5) == 1) {
    hn_tc_fence_after();
    hn_tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// CS = thread-block cluster size along M: the CS CTAs of a cluster work on CS consecutive M tiles of the same N
// tile in lock step; each loads 1/CS of the B (weight) tile and multicasts it to all of them.  RB = the layer's
// whole weight slice stays resident in shared memory.
template <int BN, int CS, bool RB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                  const __grid_constant__ ConvParams p) {
  static_assert(CS == 1 || (BN / CS) % 8 == 0, "B slices must keep whole 8-row swizzle atoms");
  static_assert(!(RB && CS > 1), "resident weights are per CTA");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_hdr = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t tmem_base = conv_prologue<BN, CS>(smem_hdr, &tm_a, &tm_b);
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may run
  // while the previous kernel of the stream is still draining; global memory is touched only after the wait.
  // The early trigger lets the next kernel do the same under this one.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  PipeState ps = {0, 0, 0, 0u, 0u};
  conv_roles<BN, CS, RB>(&tm_a, &tm_b, p, smem_hdr, tmem_base, blockIdx.x / CS, gridDim.x / CS, ps, true);
  conv_teardown<BN, CS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// Several dependent convolutions in ONE launch (the 67 tiny convolutions of the A2J pose net cost ~10 us of fixed
// latency each as separate launches).  `phases` lists the convolutions in a dependency-consistent order; tiles
// synchronise by dataflow (ConvDeps): hn_conv_multi_build finds, from the buffer pointers, which earlier convolution
// writes each input / residual and which of its M tiles every tile reads.  Launched cooperatively so that all CTAs are
// resident (a waiting tile's producers must be able to run).
// ---------------------------------------------------------------------------------------------------------------
struct PhaseDesc {
  CUtensorMap ta;
  CUtensorMap tb;
  ConvParams p;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_multi_kernel(const PhaseDesc* __restrict__ phases, int n_convs, long long* __restrict__ conv_clock) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_hdr = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t tmem_base = conv_prologue<BN, 1>(smem_hdr, nullptr, nullptr);
  PipeState ps = {0, 0, 0, 0u, 0u};
  // Every role walks the convolutions in order, straight from the plan in global memory: no block-wide or grid-wide
  // barrier between convolutions.  The producer warp of a tile waits for the producer tiles it reads (ConvDeps), the
  // epilogue announces finished tiles; the roles of one CTA may be in different convolutions at the same time.
  // Each role keeps its own copy of the current convolution's parameters in shared memory (the GroupNorm accumulator
  // area, unused here): read straight from the plan in global memory, the per-step parameter reads of the role loops
  // miss the small L1 and cost an L2 round trip each.
  static_assert(3 * sizeof(ConvParams) <= GN_SMEM_FLOATS * 4, "role parameter slots do not fit");
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int role = warp < 2 ? warp : 2;
  ConvParams* slot = reinterpret_cast<ConvParams*>(smem_hdr + HDR_BARS) + role;
  int tile_off = 0;                                      // tiles are dealt round-robin over the CTAs, continuing across convs
  for (int j = 0; j < n_convs; ++j) {
    {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&phases[j].p);
      uint32_t* dst = reinterpret_cast<uint32_t*>(slot);
      constexpr int WORDS = (int)(sizeof(ConvParams) / 4);
      if (role == 2) {
        hn_epi_bar_sync();                               // every epilogue warp is done with the previous convolution
        for (int i = threadIdx.x - 64; i < WORDS; i += EPI_THREADS) dst[i] = __ldg(src + i);
        hn_epi_bar_sync();
      } else {
        for (int i = threadIdx.x & 31; i < WORDS; i += 32) dst[i] = __ldg(src + i);
        __syncwarp();
      }
    }
    if (role == 0 && (threadIdx.x & 31) == 0 && j + 1 < n_convs) {   // descriptors of the next convolution -> descriptor cache
      hn_tma_prefetch_desc(&phases[j + 1].ta);
      hn_tma_prefetch_desc(&phases[j + 1].tb);
    }
    const ConvParams& cp = *slot;
    const int tiles = cp.m_tiles * cp.n_tiles * cp.splits;
    int first = ((int)blockIdx.x - tile_off) % (int)gridDim.x;
    if (first < 0) first += gridDim.x;
    conv_roles<BN, 1, false>(&phases[j].ta, &phases[j].tb, cp, smem_hdr, tmem_base, first, gridDim.x, ps);
    tile_off = (tile_off + tiles) % (int)gridDim.x;
    if (conv_clock != nullptr && blockIdx.x == 0 && threadIdx.x == 64) conv_clock[j] = clock64();   // bring-up only
  }
  conv_teardown<BN, 1>(tmem_base);
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
             const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    hn_set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return HN_ERR_CUDA;
  }
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides_bytes, box, ones,
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

template <int BN, int CS, bool RB>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const ConvParams& p, cudaStream_t st) {
  using C = Cfg<BN>;
  constexpr int SMEM = SMEM_BYTES_ALL;
  static bool attr_set = false;
  if (!attr_set) {
    HN_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, CS, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SMEM));
    attr_set = true;
  }
  const int super = hn_div_up(p.m_tiles, CS) * p.n_tiles * p.splits;
  // equal work per CTA: with w = ceil(tiles / SMs) waves, ceil(tiles / w) CTAs finish at the same time as a full grid
  // would and leave the other SMs to kernels of concurrent streams (graph branches)
  const int max_clusters = hn_num_sms() / CS;
  const int waves = hn_div_up(super, max_clusters);
  const int clusters = hn_div_up(super, waves);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * CS);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CS > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CS;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  HN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BN, CS, RB>, ta, tb, p));
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

bool patch_enabled() {          // HN_CONV_PATCH=1|0: patch tiles for the resident-weights 3x3 layers (see PATCH_H)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HN_CONV_PATCH");
    v = (e && e[0] != '0') ? 1 : 0;
  }
  return v == 1;
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
  int bn, cs;
  bool rb;
};

// Validate a descriptor and derive kernel parameters + tensor maps.  force_bn > 0 pins the tile width (and disables
// clusters / resident weights), as the multi-convolution kernel needs.
int build_conv(const hn_conv_desc* d, int force_bn, BuiltConv* out) {
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
  if (stem) {
    HN_REQUIRE(!force_bn && d->kh == 1 && d->stride == 1 && d->cin == 256 && d->halo_in == 0 && d->in_phases == 1 &&
                   d->cout_pad <= 128 && !d->res && !d->gn_stats && !d->splitk_ws,
               "hn_conv2d_bf16: a direct stem is a plain 1x1-over-patches convolution with cin = 256, cout_pad <= 128");
    HN_REQUIRE(d->stem_pitch_h >= 2 * d->h + 6 && d->stem_pitch_h % 2 == 0 && d->stem_pitch_w >= 2 * d->w + 8,
               "hn_conv2d_bf16: stem frame %dx%d too small for a %dx%d output (needs >= %dx%d, even height)",
               d->stem_pitch_h, d->stem_pitch_w, d->h, d->w, 2 * d->h + 6, 2 * d->w + 8);
    p.stem_tpr = hn_div_up(d->w, BLOCK_M);
    p.stem_h = d->h;
    p.m_tiles = d->n * d->h * p.stem_tpr;
  }
  int bn = force_bn ? force_bn
                    : (d->block_n ? d->block_n
                                  : pick_block_n(d->cout_pad, p.m_tiles, p.num_taps * p.cin_chunks, d->gn_stats ? 32 : 16));
  HN_REQUIRE((bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256) && d->cout_pad % bn == 0,
             "hn_conv2d_bf16: block_n=%d does not divide cout_pad=%d", bn, d->cout_pad);
  p.n_tiles = d->cout_pad / bn;
  // cluster of 2 with multicast weights when there is at least one pair of M tiles per SM pair
  // measured: pairs with multicast weights gain ~3 % on 256-wide tiles and lose elsewhere -> opt-in only
  int cs = (d->cluster && !force_bn) ? d->cluster : 1;
  HN_REQUIRE(cs == 1 || (cs == 2 && bn == 256), "hn_conv2d_bf16: cluster=%d unsupported with block_n=%d", cs, bn);
  // resident weights: the layer's whole weight slice stays in shared memory and only A boxes stream, when it is one
  // narrow N tile whose weights fit next to >= 4 A slots and there are enough tiles per CTA to amortise the load
  const int k_blocks_total = p.num_taps * p.cin_chunks;
  const long long b_bytes = (long long)k_blocks_total * bn * BLOCK_K * 2;
  bool rb = !force_bn && cs == 1 && p.n_tiles == 1 && bn <= 64 && b_bytes <= PIPE_BYTES_MAX - 4 * A_SLOT_BYTES &&
            p.m_tiles >= 2 * hn_num_sms() && !(d->debug & 16) && !stem;
  const bool uni = bn <= 128 && !rb;      // unified stages (must match the kernel's constexpr UNI)
  const bool rb3 = rb && d->kh == 3 && d->stride == 1 && !(d->debug & 32) &&
                   PIPE_BYTES_MAX - b_bytes >= 2 * 3 * A_SLOT_BYTES && p.rows > 2 * d->dilation * p.wp;
  // patch tiles (see PATCH_H): large maps only -- a 16 x 8 patch grid wastes too much on small ones.  Measured neutral
  // (layer1 60.4 vs 58.4 us, whole step 3.88-3.91 vs 3.89-3.90 ms: these layers are bound by the MMA-issuing warp, not by
  // L2 -> SM traffic), so it is opt-in: debug bit 14 or HN_CONV_PATCH=1.
  const bool patch = rb3 && d->dilation == 1 && d->halo_in >= 1 && d->h >= 64 && d->w >= 64 &&
                     ((d->debug & 16384) || patch_enabled());

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
  p.a_box_bytes = box_rows * BLOCK_K * 2;
  p.k_steps = (uni && p.uni_chunk_step > 1) ? hn_div_up(p.cin_chunks, p.uni_chunk_step) : p.n_groups * p.cin_chunks;
  if (uni) {
    p.uni_plane_bytes = box_rows * BLOCK_K * 2;
    p.uni_a_bytes = a_planes * p.uni_plane_bytes;
    p.uni_b_bytes = b_tiles * bn * BLOCK_K * 2;
    // convolutions chained in one multi-convolution launch share the ring: same stage stride and depth for all of
    // them (the largest stage: two 136-row planes + three weight tiles)
    p.uni_stride = force_bn ? 2 * A_SLOT_BYTES + 3 * bn * BLOCK_K * 2 : p.uni_a_bytes + p.uni_b_bytes;
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
    if (patch) {
      p.patch_tx = hn_div_up(d->w, PATCH_W);
      p.patch_ty = hn_div_up(d->h, PATCH_H);
      p.m_tiles = d->n * p.patch_tx * p.patch_ty;
      p.a_box_bytes = PATCH_BOX_BYTES;
    }
    const int slots = (PIPE_BYTES_MAX - p.rb_b_bytes) / (patch ? PATCH_SLOT_BYTES : (rb3 ? 3 * A_SLOT_BYTES : A_SLOT_BYTES));
    p.na_stages = slots > MAX_STAGES ? MAX_STAGES : slots;
    p.nb_stages = 1;
  } else {
    p.na_stages = Cfg<256>::NA;
    p.nb_stages = Cfg<256>::NB;
  }
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
  if (d->out_kind == 1) HN_REQUIRE(d->out_ld >= d->cout, "hn_conv2d_bf16: out_ld < cout");
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
  p.gn_stats = d->gn_stats;
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
  if (d->splitk_ws && d->splitk_counters && cs == 1 && bn >= 32) {
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
  p.epi_alt = (!force_bn && p.n_tiles == 1 && p.splits == 1 && !(d->debug & 32768) &&
               ((d->debug & 65536) || bn <= epi_alt_max_bn())) ? 1 : 0;
  CUtensorMap& ta = out->ta;
  CUtensorMap& tb = out->tb;
  if (stem) {
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
      if (!force_bn) {
        p.uni_stride = p.uni_a_bytes + p.uni_b_bytes;
        const int stages = PIPE_BYTES_MAX / p.uni_stride;
        p.uni_stages = p.na_stages = stages > MAX_STAGES ? MAX_STAGES : stages;
      }
      if (p.splits > p.k_steps) p.splits = p.k_steps;
    }
  }
  if (patch) {
    // the haloed NHWC tensor itself: (channels, padded columns, padded rows, images); out-of-range rows / columns of the
    // last patches are zero-filled by the TMA unit
    const cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)p.wp, (cuuint64_t)p.hp, (cuuint64_t)d->n};
    const cuuint64_t strides[3] = {(cuuint64_t)d->cin * 2, (cuuint64_t)p.wp * d->cin * 2, (cuuint64_t)p.hp * p.wp * d->cin * 2};
    const cuuint32_t box[4] = {BLOCK_K, PATCH_W + 2, PATCH_H + 2, 1};
    int rc = make_map(&ta, d->in, 4, dims, strides, box);
    if (rc) return rc;
  } else if (!stem && !(uni && p.uni_a_rank4)) {
    // dims (channels, rows, phases); unified stride-2 3x3: the box spans both column phases of a row phase.
    // rb3: the third dimension steps by one (dilated) image row instead -- an overlapping view of the same matrix;
    // its row count is cut by two image rows so that row + 2 * wp stays inside the allocation (the rows cut off are the
    // last image's bottom halo, which only halo outputs of kernel row 0 would read)
    const cuuint64_t dims[3] = {(cuuint64_t)d->cin, (cuuint64_t)(rb3 ? p.rows - 2 * d->dilation * p.wp : p.rows),
                                (cuuint64_t)(rb3 ? 3 : d->in_phases)};
    const cuuint64_t strides[2] = {(cuuint64_t)d->cin * 2,
                                   rb3 ? (cuuint64_t)d->dilation * p.wp * d->cin * 2 : (cuuint64_t)p.rows * d->cin * 2};
    const cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)box_rows, (cuuint32_t)(rb3 ? 3 : (uni ? a_planes : 1))};
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
    const cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)(bn / cs)};
    int rc = make_map(&tb, d->weight, 2, dims, strides, box);
    if (rc) return rc;
  }
  out->p = p;
  out->bn = bn;
  out->cs = cs;
  out->rb = rb;
  return HN_OK;
}

constexpr int MULTI_BN = 64;

// device layout of a plan: [counters: n_convs x MULTI_MAX_MT uint32][PhaseDesc array][ConvDeps array][dep tables]
constexpr int MULTI_MAX_MT = 1024;      // M tiles per convolution (the pose net's largest layer has 133)
size_t multi_phases_off(int n_convs) { return (((size_t)n_convs * MULTI_MAX_MT * 4 + 255) / 256) * 256; }
size_t multi_deps_off(int n_convs) { return multi_phases_off(n_convs) + (((size_t)n_convs * sizeof(PhaseDesc) + 255) / 256) * 256; }
size_t multi_tabs_off(int n_convs) { return multi_deps_off(n_convs) + (((size_t)n_convs * sizeof(ConvDeps) + 255) / 256) * 256; }
size_t multi_total(int n_convs) { return multi_tabs_off(n_convs) + (size_t)n_convs * MAX_DEPS * 2 * MULTI_MAX_MT * sizeof(short) + 256; }

// Which producer M tiles does M tile `mt` of a consumer read?  `rows_lo..rows_hi` = the consumer's rows (in the geometry
// of the buffer `geo` describes) that the tile touches; the answer is conservative (whole image rows).
struct BufGeo {        // a haloed NHWC buffer as a consumer sees it, or the phase-split copy (phases = 4)
  int n, hp, wp, halo, phases;
};
void producer_tile_range(const BufGeo& g, long long rows_lo, long long rows_hi, const ConvParams& prod, int* t_first,
                         int* t_last) {
  const long long img_rows = (long long)g.hp * g.wp, total = (long long)g.n * img_rows;
  *t_first = 1;
  *t_last = 0;                                         // empty by default
  if (rows_lo < 0) rows_lo = 0;
  if (rows_hi >= total) rows_hi = total - 1;
  if (rows_lo > rows_hi) return;
  const int H = g.hp - 2 * g.halo;                     // interior size of the consumer's view
  const int img_lo = (int)(rows_lo / img_rows), img_hi = (int)(rows_hi / img_rows);
  int h_lo = (int)((rows_lo - img_lo * img_rows) / g.wp) - g.halo, h_hi = (int)((rows_hi - img_hi * img_rows) / g.wp) - g.halo;
  if (h_lo < 0) h_lo = 0;
  if (h_hi > H - 1) h_hi = H - 1;
  // pixel rows of the producer's OUTPUT grid: the same grid, or twice as fine when the consumer reads the phase split
  const int ph = prod.hp - 2 * prod.halo, pw = prod.wp - 2 * prod.halo;      // producer output size (= its compute grid)
  int y_lo = g.phases == 4 ? 2 * h_lo : h_lo, y_hi = g.phases == 4 ? 2 * h_hi + 1 : h_hi;
  if (y_hi > ph - 1) y_hi = ph - 1;
  if (img_lo == img_hi && y_lo > y_hi) return;         // only halo rows
  if (y_lo > ph - 1) y_lo = ph - 1;
  const long long p_img = (long long)prod.hp * prod.wp;
  const long long m_lo = img_lo * p_img + (long long)(y_lo + prod.halo) * prod.wp + prod.halo;
  const long long m_hi = img_hi * p_img + (long long)(y_hi + prod.halo) * prod.wp + prod.halo + pw - 1;
  *t_first = (int)(m_lo / BLOCK_M);
  *t_last = (int)(m_hi / BLOCK_M);
}

}  // namespace

extern "C" int hn_conv2d_bf16(const hn_conv_desc* d, void* stream) {
  BuiltConv b;
  int rc = build_conv(d, 0, &b);
  if (rc) return rc;
  const ConvParams& p = b.p;
  const CUtensorMap& ta = b.ta;
  const CUtensorMap& tb = b.tb;
  const int bn = b.bn, cs = b.cs;
  const bool rb = b.rb;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (rb) {
    switch (bn) {
      case 64: return launch<64, 1, true>(ta, tb, p, st);
      case 32: return launch<32, 1, true>(ta, tb, p, st);
      default: return launch<16, 1, true>(ta, tb, p, st);
    }
  }
  if (cs == 2) {
    switch (bn) {
      case 256: return launch<256, 2, false>(ta, tb, p, st);
      case 128: return launch<128, 2, false>(ta, tb, p, st);
      default: return launch<64, 2, false>(ta, tb, p, st);
    }
  }
  switch (bn) {
    case 256: return launch<256, 1, false>(ta, tb, p, st);
    case 128: return launch<128, 1, false>(ta, tb, p, st);
    case 64: return launch<64, 1, false>(ta, tb, p, st);
    case 32: return launch<32, 1, false>(ta, tb, p, st);
    default: return launch<16, 1, false>(ta, tb, p, st);
  }
}

static long long* g_multi_trace = nullptr;
// Bring-up only: when set, CTA 0 of every following hn_conv_multi_run writes clock64() after each group into buf[g].
extern "C" int hn_conv_multi_set_trace(void* buf) {
  g_multi_trace = reinterpret_cast<long long*>(buf);
  return HN_OK;
}

extern "C" int64_t hn_conv_multi_plan_bytes(int n_convs, int n_groups) {
  (void)n_groups;
  return (int64_t)multi_total(n_convs);
}

extern "C" int hn_conv_multi_build(const hn_conv_desc* descs, int n_convs, const int* group_begin_host, int n_groups,
                                   void* plan_dev, int64_t plan_bytes) {
  HN_REQUIRE(descs && group_begin_host && plan_dev && n_convs > 0 && n_groups > 0, "hn_conv_multi_build: bad arguments");
  HN_REQUIRE(plan_bytes >= hn_conv_multi_plan_bytes(n_convs, n_groups), "hn_conv_multi_build: plan buffer too small");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(plan_dev) & 255) == 0, "hn_conv_multi_build: plan buffer must be 256-byte aligned");
  HN_REQUIRE(group_begin_host[0] == 0 && group_begin_host[n_groups] == n_convs, "hn_conv_multi_build: group table");
  std::vector<uint8_t> host(multi_total(n_convs), 0);
  uint8_t* dev = reinterpret_cast<uint8_t*>(plan_dev);
  PhaseDesc* ph = reinterpret_cast<PhaseDesc*>(host.data() + multi_phases_off(n_convs));
  ConvDeps* deps = reinterpret_cast<ConvDeps*>(host.data() + multi_deps_off(n_convs));
  short* tabs = reinterpret_cast<short*>(host.data() + multi_tabs_off(n_convs));
  auto dev_cnt = [&](int j) { return reinterpret_cast<unsigned*>(dev) + (size_t)j * MULTI_MAX_MT; };
  auto dev_tab = [&](int j, int d, int which) {
    return reinterpret_cast<short*>(dev + multi_tabs_off(n_convs)) + ((size_t)(j * MAX_DEPS + d) * 2 + which) * MULTI_MAX_MT;
  };
  auto host_tab = [&](int j, int d, int which) { return tabs + ((size_t)(j * MAX_DEPS + d) * 2 + which) * MULTI_MAX_MT; };
  int max_group_tiles = 1;
  for (int g = 0; g < n_groups; ++g) {
    HN_REQUIRE(group_begin_host[g + 1] > group_begin_host[g], "hn_conv_multi_build: empty group %d", g);
    int tiles = 0;
    for (int j = group_begin_host[g]; j < group_begin_host[g + 1]; ++j) {
      HN_REQUIRE(descs[j].cout_pad % MULTI_BN == 0, "hn_conv_multi_build: conv %d: cout_pad %% 64 != 0", j);
      HN_REQUIRE(descs[j].gn_stats == nullptr, "hn_conv_multi_build: GroupNorm statistics are not supported here");
      BuiltConv b;
      int rc = build_conv(&descs[j], MULTI_BN, &b);
      if (rc) return rc;
      HN_REQUIRE(b.p.m_tiles <= MULTI_MAX_MT, "hn_conv_multi_build: conv %d has %d M tiles (max %d)", j, b.p.m_tiles, MULTI_MAX_MT);
      ph[j].ta = b.ta;
      ph[j].tb = b.tb;
      ph[j].p = b.p;
      ph[j].p.deps = reinterpret_cast<const ConvDeps*>(dev + multi_deps_off(n_convs)) + j;
      tiles += b.p.m_tiles * b.p.n_tiles * b.p.splits;
    }
    if (tiles > max_group_tiles) max_group_tiles = tiles;
  }
  // dependencies from the buffer pointers: the latest earlier convolution that writes this one's input / residual
  for (int j = 0; j < n_convs; ++j) {
    const hn_conv_desc& dj = descs[j];
    const ConvParams& pj = ph[j].p;
    ConvDeps& dd = deps[j];
    dd.done = dev_cnt(j);
    dd.n_deps = 0;
    const int reach = dj.kh == 3 ? dj.dilation * pj.wp + dj.dilation : 0;     // rows a tile reads beyond its own
    struct In { const void* ptr; BufGeo geo; int reach; };
    In ins[2] = {{dj.in, {dj.n, pj.hp, pj.wp, pj.halo, dj.in_phases}, reach},
                 {dj.res, {dj.n, pj.res_hp, pj.res_wp, pj.res_halo, 1}, 0}};
    for (int k = 0; k < 2; ++k) {
      if (!ins[k].ptr) continue;
      int prod = -1, via_phase = 0;
      for (int i = j - 1; i >= 0 && prod < 0; --i) {
        if (descs[i].out == ins[k].ptr && descs[i].out_kind == 0) prod = i;
        else if (descs[i].out_phase && descs[i].out_phase == ins[k].ptr) { prod = i; via_phase = 1; }
      }
      if (prod < 0) continue;                          // written before this launch
      HN_REQUIRE((ins[k].geo.phases == 4) == (via_phase == 1), "hn_conv_multi_build: conv %d reads conv %d's %s buffer as %s",
                 j, prod, via_phase ? "phase-split" : "plain", ins[k].geo.phases == 4 ? "phase-split" : "plain");
      HN_REQUIRE(dd.n_deps < MAX_DEPS, "hn_conv_multi_build: too many dependencies");
      const int d = dd.n_deps++;
      dd.dep_done[d] = dev_cnt(prod);
      dd.dep_need[d] = ph[prod].p.n_tiles;
      dd.dep_first[d] = dev_tab(j, d, 0);
      dd.dep_last[d] = dev_tab(j, d, 1);
      short* tf = host_tab(j, d, 0);
      short* tl = host_tab(j, d, 1);
      for (int mt = 0; mt < pj.m_tiles; ++mt) {
        // consumer rows of this tile; the residual is indexed by OUTPUT pixel = compute row of the consumer, which maps
        // to the residual buffer's own geometry through the pixel (for k = 1 use the compute geometry to find pixels)
        long long lo = (long long)mt * BLOCK_M - ins[k].reach, hi = (long long)mt * BLOCK_M + BLOCK_M - 1 + ins[k].reach;
        BufGeo geo = ins[k].geo;
        if (k == 1) geo = {dj.n, pj.hp, pj.wp, pj.halo, 1};   // rows of the consumer's compute grid -> pixels
        int a = 1, b2 = 0;
        producer_tile_range(geo, lo, hi, ph[prod].p, &a, &b2);
        tf[mt] = (short)a;
        tl[mt] = (short)b2;
      }
    }
  }
  const int grid = max_group_tiles < hn_num_sms() ? max_group_tiles : hn_num_sms();
  HN_CHECK_CUDA(cudaMemcpy(plan_dev, host.data(), host.size(), cudaMemcpyHostToDevice));
  return grid;     // > 0: number of CTAs hn_conv_multi_run will launch
}

extern "C" int hn_conv_multi_run(void* plan_dev, int n_convs, int n_groups, int grid, void* stream) {
  HN_REQUIRE(plan_dev && n_convs > 0 && n_groups > 0 && grid > 0 && grid <= hn_num_sms(), "hn_conv_multi_run: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static bool attr_set = false;
  if (!attr_set) {
    HN_CHECK_CUDA(cudaFuncSetAttribute(conv_multi_kernel<MULTI_BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SMEM_BYTES_ALL));
    attr_set = true;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>(plan_dev);
  HN_CHECK_CUDA(cudaMemsetAsync(base, 0, (size_t)n_convs * MULTI_MAX_MT * 4, st));   // the tile completion counters
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES_ALL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;              // all CTAs resident: a waiting tile's producers can always run
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const PhaseDesc* phases = reinterpret_cast<const PhaseDesc*>(base + multi_phases_off(n_convs));
  HN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_multi_kernel<MULTI_BN>, phases, n_convs, g_multi_trace));
  hn_count_launch();
  return HN_OK;
}
