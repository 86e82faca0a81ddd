// conv_band2.cu - two chained 3x3 convolutions of a residual block in ONE launch (sm_100a, tcgen05 / TMEM): the
// intermediate tensor never leaves the SM.
//
// Why.  The C <= 16 levels of the ARU-Net are HBM bound layer by layer: a residual block
// conv1 -> ReLU -> convR_0 -> convR_1 -> convR_2 + conv1_preact -> ReLU (ARU_v1.py:212-227, 266-281) moves 152-176 B per
// pixel at level 0 for 64 B of compulsory traffic.  This kernel runs the block as two launches of two stages each,
//     A = conv1 (+ pre-activation export) -> convR_0         B = convR_1 -> convR_2 + residual (+ 2x2 max-pool)
// and keeps the stage-0 result in a shared-memory row FIFO: 88-112 B per pixel.  (All four stages in one launch do not
// fit: the banded weight masters + three row FIFOs + the input ring need ~275 KB, and with N = 128 columns per MMA the
// 4 KB A-operand read per MMA makes the shared-memory port the limit anyway)
//
// Both stages are k_conv_band GEMMs (conv_band.cu):  D[m = (rho, co), n = x] += A_window[m, k] * B_row[k, n]  with
// M = R x C_out = 128 rows, N = N_mma <= 128 columns, banded weight masters as A.  Geometry of one strip s:
//     stage-1 accumulator column n  <->  image column x_left + n          (x_left = s * N_out, N_out = N_mma - 2 valid)
//     stage-0 accumulator column n  <->  image column x_left - 1 + n      (one halo column on each side)
//     FIFO unit j                   <->  image column x_left - 1 + j      (units 0 .. N_mma - 1 written, N_mma, N_mma + 1 zero)
//     stage-0 tile u  = rows u R + 1 .. u R + R  (tile -1 = warm-up: one extra stage-0 tile per strip segment),
//     stage-1 tile t  = rows t R .. t R + R - 1, reading FIFO rows t R - 1 .. t R + R = the last 2 rows of stage-0 tile
//     t - 1 and all of tile t.  Every stage-0 item appends exactly R rows to the FIFO (2R + 2 slots), out-of-image
//     positions as zeros (SAME padding of the second convolution sees zeros, not conv values).
// Work items are issued in a fixed interleaved order  s0, s0, s0, s1, s0, s1, ...  (stage 1 runs one item behind, so a
// stage-0 tile has two MMA groups of time to be drained and stored into the FIFO); every role derives the same
// sequence arithmetically (struct Sched).  FIFO slots are recycled without a barrier: the stage-0 item that overwrites
// them is issued after the stage-1 item that read them, and the tensor pipe executes in order.
//
// Roles (768 threads, one persistent CTA per SM): warp 0 lane 0 MMA issuer, warp 1 producer (bulk copies of input rows),
// warp 2 TMEM allocator, warps 4-7 drain of stage 1 (TMEM -> +bias -> 16-bit -> transposition slab), warps 8-11 drain of
// stage 0 (TMEM -> +bias -> activation / border mask -> 2-byte stores straight into the position-major FIFO rows: lane
// (rho, co) owns 2 bytes of every 16-byte position), warps 12-19 store of stage 1 (slab -> + residual, ReLU, fused max-pool,
// global stores), warps 20-23 only when stage 0 exports a global copy (conv1's pre-activation = the block's residual
// operand): the drain then writes raw values and these warps store them from the FIFO and apply the activation in place.
// The two drains run in parallel - one drain/store pipeline for both stages was the bottleneck of the first version
// (profiles/r02a_profile_ops_n32.txt).  Rounding: same operands and fp32 accumulation as k_conv_band; the vertical tap
// pairs of stage 0 group rows with the other parity, so results agree with the one-launch-per-layer path to fp32
// summation order, not bit for bit.
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <vector>

#include "aru_common.cuh"
#include "kernels.h"
#include "band_common.cuh"

namespace aru {

namespace {

constexpr int NUM_THREADS2 = 768;
constexpr int SLAB2 = 32;
constexpr int SLAB2_ROW = (SLAB2 + 1) * 16;
constexpr int SLAB2_BYTES = 16 * SLAB2_ROW;
constexpr int N_SLAB2 = 4;
// shared-memory header (bytes)
constexpr int O2_INFULL = 0;                  // uint64[4]: input rows of stage-0 item j on [j & 3]
constexpr int O2_INDONE = O2_INFULL + 32;     // uint64[4]
constexpr int O2_TFULL = O2_INDONE + 32;      // uint64[4]: [stage * 2 + buffer]
constexpr int O2_TEMPTY = O2_TFULL + 32;      // uint64[4]
constexpr int O2_SFULL = O2_TEMPTY + 32;      // uint64[4]
constexpr int O2_SEMPTY = O2_SFULL + 32;      // uint64[4]
constexpr int O2_MIDFULL = O2_SEMPTY + 32;    // uint64[4]: FIFO rows of stage-0 item j on [j & 3]
constexpr int O2_MIDRAW = O2_MIDFULL + 32;    // uint64[4]: raw FIFO rows of stage-0 item j (before the in-place activation)
constexpr int O2_TMEMPTR = O2_MIDRAW + 32;    // uint32
constexpr int O2_QS = O2_TMEMPTR + 16;        // int[8]
constexpr int O2_BIAS = O2_QS + 32;           // float[2][128]
constexpr int HDR2_BYTES = ((O2_BIAS + 1024 + 127) / 128) * 128;

__device__ unsigned long long g_band2_stats[160][16];

struct ConvBand2Args {
  const act_t* in;
  long long in_plane;
  act_t* out0;             // activated stage-0 output (stored only when something else reads it), else null
  long long out0_plane;
  act_t* pre0;             // pre-activation stage-0 output (conv1: the block's residual operand), else null
  long long pre0_plane;
  act_t* out;              // stage-1 output
  long long out_plane;
  const act_t* res;        // stage-1 residual operand, else null
  long long res_plane;
  const act_t* wpack0;
  const act_t* wpack1;
  const float* bias0;
  const float* bias1;
  int cin_chunks, cop, nc_shift, R, N, n_out, n_strips, n_ty, S, RS;
  int J, masters0_bytes, masters1_bytes;
  int in_slot_units, in_row_units, mid_slot_units, mid_row_units;
  int W, H, Wp, Hp;
  long long lead;
  long long tiles;
  int act0, act1;
  int nbuf0, nbuf1, acc_stride;
  act_t* pool_out;
  long long pool_plane, lead_o;
  int Wpo, Hpo;
  int* err_flag;
  int dbg;
};

// One work item of the interleaved sequence.
struct Item2 {
  int stage;    // 0 / 1
  int n, s, u;  // page, strip, tile row index (stage 0: may be -1 = warm-up tile)
  int j;        // index within its stage's sequence
  int dep;      // stage 1: index of the stage-0 item holding its last R FIFO rows
  bool fresh;   // stage 0: first item of a strip segment (whole input window is new)
};

// The fixed issue order, derived arithmetically by every role.  Stage-1 tiles [L0, L0 + n_tiles) in (page, strip, row)
// order; a segment = a maximal run inside one strip; stage-0 sequence = per segment one warm-up tile, then its tiles.
struct Sched2 {
  int n_tiles, n_strips, n_ty;
  int i1, seg1;
  TileRef t1;
  int i0, l0;
  TileRef t0;
  bool warm_done;
  __device__ void init(long long L0, int n_tiles_, int n_strips_, int n_ty_) {
    n_tiles = n_tiles_; n_strips = n_strips_; n_ty = n_ty_;
    i1 = 0; seg1 = 0; i0 = 0; l0 = 0; warm_done = false;
    t1 = tile_ref(L0, n_strips, n_ty);
    t0 = t1;
  }
  __device__ static void advance(TileRef& t, int n_strips, int n_ty) {
    if (++t.ty == n_ty) { t.ty = 0; if (++t.s == n_strips) { t.s = 0; ++t.n; } }
  }
  __device__ bool next(Item2& it) {
    if (i1 >= n_tiles) return false;
    const bool seg_start1 = (i1 == 0) || (t1.ty == 0);
    const int dep = i1 + seg1 + (seg_start1 ? 1 : 0);
    if (l0 < n_tiles && i0 <= dep + 1) {
      it.stage = 0; it.n = t0.n; it.s = t0.s; it.j = i0; it.dep = 0;
      if (!warm_done && (l0 == 0 || t0.ty == 0)) {
        it.u = t0.ty - 1; it.fresh = true; warm_done = true;
      } else {
        it.u = t0.ty; it.fresh = false; warm_done = false;
        ++l0;
        advance(t0, n_strips, n_ty);
      }
      ++i0;
    } else {
      it.stage = 1; it.n = t1.n; it.s = t1.s; it.u = t1.ty; it.j = i1; it.dep = dep; it.fresh = false;
      if (seg_start1) ++seg1;
      ++i1;
      advance(t1, n_strips, n_ty);
    }
    return true;
  }
};

// The MMA steps of one tile (3x3): A = windows of the banded masters (start row i * cop), B = ring rows.
// NCP = channel-chunk pairs per tap (C_in / 16), 0 for C_in = 8 (tap pairs).  b_row: 16 B unit address of the window's
// first ring row; the ring spans [ring_lo, ring_end).
template <int NCP>
__device__ __forceinline__ void band2_mma_tile(uint32_t d_tmem, uint32_t a_lo0, uint32_t master_units, uint32_t cop,
                                               uint32_t b_row, uint32_t slot_units, uint32_t row_units, uint32_t ring_end,
                                               uint32_t ring_span, int rows_win, uint32_t idesc, uint32_t hi) {
  uint32_t a_row = a_lo0;
  uint32_t acc = 0;
  if constexpr (NCP == 0) {
    const uint32_t lbo1 = 1u << 16, lbo_row = slot_units << 16;
#pragma unroll 1
    for (int i = 0; i < rows_win; i += 2) {   // rows_win = R + 2 is even; windows start on even slots
      const uint32_t b_nxt = b_row + slot_units;
      umma_f16(d_tmem, desc64(hi, a_row), desc64(hi, b_row | lbo1), idesc, acc);
      umma_f16(d_tmem, desc64(hi, a_row + cop), desc64(hi, b_nxt | lbo1), idesc, 1u);
      umma_f16(d_tmem, desc64(hi, a_row + master_units), desc64(hi, (b_row + 2u) | lbo_row), idesc, 1u);
      acc = 1u;
      a_row += 2u * cop;
      b_row = b_nxt + slot_units;
      if (b_row >= ring_end) b_row -= ring_span;
    }
  } else {
    const uint32_t lbo = row_units << 16;
#pragma unroll 1
    for (int i = 0; i < rows_win; ++i) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
        for (int cp = 0; cp < NCP; ++cp) {
          umma_f16(d_tmem, desc64(hi, a_row + (uint32_t)(kx * NCP + cp) * master_units),
                   desc64(hi, (b_row + (uint32_t)kx + (uint32_t)(2 * cp) * row_units) | lbo), idesc, acc);
          acc = 1u;
        }
      }
      a_row += cop;
      b_row += slot_units;
      if (b_row >= ring_end) b_row -= ring_span;
    }
  }
}

template <int NCP0, int NCP1, bool POOL>
__global__ void __launch_bounds__(NUM_THREADS2, 1) k_conv_band2(const __grid_constant__ ConvBand2Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_infull = s_base + O2_INFULL, s_indone = s_base + O2_INDONE;
  const uint32_t s_tfull = s_base + O2_TFULL, s_tempty = s_base + O2_TEMPTY;
  const uint32_t s_sfull = s_base + O2_SFULL, s_sempty = s_base + O2_SEMPTY;
  const uint32_t s_midfull = s_base + O2_MIDFULL, s_midraw = s_base + O2_MIDRAW;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + O2_TMEMPTR);
  volatile int* qs_ring = reinterpret_cast<volatile int*>(smem + O2_QS);
  float* s_bias = reinterpret_cast<float*>(smem + O2_BIAS);
  const int m0_al = ((a.masters0_bytes + 127) / 128) * 128, m1_al = ((a.masters1_bytes + 127) / 128) * 128;
  uint8_t* masters0 = smem + HDR2_BYTES;
  uint8_t* masters1 = masters0 + m0_al;
  uint8_t* slabs = masters1 + m1_al;
  uint8_t* mid = slabs + N_SLAB2 * SLAB2_BYTES;
  uint8_t* ring = mid + (size_t)a.RS * a.mid_slot_units * 16;
  const uint32_t s_m0 = smem_u32(masters0), s_m1 = smem_u32(masters1), s_mid = smem_u32(mid), s_ring = smem_u32(ring);
  const int rows_win = a.R + 2;

  const long long base = a.tiles / gridDim.x, rem = a.tiles % gridDim.x;
  const long long L0 = (long long)blockIdx.x * base + min((long long)blockIdx.x, rem);
  const int n_tiles = (int)(base + ((long long)blockIdx.x < rem ? 1 : 0));

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(s_infull + 8 * i, 1);
      mbar_init(s_indone + 8 * i, 1);
      mbar_init(s_tfull + 8 * i, 1);
      mbar_init(s_tempty + 8 * i, 4);
      mbar_init(s_sfull + 8 * i, 4);
      mbar_init(s_sempty + 8 * i, 8);
      mbar_init(s_midfull + 8 * i, 4);
      mbar_init(s_midraw + 8 * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += NUM_THREADS2) s_bias[i] = (i < 128 ? a.bias0 : a.bias1)[(i & 127) % a.cop];
  {
    const uint4* w0 = reinterpret_cast<const uint4*>(a.wpack0);
    uint4* d0 = reinterpret_cast<uint4*>(masters0);
    for (int i = threadIdx.x; i < a.masters0_bytes / 16; i += NUM_THREADS2) d0[i] = __ldg(w0 + i);
    const uint4* w1 = reinterpret_cast<const uint4*>(a.wpack1);
    uint4* d1 = reinterpret_cast<uint4*>(masters1);
    for (int i = threadIdx.x; i < a.masters1_bytes / 16; i += NUM_THREADS2) d1[i] = __ldg(w1 + i);
    // both rings start as zeros: rows / units that are never written must be finite (they meet zero band rows) and the
    // two trailing FIFO units of every row are the zero columns right of the computed range
    uint4* r4 = reinterpret_cast<uint4*>(mid);
    const int vecs = a.RS * a.mid_slot_units + a.S * a.in_slot_units;
    for (int i = threadIdx.x; i < vecs; i += NUM_THREADS2) r4[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t acc1_col = (uint32_t)(a.nbuf0 * a.acc_stride);

  if (n_tiles > 0) {
    if (warp == 0) {
      if (lane == 0) {
        // ================= MMA issuer =================
        const uint32_t idesc = (1u << 4) | (ARU_UMMA_FMT << 7) | (ARU_UMMA_FMT << 10) | ((uint32_t)(a.N >> 3) << 17) |
                               ((128u >> 4) << 24);
        const uint32_t hi = desc_hi128();
        const uint32_t a_lbo = (uint32_t)(a.J * a.cop);
        const uint32_t a0_lo = (s_m0 >> 4) | (a_lbo << 16), a1_lo = (s_m1 >> 4) | (a_lbo << 16);
        const uint32_t master_units = (uint32_t)(2 * a.J * a.cop);
        const uint32_t cop = (uint32_t)a.cop;
        const uint32_t ring_lo = s_ring >> 4, in_slot = (uint32_t)a.in_slot_units;
        const uint32_t ring_span = (uint32_t)a.S * in_slot, ring_end = ring_lo + ring_span;
        const uint32_t mid_lo = s_mid >> 4, mid_slot = (uint32_t)a.mid_slot_units;
        const uint32_t mid_span = (uint32_t)a.RS * mid_slot, mid_end = mid_lo + mid_span;
        int q0 = 0;
        long long c_te0 = 0, c_in = 0, c_te1 = 0, c_mid = 0, c_tot = BAND_CLK(), c_m0 = 0, c_m1 = 0;
        Sched2 sc;
        sc.init(L0, n_tiles, a.n_strips, a.n_ty);
        Item2 it;
        while (sc.next(it)) {
          if (it.stage == 0) {
            const int j = it.j;
            const int b = a.nbuf0 == 2 ? (j & 1) : 0;
            const uint32_t use = a.nbuf0 == 2 ? (uint32_t)(j >> 1) : (uint32_t)j;
            if (j > 0) {
              q0 += it.fresh ? rows_win : a.R;
              if (q0 >= a.S) q0 -= a.S;
            }
            long long c0 = BAND_CLK();
            mbar_wait(s_tempty + 8 * b, (use & 1u) ^ 1u, a.err_flag, 2);
            long long c1 = BAND_CLK();
            mbar_wait(s_infull + 8 * (j & 3), (uint32_t)(j >> 2) & 1u, a.err_flag, 3);
            long long c2 = BAND_CLK();
            c_te0 += c1 - c0; c_in += c2 - c1;
            tc_fence_after();
            band2_mma_tile<NCP0>(tmem_base + (uint32_t)(b * a.acc_stride), a0_lo, master_units, cop,
                                 ring_lo + (uint32_t)q0 * in_slot, in_slot, (uint32_t)a.in_row_units, ring_end, ring_span,
                                 rows_win, idesc, hi);
            umma_commit(s_tfull + 8 * b);
            umma_commit(s_indone + 8 * (j & 3));
            c_m0 += BAND_CLK() - c2;
          } else {
            const int j = it.j;
            const int b = a.nbuf1 == 2 ? (j & 1) : 0;
            const uint32_t use = a.nbuf1 == 2 ? (uint32_t)(j >> 1) : (uint32_t)j;
            long long c0 = BAND_CLK();
            mbar_wait(s_tempty + 8 * (2 + b), (use & 1u) ^ 1u, a.err_flag, 7);
            long long c1 = BAND_CLK();
            mbar_wait(s_midfull + 8 * (it.dep & 3), (uint32_t)(it.dep >> 2) & 1u, a.err_flag, 8);
            long long c2 = BAND_CLK();
            c_te1 += c1 - c0; c_mid += c2 - c1;
            tc_fence_after();
            const int slot = (it.dep * a.R - 2) % a.RS;
            band2_mma_tile<NCP1>(tmem_base + acc1_col + (uint32_t)(b * a.acc_stride), a1_lo, master_units, cop,
                                 mid_lo + (uint32_t)slot * mid_slot, mid_slot, (uint32_t)a.mid_row_units, mid_end, mid_span,
                                 rows_win, idesc, hi);
            umma_commit(s_tfull + 8 * (2 + b));
            c_m1 += BAND_CLK() - c2;
          }
        }
        if (a.dbg & 16) {
          unsigned long long* g = g_band2_stats[blockIdx.x];
          g[0] = BAND_CLK() - c_tot; g[1] = c_te0; g[2] = c_in; g[3] = c_te1; g[4] = c_mid; g[5] = c_m0; g[6] = c_m1;
        }
      }
    } else if (warp == 1) {
      // ================= producer: input rows of the stage-0 items =================
      const uint32_t row_bytes = (uint32_t)a.in_row_units * 16u;
      const uint32_t tx = row_bytes * (uint32_t)a.cin_chunks;
      int q = 0, slot = 0, items_done = 0, free_upto = 0, j = 0;
      long long c_pd = 0;
      Sched2 sc;
      sc.init(L0, n_tiles, a.n_strips, a.n_ty);
      Item2 it;
      while (sc.next(it)) {
        if (it.stage != 0) continue;
        const int i0 = it.fresh ? 0 : 2;
        const int n_new = rows_win - i0;
        if (lane == 0) qs_ring[j & 7] = q - i0;
        __syncwarp();
        const int y_first = it.u * a.R;                 // image row of window row 0 (input rows u R .. u R + R + 1)
        const int x_first = it.s * a.n_out - 2;         // image column of slot unit 0
        const int i_lo = max(i0, -1 - y_first);         // rows above the top frame row are not loaded
        const int i_end = min(rows_win, a.H + 1 - y_first);
        while (q + n_new - 1 - a.S >= free_upto) {
          long long c0 = BAND_CLK();
          mbar_wait_sleep(s_indone + 8 * (items_done & 3), (uint32_t)(items_done >> 2) & 1u, a.err_flag, 1);
          c_pd += BAND_CLK() - c0;
          ++items_done;
          free_upto = (items_done <= j) ? qs_ring[items_done & 7] : q;
        }
        const uint32_t bar = s_infull + 8 * (j & 3);
        if (lane == 0) mbar_expect_tx(bar, tx * (uint32_t)max(i_end - i_lo, 0));
        __syncwarp();
        for (int i = i_lo + lane; i < i_end; i += 32) {
          int sl = slot + (i - i0);
          if (sl >= a.S) sl -= a.S;
          const long long pos = a.lead + ((long long)it.n * a.Hp + (y_first + i) + 1) * a.Wp + (x_first + 1);
          const uint32_t dst = s_ring + (uint32_t)sl * (uint32_t)a.in_slot_units * 16u;
          for (int c = 0; c < a.cin_chunks; ++c)
            bulk_g2s(dst + (uint32_t)c * row_bytes, a.in + ((long long)c * a.in_plane + pos) * 8, row_bytes, bar);
        }
        q += n_new;
        slot += n_new;
        if (slot >= a.S) slot -= a.S;
        ++j;
      }
      if (lane == 0 && (a.dbg & 16)) g_band2_stats[blockIdx.x][12] = c_pd;
    } else if (warp >= 4 && warp < 8) {
      // ================= drain of stage 1: TMEM -> + bias -> 16-bit -> transposition slab =================
      const int q4 = warp & 3;
      const int n_slabs = (a.N + SLAB2 - 1) / SLAB2;
      const int m = q4 * 32 + lane;
      const float bias = s_bias[128 + m];
      const int st_off = (m >> 3) * SLAB2_ROW + (m & 7) * 2;
      int sb_i = 0;
      uint32_t bpar = 0;
      long long c_tf = 0, c_se = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const int b = a.nbuf1 == 2 ? (j & 1) : 0;
        const uint32_t use = a.nbuf1 == 2 ? (uint32_t)(j >> 1) : (uint32_t)j;
        long long c0 = BAND_CLK();
        mbar_wait_sleep(s_tfull + 8 * (2 + b), use & 1u, a.err_flag, 4);
        c_tf += BAND_CLK() - c0;
        __syncwarp();
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc1_col + (uint32_t)(b * a.acc_stride);
        uint32_t v[32];
        tmem_ld32(taddr, v);
#pragma unroll 1
        for (int sb = 0; sb < n_slabs; ++sb) {
          long long c1 = BAND_CLK();
          mbar_wait_sleep(s_sempty + 8 * sb_i, bpar ^ 1u, a.err_flag, 5);
          c_se += BAND_CLK() - c1;
          uint8_t* my_st = slabs + sb_i * SLAB2_BYTES + st_off;
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const uint32_t h2 = pack2_raw(__uint_as_float(v[k]) + bias, __uint_as_float(v[k + 1]) + bias);
            *reinterpret_cast<unsigned short*>(my_st + k * 16) = (unsigned short)(h2 & 0xffffu);
            *reinterpret_cast<unsigned short*>(my_st + (k + 1) * 16) = (unsigned short)(h2 >> 16);
          }
          if (sb + 1 < n_slabs) {
            tmem_ld32(taddr + (uint32_t)((sb + 1) * SLAB2), v);
            __syncwarp();
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tempty + 8 * (2 + b));
          }
          if (lane == 0) mbar_arrive(s_sfull + 8 * sb_i);
          if (++sb_i == N_SLAB2) { sb_i = 0; bpar ^= 1u; }
        }
      }
      if (warp == 4 && lane == 0 && (a.dbg & 16)) { g_band2_stats[blockIdx.x][7] = c_tf; g_band2_stats[blockIdx.x][8] = c_se; }
    } else if (warp >= 8 && warp < 12) {
      // ================= drain of stage 0: TMEM -> + bias -> activation / border mask -> FIFO rows =================
      // Lane (rho, co) owns 2 bytes of every 16-byte position of FIFO row R-1-rho: 32 lanes = 4 rows x 8 channels, rows
      // 8 banks apart ((N + 2) * 4 words, N a multiple of 16), so the 2-byte stores are conflict-free.
      const int q4 = warp & 3;
      const int n_slabs = (a.N + SLAB2 - 1) / SLAB2;
      const int m = q4 * 32 + lane;
      const float bias = s_bias[m];
      const int rho = m >> (3 + a.nc_shift), cch = (m >> 3) & ((1 << a.nc_shift) - 1);
      const int r = a.R - 1 - rho;                       // FIFO row of this lane within the item
      const bool raw_mode = a.pre0 != nullptr || a.out0 != nullptr;   // warps 20-23 finish the rows
      const bool relu0 = a.act0 == 1 && !raw_mode;
      const uint32_t lane_off = (uint32_t)cch * (uint32_t)a.mid_row_units * 16u + (uint32_t)(m & 7) * 2u;
      const uint32_t mid_slot_bytes = (uint32_t)a.mid_slot_units * 16u;
      long long c_tf0 = 0, c_d0 = BAND_CLK();
      Sched2 sc;
      sc.init(L0, n_tiles, a.n_strips, a.n_ty);
      Item2 it;
      while (sc.next(it)) {
        if (it.stage != 0) continue;
        const int j = it.j;
        const int b = a.nbuf0 == 2 ? (j & 1) : 0;
        const uint32_t use = a.nbuf0 == 2 ? (uint32_t)(j >> 1) : (uint32_t)j;
        const int y = it.u * a.R + 1 + r;
        const bool row_ok = y >= 0 && y < a.H;
        const int x0 = it.s * a.n_out - 1;               // image column of accumulator column 0
        uint8_t* row_p = mid + (size_t)((j * a.R + r) % a.RS) * mid_slot_bytes + lane_off;
        long long c0 = BAND_CLK();
        mbar_wait_sleep(s_tfull + 8 * b, use & 1u, a.err_flag, 9);
        c_tf0 += BAND_CLK() - c0;
        // The export warps rewrite the rows of item j - 2 in place, and this item's rows alias them.  In steady state the
        // issue order already puts that behind us (s0(j) is issued after s1(dep = j - 2), which waited for those warps);
        // at the start of a CTA or strip segment three stage-0 tiles are issued back to back, so wait explicitly.
        if (raw_mode && j >= 2) mbar_wait_sleep(s_midfull + 8 * ((j - 2) & 3), (uint32_t)((j - 2) >> 2) & 1u, a.err_flag, 11);
        __syncwarp();
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(b * a.acc_stride);
        uint32_t v[32];
        tmem_ld32(taddr, v);
#pragma unroll 1
        for (int sb = 0; sb < n_slabs; ++sb) {
          const int xs = x0 + sb * SLAB2;                // image column of v[0]
          const int k_lo = max(0, -xs), k_hi = min(min(SLAB2, a.W - xs), a.N - sb * SLAB2);
          uint8_t* dst = row_p + sb * SLAB2 * 16;
          tmem_ld_wait();
          if (row_ok && k_lo == 0 && k_hi == SLAB2) {
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const float f0 = __uint_as_float(v[k]) + bias, f1 = __uint_as_float(v[k + 1]) + bias;
              const uint32_t h2 = raw_mode ? pack2_raw(f0, f1) : (relu0 ? pack2_fin<true>(f0, f1) : pack2_fin<false>(f0, f1));
              *reinterpret_cast<unsigned short*>(dst + k * 16) = (unsigned short)(h2 & 0xffffu);
              *reinterpret_cast<unsigned short*>(dst + (k + 1) * 16) = (unsigned short)(h2 >> 16);
            }
          } else {
            // tiles that touch the page border: positions outside the page are zeros (SAME padding of the next conv)
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const float f0 = __uint_as_float(v[k]) + bias, f1 = __uint_as_float(v[k + 1]) + bias;
              const uint32_t h2 = raw_mode ? pack2_raw(f0, f1) : (relu0 ? pack2_fin<true>(f0, f1) : pack2_fin<false>(f0, f1));
              const bool ok0 = row_ok && k >= k_lo && k < k_hi, ok1 = row_ok && k + 1 >= k_lo && k + 1 < k_hi;
              if (sb * SLAB2 + k < a.N) *reinterpret_cast<unsigned short*>(dst + k * 16) = ok0 ? (unsigned short)(h2 & 0xffffu) : (unsigned short)0;
              if (sb * SLAB2 + k + 1 < a.N) *reinterpret_cast<unsigned short*>(dst + (k + 1) * 16) = ok1 ? (unsigned short)(h2 >> 16) : (unsigned short)0;
            }
          }
          __syncwarp();   // the border path diverges per lane (rows); tcgen05.ld is .sync.aligned
          if (sb + 1 < n_slabs) {
            tmem_ld32(taddr + (uint32_t)((sb + 1) * SLAB2), v);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tempty + 8 * b);
          }
        }
        if (!raw_mode) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tensor pipe reads these rows
        __syncwarp();
        if (lane == 0) mbar_arrive((raw_mode ? s_midraw : s_midfull) + 8 * (j & 3));
      }
      if (warp == 8 && lane == 0 && (a.dbg & 16)) { g_band2_stats[blockIdx.x][9] = c_tf0; g_band2_stats[blockIdx.x][10] = BAND_CLK() - c_d0; }
    } else if (warp >= 12 && warp < 20) {
      // ================= store of stage 1: slab -> (+ residual) -> activation -> global stores (+ max-pool) ==========
      const int n_slabs = (a.N + SLAB2 - 1) / SLAB2;
      const bool has_res = a.res != nullptr, relu1 = a.act1 == 1;
      const int nc_mask = (1 << a.nc_shift) - 1;
      const int pp = warp - 12;
      const int cch = pp & nc_mask;
      const int rho0 = 2 * (pp >> a.nc_shift);
      const int rs0 = (rho0 << a.nc_shift) + cch, rs1 = rs0 + (1 << a.nc_shift);
      const long long out_c = (long long)cch * a.out_plane * 8, res_c = (long long)cch * a.res_plane * 8;
      const long long row8 = (long long)a.Wp * 8;
      int sb_i = 0;
      uint32_t bpar = 0;
      TileRef tr = tile_ref(L0, a.n_strips, a.n_ty);
      long long c_sf = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const int x_left = tr.s * a.n_out;
        const int y_top = tr.ty * a.R;
        const int y1 = y_top + (a.R - 2 - rho0);
        const bool gv[2] = {y1 + 1 < a.H, y1 < a.H};
        const long long gp1 = (a.lead + ((long long)tr.n * a.Hp + y1 + 1) * a.Wp + (x_left + 1)) * 8;
        long long ppos = 0;
        if constexpr (POOL)
          ppos = ((long long)cch * a.pool_plane + a.lead_o + ((long long)tr.n * a.Hpo + (y1 >> 1) + 1) * a.Wpo + (x_left >> 1) + 1) * 8;
        TileRef nx = tr;
        Sched2::advance(nx, a.n_strips, a.n_ty);
        uint4 rr[2];
        if (has_res) {
          // residual of the NEXT tile -> L2 (one 128 B line per lane and row)
          if (j + 1 < n_tiles) {
            const int cols = min(a.n_out, a.W - nx.s * a.n_out);
            const int yn = nx.ty * a.R + (a.R - 2 - rho0);
            if (lane * 8 < cols) {
              const long long pos = a.lead + ((long long)nx.n * a.Hp + yn + 1) * a.Wp + (nx.s * a.n_out + 1) + lane * 8;
              if (yn < a.H) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + res_c + pos * 8));
              if (yn + 1 < a.H) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + res_c + (pos + a.Wp) * 8));
            }
          }
          const bool xv = (lane < a.n_out) && (x_left + lane < a.W);
          rr[0] = ldg_nc_v4(a.res + res_c + gp1 + row8 + (long long)lane * 8, xv && gv[0]);
          rr[1] = ldg_nc_v4(a.res + res_c + gp1 + (long long)lane * 8, xv && gv[1]);
        }
#pragma unroll 1
        for (int sb = 0; sb < n_slabs; ++sb) {
          const int xl = sb * SLAB2 + lane;
          const bool xv = (xl < a.n_out) && (x_left + xl < a.W);
          uint4 rn[2];
          if (has_res && sb + 1 < n_slabs) {
            const int xn = xl + SLAB2;
            const bool xnv = (xn < a.n_out) && (x_left + xn < a.W);
            rn[0] = ldg_nc_v4(a.res + res_c + gp1 + row8 + (long long)xn * 8, xnv && gv[0]);
            rn[1] = ldg_nc_v4(a.res + res_c + gp1 + (long long)xn * 8, xnv && gv[1]);
          }
          long long c2 = BAND_CLK();
          mbar_wait_sleep(s_sfull + 8 * sb_i, bpar, a.err_flag, 6);
          c_sf += BAND_CLK() - c2;
          const uint8_t* slab = slabs + sb_i * SLAB2_BYTES;
          uint4 raw[2];
          raw[0] = *reinterpret_cast<const uint4*>(slab + rs0 * SLAB2_ROW + lane * 16);
          raw[1] = *reinterpret_cast<const uint4*>(slab + rs1 * SLAB2_ROW + lane * 16);
          __syncwarp();
          if (lane == 0) mbar_arrive(s_sempty + 8 * sb_i);
          if (++sb_i == N_SLAB2) { sb_i = 0; bpar ^= 1u; }
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const long long p8 = gp1 + (k == 0 ? row8 : 0) + (long long)xl * 8;
            uint4 val = raw[k];
            if (has_res) val = add8(val, rr[k]);
            const uint4 fin = relu1 ? clamp8<true>(val) : clamp8<false>(val);
            if (xv && gv[k]) *reinterpret_cast<uint4*>(a.out + out_c + p8) = fin;
            if constexpr (POOL) raw[k] = (xv && gv[k]) ? fin : make_uint4(0u, 0u, 0u, 0u);
          }
          if constexpr (POOL) {
            uint4 mx = max8(raw[0], raw[1]);
            uint4 ot;
            ot.x = __shfl_xor_sync(0xffffffffu, mx.x, 1);
            ot.y = __shfl_xor_sync(0xffffffffu, mx.y, 1);
            ot.z = __shfl_xor_sync(0xffffffffu, mx.z, 1);
            ot.w = __shfl_xor_sync(0xffffffffu, mx.w, 1);
            mx = max8(mx, ot);
            if (!(lane & 1) && xv && gv[1])
              *reinterpret_cast<uint4*>(a.pool_out + ppos + (long long)(xl >> 1) * 8) = mx;
          }
          if (has_res) {
            rr[0] = rn[0];
            rr[1] = rn[1];
          }
        }
        tr = nx;
      }
      if (warp == 12 && lane == 0 && (a.dbg & 16)) g_band2_stats[blockIdx.x][11] = c_sf;
    } else if (warp >= 20 && (a.pre0 != nullptr || a.out0 != nullptr)) {
      // ================= stage-0 export: raw FIFO rows -> global copies -> activation in place =================
      const int fw = warp - 20;
      const bool relu0 = a.act0 == 1;
      const int nc = 1 << a.nc_shift;
      const uint32_t mid_row_bytes = (uint32_t)a.mid_row_units * 16u, mid_slot_bytes = (uint32_t)a.mid_slot_units * 16u;
      Sched2 sc;
      sc.init(L0, n_tiles, a.n_strips, a.n_ty);
      Item2 it;
      while (sc.next(it)) {
        if (it.stage != 0) continue;
        const int j = it.j;
        const int x0 = it.s * a.n_out - 1;
        mbar_wait_sleep(s_midraw + 8 * (j & 3), (uint32_t)(j >> 2) & 1u, a.err_flag, 10);
        for (int r = fw; r < a.R; r += 4) {
          const int y = it.u * a.R + 1 + r;
          const bool row_ok = y >= 0 && y < a.H;
          uint8_t* row_p = mid + (size_t)((j * a.R + r) % a.RS) * mid_slot_bytes;
          const long long gp = (a.lead + ((long long)it.n * a.Hp + y + 1) * a.Wp + (x0 + 1)) * 8;
          for (int c = 0; c < nc; ++c)
            for (int xl = lane; xl < a.N; xl += 32) {
              uint4* q = reinterpret_cast<uint4*>(row_p + (size_t)c * mid_row_bytes + (size_t)xl * 16);
              const uint4 raw = *q;
              const uint4 fin = relu0 ? clamp8<true>(raw) : clamp8<false>(raw);
              *q = fin;                                    // zeros (outside the page) stay zeros
              const int x = x0 + xl;
              if (row_ok && xl >= 1 && xl <= a.n_out && x < a.W) {   // the strip's own columns
                if (a.pre0) *reinterpret_cast<uint4*>(a.pre0 + (long long)c * a.pre0_plane * 8 + gp + (long long)xl * 8) = clamp8<false>(raw);
                if (a.out0) *reinterpret_cast<uint4*>(a.out0 + (long long)c * a.out0_plane * 8 + gp + (long long)xl * 8) = fin;
              }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(s_midfull + 8 * (j & 3));
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

size_t band_masters_bytes(int cin_chunks, int R, int cop) {
  const int J = 2 * R + 1;
  const int n_masters = cin_chunks == 1 ? 2 : 3 * (cin_chunks / 2);
  return (size_t)n_masters * 2 * J * cop * 16;
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------------
ConvBand2Plan conv_band2_plan(int cin, int cmid, int cout, const Geo& g, int num_sms, size_t max_smem) {
  ConvBand2Plan p;
  p.cin_chunks = cdiv(cin, 8);
  if (cmid != cout) { p.why = "the two stages must have the same output channels"; return p; }
  int cop = 8;
  while (cop < cout) cop <<= 1;
  if (cop > 16 || cdiv(cout, 8) * 8 != cop) { p.why = "fused pairs cover C_out = 8 / 16"; return p; }
  if (p.cin_chunks != 1 && p.cin_chunks != 2 && p.cin_chunks != 4) { p.why = "C_in chunks must be 1, 2 or 4"; return p; }
  p.cop = cop;
  p.R = 128 / cop;
  p.J = 2 * p.R + 1;
  p.RS = 2 * p.R + 2;
  const int rows_win = p.R + 2;
  const int nc = cop / 8;
  // stage masters in conv_band_pack's layout
  p.st0.ok = true; p.st0.ks = 3; p.st0.cin_chunks = p.cin_chunks; p.st0.cop = cop; p.st0.R = p.R; p.st0.J = p.J;
  p.st0.n_masters = p.cin_chunks == 1 ? 2 : 3 * (p.cin_chunks / 2);
  p.st0.wpack_bytes = band_masters_bytes(p.cin_chunks, p.R, cop);
  p.st1 = p.st0;
  p.st1.cin_chunks = nc;
  p.st1.n_masters = nc == 1 ? 2 : 3 * (nc / 2);
  p.st1.wpack_bytes = band_masters_bytes(nc, p.R, cop);
  const size_t fixed = HDR2_BYTES + ((p.st0.wpack_bytes + 127) / 128) * 128 + ((p.st1.wpack_bytes + 127) / 128) * 128 +
                       N_SLAB2 * SLAB2_BYTES + 128;
  p.why = "rings do not fit in shared memory";
  int best_pref = -1;
  for (int N = 128; N >= 48; N -= 16) {
    const size_t mid = (size_t)p.RS * nc * (N + 2) * 16;
    const size_t slot = (size_t)p.cin_chunks * (N + 2) * 16;
    if (fixed + mid >= max_smem) continue;
    int S = (int)std::min<size_t>((max_smem - fixed - mid) / slot, 64);
    S = std::min(S, rows_win + 2 * p.R);
    S &= ~1;
    if (S < rows_win + 2) continue;
    // prefer a ring that prefetches at least half a tile of rows; otherwise the widest that fits at all
    const int pref = S >= rows_win + p.R / 2 ? 1 : 0;
    if (pref > best_pref) {
      best_pref = pref;
      p.ok = true;
      p.N = N;
      p.S = S;
      p.smem_bytes = fixed + mid + (size_t)S * slot;
      if (pref == 1) break;
    }
  }
  if (!p.ok) return p;
  p.why = "";
  // as few strips as the widest N allows, then the narrowest N (multiple of 16) that covers them
  p.n_strips = cdiv(g.W, p.N - 2);
  {
    const int need = cdiv(g.W, p.n_strips) + 2;
    const int N2 = std::max(48, cdiv(need, 16) * 16);
    if (N2 < p.N) {
      const size_t mid = (size_t)p.RS * nc * (N2 + 2) * 16;
      const size_t slot = (size_t)p.cin_chunks * (N2 + 2) * 16;
      int S = (int)std::min<size_t>((max_smem - fixed - mid) / slot, 64);
      S = std::min(S, rows_win + 2 * p.R) & ~1;
      p.N = N2; p.S = S; p.smem_bytes = fixed + mid + (size_t)S * slot;
    }
  }
  p.n_out = p.N - 2;
  if (p.n_out & 1) { p.ok = false; p.why = "odd strip width"; return p; }
  p.n_strips = cdiv(g.W, p.n_out);
  p.n_ty = cdiv(g.H, p.R);
  p.tiles = (long long)g.N * p.n_strips * p.n_ty;
  p.grid = (int)std::min<long long>(p.tiles, num_sms);
  p.nbuf0 = 2; p.nbuf1 = 2;
  p.acc_stride = 128;
  return p;
}

cudaError_t launch_conv_band2(cudaStream_t st, const ConvBand2Plan& p, PV in, PV out0, PV pre0, PV out, PV res,
                              const act_t* wpack0, const float* bias0, int act0, const act_t* wpack1, const float* bias1,
                              int act1, const Geo& g, int* err_flag, PV pool, const Geo* pool_geo) {
  if (!p.ok) return cudaErrorInvalidValue;
  ConvBand2Args a{};
  if (pool.p) {
    if (!pool_geo || act1 != 1 || p.R < 2 || pool_geo->H != (g.H + 1) / 2 || pool_geo->W != (g.W + 1) / 2)
      return cudaErrorInvalidValue;
    a.pool_out = pool.p; a.pool_plane = pool.plane; a.lead_o = pool_geo->lead;
    a.Wpo = pool_geo->Wp; a.Hpo = pool_geo->Hp;
  }
  a.in = in.p; a.in_plane = in.plane;
  a.out0 = out0.p; a.out0_plane = out0.plane;
  a.pre0 = pre0.p; a.pre0_plane = pre0.plane;
  a.out = out.p; a.out_plane = out.plane;
  a.res = res.p; a.res_plane = res.plane;
  a.wpack0 = wpack0; a.wpack1 = wpack1; a.bias0 = bias0; a.bias1 = bias1;
  a.cin_chunks = p.cin_chunks; a.cop = p.cop; a.R = p.R; a.N = p.N; a.n_out = p.n_out;
  a.nc_shift = 0;
  while ((8 << a.nc_shift) < p.cop) ++a.nc_shift;
  a.n_strips = p.n_strips; a.n_ty = p.n_ty; a.S = p.S; a.RS = p.RS;
  a.J = p.J; a.masters0_bytes = (int)p.st0.wpack_bytes; a.masters1_bytes = (int)p.st1.wpack_bytes;
  a.in_row_units = p.N + 2;
  a.in_slot_units = p.cin_chunks * a.in_row_units;
  a.mid_row_units = p.N + 2;
  a.mid_slot_units = (p.cop / 8) * a.mid_row_units;
  a.W = g.W; a.H = g.H; a.Wp = g.Wp; a.Hp = g.Hp;
  a.lead = g.lead;
  a.tiles = p.tiles;
  a.act0 = act0; a.act1 = act1;
  a.nbuf0 = p.nbuf0; a.nbuf1 = p.nbuf1; a.acc_stride = p.acc_stride;
  a.err_flag = err_flag;
  { const char* e = getenv("ARU_BAND_DBG"); a.dbg = e ? atoi(e) : 0; }
  using Kern = void (*)(const __grid_constant__ ConvBand2Args);
  const int ncp0 = p.cin_chunks == 1 ? 0 : p.cin_chunks / 2;
  const int ncp1 = p.cop == 8 ? 0 : p.cop / 16;
  const bool pl = pool.p != nullptr;
  Kern k = nullptr;
  int ki = -1;
#define ARU_BAND2_PICK(N0, N1, PP, II) \
  if (ncp0 == N0 && ncp1 == N1 && pl == PP) { k = k_conv_band2<N0, N1, PP>; ki = II; }
  ARU_BAND2_PICK(0, 0, false, 0) ARU_BAND2_PICK(1, 0, false, 1) ARU_BAND2_PICK(0, 1, false, 2)
  ARU_BAND2_PICK(1, 1, false, 3) ARU_BAND2_PICK(2, 1, false, 4)
  ARU_BAND2_PICK(0, 0, true, 5) ARU_BAND2_PICK(1, 0, true, 6) ARU_BAND2_PICK(0, 1, true, 7)
  ARU_BAND2_PICK(1, 1, true, 8) ARU_BAND2_PICK(2, 1, true, 9)
#undef ARU_BAND2_PICK
  if (!k) return cudaErrorInvalidValue;
  static bool configured[10] = {};
  if (!configured[ki]) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured[ki] = true;
  }
  k<<<p.grid, NUM_THREADS2, p.smem_bytes, st>>>(a);
  if (a.dbg & 16) {
    cudaDeviceSynchronize();
    static unsigned long long h[160][16];
    cudaMemcpyFromSymbol(h, g_band2_stats, sizeof(h));
    double s[16] = {0};
    for (int i = 0; i < p.grid; ++i) for (int j = 0; j < 16; ++j) s[j] += (double)h[i][j] / p.grid;
    fprintf(stderr, "band2 %dx%d cin%d cop%d N%d S%d tiles/cta %.1f res%d pre%d | issuer tot %.0f: tempty0 %.0f infull %.0f mma0 %.0f "
            "tempty1 %.0f midfull %.0f mma1 %.0f | drain1: tfull %.0f sempty %.0f | drain0: tfull %.0f tot %.0f | store sfull %.0f | producer indone %.0f\n",
            g.H, g.W, p.cin_chunks * 8, p.cop, p.N, p.S, (double)p.tiles / p.grid, a.res != nullptr, a.pre0 != nullptr, s[0], s[1], s[2], s[5],
            s[3], s[4], s[6], s[7], s[8], s[9], s[10], s[11], s[12]);
  }
  return cudaGetLastError();
}

}  // namespace aru
