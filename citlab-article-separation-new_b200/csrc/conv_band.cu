// conv_band.cu - "row-banded weights" tcgen05 convolution for the small-channel layers (C_out <= 64) on sm_100a.
//
// Why a second tensor-core kernel.  conv_tc.cu puts 128 *positions* on the MMA M axis and C_out on N.  With the A operand
// in shared memory one M=128,K=16 MMA costs >= 39 cycles whatever N is (the 4 KB operand read), so a C_in = C_out = 8
// layer pays 6 x ~50 cycles per 128 positions, more than its HBM time (tools/mma_issue_bench.cu): those
// layers were tensor-issue bound at ~55 % of the HBM roofline.  This kernel turns the GEMM around:
//
//   D[m = (rho, co), n = x] = sum_steps  A_step[m, k] * B_step[k, n]
//     n     : N (<= 256) consecutive output columns of one strip of the page           (MMA N, 112-128 cycles per MMA)
//     m     : R = 128 / C_out output rows x C_out channels, rho = R-1-r               (MMA M = 128, always full)
//     step  : one input row i of the R+KS-1 rows under the tile x one 16-wide K block (kx pair / channel-chunk pair)
//     B_step: the activations of input row i, K-major, straight out of the row ring   (positions x 16 B, any 16 B shift)
//     A_step: a 128-row *window* of a banded master matrix T_b[(j, co), k] = W[ky = j-(R-1), ...] (zero outside the
//             band); the window of input row i starts at row i*C_out, i.e. the band structure of "input row i feeds
//             output rows i-2..i" is just a descriptor start address - the masters are tiny and stay resident.
//
// A 3x3 C->C layer then needs (R+2)*3*C/16 MMAs of ~N/2 cycles per R*N outputs: 108 (C=8), 240 (C=16), 576 (C=32) tensor
// cycles per 128 outputs against 183 / 366 / 731 cycles of HBM time - every small-C layer becomes HBM bound.
//
// Pipeline (384 threads, one persistent CTA per SM, tiles walked down a strip so the KS-1 halo rows are reused):
//   warp 1 lane 0 : producer - one bulk copy (TMA engine) per input row and channel chunk into a ring of row slots
//   warp 0        : MMA issuer (elected lane), steps generated arithmetically; tcgen05.commit -> accumulator full / tile done
//   warp 2        : TMEM allocator (2 stages x 256 columns)
//   warps 4..11   : two epilogue warpgroups (even / odd tiles).  Per 32-column slab: tcgen05.ld (lane = (rho, co)) ->
//                   + bias -> 16-bit -> transposed through a padded shared-memory slab into position-major 16 B
//                   vectors -> (+ residual) (ReLU) -> 512 B coalesced stores per warp.
// Reference semantics: layers.py:191-247 (conv2d SAME + bias + activation), ARU_v1.py:212-227 (residual block).
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <vector>

#include "aru_common.cuh"
#include "kernels.h"
#include "band_common.cuh"

namespace aru {

namespace {

constexpr int NUM_THREADS = 512;
constexpr int MAX_SLOTS = 64;
constexpr int SLAB = 32;                       // output columns per epilogue slab
constexpr int SLAB_ROW = (SLAB + 1) * 16;      // bytes per (rho, chunk) row of a slab (+16 B: conflict-free 2-byte stores)
constexpr int SLAB_BYTES = 16 * SLAB_ROW;      // 16 row segments = 128 accumulator rows / 8
// shared-memory header (bytes)
constexpr int OFF_FULL = 0;                               // uint64 full[4]: the new rows of tile t land on full[t & 3]
constexpr int OFF_DONE = OFF_FULL + 32;                   // uint64 done[4]
constexpr int OFF_TFULL = OFF_DONE + 32;                  // uint64 tfull[2]
constexpr int OFF_TEMPTY = OFF_TFULL + 16;                // uint64 tempty[2]
constexpr int N_SLAB_BUF = 4;                             // transposition slabs between the drain and the store warps
constexpr int OFF_SFULL = OFF_TEMPTY + 16;                // uint64 sfull[4]
constexpr int OFF_SEMPTY = OFF_SFULL + 32;                // uint64 sempty[4]
constexpr int OFF_TMEMPTR = OFF_SEMPTY + 32;              // uint32
constexpr int OFF_QS = OFF_TMEMPTR + 16;                  // int qs[8]: first ring row of the producer's last 8 tiles
constexpr int OFF_BIAS = OFF_QS + 32;                     // float bias[128]
constexpr int HDR_BYTES = ((OFF_BIAS + 512 + 127) / 128) * 128;

struct ConvBandArgs {
  const act_t* in;
  long long in_plane;
  act_t* out;
  long long out_plane;
  act_t* out_pre;
  long long pre_plane;
  const act_t* res;
  long long res_plane;
  const act_t* wpack;
  const float* bias;
  int ks, cin_chunks, cop, nc_shift, R, N, n_strips, n_ty, S;
  int n_steps, J, masters_bytes;
  int slot_units;                 // 16 B units per row slot (all chunks)
  int row_units;                  // 16 B units per chunk row = N + ks - 1
  int W, H, Wp, Hp, NP;
  long long lead;
  long long tiles;
  int act;
  // fused 2x2 stride-2 max-pool of the (ReLU'd) output into a second chunk-planar tensor, else null
  act_t* pool_out;
  long long pool_plane, lead_o;
  int Wpo, Hpo;
  // dense float32 NHWC output [N][H][W][f32_c] with the head activation (classifier / attention logit), else null
  float* out_f32;
  int f32_c;
  int* err_flag;
  int dbg;
};

// KS: filter size.  NCP: channel-chunk pairs per tap (C_in / 16), 0 for C_in = 8 (tap pairs instead).
// HEAD: dense float32 NHWC output with the head activation instead of 16-bit planes (separate instances keep the
// plane path free of its registers and branches).
// POOL: fused max-pool output (plane path only).
template <int KS, int NCP, bool HEAD, bool POOL>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv_band(const __grid_constant__ ConvBandArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_full = s_base + OFF_FULL, s_done = s_base + OFF_DONE;
  const uint32_t s_tfull = s_base + OFF_TFULL, s_tempty = s_base + OFF_TEMPTY;
  const uint32_t s_sfull = s_base + OFF_SFULL, s_sempty = s_base + OFF_SEMPTY;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEMPTR);
  volatile int* qs_ring = reinterpret_cast<volatile int*>(smem + OFF_QS);
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  const int masters_al = ((a.masters_bytes + 127) / 128) * 128;
  uint8_t* masters = smem + HDR_BYTES;
  uint8_t* slabs = masters + masters_al;
  uint8_t* ring = slabs + 4 * SLAB_BYTES;
  const uint32_t s_masters = smem_u32(masters), s_ring = smem_u32(ring);
  const int rows_win = a.R + a.ks - 1;   // input rows under one tile

  // contiguous tile range of this CTA
  const long long base = a.tiles / gridDim.x, rem = a.tiles % gridDim.x;
  const long long L0 = (long long)blockIdx.x * base + min((long long)blockIdx.x, rem);
  const int n_tiles = (int)(base + ((long long)blockIdx.x < rem ? 1 : 0));

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(s_full + 8 * i, 1);
    for (int i = 0; i < 4; ++i) mbar_init(s_done + 8 * i, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_tfull + 8 * i, 1);
      mbar_init(s_tempty + 8 * i, 4);
    }
    for (int i = 0; i < N_SLAB_BUF; ++i) {
      mbar_init(s_sfull + 8 * i, 4);
      mbar_init(s_sempty + 8 * i, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 128; i += NUM_THREADS) s_bias[i] = a.bias[i % a.cop];
  {
    const uint4* wsrc = reinterpret_cast<const uint4*>(a.wpack);
    uint4* wdst = reinterpret_cast<uint4*>(masters);
    for (int i = threadIdx.x; i < a.masters_bytes / 16; i += NUM_THREADS) wdst[i] = __ldg(wsrc + i);
    // The ring must never hold a NaN / inf bit pattern: rows that are not loaded (below the page) still meet the zero
    // rows of the band, and 0 * NaN would poison real outputs.
    uint4* r4 = reinterpret_cast<uint4*>(ring);
    const int ring_vecs = a.S * a.slot_units;
    for (int i = threadIdx.x; i < ring_vecs; i += NUM_THREADS) r4[i] = make_uint4(0u, 0u, 0u, 0u);
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

  if (n_tiles > 0) {
    if (warp == 0) {
      if (lane == 0) {
      // ================= MMA issuer (one thread) =================
      // tcgen05.mma issue is effectively synchronous (tools/mma_issue_bench.cu: no deep queue), so every instruction
      // between two MMAs is exposed: the step loops are unrolled at compile time and only add constants.
      const uint32_t idesc = (1u << 4) | (ARU_UMMA_FMT << 7) | (ARU_UMMA_FMT << 10) | ((uint32_t)(a.N >> 3) << 17) |
                             ((128u >> 4) << 24);
      const uint32_t hi = desc_hi128();
      const uint32_t a_lbo = (uint32_t)(a.J * a.cop);            // 16 B units between the K halves of a master
      const uint32_t a_lo0 = (s_masters >> 4) | (a_lbo << 16);
      const uint32_t ring_lo = s_ring >> 4;
      const uint32_t master_units = (uint32_t)(2 * a.J * a.cop);   // 16 B units per banded master
      const uint32_t cop = (uint32_t)a.cop, slot_units = (uint32_t)a.slot_units, row_units = (uint32_t)a.row_units;
      const uint32_t ring_end = ring_lo + (uint32_t)a.S * slot_units;
      int q0 = 0;          // ring slot of input row 0 of the current tile
      TileRef tr = tile_ref(L0, a.n_strips, a.n_ty);
      long long c_tempty = 0, c_full = 0, c_tot = BAND_CLK();
      for (int t = 0; t < n_tiles; ++t) {
        const int stage = t & 1;
        long long c0 = BAND_CLK();
        mbar_wait(s_tempty + 8 * stage, ((uint32_t)(t >> 1) & 1u) ^ 1u, a.err_flag, 2);
        c_tempty += BAND_CLK() - c0;
        const uint32_t d_tmem = tmem_base + (uint32_t)(stage * 256);
        {
          // One wait per tile: per-row barriers cost the issuer ~200 cycles each (shared-memory port contention with the
          // epilogue), and the ring is a tile ahead in steady state anyway.
          long long c1 = BAND_CLK();
          mbar_wait(s_full + 8 * (t & 3), (uint32_t)(t >> 2) & 1u, a.err_flag, 3);
          c_full += BAND_CLK() - c1;
          tc_fence_after();
        }
        // K steps in input-row order.  A = the window of a banded master starting at row i * cop; B = the row slot of
        // input row i (+ kx, + channel-chunk plane).
        uint32_t a_row = a_lo0;
        uint32_t b_row = ring_lo + (uint32_t)q0 * slot_units;
        if constexpr (NCP == 0) {
          // C_in = 8: two taps per K=16 step.  3x3: (kx 0, kx 1) of each row, and kx 2 of rows (i, i+1) for even i (the
          // ring has an even number of slots and windows start on even slots, so the pair never straddles the wrap);
          // 4x4: (kx 0, kx 1) and (kx 2, kx 3) of each row.
          const uint32_t lbo1 = 1u << 16, lbo_row = slot_units << 16;
          uint32_t acc = 0;
          if constexpr (KS == 3) {
#pragma unroll 1
            for (int i = 0; i < rows_win; i += 2) {   // rows_win = R + 2 is even
              const uint32_t b_nxt = b_row + slot_units;
              umma_f16(d_tmem, desc64(hi, a_row), desc64(hi, b_row | lbo1), idesc, acc);
              umma_f16(d_tmem, desc64(hi, a_row + cop), desc64(hi, b_nxt | lbo1), idesc, 1u);
              umma_f16(d_tmem, desc64(hi, a_row + master_units), desc64(hi, (b_row + 2u) | lbo_row), idesc, 1u);
              acc = 1u;
              a_row += 2u * cop;
              b_row = b_nxt + slot_units;
              if (b_row >= ring_end) b_row -= (uint32_t)a.S * slot_units;
            }
          } else {
#pragma unroll 1
            for (int i = 0; i < rows_win; ++i) {
              umma_f16(d_tmem, desc64(hi, a_row), desc64(hi, b_row | lbo1), idesc, acc);
              umma_f16(d_tmem, desc64(hi, a_row + master_units), desc64(hi, (b_row + 2u) | lbo1), idesc, 1u);
              acc = 1u;
              a_row += cop;
              b_row += slot_units;
              if (b_row >= ring_end) b_row -= (uint32_t)a.S * slot_units;
            }
          }
        } else {
          const uint32_t lbo = row_units << 16;
          uint32_t acc = 0;
#pragma unroll 1
          for (int i = 0; i < rows_win; ++i) {
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
#pragma unroll
              for (int cp = 0; cp < NCP; ++cp) {
                umma_f16(d_tmem, desc64(hi, a_row + (uint32_t)(kx * NCP + cp) * master_units),
                         desc64(hi, (b_row + (uint32_t)kx + (uint32_t)(2 * cp) * row_units) | lbo), idesc, acc);
                acc = 1u;
              }
            }
            a_row += cop;
            b_row += slot_units;
            if (b_row >= ring_end) b_row -= (uint32_t)a.S * slot_units;
          }
        }
        umma_commit(s_tfull + 8 * stage);        // accumulators ready for the epilogue
        umma_commit(s_done + 8 * (t & 3));       // the producer may recycle the rows only this tile used
        // next tile: continuing down the strip keeps the last ks-1 rows, otherwise a fresh window
        TileRef nx = tr;
        if (++nx.ty == a.n_ty) { nx.ty = 0; if (++nx.s == a.n_strips) { nx.s = 0; ++nx.n; } }
        q0 += (nx.ty != 0) ? a.R : rows_win;
        if (q0 >= a.S) q0 -= a.S;
        tr = nx;
      }
      if (a.dbg & 16) {
        g_band_stats[blockIdx.x][0] = BAND_CLK() - c_tot;
        g_band_stats[blockIdx.x][1] = c_tempty;
        g_band_stats[blockIdx.x][2] = c_full;
      }
      }
    } else if (warp == 1) {
      // ================= producer (whole warp: lane l issues the copies of new row l of the tile) =================
      // A single thread needs ~400 cycles per row (address arithmetic + copy issue), more than a 16-row tile leaves.
      const uint32_t row_bytes = (uint32_t)a.row_units * 16u;
      const uint32_t tx = row_bytes * (uint32_t)a.cin_chunks;
      int q = 0;                 // rows loaded so far (ring sequence number of the next row)
      int slot = 0;              // q % S
      int tiles_done = 0;        // tiles whose completion has been observed
      int free_upto = 0;         // rows < free_upto are dead: row q may be loaded when q - S < free_upto
      TileRef tr = tile_ref(L0, a.n_strips, a.n_ty);
      long long c_done = 0, c_ptot = BAND_CLK();
      for (int t = 0; t < n_tiles; ++t) {
        const bool fresh = (t == 0) || (tr.ty == 0);
        const int i0 = fresh ? 0 : a.ks - 1;
        const int n_new = rows_win - i0;
        if (lane == 0) qs_ring[t & 7] = q - i0;
        __syncwarp();
        const int y_first = tr.ty * a.R - 1;     // image row of input row 0
        const int x_first = tr.s * a.N - 1;      // image column of slot position 0
        // rows -1 .. H (+1 for 4x4) exist in the padded layout (zero frame rows / the next page's frame); rows below
        // are not loaded: their slots keep stale finite data that only feeds discarded output rows
        const int i_end = min(rows_win, a.H + a.ks - 3 - y_first + 1);
        // all slots of the new rows must be dead: the last new row is q + n_new - 1
        while (q + n_new - 1 - a.S >= free_upto) {
          long long c2 = BAND_CLK();
          mbar_wait_sleep(s_done + 8 * (tiles_done & 3), (uint32_t)(tiles_done >> 2) & 1u, a.err_flag, 1);
          c_done += BAND_CLK() - c2;
          ++tiles_done;
          free_upto = (tiles_done <= t) ? qs_ring[tiles_done & 7] : q;
        }
        const uint32_t bar = s_full + 8 * (t & 3);
        if (lane == 0) mbar_expect_tx(bar, tx * (uint32_t)max(i_end - i0, 0));
        __syncwarp();
        for (int i = i0 + lane; i < i_end; i += 32) {
          int sl = slot + (i - i0);
          if (sl >= a.S) sl -= a.S;
          const long long pos = a.lead + ((long long)tr.n * a.Hp + (y_first + i) + 1) * a.Wp + (x_first + 1);
          const uint32_t dst = s_ring + (uint32_t)sl * (uint32_t)a.slot_units * 16u;
          for (int c = 0; c < a.cin_chunks; ++c)
            bulk_g2s(dst + (uint32_t)c * row_bytes, a.in + ((long long)c * a.in_plane + pos) * 8, row_bytes, bar);
        }
        q += n_new;
        slot += n_new;
        if (slot >= a.S) slot -= a.S;
        if (++tr.ty == a.n_ty) { tr.ty = 0; if (++tr.s == a.n_strips) { tr.s = 0; ++tr.n; } }
      }
      if (lane == 0 && (a.dbg & 16)) {
        g_band_stats[blockIdx.x][4] = BAND_CLK() - c_ptot;
        g_band_stats[blockIdx.x][5] = c_done;
      }
    } else if (warp >= 4) {
      // ================= epilogue: two specialised warpgroups joined by a ring of transposition slabs =================
      const int q4 = warp & 3;
      const int n_slabs = (a.N + SLAB - 1) / SLAB;
      if (warp < 8) {
        // ---- drain warps: TMEM (lane = (rho, co), registers = columns) -> + bias -> 16-bit -> slab[(rho, chunk)][x][co % 8]
        const int m = q4 * 32 + lane;            // accumulator row = rho * cop + co
        const float bias = s_bias[m];
        const int st_off = (m >> 3) * SLAB_ROW + (m & 7) * 2;
        int b = 0;
        uint32_t bpar = 0;
        long long c_tf = 0, c_se = 0;
        for (int t = 0; t < n_tiles; ++t) {
          const int stage = t & 1;
          long long c0 = BAND_CLK();
          mbar_wait_sleep(s_tfull + 8 * stage, (uint32_t)(t >> 1) & 1u, a.err_flag, 4);
          c_tf += BAND_CLK() - c0;
          __syncwarp();
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(stage * 256);
          uint32_t v[32];
          tmem_ld32(taddr, v);
#pragma unroll 1
          for (int sb = 0; sb < n_slabs; ++sb) {
            long long c1 = BAND_CLK();
            mbar_wait_sleep(s_sempty + 8 * b, bpar ^ 1u, a.err_flag, 5);
            c_se += BAND_CLK() - c1;
            uint8_t* my_st = slabs + b * SLAB_BYTES + st_off;
            tmem_ld_wait();
            if constexpr (HEAD) {
              // head: the logits of the (<= 4) real channels stay fp32 through the slab, 16 B per position
              if ((m & 7) < 4) {
                float* fs = reinterpret_cast<float*>(slabs + b * SLAB_BYTES + (m >> 3) * SLAB_ROW + (m & 7) * 4);
#pragma unroll
                for (int k = 0; k < 32; ++k) fs[k * 4] = __uint_as_float(v[k]) + bias;
              }
            } else
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const uint32_t h2 = pack2_raw(__uint_as_float(v[k]) + bias, __uint_as_float(v[k + 1]) + bias);
              *reinterpret_cast<unsigned short*>(my_st + k * 16) = (unsigned short)(h2 & 0xffffu);
              *reinterpret_cast<unsigned short*>(my_st + (k + 1) * 16) = (unsigned short)(h2 >> 16);
            }
            if (sb + 1 < n_slabs) {
              tmem_ld32(taddr + (uint32_t)((sb + 1) * SLAB), v);   // in flight while the slab is handed over
              __syncwarp();
            } else {   // last TMEM read of this stage is complete: hand the accumulators back to the issuer
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(s_tempty + 8 * stage);
            }
            if (lane == 0) mbar_arrive(s_sfull + 8 * b);
            if (++b == N_SLAB_BUF) { b = 0; bpar ^= 1u; }
          }
        }
        if (warp == 4 && lane == 0 && (a.dbg & 16)) { g_band_stats[blockIdx.x][6] = c_tf; g_band_stats[blockIdx.x][7] = c_se; }
      } else {
        // ---- store warps: slab -> position-major 16 B vectors -> (+ residual) -> activation -> 512 B coalesced stores
        const bool has_res = a.res != nullptr, has_pre = a.out_pre != nullptr, relu = a.act == 1;
        const int nc_mask = (1 << a.nc_shift) - 1;
        // Each of the 8 store warps owns two row segments of every slab (lane = column): the two image rows of a 2x2
        // pooling window in one channel chunk, so a fused max-pool never leaves the warp.  rs = rho * nc + chunk.
        const int pp = warp - 8;                       // vertical pair 0..7
        const int cch = pp & nc_mask;                  // channel chunk
        const int rho0 = 2 * (pp >> a.nc_shift);       // rho of k = 0 (odd image row); k = 1 is rho0 + 1 (the even row above)
        const int rs0 = (rho0 << a.nc_shift) + cch, rs1 = rs0 + (1 << a.nc_shift);
        const long long out_c = (long long)cch * a.out_plane * 8, pre_c = (long long)cch * a.pre_plane * 8,
                        res_c = (long long)cch * a.res_plane * 8;
        int b = 0;
        uint32_t bpar = 0;
        TileRef tr = tile_ref(L0, a.n_strips, a.n_ty);
        long long c_sf = 0;
        for (int t = 0; t < n_tiles; ++t) {
          const int y_top = tr.ty * a.R, x_left = tr.s * a.N;
          const int y1 = y_top + (a.R - 2 - rho0);     // even image row (k = 1); k = 0 is y1 + 1
          const bool gv[2] = {y1 + 1 < a.H, y1 < a.H};
          // element offset of (row y1, column x_left); row k = 0 is one padded row further
          const long long gp1 = (a.lead + ((long long)tr.n * a.Hp + y1 + 1) * a.Wp + (x_left + 1)) * 8;
          const long long row8 = (long long)a.Wp * 8;
          long long ppos = 0;   // element offset of the pooled row (column x_left / 2)
          if constexpr (POOL)
            ppos = ((long long)cch * a.pool_plane + a.lead_o + ((long long)tr.n * a.Hpo + (y1 >> 1) + 1) * a.Wpo + (x_left >> 1) + 1) * 8;
          float* fo = nullptr;
          if constexpr (HEAD) fo = a.out_f32 + (((long long)tr.n * a.H + y1) * a.W + x_left) * a.f32_c;
          uint4 rr[2];
          if (has_res) {
            // Residual of the NEXT tile -> L2 now (one 128 B line per lane and row): the register prefetch below only
            // reaches one slab ahead, which covers an L2 hit but not DRAM latency.
            TileRef nx = tr;
            if (++nx.ty == a.n_ty) { nx.ty = 0; if (++nx.s == a.n_strips) { nx.s = 0; ++nx.n; } }
            if (t + 1 < n_tiles) {
              const int cols = min(a.N, a.W - nx.s * a.N);
              const int yn = nx.ty * a.R + (a.R - 2 - rho0);
              if (lane * 8 < cols) {
                const long long pos = a.lead + ((long long)nx.n * a.Hp + yn + 1) * a.Wp + (nx.s * a.N + 1) + lane * 8;
                if (yn < a.H) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + res_c + pos * 8));
                if (yn + 1 < a.H) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + res_c + (pos + a.Wp) * 8));
              }
            }
            const bool xv = (lane < a.N) && (x_left + lane < a.W);
            rr[0] = ldg_nc_v4(a.res + res_c + gp1 + row8 + (long long)lane * 8, xv && gv[0]);
            rr[1] = ldg_nc_v4(a.res + res_c + gp1 + (long long)lane * 8, xv && gv[1]);
          }
#pragma unroll 1
          for (int sb = 0; sb < n_slabs; ++sb) {
            const int xl = sb * SLAB + lane;
            const bool xv = (xl < a.N) && (x_left + xl < a.W);
            uint4 rn[2];
            if (has_res && sb + 1 < n_slabs) {   // residual vectors of the next slab: a slab time of latency hiding
              const int xn = xl + SLAB;
              const bool xnv = (xn < a.N) && (x_left + xn < a.W);
              rn[0] = ldg_nc_v4(a.res + res_c + gp1 + row8 + (long long)xn * 8, xnv && gv[0]);
              rn[1] = ldg_nc_v4(a.res + res_c + gp1 + (long long)xn * 8, xnv && gv[1]);
            }
            long long c2 = BAND_CLK();
            mbar_wait_sleep(s_sfull + 8 * b, bpar, a.err_flag, 6);
            c_sf += BAND_CLK() - c2;
            const uint8_t* slab = slabs + b * SLAB_BYTES;
            uint4 raw[2];
            raw[0] = *reinterpret_cast<const uint4*>(slab + rs0 * SLAB_ROW + lane * 16);
            raw[1] = *reinterpret_cast<const uint4*>(slab + rs1 * SLAB_ROW + lane * 16);
            __syncwarp();
            if (lane == 0) mbar_arrive(s_sempty + 8 * b);   // the slab is in registers
            if (++b == N_SLAB_BUF) { b = 0; bpar ^= 1u; }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if constexpr (HEAD) {
                if (xv && gv[k]) {
                  // classifier / attention-logit head (cop = 8, one chunk): fp32 logits -> activation -> dense NHWC
                  const int C = a.f32_c;
                  float* o = fo + ((long long)(1 - k) * a.W + xl) * C;
                  float z[4] = {__uint_as_float(raw[k].x), __uint_as_float(raw[k].y), __uint_as_float(raw[k].z), __uint_as_float(raw[k].w)};
                  if (a.act == 2 && C == 2) {   // two-class softmax (layers.py:48-49): the shipped separator / heading nets
                    const float mx = fmaxf(z[0], z[1]);
                    const float e0 = __expf(z[0] - mx), e1 = __expf(z[1] - mx);
                    const float inv = __fdividef(1.f, e0 + e1);
                    *reinterpret_cast<float2*>(o) = make_float2(e0 * inv, e1 * inv);
                  } else {
                    if (a.act == 1) {
#pragma unroll
                      for (int j = 0; j < 4; ++j) z[j] = fmaxf(z[j], 0.f);
                    } else if (a.act == 2) {
                      float mx = z[0];
                      for (int j = 1; j < C; ++j) mx = fmaxf(mx, z[j]);
                      float sum = 0.f;
                      for (int j = 0; j < C; ++j) { z[j] = __expf(z[j] - mx); sum += z[j]; }
                      const float inv = __fdividef(1.f, sum);
                      for (int j = 0; j < C; ++j) z[j] *= inv;
                    } else if (a.act == 3) {
                      for (int j = 0; j < C; ++j) z[j] = __fdividef(1.f, 1.f + __expf(-z[j]));
                    }
                    for (int j = 0; j < C; ++j) o[j] = z[j];
                  }
                }
              } else {
                const long long p8 = gp1 + (k == 0 ? row8 : 0) + (long long)xl * 8;
                uint4 val = raw[k];
                if (has_res) val = add8(val, rr[k]);
                const uint4 fin = relu ? clamp8<true>(val) : clamp8<false>(val);
                if (xv && gv[k]) {
                  if (has_pre) *reinterpret_cast<uint4*>(a.out_pre + pre_c + p8) = clamp8<false>(val);
                  *reinterpret_cast<uint4*>(a.out + out_c + p8) = fin;
                }
                if constexpr (POOL) raw[k] = (xv && gv[k]) ? fin : make_uint4(0u, 0u, 0u, 0u);   // ReLU'd: 0 is neutral
              }
            }
            if constexpr (POOL) {
              // fused 2x2 stride-2 SAME max-pool of the ReLU'd output (layers.py:526-534): the two rows are a vertical
              // pair, lanes (2j, 2j+1) a horizontal one; cells outside the page count as 0, which max() ignores.
              uint4 mx = max8(raw[0], raw[1]);
              uint4 ot;
              ot.x = __shfl_xor_sync(0xffffffffu, mx.x, 1);
              ot.y = __shfl_xor_sync(0xffffffffu, mx.y, 1);
              ot.z = __shfl_xor_sync(0xffffffffu, mx.z, 1);
              ot.w = __shfl_xor_sync(0xffffffffu, mx.w, 1);
              mx = max8(mx, ot);
              if (!(lane & 1) && xv && gv[1])   // even column, and the even row of the pair is in the page
                *reinterpret_cast<uint4*>(a.pool_out + ppos + (long long)(xl >> 1) * 8) = mx;
            }
            if (has_res) {
              rr[0] = rn[0];
              rr[1] = rn[1];
            }
          }
          if (++tr.ty == a.n_ty) { tr.ty = 0; if (++tr.s == a.n_strips) { tr.s = 0; ++tr.n; } }
        }
        if (warp == 8 && lane == 0 && (a.dbg & 16)) g_band_stats[blockIdx.x][3] = c_sf;
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

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------------
ConvBandPlan conv_band_plan(int ks, int cin, int cout, const Geo& g, int num_sms, size_t max_smem) {
  ConvBandPlan p;
  p.ks = ks;
  p.cin_chunks = cdiv(cin, 8);
  if (ks != 3 && ks != 4) { p.why = "kernel size"; return p; }
  int cop = 8;
  while (cop < cout) cop <<= 1;
  if (cop > 64 || cin > 64) { p.why = "C > 64: the position-major kernel is at its floor there"; return p; }
  if (cdiv(cout, 8) * 8 != cop) { p.why = "C_out planes are not a power of two"; return p; }
  if (p.cin_chunks != 1 && (p.cin_chunks & 1)) { p.why = "odd number of input chunks"; return p; }
  p.cop = cop;
  p.R = 128 / cop;
  p.J = 2 * p.R + ks - 2;
  const int rows_win = p.R + ks - 1;
  const int n_cp = p.cin_chunks == 1 ? 1 : p.cin_chunks / 2;
  p.n_masters = p.cin_chunks == 1 ? 2 : ks * n_cp;
  p.n_steps = p.cin_chunks == 1 ? (ks == 3 ? rows_win + rows_win / 2 : rows_win * 2) : rows_win * ks * n_cp;
  p.wpack_bytes = (size_t)p.n_masters * 2 * p.J * cop * 16;
  if (p.wpack_bytes / 16 > 65535) { p.why = "masters too large"; return p; }
  const size_t fixed = HDR_BYTES + ((p.wpack_bytes + 127) / 128) * 128 + 4 * SLAB_BYTES + 128;
  // strips: as few as possible, then the narrowest N (multiple of 16) that covers the page width
  p.why = "row ring does not fit in shared memory";
  for (int n_strips = cdiv(g.W, 256); n_strips <= cdiv(g.W, 16) && !p.ok; ++n_strips) {
    const int N = std::max(16, cdiv(cdiv(g.W, n_strips), 16) * 16);
    if (N > 256) continue;
    const size_t slot = (size_t)p.cin_chunks * (N + ks - 1) * 16;
    if (slot / 16 > 65535) continue;
    int S = (int)std::min<size_t>((max_smem - std::min(max_smem, fixed)) / slot, MAX_SLOTS);
    S = std::min(S, rows_win + 2 * p.R);   // more than two tiles of prefetch buys nothing
    S &= ~1;                               // vertical tap pairs must not straddle the ring wrap
    // the producer loads a tile's new rows at once: the ring must hold the window plus the next tile's rows, else the
    // loads of tile t+1 cannot start before tile t has retired (try narrower strips)
    if (S < rows_win + p.R) {
      if (N == 16) break;
      continue;
    }
    p.ok = true;
    p.N = N;
    p.n_strips = n_strips;
    p.S = S;
    p.smem_bytes = fixed + (size_t)S * slot;
  }
  if (!p.ok) return p;
  p.why = "";
  p.n_ty = cdiv(g.H, p.R);
  p.tiles = (long long)g.N * p.n_strips * p.n_ty;
  p.grid = (int)std::min<long long>(p.tiles, num_sms);
  return p;
}

// Banded masters.  Master b, K half h, row (j, co), element e  ->  W[ky][kx][ci][co] of the TF filter [ks][ks][cin][cout]:
//   C_in = 8, 3x3 : b = 0  horizontal pair  ky = j-(R-1), kx = h, ci = e       b = 1  vertical pair  ky = j+h-(R-1), kx = 2
//   C_in = 8, 4x4 : b = 0  kx = h                                              b = 1  kx = 2 + h
//   C_in >= 16    : b = kx * n_cp + cp,  ky = j-(R-1),  ci = (2 cp + h) * 8 + e
void conv_band_pack(const ConvBandPlan& p, const float* w, int cin, int cout, uint16_t* dst) {
  memset(dst, 0, p.wpack_bytes);
  const int ks = p.ks, R = p.R, J = p.J, cop = p.cop;
  const int n_cp = p.cin_chunks == 1 ? 1 : p.cin_chunks / 2;
  auto W = [&](int ky, int kx, int ci, int co) -> float {
    if (ky < 0 || ky >= ks || kx < 0 || kx >= ks || ci >= cin || co >= cout) return 0.f;
    return w[(((size_t)ky * ks + kx) * cin + ci) * cout + co];
  };
  for (int b = 0; b < p.n_masters; ++b)
    for (int h = 0; h < 2; ++h)
      for (int j = 0; j < J; ++j)
        for (int co = 0; co < cop; ++co)
          for (int e = 0; e < 8; ++e) {
            float v;
            if (p.cin_chunks == 1) {
              if (ks == 3) v = b == 0 ? W(j - (R - 1), h, e, co) : W(j + h - (R - 1), 2, e, co);
              else v = W(j - (R - 1), 2 * b + h, e, co);
            } else {
              const int kx = b / n_cp, cp = b % n_cp;
              v = W(j - (R - 1), kx, (2 * cp + h) * 8 + e, co);
            }
            dst[((((size_t)b * 2 + h) * J + j) * cop + co) * 8 + e] = host_f_to_act(v);
          }
}

cudaError_t launch_conv_band(cudaStream_t st, const ConvBandPlan& p, PV in, PV out, PV out_pre, PV res, const act_t* wpack,
                             const float* bias_pad, const Geo& g, int act, int* err_flag, float* out_f32, int f32_c,
                             PV pool, const Geo* pool_geo) {
  if (!p.ok) return cudaErrorInvalidValue;
  if (out_f32 && (p.cop != 8 || f32_c < 1 || f32_c > 4)) return cudaErrorInvalidValue;
  ConvBandArgs a{};
  a.out_f32 = out_f32; a.f32_c = f32_c;
  if (pool.p) {
    // the fused pool relies on max(x, 0) = x (ReLU'd values, zero cells outside the page) and on even tile origins
    if (!pool_geo || out_f32 || act != 1 || p.R < 2 || pool_geo->H != (g.H + 1) / 2 || pool_geo->W != (g.W + 1) / 2)
      return cudaErrorInvalidValue;
    a.pool_out = pool.p; a.pool_plane = pool.plane; a.lead_o = pool_geo->lead;
    a.Wpo = pool_geo->Wp; a.Hpo = pool_geo->Hp;
  }
  a.in = in.p; a.in_plane = in.plane;
  a.out = out.p; a.out_plane = out.plane;
  a.out_pre = out_pre.p; a.pre_plane = out_pre.plane;
  a.res = res.p; a.res_plane = res.plane;
  a.wpack = wpack; a.bias = bias_pad;
  a.ks = p.ks; a.cin_chunks = p.cin_chunks; a.cop = p.cop; a.R = p.R; a.N = p.N;
  a.nc_shift = 0;
  while ((8 << a.nc_shift) < p.cop) ++a.nc_shift;
  a.n_strips = p.n_strips; a.n_ty = p.n_ty; a.S = p.S;
  a.n_steps = p.n_steps; a.J = p.J; a.masters_bytes = (int)p.wpack_bytes;
  a.row_units = p.N + p.ks - 1;
  a.slot_units = p.cin_chunks * a.row_units;
  a.W = g.W; a.H = g.H; a.Wp = g.Wp; a.Hp = g.Hp; a.NP = g.N;
  a.lead = g.lead;
  a.tiles = p.tiles;
  a.act = act;
  a.err_flag = err_flag;
  { const char* e = getenv("ARU_BAND_DBG"); a.dbg = e ? atoi(e) : 0; }
  using Kern = void (*)(const __grid_constant__ ConvBandArgs);
  const int ncp = p.cin_chunks == 1 ? 0 : p.cin_chunks / 2;
  Kern k = nullptr;
  int ki = -1;
  const bool head = out_f32 != nullptr, pl = pool.p != nullptr;
#define ARU_BAND_PICK(KK, NN, HH, PP, II) \
  if (p.ks == KK && ncp == NN && head == HH && pl == PP) { k = k_conv_band<KK, NN, HH, PP>; ki = II; }
  ARU_BAND_PICK(3, 0, false, false, 0) ARU_BAND_PICK(3, 1, false, false, 1) ARU_BAND_PICK(3, 2, false, false, 2)
  ARU_BAND_PICK(3, 4, false, false, 3) ARU_BAND_PICK(4, 0, false, false, 4) ARU_BAND_PICK(4, 1, false, false, 5)
  ARU_BAND_PICK(4, 2, false, false, 6) ARU_BAND_PICK(4, 4, false, false, 7)
  ARU_BAND_PICK(3, 0, true, false, 8) ARU_BAND_PICK(4, 0, true, false, 9)
  ARU_BAND_PICK(3, 0, false, true, 10) ARU_BAND_PICK(3, 1, false, true, 11) ARU_BAND_PICK(3, 2, false, true, 12)
  ARU_BAND_PICK(4, 0, false, true, 13) ARU_BAND_PICK(4, 1, false, true, 14) ARU_BAND_PICK(4, 2, false, true, 15)
#undef ARU_BAND_PICK
  if (!k) return cudaErrorInvalidValue;
  static bool configured[16] = {};
  if (!configured[ki]) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured[ki] = true;
  }
  k<<<p.grid, NUM_THREADS, p.smem_bytes, st>>>(a);
  if (a.dbg & 16) {
    cudaDeviceSynchronize();
    static unsigned long long h[160][8];
    cudaMemcpyFromSymbol(h, g_band_stats, sizeof(h));
    double s[8] = {0};
    for (int i = 0; i < p.grid; ++i) for (int j = 0; j < 8; ++j) s[j] += (double)h[i][j] / p.grid;
    fprintf(stderr, "band %dx%d cin%d cop%d N%d S%d tiles/cta %.1f res%d pre%d | issuer tot %.0f tempty %.0f full %.0f | producer tot %.0f done %.0f | drain tfull %.0f sempty %.0f | store sfull %.0f\n",
            g.H, g.W, p.cin_chunks * 8, p.cop, p.N, p.S, (double)p.tiles / p.grid, a.res != nullptr, a.out_pre != nullptr, s[0], s[1], s[2], s[4], s[5], s[6], s[7], s[3]);
  }
  return cudaGetLastError();
}

}  // namespace aru
