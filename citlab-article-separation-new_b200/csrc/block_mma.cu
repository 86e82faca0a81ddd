// block_mma.cu - a whole residual block of the 8 / 16 channel levels in ONE launch (sm_100a, warp-level tensor cores):
//     [conv1 ->] x0 -> ReLU -> convR_0 -> ReLU -> convR_1 -> ReLU -> convR_2 + x0 -> ReLU [-> 2x2 max-pool]
// (ARU_v1.py:212-227 down path, :266-281 up path).  The three (four) intermediates never leave the SM.
//
// Why not tcgen05 here.  With C = 8 / 16 channels a tcgen05.mma is bound by its shared-memory operand reads: both
// formulations that were built (positions on M: 4 KB of A per 32 KFLOP; banded weights on M: conv_band.cu) make one
// layer cost about its HBM time, and chaining two of them through shared memory made the shared-memory port the limit
// (conv_band2.cu, DESIGN.md 4.4: slower than two launches).  The warp-level path keeps its operands in REGISTERS:
// an ldmatrix fragment of an input row is used by the three output rows it contributes to, the filters live in
// registers for the whole launch, and `mma.sync.m16n8k16` sustains 2 clk per warp-MMA per SM whether or not an ldmatrix
// accompanies it (tools/hmma_rate_bench.cu, profiles/r02l_hmma_rate.txt: 2044 FLOP/clk/SM = 557 TFLOP/s).  A 3x3 8 -> 8
// layer then costs 9 clk per 16 pixels and SM = 0.16 ms for 32 pages of 1856x1344 against 0.47 ms of HBM time for the
// same layer run alone - and the fused block moves 36 - 48 B per pixel instead of 150 - 180.
//
// Geometry.  A CTA owns a strip of WS output columns (WS even) and streams down a segment of rows of one page in bands of
// RB = 4 rows.  All stages use the same WC = 16 * (compute warps) columns: compute column u <-> image column xs - 4 + u;
// stage k's values are valid for u in [k, WC - k), the block's output for u in [4, 4 + WS)  (WS <= WC - 8; the even
// offset keeps the 2x2 pooling windows inside a lane pair).  Warp w owns columns [16 w, 16 w + 16) of every stage.
//
// Pipeline.  Stage s (0 = conv1, 1..3 = convR_0..2) computes, in iteration t, the rows [a_s(t), a_s(t) + RB) with
// a_s(t) = a_3(t) + (3 - s)(RB + 1): every stage reads what the stage before it wrote in EARLIER iterations, so one
// __syncthreads per iteration is the only synchronisation, and the four stages of an iteration are independent work.
// Rows live in shared-memory rings (input 2RB+2 rows, x0 3(RB+1)+RB+1 rows - convR_2 adds it back 3(RB+1) rows later -,
// r0 / r1 2RB+2 rows), position-major like HBM (16 B per pixel and 8 channels), out-of-image positions stored as zeros
// (SAME padding of the next layer sees zeros, not conv values).  A loader lane streams the next band of input rows
// with bulk copies (TMA engine) one iteration ahead.
//
// MMA mapping (m16n8k16, A row-major = 16 positions x 16 k, B = filters): C_in = 8: the taps kx = -1, 0 of an input row
// form one K = 16 step (two ldmatrix halves at a one-position shift), kx = +1 is a K = 8 step; C_in >= 16: one tap and two
// input planes per K = 16 step.  An input row j feeds the output rows j-2 .. j of the band (ky = 2, 1, 0), so every
// fragment is loaded once per band and stage and used up to three times.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

#include "kernels.h"
#include "band_common.cuh"

namespace aru {
namespace {

constexpr int RB = 4;             // rows per band
constexpr int LAG = RB + 1;       // rows a stage runs behind its predecessor
constexpr int IN_ROWS = 2 * RB + 2;
constexpr int X0_ROWS = 3 * LAG + RB + 1;
constexpr int R_ROWS = 2 * RB + 2;
constexpr int ROW_BIAS = 64 * IN_ROWS * X0_ROWS;   // multiple of every ring size, makes row + ROW_BIAS non-negative

struct BlockArgs {
  const act_t* in;       // CIN_CH > 0: conv1's input planes; CIN_CH == 0: the block's x0 (pre-activation of conv1)
  long long in_plane;    // positions per plane
  act_t* out;
  long long out_plane;
  act_t* pool;           // POOL: 2x2 stride-2 SAME max-pool of the output
  long long pool_plane;
  Geo pg;
  const uint32_t* wfrag;   // per-lane B fragments of the four stages (block_mma_pack)
  const float* bias;       // [4][16]
  Geo g;
  int WC, WS, pitch;       // compute columns, output columns per strip, ring row pitch in positions (WC + 2)
  int n_strips, rows_per_unit, units_per_page, n_units;
  int wbase[4];            // first fragment register of every stage in wfrag
  int* err_flag;
};

#ifdef ARU_USE_BF16
#define ARU_MMA_T "bf16"
#else
#define ARU_MMA_T "f16"
#endif

__device__ __forceinline__ void mma16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                      uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." ARU_MMA_T "." ARU_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32." ARU_MMA_T "." ARU_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldsm4(uint32_t (&a)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm2(uint32_t (&a)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(a[0]), "=r"(a[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t relu2(uint32_t v) {
  act2_t h = *reinterpret_cast<act2_t*>(&v);
#ifdef ARU_USE_BF16
  h = __hmax2(h, __floats2bfloat162_rn(0.f, 0.f));
#else
  h = __hmax2(h, __half2half2(__ushort_as_half((unsigned short)0)));
#endif
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t max2(uint32_t x, uint32_t y) {
  act2_t h = __hmax2(*reinterpret_cast<act2_t*>(&x), *reinterpret_cast<act2_t*>(&y));
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int ring_slot(int row, int ring) { return (row + ROW_BIAS) % ring; }

// number of B-fragment registers of a stage with ICH input planes and NT output planes
__host__ __device__ constexpr int stage_regs(int ich, int nt) { return ich == 1 ? 9 * nt : 9 * ich * nt; }

template <int N>
__device__ __forceinline__ void load_wfrag(uint32_t (&w)[N], const uint32_t* __restrict__ img, int base, int lane) {
#pragma unroll
  for (int r = 0; r < N; ++r) w[r] = __ldg(img + (size_t)(base + r) * 32 + lane);
}

// One stage of one band: acc[i][nt] += conv over the RB + 2 input rows starting at page row `row_first`.
//   ring: shared-memory address of the input ring (ring_rows rows of ICH planes of `pitch` positions)
//   lane4 / lane2: per-lane ldmatrix row offsets (bytes) inside a ring row
template <int ICH, int NT, bool RELU_IN>
__device__ __forceinline__ void stage_mma(float (&acc)[RB][NT][4], uint32_t ring, int ring_rows, int row_first, int pitch,
                                          uint32_t lane4, uint32_t lane2, const uint32_t (&w)[stage_regs(ICH, NT)]) {
  int slot = ring_slot(row_first, ring_rows);
  const uint32_t row_bytes = (uint32_t)(ICH * pitch) * 16u;
#pragma unroll
  for (int j = 0; j < RB + 2; ++j) {
    const uint32_t base = ring + (uint32_t)slot * row_bytes;
    slot = slot + 1 == ring_rows ? 0 : slot + 1;
    if (ICH == 1) {
      uint32_t a[4], c[2];
      ldsm4(a, base + lane4);   // kx = -1 (k 0..7) | kx = 0 (k 8..15)
      ldsm2(c, base + lane2);   // kx = +1
      if (RELU_IN) {
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = relu2(a[q]);
        c[0] = relu2(c[0]);
        c[1] = relu2(c[1]);
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int ky = j - i;   // input row j = output row i + ky - 1 + 1
        if (ky < 0 || ky > 2) continue;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int r = (ky * NT + nt) * 3;
          mma16(acc[i][nt], a[0], a[1], a[2], a[3], w[r], w[r + 1]);
          mma8(acc[i][nt], c[0], c[1], w[r + 2]);
        }
      }
    } else {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
        for (int cp = 0; cp < ICH / 2; ++cp) {
          uint32_t a[4];
          ldsm4(a, base + lane4 + (uint32_t)(cp * 2 * pitch + kx) * 16u);   // plane 2cp (k 0..7) | plane 2cp+1 (k 8..15)
          if (RELU_IN) {
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] = relu2(a[q]);
          }
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int ky = j - i;
            if (ky < 0 || ky > 2) continue;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const int r = (((ky * 3 + kx) * (ICH / 2) + cp) * NT + nt) * 2;
              mma16(acc[i][nt], a[0], a[1], a[2], a[3], w[r], w[r + 1]);
            }
          }
        }
      }
    }
  }
}

template <int NT>
__device__ __forceinline__ void zero_acc(float (&acc)[RB][NT][4]) {
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[i][nt][0] = acc[i][nt][1] = acc[i][nt][2] = acc[i][nt][3] = 0.f;
}

// bias + (ReLU) + 16-bit store of a band into a ring; out-of-image positions become zeros
template <int NT, bool RELU>
__device__ __forceinline__ void store_ring(const float (&acc)[RB][NT][4], uint32_t ring, int ring_rows, int row0, int H,
                                           int pitch, uint32_t lane_st, bool cv0, bool cv1, const float (&b)[NT][2]) {
  int slot = ring_slot(row0, ring_rows);
#pragma unroll
  for (int i = 0; i < RB; ++i) {
    const bool rv = (unsigned)(row0 + i) < (unsigned)H;
    const uint32_t base = ring + (uint32_t)(slot * NT * pitch) * 16u + lane_st;
    slot = slot + 1 == ring_rows ? 0 : slot + 1;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t h0 = pack2_fin<RELU>(acc[i][nt][0] + b[nt][0], acc[i][nt][1] + b[nt][1]);
      uint32_t h1 = pack2_fin<RELU>(acc[i][nt][2] + b[nt][0], acc[i][nt][3] + b[nt][1]);
      if (!(rv && cv0)) h0 = 0u;
      if (!(rv && cv1)) h1 = 0u;
      sts32(base + (uint32_t)(nt * pitch) * 16u, h0);
      sts32(base + (uint32_t)(nt * pitch + 8) * 16u, h1);
    }
  }
}

// CIN_CH: input planes of conv1 (0: the launch starts from x0), NT: planes of the block (C / 8)
template <int CIN_CH, int NT, bool POOL>
__global__ void __launch_bounds__(384, 1) k_block_mma(const __grid_constant__ BlockArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_cw = a.WC >> 4;               // compute warps; warp n_cw is the loader
  const int pitch = a.pitch;
  const int H = a.g.H, W = a.g.W;
  constexpr bool CONV1 = CIN_CH > 0;
  constexpr int FIRST = CONV1 ? 0 : 1;
  constexpr int T3 = CONV1 ? (3 * LAG + 3 + RB - 1) / RB : (2 * LAG + 2 + RB - 1) / RB;
  constexpr int LD_CH = CONV1 ? CIN_CH : NT;       // planes the loader streams
  constexpr int LD_ROWS = CONV1 ? IN_ROWS : X0_ROWS;

  // shared-memory carve-up: [mbarriers 64 B][input ring (CONV1)][x0 ring][r0 ring][r1 ring]
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_bar = s_base;
  const uint32_t s_in = s_base + 128;
  const uint32_t s_x0 = s_in + (CONV1 ? (uint32_t)(IN_ROWS * CIN_CH * pitch) * 16u : 0u);
  const uint32_t s_r0 = s_x0 + (uint32_t)(X0_ROWS * NT * pitch) * 16u;
  const uint32_t s_r1 = s_r0 + (uint32_t)(R_ROWS * NT * pitch) * 16u;
  const uint32_t s_end = s_r1 + (uint32_t)(R_ROWS * NT * pitch) * 16u;
  const uint32_t s_ld = CONV1 ? s_in : s_x0;

  for (uint32_t o = 128 + threadIdx.x * 16; o < s_end - s_base; o += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + o) = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    mbar_init(s_bar, 1);
    mbar_init(s_bar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zero fill precedes bulk copies into the rings
  __syncthreads();

  // ---- per-lane constants ------------------------------------------------------------------------------------------
  const int g8 = lane >> 2, t4 = lane & 3;
  const int u0 = warp << 4;
  const int m = lane >> 3, r8 = lane & 7;
  // ldmatrix row offsets inside a ring row (position u lives at (1 + u) * 16 bytes of its plane)
  const uint32_t lane4_1 = (uint32_t)(1 + u0 + (m < 2 ? -1 : 0) + (m & 1) * 8 + r8) * 16u;              // ICH == 1
  const uint32_t lane2_1 = (uint32_t)(1 + u0 + 1 + (m & 1) * 8 + r8) * 16u;
  const uint32_t lane4_c = (uint32_t)((m >> 1) * pitch + 1 + u0 - 1 + (m & 1) * 8 + r8) * 16u;          // ICH >= 2, kx = 0
  const uint32_t lane_st = (uint32_t)(1 + u0 + g8) * 16u + (uint32_t)t4 * 4u;

  // filters and biases: registers for the whole launch
  uint32_t w0[stage_regs(CONV1 ? CIN_CH : 1, NT)];
  uint32_t w1[stage_regs(NT, NT)], w2[stage_regs(NT, NT)], w3[stage_regs(NT, NT)];
  float bs[4][NT][2];
  if (warp < n_cw) {
    if constexpr (CONV1) load_wfrag(w0, a.wfrag, a.wbase[0], lane);
    load_wfrag(w1, a.wfrag, a.wbase[1], lane);
    load_wfrag(w2, a.wfrag, a.wbase[2], lane);
    load_wfrag(w3, a.wfrag, a.wbase[3], lane);
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        bs[s][nt][0] = __ldg(a.bias + s * 16 + nt * 8 + 2 * t4);
        bs[s][nt][1] = __ldg(a.bias + s * 16 + nt * 8 + 2 * t4 + 1);
      }
  }

  uint32_t it = 0;   // iterations since the start of the launch (mbarrier it & 1, parity (it >> 1) & 1)
  for (int unit = blockIdx.x; unit < a.n_units; unit += gridDim.x) {
    const int strip = unit % a.n_strips;
    const int seg = (unit / a.n_strips) % a.units_per_page;
    const int n = unit / (a.n_strips * a.units_per_page);
    const int y0 = seg * a.rows_per_unit;
    const int y1 = min(H, y0 + a.rows_per_unit);
    const int xs = strip * a.WS;
    const int n_iter = (y1 - y0 + RB - 1) / RB + T3;
    const int a3_0 = y0 - T3 * RB;
    // image columns of this lane's two positions
    const int xc0 = xs - 4 + u0 + g8, xc1 = xc0 + 8;
    const bool cv0 = (unsigned)xc0 < (unsigned)W, cv1 = (unsigned)xc1 < (unsigned)W;
    const int x_own_end = min(xs + a.WS, W);
    const bool own0 = xc0 >= xs && xc0 < x_own_end, own1 = xc1 >= xs && xc1 < x_own_end;
    const long long page_row0 = a.g.lead + ((long long)n * a.g.Hp + 1) * a.g.Wp + 1;   // position of (y = 0, x = 0)

    auto stage_active = [&](int s, int t) {
      const int as = a3_0 + t * RB + (3 - s) * LAG;
      return as < y1 + (3 - s) && as + RB > y0 - (3 - s);
    };
    // loader: the rows the first stage reads in iteration t that no earlier iteration brought in
    auto issue_loads = [&](int t, uint32_t it_t) {
      const uint32_t bar = s_bar + (it_t & 1) * 8;
      const int af = a3_0 + t * RB + (3 - FIRST) * LAG;
      int r_lo = t == 0 ? af - 1 : af + 1, r_hi = af + RB;   // inclusive
      if (!stage_active(FIRST, t)) r_hi = r_lo - 1;
      r_lo = max(r_lo, -1);
      r_hi = min(r_hi, H);
      const int rows = max(r_hi - r_lo + 1, 0);
      const uint32_t row_bytes = (uint32_t)pitch * 16u;
      mbar_expect_tx(bar, (uint32_t)(rows * LD_CH) * row_bytes);
      for (int r = r_lo; r <= r_hi; ++r) {
        const int slot = ring_slot(r, LD_ROWS);
        const long long p = page_row0 + (long long)r * a.g.Wp + (xs - 5);
#pragma unroll
        for (int c = 0; c < LD_CH; ++c)
          bulk_g2s(s_ld + (uint32_t)((slot * LD_CH + c) * pitch) * 16u, a.in + ((long long)c * a.in_plane + p) * 8, row_bytes, bar);
      }
    };

    if (warp == n_cw && lane == 0) issue_loads(0, it);
    for (int t = 0; t < n_iter; ++t, ++it) {
      if (warp == n_cw) {
        if (lane == 0 && t + 1 < n_iter) issue_loads(t + 1, it + 1);
      } else {
        mbar_wait(s_bar + (it & 1) * 8, (it >> 1) & 1, a.err_flag, 41);
        const int a3 = a3_0 + t * RB;
        float acc[RB][NT][4];
        if constexpr (CONV1) if (stage_active(0, t)) {
          const int as = a3 + 3 * LAG;
          zero_acc<NT>(acc);
          stage_mma<CONV1 ? CIN_CH : 1, NT, false>(acc, s_in, IN_ROWS, as - 1, pitch, CIN_CH == 1 ? lane4_1 : lane4_c, lane2_1, w0);
          store_ring<NT, false>(acc, s_x0, X0_ROWS, as, H, pitch, lane_st, cv0, cv1, bs[0]);
        }
        if (stage_active(1, t)) {
          const int as = a3 + 2 * LAG;
          zero_acc<NT>(acc);
          stage_mma<NT, NT, true>(acc, s_x0, X0_ROWS, as - 1, pitch, NT == 1 ? lane4_1 : lane4_c, lane2_1, w1);
          store_ring<NT, true>(acc, s_r0, R_ROWS, as, H, pitch, lane_st, cv0, cv1, bs[1]);
        }
        if (stage_active(2, t)) {
          const int as = a3 + LAG;
          zero_acc<NT>(acc);
          stage_mma<NT, NT, false>(acc, s_r0, R_ROWS, as - 1, pitch, NT == 1 ? lane4_1 : lane4_c, lane2_1, w2);
          store_ring<NT, true>(acc, s_r1, R_ROWS, as, H, pitch, lane_st, cv0, cv1, bs[2]);
        }
        if (stage_active(3, t)) {
          zero_acc<NT>(acc);
          stage_mma<NT, NT, false>(acc, s_r1, R_ROWS, a3 - 1, pitch, NT == 1 ? lane4_1 : lane4_c, lane2_1, w3);
          // + bias + x0 (pre-activation of conv1) -> ReLU -> global (+ pooled copy)
          uint32_t hp[RB][NT][2];
          int slot = ring_slot(a3, X0_ROWS);
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int y = a3 + i;
            const bool rv = (unsigned)y < (unsigned)H && y >= y0 && y < y1;
            const uint32_t xb = s_x0 + (uint32_t)(slot * NT * pitch) * 16u + lane_st;
            slot = slot + 1 == X0_ROWS ? 0 : slot + 1;
            const long long prow = page_row0 + (long long)y * a.g.Wp;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const uint32_t q0 = lds32(xb + (uint32_t)(nt * pitch) * 16u);
              const uint32_t q1 = lds32(xb + (uint32_t)(nt * pitch + 8) * 16u);
              const float2 f0 = act2_to_f2(*reinterpret_cast<const act2_t*>(&q0));
              const float2 f1 = act2_to_f2(*reinterpret_cast<const act2_t*>(&q1));
              uint32_t h0 = pack2_fin<true>(acc[i][nt][0] + bs[3][nt][0] + f0.x, acc[i][nt][1] + bs[3][nt][1] + f0.y);
              uint32_t h1 = pack2_fin<true>(acc[i][nt][2] + bs[3][nt][0] + f1.x, acc[i][nt][3] + bs[3][nt][1] + f1.y);
              if (!(rv && cv0)) h0 = 0u;
              if (!(rv && cv1)) h1 = 0u;
              hp[i][nt][0] = h0;
              hp[i][nt][1] = h1;
              act_t* o = a.out + ((long long)nt * a.out_plane + prow) * 8 + 2 * t4;
              if (rv && own0) *reinterpret_cast<uint32_t*>(o + (long long)xc0 * 8) = h0;
              if (rv && own1) *reinterpret_cast<uint32_t*>(o + (long long)xc1 * 8) = h1;
            }
          }
          if (POOL) {
            // a3 is a multiple of RB (even): the band holds whole pooling windows; column pairs are lane pairs (g8, g8 ^ 1)
#pragma unroll
            for (int i = 0; i < RB; i += 2) {
              const int y = a3 + i;
              const bool rv = (unsigned)y < (unsigned)H && y >= y0 && y < y1;
              const long long prow = a.pg.lead + ((long long)n * a.pg.Hp + (y >> 1) + 1) * a.pg.Wp + 1;
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) {
                uint32_t m0 = max2(hp[i][nt][0], hp[i + 1][nt][0]);
                uint32_t m1 = max2(hp[i][nt][1], hp[i + 1][nt][1]);
                m0 = max2(m0, __shfl_xor_sync(0xffffffffu, m0, 4));
                m1 = max2(m1, __shfl_xor_sync(0xffffffffu, m1, 4));
                act_t* o = a.pool + ((long long)nt * a.pool_plane + prow) * 8 + 2 * t4;
                if (!(g8 & 1) && rv && own0) *reinterpret_cast<uint32_t*>(o + (long long)(xc0 >> 1) * 8) = m0;
                if (!(g8 & 1) && rv && own1) *reinterpret_cast<uint32_t*>(o + (long long)(xc1 >> 1) * 8) = m1;
              }
            }
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
size_t block_smem_bytes(int cin_ch, int nt, int wc) {
  const size_t pitch = wc + 2;
  size_t rows = (size_t)(cin_ch > 0 ? IN_ROWS * cin_ch : 0) + (size_t)X0_ROWS * nt + 2 * (size_t)R_ROWS * nt;
  return 128 + rows * pitch * 16;
}

}  // namespace

BlockMmaPlan block_mma_plan(int cin, int c, const Geo& g, int num_sms, size_t max_smem) {
  BlockMmaPlan p;
  if (c != 8) { p.why = "block channels"; return p; }   // C = 16: see DESIGN 4.6
  if (cin != 0 && cin != 8 && cin != 16) { p.why = "conv1 input channels"; return p; }
  p.c = c;
  p.nt = c / 8;
  p.cin_ch = cin / 8;
  // strip width: fewest compute columns in total; at most 11 compute warps (register budget of 384 threads)
  long long best = -1;
  for (int wc = 64; wc <= 176; wc += 16) {
    if (block_smem_bytes(p.cin_ch, p.nt, wc) > max_smem) continue;
    const int ns = cdiv(g.W, wc - 8);
    const long long cost = (long long)ns * wc;
    if (best < 0 || cost <= best) { best = cost; p.n_strips = ns; }
  }
  if (best < 0) { p.why = "shared memory"; return p; }
  p.WS = (cdiv(g.W, p.n_strips) + 1) & ~1;
  p.WC = cdiv(p.WS + 8, 16) * 16;
  p.smem_bytes = block_smem_bytes(p.cin_ch, p.nt, p.WC);
  p.threads = (p.WC / 16 + 1) * 32;
  // row segments: the makespan of ceil(units / CTAs) rounds of (rows + pipeline fill) rows
  long long best_cost = -1;
  for (int upp = 1; upp <= 64; ++upp) {
    const int rpu = cdiv(cdiv(g.H, upp), RB) * RB;
    const int eff = cdiv(g.H, rpu);
    const long long units = (long long)g.N * p.n_strips * eff;
    const long long rounds = (units + num_sms - 1) / num_sms;
    const long long cost = rounds * (rpu + 3 * RB);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      p.rows_per_unit = rpu;
      p.units_per_page = eff;
      p.n_units = (int)units;
    }
  }
  p.grid = std::min(p.n_units, num_sms);
  int base = 0;
  for (int s = 0; s < 4; ++s) {
    p.wbase[s] = base;
    const int ich = s == 0 ? p.cin_ch : p.nt;
    if (ich > 0) base += stage_regs(ich, p.nt);
  }
  p.wfrag_bytes = (size_t)std::max(base, 1) * 32 * 4;
  p.ok = true;
  return p;
}

// B fragments: register (ky, kx, plane, nt) of lane (g8, t4) = { W[ky][kx][plane*8 + 2 t4][nt*8 + g8], W[..][.. + 1][..] }
void block_mma_pack(const BlockMmaPlan& p, const float* const w_tf[4], int cin, uint32_t* dst) {
  for (int s = 0; s < 4; ++s) {
    const int ich = s == 0 ? p.cin_ch : p.nt;
    if (ich == 0) continue;
    const int ci_n = s == 0 ? cin : p.c, co_n = p.c;
    const float* w = w_tf[s];
    auto frag = [&](int ky, int kx, int plane, int nt, int lane) -> uint32_t {
      const int g8 = lane >> 2, t4 = lane & 3;
      const int co = nt * 8 + g8;
      uint32_t v = 0;
      for (int h = 0; h < 2; ++h) {
        const int ci = plane * 8 + 2 * t4 + h;
        const float f = (ci < ci_n && co < co_n) ? w[((size_t)(ky * 3 + kx) * ci_n + ci) * co_n + co] : 0.f;
        v |= (uint32_t)host_f_to_act(f) << (16 * h);
      }
      return v;
    };
    uint32_t* d = dst + (size_t)p.wbase[s] * 32;
    for (int lane = 0; lane < 32; ++lane) {
      if (ich == 1) {
        for (int ky = 0; ky < 3; ++ky)
          for (int nt = 0; nt < p.nt; ++nt)
            for (int kx = 0; kx < 3; ++kx) d[(size_t)((ky * p.nt + nt) * 3 + kx) * 32 + lane] = frag(ky, kx, 0, nt, lane);
      } else {
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx)
            for (int cp = 0; cp < ich / 2; ++cp)
              for (int nt = 0; nt < p.nt; ++nt)
                for (int h = 0; h < 2; ++h)
                  d[(size_t)((((ky * 3 + kx) * (ich / 2) + cp) * p.nt + nt) * 2 + h) * 32 + lane] = frag(ky, kx, 2 * cp + h, nt, lane);
      }
    }
  }
}

cudaError_t launch_block_mma(cudaStream_t st, const BlockMmaPlan& p, PV in, PV out, PV pool, const Geo* pool_geo,
                             const uint32_t* wfrag, const float* bias4, const Geo& g, int* err_flag) {
  if (!p.ok || out.chunks != p.nt || in.chunks != (p.cin_ch ? p.cin_ch : p.nt) || (pool.p && !pool_geo))
    return cudaErrorInvalidValue;
  BlockArgs a{};
  a.in = in.p; a.in_plane = in.plane;
  a.out = out.p; a.out_plane = out.plane;
  a.pool = pool.p; a.pool_plane = pool.plane;
  if (pool_geo) a.pg = *pool_geo;
  a.wfrag = wfrag; a.bias = bias4;
  a.g = g;
  a.WC = p.WC; a.WS = p.WS; a.pitch = p.WC + 2;
  a.n_strips = p.n_strips; a.rows_per_unit = p.rows_per_unit; a.units_per_page = p.units_per_page; a.n_units = p.n_units;
  for (int s = 0; s < 4; ++s) a.wbase[s] = p.wbase[s];
  a.err_flag = err_flag;
  void (*k)(BlockArgs) = nullptr;
  const bool pl = pool.p != nullptr;
  if (p.nt == 1 && p.cin_ch == 0) k = pl ? k_block_mma<0, 1, true> : k_block_mma<0, 1, false>;
  else if (p.nt == 1 && p.cin_ch == 1) k = pl ? k_block_mma<1, 1, true> : k_block_mma<1, 1, false>;
  else if (p.nt == 1 && p.cin_ch == 2) k = pl ? k_block_mma<2, 1, true> : k_block_mma<2, 1, false>;
  if (!k) return cudaErrorInvalidValue;
  static thread_local void* attr_done[8] = {nullptr};
  bool seen = false;
  for (void* q : attr_done) seen = seen || q == (void*)k;
  if (!seen) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    for (void*& q : attr_done) if (!q) { q = (void*)k; break; }
  }
  k<<<p.grid, p.threads, p.smem_bytes, st>>>(a);
  return cudaGetLastError();
}

}  // namespace aru
