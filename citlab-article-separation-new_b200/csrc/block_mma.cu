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
// Rings are made of BANDS of RB rows aligned with the bands their producer writes (a stage's output band, or the RB rows
// the loader brings in per iteration).  The window a consumer reads - rows a-1 .. a+RB of its own band [a, a+RB) - is
// then always "rows 2, 3 of one producer band and all four rows of the next", whatever the iteration: no per-row ring
// arithmetic, a row address is (band address, uniform) + (row offset, a per-lane constant).
constexpr int NB_IN = 3;          // input ring: the band being loaded + the two being read
constexpr int NB_X0 = 5;          // x0 ring: written by conv1 (or the loader), read by convR_0 and, 3 (RB+1) rows later, by convR_2
constexpr int NB_R = 3;

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

// number of B-fragment registers of a stage with ICH input planes and NT output planes
__host__ __device__ constexpr int stage_regs(int ich, int nt) { return ich == 1 ? 9 * nt : 9 * ich * nt; }

template <int N>
__device__ __forceinline__ void load_wfrag(uint32_t (&w)[N], const uint32_t* __restrict__ img, int base, int lane) {
#pragma unroll
  for (int r = 0; r < N; ++r) w[r] = __ldg(img + (size_t)(base + r) * 32 + lane);
}

// Per-lane byte offsets inside a band of a ring whose rows are `rb` bytes apart.
struct LaneOff {
  uint32_t a[4];    // ldmatrix.x4 of row k: (kx = -1 | kx = 0) for one input plane, (plane 2cp | plane 2cp+1) at kx = -1 otherwise
  uint32_t p[3];    // one input plane: ldmatrix.x4 of the kx = +1 halves of rows k (matrices 0, 1) and k + 1 (matrices 2, 3)
  uint32_t p_lo;    // the kx = +1 lane offset alone (row offset added by the caller)
};

// One stage of one band: acc[i][nt] += the 3x3 convolution.  The RB + 2 input rows are rows 0, 1 at bx (+ k rb) and rows
// 2 .. RB+1 at by (+ k rb).
template <int ICH, int NT, bool RELU_IN>
__device__ __forceinline__ void stage_mma(float (&acc)[RB][NT][4], uint32_t bx, uint32_t by, uint32_t rb, int pitch,
                                          const LaneOff& lo, bool upper, const uint32_t (&w)[stage_regs(ICH, NT)]) {
  static_assert(RB == 4, "window layout");
  if constexpr (ICH == 1) {
    // weights: w[(ky * NT + nt) * 3 + kx]
    uint32_t a[RB + 2][4], c[RB + 1][4];
#pragma unroll
    for (int j = 0; j < RB + 2; ++j) ldsm4(a[j], (j < 2 ? bx + lo.a[j] : by + lo.a[j - 2]));
    ldsm4(c[0], bx + lo.p[0]);
    ldsm4(c[1], (upper ? by : bx + rb) + lo.p_lo);     // row 1 (old band) | row 2 (new band)
#pragma unroll
    for (int j = 2; j < RB + 1; ++j) ldsm4(c[j], by + lo.p[j - 2]);
    if (RELU_IN) {
#pragma unroll
      for (int j = 0; j < RB + 2; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) a[j][q] = relu2(a[j][q]);
#pragma unroll
      for (int j = 0; j < RB + 1; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) c[j][q] = relu2(c[j][q]);
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int i = 0; i < RB; ++i)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          mma16(acc[i][nt], a[i + ky][0], a[i + ky][1], a[i + ky][2], a[i + ky][3], w[(ky * NT + nt) * 3], w[(ky * NT + nt) * 3 + 1]);
    // kx = +1: rows i, i+1 (ky = 0, 1) as one K = 16 step, row i+2 (ky = 2) as a K = 8 step
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        mma16(acc[i][nt], c[i][0], c[i][1], c[i][2], c[i][3], w[(0 * NT + nt) * 3 + 2], w[(1 * NT + nt) * 3 + 2]);
        if (i + 2 < RB + 1) mma8(acc[i][nt], c[i + 2][0], c[i + 2][1], w[(2 * NT + nt) * 3 + 2]);
        else mma8(acc[i][nt], c[i + 1][2], c[i + 1][3], w[(2 * NT + nt) * 3 + 2]);
      }
  } else {
    // weights: w[(((ky * 3 + kx) * (ICH / 2) + cp) * NT + nt) * 2 + {0, 1}]
#pragma unroll
    for (int j = 0; j < RB + 2; ++j) {
      const uint32_t row = j < 2 ? bx + lo.a[j] : by + lo.a[j - 2];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int cp = 0; cp < ICH / 2; ++cp) {
          uint32_t a[4];
          ldsm4(a, row + (uint32_t)(cp * 2 * pitch + kx) * 16u);
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

template <int NT>
__device__ __forceinline__ void init_acc(float (&acc)[RB][NT][4], const float (&b)[NT][2]) {
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      acc[i][nt][0] = acc[i][nt][2] = b[nt][0];
      acc[i][nt][1] = acc[i][nt][3] = b[nt][1];
    }
}

// (ReLU) + 16-bit store of a band into its ring band at `band`; MASK: out-of-image positions become zeros
template <int NT, bool RELU, bool MASK>
__device__ __forceinline__ void store_band(const float (&acc)[RB][NT][4], uint32_t band, const uint32_t (&st)[RB], int pitch,
                                           int row0, int H, bool cv0, bool cv1) {
#pragma unroll
  for (int i = 0; i < RB; ++i) {
    const bool rv = (unsigned)(row0 + i) < (unsigned)H;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t h0 = pack2_fin<RELU>(acc[i][nt][0], acc[i][nt][1]);
      uint32_t h1 = pack2_fin<RELU>(acc[i][nt][2], acc[i][nt][3]);
      if (MASK) {
        if (!(rv && cv0)) h0 = 0u;
        if (!(rv && cv1)) h1 = 0u;
      }
      sts32(band + st[i] + (uint32_t)(nt * pitch) * 16u, h0);
      sts32(band + st[i] + (uint32_t)(nt * pitch + 8) * 16u, h1);
    }
  }
}

__device__ __forceinline__ uint32_t band_addr(uint32_t base, int cur, int back, int nb, uint32_t band_bytes) {
  int q = cur - back;
  q += q < 0 ? nb : 0;
  return base + (uint32_t)q * band_bytes;
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
  constexpr int NB_LD = CONV1 ? NB_IN : NB_X0;
  constexpr int ICH0 = CONV1 ? CIN_CH : 1;

  // shared-memory carve-up: [mbarriers 128 B][input ring (CONV1)][x0 ring][r0 ring][r1 ring]
  const uint32_t rb_in = (uint32_t)(ICH0 * pitch) * 16u, rb = (uint32_t)(NT * pitch) * 16u;   // bytes per ring row
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_bar = s_base;
  const uint32_t s_in = s_base + 128;
  const uint32_t s_x0 = s_in + (CONV1 ? (uint32_t)(NB_IN * RB) * rb_in : 0u);
  const uint32_t s_r0 = s_x0 + (uint32_t)(NB_X0 * RB) * rb;
  const uint32_t s_r1 = s_r0 + (uint32_t)(NB_R * RB) * rb;
  const uint32_t s_end = s_r1 + (uint32_t)(NB_R * RB) * rb;
  const uint32_t s_ld = CONV1 ? s_in : s_x0;
  const uint32_t rb_ld = CONV1 ? rb_in : rb;

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
  const bool upper = m >= 2;
  // position u of a ring row lives at byte (1 + u) * 16 of its plane
  LaneOff lo_in, lo_x;
  {
    const uint32_t a1 = (uint32_t)(1 + u0 + (m < 2 ? -1 : 0) + (m & 1) * 8 + r8) * 16u;            // one plane: kx -1 | kx 0
    const uint32_t ac = (uint32_t)((m >> 1) * pitch + u0 + (m & 1) * 8 + r8) * 16u;                // plane pairs, kx = -1
    const uint32_t p1 = (uint32_t)(1 + u0 + 1 + (m & 1) * 8 + r8) * 16u;                           // kx = +1
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo_in.a[k] = (ICH0 == 1 ? a1 : ac) + (uint32_t)k * rb_in;
      lo_x.a[k] = (NT == 1 ? a1 : ac) + (uint32_t)k * rb;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      lo_in.p[k] = p1 + (uint32_t)(k + (upper ? 1 : 0)) * rb_in;
      lo_x.p[k] = p1 + (uint32_t)(k + (upper ? 1 : 0)) * rb;
    }
    lo_in.p_lo = p1;
    lo_x.p_lo = p1;
  }
  uint32_t st[RB];   // store offsets of this lane's first position in rows 0..3 of a band
#pragma unroll
  for (int k = 0; k < RB; ++k) st[k] = (uint32_t)(1 + u0 + g8) * 16u + (uint32_t)t4 * 4u + (uint32_t)k * rb;

  // filters and biases: registers for the whole launch
  uint32_t w0[stage_regs(ICH0, NT)];
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
    const bool cols_in = xs - 4 + u0 >= 0 && xs - 4 + u0 + 16 <= W;     // warp-uniform: no column of this warp needs a mask
    const int x_own_end = min(xs + a.WS, W);
    const bool own0 = xc0 >= xs && xc0 < x_own_end, own1 = xc1 >= xs && xc1 < x_own_end;
    const long long page_row0 = a.g.lead + ((long long)n * a.g.Hp + 1) * a.g.Wp + 1;   // position of (y = 0, x = 0)
    act_t* const out0 = a.out + (page_row0 + xc0) * 8 + 2 * t4;
    act_t* const pool0 = POOL ? a.pool + (a.pg.lead + ((long long)n * a.pg.Hp + 1) * a.pg.Wp + 1 + (xc0 >> 1)) * 8 + 2 * t4 : nullptr;

    auto stage_active = [&](int s, int t) {
      const int as = a3_0 + t * RB + (3 - s) * LAG;
      return as < y1 + (3 - s) && as + RB > y0 - (3 - s);
    };
    // loader: the rows the first stage reads in iteration t that no earlier iteration brought in = band `cur` of the
    // loader's ring (rows af + 1 .. af + RB); iteration 0 also needs rows af - 1, af = rows 2, 3 of band cur - 1
    auto issue_loads = [&](int t, uint32_t it_t, int cur) {
      const uint32_t bar = s_bar + (it_t & 1) * 8;
      const int af = a3_0 + t * RB + (3 - FIRST) * LAG;
      int r_lo = t == 0 ? af - 1 : af + 1, r_hi = af + RB;   // inclusive
      if (!stage_active(FIRST, t)) r_hi = r_lo - 1;
      const int c_lo = max(r_lo, -1), c_hi = min(r_hi, H);
      const int rows = max(c_hi - c_lo + 1, 0);
      const uint32_t plane_bytes = (uint32_t)pitch * 16u;
      mbar_expect_tx(bar, (uint32_t)(rows * LD_CH) * plane_bytes);
      for (int r = c_lo; r <= c_hi; ++r) {
        const int k = r - (af + 1);    // row inside band cur (negative: band cur - 1)
        const uint32_t dst = k >= 0 ? band_addr(s_ld, cur, 0, NB_LD, RB * rb_ld) + (uint32_t)k * rb_ld
                                    : band_addr(s_ld, cur, 1, NB_LD, RB * rb_ld) + (uint32_t)(k + RB) * rb_ld;
        const long long p = page_row0 + (long long)r * a.g.Wp + (xs - 5);
#pragma unroll
        for (int c = 0; c < LD_CH; ++c)
          bulk_g2s(dst + (uint32_t)c * plane_bytes, a.in + ((long long)c * a.in_plane + p) * 8, plane_bytes, bar);
      }
    };

    // band counters of the rings (the band their producer fills / filled for this iteration)
    int c_ld = 1, c_x0 = 0, c_r0 = 0, c_r1 = 0;
    if (warp == n_cw && lane == 0) issue_loads(0, it, c_ld);
    for (int t = 0; t < n_iter; ++t, ++it) {
      if (warp == n_cw) {
        if (lane == 0 && t + 1 < n_iter) issue_loads(t + 1, it + 1, c_ld + 1 == NB_LD ? 0 : c_ld + 1);
      } else {
        mbar_wait(s_bar + (it & 1) * 8, (it >> 1) & 1, a.err_flag, 41);
        const int a3 = a3_0 + t * RB;
        float acc[RB][NT][4];
        if constexpr (CONV1) if (stage_active(0, t)) {
          const int as = a3 + 3 * LAG;
          init_acc<NT>(acc, bs[0]);
          stage_mma<ICH0, NT, false>(acc, band_addr(s_in, c_ld, 1, NB_IN, RB * rb_in) + 2 * rb_in,
                                     band_addr(s_in, c_ld, 0, NB_IN, RB * rb_in), rb_in, pitch, lo_in, upper, w0);
          const uint32_t dst = band_addr(s_x0, c_x0, 0, NB_X0, RB * rb);
          if (cols_in && as >= 0 && as + RB <= H) store_band<NT, false, false>(acc, dst, st, pitch, as, H, cv0, cv1);
          else store_band<NT, false, true>(acc, dst, st, pitch, as, H, cv0, cv1);
        }
        if (stage_active(1, t)) {
          const int as = a3 + 2 * LAG;
          init_acc<NT>(acc, bs[1]);
          // conv1 fused: x0 bands 2 and 1 back; loader-fed: bands 1 and 0 back (the loader's band counter)
          const int cx = CONV1 ? c_x0 : c_ld, back = CONV1 ? 1 : 0;
          stage_mma<NT, NT, true>(acc, band_addr(s_x0, cx, back + 1, NB_X0, RB * rb) + 2 * rb,
                                  band_addr(s_x0, cx, back, NB_X0, RB * rb), rb, pitch, lo_x, upper, w1);
          const uint32_t dst = band_addr(s_r0, c_r0, 0, NB_R, RB * rb);
          if (cols_in && as >= 0 && as + RB <= H) store_band<NT, true, false>(acc, dst, st, pitch, as, H, cv0, cv1);
          else store_band<NT, true, true>(acc, dst, st, pitch, as, H, cv0, cv1);
        }
        if (stage_active(2, t)) {
          const int as = a3 + LAG;
          init_acc<NT>(acc, bs[2]);
          stage_mma<NT, NT, false>(acc, band_addr(s_r0, c_r0, 2, NB_R, RB * rb) + 2 * rb, band_addr(s_r0, c_r0, 1, NB_R, RB * rb),
                                   rb, pitch, lo_x, upper, w2);
          const uint32_t dst = band_addr(s_r1, c_r1, 0, NB_R, RB * rb);
          if (cols_in && as >= 0 && as + RB <= H) store_band<NT, true, false>(acc, dst, st, pitch, as, H, cv0, cv1);
          else store_band<NT, true, true>(acc, dst, st, pitch, as, H, cv0, cv1);
        }
        if (stage_active(3, t)) {
          init_acc<NT>(acc, bs[3]);
          stage_mma<NT, NT, false>(acc, band_addr(s_r1, c_r1, 2, NB_R, RB * rb) + 2 * rb, band_addr(s_r1, c_r1, 1, NB_R, RB * rb),
                                   rb, pitch, lo_x, upper, w3);
          // + x0 (pre-activation of conv1) -> ReLU -> global (+ pooled copy).  Rows a3 .. a3+3 of x0 are rows 1, 2, 3 of
          // one band and row 0 of the next: 4 / 3 bands back (conv1 fused), 3 / 2 bands back (loader-fed)
          const int cx = CONV1 ? c_x0 : c_ld, back = CONV1 ? 4 : 3;
          const uint32_t xa = band_addr(s_x0, cx, back, NB_X0, RB * rb), xb = band_addr(s_x0, cx, back - 1, NB_X0, RB * rb);
          const bool edge = !(cols_in && a3 >= 0 && a3 + RB <= H);
          uint32_t hp[RB][NT][2];
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int y = a3 + i;
            const bool rv = (unsigned)y < (unsigned)H && y >= y0 && y < y1;
            const uint32_t xr = i < 3 ? xa + st[i + 1] : xb + st[0];
            act_t* const orow = out0 + (long long)y * a.g.Wp * 8;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const uint32_t q0 = lds32(xr + (uint32_t)(nt * pitch) * 16u);
              const uint32_t q1 = lds32(xr + (uint32_t)(nt * pitch + 8) * 16u);
              const float2 f0 = act2_to_f2(*reinterpret_cast<const act2_t*>(&q0));
              const float2 f1 = act2_to_f2(*reinterpret_cast<const act2_t*>(&q1));
              uint32_t h0 = pack2_fin<true>(acc[i][nt][0] + f0.x, acc[i][nt][1] + f0.y);
              uint32_t h1 = pack2_fin<true>(acc[i][nt][2] + f1.x, acc[i][nt][3] + f1.y);
              if (POOL && edge) {   // the pooling windows of the last row / column take zeros for what lies outside
                if (!(rv && cv0)) h0 = 0u;
                if (!(rv && cv1)) h1 = 0u;
              }
              hp[i][nt][0] = h0;
              hp[i][nt][1] = h1;
              if (rv && own0) *reinterpret_cast<uint32_t*>(orow + (long long)nt * a.out_plane * 8) = h0;
              if (rv && own1) *reinterpret_cast<uint32_t*>(orow + (long long)nt * a.out_plane * 8 + 64) = h1;
            }
          }
          if (POOL) {
            // a3 is a multiple of RB (even): the band holds whole pooling windows; column pairs are lane pairs (g8, g8 ^ 1)
#pragma unroll
            for (int i = 0; i < RB; i += 2) {
              const int y = a3 + i;
              const bool rv = (unsigned)y < (unsigned)H && y >= y0 && y < y1;
              act_t* const prow = pool0 + (long long)(y >> 1) * a.pg.Wp * 8;
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) {
                uint32_t m0 = max2(hp[i][nt][0], hp[i + 1][nt][0]);
                uint32_t m1 = max2(hp[i][nt][1], hp[i + 1][nt][1]);
                m0 = max2(m0, __shfl_xor_sync(0xffffffffu, m0, 4));
                m1 = max2(m1, __shfl_xor_sync(0xffffffffu, m1, 4));
                if (!(g8 & 1) && rv && own0) *reinterpret_cast<uint32_t*>(prow + (long long)nt * a.pool_plane * 8) = m0;
                if (!(g8 & 1) && rv && own1) *reinterpret_cast<uint32_t*>(prow + (long long)nt * a.pool_plane * 8 + 32) = m1;
              }
            }
          }
        }
      }
      c_ld = c_ld + 1 == NB_LD ? 0 : c_ld + 1;
      c_x0 = c_x0 + 1 == NB_X0 ? 0 : c_x0 + 1;
      c_r0 = c_r0 + 1 == NB_R ? 0 : c_r0 + 1;
      c_r1 = c_r1 + 1 == NB_R ? 0 : c_r1 + 1;
      __syncthreads();
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
size_t block_smem_bytes(int cin_ch, int nt, int wc) {
  const size_t pitch = wc + 2;
  size_t rows = (size_t)(cin_ch > 0 ? NB_IN * RB * cin_ch : 0) + (size_t)(NB_X0 * RB) * nt + 2 * (size_t)(NB_R * RB) * nt;
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
