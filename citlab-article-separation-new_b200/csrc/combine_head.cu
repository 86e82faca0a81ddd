// combine_head.cu - the attention tail and the classifier in ONE launch (sm_100a):
//     out = act( conv4x4( sum_k upsample(det_k) * softmax_k(upsample(att_k)) ) + bias )        (ARU_v1.py:115-153, :155-160)
// The combined 8-channel map (16 B per pixel written by k_combine and read back by the classifier = 2.5 GB per 32 pages of
// 1856x1344, a third of the two launches' traffic) never reaches HBM.
//
// A CTA owns a tile of TH x 128 output pixels (TH = 16 .. 32, chosen so that three CTAs share an SM: while one waits for
// its operands the other two compute).
//   phase 0: one warp streams the tile + 4x4 SAME halo (one row / column before, two after) of the full-resolution
//            detection map into shared memory with one bulk copy (TMA engine) per row - 60 KB in flight per CTA, no
//            register staging; meanwhile all threads fill small shared-memory tables with the tile's low-resolution
//            operands: the attention logits of every scale and the channel sums of the low-resolution detection maps
//            (the ones-filter quirk of upsample_simple), each source pixel read ONCE per tile, all loads independent;
//   phase 1: in place, per position: softmax over the scales, the weighted low-resolution sums, eight FMAs, one rounding
//            to 16 bits - same arithmetic, operation order and rounding as k_combine (kernels_simple.cu), so the
//            intermediate equals the two-launch path bit for bit; out-of-image positions become zeros (SAME padding);
//   phase 2: warp w owns output columns [16 w, 16 w + 16) and walks down the tile in bands of 4 rows on the warp-level
//            tensor path: an ldmatrix fragment of an input row holds two horizontal taps (K = 16 = 2 taps x 8 channels),
//            8 mma.sync.m16n8k16 per 16 pixels, logits in fp32, softmax / sigmoid in the epilogue, float32 NHWC stores.
//            (N = 8 output channels of which n_class <= 2 are real: the launch is bound by HBM, not by these.)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels.h"
#include "band_common.cuh"

namespace aru {
namespace {

constexpr int CH_TW = 128;
constexpr int CH_COLS = CH_TW + 3;
constexpr int CH_PITCH = 132;   // positions per shared-memory row
constexpr int CH_THREADS = 256;
constexpr int CH_TILE_OFF = 128;   // the mbarrier lives in front of the tile
constexpr int CH_DET_ITEMS = 8;    // low-resolution detection cells a thread fetches per tile

struct CombineHeadArgs {
  CombineArgs c;            // the combine's operands (c.out unused)
  const uint32_t* wfrag;    // classifier B fragments: 16 registers x 32 lanes (combine_head_pack)
  float bias[8];
  float* out;               // dense float32 NHWC [N][H][W][C]
  int C, act;
  int kf;                   // the scale whose detection map has full resolution
  int TH;                   // output rows per tile (multiple of 4)
  int tab_att[ARU_COMBINE_MAX], tw_att[ARU_COMBINE_MAX];   // float offset (behind the tile) and row pitch of the tables
  int tab_det[ARU_COMBINE_MAX], tw_det[ARU_COMBINE_MAX];
  int tab_rep, tab_row, tab_wt;   // first row / column of a class, per-row offsets, softmax weights per class pair
  int low_k[ARU_COMBINE_MAX];     // the scales with an upsampled detection map, ascending
  int* err_flag;
};

#ifdef ARU_USE_BF16
#define ARU_MMA_T "bf16"
#else
#define ARU_MMA_T "f16"
#endif
__device__ __forceinline__ void mma16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." ARU_MMA_T "." ARU_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// d = a * b + {c0, c1, c0, c1}
__device__ __forceinline__ void mma16_init(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." ARU_MMA_T "." ARU_MMA_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&a)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr) : "memory");
}

// source cell of pixel coordinate v (clamped to the map) for an upsample factor 2^sh with crop offset `off`
// (combine_head_ok admits powers of two only: a general division per cell made the kernel instruction-fetch bound)
__device__ __forceinline__ int cell_of(int v, int off, int sh, int limit) { return min((v + off) >> sh, limit - 1); }
// floor(i / d) for 0 <= i < 2^16, 1 <= d < 2^10, rcp = 1.f / d
__device__ __forceinline__ int small_div(int i, float rcp) { return __float2int_rz(((float)i + 0.5f) * rcp); }

template <int A>
__global__ void __launch_bounds__(CH_THREADS, 3) k_combine_head(const __grid_constant__ CombineHeadArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NL = A - 1;   // scales whose detection map is upsampled (exactly one has full resolution)
  const CombineArgs& ca = a.c;
  const int H = ca.geo.H, W = ca.geo.W, TH = a.TH;
  const int n = blockIdx.z, ty0 = blockIdx.y * TH, tx0 = blockIdx.x * CH_TW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rows = min(TH + 3, H - ty0 + 3);   // tile rows that can feed an in-image output row
  // the in-image window of the tile (inclusive)
  const int ya = max(ty0 - 1, 0), yb = min(ty0 - 2 + rows, H - 1);
  const int xa = max(tx0 - 1, 0), xb = min(tx0 + CH_TW + 1, W - 1);
  unsigned char* const tile = smem + CH_TILE_OFF;
  float* const tabs = reinterpret_cast<float*>(tile + (size_t)(TH + 3) * CH_PITCH * 16);
  int* const itabs = reinterpret_cast<int*>(tabs);
  const uint32_t s_bar = smem_u32(smem);
  if (tid == 0) {
    mbar_init(s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // ---- phase 0: operands ------------------------------------------------------------------------------------------
  if (warp == 0) {
    const uint32_t row_bytes = (uint32_t)(xb - xa + 1) * 16u;
    if (lane == 0) mbar_expect_tx(s_bar, (uint32_t)(yb - ya + 1) * row_bytes);
    __syncwarp();
    const act_t* src = ca.det[a.kf] + ca.geo.pos(n, ya + lane, xa) * 8;
    uint32_t dst = smem_u32(tile) + (uint32_t)((ya + lane - (ty0 - 1)) * CH_PITCH + (xa - (tx0 - 1))) * 16u;
    for (int y = ya + lane; y <= yb; y += 32, src += 32LL * ca.geo.Wp * 8, dst += 32u * CH_PITCH * 16u)
      bulk_g2s(dst, src, row_bytes, s_bar);
  }
  // The low-resolution operands of the tile, every source pixel once: ALL loads of a thread are issued before the first
  // use (one exposed memory latency per CTA; they travel together with the bulk copies above).
  constexpr int NLX = NL > 0 ? NL : 1;
  int ar0[A], ac0[A];        // first attention cell of the tile per scale
  int dr0[NLX], dc0[NLX];    // first low-resolution detection cell
#pragma unroll
  for (int k = 0; k < A; ++k) {
    ar0[k] = cell_of(ya, ca.att_oy[k], ca.att_sh[k], ca.att_h[k]);
    ac0[k] = cell_of(xa, ca.att_ox[k], ca.att_sh[k], ca.att_w[k]);
  }
#pragma unroll
  for (int t = 0; t < NL; ++t) {
    const int k = a.low_k[t];
    dr0[t] = cell_of(ya, ca.det_oy[k], ca.det_sh[k], ca.det_geo[k].H);
    dc0[t] = cell_of(xa, ca.det_ox[k], ca.det_sh[k], ca.det_geo[k].W);
  }
  //   attention logits: at most one cell per thread and scale (combine_head_ok: a tile touches <= 256 cells per scale)
  float av[A];
  int a_dst[A];
#pragma unroll
  for (int k = 0; k < A; ++k) {
    const int nr = cell_of(yb, ca.att_oy[k], ca.att_sh[k], ca.att_h[k]) - ar0[k] + 1;
    const int nc = cell_of(xb, ca.att_ox[k], ca.att_sh[k], ca.att_w[k]) - ac0[k] + 1;
    av[k] = 0.f;
    a_dst[k] = -1;
    if (tid < nr * nc) {
      const int r = small_div(tid, 1.f / (float)nc), c = tid - r * nc;
      av[k] = __ldg(ca.att[k] + ((long long)n * ca.att_h[k] + ar0[k] + r) * ca.att_w[k] + ac0[k] + c);
      a_dst[k] = a.tab_att[k] + r * a.tw_att[k] + c;
    }
  }
  //   low-resolution detection maps: one flat list over the scales (combine_head_ok: <= 256 * CH_DET_ITEMS cells)
  uint4 dv[CH_DET_ITEMS];
  int d_dst[CH_DET_ITEMS];
  {
    int cnt[NLX], nc[NLX];
    float rcp[NLX];
    const uint4* base[NLX];
#pragma unroll
    for (int t = 0; t < NL; ++t) {
      const int k = a.low_k[t];
      nc[t] = cell_of(xb, ca.det_ox[k], ca.det_sh[k], ca.det_geo[k].W) - dc0[t] + 1;
      cnt[t] = (cell_of(yb, ca.det_oy[k], ca.det_sh[k], ca.det_geo[k].H) - dr0[t] + 1) * nc[t];
      rcp[t] = 1.f / (float)nc[t];
      base[t] = reinterpret_cast<const uint4*>(ca.det[k]) + ca.det_geo[k].pos(n, dr0[t], dc0[t]);
    }
#pragma unroll
    for (int i = 0; i < CH_DET_ITEMS; ++i) {
      int rem = tid + i * CH_THREADS;
      d_dst[i] = -1;
      dv[i] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int t = 0; t < NL; ++t) {
        const int k = a.low_k[t];
        const bool hit = rem >= 0 && rem < cnt[t];
        if (hit) {
          const int r = small_div(rem, rcp[t]), c = rem - r * nc[t];
          dv[i] = __ldg(base[t] + r * ca.det_geo[k].Wp + c);
          d_dst[i] = a.tab_det[k] + r * a.tw_det[k] + c;
        }
        rem = hit ? -1 : rem - cnt[t];
      }
    }
  }
  // Softmax weights change only where a scale's source cell changes: rows / columns of the tile fall into classes
  // (class = sum over the scales of the cell index relative to the tile's first cell), the weights are evaluated once
  // per (row class, column class); the thread that finds a class's first row / column records it.
  auto row_class = [&](int y) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < A; ++k)
      c += cell_of(y, ca.att_oy[k], ca.att_sh[k], ca.att_h[k]) - ar0[k];
    return c;
  };
  auto col_class = [&](int x) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < A; ++k)
      c += cell_of(x, ca.att_ox[k], ca.att_sh[k], ca.att_w[k]) - ac0[k];
    return c;
  };
  const int n_cc = col_class(xb) + 1, n_rc = row_class(yb) + 1;
  int* const rep_r = itabs + a.tab_rep;           // first tile row of a row class
  int* const rep_c = rep_r + (TH + 3);            // first tile column of a column class
  int* const rowtab = itabs + a.tab_row;          // per tile row: weight-table offset of its class, low-resolution row offsets
  float* const wt = tabs + a.tab_wt;              // [row class][column class][A]: weight of scale kf, then of the others
  if (tid < rows) {
    const int yc = min(max(ty0 - 1 + tid, ya), yb);
    const int rc = row_class(yc);
    if (tid == 0 || row_class(min(max(ty0 - 2 + tid, ya), yb)) != rc) rep_r[rc] = tid;
    rowtab[tid * A] = rc * n_cc * A;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int k = a.low_k[i];
      rowtab[tid * A + 1 + i] = a.tab_det[k] + (cell_of(yc, ca.det_oy[k], ca.det_sh[k], ca.det_geo[k].H) - dr0[i]) * a.tw_det[k];
    }
  } else if (tid >= 64 && tid < 64 + CH_COLS) {
    const int c = tid - 64;
    const int xc = min(max(tx0 - 1 + c, xa), xb);
    const int cc = col_class(xc);
    if (c == 0 || col_class(min(max(tx0 - 2 + c, xa), xb)) != cc) rep_c[cc] = c;
  }
#pragma unroll
  for (int k = 0; k < A; ++k)
    if (a_dst[k] >= 0) tabs[a_dst[k]] = av[k];
#pragma unroll
  for (int i = 0; i < CH_DET_ITEMS; ++i)
    if (d_dst[i] >= 0) {   // channel sum of the source pixel (the ones-filter quirk; k_combine's expression)
      float d[8];
      unpack8(dv[i], d);
      tabs[d_dst[i]] = ((d[0] + d[1]) + (d[2] + d[3])) + ((d[4] + d[5]) + (d[6] + d[7]));
    }
  __syncthreads();
  const float rcp_cc = 1.f / (float)n_cc;
  for (int idx = tid; idx < n_rc * n_cc; idx += CH_THREADS) {   // softmax over the scales, once per class pair
    const int rc = small_div(idx, rcp_cc), cc = idx - rc * n_cc;
    const int y = min(max(ty0 - 1 + rep_r[rc], ya), yb), x = min(max(tx0 - 1 + rep_c[cc], xa), xb);
    float wgt[A];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < A; ++k) {
      const int r = cell_of(y, ca.att_oy[k], ca.att_sh[k], ca.att_h[k]) - ar0[k];
      const int c = cell_of(x, ca.att_ox[k], ca.att_sh[k], ca.att_w[k]) - ac0[k];
      wgt[k] = tabs[a.tab_att[k] + r * a.tw_att[k] + c];
      m = fmaxf(m, wgt[k]);
    }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < A; ++k) {
      wgt[k] = __expf(wgt[k] - m);
      den += wgt[k];
    }
    const float inv = 1.f / den;
    int slot = 1;
#pragma unroll
    for (int k = 0; k < A; ++k) wt[idx * A + (k == a.kf ? 0 : slot++)] = wgt[k] * inv;
  }
  __syncthreads();
  mbar_wait_sleep(s_bar, 0, a.err_flag, 51);
  // ---- phase 1: the combined map of the tile + halo, in place -------------------------------------------------------
  {
    // per column: offset of its class in a weight-table row, offsets of its low-resolution source columns
    auto col_off = [&](int c, int& ow, int (&ol)[NL > 0 ? NL : 1]) {
      const int xc = min(max(tx0 - 1 + c, xa), xb);
      ow = col_class(xc) * A;
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        const int k = a.low_k[i];
        ol[i] = cell_of(xc, ca.det_ox[k], ca.det_sh[k], ca.det_geo[k].W) - dc0[i];
      }
    };
    auto pixel = [&](int j, int c, bool x_in, int ow, const int (&ol)[NL > 0 ? NL : 1]) {
      uint4* const cell = reinterpret_cast<uint4*>(tile + ((size_t)j * CH_PITCH + c) * 16);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (x_in && (unsigned)(ty0 - 1 + j) < (unsigned)H) {
        const int* rt = rowtab + j * A;
        const float* wp = wt + rt[0] + ow;
        float low = 0.f;   // k_combine's order: scales ascending, full-resolution map last
#pragma unroll
        for (int i = 0; i < NL; ++i) low = fmaf(tabs[rt[1 + i] + ol[i]], wp[1 + i], low);
        const float wk = wp[0];
        float d[8], acc[8];
        unpack8(*cell, d);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fmaf(d[q], wk, low);
        v = pack8_fin<false>(acc, true);
      }
      *cell = v;
    };
    {   // thread = one of the first 128 columns, every other row
      const int c = tid & 127;
      int ow, ol[NL > 0 ? NL : 1];
      col_off(c, ow, ol);
      const bool x_in = (unsigned)(tx0 - 1 + c) < (unsigned)W;
#pragma unroll 2
      for (int j = tid >> 7; j < rows; j += 2) pixel(j, c, x_in, ow, ol);
    }
    if (tid < 3 * rows) {   // the last three halo columns
      const int j = small_div(tid, 1.f / 3.f), c = 128 + (tid - 3 * j);
      int ow, ol[NL > 0 ? NL : 1];
      col_off(c, ow, ol);
      pixel(j, c, (unsigned)(tx0 - 1 + c) < (unsigned)W, ow, ol);
    }
  }
  __syncthreads();
  // ---- phase 2: 4x4 convolution + activation ----------------------------------------------------------------------
  const int g8 = lane >> 2, t4 = lane & 3, m = lane >> 3, r8 = lane & 7;
  const int u0 = warp * 16;
  if (tx0 + u0 >= W) return;
  uint32_t w[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) w[r] = __ldg(a.wfrag + r * 32 + lane);
  const float b0 = a.bias[(2 * t4) & 7], b1 = a.bias[(2 * t4 + 1) & 7];
  // matrices 0, 1 = tap kx (positions 0..7, 8..15), matrices 2, 3 = tap kx + 1: position u + kx of a tile row
  const uint32_t s_lane = smem_u32(tile) + (uint32_t)(u0 + (m >> 1) + (m & 1) * 8 + r8) * 16u;
  const int C = a.C;
  const int x0 = tx0 + u0 + g8;
  float* const o_base = a.out + ((long long)n * H * W) * C;
#pragma unroll 1
  for (int i0 = 0; i0 < TH; i0 += 4) {
    if (ty0 + i0 >= H) break;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      uint32_t f01[4], f23[4];
      const uint32_t row = s_lane + (uint32_t)((i0 + j) * CH_PITCH) * 16u;
      ldsm4(f01, row);
      ldsm4(f23, row + 32u);
      // output row i = j - ky; the two MMAs of a row are issued apart (they chain through the accumulator); the first MMA
      // of an output row (ky = 0, taps kx 0, 1) takes the bias as its C operand
#pragma unroll
      for (int ky = 3; ky >= 0; --ky) {
        const int i = j - ky;
        if (i < 0 || i > 3) continue;
        if (ky == 0) mma16_init(acc[i], f01, w[0], w[1], b0, b1);
        else mma16(acc[i], f01, w[(ky * 2) * 2], w[(ky * 2) * 2 + 1]);
      }
#pragma unroll
      for (int ky = 3; ky >= 0; --ky) {
        const int i = j - ky;
        if (i < 0 || i > 3) continue;
        mma16(acc[i], f23, w[(ky * 2 + 1) * 2], w[(ky * 2 + 1) * 2 + 1]);
      }
    }
    // channels 0, 1 of row i sit in the lanes with t4 == 0: hand row i to lane t4 == i so that all 32 lanes finish a row
    float z[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float v = __shfl_sync(0xffffffffu, acc[i][q], lane & ~3);
        if (t4 == i) z[q] = v;
      }
    }
    const int y = ty0 + i0 + t4;
    if (y < H) {
      float* const orow = o_base + (long long)y * W * C;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int x = x0 + 8 * h;
        if (x >= W) continue;
        float z0 = z[2 * h], z1 = z[2 * h + 1];
        if (a.act == 2 && C == 2) {   // two-class softmax: p0 = 1 / (1 + e^(z1 - z0)), p1 = e^(z1 - z0) * p0
          const float e = __expf(fminf(z1 - z0, 80.f));
          const float p0 = __fdividef(1.f, 1.f + e);
          *reinterpret_cast<float2*>(orow + (long long)x * 2) = make_float2(p0, e * p0);
        } else {
          if (a.act == 1) { z0 = fmaxf(z0, 0.f); z1 = fmaxf(z1, 0.f); }
          else if (a.act == 2) { z0 = 1.f; }                                  // softmax over one class
          else if (a.act == 3) { z0 = __fdividef(1.f, 1.f + __expf(-z0)); z1 = __fdividef(1.f, 1.f + __expf(-z1)); }
          orow[(long long)x * C] = z0;
          if (C == 2) orow[(long long)x * C + 1] = z1;
        }
      }
    }
  }
}

// Shared-memory plan: the tallest tile (multiple of 4 rows) that lets three CTAs share an SM.
struct HeadPlan {
  bool ok = false;
  int kf = -1, TH = 0;
  int tab_att[ARU_COMBINE_MAX], tw_att[ARU_COMBINE_MAX], tab_det[ARU_COMBINE_MAX], tw_det[ARU_COMBINE_MAX];
  int tab_rep = 0, tab_row = 0, tab_wt = 0;
  size_t smem = 0;
};

HeadPlan head_plan(const CombineArgs& c) {
  HeadPlan p;
  int n_full = 0;
  for (int k = 0; k < c.A; ++k)
    if (c.det_up[k] == 1) { p.kf = k; ++n_full; }
  if (n_full != 1 || c.out_chunks != 1) return p;
  for (int k = 0; k < c.A; ++k)
    if (c.det_chunks[k] != 1) return p;
  for (int k = 0; k < c.A; ++k)
    if (c.att_up[k] < 1 || c.det_up[k] < 1 || (c.att_up[k] & (c.att_up[k] - 1)) || (c.det_up[k] & (c.det_up[k] - 1)) || c.att_oy[k] < 0 || c.att_ox[k] < 0 || c.det_oy[k] < 0 || c.det_ox[k] < 0) return p;
  static size_t sm_smem = 0;
  if (!sm_smem) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    sm_smem = (size_t)v;
  }
  for (int TH = 32; TH >= 16; TH -= 4) {
    int off = 0;
    auto table = [&](int up, int* tab, int* tw) {   // cells a window of (TH + 3) x CH_COLS pixels can touch
      *tab = off;
      *tw = (CH_COLS - 1) / up + 2;
      off += ((TH + 2) / up + 2) * *tw;
    };
    int n_rc = 1, n_cc = 1;   // upper bounds of the row / column classes of a tile
    int det_cells = 0;
    bool fits = true;         // the kernel's per-thread load lists
    for (int k = 0; k < c.A; ++k) {
      int before = off;
      table(c.att_up[k], &p.tab_att[k], &p.tw_att[k]);
      fits = fits && off - before <= CH_THREADS;
      before = off;
      if (c.det_up[k] != 1) table(c.det_up[k], &p.tab_det[k], &p.tw_det[k]);
      else p.tab_det[k] = p.tw_det[k] = 0;
      det_cells += off - before;
      n_rc += (TH + 2) / c.att_up[k] + 1;
      n_cc += (CH_COLS - 1) / c.att_up[k] + 1;
    }
    p.tab_rep = off; off += (TH + 3) + CH_COLS;
    p.tab_row = off; off += (TH + 3) * c.A;
    p.tab_wt = off; off += n_rc * n_cc * c.A;
    p.smem = CH_TILE_OFF + (size_t)(TH + 3) * CH_PITCH * 16 + (size_t)off * 4;
    if (!fits || det_cells > CH_THREADS * CH_DET_ITEMS) continue;
    if (3 * (p.smem + 1024) <= sm_smem) {
      p.TH = TH;
      p.ok = true;
      return p;
    }
  }
  return p;
}

}  // namespace

bool combine_head_ok(const CombineArgs& c, int ks, int cin, int cout, int act) {
  if (!(ks == 4 && cin == 8 && (cout == 1 || cout == 2) && c.A >= 1 && c.A <= ARU_COMBINE_MAX && c.geo.N <= 65535 &&
        act >= 0 && act <= 3))
    return false;
  const HeadPlan p = head_plan(c);
  return p.ok && cdiv(c.geo.H, p.TH) <= 65535;
}

// TF filter [4][4][8][cout] (float32) -> B fragments: register (ky, kx pair, half) of lane (g8, t4) =
// { W[ky][2 kxp + half][2 t4][g8], W[ky][2 kxp + half][2 t4 + 1][g8] }, zero for g8 >= cout
void combine_head_pack(const float* w_tf, int cout, uint32_t* dst) {
  for (int lane = 0; lane < 32; ++lane) {
    const int g8 = lane >> 2, t4 = lane & 3;
    for (int ky = 0; ky < 4; ++ky)
      for (int kxp = 0; kxp < 2; ++kxp)
        for (int half = 0; half < 2; ++half) {
          uint32_t v = 0;
          for (int h = 0; h < 2; ++h) {
            const int ci = 2 * t4 + h, kx = 2 * kxp + half;
            const float f = g8 < cout ? w_tf[((size_t)(ky * 4 + kx) * 8 + ci) * cout + g8] : 0.f;
            v |= (uint32_t)host_f_to_act(f) << (16 * h);
          }
          dst[(size_t)((ky * 2 + kxp) * 2 + half) * 32 + lane] = v;
        }
  }
}

cudaError_t launch_combine_head(cudaStream_t st, const CombineArgs& c_in, const uint32_t* wfrag, const float* bias_host,
                                int cout, int act, float* out, int* err_flag) {
  CombineHeadArgs a{};
  a.c = c_in;
  auto log2_or_neg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  for (int k = 0; k < a.c.A; ++k) {
    a.c.att_sh[k] = log2_or_neg(a.c.att_up[k]);
    a.c.det_sh[k] = log2_or_neg(a.c.det_up[k]);
  }
  const HeadPlan p = head_plan(a.c);
  if (!p.ok) return cudaErrorInvalidValue;
  a.kf = p.kf;
  a.TH = p.TH;
  for (int k = 0; k < a.c.A; ++k) {
    a.tab_att[k] = p.tab_att[k]; a.tw_att[k] = p.tw_att[k];
    a.tab_det[k] = p.tab_det[k]; a.tw_det[k] = p.tw_det[k];
  }
  a.tab_rep = p.tab_rep; a.tab_row = p.tab_row; a.tab_wt = p.tab_wt;
  for (int k = 0, i = 0; k < a.c.A; ++k)
    if (k != p.kf) a.low_k[i++] = k;
  a.wfrag = wfrag;
  for (int i = 0; i < 8; ++i) a.bias[i] = i < cout ? bias_host[i] : 0.f;
  a.out = out;
  a.C = cout;
  a.act = act;
  a.err_flag = err_flag;
  const dim3 grid((unsigned)cdiv(a.c.geo.W, CH_TW), (unsigned)cdiv(a.c.geo.H, p.TH), (unsigned)a.c.geo.N);
  void (*k)(CombineHeadArgs) = nullptr;
  switch (a.c.A) {
    case 1: k = k_combine_head<1>; break;
    case 2: k = k_combine_head<2>; break;
    case 3: k = k_combine_head<3>; break;
    case 4: k = k_combine_head<4>; break;
    case 5: k = k_combine_head<5>; break;
    case 6: k = k_combine_head<6>; break;
    case 7: k = k_combine_head<7>; break;
    case 8: k = k_combine_head<8>; break;
    default: return cudaErrorInvalidValue;
  }
  // three CTAs per SM need the whole shared-memory carve-out (the default leaves room for ONE CTA of this size)
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  if (getenv("ARU_HEAD_DBG")) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, CH_THREADS, p.smem);
    fprintf(stderr, "combine_head: TH %d, %zu B shared memory, %d CTAs per SM, grid %u x %u x %u\n", p.TH, p.smem, nb, grid.x,
            grid.y, grid.z);
  }
  k<<<grid, CH_THREADS, p.smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace aru
