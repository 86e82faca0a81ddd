// kernels_simple.cu - CUDA-core kernels of the ARU-Net forward pass.
//
// Two groups:
//  (1) the memory-bound ops that stay on CUDA cores by design (HBM roofline): 1-channel stem convs,
//      small-C_out convs (attention logit, classifier + softmax), pools, the fused
//      upsample+softmax-over-scales+weighted-sum ("combine"), uint8/mask quantisation;
//  (2) a direct convolution / transposed convolution on CUDA cores that covers every channel count.
//      It is the validation twin of the tcgen05 implicit-GEMM kernel (conv_tc.cu): tests run both
//      against the CPU oracle; ARU_OPT_CONV_PATH=1 forces it.
// Semantics follow TF's SAME rule; reference call sites are cited per kernel.  Thanks to the zero
// frame of the activation layout (aru_common.cuh) the chunk-planar kernels need no bounds checks on
// their reads; they only ever store to in-image positions.
#include "aru_common.cuh"
#include <cstdlib>

#include "kernels.h"

namespace aru {

static __device__ __forceinline__ float apply_act(float v, int act) { return act == 1 ? fmaxf(v, 0.f) : v; }

// pixel index -> (n, y, x)
static __device__ __forceinline__ void pix_coords(long long i, int H, int W, int& n, int& y, int& x) {
  x = (int)(i % W);
  const long long r = i / W;
  y = (int)(r % H);
  n = (int)(r / H);
}

// ---------------------------------------------------------------------------------------------
// Stem conv: C_in = 1 float32 plane -> C_out <= 16 chunk-planar.  layers.py:191-247 with the 1-channel
// inputs of ARU_v1.py:173 (attention conv1, 4x4) and :212 (unet_down_0/conv1, 3x3).  HBM-bound by bytes
// (4 B in, 16-32 B out per pixel), so the work is to keep the FMA count off the critical path: the
// filter lives in the kernel parameters (constant bank), every FFMA takes its weight as a constant
// operand (no shared-memory or register traffic for weights), the loops are fully unrolled over the
// real C_out.
// One thread = one image column, walking 16 rows with a sliding KS x KS window in registers (KS coalesced
// loads per pixel); grid = (ceil(W/128), ceil(H/16), N): no index divisions.
// ---------------------------------------------------------------------------------------------
struct StemW {
  float w[16 * 16];  // [tap][16]
  float b[16];
};

constexpr int STEM_ROWS = 16;  // image rows one thread walks down (sliding KS x KS window in registers)

template <int KS, int COUT>
__global__ void __launch_bounds__(128) k_conv_stem(const float* __restrict__ in, act_t* __restrict__ out,
                                                   long long out_plane, act_t* __restrict__ out_pre,
                                                   long long pre_plane, const __grid_constant__ StemW sw, Geo g,
                                                   int act) {
  constexpr int CH = (COUT + 7) / 8, PB = (KS - 1) / 2;
  const int n = blockIdx.z, y0 = blockIdx.y * STEM_ROWS, x = blockIdx.x * 128 + threadIdx.x;
  if (x >= g.W) return;
  const float* img = in + (long long)n * g.H * g.W;
  bool okx[KS];
#pragma unroll
  for (int kx = 0; kx < KS; ++kx) okx[kx] = (unsigned)(x + kx - PB) < (unsigned)g.W;
  auto load_row = [&](int yy, float r[KS]) {
    const bool oky = (unsigned)yy < (unsigned)g.H;
    const float* row = img + (long long)yy * g.W + (x - PB);
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) r[kx] = (oky && okx[kx]) ? __ldg(row + kx) : 0.f;
  };
  float v[KS][KS];
#pragma unroll
  for (int ky = 0; ky < KS - 1; ++ky) load_row(y0 + ky - PB, v[ky + 1]);   // rows ky = 1..KS-1 of the first window, shifted below
  const int y1 = min(y0 + STEM_ROWS, g.H);
  long long p = g.pos(n, y0, x);
  for (int y = y0; y < y1; ++y, p += g.Wp) {
#pragma unroll
    for (int ky = 0; ky < KS - 1; ++ky)
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) v[ky][kx] = v[ky + 1][kx];
    load_row(y + KS - 1 - PB, v[KS - 1]);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = (c * 8 + j < COUT) ? sw.b[c * 8 + j] : 0.f;
#pragma unroll
      for (int t = 0; t < KS * KS; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c * 8 + j < COUT) acc[j] = fmaf(v[t / KS][t % KS], sw.w[t * 16 + c * 8 + j], acc[j]);
      if (out_pre) *reinterpret_cast<uint4*>(out_pre + (c * pre_plane + p) * 8) = pack8_fin<false>(acc, true);
      if (out)   // null: only the pre-activation is wanted (the fused residual block starts from it, block_mma.cu)
        *reinterpret_cast<uint4*>(out + (c * out_plane + p) * 8) =
            act == 1 ? pack8_fin<true>(acc, true) : pack8_fin<false>(acc, true);
    }
  }
}

// Stem conv + ReLU + 2x2 stride-2 SAME max-pool in one launch (ARU_v1.py:173-176: the attention CNN's conv1 feeds
// nothing but its pool).  Same walk as k_conv_stem; the packed 16-bit results of an even row are kept in registers, the
// odd row is max-ed in (packed hmax2; the maximum commutes with the rounding, so this equals pooling the stored
// 16-bit tensor), one shfl.xor brings the neighbouring column and even lanes store the pooled pixel.  Post-ReLU values
// are >= 0, so columns / rows beyond the image contribute zeros (TF pads the pool with -inf: same maximum).
// FULL: also store the full-resolution tensor (needed only when something other than the pool reads it).
static __device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
#ifdef ARU_USE_BF16
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
#else
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
#endif
  return *reinterpret_cast<uint32_t*>(&r);
}
static __device__ __forceinline__ uint4 hmax2_u4(uint4 a, uint4 b) {
  return make_uint4(hmax2_u32(a.x, b.x), hmax2_u32(a.y, b.y), hmax2_u32(a.z, b.z), hmax2_u32(a.w, b.w));
}

template <int KS, int COUT, bool FULL>
__global__ void __launch_bounds__(128) k_conv_stem_pool(const float* __restrict__ in, act_t* __restrict__ out,
                                                        long long out_plane, act_t* __restrict__ pool,
                                                        long long pool_plane, const __grid_constant__ StemW sw, Geo g,
                                                        Geo gp) {
  constexpr int CH = (COUT + 7) / 8, PB = (KS - 1) / 2;
  const int n = blockIdx.z, y0 = blockIdx.y * STEM_ROWS, x = blockIdx.x * 128 + threadIdx.x;
  const bool live = x < g.W;   // dead lanes stay for the shuffles and contribute zeros
  const float* img = in + (long long)n * g.H * g.W;
  bool okx[KS];
#pragma unroll
  for (int kx = 0; kx < KS; ++kx) okx[kx] = (unsigned)(x + kx - PB) < (unsigned)g.W;
  auto load_row = [&](int yy, float r[KS]) {
    const bool oky = (unsigned)yy < (unsigned)g.H;
    const float* row = img + (long long)yy * g.W + (x - PB);
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) r[kx] = (oky && okx[kx]) ? __ldg(row + kx) : 0.f;
  };
  float v[KS][KS];
#pragma unroll
  for (int ky = 0; ky < KS - 1; ++ky) load_row(y0 + ky - PB, v[ky + 1]);
  const int y1 = min(y0 + STEM_ROWS, g.H);
  long long p = g.pos(n, y0, x);
  uint4 even[CH];
  for (int y = y0; y < y1; ++y, p += g.Wp) {
#pragma unroll
    for (int ky = 0; ky < KS - 1; ++ky)
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) v[ky][kx] = v[ky + 1][kx];
    load_row(y + KS - 1 - PB, v[KS - 1]);
    const bool odd = y & 1, last = y + 1 == g.H;   // STEM_ROWS is even: row pairs never straddle two blocks
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = (c * 8 + j < COUT) ? sw.b[c * 8 + j] : 0.f;
#pragma unroll
      for (int t = 0; t < KS * KS; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c * 8 + j < COUT) acc[j] = fmaf(v[t / KS][t % KS], sw.w[t * 16 + c * 8 + j], acc[j]);
      const uint4 cur = pack8_fin<true>(acc, live);
      if (FULL && live) *reinterpret_cast<uint4*>(out + (c * out_plane + p) * 8) = cur;
      if (!odd && !last) {
        even[c] = cur;
      } else {
        uint4 m = odd ? hmax2_u4(even[c], cur) : cur;
        uint4 nb;
        nb.x = __shfl_xor_sync(0xffffffffu, m.x, 1);
        nb.y = __shfl_xor_sync(0xffffffffu, m.y, 1);
        nb.z = __shfl_xor_sync(0xffffffffu, m.z, 1);
        nb.w = __shfl_xor_sync(0xffffffffu, m.w, 1);
        m = hmax2_u4(m, nb);
        if (live && !(x & 1))
          *reinterpret_cast<uint4*>(pool + (c * pool_plane + gp.pos(n, y >> 1, x >> 1)) * 8) = m;
      }
    }
  }
}

// 4x4 stem conv + ReLU + 2x2 max-pool on the tensor cores (warp-level mma.sync: the 16 taps of the 4x4 window are
// exactly one K = 16 step).  The FFMA version above is CUDA-core bound (192 FFMA per pixel, measured 43 % of the fp32
// peak and 3x the HBM time of the pooled output); here a warp turns 32 pixels of a row into 2 (pixel tiles of 16) x NT
// (channel tiles of 8) m16n8k16 MMAs:
//   A[pixel][tap]   gathered as 32-bit (two adjacent kx) loads from a 16-bit shared-memory copy of the (16+3) x (128+3)
//                   input tile - stored twice, the second copy shifted by one element, so that the pair (c, c+1) is
//                   4-byte aligned in one of them whatever the parity of c;
//   B[tap][channel] the filter, four 32-bit registers per thread for the whole kernel;
//   D[pixel][ch]    lane (g, t) holds channels 2t, 2t+1 of pixels g, g+8: bias + ReLU + 16-bit pack per lane, the row pair
//                   is max-ed in registers, the column pair comes from lane ^ 4, and a warp store writes 4-byte pieces
//                   that tile 16-byte position vectors contiguously.
// Operands are rounded to the 16-bit storage type like every other layer's (fp32 accumulate).
static __device__ __forceinline__ void mma_16816(float d[4], const uint32_t a[4], const uint32_t b[2], const float c[2]) {
  // D = A B + C with C = (c0, c1, c0, c1): the bias of this lane's two channels, never overwritten
#ifdef ARU_USE_BF16
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};"
#else
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};"
#endif
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(c[0]), "f"(c[1]));
}
// two fp32 -> packed 16-bit pair, plain round-to-nearest (an overflow becomes +-inf and is clamped by finish2)
static __device__ __forceinline__ uint32_t pack2_raw(float x, float y) {
  act2_t h;
#ifdef ARU_USE_BF16
  h = __floats2bfloat162_rn(x, y);
#else
  h = __floats2half2_rn(x, y);
#endif
  return *reinterpret_cast<uint32_t*>(&h);
}
// saturate to the finite range and ReLU a packed pair (both commute with the maximum of the pooling window)
static __device__ __forceinline__ uint32_t finish2_relu(uint32_t v) {
#ifdef ARU_USE_BF16
  return hmax2_u32(v, 0u);
#else
  __half2 h = __hmin2(*reinterpret_cast<__half2*>(&v), __half2half2(__ushort_as_half((unsigned short)0x7bff)));
  h = __hmax2(h, __half2half2(__ushort_as_half((unsigned short)0)));
  return *reinterpret_cast<uint32_t*>(&h);
#endif
}
static __device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Tile rows are 144 halves (72 words: the two tap rows a warp reads land 8 banks apart) and the second copy starts 16
// banks after the first, so the 32-bit fragment loads of a warp are conflict free.
constexpr int SM_TILE_W = 128, SM_ROWS = STEM_ROWS + 3, SM_STRIDE = 144, SM_COPY = SM_ROWS * SM_STRIDE + 48;

template <int NT, bool FULL>
__global__ void __launch_bounds__(128) k_conv_stem4_pool_mma(const float* __restrict__ in, act_t* __restrict__ out,
                                                             long long out_plane, act_t* __restrict__ pool,
                                                             long long pool_plane, const float* __restrict__ w_dev,
                                                             const float* __restrict__ bias_dev, Geo g, Geo gp) {
  // w_dev: filter [16 taps][16] (float32, device memory; per-lane fragment gathers from the constant bank would replay
  // once per distinct address), bias_dev [16]
  __shared__ __align__(16) act_t tile_mem[2 * SM_COPY];
  auto tile = [&](int copy, int r, int c) -> act_t& { return tile_mem[copy * SM_COPY + r * SM_STRIDE + c]; };
  const int n = blockIdx.z, y0 = blockIdx.y * STEM_ROWS, x0 = blockIdx.x * SM_TILE_W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const float* img = in + (long long)n * g.H * g.W;
  // tile[.][r][c] = pixel (y0 + r - 1, x0 + c - 1); copy 1 holds element c at index c + 1.  All loads of a thread are
  // issued before the first conversion (one exposed memory latency per block instead of one per tile row).
  {
    float v[SM_ROWS];
    const int gx = x0 + tid - 1;
    const bool okx = (unsigned)gx < (unsigned)g.W;
#pragma unroll
    for (int r = 0; r < SM_ROWS; ++r) {
      const int gy = y0 + r - 1;
      v[r] = (okx && (unsigned)gy < (unsigned)g.H) ? __ldg(img + (long long)gy * g.W + gx) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < SM_ROWS; ++r) {
      const act_t h = f_to_act(v[r]);
      tile(0, r, tid) = h;
      tile(1, r, tid + 1) = h;
    }
    if (tid < 3 * SM_ROWS) {   // the three extra columns 128..130
      const int r = tid / 3, c = SM_TILE_W + tid % 3;
      const int gy = y0 + r - 1, gx2 = x0 + c - 1;
      const float u = ((unsigned)gy < (unsigned)g.H && (unsigned)gx2 < (unsigned)g.W) ? __ldg(img + (long long)gy * g.W + gx2) : 0.f;
      const act_t h = f_to_act(u);
      tile(0, r, c) = h;
      tile(1, r, c + 1) = h;
    }
  }
  // filter fragments: b[nt][0] = taps (2t, 2t+1), b[nt][1] = taps (2t+8, 2t+9) of channel 8 nt + g
  uint32_t bw[NT][2];
  float bias[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int ch = nt * 8 + gq;
    bw[nt][0] = pack2(__ldg(w_dev + (2 * t) * 16 + ch), __ldg(w_dev + (2 * t + 1) * 16 + ch));
    bw[nt][1] = pack2(__ldg(w_dev + (2 * t + 8) * 16 + ch), __ldg(w_dev + (2 * t + 9) * 16 + ch));
    bias[nt][0] = __ldg(bias_dev + nt * 8 + 2 * t);
    bias[nt][1] = __ldg(bias_dev + nt * 8 + 2 * t + 1);
  }
  __syncthreads();
  if (y0 >= g.H) return;
  const int ky0 = t >> 1, kx0 = 2 * (t & 1);
  const int copy = gq & 1;                           // parity of the tile column of this lane's pairs
  const int pl0 = warp * 32 + gq;                    // tile-local pixel column of accumulator row g of pixel tile 0
  // byte address of A register 0 for tile row 0; + 32 B per pixel tile, + 16 B for rows g + 8, + 2 tile rows for taps 8..15
  uint32_t ta = (uint32_t)__cvta_generic_to_shared(&tile(copy, ky0, pl0 + kx0 + copy));
  constexpr uint32_t ROWB = SM_STRIDE * 2;
  const int xa0 = x0 + pl0;
  bool live[2][2], mate[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int x = xa0 + mt * 16 + h * 8;
      live[mt][h] = x < g.W;
      mate[mt][h] = (x ^ 1) < g.W;                    // the other column of the 2x2 window exists
    }
  act_t* pp = pool + gp.pos(n, y0 >> 1, xa0 >> 1) * 8 + 2 * t;   // + 64 halves per pixel tile, + 32 for rows g + 8
  act_t* po = FULL ? out + g.pos(n, y0, xa0) * 8 + 2 * t : nullptr;
  const int rows = min(STEM_ROWS, g.H - y0);
  for (int r = 0; r < rows; r += 2, ta += 2 * ROWB, pp += (long long)gp.Wp * 8) {
    const bool two = r + 1 < rows;
    uint32_t m[2][NT][2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      if (rr == 1 && !two) break;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t base = ta + rr * ROWB + mt * 32;
        uint32_t a[4];
        a[0] = lds32(base);
        a[1] = lds32(base + 16);
        a[2] = lds32(base + 2 * ROWB);
        a[3] = lds32(base + 2 * ROWB + 16);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          float d[4];
          mma_16816(d, a, bw[nt], bias[nt]);
          uint32_t c0 = pack2_raw(d[0], d[1]), c1 = pack2_raw(d[2], d[3]);
          if (FULL) {
            c0 = finish2_relu(c0);
            c1 = finish2_relu(c1);
            act_t* o = po + ((long long)nt * out_plane + (long long)rr * g.Wp + mt * 16) * 8;
            if (live[mt][0]) *reinterpret_cast<uint32_t*>(o) = c0;
            if (live[mt][1]) *reinterpret_cast<uint32_t*>(o + 64) = c1;
          }
          if (rr == 0) {
            m[mt][nt][0] = c0;
            m[mt][nt][1] = c1;
          } else {
            m[mt][nt][0] = hmax2_u32(m[mt][nt][0], c0);
            m[mt][nt][1] = hmax2_u32(m[mt][nt][1], c1);
          }
        }
      }
    }
    if (FULL) po += 2LL * g.Wp * 8;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v = m[mt][nt][h];
          const uint32_t nb = __shfl_xor_sync(0xffffffffu, v, 4);   // the other column of the 2x2 window: g ^ 1
          if (mate[mt][h]) v = hmax2_u32(v, nb);
          if (!FULL) v = finish2_relu(v);
          if (!(gq & 1) && live[mt][h])
            *reinterpret_cast<uint32_t*>(pp + (long long)nt * pool_plane * 8 + mt * 64 + h * 32) = v;
        }
  }
}

// The 3x3 stem of a fused residual block (block_mma.cu starts from the stem's PRE-ACTIVATION): 1 -> 8 channels, only the
// un-activated result is stored.  Same scheme as k_conv_stem4_pool_mma - the 3x3 SAME window (one row / column before,
// one after) is the 4x4 window of that kernel with a zero fourth row and column, so the tile, its two copies and the
// fragment addressing are identical and the 9 taps are one K = 16 step of mma.sync.m16n8k16: 1 HMMA per 16 pixels
// instead of 72 FFMA per pixel (the CUDA-core kernel ran at 0.64 of the HBM roofline, FMA bound).
__global__ void __launch_bounds__(128) k_conv_stem3_pre_mma(const float* __restrict__ in, act_t* __restrict__ out,
                                                            act_t* __restrict__ out_pre, const float* __restrict__ w_dev,
                                                            const float* __restrict__ bias_dev, Geo g, int act) {
  // out (may be null): the activated result; out_pre (may be null): the un-activated one
  // w_dev: filter [9 taps][16] (float32, device memory), bias_dev [16]
  __shared__ __align__(16) act_t tile_mem[2 * SM_COPY];
  auto tile = [&](int copy, int r, int c) -> act_t& { return tile_mem[copy * SM_COPY + r * SM_STRIDE + c]; };
  const int n = blockIdx.z, y0 = blockIdx.y * STEM_ROWS, x0 = blockIdx.x * SM_TILE_W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const float* img = in + (long long)n * g.H * g.W;
  {   // tile[.][r][c] = pixel (y0 + r - 1, x0 + c - 1); copy 1 holds element c at index c + 1
    float v[SM_ROWS];
    const int gx = x0 + tid - 1;
    const bool okx = (unsigned)gx < (unsigned)g.W;
#pragma unroll
    for (int r = 0; r < SM_ROWS; ++r) {
      const int gy = y0 + r - 1;
      v[r] = (okx && (unsigned)gy < (unsigned)g.H) ? __ldg(img + (long long)gy * g.W + gx) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < SM_ROWS; ++r) {
      const act_t h = f_to_act(v[r]);
      tile(0, r, tid) = h;
      tile(1, r, tid + 1) = h;
    }
    if (tid < 3 * SM_ROWS) {   // the three extra columns 128..130
      const int r = tid / 3, c = SM_TILE_W + tid % 3;
      const int gy = y0 + r - 1, gx2 = x0 + c - 1;
      const float u = ((unsigned)gy < (unsigned)g.H && (unsigned)gx2 < (unsigned)g.W) ? __ldg(img + (long long)gy * g.W + gx2) : 0.f;
      const act_t h = f_to_act(u);
      tile(0, r, c) = h;
      tile(1, r, c + 1) = h;
    }
  }
  // filter fragments over the 4 x 4 window index k = 4 ky + kx (zero for ky = 3 or kx = 3):
  // b0 = taps k = 2t, 2t + 1, b1 = taps k + 8 of channel g
  auto tap = [&](int k) {
    const int ky = k >> 2, kx = k & 3;
    return (ky < 3 && kx < 3) ? __ldg(w_dev + (ky * 3 + kx) * 16 + gq) : 0.f;
  };
  uint32_t bw[2];
  bw[0] = pack2(tap(2 * t), tap(2 * t + 1));
  bw[1] = pack2(tap(2 * t + 8), tap(2 * t + 9));
  float bias[2] = {__ldg(bias_dev + 2 * t), __ldg(bias_dev + 2 * t + 1)};
  __syncthreads();
  if (y0 >= g.H) return;
  const int ky0 = t >> 1, kx0 = 2 * (t & 1);
  const int copy = gq & 1;
  const int pl0 = warp * 32 + gq;
  uint32_t ta = (uint32_t)__cvta_generic_to_shared(&tile(copy, ky0, pl0 + kx0 + copy));
  constexpr uint32_t ROWB = SM_STRIDE * 2;
  const int xa0 = x0 + pl0;
  bool live[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) live[mt][h] = xa0 + mt * 16 + h * 8 < g.W;
  const long long p0 = g.pos(n, y0, xa0) * 8 + 2 * t;
  const int rows = min(STEM_ROWS, g.H - y0);
  long long po = p0;
  for (int r = 0; r < rows; ++r, ta += ROWB, po += (long long)g.Wp * 8) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const uint32_t base = ta + mt * 32;
      uint32_t a[4];
      a[0] = lds32(base);
      a[1] = lds32(base + 16);
      a[2] = lds32(base + 2 * ROWB);
      a[3] = lds32(base + 2 * ROWB + 16);
      float d[4];
      mma_16816(d, a, bw, bias);
      if (out_pre) {
        if (live[mt][0]) *reinterpret_cast<uint32_t*>(out_pre + po + mt * 128) = pack2_fin<false>(d[0], d[1]);
        if (live[mt][1]) *reinterpret_cast<uint32_t*>(out_pre + po + mt * 128 + 64) = pack2_fin<false>(d[2], d[3]);
      }
      if (out) {
        if (live[mt][0]) *reinterpret_cast<uint32_t*>(out + po + mt * 128) = act == 1 ? pack2_fin<true>(d[0], d[1]) : pack2_fin<false>(d[0], d[1]);
        if (live[mt][1]) *reinterpret_cast<uint32_t*>(out + po + mt * 128 + 64) = act == 1 ? pack2_fin<true>(d[2], d[3]) : pack2_fin<false>(d[2], d[3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Direct conv, chunk-planar -> chunk-planar (validation twin of conv_tc).  One thread = one pixel x
// one output chunk.  Weights packed [tap][cin_chunk][cout_chunk][ci 8][co 8] (16-bit).
// out = act(conv + bias [+ res]);  out_pre = conv + bias [+ res] (ARU_v1.py:212-227).
// ---------------------------------------------------------------------------------------------
static __device__ __forceinline__ void mac8x8(const uint4& xin, const uint4* __restrict__ wp, float acc[8]) {
  float xi[8];
  unpack8(xin, xi);
#pragma unroll
  for (int ci = 0; ci < 8; ++ci) {
    float wf[8];
    unpack8(__ldg(wp + ci), wf);
#pragma unroll
    for (int co = 0; co < 8; ++co) acc[co] = fmaf(xi[ci], wf[co], acc[co]);
  }
}

template <int KS>
__global__ void __launch_bounds__(256) k_conv_direct(const act_t* __restrict__ in, long long in_plane, int cin_chunks,
                                                     act_t* __restrict__ out, act_t* __restrict__ out_pre,
                                                     const act_t* __restrict__ res, long long out_plane,
                                                     long long pre_plane, long long res_plane, int cout_chunks,
                                                     const act_t* __restrict__ w, const float* __restrict__ bias,
                                                     Geo g, int act) {
  const long long total = (long long)g.N * g.H * g.W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int coc = blockIdx.y;
  if (i >= total) return;
  int n, y, x;
  pix_coords(i, g.H, g.W, n, y, x);
  const long long p = g.pos(n, y, x);
  constexpr int PB = (KS - 1) / 2;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = __ldg(bias + coc * 8 + j);
  for (int ky = 0; ky < KS; ++ky) {
    for (int kx = 0; kx < KS; ++kx) {
      const long long q = p + (long long)(ky - PB) * g.Wp + (kx - PB);  // frame positions read as zero
      const int tap = ky * KS + kx;
      for (int cic = 0; cic < cin_chunks; ++cic) {
        const uint4 xin = __ldg(reinterpret_cast<const uint4*>(in + (cic * in_plane + q) * 8));
        const uint4* wp = reinterpret_cast<const uint4*>(w) + ((long long)(tap * cin_chunks + cic) * cout_chunks + coc) * 8;
        mac8x8(xin, wp, acc);
      }
    }
  }
  if (res) {
    float r[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(res + (coc * res_plane + p) * 8)), r);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += r[j];
  }
  if (out_pre) *reinterpret_cast<uint4*>(out_pre + (coc * pre_plane + p) * 8) = pack8(acc);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = apply_act(acc[j], act);
  *reinterpret_cast<uint4*>(out + (coc * out_plane + p) * 8) = pack8(acc);
}

// ---------------------------------------------------------------------------------------------
// Small-C_out conv to float32: attention logit (4x4, 32->1, ReLU; ARU_v1.py:182-183) and the classifier
// (4x4, 8->n_class, identity + output Softmax/Sigmoid; ARU_v1.py:158-160 + exporter's 'output' op,
// contract helper.py:70).  Output is dense float32 NHWC [N,H,W,cout] (cout == 1 -> a plane).
// Weights float32 [tap][cin_pad][cout].
// ---------------------------------------------------------------------------------------------
template <int KS, int COUT>
__global__ void __launch_bounds__(128) k_conv_small(const act_t* __restrict__ in, long long in_plane, int cin_chunks,
                                                    float* __restrict__ out, const float* __restrict__ w,
                                                    const float* __restrict__ bias, Geo g, int act) {
  extern __shared__ float sw[];  // [KS*KS][cin_chunks*8][COUT]
  const int nw = KS * KS * cin_chunks * 8 * COUT;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long total = (long long)g.N * g.H * g.W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int n, y, x;
  pix_coords(i, g.H, g.W, n, y, x);
  const long long p = g.pos(n, y, x);
  constexpr int PB = (KS - 1) / 2;
  float acc[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) acc[j] = __ldg(bias + j);
  for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) {
      const long long q = p + (long long)(ky - PB) * g.Wp + (kx - PB);
      const float* wt = sw + (ky * KS + kx) * cin_chunks * 8 * COUT;
      for (int cic = 0; cic < cin_chunks; ++cic) {
        float xi[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(in + (cic * in_plane + q) * 8)), xi);
#pragma unroll
        for (int ci = 0; ci < 8; ++ci)
#pragma unroll
          for (int j = 0; j < COUT; ++j) acc[j] = fmaf(xi[ci], wt[(cic * 8 + ci) * COUT + j], acc[j]);
      }
    }
  }
  if (act == 1) {
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = fmaxf(acc[j], 0.f);
  } else if (act == 2) {  // softmax over channels, fp32 (layers.py:48-49)
    float m = acc[0];
#pragma unroll
    for (int j = 1; j < COUT; ++j) m = fmaxf(m, acc[j]);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
      acc[j] = expf(acc[j] - m);
      s += acc[j];
    }
    const float inv = 1.f / s;
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] *= inv;
  } else if (act == 3) {
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = 1.f / (1.f + expf(-acc[j]));
  }
#pragma unroll
  for (int j = 0; j < COUT; ++j) out[i * COUT + j] = acc[j];
}

// ---------------------------------------------------------------------------------------------
// Transposed conv 3x3 stride 2 SAME (layers.py:342-367: tf.nn.conv2d_transpose, filter [k,k,Cout,Cin]).
// full[2*iy+ky, 2*ix+kx] += in[iy,ix] * W[ky,kx];  out = full[oy:oy+Ho, ox:ox+Wo],
// oy = (2*Hi + 1 - Ho) / 2  (0 when Ho even, 1 when odd).  Weights packed like k_conv_direct.
// Input reads at iy = -1 / Hi or ix = -1 / Wi land on the zero frame.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_deconv_direct(const act_t* __restrict__ in, long long in_plane, int cin_chunks,
                                                       act_t* __restrict__ out, long long out_plane, int cout_chunks,
                                                       const act_t* __restrict__ w, const float* __restrict__ bias,
                                                       Geo gi, Geo go, int act) {
  const long long total = (long long)go.N * go.H * go.W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int coc = blockIdx.y;
  if (i >= total) return;
  int n, y, x;
  pix_coords(i, go.H, go.W, n, y, x);
  const int oy = (2 * gi.H + 1 - go.H) / 2, ox = (2 * gi.W + 1 - go.W) / 2;
  const int fy = y + oy, fx = x + ox;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = __ldg(bias + coc * 8 + j);
  for (int ky = (fy & 1); ky < 3; ky += 2) {
    const int iy = (fy - ky) >> 1;  // in [-1, Hi]
    for (int kx = (fx & 1); kx < 3; kx += 2) {
      const int ix = (fx - kx) >> 1;
      const long long q = gi.pos(n, iy, ix);
      const int tap = ky * 3 + kx;
      for (int cic = 0; cic < cin_chunks; ++cic) {
        const uint4 xin = __ldg(reinterpret_cast<const uint4*>(in + (cic * in_plane + q) * 8));
        const uint4* wp = reinterpret_cast<const uint4*>(w) + ((long long)(tap * cin_chunks + cic) * cout_chunks + coc) * 8;
        mac8x8(xin, wp, acc);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = apply_act(acc[j], act);
  *reinterpret_cast<uint4*>(out + (coc * out_plane + go.pos(n, y, x)) * 8) = pack8(acc);
}

// ---------------------------------------------------------------------------------------------
// 2x2 stride-2 SAME pools (layers.py:526-544).  out = ceil(in/2); the extra row/col of odd sizes sees
// -inf (max) or is left out of the divisor (avg).
// ---------------------------------------------------------------------------------------------
template <bool IS_MAX>
__global__ void __launch_bounds__(256) k_pool(const act_t* __restrict__ in, long long in_plane, act_t* __restrict__ out,
                                              long long out_plane, Geo gi, Geo go) {
  const long long total = (long long)go.N * go.H * go.W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (i >= total) return;
  int n, y, x;
  pix_coords(i, go.H, go.W, n, y, x);
  const act_t* src = in + (c * in_plane + gi.pos(n, 2 * y, 2 * x)) * 8;
  const bool hx = 2 * x + 1 < gi.W, hy = 2 * y + 1 < gi.H;
  float m[8], t[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(src)), m);
  int cnt = 1;
  if (hx) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + 8)), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = IS_MAX ? fmaxf(m[j], t[j]) : m[j] + t[j];
    ++cnt;
  }
  if (hy) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + (long long)gi.Wp * 8)), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = IS_MAX ? fmaxf(m[j], t[j]) : m[j] + t[j];
    ++cnt;
    if (hx) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + (long long)gi.Wp * 8 + 8)), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = IS_MAX ? fmaxf(m[j], t[j]) : m[j] + t[j];
      ++cnt;
    }
  }
  if (!IS_MAX) {
    const float inv = 1.f / cnt;
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] *= inv;
  }
  *reinterpret_cast<uint4*>(out + (c * out_plane + go.pos(n, y, x)) * 8) = pack8(m);
}

// float32 1-channel planes (input pyramid, ARU_v1.py:105-109)
__global__ void __launch_bounds__(256) k_pool_f32(const float* __restrict__ in, float* __restrict__ out, int N, int Hi,
                                                  int Wi, int Ho, int Wo, int is_max) {
  const long long total = (long long)N * Ho * Wo;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int x = (int)(p % Wo);
  const int y = (int)((p / Wo) % Ho);
  const int n = (int)(p / ((long long)Wo * Ho));
  const float* src = in + ((long long)n * Hi + 2 * y) * Wi + 2 * x;
  const bool hx = 2 * x + 1 < Wi, hy = 2 * y + 1 < Hi;
  float a = src[0];
  if (is_max) {
    if (hx) a = fmaxf(a, src[1]);
    if (hy) a = fmaxf(a, src[Wi]);
    if (hx && hy) a = fmaxf(a, src[Wi + 1]);
  } else {
    // the divisor counts the valid cells only: 1, 2 or 4 - a multiplication by its reciprocal is exact
    float inv = 1.f;
    if (hx) { a += src[1]; inv = 0.5f; }
    if (hy) { a += src[Wi]; inv *= 0.5f; }
    if (hx && hy) a += src[Wi + 1];
    a *= inv;
  }
  out[p] = a;
}

// widths that are multiples of 4: two outputs per thread from one 16-byte load per input row (same sums in the same
// order as k_pool_f32: (a + b) + c + d with a, b of the upper row)
__global__ void __launch_bounds__(256) k_pool_f32_x2(const float4* __restrict__ in, float2* __restrict__ out, int N, int Hi,
                                                     int Wi4, int Ho, int Wo2, int is_max) {
  const long long total = (long long)N * Ho * Wo2;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int x = (int)(p % Wo2);
  const int y = (int)((p / Wo2) % Ho);
  const int n = (int)(p / ((long long)Wo2 * Ho));
  const float4* src = in + ((long long)n * Hi + 2 * y) * Wi4 + x;
  const bool hy = 2 * y + 1 < Hi;
  const float4 u = __ldg(src);
  float2 r;
  if (is_max) {
    r = make_float2(fmaxf(u.x, u.y), fmaxf(u.z, u.w));
    if (hy) {
      const float4 d = __ldg(src + Wi4);
      r.x = fmaxf(fmaxf(r.x, d.x), d.y);
      r.y = fmaxf(fmaxf(r.y, d.z), d.w);
    }
  } else {
    r = make_float2(u.x + u.y, u.z + u.w);
    float inv = 0.5f;
    if (hy) {
      const float4 d = __ldg(src + Wi4);
      r.x = (r.x + d.x) + d.y;
      r.y = (r.y + d.z) + d.w;
      inv = 0.25f;
    }
    r.x *= inv;
    r.y *= inv;
  }
  out[p] = r;
}

// ---------------------------------------------------------------------------------------------
// Fused attention tail (ARU_v1.py:115,137,145-153):
//   a_k  = upsample_simple(att_k)  : NN-upsample x up_att[k] of a 1-channel float plane, cropped at
//          offset (Hin*up - H)/2 (layers.py:716-720 via conv2d_transpose SAME)
//   s    = softmax_k(a_k)          (fp32)
//   d_k  = det_k (k with up_det == 1) or upsample_simple(det_k): with the ones filter [up,up,C,C] every
//          output channel is the SUM over all input channels (the reference's quirk, kept)
//   out  = sum_k d_k * s_k
// One thread per output pixel; low-resolution operands are re-read through L1/L2.
// ---------------------------------------------------------------------------------------------
static __device__ __forceinline__ int div_up(int v, int up, int sh) { return sh >= 0 ? (v >> sh) : (v / up); }

constexpr int COMBINE_ROWS = 16;

// One thread = one output column walking 16 rows; grid = (ceil(W/128), ceil(H/16), N).  Column-wise source indices
// are computed once per thread, row-wise ones are shifts (the upsample factors of the ARU topology are powers of two;
// other factors take the division).  Templated on the number of scales so the per-scale state lives in registers.
// FAST: one output plane and exactly one full-resolution detection map (index kf) - the ARU topology; its 16-byte
// vectors are fetched three rows ahead so that four loads per thread are in flight (the kernel is latency bound otherwise).
template <int A, bool FAST>
__global__ void __launch_bounds__(128) k_combine(const __grid_constant__ CombineArgs a, int kf) {
  const Geo& g = a.geo;
  const int n = blockIdx.z, y0 = blockIdx.y * COMBINE_ROWS, x = blockIdx.x * 128 + threadIdx.x;
  if (x >= g.W) return;
  const float* att_col[A];
  long long det_col[A];   // position of (n, row 0, source column) in the low-resolution detection map
#pragma unroll
  for (int k = 0; k < A; ++k) {
    att_col[k] = a.att[k] + (long long)n * a.att_h[k] * a.att_w[k] + div_up(x + a.att_ox[k], a.att_up[k], a.att_sh[k]);
    det_col[k] = a.det_geo[k].pos(n, 0, div_up(x + a.det_ox[k], a.det_up[k], a.det_sh[k]));
  }
  const int y1 = min(y0 + COMBINE_ROWS, g.H);
  long long p = g.pos(n, y0, x);
  // Upsampled sources change only every `up` rows: the attention logits (and with them the softmax weights) and the
  // channel sums of the low-resolution detection maps are kept in registers and refreshed when their source row moves
  // (the test is uniform over the block: all threads of a block share y).
  float raw[A], wgt[A], lowsum[A];
  int att_row[A], det_row[A];
#pragma unroll
  for (int k = 0; k < A; ++k) { att_row[k] = -1; det_row[k] = -1; raw[k] = 0.f; wgt[k] = 0.f; lowsum[k] = 0.f; }
  const uint4* full = nullptr;   // FAST: the full-resolution detection vector of (row y, this column)
  uint4 ahead[3];
  if (FAST) {
    full = reinterpret_cast<const uint4*>(a.det[kf]) + p;
#pragma unroll
    for (int j = 0; j < 3; ++j) ahead[j] = (y0 + j < y1) ? __ldg(full + (long long)j * g.Wp) : make_uint4(0u, 0u, 0u, 0u);
  }
  for (int y = y0; y < y1; ++y, p += g.Wp) {
    uint4 cur = make_uint4(0u, 0u, 0u, 0u);
    if (FAST) {
      cur = ahead[0];
      ahead[0] = ahead[1];
      ahead[1] = ahead[2];
      if (y + 3 < y1) ahead[2] = __ldg(full + 3LL * g.Wp);
      full += g.Wp;
    }
    bool moved = false;
#pragma unroll
    for (int k = 0; k < A; ++k) {
      const int r = div_up(y + a.att_oy[k], a.att_up[k], a.att_sh[k]);
      if (r != att_row[k]) {
        att_row[k] = r;
        raw[k] = __ldg(att_col[k] + (long long)r * a.att_w[k]);
        moved = true;
      }
    }
    if (moved) {
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < A; ++k) m = fmaxf(m, raw[k]);
      float den = 0.f;
#pragma unroll
      for (int k = 0; k < A; ++k) {
        wgt[k] = __expf(raw[k] - m);
        den += wgt[k];
      }
      const float inv = 1.f / den;
#pragma unroll
      for (int k = 0; k < A; ++k) wgt[k] *= inv;
    }
    // low-resolution detection maps: channel sum of the source pixel (the ones-filter quirk), shared by all chunks
    float low = 0.f;
#pragma unroll
    for (int k = 0; k < A; ++k) {
      if (a.det_up[k] != 1) {
        const int r = div_up(y + a.det_oy[k], a.det_up[k], a.det_sh[k]);
        if (r != det_row[k]) {
          det_row[k] = r;
          const long long q = det_col[k] + (long long)r * a.det_geo[k].Wp;
          float sum = 0.f;
          for (int c = 0; c < a.det_chunks[k]; ++c) {
            float d[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(a.det[k] + (c * a.det_plane[k] + q) * 8)), d);
            sum += ((d[0] + d[1]) + (d[2] + d[3])) + ((d[4] + d[5]) + (d[6] + d[7]));
          }
          lowsum[k] = sum;
        }
        low = fmaf(lowsum[k], wgt[k], low);
      }
    }
    if (FAST) {
      float acc[8], d[8];
      unpack8(cur, d);
      float wk = 0.f;
#pragma unroll
      for (int k = 0; k < A; ++k)
        if (k == kf) wk = wgt[k];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(d[j], wk, low);
      *reinterpret_cast<uint4*>(a.out + p * 8) = pack8_fin<false>(acc, true);
      continue;
    }
    for (int oc = 0; oc < a.out_chunks; ++oc) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = low;
#pragma unroll
      for (int k = 0; k < A; ++k) {
        if (a.det_up[k] == 1) {
          float d[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(a.det[k] + (oc * a.det_plane[k] + p) * 8)), d);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(d[j], wgt[k], acc[j]);
        }
      }
      *reinterpret_cast<uint4*>(a.out + (oc * a.out_plane + p) * 8) = pack8_fin<false>(acc, true);
    }
  }
}

// stand-alone upsample_simple (only reached by graphs that use it outside the attention pattern)
__global__ void __launch_bounds__(256) k_upsum(const act_t* __restrict__ in, long long in_plane, int in_chunks, Geo gi,
                                               act_t* __restrict__ out, long long out_plane, int out_chunks, int out_c,
                                               Geo go, int up, int oy, int ox) {
  const long long total = (long long)go.N * go.H * go.W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int n, y, x;
  pix_coords(i, go.H, go.W, n, y, x);
  const long long q = gi.pos(n, (y + oy) / up, (x + ox) / up);
  float sum = 0.f;
  for (int c = 0; c < in_chunks; ++c) {
    float d[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(in + (c * in_plane + q) * 8)), d);
    for (int j = 0; j < 8; ++j) sum += d[j];
  }
  const long long p = go.pos(n, y, x);
  for (int oc = 0; oc < out_chunks; ++oc) {
    float v[8];
    for (int j = 0; j < 8; ++j) v[j] = (oc * 8 + j < out_c) ? sum : 0.f;
    *reinterpret_cast<uint4*>(out + (oc * out_plane + p) * 8) = pack8(v);
  }
}

__global__ void __launch_bounds__(256) k_upsum_f32(const float* __restrict__ in, int Hi, int Wi, float* __restrict__ out,
                                                   int N, int H, int W, int up, int oy, int ox) {
  const long long total = (long long)N * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int x = (int)(p % W);
  const int y = (int)((p / W) % H);
  const int n = (int)(p / ((long long)W * H));
  out[p] = in[((long long)n * Hi + (y + oy) / up) * Wi + (x + ox) / up];
}

__global__ void __launch_bounds__(256) k_copy16(const uint4* __restrict__ in, uint4* __restrict__ out, long long n16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n16) out[i] = in[i];
}

// consumers' integer forms of the probability map: u8 = trunc(p*255) (separator_net_post_processor.py:147,
// heading_net_post_processor.py:287), mask = 255 * (u8[...,0] > thr*255) (helper.py:75-78).
__global__ void __launch_bounds__(256) k_quantize(const float* __restrict__ prob, uint8_t* __restrict__ u8,
                                                  uint8_t* __restrict__ mask, long long npix, int C, float thr255,
                                                  int cut, int Cw) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  uint8_t first = 0;
  for (int c = 0; c < C; ++c) {
    // numpy: np.array(p*255, dtype=np.uint8) - float32 multiply, truncation toward zero
    const float v = prob[p * C + c] * 255.f;
    const uint8_t q = (uint8_t)(int)v;
    if (c == 0) first = q;
    if (u8 && c < Cw) u8[p * Cw + c] = q;    // Cw leading channels only (the consumers read channel 0, sep:33, head:209)
  }
  // helper.py:75-78: u8 > threshold*255; cut >= 0 carries the comparison as an integer (threshold given as a double)
  if (mask) mask[p] = (cut >= 0 ? (int)first >= cut : (float)first > thr255) ? 255 : 0;
}

// two-class maps, 4 pixels per thread: two 128-bit loads, one 64-bit and one 32-bit store
__global__ void __launch_bounds__(256) k_quantize_c2x4(const float4* __restrict__ prob, uint2* __restrict__ u8,
                                                       uint32_t* __restrict__ mask, long long nquad, float thr255, int cut) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nquad) return;
  const float4 a = prob[2 * q], b = prob[2 * q + 1];
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  unsigned qv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) qv[j] = (unsigned)(uint8_t)(int)(v[j] * 255.f);
  if (u8) u8[q] = make_uint2(qv[0] | (qv[1] << 8) | (qv[2] << 16) | (qv[3] << 24),
                             qv[4] | (qv[5] << 8) | (qv[6] << 16) | (qv[7] << 24));
  if (mask) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned first = qv[2 * j];
      if (cut >= 0 ? (int)first >= cut : (float)first > thr255) m |= 0xffu << (8 * j);
    }
    mask[q] = m;
  }
}

__global__ void __launch_bounds__(256) k_unpack_nhwc(const act_t* __restrict__ in, long long in_plane, int C, Geo g,
                                                     float* __restrict__ out) {
  const long long total = (long long)g.N * g.H * g.W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int n, y, x;
  pix_coords(i, g.H, g.W, n, y, x);
  const long long p = g.pos(n, y, x);
  for (int c = 0; c < C; ++c) out[i * C + c] = act_to_f(in[((c >> 3) * in_plane + p) * 8 + (c & 7)]);
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }
static inline long long npix(const Geo& g) { return (long long)g.N * g.H * g.W; }

// ARU_STEM_FFMA=1 keeps the CUDA-core kernel for the pooled 4x4 stem (A/B comparisons)
static bool stem_force_ffma() {
  static const bool v = [] { const char* e = getenv("ARU_STEM_FFMA"); return e && e[0] == '1'; }();
  return v;
}

cudaError_t launch_conv_stem(cudaStream_t st, int ks, const float* in, PV out, PV out_pre, const float* w_host,
                             const float* bias_host, const Geo& g, int act, PV pool, const Geo* pool_geo, bool full,
                             const float* w_dev16, const float* bias_dev16) {
  // w_host: the TF filter [ks][ks][1][C_out] (host memory), bias_host [C_out]; both travel in the kernel parameters
  const int cout = out.C;
  if (cout < 1 || cout > 16 || (ks != 3 && ks != 4) || g.H > 65535 || g.N > 65535) return cudaErrorInvalidValue;
  StemW sw;
  for (int i = 0; i < 16 * 16; ++i) sw.w[i] = 0.f;
  for (int i = 0; i < 16; ++i) sw.b[i] = i < cout ? bias_host[i] : 0.f;
  for (int t = 0; t < ks * ks; ++t)
    for (int c = 0; c < cout; ++c) sw.w[t * 16 + c] = w_host[(size_t)t * cout + c];
  const dim3 grid((unsigned)cdiv(g.W, 128), (unsigned)cdiv(g.H, STEM_ROWS), (unsigned)g.N);
  const int cc = cout <= 8 ? 8 : (cout <= 12 ? 12 : 16);  // instantiated widths (padding channels have zero weights)
  if (pool.p) {   // conv + ReLU + 2x2 max-pool (full == false: the full-resolution tensor is not stored at all)
    if (act != 1 || out_pre.p || !pool_geo) return cudaErrorInvalidValue;
    if (ks == 4 && w_dev16 && bias_dev16 && !stem_force_ffma()) {   // the 16 taps are one K = 16 tensor-core step
      const bool two = cout > 8;
      if (two && full) k_conv_stem4_pool_mma<2, true><<<grid, 128, 0, st>>>(in, out.p, out.plane, pool.p, pool.plane, w_dev16, bias_dev16, g, *pool_geo);
      else if (two) k_conv_stem4_pool_mma<2, false><<<grid, 128, 0, st>>>(in, out.p, out.plane, pool.p, pool.plane, w_dev16, bias_dev16, g, *pool_geo);
      else if (full) k_conv_stem4_pool_mma<1, true><<<grid, 128, 0, st>>>(in, out.p, out.plane, pool.p, pool.plane, w_dev16, bias_dev16, g, *pool_geo);
      else k_conv_stem4_pool_mma<1, false><<<grid, 128, 0, st>>>(in, out.p, out.plane, pool.p, pool.plane, w_dev16, bias_dev16, g, *pool_geo);
      return cudaGetLastError();
    }
#define ARU_STEM_POOL(KK, CC)                                                                                          \
  if (ks == KK && cc == CC) {                                                                                          \
    if (full) k_conv_stem_pool<KK, CC, true><<<grid, 128, 0, st>>>(in, out.p, out.plane, pool.p, pool.plane, sw, g, *pool_geo); \
    else k_conv_stem_pool<KK, CC, false><<<grid, 128, 0, st>>>(in, out.p, out.plane, pool.p, pool.plane, sw, g, *pool_geo);     \
  }
    ARU_STEM_POOL(3, 8) ARU_STEM_POOL(3, 12) ARU_STEM_POOL(3, 16) ARU_STEM_POOL(4, 8) ARU_STEM_POOL(4, 12) ARU_STEM_POOL(4, 16)
#undef ARU_STEM_POOL
    return cudaGetLastError();
  }
  if (ks == 3 && cout <= 8 && (out.p || out_pre.p) && (act == 0 || act == 1) && w_dev16 && bias_dev16 && !stem_force_ffma()) {
    // 1 -> 8 channels (unet_down_0/conv1): the 9 taps as one tensor-core K step; with a fused residual block behind it
    // only the pre-activation is stored
    k_conv_stem3_pre_mma<<<grid, 128, 0, st>>>(in, out.p, out_pre.p, w_dev16, bias_dev16, g, act);
    return cudaGetLastError();
  }
#define ARU_STEM(KK, CC)                                                                                              \
  if (ks == KK && cc == CC)                                                                                           \
    k_conv_stem<KK, CC><<<grid, 128, 0, st>>>(in, out.p, out.plane, out_pre.p, out_pre.plane, sw, g, act);
  ARU_STEM(3, 8) ARU_STEM(3, 12) ARU_STEM(3, 16) ARU_STEM(4, 8) ARU_STEM(4, 12) ARU_STEM(4, 16)
#undef ARU_STEM
  return cudaGetLastError();
}

cudaError_t launch_conv_direct(cudaStream_t st, int ks, PV in, PV out, PV out_pre, PV res, const act_t* w,
                               const float* bias, const Geo& g, int act) {
  dim3 grid(blocks_for(npix(g), 256), out.chunks);
  if (ks == 3)
    k_conv_direct<3><<<grid, 256, 0, st>>>(in.p, in.plane, in.chunks, out.p, out_pre.p, res.p, out.plane, out_pre.plane,
                                           res.plane, out.chunks, w, bias, g, act);
  else if (ks == 4)
    k_conv_direct<4><<<grid, 256, 0, st>>>(in.p, in.plane, in.chunks, out.p, out_pre.p, res.p, out.plane, out_pre.plane,
                                           res.plane, out.chunks, w, bias, g, act);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

template <int KS>
static cudaError_t launch_small_ks(cudaStream_t st, int cout, PV in, float* out, const float* w, const float* bias,
                                   const Geo& g, int act) {
  const size_t smem = (size_t)KS * KS * in.chunks * 8 * cout * sizeof(float);
  const unsigned nb = blocks_for(npix(g), 128);
#define ARU_SMALL(C)                                                                                       \
  case C:                                                                                                  \
    if (smem > 48 * 1024)                                                                                  \
      cudaFuncSetAttribute(k_conv_small<KS, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    k_conv_small<KS, C><<<nb, 128, smem, st>>>(in.p, in.plane, in.chunks, out, w, bias, g, act);          \
    break;
  switch (cout) {
    ARU_SMALL(1) ARU_SMALL(2) ARU_SMALL(3) ARU_SMALL(4) ARU_SMALL(5) ARU_SMALL(6) ARU_SMALL(7) ARU_SMALL(8)
    default:
      return cudaErrorInvalidValue;
  }
#undef ARU_SMALL
  return cudaGetLastError();
}

cudaError_t launch_conv_small(cudaStream_t st, int ks, int cout, PV in, float* out, const float* w, const float* bias,
                              const Geo& g, int act) {
  if (ks == 3) return launch_small_ks<3>(st, cout, in, out, w, bias, g, act);
  if (ks == 4) return launch_small_ks<4>(st, cout, in, out, w, bias, g, act);
  return cudaErrorInvalidValue;
}

cudaError_t launch_deconv_direct(cudaStream_t st, PV in, const Geo& gi, PV out, const Geo& go, const act_t* w,
                                 const float* bias, int act) {
  dim3 grid(blocks_for(npix(go), 256), out.chunks);
  k_deconv_direct<<<grid, 256, 0, st>>>(in.p, in.plane, in.chunks, out.p, out.plane, out.chunks, w, bias, gi, go, act);
  return cudaGetLastError();
}

cudaError_t launch_pool(cudaStream_t st, bool is_max, PV in, const Geo& gi, PV out, const Geo& go) {
  dim3 grid(blocks_for(npix(go), 256), out.chunks);
  if (is_max)
    k_pool<true><<<grid, 256, 0, st>>>(in.p, in.plane, out.p, out.plane, gi, go);
  else
    k_pool<false><<<grid, 256, 0, st>>>(in.p, in.plane, out.p, out.plane, gi, go);
  return cudaGetLastError();
}

cudaError_t launch_pool_f32(cudaStream_t st, bool is_max, const float* in, float* out, int N, int Hi, int Wi, int Ho,
                            int Wo) {
  const long long total = (long long)N * Ho * Wo;
  if (Wi % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0)
    k_pool_f32_x2<<<blocks_for(total / 2, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(in),
                                                               reinterpret_cast<float2*>(out), N, Hi, Wi / 4, Ho, Wo / 2,
                                                               is_max ? 1 : 0);
  else
    k_pool_f32<<<blocks_for(total, 256), 256, 0, st>>>(in, out, N, Hi, Wi, Ho, Wo, is_max ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_combine(cudaStream_t st, const CombineArgs& a_in) {
  CombineArgs a = a_in;
  if (a.A < 1 || a.A > ARU_COMBINE_MAX || a.geo.H > 65535 || a.geo.N > 65535) return cudaErrorInvalidValue;
  auto log2_or_neg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  for (int k = 0; k < a.A; ++k) {
    a.att_sh[k] = log2_or_neg(a.att_up[k]);
    a.det_sh[k] = log2_or_neg(a.det_up[k]);
  }
  const dim3 grid((unsigned)cdiv(a.geo.W, 128), (unsigned)cdiv(a.geo.H, COMBINE_ROWS), (unsigned)a.geo.N);
  int kf = -1, n_full = 0;
  for (int k = 0; k < a.A; ++k)
    if (a.det_up[k] == 1) { kf = k; ++n_full; }
  const bool fast = a.out_chunks == 1 && n_full == 1;
#define ARU_COMBINE(AA)                                                    \
  case AA:                                                                 \
    if (fast) k_combine<AA, true><<<grid, 128, 0, st>>>(a, kf);            \
    else k_combine<AA, false><<<grid, 128, 0, st>>>(a, kf);                \
    break;
  switch (a.A) {
    ARU_COMBINE(1) ARU_COMBINE(2) ARU_COMBINE(3) ARU_COMBINE(4) ARU_COMBINE(5) ARU_COMBINE(6) ARU_COMBINE(7)
    default:
      if (fast) k_combine<8, true><<<grid, 128, 0, st>>>(a, kf);
      else k_combine<8, false><<<grid, 128, 0, st>>>(a, kf);
      break;
  }
#undef ARU_COMBINE
  return cudaGetLastError();
}

cudaError_t launch_upsum(cudaStream_t st, PV in, const Geo& gi, PV out, const Geo& go, int up, int oy, int ox) {
  k_upsum<<<blocks_for(npix(go), 256), 256, 0, st>>>(in.p, in.plane, in.chunks, gi, out.p, out.plane, out.chunks, out.C,
                                                     go, up, oy, ox);
  return cudaGetLastError();
}

cudaError_t launch_upsum_f32(cudaStream_t st, const float* in, int Hi, int Wi, float* out, int N, int H, int W, int up,
                             int oy, int ox) {
  const long long total = (long long)N * H * W;
  k_upsum_f32<<<blocks_for(total, 256), 256, 0, st>>>(in, Hi, Wi, out, N, H, W, up, oy, ox);
  return cudaGetLastError();
}

cudaError_t launch_copy(cudaStream_t st, const void* in, void* out, long long bytes) {
  const long long n16 = bytes / 16;
  k_copy16<<<blocks_for(n16, 256), 256, 0, st>>>((const uint4*)in, (uint4*)out, n16);
  return cudaGetLastError();
}

cudaError_t launch_quantize(cudaStream_t st, const float* prob, uint8_t* u8, uint8_t* mask, long long np, int C,
                            float thr, int cut, int u8_channels) {
  const int Cw = (u8_channels > 0 && u8_channels < C) ? u8_channels : C;
  long long done = 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(prob) & 15) == 0 && (reinterpret_cast<uintptr_t>(u8) & 7) == 0 &&
                       (reinterpret_cast<uintptr_t>(mask) & 3) == 0;
  if (C == 2 && Cw == C && aligned && np >= 4) {
    const long long nquad = np / 4;
    k_quantize_c2x4<<<blocks_for(nquad, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(prob),
                                                            reinterpret_cast<uint2*>(u8), reinterpret_cast<uint32_t*>(mask),
                                                            nquad, thr * 255.f, cut);
    done = nquad * 4;
  }
  if (done < np)
    k_quantize<<<blocks_for(np - done, 256), 256, 0, st>>>(prob + done * C, u8 ? u8 + done * Cw : nullptr,
                                                           mask ? mask + done : nullptr, np - done, C, thr * 255.f, cut, Cw);
  return cudaGetLastError();
}

// Range check of stored 16-bit activations: values at the storage limit (fp16: |v| = 65504, where stores saturate) or
// not finite.  One counter per call; 8 values per thread.
__global__ void __launch_bounds__(256) k_range_scan(const uint4* __restrict__ v, long long n_vec,
                                                    unsigned long long* __restrict__ count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned bad = 0;
  if (i < n_vec) {
    const uint4 q = __ldg(v + i);
    const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const unsigned e = (w[k] >> (16 * h)) & 0x7fffu;
#ifdef ARU_USE_BF16
        bad += e >= 0x7f80u;                 // inf / nan
#else
        bad += e >= 0x7bffu;                 // 65504 (saturated store), inf, nan
#endif
      }
    }
  }
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, (unsigned long long)bad);
}

cudaError_t launch_range_scan(cudaStream_t st, const act_t* p, long long elements, unsigned long long* count) {
  const long long n_vec = elements / 8;
  if (n_vec <= 0) return cudaSuccess;
  k_range_scan<<<blocks_for(n_vec, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(p), n_vec, count);
  return cudaGetLastError();
}

cudaError_t launch_unpack_nhwc(cudaStream_t st, PV in, const Geo& g, float* out) {
  k_unpack_nhwc<<<blocks_for(npix(g), 256), 256, 0, st>>>(in.p, in.plane, in.C, g, out);
  return cudaGetLastError();
}

}  // namespace aru
