// aru_common.cuh - shared types of the B200 ARU-Net engine (device + host).
//
// Activation storage ("flattened padded chunk-planar"): a tensor with C channels for N pages of
// H x W lives in ceil(C/8) planes; every plane is a 1-D array of *positions*, one position = the 8
// channels of one pixel = one 16-byte vector of 16-bit floats.  Pages are stacked with a one-pixel
// zero frame that neighbouring rows / pages share:
//     Wp = W + 2,  Hp = H + 2
//     pos(n, y, x) = lead + (n*Hp + y + 1)*Wp + (x + 1)
// Position 'pos + dy*Wp + dx' is the neighbour (y+dy, x+dx); out-of-image neighbours land on frame
// positions, which are zero (the arena is cleared once per plan and kernels only ever store to
// in-image positions).  Consequences, all of them deliberate (B200):
//  (1) a KxK SAME convolution is a 1-D correlation over positions with K*K constant offsets - the
//      implicit-GEMM A operand of tap (ky,kx) for 128 consecutive output positions is a *contiguous*
//      128 x 16 B run, i.e. 16 canonical no-swizzle K-major core matrices that a tcgen05 shared-memory
//      descriptor addresses at any 16-byte shift.  TF's asymmetric padding of even kernels
//      (1 before / 2 after for 4x4) is just the offset table;
//  (2) the tile stream of a CTA is a contiguous range of positions, so operand staging is a ring of
//      2 KB bulk-copy (TMA engine) chunks with no 2-D bookkeeping, read amplification ~1.0;
//  (3) the epilogue's one-lane-per-position TMEM read-out stores 16 B per lane = 512 B contiguous per
//      warp and plane; a ConcatV2 along channels is "more planes" of one buffer.
// 1-channel tensors (input page pyramid, attention logits) are dense float32 [N][H][W]; the network
// output is dense float32 NHWC.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef ARU_USE_BF16
typedef __nv_bfloat16 act_t;
typedef __nv_bfloat162 act2_t;
#define ARU_ACT_NAME "bf16"
#define ARU_UMMA_FMT 1u
#else
typedef __half act_t;       // fp16 storage / tensor-core operands, fp32 accumulate (DESIGN.md "numerics")
typedef __half2 act2_t;
#define ARU_ACT_NAME "fp16"
#define ARU_UMMA_FMT 0u
#endif

namespace aru {

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Geometry of a chunk-planar tensor (identical for all tensors with the same N,H,W).
struct Geo {
  int N, H, W;
  int Wp, Hp;
  long long lead;   // positions before the frame of page 0
  long long plane;  // positions per plane (multiple of 128)
  __host__ __device__ long long pos(int n, int y, int x) const {
    return lead + ((long long)n * Hp + y + 1) * Wp + (x + 1);
  }
};

inline Geo make_geo(int N, int H, int W) {
  Geo g;
  g.N = N; g.H = H; g.W = W;
  g.Wp = W + 2; g.Hp = H + 2;
  // margins: conv_tc loads whole 512-position units around the first / last in-image position
  g.lead = ((long long)(g.Wp + 1 + 1024 + 127) / 128) * 128;
  long long body = (long long)N * g.Hp * g.Wp;
  g.plane = ((g.lead + body + 2LL * g.Wp + 1280 + 127) / 128) * 128;
  return g;
}

__device__ __forceinline__ float act_to_f(act_t v) {
#ifdef ARU_USE_BF16
  return __bfloat162float(v);
#else
  return __half2float(v);
#endif
}
__device__ __forceinline__ act_t f_to_act(float v) {
#ifdef ARU_USE_BF16
  return __float2bfloat16_rn(v);
#else
  // saturate instead of producing inf: an overflowing activation must not poison later 0*x products
  return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
#endif
}
__device__ __forceinline__ float2 act2_to_f2(act2_t v) {
#ifdef ARU_USE_BF16
  return __bfloat1622float2(v);
#else
  return __half22float2(v);
#endif
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  act_t lo = f_to_act(a), hi = f_to_act(b);
  return (uint32_t)(*reinterpret_cast<uint16_t*>(&lo)) | ((uint32_t)(*reinterpret_cast<uint16_t*>(&hi)) << 16);
}
__device__ __forceinline__ void unpack8(const uint4& v, float f[8]) {
  const act2_t* p = reinterpret_cast<const act2_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = act2_to_f2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float f[8]) {
  uint4 v;
  v.x = pack2(f[0], f[1]);
  v.y = pack2(f[2], f[3]);
  v.z = pack2(f[4], f[5]);
  v.w = pack2(f[6], f[7]);
  return v;
}

// ---- 16-bit packing of kernel results (packed half2 path: 4 converts + 8 min/max per 8 values) ----
// two fp32 -> packed 16-bit pair, saturated to the finite range (fp16: an overflowing activation must not
// become inf and poison later 0*x products), optionally ReLU'd.  max(.,0) commutes with the rounding.
template <bool RELU>
__device__ __forceinline__ uint32_t pack2_fin(float x, float y) {
  // one F2FP: round to nearest even, clamp to the largest finite value, optional ReLU (x -> low half, y -> high half)
  uint32_t d;
#ifdef ARU_USE_BF16
  if (RELU) asm("cvt.rn.relu.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(y), "f"(x));
  else asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(y), "f"(x));
#else
  if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(y), "f"(x));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(y), "f"(x));
#endif
  return d;
}
template <bool RELU>
__device__ __forceinline__ uint4 pack8_fin(const float f[8], bool valid) {
  uint4 v;
  v.x = pack2_fin<RELU>(f[0], f[1]);
  v.y = pack2_fin<RELU>(f[2], f[3]);
  v.z = pack2_fin<RELU>(f[4], f[5]);
  v.w = pack2_fin<RELU>(f[6], f[7]);
  if (!valid) v = make_uint4(0u, 0u, 0u, 0u);   // frame / margin positions are (re)written with zeros
  return v;
}

}  // namespace aru
