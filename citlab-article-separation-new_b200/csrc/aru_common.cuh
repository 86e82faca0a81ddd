// aru_common.cuh - shared types of the B200 ARU-Net engine (device + host).
//
// Activation storage ("chunk-planar"): a tensor with C channels, N pages of H x W is stored as
//   [ceil(C/8)] [N] [H] [W] [8]  16-bit floats
// i.e. 8 channels of one pixel are one 16-byte vector and consecutive pixels of a row are contiguous.
// Why (B200): (1) a 3x3 tap of the implicit GEMM is then a *contiguous* K-major core matrix of
// 8 pixels x 8 channels = 128 B that TMA lands in shared memory without any transposition and that
// a tcgen05 no-swizzle descriptor addresses at any 16-byte shift; (2) the epilogue's one-thread-per-
// pixel TMEM read-out stores 16 B per lane = 512 B contiguous per warp; (3) a ConcatV2 along
// channels is just "more chunk planes" of one buffer. 1-channel tensors (input page, attention
// logits) and the network output stay float32.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef ARU_USE_BF16
typedef __nv_bfloat16 act_t;
typedef __nv_bfloat162 act2_t;
#define ARU_ACT_NAME "bf16"
#else
typedef __half act_t;       // fp16 storage / tensor-core operands, fp32 accumulate (DESIGN.md "numerics")
typedef __half2 act2_t;
#define ARU_ACT_NAME "fp16"
#endif

namespace aru {

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

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
  // saturate instead of producing inf: an overflowing activation must not poison the 0*x products of
  // zero-padded channels downstream
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

// A channel-slice view of a chunk-planar tensor (or a float32 plane when chunks == 0).
struct TView {
  void* base;        // first chunk of the view (act_t*) or float plane
  long long plane;   // elements between consecutive chunks: N*H*W*8
  int N, H, W;
  int C;             // logical channels of the view
  int chunks;        // ceil(C/8); 0 for float32 1-channel planes / NHWC float output
};

}  // namespace aru
