// hmma_rate_bench.cu - microbenchmark: what does the warp-level tensor path (mma.sync.m16n8k16 f16 -> f32) sustain
// on sm_100a, with operands in registers and with the A fragment re-fetched from shared memory (ldmatrix.x4)?
// Decides whether a fused residual block for the 8 / 16 channel levels can run on register operands.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hmma_rate_bench tools/hmma_rate_bench.cu && /tmp/hmma_rate_bench
// Variants:
//   0  MMAs only: NACC independent accumulators per warp, A / B fragments constant in registers
//   1  one ldmatrix.x4 (512 B of shared memory per warp) per MMA
//   2  one ldmatrix.x4 per 3 MMAs (the vertical-tap reuse of a fused 3x3 block)
//   4  m16n8k8 only (FLOP figures printed assume 4096 per MMA: halve them)      5  alternating m16n8k16 / m16n8k8
//   6 / 7 / 8  MMAs only as in 0, plus 2 / 4 / 8 independent FFMAs per MMA (does ALU work issue under a busy tensor pipe?)
//   3  like 2 plus one 16-byte shared-memory store per lane per 15 MMAs (an intermediate row written back)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <algorithm>

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldsm4(uint32_t (&a)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}

template <int VAR, int NACC>
__global__ void __launch_bounds__(NACC > 8 ? 512 : 1024) k_rate(int iters, float* out, long long* cyc) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  __syncthreads();
  float acc[NACC][4];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
  uint32_t b0 = 0x2c002c00u + lane, b1 = 0x2c002c00u;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem) + (warp & 7) * 2048 + lane * 16;
  float f[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) f[q] = (float)(lane + q);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      if (VAR == 1) ldsm4(a, sbase + ((it + j) & 3) * 512);
      if ((VAR == 2 || VAR == 3) && (j % 3) == 0) ldsm4(a, sbase + ((it + j) & 3) * 512);
      if (VAR == 4) mma1688(acc[j], a[0], a[1], b0);
      else if (VAR == 5 && (j & 1)) mma1688(acc[j], a[0], a[1], b0);
      else mma16816(acc[j], a, b0, b1);
      if (VAR >= 6 && VAR <= 8) {
        constexpr int NF = VAR == 6 ? 2 : VAR == 7 ? 4 : 8;
#pragma unroll
        for (int q = 0; q < NF; ++q) f[q] = fmaf(f[q], 1.0001f, 0.5f);
      }
    }
    if (VAR == 3 && (it % 2) == 0) {
      uint4 v = make_uint4(__float_as_uint(acc[0][0]), __float_as_uint(acc[1][1]), it, lane);
      *reinterpret_cast<uint4*>(smem + 16384 + threadIdx.x * 16) = v;
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NACC; ++j) s += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
#pragma unroll
  for (int q = 0; q < 8; ++q) s += f[q];
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int VAR, int NACC>
static void run(int warps, int iters, int sms) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  const int smem = 16384 + 1024 * 16;
  cudaFuncSetAttribute(k_rate<VAR, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_rate<VAR, NACC><<<sms, warps * 32, smem>>>(iters, out, cyc);
  cudaEventRecord(e0);
  k_rate<VAR, NACC><<<sms, warps * 32, smem>>>(iters, out, cyc);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const double mmas = (double)warps * iters * NACC;
  const double flop = mmas * 4096.0;
  printf("var %d nacc %2d warps %2d: %7.1f FLOP/clk/SM  (%.2f clk per warp-MMA per SM)  %8.1f TFLOP/s over %d SMs  [%s]\n", VAR, NACC,
         warps, flop / (double)h[sms / 2], (double)h[sms / 2] / mmas, flop * sms / (ms * 1e-3) / 1e12, sms,
         cudaGetErrorString(err));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs, %d kHz\n", p.name, sms, p.clockRate);
  const int it = 20000;
  for (int w : {4, 8, 16, 32}) run<0, 8>(w, it, sms);
  for (int w : {4, 8, 16}) run<0, 15>(w, it, sms);
  for (int w : {4, 8, 16, 32}) run<1, 8>(w, it, sms);
  for (int w : {4, 8, 16}) run<2, 15>(w, it, sms);
  for (int w : {4, 8, 16}) run<3, 15>(w, it, sms);
  for (int w : {4, 12}) run<6, 8>(w, it, sms);
  for (int w : {4, 12}) run<7, 8>(w, it, sms);
  for (int w : {4, 12}) run<8, 8>(w, it, sms);
  for (int w : {4, 8, 16}) run<4, 8>(w, it, sms);
  for (int w : {4, 8, 16}) run<5, 8>(w, it, sms);
  return 0;
}
