// mma_issue_bench.cu - microbenchmark: how fast can tcgen05.mma (M=128, K=16, small N) be *issued* on sm_100a?
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_issue_bench tools/mma_issue_bench.cu && /tmp/mma_issue_bench
// Variants (cycles per MMA, one CTA per SM on all SMs, median over CTAs):
//   0 single lane, descriptors from a shared-memory table (2 LDS per MMA)
//   1 single lane, descriptor = base + i*stride (pure arithmetic), rolled loop
//   2 single lane, arithmetic, unrolled x8
//   3 whole warp, elect.sync around every MMA, arithmetic, unrolled x8
//   4 single lane, one inline-asm block issuing 8 MMAs with descriptor increments inside the asm
//   5 like 2, two issuer warps (different accumulators)      6 like 2, four issuer warps
//   7 like 4, four issuer warps
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// 8 MMAs in one asm block: A descriptor low word advances by `astep` each time, B fixed.
__device__ __forceinline__ void umma_x8(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t astep) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "mov.b32 al, %1;\n\t"
      "mov.b64 da, {al, %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\tadd.u32 al, al, %6;\n\t"
      "mov.b64 da, {al, %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\tadd.u32 al, al, %6;\n\t"
      "mov.b64 da, {al, %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\tadd.u32 al, al, %6;\n\t"
      "mov.b64 da, {al, %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\tadd.u32 al, al, %6;\n\t"
      "mov.b64 da, {al, %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\tadd.u32 al, al, %6;\n\t"
      "mov.b64 da, {al, %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\tadd.u32 al, al, %6;\n\t"
      "mov.b64 da, {al, %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\tadd.u32 al, al, %6;\n\t"
      "mov.b64 da, {al, %2};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(astep) : "memory");
}

constexpr int NITER = 4080;  // multiple of 24  // MMAs per issuer

template <int VARIANT>
__global__ void __launch_bounds__(256, 1) k_bench(int npad, long long* out_cycles, int a_off) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_ptr;
  __shared__ int2 tab[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const uint32_t s_a = smem_u32(smem), s_b = s_a + 40 * 1024;
  const uint32_t hi = (128u >> 4) | (1u << 14);
  const uint32_t a_lo0 = ((s_a >> 4) + (uint32_t)a_off) | ((16u >> 4) << 16);  // LBO = 16 B (paired-tap style); a_off: misalignment in 16 B units
  const uint32_t b_lo0 = (s_b >> 4) | ((((uint32_t)npad * 16u) >> 4) << 16);
  if (threadIdx.x < 64) tab[threadIdx.x] = make_int2((int)(a_lo0 + threadIdx.x * 8), (int)b_lo0);
  if (threadIdx.x == 0)
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 7) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(npad >> 3) << 17) | ((128u >> 4) << 24);
  constexpr int NISS = VARIANT == 5 ? 2 : (VARIANT == 6 || VARIANT == 7) ? 4 : 1;
  long long t0 = 0, t1 = 0;
  if (warp < NISS) {
    const uint32_t d = tmem + (uint32_t)(warp * 64);
    const uint32_t bar = smem_u32(&bars[warp]);
    if (VARIANT == 3) {
      t0 = clock64();
#pragma unroll 8
      for (int i = 0; i < NITER; ++i) {
        const uint32_t a_lo = a_lo0 + (uint32_t)((i & 63) * 8);
        if (elect_one()) umma_f16(d, desc64(hi, a_lo), desc64(hi, b_lo0), idesc, 1u);
      }
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      while (!mbar_try_wait(bar, 0)) {}
      t1 = clock64();
    } else if (VARIANT == 8 || VARIANT == 9) {
      // 8: accumulator rotates every MMA over 4 tiles, B changes every 4 MMAs (the conv kernel's t-inner order)
      // 9: 6 consecutive MMAs into one accumulator with 6 different B images, then the next tile (k-inner order)
      const uint32_t bstep = (uint32_t)(32 * npad) >> 4;
      t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < NITER; i += 24) {
        const uint32_t a_base = a_lo0 + (uint32_t)((i & 63) * 8);
#pragma unroll
        for (int j = 0; j < 24; ++j) {
          const int t = VARIANT == 8 ? (j & 3) : (j / 6), k = VARIANT == 8 ? (j >> 2) : (j % 6);
          if (elect_one())
            umma_f16(d + (uint32_t)(t * npad), desc64(hi, a_base + (uint32_t)(t * 256 + (k >> 1) * 84 + (k & 1) * 2)),
                     desc64(hi, b_lo0 + (uint32_t)k * bstep), idesc, k ? 1u : 0u);
        }
      }
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      while (!mbar_try_wait(bar, 0)) {}
      t1 = clock64();
    } else if (lane == 0) {
      t0 = clock64();
      if (VARIANT == 0) {
#pragma unroll 2
        for (int i = 0; i < NITER; ++i) {
          const int2 e = tab[i & 63];
          umma_f16(d, desc64(hi, (uint32_t)e.x), desc64(hi, (uint32_t)e.y), idesc, 1u);
        }
      } else if (VARIANT == 1) {
#pragma unroll 1
        for (int i = 0; i < NITER; ++i) {
          const uint32_t a_lo = a_lo0 + (uint32_t)((i & 63) * 8);
          umma_f16(d, desc64(hi, a_lo), desc64(hi, b_lo0), idesc, 1u);
        }
      } else if (VARIANT == 2 || VARIANT == 5 || VARIANT == 6) {
#pragma unroll 8
        for (int i = 0; i < NITER; ++i) {
          const uint32_t a_lo = a_lo0 + (uint32_t)((i & 63) * 8);
          umma_f16(d, desc64(hi, a_lo), desc64(hi, b_lo0), idesc, 1u);
        }
      } else if (VARIANT == 4 || VARIANT == 7) {
#pragma unroll 1
        for (int i = 0; i < NITER; i += 8) {
          const uint32_t a_lo = a_lo0 + (uint32_t)((i & 63) * 8);
          umma_x8(d, a_lo, hi, b_lo0, hi, idesc, 8u);
        }
      }
      umma_commit(bar);
      while (!mbar_try_wait(bar, 0)) {}
      t1 = clock64();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  if (warp == 7) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// Queue depth: time for the elected lane to get n MMAs *accepted* (clock right after the last issue), and until they retire.
__global__ void __launch_bounds__(256, 1) k_queue(int npad, long long* out, int n) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const uint32_t s_a = smem_u32(smem), s_b = s_a + 40 * 1024;
  const uint32_t hi = (128u >> 4) | (1u << 14);
  const uint32_t a_lo0 = (s_a >> 4) | ((16u >> 4) << 16);
  const uint32_t b_lo0 = (s_b >> 4) | ((((uint32_t)npad * 16u) >> 4) << 16);
  if (threadIdx.x == 0)
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 7) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(npad >> 3) << 17) | ((128u >> 4) << 24);
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    const uint32_t bar = smem_u32(&bars[0]);
    uint32_t par = 0;
    for (int rep = 0; rep < 3; ++rep) {  // last repetition is reported (warm instruction cache)
      t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < n; ++i)
        if (elect_one()) umma_f16(tmem, desc64(hi, a_lo0 + (uint32_t)((i & 63) * 8)), desc64(hi, b_lo0), idesc, 1u);
      t1 = clock64();
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      while (!mbar_try_wait(bar, par)) {}
      par ^= 1;
      t2 = clock64();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = t2 - t0; }
  if (warp == 7) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

void run_queue(int npad) {
  long long* d;
  cudaMalloc(&d, 2 * sizeof(long long));
  cudaFuncSetAttribute(k_queue, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int n : {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64}) {
    k_queue<<<1, 256, 64 * 1024>>>(npad, d, n);
    cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("queue N=%3d: %2d MMAs accepted after %5lld cycles, retired after %5lld cycles\n", npad, n, h[0], h[1]);
  }
  cudaFree(d);
}

template <int V>
void run(const char* name, int npad, int a_off = 0) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  cudaFuncSetAttribute(k_bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) k_bench<V><<<sms, 256, 64 * 1024>>>(npad, d, a_off);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  printf("variant %d (%-44s) N=%3d a_off=%d: %7.1f cycles per MMA per issuer (median CTA)  %s\n", V, name, npad, a_off,
         (double)h[sms / 2] / NITER, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1) {  // alignment sweep of the best variant: does a 16-byte-misaligned A start cost tensor-pipe time?
    for (int npad : {16, 64})
      for (int off : {0, 1, 3, 4, 7}) run<3>("warp + elect, unroll 8", npad, off);
    run_queue(16);
    run_queue(64);
    for (int npad : {16, 32, 64}) {
      run<8>("warp + elect, D rotates per MMA (t-inner)", npad);
      run<9>("warp + elect, 6 MMAs per D (k-inner)", npad);
    }
    return 0;
  }
  for (int npad : {16, 64, 256}) {
    run<0>("1 lane, smem table, unroll 2", npad);
    run<1>("1 lane, arithmetic, rolled", npad);
    run<2>("1 lane, arithmetic, unroll 8", npad);
    run<3>("warp + elect, unroll 8", npad);
    run<4>("1 lane, 8 MMAs per asm block", npad);
    run<5>("2 issuer warps, unroll 8", npad);
    run<6>("4 issuer warps, unroll 8", npad);
    run<7>("4 issuer warps, 8 MMAs per asm block", npad);
  }
  return 0;
}
