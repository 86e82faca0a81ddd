// band_common.cuh - PTX wrappers (mbarrier, bulk copy, tcgen05) and packed 16-bit helpers shared by the row-banded
// tensor-core kernels (conv_band.cu, conv_band2.cu).
#pragma once
#include "aru_common.cuh"

namespace aru {
namespace {

// ---- PTX wrappers (same conventions as conv_tc.cu) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait (0.2 s): a protocol bug records its barrier class and traps instead of hanging the GPU.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  const unsigned long long t0 = global_ns();
  while (true) {
    for (int i = 0; i < 64; ++i)
      if (mbar_try_wait(bar, parity)) return;
    if (global_ns() - t0 > 200000000ull) {
      atomicCAS(err_flag, 0, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, err_flag, code);
}
// Waiters that are not on the critical path (epilogue warps, producer) back off between polls: nine warps spinning on
// try_wait saturate the shared-memory port and every shared-memory access of the MMA issuer then takes ~200 cycles.
__device__ __noinline__ void mbar_wait_sleep_slow(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  const unsigned long long t0 = global_ns();
  while (true) {
    for (int i = 0; i < 32; ++i) {
      __nanosleep(48);
      if (mbar_try_wait(bar, parity)) return;
    }
    if (global_ns() - t0 > 200000000ull) {
      atomicCAS(err_flag, 0, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_sleep_slow(bar, parity, err_flag, code);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// descriptor high word: SBO = 128 B (8 rows of a core matrix are contiguous), no swizzle, sm_100 descriptor version
__device__ __forceinline__ uint32_t desc_hi128() { return ((128u >> 4) & 0x3FFF) | (1u << 14); }
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// unsaturated 16-bit pair (the position-major pass clamps: an inf formed here becomes +-65504 there)
__device__ __forceinline__ uint32_t pack2_raw(float x, float y) {
#ifdef ARU_USE_BF16
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
#else
  __half2 h = __floats2half2_rn(x, y);
#endif
  return *reinterpret_cast<uint32_t*>(&h);
}

// wait-time counters of the pipeline roles (ARU_BAND_DBG=16 prints them); the clock reads sit in every hot loop, so
// they are compiled in only with -DARU_BAND_STATS
#ifdef ARU_BAND_STATS
#define BAND_CLK() clock64()
#else
#define BAND_CLK() 0LL
#endif
__device__ unsigned long long g_band_stats[160][8];

// 16-byte read-only load that the compiler cannot sink to its use (the point is to have it in flight early)
__device__ __forceinline__ uint4 ldg_nc_v4(const act_t* p, bool pred) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (pred)
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// finite clamp (+ ReLU) of four packed 16-bit pairs without leaving the packed domain
template <bool RELU>
__device__ __forceinline__ uint4 clamp8(uint4 v) {
#ifdef ARU_USE_BF16
  if (RELU) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
    const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __hmax2(h[i], z);
  }
#else
  __half2* h = reinterpret_cast<__half2*>(&v);
  const __half2 hi = __half2half2(__ushort_as_half((unsigned short)0x7bff));
  const __half2 lo = RELU ? __half2half2(__ushort_as_half((unsigned short)0)) : __half2half2(__ushort_as_half((unsigned short)0xfbff));
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __hmax2(__hmin2(h[i], hi), lo);
#endif
  return v;
}

// packed 16-bit add: a correctly rounded add of two 16-bit floats equals the fp32 add rounded once, so this matches the
// fp32 epilogue of the other kernels bit for bit (the sum then goes through clamp8)
__device__ __forceinline__ uint4 add8(uint4 x, const uint4& y) {
  act2_t* a = reinterpret_cast<act2_t*>(&x);
  const act2_t* b = reinterpret_cast<const act2_t*>(&y);
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __hadd2(a[i], b[i]);
  return x;
}

__device__ __forceinline__ uint4 max8(uint4 x, const uint4& y) {
  act2_t* a = reinterpret_cast<act2_t*>(&x);
  const act2_t* b = reinterpret_cast<const act2_t*>(&y);
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __hmax2(a[i], b[i]);
  return x;
}

// tile index -> page, strip, row tile
struct TileRef {
  int n, s, ty;
};
__device__ __forceinline__ TileRef tile_ref(long long L, int n_strips, int n_ty) {
  TileRef t;
  const int per_page = n_strips * n_ty;
  t.n = (int)(L / per_page);
  const int r = (int)(L - (long long)t.n * per_page);
  t.s = r / n_ty;
  t.ty = r - t.s * n_ty;
  return t;
}

}  // namespace
}  // namespace aru
