// conv_tc.cu - tcgen05 / TMEM implicit-GEMM convolution for sm_100a.
//
// Computes  out = act(conv_KxK_SAME(in) + bias [+ res])  (layers.py:191-247, ARU_v1.py:212-227) on the
// flattened padded chunk-planar layout of aru_common.cuh.  In that layout the convolution is a 1-D
// correlation over positions:  out[p] = sum_{tap} W[tap] . in[p + off(tap)],  off = (ky-1)*Wp + (kx-1).
//
// GEMM view per 128-position tile:  D[128, N=C_out] = sum over K-steps of A[128, 16] * B[16, N]
//   * A (activations): for tap t and channel chunks (2c, 2c+1) the 128 x 16 operand is two contiguous
//     runs of 128 x 16 B in shared memory -> canonical no-swizzle K-major core matrices,
//     SBO = 128 B (next 8 positions), LBO = ring plane stride (next 8 channels).  For C_in = 8 two
//     horizontally adjacent taps form one K=16 step (LBO = 16 B: the same run shifted by one position).
//   * B (weights): pre-packed on the host into the exact shared-memory image.  Resident for the whole
//     kernel when it fits (weight-stationary persistent CTAs), otherwise streamed per pass in groups of
//     K-steps through a small ring (C >= 128 layers).
//   * D: fp32 in TMEM; a *pass* is T consecutive tiles (T x N columns per accumulator stage), and there
//     are 2 or 4 stages so the epilogue of pass i overlaps the MMAs of pass i+1.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over a contiguous range of passes):
//   warp 1 lane 0 : producer - streams the input as bulk copies (TMA engine, cp.async.bulk) of one *unit*
//                   (128*T positions x all input chunks) into a ring; the head of the ring is mirrored
//                   behind its end so that every 128(+1)-position operand run is contiguous even when it
//                   wraps.  Read amplification is ~1: every position is fetched once per CTA.
//   warp 0 lane 0 : MMA issuer - tcgen05.mma.cta_group::1.kind::f16.  tools/mma_issue_bench.cu measured the
//                   issue side on B200: an M=128,K=16 MMA with N <= 64 occupies the pipe for 39-48 cycles
//                   (the 4 KB shared-memory A read), one thread reaches ~50 cycles/MMA with straight-line
//                   code and arithmetic descriptors, table look-ups or rolled loops cost 2-3x and extra
//                   issuer warps do not help.  Hence the kernel is templated on the filter size so the tap
//                   loops unroll completely and descriptors are base + compile-time-shaped offsets.
//                   tcgen05.commit frees ring units / weight stages and publishes accumulators
//   warp 2        : TMEM allocator
//   warps 4..11   : two epilogue warpgroups (alternating passes): tcgen05.ld -> +bias (+res) -> act ->
//                   16-bit pack -> 16 B coalesced global stores (in-image positions only; the zero frame of
//                   the output tensor is never written)
// All mbarrier waits are bounded: a protocol bug records its barrier class in *err_flag and traps (the launch fails
// loudly with a CUDA error) instead of hanging the GPU.
#include <cstring>
#include <vector>

#include "aru_common.cuh"
#include "kernels.h"

namespace aru {

namespace {

constexpr int TILE = 128;
constexpr int NUM_THREADS = 384;
constexpr int MAX_UNITS = 96;
constexpr int MAX_STEPS = 256;
constexpr int MIRROR = 136;          // positions mirrored behind the ring (129 needed, rounded up)
constexpr int PREFETCH_UNITS = 2;
constexpr int MAX_WSTAGES = 4;
// shared-memory header layout (bytes)
constexpr int OFF_FULL = 0;                                // uint64 full[MAX_UNITS]
constexpr int OFF_EMPTY = OFF_FULL + 8 * MAX_UNITS;        // uint64 empty[MAX_UNITS]
constexpr int OFF_TFULL = OFF_EMPTY + 8 * MAX_UNITS;       // uint64 tmem_full[4]
constexpr int OFF_TEMPTY = OFF_TFULL + 32;                 // uint64 tmem_empty[4]
constexpr int OFF_WFULL = OFF_TEMPTY + 32;                 // uint64 w_full[4]
constexpr int OFF_WEMPTY = OFF_WFULL + 32;                 // uint64 w_empty[4]
constexpr int OFF_TMEMPTR = OFF_WEMPTY + 32;               // uint32 tmem base
constexpr int OFF_BIAS = OFF_TMEMPTR + 16;                 // float bias[256]
constexpr int HDR_BYTES = ((OFF_BIAS + 4 * 256 + 127) / 128) * 128;
constexpr int TAB_MAX = 768;         // A-operand start table entries (n_units * T * KS), lives in the kernel parameters

struct ConvTcArgs {
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
  int ks, cin_chunks, cout_chunks, npad, n_steps;
  int n_units, n_stages, tmem_cols;
  int w_stream, w_group, n_wst;   // weight streaming: steps per group, stages
  int W, H, Wp, Hp, N;
  int lead, body_end, rel_bias;
  int pass_begin, pass_end;
  int dlo, dhi, min_off;
  int dc128, dr128, dcT, drT;     // (col,row) advance of 128 and 128*T positions
  int dcS, drS;                   // (col,row) advance of 2*128*T positions (one epilogue warpgroup stride)
  int act;
  // transposed convolution (3x3 stride 2 as a 2x2 correlation with four output-parity classes in N)
  int deconv, cls_chunks, bias_pages;
  int Wo, Ho, Wpo, Hpo, lead_o, offy, offx;
  // dense float32 NHWC output [N][H][W][f32_c] with the head activation (classifier / attention logit), else null
  float* out_f32;
  int f32_c;
  int* err_flag;
  // Ring position (16-byte units) of the A operand of tile t, filter row ky, for ring phase s = pass % n_units:
  // tab[(s*T + t)*KS + ky].  Precomputed on the host so that the issuer's descriptor arithmetic is
  // base + table entry + compile-time tap offset: no ring-wrap arithmetic on the single issuing thread.
  uint32_t tab[TAB_MAX];
};

// ---- PTX wrappers ------------------------------------------------------------------------------
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
// Bounded wait: when the barrier did not flip within 0.2 s the barrier class is recorded and the kernel traps.
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
// Fully inlined variant for the MMA issuer: a CALL in its loop would force every loop-carried descriptor value out of
// the uniform registers (they are not preserved across calls) and back through R2UR before each UTCHMMA.
__device__ __forceinline__ void mbar_wait_inline(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (global_ns() - t0 > 200000000ull) {
      atomicCAS(err_flag, 0, code);
      __threadfence_system();
      __trap();
    }
  }
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t v[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a converged warp.  Issuing the UMMA instructions under this predicate from warp-uniform control flow
// lets the compiler keep descriptors in uniform registers instead of a per-lane waterfall around every instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// T MMAs of one K-step (the T tiles of a pass) in ONE asm block under ONE election: A descriptor low words
// a_lo[t] + a_add, a common B descriptor, accumulators d0 + t * dstep.  With warp-uniform operands ptxas keeps all of
// it on the uniform datapath: ~2-3 instructions per UTCHMMA instead of ~15 with a C++-level election per MMA.
#define ARU_MMA_HEAD                                                                                       \
  "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 al, dd;\n\t"                                      \
  "elect.sync _|p, 0xffffffff;\n\t"                                                                        \
  "setp.ne.b32 q, %4, 0;\n\t"                                                                              \
  "mov.b64 db, {%1, %2};\n\t"                                                                              \
  "mov.b32 dd, %5;\n\t"
#define ARU_MMA_ONE(AREG)                                                                                  \
  "add.u32 al, " AREG ", %0;\n\tmov.b64 da, {al, %2};\n\t"                                                  \
  "@p tcgen05.mma.cta_group::1.kind::f16 [dd], da, db, %3, q;\n\t"                                         \
  "add.u32 dd, dd, %6;\n\t"
template <int T>
__device__ __forceinline__ void umma_tiles(const uint32_t (&a_lo)[T], uint32_t a_add, uint32_t b_lo, uint32_t hi,
                                           uint32_t d0, uint32_t dstep, uint32_t idesc, uint32_t acc) {
  static_assert(T == 1 || T == 2 || T == 4, "tiles per pass");
  if constexpr (T == 1) {
    asm volatile(ARU_MMA_HEAD ARU_MMA_ONE("%7") "}" ::"r"(a_add), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc), "r"(d0),
                 "r"(dstep), "r"(a_lo[0])
                 : "memory");
  } else if constexpr (T == 2) {
    asm volatile(ARU_MMA_HEAD ARU_MMA_ONE("%7") ARU_MMA_ONE("%8") "}" ::"r"(a_add), "r"(b_lo), "r"(hi), "r"(idesc),
                 "r"(acc), "r"(d0), "r"(dstep), "r"(a_lo[0]), "r"(a_lo[1])
                 : "memory");
  } else {
    asm volatile(ARU_MMA_HEAD ARU_MMA_ONE("%7") ARU_MMA_ONE("%8") ARU_MMA_ONE("%9") ARU_MMA_ONE("%10") "}" ::"r"(a_add),
                 "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc), "r"(d0), "r"(dstep), "r"(a_lo[0]), "r"(a_lo[1]), "r"(a_lo[2]),
                 "r"(a_lo[3])
                 : "memory");
  }
}
#undef ARU_MMA_HEAD
#undef ARU_MMA_ONE

// High word of a shared-memory matrix descriptor (no swizzle, K-major): SBO in 16-byte units at bits 32..45,
// descriptor version 1 (sm_100) at bits 46..47.  Low word = start address >> 4 | (LBO >> 4) << 16.
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14); }
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

__device__ __forceinline__ void walk_adv(int& col, int& rin, int& pg, int dc, int dr, int Wp, int Hp) {
  col += dc;
  rin += dr;
  if (col >= Wp) { col -= Wp; ++rin; }
  if (rin >= Hp) {
    rin -= Hp;
    ++pg;
    if (rin >= Hp) { pg += rin / Hp; rin %= Hp; }
  }
}

// ---- the kernel --------------------------------------------------------------------------------
// MODE 0: C_in = 8 (paired taps)   1: resident weights, >= 2 input chunks   2: streamed weights
template <int T, int KS, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv_tc(const __grid_constant__ ConvTcArgs a) {
  constexpr bool CIN1 = (MODE == 0);
  extern __shared__ __align__(128) uint8_t smem[];

  constexpr int UNIT = TILE * T;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_full = s_base + OFF_FULL, s_empty = s_base + OFF_EMPTY;
  const uint32_t s_tfull = s_base + OFF_TFULL, s_tempty = s_base + OFF_TEMPTY;
  const uint32_t s_wfull = s_base + OFF_WFULL, s_wempty = s_base + OFF_WEMPTY;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEMPTR);
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  const int step_bytes = 32 * a.npad;
  const int w_smem_bytes = a.w_stream ? a.n_wst * a.w_group * step_bytes : a.n_steps * step_bytes;
  const uint32_t s_w = s_base + HDR_BYTES;
  const uint32_t s_ring = s_w + ((w_smem_bytes + 127) / 128) * 128;
  const int L = a.n_units * UNIT;        // ring length in positions
  const int ring_plane = L + MIRROR;

  // static partition of the pass range over the persistent CTAs
  const int total = a.pass_end - a.pass_begin;
  const int per_cta = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int u0 = a.pass_begin + (int)blockIdx.x * per_cta;
  const int n_pass = max(min(u0 + per_cta, a.pass_end) - u0, 0);
  const int win = a.dhi - a.dlo + 1;     // units a pass reads

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.n_units; ++i) {
      mbar_init(s_full + 8 * i, 1);
      mbar_init(s_empty + 8 * i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(s_tfull + 8 * i, 1);
      mbar_init(s_tempty + 8 * i, 4);
      mbar_init(s_wfull + 8 * i, 1);
      mbar_init(s_wempty + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    for (int i = threadIdx.x; i < (a.deconv ? a.cls_chunks * 8 : a.npad); i += NUM_THREADS) s_bias[i] = a.bias[i];
    if (!a.w_stream) {  // resident B image
      const uint4* wsrc = reinterpret_cast<const uint4*>(a.wpack);
      uint4* wdst = reinterpret_cast<uint4*>(smem + HDR_BYTES);
      for (int i = threadIdx.x; i < w_smem_bytes / 16; i += NUM_THREADS) wdst[i] = __ldg(wsrc + i);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA) reads
    }
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int n_groups = a.w_stream ? (a.n_steps + a.w_group - 1) / a.w_group : 0;

  if (n_pass > 0) {
    if (warp == 0) {
      // ================= MMA issuer =================
      // The whole warp runs this role convergently and one elected lane issues each instruction: ptxas then keeps
      // descriptors in uniform registers and emits back-to-back UTCHMMAs (39 cycles / MMA at N = 16 measured by
      // tools/mma_issue_bench.cu variant 3, against 50 for a single-lane region with its per-instruction waterfall).
      const uint32_t idesc = (1u << 4) | (ARU_UMMA_FMT << 7) | (ARU_UMMA_FMT << 10) | ((uint32_t)(a.npad >> 3) << 17) |
                             ((uint32_t)(TILE >> 4) << 24);
      const uint32_t hi = desc_hi(128);
      const uint32_t wstage_units = (uint32_t)(a.w_group * step_bytes) >> 4;
      const uint32_t step_units = (uint32_t)step_bytes >> 4;
      const uint32_t a_lbo = CIN1 ? 16u : (uint32_t)ring_plane * 16u;
      const uint32_t a_lo0 = (s_ring >> 4) | ((a_lbo >> 4) << 16);
      const uint32_t b_lo0 = (s_w >> 4) | ((((uint32_t)a.npad * 16u) >> 4) << 16);
      const uint32_t cp_stride = (uint32_t)(2 * ring_plane);  // descriptor units between channel-chunk pairs
      const int n_cp = CIN1 ? 1 : a.cin_chunks / 2;
      const uint32_t npad = (uint32_t)a.npad;
      // incremental ring bookkeeping (no divisions in the loop)
      int stage = 0, f_slot = 0, e_slot = 0, wst = 0, phase = 0;
      uint32_t stage_par = 0, f_par = 0, w_par = 0;
      for (int pi = 0; pi < n_pass; ++pi) {
        mbar_wait_inline(s_tempty + 8 * stage, stage_par ^ 1, a.err_flag, 2);
        // operands: pass pi reads units pi .. pi+win-1 (the first pass waits for its whole window)
        for (int k = (pi == 0 ? 0 : win - 1); k < win; ++k) {
          mbar_wait_inline(s_full + 8 * f_slot, f_par, a.err_flag, 3);
          if (++f_slot == a.n_units) { f_slot = 0; f_par ^= 1; }
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(stage * T) * npad;
        const uint32_t* tab = a.tab + phase * (T * KS);
        // K loop in the order the B image was packed: taps (ky, kx) unrolled at compile time, channel-chunk
        // pairs inside.  C_in = 8 (CIN1): horizontally adjacent taps are paired into one K=16 step (the B
        // image holds zeros for the missing partner of the last tap of an odd-width kernel).
        uint32_t b_lo = b_lo0;
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          uint32_t a_row[T];
#pragma unroll
          for (int t = 0; t < T; ++t) a_row[t] = a_lo0 + tab[t * KS + ky];
#pragma unroll
          for (int kx = 0; kx < KS; kx += (CIN1 ? 2 : 1)) {
            if constexpr (MODE == 0) {
              umma_tiles<T>(a_row, (uint32_t)kx, b_lo, hi, d_tmem, npad, idesc, (ky | kx) ? 1u : 0u);
              b_lo += step_units;
            } else if constexpr (MODE == 1) {
              // rolled over the channel-chunk pairs (keeps the kernel small enough for the instruction cache)
#pragma unroll 1
              for (int cp = 0; cp < n_cp; ++cp) {
                umma_tiles<T>(a_row, (uint32_t)kx + (uint32_t)cp * cp_stride, b_lo, hi, d_tmem, npad, idesc,
                              (ky | kx | cp) ? 1u : 0u);
                b_lo += step_units;
              }
            } else {
              // streamed weights: groups of w_group steps (w_group divides n_cp, so a group never straddles taps)
#pragma unroll 1
              for (int cp0 = 0; cp0 < n_cp; cp0 += a.w_group) {
                mbar_wait_inline(s_wfull + 8 * wst, w_par, a.err_flag, 6);
                tc_fence_after();
                b_lo = b_lo0 + (uint32_t)wst * wstage_units;
#pragma unroll 1
                for (int cp = cp0; cp < cp0 + a.w_group; ++cp) {
                  umma_tiles<T>(a_row, (uint32_t)kx + (uint32_t)cp * cp_stride, b_lo, hi, d_tmem, npad, idesc,
                                (ky | kx | cp) ? 1u : 0u);
                  b_lo += step_units;
                }
                if (elect_one()) umma_commit(s_wempty + 8 * wst);  // stage may be refilled once these MMAs retire
                if (++wst == a.n_wst) { wst = 0; w_par ^= 1; }
              }
            }
          }
        }
        if (elect_one()) {
          umma_commit(s_empty + 8 * e_slot);  // unit pi is not needed by later passes
          umma_commit(s_tfull + 8 * stage);   // accumulators ready for the epilogue
        }
        if (++e_slot == a.n_units) e_slot = 0;
        if (++stage == a.n_stages) { stage = 0; stage_par ^= 1; }
        if (++phase == a.n_units) phase = 0;
      }
    } else if (warp == 1) {
      if (lane == 0) {
      // ================= producer =================
      const int n_load = n_pass + win - 1;
      const long long first_pos = (long long)(u0 + a.dlo) * UNIT;  // ring origin in plane positions
      const uint32_t tx = (uint32_t)a.cin_chunks * UNIT * 16;
      const uint32_t tx_mirror = (uint32_t)a.cin_chunks * MIRROR * 16;
      int slot = 0;            // k % n_units and the parity of k / n_units, kept incrementally (no divisions
      uint32_t use_par = 0;    // in the steady-state loops of any role)
      auto load_unit = [&](int k) {
        mbar_wait(s_empty + 8 * slot, use_par ^ 1, a.err_flag, 1);
        const uint32_t bar = s_full + 8 * slot;
        const act_t* src = a.in + (first_pos + (long long)k * UNIT) * 8;
        mbar_expect_tx(bar, slot == 0 ? tx + tx_mirror : tx);
        for (int c = 0; c < a.cin_chunks; ++c) {
          const uint32_t dst = s_ring + (uint32_t)(c * ring_plane + slot * UNIT) * 16;
          bulk_g2s(dst, src + c * a.in_plane * 8, UNIT * 16, bar);
          if (slot == 0) bulk_g2s(dst + (uint32_t)L * 16, src + c * a.in_plane * 8, MIRROR * 16, bar);
        }
        if (++slot == a.n_units) { slot = 0; use_par ^= 1; }
      };
      if (!a.w_stream) {
        for (int k = 0; k < n_load; ++k) load_unit(k);
      } else {
        // consumption order: window of pass 0 (+ prefetch), then per pass its weight groups and one more unit
        int next = 0;
        for (; next < min(n_load, a.n_units - 1); ++next) load_unit(next);
        int ws = 0;            // weight stage and its use parity
        uint32_t w_par = 0;
        for (int pi = 0; pi < n_pass; ++pi) {
          for (int g = 0; g < n_groups; ++g) {
            mbar_wait(s_wempty + 8 * ws, w_par ^ 1, a.err_flag, 5);
            const int steps = min(a.w_group, a.n_steps - g * a.w_group);
            const uint32_t bytes = (uint32_t)(steps * step_bytes);
            mbar_expect_tx(s_wfull + 8 * ws, bytes);
            bulk_g2s(s_w + (uint32_t)(ws * a.w_group * step_bytes),
                     reinterpret_cast<const uint8_t*>(a.wpack) + (size_t)g * a.w_group * step_bytes, bytes, s_wfull + 8 * ws);
            if (++ws == a.n_wst) { ws = 0; w_par ^= 1; }
          }
          if (next < n_load) load_unit(next++);
        }
      }
      }
    } else if (warp >= 4) {
      // ================= epilogue =================
      const int wg = (warp - 4) >> 2;  // warpgroup 0/1 -> even/odd passes
      const int q = warp & 3;          // TMEM lane quarter this warp may access
      const int m = q * 32 + lane;
      // (col, rin, pg) of this lane's position in tile 0 of its first pass; rel_bias makes the numerator non-negative
      int p = (u0 + wg) * UNIT + m;
      int col, rin, pg;
      {
        const int rel = p - a.lead + a.rel_bias;
        const int row = rel / a.Wp;
        col = rel - row * a.Wp;
        pg = row / a.Hp;
        rin = row - pg * a.Hp;
        pg -= a.bias_pages;
      }
      int stage = wg;             // pi % n_stages and the parity of pi / n_stages (n_stages is 2 or 4)
      uint32_t stage_par = 0;
      if (a.out_f32) {
        // Small-C_out head (attention logit: ReLU; classifier: softmax / sigmoid over the classes, fp32) written
        // as dense float32 NHWC: lane = pixel, so a warp stores 32 consecutive pixels.
        const int C = a.f32_c;
        for (int pi = wg; pi < n_pass; pi += 2) {
          long long px[T];
          bool vt[T];
          {
            int c2 = col, r2 = rin, g2 = pg;
#pragma unroll
            for (int t = 0; t < T; ++t) {
              vt[t] = ((unsigned)(c2 - 1) < (unsigned)a.W) && ((unsigned)(r2 - 1) < (unsigned)a.H) && ((unsigned)g2 < (unsigned)a.N);
              px[t] = (((long long)g2 * a.H + (r2 - 1)) * a.W + (c2 - 1)) * C;
              walk_adv(c2, r2, g2, a.dc128, a.dr128, a.Wp, a.Hp);
            }
          }
          mbar_wait(s_tfull + 8 * stage, stage_par, a.err_flag, 4);
          __syncwarp();
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(stage * T * a.npad);
          uint32_t v[T][8];
#pragma unroll
          for (int t = 0; t < T; ++t) tmem_ld8(taddr + (uint32_t)(t * a.npad), v[t]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_tempty + 8 * stage);
#pragma unroll
          for (int t = 0; t < T; ++t) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = __uint_as_float(v[t][j]) + s_bias[j];
            if (a.act == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
            } else if (a.act == 2) {  // softmax over the C classes (layers.py:48-49)
              float mx = acc[0];
#pragma unroll
              for (int j = 1; j < 8; ++j) if (j < C) mx = fmaxf(mx, acc[j]);
              float sum = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) if (j < C) { acc[j] = expf(acc[j] - mx); sum += acc[j]; }
              const float inv = 1.f / sum;
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] *= inv;
            } else if (a.act == 3) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = 1.f / (1.f + expf(-acc[j]));
            }
            if (vt[t]) {
              float* o = a.out_f32 + px[t];
              if (C == 2) *reinterpret_cast<float2*>(o) = make_float2(acc[0], acc[1]);
              else if (C == 1) *o = acc[0];
              else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (j < C) o[j] = acc[j];
              }
            }
          }
          walk_adv(col, rin, pg, a.dcS, a.drS, a.Wp, a.Hp);
          stage += 2;
          if (stage >= a.n_stages) { stage -= a.n_stages; stage_par ^= 1; }
        }
      } else if (!a.deconv) {
        // One lane = one position of each of the T tiles of the pass (unrolled: T independent chains).  Every
        // position of the pass range is stored: in-image ones with the result, frame / margin ones with zero
        // (they are zero anyway), so there is no divergence and no bounds logic beyond three range checks.
        const bool has_res = a.res != nullptr, has_pre = a.out_pre != nullptr, relu = a.act == 1;
        for (int pi = wg; pi < n_pass; pi += 2) {
          long long pt[T];
          bool vt[T];
          {
            int c2 = col, r2 = rin, g2 = pg;
#pragma unroll
            for (int t = 0; t < T; ++t) {
              pt[t] = (long long)(p + t * TILE) * 8;
              vt[t] = ((unsigned)(c2 - 1) < (unsigned)a.W) && ((unsigned)(r2 - 1) < (unsigned)a.H) && ((unsigned)g2 < (unsigned)a.N);
              walk_adv(c2, r2, g2, a.dc128, a.dr128, a.Wp, a.Hp);
            }
          }
          // Residual vectors travel in a ring PD output chunks deep, filled before the accumulators are waited for and
          // refilled right after each use: with one load in flight per lane the epilogue of a T = 1 pass (C_out = 64:
          // eight chunks) was a chain of eight exposed memory latencies (measured 91 us against 52 us without residual).
          constexpr int PD = T == 1 ? 4 : 2;
          uint4 rr[PD][T];
          // ... and the residual of the pass this warpgroup handles NEXT (two units further) is pulled into L2 now: when
          // the accumulators are ready on arrival (the epilogue is the slower side of a residual launch) the first use
          // of rr below waits for its loads, and an L2 hit costs a third of a DRAM access (ncu: 37 % of the samples of
          // the 16 -> 16 + residual launch sat on that wait, profiles/r02z_ncu_conv_tc_16x16_res.txt)
          if (has_res && pi + 2 < n_pass && (lane & 7) == 0) {
#pragma unroll 1
            for (int d = 0; d < a.cout_chunks; ++d)
#pragma unroll
              for (int t = 0; t < T; ++t)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + (long long)d * a.res_plane * 8 + pt[t] + 2LL * UNIT * 8));
          }
          if (has_res) {
#pragma unroll
            for (int d = 0; d < PD; ++d)
              if (d < a.cout_chunks) {
#pragma unroll
                for (int t = 0; t < T; ++t)
                  rr[d][t] = __ldg(reinterpret_cast<const uint4*>(a.res + (long long)d * a.res_plane * 8 + pt[t]));
              }
          }
          mbar_wait(s_tfull + 8 * stage, stage_par, a.err_flag, 4);
          __syncwarp();
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(stage * T * a.npad);
#pragma unroll 1
          for (int c0 = 0; c0 < a.cout_chunks; c0 += PD) {
            // the TMEM reads of the whole group are issued together and waited for once
            uint32_t v[PD][T][8];
#pragma unroll
            for (int d = 0; d < PD; ++d)
              if (c0 + d < a.cout_chunks) {
#pragma unroll
                for (int t = 0; t < T; ++t) tmem_ld8(taddr + (uint32_t)(t * a.npad + (c0 + d) * 8), v[d][t]);
              }
            tmem_ld_wait();
            if (c0 + PD >= a.cout_chunks) {  // last TMEM read of this stage: hand the accumulators back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(s_tempty + 8 * stage);
            }
#pragma unroll
            for (int d = 0; d < PD; ++d) {
              const int c = c0 + d;
              if (c >= a.cout_chunks) break;
              const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c * 8 + 4);
#pragma unroll
              for (int t = 0; t < T; ++t) {
                float acc[8];
                acc[0] = __uint_as_float(v[d][t][0]) + b0.x; acc[1] = __uint_as_float(v[d][t][1]) + b0.y;
                acc[2] = __uint_as_float(v[d][t][2]) + b0.z; acc[3] = __uint_as_float(v[d][t][3]) + b0.w;
                acc[4] = __uint_as_float(v[d][t][4]) + b1.x; acc[5] = __uint_as_float(v[d][t][5]) + b1.y;
                acc[6] = __uint_as_float(v[d][t][6]) + b1.z; acc[7] = __uint_as_float(v[d][t][7]) + b1.w;
                if (has_res) {
                  float r[8];
                  unpack8(rr[d][t], r);
#pragma unroll
                  for (int j = 0; j < 8; ++j) acc[j] += r[j];
                }
                if (has_pre)
                  *reinterpret_cast<uint4*>(a.out_pre + (long long)c * a.pre_plane * 8 + pt[t]) = pack8_fin<false>(acc, vt[t]);
                const uint4 o = relu ? pack8_fin<true>(acc, vt[t]) : pack8_fin<false>(acc, vt[t]);
                *reinterpret_cast<uint4*>(a.out + (long long)c * a.out_plane * 8 + pt[t]) = o;
              }
              if (has_res && c + PD < a.cout_chunks) {
#pragma unroll
                for (int t = 0; t < T; ++t)
                  rr[d][t] = __ldg(reinterpret_cast<const uint4*>(a.res + (long long)(c + PD) * a.res_plane * 8 + pt[t]));
              }
            }
          }
          walk_adv(col, rin, pg, a.dcS, a.drS, a.Wp, a.Hp);
          p += 2 * UNIT;
          stage += 2;
          if (stage >= a.n_stages) { stage -= a.n_stages; stage_par ^= 1; }
        }
      } else {
        // transposed convolution: this lane's *input* pixel (j, i) = (rin-1, col-1), including the frame row /
        // column j = H, i = W; it owns the four outputs full[2j+py, 2i+px], cropped by (offy, offx)
        // (layers.py:342-367).  The (tile, channel chunk) steps of a pass are software-pipelined over two register
        // buffers: the TMEM reads of step i + 1 (four output-parity classes) are in flight while step i is converted and
        // stored - the read-out was a chain of exposed TMEM latencies (ncu: the first FADD / F2FP after every
        // tcgen05.wait::ld held 25 % of the kernel's samples).
        const int n_it = T * a.cls_chunks;
        for (int pi = wg; pi < n_pass; pi += 2) {
          mbar_wait(s_tfull + 8 * stage, stage_par, a.err_flag, 4);
          __syncwarp();
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(stage * T * a.npad);
          int c2 = col, r2 = rin, g2 = pg;
          bool v[4];
          long long po[4];
          auto tile_coords = [&]() {   // validity and output positions of the four classes of the current tile
            const bool in_dom = ((unsigned)g2 < (unsigned)a.N) && (c2 >= 1) && (c2 <= a.W + 1) && (r2 >= 1) && (r2 <= a.H + 1);
            const int oy0 = 2 * (r2 - 1) - a.offy, ox0 = 2 * (c2 - 1) - a.offx;
            const long long po0 = (long long)a.lead_o + ((long long)g2 * a.Hpo + oy0 + 1) * a.Wpo + ox0 + 1;
#pragma unroll
            for (int cls = 0; cls < 4; ++cls) {
              const int oy = oy0 + (cls >> 1), ox = ox0 + (cls & 1);
              v[cls] = in_dom && (oy >= 0) && (oy < a.Ho) && (ox >= 0) && (ox < a.Wo);
              po[cls] = po0 + (long long)(cls >> 1) * a.Wpo + (cls & 1);
            }
          };
          int t_ld = 0, c_ld = 0;   // the step whose TMEM reads are issued next
          auto issue = [&](uint32_t (&dst)[4][8]) {
#pragma unroll
            for (int cls = 0; cls < 4; ++cls)
              tmem_ld8(taddr + (uint32_t)(t_ld * a.npad + (cls * a.cls_chunks + c_ld) * 8), dst[cls]);
            if (++c_ld == a.cls_chunks) { c_ld = 0; ++t_ld; }
          };
          int c_cur = 0;
          auto finish = [&](const uint32_t (&src)[4][8]) {
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c_cur * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c_cur * 8 + 4);
#pragma unroll
            for (int cls = 0; cls < 4; ++cls) {
              if (v[cls]) {
                float acc[8];
                acc[0] = __uint_as_float(src[cls][0]) + b0.x; acc[1] = __uint_as_float(src[cls][1]) + b0.y;
                acc[2] = __uint_as_float(src[cls][2]) + b0.z; acc[3] = __uint_as_float(src[cls][3]) + b0.w;
                acc[4] = __uint_as_float(src[cls][4]) + b1.x; acc[5] = __uint_as_float(src[cls][5]) + b1.y;
                acc[6] = __uint_as_float(src[cls][6]) + b1.z; acc[7] = __uint_as_float(src[cls][7]) + b1.w;
                *reinterpret_cast<uint4*>(a.out + ((long long)c_cur * a.out_plane + po[cls]) * 8) =
                    a.act == 1 ? pack8_fin<true>(acc, true) : pack8_fin<false>(acc, true);
              }
            }
            if (++c_cur == a.cls_chunks) {   // next tile
              c_cur = 0;
              walk_adv(c2, r2, g2, a.dc128, a.dr128, a.Wp, a.Hp);
              tile_coords();
            }
          };
          auto release = [&]() {   // every TMEM read of this stage has completed: hand the accumulators back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tempty + 8 * stage);
          };
          uint32_t va[4][8], vb[4][8];
          tile_coords();
          issue(va);
#pragma unroll 1
          for (int it = 0; it < n_it; it += 2) {
            tmem_ld_wait();
            if (it + 1 < n_it) issue(vb); else release();
            finish(va);
            if (it + 1 < n_it) {
              tmem_ld_wait();
              if (it + 2 < n_it) issue(va); else release();
              finish(vb);
            }
          }
          walk_adv(col, rin, pg, a.dcS, a.drS, a.Wp, a.Hp);
          stage += 2;
          if (stage >= a.n_stages) { stage -= a.n_stages; stage_par ^= 1; }
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
}

struct Window {
  int min_off, max_off, dlo, dhi, win;
};

// KxK SAME correlation: taps reach PB positions back in both directions (TF pads 1 before for 3x3 and 4x4);
// the 2x2 form of the transposed convolution looks back only (PB = 1).
Window make_window(int ks, int cin_chunks, int Wp, int unit) {
  Window w;
  const int PB = ks == 2 ? 1 : (ks - 1) / 2;
  w.min_off = -PB * (Wp + 1);
  w.max_off = (ks - 1 - PB) * (Wp + 1);
  if (cin_chunks == 1) w.max_off += 1;  // paired taps read one position further
  w.dlo = -cdiv(-w.min_off, unit);
  w.dhi = (unit - 1 + w.max_off) / unit;
  w.win = w.dhi - w.dlo + 1;
  return w;
}

}  // namespace

// ---- host side -------------------------------------------------------------------------------
uint16_t host_f_to_act(float v) {
#ifdef ARU_USE_BF16
  uint32_t u;
  memcpy(&u, &v, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
#else
  if (v > 65504.f) v = 65504.f;
  if (v < -65504.f) v = -65504.f;
  __half h = __float2half_rn(v);
  uint16_t r;
  memcpy(&r, &h, 2);
  return r;
#endif
}

ConvTcPlan conv_tc_plan(int ks, int cin, int cout, const Geo& g, int num_sms, size_t max_smem, bool deconv) {
  ConvTcPlan p;
  p.deconv = deconv;
  p.ks = deconv ? 2 : ks;  // 3x3 stride-2 transposed conv == 2x2 correlation over the input with 4 parity classes in N
  p.cin_chunks = cdiv(cin, 8);
  p.cout_chunks = cdiv(cout, 8);
  p.npad = cdiv((deconv ? 4 : 1) * p.cout_chunks * 8, 16) * 16;
  if (deconv ? ks != 3 : (ks != 3 && ks != 4)) { p.why = "kernel size"; return p; }
  if (deconv && p.cin_chunks == 1) { p.why = "transposed conv needs >= 16 input channels"; return p; }
  ks = p.ks;
  if (p.cin_chunks != 1 && (p.cin_chunks & 1)) { p.why = "odd number of input chunks"; return p; }
  if (p.npad > 256) { p.why = "C_out > 256"; return p; }
  if (g.plane >= (1LL << 31) / 16) { p.why = "plane too large for 32-bit positions"; return p; }
  p.n_steps = p.cin_chunks == 1 ? ks * ((ks + 1) / 2) : ks * ks * (p.cin_chunks / 2);
  if (p.n_steps > MAX_STEPS) { p.why = "too many K steps"; return p; }
  p.wpack_bytes = (size_t)p.n_steps * 32 * p.npad;
  const size_t w_resident = ((p.wpack_bytes + 127) / 128) * 128;
  // streamed weight groups: about 16 KB, a divisor of the channel-chunk pairs per tap
  const int n_cp = p.cin_chunks == 1 ? 1 : p.cin_chunks / 2;
  p.w_group = std::max(1, std::min(n_cp, 16384 / (32 * p.npad)));
  while (n_cp % p.w_group) --p.w_group;
  p.n_wst = 3;
  const size_t w_streamed = (((size_t)p.n_wst * p.w_group * 32 * p.npad) + 127) / 128 * 128;
  p.why = "weights + operand ring exceed shared memory";
  const int combos[3][2] = {{0, PREFETCH_UNITS}, {1, PREFETCH_UNITS}, {1, 1}};  // {stream weights, prefetch units}
  for (int ci = 0; ci < 3 && !p.ok; ++ci) {
    const int stream = combos[ci][0], pf = combos[ci][1];
    for (int T = 4; T >= 1 && !p.ok; T >>= 1) {
      if (T * p.npad * 2 > 512) continue;  // at least two accumulator stages
      if (stream && T > 2) continue;
      const Window w = make_window(ks, p.cin_chunks, g.Wp, TILE * T);
      const int n_units = w.win + pf;
      if (n_units > MAX_UNITS || n_units * T * ks > TAB_MAX) continue;
      const size_t ring = (size_t)p.cin_chunks * (n_units * TILE * T + MIRROR) * 16;
      const size_t smem = HDR_BYTES + (stream ? w_streamed : w_resident) + ring;
      if (smem > max_smem) continue;
      p.ok = true;
      p.T = T;
      p.w_stream = stream;
      p.n_slots = n_units;
      p.smem_bytes = smem;
      p.n_stages = (T * p.npad * 4 <= 512) ? 4 : 2;
      // Deepen the prefetch with whatever shared memory is left: HBM latency under load is ~1-2 us, so a
      // CTA needs tens of KB of loads in flight to sustain its share (~45 GB/s) of the memory bandwidth.
      const size_t unit_bytes = (size_t)p.cin_chunks * TILE * T * 16;
      while (p.n_slots < MAX_UNITS && (p.n_slots + 1) * T * ks <= TAB_MAX && p.smem_bytes + unit_bytes <= max_smem &&
             (size_t)(p.n_slots + 1 - w.win) * unit_bytes <= 96 * 1024) {
        ++p.n_slots;
        p.smem_bytes += unit_bytes;
      }
    }
  }
  if (!p.ok) return p;
  p.why = "";
  const long long unit = TILE * p.T;
  p.tile_begin = g.pos(0, 0, 0) / unit;                          // in passes
  // the transposed conv also computes at the frame row / column (input pixel j = H, i = W feeds output 2H, 2W)
  p.tile_end = (deconv ? g.pos(g.N - 1, g.H, g.W) : g.pos(g.N - 1, g.H - 1, g.W - 1)) / unit + 1;
  const long long passes = p.tile_end - p.tile_begin;
  p.grid = (int)(passes < num_sms ? passes : num_sms);
  return p;
}

void conv_tc_pack_weights(const ConvTcPlan& p, const float* w, int cin, int cout, int ci_begin, uint16_t* dst) {
  const int ks = p.ks;
  memset(dst, 0, p.wpack_bytes);
  auto W = [&](int ky, int kx, int ci, int co) -> float {
    // ci is local to this K-part: channels [ci_begin, ci_begin + 8*cin_chunks) of the full filter
    if (kx >= ks || ci >= p.cin_chunks * 8 || ci_begin + ci >= cin || co >= cout) return 0.f;
    return w[(((size_t)ky * ks + kx) * cin + ci_begin + ci) * cout + co];
  };
  for (int s = 0; s < p.n_steps; ++s)
    for (int h = 0; h < 2; ++h)
      for (int n = 0; n < p.npad; ++n)
        for (int j = 0; j < 8; ++j) {
          float v;
          if (p.cin_chunks == 1) {
            const int npairs = (ks + 1) / 2;
            v = W(s / npairs, 2 * (s % npairs) + h, j, n);
          } else {
            const int cp = p.cin_chunks / 2;
            const int tap = s / cp;
            v = W(tap / ks, tap % ks, (2 * (s % cp) + h) * 8 + j, n);
          }
          dst[(((size_t)s * 2 + h) * p.npad + n) * 8 + j] = host_f_to_act(v);
        }
}

// Transposed-conv B image.  K = (2x2 look-back tap, input channel), N = (output parity class, output channel):
//   full[2j+py, 2i+px] = sum over (ky2, kx2) in {0,1}^2 of in[j-1+ky2, i-1+kx2] . Wd[ky][kx],
//   with ky = py (ky2 == 1: the pixel itself) or ky = 2 (ky2 == 0 and py == 0: the pixel above); same in x.
// Wd is the TF filter [3][3][cout][cin] of Conv2DBackpropInput (layers.py:356-360).
void conv_tc_pack_deconv(const ConvTcPlan& p, const float* w, int cin, int cout, uint16_t* dst) {
  memset(dst, 0, p.wpack_bytes);
  const int n_cp = p.cin_chunks / 2, cpad = p.cout_chunks * 8;
  auto tap_of = [](int k2, int par) -> int { return k2 == 1 ? par : (par == 0 ? 2 : -1); };
  for (int s = 0; s < p.n_steps; ++s) {
    const int tap2 = s / n_cp, cp = s % n_cp;
    const int ky2 = tap2 / 2, kx2 = tap2 % 2;
    for (int h = 0; h < 2; ++h)
      for (int cls = 0; cls < 4; ++cls) {
        const int ky = tap_of(ky2, cls >> 1), kx = tap_of(kx2, cls & 1);
        if (ky < 0 || kx < 0) continue;
        for (int co = 0; co < cout; ++co)
          for (int j = 0; j < 8; ++j) {
            const int ci = (2 * cp + h) * 8 + j;
            if (ci >= cin) continue;
            const float v = w[(((size_t)ky * 3 + kx) * cout + co) * cin + ci];
            dst[(((size_t)s * 2 + h) * p.npad + cls * cpad + co) * 8 + j] = host_f_to_act(v);
          }
      }
  }
}

static cudaError_t launch_tc_common(cudaStream_t st, const ConvTcPlan& p, ConvTcArgs& a, const Geo& g);

cudaError_t launch_deconv_tc(cudaStream_t st, const ConvTcPlan& p, PV in, PV out, const act_t* wpack, const float* bias_pad,
                             const Geo& gi, const Geo& go, int act, int* err_flag) {
  if (!p.ok || !p.deconv) return cudaErrorInvalidValue;
  ConvTcArgs a{};
  a.in = in.p; a.in_plane = in.plane;
  a.out = out.p; a.out_plane = out.plane;
  a.wpack = wpack; a.bias = bias_pad;
  a.act = act;
  a.err_flag = err_flag;
  a.deconv = 1;
  a.cls_chunks = p.cout_chunks;
  a.Wo = go.W; a.Ho = go.H; a.Wpo = go.Wp; a.Hpo = go.Hp; a.lead_o = (int)go.lead;
  a.offy = (2 * gi.H + 1 - go.H) / 2;
  a.offx = (2 * gi.W + 1 - go.W) / 2;
  return launch_tc_common(st, p, a, gi);
}

cudaError_t launch_conv_tc(cudaStream_t st, const ConvTcPlan& p, PV in, PV out, PV out_pre, PV res, const act_t* wpack,
                           const float* bias_pad, const Geo& g, int act, int* err_flag) {
  if (!p.ok || p.deconv) return cudaErrorInvalidValue;
  ConvTcArgs a{};
  a.in = in.p; a.in_plane = in.plane;
  a.out = out.p; a.out_plane = out.plane;
  a.out_pre = out_pre.p; a.pre_plane = out_pre.plane;
  a.res = res.p; a.res_plane = res.plane;
  a.wpack = wpack; a.bias = bias_pad;
  a.act = act;
  a.err_flag = err_flag;
  return launch_tc_common(st, p, a, g);
}

cudaError_t launch_conv_tc_f32(cudaStream_t st, const ConvTcPlan& p, PV in, float* out_f32, int cout, const act_t* wpack,
                               const float* bias_pad, const Geo& g, int act, int* err_flag) {
  if (!p.ok || p.deconv || cout < 1 || cout > 8 || p.cout_chunks != 1) return cudaErrorInvalidValue;
  ConvTcArgs a{};
  a.in = in.p; a.in_plane = in.plane;
  a.out_f32 = out_f32; a.f32_c = cout;
  a.wpack = wpack; a.bias = bias_pad;
  a.act = act;
  a.err_flag = err_flag;
  return launch_tc_common(st, p, a, g);
}

static cudaError_t launch_tc_common(cudaStream_t st, const ConvTcPlan& p, ConvTcArgs& a, const Geo& g) {
  a.ks = p.ks; a.cin_chunks = p.cin_chunks; a.cout_chunks = p.cout_chunks; a.npad = p.npad;
  a.n_steps = p.n_steps; a.n_units = p.n_slots; a.n_stages = p.n_stages;
  a.w_stream = p.w_stream; a.w_group = p.w_group; a.n_wst = p.n_wst;
  int cols = p.n_stages * p.T * p.npad;
  int pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  a.tmem_cols = pow2;
  a.W = g.W; a.H = g.H; a.Wp = g.Wp; a.Hp = g.Hp; a.N = g.N;
  a.lead = (int)g.lead;
  a.body_end = (int)(g.lead + (long long)g.N * g.Hp * g.Wp);
  const long long page = (long long)g.Wp * g.Hp;
  a.rel_bias = (int)(((g.lead + page - 1) / page) * page);
  a.pass_begin = (int)p.tile_begin; a.pass_end = (int)p.tile_end;
  const int unit = TILE * p.T;
  const Window w = make_window(p.ks, p.cin_chunks, g.Wp, unit);
  a.dlo = w.dlo; a.dhi = w.dhi; a.min_off = w.min_off;
  a.dc128 = TILE % g.Wp; a.dr128 = TILE / g.Wp;
  a.dcT = unit % g.Wp; a.drT = unit / g.Wp;
  a.dcS = (2 * unit) % g.Wp; a.drS = (2 * unit) / g.Wp;
  {
    // ring position of the A operand of (phase s, tile t, filter row ky); the tap column kx is added by the issuer
    // (the mirror behind the ring keeps start + kx + 127 contiguous)
    const long long L = (long long)p.n_slots * unit;
    const long long ws0 = ((((long long)(-w.dlo) * unit + w.min_off) % L) + L) % L;
    if (p.n_slots * p.T * p.ks > TAB_MAX) return cudaErrorInvalidValue;
    for (int s = 0; s < p.n_slots; ++s)
      for (int t = 0; t < p.T; ++t)
        for (int ky = 0; ky < p.ks; ++ky)
          a.tab[(s * p.T + t) * p.ks + ky] = (uint32_t)((ws0 + (long long)s * unit + (long long)t * TILE + (long long)ky * g.Wp) % L);
  }
  a.bias_pages = (int)(a.rel_bias / page);
  using Kern = void (*)(const __grid_constant__ ConvTcArgs);
  Kern k = nullptr;
  const int mode = p.cin_chunks == 1 ? 0 : (p.w_stream ? 2 : 1);
#define ARU_TC_PICK(TT, KK, MM) if (p.T == TT && p.ks == KK && mode == MM) k = k_conv_tc<TT, KK, MM>;
  ARU_TC_PICK(1, 3, 0) ARU_TC_PICK(2, 3, 0) ARU_TC_PICK(4, 3, 0)
  ARU_TC_PICK(1, 3, 1) ARU_TC_PICK(2, 3, 1) ARU_TC_PICK(4, 3, 1)
  ARU_TC_PICK(1, 3, 2) ARU_TC_PICK(2, 3, 2)
  ARU_TC_PICK(1, 4, 0) ARU_TC_PICK(2, 4, 0) ARU_TC_PICK(4, 4, 0)
  ARU_TC_PICK(1, 4, 1) ARU_TC_PICK(2, 4, 1) ARU_TC_PICK(4, 4, 1)
  ARU_TC_PICK(1, 4, 2) ARU_TC_PICK(2, 4, 2)
  ARU_TC_PICK(1, 2, 1) ARU_TC_PICK(2, 2, 1) ARU_TC_PICK(4, 2, 1)
  ARU_TC_PICK(1, 2, 2) ARU_TC_PICK(2, 2, 2)
#undef ARU_TC_PICK
  if (!k) return cudaErrorInvalidValue;
  static Kern configured[32];
  static int n_configured = 0;
  bool done = false;
  for (int i = 0; i < n_configured; ++i) done = done || configured[i] == k;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    if (n_configured < 32) configured[n_configured++] = k;
  }
  k<<<p.grid, NUM_THREADS, p.smem_bytes, st>>>(a);
  return cudaGetLastError();
}

}  // namespace aru
