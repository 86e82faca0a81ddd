// conv_tc.cu - tcgen05 / TMEM implicit-GEMM convolution for sm_100a.
//
// Computes  out = act(conv_KxK_SAME(in) + bias [+ res])  (layers.py:191-247, ARU_v1.py:212-227) on the
// flattened padded chunk-planar layout of aru_common.cuh.  In that layout the convolution is a 1-D
// correlation over positions:  out[p] = sum_{tap} W[tap] . in[p + off(tap)],  off = (ky-1)*Wp + (kx-1).
//
// GEMM view per tile:  D[128 positions, N=C_out] = sum over K-steps of A[128, 16] * B[16, N]
//   * A (activations): for tap t and channel chunks (2c, 2c+1) the 128 x 16 operand is two contiguous
//     runs of 128 x 16 B in shared memory -> canonical no-swizzle K-major core matrices,
//     SBO = 128 B (next 8 positions), LBO = ring plane stride (next 8 channels).  For C_in = 8 two
//     horizontally adjacent taps form one K=16 step (LBO = 16 B: the same run shifted by one position).
//   * B (weights): pre-packed on the host into the exact shared-memory image, resident for the whole
//     kernel (weight-stationary persistent CTAs).
//   * D: fp32 in TMEM, n_stages accumulator stages so the epilogue of tile i overlaps the MMAs of i+1..
//
// Warp roles (384 threads, 1 CTA / SM, persistent over a contiguous range of tiles):
//   warp 0 lane 0 : producer - streams the input as 2 KB bulk copies (TMA engine, cp.async.bulk) into a
//                   ring of 128-position slots; slot 0 is mirrored behind the ring so that every
//                   128(+1)-position operand run is contiguous even when it wraps
//   warp 1 lane 0 : MMA issuer - tcgen05.mma.cta_group::1.kind::f16, commits free ring slots / publish D
//   warp 2        : TMEM allocator
//   warps 4..11   : two epilogue warpgroups (alternating tiles): tcgen05.ld -> +bias (+res) -> act ->
//                   16-bit pack -> 16 B coalesced global stores (in-image positions only)
// All mbarrier waits are bounded: a protocol bug sets *err_flag and drains instead of hanging the GPU.
#include <cstring>
#include <vector>

#include "aru_common.cuh"
#include "kernels.h"

namespace aru {

namespace {

constexpr int TILE = 128;
constexpr int NUM_THREADS = 384;
constexpr int MAX_SLOTS = 128;
constexpr int MAX_STEPS = 256;
constexpr int PREFETCH_SLOTS = 3;
// shared-memory header layout (bytes)
constexpr int OFF_FULL = 0;                       // uint64 full[MAX_SLOTS]
constexpr int OFF_EMPTY = OFF_FULL + 8 * MAX_SLOTS;
constexpr int OFF_TFULL = OFF_EMPTY + 8 * MAX_SLOTS;  // uint64 tmem_full[4]
constexpr int OFF_TEMPTY = OFF_TFULL + 32;            // uint64 tmem_empty[4]
constexpr int OFF_TMEMPTR = OFF_TEMPTY + 32;          // uint32 tmem base, uint32 abort flag
constexpr int OFF_STEPOFF = OFF_TMEMPTR + 16;         // int step_off[MAX_STEPS]
constexpr int OFF_STEPCH = OFF_STEPOFF + 4 * MAX_STEPS;  // int step_chunk[MAX_STEPS]
constexpr int OFF_BIAS = OFF_STEPCH + 4 * MAX_STEPS;     // float bias[256]
constexpr int HDR_BYTES = ((OFF_BIAS + 4 * 256 + 127) / 128) * 128;

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
  int ks, cin_chunks, cout_chunks, npad, n_steps, n_slots, n_stages, tmem_cols;
  int N, H, W, Wp, Hp;
  int lead;
  int tile_begin, tile_end;
  int dlo, dhi;
  int act;
  int* err_flag;
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
// Bounded wait: returns false (and raises the abort flags) when the barrier did not flip within 0.2 s.
__device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, volatile uint32_t* abort_smem, int* err_flag,
                                            int code) {
  const unsigned long long t0 = global_ns();
  while (true) {
    for (int i = 0; i < 64; ++i)
      if (mbar_try_wait(bar, parity)) return true;
    if (*abort_smem) return false;
    if (global_ns() - t0 > 200000000ull) {
      *abort_smem = 1;
      atomicCAS(err_flag, 0, code);
      return false;
    }
  }
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile uint32_t* abort_smem, int* err_flag,
                                          int code) {
  if (mbar_try_wait(bar, parity)) return true;
  return mbar_wait_slow(bar, parity, abort_smem, err_flag, code);
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

// shared-memory matrix descriptor, no swizzle, K-major: start / LBO / SBO in 16-byte units, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// ---- the kernel --------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv_tc(const ConvTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_full = s_base + OFF_FULL, s_empty = s_base + OFF_EMPTY;
  const uint32_t s_tfull = s_base + OFF_TFULL, s_tempty = s_base + OFF_TEMPTY;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEMPTR);
  volatile uint32_t* abort_smem = tmem_ptr + 1;
  int* step_off = reinterpret_cast<int*>(smem + OFF_STEPOFF);
  int* step_chunk = reinterpret_cast<int*>(smem + OFF_STEPCH);
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  const int wbytes = a.n_steps * 32 * a.npad;
  const uint32_t s_w = s_base + HDR_BYTES;
  const uint32_t s_ring = s_w + ((wbytes + 127) / 128) * 128;
  const int L = a.n_slots * TILE;        // ring length in positions
  const int ring_plane = L + TILE;       // + mirror of slot 0

  // static partition of the tile range over the persistent CTAs
  const int total_tiles = a.tile_end - a.tile_begin;
  const int per_cta = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t0 = a.tile_begin + (int)blockIdx.x * per_cta;
  const int t1 = min(t0 + per_cta, a.tile_end);
  const int n_tiles = max(t1 - t0, 0);
  const int win = a.dhi - a.dlo + 1;     // chunks a tile reads

  // a previous launch hit a protocol timeout: do nothing (the host reports the error after the pass)
  if (*reinterpret_cast<volatile int*>(a.err_flag) != 0) return;

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    *abort_smem = 0;
    for (int i = 0; i < a.n_slots; ++i) {
      mbar_init(s_full + 8 * i, 1);
      mbar_init(s_empty + 8 * i, 1);
    }
    for (int i = 0; i < a.n_stages; ++i) {
      mbar_init(s_tfull + 8 * i, 1);
      mbar_init(s_tempty + 8 * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {  // K-step table: offset (positions) and first channel chunk of each step
    const int PB = (a.ks - 1) / 2;
    for (int s = threadIdx.x; s < a.n_steps; s += NUM_THREADS) {
      int off, ch;
      if (a.cin_chunks == 1) {
        const int npairs = (a.ks + 1) / 2;
        const int ky = s / npairs, pi = s % npairs;
        off = (ky - PB) * a.Wp + (2 * pi - PB);
        ch = 0;
      } else {
        const int cp = a.cin_chunks / 2;
        const int tap = s / cp;
        off = (tap / a.ks - PB) * a.Wp + (tap % a.ks - PB);
        ch = 2 * (s % cp);
      }
      step_off[s] = off;
      step_chunk[s] = ch;
    }
    for (int i = threadIdx.x; i < a.npad; i += NUM_THREADS) s_bias[i] = a.bias[i];
    // resident B image
    const uint4* wsrc = reinterpret_cast<const uint4*>(a.wpack);
    uint4* wdst = reinterpret_cast<uint4*>(smem + HDR_BYTES);
    for (int i = threadIdx.x; i < wbytes / 16; i += NUM_THREADS) wdst[i] = __ldg(wsrc + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA) reads
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

  if (n_tiles > 0) {
    if (warp == 0 && lane == 0) {
      // ================= producer =================
      const int n_chunks = n_tiles + win - 1;
      const long long first_pos = (long long)(t0 + a.dlo) * TILE;  // ring origin in plane positions
      const uint32_t tx = (uint32_t)a.cin_chunks * TILE * 16;
      for (int k = 0; k < n_chunks; ++k) {
        const int slot = k % a.n_slots, use = k / a.n_slots;
        if (!mbar_wait(s_empty + 8 * slot, (use & 1) ^ 1, abort_smem, a.err_flag, 1)) break;
        const uint32_t bar = s_full + 8 * slot;
        mbar_expect_tx(bar, slot == 0 ? 2 * tx : tx);
        const act_t* src = a.in + (first_pos + (long long)k * TILE) * 8;
        for (int c = 0; c < a.cin_chunks; ++c) {
          const uint32_t dst = s_ring + (uint32_t)(c * ring_plane + slot * TILE) * 16;
          bulk_g2s(dst, src + c * a.in_plane * 8, TILE * 16, bar);
          if (slot == 0) bulk_g2s(dst + (uint32_t)L * 16, src + c * a.in_plane * 8, TILE * 16, bar);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ================= MMA issuer =================
      const uint32_t idesc = (1u << 4) | (ARU_UMMA_FMT << 7) | (ARU_UMMA_FMT << 10) | ((uint32_t)(a.npad >> 3) << 17) |
                             ((uint32_t)(TILE >> 4) << 24);
      const uint32_t a_lbo = a.cin_chunks == 1 ? 16u : (uint32_t)ring_plane * 16u;
      const uint32_t b_lbo = (uint32_t)a.npad * 16u;
      int tile_ring = (-a.dlo * TILE) % L;  // ring position of the tile's own position 0
      bool alive = true;
      for (int ti = 0; ti < n_tiles && alive; ++ti) {
        const int stage = ti % a.n_stages;
        alive = mbar_wait(s_tempty + 8 * stage, ((ti / a.n_stages) & 1) ^ 1, abort_smem, a.err_flag, 2);
        if (!alive) break;
        // operands: tile ti reads chunks ti .. ti+win-1 (the first tile waits for its whole window)
        for (int k = (ti == 0 ? 0 : ti + win - 1); k <= ti + win - 1; ++k) {
          alive = mbar_wait(s_full + 8 * (k % a.n_slots), (k / a.n_slots) & 1, abort_smem, a.err_flag, 3);
          if (!alive) break;
        }
        if (!alive) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(stage * a.npad);
        for (int s = 0; s < a.n_steps; ++s) {
          int rp = tile_ring + step_off[s];
          rp += (rp < 0) ? L : 0;
          rp -= (rp >= L) ? L : 0;
          const uint32_t a_addr = s_ring + (uint32_t)(step_chunk[s] * ring_plane + rp) * 16u;
          const uint32_t b_addr = s_w + (uint32_t)s * 32u * (uint32_t)a.npad;
          umma_f16(d_tmem, make_desc(a_addr, a_lbo, 128), make_desc(b_addr, b_lbo, 128), idesc, s > 0 ? 1u : 0u);
        }
        umma_commit(s_empty + 8 * (ti % a.n_slots));  // chunk ti is not needed by later tiles
        umma_commit(s_tfull + 8 * stage);             // accumulator ready for the epilogue
        tile_ring += TILE;
        tile_ring -= (tile_ring >= L) ? L : 0;
      }
    } else if (warp >= 4) {
      // ================= epilogue =================
      const int wg = (warp - 4) >> 2;  // warpgroup 0/1 -> even/odd tiles
      const int q = warp & 3;          // TMEM lane quarter this warp may access
      const int m = q * 32 + lane;
      const int body_rows = a.N * a.Hp;
      bool alive = true;
      for (int ti = wg; ti < n_tiles && alive; ti += 2) {
        const int stage = ti % a.n_stages;
        const int p = (t0 + ti) * TILE + m;
        const int rel = p - a.lead;
        bool valid = false;
        if (rel >= 0) {
          const int row = rel / a.Wp, col = rel - row * a.Wp;
          const int rin = row % a.Hp;
          valid = (col >= 1) && (col <= a.W) && (rin >= 1) && (rin <= a.H) && (row < body_rows);
        }
        alive = mbar_wait(s_tfull + 8 * stage, (ti / a.n_stages) & 1, abort_smem, a.err_flag, 4);
        alive = __shfl_sync(0xffffffffu, alive ? 1 : 0, 0) != 0;
        if (!alive) break;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(stage * a.npad);
        for (int c0 = 0; c0 < a.cout_chunks; c0 += 2) {
          uint32_t v0[8], v1[8];
          const bool two = (c0 + 1 < a.cout_chunks);
          tmem_ld8(taddr + c0 * 8, v0);
          if (two) tmem_ld8(taddr + c0 * 8 + 8, v1);
          tmem_ld_wait();
          if (c0 + 2 >= a.cout_chunks) {  // last TMEM read of this stage: hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_tempty + 8 * stage);
          }
          if (valid) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 1 && !two) break;
              const int c = c0 + h;
              float acc[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = __uint_as_float(h ? v1[j] : v0[j]) + s_bias[c * 8 + j];
              if (a.res) {
                float r[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(a.res + ((long long)c * a.res_plane + p) * 8)), r);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += r[j];
              }
              if (a.out_pre)
                *reinterpret_cast<uint4*>(a.out_pre + ((long long)c * a.pre_plane + p) * 8) = pack8(acc);
              if (a.act == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
              }
              *reinterpret_cast<uint4*>(a.out + ((long long)c * a.out_plane + p) * 8) = pack8(acc);
            }
          }
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

ConvTcPlan conv_tc_plan(int ks, int cin, int cout, const Geo& g, int num_sms, size_t max_smem) {
  ConvTcPlan p;
  p.ks = ks;
  p.cin_chunks = cdiv(cin, 8);
  p.cout_chunks = cdiv(cout, 8);
  p.npad = cdiv(p.cout_chunks * 8, 16) * 16;
  if (ks != 3 && ks != 4) { p.why = "kernel size"; return p; }
  if (p.cin_chunks != 1 && (p.cin_chunks & 1)) { p.why = "odd number of input chunks"; return p; }
  if (p.npad > 256) { p.why = "C_out > 256"; return p; }
  if (g.plane >= (1LL << 31) / 16) { p.why = "plane too large for 32-bit positions"; return p; }
  const int PB = (ks - 1) / 2;
  const int min_off = -PB * (g.Wp + 1);
  int max_off = (ks - 1 - PB) * (g.Wp + 1);
  if (p.cin_chunks == 1) max_off += 1;  // paired taps read one position further
  const int dlo = -cdiv(-min_off, TILE);
  const int dhi = (TILE - 1 + max_off) / TILE;
  const int win = dhi - dlo + 1;
  p.n_slots = win + PREFETCH_SLOTS;
  if (p.n_slots > MAX_SLOTS) { p.why = "row pitch too large for the operand ring"; return p; }
  p.n_steps = p.cin_chunks == 1 ? ks * ((ks + 1) / 2) : ks * ks * (p.cin_chunks / 2);
  if (p.n_steps > MAX_STEPS) { p.why = "too many K steps"; return p; }
  p.n_stages = p.npad <= 64 ? 4 : 2;
  p.wpack_bytes = (size_t)p.n_steps * 32 * p.npad;
  const size_t ring = (size_t)p.cin_chunks * (p.n_slots * TILE + TILE) * 16;
  p.smem_bytes = HDR_BYTES + ((p.wpack_bytes + 127) / 128) * 128 + ring;
  if (p.smem_bytes > max_smem) { p.why = "weights + operand ring exceed shared memory"; return p; }
  p.tile_begin = g.pos(0, 0, 0) / TILE;
  p.tile_end = g.pos(g.N - 1, g.H - 1, g.W - 1) / TILE + 1;
  const long long tiles = p.tile_end - p.tile_begin;
  p.grid = (int)(tiles < num_sms ? tiles : num_sms);
  p.ok = true;
  return p;
}

void conv_tc_pack_weights(const ConvTcPlan& p, const float* w, int cin, int cout, uint16_t* dst) {
  const int ks = p.ks;
  memset(dst, 0, p.wpack_bytes);
  auto W = [&](int ky, int kx, int ci, int co) -> float {
    if (kx >= ks || ci >= cin || co >= cout) return 0.f;
    return w[(((size_t)ky * ks + kx) * cin + ci) * cout + co];
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

cudaError_t launch_conv_tc(cudaStream_t st, const ConvTcPlan& p, PV in, PV out, PV out_pre, PV res, const act_t* wpack,
                           const float* bias_pad, const Geo& g, int act, int* err_flag) {
  if (!p.ok) return cudaErrorInvalidValue;
  ConvTcArgs a;
  a.in = in.p; a.in_plane = in.plane;
  a.out = out.p; a.out_plane = out.plane;
  a.out_pre = out_pre.p; a.pre_plane = out_pre.plane;
  a.res = res.p; a.res_plane = res.plane;
  a.wpack = wpack; a.bias = bias_pad;
  a.ks = p.ks; a.cin_chunks = p.cin_chunks; a.cout_chunks = p.cout_chunks; a.npad = p.npad;
  a.n_steps = p.n_steps; a.n_slots = p.n_slots; a.n_stages = p.n_stages;
  int cols = p.n_stages * p.npad;
  int pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  a.tmem_cols = pow2;
  a.N = g.N; a.H = g.H; a.W = g.W; a.Wp = g.Wp; a.Hp = g.Hp; a.lead = (int)g.lead;
  a.tile_begin = (int)p.tile_begin; a.tile_end = (int)p.tile_end;
  const int PB = (p.ks - 1) / 2;
  const int min_off = -PB * (g.Wp + 1);
  int max_off = (p.ks - 1 - PB) * (g.Wp + 1);
  if (p.cin_chunks == 1) max_off += 1;
  a.dlo = -cdiv(-min_off, TILE);
  a.dhi = (TILE - 1 + max_off) / TILE;
  a.act = act;
  a.err_flag = err_flag;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_conv_tc<<<p.grid, NUM_THREADS, p.smem_bytes, st>>>(a);
  return cudaGetLastError();
}

}  // namespace aru
