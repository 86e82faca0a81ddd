// kernels_simple.cu - CUDA-core kernels of the ARU-Net forward pass.
//
// Two groups:
//  (1) the memory-bound ops that stay on CUDA cores by design (HBM roofline): 1-channel stem convs,
//      small-C_out convs (attention logit, classifier + softmax), pools, the fused
//      upsample+softmax-over-scales+weighted-sum ("combine"), uint8/mask quantisation;
//  (2) a direct convolution / transposed convolution on CUDA cores that covers every channel count.
//      It is the validation twin of the tcgen05 implicit-GEMM kernels (conv_tc.cu): tests run both
//      against the CPU oracle; ARU_OPT_CONV_PATH=1 forces it.
// Semantics follow TF's SAME rule; reference call sites are cited per kernel.
#include "aru_common.cuh"
#include "kernels.h"

namespace aru {

static __device__ __forceinline__ float apply_act(float v, int act) { return act == 1 ? fmaxf(v, 0.f) : v; }

// ---------------------------------------------------------------------------------------------
// Stem conv: C_in = 1 float32 plane -> C_out <= 16 chunk-planar.  layers.py:191-247 with the 1-channel
// inputs of ARU_v1.py:173 (attention conv1, 4x4) and :212 (unet_down_0/conv1, 3x3).  HBM-bound.
// ---------------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(256) k_conv_stem(const float* __restrict__ in, act_t* __restrict__ out,
                                                   act_t* __restrict__ out_pre, const float* __restrict__ w,
                                                   const float* __restrict__ bias, int N, int H, int W, int chunks,
                                                   long long plane, int act) {
  __shared__ float sw[KS * KS * 16];
  __shared__ float sb[16];
  const int cpad = chunks * 8;
  for (int i = threadIdx.x; i < KS * KS * cpad; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < cpad; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const long long total = (long long)N * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int x = (int)(p % W);
  const int y = (int)((p / W) % H);
  const float* img = in + (p - (long long)y * W - x);
  constexpr int PB = (KS - 1) / 2;
  float v[KS * KS];
#pragma unroll
  for (int ky = 0; ky < KS; ++ky)
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) {
      const int yy = y + ky - PB, xx = x + kx - PB;
      v[ky * KS + kx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (long long)yy * W + xx) : 0.f;
    }
  for (int c = 0; c < chunks; ++c) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = sb[c * 8 + j];
#pragma unroll
    for (int t = 0; t < KS * KS; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[t], sw[t * cpad + c * 8 + j], acc[j]);
    if (out_pre) *reinterpret_cast<uint4*>(out_pre + c * plane + p * 8) = pack8(acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = apply_act(acc[j], act);
    *reinterpret_cast<uint4*>(out + c * plane + p * 8) = pack8(acc);
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
                                                     int N, int H, int W, int act) {
  const long long total = (long long)N * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int coc = blockIdx.y;
  if (p >= total) return;
  const int x = (int)(p % W);
  const int y = (int)((p / W) % H);
  constexpr int PB = (KS - 1) / 2;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = __ldg(bias + coc * 8 + j);
  for (int ky = 0; ky < KS; ++ky) {
    const int yy = y + ky - PB;
    if (yy < 0 || yy >= H) continue;
    for (int kx = 0; kx < KS; ++kx) {
      const int xx = x + kx - PB;
      if (xx < 0 || xx >= W) continue;
      const long long q = p + (long long)(ky - PB) * W + (kx - PB);
      const int tap = ky * KS + kx;
      for (int cic = 0; cic < cin_chunks; ++cic) {
        const uint4 xin = __ldg(reinterpret_cast<const uint4*>(in + cic * in_plane + q * 8));
        const uint4* wp = reinterpret_cast<const uint4*>(w) + ((long long)(tap * cin_chunks + cic) * cout_chunks + coc) * 8;
        mac8x8(xin, wp, acc);
      }
    }
  }
  if (res) {
    float r[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(res + coc * res_plane + p * 8)), r);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += r[j];
  }
  if (out_pre) *reinterpret_cast<uint4*>(out_pre + coc * pre_plane + p * 8) = pack8(acc);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = apply_act(acc[j], act);
  *reinterpret_cast<uint4*>(out + coc * out_plane + p * 8) = pack8(acc);
}

// ---------------------------------------------------------------------------------------------
// Small-C_out conv to float32: attention logit (4x4, 32->1, ReLU; ARU_v1.py:182-183) and the classifier
// (4x4, 8->n_class, identity + output Softmax/Sigmoid; ARU_v1.py:158-160 + exporter's 'output' op,
// contract helper.py:70).  Output is float32 NHWC [N,H,W,cout] (cout == 1 -> a plane).
// Weights float32 [tap][cin_pad][cout].  Optionally also emits the consumers' uint8 / mask forms
// (separator_net_post_processor.py:147-149, helper.py:75-78).
// ---------------------------------------------------------------------------------------------
template <int KS, int COUT>
__global__ void __launch_bounds__(128) k_conv_small(const act_t* __restrict__ in, long long in_plane, int cin_chunks,
                                                    float* __restrict__ out, const float* __restrict__ w,
                                                    const float* __restrict__ bias, int N, int H, int W, int act) {
  extern __shared__ float sw[];  // [KS*KS][cin_chunks*8][COUT]
  const int nw = KS * KS * cin_chunks * 8 * COUT;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long total = (long long)N * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int x = (int)(p % W);
  const int y = (int)((p / W) % H);
  constexpr int PB = (KS - 1) / 2;
  float acc[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) acc[j] = __ldg(bias + j);
  for (int ky = 0; ky < KS; ++ky) {
    const int yy = y + ky - PB;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) {
      const int xx = x + kx - PB;
      if (xx < 0 || xx >= W) continue;
      const long long q = p + (long long)(ky - PB) * W + (kx - PB);
      const float* wt = sw + (ky * KS + kx) * cin_chunks * 8 * COUT;
      for (int cic = 0; cic < cin_chunks; ++cic) {
        float xi[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(in + cic * in_plane + q * 8)), xi);
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
  for (int j = 0; j < COUT; ++j) out[p * COUT + j] = acc[j];
}

// ---------------------------------------------------------------------------------------------
// Transposed conv 3x3 stride 2 SAME (layers.py:342-367: tf.nn.conv2d_transpose, filter [k,k,Cout,Cin]).
// full[2*iy+ky, 2*ix+kx] += in[iy,ix] * W[ky,kx];  out = full[oy:oy+Ho, ox:ox+Wo],
// oy = (2*Hi + 1 - Ho) / 2  (0 when Ho even, 1 when odd).  Weights packed like k_conv_direct.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_deconv_direct(const act_t* __restrict__ in, long long in_plane, int cin_chunks,
                                                       act_t* __restrict__ out, long long out_plane, int cout_chunks,
                                                       const act_t* __restrict__ w, const float* __restrict__ bias,
                                                       int N, int Hi, int Wi, int Ho, int Wo, int act) {
  const long long total = (long long)N * Ho * Wo;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int coc = blockIdx.y;
  if (p >= total) return;
  const int x = (int)(p % Wo);
  const int y = (int)((p / Wo) % Ho);
  const int n = (int)(p / ((long long)Wo * Ho));
  const int oy = (2 * Hi + 1 - Ho) / 2, ox = (2 * Wi + 1 - Wo) / 2;
  const int fy = y + oy, fx = x + ox;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = __ldg(bias + coc * 8 + j);
  for (int ky = (fy & 1); ky < 3; ky += 2) {
    const int iy = (fy - ky) >> 1;
    if (iy < 0 || iy >= Hi) continue;
    for (int kx = (fx & 1); kx < 3; kx += 2) {
      const int ix = (fx - kx) >> 1;
      if (ix < 0 || ix >= Wi) continue;
      const long long q = ((long long)n * Hi + iy) * Wi + ix;
      const int tap = ky * 3 + kx;
      for (int cic = 0; cic < cin_chunks; ++cic) {
        const uint4 xin = __ldg(reinterpret_cast<const uint4*>(in + cic * in_plane + q * 8));
        const uint4* wp = reinterpret_cast<const uint4*>(w) + ((long long)(tap * cin_chunks + cic) * cout_chunks + coc) * 8;
        mac8x8(xin, wp, acc);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = apply_act(acc[j], act);
  *reinterpret_cast<uint4*>(out + coc * out_plane + p * 8) = pack8(acc);
}

// ---------------------------------------------------------------------------------------------
// 2x2 stride-2 SAME pools (layers.py:526-544).  out = ceil(in/2); the extra row/col of odd sizes sees
// -inf (max) or is left out of the divisor (avg).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_maxpool(const act_t* __restrict__ in, long long in_plane,
                                                 act_t* __restrict__ out, long long out_plane, int N, int Hi, int Wi,
                                                 int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (p >= total) return;
  const int x = (int)(p % Wo);
  const int y = (int)((p / Wo) % Ho);
  const int n = (int)(p / ((long long)Wo * Ho));
  const act_t* src = in + c * in_plane + (((long long)n * Hi + 2 * y) * Wi + 2 * x) * 8;
  const bool hx = 2 * x + 1 < Wi, hy = 2 * y + 1 < Hi;
  float m[8], t[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(src)), m);
  if (hx) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + 8)), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], t[j]);
  }
  if (hy) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + (long long)Wi * 8)), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], t[j]);
    if (hx) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + (long long)Wi * 8 + 8)), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], t[j]);
    }
  }
  *reinterpret_cast<uint4*>(out + c * out_plane + p * 8) = pack8(m);
}

__global__ void __launch_bounds__(256) k_avgpool(const act_t* __restrict__ in, long long in_plane,
                                                 act_t* __restrict__ out, long long out_plane, int N, int Hi, int Wi,
                                                 int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (p >= total) return;
  const int x = (int)(p % Wo);
  const int y = (int)((p / Wo) % Ho);
  const int n = (int)(p / ((long long)Wo * Ho));
  const act_t* src = in + c * in_plane + (((long long)n * Hi + 2 * y) * Wi + 2 * x) * 8;
  const bool hx = 2 * x + 1 < Wi, hy = 2 * y + 1 < Hi;
  float s[8], t[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(src)), s);
  int cnt = 1;
  if (hx) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + 8)), t);
    for (int j = 0; j < 8; ++j) s[j] += t[j];
    ++cnt;
  }
  if (hy) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + (long long)Wi * 8)), t);
    for (int j = 0; j < 8; ++j) s[j] += t[j];
    ++cnt;
    if (hx) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + (long long)Wi * 8 + 8)), t);
      for (int j = 0; j < 8; ++j) s[j] += t[j];
      ++cnt;
    }
  }
  const float inv = 1.f / cnt;
  for (int j = 0; j < 8; ++j) s[j] *= inv;
  *reinterpret_cast<uint4*>(out + c * out_plane + p * 8) = pack8(s);
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
    // same summation order as a row-major 2x2 window walk
    int cnt = 1;
    if (hx) { a += src[1]; ++cnt; }
    if (hy) { a += src[Wi]; ++cnt; }
    if (hx && hy) { a += src[Wi + 1]; ++cnt; }
    a = a / cnt;
  }
  out[p] = a;
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
struct CombineArgs {
  const float* att[8];
  const act_t* det[8];
  long long det_plane[8];
  int att_h[8], att_w[8], att_up[8], att_oy[8], att_ox[8];
  int det_h[8], det_w[8], det_up[8], det_oy[8], det_ox[8], det_chunks[8];
  int A;
  act_t* out;
  long long out_plane;
  int out_chunks;
  int N, H, W;
};

__global__ void __launch_bounds__(256) k_combine(const CombineArgs a) {
  const long long total = (long long)a.N * a.H * a.W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int x = (int)(p % a.W);
  const int y = (int)((p / a.W) % a.H);
  const int n = (int)(p / ((long long)a.W * a.H));
  float s[8];
  float m = -INFINITY;
  for (int k = 0; k < a.A; ++k) {
    const int sy = (y + a.att_oy[k]) / a.att_up[k], sx = (x + a.att_ox[k]) / a.att_up[k];
    s[k] = __ldg(a.att[k] + ((long long)n * a.att_h[k] + sy) * a.att_w[k] + sx);
    m = fmaxf(m, s[k]);
  }
  float den = 0.f;
  for (int k = 0; k < a.A; ++k) {
    s[k] = expf(s[k] - m);
    den += s[k];
  }
  const float inv = 1.f / den;
  for (int oc = 0; oc < a.out_chunks; ++oc) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int k = 0; k < a.A; ++k) {
      const float wk = s[k] * inv;
      if (a.det_up[k] == 1) {
        float d[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(a.det[k] + oc * a.det_plane[k] + p * 8)), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(d[j], wk, acc[j]);
      } else {
        const int sy = (y + a.det_oy[k]) / a.det_up[k], sx = (x + a.det_ox[k]) / a.det_up[k];
        const long long q = ((long long)n * a.det_h[k] + sy) * a.det_w[k] + sx;
        float sum = 0.f;
        for (int c = 0; c < a.det_chunks[k]; ++c) {
          float d[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(a.det[k] + c * a.det_plane[k] + q * 8)), d);
#pragma unroll
          for (int j = 0; j < 8; ++j) sum += d[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(sum, wk, acc[j]);
      }
    }
    *reinterpret_cast<uint4*>(a.out + oc * a.out_plane + p * 8) = pack8(acc);
  }
}

// stand-alone upsample_simple (only reached by graphs that use it outside the attention pattern)
__global__ void __launch_bounds__(256) k_upsum(const act_t* __restrict__ in, long long in_plane, int in_chunks, int Hi,
                                               int Wi, act_t* __restrict__ out, long long out_plane, int out_chunks,
                                               int out_c, int N, int H, int W, int up, int oy, int ox) {
  const long long total = (long long)N * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int x = (int)(p % W);
  const int y = (int)((p / W) % H);
  const int n = (int)(p / ((long long)W * H));
  const long long q = ((long long)n * Hi + (y + oy) / up) * Wi + (x + ox) / up;
  float sum = 0.f;
  for (int c = 0; c < in_chunks; ++c) {
    float d[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(in + c * in_plane + q * 8)), d);
    for (int j = 0; j < 8; ++j) sum += d[j];
  }
  for (int oc = 0; oc < out_chunks; ++oc) {
    float v[8];
    for (int j = 0; j < 8; ++j) v[j] = (oc * 8 + j < out_c) ? sum : 0.f;
    *reinterpret_cast<uint4*>(out + oc * out_plane + p * 8) = pack8(v);
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
                                                  uint8_t* __restrict__ mask, long long npix, int C, float thr255) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  uint8_t first = 0;
  for (int c = 0; c < C; ++c) {
    // numpy: np.array(p*255, dtype=np.uint8) - float32 multiply, truncation toward zero
    const float v = prob[p * C + c] * 255.f;
    const uint8_t q = (uint8_t)(int)v;
    if (c == 0) first = q;
    if (u8) u8[p * C + c] = q;
  }
  if (mask) mask[p] = ((float)first > thr255) ? 255 : 0;
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

cudaError_t launch_conv_stem(cudaStream_t st, int ks, const float* in, act_t* out, act_t* out_pre, const float* w,
                             const float* bias, int N, int H, int W, int chunks, long long plane, int act) {
  const long long total = (long long)N * H * W;
  if (ks == 3)
    k_conv_stem<3><<<blocks_for(total, 256), 256, 0, st>>>(in, out, out_pre, w, bias, N, H, W, chunks, plane, act);
  else
    k_conv_stem<4><<<blocks_for(total, 256), 256, 0, st>>>(in, out, out_pre, w, bias, N, H, W, chunks, plane, act);
  return cudaGetLastError();
}

cudaError_t launch_conv_direct(cudaStream_t st, int ks, const act_t* in, long long in_plane, int cin_chunks, act_t* out,
                               act_t* out_pre, const act_t* res, long long out_plane, long long pre_plane,
                               long long res_plane, int cout_chunks, const act_t* w, const float* bias, int N, int H,
                               int W, int act) {
  const long long total = (long long)N * H * W;
  dim3 grid(blocks_for(total, 256), cout_chunks);
  if (ks == 3)
    k_conv_direct<3><<<grid, 256, 0, st>>>(in, in_plane, cin_chunks, out, out_pre, res, out_plane, pre_plane, res_plane,
                                           cout_chunks, w, bias, N, H, W, act);
  else
    k_conv_direct<4><<<grid, 256, 0, st>>>(in, in_plane, cin_chunks, out, out_pre, res, out_plane, pre_plane, res_plane,
                                           cout_chunks, w, bias, N, H, W, act);
  return cudaGetLastError();
}

template <int KS>
static cudaError_t launch_small_ks(cudaStream_t st, int cout, const act_t* in, long long in_plane, int cin_chunks,
                                   float* out, const float* w, const float* bias, int N, int H, int W, int act) {
  const long long total = (long long)N * H * W;
  const size_t smem = (size_t)KS * KS * cin_chunks * 8 * cout * sizeof(float);
  const unsigned nb = blocks_for(total, 128);
#define ARU_SMALL(C)                                                                                              \
  case C:                                                                                                         \
    cudaFuncSetAttribute(k_conv_small<KS, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k_conv_small<KS, C><<<nb, 128, smem, st>>>(in, in_plane, cin_chunks, out, w, bias, N, H, W, act);            \
    break;
  switch (cout) {
    ARU_SMALL(1) ARU_SMALL(2) ARU_SMALL(3) ARU_SMALL(4) ARU_SMALL(5) ARU_SMALL(6) ARU_SMALL(7) ARU_SMALL(8)
    default:
      return cudaErrorInvalidValue;
  }
#undef ARU_SMALL
  return cudaGetLastError();
}

cudaError_t launch_conv_small(cudaStream_t st, int ks, int cout, const act_t* in, long long in_plane, int cin_chunks,
                              float* out, const float* w, const float* bias, int N, int H, int W, int act) {
  return ks == 3 ? launch_small_ks<3>(st, cout, in, in_plane, cin_chunks, out, w, bias, N, H, W, act)
                 : launch_small_ks<4>(st, cout, in, in_plane, cin_chunks, out, w, bias, N, H, W, act);
}

cudaError_t launch_deconv_direct(cudaStream_t st, const act_t* in, long long in_plane, int cin_chunks, act_t* out,
                                 long long out_plane, int cout_chunks, const act_t* w, const float* bias, int N, int Hi,
                                 int Wi, int Ho, int Wo, int act) {
  const long long total = (long long)N * Ho * Wo;
  dim3 grid(blocks_for(total, 256), cout_chunks);
  k_deconv_direct<<<grid, 256, 0, st>>>(in, in_plane, cin_chunks, out, out_plane, cout_chunks, w, bias, N, Hi, Wi, Ho,
                                        Wo, act);
  return cudaGetLastError();
}

cudaError_t launch_pool(cudaStream_t st, bool is_max, const act_t* in, long long in_plane, act_t* out,
                        long long out_plane, int chunks, int N, int Hi, int Wi, int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo;
  dim3 grid(blocks_for(total, 256), chunks);
  if (is_max)
    k_maxpool<<<grid, 256, 0, st>>>(in, in_plane, out, out_plane, N, Hi, Wi, Ho, Wo);
  else
    k_avgpool<<<grid, 256, 0, st>>>(in, in_plane, out, out_plane, N, Hi, Wi, Ho, Wo);
  return cudaGetLastError();
}

cudaError_t launch_pool_f32(cudaStream_t st, bool is_max, const float* in, float* out, int N, int Hi, int Wi, int Ho,
                            int Wo) {
  const long long total = (long long)N * Ho * Wo;
  k_pool_f32<<<blocks_for(total, 256), 256, 0, st>>>(in, out, N, Hi, Wi, Ho, Wo, is_max ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_combine(cudaStream_t st, const CombineArgs& a) {
  const long long total = (long long)a.N * a.H * a.W;
  k_combine<<<blocks_for(total, 256), 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_upsum(cudaStream_t st, const act_t* in, long long in_plane, int in_chunks, int Hi, int Wi, act_t* out,
                         long long out_plane, int out_chunks, int out_c, int N, int H, int W, int up, int oy, int ox) {
  const long long total = (long long)N * H * W;
  k_upsum<<<blocks_for(total, 256), 256, 0, st>>>(in, in_plane, in_chunks, Hi, Wi, out, out_plane, out_chunks, out_c, N,
                                                  H, W, up, oy, ox);
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

cudaError_t launch_quantize(cudaStream_t st, const float* prob, uint8_t* u8, uint8_t* mask, long long npix, int C,
                            float thr) {
  k_quantize<<<blocks_for(npix, 256), 256, 0, st>>>(prob, u8, mask, npix, C, thr * 255.f);
  return cudaGetLastError();
}

}  // namespace aru
