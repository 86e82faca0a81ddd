// post.cu - the integer steps either side of the ARU-Net forward pass, on the device.
//
//   f1  load_and_scale_image's colour step (net_post_processing_helper.py:28-33):
//         cv2.cvtColor(image, COLOR_BGR2GRAY) / 255.0   ->  k_pages_to_input (uint8 gray or BGR in, float32 net input out)
//   f2  SeparatorNetPostProcessor.post_process (separator_net_post_processor.py:25-99):
//         apply_cc_analysis (region_net_post_processor_base.py:230-251): 8-connected components, keep area >= min_size
//           -> k_bits_pack, k_cc_init / k_cc_merge / k_cc_count / k_cc_filter  (union-find over in-word runs, atomicMin hooks)
//         cv2.morphologyEx(MORPH_OPEN, RECT (kw,1)) / (1,kh), cv2.subtract, MORPH_OPEN (kw2,1)
//           -> k_bits_h / k_bits_v / k_bits_andnot on a 1-bit-per-pixel image (32 pixels per word), k_bits_unpack
//
// Everything here is integer / bit work and bit-exact against OpenCV's rules (oracle/separator_post_oracle.py, pinned
// against the reference's own post_process by tests/golden/make_post_golden.py):
//   * erode / dilate with a (k x 1) or (1 x k) rectangle and the default anchor k/2 read src(x + j - k/2), j in [0,k);
//     taps outside the image are ignored; erode and dilate use the SAME offsets (an even k shifts the opening by +1);
//   * uint8 subtract saturates: on {0,255} images it is h AND NOT v.
// All kernels are HBM/L2-bound byte work: the masks are read once (1 B/px), the label / area arrays are int32 per pixel,
// and the morphology runs on 1/8 B per pixel.
#include "kernels.h"

namespace aru {

// ---- f1: uint8 page -> float32 net input ----------------------------------------------------------------------------
// float32(u8 / 255.0 in float64) == float32(u8) / 255.f for all 256 values (both correctly rounded; checked exhaustively
// in tests/test_post_oracle.py); BGR -> gray is OpenCV's 8-bit fixed point (B*3735 + G*19235 + R*9798 + 2^14) >> 15.
template <int CH>
__global__ void __launch_bounds__(256) k_pages_to_input(const uint8_t* __restrict__ pages, long long npix,
                                                        float* __restrict__ out, uint8_t* __restrict__ gray_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  unsigned g;
  if (CH == 1) {
    g = pages[i];
  } else {
    const unsigned b = pages[i * 3], gg = pages[i * 3 + 1], r = pages[i * 3 + 2];
    g = (b * 3735u + gg * 19235u + r * 9798u + (1u << 14)) >> 15;
  }
  out[i] = __fdiv_rn((float)g, 255.f);
  if (gray_out) gray_out[i] = (uint8_t)g;
}

// 4 pixels per thread: one 32-bit (gray) or three 32-bit (BGR) loads, one 128-bit store
template <int CH>
__global__ void __launch_bounds__(256) k_pages_to_input4(const uint32_t* __restrict__ pages, long long nquad,
                                                         float4* __restrict__ out, uint32_t* __restrict__ gray_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nquad) return;
  unsigned g[4];
  if (CH == 1) {
    const uint32_t v = pages[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] = (v >> (8 * j)) & 255u;
  } else {
    const uint32_t a = pages[i * 3], b = pages[i * 3 + 1], c = pages[i * 3 + 2];
    const unsigned by[12] = {a & 255u, (a >> 8) & 255u, (a >> 16) & 255u, a >> 24, b & 255u, (b >> 8) & 255u,
                             (b >> 16) & 255u, b >> 24, c & 255u, (c >> 8) & 255u, (c >> 16) & 255u, c >> 24};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      g[j] = (by[3 * j] * 3735u + by[3 * j + 1] * 19235u + by[3 * j + 2] * 9798u + (1u << 14)) >> 15;
  }
  out[i] = make_float4(__fdiv_rn((float)g[0], 255.f), __fdiv_rn((float)g[1], 255.f), __fdiv_rn((float)g[2], 255.f),
                       __fdiv_rn((float)g[3], 255.f));
  if (gray_out) gray_out[i] = g[0] | (g[1] << 8) | (g[2] << 16) | (g[3] << 24);
}

cudaError_t launch_pages_to_input(cudaStream_t st, const uint8_t* pages, int channels, long long npix, float* out,
                                  uint8_t* gray_out) {
  if (channels != 1 && channels != 3) return cudaErrorInvalidValue;
  long long done = 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(pages) | reinterpret_cast<uintptr_t>(gray_out)) & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (aligned && npix >= 4) {
    const long long nquad = npix / 4;
    const unsigned blocks = (unsigned)((nquad + 255) / 256);
    if (channels == 1)
      k_pages_to_input4<1><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(pages), nquad,
                                                   reinterpret_cast<float4*>(out), reinterpret_cast<uint32_t*>(gray_out));
    else
      k_pages_to_input4<3><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(pages), nquad,
                                                   reinterpret_cast<float4*>(out), reinterpret_cast<uint32_t*>(gray_out));
    done = nquad * 4;
  }
  if (done < npix) {   // tail (or unaligned buffers): one pixel per thread
    const long long rest = npix - done;
    const unsigned blocks = (unsigned)((rest + 255) / 256);
    uint8_t* go = gray_out ? gray_out + done : nullptr;
    if (channels == 1) k_pages_to_input<1><<<blocks, 256, 0, st>>>(pages + done, rest, out + done, go);
    else k_pages_to_input<3><<<blocks, 256, 0, st>>>(pages + done * 3, rest, out + done, go);
  }
  return cudaGetLastError();
}

// ---- f1b: cv2.resize(..., INTER_AREA) of uint8 pages (scale_image, net_post_processing_helper.py:14-25, sc < 1) -------
// OpenCV's area resampling, restated operation by operation (oracle/resize_oracle.py; bit-exact against cv2.resize):
//   general scale: per destination column / row a table of (source index, float weight) entries (host, double
//     arithmetic, see area_table() in engine.cu); a source row is first reduced horizontally, buf = sum_k S[sx_k]*alpha_k
//     (float, entries in order), then rows are combined, sum = beta_0*buf_0 (+)= beta_j*buf_j; multiplies and adds are
//     separate roundings (no FMA); the result is rounded half to even and saturated;
//   integer scale (both axes): integer box sum; (sum + 2) >> 2 for 2x2, else round(float(sum) * (1.f / area)).
// One thread per destination pixel; the footprint of a thread is ceil(scale)+1 source rows / columns.
struct AreaTabs {
  const int* x_start; const int* x_si; const float* x_alpha;   // entries of column dx: [x_start[dx], x_start[dx+1])
  const int* y_start; const int* y_si; const float* y_alpha;
  int fast;        // 1: integer scales ix, iy
  int ix, iy;
  float inv_area;  // 1.f / (ix * iy)
};

template <int CH>
__global__ void __launch_bounds__(256) k_resize_area(const uint8_t* __restrict__ src, int n, int sh, int sw,
                                                     uint8_t* __restrict__ dst, int dh, int dw, AreaTabs t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * dh * dw) return;
  const int dx = (int)(i % dw);
  const long long r = i / dw;
  const int dy = (int)(r % dh), pg = (int)(r / dh);
  const uint8_t* S = src + (long long)pg * sh * sw * CH;
  int res[CH];
  if (t.fast) {
    int sum[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) sum[c] = 0;
    for (int yy = 0; yy < t.iy; ++yy) {
      const uint8_t* row = S + ((long long)(dy * t.iy + yy) * sw + (long long)dx * t.ix) * CH;
      for (int xx = 0; xx < t.ix; ++xx)
#pragma unroll
        for (int c = 0; c < CH; ++c) sum[c] += row[xx * CH + c];
    }
#pragma unroll
    for (int c = 0; c < CH; ++c)
      res[c] = (t.ix == 2 && t.iy == 2) ? (sum[c] + 2) >> 2 : __float2int_rn(__fmul_rn((float)sum[c], t.inv_area));
  } else {
    float acc[CH];
    const int y0 = t.y_start[dy], y1 = t.y_start[dy + 1], x0 = t.x_start[dx], x1 = t.x_start[dx + 1];
    for (int j = y0; j < y1; ++j) {
      const uint8_t* row = S + (long long)t.y_si[j] * sw * CH;
      const float beta = t.y_alpha[j];
      float buf[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) buf[c] = 0.f;
      for (int k = x0; k < x1; ++k) {
        const uint8_t* px = row + (long long)t.x_si[k] * CH;
        const float alpha = t.x_alpha[k];
#pragma unroll
        for (int c = 0; c < CH; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)px[c], alpha));
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float v = __fmul_rn(beta, buf[c]);
        acc[c] = j == y0 ? v : __fadd_rn(acc[c], v);
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) res[c] = y1 > y0 ? __float2int_rn(acc[c]) : 0;
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) dst[i * CH + c] = (uint8_t)min(max(res[c], 0), 255);
}

cudaError_t launch_resize_area(cudaStream_t st, const uint8_t* src, int channels, int n, int sh, int sw, uint8_t* dst,
                               int dh, int dw, const int* x_start, const int* x_si, const float* x_alpha,
                               const int* y_start, const int* y_si, const float* y_alpha, int fast, int ix, int iy) {
  if (channels != 1 && channels != 3) return cudaErrorInvalidValue;
  AreaTabs t{x_start, x_si, x_alpha, y_start, y_si, y_alpha, fast, ix, iy, fast ? 1.f / (float)(ix * iy) : 0.f};
  const long long total = (long long)n * dh * dw;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (channels == 1) k_resize_area<1><<<blocks, 256, 0, st>>>(src, n, sh, sw, dst, dh, dw, t);
  else k_resize_area<3><<<blocks, 256, 0, st>>>(src, n, sh, sw, dst, dh, dw, t);
  return cudaGetLastError();
}

// cv2.resize(INTER_CUBIC) for enlarging (scale_image with sc > 1, net_post_processing_helper.py:21-23): OpenCV's 8-bit
// path in its scalar form - per destination column / row four source indices (clamped to the image) and four bicubic
// weights (A = -0.75, computed in float32 from fx = (dx + 0.5) / sc - 0.5) as 11-bit fixed point, horizontal pass first,
// result (sum + 2^21) >> 22 saturated.  OpenCV's SIMD builds evaluate the vertical pass in float and differ from this
// scalar form by at most one grey level on a few percent of the pixels (tests: |device - cv2| <= 1), which is the
// stated tolerance of this path; against the restatement (oracle/resize_oracle.py) it is bit-exact.
template <int CH>
__global__ void __launch_bounds__(256) k_resize_cubic(const uint8_t* __restrict__ src, int n, int sh, int sw,
                                                      uint8_t* __restrict__ dst, int dh, int dw,
                                                      const int* __restrict__ x_idx, const int* __restrict__ x_coef,
                                                      const int* __restrict__ y_idx, const int* __restrict__ y_coef) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * dh * dw) return;
  const int dx = (int)(i % dw);
  const long long r = i / dw;
  const int dy = (int)(r % dh), pg = (int)(r / dh);
  const uint8_t* S = src + (long long)pg * sh * sw * CH;
  long long acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = 0;
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const uint8_t* row = S + (long long)y_idx[dy * 4 + ky] * sw * CH;
    int hor[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) hor[c] = 0;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const uint8_t* px = row + (long long)x_idx[dx * 4 + kx] * CH;
      const int a = x_coef[dx * 4 + kx];
#pragma unroll
      for (int c = 0; c < CH; ++c) hor[c] += a * px[c];
    }
    const int b = y_coef[dy * 4 + ky];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] += (long long)b * hor[c];
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const long long v = (acc[c] + (1LL << 21)) >> 22;
    dst[i * CH + c] = (uint8_t)min(max(v, 0LL), 255LL);
  }
}

cudaError_t launch_resize_cubic(cudaStream_t st, const uint8_t* src, int channels, int n, int sh, int sw, uint8_t* dst,
                                int dh, int dw, const int* x_idx, const int* x_coef, const int* y_idx, const int* y_coef) {
  if (channels != 1 && channels != 3) return cudaErrorInvalidValue;
  const long long total = (long long)n * dh * dw;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (channels == 1) k_resize_cubic<1><<<blocks, 256, 0, st>>>(src, n, sh, sw, dst, dh, dw, x_idx, x_coef, y_idx, y_coef);
  else k_resize_cubic<3><<<blocks, 256, 0, st>>>(src, n, sh, sw, dst, dh, dw, x_idx, x_coef, y_idx, y_coef);
  return cudaGetLastError();
}

// ---- f2a: connected-component size filter ---------------------------------------------------------------------------
// Works on the 1-bit image (bit b of word i of a row = pixel 32 i + b, pad bits 0), one thread per 32-pixel word.  The
// union-find nodes are the in-word runs of ones: a label is the page-local pixel index of a run's first pixel, and the
// int32 label / area arrays are only ever touched there (a thresholded separator map is a few percent foreground, so the
// traffic is the 1/8 B/px bit image plus sparse sectors).  k_cc_init makes every run its own root, k_cc_merge hooks runs
// across word boundaries and to the touching runs of the row above, (only where the link is not implied by a neighbour's own link), k_cc_count flattens and accumulates

__device__ __forceinline__ int run_start(unsigned bits, int lane) {
  const unsigned zeros_below = ~bits & ((1u << lane) - 1u);
  return zeros_below ? 32 - __clz(zeros_below) : 0;
}

// uint8 mask (non-zero = foreground) -> bits.  A warp packs 128 pixels of one row: 4 per lane, nibbles merged by shuffles.
template <bool VEC4>
__global__ void __launch_bounds__(256) k_bits_pack(const uint8_t* __restrict__ mask, int H, int W, int Wd,
                                                   uint32_t* __restrict__ bits_out) {
  const int lane = threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (y >= H) return;
  const int x = blockIdx.x * 128 + lane * 4;
  const uint8_t* row = mask + ((long long)blockIdx.z * H + y) * W;
  unsigned nib = 0;
  if (VEC4) {   // W % 4 == 0: every row is 4-byte aligned and x + 3 < W whenever x < W
    if (x < W) {
      const uchar4 v = *reinterpret_cast<const uchar4*>(row + x);
      nib = (v.x ? 1u : 0u) | (v.y ? 2u : 0u) | (v.z ? 4u : 0u) | (v.w ? 8u : 0u);
    }
  } else {
    for (int j = 0; j < 4; ++j)
      if (x + j < W && row[x + j]) nib |= 1u << j;
  }
  unsigned v = nib << (4 * (lane & 7));
  v |= __shfl_xor_sync(0xffffffffu, v, 1);
  v |= __shfl_xor_sync(0xffffffffu, v, 2);
  v |= __shfl_xor_sync(0xffffffffu, v, 4);
  const int wi = blockIdx.x * 4 + (lane >> 3);
  if ((lane & 7) == 0 && wi < Wd) bits_out[((long long)blockIdx.z * H + y) * Wd + wi] = v;
}

// ---- one thread per 32-pixel word; labels exist only at the first pixel of every in-word run ----
struct WordAt {
  long long t;   // word index in the batch
  int n, y, i;   // page, row, word of the row
};
__device__ __forceinline__ bool word_at(int N, int H, int Wd, WordAt& a) {
  a.t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_page = (long long)H * Wd;
  if (a.t >= per_page * N) return false;
  a.n = (int)(a.t / per_page);
  const long long rem = a.t - a.n * per_page;
  a.y = (int)(rem / Wd);
  a.i = (int)(rem - (long long)a.y * Wd);
  return true;
}
// next run of ones in `rem` (a suffix of `bits`): [s, s + len); strips it from rem
__device__ __forceinline__ void next_run(unsigned bits, unsigned& rem, int& s, int& len) {
  s = __ffs(rem) - 1;
  const unsigned above = ~(bits >> s);
  len = above ? __ffs(above) - 1 : 32 - s;
  rem = (s + len >= 32) ? 0u : (rem & ~((1u << (s + len)) - 1u));
}

__global__ void __launch_bounds__(256) k_cc_init(const uint32_t* __restrict__ bits_in, int N, int H, int W, int Wd,
                                                 int* __restrict__ label, int* __restrict__ area) {
  WordAt a;
  if (!word_at(N, H, Wd, a)) return;
  const unsigned bits = bits_in[a.t];
  if (!bits) return;
  const long long page = (long long)a.n * H * W;
  const int base = a.y * W + a.i * 32;
  unsigned rem = bits;
  while (rem) {
    int s, len;
    next_run(bits, rem, s, len);
    label[page + base + s] = base + s;
    area[page + base + s] = 0;
  }
}

__device__ __forceinline__ int uf_find(volatile int* L, int a) {
  int p;
  while ((p = L[a]) != a) a = p;
  return a;
}

// find with full path compression: every node on the walked path is re-pointed at the root with atomicMin (a plain
// store could overwrite a concurrent hook of a node that is still a root; the minimum of two ancestors is an ancestor)
__device__ __forceinline__ int uf_find_compress(int* L, int a) {
  volatile int* V = L;
  int r = a, p;
  while ((p = V[r]) != r) r = p;
  while ((p = V[a]) > r) { atomicMin(&L[a], r); a = p; }
  return r;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  for (;;) {
    a = uf_find_compress(L, a);
    b = uf_find_compress(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&L[b], a);  // hook the larger root under the smaller one
    if (old == b) return;
    b = old;                              // somebody re-hooked b meanwhile: keep merging its new parent with a
  }
}

// Every in-word run is hooked to (a) the run ending at bit 31 of the word to its left when it starts at bit 0, and
// (b) every run fragment of the row above that touches it 8-connectedly: columns [s-1, s+len] of the row above, taken
// from a 34-bit view U of that row (bit 0 = last pixel of the word up-left, bits 1..32 = the word above, bit 33 = first
// pixel of the word up-right).
__global__ void __launch_bounds__(256) k_cc_merge(const uint32_t* __restrict__ bits_in, int N, int H, int W, int Wd,
                                                  int* __restrict__ label_all) {
  WordAt a;
  if (!word_at(N, H, Wd, a)) return;
  const unsigned cur = bits_in[a.t];
  if (!cur) return;
  int* L = label_all + (long long)a.n * H * W;
  const int base = a.y * W + a.i * 32;
  const unsigned left = a.i > 0 ? bits_in[a.t - 1] : 0u;
  unsigned up = 0, upl = 0, upr = 0;
  if (a.y > 0) {
    up = bits_in[a.t - Wd];
    if (a.i > 0) upl = bits_in[a.t - Wd - 1];
    if (a.i + 1 < Wd) upr = bits_in[a.t - Wd + 1];
  }
  const unsigned long long U = ((unsigned long long)up << 1) | (upl >> 31) | ((unsigned long long)(upr & 1u) << 33);
  const int ubase = base - W;   // index of bit 1 of U (pixel 32 i of the row above)
  unsigned rem = cur;
  while (rem) {
    int s, len;
    next_run(cur, rem, s, len);
    const int me = base + s;
    if (s == 0 && (left >> 31)) uf_union(L, me, base - 32 + run_start(left, 31));
    // bits [s, s+len+1] of U
    unsigned long long touch = (U >> s) & ((2ull << (len + 1)) - 1ull);
    while (touch) {
      const int b = __ffsll((long long)touch) - 1 + s;       // bit position in U of the first pixel of a fragment
      int other;
      if (b == 0) other = ubase - 32 + run_start(upl, 31);
      else if (b == 33) other = ubase + 32;
      else other = ubase + run_start(up, b - 1);
      uf_union(L, me, other);
      // strip this fragment (consecutive ones from b upwards, within U)
      const unsigned long long from = touch >> (b - s);
      const unsigned long long inv = ~from;
      const int flen = inv ? __ffsll((long long)inv) - 1 : 64;
      touch = (b - s + flen >= 64) ? 0ull : (touch & ~((1ull << (b - s + flen)) - 1ull));
    }
  }
}

__global__ void __launch_bounds__(256) k_cc_count(const uint32_t* __restrict__ bits_in, int N, int H, int W, int Wd,
                                                  int* __restrict__ label_all, int* __restrict__ area_all) {
  WordAt a;
  if (!word_at(N, H, Wd, a)) return;
  const unsigned bits = bits_in[a.t];
  if (!bits) return;
  const long long page = (long long)a.n * H * W;
  int* L = label_all + page;
  const int base = a.y * W + a.i * 32;
  unsigned rem = bits;
  while (rem) {
    int s, len;
    next_run(bits, rem, s, len);
    const int root = uf_find(L, base + s);
    L[base + s] = root;
    atomicAdd(&area_all[page + root], len);
  }
}

__global__ void __launch_bounds__(256) k_cc_filter(const uint32_t* __restrict__ bits_in, int N, int H, int W, int Wd,
                                                   const int* __restrict__ label_all, const int* __restrict__ area_all,
                                                   int min_size, uint32_t* __restrict__ bits_out) {
  WordAt a;
  if (!word_at(N, H, Wd, a)) return;
  const unsigned bits = bits_in[a.t];
  unsigned keep = 0;
  if (bits) {
    const long long page = (long long)a.n * H * W;
    const int base = a.y * W + a.i * 32;
    unsigned rem = bits;
    while (rem) {
      int s, len;
      next_run(bits, rem, s, len);
      if (area_all[page + label_all[page + base + s]] >= min_size)
        keep |= (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s;
    }
  }
  bits_out[a.t] = keep;
}

// ---- f2b: rectangular erode / dilate on the 1-bit image -------------------------------------------------------------
// Bit b of word i of a row = pixel x = 32 i + b; bits at x >= W are kept 0.  One thread per output word.
template <bool ERODE>
__global__ void __launch_bounds__(256) k_bits_h(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long rows,
                                                int W, int Wd, int k) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * Wd) return;
  const long long row = t / Wd;
  const int i = (int)(t - row * Wd);
  const uint32_t* r = in + row * Wd;
  const uint32_t tail = (W & 31) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;  // valid bits of the last word
  const uint32_t fill = ERODE ? 0xffffffffu : 0u;                         // ignored taps never decide
  auto fetch = [&](int wi) -> uint32_t {
    if (wi < 0 || wi >= Wd) return fill;
    const uint32_t v = r[wi];
    return (ERODE && wi == Wd - 1) ? (v | ~tail) : v;
  };
  const int a = k >> 1;
  uint32_t acc = fill;
  int q = -((a + 31) >> 5);             // word offset of the first tap: floor(-a / 32)
  int rbit = -a - q * 32;               // in [0, 31]
  uint32_t lo = fetch(i + q), hi = fetch(i + q + 1);
  for (int j = 0; j < k; ++j) {
    const uint32_t v = __funnelshift_r(lo, hi, rbit);   // pixels x + j - a for the 32 x of this word
    acc = ERODE ? (acc & v) : (acc | v);
    if (++rbit == 32) { rbit = 0; ++q; lo = hi; hi = fetch(i + q + 1); }
  }
  if (i == Wd - 1) acc &= tail;
  out[t] = acc;
}

template <bool ERODE>
__global__ void __launch_bounds__(256) k_bits_v(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int N, int H,
                                                int Wd, int k) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_page = (long long)H * Wd;
  if (t >= per_page * N) return;
  const long long n = t / per_page;
  const long long rem = t - n * per_page;
  const int y = (int)(rem / Wd), i = (int)(rem - (long long)y * Wd);
  const uint32_t* p = in + n * per_page + i;
  const int y0 = max(y - (k >> 1), 0), y1 = min(y - (k >> 1) + k - 1, H - 1);
  uint32_t acc = ERODE ? 0xffffffffu : 0u;
  for (int yy = y0; yy <= y1; ++yy) {
    const uint32_t v = p[(long long)yy * Wd];
    acc = ERODE ? (acc & v) : (acc | v);
  }
  out[t] = acc;   // pad bits: AND / OR of zeros stays zero (y is always inside its own window)
}

__global__ void __launch_bounds__(256) k_bits_andnot(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                     uint32_t* __restrict__ out, long long words) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < words) out[t] = a[t] & ~b[t];
}

// bits -> uint8 {0,255}; one thread per 16 pixels of a row (one 128-bit store) when W % 16 == 0
__global__ void __launch_bounds__(256) k_bits_unpack16(const uint32_t* __restrict__ bits, uint8_t* __restrict__ out,
                                                       long long rows, int W, int Wd) {
  const int per_row = W >> 4;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * per_row) return;
  const long long row = t / per_row;
  const int x = (int)(t - row * per_row) * 16;
  const uint32_t h = (bits[row * Wd + (x >> 5)] >> (x & 31)) & 0xffffu;
  uint4 v;
  uint32_t* o = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t nib = (h >> (4 * j)) & 15u;
    o[j] = ((nib & 1u) ? 0xffu : 0u) | ((nib & 2u) ? 0xff00u : 0u) | ((nib & 4u) ? 0xff0000u : 0u) |
           ((nib & 8u) ? 0xff000000u : 0u);
  }
  *reinterpret_cast<uint4*>(out + row * W + x) = v;
}

// general widths: one thread per 4 pixels of a row (one 32-bit store when the row pitch allows it)
__global__ void __launch_bounds__(256) k_bits_unpack(const uint32_t* __restrict__ bits, uint8_t* __restrict__ out,
                                                     long long rows, int W, int Wd) {
  const int quads = (W + 3) >> 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * quads) return;
  const long long row = t / quads;
  const int x = (int)(t - row * quads) * 4;
  const uint32_t nib = (bits[row * Wd + (x >> 5)] >> (x & 31)) & 15u;
  uint8_t* o = out + row * W + x;
  if ((W & 3) == 0) {
    const uint32_t v = ((nib & 1u) ? 0xffu : 0u) | ((nib & 2u) ? 0xff00u : 0u) | ((nib & 4u) ? 0xff0000u : 0u) |
                       ((nib & 8u) ? 0xff000000u : 0u);
    *reinterpret_cast<uint32_t*>(o) = v;
  } else {
    for (int j = 0; j < 4 && x + j < W; ++j) o[j] = ((nib >> j) & 1u) ? 255 : 0;
  }
}

// ---- f3: per-text-line sums of the heading map ----------------------------------------------------------------------
// HeadingNetPostProcessor.get_net_prob_for_text_line (heading_net_post_processor.py:247-270) sums net_output[ya:yb, xa:xb]
// (channel 0 of the uint8 map / 255) inside the bounding box of every text line.  One block per box; the sum is an
// exact integer (the caller divides by 255 and by the box area), so nothing but a few numbers per page leaves the GPU.
// boxes: [n_boxes][5] = page, y0, y1, x0, x1 (half open, already clipped to the page); boxes of pages outside
// [page0, page0 + n) are left to the micro-batch that holds their page.
__global__ void __launch_bounds__(256) k_box_sums(const uint8_t* __restrict__ u8, int n, int H, int W, int C, int page0,
                                                  const int* __restrict__ boxes, int n_boxes,
                                                  unsigned long long* __restrict__ sums) {
  const int b = blockIdx.x;
  if (b >= n_boxes) return;
  const int pg = boxes[b * 5] - page0;
  if (pg < 0 || pg >= n) return;
  const int y0 = boxes[b * 5 + 1], y1 = boxes[b * 5 + 2], x0 = boxes[b * 5 + 3], x1 = boxes[b * 5 + 4];
  unsigned long long acc = 0;
  const int bw = x1 - x0;
  if (bw > 0 && y1 > y0) {
    // threads cover the box row-major in chunks of 256 pixels: consecutive lanes read consecutive pixels of a row
    const long long total = (long long)bw * (y1 - y0);
    for (long long i = threadIdx.x; i < total; i += 256) {
      const int r = (int)(i / bw), c = (int)(i - (long long)r * bw);
      acc += u8[(((long long)pg * H + y0 + r) * W + x0 + c) * C];
    }
  }
  __shared__ unsigned long long part[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i];
    sums[b] = t;
  }
}

cudaError_t launch_box_sums(cudaStream_t st, const uint8_t* u8, int n, int h, int w, int c, int page0, const int* boxes,
                            int n_boxes, unsigned long long* sums) {
  if (n_boxes <= 0) return cudaSuccess;
  k_box_sums<<<n_boxes, 256, 0, st>>>(u8, n, h, w, c, page0, boxes, n_boxes, sums);
  return cudaGetLastError();
}

static inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

static void bits_pack(cudaStream_t st, const uint8_t* mask, int n, int h, int w, int Wd, uint32_t* bits) {
  const dim3 blk(32, 8), grd((w + 127) / 128, (h + 7) / 8, n);
  if ((w & 3) == 0 && (reinterpret_cast<uintptr_t>(mask) & 3) == 0) k_bits_pack<true><<<grd, blk, 0, st>>>(mask, h, w, Wd, bits);
  else k_bits_pack<false><<<grd, blk, 0, st>>>(mask, h, w, Wd, bits);
}

static void bits_unpack(cudaStream_t st, const uint32_t* bits, uint8_t* out, int n, int h, int w, int Wd) {
  const long long rows = (long long)n * h;
  if ((w & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    k_bits_unpack16<<<nblk(rows * (w >> 4)), 256, 0, st>>>(bits, out, rows, w, Wd);
  else
    k_bits_unpack<<<nblk(rows * ((w + 3) / 4)), 256, 0, st>>>(bits, out, rows, w, Wd);
}

static void bits_open(cudaStream_t st, const uint32_t* in, uint32_t* tmp, uint32_t* out, int N, int H, int W, int Wd,
                      int k, bool horizontal) {
  const long long rows = (long long)N * H, words = rows * Wd;
  if (horizontal) {
    k_bits_h<true><<<nblk(words), 256, 0, st>>>(in, tmp, rows, W, Wd, k);
    k_bits_h<false><<<nblk(words), 256, 0, st>>>(tmp, out, rows, W, Wd, k);
  } else {
    k_bits_v<true><<<nblk(words), 256, 0, st>>>(in, tmp, N, H, Wd, k);
    k_bits_v<false><<<nblk(words), 256, 0, st>>>(tmp, out, N, H, Wd, k);
  }
}

size_t separator_post_scratch_bytes(int n, int h, int w) {
  const size_t px = (size_t)n * h * w, words = (size_t)n * h * ((w + 31) / 32);
  return 2 * px * sizeof(int) + 4 * words * sizeof(uint32_t) + 1024;
}

int separator_post_launches() { return 5 + 2 + 2 + 1 + 2 + 2; }

cudaError_t launch_separator_post(cudaStream_t st, const uint8_t* mask, int n, int h, int w, int min_size, int k_h1,
                                  int k_v, int k_h2, void* scratch, uint8_t* out_h, uint8_t* out_v) {
  if (n <= 0 || h <= 0 || w <= 0 || k_h1 < 1 || k_v < 1 || k_h2 < 1 || (long long)h * w >= (1LL << 31) || n > 65535)
    return cudaErrorInvalidValue;
  const int Wd = (w + 31) / 32;
  const size_t px = (size_t)n * h * w, words = (size_t)n * h * Wd;
  int* label = reinterpret_cast<int*>(scratch);
  int* area = label + px;
  uint32_t* b0 = reinterpret_cast<uint32_t*>(area + px);
  uint32_t *b1 = b0 + words, *b2 = b1 + words, *b3 = b2 + words;
  bits_pack(st, mask, n, h, w, Wd, b3);
  k_cc_init<<<nblk(words), 256, 0, st>>>(b3, n, h, w, Wd, label, area);
  k_cc_merge<<<nblk(words), 256, 0, st>>>(b3, n, h, w, Wd, label);
  k_cc_count<<<nblk(words), 256, 0, st>>>(b3, n, h, w, Wd, label, area);
  k_cc_filter<<<nblk(words), 256, 0, st>>>(b3, n, h, w, Wd, label, area, min_size, b0);
  bits_open(st, b0, b3, b1, n, h, w, Wd, k_h1, true);    // b1 = horizontal opening
  bits_open(st, b0, b3, b2, n, h, w, Wd, k_v, false);    // b2 = vertical opening
  k_bits_andnot<<<nblk(words), 256, 0, st>>>(b1, b2, b0, (long long)words);
  bits_open(st, b0, b3, b1, n, h, w, Wd, k_h2, true);    // b1 = cleaned horizontal mask
  bits_unpack(st, b1, out_h, n, h, w, Wd);
  bits_unpack(st, b2, out_v, n, h, w, Wd);
  return cudaGetLastError();
}

// RegionNetPostProcessor.apply_cc_analysis alone (region_net_post_processor_base.py:230-251; also what
// TextBlockNetPostProcessor.post_process consists of, text_block_net_post_processor.py:12-24)
cudaError_t launch_cc_filter(cudaStream_t st, const uint8_t* mask, int n, int h, int w, int min_size, void* scratch,
                             uint8_t* out) {
  if (n <= 0 || h <= 0 || w <= 0 || (long long)h * w >= (1LL << 31) || n > 65535) return cudaErrorInvalidValue;
  const int Wd = (w + 31) / 32;
  const size_t px = (size_t)n * h * w, words = (size_t)n * h * Wd;
  int* label = reinterpret_cast<int*>(scratch);
  int* area = label + px;
  uint32_t* b0 = reinterpret_cast<uint32_t*>(area + px);
  uint32_t* b1 = b0 + words;
  bits_pack(st, mask, n, h, w, Wd, b0);
  k_cc_init<<<nblk(words), 256, 0, st>>>(b0, n, h, w, Wd, label, area);
  k_cc_merge<<<nblk(words), 256, 0, st>>>(b0, n, h, w, Wd, label);
  k_cc_count<<<nblk(words), 256, 0, st>>>(b0, n, h, w, Wd, label, area);
  k_cc_filter<<<nblk(words), 256, 0, st>>>(b0, n, h, w, Wd, label, area, min_size, b1);
  bits_unpack(st, b1, out, n, h, w, Wd);
  return cudaGetLastError();
}

// morphology alone (tests): open_rect(mask, kw, kh) with kw == 1 or kh == 1
cudaError_t launch_open_rect(cudaStream_t st, const uint8_t* mask, int n, int h, int w, int kw, int kh, void* scratch,
                             uint8_t* out) {
  if (n <= 0 || h <= 0 || w <= 0 || kw < 1 || kh < 1 || (kw > 1 && kh > 1) || n > 65535) return cudaErrorInvalidValue;
  const int Wd = (w + 31) / 32;
  const size_t px = (size_t)n * h * w, words = (size_t)n * h * Wd;
  uint32_t* b0 = reinterpret_cast<uint32_t*>(reinterpret_cast<int*>(scratch) + 2 * px);
  uint32_t *b1 = b0 + words, *b3 = b1 + 2 * words;
  bits_pack(st, mask, n, h, w, Wd, b0);
  const uint32_t* res = b0;
  if (kw > 1) { bits_open(st, b0, b3, b1, n, h, w, Wd, kw, true); res = b1; }
  else if (kh > 1) { bits_open(st, b0, b3, b1, n, h, w, Wd, kh, false); res = b1; }
  bits_unpack(st, res, out, n, h, w, Wd);
  return cudaGetLastError();
}

}  // namespace aru
