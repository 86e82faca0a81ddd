// post.cu - the integer steps either side of the ARU-Net forward pass, on the device.
//
//   f1  load_and_scale_image's colour step (net_post_processing_helper.py:28-33):
//         cv2.cvtColor(image, COLOR_BGR2GRAY) / 255.0   ->  k_pages_to_input (uint8 gray or BGR in, float32 net input out)
//   f2  SeparatorNetPostProcessor.post_process (separator_net_post_processor.py:25-99):
//         apply_cc_analysis (region_net_post_processor_base.py:230-251): 8-connected components, keep area >= min_size
//           -> k_cc_init / k_cc_merge / k_cc_count / k_cc_filter_pack  (union-find on pixel indices, atomicMin hooks)
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

cudaError_t launch_pages_to_input(cudaStream_t st, const uint8_t* pages, int channels, long long npix, float* out,
                                  uint8_t* gray_out) {
  if (channels != 1 && channels != 3) return cudaErrorInvalidValue;
  const unsigned blocks = (unsigned)((npix + 255) / 256);
  if (channels == 1) k_pages_to_input<1><<<blocks, 256, 0, st>>>(pages, npix, out, gray_out);
  else k_pages_to_input<3><<<blocks, 256, 0, st>>>(pages, npix, out, gray_out);
  return cudaGetLastError();
}

// ---- f2a: connected-component size filter ---------------------------------------------------------------------------
// One warp = 32 consecutive pixels of one row (blockDim = (32, 8), grid = (ceil(W/32), ceil(H/8), N)); a label is a page-
// local pixel index, -1 = background.  k_cc_init points every foreground pixel at the first pixel of its run inside the
// 32-pixel word (so horizontal runs are pre-merged), k_cc_merge hooks runs across word boundaries and to the row above
// (only where the link is not implied by a neighbour's link), k_cc_count flattens and accumulates areas per root (one
// atomicAdd per in-word run), k_cc_filter_pack keeps area >= min_size and writes the result as bits.
__device__ __forceinline__ int run_start(unsigned bits, int lane) {
  const unsigned zeros_below = ~bits & ((1u << lane) - 1u);
  return zeros_below ? 32 - __clz(zeros_below) : 0;
}

__global__ void __launch_bounds__(256) k_cc_init(const uint8_t* __restrict__ mask, int H, int W, int* __restrict__ label,
                                                 int* __restrict__ area) {
  const int lane = threadIdx.x;
  const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + threadIdx.y;
  if (y >= H) return;
  const long long page = (long long)blockIdx.z * H * W;
  const bool in = x < W;
  const bool fg = in && mask[page + (long long)y * W + x] != 0;
  const unsigned bits = __ballot_sync(0xffffffffu, fg);
  if (!in) return;
  const int idx = y * W + x;
  label[page + idx] = fg ? idx - lane + run_start(bits, lane) : -1;
  area[page + idx] = 0;
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

__global__ void __launch_bounds__(256) k_cc_merge(int H, int W, int* __restrict__ label_all) {
  const int lane = threadIdx.x;
  const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + threadIdx.y;
  if (y >= H || x >= W) return;
  int* L = label_all + (long long)blockIdx.z * H * W;
  const int idx = y * W + x;
  if (L[idx] < 0) return;
  const bool wf = x > 0 && L[idx - 1] >= 0;
  if (lane == 0 && wf) uf_union(L, idx, idx - 1);
  if (y == 0) return;
  const int up = idx - W;
  const bool nf = L[up] >= 0;
  const bool nwf = x > 0 && L[up - 1] >= 0;
  if (nf) {
    if (!(wf && nwf)) uf_union(L, idx, up);   // otherwise this ~ W ~ NW ~ N through the neighbours' own links
  } else {
    if (nwf && !wf) uf_union(L, idx, up - 1);
    if (x + 1 < W && L[up + 1] >= 0) uf_union(L, idx, up + 1);
  }
}

__global__ void __launch_bounds__(256) k_cc_count(int H, int W, int* __restrict__ label_all, int* __restrict__ area_all) {
  const int lane = threadIdx.x;
  const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + threadIdx.y;
  if (y >= H) return;
  const long long page = (long long)blockIdx.z * H * W;
  int* L = label_all + page;
  const int idx = y * W + x;
  const bool fg = x < W && L[idx] >= 0;
  const unsigned bits = __ballot_sync(0xffffffffu, fg);
  if (!fg) return;
  const int root = uf_find(L, idx);
  L[idx] = root;
  if (run_start(bits, lane) == lane) {          // leader of an in-word run: all its pixels share the root
    const unsigned above = ~(bits >> lane);     // first zero above the leader ends the run
    const int len = above ? __ffs(above) - 1 : 32 - lane;
    atomicAdd(&area_all[page + root], len);
  }
}

__global__ void __launch_bounds__(256) k_cc_filter_pack(int H, int W, int Wd, const int* __restrict__ label_all,
                                                        const int* __restrict__ area_all, int min_size,
                                                        uint32_t* __restrict__ bits_out) {
  const int lane = threadIdx.x;
  const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + threadIdx.y;
  if (y >= H) return;
  const long long page = (long long)blockIdx.z * H * W;
  bool keep = false;
  if (x < W) {
    const int root = label_all[page + (long long)y * W + x];
    keep = root >= 0 && area_all[page + root] >= min_size;
  }
  const unsigned bits = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) bits_out[((long long)blockIdx.z * H + y) * Wd + blockIdx.x] = bits;
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

// bits -> uint8 {0,255}; one thread per 4 pixels of a row (one 32-bit store when the row pitch allows it)
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

// uint8 mask -> bits (used when the component filter is skipped in tests of the morphology alone)
__global__ void __launch_bounds__(256) k_bits_pack(const uint8_t* __restrict__ mask, int H, int W, int Wd,
                                                   uint32_t* __restrict__ bits_out) {
  const int lane = threadIdx.x;
  const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + threadIdx.y;
  if (y >= H) return;
  const bool fg = x < W && mask[((long long)blockIdx.z * H + y) * W + x] != 0;
  const unsigned bits = __ballot_sync(0xffffffffu, fg);
  if (lane == 0) bits_out[((long long)blockIdx.z * H + y) * Wd + blockIdx.x] = bits;
}

static inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

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

int separator_post_launches() { return 4 + 2 + 2 + 1 + 2 + 2; }

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
  const dim3 blk(32, 8), grd(Wd, (h + 7) / 8, n);
  k_cc_init<<<grd, blk, 0, st>>>(mask, h, w, label, area);
  k_cc_merge<<<grd, blk, 0, st>>>(h, w, label);
  k_cc_count<<<grd, blk, 0, st>>>(h, w, label, area);
  k_cc_filter_pack<<<grd, blk, 0, st>>>(h, w, Wd, label, area, min_size, b0);
  bits_open(st, b0, b3, b1, n, h, w, Wd, k_h1, true);    // b1 = horizontal opening
  bits_open(st, b0, b3, b2, n, h, w, Wd, k_v, false);    // b2 = vertical opening
  k_bits_andnot<<<nblk(words), 256, 0, st>>>(b1, b2, b0, (long long)words);
  bits_open(st, b0, b3, b1, n, h, w, Wd, k_h2, true);    // b1 = cleaned horizontal mask
  const long long rows = (long long)n * h, quads = rows * ((w + 3) / 4);
  k_bits_unpack<<<nblk(quads), 256, 0, st>>>(b1, out_h, rows, w, Wd);
  k_bits_unpack<<<nblk(quads), 256, 0, st>>>(b2, out_v, rows, w, Wd);
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
  const dim3 blk(32, 8), grd(Wd, (h + 7) / 8, n);
  k_bits_pack<<<grd, blk, 0, st>>>(mask, h, w, Wd, b0);
  const uint32_t* res = b0;
  if (kw > 1) { bits_open(st, b0, b3, b1, n, h, w, Wd, kw, true); res = b1; }
  else if (kh > 1) { bits_open(st, b0, b3, b1, n, h, w, Wd, kh, false); res = b1; }
  const long long rows = (long long)n * h;
  k_bits_unpack<<<nblk(rows * ((w + 3) / 4)), 256, 0, st>>>(res, out, rows, w, Wd);
  return cudaGetLastError();
}

}  // namespace aru
