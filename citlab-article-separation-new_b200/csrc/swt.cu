// swt.cu - the SWT feature image of the heading post-processor on the device:
//   StrokeWidthDistanceTransform.distance_transform (python_util/image_processing/swt_dist_trafo.py:18-29), called twice per
//   page at full image resolution by HeadingNetPostProcessor (heading_net_post_processor.py:86,297):
//     image = 255 - gray                                  (uint8 "-image + 255", dark_on_bright)
//     blur  = cv2.GaussianBlur(image, (5, 5), 0)          fixed kernel [1 4 6 4 1]/16, BORDER_REFLECT_101, one rounding
//     t, b  = cv2.threshold(blur, 0, 255, BINARY + OTSU)  256-bin histogram, OpenCV's getThreshVal_Otsu_8u on the host
//     d     = cv2.distanceTransform(b, DIST_L2, DIST_MASK_PRECISE)   exact Euclidean distance to the nearest zero pixel
//     return d.astype(np.uint8)                            truncation, distances >= 256 wrap
// Bit-exact against the reference's own function (oracle/swt_oracle.py pinned by tests/golden/make_post_golden.py).
// All byte / integer work, HBM / L2 bound: blur reads 1 B/px (25 taps through L1) and writes 1 B/px; the column pass
// reads 1 B/px and writes 2 B/px; the row pass reads 2 B/px (neighbours through L1) and writes 1 B/px.
#include <algorithm>
#include <vector>

#include "kernels.h"

namespace aru {

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// invert + 5x5 Gaussian + per-page histogram of the result (shared-memory bins, one global atomic per bin and block)
__global__ void __launch_bounds__(256) k_swt_blur(const uint8_t* __restrict__ gray, int n, int H, int W, int invert,
                                                  uint8_t* __restrict__ blur, unsigned* __restrict__ hist) {
  __shared__ unsigned bins[256];
  bins[threadIdx.x] = 0;
  __syncthreads();
  const int page = blockIdx.y;
  const long long px = (long long)H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < px) {
    const int y = (int)(i / W), x = (int)(i - (long long)y * W);
    const uint8_t* g = gray + (long long)page * px;
    const int k[5] = {1, 4, 6, 4, 1};
    int acc = 0;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
      const uint8_t* row = g + (long long)reflect101(y + dy - 2, H) * W;
      int r = 0;
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
        const int v = row[reflect101(x + dx - 2, W)];
        r += k[dx] * (invert ? 255 - v : v);
      }
      acc += k[dy] * r;
    }
    const int b = (acc + 128) >> 8;
    blur[(long long)page * px + i] = (uint8_t)b;
    atomicAdd(&bins[b], 1u);
  }
  __syncthreads();
  if (bins[threadIdx.x]) atomicAdd(&hist[page * 256 + threadIdx.x], bins[threadIdx.x]);
}

// column pass: g[y][x] = distance along the column to the nearest pixel with blur <= thr (0 for such a pixel), capped
__global__ void __launch_bounds__(128) k_swt_cols(const uint8_t* __restrict__ blur, int H, int W, const int* __restrict__ thr,
                                                  unsigned short* __restrict__ g) {
  const int page = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const long long px = (long long)H * W;
  const uint8_t* b = blur + (long long)page * px;
  unsigned short* gp = g + (long long)page * px;
  const int t = thr[page];
  int d = 60000;                                  // "no zero pixel above": larger than any page dimension
  for (int y = 0; y < H; ++y) {
    d = b[(long long)y * W + x] > t ? min(d + 1, 60000) : 0;
    gp[(long long)y * W + x] = (unsigned short)d;
  }
  d = 60000;
  for (int y = H - 1; y >= 0; --y) {
    const int down = gp[(long long)y * W + x];
    d = down == 0 ? 0 : min(d + 1, 60000);
    if (d < down) gp[(long long)y * W + x] = (unsigned short)d;
  }
}

// row pass: exact squared distance = min over x' of (x - x')^2 + g[x']^2; the search stops once (x - x')^2 reaches the
// best value so far (it starts at g[x]^2, so text strokes take a handful of steps)
__global__ void __launch_bounds__(256) k_swt_rows(const unsigned short* __restrict__ g, int H, int W,
                                                  uint8_t* __restrict__ out) {
  const int page = blockIdx.y;
  const long long px = (long long)H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= px) return;
  const int y = (int)(i / W), x = (int)(i - (long long)y * W);
  const unsigned short* row = g + (long long)page * px + (long long)y * W;
  long long best = (long long)row[x] * row[x];
  if (best != 0) {
    for (int k = 1; (long long)k * k < best; ++k) {
      const long long k2 = (long long)k * k;
      if (x - k >= 0) { const long long v = row[x - k]; best = min(best, k2 + v * v); }
      if (x + k < W) { const long long v = row[x + k]; best = min(best, k2 + v * v); }
      if (x - k < 0 && x + k >= W) break;
    }
  }
  // float32 sqrt of the exact integer (cv2 computes sqrt(float(dist2))), truncation, modulo 256 (numpy's uint8 cast)
  const float d = __fsqrt_rn((float)best);
  out[(long long)page * px + i] = (uint8_t)((long long)d & 0xFF);
}

}  // namespace

// scratch layout: histograms [n][256] u32 | thresholds [n] i32 | (256 B aligned) column distances u16 | blurred pages u8
static size_t swt_head_bytes(int n) { return (((size_t)n * 257 * 4) + 255) / 256 * 256; }
size_t swt_scratch_bytes(int n, int h, int w) {
  const size_t px = (size_t)n * h * w;
  return swt_head_bytes(n) + 2 * px + px + 256;
}

// OpenCV's getThreshVal_Otsu_8u on one 256-bin histogram (double arithmetic, first maximum of the between-class variance)
int swt_otsu_from_hist(const unsigned* hist, long long size) {
  const double scale = 1.0 / (double)size;
  double mu = 0;
  for (int i = 0; i < 256; ++i) mu += (double)i * (double)hist[i];
  mu *= scale;
  double mu1 = 0, q1 = 0, max_sigma = 0;
  int max_val = 0;
  const double eps = 1.1920928955078125e-07;   // FLT_EPSILON
  for (int i = 0; i < 256; ++i) {
    const double p_i = hist[i] * scale;
    mu1 *= q1;
    q1 += p_i;
    const double q2 = 1.0 - q1;
    if (std::min(q1, q2) < eps || std::max(q1, q2) > 1.0 - eps) continue;
    mu1 = (mu1 + i * p_i) / q1;
    const double mu2 = (mu - q1 * mu1) / q2;
    const double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
    if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
  }
  return max_val;
}

// gray: device uint8 [n][h][w]; out: device uint8 [n][h][w]; thr_host (nullable): the Otsu thresholds.  Synchronises the
// stream once (the thresholds are computed on the host from the device histograms).
cudaError_t launch_swt_distance(cudaStream_t st, const uint8_t* gray, int n, int h, int w, int dark_on_bright, void* scratch,
                                uint8_t* out, int* thr_host) {
  const size_t px = (size_t)h * w;
  unsigned* hist = static_cast<unsigned*>(scratch);
  int* thr = reinterpret_cast<int*>(hist + (size_t)n * 256);
  unsigned short* g = reinterpret_cast<unsigned short*>(static_cast<uint8_t*>(scratch) + swt_head_bytes(n));
  uint8_t* blur = reinterpret_cast<uint8_t*>(g + (size_t)n * px);
  cudaError_t err = cudaMemsetAsync(hist, 0, (size_t)n * 256 * sizeof(unsigned), st);
  if (err != cudaSuccess) return err;
  const dim3 grid_px((unsigned)((px + 255) / 256), (unsigned)n);
  k_swt_blur<<<grid_px, 256, 0, st>>>(gray, n, h, w, dark_on_bright ? 1 : 0, blur, hist);
  std::vector<unsigned> hh((size_t)n * 256);
  err = cudaMemcpyAsync(hh.data(), hist, hh.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) return err;
  std::vector<int> th(n);
  bool any_empty = false;
  std::vector<char> no_zero(n, 0);
  for (int p = 0; p < n; ++p) {
    th[p] = swt_otsu_from_hist(hh.data() + (size_t)p * 256, (long long)px);
    unsigned long long zeros = 0;
    for (int i = 0; i <= th[p]; ++i) zeros += hh[(size_t)p * 256 + i];
    no_zero[p] = zeros == 0;      // no pixel at or below the threshold: OpenCV's distances are infinite, the cast gives 0
    any_empty |= no_zero[p] != 0;
    if (thr_host) thr_host[p] = th[p];
  }
  err = cudaMemcpyAsync(thr, th.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st);
  if (err != cudaSuccess) return err;
  k_swt_cols<<<dim3((unsigned)((w + 127) / 128), (unsigned)n), 128, 0, st>>>(blur, h, w, thr, g);
  k_swt_rows<<<grid_px, 256, 0, st>>>(g, h, w, out);
  if (any_empty)
    for (int p = 0; p < n && err == cudaSuccess; ++p)
      if (no_zero[p]) err = cudaMemsetAsync(out + (size_t)p * px, 0, px, st);
  if (err != cudaSuccess) return err;
  // th is read by the asynchronous copy above: keep it alive until the stream has consumed it
  err = cudaStreamSynchronize(st);
  return err != cudaSuccess ? err : cudaGetLastError();
}

}  // namespace aru
