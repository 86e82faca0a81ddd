// kernels.h - host-side launchers of the engine's CUDA kernels (kernels_simple.cu, conv_tc.cu).
#pragma once
#include "aru_common.cuh"

namespace aru {

// A channel-slice view of a chunk-planar tensor: `p` points at position 0 of the first plane.
struct PV {
  act_t* p = nullptr;
  long long plane = 0;  // positions between consecutive planes
  int chunks = 0;       // planes in the view
  int C = 0;            // logical channels
};

#define ARU_COMBINE_MAX 8
struct CombineArgs {
  const float* att[8];
  const act_t* det[8];
  long long det_plane[8];
  Geo det_geo[8];
  int att_h[8], att_w[8], att_up[8], att_oy[8], att_ox[8];
  int det_up[8], det_oy[8], det_ox[8], det_chunks[8];
  int att_sh[8], det_sh[8];  // log2 of the upsample factors (-1: not a power of two), filled by launch_combine
  int A;
  act_t* out;
  long long out_plane;
  int out_chunks;
  Geo geo;  // output geometry
};

// ---- CUDA-core kernels (kernels_simple.cu) -------------------------------------------------------
// w_host / bias_host: HOST pointers to the TF filter [ks][ks][1][C_out] and bias [C_out] (passed as kernel parameters)
// pool.p != null: the launch also performs the 2x2 stride-2 SAME max-pool of its ReLU'd output (pool_geo = the pooled
// geometry); full == false then skips the store of the full-resolution tensor (nothing but the pool reads it)
cudaError_t launch_conv_stem(cudaStream_t st, int ks, const float* in, PV out, PV out_pre, const float* w_host,
                             const float* bias_host, const Geo& g, int act, PV pool = PV(), const Geo* pool_geo = nullptr,
                             bool full = true, const float* w_dev16 = nullptr, const float* bias_dev16 = nullptr);
// w_dev16 / bias_dev16: device copies of the filter as [tap][16] float32 and of the bias padded to 16 (the tensor-core
// variant of the pooled 4x4 stem gathers its per-lane fragments from them)
cudaError_t launch_conv_direct(cudaStream_t st, int ks, PV in, PV out, PV out_pre, PV res, const act_t* w,
                               const float* bias, const Geo& g, int act);
cudaError_t launch_conv_small(cudaStream_t st, int ks, int cout, PV in, float* out, const float* w, const float* bias,
                              const Geo& g, int act);
cudaError_t launch_deconv_direct(cudaStream_t st, PV in, const Geo& gi, PV out, const Geo& go, const act_t* w,
                                 const float* bias, int act);
cudaError_t launch_pool(cudaStream_t st, bool is_max, PV in, const Geo& gi, PV out, const Geo& go);
cudaError_t launch_pool_f32(cudaStream_t st, bool is_max, const float* in, float* out, int N, int Hi, int Wi, int Ho,
                            int Wo);
cudaError_t launch_combine(cudaStream_t st, const CombineArgs& a);
cudaError_t launch_upsum(cudaStream_t st, PV in, const Geo& gi, PV out, const Geo& go, int up, int oy, int ox);
cudaError_t launch_upsum_f32(cudaStream_t st, const float* in, int Hi, int Wi, float* out, int N, int H, int W, int up,
                             int oy, int ox);
cudaError_t launch_copy(cudaStream_t st, const void* in, void* out, long long bytes);
// mask = 255 where u8[...,0] > thr*255 (float compare), or where u8[...,0] >= cut when cut >= 0
// u8_channels > 0: only that many leading channels are written, u8 is [npix][u8_channels] (0: all C)
cudaError_t launch_quantize(cudaStream_t st, const float* prob, uint8_t* u8, uint8_t* mask, long long npix, int C,
                            float thr, int cut = -1, int u8_channels = 0);

// ---- integer pre / post-processing around the net (post.cu) --------------------------------------
// uint8 gray (channels 1) or BGR (channels 3) pages -> float32 net input gray/255 (+ optional uint8 gray copy)
cudaError_t launch_pages_to_input(cudaStream_t st, const uint8_t* pages, int channels, long long npix, float* out,
                                  uint8_t* gray_out);
// cv2.resize(INTER_AREA) of n uint8 pages sh x sw -> dh x dw; the (source index, weight) tables are device arrays built
// by the engine (general scale) or unused (fast = 1: integer scales ix, iy)
cudaError_t launch_resize_area(cudaStream_t st, const uint8_t* src, int channels, int n, int sh, int sw, uint8_t* dst,
                               int dh, int dw, const int* x_start, const int* x_si, const float* x_alpha,
                               const int* y_start, const int* y_si, const float* y_alpha, int fast, int ix, int iy);
// cv2.resize(INTER_CUBIC) of n uint8 pages (enlarging): per destination index 4 clamped source indices and 4 fixed-point
// (11-bit) weights, device arrays [d][4] built by the engine
cudaError_t launch_resize_cubic(cudaStream_t st, const uint8_t* src, int channels, int n, int sh, int sw, uint8_t* dst,
                                int dh, int dw, const int* x_idx, const int* x_coef, const int* y_idx, const int* y_coef);
// SeparatorNetPostProcessor.post_process on n thresholded masks of h x w: component size filter, three rectangular
// openings and the saturating subtract; scratch >= separator_post_scratch_bytes(n, h, w)
size_t separator_post_scratch_bytes(int n, int h, int w);
int separator_post_launches();
cudaError_t launch_separator_post(cudaStream_t st, const uint8_t* mask, int n, int h, int w, int min_size, int k_h1,
                                  int k_v, int k_h2, void* scratch, uint8_t* out_h, uint8_t* out_v);
// sums[b] = sum of channel 0 of the uint8 map over box b = (page, y0, y1, x0, x1), for the boxes whose page lies in
// [page0, page0 + n); u8 holds those n pages as [n][h][w][c]
cudaError_t launch_box_sums(cudaStream_t st, const uint8_t* u8, int n, int h, int w, int c, int page0, const int* boxes,
                            int n_boxes, unsigned long long* sums);
// apply_cc_analysis: keep the 8-connected components of the non-zero pixels with area >= min_size (output {0,255})
cudaError_t launch_cc_filter(cudaStream_t st, const uint8_t* mask, int n, int h, int w, int min_size, void* scratch,
                             uint8_t* out);
cudaError_t launch_open_rect(cudaStream_t st, const uint8_t* mask, int n, int h, int w, int kw, int kh, void* scratch,
                             uint8_t* out);
// StrokeWidthDistanceTransform.distance_transform (swt_dist_trafo.py:18-29) of n gray pages (swt.cu): invert, 5x5 Gaussian,
// Otsu, exact Euclidean distance transform, uint8 truncation.  scratch >= swt_scratch_bytes(n, h, w); synchronises `st`.
size_t swt_scratch_bytes(int n, int h, int w);
int swt_otsu_from_hist(const unsigned* hist, long long size);
cudaError_t launch_swt_distance(cudaStream_t st, const uint8_t* gray, int n, int h, int w, int dark_on_bright, void* scratch,
                                uint8_t* out, int* thr_host);
// *count += number of stored 16-bit values at the storage limit (fp16: |v| = 65504) or not finite
cudaError_t launch_range_scan(cudaStream_t st, const act_t* p, long long elements, unsigned long long* count);
// debug read-back: chunk-planar view -> dense float32 NHWC [N][H][W][C]
cudaError_t launch_unpack_nhwc(cudaStream_t st, PV in, const Geo& g, float* out);

// ---- tcgen05 implicit-GEMM convolution (conv_tc.cu) ---------------------------------------------
struct ConvTcPlan {
  bool deconv = false;   // 3x3 stride-2 transposed convolution (2x2 look-back taps x 4 output-parity classes)
  bool ok = false;       // false: shape not covered by the tensor-core kernel (caller must use another kernel)
  int ks = 0;
  int cin_chunks = 0;
  int cout_chunks = 0;   // real output planes
  int npad = 0;          // UMMA N (multiple of 16)
  int n_steps = 0;       // K=16 MMA steps per tile
  int T = 1;             // 128-position tiles per pass
  int n_slots = 0;       // ring units (128*T positions each)
  int n_stages = 0;      // TMEM accumulator stages
  int w_stream = 0;      // 1: filters do not fit next to the operand ring and are streamed per pass
  int w_group = 0;       // K-steps per streamed weight group
  int n_wst = 0;         // weight ring stages
  int grid = 0;
  size_t smem_bytes = 0;
  size_t wpack_bytes = 0;  // packed B image (16-bit elements)
  long long tile_begin = 0, tile_end = 0;  // pass range [begin, end) in units of 128*T positions
  const char* why = "";
};

// Decide whether / how the tcgen05 kernel runs conv (ks x ks, cin -> cout) on geometry g.
ConvTcPlan conv_tc_plan(int ks, int cin, int cout, const Geo& g, int num_sms, size_t max_smem, bool deconv = false);
// Pack input channels [ci_begin, ci_begin + 8*plan.cin_chunks) of a TF-layout filter [ks][ks][cin][cout]
// (float32) into the kernel's shared-memory B image.
void conv_tc_pack_weights(const ConvTcPlan& plan, const float* w_tf, int cin, int cout, int ci_begin, uint16_t* dst);
// Transposed conv: filter [3][3][cout][cin] -> B image; launch over the *input* geometry gi, stores to go.
void conv_tc_pack_deconv(const ConvTcPlan& plan, const float* w_tf, int cin, int cout, uint16_t* dst);
cudaError_t launch_deconv_tc(cudaStream_t st, const ConvTcPlan& plan, PV in, PV out, const act_t* wpack,
                             const float* bias_pad, const Geo& gi, const Geo& go, int act, int* err_flag);
cudaError_t launch_conv_tc(cudaStream_t st, const ConvTcPlan& plan, PV in, PV out, PV out_pre, PV res,
                           const act_t* wpack, const float* bias_pad, const Geo& g, int act, int* err_flag);

// Small-C_out head (C_out <= 8) writing dense float32 NHWC with ReLU / softmax / sigmoid (enum aru_act) in the epilogue.
cudaError_t launch_conv_tc_f32(cudaStream_t st, const ConvTcPlan& plan, PV in, float* out_f32, int cout,
                               const act_t* wpack, const float* bias_pad, const Geo& g, int act, int* err_flag);

// ---- tcgen05 row-banded convolution for C_out <= 64 (conv_band.cu) ------------------------------
// Output rows x channels on the MMA M axis, N = a strip of up to 256 output columns, banded weight masters as A.
struct ConvBandPlan {
  bool ok = false;
  int ks = 0;
  int cin_chunks = 0;
  int cop = 0;           // C_out padded to a power of two (8..64); planes written = cop / 8
  int R = 0;             // output rows per tile = 128 / cop
  int J = 0;             // row groups of a banded master = 2R + ks - 2
  int N = 0;             // output columns per tile (UMMA N, multiple of 16)
  int n_strips = 0, n_ty = 0;
  int S = 0;             // row slots of the input ring (even)
  int n_masters = 0, n_steps = 0;
  int grid = 0;
  long long tiles = 0;
  size_t smem_bytes = 0;
  size_t wpack_bytes = 0;
  const char* why = "";
};
ConvBandPlan conv_band_plan(int ks, int cin, int cout, const Geo& g, int num_sms, size_t max_smem);
// TF filter [ks][ks][cin][cout] (float32) -> banded masters (16-bit), the kernel's resident shared-memory image
void conv_band_pack(const ConvBandPlan& plan, const float* w_tf, int cin, int cout, uint16_t* dst);
// out_f32 != null: small-C_out head (C_out <= 4) writing dense float32 NHWC with ReLU / softmax / sigmoid instead of planes
cudaError_t launch_conv_band(cudaStream_t st, const ConvBandPlan& plan, PV in, PV out, PV out_pre, PV res,
                             const act_t* wpack, const float* bias_pad, const Geo& g, int act, int* err_flag,
                             float* out_f32 = nullptr, int f32_c = 0, PV pool = PV(), const Geo* pool_geo = nullptr);
// conv_band_can_pool: the launch can also write the 2x2 stride-2 SAME max-pool of its ReLU'd output (pool = its view)
inline bool conv_band_can_pool(const ConvBandPlan& plan, int act, int cin, int ks) {
  return plan.ok && act == 1 && plan.R >= 2 && cin <= 32 && (ks == 3 || ks == 4);
}

// ---- two chained 3x3 convolutions in one launch (conv_band2.cu) ---------------------------------
// conv (cin -> c) [+ pre-activation export] -> act -> conv (c -> c) [+ residual] -> act [-> 2x2 max-pool], c = 8 / 16;
// the intermediate lives in a shared-memory row FIFO.  st0 / st1 describe the banded masters of the two stages
// (conv_band_pack's layout).
struct ConvBand2Plan {
  bool ok = false;
  int cin_chunks = 0, cop = 0, R = 0, J = 0;
  int N = 0;             // UMMA N of both stages (multiple of 16, <= 128)
  int n_out = 0;         // valid output columns per strip = N - 2
  int n_strips = 0, n_ty = 0;
  int S = 0;             // row slots of the input ring (even)
  int RS = 0;            // row slots of the intermediate FIFO = 2R + 2
  int nbuf0 = 2, nbuf1 = 2, acc_stride = 128;   // TMEM accumulator buffers per stage and their column stride
  int grid = 0;
  long long tiles = 0;
  size_t smem_bytes = 0;
  ConvBandPlan st0, st1;
  const char* why = "";
};
ConvBand2Plan conv_band2_plan(int cin, int cmid, int cout, const Geo& g, int num_sms, size_t max_smem);
// out0 / pre0: optional global copies of the stage-0 result (activated / pre-activation); res: stage-1 residual operand
cudaError_t launch_conv_band2(cudaStream_t st, const ConvBand2Plan& plan, PV in, PV out0, PV pre0, PV out, PV res,
                              const act_t* wpack0, const float* bias0, int act0, const act_t* wpack1, const float* bias1,
                              int act1, const Geo& g, int* err_flag, PV pool = PV(), const Geo* pool_geo = nullptr);

// ---- a whole residual block per launch on the warp-level tensor path (block_mma.cu) --------------
// [conv1 (cin -> c) ->] x0 -> ReLU -> convR_0 -> ReLU -> convR_1 -> ReLU -> convR_2 + x0 -> ReLU [-> 2x2 max-pool], c = 8;
// cin == 0: the launch starts from x0 (conv1's pre-activation, written by another kernel).
struct BlockMmaPlan {
  bool ok = false;
  int c = 0, nt = 0, cin_ch = 0;
  int WC = 0, WS = 0, n_strips = 0;          // compute columns per CTA, output columns per strip
  int rows_per_unit = 0, units_per_page = 0, n_units = 0;
  int grid = 0, threads = 0;
  int wbase[4] = {0, 0, 0, 0};               // first B-fragment register of every stage
  size_t smem_bytes = 0;
  size_t wfrag_bytes = 0;
  const char* why = "";
};
BlockMmaPlan block_mma_plan(int cin, int c, const Geo& g, int num_sms, size_t max_smem);
// w_tf[s]: TF filter [3][3][ci][c] (float32) of stage s (w_tf[0] unused when cin == 0) -> per-lane B fragments
void block_mma_pack(const BlockMmaPlan& plan, const float* const w_tf[4], int cin, uint32_t* dst);
// bias4: device [4][16] float32 (stage s at bias4 + 16 s)
cudaError_t launch_block_mma(cudaStream_t st, const BlockMmaPlan& plan, PV in, PV out, PV pool, const Geo* pool_geo,
                             const uint32_t* wfrag, const float* bias4, const Geo& g, int* err_flag);

// ---- attention combine + 4x4 classifier head in one launch (combine_head.cu) ---------------------
// out = act(conv4x4(combine(...)) + bias) as dense float32 NHWC; the combined 8-channel map stays in shared memory
bool combine_head_ok(const CombineArgs& c, int ks, int cin, int cout, int act);
void combine_head_pack(const float* w_tf, int cout, uint32_t* dst);   // dst: 16 * 32 words
cudaError_t launch_combine_head(cudaStream_t st, const CombineArgs& c, const uint32_t* wfrag, const float* bias_host,
                                int cout, int act, float* out, int* err_flag);

// host-side 16-bit conversion matching act_t
uint16_t host_f_to_act(float v);

}  // namespace aru
