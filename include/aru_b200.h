/*
 * aru_b200.h - C ABI of libaru_b200.so: the B200 (sm_100a) ARU-Net forward engine.
 *
 * Drop-in boundary for ONE call of the reference (paths relative to the reference repo):
 *   article_separation/image_segmentation/net_post_processing/net_post_processing_helper.py
 *     :36-53  load_graph(path_to_pb)            -> aru_create()  (+ the Python GraphDef loader)
 *     :56-72  get_net_output(image, graph, dev) -> aru_forward() (replaces tf.Session.run)
 * and, one level up, for the integer steps either side of it:
 *     :14-25  scale_image (INTER_AREA / INTER_CUBIC)                -> aru_scale_pages(), aru_separator_images(), aru_heading_images()
 *     :31     cv2.cvtColor(BGR2GRAY) / 255.0                        -> aru_pages_to_input(), aru_separator_pages(), aru_heading_pages()
 *     :75-78  apply_threshold; separator_net_post_processor.py:147  -> the uint8 / mask outputs of aru_forward() / aru_separator_pages()
 *   region_net_post_processor_base.py:230-251  apply_cc_analysis   -> aru_cc_filter()
 *   separator_net_post_processor.py:25-99      post_process        -> aru_separator_post(), aru_separator_pages()
 *   heading_net_post_processor.py:247-270      get_net_prob_for_text_line -> aru_heading_pages(), aru_box_sums()
 * The reference has no FFI of its own (it is Python over TensorFlow 1.x), so these entry points are
 * what a ctypes binding in that helper module would bind; INTEGRATION.md shows the stub.
 *
 * Plain pointers and sizes only; no torch / CUDA types in the signatures (streams are void*).
 * All functions return 0 on success and a non-zero ARU_E* code on failure; aru_last_error()
 * returns the message of the last failure on that engine (or the global one for aru_create).
 * One engine per (process, device); an engine is not re-entrant.
 */
#ifndef ARU_B200_H
#define ARU_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARU_ABI_VERSION 1
#define ARU_PROGRAM_MAGIC 0x31555241u /* "ARU1" */
#define ARU_MAX_SCALES 8

/* error codes */
#define ARU_OK 0
#define ARU_EINVAL 1   /* bad argument / malformed program */
#define ARU_ECUDA 2    /* CUDA runtime or driver error     */
#define ARU_ENOMEM 3   /* device or pinned-host allocation failed */
#define ARU_EUNSUP 4   /* program uses a shape / op the kernels do not cover */
#define ARU_ENODEV 5   /* no CUDA device: there is no CPU fallback by design */

/* op kinds of the lowered program (produced by the Python GraphDef loader, program.py) */
enum aru_op_kind {
  ARU_OP_CONV = 1,    /* Conv2D SAME s1 + BiasAdd [+ residual Add] [+ Relu]; layers.py:191-247, ARU_v1.py:212-227 */
  ARU_OP_DECONV = 2,  /* Conv2DBackpropInput 3x3 s2 SAME + BiasAdd + act; layers.py:342-367                       */
  ARU_OP_MAXPOOL = 3, /* MaxPool 2x2 s2 SAME; layers.py:543-544                                                   */
  ARU_OP_AVGPOOL = 4, /* AvgPool 2x2 s2 SAME (divisor = valid cells); layers.py:526-527                           */
  ARU_OP_COMBINE = 5, /* upsample_simple x A, softmax over scales, Mul, AddN; ARU_v1.py:115,137,145-153           */
  ARU_OP_UPSUM = 6,   /* stand-alone upsample_simple (ones-filter transposed conv); layers.py:716-720             */
  ARU_OP_COPY = 7     /* channel-slice copy (a ConcatV2 input that could not be produced in place)                */
};

/* activations */
enum aru_act { ARU_ACT_NONE = 0, ARU_ACT_RELU = 1, ARU_ACT_SOFTMAX = 2, ARU_ACT_SIGMOID = 3 };

/* a channel slice [ch_off, ch_off+ch) of a buffer; buf < 0 means "absent" */
typedef struct aru_view {
  int32_t buf;
  int32_t ch_off;
  int32_t ch;
} aru_view;

typedef struct aru_op {
  int32_t kind;     /* enum aru_op_kind */
  int32_t ksize;    /* conv / deconv kernel size (3 or 4) */
  int32_t stride;   /* deconv stride (2) / UPSUM factor */
  int32_t act;      /* enum aru_act */
  int32_t n_scales; /* COMBINE: number of attention scales A */
  int32_t like_buf; /* DECONV/UPSUM/COMBINE: buffer whose spatial dims the output takes */
  int64_t w_off;    /* offset (floats) of the filter in the weight blob, TF layout; -1 = none */
  int64_t b_off;    /* offset (floats) of the bias; -1 = none */
  aru_view in;      /* main input */
  aru_view out;     /* main output (post-activation) */
  aru_view out_pre; /* CONV: optional copy of the pre-activation value (bias added, before Relu) */
  aru_view res;     /* CONV: optional residual operand added before the activation */
  aru_view att[ARU_MAX_SCALES]; /* COMBINE: low-resolution 1-channel attention maps */
  aru_view det[ARU_MAX_SCALES]; /* COMBINE: detection feature maps */
  int32_t up_att[ARU_MAX_SCALES]; /* COMBINE: upsample factor of att[i] */
  int32_t up_det[ARU_MAX_SCALES]; /* COMBINE: upsample factor of det[i] (1 = already full size) */
} aru_op;

typedef struct aru_buffer {
  int32_t channels;
  int32_t role; /* 0 = intermediate, 1 = graph input (inImg:0), 2 = graph output (output:0) */
} aru_buffer;

/* the lowered graph handed to aru_create(); all pointers are host memory, copied by the engine */
typedef struct aru_graph_desc {
  uint32_t magic;       /* ARU_PROGRAM_MAGIC */
  uint32_t abi_version; /* ARU_ABI_VERSION */
  int32_t n_buffers;
  int32_t n_ops;
  int64_t n_weights; /* floats in `weights` */
  const aru_buffer* buffers;
  const aru_op* ops;
  const float* weights;
} aru_graph_desc;

typedef struct aru_engine aru_engine;

/* engine options (aru_set_option) */
#define ARU_OPT_CONV_PATH 1   /* 0 = auto (tcgen05: row-banded kernel for C_out <= 16, position-major kernel above - a
                                     fixed rule, so results do not depend on the process; ARU_AUTOTUNE=1 in the
                                     environment times both per layer instead),
                                 1 = force the CUDA-core kernels (validation), 2 = tcgen05 position-major kernel only,
                                 3 = row-banded kernel wherever it applies, one launch per layer */
#define ARU_OPT_USE_GRAPH 2   /* 1 = replay a captured CUDA graph per plan (default), 0 = plain launches */
#define ARU_OPT_MICRO_BATCH 3 /* pages per pass through the net inside aru_forward (0 = auto) */
#define ARU_OPT_KEEP_ALL 4    /* 1 = store every lowered tensor (per-layer checks through aru_read_buffer); 0 (default) =
                                 a tensor read only by a pool that its producer's launch performs is not stored */

#define ARU_OPT_FUSE_PAIRS 5  /* 1 = with ARU_OPT_CONV_PATH 0, two chained 3x3 convolutions with C_out = 8 / 16 (conv1 ->
                                 convR_0, convR_1 -> convR_2 of a residual block, ARU_v1.py:212-227) run as one launch with
                                 the intermediate in a shared-memory row FIFO (conv_band2.cu); 0 (default) = one launch per
                                 layer.  Measured slower than the per-layer kernels on B200 (the banded-weight MMAs make the
                                 shared-memory port the limit), kept selectable and parity-tested.
                                 ARU_FUSE_PAIRS=1 in the environment turns it on as well. */

#define ARU_OPT_U8_CHANNELS 6 /* k > 0: the uint8 outputs (out_u8) of the host-buffer calls hold only the k leading channels,
                                 [n,h,w,k] - every consumer reads channel 0 alone (separator_net_post_processor.py:33,
                                 heading_net_post_processor.py:209), and the device->host copy shrinks with it; 0 = all */

#define ARU_OPT_ASYNC 7       /* 1 = aru_forward / aru_separator_pages / aru_separator_images return as soon as their work is
                                 enqueued; aru_last_ticket() names the call and aru_wait() blocks until its outputs are in
                                 the host buffers (which, like the inputs, must stay alive until then).  Consecutive calls
                                 then overlap: the next call's first copy-in runs under this call's tail.  At most 8 calls
                                 may be in flight.  The heading calls (they return sums by value) always complete. */

#define ARU_OPT_FUSE_BLOCKS 8  /* 1 (default) = with ARU_OPT_CONV_PATH 0 and ARU_OPT_KEEP_ALL 0, a residual block of the
                                 8-channel levels (conv1 -> convR_0 -> convR_1 -> convR_2 + conv1's pre-activation -> ReLU
                                 [-> max-pool], ARU_v1.py:212-227, 266-281) runs as ONE launch on the warp-level tensor
                                 path with its intermediates in shared memory (block_mma.cu); 0 = one launch per layer.
                                 ARU_FUSE_BLOCKS=0 / 1 in the environment overrides the option. */
#define ARU_OPT_BRANCH_STREAMS 9 /* 1 (default) = the independent branches of a pass (the smaller scales of the pyramid, the
                                    attention CNNs) are enqueued on the engine's own low-priority streams and joined before
                                    the attention combine, inside the captured graph as well; 0 = one stream.  Same bits
                                    either way.  ARU_BRANCH_STREAMS=0 / 1 in the environment overrides the option. */

int aru_abi_version(void);
int aru_device_count(void);

/* One process per GPU: pin the calling thread to the CPUs of the NUMA node CUDA device `device` is
 * attached to and prefer that node for new pages, so that the buffers the rank page-locks afterwards are local to its
 * GPU (sysfs local_cpulist / numa_node of the PCI device).  aru_create() calls it unless ARU_NUMA_BIND=0 is set.
 * *numa_node receives the node (-1: unknown); ARU_EUNSUP when the topology cannot be read or applied (nothing changed). */
int aru_bind_host_to_device(int device, int* numa_node);

/* get_net_output receives a float64 page (image / 255.0, net_post_processing_helper.py:31,56-59): dst[i] = (float)src[i]
 * for `count` values on up to `threads` host threads (0 = the process's share of the cores - its affinity mask divided by
 * LOCAL_WORLD_SIZE when a launcher exports it - at most 16) - the one CPU pass
 * over the page before it is copied up; dst is normally a page-locked buffer from aru_host_alloc. */
int aru_f64_to_f32(const double* src, float* dst, long long count, int threads);

/* load_graph: build an engine for `g` on CUDA device `device`. Fails with ARU_ENODEV without a GPU. */
int aru_create(const aru_graph_desc* g, int device, aru_engine** out);
void aru_destroy(aru_engine* e);

int aru_set_option(aru_engine* e, int option, int64_t value);
int aru_num_classes(const aru_engine* e);

/* Allocate the arena and build launch plans / CUDA graph for n pages of h x w. Implicit in aru_forward. */
int aru_plan(aru_engine* e, int n, int h, int w);

/*
 * get_net_output: run the net on n pages of h x w.
 *   in        float32 [n,h,w]   (gray/255), host memory (pageable or pinned)
 *   out       float32 [n,h,w,c] class probabilities, host memory; may be NULL
 *   out_u8    uint8   [n,h,w,c] = trunc(p*255) (separator_net_post_processor.py:147); may be NULL
 *   out_mask  uint8   [n,h,w]   = 255 where out_u8[...,0] > thr*255 else 0 (helper.py:75-78); may be NULL
 * Copies are issued on the engine's stream and the call returns after they completed.
 */
int aru_forward(aru_engine* e, const float* in, int n, int h, int w, float* out, uint8_t* out_u8,
                uint8_t* out_mask, float thr);

/* Same, but `in` / `out*` are DEVICE pointers on the engine's device and the work is only enqueued on
 * `stream` (a cudaStream_t, NULL = the engine's own stream); n must not exceed the planned batch.
 * When `out` is given and `in` / `out` are 16-byte aligned the buffers are bound IN PLACE: the first layers read `in`,
 * the classifier writes `out`, nothing is staged (one captured graph per (in, out) pair, four pairs kept per page
 * shape); `in` is never written.  Otherwise the pages go through the plan's own staging buffers. */
int aru_forward_device(aru_engine* e, const float* in, int n, int h, int w, float* out, uint8_t* out_u8,
                       uint8_t* out_mask, float thr, void* stream);

/*
 * The integer steps either side of the net, on the device, so that only uint8
 * pages go up and only the two separator masks come down:
 *
 * aru_separator_pages = one iteration of SeparatorNetPostProcessor.run up to the polygon step
 *   (separator_net_post_processor.py:141-151):
 *     cv2.cvtColor(BGR2GRAY) / 255.0          (net_post_processing_helper.py:31; channels == 3, else gray pages)
 *     get_net_output                          (helper.py:56-72)
 *     np.array(net_output * 255, uint8)       (sep:147)
 *     apply_threshold(net_output, threshold)  (sep:149, helper.py:75-78; `thr` is the Python double)
 *     post_process                            (sep:25-99: component size filter base:230-251, cv2 MORPH_OPEN with
 *                                              (int(15W/1000) x 1) and (1 x int(30H/1500)), cv2.subtract, MORPH_OPEN
 *                                              (int(10W/1000) x 1))
 *   pages      uint8 [n,h,w,channels] host memory (channels 1 = gray, 3 = B,G,R as cv2.imread returns them)
 *   out        float32 [n,h,w,c] or NULL;  out_u8  uint8 [n,h,w,c] or NULL;  out_mask  uint8 [n,h,w] or NULL
 *   horizontal / vertical   uint8 [n,h,w] in {0,255}, both or neither (NULL): post_process()["horizontal"/"vertical"]
 * Fails with ARU_EINVAL when a structuring element would be empty (W < 100 or H < 50): OpenCV raises there too.
 */
int aru_separator_pages(aru_engine* e, const uint8_t* pages, int channels, int n, int h, int w, double thr, float* out,
                        uint8_t* out_u8, uint8_t* out_mask, uint8_t* horizontal, uint8_t* vertical);

/*
 * aru_heading_pages = one iteration of HeadingNetPostProcessor.run up to the per-text-line network feature
 *   (heading_net_post_processor.py:280-291 and get_net_prob_for_text_line :247-270): colour step, net,
 *   np.array(net_output * 255, uint8), and for every text-line bounding box the sum of channel 0 of that uint8 map
 *   over net_output[ya:yb, xa:xb].  The reference's value is sums[i] / 255 / (box width * box height).
 *   boxes   int32 [n_boxes][5] = page index, y0, y1, x0, x1: half-open ranges already clipped to the page (numpy slice
 *           semantics are the caller's, see Engine.heading_pages); a box outside its page is ARU_EINVAL
 *   sums    uint64 [n_boxes] exact integer sums;  out_u8  uint8 [n,h,w,c] or NULL
 */
int aru_heading_pages(aru_engine* e, const uint8_t* pages, int channels, int n, int h, int w, const int32_t* boxes,
                      int n_boxes, uint64_t* sums, uint8_t* out_u8);

/* scale_image + aru_heading_pages in one call; boxes are in the coordinates of the scaled page. */
int aru_heading_images(aru_engine* e, const uint8_t* images, int channels, int n, int src_h, int src_w, double sc,
                       const int32_t* boxes, int n_boxes, uint64_t* sums, uint8_t* out_u8);

/* The box sums alone on a host uint8 map [n,h,w,c] (channel 0 is summed). */
int aru_box_sums(aru_engine* e, const uint8_t* u8, int n, int h, int w, int c, const int32_t* boxes, int n_boxes,
                 uint64_t* sums);

/*
 * scale_image on the device (net_post_processing_helper.py:14-25):
 *   sc < 1  cv2.resize(image, None, fx=sc, fy=sc, interpolation=cv2.INTER_AREA), bit-exact against OpenCV (general and
 *           integer scales);
 *   sc > 1  cv2.resize(..., interpolation=cv2.INTER_CUBIC) in OpenCV's scalar fixed-point form: the destination size is
 *           OpenCV's, the values are within ONE grey level of cv2.resize (its SIMD builds evaluate the vertical pass in
 *           float; ~6 % of the pixels differ by 1) - the stated tolerance of the enlarging path.
 * The destination size is (cvRound(src_h*sc), cvRound(src_w*sc)) = aru_scaled_size().
 *   aru_scale_pages       uint8 [n,src_h,src_w,channels] -> uint8 [n,h,w,channels], host memory in and out
 *   aru_separator_images  = scale_image + aru_separator_pages in one call: the unscaled uint8 images go up, the
 *                           outputs have the scaled size (sc == 1: no resize, as in the reference)
 */
int aru_scaled_size(int src_h, int src_w, double sc, int* h, int* w);
int aru_scale_pages(aru_engine* e, const uint8_t* pages, int channels, int n, int src_h, int src_w, double sc, uint8_t* out);
int aru_separator_images(aru_engine* e, const uint8_t* images, int channels, int n, int src_h, int src_w, double sc,
                         double thr, float* out, uint8_t* out_u8, uint8_t* out_mask, uint8_t* horizontal, uint8_t* vertical);

/* post_process alone on n thresholded masks (uint8 [n,h,w], non-zero = foreground), host memory in and out. */
int aru_separator_post(aru_engine* e, const uint8_t* mask, int n, int h, int w, uint8_t* horizontal, uint8_t* vertical);

/* RegionNetPostProcessor.apply_cc_analysis (region_net_post_processor_base.py:230-251) alone - also the whole of
 * TextBlockNetPostProcessor.post_process (text_block_net_post_processor.py:12-24): keep the 8-connected components of the
 * non-zero pixels whose area is >= min_size; uint8 [n,h,w] host memory in, {0,255} out.  The reference's limit is
 * int(net_output.size * threshold) with Python float arithmetic - computed by the caller (net_boundary.apply_cc_analysis). */
int aru_cc_filter(aru_engine* e, const uint8_t* mask, int n, int h, int w, int min_size, uint8_t* out);

/* cv2.morphologyEx(mask, MORPH_OPEN, getStructuringElement(MORPH_RECT, (kw, kh))) for kw == 1 or kh == 1 on binary
 * masks (non-zero = 255), host memory in and out; exposed for the parity tests of the morphology kernels. */
int aru_open_rect(aru_engine* e, const uint8_t* mask, int n, int h, int w, int kw, int kh, uint8_t* out);

/*
 * StrokeWidthDistanceTransform.distance_transform (python_util/image_processing/swt_dist_trafo.py:18-29) - the SWT feature
 * image HeadingNetPostProcessor computes twice per page at full image resolution (heading_net_post_processor.py:86,297):
 * 255 - gray (dark_on_bright), cv2.GaussianBlur 5x5 sigma 0, cv2.threshold OTSU, cv2.distanceTransform(DIST_L2,
 * DIST_MASK_PRECISE), astype(uint8).  Bit-exact against the reference's function (the decode, cv2.imread(GRAYSCALE),
 * stays with the caller).
 *   gray        uint8 [n,h,w] host memory;   out  uint8 [n,h,w];   thresholds  int32 [n] Otsu thresholds, or NULL
 */
int aru_swt_distance(aru_engine* e, const uint8_t* gray, int n, int h, int w, int dark_on_bright, uint8_t* out,
                     int32_t* thresholds);

/* The colour step alone: uint8 pages [n,h,w,channels] -> float32 [n,h,w] = gray / 255.0 (helper.py:31). */
int aru_pages_to_input(aru_engine* e, const uint8_t* pages, int channels, int n, int h, int w, float* out);

/* ARU_OPT_ASYNC: ticket of the most recent host-buffer call, and completion of one call. */
uint64_t aru_last_ticket(const aru_engine* e);
int aru_wait(aru_engine* e, uint64_t ticket);

/* Block until everything enqueued on the engine's stream has finished. */
int aru_sync(aru_engine* e);

/* Introspection used by tests / bench: number of kernels one forward of the current plan launches,
 * and a copy-out of an intermediate buffer as float32 NHWC (debugging / per-layer parity). */
int aru_launches_per_forward(const aru_engine* e);
int aru_read_buffer(aru_engine* e, int buf, int n_index, float* out_nhwc, size_t out_floats);
int aru_buffer_dims(const aru_engine* e, int buf, int* h, int* w, int* c);

/* Per-op timing of the current plan (CUDA events on the engine stream, plain launches):
 * ms[i] receives the mean duration of op i over `iters` runs; n_ms must be >= n_ops. */
int aru_profile_ops(aru_engine* e, int iters, float* ms, int n_ms);
/* Kernel label ("conv_tc", "conv_direct", ...) the plan chose for op i. */
const char* aru_op_kernel_name(const aru_engine* e, int op);

const char* aru_last_error(const aru_engine* e);

/* Dynamic-range check: the first host-buffer pass of an engine on real data is followed by one scan of every stored
 * tensor for 16-bit values at the storage limit (fp16 stores saturate at +-65504) or not finite.  A non-empty string
 * says how many were found and where; it is also printed to stderr once.  The synthetic nets stay below 40; a trained
 * ReLU net without normalisation could exceed the limit - then the bf16 build is the one to use.  ARU_RANGE_CHECK=0
 * in the environment skips the scan. */
const char* aru_last_warning(const aru_engine* e);

/* Page-locked host memory for the in / out arrays of aru_forward(): with pinned buffers the
 * host<->device copies of consecutive micro-batches overlap the kernels (pageable memory works too,
 * but serialises). */
int aru_host_alloc(void** ptr, size_t bytes);
void aru_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* ARU_B200_H */
