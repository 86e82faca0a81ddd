// engine.cu - C ABI of libaru_b200.so (include/aru_b200.h): engine creation from the lowered op
// program, per-shape planning (arena, launch list, CUDA graphs) and the pipelined forward pass that
// replaces the reference's  sess.run('output:0', {'inImg:0': x})  (net_post_processing_helper.py:56-72).
//
// Threading model: one engine per (process, device), calls are not re-entrant (the reference calls the
// net from a plain per-page loop, separator_net_post_processor.py:141).  There is no CPU fallback: without
// a CUDA device aru_create fails with ARU_ENODEV.
#include <cuda_runtime.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <tuple>
#include <string>
#include <thread>
#include <vector>

#include "../../include/aru_b200.h"
#include "aru_common.cuh"
#include "kernels.h"

using namespace aru;

namespace {

thread_local std::string g_error;  // message of a failed aru_create

enum BufKind { KIND_PLANAR = 0, KIND_F32 = 1, KIND_OUT = 2 };

struct OpWeights {
  act_t* w16 = nullptr;     // direct conv / deconv pack [tap][cin_chunk][cout_chunk][8][8]
  // tcgen05 B images, keyed by (first input chunk, number of input chunks): the whole filter, or the
  // K-parts of a split launch (built lazily at plan time)
  struct TcImage { int ci_chunk_begin, ci_chunks; act_t* w; };
  std::vector<TcImage> tc_images;
  act_t* band_w = nullptr;  // banded masters of conv_band.cu (depend on the filter only)
  size_t band_bytes = 0;
  uint32_t* blk_w = nullptr;   // block_mma.cu: B fragments of the residual block this op ends (with / without conv1)
  float* blk_bias = nullptr;   // [4][16]
  int blk_cin = -1;            // conv1 input channels the fragments were packed for (0: block starts from x0)
  uint32_t* head_w = nullptr;  // combine_head.cu: B fragments of a classifier that runs in the combine's launch
  float* w32 = nullptr;     // stem [tap][cpad] / small [tap][cin_pad][cout]
  float* bias = nullptr;    // zero-padded to a multiple of 16
  int cin = 0, cout = 0;
};

struct BufPlan {
  int kind = KIND_PLANAR;
  int channels = 0, chunks = 0;
  int h = 0, w = 0;
  bool sized = false;
  Geo geo{};
  size_t offset = 0;  // bytes into the arena
  size_t bytes = 0;
};

struct TcPart {
  ConvTcPlan plan;
  int ci_chunk_begin = 0, ci_chunks = 0;
  const act_t* w = nullptr;
};

struct Plan {
  int n = 0, h = 0, w = 0;
  std::vector<BufPlan> bufs;
  std::vector<std::vector<TcPart>> tc;  // per op: K-parts of the tensor-core launch (empty: other kernel)
  std::vector<ConvBandPlan> band;       // per op: row-banded tensor-core launch (ok == false: other kernel)
  std::vector<char> use_band;           // per op: the row-banded kernel won the plan-time timing (or is the only one)
  std::vector<int> fused_pool;          // per conv op: the MAXPOOL op its row-banded launch also performs (-1: none)
  std::vector<char> skip;               // per op: performed by another op's launch
  std::vector<char> pool_only;          // per stem conv op with a fused pool: nothing else reads its full-resolution output
  std::vector<int> pair_first;          // per conv op k: the conv op i whose launch it absorbs (conv_band2.cu), -1: none
  std::vector<ConvBand2Plan> pair;      // per conv op k with pair_first[k] >= 0
  std::vector<char> pair_store0;        // per conv op k: the stage-0 (op i) activated output must also be stored
  // whole residual blocks as one launch (block_mma.cu): per conv op c = convR_2 of a block, its conv1 / convR_0 / convR_1
  std::vector<int> blk_conv1, blk_r0, blk_r1;   // -1: op c does not end a fused block
  std::vector<BlockMmaPlan> blk;
  std::vector<char> blk_from_x0;        // per op c: conv1 runs as its own launch and the block starts from its pre-activation
  std::vector<char> pre_only;           // per conv1 op of such a block: only the pre-activation is stored
  // Independent branches of a pass (the smaller scales of the pyramid, the attention CNNs) run on their own streams:
  // op_stream[i] = 0 (the caller's stream) or 1 + aux stream; op_event[i] is recorded after op i when an op of another
  // stream reads what it wrote; op_waits[i] = the ops whose events op i waits for.
  bool branched = false;
  std::vector<int> op_stream;
  std::vector<cudaEvent_t> op_event;
  std::vector<std::vector<int>> op_waits;
  cudaEvent_t ev_start = nullptr, ev_end[3] = {nullptr, nullptr, nullptr};
  std::vector<int> head_combine;        // per classifier conv op: the COMBINE op its launch also performs (-1: none)
  act_t* scratch = nullptr;             // partial sums of split-K launches
  size_t scratch_bytes = 0;
  std::vector<const char*> kernel;   // per op label
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<char> buf_dead;   // per buffer: no launch of this plan touches it (an intermediate of a fused launch): not in the arena
  // [0], [1]: the plan's own double-buffered staging; [2 .. 2 + EXT_SLOTS): caller-owned device buffers bound in place
  // by aru_forward_device (no staging copy; one captured graph per binding, least recently used slot recycled)
  static constexpr int EXT_SLOTS = 4;
  float* in_dev[2 + EXT_SLOTS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float* out_dev[2 + EXT_SLOTS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  unsigned long long ext_use[2 + EXT_SLOTS] = {0, 0, 0, 0, 0, 0};
  unsigned long long ext_clock = 0;
  uint8_t* u8_dev[2] = {nullptr, nullptr};
  uint8_t* mask_dev[2] = {nullptr, nullptr};
  uint8_t* page_dev[2] = {nullptr, nullptr};   // uint8 gray / BGR pages (aru_separator_pages)
  int page_dev_channels = 0;
  uint8_t* src_dev[2] = {nullptr, nullptr};    // unscaled uint8 pages (aru_*_images: cv2.resize on the device)
  size_t src_dev_bytes = 0;
  uint8_t* hor_dev[2] = {nullptr, nullptr};    // separator post-processing results
  uint8_t* ver_dev[2] = {nullptr, nullptr};
  void* post_scratch = nullptr;                // labels, areas, bit planes (used on the compute stream only)
  cudaGraphExec_t graph[2 + EXT_SLOTS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  bool d2h_pending[2] = {false, false};
  int launches = 0;
  uint64_t last_use = 0;
  int eager_passes[2 + EXT_SLOTS] = {0, 0, 0, 0, 0, 0};   // passes enqueued without a graph, per parity (the graph is captured on the second)
  Plan() = default;
  Plan(const Plan&) = delete;
  Plan& operator=(const Plan&) = delete;
  ~Plan();   // releases the arena, staging buffers, events and graphs (every error return of build_plan relies on it)
};

}  // namespace

struct aru_engine {
  int device = 0;
  int num_sms = 0;
  size_t max_smem = 0;
  std::vector<aru_buffer> buffers;
  std::vector<aru_op> ops;
  std::vector<float> weights;
  std::vector<OpWeights> opw;
  std::vector<int> kind;  // per buffer
  int input_buf = -1, output_buf = -1, n_class = 0;
  int conv_path = 0, use_graph = 1, micro_batch = 0, keep_all = 0, fuse_pairs = 0, fuse_blocks = 1, u8_channels = 0, async_calls = 0, branch_streams = 1;
  // ARU_OPT_ASYNC: completion events of the host-buffer calls in flight (ticket t lives in tickets[t % 8])
  cudaEvent_t tickets[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t last_ticket = 0;
  cudaStream_t s_comp = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  static constexpr int N_AUX = 3;
  cudaStream_t s_aux[N_AUX] = {nullptr, nullptr, nullptr};   // branch streams of a pass (low priority), see plan_branches
  int* err_flag = nullptr;
  float* zero_bias = nullptr;
  std::vector<std::unique_ptr<Plan>> plans;
  // plan-time kernel choice per (op, page height, page width): independent of the batch size, so that a page gives
  // bit-identical results whatever micro-batch it travels in
  std::map<std::tuple<int, int, int>, char> tune_cache;
  Plan* cur = nullptr;
  uint64_t tick = 0;
  std::string error;
  // dynamic-range check of the first real pass (fp16 stores saturate at 65504: real weights could exceed what the
  // synthetic nets do); ARU_RANGE_CHECK=0 skips it
  int range_checked = 0;
  std::string warning;
};

namespace {

int fail(aru_engine* e, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->error = buf; else g_error = buf;
  return code;
}

#define CU(e, call)                                                                                      \
  do {                                                                                                   \
    cudaError_t err__ = (call);                                                                          \
    if (err__ != cudaSuccess)                                                                            \
      return fail(e, err__ == cudaErrorMemoryAllocation ? ARU_ENOMEM : ARU_ECUDA, "%s failed: %s (%s:%d)", #call, \
                  cudaGetErrorString(err__), __FILE__, __LINE__);                                        \
  } while (0)

template <typename T>
int upload(aru_engine* e, const std::vector<T>& host, T** dev) {
  CU(e, cudaMalloc((void**)dev, std::max<size_t>(host.size() * sizeof(T), 16)));
  CU(e, cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return ARU_OK;
}

bool view_ok(const aru_engine* e, const aru_view& v) {
  return v.buf >= 0 && v.buf < (int)e->buffers.size() && v.ch > 0 && v.ch_off >= 0 &&
         v.ch_off + v.ch <= e->buffers[v.buf].channels;
}

// ---- weight packing ------------------------------------------------------------------------------
int pack_op_weights(aru_engine* e, int oi) {
  const aru_op& op = e->ops[oi];
  OpWeights& ow = e->opw[oi];
  if (op.kind != ARU_OP_CONV && op.kind != ARU_OP_DECONV) return ARU_OK;
  const int ks = op.ksize, cin = op.in.ch, cout = op.out.ch;
  ow.cin = cin;
  ow.cout = cout;
  const size_t nw = (size_t)ks * ks * cin * cout;
  if (op.w_off < 0 || op.b_off < 0 || (size_t)op.w_off + nw > e->weights.size() ||
      (size_t)op.b_off + cout > e->weights.size())
    return fail(e, ARU_EINVAL, "op %d: weight / bias offsets out of range", oi);
  const float* w = e->weights.data() + op.w_off;
  const float* b = e->weights.data() + op.b_off;
  const int cinc = cdiv(cin, 8), coutc = cdiv(cout, 8);
  std::vector<float> bias(std::max(16, cdiv(coutc * 8, 16) * 16), 0.f);
  for (int i = 0; i < cout; ++i) bias[i] = b[i];
  int rc = upload(e, bias, &ow.bias);
  if (rc) return rc;
  const int ik = e->kind[op.in.buf], okind = e->kind[op.out.buf];
  if (op.kind == ARU_OP_CONV && ik == KIND_F32) {
    // stem: [tap][16] (filter is [kh][kw][1][cout], cout <= 16)
    const int cpad = 16;
    if (cout > 16) return fail(e, ARU_EUNSUP, "op %d: 1-channel convolution with %d > 16 output channels", oi, cout);
    std::vector<float> p((size_t)ks * ks * cpad, 0.f);
    for (int t = 0; t < ks * ks; ++t)
      for (int c = 0; c < cout; ++c) p[(size_t)t * cpad + c] = w[(size_t)t * cout + c];
    return upload(e, p, &ow.w32);
  }
  if (op.kind == ARU_OP_CONV && okind != KIND_PLANAR) {
    // small: [tap][cin_pad][cout]
    const int cpad = cinc * 8;
    std::vector<float> p((size_t)ks * ks * cpad * cout, 0.f);
    for (int t = 0; t < ks * ks; ++t)
      for (int ci = 0; ci < cin; ++ci)
        for (int c = 0; c < cout; ++c) p[((size_t)t * cpad + ci) * cout + c] = w[((size_t)t * cin + ci) * cout + c];
    return upload(e, p, &ow.w32);
  }
  // chunk-planar -> chunk-planar: direct pack [tap][cin_chunk][cout_chunk][ci][co]
  std::vector<uint16_t> p((size_t)ks * ks * cinc * coutc * 64, host_f_to_act(0.f));
  for (int t = 0; t < ks * ks; ++t)
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co) {
        // Conv2D filter [kh][kw][cin][cout]; Conv2DBackpropInput filter [kh][kw][cout][cin] (layers.py:356)
        const float v = op.kind == ARU_OP_CONV ? w[((size_t)t * cin + ci) * cout + co] : w[((size_t)t * cout + co) * cin + ci];
        p[((((size_t)t * cinc + ci / 8) * coutc + co / 8) * 8 + ci % 8) * 8 + co % 8] = host_f_to_act(v);
      }
  rc = upload(e, p, reinterpret_cast<uint16_t**>(&ow.w16));
  if (rc) return rc;
  return ARU_OK;
}

// B image of input chunks [begin, begin+chunks) of op oi for the tensor-core kernel (cached per engine)
int get_tc_image(aru_engine* e, int oi, const ConvTcPlan& tp, int begin, int chunks, const act_t** out) {
  OpWeights& ow = e->opw[oi];
  for (auto& im : ow.tc_images)
    if (im.ci_chunk_begin == begin && im.ci_chunks == chunks) { *out = im.w; return ARU_OK; }
  const aru_op& op = e->ops[oi];
  std::vector<uint16_t> img(tp.wpack_bytes / 2);
  if (tp.deconv)
    conv_tc_pack_deconv(tp, e->weights.data() + op.w_off, op.in.ch, op.out.ch, img.data());
  else
    conv_tc_pack_weights(tp, e->weights.data() + op.w_off, op.in.ch, op.out.ch, begin * 8, img.data());
  uint16_t* dev = nullptr;
  int rc = upload(e, img, &dev);
  if (rc) return rc;
  ow.tc_images.push_back({begin, chunks, reinterpret_cast<act_t*>(dev)});
  *out = reinterpret_cast<act_t*>(dev);
  return ARU_OK;
}

// ---- planning ------------------------------------------------------------------------------------
// Idempotent: every pointer is cleared, so the explicit calls on error paths and the destructor can both run.
void free_plan(Plan* p) {
  if (!p) return;
  auto dev_free = [](auto*& q) { if (q) { cudaFree(q); q = nullptr; } };
  auto ev_free = [](cudaEvent_t& ev) { if (ev) { cudaEventDestroy(ev); ev = nullptr; } };
  for (int i = 0; i < 2; ++i) {
    if (p->graph[i]) { cudaGraphExecDestroy(p->graph[i]); p->graph[i] = nullptr; }
    dev_free(p->in_dev[i]); dev_free(p->out_dev[i]); dev_free(p->u8_dev[i]); dev_free(p->mask_dev[i]);
    dev_free(p->page_dev[i]); dev_free(p->src_dev[i]); dev_free(p->hor_dev[i]); dev_free(p->ver_dev[i]);
    ev_free(p->ev_h2d[i]); ev_free(p->ev_comp[i]); ev_free(p->ev_d2h[i]);
  }
  for (int i = 2; i < 2 + Plan::EXT_SLOTS; ++i) {   // bindings of caller-owned buffers: only the graphs are ours
    if (p->graph[i]) { cudaGraphExecDestroy(p->graph[i]); p->graph[i] = nullptr; }
    p->in_dev[i] = p->out_dev[i] = nullptr;
  }
  for (cudaEvent_t& ev : p->op_event) ev_free(ev);
  ev_free(p->ev_start);
  for (cudaEvent_t& ev : p->ev_end) ev_free(ev);
  dev_free(p->arena); dev_free(p->scratch); dev_free(p->post_scratch);
  p->src_dev_bytes = 0;
}

}  // namespace
Plan::~Plan() { free_plan(this); }
namespace {

int set_dims(aru_engine* e, Plan* p, int buf, int h, int w, int op) {
  BufPlan& b = p->bufs[buf];
  if (b.sized) {
    if (b.h != h || b.w != w)
      return fail(e, ARU_EINVAL, "op %d: buffer %d written with two shapes (%dx%d vs %dx%d)", op, buf, b.h, b.w, h, w);
    return ARU_OK;
  }
  b.h = h; b.w = w; b.sized = true;
  return ARU_OK;
}

PV make_pv(const aru_engine* e, const Plan* p, const aru_view& v) {
  PV pv;
  if (v.buf < 0) return pv;
  const BufPlan& b = p->bufs[v.buf];
  pv.plane = b.geo.plane;
  pv.chunks = cdiv(v.ch, 8);
  pv.C = v.ch;
  pv.p = reinterpret_cast<act_t*>(p->arena + b.offset) + (long long)(v.ch_off / 8) * b.geo.plane * 8;
  return pv;
}

float* f32_ptr(const aru_engine* e, const Plan* p, int buf, int parity) {
  if (buf == e->input_buf) return p->in_dev[parity];
  if (buf == e->output_buf) return p->out_dev[parity];
  return reinterpret_cast<float*>(p->arena + p->bufs[buf].offset);
}

CombineArgs combine_args(const aru_engine* e, const Plan* p, const aru_op& op, int parity) {
  CombineArgs a{};
  const BufPlan& bo = p->bufs[op.out.buf];
  a.A = op.n_scales;
  a.geo = bo.geo;
  PV out = make_pv(e, p, op.out);
  a.out = out.p; a.out_plane = out.plane; a.out_chunks = out.chunks;
  for (int k = 0; k < op.n_scales; ++k) {
    const BufPlan& ba = p->bufs[op.att[k].buf];
    a.att[k] = f32_ptr(e, p, op.att[k].buf, parity);
    a.att_h[k] = ba.h; a.att_w[k] = ba.w; a.att_up[k] = op.up_att[k];
    a.att_oy[k] = (ba.h * op.up_att[k] - bo.h) / 2;
    a.att_ox[k] = (ba.w * op.up_att[k] - bo.w) / 2;
    const BufPlan& bd = p->bufs[op.det[k].buf];
    PV det = make_pv(e, p, op.det[k]);
    a.det[k] = det.p; a.det_plane[k] = det.plane; a.det_chunks[k] = det.chunks; a.det_geo[k] = bd.geo;
    a.det_up[k] = op.up_det[k];
    a.det_oy[k] = (bd.h * op.up_det[k] - bo.h) / 2;
    a.det_ox[k] = (bd.w * op.up_det[k] - bo.w) / 2;
  }
  return a;
}

// Enqueue op `oi` of plan p on stream st.
int run_op(aru_engine* e, Plan* p, int oi, int parity, cudaStream_t st) {
  const aru_op& op = e->ops[oi];
  const OpWeights& ow = e->opw[oi];
  cudaError_t err = cudaSuccess;
  const char* label = "?";
  switch (op.kind) {
    case ARU_OP_CONV: {
      const BufPlan& bi = p->bufs[op.in.buf];
      const BufPlan& bo = p->bufs[op.out.buf];
      if (bi.kind == KIND_F32) {
        const int pj = p->fused_pool[oi];
        label = pj >= 0 ? (p->pool_only[oi] ? "conv_stem_pool" : "conv_stem_fpool") : "conv_stem";
        PV stem_out = make_pv(e, p, op.out);
        if (p->pre_only[oi]) { stem_out.p = nullptr; label = "conv_stem_pre"; }   // the fused block reads the pre-activation only
        err = launch_conv_stem(st, op.ksize, f32_ptr(e, p, op.in.buf, parity), stem_out,
                               make_pv(e, p, op.out_pre), e->weights.data() + op.w_off, e->weights.data() + op.b_off,
                               bo.geo, op.act, pj >= 0 ? make_pv(e, p, e->ops[pj].out) : PV(),
                               pj >= 0 ? &p->bufs[e->ops[pj].out.buf].geo : nullptr, !p->pool_only[oi], ow.w32, ow.bias);
      } else if (bo.kind != KIND_PLANAR && p->head_combine[oi] >= 0) {
        label = "combine_head";   // attention combine + classifier, the combined map stays in shared memory
        err = launch_combine_head(st, combine_args(e, p, e->ops[p->head_combine[oi]], parity), ow.head_w,
                                  e->weights.data() + op.b_off, op.out.ch, op.act, f32_ptr(e, p, op.out.buf, parity), e->err_flag);
      } else if (bo.kind != KIND_PLANAR && p->band[oi].ok && p->use_band[oi]) {
        label = "conv_band_head";
        err = launch_conv_band(st, p->band[oi], make_pv(e, p, op.in), PV(), PV(), PV(), ow.band_w, ow.bias, bi.geo, op.act,
                               e->err_flag, f32_ptr(e, p, op.out.buf, parity), op.out.ch);
      } else if (bo.kind != KIND_PLANAR && !p->tc[oi].empty()) {
        label = "conv_tc_head";
        err = launch_conv_tc_f32(st, p->tc[oi][0].plan, make_pv(e, p, op.in), f32_ptr(e, p, op.out.buf, parity), op.out.ch,
                                 p->tc[oi][0].w, ow.bias, bi.geo, op.act, e->err_flag);
      } else if (bo.kind != KIND_PLANAR) {
        label = "conv_small";
        err = launch_conv_small(st, op.ksize, op.out.ch, make_pv(e, p, op.in), f32_ptr(e, p, op.out.buf, parity), ow.w32,
                                ow.bias, bi.geo, op.act);
      } else if (p->skip[oi]) {
        label = "pair_fused";   // performed by the launch of the convolution that consumes it (conv_band2.cu, block_mma.cu)
      } else if (p->blk_conv1[oi] >= 0) {
        // the whole residual block in one launch, at the position of its last convolution
        const aru_op& o1 = e->ops[p->blk_conv1[oi]];
        const int pj = p->fused_pool[oi];
        label = pj >= 0 ? "block_mma_pool" : "block_mma";
        err = launch_block_mma(st, p->blk[oi], p->blk_from_x0[oi] ? make_pv(e, p, o1.out_pre) : make_pv(e, p, o1.in),
                               make_pv(e, p, op.out), pj >= 0 ? make_pv(e, p, e->ops[pj].out) : PV(),
                               pj >= 0 ? &p->bufs[e->ops[pj].out.buf].geo : nullptr, ow.blk_w, ow.blk_bias, bo.geo, e->err_flag);
      } else if (p->pair_first[oi] >= 0) {
        // two chained convolutions in one launch: op `fi` (stage 0) feeds this op through a shared-memory row FIFO
        const int fi = p->pair_first[oi];
        const aru_op& o0 = e->ops[fi];
        const int pj = p->fused_pool[oi];
        label = pj >= 0 ? "conv_band2_pool" : "conv_band2";
        err = launch_conv_band2(st, p->pair[oi], make_pv(e, p, o0.in), p->pair_store0[oi] ? make_pv(e, p, o0.out) : PV(),
                                make_pv(e, p, o0.out_pre), make_pv(e, p, op.out), make_pv(e, p, op.res), e->opw[fi].band_w,
                                e->opw[fi].bias, o0.act, ow.band_w, ow.bias, op.act, bo.geo, e->err_flag,
                                pj >= 0 ? make_pv(e, p, e->ops[pj].out) : PV(),
                                pj >= 0 ? &p->bufs[e->ops[pj].out.buf].geo : nullptr);
      } else if (p->band[oi].ok && p->use_band[oi]) {
        const int pj = p->fused_pool[oi];
        label = pj >= 0 ? "conv_band_pool" : "conv_band";
        err = launch_conv_band(st, p->band[oi], make_pv(e, p, op.in), make_pv(e, p, op.out), make_pv(e, p, op.out_pre),
                               make_pv(e, p, op.res), ow.band_w, ow.bias, bo.geo, op.act, e->err_flag, nullptr, 0,
                               pj >= 0 ? make_pv(e, p, e->ops[pj].out) : PV(),
                               pj >= 0 ? &p->bufs[e->ops[pj].out.buf].geo : nullptr);
      } else if (!p->tc[oi].empty()) {
        // K-parts chain through the scratch buffer: part 0 adds the op's residual, the last part adds the
        // bias, applies the activation and writes the real outputs
        const auto& parts = p->tc[oi];
        label = parts.size() == 1 ? (parts[0].plan.w_stream ? "conv_tc_ws" : "conv_tc") : "conv_tc_splitk";
        PV scratch;
        scratch.p = p->scratch; scratch.plane = bo.geo.plane; scratch.chunks = cdiv(op.out.ch, 8); scratch.C = op.out.ch;
        for (size_t k = 0; k < parts.size() && err == cudaSuccess; ++k) {
          const bool first = k == 0, last = k + 1 == parts.size();
          PV in = make_pv(e, p, op.in);
          in.p += (long long)parts[k].ci_chunk_begin * in.plane * 8;
          in.chunks = parts[k].ci_chunks;
          err = launch_conv_tc(st, parts[k].plan, in, last ? make_pv(e, p, op.out) : scratch,
                               last ? make_pv(e, p, op.out_pre) : PV(), first ? make_pv(e, p, op.res) : scratch,
                               parts[k].w, last ? ow.bias : e->zero_bias, bo.geo, last ? op.act : ARU_ACT_NONE, e->err_flag);
        }
      } else {
        label = "conv_direct";
        err = launch_conv_direct(st, op.ksize, make_pv(e, p, op.in), make_pv(e, p, op.out), make_pv(e, p, op.out_pre),
                                 make_pv(e, p, op.res), ow.w16, ow.bias, bo.geo, op.act);
      }
      break;
    }
    case ARU_OP_DECONV:
      if (!p->tc[oi].empty()) {
        label = "deconv_tc";
        err = launch_deconv_tc(st, p->tc[oi][0].plan, make_pv(e, p, op.in), make_pv(e, p, op.out), p->tc[oi][0].w, ow.bias,
                               p->bufs[op.in.buf].geo, p->bufs[op.out.buf].geo, op.act, e->err_flag);
      } else {
        label = "deconv_direct";
        err = launch_deconv_direct(st, make_pv(e, p, op.in), p->bufs[op.in.buf].geo, make_pv(e, p, op.out),
                                   p->bufs[op.out.buf].geo, ow.w16, ow.bias, op.act);
      }
      break;
    case ARU_OP_MAXPOOL:
    case ARU_OP_AVGPOOL: {
      const BufPlan& bi = p->bufs[op.in.buf];
      const BufPlan& bo = p->bufs[op.out.buf];
      if (p->skip[oi]) {
        label = "pool_fused";   // written by the producing convolution's launch
      } else if (bi.kind == KIND_F32) {
        label = "pool_f32";
        err = launch_pool_f32(st, op.kind == ARU_OP_MAXPOOL, f32_ptr(e, p, op.in.buf, parity),
                              f32_ptr(e, p, op.out.buf, parity), p->n, bi.h, bi.w, bo.h, bo.w);
      } else {
        label = "pool";
        err = launch_pool(st, op.kind == ARU_OP_MAXPOOL, make_pv(e, p, op.in), bi.geo, make_pv(e, p, op.out), bo.geo);
      }
      break;
    }
    case ARU_OP_COMBINE: {
      if (p->skip[oi]) {
        label = "head_fused";   // performed by the classifier's launch (combine_head.cu)
        break;
      }
      label = "combine";
      err = launch_combine(st, combine_args(e, p, op, parity));
      break;
    }
    case ARU_OP_UPSUM: {
      const BufPlan& bi = p->bufs[op.in.buf];
      const BufPlan& bo = p->bufs[op.out.buf];
      const int oy = (bi.h * op.stride - bo.h) / 2, ox = (bi.w * op.stride - bo.w) / 2;
      if (bi.kind == KIND_F32) {
        label = "upsum_f32";
        err = launch_upsum_f32(st, f32_ptr(e, p, op.in.buf, parity), bi.h, bi.w, f32_ptr(e, p, op.out.buf, parity), p->n,
                               bo.h, bo.w, op.stride, oy, ox);
      } else {
        label = "upsum";
        err = launch_upsum(st, make_pv(e, p, op.in), bi.geo, make_pv(e, p, op.out), bo.geo, op.stride, oy, ox);
      }
      break;
    }
    case ARU_OP_COPY: {
      label = "copy";
      PV in = make_pv(e, p, op.in), out = make_pv(e, p, op.out);
      for (int c = 0; c < in.chunks && err == cudaSuccess; ++c)
        err = launch_copy(st, in.p + c * in.plane * 8, out.p + c * out.plane * 8, in.plane * 16);
      break;
    }
    default:
      return fail(e, ARU_EUNSUP, "op %d: unknown kind %d", oi, op.kind);
  }
  p->kernel[oi] = label;
  if (err != cudaSuccess) return fail(e, ARU_ECUDA, "op %d (%s): launch failed: %s", oi, label, cudaGetErrorString(err));
  return ARU_OK;
}

int run_all(aru_engine* e, Plan* p, int parity, cudaStream_t st) {
  const int no = (int)e->ops.size();
  if (!p->branched) {
    for (int i = 0; i < no; ++i) {
      int rc = run_op(e, p, i, parity, st);
      if (rc) return rc;
    }
    return ARU_OK;
  }
  // fork: the branch streams start after everything `st` holds so far (the previous pass included) ...
  CU(e, cudaEventRecord(p->ev_start, st));
  for (int a = 0; a < aru_engine::N_AUX; ++a) CU(e, cudaStreamWaitEvent(e->s_aux[a], p->ev_start, 0));
  for (int i = 0; i < no; ++i) {
    cudaStream_t s = p->op_stream[i] ? e->s_aux[p->op_stream[i] - 1] : st;
    for (int w : p->op_waits[i]) CU(e, cudaStreamWaitEvent(s, p->op_event[w], 0));
    int rc = run_op(e, p, i, parity, s);
    if (rc) return rc;
    if (p->op_event[i]) CU(e, cudaEventRecord(p->op_event[i], s));
  }
  // ... and join: the pass is complete on `st` when every branch is
  for (int a = 0; a < aru_engine::N_AUX; ++a) {
    CU(e, cudaEventRecord(p->ev_end[a], e->s_aux[a]));
    CU(e, cudaStreamWaitEvent(st, p->ev_end[a], 0));
  }
  return ARU_OK;
}

// Stream assignment of a pass.  The pyramid scales of the shared-weight RU-Net and the attention CNNs are independent
// until the attention combine (ARU_v1.py:105-153); run one after the other, the small scales spend most of their time in
// launch prologues and tails (a sixteenth of the pixels takes 1.25 of 15.65 ms).  On streams of their own their kernels
// fill the SMs the full-resolution chain leaves idle.  Scale k of the detection branch -> stream k mod 3 (scale 0 = the
// caller's stream), ops that only feed attention operands -> the last branch stream, everything else -> the caller's.
// Dependencies across streams become events.  Buffers are written once per pass (no aliasing), so there are no other
// hazards; the split-K scratch is only used on the caller's stream (checked).
int plan_branches(aru_engine* e, Plan* p) {
  const int no = (int)e->ops.size(), nb = (int)e->buffers.size();
  p->op_stream.assign(no, 0);
  p->op_event.assign(no, nullptr);
  p->op_waits.assign(no, {});
  p->branched = false;
  const char* env = getenv("ARU_BRANCH_STREAMS");
  if (!(env ? env[0] == '1' : e->branch_streams != 0) || e->keep_all || e->conv_path != 0) return ARU_OK;
  int c = -1;
  for (int i = 0; i < no && c < 0; ++i)
    if (e->ops[i].kind == ARU_OP_COMBINE) c = i;
  if (c < 0 || e->ops[c].n_scales < 2) return ARU_OK;
  auto reads = [&](const aru_op& op, std::vector<int>& out) {
    out.clear();
    if (op.kind == ARU_OP_COMBINE) {
      for (int k = 0; k < op.n_scales; ++k) { out.push_back(op.att[k].buf); out.push_back(op.det[k].buf); }
    } else {
      if (op.in.buf >= 0) out.push_back(op.in.buf);
      if (op.res.buf >= 0) out.push_back(op.res.buf);
    }
  };
  // producers of every op's operands (program order = a topological order)
  std::vector<int> writer(nb, -1), rd;
  std::vector<std::vector<int>> deps(no);
  for (int i = 0; i < no; ++i) {
    reads(e->ops[i], rd);
    for (int b : rd)
      if (writer[b] >= 0) deps[i].push_back(writer[b]);
    if (e->ops[i].out.buf >= 0) writer[e->ops[i].out.buf] = i;
    if (e->ops[i].out_pre.buf >= 0) writer[e->ops[i].out_pre.buf] = i;
  }
  // ancestors of the combine's operands: bit k = feeds det[k], bit 16 = feeds an attention operand
  std::vector<unsigned> feeds(no, 0u);
  {
    std::vector<int> w2(nb, -1);
    for (int i = 0; i < c; ++i) {
      if (e->ops[i].out.buf >= 0) w2[e->ops[i].out.buf] = i;
      if (e->ops[i].out_pre.buf >= 0) w2[e->ops[i].out_pre.buf] = i;
    }
    const aru_op& oc = e->ops[c];
    for (int k = 0; k < oc.n_scales; ++k) {
      if (w2[oc.det[k].buf] >= 0) feeds[w2[oc.det[k].buf]] |= 1u << std::min(k, 15);
      if (w2[oc.att[k].buf] >= 0) feeds[w2[oc.att[k].buf]] |= 1u << 16;
    }
    for (int i = c - 1; i >= 0; --i)
      for (int d : deps[i]) feeds[d] |= feeds[i];
  }
  bool any = false;
  for (int i = 0; i < c; ++i) {
    const unsigned det = feeds[i] & 0xffffu;
    int s = 0;
    if (det) { int k = 0; while (!((det >> k) & 1u)) ++k; s = k % 3; }      // the first scale it feeds
    else if (feeds[i] & (1u << 16)) s = aru_engine::N_AUX;                   // attention only
    p->op_stream[i] = s;
    any = any || s != 0;
  }
  if (!any) return ARU_OK;
  for (int i = 0; i < no; ++i)
    if (p->tc[i].size() > 1) p->op_stream[i] = 0;   // split-K launches share one scratch buffer: the caller's stream only
  // events for the edges that cross streams; an op without a launch of its own must not be such a producer unless the
  // launch that performs its work comes earlier on the same stream (fused pools), which stream order already covers
  for (int i = 0; i < no; ++i)
    for (int d : deps[i]) {
      if (p->op_stream[d] == p->op_stream[i]) continue;
      if (p->skip[d]) { p->op_stream.assign(no, 0); p->op_waits.assign(no, {}); return ARU_OK; }   // fused away: stay serial
      if (std::find(p->op_waits[i].begin(), p->op_waits[i].end(), d) == p->op_waits[i].end()) p->op_waits[i].push_back(d);
    }
  for (int i = 0; i < no; ++i)
    for (int d : p->op_waits[i])
      if (!p->op_event[d]) CU(e, cudaEventCreateWithFlags(&p->op_event[d], cudaEventDisableTiming));
  CU(e, cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
  for (cudaEvent_t& ev : p->ev_end) CU(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  p->branched = true;
  return ARU_OK;
}

int check_err_flag(aru_engine* e) {
  int flag = 0;
  CU(e, cudaMemcpy(&flag, e->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    cudaMemset(e->err_flag, 0, sizeof(int));
    return fail(e, ARU_ECUDA, "conv_tc pipeline timeout (barrier class %d): protocol error in the tcgen05 kernel", flag);
  }
  return ARU_OK;
}

int build_plan(aru_engine* e, int n, int h, int w, Plan** out) {
  // ARU_PLAN_TRACE=1: where the time of meeting a new shape goes (stderr)
  static const bool trace = [] { const char* v = getenv("ARU_PLAN_TRACE"); return v && v[0] == '1'; }();
  const auto t_begin = std::chrono::steady_clock::now();
  auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  };
  std::unique_ptr<Plan> p(new Plan);
  p->n = n; p->h = h; p->w = w;
  const int nb = (int)e->buffers.size(), no = (int)e->ops.size();
  p->bufs.resize(nb);
  p->tc.resize(no);
  p->band.resize(no);
  p->use_band.assign(no, 0);
  p->fused_pool.assign(no, -1);
  p->skip.assign(no, 0);
  p->pool_only.assign(no, 0);
  p->pair_first.assign(no, -1);
  p->pair.resize(no);
  p->pair_store0.assign(no, 0);
  p->blk_conv1.assign(no, -1);
  p->blk_r0.assign(no, -1);
  p->blk_r1.assign(no, -1);
  p->blk.resize(no);
  p->blk_from_x0.assign(no, 0);
  p->pre_only.assign(no, 0);
  p->head_combine.assign(no, -1);
  p->kernel.assign(no, "?");
  for (int i = 0; i < nb; ++i) {
    p->bufs[i].kind = e->kind[i];
    p->bufs[i].channels = e->buffers[i].channels;
    p->bufs[i].chunks = cdiv(e->buffers[i].channels, 8);
  }
  int rc = set_dims(e, p.get(), e->input_buf, h, w, -1);
  if (rc) return rc;
  // shape inference in program order (the lowering emits ops topologically)
  for (int i = 0; i < no; ++i) {
    const aru_op& op = e->ops[i];
    auto need = [&](int buf) -> bool { return buf >= 0 && p->bufs[buf].sized; };
    switch (op.kind) {
      case ARU_OP_CONV:
      case ARU_OP_COPY: {
        if (!need(op.in.buf)) return fail(e, ARU_EINVAL, "op %d reads buffer %d before it is produced", i, op.in.buf);
        const BufPlan& bi = p->bufs[op.in.buf];
        if ((rc = set_dims(e, p.get(), op.out.buf, bi.h, bi.w, i))) return rc;
        if (op.out_pre.buf >= 0 && (rc = set_dims(e, p.get(), op.out_pre.buf, bi.h, bi.w, i))) return rc;
        if (op.res.buf >= 0) {
          if (!need(op.res.buf) || p->bufs[op.res.buf].h != bi.h || p->bufs[op.res.buf].w != bi.w)
            return fail(e, ARU_EINVAL, "op %d: residual operand shape mismatch", i);
        }
        break;
      }
      case ARU_OP_MAXPOOL:
      case ARU_OP_AVGPOOL: {
        if (!need(op.in.buf)) return fail(e, ARU_EINVAL, "op %d reads buffer %d before it is produced", i, op.in.buf);
        const BufPlan& bi = p->bufs[op.in.buf];
        if ((rc = set_dims(e, p.get(), op.out.buf, cdiv(bi.h, 2), cdiv(bi.w, 2), i))) return rc;
        break;
      }
      case ARU_OP_DECONV:
      case ARU_OP_UPSUM: {
        if (!need(op.in.buf) || !need(op.like_buf)) return fail(e, ARU_EINVAL, "op %d: operands not produced yet", i);
        const BufPlan& bi = p->bufs[op.in.buf];
        const BufPlan& bl = p->bufs[op.like_buf];
        if (bi.h != cdiv(bl.h, op.stride) || bi.w != cdiv(bl.w, op.stride))
          return fail(e, ARU_EINVAL, "op %d: transposed conv input %dx%d is not ceil(%dx%d / %d)", i, bi.h, bi.w, bl.h,
                      bl.w, op.stride);
        if ((rc = set_dims(e, p.get(), op.out.buf, bl.h, bl.w, i))) return rc;
        break;
      }
      case ARU_OP_COMBINE: {
        if (!need(op.like_buf)) return fail(e, ARU_EINVAL, "op %d: like buffer not produced yet", i);
        const BufPlan& bl = p->bufs[op.like_buf];
        for (int k = 0; k < op.n_scales; ++k) {
          if (!need(op.att[k].buf) || !need(op.det[k].buf))
            return fail(e, ARU_EINVAL, "op %d: attention operands not produced yet", i);
          const BufPlan& ba = p->bufs[op.att[k].buf];
          const BufPlan& bd = p->bufs[op.det[k].buf];
          if (ba.h != cdiv(bl.h, op.up_att[k]) || ba.w != cdiv(bl.w, op.up_att[k]) ||
              bd.h != cdiv(bl.h, op.up_det[k]) || bd.w != cdiv(bl.w, op.up_det[k]))
            return fail(e, ARU_EINVAL, "op %d: scale %d operand sizes do not match the upsample factors", i, k);
        }
        if ((rc = set_dims(e, p.get(), op.out.buf, bl.h, bl.w, i))) return rc;
        break;
      }
      default:
        return fail(e, ARU_EUNSUP, "op %d: unknown kind %d", i, op.kind);
    }
  }
  // arena: every tensor a launch of this plan touches gets a region of its own (no aliasing: the zero frame of a tensor
  // must stay intact); the intermediates of the fused launches (block_mma.cu, combine_head.cu) are never materialised
  // and get none - they are known once the fusion decisions below are taken, so the arena is laid out after them
  // (at once when the plan-time tuner has to run launches)
  p->buf_dead.assign(nb, 0);
  for (int i = 0; i < nb; ++i) {
    BufPlan& b = p->bufs[i];
    b.offset = 0;
    if (!b.sized) { b.bytes = 0; continue; }
    b.geo = make_geo(n, b.h, b.w);
    b.bytes = b.kind == KIND_PLANAR ? (size_t)b.chunks * b.geo.plane * 16 : (size_t)n * b.h * b.w * b.channels * sizeof(float);
  }
  auto layout_arena = [&]() -> int {
    if (p->arena) return ARU_OK;
    size_t off = 0;
    for (int i = 0; i < nb; ++i) {
      BufPlan& b = p->bufs[i];
      if (!b.sized || p->buf_dead[i] || i == e->input_buf || i == e->output_buf) continue;   // in / out: separate, double-buffered
      b.offset = off;
      off += (b.bytes + 255) / 256 * 256;
    }
    p->arena_bytes = std::max<size_t>(off, 256);
    CU(e, cudaMalloc((void**)&p->arena, p->arena_bytes));
    CU(e, cudaMemsetAsync(p->arena, 0, p->arena_bytes, e->s_comp));
    return ARU_OK;
  };
  // a view's buffer is dead when the view covers it and every op that names it is one of `ops` (the members of a fused launch)
  auto mark_dead = [&](const aru_view& v, std::initializer_list<int> ops) {
    if (v.buf < 0 || v.buf == e->input_buf || v.buf == e->output_buf) return;
    if (v.ch_off != 0 || v.ch != e->buffers[v.buf].channels) return;
    for (int j = 0; j < no; ++j) {
      const aru_op& o = e->ops[j];
      bool names = o.in.buf == v.buf || o.res.buf == v.buf || o.out.buf == v.buf || o.out_pre.buf == v.buf;
      if (o.kind == ARU_OP_COMBINE)
        for (int k = 0; k < o.n_scales; ++k) names = names || o.att[k].buf == v.buf || o.det[k].buf == v.buf;
      if (names && std::find(ops.begin(), ops.end(), j) == ops.end()) return;
    }
    p->buf_dead[v.buf] = 1;
  };
  const double t_shapes = ms_since(t_begin);
  {
    const char* tune_env = getenv("ARU_AUTOTUNE");
    if (tune_env && tune_env[0] == '1' && (rc = layout_arena())) { free_plan(p.get()); return rc; }
  }
  const size_t in_bytes = (size_t)n * h * w * sizeof(float);
  const size_t out_bytes = (size_t)n * h * w * e->n_class * sizeof(float);
  for (int i = 0; i < 2; ++i) {
    CU(e, cudaMalloc((void**)&p->in_dev[i], in_bytes));
    CU(e, cudaMalloc((void**)&p->out_dev[i], out_bytes));
    CU(e, cudaMemsetAsync(p->in_dev[i], 0, in_bytes, e->s_comp));
    CU(e, cudaEventCreateWithFlags(&p->ev_h2d[i], cudaEventDisableTiming));
    CU(e, cudaEventCreateWithFlags(&p->ev_comp[i], cudaEventDisableTiming));
    CU(e, cudaEventCreateWithFlags(&p->ev_d2h[i], cudaEventDisableTiming));
  }
  const double t_alloc = ms_since(t_begin);
  // row-banded tensor-core plan of conv op i (+ its banded weight masters, cached per engine)
  // banded weight masters of conv op i (they depend on the filter and on (ks, C_in, C_out) only; cached per engine)
  auto band_weights = [&](int i, const ConvBandPlan& bp) -> int {
    const aru_op& op = e->ops[i];
    OpWeights& ow = e->opw[i];
    if (!ow.band_w || ow.band_bytes != bp.wpack_bytes) {
      if (ow.band_w) { cudaFree(ow.band_w); ow.band_w = nullptr; }
      std::vector<uint16_t> img(bp.wpack_bytes / 2);
      conv_band_pack(bp, e->weights.data() + op.w_off, op.in.ch, op.out.ch, img.data());
      uint16_t* dev = nullptr;
      int rc2 = upload(e, img, &dev);
      if (rc2) return rc2;
      ow.band_w = reinterpret_cast<act_t*>(dev);
      ow.band_bytes = bp.wpack_bytes;
    }
    return ARU_OK;
  };
  auto plan_band = [&](int i, const Geo& geo) -> int {
    const aru_op& op = e->ops[i];
    ConvBandPlan bp = conv_band_plan(op.ksize, op.in.ch, op.out.ch, geo, e->num_sms, e->max_smem);
    if (!bp.ok) return ARU_OK;
    int rc2 = band_weights(i, bp);
    if (rc2) return rc2;
    p->band[i] = bp;
    p->use_band[i] = 1;
    return ARU_OK;
  };
  // kernel selection for the chunk-planar convolutions
  for (int i = 0; i < no; ++i) {
    const aru_op& op = e->ops[i];
    if (e->conv_path == 1) continue;
    if (op.kind == ARU_OP_DECONV) {
      ConvTcPlan tp = conv_tc_plan(3, op.in.ch, op.out.ch, p->bufs[op.in.buf].geo, e->num_sms, e->max_smem, true);
      if (tp.ok) {
        TcPart part;
        part.plan = tp;
        part.ci_chunks = cdiv(op.in.ch, 8);
        if ((rc = get_tc_image(e, i, tp, -1, part.ci_chunks, &part.w))) { free_plan(p.get()); return rc; }
        p->tc[i].push_back(part);
      }
      continue;
    }
    if (op.kind != ARU_OP_CONV) continue;
    if (p->bufs[op.in.buf].kind != KIND_PLANAR) continue;
    if (p->bufs[op.out.buf].kind != KIND_PLANAR) {
      // float32-output head (attention logit, classifier): one launch, no split-K
      if ((e->conv_path == 0 || e->conv_path == 3) && op.out.ch <= 4 && op.in.ch <= 8) {   // the classifier head
        if ((rc = plan_band(i, p->bufs[op.in.buf].geo))) { free_plan(p.get()); return rc; }
        if (p->band[i].ok && e->conv_path == 3) continue;
      }
      ConvTcPlan tp = conv_tc_plan(op.ksize, op.in.ch, op.out.ch, p->bufs[op.in.buf].geo, e->num_sms, e->max_smem);
      if (tp.ok && tp.cout_chunks == 1) {
        TcPart part;
        part.plan = tp;
        part.ci_chunks = cdiv(op.in.ch, 8);
        if ((rc = get_tc_image(e, i, tp, 0, part.ci_chunks, &part.w))) { free_plan(p.get()); return rc; }
        p->tc[i].push_back(part);
      }
      continue;
    }
    const Geo& geo = p->bufs[op.out.buf].geo;
    const int cin_chunks = cdiv(op.in.ch, 8);
    if (e->conv_path == 0 || e->conv_path == 3) {
      // small C_out: output rows x channels on the MMA M axis (HBM bound instead of tensor-issue bound)
      if ((rc = plan_band(i, geo))) { free_plan(p.get()); return rc; }
      if (p->band[i].ok && e->conv_path == 3) continue;   // forced: no position-major alternative is planned
    }
    for (int parts = 1; parts <= 4 && p->tc[i].empty(); parts *= 2) {
      if (cin_chunks % parts || (parts > 1 && (cin_chunks / parts) % 2)) break;
      const int pc = cin_chunks / parts;
      ConvTcPlan tp = conv_tc_plan(op.ksize, parts == 1 ? op.in.ch : pc * 8, op.out.ch, geo, e->num_sms, e->max_smem);
      if (!tp.ok) continue;
      for (int k = 0; k < parts; ++k) {
        TcPart part;
        part.plan = tp;
        part.ci_chunk_begin = k * pc;
        part.ci_chunks = pc;
        if ((rc = get_tc_image(e, i, tp, part.ci_chunk_begin, pc, &part.w))) { free_plan(p.get()); return rc; }
        p->tc[i].push_back(part);
      }
      if (parts > 1) p->scratch_bytes = std::max<size_t>(p->scratch_bytes, (size_t)cdiv(op.out.ch, 8) * (size_t)geo.plane * 16);
    }
  }
  if (p->scratch_bytes) CU(e, cudaMalloc((void**)&p->scratch, p->scratch_bytes));
  // Ops both tensor-core kernels cover.
  {
    // Default: a fixed rule (row-banded kernel for C_out <= 16, position-major above), so that a page gives the same
    // bits in every process, rank and run - the two kernels round differently and masks near the 13/255 cut would
    // flip.  ARU_AUTOTUNE=1 times both on this plan's geometry and keeps the faster (benchmarking only).
    const char* env = getenv("ARU_AUTOTUNE");
    const bool tune = env && env[0] == '1';
    // The rule is a function of the op's signature only (never of the batch or the page size: a page must give the same
    // bits alone and inside a batch).  Row-banded for C_out <= 16 - except the two 16-channel shapes whose extra stream
    // loads the banded kernel's store warps: a residual operand, or a pre-activation export behind 8 input channels;
    // those run position-major (measured at 32 x 1856x1344, profiles/r02q_profile_autotune.txt: 0.41 -> 0.33 ms,
    // 0.37 -> 0.30 ms, 0.48 -> 0.33 + 0.12 ms for the separate pool; -0.3 ms per pass).
    auto fixed_rule = [&](int i) -> char {
      const aru_op& op = e->ops[i];
      if (p->band[i].cop > 16) return 0;
      if (p->band[i].cop == 16 && (op.res.buf >= 0 || (op.out_pre.buf >= 0 && op.in.ch <= 8))) return 0;
      return 1;
    };
    for (int i = 0; i < no && !tune; ++i)
      if (p->band[i].ok && !p->tc[i].empty()) p->use_band[i] = fixed_rule(i);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    for (int i = 0; i < no && tune; ++i) {
      if (!p->band[i].ok || p->tc[i].empty()) continue;
      const auto key = std::make_tuple(i, h, w);
      const auto hit = e->tune_cache.find(key);
      if (hit != e->tune_cache.end()) { p->use_band[i] = hit->second; continue; }
      {
        // small launches are timing noise: fixed rule (keeps small-page results identical from process to process)
        const Geo& og = p->bufs[e->ops[i].in.buf].geo;
        if ((long long)og.N * og.H * og.W < (1LL << 21)) {
          p->use_band[i] = fixed_rule(i);
          e->tune_cache[key] = p->use_band[i];
          continue;
        }
      }
      if (!ev0) { CU(e, cudaEventCreate(&ev0)); CU(e, cudaEventCreate(&ev1)); }
      float best[2] = {1e30f, 1e30f};
      for (int which = 0; which < 2; ++which) {
        p->use_band[i] = (char)which;
        for (int rep = 0; rep < 3; ++rep) {
          CU(e, cudaEventRecord(ev0, e->s_comp));
          if ((rc = run_op(e, p.get(), i, 0, e->s_comp))) { free_plan(p.get()); return rc; }
          CU(e, cudaEventRecord(ev1, e->s_comp));
          CU(e, cudaEventSynchronize(ev1));
          float ms = 0.f;
          CU(e, cudaEventElapsedTime(&ms, ev0, ev1));
          if (rep > 0) best[which] = std::min(best[which], ms);
        }
      }
      p->use_band[i] = best[1] < 1.03f * best[0] ? 1 : 0;   // ties go to the row-banded kernel
      e->tune_cache[key] = p->use_band[i];
    }
    if (ev0) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); }
    if ((rc = check_err_flag(e))) { free_plan(p.get()); return rc; }
  }
  // Whole residual blocks of the 8-channel levels (conv1 -> convR_0 -> convR_1 -> convR_2 + conv1's pre-activation,
  // ARU_v1.py:212-227, 266-281) run as ONE launch on the warp-level tensor path (block_mma.cu); the intermediates are
  // never stored.  The launch sits at convR_2's position.  conv1 joins the launch when its input is an 8 / 16 channel
  // tensor; a conv1 that reads the float32 page (the stem) stays its own launch and stores its pre-activation only.
  {
    const char* env = getenv("ARU_FUSE_BLOCKS");
    const bool fuse = e->conv_path == 0 && !e->keep_all && (env ? env[0] == '1' : e->fuse_blocks != 0);
    auto reads = [&](const aru_op& ro, int buf) {
      if (ro.in.buf == buf || ro.res.buf == buf) return true;
      if (ro.kind == ARU_OP_COMBINE)
        for (int a = 0; a < ro.n_scales; ++a)
          if (ro.att[a].buf == buf || ro.det[a].buf == buf) return true;
      return false;
    };
    auto conv3 = [&](const aru_op& o) {
      return o.kind == ARU_OP_CONV && o.ksize == 3 && p->bufs[o.out.buf].kind == KIND_PLANAR && o.act == ARU_ACT_RELU;
    };
    // the only reader of view v after op i, or -1
    auto sole_reader = [&](int i, const aru_view& v) {
      int k = -1, cnt = 0;
      for (int j = i + 1; j < no; ++j)
        if (reads(e->ops[j], v.buf)) { if (k < 0) k = j; ++cnt; }
      return cnt == 1 ? k : -1;
    };
    auto same_view = [](const aru_view& a, const aru_view& b) { return a.buf == b.buf && a.ch_off == b.ch_off && a.ch == b.ch; };
    for (int i = 0; i < no && fuse; ++i) {
      const aru_op& o1 = e->ops[i];
      if (!conv3(o1) || o1.out_pre.buf < 0 || o1.res.buf >= 0 || o1.out.ch != 8) continue;
      if (o1.out.buf == e->output_buf || o1.out_pre.buf == e->output_buf) continue;
      const int ia = sole_reader(i, o1.out);
      if (ia < 0) continue;
      const aru_op& oa = e->ops[ia];
      if (!conv3(oa) || !same_view(oa.in, o1.out) || oa.res.buf >= 0 || oa.out_pre.buf >= 0 || oa.out.ch != 8) continue;
      const int ib = sole_reader(ia, oa.out);
      if (ib < 0) continue;
      const aru_op& ob = e->ops[ib];
      if (!conv3(ob) || !same_view(ob.in, oa.out) || ob.res.buf >= 0 || ob.out_pre.buf >= 0 || ob.out.ch != 8) continue;
      const int ic = sole_reader(ib, ob.out);
      if (ic < 0) continue;
      const aru_op& oc = e->ops[ic];
      if (!conv3(oc) || !same_view(oc.in, ob.out) || !same_view(oc.res, o1.out_pre) || oc.out_pre.buf >= 0 || oc.out.ch != 8)
        continue;
      if (sole_reader(i, o1.out_pre) != ic) continue;
      if (oa.out.buf == e->output_buf || ob.out.buf == e->output_buf) continue;
      const bool planar_in = p->bufs[o1.in.buf].kind == KIND_PLANAR && (o1.in.ch == 8 || o1.in.ch == 16) && o1.in.ch_off % 8 == 0;
      const int cin = planar_in ? o1.in.ch : 0;
      BlockMmaPlan bp = block_mma_plan(cin, 8, p->bufs[oc.out.buf].geo, e->num_sms, e->max_smem);
      if (!bp.ok) continue;
      OpWeights& ow = e->opw[ic];
      if (!ow.blk_w || ow.blk_cin != cin) {
        if (ow.blk_w) { cudaFree(ow.blk_w); ow.blk_w = nullptr; }
        if (ow.blk_bias) { cudaFree(ow.blk_bias); ow.blk_bias = nullptr; }
        std::vector<uint32_t> img(bp.wfrag_bytes / 4, 0u);
        const float* wt[4] = {e->weights.data() + o1.w_off, e->weights.data() + oa.w_off, e->weights.data() + ob.w_off,
                              e->weights.data() + oc.w_off};
        block_mma_pack(bp, wt, cin, img.data());
        std::vector<float> b4(64, 0.f);
        const long long offs[4] = {o1.b_off, oa.b_off, ob.b_off, oc.b_off};
        for (int s2 = 0; s2 < 4; ++s2)
          for (int c2 = 0; c2 < 8; ++c2) b4[s2 * 16 + c2] = e->weights[(size_t)offs[s2] + c2];
        if ((rc = upload(e, img, &ow.blk_w)) || (rc = upload(e, b4, &ow.blk_bias))) { free_plan(p.get()); return rc; }
        ow.blk_cin = cin;
      }
      p->blk[ic] = bp;
      p->blk_conv1[ic] = i;
      p->blk_r0[ic] = ia;
      p->blk_r1[ic] = ib;
      p->blk_from_x0[ic] = cin == 0;
      p->skip[ia] = p->skip[ib] = 1;
      if (cin > 0) p->skip[i] = 1;
      else if (p->bufs[o1.in.buf].kind == KIND_F32) p->pre_only[i] = 1;
      if (!p->arena) {   // the intermediates of the launch stay in shared memory
        mark_dead(oa.out, {i, ia, ib, ic});
        mark_dead(ob.out, {i, ia, ib, ic});
        if (cin > 0 || p->pre_only[i]) mark_dead(o1.out, {i, ia, ib, ic});
        if (cin > 0) mark_dead(o1.out_pre, {i, ia, ib, ic});
      }
    }
  }
  // The attention combine runs inside the launch of the classifier that is its only reader (combine_head.cu).
  {
    const char* env = getenv("ARU_FUSE_HEAD");
    const bool fuse = e->conv_path == 0 && !e->keep_all && (env ? env[0] == '1' : e->fuse_blocks != 0);
    for (int c = 0; c < no && fuse; ++c) {
      const aru_op& oc = e->ops[c];
      if (oc.kind != ARU_OP_COMBINE || oc.out.buf == e->output_buf) continue;
      int h = -1, readers = 0;
      for (int j = c + 1; j < no; ++j) {
        const aru_op& ro = e->ops[j];
        bool rd = ro.in.buf == oc.out.buf || ro.res.buf == oc.out.buf;
        if (ro.kind == ARU_OP_COMBINE)
          for (int k = 0; k < ro.n_scales; ++k) rd = rd || ro.att[k].buf == oc.out.buf || ro.det[k].buf == oc.out.buf;
        if (rd) { if (h < 0) h = j; ++readers; }
      }
      if (readers != 1) continue;
      const aru_op& oh = e->ops[h];
      if (oh.kind != ARU_OP_CONV || oh.in.buf != oc.out.buf || oh.in.ch_off != oc.out.ch_off || oh.in.ch != oc.out.ch ||
          oh.res.buf >= 0 || oh.out_pre.buf >= 0 || p->bufs[oh.out.buf].kind == KIND_PLANAR)
        continue;
      if (!combine_head_ok(combine_args(e, p.get(), oc, 0), oh.ksize, oh.in.ch, oh.out.ch, oh.act)) continue;
      OpWeights& ow = e->opw[h];
      if (!ow.head_w) {
        std::vector<uint32_t> img(16 * 32, 0u);
        combine_head_pack(e->weights.data() + oh.w_off, oh.out.ch, img.data());
        if ((rc = upload(e, img, &ow.head_w))) { free_plan(p.get()); return rc; }
      }
      p->head_combine[h] = c;
      p->skip[c] = 1;
      if (!p->arena) mark_dead(oc.out, {c, h});
    }
  }
  // Pairs of chained 3x3 convolutions with C_out = 8 / 16 run as one launch (conv_band2.cu): greedy in program order,
  // op i -> its only consumer k.  The launch happens at k's position, so nothing between i and k may read i's outputs.
  const char* fuse_env = getenv("ARU_FUSE_PAIRS");
  if (e->conv_path == 0 && (e->fuse_pairs || (fuse_env && fuse_env[0] == '1'))) {
    const bool fuse = true;
    auto reads = [&](const aru_op& ro, int buf) {
      if (ro.in.buf == buf || ro.res.buf == buf) return true;
      if (ro.kind == ARU_OP_COMBINE)
        for (int a = 0; a < ro.n_scales; ++a)
          if (ro.att[a].buf == buf || ro.det[a].buf == buf) return true;
      return false;
    };
    auto planar_conv3 = [&](const aru_op& o) {
      return o.kind == ARU_OP_CONV && o.ksize == 3 && p->bufs[o.in.buf].kind == KIND_PLANAR &&
             p->bufs[o.out.buf].kind == KIND_PLANAR && (o.act == ARU_ACT_NONE || o.act == ARU_ACT_RELU);
    };
    for (int i = 0; i < no && fuse; ++i) {
      const aru_op& o0 = e->ops[i];
      if (!planar_conv3(o0) || o0.res.buf >= 0 || p->skip[i] || p->pair_first[i] >= 0 || p->blk_conv1[i] >= 0) continue;
      // consumers of o0.out
      int k = -1, n_readers = 0;
      for (int j = i + 1; j < no; ++j)
        if (reads(e->ops[j], o0.out.buf)) { if (k < 0) k = j; ++n_readers; }
      if (k < 0) continue;
      const aru_op& o1 = e->ops[k];
      if (p->blk_conv1[k] >= 0 || p->skip[k]) continue;
      if (!planar_conv3(o1) || o1.out_pre.buf >= 0 || o1.in.buf != o0.out.buf || o1.in.ch_off != o0.out.ch_off ||
          o1.in.ch != o0.out.ch || o1.out.ch != o0.out.ch || o1.res.buf == o0.out.buf)
        continue;
      bool clean = true;
      for (int j = i + 1; j < k && clean; ++j)
        clean = !(o0.out_pre.buf >= 0 && reads(e->ops[j], o0.out_pre.buf));
      if (!clean) continue;
      ConvBand2Plan pp = conv_band2_plan(o0.in.ch, o0.out.ch, o1.out.ch, p->bufs[o1.out.buf].geo, e->num_sms, e->max_smem);
      if (!pp.ok) continue;
      if ((rc = band_weights(i, pp.st0)) || (rc = band_weights(k, pp.st1))) { free_plan(p.get()); return rc; }
      p->pair[k] = pp;
      p->pair_first[k] = i;
      p->pair_store0[k] = (n_readers > 1 || e->keep_all || o0.out.buf == e->output_buf) ? 1 : 0;
      p->skip[i] = 1;
    }
  }
  // max-pools whose input is the ReLU'd output of a row-banded convolution are written by that launch
  {
    const char* env = getenv("ARU_FUSE_POOL");
    const bool fuse = !(env && env[0] == '0');
    for (int j = 0; j < no && fuse; ++j) {
      const aru_op& po = e->ops[j];
      if (po.kind != ARU_OP_MAXPOOL || p->bufs[po.in.buf].kind != KIND_PLANAR) continue;
      for (int i = j - 1; i >= 0; --i) {
        const aru_op& co = e->ops[i];
        if (co.out.buf != po.in.buf) continue;
        if (co.kind == ARU_OP_CONV && co.out.ch_off == po.in.ch_off && co.out.ch == po.in.ch &&
            (p->pair_first[i] >= 0 || p->blk_conv1[i] >= 0) && p->fused_pool[i] < 0 && co.act == ARU_ACT_RELU) {
          p->fused_pool[i] = j;
          p->skip[j] = 1;
        } else if (co.kind == ARU_OP_CONV && co.out.ch_off == po.in.ch_off && co.out.ch == po.in.ch && p->band[i].ok &&
            p->use_band[i] && p->fused_pool[i] < 0 && p->bufs[co.out.buf].kind == KIND_PLANAR && !p->skip[i] &&
            conv_band_can_pool(p->band[i], co.act, co.in.ch, co.ksize)) {
          p->fused_pool[i] = j;
          p->skip[j] = 1;
        } else if (co.kind == ARU_OP_CONV && co.out.ch_off == po.in.ch_off && co.out.ch == po.in.ch &&
                   p->bufs[co.in.buf].kind == KIND_F32 && p->bufs[co.out.buf].kind == KIND_PLANAR && co.act == ARU_ACT_RELU &&
                   co.out_pre.buf < 0 && co.out.ch <= 16 && p->fused_pool[i] < 0) {
          // stem convolution (1-channel float32 input, CUDA cores): the launch pools too, and when nothing but the pool
          // reads the full-resolution tensor it is not stored at all (ARU_OPT_KEEP_ALL = 1 stores it anyway)
          p->fused_pool[i] = j;
          p->skip[j] = 1;
          bool other_reader = false;
          for (int k = 0; k < no && !other_reader; ++k) {
            if (k == j || k == i) continue;
            const aru_op& ro = e->ops[k];
            other_reader = ro.in.buf == co.out.buf || ro.res.buf == co.out.buf;
            for (int a = 0; a < ARU_MAX_SCALES && !other_reader; ++a)
              other_reader = ro.att[a].buf == co.out.buf || ro.det[a].buf == co.out.buf;
          }
          p->pool_only[i] = !other_reader && !e->keep_all && co.out.buf != e->output_buf;
        }
        break;   // the last writer of the pool's input decides
      }
    }
  }
  const double t_kernels = ms_since(t_begin);
  if ((rc = layout_arena())) { free_plan(p.get()); return rc; }
  if ((rc = plan_branches(e, p.get()))) { free_plan(p.get()); return rc; }
  // one eager pass (sets function attributes, validates every launch), then capture
  rc = run_all(e, p.get(), 0, e->s_comp);
  if (rc) { free_plan(p.get()); return rc; }
  {
    cudaError_t err = cudaStreamSynchronize(e->s_comp);
    if (err != cudaSuccess) {
      free_plan(p.get());
      return fail(e, ARU_ECUDA, "first pass of the plan failed: %s", cudaGetErrorString(err));
    }
  }
  if ((rc = check_err_flag(e))) { free_plan(p.get()); return rc; }
  p->launches = no;
  for (int i = 0; i < no; ++i)
    if (e->ops[i].kind == ARU_OP_COPY) p->launches += cdiv(e->ops[i].in.ch, 8) - 1;
    else if (p->tc[i].size() > 1 && !(p->band[i].ok && p->use_band[i])) p->launches += (int)p->tc[i].size() - 1;
  for (int i = 0; i < no; ++i) p->launches -= p->skip[i];
  if (trace)
    fprintf(stderr, "plan %dx%dx%d: shapes %.2f ms, allocations %.2f ms (arena %.2f GB), kernel plans / weights / timing %.2f ms, "
            "first pass %.2f ms\n", n, h, w, t_shapes, t_alloc - t_shapes, p->arena_bytes / 1e9, t_kernels - t_alloc,
            ms_since(t_begin) - t_kernels);
  // CUDA graphs are captured lazily, on the second pass of a plan and parity (enqueue_forward): pages scaled to a fixed
  // height differ in width from scan to scan, and a shape that is met once should not pay for two captures and
  // instantiations (they were half of the ~15 ms a new shape costs)
  *out = p.get();
  e->plans.push_back(std::move(p));
  return ARU_OK;
}

int get_plan(aru_engine* e, int n, int h, int w, Plan** out) {
  for (auto& p : e->plans)
    if (p->n == n && p->h == h && p->w == w) {
      p->last_use = ++e->tick;
      *out = p.get();
      return ARU_OK;
    }
  // at most 8 plans and ~120 GB of arenas per engine; the least recently used ones go first (never the new one)
  auto evict_lru = [&](const Plan* keep) {
    size_t lru = e->plans.size();
    for (size_t i = 0; i < e->plans.size(); ++i)
      if (e->plans[i].get() != keep && (lru == e->plans.size() || e->plans[i]->last_use < e->plans[lru]->last_use)) lru = i;
    if (lru == e->plans.size()) return false;
    cudaDeviceSynchronize();
    if (e->cur == e->plans[lru].get()) e->cur = nullptr;
    free_plan(e->plans[lru].get());
    e->plans.erase(e->plans.begin() + lru);
    return true;
  };
  if (e->plans.size() >= 8) evict_lru(nullptr);
  int rc = build_plan(e, n, h, w, out);
  if (rc == ARU_ENOMEM && evict_lru(nullptr)) {   // make room and try once more
    while (e->plans.size() > 0 && evict_lru(nullptr)) {}
    rc = build_plan(e, n, h, w, out);
  }
  if (rc) return rc;
  (*out)->last_use = ++e->tick;
  for (;;) {
    size_t total = 0;
    for (auto& p : e->plans) total += p->arena_bytes;
    if (total <= ((size_t)120 << 30) || !evict_lru(*out)) break;
  }
  return ARU_OK;
}

int capture_graph(aru_engine* e, Plan* p, int par, cudaStream_t st) {
  cudaGraph_t g = nullptr;
  CU(e, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = run_all(e, p, par, st);
  cudaError_t err = cudaStreamEndCapture(st, &g);
  if (rc || err != cudaSuccess) {
    if (g) cudaGraphDestroy(g);
    return rc ? rc : fail(e, ARU_ECUDA, "graph capture failed: %s", cudaGetErrorString(err));
  }
  err = cudaGraphInstantiate(&p->graph[par], g, 0);
  cudaGraphDestroy(g);
  if (err != cudaSuccess) return fail(e, ARU_ECUDA, "graph instantiation failed: %s", cudaGetErrorString(err));
  return ARU_OK;
}

int enqueue_forward(aru_engine* e, Plan* p, int parity, cudaStream_t st) {
  if (e->use_graph) {
    if (!p->graph[parity] && p->eager_passes[parity] >= 1) {   // second pass of this plan and parity: worth a graph
      const int rc = capture_graph(e, p, parity, st);
      if (rc) return rc;
    }
    if (p->graph[parity]) {
      CU(e, cudaGraphLaunch(p->graph[parity], st));
      return ARU_OK;
    }
    ++p->eager_passes[parity];
  }
  return run_all(e, p, parity, st);
}

int ensure_quant(aru_engine* e, Plan* p, bool u8, bool mask) {
  for (int i = 0; i < 2; ++i) {
    if (u8 && !p->u8_dev[i]) CU(e, cudaMalloc((void**)&p->u8_dev[i], (size_t)p->n * p->h * p->w * e->n_class));
    if (mask && !p->mask_dev[i]) CU(e, cudaMalloc((void**)&p->mask_dev[i], (size_t)p->n * p->h * p->w));
  }
  return ARU_OK;
}

int ensure_pages(aru_engine* e, Plan* p, int channels) {
  if (p->page_dev[0] && p->page_dev_channels >= channels) return ARU_OK;
  for (int i = 0; i < 2; ++i) {
    if (p->page_dev[i]) { CU(e, cudaStreamSynchronize(e->s_h2d)); CU(e, cudaStreamSynchronize(e->s_comp)); cudaFree(p->page_dev[i]); p->page_dev[i] = nullptr; }
    CU(e, cudaMalloc((void**)&p->page_dev[i], (size_t)p->n * p->h * p->w * channels));
  }
  p->page_dev_channels = channels;
  return ARU_OK;
}

int ensure_post(aru_engine* e, Plan* p) {
  for (int i = 0; i < 2; ++i) {
    if (!p->hor_dev[i]) CU(e, cudaMalloc((void**)&p->hor_dev[i], (size_t)p->n * p->h * p->w));
    if (!p->ver_dev[i]) CU(e, cudaMalloc((void**)&p->ver_dev[i], (size_t)p->n * p->h * p->w));
  }
  if (!p->post_scratch) CU(e, cudaMalloc(&p->post_scratch, separator_post_scratch_bytes(p->n, p->h, p->w)));
  return ARU_OK;
}

// int(net_output.size * (1 / net_output.size * 100)) in IEEE doubles, as Python evaluates it
// (separator_net_post_processor.py:37, region_net_post_processor_base.py:244): 100, sometimes 99
int cc_min_size(long long page_px) {
  const double thr = 1.0 / (double)page_px * 100.0;
  return (int)((double)page_px * thr);
}

// apply_threshold on the uint8 map (helper.py:75-78): u8 > threshold * 255 (double)  <=>  u8 >= floor(t) + 1
int mask_cut(double thr) {
  const double t = thr * 255.0;
  if (t < 0.0) return 0;
  if (t >= 255.0) return 256;
  return (int)std::floor(t) + 1;
}

int pick_micro_batch(const aru_engine* e, int n, int h, int w) {
  if (e->micro_batch > 0) return std::min(n, e->micro_batch);
  const long long px = (long long)h * w;
  // ~80 Mpx per pass: every launch carries ~10 us of fixed cost (prologue, pipeline fill, tail) and the deep levels of
  // the pyramid are small, so a pass of 32 pages of 2.5 Mpx is 7 % faster per page than one of 16 (measured); the arena
  // of such a pass is ~30 GB of the 180 GB
  long long mb = (80LL << 20) / std::max<long long>(px, 1);
  mb = std::max<long long>(1, std::min<long long>(mb, 32));
  return (int)std::min<long long>(mb, n);
}

// sizes of the three structuring elements of SeparatorNetPostProcessor.post_process (sep:71,76,85):
// int(15 * W / 1000), int(30 * H / 1500), int(10 * W / 1000) - Python float division, then truncation
void separator_kernel_sizes(int h, int w, int* k_h1, int* k_v, int* k_h2) {
  *k_h1 = (int)(15.0 * w / 1000.0);
  *k_v = (int)(30.0 * h / 1500.0);
  *k_h2 = (int)(10.0 * w / 1000.0);
}

// cv2.resize(image, None, fx=sc, fy=sc, interpolation=cv2.INTER_AREA) (scale_image, helper.py:14-25): destination size
// and the (source index, weight) tables of OpenCV's area resampling, in its own double / float arithmetic.
struct ScalePlan {
  int sh = 0, sw = 0, dh = 0, dw = 0;
  int fast = 0, ix = 1, iy = 1;
  int *x_start = nullptr, *x_si = nullptr, *y_start = nullptr, *y_si = nullptr;
  float *x_alpha = nullptr, *y_alpha = nullptr;
  int cubic = 0;   // enlarging: INTER_CUBIC tables [d][4] (indices, 11-bit fixed-point weights)
  int *cx_idx = nullptr, *cx_coef = nullptr, *cy_idx = nullptr, *cy_coef = nullptr;
};

void free_scale_plan(ScalePlan* sp) {
  cudaFree(sp->x_start); cudaFree(sp->x_si); cudaFree(sp->x_alpha);
  cudaFree(sp->y_start); cudaFree(sp->y_si); cudaFree(sp->y_alpha);
  cudaFree(sp->cx_idx); cudaFree(sp->cx_coef); cudaFree(sp->cy_idx); cudaFree(sp->cy_coef);
  *sp = ScalePlan();
}

// OpenCV's cubic tables (resize.cpp, INTER_CUBIC, 8-bit): fx = (float)((dx + 0.5) * scale - 0.5), sx = floor(fx),
// weights of taps sx - 1 .. sx + 2 with A = -0.75 in float32, stored as cvRound(w * 2048); indices clamped to the image
void cubic_table(int ssize, int dsize, double scale, std::vector<int>* idx, std::vector<int>* coef) {
  idx->resize((size_t)dsize * 4);
  coef->resize((size_t)dsize * 4);
  const float A = -0.75f;
  for (int d = 0; d < dsize; ++d) {
    float fx = (float)((d + 0.5) * scale - 0.5);
    const int sx = (int)std::floor(fx);
    fx -= (float)sx;
    float c[4];
    c[0] = ((A * (fx + 1) - 5 * A) * (fx + 1) + 8 * A) * (fx + 1) - 4 * A;
    c[1] = ((A + 2) * fx - (A + 3)) * fx * fx + 1;
    c[2] = ((A + 2) * (1 - fx) - (A + 3)) * (1 - fx) * (1 - fx) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
    for (int k = 0; k < 4; ++k) {
      (*idx)[(size_t)d * 4 + k] = std::min(std::max(sx - 1 + k, 0), ssize - 1);
      (*coef)[(size_t)d * 4 + k] = (int)std::nearbyint(c[k] * 2048.f);
    }
  }
}

int cv_round(double v) { return (int)std::nearbyint(v); }   // cvRound / saturate_cast<int>(double): half to even

void area_table(int ssize, int dsize, double scale, std::vector<int>* start, std::vector<int>* si, std::vector<float>* alpha) {
  start->assign(1, 0);
  for (int dx = 0; dx < dsize; ++dx) {
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    const double cell = std::min(scale, ssize - fsx1);
    int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) { si->push_back(sx1 - 1); alpha->push_back((float)((sx1 - fsx1) / cell)); }
    for (int sx = sx1; sx < sx2; ++sx) { si->push_back(sx); alpha->push_back((float)(1.0 / cell)); }
    if (fsx2 - sx2 > 1e-3) { si->push_back(sx2); alpha->push_back((float)(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell)); }
    start->push_back((int)si->size());
  }
}

int make_scale_plan(aru_engine* e, int sh, int sw, double sc, ScalePlan* sp) {
  if (!(sc > 0.0) || sc == 1.0) return fail(e, ARU_EINVAL, "bad scale %.6f", sc);
  sp->sh = sh; sp->sw = sw;
  sp->dw = cv_round(sw * sc);
  sp->dh = cv_round(sh * sc);
  if (sp->dw < 1 || sp->dh < 1) return fail(e, ARU_EINVAL, "scaled page would be empty");
  if (sc > 1.0) {
    // enlarging: cv2.INTER_CUBIC (helper.py:21-23); scale = 1 / fx as OpenCV computes it for fx-given calls
    std::vector<int> xi, xc, yi, yc;
    cubic_table(sw, sp->dw, 1.0 / sc, &xi, &xc);
    cubic_table(sh, sp->dh, 1.0 / sc, &yi, &yc);
    sp->cubic = 1;
    int rc;
    if ((rc = upload(e, xi, &sp->cx_idx)) || (rc = upload(e, xc, &sp->cx_coef)) || (rc = upload(e, yi, &sp->cy_idx)) ||
        (rc = upload(e, yc, &sp->cy_coef))) {
      free_scale_plan(sp);
      return rc;
    }
    return ARU_OK;
  }
  const double scale = 1.0 / sc;
  const int iscale = cv_round(scale);
  if (std::fabs(scale - iscale) < 2.220446049250313e-16) {   // OpenCV's integer-scale path
    if ((long long)sp->dw * iscale > sw || (long long)sp->dh * iscale > sh)
      return fail(e, ARU_EUNSUP, "integer-scale area resize with a partial last cell (%dx%d / %d) is not covered on the device",
                  sh, sw, iscale);
    sp->fast = 1; sp->ix = sp->iy = iscale;
    return ARU_OK;
  }
  std::vector<int> xs, xi, ys, yi;
  std::vector<float> xa, ya;
  area_table(sw, sp->dw, scale, &xs, &xi, &xa);
  area_table(sh, sp->dh, scale, &ys, &yi, &ya);
  int rc;
  if ((rc = upload(e, xs, &sp->x_start)) || (rc = upload(e, xi, &sp->x_si)) || (rc = upload(e, xa, &sp->x_alpha)) ||
      (rc = upload(e, ys, &sp->y_start)) || (rc = upload(e, yi, &sp->y_si)) || (rc = upload(e, ya, &sp->y_alpha))) {
    free_scale_plan(sp);
    return rc;
  }
  return ARU_OK;
}

// After the first pass on real data: count the stored activations that sit on the storage limit or are not finite.
int range_check(aru_engine* e, Plan* p) {
  unsigned long long* cnt = nullptr;
  const int nb = (int)p->bufs.size();
  CU(e, cudaMalloc((void**)&cnt, (size_t)nb * sizeof(unsigned long long)));
  CU(e, cudaMemsetAsync(cnt, 0, (size_t)nb * sizeof(unsigned long long), e->s_comp));
  for (int b = 0; b < nb; ++b) {
    const BufPlan& bp = p->bufs[b];
    if (!bp.sized || p->buf_dead[b] || bp.kind != KIND_PLANAR || b == e->input_buf || b == e->output_buf) continue;
    cudaError_t err = launch_range_scan(e->s_comp, reinterpret_cast<const act_t*>(p->arena + bp.offset),
                                        (long long)bp.chunks * bp.geo.plane * 8, cnt + b);
    if (err != cudaSuccess) { cudaFree(cnt); return fail(e, ARU_ECUDA, "range scan: %s", cudaGetErrorString(err)); }
  }
  std::vector<unsigned long long> h(nb);
  cudaError_t err = cudaMemcpyAsync(h.data(), cnt, (size_t)nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->s_comp);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->s_comp);
  cudaFree(cnt);
  if (err != cudaSuccess) return fail(e, ARU_ECUDA, "range scan: %s", cudaGetErrorString(err));
  unsigned long long total = 0;
  int first = -1, n_bufs = 0;
  for (int b = 0; b < nb; ++b)
    if (h[b]) { total += h[b]; ++n_bufs; if (first < 0) first = b; }
  if (total) {
    int writer = -1;
    for (int i = 0; i < (int)e->ops.size() && writer < 0; ++i)
      if (e->ops[i].out.buf == first || e->ops[i].out_pre.buf == first) writer = i;
    char msg[512];
    snprintf(msg, sizeof msg, "%llu stored activations in %d tensor(s) reached the " ARU_ACT_NAME " storage limit or are not "
             "finite (first: buffer %d written by op %d): the probability maps of this graph are not trustworthy in this "
             "build%s", total, n_bufs, first, writer,
#ifdef ARU_USE_BF16
             "");
#else
             "; fp16 stores saturate at +-65504 - use the bf16 library (ARU_B200_LIB=.../libaru_b200_bf16.so)");
#endif
    e->warning = msg;
    fprintf(stderr, "aru_b200 warning: %s\n", msg);
  }
  return ARU_OK;
}

struct HostIO {
  const ScalePlan* scale = nullptr;   // in_u8 holds unscaled pages sh x sw; they are resized on the device first
  const float* in_f32 = nullptr;   // float32 [n,h,w] pages (gray / 255) ...
  const uint8_t* in_u8 = nullptr;  // ... or uint8 [n,h,w,channels] pages (gray, or BGR as cv2.imread returns them)
  int channels = 1;
  float* out = nullptr;            // float32 [n,h,w,c]
  uint8_t* out_u8 = nullptr;       // uint8 [n,h,w,c]
  uint8_t* out_mask = nullptr;     // uint8 [n,h,w]
  uint8_t* out_h = nullptr;        // uint8 [n,h,w] horizontal separators (post_process)
  uint8_t* out_v = nullptr;        // uint8 [n,h,w] vertical separators
  float thr = 0.f;                 // float threshold (aru_forward) ...
  int cut = -1;                    // ... or the integer cut of the double threshold (>= 0)
  const int* boxes_dev = nullptr;  // device [n_boxes][5] = page, y0, y1, x0, x1 (aru_heading_pages)
  int n_boxes = 0;
  unsigned long long* sums_dev = nullptr;
};

// Host-buffer forward: micro-batches, double-buffered device staging, copies on their own streams.
int forward_host_impl(aru_engine* e, const HostIO& io, int n, int h, int w);
int forward_host(aru_engine* e, const HostIO& io, int n, int h, int w) {
  const int rc = forward_host_impl(e, io, n, h, w);
  if (rc != ARU_OK) {
    // A failure in the middle of the schedule leaves copies of earlier micro-batches in flight: drain the three streams
    // before the caller may recycle its (pinned) buffers, and forget the pending flags.
    const std::string msg = e->error;
    cudaStreamSynchronize(e->s_h2d);
    cudaStreamSynchronize(e->s_comp);
    cudaStreamSynchronize(e->s_d2h);
    cudaGetLastError();
    for (auto& pl : e->plans) pl->d2h_pending[0] = pl->d2h_pending[1] = false;
    e->error = msg;
  }
  return rc;
}
int forward_host_impl(aru_engine* e, const HostIO& io, int n, int h, int w) {
  CU(e, cudaSetDevice(e->device));
  const int mb = pick_micro_batch(e, n, h, w);
  const size_t page_px = (size_t)h * w;
  const int C = e->n_class;
  // channels of the uint8 map that are produced / copied out (ARU_OPT_U8_CHANNELS; the box sums read channel 0 only)
  const int Cu = (e->u8_channels > 0 && e->u8_channels < C) ? e->u8_channels : C;
  const bool want_post = io.out_h != nullptr;
  const bool want_mask = io.out_mask != nullptr || want_post;
  int k1 = 0, kv = 0, k2 = 0;
  separator_kernel_sizes(h, w, &k1, &kv, &k2);
  int slot = 0;
  // Micro-batch schedule.  The host->device copy of the first pass and the device->host copy of the last one cannot
  // overlap with compute, and a pass hides at most the copies that fit in its own duration: the first pass is a
  // quarter batch, the tail halves down (mb/2, mb/4, mb/8) so that the exposed last copy is small and every earlier
  // copy-out runs under the next pass: [mb/4, mb, ..., (rest), mb/2, mb/4, mb/8].  A three-stream pipeline model with
  // the measured rates (0.18 / 0.55 / 0.36 ms per 2.5 Mpx page for in / net / out, 1.2 ms per pass) puts 64 pages at
  // 44.1 ms against 46.2 ms for [8, 32, 16, 8] (measured: 1468 against 1411 pages/s).  The ramp is for the float32 maps
  // (8 B/px out); with uint8 / mask outputs only (1-3 B/px) the copies are short and the extra small passes cost more
  // than they hide, so those calls keep [mb/4, mb, ..., (rest), mb/4].
  std::vector<int> sched;
  if (e->async_calls && io.n_boxes == 0) {
    // calls in flight overlap each other: the exposed first copy-in / last copy-out run under the neighbouring calls,
    // so plain full passes (no small head / ramped tail, which cost fixed launch time) are the fastest schedule
    // When float32 maps go back (8 B/px down, 4 B/px up: the copies of a pass take almost as long as its kernels) one
    // pass more than necessary, all of about equal size, keeps the copy streams ahead: 64 pages as [22, 22, 20] instead
    // of [32, 32] measured 1 930 - 1 976 against 1 763 - 1 826 pages/s (A/B in one call; uint8 / mask calls do not care).
    int per = mb;
    if (io.out != nullptr && e->micro_batch == 0 && n >= 16) per = cdiv(n, cdiv(n, mb) + 1);
    for (int left = n; left > 0; left -= per) sched.push_back(std::min(per, left));
  } else if (n > mb && mb >= 8 && io.out != nullptr) {
    const int head = mb / 4;
    std::vector<int> tail = {mb / 2, mb / 4, mb / 8};
    int tail_sum = tail[0] + tail[1] + tail[2];
    while (!tail.empty() && head + tail_sum > n) { tail_sum -= tail.front(); tail.erase(tail.begin()); }
    sched.push_back(head);
    int left = n - head - tail_sum;
    while (left > 0) { sched.push_back(std::min(mb, left)); left -= sched.back(); }
    for (int t : tail) sched.push_back(t);
  } else if (n > mb && mb >= 4) {
    const int edge = std::max(1, mb / 4);
    sched.push_back(edge);
    int left = n - 2 * edge;
    while (left > 0) { sched.push_back(std::min(mb, left)); left -= sched.back(); }
    sched.push_back(edge);
  } else {
    for (int left = n; left > 0; left -= mb) sched.push_back(std::min(mb, left));
  }
  int i0 = 0;
  for (size_t si = 0; si < sched.size(); i0 += sched[si], ++si, ++slot) {
    const int cnt = sched[si];
    Plan* p = nullptr;
    int rc = get_plan(e, cnt, h, w, &p);
    if (rc) return rc;
    e->cur = p;
    const bool want_u8 = io.out_u8 != nullptr || io.n_boxes > 0;
    rc = ensure_quant(e, p, want_u8, want_mask);
    if (rc) return rc;
    if (io.in_u8 && (rc = ensure_pages(e, p, io.channels))) return rc;
    const size_t src_px = io.scale ? (size_t)io.scale->sh * io.scale->sw : 0;
    if (io.scale && p->src_dev_bytes < (size_t)p->n * src_px * io.channels) {
      CU(e, cudaStreamSynchronize(e->s_h2d));
      CU(e, cudaStreamSynchronize(e->s_comp));
      for (int i = 0; i < 2; ++i) {
        if (p->src_dev[i]) { cudaFree(p->src_dev[i]); p->src_dev[i] = nullptr; }
        CU(e, cudaMalloc((void**)&p->src_dev[i], (size_t)p->n * src_px * io.channels));
      }
      p->src_dev_bytes = (size_t)p->n * src_px * io.channels;
    }
    if (want_post && (rc = ensure_post(e, p))) return rc;
    const int par = slot & 1;
    // buffers of this parity are free once the device->host copies of their previous use are done
    if (p->d2h_pending[par]) {
      CU(e, cudaStreamWaitEvent(e->s_h2d, p->ev_d2h[par], 0));
      CU(e, cudaStreamWaitEvent(e->s_comp, p->ev_d2h[par], 0));
    }
    if (io.in_u8 && io.scale)
      CU(e, cudaMemcpyAsync(p->src_dev[par], io.in_u8 + (size_t)i0 * src_px * io.channels,
                            (size_t)cnt * src_px * io.channels, cudaMemcpyHostToDevice, e->s_h2d));
    else if (io.in_u8)
      CU(e, cudaMemcpyAsync(p->page_dev[par], io.in_u8 + (size_t)i0 * page_px * io.channels,
                            (size_t)cnt * page_px * io.channels, cudaMemcpyHostToDevice, e->s_h2d));
    else
      CU(e, cudaMemcpyAsync(p->in_dev[par], io.in_f32 + (size_t)i0 * page_px, (size_t)cnt * page_px * sizeof(float),
                            cudaMemcpyHostToDevice, e->s_h2d));
    CU(e, cudaEventRecord(p->ev_h2d[par], e->s_h2d));
    CU(e, cudaStreamWaitEvent(e->s_comp, p->ev_h2d[par], 0));
    if (io.scale) {
      const ScalePlan& sp = *io.scale;
      cudaError_t err = sp.cubic
          ? launch_resize_cubic(e->s_comp, p->src_dev[par], io.channels, cnt, sp.sh, sp.sw, p->page_dev[par], h, w, sp.cx_idx,
                                sp.cx_coef, sp.cy_idx, sp.cy_coef)
          : launch_resize_area(e->s_comp, p->src_dev[par], io.channels, cnt, sp.sh, sp.sw, p->page_dev[par], h, w,
                               sp.x_start, sp.x_si, sp.x_alpha, sp.y_start, sp.y_si, sp.y_alpha, sp.fast, sp.ix, sp.iy);
      if (err != cudaSuccess) return fail(e, ARU_ECUDA, "resize_area launch: %s", cudaGetErrorString(err));
    }
    if (io.in_u8) {
      cudaError_t err = launch_pages_to_input(e->s_comp, p->page_dev[par], io.channels, (long long)cnt * page_px,
                                              p->in_dev[par], nullptr);
      if (err != cudaSuccess) return fail(e, ARU_ECUDA, "pages_to_input launch: %s", cudaGetErrorString(err));
    }
    rc = enqueue_forward(e, p, par, e->s_comp);
    if (rc) return rc;
    if (!e->range_checked) {
      e->range_checked = 1;
      const char* rcv = getenv("ARU_RANGE_CHECK");
      if (!(rcv && rcv[0] == '0') && (rc = range_check(e, p))) return rc;
    }
    if (want_u8 || want_mask) {
      cudaError_t err = launch_quantize(e->s_comp, p->out_dev[par], want_u8 ? p->u8_dev[par] : nullptr,
                                        want_mask ? p->mask_dev[par] : nullptr, (long long)cnt * page_px, C, io.thr, io.cut, Cu);
      if (err != cudaSuccess) return fail(e, ARU_ECUDA, "quantize launch: %s", cudaGetErrorString(err));
    }
    if (io.n_boxes > 0) {
      cudaError_t err = launch_box_sums(e->s_comp, p->u8_dev[par], cnt, h, w, Cu, i0, io.boxes_dev, io.n_boxes, io.sums_dev);
      if (err != cudaSuccess) return fail(e, ARU_ECUDA, "box_sums launch: %s", cudaGetErrorString(err));
    }
    if (want_post) {
      cudaError_t err = launch_separator_post(e->s_comp, p->mask_dev[par], cnt, h, w, cc_min_size((long long)page_px), k1,
                                              kv, k2, p->post_scratch, p->hor_dev[par], p->ver_dev[par]);
      if (err != cudaSuccess) return fail(e, ARU_ECUDA, "separator_post launch: %s", cudaGetErrorString(err));
    }
    CU(e, cudaEventRecord(p->ev_comp[par], e->s_comp));
    CU(e, cudaStreamWaitEvent(e->s_d2h, p->ev_comp[par], 0));
    if (io.out)
      CU(e, cudaMemcpyAsync(io.out + (size_t)i0 * page_px * C, p->out_dev[par], (size_t)cnt * page_px * C * sizeof(float),
                            cudaMemcpyDeviceToHost, e->s_d2h));
    if (io.out_u8)
      CU(e, cudaMemcpyAsync(io.out_u8 + (size_t)i0 * page_px * Cu, p->u8_dev[par], (size_t)cnt * page_px * Cu,
                            cudaMemcpyDeviceToHost, e->s_d2h));
    if (io.out_mask)
      CU(e, cudaMemcpyAsync(io.out_mask + (size_t)i0 * page_px, p->mask_dev[par], (size_t)cnt * page_px,
                            cudaMemcpyDeviceToHost, e->s_d2h));
    if (want_post) {
      CU(e, cudaMemcpyAsync(io.out_h + (size_t)i0 * page_px, p->hor_dev[par], (size_t)cnt * page_px,
                            cudaMemcpyDeviceToHost, e->s_d2h));
      CU(e, cudaMemcpyAsync(io.out_v + (size_t)i0 * page_px, p->ver_dev[par], (size_t)cnt * page_px,
                            cudaMemcpyDeviceToHost, e->s_d2h));
    }
    CU(e, cudaEventRecord(p->ev_d2h[par], e->s_d2h));
    p->d2h_pending[par] = true;
  }
  if (e->async_calls && io.n_boxes == 0) {
    // ARU_OPT_ASYNC: return once everything is enqueued; the copy stream's last event is the call's ticket (every
    // micro-batch's copy-out waits for its compute, so the event covers the whole call).  The staging buffers of the
    // plans stay guarded by their own events, so the next call's first copy-in overlaps this call's tail.
    const uint64_t t = ++e->last_ticket;
    cudaEvent_t& ev = e->tickets[t % 8];
    if (!ev) CU(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(e, cudaEventRecord(ev, e->s_d2h));
    return ARU_OK;
  }
  CU(e, cudaStreamSynchronize(e->s_d2h));
  CU(e, cudaStreamSynchronize(e->s_comp));
  for (auto& pl : e->plans) pl->d2h_pending[0] = pl->d2h_pending[1] = false;   // everything has completed
  return check_err_flag(e);
}

}  // namespace

// ---- C ABI -----------------------------------------------------------------------------------------
extern "C" {

int aru_abi_version(void) { return ARU_ABI_VERSION; }

int aru_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* aru_last_error(const aru_engine* e) { return e ? e->error.c_str() : g_error.c_str(); }

const char* aru_last_warning(const aru_engine* e) { return e ? e->warning.c_str() : ""; }

// float64 page -> float32 staging buffer on a few host threads (numpy's single-threaded astype of a 2.5 Mpx page costs
// more than the page's whole forward pass)
int aru_f64_to_f32(const double* src, float* dst, long long count, int threads) {
  if (count < 0 || (count > 0 && (!src || !dst))) return ARU_EINVAL;
  cpu_set_t set;
  int avail = 1;
  if (sched_getaffinity(0, sizeof set, &set) == 0) avail = std::max(1, CPU_COUNT(&set));
  // one process per GPU: the ranks of a box share its cores (torchrun exports LOCAL_WORLD_SIZE)
  int ranks = 1;
  if (const char* lw = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(lw));
  int nt = threads > 0 ? threads : std::max(1, std::min(avail / ranks, 16));
  nt = (int)std::max<long long>(1, std::min<long long>(nt, count / 65536));
  // Streaming (non-temporal) stores where the target is 16-byte aligned: the staging buffer is read next by the copy
  // engine, not by the CPU, and lines left dirty in the caches slow that read down.
  const bool stream_ok = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  auto work = [=](int t) {
    long long b = (count * t / nt) & ~3LL, e2 = t + 1 == nt ? count : (count * (t + 1) / nt) & ~3LL;
    long long i = b;
#if defined(__x86_64__) && defined(__SSE2__)
    if (stream_ok) {
      for (; i + 4 <= e2; i += 4) {
        const __m128 lo = _mm_cvtpd_ps(_mm_loadu_pd(src + i)), hi = _mm_cvtpd_ps(_mm_loadu_pd(src + i + 2));
        _mm_stream_ps(dst + i, _mm_movelh_ps(lo, hi));
      }
      _mm_sfence();
    }
#endif
    for (; i < e2; ++i) dst[i] = (float)src[i];
  };
  if (nt == 1) { work(0); return ARU_OK; }
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  try {
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
  } catch (...) {   // no more threads to be had: the caller's thread does the rest
    const int started = (int)pool.size();
    for (int t = started + 1; t < nt; ++t) work(t);
  }
  work(0);
  for (auto& th : pool) th.join();
  return ARU_OK;
}

// One process per GPU: the host buffers a rank page-locks must live on the NUMA node its GPU hangs off, or the
// host<->device copies of the ranks on the other socket cross the inter-socket link and the 8-GPU end-to-end rate
// collapses (profiles/r02*_host_bw_probe*.json).  Linux: the PCI device's local_cpulist / numa_node in sysfs.
int aru_bind_host_to_device(int device, int* numa_node) {
  if (numa_node) *numa_node = -1;
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return ARU_ECUDA; }
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  auto read_line = [&](const char* leaf, char* buf, size_t n) -> bool {
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/%s", bus, leaf);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    const bool ok = fgets(buf, (int)n, f) != nullptr;
    fclose(f);
    return ok;
  };
  char line[4096];
  int node = -1;
  if (read_line("numa_node", line, sizeof line)) node = atoi(line);
  if (numa_node) *numa_node = node;
  if (!read_line("local_cpulist", line, sizeof line)) return ARU_EUNSUP;
  cpu_set_t set;
  CPU_ZERO(&set);
  int n_cpus = 0;
  for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int a = 0, b = 0;
    const int k = sscanf(tok, "%d-%d", &a, &b);
    if (k < 1) continue;
    if (k == 1) b = a;
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &set); ++n_cpus; }
  }
  if (n_cpus == 0) return ARU_EUNSUP;
  // keep the CPUs the process was allowed to use (containers, taskset): intersect, and give up if nothing is left
  cpu_set_t cur, both;
  if (sched_getaffinity(0, sizeof cur, &cur) == 0) {
    CPU_AND(&both, &cur, &set);
    if (CPU_COUNT(&both) == 0) return ARU_EUNSUP;
    set = both;
  }
  if (sched_setaffinity(0, sizeof set, &set) != 0) return ARU_EUNSUP;
#ifdef SYS_set_mempolicy
  if (node >= 0 && node < 1024) {   // MPOL_PREFERRED = 1: new pages (and what cudaHostAlloc pins) come from the local node
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1UL << (node % (8 * sizeof(unsigned long)));
    syscall(SYS_set_mempolicy, 1, mask, (unsigned long)(8 * sizeof mask));
  }
#endif
  return ARU_OK;
}

int aru_create(const aru_graph_desc* g, int device, aru_engine** out) {
  if (!g || !out) return fail(nullptr, ARU_EINVAL, "aru_create: null argument");
  *out = nullptr;
  if (g->magic != ARU_PROGRAM_MAGIC || g->abi_version != ARU_ABI_VERSION)
    return fail(nullptr, ARU_EINVAL, "aru_create: bad magic / ABI version (got %08x / %u)", g->magic, g->abi_version);
  if (g->n_buffers <= 0 || g->n_ops <= 0 || !g->buffers || !g->ops || (g->n_weights > 0 && !g->weights))
    return fail(nullptr, ARU_EINVAL, "aru_create: empty program");
  int ndev = aru_device_count();
  if (ndev <= 0) return fail(nullptr, ARU_ENODEV, "no CUDA device: the B200 engine has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(nullptr, ARU_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
  std::unique_ptr<aru_engine> e(new aru_engine);
  e->device = device;
  {
    // ARU_NUMA_BIND=0 leaves the calling thread's CPU affinity / memory policy alone
    const char* nb = getenv("ARU_NUMA_BIND");
    if (!(nb && nb[0] == '0')) aru_bind_host_to_device(device, nullptr);
  }
  cudaError_t err = cudaSetDevice(device);
  if (err != cudaSuccess) return fail(nullptr, ARU_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(err));
  cudaDeviceProp prop;
  err = cudaGetDeviceProperties(&prop, device);
  if (err != cudaSuccess) return fail(nullptr, ARU_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(err));
  if (prop.major != 10)
    return fail(nullptr, ARU_EUNSUP, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
  e->num_sms = prop.multiProcessorCount;
  e->max_smem = prop.sharedMemPerBlockOptin;
  e->buffers.assign(g->buffers, g->buffers + g->n_buffers);
  e->ops.assign(g->ops, g->ops + g->n_ops);
  e->weights.assign(g->weights, g->weights + g->n_weights);
  e->kind.resize(g->n_buffers);
  for (int i = 0; i < g->n_buffers; ++i) {
    const aru_buffer& b = e->buffers[i];
    if (b.channels <= 0) return fail(nullptr, ARU_EINVAL, "buffer %d has %d channels", i, b.channels);
    if (b.role == 1) {
      if (e->input_buf >= 0 || b.channels != 1) return fail(nullptr, ARU_EUNSUP, "exactly one 1-channel input is supported");
      e->input_buf = i;
    }
    if (b.role == 2) {
      if (e->output_buf >= 0) return fail(nullptr, ARU_EUNSUP, "exactly one output is supported");
      e->output_buf = i;
      e->n_class = b.channels;
    }
    e->kind[i] = b.role == 2 ? KIND_OUT : (b.channels == 1 ? KIND_F32 : KIND_PLANAR);
  }
  if (e->input_buf < 0 || e->output_buf < 0) return fail(nullptr, ARU_EINVAL, "program lacks an input or output buffer");
  // static validation of the ops
  for (int i = 0; i < g->n_ops; ++i) {
    const aru_op& op = e->ops[i];
    auto bad = [&](const char* what) { return fail(nullptr, ARU_EINVAL, "op %d: %s", i, what); };
    if (op.kind == ARU_OP_COMBINE) {
      if (op.n_scales < 1 || op.n_scales > ARU_MAX_SCALES) return bad("bad number of attention scales");
      if (!view_ok(e.get(), op.out) || e->kind[op.out.buf] != KIND_PLANAR || op.out.ch_off % 8) return bad("bad output view");
      if (op.like_buf < 0 || op.like_buf >= g->n_buffers) return bad("bad like buffer");
      for (int k = 0; k < op.n_scales; ++k) {
        if (!view_ok(e.get(), op.att[k]) || e->kind[op.att[k].buf] != KIND_F32) return bad("attention map must be a 1-channel plane");
        if (!view_ok(e.get(), op.det[k]) || e->kind[op.det[k].buf] != KIND_PLANAR || op.det[k].ch_off % 8)
          return bad("bad detection map view");
        if (op.up_att[k] < 1 || op.up_det[k] < 1) return bad("bad upsample factor");
        if (op.up_det[k] == 1 && op.det[k].ch != op.out.ch) return bad("full-resolution detection map channel mismatch");
      }
      continue;
    }
    if (!view_ok(e.get(), op.in) || !view_ok(e.get(), op.out)) return bad("bad input / output view");
    const int ik = e->kind[op.in.buf], ok = e->kind[op.out.buf];
    if (ik == KIND_PLANAR && op.in.ch_off % 8) return bad("input channel offset must be a multiple of 8");
    if (ok == KIND_PLANAR && op.out.ch_off % 8) return bad("output channel offset must be a multiple of 8");
    if (ik == KIND_OUT) return bad("the network output cannot be read by an op");
    switch (op.kind) {
      case ARU_OP_CONV:
        if (op.ksize != 3 && op.ksize != 4) return bad("only 3x3 and 4x4 convolutions are supported");
        if (ik == KIND_F32) {
          if (ok != KIND_PLANAR || op.out.ch > 16) return bad("1-channel input conv must produce <= 16 planar channels");
          if (op.act != ARU_ACT_NONE && op.act != ARU_ACT_RELU) return bad("unsupported activation on a stem conv");
          if (op.res.buf >= 0) return bad("residual on a stem conv");
        } else if (ok != KIND_PLANAR) {
          if (op.out.ch > 8) return bad("float-output conv supports C_out <= 8");
          if (op.res.buf >= 0 || op.out_pre.buf >= 0) return bad("residual / pre-activation export on a float-output conv");
          if (op.out.ch_off != 0 || op.out.ch != e->buffers[op.out.buf].channels) return bad("float outputs cannot be sliced");
        } else {
          if (op.act != ARU_ACT_NONE && op.act != ARU_ACT_RELU) return bad("unsupported activation");
        }
        if (op.out_pre.buf >= 0 && (!view_ok(e.get(), op.out_pre) || e->kind[op.out_pre.buf] != KIND_PLANAR ||
                                    op.out_pre.ch != op.out.ch || op.out_pre.ch_off % 8))
          return bad("bad pre-activation view");
        if (op.res.buf >= 0 && (!view_ok(e.get(), op.res) || e->kind[op.res.buf] != KIND_PLANAR || op.res.ch != op.out.ch ||
                                op.res.ch_off % 8))
          return bad("bad residual view");
        break;
      case ARU_OP_DECONV:
        if (op.ksize != 3 || op.stride != 2) return bad("only 3x3 stride-2 transposed convolutions are supported");
        if (ik != KIND_PLANAR || ok != KIND_PLANAR) return bad("transposed conv needs planar tensors");
        if (op.act != ARU_ACT_NONE && op.act != ARU_ACT_RELU) return bad("unsupported activation");
        if (op.like_buf < 0 || op.like_buf >= g->n_buffers) return bad("bad like buffer");
        break;
      case ARU_OP_MAXPOOL:
      case ARU_OP_AVGPOOL:
        if (ik != ok && !(ik == KIND_F32 && ok == KIND_F32)) return bad("pool input / output kinds differ");
        if (op.in.ch != op.out.ch) return bad("pool channel mismatch");
        break;
      case ARU_OP_UPSUM:
        if (op.stride < 1 || op.like_buf < 0 || op.like_buf >= g->n_buffers) return bad("bad upsample");
        if (ik != ok) return bad("upsample input / output kinds differ");
        break;
      case ARU_OP_COPY:
        if (ik != KIND_PLANAR || ok != KIND_PLANAR || op.in.ch != op.out.ch) return bad("bad copy");
        break;
      default:
        return fail(nullptr, ARU_EUNSUP, "op %d: unknown kind %d", i, op.kind);
    }
  }
  err = cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking);
  if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking);
  if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking);
  {
    int prio_lo = 0, prio_hi = 0;   // branch streams yield to the full-resolution chain on the caller's stream
    if (err == cudaSuccess) err = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    for (int a2 = 0; a2 < aru_engine::N_AUX && err == cudaSuccess; ++a2)
      err = cudaStreamCreateWithPriority(&e->s_aux[a2], cudaStreamNonBlocking, prio_lo);
  }
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->err_flag, sizeof(int));
  if (err == cudaSuccess) err = cudaMemset(e->err_flag, 0, sizeof(int));
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->zero_bias, 256 * sizeof(float));
  if (err == cudaSuccess) err = cudaMemset(e->zero_bias, 0, 256 * sizeof(float));
  if (err != cudaSuccess) return fail(nullptr, ARU_ECUDA, "stream / flag setup: %s", cudaGetErrorString(err));
  e->opw.resize(g->n_ops);
  for (int i = 0; i < g->n_ops; ++i) {
    int rc = pack_op_weights(e.get(), i);
    if (rc) {
      g_error = e->error;
      aru_destroy(e.release());
      return rc;
    }
  }
  *out = e.release();
  return ARU_OK;
}

void aru_destroy(aru_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& p : e->plans) free_plan(p.get());
  for (auto& w : e->opw) {
    if (w.w16) cudaFree(w.w16);
    for (auto& im : w.tc_images) cudaFree(im.w);
    if (w.band_w) cudaFree(w.band_w);
    if (w.blk_w) cudaFree(w.blk_w);
    if (w.blk_bias) cudaFree(w.blk_bias);
    if (w.head_w) cudaFree(w.head_w);
    if (w.w32) cudaFree(w.w32);
    if (w.bias) cudaFree(w.bias);
  }
  for (auto& ev : e->tickets)
    if (ev) cudaEventDestroy(ev);
  if (e->err_flag) cudaFree(e->err_flag);
  if (e->zero_bias) cudaFree(e->zero_bias);
  if (e->s_comp) cudaStreamDestroy(e->s_comp);
  if (e->s_h2d) cudaStreamDestroy(e->s_h2d);
  if (e->s_d2h) cudaStreamDestroy(e->s_d2h);
  for (cudaStream_t q : e->s_aux) if (q) cudaStreamDestroy(q);
  delete e;
}

int aru_set_option(aru_engine* e, int option, int64_t value) {
  if (!e) return ARU_EINVAL;
  switch (option) {
    case ARU_OPT_CONV_PATH:
      if (value < 0 || value > 3) return fail(e, ARU_EINVAL, "ARU_OPT_CONV_PATH must be 0..3");
      if (e->conv_path != (int)value) {  // plans bake the kernel choice in
        cudaSetDevice(e->device);
        cudaDeviceSynchronize();
        e->plans.clear();
        e->cur = nullptr;
      }
      e->conv_path = (int)value;
      return ARU_OK;
    case ARU_OPT_USE_GRAPH:
      e->use_graph = value ? 1 : 0;
      return ARU_OK;
    case ARU_OPT_KEEP_ALL:
      if (e->keep_all != (value ? 1 : 0)) {
        cudaSetDevice(e->device);
        cudaDeviceSynchronize();
        e->plans.clear();
        e->cur = nullptr;
      }
      e->keep_all = value ? 1 : 0;
      return ARU_OK;
    case ARU_OPT_FUSE_PAIRS:
      if (e->fuse_pairs != (value ? 1 : 0)) {
        cudaSetDevice(e->device);
        cudaDeviceSynchronize();
        e->plans.clear();
        e->cur = nullptr;
      }
      e->fuse_pairs = value ? 1 : 0;
      return ARU_OK;
    case ARU_OPT_FUSE_BLOCKS:
      if (e->fuse_blocks != (value ? 1 : 0)) {
        cudaSetDevice(e->device);
        cudaDeviceSynchronize();
        e->plans.clear();
        e->cur = nullptr;
      }
      e->fuse_blocks = value ? 1 : 0;
      return ARU_OK;
    case ARU_OPT_BRANCH_STREAMS:
      if (e->branch_streams != (value ? 1 : 0)) {
        cudaSetDevice(e->device);
        cudaDeviceSynchronize();
        e->plans.clear();
        e->cur = nullptr;
      }
      e->branch_streams = value ? 1 : 0;
      return ARU_OK;
    case ARU_OPT_ASYNC:   // calls already in flight stay in flight; any later synchronous call or aru_sync completes them
      e->async_calls = value ? 1 : 0;
      return ARU_OK;
    case ARU_OPT_U8_CHANNELS:
      if (value < 0 || value > 64) return fail(e, ARU_EINVAL, "bad number of uint8 channels");
      e->u8_channels = (int)value;
      return ARU_OK;
    case ARU_OPT_MICRO_BATCH:
      if (value < 0 || value > 4096) return fail(e, ARU_EINVAL, "bad micro batch");
      e->micro_batch = (int)value;
      return ARU_OK;
    default:
      return fail(e, ARU_EINVAL, "unknown option %d", option);
  }
}

int aru_num_classes(const aru_engine* e) { return e ? e->n_class : 0; }

int aru_plan(aru_engine* e, int n, int h, int w) {
  if (!e || n <= 0 || h <= 0 || w <= 0) return e ? fail(e, ARU_EINVAL, "aru_plan: bad shape") : ARU_EINVAL;
  CU(e, cudaSetDevice(e->device));
  Plan* p = nullptr;
  int rc = get_plan(e, n, h, w, &p);
  if (rc) return rc;
  e->cur = p;
  return ARU_OK;
}

int aru_forward(aru_engine* e, const float* in, int n, int h, int w, float* out, uint8_t* out_u8, uint8_t* out_mask,
                float thr) {
  if (!e || !in || n <= 0 || h <= 0 || w <= 0) return e ? fail(e, ARU_EINVAL, "aru_forward: bad argument") : ARU_EINVAL;
  HostIO io;
  io.in_f32 = in;
  io.out = out;
  io.out_u8 = out_u8;
  io.out_mask = out_mask;
  io.thr = thr;
  return forward_host(e, io, n, h, w);
}

int aru_separator_pages(aru_engine* e, const uint8_t* pages, int channels, int n, int h, int w, double thr, float* out,
                        uint8_t* out_u8, uint8_t* out_mask, uint8_t* horizontal, uint8_t* vertical) {
  if (!e || !pages || n <= 0 || h <= 0 || w <= 0 || (channels != 1 && channels != 3))
    return e ? fail(e, ARU_EINVAL, "aru_separator_pages: bad argument") : ARU_EINVAL;
  if ((horizontal == nullptr) != (vertical == nullptr))
    return fail(e, ARU_EINVAL, "aru_separator_pages: horizontal and vertical go together");
  if (horizontal) {
    int k1, kv, k2;
    separator_kernel_sizes(h, w, &k1, &kv, &k2);
    if (k1 < 1 || kv < 1 || k2 < 1)
      return fail(e, ARU_EINVAL, "aru_separator_pages: page %dx%d too small (structuring element %dx1 / 1x%d / %dx1; "
                  "OpenCV rejects an empty element in the reference too)", h, w, k1, kv, k2);
  }
  HostIO io;
  io.in_u8 = pages;
  io.channels = channels;
  io.out = out;
  io.out_u8 = out_u8;
  io.out_mask = out_mask;
  io.out_h = horizontal;
  io.out_v = vertical;
  io.cut = mask_cut(thr);
  return forward_host(e, io, n, h, w);
}

int aru_scaled_size(int src_h, int src_w, double sc, int* h, int* w) {
  if (!h || !w || src_h <= 0 || src_w <= 0 || !(sc > 0.0)) return ARU_EINVAL;
  *w = cv_round(src_w * sc);
  *h = cv_round(src_h * sc);
  return ARU_OK;
}

int aru_scale_pages(aru_engine* e, const uint8_t* pages, int channels, int n, int src_h, int src_w, double sc, uint8_t* out) {
  if (!e || !pages || !out || n <= 0 || src_h <= 0 || src_w <= 0 || (channels != 1 && channels != 3))
    return e ? fail(e, ARU_EINVAL, "aru_scale_pages: bad argument") : ARU_EINVAL;
  CU(e, cudaSetDevice(e->device));
  ScalePlan sp;
  int rc = make_scale_plan(e, src_h, src_w, sc, &sp);
  if (rc) return rc;
  const size_t src_b = (size_t)src_h * src_w * channels, dst_b = (size_t)sp.dh * sp.dw * channels;
  const int mb = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, ((size_t)256 << 20) / src_b));
  uint8_t *d_in = nullptr, *d_out = nullptr;
  if (cudaMalloc((void**)&d_in, mb * src_b) || cudaMalloc((void**)&d_out, mb * dst_b)) {
    cudaFree(d_in); cudaFree(d_out); free_scale_plan(&sp);
    return fail(e, ARU_ENOMEM, "aru_scale_pages: device allocation failed");
  }
  for (int i0 = 0; i0 < n && rc == ARU_OK; i0 += mb) {
    const int cnt = std::min(mb, n - i0);
    cudaError_t err = cudaMemcpyAsync(d_in, pages + (size_t)i0 * src_b, cnt * src_b, cudaMemcpyHostToDevice, e->s_comp);
    if (err == cudaSuccess)
      err = sp.cubic ? launch_resize_cubic(e->s_comp, d_in, channels, cnt, src_h, src_w, d_out, sp.dh, sp.dw, sp.cx_idx,
                                           sp.cx_coef, sp.cy_idx, sp.cy_coef)
                     : launch_resize_area(e->s_comp, d_in, channels, cnt, src_h, src_w, d_out, sp.dh, sp.dw, sp.x_start,
                                          sp.x_si, sp.x_alpha, sp.y_start, sp.y_si, sp.y_alpha, sp.fast, sp.ix, sp.iy);
    if (err == cudaSuccess) err = cudaMemcpyAsync(out + (size_t)i0 * dst_b, d_out, cnt * dst_b, cudaMemcpyDeviceToHost, e->s_comp);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->s_comp);
    if (err != cudaSuccess) rc = fail(e, ARU_ECUDA, "aru_scale_pages: %s", cudaGetErrorString(err));
  }
  cudaFree(d_in); cudaFree(d_out); free_scale_plan(&sp);
  return rc;
}

int aru_separator_images(aru_engine* e, const uint8_t* images, int channels, int n, int src_h, int src_w, double sc,
                         double thr, float* out, uint8_t* out_u8, uint8_t* out_mask, uint8_t* horizontal, uint8_t* vertical) {
  if (!e || !images || n <= 0 || src_h <= 0 || src_w <= 0 || (channels != 1 && channels != 3))
    return e ? fail(e, ARU_EINVAL, "aru_separator_images: bad argument") : ARU_EINVAL;
  if (sc == 1.0)
    return aru_separator_pages(e, images, channels, n, src_h, src_w, thr, out, out_u8, out_mask, horizontal, vertical);
  if ((horizontal == nullptr) != (vertical == nullptr))
    return fail(e, ARU_EINVAL, "aru_separator_images: horizontal and vertical go together");
  CU(e, cudaSetDevice(e->device));
  ScalePlan sp;
  int rc = make_scale_plan(e, src_h, src_w, sc, &sp);
  if (rc) return rc;
  if (horizontal) {
    int k1, kv, k2;
    separator_kernel_sizes(sp.dh, sp.dw, &k1, &kv, &k2);
    if (k1 < 1 || kv < 1 || k2 < 1) {
      free_scale_plan(&sp);
      return fail(e, ARU_EINVAL, "aru_separator_images: scaled page %dx%d too small for the structuring elements", sp.dh, sp.dw);
    }
  }
  HostIO io;
  io.scale = &sp;
  io.in_u8 = images;
  io.channels = channels;
  io.out = out;
  io.out_u8 = out_u8;
  io.out_mask = out_mask;
  io.out_h = horizontal;
  io.out_v = vertical;
  io.cut = mask_cut(thr);
  rc = forward_host(e, io, n, sp.dh, sp.dw);
  free_scale_plan(&sp);
  return rc;
}

static int check_boxes(aru_engine* e, const int32_t* boxes, int n_boxes, int n, int h, int w) {
  for (int b = 0; b < n_boxes; ++b) {
    const int32_t* q = boxes + 5 * b;
    if (q[0] < 0 || q[0] >= n || q[1] < 0 || q[2] > h || q[3] < 0 || q[4] > w)
      return fail(e, ARU_EINVAL, "box %d = (page %d, rows [%d,%d), columns [%d,%d)) leaves the %d pages of %dx%d", b, q[0],
                  q[1], q[2], q[3], q[4], n, h, w);
  }
  return ARU_OK;
}

int aru_heading_pages(aru_engine* e, const uint8_t* pages, int channels, int n, int h, int w, const int32_t* boxes,
                      int n_boxes, uint64_t* sums, uint8_t* out_u8) {
  if (!e || !pages || n <= 0 || h <= 0 || w <= 0 || (channels != 1 && channels != 3) || n_boxes < 0 ||
      (n_boxes > 0 && (!boxes || !sums)))
    return e ? fail(e, ARU_EINVAL, "aru_heading_pages: bad argument") : ARU_EINVAL;
  int rc = check_boxes(e, boxes, n_boxes, n, h, w);
  if (rc) return rc;
  CU(e, cudaSetDevice(e->device));
  int* boxes_dev = nullptr;
  unsigned long long* sums_dev = nullptr;
  if (n_boxes > 0) {
    if (cudaMalloc((void**)&boxes_dev, (size_t)n_boxes * 5 * sizeof(int)) || cudaMalloc((void**)&sums_dev, (size_t)n_boxes * 8)) {
      cudaFree(boxes_dev);
      return fail(e, ARU_ENOMEM, "aru_heading_pages: device allocation failed");
    }
    cudaMemcpyAsync(boxes_dev, boxes, (size_t)n_boxes * 5 * sizeof(int), cudaMemcpyHostToDevice, e->s_comp);
  }
  HostIO io;
  io.in_u8 = pages;
  io.channels = channels;
  io.out_u8 = out_u8;
  io.boxes_dev = boxes_dev;
  io.n_boxes = n_boxes;
  io.sums_dev = sums_dev;
  rc = forward_host(e, io, n, h, w);
  if (rc == ARU_OK && n_boxes > 0) {
    cudaError_t err = cudaMemcpy(sums, sums_dev, (size_t)n_boxes * 8, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) rc = fail(e, ARU_ECUDA, "aru_heading_pages: %s", cudaGetErrorString(err));
  }
  cudaFree(boxes_dev);
  cudaFree(sums_dev);
  return rc;
}

int aru_heading_images(aru_engine* e, const uint8_t* images, int channels, int n, int src_h, int src_w, double sc,
                       const int32_t* boxes, int n_boxes, uint64_t* sums, uint8_t* out_u8) {
  if (!e || !images || n <= 0 || src_h <= 0 || src_w <= 0 || (channels != 1 && channels != 3) || n_boxes < 0 ||
      (n_boxes > 0 && (!boxes || !sums)))
    return e ? fail(e, ARU_EINVAL, "aru_heading_images: bad argument") : ARU_EINVAL;
  if (sc == 1.0) return aru_heading_pages(e, images, channels, n, src_h, src_w, boxes, n_boxes, sums, out_u8);
  CU(e, cudaSetDevice(e->device));
  ScalePlan sp;
  int rc = make_scale_plan(e, src_h, src_w, sc, &sp);
  if (rc) return rc;
  rc = check_boxes(e, boxes, n_boxes, n, sp.dh, sp.dw);
  int* boxes_dev = nullptr;
  unsigned long long* sums_dev = nullptr;
  if (rc == ARU_OK && n_boxes > 0) {
    if (cudaMalloc((void**)&boxes_dev, (size_t)n_boxes * 5 * sizeof(int)) || cudaMalloc((void**)&sums_dev, (size_t)n_boxes * 8))
      rc = fail(e, ARU_ENOMEM, "aru_heading_images: device allocation failed");
    else
      cudaMemcpyAsync(boxes_dev, boxes, (size_t)n_boxes * 5 * sizeof(int), cudaMemcpyHostToDevice, e->s_comp);
  }
  if (rc == ARU_OK) {
    HostIO io;
    io.scale = &sp;
    io.in_u8 = images;
    io.channels = channels;
    io.out_u8 = out_u8;
    io.boxes_dev = boxes_dev;
    io.n_boxes = n_boxes;
    io.sums_dev = sums_dev;
    rc = forward_host(e, io, n, sp.dh, sp.dw);
  }
  if (rc == ARU_OK && n_boxes > 0) {
    cudaError_t err = cudaMemcpy(sums, sums_dev, (size_t)n_boxes * 8, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) rc = fail(e, ARU_ECUDA, "aru_heading_images: %s", cudaGetErrorString(err));
  }
  cudaFree(boxes_dev);
  cudaFree(sums_dev);
  free_scale_plan(&sp);
  return rc;
}

int aru_box_sums(aru_engine* e, const uint8_t* u8, int n, int h, int w, int c, const int32_t* boxes, int n_boxes,
                 uint64_t* sums) {
  if (!e || !u8 || n <= 0 || h <= 0 || w <= 0 || c <= 0 || n_boxes < 0 || (n_boxes > 0 && (!boxes || !sums)))
    return e ? fail(e, ARU_EINVAL, "aru_box_sums: bad argument") : ARU_EINVAL;
  if (n_boxes == 0) return ARU_OK;
  int rc = check_boxes(e, boxes, n_boxes, n, h, w);
  if (rc) return rc;
  CU(e, cudaSetDevice(e->device));
  const size_t bytes = (size_t)n * h * w * c;
  uint8_t* d_u8 = nullptr;
  int* boxes_dev = nullptr;
  unsigned long long* sums_dev = nullptr;
  auto cleanup = [&]() { cudaFree(d_u8); cudaFree(boxes_dev); cudaFree(sums_dev); };
  if (cudaMalloc((void**)&d_u8, bytes) || cudaMalloc((void**)&boxes_dev, (size_t)n_boxes * 5 * sizeof(int)) ||
      cudaMalloc((void**)&sums_dev, (size_t)n_boxes * 8)) {
    cleanup();
    return fail(e, ARU_ENOMEM, "aru_box_sums: device allocation failed");
  }
  cudaError_t err = cudaMemcpyAsync(d_u8, u8, bytes, cudaMemcpyHostToDevice, e->s_comp);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(boxes_dev, boxes, (size_t)n_boxes * 5 * sizeof(int), cudaMemcpyHostToDevice, e->s_comp);
  if (err == cudaSuccess) err = launch_box_sums(e->s_comp, d_u8, n, h, w, c, 0, boxes_dev, n_boxes, sums_dev);
  if (err == cudaSuccess) err = cudaMemcpyAsync(sums, sums_dev, (size_t)n_boxes * 8, cudaMemcpyDeviceToHost, e->s_comp);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->s_comp);
  cleanup();
  if (err != cudaSuccess) return fail(e, ARU_ECUDA, "aru_box_sums: %s", cudaGetErrorString(err));
  return ARU_OK;
}

int aru_separator_post(aru_engine* e, const uint8_t* mask, int n, int h, int w, uint8_t* horizontal, uint8_t* vertical) {
  if (!e || !mask || !horizontal || !vertical || n <= 0 || h <= 0 || w <= 0)
    return e ? fail(e, ARU_EINVAL, "aru_separator_post: bad argument") : ARU_EINVAL;
  int k1, kv, k2;
  separator_kernel_sizes(h, w, &k1, &kv, &k2);
  if (k1 < 1 || kv < 1 || k2 < 1)
    return fail(e, ARU_EINVAL, "aru_separator_post: page %dx%d too small (structuring element %dx1 / 1x%d / %dx1)", h, w,
                k1, kv, k2);
  CU(e, cudaSetDevice(e->device));
  const size_t page_px = (size_t)h * w;
  const int mb = (int)std::max<long long>(1, std::min<long long>(n, (64LL << 20) / (long long)page_px));
  uint8_t *d_in = nullptr, *d_h = nullptr, *d_v = nullptr;
  void* scratch = nullptr;
  int rc = ARU_OK;
  auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_h); cudaFree(d_v); cudaFree(scratch); };
  if (cudaMalloc((void**)&d_in, mb * page_px) || cudaMalloc((void**)&d_h, mb * page_px) ||
      cudaMalloc((void**)&d_v, mb * page_px) || cudaMalloc(&scratch, separator_post_scratch_bytes(mb, h, w))) {
    cleanup();
    return fail(e, ARU_ENOMEM, "aru_separator_post: device allocation failed");
  }
  for (int i0 = 0; i0 < n && rc == ARU_OK; i0 += mb) {
    const int cnt = std::min(mb, n - i0);
    cudaError_t err = cudaMemcpyAsync(d_in, mask + (size_t)i0 * page_px, cnt * page_px, cudaMemcpyHostToDevice, e->s_comp);
    if (err == cudaSuccess)
      err = launch_separator_post(e->s_comp, d_in, cnt, h, w, cc_min_size((long long)page_px), k1, kv, k2, scratch, d_h, d_v);
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(horizontal + (size_t)i0 * page_px, d_h, cnt * page_px, cudaMemcpyDeviceToHost, e->s_comp);
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(vertical + (size_t)i0 * page_px, d_v, cnt * page_px, cudaMemcpyDeviceToHost, e->s_comp);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->s_comp);
    if (err != cudaSuccess) rc = fail(e, ARU_ECUDA, "aru_separator_post: %s", cudaGetErrorString(err));
  }
  cleanup();
  return rc;
}

int aru_cc_filter(aru_engine* e, const uint8_t* mask, int n, int h, int w, int min_size, uint8_t* out) {
  if (!e || !mask || !out || n <= 0 || h <= 0 || w <= 0)
    return e ? fail(e, ARU_EINVAL, "aru_cc_filter: bad argument") : ARU_EINVAL;
  CU(e, cudaSetDevice(e->device));
  const size_t page_px = (size_t)h * w;
  const int mb = (int)std::max<long long>(1, std::min<long long>(n, (64LL << 20) / (long long)page_px));
  uint8_t *d_in = nullptr, *d_out = nullptr;
  void* scratch = nullptr;
  auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_out); cudaFree(scratch); };
  if (cudaMalloc((void**)&d_in, mb * page_px) || cudaMalloc((void**)&d_out, mb * page_px) ||
      cudaMalloc(&scratch, separator_post_scratch_bytes(mb, h, w))) {
    cleanup();
    return fail(e, ARU_ENOMEM, "aru_cc_filter: device allocation failed");
  }
  int rc = ARU_OK;
  for (int i0 = 0; i0 < n && rc == ARU_OK; i0 += mb) {
    const int cnt = std::min(mb, n - i0);
    cudaError_t err = cudaMemcpyAsync(d_in, mask + (size_t)i0 * page_px, cnt * page_px, cudaMemcpyHostToDevice, e->s_comp);
    if (err == cudaSuccess) err = launch_cc_filter(e->s_comp, d_in, cnt, h, w, min_size, scratch, d_out);
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(out + (size_t)i0 * page_px, d_out, cnt * page_px, cudaMemcpyDeviceToHost, e->s_comp);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->s_comp);
    if (err != cudaSuccess) rc = fail(e, ARU_ECUDA, "aru_cc_filter: %s", cudaGetErrorString(err));
  }
  cleanup();
  return rc;
}

int aru_open_rect(aru_engine* e, const uint8_t* mask, int n, int h, int w, int kw, int kh, uint8_t* out) {
  if (!e || !mask || !out || n <= 0 || h <= 0 || w <= 0 || kw < 1 || kh < 1 || (kw > 1 && kh > 1))
    return e ? fail(e, ARU_EINVAL, "aru_open_rect: bad argument (one of kw, kh must be 1)") : ARU_EINVAL;
  CU(e, cudaSetDevice(e->device));
  const size_t px = (size_t)n * h * w;
  uint8_t *d_in = nullptr, *d_out = nullptr;
  void* scratch = nullptr;
  auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_out); cudaFree(scratch); };
  if (cudaMalloc((void**)&d_in, px) || cudaMalloc((void**)&d_out, px) ||
      cudaMalloc(&scratch, separator_post_scratch_bytes(n, h, w))) {
    cleanup();
    return fail(e, ARU_ENOMEM, "aru_open_rect: device allocation failed");
  }
  cudaError_t err = cudaMemcpyAsync(d_in, mask, px, cudaMemcpyHostToDevice, e->s_comp);
  if (err == cudaSuccess) err = launch_open_rect(e->s_comp, d_in, n, h, w, kw, kh, scratch, d_out);
  if (err == cudaSuccess) err = cudaMemcpyAsync(out, d_out, px, cudaMemcpyDeviceToHost, e->s_comp);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->s_comp);
  cleanup();
  if (err != cudaSuccess) return fail(e, ARU_ECUDA, "aru_open_rect: %s", cudaGetErrorString(err));
  return ARU_OK;
}

int aru_swt_distance(aru_engine* e, const uint8_t* gray, int n, int h, int w, int dark_on_bright, uint8_t* out,
                     int32_t* thresholds) {
  if (!e || !gray || !out || n <= 0 || h <= 0 || w <= 0)
    return e ? fail(e, ARU_EINVAL, "aru_swt_distance: bad argument") : ARU_EINVAL;
  if (h >= 60000 || w >= 60000) return fail(e, ARU_EUNSUP, "aru_swt_distance: page dimensions must be below 60000");
  CU(e, cudaSetDevice(e->device));
  const size_t page_px = (size_t)h * w;
  const int mb = (int)std::max<long long>(1, std::min<long long>(n, (256LL << 20) / (long long)page_px));
  uint8_t *d_in = nullptr, *d_out = nullptr;
  void* scratch = nullptr;
  auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_out); cudaFree(scratch); };
  if (cudaMalloc((void**)&d_in, mb * page_px) || cudaMalloc((void**)&d_out, mb * page_px) ||
      cudaMalloc(&scratch, swt_scratch_bytes(mb, h, w))) {
    cleanup();
    cudaGetLastError();
    return fail(e, ARU_ENOMEM, "aru_swt_distance: device allocation failed");
  }
  int rc = ARU_OK;
  std::vector<int> thr(mb);
  for (int i0 = 0; i0 < n && rc == ARU_OK; i0 += mb) {
    const int cnt = std::min(mb, n - i0);
    cudaError_t err = cudaMemcpyAsync(d_in, gray + (size_t)i0 * page_px, cnt * page_px, cudaMemcpyHostToDevice, e->s_comp);
    if (err == cudaSuccess) err = launch_swt_distance(e->s_comp, d_in, cnt, h, w, dark_on_bright, scratch, d_out, thr.data());
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(out + (size_t)i0 * page_px, d_out, cnt * page_px, cudaMemcpyDeviceToHost, e->s_comp);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->s_comp);
    if (err != cudaSuccess) rc = fail(e, ARU_ECUDA, "aru_swt_distance: %s", cudaGetErrorString(err));
    if (thresholds)
      for (int i = 0; i < cnt; ++i) thresholds[i0 + i] = thr[i];
  }
  cleanup();
  return rc;
}

int aru_pages_to_input(aru_engine* e, const uint8_t* pages, int channels, int n, int h, int w, float* out) {
  if (!e || !pages || !out || n <= 0 || h <= 0 || w <= 0 || (channels != 1 && channels != 3))
    return e ? fail(e, ARU_EINVAL, "aru_pages_to_input: bad argument") : ARU_EINVAL;
  CU(e, cudaSetDevice(e->device));
  const size_t px = (size_t)n * h * w;
  uint8_t* d_in = nullptr;
  float* d_out = nullptr;
  if (cudaMalloc((void**)&d_in, px * channels) || cudaMalloc((void**)&d_out, px * sizeof(float))) {
    cudaFree(d_in);
    cudaFree(d_out);
    return fail(e, ARU_ENOMEM, "aru_pages_to_input: device allocation failed");
  }
  cudaError_t err = cudaMemcpyAsync(d_in, pages, px * channels, cudaMemcpyHostToDevice, e->s_comp);
  if (err == cudaSuccess) err = launch_pages_to_input(e->s_comp, d_in, channels, (long long)px, d_out, nullptr);
  if (err == cudaSuccess) err = cudaMemcpyAsync(out, d_out, px * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->s_comp);
  cudaFree(d_in);
  cudaFree(d_out);
  if (err != cudaSuccess) return fail(e, ARU_ECUDA, "aru_pages_to_input: %s", cudaGetErrorString(err));
  return ARU_OK;
}

int aru_forward_device(aru_engine* e, const float* in, int n, int h, int w, float* out, uint8_t* out_u8,
                       uint8_t* out_mask, float thr, void* stream) {
  if (!e || !in || n <= 0 || h <= 0 || w <= 0) return e ? fail(e, ARU_EINVAL, "aru_forward_device: bad argument") : ARU_EINVAL;
  CU(e, cudaSetDevice(e->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->s_comp;
  const int mb = pick_micro_batch(e, n, h, w);
  const size_t page_px = (size_t)h * w;
  const int C = e->n_class;
  for (int i0 = 0; i0 < n; i0 += mb) {
    const int cnt = std::min(mb, n - i0);
    Plan* p = nullptr;
    int rc = get_plan(e, cnt, h, w, &p);
    if (rc) return rc;
    e->cur = p;
    const float* in_i = in + (size_t)i0 * page_px;
    float* out_i = out ? out + (size_t)i0 * page_px * C : nullptr;
    int slot = 0;
    // The caller's buffers are bound in place when they can be (16-byte aligned, a float32 result is wanted): the first
    // layers read `in`, the classifier writes `out`, no staging copies.  One graph per (in, out) pair.
    if (out_i && !e->keep_all && ((reinterpret_cast<uintptr_t>(in_i) | reinterpret_cast<uintptr_t>(out_i)) & 15) == 0 &&
        getenv("ARU_NO_INPLACE") == nullptr) {
      int lru = 2;
      for (int s2 = 2; s2 < 2 + Plan::EXT_SLOTS; ++s2) {
        if (p->in_dev[s2] == in_i && p->out_dev[s2] == out_i) { slot = s2; break; }
        if (p->ext_use[s2] < p->ext_use[lru]) lru = s2;
      }
      if (!slot) {
        slot = lru;
        if (p->graph[slot]) {
          CU(e, cudaStreamSynchronize(st));   // the graph may still be running
          cudaGraphExecDestroy(p->graph[slot]);
          p->graph[slot] = nullptr;
        }
        p->in_dev[slot] = const_cast<float*>(in_i);
        p->out_dev[slot] = out_i;
        p->eager_passes[slot] = 0;
      }
      p->ext_use[slot] = ++p->ext_clock;
    }
    if (!slot)
      CU(e, cudaMemcpyAsync(p->in_dev[0], in_i, (size_t)cnt * page_px * sizeof(float), cudaMemcpyDeviceToDevice, st));
    rc = enqueue_forward(e, p, slot, st);
    if (rc) return rc;
    if (out && !slot)
      CU(e, cudaMemcpyAsync(out_i, p->out_dev[0], (size_t)cnt * page_px * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (out_u8 || out_mask) {
      cudaError_t err = launch_quantize(st, p->out_dev[slot], out_u8 ? out_u8 + (size_t)i0 * page_px * C : nullptr,
                                        out_mask ? out_mask + (size_t)i0 * page_px : nullptr, (long long)cnt * page_px, C, thr, -1);
      if (err != cudaSuccess) return fail(e, ARU_ECUDA, "quantize launch: %s", cudaGetErrorString(err));
    }
  }
  return ARU_OK;
}

uint64_t aru_last_ticket(const aru_engine* e) { return e ? e->last_ticket : 0; }

int aru_wait(aru_engine* e, uint64_t ticket) {
  if (!e) return ARU_EINVAL;
  if (ticket == 0 || ticket > e->last_ticket) return fail(e, ARU_EINVAL, "aru_wait: unknown ticket");
  CU(e, cudaSetDevice(e->device));
  if (e->last_ticket - ticket >= 8) return check_err_flag(e);   // its event was recycled: 8 later calls were submitted, and
                                                               // a call is only submitted after ... its predecessors' waits
  cudaEvent_t ev = e->tickets[ticket % 8];
  if (ev) CU(e, cudaEventSynchronize(ev));
  return check_err_flag(e);
}

int aru_sync(aru_engine* e) {
  if (!e) return ARU_EINVAL;
  CU(e, cudaSetDevice(e->device));
  CU(e, cudaStreamSynchronize(e->s_h2d));
  CU(e, cudaStreamSynchronize(e->s_comp));
  CU(e, cudaStreamSynchronize(e->s_d2h));
  return check_err_flag(e);
}

int aru_launches_per_forward(const aru_engine* e) { return (e && e->cur) ? e->cur->launches : 0; }

int aru_buffer_dims(const aru_engine* e, int buf, int* h, int* w, int* c) {
  if (!e || !e->cur || buf < 0 || buf >= (int)e->buffers.size()) return ARU_EINVAL;
  const BufPlan& b = e->cur->bufs[buf];
  if (h) *h = b.h;
  if (w) *w = b.w;
  if (c) *c = b.channels;
  return b.sized ? ARU_OK : ARU_EINVAL;
}

int aru_read_buffer(aru_engine* e, int buf, int n_index, float* out_nhwc, size_t out_floats) {
  if (!e || !e->cur || buf < 0 || buf >= (int)e->buffers.size() || !out_nhwc)
    return e ? fail(e, ARU_EINVAL, "aru_read_buffer: bad argument") : ARU_EINVAL;
  Plan* p = e->cur;
  const BufPlan& b = p->bufs[buf];
  if (!b.sized || n_index < 0 || n_index >= p->n) return fail(e, ARU_EINVAL, "aru_read_buffer: bad buffer / page index");
  if (p->buf_dead[buf]) return fail(e, ARU_EINVAL, "aru_read_buffer: buffer %d is an intermediate of a fused launch (set ARU_OPT_KEEP_ALL)", buf);
  const size_t cnt = (size_t)b.h * b.w * b.channels;
  if (out_floats < cnt) return fail(e, ARU_EINVAL, "aru_read_buffer: output too small (%zu < %zu)", out_floats, cnt);
  CU(e, cudaSetDevice(e->device));
  CU(e, cudaStreamSynchronize(e->s_comp));
  if (b.kind != KIND_PLANAR) {
    const float* src = f32_ptr(e, p, buf, 0) + (size_t)n_index * cnt;
    CU(e, cudaMemcpy(out_nhwc, src, cnt * sizeof(float), cudaMemcpyDeviceToHost));
    return ARU_OK;
  }
  float* tmp = nullptr;
  CU(e, cudaMalloc((void**)&tmp, (size_t)p->n * cnt * sizeof(float)));
  aru_view v{buf, 0, b.channels};
  cudaError_t err = launch_unpack_nhwc(e->s_comp, make_pv(e, p, v), b.geo, tmp);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->s_comp);
  if (err == cudaSuccess) err = cudaMemcpy(out_nhwc, tmp + (size_t)n_index * cnt, cnt * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(tmp);
  if (err != cudaSuccess) return fail(e, ARU_ECUDA, "aru_read_buffer: %s", cudaGetErrorString(err));
  return ARU_OK;
}

int aru_profile_ops(aru_engine* e, int iters, float* ms, int n_ms) {
  if (!e || !e->cur || !ms || iters <= 0 || n_ms < (int)e->ops.size())
    return e ? fail(e, ARU_EINVAL, "aru_profile_ops: bad argument / no plan") : ARU_EINVAL;
  CU(e, cudaSetDevice(e->device));
  Plan* p = e->cur;
  cudaEvent_t a, b;
  CU(e, cudaEventCreate(&a));
  CU(e, cudaEventCreate(&b));
  for (int i = 0; i < (int)e->ops.size(); ++i) {
    int rc = run_op(e, p, i, 0, e->s_comp);  // warm
    if (rc) return rc;
    CU(e, cudaEventRecord(a, e->s_comp));
    for (int k = 0; k < iters; ++k)
      if ((rc = run_op(e, p, i, 0, e->s_comp))) return rc;
    CU(e, cudaEventRecord(b, e->s_comp));
    CU(e, cudaEventSynchronize(b));
    float t = 0.f;
    CU(e, cudaEventElapsedTime(&t, a, b));
    ms[i] = t / iters;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return check_err_flag(e);
}

const char* aru_op_kernel_name(const aru_engine* e, int op) {
  if (!e || !e->cur || op < 0 || op >= (int)e->ops.size()) return "";
  return e->cur->kernel[op];
}

int aru_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return ARU_EINVAL;
  cudaError_t err = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
  if (err != cudaSuccess) {
    g_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(err);
    cudaGetLastError();
    return ARU_ENOMEM;
  }
  return ARU_OK;
}

void aru_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
