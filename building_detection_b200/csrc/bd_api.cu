// bd_api.cu -- C ABI (include/bd_b200.h): context, network plans (native launch lists), tiler/stitcher.
// Fusion and contour entry points live in post.cu.
#include "../../include/bd_b200.h"

#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#include "common.cuh"
#include "conv_umma.cuh"
#include "dw_tma.cuh"
#include "kernels.cuh"
#include "post_ws.cuh"

namespace bd {
std::string& last_error() {
  static thread_local std::string e;
  return e;
}
int fail(const std::string& msg) {
  last_error() = msg;
  return 1;
}
}  // namespace bd

using namespace bd;


namespace {

struct BufInfo {
  int H, W, C, dtype, kind;
  size_t bytes, offset;
  int first = INT32_MAX, last = -1;  // first / last plan step (bd_plan_add_* call) that touches the buffer
  bool shared = false;               // its arena range is reused by buffers with disjoint lifetimes
  std::vector<std::pair<int, int>> written;  // channel intervals some op writes
  int hard_read_end = 0;             // highest channel (+1) read by a kernel that cannot clip its reads
  int valid_c = -1;                  // shared buffers: channels [valid_c, C) are never written (TMA reads clip there)
};

struct Op {
  int kclass;     // 0 conv umma, 1 conv direct, 2 other
  double flops;
  int launches;
  std::function<int(cudaStream_t)> run;
};

// programmatic dependent launch of every plan kernel (BD_PDL=0 turns it off for A/B measurements)
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("BD_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
#define BD_LAUNCH(...) BD_CUDA(::bd::launch_k(pdl_enabled(), __VA_ARGS__))

inline int grid_for(size_t total, int num_sms) {
  size_t b = (total + k::TPB - 1) / k::TPB;
  size_t cap = static_cast<size_t>(num_sms) * 8;
  return static_cast<int>(std::max<size_t>(1, std::min(b, cap)));
}

}  // namespace

// Every bd_plan_add_* / finalize call is also appended to the plan's log (op code + plain arguments + weight arrays):
// bd_plan_save writes the log, bd_plan_load replays it through the same entry points -- a host without the Python graph
// builder loads pre-lowered plans (tools/export_plans.py) and runs the whole path through the C ABI.
enum LogOp : uint32_t { LOG_BUFFER = 1, LOG_CONV, LOG_DWCONV, LOG_MAXPOOL, LOG_ADDN, LOG_GAP, LOG_DENSE, LOG_GATE, LOG_SKFUSE,
                        LOG_BCAST, LOG_FINALIZE };
struct LogWriter {
  std::vector<uint8_t>& b;
  template <class T> void pod(const T& v) { const uint8_t* p = reinterpret_cast<const uint8_t*>(&v); b.insert(b.end(), p, p + sizeof(T)); }
  void arr(const void* p, size_t bytes) {
    pod(static_cast<uint64_t>(bytes));
    const uint8_t* q = static_cast<const uint8_t*>(p);
    b.insert(b.end(), q, q + bytes);
    b.insert(b.end(), (8 - bytes % 8) % 8, 0);
  }
};
struct LogReader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  template <class T> T pod() {
    T v{};
    if (p + sizeof(T) > end) { ok = false; return v; }
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  const void* arr(size_t* bytes) {
    const uint64_t n = pod<uint64_t>();
    const size_t padded = n + (8 - n % 8) % 8;
    if (!ok || p + padded > end) { ok = false; *bytes = 0; return nullptr; }
    const void* q = p;
    p += padded;
    *bytes = n;
    return q;
  }
};

struct bd_plan {
  bd_ctx* ctx = nullptr;
  int batch = 0;
  bool finalized = false;
  std::vector<uint8_t> log;
  std::vector<BufInfo> bufs;
  std::vector<std::function<int(bd_plan*)>> builders;  // run at finalize, append to ops
  std::vector<Op> ops;
  std::vector<void*> dev_allocs;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  // Arena reuse: map buffers whose lifetimes (first .. last step touching them) are disjoint share address ranges.
  // On by default; bd_plan_set_arena_reuse(plan, 0) gives every buffer a range of its own (readable after a forward).
  bool reuse = true;
  size_t arena_bytes_flat = 0;  // what the arena would take without reuse
  int input_buf = -1, logits_buf = -1, logits_up = 1;
  void touch(int buf) {  // called by bd_plan_add_* for every map buffer the op reads or writes
    if (buf < 0 || buf >= static_cast<int>(bufs.size())) return;
    const int step = static_cast<int>(builders.size());
    bufs[buf].first = std::min(bufs[buf].first, step);
    bufs[buf].last = std::max(bufs[buf].last, step);
  }
  void touch_w(const bd_tref& r) {
    touch(r.buf);
    if (r.buf >= 0 && r.buf < static_cast<int>(bufs.size())) bufs[r.buf].written.emplace_back(r.c0, r.c0 + r.c);
  }
  // clip_ok: the reader goes through a TMA tensor map whose channel extent can be cut at the written channels (reads
  // beyond it return zero); every other kernel reads its whole slice from memory
  void touch_r(const bd_tref& r, bool clip_ok = false) {
    touch(r.buf);
    if (!clip_ok && r.buf >= 0 && r.buf < static_cast<int>(bufs.size()))
      bufs[r.buf].hard_read_end = std::max(bufs[r.buf].hard_read_end, r.c0 + r.c);
  }
  // channels of a read slice that hold written data (the rest of a shared buffer is another tenant's garbage)
  int valid_channels(const bd_tref& r) const {
    const BufInfo& b = bufs[r.buf];
    if (b.valid_c < 0) return r.c;
    return std::max(0, std::min(r.c, b.valid_c - r.c0));
  }
  float* cur_probs = nullptr;
  uint8_t* cur_mask = nullptr;
  // CUDA graph of one forward writing the argmax masks to graph_mask (bd_scene_run): captured on first use
  cudaGraphExec_t graph_exec = nullptr;
  uint8_t* graph_mask = nullptr;
  bool graph_failed = false;

  int upload(const void* host, size_t bytes, void** out) {
    void* d = nullptr;
    BD_CUDA(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
    dev_allocs.push_back(d);
    BD_CUDA(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
    *out = d;
    return 0;
  }
  int scratch(size_t bytes, void** out) {
    void* d = nullptr;
    BD_CUDA(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
    dev_allocs.push_back(d);
    *out = d;
    return 0;
  }
  int check_ref(const bd_tref& r, bool vec8) const {
    BD_CHECK(r.buf >= 0 && r.buf < static_cast<int>(bufs.size()), "tensor ref: bad buffer id");
    const BufInfo& b = bufs[r.buf];
    BD_CHECK(b.kind == BD_MAP, "tensor ref: not a map buffer");
    BD_CHECK(r.c0 >= 0 && r.c > 0 && r.c0 + r.c <= b.C, "tensor ref: slice out of range");
    if (vec8)
      BD_CHECK(b.dtype == BD_F16 && r.c % 8 == 0 && r.c0 % 8 == 0 && b.C % 8 == 0,
               "tensor ref: op needs h16 slices aligned to 8 channels");
    return 0;
  }
  int check_vec(int id, int c) const {
    BD_CHECK(id >= 0 && id < static_cast<int>(bufs.size()) && bufs[id].kind == BD_VEC, "bad vector buffer id");
    BD_CHECK(c < 0 || bufs[id].C == c, "vector buffer has the wrong length");
    return 0;
  }
  TView tview(const bd_tref& r) const {
    const BufInfo& b = bufs[r.buf];
    TView v;
    v.base = arena + b.offset; v.N = batch; v.H = b.H; v.W = b.W; v.ctot = b.C; v.c0 = r.c0; v.c = r.c;
    v.f32 = (b.dtype == BD_F32);
    return v;
  }
  k::View kview(const bd_tref& r) const {
    const BufInfo& b = bufs[r.buf];
    k::View v;
    v.base = arena + b.offset; v.H = b.H; v.W = b.W; v.ctot = b.C; v.c0 = r.c0; v.c = r.c; v.f32 = (b.dtype == BD_F32);
    return v;
  }
  float* vecptr(int id) const { return reinterpret_cast<float*>(arena + bufs[id].offset); }
};

extern "C" {

const char* bd_last_error(void) { return bd::last_error().c_str(); }
const char* bd_version(void) { return "bd_b200 0.1 (sm_100a)"; }

int bd_create(int device, bd_ctx** out) {
  BD_CHECK(out != nullptr, "null out pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(std::string("no CUDA device available (this library has no CPU path): ") + cudaGetErrorString(e));
  BD_CHECK(device >= 0 && device < count, "device index out of range");
  DeviceGuard guard(device);  // the caller's current device is restored on return
  cudaDeviceProp prop;
  BD_CUDA(cudaGetDeviceProperties(&prop, device));
  BD_CHECK(prop.major == 10, "this library is built for sm_100a (Blackwell B200) only");
  bd_ctx* c = new bd_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  BD_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->h_scalar), 64));
  if (const char* s = getenv("BD_UMMA_SMEM_KB")) c->umma_smem_kb = std::max(48, std::min(226, atoi(s)));
  if (const char* s = getenv("BD_UMMA_GROUP")) c->umma_group = std::max(0, std::min(9, atoi(s)));
  if (const char* s = getenv("BD_UMMA_MAX_N")) c->umma_max_block_n = std::max(16, std::min(256, atoi(s) / 16 * 16));
  *out = c;
  return 0;
}

void bd_destroy(bd_ctx* ctx) {
  BD_ON_CTX(ctx);
  if (!ctx) return;
  if (ctx->d_ys) cudaFree(ctx->d_ys);
  if (ctx->d_xs) cudaFree(ctx->d_xs);
  if (ctx->d_all_ys) cudaFree(ctx->d_all_ys);
  if (ctx->d_all_xs) cudaFree(ctx->d_all_xs);
  ctx->arena.release();
  ctx->pool.release();
  if (ctx->h_scalar) cudaFreeHost(ctx->h_scalar);
  for (void* hp : ctx->h_pts)
    if (hp) cudaFreeHost(hp);
  if (ctx->capture_stream) cudaStreamDestroy(ctx->capture_stream);
  delete ctx;
}

int64_t bd_launch_count(bd_ctx* ctx) { return ctx ? ctx->launches : 0; }

int bd_debug_read_trace(bd_ctx* ctx, long long* host_dst, int max_events) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && ctx->trace_buf && host_dst, "no trace buffer (set BD_UMMA_TRACE=1 before building the plan)");
  BD_CUDA(cudaDeviceSynchronize());
  (void)max_events;
  BD_CUDA(cudaMemcpy(host_dst, ctx->trace_buf, sizeof(long long) * 4 * 4096, cudaMemcpyDeviceToHost));
  BD_CUDA(cudaMemset(ctx->trace_buf, 0, sizeof(long long) * 4 * 4096));
  return 0;
}

// ------------------------------------------------------------------------------------------ plans
int bd_plan_create(bd_ctx* ctx, int batch, bd_plan** out) {
  BD_CHECK(ctx && out, "null argument");
  BD_CHECK(batch >= 1 && batch <= 64, "batch must be in [1,64]");
  bd_plan* p = new bd_plan();
  p->ctx = ctx;
  p->batch = batch;
  {  // arena reuse is the default (BD_ARENA_REUSE=0 or bd_plan_set_arena_reuse(plan, 0) keep every buffer apart)
    const char* e = getenv("BD_ARENA_REUSE");
    p->reuse = !(e && e[0] == '0');
  }
  *out = p;
  return 0;
}

void bd_plan_destroy(bd_plan* p) {
  BD_ON_PLAN(p);
  if (!p) return;
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  for (void* d : p->dev_allocs) cudaFree(d);
  if (p->arena) cudaFree(p->arena);
  delete p;
}

int bd_plan_add_buffer(bd_plan* p, int h, int w, int c, int dtype, int kind) {
  if (!p || p->finalized || h < 1 || w < 1 || c < 1) {
    fail("bd_plan_add_buffer: bad arguments");
    return -1;
  }
  { LogWriter lw{p->log}; lw.pod(LOG_BUFFER); lw.pod(h); lw.pod(w); lw.pod(c); lw.pod(dtype); lw.pod(kind); }
  BufInfo b;
  b.H = h; b.W = w; b.C = c; b.dtype = dtype; b.kind = kind;
  if (kind == BD_VEC) b.bytes = static_cast<size_t>(p->batch) * c * 4;
  else b.bytes = static_cast<size_t>(p->batch) * h * w * c * (dtype == BD_F32 ? 4 : 2);
  b.offset = 0;
  p->bufs.push_back(b);
  return static_cast<int>(p->bufs.size()) - 1;
}

int bd_plan_add_conv(bd_plan* p, const bd_conv_desc* dptr) {
  BD_CHECK(p && dptr && !p->finalized, "bad arguments");
  bd_conv_desc d = *dptr;
  BD_CHECK(d.ntaps >= 1 && d.ntaps <= BD_MAX_TAPS, "conv: bad tap count");
  if (p->check_ref(d.x, false) || p->check_ref(d.y, false)) return 1;
  const bool has_res = d.res.buf >= 0;
  if (has_res && p->check_ref(d.res, false)) return 1;
  const int cin = d.x.c, cout = d.y.c;
  const BufInfo& yb = p->bufs[d.y.buf];
  BD_CHECK(yb.H == d.ho * d.out_scale && yb.W == d.wo * d.out_scale, "conv: output geometry mismatch");
  if (has_res) BD_CHECK(d.res.c == cout && p->bufs[d.res.buf].H == yb.H && p->bufs[d.res.buf].W == yb.W,
                        "conv: residual geometry mismatch");
  std::shared_ptr<std::vector<uint16_t>> w(new std::vector<uint16_t>(d.w_host, d.w_host + static_cast<size_t>(d.ntaps) * cout * cin));
  std::shared_ptr<std::vector<float>> b(new std::vector<float>(d.bias_host, d.bias_host + cout));
  // fused SeparableConv2D: depthwise weights [9][cin] -> fp16 (the values are fp16-representable, graph.py)
  std::shared_ptr<std::vector<uint16_t>> dww;
  if (d.dw_w_host) {
    BD_CHECK(d.path == BD_CONV_UMMA && d.ntaps == 1 && d.stride == 1 && d.out_scale == 1,
             "fused separable conv: the pointwise stage must be a stride-1 1x1 UMMA convolution");
    dww.reset(new std::vector<uint16_t>(static_cast<size_t>(9) * cin));
    for (size_t i = 0; i < dww->size(); ++i) {
      const __half h = __float2half_rn(d.dw_w_host[i]);
      memcpy(&(*dww)[i], &h, 2);
    }
  }
  const int dw_relu = d.dw_relu_in;
  {
    LogWriter lw{p->log};
    lw.pod(LOG_CONV);
    bd_conv_desc plain = d;
    plain.w_host = nullptr; plain.bias_host = nullptr; plain.dw_w_host = nullptr;
    lw.pod(plain);
    lw.arr(w->data(), w->size() * 2);
    lw.arr(b->data(), b->size() * 4);
    lw.pod(static_cast<int32_t>(d.dw_w_host ? 1 : 0));
    if (d.dw_w_host) lw.arr(d.dw_w_host, static_cast<size_t>(9) * cin * 4);
  }
  d.w_host = nullptr; d.bias_host = nullptr; d.dw_w_host = nullptr;
  p->touch_r(d.x, d.path == BD_CONV_UMMA); p->touch_w(d.y); if (has_res) p->touch_r(d.res);
  p->builders.push_back([d, w, b, dww, dw_relu, has_res, cin, cout](bd_plan* pl) -> int {
    void *wd = nullptr, *bdv = nullptr, *dwd = nullptr;
    if (pl->upload(w->data(), w->size() * 2, &wd) || pl->upload(b->data(), b->size() * 4, &bdv)) return 1;
    if (dww && pl->upload(dww->data(), dww->size() * 2, &dwd)) return 1;
    Op op;
    op.flops = 2.0 * pl->batch * d.ho * d.wo * (static_cast<double>(cout) * cin * d.ntaps + (dww ? 9.0 * cin : 0.0));
    op.launches = 1;
    bd_ctx* ctx = pl->ctx;
    if (d.path == BD_CONV_SMALL) {
      // CUDA-core path for <= 16 output channels (k::conv_small_kernel): fp32 weights [tap][ci][CO], only the channels
      // that carry non-zero weights / bias
      BD_CHECK(!has_res && d.stride == 1 && d.out_scale == 1 && d.ntaps <= 9 && d.act_post == BD_ACT_NONE,
               "small conv: stride 1, no residual, at most 9 taps");
      const BufInfo& xb = pl->bufs[d.x.buf];
      const BufInfo& yb = pl->bufs[d.y.buf];
      BD_CHECK(xb.dtype == BD_F16 && d.x.c % 8 == 0 && d.x.c0 % 8 == 0 && xb.C % 8 == 0, "small conv: fp16 input in 8-channel vectors");
      BD_CHECK(yb.dtype == BD_F32 || (d.y.c % 8 == 0 && d.y.c0 % 8 == 0 && yb.C % 8 == 0), "small conv: output slice alignment");
      int cin_used = 0, cout_used = 0;
      auto h2f = [](uint16_t b) { return __half2float(__ushort_as_half(b)); };
      for (int t = 0; t < d.ntaps; ++t)
        for (int co = 0; co < cout; ++co)
          for (int ci = 0; ci < cin; ++ci)
            if ((*w)[(static_cast<size_t>(t) * cout + co) * cin + ci] & 0x7FFF) { cin_used = std::max(cin_used, ci + 1); cout_used = std::max(cout_used, co + 1); }
      for (int co = 0; co < cout; ++co)
        if ((*b)[co] != 0.0f) cout_used = std::max(cout_used, co + 1);
      cin_used = std::max(cin_used, 1); cout_used = std::max(cout_used, 1);
      BD_CHECK(cout_used <= 16, "small conv: more than 16 output channels in use");
      const int CO = cout_used <= 2 ? 2 : cout_used <= 4 ? 4 : cout_used <= 8 ? 8 : 16;
      std::vector<float> wf(static_cast<size_t>(d.ntaps) * cin_used * CO, 0.0f), bf(CO, 0.0f);
      for (int t = 0; t < d.ntaps; ++t)
        for (int ci = 0; ci < cin_used; ++ci)
          for (int co = 0; co < cout_used; ++co)
            wf[(static_cast<size_t>(t) * cin_used + ci) * CO + co] = h2f((*w)[(static_cast<size_t>(t) * cout + co) * cin + ci]);
      for (int co = 0; co < cout_used; ++co) bf[co] = (*b)[co];
      void *wfd = nullptr, *bfd = nullptr;
      if (pl->upload(wf.data(), wf.size() * 4, &wfd) || pl->upload(bf.data(), bf.size() * 4, &bfd)) return 1;
      k::SmallParams q;
      memset(&q, 0, sizeof(q));
      q.x = pl->kview(d.x); q.y = pl->kview(d.y);
      q.N = pl->batch; q.Ho = d.ho; q.Wo = d.wo; q.ntaps = d.ntaps;
      for (int t = 0; t < d.ntaps; ++t) { q.dy[t] = d.dy[t]; q.dx[t] = d.dx[t]; }
      q.cin_used = cin_used; q.cout_used = cout_used; q.act = d.act_pre;
      q.w = static_cast<const float*>(wfd); q.bias = static_cast<const float*>(bfd);
      const int nvec = cdiv(cin_used, 8);
      const int LPP = nvec <= 1 ? 1 : nvec <= 2 ? 2 : nvec <= 4 ? 4 : 8;  // lanes per pixel
      const size_t total = static_cast<size_t>(pl->batch) * d.ho * d.wo;
      BD_CHECK(total < (1ull << 31), "small conv: too many output pixels for 32-bit indices");
      const bool fast1x1 = d.ntaps == 1 && d.dy[0] == 0 && d.dx[0] == 0 && cin_used <= 8 * LPP;  // 4 pixels per lane group
      const int grid = grid_for(fast1x1 ? (total * LPP + 3) / 4 : total * LPP, ctx->num_sms * 4);
      const size_t smem = (wf.size() + CO) * sizeof(float);
      BD_CHECK(smem <= 48 * 1024, "small conv: weights do not fit into shared memory");
      op.kclass = 2;
      op.flops = 2.0 * pl->batch * d.ho * d.wo * static_cast<double>(cout_used) * cin_used * d.ntaps;
      typedef void (*SmallKernel)(k::SmallParams);
      static const SmallKernel table[4][4] = {
          {k::conv_small_kernel<2, 1>, k::conv_small_kernel<2, 2>, k::conv_small_kernel<2, 4>, k::conv_small_kernel<2, 8>},
          {k::conv_small_kernel<4, 1>, k::conv_small_kernel<4, 2>, k::conv_small_kernel<4, 4>, k::conv_small_kernel<4, 8>},
          {k::conv_small_kernel<8, 1>, k::conv_small_kernel<8, 2>, k::conv_small_kernel<8, 4>, k::conv_small_kernel<8, 8>},
          {k::conv_small_kernel<16, 1>, k::conv_small_kernel<16, 2>, k::conv_small_kernel<16, 4>, k::conv_small_kernel<16, 8>}};
      const SmallKernel kern = table[CO == 2 ? 0 : CO == 4 ? 1 : CO == 8 ? 2 : 3][LPP == 1 ? 0 : LPP == 2 ? 1 : LPP == 4 ? 2 : 3];
      op.run = [q, grid, smem, kern, ctx](cudaStream_t s) -> int {
        ctx->launches++;
        BD_LAUNCH(kern, dim3(grid), dim3(k::TPB), smem, s, q);
        BD_CUDA(cudaGetLastError());
        return 0;
      };
    } else
    if (d.path == BD_CONV_UMMA) {
      std::shared_ptr<umma::Launch> L(new umma::Launch());
      TView x = pl->tview(d.x), y = pl->tview(d.y), r;
      if (has_res) r = pl->tview(d.res);
      if (umma::prepare(L.get(), x, y, has_res ? &r : nullptr, d.ntaps, d.dy, d.dx, d.stride, d.ho, d.wo, d.act_pre,
                        d.act_post, d.out_scale, d.out_oy, d.out_ox, static_cast<const h16*>(wd),
                        static_cast<const float*>(bdv), ctx->umma_smem_kb, ctx->umma_max_block_n, ctx->num_sms, ctx->umma_group,
                        static_cast<const h16*>(dwd), dw_relu, pl->valid_channels(d.x)))
        return 1;
      if (const char* tr = getenv("BD_UMMA_TRACE")) {  // debug: event trace of CTA 0 (tools/umma_trace.py)
        void* tbuf = nullptr;
        if (pl->scratch(sizeof(long long) * 4 * 4096, &tbuf)) return 1;
        cudaMemset(tbuf, 0, sizeof(long long) * 4 * 4096);
        L->p.trace = static_cast<long long*>(tbuf);
        ctx->trace_buf = tbuf;
        (void)tr;
      }
      op.kclass = 0;
      op.run = [L, ctx](cudaStream_t s) -> int { ctx->launches++; return umma::launch(*L, s, pdl_enabled()); };
    } else {
      k::DirectParams q;
      memset(&q, 0, sizeof(q));
      q.x = pl->kview(d.x); q.y = pl->kview(d.y);
      if (has_res) q.res = pl->kview(d.res);
      q.N = pl->batch; q.Ho = d.ho; q.Wo = d.wo; q.stride = d.stride; q.ntaps = d.ntaps;
      for (int t = 0; t < d.ntaps; ++t) { q.dy[t] = d.dy[t]; q.dx[t] = d.dx[t]; }
      q.act_pre = d.act_pre; q.act_post = d.act_post; q.out_scale = d.out_scale; q.out_oy = d.out_oy; q.out_ox = d.out_ox;
      q.w = static_cast<const h16*>(wd); q.bias = static_cast<const float*>(bdv);
      const size_t total = static_cast<size_t>(pl->batch) * d.ho * d.wo * cdiv(cout, k::DC_CO);
      const int grid = static_cast<int>(std::min<size_t>((total + k::TPB - 1) / k::TPB, 1u << 20));
      op.kclass = 1;
      op.run = [q, grid, ctx](cudaStream_t s) -> int {
        ctx->launches++;
        BD_LAUNCH(k::conv_direct_kernel, dim3(grid), dim3(k::TPB), 0, s, q);
        BD_CUDA(cudaGetLastError());
        return 0;
      };
    }
    pl->ops.push_back(op);
    return 0;
  });
  return 0;
}

int bd_plan_add_dwconv(bd_plan* p, bd_tref x, bd_tref y, int stride, int pad_t, int pad_l, int relu_in,
                       const float* w_host) {
  BD_CHECK(p && !p->finalized && w_host, "bad arguments");
  if (p->check_ref(x, true) || p->check_ref(y, true)) return 1;
  BD_CHECK(x.c == y.c, "dwconv: channel mismatch");
  { LogWriter lw{p->log}; lw.pod(LOG_DWCONV); lw.pod(x); lw.pod(y); lw.pod(stride); lw.pod(pad_t); lw.pod(pad_l); lw.pod(relu_in);
    lw.arr(w_host, static_cast<size_t>(9) * x.c * 4); }
  // depthwise weights are stored as fp16 on the device (the host passes fp16-representable values)
  std::shared_ptr<std::vector<uint16_t>> w(new std::vector<uint16_t>(9 * static_cast<size_t>(x.c)));
  for (size_t i = 0; i < w->size(); ++i) {
    const __half hv = __float2half_rn(std::max(-65504.0f, std::min(65504.0f, w_host[i])));
    memcpy(&(*w)[i], &hv, 2);
  }
  p->touch_r(x); p->touch_w(y);
  p->builders.push_back([=](bd_plan* pl) -> int {
    void* wd = nullptr;
    if (pl->upload(w->data(), w->size() * 2, &wd)) return 1;
    k::DwParams q;
    q.x = pl->kview(x); q.y = pl->kview(y);
    q.N = pl->batch; q.Ho = q.y.H; q.Wo = q.y.W; q.stride = stride; q.pad_t = pad_t; q.pad_l = pad_l; q.relu_in = relu_in;
    q.w = static_cast<const h16*>(wd);
    BD_CHECK(stride == 1 || stride == 2, "dwconv: stride must be 1 or 2");
    {
      // shared-memory tiles moved by TMA (dw_tma.cuh) for the stride-1 'same' case on maps small enough that the
      // register-window kernel is latency bound (BD_DW_TMA=0 turns it off)
      TView xv = pl->tview(x), yv = pl->tview(y);
      static const bool dw_tma_on = [] { const char* e = getenv("BD_DW_TMA"); return !(e && e[0] == '0'); }();
      if (dw_tma_on && dwt::eligible(xv, yv, stride, pad_t, pad_l)) {
        std::shared_ptr<dwt::Launch> L(new dwt::Launch());
        if (dwt::prepare(L.get(), xv, yv, static_cast<const h16*>(wd), relu_in, pl->ctx->num_sms)) return 1;
        bd_ctx* ctx = pl->ctx;
        Op op;
        op.kclass = 2; op.launches = 1;
        op.flops = 2.0 * pl->batch * q.Ho * q.Wo * static_cast<double>(x.c) * 9;
        op.run = [L, ctx](cudaStream_t s) -> int { ctx->launches++; return dwt::launch(*L, s, pdl_enabled()); };
        pl->ops.push_back(op);
        return 0;
      }
    }
    // 4-channel vectors (half the registers, twice the resident warps) whenever the 8-channel version could not
    // fill the machine with threads; both are exact and round identically
    const size_t total8 = static_cast<size_t>(pl->batch) * cdiv(q.Ho, k::DW_ROWS) * q.Wo * (x.c / 8);
    bd_ctx* ctx = pl->ctx;
    const int vec = (x.c % 8 != 0 || total8 < static_cast<size_t>(ctx->num_sms) * 2048 * 2) ? 4 : 8;
    BD_CHECK(q.x.c % vec == 0 && q.x.c0 % vec == 0 && q.x.ctot % vec == 0 && q.y.c0 % vec == 0 && q.y.ctot % vec == 0,
             "dwconv: channel slices must be vector aligned");
    const size_t total = total8 * (8 / vec);
    BD_CHECK(total < (1ull << 31), "dwconv: too many work items");
    const int grid = static_cast<int>(cdiv64(static_cast<int64_t>(total), k::TPB));
    Op op;
    op.kclass = 2; op.launches = 1;
    op.flops = 2.0 * pl->batch * q.Ho * q.Wo * static_cast<double>(x.c) * 9;
    op.run = [q, grid, ctx, stride, vec](cudaStream_t s) -> int {
      ctx->launches++;
      if (stride == 1 && vec == 8) BD_LAUNCH(k::dwconv3x3_kernel<1, 8>, dim3(grid), dim3(k::TPB), 0, s, q);
      else if (stride == 1) BD_LAUNCH(k::dwconv3x3_kernel<1, 4>, dim3(grid), dim3(k::TPB), 0, s, q);
      else if (vec == 8) BD_LAUNCH(k::dwconv3x3_kernel<2, 8>, dim3(grid), dim3(k::TPB), 0, s, q);
      else BD_LAUNCH(k::dwconv3x3_kernel<2, 4>, dim3(grid), dim3(k::TPB), 0, s, q);
      BD_CUDA(cudaGetLastError());
      return 0;
    };
    pl->ops.push_back(op);
    return 0;
  });
  return 0;
}

int bd_plan_add_maxpool(bd_plan* p, bd_tref x, bd_tref y, int kk, int stride, int pad_t, int pad_l) {
  BD_CHECK(p && !p->finalized, "bad arguments");
  if (p->check_ref(x, true) || p->check_ref(y, true)) return 1;
  BD_CHECK(x.c == y.c && kk >= 1 && kk <= 3, "maxpool: bad arguments");
  { LogWriter lw{p->log}; lw.pod(LOG_MAXPOOL); lw.pod(x); lw.pod(y); lw.pod(kk); lw.pod(stride); lw.pod(pad_t); lw.pod(pad_l); }
  p->touch_r(x); p->touch_w(y);
  p->builders.push_back([=](bd_plan* pl) -> int {
    k::PoolParams q;
    q.x = pl->kview(x); q.y = pl->kview(y);
    q.N = pl->batch; q.Ho = q.y.H; q.Wo = q.y.W; q.k = kk; q.stride = stride; q.pad_t = pad_t; q.pad_l = pad_l;
    const size_t total = static_cast<size_t>(pl->batch) * q.Ho * q.Wo * (x.c / 8);
    BD_CHECK(total < (1ull << 31), "maxpool: too many vectors for 32-bit indices");
    q.fd_cg = make_fastdiv(x.c / 8); q.fd_wo = make_fastdiv(q.Wo); q.fd_ho = make_fastdiv(q.Ho);
    bd_ctx* ctx = pl->ctx;
    const int grid = grid_for(total, ctx->num_sms * 4);
    Op op;
    op.kclass = 2; op.launches = 1; op.flops = 0;
    op.run = [q, grid, ctx, kk](cudaStream_t s) -> int {
      ctx->launches++;
      if (kk == 1) BD_LAUNCH(k::maxpool_kernel<1>, dim3(grid), dim3(k::TPB), 0, s, q);
      else if (kk == 2) BD_LAUNCH(k::maxpool_kernel<2>, dim3(grid), dim3(k::TPB), 0, s, q);
      else BD_LAUNCH(k::maxpool_kernel<3>, dim3(grid), dim3(k::TPB), 0, s, q);
      BD_CUDA(cudaGetLastError());
      return 0;
    };
    pl->ops.push_back(op);
    return 0;
  });
  return 0;
}

int bd_plan_add_addn(bd_plan* p, int n, const bd_tref* xs, const int32_t* fs, bd_tref y, int act) {
  BD_CHECK(p && !p->finalized && xs && fs && n >= 1 && n <= 4, "bad arguments");
  if (p->check_ref(y, true)) return 1;
  std::vector<bd_tref> xv(xs, xs + n);
  std::vector<int> fv(fs, fs + n);
  { LogWriter lw{p->log}; lw.pod(LOG_ADDN); lw.pod(n); for (int i = 0; i < n; ++i) { lw.pod(xs[i]); lw.pod(fs[i]); } lw.pod(y); lw.pod(act); }
  for (int i = 0; i < n; ++i) {
    if (p->check_ref(xv[i], true)) return 1;
    const BufInfo& b = p->bufs[xv[i].buf];
    BD_CHECK(xv[i].c == y.c && b.H * fv[i] == p->bufs[y.buf].H && b.W * fv[i] == p->bufs[y.buf].W,
             "addn: geometry mismatch");
  }
  for (int i = 0; i < n; ++i) p->touch_r(xv[i]);
  p->touch_w(y);
  p->builders.push_back([=](bd_plan* pl) -> int {
    k::AddnParams q;
    memset(&q, 0, sizeof(q));
    for (int i = 0; i < n; ++i) {
      q.x[i] = pl->kview(xv[i]);
      int sh = 0;
      while ((1 << sh) < fv[i]) ++sh;
      BD_CHECK((1 << sh) == fv[i], "addn: up-sampling factors must be powers of two");
      q.sh[i] = sh;
    }
    q.y = pl->kview(y); q.n_in = n; q.N = pl->batch; q.act = act;
    const size_t total = static_cast<size_t>(pl->batch) * q.y.H * q.y.W * (y.c / 8);
    BD_CHECK(total < (1ull << 31), "addn: too many vectors for 32-bit indices");
    q.fd_cg = make_fastdiv(y.c / 8); q.fd_w = make_fastdiv(q.y.W); q.fd_h = make_fastdiv(q.y.H);
    bd_ctx* ctx = pl->ctx;
    const int grid = grid_for((total + 1) / 2, ctx->num_sms * 4);
    Op op;
    op.kclass = 2; op.launches = 1; op.flops = 0;
    op.run = [q, grid, ctx](cudaStream_t s) -> int {
      ctx->launches++;
      BD_LAUNCH(k::addn_kernel, dim3(grid), dim3(k::TPB), 0, s, q);
      BD_CUDA(cudaGetLastError());
      return 0;
    };
    pl->ops.push_back(op);
    return 0;
  });
  return 0;
}

int bd_plan_add_gap(bd_plan* p, bd_tref x, int y_vec) {
  BD_CHECK(p && !p->finalized, "bad arguments");
  if (p->check_ref(x, true) || p->check_vec(y_vec, x.c)) return 1;
  BD_CHECK(x.c / 8 <= k::TPB, "gap: too many channels");
  { LogWriter lw{p->log}; lw.pod(LOG_GAP); lw.pod(x); lw.pod(y_vec); }
  p->touch_r(x);
  p->builders.push_back([=](bd_plan* pl) -> int {
    k::GapParams q;
    q.x = pl->kview(x); q.N = pl->batch;
    const int HW = q.x.H * q.x.W;
    // enough blocks to fill the machine four times over, at least 32 pixels and (for the large maps) about 2048
    // pixels per block
    bd_ctx* ctx = pl->ctx;
    // (sized for batch 16, and NOT a function of the batch: a tile's result must not depend on its batch neighbours)
    // (about 2048 pixels per block on the large maps: with 512 the per-block reduction, fence and ticket were ~20 % of a
    // block's life; 37 x batch blocks still fill the machine)
    const int want = std::max(cdiv(HW, 2048), cdiv(4 * 148, 16));
    q.splits = std::max(1, std::min(std::min(want, 1024), std::max(1, HW / 32)));
    void *part = nullptr, *tick = nullptr;
    if (pl->scratch(static_cast<size_t>(pl->batch) * q.splits * x.c * 4, &part)) return 1;
    if (pl->scratch(static_cast<size_t>(pl->batch) * 4, &tick)) return 1;
    BD_CUDA(cudaMemset(tick, 0, static_cast<size_t>(pl->batch) * 4));
    q.partial = static_cast<float*>(part);
    q.tickets = static_cast<unsigned*>(tick);
    q.out = pl->vecptr(y_vec);
    const int rows = k::TPB / (x.c / 8);
    const size_t smem = static_cast<size_t>(rows) * x.c * 4;
    Op op;
    op.kclass = 2; op.launches = 1; op.flops = 0;
    op.run = [q, smem, ctx](cudaStream_t s) -> int {
      ctx->launches += 1;
      BD_LAUNCH(k::gap_kernel, dim3(dim3(q.splits, q.N)), dim3(k::TPB), smem, s, q);
      BD_CUDA(cudaGetLastError());
      return 0;
    };
    pl->ops.push_back(op);
    return 0;
  });
  return 0;
}

int bd_plan_add_dense(bd_plan* p, int n_in, const int32_t* x_vecs, int y_vec, int cin, int cout, int act,
                      const float* w_host, const float* b_host) {
  BD_CHECK(p && !p->finalized && x_vecs && w_host && b_host && n_in >= 1 && n_in <= 5, "bad arguments");
  std::vector<int> xv(x_vecs, x_vecs + n_in);
  for (int i = 0; i < n_in; ++i)
    if (p->check_vec(xv[i], cin)) return 1;
  if (p->check_vec(y_vec, cout)) return 1;
  std::shared_ptr<std::vector<float>> w(new std::vector<float>(w_host, w_host + static_cast<size_t>(cin) * cout));
  std::shared_ptr<std::vector<float>> b(new std::vector<float>(b_host, b_host + cout));
  { LogWriter lw{p->log}; lw.pod(LOG_DENSE); lw.pod(n_in); for (int i = 0; i < n_in; ++i) lw.pod(x_vecs[i]); lw.pod(y_vec); lw.pod(cin);
    lw.pod(cout); lw.pod(act); lw.arr(w_host, static_cast<size_t>(cin) * cout * 4); lw.arr(b_host, static_cast<size_t>(cout) * 4); }
  p->builders.push_back([=](bd_plan* pl) -> int {
    void *wd = nullptr, *bdv = nullptr;
    if (pl->upload(w->data(), w->size() * 4, &wd) || pl->upload(b->data(), b->size() * 4, &bdv)) return 1;
    k::DenseParams q;
    memset(&q, 0, sizeof(q));
    for (int i = 0; i < n_in; ++i) q.x[i] = pl->vecptr(xv[i]);
    q.y = pl->vecptr(y_vec); q.w = static_cast<const float*>(wd); q.b = static_cast<const float*>(bdv);
    q.n_in = n_in; q.N = pl->batch; q.cin = cin; q.cout = cout; q.act = act;
    const int grid = cdiv(pl->batch * cout * 32, k::TPB);
    bd_ctx* ctx = pl->ctx;
    Op op;
    op.kclass = 2; op.launches = 1; op.flops = 2.0 * pl->batch * cin * static_cast<double>(cout);
    op.run = [q, grid, ctx](cudaStream_t s) -> int {
      ctx->launches++;
      BD_LAUNCH(k::dense_kernel, dim3(grid), dim3(k::TPB), 0, s, q);
      BD_CUDA(cudaGetLastError());
      return 0;
    };
    pl->ops.push_back(op);
    return 0;
  });
  return 0;
}

int bd_plan_add_gate(bd_plan* p, int mode, bd_tref x, bd_tref y, int v_vec, bd_tref sref, const float* w_host,
                     float bscalar) {
  BD_CHECK(p && !p->finalized && mode >= 0 && mode <= 2, "bad arguments");
  if (p->check_ref(x, true) || p->check_ref(y, true) || p->check_vec(v_vec, x.c)) return 1;
  BD_CHECK(x.c == y.c, "gate: channel mismatch");
  if (mode == BD_GATE_BAM) {
    if (p->check_ref(sref, false)) return 1;
    BD_CHECK(sref.c == 1, "gate: BAM spatial logits must have one channel");
  }
  std::shared_ptr<std::vector<float>> w;
  if (mode == BD_GATE_SCSE) {
    BD_CHECK(w_host != nullptr, "gate: scSE needs spatial weights");
    w.reset(new std::vector<float>(w_host, w_host + x.c));
  }
  { LogWriter lw{p->log}; lw.pod(LOG_GATE); lw.pod(mode); lw.pod(x); lw.pod(y); lw.pod(v_vec); lw.pod(sref); lw.pod(bscalar);
    lw.pod(static_cast<int32_t>(w ? 1 : 0)); if (w) lw.arr(w->data(), w->size() * 4); }
  p->touch_r(x); p->touch_w(y); if (mode == BD_GATE_BAM) p->touch_r(sref);
  p->builders.push_back([=](bd_plan* pl) -> int {
    k::GateParams q;
    memset(&q, 0, sizeof(q));
    q.x = pl->kview(x); q.y = pl->kview(y); q.v = pl->vecptr(v_vec); q.mode = mode; q.N = pl->batch; q.b = bscalar;
    if (mode == BD_GATE_BAM) q.s = pl->kview(sref);
    if (mode == BD_GATE_SCSE) {
      void* wd = nullptr;
      if (pl->upload(w->data(), w->size() * 4, &wd)) return 1;
      q.w = static_cast<const float*>(wd);
    }
    const size_t npix = static_cast<size_t>(pl->batch) * q.x.H * q.x.W;
    BD_CHECK(npix * (x.c / 8) < (1ull << 31), "gate: too many vectors for 32-bit indices");
    q.fd_cg = make_fastdiv(x.c / 8); q.fd_hw = make_fastdiv(static_cast<uint32_t>(q.x.H) * q.x.W);
    bd_ctx* ctx = pl->ctx;
    Op op;
    op.kclass = 2; op.launches = 1;
    op.flops = mode == BD_GATE_SCSE ? 2.0 * npix * x.c : 0.0;
    if (mode == BD_GATE_SCSE) {
      int lanes = 1;
      while (lanes < 32 && lanes * 2 <= x.c / 8) lanes *= 2;
      const int vpl = (x.c / 8) / lanes;  // vectors per lane
      BD_CHECK(vpl * lanes == x.c / 8 && (vpl == 1 || vpl == 2 || vpl == 4),
               "gate: scSE needs a channel count of 8 * 2^k up to 1024");
      const int U = 4 / vpl;  // pixels in flight per lane group
      const size_t warps = (npix + (32 / lanes) * U - 1) / ((32 / lanes) * U);
      const int grid = grid_for(warps * 32, ctx->num_sms * 4);
      op.run = [q, grid, lanes, vpl, ctx](cudaStream_t s) -> int {
        ctx->launches++;
        if (vpl == 1) BD_LAUNCH((k::gate_scse_kernel<1, 4>), dim3(grid), dim3(k::TPB), 0, s, q, lanes);
        else if (vpl == 2) BD_LAUNCH((k::gate_scse_kernel<2, 2>), dim3(grid), dim3(k::TPB), 0, s, q, lanes);
        else BD_LAUNCH((k::gate_scse_kernel<4, 1>), dim3(grid), dim3(k::TPB), 0, s, q, lanes);
        BD_CUDA(cudaGetLastError());
        return 0;
      };
    } else {
      const int grid = grid_for((npix * (x.c / 8) + 3) / 4, ctx->num_sms * 4);
      op.run = [q, grid, ctx](cudaStream_t s) -> int {
        ctx->launches++;
        BD_LAUNCH(k::gate_kernel, dim3(grid), dim3(k::TPB), 0, s, q);
        BD_CUDA(cudaGetLastError());
        return 0;
      };
    }
    pl->ops.push_back(op);
    return 0;
  });
  return 0;
}

int bd_plan_add_skfuse(bd_plan* p, const bd_tref* xs4, int g_vec, const int32_t* logit_vecs5, bd_tref y,
                       const float* scale_host, const float* shift_host) {
  BD_CHECK(p && !p->finalized && xs4 && logit_vecs5 && scale_host && shift_host, "bad arguments");
  if (p->check_ref(y, true) || p->check_vec(g_vec, y.c)) return 1;
  std::vector<bd_tref> xv(xs4, xs4 + 4);
  std::vector<int> lv(logit_vecs5, logit_vecs5 + 5);
  for (int i = 0; i < 4; ++i) {
    if (p->check_ref(xv[i], true)) return 1;
    BD_CHECK(xv[i].c == y.c, "skfuse: channel mismatch");
  }
  for (int i = 0; i < 5; ++i)
    if (p->check_vec(lv[i], y.c)) return 1;
  std::shared_ptr<std::vector<float>> sc(new std::vector<float>(scale_host, scale_host + y.c));
  std::shared_ptr<std::vector<float>> sh(new std::vector<float>(shift_host, shift_host + y.c));
  { LogWriter lw{p->log}; lw.pod(LOG_SKFUSE); for (int i = 0; i < 4; ++i) lw.pod(xs4[i]); lw.pod(g_vec); for (int i = 0; i < 5; ++i) lw.pod(logit_vecs5[i]);
    lw.pod(y); lw.arr(scale_host, static_cast<size_t>(y.c) * 4); lw.arr(shift_host, static_cast<size_t>(y.c) * 4); }
  for (int i = 0; i < 4; ++i) p->touch_r(xv[i]);
  p->touch_w(y);
  p->builders.push_back([=](bd_plan* pl) -> int {
    void *scd = nullptr, *shd = nullptr;
    if (pl->upload(sc->data(), sc->size() * 4, &scd) || pl->upload(sh->data(), sh->size() * 4, &shd)) return 1;
    k::SkParams q;
    memset(&q, 0, sizeof(q));
    for (int i = 0; i < 4; ++i) q.x[i] = pl->kview(xv[i]);
    q.y = pl->kview(y); q.g = pl->vecptr(g_vec);
    for (int i = 0; i < 5; ++i) q.lg[i] = pl->vecptr(lv[i]);
    q.scale = static_cast<const float*>(scd); q.shift = static_cast<const float*>(shd); q.N = pl->batch;
    const size_t total = static_cast<size_t>(pl->batch) * q.y.H * q.y.W * (y.c / 8);
    bd_ctx* ctx = pl->ctx;
    const int grid = grid_for(total, ctx->num_sms * 4);
    Op op;
    op.kclass = 2; op.launches = 1; op.flops = 0;
    op.run = [q, grid, ctx](cudaStream_t s) -> int {
      ctx->launches++;
      BD_LAUNCH(k::skfuse_kernel, dim3(grid), dim3(k::TPB), 0, s, q);
      BD_CUDA(cudaGetLastError());
      return 0;
    };
    pl->ops.push_back(op);
    return 0;
  });
  return 0;
}

int bd_plan_add_bcast(bd_plan* p, int v_vec, bd_tref y) {
  BD_CHECK(p && !p->finalized, "bad arguments");
  if (p->check_ref(y, true) || p->check_vec(v_vec, y.c)) return 1;
  { LogWriter lw{p->log}; lw.pod(LOG_BCAST); lw.pod(v_vec); lw.pod(y); }
  p->touch_w(y);
  p->builders.push_back([=](bd_plan* pl) -> int {
    k::BcastParams q;
    q.y = pl->kview(y); q.v = pl->vecptr(v_vec); q.N = pl->batch;
    const size_t total = static_cast<size_t>(pl->batch) * q.y.H * q.y.W * (y.c / 8);
    bd_ctx* ctx = pl->ctx;
    const int grid = grid_for(total, ctx->num_sms * 4);
    Op op;
    op.kclass = 2; op.launches = 1; op.flops = 0;
    op.run = [q, grid, ctx](cudaStream_t s) -> int {
      ctx->launches++;
      BD_LAUNCH(k::bcast_kernel, dim3(grid), dim3(k::TPB), 0, s, q);
      BD_CUDA(cudaGetLastError());
      return 0;
    };
    pl->ops.push_back(op);
    return 0;
  });
  return 0;
}

int bd_plan_finalize(bd_plan* p, int input_buf, int logits_buf, int logits_up) {
  BD_ON_PLAN(p);
  BD_CHECK(p && !p->finalized, "bad arguments");
  { LogWriter lw{p->log}; lw.pod(LOG_FINALIZE); lw.pod(input_buf); lw.pod(logits_buf); lw.pod(logits_up); }
  const int nb = static_cast<int>(p->bufs.size());
  if (input_buf >= 0) {
    const BufInfo& ib = p->bufs[input_buf];
    BD_CHECK(input_buf < nb && ib.dtype == BD_F16 && ib.C == 32 && ib.kind == BD_MAP && ib.H == ib.W &&
                 (ib.H == 512 || ib.H == 256),
             "bad input buffer (expected the im2col'ed stem input: fp16, 32 channels, 512x512 or 256x256)");
  }
  if (logits_buf >= 0) {
    BD_CHECK(logits_buf < nb && p->bufs[logits_buf].dtype == BD_F32 && p->bufs[logits_buf].C == 2 && logits_up >= 1,
             "bad logits buffer");
  }
  size_t off = 0;
  for (BufInfo& b : p->bufs) {
    b.offset = off;
    off += (b.bytes + 1023) / 1024 * 1024;
  }
  p->arena_bytes_flat = std::max<size_t>(off, 1024);
  if (p->reuse) {
    // Lifetime of a map buffer = first .. last step that reads or writes it (the input is written before step 0, the
    // logits are read by the softmax head after the last).  Buffers are placed in order of first use at the lowest
    // offset that is free of every already placed buffer with an intersecting lifetime (an op's outputs therefore
    // never alias its inputs).  Every kernel waits for the complete previous kernel before it touches data
    // (griddepcontrol.wait), so step order is execution order.  Pooled vectors and untouched buffers keep ranges of
    // their own.  Nothing may rely on the zero fill of the arena: channel padding is never read (TMA boxes are clipped
    // by the tensor map, the CUDA-core kernels loop over the slice) and padded slices are written whole.
    const int NB = static_cast<int>(p->bufs.size());
    if (input_buf >= 0) p->bufs[input_buf].first = -1;
    if (logits_buf >= 0) p->bufs[logits_buf].last = INT32_MAX;
    std::vector<int> order;
    size_t fixed_end = 0;
    for (int i = 0; i < NB; ++i) {
      BufInfo& b = p->bufs[i];
      bool movable = b.kind == BD_MAP && b.last >= 0 && b.first != INT32_MAX;
      if (movable && i != input_buf) {
        // Channels no op writes (728-channel maps are stored with a 768-channel pitch and read at that width with zero
        // weights) hold zeros only in a range of the buffer's own.  In a shared range they are garbage: the buffer may
        // share only if the written channels are one run [0, valid_c) and every reader beyond valid_c is a TMA map
        // that can be cut there (out-of-range elements read as zero).
        std::sort(b.written.begin(), b.written.end());
        int run_end = 0;
        bool holes = false;
        for (const auto& iv : b.written) {
          if (iv.first > run_end) { holes = true; break; }
          run_end = std::max(run_end, iv.second);
        }
        if (holes || run_end == 0 || b.hard_read_end > run_end) movable = false;
        else b.valid_c = run_end;
      }
      if (movable) { order.push_back(i); b.shared = !(i == input_buf || i == logits_buf); }
      else { b.offset = fixed_end; fixed_end += (b.bytes + 1023) / 1024 * 1024; }
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) {
      const BufInfo &x = p->bufs[a], &y = p->bufs[c];
      return x.first != y.first ? x.first < y.first : x.bytes > y.bytes;
    });
    std::vector<int> placed;
    size_t end = fixed_end;
    for (int i : order) {
      BufInfo& b = p->bufs[i];
      const size_t need = (b.bytes + 1023) / 1024 * 1024;
      std::vector<std::pair<size_t, size_t>> busy;  // ranges of placed buffers alive during b's lifetime
      for (int j : placed) {
        const BufInfo& o = p->bufs[j];
        if (o.first <= b.last && b.first <= o.last) busy.emplace_back(o.offset, o.offset + (o.bytes + 1023) / 1024 * 1024);
      }
      std::sort(busy.begin(), busy.end());
      size_t at = fixed_end;
      for (const auto& r : busy) {
        if (at + need <= r.first) break;
        at = std::max(at, r.second);
      }
      b.offset = at;
      end = std::max(end, at + need);
      placed.push_back(i);
    }
    off = end;
  }
  p->arena_bytes = std::max<size_t>(off, 1024);
  BD_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->arena), p->arena_bytes));
  BD_CUDA(cudaMemset(p->arena, 0, p->arena_bytes));
  p->input_buf = input_buf; p->logits_buf = logits_buf; p->logits_up = logits_up;
  for (auto& b : p->builders)
    if (b(p)) return 1;
  p->builders.clear();
  if (logits_buf >= 0) {
    const BufInfo& lb = p->bufs[logits_buf];
    const float* lg = reinterpret_cast<const float*>(p->arena + lb.offset);
    const int H = lb.H * logits_up, W = lb.W * logits_up, N = p->batch, up = logits_up;
    bd_ctx* ctx = p->ctx;
    BD_CHECK(W % 4 == 0, "softmax head: the output width must be a multiple of 4");
    const int grid = grid_for(static_cast<size_t>(N) * H * W / 4, ctx->num_sms * 4);
    Op op;
    op.kclass = 2; op.launches = 1; op.flops = 0;
    op.run = [p, lg, N, H, W, up, grid, ctx](cudaStream_t s) -> int {
      if (!p->cur_probs && !p->cur_mask) return 0;
      BD_CHECK((reinterpret_cast<uintptr_t>(p->cur_probs) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->cur_mask) & 3) == 0,
               "bd_plan_run: probs must be 16-byte aligned and mask 4-byte aligned");
      ctx->launches++;
      BD_LAUNCH(k::softmax2_kernel, dim3(grid), dim3(k::TPB), 0, s, lg, N, H, W, up, p->cur_probs, p->cur_mask);
      BD_CUDA(cudaGetLastError());
      return 0;
    };
    p->ops.push_back(op);
  }
  BD_CUDA(cudaDeviceSynchronize());
  p->finalized = true;
  return 0;
}

int bd_plan_run(bd_plan* p, const float* x_dev, float* probs_dev, uint8_t* mask_dev, void* stream) {
  BD_ON_PLAN(p);
  BD_CHECK(p && p->finalized, "plan not finalized");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x_dev) {
    BD_CHECK(p->input_buf >= 0, "plan has no input buffer");
    const BufInfo& ib = p->bufs[p->input_buf];
    const size_t work = static_cast<size_t>(p->batch) * ib.H * ib.W * 4;
    p->ctx->launches++;
    h16* dst = reinterpret_cast<h16*>(p->arena + ib.offset);
    if (ib.H == 512) BD_LAUNCH(k::input_convert_kernel<1>, dim3(grid_for(work, p->ctx->num_sms * 4)), dim3(k::TPB), 0, s, x_dev, p->batch, dst);
    else BD_LAUNCH(k::input_convert_kernel<2>, dim3(grid_for(work, p->ctx->num_sms * 4)), dim3(k::TPB), 0, s, x_dev, p->batch, dst);
    BD_CUDA(cudaGetLastError());
  }
  p->cur_probs = probs_dev;
  p->cur_mask = mask_dev;
  for (Op& op : p->ops)
    if (op.run(s)) return 1;
  return 0;
}

int bd_plan_run_head(bd_plan* p, float* probs_dev, uint8_t* mask_dev, void* stream) {
  BD_ON_PLAN(p);
  BD_CHECK(p && p->finalized && p->logits_buf >= 0, "plan has no softmax head");
  p->cur_probs = probs_dev;
  p->cur_mask = mask_dev;
  return p->ops.back().run(static_cast<cudaStream_t>(stream));
}

int bd_plan_run_host(bd_plan* p, const float* x_host, float* probs_host, uint8_t* mask_host) {
  BD_ON_PLAN(p);
  BD_CHECK(p && p->finalized && p->input_buf >= 0 && p->logits_buf >= 0, "plan not runnable from host buffers");
  const BufInfo& ib = p->bufs[p->input_buf];
  const BufInfo& lb = p->bufs[p->logits_buf];
  const size_t npix = static_cast<size_t>(p->batch) * lb.H * p->logits_up * lb.W * p->logits_up;
  float* dprobs = nullptr;
  uint8_t* dmask = nullptr;
  float* dx = nullptr;
  if (x_host) {
    const size_t xbytes = static_cast<size_t>(p->batch) * 512 * 512 * 3 * sizeof(float);
    BD_CUDA(cudaMalloc(reinterpret_cast<void**>(&dx), xbytes));
    BD_CUDA(cudaMemcpy(dx, x_host, xbytes, cudaMemcpyHostToDevice));
  }
  if (probs_host) BD_CUDA(cudaMalloc(reinterpret_cast<void**>(&dprobs), npix * 2 * 4));
  if (mask_host) BD_CUDA(cudaMalloc(reinterpret_cast<void**>(&dmask), npix));
  int rc = bd_plan_run(p, dx, dprobs, dmask, nullptr);
  if (!rc) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = fail(std::string("forward failed: ") + cudaGetErrorString(e));
  }
  if (!rc && probs_host) cudaMemcpy(probs_host, dprobs, npix * 2 * 4, cudaMemcpyDeviceToHost);
  if (!rc && mask_host) cudaMemcpy(mask_host, dmask, npix, cudaMemcpyDeviceToHost);
  p->cur_probs = nullptr;
  p->cur_mask = nullptr;
  if (dx) cudaFree(dx);
  if (dprobs) cudaFree(dprobs);
  if (dmask) cudaFree(dmask);
  return rc;
}

void* bd_plan_buffer_ptr(bd_plan* p, int buf) {
  if (!p || !p->finalized || buf < 0 || buf >= static_cast<int>(p->bufs.size())) return nullptr;
  return p->arena + p->bufs[buf].offset;
}
size_t bd_plan_buffer_bytes(bd_plan* p, int buf) {
  if (!p || buf < 0 || buf >= static_cast<int>(p->bufs.size())) return 0;
  return p->bufs[buf].bytes;
}
int bd_plan_read_buffer(bd_plan* p, int buf, void* host_dst, size_t bytes) {
  BD_ON_PLAN(p);
  BD_CHECK(p && p->finalized && buf >= 0 && buf < static_cast<int>(p->bufs.size()) && bytes <= p->bufs[buf].bytes,
           "bad arguments");
  BD_CHECK(!p->bufs[buf].shared, "this buffer shares its arena range with others (arena reuse): build the plan with "
                                 "bd_plan_set_arena_reuse(plan, 0) to read intermediates");
  BD_CUDA(cudaDeviceSynchronize());
  BD_CUDA(cudaMemcpy(host_dst, p->arena + p->bufs[buf].offset, bytes, cudaMemcpyDeviceToHost));
  return 0;
}
int bd_plan_write_buffer(bd_plan* p, int buf, const void* host_src, size_t bytes) {
  BD_ON_PLAN(p);
  BD_CHECK(p && p->finalized && buf >= 0 && buf < static_cast<int>(p->bufs.size()) && bytes <= p->bufs[buf].bytes,
           "bad arguments");
  BD_CUDA(cudaMemcpy(p->arena + p->bufs[buf].offset, host_src, bytes, cudaMemcpyHostToDevice));
  return 0;
}
size_t bd_plan_arena_bytes(bd_plan* p) { return p ? p->arena_bytes : 0; }
size_t bd_plan_arena_bytes_flat(bd_plan* p) { return p ? p->arena_bytes_flat : 0; }
int bd_plan_buffer_lifetime(bd_plan* p, int buf, int* first, int* last, int* shared) {
  BD_CHECK(p && p->finalized && buf >= 0 && buf < static_cast<int>(p->bufs.size()), "bad arguments");
  if (first) *first = p->bufs[buf].first;
  if (last) *last = p->bufs[buf].last;
  if (shared) *shared = p->bufs[buf].shared ? 1 : 0;
  return 0;
}
int bd_plan_set_arena_reuse(bd_plan* p, int on) {
  BD_CHECK(p && !p->finalized, "bd_plan_set_arena_reuse: call before bd_plan_finalize");
  p->reuse = on != 0;
  return 0;
}
int bd_plan_num_launches(bd_plan* p) {
  int n = 0;
  if (p) for (const Op& op : p->ops) n += op.launches;
  return n;
}
int bd_plan_num_ops(bd_plan* p) { return p ? static_cast<int>(p->ops.size()) : 0; }
int bd_plan_op_info(bd_plan* p, int i, int* kind, double* flops) {
  BD_CHECK(p && i >= 0 && i < static_cast<int>(p->ops.size()), "bad op index");
  if (kind) *kind = p->ops[i].kclass;
  if (flops) *flops = p->ops[i].flops;
  return 0;
}
int bd_plan_time_ops(bd_plan* p, float* ms_out, void* stream) {
  BD_ON_PLAN(p);
  BD_CHECK(p && p->finalized && ms_out, "bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = p->ops.size();
  p->cur_probs = nullptr;  // the head kernel is skipped: its outputs belong to the caller of bd_plan_run
  p->cur_mask = nullptr;
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) BD_CUDA(cudaEventCreate(&e));
  BD_CUDA(cudaEventRecord(ev[0], s));
  const bool sync_each = getenv("BD_SYNC_EACH_OP") != nullptr;  // debugging aid: name the op that faults
  for (size_t i = 0; i < n; ++i) {
    if (p->ops[i].run(s)) return 1;
    BD_CUDA(cudaEventRecord(ev[i + 1], s));
    if (sync_each) {
      const cudaError_t e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) return fail("native op " + std::to_string(i) + " failed: " + cudaGetErrorString(e));
    }
  }
  BD_CUDA(cudaStreamSynchronize(s));
  for (size_t i = 0; i < n; ++i) BD_CUDA(cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]));
  for (auto& e : ev) cudaEventDestroy(e);
  return 0;
}

// ------------------------------------------------------------------------------------------ tiler / stitcher
static int ensure_tile_scratch(bd_ctx* ctx, int n) {
  if (n <= ctx->tile_cap) return 0;
  if (ctx->d_ys) cudaFree(ctx->d_ys);
  if (ctx->d_xs) cudaFree(ctx->d_xs);
  ctx->d_ys = ctx->d_xs = nullptr;
  ctx->tile_cap = 0;
  BD_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->d_ys), sizeof(int) * n));
  BD_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->d_xs), sizeof(int) * n));
  ctx->tile_cap = n;
  return 0;
}

int bd_tiles_gather(bd_ctx* ctx, const uint8_t* scene_bgr_dev, int h, int w, const int32_t* ys_host,
                    const int32_t* xs_host, int n, void* x_dev, int stem_stride, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && scene_bgr_dev && ys_host && xs_host && x_dev && n >= 1 && h >= 1 && w >= 1, "bad arguments");
  BD_CHECK(stem_stride == 1 || stem_stride == 2, "stem stride must be 1 or 2");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ensure_tile_scratch(ctx, std::max(n, 64))) return 1;
  BD_CUDA(cudaMemcpyAsync(ctx->d_ys, ys_host, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  BD_CUDA(cudaMemcpyAsync(ctx->d_xs, xs_host, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  ctx->launches++;
  const size_t work = static_cast<size_t>(n) * (512 / stem_stride) * (512 / stem_stride) * 4;
  if (stem_stride == 1)
    BD_LAUNCH(k::tiles_gather_kernel<1>, dim3(grid_for(work, ctx->num_sms * 4)), dim3(k::TPB), 0, s, scene_bgr_dev, h, w, ctx->d_ys, ctx->d_xs, n,
                                                                                static_cast<h16*>(x_dev));
  else
    BD_LAUNCH(k::tiles_gather_kernel<2>, dim3(grid_for(work, ctx->num_sms * 4)), dim3(k::TPB), 0, s, scene_bgr_dev, h, w, ctx->d_ys, ctx->d_xs, n,
                                                                                static_cast<h16*>(x_dev));
  BD_CUDA(cudaGetLastError());
  return 0;
}

int bd_stitch_or(bd_ctx* ctx, const uint8_t* tile_masks_dev, const int32_t* ys_host, const int32_t* xs_host, int n,
                 uint8_t* scene_mask_dev, int h, int w, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && tile_masks_dev && ys_host && xs_host && scene_mask_dev && n >= 1, "bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ensure_tile_scratch(ctx, std::max(n, 64))) return 1;
  BD_CUDA(cudaMemcpyAsync(ctx->d_ys, ys_host, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  BD_CUDA(cudaMemcpyAsync(ctx->d_xs, xs_host, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  ctx->launches++;
  BD_LAUNCH(k::stitch_or_kernel, dim3(grid_for(static_cast<size_t>(n) * 512 * 512, ctx->num_sms * 4)), dim3(k::TPB), 0, s, 
      tile_masks_dev, ctx->d_ys, ctx->d_xs, n, scene_mask_dev, h, w);
  BD_CUDA(cudaGetLastError());
  return 0;
}

// The origins of ALL tiles of a scene, uploaded once: the per-batch calls below index into them, so the scene loop
// issues no host->device copies between the kernels of consecutive batches.
int bd_tiles_set_origins(bd_ctx* ctx, const int32_t* ys_host, const int32_t* xs_host, int n, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && ys_host && xs_host && n >= 1, "bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n > ctx->origin_cap) {
    if (ctx->d_all_ys) cudaFree(ctx->d_all_ys);
    if (ctx->d_all_xs) cudaFree(ctx->d_all_xs);
    ctx->d_all_ys = ctx->d_all_xs = nullptr;
    ctx->origin_cap = 0;
    BD_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->d_all_ys), sizeof(int) * n));
    BD_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->d_all_xs), sizeof(int) * n));
    ctx->origin_cap = n;
  }
  BD_CUDA(cudaMemcpyAsync(ctx->d_all_ys, ys_host, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  BD_CUDA(cudaMemcpyAsync(ctx->d_all_xs, xs_host, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  BD_CUDA(cudaStreamSynchronize(s));  // the host arrays may be temporaries
  ctx->n_origins = n;
  return 0;
}

int bd_tiles_gather_at(bd_ctx* ctx, const uint8_t* scene_bgr_dev, int h, int w, int first, int n, void* x_dev,
                       int stem_stride, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && scene_bgr_dev && x_dev && n >= 1 && h >= 1 && w >= 1, "bad arguments");
  BD_CHECK(first >= 0 && first + n <= ctx->n_origins, "tile range outside the origins set by bd_tiles_set_origins");
  BD_CHECK(stem_stride == 1 || stem_stride == 2, "stem stride must be 1 or 2");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ctx->launches++;
  const size_t work = static_cast<size_t>(n) * (512 / stem_stride) * (512 / stem_stride) * 4;
  if (stem_stride == 1)
    BD_LAUNCH(k::tiles_gather_kernel<1>, dim3(grid_for(work, ctx->num_sms * 4)), dim3(k::TPB), 0, s, scene_bgr_dev, h, w,
              static_cast<const int*>(ctx->d_all_ys + first), static_cast<const int*>(ctx->d_all_xs + first), n, static_cast<h16*>(x_dev));
  else
    BD_LAUNCH(k::tiles_gather_kernel<2>, dim3(grid_for(work, ctx->num_sms * 4)), dim3(k::TPB), 0, s, scene_bgr_dev, h, w,
              static_cast<const int*>(ctx->d_all_ys + first), static_cast<const int*>(ctx->d_all_xs + first), n, static_cast<h16*>(x_dev));
  return 0;
}

int bd_stitch_or_at(bd_ctx* ctx, const uint8_t* tile_masks_dev, int first, int n, uint8_t* scene_mask_dev, int h, int w,
                    void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && tile_masks_dev && scene_mask_dev && n >= 1, "bad arguments");
  BD_CHECK(first >= 0 && first + n <= ctx->n_origins, "tile range outside the origins set by bd_tiles_set_origins");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ctx->launches++;
  BD_LAUNCH(k::stitch_or_kernel, dim3(grid_for(static_cast<size_t>(n) * 512 * 512, ctx->num_sms * 4)), dim3(k::TPB), 0, s,
            tile_masks_dev, static_cast<const int*>(ctx->d_all_ys + first), static_cast<const int*>(ctx->d_all_xs + first), n,
            scene_mask_dev, h, w);
  return 0;
}


// ------------------------------------------------------------------------------------------ whole-scene entry
// One forward of `p` writing the argmax masks to mask_dev, replayed from a CUDA graph when it can be captured (the
// ~140 kernels of a plan become one launch; BD_GRAPHS=0 turns it off).  Falls back to direct launches.
static int plan_run_masks(bd_plan* p, uint8_t* mask_dev, cudaStream_t s) {
  static const bool graphs_on = [] { const char* e = getenv("BD_GRAPHS"); return !(e && e[0] == '0'); }();
  if (graphs_on && !p->graph_failed && (!p->graph_exec || p->graph_mask != mask_dev)) {
    if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
    cudaGraph_t g = nullptr;
    // capture on an internal stream (the caller's may be the legacy default stream, which cannot capture); the
    // instantiated graph is launched on the caller's stream
    bd_ctx* ctx = p->ctx;
    bool ok = true;
    if (!ctx->capture_stream) ok = cudaStreamCreateWithFlags(&ctx->capture_stream, cudaStreamNonBlocking) == cudaSuccess;
    cudaStream_t cs = ctx->capture_stream;
    ok = ok && cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      p->cur_probs = nullptr;
      p->cur_mask = mask_dev;
      const int64_t launches0 = ctx->launches;
      int rc = 0;
      for (Op& op : p->ops)
        if ((rc = op.run(cs))) break;
      ctx->launches = launches0;  // capturing is not launching
      ok = cudaStreamEndCapture(cs, &g) == cudaSuccess && rc == 0 && g != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&p->graph_exec, g, 0) == cudaSuccess;
    if (g) cudaGraphDestroy(g);
    if (!ok) {
      cudaGetLastError();  // clear the sticky-free error of a refused capture
      p->graph_exec = nullptr;
      p->graph_failed = true;
    } else {
      p->graph_mask = mask_dev;
    }
  }
  if (p->graph_exec) {
    BD_CUDA(cudaGraphLaunch(p->graph_exec, s));
    p->ctx->launches += bd_plan_num_launches(p);
    return 0;
  }
  return bd_plan_run(p, nullptr, nullptr, mask_dev, s);
}

// The tiled forward of a scene, predict.py:98-114 for every model: the batch loop lives here, so a scene is ONE call
// from the host language (196 batches x 5 models x {gather, forward, stitch} at 20 000^2).
int bd_scene_run(bd_ctx* ctx, bd_plan* const* plans, int n_plans, const uint8_t* scene_bgr_dev, int h, int w,
                 const int32_t* ys_host, const int32_t* xs_host, int n_tiles, uint8_t* masks_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && plans && n_plans >= 1 && scene_bgr_dev && masks_dev && h >= 1 && w >= 1 && n_tiles >= 0, "bad arguments");
  if (n_tiles == 0) return 0;
  BD_CHECK(ys_host && xs_host, "bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = plans[0]->batch;
  for (int k = 0; k < n_plans; ++k) {
    BD_CHECK(plans[k] && plans[k]->finalized && plans[k]->ctx == ctx && plans[k]->batch == B && plans[k]->input_buf >= 0 &&
                 plans[k]->logits_buf >= 0,
             "bd_scene_run: plans must be finalized on this context with one common batch size");
  }
  const size_t tm_bytes = static_cast<size_t>(B) * 512 * 512;
  void* tm = nullptr;
  if (ctx->pool.get(bd::post::SLOT_TILEMASK, tm_bytes, &tm)) return 1;
  if (bd_tiles_set_origins(ctx, ys_host, xs_host, n_tiles, stream)) return 1;
  const size_t plane = static_cast<size_t>(h) * w;
  NvtxRange nvtx("bd:scene_forward");
  for (int b0 = 0; b0 < n_tiles; b0 += B) {
    const int n = std::min(B, n_tiles - b0);  // a ragged last batch leaves stale tiles in the other slots: ignored
    for (int k = 0; k < n_plans; ++k) {
      bd_plan* p = plans[k];
      NvtxRange nvtx_model(k == 0 ? "bd:model0" : k == 1 ? "bd:model1" : k == 2 ? "bd:model2" : k == 3 ? "bd:model3" : "bd:model4+");
      const BufInfo& ib = p->bufs[p->input_buf];
      if (bd_tiles_gather_at(ctx, scene_bgr_dev, h, w, b0, n, p->arena + ib.offset, ib.H == 512 ? 1 : 2, stream)) return 1;
      if (plan_run_masks(p, static_cast<uint8_t*>(tm), s)) return 1;
      if (bd_stitch_or_at(ctx, static_cast<const uint8_t*>(tm), b0, n, masks_dev + k * plane, h, w, stream)) return 1;
    }
  }
  BD_CUDA(cudaGetLastError());
  return 0;
}

// ---- plan files: the lowered network (buffers, fused ops, BN-folded fp16 weights) as the sequence of C-ABI calls that
// built it.  Written once by the Python builder (tools/export_plans.py), loaded by any host language.
static const char PLAN_MAGIC[8] = {'B', 'D', 'P', 'L', 'A', 'N', '0', '1'};
int bd_plan_save(bd_plan* p, const char* path) {
  BD_CHECK(p && p->finalized && path, "bd_plan_save: a finalized plan and a path");
  FILE* f = fopen(path, "wb");
  BD_CHECK(f != nullptr, "bd_plan_save: cannot open the file for writing");
  const int32_t batch = p->batch;
  const uint64_t n = p->log.size();
  const bool ok = fwrite(PLAN_MAGIC, 1, 8, f) == 8 && fwrite(&batch, 4, 1, f) == 1 && fwrite(&n, 8, 1, f) == 1 &&
                  fwrite(p->log.data(), 1, n, f) == n;
  fclose(f);
  BD_CHECK(ok, "bd_plan_save: short write");
  return 0;
}

int bd_plan_load(bd_ctx* ctx, const char* path, bd_plan** out) {
  BD_CHECK(ctx && path && out, "bad arguments");
  FILE* f = fopen(path, "rb");
  BD_CHECK(f != nullptr, "bd_plan_load: cannot open the file");
  char magic[8];
  int32_t batch = 0;
  uint64_t n = 0;
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, PLAN_MAGIC, 8) == 0 && fread(&batch, 4, 1, f) == 1 && fread(&n, 8, 1, f) == 1 &&
            n < (1ull << 34);
  std::vector<uint8_t> log;
  if (ok) { log.resize(n); ok = fread(log.data(), 1, n, f) == n; }
  fclose(f);
  BD_CHECK(ok, "bd_plan_load: not a plan file (or truncated)");
  bd_plan* p = nullptr;
  if (bd_plan_create(ctx, batch, &p)) return 1;
  LogReader r{log.data(), log.data() + log.size()};
  int rc = 0;
  bool done = false;
  while (!rc && r.ok && r.p < r.end && !done) {
    const uint32_t op = r.pod<uint32_t>();
    size_t nb = 0;
    switch (op) {
      case LOG_BUFFER: { const int h = r.pod<int>(), w = r.pod<int>(), c = r.pod<int>(), dt = r.pod<int>(), kd = r.pod<int>();
        rc = bd_plan_add_buffer(p, h, w, c, dt, kd) < 0; break; }
      case LOG_CONV: { bd_conv_desc d = r.pod<bd_conv_desc>();
        d.w_host = static_cast<const uint16_t*>(r.arr(&nb)); d.bias_host = static_cast<const float*>(r.arr(&nb));
        d.dw_w_host = r.pod<int32_t>() ? static_cast<const float*>(r.arr(&nb)) : nullptr;
        if (r.ok) rc = bd_plan_add_conv(p, &d); break; }
      case LOG_DWCONV: { const bd_tref x = r.pod<bd_tref>(), y = r.pod<bd_tref>(); const int st = r.pod<int>(), pt = r.pod<int>(), pl = r.pod<int>(), ri = r.pod<int>();
        const float* w = static_cast<const float*>(r.arr(&nb)); if (r.ok) rc = bd_plan_add_dwconv(p, x, y, st, pt, pl, ri, w); break; }
      case LOG_MAXPOOL: { const bd_tref x = r.pod<bd_tref>(), y = r.pod<bd_tref>(); const int kk = r.pod<int>(), st = r.pod<int>(), pt = r.pod<int>(), pl = r.pod<int>();
        rc = bd_plan_add_maxpool(p, x, y, kk, st, pt, pl); break; }
      case LOG_ADDN: { const int nn = r.pod<int>(); bd_tref xs[4]; int32_t fs[4];
        if (nn < 1 || nn > 4) { r.ok = false; break; }
        for (int i = 0; i < nn; ++i) { xs[i] = r.pod<bd_tref>(); fs[i] = r.pod<int32_t>(); }
        const bd_tref y = r.pod<bd_tref>(); const int act = r.pod<int>(); rc = bd_plan_add_addn(p, nn, xs, fs, y, act); break; }
      case LOG_GAP: { const bd_tref x = r.pod<bd_tref>(); const int yv = r.pod<int>(); rc = bd_plan_add_gap(p, x, yv); break; }
      case LOG_DENSE: { const int ni = r.pod<int>(); int32_t xv[5];
        if (ni < 1 || ni > 5) { r.ok = false; break; }
        for (int i = 0; i < ni; ++i) xv[i] = r.pod<int32_t>();
        const int yv = r.pod<int>(), cin = r.pod<int>(), cout = r.pod<int>(), act = r.pod<int>();
        const float* w = static_cast<const float*>(r.arr(&nb)); const float* b = static_cast<const float*>(r.arr(&nb));
        if (r.ok) rc = bd_plan_add_dense(p, ni, xv, yv, cin, cout, act, w, b); break; }
      case LOG_GATE: { const int mode = r.pod<int>(); const bd_tref x = r.pod<bd_tref>(), y = r.pod<bd_tref>(); const int vv = r.pod<int>();
        const bd_tref sr = r.pod<bd_tref>(); const float bs = r.pod<float>(); const float* w = r.pod<int32_t>() ? static_cast<const float*>(r.arr(&nb)) : nullptr;
        if (r.ok) rc = bd_plan_add_gate(p, mode, x, y, vv, sr, w, bs); break; }
      case LOG_SKFUSE: { bd_tref xs[4]; int32_t lv[5]; for (int i = 0; i < 4; ++i) xs[i] = r.pod<bd_tref>(); const int gv = r.pod<int>();
        for (int i = 0; i < 5; ++i) lv[i] = r.pod<int32_t>(); const bd_tref y = r.pod<bd_tref>();
        const float* sc = static_cast<const float*>(r.arr(&nb)); const float* sh = static_cast<const float*>(r.arr(&nb));
        if (r.ok) rc = bd_plan_add_skfuse(p, xs, gv, lv, y, sc, sh); break; }
      case LOG_BCAST: { const int vv = r.pod<int>(); const bd_tref y = r.pod<bd_tref>(); rc = bd_plan_add_bcast(p, vv, y); break; }
      case LOG_FINALIZE: { const int ib = r.pod<int>(), lb = r.pod<int>(), up = r.pod<int>(); rc = bd_plan_finalize(p, ib, lb, up); done = true; break; }
      default: r.ok = false;
    }
  }
  if (rc || !r.ok || !done) {
    const std::string why = rc ? bd::last_error() : std::string("bd_plan_load: corrupt plan file");
    bd_plan_destroy(p);
    return fail(why);
  }
  *out = p;
  return 0;
}

// stride of the 3x3 stem the plan's input buffer is laid out for (1 or 2): what bd_tiles_gather needs
int bd_plan_input_stride(bd_plan* p) {
  if (!p || p->input_buf < 0) return 0;
  return p->bufs[p->input_buf].H == 512 ? 1 : 2;
}

// 1 when the plan's forward is replayed from a captured CUDA graph (bd_scene_run), 0 when it is launched kernel by kernel
int bd_plan_uses_graph(bd_plan* p) { return (p && p->graph_exec) ? 1 : 0; }

// Opt-in fusion by probability averaging (the reference votes on argmax masks): per tile pixel
// the mean over the models of P(building) > 0.5, tiles OR-stitched into ONE scene mask.  The caller applies the
// reference's final clean-up (bd_mask_cleanup) and bd_contours to it.
int bd_scene_run_average(bd_ctx* ctx, bd_plan* const* plans, int n_plans, const uint8_t* scene_bgr_dev, int h, int w,
                         const int32_t* ys_host, const int32_t* xs_host, int n_tiles, uint8_t* mask_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && plans && n_plans >= 1 && scene_bgr_dev && mask_dev && h >= 1 && w >= 1 && n_tiles >= 0, "bad arguments");
  if (n_tiles == 0) return 0;
  BD_CHECK(ys_host && xs_host, "bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = plans[0]->batch;
  for (int k = 0; k < n_plans; ++k)
    BD_CHECK(plans[k] && plans[k]->finalized && plans[k]->ctx == ctx && plans[k]->batch == B && plans[k]->input_buf >= 0 &&
                 plans[k]->logits_buf >= 0, "bd_scene_run_average: plans must be finalized on this context with one common batch size");
  const size_t npx = static_cast<size_t>(B) * 512 * 512;
  void *tm = nullptr, *probs = nullptr, *acc = nullptr;
  if (ctx->pool.get(bd::post::SLOT_TILEMASK, npx, &tm) || ctx->pool.get(bd::post::SLOT_PROBS, npx * 2 * sizeof(float), &probs) ||
      ctx->pool.get(bd::post::SLOT_PROBACC, npx * sizeof(float), &acc))
    return 1;
  if (bd_tiles_set_origins(ctx, ys_host, xs_host, n_tiles, stream)) return 1;
  NvtxRange nvtx("bd:scene_forward_average");
  for (int b0 = 0; b0 < n_tiles; b0 += B) {
    const int n = std::min(B, n_tiles - b0);
    for (int k = 0; k < n_plans; ++k) {
      bd_plan* p = plans[k];
      const BufInfo& ib = p->bufs[p->input_buf];
      if (bd_tiles_gather_at(ctx, scene_bgr_dev, h, w, b0, n, p->arena + ib.offset, ib.H == 512 ? 1 : 2, stream)) return 1;
      if (bd_plan_run(p, nullptr, static_cast<float*>(probs), nullptr, s)) return 1;
      ctx->launches++;
      BD_LAUNCH(k::prob_accum_kernel, dim3(grid_for(npx, ctx->num_sms * 4)), dim3(k::TPB), 0, s, static_cast<const float*>(probs),
                static_cast<float*>(acc), npx, k == 0 ? 1 : 0, 0.5f * n_plans, k + 1 == n_plans ? static_cast<uint8_t*>(tm) : nullptr);
    }
    if (bd_stitch_or_at(ctx, static_cast<const uint8_t*>(tm), b0, n, mask_dev, h, w, stream)) return 1;
  }
  BD_CUDA(cudaGetLastError());
  return 0;
}

// device bytes a context holds for the scene-level stages at (h, w): fusion / contour arena + pools (weights and
// activation arenas belong to the plans: bd_plan_arena_bytes)
size_t bd_workspace_bytes(bd_ctx* ctx) {
  if (!ctx) return 0;
  size_t b = ctx->arena.cap;
  for (int i = 0; i < bd::post::DevPool::SLOTS; ++i) b += ctx->pool.cap[i];
  return b;
}

}  // extern "C"
