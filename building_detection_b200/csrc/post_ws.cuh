// post_ws.cuh -- bd_ctx (shared by bd_api.cu / post.cu / contours.cu) and the scene-sized workspace of the
// fusion and contour stages.  The workspace grows on demand and is reused across calls.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace bd {
namespace post {
struct Workspace {
  size_t cap = 0;            // pixels
  int *L = nullptr, *Lh = nullptr, *Lv = nullptr;              // parent / fragment labels
  long long *a2 = nullptr, *a2h = nullptr, *a2v = nullptr;    // 2 x signed polygon area, indexed by root pixel
  int *cntH = nullptr, *survH = nullptr, *cntV = nullptr, *survV = nullptr;
  uint8_t *filled = nullptr, *keep = nullptr, *er = nullptr, *voted = nullptr, *cleaned = nullptr;  // cleaned: 5 masks
  void release() {
    void* ptrs[] = {L, Lh, Lv, a2, a2h, a2v, cntH, survH, cntV, survV, filled, keep, er, voted, cleaned};
    for (void* p : ptrs)
      if (p) cudaFree(p);
    *this = Workspace();
  }
};
}  // namespace post
}  // namespace bd

namespace bd {
namespace post {
// grow-only device scratch slots for the contour stage (no cudaMalloc / cudaFree per call once warmed up)
struct DevPool {
  static constexpr int SLOTS = 32;
  void* ptr[SLOTS] = {};
  size_t cap[SLOTS] = {};
  int get(int slot, size_t bytes, void** out) {
    if (bytes > cap[slot]) {
      if (ptr[slot]) cudaFree(ptr[slot]);
      ptr[slot] = nullptr; cap[slot] = 0;
      const size_t want = bytes + bytes / 4 + 256;
      BD_CUDA(cudaMalloc(&ptr[slot], want));
      cap[slot] = want;
    }
    *out = ptr[slot];
    return 0;
  }
  void release() {
    for (int i = 0; i < SLOTS; ++i) { if (ptr[i]) cudaFree(ptr[i]); ptr[i] = nullptr; cap[i] = 0; }
  }
};
}  // namespace post
}  // namespace bd

struct bd_ctx {
  int device = 0;
  int num_sms = 148;
  int64_t launches = 0;
  int umma_smem_kb = 226;   // per-CTA smem budget of the persistent tcgen05 conv (1 CTA / SM)
  int umma_max_block_n = 256;
  int umma_group = 0;       // k-blocks per smem stage: 0 = automatic, 1 = off (BD_UMMA_GROUP)
  int* d_ys = nullptr;      // tile origin scratch
  int* d_xs = nullptr;
  int tile_cap = 0;
  int* d_all_ys = nullptr;  // origins of all tiles of the current scene (bd_tiles_set_origins)
  int* d_all_xs = nullptr;
  int origin_cap = 0, n_origins = 0;
  bd::post::Workspace post_ws;
  bd::post::DevPool pool;
  void* trace_buf = nullptr;  // BD_UMMA_TRACE debug buffer of the most recently built conv
};

namespace bd {
namespace post {
// defined in post.cu
int label8(bd_ctx* ctx, const uint8_t* m, int* L, int H, int W, cudaStream_t s);                 // 8-connected labels of m
int fill(bd_ctx* ctx, const uint8_t* m, int* Lbg, uint8_t* out, int H, int W, cudaStream_t s);  // hole fill
inline int ctx_sms(bd_ctx* c) { return c->num_sms; }
inline void ctx_count(bd_ctx* c, int n) { c->launches += n; }

inline int workspace(bd_ctx* ctx, int H, int W, Workspace** out) {
  Workspace& w = ctx->post_ws;
  const size_t n = static_cast<size_t>(H) * W;
  if (n > w.cap) {
    w.release();
    const size_t cap = n + 64;
#define BD_WS_ALLOC(field, type, count) BD_CUDA(cudaMalloc(reinterpret_cast<void**>(&w.field), sizeof(type) * (count)))
    BD_WS_ALLOC(L, int, cap); BD_WS_ALLOC(Lh, int, cap); BD_WS_ALLOC(Lv, int, cap);
    BD_WS_ALLOC(a2, long long, cap); BD_WS_ALLOC(a2h, long long, cap); BD_WS_ALLOC(a2v, long long, cap);
    BD_WS_ALLOC(cntH, int, cap); BD_WS_ALLOC(survH, int, cap); BD_WS_ALLOC(cntV, int, cap); BD_WS_ALLOC(survV, int, cap);
    BD_WS_ALLOC(filled, uint8_t, cap); BD_WS_ALLOC(keep, uint8_t, cap); BD_WS_ALLOC(er, uint8_t, cap);
    BD_WS_ALLOC(voted, uint8_t, cap); BD_WS_ALLOC(cleaned, uint8_t, 5 * cap);
#undef BD_WS_ALLOC
    w.cap = n;
  }
  *out = &w;
  return 0;
}
}  // namespace post
}  // namespace bd
