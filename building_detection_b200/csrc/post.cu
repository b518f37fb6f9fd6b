// post.cu -- mask fusion (reference model_fuse.py) on the GPU; contour extraction lives in contours.cu.
//
// model_fuse.py processes every object on its own full-frame image (O(#objects x H x W)); here each step is one
// pass over the scene with per-component results keyed by the component's root pixel:
//   clean-up(mask) = fill holes -> label -> drop polygon area <= 1000 -> 1x21 and 21x1 erosion -> label the
//   fragments, drop polygon area <= 500 -> per-object decision (keep / drop / replace by its fragments dilated
//   back) -> rasterise;   fuse = clean-up x5 -> vote >= 3 -> clean-up.
// Different objects are never 8-adjacent, so the per-object erosions / dilations of the reference equal one
// global erosion / a dilation restricted to the fragments' own object.
#include "../../include/bd_b200.h"
#include "ccl.cuh"
#include "post_ws.cuh"

using namespace bd;

namespace bd {
namespace post {

constexpr int TPB = ccl::TPB;

// For every fragment root f: parent = L[f]; cnt[parent]++; surv[parent]++ when the fragment's polygon area > 500.
static __global__ void __launch_bounds__(TPB) count_fragments(const int* __restrict__ Lf, const long long* __restrict__ area2f,
                                                       const int* __restrict__ L, int* cnt, int* surv, size_t n,
                                                       long long thr2) {
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB)
    if (Lf[i] == static_cast<int>(i)) {
      const int parent = L[i];
      atomicAdd(cnt + parent, 1);
      if (llabs(area2f[i]) > thr2) atomicAdd(surv + parent, 1);
    }
}
// eroede_dilate_process decision + only_plt rasterisation (model_fuse.py:173-218, 265-268) per pixel of a kept object
static __global__ void __launch_bounds__(TPB) rasterise_objects(const uint8_t* __restrict__ keep, const int* __restrict__ L,
                                                         const int* __restrict__ Lh, const int* __restrict__ Lv,
                                                         const long long* __restrict__ a2h, const long long* __restrict__ a2v,
                                                         const int* __restrict__ cntH, const int* __restrict__ survH,
                                                         const int* __restrict__ cntV, const int* __restrict__ survV,
                                                         uint8_t* __restrict__ out, int H, int W, int half, long long thr2) {
  const size_t n = static_cast<size_t>(H) * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB) {
    uint8_t v = 0;
    if (keep[i]) {
      const int r = L[i];
      const int ch = cntH[r], sh = survH[r], cv_ = cntV[r], sv = survV[r];
      // erode_process returns False when fragments were erased and none is left (:81-83)
      const bool falseH = (ch != 1) && (sh < ch) && (sh == 0);
      const bool falseV = (cv_ != 1) && (sv < cv_) && (sv == 0);
      if (falseH || falseV) v = 0;
      else if (ch == 1 && cv_ == 1) v = 255;
      else {
        const int x = static_cast<int>(i % W), y = static_cast<int>(i / W);
        if (ch != 1) {  // pieces of the horizontal split: surviving fragments dilated back by 1x21
          const int x0 = max(0, x - half), x1 = min(W - 1, x + half);
          const size_t row = static_cast<size_t>(y) * W;
          for (int xx = x0; xx <= x1 && !v; ++xx) {
            const int f = Lh[row + xx];
            if (f >= 0 && llabs(a2h[f]) > thr2) v = 255;
          }
        }
        if (cv_ != 1 && !v) {
          const int y0 = max(0, y - half), y1 = min(H - 1, y + half);
          for (int yy = y0; yy <= y1 && !v; ++yy) {
            const int f = Lv[static_cast<size_t>(yy) * W + x];
            if (f >= 0 && llabs(a2v[f]) > thr2) v = 255;
          }
        }
      }
    }
    out[i] = v;
  }
}
static __global__ void __launch_bounds__(TPB) vote3of5(const uint8_t* __restrict__ m, size_t n, uint8_t* __restrict__ out) {
  // 16 pixels per thread (128-bit loads) when the five planes stay 16-byte aligned; l_k // 255 summed, >= 3 -> 255
  // (model_fuse.py:315-324)
  const size_t nv = (n % 16 == 0) ? n / 16 : 0;
  for (size_t t = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; t < nv; t += static_cast<size_t>(gridDim.x) * TPB) {
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const uint4 v = reinterpret_cast<const uint4*>(m + k * n)[t];
      // bytes are 0 or 255: (v >> 7) & 0x01010101 gives 0/1 per byte
      acc.x += (v.x >> 7) & 0x01010101u; acc.y += (v.y >> 7) & 0x01010101u;
      acc.z += (v.z >> 7) & 0x01010101u; acc.w += (v.w >> 7) & 0x01010101u;
    }
    auto thr = [](unsigned a) {  // per byte: a >= 3 -> 0xFF  (a in 0..5: add 5, bit 3 set iff a >= 3)
      const unsigned b = ((a + 0x05050505u) >> 3) & 0x01010101u;
      return b * 255u;
    };
    reinterpret_cast<uint4*>(out)[t] = make_uint4(thr(acc.x), thr(acc.y), thr(acc.z), thr(acc.w));
  }
  for (size_t i = nv * 16 + blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB) {
    int s = 0;
    for (int k = 0; k < 5; ++k) s += m[k * n + i] / 255;
    out[i] = s >= 3 ? 255 : 0;
  }
}

static inline int grid_for(size_t n, int sms) {
  return static_cast<int>(std::max<size_t>(1, std::min<size_t>((n + TPB - 1) / TPB, static_cast<size_t>(sms) * 16)));
}

// labels of the set pixels of m (8-connectivity) into L
int label8(bd_ctx* ctx, const uint8_t* m, int* L, int H, int W, cudaStream_t s) {
  const size_t n = static_cast<size_t>(H) * W;
  const int g = grid_for(n, ctx_sms(ctx));
  ccl::init_runs<<<grid_for(static_cast<size_t>(H) * 32, ctx_sms(ctx)), TPB, 0, s>>>(m, L, H, W, 1);
  ccl::merge_runs<true><<<g, TPB, 0, s>>>(L, H, W);
  ccl::compress_runs<<<g, TPB, 0, s>>>(L, n, W);
  ccl::flatten_runs<<<g, TPB, 0, s>>>(L, n, W);
  ccl::flatten_pixels<<<g, TPB, 0, s>>>(L, n, W);
  ctx_count(ctx, 5);
  BD_CUDA(cudaGetLastError());
  return 0;
}

// hole fill: out = m with every enclosed region set (cv::fillPoly of all external contours)
int fill(bd_ctx* ctx, const uint8_t* m, int* Lbg, uint8_t* out, int H, int W, cudaStream_t s) {
  const size_t n = static_cast<size_t>(H) * W;
  const int g = grid_for(n, ctx_sms(ctx));
  ccl::init_runs<<<grid_for(static_cast<size_t>(H) * 32, ctx_sms(ctx)), TPB, 0, s>>>(m, Lbg, H, W, 0);
  ccl::merge_runs<false><<<g, TPB, 0, s>>>(Lbg, H, W);
  ccl::compress_runs<<<g, TPB, 0, s>>>(Lbg, n, W);
  ccl::flatten_runs<<<g, TPB, 0, s>>>(Lbg, n, W);
  ccl::flatten_pixels<<<g, TPB, 0, s>>>(Lbg, n, W);
  ccl::mark_outside<<<grid_for(2 * (static_cast<size_t>(H) + W), ctx_sms(ctx)), TPB, 0, s>>>(Lbg, H, W);
  ccl::fill_holes<<<g, TPB, 0, s>>>(m, Lbg, out, n);
  ctx_count(ctx, 7);
  BD_CUDA(cudaGetLastError());
  return 0;
}

// One clean-up pass (fill_and_delete + eroede_dilate_process + only_plt, model_fuse.py:9-218,265-268).
int cleanup(bd_ctx* ctx, const uint8_t* mask, int H, int W, uint8_t* out, cudaStream_t s) {
  Workspace* ws = nullptr;
  if (workspace(ctx, H, W, &ws)) return 1;
  const size_t n = static_cast<size_t>(H) * W;
  const int g = grid_for(n, ctx_sms(ctx));
  const int gv = grid_for(static_cast<size_t>(H + 1) * (W + 1), ctx_sms(ctx));
  // fill_and_delete: fill, label, polygon area, drop <= 1000
  if (fill(ctx, mask, ws->Lh, ws->filled, H, W, s)) return 1;
  if (label8(ctx, ws->filled, ws->L, H, W, s)) return 1;
  ccl::zero_at_roots<<<g, TPB, 0, s>>>(ws->L, n, ws->a2, ws->cntH, ws->survH, ws->cntV, ws->survV);
  ccl::polygon_area2<<<gv, TPB, 0, s>>>(ws->L, H, W, ws->a2);
  ccl::drop_small<<<g, TPB, 0, s>>>(ws->L, ws->a2, 2 * 1000, ws->keep, n, 0);
  // eroede_dilate_process: 1x5 / 5x1 kernels, 5 iterations == one 1x21 / 21x1 erosion
  ccl::erode_line<<<g, TPB, 0, s>>>(ws->keep, ws->er, H, W, 10, 0);
  ctx_count(ctx, 4);
  if (label8(ctx, ws->er, ws->Lh, H, W, s)) return 1;
  ccl::zero_at_roots<<<g, TPB, 0, s>>>(ws->Lh, n, ws->a2h, nullptr, nullptr, nullptr, nullptr);
  ccl::polygon_area2<<<gv, TPB, 0, s>>>(ws->Lh, H, W, ws->a2h);
  count_fragments<<<g, TPB, 0, s>>>(ws->Lh, ws->a2h, ws->L, ws->cntH, ws->survH, n, 2 * 500);
  ccl::erode_line<<<g, TPB, 0, s>>>(ws->keep, ws->er, H, W, 10, 1);
  ctx_count(ctx, 4);
  if (label8(ctx, ws->er, ws->Lv, H, W, s)) return 1;
  ccl::zero_at_roots<<<g, TPB, 0, s>>>(ws->Lv, n, ws->a2v, nullptr, nullptr, nullptr, nullptr);
  ccl::polygon_area2<<<gv, TPB, 0, s>>>(ws->Lv, H, W, ws->a2v);
  count_fragments<<<g, TPB, 0, s>>>(ws->Lv, ws->a2v, ws->L, ws->cntV, ws->survV, n, 2 * 500);
  rasterise_objects<<<g, TPB, 0, s>>>(ws->keep, ws->L, ws->Lh, ws->Lv, ws->a2h, ws->a2v, ws->cntH, ws->survH, ws->cntV,
                                      ws->survV, out, H, W, 10, 2 * 500);
  ctx_count(ctx, 4);
  BD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace post
}  // namespace bd

extern "C" {

int bd_mask_cleanup(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, uint8_t* out_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && mask_dev && out_dev && h >= 1 && w >= 1, "bad arguments");
  BD_CHECK(static_cast<size_t>(h) * w < (1ull << 31), "scene too large for int32 pixel labels");
  return post::cleanup(ctx, mask_dev, h, w, out_dev, static_cast<cudaStream_t>(stream));
}

int bd_fuse(bd_ctx* ctx, const uint8_t* masks5_dev, int h, int w, uint8_t* fused_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && masks5_dev && fused_dev && h >= 1 && w >= 1, "bad arguments");
  BD_CHECK(static_cast<size_t>(h) * w < (1ull << 31), "scene too large for int32 pixel labels");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  post::Workspace* ws = nullptr;
  if (post::workspace(ctx, h, w, &ws)) return 1;
  const size_t n = static_cast<size_t>(h) * w;
  for (int k = 0; k < 5; ++k)
    if (post::cleanup(ctx, masks5_dev + k * n, h, w, ws->cleaned + k * n, s)) return 1;
  return bd_fuse_cleaned(ctx, ws->cleaned, h, w, fused_dev, stream);
}

int bd_fuse_cleaned(bd_ctx* ctx, const uint8_t* cleaned5_dev, int h, int w, uint8_t* fused_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && cleaned5_dev && fused_dev && h >= 1 && w >= 1, "bad arguments");
  BD_CHECK(static_cast<size_t>(h) * w < (1ull << 31), "scene too large for int32 pixel labels");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  post::Workspace* ws = nullptr;
  if (post::workspace(ctx, h, w, &ws)) return 1;
  const size_t n = static_cast<size_t>(h) * w;
  post::vote3of5<<<post::grid_for(n / 16 + 1, post::ctx_sms(ctx)), post::TPB, 0, s>>>(cleaned5_dev, n, ws->voted);
  post::ctx_count(ctx, 1);
  BD_CUDA(cudaGetLastError());
  return post::cleanup(ctx, ws->voted, h, w, fused_dev, s);
}

}  // extern "C"
