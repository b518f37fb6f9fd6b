// post.cu -- mask fusion (reference model_fuse.py) on the GPU, on bit-packed planes and run-based labels (rle.cuh);
// contour extraction lives in contours.cu.
//
// model_fuse.py processes every object on its own full-frame image (O(#objects x H x W)); here each step is one pass
// over a 1-bit-per-pixel plane with per-component results keyed by the component's root run:
//   clean-up(mask) = fill holes (4-connected labelling of the background) -> label (8-connected) -> drop polygon
//   area <= 1000 -> 1x21 and 21x1 erosion -> label the fragments, drop polygon area <= 500 -> per-object decision
//   (keep / drop / replace by its fragments dilated back) -> rasterise;   fuse = clean-up x5 -> vote >= 3 -> clean-up.
// Different objects are never 8-adjacent, so the per-object erosions / dilations of the reference equal one global
// erosion / dilation; the opening of a hole-free object by a line segment is a subset of the object and cannot
// enclose a hole, so `drawContours(FILLED)` of a dilated fragment is the dilated fragment.
#include "../../include/bd_b200.h"
#include "post_ws.cuh"
#include "rle.cuh"

using namespace bd;

namespace bd {
namespace post {

using rle::Plane;
using rle::RunSet;
constexpr int TPB = rle::TPB;

int grid_words(bd_ctx* ctx, size_t words) {
  return static_cast<int>(std::max<size_t>(1, std::min<size_t>((words + TPB - 1) / TPB, static_cast<size_t>(ctx->num_sms) * 16)));
}
static int grid_rows(bd_ctx* ctx, int H) {
  return std::max(1, std::min((H + TPB / 32 - 1) / (TPB / 32), ctx->num_sms * 8));
}

Plane take_plane(Arena& a, int H, int W) {
  Plane p;
  p.H = H; p.W = W; p.wp = rle::words_per_row(W);
  p.w = a.take<uint32_t>(static_cast<size_t>(H) * p.wp);
  return p;
}

int pack(bd_ctx* ctx, const uint8_t* src, Plane p, cudaStream_t s) {
  const size_t words = static_cast<size_t>(p.H) * p.wp;
  const bool vec = (p.W % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0);
  if (vec) rle::pack_u8<true><<<grid_words(ctx, words), TPB, 0, s>>>(src, p);
  else rle::pack_u8<false><<<grid_words(ctx, words), TPB, 0, s>>>(src, p);
  ctx->launches++;
  BD_CUDA(cudaGetLastError());
  return 0;
}
int unpack(bd_ctx* ctx, Plane p, uint8_t* dst, cudaStream_t s) {
  const size_t words = static_cast<size_t>(p.H) * p.wp;
  const bool vec = (p.W % 16 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  if (vec) rle::unpack_u8<true><<<grid_words(ctx, words), TPB, 0, s>>>(p, dst);
  else rle::unpack_u8<false><<<grid_words(ctx, words), TPB, 0, s>>>(p, dst);
  ctx->launches++;
  BD_CUDA(cudaGetLastError());
  return 0;
}

// Number the runs of plane p and label its components (8- or 4-connected).  `slot`: DevPool slot of the parent array.
// One host synchronisation (the run count sizes the per-run arrays).
int build_runs(bd_ctx* ctx, Arena& a, Plane p, bool conn8, int slot, RunSet* out, cudaStream_t s) {
  RunSet r;
  r.p = p;
  r.wprefix = a.take<uint32_t>(static_cast<size_t>(p.H) * p.wp);
  int* rows = a.take<int>(static_cast<size_t>(p.H) + 1);
  int* d_total = rows + p.H;
  rle::count_rows<<<grid_rows(ctx, p.H), TPB, 0, s>>>(p, rows);
  rle::scan_rows<<<1, 1024, 0, s>>>(rows, p.H, d_total);
  BD_CUDA(cudaMemcpyAsync(ctx->h_scalar, d_total, sizeof(int), cudaMemcpyDeviceToHost, s));
  BD_CUDA(cudaStreamSynchronize(s));
  r.nruns = *ctx->h_scalar;
  BD_CHECK(r.nruns >= 0, "run count overflow");
  void* P = nullptr;
  if (ctx->pool.get(slot, sizeof(int) * (static_cast<size_t>(r.nruns) + 1), &P)) return 1;
  r.P = static_cast<int*>(P);
  rle::emit_prefix<<<grid_rows(ctx, p.H), TPB, 0, s>>>(p, rows, r.wprefix, r.P);
  ctx->launches += 3;
  if (r.nruns > 0) {
    const size_t words = static_cast<size_t>(p.H) * p.wp;
    if (p.H > 1) {
      if (conn8) rle::merge_rows<true><<<grid_words(ctx, words), TPB, 0, s>>>(r);
      else rle::merge_rows<false><<<grid_words(ctx, words), TPB, 0, s>>>(r);
    }
    const int g = grid_words(ctx, static_cast<size_t>(r.nruns));
    rle::compress_runs<<<g, TPB, 0, s>>>(r.P, r.nruns);
    rle::flatten_runs<<<g, TPB, 0, s>>>(r.P, r.nruns);
    ctx->launches += 3;
  }
  BD_CUDA(cudaGetLastError());
  *out = r;
  return 0;
}

// hole fill: filled = fg | every background region that does not reach the frame.  Slots slot, slot+1.
int fill(bd_ctx* ctx, Arena& a, Plane fg, Plane filled, int slot, cudaStream_t s) {
  const size_t words = static_cast<size_t>(fg.H) * fg.wp;
  Plane bgp = take_plane(a, fg.H, fg.W);
  rle::complement<<<grid_words(ctx, words), TPB, 0, s>>>(fg, bgp);
  ctx->launches++;
  RunSet bg;
  if (build_runs(ctx, a, bgp, false, slot, &bg, s)) return 1;
  void* outside = nullptr;
  if (ctx->pool.get(slot + 1, static_cast<size_t>(bg.nruns) + 1, &outside)) return 1;
  BD_CUDA(cudaMemsetAsync(outside, 0, static_cast<size_t>(bg.nruns) + 1, s));
  rle::mark_outside<<<grid_words(ctx, words), TPB, 0, s>>>(bg, static_cast<uint8_t*>(outside));
  rle::fill_holes<<<grid_words(ctx, words), TPB, 0, s>>>(fg, bg, static_cast<const uint8_t*>(outside), filled);
  ctx->launches += 2;
  BD_CUDA(cudaGetLastError());
  return 0;
}

// label `p` (8-connected) and compute every component's polygon area; slots slot (parents), slot+1 (areas)
int label_area(bd_ctx* ctx, Arena& a, Plane p, int slot, RunSet* rs, long long** area2, cudaStream_t s) {
  if (build_runs(ctx, a, p, true, slot, rs, s)) return 1;
  void* ar = nullptr;
  if (ctx->pool.get(slot + 1, sizeof(long long) * (static_cast<size_t>(rs->nruns) + 1), &ar)) return 1;
  BD_CUDA(cudaMemsetAsync(ar, 0, sizeof(long long) * (static_cast<size_t>(rs->nruns) + 1), s));
  *area2 = static_cast<long long*>(ar);
  if (rs->nruns > 0) {
    rle::polygon_area2<<<grid_words(ctx, static_cast<size_t>(p.H + 1) * (p.wp + 1)), TPB, 0, s>>>(*rs, *area2);
    ctx->launches++;
  }
  BD_CUDA(cudaGetLastError());
  return 0;
}

// One clean-up pass (fill_and_delete + eroede_dilate_process + only_plt, model_fuse.py:9-218,265-268) on planes.
// stage_out / stage: debugging hook (bd_debug_cleanup_stage) -- copies an intermediate plane out.
int cleanup_plane(bd_ctx* ctx, Arena& a, Plane in, Plane out, cudaStream_t s, int stage, Plane* stage_out) {
  NvtxRange nvtx("bd:cleanup");
  const PostConstants& K = ctx->consts;
  const int H = in.H, W = in.W;
  const size_t words = static_cast<size_t>(H) * in.wp;
  const int g = grid_words(ctx, words);
  const size_t mark = a.off;  // scratch below is released on return
  auto emit_stage = [&](int id, Plane p) {
    if (stage == id && stage_out) cudaMemcpyAsync(stage_out->w, p.w, words * 4, cudaMemcpyDeviceToDevice, s);
  };
  // fill_and_delete: fill, label, polygon area, drop <= 1000
  Plane filled = take_plane(a, H, W);
  if (fill(ctx, a, in, filled, SLOT_BG, s)) return 1;
  emit_stage(0, filled);
  RunSet F;
  long long* a2 = nullptr;
  if (label_area(ctx, a, filled, SLOT_F, &F, &a2, s)) return 1;
  Plane keep = take_plane(a, H, W);
  rle::keep_large<<<g, TPB, 0, s>>>(F, a2, 2LL * K.fuse_min_area, 0, keep);
  emit_stage(1, keep);
  // eroede_dilate_process: 1x5 / 5x1 kernels, 5 iterations == one 1x21 / 21x1 erosion
  Plane eh = take_plane(a, H, W), ev = take_plane(a, H, W);
  rle::morph_h<true><<<g, TPB, 0, s>>>(keep, eh, K.fuse_split_half);
  rle::morph_v<true><<<g, TPB, 0, s>>>(keep, ev, K.fuse_split_half);
  ctx->launches += 3;
  emit_stage(2, eh);
  emit_stage(3, ev);
  RunSet EH, EV;
  long long *a2h = nullptr, *a2v = nullptr;
  if (label_area(ctx, a, eh, SLOT_EH, &EH, &a2h, s)) return 1;
  if (label_area(ctx, a, ev, SLOT_EV, &EV, &a2v, s)) return 1;
  // per object: number of fragments and of fragments that survive the <= 500 filter, in both directions
  void* cnt = nullptr;
  const size_t nobj = static_cast<size_t>(F.nruns) + 1;
  if (ctx->pool.get(SLOT_CNT, sizeof(int) * 4 * nobj, &cnt)) return 1;
  BD_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * 4 * nobj, s));
  int *cntH = static_cast<int*>(cnt), *survH = cntH + nobj, *cntV = survH + nobj, *survV = cntV + nobj;
  const long long frag2 = 2LL * K.fuse_min_fragment;
  rle::count_fragments<<<g, TPB, 0, s>>>(EH, a2h, F, cntH, survH, frag2);
  rle::count_fragments<<<g, TPB, 0, s>>>(EV, a2v, F, cntV, survV, frag2);
  // decision + rasterisation: whole objects, surviving fragments dilated back
  Plane whole = take_plane(a, H, W), sh = take_plane(a, H, W), sv = take_plane(a, H, W);
  Plane dh = take_plane(a, H, W), dv = take_plane(a, H, W);
  rle::raster_whole<<<g, TPB, 0, s>>>(F, keep, cntH, survH, cntV, survV, whole);
  rle::raster_seeds<<<g, TPB, 0, s>>>(EH, a2h, F, cntH, survH, cntV, survV, 0, frag2, sh);
  rle::raster_seeds<<<g, TPB, 0, s>>>(EV, a2v, F, cntH, survH, cntV, survV, 1, frag2, sv);
  rle::morph_h<false><<<g, TPB, 0, s>>>(sh, dh, K.fuse_split_half);
  rle::morph_v<false><<<g, TPB, 0, s>>>(sv, dv, K.fuse_split_half);
  rle::or3<<<g, TPB, 0, s>>>(whole, dh, dv, out);
  ctx->launches += 8;
  emit_stage(4, whole);
  emit_stage(5, sh);
  emit_stage(6, sv);
  emit_stage(7, out);
  BD_CUDA(cudaGetLastError());
  a.off = mark;
  return 0;
}

// scratch one clean-up needs beyond its in / out planes (bytes), for Arena::reserve
size_t cleanup_scratch_bytes(int H, int W) {
  const size_t plane = static_cast<size_t>(H) * rle::words_per_row(W) * 4 + 256;
  const size_t rows = (static_cast<size_t>(H) + 1) * 4 + 256;
  return 12 * plane /* planes */ + 4 * (plane + rows) /* run sets: wprefix + row table */ + 4096;
}

}  // namespace post
}  // namespace bd

// test hook: per-pixel component labels (raster index of the component's first pixel, -1 = not in the set) of the
// set (fg != 0) or clear (fg == 0) pixels, 8- or 4-connected -- the run-based labelling of rle.cuh made comparable
// with cv2.connectedComponents
namespace bd {
namespace post {
static __global__ void __launch_bounds__(rle::TPB) labels_to_pixels(rle::RunSet r, int* __restrict__ root_pixel, int* __restrict__ out) {
  const rle::Plane& p = r.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(rle::TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * rle::TPB) {
    const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
    const uint32_t cur = p.w[i], prev = wd ? p.w[i - 1] : 0u;
    if (root_pixel) {  // pass 1: every run start records its pixel index
      uint32_t st = rle::starts_of(cur, prev);
      int rid = static_cast<int>(r.wprefix[i]);
      while (st) {
        const int j = __ffs(st) - 1;
        st &= st - 1;
        root_pixel[rid++] = y * p.W + wd * 32 + j;
      }
    } else {
      for (int j = 0; j < 32 && wd * 32 + j < p.W; ++j) out[static_cast<size_t>(y) * p.W + wd * 32 + j] = -1;
    }
  }
}
static __global__ void __launch_bounds__(rle::TPB) labels_write(rle::RunSet r, const int* __restrict__ root_pixel, int* __restrict__ out) {
  const rle::Plane& p = r.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(rle::TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * rle::TPB) {
    const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
    const uint32_t cur = p.w[i], prev = wd ? p.w[i - 1] : 0u;
    rle::for_runs(cur, prev, r.wprefix[i], [&](int rid, uint32_t mask, int) {
      const int lab = root_pixel[r.P[rid]];
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        out[static_cast<size_t>(y) * p.W + wd * 32 + j] = lab;
      }
    });
  }
}
}  // namespace post
}  // namespace bd


extern "C" {

int bd_post_constants(bd_ctx* ctx, bd_post_constants_t* out) {
  BD_CHECK(ctx && out, "bad arguments");
  const post::PostConstants& K = ctx->consts;
  out->fuse_min_area = K.fuse_min_area; out->fuse_min_fragment = K.fuse_min_fragment;
  out->fuse_split_width = 2 * K.fuse_split_half + 1; out->fuse_votes = K.fuse_votes;
  out->edge_min_area = K.edge_min_area; out->edge_min_fragment = K.edge_min_fragment;
  out->edge_split_width = 2 * K.edge_split_half + 1; out->edge_iou = K.edge_iou; out->edge_min_moment = K.edge_min_moment;
  out->tier_small = K.tier_small; out->tier_mid = K.tier_mid; out->tier_big0 = K.tier_big0; out->tier_big1 = K.tier_big1;
  out->tier_big2 = K.tier_big2; out->eps_default = K.eps_default; out->eps_mid_mult = K.eps_mid_mult;
  out->eps_big0 = K.eps_big0; out->eps_big1 = K.eps_big1; out->eps_big2 = K.eps_big2;
  return 0;
}

size_t bd_plane_words_per_row(int w) { return static_cast<size_t>(rle::words_per_row(w)); }

int bd_mask_pack(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, uint32_t* plane_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && mask_dev && plane_dev && h >= 1 && w >= 1, "bad arguments");
  rle::Plane p{plane_dev, h, w, rle::words_per_row(w)};
  return post::pack(ctx, mask_dev, p, static_cast<cudaStream_t>(stream));
}

int bd_mask_unpack(bd_ctx* ctx, const uint32_t* plane_dev, int h, int w, uint8_t* mask_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && mask_dev && plane_dev && h >= 1 && w >= 1, "bad arguments");
  rle::Plane p{const_cast<uint32_t*>(plane_dev), h, w, rle::words_per_row(w)};
  return post::unpack(ctx, p, mask_dev, static_cast<cudaStream_t>(stream));
}

int bd_mask_cleanup(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, uint8_t* out_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && mask_dev && out_dev && h >= 1 && w >= 1, "bad arguments");
  BD_CHECK(static_cast<size_t>(h) * w < (1ull << 31), "scene too large for int32 pixel indices");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  post::Arena& a = ctx->arena;
  const size_t plane = static_cast<size_t>(h) * rle::words_per_row(w) * 4 + 256;
  if (a.reserve(2 * plane + post::cleanup_scratch_bytes(h, w))) return 1;
  rle::Plane in = post::take_plane(a, h, w), out = post::take_plane(a, h, w);
  if (post::pack(ctx, mask_dev, in, s)) return 1;
  if (post::cleanup_plane(ctx, a, in, out, s, -1, nullptr)) return 1;
  return post::unpack(ctx, out, out_dev, s);
}

// planes5_dev: five planes back to back (each h * bd_plane_words_per_row(w) words)
int bd_fuse_planes(bd_ctx* ctx, const uint32_t* planes5_dev, int cleaned, int h, int w, uint8_t* fused_dev,
                   uint32_t* fused_plane_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && planes5_dev && (fused_dev || fused_plane_dev) && h >= 1 && w >= 1, "bad arguments");
  NvtxRange nvtx("bd:fuse");
  BD_CHECK(static_cast<size_t>(h) * w < (1ull << 31), "scene too large for int32 pixel indices");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  post::Arena& a = ctx->arena;
  const int wp = rle::words_per_row(w);
  const size_t words = static_cast<size_t>(h) * wp;
  const size_t plane = words * 4 + 256;
  if (a.reserve(8 * plane + post::cleanup_scratch_bytes(h, w))) return 1;
  uint32_t* cleaned5 = a.take<uint32_t>(5 * words);
  rle::Plane voted = post::take_plane(a, h, w), out = post::take_plane(a, h, w);
  const uint32_t* vote_src = planes5_dev;
  if (!cleaned) {
    for (int k = 0; k < 5; ++k) {
      rle::Plane in{const_cast<uint32_t*>(planes5_dev) + k * words, h, w, wp};
      rle::Plane ck{cleaned5 + k * words, h, w, wp};
      if (post::cleanup_plane(ctx, a, in, ck, s, -1, nullptr)) return 1;
    }
    vote_src = cleaned5;
  }
  rle::vote3of5<<<post::grid_words(ctx, words), post::TPB, 0, s>>>(vote_src, words, voted.w);
  ctx->launches++;
  BD_CUDA(cudaGetLastError());
  if (post::cleanup_plane(ctx, a, voted, out, s, -1, nullptr)) return 1;
  if (fused_plane_dev) BD_CUDA(cudaMemcpyAsync(fused_plane_dev, out.w, words * 4, cudaMemcpyDeviceToDevice, s));
  if (fused_dev && post::unpack(ctx, out, fused_dev, s)) return 1;
  return 0;
}

static int fuse_u8(bd_ctx* ctx, const uint8_t* masks5_dev, int cleaned, int h, int w, uint8_t* fused_dev, cudaStream_t s) {
  const int wp = rle::words_per_row(w);
  const size_t words = static_cast<size_t>(h) * wp;
  void* planes = nullptr;
  if (ctx->pool.get(post::SLOT_IN5, 5 * words * 4, &planes)) return 1;
  const size_t n = static_cast<size_t>(h) * w;
  for (int k = 0; k < 5; ++k) {
    rle::Plane p{static_cast<uint32_t*>(planes) + k * words, h, w, wp};
    if (post::pack(ctx, masks5_dev + k * n, p, s)) return 1;
  }
  return bd_fuse_planes(ctx, static_cast<const uint32_t*>(planes), cleaned, h, w, fused_dev, nullptr, s);
}

int bd_fuse(bd_ctx* ctx, const uint8_t* masks5_dev, int h, int w, uint8_t* fused_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && masks5_dev && fused_dev && h >= 1 && w >= 1, "bad arguments");
  return fuse_u8(ctx, masks5_dev, 0, h, w, fused_dev, static_cast<cudaStream_t>(stream));
}

int bd_fuse_cleaned(bd_ctx* ctx, const uint8_t* cleaned5_dev, int h, int w, uint8_t* fused_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && cleaned5_dev && fused_dev && h >= 1 && w >= 1, "bad arguments");
  return fuse_u8(ctx, cleaned5_dev, 1, h, w, fused_dev, static_cast<cudaStream_t>(stream));
}

// debugging / test hook: an intermediate plane of one clean-up pass as a u8 mask.
// stage: 0 holes filled, 1 area filter, 2 / 3 horizontal / vertical erosion, 4 objects kept whole,
// 5 / 6 surviving fragments of the horizontal / vertical split, 7 result
int bd_debug_cleanup_stage(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, int stage, uint8_t* out_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && mask_dev && out_dev && h >= 1 && w >= 1 && stage >= 0 && stage <= 7, "bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  post::Arena& a = ctx->arena;
  const size_t plane = static_cast<size_t>(h) * rle::words_per_row(w) * 4 + 256;
  if (a.reserve(3 * plane + post::cleanup_scratch_bytes(h, w))) return 1;
  rle::Plane in = post::take_plane(a, h, w), out = post::take_plane(a, h, w), st = post::take_plane(a, h, w);
  if (post::pack(ctx, mask_dev, in, s)) return 1;
  if (post::cleanup_plane(ctx, a, in, out, s, stage, &st)) return 1;
  return post::unpack(ctx, st, out_dev, s);
}

int bd_debug_labels(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, int fg, int conn8, int32_t* labels_dev, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && mask_dev && labels_dev && h >= 1 && w >= 1, "bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  post::Arena& a = ctx->arena;
  if (a.reserve(post::cleanup_scratch_bytes(h, w))) return 1;
  rle::Plane in = post::take_plane(a, h, w), set = in;
  if (post::pack(ctx, mask_dev, in, s)) return 1;
  const size_t words = static_cast<size_t>(h) * in.wp;
  const int g = post::grid_words(ctx, words);
  if (!fg) {
    set = post::take_plane(a, h, w);
    rle::complement<<<g, rle::TPB, 0, s>>>(in, set);
  }
  rle::RunSet r;
  if (post::build_runs(ctx, a, set, conn8 != 0, post::SLOT_F, &r, s)) return 1;
  void* rp = nullptr;
  if (ctx->pool.get(post::SLOT_F + 1, sizeof(int) * (static_cast<size_t>(r.nruns) + 1), &rp)) return 1;
  post::labels_to_pixels<<<g, rle::TPB, 0, s>>>(r, nullptr, labels_dev);
  post::labels_to_pixels<<<g, rle::TPB, 0, s>>>(r, static_cast<int*>(rp), nullptr);
  post::labels_write<<<g, rle::TPB, 0, s>>>(r, static_cast<const int*>(rp), labels_dev);
  BD_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
