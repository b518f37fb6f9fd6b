// rle_emul.cpp -- TEST INFRASTRUCTURE (CPU): the word-parallel kernels of building_detection_b200/csrc/rle.cuh compiled
// with g++ (host_emul.h) and driven in the order csrc/post.cu and csrc/contours.cu drive them on the GPU.  The three
// cooperative kernels (count_rows / scan_rows / emit_prefix) are restated here as plain loops.  Built and called by
// tests/test_rle_emul.py; never part of the product.
#define BD_HOST_EMUL 1
#include "host_emul.h"

#include <vector>

#include "rle.cuh"

using namespace bd::rle;

namespace {

struct Pl {
  std::vector<uint32_t> buf;
  Plane p;
  Pl(int H, int W) : buf(static_cast<size_t>(H) * words_per_row(W), 0u) { p = Plane{buf.data(), H, W, words_per_row(W)}; }
};
struct RS {
  std::vector<uint32_t> wprefix;
  std::vector<int> P;
  RunSet r;
};

template <class K, class... A>
void launch(K k, A... a) { emul_launch(TPB, k, a...); }

void number_runs(Plane p, RS* rs) {  // count_rows + scan_rows + emit_prefix
  rs->wprefix.assign(static_cast<size_t>(p.H) * p.wp, 0u);
  int n = 0;
  for (int y = 0; y < p.H; ++y)
    for (int wd = 0; wd < p.wp; ++wd) {
      const size_t i = static_cast<size_t>(y) * p.wp + wd;
      rs->wprefix[i] = n;
      n += __popc(starts_of(p.w[i], wd ? p.w[i - 1] : 0u));
    }
  rs->P.resize(n + 1);
  for (int i = 0; i <= n; ++i) rs->P[i] = i;
  rs->r = RunSet{p, rs->wprefix.data(), rs->P.data(), n};
}
void build_runs(Plane p, bool conn8, RS* rs) {
  number_runs(p, rs);
  if (rs->r.nruns > 0) {
    if (p.H > 1) {
      if (conn8) launch(merge_rows<true>, rs->r);
      else launch(merge_rows<false>, rs->r);
    }
    launch(compress_runs, rs->r.P, rs->r.nruns);
    launch(flatten_runs, rs->r.P, rs->r.nruns);
  }
}
void fill(Plane fg, Plane filled) {
  Pl bg(fg.H, fg.W);
  launch(complement, fg, bg.p);
  RS B;
  build_runs(bg.p, false, &B);
  std::vector<uint8_t> outside(B.r.nruns + 1, 0);
  launch(mark_outside, B.r, outside.data());
  launch(fill_holes, fg, B.r, static_cast<const uint8_t*>(outside.data()), filled);
}
void label_area(Plane p, RS* rs, std::vector<long long>* a2) {
  build_runs(p, true, rs);
  a2->assign(rs->r.nruns + 1, 0);
  if (rs->r.nruns > 0) launch(polygon_area2, rs->r, a2->data());
}
void pack(const uint8_t* src, Plane p) {
  if (p.W % 16 == 0) launch(pack_u8<true>, src, p);
  else launch(pack_u8<false>, src, p);
}
void unpack(Plane p, uint8_t* dst) {
  if (p.W % 16 == 0) launch(unpack_u8<true>, p, dst);
  else launch(unpack_u8<false>, p, dst);
}

// csrc/post.cu: cleanup_plane
void cleanup_plane(Plane in, Plane out, int stage, Plane* stage_out, int min_area, int min_frag, int half) {
  const int H = in.H, W = in.W;
  auto emit = [&](int id, Plane p) {
    if (stage == id && stage_out) memcpy(stage_out->w, p.w, static_cast<size_t>(H) * p.wp * 4);
  };
  Pl filled(H, W), keep(H, W), eh(H, W), ev(H, W), whole(H, W), sh(H, W), sv(H, W), dh(H, W), dv(H, W);
  fill(in, filled.p);
  emit(0, filled.p);
  RS F, EH, EV;
  std::vector<long long> a2, a2h, a2v;
  label_area(filled.p, &F, &a2);
  launch(keep_large, F.r, static_cast<const long long*>(a2.data()), 2LL * min_area, 0, keep.p);
  emit(1, keep.p);
  launch(morph_h<true>, keep.p, eh.p, half);
  launch(morph_v<true>, keep.p, ev.p, half);
  emit(2, eh.p);
  emit(3, ev.p);
  label_area(eh.p, &EH, &a2h);
  label_area(ev.p, &EV, &a2v);
  const size_t nobj = F.r.nruns + 1;
  std::vector<int> cnt(4 * nobj, 0);
  int *cntH = cnt.data(), *survH = cntH + nobj, *cntV = survH + nobj, *survV = cntV + nobj;
  const long long frag2 = 2LL * min_frag;
  launch(count_fragments, EH.r, static_cast<const long long*>(a2h.data()), F.r, cntH, survH, frag2);
  launch(count_fragments, EV.r, static_cast<const long long*>(a2v.data()), F.r, cntV, survV, frag2);
  launch(raster_whole, F.r, keep.p, static_cast<const int*>(cntH), static_cast<const int*>(survH), static_cast<const int*>(cntV),
         static_cast<const int*>(survV), whole.p);
  launch(raster_seeds, EH.r, static_cast<const long long*>(a2h.data()), F.r, static_cast<const int*>(cntH), static_cast<const int*>(survH),
         static_cast<const int*>(cntV), static_cast<const int*>(survV), 0, frag2, sh.p);
  launch(raster_seeds, EV.r, static_cast<const long long*>(a2v.data()), F.r, static_cast<const int*>(cntH), static_cast<const int*>(survH),
         static_cast<const int*>(cntV), static_cast<const int*>(survV), 1, frag2, sv.p);
  launch(morph_h<false>, sh.p, dh.p, half);
  launch(morph_v<false>, sv.p, dv.p, half);
  launch(or3, whole.p, dh.p, dv.p, out);
  emit(4, whole.p);
  emit(5, sh.p);
  emit(6, sv.p);
  emit(7, out);
}

// csrc/contours.cu: trace_set -- every component of the plane, in OpenCV's order (descending first pixel)
void trace_all(Plane p, std::vector<int>* npts_out, std::vector<int2>* pts_out) {
  RS R;
  build_runs(p, true, &R);
  std::vector<int> roots;
  for (int y = 0; y < p.H; ++y)
    for (int wd = 0; wd < p.wp; ++wd) {
      const size_t i = static_cast<size_t>(y) * p.wp + wd;
      uint32_t st = starts_of(p.w[i], wd ? p.w[i - 1] : 0u);
      int rid = R.wprefix[i];
      while (st) { const int j = __ffs(st) - 1; st &= st - 1; if (R.P[rid] == rid) roots.push_back(y * p.W + wd * 32 + j); ++rid; }
    }
  std::sort(roots.begin(), roots.end(), [](int a, int b) { return a > b; });
  const int n = static_cast<int>(roots.size());
  // count_row_cracks + scan_rows + emit_crack_base
  std::vector<uint32_t> cbase(static_cast<size_t>(p.H) * p.wp);
  int nc = 0;
  for (int y = 0; y < p.H; ++y)
    for (int wd = 0; wd < p.wp; ++wd) {
      cbase[static_cast<size_t>(y) * p.wp + wd] = nc;
      const CrackMasks k = crack_masks(p, y, wd);
      nc += __popc(k.m[0]) + __popc(k.m[1]) + __popc(k.m[2]) + __popc(k.m[3]);
    }
  std::vector<int> start_of(nc + 1, 0), start_crack(n + 1), npts(n + 1, 0), nxt(nc + 1), ws(nc + 1), nxt2(nc + 1), ws2(nc + 1), term_of(nc + 1);
  launch(mark_starts, p, static_cast<const uint32_t*>(cbase.data()), static_cast<const int*>(roots.data()), n, start_of.data(),
         start_crack.data(), npts.data());
  launch(init_cracks, p, static_cast<const uint32_t*>(cbase.data()), static_cast<const int*>(start_of.data()), nxt.data(), ws.data(),
         term_of.data());
  int rounds = 1;
  while ((1ll << rounds) < nc + 1) ++rounds;
  int *a = nxt.data(), *b = ws.data(), *a2 = nxt2.data(), *b2 = ws2.data();
  for (int r = 0; r < rounds; ++r) {
    launch(jump_cracks, static_cast<const int*>(a), static_cast<const int*>(b), a2, b2, nc);
    std::swap(a, a2);
    std::swap(b, b2);
  }
  launch(contour_totals, static_cast<const int*>(start_crack.data()), static_cast<const int*>(b), n, npts.data());
  std::vector<long long> off(n + 1, 0);
  for (int c = 0; c < n; ++c) off[c + 1] = off[c] + npts[c];
  pts_out->assign(off[n] + 1, make_int2(-1, -1));
  launch(scatter_points, p, static_cast<const uint32_t*>(cbase.data()), static_cast<const int*>(a), static_cast<const int*>(b),
         static_cast<const int*>(term_of.data()), static_cast<const int*>(npts.data()), static_cast<const long long*>(off.data()),
         static_cast<const int*>(roots.data()), n, pts_out->data());
  pts_out->resize(off[n]);
  npts.resize(n);
  *npts_out = npts;
}

}  // namespace

extern "C" {

// external contours of every component of an (h, w) u8 mask, OpenCV order; returns the number of contours, fills
// npts[contour] and xy[2 * point] (capacities given)
int emul_contours(const uint8_t* mask, int h, int w, int* npts, int cap_contours, int* xy, int cap_points) {
  Pl in(h, w);
  pack(mask, in.p);
  std::vector<int> n;
  std::vector<int2> pts;
  trace_all(in.p, &n, &pts);
  if (static_cast<int>(n.size()) > cap_contours || static_cast<int>(pts.size()) > cap_points) return -1;
  for (size_t i = 0; i < n.size(); ++i) npts[i] = n[i];
  for (size_t i = 0; i < pts.size(); ++i) { xy[2 * i] = pts[i].x; xy[2 * i + 1] = pts[i].y; }
  return static_cast<int>(n.size());
}
// bounding boxes per pixel's component (x0, y0, x1, y1 as cv::boundingRect gives x, y, x+w, y+h)
void emul_bboxes(const uint8_t* mask, int h, int w, int* bb_px) {
  Pl in(h, w);
  pack(mask, in.p);
  RS R;
  build_runs(in.p, true, &R);
  std::vector<int> bmin(2 * (R.r.nruns + 1), 0x7f7f7f7f), bmax(2 * (R.r.nruns + 1), -1), rr(R.r.nruns + 1), bb(4 * (R.r.nruns + 1));
  for (int i = 0; i <= R.r.nruns; ++i) rr[i] = i;
  launch(run_bboxes, R.r, bmin.data(), bmax.data());
  launch(gather_bboxes, static_cast<const int*>(bmin.data()), static_cast<const int*>(bmax.data()), static_cast<const int*>(rr.data()),
         R.r.nruns, bb.data());
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const size_t i = static_cast<size_t>(y) * in.p.wp + (x >> 5);
      const bool on = (in.p.w[i] >> (x & 31)) & 1u;
      for (int k = 0; k < 4; ++k) bb_px[(static_cast<size_t>(y) * w + x) * 4 + k] = on ? bb[4 * R.P[run_at(R.r, y, x)] + k] : 0;
    }
}

// one clean-up pass (or an intermediate plane of it, stage 0..7; -1 / 7 = result) of an (h, w) u8 mask
void emul_cleanup(const uint8_t* mask, int h, int w, int stage, uint8_t* out) {
  Pl in(h, w), o(h, w), st(h, w);
  pack(mask, in.p);
  cleanup_plane(in.p, o.p, stage, &st.p, 1000, 500, 10);
  unpack(stage >= 0 && stage < 7 ? st.p : o.p, out);
}
// model_confuse on five (h, w) u8 masks stored back to back
void emul_fuse(const uint8_t* masks5, int h, int w, uint8_t* out) {
  const size_t words = static_cast<size_t>(h) * words_per_row(w);
  std::vector<uint32_t> cleaned(5 * words);
  for (int k = 0; k < 5; ++k) {
    Pl in(h, w);
    pack(masks5 + static_cast<size_t>(k) * h * w, in.p);
    Plane ck{cleaned.data() + k * words, h, w, words_per_row(w)};
    cleanup_plane(in.p, ck, -1, nullptr, 1000, 500, 10);
  }
  Pl voted(h, w), o(h, w);
  launch(vote3of5, static_cast<const uint32_t*>(cleaned.data()), words, voted.p.w);
  cleanup_plane(voted.p, o.p, -1, nullptr, 1000, 500, 10);
  unpack(o.p, out);
}
// per-pixel labels (raster index of the component's first pixel, -1 outside the set)
void emul_labels(const uint8_t* mask, int h, int w, int fg, int conn8, int32_t* labels) {
  Pl in(h, w), cm(h, w);
  pack(mask, in.p);
  Plane set = in.p;
  if (!fg) { launch(complement, in.p, cm.p); set = cm.p; }
  RS R;
  build_runs(set, conn8 != 0, &R);
  std::vector<int> first(R.r.nruns + 1, -1);
  for (int y = 0; y < h; ++y)
    for (int wd = 0; wd < set.wp; ++wd) {
      const size_t i = static_cast<size_t>(y) * set.wp + wd;
      uint32_t st = starts_of(set.w[i], wd ? set.w[i - 1] : 0u);
      int rid = R.wprefix[i];
      while (st) { const int j = __ffs(st) - 1; st &= st - 1; first[rid++] = y * w + wd * 32 + j; }
    }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const size_t i = static_cast<size_t>(y) * set.wp + (x >> 5);
      labels[static_cast<size_t>(y) * w + x] = ((set.w[i] >> (x & 31)) & 1u) ? first[R.P[run_at(R.r, y, x)]] : -1;
    }
}
// 2 x signed polygon area per pixel's component (0 outside the set), 8-connected
void emul_area2(const uint8_t* mask, int h, int w, long long* area2_px) {
  Pl in(h, w);
  pack(mask, in.p);
  RS R;
  std::vector<long long> a2;
  label_area(in.p, &R, &a2);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const size_t i = static_cast<size_t>(y) * in.p.wp + (x >> 5);
      area2_px[static_cast<size_t>(y) * w + x] = ((in.p.w[i] >> (x & 31)) & 1u) ? a2[R.P[run_at(R.r, y, x)]] : 0;
    }
}
// 1 x K / K x 1 erosion or dilation (K = 2 half + 1)
void emul_morph(const uint8_t* mask, int h, int w, int half, int vertical, int erode, uint8_t* out) {
  Pl in(h, w), o(h, w);
  pack(mask, in.p);
  if (vertical) { if (erode) launch(morph_v<true>, in.p, o.p, half); else launch(morph_v<false>, in.p, o.p, half); }
  else { if (erode) launch(morph_h<true>, in.p, o.p, half); else launch(morph_h<false>, in.p, o.p, half); }
  unpack(o.p, out);
}

}  // extern "C"
