// contours.cu -- contour extraction and polygon simplification (reference edge_3.py:_detection, 310-387).
//
// Device side: hole fill, 8-connected labelling, polygon-area filter (<= 100), 1x7 and 7x1 erosions with their
// fragment filter (< 50) on bit-packed planes with run-based labels (rle.cuh, shared with the fusion stage); then
// border following of every kept component on the unpacked u8 planes (one thread per
// contour; the traced sequence is the one cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) returns: start at
// the component's first raster pixel, first step down/left, contours listed by descending start pixel), and the
// all-pairs bounding-box IoU matching of process_td / process_rl.
// Host side (O(#boundary points), double precision, compiled without FMA contraction so that it rounds like
// OpenCV's scalar code): the list surgery of detction_overlap_building and the area-tiered Douglas-Peucker
// simplification (restatement of cv::approxPolyDP / arcLength / contourArea for integer contours).
#include "../../include/bd_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "post_ws.cuh"
#include "rle.cuh"

using namespace bd;

namespace bd {
namespace post {  // post.cu
int grid_words(bd_ctx* ctx, size_t words);
rle::Plane take_plane(Arena& a, int H, int W);
int pack(bd_ctx* ctx, const uint8_t* src, rle::Plane p, cudaStream_t s);
int unpack(bd_ctx* ctx, rle::Plane p, uint8_t* dst, cudaStream_t s);
int fill(bd_ctx* ctx, Arena& a, rle::Plane fg, rle::Plane filled, int slot, cudaStream_t s);
int label_area(bd_ctx* ctx, Arena& a, rle::Plane p, int slot, rle::RunSet* rs, long long** area2, cudaStream_t s);
size_t cleanup_scratch_bytes(int H, int W);
}  // namespace post
}  // namespace bd

namespace bd {
namespace cont {

constexpr int TPB = 256;

// edge_3.py:26-47 for every initial box against all eroded boxes: index of the first maximum IoU if any IoU
// exceeds 0.5, else -1.  Boxes are (x0, y0, x1, y1); IoU in float64 exactly like numpy's int64 / int64.
static __global__ void __launch_bounds__(TPB) match_boxes(const int* __restrict__ a, int na, const int* __restrict__ b, int nb,
                                                   int* __restrict__ res) {
  __shared__ int sb[TPB * 4];
  const int i = blockIdx.x * TPB + threadIdx.x;
  int ax0 = 0, ay0 = 0, ax1 = 0, ay1 = 0;
  if (i < na) { ax0 = a[4 * i]; ay0 = a[4 * i + 1]; ax1 = a[4 * i + 2]; ay1 = a[4 * i + 3]; }
  const long long aarea = static_cast<long long>(ax1 - ax0) * (ay1 - ay0);
  double best = -1.0;
  int besti = 0;
  bool any = false;
  for (int j0 = 0; j0 < nb; j0 += TPB) {
    const int m = min(TPB, nb - j0);
    __syncthreads();
    for (int t = threadIdx.x; t < m * 4; t += TPB) sb[t] = b[4 * j0 + t];
    __syncthreads();
    if (i < na)
      for (int j = 0; j < m; ++j) {
        const int bx0 = sb[4 * j], by0 = sb[4 * j + 1], bx1 = sb[4 * j + 2], by1 = sb[4 * j + 3];
        const long long iw = max(min(ax1, bx1) - max(ax0, bx0), 0), ih = max(min(ay1, by1) - max(ay0, by0), 0);
        const long long inter = iw * ih;
        if (inter == 0) {  // disjoint boxes (almost all pairs): IoU 0 without the float64 division
          if (best < 0.0) { best = 0.0; besti = j0 + j; }
          continue;
        }
        const long long uni = aarea + static_cast<long long>(bx1 - bx0) * (by1 - by0) - inter;
        const double v = static_cast<double>(inter) / static_cast<double>(uni);
        if (v > best) { best = v; besti = j0 + j; }
        any |= v > 0.5;
      }
  }
  if (i < na) res[i] = any ? besti : -1;
}

// ------------------------------------------------------------------------------------------ host geometry
struct Pt { int x, y; };

// cv::contourArea for integer points: |sum(x_{i-1} y_i - x_i y_{i-1})| / 2 (exact in double)
static double contour_area(const Pt* p, int n) {
  if (n == 0) return 0.0;
  double a = 0;
  Pt prev = p[n - 1];
  for (int i = 0; i < n; ++i) {
    a += static_cast<double>(prev.x) * p[i].y - static_cast<double>(prev.y) * p[i].x;
    prev = p[i];
  }
  return std::fabs(a * 0.5);
}
// cv::arcLength(closed): float32 segment lengths summed in double
static double arc_length_closed(const Pt* p, int n) {
  if (n <= 1) return 0.0;
  double per = 0;
  float px = static_cast<float>(p[n - 1].x), py = static_cast<float>(p[n - 1].y);
  for (int i = 0; i < n; ++i) {
    const float x = static_cast<float>(p[i].x), y = static_cast<float>(p[i].y);
    const float dx = x - px, dy = y - py;
    const float d2 = dx * dx + dy * dy;
    per += static_cast<double>(std::sqrt(d2));
    px = x; py = y;
  }
  return per;
}
// cv::approxPolyDP(closed=true) for integer contours (Douglas-Peucker with OpenCV's start-point search and its
// final collinear clean-up).
static void approx_poly_closed(const Pt* src, int count, double eps, std::vector<Pt>& dst) {
  dst.clear();
  if (count == 0) return;
  struct Range { int start, end; };
  std::vector<Range> stack;
  eps *= eps;
  Range slice{0, 0}, right{0, 0};
  Pt start_pt{-1000000, -1000000}, end_pt{0, 0}, pt{0, 0};
  int pos = 0;
  bool le_eps = false;
  auto read = [&](Pt& q, int& ps) { q = src[ps]; if (++ps >= count) ps = 0; };
  // 1. approximately the two farthest points
  right.start = 0;
  for (int i = 0; i < 3; ++i) {
    double max_dist = 0;
    pos = (pos + right.start) % count;
    read(start_pt, pos);
    for (int j = 1; j < count; ++j) {
      read(pt, pos);
      const double dx = pt.x - start_pt.x, dy = pt.y - start_pt.y;
      const double dist = dx * dx + dy * dy;
      if (dist > max_dist) { max_dist = dist; right.start = j; }
    }
    le_eps = max_dist <= eps;
  }
  // 2. initial two slices
  if (!le_eps) {
    right.end = slice.start = pos % count;
    slice.end = right.start = (right.start + slice.start) % count;
    stack.push_back(right);
    stack.push_back(slice);
  } else {
    dst.push_back(start_pt);
  }
  // 3. recursive subdivision
  while (!stack.empty()) {
    slice = stack.back();
    stack.pop_back();
    end_pt = src[slice.end];
    pos = slice.start;
    read(start_pt, pos);
    if (pos != slice.end) {
      // distance to the SEGMENT start-end (cv2 >= 4.10 measures points that project beyond an end point to that
      // end point; pinned by the differential test against cv2 4.13 in tests/test_post_cpu.py)
      double max_dist = 0;
      const double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y;
      const double len2 = dx * dx + dy * dy;
      while (pos != slice.end) {
        read(pt, pos);
        const double px = pt.x - start_pt.x, py = pt.y - start_pt.y;
        const double dot = px * dx + py * dy;
        double dist;
        if (len2 == 0 || dot < 0) dist = std::sqrt(px * px + py * py);
        else if (dot > len2) {
          const double qx = pt.x - end_pt.x, qy = pt.y - end_pt.y;
          dist = std::sqrt(qx * qx + qy * qy);
        } else dist = std::fabs(py * dx - px * dy) / std::sqrt(len2);
        if (dist > max_dist) { max_dist = dist; right.start = (pos + count - 1) % count; }
      }
      le_eps = max_dist * max_dist <= eps;
    } else {
      le_eps = true;
      start_pt = src[slice.start];
    }
    if (le_eps) {
      dst.push_back(start_pt);
    } else {
      right.end = slice.end;
      slice.end = right.start;
      stack.push_back(right);
      stack.push_back(slice);
    }
  }
  // 4. remove points on (almost) straight lines
  int new_count = static_cast<int>(dst.size());
  count = new_count;
  auto read_dst = [&](Pt& q, int& ps) { q = dst[ps]; if (++ps >= count) ps = 0; };
  pos = count - 1;
  read_dst(start_pt, pos);
  int wpos = pos;
  read_dst(pt, pos);
  for (int i = 0; i < count && new_count > 2; ++i) {
    read_dst(end_pt, pos);
    const double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y;
    const double dist = std::fabs((pt.x - start_pt.x) * dy - (pt.y - start_pt.y) * dx);
    const double sip = static_cast<double>(pt.x - start_pt.x) * (end_pt.x - pt.x) +
                       static_cast<double>(pt.y - start_pt.y) * (end_pt.y - pt.y);
    if (dist * dist <= 0.5 * eps * (dx * dx + dy * dy) && dx != 0 && dy != 0 && sip >= 0) {
      new_count--;
      dst[wpos] = start_pt = end_pt;
      if (++wpos >= count) wpos = 0;
      read_dst(pt, pos);
      i++;
      continue;
    }
    dst[wpos] = start_pt = pt;
    if (++wpos >= count) wpos = 0;
    pt = end_pt;
  }
  dst.resize(new_count);
}

// edge_3.py:351-378.  Returns 0 = skip (m00 <= 10), 1 = polygon in `out`, 2 = the 4-vertex search failed: the
// caller must take cv::boxPoints(cv::minAreaRect(contour)) (float32 libm trigonometry, kept on the host side).
static int simplify(const Pt* c, int n, std::vector<Pt>& out, const post::PostConstants& K) {
  const double area = contour_area(c, n);
  const double per = arc_length_closed(c, n);
  double eps = K.eps_default * per;
  if (area <= K.edge_min_moment) return 0;  // cv::moments(contour)["m00"] equals contourArea for a closed integer contour
  if (area < K.tier_small) {                // small_target, edge_3.py:265-286
    approx_poly_closed(c, n, eps, out);
    double rate = K.small_rate0;
    int tries = 0;
    while (out.size() != 4) {
      eps = rate * per;
      rate = rate + K.small_rate_step;
      approx_poly_closed(c, n, eps, out);
      if (++tries > K.small_max_tries) break;
    }
    return out.size() == 4 ? 1 : 2;
  }
  if (K.tier_small < area && area < K.tier_mid) eps = K.eps_mid_mult * eps;
  else if (K.tier_big0 < area && area < K.tier_big1) eps = K.eps_big0 * per;
  else if (K.tier_big1 < area && area <= K.tier_big2) eps = K.eps_big1 * per;
  else if (area > K.tier_big2) eps = K.eps_big2 * per;
  approx_poly_closed(c, n, eps, out);
  return 1;
}

struct HostSet {           // one traced contour list, in cv::findContours order
  int n = 0;
  std::vector<int> bbox;   // 4 per contour
  std::vector<long long> off;  // n + 1
  const Pt* pts = nullptr; // off[n] points in the context's pinned host buffer of this set (valid during the call)
  int* d_bbox = nullptr;   // device copy (matching kernel)
};

static int grid_rows(bd_ctx* ctx, int H) {
  return std::max(1, std::min((H + rle::TPB / 32 - 1) / (rle::TPB / 32), ctx->num_sms * 8));
}

// External contours (cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE)) of every component of rs whose polygon area
// passes the threshold.  `p` is the plane that holds exactly those components (their runs are runs of rs.p).  The
// borders are followed in parallel (rle.cuh: crack successors + pointer jumping), so a scene-sized contour costs the
// same as twenty thousand small ones.  `which` (0..2): the set's pool slots and pinned host buffer.
static int trace_set(bd_ctx* ctx, rle::Plane p, const rle::RunSet& rs, const long long* a2, long long thr2, int strict,
                     cudaStream_t s, HostSet* out, int which, int* n_launch) {
  const int slot0 = which * 8;
  const int H = p.H;
  const size_t words = static_cast<size_t>(H) * p.wp;
  post::DevPool& pool = ctx->pool;
  const bool timing = getenv("BD_POST_TIMING") != nullptr;  // per-step wall clock of this set (synchronises)
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(s);
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[trace_set %d] %-28s %8.2f ms\n", which, what, std::chrono::duration<double, std::milli>(t - t_last).count());
    t_last = t;
  };
  out->n = 0;
  out->off.assign(1, 0);
  out->bbox.clear();
  out->pts = nullptr;
  if (rs.nruns == 0) return 0;
  const int g = post::grid_words(ctx, words);
  // roots + crack numbering, one synchronisation for both counts
  int* d_count = nullptr;  // [0] components, [1] cracks
  if (pool.get(slot0 + 0, 2 * sizeof(int), reinterpret_cast<void**>(&d_count))) return 1;
  BD_CUDA(cudaMemsetAsync(d_count, 0, 2 * sizeof(int), s));
  const int cap = rs.nruns;  // a component has at least one run
  int* d_list = nullptr;     // [cap] first pixels, then [cap] root runs
  if (pool.get(slot0 + 1, sizeof(int) * 2 * static_cast<size_t>(cap), reinterpret_cast<void**>(&d_list))) return 1;
  int* d_rid = d_list + cap;
  rle::collect_roots<<<g, rle::TPB, 0, s>>>(rs, a2, thr2, strict, d_list, d_rid, d_count, cap);
  char* geo = nullptr;  // cbase [words] u32 | rows [H + 1] | bbmin [2 nruns] | bbmax [2 nruns]
  const size_t geo_bytes = words * 4 + (static_cast<size_t>(H) + 1) * 4 + 16 * static_cast<size_t>(rs.nruns) + 64;
  if (pool.get(slot0 + 7, geo_bytes, reinterpret_cast<void**>(&geo))) return 1;
  uint32_t* cbase = reinterpret_cast<uint32_t*>(geo);
  int* rows = reinterpret_cast<int*>(geo + words * 4);
  int* bbmin = rows + H + 1;
  int* bbmax = bbmin + 2 * static_cast<size_t>(rs.nruns);
  rle::count_row_cracks<<<grid_rows(ctx, H), rle::TPB, 0, s>>>(p, rows);
  rle::scan_rows<<<1, 1024, 0, s>>>(rows, H, d_count + 1);
  rle::emit_crack_base<<<grid_rows(ctx, H), rle::TPB, 0, s>>>(p, rows, cbase);
  BD_CUDA(cudaMemsetAsync(bbmin, 0x7f, 8 * static_cast<size_t>(rs.nruns), s));
  BD_CUDA(cudaMemsetAsync(bbmax, 0xff, 8 * static_cast<size_t>(rs.nruns), s));
  rle::run_bboxes<<<g, rle::TPB, 0, s>>>(rs, bbmin, bbmax);
  *n_launch += 5;
  int counts[2] = {0, 0};
  BD_CUDA(cudaMemcpyAsync(counts, d_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  BD_CUDA(cudaStreamSynchronize(s));
  const int cnt = counts[0], nc = counts[1];
  BD_CHECK(cnt <= cap, "too many components for the contour stage");
  out->n = cnt;
  out->off.assign(cnt + 1, 0);
  out->bbox.assign(static_cast<size_t>(cnt) * 4, 0);
  if (cnt == 0) return 0;
  std::vector<int> roots(cnt), rids(cnt);
  BD_CUDA(cudaMemcpyAsync(roots.data(), d_list, sizeof(int) * cnt, cudaMemcpyDeviceToHost, s));
  BD_CUDA(cudaMemcpyAsync(rids.data(), d_rid, sizeof(int) * cnt, cudaMemcpyDeviceToHost, s));
  BD_CUDA(cudaStreamSynchronize(s));
  lap("roots + crack numbering");
  {  // findContours lists the last-found first: descending first pixel (== descending run number)
    std::vector<int> order(cnt);
    for (int i = 0; i < cnt; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return roots[a] > roots[b]; });
    std::vector<int> r2(cnt), q2(cnt);
    for (int i = 0; i < cnt; ++i) { r2[i] = roots[order[i]]; q2[i] = rids[order[i]]; }
    roots.swap(r2);
    rids.swap(q2);
  }
  BD_CUDA(cudaMemcpyAsync(d_list, roots.data(), sizeof(int) * cnt, cudaMemcpyHostToDevice, s));
  BD_CUDA(cudaMemcpyAsync(d_rid, rids.data(), sizeof(int) * cnt, cudaMemcpyHostToDevice, s));
  int *d_npts = nullptr, *d_bbox = nullptr, *d_cr = nullptr;
  if (pool.get(slot0 + 2, sizeof(int) * 2 * static_cast<size_t>(cnt), reinterpret_cast<void**>(&d_npts))) return 1;
  int* d_start = d_npts + cnt;
  if (pool.get(slot0 + 3, sizeof(int) * 4 * static_cast<size_t>(cnt), reinterpret_cast<void**>(&d_bbox))) return 1;
  // per crack: start_of | term_of | nxt | ws | nxt2 | ws2
  const size_t ncp = static_cast<size_t>(nc) + 1;
  if (pool.get(slot0 + 6, sizeof(int) * 6 * ncp, reinterpret_cast<void**>(&d_cr))) return 1;
  int *start_of = d_cr, *term_of = d_cr + ncp, *nxt = d_cr + 2 * ncp, *ws = d_cr + 3 * ncp, *nxt2 = d_cr + 4 * ncp, *ws2 = d_cr + 5 * ncp;
  BD_CUDA(cudaMemsetAsync(start_of, 0, sizeof(int) * ncp, s));
  const int gc = post::grid_words(ctx, static_cast<size_t>(cnt));
  rle::mark_starts<<<gc, rle::TPB, 0, s>>>(p, cbase, d_list, cnt, start_of, d_start, d_npts);
  rle::init_cracks<<<g, rle::TPB, 0, s>>>(p, cbase, start_of, nxt, ws, term_of);
  int rounds = 1;
  while ((1ll << rounds) < static_cast<long long>(nc) + 1 && rounds < 31) ++rounds;
  const int gk = post::grid_words(ctx, static_cast<size_t>(nc));
  for (int r = 0; r < rounds; ++r) {
    rle::jump_cracks<<<gk, rle::TPB, 0, s>>>(nxt, ws, nxt2, ws2, nc);
    std::swap(nxt, nxt2);
    std::swap(ws, ws2);
  }
  rle::contour_totals<<<gc, rle::TPB, 0, s>>>(d_start, ws, cnt, d_npts);
  rle::gather_bboxes<<<gc, rle::TPB, 0, s>>>(bbmin, bbmax, d_rid, cnt, d_bbox);
  *n_launch += 4 + rounds;
  std::vector<int> npts(cnt);
  BD_CUDA(cudaMemcpyAsync(npts.data(), d_npts, sizeof(int) * cnt, cudaMemcpyDeviceToHost, s));
  BD_CUDA(cudaMemcpyAsync(out->bbox.data(), d_bbox, sizeof(int) * 4 * cnt, cudaMemcpyDeviceToHost, s));
  BD_CUDA(cudaStreamSynchronize(s));
  lap("successors + ranking");
  for (int i = 0; i < cnt; ++i) out->off[i + 1] = out->off[i] + npts[i];
  const long long total = out->off[cnt];
  if (timing) fprintf(stderr, "[trace_set %d] %d contours, %d cracks, %d jump rounds, %lld points\n", which, cnt, nc, rounds, total);
  long long* d_off = nullptr;
  int2* d_pts = nullptr;
  if (pool.get(slot0 + 4, sizeof(long long) * (static_cast<size_t>(cnt) + 1), reinterpret_cast<void**>(&d_off))) return 1;
  if (pool.get(slot0 + 5, sizeof(int2) * static_cast<size_t>(std::max<long long>(total, 1)), reinterpret_cast<void**>(&d_pts))) return 1;
  BD_CUDA(cudaMemcpyAsync(d_off, out->off.data(), sizeof(long long) * (cnt + 1), cudaMemcpyHostToDevice, s));
  rle::scatter_points<<<g, rle::TPB, 0, s>>>(p, cbase, nxt, ws, term_of, d_npts, d_off, d_list, cnt, d_pts);
  ++*n_launch;
  // points -> pinned host memory (pageable vectors cost 10+ ms for the 30 MB of a 20 000^2 scene)
  const size_t pbytes = sizeof(int2) * static_cast<size_t>(std::max<long long>(total, 1));
  if (pbytes > ctx->h_pts_cap[which]) {
    if (ctx->h_pts[which]) cudaFreeHost(ctx->h_pts[which]);
    ctx->h_pts[which] = nullptr; ctx->h_pts_cap[which] = 0;
    BD_CUDA(cudaMallocHost(&ctx->h_pts[which], pbytes + pbytes / 4));
    ctx->h_pts_cap[which] = pbytes + pbytes / 4;
  }
  BD_CUDA(cudaMemcpyAsync(ctx->h_pts[which], d_pts, sizeof(int2) * static_cast<size_t>(total), cudaMemcpyDeviceToHost, s));
  BD_CUDA(cudaStreamSynchronize(s));
  lap("scatter + copy out");
  out->pts = static_cast<const Pt*>(ctx->h_pts[which]);
  out->d_bbox = d_bbox;
  return 0;
}

// _match of the oracle (edge_3.py process_td / process_rl): lost = initial contours without counterpart,
// fresh = eroded contours nobody claimed (ascending index)
static int match_sets(bd_ctx* ctx, const HostSet& ini, const HostSet& ero, cudaStream_t s, std::vector<int>* lost,
                      std::vector<int>* fresh, int slot) {
  lost->clear();
  fresh->clear();
  if (ini.n == 0) {
    for (int j = 0; j < ero.n; ++j) fresh->push_back(j);
    return 0;
  }
  if (ero.n == 0) {
    last_error() = "IndexError: too many indices for array (edge_3.py:33 with no eroded contours)";
    return 2;
  }
  int* d_res = nullptr;
  if (ctx->pool.get(slot, sizeof(int) * ini.n, reinterpret_cast<void**>(&d_res))) return 1;
  match_boxes<<<(ini.n + TPB - 1) / TPB, TPB, 0, s>>>(ini.d_bbox, ini.n, ero.d_bbox, ero.n, d_res);
  ctx->launches++;
  std::vector<int> res(ini.n);
  BD_CUDA(cudaMemcpyAsync(res.data(), d_res, sizeof(int) * ini.n, cudaMemcpyDeviceToHost, s));
  BD_CUDA(cudaStreamSynchronize(s));
  std::vector<char> claimed(ero.n, 0);
  for (int i = 0; i < ini.n; ++i) {
    if (res[i] < 0) lost->push_back(i);
    else claimed[res[i]] = 1;
  }
  for (int j = 0; j < ero.n; ++j)
    if (!claimed[j]) fresh->push_back(j);
  return 0;
}

static int best_iou_host(const int* box, const std::vector<const int*>& others) {  // edge_3.py:26-47 on small lists
  double best = -1;
  int besti = 0;
  bool any = false;
  const long long aarea = static_cast<long long>(box[2] - box[0]) * (box[3] - box[1]);
  for (size_t j = 0; j < others.size(); ++j) {
    const int* o = others[j];
    const long long iw = std::max(std::min(box[2], o[2]) - std::max(box[0], o[0]), 0);
    const long long ih = std::max(std::min(box[3], o[3]) - std::max(box[1], o[1]), 0);
    const long long inter = iw * ih;
    const long long uni = aarea + static_cast<long long>(o[2] - o[0]) * (o[3] - o[1]) - inter;
    const double v = static_cast<double>(inter) / static_cast<double>(uni);
    if (v > best) { best = v; besti = static_cast<int>(j); }
    any |= v > 0.5;
  }
  return any ? besti : -1;
}

}  // namespace cont
}  // namespace bd

extern "C" {

void bd_polys_free(bd_polys* p) {
  if (!p) return;
  free(p->offsets); free(p->xs); free(p->ys); free(p->is_float);
  memset(p, 0, sizeof(*p));
}

int bd_contours(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, bd_polys* out, void* stream) {
  BD_ON_CTX(ctx);
  BD_CHECK(ctx && mask_dev && out && h >= 1 && w >= 1, "bad arguments");
  BD_CHECK(static_cast<size_t>(h) * w < (1ull << 31), "scene too large for int32 pixel labels");
  using namespace bd::cont;
  NvtxRange nvtx("bd:contours");
  memset(out, 0, sizeof(*out));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const post::PostConstants& K = ctx->consts;
  post::Arena& ar = ctx->arena;
  const size_t words = static_cast<size_t>(h) * rle::words_per_row(w);
  if (ar.reserve(post::cleanup_scratch_bytes(h, w))) return 1;
  const int g = post::grid_words(ctx, words);
  // BD_POST_TIMING=1: wall-clock of every phase (synchronises the stream; debugging aid)
  const bool timing = getenv("BD_POST_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto t_prev = now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(s);
    auto t = now();
    fprintf(stderr, "[bd_contours] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
    t_prev = t;
  };

  // edge_3.py:317-329: fill every external contour, erase polygon area <= 100 -> initial_img (`keep`)
  rle::Plane in = post::take_plane(ar, h, w), filled = post::take_plane(ar, h, w), keep = post::take_plane(ar, h, w);
  rle::Plane eh = post::take_plane(ar, h, w), ev = post::take_plane(ar, h, w);
  if (post::pack(ctx, mask_dev, in, s)) return 1;
  if (post::fill(ctx, ar, in, filled, post::SLOT_BG, s)) return 1;
  rle::RunSet F, EH, EV;
  long long *a2 = nullptr, *a2h = nullptr, *a2v = nullptr;
  if (post::label_area(ctx, ar, filled, post::SLOT_F, &F, &a2, s)) return 1;
  rle::keep_large<<<g, rle::TPB, 0, s>>>(F, a2, 2LL * K.edge_min_area, 0, keep);
  // :172-199: 1x7 and 7x1 erosion (one iteration), fragments of area < 50 erased
  rle::morph_h<true><<<g, rle::TPB, 0, s>>>(keep, eh, K.edge_split_half);
  rle::morph_v<true><<<g, rle::TPB, 0, s>>>(keep, ev, K.edge_split_half);
  ctx->launches += 3;
  if (post::label_area(ctx, ar, eh, post::SLOT_EH, &EH, &a2h, s)) return 1;
  if (post::label_area(ctx, ar, ev, post::SLOT_EV, &EV, &a2v, s)) return 1;
  // the planes the borders are followed on hold exactly the traced components: eroded fragments of area < 50 are erased
  rle::Plane ehk = post::take_plane(ar, h, w), evk = post::take_plane(ar, h, w);
  rle::keep_large<<<g, rle::TPB, 0, s>>>(EH, a2h, 2LL * K.edge_min_fragment, 1, ehk);
  rle::keep_large<<<g, rle::TPB, 0, s>>>(EV, a2v, 2LL * K.edge_min_fragment, 1, evk);
  ctx->launches += 2;
  BD_CUDA(cudaGetLastError());
  lap("fill/label/area/erode passes");

  // The three contour sets are traced concurrently, each on its own stream from its own host thread: border
  // following is one thread per component, so a scene-sized component (random-init masks fuse into one) leaves
  // the GPU empty while it is walked -- three walks at once cost the time of the longest.
  HostSet ini, td, rl;
  {
    static thread_local cudaStream_t aux[2] = {nullptr, nullptr};
    static thread_local cudaEvent_t ready = nullptr;
    if (!aux[0]) {
      BD_CUDA(cudaStreamCreateWithFlags(&aux[0], cudaStreamNonBlocking));
      BD_CUDA(cudaStreamCreateWithFlags(&aux[1], cudaStreamNonBlocking));
      BD_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    }
    BD_CUDA(cudaEventRecord(ready, s));
    BD_CUDA(cudaStreamWaitEvent(aux[0], ready, 0));
    BD_CUDA(cudaStreamWaitEvent(aux[1], ready, 0));
    int rc[3] = {0, 0, 0}, nl[3] = {0, 0, 0};
    std::string err[3];
    const int device = ctx->device;
    auto job = [&](int k, rle::Plane pl, const rle::RunSet* rs, const long long* ar2, long long thr2, int strict,
                   cudaStream_t st, HostSet* out_set) {
      if (cudaSetDevice(device) != cudaSuccess) { rc[k] = 1; err[k] = "cudaSetDevice failed in a contour worker"; return; }
      rc[k] = trace_set(ctx, pl, *rs, ar2, thr2, strict, st, out_set, k, &nl[k]);
      if (rc[k]) err[k] = bd::last_error();  // thread-local message of the worker
    };
    std::thread t1(job, 1, ehk, &EH, a2h, 2LL * K.edge_min_fragment, 1, aux[0], &td);
    std::thread t2(job, 2, evk, &EV, a2v, 2LL * K.edge_min_fragment, 1, aux[1], &rl);
    job(0, keep, &F, a2, 2LL * K.edge_min_area, 0, s, &ini);
    t1.join();
    t2.join();
    ctx->launches += nl[0] + nl[1] + nl[2];
    for (int k = 0; k < 3; ++k)
      if (rc[k]) return bd::fail(err[k]);
  }
  lap("trace initial + eroded contours (three streams)");

  // detction_overlap_building (:159-262): final list of (set, index); set < 0 marks None
  struct Ref { const HostSet* set; int idx; };
  std::vector<Ref> finals;
  for (int i = 0; i < ini.n; ++i) finals.push_back({&ini, i});
  if (!(td.n == ini.n && rl.n == ini.n)) {
    std::vector<int> lost_td, new_td, lost_rl, new_rl;
    const bool do_td = td.n != ini.n, do_rl = rl.n != ini.n;
    if (do_td) { int rc = match_sets(ctx, ini, td, s, &lost_td, &new_td, 24); if (rc) return rc; }
    if (do_rl) { int rc = match_sets(ctx, ini, rl, s, &lost_rl, &new_rl, 25); if (rc) return rc; }
    for (int i : lost_td) finals[i].set = nullptr;
    for (int i : lost_rl) finals[i].set = nullptr;
    if (do_td && do_rl) {
      if (!new_td.empty() && !new_rl.empty()) {
        std::vector<const int*> rl_boxes;
        for (int j : new_rl) rl_boxes.push_back(&rl.bbox[4 * static_cast<size_t>(j)]);
        std::vector<char> dup(new_rl.size(), 0);
        for (int j : new_td) {
          const int r = best_iou_host(&td.bbox[4 * static_cast<size_t>(j)], rl_boxes);
          finals.push_back({&td, j});
          if (r >= 0) dup[r] = 1;
        }
        for (size_t i = 0; i < new_rl.size(); ++i)
          if (!dup[i]) finals.push_back({&rl, new_rl[i]});
      } else if (!new_td.empty()) {
        for (int j : new_td) finals.push_back({&td, j});
      } else {
        for (int j : new_rl) finals.push_back({&rl, j});
      }
    } else if (do_td) {
      for (int j : new_td) finals.push_back({&td, j});
    } else {
      for (int j : new_rl) finals.push_back({&rl, j});
    }
  }

  lap("match + list surgery");
  // :351-385 per contour.  The Douglas-Peucker passes are independent per contour: spread them over host threads
  // (12 000 buildings took 73 ms on one core), then concatenate in list order.
  struct Simp { int kind = 0; std::vector<Pt> poly; };
  std::vector<Simp> simp(finals.size());
  {
    const size_t nf = finals.size();
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const unsigned nt = static_cast<unsigned>(std::min<size_t>(hw, (nf + 255) / 256));
    auto work = [&](unsigned t, unsigned T) {
      for (size_t i = t; i < nf; i += T) {
        const Ref& r = finals[i];
        if (!r.set) continue;
        const Pt* c = r.set->pts + r.set->off[r.idx];
        const int cn = static_cast<int>(r.set->off[r.idx + 1] - r.set->off[r.idx]);
        simp[i].kind = simplify(c, cn, simp[i].poly, K);
      }
    };
    if (nt <= 1) work(0, 1);
    else {
      std::vector<std::thread> th;
      for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t, nt);
      work(0, nt);
      for (auto& x : th) x.join();
    }
  }
  std::vector<int> offsets{0};
  std::vector<float> xs, ys;
  std::vector<uint8_t> kinds;
  for (size_t fi = 0; fi < finals.size(); ++fi) {
    const Ref& r = finals[fi];
    if (!r.set) continue;
    const int kind = simp[fi].kind;
    if (kind == 0) continue;
    const std::vector<Pt>& poly = simp[fi].poly;
    if (kind == 1) {
      for (const Pt& p : poly) { xs.push_back(static_cast<float>(p.x)); ys.push_back(static_cast<float>(p.y)); }
      xs.push_back(static_cast<float>(poly[0].x)); ys.push_back(static_cast<float>(poly[0].y));  // closed (:379-384)
    } else {
      const Pt* c = r.set->pts + r.set->off[r.idx];
      const int cn = static_cast<int>(r.set->off[r.idx + 1] - r.set->off[r.idx]);
      for (int i = 0; i < cn; ++i) { xs.push_back(static_cast<float>(c[i].x)); ys.push_back(static_cast<float>(c[i].y)); }
    }
    kinds.push_back(kind == 1 ? 0 : 2);
    offsets.push_back(static_cast<int>(xs.size()));
  }
  lap("simplify (host)");
  out->n_polys = static_cast<int>(kinds.size());
  out->n_points = static_cast<int>(xs.size());
  out->offsets = static_cast<int32_t*>(malloc(sizeof(int32_t) * offsets.size()));
  out->xs = static_cast<float*>(malloc(sizeof(float) * std::max<size_t>(xs.size(), 1)));
  out->ys = static_cast<float*>(malloc(sizeof(float) * std::max<size_t>(ys.size(), 1)));
  out->is_float = static_cast<uint8_t*>(malloc(std::max<size_t>(kinds.size(), 1)));
  BD_CHECK(out->offsets && out->xs && out->ys && out->is_float, "out of host memory");
  memcpy(out->offsets, offsets.data(), sizeof(int32_t) * offsets.size());
  if (!xs.empty()) { memcpy(out->xs, xs.data(), sizeof(float) * xs.size()); memcpy(out->ys, ys.data(), sizeof(float) * ys.size()); }
  if (!kinds.empty()) memcpy(out->is_float, kinds.data(), kinds.size());
  return 0;
}

// host-side geometry exposed for CPU differential tests against cv2 (no GPU involved)
double bd_host_contour_area(const int32_t* xy, int n) { return cont::contour_area(reinterpret_cast<const cont::Pt*>(xy), n); }
double bd_host_arc_length(const int32_t* xy, int n) { return cont::arc_length_closed(reinterpret_cast<const cont::Pt*>(xy), n); }
int bd_host_approx_poly(const int32_t* xy, int n, double eps, int32_t* out_xy) {
  std::vector<cont::Pt> dst;
  cont::approx_poly_closed(reinterpret_cast<const cont::Pt*>(xy), n, eps, dst);
  memcpy(out_xy, dst.data(), sizeof(cont::Pt) * dst.size());
  return static_cast<int>(dst.size());
}
int bd_host_simplify(const int32_t* xy, int n, int32_t* out_xy, int* out_n) {
  std::vector<cont::Pt> dst;
  const post::PostConstants K;
  const int kind = cont::simplify(reinterpret_cast<const cont::Pt*>(xy), n, dst, K);
  *out_n = kind == 1 ? static_cast<int>(dst.size()) : 0;
  if (kind == 1) memcpy(out_xy, dst.data(), sizeof(cont::Pt) * dst.size());
  return kind;
}

}  // extern "C"
