// ccl.cuh -- connected-component labelling and the mask primitives shared by the fusion (model_fuse.py)
// and contour (edge_3.py) stages.  All kernels are HBM-bound integer/byte work: one thread per pixel,
// consecutive threads on consecutive pixels of a row (coalesced), grid-stride loops.
//
// Labels: int32 per pixel, -1 = not in the set, otherwise the raster index of the component's FIRST pixel in
// raster order (union-find with atomicMin links the larger root under the smaller).  That root is exactly the
// pixel cv::findContours starts a contour from, and reverse root order is the order it returns them in.
#pragma once
#include "common.cuh"

namespace bd {
namespace ccl {

constexpr int TPB = 256;

// find with path halving: every visited node is re-pointed at its grandparent.  Labels only ever decrease towards
// the root, so the racy plain stores are benign (any stored value is an ancestor).
__device__ __forceinline__ int find_root(int* L, int i) {
  int p = L[i];
  while (p != i) {
    const int g = L[p];
    if (g != p) L[i] = g;
    i = p;
    p = g;
  }
  return i;
}
__device__ __forceinline__ int find_root_ro(const int* L, int i) {  // read-only walk
  int p = L[i];
  while (p != i) {
    i = p;
    p = L[i];
  }
  return i;
}
__device__ __forceinline__ void unite(int* L, int a, int b) {
  while (true) {
    a = find_root(L, a);
    b = find_root(L, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }  // a > b: hang a under b
    const int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

// Run-based labelling.  Pass 1: one warp per image row labels every pixel of the set with the index of the first
// pixel of its horizontal run (ballot + carry, 32 pixels per step); unions then only happen where runs of
// adjacent rows start to touch, O(#run adjacencies) instead of O(#pixels), and the flatten walks run starts only.
// set = pixels with (m != 0) == (fg != 0).
static __global__ void __launch_bounds__(TPB) init_runs(const uint8_t* __restrict__ m, int* __restrict__ L, int H, int W, int fg) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = TPB / 32;
  for (int y = blockIdx.x * warps_per_block + (threadIdx.x >> 5); y < H; y += gridDim.x * warps_per_block) {
    const size_t row = static_cast<size_t>(y) * W;
    int carry = -1;  // run start (pixel index) of a run reaching the previous chunk's last pixel, else -1
    for (int x0 = 0; x0 < W; x0 += 32) {
      const int x = x0 + lane;
      const bool in = x < W && ((m[row + x] != 0) == (fg != 0));
      const unsigned mask = __ballot_sync(0xffffffffu, in);
      int lab = -1;
      if (in) {
        const unsigned zeros_below = ~mask & ((1u << lane) - 1u);
        if (zeros_below == 0) lab = carry >= 0 ? carry : static_cast<int>(row) + x0;
        else lab = static_cast<int>(row) + x0 + (32 - __clz(zeros_below));
      }
      if (x < W) L[row + x] = lab;
      carry = __shfl_sync(0xffffffffu, lab, 31);  // -1 if the chunk's last pixel is not in the set
    }
  }
}
// unions between runs of adjacent rows; CONN8: also diagonal contact
template <bool CONN8>
static __global__ void __launch_bounds__(TPB) merge_runs(int* L, int H, int W) {
  const size_t n = static_cast<size_t>(H) * W;
  for (size_t i = static_cast<size_t>(W) + blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * TPB) {
    const int lp = L[i];
    if (lp < 0) continue;
    const int x = static_cast<int>(i % W);
    const int up = L[i - W];
    const int ul = x > 0 ? L[i - W - 1] : -1;
    const bool start = x == 0 || L[i - 1] < 0;  // first pixel of its run (signs never change; L[i] itself may
                                                 // already point at another run's start after a concurrent union)
    if (up >= 0) {
      if (start || ul < 0) unite(L, lp, up);  // where the lower run or the upper run begins inside the overlap
    } else if (CONN8) {
      if (ul >= 0 && start) unite(L, lp, ul);  // upper run ends diagonally before this run starts
      if (x + 1 < W) {
        const int ur = L[i - W + 1];
        if (ur >= 0 && L[i + 1] < 0) unite(L, lp, ur);  // upper run starts diagonally after this run ends
      }
    }
  }
}
// Flatten in three passes.  (1) every run start walks to its root with path halving: racy, leaves a shallow but
// not necessarily flat forest (a halving store may land after another thread's final store, so the result of
// this pass is not used); (2) every run start takes its root by a read-only walk (all concurrent stores write
// roots or leave ancestors, both valid for readers); (3) every other pixel takes its run start's root.
static __global__ void __launch_bounds__(TPB) compress_runs(int* L, size_t n, int W) {
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB) {
    if (L[i] < 0) continue;
    if ((i % W == 0) || L[i - 1] < 0) find_root(L, static_cast<int>(i));
  }
}
static __global__ void __launch_bounds__(TPB) flatten_runs(int* L, size_t n, int W) {
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB) {
    if (L[i] < 0) continue;
    if ((i % W == 0) || L[i - 1] < 0) L[i] = find_root_ro(L, static_cast<int>(i));
  }
}
static __global__ void __launch_bounds__(TPB) flatten_pixels(int* L, size_t n, int W) {
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB) {
    const int l = L[i];
    if (l < 0) continue;
    const bool start = (i % W == 0) || L[i - 1] < 0;
    if (!start) L[i] = L[l];  // l is this pixel's run start (a run start never has an in-set left neighbour)
  }
}

// ---- hole filling: cv::fillPoly / drawContours(FILLED) of an external contour = the component plus every pixel
// it encloses = complement of the background that is 4-connected to the outside of the image.
// Pass 1 (after flatten of the background labels): every background pixel on the image frame marks its root as
// outside by overwriting the root's own label with OUTSIDE.
constexpr int OUTSIDE = -2;
static __global__ void __launch_bounds__(TPB) mark_outside(int* L, int H, int W) {
  const int per = 2 * (H + W);
  for (int t = blockIdx.x * TPB + threadIdx.x; t < per; t += gridDim.x * TPB) {
    int x, y;
    if (t < W) { x = t; y = 0; }
    else if (t < 2 * W) { x = t - W; y = H - 1; }
    else if (t < 2 * W + H) { x = 0; y = t - 2 * W; }
    else { x = W - 1; y = t - 2 * W - H; }
    const size_t i = static_cast<size_t>(y) * W + x;
    const int r = L[i];
    if (r >= 0) L[r] = OUTSIDE;  // benign race: all writers store the same value; non-root frame pixels keep r
  }
}
// out = 255 for set pixels and for background pixels whose region is not outside
static __global__ void __launch_bounds__(TPB) fill_holes(const uint8_t* __restrict__ m, const int* __restrict__ L,
                                                  uint8_t* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB) {
    uint8_t v = 255;
    if (m[i] == 0) {
      const int r = L[i];  // r == OUTSIDE: this pixel is an outside root; else look at the root's label
      v = (r == OUTSIDE || L[r] == OUTSIDE) ? 0 : 255;
    }
    out[i] = v;
  }
}

// ---- polygon area of every component's external contour (cv::contourArea = |shoelace| / 2 over the boundary
// pixel centres) without tracing: the border following of a hole-free component walks the pixel cracks with the
// component on one side, and the step it takes at a grid vertex is determined by the 2x2 pixels around that
// vertex.  Summing cross(p, q) of those steps per component gives twice the signed area; vertices with one, four
// or two diagonal set pixels contribute nothing (same pixel / no crack / out-and-back).  Verified against
// cv2.contourArea in tests/test_post_cpu.py (numpy twin of this kernel).  area2 must be zero at the roots.
static __global__ void __launch_bounds__(TPB) polygon_area2(const int* __restrict__ L, int H, int W,
                                                     long long* __restrict__ area2) {
  const size_t nv = static_cast<size_t>(H + 1) * (W + 1);
  for (size_t t = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; t < nv; t += static_cast<size_t>(gridDim.x) * TPB) {
    const int x = static_cast<int>(t % (W + 1)), y = static_cast<int>(t / (W + 1));
    // a=(x-1,y-1) b=(x,y-1) c=(x-1,y) d=(x,y)
    const int la = (x > 0 && y > 0) ? L[static_cast<size_t>(y - 1) * W + x - 1] : -1;
    const int lb = (x < W && y > 0) ? L[static_cast<size_t>(y - 1) * W + x] : -1;
    const int lc = (x > 0 && y < H) ? L[static_cast<size_t>(y) * W + x - 1] : -1;
    const int ld = (x < W && y < H) ? L[static_cast<size_t>(y) * W + x] : -1;
    const int code = (la >= 0) | ((lb >= 0) << 1) | ((lc >= 0) << 2) | ((ld >= 0) << 3);
    // step p -> q with p, q in {a,b,c,d}: encoded as (p index << 2 | q index), 0xFF = none
    int p, q, lab;
    switch (code) {
      case 0x3: p = 0; q = 1; lab = la; break;  // a,b   : a -> b
      case 0xC: p = 3; q = 2; lab = lc; break;  // c,d   : d -> c
      case 0x5: p = 2; q = 0; lab = la; break;  // a,c   : c -> a
      case 0xA: p = 1; q = 3; lab = lb; break;  // b,d   : b -> d
      case 0x7: p = 2; q = 1; lab = la; break;  // a,b,c : c -> b
      case 0xB: p = 0; q = 3; lab = la; break;  // a,b,d : a -> d
      case 0xD: p = 3; q = 0; lab = la; break;  // a,c,d : d -> a
      case 0xE: p = 1; q = 2; lab = lb; break;  // b,c,d : b -> c
      default: continue;
    }
    const long long px = x - 1 + (p & 1), py = y - 1 + (p >> 1);
    const long long qx = x - 1 + (q & 1), qy = y - 1 + (q >> 1);
    atomicAdd(reinterpret_cast<unsigned long long*>(area2 + lab), static_cast<unsigned long long>(px * qy - qx * py));
  }
}

// ---- 1 x K / K x 1 erosion (K odd) of a binary mask; pixels outside the image count as set (cv::erode's default
// border value), so a component touching the frame is not eroded from that side.
static __global__ void __launch_bounds__(TPB) erode_line(const uint8_t* __restrict__ m, uint8_t* __restrict__ out, int H,
                                                  int W, int half, int vertical) {
  const size_t n = static_cast<size_t>(H) * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB) {
    uint8_t v = m[i];
    if (v) {
      const int x = static_cast<int>(i % W), y = static_cast<int>(i / W);
      if (vertical) {
        const int y0 = max(0, y - half), y1 = min(H - 1, y + half);
        for (int yy = y0; yy <= y1 && v; ++yy) v = m[static_cast<size_t>(yy) * W + x] ? 255 : 0;
      } else {
        const int x0 = max(0, x - half), x1 = min(W - 1, x + half);
        const uint8_t* row = m + static_cast<size_t>(y) * W;
        for (int xx = x0; xx <= x1 && v; ++xx) v = row[xx] ? 255 : 0;
      }
    }
    out[i] = v ? 255 : 0;
  }
}

// roots get their accumulators zeroed (cheaper than a memset of the scene-sized arrays)
static __global__ void __launch_bounds__(TPB) zero_at_roots(const int* __restrict__ L, size_t n, long long* a0, int* c0,
                                                            int* c1, int* c2, int* c3) {
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB)
    if (L[i] == static_cast<int>(i)) {
      if (a0) a0[i] = 0;
      if (c0) c0[i] = 0;
      if (c1) c1[i] = 0;
      if (c2) c2[i] = 0;
      if (c3) c3[i] = 0;
    }
}
// keep[p] = 255 for set pixels whose component's polygon area exceeds thr2/2 (strict: is at least thr2/2)
static __global__ void __launch_bounds__(TPB) drop_small(const int* __restrict__ L, const long long* __restrict__ area2,
                                                         long long thr2, uint8_t* __restrict__ keep, size_t n, int strict) {
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB) {
    const int r = L[i];
    uint8_t v = 0;
    if (r >= 0) {
      const long long a = llabs(area2[r]);
      v = (strict ? a >= thr2 : a > thr2) ? 255 : 0;
    }
    keep[i] = v;
  }
}

}  // namespace ccl
}  // namespace bd
