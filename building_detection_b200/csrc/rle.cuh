// rle.cuh -- bit-packed masks and run-based connected components for the fusion (model_fuse.py) and contour
// (edge_3.py) stages.
//
// Layout.  A mask is a *plane*: 1 bit per pixel, 32 pixels per word (bit j of word wd of row y = pixel (32 wd + j, y)),
// rows padded to a multiple of 4 words; bits beyond the image width are always zero.  A 20 000^2 scene is 50 MB per
// plane, so every pass below runs out of the 126 MB L2 instead of moving 400 MB (u8) or 1.6 GB (int32 labels) of HBM.
//
// Runs.  The unit of labelling is a maximal horizontal run of set pixels.  Runs are numbered in raster order (row by
// row, left to right); `wprefix[y][wd]` holds the number of the first run that STARTS in word wd of row y, so the run
// covering any set pixel is  wprefix + popc(starts at or before the pixel) - 1  -- no per-run geometry is stored.
// Union-find works on run numbers (atomicMin: the root is the component's first run in raster order, whose first
// pixel is the pixel cv::findContours starts the component's contour from).  All kernels are one thread per word
// (consecutive threads on consecutive words of a row) or one warp per row.
#pragma once
#ifdef BD_HOST_EMUL
#include "host_emul.h"  // tests/emul: the same kernels compiled for the CPU
#else
#include "common.cuh"
#endif

namespace bd {
namespace rle {

constexpr int TPB = 256;

// ---- union-find on run numbers.  find with path halving: every visited node is re-pointed at its grandparent.
// Parents only ever decrease towards the root, so the racy plain stores are benign (any stored value is an ancestor).
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
    if (a < b) { const int t = a; a = b; b = t; }  // a > b: hang a under b (the root is the first run in raster order)
    const int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

struct Plane {
  uint32_t* w;
  int H, W, wp;  // wp: words per row
};
struct RunSet {
  Plane p;
  uint32_t* wprefix;  // [H][wp]
  int* P;             // [nruns] union-find parent -> after flatten(): root run of the component
  int nruns;
};

inline int words_per_row(int W) { return ((W + 31) / 32 + 3) & ~3; }

__device__ __forceinline__ uint32_t valid_mask(int W, int wd) {
  const int rem = W - wd * 32;
  return rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}
__device__ __forceinline__ uint32_t starts_of(uint32_t cur, uint32_t prev) { return cur & ~((cur << 1) | (prev >> 31)); }
__device__ __forceinline__ uint32_t upto(int j) { return (2u << j) - 1u; }   // bits 0..j (j = 31 -> all ones)
__device__ __forceinline__ uint32_t below(int j) { return (1u << j) - 1u; }  // bits 0..j-1

// number of the run covering the set pixel (x, y)
__device__ __forceinline__ int run_at(const RunSet& r, int y, int x) {
  const int wd = x >> 5, j = x & 31;
  const size_t o = static_cast<size_t>(y) * r.p.wp + wd;
  const uint32_t cur = r.p.w[o], prev = wd ? r.p.w[o - 1] : 0u;
  return static_cast<int>(r.wprefix[o]) + __popc(starts_of(cur, prev) & upto(j)) - 1;
}

// f(run number, bit mask of the run's pixels inside this word, first bit) for every run that has pixels in the word
template <class F>
__device__ __forceinline__ void for_runs(uint32_t cur, uint32_t prev, uint32_t pref, F f) {
  const uint32_t st = starts_of(cur, prev);
  uint32_t rem = cur;
  int k = 0;
  while (rem) {
    const int j = __ffs(rem) - 1;
    const uint32_t inv = ~(rem >> j);
    const int len = inv ? __ffs(inv) - 1 : 32;
    const uint32_t mask = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << j;
    const int rid = ((st >> j) & 1u) ? static_cast<int>(pref) + k++ : static_cast<int>(pref) - 1;  // else: continues from the previous word
    f(rid, mask, j);
    rem &= ~mask;
  }
}

// ---------------------------------------------------------------------------------------------- u8 <-> bits
// nonzero bytes of a 32-bit word -> 4 bits
__device__ __forceinline__ uint32_t nz4(uint32_t q) {
  const uint32_t h = (((q & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | q) & 0x80808080u;
  return (((h >> 7) * 0x00204081u) >> 21) & 0xFu;
}
// plane row y = (src row y != 0); VEC: rows are 16-byte aligned (W % 16 == 0 and aligned base)
template <bool VEC>
static __global__ void __launch_bounds__(TPB) pack_u8(const uint8_t* __restrict__ src, Plane p) {
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
    const int x0 = wd * 32;
    uint32_t out = 0;
    if (x0 < p.W) {
      const uint8_t* s = src + static_cast<size_t>(y) * p.W + x0;
      if (VEC && x0 + 32 <= p.W) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(s)), b = __ldg(reinterpret_cast<const uint4*>(s) + 1);
        out = nz4(a.x) | (nz4(a.y) << 4) | (nz4(a.z) << 8) | (nz4(a.w) << 12) | (nz4(b.x) << 16) | (nz4(b.y) << 20) |
              (nz4(b.z) << 24) | (nz4(b.w) << 28);
      } else {
        const int n = min(32, p.W - x0);
        for (int j = 0; j < n; ++j) out |= (s[j] ? 1u : 0u) << j;
      }
    }
    p.w[i] = out;
  }
}
template <bool VEC>
static __global__ void __launch_bounds__(TPB) unpack_u8(Plane p, uint8_t* __restrict__ dst) {
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
    const int x0 = wd * 32;
    if (x0 >= p.W) continue;
    const uint32_t v = p.w[i];
    uint8_t* d = dst + static_cast<size_t>(y) * p.W + x0;
    if (VEC && x0 + 32 <= p.W) {
      auto ex = [](uint32_t nib) {  // 4 bits -> 4 bytes of 0 / 255
        const uint32_t s = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
        return s * 255u;
      };
      uint4 a, b;
      a.x = ex(v & 15u); a.y = ex((v >> 4) & 15u); a.z = ex((v >> 8) & 15u); a.w = ex((v >> 12) & 15u);
      b.x = ex((v >> 16) & 15u); b.y = ex((v >> 20) & 15u); b.z = ex((v >> 24) & 15u); b.w = ex(v >> 28);
      reinterpret_cast<uint4*>(d)[0] = a;
      reinterpret_cast<uint4*>(d)[1] = b;
    } else {
      const int n = min(32, p.W - x0);
      for (int j = 0; j < n; ++j) d[j] = ((v >> j) & 1u) ? 255 : 0;
    }
  }
}

// dst = ~src inside the image (the background plane of the hole fill)
static __global__ void __launch_bounds__(TPB) complement(Plane src, Plane dst) {
  const size_t total = static_cast<size_t>(src.H) * src.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB)
    dst.w[i] = ~src.w[i] & valid_mask(src.W, static_cast<int>(i % src.wp));
}

// >= 3 of 5 planes (model_fuse.py:315-324) with a bit-sliced counter
static __global__ void __launch_bounds__(TPB) vote3of5(const uint32_t* __restrict__ planes5, size_t words, uint32_t* __restrict__ out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < words; i += static_cast<size_t>(gridDim.x) * TPB) {
    uint32_t ones = 0, twos = 0, fours = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const uint32_t x = planes5[k * words + i];
      const uint32_t c1 = ones & x;
      ones ^= x;
      const uint32_t c2 = twos & c1;
      twos ^= c1;
      fours |= c2;
    }
    out[i] = fours | (twos & ones);
  }
}

// ---------------------------------------------------------------------------------------------- run numbering
#ifndef BD_HOST_EMUL  // warp-cooperative / block-cooperative: restated in tests/emul/rle_emul.cpp
static __global__ void __launch_bounds__(TPB) count_rows(Plane p, int* __restrict__ rowcount) {
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (TPB / 32);
  for (int y = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); y < p.H; y += nwarps) {
    const uint32_t* row = p.w + static_cast<size_t>(y) * p.wp;
    int c = 0;
    for (int wd = lane; wd < p.wp; wd += 32) c += __popc(starts_of(row[wd], wd ? row[wd - 1] : 0u));
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) rowcount[y] = c;
  }
}
// exclusive scan of rowcount[0..H) in place (one block), total -> *total_out
static __global__ void __launch_bounds__(1024) scan_rows(int* __restrict__ rows, int H, int* __restrict__ total_out) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  const int per = (H + 1023) / 1024;
  const int b = t * per, e = min(H, b + per);
  int s = 0;
  for (int i = b; i < e; ++i) s += rows[i];
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int run = part[t] - s;  // exclusive prefix of this thread's chunk
  for (int i = b; i < e; ++i) {
    const int c = rows[i];
    rows[i] = run;
    run += c;
  }
  if (t == 1023) *total_out = part[1023];
}
static __global__ void __launch_bounds__(TPB) emit_prefix(Plane p, const int* __restrict__ rowbase, uint32_t* __restrict__ wprefix,
                                                   int* __restrict__ P) {
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (TPB / 32);
  for (int y = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); y < p.H; y += nwarps) {
    const uint32_t* row = p.w + static_cast<size_t>(y) * p.wp;
    int base = rowbase[y];
    for (int wd0 = 0; wd0 < p.wp; wd0 += 32) {
      const int wd = wd0 + lane;
      const int c = wd < p.wp ? __popc(starts_of(row[wd], wd ? row[wd - 1] : 0u)) : 0;
      int incl = c;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      if (wd < p.wp) {
        const int first = base + incl - c;
        wprefix[static_cast<size_t>(y) * p.wp + wd] = static_cast<uint32_t>(first);
        for (int k = 0; k < c; ++k) P[first + k] = first + k;
      }
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

#endif  // BD_HOST_EMUL

// Unions between the runs of adjacent rows.  Two runs touch when their extents overlap (4-connectivity) or overlap after
// growing one of them by a pixel on each side (8-connectivity).  Every touching pair is found from the run that starts
// LATER: a lower run starting at s touches the upper run covering s (or s-1 with CONN8); an upper run starting at t
// touches the lower run covering t (or t-1).  So each run start issues at most one union.
template <bool CONN8>
static __global__ void __launch_bounds__(TPB) merge_rows(RunSet r) {
  const Plane& p = r.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = p.wp + blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const int wd = static_cast<int>(i % p.wp);
    const uint32_t L = p.w[i], U = p.w[i - p.wp];
    const uint32_t Lp = wd ? p.w[i - 1] : 0u, Up = wd ? p.w[i - p.wp - 1] : 0u;
    if (!L && !U) continue;  // (one of them empty: a start at bit 0 may still touch a run ending in the previous word)
    const uint32_t Ls = starts_of(L, Lp), Us = starts_of(U, Up);
    const uint32_t Ucov = CONN8 ? (U | (U << 1) | (Up >> 31)) : U;
    const uint32_t Lcov = CONN8 ? (L | (L << 1) | (Lp >> 31)) : L;
    const int prefL = static_cast<int>(r.wprefix[i]), prefU = static_cast<int>(r.wprefix[i - p.wp]);
    uint32_t h1 = Ls & Ucov;
    while (h1) {
      const int j = __ffs(h1) - 1;
      h1 &= h1 - 1;
      const int lower = prefL + __popc(Ls & below(j));
      int upper;
      if ((U >> j) & 1u) upper = prefU + __popc(Us & upto(j)) - 1;
      else if (j > 0) upper = prefU + __popc(Us & upto(j - 1)) - 1;
      else upper = prefU - 1;  // the run that reaches the last pixel of the previous word
      unite(r.P, lower, upper);
    }
    uint32_t h2 = Us & Lcov;
    while (h2) {
      const int j = __ffs(h2) - 1;
      h2 &= h2 - 1;
      const int upper = prefU + __popc(Us & below(j));
      int lower;
      if ((L >> j) & 1u) lower = prefL + __popc(Ls & upto(j)) - 1;
      else if (j > 0) lower = prefL + __popc(Ls & upto(j - 1)) - 1;
      else lower = prefL - 1;
      unite(r.P, lower, upper);
    }
  }
}
static __global__ void __launch_bounds__(TPB) compress_runs(int* P, int n) {
  for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) find_root(P, i);
}
static __global__ void __launch_bounds__(TPB) flatten_runs(int* P, int n) {
  for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) P[i] = find_root_ro(P, i);
}

// ---------------------------------------------------------------------------------------------- hole fill
// background runs (4-connected labelling of the complement) that reach the image frame mark their component
static __global__ void __launch_bounds__(TPB) mark_outside(RunSet bg, uint8_t* __restrict__ outside) {
  const Plane& p = bg.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  const int wl = (p.W - 1) >> 5, jl = (p.W - 1) & 31;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const uint32_t cur = p.w[i];
    if (!cur) continue;
    const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
    const uint32_t prev = wd ? p.w[i - 1] : 0u;
    if (y == 0 || y == p.H - 1) {
      for_runs(cur, prev, bg.wprefix[i], [&](int rid, uint32_t, int) { outside[bg.P[rid]] = 1; });
    } else {
      if (wd == 0 && (cur & 1u)) outside[bg.P[bg.wprefix[i]]] = 1;
      if (wd == wl && ((cur >> jl) & 1u)) outside[bg.P[static_cast<int>(bg.wprefix[i]) + __popc(starts_of(cur, prev) & upto(jl)) - 1]] = 1;
    }
  }
}
// filled = fg | background regions that do not reach the frame (cv::fillPoly / drawContours(FILLED) of every
// external contour)
static __global__ void __launch_bounds__(TPB) fill_holes(Plane fg, RunSet bg, const uint8_t* __restrict__ outside, Plane filled) {
  const Plane& p = bg.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const uint32_t cur = p.w[i];
    uint32_t out = fg.w[i];
    if (cur) {
      const uint32_t prev = (i % p.wp) ? p.w[i - 1] : 0u;
      for_runs(cur, prev, bg.wprefix[i], [&](int rid, uint32_t mask, int) {
        if (!outside[bg.P[rid]]) out |= mask;
      });
    }
    filled.w[i] = out;
  }
}

// ---------------------------------------------------------------------------------------------- polygon area
// 2 x signed area of every component's external contour (cv::contourArea = |shoelace| / 2 over the boundary pixel
// centres) without tracing: the border following of a hole-free component walks the pixel cracks with the component on
// one side, and the step it takes at a grid vertex is determined by the 2x2 pixels a b / c d around that vertex --
//   a,b: a->b   c,d: d->c   a,c: c->a   b,d: b->d   a,b,c: c->b   a,b,d: a->d   a,c,d: d->a   b,c,d: b->c
// Summing cross(p, q) of those steps per component gives twice the signed area; vertices with one, four or two
// diagonal set pixels contribute nothing (same pixel / no crack / out-and-back).  Verified against cv2.contourArea in
// tests/test_post_cpu.py (numpy twin) and tests/test_rle_emul.py (this kernel compiled for the host).
// One thread per word of VERTICES: vertex (x, v) sits between pixel rows v-1 / v and columns x-1 / x.
// area2 is indexed by root run and must be zero on entry.
static __global__ void __launch_bounds__(TPB) polygon_area2(RunSet r, long long* __restrict__ area2) {
  const Plane& p = r.p;
  const int wv = p.wp + 1;  // words of vertices per row (vertex x = W may need one more word)
  const size_t total = static_cast<size_t>(p.H + 1) * wv;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const int v = static_cast<int>(i / wv), wd = static_cast<int>(i % wv);
    auto ld = [&](int y, int w_) -> uint32_t {
      return (y >= 0 && y < p.H && w_ >= 0 && w_ < p.wp) ? p.w[static_cast<size_t>(y) * p.wp + w_] : 0u;
    };
    const uint32_t A = ld(v - 1, wd), Ap = ld(v - 1, wd - 1), C = ld(v, wd), Cp = ld(v, wd - 1);
    const uint32_t a = (A << 1) | (Ap >> 31), b = A, c = (C << 1) | (Cp >> 31), d = C;
    // vertices with exactly two edge-adjacent, or three, set pixels take a step
    const uint32_t n1 = a ^ b ^ c ^ d;                                      // odd count
    const uint32_t two_adj = ((a & b & ~c & ~d) | (c & d & ~a & ~b) | (a & c & ~b & ~d) | (b & d & ~a & ~c));
    const uint32_t three = n1 & ((a & b) | (c & d));                         // odd and at least one full row -> 3 set
    uint32_t M = two_adj | three;
    int cur_root = -1;
    long long acc = 0;
    while (M) {
      const int j = __ffs(M) - 1;
      M &= M - 1;
      const int x = wd * 32 + j;
      const int code = ((a >> j) & 1u) | (((b >> j) & 1u) << 1) | (((c >> j) & 1u) << 2) | (((d >> j) & 1u) << 3);
      long long t;
      int ly, lx;  // a set pixel of the component the step belongs to
      switch (code) {
        case 0x3: t = -(v - 1); ly = v - 1; lx = x - 1; break;          // a,b   : a -> b
        case 0xC: t = v; ly = v; lx = x - 1; break;                     // c,d   : d -> c
        case 0x5: t = -(x - 1); ly = v - 1; lx = x - 1; break;          // a,c   : c -> a
        case 0xA: t = x; ly = v - 1; lx = x; break;                     // b,d   : b -> d
        case 0x7: t = 1 - x - v; ly = v - 1; lx = x - 1; break;         // a,b,c : c -> b
        case 0xB: t = x - v; ly = v - 1; lx = x - 1; break;             // a,b,d : a -> d
        case 0xD: t = v - x; ly = v - 1; lx = x - 1; break;             // a,c,d : d -> a
        default:  t = x + v - 1; ly = v - 1; lx = x; break;             // 0xE b,c,d : b -> c
      }
      const int root = r.P[run_at(r, ly, lx)];
      if (root != cur_root) {
        if (cur_root >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(area2 + cur_root), static_cast<unsigned long long>(acc));
        cur_root = root;
        acc = 0;
      }
      acc += t;
    }
    if (cur_root >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(area2 + cur_root), static_cast<unsigned long long>(acc));
  }
}

// out = runs of r whose component's polygon area exceeds thr2 / 2 (strict: is at least thr2 / 2)
static __global__ void __launch_bounds__(TPB) keep_large(RunSet r, const long long* __restrict__ area2, long long thr2, int strict,
                                                  Plane out) {
  const Plane& p = r.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const uint32_t cur = p.w[i];
    uint32_t o = 0;
    if (cur) {
      const uint32_t prev = (i % p.wp) ? p.w[i - 1] : 0u;
      for_runs(cur, prev, r.wprefix[i], [&](int rid, uint32_t mask, int) {
        const long long a = llabs(area2[r.P[rid]]);
        if (strict ? a >= thr2 : a > thr2) o |= mask;
      });
    }
    out.w[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------- morphology
// 1 x (2 half + 1) erosion / dilation, half <= 16: a 64-bit window (16 pixels of the previous word, this word, 16 of the
// next) and log-step shifts.  Erosion: pixels outside the image count as set (cv::erode's default border value), so an
// object touching the frame is not eroded from that side; dilation: outside pixels count as clear.
template <bool ERODE>
static __global__ void __launch_bounds__(TPB) morph_h(Plane src, Plane dst, int half) {
  const size_t total = static_cast<size_t>(src.H) * src.wp;
  const int K = 2 * half + 1;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const int wd = static_cast<int>(i % src.wp);
    const uint32_t vm = valid_mask(src.W, wd);
    if (!vm) { dst.w[i] = 0; continue; }
    auto ld = [&](int w_) -> uint32_t {
      if (w_ < 0 || w_ >= src.wp) return ERODE ? 0xFFFFFFFFu : 0u;
      const uint32_t x = src.w[i - wd + w_];
      return ERODE ? (x | ~valid_mask(src.W, w_)) : x;
    };
    const uint64_t v = (static_cast<uint64_t>(ld(wd - 1)) >> 16) | (static_cast<uint64_t>(ld(wd)) << 16) |
                       (static_cast<uint64_t>(ld(wd + 1)) << 48);
    uint64_t r = v;  // bit i of r = AND / OR of v[i .. i+span-1]
    int span = 1;
    while (span * 2 <= K) {
      r = ERODE ? (r & (r >> span)) : (r | (r >> span));
      span *= 2;
    }
    if (span < K) r = ERODE ? (r & (r >> (K - span))) : (r | (r >> (K - span)));
    dst.w[i] = static_cast<uint32_t>(r >> (16 - half)) & vm;
  }
}
// (2 half + 1) x 1
template <bool ERODE>
static __global__ void __launch_bounds__(TPB) morph_v(Plane src, Plane dst, int half) {
  const size_t total = static_cast<size_t>(src.H) * src.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const int y = static_cast<int>(i / src.wp);
    const int y0 = max(0, y - half), y1 = min(src.H - 1, y + half);
    uint32_t r = ERODE ? src.w[i] : 0u;
    if (ERODE) {
      for (int yy = y0; yy <= y1 && r; ++yy) r &= src.w[i + static_cast<ptrdiff_t>(yy - y) * src.wp];
    } else {
      for (int yy = y0; yy <= y1; ++yy) r |= src.w[i + static_cast<ptrdiff_t>(yy - y) * src.wp];
    }
    dst.w[i] = r;
  }
}
static __global__ void __launch_bounds__(TPB) or3(Plane a, Plane b, Plane c, Plane out) {
  const size_t total = static_cast<size_t>(a.H) * a.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB)
    out.w[i] = a.w[i] | b.w[i] | c.w[i];
}

// ---------------------------------------------------------------------------------------------- split decision
// model_fuse.py:65-117.  For every fragment (component of the eroded plane e): its object is the component of `obj`
// that contains it; cnt[object]++ and, when the fragment's polygon area exceeds thr2 / 2, surv[object]++.
static __global__ void __launch_bounds__(TPB) count_fragments(RunSet e, const long long* __restrict__ area2e, RunSet obj,
                                                       int* __restrict__ cnt, int* __restrict__ surv, long long thr2) {
  const Plane& p = e.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const uint32_t cur = p.w[i];
    if (!cur) continue;
    const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
    uint32_t st = starts_of(cur, wd ? p.w[i - 1] : 0u);
    int rid = static_cast<int>(e.wprefix[i]);
    while (st) {
      const int j = __ffs(st) - 1;
      st &= st - 1;
      if (e.P[rid] == rid) {  // the fragment's first run
        const int parent = obj.P[run_at(obj, y, wd * 32 + j)];
        atomicAdd(cnt + parent, 1);
        if (llabs(area2e[rid]) > thr2) atomicAdd(surv + parent, 1);
      }
      ++rid;
    }
  }
}
// eroede_dilate_process (model_fuse.py:173-218) per object from the fragment counts of its two erosions:
//   0 = dropped (a split found fragments but erased them all: `False`), 1 = kept whole (both splits `None`),
//   2 = replaced by its surviving fragments (possibly none: an empty list drops the object)
__device__ __forceinline__ int object_decision(int ch, int sh, int cv_, int sv) {
  const bool falseH = (ch != 1) && (sh < ch) && (sh == 0);
  const bool falseV = (cv_ != 1) && (sv < cv_) && (sv == 0);
  if (falseH || falseV) return 0;
  if (ch == 1 && cv_ == 1) return 1;
  return 2;
}
// whole = runs of kept objects (obj restricted to `keep`) whose decision is "kept whole"
static __global__ void __launch_bounds__(TPB) raster_whole(RunSet obj, Plane keep, const int* __restrict__ cntH, const int* __restrict__ survH,
                                                    const int* __restrict__ cntV, const int* __restrict__ survV, Plane whole) {
  const Plane& p = obj.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const uint32_t kp = keep.w[i];
    uint32_t o = 0;
    if (kp) {
      const uint32_t cur = p.w[i], prev = (i % p.wp) ? p.w[i - 1] : 0u;
      for_runs(cur, prev, obj.wprefix[i], [&](int rid, uint32_t mask, int) {
        if (!(mask & kp)) return;
        const int root = obj.P[rid];
        if (object_decision(cntH[root], survH[root], cntV[root], survV[root]) == 1) o |= mask;
      });
    }
    whole.w[i] = o;
  }
}
// seeds = runs of surviving fragments of objects that are replaced by their fragments in this direction
static __global__ void __launch_bounds__(TPB) raster_seeds(RunSet e, const long long* __restrict__ area2e, RunSet obj,
                                                    const int* __restrict__ cntH, const int* __restrict__ survH,
                                                    const int* __restrict__ cntV, const int* __restrict__ survV, int vertical,
                                                    long long thr2, Plane seeds) {
  const Plane& p = e.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const uint32_t cur = p.w[i];
    uint32_t o = 0;
    if (cur) {
      const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
      const uint32_t prev = wd ? p.w[i - 1] : 0u;
      for_runs(cur, prev, e.wprefix[i], [&](int rid, uint32_t mask, int j) {
        if (llabs(area2e[e.P[rid]]) <= thr2) return;
        const int root = obj.P[run_at(obj, y, wd * 32 + j)];
        const int ch = cntH[root], cv_ = cntV[root];
        if (object_decision(ch, survH[root], cv_, survV[root]) != 2) return;
        if ((vertical ? cv_ : ch) != 1) o |= mask;
      });
    }
    seeds.w[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------- contour stage helpers
// first pixel (raster index) and run number of every component of r whose polygon area passes the threshold
static __global__ void __launch_bounds__(TPB) collect_roots(RunSet r, const long long* __restrict__ area2, long long thr2, int strict,
                                                     int* __restrict__ list_pix, int* __restrict__ list_rid, int* __restrict__ count,
                                                     int cap) {
  const Plane& p = r.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const uint32_t cur = p.w[i];
    if (!cur) continue;
    const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
    uint32_t st = starts_of(cur, wd ? p.w[i - 1] : 0u);
    int rid = static_cast<int>(r.wprefix[i]);
    while (st) {
      const int j = __ffs(st) - 1;
      st &= st - 1;
      if (r.P[rid] == rid) {
        const long long a = llabs(area2[rid]);
        if (strict ? a >= thr2 : a > thr2) {
          const int k = atomicAdd(count, 1);
          if (k < cap) { list_pix[k] = y * p.W + wd * 32 + j; list_rid[k] = rid; }
        }
      }
      ++rid;
    }
  }
}
// ---------------------------------------------------------------------------------------------- parallel border following
// cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) without a serial walk.  A *crack* is a pixel edge between a set
// pixel and a clear one (or the frame), directed so that the set pixel lies on its left: left cracks run down, bottom
// cracks right, right cracks up, top cracks left (counter-clockwise in image coordinates, the direction OpenCV's
// border follower takes from the first raster pixel).  Every crack has exactly one successor, decided by the 2x2
// pixels around its end vertex -- turn towards a diagonal neighbour first (8-connectivity), else go straight, else turn
// around the own pixel's corner -- so the cracks of a component's outer border form one cycle.  The contour OpenCV
// returns is the sequence of the cracks' pixels along that cycle, starting at the left crack of the component's first
// raster pixel, with consecutive repetitions of a pixel collapsed (cyclically).  [tests/test_rle_emul.py walks this
// rule against cv2.findContours; tests/test_post_gpu.py checks the device result point for point.]
// The cycle is cut in front of the start crack and ranked by pointer jumping, carrying the number of points emitted
// from a crack to the end of its list, so every crack knows where its point goes.
//
// Numbering: cbase[y][wd] = number of the first crack of word wd of row y; inside a word the left cracks come first
// (by bit), then bottom, right, top.
enum { CRACK_L = 0, CRACK_B = 1, CRACK_R = 2, CRACK_T = 3 };

struct CrackMasks { uint32_t m[4]; };
__device__ __forceinline__ uint32_t plane_word(const Plane& p, int y, int wd) {
  return (y >= 0 && y < p.H && wd >= 0 && wd < p.wp) ? p.w[static_cast<size_t>(y) * p.wp + wd] : 0u;
}
__device__ __forceinline__ bool plane_bit(const Plane& p, int x, int y) {
  return x >= 0 && x < p.W && y >= 0 && y < p.H && ((p.w[static_cast<size_t>(y) * p.wp + (x >> 5)] >> (x & 31)) & 1u);
}
__device__ __forceinline__ CrackMasks crack_masks(const Plane& p, int y, int wd) {
  const uint32_t cur = plane_word(p, y, wd);
  CrackMasks c;
  c.m[CRACK_L] = cur & ~((cur << 1) | (plane_word(p, y, wd - 1) >> 31));
  c.m[CRACK_B] = cur & ~plane_word(p, y + 1, wd);
  c.m[CRACK_R] = cur & ~((cur >> 1) | (plane_word(p, y, wd + 1) << 31));
  c.m[CRACK_T] = cur & ~plane_word(p, y - 1, wd);
  return c;
}
__device__ __forceinline__ int crack_rank(const CrackMasks& c, int t, int j) {  // number inside the word
  int r = __popc(c.m[t] & below(j));
  if (t > CRACK_L) r += __popc(c.m[CRACK_L]);
  if (t > CRACK_B) r += __popc(c.m[CRACK_B]);
  if (t > CRACK_R) r += __popc(c.m[CRACK_R]);
  return r;
}
__device__ __forceinline__ int crack_id(const Plane& p, const uint32_t* __restrict__ cbase, int x, int y, int t) {
  const int wd = x >> 5;
  return static_cast<int>(cbase[static_cast<size_t>(y) * p.wp + wd]) + crack_rank(crack_masks(p, y, wd), t, x & 31);
}
// successor of crack t of pixel (x, y): pixel and type
__device__ __forceinline__ void crack_succ(const Plane& p, int x, int y, int t, int* nx, int* ny, int* nt) {
  switch (t) {
    case CRACK_L:
      if (plane_bit(p, x - 1, y + 1)) { *nx = x - 1; *ny = y + 1; *nt = CRACK_T; }
      else if (plane_bit(p, x, y + 1)) { *nx = x; *ny = y + 1; *nt = CRACK_L; }
      else { *nx = x; *ny = y; *nt = CRACK_B; }
      break;
    case CRACK_B:
      if (plane_bit(p, x + 1, y + 1)) { *nx = x + 1; *ny = y + 1; *nt = CRACK_L; }
      else if (plane_bit(p, x + 1, y)) { *nx = x + 1; *ny = y; *nt = CRACK_B; }
      else { *nx = x; *ny = y; *nt = CRACK_R; }
      break;
    case CRACK_R:
      if (plane_bit(p, x + 1, y - 1)) { *nx = x + 1; *ny = y - 1; *nt = CRACK_B; }
      else if (plane_bit(p, x, y - 1)) { *nx = x; *ny = y - 1; *nt = CRACK_R; }
      else { *nx = x; *ny = y; *nt = CRACK_T; }
      break;
    default:
      if (plane_bit(p, x - 1, y - 1)) { *nx = x - 1; *ny = y - 1; *nt = CRACK_R; }
      else if (plane_bit(p, x - 1, y)) { *nx = x - 1; *ny = y; *nt = CRACK_T; }
      else { *nx = x; *ny = y; *nt = CRACK_L; }
      break;
  }
}
#ifndef BD_HOST_EMUL
static __global__ void __launch_bounds__(TPB) count_row_cracks(Plane p, int* __restrict__ rowcount) {
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (TPB / 32);
  for (int y = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); y < p.H; y += nwarps) {
    int c = 0;
    for (int wd = lane; wd < p.wp; wd += 32) {
      const CrackMasks k = crack_masks(p, y, wd);
      c += __popc(k.m[0]) + __popc(k.m[1]) + __popc(k.m[2]) + __popc(k.m[3]);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) rowcount[y] = c;
  }
}
static __global__ void __launch_bounds__(TPB) emit_crack_base(Plane p, const int* __restrict__ rowbase, uint32_t* __restrict__ cbase) {
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (TPB / 32);
  for (int y = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); y < p.H; y += nwarps) {
    int base = rowbase[y];
    for (int wd0 = 0; wd0 < p.wp; wd0 += 32) {
      const int wd = wd0 + lane;
      int c = 0;
      if (wd < p.wp) {
        const CrackMasks k = crack_masks(p, y, wd);
        c = __popc(k.m[0]) + __popc(k.m[1]) + __popc(k.m[2]) + __popc(k.m[3]);
      }
      int incl = c;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      if (wd < p.wp) cbase[static_cast<size_t>(y) * p.wp + wd] = static_cast<uint32_t>(base + incl - c);
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}
#endif  // BD_HOST_EMUL
// For contour c (first pixel roots[c], in output order): isolated pixel -> npts[c] = 1 and its start crack stays
// unmarked (its four cracks rank into a list nobody reads); otherwise start_of[crack id of the pixel's left crack] =
// c + 1 and start_crack[c] = that id.
static __global__ void __launch_bounds__(TPB) mark_starts(Plane p, const uint32_t* __restrict__ cbase, const int* __restrict__ roots,
                                                   int n, int* __restrict__ start_of, int* __restrict__ start_crack,
                                                   int* __restrict__ npts) {
  for (int c = blockIdx.x * TPB + threadIdx.x; c < n; c += gridDim.x * TPB) {
    const int x = roots[c] % p.W, y = roots[c] / p.W;
    const int id = crack_id(p, cbase, x, y, CRACK_L);
    start_crack[c] = id;
    const bool isolated = !plane_bit(p, x + 1, y) && !plane_bit(p, x - 1, y + 1) && !plane_bit(p, x, y + 1) && !plane_bit(p, x + 1, y + 1);
    if (isolated) npts[c] = 1;
    else { npts[c] = 0; start_of[id] = c + 1; }
  }
}
// per crack: successor (a crack whose successor is a marked start crack ends its list: nxt = itself, term_of = contour
// + 1) and ws = 1 when the crack is the last one of its pixel's visit (the successor sits on another pixel)
static __global__ void __launch_bounds__(TPB) init_cracks(Plane p, const uint32_t* __restrict__ cbase, const int* __restrict__ start_of,
                                                   int* __restrict__ nxt, int* __restrict__ ws, int* __restrict__ term_of) {
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    if (!p.w[i]) continue;
    const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
    const CrackMasks k = crack_masks(p, y, wd);
    int id = static_cast<int>(cbase[i]);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      uint32_t m = k.m[t];
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const int x = wd * 32 + j;
        int nx, ny, nt;
        crack_succ(p, x, y, t, &nx, &ny, &nt);
        const int sid = (ny == y && (nx >> 5) == wd) ? static_cast<int>(cbase[i]) + crack_rank(k, nt, nx & 31) : crack_id(p, cbase, nx, ny, nt);
        const int st = start_of[sid];
        if (st) { nxt[id] = id; ws[id] = 0; term_of[id] = st; }
        else { nxt[id] = sid; ws[id] = (nx != x || ny != y) ? 1 : 0; term_of[id] = 0; }
        ++id;
      }
    }
  }
}
// one pointer-jumping round (double buffered): ws covers [crack, nxt) afterwards [crack, nxt[nxt])
static __global__ void __launch_bounds__(TPB) jump_cracks(const int* __restrict__ nxt, const int* __restrict__ ws, int* __restrict__ nxt2,
                                                   int* __restrict__ ws2, int n) {
  for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
    const int s = nxt[i];
    ws2[i] = ws[i] + ws[s];  // ws of a list end is 0
    nxt2[i] = nxt[s];
  }
}
static __global__ void __launch_bounds__(TPB) contour_totals(const int* __restrict__ start_crack, const int* __restrict__ ws, int n,
                                                      int* __restrict__ npts) {
  for (int c = blockIdx.x * TPB + threadIdx.x; c < n; c += gridDim.x * TPB)
    if (npts[c] == 0) npts[c] = ws[start_crack[c]];  // (1: an isolated pixel, set by mark_starts)
}
// write the points: the crack that ends a pixel visit puts its pixel at off[contour] + (total - points from here on)
static __global__ void __launch_bounds__(TPB) scatter_points(Plane p, const uint32_t* __restrict__ cbase, const int* __restrict__ nxt,
                                                      const int* __restrict__ ws, const int* __restrict__ term_of,
                                                      const int* __restrict__ npts, const long long* __restrict__ off,
                                                      const int* __restrict__ roots, int ncontours, int2* __restrict__ pts) {
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    if (!p.w[i]) continue;
    const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
    const CrackMasks k = crack_masks(p, y, wd);
    int id = static_cast<int>(cbase[i]);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      uint32_t m = k.m[t];
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const int e = nxt[id];      // list end
        const int c = term_of[e];   // contour + 1 (0: a list nobody reads -- isolated pixels, untraced components)
        if (c && e != id) {
          const int x = wd * 32 + j;
          int nx, ny, nt;
          crack_succ(p, x, y, t, &nx, &ny, &nt);
          if (nx != x || ny != y) pts[off[c - 1] + (npts[c - 1] - ws[id])] = make_int2(x, y);
        }
        ++id;
      }
    }
  }
  // isolated pixels: one point
  for (int c = blockIdx.x * TPB + threadIdx.x; c < ncontours; c += gridDim.x * TPB)
    if (off[c + 1] - off[c] == 1 && npts[c] == 1) pts[off[c]] = make_int2(roots[c] % p.W, roots[c] / p.W);
}
// bounding boxes of components from their runs, indexed by root run: bbmin[2 r] = min x, min y (memset 0x7f),
// bbmax[2 r] = max x + 1, max y + 1 (memset 0xff) -- cv::boundingRect's x, y, x + w, y + h
static __global__ void __launch_bounds__(TPB) run_bboxes(RunSet r, int* __restrict__ bbmin, int* __restrict__ bbmax) {
  // one update per RUN, issued where the run starts (x0, row) and where it ends (x1 + 1): a scene-sized component has
  // 2 x H updates instead of one per word (12.5 M same-address atomics cost 30+ ms at 20 000^2)
  const Plane& p = r.p;
  const size_t total = static_cast<size_t>(p.H) * p.wp;
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * TPB) {
    const uint32_t cur = p.w[i];
    if (!cur) continue;
    const int y = static_cast<int>(i / p.wp), wd = static_cast<int>(i % p.wp);
    const uint32_t prev = wd ? p.w[i - 1] : 0u, next = wd + 1 < p.wp ? p.w[i + 1] : 0u;
    const uint32_t st = starts_of(cur, prev);
    const uint32_t en = cur & ~((cur >> 1) | (next << 31));  // last pixels of runs
    for_runs(cur, prev, r.wprefix[i], [&](int rid, uint32_t mask, int j) {
      const size_t o = 2 * static_cast<size_t>(r.P[rid]);
      if (mask & st) {
        atomicMin(bbmin + o, wd * 32 + j);
        atomicMin(bbmin + o + 1, y);
        atomicMax(bbmax + o + 1, y + 1);
      }
      if (mask & en) atomicMax(bbmax + o, wd * 32 + (32 - __clz(mask)));
    });
  }
}
static __global__ void __launch_bounds__(TPB) gather_bboxes(const int* __restrict__ bbmin, const int* __restrict__ bbmax,
                                                     const int* __restrict__ root_runs, int n, int* __restrict__ bbox4) {
  for (int c = blockIdx.x * TPB + threadIdx.x; c < n; c += gridDim.x * TPB) {
    const size_t o = 2 * static_cast<size_t>(root_runs[c]);
    bbox4[4 * c] = bbmin[o]; bbox4[4 * c + 1] = bbmin[o + 1]; bbox4[4 * c + 2] = bbmax[o]; bbox4[4 * c + 3] = bbmax[o + 1];
  }
}

}  // namespace rle
}  // namespace bd
