// kernels.cuh -- CUDA-core kernels of the network forward: the memory-bound ops (depthwise 3x3,
// pooling, nearest-upsample/add, global average pooling, attention gates, SK fusion, softmax head)
// and the small-channel direct convolution (3-channel stems, 1/2-channel gates and heads, BAM C/16
// convs) that is not worth a tensor-core tile.  Feature maps are NHWC; wherever channel counts are
// multiples of 8 the access unit is one 16-byte vector of 8 h16, with threads mapped channel-group
// fastest so that a warp touches consecutive 16-byte vectors.
#pragma once
#include "common.cuh"

namespace bd {
namespace k {

constexpr int TPB = 256;

struct View {  // device-side channel-slice view
  void* base;
  int H, W, ctot, c0, c, f32;
};

__device__ __forceinline__ float ld1(const View& v, size_t pix, int ch) {
  const size_t i = pix * v.ctot + v.c0 + ch;
  return v.f32 ? static_cast<const float*>(v.base)[i] : __half2float(static_cast<const h16*>(v.base)[i]);
}
__device__ __forceinline__ void st1(const View& v, size_t pix, int ch, float x) {
  const size_t i = pix * v.ctot + v.c0 + ch;
  if (v.f32) static_cast<float*>(v.base)[i] = x;
  else static_cast<h16*>(v.base)[i] = to_h16(x);
}
__device__ __forceinline__ void ld8(const View& v, size_t pix, int ch, float* f) {
  unpack8(*reinterpret_cast<const h16x8*>(static_cast<const h16*>(v.base) + pix * v.ctot + v.c0 + ch), f);
}
__device__ __forceinline__ void st8(const View& v, size_t pix, int ch, const float* f) {
  *reinterpret_cast<h16x8*>(static_cast<h16*>(v.base) + pix * v.ctot + v.c0 + ch) = pack8(f);
}
__device__ __forceinline__ float actf(float v, int act) {
  return act == 1 ? fmaxf(v, 0.0f) : (act == 2 ? sigmoidf_(v) : v);
}

// ---------------------------------------------------------------------------------- direct conv
struct DirectParams {
  View x, y, res;  // res.base == nullptr: none
  int N, Ho, Wo, stride, ntaps;
  int dy[18], dx[18];
  int act_pre, act_post, out_scale, out_oy, out_ox;
  const h16* w;  // [ntaps][Cout][Cin]
  const float* bias;
};
constexpr int DC_CO = 4;  // output channels per thread

// thread = (output pixel, group of DC_CO output channels); weights are warp-uniform broadcasts.
__global__ void __launch_bounds__(TPB) conv_direct_kernel(const __grid_constant__ DirectParams p) {
  pdl_prologue();
  const int cog = (p.y.c + DC_CO - 1) / DC_CO;
  const size_t total = static_cast<size_t>(p.N) * p.Ho * p.Wo * cog;
  for (size_t idx = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * TPB) {
    // pixel fastest within a warp so that the weight reads are uniform
    const size_t pix = idx % (static_cast<size_t>(p.N) * p.Ho * p.Wo);
    const int g = static_cast<int>(idx / (static_cast<size_t>(p.N) * p.Ho * p.Wo));
    const int ow = static_cast<int>(pix % p.Wo), oh = static_cast<int>((pix / p.Wo) % p.Ho);
    const int n = static_cast<int>(pix / (static_cast<size_t>(p.Wo) * p.Ho));
    const int co0 = g * DC_CO;
    const int Cin = p.x.c, Cout = p.y.c;
    float acc[DC_CO];
#pragma unroll
    for (int j = 0; j < DC_CO; ++j) acc[j] = 0.0f;
    for (int t = 0; t < p.ntaps; ++t) {
      const int ih = oh * p.stride + p.dy[t], iw = ow * p.stride + p.dx[t];
      if (ih < 0 || ih >= p.x.H || iw < 0 || iw >= p.x.W) continue;
      const size_t ipix = (static_cast<size_t>(n) * p.x.H + ih) * p.x.W + iw;
      const h16* wt = p.w + (static_cast<size_t>(t) * Cout + co0) * Cin;
      for (int ci = 0; ci < Cin; ++ci) {
        const float xv = ld1(p.x, ipix, ci);
#pragma unroll
        for (int j = 0; j < DC_CO; ++j)
          if (co0 + j < Cout) acc[j] = fmaf(xv, __half2float(wt[static_cast<size_t>(j) * Cin + ci]), acc[j]);
      }
    }
    const size_t opix = (static_cast<size_t>(n) * p.y.H + (oh * p.out_scale + p.out_oy)) * p.y.W +
                        (ow * p.out_scale + p.out_ox);
#pragma unroll
    for (int j = 0; j < DC_CO; ++j) {
      if (co0 + j >= Cout) break;
      float v = actf(acc[j] + p.bias[co0 + j], p.act_pre);
      if (p.res.base) v += ld1(p.res, opix, co0 + j);
      st1(p.y, opix, co0 + j, actf(v, p.act_post));
    }
  }
}

// ---------------------------------------------------------------------------------- small-channel conv (CUDA cores)
// Convolutions with at most 16 output channels and a few hundred multiply-adds per pixel: BAM's spatial-gate
// convolutions on C/16 = 4 or 8 channels (3x3, dilation 4; stored padded to 16 channels), its 1-channel gate logits,
// and the 1x1 two-channel heads.  On the tensor-core path these are 128-pixel tiles of N = 16 columns that spend their
// time in per-k-block bookkeeping (bam 3x3 4->4 @256^2: 1.8 TFLOP/s, 17x its HBM time); here they are HBM-bound:
// thread = output pixel, 128-bit loads of 8 input channels, the (tap, ci, co) weights as fp32 in shared memory
// (warp-uniform broadcast reads), only the channels that carry non-zero weights are touched.
struct SmallParams {
  View x, y;
  int N, Ho, Wo, ntaps;
  int dy[9], dx[9];
  int cin_used, cout_used, act;
  const float* w;     // [ntaps][cin_used][CO] fp32
  const float* bias;  // [CO]
};
// LPP lanes share a pixel: lane l of the group takes the 8-channel vectors l, l + LPP, ... so that the LPP loads of a
// pixel are one contiguous 16 x LPP-byte segment (thread-per-pixel with 64 input channels touches 32 different
// 128-byte lines per load instruction: measured 3x slower than the tensor-core tile it replaced), partial sums are
// combined with shuffles.
template <int CO, int LPP>
__global__ void __launch_bounds__(TPB) conv_small_kernel(const __grid_constant__ SmallParams p) {
  extern __shared__ float sw[];
  pdl_trigger();
  const int nw = p.ntaps * p.cin_used * CO;
  for (int i = threadIdx.x; i < nw + CO; i += TPB) sw[i] = i < nw ? p.w[i] : p.bias[i - nw];  // constant data
  __syncthreads();
  pdl_wait();
  constexpr int PPW = 32 / LPP;  // pixels per warp
  // 32-bit index arithmetic (the host checks N * Ho * Wo < 2^31): 64-bit divisions per pixel cost more than the loads
  const unsigned total = static_cast<unsigned>(p.N) * p.Ho * p.Wo;
  const h16* xb = static_cast<const h16*>(p.x.base) + p.x.c0;
  const int lane = threadIdx.x & 31, sub = lane % LPP;
  const unsigned warp0 = (blockIdx.x * TPB + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * TPB) >> 5;
  const bool pointwise = p.ntaps == 1 && p.dy[0] == 0 && p.dx[0] == 0;  // 1x1: input pixel == output pixel
  const unsigned HW = static_cast<unsigned>(p.Ho) * p.Wo;
  auto finish = [&](float* acc, unsigned pix) {  // bias, activation, store (one lane per pixel)
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[j] = actf(acc[j] + sw[nw + j], p.act);
    if (p.y.f32) {
      float* yp = static_cast<float*>(p.y.base) + static_cast<size_t>(pix) * p.y.ctot + p.y.c0;
      if (CO == 2 && p.y.c == 2 && ((p.y.ctot | p.y.c0) & 1) == 0) {
        *reinterpret_cast<float2*>(yp) = make_float2(acc[0], acc[1]);
      } else {
#pragma unroll
        for (int j = 0; j < CO; ++j)
          if (j < p.y.c) yp[j] = acc[j];
      }
    } else {
      // the whole (padded) slice is written: channels beyond cout_used carry zero weights and zero bias
      for (int c0 = 0; c0 < p.y.c; c0 += 8) {
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = 0.0f;
#pragma unroll
        for (int j = 0; j < CO; ++j)
          if (j >= c0 && j < c0 + 8) f[j - c0] = acc[j];
        st8(p.y, pix, c0, f);
      }
    }
  };
  if (pointwise && p.cin_used <= 8 * LPP) {
    // 1x1 with one 16-byte vector per lane (the two-channel heads on 32 / 64 channels, BAM's 64 -> 4 reduce): U pixels
    // per lane group in flight -- one load per thread and iteration left the kernel at a quarter of the HBM rate
    constexpr int U = 4;
    const bool has_vec = sub * 8 < p.cin_used;
    constexpr bool WREG = CO <= 4;         // the lane's weight rows in registers (shared memory for wider outputs)
    float wr[WREG ? 8 : 1][WREG ? CO : 1];
    if (WREG) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int j = 0; j < CO; ++j) wr[WREG ? k : 0][WREG ? j : 0] = (has_vec && sub * 8 + k < p.cin_used) ? sw[(sub * 8 + k) * CO + j] : 0.0f;
    }
    for (unsigned base = warp0 * PPW; base < total; base += nwarps * PPW * U) {  // warp-uniform trip count
      h16x8 xr[U];
      unsigned pixs[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pixs[u] = base + u * nwarps * PPW + lane / LPP;
        if (pixs[u] < total && has_vec)
          xr[u] = *reinterpret_cast<const h16x8*>(xb + static_cast<size_t>(pixs[u]) * p.x.ctot + sub * 8);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool valid = pixs[u] < total;
        float acc[CO];
#pragma unroll
        for (int j = 0; j < CO; ++j) acc[j] = 0.0f;
        if (valid && has_vec) {
          float f[8];
          unpack8(xr[u], f);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (WREG) {
#pragma unroll
              for (int j = 0; j < CO; ++j) acc[j] = fmaf(f[k], wr[WREG ? k : 0][WREG ? j : 0], acc[j]);
            } else if (sub * 8 + k < p.cin_used) {
              const float* wrow = sw + (sub * 8 + k) * CO;
#pragma unroll
              for (int j = 0; j < CO; ++j) acc[j] = fmaf(f[k], wrow[j], acc[j]);
            }
          }
        }
#pragma unroll
        for (int off = LPP / 2; off > 0; off >>= 1) {
#pragma unroll
          for (int j = 0; j < CO; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
        }
        if (valid && sub == 0) finish(acc, pixs[u]);
      }
    }
    return;
  }
  for (unsigned base = warp0 * PPW; base < total; base += nwarps * PPW) {  // warp-uniform trip count (shuffles below)
    const unsigned pix = base + lane / LPP;
    const bool valid = pix < total;
    const unsigned pc = valid ? pix : total - 1;
    float acc[CO];
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[j] = 0.0f;
    if (pointwise) {
      const h16* xp = xb + static_cast<size_t>(pc) * p.x.ctot;
      for (int c0 = sub * 8; c0 < p.cin_used; c0 += 8 * LPP) {
        float f[8];
        unpack8(*reinterpret_cast<const h16x8*>(xp + c0), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (c0 + k < p.cin_used) {
            const float* wr = sw + (c0 + k) * CO;
#pragma unroll
            for (int j = 0; j < CO; ++j) acc[j] = fmaf(f[k], wr[j], acc[j]);
          }
        }
      }
    } else {
      const unsigned n = pc / HW, rem = pc - n * HW;
      const int oh = static_cast<int>(rem / p.Wo), ow = static_cast<int>(rem - oh * p.Wo);
      for (int t = 0; t < p.ntaps; ++t) {
        const int ih = oh + p.dy[t], iw = ow + p.dx[t];
        if (ih < 0 || ih >= p.x.H || iw < 0 || iw >= p.x.W) continue;  // 'same' padding reads zero
        const h16* xp = xb + (static_cast<size_t>(n * p.x.H + ih) * p.x.W + iw) * p.x.ctot;
        const float* wt = sw + t * p.cin_used * CO;
        for (int c0 = sub * 8; c0 < p.cin_used; c0 += 8 * LPP) {
          float f[8];
          unpack8(*reinterpret_cast<const h16x8*>(xp + c0), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (c0 + k < p.cin_used) {
              const float* wr = wt + (c0 + k) * CO;
#pragma unroll
              for (int j = 0; j < CO; ++j) acc[j] = fmaf(f[k], wr[j], acc[j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int off = LPP / 2; off > 0; off >>= 1) {
#pragma unroll
      for (int j = 0; j < CO; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
    }
    if (!valid || sub != 0) continue;
    finish(acc, pix);
  }
}

// ---------------------------------------------------------------------------------- depthwise 3x3
struct DwParams {
  View x, y;
  int N, Ho, Wo, stride, pad_t, pad_l, relu_in;
  const h16* w;  // [9][C] fp16
};
// thread = (VEC-channel group, output column, band of DW_ROWS output rows, image).  The 9 x VEC fp16 weights stay
// packed in registers; the input window is a ring of row slots (3 rows in use + the rows of the NEXT output row,
// whose loads are issued one iteration ahead of their use), fully unrolled so the ring indices are compile-time
// and the loads of row r+1 overlap the FMAs of row r.  The 9-tap sum runs on packed half2 FMAs (fp16 accumulate,
// one rounding per tap, tap order kh-major); oracle/plan_interp.py mirrors this rounding sequence exactly.
// Consecutive threads are consecutive channel groups (coalesced vectors).  VEC = 4 halves the registers per
// thread (twice the resident warps) for the latency-bound small maps.
constexpr int DW_ROWS = 8;
template <int VEC> struct DwVec;
template <> struct DwVec<8> { typedef uint4 T; };
template <> struct DwVec<4> { typedef uint2 T; };
template <int STRIDE, int VEC>
__global__ void __launch_bounds__(TPB) dwconv3x3_kernel(const __grid_constant__ DwParams p) {
  pdl_prologue();
  typedef typename DwVec<VEC>::T V;
  constexpr int NH = VEC / 2;                       // half2 per vector
  constexpr int SLOTS = STRIDE == 1 ? 4 : 5;        // input row r lives in slot r % SLOTS
  union U { V v; __half2 h[NH]; };
  const int C = p.x.c, cg = C / VEC;
  const int rb = (p.Ho + DW_ROWS - 1) / DW_ROWS;
  const unsigned total = static_cast<unsigned>(p.N) * rb * p.Wo * cg;
  const __half2 zero2 = __float2half2_rn(0.0f);
  for (unsigned idx = blockIdx.x * TPB + threadIdx.x; idx < total; idx += gridDim.x * TPB) {
    const int g = static_cast<int>(idx % cg);
    unsigned t = idx / cg;
    const int ow = static_cast<int>(t % p.Wo);
    t /= p.Wo;
    const int ob = static_cast<int>(t % rb), n = static_cast<int>(t / rb);
    U wp[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wp[k].v = *reinterpret_cast<const V*>(p.w + k * C + g * VEC);
    const int iw0 = ow * STRIDE - p.pad_l;
    const int oh0 = ob * DW_ROWS;
    const int ih0 = oh0 * STRIDE - p.pad_t;
    const h16* xb = static_cast<const h16*>(p.x.base) + p.x.c0 + g * VEC +
                    (static_cast<size_t>(n) * p.x.H * p.x.W) * p.x.ctot;
    h16* yb = static_cast<h16*>(p.y.base) + p.y.c0 + g * VEC + (static_cast<size_t>(n) * p.Ho * p.Wo + ow) * p.y.ctot;
    const bool cok0 = iw0 >= 0, cok1 = iw0 + 1 < p.x.W, cok2 = iw0 + 2 < p.x.W;  // iw0 + 1 >= 0 always
    U win[SLOTS][3];
    auto load_row = [&](int rel, U* row) {  // 3 taps of input row ih0 + rel (zero outside the map), ReLU applied once
      const int ih = ih0 + rel;
      const bool rok = ih >= 0 && ih < p.x.H;
      const h16* xr = xb + (static_cast<size_t>(ih) * p.x.W + iw0) * p.x.ctot;
      const V z = {};
      row[0].v = (rok && cok0) ? *reinterpret_cast<const V*>(xr) : z;
      row[1].v = (rok && cok1) ? *reinterpret_cast<const V*>(xr + p.x.ctot) : z;
      row[2].v = (rok && cok2) ? *reinterpret_cast<const V*>(xr + 2 * p.x.ctot) : z;
    };
    auto relu_row = [&](U* row) {
      if (p.relu_in) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int j = 0; j < NH; ++j) row[kw].h[j] = __hmax2(row[kw].h[j], zero2);
      }
    };
    load_row(0, win[0]);
    load_row(1, win[1]);
    load_row(2, win[2]);
    relu_row(win[0]); relu_row(win[1]); relu_row(win[2]);
#pragma unroll
    for (int r = 0; r < DW_ROWS; ++r) {
      const int oh = oh0 + r;
      if (oh >= p.Ho) break;
      // rows of the next output row: STRIDE new ones, into the slots the current row no longer needs after this step
      if (r + 1 < DW_ROWS) {
#pragma unroll
        for (int q = 0; q < STRIDE; ++q) load_row(STRIDE * r + 3 + q, win[(STRIDE * r + 3 + q) % SLOTS]);
      }
      U acc;
#pragma unroll
      for (int j = 0; j < NH; ++j) acc.h[j] = zero2;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int j = 0; j < NH; ++j)
            acc.h[j] = __hfma2(win[(STRIDE * r + kh) % SLOTS][kw].h[j], wp[kh * 3 + kw].h[j], acc.h[j]);
      *reinterpret_cast<V*>(yb + static_cast<size_t>(oh) * p.Wo * p.y.ctot) = acc.v;
      if (r + 1 < DW_ROWS) {
#pragma unroll
        for (int q = 0; q < STRIDE; ++q) relu_row(win[(STRIDE * r + 3 + q) % SLOTS]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------- max pool
// The element-wise kernels below share one shape: a grid-stride loop over 16-byte vectors with 32-bit indices
// (multiply-shift division by launch constants; the host checks that the vector count stays below 2^31) that issues
// every load of an iteration before the first use.  The first versions took three 64-bit divisions per vector
// (~200 instructions for one 16-byte load and store) and were bound by instruction issue at 0.3-0.6 of the HBM peak.
struct PoolParams {
  View x, y;
  int N, Ho, Wo, k, stride, pad_t, pad_l;
  FastDiv fd_cg, fd_wo, fd_ho;
};
// max is exact in any precision: it runs on packed fp16 (stored maps never hold inf: stores saturate)
template <int K>
__global__ void __launch_bounds__(TPB) maxpool_kernel(const __grid_constant__ PoolParams p) {
  pdl_prologue();
  const unsigned cg = static_cast<unsigned>(p.x.c) >> 3;
  const unsigned total = static_cast<unsigned>(p.N) * p.Ho * p.Wo * cg;
  const h16* xb = static_cast<const h16*>(p.x.base) + p.x.c0;
  h16* yb = static_cast<h16*>(p.y.base) + p.y.c0;
  const __half2 ninf = __float2half2_rn(-INFINITY);
  for (unsigned idx = blockIdx.x * TPB + threadIdx.x; idx < total; idx += gridDim.x * TPB) {
    const unsigned pix = fd_div(idx, p.fd_cg), g = idx - pix * cg;
    const unsigned t = fd_div(pix, p.fd_wo), ow = pix - t * p.Wo;
    const unsigned n = fd_div(t, p.fd_ho), oh = t - n * p.Ho;
    const int ih0 = static_cast<int>(oh) * p.stride - p.pad_t, iw0 = static_cast<int>(ow) * p.stride - p.pad_l;
    const h16* xn = xb + static_cast<size_t>(n) * p.x.H * p.x.W * p.x.ctot + g * 8;
    h16x8 v[K * K];
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const int ih = ih0 + kh, iw = iw0 + kw;
        if (ih >= 0 && ih < p.x.H && iw >= 0 && iw < p.x.W)
          v[kh * K + kw] = *reinterpret_cast<const h16x8*>(xn + (static_cast<size_t>(ih) * p.x.W + iw) * p.x.ctot);
        else
          v[kh * K + kw].v[0] = v[kh * K + kw].v[1] = v[kh * K + kw].v[2] = v[kh * K + kw].v[3] = ninf;
      }
    h16x8 m = v[0];
#pragma unroll
    for (int i = 1; i < K * K; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) m.v[j] = __hmax2(m.v[j], v[i].v[j]);
    *reinterpret_cast<h16x8*>(yb + static_cast<size_t>(pix) * p.y.ctot + g * 8) = m;
  }
}

// ---------------------------------------------------------------------------------- add-N with nearest upsample
struct AddnParams {
  View x[4], y;
  int sh[4];  // log2 of the nearest-upsampling factor of input i
  int n_in, N, act;
  FastDiv fd_cg, fd_w, fd_h;
};
__global__ void __launch_bounds__(TPB) addn_kernel(const __grid_constant__ AddnParams p) {
  pdl_prologue();
  constexpr int U = 2;
  const unsigned cg = static_cast<unsigned>(p.y.c) >> 3;
  const unsigned total = static_cast<unsigned>(p.N) * p.y.H * p.y.W * cg;
  const unsigned stride = gridDim.x * TPB;
  for (unsigned i0 = blockIdx.x * TPB + threadIdx.x; i0 < total; i0 += U * stride) {
    h16x8 v[U][4];
    unsigned pixs[U], gs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned idx = i0 + u * stride;
      if (idx >= total) continue;
      const unsigned pix = fd_div(idx, p.fd_cg), g = idx - pix * cg;
      const unsigned t = fd_div(pix, p.fd_w), ow = pix - t * p.y.W;
      const unsigned n = fd_div(t, p.fd_h), oh = t - n * p.y.H;
      pixs[u] = pix; gs[u] = g;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < p.n_in)
          v[u][i] = *reinterpret_cast<const h16x8*>(
              static_cast<const h16*>(p.x[i].base) + p.x[i].c0 + g * 8 +
              ((static_cast<size_t>(n) * p.x[i].H + (oh >> p.sh[i])) * p.x[i].W + (ow >> p.sh[i])) * p.x[i].ctot);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * stride >= total) continue;
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < p.n_in) {
          float xv[8];
          unpack8(v[u][i], xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += xv[j];
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = actf(acc[j], p.act);
      st8(p.y, pixs[u], gs[u] * 8, acc);
    }
  }
}

// ---------------------------------------------------------------------------------- global average pool
// One kernel: block (split, n) sums its pixel range into partial[n][split][c]; the LAST block of image n to finish
// (ticket counter) adds the partials in split order and writes out[n][c] = sum / (H*W).  Deterministic: the
// summation order is fixed by the grid, not by arrival order.
struct GapParams {
  View x;
  int N, splits;
  float* partial;      // [N][splits][C]
  float* out;          // [N][C]
  unsigned* tickets;   // [N], zero between launches (the last block resets its counter)
};
__global__ void __launch_bounds__(TPB) gap_kernel(const __grid_constant__ GapParams p) {
  pdl_prologue();
  extern __shared__ float sh[];  // [TPB/cg rows][C]
  __shared__ unsigned s_last;
  const int C = p.x.c, cg = C >> 3;
  const int n = blockIdx.y, split = blockIdx.x;
  const int rows = TPB / cg;  // pixel lanes per block (host guarantees cg <= TPB)
  const int g = threadIdx.x % cg, row = threadIdx.x / cg;
  const int HW = p.x.H * p.x.W;
  const int per = (HW + p.splits - 1) / p.splits;
  const int beg = split * per, end = min(HW, beg + per);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (row < rows) {
    const h16* xb = static_cast<const h16*>(p.x.base) + p.x.c0 + g * 8 + static_cast<size_t>(n) * HW * p.x.ctot;
    int i = beg + row;
    for (; i + 3 * rows < end; i += 4 * rows) {  // four independent 16-byte loads in flight
      h16x8 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const h16x8*>(xb + static_cast<size_t>(i + u * rows) * p.x.ctot);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float xv[8];
        unpack8(v[u], xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += xv[j];
      }
    }
    for (; i < end; i += rows) {
      float xv[8];
      unpack8(*reinterpret_cast<const h16x8*>(xb + static_cast<size_t>(i) * p.x.ctot), xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += xv[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[row * C + g * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += TPB) {
    float s = 0.0f;
    for (int r2 = 0; r2 < rows; ++r2) s += sh[r2 * C + c];
    p.partial[(static_cast<size_t>(n) * p.splits + split) * C + c] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(p.tickets + n, 1u) == static_cast<unsigned>(p.splits - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += TPB) {
    float s = 0.0f;
    for (int k2 = 0; k2 < p.splits; ++k2) s += __ldcg(p.partial + (static_cast<size_t>(n) * p.splits + k2) * C + c);
    p.out[static_cast<size_t>(n) * C + c] = s / static_cast<float>(HW);
  }
  if (threadIdx.x == 0) p.tickets[n] = 0u;
}

// ---------------------------------------------------------------------------------- dense on pooled vectors
struct DenseParams {
  const float* x[5];
  float* y;
  const float* w;  // [Cout][Cin]
  const float* b;
  int n_in, N, cin, cout, act;
};
// one warp per (n, output channel)
__global__ void __launch_bounds__(TPB) dense_kernel(const __grid_constant__ DenseParams p) {
  pdl_prologue();
  const int wid = (blockIdx.x * TPB + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= p.N * p.cout) return;
  const int n = wid / p.cout, co = wid % p.cout;
  float s = 0.0f;
  for (int ci = lane; ci < p.cin; ci += 32) {
    float xv = 0.0f;
    for (int i = 0; i < p.n_in; ++i) xv += p.x[i][static_cast<size_t>(n) * p.cin + ci];
    s = fmaf(xv, p.w[static_cast<size_t>(co) * p.cin + ci], s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) p.y[wid] = actf(s + p.b[co], p.act);
}

// ---------------------------------------------------------------------------------- attention gates
struct GateParams {
  View x, y, s;    // s: 1-channel spatial logits (BAM)
  const float* v;  // [N][C]
  const float* w;  // [C] spatial-squeeze weights (scSE)
  float b;
  int mode, N;
  FastDiv fd_cg, fd_hw;
};
// SE: y = x*v ; BAM: y = x*(1+sigmoid(v+s))
__global__ void __launch_bounds__(TPB) gate_kernel(const __grid_constant__ GateParams p) {
  pdl_prologue();
  constexpr int U = 4;
  const unsigned C = p.x.c, cg = C >> 3;
  const unsigned HW = static_cast<unsigned>(p.x.H) * p.x.W;
  const unsigned total = static_cast<unsigned>(p.N) * HW * cg;
  const unsigned stride = gridDim.x * TPB;
  const h16* xb = static_cast<const h16*>(p.x.base) + p.x.c0;
  for (unsigned i0 = blockIdx.x * TPB + threadIdx.x; i0 < total; i0 += U * stride) {
    h16x8 xr[U];
    float sg[U];
    unsigned pixs[U], gs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned idx = i0 + u * stride;
      if (idx >= total) continue;
      const unsigned pix = fd_div(idx, p.fd_cg), g = idx - pix * cg;
      pixs[u] = pix; gs[u] = g;
      xr[u] = *reinterpret_cast<const h16x8*>(xb + static_cast<size_t>(pix) * p.x.ctot + g * 8);
      if (p.mode != 0) sg[u] = ld1(p.s, pix, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * stride >= total) continue;
      const unsigned n = fd_div(pixs[u], p.fd_hw);
      const float4* vv = reinterpret_cast<const float4*>(p.v + static_cast<size_t>(n) * C + gs[u] * 8);
      const float4 va = __ldg(vv), vb = __ldg(vv + 1);
      const float vf[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
      float xv[8];
      unpack8(xr[u], xv);
      if (p.mode == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] *= vf[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] *= 1.0f + sigmoidf_(vf[j] + sg[u]);
      }
      st8(p.y, pixs[u], gs[u] * 8, xv);
    }
  }
}
// scSE: y = x*(sigmoid(w.x + b) + v): a group of LPP = min(32, C/8) lanes owns one pixel and VPL = C/8/LPP vectors per
// lane; the per-pixel dot product over channels is a shuffle reduction inside the group.  U pixels per group and
// iteration, all U x VPL loads issued before the first reduction; x stays in registers between the dot product and
// the gated store (read once, written once).
template <int VPL, int U>
__global__ void __launch_bounds__(TPB) gate_scse_kernel(const __grid_constant__ GateParams p, int lanes_per_pix) {
  pdl_prologue();
  const unsigned C = p.x.c;
  const unsigned HW = static_cast<unsigned>(p.x.H) * p.x.W;
  const unsigned npix = static_cast<unsigned>(p.N) * HW;
  const unsigned lane = threadIdx.x & 31;
  const unsigned sub = lane % lanes_per_pix;
  const unsigned ppw = 32 / lanes_per_pix;  // pixels per warp and step
  const unsigned warp_global = (blockIdx.x * TPB + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * TPB) >> 5;
  const h16* xb = static_cast<const h16*>(p.x.base) + p.x.c0;
  float wv[VPL][8];  // the lane's spatial-squeeze weights
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[k][j] = p.w[(sub + k * lanes_per_pix) * 8 + j];
  for (unsigned base = warp_global * ppw; base < npix; base += nwarps * ppw * U) {  // warp-uniform trip count
    h16x8 xr[U][VPL];
    unsigned pixs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      pixs[u] = base + u * nwarps * ppw + lane / lanes_per_pix;
      if (pixs[u] < npix) {
        const h16* xp = xb + static_cast<size_t>(pixs[u]) * p.x.ctot + sub * 8;
#pragma unroll
        for (int k = 0; k < VPL; ++k) xr[u][k] = *reinterpret_cast<const h16x8*>(xp + k * lanes_per_pix * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = pixs[u] < npix;
      float xv[VPL][8];
      float dot = 0.0f;
      if (ok) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          unpack8(xr[u][k], xv[k]);
#pragma unroll
          for (int j = 0; j < 8; ++j) dot = fmaf(xv[k][j], wv[k][j], dot);
        }
      }
      for (int o = lanes_per_pix >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (!ok) continue;
      const float sp = sigmoidf_(dot + p.b);
      const unsigned n = fd_div(pixs[u], p.fd_hw);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const unsigned g = sub + k * lanes_per_pix;
        const float4* vv = reinterpret_cast<const float4*>(p.v + static_cast<size_t>(n) * C + g * 8);
        const float4 va = __ldg(vv), vb = __ldg(vv + 1);
        const float vf[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[k][j] = xv[k][j] * sp + xv[k][j] * vf[j];
        st8(p.y, pixs[u], g * 8, xv[k]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------- selective-kernel fusion
struct SkParams {
  View x[4], y;
  const float* g;       // [N][C]
  const float* lg[5];   // [N][C] each
  const float* scale;   // [C]
  const float* shift;
  int N;
};
__global__ void __launch_bounds__(TPB) skfuse_kernel(const __grid_constant__ SkParams p) {
  pdl_prologue();
  const int C = p.y.c, cg = C >> 3;
  const size_t HW = static_cast<size_t>(p.y.H) * p.y.W;
  const size_t total = static_cast<size_t>(p.N) * HW * cg;
  for (size_t idx = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * TPB) {
    const int g = static_cast<int>(idx % cg);
    const size_t pix = idx / cg;
    const int n = static_cast<int>(pix / HW);
    float xs[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld8(p.x[i], pix, g * 8, xs[i]);
    float out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t vi = static_cast<size_t>(n) * C + g * 8 + j;
      float l[5], m = -INFINITY;
#pragma unroll
      for (int i = 0; i < 5; ++i) { l[i] = p.lg[i][vi]; m = fmaxf(m, l[i]); }
      float den = 0.0f;
#pragma unroll
      for (int i = 0; i < 5; ++i) { l[i] = __expf(l[i] - m); den += l[i]; }
      const float inv = 1.0f / den;
      float acc = p.g[vi] * l[4] * inv;
#pragma unroll
      for (int i = 0; i < 4; ++i) acc = fmaf(xs[i][j], l[i] * inv, acc);
      out[j] = fmaxf(fmaf(acc, p.scale[g * 8 + j], p.shift[g * 8 + j]), 0.0f);
    }
    st8(p.y, pix, g * 8, out);
  }
}

// ---------------------------------------------------------------------------------- broadcast vector over a map slice
struct BcastParams {
  View y;
  const float* v;
  int N;
};
__global__ void __launch_bounds__(TPB) bcast_kernel(const __grid_constant__ BcastParams p) {
  pdl_prologue();
  const int C = p.y.c, cg = C >> 3;
  const size_t HW = static_cast<size_t>(p.y.H) * p.y.W;
  const size_t total = static_cast<size_t>(p.N) * HW * cg;
  for (size_t idx = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * TPB) {
    const int g = static_cast<int>(idx % cg);
    const size_t pix = idx / cg;
    const int n = static_cast<int>(pix / HW);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = p.v[static_cast<size_t>(n) * C + g * 8 + j];
    st8(p.y, pix, g * 8, f);
  }
}

// ---------------------------------------------------------------------------------- softmax head
// logits fp32 (N, 512/up, 512/up, 2) -> probs fp32 (N,512,512,2) and/or mask u8 (argmax of the probabilities, ties -> class 0)
__global__ void __launch_bounds__(TPB) softmax2_kernel(const float* __restrict__ logits, int N, int H, int W, int up,
                                                       float* __restrict__ probs, uint8_t* __restrict__ mask) {
  pdl_prologue();
  // thread = 4 consecutive output pixels of a row (W % 4 == 0: W is 512): one 4-byte mask store, two float4 prob stores
  const unsigned w4 = static_cast<unsigned>(W) >> 2;
  const unsigned total = static_cast<unsigned>(N) * H * w4;
  const int h2 = H / up, w2 = W / up;
  const int sh = up == 1 ? 0 : up == 2 ? 1 : up == 4 ? 2 : -1;
  for (unsigned idx = blockIdx.x * TPB + threadIdx.x; idx < total; idx += gridDim.x * TPB) {
    const unsigned row = idx / w4, q = idx - row * w4;     // row = n * H + oh
    const unsigned n = row / H, oh = row - n * H;
    const unsigned ow = q * 4;
    const unsigned sy = sh >= 0 ? oh >> sh : oh / up;
    const float* lrow = logits + (static_cast<size_t>(n) * h2 + sy) * w2 * 2;
    float2 l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) l[j] = *reinterpret_cast<const float2*>(lrow + 2 * (sh >= 0 ? (ow + j) >> sh : (ow + j) / up));
    // The mask is the argmax of the float32 PROBABILITIES (predict.py:109-110: tf.argmax of the softmax output, first
    // maximum on ties), not of the logits: logits closer than one float32 ulp of 0.5 tie after the softmax.
    float pr[8];
    uint32_t mk = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float m = fmaxf(l[j].x, l[j].y);
      const float e0 = expf(l[j].x - m), e1 = expf(l[j].y - m);
      const float inv = 1.0f / (e0 + e1);
      pr[2 * j] = e0 * inv; pr[2 * j + 1] = e1 * inv;
      mk |= (pr[2 * j + 1] > pr[2 * j] ? 1u : 0u) << (8 * j);
    }
    const size_t o = static_cast<size_t>(row) * W + ow;
    if (probs) {
      *reinterpret_cast<float4*>(probs + o * 2) = make_float4(pr[0], pr[1], pr[2], pr[3]);
      *reinterpret_cast<float4*>(probs + o * 2 + 4) = make_float4(pr[4], pr[5], pr[6], pr[7]);
    }
    if (mask) *reinterpret_cast<uint32_t*>(mask + o) = mk;
  }
}

// ---------------------------------------------------------------------------------- probability-averaging fusion (opt-in)
// acc[i] (+)= P(class 1) of one model; after the last model mask[i] = acc[i] > thr (thr = n_models / 2: mean > 0.5)
__global__ void __launch_bounds__(TPB) prob_accum_kernel(const float* __restrict__ probs, float* __restrict__ acc, size_t n,
                                                         int first, float thr, uint8_t* __restrict__ mask) {
  pdl_prologue();
  for (size_t i = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * TPB) {
    const float v = (first ? 0.0f : acc[i]) + probs[2 * i + 1];
    acc[i] = v;
    if (mask) mask[i] = v > thr ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------- tiler / stitcher
// Network input layout (graph.Net.input): fp16, 32 channels per pixel of the STEM OUTPUT grid, channel
// (kh*3+kw)*3+c = 255 * x[oy*S+kh-PAD, ox*S+kw-PAD, c] for the 3x3 stem conv of stride S with TF 'same' padding
// (PAD = 1 for S = 1, 0 for S = 2), zero outside the 512x512 tile.  predict.py:91-108 computes x = pixel/127.5 - 1
// (float64, cast to float32 by Keras) and zero-pads in normalised space; 255*x = 2*pixel - 255 is an integer in
// [-255, 255], exact in fp16; the stem weights carry the 1/255.  thread = (output pixel, 8-channel quarter).
template <int S>
__global__ void __launch_bounds__(TPB) tiles_gather_kernel(const uint8_t* __restrict__ scene, int H, int W,
                                                           const int* __restrict__ ys, const int* __restrict__ xs,
                                                           int n, h16* __restrict__ out) {
  pdl_prologue();
  constexpr int O = 512 / S, PAD = S == 1 ? 1 : 0;
  const size_t total = static_cast<size_t>(n) * O * O * 4;
  for (size_t idx = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * TPB) {
    const int q = static_cast<int>(idx & 3);
    const size_t pix = idx >> 2;
    const int ox = static_cast<int>(pix % O), oy = static_cast<int>((pix / O) % O);
    const int t = static_cast<int>(pix / (static_cast<size_t>(O) * O));
    const int y0 = ys[t], x0 = xs[t];
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = q * 8 + j, tap = k / 3, c = k - tap * 3;
      float v = 0.0f;
      if (k < 27) {
        const int ty = oy * S + tap / 3 - PAD, tx = ox * S + tap % 3 - PAD;  // position inside the tile
        const int y = y0 + ty, x = x0 + tx;
        if (ty >= 0 && ty < 512 && tx >= 0 && tx < 512 && y < H && x < W)
          v = static_cast<float>(2 * static_cast<int>(scene[(static_cast<size_t>(y) * W + x) * 3 + (2 - c)]) - 255);  // BGR -> RGB
      }
      f[j] = v;
    }
    *reinterpret_cast<h16x8*>(out + idx * 8) = pack8(f);
  }
}
// model.predict(x) path: fp32 (N,512,512,3) in [-1,1] -> the same layout (255*x rounded to fp16)
template <int S>
__global__ void __launch_bounds__(TPB) input_convert_kernel(const float* __restrict__ x, int n, h16* __restrict__ out) {
  pdl_prologue();
  constexpr int O = 512 / S, PAD = S == 1 ? 1 : 0;
  const size_t total = static_cast<size_t>(n) * O * O * 4;
  for (size_t idx = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * TPB) {
    const int q = static_cast<int>(idx & 3);
    const size_t pix = idx >> 2;
    const int ox = static_cast<int>(pix % O), oy = static_cast<int>((pix / O) % O);
    const size_t t = pix / (static_cast<size_t>(O) * O);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = q * 8 + j, tap = k / 3, c = k - tap * 3;
      float v = 0.0f;
      if (k < 27) {
        const int ty = oy * S + tap / 3 - PAD, tx = ox * S + tap % 3 - PAD;
        if (ty >= 0 && ty < 512 && tx >= 0 && tx < 512) v = x[((t * 512 + ty) * 512 + tx) * 3 + c] * 255.0f;
      }
      f[j] = v;
    }
    *reinterpret_cast<h16x8*>(out + idx * 8) = pack8(f);
  }
}
// predict.py:113-114: scene[y,x] = 255 where any covering tile predicts class 1 (all writers store 255).
__global__ void __launch_bounds__(TPB) stitch_or_kernel(const uint8_t* __restrict__ tiles, const int* __restrict__ ys,
                                                        const int* __restrict__ xs, int n, uint8_t* __restrict__ scene,
                                                        int H, int W) {
  pdl_prologue();
  const size_t total = static_cast<size_t>(n) * 512 * 512;
  for (size_t idx = blockIdx.x * static_cast<size_t>(TPB) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * TPB) {
    if (!tiles[idx]) continue;
    const int tx = static_cast<int>(idx % 512), ty = static_cast<int>((idx / 512) % 512);
    const int t = static_cast<int>(idx / (512 * 512));
    const int y = ys[t] + ty, x = xs[t] + tx;
    if (y < H && x < W) scene[static_cast<size_t>(y) * W + x] = 255;
  }
}

}  // namespace k
}  // namespace bd
