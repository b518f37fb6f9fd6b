// conv_umma.cuh -- implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
// One CTA computes a 128-pixel x BLOCK_N-channel output tile.  GEMM view of the convolution:
//   D[m, n] = sum_{tap, c} A_tap[m, c] * W[tap][n][c]
// where m runs over a (bn x bh x bw) box of output pixels (bn*bh*bw = 128) and A_tap is the NHWC
// input box shifted by the tap offset (dy, dx).  The im2col matrix is never materialised: for every
// (tap, 64-channel chunk) the TMA engine loads the shifted 4-D box straight from the activation
// tensor into shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes, and its
// out-of-bounds zero fill IS the 'same' padding.  Stride-2 convolutions use up to four parity views
// of the input (one tensor map per (row parity, column parity)), each again a plain tiled map.
//
// Roles (128 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (+ TMEM alloc by
// warp 1), then all four warps run the epilogue: tcgen05.ld the fp32 accumulators (warp w owns TMEM
// lanes 32w..32w+31 = tile rows), add bias (BatchNorm folded), optional ReLU / residual / ReLU, pack
// to h16 and store 16-byte vectors into the destination channel slice (concat elision; pixel
// scatter for the sub-pixel transposed convolutions).
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace bd {
namespace umma {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // h16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int MAX_TAPS = 9;

struct Params {
  int N, Ho, Wo, Cout;
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n, n_tiles;
  int block_n, kchunks, ntaps, stages, tmem_cols;
  int tap_map[MAX_TAPS], tap_dy[MAX_TAPS], tap_dx[MAX_TAPS];
  h16* y;
  int y_ctot, y_c0, y_H, y_W, out_scale, out_oy, out_ox;
  const h16* res;
  int res_ctot, res_c0;
  const float* bias;
  int act_pre, act_post;
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a tile finishes in microseconds, so ~2^24 polls mean a protocol bug -> trap instead of
// hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  printf("bd conv_umma: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar,
         parity);
  __trap();
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (8 rows * 128 B) |
// version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A and B formats at [7,10) / [10,13) = 0 (fp16; 1 would be
// bf16), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ float apply_act(float v, int act) { return act == 1 ? fmaxf(v, 0.0f) : v; }

// ------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(128) conv_umma_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                        const __grid_constant__ CUtensorMap tmA1,
                                                        const __grid_constant__ CUtensorMap tmA2,
                                                        const __grid_constant__ CUtensorMap tmA3,
                                                        const __grid_constant__ CUtensorMap tmB,
                                                        const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_bytes = static_cast<uint32_t>(p.block_n) * 128u;
  const uint32_t stage_bytes = A_STAGE_BYTES + b_bytes;
  const uint32_t bar_base = smem_base + static_cast<uint32_t>(p.stages) * stage_bytes;  // 8-byte slots
  const uint32_t full0 = bar_base, empty0 = bar_base + 8u * p.stages, accum_bar = bar_base + 16u * p.stages;
  const uint32_t holder = accum_bar + 8u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (holder - smem_base));

  const int nt = blockIdx.x % p.n_tiles;
  const int mt = blockIdx.x / p.n_tiles;
  const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
  const int n_base = nt * p.block_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;

  const int num_kb = p.ntaps * p.kchunks;
  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmB);
    const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % p.stages;
      const uint32_t ph = (kb / p.stages) & 1;
      mbar_wait(empty0 + 8u * s, ph ^ 1u);
      const uint32_t fb = full0 + 8u * s;
      mbar_expect_tx(fb, stage_bytes);
      const int tap = kb / p.kchunks, kc = kb - tap * p.kchunks;
      const int m = p.tap_map[tap];
      const CUtensorMap* tm = (m == 0) ? &tmA0 : (m == 1) ? &tmA1 : (m == 2) ? &tmA2 : &tmA3;
      const uint32_t a_s = smem_base + s * stage_bytes;
      tma_load_4d(a_s, tm, fb, kc * BLOCK_K, w0 + p.tap_dx[tap], h0 + p.tap_dy[tap], n0);
      tma_load_3d(a_s + A_STAGE_BYTES, &tmB, fb, kc * BLOCK_K, n_base, tap);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer
    const uint32_t idesc = make_idesc(p.block_n);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % p.stages;
      const uint32_t ph = (kb / p.stages) & 1;
      mbar_wait(full0 + 8u * s, ph);
      tc_fence_after();
      const uint32_t a_s = smem_base + s * stage_bytes;
      const uint64_t adesc = make_sdesc(a_s), bdesc = make_sdesc(a_s + A_STAGE_BYTES);
#pragma unroll
      for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
        // advance 16 h16 = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
        tc_mma_f16(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
      tc_commit(empty0 + 8u * s);  // frees the smem stage once these MMAs have read it
    }
    tc_commit(accum_bar);  // accumulators complete
  }
  __syncwarp();

  // ---------------- epilogue (all 4 warps)
  mbar_wait(accum_bar, 0);
  tc_fence_after();
  const int r = warp * 32 + lane;  // tile row = TMEM lane
  const int wl = r % p.bw, hl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);
  const int ow = tw * p.bw + wl, oh = th * p.bh + hl, on = tn * p.bn + nl;
  const bool pix_ok = (ow < p.Wo) && (oh < p.Ho) && (on < p.N);
  const size_t ypix = (static_cast<size_t>(on) * p.y_H + (oh * p.out_scale + p.out_oy)) * p.y_W +
                      (ow * p.out_scale + p.out_ox);
  h16* yrow = p.y + ypix * p.y_ctot + p.y_c0;
  const h16* rrow = p.res ? p.res + ypix * p.res_ctot + p.res_c0 : nullptr;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c0 = 0; c0 < p.block_n; c0 += 16) {
    uint32_t acc[16];
    tmem_ld16(trow + c0, acc);
    tmem_ld_wait();
    const int ch0 = n_base + c0;
    if (pix_ok && ch0 < p.Cout) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int ch = ch0 + 8 * g;
        if (ch < p.Cout) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = apply_act(__uint_as_float(acc[8 * g + j]) + __ldg(p.bias + ch + j), p.act_pre);
          if (rrow) {
            float rf[8];
            unpack8(*reinterpret_cast<const h16x8*>(rrow + ch), rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += rf[j];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], p.act_post);
          *reinterpret_cast<h16x8*>(yrow + ch) = pack8(v);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// h16 tensor map, 128-byte swizzle, zero OOB fill.  dims/strides innermost first; strides[i] is the
// byte stride of dim i+1.
inline int encode_h16(CUtensorMap* tm, void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                       const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
  return 0;
}

struct Launch {
  CUtensorMap tmA[4], tmB;
  Params p;
  dim3 grid;
  int smem_bytes;
};

inline int floor_pow2(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}

// x: input view, y: output view (h16 both).  w_dev: [ntaps][Cout][Cin] h16.
inline int prepare(Launch* L, const TView& x, const TView& y, const TView* res, int ntaps, const int* dy,
                   const int* dx, int stride, int Ho, int Wo, int act_pre, int act_post, int out_scale, int out_oy,
                   int out_ox, const h16* w_dev, const float* bias_dev, int smem_budget_kb, int max_block_n) {
  const int Cin = x.c, Cout = y.c;
  BD_CHECK(!x.f32 && !y.f32, "umma conv needs h16 maps");
  BD_CHECK(Cin % 8 == 0 && Cout % 8 == 0 && x.c0 % 8 == 0 && x.ctot % 8 == 0 && y.c0 % 8 == 0 && y.ctot % 8 == 0,
           "umma conv needs 16-byte aligned channel slices");
  BD_CHECK(stride == 1 || stride == 2, "umma conv stride must be 1 or 2");
  BD_CHECK(ntaps >= 1 && ntaps <= MAX_TAPS, "bad tap count");
  Params& p = L->p;
  memset(&p, 0, sizeof(p));
  p.N = x.N; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
  p.bw = std::min(16, floor_pow2(Wo));
  p.bh = std::min(BLOCK_M / p.bw, floor_pow2(Ho));
  p.bn = BLOCK_M / (p.bw * p.bh);
  p.tiles_w = cdiv(Wo, p.bw); p.tiles_h = cdiv(Ho, p.bh); p.tiles_n = cdiv(x.N, p.bn);
  const int cout16 = cdiv(Cout, 16) * 16;
  const int ntile = cdiv(cout16, max_block_n);
  p.block_n = cdiv(cdiv(cout16, ntile), 16) * 16;
  p.n_tiles = cdiv(Cout, p.block_n);
  p.kchunks = cdiv(Cin, BLOCK_K);
  p.ntaps = ntaps;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.block_n) p.tmem_cols *= 2;
  const int stage_bytes = A_STAGE_BYTES + p.block_n * 128;
  p.stages = std::max(2, std::min(8, (smem_budget_kb * 1024 - 2048) / stage_bytes));
  p.stages = std::min(p.stages, std::max(2, ntaps * p.kchunks));
  L->smem_bytes = p.stages * stage_bytes + 1024 + 256;
  BD_CHECK(L->smem_bytes <= 227 * 1024, "umma conv smem budget exceeded");

  // parity views of the input for stride 2 (a single plain view for stride 1)
  bool used[4] = {false, false, false, false};
  for (int t = 0; t < ntaps; ++t) {
    int py = 0, px = 0, oy = dy[t], ox = dx[t];
    if (stride == 2) {
      py = ((dy[t] % 2) + 2) % 2; px = ((dx[t] % 2) + 2) % 2;
      oy = (dy[t] - py) / 2; ox = (dx[t] - px) / 2;
    }
    p.tap_map[t] = py * 2 + px; p.tap_dy[t] = oy; p.tap_dx[t] = ox;
    used[py * 2 + px] = true;
  }
  const uint64_t pitch = static_cast<uint64_t>(x.ctot) * 2;
  int first = -1;
  for (int m = 0; m < 4; ++m) {
    if (!used[m]) continue;
    const int py = m / 2, px = m % 2;
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>((x.W - px + stride - 1) / stride),
                        static_cast<uint64_t>((x.H - py + stride - 1) / stride), static_cast<uint64_t>(x.N)};
    uint64_t strides[3] = {pitch * stride, pitch * x.W * stride, pitch * x.W * x.H};
    uint32_t box[4] = {BLOCK_K, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), static_cast<uint32_t>(p.bn)};
    char* base = static_cast<char*>(x.base) + (static_cast<size_t>(py) * x.W + px) * pitch + static_cast<size_t>(x.c0) * 2;
    if (encode_h16(&L->tmA[m], base, 4, dims, strides, box)) return 1;
    if (first < 0) first = m;
  }
  for (int m = 0; m < 4; ++m)
    if (!used[m]) L->tmA[m] = L->tmA[first];
  {
    uint64_t dims[3] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Cout), static_cast<uint64_t>(ntaps)};
    uint64_t strides[2] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Cin) * Cout * 2};
    uint32_t box[3] = {BLOCK_K, static_cast<uint32_t>(p.block_n), 1};
    if (encode_h16(&L->tmB, const_cast<h16*>(w_dev), 3, dims, strides, box)) return 1;
  }
  p.y = static_cast<h16*>(y.base);
  p.y_ctot = y.ctot; p.y_c0 = y.c0; p.y_H = y.H; p.y_W = y.W;
  p.out_scale = out_scale; p.out_oy = out_oy; p.out_ox = out_ox;
  if (res) {
    BD_CHECK(!res->f32 && res->c0 % 8 == 0 && res->ctot % 8 == 0 && out_scale == 1, "bad residual view");
    p.res = static_cast<const h16*>(res->base); p.res_ctot = res->ctot; p.res_c0 = res->c0;
  }
  p.bias = bias_dev; p.act_pre = act_pre; p.act_post = act_post;
  L->grid = dim3(static_cast<unsigned>(p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles));
  return 0;
}

inline int launch(const Launch& L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    BD_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  conv_umma_kernel<<<L.grid, 128, L.smem_bytes, stream>>>(L.tmA[0], L.tmA[1], L.tmA[2], L.tmA[3], L.tmB, L.p);
  BD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace umma
}  // namespace bd
