// dw_tma.cuh -- depthwise 3x3 convolution ('same', stride 1) through shared-memory tiles moved by TMA.
//
// The register-window kernel (k::dwconv3x3_kernel) is latency bound on the small 728-channel maps of the Xception
// middle flow (each thread chases ~40 dependent L2 round trips).  Here the memory side is asynchronous:
//   warp 0 (one lane)   producer: per work item (image, 8 x 16 pixel tile, 64-channel chunk) one TMA box load of the
//                       (8+2) x (16+2) halo (zero fill outside the map = the 'same' padding), ring of stages
//   warps 1-4           compute: thread = (4-channel group, 4x4 output patch); a 6x6 window of 8-byte vectors from
//                       the swizzled halo tile, nine taps on packed half2 FMAs in the order and rounding of
//                       k::dwconv3x3_kernel (bit-identical results), 16 outputs into a swizzled 16 KB output tile
//   one thread          TMA store of the output tile (clipped by the tensor map at ragged edges / channel tails)
// Two CTAs per SM (8 compute warps) whenever the weights of all chunks fit next to two halo stages.
#pragma once
#include "conv_umma.cuh"

namespace bd {
namespace dwt {

constexpr int THREADS = 160;
constexpr int OUT_TILE_BYTES = 128 * 128;  // 128 pixels x 64 channels fp16

struct Params {
  int N, H, W, C;          // map geometry, channels of the slice
  int kchunks, tiles_w, tiles_h, total_items, stages, relu_in;
  int out_tiles;           // output staging tiles (2, or 1 when that buys a third halo stage)
  int w_bytes;             // staged weights [9][kchunks*64] fp16
  umma::FastDiv fd_kc, fd_tw, fd_th;
  const h16* w;            // [9][C] fp16
};
struct alignas(64) Maps {
  CUtensorMap x, y;
};

__device__ __forceinline__ void item_coords(const Params& p, int item, int& kc, int& tw, int& th, int& n) {
  const uint32_t t = static_cast<uint32_t>(item);
  const uint32_t m = umma::fd_div(t, p.fd_kc);
  kc = static_cast<int>(t - m * p.kchunks);
  const uint32_t t2 = umma::fd_div(m, p.fd_tw);
  tw = static_cast<int>(m - t2 * p.tiles_w);
  const uint32_t t3 = umma::fd_div(t2, p.fd_th);
  th = static_cast<int>(t2 - t3 * p.tiles_h);
  n = static_cast<int>(t3);
}

__global__ void __launch_bounds__(THREADS, 2) dwconv_tma_kernel(const __grid_constant__ Maps maps,
                                                                const __grid_constant__ Params p) {
  using namespace umma;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t out0 = smem_base + static_cast<uint32_t>(p.stages) * HALO_STAGE;  // two output tiles
  const uint32_t w0 = out0 + static_cast<uint32_t>(p.out_tiles) * OUT_TILE_BYTES;   // staged weights
  const uint32_t full0 = w0 + static_cast<uint32_t>(p.w_bytes), empty0 = full0 + 8u * p.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, 4);  // one arrival per compute warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
    prefetch_tmap(&maps.x);
    prefetch_tmap(&maps.y);
  }
  {  // weights [9][C] -> [9][kchunks*64] fp16, zero tail (constant data: before the dependency wait)
    const int v8 = p.kchunks * 8;  // 8-channel vectors per tap (C % 8 == 0)
    uint4* sw = reinterpret_cast<uint4*>(gen_base + (w0 - smem_base));
    for (int i = threadIdx.x; i < 9 * v8; i += THREADS) {
      const int k = i / v8, c = (i - k * v8) * 8;
      sw[i] = c < p.C ? __ldg(reinterpret_cast<const uint4*>(p.w + k * p.C + c)) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();
  pdl_wait();
  if (warp == 0) {
    // ---------------- producer
    uint32_t s = 0, ph = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int kc, tw, th, n;
      item_coords(p, item, kc, tw, th, n);
      mbar_wait(empty0 + 8u * s, ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(full0 + 8u * s, HALO_BYTES);
        tma_load_4d(smem_base + s * HALO_STAGE, &maps.x, full0 + 8u * s, kc * 64, tw * 8 - 1, th * 16 - 1, n);
      }
      __syncwarp();
      if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; }
    }
  } else {
    // ---------------- compute warps 1..4
    const int dt = threadIdx.x - 32;  // 0..127
    const int g4 = dt & 15, patch = dt >> 4, px0 = (patch & 1) * 4, py0 = (patch >> 1) * 4;
    const uint32_t sub8 = (g4 & 1) * 8;  // byte offset of my 4 channels inside their 16-byte chunk
    const int gc = g4 >> 1;              // 16-byte chunk of my channels
    const __half2 zero2 = __float2half2_rn(0.0f);
    union U { uint2 v; __half2 h[2]; };
    const int cpad = p.kchunks * 64;
    uint32_t s = 0, ph = 0, ob = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ob ^= 1u) {
      int kc, tw, th, n;
      item_coords(p, item, kc, tw, th, n);
      U wp[9];
#pragma unroll
      for (int k = 0; k < 9; ++k)
        wp[k].v = *reinterpret_cast<const uint2*>(gen_base + (w0 - smem_base) + (static_cast<size_t>(k) * cpad + kc * 64 + g4 * 4) * 2);
      mbar_wait(full0 + 8u * s, ph);
      const uint8_t* halo = gen_base + s * HALO_STAGE;
      U win[6][6];  // halo rows py0 .. py0+5, columns px0 .. px0+5
#pragma unroll
      for (int hy = 0; hy < 6; ++hy)
#pragma unroll
        for (int hx = 0; hx < 6; ++hx) {
          const int P = (py0 + hy) * HALO_W + px0 + hx;  // halo pixel; its 16-byte chunks are XOR-swizzled by P % 8
          win[hy][hx].v = *reinterpret_cast<const uint2*>(halo + P * 128 + ((gc ^ (P & 7)) << 4) + sub8);
        }
      if (p.relu_in) {
#pragma unroll
        for (int hy = 0; hy < 6; ++hy)
#pragma unroll
          for (int hx = 0; hx < 6; ++hx) {
            win[hy][hx].h[0] = __hmax2(win[hy][hx].h[0], zero2);
            win[hy][hx].h[1] = __hmax2(win[hy][hx].h[1], zero2);
          }
      }
      // the window lives in registers: hand the halo stage back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8u * s);
      if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; }
      // the TMA store that last read this output tile (two items ago; the previous item with a single tile) must have
      // finished reading it
      if (dt == 0) {
        if (p.out_tiles == 2) tma_store_wait_read<1>();
        else tma_store_wait_read<0>();
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (p.out_tiles == 1) ob = 0u;
      uint8_t* otile = gen_base + (out0 - smem_base) + ob * OUT_TILE_BYTES;
#pragma unroll
      for (int oy = 0; oy < 4; ++oy)
#pragma unroll
        for (int ox = 0; ox < 4; ++ox) {
          U acc;
          acc.h[0] = zero2; acc.h[1] = zero2;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              acc.h[0] = __hfma2(win[oy + kh][ox + kw].h[0], wp[kh * 3 + kw].h[0], acc.h[0]);
              acc.h[1] = __hfma2(win[oy + kh][ox + kw].h[1], wp[kh * 3 + kw].h[1], acc.h[1]);
            }
          const int m = (py0 + oy) * 8 + px0 + ox;  // tile row (w fastest)
          *reinterpret_cast<uint2*>(otile + m * 128 + ((gc ^ (m & 7)) << 4) + sub8) = acc.v;
        }
      fence_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (dt == 0) {
        tma_store_4d(&maps.y, out0 + ob * OUT_TILE_BYTES, kc * 64, tw * 8, th * 16, n);
        tma_store_commit();
      }
    }
    if (dt == 0) tma_store_wait_all();
  }
}

struct Launch {
  Maps maps;
  Params p;
  dim3 grid;
  int smem_bytes;
};

// eligibility: stride 1, 'same' padding (pad 1), fp16 maps, 16-byte aligned slices, map at least 8 x 16
inline bool eligible(const TView& x, const TView& y, int stride, int pad_t, int pad_l) {
  return stride == 1 && pad_t == 1 && pad_l == 1 && !x.f32 && !y.f32 && x.c == y.c && x.c % 8 == 0 && x.c0 % 8 == 0 &&
         x.ctot % 8 == 0 && y.c0 % 8 == 0 && y.ctot % 8 == 0 && x.W >= 8 && x.H >= 16 && x.H == y.H && x.W == y.W;
}

inline int prepare(Launch* L, const TView& x, const TView& y, const h16* w_dev, int relu_in, int num_sms) {
  Params& p = L->p;
  memset(&p, 0, sizeof(p));
  p.N = x.N; p.H = x.H; p.W = x.W; p.C = x.c;
  p.kchunks = cdiv(x.c, 64);
  p.tiles_w = cdiv(x.W, 8); p.tiles_h = cdiv(x.H, 16);
  p.total_items = p.N * p.tiles_w * p.tiles_h * p.kchunks;
  p.relu_in = relu_in;
  p.w = w_dev;
  p.w_bytes = (9 * p.kchunks * 64 * 2 + 127) / 128 * 128;
  p.fd_kc = umma::make_fastdiv(p.kchunks); p.fd_tw = umma::make_fastdiv(p.tiles_w); p.fd_th = umma::make_fastdiv(p.tiles_h);
  // two CTAs per SM when two halo stages + everything else fit into half the shared memory.  The kernel is bound by the
  // latency of the halo loads (ncu: long-scoreboard stalls, 20 % of the DRAM rate), i.e. by the bytes in flight: when two
  // output tiles leave room for only two halo stages (728 channels: 13.8 KB of weights) a single output tile buys a third
  p.out_tiles = 2;
  int fixed = 2 * OUT_TILE_BYTES + p.w_bytes + 1024 + 256;
  const int half_sm = 113 * 1024;
  int ctas_per_sm = 2;
  p.stages = (half_sm - fixed) / umma::HALO_STAGE;
  if (p.stages < 3 && (half_sm - (fixed - OUT_TILE_BYTES)) / umma::HALO_STAGE >= 3) {
    p.out_tiles = 1;
    fixed -= OUT_TILE_BYTES;
    p.stages = (half_sm - fixed) / umma::HALO_STAGE;
  }
  if (p.stages < 2) { ctas_per_sm = 1; p.stages = std::min(4, (226 * 1024 - fixed) / umma::HALO_STAGE); }
  p.stages = std::min(p.stages, 4);
  BD_CHECK(p.stages >= 2, "dwconv_tma: shared memory budget too small");
  L->smem_bytes = p.stages * umma::HALO_STAGE + fixed;
  L->grid = dim3(static_cast<unsigned>(std::min(p.total_items, ctas_per_sm * num_sms)));
  {
    const uint64_t pitch = static_cast<uint64_t>(x.ctot) * 2;
    uint64_t dims[4] = {static_cast<uint64_t>(x.c), static_cast<uint64_t>(x.W), static_cast<uint64_t>(x.H), static_cast<uint64_t>(x.N)};
    uint64_t strides[3] = {pitch, pitch * x.W, pitch * x.W * x.H};
    uint32_t box[4] = {64, umma::HALO_W, umma::HALO_H, 1};
    if (umma::encode_h16(&L->maps.x, static_cast<char*>(x.base) + static_cast<size_t>(x.c0) * 2, 4, dims, strides, box)) return 1;
  }
  {
    const uint64_t pitch = static_cast<uint64_t>(y.ctot) * 2;
    uint64_t dims[4] = {static_cast<uint64_t>(y.c), static_cast<uint64_t>(y.W), static_cast<uint64_t>(y.H), static_cast<uint64_t>(y.N)};
    uint64_t strides[3] = {pitch, pitch * y.W, pitch * y.W * y.H};
    uint32_t box[4] = {64, 8, 16, 1};
    if (umma::encode_h16(&L->maps.y, static_cast<char*>(y.base) + static_cast<size_t>(y.c0) * 2, 4, dims, strides, box)) return 1;
  }
  return 0;
}

inline int launch(const Launch& L, cudaStream_t stream, bool pdl) {
  static bool attr_set_dev[64] = {};  // per-device attribute
  int dev_ = 0;
  BD_CUDA(cudaGetDevice(&dev_));
  bool& attr_set = attr_set_dev[dev_ & 63];
  if (!attr_set) {
    BD_CUDA(cudaFuncSetAttribute(dwconv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  BD_CUDA(launch_k(pdl, dwconv_tma_kernel, L.grid, dim3(THREADS), static_cast<size_t>(L.smem_bytes), stream, L.maps, L.p));
  return 0;
}

}  // namespace dwt
}  // namespace bd
