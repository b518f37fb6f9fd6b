// tma_rate.cu -- micro-benchmark: how fast can one elected thread stream TMA tensor loads into a ring of shared
// memory stages (no MMA)?  Varies box shape / rank / swizzle / descriptor location.  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_rate tools/micro/tma_rate.cu -lcuda && /tmp/tma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  uint32_t d = 0;
  while (!d) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(d) : "r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6}], [%2];"
               ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3}], [%4];"
               ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// mode 0: 4-D box (64ch, 16w, 8h, 1n) walking over an NHWC image; mode 1: 2-D box (64, rows)
template <int MODE>
__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap tm, const CUtensorMap* tm_global, int use_global,
                                           int iters, int stages, int box_bytes, int ops_per_stage, int W, int H, long long* out) {
  extern __shared__ uint8_t raw[];
  uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint32_t bars = base + stages * box_bytes * ops_per_stage;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (stages + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const CUtensorMap* t = use_global ? tm_global : &tm;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {  // producer
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      int s = it % stages; uint32_t ph = (it / stages) & 1;
      mbar_wait(bars + 8 * (stages + s), ph ^ 1);
      mbar_expect(bars + 8 * s, box_bytes * ops_per_stage);
      for (int o = 0; o < ops_per_stage; ++o) {
        int idx = (blockIdx.x * 7 + it * ops_per_stage + o);
        uint32_t dst = base + (s * ops_per_stage + o) * box_bytes;
        if (MODE == 0) tma4(dst, t, bars + 8 * s, 0, (idx * 16) % W, ((idx * 16) / W * 8) % H, (idx / 1024) % 16);
        else tma2(dst, t, bars + 8 * s, 0, (idx * (box_bytes / 128)) % (W * H));
      }
    }
  } else if (threadIdx.x == 32) {  // consumer: just frees the stage
    for (int it = 0; it < iters; ++it) {
      int s = it % stages; uint32_t ph = (it / stages) & 1;
      mbar_wait(bars + 8 * s, ph);
      mbar_arrive(bars + 8 * (stages + s));
    }
    t1 = clock64();
    out[blockIdx.x] = t1;
  }
  if (threadIdx.x == 0) out[gridDim.x + blockIdx.x] = t0;
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  Enc enc = (Enc)fp;
  const int N = 16, H = 512, W = 512, C = 64;
  size_t bytes = (size_t)N * H * W * C * 2;
  void* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 0, bytes));
  long long* out; CK(cudaMalloc(&out, sizeof(long long) * 2 * 148));
  CUtensorMap* tmg; CK(cudaMalloc(&tmg, sizeof(CUtensorMap)));
  std::vector<long long> h(2 * 148);
  struct Cfg { const char* name; int mode, swz, promo, use_global, box_rows, ops, stages; };
  Cfg cfgs[] = {
      {"4D 64x16x8  sw128 promo256 param  1op/stage 6st", 0, 1, 2, 0, 128, 1, 6},
      {"4D 64x16x8  sw128 promo256 param  2op/stage 6st", 0, 1, 2, 0, 128, 2, 6},
      {"4D 64x16x8  sw128 promo256 GLOBAL 1op/stage 6st", 0, 1, 2, 1, 128, 1, 6},
      {"4D 64x16x8  sw128 promo128 param  1op/stage 6st", 0, 1, 1, 0, 128, 1, 6},
      {"4D 64x16x8  sw128 promoNONE param 1op/stage 6st", 0, 1, 0, 0, 128, 1, 6},
      {"4D 64x16x8  swNONE promo256 param 1op/stage 6st", 0, 0, 2, 0, 128, 1, 6},
      {"2D 64x128   sw128 promo256 param  1op/stage 6st", 1, 1, 2, 0, 128, 1, 6},
      {"2D 64x256   sw128 promo256 param  1op/stage 4st", 1, 1, 2, 0, 256, 1, 4},
      {"2D 64x64    sw128 promo256 param  1op/stage 8st", 1, 1, 2, 0, 64, 1, 8},
      {"2D 64x32    sw128 promo256 param  1op/stage 8st", 1, 1, 2, 0, 32, 1, 8},
      {"4D 64x16x8  sw128 promo256 param  1op/stage 12st", 0, 1, 2, 0, 128, 1, 12},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap tm;
    CUtensorMapSwizzle sw = c.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUtensorMapL2promotion pr = c.promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : c.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    CUresult r;
    if (c.mode == 0) {
      cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      cuuint64_t gs[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
      cuuint32_t bx[4] = {64, 16, 8, 1}, es[4] = {1, 1, 1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t gd[2] = {(cuuint64_t)C, (cuuint64_t)N * H * W};
      cuuint64_t gs[1] = {(cuuint64_t)C * 2};
      cuuint32_t bx[2] = {64, (cuuint32_t)c.box_rows}, es[2] = {1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    CK(cudaMemcpy(tmg, &tm, sizeof(tm), cudaMemcpyHostToDevice));
    const int box_bytes = c.box_rows * 128, iters = 2000;
    const int smem = c.stages * box_bytes * c.ops + 1024 + 256;
    for (int grid : {1, 148}) {
      float ms = 0;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (c.mode == 0) {
          CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          k<0><<<grid, 64, smem>>>(tm, tmg, c.use_global, iters, c.stages, box_bytes, c.ops, W, H, out);
        } else {
          CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          k<1><<<grid, 64, smem>>>(tm, tmg, c.use_global, iters, c.stages, box_bytes, c.ops, W, H, out);
        }
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
      }
      CK(cudaMemcpy(h.data(), out, sizeof(long long) * 2 * 148, cudaMemcpyDeviceToHost));
      double clk = (double)(h[0] - h[grid]) / iters;  // block 0: end - start
      printf("%-52s grid %3d: %7.1f clk/stage  %6.1f clk/op  %6.1f B/clk/SM   (%.3f ms)\n", c.name, grid, clk, clk / c.ops,
             box_bytes * c.ops / clk, ms);
    }
  }
  return 0;
}
