// common.cuh -- shared helpers for the sm_100a kernels of the building-detection hot path.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <nvtx3/nvToolsExt.h>

#include <string>
#include <utility>

namespace bd {

// Storage type of feature maps and conv weights: IEEE fp16 (11-bit significand).  bf16 runs at the same
// tcgen05 kind::f16 rate but its 8-bit significand accumulates ~8x more rounding error through the
// 60-200 layer chains of these networks than the 2e-2 probability tolerance allows (DESIGN.md
// "Numerics"); the fp32 -> fp16 conversion saturates instead of overflowing to inf.
typedef __half h16;
constexpr float H16_MAX = 65504.0f;

// thread-local last-error string behind bd_last_error()
std::string& last_error();
int fail(const std::string& msg);

#define BD_CUDA(call)                                                                             \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return ::bd::fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ +    \
                        ":" + std::to_string(__LINE__));                                          \
  } while (0)

#define BD_CHECK(cond, msg)                                                                       \
  do {                                                                                            \
    if (!(cond)) return ::bd::fail(std::string(msg) + " (" #cond ") @" + __FILE__ + ":" +         \
                                   std::to_string(__LINE__));                                     \
  } while (0)

// Programmatic dependent launch: every kernel of a plan starts with pdl_prologue() -- it lets the NEXT kernel of the
// stream begin launching (its CTAs run their own prologue as SM resources free up) and then waits until the
// PREVIOUS kernel has completed and its writes are visible.  Launched through launch_k() with pdl = true this hides
// the drain + launch latency between the ~700 dependent kernels of a batch; without the attribute both
// instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }

template <typename... Params, typename... Args>
inline cudaError_t launch_k(bool pdl, void (*kern)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// NVTX range per stage / model: shows up in nsys / ncu timelines, costs nothing without a tool attached
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// Every C-ABI entry point runs on its context's device and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev = -1, dev;
  explicit DeviceGuard(int d) : dev(d) {  // d < 0: no-op (null context; the entry point rejects it itself)
    if (dev < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (dev >= 0 && prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define BD_ON_CTX(ctx_) ::bd::DeviceGuard bd_dev_guard_((ctx_) ? (ctx_)->device : -1)
#define BD_ON_PLAN(p_) ::bd::DeviceGuard bd_dev_guard_(((p_) && (p_)->ctx) ? (p_)->ctx->device : -1)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// division by a launch constant: q = (umulhi(mul, n) + n) >> shr, exact for n < 2^31 (Granlund-Montgomery)
struct FastDiv { uint32_t mul, shr; };
inline FastDiv make_fastdiv(uint32_t d) {
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  return FastDiv{static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - d)) / d + 1), l};
}
__device__ __forceinline__ uint32_t fd_div(uint32_t n, FastDiv f) { return (__umulhi(f.mul, n) + n) >> f.shr; }

// A channel-slice view of an NHWC feature map resident in the plan arena.
struct TView {
  void* base;      // buffer base (element 0 of channel 0)
  int N, H, W;     // map geometry
  int ctot;        // channels of the underlying buffer (pixel pitch in elements)
  int c0, c;       // slice
  int f32;         // element type: 0 fp16, 1 fp32
};

struct alignas(16) h16x8 {
  __half2 v[4];
};

__device__ __forceinline__ void unpack8(const h16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ float sat16(float v) { return fminf(fmaxf(v, -H16_MAX), H16_MAX); }
__device__ __forceinline__ h16 to_h16(float v) { return __float2half_rn(sat16(v)); }
__device__ __forceinline__ h16x8 pack8(const float* f) {
  h16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2half2_rn(sat16(f[2 * i]), sat16(f[2 * i + 1]));
  return p;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

}  // namespace bd
