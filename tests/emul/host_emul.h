// host_emul.h -- TEST INFRASTRUCTURE: just enough of the CUDA device environment to compile the word-parallel kernels
// of csrc/rle.cuh with g++ and run them on the CPU, one "thread" after the other (every kernel there is a grid-stride
// loop without intra-block synchronisation; the warp-cooperative ones are compiled out and restated in the harness).
// Logic errors in the kernels show up in the CPU suite; races and launch geometry are what the GPU tests add.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
static thread_local uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, gridDim = {1, 1, 1}, blockDim = {1, 1, 1};

static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int atomicMin(int* a, int v) { int o = *a; if (v < o) *a = v; return o; }
static inline int atomicAdd(int* a, int v) { int o = *a; *a += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { unsigned long long o = *a; *a += v; return o; }
static inline int atomicMax(int* a, int v) { int o = *a; if (v > o) *a = v; return o; }
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }
static inline int atomicExch(int* a, int v) { int o = *a; *a = v; return o; }
using std::max;
using std::min;

// run a kernel written as `for (i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB)` over one block
template <class K, class... A>
static inline void emul_launch(int tpb, K kern, A... args) {
  gridDim.x = 1; blockIdx.x = 0; blockDim.x = tpb;
  for (int t = 0; t < tpb; ++t) { threadIdx.x = t; kern(args...); }
  threadIdx.x = 0;
}
