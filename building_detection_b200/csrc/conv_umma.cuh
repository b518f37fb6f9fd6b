// conv_umma.cuh -- implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
// GEMM view of the convolution:  D[m, n] = sum_{tap, c} A_tap[m, c] * W[tap][n][c]
// where m runs over a (bn x bh x bw) box of output pixels (bn*bh*bw = 128) and A_tap is the NHWC input box
// shifted by the tap offset (dy, dx).  The im2col matrix is never materialised: for every (tap, 64-channel
// chunk) the TMA engine loads the shifted 4-D box straight from the activation tensor into shared memory in
// the 128-byte-swizzled K-major layout tcgen05.mma consumes, and its out-of-bounds zero fill IS the 'same'
// padding (and the channel padding when Cin is not a multiple of 64).  Stride-2 convolutions use up to four
// parity views of the input (one tensor map per (row parity, column parity)), each again a plain tiled map.
//
// Persistent kernel, one CTA per SM, 352 threads, warp-specialised.  Template instances: <0,1> generic ring, <3,1> halo
// path with resident weights (<3,2> split weights, <3,0> tap subset), <5,1> halo path with streamed weights, <4,1> fused
// separable convolution, and the CTA-pair variants (clusters of two, tcgen05.mma.cta_group::2 with M = 256, each CTA
// loads its own pixel tile and half of every weight tile): <6,1> generic ring, <7,1> streamed-weight halo path, <8,1>
// resident-weight halo path.
//   warp 0 (one lane)  TMA producer: A box + W tile per k-block into a ring of shared-memory stages (halo path: one
//                      (16+2) x (8+2) halo box per 64-channel chunk, the weights of all taps resident)
//   warp 1 (one lane)  MMA issuer: tcgen05.mma kind::f16 (fp16 x fp16 -> fp32) into one of TWO TMEM accumulator
//                      stages; tcgen05.commit releases the smem stage / publishes the accumulator
//   warp 10            second MMA issuer for the odd tiles on the 3x3 one-chunk halo path (BD_UMMA_ISSUERS=1: off)
//   warps 2-5          in the fused separable instance: depthwise warps that compute the A tile from a halo box
//   warps 2-9          epilogue, two warps per TMEM lane quadrant (warp w may read TMEM lanes 32*(w%4)..+31 = tile
//                      rows).  The work units are (tile, 64-column chunk) pairs; the two warps of a quadrant take
//                      alternate units and never synchronise with any other warp: tcgen05.ld its 32 rows x 64
//                      columns, + bias (BatchNorm folded, staged in shared memory) / ReLU / residual / ReLU, pack to
//                      fp16 into the warp's own swizzled 4 KB staging tile, one TMA STORE per unit (full-line
//                      writes, ragged tiles and the destination channel slice clipped by the tensor map; the
//                      sub-pixel scatter of the transposed convolutions is a strided output map).  Residual tiles
//                      arrive by per-warp TMA loads issued one unit ahead.  fp32 outputs (2-channel logits,
//                      1-channel gates) are stored directly.
// so the epilogue of tile i overlaps the main loop of tile i+1.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace bd {
namespace umma {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // fp16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int OUT_CHUNK = 64;                             // channels per staging tile / TMA store
constexpr int EPI_TILE_BYTES = 32 * OUT_CHUNK * 2;        // 4 KB: one warp's 32 rows x 64 channels
constexpr int MAX_TAPS = 18;
// halo path: one (16+2) x (8+2) pixel halo box per 64-channel chunk feeds all nine taps of a 3x3 convolution
constexpr int HALO_W = 10, HALO_H = 18, HALO_BYTES = HALO_W * HALO_H * 128, HALO_STAGE = 23 * 1024;
constexpr int EPI_WARPS = 8;                 // two warps per TMEM lane quadrant, each takes half of a column chunk
constexpr int MMA2_WARP = 2 + EPI_WARPS;       // second MMA issuer (tiles of odd index), see the kernel
constexpr int THREADS = 64 + 32 * EPI_WARPS + 32;

using ::bd::FastDiv;  // division by a launch constant (common.cuh)
using ::bd::make_fastdiv;
using ::bd::fd_div;

struct Params {
  int N, Ho, Wo, Cout, Cin;
  FastDiv fd_nt, fd_tw, fd_th;  // dividers by n_tiles, tiles_w, tiles_h
  FastDiv fd_mt;                // divider by the number of pixel tiles (m_fast order)
  int m_total, m_fast;          // m_fast: pixel tiles run fastest in the tile index (all CTAs on one N tile at a time)
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n, n_tiles, total_tiles;
  int block_n, kchunks, ntaps, stages, group, tmem_cols;  // group: k-blocks per shared-memory stage
  int pair;           // spec 6: CTA pairs (cluster of 2): tile index t -> pair tile t >> 1, the CTA's M tile by t & 1
  int nacc, nacc_sh;  // TMEM accumulator stages (2, or 4 when four N tiles fit into the 512 columns) and log2 of it
  int dbg;      // debug switches (BD_UMMA_DBG): 1 = every thread waits for the previous grid before the role split, 2 = no early launch_dependents
  int halo_subset;  // spec 3 with a runtime tap list (kernel instance <3, 0>)
  int prefetch; // 1: the producer prefetches the next tile's activation boxes into L2 (BD_UMMA_PREFETCH=0: off)
  int issuers;  // MMA-issuing warps: 2 = warp 1 takes the even tiles of a CTA, warp MMA2_WARP the odd ones
  int spec;  // 0 generic loops; 1 = 9 taps x 1 chunk, group 3; 2 = 9 taps x 2 chunks, group 2 (fully unrolled loops);
             // 3 = halo path: 3x3 stride 1, one halo box per chunk, weights resident in shared memory;
             // 5 = halo path with streamed weights: halo box per chunk in two slots, the ring holds weight tiles;
             // 4 = fused separable convolution: warps 2-5 compute the depthwise 3x3 of a halo box into the A tile of
             //     the pointwise 1x1 GEMM, warps 6-9 are the epilogue
  int wres_bytes;  // bytes of resident weights (spec 3), 0 otherwise
  int bias_bytes;  // shared-memory bytes of the staged bias vector (all N tiles)
  int aux_bytes;   // spec 4: shared-memory bytes of the staged depthwise weights [9][Cin]
  const h16* dw_w; // spec 4: depthwise 3x3 weights [9][Cin] fp16 (nullptr otherwise)
  int dw_relu;     // spec 4: ReLU on the depthwise input
  int tap_map[MAX_TAPS], tap_dy[MAX_TAPS], tap_dx[MAX_TAPS];
  float* y32;  // fp32 output path (Cout <= 16): direct stores
  int y_ctot, y_c0, y_H, y_W, out_scale, out_oy, out_ox;
  const h16* res;
  int res_ctot, res_c0;
  const float* bias;
  int act_pre, act_post;
  long long* trace;  // optional (debug): per-event clock64 of CTA 0, see tools/umma_trace.py
};

// event record (debug): role r owns trace[r*4096 ...]: [count, (a, b, clock) triples]; plain stores, no atomics
constexpr int TRACE_PER_ROLE = 4096;
__device__ __forceinline__ void trace_ev(const Params& p, int role, int& idx, int a, int b) {
  if (p.trace && blockIdx.x == 0 && idx < 1300) {
    long long* e = p.trace + role * TRACE_PER_ROLE + 1 + 3 * idx;
    e[0] = a; e[1] = b; e[2] = clock64();
    p.trace[role * TRACE_PER_ROLE] = ++idx;
  }
}

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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a tile finishes in microseconds, so ~2^26 polls mean a protocol bug -> trap instead of
// hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 25); ++it) {
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
// Non-blocking phase test.  A wait on an already completed mbarrier still costs a 200-300 cycle shared-memory round
// trip in the single-threaded MMA loop; the halo path (one wait per 36 MMAs) tests the barriers it will need NEXT
// while two thirds of the current chunk's MMAs are queued and only falls back to mbar_wait when the early test failed.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
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
// L2 prefetch of a tensor-map box (no shared-memory destination, no barrier): the producer asks for the boxes of its
// NEXT tile while it loads the current one, so that the loads which would miss to DRAM (the first touch of an
// activation region) find their lines in L2 -- the ring holds ~2000 clk of work, a DRAM round trip under load is longer
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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
// ---- CTA pair (cta_group::2): two CTAs of a cluster on one TPC run ONE M = 256 MMA per instruction; each holds its own
// 128 pixel rows of A and HALF of the weight tile, so a CTA pulls 16 KB less per k-block through L2.  The leader (rank
// 0) issues the MMAs; both producers' TMA bytes complete on the LEADER's full barrier (address with the rank bit
// cleared), the leader's commits arrive on both CTAs' barriers by multicast.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even (leader) CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t leader_bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t leader_bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {  // remote arrive on the leader CTA's barrier
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
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
// same layout, explicit stride between 8-row groups.  The swizzle is a function of the absolute shared-memory
// address (measured: tools/micro/desc_shift.cu), so a descriptor may start at any 128-byte row of a TMA-written
// tile and step between row groups by any multiple of 128 bytes -- which is what reading the nine shifted taps of
// a 3x3 convolution out of one halo tile needs.
__device__ __forceinline__ uint64_t make_sdesc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A and B formats at [7,10) / [10,13) = 0 (fp16; 1 would be
// bf16), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int n, int m = BLOCK_M) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// two fp32 -> packed fp16x2 (lo = first argument), round to nearest, saturating at +-65504 (one F2FP instruction)
__device__ __forceinline__ uint32_t pack2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// all tensor maps of one launch as a single kernel parameter: the producer selects the input parity view by
// pointer arithmetic instead of a chain of selects
struct alignas(64) Maps {
  CUtensorMap a[4];
  CUtensorMap b;
  CUtensorMap y;
  CUtensorMap r;  // residual (same boxes as y), only valid when Params::res != nullptr
};

// tile index -> (N tile, column tile, row tile, image tile); the N tile runs fastest
__device__ __forceinline__ void tile_coords(const Params& p, int tile, int& nt, int& tw, int& th, int& tn) {
  const uint32_t t = static_cast<uint32_t>(tile);
  uint32_t mt;
  if (p.pair) {
    // CTAs 2P and 2P+1 walk tiles t and t+1 (t even): the same N tile, adjacent pixel tiles
    const uint32_t q = t >> 1, mp = fd_div(q, p.fd_nt);
    nt = static_cast<int>(q - mp * p.n_tiles);
    mt = 2u * mp + (t & 1u);
  } else if (p.m_fast) {
    const uint32_t q = fd_div(t, p.fd_mt);
    nt = static_cast<int>(q);
    mt = t - q * p.m_total;
  } else {
    mt = fd_div(t, p.fd_nt);
    nt = static_cast<int>(t - mt * p.n_tiles);
  }
  const uint32_t t2 = fd_div(mt, p.fd_tw);
  tw = static_cast<int>(mt - t2 * p.tiles_w);
  const uint32_t t3 = fd_div(t2, p.fd_th);
  th = static_cast<int>(t2 - t3 * p.tiles_h);
  tn = static_cast<int>(t3);
}

// ------------------------------------------------------------------------------------------ unrolled loops
// The small-channel 3x3 convolutions (Cin <= 128: most of res34 / scse / hrnet's time) have k-blocks whose MMAs take
// 100-300 cycles, so the per-k-block bookkeeping of the generic loops below dominates.  For them one tile's whole
// main loop is unrolled at compile time: tap offsets come straight from the constant bank, descriptor offsets are
// immediates, and a barrier round covers G k-blocks.
struct Ring {
  uint32_t s, ph, off, fb, eb;  // stage index, phase, byte offset of the stage, full / empty barrier addresses
};
__device__ __forceinline__ void ring_advance(Ring& r, const Params& p, uint32_t stage_bytes, uint32_t full0, uint32_t empty0) {
  if (++r.s == static_cast<uint32_t>(p.stages)) { r.s = 0; r.ph ^= 1u; r.off = 0; r.fb = full0; r.eb = empty0; }
  else { r.off += stage_bytes; r.fb += 8u; r.eb += 8u; }
}
// early test of the full barrier of the stage AFTER r
__device__ __forceinline__ uint32_t ring_test_next_full(const Ring& r, const Params& p, uint32_t full0) {
  const bool wrap = r.s + 1u == static_cast<uint32_t>(p.stages);
  return mbar_test(wrap ? full0 : r.fb + 8u, wrap ? r.ph ^ 1u : r.ph);
}
template <int NTAPS, int KCH, int G>
__device__ __forceinline__ void produce_tile(const Maps& maps, const Params& p, Ring& r, uint32_t smem_base, uint32_t sub_bytes,
                                             uint32_t full0, uint32_t empty0, int w0, int h0, int n0, int n_base, int& tr_i,
                                             int tile, bool pf, int pw0, int ph0, int pn0) {
  static_assert((NTAPS * KCH) % G == 0, "group must divide the k-block count");
  const char* maps_a = reinterpret_cast<const char*>(&maps.a[0]);
#pragma unroll
  for (int grp = 0; grp < NTAPS * KCH / G; ++grp) {
    mbar_wait(r.eb, r.ph ^ 1u);
    if (elect_one()) {
      trace_ev(p, 0, tr_i, tile, grp);
      mbar_expect_tx(r.fb, sub_bytes * G);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int kb = grp * G + g, tap = kb / KCH, kc = kb % KCH;  // compile-time
        const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(maps_a + p.tap_map[tap] * sizeof(CUtensorMap));
        const uint32_t dst = smem_base + r.off + g * sub_bytes;
        tma_load_4d(dst, tm, r.fb, kc * BLOCK_K, w0 + p.tap_dx[tap], h0 + p.tap_dy[tap], n0);
        tma_load_3d(dst + A_STAGE_BYTES, &maps.b, r.fb, kc * BLOCK_K, n_base, tap);
        if (tap == NTAPS / 2 && pf) tma_prefetch_4d(tm, kc * BLOCK_K, pw0 + p.tap_dx[tap], ph0 + p.tap_dy[tap], pn0);
      }
    }
    __syncwarp();
    ring_advance(r, p, sub_bytes * G, full0, empty0);
  }
}
template <int NTAPS, int KCH, int G>
__device__ __forceinline__ void mma_tile(const Params& p, Ring& r, uint64_t desc0, uint32_t sub_bytes, uint32_t full0,
                                         uint32_t empty0, uint32_t tacc, uint32_t idesc, uint32_t tfull_bar, int& tr_i,
                                         int tile) {
  const int last_c = p.Cin - (KCH - 1) * BLOCK_K;  // channels in the last chunk of a tap (runtime, 8..64)
#pragma unroll
  for (int grp = 0; grp < NTAPS * KCH / G; ++grp) {
    mbar_wait(r.fb, r.ph);
    tc_fence_after();
    if (elect_one()) {
      trace_ev(p, 1, tr_i, tile, grp);
      const uint64_t d = desc0 + (r.off >> 4);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int kb = grp * G + g, kc = kb % KCH;
        const uint64_t adesc = d + g * (sub_bytes >> 4), bdesc = adesc + (A_STAGE_BYTES >> 4);
        tc_mma_f16(tacc, adesc, bdesc, idesc, kb > 0 ? 1u : 0u);
        if (kc < KCH - 1) {
          tc_mma_f16(tacc, adesc + 2u, bdesc + 2u, idesc, 1u);
          tc_mma_f16(tacc, adesc + 4u, bdesc + 4u, idesc, 1u);
          tc_mma_f16(tacc, adesc + 6u, bdesc + 6u, idesc, 1u);
        } else {
          if (last_c > 16) tc_mma_f16(tacc, adesc + 2u, bdesc + 2u, idesc, 1u);
          if (last_c > 32) tc_mma_f16(tacc, adesc + 4u, bdesc + 4u, idesc, 1u);
          if (last_c > 48) tc_mma_f16(tacc, adesc + 6u, bdesc + 6u, idesc, 1u);
        }
      }
      tc_commit(r.eb);
      if (grp == NTAPS * KCH / G - 1) tc_commit(tfull_bar);
    }
    __syncwarp();
    ring_advance(r, p, sub_bytes * G, full0, empty0);
  }
}

// ------------------------------------------------------------------------------------------ kernel
// The producer and the MMA issuer are single threads executing dependent scalar code: every instruction in
// their per-k-block loops costs several cycles of latency (a k-block's four N=64 MMAs take only 128 cycles), so
// those loops keep their ring position incrementally and contain no integer division.
// KSPEC: 0 = generic / unrolled ring paths (Params::spec 0-2), 3 = halo path, 4 = fused separable convolution; REPS:
// halo path only, 2 = hi/lo-split weights (the nine offsets twice).  Separate instantiations, so that the register
// allocation and scheduling of one path's hot loops do not depend on the code of the others (adding the spec-4 role to
// a single kernel cost the halo path 9 %).
template <int KSPEC, int REPS>
__global__ void __launch_bounds__(THREADS, 1) conv_umma_kernel(const __grid_constant__ Maps maps,
                                                               const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  if (!(p.dbg & 2)) pdl_trigger();  // the next kernel may begin launching; it waits for THIS grid's completion in its own pdl_wait()
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool PAIR = KSPEC == 6 || KSPEC == 7 || KSPEC == 8;  // CTA pairs: generic ring (6), streamed-weight halo path (7),
                                                                 // resident-weight halo path (8)
  constexpr bool HSTREAM = KSPEC == 5 || KSPEC == 7; // halo path with streamed weights
  constexpr bool HALO3 = KSPEC == 3 || KSPEC == 8;   // halo path with resident weights
  auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t acc) {
    if (PAIR) tc_mma_f16_2sm(d, ad, bd, id, acc);
    else tc_mma_f16(d, ad, bd, id, acc);
  };
  auto commit = [&](uint32_t bar) {  // pair: arrives on the barrier at this offset in BOTH CTAs
    if (PAIR) tc_commit_2sm(bar);
    else tc_commit(bar);
  };
  const uint32_t pair_rank = PAIR ? cluster_ctarank() : 0u;
  const uint32_t b_bytes = static_cast<uint32_t>(PAIR ? p.block_n / 2 : p.block_n) * 128u;  // pair: half of the weight tile
  const uint32_t sub_bytes = A_STAGE_BYTES + b_bytes;                       // one k-block: A box + W tile
  const uint32_t stage_bytes = HALO3 ? static_cast<uint32_t>(HALO_STAGE)
                               : HSTREAM ? b_bytes * static_cast<uint32_t>(p.group)   // `group` weight tiles (taps)
                               : (KSPEC == 4) ? static_cast<uint32_t>(HALO_STAGE) + b_bytes  // halo box + pointwise W tile
                                             : sub_bytes * static_cast<uint32_t>(p.group);  // a stage holds `group` k-blocks
  const uint32_t wres0 = smem_base + static_cast<uint32_t>(p.stages) * stage_bytes;      // resident weights (spec 3)
  const uint32_t out0 = wres0 + static_cast<uint32_t>(p.wres_bytes);                      // 8 per-warp staging tiles
  const uint32_t n_epi_warps = (KSPEC == 4) ? EPI_WARPS / 2 : EPI_WARPS;              // spec 4: warps 2-5 are depthwise warps
  const uint32_t res0 = out0 + n_epi_warps * EPI_TILE_BYTES;                         // per-warp residual tiles (if any)
  const uint32_t bias0 = res0 + (p.res ? n_epi_warps * EPI_TILE_BYTES : 0u);         // bias of all N tiles, fp32
  const uint32_t aux0 = bias0 + static_cast<uint32_t>(p.bias_bytes);                 // depthwise weights (spec 4)
  const uint32_t bar_base = aux0 + static_cast<uint32_t>(p.aux_bytes);               // 8-byte slots
  const uint32_t full0 = bar_base, empty0 = bar_base + 8u * p.stages;
  const uint32_t tfull0 = bar_base + 16u * p.stages, tempty0 = tfull0 + 32u;  // up to 4 accumulator stages; tempty: [stage][epilogue group]
  const uint32_t rbar0 = tempty0 + 64u;  // one per epilogue warp
  const uint32_t wbar = rbar0 + 8u * EPI_WARPS;
  const uint32_t afull0 = wbar + 16u, aempty0 = afull0 + 16u;  // spec 4: two depthwise-output (A tile) slots
  const uint32_t holder = aempty0 + 16u;
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (holder - smem_base));

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, (KSPEC == 4) ? 5 : 1);  // spec 4: four depthwise warps + the MMA commit free a stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(afull0 + 8u * a, HSTREAM ? 1 : 4);   // one arrival per depthwise warp (spec 5: the producer's expect_tx)
      mbar_init(aempty0 + 8u * a, 1);  // MMA commit
    }
    for (int a = 0; a < 4; ++a) {
      mbar_init(tfull0 + 8u * a, 1);
      // one arrival per warp of epilogue group 0 / group 1 (pair: the warps of BOTH CTAs arrive on the leader's barrier)
      mbar_init(tempty0 + 16u * a, (PAIR ? 2 : 1) * EPI_WARPS / 2);
      mbar_init(tempty0 + 16u * a + 8u, (PAIR ? 2 : 1) * EPI_WARPS / 2);
    }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(rbar0 + 8u * w, 1);
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == 1) {
    if (PAIR) {  // the same warp of both CTAs allocates the pair's columns
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;
  // Barrier set-up and TMEM allocation above overlap the previous kernel's tail.  Each role calls pdl_wait() itself,
  // after whatever it can do on CONSTANT data (resident weights, bias, depthwise weights are never written by a
  // kernel) and before its first access to activations; the MMA warp touches only shared memory and TMEM.

  if (p.dbg & 1) pdl_wait();
  int tr_i = 0, tr_j = 0;  // debug trace cursors

  // Warps 0 and 1 run their loops with all 32 lanes (every value is warp-uniform, so the compiler keeps the ring
  // state in uniform registers next to the UTMALDG / UTCHMMA operands); one elected lane issues.
  if (warp == 0) {
    // ---------------- TMA producer
    uint32_t s = 0, sub = 0, ph = 0, a_s = smem_base, fb = full0, eb = empty0;  // ring position, kept incrementally
    const char* maps_a = reinterpret_cast<const char*>(&maps.a[0]);
    Ring ring{0u, 0u, 0u, full0, empty0};
    uint32_t hcount = 0;  // spec 5: running chunk counter (halo slot / parity)
    if (HALO3 && elect_one()) {  // all weight tiles once: [chunk][tap] blocks of block_n x 128 B
      uint32_t dst = wres0;
      if (PAIR) {  // each CTA keeps its half of the columns; both halves complete on the leader's barrier
        if (pair_rank == 0) mbar_expect_tx(wbar, 2u * static_cast<uint32_t>(p.wres_bytes));
        for (int c = 0; c < p.Cin; c += BLOCK_K)
          for (int tap = 0; tap < p.ntaps; ++tap, dst += b_bytes)
            tma_load_3d_2sm(dst, &maps.b, wbar & PEER_BIT_MASK, c, static_cast<int>(pair_rank) * (p.block_n / 2), tap);
      } else {
        mbar_expect_tx(wbar, static_cast<uint32_t>(p.wres_bytes));
        for (int c = 0; c < p.Cin; c += BLOCK_K)
          for (int tap = 0; tap < p.ntaps; ++tap, dst += b_bytes) tma_load_3d(dst, &maps.b, wbar, c, 0, tap);
      }
    }
    __syncwarp();
    pdl_wait();  // activations of the previous kernel from here on
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int nt, tw, th, tn;
      tile_coords(p, tile, nt, tw, th, tn);
      const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn, n_base = nt * p.block_n;
      // the next tile of this CTA: its activation boxes are prefetched into L2 while this tile loads
      int pw0 = 0, ph0 = 0, pn0 = 0;
      const bool pf = p.prefetch && tile + static_cast<int>(gridDim.x) < p.total_tiles;
      if (pf) {
        int nt2, tw2, th2, tn2;
        tile_coords(p, tile + gridDim.x, nt2, tw2, th2, tn2);
        pw0 = tw2 * p.bw; ph0 = th2 * p.bh; pn0 = tn2 * p.bn;
      }
      if (HALO3) {  // one halo box per 64-channel chunk
        for (int c = 0; c < p.Cin; c += BLOCK_K) {
          mbar_wait(ring.eb, ring.ph ^ 1u);
          if (elect_one()) {
            trace_ev(p, 0, tr_i, tile, c);
            if (PAIR) {
              if (pair_rank == 0) mbar_expect_tx(ring.fb, 2u * HALO_BYTES);
              tma_load_4d_2sm(smem_base + ring.off, &maps.a[0], ring.fb & PEER_BIT_MASK, c, w0 - 1, h0 - 1, n0);
            } else {
              mbar_expect_tx(ring.fb, HALO_BYTES);
              tma_load_4d(smem_base + ring.off, &maps.a[0], ring.fb, c, w0 - 1, h0 - 1, n0);
            }
            if (pf) tma_prefetch_4d(&maps.a[0], c, pw0 - 1, ph0 - 1, pn0);
          }
          __syncwarp();
          ring_advance(ring, p, HALO_STAGE, full0, empty0);
        }
        continue;
      }
      if (HSTREAM) {
        // halo path with STREAMED weights (3x3 layers whose weights do not fit into shared memory): per 64-channel
        // chunk one halo box into one of two slots (it feeds all nine taps), then the nine weight tiles through the
        // ring, `group` taps per stage
        for (int c = 0; c < p.Cin; c += BLOCK_K, ++hcount) {
          const uint32_t slot = hcount & 1u;
          mbar_wait(aempty0 + 8u * slot, ((hcount >> 1) & 1u) ^ 1u);
          if (elect_one()) {
            trace_ev(p, 0, tr_i, tile, c);
            if (PAIR) {  // both halo boxes complete on the leader's barrier
              if (pair_rank == 0) mbar_expect_tx(afull0 + 8u * slot, 2u * HALO_BYTES);
              tma_load_4d_2sm(wres0 + slot * HALO_STAGE, &maps.a[0], (afull0 + 8u * slot) & PEER_BIT_MASK, c, w0 - 1, h0 - 1, n0);
            } else {
              mbar_expect_tx(afull0 + 8u * slot, HALO_BYTES);
              tma_load_4d(wres0 + slot * HALO_STAGE, &maps.a[0], afull0 + 8u * slot, c, w0 - 1, h0 - 1, n0);
            }
            if (pf) tma_prefetch_4d(&maps.a[0], c, pw0 - 1, ph0 - 1, pn0);
          }
          __syncwarp();
          for (int t0 = 0; t0 < 9; t0 += p.group) {
            mbar_wait(ring.eb, ring.ph ^ 1u);
            if (elect_one()) {
              if (PAIR) {  // each CTA streams its half of every weight tile
                if (pair_rank == 0) mbar_expect_tx(ring.fb, 2u * stage_bytes);
                for (int g = 0; g < p.group; ++g)
                  tma_load_3d_2sm(smem_base + ring.off + g * b_bytes, &maps.b, ring.fb & PEER_BIT_MASK, c,
                                  n_base + static_cast<int>(pair_rank) * (p.block_n / 2), t0 + g);
              } else {
                mbar_expect_tx(ring.fb, stage_bytes);
                for (int g = 0; g < p.group; ++g)
                  tma_load_3d(smem_base + ring.off + g * b_bytes, &maps.b, ring.fb, c, n_base, t0 + g);
              }
            }
            __syncwarp();
            ring_advance(ring, p, stage_bytes, full0, empty0);
          }
        }
        continue;
      }
      if ((KSPEC == 4)) {  // per 64-channel chunk: the halo box of the depthwise input + the pointwise weight tile
        for (int c = 0; c < p.Cin; c += BLOCK_K) {
          mbar_wait(ring.eb, ring.ph ^ 1u);
          if (elect_one()) {
            trace_ev(p, 0, tr_i, tile, c);
            mbar_expect_tx(ring.fb, HALO_BYTES + b_bytes);
            tma_load_4d(smem_base + ring.off, &maps.a[0], ring.fb, c, w0 - 1, h0 - 1, n0);
            tma_load_3d(smem_base + ring.off + HALO_STAGE, &maps.b, ring.fb, c, n_base, 0);
            if (pf) tma_prefetch_4d(&maps.a[0], c, pw0 - 1, ph0 - 1, pn0);
          }
          __syncwarp();
          ring_advance(ring, p, stage_bytes, full0, empty0);
        }
        continue;
      }
      if (KSPEC == 0 && p.spec == 1) { produce_tile<9, 1, 3>(maps, p, ring, smem_base, sub_bytes, full0, empty0, w0, h0, n0, n_base, tr_i, tile, pf, pw0, ph0, pn0); continue; }
      if (KSPEC == 0 && p.spec == 2) { produce_tile<9, 2, 2>(maps, p, ring, smem_base, sub_bytes, full0, empty0, w0, h0, n0, n_base, tr_i, tile, pf, pw0, ph0, pn0); continue; }
      const int pf_tap = p.ntaps >> 1;  // the centre tap of a 3x3: its box covers the region all taps read (+- 1 pixel)
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(maps_a + p.tap_map[tap] * sizeof(CUtensorMap));
        const int cx = w0 + p.tap_dx[tap], cy = h0 + p.tap_dy[tap];
        for (int c = 0; c < p.Cin; c += BLOCK_K) {
          if (sub == 0) mbar_wait(eb, ph ^ 1u);
          if (elect_one()) {
            trace_ev(p, 0, tr_i, tile, tap);
            if (PAIR) {
              // both CTAs' bytes (own pixel tile + own half of the weight tile) complete on the LEADER's full barrier
              if (sub == 0 && pair_rank == 0) mbar_expect_tx(fb, 2u * stage_bytes);
              tma_load_4d_2sm(a_s, tm, fb & PEER_BIT_MASK, c, cx, cy, n0);
              tma_load_3d_2sm(a_s + A_STAGE_BYTES, &maps.b, fb & PEER_BIT_MASK, c,
                              n_base + static_cast<int>(pair_rank) * (p.block_n / 2), tap);
            } else {
              if (sub == 0) mbar_expect_tx(fb, stage_bytes);
              tma_load_4d(a_s, tm, fb, c, cx, cy, n0);
              tma_load_3d(a_s + A_STAGE_BYTES, &maps.b, fb, c, n_base, tap);
            }
            if (pf && tap == pf_tap) tma_prefetch_4d(tm, c, pw0 + p.tap_dx[tap], ph0 + p.tap_dy[tap], pn0);
          }
          __syncwarp();
          a_s += sub_bytes;
          if (++sub == static_cast<uint32_t>(p.group)) {
            sub = 0;
            if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; a_s = smem_base; fb = full0; eb = empty0; }
            else { fb += 8u; eb += 8u; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == MMA2_WARP) {
    // ---------------- MMA issuers.  A single thread issues a tcgen05.mma every ~50 clk but spends 450-700 clk per
    // tile on barrier round trips and bookkeeping, during which the tensor pipe drains (a 3x3 64->64 tile is only
    // 36 MMAs).  With two issuers, warp 1 takes the even tiles of this CTA (TMEM stage 0) and warp MMA2_WARP the odd
    // ones (stage 1): while one sits between two tiles the other keeps the pipe fed.  Both walk the shared-memory
    // ring in tile order and skip the stages of the other's tiles; each commits the stages it consumed.
    const uint32_t issuer = warp == 1 ? 0u : 1u;
    const bool dual = p.issuers == 2;
    if (issuer == 1) tr_i = 1 << 20;  // (debug trace: issuer 0 only)
    const uint32_t idesc = PAIR ? make_idesc(p.block_n, 2 * BLOCK_M) : make_idesc(p.block_n);
    const uint64_t desc0 = make_sdesc(smem_base);
    const uint32_t dsub = sub_bytes >> 4;  // descriptor start-address field counts 16-byte units
    uint32_t s = 0, sub = 0, ph = 0, doff = 0, fb = full0, eb = empty0, ti = 0;
    Ring ring{0u, 0u, 0u, full0, empty0};
    uint32_t f_ready = 0, te_ready = 0;  // early test results: next full barrier / this tile's accumulator stage
    uint32_t it4 = 0;                    // spec 4: running k-block counter (A slot / parity)
    const bool single_unit = p.block_n <= OUT_CHUNK;  // one epilogue unit per tile
    const int num_kb_all = p.ntaps * p.kchunks;
    // (pair: the leader CTA issues for both; the peer's MMA warp only allocated its half of the tensor memory)
    for (int tile = blockIdx.x; tile < p.total_tiles && (dual || issuer == 0) && !(PAIR && pair_rank != 0);
         tile += gridDim.x, ++ti) {
      // accumulator stage of this tile and the phase of its barriers (two stages; four with BD_UMMA_NACC=4 and N <= 128:
      // an issuer that owns every other tile then alternates between two stages of its own)
      const uint32_t a = ti & static_cast<uint32_t>(p.nacc - 1), aph = (ti >> p.nacc_sh) & 1u;
      if (dual && (ti & 1u) != issuer) {
        // The other issuer's tile: step over its ring stages WITHOUT touching their barriers.  An mbarrier parity wait
        // is only sound when the waiter has seen the previous phase of that barrier complete, so the two issuers must
        // never share a stage: the host enables two issuers only when a tile is ONE stage and the ring has an EVEN
        // number of stages -- then issuer 0 owns the even stages and issuer 1 the odd ones, and each sees every phase
        // of its own barriers in order.  (Round 1 ran odd rings, 5 or 3 stages for 64->64: a stage alternated between
        // the issuers, an issuer that ran ahead could read "parity k done" from phase k-2 while the other's data was
        // still landing -- the intermittent launch failures.  Making the skipping issuer wait on the skipped full
        // barriers instead deadlocks when it falls a whole ring behind: tried, times out at batch 32.)
        if (HALO3) {
          for (int c = 0; c < p.kchunks; ++c) ring_advance(ring, p, HALO_STAGE, full0, empty0);
        } else if (KSPEC == 0 && p.spec == 1) {
          for (int g2 = 0; g2 < 3; ++g2) ring_advance(ring, p, sub_bytes * 3, full0, empty0);
        } else if (KSPEC == 0 && p.spec == 2) {
          for (int g2 = 0; g2 < 9; ++g2) ring_advance(ring, p, sub_bytes * 2, full0, empty0);
        } else {
          for (int kb = 0; kb < num_kb_all; ++kb) {
            doff += dsub;
            if (++sub == static_cast<uint32_t>(p.group)) {
              sub = 0;
              if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; doff = 0; fb = full0; eb = empty0; }
              else { fb += 8u; eb += 8u; }
            }
          }
        }
        continue;
      }
      if ((KSPEC == 4)) {
        mbar_wait(tempty0 + 16u * a + 8u, aph ^ 1u);  // the single epilogue group (warps 6-9) is group 1
        tc_fence_after();
        const uint32_t tacc4 = tmem_base + a * static_cast<uint32_t>(p.block_n);
        for (int c = 0; c < p.Cin; c += BLOCK_K, ++it4) {
          const uint32_t slot = it4 & 1u;
          mbar_wait(ring.fb, ring.ph);                            // pointwise weight tile has landed
          mbar_wait(afull0 + 8u * slot, (it4 >> 1) & 1u);         // depthwise warps have written the A tile
          tc_fence_after();
          if (elect_one()) {
            trace_ev(p, 1, tr_i, tile, c);
            const uint64_t adesc = make_sdesc(wres0 + slot * A_STAGE_BYTES);
            const uint64_t bdesc = make_sdesc(smem_base + ring.off + HALO_STAGE);
            tc_mma_f16(tacc4, adesc, bdesc, idesc, c > 0 ? 1u : 0u);
            tc_mma_f16(tacc4, adesc + 2u, bdesc + 2u, idesc, 1u);
            tc_mma_f16(tacc4, adesc + 4u, bdesc + 4u, idesc, 1u);
            tc_mma_f16(tacc4, adesc + 6u, bdesc + 6u, idesc, 1u);
            tc_commit(aempty0 + 8u * slot);
            tc_commit(ring.eb);
            if (c + BLOCK_K >= p.Cin) tc_commit(tfull0 + 8u * a);
          }
          __syncwarp();
          ring_advance(ring, p, stage_bytes, full0, empty0);
        }
        continue;
      }
      // The epilogue groups that read accumulator stage a two tiles ago have drained it.  A tile of one 64-column
      // chunk belongs to ONE group (group = tile parity = a), wider tiles to both; each (stage, group) barrier
      // completes one phase per tile on that stage, so the parity is the same for all of them.
      if (!te_ready) {
        if (single_unit) mbar_wait(tempty0 + 16u * a + 8u * (ti & 1u), aph ^ 1u);
        else { mbar_wait(tempty0 + 16u * a, aph ^ 1u); mbar_wait(tempty0 + 16u * a + 8u, aph ^ 1u); }
      }
      tc_fence_after();
      const uint32_t tacc = tmem_base + a * static_cast<uint32_t>(p.block_n);
      // the NEXT tile's accumulator stage / phase, tested early during this tile's last k-block
      const uint32_t te_bar = tempty0 + 16u * ((ti + 1u) & static_cast<uint32_t>(p.nacc - 1)),
                     te_par = (((ti + 1u) >> p.nacc_sh) & 1u) ^ 1u;
      if (HSTREAM) {
        for (int c = p.Cin; c > 0; c -= BLOCK_K, ++it4) {  // c = channels left; it4 = running chunk counter
          const uint32_t slot = it4 & 1u;
          mbar_wait(afull0 + 8u * slot, (it4 >> 1) & 1u);  // halo box of this chunk
          const uint64_t hdesc = make_sdesc_sbo(wres0 + slot * HALO_STAGE, HALO_W * 128);
          for (int t0 = 0; t0 < 9; t0 += p.group) {
            if (!f_ready) mbar_wait(ring.fb, ring.ph);
            tc_fence_after();
            if (elect_one()) {
              if (t0 == 0) trace_ev(p, 1, tr_i, tile, c);
              const uint64_t bdesc0 = make_sdesc(smem_base + ring.off);
              for (int g = 0; g < p.group; ++g) {
                const int tap = t0 + g;
                const uint64_t adesc = hdesc + static_cast<uint32_t>(((tap / 3) * HALO_W + tap % 3) * 8);
                const uint64_t bd = bdesc0 + static_cast<uint32_t>(g) * (b_bytes >> 4);
                if (PAIR) {
                  tc_mma_f16_2sm(tacc, adesc, bd, idesc, (tap > 0 || c < p.Cin) ? 1u : 0u);
                  if (c > 16) tc_mma_f16_2sm(tacc, adesc + 2u, bd + 2u, idesc, 1u);
                  if (c > 32) tc_mma_f16_2sm(tacc, adesc + 4u, bd + 4u, idesc, 1u);
                  if (c > 48) tc_mma_f16_2sm(tacc, adesc + 6u, bd + 6u, idesc, 1u);
                } else {
                  tc_mma_f16(tacc, adesc, bd, idesc, (tap > 0 || c < p.Cin) ? 1u : 0u);
                  if (c > 16) tc_mma_f16(tacc, adesc + 2u, bd + 2u, idesc, 1u);
                  if (c > 32) tc_mma_f16(tacc, adesc + 4u, bd + 4u, idesc, 1u);
                  if (c > 48) tc_mma_f16(tacc, adesc + 6u, bd + 6u, idesc, 1u);
                }
              }
              if (PAIR) {  // multicast: the stage / slot / accumulator barriers of BOTH CTAs
                tc_commit_2sm(ring.eb);
                if (t0 + p.group >= 9) {
                  tc_commit_2sm(aempty0 + 8u * slot);
                  if (c <= BLOCK_K) tc_commit_2sm(tfull0 + 8u * a);
                }
              } else {
                tc_commit(ring.eb);                                   // weight stage free once these MMAs have read it
                if (t0 + p.group >= 9) {
                  tc_commit(aempty0 + 8u * slot);                     // ... and the halo slot after the ninth tap
                  if (c <= BLOCK_K) tc_commit(tfull0 + 8u * a);       // accumulator complete
                }
              }
            }
            __syncwarp();
            // these MMAs are queued: test what the next round will wait for
            f_ready = ring_test_next_full(ring, p, full0);
            ring_advance(ring, p, stage_bytes, full0, empty0);
          }
        }
        te_ready = 0u;
        continue;
      }
      if (HALO3) {
        if (ti < 2) mbar_wait(wbar, 0);  // (each issuer before its first tile; completes once, parity 0)
        const uint32_t dkb = b_bytes >> 4;   // descriptor units per weight tile
        uint64_t bdesc = make_sdesc(wres0);  // weights are laid out [chunk][tap]: a running descriptor
        for (int c = p.Cin; c > 0; c -= BLOCK_K) {  // c = channels left
          if (!f_ready) mbar_wait(ring.fb, ring.ph);
          tc_fence_after();
          const uint64_t hdesc = make_sdesc_sbo(smem_base + ring.off, HALO_W * 128);
          if (REPS == 0) {
            // tap subset (sub-pixel phases of up-sample+conv / transposed convolutions: 2 or 4 taps of the 3x3
            // window, any order): runtime tap list, halo offset of tap t from its (dy, dx)
            if (elect_one()) {
              trace_ev(p, 1, tr_i, tile, c);
              for (int t = 0; t < p.ntaps; ++t) {
                const uint64_t adesc = hdesc + static_cast<uint32_t>(((p.tap_dy[t] + 1) * HALO_W + p.tap_dx[t] + 1) * 8);
                const uint64_t bd = bdesc + static_cast<uint32_t>(t) * dkb;
                mma(tacc, adesc, bd, idesc, (t > 0 || c < p.Cin) ? 1u : 0u);
                if (c > 16) mma(tacc, adesc + 2u, bd + 2u, idesc, 1u);
                if (c > 32) mma(tacc, adesc + 4u, bd + 4u, idesc, 1u);
                if (c > 48) mma(tacc, adesc + 6u, bd + 6u, idesc, 1u);
              }
              commit(ring.eb);
              if (c <= BLOCK_K) commit(tfull0 + 8u * a);
              trace_ev(p, 1, tr_i, tile, -1);
            }
            __syncwarp();
            bdesc += static_cast<uint32_t>(p.ntaps) * dkb;
          }
          constexpr int reps = REPS;  // 1, or 2 with hi/lo-split weights (the nine offsets twice); 0: tap subset above
#pragma unroll
          for (int rep = 0; rep < reps; ++rep) {
            const bool last_rep = rep + 1 == reps;
            if (elect_one()) {
              if (rep == 0) trace_ev(p, 1, tr_i, tile, c);
#pragma unroll
              for (int tap = 0; tap < 6; ++tap) {
                // tap (kh, kw) reads the halo rows shifted by kh halo rows and kw pixels
                const uint64_t adesc = hdesc + (((tap / 3) * HALO_W + tap % 3) * 128 >> 4);
                const uint64_t bd = bdesc + static_cast<uint32_t>(tap) * dkb;
                mma(tacc, adesc, bd, idesc, (tap > 0 || rep > 0 || c < p.Cin) ? 1u : 0u);
                if (c > 16) mma(tacc, adesc + 2u, bd + 2u, idesc, 1u);
                if (c > 32) mma(tacc, adesc + 4u, bd + 4u, idesc, 1u);
                if (c > 48) mma(tacc, adesc + 6u, bd + 6u, idesc, 1u);
              }
            }
            __syncwarp();
            if (last_rep && !dual) {
              // most of this chunk's MMAs are queued: test what the next chunk / tile will wait for
              f_ready = ring_test_next_full(ring, p, full0);
              if (c <= BLOCK_K)
                te_ready = single_unit ? mbar_test(te_bar + 8u * ((ti + 1u) & 1u), te_par)
                                       : (mbar_test(te_bar, te_par) & mbar_test(te_bar + 8u, te_par));
            }
            if (elect_one()) {
#pragma unroll
              for (int tap = 6; tap < 9; ++tap) {
                const uint64_t adesc = hdesc + (((tap / 3) * HALO_W + tap % 3) * 128 >> 4);
                const uint64_t bd = bdesc + static_cast<uint32_t>(tap) * dkb;
                mma(tacc, adesc, bd, idesc, 1u);
                if (c > 16) mma(tacc, adesc + 2u, bd + 2u, idesc, 1u);
                if (c > 32) mma(tacc, adesc + 4u, bd + 4u, idesc, 1u);
                if (c > 48) mma(tacc, adesc + 6u, bd + 6u, idesc, 1u);
              }
              if (last_rep) {
                commit(ring.eb);
                if (c <= BLOCK_K) commit(tfull0 + 8u * a);
                trace_ev(p, 1, tr_i, tile, -1);
              }
            }
            __syncwarp();
            bdesc += 9u * dkb;
          }
          ring_advance(ring, p, HALO_STAGE, full0, empty0);
        }
        continue;
      }
      if (KSPEC == 0 && p.spec == 1) { mma_tile<9, 1, 3>(p, ring, desc0, sub_bytes, full0, empty0, tacc, idesc, tfull0 + 8u * a, tr_i, tile); f_ready = te_ready = 0u; continue; }
      if (KSPEC == 0 && p.spec == 2) { mma_tile<9, 2, 2>(p, ring, desc0, sub_bytes, full0, empty0, tacc, idesc, tfull0 + 8u * a, tr_i, tile); f_ready = te_ready = 0u; continue; }
      uint32_t accum = 0;
      for (int tap = 0; tap < p.ntaps; ++tap) {
        for (int c = p.Cin; c > 0; c -= BLOCK_K) {  // c = channels left in this tap
          if (sub == 0) {
            mbar_wait(fb, ph);
            tc_fence_after();
          }
          if (elect_one()) {
            trace_ev(p, 1, tr_i, tile, tap);
            const uint64_t adesc = desc0 + doff, bdesc = adesc + (A_STAGE_BYTES >> 4);
            // 16 fp16 = 32 bytes inside the swizzle atom per MMA: +2 in the (addr >> 4) field
            if (PAIR) {
              tc_mma_f16_2sm(tacc, adesc, bdesc, idesc, accum);
              if (c > 16) tc_mma_f16_2sm(tacc, adesc + 2u, bdesc + 2u, idesc, 1u);
              if (c > 32) tc_mma_f16_2sm(tacc, adesc + 4u, bdesc + 4u, idesc, 1u);
              if (c > 48) tc_mma_f16_2sm(tacc, adesc + 6u, bdesc + 6u, idesc, 1u);
              if (sub + 1 == static_cast<uint32_t>(p.group)) tc_commit_2sm(eb);  // frees the stage in both CTAs
            } else {
              tc_mma_f16(tacc, adesc, bdesc, idesc, accum);
              if (c > 16) tc_mma_f16(tacc, adesc + 2u, bdesc + 2u, idesc, 1u);
              if (c > 32) tc_mma_f16(tacc, adesc + 4u, bdesc + 4u, idesc, 1u);
              if (c > 48) tc_mma_f16(tacc, adesc + 6u, bdesc + 6u, idesc, 1u);
              if (sub + 1 == static_cast<uint32_t>(p.group)) tc_commit(eb);  // frees the stage once these MMAs have read it
            }
          }
          __syncwarp();
          accum = 1u;
          doff += dsub;
          if (++sub == static_cast<uint32_t>(p.group)) {
            sub = 0;
            if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; doff = 0; fb = full0; eb = empty0; }
            else { fb += 8u; eb += 8u; }
          }
        }
      }
      if (elect_one()) {  // accumulator complete (pair: in both CTAs)
        if (PAIR) tc_commit_2sm(tfull0 + 8u * a);
        else tc_commit(tfull0 + 8u * a);
      }
      __syncwarp();
    }
  } else if ((KSPEC == 4) && warp < 6) {
    // ---------------- depthwise warps 2..5 (spec 4): A tile of the pointwise GEMM = depthwise 3x3 of the halo box.
    // thread = (4-channel group g4 of the 64-channel chunk, 4x4 output patch of the 8 x 16 tile): a 6x6 window of
    // 8-byte vectors in registers serves 16 outputs (2.25 shared-memory loads per output instead of 9 -- the
    // shared-memory pipe is shared with the TMA writes and the tensor core's operand reads and is what bounds this
    // kernel).  The nine taps run on packed half2 FMAs in exactly the order and rounding of k::dwconv3x3_kernel
    // (fp16 accumulate, kh-major), so the fused result is bit-identical to the two-kernel sequence.  The 16 threads
    // of a half-warp phase read / write the 16 x 8 bytes of ONE 128-byte pixel row: conflict-free.
    const int dt = threadIdx.x - 64;  // 0..127
    const int g4 = dt & 15, patch = dt >> 4, px0 = (patch & 1) * 4, py0 = (patch >> 1) * 4;
    {  // stage the depthwise weights [9][Cin] (constant data: no dependence on the previous kernel's output)
      const uint4* src = reinterpret_cast<const uint4*>(p.dw_w);
      uint4* dst = reinterpret_cast<uint4*>(gen_base + (aux0 - smem_base));
      for (int i = dt; i < 9 * p.Cin / 8; i += 128) dst[i] = __ldg(src + i);
      asm volatile("bar.sync 2, 128;" ::: "memory");
    }
    const __half2 zero2 = __float2half2_rn(0.0f);
    union U { uint2 v; __half2 h[2]; };
    Ring ring{0u, 0u, 0u, full0, empty0};
    uint32_t it = 0;
    const uint32_t sub8 = (g4 & 1) * 8;  // byte offset of my 4 channels inside their 16-byte chunk
    const int gc = g4 >> 1;              // 16-byte chunk of my channels
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int c = 0; c < p.Cin; c += BLOCK_K, ++it) {
        const uint32_t slot = it & 1u;
        U wp[9];
#pragma unroll
        for (int k = 0; k < 9; ++k)
          wp[k].v = *reinterpret_cast<const uint2*>(gen_base + (aux0 - smem_base) + (static_cast<size_t>(k) * p.Cin + c + g4 * 4) * 2);
        if (dt == 0) trace_ev(p, 3, tr_j, tile, 1);
        mbar_wait(ring.fb, ring.ph);                                // halo box has landed
        mbar_wait(aempty0 + 8u * slot, ((it >> 1) & 1u) ^ 1u);      // the MMAs that read this A slot are done
        if (dt == 0) trace_ev(p, 3, tr_j, tile, 3);
        const uint8_t* halo = gen_base + ring.off;
        uint8_t* atile = gen_base + (wres0 - smem_base) + slot * A_STAGE_BYTES;
        U win[6][6];  // halo rows py0 .. py0+5, columns px0 .. px0+5
#pragma unroll
        for (int hy = 0; hy < 6; ++hy)
#pragma unroll
          for (int hx = 0; hx < 6; ++hx) {
            const int P = (py0 + hy) * HALO_W + px0 + hx;  // halo pixel; its 16-byte chunks are XOR-swizzled by P % 8
            win[hy][hx].v = *reinterpret_cast<const uint2*>(halo + P * 128 + ((gc ^ (P & 7)) << 4) + sub8);
          }
        if (p.dw_relu) {
#pragma unroll
          for (int hy = 0; hy < 6; ++hy)
#pragma unroll
            for (int hx = 0; hx < 6; ++hx) {
              win[hy][hx].h[0] = __hmax2(win[hy][hx].h[0], zero2);
              win[hy][hx].h[1] = __hmax2(win[hy][hx].h[1], zero2);
            }
        }
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
            *reinterpret_cast<uint2*>(atile + m * 128 + ((gc ^ (m & 7)) << 4) + sub8) = acc.v;
          }
        if (dt == 0) trace_ev(p, 3, tr_j, tile, 5);
        fence_async_smem();  // generic-proxy writes of the A tile -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(afull0 + 8u * slot);
          mbar_arrive(ring.eb);  // this warp is done with the halo box
        }
        if (dt == 0) trace_ev(p, 3, tr_j, tile, 0);
        ring_advance(ring, p, stage_bytes, full0, empty0);
      }
    }
  } else {
    // ---------------- epilogue warps (2..9; 6..9 in spec 4): warp w reads TMEM lane quadrant w % 4 (hardware rule).
    // Units = (tile, 64-column chunk) in launch order; with two groups the two warps of a quadrant take alternate
    // units (group = unit parity), with one group (spec 4) warps 6..9 take them all.
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;  // epilogue group of this warp: 0 / 1
    const int ew = warp - ((KSPEC == 4) ? 6 : 2);  // 0 .. (epilogue warps)-1: staging / residual tile, residual barrier
    const int r = q * 32 + lane;       // tile row = TMEM lane
    const int ustep = (KSPEC == 4) ? 1 : 2;                       // number of epilogue groups = unit stride
    const int et = threadIdx.x - ((KSPEC == 4) ? 192 : 64);       // 0 .. 32 * (epilogue warps) - 1
    const int n_epi = (KSPEC == 4) ? 128 : 32 * EPI_WARPS;        // epilogue threads
    // the quadrant's 32 rows as a (bw x qbh x qbn) sub-box of the tile box, starting at row qh of image qn
    const int hrows = 32 / p.bw, hq = q * hrows;
    const int qh = hq % p.bh, qn = hq / p.bh;
    const int wl = r % p.bw, hl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);  // (fp32 path: per-row coordinates)
    const uint32_t stg = out0 + static_cast<uint32_t>(ew) * EPI_TILE_BYTES;   // this warp's staging tile
    const uint32_t rsb = res0 + static_cast<uint32_t>(ew) * EPI_TILE_BYTES;   // this warp's residual tile
    uint8_t* srow = gen_base + (stg - smem_base) + lane * 128;
    const uint8_t* rsrow = gen_base + (rsb - smem_base) + lane * 128;
    const uint32_t rbar = rbar0 + 8u * static_cast<uint32_t>(ew);
    const int swz = lane & 7;  // 128-byte swizzle: 16-byte chunk index XOR (row % 8)
    const bool has_res = p.res != nullptr;
    const int nch = (p.block_n + OUT_CHUNK - 1) / OUT_CHUNK;  // units per tile
    const float lo_pre = p.act_pre == 1 ? 0.0f : -INFINITY, lo_post = p.act_post == 1 ? 0.0f : -INFINITY;
    // bias of every N tile -> shared memory (broadcast reads in the unit loop instead of exposed global latency)
    float* sbias = reinterpret_cast<float*>(gen_base + (bias0 - smem_base));
    for (int i = et; i < p.n_tiles * p.block_n; i += n_epi) sbias[i] = i < p.Cout ? __ldg(p.bias + i) : 0.0f;
    asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory");
    pdl_wait();  // residual loads and output stores touch activations
    // residual tile of unit (tile, ci): 32 rows x 64 channels by TMA into this warp's buffer
    auto issue_res = [&](int tile, int ci) {
      int nt, tw, th, tn;
      tile_coords(p, tile, nt, tw, th, tn);
      if (lane == 0) {
        mbar_expect_tx(rbar, EPI_TILE_BYTES);
        tma_load_4d(rsb, &maps.r, rbar, nt * p.block_n + ci * OUT_CHUNK, tw * p.bw, th * p.bh + qh, tn * p.bn + qn);
      }
    };
    if (has_res) {  // first unit of this warp
      int t0 = blockIdx.x, c = ustep == 1 ? 0 : half;
      while (c >= nch) { c -= nch; t0 += gridDim.x; }
      if (t0 < p.total_tiles) issue_res(t0, c);
    }
    uint32_t ti = 0, u = 0, rcount = 0;  // tiles done, units before this tile, residual tiles consumed
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti, u += nch) {
      const uint32_t a = ti & static_cast<uint32_t>(p.nacc - 1), aph = (ti >> p.nacc_sh) & 1u;
      int nt, tw, th, tn;
      tile_coords(p, tile, nt, tw, th, tn);
      const int n_base = nt * p.block_n;
      const int first = (ustep == 1 || (u & 1u) == static_cast<uint32_t>(half)) ? 0 : 1;  // my first chunk of this tile
      if (first >= nch) continue;  // no unit of this tile is mine (the MMA warp does not wait for my group then)
      const int last = first + (nch - 1 - first) / ustep * ustep;                          // my last one
      const uint32_t tempty = tempty0 + 16u * a + 8u * static_cast<uint32_t>(half);
      mbar_wait(tfull0 + 8u * a, aph);
      tc_fence_after();
      if (et == 0) trace_ev(p, 2, tr_i, tile, 0);
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * static_cast<uint32_t>(p.block_n);
      if (p.y32) {
        // fp32 output (logits / gate maps, Cout <= 16): one 16-column read, direct stores
        uint32_t acc[16];
        tmem_ld16(trow, acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty); else mbar_arrive(tempty); }
        const int ow = tw * p.bw + wl, oh = th * p.bh + hl, on = tn * p.bn + nl;
        if ((ow < p.Wo) && (oh < p.Ho) && (on < p.N)) {
          const size_t ypix = (static_cast<size_t>(on) * p.y_H + (oh * p.out_scale + p.out_oy)) * p.y_W +
                              (ow * p.out_scale + p.out_ox);
          float* yrow = p.y32 + ypix * p.y_ctot + p.y_c0;
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c < p.Cout) yrow[c] = fmaxf(fmaxf(__uint_as_float(acc[c]) + sbias[c], lo_pre), lo_post);
        }
        continue;
      }
      for (int ci = first; ci < nch; ci += ustep) {
        const int c0 = ci * OUT_CHUNK;
        const int cw = min(OUT_CHUNK, p.block_n - c0);  // 16 / 32 / 48 / 64 columns
        uint32_t acc[64];
        tmem_ld16(trow + c0, acc);
        if (cw > 16) tmem_ld16(trow + c0 + 16, acc + 16);
        if (cw > 32) tmem_ld16(trow + c0 + 32, acc + 32);
        if (cw > 48) tmem_ld16(trow + c0 + 48, acc + 48);
        // bias of the first 32 columns while the TMEM loads are in flight (broadcast reads)
        const float4* bch = reinterpret_cast<const float4*>(sbias + n_base + c0);
        float4 bb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) bb[k] = bch[k];
        tmem_ld_wait();
        if (et == 0 && (KSPEC != 4)) trace_ev(p, 3, tr_j, tile, 2);
        if (ci == last) {  // the accumulator now lives in registers: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty); else mbar_arrive(tempty); }
        }
        // bias + activation in place (columns beyond cw hold garbage that the output map clips)
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = 32 * hlf + 4 * k;
            acc[c + 0] = __float_as_uint(fmaxf(__uint_as_float(acc[c + 0]) + bb[k].x, lo_pre));
            acc[c + 1] = __float_as_uint(fmaxf(__uint_as_float(acc[c + 1]) + bb[k].y, lo_pre));
            acc[c + 2] = __float_as_uint(fmaxf(__uint_as_float(acc[c + 2]) + bb[k].z, lo_pre));
            acc[c + 3] = __float_as_uint(fmaxf(__uint_as_float(acc[c + 3]) + bb[k].w, lo_pre));
          }
          if (hlf == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) bb[k] = bch[8 + k];
          }
        }
        if (has_res) {  // same swizzled position in the residual tile as in the staging tile
          mbar_wait(rbar, rcount & 1u);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint4 rv = *reinterpret_cast<const uint4*>(rsrow + ((g ^ swz) * 16));
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rw[j]));
              acc[8 * g + 2 * j] = __float_as_uint(__uint_as_float(acc[8 * g + 2 * j]) + rf.x);
              acc[8 * g + 2 * j + 1] = __float_as_uint(__uint_as_float(acc[8 * g + 2 * j + 1]) + rf.y);
            }
          }
          // every lane has read the residual tile: fetch the one of my next unit
          __syncwarp();
          ++rcount;
          int t2 = tile, c2 = ci + ustep;
          while (c2 >= nch) { c2 -= nch; t2 += gridDim.x; }
          if (t2 < p.total_tiles) issue_res(t2, c2);
        }
        if (lane == 0) tma_store_wait_read<0>();  // my previous TMA store has finished reading the staging tile
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint4 o;
          o.x = pack2_sat(fmaxf(__uint_as_float(acc[8 * g + 0]), lo_post), fmaxf(__uint_as_float(acc[8 * g + 1]), lo_post));
          o.y = pack2_sat(fmaxf(__uint_as_float(acc[8 * g + 2]), lo_post), fmaxf(__uint_as_float(acc[8 * g + 3]), lo_post));
          o.z = pack2_sat(fmaxf(__uint_as_float(acc[8 * g + 4]), lo_post), fmaxf(__uint_as_float(acc[8 * g + 5]), lo_post));
          o.w = pack2_sat(fmaxf(__uint_as_float(acc[8 * g + 6]), lo_post), fmaxf(__uint_as_float(acc[8 * g + 7]), lo_post));
          *reinterpret_cast<uint4*>(srow + ((g ^ swz) * 16)) = o;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&maps.y, stg, n_base + c0, tw * p.bw, th * p.bh + qh, tn * p.bn + qn);
          tma_store_commit();
        }
        if (et == 0 && (KSPEC != 4)) trace_ev(p, 3, tr_j, tile, 5);
      }
      if (et == 0 && (KSPEC != 4)) trace_ev(p, 3, tr_j, tile, 0);
    }
    if (lane == 0) tma_store_wait_all();
  }
  if (p.dbg & 4) __nanosleep(5000);
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no remote arrive / multicast may target a CTA that has left
  if (warp == 1) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
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

// fp16 tensor map, 128-byte swizzle, zero OOB fill.  dims/strides innermost first; strides[i] is the
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
  Maps maps;
  Params p;
  dim3 grid;
  int smem_bytes;
};

inline int floor_pow2(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}

// x: input view (fp16), y: output view (fp16, or fp32 with at most 16 channels).  w_dev: [ntaps][Cout][Cin] fp16.
// BD_HALO_STREAM: 0 = off, 1 = N tiles of at most 128 columns, 2 = every width (default: measured faster on every
// 3x3 layer of the five networks, profiles/r2r_op_table_*)
inline bool halo_stream_mode(int block_n) {
  static const int mode = [] { const char* e = getenv("BD_HALO_STREAM"); return e ? atoi(e) : 2; }();
  return mode >= 2 || (mode == 1 && block_n <= 128);
}
inline int prepare(Launch* L, const TView& x, const TView& y, const TView* res, int ntaps, const int* dy,
                   const int* dx, int stride, int Ho, int Wo, int act_pre, int act_post, int out_scale, int out_oy,
                   int out_ox, const h16* w_dev, const float* bias_dev, int smem_budget_kb, int max_block_n,
                   int num_sms, int group_hint = 0, const h16* dw_w_dev = nullptr, int dw_relu = 0, int cin_valid = -1) {
  // cin_valid < Cin: only the first cin_valid channels of the input slice hold data (the rest of a buffer that shares
  // its arena range is garbage): the A tensor maps end there and the TMA zero-fills the tail of the last k-block
  const int Cin = x.c, Cout = y.c;
  BD_CHECK(!x.f32, "umma conv needs an fp16 input map");
  BD_CHECK(Cin % 8 == 0 && x.c0 % 8 == 0 && x.ctot % 8 == 0, "umma conv needs 16-byte aligned input channel slices");
  if (y.f32) BD_CHECK(Cout <= 16 && res == nullptr && out_scale == 1, "fp32 umma output: at most 16 channels, no residual");
  else BD_CHECK(Cout % 8 == 0 && y.c0 % 8 == 0 && y.ctot % 8 == 0, "umma conv needs 16-byte aligned output channel slices");
  BD_CHECK(stride == 1 || stride == 2, "umma conv stride must be 1 or 2");
  BD_CHECK(ntaps >= 1 && ntaps <= MAX_TAPS, "bad tap count");
  Params& p = L->p;
  memset(&p, 0, sizeof(p));
  p.N = x.N; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.Cin = Cin;
  p.bw = std::min(16, floor_pow2(Wo));
  p.bh = std::min(BLOCK_M / p.bw, floor_pow2(Ho));
  p.bn = BLOCK_M / (p.bw * p.bh);
  // halo path candidate: plain 3x3, stride 1, dilation 1 ('same' padding), map at least 8 x 16
  // (18 taps: a hi/lo weight split, the 3x3 offsets twice; 2..8 taps: any subset of the 3x3 window, e.g. the
  // sub-pixel phases of up-sample+conv heads and 3x3 transposed convolutions)
  bool full3x3 = (ntaps == 9 || ntaps == 18);
  for (int t = 0; t < ntaps && full3x3; ++t) full3x3 = dy[t] == (t % 9) / 3 - 1 && dx[t] == (t % 9) % 3 - 1;
  static const bool subset_on = [] { const char* e = getenv("BD_HALO_SUBSET"); return !(e && e[0] == '0'); }();
  bool subset = subset_on && !full3x3 && ntaps >= 2 && ntaps <= 9;
  for (int t = 0; t < ntaps && subset; ++t) subset = dy[t] >= -1 && dy[t] <= 1 && dx[t] >= -1 && dx[t] <= 1;
  const bool halo = group_hint == 0 && (full3x3 || subset) && stride == 1 && (out_scale == 1 || subset) && Wo >= 8 &&
                    Ho >= 16 && Wo == x.W && Ho == x.H;
  p.halo_subset = subset ? 1 : 0;
  p.tiles_w = cdiv(Wo, p.bw); p.tiles_h = cdiv(Ho, p.bh); p.tiles_n = cdiv(x.N, p.bn);
  const int cout16 = cdiv(Cout, 16) * 16;
  if (cout16 <= max_block_n) {
    p.block_n = cout16;  // a single N tile; staging chunks that overhang Cout are clipped by the output map
  } else {
    // several N tiles: multiples of 64 so that no staging chunk spills into the next tile's channels.  (A width chosen
    // per layer to fill whole waves -- 4 x 192 instead of 3 x 256 for 768->728 @32^2 at batch 32 -- measured 2 % SLOWER:
    // the per-tile costs outweigh the saved tail, profiles/README.md round 2.)
    const int cap = std::max(64, max_block_n / 64 * 64);
    const int ntile = cdiv(cout16, cap);
    p.block_n = std::min(cap, cdiv(cdiv(cout16, ntile), 64) * 64);
  }
  p.n_tiles = cdiv(Cout, p.block_n);
  p.kchunks = cdiv(Cin, BLOCK_K);
  p.ntaps = ntaps;
  {
    // Four accumulator stages (BD_UMMA_NACC=4) measured no different from two on every layer (profiles/r2v_*): the
    // small-N tiles are bound by the ~70 clk a tcgen05.mma costs whatever its N (18 MMAs = 1270 clk per 32->32 tile,
    // two issuers alternating), not by the accumulator hand-off.  Two stays the default.
    static const int env_nacc = [] { const char* e = getenv("BD_UMMA_NACC"); return e ? atoi(e) : 2; }();
    p.nacc = (env_nacc == 4 && 4 * p.block_n <= 512) ? 4 : 2;
    p.nacc_sh = p.nacc == 4 ? 2 : 1;
  }
  p.tmem_cols = 32;
  while (p.tmem_cols < p.nacc * p.block_n) p.tmem_cols *= 2;
  const int sub_bytes = A_STAGE_BYTES + p.block_n * 128;
  p.bias_bytes = (p.n_tiles * p.block_n * 4 + 127) / 128 * 128;
  // per-warp staging (+ residual) tiles, staged bias, alignment slack, barriers
  const int fixed = (res ? 2 : 1) * EPI_WARPS * EPI_TILE_BYTES + p.bias_bytes + 1024 + 1024;
  const int avail = smem_budget_kb * 1024 - fixed;
  // Small-N layers are bound by the fixed cost of a barrier round in the single-thread producer / MMA loops, so
  // several k-blocks share one stage (one wait + one expect_tx + one commit per `group` k-blocks) whenever the
  // k-block count divides evenly and at least two such stages fit.
  const int num_kb = ntaps * p.kchunks;
  p.group = 1;
  for (int g : {4, 3, 2})
    if (group_hint != 1 && num_kb % g == 0 && 2 * g * sub_bytes <= avail && p.block_n <= 128) { p.group = g; break; }
  if (group_hint > 1 && num_kb % group_hint == 0 && 2 * group_hint * sub_bytes <= avail) p.group = group_hint;
  p.spec = 0;
  p.wres_bytes = 0;
  p.aux_bytes = 0;
  p.dw_w = dw_w_dev; p.dw_relu = dw_relu;
  const int wbytes = ntaps * p.kchunks * p.block_n * 128;
  if (dw_w_dev) {
    // fused separable convolution: x is the DEPTHWISE input; per k-block a halo box + the pointwise weight tile
    BD_CHECK(ntaps == 1 && stride == 1 && out_scale == 1 && !y.f32 && Cin % BLOCK_K == 0 && Wo >= 8 && Ho >= 16 &&
                 Wo == x.W && Ho == x.H && dy[0] == 0 && dx[0] == 0,
             "fused separable conv: needs a stride-1 1x1 pointwise stage, Cin % 64 == 0 and a map of at least 8 x 16");
    p.spec = 4; p.group = 1;
    const int fixed4 = fixed - (res ? 2 : 1) * (EPI_WARPS / 2) * EPI_TILE_BYTES;  // four epilogue warps only
    const int avail4 = smem_budget_kb * 1024 - fixed4;
    p.wres_bytes = 2 * A_STAGE_BYTES;                       // two A-tile slots
    p.aux_bytes = (9 * Cin * 2 + 127) / 128 * 128;          // staged depthwise weights
    p.bw = 8; p.bh = 16; p.bn = 1;
    p.tiles_w = cdiv(Wo, p.bw); p.tiles_h = cdiv(Ho, p.bh); p.tiles_n = x.N;
    const int stage4 = HALO_STAGE + p.block_n * 128;
    p.stages = std::min(6, (avail4 - p.wres_bytes - p.aux_bytes) / stage4);
    BD_CHECK(p.stages >= 2, "fused separable conv: shared memory budget too small");
    L->smem_bytes = p.stages * stage4 + p.wres_bytes + p.aux_bytes + fixed4;
  } else
  if (halo && p.n_tiles == 1 && wbytes + 2 * HALO_STAGE <= avail && wbytes <= 148 * 1024) {
    // halo path: weights resident, the ring holds one halo box per 64-channel chunk
    p.spec = 3; p.group = 1; p.wres_bytes = wbytes;
    p.bw = 8; p.bh = 16; p.bn = 1;
    p.tiles_w = cdiv(Wo, p.bw); p.tiles_h = cdiv(Ho, p.bh); p.tiles_n = x.N;
    {  // CTA pairs (BD_UMMA_PAIR_RES=0: off): each CTA keeps half of the columns of the resident weights.  Measured at
       // batch 32: 64->64 @512^2 915 -> 1188, 128->64 @512^2 1012 -> 1219, 32->32 @256^2 366 -> 388 TFLOP/s
      static const int env_pair3 = [] { const char* e = getenv("BD_UMMA_PAIR_RES"); return e ? atoi(e) : 1; }();
      if (env_pair3 && full3x3 && ntaps == 9 && !y.f32 && p.block_n >= 32 && p.block_n % 32 == 0 &&
          (p.tiles_w * p.tiles_h * p.tiles_n) % 2 == 0 && num_sms >= 2) {
        p.spec = 8; p.pair = 1; p.wres_bytes = wbytes / 2;
      }
    }
    p.stages = std::max(2, std::min(12, (avail - p.wres_bytes) / HALO_STAGE));
    L->smem_bytes = p.stages * HALO_STAGE + p.wres_bytes + fixed;
  } else
  if (halo && full3x3 && ntaps == 9 && halo_stream_mode(p.block_n) &&
      2 * HALO_STAGE + 2 * (p.block_n <= 128 ? 3 : 1) * p.block_n * 128 <= avail) {
    // halo path with streamed weights: the weights do not fit, but the activations still arrive as ONE halo box per
    // 64-channel chunk (two slots) instead of nine shifted boxes; the ring holds weight tiles only, `group` taps per
    // stage (three 16 KB tiles for N <= 128, where a single tile's four MMAs are shorter than a barrier round trip)
    p.spec = 5; p.group = p.block_n <= 128 ? 3 : 1;
    p.wres_bytes = 2 * HALO_STAGE;
    p.bw = 8; p.bh = 16; p.bn = 1;
    p.tiles_w = cdiv(Wo, p.bw); p.tiles_h = cdiv(Ho, p.bh); p.tiles_n = x.N;
    {  // CTA pairs (BD_UMMA_PAIR_HALO=0: off): each CTA its own halo box and half of every weight tile
      static const int env_pair5 = [] { const char* e = getenv("BD_UMMA_PAIR_HALO"); return e ? atoi(e) : 1; }();
      if (env_pair5 && !y.f32 && p.block_n >= 128 && p.block_n % 32 == 0 && (p.tiles_w * p.tiles_h * p.tiles_n) % 2 == 0 &&
          num_sms >= 2) {
        p.spec = 7; p.pair = 1;
      }
    }
    const int stage5 = p.group * (p.pair ? p.block_n / 2 : p.block_n) * 128;
    p.stages = std::max(2, std::min(12, (avail - p.wres_bytes) / stage5));
    L->smem_bytes = p.stages * stage5 + p.wres_bytes + fixed;
  } else
  if (group_hint == 0 && ntaps == 9 && p.kchunks == 1 && 2 * 3 * sub_bytes <= avail) { p.spec = 1; p.group = 3; }
  else if (group_hint == 0 && ntaps == 9 && p.kchunks == 2 && 2 * 2 * sub_bytes <= avail) { p.spec = 2; p.group = 2; }
  {
    // CTA pairs (BD_UMMA_PAIR=1) for the generic ring: 1x1 / dilated layers with N tiles of at least 128 columns and an
    // even number of pixel tiles -- M = 256 MMAs issued by the even CTA of a 2-CTA cluster, each CTA loads its own pixel
    // tile and half of the weight tile (the generic ring is bound by L2->SM traffic: 48 -> 32 KB per k-block)
    // Measured per layer (profiles/r2z_op_table_b32_pair.txt): +8 ... +20 % on the compute-heavy layers (K = taps x Cin
    // >= 512: 768->728 684 -> 750, dilated 2048->256 1173 -> 1418, 1536->2048 1274 -> 1459 TFLOP/s), -5 ... -10 % on
    // the HBM-bound short-K 1x1 layers (128->128 @256^2), which therefore stay single.  BD_UMMA_PAIR=0 turns it off.
    static const int env_pair = [] { const char* e = getenv("BD_UMMA_PAIR"); return e ? atoi(e) : 1; }();
    const int m_total = p.tiles_w * p.tiles_h * p.tiles_n;
    if (env_pair && p.spec == 0 && !dw_w_dev && stride == 1 && !y.f32 && p.block_n >= 128 && p.block_n % 32 == 0 &&
        m_total % 2 == 0 && num_sms >= 2 && group_hint == 0 && ntaps * Cin >= 512) {
      p.spec = 6; p.pair = 1;
    }
  }
  if (p.spec != 3 && p.spec != 4 && p.spec != 5 && p.spec != 7 && p.spec != 8) {
    const int stage_bytes = (p.pair ? A_STAGE_BYTES + p.block_n / 2 * 128 : sub_bytes) * p.group;
    p.stages = std::max(2, std::min(12, avail / stage_bytes));
    // the ring covers two tiles, and at least ~190 KB / 8 stages of loads in flight: the short-K 1x1 layers (the K = 32
    // stems: one 24 KB stage per tile) are bound by load latency, not by shared memory
    p.stages = std::min(p.stages, std::max({2, 2 * num_kb / p.group, std::min(8, 192 * 1024 / stage_bytes)}));
    L->smem_bytes = p.stages * stage_bytes + fixed;
  }
  BD_CHECK(L->smem_bytes <= 227 * 1024, "umma conv smem budget exceeded");
  {
    // Two MMA issuers when a whole tile fits into the shared-memory ring (BD_UMMA_ISSUERS=1 turns it off for A/B
    // measurements; the fused separable path has one).  An issuer that has finished its tile steps over the other
    // issuer's ring stages and waits on a full barrier that far ahead; mbarrier parity waits are only sound at most
    // one phase ahead of the barrier: the stage `stages` before the awaited one must already have landed, which is
    // guaranteed (it belongs to the tile this issuer has just consumed) iff a tile spans fewer than `stages` stages.
    // These are exactly the short-K tiles that lose the most to per-tile bookkeeping.
    static const int env_dbg = [] { const char* e = getenv("BD_UMMA_DBG"); return e ? atoi(e) : 0; }();
    p.dbg = env_dbg;
    static const int env_pf = [] { const char* e = getenv("BD_UMMA_PREFETCH"); return e ? atoi(e) : 1; }();
    p.prefetch = env_pf;
    // ON by default since round 2: the intermittent failures of round 1 came from odd ring sizes (a stage shared by the
    // two issuers -> mbarrier parity aliasing, see the issuer loop); with the even-ring rule below the scheme ran
    // tools/stress2.py clean at batch 16 and 32 (result digests + the per-op timing path).  BD_UMMA_ISSUERS=1: one.
    static const int env_issuers = [] { const char* e = getenv("BD_UMMA_ISSUERS"); return e ? atoi(e) : 2; }();
    const bool halo3 = p.spec == 3 || p.spec == 8;
    const int stages_per_tile = halo3 ? p.kchunks : num_kb / p.group;
    // Only on the halo path: there one elected lane issues a whole tile (36 MMAs + both commits) in one go.  On the
    // grouped generic ring (several elect blocks per stage) two issuers showed an intermittent hang on B200 that
    // is not understood yet, so those layers keep the single issuer.
    // ... and there only for the configuration that has been stress-tested: the full 3x3 with one 64-channel chunk
    // per tile (the 64->64 and 32->32 layers, which are the ones that gain).  With the tap-subset variant (two
    // chunks per tile) the two-issuer scheme faulted under tools/stress.py; single-issuer it is clean.
    p.issuers = (!halo3 || p.halo_subset || p.kchunks != 1 || ntaps != 9 || env_issuers != 2 ||
                 stages_per_tile >= p.stages) ? 1 : 2;
    if (p.issuers == 2 && (p.stages & 1)) {
      // one stage per tile and an even ring: each issuer owns every other stage (see the issuer loop)
      if (p.stages >= 3) { p.stages -= 1; L->smem_bytes -= HALO_STAGE; }
      else p.issuers = 1;
    }
  }

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
    uint64_t dims[4] = {static_cast<uint64_t>(cin_valid >= 0 && cin_valid < Cin ? std::max(cin_valid, 8) : Cin),
                        static_cast<uint64_t>((x.W - px + stride - 1) / stride),
                        static_cast<uint64_t>((x.H - py + stride - 1) / stride), static_cast<uint64_t>(x.N)};
    uint64_t strides[3] = {pitch * stride, pitch * x.W * stride, pitch * x.W * x.H};
    uint32_t box[4] = {BLOCK_K, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), static_cast<uint32_t>(p.bn)};
    if (p.spec == 3 || p.spec == 4 || p.spec == 5 || p.spec == 7 || p.spec == 8) { box[1] = HALO_W; box[2] = HALO_H; box[3] = 1; }
    char* base = static_cast<char*>(x.base) + (static_cast<size_t>(py) * x.W + px) * pitch + static_cast<size_t>(x.c0) * 2;
    if (encode_h16(&L->maps.a[m], base, 4, dims, strides, box)) return 1;
    if (first < 0) first = m;
  }
  for (int m = 0; m < 4; ++m)
    if (!used[m]) L->maps.a[m] = L->maps.a[first];
  {
    uint64_t dims[3] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Cout), static_cast<uint64_t>(ntaps)};
    uint64_t strides[2] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Cin) * Cout * 2};
    uint32_t box[3] = {BLOCK_K, static_cast<uint32_t>(p.pair ? p.block_n / 2 : p.block_n), 1};  // pair: half a tile per CTA
    if (encode_h16(&L->maps.b, const_cast<h16*>(w_dev), 3, dims, strides, box)) return 1;
  }
  // an epilogue warp stores (and fetches residuals for) its TMEM lane quadrant: 32 consecutive tile rows =
  // a (bw x qbh x qbn) sub-box of the (bw x bh x bn) tile box
  const int qbh = std::min(p.bh, 32 / p.bw), qbn = 32 / (p.bw * qbh);
  p.fd_nt = make_fastdiv(p.n_tiles); p.fd_tw = make_fastdiv(p.tiles_w); p.fd_th = make_fastdiv(p.tiles_h);
  p.m_total = p.tiles_w * p.tiles_h * p.tiles_n;
  p.fd_mt = make_fastdiv(p.m_total);
  {
    // BD_UMMA_TILE_ORDER: 0 = N tile fastest (CTAs that run together share pixel tiles), 1 = pixel tile fastest (they
    // share one weight tile)
    static const int env_order = [] { const char* e = getenv("BD_UMMA_TILE_ORDER"); return e ? atoi(e) : 0; }();
    p.m_fast = (env_order == 1 && p.n_tiles > 1) ? 1 : 0;
  }
  p.y_ctot = y.ctot; p.y_c0 = y.c0; p.y_H = y.H; p.y_W = y.W;
  p.out_scale = out_scale; p.out_oy = out_oy; p.out_ox = out_ox;
  if (y.f32) {
    p.y32 = static_cast<float*>(y.base);
    L->maps.y = L->maps.b;  // unused
  } else {
    // output map over the destination channel slice; the sub-pixel phases of a transposed convolution are a
    // strided view (every out_scale-th pixel starting at (out_oy, out_ox))
    const uint64_t ypitch = static_cast<uint64_t>(y.ctot) * 2;
    uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                        static_cast<uint64_t>(x.N)};
    uint64_t strides[3] = {ypitch * out_scale, ypitch * y.W * out_scale, ypitch * y.W * y.H};
    uint32_t box[4] = {OUT_CHUNK, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(qbh), static_cast<uint32_t>(qbn)};
    char* base = static_cast<char*>(y.base) + (static_cast<size_t>(out_oy) * y.W + out_ox) * ypitch + static_cast<size_t>(y.c0) * 2;
    if (encode_h16(&L->maps.y, base, 4, dims, strides, box)) return 1;
  }
  L->maps.r = L->maps.y;
  if (res) {
    BD_CHECK(!res->f32 && res->c0 % 8 == 0 && res->ctot % 8 == 0 && out_scale == 1, "bad residual view");
    p.res = static_cast<const h16*>(res->base); p.res_ctot = res->ctot; p.res_c0 = res->c0;
    const uint64_t rpitch = static_cast<uint64_t>(res->ctot) * 2;
    uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                        static_cast<uint64_t>(x.N)};
    uint64_t strides[3] = {rpitch, rpitch * res->W, rpitch * res->W * res->H};
    uint32_t box[4] = {OUT_CHUNK, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(qbh), static_cast<uint32_t>(qbn)};
    char* base = static_cast<char*>(res->base) + static_cast<size_t>(res->c0) * 2;
    if (encode_h16(&L->maps.r, base, 4, dims, strides, box)) return 1;
  }
  p.bias = bias_dev; p.act_pre = act_pre; p.act_post = act_post;
  p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  L->grid = dim3(static_cast<unsigned>(std::min(p.total_tiles, num_sms)));
  if (p.pair) L->grid = dim3(L->grid.x & ~1u);  // whole clusters of two (total_tiles is even)
  return 0;
}

inline int launch(const Launch& L, cudaStream_t stream, bool pdl = false) {
  // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per device of this process
  static bool attr_set_dev[64] = {};
  int dev_ = 0;
  BD_CUDA(cudaGetDevice(&dev_));
  bool& attr_set = attr_set_dev[dev_ & 63];
  if (!attr_set) {
    BD_CUDA(cudaFuncSetAttribute(conv_umma_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BD_CUDA(cudaFuncSetAttribute(conv_umma_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BD_CUDA(cudaFuncSetAttribute(conv_umma_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BD_CUDA(cudaFuncSetAttribute(conv_umma_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BD_CUDA(cudaFuncSetAttribute(conv_umma_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BD_CUDA(cudaFuncSetAttribute(conv_umma_kernel<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BD_CUDA(cudaFuncSetAttribute(conv_umma_kernel<6, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BD_CUDA(cudaFuncSetAttribute(conv_umma_kernel<7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BD_CUDA(cudaFuncSetAttribute(conv_umma_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  void (*kern)(Maps, Params) = conv_umma_kernel<0, 1>;
  if (L.p.spec == 3) kern = L.p.halo_subset ? conv_umma_kernel<3, 0> : L.p.ntaps == 18 ? conv_umma_kernel<3, 2> : conv_umma_kernel<3, 1>;
  else if (L.p.spec == 4) kern = conv_umma_kernel<4, 1>;
  else if (L.p.spec == 5) kern = conv_umma_kernel<5, 1>;
  if (L.p.spec == 6 || L.p.spec == 7 || L.p.spec == 8) {  // CTA pairs: clusters of two
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = L.grid; cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = static_cast<size_t>(L.smem_bytes); cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
    if (L.p.spec == 6) BD_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<6, 1>, L.maps, L.p));
    else if (L.p.spec == 7) BD_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<7, 1>, L.maps, L.p));
    else BD_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<8, 1>, L.maps, L.p));
    return 0;
  }
  BD_CUDA(launch_k(pdl, kern, L.grid, dim3(THREADS), static_cast<size_t>(L.smem_bytes), stream, L.maps, L.p));
  return 0;
}

}  // namespace umma
}  // namespace bd
