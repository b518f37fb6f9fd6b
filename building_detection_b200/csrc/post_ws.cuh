// post_ws.cuh -- bd_ctx (shared by bd_api.cu / post.cu / contours.cu), the constants of the fusion and contour stages,
// and their device scratch: a bump arena for the scene-sized planes and grow-only slots for the arrays whose size
// depends on the number of runs / contours.  Nothing is allocated per call once warmed up.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace bd {
namespace post {

// The literals of model_fuse.py and edge_3.py in one place, with the reference line each one
// comes from.  bd_post_constants() exposes them; building_detection_b200/constants.py mirrors them on the host.
struct PostConstants {
  // model_fuse.py
  int fuse_min_area = 1000;      // fill_and_delete: polygons of area <= 1000 are erased            (:22)
  int fuse_min_fragment = 500;   // fill_small_target: fragments of area <= 500 are erased          (:57)
  int fuse_split_half = 10;      // 1x5 / 5x1 kernel, 5 iterations == one 1x21 / 21x1 erosion       (:180-181, :67)
  int fuse_votes = 3;            // sum of five masks >= 3                                            (:323)
  // edge_3.py
  int edge_min_area = 100;       // _detection: polygons of area <= 100 are erased                   (:326)
  int edge_min_fragment = 50;    // erode_images_process: fragments of area < 50 are erased          (:131)
  int edge_split_half = 3;       // 1x7 / 7x1 erosion, one iteration                                  (:128)
  double edge_iou = 0.5;         // process_td / process_rl: boxes match above IoU 0.5               (:42)
  double edge_min_moment = 10;   // contours with m00 <= 10 are skipped                              (:331)
  // area tiers of the polygon simplification (:351-378): epsilon = factor x perimeter
  double tier_small = 150, tier_mid = 300, tier_big0 = 3000, tier_big1 = 8000, tier_big2 = 15000;
  double eps_default = 0.01, eps_mid_mult = 5, eps_big0 = 0.005, eps_big1 = 0.004, eps_big2 = 0.002;
  double small_rate0 = 0.002, small_rate_step = 0.002;  // small_target retries (:265-286)
  int small_max_tries = 10;
};

struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  // make room for `bytes` and start carving from the beginning
  int reserve(size_t bytes) {
    if (bytes > cap) {
      if (base) cudaFree(base);
      base = nullptr; cap = 0;
      const size_t want = bytes + bytes / 8 + (1 << 20);
      BD_CUDA(cudaMalloc(reinterpret_cast<void**>(&base), want));
      cap = want;
    }
    off = 0;
    return 0;
  }
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;  // (reserve() was sized for the whole call; overflow is checked by the callers' byte budgets)
  }
  void release() {
    if (base) cudaFree(base);
    base = nullptr; cap = off = 0;
  }
};

// grow-only device scratch slots (no cudaMalloc / cudaFree per call once warmed up)
struct DevPool {
  static constexpr int SLOTS = 64;
  void* ptr[SLOTS] = {};
  size_t cap[SLOTS] = {};
  int get(int slot, size_t bytes, void** out) {
    if (bytes > cap[slot]) {
      if (ptr[slot]) cudaFree(ptr[slot]);
      ptr[slot] = nullptr; cap[slot] = 0;
      const size_t want = bytes + bytes / 4 + 256;
      BD_CUDA(cudaMalloc(&ptr[slot], want));
      cap[slot] = want;
    }
    *out = ptr[slot];
    return 0;
  }
  void release() {
    for (int i = 0; i < SLOTS; ++i) { if (ptr[i]) cudaFree(ptr[i]); ptr[i] = nullptr; cap[i] = 0; }
  }
};

// slot map: 0..31 contour tracing (three sets x 8, two for the box matching), 32.. fusion / labelling
enum {
  SLOT_IN5 = 32,  // five packed input planes of bd_fuse
  SLOT_BG = 33,   // +1: background parents, outside flags
  SLOT_F = 35,    // +1: object parents, areas
  SLOT_EH = 37,   // +1
  SLOT_EV = 39,   // +1
  SLOT_CNT = 41,  // fragment counters
  SLOT_TILEMASK = 42,  // argmax masks of one batch of tiles (bd_scene_run)
  SLOT_PROBS = 43,     // probabilities of one batch of tiles, summed P(building) (bd_scene_run_average)
  SLOT_PROBACC = 44,
};

}  // namespace post
}  // namespace bd

struct bd_ctx {
  int device = 0;
  int num_sms = 148;
  int64_t launches = 0;
  int umma_smem_kb = 226;   // per-CTA smem budget of the persistent tcgen05 conv (1 CTA / SM)
  int umma_max_block_n = 256;
  int umma_group = 0;       // k-blocks per smem stage: 0 = automatic, 1 = off (BD_UMMA_GROUP)
  int* d_ys = nullptr;      // tile origin scratch
  int* d_xs = nullptr;
  int tile_cap = 0;
  int* d_all_ys = nullptr;  // origins of all tiles of the current scene (bd_tiles_set_origins)
  int* d_all_xs = nullptr;
  int origin_cap = 0, n_origins = 0;
  bd::post::PostConstants consts;
  bd::post::Arena arena;
  bd::post::DevPool pool;
  int* h_scalar = nullptr;    // pinned host word for device -> host counters
  void* h_pts[3] = {nullptr, nullptr, nullptr};  // pinned host buffers of the three traced contour sets (grow-only)
  size_t h_pts_cap[3] = {0, 0, 0};
  cudaStream_t capture_stream = nullptr;  // CUDA-graph capture of the plans (bd_scene_run)
  void* trace_buf = nullptr;  // BD_UMMA_TRACE debug buffer of the most recently built conv
};
