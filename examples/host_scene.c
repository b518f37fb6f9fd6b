/* host_scene.c -- the whole hot path from a plain C host through the C ABI (include/bd_b200.h), no Python at run time:
 *
 *     bd_create -> bd_plan_load x N (plan files written by tools/export_plans.py) -> upload a BGR scene ->
 *     bd_scene_run (tiling + N network forwards + OR-stitch, predict.py:75-116) -> bd_fuse (model_fuse.py:271-350) ->
 *     bd_contours (edge_3.py:310-387) -> fused mask + polygons
 *
 * This is what a binding in another language (cgo, JNI, N-API ...) wraps.  Build and run:
 *     gcc -O2 examples/host_scene.c -Iinclude -I/usr/local/cuda/include -Lbuilding_detection_b200 -lbd_b200 \
 *         -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/building_detection_b200 -o host_scene
 *     ./host_scene scene.bgr H W out_mask.u8 out_polys.txt plan0.bdplan ... plan4.bdplan
 * scene.bgr: H*W*3 bytes (what cv.imread returns); out_mask.u8: H*W bytes {0,255}; out_polys.txt: one polygon per line,
 * "x,y " per vertex (predict.py:119-132).  With fewer than five plans only the stitched masks are written. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bd_b200.h"

#define CK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, bd_last_error()); return 1; } } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s scene.bgr H W out_mask.u8 out_polys.txt plan.bdplan [...]\n", argv[0]); return 2; }
  const int h = atoi(argv[2]), w = atoi(argv[3]), n_plans = argc - 6;
  const size_t npx = (size_t)h * w;
  uint8_t* scene = (uint8_t*)malloc(npx * 3);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(scene, 1, npx * 3, f) != npx * 3) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
  fclose(f);

  bd_ctx* ctx = NULL;
  CK(bd_create(0, &ctx));
  bd_plan* plans[8];
  for (int k = 0; k < n_plans && k < 8; ++k) CK(bd_plan_load(ctx, argv[6 + k], &plans[k]));

  /* tile origins in the reference's order (predict.py:98-107; the column loop runs over new_h: square scenes) */
  const int h_num = (int)ceil((h - 152) / 360.0), new_h = h_num * 360 + 152;
  int n_tiles = 0;
  int32_t *ys = (int32_t*)malloc(sizeof(int32_t) * (size_t)(h_num + 1) * (h_num + 1)), *xs = (int32_t*)malloc(sizeof(int32_t) * (size_t)(h_num + 1) * (h_num + 1));
  for (int i = 0; i < new_h - 152; i += 360)
    for (int j = 0; j < new_h - 152; j += 360) { ys[n_tiles] = i; xs[n_tiles] = j; ++n_tiles; }

  uint8_t *d_scene, *d_masks, *d_fused;
  CU(cudaMalloc((void**)&d_scene, npx * 3));
  CU(cudaMalloc((void**)&d_masks, npx * n_plans));
  CU(cudaMalloc((void**)&d_fused, npx));
  CU(cudaMemcpy(d_scene, scene, npx * 3, cudaMemcpyHostToDevice));
  CU(cudaMemset(d_masks, 0, npx * n_plans));
  CK(bd_scene_run(ctx, plans, n_plans, d_scene, h, w, ys, xs, n_tiles, d_masks, NULL));
  CU(cudaDeviceSynchronize());

  uint8_t* out = (uint8_t*)malloc(npx * (n_plans == 5 ? 1 : n_plans));
  FILE* fp = fopen(argv[5], "w");
  if (n_plans == 5) {
    CK(bd_fuse(ctx, d_masks, h, w, d_fused, NULL));
    bd_polys polys;
    const int rc = bd_contours(ctx, d_fused, h, w, &polys, NULL);
    if (rc != 0) fprintf(fp, "ERROR %s\n", bd_last_error());
    else {
      for (int i = 0; i < polys.n_polys; ++i) {
        if (polys.is_float[i] == 2) { fprintf(fp, "RAW\n"); continue; } /* needs cv::minAreaRect on the host (edge_3.py:281-285) */
        for (int k = polys.offsets[i]; k < polys.offsets[i + 1]; ++k) fprintf(fp, "%d,%d ", (int)polys.xs[k], (int)polys.ys[k]);
        fprintf(fp, "\n");
      }
      bd_polys_free(&polys);
    }
    CU(cudaMemcpy(out, d_fused, npx, cudaMemcpyDeviceToHost));
  } else {
    CU(cudaMemcpy(out, d_masks, npx * n_plans, cudaMemcpyDeviceToHost));
  }
  fclose(fp);
  f = fopen(argv[4], "wb");
  fwrite(out, 1, npx * (n_plans == 5 ? 1 : n_plans), f);
  fclose(f);
  printf("host_scene: %d tiles, %d plans, %lld kernel launches\n", n_tiles, n_plans, (long long)bd_launch_count(ctx));
  for (int k = 0; k < n_plans; ++k) bd_plan_destroy(plans[k]);
  bd_destroy(ctx);
  return 0;
}
