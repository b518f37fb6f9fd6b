// post.cu -- mask fusion (model_fuse.py) and contour extraction (edge_3.py) entry points.
#include "../../include/bd_b200.h"
#include "common.cuh"

using namespace bd;

extern "C" {
int bd_fuse(bd_ctx*, const uint8_t*, int, int, uint8_t*, void*) { return fail("bd_fuse: not implemented yet"); }
int bd_mask_cleanup(bd_ctx*, const uint8_t*, int, int, uint8_t*, void*) { return fail("bd_mask_cleanup: not implemented yet"); }
int bd_contours(bd_ctx*, const uint8_t*, int, int, bd_polys*, void*) { return fail("bd_contours: not implemented yet"); }
void bd_polys_free(bd_polys*) {}
}
