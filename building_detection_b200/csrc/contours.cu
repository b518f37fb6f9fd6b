// contours.cu -- contour extraction (reference edge_3.py) entry points.
#include "../../include/bd_b200.h"
#include "common.cuh"
using namespace bd;
extern "C" {
int bd_contours(bd_ctx*, const uint8_t*, int, int, bd_polys*, void*) { return fail("bd_contours: not implemented yet"); }
void bd_polys_free(bd_polys*) {}
}
