// placeholder, replaced below
#include "te_oracle.h"
extern "C" int teo_footprint(const teo_geometry*, const teo_footprint_params*, const float*, const float*, const float*, const float*, float*, float*, float*, int) { return 99; }
extern "C" int teo_spiral_offsets(double, double, int32_t*, int32_t*, int) { return -1; }
