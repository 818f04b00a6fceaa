// fast-path kernels for the "r5b" group of line lengths (60 / 50 points per lane; see fb_fast_groups.h)
#include "fb_fast_groups.h"
#define FB_TU_SIZES(X) FB_FAST_SIZES_R5B(X)
#define FB_TU_G r5b
#define FB_TU_EXTRA 0
#include "fb_fast_tu.inc"
