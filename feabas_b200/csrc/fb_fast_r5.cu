// fast-path kernels for the "r5" group of line lengths (see fb_fast_groups.h)
#include "fb_fast_groups.h"
#define FB_TU_SIZES(X) FB_FAST_SIZES_R5(X)
#define FB_TU_G r5
#define FB_TU_EXTRA 0
#include "fb_fast_tu.inc"
