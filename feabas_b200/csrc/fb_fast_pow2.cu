// fast-path kernels for the "pow2" group of line lengths (see fb_fast_groups.h)
#include "fb_fast_groups.h"
#define FB_TU_SIZES(X) FB_FAST_SIZES_POW2(X)
#define FB_TU_G pow2
#define FB_TU_EXTRA 1
#include "fb_fast_tu.inc"
