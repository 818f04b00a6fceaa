// fast-path kernels for the "big" group of line lengths (see fb_fast_groups.h)
#include "fb_fast_groups.h"
#define FB_TU_SIZES(X) FB_FAST_SIZES_BIG(X)
#define FB_TU_G big
#define FB_TU_EXTRA 0
#include "fb_fast_tu.inc"
