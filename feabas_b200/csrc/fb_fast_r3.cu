// fast-path kernels for the "r3" group of line lengths (see fb_fast_groups.h)
#include "fb_fast_groups.h"
#define FB_TU_SIZES(X) FB_FAST_SIZES_R3(X)
#define FB_TU_G r3
#define FB_TU_EXTRA 0
#include "fb_fast_tu.inc"
