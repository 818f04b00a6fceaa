// warp-fused kernel, size group b (fb_wf_groups.h)
#include "fb_wf_groups.h"
#define FB_TU_SIZES(X) FB_WF_SIZES_B(X)
#define FB_TU_G b
#include "fb_wf_tu.inc"
