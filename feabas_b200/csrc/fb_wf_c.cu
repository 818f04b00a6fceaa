// warp-fused kernel, size group c (fb_wf_groups.h)
#include "fb_wf_groups.h"
#define FB_TU_SIZES(X) FB_WF_SIZES_C(X)
#define FB_TU_G c
#include "fb_wf_tu.inc"
