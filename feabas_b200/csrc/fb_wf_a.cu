// warp-fused kernel, size group a (fb_wf_groups.h)
#include "fb_wf_groups.h"
#define FB_TU_SIZES(X) FB_WF_SIZES_A(X)
#define FB_TU_G a
#include "fb_wf_tu.inc"
