// fb_common.h -- error reporting and launch accounting shared by the translation units of
// libfeabas_cuda.so (defined in fb_xcorr.cu).
#pragma once
#include <cstdarg>
#include <cstdio>

int fb_set_error(int code, const char* msg);      // stores the thread-local message, returns code
void fb_count_launches(int n);

static inline int fb_failf(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return fb_set_error(code, buf);
}

#define FB_CU(call)                                                                                                  \
    do {                                                                                                             \
        cudaError_t e_ = (call);                                                                                     \
        if (e_ != cudaSuccess)                                                                                       \
            return fb_failf(FB_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);        \
    } while (0)
