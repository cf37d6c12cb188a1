// Error reporting shared by every translation unit of libb200seg.so.
// Entry points return 0 or a negative B200_E* code; the text of the last failure on the calling
// thread is available through b200_last_error().
#pragma once
#include <cuda_runtime.h>

#define B200_OK 0
#define B200_EINVAL (-1)   /* bad shape / alignment / argument */
#define B200_ECUDA (-2)    /* CUDA runtime error (launch, attribute, memcpy) */
#define B200_EDRIVER (-3)  /* driver entry point / tensor-map encoding failure */
#define B200_EARCH (-4)    /* device is not compute capability 10.x */

namespace b200 {
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
}  // namespace b200
