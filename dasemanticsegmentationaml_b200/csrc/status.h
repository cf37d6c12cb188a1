// Error reporting shared by every translation unit of libb200seg.so.
// Entry points return 0 or a negative B200_E* code; the text of the last failure on the calling
// thread is available through b200_last_error().
#pragma once
#include <cuda_runtime.h>

#include "b200seg.h"







namespace b200 {
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
}  // namespace b200
