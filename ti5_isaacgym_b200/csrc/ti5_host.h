// Host-side plumbing shared by the C-ABI translation units: argument checks, launch checks,
// the thread-local last-error string.
#pragma once
#include <cuda_runtime.h>

#include "../../include/ti5_step.h"

void ti5_set_error(const char* fmt, ...);
int ti5_check_launch(const char* what);

#define TI5_CHECK_ARGS(cond)                                                        \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ti5_set_error("%s: invalid argument: !(%s)", __func__, #cond);                \
      return TI5_EINVAL;                                                            \
    }                                                                               \
  } while (0)

