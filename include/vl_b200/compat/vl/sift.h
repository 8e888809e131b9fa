/* vl_b200/compat/vl/sift.h -- `#include "vl/sift.h"` (ImageProcess.h:36) resolved to the B200 shim. */
#ifndef VL_B200_COMPAT_SIFT_H
#define VL_B200_COMPAT_SIFT_H
#include "../../sift.h"
#endif
