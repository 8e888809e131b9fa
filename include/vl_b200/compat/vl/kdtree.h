/* vl_b200/compat/vl/kdtree.h -- `#include "vl/kdtree.h"` (ImageProcess.h:37) resolved to the B200 shim. */
#ifndef VL_B200_COMPAT_KDTREE_H
#define VL_B200_COMPAT_KDTREE_H
#include "../../kdtree.h"
#endif
