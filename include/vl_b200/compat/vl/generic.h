/* vl_b200/compat/vl/generic.h -- stands in for VLFeat's vl/generic.h for sources that keep their
 * `#include "vl/generic.h"` (ImageProcess.h:35).  Put include/vl_b200/compat on the include path INSTEAD of the VLFeat
 * tree and link libpano_b200.so instead of libvl / vl.dll: the error codes, basic types and type ids the stitcher uses
 * (vl/generic.h:18-22, 108-113; vl/host.h:382-394) come from the two shim headers. */
#ifndef VL_B200_COMPAT_GENERIC_H
#define VL_B200_COMPAT_GENERIC_H
/* the standard headers vl/generic.h:7-10 and vl/sift.h:4 pull in, which callers rely on transitively
 * (ImageProcess.cpp:643 uses assert without including <cassert>) */
#include <stdlib.h>
#include <stddef.h>
#include <time.h>
#include <assert.h>
#include <stdio.h>
#include <math.h>
#include <float.h>
#include <limits.h>
#include "../../sift.h"
#include "../../kdtree.h"
#endif
