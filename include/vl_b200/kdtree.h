/* vl_b200/kdtree.h -- drop-in replacement for the part of VLFeat's vl/kdtree.h the stitcher binds
 * (ImageProcess::getImgPair, ImageProcess.cpp:273-351), backed by libpano_b200.so.
 *
 * The reference builds a 1-tree kd-forest over image A's descriptors (VL_TYPE_FLOAT, 128 dimensions, VlDistanceL1,
 * exact search: max_num_comparisons = 0) and asks it for the 2 nearest neighbours of every descriptor of image B.
 * In 128 dimensions the tree prunes nothing, so the GPU implementation is an exact brute-force scan
 * with VLFeat's distance arithmetic (float accumulator over the dimensions in order, vl/mathop.c:307-318 for L1,
 * :296-305 for squared L2).  vl_kdforest_build uploads the table to HBM once; every query entry point returns the
 * neighbours in ascending distance order with `distance` widened to double exactly as VLFeat does
 * (vl/kdtree.c:773-847).
 *
 *   reference call (vl/kdtree.h)                                replacement
 *   vl_kdforest_new (:137)                                      same signature; float data, L1 or L2, any dimension
 *   vl_kdforest_build (:146)                                    same (the data IS copied, to the device)
 *   vl_kdforest_new_searcher / vl_kdforestsearcher_delete       same
 *   vl_kdforestsearcher_query, vl_kdforest_query (:150, :162)   same, numNeighbors <= VL_B200_KDFOREST_MAX_NEIGHBORS;
 *                                                               one small kernel launch + one synchronisation per call
 *   vl_kdforest_query_with_array (:155)                         same; ALL queries in one launch -- the call to prefer
 *   vl_kdforest_delete (:140)                                   same
 *
 * Differences from VLFeat, all outside what the reference relies on: equidistant neighbours are returned in
 * ascending index order (VLFeat: tree visiting order); approximate search (max_num_comparisons > 0) is accepted and
 * answered exactly; VL_TYPE_DOUBLE data and the other VlVectorComparisonType values are refused (vl_kdforest_new
 * returns NULL and prints the reason).  The return value of the query functions is the number of data points compared
 * (VLFeat: number of leaves visited); the reference ignores it.
 */
#ifndef VL_B200_KDTREE_H
#define VL_B200_KDTREE_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef VL_B200_BASIC_TYPES
#define VL_B200_BASIC_TYPES
typedef unsigned long long vl_size;   /* vl/host.h:392 */
typedef unsigned long long vl_uindex; /* vl/host.h:394 */
typedef unsigned int vl_uint32;       /* vl/host.h:382 */
typedef vl_uint32 vl_type;            /* vl/generic.h:18 */
#define VL_TYPE_FLOAT 1               /* vl/generic.h:21 */
#define VL_TYPE_DOUBLE 2              /* vl/generic.h:22 */
#endif

#ifndef VL_B200_VECTOR_COMPARISON_TYPE
#define VL_B200_VECTOR_COMPARISON_TYPE
typedef enum _VlVectorComparisonType { /* vl/mathop.h:628-640 (values must match) */
    VlDistanceL1, VlDistanceL2, VlDistanceChi2, VlDistanceHellinger, VlDistanceJS, VlDistanceMahalanobis,
    VlKernelL1, VlKernelL2, VlKernelChi2, VlKernelHellinger, VlKernelJS
} VlVectorComparisonType;
#endif

typedef enum _VlKDTreeThresholdingMethod { VL_KDTREE_MEDIAN, VL_KDTREE_MEAN } VlKDTreeThresholdingMethod; /* :54-57 */

typedef struct _VlKDForestNeighbor { /* vl/kdtree.h:60-63 */
    double distance;
    vl_uindex index;
} VlKDForestNeighbor;

#define VL_B200_KDFOREST_MAX_NEIGHBORS 8

typedef struct _VlKDForest VlKDForest;                 /* opaque */
typedef struct _VlKDForestSearcher VlKDForestSearcher; /* opaque */

VlKDForest* vl_kdforest_new(vl_type dataType, vl_size dimension, vl_size numTrees, VlVectorComparisonType normType);
VlKDForestSearcher* vl_kdforest_new_searcher(VlKDForest* kdforest);
void vl_kdforest_delete(VlKDForest* self);
void vl_kdforestsearcher_delete(VlKDForestSearcher* searcher);
void vl_kdforest_build(VlKDForest* self, vl_size numData, void const* data);
vl_size vl_kdforest_query(VlKDForest* self, VlKDForestNeighbor* neighbors, vl_size numNeighbors, void const* query);
vl_size vl_kdforest_query_with_array(VlKDForest* self, vl_uint32* index, vl_size numNeighbors, vl_size numQueries,
                                     void* distance, void const* queries);
vl_size vl_kdforestsearcher_query(VlKDForestSearcher* self, VlKDForestNeighbor* neighbors, vl_size numNeighbors,
                                  void const* query);

vl_size vl_kdforest_get_num_trees(VlKDForest const* self);
vl_size vl_kdforest_get_data_dimension(VlKDForest const* self);
vl_type vl_kdforest_get_data_type(VlKDForest const* self);
void vl_kdforest_set_max_num_comparisons(VlKDForest* self, vl_size n);
vl_size vl_kdforest_get_max_num_comparisons(VlKDForest* self);
void vl_kdforest_set_thresholding_method(VlKDForest* self, VlKDTreeThresholdingMethod method);
VlKDTreeThresholdingMethod vl_kdforest_get_thresholding_method(VlKDForest const* self);
VlKDForest* vl_kdforest_searcher_get_forest(VlKDForestSearcher const* self);

/* B200 extension: device used by subsequent vl_kdforest_new calls (default 0) */
void vl_b200_kdforest_set_device(int device);

#ifdef __cplusplus
}
#endif
#endif
