// pbsc_status.h — internal per-walk / per-read status codes shared by the kernels and the host-side retry policy.
#ifndef PBSC_STATUS_H
#define PBSC_STATUS_H
#include <cuda_runtime.h>

namespace pbsc {
#define PBSC_WALK_OVERFLOW (-100)   // a fixed-size slot was too small (out slot, query scratch): the host re-runs with larger scratch
// distinct causes, so that pbsc_batch_run grows only the capacity that ran out (and reports a DP limit at once)
#define PBSC_OVF_PIECES (-110)      // a read's piece / bounds region (stitch_kernel): piece_factor
#define PBSC_OVF_TREE (-111)        // label tree, result list or history rings of the full-capacity pass: node_cap
#define PBSC_OVF_POOL (-112)        // shared label-tree pool of successful light walks: pool_factor
#define PBSC_OVF_DP (-113)          // DP fallback outside this build's limits: not a capacity, reported as PBSC_ERR_LIMIT
__host__ __device__ __forceinline__ bool is_overflow(int st) { return st == PBSC_WALK_OVERFLOW || (st <= PBSC_OVF_PIECES && st >= PBSC_OVF_DP); }
#define PBSC_WALK_UNSUPPORTED (-101)

}  // namespace pbsc
#endif
