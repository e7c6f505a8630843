// pbsc_task.cuh — one seed pair of a read as the extend phase schedules it (shared by the thread engine,
// pbsc_extend_thread.cu, and the DP / multiple-alignment fallback, pbsc_dp.cu).
#ifndef PBSC_TASK_CUH
#define PBSC_TASK_CUH

#include "pbsc_walk_thread.cuh"

namespace pbsc {

#define PBSC_TASK_PENDING (-999)

// dp_status of a task whose FM walk failed (correctByMSAlignment, PacBioSelfCorrectionProcess.cpp:208-245)
#define PBSC_DP_NONE 0        // fallback not run (--nodp, look-ahead request, or the walk succeeded)
#define PBSC_DP_OK 1          // consensus of out_len bases at out_off
#define PBSC_DP_FEW_ROWS (-1) // maquery.getNumRows() <= 3: the caller appends the raw read instead

struct __align__(16) WalkTask
{
    uint64_t src_hi, src_lo;   // last k bases of the source piece, newest base in the top two bits of src_hi
    uint64_t out_off;          // where the merged sequence goes in the output pool
    uint32_t read;
    int32_t src_end;           // source.seedEndPos in the raw read
    int32_t trg_start, trg_len;
    int32_t k, rtou;
    int32_t status;
    uint32_t out_len, out_cap, valid;
    int32_t freq_sum;          // source.maxFixedMerFreq + target.maxFixedMerFreq
    int32_t dp_wanted;         // run the DP fallback if the walk fails
    int32_t dp_status;
    uint32_t pad;
};
static_assert(sizeof(WalkTask) == 80, "WalkTask must be 80 bytes");

// base x of beginningkmer + strBetweenSrcTarget + targetSeed in read orientation (PacBioSelfCorrectionProcess.cpp:168-171,
// 220-224), before any isFromRtoU flip
__device__ __forceinline__ uint8_t task_query_base(const WalkTask& tk, const uint8_t* __restrict__ read, uint32_t x)
{
    const uint32_t k = (uint32_t)tk.k;
    const uint32_t interval = (uint32_t)(tk.trg_start - tk.src_end - 1);
    if (x < k) return (uint8_t)tail_base(tk.src_hi, tk.src_lo, (int)(k - 1 - x));
    if (x < k + interval) return read[tk.src_end + 1 + (x - k)];
    return read[tk.trg_start + (x - k - interval)];
}

}  // namespace pbsc
#endif
