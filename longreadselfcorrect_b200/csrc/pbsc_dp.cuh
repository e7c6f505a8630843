// pbsc_dp.cuh — host entry of the DP / multiple-alignment fallback (pbsc_dp.cu).
#ifndef PBSC_DP_CUH
#define PBSC_DP_CUH

#include "pbsc_batch.cuh"

namespace pbsc {

struct DpStats { uint64_t jobs = 0, rows = 0, chunks = 0, bad = 0, thread_rows = 0; float ms = 0; };
DpStats& last_dp_stats();

// For every task of the list whose FM walk failed and that asked for it (dp_wanted): retrieve the overlapping reads, align
// them to the pair's query, build the multiple alignment and write the consensus to the task's slot of `outpool`
// (dp_status / out_len).  `tasks` is a device array of WalkTask; `list` (optional) selects n_items of them;
// q_cap bounds the query length of any task of this batch.
int run_dp_fallback(pbsc_index* idx, const pbsc_params* p, DeviceBatch& b, void* tasks, uint64_t n_items, const uint32_t* list, uint8_t* outpool,
                    uint32_t q_cap, uint64_t* launches);

}  // namespace pbsc
#endif
