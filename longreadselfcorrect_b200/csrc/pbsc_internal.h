// pbsc_internal.h — host-side state shared by the translation units of libpbsc.so.
#ifndef PBSC_INTERNAL_H
#define PBSC_INTERNAL_H

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <exception>
#include <map>
#include <new>
#include <mutex>
#include <condition_variable>
#include <string>
#include <vector>
#include "../../include/pbsc.h"
#include "fm_table.cuh"
#include "pbsc_status.h"

struct pbsc_index
{
    int device = 0;
    pbsc::FmIndexDev dev;
    pbsc::FmBlock* d_blocks[2] = {nullptr, nullptr};
    uint32_t* d_dollar[2] = {nullptr, nullptr};
    uint64_t* d_dmask[2] = {nullptr, nullptr};
    pbsc::PrefixEntry* d_prefix = nullptr;
    uint8_t* d_idmer_valid = nullptr;
    uint64_t n_symbols[2] = {0, 0}, n_strings[2] = {0, 0}, n_blocks[2] = {0, 0};
    size_t device_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;   // side stream: the heavy walk pass runs here while the DP fallback of the light pass runs on `stream`
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    int sm_count = 0;   // cached: cudaGetDeviceProperties costs tens of milliseconds
    // capacities a batch had to grow to (label-tree nodes per walk, piece bytes per read base): later batches start there
    uint32_t learned_node_cap = 0;
    float learned_piece_factor = 0;
    uint32_t learned_pool_nodes = 0;
    // one batch RUNS at a time per index (the arena below and `stream` are shared); uploads and fetches of other batches
    // proceed on their own streams meanwhile (pbsc_pipeline.cu)
    std::mutex run_mu;
    // geometry of the source files (0 when the index was not loaded from .bwt/.rbwt): checked against PREFIX.fmg
    uint64_t src_runs[2] = {0, 0};
    void* blob = nullptr;        // single allocation that holds every table when the index came from a blob (import / .fmg / clone)
    size_t blob_bytes = 0;
    // Lanes: up to PBSC_MAX_LANES batches of one index run at the same time, each on its own stream with its own arena, so
    // that the tail of one batch's rounds (a few long walks, the long alignment pile-ups) overlaps the dense kernels of
    // another (replaces the worker threads of Concurrency/SequenceProcessFramework.h:91-230).  Lane 0 is the index itself;
    // the others are shadow objects that borrow its tables (`primary` set, nothing owned but stream + arena).
    pbsc_index* primary = nullptr;
    std::vector<pbsc_index*> shadows;
    std::vector<char> lane_busy;          // [0] = this object, [i] = shadows[i-1]
    std::condition_variable lane_cv;      // guarded by run_mu
    int lanes_active = 0;
    // named scratch buffers that survive across batches (grow-only), so that the hot path does not pay
    // cudaMalloc/cudaFree of gigabytes per batch; one batch runs at a time per index
    struct ArenaBuf { void* p = nullptr; size_t cap = 0; };
    std::map<std::string, ArenaBuf> arena;
};

namespace pbsc {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

// no C++ exception may cross the C ABI: entry points that allocate host memory are function-try-blocks closed by this
#define PBSC_CATCH_ALL(who)                                                                                               \
    catch (const std::bad_alloc&) { pbsc::set_error("%s: out of host memory", who); return PBSC_ERR_LIMIT; }              \
    catch (const std::exception& e) { pbsc::set_error("%s: %s", who, e.what()); return PBSC_ERR_INTERNAL; }               \
    catch (...) { pbsc::set_error("%s: unknown exception", who); return PBSC_ERR_INTERNAL; }

#define PBSC_CUDA(call)                                                             \
    do {                                                                            \
        cudaError_t _e = (call);                                                    \
        if (_e != cudaSuccess) return pbsc::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// Device blocks of per-batch buffers are kept in a per-device cache instead of going back to the driver: cudaMalloc/cudaFree
// of gigabytes per batch cost ~0.2 s per call of the end-to-end entry point (host trace, tools/e2e_trace.py).  pbsc_trim()
// (and the destruction of the last index on a device) returns the cache to the driver.
cudaError_t dev_cache_alloc(void** p, size_t bytes);
void dev_cache_free(void* p);
void dev_cache_trim(int device);

// RAII device buffer (released on scope exit); keeps the C ABI functions leak-free on error paths
template <class T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    ~DevBuf() { if (p) dev_cache_free(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t count)
    {
        if (p) { dev_cache_free(p); p = nullptr; }
        n = count;
        return dev_cache_alloc((void**)&p, (count ? count : 1) * sizeof(T));
    }
    void release() { if (p) dev_cache_free(p); p = nullptr; n = 0; }
};

// grow-only named device buffer owned by the index
inline cudaError_t arena_get(pbsc_index* idx, const char* name, size_t bytes, void** out)
{
    pbsc_index::ArenaBuf& a = idx->arena[name];
    if (a.cap < bytes || a.p == nullptr)
    {
        const bool tr = getenv("PBSC_ROUND_TRACE") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        if (a.p) cudaFree(a.p);
        a.p = nullptr; a.cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&a.p, want);
        if (e != cudaSuccess) return e;
        a.cap = want;
        if (tr) fprintf(stderr, "[pbsc round trace]     arena %-16s -> %8.1f MB in %7.2f ms\n", name, want / 1048576.0,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    *out = a.p;
    return cudaSuccess;
}

struct Timing
{
    float h2d_ms = 0, seed_ms = 0, extend_ms = 0, d2h_ms = 0, total_ms = 0;
    uint64_t kernel_launches = 0, seed_pairs = 0, rank_queries = 0;
    float dp_ms = 0;                  // part of extend_ms spent in the DP / multiple-alignment fallback
    uint64_t dp_jobs = 0, dp_rows = 0;
    float walk_ms = 0;                // walk_levels_kernel alone (CUDA events around each of its launches)
    uint64_t walk_launches = 0;
    uint64_t dp_thread_rows = 0;      // rows aligned by dp_align_thread_kernel
};
Timing& last_timing();

#define PBSC_MAX_LANES 4
// take a free lane of `primary` (blocks while all are busy); the returned object has the primary's tables and its own
// stream and arena.  `idmer_len` is the -i of the batch about to run: the idmer validity table is (re)built first, with no
// other lane running, when it differs from the table's.
int lane_acquire(pbsc_index* primary, int idmer_len, pbsc_index** lane);
void lane_release(pbsc_index* primary, pbsc_index* lane);
int ensure_idmer_table(pbsc_index* idx, int idmer_len);   // pbsc_extend_thread.cu

// 2-bit code of a base; -1 for anything else
inline int base_code(char b) { switch (b) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; } }

}  // namespace pbsc
#endif
