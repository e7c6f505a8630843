// pbsc_extend.cu — FM-extend phase.
//
//   extend_pairs_kernel   explicit (source, path, target) triples: LongReadSelfCorrectByOverlap ctor +
//                         extendOverlap as called from correctByFMExtension
//                         (PacBio/PacBioSelfCorrectionProcess.cpp:186-192); used by the parity tests
//   correct_reads_kernel  the per-read chain of PacBioSelfCorrectionProcess::initCorrect (:56-157) with
//                         --nodp semantics for failed walks: warp per read, walks in seed order because the
//                         source of walk i+1 is the corrected piece produced by walk i
//
// Both are persistent: each warp pulls work items from an atomic counter (reads longest first).
#include <algorithm>
#include <numeric>
#include <stdlib.h>
#include <string.h>
#include "pbsc_batch.cuh"
#include "pbsc_walk.cuh"

namespace pbsc {

constexpr int WARPS_PER_BLOCK = 4;
#ifndef PBSC_MIN_BLOCKS
#define PBSC_MIN_BLOCKS 2   // resident blocks per SM the register allocation is capped for
#endif

__device__ __forceinline__ unsigned long long warp_next(unsigned long long* counter)
{
    unsigned long long v = 0;
    if (lane_id() == 0) v = atomicAdd(counter, 1ull);
    return __shfl_sync(FULL, v, 0);
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, PBSC_MIN_BLOCKS)
extend_pairs_kernel(const __grid_constant__ FmIndexDev idx, const __grid_constant__ ExtParamsDev P, uint8_t* scratch, size_t scratch_stride, unsigned long long* counter,
                    uint64_t n_pairs, const uint8_t* __restrict__ src, const uint64_t* __restrict__ src_off,
                    const uint8_t* __restrict__ path, const uint64_t* __restrict__ path_off,
                    const uint8_t* __restrict__ trg, const uint64_t* __restrict__ trg_off,
                    const int32_t* __restrict__ dis, const int32_t* __restrict__ kk, const int32_t* __restrict__ min_sa,
                    int32_t* __restrict__ status, uint8_t* __restrict__ out, const uint64_t* __restrict__ out_region,
                    uint32_t* __restrict__ out_len)
{
    __shared__ WarpShared shared[WARPS_PER_BLOCK];
    const int warp_in_block = threadIdx.x >> 5;
    const int lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * WARPS_PER_BLOCK + warp_in_block;
    WarpShared& sh = shared[warp_in_block];
    WarpScratch ws;
    carve_scratch(scratch + warp_global * scratch_stride, P, ws);
    for (;;)
    {
        const unsigned long long i = warp_next(counter);
        if (i >= n_pairs) break;
        const uint32_t sl = (uint32_t)(src_off[i + 1] - src_off[i]);
        const uint32_t pl = (uint32_t)(path_off[i + 1] - path_off[i]);
        const uint32_t tl = (uint32_t)(trg_off[i + 1] - trg_off[i]);
        const int32_t k = kk[i];
        int st;
        uint32_t mlen = 0;
        if (k <= 0 || (uint32_t)k > sl || dis[i] != (int32_t)pl || (uint64_t)k + pl + tl > P.q_cap) st = PBSC_WALK_UNSUPPORTED;
        else
        {
            const uint8_t* s = src + src_off[i] + (sl - k);
            for (uint32_t x = lane; x < (uint32_t)k; x += 32) ws.q[x] = s[x];
            for (uint32_t x = lane; x < pl; x += 32) ws.q[k + x] = path[path_off[i] + x];
            for (uint32_t x = lane; x < tl; x += 32) ws.q[k + pl + x] = trg[trg_off[i] + x];
            __syncwarp();
            st = walk_pair(idx, P, ws, sh, (uint32_t)k + pl + tl, (uint32_t)k, dis[i], tl, (uint64_t)min_sa[i], &mlen);
            __syncwarp();
        }
        if (st == 1)
        {
            const uint64_t cap = out_region[i + 1] - out_region[i];
            if (mlen > cap) st = PBSC_WALK_OVERFLOW;
            else for (uint32_t x = lane; x < mlen; x += 32) out[out_region[i] + x] = ws.merged[x];
        }
        if (lane == 0) { status[i] = st; out_len[i] = st == 1 ? mlen : 0; }
        __syncwarp();
    }
}

struct ChainParamsDev
{
    int32_t start_kmer, next_target, split, pb_coverage;
};

// per-read chain: initCorrect + correctByFMExtension (PacBioSelfCorrectionProcess.cpp:56-206), failed walks take the
// --nodp branch (:146-153)
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, PBSC_MIN_BLOCKS)
correct_reads_kernel(const __grid_constant__ FmIndexDev idx, const __grid_constant__ ExtParamsDev P, ChainParamsDev C, uint8_t* scratch, size_t scratch_stride,
                     unsigned long long* counter, uint64_t n_reads, const uint32_t* __restrict__ order,
                     const uint8_t* __restrict__ codes, const uint64_t* __restrict__ offsets,
                     const pbsc_seed* __restrict__ seeds, const uint64_t* __restrict__ region, const uint32_t* __restrict__ seed_count,
                     uint8_t* __restrict__ pieces, const uint64_t* __restrict__ piece_region,
                     uint32_t* __restrict__ piece_bounds, const uint64_t* __restrict__ bounds_region,
                     pbsc_read_stats* __restrict__ stats, int32_t* __restrict__ read_status, unsigned long long* walk_counter)
{
    __shared__ WarpShared shared[WARPS_PER_BLOCK];
    const int warp_in_block = threadIdx.x >> 5;
    const int lane = lane_id();
    const size_t warp_global = (size_t)blockIdx.x * WARPS_PER_BLOCK + warp_in_block;
    WarpShared& sh = shared[warp_in_block];
    WarpScratch ws;
    carve_scratch(scratch + warp_global * scratch_stride, P, ws);
    for (;;)
    {
        const unsigned long long it = warp_next(counter);
        if (it >= n_reads) break;
        const uint32_t r = order[it];
        const uint8_t* read = codes + offsets[r];
        const int64_t L = (int64_t)(offsets[r + 1] - offsets[r]);
        const pbsc_seed* sv = seeds + region[r];
        const uint32_t ns = seed_count[r];
        uint8_t* piece = pieces + piece_region[r];
        const uint64_t pieceCap = piece_region[r + 1] - piece_region[r];
        uint32_t* bounds = piece_bounds + bounds_region[r];
        const uint64_t boundsCap = bounds_region[r + 1] - bounds_region[r];
        pbsc_read_stats st;
        memset(&st, 0, sizeof st);
        st.total_reads_len = L;
        st.total_seed_num = ns;
        int rstatus = 0;
        uint64_t plen = 0;       // bytes written to `piece` so far (all pieces)
        uint32_t nPieces = 0;
        if (ns >= 2)
        {
            // pieceVec.push_back(seedVec[0])
            uint64_t curStart = 0;      // start of the current piece inside `piece`
            pbsc_seed s0 = sv[0];
            if ((uint64_t)s0.len > pieceCap) rstatus = PBSC_WALK_OVERFLOW;
            else
            {
                for (int x = lane; x < s0.len; x += 32) piece[x] = read[s0.start + x];
                plen = s0.len;
            }
            __syncwarp();
            int64_t srcLen = s0.len;                       // source.seedLen (whole current piece)
            int srcEnd = s0.start + s0.len - 1;            // source.seedEndPos
            int srcEndBest = s0.end_best_k;
            int srcRepeat = s0.is_repeat;
            nPieces = 1;
            if (boundsCap) { if (lane == 0) bounds[0] = 0; } else rstatus = PBSC_WALK_OVERFLOW;
            for (uint32_t t = 1; t < ns && rstatus == 0; t++)
            {
                int success = 0, firstType = 0;
                for (int next = 0; next < C.next_target && t + next < ns; next++)
                {
                    const pbsc_seed tg = sv[t + next];
                    // ---- correctByFMExtension ----
                    const int interval = tg.start - srcEnd - 1;
                    int k = min(srcEndBest, tg.start_best_k) - 2;
                    if (srcRepeat || tg.is_repeat)
                    {
                        k = (int)min(srcLen, (int64_t)tg.len);
                        k = min(k, C.start_kmer + 2);
                    }
                    const uint64_t minSA = C.pb_coverage > 60 ? (uint64_t)((C.pb_coverage / 60) * 3) : 3;
                    const bool rtou = srcRepeat && !tg.is_repeat;
                    int stw;
                    uint32_t mlen = 0;
                    if (k <= 0 || k > srcLen || interval < 0) stw = PBSC_WALK_UNSUPPORTED;
                    else
                    {
                        const uint8_t* srcK = piece + plen - k;      // last k bases of the current piece
                        const uint8_t* pth = read + srcEnd + 1;
                        const uint8_t* trgS = read + tg.start;
                        uint32_t trgLen, qlen;
                        if (!rtou)
                        {
                            trgLen = tg.len; qlen = k + interval + trgLen;
                            if (qlen > P.q_cap) stw = PBSC_WALK_OVERFLOW;
                            else
                            {
                                for (int x = lane; x < k; x += 32) ws.q[x] = srcK[x];
                                for (int x = lane; x < interval; x += 32) ws.q[k + x] = pth[x];
                                for (uint32_t x = lane; x < trgLen; x += 32) ws.q[k + interval + x] = trgS[x];
                                stw = 0;
                            }
                        }
                        else
                        {
                            // query = revcomp(src_k + path + target[0..k))
                            trgLen = k; qlen = 2 * k + interval;
                            if (qlen > P.q_cap) stw = PBSC_WALK_OVERFLOW;
                            else
                            {
                                for (uint32_t x = lane; x < qlen; x += 32)
                                {
                                    const uint32_t y = qlen - 1 - x;   // index into src_k + path + target prefix
                                    const uint8_t c = y < (uint32_t)k ? srcK[y] : (y < (uint32_t)(k + interval) ? pth[y - k] : trgS[y - k - interval]);
                                    ws.q[x] = 3 - c;
                                }
                                stw = 0;
                            }
                        }
                        if (stw == 0)
                        {
                            __syncwarp();
                            if (lane == 0) atomicAdd(walk_counter, 1ull);
                            stw = walk_pair(idx, P, ws, sh, qlen, (uint32_t)k, interval, trgLen, minSA, &mlen);
                            __syncwarp();
                        }
                    }
                    if (stw == PBSC_WALK_OVERFLOW || stw == PBSC_WALK_UNSUPPORTED) { rstatus = stw; break; }
                    if (next == 0) firstType = stw;
                    if (stw > 0)
                    {
                        // out = mergedSeq.erase(0, extendKmerSize), after undoing the strand swap (:195-201)
                        uint64_t outLen;
                        if (!rtou)
                        {
                            outLen = mlen - k;
                            if (plen + outLen > pieceCap) { rstatus = PBSC_WALK_OVERFLOW; break; }
                            for (uint32_t x = lane; x < outLen; x += 32) piece[plen + x] = ws.merged[k + x];
                        }
                        else
                        {
                            // merged = revcomp(merged') + target.substr(k); then drop the first k
                            const uint64_t tailLen = tg.len - k;
                            outLen = (mlen - k) + tailLen;
                            if (plen + outLen > pieceCap) { rstatus = PBSC_WALK_OVERFLOW; break; }
                            for (uint32_t x = lane; x < mlen - k; x += 32) piece[plen + x] = 3 - ws.merged[mlen - 1 - (k + x)];
                            for (uint32_t x = lane; x < tailLen; x += 32) piece[plen + (mlen - k) + x] = read[tg.start + k + x];
                        }
                        __syncwarp();
                        plen += outLen;
                        st.corrected_len += outLen;
                        st.seed_dis += interval;
                        st.fm_num++;
                        st.total_walk_num++;
                        // source.append(mergedSeq, target)
                        srcLen += outLen;
                        srcEnd = tg.start + tg.len - 1; srcEndBest = tg.end_best_k; srcRepeat = tg.is_repeat;
                        t += next;
                        success = 1;
                        break;
                    }
                }
                if (rstatus) break;
                if (!success)
                {
                    const pbsc_seed tg = sv[t];
                    if (firstType == -1) st.high_error_num++;
                    else if (firstType == -2) st.exceed_depth_num++;
                    else if (firstType == -3) st.exceed_leave_num++;
                    else { rstatus = PBSC_WALK_NO_PATH; break; }   // "Does it really happen?" (:125-127)
                    st.total_walk_num++;
                    if (C.split)
                    {
                        // pieceVec.push_back(target)
                        if (plen + tg.len > pieceCap || nPieces >= boundsCap) { rstatus = PBSC_WALK_OVERFLOW; break; }
                        if (lane == 0) bounds[nPieces] = (uint32_t)plen;
                        nPieces++;
                        curStart = plen;
                        for (int x = lane; x < tg.len; x += 32) piece[plen + x] = read[tg.start + x];
                        plen += tg.len;
                        srcLen = tg.len;
                    }
                    else
                    {
                        // mergedSeq = readSeq.substr(source.seedEndPos + 1, target.seedEndPos - source.seedEndPos)
                        const int tgEnd = tg.start + tg.len - 1;
                        const uint64_t n = (uint64_t)(tgEnd - srcEnd);
                        if (plen + n > pieceCap) { rstatus = PBSC_WALK_OVERFLOW; break; }
                        for (uint32_t x = lane; x < n; x += 32) piece[plen + x] = read[srcEnd + 1 + x];
                        plen += n;
                        srcLen += n;
                    }
                    __syncwarp();
                    srcEnd = tg.start + tg.len - 1; srcEndBest = tg.end_best_k; srcRepeat = tg.is_repeat;
                    st.corrected_len += tg.len;
                }
            }
            (void)curStart;
        }
        st.merge = (ns >= 2) ? 1 : 0;
        st.n_pieces = (int32_t)nPieces;
        if (lane == 0)
        {
            if (nPieces && nPieces < boundsCap) bounds[nPieces] = (uint32_t)plen;
            stats[r] = st;
            read_status[r] = rstatus;
        }
        __syncwarp();
    }
}

// largest query (k + gap + target) any walk of the batch can need, and the largest gap
__global__ void chain_bounds_kernel(uint64_t n_reads, const pbsc_seed* __restrict__ seeds, const uint64_t* __restrict__ region,
                                    const uint32_t* __restrict__ seed_count, int next_target, unsigned int* max_gap, unsigned int* max_trg)
{
    uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const pbsc_seed* sv = seeds + region[r];
    const uint32_t ns = seed_count[r];
    unsigned int g = 0, tl = 0;
    for (uint32_t i = 0; i + 1 < ns; i++)
    {
        const int end = sv[i].start + sv[i].len - 1;
        for (int j = 1; j <= next_target && i + j < ns; j++)
        {
            g = max(g, (unsigned int)max(0, sv[i + j].start - end - 1));
            tl = max(tl, (unsigned int)sv[i + j].len);
        }
    }
    if (ns) tl = max(tl, (unsigned int)sv[0].len);
    atomicMax(max_gap, g);
    atomicMax(max_trg, tl);
}

void make_ext_params(const pbsc_params* p, ExtParamsDev& d, uint32_t q_cap, uint32_t node_cap, uint32_t merged_cap)
{
    memset(&d, 0, sizeof d);
    d.max_leaves = p->max_leaves; d.seed_size = p->idmer_len; d.min_overlap = p->min_kmer; d.pb_coverage = p->pb_coverage;
    // isInsufficientFreqs, LongReadCorrectByOverlap.cpp:339
    d.high_freq_thr = (size_t)p->pb_coverage > 60 ? (int)((size_t)((size_t)p->pb_coverage / 60) * 3) : 3;
    d.redeem_a = (size_t)(p->idmer_len - 1) * p->error_rate;
    d.redeem_b = 1 - p->error_rate;
    d.walk_error_rate = 0.25;
    for (int i = 0; i <= 100; i++) d.freq_int[i] = (int)p->freqs_of_kmer[i];
    d.q_cap = q_cap; d.node_cap = node_cap; d.merged_cap = merged_cap;
}

int launch_geometry(int device, int* blocks)
{
    int sms = 0;
    PBSC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    int per_sm = 0;
    PBSC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, correct_reads_kernel, WARPS_PER_BLOCK * 32, 0));
    if (per_sm < 1) per_sm = 1;
    *blocks = sms * per_sm;
    return PBSC_OK;
}

int alloc_extend_workspace(pbsc_index* idx, const pbsc_params* p, const std::vector<uint64_t>& h_offsets, DeviceBatch& b, SeedBuffers& s, Workspace& w, cudaStream_t st)
{
    const uint64_t n = b.n_reads;
    // per-read output regions sized from the read length alone, so nothing has to come back from the seed phase
    w.h_piece_region.assign(n + 1, 0);
    w.h_bounds_region.assign(n + 1, 0);
    int min_static = p->start_kmer;
    for (int m = 0; m < 3; m++) min_static = std::min(min_static, p->start_kmer + p->offset[m]);
    for (uint64_t r = 0; r < n; r++)
    {
        const uint64_t L = h_offsets[r + 1] - h_offsets[r];
        const uint64_t cap = (int64_t)L >= p->start_kmer ? (uint64_t)(w.piece_factor * (double)L) + 2048 : 0;
        w.h_piece_region[r + 1] = w.h_piece_region[r] + align_up(cap, 16);
        w.h_bounds_region[r + 1] = w.h_bounds_region[r] + (cap ? (p->split ? L / (uint64_t)std::max(min_static, 1) + 4 : 2) : 0);
    }
    PBSC_CUDA(w.pieces.alloc(w.h_piece_region[n])); PBSC_CUDA(w.piece_region.alloc(n + 1)); PBSC_CUDA(w.bounds_region.alloc(n + 1));
    PBSC_CUDA(w.bounds.alloc(w.h_bounds_region[n])); PBSC_CUDA(w.order.alloc(n)); PBSC_CUDA(w.stats.alloc(n)); PBSC_CUDA(w.status.alloc(n));
    PBSC_CUDA(w.counters.alloc(2)); PBSC_CUDA(w.maxima.alloc(2));
    // longest reads first: the chain of a read is sequential, so the tail of the batch is the longest chain
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return (h_offsets[a + 1] - h_offsets[a]) > (h_offsets[c + 1] - h_offsets[c]); });
    PBSC_CUDA(cudaMemcpyAsync(w.order.p, order.data(), n * 4, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(w.piece_region.p, w.h_piece_region.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(w.bounds_region.p, w.h_bounds_region.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    int rc = launch_geometry(idx->device, &w.blocks);
    if (rc != PBSC_OK) return rc;
    if ((uint64_t)w.blocks * WARPS_PER_BLOCK > n) w.blocks = (int)((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    if (w.blocks < 1) w.blocks = 1;
    if (w.q_cap == 0) w.q_cap = 2048;
    w.merged_cap = (uint32_t)align_up((size_t)(1.2 * (w.q_cap + 10)) + 256, 16);
    w.thread_engine = use_thread_engine();
    if (w.node_cap == 0) w.node_cap = w.thread_engine ? (1u << 13) : (1u << 15);
    w.scratch_stride = warp_scratch_bytes(w.q_cap, w.node_cap, w.merged_cap);
    if (!w.thread_engine) PBSC_CUDA(w.scratch.alloc(w.scratch_stride * (size_t)w.blocks * WARPS_PER_BLOCK));
    PBSC_CUDA(cudaStreamSynchronize(st));
    return PBSC_OK;
}

bool use_thread_engine()
{
    const char* e = getenv("PBSC_ENGINE");
    return !(e && strcmp(e, "warp") == 0);
}

int run_extend_chain(pbsc_index* idx, const pbsc_params* p, DeviceBatch& b, SeedBuffers& s, Workspace& w, uint64_t* launches)
{
    cudaStream_t st = idx->stream;
    const uint64_t n = b.n_reads;
    if (n == 0) return PBSC_OK;
    // ---- does the largest seed gap of this batch fit the per-warp query scratch? (8-byte round trip) ----
    PBSC_CUDA(cudaMemsetAsync(w.maxima.p, 0, 8, st));
    chain_bounds_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, s.seeds.p, s.region.p, s.count.p, p->next_target, w.maxima.p, w.maxima.p + 1);
    unsigned int hmax[2] = {0, 0};
    PBSC_CUDA(cudaMemcpyAsync(hmax, w.maxima.p, 8, cudaMemcpyDeviceToHost, st));
    PBSC_CUDA(cudaStreamSynchronize(st));
    const uint32_t need_q = (uint32_t)align_up((size_t)hmax[0] + hmax[1] + 64 + 16, 16);
    if (need_q > w.q_cap)
    {
        w.q_cap = need_q;
        w.merged_cap = (uint32_t)align_up((size_t)(1.2 * (w.q_cap + 10)) + 256, 16);
        w.scratch_stride = warp_scratch_bytes(w.q_cap, w.node_cap, w.merged_cap);
        PBSC_CUDA(w.scratch.alloc(w.scratch_stride * (size_t)w.blocks * WARPS_PER_BLOCK));
    }
    ExtParamsDev P;
    make_ext_params(p, P, w.q_cap, w.node_cap, w.merged_cap);
    ChainParamsDev C;
    C.start_kmer = p->start_kmer; C.next_target = p->next_target; C.split = p->split; C.pb_coverage = p->pb_coverage;
    PBSC_CUDA(cudaMemsetAsync(w.counters.p, 0, 16, st));
    correct_reads_kernel<<<w.blocks, WARPS_PER_BLOCK * 32, 0, st>>>(idx->dev, P, C, w.scratch.p, w.scratch_stride, w.counters.p, n, w.order.p, b.codes.p, b.offsets.p,
                                                                    s.seeds.p, s.region.p, s.count.p, w.pieces.p, w.piece_region.p, w.bounds.p,
                                                                    w.bounds_region.p, w.stats.p, w.status.p, w.counters.p + 1);
    if (launches) *launches += 2;
    PBSC_CUDA(cudaGetLastError());
    return PBSC_OK;
}

}  // namespace pbsc

using namespace pbsc;

extern "C" int pbsc_extend_batch(pbsc_index* idx, const pbsc_params* p, uint64_t n_pairs,
                                 const char* src, const uint64_t* src_off, const char* path, const uint64_t* path_off,
                                 const char* trg, const uint64_t* trg_off, const int32_t* dis, const int32_t* k,
                                 const int32_t* min_sa, int32_t* status, char* out, uint64_t out_cap, uint64_t* out_offsets)
try
{
    if (!idx || !p || !src || !src_off || !path || !path_off || !trg || !trg_off || !dis || !k || !min_sa || !status || !out_offsets)
    { set_error("pbsc_extend_batch: null argument"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    const uint64_t n = n_pairs;
    out_offsets[0] = 0;
    if (n == 0) return PBSC_OK;
    auto to_codes = [&](const char* s, uint64_t len, std::vector<uint8_t>& v) -> bool {
        v.resize(len + 1);
        for (uint64_t i = 0; i < len; i++) { int c = base_code(s[i]); if (c < 0) return false; v[i] = (uint8_t)c; }
        return true;
    };
    std::vector<uint8_t> hs, hp, ht;
    if (!to_codes(src, src_off[n], hs) || !to_codes(path, path_off[n], hp) || !to_codes(trg, trg_off[n], ht)) { set_error("pbsc_extend_batch: non-ACGT character"); return PBSC_ERR_ARG; }
    uint32_t maxq = 0, maxm = 0;
    std::vector<uint64_t> region(n + 1, 0);
    for (uint64_t i = 0; i < n; i++)
    {
        const uint64_t tl = trg_off[i + 1] - trg_off[i], pl = path_off[i + 1] - path_off[i];
        const uint64_t q = (uint64_t)std::max(k[i], 0) + pl + tl;
        const uint64_t m = (uint64_t)(1.2 * ((double)pl + 10)) + 2 * (uint64_t)std::max(k[i], 0) + tl + 16;
        maxq = std::max<uint64_t>(maxq, q); maxm = std::max<uint64_t>(maxm, m);
        region[i + 1] = region[i] + m;
    }
    ExtParamsDev P;
    make_ext_params(p, P, (uint32_t)align_up(maxq + 16, 16), 1u << 16, (uint32_t)align_up(maxm + 16, 16));
    DevBuf<uint8_t> ds, dp, dt, dout, dscratch; DevBuf<uint64_t> dso, dpo, dto, dreg; DevBuf<int32_t> ddis, dk, dsa, dstat; DevBuf<uint32_t> dlen;
    DevBuf<unsigned long long> dctr;
    PBSC_CUDA(ds.alloc(hs.size())); PBSC_CUDA(dp.alloc(hp.size())); PBSC_CUDA(dt.alloc(ht.size())); PBSC_CUDA(dout.alloc(region[n]));
    PBSC_CUDA(dso.alloc(n + 1)); PBSC_CUDA(dpo.alloc(n + 1)); PBSC_CUDA(dto.alloc(n + 1)); PBSC_CUDA(dreg.alloc(n + 1));
    PBSC_CUDA(ddis.alloc(n)); PBSC_CUDA(dk.alloc(n)); PBSC_CUDA(dsa.alloc(n)); PBSC_CUDA(dstat.alloc(n)); PBSC_CUDA(dlen.alloc(n)); PBSC_CUDA(dctr.alloc(1));
    PBSC_CUDA(cudaMemcpyAsync(ds.p, hs.data(), hs.size(), cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(dp.p, hp.data(), hp.size(), cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(dt.p, ht.data(), ht.size(), cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(dso.p, src_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(dpo.p, path_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(dto.p, trg_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(dreg.p, region.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(ddis.p, dis, n * 4, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(dk.p, k, n * 4, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(dsa.p, min_sa, n * 4, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemsetAsync(dctr.p, 0, 8, st));
    int blocks = 0;
    int rc = launch_geometry(idx->device, &blocks);
    if (rc != PBSC_OK) return rc;
    if ((uint64_t)blocks * WARPS_PER_BLOCK > n) blocks = (int)((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    const size_t stride = warp_scratch_bytes(P.q_cap, P.node_cap, P.merged_cap);
    PBSC_CUDA(dscratch.alloc(stride * (size_t)blocks * WARPS_PER_BLOCK));
    extend_pairs_kernel<<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(idx->dev, P, dscratch.p, stride, dctr.p, n, ds.p, dso.p, dp.p, dpo.p, dt.p, dto.p,
                                                                 ddis.p, dk.p, dsa.p, dstat.p, dout.p, dreg.p, dlen.p);
    PBSC_CUDA(cudaGetLastError());
    std::vector<uint8_t> hout(region[n]);
    std::vector<uint32_t> hlen(n);
    PBSC_CUDA(cudaMemcpyAsync(status, dstat.p, n * 4, cudaMemcpyDeviceToHost, st));
    PBSC_CUDA(cudaMemcpyAsync(hlen.data(), dlen.p, n * 4, cudaMemcpyDeviceToHost, st));
    PBSC_CUDA(cudaMemcpyAsync(hout.data(), dout.p, region[n], cudaMemcpyDeviceToHost, st));
    PBSC_CUDA(cudaStreamSynchronize(st));
    for (uint64_t i = 0; i < n; i++) out_offsets[i + 1] = out_offsets[i] + hlen[i];
    if (out_offsets[n] > out_cap || (!out && out_offsets[n])) { set_error("pbsc_extend_batch: output needs %llu bytes", (unsigned long long)out_offsets[n]); return PBSC_ERR_LIMIT; }
    for (uint64_t i = 0; i < n; i++)
        for (uint32_t x = 0; x < hlen[i]; x++) out[out_offsets[i] + x] = "ACGT"[hout[region[i] + x] & 3];
    return PBSC_OK;
}
PBSC_CATCH_ALL("pbsc_extend_batch")
