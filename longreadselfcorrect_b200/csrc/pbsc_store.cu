// pbsc_store.cu — the index as ONE relocatable blob (header + tables), and what follows from that:
//   * PREFIX.fmg, the persisted flat index: a later run maps the tables straight into HBM instead of
//     decoding the run-length bytes of PREFIX.bwt/.rbwt (SuffixTools/BWTReaderBinary.cpp:55-85, RLBWT::initializeFMIndex,
//     SuffixTools/RLBWT.cpp:109-248) and rebuilding 2.1 GB of prefix table;
//   * pbsc_index_clone: GPU 0 loads the index once, the other GPUs of the box receive it over NVLink with one peer copy per
//     table (plain cudaMemcpyPeerAsync) instead of eight file reads, eight host decodes and eight PCIe uploads
//     (replaces the `omp parallel sections` load of StriDe/PacBioSelfCorrection.cpp:155-172 for the multi-GPU case);
//   * pbsc_index_export_blob / pbsc_index_import_blob: the same bytes through a caller-owned device buffer, for callers that
//     distribute the index themselves (bench.py broadcasts it with NCCL between the one-process-per-GPU ranks);
//   * lanes: several batches of one index in flight at once (see pbsc_internal.h).
#include <string.h>
#include <sys/stat.h>
#include <algorithm>
#include <chrono>
#include <fstream>
#include "pbsc_internal.h"

namespace pbsc {

struct FmgHeader
{
    char magic[8];                 // "PBSCFMG" + format version byte
    uint32_t header_bytes, flags;
    uint64_t total_bytes;
    uint64_t n_symbols[2], n_strings[2], n_blocks[2], n_dollar[2];
    uint64_t src_runs[2];          // run bytes of PREFIX.bwt / PREFIX.rbwt the tables were decoded from (0 = not from files)
    uint64_t total[2][5];          // symbol counts $ A C G T
    int32_t k0, idmer_len;
    uint64_t off_blocks[2], off_dollar[2], off_dmask[2], off_prefix, off_idmer;   // byte offsets from the start of the blob
    uint64_t bytes_prefix, bytes_idmer;
    uint64_t checksum;             // FNV-1a over every header byte before this field
    uint8_t pad[512 - 8 - 8 - 8 - 8 * 8 - 16 - 80 - 8 - 8 * 8 - 16 - 8];
};
static_assert(sizeof(FmgHeader) == 512, "FmgHeader is 512 bytes on disk");
static const char FMG_MAGIC[8] = {'P', 'B', 'S', 'C', 'F', 'M', 'G', 1};

static uint64_t fnv1a(const void* p, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static uint64_t up256(uint64_t x) { return (x + 255) / 256 * 256; }

static void make_header(const pbsc_index* idx, FmgHeader& H)
{
    memset(&H, 0, sizeof H);
    memcpy(H.magic, FMG_MAGIC, 8);
    H.header_bytes = sizeof(FmgHeader);
    uint64_t off = sizeof(FmgHeader);
    for (int w = 0; w < 2; w++)
    {
        const FmTable& t = idx->dev.t[w];
        H.n_symbols[w] = idx->n_symbols[w]; H.n_strings[w] = idx->n_strings[w]; H.n_blocks[w] = idx->n_blocks[w];
        H.n_dollar[w] = t.n_dollar; H.src_runs[w] = idx->src_runs[w];
        H.total[w][0] = t.C[0];
        for (int c = 0; c < 4; c++) H.total[w][c + 1] = t.total[c];
    }
    for (int w = 0; w < 2; w++) { H.off_blocks[w] = off; off = up256(off + H.n_blocks[w] * sizeof(FmBlock)); }
    for (int w = 0; w < 2; w++) { H.off_dollar[w] = off; off = up256(off + (H.n_dollar[w] + 1) * sizeof(uint32_t)); }
    for (int w = 0; w < 2; w++) { H.off_dmask[w] = off; off = up256(off + H.n_blocks[w] * sizeof(uint64_t)); }
    H.k0 = idx->dev.prefix ? idx->dev.k0 : 0;
    H.bytes_prefix = H.k0 ? (sizeof(PrefixEntry) << (2 * H.k0)) : 0;
    H.off_prefix = H.bytes_prefix ? off : 0; off = up256(off + H.bytes_prefix);
    H.idmer_len = idx->dev.idmer_valid ? idx->dev.idmer_len : 0;
    H.bytes_idmer = H.idmer_len ? (1ull << (2 * H.idmer_len)) : 0;
    H.off_idmer = H.bytes_idmer ? off : 0; off = up256(off + H.bytes_idmer);
    H.total_bytes = off;
    H.checksum = fnv1a(&H, offsetof(FmgHeader, checksum));
}

static int check_header(const FmgHeader& H, uint64_t have_bytes, const char* what)
{
    if (memcmp(H.magic, FMG_MAGIC, 7) != 0) { set_error("%s is not a flat-index blob (bad magic)", what); return PBSC_ERR_FORMAT; }
    if (H.magic[7] != FMG_MAGIC[7]) { set_error("%s has format version %d, this build reads version %d", what, (int)H.magic[7], (int)FMG_MAGIC[7]); return PBSC_ERR_FORMAT; }
    if (H.header_bytes != sizeof(FmgHeader) || H.checksum != fnv1a(&H, offsetof(FmgHeader, checksum))) { set_error("%s: corrupt header", what); return PBSC_ERR_FORMAT; }
    if (have_bytes < H.total_bytes) { set_error("%s is truncated: %llu of %llu bytes", what, (unsigned long long)have_bytes, (unsigned long long)H.total_bytes); return PBSC_ERR_FORMAT; }
    for (int w = 0; w < 2; w++)
        if (H.n_blocks[w] != H.n_symbols[w] / 64 + 1 || H.off_blocks[w] + H.n_blocks[w] * sizeof(FmBlock) > H.total_bytes || H.off_dmask[w] + H.n_blocks[w] * 8 > H.total_bytes ||
            H.off_dollar[w] + (H.n_dollar[w] + 1) * 4 > H.total_bytes)
        { set_error("%s: inconsistent table geometry", what); return PBSC_ERR_FORMAT; }
    if (H.k0 < 0 || H.k0 > 15 || H.off_prefix + H.bytes_prefix > H.total_bytes || H.off_idmer + H.bytes_idmer > H.total_bytes) { set_error("%s: inconsistent prefix table", what); return PBSC_ERR_FORMAT; }
    return PBSC_OK;
}

struct Section { const void* src; uint64_t off, bytes; };
static int sections_of(const pbsc_index* idx, const FmgHeader& H, Section out[8])
{
    int n = 0;
    for (int w = 0; w < 2; w++) out[n++] = Section{idx->d_blocks[w], H.off_blocks[w], H.n_blocks[w] * sizeof(FmBlock)};
    for (int w = 0; w < 2; w++) out[n++] = Section{idx->d_dollar[w], H.off_dollar[w], (H.n_dollar[w] + 1) * sizeof(uint32_t)};
    for (int w = 0; w < 2; w++) out[n++] = Section{idx->d_dmask[w], H.off_dmask[w], H.n_blocks[w] * sizeof(uint64_t)};
    if (H.bytes_prefix) out[n++] = Section{idx->d_prefix, H.off_prefix, H.bytes_prefix};
    if (H.bytes_idmer) out[n++] = Section{idx->d_idmer_valid, H.off_idmer, H.bytes_idmer};
    return n;
}

// a new index object over a blob that already sits in `blob` on `device` (ownership passes to the index)
static int adopt_blob(void* blob, const FmgHeader& H, int device, pbsc_index** out)
{
    pbsc_index* idx = new pbsc_index();
    idx->device = device;
    idx->blob = blob; idx->blob_bytes = H.total_bytes;
    uint8_t* b = (uint8_t*)blob;
    for (int w = 0; w < 2; w++)
    {
        idx->d_blocks[w] = (FmBlock*)(b + H.off_blocks[w]);
        idx->d_dollar[w] = (uint32_t*)(b + H.off_dollar[w]);
        idx->d_dmask[w] = (uint64_t*)(b + H.off_dmask[w]);
        idx->n_symbols[w] = H.n_symbols[w]; idx->n_strings[w] = H.n_strings[w]; idx->n_blocks[w] = H.n_blocks[w]; idx->src_runs[w] = H.src_runs[w];
        FmTable& t = idx->dev.t[w];
        t.blocks = idx->d_blocks[w]; t.dollar_pos = idx->d_dollar[w]; t.dollar_mask = idx->d_dmask[w];
        t.n = H.n_symbols[w]; t.n_dollar = (uint32_t)H.n_dollar[w];
        t.C[0] = H.total[w][0]; t.C[1] = t.C[0] + H.total[w][1]; t.C[2] = t.C[1] + H.total[w][2]; t.C[3] = t.C[2] + H.total[w][3];
        for (int c = 0; c < 4; c++) t.total[c] = H.total[w][c + 1];
    }
    idx->d_prefix = H.bytes_prefix ? (PrefixEntry*)(b + H.off_prefix) : nullptr;
    idx->dev.prefix = idx->d_prefix; idx->dev.k0 = H.bytes_prefix ? H.k0 : 0;
    idx->d_idmer_valid = H.bytes_idmer ? b + H.off_idmer : nullptr;
    idx->dev.idmer_valid = idx->d_idmer_valid; idx->dev.idmer_len = H.bytes_idmer ? H.idmer_len : 0;
    idx->device_bytes = H.total_bytes - sizeof(FmgHeader);
    if (cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking) != cudaSuccess)
    { const int rc = cuda_fail(cudaGetLastError(), "cudaStreamCreate", __FILE__, __LINE__); idx->blob = nullptr; delete idx; return rc; }
    cudaDeviceGetAttribute(&idx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = idx;
    return PBSC_OK;
}

static int check_device(int device, const char* who)
{
    int ndev = 0;
    PBSC_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("%s: device %d not available (%d devices)", who, device, ndev); return PBSC_ERR_CUDA; }
    return PBSC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// lanes
// ---------------------------------------------------------------------------------------------------------------
int lane_acquire(pbsc_index* primary, int idmer_len, pbsc_index** lane)
{
    std::unique_lock<std::mutex> lk(primary->run_mu);
    if (primary->lane_busy.empty()) primary->lane_busy.assign(1, 0);
    if (primary->dev.idmer_len != idmer_len || !primary->dev.idmer_valid)
    {
        // the table is shared by every lane: rebuild it with nobody running
        primary->lane_cv.wait(lk, [&] { return primary->lanes_active == 0; });
        const int rc = ensure_idmer_table(primary, idmer_len);
        if (rc != PBSC_OK) return rc;
    }
    size_t pick = 0;
    primary->lane_cv.wait(lk, [&] {
        for (size_t i = 0; i < primary->lane_busy.size(); i++) if (!primary->lane_busy[i]) { pick = i; return true; }
        return false;
    });
    primary->lane_busy[pick] = 1;
    primary->lanes_active++;
    pbsc_index* L = pick == 0 ? primary : primary->shadows[pick - 1];
    if (pick)
    {
        L->dev = primary->dev;   // tables may have been rebuilt (prefix table, idmer table) since the lane last ran
        L->d_prefix = primary->d_prefix; L->d_idmer_valid = primary->d_idmer_valid;
        L->learned_node_cap = primary->learned_node_cap; L->learned_piece_factor = primary->learned_piece_factor; L->learned_pool_nodes = primary->learned_pool_nodes;
    }
    *lane = L;
    return PBSC_OK;
}

void lane_release(pbsc_index* primary, pbsc_index* lane)
{
    {
        std::lock_guard<std::mutex> lk(primary->run_mu);
        size_t pick = 0;
        for (size_t i = 0; i < primary->shadows.size(); i++) if (primary->shadows[i] == lane) pick = i + 1;
        primary->lane_busy[pick] = 0;
        primary->lanes_active--;
    }
    primary->lane_cv.notify_all();
}

}  // namespace pbsc

using namespace pbsc;

extern "C" {

int pbsc_index_set_lanes(pbsc_index* idx, int lanes)
try
{
    if (!idx || idx->primary || lanes < 1 || lanes > PBSC_MAX_LANES) { set_error("pbsc_index_set_lanes: lanes must be in 1..%d", PBSC_MAX_LANES); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    std::unique_lock<std::mutex> lk(idx->run_mu);
    idx->lane_cv.wait(lk, [&] { return idx->lanes_active == 0; });
    while ((int)idx->shadows.size() + 1 > lanes)
    {
        pbsc_index* s = idx->shadows.back();
        idx->shadows.pop_back();
        for (auto& kv : s->arena) if (kv.second.p) cudaFree(kv.second.p);
        if (s->stream) cudaStreamDestroy(s->stream);
        if (s->stream2) cudaStreamDestroy(s->stream2);
        if (s->ev_a) cudaEventDestroy(s->ev_a);
        if (s->ev_b) cudaEventDestroy(s->ev_b);
        delete s;
    }
    while ((int)idx->shadows.size() + 1 < lanes)
    {
        pbsc_index* s = new pbsc_index();
        s->primary = idx; s->device = idx->device; s->dev = idx->dev; s->sm_count = idx->sm_count;
        for (int w = 0; w < 2; w++) { s->d_blocks[w] = idx->d_blocks[w]; s->d_dollar[w] = idx->d_dollar[w]; s->d_dmask[w] = idx->d_dmask[w]; s->n_symbols[w] = idx->n_symbols[w]; s->n_strings[w] = idx->n_strings[w]; s->n_blocks[w] = idx->n_blocks[w]; }
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { delete s; return cuda_fail(cudaGetLastError(), "cudaStreamCreate", __FILE__, __LINE__); }
        idx->shadows.push_back(s);
    }
    idx->lane_busy.assign((size_t)lanes, 0);
    return PBSC_OK;
}
PBSC_CATCH_ALL("pbsc_index_set_lanes")

int pbsc_index_lanes(const pbsc_index* idx) { return idx ? (int)idx->shadows.size() + 1 : 0; }

int pbsc_index_blob_size(const pbsc_index* idx, uint64_t* bytes)
{
    if (!idx || !bytes) { set_error("pbsc_index_blob_size: null argument"); return PBSC_ERR_ARG; }
    FmgHeader H;
    make_header(idx, H);
    *bytes = H.total_bytes;
    return PBSC_OK;
}

int pbsc_index_export_blob(const pbsc_index* idx, void* d_dst, uint64_t cap)
{
    if (!idx || !d_dst) { set_error("pbsc_index_export_blob: null argument"); return PBSC_ERR_ARG; }
    FmgHeader H;
    make_header(idx, H);
    if (cap < H.total_bytes) { set_error("pbsc_index_export_blob: the blob needs %llu bytes", (unsigned long long)H.total_bytes); return PBSC_ERR_LIMIT; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    PBSC_CUDA(cudaMemcpyAsync(d_dst, &H, sizeof H, cudaMemcpyHostToDevice, idx->stream));
    Section sec[8];
    const int ns = sections_of(idx, H, sec);
    for (int i = 0; i < ns; i++) PBSC_CUDA(cudaMemcpyAsync((uint8_t*)d_dst + sec[i].off, sec[i].src, sec[i].bytes, cudaMemcpyDeviceToDevice, idx->stream));
    PBSC_CUDA(cudaStreamSynchronize(idx->stream));
    return PBSC_OK;
}

int pbsc_index_import_blob(const void* src, uint64_t bytes, int src_device, int device, pbsc_index** out)
try
{
    if (!src || !out || bytes < sizeof(FmgHeader)) { set_error("pbsc_index_import_blob: bad argument"); return PBSC_ERR_ARG; }
    *out = nullptr;
    int rc = check_device(device, "pbsc_index_import_blob");
    if (rc != PBSC_OK) return rc;
    PBSC_CUDA(cudaSetDevice(device));
    FmgHeader H;
    if (src_device < 0) memcpy(&H, src, sizeof H);
    else if (src_device == device) PBSC_CUDA(cudaMemcpy(&H, src, sizeof H, cudaMemcpyDeviceToHost));
    else { PBSC_CUDA(cudaSetDevice(src_device)); PBSC_CUDA(cudaMemcpy(&H, src, sizeof H, cudaMemcpyDeviceToHost)); PBSC_CUDA(cudaSetDevice(device)); }
    rc = check_header(H, bytes, "blob");
    if (rc != PBSC_OK) return rc;
    void* blob = nullptr;
    PBSC_CUDA(cudaMalloc(&blob, H.total_bytes));
    cudaError_t e;
    if (src_device < 0) e = cudaMemcpy(blob, src, H.total_bytes, cudaMemcpyHostToDevice);
    else if (src_device == device) e = cudaMemcpy(blob, src, H.total_bytes, cudaMemcpyDeviceToDevice);
    else e = cudaMemcpyPeer(blob, device, src, src_device, H.total_bytes);
    if (e != cudaSuccess) { cudaFree(blob); return cuda_fail(e, "copy of the index blob", __FILE__, __LINE__); }
    rc = adopt_blob(blob, H, device, out);
    if (rc != PBSC_OK) cudaFree(blob);
    return rc;
}
PBSC_CATCH_ALL("pbsc_index_import_blob")

// GPU `device` receives the tables of `src` (resident on another GPU of the box, or the same one) by peer copies: one
// cudaMemcpyPeerAsync per table over NVLink, no host decode, no PCIe upload
int pbsc_index_clone(const pbsc_index* src, int device, pbsc_index** out)
try
{
    if (!src || !out) { set_error("pbsc_index_clone: null argument"); return PBSC_ERR_ARG; }
    *out = nullptr;
    int rc = check_device(device, "pbsc_index_clone");
    if (rc != PBSC_OK) return rc;
    FmgHeader H;
    make_header(src, H);
    PBSC_CUDA(cudaSetDevice(device));
    if (device != src->device)
    {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, device, src->device);
        if (can) { const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else cudaGetLastError(); }
        // without peer access cudaMemcpyPeerAsync stages through the host: slower, still correct
    }
    void* blob = nullptr;
    PBSC_CUDA(cudaMalloc(&blob, H.total_bytes));
    cudaStream_t st = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMemcpyAsync(blob, &H, sizeof H, cudaMemcpyHostToDevice, st);
    Section sec[8];
    const int ns = sections_of(src, H, sec);
    for (int i = 0; i < ns && e == cudaSuccess; i++)
        e = device == src->device ? cudaMemcpyAsync((uint8_t*)blob + sec[i].off, sec[i].src, sec[i].bytes, cudaMemcpyDeviceToDevice, st)
                                  : cudaMemcpyPeerAsync((uint8_t*)blob + sec[i].off, device, sec[i].src, src->device, sec[i].bytes, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (st) cudaStreamDestroy(st);
    if (e != cudaSuccess) { cudaFree(blob); return cuda_fail(e, "peer copy of the index", __FILE__, __LINE__); }
    rc = adopt_blob(blob, H, device, out);
    if (rc != PBSC_OK) cudaFree(blob);
    return rc;
}
PBSC_CATCH_ALL("pbsc_index_clone")

// PREFIX.fmg: header + tables exactly as they sit in HBM
int pbsc_index_save(const pbsc_index* idx, const char* path)
try
{
    if (!idx || !path) { set_error("pbsc_index_save: null argument"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    FmgHeader H;
    make_header(idx, H);
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) { set_error("cannot write %s", tmp.c_str()); return PBSC_ERR_IO; }
    const size_t CH = 64u << 20;
    void* stage = nullptr;
    if (cudaHostAlloc(&stage, CH, cudaHostAllocDefault) != cudaSuccess) { fclose(f); remove(tmp.c_str()); return cuda_fail(cudaGetLastError(), "cudaHostAlloc", __FILE__, __LINE__); }
    bool ok = fwrite(&H, sizeof H, 1, f) == 1;
    uint64_t pos = sizeof H;
    Section sec[8];
    const int ns = sections_of(idx, H, sec);
    static const char zeros[256] = {0};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < ns && ok && e == cudaSuccess; i++)
    {
        while (pos < sec[i].off && ok) { const size_t z = (size_t)std::min<uint64_t>(256, sec[i].off - pos); ok = fwrite(zeros, 1, z, f) == z; pos += z; }
        for (uint64_t done = 0; done < sec[i].bytes && ok; done += CH)
        {
            const size_t nb = (size_t)std::min<uint64_t>(CH, sec[i].bytes - done);
            e = cudaMemcpy(stage, (const uint8_t*)sec[i].src + done, nb, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) break;
            ok = fwrite(stage, 1, nb, f) == nb;
            pos += nb;
        }
    }
    while (pos < H.total_bytes && ok) { const size_t z = (size_t)std::min<uint64_t>(256, H.total_bytes - pos); ok = fwrite(zeros, 1, z, f) == z; pos += z; }
    cudaFreeHost(stage);
    ok = (fclose(f) == 0) && ok;
    if (e != cudaSuccess) { remove(tmp.c_str()); return cuda_fail(e, "copy of the index to the host", __FILE__, __LINE__); }
    if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); set_error("cannot write %s", path); return PBSC_ERR_IO; }
    return PBSC_OK;
}
PBSC_CATCH_ALL("pbsc_index_save")

int pbsc_index_load_fmg(const char* path, int device, pbsc_index** out)
try
{
    if (!path || !out) { set_error("pbsc_index_load_fmg: null argument"); return PBSC_ERR_ARG; }
    *out = nullptr;
    int rc = check_device(device, "pbsc_index_load_fmg");
    if (rc != PBSC_OK) return rc;
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return PBSC_ERR_IO; }
    struct stat sb;
    FmgHeader H;
    if (fstat(fileno(f), &sb) != 0 || fread(&H, sizeof H, 1, f) != 1) { fclose(f); set_error("%s: truncated header", path); return PBSC_ERR_FORMAT; }
    rc = check_header(H, (uint64_t)sb.st_size, path);
    if (rc != PBSC_OK) { fclose(f); return rc; }
    cudaError_t e = cudaSetDevice(device);
    void* blob = nullptr;
    void* stage[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    const size_t CH = 64u << 20;
    if (e == cudaSuccess) e = cudaMalloc(&blob, H.total_bytes);
    if (e == cudaSuccess) e = cudaHostAlloc(&stage[0], CH, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(&stage[1], CH, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    bool ok = true;
    if (e == cudaSuccess) e = cudaMemcpyAsync(blob, &H, sizeof H, cudaMemcpyHostToDevice, st);
    // double-buffered: the file read of chunk i+1 overlaps the H2D copy of chunk i
    int which = 0;
    for (uint64_t pos = sizeof H; pos < H.total_bytes && ok && e == cudaSuccess; which ^= 1)
    {
        const size_t nb = (size_t)std::min<uint64_t>(CH, H.total_bytes - pos);
        e = cudaEventSynchronize(ev[which]);   // the copy that last used this staging buffer
        if (e != cudaSuccess) break;
        ok = fread(stage[which], 1, nb, f) == nb;
        if (!ok) break;
        e = cudaMemcpyAsync((uint8_t*)blob + pos, stage[which], nb, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(ev[which], st);
        pos += nb;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    fclose(f);
    for (int i = 0; i < 2; i++) { if (ev[i]) cudaEventDestroy(ev[i]); if (stage[i]) cudaFreeHost(stage[i]); }
    if (st) cudaStreamDestroy(st);
    if (e != cudaSuccess) { if (blob) cudaFree(blob); return cuda_fail(e, "upload of the flat index", __FILE__, __LINE__); }
    if (!ok) { cudaFree(blob); set_error("%s: short read", path); return PBSC_ERR_FORMAT; }
    rc = adopt_blob(blob, H, device, out);
    if (rc != PBSC_OK) cudaFree(blob);
    return rc;
}
PBSC_CATCH_ALL("pbsc_index_load_fmg")

static bool read_bwt_header(const std::string& path, uint64_t& n_strings, uint64_t& n_symbols, uint64_t& n_runs)
{
    std::ifstream in(path.c_str(), std::ios::binary);
    uint16_t magic = 0;
    in.read((char*)&magic, 2); in.read((char*)&n_strings, 8); in.read((char*)&n_symbols, 8); in.read((char*)&n_runs, 8);
    return (bool)in && magic == 0xCACA;
}

// What `pbcorrect` does with -p PREFIX: use PREFIX.fmg when it exists, is well-formed and describes the same
// PREFIX.bwt / PREFIX.rbwt (string, symbol and run counts of both headers) with the wanted prefix table; otherwise decode
// the run-length files, build the prefix table and, if write_fmg, leave PREFIX.fmg for the next run.
// *from_fmg (optional) tells which way it went.
int pbsc_index_open(const char* prefix, int device, int require_sai, int k0, int write_fmg, int* from_fmg, pbsc_index** out)
try
{
    if (!prefix || !out) { set_error("pbsc_index_open: null argument"); return PBSC_ERR_ARG; }
    *out = nullptr;
    if (from_fmg) *from_fmg = 0;
    const std::string p(prefix), fmg = p + ".fmg";
    uint64_t ns[2], nsym[2], nr[2];
    const bool have_bwt = read_bwt_header(p + ".bwt", ns[0], nsym[0], nr[0]) && read_bwt_header(p + ".rbwt", ns[1], nsym[1], nr[1]);
    struct stat sb;
    if (require_sai && stat((p + ".sai").c_str(), &sb) != 0) { set_error("cannot open %s.sai", prefix); return PBSC_ERR_IO; }
    if (have_bwt && stat(fmg.c_str(), &sb) == 0)
    {
        FmgHeader H;
        FILE* f = fopen(fmg.c_str(), "rb");
        const bool got = f && fread(&H, sizeof H, 1, f) == 1;
        if (f) fclose(f);
        bool match = got && check_header(H, (uint64_t)sb.st_size, fmg.c_str()) == PBSC_OK && H.k0 == k0;
        for (int w = 0; w < 2 && match; w++) match = H.n_strings[w] == ns[w] && H.n_symbols[w] == nsym[w] && H.src_runs[w] == nr[w];
        if (match)
        {
            const auto t0 = std::chrono::steady_clock::now();
            const int rc = pbsc_index_load_fmg(fmg.c_str(), device, out);
            if (getenv("PBSC_TRACE")) fprintf(stderr, "[pbsc] %s: %.2f GB in %.2f s\n", fmg.c_str(), H.total_bytes / 1e9, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            if (rc == PBSC_OK) { if (from_fmg) *from_fmg = 1; return rc; }
        }
        else if (getenv("PBSC_TRACE")) fprintf(stderr, "[pbsc] %s does not match the index files or k0 = %d: not used\n", fmg.c_str(), k0);
        // stale or damaged: fall through to the run-length files (and rewrite it below)
    }
    const auto t1 = std::chrono::steady_clock::now();
    int rc = pbsc_index_load(prefix, device, 0, out);
    if (rc != PBSC_OK) return rc;
    const auto t2 = std::chrono::steady_clock::now();
    if (k0 > 0) rc = pbsc_index_build_prefix_table(*out, k0);
    if (getenv("PBSC_TRACE")) fprintf(stderr, "[pbsc] %s.bwt/.rbwt read + decoded in %.2f s, prefix table in %.2f s\n", prefix, std::chrono::duration<double>(t2 - t1).count(),
                                      std::chrono::duration<double>(std::chrono::steady_clock::now() - t2).count());
    if (rc != PBSC_OK) { pbsc_index_destroy(*out); *out = nullptr; return rc; }
    if (write_fmg && pbsc_index_save(*out, fmg.c_str()) != PBSC_OK) fprintf(stderr, "[pbsc] warning: %s\n", pbsc_last_error());
    return PBSC_OK;
}
PBSC_CATCH_ALL("pbsc_index_open")

}  // extern "C"
