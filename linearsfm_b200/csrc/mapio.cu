// Host <-> device movement of local maps (one arena + one pinned staging buffer per batch).
//
// Replaces the heap-per-array ownership of LocalMapInfoStereo (LinearSFMImp.h:75-121): a batch of
// host maps is validated, packed by several host threads into ONE pinned buffer that mirrors the
// device arena byte for byte, and moved with ONE cudaMemcpyAsync.  The host-side conversions
// (stno -> poseNo/featNo, FBlock/feature -> CSR wPtr) happen during packing.
#include "mapio.h"
#include <cstring>
#include <atomic>
#include <thread>
#include <algorithm>

namespace {

struct Pinned {
    char *p = nullptr;
    size_t bytes = 0;
    char *get(size_t need)
    {
        if (need > bytes) {
            if (p) cudaFreeHost(p);
            size_t cap = std::max(need, bytes * 2);
            CUDA_CHECK(cudaHostAlloc((void **)&p, cap, cudaHostAllocDefault));
            bytes = cap;
        }
        return p;
    }
    ~Pinned() { if (p) cudaFreeHost(p); }
};
Pinned g_stage_up, g_stage_down;
cudaEvent_t g_up_done = nullptr;      // last upload's copies out of g_stage_up
bool g_up_pending = false;

// sizes only (needed before anything is allocated)
void validate_shape(const lsfm_map &M, int idx)
{
    auto bad = [&](const std::string &why) {
        throw LsfmError(LSFM_ERR_FORMAT, "local map " + std::to_string(idx) + ": " + why);
    };
    if (M.m < 0 || M.n < 0 || M.nU < 0 || M.nW < 0) bad("negative size");
    if (M.r != 6 * M.m + 3 * M.n) bad("r != 6m+3n");
}

// contents; runs on the packing threads
void validate_map(const lsfm_map &M, int idx)
{
    auto bad = [&](const std::string &why) {
        throw LsfmError(LSFM_ERR_FORMAT, "local map " + std::to_string(idx) + ": " + why);
    };
    for (int p = 0; p < M.m; p++)
        if (M.stno[6 * p] > 0) bad("pose row with positive stno");
    for (int f = 0; f < M.n; f++)
        if (M.stno[6 * M.m + 3 * f] <= 0) bad("feature row with non-positive stno");
    for (int b = 0; b < M.nU; b++)
        if (M.Ui[b] < 0 || M.Uj[b] >= M.m || M.Ui[b] > M.Uj[b]) bad("U block index out of order/range");
    int prev = 0;
    for (int j = 0; j < M.nW; j++) {
        if (M.photo[j] < 0 || M.photo[j] >= M.m) bad("photo index out of range");
        if (M.feature[j] < prev || M.feature[j] >= M.n) bad("feature index not sorted / out of range");
        prev = M.feature[j];
    }
}

} // namespace

std::vector<MapHandle> upload_maps(Context &ctx, const lsfm_map *maps, int K, bool validate)
{
    std::vector<DMap> shapes(K);
    for (int k = 0; k < K; k++) {
        if (validate) validate_shape(maps[k], k);
        DMap &d = shapes[k];
        memset(&d, 0, sizeof(d));
        d.Ref = maps[k].Ref; d.FRef = maps[k].FRef; d.m = maps[k].m; d.n = maps[k].n;
        d.nU = maps[k].nU; d.nW = maps[k].nW;
        d.ScaP = maps[k].ScaP; d.Fix = maps[k].Fix; d.Sign = maps[k].Sign;
        d.FScaP = maps[k].FScaP; d.FFix = maps[k].FFix;
    }
    std::vector<MapHandle> out = alloc_maps(ctx, shapes);
    if (K == 0) return out;
    Arena &A = *out[0].arena;
    if (g_up_pending) { CUDA_CHECK(cudaEventSynchronize(g_up_done)); g_up_pending = false; }
    char *stage = g_stage_up.get(A.used);
    char *base = A.base;
    auto H = [&](const void *devp) { return stage + ((const char *)devp - base); };

    // Packing and the host-to-device copy are pipelined: the maps are cut into groups of ~8 MB
    // (contiguous in the arena); worker threads pack groups in order, the calling thread issues one
    // asynchronous copy per group as soon as it is packed, so the PCIe transfer of group g overlaps
    // the packing of the groups behind it.
    int nthreads = std::min<int>(std::max(1u, std::thread::hardware_concurrency()), 16);
    if (A.used < (1u << 22)) nthreads = 1;
    auto map_begin = [&](int k) { return (size_t)((const char *)out[k].d.poseNo - base); };
    std::vector<int> gstart{0};
    for (int k = 1; k < K; k++)
        if (map_begin(k) - map_begin(gstart.back()) >= (size_t)(8u << 20)) gstart.push_back(k);
    const int G = (int)gstart.size();
    gstart.push_back(K);
    auto pack_map = [&](int k) {
        const lsfm_map &M = maps[k];
        if (validate) validate_map(M, k);
        const DMap &d = out[k].d;
        int *poseNo = (int *)H(d.poseNo);
        double *poseVal = (double *)H(d.poseVal);
        for (int p = 0; p < M.m; p++) poseNo[p] = M.stno[6 * p];
        if (M.m) memcpy(poseVal, M.stVal, sizeof(double) * 6 * (size_t)M.m);
        int *featNo = (int *)H(d.featNo);
        for (int f = 0; f < M.n; f++) featNo[f] = M.stno[6 * M.m + 3 * f];
        if (M.n) memcpy(H(d.featVal), M.stVal + 6 * (size_t)M.m, sizeof(double) * 3 * (size_t)M.n);
        if (M.nU) {
            memcpy(H(d.U), M.U, sizeof(double) * 36 * (size_t)M.nU);
            memcpy(H(d.Ui), M.Ui, sizeof(int) * M.nU);
            memcpy(H(d.Uj), M.Uj, sizeof(int) * M.nU);
        }
        if (M.nW) {
            memcpy(H(d.W), M.W, sizeof(double) * 18 * (size_t)M.nW);
            memcpy(H(d.photo), M.photo, sizeof(int) * M.nW);
            memcpy(H(d.feature), M.feature, sizeof(int) * M.nW);
        }
        if (M.n) memcpy(H(d.V), M.V, sizeof(double) * 9 * (size_t)M.n);
        int *wPtr = (int *)H(d.wPtr);
        int j = 0;
        for (int f = 0; f < M.n; f++) {
            wPtr[f] = j;
            while (j < M.nW && M.feature[j] == f) j++;
            if (validate && M.FBlock) {
                int fb = (j > wPtr[f]) ? wPtr[f] : -1;
                if (M.FBlock[f] != fb)
                    throw LsfmError(LSFM_ERR_FORMAT, "local map " + std::to_string(k) +
                                                         ": FBlock inconsistent with feature[]");
            }
        }
        wPtr[M.n] = M.nW;
    };
    std::vector<std::string> errs(nthreads);
    if (nthreads == 1) {
        try { for (int k = 0; k < K; k++) pack_map(k); } catch (const std::exception &e) { errs[0] = e.what(); }
        if (errs[0].empty()) CUDA_CHECK(cudaMemcpyAsync(A.base, stage, A.used, cudaMemcpyHostToDevice, ctx.stream));
    } else {
        std::atomic<int> next(0);
        std::vector<std::atomic<int>> done(G);
        for (auto &d : done) d.store(0);
        auto work = [&](int tid) {
            for (;;) {
                int g = next.fetch_add(1);
                if (g >= G) break;
                try {
                    for (int k = gstart[g]; k < gstart[g + 1]; k++) pack_map(k);
                } catch (const std::exception &e) { errs[tid] = e.what(); }
                done[g].store(1, std::memory_order_release);
            }
        };
        // the packing threads are joined on EVERY exit path (a throwing CUDA call below must surface as
        // an error code, not as std::terminate on a joinable thread)
        struct Joiner {
            std::vector<std::thread> th;
            ~Joiner() { for (auto &t : th) if (t.joinable()) t.join(); }
        } joiner;
        std::vector<std::thread> &th = joiner.th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
        bool failed = false;
        for (int g = 0; g < G; g++) {
            while (!done[g].load(std::memory_order_acquire)) std::this_thread::yield();
            for (auto &e : errs) if (!e.empty()) failed = true;
            if (failed) continue;
            size_t b0 = map_begin(gstart[g]), b1 = (g + 1 < G) ? map_begin(gstart[g + 1]) : A.used;
            CUDA_CHECK(cudaMemcpyAsync(A.base + b0, stage + b0, b1 - b0, cudaMemcpyHostToDevice, ctx.stream));
        }
        for (auto &t : th) t.join();
    }
    for (auto &e : errs) if (!e.empty()) throw LsfmError(LSFM_ERR_FORMAT, e);
    // the staging buffer is reused by the next upload: that one waits for this copy (the caller's
    // own buffers have been read completely by now, and the stream orders the copy before any kernel)
    if (!g_up_done) CUDA_CHECK(cudaEventCreateWithFlags(&g_up_done, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(g_up_done, ctx.stream));
    g_up_pending = true;
    return out;
}

void mapio_shutdown()
{
    if (g_up_pending && g_up_done) cudaEventSynchronize(g_up_done);
    g_up_pending = false;
    if (g_up_done) cudaEventDestroy(g_up_done);
    g_up_done = nullptr;
}

static void alloc_host_map(lsfm_map *o, const DMap &d)
{
    memset(o, 0, sizeof(*o));
    o->Ref = d.Ref; o->FRef = d.FRef; o->m = d.m; o->n = d.n; o->nU = d.nU; o->nW = d.nW;
    o->r = 6 * d.m + 3 * d.n;
    o->ScaP = d.ScaP; o->Fix = d.Fix; o->Sign = d.Sign; o->FScaP = d.FScaP; o->FFix = d.FFix;
    auto A = [](size_t n, size_t sz) { return malloc((n * sz) != 0 ? (n * sz) : 1); };
    o->stno = (int *)A(o->r, sizeof(int));
    o->stVal = (double *)A(o->r, sizeof(double));
    o->U = (double *)A(36 * (size_t)d.nU, sizeof(double));
    o->Ui = (int *)A(d.nU, sizeof(int));
    o->Uj = (int *)A(d.nU, sizeof(int));
    o->W = (double *)A(18 * (size_t)d.nW, sizeof(double));
    o->photo = (int *)A(d.nW, sizeof(int));
    o->feature = (int *)A(d.nW, sizeof(int));
    o->V = (double *)A(9 * (size_t)d.n, sizeof(double));
    o->FBlock = (int *)A(d.n, sizeof(int));
}

void download_state(Context &ctx, const MapHandle &h, int *stno, double *stVal)
{
    const DMap &d = h.d;
    size_t need = sizeof(int) * (d.m + d.n) + 256;
    char *stage = g_stage_down.get(need);
    int *poseNo = (int *)stage, *featNo = poseNo + d.m;
    cudaStream_t s = ctx.stream;
    if (d.m) {
        CUDA_CHECK(cudaMemcpyAsync(poseNo, d.poseNo, sizeof(int) * d.m, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaMemcpyAsync(stVal, d.poseVal, sizeof(double) * 6 * (size_t)d.m, cudaMemcpyDeviceToHost, s));
    }
    if (d.n) {
        CUDA_CHECK(cudaMemcpyAsync(featNo, d.featNo, sizeof(int) * d.n, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaMemcpyAsync(stVal + 6 * (size_t)d.m, d.featVal, sizeof(double) * 3 * (size_t)d.n,
                                   cudaMemcpyDeviceToHost, s));
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (int p = 0; p < d.m; p++)
        for (int q = 0; q < 6; q++) stno[6 * p + q] = poseNo[p];
    int *sf = stno + 6 * (size_t)d.m;
    for (int f = 0; f < d.n; f++) { sf[3 * f] = sf[3 * f + 1] = sf[3 * f + 2] = featNo[f]; }
}

void download_map(Context &ctx, const MapHandle &h, lsfm_map *o)
{
    const DMap &d = h.d;
    alloc_host_map(o, d);
    cudaStream_t s = ctx.stream;
    auto D2H = [&](void *dst, const void *src, size_t b) {
        if (b) CUDA_CHECK(cudaMemcpyAsync(dst, src, b, cudaMemcpyDeviceToHost, s));
    };
    std::vector<int> wPtr(d.n + 1);
    D2H(o->U, d.U, sizeof(double) * 36 * (size_t)d.nU);
    D2H(o->Ui, d.Ui, sizeof(int) * d.nU);
    D2H(o->Uj, d.Uj, sizeof(int) * d.nU);
    D2H(o->W, d.W, sizeof(double) * 18 * (size_t)d.nW);
    D2H(o->photo, d.photo, sizeof(int) * d.nW);
    D2H(o->feature, d.feature, sizeof(int) * d.nW);
    D2H(o->V, d.V, sizeof(double) * 9 * (size_t)d.n);
    D2H(wPtr.data(), d.wPtr, sizeof(int) * (d.n + 1));
    download_state(ctx, h, o->stno, o->stVal);      // synchronises
    for (int f = 0; f < d.n; f++) o->FBlock[f] = (wPtr[f + 1] > wPtr[f]) ? wPtr[f] : -1;
}

void free_host_map(lsfm_map *m)
{
    if (!m) return;
    free(m->stno); free(m->stVal); free(m->U); free(m->Ui); free(m->Uj); free(m->W);
    free(m->photo); free(m->feature); free(m->V); free(m->FBlock);
    memset(m, 0, sizeof(*m));
}
