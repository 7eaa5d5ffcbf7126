// C ABI of liblinearsfm_b200.so (declared in include/linearsfm_b200.h).
#include "../../include/linearsfm_b200.h"
#include "mapio.h"
#include "scheduler.h"
#include "chol_symbolic.h"
#include "builder.h"
#include <cstring>
#include <sstream>
#include <algorithm>
#include <mutex>

namespace {

Context *g_ctx = nullptr;
std::string g_err;
std::string g_stats_json;
SolveDebug g_dbg;
bool g_dbg_valid = false;

int fail(int code, const std::string &msg) { g_err = msg; return code; }
} // namespace
// the one error string lsfm_last_error() returns; file I/O (fileio.cpp) reports through it too
void lsfm_set_error(const std::string &msg) { g_err = msg; }
// the process's context for library-internal tools (pcg.cu); nullptr without a device
Context *lsfm_internal_ctx() { return (g_ctx || lsfm_init(0) == LSFM_OK) ? g_ctx : nullptr; }
namespace {

int ensure_ctx()
{
    if (g_ctx) return LSFM_OK;
    return lsfm_init(0);
}

template <class F> int guarded(F &&f)
{
    try {
        int rc = ensure_ctx();
        if (rc != LSFM_OK) return rc;
        f();
        return LSFM_OK;
    } catch (const LsfmError &e) {
        return fail(e.code, e.what());
    } catch (const std::exception &e) {
        return fail(LSFM_ERR_CUDA, e.what());
    }
}

} // namespace

struct lsfm_tree {
    std::vector<MapHandle> uploaded;   // what the last create/set_maps put in HBM
    std::vector<MapHandle> leaves;     // current input set of lsfm_tree_solve
    std::vector<MapHandle> result;
    double last_ms = 0.0;            // device time of the last solve (CUDA events on our stream)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

extern "C" {

int lsfm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int lsfm_init(int device)
{
    try {
        int n = lsfm_device_count();
        if (n <= 0)
            return fail(LSFM_ERR_NO_DEVICE, "no CUDA device visible: the LinearSFM hot path has no CPU fallback");
        if (device < 0 || device >= n) return fail(LSFM_ERR_ARG, "bad device index");
        if (g_ctx && g_ctx->device == device) return LSFM_OK;
        if (g_ctx) lsfm_shutdown();
        CUDA_CHECK(cudaSetDevice(device));
        DevicePool::get().set_device(device);
        Context *c = new Context();
        c->device = device;
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreate(&c->ev0));
        CUDA_CHECK(cudaEventCreate(&c->ev1));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        c->num_sms = prop.multiProcessorCount;
        // keep freed blocks in the stream-ordered pool: allocation becomes a pointer bump
        cudaMemPool_t pool;
        CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long thr = ~0ull;
        CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        g_ctx = c;
        return LSFM_OK;
    } catch (const LsfmError &e) {
        return fail(e.code, e.what());
    }
}

void lsfm_shutdown(void)
{
    if (!g_ctx) return;
    cudaStreamSynchronize(g_ctx->stream);
    mapio_shutdown();                       // upload event / staging state belong to this device
    g_ctx->release_error_flag();
    DevicePool::get().release_cached();     // cached blocks of this device go back to the driver
    cudaEventDestroy(g_ctx->ev0);
    cudaEventDestroy(g_ctx->ev1);
    cudaStreamDestroy(g_ctx->stream);
    delete g_ctx;
    g_ctx = nullptr;
}

const char *lsfm_last_error(void) { return g_err.c_str(); }

void lsfm_free_map(lsfm_map *m) { free_host_map(m); }

void lsfm_stats_reset(int enable_stage_timing)
{
    if (ensure_ctx() != LSFM_OK) return;
    g_ctx->stats.clear();
    g_ctx->launches = 0;
    g_ctx->objectives.clear();
    g_ctx->timing = (enable_stage_timing & 1) != 0;
    g_ctx->want_objective = (enable_stage_timing & 2) != 0;
}

const char *lsfm_stats_json(void)
{
    std::ostringstream os;
    os.precision(17);
    os << "{\"launches\": " << (g_ctx ? g_ctx->launches : 0) << ", \"stages\": {";
    bool first = true;
    if (g_ctx)
        for (auto &kv : g_ctx->stats) {
            if (!first) os << ", ";
            first = false;
            os << "\"" << kv.first << "\": {\"ms\": " << kv.second.ms << ", \"launches\": " << kv.second.launches
               << ", \"bytes\": " << kv.second.bytes << ", \"flops\": " << kv.second.flops << "}";
        }
    os << "}, \"objectives\": [";
    if (g_ctx)
        for (size_t i = 0; i < g_ctx->objectives.size(); i++) os << (i ? ", " : "") << g_ctx->objectives[i];
    os << "]}";
    g_stats_json = os.str();
    return g_stats_json.c_str();
}

int lsfm_transform_stereo_batch(const lsfm_map *in, const int *Ref, int K, lsfm_map *out)
{
    return guarded([&] {
        std::vector<MapHandle> h = upload_maps(*g_ctx, in, K, true);
        std::vector<int> refs(Ref, Ref + K);
        std::vector<MapHandle> r = transform_or_pass(*g_ctx, h, refs);
        for (int k = 0; k < K; k++) download_map(*g_ctx, r[k], &out[k]);
    });
}

int lsfm_transform_stereo(const lsfm_map *in, int Ref, lsfm_map *out)
{
    return lsfm_transform_stereo_batch(in, &Ref, 1, out);
}

int lsfm_join_stereo_batch(const lsfm_map *end, const lsfm_map *cur, int K, lsfm_map *out)
{
    return guarded([&] {
        std::vector<MapHandle> E = upload_maps(*g_ctx, end, K, true);
        std::vector<MapHandle> C = upload_maps(*g_ctx, cur, K, true);
        std::vector<MapHandle> J = join_stereo_batch(*g_ctx, E, C);
        g_ctx->check_errors();
        for (int k = 0; k < K; k++) download_map(*g_ctx, J[k], &out[k]);
    });
}

int lsfm_transform_mono_batch(const lsfm_map *in, const int *Ref, const int *ScaP, const int *Fix, int K,
                              lsfm_map *out)
{
    return guarded([&] {
        std::vector<MapHandle> h = upload_maps(*g_ctx, in, K, true);
        std::vector<MapHandle> todo;
        std::vector<int> r, sc, fx, where;
        for (int k = 0; k < K; k++)
            if (!(in[k].Ref == Ref[k] && in[k].ScaP == ScaP[k])) {          // LinearSFMImp.cpp:3176
                todo.push_back(h[k]); r.push_back(Ref[k]); sc.push_back(ScaP[k]); fx.push_back(Fix[k]);
                where.push_back(k);
            }
        std::vector<MapHandle> done = transform_mono_batch(*g_ctx, todo, r, sc, fx);
        for (size_t j = 0; j < where.size(); j++) h[where[j]] = done[j];
        for (int k = 0; k < K; k++) download_map(*g_ctx, h[k], &out[k]);
    });
}

int lsfm_transform_mono(const lsfm_map *in, int Ref, int ScaP, int Fix, lsfm_map *out)
{
    return lsfm_transform_mono_batch(in, &Ref, &ScaP, &Fix, 1, out);
}

int lsfm_join_mono_batch(const lsfm_map *end, const lsfm_map *cur, int K, lsfm_map *out)
{
    return guarded([&] {
        std::vector<MapHandle> E = upload_maps(*g_ctx, end, K, true);
        std::vector<MapHandle> C = upload_maps(*g_ctx, cur, K, true);
        std::vector<MapHandle> J = join_mono_batch(*g_ctx, E, C);
        g_ctx->check_errors();
        for (int k = 0; k < K; k++) download_map(*g_ctx, J[k], &out[k]);
    });
}

int lsfm_join_mono(const lsfm_map *end, const lsfm_map *cur, lsfm_map *out)
{
    return lsfm_join_mono_batch(end, cur, 1, out);
}

int lsfm_run_mono(const lsfm_map *maps, int num, lsfm_map *out) { return lsfm_run_mono_ex(maps, num, 0, out); }

int lsfm_run_mono_ex(const lsfm_map *maps, int num, int verbose, lsfm_map *out)
{
    return guarded([&] {
        if (num < 1) throw LsfmError(LSFM_ERR_ARG, "need at least one local map");
        std::vector<MapHandle> leaves = upload_maps(*g_ctx, maps, num, true);
        std::vector<MapHandle> top = solve_tree_mono(*g_ctx, std::move(leaves), verbose != 0);
        g_ctx->check_errors();
        download_map(*g_ctx, top[0], out);
    });
}

int lsfm_join_stereo(const lsfm_map *end, const lsfm_map *cur, lsfm_map *out)
{
    return lsfm_join_stereo_batch(end, cur, 1, out);
}

int lsfm_solve_stereo(double *stVal, const double *eb, const double *ea, const double *U,
                      const double *W, const double *V, const int *Ui, const int *Uj,
                      const int *photo, const int *feature, int m, int n, int nU, int nW)
{
    return guarded([&] {
        if (m <= 0 || n < 0) throw LsfmError(LSFM_ERR_ARG, "solve: bad sizes");
        // wrap the raw arrays into a map (state values are outputs)
        lsfm_map M;
        memset(&M, 0, sizeof(M));
        M.m = m; M.n = n; M.nU = nU; M.nW = nW; M.r = 6 * m + 3 * n;
        std::vector<int> stno(M.r);
        for (int i = 0; i < 6 * m; i++) stno[i] = -(i / 6 + 1);
        for (int i = 0; i < 3 * n; i++) stno[6 * m + i] = i / 3 + 1;
        std::vector<double> zeros(M.r, 0.0);
        M.stno = stno.data(); M.stVal = zeros.data();
        M.U = (double *)U; M.Ui = (int *)Ui; M.Uj = (int *)Uj;
        M.W = (double *)W; M.photo = (int *)photo; M.feature = (int *)feature;
        M.V = (double *)V; M.FBlock = nullptr;
        std::vector<MapHandle> h = upload_maps(*g_ctx, &M, 1, true);
        OpMaps J;
        J.build(h, g_ctx->stream);
        DevBuf<double> eP(6 * (size_t)m, g_ctx->stream), eF(3 * (size_t)n, g_ctx->stream);
        eP.upload(ea, 6 * (size_t)m);
        if (n) eF.upload(eb, 3 * (size_t)n);
        g_dbg = SolveDebug();
        solve_stereo_batch(*g_ctx, J, eP.p, eF.p, &g_dbg);
        g_ctx->check_errors();
        g_dbg_valid = true;
        std::vector<int> tmpno(M.r);
        download_state(*g_ctx, h[0], tmpno.data(), stVal);
    });
}

// one-element gauge arrays of the standalone mono solve operator
static __global__ void k_set_gauge(int *g, int ref, int fix, int sign) { g[0] = ref; g[1] = fix; g[2] = sign; }

int lsfm_solve_mono(double *stVal, const double *eb, const double *ea, const double *U, const double *W,
                    const double *V, const int *Ui, const int *Uj, const int *photo, const int *feature,
                    int m, int n, int nU, int nW, int Ref, int ScaP, int Fix, int Sign, int FixBlk)
{
    return guarded([&] {
        if (m <= 0 || n < 0) throw LsfmError(LSFM_ERR_ARG, "solve: bad sizes");
        if (Ref < 0 || Ref >= m || ScaP != 6 * Ref || Fix < 0 || Fix >= 6 * m)
            throw LsfmError(LSFM_ERR_ARG, "solve_mono: gauge arguments out of range (ScaP must be 6 * Ref)");
        (void)FixBlk;                     // only steers the reference's scalar permutation (7083-7105)
        lsfm_map M;
        memset(&M, 0, sizeof(M));
        M.m = m; M.n = n; M.nU = nU; M.nW = nW; M.r = 6 * m + 3 * n;
        std::vector<int> stno(M.r);
        for (int i = 0; i < 6 * m; i++) stno[i] = -(i / 6 + 1);
        for (int i = 0; i < 3 * n; i++) stno[6 * m + i] = i / 3 + 1;
        std::vector<double> zeros(M.r, 0.0);
        M.stno = stno.data(); M.stVal = zeros.data();
        M.U = (double *)U; M.Ui = (int *)Ui; M.Uj = (int *)Uj;
        M.W = (double *)W; M.photo = (int *)photo; M.feature = (int *)feature;
        M.V = (double *)V; M.FBlock = nullptr;
        std::vector<MapHandle> h = upload_maps(*g_ctx, &M, 1, true);
        OpMaps J;
        J.build(h, g_ctx->stream);
        DevBuf<double> eP(6 * (size_t)m, g_ctx->stream), eF(3 * (size_t)std::max(n, 1), g_ctx->stream);
        eP.upload(ea, 6 * (size_t)m);
        if (n) eF.upload(eb, 3 * (size_t)n);
        DevBuf<int> g(3, g_ctx->stream);
        k_set_gauge<<<1, 1, 0, g_ctx->stream>>>(g.p, Ref, Fix, Sign);
        MonoGauge gauge{g.p, g.p + 1, g.p + 2};
        g_dbg = SolveDebug();
        solve_stereo_batch(*g_ctx, J, eP.p, eF.p, &g_dbg, &gauge);
        g_ctx->check_errors();
        g_dbg_valid = true;
        std::vector<int> tmpno(M.r);
        download_state(*g_ctx, h[0], tmpno.data(), stVal);
    });
}

int lsfm_debug_last_solve(int *m, const int **rowptr, const int **colidx, const double **S,
                          const double **E, const int **perm)
{
    if (!g_dbg_valid) return fail(LSFM_ERR_ARG, "no solve captured");
    if (m) *m = (int)g_dbg.rowptr.size() - 1;
    if (rowptr) *rowptr = g_dbg.rowptr.data();
    if (colidx) *colidx = g_dbg.colidx.data();
    if (S) *S = g_dbg.S.data();
    if (E) *E = g_dbg.E.data();
    if (perm) *perm = g_dbg.perm.data();
    return LSFM_OK;
}

int lsfm_block_ordering(int m, const int *Ap, const int *Ai, int *perm)
{
    // host-only integer routine (no device needed): upper CSC block pattern -> LSFM-ND ordering
    try {
        std::vector<int> ptr(m + 1, 0);
        for (int j = 0; j < m; j++)
            for (int p = Ap[j]; p < Ap[j + 1]; p++)
                if (Ai[p] != j) { ptr[Ai[p] + 1]++; ptr[j + 1]++; }
        for (int i = 0; i < m; i++) ptr[i + 1] += ptr[i];
        std::vector<int> adj(ptr[m]), fill(ptr.begin(), ptr.end() - 1);
        for (int j = 0; j < m; j++)
            for (int p = Ap[j]; p < Ap[j + 1]; p++)
                if (Ai[p] != j) { adj[fill[Ai[p]]++] = j; adj[fill[j]++] = Ai[p]; }
        for (int v = 0; v < m; v++) std::sort(adj.begin() + ptr[v], adj.begin() + ptr[v + 1]);
        std::vector<int> p, nodes;
        lsfm_order(m, ptr.data(), adj.data(), p, nodes);
        for (int i = 0; i < m; i++) perm[i] = p[i];
        return LSFM_OK;
    } catch (const std::exception &e) {
        return fail(LSFM_ERR_ARG, e.what());
    }
}

int lsfm_build_localmaps_stereo(const lsfm_stereo_pair *pairs, int num, const lsfm_stereo_cam *cam,
                                int max_iters, double tol, lsfm_map *out, int *iters_done)
{
    return guarded([&] {
        if (!pairs || !cam || !out || num < 0) throw LsfmError(LSFM_ERR_ARG, "builder: null argument");
        build_localmaps_stereo(*g_ctx, pairs, num, *cam, max_iters, tol, out, iters_done);
    });
}

int lsfm_run_stereo(const lsfm_map *maps, int num, lsfm_map *out)
{
    return guarded([&] {
        if (num < 1) throw LsfmError(LSFM_ERR_ARG, "need at least one local map");
        std::vector<MapHandle> leaves = upload_maps(*g_ctx, maps, num, true);
        std::vector<MapHandle> top = solve_tree_stereo(*g_ctx, std::move(leaves), false, 0, -1);
        MapHandle root = final_rebase_stereo(*g_ctx, top[0]);
        g_ctx->check_errors();
        download_map(*g_ctx, root, out);
    });
}

int lsfm_tree_create_stereo(const lsfm_map *maps, int num, lsfm_tree **tree)
{
    return guarded([&] {
        lsfm_tree *t = new lsfm_tree();
        t->uploaded = upload_maps(*g_ctx, maps, num, true);
        t->leaves = t->uploaded;
        *tree = t;
    });
}

int lsfm_tree_set_maps(lsfm_tree *tree, const lsfm_map *maps, int num)
{
    return guarded([&] {
        tree->result.clear();
        tree->leaves.clear();
        tree->uploaded.clear();
        tree->uploaded = upload_maps(*g_ctx, maps, num, true);
        tree->leaves = tree->uploaded;
    });
}

static DMap shape_of(const lsfm_map *s)
{
    DMap d;
    memset(&d, 0, sizeof(d));
    d.Ref = s->Ref; d.FRef = s->FRef; d.m = s->m; d.n = s->n; d.nU = s->nU; d.nW = s->nW;
    return d;
}

size_t lsfm_map_device_bytes(const lsfm_map *shape)
{
    DMap d = shape_of(shape);
    return map_layout(d, nullptr);
}

int lsfm_tree_export_device(const lsfm_tree *tree, int idx, void *dst_device, size_t bytes)
{
    return guarded([&] {
        if (idx < 0 || idx >= (int)tree->result.size()) throw LsfmError(LSFM_ERR_ARG, "bad result index");
        const DMap &s = tree->result[idx].d;
        DMap d = s;
        size_t need = map_layout(d, (char *)dst_device);
        if (bytes < need) throw LsfmError(LSFM_ERR_ARG, "export buffer too small");
        cudaStream_t st = g_ctx->stream;
        auto cp = [&](void *dst, const void *src, size_t b) {
            if (b) CUDA_CHECK(cudaMemcpyAsync(dst, src, b, cudaMemcpyDeviceToDevice, st));
        };
        cp(d.poseNo, s.poseNo, sizeof(int) * s.m);
        cp(d.poseVal, s.poseVal, sizeof(double) * 6 * (size_t)s.m);
        cp(d.featNo, s.featNo, sizeof(int) * s.n);
        cp(d.featVal, s.featVal, sizeof(double) * 3 * (size_t)s.n);
        cp(d.U, s.U, sizeof(double) * 36 * (size_t)s.nU);
        cp(d.Ui, s.Ui, sizeof(int) * s.nU);
        cp(d.Uj, s.Uj, sizeof(int) * s.nU);
        cp(d.W, s.W, sizeof(double) * 18 * (size_t)s.nW);
        cp(d.photo, s.photo, sizeof(int) * s.nW);
        cp(d.feature, s.feature, sizeof(int) * s.nW);
        cp(d.V, s.V, sizeof(double) * 9 * (size_t)s.n);
        cp(d.wPtr, s.wPtr, sizeof(int) * (s.n + 1));
        CUDA_CHECK(cudaStreamSynchronize(st));     // the caller hands the buffer to NCCL on another stream
    });
}

int lsfm_tree_append_device(lsfm_tree *tree, const lsfm_map *shape, const void *src_device, size_t bytes)
{
    return guarded([&] {
        std::vector<DMap> shapes{shape_of(shape)};
        std::vector<MapHandle> h = alloc_maps(*g_ctx, shapes);
        DMap probe = shapes[0];
        size_t need = map_layout(probe, nullptr);
        if (bytes < need) throw LsfmError(LSFM_ERR_ARG, "import buffer too small");
        CUDA_CHECK(cudaMemcpyAsync(h[0].arena->base, src_device, need, cudaMemcpyDeviceToDevice, g_ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(g_ctx->stream));
        tree->leaves.push_back(h[0]);
    });
}

int lsfm_tree_reset(lsfm_tree *tree)
{
    tree->result.clear();
    tree->leaves = tree->uploaded;
    return LSFM_OK;
}

int lsfm_tree_solve(lsfm_tree *tree, int verbose, int first_index, int max_levels)
{
    return guarded([&] {
        tree->result.clear();
        if (!tree->e0) { CUDA_CHECK(cudaEventCreate(&tree->e0)); CUDA_CHECK(cudaEventCreate(&tree->e1)); }
        CUDA_CHECK(cudaEventRecord(tree->e0, g_ctx->stream));
        std::vector<MapHandle> top = solve_tree_stereo(*g_ctx, tree->leaves, verbose != 0, first_index, max_levels);
        if (max_levels < 0 && top.size() == 1) top[0] = final_rebase_stereo(*g_ctx, top[0]);
        tree->result = std::move(top);
        CUDA_CHECK(cudaEventRecord(tree->e1, g_ctx->stream));
        CUDA_CHECK(cudaEventSynchronize(tree->e1));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, tree->e0, tree->e1));
        tree->last_ms = ms;
        g_ctx->check_errors();
    });
}

double lsfm_tree_last_solve_ms(const lsfm_tree *tree) { return tree->last_ms; }

int lsfm_tree_adopt_result(lsfm_tree *tree)
{
    tree->leaves = std::move(tree->result);
    tree->result.clear();
    return LSFM_OK;
}

int lsfm_tree_append_maps(lsfm_tree *tree, const lsfm_map *maps, int num)
{
    return guarded([&] {
        std::vector<MapHandle> h = upload_maps(*g_ctx, maps, num, true);
        tree->leaves.insert(tree->leaves.end(), h.begin(), h.end());
    });
}

int lsfm_tree_result_count(const lsfm_tree *tree) { return (int)tree->result.size(); }

int lsfm_tree_result_shape(const lsfm_tree *tree, int idx, lsfm_map *o)
{
    if (idx < 0 || idx >= (int)tree->result.size()) return fail(LSFM_ERR_ARG, "bad result index");
    const DMap &d = tree->result[idx].d;
    memset(o, 0, sizeof(*o));
    o->Ref = d.Ref; o->FRef = d.FRef; o->m = d.m; o->n = d.n; o->nU = d.nU; o->nW = d.nW;
    o->r = 6 * d.m + 3 * d.n;
    return LSFM_OK;
}

int lsfm_tree_download(const lsfm_tree *tree, int idx, lsfm_map *out)
{
    return guarded([&] {
        if (idx < 0 || idx >= (int)tree->result.size()) throw LsfmError(LSFM_ERR_ARG, "bad result index");
        download_map(*g_ctx, tree->result[idx], out);
    });
}

// ---------------------------------------------------------------------------------------------
// Marginal covariances of selected state blocks (the reference frees the final
// information matrix without using it, LinearSFMImp.cpp:2081-2096).  Sigma = I^-1 of the map's block
// information matrix I = [[U, W], [W^T, V]]: column c of Sigma is the solution of I x = e_c, so the
// d (6 or 3) unit right-hand sides of a selected block go through the joint solver of the merge tree
// (per-feature 3x3 inversion, Schur complement, multifrontal Cholesky, back-substitution) as ONE
// segmented batch whose K entries share the map's U / W / V arrays and differ in the right-hand
// side and in the state buffers the solutions are written to.
// ---------------------------------------------------------------------------------------------
static __global__ void k_unit_rhs(double *eP, double *eF, const long long *pos, int K)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const long long q = pos[k];                    // >= 0: index into eP, < 0: -(index into eF) - 1
    if (q >= 0) eP[q] = 1.0; else eF[-q - 1] = 1.0;
}

static __global__ void k_cov_gather(const double *xP, const double *xF, int m, int n, const int *colBlk,
                                    const int *colOff, int K, double *out)
{
    // column k (unit vector `colOff[k]` of block colBlk[k]) -> the d entries of the block's own rows
    int k = blockIdx.x, r = threadIdx.x;
    if (k >= K) return;
    const int b = colBlk[k];
    const int d = b < m ? 6 : 3;
    if (r >= d) return;
    const double v = b < m ? xP[6 * (size_t)m * k + 6 * (size_t)b + r] : xF[3 * (size_t)n * k + 3 * (size_t)(b - m) + r];
    out[(size_t)8 * k + r] = v;                    // 8 doubles per column (padded)
}

static void marginal_cov(Context &ctx, const MapHandle &h, int nsel, const int *sel, double *cov)
{
    const DMap &d = h.d;
    const int m = d.m, n = d.n;
    if (nsel < 0 || (nsel > 0 && (!sel || !cov))) throw LsfmError(LSFM_ERR_ARG, "marginal_cov: bad arguments");
    struct Col { int blk, comp; size_t out; };
    std::vector<Col> cols;
    size_t off = 0;
    for (int i = 0; i < nsel; i++) {
        if (sel[i] < 0 || sel[i] >= m + n) throw LsfmError(LSFM_ERR_ARG, "marginal_cov: block index out of range");
        const int dd = sel[i] < m ? 6 : 3;
        for (int c = 0; c < dd; c++) cols.push_back({sel[i], c, off});
        off += (size_t)dd * dd;
    }
    cudaStream_t s = ctx.stream;
    // columns per batch: bounded by the K state buffers (8 (6 m + 3 n) K bytes) and by K copies of S
    const size_t perCol = sizeof(double) * (6 * (size_t)m + 3 * (size_t)n) * 2;
    const int KB = (int)std::max<size_t>(1, std::min<size_t>(48, ((size_t)1 << 30) / std::max<size_t>(perCol, 1)));
    for (size_t c0 = 0; c0 < cols.size(); c0 += KB) {
        const int K = (int)std::min<size_t>(KB, cols.size() - c0);
        DevBuf<double> xP(6 * (size_t)m * K, s), xF(3 * (size_t)std::max(n, 1) * K, s);
        DevBuf<double> eP(6 * (size_t)m * K, s), eF(3 * (size_t)std::max(n, 1) * K, s);
        xP.zero(); xF.zero(); eP.zero(); eF.zero();
        std::vector<DMap> maps(K, d);
        std::vector<long long> pos(K);
        std::vector<int> colBlk(K), colOff(K);
        for (int k = 0; k < K; k++) {
            const Col &c = cols[c0 + k];
            maps[k].poseVal = xP.p + 6 * (size_t)m * k;
            maps[k].featVal = xF.p + 3 * (size_t)n * k;
            colBlk[k] = c.blk; colOff[k] = c.comp;
            pos[k] = c.blk < m ? (long long)(6 * (size_t)m * k + 6 * (size_t)c.blk + c.comp)
                               : -(long long)(3 * (size_t)n * k + 3 * (size_t)(c.blk - m) + c.comp) - 1;
        }
        OpMaps J;
        J.build(maps, s);
        DevBuf<long long> dPos(K, s);
        DevBuf<int> dBlk(K, s), dOff(K, s);
        dPos.upload(pos); dBlk.upload(colBlk); dOff.upload(colOff);
        k_unit_rhs<<<ceil_div(K, 64), 64, 0, s>>>(eP.p, eF.p, dPos.p, K);
        solve_stereo_batch(ctx, J, eP.p, eF.p, nullptr);
        ctx.check_errors();
        DevBuf<double> outD(8 * (size_t)K, s);
        k_cov_gather<<<K, 8, 0, s>>>(xP.p, xF.p, m, n, dBlk.p, dOff.p, K, outD.p);
        std::vector<double> outH = outD.to_host();
        for (int k = 0; k < K; k++) {
            const Col &c = cols[c0 + k];
            const int dd = c.blk < m ? 6 : 3;
            for (int r = 0; r < dd; r++) cov[c.out + (size_t)r * dd + c.comp] = outH[8 * (size_t)k + r];
        }
    }
}

int lsfm_marginal_cov_stereo(const lsfm_map *map, int nsel, const int *sel, double *cov)
{
    return guarded([&] {
        if (!map) throw LsfmError(LSFM_ERR_ARG, "marginal_cov: null map");
        std::vector<MapHandle> h = upload_maps(*g_ctx, map, 1, true);
        marginal_cov(*g_ctx, h[0], nsel, sel, cov);
    });
}

int lsfm_tree_marginal_cov(const lsfm_tree *tree, int idx, int nsel, const int *sel, double *cov)
{
    return guarded([&] {
        if (!tree || idx < 0 || idx >= (int)tree->result.size()) throw LsfmError(LSFM_ERR_ARG, "bad result index");
        marginal_cov(*g_ctx, tree->result[idx], nsel, sel, cov);
    });
}

int lsfm_tree_download_state(const lsfm_tree *tree, int idx, int *stno, double *stVal)
{
    return guarded([&] {
        if (idx < 0 || idx >= (int)tree->result.size()) throw LsfmError(LSFM_ERR_ARG, "bad result index");
        download_state(*g_ctx, tree->result[idx], stno, stVal);
    });
}

void lsfm_tree_free(lsfm_tree *tree)
{
    if (!tree) return;
    if (g_ctx) cudaStreamSynchronize(g_ctx->stream);
    if (tree->e0) { cudaEventDestroy(tree->e0); cudaEventDestroy(tree->e1); }
    delete tree;
}

} // extern "C"
