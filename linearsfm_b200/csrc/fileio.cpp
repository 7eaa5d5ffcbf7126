// Drop-in process boundary: the reference's file formats and command line.
//   lmj_readInformationStereo       LinearSFMImp.cpp:3044-3132   (format: SURVEY Appendix A.1)
//   lmj_SaveStateVector             LinearSFMImp.cpp:2102-2117
//   lmj_SavePoses_3DPF              LinearSFMImp.cpp:7876-7967
//   run / lmj_parseArgs / printHelp LinearSFMImp.cpp:7972-8106
// The reader slurps the file and tokenises with strtol/strtod (the reference calls fscanf once per
// number); files of one run are parsed by several host threads.
#include "../../include/linearsfm_b200.h"
#include "common.h"
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <map>
#include <thread>
#include <functional>
#include <chrono>
#include <algorithm>

#include "fast_num.h"

void lsfm_set_error(const std::string &msg);          // capi.cu: the string lsfm_last_error() returns

namespace {

// I/O errors of the single-map entry points go to lsfm_last_error(); the CLI's per-thread loaders keep
// their own strings
struct IoErr {
    std::string s;
    IoErr &operator=(const std::string &m) { s = m; lsfm_set_error(m); return *this; }
} g_io_err;

struct Tok {
    char *p, *end;
    bool fail = false;
    // fastnum::parse_long / parse_double return exactly what strtol / strtod return (fast_num.h), about three
    // times faster on the 15-17 digit numbers of a localmap file
    long next_int()
    {
        char *q;
        long v = fastnum::parse_long(p, &q);
        if (q == p) fail = true;
        p = q;
        return v;
    }
    double next_dbl()
    {
        char *q;
        double v = fastnum::parse_double(p, &q);
        if (q == p) fail = true;
        p = q;
        return v;
    }
};

int load_map_impl(const char *path, lsfm_map *M, std::string &err, bool mono)
{
    FILE *f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return LSFM_ERR_IO; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); err = std::string("cannot determine the size of ") + path; return LSFM_ERR_IO; }
    // every count in the file is bounded by the file size (each entry takes >= 2 bytes of text), so a
    // corrupt header cannot trigger a huge allocation
    const size_t maxCount = (size_t)sz / 2 + 1;
    std::vector<char> buf(sz + 1);
    size_t got = fread(buf.data(), 1, sz, f);
    fclose(f);
    buf[got] = 0;
    Tok t{buf.data(), buf.data() + got};
    memset(M, 0, sizeof(*M));
    M->Ref = (int)t.next_int();
    M->FRef = M->Ref;                                   // LinearSFMImp.cpp:3053 / 6668
    if (mono) {                                         // header Ref ScaP Fix Sign (6669-6676)
        M->ScaP = (int)t.next_int(); M->FScaP = M->ScaP;
        M->Fix = (int)t.next_int(); M->FFix = M->Fix;
        M->Sign = (int)t.next_int();
    }
    M->r = (int)t.next_int();
    if (t.fail || M->r < 0 || (size_t)M->r > maxCount) { err = std::string(path) + ": bad header"; M->r = 0; return LSFM_ERR_FORMAT; }
    bool oom = false;
    auto A = [&oom](size_t n, size_t s) { void *q = malloc((n * s) != 0 ? (n * s) : 1); if (!q) oom = true; return q; };
    M->stno = (int *)A(M->r, sizeof(int));
    M->stVal = (double *)A(M->r, sizeof(double));
    if (oom) { err = std::string(path) + ": out of host memory"; lsfm_free_map(M); return LSFM_ERR_IO; }
    for (int i = 0; i < M->r; i++) { M->stno[i] = (int)t.next_int(); M->stVal[i] = t.next_dbl(); }
    M->m = (int)t.next_int();
    M->n = (int)t.next_int();
    M->nU = (int)t.next_int();
    if (t.fail || M->m < 0 || M->n < 0 || M->nU < 0 || M->r != 6 * M->m + 3 * M->n || 36 * (size_t)M->nU > maxCount) {
        err = std::string(path) + ": inconsistent sizes"; lsfm_free_map(M); return LSFM_ERR_FORMAT;
    }
    M->U = (double *)A(36 * (size_t)M->nU, sizeof(double));
    M->Ui = (int *)A(M->nU, sizeof(int));
    M->Uj = (int *)A(M->nU, sizeof(int));
    if (oom) { err = std::string(path) + ": out of host memory"; lsfm_free_map(M); return LSFM_ERR_IO; }
    for (size_t i = 0; i < 36 * (size_t)M->nU; i++) M->U[i] = t.next_dbl();
    for (int i = 0; i < M->nU; i++) M->Ui[i] = (int)t.next_int();
    for (int i = 0; i < M->nU; i++) M->Uj[i] = (int)t.next_int();
    M->nW = (int)t.next_int();
    if (t.fail || M->nW < 0 || 18 * (size_t)M->nW > maxCount) { err = std::string(path) + ": bad nW"; lsfm_free_map(M); return LSFM_ERR_FORMAT; }
    M->W = (double *)A(18 * (size_t)M->nW, sizeof(double));
    M->photo = (int *)A(M->nW, sizeof(int));
    M->feature = (int *)A(M->nW, sizeof(int));
    M->V = (double *)A(9 * (size_t)M->n, sizeof(double));
    M->FBlock = (int *)A(M->n, sizeof(int));
    if (oom) { err = std::string(path) + ": out of host memory"; lsfm_free_map(M); return LSFM_ERR_IO; }
    for (size_t i = 0; i < 18 * (size_t)M->nW; i++) M->W[i] = t.next_dbl();
    for (int i = 0; i < M->nW; i++) M->photo[i] = (int)t.next_int();
    for (int i = 0; i < M->nW; i++) M->feature[i] = (int)t.next_int();
    for (size_t i = 0; i < 9 * (size_t)M->n; i++) M->V[i] = t.next_dbl();
    for (int i = 0; i < M->n; i++) M->FBlock[i] = (int)t.next_int();
    if (t.fail) { err = std::string(path) + ": truncated file"; lsfm_free_map(M); return LSFM_ERR_FORMAT; }
    return LSFM_OK;
}

void print_help()
{
    printf("Linear SFM Solution General Options\n\n");
    printf("-path			Set Data Path.\n");
    printf("-st            Set Path to Save Final State Vector\n");
    printf("-p			Set Path to Save Poses\n");
    printf("-f			Set Path to Save Features\n");
    printf("-num			Number of Initial Recontruction\n");
    printf("-type			Set Data Type.\n");
    printf("Data Type Listed As Following:\n");
    printf("			I  : Monocular\n");
    printf("			II : Stereo\n");
    printf("\n");
}

} // namespace

extern "C" {

int lsfm_load_localmap_stereo(const char *path, lsfm_map *out)
{
    { std::string e; int rc = load_map_impl(path, out, e, false); if (rc != LSFM_OK) g_io_err = e; return rc; }
}

int lsfm_load_localmap_mono(const char *path, lsfm_map *out)
{
    { std::string e; int rc = load_map_impl(path, out, e, true); if (rc != LSFM_OK) g_io_err = e; return rc; }
}

int lsfm_save_localmap(const lsfm_map *M, const char *path, int mono)
{
    // the reference's localmap_<i>.txt format (readers LinearSFMImp.cpp:3044-3132,
    // 6660-6754), doubles with 17 significant digits so they survive fscanf("%lf") bit for bit
    FILE *fp = fopen(path, "w");
    if (!fp) { g_io_err = std::string("cannot open ") + path; return LSFM_ERR_IO; }
    std::vector<char> buf(1 << 20);
    setvbuf(fp, buf.data(), _IOFBF, buf.size());
    if (mono) fprintf(fp, "%d %d %d %d\n%d\n", M->Ref, M->ScaP, M->Fix, M->Sign, M->r);
    else fprintf(fp, "%d\n%d\n", M->Ref, M->r);
    // The joined map of the NC3500-size scene is 112 M doubles of W alone (2.4 GB of text): the elements are
    // formatted by several host threads, a wave of 4 M elements at a time (same conversions, same bytes as one
    // fprintf per element), and written in order.
    auto elems = [&](size_t cnt, const std::function<int(size_t, char *)> &fmt) {
        const size_t WAVE = (size_t)1 << 22;
        int nth = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 32);
        if (cnt < 50000) nth = 1;
        std::vector<std::string> chunks(nth);
        for (size_t w0 = 0; w0 < cnt; w0 += WAVE) {
            const size_t w1 = std::min(cnt, w0 + WAVE), per = (w1 - w0 + nth - 1) / nth;
            auto work = [&](int t) {
                const size_t a0 = std::min(w1, w0 + per * t), a1 = std::min(w1, a0 + per);
                std::string &out = chunks[t];
                out.clear();
                out.reserve((a1 - a0) * 26);
                char tmp[64];                      // "%d %.17g\n": at most 11 + 1 + 24 + 1 characters
                for (size_t i = a0; i < a1; i++) out.append(tmp, (size_t)fmt(i, tmp));
            };
            if (nth == 1) work(0);
            else {
                std::vector<std::thread> th;
                for (int t = 0; t < nth; t++) th.emplace_back(work, t);
                for (auto &t : th) t.join();
            }
            for (auto &c : chunks) fwrite(c.data(), 1, c.size(), fp);
        }
    };
    elems((size_t)M->r, [&](size_t i, char *o) { return snprintf(o, 64, "%d %.17g\n", M->stno[i], M->stVal[i]); });
    fprintf(fp, "%d %d\n%d\n", M->m, M->n, M->nU);
    auto doubles = [&](const double *x, size_t cnt, size_t per_line) {
        elems(cnt, [&](size_t i, char *o) { return snprintf(o, 64, "%.17g%c", x[i], ((i + 1) % per_line == 0 || i + 1 == cnt) ? '\n' : ' '); });
        if (cnt == 0) fputc('\n', fp);
    };
    auto ints = [&](const int *x, size_t cnt) {
        elems(cnt, [&](size_t i, char *o) { int n = fastnum::put_int(o, x[i]); o[n++] = (i + 1 == cnt) ? '\n' : ' '; return n; });
        if (cnt == 0) fputc('\n', fp);
    };
    doubles(M->U, 36 * (size_t)M->nU, 6);
    ints(M->Ui, M->nU);
    ints(M->Uj, M->nU);
    fprintf(fp, "%d\n", M->nW);
    doubles(M->W, 18 * (size_t)M->nW, 3);
    ints(M->photo, M->nW);
    ints(M->feature, M->nW);
    doubles(M->V, 9 * (size_t)M->n, 3);
    ints(M->FBlock, M->n);
    bool ok = !ferror(fp);
    ok = (fclose(fp) == 0) && ok;
    if (!ok) { g_io_err = std::string("write failed: ") + path; return LSFM_ERR_IO; }
    return LSFM_OK;
}

int lsfm_save_outputs(const lsfm_map *M, const char *st, const char *pose, const char *feat)
{
    // Same bytes as the reference's writers (lmj_SaveStateVector 2102-2117, lmj_SavePoses_3DPF
    // 7876-7967): every line is produced by the same printf conversions ("%d %lf", ...), but the rows
    // are formatted by several host threads into per-chunk buffers and written in order -- at the
    // NC3500 size (1.3 M feature rows) the single-threaded fprintf loop took 0.8 s, 20x the solve.
    auto write_rows = [&](const char *path, size_t nrows, const std::function<int(size_t, char *, size_t)> &fmt) -> int {
        FILE *fp = fopen(path, "w");
        if (!fp) { g_io_err = std::string("cannot open ") + path; return LSFM_ERR_IO; }
        int nth = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 32);
        if (nrows < 20000) nth = 1;
        const size_t per = (nrows + nth - 1) / std::max(nth, 1);
        std::vector<std::string> chunks(nth);
        auto work = [&](int t) {
            size_t r0 = std::min(nrows, per * t), r1 = std::min(nrows, r0 + per);
            std::string &out = chunks[t];
            out.reserve((r1 - r0) * 48);
            char line[2304];                  // six "%lf" of any finite double fit (317 characters each)
            for (size_t r = r0; r < r1; r++) {
                int len = fmt(r, line, sizeof(line));
                out.append(line, (size_t)len);
            }
        };
        if (nth == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nth; t++) th.emplace_back(work, t);
            for (auto &t : th) t.join();
        }
        bool ok = true;
        for (auto &c : chunks) ok = ok && (fwrite(c.data(), 1, c.size(), fp) == c.size());
        ok = (fclose(fp) == 0) && ok;
        return ok ? LSFM_OK : LSFM_ERR_IO;
    };
    if (st) {
        int rc = write_rows(st, (size_t)M->r, [&](size_t i, char *buf, size_t cap) {
            // "%d %lf\n" (fastnum::put_*: the same bytes as printf, tests/test_fast_num.py)
            int n = fastnum::put_int(buf, M->stno[i]);
            buf[n++] = ' ';
            n += fastnum::put_f6(buf + n, cap - n, M->stVal[i]);
            buf[n++] = '\n';
            return n;
        });
        if (rc != LSFM_OK) { printf("Please Input Path to Save Final State Vector!"); return rc; }
    }
    if (!pose && !feat) return LSFM_OK;
    // sorted by id; a repeated id keeps its last occurrence (std::map overwrite, 7907/7925)
    std::vector<std::pair<int, int>> poseIdx, featIdx;
    for (int i = 0; i < M->r;) {
        if (M->stno[i] <= 0) { poseIdx.push_back({-M->stno[i], i}); i += 6; }
        else { featIdx.push_back({M->stno[i], i}); i += 3; }
    }
    auto sort_unique_last = [](std::vector<std::pair<int, int>> &v) {
        std::stable_sort(v.begin(), v.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first < b.first; });
        size_t w = 0;
        for (size_t i = 0; i < v.size(); i++) {
            if (i + 1 < v.size() && v[i + 1].first == v[i].first) continue;     // a later occurrence wins
            v[w++] = v[i];
        }
        v.resize(w);
    };
    sort_unique_last(poseIdx);
    sort_unique_last(featIdx);
    if (pose) {
        int rc = write_rows(pose, poseIdx.size(), [&](size_t k, char *buf, size_t cap) {
            const double *x = M->stVal + poseIdx[k].second;
            // "%d  %lf  %lf  %lf %lf  %lf  %lf\n"
            int n = fastnum::put_int(buf, poseIdx[k].first);
            for (int q = 0; q < 6; q++) {
                buf[n++] = ' ';
                if (q != 3) buf[n++] = ' ';
                n += fastnum::put_f6(buf + n, cap - n, x[q]);
            }
            buf[n++] = '\n';
            return n;
        });
        if (rc != LSFM_OK) return rc;
    }
    if (feat) {
        int rc = write_rows(feat, featIdx.size(), [&](size_t k, char *buf, size_t cap) {
            const double *x = M->stVal + featIdx[k].second;
            // "%d  %lf  %lf %lf\n"
            int n = fastnum::put_int(buf, featIdx[k].first);
            for (int q = 0; q < 3; q++) {
                buf[n++] = ' ';
                if (q != 2) buf[n++] = ' ';
                n += fastnum::put_f6(buf + n, cap - n, x[q]);
            }
            buf[n++] = '\n';
            return n;
        });
        if (rc != LSFM_OK) return rc;
    }
    return LSFM_OK;
}

// ---- binary input cache ------------------------------------------------------
// The reference parses 3499 text files with one fscanf per number on every run (LinearSFMImp.cpp:3062-3128).
// The cache is the same maps as raw arrays in ONE file: header {magic, version, num, mono}, then per map the
// 12 header ints and the ten arrays of struct lsfm_map in declaration order.  Native endianness and sizes.
static const char CACHE_MAGIC[8] = {'L', 'S', 'F', 'M', 'B', 'C', '0', '1'};

int lsfm_save_cache(const char *file, const lsfm_map *maps, int num, int mono)
{
    FILE *f = fopen(file, "wb");
    if (!f) { lsfm_set_error(std::string("cannot write ") + file); return LSFM_ERR_IO; }
    bool ok = fwrite(CACHE_MAGIC, 1, 8, f) == 8;
    int hdr[2] = {num, mono ? 1 : 0};
    ok = ok && fwrite(hdr, sizeof(int), 2, f) == 2;
    auto put = [&](const void *p, size_t bytes) { if (ok && bytes) ok = fwrite(p, 1, bytes, f) == bytes; };
    for (int i = 0; i < num && ok; i++) {
        const lsfm_map &M = maps[i];
        int h[12] = {M.Ref, M.FRef, M.m, M.n, M.nU, M.nW, M.r, M.ScaP, M.Fix, M.Sign, M.FScaP, M.FFix};
        put(h, sizeof(h));
        put(M.stno, sizeof(int) * (size_t)M.r);
        put(M.stVal, sizeof(double) * (size_t)M.r);
        put(M.U, sizeof(double) * 36 * (size_t)M.nU);
        put(M.Ui, sizeof(int) * (size_t)M.nU);
        put(M.Uj, sizeof(int) * (size_t)M.nU);
        put(M.W, sizeof(double) * 18 * (size_t)M.nW);
        put(M.photo, sizeof(int) * (size_t)M.nW);
        put(M.feature, sizeof(int) * (size_t)M.nW);
        put(M.V, sizeof(double) * 9 * (size_t)M.n);
        put(M.FBlock, sizeof(int) * (size_t)M.n);
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) { lsfm_set_error(std::string("write error on ") + file); return LSFM_ERR_IO; }
    return LSFM_OK;
}

int lsfm_load_cache(const char *file, lsfm_map **maps_out, int *num_out, int *mono_out)
{
    if (!maps_out || !num_out) { lsfm_set_error("load_cache: null output"); return LSFM_ERR_ARG; }
    *maps_out = nullptr; *num_out = 0;
    FILE *f = fopen(file, "rb");
    if (!f) { lsfm_set_error(std::string("cannot open ") + file); return LSFM_ERR_IO; }
    char magic[8];
    int hdr[2] = {0, 0};
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, CACHE_MAGIC, 8) != 0 || fread(hdr, sizeof(int), 2, f) != 2 ||
        hdr[0] < 0) {
        fclose(f);
        lsfm_set_error(std::string(file) + ": not a LinearSFM map cache");
        return LSFM_ERR_FORMAT;
    }
    const int num = hdr[0];
    lsfm_map *maps = (lsfm_map *)calloc((size_t)std::max(num, 1), sizeof(lsfm_map));
    bool ok = maps != nullptr;
    auto get = [&](void **dst, size_t bytes) {
        if (!ok) return;
        *dst = malloc(bytes ? bytes : 1);
        if (!*dst) { ok = false; return; }
        if (bytes && fread(*dst, 1, bytes, f) != bytes) ok = false;
    };
    for (int i = 0; i < num && ok; i++) {
        lsfm_map &M = maps[i];
        int h[12];
        if (fread(h, sizeof(int), 12, f) != 12) { ok = false; break; }
        M.Ref = h[0]; M.FRef = h[1]; M.m = h[2]; M.n = h[3]; M.nU = h[4]; M.nW = h[5]; M.r = h[6];
        M.ScaP = h[7]; M.Fix = h[8]; M.Sign = h[9]; M.FScaP = h[10]; M.FFix = h[11];
        if (M.m < 0 || M.n < 0 || M.nU < 0 || M.nW < 0 || M.r != 6 * M.m + 3 * M.n) { ok = false; break; }
        get((void **)&M.stno, sizeof(int) * (size_t)M.r);
        get((void **)&M.stVal, sizeof(double) * (size_t)M.r);
        get((void **)&M.U, sizeof(double) * 36 * (size_t)M.nU);
        get((void **)&M.Ui, sizeof(int) * (size_t)M.nU);
        get((void **)&M.Uj, sizeof(int) * (size_t)M.nU);
        get((void **)&M.W, sizeof(double) * 18 * (size_t)M.nW);
        get((void **)&M.photo, sizeof(int) * (size_t)M.nW);
        get((void **)&M.feature, sizeof(int) * (size_t)M.nW);
        get((void **)&M.V, sizeof(double) * 9 * (size_t)M.n);
        get((void **)&M.FBlock, sizeof(int) * (size_t)M.n);
    }
    fclose(f);
    if (!ok) {
        if (maps) { for (int i = 0; i < num; i++) lsfm_free_map(&maps[i]); free(maps); }
        lsfm_set_error(std::string(file) + ": truncated or corrupt map cache");
        return LSFM_ERR_FORMAT;
    }
    *maps_out = maps; *num_out = num;
    if (mono_out) *mono_out = hdr[1];
    return LSFM_OK;
}

void lsfm_free_cache(lsfm_map *maps, int num)
{
    if (!maps) return;
    for (int i = 0; i < num; i++) lsfm_free_map(&maps[i]);
    free(maps);
}

int lsfm_cli_main(int argc, char **argv)
{
    std::string path, st, pose, feat, type, mapout, cache, covout;
    int num = 0;
    bool hasPath = false, hasNum = false, hasType = false;
    for (int i = 1; i < argc; i++) {
        std::string name = argv[i];
        if (name[0] != '-') return 0;                    // LinearSFMImp.cpp:8000-8002: silent return
        size_t d = name.find_first_not_of('-');
        if (d != std::string::npos) name = name.substr(d);
        if (name == "help") { print_help(); return 0; }
        auto arg = [&](std::string &dst) { if (i + 1 < argc) dst = argv[++i]; };
        if (name == "path") { arg(path); hasPath = true; }
        else if (name == "st") arg(st);
        else if (name == "p") arg(pose);
        else if (name == "f") arg(feat);
        else if (name == "map") arg(mapout);          // extension: the joined map (state + information) in localmap format
        else if (name == "cache") arg(cache);         // extension: binary cache of the parsed input maps
        else if (name == "cov") arg(covout);          // extension: marginal pose covariances of the joined map
        else if (name == "num") { std::string v; arg(v); num = atoi(v.c_str()); hasNum = true; }
        else if (name == "type") {
            std::string v; arg(v);
            if (v == "Monocular" || v == "Stereo") { type = v; hasType = true; }
        }
    }
    if (!hasPath) { printf("LinerSFM Error: Please Input Right File Path:\n"); return 0; }
    if (!hasNum) { printf("LinerSFM Error: Please Set Local Map Number:\n"); return 0; }
    if (!hasType) { printf("LinerSFM Error: Please Set Data Type:\n"); return 0; }
    const bool mono = (type == "Monocular");
    if (num < 1) return 0;
    if (lsfm_init(0) != LSFM_OK) { fprintf(stderr, "LinearSFM (B200): %s\n", lsfm_last_error()); return 1; }

    std::vector<lsfm_map> maps(num);
    // -cache <file>: an existing cache with the same map count and type replaces the text parse; otherwise
    // the text files are parsed and the cache is (re)written
    bool fromCache = false;
    if (!cache.empty()) {
        lsfm_map *cm = nullptr;
        int cn = 0, cmono = 0;
        if (lsfm_load_cache(cache.c_str(), &cm, &cn, &cmono) == LSFM_OK) {
            if (cn == num && (cmono != 0) == mono) {
                for (int i = 0; i < num; i++) maps[i] = cm[i];      // the arrays move into `maps`
                free(cm);
                fromCache = true;
            } else lsfm_free_cache(cm, cn);
        }
    }
    if (!fromCache) {
        std::vector<int> rc(num, LSFM_OK);
        std::vector<std::string> errs(num);
        int nth = std::min<int>(std::max(1u, std::thread::hardware_concurrency()), 32);
        nth = std::min(nth, num);
        auto work = [&](int tid) {
            for (int i = tid; i < num; i += nth) {
                std::string p = path + "/localmap_" + std::to_string(i + 1) + ".txt";
                rc[i] = load_map_impl(p.c_str(), &maps[i], errs[i], mono);
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < nth; t++) th.emplace_back(work, t);
        for (auto &t : th) t.join();
        for (int i = 0; i < num; i++)
            if (rc[i] != LSFM_OK) { fprintf(stderr, "LinearSFM (B200): %s\n", errs[i].c_str()); return 1; }
        if (!cache.empty() && lsfm_save_cache(cache.c_str(), maps.data(), num, mono ? 1 : 0) != LSFM_OK)
            fprintf(stderr, "LinearSFM (B200): %s (continuing without the cache)\n", lsfm_last_error());
    }

    if (mono) {
        // CLinearSFMImp::runMono (LinearSFMImp.cpp:3136-3152)
        lsfm_map outm;
        auto tm0 = std::chrono::steady_clock::now();
        if (lsfm_run_mono_ex(maps.data(), num, 1, &outm) != LSFM_OK) {
            fprintf(stderr, "LinearSFM (B200): %s\n", lsfm_last_error()); return 1;
        }
        double secm = std::chrono::duration<double>(std::chrono::steady_clock::now() - tm0).count();
        printf("Total Used Time:  %lf  sec\n\n", secm);        // LinearSFMImp.cpp:6639 (incl. H2D/D2H here)
        for (auto &m : maps) lsfm_free_map(&m);
        int wrc = LSFM_OK;
        if (!st.empty()) wrc |= lsfm_save_outputs(&outm, st.c_str(), nullptr, nullptr);
        if (!pose.empty() && !feat.empty()) wrc |= lsfm_save_outputs(&outm, nullptr, pose.c_str(), feat.c_str());
        if (!mapout.empty()) wrc |= lsfm_save_localmap(&outm, mapout.c_str(), 1);
        lsfm_free_map(&outm);
        if (wrc != LSFM_OK) { fprintf(stderr, "LinearSFM (B200): %s\n", lsfm_last_error()); return 1; }
        return 0;
    }
    lsfm_tree *tree = nullptr;
    if (lsfm_tree_create_stereo(maps.data(), num, &tree) != LSFM_OK) {
        fprintf(stderr, "LinearSFM (B200): %s\n", lsfm_last_error()); return 1;
    }
    for (auto &m : maps) lsfm_free_map(&m);
    auto t0 = std::chrono::steady_clock::now();
    if (lsfm_tree_solve(tree, 1, 0, -1) != LSFM_OK) {
        fprintf(stderr, "LinearSFM (B200): %s\n", lsfm_last_error()); return 1;
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("Total Used Time:  %lf  sec\n\n", sec);          // LinearSFMImp.cpp:2072
    lsfm_map out;
    if (!mapout.empty()) {
        if (lsfm_tree_download(tree, 0, &out) != LSFM_OK) {
            fprintf(stderr, "LinearSFM (B200): %s\n", lsfm_last_error()); return 1;
        }
    } else {
        // the reference's outputs only need the state vector: 12 bytes per row come back from the
        // device instead of the whole information matrix (~1 GB of W blocks at the NC3500 size)
        if (lsfm_tree_result_shape(tree, 0, &out) != LSFM_OK) {
            fprintf(stderr, "LinearSFM (B200): %s\n", lsfm_last_error()); return 1;
        }
        out.stno = (int *)malloc(sizeof(int) * (size_t)std::max(out.r, 1));
        out.stVal = (double *)malloc(sizeof(double) * (size_t)std::max(out.r, 1));
        if (lsfm_tree_download_state(tree, 0, out.stno, out.stVal) != LSFM_OK) {
            fprintf(stderr, "LinearSFM (B200): %s\n", lsfm_last_error()); return 1;
        }
    }
    if (!covout.empty()) {
        // -cov <file> (Stereo): 6x6 marginal covariance of at most 64 poses, evenly spaced, the last pose always
        // among them; one line per pose: pose id, then the 36 entries row by row
        const int mP = out.m, stride = std::max(1, (mP + 62) / 63);
        std::vector<int> sel;
        for (int p = 0; p < mP; p += stride) sel.push_back(p);
        if (mP > 0 && sel.back() != mP - 1) sel.push_back(mP - 1);
        std::vector<double> cv(36 * sel.size());
        if (lsfm_tree_marginal_cov(tree, 0, (int)sel.size(), sel.data(), cv.data()) != LSFM_OK) {
            fprintf(stderr, "LinearSFM (B200): %s\n", lsfm_last_error()); return 1;
        }
        FILE *fc = fopen(covout.c_str(), "w");
        if (!fc) { fprintf(stderr, "LinearSFM (B200): cannot write %s\n", covout.c_str()); return 1; }
        for (size_t i = 0; i < sel.size(); i++) {
            fprintf(fc, "%d", -out.stno[6 * (size_t)sel[i]]);
            for (int q = 0; q < 36; q++) fprintf(fc, " %.17g", cv[36 * i + q]);
            fprintf(fc, "\n");
        }
        fclose(fc);
    }
    lsfm_tree_free(tree);
    // usage problems return 0 like the reference (LinearSFM.cpp:17); I/O and solve failures return 1
    int wrc = LSFM_OK;
    if (!st.empty()) wrc |= lsfm_save_outputs(&out, st.c_str(), nullptr, nullptr);
    if (!pose.empty() && !feat.empty())                     // both required, LinearSFMImp.cpp:2078
        wrc |= lsfm_save_outputs(&out, nullptr, pose.c_str(), feat.c_str());
    if (!mapout.empty()) wrc |= lsfm_save_localmap(&out, mapout.c_str(), 0);
    lsfm_free_map(&out);
    if (wrc != LSFM_OK) { fprintf(stderr, "LinearSFM (B200): %s\n", lsfm_last_error()); return 1; }
    return 0;
}

} // extern "C"
