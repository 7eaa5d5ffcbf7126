/*
 * oracle/cholmod_shim.c -- TEST INFRASTRUCTURE ONLY (CPU). Not part of the product path.
 *
 * The reference (LinearSFMImp.cpp) calls 12 CHOLMOD entry points whose library (SuiteSparse
 * CHOLMOD 1.6.0 + AMD + BLAS, see windows/include/cholmod/cholmod_core.h:247-251) is NOT shipped
 * under /root/reference for Linux.  This file restates the *published semantics* of those entry
 * points (cholmod_cholesky.h:75-180, 278-290; cholmod_core.h:1094-1142, 1476-1595, 1755-1766):
 *   upper-triangular stype=1 CSC input, fill-reducing permutation P, P A P' = L L', solve A x = b.
 * so that the UNMODIFIED reference translation unit links and runs here.
 *
 * Call sites it serves (reference file:line):
 *   cholmod_start/finish            Imp.cpp:86, 92, 2335, 2376, 6976, 7039
 *   cholmod_zeros                   Imp.cpp:2340, 6981
 *   cholmod_allocate_sparse         Imp.cpp:2344, 6985
 *   cholmod_amd                     Imp.cpp:2413, 7081
 *   cholmod_analyze / analyze_p     Imp.cpp:2389, 2440, 7056, 7112
 *   cholmod_factorize / solve       Imp.cpp:2444-2445, 7116-7117
 *   cholmod_free_{factor,sparse,dense}  Imp.cpp:2372-2375, 7035-7038
 *
 * PARITY STATUS: "parity unpinned" w.r.t. the real SuiteSparse binaries (absent).  The linear
 * system is SPD, so its solution is unique; the shim's answer is checked by residual and against
 * scipy in tests/test_oracle_shim.py.  The ORDERING is not SuiteSparse AMD: it is the documented
 * deterministic "LSFM-ND" ordering (DESIGN.md section "Ordering"), implemented here naively and
 * independently of the product's implementation (linearsfm_b200/csrc/chol_symbolic.cpp) so the
 * bit-exact ordering test compares two separate codes.
 *
 * Numeric method: simplicial up-looking LL' (the row-by-row algorithm CHOLMOD's simplicial
 * factorisation and CSparse's cs_chol use: elimination tree, row-subtree reach, sparse
 * triangular solve per row).
 *
 * Hazard handled: pba_solveCholmodLM/GN take cholmod_common BY VALUE, objects are
 * allocated under one common and freed under another -> the shim keeps no per-common state.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include "suitesparse/cholmod.h"

/* ------------------------------------------------------------------------------------------ */
/* capture buffers: tests read the last system the reference handed to "CHOLMOD"              */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int m;            /* block pattern dimension (cholmod_amd input) */
    int *Ap, *Ai;     /* block pattern, upper CSC */
    int *bperm;       /* block permutation returned by cholmod_amd */
    int n;            /* scalar dimension */
    int *Sp, *Si;     /* scalar upper CSC */
    double *Sx;
    int *perm;        /* scalar permutation used by analyze */
    double *b, *x;    /* rhs / solution of last cholmod_solve */
    long n_amd, n_factorize, n_solve;
    double t_amd, t_analyze, t_factorize, t_solve; /* seconds spent inside the shim */
} shim_capture_t;

static shim_capture_t g_cap;
static int g_capture_enabled = 0;

#include <time.h>
static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void lsfm_shim_capture_enable(int on) { g_capture_enabled = on; }
const shim_capture_t *lsfm_shim_capture(void) { return &g_cap; }
void lsfm_shim_reset_timers(void)
{
    g_cap.n_amd = g_cap.n_factorize = g_cap.n_solve = 0;
    g_cap.t_amd = g_cap.t_analyze = g_cap.t_factorize = g_cap.t_solve = 0.0;
}
double lsfm_shim_time(int which)
{
    switch (which) { case 0: return g_cap.t_amd; case 1: return g_cap.t_analyze;
                     case 2: return g_cap.t_factorize; default: return g_cap.t_solve; }
}

static void cap_ints(int **dst, const int *src, size_t n)
{
    free(*dst); *dst = (int *)malloc((n ? n : 1) * sizeof(int));
    if (n) memcpy(*dst, src, n * sizeof(int));
}
static void cap_dbls(double **dst, const double *src, size_t n)
{
    free(*dst); *dst = (double *)malloc((n ? n : 1) * sizeof(double));
    if (n) memcpy(*dst, src, n * sizeof(double));
}

/* ------------------------------------------------------------------------------------------ */
/* LSFM-ND ordering (spec in DESIGN.md): naive recursive reference implementation             */
/*   n <= 32 : identity                                                                        */
/*   dissect(L), L ascending:                                                                  */
/*       |L| <= 8 : emit L                                                                     */
/*       A = first floor(|L|/2), B = rest                                                      */
/*       S = greedy vertex cover of the A-B cross edges: take the vertex with most uncovered   */
/*           cross edges (tie: smallest index) until every cross edge is covered               */
/*       dissect(A \ S); dissect(B \ S); emit S ascending                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int n; int *ptr; int *adj; } graph_t;

static graph_t build_graph_upper(int n, const int *Ap, const int *Ai)
{
    graph_t g; g.n = n;
    int *cnt = (int *)calloc(n + 1, sizeof(int));
    for (int j = 0; j < n; j++)
        for (int p = Ap[j]; p < Ap[j + 1]; p++) {
            int i = Ai[p];
            if (i != j) { cnt[i]++; cnt[j]++; }
        }
    g.ptr = (int *)malloc((n + 1) * sizeof(int));
    g.ptr[0] = 0;
    for (int i = 0; i < n; i++) g.ptr[i + 1] = g.ptr[i] + cnt[i];
    g.adj = (int *)malloc((g.ptr[n] ? g.ptr[n] : 1) * sizeof(int));
    memset(cnt, 0, (n + 1) * sizeof(int));
    for (int j = 0; j < n; j++)
        for (int p = Ap[j]; p < Ap[j + 1]; p++) {
            int i = Ai[p];
            if (i != j) { g.adj[g.ptr[i] + cnt[i]++] = j; g.adj[g.ptr[j] + cnt[j]++] = i; }
        }
    free(cnt);
    return g;
}

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

static int g_nd_max_sep;   /* largest separator of the current dissection */
/* where[v] : 0 outside the current list, 1 = A, 2 = B, 3 = separator; restored to 0 on return */
static void nd_dissect(const graph_t *g, const int *L, int len, char *where, int *out, int *nout)
{
    if (len <= 0) return;
    if (len <= 8) { for (int i = 0; i < len; i++) out[(*nout)++] = L[i]; return; }
    int h = len / 2;
    for (int i = 0; i < len; i++) where[L[i]] = (i < h) ? 1 : 2;
    int *sep = (int *)malloc(len * sizeof(int)), ns = 0;
    for (;;) {
        /* recount from scratch every round: O(|S| * edges), fine for an oracle */
        int best = -1, bc = 0;
        for (int i = 0; i < len; i++) {
            int v = L[i];
            if (where[v] == 3) continue;
            char other = (where[v] == 1) ? 2 : 1;
            int c = 0;
            for (int p = g->ptr[v]; p < g->ptr[v + 1]; p++) c += (where[g->adj[p]] == other);
            if (c > bc) { bc = c; best = v; }
        }
        if (best < 0) break;
        where[best] = 3;
        sep[ns++] = best;
    }
    qsort(sep, ns, sizeof(int), cmp_int);
    if (ns > g_nd_max_sep) g_nd_max_sep = ns;
    int *A = (int *)malloc(len * sizeof(int)), *B = (int *)malloc(len * sizeof(int)), na = 0, nb = 0;
    for (int i = 0; i < len; i++) {
        if (where[L[i]] == 1) A[na++] = L[i];
        else if (where[L[i]] == 2) B[nb++] = L[i];
    }
    for (int i = 0; i < len; i++) where[L[i]] = 0;
    nd_dissect(g, A, na, where, out, nout);
    nd_dissect(g, B, nb, where, out, nout);
    for (int i = 0; i < ns; i++) out[(*nout)++] = sep[i];
    free(A); free(B); free(sep);
}

/* LSFM-MD (spec in DESIGN.md): exact minimum degree on the block graph, naive set implementation.  */
/*   repeat n times: v = remaining vertex with the fewest remaining neighbours (tie: smallest index); */
/*   emit v; the neighbours of v become pairwise adjacent; v leaves the graph                         */
static void lsfm_md_order(const graph_t *g, int *perm)
{
    int n = g->n;
    unsigned char *a = (unsigned char *)calloc((size_t)n * n, 1);   /* dense adjacency: an oracle may */
    char *gone = (char *)calloc(n, 1);
    int *nb = (int *)malloc(n * sizeof(int));
    for (int v = 0; v < n; v++)
        for (int p = g->ptr[v]; p < g->ptr[v + 1]; p++) a[(size_t)v * n + g->adj[p]] = 1;
    int *deg = (int *)calloc(n, sizeof(int));
    for (int v = 0; v < n; v++) deg[v] = g->ptr[v + 1] - g->ptr[v];
    for (int k = 0; k < n; k++) {
        int v = -1;
        for (int i = 0; i < n; i++)
            if (!gone[i] && (v < 0 || deg[i] < deg[v])) v = i;
        perm[k] = v;
        gone[v] = 1;
        int c = 0;
        for (int u = 0; u < n; u++)
            if (a[(size_t)v * n + u] && !gone[u]) nb[c++] = u;
        for (int x = 0; x < c; x++) {
            unsigned char *ru = a + (size_t)nb[x] * n;
            for (int y = 0; y < c; y++)
                if (y != x) ru[nb[y]] = 1;
            ru[v] = 0;
            int d = 0;
            for (int u = 0; u < n; u++) d += (ru[u] && !gone[u]);
            deg[nb[x]] = d;
        }
    }
    free(a); free(gone); free(nb); free(deg);
}

/* the ordering rule: LSFM-ND unless one of its separators has more than 64 vertices, then LSFM-MD */
static void lsfm_nd_order(int n, const int *Ap, const int *Ai, int *perm)
{
    if (n <= 32) { for (int i = 0; i < n; i++) perm[i] = i; return; }
    graph_t g = build_graph_upper(n, Ap, Ai);
    int *L = (int *)malloc(n * sizeof(int));
    for (int i = 0; i < n; i++) L[i] = i;
    char *where = (char *)calloc(n, 1);
    int nout = 0;
    g_nd_max_sep = 0;
    nd_dissect(&g, L, n, where, perm, &nout);
    if (g_nd_max_sep > 64) lsfm_md_order(&g, perm);
    free(where); free(L); free(g.ptr); free(g.adj);
}

/* exported for the ordering parity test (block pattern in, permutation out) */
void lsfm_shim_order(int n, const int *Ap, const int *Ai, int *perm) { lsfm_nd_order(n, Ap, Ai, perm); }

/* ------------------------------------------------------------------------------------------ */
/* simplicial up-looking LL'                                                                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int n;
    int *perm, *iperm;
    int *parent;      /* elimination tree of P A P' */
    int *Lp, *Li;     /* CSC of L (row indices in increasing order of discovery) */
    double *Lx;
    int *Cp, *Ci;     /* upper CSC pattern of C = P A P' and map back into A->x */
    int *Cmap;
    int numeric_ok;
} shim_factor_t;

/* C = P A P' (upper part), column j of C holds entries (i,j) with i<=j in new numbering */
static void sym_permute_upper(int n, const int *Ap, const int *Ai, const int *iperm,
                              int **Cp_out, int **Ci_out, int **Cmap_out)
{
    int *cnt = (int *)calloc(n + 1, sizeof(int));
    for (int j = 0; j < n; j++)
        for (int p = Ap[j]; p < Ap[j + 1]; p++) {
            int i = Ai[p]; if (i > j) continue;
            int i2 = iperm[i], j2 = iperm[j];
            cnt[i2 > j2 ? i2 : j2]++;
        }
    int *Cp = (int *)malloc((n + 1) * sizeof(int));
    Cp[0] = 0;
    for (int j = 0; j < n; j++) Cp[j + 1] = Cp[j] + cnt[j];
    int nz = Cp[n];
    int *Ci = (int *)malloc((nz ? nz : 1) * sizeof(int));
    int *Cm = (int *)malloc((nz ? nz : 1) * sizeof(int));
    memset(cnt, 0, (n + 1) * sizeof(int));
    for (int j = 0; j < n; j++)
        for (int p = Ap[j]; p < Ap[j + 1]; p++) {
            int i = Ai[p]; if (i > j) continue;
            int i2 = iperm[i], j2 = iperm[j];
            int r = i2 < j2 ? i2 : j2, c = i2 > j2 ? i2 : j2;
            int q = Cp[c] + cnt[c]++;
            Ci[q] = r; Cm[q] = p;
        }
    free(cnt);
    *Cp_out = Cp; *Ci_out = Ci; *Cmap_out = Cm;
}

static void etree_upper(int n, const int *Cp, const int *Ci, int *parent)
{
    int *anc = (int *)malloc(n * sizeof(int));
    for (int k = 0; k < n; k++) {
        parent[k] = -1; anc[k] = -1;
        for (int p = Cp[k]; p < Cp[k + 1]; p++) {
            int i = Ci[p];
            while (i != -1 && i < k) {
                int nx = anc[i]; anc[i] = k;
                if (nx == -1) parent[i] = k;
                i = nx;
            }
        }
    }
    free(anc);
}

/* nonzero pattern of row k of L: s[top..n-1], topological order. w = marks (w[i]==k => visited) */
static int row_reach(int k, const int *Cp, const int *Ci, const int *parent, int *s, int *w, int n)
{
    int top = n;
    w[k] = k;
    for (int p = Cp[k]; p < Cp[k + 1]; p++) {
        int i = Ci[p];
        if (i > k) continue;
        int len = 0;
        for (; w[i] != k; i = parent[i]) { s[len++] = i; w[i] = k; }
        while (len > 0) s[--top] = s[--len];
    }
    return top;
}

static shim_factor_t *shim_symbolic(cholmod_sparse *A, const int *perm)
{
    int n = (int)A->ncol;
    const int *Ap = (const int *)A->p, *Ai = (const int *)A->i;
    shim_factor_t *F = (shim_factor_t *)calloc(1, sizeof(shim_factor_t));
    F->n = n;
    F->perm = (int *)malloc((n ? n : 1) * sizeof(int));
    F->iperm = (int *)malloc((n ? n : 1) * sizeof(int));
    for (int k = 0; k < n; k++) { F->perm[k] = perm ? perm[k] : k; }
    for (int k = 0; k < n; k++) F->iperm[F->perm[k]] = k;
    sym_permute_upper(n, Ap, Ai, F->iperm, &F->Cp, &F->Ci, &F->Cmap);
    F->parent = (int *)malloc((n ? n : 1) * sizeof(int));
    etree_upper(n, F->Cp, F->Ci, F->parent);
    /* column counts by one reach pass */
    int *cnt = (int *)calloc(n + 1, sizeof(int));
    int *s = (int *)malloc((n ? n : 1) * sizeof(int));
    int *w = (int *)malloc((n ? n : 1) * sizeof(int));
    for (int i = 0; i < n; i++) w[i] = -1;
    for (int k = 0; k < n; k++) {
        int top = row_reach(k, F->Cp, F->Ci, F->parent, s, w, n);
        for (int t = top; t < n; t++) cnt[s[t]]++;
        cnt[k]++;
    }
    F->Lp = (int *)malloc((n + 1) * sizeof(int));
    F->Lp[0] = 0;
    for (int j = 0; j < n; j++) F->Lp[j + 1] = F->Lp[j] + cnt[j];
    size_t lnz = (size_t)F->Lp[n];
    F->Li = (int *)malloc((lnz ? lnz : 1) * sizeof(int));
    F->Lx = (double *)malloc((lnz ? lnz : 1) * sizeof(double));
    free(cnt); free(s); free(w);
    return F;
}

static int shim_numeric(cholmod_sparse *A, shim_factor_t *F)
{
    int n = F->n;
    const double *Ax = (const double *)A->x;
    int *c = (int *)malloc((n ? n : 1) * sizeof(int));      /* next free slot per column */
    int *s = (int *)malloc((n ? n : 1) * sizeof(int));
    int *w = (int *)malloc((n ? n : 1) * sizeof(int));
    double *xw = (double *)calloc((n ? n : 1), sizeof(double));
    for (int i = 0; i < n; i++) { c[i] = F->Lp[i]; w[i] = -1; }
    int ok = 1;
    for (int k = 0; k < n; k++) {
        int top = row_reach(k, F->Cp, F->Ci, F->parent, s, w, n);
        xw[k] = 0.0;
        for (int p = F->Cp[k]; p < F->Cp[k + 1]; p++)
            if (F->Ci[p] <= k) xw[F->Ci[p]] = Ax[F->Cmap[p]];
        double d = xw[k]; xw[k] = 0.0;
        for (; top < n; top++) {
            int i = s[top];
            double lki = xw[i] / F->Lx[F->Lp[i]];
            xw[i] = 0.0;
            for (int p = F->Lp[i] + 1; p < c[i]; p++) xw[F->Li[p]] -= F->Lx[p] * lki;
            d -= lki * lki;
            int p = c[i]++;
            F->Li[p] = k; F->Lx[p] = lki;
        }
        if (!(d > 0.0)) { ok = 0; d = 1.0; fprintf(stderr, "cholmod_shim: non-positive pivot at %d\n", k); }
        int p = c[k]++;
        F->Li[p] = k; F->Lx[p] = sqrt(d);
    }
    free(c); free(s); free(w); free(xw);
    F->numeric_ok = ok;
    return ok;
}

/* ------------------------------------------------------------------------------------------ */
/* CHOLMOD entry points                                                                        */
/* ------------------------------------------------------------------------------------------ */
int cholmod_start(cholmod_common *Common)
{
    if (!Common) return 0;
    memset(Common, 0, sizeof(*Common));
    Common->status = CHOLMOD_OK;
    Common->nmethods = 0;
    Common->itype = CHOLMOD_INT;
    Common->dtype = CHOLMOD_DOUBLE;
    return 1;
}

int cholmod_finish(cholmod_common *Common) { (void)Common; return 1; }

cholmod_dense *cholmod_zeros(size_t nrow, size_t ncol, int xtype, cholmod_common *Common)
{
    (void)Common;
    cholmod_dense *X = (cholmod_dense *)calloc(1, sizeof(cholmod_dense));
    X->nrow = nrow; X->ncol = ncol; X->d = nrow; X->nzmax = nrow * ncol;
    X->x = calloc(X->nzmax ? X->nzmax : 1, sizeof(double));
    X->z = NULL; X->xtype = xtype; X->dtype = CHOLMOD_DOUBLE;
    return X;
}

cholmod_sparse *cholmod_allocate_sparse(size_t nrow, size_t ncol, size_t nzmax, int sorted,
                                        int packed, int stype, int xtype, cholmod_common *Common)
{
    (void)Common;
    cholmod_sparse *A = (cholmod_sparse *)calloc(1, sizeof(cholmod_sparse));
    A->nrow = nrow; A->ncol = ncol; A->nzmax = nzmax ? nzmax : 1;
    A->p = calloc(ncol + 1, sizeof(int));
    A->i = calloc(A->nzmax, sizeof(int));
    A->nz = packed ? NULL : calloc(ncol ? ncol : 1, sizeof(int));
    A->x = (xtype == CHOLMOD_PATTERN) ? NULL : calloc(A->nzmax, sizeof(double));
    A->z = NULL;
    A->stype = stype; A->itype = CHOLMOD_INT; A->xtype = xtype; A->dtype = CHOLMOD_DOUBLE;
    A->sorted = sorted; A->packed = packed;
    return A;
}

int cholmod_amd(cholmod_sparse *A, int *fset, size_t fsize, int *Perm, cholmod_common *Common)
{
    (void)fset; (void)fsize; (void)Common;
    double t0 = now_s();
    int n = (int)A->ncol;
    lsfm_nd_order(n, (const int *)A->p, (const int *)A->i, Perm);
    g_cap.t_amd += now_s() - t0; g_cap.n_amd++;
    if (g_capture_enabled) {
        g_cap.m = n;
        cap_ints(&g_cap.Ap, (const int *)A->p, n + 1);
        cap_ints(&g_cap.Ai, (const int *)A->i, ((const int *)A->p)[n]);
        cap_ints(&g_cap.bperm, Perm, n);
    }
    return 1;
}

static cholmod_factor *wrap_factor(shim_factor_t *F, int ordering)
{
    cholmod_factor *L = (cholmod_factor *)calloc(1, sizeof(cholmod_factor));
    L->n = F->n; L->minor = F->n;
    L->Perm = F->perm; L->p = F->Lp; L->i = F->Li; L->x = F->Lx;
    L->nzmax = (size_t)F->Lp[F->n];
    L->ordering = ordering; L->is_ll = 1; L->is_super = 0; L->is_monotonic = 1;
    L->itype = CHOLMOD_INT; L->xtype = CHOLMOD_PATTERN; L->dtype = CHOLMOD_DOUBLE;
    L->z = (void *)F;             /* private handle (z is unused for real factors) */
    return L;
}

cholmod_factor *cholmod_analyze_p(cholmod_sparse *A, int *UserPerm, int *fset, size_t fsize,
                                  cholmod_common *Common)
{
    (void)fset; (void)fsize;
    double t0 = now_s();
    int n = (int)A->ncol;
    int *perm = (int *)malloc((n ? n : 1) * sizeof(int));
    int ordering = CHOLMOD_NATURAL;
    if (Common && Common->nmethods >= 1) ordering = Common->method[0].ordering;
    if (UserPerm && (ordering == CHOLMOD_GIVEN || !Common || Common->nmethods == 0)) {
        memcpy(perm, UserPerm, n * sizeof(int)); ordering = CHOLMOD_GIVEN;
    } else if (ordering == CHOLMOD_AMD || (Common && Common->nmethods == 0)) {
        /* default strategy of CHOLMOD tries AMD first; the shim's "AMD" is LSFM-ND */
        lsfm_nd_order(n, (const int *)A->p, (const int *)A->i, perm); ordering = CHOLMOD_AMD;
    } else {
        for (int i = 0; i < n; i++) perm[i] = i;
    }
    shim_factor_t *F = shim_symbolic(A, perm);
    free(perm);
    g_cap.t_analyze += now_s() - t0;
    if (g_capture_enabled) { g_cap.n = n; cap_ints(&g_cap.perm, F->perm, n); }
    return wrap_factor(F, ordering);
}

cholmod_factor *cholmod_analyze(cholmod_sparse *A, cholmod_common *Common)
{
    return cholmod_analyze_p(A, NULL, NULL, 0, Common);
}

int cholmod_factorize(cholmod_sparse *A, cholmod_factor *L, cholmod_common *Common)
{
    double t0 = now_s();
    shim_factor_t *F = (shim_factor_t *)L->z;
    int ok = shim_numeric(A, F);
    L->xtype = CHOLMOD_REAL;
    if (Common) Common->status = ok ? CHOLMOD_OK : 1 /* CHOLMOD_NOT_POSDEF */;
    g_cap.t_factorize += now_s() - t0; g_cap.n_factorize++;
    if (g_capture_enabled) {
        int n = (int)A->ncol; g_cap.n = n;
        cap_ints(&g_cap.Sp, (const int *)A->p, n + 1);
        cap_ints(&g_cap.Si, (const int *)A->i, ((const int *)A->p)[n]);
        cap_dbls(&g_cap.Sx, (const double *)A->x, ((const int *)A->p)[n]);
    }
    return 1;
}

cholmod_dense *cholmod_solve(int sys, cholmod_factor *L, cholmod_dense *B, cholmod_common *Common)
{
    double t0 = now_s();
    shim_factor_t *F = (shim_factor_t *)L->z;
    int n = F->n;
    cholmod_dense *X = cholmod_zeros(B->nrow, B->ncol, CHOLMOD_REAL, Common);
    if (sys != CHOLMOD_A) { fprintf(stderr, "cholmod_shim: only CHOLMOD_A supported\n"); return X; }
    double *y = (double *)malloc((n ? n : 1) * sizeof(double));
    for (size_t col = 0; col < B->ncol; col++) {
        const double *b = (const double *)B->x + col * B->d;
        double *x = (double *)X->x + col * X->d;
        for (int k = 0; k < n; k++) y[k] = b[F->perm[k]];
        for (int j = 0; j < n; j++) {                     /* L y = Pb */
            y[j] /= F->Lx[F->Lp[j]];
            for (int p = F->Lp[j] + 1; p < F->Lp[j + 1]; p++) y[F->Li[p]] -= F->Lx[p] * y[j];
        }
        for (int j = n - 1; j >= 0; j--) {                /* L' z = y */
            for (int p = F->Lp[j] + 1; p < F->Lp[j + 1]; p++) y[j] -= F->Lx[p] * y[F->Li[p]];
            y[j] /= F->Lx[F->Lp[j]];
        }
        for (int k = 0; k < n; k++) x[F->perm[k]] = y[k];
    }
    free(y);
    g_cap.t_solve += now_s() - t0; g_cap.n_solve++;
    if (g_capture_enabled) {
        cap_dbls(&g_cap.b, (const double *)B->x, n);
        cap_dbls(&g_cap.x, (const double *)X->x, n);
    }
    return X;
}

int cholmod_free_factor(cholmod_factor **L, cholmod_common *Common)
{
    (void)Common;
    if (!L || !*L) return 1;
    shim_factor_t *F = (shim_factor_t *)(*L)->z;
    if (F) {
        free(F->perm); free(F->iperm); free(F->parent); free(F->Lp); free(F->Li); free(F->Lx);
        free(F->Cp); free(F->Ci); free(F->Cmap); free(F);
    }
    free(*L); *L = NULL;
    return 1;
}

int cholmod_free_sparse(cholmod_sparse **A, cholmod_common *Common)
{
    (void)Common;
    if (!A || !*A) return 1;
    free((*A)->p); free((*A)->i); free((*A)->nz); free((*A)->x); free(*A); *A = NULL;
    return 1;
}

int cholmod_free_dense(cholmod_dense **X, cholmod_common *Common)
{
    (void)Common;
    if (!X || !*X) return 1;
    free((*X)->x); free(*X); *X = NULL;
    return 1;
}
