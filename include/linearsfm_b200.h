/*
 * linearsfm_b200.h -- C ABI of the B200-native LinearSFM hot path (liblinearsfm_b200.so).
 *
 * The reference (LiangZhaoPKUImperial/LinearSFM) has no FFI; its in-process boundary is the public
 * member functions of class CLinearSFMImp (linux/src/LinearSFMImp/LinearSFMImp.h:181-253) working
 * on the POD containers LocalMapInfoStereo / LocalMapInfo (LinearSFMImp.h:75-178).  Every entry
 * point below names the reference interface it replaces.  All pointers are HOST pointers, sizes
 * are plain ints, no C++/torch types cross the boundary.  Device residency is hidden behind the
 * opaque lsfm_tree handle.  Every function returns an int status (LSFM_OK = 0); the text of the
 * last error is available from lsfm_last_error().  There is NO CPU fallback: without a CUDA
 * device every compute entry returns LSFM_ERR_NO_DEVICE.
 *
 * Threading: one context per process, not re-entrant (same as the reference, which keeps scratch
 * in members m_sparseS/m_factorS/..., LinearSFMImp.h:244-247).
 */
#ifndef LINEARSFM_B200_H
#define LINEARSFM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSFM_OK 0
#define LSFM_ERR_CUDA 1
#define LSFM_ERR_ARG 2
#define LSFM_ERR_REF_NOT_FOUND 3
#define LSFM_ERR_NOT_SPD 4
#define LSFM_ERR_IO 5
#define LSFM_ERR_FORMAT 6
#define LSFM_ERR_NO_DEVICE 7

/* Field-for-field mirror of LocalMapInfoStereo (LinearSFMImp.h:75-121) and LocalMapInfo
 * (LinearSFMImp.h:124-178; the mono-only fields are ignored by the stereo entry points).
 * Arrays are malloc()'d by the producer and released with lsfm_free_map() (the reference's
 * ownership convention: producer mallocs, consumer frees).
 *   stno[r], stVal[r]  r = 6m+3n; pose rows carry -poseID (x6), feature rows +featID (x3)
 *   U[36 nU] row-major 6x6, Ui/Uj[nU] with Ui<=Uj;  W[18 nW] row-major 6x3, photo/feature[nW]
 *   (feature non-decreasing);  V[9 n];  FBlock[n] = first W block of the feature or -1.        */
typedef struct lsfm_map {
    int Ref, FRef, r, m, n, nU, nW;
    int ScaP, Fix, Sign, FScaP, FFix;
    int *stno;
    double *stVal;
    double *U;
    int *Ui, *Uj;
    double *W;
    int *photo, *feature;
    double *V;
    int *FBlock;
} lsfm_map;

typedef struct lsfm_tree lsfm_tree; /* opaque: leaf maps resident in HBM + the last result */

/* ---- context ------------------------------------------------------------------------------ */
int lsfm_init(int device);                 /* replaces CLinearSFMImp::CLinearSFMImp (Imp.cpp:80-87) */
void lsfm_shutdown(void);                  /* replaces ~CLinearSFMImp (Imp.cpp:90-93)               */
const char *lsfm_last_error(void);
int lsfm_device_count(void);
void lsfm_free_map(lsfm_map *m);
/* per-stage statistics of the calls since the last reset, as a JSON object string */
void lsfm_stats_reset(int enable_stage_timing);
const char *lsfm_stats_json(void);

/* ---- operators (host buffers in / host buffers out) ---------------------------------------- */
/* void lmj_Transform_PF3DStereo(LocalMapInfoStereo& out, int Ref) with m_GMapS = *in
 * (LinearSFMImp.h:206, LinearSFMImp.cpp:349-1924).  Ref == in->Ref returns a copy.            */
int lsfm_transform_stereo(const lsfm_map *in, int Ref, lsfm_map *out);
/* K independent transforms in one segmented launch set (what the scheduler uses per level).   */
int lsfm_transform_stereo_batch(const lsfm_map *in, const int *Ref, int K, lsfm_map *out);

/* void lmj_Transform_PF3DMono(LocalMapInfo& out, int Ref, int ScaP, int Fix) with m_GMap = *in
 * (LinearSFMImp.h:216, LinearSFMImp.cpp:3173-6509).  Same Ref and ScaP returns a copy (3176).     */
int lsfm_transform_mono(const lsfm_map *in, int Ref, int ScaP, int Fix, lsfm_map *out);
int lsfm_transform_mono_batch(const lsfm_map *in, const int *Ref, const int *ScaP, const int *Fix, int K,
                              lsfm_map *out);

/* void lmj_LinearLS_PF3DStereo(LocalMapInfoStereo& End, LocalMapInfoStereo& Cur), result in
 * m_GMapS (LinearSFMImp.h:207, LinearSFMImp.cpp:2551-2978).  Inputs are NOT freed here.       */
int lsfm_join_stereo(const lsfm_map *end, const lsfm_map *cur, lsfm_map *out);
int lsfm_join_stereo_batch(const lsfm_map *end, const lsfm_map *cur, int K, lsfm_map *out);

/* void lmj_LinearLS_PF3DMono(LocalMapInfo& End, LocalMapInfo& Cur), result in m_GMap
 * (LinearSFMImp.h:219, LinearSFMImp.cpp:7282-7874) incl. lmj_solveLinearSFMMono (6756-7041).      */
int lsfm_join_mono(const lsfm_map *end, const lsfm_map *cur, lsfm_map *out);
int lsfm_join_mono_batch(const lsfm_map *end, const lsfm_map *cur, int K, lsfm_map *out);
/* void lmj_PF3D_Divide_ConquerMono(int nLocalMapCount) over m_LMset (LinearSFMImp.cpp:6511-6658)    */
int lsfm_run_mono(const lsfm_map *maps, int num, lsfm_map *out);
int lsfm_run_mono_ex(const lsfm_map *maps, int num, int verbose, lsfm_map *out);   /* verbose: reference stdout lines */

/* void lmj_solveLinearSFMStereo(double* stVal, double* eb, double* ea, double* U, double* W,
 *      double* V, int* Ui, int* Uj, int* photo, int* feature, int m, int n, int nU, int nW)
 * (LinearSFMImp.h:209, LinearSFMImp.cpp:2119-2378): same argument order and meaning; stVal[6m+3n]
 * is caller-allocated output; V is left untouched (the reference overwrites and restores it).  */
int lsfm_solve_stereo(double *stVal, const double *eb, const double *ea, const double *U,
                      const double *W, const double *V, const int *Ui, const int *Uj,
                      const int *photo, const int *feature, int m, int n, int nU, int nW);

/* void lmj_solveLinearSFMMono(double* stVal, double* eb, double* ea, double* U, double* W, double* V,
 *      int* Ui, int* Uj, int* photo, int* feature, int m, int n, int nU, int nW,
 *      int Ref, int ScaP, int Fix, int Sign, int FixBlk)     (LinearSFMImp.h:223, LinearSFMImp.cpp:6756-7041)
 * Same argument order and meaning: Ref = block index of the all-zero pose, ScaP = its first scalar row
 * (= 6 Ref; rows ScaP..ScaP+5 are removed from the system, 6981-7001), Fix = scalar row of the pinned
 * translation component (also removed), Sign = its value written back at the end (7026).  FixBlk only
 * steers the reference's scalar permutation and is ignored.  Here the gauge rows stay in the block
 * system as identity rows with a zero right-hand side (same solution).          */
int lsfm_solve_mono(double *stVal, const double *eb, const double *ea, const double *U, const double *W,
                    const double *V, const int *Ui, const int *Uj, const int *photo, const int *feature,
                    int m, int n, int nU, int nW, int Ref, int ScaP, int Fix, int Sign, int FixBlk);

/* Debug capture of the last lsfm_solve_stereo / single join (integer parity tests):
 * block CRS of S (the reference's Sidxij, LinearSFMImp.cpp:2190-2205), block elimination ordering
 * (replaces cholmod_amd, LinearSFMImp.cpp:2413).  Pointers stay valid until the next solve.    */
int lsfm_debug_last_solve(int *m, const int **rowptr, const int **colidx, const double **S,
                          const double **E, const int **perm);
/* Block-Jacobi preconditioned conjugate gradients on S x = E (north_star: "block-Jacobi PCG as a
 * cross-check" of the Cholesky that replaces pba_solveCholmodLM, LinearSFMImp.cpp:2380-2449).  S is the
 * upper block triangle in block CRS exactly as lsfm_debug_last_solve() returns it (36 doubles per
 * block, row-major, diagonal blocks full).  Stops at ||r|| <= tol ||E|| or after max_iters.        */
int lsfm_pcg_block(int m, const int *rowptr, const int *colidx, const double *S, const double *E,
                   double *x, double tol, int max_iters, int *iters_done, double *rel_residual);
/* the ordering routine alone: upper block pattern (CSC: Ap[m+1], Ai) -> perm[m]               */
int lsfm_block_ordering(int m, const int *Ap, const int *Ai, int *perm);

/* ---- whole merge tree ----------------------------------------------------------------------- */
/* void lmj_PF3D_Divide_ConquerStereo(int nLocalMapCount) over m_LMsetS (LinearSFMImp.cpp:1926-2099)
 * with host maps in, final map out (H2D of the leaves and D2H of the result inside the call).   */
int lsfm_run_stereo(const lsfm_map *maps, int num, lsfm_map *out);

/* Resident variant: upload once, solve many times (bench), download state or full map.
 * first_index = global index of maps[0] in the full sequence (multi-GPU sharding: the re-base
 * rule of LinearSFMImp.cpp:1997 depends on the global output index).                           */
int lsfm_tree_create_stereo(const lsfm_map *maps, int num, lsfm_tree **tree);
int lsfm_tree_solve(lsfm_tree *tree, int verbose, int first_index, int max_levels);
/* device time (ms, CUDA events on the library's stream) of the last lsfm_tree_solve            */
double lsfm_tree_last_solve_ms(const lsfm_tree *tree);
/* make the result of the last solve the new input set (no copy); append uploaded host maps     */
int lsfm_tree_adopt_result(lsfm_tree *tree);
/* input set := the maps of the last create/set_maps (still resident; no copy)                  */
int lsfm_tree_reset(lsfm_tree *tree);
int lsfm_tree_append_maps(lsfm_tree *tree, const lsfm_map *maps, int num);
int lsfm_tree_result_count(const lsfm_tree *tree);
int lsfm_tree_result_shape(const lsfm_tree *tree, int idx, lsfm_map *shape_only);
int lsfm_tree_download(const lsfm_tree *tree, int idx, lsfm_map *out);
int lsfm_tree_download_state(const lsfm_tree *tree, int idx, int *stno, double *stVal);
/* Marginal covariances of selected state blocks of a map (the reference frees the final
 * information matrix unused, LinearSFMImp.cpp:2081-2096).  sel[i] in [0, m) selects pose block i, in
 * [m, m + n) feature sel[i] - m; cov receives, in the order of sel, the row-major 6x6 (pose) or 3x3
 * (feature) diagonal block of the inverse of the map's information matrix [[U, W], [W^T, V]]: 36 or 9
 * doubles per selected block.  Every column is one unit right-hand side through the joint solver (Schur
 * complement + multifrontal Cholesky + back-substitution), up to 48 columns per segmented batch.
 * lsfm_tree_marginal_cov works on result `idx` of a solved tree without leaving HBM. */
int lsfm_marginal_cov_stereo(const lsfm_map *map, int nsel, const int *sel, double *cov);
int lsfm_tree_marginal_cov(const lsfm_tree *tree, int idx, int nsel, const int *sel, double *cov);
/* Device-to-device hand-over of a whole map between ranks (multi-GPU tree levels): a map is
 * packed into / unpacked from ONE contiguous device buffer whose layout depends only on its shape
 * (m, n, nU, nW), so the buffer can be moved by NCCL send/recv or a peer copy as raw bytes.       */
size_t lsfm_map_device_bytes(const lsfm_map *shape);
int lsfm_tree_export_device(const lsfm_tree *tree, int idx, void *dst_device, size_t bytes);
int lsfm_tree_append_device(lsfm_tree *tree, const lsfm_map *shape, const void *src_device, size_t bytes);
/* replace leaf/result set by host maps (multi-GPU hand-over between ranks)                      */
int lsfm_tree_set_maps(lsfm_tree *tree, const lsfm_map *maps, int num);
void lsfm_tree_free(lsfm_tree *tree);

/* ---- files + CLI (drop-in process boundary) ------------------------------------------------- */
/* lmj_readInformationStereo (LinearSFMImp.cpp:3044-3132)                                        */
int lsfm_load_localmap_stereo(const char *path, lsfm_map *out);
/* lmj_readInformationMono (LinearSFMImp.cpp:6660-6754): header `Ref ScaP Fix Sign r`           */
int lsfm_load_localmap_mono(const char *path, lsfm_map *out);
/* lmj_SaveStateVector (2102-2117) and lmj_SavePoses_3DPF (7876-7967); NULL = skip              */
/* Writes a map (state + block information) in the reference's localmap_<i>.txt format (the format
 * lmj_readInformationStereo / lmj_readInformationMono parse, LinearSFMImp.cpp:3044-3132, 6660-6754),
 * doubles with 17 significant digits.  The reference has no writer for it (it frees the joined map,
 * LinearSFMImp.cpp:2081-2096); this is SURVEY 8(f)-3: the final information matrix as an output, so
 * that a joined map can be joined again later.  CLI: -map <file>.                                  */
int lsfm_save_localmap(const lsfm_map *m, const char *path, int mono);
int lsfm_save_outputs(const lsfm_map *m, const char *state_path, const char *pose_path,
                      const char *feature_path);
/* Binary cache of parsed input maps: the reference re-parses every localmap_<i>.txt with one
 * fscanf per number on each run (LinearSFMImp.cpp:3062-3128, 6678-6750); lsfm_save_cache writes `num` maps as raw
 * arrays into ONE file, lsfm_load_cache reads them back bit for bit (*maps_out: calloc'd array of `*num_out`
 * maps whose arrays are malloc'd; release with lsfm_free_cache).  No reference interface is replaced.
 * CLI: -cache <file> (used when present and matching -num / -type, otherwise written after the text parse). */
int lsfm_save_cache(const char *file, const lsfm_map *maps, int num, int mono);
int lsfm_load_cache(const char *file, lsfm_map **maps_out, int *num_out, int *mono_out);
void lsfm_free_cache(lsfm_map *maps, int num);
/* CLinearSFMImp::run(argc, argv) (LinearSFMImp.cpp:7972-8106): same flags
 *   -path <dir> -num <N> -type {Monocular|Stereo} [-p <pose>] [-f <feature>] [-st <state>] [-help] */
int lsfm_cli_main(int argc, char **argv);

/* ---- local-map builder -----------------------------------------------------
 * The step UPSTREAM of the reference: LinearSFM starts from finished localmap_*.txt files (DOC p.1,
 * "initial reconstructions"), there is no reference interface to replace.  One local map = two
 * consecutive stereo frames; conventions follow LinearSFMImp.cpp:132-143 (X_cam = R (X - t),
 * R = Rx(gamma) Ry(beta) Rz(alpha)) and SURVEY Appendix A: map frame = first frame, state = pose of
 * the second frame + landmarks, camera along +x (y left, z up), right camera at y = -baseline,
 * measurement (uL, vL, uR) = (cx - f y/x, cy - f z/x, cx - f (y+b)/x).                           */
typedef struct lsfm_stereo_pair {
    int Ref;               /* pose id of the first frame (the map's frame; becomes lsfm_map.Ref)   */
    int pose_id;           /* pose id of the second frame (state rows carry -pose_id)              */
    int n;                 /* landmarks observed in both frames                                     */
    const int *feat_id;    /* [n]  landmark ids (> 0)                                               */
    const double *z0;      /* [3n] (uL, vL, uR) in the first frame                                  */
    const double *z1;      /* [3n] (uL, vL, uR) in the second frame                                 */
    const double *pose0;   /* [6]  initial guess of the second frame's pose in the first frame      */
    const double *X0;      /* [3n] initial landmark positions, or NULL: triangulated from z0        */
} lsfm_stereo_pair;
typedef struct lsfm_stereo_cam { double f, baseline, cx, cy, sigma; } lsfm_stereo_cam;
/* Gauss-Newton two-view bundle adjustment of every pair (one CTA per pair, one launch), stopping
 * when max|d pose| < tol or after max_iters; out[k] is the local map (m = 1, nU = 1, nW = n) with
 * the information blocks evaluated at the estimate.  iters_done[K] may be NULL.                   */
int lsfm_build_localmaps_stereo(const lsfm_stereo_pair *pairs, int num, const lsfm_stereo_cam *cam,
                                int max_iters, double tol, lsfm_map *out, int *iters_done);

#ifdef __cplusplus
}
#endif
#endif
