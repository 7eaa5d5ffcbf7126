// Common types for the B200-native LinearSFM hot path (device map descriptors, error handling).
//
// Device data layout ("map descriptor" = DMap): the reference keeps one heap object per array of
// LocalMapInfoStereo (LinearSFMImp.h:75-121).  Here every map of a merge-tree level lives inside a
// few large device arenas, and a DMap is a POD of pointers + sizes into them, so that one kernel
// launch can process every map / every join of a level ("segmented" launches).
//   poseNo[m]   = stno of the pose's six rows (<=0, LinearSFMImp.cpp:423)      int32
//   poseVal[6m] = x y z alpha beta gamma per pose                                f64
//   featNo[n]   = stno of the feature's three rows (>0)                          int32
//   featVal[3n]                                                                  f64
//   U[36 nU] row-major 6x6, Ui/Uj[nU] (Ui<=Uj)                                   f64 / int32
//   W[18 nW] row-major 6x3, photo/feature[nW] (feature non-decreasing)           f64 / int32
//   V[9 n] row-major 3x3                                                         f64
//   wPtr[n+1]   = CSR start of every feature's W blocks; the reference's FBlock[f]
//                 (LinearSFMImp.cpp:1304, 2855-2858) is wPtr[f] if the feature has blocks else -1.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#define LSFM_OK 0
#define LSFM_ERR_CUDA 1
#define LSFM_ERR_ARG 2
#define LSFM_ERR_REF_NOT_FOUND 3     // Transform: pose -Ref not in the state (reference: UB)
#define LSFM_ERR_NOT_SPD 4
#define LSFM_ERR_IO 5
#define LSFM_ERR_FORMAT 6
#define LSFM_ERR_NO_DEVICE 7

struct LsfmError : public std::runtime_error {
    int code;
    LsfmError(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define CUDA_CHECK(expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            throw LsfmError(LSFM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) +  \
                                               " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)
#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())
#endif

struct DMap {
    int Ref, FRef, m, n, nU, nW;
    int *poseNo;
    double *poseVal;
    int *featNo;
    double *featVal;
    double *U;
    int *Ui, *Uj;
    double *W;
    int *photo, *feature;
    double *V;
    int *wPtr;
    // monocular maps only (LocalMapInfo, LinearSFMImp.h:172-176)
    int ScaP, Fix, Sign, FScaP, FFix;
};

static inline int ceil_div(long long a, int b) { return (int)((a + b - 1) / b); }
