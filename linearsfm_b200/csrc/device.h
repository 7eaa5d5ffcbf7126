// Device-side plumbing: stream-ordered arenas, RAII scratch buffers, the solver context.
#pragma once
#include "common.h"
#include <cuda_runtime.h>
#include <memory>
#include <vector>
#include <map>
#include <string>

struct Arena {
    char *base = nullptr;
    size_t bytes = 0, used = 0;
    cudaStream_t stream = 0;
    Arena(size_t nbytes, cudaStream_t s) : bytes(nbytes ? nbytes : 256), stream(s)
    {
        CUDA_CHECK(cudaMallocAsync((void **)&base, bytes, stream));
    }
    ~Arena()
    {
        if (base) cudaFreeAsync(base, stream);
    }
    Arena(const Arena &) = delete;
    Arena &operator=(const Arena &) = delete;
    static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
    template <class T> T *take(size_t count)
    {
        size_t b = pad(count * sizeof(T));
        if (used + b > bytes) throw LsfmError(LSFM_ERR_ARG, "arena overflow");
        T *p = (T *)(base + used);
        used += b;
        return p;
    }
};

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t stream = 0;
    DevBuf() {}
    DevBuf(size_t count, cudaStream_t s) { alloc(count, s); }
    void alloc(size_t count, cudaStream_t s)
    {
        release();
        n = count; stream = s;
        CUDA_CHECK(cudaMallocAsync((void **)&p, (count ? count : 1) * sizeof(T), s));
    }
    void release()
    {
        if (p) cudaFreeAsync(p, stream);
        p = nullptr; n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), stream(o.stream) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; n = o.n; stream = o.stream; o.p = nullptr; o.n = 0; }
        return *this;
    }
    void zero() { CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), stream)); }
    void upload(const T *h, size_t count)
    {
        CUDA_CHECK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, stream));
    }
    void upload(const std::vector<T> &h) { upload(h.data(), h.size()); }
    void download(T *h, size_t count) const
    {
        CUDA_CHECK(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, stream));
    }
    std::vector<T> to_host() const
    {
        std::vector<T> h(n);
        download(h.data(), n);
        CUDA_CHECK(cudaStreamSynchronize(stream));
        return h;
    }
};

// A map living on the device: descriptor + the arena that owns its storage.
struct MapHandle {
    DMap d;
    std::shared_ptr<Arena> arena;
};

// Per-stage accounting (kernel launches, device time, algorithmic bytes / flops).
struct StageStat {
    double ms = 0.0;
    long long launches = 0;
    double bytes = 0.0;
    double flops = 0.0;
};

struct Context {
    int device = 0;
    cudaStream_t stream = 0;
    bool timing = false;                     // per-stage CUDA-event timing (adds syncs)
    std::map<std::string, StageStat> stats;
    long long launches = 0;
    std::vector<double> objectives;          // optional per-join objective values
    bool want_objective = false;
    int num_sms = 148;

    // event pool for stage timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string cur_stage;

    void begin(const char *stage)
    {
        if (!timing) return;
        cur_stage = stage;
        CUDA_CHECK(cudaEventRecord(ev0, stream));
    }
    void end(double bytes = 0.0, double flops = 0.0, int nlaunch = 0)
    {
        launches += nlaunch;
        if (!timing) return;
        CUDA_CHECK(cudaEventRecord(ev1, stream));
        CUDA_CHECK(cudaEventSynchronize(ev1));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
        StageStat &s = stats[cur_stage];
        s.ms += ms; s.bytes += bytes; s.flops += flops; s.launches += nlaunch;
    }
};

// binary search: largest k with pre[k] <= g  (pre is a K+1 prefix-sum array)
__host__ __device__ static inline int seg_find(const int *pre, int K, int g)
{
    int lo = 0, hi = K;            // invariant: pre[lo] <= g < pre[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (pre[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}
