// Device-side plumbing: stream-ordered arenas, RAII scratch buffers, the solver context.
#pragma once
#include "common.h"
#include <cuda_runtime.h>
#include <memory>
#include <vector>
#include <map>
#include <chrono>
#include <string>

// Caching device allocator.  Every allocation / free of the library happens in program order on ONE
// stream, so a freed block may be handed out again immediately: the kernels that used it were
// enqueued before the kernels of its next owner.  Blocks are never returned to the driver while the
// context lives; a merge-tree solve repeats the same size sequence, so after the first solve no
// cudaMalloc is issued at all.  (cudaMallocAsync showed 2x run-to-run variation of the whole solve.)
struct DevicePool {
    // cached and live blocks carry the device they were allocated on: after lsfm_init() switches to
    // another GPU, blocks of the previous one are never handed out (and lsfm_shutdown drains the cache)
    struct Live { size_t bytes; int dev; };
    std::map<int, std::multimap<size_t, void *>> freeByDev_;
    std::map<void *, Live> live_;
    int dev_ = 0;                          // device of the current context
    size_t reserved = 0;
    void set_device(int d) { dev_ = d; }
    static size_t round_up(size_t b)
    {
        if (b < 4096) return (b + 511) & ~(size_t)511;
        if (b < (1u << 20)) return (b + 4095) & ~(size_t)4095;
        if (b < (1u << 26)) return (b + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
        return (b + (1u << 24) - 1) & ~(size_t)((1u << 24) - 1);
    }
    long long misses = 0;                    // allocations that went to cudaMalloc (LSFM_DEBUG)
    void *alloc(size_t bytes)
    {
        size_t need = round_up(bytes ? bytes : 1);
        std::multimap<size_t, void *> &free_ = freeByDev_[dev_];
        auto it = free_.lower_bound(need);
        // accept a cached block unless it wastes more than 25 % (+64 MB slack for the big arenas)
        if (it != free_.end() && it->first <= need + need / 4 + (need >= (1u << 26) ? (1u << 26) : 0)) {
            void *p = it->second;
            live_[p] = Live{it->first, dev_};
            free_.erase(it);
            return p;
        }
        void *p = nullptr;
        static const double limit_gb = getenv("LSFM_POOL_LIMIT_GB") ? atof(getenv("LSFM_POOL_LIMIT_GB")) : 0.0;
        cudaError_t e = cudaErrorMemoryAllocation;
        misses++;
        if (limit_gb <= 0.0 || (double)(reserved + need) <= limit_gb * 1e9) e = cudaMalloc(&p, need);
        if (e != cudaSuccess) {
            // out of memory: drop the cache and retry once
            cudaGetLastError();
            release_cached();
            e = cudaMalloc(&p, need);
            if (e != cudaSuccess)
                throw LsfmError(LSFM_ERR_CUDA, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        }
        reserved += need;
        live_[p] = Live{need, dev_};
        return p;
    }
    void free(void *p)
    {
        auto it = live_.find(p);
        if (it == live_.end()) return;
        freeByDev_[it->second.dev].insert({it->second.bytes, p});
        live_.erase(it);
    }
    void release_cached()
    {
        cudaDeviceSynchronize();
        for (auto &dv : freeByDev_) {
            for (auto &kv : dv.second) { cudaFree(kv.second); reserved -= kv.first; }
            dv.second.clear();
        }
    }
    static DevicePool &get()
    {
        static DevicePool pool;
        return pool;
    }
};

struct Arena {
    char *base = nullptr;
    size_t bytes = 0, used = 0;
    cudaStream_t stream = 0;
    Arena(size_t nbytes, cudaStream_t s) : bytes(nbytes ? nbytes : 256), stream(s)
    {
        base = (char *)DevicePool::get().alloc(bytes);
    }
    ~Arena()
    {
        if (base) DevicePool::get().free(base);
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
        p = (T *)DevicePool::get().alloc((count ? count : 1) * sizeof(T));
    }
    void release()
    {
        if (p) DevicePool::get().free(p);
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
    // Feature chunks of the chunk kernels (pattern / Schur / Transform W-V pass): ascending start indices
    // ending with n.  nullptr = 128 consecutive features per chunk.  Set by the join whose pattern stage had
    // to split chunks that saw more than 30 distinct poses (loop closures, dense overlap); Transforms keep it.
    std::shared_ptr<const std::vector<int>> chunkStarts;
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

    // persistent pinned host scratch for the per-level device->host hand-over (pattern keys, row
    // pointers): no per-level allocation / page faults, true asynchronous copies
    struct Pinned {
        char *p = nullptr;
        size_t cap = 0;
        ~Pinned() { if (p) cudaFreeHost(p); }
        char *need(size_t bytes)
        {
            if (bytes > cap) {
                if (p) cudaFreeHost(p);
                cap = bytes + bytes / 2 + 65536;
                CUDA_CHECK(cudaMallocHost((void **)&p, cap));
            }
            return p;
        }
    } pinDown;

    // largest dynamic shared-memory size already granted per kernel on THIS context's device
    std::map<const void *, size_t> smemGranted;
    void ensure_smem(const void *func, size_t bytes)
    {
        if (bytes <= 48 * 1024) return;
        size_t &g = smemGranted[func];
        if (bytes > g) {
            CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            g = bytes;
        }
    }

    // sticky device-side error flag (bit 0: non-positive pivot in the Cholesky); read by check_errors()
    int *d_err = nullptr;
    int *error_flag()
    {
        if (!d_err) {
            d_err = (int *)DevicePool::get().alloc(sizeof(int));
            CUDA_CHECK(cudaMemsetAsync(d_err, 0, sizeof(int), stream));
        }
        return d_err;
    }
    void release_error_flag()
    {
        if (d_err) DevicePool::get().free(d_err);
        d_err = nullptr;
    }
    void check_errors()
    {
        if (!d_err) return;
        int h = 0;
        CUDA_CHECK(cudaMemcpyAsync(&h, d_err, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        if (h) {
            CUDA_CHECK(cudaMemsetAsync(d_err, 0, sizeof(int), stream));
            if (h & 4)
                throw LsfmError(LSFM_ERR_CUDA, "Cholesky: a front waited for more than 2 s for its children (single-launch "
                                               "factorisation); rerun with LSFM_CHOL_LEVELS=1");
            throw LsfmError(LSFM_ERR_NOT_SPD, "reduced camera system is not positive definite");
        }
    }

    // host time between a size read-back (stream drained) and the next launch = GPU idle; LSFM_DEBUG
    double idle_ms[4] = {0, 0, 0, 0};
    std::chrono::steady_clock::time_point idle_t0;
    void idle_begin() { idle_t0 = std::chrono::steady_clock::now(); }
    void idle_end(int site)
    {
        idle_ms[site] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - idle_t0).count();
    }

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
