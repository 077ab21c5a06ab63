// host.hpp -- host-side plumbing shared by the ABI translation units: errors, launch counting,
// grow-only device buffers.  No torch types anywhere: the library only needs the CUDA runtime.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/genpf.h"

namespace genpf {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

inline int32_t fail(int32_t code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

#define GENPF_CUDA_TRY(expr)                                                                                 \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) {                                                                             \
            char _b[512];                                                                                    \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, \
                     #expr);                                                                                 \
            cudaGetLastError();                                                                              \
            return ::genpf::fail(GENPF_ERR_CUDA, _b);                                                        \
        }                                                                                                    \
    } while (0)

#define GENPF_TRY(expr)               \
    do {                              \
        int32_t _s = (expr);          \
        if (_s != GENPF_OK) return _s; \
    } while (0)

// optional per-kernel CUDA-event timing (genpf_profile_begin/end): events are recorded on the launching
// stream around every kernel, so the durations are device times of exactly the launches of the timed region
struct ProfRec {
    const char *name;
    cudaEvent_t e0, e1;
};
extern bool g_prof_on;
extern std::vector<ProfRec> g_prof;
extern std::mutex g_prof_mu;  // launches may come from several host threads (one workspace / filter each)
inline cudaEvent_t prof_mark(cudaStream_t s) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    return e;
}
inline void prof_push(const char *name, cudaEvent_t e0, cudaEvent_t e1) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back({name, e0, e1});
}

// kernel launch + count (bench.py's gpu_launches) + launch-error check
#define GENPF_LAUNCH(kernel, grid, block, stream, ...) GENPF_LAUNCH_SMEM(kernel, grid, block, 0, stream, __VA_ARGS__)
#define GENPF_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)            \
    do {                                                                     \
        cudaEvent_t _e0 = nullptr;                                           \
        if (::genpf::g_prof_on) _e0 = ::genpf::prof_mark(stream);            \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);          \
        if (::genpf::g_prof_on) ::genpf::prof_push(#kernel, _e0, ::genpf::prof_mark(stream)); \
        ::genpf::g_launches.fetch_add(1, std::memory_order_relaxed);         \
        GENPF_CUDA_TRY(cudaGetLastError());                                  \
    } while (0)

// Programmatic dependent launch for the kernel chain of a step (finalize -> combine -> scan -> fused / push ->
// boundary): the next kernel's blocks are scheduled while the previous kernel's last wave drains and park at
// griddepcontrol.wait (common.cuh::pdl_enter) until it has completed and flushed -- same dependencies, but the
// launch latency and the ramp of each kernel are hidden (4 boundaries per step).  Only kernels that call pdl_enter()
// first thing may be launched this way.  GENPF_PDL=0 in the environment falls back to plain launches.
extern bool g_pdl;
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A &&...a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(std::forward<A>(a))...);
}
#define GENPF_LAUNCH_PDL(kernel, grid, block, stream, ...)                                                     \
    do {                                                                                                       \
        cudaEvent_t _e0 = nullptr;                                                                             \
        if (::genpf::g_prof_on) _e0 = ::genpf::prof_mark(stream);                                              \
        cudaError_t _le = ::genpf::launch_pdl(kernel, dim3(grid), dim3(block), 0, (stream), __VA_ARGS__);      \
        if (::genpf::g_prof_on) ::genpf::prof_push(#kernel, _e0, ::genpf::prof_mark(stream));                  \
        ::genpf::g_launches.fetch_add(1, std::memory_order_relaxed);                                           \
        GENPF_CUDA_TRY(_le);                                                                                   \
        GENPF_CUDA_TRY(cudaGetLastError());                                                                    \
    } while (0)

// grow-only device allocation
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int32_t ensure(size_t bytes) {
        if (bytes <= cap) return GENPF_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return fail(GENPF_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        }
        cap = want;
        return GENPF_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline unsigned grid_1d(int64_t n, int block = 256, int64_t max_blocks = 148 * 16) {
    int64_t g = ceil_div(n, block);
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (unsigned)g;
}

// stable descending sort of fp64 keys with Julia `isless` ordering (sort.cu; K7).
// order32[k] = original index of the k-th largest key; ties keep ascending original index.
int32_t sort_desc_stable(const double *keys, int64_t n, double *keys_sorted, int32_t *order32, DevBuf &tmp,
                         cudaStream_t stream);
// the same for every filter of a batch (filter f at [f*n, (f+1)*n)): one launch for small filters (n <= 4096)
int32_t sort_desc_stable_batched(const double *keys, int64_t n, int64_t nf, double *keys_sorted, int32_t *order32,
                                 DevBuf &tmp, cudaStream_t stream);
// stable ascending sort of int64 keys with index payload (coalesce)
int32_t sort_keys_i64(const int64_t *keys, int64_t n, int64_t *keys_sorted, int32_t *order32, DevBuf &tmp,
                      cudaStream_t stream);

}  // namespace genpf
