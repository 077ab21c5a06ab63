// sort.cu -- K7: stable descending argsort of fp64 keys (sortperm(log_priorities, rev=true),
// reference src/resample.jl:156-157) and the key sort behind pf_coalesce! (resize.jl:309-334).
//
// STOPGAP (round 1): the radix passes are CUB's DeviceRadixSort (library code, like calling cuBLAS);
// this path is only taken for sort_particles=true and coalesce, never by the headline filter step.
// The key transform that reproduces Julia's `isless` total order (-0.0 < 0.0, stable ties by
// ascending index) is ours.
#include <cub/device/device_radix_sort.cuh>

#include "host.hpp"

namespace genpf {

__device__ __forceinline__ uint64_t order_bits(double x) {
    uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);  // monotone fp64 -> u64, -0.0 < +0.0
}
__device__ __forceinline__ double order_bits_inv(uint64_t t) {
    uint64_t b = (t >> 63) ? (t & 0x7FFFFFFFFFFFFFFFull) : ~t;
    return __longlong_as_double((long long)b);
}
__global__ void k_sort_prepare(const double *keys, int64_t n, uint64_t *k_out, int32_t *idx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        k_out[i] = ~order_bits(keys[i]);  // ascending radix order == descending key order
        idx[i] = (int32_t)i;
    }
}
__global__ void k_sort_finish(const uint64_t *k_sorted, int64_t n, double *keys_sorted) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys_sorted[i] = order_bits_inv(~k_sorted[i]);
}
__global__ void k_sort_prepare_i64(const int64_t *keys, int64_t n, uint64_t *k_out, int32_t *idx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        k_out[i] = (uint64_t)keys[i] ^ 0x8000000000000000ull;
        idx[i] = (int32_t)i;
    }
}
__global__ void k_sort_finish_i64(const uint64_t *k_sorted, int64_t n, int64_t *keys_sorted) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys_sorted[i] = (int64_t)(k_sorted[i] ^ 0x8000000000000000ull);
}

static int32_t sort_pairs_u64(uint64_t *k_in, uint64_t *k_out, int32_t *v_in, int32_t *v_out, int64_t n,
                              char *cub_tmp, size_t cub_bytes, cudaStream_t stream) {
    GENPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k_in, k_out, v_in, v_out, (int)n, 0, 64, stream));
    g_launches.fetch_add(8, std::memory_order_relaxed);
    return GENPF_OK;
}

static int32_t layout_tmp(int64_t n, DevBuf &tmp, uint64_t *&k_in, uint64_t *&k_out, int32_t *&v_in, char *&cub_tmp,
                          size_t &cub_bytes) {
    if (n > 0x7FFFFFFFll) return fail(GENPF_ERR_UNSUPPORTED, "sort: n must be < 2^31");
    cub_bytes = 0;
    GENPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                                   (int32_t *)nullptr, (int32_t *)nullptr, (int)n, 0, 64));
    size_t a = ((size_t)n * 8 + 255) & ~(size_t)255, b = ((size_t)n * 4 + 255) & ~(size_t)255;
    GENPF_TRY(tmp.ensure(2 * a + b + cub_bytes + 256));
    char *base = tmp.as<char>();
    k_in = (uint64_t *)base;
    k_out = (uint64_t *)(base + a);
    v_in = (int32_t *)(base + 2 * a);
    cub_tmp = base + 2 * a + b;
    return GENPF_OK;
}

int32_t sort_desc_stable(const double *keys, int64_t n, double *keys_sorted, int32_t *order32, DevBuf &tmp,
                         cudaStream_t stream) {
    uint64_t *k_in, *k_out;
    int32_t *v_in;
    char *cub_tmp;
    size_t cub_bytes;
    GENPF_TRY(layout_tmp(n, tmp, k_in, k_out, v_in, cub_tmp, cub_bytes));
    GENPF_LAUNCH(k_sort_prepare, grid_1d(n), 256, stream, keys, n, k_in, v_in);
    GENPF_TRY(sort_pairs_u64(k_in, k_out, v_in, order32, n, cub_tmp, cub_bytes, stream));
    if (keys_sorted) GENPF_LAUNCH(k_sort_finish, grid_1d(n), 256, stream, k_out, n, keys_sorted);
    return GENPF_OK;
}

int32_t sort_keys_i64(const int64_t *keys, int64_t n, int64_t *keys_sorted, int32_t *order32, DevBuf &tmp,
                      cudaStream_t stream) {
    uint64_t *k_in, *k_out;
    int32_t *v_in;
    char *cub_tmp;
    size_t cub_bytes;
    GENPF_TRY(layout_tmp(n, tmp, k_in, k_out, v_in, cub_tmp, cub_bytes));
    GENPF_LAUNCH(k_sort_prepare_i64, grid_1d(n), 256, stream, keys, n, k_in, v_in);
    GENPF_TRY(sort_pairs_u64(k_in, k_out, v_in, order32, n, cub_tmp, cub_bytes, stream));
    if (keys_sorted) GENPF_LAUNCH(k_sort_finish_i64, grid_1d(n), 256, stream, k_out, n, keys_sorted);
    return GENPF_OK;
}

}  // namespace genpf
