// plugin.hpp -- model registry (built-in + run-time compiled plugins) and typed dynamic kernel launches.
#pragma once
#include <mutex>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "filter_state.hpp"

namespace genpf {

constexpr int kPluginIdBase = 100;  // model ids >= 100 are plugins registered at run time
// kernel slots of a plugin image: k_propagate<M, Noise, INIT> (noise x {update, init}), k_mh<M, Noise, REWEIGHT>
// (noise x {accept, reweight}), k_step_fused<M, Noise, int, MH> (noise x MH in {1, 0, -1}), k_introduce<M, Noise> (image
// version 2); noise order Lean, Philox53, Cols
enum { kPlugProp = 0, kPlugMh = 6, kPlugFused = 12, kPlugIntro = 21, kPlugKernels = 24, kPlugKernelsV1 = 21 };
enum { kNzLean = 0, kNzPhilox53 = 1, kNzCols = 2 };

struct PluginModel {
    std::string name, name_store;
    ModelInfo info{};
    std::vector<char> cubin;
    std::string lowered[kPlugKernels];
    void *lib = nullptr;  // cudaLibrary_t, loaded on first use
    const void *kern[kPlugKernels] = {};
    std::mutex mu;
};

const ModelInfo *model_info(int32_t id);
PluginModel *plugin_model(int32_t id);
int32_t find_model(const char *name);
int32_t plugin_kernel(PluginModel *pm, int slot, const void **fn);

template <class Noise> struct NoiseIndex;
template <> struct NoiseIndex<NoiseLean> { static constexpr int value = kNzLean; };
template <> struct NoiseIndex<NoisePhilox53> { static constexpr int value = kNzPhilox53; };
template <> struct NoiseIndex<NoiseCols> { static constexpr int value = kNzCols; };

// Launch `fn` (a __global__ function of this library or a cudaKernel_t of a plugin image) with the parameter list
// of `sig`: every argument is converted to the exact parameter type first, as <<<>>> would.
// pdl: programmatic dependent launch (host.hpp::launch_pdl) -- only for kernels that start with pdl_enter()
template <typename... P, typename... A, size_t... I>
inline cudaError_t launch_typed_impl(void (*)(P...), const void *fn, dim3 grid, dim3 block, cudaStream_t s, bool pdl,
                                     std::index_sequence<I...>, A &&...a) {
    std::tuple<P...> args{static_cast<P>(std::forward<A>(a))...};
    void *argv[] = {(void *)&std::get<I>(args)...};
    if (!(pdl && g_pdl)) return cudaLaunchKernel(fn, grid, block, argv, 0, s);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, fn, argv);
}
template <typename... P, typename... A>
inline int32_t launch_typed(const char *name, void (*sig)(P...), const void *fn, dim3 grid, dim3 block, cudaStream_t s,
                            A &&...a) {
    static_assert(sizeof...(P) == sizeof...(A), "argument count differs from the kernel's parameter list");
    cudaEvent_t e0 = nullptr;
    if (g_prof_on) e0 = prof_mark(s);
    const bool pdl = name[0] == 'k' && name[2] == 's' && name[3] == 't';  // "k_step_fused": the one chain kernel launched here
    cudaError_t e = launch_typed_impl(sig, fn, grid, block, s, pdl, std::index_sequence_for<P...>{}, std::forward<A>(a)...);
    if (g_prof_on) prof_push(name, e0, prof_mark(s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(GENPF_ERR_CUDA, std::string("plugin kernel launch failed: ") + cudaGetErrorString(e));
    }
    GENPF_CUDA_TRY(cudaGetLastError());
    return GENPF_OK;
}

}  // namespace genpf
