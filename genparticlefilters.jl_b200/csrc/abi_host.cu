// abi_host.cu -- C ABI, host-array path: the caller (Julia via ccall) keeps its traces and hands over
// only log_weights / uniforms; ancestors and new weights come back.  See include/genpf.h for the
// reference function each entry point replaces.  No CPU fallback: every path launches CUDA kernels.
#include <cstdlib>
#include <cstring>

#include "coalesce.cuh"
#include "optimal.cuh"
#include "engine.cuh"

namespace genpf {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
bool g_prof_on = false;
std::mutex g_prof_mu;
bool g_pdl = []() { const char *e = getenv("GENPF_PDL"); return !(e && e[0] == '0'); }();
std::vector<ProfRec> g_prof;

// per-host-thread workspace: one stream, staging buffers, scratch
struct HostWs {
    cudaStream_t stream = nullptr;
    Scratch sc;
    DevBuf lw, lp, u, parents, lw_out, x, keys, aux1, aux2, aux3;
    CoalesceBufs cb;
    OptimalBufs ob;
    OptCtrl *h_ctrl = nullptr;  // pinned
    Stats *h_stats = nullptr;  // pinned, 4 entries
    double *h_scalars = nullptr;  // pinned, 8 doubles
    int device = -1;
    int32_t init() {
        int dev = 0;
        GENPF_CUDA_TRY(cudaGetDevice(&dev));
        if (stream && dev == device) return GENPF_OK;
        if (stream) destroy();
        device = dev;
        GENPF_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        GENPF_CUDA_TRY(cudaMallocHost(&h_stats, sizeof(Stats) * 4));
        GENPF_CUDA_TRY(cudaMallocHost(&h_scalars, sizeof(double) * 8));
        GENPF_CUDA_TRY(cudaMallocHost(&h_ctrl, sizeof(OptCtrl)));
        return GENPF_OK;
    }
    void destroy() {
        sc.release();
        for (DevBuf *b : {&lw, &lp, &u, &parents, &lw_out, &x, &keys, &aux1, &aux2, &aux3}) b->release();
        if (h_stats) cudaFreeHost(h_stats);
        if (h_scalars) cudaFreeHost(h_scalars);
        if (h_ctrl) cudaFreeHost(h_ctrl);
        h_ctrl = nullptr;
        ob.release();
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
        h_stats = nullptr;
        h_scalars = nullptr;
    }
    ~HostWs() {}  // process teardown: leave to the driver (CUDA context may already be gone)
};
static thread_local HostWs g_ws;

// stage an input array: returns the device pointer to use
template <typename T>
static int32_t stage_in(HostWs &ws, DevBuf &buf, const T *src, int64_t n, bool device_ptrs, const T **out) {
    if (!src) {
        *out = nullptr;
        return GENPF_OK;
    }
    if (device_ptrs) {
        *out = src;
        return GENPF_OK;
    }
    GENPF_TRY(buf.ensure((size_t)n * sizeof(T)));
    GENPF_CUDA_TRY(cudaMemcpyAsync(buf.p, src, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, ws.stream));
    *out = buf.as<T>();
    return GENPF_OK;
}
template <typename T>
static int32_t stage_out(DevBuf &buf, T *dst, int64_t n, bool device_ptrs, T **out) {
    if (!dst) {
        *out = nullptr;
        return GENPF_OK;
    }
    if (device_ptrs) {
        *out = dst;
        return GENPF_OK;
    }
    GENPF_TRY(buf.ensure((size_t)n * sizeof(T)));
    *out = buf.as<T>();
    return GENPF_OK;
}
template <typename T>
static int32_t copy_out(HostWs &ws, const T *dev, T *dst, int64_t n, bool device_ptrs) {
    if (!dst || device_ptrs) return GENPF_OK;
    GENPF_CUDA_TRY(cudaMemcpyAsync(dst, dev, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, ws.stream));
    return GENPF_OK;
}

static int32_t reduce_to_host(HostWs &ws, const double *d_lw, int64_t n, Stats *out) {
    GENPF_TRY(ws.sc.ensure(n, 1));
    LwSrc src{d_lw, 1.0};
    GENPF_TRY(launch_reduce(ws.stream, src, n, 1, ws.sc.partials(0)));
    GENPF_TRY(launch_finalize(ws.stream, ws.sc, ws.sc.partials(0), n, 1, ws.sc.st(0, 1), nullptr, -1.0, nullptr));
    GENPF_CUDA_TRY(cudaMemcpyAsync(ws.h_stats, ws.sc.st(0, 1), sizeof(Stats), cudaMemcpyDeviceToHost, ws.stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    *out = ws.h_stats[0];
    return GENPF_OK;
}

}  // namespace genpf

using namespace genpf;

extern "C" {

int32_t genpf_version(void) { return GENPF_VERSION; }
const char *genpf_last_error(void) { return g_last_error.c_str(); }
int64_t genpf_launch_count(void) { return g_launches.load(); }

// per-kernel device timing of everything launched between begin and end (single host thread)
int32_t genpf_profile_begin(void) {
    for (auto &r : g_prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    g_prof_on = true;
    return GENPF_OK;
}
// writes "name count total_ms\n" lines into buf (NUL terminated); returns GENPF_OK
int32_t genpf_profile_end(char *buf, int64_t buf_len) {
    g_prof_on = false;
    GENPF_CUDA_TRY(cudaDeviceSynchronize());
    std::vector<std::string> names;
    std::vector<double> total;
    std::vector<int64_t> count;
    for (auto &r : g_prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        size_t k = 0;
        for (; k < names.size(); ++k)
            if (names[k] == r.name) break;
        if (k == names.size()) {
            names.push_back(r.name);
            total.push_back(0.0);
            count.push_back(0);
        }
        total[k] += ms;
        count[k] += 1;
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    std::string out;
    for (size_t k = 0; k < names.size(); ++k) {
        char line[256];
        snprintf(line, sizeof(line), "%s\t%lld\t%.6f\n", names[k].c_str(), (long long)count[k], total[k]);
        out += line;
    }
    if (buf && buf_len > 0) {
        size_t m = out.size() < (size_t)buf_len - 1 ? out.size() : (size_t)buf_len - 1;
        memcpy(buf, out.data(), m);
        buf[m] = 0;
    }
    return GENPF_OK;
}

int32_t genpf_device_count(int32_t *count) {
    if (!count) return fail(GENPF_ERR_INVALID_ARG, "count is NULL");
    int c = 0;
    GENPF_CUDA_TRY(cudaGetDeviceCount(&c));
    *count = c;
    return GENPF_OK;
}
int32_t genpf_set_device(int32_t device) {
    GENPF_CUDA_TRY(cudaSetDevice(device));
    return GENPF_OK;
}
int32_t genpf_synchronize(void) {
    GENPF_CUDA_TRY(cudaDeviceSynchronize());
    return GENPF_OK;
}

int32_t genpf_logsumexp(const double *lw, int64_t n, uint32_t flags, double *out) {
    if (!lw || !out || n <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_logsumexp: bad arguments (n must be > 0)");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const double *d_lw;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n, flags & GENPF_DEVICE_PTRS, &d_lw));
    Stats st;
    GENPF_TRY(reduce_to_host(ws, d_lw, n, &st));
    // Gen.logsumexp: NaN input propagates through maximum()
    *out = (st.invalid_kind == GENPF_INV_NAN_INPUT || st.invalid_kind == GENPF_INV_NAN_TOTAL) ? NAN : st.lse;
    return GENPF_OK;
}

int32_t genpf_ess(const double *lw, int64_t n, uint32_t flags, double *out) {
    if (!lw || !out || n <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_ess: bad arguments (n must be > 0)");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const double *d_lw;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n, flags & GENPF_DEVICE_PTRS, &d_lw));
    Stats st;
    GENPF_TRY(reduce_to_host(ws, d_lw, n, &st));
    *out = st.invalid_kind == GENPF_VALID ? st.ess : NAN;
    return GENPF_OK;
}

int32_t genpf_normalize(const double *lw, int64_t n, uint32_t flags, double *log_norm, double *norm, double *lse,
                        double *ess, int32_t *invalid_kind) {
    if (!lw || n <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_normalize: bad arguments (n must be > 0)");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const double *d_lw;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n, dp, &d_lw));
    Stats st;
    GENPF_TRY(reduce_to_host(ws, d_lw, n, &st));
    if (lse) *lse = st.lse;
    if (ess) *ess = st.ess;
    if (invalid_kind) *invalid_kind = st.invalid_kind;
    if (log_norm || norm) {
        double *d_ln, *d_nm;
        GENPF_TRY(stage_out(ws.aux1, log_norm, n, dp, &d_ln));
        GENPF_TRY(stage_out(ws.aux2, norm, n, dp, &d_nm));
        GENPF_LAUNCH(k_normalize_out, grid_1d(n), 256, ws.stream, d_lw, n, ws.sc.st(0, 1), d_ln, d_nm);
        GENPF_TRY(copy_out(ws, d_ln, log_norm, n, dp));
        GENPF_TRY(copy_out(ws, d_nm, norm, n, dp));
        GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    }
    return GENPF_OK;
}

// shared body of genpf_resample / genpf_resample_segmented (one segment)
static int32_t resample_one(HostWs &ws, int32_t method, const double *d_lw, const double *d_lp, int64_t n_in,
                            int64_t n_out, const double *d_u, uint64_t seed, int64_t slot_offset, uint32_t flags,
                            int64_t *d_parents, double *d_lw_out, double *lml_increment, int32_t *invalid_kind) {
    cudaStream_t s = ws.stream;
    Scratch &sc = ws.sc;
    const bool substate = flags & GENPF_SUBSTATE;
    GENPF_TRY(sc.ensure(n_in > n_out ? n_in : n_out, 1));
    LwSrc lw_src{d_lw, 1.0};
    LwSrc sel = d_lp ? LwSrc{d_lp, 1.0} : lw_src;
    Stats *st_lw = sc.st(0, 1), *st_sel = d_lp ? sc.st(1, 1) : st_lw, *st_d = sc.st(2, 1);
    GENPF_TRY(launch_reduce(s, lw_src, n_in, 1, sc.partials(0)));
    GENPF_TRY(launch_finalize(s, sc, sc.partials(0), n_in, 1, st_lw, d_lp ? nullptr : sc.tile_off.as<double>(), -1.0, nullptr));
    if (d_lp) {
        GENPF_TRY(launch_reduce(s, sel, n_in, 1, sc.partials(1)));
        GENPF_TRY(launch_finalize(s, sc, sc.partials(1), n_in, 1, st_sel, sc.tile_off.as<double>(), -1.0, nullptr));
    }
    GENPF_CUDA_TRY(cudaMemcpyAsync(ws.h_stats, sc.stats.p, sizeof(Stats) * 2, cudaMemcpyDeviceToHost, s));
    GENPF_CUDA_TRY(cudaStreamSynchronize(s));
    const Stats h_lw = ws.h_stats[0], h_sel = d_lp ? ws.h_stats[1] : ws.h_stats[0];
    const int kind = h_sel.invalid_kind;
    if (invalid_kind) *invalid_kind = kind;
    // check == true && invalid && error("Invalid weights.")  -- before any mutation (resample.jl:55,92,151)
    if ((flags & GENPF_CHECK) && kind != GENPF_VALID) return fail(GENPF_ERR_INVALID_WEIGHTS, "Invalid weights.");
    // update_lml_est!: uses log_weights (not priorities) and the pre-resize n (resample.jl:178-182)
    if (lml_increment) *lml_increment = substate ? 0.0 : h_lw.lse - log((double)n_in);
    if (kind == GENPF_INV_NAN_INPUT || kind == GENPF_INV_NAN_TOTAL) return GENPF_OK;  // reference crashes here

    UniSrc uni{d_u, seed, make_stream(0, 0), slot_offset};
    const int64_t base = (flags & GENPF_INDEX_BASE1) ? 1 : 0;
    // update_weights! (resample.jl:190-218, resize.jl:424-438): without priorities the constant weight is written
    // by the kernels that write the ancestors
    GENPF_TRY(select_ancestors<long long>(s, sc, method, sel, n_in, n_out, 1, st_sel, uni, flags,
                                          reinterpret_cast<long long *>(d_parents), base, 0, nullptr,
                                          LwFill{d_lp ? nullptr : d_lw_out, st_lw, substate ? 1 : 0}));
    if (d_lp) {
        GENPF_LAUNCH((k_prio_ratio<long long>), dim3(grid_1d(n_out), 1), 256, s, d_lw, sel,
                     reinterpret_cast<const long long *>(d_parents), base, n_in, n_out, d_lw_out);
        LwSrc dsrc{d_lw_out, 1.0};
        GENPF_TRY(launch_reduce(s, dsrc, n_out, 1, sc.partials(2)));
        GENPF_TRY(launch_finalize(s, sc, sc.partials(2), n_out, 1, st_d, nullptr, -1.0, nullptr));
        GENPF_LAUNCH(k_prio_shift, dim3(grid_1d(n_out), 1), 256, s, d_lw_out, n_out, st_d, st_lw, substate ? 1 : 0);
    }
    return GENPF_OK;
}

int32_t genpf_resample(int32_t method, const double *lw, const double *log_prio, int64_t n_in, int64_t n_out,
                       const double *uniforms, uint64_t seed, uint32_t flags, int64_t *parents_out, double *lw_out,
                       double *lml_increment, int32_t *invalid_kind) {
    if (!lw || !parents_out || !lw_out) return fail(GENPF_ERR_INVALID_ARG, "genpf_resample: NULL array argument");
    if (n_in <= 0 || n_out <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_resample: empty particle set");
    if (method != GENPF_MULTINOMIAL && method != GENPF_RESIDUAL && method != GENPF_STRATIFIED)
        return fail(GENPF_ERR_UNKNOWN_METHOD, "Resampling method not recognized.");
    if (method == GENPF_STRATIFIED && n_out != n_in)
        return fail(GENPF_ERR_INVALID_ARG, "stratified resampling cannot resize (resize.jl:16-27)");
    if ((flags & GENPF_SUBSTATE) && n_out != n_in)
        return fail(GENPF_ERR_INVALID_ARG, "a sub-state cannot be resized");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const double *d_lw, *d_lp, *d_u;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n_in, dp, &d_lw));
    GENPF_TRY(stage_in(ws, ws.lp, log_prio, n_in, dp, &d_lp));
    GENPF_TRY(stage_in(ws, ws.u, uniforms, n_out, dp, &d_u));
    int64_t *d_par;
    double *d_out;
    GENPF_TRY(stage_out(ws.parents, parents_out, n_out, dp, &d_par));
    GENPF_TRY(stage_out(ws.lw_out, lw_out, n_out, dp, &d_out));
    int32_t kind = 0;
    GENPF_TRY(resample_one(ws, method, d_lw, d_lp, n_in, n_out, d_u, seed, 0, flags, d_par, d_out, lml_increment, &kind));
    if (invalid_kind) *invalid_kind = kind;
    if (kind != GENPF_INV_NAN_INPUT && kind != GENPF_INV_NAN_TOTAL) {
        GENPF_TRY(copy_out(ws, d_par, parents_out, n_out, dp));
        GENPF_TRY(copy_out(ws, d_out, lw_out, n_out, dp));
    }
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    return GENPF_OK;
}

int32_t genpf_resample_segmented(int32_t method, const double *lw, const double *log_prio, int64_t n,
                                 const int64_t *seg_offsets, int64_t n_seg, const double *uniforms, uint64_t seed,
                                 uint32_t flags, int64_t *parents_out, double *lw_out, int32_t *invalid_kinds) {
    if (!lw || !parents_out || !lw_out || !seg_offsets || n <= 0 || n_seg <= 0)
        return fail(GENPF_ERR_INVALID_ARG, "genpf_resample_segmented: bad arguments");
    if (method != GENPF_MULTINOMIAL && method != GENPF_RESIDUAL && method != GENPF_STRATIFIED)
        return fail(GENPF_ERR_UNKNOWN_METHOD, "Resampling method not recognized.");
    for (int64_t sgi = 0; sgi < n_seg; ++sgi)
        if (seg_offsets[sgi] < 0 || seg_offsets[sgi + 1] <= seg_offsets[sgi] || seg_offsets[sgi + 1] > n)
            return fail(GENPF_ERR_INVALID_ARG, "genpf_resample_segmented: segments must be non-empty, ordered, in range");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const double *d_lw, *d_lp, *d_u;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n, dp, &d_lw));
    GENPF_TRY(stage_in(ws, ws.lp, log_prio, n, dp, &d_lp));
    GENPF_TRY(stage_in(ws, ws.u, uniforms, n, dp, &d_u));
    int64_t *d_par;
    double *d_out;
    GENPF_TRY(stage_out(ws.parents, parents_out, n, dp, &d_par));
    GENPF_TRY(stage_out(ws.lw_out, lw_out, n, dp, &d_out));
    if (!dp) {  // positions outside every segment keep their input weight / identity parent
        GENPF_CUDA_TRY(cudaMemcpyAsync(d_out, d_lw, (size_t)n * 8, cudaMemcpyDeviceToDevice, ws.stream));
        GENPF_CUDA_TRY(cudaMemsetAsync(d_par, 0, (size_t)n * 8, ws.stream));
    }
    const uint32_t f2 = flags | GENPF_SUBSTATE;
    for (int64_t sgi = 0; sgi < n_seg; ++sgi) {
        const int64_t a = seg_offsets[sgi], len = seg_offsets[sgi + 1] - a;
        int32_t kind = 0;
        GENPF_TRY(resample_one(ws, method, d_lw + a, d_lp ? d_lp + a : nullptr, len, len, d_u ? d_u + a : nullptr, seed,
                               a, f2, d_par + a, d_out + a, nullptr, &kind));
        if (invalid_kinds) invalid_kinds[sgi] = kind;
    }
    GENPF_TRY(copy_out(ws, d_par, parents_out, n, dp));
    GENPF_TRY(copy_out(ws, d_out, lw_out, n, dp));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    return GENPF_OK;
}

int32_t genpf_weighted_mean_var(const double *lw, const double *x, int64_t n, uint32_t flags, double *mean,
                                double *var) {
    if (!lw || !x || n <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_weighted_mean_var: bad arguments");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const double *d_lw, *d_x;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n, dp, &d_lw));
    GENPF_TRY(stage_in(ws, ws.x, x, n, dp, &d_x));
    GENPF_TRY(ws.sc.ensure(n, 1));
    LwSrc src{d_lw, 1.0};
    GENPF_TRY(launch_reduce(ws.stream, src, n, 1, ws.sc.partials(0)));
    GENPF_TRY(launch_finalize(ws.stream, ws.sc, ws.sc.partials(0), n, 1, ws.sc.st(0, 1), nullptr, -1.0, nullptr));
    XSrc xs{d_x, nullptr};
    GENPF_TRY(launch_mean_var(ws.stream, ws.sc, d_lw, xs, n, 1, ws.sc.st(0, 1)));
    GENPF_CUDA_TRY(cudaMemcpyAsync(ws.h_scalars, ws.sc.moment_out.p, 16, cudaMemcpyDeviceToHost, ws.stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    if (mean) *mean = ws.h_scalars[0];
    if (var) *var = ws.h_scalars[1];
    return GENPF_OK;
}

int32_t genpf_replicate_host(const double *lw, int64_t n, int64_t k, int32_t layout, uint32_t flags,
                             int64_t *parents_out, double *lw_out) {
    if (!lw || !parents_out || !lw_out || n <= 0 || k <= 0)
        return fail(GENPF_ERR_INVALID_ARG, "genpf_replicate_host: bad arguments");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const double *d_lw;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n, dp, &d_lw));
    int64_t *d_par;
    double *d_out;
    GENPF_TRY(stage_out(ws.parents, parents_out, n * k, dp, &d_par));
    GENPF_TRY(stage_out(ws.lw_out, lw_out, n * k, dp, &d_out));
    GENPF_LAUNCH((k_replicate<long long>), dim3(grid_1d(n * k), 1), 256, ws.stream, d_lw, n, k,
                 layout == GENPF_LAYOUT_INTERLEAVED ? 1 : 0, reinterpret_cast<long long *>(d_par),
                 (int64_t)((flags & GENPF_INDEX_BASE1) ? 1 : 0), d_out);
    GENPF_TRY(copy_out(ws, d_par, parents_out, n * k, dp));
    GENPF_TRY(copy_out(ws, d_out, lw_out, n * k, dp));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    return GENPF_OK;
}

int32_t genpf_dereplicate_host(const double *lw, int64_t n, int64_t k, int32_t layout, int32_t method,
                               const double *uniforms, uint64_t seed, uint32_t flags, int64_t *parents_out,
                               double *lw_out) {
    if (!lw || !parents_out || !lw_out || n <= 0 || k <= 0)
        return fail(GENPF_ERR_INVALID_ARG, "genpf_dereplicate_host: bad arguments");
    if (n % k != 0) return fail(GENPF_ERR_INVALID_ARG, "n must be a multiple of n_replicates (resize.jl:270)");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const int64_t n_new = n / k;
    const double *d_lw, *d_u;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n, dp, &d_lw));
    GENPF_TRY(stage_in(ws, ws.u, uniforms, n_new, dp, &d_u));
    int64_t *d_par;
    double *d_out;
    GENPF_TRY(stage_out(ws.parents, parents_out, n_new, dp, &d_par));
    GENPF_TRY(stage_out(ws.lw_out, lw_out, n_new, dp, &d_out));
    UniSrc uni{d_u, seed, make_stream(kPurposeDerep, 0), 0};
    GENPF_LAUNCH((k_dereplicate<long long>), dim3(grid_1d(n_new), 1), 256, ws.stream, d_lw, n, k,
                 layout == GENPF_LAYOUT_INTERLEAVED ? 1 : 0, method == GENPF_SAMPLE ? 1 : 0, uni,
                 reinterpret_cast<long long *>(d_par), (int64_t)((flags & GENPF_INDEX_BASE1) ? 1 : 0), d_out);
    GENPF_TRY(copy_out(ws, d_par, parents_out, n_new, dp));
    GENPF_TRY(copy_out(ws, d_out, lw_out, n_new, dp));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    return GENPF_OK;
}

int32_t genpf_proportionmap_host(const double *lw, const int64_t *keys, int64_t n, uint32_t flags,
                                 int64_t *first_index_out, double *prop_out, int64_t *n_unique) {
    if (!lw || !keys || !first_index_out || !prop_out || n <= 0)
        return fail(GENPF_ERR_INVALID_ARG, "genpf_proportionmap_host: bad arguments");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const double *d_lw;
    const int64_t *d_keys;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n, dp, &d_lw));
    GENPF_TRY(stage_in(ws, ws.keys, keys, n, dp, &d_keys));
    int64_t *d_first;
    double *d_prop;
    GENPF_TRY(stage_out(ws.parents, first_index_out, n, dp, &d_first));
    GENPF_TRY(stage_out(ws.lw_out, prop_out, n, dp, &d_prop));
    GENPF_TRY(ws.sc.ensure(n, 1));
    LwSrc src{d_lw, 1.0};
    GENPF_TRY(launch_reduce(ws.stream, src, n, 1, ws.sc.partials(0)));
    GENPF_TRY(launch_finalize(ws.stream, ws.sc, ws.sc.partials(0), n, 1, ws.sc.st(0, 1), nullptr, -1.0, nullptr));
    long long *n_dev = nullptr;
    GENPF_TRY(launch_coalesce<long long>(ws.stream, ws.cb, d_lw, d_keys, n, reinterpret_cast<long long *>(d_first),
                                         (int64_t)((flags & GENPF_INDEX_BASE1) ? 1 : 0), d_prop, &n_dev,
                                         (const Stats *)ws.sc.st(0, 1)));
    GENPF_CUDA_TRY(cudaMemcpyAsync(ws.h_scalars, n_dev, 8, cudaMemcpyDeviceToHost, ws.stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    const int64_t g = (int64_t) * reinterpret_cast<long long *>(ws.h_scalars);
    if (n_unique) *n_unique = g;
    GENPF_TRY(copy_out(ws, d_first, first_index_out, g, dp));
    GENPF_TRY(copy_out(ws, d_prop, prop_out, g, dp));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    return GENPF_OK;
}

int32_t genpf_optimal_resize(const double *lw, int64_t n_in, int64_t n_out, const double *uniform, uint64_t seed,
                             uint32_t flags, int64_t *parents_out, double *lw_out, int64_t *n_keep,
                             double *inv_w_threshold, int32_t *invalid_kinds) {
    if (!lw || !parents_out || !lw_out) return fail(GENPF_ERR_INVALID_ARG, "genpf_optimal_resize: NULL array argument");
    if (n_in <= 0 || n_out <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_optimal_resize: empty particle set");
    if (n_out > n_in) return fail(GENPF_ERR_ASSERT, "optimal resize cannot grow the filter (@assert n_particles <= n_old, resize.jl:183)");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const double *d_lw;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n_in, dp, &d_lw));
    int64_t *d_par;
    double *d_out;
    GENPF_TRY(stage_out(ws.parents, parents_out, n_out, dp, &d_par));
    GENPF_TRY(stage_out(ws.lw_out, lw_out, n_out, dp, &d_out));
    OptResult res;
    UniSrc uni{nullptr, seed, make_stream(0, 0), 0};
    const int32_t st = optimal_resize_core<long long>(ws.stream, ws.sc, ws.ob, ws.h_ctrl, ws.h_stats, d_lw, n_in, n_out,
                                                      uniform, uni, flags, reinterpret_cast<long long *>(d_par),
                                                      (int64_t)((flags & GENPF_INDEX_BASE1) ? 1 : 0), d_out, &res);
    if (n_keep) *n_keep = res.n_keep;
    if (inv_w_threshold) *inv_w_threshold = res.inv_w;
    if (invalid_kinds) {
        invalid_kinds[0] = res.kind;
        invalid_kinds[1] = res.kind_strat;
    }
    if (st != GENPF_OK) return st;
    GENPF_TRY(copy_out(ws, d_par, parents_out, n_out, dp));
    GENPF_TRY(copy_out(ws, d_out, lw_out, n_out, dp));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    return GENPF_OK;
}

int32_t genpf_uniforms(uint64_t seed, uint64_t stream, int64_t n, uint32_t flags, double *out) {
    if (!out || n <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_uniforms: bad arguments");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    double *d_out;
    GENPF_TRY(stage_out(ws.aux1, out, n, dp, &d_out));
    UniSrc uni{nullptr, seed, stream, 0};
    GENPF_LAUNCH(k_uniforms, grid_1d(n), 256, ws.stream, uni, n, d_out, (flags & GENPF_UNIFORMS_STRATA) ? 1 : 0);
    GENPF_TRY(copy_out(ws, d_out, out, n, dp));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    return GENPF_OK;
}

int32_t genpf_debug_sortperm(const double *keys, int64_t n, uint32_t flags, int64_t *order_out) {
    if (!keys || !order_out || n <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_debug_sortperm: bad arguments");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const double *d_keys;
    GENPF_TRY(stage_in(ws, ws.lw, keys, n, dp, &d_keys));
    int64_t *d_out;
    GENPF_TRY(stage_out(ws.parents, order_out, n, dp, &d_out));
    GENPF_TRY(ws.sc.order.ensure((size_t)n * 4));
    GENPF_TRY(sort_desc_stable(d_keys, n, nullptr, ws.sc.order.as<int32_t>(), ws.sc.sort_tmp, ws.stream));
    GENPF_LAUNCH((k_convert_idx<int32_t, long long>), grid_1d(n), 256, ws.stream, (const int32_t *)ws.sc.order.as<int32_t>(),
                 reinterpret_cast<long long *>(d_out), n, (int64_t)((flags & GENPF_INDEX_BASE1) ? 1 : 0));
    GENPF_TRY(copy_out(ws, d_out, order_out, n, dp));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    return GENPF_OK;
}

int32_t genpf_debug_cumweights(const double *lw, int64_t n, uint32_t flags, double *W_out) {
    if (!lw || !W_out || n <= 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_debug_cumweights: bad arguments");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const double *d_lw;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n, dp, &d_lw));
    double *d_W;
    GENPF_TRY(stage_out(ws.aux1, W_out, n, dp, &d_W));
    GENPF_TRY(ws.sc.ensure(n, 1));
    LwSrc src{d_lw, 1.0};
    const int64_t tpf = ceil_div(n, kTile);
    GENPF_TRY(launch_reduce(ws.stream, src, n, 1, ws.sc.partials(0)));
    GENPF_TRY(launch_finalize(ws.stream, ws.sc, ws.sc.partials(0), n, 1, ws.sc.st(0, 1), ws.sc.tile_off.as<double>(), -1.0, nullptr));
    UniSrc uni{nullptr, 0, 0, 0};
    StratArgs none = make_strat(uni, n);
    GENPF_LAUNCH((k_scan<int32_t>), (unsigned)tpf, kScanThreads, ws.stream, src, n, tpf, ws.sc.st(0, 1),
                 ws.sc.tile_off.as<double>(), WTables{d_W}, (int32_t *)nullptr, (int32_t *)nullptr, none, 0,
                 (const double *)nullptr, (int64_t)0, ws.sc.chunk_info_ptr(n), Scratch::kChunkTiles);
    GENPF_TRY(copy_out(ws, d_W, W_out, n, dp));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    return GENPF_OK;
}

int32_t genpf_coalesce_host(const double *lw, const int64_t *keys, int64_t n, uint32_t flags, int64_t *parents_out,
                            double *lw_out, int64_t *n_new) {
    if (!lw || !keys || !parents_out || !lw_out || !n_new || n <= 0)
        return fail(GENPF_ERR_INVALID_ARG, "genpf_coalesce_host: bad arguments");
    if (n >= 0x7FFFFFF0ll) return fail(GENPF_ERR_UNSUPPORTED, "coalesce: n must be < 2^31");
    HostWs &ws = g_ws;
    GENPF_TRY(ws.init());
    const bool dp = flags & GENPF_DEVICE_PTRS;
    const double *d_lw;
    const int64_t *d_keys;
    GENPF_TRY(stage_in(ws, ws.lw, lw, n, dp, &d_lw));
    GENPF_TRY(stage_in(ws, ws.keys, keys, n, dp, &d_keys));
    int64_t *d_par;
    double *d_out;
    GENPF_TRY(stage_out(ws.parents, parents_out, n, dp, &d_par));
    GENPF_TRY(stage_out(ws.lw_out, lw_out, n, dp, &d_out));
    long long *n_new_dev = nullptr;
    GENPF_TRY(launch_coalesce<long long>(ws.stream, ws.cb, d_lw, d_keys, n, reinterpret_cast<long long *>(d_par),
                                         (int64_t)((flags & GENPF_INDEX_BASE1) ? 1 : 0), d_out, &n_new_dev));
    GENPF_CUDA_TRY(cudaMemcpyAsync(ws.h_scalars, n_new_dev, 8, cudaMemcpyDeviceToHost, ws.stream));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    const int64_t nn = (int64_t) * reinterpret_cast<long long *>(ws.h_scalars);
    *n_new = nn;
    GENPF_TRY(copy_out(ws, d_par, parents_out, nn, dp));
    GENPF_TRY(copy_out(ws, d_out, lw_out, nn, dp));
    GENPF_CUDA_TRY(cudaStreamSynchronize(ws.stream));
    return GENPF_OK;
}

}  // extern "C"
