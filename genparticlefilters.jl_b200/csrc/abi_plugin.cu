// abi_plugin.cu -- user-registrable device plugins (BASELINE north star: "models registered as device plugins (a C-ABI
// transition / observation log-density / MH-proposal kernel over struct-of-arrays particle state)"; SURVEY 8b
// `genpf_model_load_cubin`, 8f rank 4).
//
// A plugin is CUDA C++ source defining ONE struct with the interface of models.cuh (NF/NB/NP/NAUX, initial,
// transition, obs_logpdf, and optionally constrain / propose / proposal_logpdf / transition_logpdf / translate).
// genpf_model_compile hands the source to NVRTC together with THIS library's own kernel headers (embedded at build
// time, plugin_sources.inc), so the model's functions are inlined into the same k_propagate / k_mh / k_step_fused
// templates the built-in models use -- same arithmetic, same tile partition, same Philox streams: a model
// registered from source is bit-identical to the same model compiled into the library.  The result is an sm_100a
// cubin (no driver JIT), loaded with cudaLibraryLoadData when the first filter of that model is created; kernels
// are launched through cudaLaunchKernel with the exact parameter lists of the templates.
// genpf_model_export / genpf_model_load_image move a compiled plugin as one self-describing image (names + cubin)
// so a deployment needs NVRTC only once.
// libnvrtc is opened with dlopen at the first compile: the library itself keeps depending on the CUDA runtime only.
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "filter_state.hpp"
#include "plugin.hpp"

namespace genpf {

// ---- the library's own headers, as NVRTC include files
struct EmbeddedHeader {
    const char *name;
    const char *text;
};
#include "plugin_sources.inc"  // static const EmbeddedHeader kPluginHeaders[]; static const int kNumPluginHeaders

// ---- NVRTC through dlopen
typedef void *nvrtcProgram;
struct Nvrtc {
    void *h = nullptr;
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    int (*DestroyProgram)(nvrtcProgram *);
    int (*CompileProgram)(nvrtcProgram, int, const char *const *);
    int (*GetProgramLogSize)(nvrtcProgram, size_t *);
    int (*GetProgramLog)(nvrtcProgram, char *);
    int (*GetCUBINSize)(nvrtcProgram, size_t *);
    int (*GetCUBIN)(nvrtcProgram, char *);
    int (*AddNameExpression)(nvrtcProgram, const char *);
    int (*GetLoweredName)(nvrtcProgram, const char *, const char **);
    const char *(*GetErrorString)(int);
};
static Nvrtc g_nvrtc;
static std::mutex g_plugin_mu;

static int32_t load_nvrtc() {
    if (g_nvrtc.h) return GENPF_OK;
    const char *cands[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    void *h = nullptr;
    for (const char *c : cands)
        if ((h = dlopen(c, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) return fail(GENPF_ERR_UNSUPPORTED, "genpf_model_compile: libnvrtc.so.12 not found (needed to compile plugin source)");
#define GENPF_NVRTC_SYM(field, sym)                                                     \
    *(void **)(&g_nvrtc.field) = dlsym(h, sym);                                         \
    if (!g_nvrtc.field) return fail(GENPF_ERR_UNSUPPORTED, std::string("libnvrtc lacks ") + sym)
    GENPF_NVRTC_SYM(CreateProgram, "nvrtcCreateProgram");
    GENPF_NVRTC_SYM(DestroyProgram, "nvrtcDestroyProgram");
    GENPF_NVRTC_SYM(CompileProgram, "nvrtcCompileProgram");
    GENPF_NVRTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    GENPF_NVRTC_SYM(GetProgramLog, "nvrtcGetProgramLog");
    GENPF_NVRTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    GENPF_NVRTC_SYM(GetCUBIN, "nvrtcGetCUBIN");
    GENPF_NVRTC_SYM(AddNameExpression, "nvrtcAddNameExpression");
    GENPF_NVRTC_SYM(GetLoweredName, "nvrtcGetLoweredName");
    GENPF_NVRTC_SYM(GetErrorString, "nvrtcGetErrorString");
#undef GENPF_NVRTC_SYM
    g_nvrtc.h = h;
    return GENPF_OK;
}

// ---- registry
static std::vector<PluginModel *> g_plugins;

const ModelInfo *model_info(int32_t id) {
    if (id >= 0 && id < kNumModels) return &kModels[id];
    std::lock_guard<std::mutex> lk(g_plugin_mu);
    const int k = id - kPluginIdBase;
    if (k >= 0 && k < (int)g_plugins.size()) return &g_plugins[(size_t)k]->info;
    return nullptr;
}
PluginModel *plugin_model(int32_t id) {
    std::lock_guard<std::mutex> lk(g_plugin_mu);
    const int k = id - kPluginIdBase;
    return (k >= 0 && k < (int)g_plugins.size()) ? g_plugins[(size_t)k] : nullptr;
}
int32_t find_model(const char *name) {
    for (int i = 0; i < kNumModels; ++i)
        if (strcmp(name, kModels[i].name) == 0) return i;
    std::lock_guard<std::mutex> lk(g_plugin_mu);
    for (size_t k = 0; k < g_plugins.size(); ++k)
        if (g_plugins[k]->name == name) return kPluginIdBase + (int)k;
    return -1;
}

static const char *kNoiseNames[3] = {"genpf::NoiseLean", "genpf::NoisePhilox53", "genpf::NoiseCols"};
static const char *kFusedMh[3] = {"1", "0", "-1"};

// the name expression of kernel slot `k` for struct `S`
static std::string kernel_expr(int k, const std::string &S) {
    if (k < kPlugMh) {
        const int nz = (k - kPlugProp) / 2, init = (k - kPlugProp) % 2;
        return "genpf::k_propagate<" + S + ", " + kNoiseNames[nz] + ", " + (init ? "true" : "false") + ">";
    }
    if (k < kPlugFused) {
        const int nz = (k - kPlugMh) / 2, rw = (k - kPlugMh) % 2;
        return "genpf::k_mh<" + S + ", " + kNoiseNames[nz] + ", " + (rw ? "true" : "false") + ">";
    }
    if (k >= kPlugIntro) return "genpf::k_introduce<" + S + ", " + kNoiseNames[k - kPlugIntro] + ">";
    const int nz = (k - kPlugFused) / 3, mh = (k - kPlugFused) % 3;
    return "genpf::k_step_fused<" + S + ", " + kNoiseNames[nz] + ", int, " + kFusedMh[mh] + ">";
}

// "…plugin_dimsILi1ELi0ELi5ELi0ELi1ELi0EEvv" -> six integers
static bool parse_dims(const char *lowered, int out[6]) {
    const char *p = strstr(lowered, "plugin_dimsI");
    if (!p) return false;
    p += strlen("plugin_dimsI");
    for (int i = 0; i < 6; ++i) {
        if (p[0] != 'L' || (p[1] != 'i' && p[1] != 'b')) return false;
        p += 2;
        bool neg = false;
        if (*p == 'n') { neg = true; ++p; }
        int v = 0;
        while (*p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
        if (*p != 'E') return false;
        ++p;
        out[i] = neg ? -v : v;
    }
    return true;
}

static int32_t register_plugin(PluginModel *pm, int32_t *model_id) {
    std::lock_guard<std::mutex> lk(g_plugin_mu);
    for (size_t k = 0; k < g_plugins.size(); ++k)
        if (g_plugins[k]->name == pm->name) {  // re-registration replaces the model of that name (new id)
            g_plugins[k]->name += "#superseded";
        }
    g_plugins.push_back(pm);
    *model_id = kPluginIdBase + (int)g_plugins.size() - 1;
    return GENPF_OK;
}

// make sure the cubin is loaded on the current device and kernel slot k is resolved
int32_t plugin_kernel(PluginModel *pm, int k, const void **fn) {
    std::lock_guard<std::mutex> lk(pm->mu);
    if (!pm->lib) {
        cudaLibrary_t lib = nullptr;
        cudaError_t e = cudaLibraryLoadData(&lib, pm->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(GENPF_ERR_CUDA, std::string("cudaLibraryLoadData failed for plugin '") + pm->name + "': " + cudaGetErrorString(e));
        }
        pm->lib = lib;
    }
    if (!pm->kern[k]) {
        if (pm->lowered[k].empty()) return fail(GENPF_ERR_UNSUPPORTED, "plugin image lacks a kernel this call needs");
        cudaKernel_t kk = nullptr;
        cudaError_t e = cudaLibraryGetKernel(&kk, (cudaLibrary_t)pm->lib, pm->lowered[k].c_str());
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(GENPF_ERR_CUDA, std::string("cudaLibraryGetKernel(") + pm->lowered[k] + "): " + cudaGetErrorString(e));
        }
        pm->kern[k] = (const void *)kk;
    }
    *fn = pm->kern[k];
    return GENPF_OK;
}

}  // namespace genpf

using namespace genpf;

extern "C" {

int32_t genpf_model_compile(const char *name, const char *source, const char *struct_name, const char *options,
                            int32_t *model_id) {
    if (!name || !source || !struct_name || !model_id) return fail(GENPF_ERR_INVALID_ARG, "genpf_model_compile: NULL argument");
    for (int i = 0; i < kNumModels; ++i)
        if (strcmp(name, kModels[i].name) == 0) return fail(GENPF_ERR_INVALID_ARG, "genpf_model_compile: name of a built-in model");
    GENPF_TRY(load_nvrtc());
    const std::string S(struct_name);
    // translation unit = the user's source + the dimension probe
    std::string tu = "#include \"genpf_plugin.h\"\n#line 1 \"";
    tu += name;
    tu += ".cu\"\n";
    tu += source;
    tu += "\nnamespace genpf { template <int NF, int NB, int NP, int NAUX, int PROP, int TRANS> __global__ void plugin_dims() {} }\n";
    std::vector<const char *> hdr_names, hdr_texts;
    for (int i = 0; i < kNumPluginHeaders; ++i) {
        hdr_names.push_back(kPluginHeaders[i].name);
        hdr_texts.push_back(kPluginHeaders[i].text);
    }
    static const char kPluginH[] = "#pragma once\n#ifndef GENPF_PLUGIN_BUILD\n#define GENPF_PLUGIN_BUILD 1\n#endif\n#include \"fused.cuh\"\n";
    hdr_names.push_back("genpf_plugin.h");
    hdr_texts.push_back(kPluginH);
    nvrtcProgram prog = nullptr;
    int r = g_nvrtc.CreateProgram(&prog, tu.c_str(), (std::string(name) + ".cu").c_str(), (int)hdr_names.size(), hdr_texts.data(),
                                  hdr_names.data());
    if (r != 0) return fail(GENPF_ERR_CUDA, std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r));
    const std::string dims_expr = "genpf::plugin_dims<" + S + "::NF, " + S + "::NB, " + S + "::NP, " + S + "::NAUX, genpf::has_proposal<" +
                                  S + ">::value, genpf::has_translate<" + S + ">::value>";
    std::vector<std::string> exprs;
    for (int k = 0; k < kPlugKernels; ++k) exprs.push_back(kernel_expr(k, S));
    g_nvrtc.AddNameExpression(prog, dims_expr.c_str());
    for (auto &e : exprs) g_nvrtc.AddNameExpression(prog, e.c_str());
    std::vector<std::string> opt_store = {"--gpu-architecture=sm_100a", "--std=c++17", "-DGENPF_PLUGIN_BUILD=1",
                                          "-DGENPF_STATE_THREADS=" + std::to_string(kStateThreads)};
    if (options && *options) {  // extra options, space separated
        std::string o(options);
        size_t a = 0;
        while (a < o.size()) {
            size_t b = o.find(' ', a);
            if (b == std::string::npos) b = o.size();
            if (b > a) opt_store.push_back(o.substr(a, b - a));
            a = b + 1;
        }
    }
    std::vector<const char *> opts;
    for (auto &o : opt_store) opts.push_back(o.c_str());
    r = g_nvrtc.CompileProgram(prog, (int)opts.size(), opts.data());
    if (r != 0) {
        size_t n = 0;
        g_nvrtc.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) g_nvrtc.GetProgramLog(prog, &log[0]);
        g_nvrtc.DestroyProgram(&prog);
        if (log.size() > 6000) log.resize(6000);
        return fail(GENPF_ERR_INVALID_ARG, std::string("plugin '") + name + "' does not compile (" + g_nvrtc.GetErrorString(r) + "):\n" + log);
    }
    PluginModel *pm = new PluginModel();
    pm->name = name;
    const char *low = nullptr;
    int dims[6] = {0, 0, 0, 0, 0, 0};
    if (g_nvrtc.GetLoweredName(prog, dims_expr.c_str(), &low) != 0 || !low || !parse_dims(low, dims)) {
        g_nvrtc.DestroyProgram(&prog);
        delete pm;
        return fail(GENPF_ERR_INVALID_ARG, "plugin: cannot read NF/NB/NP/NAUX of the model struct");
    }
    if (dims[0] < 0 || dims[0] > kMaxF || dims[1] < 0 || dims[1] > kMaxB || dims[2] < 0 || dims[2] > kMaxParams || dims[3] < 0 ||
        dims[3] > kMaxAux || dims[0] + dims[1] < 1) {
        g_nvrtc.DestroyProgram(&prog);
        delete pm;
        return fail(GENPF_ERR_INVALID_ARG, "plugin: need 1 <= NF + NB, NF <= 2, NB <= 2, NP <= 8, NAUX <= 4");
    }
    for (int k = 0; k < kPlugKernels; ++k) {
        low = nullptr;
        if (g_nvrtc.GetLoweredName(prog, exprs[(size_t)k].c_str(), &low) == 0 && low) pm->lowered[k] = low;
    }
    size_t sz = 0;
    g_nvrtc.GetCUBINSize(prog, &sz);
    pm->cubin.resize(sz);
    if (sz) g_nvrtc.GetCUBIN(prog, pm->cubin.data());
    g_nvrtc.DestroyProgram(&prog);
    if (!sz) {
        delete pm;
        return fail(GENPF_ERR_CUDA, "plugin: NVRTC produced no cubin");
    }
    pm->name_store = pm->name;
    pm->info = ModelInfo{pm->name_store.c_str(), dims[0], dims[1], dims[2], dims[3], (dims[4] ? 1 : 0) | (dims[5] ? 2 : 0)};
    return register_plugin(pm, model_id);
}

// image = "GENPFPLG" | u32 version | i32 dims[6] | u32 name_len | name | per kernel: u32 len | lowered name | u64 cubin size | cubin
int32_t genpf_model_export(int32_t model_id, void *buf, int64_t cap, int64_t *size) {
    PluginModel *pm = plugin_model(model_id);
    if (!pm || !size) return fail(GENPF_ERR_INVALID_ARG, "genpf_model_export: not a plugin model");
    std::string img = "GENPFPLG";
    auto put32 = [&](uint32_t v) { img.append(reinterpret_cast<const char *>(&v), 4); };
    put32(2);  // version 2: 24 kernel slots (k_introduce added); version 1 images (21 slots) still load
    const int32_t dims[6] = {pm->info.nf, pm->info.nb, pm->info.np, pm->info.naux, pm->info.caps & 1, (pm->info.caps >> 1) & 1};
    img.append(reinterpret_cast<const char *>(dims), sizeof(dims));
    put32((uint32_t)pm->name.size());
    img += pm->name;
    for (int k = 0; k < kPlugKernels; ++k) {
        put32((uint32_t)pm->lowered[k].size());
        img += pm->lowered[k];
    }
    const uint64_t cs = pm->cubin.size();
    img.append(reinterpret_cast<const char *>(&cs), 8);
    img.append(pm->cubin.data(), pm->cubin.size());
    *size = (int64_t)img.size();
    if (buf) {
        if (cap < (int64_t)img.size()) return fail(GENPF_ERR_INVALID_ARG, "genpf_model_export: buffer too small");
        memcpy(buf, img.data(), img.size());
    }
    return GENPF_OK;
}

int32_t genpf_model_load_image(const void *image, int64_t size, int32_t *model_id) {
    if (!image || !model_id || size < 48) return fail(GENPF_ERR_INVALID_ARG, "genpf_model_load_image: bad arguments");
    const char *p = reinterpret_cast<const char *>(image), *end = p + size;
    if (memcmp(p, "GENPFPLG", 8) != 0) return fail(GENPF_ERR_INVALID_ARG, "not a genpf plugin image");
    p += 8;
    auto get32 = [&](uint32_t &v) {
        if (p + 4 > end) return false;
        memcpy(&v, p, 4);
        p += 4;
        return true;
    };
    uint32_t ver = 0, len = 0;
    if (!get32(ver) || (ver != 1 && ver != 2)) return fail(GENPF_ERR_INVALID_ARG, "plugin image: unknown version");
    const int n_slots = ver == 1 ? kPlugKernelsV1 : kPlugKernels;
    int32_t dims[6];
    if (p + sizeof(dims) > end) return fail(GENPF_ERR_INVALID_ARG, "plugin image truncated");
    memcpy(dims, p, sizeof(dims));
    p += sizeof(dims);
    PluginModel *pm = new PluginModel();
    bool ok = get32(len) && p + len <= end;
    if (ok) {
        pm->name.assign(p, len);
        p += len;
    }
    for (int k = 0; ok && k < n_slots; ++k) {
        ok = get32(len) && p + len <= end;
        if (ok) {
            pm->lowered[k].assign(p, len);
            p += len;
        }
    }
    uint64_t cs = 0;
    if (ok && p + 8 <= end) {
        memcpy(&cs, p, 8);
        p += 8;
        ok = p + cs <= end;
    } else {
        ok = false;
    }
    if (!ok) {
        delete pm;
        return fail(GENPF_ERR_INVALID_ARG, "plugin image truncated");
    }
    pm->cubin.assign(p, p + cs);
    pm->name_store = pm->name;
    pm->info = ModelInfo{pm->name_store.c_str(), dims[0], dims[1], dims[2], dims[3], (dims[4] ? 1 : 0) | (dims[5] ? 2 : 0)};
    return register_plugin(pm, model_id);
}

// the interface header a plugin author compiles against (the text NVRTC sees as "genpf_plugin.h" plus the tree)
int32_t genpf_model_plugin_sources(int32_t index, const char **name, const char **text) {
    if (index < 0 || index >= kNumPluginHeaders || !name || !text) return fail(GENPF_ERR_INVALID_ARG, "no such embedded header");
    *name = kPluginHeaders[index].name;
    *text = kPluginHeaders[index].text;
    return GENPF_OK;
}

}  // extern "C"
