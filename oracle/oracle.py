"""ctypes binding of the C oracle (oracle/genpf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
legs, never by the product package.  Indices returned here are 0-based numpy int64
(the C oracle itself speaks Julia's 1-based indices).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MULTINOMIAL, RESIDUAL, STRATIFIED = 0, 1, 2
FLAG_SORT, FLAG_SUBSTATE, FLAG_EXACT_CUMSUM = 1, 2, 4
METHODS = {"multinomial": 0, "residual": 1, "stratified": 2}

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_libs = {}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load(omp=False):
    name = "libgenpf_oracle_omp.so" if omp else "libgenpf_oracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    d, i64, i32, u32, u64 = C.c_double, C.c_int64, C.c_int32, C.c_uint32, C.c_uint64
    sig = {
        "orc_uniform53": (d, [u64, u64, u64]),
        "orc_fill_uniform53": (None, [u64, u64, i64, _vp]),
        "orc_fill_uniform_strata": (None, [u64, u64, i64, _vp]),
        "orc_sum_pairwise": (d, [_vp, i64]),
        "orc_logsumexp": (d, [_vp, i64]),
        "orc_lognorm": (None, [_vp, i64, _vp]),
        "orc_softmax": (None, [_vp, i64, _vp]),
        "orc_safe_softmax": (i32, [_vp, i64, _vp]),
        "orc_ess": (d, [_vp, i64]),
        "orc_lml_estimate": (d, [d, _vp, i64]),
        "orc_sortperm_desc": (None, [_vp, i64, _vp]),
        "orc_select_multinomial": (None, [_vp, i64, _vp, i64, _vp]),
        "orc_select_stratified": (None, [_vp, _vp, i64, _vp, u32, _vp]),
        "orc_select_stratified_search": (None, [_vp, _vp, i64, _vp, u32, _vp]),
        "orc_select_residual": (None, [_vp, i64, _vp, i64, _vp, C.POINTER(i64)]),
        "orc_cumweights": (None, [_vp, _vp, i64, u32, _vp]),
        "orc_resample": (i32, [i32, _vp, _vp, i64, i64, _vp, u32, _vp, _vp, _dp, C.POINTER(i32)]),
        "orc_mean_var": (None, [_vp, _vp, i64, _dp, _dp]),
        "orc_replicate": (None, [_vp, i64, i64, i32, _vp, _vp]),
        "orc_dereplicate": (None, [_vp, i64, i64, i32, i32, _vp, _vp, _vp]),
        "orc_coalesce": (i64, [_vp, _vp, i64, _vp, _vp]),
        "orc_om_transition": (None, [_vp, i64, _vp, _vp, d, _vp, _vp, _vp, _vp]),
        "orc_om_obs_logpdf_add": (None, [_vp, i64, _vp, d, _vp, i32]),
        "orc_om_mh": (None, [_vp, i64, _vp, _vp, _vp, _vp, d, d, _vp, _vp, _vp, _vp]),
        "orc_lg_transition": (None, [_vp, i64, _vp, _vp, _vp]),
        "orc_lg_obs_logpdf_add": (None, [_vp, i64, _vp, d, _vp, i32]),
        "orc_lg_mh": (None, [_vp, i64, _vp, _vp, d, _vp, _vp, _vp]),
        "orc_normal_logpdf": (d, [d, d, d]),
        "orc_om_filter_create": (_vp, [i64, u64]),
        "orc_om_filter_destroy": (None, [_vp]),
        "orc_om_filter_init": (None, [_vp, _vp, d, d]),
        "orc_om_filter_step": (d, [_vp, _vp, i64, d, d, d, d]),
        "orc_num_threads": (i32, []),
        "orc_set_num_threads": (None, [i32]),
    }
    for k, (res, args) in sig.items():
        fn = getattr(lib, k)
        fn.restype = res
        fn.argtypes = args
    _libs[name] = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def uniforms(seed, stream, n):
    out = np.empty(n)
    load().orc_fill_uniform53(seed, stream, n, _p(out))
    return out


def uniforms_strata(seed, stream, n):
    """Stratum uniforms as the CUDA library generates them (32-bit, four strata per Philox block)."""
    out = np.empty(n)
    load().orc_fill_uniform_strata(seed, stream, n, _p(out))
    return out


def logsumexp(v):
    v = _f(v)
    return load().orc_logsumexp(_p(v), v.size)


def lognorm(v):
    v = _f(v)
    out = np.empty_like(v)
    load().orc_lognorm(_p(v), v.size, _p(out))
    return out


def softmax(v):
    v = _f(v)
    out = np.empty_like(v)
    load().orc_softmax(_p(v), v.size, _p(out))
    return out


def safe_softmax(v):
    v = _f(v)
    out = np.empty_like(v)
    kind = load().orc_safe_softmax(_p(v), v.size, _p(out))
    return out, kind


def ess(lw):
    lw = _f(lw)
    return load().orc_ess(_p(lw), lw.size)


def sortperm_desc(keys):
    keys = _f(keys)
    out = np.empty(keys.size, dtype=np.int64)
    load().orc_sortperm_desc(_p(keys), keys.size, _p(out))
    return out - 1


def cumweights(w, order=None, exact=False):
    w = _f(w)
    o1 = None if order is None else np.ascontiguousarray(order, dtype=np.int64) + 1
    W = np.empty_like(w)
    load().orc_cumweights(_p(w), _p(o1), w.size, FLAG_EXACT_CUMSUM if exact else 0, _p(W))
    return W


def select_stratified(w, r, order=None, exact=False, search=False):
    w, r = _f(w), _f(r)
    o1 = None if order is None else np.ascontiguousarray(order, dtype=np.int64) + 1
    out = np.empty(w.size, dtype=np.int64)
    fn = load().orc_select_stratified_search if search else load().orc_select_stratified
    fn(_p(w), _p(o1), w.size, _p(r), FLAG_EXACT_CUMSUM if exact else 0, _p(out))
    return out - 1


def select_multinomial(w, u):
    w, u = _f(w), _f(u)
    out = np.empty(u.size, dtype=np.int64)
    load().orc_select_multinomial(_p(w), w.size, _p(u), u.size, _p(out))
    return out - 1


def select_residual(w, u, n_out=None):
    w, u = _f(w), _f(u)
    n_out = u.size if n_out is None else n_out
    out = np.empty(n_out, dtype=np.int64)
    nd = C.c_int64()
    load().orc_select_residual(_p(w), w.size, _p(u), n_out, _p(out), C.byref(nd))
    return out - 1, nd.value


def resample(method, lw, uniforms_, lp=None, n_out=None, sort=False, substate=False, exact=False):
    """Returns (parents0, lw_out, lml_increment, invalid_kind)."""
    lw, u = _f(lw), _f(uniforms_)
    lp = None if lp is None else _f(lp)
    n_in = lw.size
    n_out = n_in if n_out is None else n_out
    parents = np.zeros(n_out, dtype=np.int64)
    lw_out = np.zeros(n_out)
    inc, kind = C.c_double(), C.c_int32()
    flags = (FLAG_SORT if sort else 0) | (FLAG_SUBSTATE if substate else 0) | (FLAG_EXACT_CUMSUM if exact else 0)
    m = METHODS[method] if isinstance(method, str) else method
    st = load().orc_resample(m, _p(lw), _p(lp), n_in, n_out, _p(u), flags, _p(parents), _p(lw_out), C.byref(inc),
                             C.byref(kind))
    if st != 0:
        raise ValueError(f"orc_resample status {st}")
    return parents - 1, lw_out, inc.value, kind.value


def mean_var(lw, x):
    lw, x = _f(lw), _f(x)
    m, v = C.c_double(), C.c_double()
    load().orc_mean_var(_p(lw), _p(x), lw.size, C.byref(m), C.byref(v))
    return m.value, v.value


def replicate(lw, k, interleaved=False):
    lw = _f(lw)
    parents = np.empty(lw.size * k, dtype=np.int64)
    out = np.empty(lw.size * k)
    load().orc_replicate(_p(lw), lw.size, k, int(interleaved), _p(parents), _p(out))
    return parents - 1, out


def dereplicate(lw, k, interleaved=False, sample=False, u=None):
    lw = _f(lw)
    u = None if u is None else _f(u)
    parents = np.empty(lw.size // k, dtype=np.int64)
    out = np.empty(lw.size // k)
    load().orc_dereplicate(_p(lw), lw.size, k, int(interleaved), int(sample), _p(u), _p(parents), _p(out))
    return parents - 1, out


def find_inv_w_threshold(w, n_particles):
    w = _f(w)
    fn = load().orc_find_inv_w_threshold
    fn.restype = C.c_double
    fn.argtypes = [_vp, C.c_int64, C.c_int64]
    return fn(_p(w), w.size, n_particles)


def optimal_resize(lw, n_out, u_rand):
    """pf_optimal_resize! (resize.jl:149-196).  Returns dict(parents0, lw_out, n_keep, inv_w, kind, kind_strat,
    n_selected, status); status -3 = the reference's @assert (resize.jl:181) would fail."""
    lw = _f(lw)
    parents = np.zeros(n_out, dtype=np.int64)
    lw_out = np.zeros(n_out)
    n_keep, n_sel = C.c_int64(), C.c_int64()
    inv_w = C.c_double()
    kind, kind2 = C.c_int32(), C.c_int32()
    fn = load().orc_optimal_resize
    fn.restype = C.c_int32
    fn.argtypes = [_vp, C.c_int64, C.c_int64, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    st = fn(_p(lw), lw.size, n_out, float(u_rand), _p(parents), _p(lw_out), C.byref(n_keep), C.byref(inv_w),
            C.byref(kind), C.byref(kind2), C.byref(n_sel))
    return dict(parents0=parents - 1, lw_out=lw_out, n_keep=n_keep.value, inv_w=inv_w.value, kind=kind.value,
                kind_strat=kind2.value, n_selected=n_sel.value, status=st)


def coalesce(lw, keys):
    lw = _f(lw)
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    parents = np.empty(lw.size, dtype=np.int64)
    out = np.empty(lw.size)
    n_new = load().orc_coalesce(_p(lw), _p(keys), lw.size, _p(parents), _p(out))
    return parents[:n_new] - 1, out[:n_new]


class OMParams(C.Structure):
    _fields_ = [("p_stay", C.c_double), ("p_start", C.c_double), ("sigma_proc", C.c_double), ("sigma_obs", C.c_double)]


class LGParams(C.Structure):
    _fields_ = [("a", C.c_double), ("q", C.c_double), ("r", C.c_double), ("m0", C.c_double), ("s0", C.c_double)]


OM_DEFAULT = (0.75, 0.25, 0.01, 0.25)  # README.md:47-50


def normal_logpdf(x, mu, sigma):
    """Gen logpdf(normal, x, mu, sigma), vectorised over x and mu (orc_normal_logpdf per element)."""
    fn = load().orc_normal_logpdf
    xs, mus = np.broadcast_arrays(np.asarray(x, dtype=np.float64), np.asarray(mu, dtype=np.float64))
    out = np.array([fn(float(a), float(b), float(sigma)) for a, b in zip(xs.ravel(), mus.ravel())])
    return out.reshape(xs.shape) if xs.shape else float(out[0])


def om_transition(y_prev, m_prev, vel, U, Z, params=OM_DEFAULT):
    p = OMParams(*params)
    n = len(U)
    y_prev = None if y_prev is None else _f(y_prev)
    m_prev = None if m_prev is None else np.ascontiguousarray(m_prev, dtype=np.uint8)
    U, Z = _f(U), _f(Z)
    y, m = np.empty(n), np.empty(n, dtype=np.uint8)
    load().orc_om_transition(C.byref(p), n, _p(y_prev), _p(m_prev), vel, _p(U), _p(Z), _p(y), _p(m))
    return y, m


def om_obs_logpdf(y, obs, lw=None, params=OM_DEFAULT):
    p = OMParams(*params)
    y = _f(y)
    out = np.zeros(y.size) if lw is None else _f(lw).copy()
    load().orc_om_obs_logpdf_add(C.byref(p), y.size, _p(y), obs, _p(out), 1 if lw is None else 0)
    return out


def om_mh(y_pp, m_pp, y_cur, m_cur, vel, obs, U2, Z2, U3, params=OM_DEFAULT):
    p = OMParams(*params)
    n = len(U2)
    y_pp = None if y_pp is None else _f(y_pp)
    m_pp = None if m_pp is None else np.ascontiguousarray(m_pp, dtype=np.uint8)
    y, m = _f(y_cur).copy(), np.ascontiguousarray(m_cur, dtype=np.uint8).copy()
    U2, Z2, U3 = _f(U2), _f(Z2), _f(U3)
    acc = np.empty(n, dtype=np.uint8)
    load().orc_om_mh(C.byref(p), n, _p(y_pp), _p(m_pp), _p(y), _p(m), vel, obs, _p(U2), _p(Z2), _p(U3), _p(acc))
    return y, m, acc.astype(bool)


def lg_transition(x_prev, Z, params):
    p = LGParams(*params)
    x_prev, Z = _f(x_prev), _f(Z)
    out = np.empty_like(Z)
    load().orc_lg_transition(C.byref(p), Z.size, _p(x_prev), _p(Z), _p(out))
    return out


def lg_obs_logpdf(x, obs, params, lw=None):
    p = LGParams(*params)
    x = _f(x)
    out = np.zeros(x.size) if lw is None else _f(lw).copy()
    load().orc_lg_obs_logpdf_add(C.byref(p), x.size, _p(x), obs, _p(out), 1 if lw is None else 0)
    return out


def lg_mh(x_pp, x_cur, obs, Z2, U3, params):
    p = LGParams(*params)
    x_pp, Z2, U3 = _f(x_pp), _f(Z2), _f(U3)
    x = _f(x_cur).copy()
    acc = np.empty(x.size, dtype=np.uint8)
    load().orc_lg_mh(C.byref(p), x.size, _p(x_pp), _p(x), obs, _p(Z2), _p(U3), _p(acc))
    return x, acc.astype(bool)


class OMFilter:
    """CPU baseline: README loop on object_motion (bench.py only)."""

    def __init__(self, n, seed=0, omp=True, params=OM_DEFAULT, threads=None):
        self.lib = load(omp=omp)
        if threads:
            self.lib.orc_set_num_threads(int(threads))
        self.p = OMParams(*params)
        self.h = self.lib.orc_om_filter_create(n, seed)
        self.n = n

    def init(self, vel1, obs1):
        self.lib.orc_om_filter_init(self.h, C.byref(self.p), vel1, obs1)

    def step(self, t, vel_prev, obs_prev, vel_t, obs_t):
        return self.lib.orc_om_filter_step(self.h, C.byref(self.p), t, vel_prev, obs_prev, vel_t, obs_t)

    def threads(self):
        return self.lib.orc_num_threads()

    def __del__(self):
        try:
            self.lib.orc_om_filter_destroy(self.h)
        except Exception:
            pass
