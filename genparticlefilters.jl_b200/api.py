"""Host-side mirror of GenParticleFilters.jl's exported API over libgenpf_cuda.so.

Same names, argument meaning and error behaviour as the reference (file:line cited per
function, relative to the reference root) so tests read like the reference's own.
Two state types, as in BASELINE.json's north star:

* ``ParticleFilterState``  -- arbitrary models: ``traces`` are host objects (anything with the
  small generative-function protocol below); only ``log_weights`` and ancestor indices go to
  the GPU and traces are gathered on the host (``new_traces .= view(traces, parents)``).
* ``DevicePFState``        -- models registered as device plugins (``DeviceModel("object_motion")``,
  ``DeviceModel("lingauss1d")``): state lives in HBM, every call is a CUDA kernel sequence.

Python indices are 0-based (``parents``); the C ABI writes 1-based for Julia on request.
The host-model protocol (stand-in for Gen's GFI, reference SURVEY 8b):
  model.generate(args, observations) -> (trace, log_weight)
  trace.update(new_args, argdiffs, observations) -> (new_trace, log_weight_increment, retdiff, discard)
  kern(trace, *kern_args, **kwargs) -> (trace, accepted)            (move-accept)
  kern(trace, *kern_args, **kwargs) -> (trace, rel_log_weight)      (move-reweight)
  trace[addr] -> value
"""
import ctypes as C
import logging
import math
import warnings

import numpy as np

from . import _lib as L

log = logging.getLogger("genpf")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _fresh_seed(seed):
    """The reference draws from Julia's global RNG: every call sees fresh randomness.  The C ABI is stateless
    (Philox keyed by `seed`), so `seed=None` draws a new 64-bit key per call from numpy's global generator
    (`np.random.seed` makes a run reproducible, like `Random.seed!`); an explicit seed pins the draw."""
    if seed is None:
        return int(np.random.randint(0, np.iinfo(np.int64).max, dtype=np.int64))
    return int(seed)


# --------------------------------------------------------------------------- states
class ParticleFilterState:
    """Gen.ParticleFilterState restricted to what this path touches (initialize.jl:4-10,42-43)."""

    def __init__(self, traces, log_weights=None, log_ml_est=0.0):
        n = len(traces)
        self.traces = list(traces)
        self.new_traces = [None] * n
        self.log_weights = np.zeros(n) if log_weights is None else _f64(log_weights).copy()
        self.log_ml_est = float(log_ml_est)
        self.parents = np.arange(n, dtype=np.int64)

    def __len__(self):
        return len(self.traces)

    def __getitem__(self, idxs):  # view.jl:35-48
        return ParticleFilterSubState(self, idxs)


class ParticleFilterSubState:
    """ParticleFilterSubState (view.jl:16-22): a view on a subset of particles of `source`."""

    def __init__(self, source, idxs):
        n = len(source.traces)
        if isinstance(idxs, slice):
            idxs = np.arange(n)[idxs]
        self.idxs = np.asarray(idxs, dtype=np.int64)
        self.source = source

    def __len__(self):
        return len(self.idxs)

    @property
    def traces(self):
        return [self.source.traces[i] for i in self.idxs]

    @property
    def log_weights(self):
        return self.source.log_weights[self.idxs]

    @property
    def parents(self):
        return self.source.parents[self.idxs]


_PLUGIN_META = {}  # name -> dict(fields, bool_fields, aux_fn) of models registered from source in this process


class DeviceModel:
    """A model registered as a device plugin of libgenpf_cuda.so (include/genpf.h): one of the built-in ones
    ("object_motion", "lingauss1d") or one compiled at run time from CUDA C++ source (`DeviceModel.from_source`)."""

    def __init__(self, name, params=None):
        lib = L.load()
        mid = C.c_int32()
        L.check(lib.genpf_model_builtin(name.encode(), C.byref(mid)))
        self.name, self.model_id = name, mid.value
        nf, nb, npar, naux, caps = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        L.check(lib.genpf_model_info(mid, C.byref(nf), C.byref(nb), C.byref(npar), C.byref(naux)))
        L.check(lib.genpf_model_caps(mid, C.byref(caps)))
        self.n_f64, self.n_u8, self.n_params, self.n_aux = nf.value, nb.value, npar.value, naux.value
        self.has_proposal, self.has_translator = bool(caps.value & 1), bool(caps.value & 2)
        self.params = None if params is None else _f64(params)
        builtin = {"object_motion": ({"y": 0, "moving": 1}, ("moving",)), "lingauss1d": ({"x": 0}, ())}
        if name in builtin:
            self.fields, self.bool_fields = builtin[name]
            self._aux_fn = (lambda t: [math.sin(float(t))]) if name == "object_motion" else None
        else:
            meta = _PLUGIN_META.get(name, {})
            default = {f"f{i}": i for i in range(self.n_f64)}
            default.update({f"b{i}": self.n_f64 + i for i in range(self.n_u8)})
            self.fields = meta.get("fields") or default
            self.bool_fields = meta.get("bool_fields") or tuple(k for k, v in self.fields.items() if v >= self.n_f64)
            self._aux_fn = meta.get("aux_fn")

    @classmethod
    def from_source(cls, name, source, struct_name, *, fields=None, bool_fields=None, aux_fn=None, params=None,
                    options=None):
        """Compile `source` (CUDA C++ defining `struct_name` with the plugin interface of csrc/models.cuh) with NVRTC
        for sm_100a and register it under `name` (C ABI genpf_model_compile).  fields: {"name": index}, fp64 fields
        first; aux_fn(t) -> the model's per-step host scalars (NAUX of them) or None."""
        mid = C.c_int32()
        st = L.load().genpf_model_compile(name.encode(), source.encode(), struct_name.encode(),
                                          None if options is None else options.encode(), C.byref(mid))
        if st != L.OK:
            raise GenPFErrorException(L.load().genpf_last_error().decode("utf-8", "replace"))
        _PLUGIN_META[name] = dict(fields=fields, bool_fields=bool_fields, aux_fn=aux_fn)
        return cls(name, params)

    def export_image(self):
        """The compiled plugin as bytes (kernel names + cubin); `DeviceModel.from_image` loads it without NVRTC."""
        size = C.c_int64()
        L.check(L.load().genpf_model_export(self.model_id, None, 0, C.byref(size)))
        buf = C.create_string_buffer(size.value)
        L.check(L.load().genpf_model_export(self.model_id, buf, size.value, C.byref(size)))
        return buf.raw

    @classmethod
    def from_image(cls, image, *, fields=None, bool_fields=None, aux_fn=None, params=None):
        mid = C.c_int32()
        buf = C.create_string_buffer(image, len(image))
        L.check(L.load().genpf_model_load_image(buf, len(image), C.byref(mid)))
        # the image carries the model's name: look it up through the registry
        name = _image_name(image)
        _PLUGIN_META[name] = dict(fields=fields, bool_fields=bool_fields, aux_fn=aux_fn)
        return cls(name, params)

    def aux(self, t):
        # object_motion: vel_y = sin(t) with integer t in radians, computed by the caller (README.md:48)
        return None if self._aux_fn is None else _f64(self._aux_fn(t))


def _image_name(image):
    import struct
    off = 8 + 4 + 24
    (ln,) = struct.unpack_from("<I", image, off)
    return image[off + 4: off + 4 + ln].decode()


class DevicePFState:
    """Device-resident population (n_filters independent filters of n_particles)."""

    def __init__(self, model, n_particles, n_filters=1, seed=0, keep_history=False, noise="lean", first_filter=0):
        """first_filter: position of this handle's filter 0 inside a larger batch that is split over several handles /
        processes / GPUs (batch sharding needs no communication; the draws are those of the one big batch)."""
        lib = L.load()
        self.model = model
        flags = (L.KEEP_HISTORY if keep_history else 0) | (L.NOISE_PHILOX53 if noise == "philox53" else 0)
        h = C.c_void_p()
        L.check(lib.genpf_filter_create(model.model_id, L.ptr(model.params),
                                        0 if model.params is None else len(model.params), n_particles, n_filters,
                                        seed, flags, C.byref(h)))
        self._h = h
        self.n_filters = n_filters
        self.t = 0
        if first_filter:
            L.check(lib.genpf_filter_set_first_filter(h, int(first_filter)))
        self._ctor = dict(n_particles=int(n_particles), n_filters=int(n_filters), seed=int(seed),
                          keep_history=bool(keep_history), noise=noise, first_filter=int(first_filter))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                L.load().genpf_filter_destroy(h)
            except Exception:
                pass

    def __len__(self):
        n, f = C.c_int64(), C.c_int64()
        L.check(L.load().genpf_filter_size(self._h, C.byref(n), C.byref(f)))
        return n.value

    def _obs(self, obs):
        o = _f64(np.atleast_1d(obs))
        if o.size != self.n_filters:
            raise ValueError("need one observation per filter")
        return o

    # accessors
    @property
    def log_weights(self):
        out = np.empty(len(self) * self.n_filters)
        L.check(L.load().genpf_get_log_weights(self._h, L.ptr(out)))
        return out

    @log_weights.setter
    def log_weights(self, v):
        v = _f64(v)
        assert v.size == len(self) * self.n_filters
        L.check(L.load().genpf_set_log_weights(self._h, L.ptr(v)))

    @property
    def parents(self):
        out = np.empty(len(self) * self.n_filters, dtype=np.int64)
        L.check(L.load().genpf_get_parents(self._h, L.ptr(out), 0))
        return out

    @property
    def log_ml_est(self):
        """The accumulated field (update_lml_est!, resample.jl:178-182), without the current weights' term; one value
        per filter for a batch."""
        t, r = C.c_int64(), C.c_int64()
        acc = np.empty(self.n_filters)
        L.check(L.load().genpf_filter_get_progress(self._h, C.byref(t), C.byref(r), L.ptr(acc)))
        return float(acc[0]) if self.n_filters == 1 else acc

    def field(self, name_or_idx, t=None):
        f = self.model.fields[name_or_idx] if isinstance(name_or_idx, str) else name_or_idx
        out = np.empty(len(self) * self.n_filters)
        L.check(L.load().genpf_get_field(self._h, f, self.t if t is None else t, L.ptr(out)))
        return out

    def set_field(self, name_or_idx, t, values):
        f = self.model.fields[name_or_idx] if isinstance(name_or_idx, str) else name_or_idx
        v = _f64(values)
        L.check(L.load().genpf_set_field(self._h, f, t, L.ptr(v)))

    @property
    def accepts(self):
        out = np.empty(len(self) * self.n_filters, dtype=np.uint8)
        L.check(L.load().genpf_get_accepts(self._h, L.ptr(out)))
        return out.astype(bool)

    def sync(self):
        L.check(L.load().genpf_filter_sync(self._h))

    # checkpoint / resume (column dump of the resident window; the reference has none)
    def save(self, path):
        """Writes an .npz from which `DevicePFState.load` continues bit-identically: resident window fields,
        log-weights, accumulated log_ml_est, step / resample counters, seed and noise policy."""
        if self._ctor["keep_history"]:
            raise GenPFErrorException("checkpointing a filter that keeps its history is not supported")
        t, nres = C.c_int64(), C.c_int64()
        lml = np.empty(self.n_filters)
        L.check(L.load().genpf_filter_get_progress(self._h, C.byref(t), C.byref(nres), L.ptr(lml)))
        ctor = dict(self._ctor, n_particles=len(self))
        out = dict(model=self.model.name, params=np.zeros(0) if self.model.params is None else self.model.params,
                   t=t.value, n_resamples=nres.value, lml=lml, log_weights=self.log_weights,
                   **{"ctor_" + k: v for k, v in ctor.items()})
        for tau in (t.value - 1, t.value):
            if tau >= 1:
                for name in self.model.fields:
                    out[f"field_{name}_{tau}"] = self.field(name, tau)
        with open(path, "wb") as fh:  # np.savez on a file object keeps the caller's path verbatim (no '.npz' appended)
            np.savez(fh, **out)

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        params = z["params"]
        model = DeviceModel(str(z["model"]), None if params.size == 0 else params)
        state = cls(model, int(z["ctor_n_particles"]), n_filters=int(z["ctor_n_filters"]), seed=int(z["ctor_seed"]),
                    keep_history=False, noise=str(z["ctor_noise"]),
                    first_filter=int(z["ctor_first_filter"]) if "ctor_first_filter" in z else 0)
        t = int(z["t"])
        lml = _f64(z["lml"])
        L.check(L.load().genpf_filter_set_progress(state._h, t, int(z["n_resamples"]), L.ptr(lml)))
        state.t = t
        for tau in (t - 1, t):
            if tau >= 1:
                for name in model.fields:
                    state.set_field(name, tau, z[f"field_{name}_{tau}"])
        state.log_weights = z["log_weights"]
        return state


def logsumexp_host(v):
    """Gen.logsumexp on the GPU (C ABI genpf_logsumexp)."""
    v = _f64(v)
    out = C.c_double()
    L.check(L.load().genpf_logsumexp(L.ptr(v), v.size, 0, C.byref(out)))
    return out.value


# --------------------------------------------------------------------------- utils.jl
def get_log_norm_weights(state):
    """utils.jl:148"""
    lw = _f64(state.log_weights)
    out = np.empty_like(lw)
    L.check(L.load().genpf_normalize(L.ptr(lw), lw.size, 0, L.ptr(out), None, None, None, None))
    return out


def get_norm_weights(state):
    """utils.jl:156"""
    lw = _f64(state.log_weights)
    out = np.empty_like(lw)
    L.check(L.load().genpf_normalize(L.ptr(lw), lw.size, 0, None, L.ptr(out), None, None, None))
    return out


def effective_sample_size(state):
    """utils.jl:163-164 (per filter for a batched device state)."""
    if isinstance(state, DevicePFState):
        out = np.empty(state.n_filters)
        L.check(L.load().genpf_ess_dev(state._h, L.ptr(out)))
        return out[0] if state.n_filters == 1 else out
    lw = _f64(state.log_weights)
    out = C.c_double()
    L.check(L.load().genpf_ess(L.ptr(lw), lw.size, 0, C.byref(out)))
    return out.value


get_ess = effective_sample_size  # utils.jl:171


def log_ml_estimate(state):
    """Gen.log_ml_estimate; sub-states use the source's estimate (utils.jl:174-178)."""
    if isinstance(state, DevicePFState):
        out = np.empty(state.n_filters)
        L.check(L.load().genpf_lml_dev(state._h, L.ptr(out)))
        return out[0] if state.n_filters == 1 else out
    base = state.source.log_ml_est if isinstance(state, ParticleFilterSubState) else state.log_ml_est
    return base + logsumexp_host(state.log_weights) - math.log(len(state))


get_lml_est = log_ml_estimate  # utils.jl:186


# --------------------------------------------------------------------------- initialize.jl / update.jl
def pf_initialize(model, model_args, observations, n_particles, *, strata=None, layout="contiguous", proposal=None,
                  proposal_args=(), **kw):
    """initialize.jl:31-44; with `strata` the stratified form (initialize.jl:93-108).  Device models:
    observations = y_obs_1 (one per filter), strata = (field_name, values); host models: strata = iterable of
    constraint dicts merged into the observations."""
    if isinstance(model, DeviceModel):
        state = DevicePFState(model, n_particles, **kw)
        if proposal is not None:  # initialize.jl:46-62: the plugin's own proposal; weight = model - proposal score
            L.check(L.load().genpf_initialize_proposal(state._h, L.ptr(state._obs(observations)), L.ptr(model.aux(1)),
                                                       None, None))
        elif strata is None:
            L.check(L.load().genpf_initialize(state._h, L.ptr(state._obs(observations)), L.ptr(model.aux(1))))
        else:
            name, values = strata
            vals = _f64([float(v) for v in values])
            lay = L.LAYOUT_CONTIGUOUS if layout == "contiguous" else L.LAYOUT_INTERLEAVED
            L.check(L.load().genpf_initialize_stratified(state._h, L.ptr(state._obs(observations)), L.ptr(model.aux(1)),
                                                         model.fields[name], L.ptr(vals), vals.size, lay, None, None))
        state.t = 1
        return state
    if proposal is not None and strata is None:  # initialize.jl:46-62 (host models): proposal(*args) -> (choices, log q)
        traces, lws = [], np.empty(n_particles)
        for i in range(n_particles):
            choices, log_q = proposal(*proposal_args)
            tr, w = model.generate(model_args, {**observations, **choices})
            traces.append(tr)
            lws[i] = w - log_q
        return ParticleFilterState(traces, lws)
    if strata is not None:
        strata = list(strata)
        k_n, block = len(strata), n_particles // len(strata)
        traces, lws = [None] * n_particles, np.empty(n_particles)
        assign = [(i // block if layout == "contiguous" else i % k_n) for i in range(block * k_n)]
        assign += list(np.random.randint(0, k_n, n_particles - block * k_n))  # sample(strata, n_remaining)
        for i, k in enumerate(assign):
            traces[i], w = model.generate(model_args, {**strata[k], **observations})
            lws[i] = w + math.log(k_n)  # initialize.jl:104
        return ParticleFilterState(traces, lws)
    traces, lws = [], np.empty(n_particles)
    for i in range(n_particles):
        tr, w = model.generate(model_args, observations)
        traces.append(tr)
        lws[i] = w
    return ParticleFilterState(traces, lws)


def pf_update(state, new_args=None, argdiffs=None, observations=None, *, strata=None, layout="interleaved",
              proposal=None, proposal_args=(), translator=None, **translator_kw):
    """pf_update!, update.jl:12-25; with `strata` the stratified form (update.jl:193-210): device states take
    strata = (field_name, values), host states an iterable of constraint dicts.  `proposal`: the custom-proposal
    form (update.jl:79-96); `translator`: the generic form pf_update!(state, translator) (update.jl:35-44) -- a
    callable trace -> (new_trace, log_weight) for host states, True for a device plugin that defines `translate`."""
    if translator is not None and not isinstance(state, DevicePFState):
        src, idxs = _resolve(state)
        for i in idxs:
            src.new_traces[i], log_weight = translator(src.traces[i], **translator_kw)
            src.log_weights[i] += log_weight
        _update_refs(state)
        return state
    if isinstance(state, DevicePFState):
        t = int(new_args[0])
        if translator is not None or proposal is not None:
            fn = L.load().genpf_update_translate if translator is not None else L.load().genpf_update_proposal
            L.check(fn(state._h, t, L.ptr(state._obs(observations)), L.ptr(state.model.aux(t)), None, None))
        elif strata is None:
            L.check(L.load().genpf_update(state._h, t, L.ptr(state._obs(observations)), L.ptr(state.model.aux(t))))
        else:
            name, values = strata
            vals = _f64([float(v) for v in values])
            lay = L.LAYOUT_CONTIGUOUS if layout == "contiguous" else L.LAYOUT_INTERLEAVED
            L.check(L.load().genpf_update_stratified(state._h, t, L.ptr(state._obs(observations)),
                                                     L.ptr(state.model.aux(t)), state.model.fields[name], L.ptr(vals),
                                                     vals.size, lay, None, None))
        state.t = t
        return state
    src, idxs = _resolve(state)
    assign = None
    if strata is not None:
        strata = list(strata)
        k_n, n_p = len(strata), len(idxs)
        block = n_p // k_n
        assign = [(j // block if layout == "contiguous" else j % k_n) for j in range(block * k_n)]
        assign += list(np.random.randint(0, k_n, n_p - block * k_n))  # sample(strata, n_remaining)
    for j, i in enumerate(idxs):
        cons = observations if assign is None else {**strata[assign[j]], **observations}
        prop_w = 0.0
        if proposal is not None:  # update.jl:79-96: proposal(trace, *args) -> (choices, log q)
            choices, prop_w = proposal(src.traces[i], *proposal_args)
            cons = {**cons, **choices}
        new_tr, incr, _, discard = src.traces[i].update(new_args, argdiffs, cons)
        incr = incr - prop_w
        if discard:
            raise GenPFErrorException(f"Choices were updated or deleted: {discard}")  # update.jl:18-20
        src.new_traces[i] = new_tr
        src.log_weights[i] += incr + (0.0 if assign is None else math.log(len(strata)))
    _update_refs(state)
    return state


class GenPFErrorException(RuntimeError):
    """Julia's ErrorException (error(...)) on this path."""


def _resolve(state):
    if isinstance(state, ParticleFilterSubState):
        return state.source, state.idxs
    return state, np.arange(len(state.traces))


def _update_refs(state):
    """update_refs!, utils.jl:10-20: swap for a full state, copy-back for a view."""
    if isinstance(state, ParticleFilterSubState):
        for i in state.idxs:
            state.source.traces[i] = state.source.new_traces[i]
    else:
        state.traces, state.new_traces = state.new_traces, state.traces


# --------------------------------------------------------------------------- resample.jl / resize.jl
def _emit_check(kind, check):
    if kind == L.VALID:
        return
    if check is True:
        raise GenPFErrorException("Invalid weights.")  # resample.jl:55,92,151
    if check == "warn":
        warnings.warn(L.WARNINGS[kind])  # utils.jl:120,124,132,135


def pf_resample(state, method="multinomial", *, priority_fn=None, check="warn", sort_particles=True,
                uniforms=None, seed=None, n_out=None):
    """pf_resample!, resample.jl:19-30 (dispatch), 48-65 / 85-120 / 143-175.

    `uniforms` (optional) are the reference RNG's draws exported per output slot / stratum; without
    them the library draws Philox4x32-10(seed) on the device.
    """
    if method not in L.METHODS:
        raise GenPFErrorException(f"Resampling method {method} not recognized.")  # resample.jl:28
    m = L.METHODS[method]
    lib = L.load()
    flags = L.SORT_PARTICLES if (method == "stratified" and sort_particles) else 0
    if isinstance(state, DevicePFState):
        kinds = np.zeros(state.n_filters, dtype=np.int32)
        prio_kind, prio_param, prio_col = L.PRIO_NONE, 1.0, None
        if priority_fn is not None:
            if isinstance(priority_fn, (int, float)):
                prio_kind, prio_param = L.PRIO_SCALE, float(priority_fn)  # w -> alpha*w
            else:
                prio_kind, prio_col = L.PRIO_COLUMN, _f64(priority_fn(state.log_weights))
        u = None if uniforms is None else _f64(uniforms)
        st = lib.genpf_resample_dev(state._h, m, prio_kind, prio_param, L.ptr(prio_col),
                                    0 if n_out is None else n_out, flags | (L.CHECK if check is True else 0),
                                    L.ptr(u), L.ptr(kinds))
        if st == L.ERR_INVALID_WEIGHTS:
            raise GenPFErrorException("Invalid weights.")
        L.check(st)
        for k in kinds:
            _emit_check(int(k), check)
        return state
    src, idxs = _resolve(state)
    sub = isinstance(state, ParticleFilterSubState)
    n_in = len(idxs)
    n_out = n_in if n_out is None else int(n_out)
    lw = _f64(src.log_weights[idxs])
    lp = None if priority_fn is None else _f64(priority_fn(lw))
    u = None if uniforms is None else _f64(uniforms)
    parents = np.empty(n_out, dtype=np.int64)
    lw_out = np.empty(n_out)
    lml_inc, kind = C.c_double(), C.c_int32()
    st = lib.genpf_resample(m, L.ptr(lw), L.ptr(lp), n_in, n_out, L.ptr(u), _fresh_seed(seed),
                            flags | (L.SUBSTATE if sub else 0) | (L.CHECK if check is True else 0),
                            L.ptr(parents), L.ptr(lw_out), C.byref(lml_inc), C.byref(kind))
    if st == L.ERR_INVALID_WEIGHTS:
        raise GenPFErrorException("Invalid weights.")
    L.check(st)
    _emit_check(kind.value, check)
    if kind.value in (L.INV_NAN_INPUT, L.INV_NAN_TOTAL):
        # the reference crashes inside Categorical/floor(Int, NaN) here
        raise GenPFErrorException("NaN weights: cannot resample.")
    if sub:
        # local parents -> positions of the view; traces copied back in place (utils.jl:17-20)
        new = [src.traces[idxs[p]] for p in parents]
        for j, i in enumerate(idxs):
            src.new_traces[i] = new[j]
            src.traces[i] = new[j]
        src.parents[idxs] = parents
        src.log_weights[idxs] = lw_out
    else:
        src.log_ml_est += lml_inc.value  # update_lml_est!, resample.jl:178-182
        src.new_traces = [src.traces[p] for p in parents]  # new_traces .= view(traces, parents)
        src.parents = parents
        src.log_weights = lw_out
        src.traces, src.new_traces = src.new_traces, [None] * n_out  # update_refs! (+ resize.jl:441-449)
    return state


def pf_multinomial_resample(state, **kw):
    return pf_resample(state, "multinomial", **kw)


def pf_residual_resample(state, **kw):
    return pf_resample(state, "residual", **kw)


def pf_stratified_resample(state, **kw):
    return pf_resample(state, "stratified", **kw)


def pf_resize(state, n_particles, method="multinomial", **kw):
    """pf_resize!, resize.jl:16-27 (:multinomial 46-67, :residual 87-124)."""
    if method == "optimal":
        return pf_optimal_resize(state, n_particles, **kw)
    if method not in ("multinomial", "residual"):
        raise GenPFErrorException(f"Resampling method {method} not recognized.")  # resize.jl:25
    if isinstance(state, ParticleFilterSubState):
        raise TypeError("pf_resize! takes a full ParticleFilterState")
    return pf_resample(state, method, n_out=n_particles, **kw)


def pf_optimal_resize(state, n_particles, *, check="warn", uniform=None, seed=None):
    """pf_optimal_resize!, resize.jl:149-196 (Fearnhead-Clifford optimal resampling; n_particles <= current size).
    `uniform` stands for the single rand() of resize.jl:171 (None: the library's Philox draw)."""
    if isinstance(state, ParticleFilterSubState):
        raise TypeError("pf_resize! takes a full ParticleFilterState")
    u = None if uniform is None else C.byref(C.c_double(float(uniform)))
    n_keep, inv_w = C.c_int64(), C.c_double()
    kinds = (C.c_int32 * 2)()
    flags = L.CHECK if check is True else 0
    if isinstance(state, DevicePFState):
        st = L.load().genpf_optimal_resize_dev(state._h, n_particles, u, flags, C.byref(n_keep), C.byref(inv_w), kinds)
    else:
        lw = _f64(state.log_weights)
        parents = np.empty(n_particles, dtype=np.int64)
        lw_out = np.empty(n_particles)
        st = L.load().genpf_optimal_resize(L.ptr(lw), lw.size, n_particles, u, _fresh_seed(seed), flags, L.ptr(parents),
                                           L.ptr(lw_out), C.byref(n_keep), C.byref(inv_w), kinds)
    if st == L.ERR_INVALID_WEIGHTS:
        raise GenPFErrorException("Invalid weights.")
    if st == L.ERR_ASSERT:
        raise AssertionError(L.load().genpf_last_error().decode("utf-8", "replace"))  # @assert, resize.jl:181,183
    L.check(st)
    _emit_check(kinds[0], check)
    if kinds[1] != kinds[0]:
        _emit_check(kinds[1], check)
    if not isinstance(state, DevicePFState):
        _apply_resize(state, parents, lw_out)
    return state


def pf_multinomial_resize(state, n_particles, **kw):
    return pf_resize(state, n_particles, "multinomial", **kw)


def pf_residual_resize(state, n_particles, **kw):
    return pf_resize(state, n_particles, "residual", **kw)


def pf_replicate(state, n_replicates, *, layout="contiguous"):
    """pf_replicate!, resize.jl:236-244."""
    lay = L.LAYOUT_CONTIGUOUS if layout == "contiguous" else L.LAYOUT_INTERLEAVED
    if isinstance(state, DevicePFState):
        L.check(L.load().genpf_replicate(state._h, n_replicates, lay))
        return state
    n = len(state.traces)
    lw = _f64(state.log_weights)
    parents = np.empty(n * n_replicates, dtype=np.int64)
    lw_out = np.empty(n * n_replicates)
    L.check(L.load().genpf_replicate_host(L.ptr(lw), n, n_replicates, lay, 0, L.ptr(parents), L.ptr(lw_out)))
    _apply_resize(state, parents, lw_out)
    return state


def pf_dereplicate(state, n_replicates, *, layout="contiguous", method="keepfirst", uniforms=None, seed=None):
    """pf_dereplicate!, resize.jl:267-297."""
    lay = L.LAYOUT_CONTIGUOUS if layout == "contiguous" else L.LAYOUT_INTERLEAVED
    meth = L.KEEPFIRST if method == "keepfirst" else L.SAMPLE
    u = None if uniforms is None else _f64(uniforms)
    if isinstance(state, DevicePFState):
        L.check(L.load().genpf_dereplicate(state._h, n_replicates, lay, meth, L.ptr(u)))
        return state
    n = len(state.traces)
    assert n % n_replicates == 0  # resize.jl:270
    lw = _f64(state.log_weights)
    parents = np.empty(n // n_replicates, dtype=np.int64)
    lw_out = np.empty(n // n_replicates)
    L.check(L.load().genpf_dereplicate_host(L.ptr(lw), n, n_replicates, lay, meth, L.ptr(u), _fresh_seed(seed), 0,
                                            L.ptr(parents), L.ptr(lw_out)))
    _apply_resize(state, parents, lw_out)
    return state


def pf_coalesce(state, *, by=None):
    """pf_coalesce!, resize.jl:309-334.  `by(trace)` must return something hashable."""
    if isinstance(state, DevicePFState):
        if by is not None:
            raise GenPFErrorException("pf_coalesce!(device state): particles are grouped by their resident window "
                                      "(all fields of slices t-1 and t); a custom `by` is not supported")
        n_new = C.c_int64()
        L.check(L.load().genpf_coalesce(state._h, C.byref(n_new)))
        return state
    if not state.traces:
        return state
    vals = [tr if by is None else by(tr) for tr in state.traces]
    codes = {}
    keys = np.array([codes.setdefault(v, len(codes)) for v in vals], dtype=np.int64)
    n = len(keys)
    lw = _f64(state.log_weights)
    parents = np.empty(n, dtype=np.int64)
    lw_out = np.empty(n)
    n_new = C.c_int64()
    L.check(L.load().genpf_coalesce_host(L.ptr(lw), L.ptr(keys), n, 0, L.ptr(parents), L.ptr(lw_out), C.byref(n_new)))
    _apply_resize(state, parents[: n_new.value].copy(), lw_out[: n_new.value].copy())
    return state


def pf_introduce(state, model, model_args, observations, n_particles, *, proposal=None, proposal_args=()):
    """pf_introduce!, resize.jl:351-421 (host models): append `n_particles` freshly generated traces.  The
    accumulated log_ml_est is folded into the existing weights first (resize.jl:362-365); `model` / `model_args`
    None reuse those of the first trace (trace.model, trace.args).  With a proposal (callable returning
    (choices, log_q)) the new weight is model_weight - log_q (resize.jl:410-413)."""
    if isinstance(state, DevicePFState):
        # device plugins: `observations` is the history [y_1, ..., y_t] (each a scalar or one value per filter); every
        # new trace is a whole chain generated under it (genpf_introduce); proposal=True uses the plugin's proposal
        t = state.t
        obs = _f64(np.stack([state._obs(o) for o in observations]))
        if obs.shape[0] != t:
            raise ValueError(f"pf_introduce! on a device filter at time {t} needs {t} observations, got {obs.shape[0]}")
        m = state.model
        aux = _f64(np.concatenate([m.aux(tau) for tau in range(1, t + 1)])) if m.n_aux else None
        L.check(L.load().genpf_introduce(state._h, int(n_particles), L.ptr(obs), L.ptr(aux), 1 if proposal else 0, None, None))
        return state
    if isinstance(state, ParticleFilterSubState):
        raise TypeError("pf_introduce! takes a full ParticleFilterState")
    model = state.traces[0].model if model is None else model
    model_args = state.traces[0].args if model_args is None else model_args
    if state.log_ml_est != 0.0:
        state.log_weights = state.log_weights + state.log_ml_est
        state.log_ml_est = 0.0
    n_old = len(state.traces)
    new_lw = np.empty(n_particles)
    for i in range(n_particles):
        if proposal is None:
            tr, w = model.generate(model_args, observations)
        else:
            choices, log_q = proposal(*proposal_args)
            tr, w = model.generate(model_args, {**observations, **choices})
            w = w - log_q
        state.traces.append(tr)
        new_lw[i] = w
    state.log_weights = np.concatenate([state.log_weights, new_lw])
    state.parents = np.concatenate([state.parents, np.zeros(n_particles, dtype=np.int64)])  # resize!: unspecified
    state.new_traces = [None] * (n_old + n_particles)
    return state


def choiceproduct(*choices):
    """choiceproduct, utils.jl:56-95: iterator over constraint dicts, the Cartesian product of (addr, values) pairs
    (or of a dict addr -> values); the strata argument of the stratified pf_initialize."""
    import itertools
    if len(choices) == 1 and isinstance(choices[0], dict):
        choices = tuple(choices[0].items())
    return (dict(cs) for cs in itertools.product(*[[(addr, v) for v in vals] for addr, vals in choices]))


def get_traces(state):
    """Gen.get_traces"""
    return list(state.traces)


def get_log_weights(state):
    """Gen.get_log_weights"""
    return _f64(state.log_weights).copy()


def sample_unweighted_traces(state, n_samples, *, uniforms=None, seed=None):
    """Gen.sample_unweighted_traces / its sub-state method (utils.jl:189-194): n_samples traces drawn with
    replacement in proportion to the normalised weights (inverse-CDF draws on the GPU); the state is not changed."""
    lw = _f64(state.log_weights)
    u = None if uniforms is None else _f64(uniforms)
    parents = np.empty(n_samples, dtype=np.int64)
    lw_out = np.empty(n_samples)
    inc, kind = C.c_double(), C.c_int32()
    L.check(L.load().genpf_resample(L.METHODS["multinomial"], L.ptr(lw), None, lw.size, n_samples, L.ptr(u),
                                    _fresh_seed(seed),
                                    L.SUBSTATE if n_samples == lw.size else 0, L.ptr(parents), L.ptr(lw_out),
                                    C.byref(inc), C.byref(kind)))
    traces = state.traces
    return [traces[p] for p in parents]


def _apply_resize(state, parents, lw_out):
    state.new_traces = [state.traces[p] for p in parents]
    state.parents = parents
    state.log_weights = lw_out
    state.traces, state.new_traces = state.new_traces, [None] * len(parents)


# --------------------------------------------------------------------------- rejuvenate.jl
def pf_rejuvenate(state, kern, kern_args=(), n_iters=1, *, method="move", **kwargs):
    """pf_rejuvenate!, rejuvenate.jl:18-27."""
    if method == "move":
        return pf_move_accept(state, kern, kern_args, n_iters, **kwargs)
    if method == "reweight":
        return pf_move_reweight(state, kern, kern_args, n_iters, **kwargs)
    raise GenPFErrorException(f"Method {method} not recognized.")  # rejuvenate.jl:25


def mh(*a, **k):
    """Marker for Gen.mh on a device state: pf_rejuvenate(state, mh, (tau, obs_tau))."""
    raise TypeError("mh is only a marker for device states")


def pf_move_accept(state, kern, kern_args=(), n_iters=1, **kwargs):
    """pf_move_accept!, rejuvenate.jl:40-53."""
    if isinstance(state, DevicePFState):
        if kern is not mh:
            raise TypeError("device states rejuvenate with the built-in mh kernel")
        tau, obs = kern_args
        if not 0 <= int(n_iters) <= 256:
            raise GenPFErrorException("device mh rejuvenation supports at most 256 iterations per call")
        acc = np.zeros(state.n_filters, dtype=np.int64)
        L.check(L.load().genpf_rejuvenate_mh(state._h, int(tau), L.ptr(state._obs(obs)),
                                             L.ptr(state.model.aux(tau)), n_iters, L.ptr(acc)))
        state.last_n_accept = acc
        return state
    src, idxs = _resolve(state)
    for i in idxs:
        trace = src.traces[i]
        for _ in range(n_iters):
            trace, accept = kern(trace, *kern_args, **kwargs)
            log.debug("Accepted: %s", accept)  # rejuvenate.jl:47
        src.new_traces[i] = trace
    _update_refs(state)
    return state


def move_reweight(*a, **k):
    """Marker for move_reweight(trace, selection) on a device state: pf_move_reweight(state, move_reweight, (tau, obs))."""
    raise TypeError("move_reweight is only a marker for device states")


def pf_move_reweight(state, kern, kern_args=(), n_iters=1, **kwargs):
    """pf_move_reweight!, rejuvenate.jl:74-90; device states: the built-in regenerate-and-reweight of slice tau
    (`proposal=True`: move_reweight(trace, proposal, proposal_args), rejuvenate.jl:134-148, with the plugin's proposal)."""
    if isinstance(state, DevicePFState):
        if kern is not move_reweight:
            raise TypeError("device states reweight with the built-in move_reweight kernel")
        tau, obs = kern_args
        if kwargs.get("proposal"):
            L.check(L.load().genpf_rejuvenate_reweight_proposal(state._h, int(tau), L.ptr(state._obs(obs)),
                                                                L.ptr(state.model.aux(tau)), n_iters, None, None))
            return state
        L.check(L.load().genpf_rejuvenate_reweight(state._h, int(tau), L.ptr(state._obs(obs)),
                                                   L.ptr(state.model.aux(tau)), n_iters))
        return state
    src, idxs = _resolve(state)
    for i in idxs:
        trace = src.traces[i]
        for _ in range(n_iters):
            trace, rel_weight = kern(trace, *kern_args, **kwargs)
            src.log_weights[i] += rel_weight
            log.debug("Rel. Weight: %s", rel_weight)  # rejuvenate.jl:83
        src.new_traces[i] = trace
    _update_refs(state)
    return state


# --------------------------------------------------------------------------- statistics.jl
def _mean_var(state, addr):
    lib = L.load()
    if isinstance(state, DevicePFState):
        tau, name = addr  # (t, :field), like `5 => :moving`
        m, v = np.empty(state.n_filters), np.empty(state.n_filters)
        L.check(lib.genpf_mean_var(state._h, state.model.fields[name], int(tau), L.ptr(m), L.ptr(v)))
        return (m[0], v[0]) if state.n_filters == 1 else (m, v)
    lw = _f64(state.log_weights)
    x = _f64([tr[addr] for tr in state.traces])
    m, v = C.c_double(), C.c_double()
    L.check(lib.genpf_weighted_mean_var(L.ptr(lw), L.ptr(x), lw.size, 0, C.byref(m), C.byref(v)))
    return m.value, v.value


def mean(state, addr):
    """Statistics.mean(state, addr), statistics.jl:13-14."""
    return _mean_var(state, addr)[0]


def var(state, addr):
    """Statistics.var(state, addr), statistics.jl:48-50 (uncorrected, two-pass)."""
    return _mean_var(state, addr)[1]


def proportionmap(state, addr, f=None, max_values=1 << 16):
    """StatsBase.proportionmap(state, addr) / proportionmap(f, state, addr), statistics.jl:91-130: dict mapping each
    distinct value at `addr` to the sum of normalised weights of the particles holding it."""
    lib = L.load()
    n_unique = C.c_int64()
    if isinstance(state, DevicePFState):
        if f is not None:
            raise GenPFErrorException("proportionmap(f, ...) on a device state: apply f to the keys of the result")
        tau, name = addr
        vals, props = np.empty(max_values), np.empty(max_values)
        L.check(lib.genpf_proportionmap(state._h, state.model.fields[name], int(tau), vals.ctypes.data_as(L._dp),
                                        props.ctypes.data_as(L._dp), max_values, C.byref(n_unique)))
        if n_unique.value > max_values:
            warnings.warn(f"proportionmap: {n_unique.value} distinct values, only the first {max_values} returned")
        g = min(n_unique.value, max_values)
        is_bool = name in getattr(state.model, "bool_fields", ())
        return {(bool(v) if is_bool else float(v)): float(p) for v, p in zip(vals[:g], props[:g])}
    vs = [tr[addr] for tr in state.traces]
    if f is not None:
        vs = [f(v) for v in vs]
    codes = {}
    keys = np.array([codes.setdefault(v, len(codes)) for v in vs], dtype=np.int64)
    lw = _f64(state.log_weights)
    first, props = np.empty(lw.size, dtype=np.int64), np.empty(lw.size)
    L.check(lib.genpf_proportionmap_host(L.ptr(lw), L.ptr(keys), lw.size, 0, L.ptr(first), L.ptr(props),
                                         C.byref(n_unique)))
    return {vs[i]: float(p) for i, p in zip(first[:n_unique.value], props[:n_unique.value])}


def pf_step(state, t, obs_prev, obs_t, *, method="stratified", ess_thresh=0.5, mh_iters=1, return_ess=True):
    """One iteration of the README loop (README.md:66-77) in a single C-ABI call on a device state."""
    m = state.model
    ess = np.empty(state.n_filters) if return_ess else None
    L.check(L.load().genpf_step(state._h, int(t), L.ptr(state._obs(obs_prev)), L.ptr(m.aux(t - 1)),
                                L.ptr(state._obs(obs_t)), L.ptr(m.aux(t)), L.METHODS[method], float(ess_thresh),
                                int(mh_iters), L.ptr(ess)))
    state.t = int(t)
    return ess


def pf_step_with_noise(state, t, obs_prev, obs_t, *, method="stratified", ess_thresh=1.0, mh_iters=1, uniforms=None,
                       U2=None, Z2=None, U3=None, U1=None, Z1=None):
    """The README iteration in parity mode: every draw supplied as a column indexed by output particle
    (uniforms=None: the library's own Philox stratum / inverse-CDF draws).  ess_thresh >= 1 resamples every filter,
    below 1 each filter of a batch decides for itself (ess < ess_thresh * n)."""
    m = state.model
    cols = [None if c is None else _f64(c) for c in (uniforms, U2, Z2, U3, U1, Z1)]
    L.check(L.load().genpf_step_with_noise(state._h, int(t), L.ptr(state._obs(obs_prev)), L.ptr(m.aux(t - 1)),
                                           L.ptr(state._obs(obs_t)), L.ptr(m.aux(t)), L.METHODS[method],
                                           float(ess_thresh), int(mh_iters), *[L.ptr(c) for c in cols]))
    state.t = int(t)
    return state


def pf_run(state, t_first, observations, *, method="stratified", ess_thresh=0.5, mh_iters=1, graph=False):
    """The README loop (README.md:66-77) for steps t_first .. t_first + T - 1 in one asynchronous C-ABI call
    (genpf_run_steps): `observations` has T + 1 rows, row 0 being the observation of step t_first - 1.  graph=True
    replays steps 2.. as one CUDA graph.  Nothing is copied back; read ESS / fields afterwards."""
    m = state.model
    obs = _f64(np.asarray(observations, dtype=np.float64).reshape(-1, state.n_filters))
    T = obs.shape[0] - 1
    aux = None
    if m.n_aux:
        aux = _f64(np.concatenate([m.aux(t_first - 1 + r) for r in range(T + 1)]))
    L.check(L.load().genpf_run_steps(state._h, int(t_first), T, L.ptr(obs), L.ptr(aux), L.METHODS[method],
                                     float(ess_thresh), int(mh_iters), L.RUN_GRAPH if graph else 0))
    state.t = int(t_first) + T - 1
    return state
