"""Multi-GPU particle sharding: one process per GPU, torch.distributed for the plumbing.

``ShardedFilter`` is ONE particle filter of ``world * n_local`` particles whose slots are split contiguously
across ranks.  exchange="p2p" (default): one library call per step (genpf_shard_step_p2p) -- the per-shard totals and
the closing barrier are exchanged by the step's own kernels over peer-mapped memory (NVLink stores + epoch flags, two
synchronisation points, no NCCL inside the step).  exchange="nccl": the library launches its kernels on the filter's
CUDA stream and this class issues three tiny collectives (all-gather of 3 doubles, all-gather of one int64, a barrier)
with NCCL *on that same stream*.  Either way there is no host synchronisation inside a step, and offspring travel to
their owner as NVLink P2P stores from inside the fused step kernel.

The exchange logic is backend agnostic: tests/test_shard_host.py drives ``exchange_plan`` / ``ShardExchange``
with the gloo backend at world_size 2 on CPU.
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L


def exchange_plan(oend_all, n_total):
    """Output range [begin, end) each rank parents, from the all-gathered closing offspring counts
    (mirror of k_shard_ranges in csrc/fused.cuh; monotone by construction, covers [0, n_total) exactly)."""
    ranges, begin = [], 0
    world = len(oend_all)
    for g in range(world):
        end = max(begin, int(oend_all[g]))
        if g == world - 1:
            end = int(n_total)
        ranges.append((begin, end))
        begin = max(begin, end)
    return ranges


def closing_counts_from_totals(totals, n_total, r):
    """Host mirror of the device's exchange-free closing counts (csrc/kernels.cuh::xchg_stats_combine).
    totals: per shard (max, sum e^{v-max}) of its log-weights -- what the ranks exchange; r: the n_total stratum
    uniforms.  Every rank derives the same cumulative mass W_end(g) = prefix(g+1) (left-to-right sum of the shares) and
    the same count O_end(g) = #{i : u_i <= W_end(g)} with u_i = r_i/n + (i-1)/n (resample.jl:162); the last shard closes
    at n_total.  The scan pins each shard's counts to [O_end(g-1), O_end(g)], so exchange_plan(...) of the result is an
    exact cover that agrees with the unsharded ancestors except at cumulative-sum ties on a shard boundary."""
    totals = np.asarray(totals, dtype=np.float64).reshape(-1, 2)
    world = totals.shape[0]
    M = totals[:, 0].max()
    a = totals[:, 1] * np.exp(totals[:, 0] - M)
    S = 0.0
    for g in range(world):  # rank order, like the device
        S += a[g]
    u = np.asarray(r, dtype=np.float64) * (1.0 / n_total) + np.arange(n_total) / n_total
    oend, run = [], 0.0
    for g in range(world):
        run = run + a[g] / S
        oend.append(int(n_total) if g == world - 1 else int(np.searchsorted(u, run, side="right")))
    return oend


def cross_shard_fraction(ranges, n_local):
    """Fraction of offspring whose owner differs from their parent's rank (NVLink traffic share)."""
    total = cross = 0
    for g, (b, e) in enumerate(ranges):
        own_b, own_e = g * n_local, (g + 1) * n_local
        inside = max(0, min(e, own_e) - max(b, own_b))
        total += e - b
        cross += (e - b) - inside
    return cross / max(total, 1)


class ShardExchange:
    """The collectives of one sharded step over torch.distributed (NCCL on GPU, gloo in the CPU tests)."""

    def __init__(self, device, dist=None):
        import torch
        import torch.distributed as tdist
        self.torch = torch
        self.dist = dist or tdist
        self.rank, self.world = self.dist.get_rank(), self.dist.get_world_size()
        self.stats_local = torch.zeros(3, dtype=torch.float64, device=device)
        self.stats_all = torch.zeros(3 * self.world, dtype=torch.float64, device=device)
        self.oend_local = torch.zeros(1, dtype=torch.int64, device=device)
        self.oend_all = torch.zeros(self.world, dtype=torch.int64, device=device)
        self.token = torch.zeros(1, dtype=torch.int32, device=device)

    def gather_stats(self):
        self.dist.all_gather_into_tensor(self.stats_all, self.stats_local)

    def gather_oend(self):
        self.dist.all_gather_into_tensor(self.oend_all, self.oend_local)

    def barrier(self):
        self.dist.all_reduce(self.token)

    def all_gather_bytes(self, payload: bytes):
        out = [None] * self.world
        self.dist.all_gather_object(out, payload)
        return out


class ShardedFilter:
    """One device-plugin particle filter sharded over the ranks of the default process group."""

    def __init__(self, model, n_local, seed=0, noise="lean", exchange="p2p"):
        """exchange="p2p": the library's own peer-memory flag protocol (NVLink stores + epoch flags, no NCCL in
        the step); exchange="nccl": torch.distributed collectives issued on the filter's stream."""
        import torch
        self.exchange = exchange
        from .api import DevicePFState
        self.torch = torch
        self.lib = L.load()
        self.model = model
        self.state = DevicePFState(model, n_local, seed=seed, noise=noise)
        dev = torch.device("cuda", torch.cuda.current_device())
        sp = C.c_void_p()
        L.check(self.lib.genpf_filter_stream(self.state._h, C.byref(sp)))
        self.stream = torch.cuda.ExternalStream(sp.value)
        with torch.cuda.stream(self.stream):
            self.ex = ShardExchange(dev)
        self.rank, self.world = self.ex.rank, self.ex.world
        self.n_local, self.n_total = n_local, n_local * self.world
        nb = C.c_int64()
        L.check(self.lib.genpf_shard_ipc_export(self.state._h, None, C.byref(nb)))
        handles = C.create_string_buffer(64 * nb.value)
        L.check(self.lib.genpf_shard_ipc_export(self.state._h, handles, C.byref(nb)))
        allh = b"".join(self.ex.all_gather_bytes(handles.raw)) if self.world > 1 else handles.raw
        self._allh = C.create_string_buffer(allh, len(allh))
        L.check(self.lib.genpf_shard_attach(self.state._h, self.rank, self.world, self._allh,
                                            self.ex.stats_local.data_ptr(), self.ex.stats_all.data_ptr(),
                                            self.ex.oend_local.data_ptr(), self.ex.oend_all.data_ptr()))
        self.t = 0

    def _aux(self, t):
        return self.model.aux(t)

    def initialize(self, obs1):
        o = np.array([float(obs1)])
        L.check(self.lib.genpf_shard_initialize(self.state._h, L.ptr(o), L.ptr(self._aux(1))))
        self.state.t = self.t = 1
        # line the ranks up before the first step: the in-kernel exchange spins (bounded) on peers' flags
        self.state.sync()
        if self.world > 1:
            self.ex.dist.barrier()

    def step(self, t, obs_prev, obs_t, mh_iters=1):
        """ESS -> stratified resample -> mh(t-1) -> update(t) over the whole sharded population; asynchronous."""
        h = self.state._h
        op, ot = np.array([float(obs_prev)]), np.array([float(obs_t)])
        if self.exchange == "p2p":
            L.check(self.lib.genpf_shard_step_p2p(h, int(t), L.ptr(op), L.ptr(self._aux(t - 1)), L.ptr(ot),
                                                  L.ptr(self._aux(t)), int(mh_iters)))
            self.state.t = self.t = int(t)
            return
        with self.torch.cuda.stream(self.stream):
            L.check(self.lib.genpf_shard_begin_step(h))
            self.ex.gather_stats()
            L.check(self.lib.genpf_shard_scan(h))
            self.ex.gather_oend()
            L.check(self.lib.genpf_shard_push(h, int(t), L.ptr(op), L.ptr(self._aux(t - 1)), L.ptr(ot),
                                              L.ptr(self._aux(t)), int(mh_iters)))
            self.ex.barrier()
            L.check(self.lib.genpf_shard_finish(h))
        self.state.t = self.t = int(t)

    def step_with_noise(self, t, obs_prev, obs_t, *, mh_iters=1, uniforms=None, U2=None, Z2=None, U3=None, U1=None,
                        Z1=None):
        """Parity mode of `step`: GLOBAL-length noise columns (world * n_local), identical on every rank."""
        h = self.state._h
        op, ot = np.array([float(obs_prev)]), np.array([float(obs_t)])
        cols = [None if c is None else np.ascontiguousarray(c, dtype=np.float64) for c in (uniforms, U2, Z2, U3, U1, Z1)]
        for c in cols:
            assert c is None or c.size == self.n_total
        L.check(self.lib.genpf_shard_step_p2p_with_noise(h, int(t), L.ptr(op), L.ptr(self._aux(t - 1)), L.ptr(ot),
                                                         L.ptr(self._aux(t)), int(mh_iters), *[L.ptr(c) for c in cols]))
        self.state.t = self.t = int(t)

    def stats(self):
        """(global ESS before the last resample, accumulated log_ml_est, invalid kind) -- synchronises."""
        ess, lml, kind = C.c_double(), C.c_double(), C.c_int32()
        L.check(self.lib.genpf_shard_stats(self.state._h, C.byref(ess), C.byref(lml), C.byref(kind)))
        return ess.value, lml.value, kind.value

    def exchange_summary(self):
        """Output ranges per rank and the cross-shard offspring fraction of the last step (synchronises)."""
        if self.exchange == "p2p":
            oend = (C.c_longlong * 8)()
            err = C.c_int32()
            L.check(self.lib.genpf_shard_oend(self.state._h, oend, C.byref(err)))
            if err.value:
                raise L.GenPFError(L.ERR_STATE, "a peer timed out in the shard exchange")
            oend_all = np.array(list(oend)[: self.world])
        else:
            self.state.sync()
            oend_all = self.ex.oend_all.cpu().numpy()
        ranges = exchange_plan(oend_all, self.n_total)
        return ranges, cross_shard_fraction(ranges, self.n_local)

    def log_ml_estimate(self):
        """log_ml_est + logsumexp(lw) - log(n) over the whole population (one host all-reduce)."""
        lw = self.state.log_weights
        m = float(lw.max())
        s = float(np.exp(lw - m).sum())
        t = self.torch.tensor([m, s], dtype=self.torch.float64, device="cuda")
        allt = self.torch.zeros(2 * self.world, dtype=self.torch.float64, device="cuda")
        self.ex.dist.all_gather_into_tensor(allt, t)
        a = allt.cpu().numpy().reshape(self.world, 2)
        M = a[:, 0].max()
        S = float((a[:, 1] * np.exp(a[:, 0] - M)).sum())
        _, lml, _ = self.stats()
        return lml + M + math.log(S) - math.log(self.n_total)

    def close(self):
        self.lib.genpf_shard_detach(self.state._h)
