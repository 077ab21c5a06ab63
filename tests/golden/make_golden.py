"""Generates tests/golden/golden_v1.npz.

The reference (Julia) cannot run in this image and ships no seeded vectors, so these
fixtures are produced by a SECOND, independent restatement of the reference algorithms written in
pure-Python IEEE-double arithmetic (math.exp/log = the same libm as the C oracle, no FMA, literal
loops), each function citing the reference file:line.  tests/test_oracle_kat.py requires the C oracle
to reproduce them bit-for-bit; the GPU tests then check the CUDA path against the same files.
Run:  python tests/golden/make_golden.py
"""
import math
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
M32 = 0xFFFFFFFF


# ---- Philox4x32-10 (shared RNG convention; not reference arithmetic)
def philox(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def uniform53(seed, stream, idx):
    o = philox((idx & M32, idx >> 32, stream & M32, stream >> 32), (seed & M32, seed >> 32))
    x = (o[1] << 32) | o[0]
    return (x >> 11) * 2.0 ** -53


def normal_from(seed, stream, idx):  # inputs only; any deterministic generator would do
    u1 = 1.0 - uniform53(seed, stream, 2 * idx)
    u2 = uniform53(seed, stream, 2 * idx + 1)
    return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)


# ---- Julia Base.sum (pairwise, 1024 block) and Gen.logsumexp
def jl_sum(a, lo=0, hi=None):
    hi = len(a) - 1 if hi is None else hi
    if lo == hi:
        return a[lo]
    if hi - lo < 1024:
        v = a[lo] + a[lo + 1]
        for i in range(lo + 2, hi + 1):
            v += a[i]
        return v
    mid = lo + ((hi - lo) >> 1)
    return jl_sum(a, lo, mid) + jl_sum(a, mid + 1, hi)


def logsumexp(v):
    m = max(v)
    if m == -math.inf:
        return -math.inf
    return m + math.log(jl_sum([math.exp(x - m) for x in v]))


def softmax(v):  # utils.jl:103-107
    m = max(v)
    ws = [math.exp(x - m) for x in v]
    s = jl_sum(ws)
    return [w / s for w in ws]


def safe_softmax(v):  # utils.jl:117-140
    n = len(v)
    if any(math.isnan(x) for x in v):
        return [math.nan] * n, 1
    if all(x == -math.inf for x in v):
        return [1.0 / n] * n, 2
    m = max(v)
    ws = [math.exp(x - m) for x in v]
    total = jl_sum(ws)
    if total == 0.0:
        return [1.0 / n] * n, 3
    if math.isnan(total):
        return [math.nan] * n, 4
    s = jl_sum(ws)
    return [w / s for w in ws], 0


def ess(lw):  # utils.jl:163-164 + Gen.effective_sample_size
    l = logsumexp(lw)
    return math.exp(-logsumexp([2.0 * (x - l) for x in lw]))


def sortperm_desc(keys):  # sortperm(v, rev=true): stable, ties by ascending index
    return sorted(range(len(keys)), key=lambda i: (-keys[i], i))


def stratified(w, r, order=None):  # resample.jl:159-170, literal
    n = len(w)
    order = list(range(n)) if order is None else order
    parents = [0] * n
    i_old, step, accum = 0, 1.0 / n, 0.0
    for i_new in range(1, n + 1):
        lower = (i_new - 1) / n
        if lower + step > accum:
            u = r[i_new - 1] * step
            u = u + lower
            while accum < u and i_old < n:
                accum += w[order[i_old]]
                i_old += 1
        parents[i_new - 1] = order[max(i_old, 1) - 1]
    return parents


def multinomial(w, us):  # inverse CDF, Distributions single-draw rule (resize.jl:284): first W_i > u
    out = []
    n = len(w)
    W, acc = [], 0.0
    for x in w:
        acc += x
        W.append(acc)
    for u in us:
        i = 0
        while W[i] <= u and i < n - 1:
            i += 1
        out.append(i)
    return out


def residual(w, us, n_out):  # resample.jl:96-115 / resize.jl:100-119
    parents = []
    for i, x in enumerate(w):
        c = int(math.floor(n_out * x))
        parents += [i] * c
    parents = parents[:n_out]
    nd = len(parents)
    if nd < n_out:
        rw = [n_out * x - math.floor(n_out * x) for x in w]
        s = jl_sum(rw)
        rw = [x / s for x in rw]
        parents += multinomial(rw, us[nd:n_out])
    return parents, nd


def reweight(lw, lp, parents, n_out, substate):  # resample.jl:190-218, resize.jl:424-438
    if lp is None:
        v = logsumexp(lw) - math.log(len(lw)) if substate else 0.0
        return [v] * n_out
    d = [lw[p] - lp[p] for p in parents]
    shift = (logsumexp(lw) - logsumexp(d)) if substate else (math.log(n_out) - logsumexp(d))
    return [x + shift for x in d]


def mean_var(lw, x):  # statistics.jl:13-17,48-54
    w = softmax(lw)
    mu = jl_sum([a * b for a, b in zip(w, x)])
    return mu, jl_sum([a * ((b - mu) * (b - mu)) for a, b in zip(w, x)])


def find_inv_w_threshold(w, n_particles):  # resize.jl:199-216, literal
    ws = sorted(w)
    A, B = len(ws), 0.0
    for kappa in ws:
        A -= 1
        B += kappa
        n_check = B / kappa + A if kappa != 0.0 else math.nan
        eps = math.ulp(abs(n_check)) if n_check == n_check else math.nan
        if n_check <= n_particles + eps:
            return (n_particles - A) / B
    return float(n_particles)


def optimal_resize(lw, n_particles, u_rand):  # resize.jl:149-196, literal; 0-based parents
    n = len(lw)
    w, _ = safe_softmax(lw)
    c = find_inv_w_threshold(w, n_particles)
    keep = [i for i in range(n) if c * w[i] >= 1]
    strat = [i for i in range(n) if not c * w[i] >= 1]
    n_keep, n_res = len(keep), n_particles - len(keep)
    parents = list(keep)
    if strat:
        sw, _ = safe_softmax([lw[i] for i in strat])
        step = 1 / n_res if n_res else math.inf
        u = u_rand * step
        for q, i in enumerate(strat):
            u = u - sw[q]
            if u < 0:
                parents.append(i)
                u += step
    assert len(parents) == n_particles  # resize.jl:181
    log_n_ratio = math.log(n_particles) - math.log(n)
    res_lw = logsumexp(lw) - math.log(c)
    lw_out = [lw[i] + log_n_ratio for i in keep] + [res_lw + log_n_ratio] * n_res
    return parents, lw_out, n_keep, c


def main_optimal():
    """tests/golden/golden_v2_optimal.npz: pf_optimal_resize! cases on the inputs of golden_v1 (kept in a second
    file so that golden_v1.npz stays byte-identical)."""
    out = {}
    cases = [("n100_s1", 100, 1.0, 11), ("n1000_s5", 1000, 5.0, 12), ("n2048_s2", 2048, 2.0, 13),
             ("n3000_s1", 3000, 1.0, 14)]
    for name, n, sigma, seed in cases:
        lw = [sigma * normal_from(seed, 7, i) for i in range(n)]
        for N in (max(1, n // 4), n // 2, n - 1):
            u = uniform53(seed, 3, N)
            parents, lw_out, n_keep, c = optimal_resize(lw, N, u)
            out[f"{name}/optimal_{N}/u"] = np.array(u)
            out[f"{name}/optimal_{N}/parents"] = np.array(parents)
            out[f"{name}/optimal_{N}/lw_out"] = np.array(lw_out)
            out[f"{name}/optimal_{N}/n_keep"] = np.array(n_keep)
            out[f"{name}/optimal_{N}/inv_w"] = np.array(c)
    np.savez_compressed(os.path.join(HERE, "golden_v2_optimal.npz"), **out)
    print("wrote", len(out), "arrays (optimal)")


def replicate(lw, k, interleaved):  # resize.jl:236-244: repeat(x; inner=k) | repeat(x, k)
    n = len(lw)
    parents = [i for _ in range(k) for i in range(n)] if interleaved else [i for i in range(n) for _ in range(k)]
    return parents, [lw[i] for i in parents]


def dereplicate(lw, k, interleaved, sample, us):  # resize.jl:267-297
    n_old = len(lw)
    assert n_old % k == 0
    n_new = n_old // k
    blocks = [list(range(b, n_old, n_new)) for b in range(n_new)] if interleaved else \
             [list(range(b * k, (b + 1) * k)) for b in range(n_new)]
    if not sample:
        idxs = [blk[0] for blk in blocks]
        return idxs, [lw[i] for i in idxs]
    idxs, out = [], []
    for b, blk in enumerate(blocks):
        w = softmax([lw[i] for i in blk])
        j, cp = 0, w[0]  # rand(Categorical(w)): inverse CDF, first cumulative > u (Distributions single draw)
        while cp <= us[b] and j < k - 1:
            j += 1
            cp += w[j]
        idxs.append(blk[j])
        out.append(logsumexp([lw[i] for i in blk]) - math.log(k))
    return idxs, out


def coalesce(lw, keys):  # resize.jl:309-334; groups emitted in ascending first-index order
    first, acc = {}, {}
    for i, (v, w) in enumerate(zip(keys, lw)):
        j = first.setdefault(v, i)
        acc[j] = acc.get(j, 0.0) + math.exp(w)
    n_new = len(first)
    ratio = math.log(n_new) - math.log(len(lw))
    parents = sorted(first.values())
    return parents, [math.log(acc[j]) + ratio for j in parents]


def main_resize():
    """tests/golden/golden_v3_resize.npz: pf_replicate! / pf_dereplicate! / pf_coalesce! on golden_v1's inputs."""
    out = {}
    for name, n, sigma, seed in [("n100_s1", 100, 1.0, 11), ("n3000_s1", 3000, 1.0, 14)]:
        lw = [sigma * normal_from(seed, 7, i) for i in range(n)]
        for tag, inter in (("contiguous", False), ("interleaved", True)):
            p, w = replicate(lw, 3, inter)
            out[f"{name}/replicate3_{tag}/parents"] = np.array(p)
            out[f"{name}/replicate3_{tag}/lw_out"] = np.array(w)
            us = [uniform53(seed, 5, b) for b in range(n // 5)]
            out[f"{name}/u_derep"] = np.array(us)
            for mtag, smp in (("keepfirst", False), ("sample", True)):
                p, w = dereplicate(lw, 5, inter, smp, us)
                out[f"{name}/dereplicate5_{tag}_{mtag}/parents"] = np.array(p)
                out[f"{name}/dereplicate5_{tag}_{mtag}/lw_out"] = np.array(w)
        keys = [int(uniform53(seed, 6, i) * 7) - 3 for i in range(n)]
        out[f"{name}/coalesce/keys"] = np.array(keys, dtype=np.int64)
        p, w = coalesce(lw, keys)
        out[f"{name}/coalesce/parents"] = np.array(p)
        out[f"{name}/coalesce/lw_out"] = np.array(w)
    np.savez_compressed(os.path.join(HERE, "golden_v3_resize.npz"), **out)
    print("wrote", len(out), "arrays (resize)")


def main():
    out = {}
    cases = [("n100_s1", 100, 1.0, 11), ("n1000_s5", 1000, 5.0, 12), ("n2048_s2", 2048, 2.0, 13),
             ("n3000_s1", 3000, 1.0, 14)]
    for name, n, sigma, seed in cases:
        lw = [sigma * normal_from(seed, 7, i) for i in range(n)]
        u = [uniform53(seed, 0, i) for i in range(n)]
        out[f"{name}/lw"] = np.array(lw)
        out[f"{name}/u"] = np.array(u)
        out[f"{name}/lse"] = np.array(logsumexp(lw))
        out[f"{name}/ess"] = np.array(ess(lw))
        w, kind = safe_softmax(lw)
        assert kind == 0
        out[f"{name}/w"] = np.array(w)
        x = [normal_from(seed, 9, i) for i in range(n)]
        out[f"{name}/x"] = np.array(x)
        out[f"{name}/mean_var"] = np.array(mean_var(lw, x))
        # stratified, unsorted and sorted, with and without priorities w -> w/2 (test/resample.jl:90-119)
        p = stratified(w, u)
        out[f"{name}/strat/parents"] = np.array(p)
        order = sortperm_desc(lw)
        out[f"{name}/order"] = np.array(order)
        out[f"{name}/strat_sorted/parents"] = np.array(stratified(w, u, order))
        lp = [x_ / 2 for x_ in lw]
        wp, _ = safe_softmax(lp)
        pp = stratified(wp, u)
        out[f"{name}/strat_prio/parents"] = np.array(pp)
        out[f"{name}/strat_prio/lw_out"] = np.array(reweight(lw, lp, pp, n, False))
        out[f"{name}/strat_prio/lw_out_sub"] = np.array(reweight(lw, lp, pp, n, True))
        out[f"{name}/lw_out_sub"] = np.array(reweight(lw, None, p, n, True))
        # multinomial and residual, plus resize to n/2 and 3n/2 (test/resize.jl:3-84)
        for n_out in (n, n // 2, n + n // 2):
            us = [uniform53(seed, 1, i) for i in range(n_out)]
            out[f"{name}/u_{n_out}"] = np.array(us)
            out[f"{name}/multi_{n_out}/parents"] = np.array(multinomial(w, us))
            rp, nd = residual(w, us, n_out)
            out[f"{name}/resid_{n_out}/parents"] = np.array(rp)
            out[f"{name}/resid_{n_out}/n_det"] = np.array(nd)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
    main_optimal()
    main_resize()
