"""Parity against REAL reference output (tests/golden/reference_julia.npz, produced by tools/export_golden.jl +
tools/import_golden.py on a machine with Julia/Gen).  The build image has no Julia, so the file is absent there and
these tests skip; committing it turns "parity unpinned" (oracle/genpf_oracle.h) into a pinned oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from util import check_parents, strat_u

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_julia.npz")
needs_file = pytest.mark.skipif(not os.path.exists(PATH), reason="no reference_julia.npz (needs Julia: tools/export_golden.jl)")
RTOL = 1e-10


def cases(z):
    return sorted({k.split("/")[0] for k in z.files})


@needs_file
def test_oracle_against_julia(orc):
    z = np.load(PATH)
    for c in cases(z):
        lw = z[f"{c}/lw"]
        n = lw.size
        assert orc.logsumexp(lw) == pytest.approx(z[f"{c}/lse"][0], rel=RTOL)
        assert orc.ess(lw) == pytest.approx(z[f"{c}/ess"][0], rel=RTOL)
        np.testing.assert_allclose(orc.softmax(lw), z[f"{c}/norm_w"], rtol=RTOL)
        for tag, sort in (("strat", False), ("strat_sorted", True)):
            p, _, inc, _ = orc.resample("stratified", lw, z[f"{c}/{tag}/r"], sort=sort)
            np.testing.assert_array_equal(p, z[f"{c}/{tag}/parents"])  # same sequential arithmetic: bit-exact
            assert inc == pytest.approx(z[f"{c}/{tag}/lml"][0], rel=RTOL, abs=1e-12)
        for key in [k for k in z.files if k.startswith(f"{c}/optimal_") and k.endswith("/u")]:
            N = int(key.split("/")[1].split("_")[1])
            r = orc.optimal_resize(lw, N, float(z[key][0]))
            assert r["inv_w"] == pytest.approx(z[f"{c}/optimal_{N}/inv_w"][0], rel=1e-12)
            np.testing.assert_array_equal(r["parents0"], z[f"{c}/optimal_{N}/parents"])
            np.testing.assert_allclose(r["lw_out"], z[f"{c}/optimal_{N}/lw_out"], rtol=RTOL)


@needs_file
@pytest.mark.gpu
def test_cuda_against_julia(g, orc):
    from test_gpu_host_path import raw_optimal_resize, raw_resample
    z = np.load(PATH)
    for c in cases(z):
        lw = z[f"{c}/lw"]
        n = lw.size
        W_ref = orc.cumweights(orc.softmax(lw))
        for tag, flags in (("strat", 0), ("strat_sorted", g._lib.SORT_PARTICLES)):
            r = z[f"{c}/{tag}/r"]
            st, p, lw_out, inc, kind = raw_resample(g, "stratified", lw, r, flags=flags)
            assert st == 0 and kind == 0
            if flags:
                order = orc.sortperm_desc(lw)
                check_parents(p, z[f"{c}/{tag}/parents"], orc.cumweights(orc.softmax(lw), order), strat_u(r, n), order=order)
            else:
                check_parents(p, z[f"{c}/{tag}/parents"], W_ref, strat_u(r, n))
            assert inc == pytest.approx(z[f"{c}/{tag}/lml"][0], rel=RTOL, abs=1e-12)
        for key in [k for k in z.files if k.startswith(f"{c}/optimal_") and k.endswith("/u")]:
            N = int(key.split("/")[1].split("_")[1])
            st, p, lw_out, n_keep, inv_w, _ = raw_optimal_resize(g, lw, N, float(z[key][0]))
            assert st == 0 and inv_w == pytest.approx(z[f"{c}/optimal_{N}/inv_w"][0], rel=1e-9)
            ref = z[f"{c}/optimal_{N}/parents"]
            assert np.mean(p == ref) > 0.999
            np.testing.assert_allclose(lw_out, z[f"{c}/optimal_{N}/lw_out"], rtol=RTOL)
