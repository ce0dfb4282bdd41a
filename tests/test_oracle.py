"""Pins the oracle.  oracle/_ref/ref_driver is the UNMODIFIED reference compiled from /root/reference (oracle/Makefile);
here it is checked against (1) the known answers of the reference's own test driver Tests/LinearSolvers/ABecLaplacian_C
(analytic solutions of initProb_K.H; iteration counts and error norms recorded in BASELINE.md section 2 from the
reference's GNUmake build of that test) and (2) the committed golden files."""
import glob
import json
import os

import pytest

from common import GOLDEN, have_ref, run_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/ref_driver not built")


def test_reference_known_answers_config1():
    # 128^3 MLPoisson, mgs 64: 10 iterations, max-norm error 2.8176e-4, rhs norm 166.130398
    res, _ = run_ref(mode="solve", prob_type=1, n_cell=128, max_grid_size=64, linop_maxorder=2, agg_grid_size=8)
    assert res["iters"] == 10
    assert res["err_inf"][0] == pytest.approx(2.8176e-4, rel=1e-4)
    assert res["rhsnorm0"] == pytest.approx(166.130398, rel=1e-8)
    assert res["history"][-1] / res["rhsnorm0"] == pytest.approx(6.2828e-11, rel=2e-3)


def test_reference_known_answers_abeclap():
    # config #2 at 1/8 size: second-order convergence of the discretisation error against the analytic solution
    r64, _ = run_ref(mode="solve", prob_type=2, n_cell=64, max_grid_size=32, linop_maxorder=2, agg_grid_size=32)
    r128, _ = run_ref(mode="solve", prob_type=2, n_cell=128, max_grid_size=64, linop_maxorder=2, agg_grid_size=32)
    assert 3.5 < r64["err_inf"][0] / r128["err_inf"][0] < 4.5
    assert r128["history"][-1] <= 1e-10 * max(r128["rhsnorm0"], r128["resnorm0"])


SMALL = ["p1_n64_g32", "p2_n64_g32", "p2_n64_g32_mo3", "p5_n64_g32", "p1_n64_g32_cg", "p2_n64_g32_lev1_mo3"]


@pytest.mark.parametrize("name", SMALL)
def test_reference_reproduces_golden(name):
    g = json.load(open(os.path.join(GOLDEN, f"solve_{name}.json")))
    res, _ = run_ref(mode="solve", **g["_args"])
    assert res["iters"] == g["iters"]
    assert res["cg_iters"] == g["cg_iters"]
    for a, b in zip(res["history"], g["history"]):
        # OpenMP reduction order differs between hosts/runs: rounding noise of order eps*|rhs| sits under every residual
        assert a == pytest.approx(b, rel=1e-6, abs=1e-13 * g["rhsnorm0"])
    assert res["err_inf"] == pytest.approx(g["err_inf"], rel=1e-9)


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLDEN, "meta_*.json"))) >= 9
    assert len(glob.glob(os.path.join(GOLDEN, "solve_*.json"))) >= 12
    assert len(glob.glob(os.path.join(GOLDEN, "prim_*.npz"))) >= 3
    big = json.load(open(os.path.join(GOLDEN, "solve_p2_n512_g128.json")))
    assert big["iters"] == 10 and big["err_inf"][0] == pytest.approx(3.7635e-5, rel=1e-4)
