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


SMALL = ["p1_n64_g32", "p2_n64_g32", "p2_n64_g32_mo3", "p5_n64_g32", "p1_n64_g32_cg", "p2_n64_g32_lev1_mo3",
         # section 8f options: Jacobi smoother, inhomogeneous Neumann data, F-cycles, GMRES preconditioned by MLMG
         "p2_n64_g32_jacobi", "p1_n64_g32_jacobi", "p3_n64_g32", "p2_n64_g32_fmg2", "p1_n64_g32_fmg2",
         "p2_n64_g32_gmres", "p1_n64_g32_gmres", "p2_n64_g32_lev1_levelsolve", "p1_n64_g32_lev1_levelsolve"]


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
    assert res["final_resnorm"] == pytest.approx(g["final_resnorm"], rel=1e-5, abs=1e-13 * max(g["rhsnorm0"], 1.0))


def test_reference_option_known_answers():
    """What the options must do, independent of the stored numbers: GMRES + one V-cycle converges in fewer iterations than
    plain V-cycles and reaches the same discretisation error; Jacobi needs about twice the V-cycles of red-black
    Gauss-Seidel; two F-cycles do not cost iterations."""
    g = {n: json.load(open(os.path.join(GOLDEN, f"solve_{n}.json"))) for n in
         ("p2_n64_g32", "p2_n64_g32_gmres", "p2_n64_g32_jacobi", "p2_n64_g32_fmg2")}
    assert g["p2_n64_g32_gmres"]["iters"] < g["p2_n64_g32"]["iters"]
    assert g["p2_n64_g32_gmres"]["err_inf"][0] == pytest.approx(g["p2_n64_g32"]["err_inf"][0], rel=1e-6)
    assert 1.8 * g["p2_n64_g32"]["iters"] <= g["p2_n64_g32_jacobi"]["iters"] <= 2.6 * g["p2_n64_g32"]["iters"]
    assert g["p2_n64_g32_fmg2"]["iters"] <= g["p2_n64_g32"]["iters"]


def test_reference_plotfile_roundtrip_through_fcompare(tmp_path):
    """The reference's plotfile writer and its comparison tool (both compiled into oracle/_ref) agree with themselves and
    notice a different field: the checker of tests/test_solve_gpu.py::test_plotfile_judged_by_reference_fcompare works."""
    import subprocess
    from common import REF_DRIVER
    fcompare = os.path.join(os.path.dirname(REF_DRIVER), "fcompare")
    if not os.access(fcompare, os.X_OK):
        pytest.skip("oracle/_ref/fcompare not built")
    a, b, c = str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "c")
    run_ref(mode="solve", prob_type=2, n_cell=32, max_grid_size=16, agg_grid_size=32, plotfile=a)
    run_ref(mode="solve", prob_type=2, n_cell=32, max_grid_size=16, agg_grid_size=32, plotfile=b)
    run_ref(mode="solve", prob_type=2, n_cell=32, max_grid_size=16, agg_grid_size=32, linop_maxorder=3, tol_rel=1e-4, plotfile=c)
    assert open(os.path.join(a, "Header")).read().splitlines()[:6] == ["HyperCLaw-V1.1", "4", "solution", "rhs", "exact_solution", "error"]
    assert subprocess.run([fcompare, "-r", "1e-9", a, b], capture_output=True, text=True, timeout=300).returncode == 0
    assert subprocess.run([fcompare, "-r", "1e-9", a, c], capture_output=True, text=True, timeout=300).returncode != 0


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLDEN, "meta_*.json"))) >= 9
    assert len(glob.glob(os.path.join(GOLDEN, "solve_*.json"))) >= 12
    assert len(glob.glob(os.path.join(GOLDEN, "prim_*.npz"))) >= 3
    big = json.load(open(os.path.join(GOLDEN, "solve_p2_n512_g128.json")))
    assert big["iters"] == 10 and big["err_inf"][0] == pytest.approx(3.7635e-5, rel=1e-4)
