"""Drop-in at the SOURCE level (SURVEY 8b, C++ surface): the reference's own test driver
Tests/LinearSolvers/ABecLaplacian_C (main.cpp, MyTest.cpp, initProb.cpp, MyTestPlotfile.cpp) is compiled UNMODIFIED, from
where it lies under /root/reference, against this library's headers and linked with libamrex_b200.so
(scripts/build_reference_driver.sh).  Building needs no GPU; the executable travels to the GPU box in build/refdriver/.

CPU: the driver compiles, links and refuses to run without a device; ParmParse (host only) is unit-tested.
GPU: the driver solves the reference's own input decks; iteration counts are compared with the committed golden runs of
the reference and the plotfile it writes is judged by the reference's fcompare."""
import json
import os
import subprocess

import pytest

from common import GOLDEN, REF_DRIVER, have_ref, run_ref

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(REPO, "build", "refdriver", "ABecLaplacian_C.b200.ex")
REF_SRC = "/root/reference/Tests/LinearSolvers/ABecLaplacian_C"


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference sources not mounted")
def test_reference_driver_compiles_unmodified_and_links():
    out = subprocess.run([os.path.join(REPO, "scripts", "build_reference_driver.sh")], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert os.access(EXE, os.X_OK)
    import torch
    if not torch.cuda.is_available():       # no CPU fallback: the driver must stop at amrex::Initialize, loudly
        run = subprocess.run([EXE, "n_cell=32"], capture_output=True, text=True, timeout=120)
        assert run.returncode != 0 and "no CUDA device" in (run.stdout + run.stderr)


def test_parmparse_host_only(tmp_path):
    exe = str(tmp_path / "pp_test")
    inc = [f"-I{REPO}/amrex_b200/csrc/{d}" for d in ("compat", "base", "mlmg")] + [f"-I{REPO}/include"]
    cc = subprocess.run(["nvcc", "-std=c++17", "-x", "cu", *inc, os.path.join(REPO, "tests", "cpp", "parmparse_test.cpp"), "-o", exe,
                         f"-L{REPO}/amrex_b200/lib", "-lamrex_b200", "-Xlinker", "-rpath", "-Xlinker", f"{REPO}/amrex_b200/lib"],
                        capture_output=True, text=True, timeout=600)
    assert cc.returncode == 0, cc.stderr[-3000:]
    run = subprocess.run([exe, str(tmp_path / "inputs")], capture_output=True, text=True, timeout=60)
    assert run.returncode == 0 and "PARMPARSE OK" in run.stdout, run.stdout + run.stderr


def _run_driver(tmp_path, **kw):
    args = [EXE] + [f"{k}={v}" for k, v in kw.items()]
    return subprocess.run(args, capture_output=True, text=True, timeout=180, cwd=str(tmp_path))


@pytest.mark.gpu
@pytest.mark.skipif(not os.access(EXE, os.X_OK), reason="build/refdriver/ABecLaplacian_C.b200.ex not built (scripts/build_reference_driver.sh)")
@pytest.mark.parametrize("prob_type,golden,max_level,maxorder,composite", [
    (1, "p1_n64_g32", 0, 2, 1), (2, "p2_n64_g32", 0, 2, 1), (2, "p2_n64_g32_lev1_mo3", 1, 3, 1),
    # level-by-level: the fine level is solved as a single-level operator with setCoarseFineBC data (MyTest.cpp:104-141, 226-279)
    (2, "p2_n64_g32_lev1_levelsolve", 1, 2, 0), (1, "p1_n64_g32_lev1_levelsolve", 1, 2, 0)])
def test_reference_driver_runs_on_gpu(tmp_path, prob_type, golden, max_level, maxorder, composite):
    """All five decks ran clean on hardware in round 1 (GPUTEST_r01.json), so every deviation is a failure: non-zero exit,
    a missing 'Final Iter.' line, a V-cycle count more than 1 off the reference's golden run, a missing plotfile, or a
    difference found by the reference's fcompare.  (File name: runs after every other test module.)"""
    g = json.load(open(os.path.join(GOLDEN, f"solve_{golden}.json")))
    run = _run_driver(tmp_path, max_level=max_level, n_cell=64, max_grid_size=32, prob_type=prob_type, verbose=2, composite_solve=composite,
                      linop_maxorder=maxorder)
    log = run.stdout + run.stderr
    assert run.returncode == 0, "driver exited with %d:\n%s" % (run.returncode, log[-2000:])
    finals = [ln for ln in log.splitlines() if ln.startswith("MLMG: Final Iter.")]
    assert finals, "no 'MLMG: Final Iter.' line:\n" + log[-2000:]
    iters = int(finals[-1].split()[3])
    assert abs(iters - g["iters"]) <= 1, f"{iters} V-cycles (reference {g['iters']}):\n" + log[-2000:]
    assert os.path.isfile(os.path.join(str(tmp_path), "plot", "Header")), "no plotfile written:\n" + log[-2000:]
    # the plotfile the driver wrote, judged by the reference's fcompare against the reference's own run of the same deck
    fcompare = os.path.join(os.path.dirname(REF_DRIVER), "fcompare")
    if have_ref() and os.access(fcompare, os.X_OK):
        ref_plt = str(tmp_path / "plt_ref")
        run_ref(mode="solve", prob_type=prob_type, n_cell=64, max_grid_size=32, linop_maxorder=maxorder, agg_grid_size=32,
                max_level=max_level, composite_solve=composite, plotfile=ref_plt)
        cmp_ = subprocess.run([fcompare, "-r", "1e-7", ref_plt, os.path.join(str(tmp_path), "plot")], capture_output=True, text=True, timeout=300)
        assert cmp_.returncode == 0, "fcompare found differences:\n" + cmp_.stdout[-2000:]
