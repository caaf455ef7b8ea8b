"""GPU tests for the finished -soil_spectra path: the table lookup kernel against the restatement, and the gortt
command line (-soil_spectra runs the model with the file's soil spectrum; -soil_dump reproduces the reference stub's
output byte for byte)."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

import gort_b200
from gort_b200 import workloads as wk

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
CLI = ROOT / "gort_b200" / "bin" / "gortt"
CASES = json.loads((ROOT / "tests" / "golden" / "soil_cases.json").read_text())


def test_lookup_kernel_matches_restatement(gort, oracle, tmp_path):
    rng = np.random.Generator(np.random.PCG64(9))
    for name in ("irregular", "grid5", "sparse3"):
        p = tmp_path / (name + ".txt")
        p.write_text(next(c for c in CASES if c["name"] == name)["file"])
        tab = gort_b200.soil_table_read(p)
        wl = np.concatenate([[400.0, 2500.0, 858.5, 2499.999], np.arange(400.0, 2501.0, 7.0), rng.uniform(400, 2500, 200)])
        got = gort.soil_from_table(tab, wl, n_sets=3)
        want = oracle.soil_lookup(tab, wl)
        assert got.shape == (3, wl.size)
        for m in range(3):
            assert np.array_equal(got[m], want)            # products and sums of two table rows: exact in both
    with pytest.raises(gort_b200.GortError) as ei:
        gort.soil_from_table(tab, np.array([399.0]))
    assert ei.value.code == 3


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_cli_soil_dump_reproduces_the_reference_stub(case, tmp_path):
    name = "soil.txt" if case["file"] is not None else "missing.txt"
    if case["file"] is not None:
        (tmp_path / name).write_text(case["file"])
    r = subprocess.run(["gortt", "-soil_dump", name], executable=str(CLI), input="", capture_output=True, text=True, cwd=tmp_path)
    assert (r.returncode, r.stdout, r.stderr) == (case["rc"], case["stdout"], case["stderr"])
    if case["file"] is None or not case["stdout"]:
        # the finished option fails on the same files with the same messages
        r = subprocess.run(["gortt", "-soil_spectra", name, "-LAI", "4.0"], executable=str(CLI), input="1 2 450 800\n10 0 30 20\n",
                           capture_output=True, text=True, cwd=tmp_path)
        assert (r.returncode, r.stdout, r.stderr) == (case["rc"], "", case["stderr"])


def test_cli_soil_spectra_runs_the_model_with_the_file(gort, oracle, tmp_path):
    (tmp_path / "soil.txt").write_text(next(c for c in CASES if c["name"] == "irregular")["file"])
    wl = np.array([450.0, 600.5, 858.5, 1640.0, 2500.0])
    stdin = "2 5 450 600.5 858.5 1640 2500\n10 0 30 20\n-25 40 55 300\n"
    r = subprocess.run(["gortt", "-LAI", "4.0", "-soil_spectra", "soil.txt", "-energy"], executable=str(CLI), input=stdin,
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and r.stderr == ""
    got = np.array([[float(x) for x in ln.split()] for ln in r.stdout.strip().split("\n")[1:]])
    st = gort_b200.structure_from_options(lai=4.0)
    _, tab, _ = oracle.soil_table(tmp_path / "soil.txt")
    rl, tl, _ = oracle.spectra(wk.DEFAULT_LEAF, wk.DEFAULT_SOIL, wl)
    rs = oracle.soil_lookup(tab, wl)
    ang = np.array([[10.0, 0.0, 30.0, 20.0], [-25.0, 40.0, 55.0, 300.0]])
    lut = oracle.lut(st)
    rsurf, _, _ = oracle.brdf(st, lut, ang, rl, tl, rs)
    alb, fv, fs = oracle.energy(st, lut, ang, rl, tl, rs)
    # line layout: angles, rsurf per band, then (albedo, favegt, fasoil) per band (gortt.c:309-325)
    want = np.concatenate([ang, rsurf, np.stack([alb, fv, fs], axis=2).reshape(2, -1)], axis=1)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 1.000001e-6          # "%f" text: 6 decimals
    # and it differs from the Price soil run: the option is live
    r2 = subprocess.run(["gortt", "-LAI", "4.0", "-energy"], executable=str(CLI), input=stdin, capture_output=True, text=True)
    assert r2.stdout != r.stdout
