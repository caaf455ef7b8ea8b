"""CPU tests for the soil-spectrum file path: the reference's -soil_spectra is a stub that builds
the 1-nm table, prints it with "%d %lf" and exits (gortt.c:1388-1451).  tests/golden/soil_cases.json holds what the
reference BINARY prints for a set of files (made by tests/golden/make_golden.py).  Checked here, without a GPU:
  * the restatement of the interpolation loop (oracle) reproduces those bytes and error cases;
  * the product's host-side reader gort_soil_table_read (no CUDA involved) gives the restatement's bits and the
    reference's messages."""
import json
from pathlib import Path

import numpy as np
import pytest

import gort_b200

CASES = json.loads((Path(__file__).resolve().parent / "golden" / "soil_cases.json").read_text())
OK = [c for c in CASES if c["stdout"]]
ERR = [c for c in CASES if not c["stdout"]]


def as_text(table):
    return "".join("%d %f\n" % (i + 400, v) for i, v in enumerate(table))


@pytest.mark.parametrize("case", OK, ids=lambda c: c["name"])
def test_table_matches_reference_binary(case, oracle, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    Path("soil.txt").write_text(case["file"])
    rc, tab, _ = oracle.soil_table("soil.txt")
    assert rc == 0 and as_text(tab) == case["stdout"]
    got = gort_b200.soil_table_read("soil.txt")
    assert np.array_equal(got, tab)                       # same loop, same bits


@pytest.mark.parametrize("case", ERR, ids=lambda c: c["name"])
def test_error_messages_match_reference_binary(case, oracle, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    name = "soil.txt"
    if case["file"] is None:
        name = "missing.txt"
    else:
        Path(name).write_text(case["file"])
    rc, _, _ = oracle.soil_table(name)
    assert rc != 0
    with pytest.raises(gort_b200.GortError) as ei:
        gort_b200.soil_table_read(name)
    assert ei.value.code == 5
    assert "gortt: %s\n" % str(ei.value).split(": ", 1)[1] == case["stderr"]


def test_lookup_restatement_on_grid_and_between(oracle, tmp_path):
    p = tmp_path / "s.txt"
    p.write_text(next(c for c in OK if c["name"] == "irregular")["file"])
    _, tab, _ = oracle.soil_table(p)
    wl = np.array([400.0, 401.0, 1234.0, 2500.0, 400.5, 858.5, 2499.75])
    r = oracle.soil_lookup(tab, wl)
    assert np.array_equal(r[:4], tab[[0, 1, 834, 2100]])
    assert r[4] == 0.5 * tab[0] + 0.5 * tab[1] and r[5] == 0.5 * tab[458] + 0.5 * tab[459]
    assert abs(r[6] - (0.25 * tab[2099] + 0.75 * tab[2100])) < 1e-16
