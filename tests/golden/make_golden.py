#!/usr/bin/env python3
"""Generate the committed golden fixtures from the UNMODIFIED reference compiled into oracle/_ref.

Run in the build container only (needs /root/reference to have been compiled by oracle/Makefile):
    python tests/golden/make_golden.py
Outputs (committed, small):
    tests/golden/cli_cases.json     command lines + stdin -> exact stdout / stderr / exit code of the
                                    reference binary oracle/_ref/gortt_ref
    tests/golden/soil_cases.json    soil-spectrum files -> the 1-nm table the reference binary prints for them
    tests/golden/ref_vectors.npz    in-memory doubles from oracle/_ref/libgortt_ref.so (LUTs, BRDF, energy,
                                    spectra) on seeded inputs
The reference ships no golden vectors of its own; these pin the oracle restatement and
the CUDA path on boxes where /root/reference does not exist.  PROSPECT-D inside the reference binary is
our restatement (no Fortran compiler): cases that use it pin the C interface around it, not the
Fortran arithmetic.
"""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from checkers import ref, REF_BIN  # noqa: E402
from gort_b200 import workloads as wk  # noqa: E402

README_IN = "1 4 450 600 800 1000\n10 0 30 20\n"
MODIS_IN = "3 7 469 555 645 858.5 1240 1640 2130\n0 0 30 0\n30 0 45 180\n-30 0 45 0\n"
SWEEP_IN = "6 3 450 650 850\n0 0 0 0\n20 90 40 10\n60 200 10 300\n-15 10 -35 20\n85 0 60 0\n40 400 50 -400\n"


def cli_cases():
    lutfile = HERE / "lut_lai4.txt"
    cases = [
        dict(name="readme_prospect", args=["-LAI", "4.0"], stdin=README_IN),
        dict(name="e1_alb_leaf", args=["-LAI", "4.0", "-alb_leaf", "0.5"], stdin=README_IN),
        dict(name="e2_q08", args=["-LAI", "4.0", "-alb_leaf", "0.5", "-q08_pn_kopen"], stdin=README_IN),
        dict(name="e3_new_style", args=["-alb_leaf", "0.5", "-alb_soil", "0.2", "-HB", "2", "-BR", "1.5", "-PCC", "0.6", "-LAI", "3.7"],
             stdin="3 2 650 850\n0 0 45 0\n30 0 45 180\n-30 0 45 0\n"),
        dict(name="e5_modis_prnprop_energy", args=["-LAI", "4.0", "-alb_leaf", "0.5", "-prnprop", "-energy"], stdin=MODIS_IN),
        dict(name="prnspec_prnprop", args=["-LAI", "2.5", "-prnspec", "-prnprop"], stdin=SWEEP_IN),
        dict(name="old_style_structure", args=["-lambda", "0.2", "-r", "1.1", "-b", "2.0", "-h1", "4", "-h2", "11", "-favd", "0.7"], stdin=SWEEP_IN),
        dict(name="beta_diffuse", args=["-LAI", "3", "-beta", "0.4", "-diffuse", "0.3"], stdin=SWEEP_IN),
        dict(name="prospect_price_options", args=["-LAI", "3", "-N", "1.8", "-cab", "45", "-car", "9", "-canth", "2", "-cbrown", "0.1",
                                                  "-cw", "0.02", "-cm", "0.006", "-rsl1", "0.3", "-rsl2", "-0.05", "-rsl3", "0.01", "-rsl4", "0.002"],
             stdin=MODIS_IN),
        dict(name="write_lut", args=["-LAI", "4.0", "-W"], stdin=""),
        dict(name="write_lut_q08", args=["-LAI", "4.0", "-W", "-q08_pn_kopen"], stdin=""),
        dict(name="read_lut", args=["-LAI", "4.0", "-alb_leaf", "0.5", "-P", "@LUT@"], stdin=SWEEP_IN),
        dict(name="energy_alb_soil", args=["-LAI", "1.5", "-alb_soil", "0.15", "-energy"], stdin="2 3 500 700 900\n0 0 20 0\n10 0 55 100\n"),
        # error paths
        dict(name="err_unknown_option", args=["-zzz"], stdin=""),
        dict(name="err_unknown_argument", args=["-LAI", "4", "foo"], stdin=""),
        dict(name="err_wavelength_count", args=["-LAI", "4"], stdin="1 3 450 600\n10 0 30 20\n"),
        dict(name="err_wavelength_range", args=["-LAI", "4"], stdin="1 2 399 600\n10 0 30 20\n"),
        dict(name="err_bad_line", args=["-LAI", "4", "-alb_leaf", "0.5"], stdin="3 2 650 850\n0 0 45 0\n30 0 oops 180\n-30 0 45 0\n"),
        dict(name="err_angle_count", args=["-LAI", "4", "-alb_leaf", "0.5"], stdin="3 2 650 850\n0 0 45 0\n30 0 45 180\n"),
        dict(name="err_no_stdin", args=["-LAI", "4"], stdin=""),
        dict(name="err_missing_lut", args=["-P", "/nonexistent/lut.txt"], stdin=README_IN),
    ]
    # the LUT text the read_lut case consumes: the reference's own -W output
    w = subprocess.run([str(REF_BIN), "-LAI", "4.0", "-W"], input="", capture_output=True, text=True)
    lutfile.write_text(w.stdout)
    out = []
    for c in cases:
        args = [a.replace("@LUT@", str(lutfile)) for a in c["args"]]
        r = subprocess.run(["gortt"] + args, executable=str(REF_BIN), input=c["stdin"], capture_output=True, text=True)
        out.append(dict(name=c["name"], args=c["args"], stdin=c["stdin"], stdout=r.stdout, stderr=r.stderr, rc=r.returncode))
        print("%-28s rc=%d stdout=%dB stderr=%dB" % (c["name"], r.returncode, len(r.stdout), len(r.stderr)))
    (HERE / "cli_cases.json").write_text(json.dumps(out, indent=1))


def soil_files():
    """name -> text of the soil-spectrum files of the soil cases (deterministic)."""
    rng = np.random.Generator(np.random.PCG64(77))
    f = {}
    wl = np.arange(350.0, 2600.1, 5.0)
    f["grid5"] = "".join("%g %.6f\n" % (w, 0.08 + 0.25 * (1 - np.exp(-(w - 350) / 600.0)) + 0.02 * np.sin(w / 90.0)) for w in wl)
    w = np.sort(np.concatenate([[399.3, 2500.7], rng.uniform(400, 2500, 300)]))
    f["irregular"] = "".join("%.4f %.7f\n" % (a, b) for a, b in zip(w, 0.05 + 0.3 * rng.uniform(0, 1, w.size)))
    f["exact1nm"] = "".join("%d %.5f\n" % (a, 0.1 + 0.0001 * (a - 400)) for a in range(400, 2501))
    f["sparse3"] = "400 0.1\n1000 0.35\n2500 0.2\n"
    f["wide_margins"] = "100 0.5\n399.5 0.1\n400.5 0.2\n2499.5 0.3\n3000 0.9\n"
    f["err_first"] = "401 0.1\n2500 0.2\n"
    f["err_last"] = "400 0.1\n2499 0.2\n"
    f["err_line"] = "400 0.1\n1000 oops\n2500 0.2\n"
    return f


def soil_cases():
    """The reference's unfinished -soil_spectra prints the interpolated 1-nm table and exits (gortt.c:1441-1442):
    stdout / stderr / exit code of the reference binary for each soil file, run with the file in the working directory
    (the messages quote the path as given)."""
    import tempfile
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, text in soil_files().items():
            (Path(tmp) / "soil.txt").write_text(text)
            r = subprocess.run(["gortt", "-soil_spectra", "soil.txt"], executable=str(REF_BIN), input="", capture_output=True,
                               text=True, cwd=tmp)
            out.append(dict(name=name, file=text, stdout=r.stdout, stderr=r.stderr, rc=r.returncode))
            print("soil %-14s rc=%d stdout=%dB stderr=%r" % (name, r.returncode, len(r.stdout), r.stderr[:80]))
        r = subprocess.run(["gortt", "-soil_spectra", "missing.txt"], executable=str(REF_BIN), input="", capture_output=True, text=True, cwd=tmp)
        out.append(dict(name="err_missing", file=None, stdout=r.stdout, stderr=r.stderr, rc=r.returncode))
    (HERE / "soil_cases.json").write_text(json.dumps(out, indent=0))


def vectors():
    r = ref()
    rng = np.random.Generator(np.random.PCG64(2026))
    d = {}
    # structures: CLI defaults + LAI 4, new-style example, and random C3-range sets
    sts = [wk.structure_from_options(lai=4.0), wk.structure_from_options(hb=2, br=1.5, pcc=0.6, lai=3.7)]
    sts += list(wk.random_structures(rng, 6).T)
    st = np.array(sts)                                   # [8][6]
    d["structure"] = st
    d["lut_full"] = np.array([r.lut(s, 0) for s in st])
    d["lut_q08"] = np.array([r.lut(s, 1) for s in st])
    leaf = wk.random_leaves(rng, 8).T                    # [8][7]
    leaf[0] = wk.DEFAULT_LEAF
    soil = np.tile(wk.DEFAULT_SOIL, (8, 1))
    soil[3:] += rng.uniform(-0.02, 0.02, (5, 4))
    wl = np.array([400.0, 469.0, 555.0, 645.0, 858.5, 1240.0, 1640.0, 2130.0, 2500.0, 1234.56])
    d["leaf"], d["soil"], d["wavelength"] = leaf, soil, wl
    sp = [r.spectra(leaf[k], soil[k], wl) for k in range(8)]
    d["rleaf"] = np.array([s[0] for s in sp]); d["tleaf"] = np.array([s[1] for s in sp]); d["rsoil"] = np.array([s[2] for s in sp])
    ang = np.stack([rng.uniform(-80, 89, 24), rng.uniform(-400, 400, 24), rng.uniform(-80, 89, 24), rng.uniform(-400, 400, 24)], axis=1)
    ang[0] = [10, 0, 30, 20]; ang[1] = [0, 0, 0, 0]; ang[2] = [89.6, 10, 30, 0]; ang[3] = [20, 0, 89.4, 180]
    ang[4:8, 2] = 35.0; ang[4:8, 3] = 100.0
    d["angles"] = ang
    rs, sc, kp = [], [], []
    for k in range(8):
        a, b, c = r.brdf(st[k], d["lut_full"][k], ang, d["rleaf"][k], d["tleaf"][k], d["rsoil"][k])
        rs.append(a); sc.append(b); kp.append(c)
    d["rsurf"], d["scomp"], d["kprop"] = np.array(rs), np.array(sc), np.array(kp)
    a, _, _ = r.brdf(st[0], d["lut_full"][0], ang, d["rleaf"][0], d["tleaf"][0], d["rsoil"][0], beta=0.3, fd=0.8)
    d["rsurf_beta_fd"] = a
    en = [r.energy(st[k], d["lut_full"][k], ang[:3], d["rleaf"][k], d["tleaf"][k], d["rsoil"][k]) for k in range(3)]
    d["albedo"] = np.array([e[0] for e in en]); d["favegt"] = np.array([e[1] for e in en]); d["fasoil"] = np.array([e[2] for e in en])
    x, w = r.gauleg(32)
    d["gauleg_x"], d["gauleg_w"] = x, w
    np.savez_compressed(HERE / "ref_vectors.npz", **d)
    print("ref_vectors.npz:", {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    if ref() is None or not REF_BIN.exists():
        raise SystemExit("oracle/_ref is not built: run `make -C oracle` where /root/reference exists")
    only = sys.argv[1:]
    if not only or "cli" in only:
        cli_cases()
    if not only or "soil" in only:
        soil_cases()
    if not only or "vectors" in only:
        vectors()
