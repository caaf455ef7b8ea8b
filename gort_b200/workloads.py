"""Synthetic workloads C1-C5 of BASELINE.json / SURVEY.md 8(d), generated with
numpy.random.Generator(PCG64(seed)).  Shared by tests/ and bench.py so that the oracle and the GPU
path see byte-identical inputs.  Pure input generation: no model arithmetic.

Layouts follow include/gort_b200.h: structure [6][M] (lambda, r, b, h1, h2, favd), leaf [7][M],
soil [4][M], angles [4][G] or [4][M][G] in degrees (vza, vaa, sza, saa).
"""
import numpy as np

from .api import structure_from_options

DEFAULT_LEAF = np.array([1.2, 30.0, 10.0, 1.0, 0.0, 0.015, 0.009])        # gortt.c:53-59
DEFAULT_SOIL = np.array([0.2, 0.1, 0.03726, -0.002426])                   # gortt.c:38-41
MODIS_BANDS = np.array([645.0, 858.5, 469.0, 555.0, 1240.0, 1640.0, 2130.0])


def c1_readme():
    """README example: gortt -LAI 4.0, one geometry (vza 10, sza 30, raz 20), 450/600/800/1000 nm."""
    st = structure_from_options(lai=4.0).reshape(6, 1)
    ang = np.array([[10.0], [0.0], [30.0], [20.0]])
    wl = np.array([450.0, 600.0, 800.0, 1000.0])
    return dict(structure=st, angles=ang, wavelength=wl, leaf=DEFAULT_LEAF.reshape(7, 1).copy(),
                soil=DEFAULT_SOIL.reshape(4, 1).copy())


def c2_hemisphere(wl_step=1, sets=1, lai0=4.0):
    """Hemispherical BRDF sweep: vza, sza in {0,5,..,85}, view azimuth in {0,10,..,350}, sun azimuth 0;
    400-2500 nm at `wl_step` nm; structure defaults + -LAI 4 (set k > 0 uses LAI lai0 + 0.25 k: the
    weak-scaling variant gives every extra GPU its own forest)."""
    vz = np.arange(0.0, 90.0, 5.0)
    sz = np.arange(0.0, 90.0, 5.0)
    az = np.arange(0.0, 360.0, 10.0)
    V, S, A = np.meshgrid(vz, sz, az, indexing="ij")     # azimuth fastest: 36 consecutive lines share the sun
    ang = np.stack([V.ravel(), A.ravel(), S.ravel(), np.zeros(V.size)])
    wl = np.arange(400.0, 2500.0 + 0.5, float(wl_step))
    st = np.stack([structure_from_options(lai=lai0 + 0.25 * k) for k in range(sets)], axis=1)
    return dict(structure=st, angles=ang, wavelength=wl,
                leaf=np.repeat(DEFAULT_LEAF.reshape(7, 1), sets, axis=1),
                soil=np.repeat(DEFAULT_SOIL.reshape(4, 1), sets, axis=1))


def random_structures(rng, n):
    """C3 ranges: r~U[0.3,3], b/r~U[0.5,4], h1=b+U[0,5], h2=h1+U[0.5,15], cover=lambda pi r^2~U[0.05,0.9],
    LAI~U[0.5,8] (-> favd by gortt.c:1129)."""
    r = rng.uniform(0.3, 3.0, n)
    b = r * rng.uniform(0.5, 4.0, n)
    h1 = b + rng.uniform(0.0, 5.0, n)
    h2 = h1 + rng.uniform(0.5, 15.0, n)
    cover = rng.uniform(0.05, 0.9, n)
    lam = cover / (np.pi * r * r)
    lai = rng.uniform(0.5, 8.0, n)
    favd = lai * 3.0 / (lam * r * r * np.pi * b * 4.0)
    return np.stack([lam, r, b, h1, h2, favd])


def random_leaves(rng, n):
    """PROSPECT N~U[1,3], Cab~U[5,80], Car=Cab/4.5, Anth~U[0,5], Cbrown~U[0,0.5], Cw~U[0.004,0.04],
    Cm~U[0.002,0.016]."""
    N = rng.uniform(1.0, 3.0, n)
    cab = rng.uniform(5.0, 80.0, n)
    car = cab / 4.5
    anth = rng.uniform(0.0, 5.0, n)
    cbrown = rng.uniform(0.0, 0.5, n)
    cw = rng.uniform(0.004, 0.04, n)
    cm = rng.uniform(0.002, 0.016, n)
    return np.stack([N, cab, car, anth, cbrown, cw, cm])


def c3_albedo(n_sets=10000, seed=1001, wl_step=10):
    """Spectral albedo + fAPAR for n_sets canopy parameter sets with PROSPECT-D leaf optics;
    sza in {0,30,60}; 400-2500 nm step 10 (211 bands)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    st = random_structures(rng, n_sets)
    leaf = random_leaves(rng, n_sets)
    soil = np.repeat(DEFAULT_SOIL.reshape(4, 1), n_sets, axis=1)
    ang = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 30.0, 60.0], [0.0, 0.0, 0.0]])
    wl = np.arange(400.0, 2500.0 + 0.5, float(wl_step))
    return dict(structure=st, angles=ang, wavelength=wl, leaf=leaf, soil=soil)


def c4_enkf(n_members=100000, seed=1002, vary_structure=True, n_geom=16):
    """EnKF-style forward operator: members x 16 MODIS-like geometries x 7 bands.  4a: structure shared,
    LAI-only varying (vary_structure=False); 4b: everything varying."""
    rng = np.random.Generator(np.random.PCG64(seed))
    st = random_structures(rng, n_members)
    leaf = random_leaves(rng, n_members)
    if not vary_structure:
        base = st[:, :1].copy()
        lam, r, b = base[0, 0], base[1, 0], base[2, 0]
        lai = rng.uniform(0.5, 8.0, n_members)
        st = np.repeat(base, n_members, axis=1)
        st[5] = lai * 3.0 / (lam * r * r * np.pi * b * 4.0)
    soil = np.repeat(DEFAULT_SOIL.reshape(4, 1), n_members, axis=1)
    vza = rng.uniform(0.0, 60.0, (n_members, n_geom))
    vaa = rng.uniform(0.0, 360.0, (n_members, n_geom))
    sza = rng.uniform(20.0, 70.0, (n_members, n_geom))
    saa = rng.uniform(0.0, 360.0, (n_members, n_geom))
    ang = np.stack([vza, vaa, sza, saa])
    return dict(structure=st, angles=ang, wavelength=MODIS_BANDS.copy(), leaf=leaf, soil=soil)


def c5_lut_grid(n=(8, 8, 4, 8, 8, 8), seed=1003):
    """Tensor grid r x b/r x h1 x (h2-h1) x cover x favd over the C3 ranges (default 131 072 LUTs)."""
    r = np.linspace(0.3, 3.0, n[0])
    br = np.linspace(0.5, 4.0, n[1])
    dh1 = np.linspace(0.0, 5.0, n[2])
    dh = np.linspace(0.5, 15.0, n[3])
    cover = np.linspace(0.05, 0.9, n[4])
    favd = np.linspace(0.1, 3.0, n[5])
    R, BR, DH1, DH, CV, FV = [a.ravel() for a in np.meshgrid(r, br, dh1, dh, cover, favd, indexing="ij")]
    b = R * BR
    h1 = b + DH1
    h2 = h1 + DH
    lam = CV / (np.pi * R * R)
    return dict(structure=np.stack([lam, R, b, h1, h2, FV]))
