"""CPU tests of the oracle's constitutive restatement: analytic Jacobian vs finite differences,
tangent consistency, elastic limit, hardening saturation."""
import numpy as np
import pytest

import refcases


def _hist(orc, xtal, kin, props, seed=0):
    rng = np.random.default_rng(seed)
    h = orc.hist_init(xtal, kin, props)
    q = rng.normal(size=4)
    h[9:13] = q / np.linalg.norm(q)
    h[4:9] = 2e-4 * rng.normal(size=5)
    return h


@pytest.mark.parametrize("xtal,kin,pk", [(0, 0, "props_cp_voce"), (1, 0, "props_cp_voce"), (0, 2, "props_cp_mts"),
                                         (1, 2, "props_cp_mts"), (2, 2, "hcp")])
def test_local_jacobian_matches_finite_differences(orc, xtal, kin, pk):
    props = refcases.hcp_props() if pk == "hcp" else refcases.goldens()[pk]
    h = _hist(orc, xtal, kin, props, seed=3)
    rng = np.random.default_rng(4)
    d = np.zeros(7)
    d[:6] = 1e-3 * rng.normal(size=6)
    d[:3] -= d[:3].mean()
    w = 1e-3 * rng.normal(size=3)
    x0 = 0.3 * rng.normal(size=8)
    tK = 300.0
    R0, J = orc.local_problem(xtal, kin, props, 0.1, d, w, 1.0002, h, tK, x0)
    Jfd = np.zeros((8, 8))
    for j in range(8):
        eps = 1e-6
        xp, xm = x0.copy(), x0.copy()
        xp[j] += eps
        xm[j] -= eps
        Rp, _ = orc.local_problem(xtal, kin, props, 0.1, d, w, 1.0002, h, tK, xp)
        Rm, _ = orc.local_problem(xtal, kin, props, 0.1, d, w, 1.0002, h, tK, xm)
        Jfd[:, j] = (Rp - Rm) / (2 * eps)
    assert np.abs(J - Jfd).max() / np.abs(Jfd).max() < 1e-6


def _one_point(orc, xtal, kin, props, L, dt, stress0, hist0):
    """model_setup on a single unit-cube element with homogeneous velocity gradient L."""
    e2n, coords = orc.voxel_mesh(1, 1, 1)
    G, W = orc.hex8_dshape()
    jac = orc.jacobians(G, orc.gather(e2n, coords))
    X = coords.reshape(3, 8)
    vel = (L @ X).ravel()
    s1, h1, dd, nfail = orc.model_setup(xtal, kin, props, dt, 298.0, jac, G, orc.gather(e2n, vel),
                                        np.tile(stress0, 8), np.tile(hist0, 8))
    assert nfail == 0
    return s1[:6], h1[:hist0.size], dd[:36].reshape(6, 6).T  # K[i,j]


def test_elastic_limit_is_cubic_hooke(orc):
    props = refcases.goldens()["props_cp_voce"]
    h = orc.hist_init(0, 0, props)  # identity orientation
    L = np.diag([1e-6, -2e-6, 0.5e-6]) + 1e-6 * np.array([[0, 1, 0], [1, 0, 2], [0, 2, 0.0]])
    s, h1, K = _one_point(orc, 0, 0, props, L, 1.0, np.zeros(6), h)
    c11, c12, c44 = 168.4, 121.4, 75.2
    eps = 0.5 * (L + L.T)
    tr = np.trace(eps)
    dev = eps - tr / 3 * np.eye(3)
    # deviatoric part via (c11-c12) / 2c44, pressure via the bulk modulus
    sig = np.zeros((3, 3))
    for i in range(3):
        sig[i, i] = (c11 - c12) * dev[i, i]
    sig[1, 2] = sig[2, 1] = 2 * c44 * eps[1, 2]
    sig[0, 2] = sig[2, 0] = 2 * c44 * eps[0, 2]
    sig[0, 1] = sig[1, 0] = 2 * c44 * eps[0, 1]
    sig += (c11 + 2 * c12) / 3 * tr * np.eye(3)
    ref = np.array([sig[0, 0], sig[1, 1], sig[2, 2], sig[1, 2], sig[0, 2], sig[0, 1]])
    assert np.abs(s - ref).max() / np.abs(ref).max() < 1e-5
    # tangent = cubic stiffness (engineering shear columns)
    Kref = np.zeros((6, 6))
    Kref[:3, :3] = c12
    Kref[np.arange(3), np.arange(3)] = c11
    Kref[np.arange(3, 6), np.arange(3, 6)] = c44
    assert np.abs(K - Kref).max() / c11 < 1e-4


@pytest.mark.parametrize("xtal,kin,pk", [(0, 0, "props_cp_voce"), (1, 2, "props_cp_mts")])
def test_tangent_is_consistent(orc, xtal, kin, pk):
    """d sigma / d (D dt) from the implicit-function tangent matches finite differences of the update."""
    props = refcases.goldens()[pk]
    h = _hist(orc, xtal, kin, props, seed=5)
    h[4:9] = 0.0
    L0 = 2e-3 * np.array([[-0.4, 0.1, 0.0], [0.1, -0.3, 0.05], [0.0, 0.05, 1.0]])
    dt = 0.2
    # advance into the plastic regime first
    s, hh = np.zeros(6), h
    for _ in range(3):
        s, hh, K = _one_point(orc, xtal, kin, props, L0, dt, s, hh)
    s1, _, K = _one_point(orc, xtal, kin, props, L0, dt, s, hh)
    vmap = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]
    Kfd = np.zeros((6, 6))
    for j, (a, b) in enumerate(vmap):
        de = 1e-7
        dL = np.zeros((3, 3))
        if a == b:
            dL[a, a] = de / dt
        else:
            dL[a, b] = dL[b, a] = 0.5 * de / dt  # engineering shear increment de
        sp, _, _ = _one_point(orc, xtal, kin, props, L0 + dL, dt, s, hh)
        sm, _, _ = _one_point(orc, xtal, kin, props, L0 - dL, dt, s, hh)
        Kfd[:, j] = (sp - sm) / (2 * de)
    # the analytic tangent ignores the geometric change of the element (J) and of V_new's
    # dependence entering through the reference-side kernel_setup only at O(strain)
    assert np.abs(K - Kfd).max() / np.abs(Kfd).max() < 2e-3


def test_voce_hardening_saturates(orc):
    props = refcases.goldens()["props_cp_voce"]
    h = orc.hist_init(0, 0, props)
    L = 1e-2 * np.diag([-0.5, -0.5, 1.0])
    s, hh = np.zeros(6), h
    for _ in range(60):
        s, hh, _ = _one_point(orc, 0, 0, props, L, 1.0, s, hh)
    assert 17e-3 < hh[13] <= 122.4e-3 + 1e-12
    assert hh[13] > 0.12  # close to saturation after 60% strain
    assert abs(np.linalg.norm(hh[9:13]) - 1.0) < 1e-12


def test_hcp_slip_systems_and_update(orc):
    """HCP KMBalD (24 systems, per-family resistances; synthetic properties): the update converges, keeps the
    quaternion normalised, yields a transversely-isotropic elastic tangent and plastic flow at large strain."""
    props = refcases.hcp_props()
    h = orc.hist_init(2, 2, props)
    assert h.size == 40
    L = 2e-3 * np.diag([-0.5, -0.5, 1.0])
    s, hh = np.zeros(6), h
    for _ in range(6):
        s, hh, K = _one_point(orc, 2, 2, props, L, 0.5, s, hh)
    assert abs(np.linalg.norm(hh[9:13]) - 1.0) < 1e-12
    assert np.abs(hh[14:38]).sum() > 1e-4          # slip is active
    assert s[2] > 0 and abs(s[0] - s[1]) < 1e-9 * abs(s[2])  # c-axis loading keeps the basal plane isotropic
    s_el, _, K_el = _one_point(orc, 2, 2, props, 1e-6 * np.diag([1.0, -0.3, 0.2]), 1.0, np.zeros(6), h)
    c11, c12, c13, c33, c44 = props[3:8]
    Kref = np.array([[c11, c12, c13], [c12, c11, c13], [c13, c13, c33]])
    assert np.abs(K_el[:3, :3] - Kref).max() / c11 < 1e-4
    assert abs(K_el[3, 3] - c44) / c44 < 1e-4 and abs(K_el[5, 5] - 0.5 * (c11 - c12)) / c11 < 1e-4
