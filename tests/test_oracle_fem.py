"""CPU tests of the oracle's FEM side against the reference's own analytic identities
(test/mechanics_test.cpp:780-814: PA action == EA action == element matrix action to 1e-14;
test/grad_test.cpp:86-103,192-196: analytic gradient of a linear field)."""
import numpy as np
import pytest


def _mesh(orc, n=2, distort=0.0, seed=0):
    e2n, coords = orc.voxel_mesh(n, n, n)
    if distort:
        rng = np.random.default_rng(seed)
        coords = coords + distort / n * (rng.random(coords.size) - 0.5)
    G, W = orc.hex8_dshape()
    jac = orc.jacobians(G, orc.gather(e2n, coords))
    return e2n, coords, G, W, jac


def _cubic_tangent(npts):
    # test/mechanics_test.cpp:749-778: cubic material {100, 75, 50}
    K = np.zeros((6, 6))
    K[:3, :3] = 75.0
    K[np.arange(3), np.arange(3)] = 100.0
    K[np.arange(3, 6), np.arange(3, 6)] = 50.0
    return np.tile(K.T.ravel(), npts)  # column-major storage k[j*6+i]


def test_dshape_partition_of_unity(orc):
    G, W = orc.hex8_dshape()
    G = G.reshape(8, 3, 8)
    assert np.allclose(G.sum(axis=2), 0.0, atol=1e-15)
    assert np.isclose(W.sum(), 1.0)


def test_grad_calc_linear_field(orc):
    # u = (2x+3y+4z, 4x+2y+3z, 3x+4y+2z): gradient [[2,3,4],[4,2,3],[3,4,2]] (grad_test.cpp:86-103)
    e2n, coords, G, W, jac = _mesh(orc, 2, distort=0.2)
    nn = coords.size // 3
    X = coords.reshape(3, nn)
    A = np.array([[2.0, 3, 4], [4, 2, 3], [3, 4, 2]])
    u = (A @ X).ravel()
    g = orc.grad_calc(jac, G, orc.gather(e2n, u)).reshape(-1, 3, 3)  # [pt][t][i]
    err = np.linalg.norm(g - A.T[None]) / g.shape[0]
    assert err < 3e-15


@pytest.mark.parametrize("tangent", ["ones", "cubic", "random"])
def test_pa_equals_ea_action(orc, tangent):
    e2n, coords, G, W, jac = _mesh(orc, 2, distort=0.2)
    ne = e2n.size // 8
    npts = ne * 8
    rng = np.random.default_rng(1)
    if tangent == "ones":
        k36 = np.ones(npts * 36)
    elif tangent == "cubic":
        k36 = _cubic_tangent(npts)
    else:
        S = rng.normal(size=(npts, 6, 6))
        k36 = (S + S.transpose(0, 2, 1)).ravel()  # symmetric => PA (K) and EA (K^T) agree
    x = np.arange(1, ne * 24 + 1, dtype=float)  # x_i = i + 1 (mechanics_test.cpp:111-113)
    dt = 1.0
    y_pa = orc.grad_mult_pa(dt, jac, W, G, k36, x)
    ea = orc.assemble_ea(dt, jac, W, G, k36)
    y_ea = orc.ea_mult(ea, x)
    scale = np.abs(y_pa).max()
    assert np.abs(y_pa - y_ea).max() / scale < 1e-13
    # element matrices are symmetric for symmetric K and their diagonal equals the PA diagonal
    E = ea.reshape(ne, 24, 24)
    assert np.abs(E - E.transpose(0, 2, 1)).max() / np.abs(E).max() < 1e-13
    d_pa = orc.grad_diag_pa(dt, jac, W, G, k36)
    assert np.abs(d_pa - orc.ea_diag(ea)).max() / np.abs(d_pa).max() < 1e-13


def test_residual_unit_stress(orc):
    # sigma == 1 in every component (ExaNLFIntegratorPAVecTest, mechanics_test.cpp:184-303):
    # the residual of a constant stress field sums to zero over the nodes of each element
    e2n, coords, G, W, jac = _mesh(orc, 2, distort=0.2)
    ne = e2n.size // 8
    y = orc.residual_pa(jac, W, G, np.ones(ne * 48)).reshape(ne, 3, 8)
    assert np.abs(y.sum(axis=2)).max() < 1e-14
    # and assembled over a closed mesh it vanishes at interior nodes
    nn = coords.size // 3
    yL = orc.scatter_add(e2n, y.ravel(), nn).reshape(3, nn)
    interior = 13  # centre node of the 3x3x3 node grid
    assert np.abs(yL[:, interior]).max() < 1e-14


def test_bbar_reduces_to_standard_for_deviatoric(orc):
    # with a trace-free stress the B-bar residual equals the standard one; the B-bar element
    # matrix equals the standard one for a purely deviatoric tangent
    e2n, coords, G, W, jac = _mesh(orc, 2, distort=0.2)
    ne = e2n.size // 8
    rng = np.random.default_rng(2)
    s = rng.normal(size=(ne * 8, 6))
    s[:, :3] -= s[:, :3].mean(axis=1, keepdims=True)
    eds = orc.ic_eds(jac, W, G)
    y0 = orc.residual_pa(jac, W, G, s.ravel())
    y1 = orc.ic_residual_pa(jac, W, G, eds, s.ravel())
    assert np.abs(y0 - y1).max() / np.abs(y0).max() < 1e-13
    # B-bar diagonal == diagonal of the B-bar element matrices
    S = rng.normal(size=(ne * 8, 6, 6))
    k36 = (S + S.transpose(0, 2, 1)).ravel()
    ea = orc.ic_assemble_ea(0.5, jac, W, G, eds, k36)
    d = orc.ic_grad_diag_pa(0.5, jac, W, G, eds, k36)
    assert np.abs(orc.ea_diag(ea) - d).max() / np.abs(d).max() < 1e-13


def test_vol_sum(orc):
    e2n, coords, G, W, jac = _mesh(orc, 3, distort=0.0)
    ne = e2n.size // 8
    sums, vol = orc.vol_sum(jac, W, np.ones(ne * 8 * 2), 2)
    assert np.isclose(vol, 1.0, rtol=1e-14)
    assert np.allclose(sums, 1.0, rtol=1e-14)
