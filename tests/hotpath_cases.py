"""Shared builders for hot-path parity cases: the same seeded inputs are pushed through the CPU
oracle (oracle/) and through the CUDA library via its C ABI (exaconstit_b200.capi)."""
import numpy as np

import refcases


def make_case(n=4, seed=0, ngrains=4, xtal=0, kin=0, props=None, rate=5e-3, dt=0.1, distort=0.15, presteps=1,
              assembly=0, integ=0):
    """Voxel mesh n^3 with randomly displaced interior nodes, random grains/orientations, a velocity field
    = uniaxial-ish gradient + noise.  `presteps` oracle updates advance the state into the plastic regime
    so that the tested step starts from non-trivial stress/history."""
    from oracle import orc
    rng = np.random.default_rng(seed)
    g = refcases.goldens()
    if props is None:
        if xtal == 2:
            props = refcases.hcp_props()
        else:
            props = g["props_cp_mts"] if kin == 2 else (g["props_cp_vocenl"] if kin == 1 else g["props_cp_voce"])
    nx = ny = nz = n
    e2n, coords = orc.voxel_mesh(nx, ny, nz)
    ne, nn = nx * ny * nz, (nx + 1) ** 3
    h = 1.0 / n
    coords = coords + distort * h * (rng.random(coords.size) - 0.5)
    # velocity: v = L x + noise
    L = rate * np.array([[-0.35, 0.1, 0.0], [0.05, -0.3, 0.1], [0.0, -0.1, 1.0]])
    X = coords.reshape(3, nn)
    vel = (L @ X).ravel() + 0.05 * rate * h * (rng.random(3 * nn) - 0.5)
    grains = rng.integers(1, ngrains + 1, size=ne).astype(np.int32)
    quats = rng.normal(size=(ngrains, 4))
    quats /= np.linalg.norm(quats, axis=1)[:, None]
    nsv = orc.nhist(xtal, kin)
    hinit = orc.hist_init(xtal, kin, props)
    hist0 = np.tile(hinit, ne * 8).reshape(ne * 8, nsv)
    hist0[:, 9:13] = np.repeat(quats[grains - 1], 8, axis=0)
    hist0 = hist0.ravel().copy()
    stress0 = np.zeros(ne * 8 * 6)
    G, W = orc.hex8_dshape()
    case = dict(n=n, ne=ne, nn=nn, e2n=e2n, xbeg=coords.copy(), vel=vel, dt=dt, xtal=xtal, kin=kin, props=props,
                temp_k=298.0, nsv=nsv, G=G, W=W, assembly=assembly, integ=integ, seed=seed)
    x = coords.copy()
    for _ in range(presteps):
        xend = x + dt * vel
        jac = orc.jacobians(G, orc.gather(e2n, xend))
        s1, h1, _, nfail = orc.model_setup(xtal, kin, props, dt, 298.0, jac, G, orc.gather(e2n, vel), stress0, hist0)
        assert nfail == 0
        stress0, hist0, x = s1, h1, xend
    case["xbeg"] = x
    case["stress0"], case["hist0"] = stress0, hist0
    # essential mask: z-min face all comps, x-min face x comp
    mask = np.zeros(nn, dtype=np.uint8)
    idx = np.arange(nn)
    i, k = idx % (n + 1), idx // ((n + 1) ** 2)
    mask[k == 0] |= 7
    mask[i == 0] |= 1
    case["essmask"] = mask
    case["xvec"] = rng.normal(size=3 * nn)
    return case


def ess_dofs(case):
    m, nn = case["essmask"], case["nn"]
    return np.concatenate([np.nonzero(m & (1 << c))[0] + c * nn for c in range(3)])


def run_oracle_hot_path(case):
    from oracle import orc
    c = case
    e2n, G, W, dt = c["e2n"], c["G"], c["W"], c["dt"]
    xend = c["xbeg"] + dt * c["vel"]
    jac = orc.jacobians(G, orc.gather(e2n, xend))
    velE = orc.gather(e2n, c["vel"])
    s1, h1, mg, nfail = orc.model_setup(c["xtal"], c["kin"], c["props"], dt, c["temp_k"], jac, G, velE, c["stress0"],
                                        c["hist0"])
    ess = ess_dofs(c)
    out = dict(jac=jac, stress1=s1, hist1=h1, matgrad=mg, nfail=nfail, velE=velE)
    # residual (MultVec)
    if c["integ"] == 1:
        eds = orc.ic_eds(jac, W, G)
        rE = orc.ic_residual_pa(jac, W, G, eds, s1)
    else:
        rE = orc.residual_pa(jac, W, G, s1)
    r = orc.scatter_add(e2n, rE, c["nn"])
    r[ess] = 0.0
    out["resid_E"], out["resid"] = rE, r
    # gradient apply (TMult<false>) and diagonal
    xm = c["xvec"].copy()
    xm[ess] = 0.0
    xE = orc.gather(e2n, xm)
    if c["assembly"] == 0:
        # PA: plain gradient operator also with B-bar integration (the reference's ICExaNLFIntegrator inherits
        # AssembleGradPA / AddMultGradPA); the diagonal is the B-bar one then (AssembleGradDiagonalPA override)
        yE = orc.grad_mult_pa(dt, jac, W, G, mg, xE)
        dE = orc.ic_grad_diag_pa(dt, jac, W, G, eds, mg) if c["integ"] == 1 else orc.grad_diag_pa(dt, jac, W, G, mg)
        out["y_grad_E_full"] = orc.grad_mult_pa(dt, jac, W, G, mg, orc.gather(e2n, c["xvec"]))
    else:
        ea = orc.ic_assemble_ea(dt, jac, W, G, eds, mg) if c["integ"] == 1 else orc.assemble_ea(dt, jac, W, G, mg)
        yE = orc.ea_mult(ea, xE)
        dE = orc.ea_diag(ea)
        out["ea"] = ea
        out["y_grad_E_full"] = orc.ea_mult(ea, orc.gather(e2n, c["xvec"]))
    y = orc.scatter_add(e2n, yE, c["nn"])
    y[ess] = 0.0
    d = orc.scatter_add(e2n, dE, c["nn"])
    d[ess] = 1.0
    out["y_grad"], out["diag"], out["diag_E"] = y, d, dE
    sums, vol = orc.vol_sum(jac, W, s1, 6)
    out["avg_stress"] = sums / vol
    out["vol"] = vol
    return out


def run_gpu_hot_path(case, evec=False, ctas_per_sm=None):
    import torch
    from exaconstit_b200 import capi
    c = case
    dev = torch.device("cuda")
    f64 = dict(dtype=torch.float64, device=dev)
    T = lambda a: torch.tensor(np.ascontiguousarray(a), **f64)
    ne, nn, nsv, dt = c["ne"], c["nn"], c["nsv"], c["dt"]
    ctx = capi.Context(c["xtal"], c["kin"], c["props"], c["temp_k"], ne, nn, c["e2n"], c["assembly"], c["integ"])
    if ctas_per_sm:
        ctx.set_tuning(ctas_per_sm)
    assert ctx.nstatev == nsv
    ctx.set_essential_mask(c["essmask"])
    xbeg, vel = T(c["xbeg"]), T(c["vel"])
    jac = torch.empty(ne * 72, **f64)
    ctx.setup_jacobians(xbeg, vel, dt, jac)
    s0, h0 = T(c["stress0"]), T(c["hist0"])
    s1 = torch.empty_like(s0)
    h1 = torch.empty_like(h0)
    mg = torch.empty(ne * 8 * 36, **f64)
    if evec:
        from oracle import orc
        ctx.model_setup_evec(dt, jac, T(orc.gather(c["e2n"], c["vel"])), s0, h0, s1, h1, mg)
    else:
        ctx.model_setup(dt, jac, vel, s0, h0, s1, h1, mg)
    nfail = ctx.failed_points()
    out = dict(jac=jac.cpu().numpy(), stress1=s1.cpu().numpy(), hist1=h1.cpu().numpy(), matgrad=mg.cpu().numpy(),
               nfail=nfail)
    r = torch.empty(3 * nn, **f64)
    ctx.residual(jac, s1, r)
    out["resid"] = r.cpu().numpy()
    rE = torch.zeros(ne * 24, **f64)
    ctx.residual_evec(jac, s1, rE)
    out["resid_E"] = rE.cpu().numpy()
    ctx.grad_setup(dt, mg, jac)
    x = T(c["xvec"])
    y = torch.empty(3 * nn, **f64)
    ctx.grad_mult(x, y)
    out["y_grad"] = y.cpu().numpy()
    from oracle import orc
    xE = T(orc.gather(c["e2n"], c["xvec"]))
    yE = torch.zeros(ne * 24, **f64)
    ctx.grad_mult_evec(xE, yE)
    out["y_grad_E_full"] = yE.cpu().numpy()
    d = torch.empty(3 * nn, **f64)
    ctx.grad_diag(d)
    out["diag"] = d.cpu().numpy()
    dE = torch.zeros(ne * 24, **f64)
    ctx.grad_diag_evec(dE)
    out["diag_E"] = dE.cpu().numpy()
    if c["assembly"] == 1:
        ea = torch.zeros(ne * 576, **f64)
        ctx.ea_assemble(dt, mg, jac, ea)
        out["ea"] = ea.cpu().numpy()
    vs = torch.empty(7, **f64)
    ctx.vol_sum(jac, s1, 6, vs)
    vs = vs.cpu().numpy()
    out["avg_stress"] = vs[:6] / vs[6]
    out["vol"] = vs[6]
    out["launches"] = ctx.launch_count()
    torch.cuda.synchronize()
    ctx.close()
    return out


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
