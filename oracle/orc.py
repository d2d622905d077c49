"""ORACLE (test infrastructure only).  ctypes binding of oracle/build/liborc.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path (exaconstit_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "build", "liborc.so")

FCC, BCC, HCP = 0, 1, 2
VOCE, VOCE_NL, KMBALD = 0, 1, 2


def build(force=False):
    """make decides whether the library is stale (its rule lists every source); a box without make uses what is there"""
    try:
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        if not os.path.exists(_LIB):
            raise
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_sim_run2.restype = C.c_int
        _lib.orc_sim_create.restype = C.c_void_p
    return _lib


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# (kirchhoff_rss, eos_temperature, hard_lag, av_power, slip_stretch_terms, eos_mu_form, vol_convect): see ecm::Options
DEFAULT_OPTS = (1.0, 1.0, 1.0, 0.0, 0.0, 1.0, 1.0)


def num_threads():
    return lib().orc_num_threads()


def use_all_host_threads():
    """Launchers (torch.distributed.run) export OMP_NUM_THREADS=1; bench.py's CPU legs must use every host thread
    this process may run on.  Returns the count now in effect."""
    n = len(os.sched_getaffinity(0))
    lib().orc_set_num_threads(n)
    assert num_threads() == n, (num_threads(), n)
    return n


def hex8_dshape():
    G = np.zeros(192)
    W = np.zeros(8)
    lib().orc_hex8_dshape(_p(G), _p(W))
    return G, W


def voxel_mesh(nx, ny, nz, length=(1.0, 1.0, 1.0)):
    ne = nx * ny * nz
    nn = (nx + 1) * (ny + 1) * (nz + 1)
    e2n = np.zeros(ne * 8, dtype=np.int32)
    coords = np.zeros(3 * nn)
    lib().orc_voxel_mesh(nx, ny, nz, _p(_d(length)), _p(e2n), _p(coords))
    return e2n, coords


def gather(e2n, xL):
    ne = e2n.size // 8
    nn = xL.size // 3
    xE = np.zeros(ne * 24)
    lib().orc_gather(C.c_long(ne), C.c_long(nn), _p(e2n), _p(_d(xL)), _p(xE))
    return xE


def scatter_add(e2n, yE, nn):
    ne = e2n.size // 8
    yL = np.zeros(3 * nn)
    lib().orc_scatter_add(C.c_long(ne), C.c_long(nn), _p(e2n), _p(_d(yE)), _p(yL))
    return yL


def jacobians(G, xE):
    ne = xE.size // 24
    jac = np.zeros(ne * 72)
    lib().orc_jacobians(C.c_long(ne), _p(G), _p(_d(xE)), _p(jac))
    return jac


def grad_calc(jac, G, xE):
    ne = xE.size // 24
    out = np.zeros(ne * 72)
    lib().orc_grad_calc(C.c_long(ne), _p(jac), _p(G), _p(_d(xE)), _p(out))
    return out


def transform_matgrad_4d(k36):
    npts = k36.size // 36
    c81 = np.zeros(npts * 81)
    lib().orc_transform_matgrad_4d(C.c_long(npts), _p(_d(k36)), _p(c81))
    return c81


def residual_pa(jac, W, G, stress):
    ne = jac.size // 72
    d = np.zeros(ne * 72)
    y = np.zeros(ne * 24)
    lib().orc_assemble_pa(C.c_long(ne), _p(jac), _p(W), _p(_d(stress)), _p(d))
    lib().orc_addmult_pa(C.c_long(ne), _p(G), _p(d), _p(y))
    return y


def grad_mult_pa(dt, jac, W, G, k36, xE):
    ne = jac.size // 72
    c81 = transform_matgrad_4d(k36)
    D = np.zeros(ne * 8 * 81)
    lib().orc_assemble_grad_pa(C.c_long(ne), C.c_double(dt), _p(jac), _p(W), _p(c81), _p(D))
    y = np.zeros(ne * 24)
    lib().orc_addmult_grad_pa(C.c_long(ne), _p(G), _p(D), _p(_d(xE)), _p(y))
    return y


def grad_diag_pa(dt, jac, W, G, k36):
    ne = jac.size // 72
    d = np.zeros(ne * 24)
    lib().orc_assemble_grad_diag_pa(C.c_long(ne), C.c_double(dt), _p(jac), _p(W), _p(G), _p(_d(k36)), _p(d))
    return d


def assemble_ea(dt, jac, W, G, k36):
    ne = jac.size // 72
    ea = np.zeros(ne * 576)
    lib().orc_assemble_ea(C.c_long(ne), C.c_double(dt), _p(jac), _p(W), _p(G), _p(_d(k36)), _p(ea))
    return ea


def ea_mult(ea, xE):
    ne = xE.size // 24
    y = np.zeros(ne * 24)
    lib().orc_ea_mult(C.c_long(ne), _p(ea), _p(_d(xE)), _p(y))
    return y


def ea_diag(ea):
    ne = ea.size // 576
    d = np.zeros(ne * 24)
    lib().orc_ea_diag(C.c_long(ne), _p(ea), _p(d))
    return d


def ic_eds(jac, W, G):
    ne = jac.size // 72
    eds = np.zeros(ne * 24)
    lib().orc_ic_assemble_eds(C.c_long(ne), _p(jac), _p(W), _p(G), _p(eds))
    return eds


def ic_residual_pa(jac, W, G, eds, stress):
    ne = jac.size // 72
    y = np.zeros(ne * 24)
    lib().orc_ic_addmult_pa(C.c_long(ne), _p(jac), _p(W), _p(G), _p(eds), _p(_d(stress)), _p(y))
    return y


def ic_assemble_ea(dt, jac, W, G, eds, k36):
    ne = jac.size // 72
    ea = np.zeros(ne * 576)
    lib().orc_ic_assemble_ea(C.c_long(ne), C.c_double(dt), _p(jac), _p(W), _p(G), _p(eds), _p(_d(k36)), _p(ea))
    return ea


def ic_grad_diag_pa(dt, jac, W, G, eds, k36):
    ne = jac.size // 72
    d = np.zeros(ne * 24)
    lib().orc_ic_assemble_grad_diag_pa(C.c_long(ne), C.c_double(dt), _p(jac), _p(W), _p(G), _p(eds), _p(_d(k36)), _p(d))
    return d


def vol_sum(jac, W, qf, vdim):
    ne = jac.size // 72
    sums = np.zeros(vdim)
    vol = np.zeros(1)
    lib().orc_vol_sum(C.c_long(ne), vdim, _p(jac), _p(W), _p(_d(qf)), _p(sums), _p(vol))
    return sums, float(vol[0])


def nhist(xtal, kin):
    return lib().orc_nhist(xtal, kin)


def hist_init(xtal, kin, props):
    h = np.zeros(nhist(xtal, kin))
    props = _d(props)
    rc = lib().orc_hist_init(xtal, kin, _p(props), props.size, _p(h))
    if rc:
        raise ValueError("bad property vector rc=%d" % rc)
    return h


def model_setup(xtal, kin, props, dt, temp_k, jac, G, velE, stress0, hist0, opts=DEFAULT_OPTS):
    """ExaCMechModel::ModelSetup -> (stress1, hist1, ddsdde, nfail)."""
    ne = jac.size // 72
    props = _d(props)
    stress1 = np.zeros(ne * 48)
    hist1 = np.zeros_like(hist0)
    dd = np.zeros(ne * 8 * 36)
    o = _d(opts)
    rc = lib().orc_model_setup(xtal, kin, _p(props), props.size, _p(o), C.c_long(ne), C.c_double(dt),
                               C.c_double(temp_k), _p(_d(jac)), _p(_d(G)), _p(_d(velE)), _p(_d(stress0)),
                               _p(_d(hist0)), _p(stress1), _p(hist1), _p(dd))
    return stress1, hist1, dd, rc


def local_problem(xtal, kin, props, dt, d_svec_p, w_vec, vnew, hist, tK, x, opts=DEFAULT_OPTS):
    props = _d(props)
    R = np.zeros(8)
    J = np.zeros(64)
    o = _d(opts)
    rc = lib().orc_local_problem(xtal, kin, _p(props), props.size, _p(o), C.c_double(dt), _p(_d(d_svec_p)),
                                 _p(_d(w_vec)), C.c_double(vnew), _p(_d(hist)), C.c_double(tK), _p(_d(x)),
                                 _p(R), _p(J))
    assert rc == 0
    return R, J.reshape(8, 8)


def sim_run(n, length, xtal, kin, props, temp_k, grain_ids, quats, dts, bcs, assembly=0, integ=0,
            nl_solver=0, nr=(5e-5, 5e-10, 25), kr=(1e-7, 1e-27, 1000), true_jacobi=False,
            opts=DEFAULT_OPTS, verbose=0, want_state=False, auto_time=None):
    """bcs: list of (step, ids, comps, vals[, vgrad 3x3]); comps < 0 = velocity-gradient BC.
    auto_time: dict(dt_start, dt_min, dt_scale, t_final) -> Time.Auto stepping (dts is ignored).  Returns dict."""
    nx, ny, nz = n
    props = _d(props)
    grain_ids = _i(grain_ids)
    quats = _d(quats)
    at = np.zeros(5)
    if auto_time:
        at[:] = [1.0, auto_time["dt_start"], auto_time["dt_min"], auto_time["dt_scale"], auto_time["t_final"]]
        dts = np.zeros(int(np.ceil(auto_time["t_final"] / auto_time["dt_min"])))
    dts = _d(dts)
    nsteps = dts.size
    bc_steps = _i([b[0] for b in bcs])
    bc_counts = _i([len(b[1]) for b in bcs])
    bc_ids = _i(np.concatenate([b[1] for b in bcs]))
    bc_comps = _i(np.concatenate([b[2] for b in bcs]))
    bc_vals = _d(np.concatenate([np.asarray(b[3], dtype=float).ravel() for b in bcs]))
    bc_vgrads = _d(np.concatenate([np.asarray(b[4], dtype=float).ravel() if len(b) > 4 else np.zeros(9) for b in bcs]))
    out_stress = np.zeros((nsteps, 6))
    out_extra = np.zeros((nsteps, 16))
    out_iters = np.zeros((nsteps, 2), dtype=np.int32)
    out_stats = np.zeros(7)
    out_dts = np.zeros(nsteps)
    nh = nhist(xtal, kin)
    npts = nx * ny * nz * 8
    out_hist = np.zeros(npts * nh) if want_state else None
    out_sq = np.zeros(npts * 6) if want_state else None
    o = _d(opts)
    rc = lib().orc_sim_run2(nx, ny, nz, _p(_d(length)), xtal, kin, _p(props), props.size, C.c_double(temp_k),
                            _p(grain_ids), _p(quats), quats.size // 4, _p(dts), nsteps, len(bcs), _p(bc_steps),
                            _p(bc_counts), _p(bc_ids), _p(bc_comps), _p(bc_vals), _p(bc_vgrads), assembly, integ,
                            nl_solver, _p(_d(nr)), _p(_d(kr)), int(true_jacobi), _p(o), verbose, _p(at),
                            _p(out_stress), _p(out_extra), _p(out_iters), _p(out_stats), _p(out_hist), _p(out_sq),
                            _p(out_dts))
    taken = int(out_stats[6]) if rc == 0 else nsteps
    return dict(rc=rc, stress=out_stress[:taken], extra=out_extra[:taken], iters=out_iters[:taken], dts=out_dts[:taken],
                stats=dict(newton_iters=int(out_stats[0]), pcg_iters=int(out_stats[1]),
                           model_setups=int(out_stats[2]), grad_mults=int(out_stats[3]),
                           failed_points=int(out_stats[4]), seconds=float(out_stats[5])),
                hist=out_hist, stress_qp=out_sq)


class SimStepper:
    """The same simulation as sim_run, driven one time step at a time (bench.py's CPU legs time single steps)."""

    def __init__(self, n, length, xtal, kin, props, temp_k, grain_ids, quats, bcs, assembly=0, integ=0, nl_solver=0,
                 nr=(5e-5, 5e-10, 25), kr=(1e-7, 1e-27, 1000), true_jacobi=False, opts=DEFAULT_OPTS, verbose=0):
        nx, ny, nz = n
        props = _d(props)
        quats = _d(quats)
        bc_steps = _i([b[0] for b in bcs])
        bc_counts = _i([len(b[1]) for b in bcs])
        bc_ids = _i(np.concatenate([b[1] for b in bcs]))
        bc_comps = _i(np.concatenate([b[2] for b in bcs]))
        bc_vals = _d(np.concatenate([np.asarray(b[3], dtype=float).ravel() for b in bcs]))
        bc_vgrads = _d(np.concatenate([np.asarray(b[4], dtype=float).ravel() if len(b) > 4 else np.zeros(9) for b in bcs]))
        self._h = C.c_void_p(lib().orc_sim_create(
            nx, ny, nz, _p(_d(length)), xtal, kin, _p(props), props.size, C.c_double(temp_k), _p(_i(grain_ids)), _p(quats),
            quats.size // 4, len(bcs), _p(bc_steps), _p(bc_counts), _p(bc_ids), _p(bc_comps), _p(bc_vals), _p(bc_vgrads),
            assembly, integ, nl_solver, _p(_d(nr)), _p(_d(kr)), int(true_jacobi), _p(_d(opts)), verbose))
        if not self._h:
            raise ValueError("orc_sim_create failed")

    def step(self, dt):
        out = np.zeros(12)
        rc = lib().orc_sim_step(self._h, C.c_double(dt), _p(out))
        return dict(rc=rc, newton_iters=int(out[0]), pcg_iters=int(out[1]), model_setups=int(out[2]), grad_mults=int(out[3]),
                    seconds=float(out[4]), avg_stress=out[5:11].copy(), pcg_seconds=float(out[11]))

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()
