"""ctypes binding of the C++ host layer (include/exahost.h): the SystemDriver / NonlinearMechOperator /
ExaNewtonSolver / CGSolver mirror that drives the exab200 kernels for one z-slab of a voxel mesh."""
import ctypes as C
import os

import numpy as np

from . import capi, voxel

_HERE = os.path.dirname(os.path.abspath(__file__))
# EXAB200_LIBDIR: alternative build directory (tuning experiments only)
LIB_PATH = os.path.join(os.environ.get("EXAB200_LIBDIR", os.path.join(_HERE, "lib")), "libexahost.so")


class HostConfig(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz_local", C.c_int), ("z0", C.c_int), ("nz_total", C.c_int),
                ("length", C.c_double * 3), ("xtal", C.c_int), ("slip", C.c_int), ("nprops", C.c_int),
                ("props", C.POINTER(C.c_double)), ("temp_k", C.c_double), ("grain_ids", C.POINTER(C.c_int)),
                ("quats", C.POINTER(C.c_double)), ("ngrains", C.c_int), ("assembly", C.c_int), ("integ", C.c_int),
                ("nl_solver", C.c_int), ("newton_rel_tol", C.c_double), ("newton_abs_tol", C.c_double),
                ("newton_iter", C.c_int), ("krylov_rel_tol", C.c_double), ("krylov_abs_tol", C.c_double),
                ("krylov_iter", C.c_int), ("true_jacobi", C.c_int), ("rank", C.c_int), ("nranks", C.c_int),
                ("device", C.c_int), ("nccl_id", C.c_void_p), ("verbose", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        capi.lib()  # libexab200 first (RPATH $ORIGIN also finds it)
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("exahost: %s not built" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.exahost_last_error.restype = C.c_char_p
        _lib.exahost_counter.restype = C.c_long
        _lib.exahost_stream.restype = C.c_void_p
        _lib.exahost_ctx.restype = C.c_void_p
    return _lib


class HostError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise HostError(lib().exahost_last_error().decode())


def nccl_unique_id():
    buf = (C.c_char * 128)()
    _chk(lib().exahost_nccl_unique_id(buf))
    return bytes(buf)


_LAYOUT_KEYS = ("z0", "nz_local", "nnodes", "plane", "n_owned", "lo_offset", "hi_offset", "nelems", "has_lo", "has_hi",
                "lo_tiles", "hi_tile_start")


def slab_layout(nx, ny, nz, rank, nranks):
    """exahost_slab_layout: this rank's z-slab of the voxel mesh and the ownership / interface-plane index sets of the
    exchanges (the C++ host layer's own partition arithmetic; no device needed)."""
    out = (C.c_long * 12)()
    _chk(lib().exahost_slab_layout(nx, ny, nz, rank, nranks, out))
    return dict(zip(_LAYOUT_KEYS, [int(v) for v in out]))


class VoxelSim:
    """One rank's slab of a voxel-mesh simulation.  `n` = (nx, ny, nz_total); with nranks > 1 the element
    layers are split by voxel.slab_partition and `grain_ids` is the GLOBAL x-fastest array."""

    def __init__(self, n, length, xtal, slip, props, temp_k, grain_ids, quats, assembly=0, integ=0, nl_solver=0,
                 nr=(5e-5, 5e-10, 25), kr=(1e-7, 1e-27, 1000), true_jacobi=False, rank=0, nranks=1, device=0,
                 nccl_id=None, verbose=0):
        nx, ny, nz = n
        lay = slab_layout(nx, ny, nz, rank, nranks)
        self.z0, self.nzl = lay["z0"], lay["nz_local"]
        self.nx, self.ny, self.nz = nx, ny, nz
        self.rank, self.nranks = rank, nranks
        self.nelems = nx * ny * self.nzl
        self.plane = (nx + 1) * (ny + 1)
        self.nnodes = self.plane * (self.nzl + 1)
        gl = np.ascontiguousarray(np.asarray(grain_ids, dtype=np.int32)[nx * ny * self.z0: nx * ny * (self.z0 + self.nzl)])
        self._keep = dict(props=np.ascontiguousarray(props, dtype=np.float64), grains=gl,
                          quats=np.ascontiguousarray(quats, dtype=np.float64).reshape(-1))
        cfg = HostConfig()
        cfg.nx, cfg.ny, cfg.nz_local, cfg.z0, cfg.nz_total = nx, ny, self.nzl, self.z0, nz
        cfg.length = (C.c_double * 3)(*length)
        cfg.xtal, cfg.slip, cfg.nprops = xtal, slip, self._keep["props"].size
        cfg.props = self._keep["props"].ctypes.data_as(C.POINTER(C.c_double))
        cfg.temp_k = temp_k
        cfg.grain_ids = gl.ctypes.data_as(C.POINTER(C.c_int))
        cfg.quats = self._keep["quats"].ctypes.data_as(C.POINTER(C.c_double))
        cfg.ngrains = self._keep["quats"].size // 4
        cfg.assembly, cfg.integ, cfg.nl_solver = assembly, integ, nl_solver
        cfg.newton_rel_tol, cfg.newton_abs_tol, cfg.newton_iter = nr[0], nr[1], int(nr[2])
        cfg.krylov_rel_tol, cfg.krylov_abs_tol, cfg.krylov_iter = kr[0], kr[1], int(kr[2])
        cfg.true_jacobi = int(true_jacobi)
        cfg.rank, cfg.nranks, cfg.device = rank, nranks, device
        if nranks > 1:
            assert nccl_id is not None and len(nccl_id) == 128
            self._keep["id"] = C.create_string_buffer(nccl_id, 128)
            cfg.nccl_id = C.cast(self._keep["id"], C.c_void_p)
        cfg.verbose = verbose
        h = C.c_void_p()
        _chk(lib().exahost_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.nstatev = lib().exahost_counter(self._h, 7)

    def close(self):
        if getattr(self, "_h", None):
            lib().exahost_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_bcs(self, ids, comps, vals, vgrad=None):
        """UpdateEssBdr: essential boundary attributes / component codes / velocities (src/BCManager.cpp); negative
        component codes mark velocity-gradient attributes driven by the 3x3 `vgrad` (BCs.essential_vel_grad)."""
        mask, val = voxel.essential_bcs(self.nx, self.ny, self.nzl, ids, comps, vals, self.z0, self.nz)
        vg = voxel.vgrad_mask(self.nx, self.ny, self.nzl, ids, comps, self.z0, self.nz)
        val[np.concatenate([np.nonzero(vg & (1 << d))[0] + d * self.nnodes for d in range(3)])] = 0.0
        self._ess_val = val
        _chk(lib().exahost_set_bcs(self._h, mask.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p)))
        if any(c < 0 for c in comps):
            L = np.ascontiguousarray(np.asarray(vgrad, dtype=np.float64).reshape(9))
            _chk(lib().exahost_set_vgrad(self._h, vg.ctypes.data_as(C.c_void_p), L.ctypes.data_as(C.c_void_p)))
        else:
            _chk(lib().exahost_set_vgrad(self._h, None, None))
        return mask, val

    def step(self, dt, bc_changed=False, ess_val_host=None, vel_out_host=None):
        out = np.zeros(16)
        pin = ess_val_host.ctypes.data_as(C.c_void_p) if ess_val_host is not None else None
        pout = vel_out_host.ctypes.data_as(C.c_void_p) if vel_out_host is not None else None
        _chk(lib().exahost_step(self._h, C.c_double(dt), int(bc_changed), pin, pout, out.ctypes.data_as(C.c_void_p)))
        return dict(newton_iters=int(out[0]), pcg_iters=int(out[1]), converged=bool(out[2]), model_setups=int(out[3]),
                    grad_mults=int(out[4]), seconds=float(out[5]), avg_stress=out[6:12].copy(),
                    dev_ms=float(out[12]), e2e_ms=float(out[13]))

    def step_auto(self, ctl, bc_changed=False):
        """One Time.Auto step; ctl = np.array([dt_class, t, dt_min, dt_scale, t_final, last_step]) is updated in place."""
        out = np.zeros(16)
        _chk(lib().exahost_step_auto(self._h, ctl.ctypes.data_as(C.c_void_p), int(bc_changed), None, None,
                                     out.ctypes.data_as(C.c_void_p)))
        return dict(newton_iters=int(out[0]), pcg_iters=int(out[1]), converged=bool(out[2]), model_setups=int(out[3]),
                    grad_mults=int(out[4]), seconds=float(out[5]), avg_stress=out[6:12].copy(),
                    dev_ms=float(out[12]), e2e_ms=float(out[13]), dt=float(out[14]))

    def extra_avgs(self):
        """additional_avgs of UpdateModel: plastic-work integral, <F> (9, [t*3+i]), <D^p> (6 Voigt)."""
        out = np.zeros(16)
        _chk(lib().exahost_extra_avgs(self._h, out.ctypes.data_as(C.c_void_p)))
        return dict(pl_work=float(out[0]), def_grad=out[1:10].copy(), dp=out[10:16].copy())

    def get(self, which):
        sizes = {"stress": self.nelems * 48, "hist": self.nelems * 8 * self.nstatev, "vel": 3 * self.nnodes,
                 "xbeg": 3 * self.nnodes}
        idx = {"stress": 0, "hist": 1, "vel": 2, "xbeg": 3}[which]
        out = np.zeros(sizes[which])
        _chk(lib().exahost_get(self._h, idx, out.ctypes.data_as(C.c_void_p)))
        return out

    def counter(self, name):
        idx = {"launches": 0, "allreduces": 1, "halos": 2, "model_setups": 3, "grad_mults": 4, "pcg_iters": 5,
               "newton_iters": 6}[name]
        return lib().exahost_counter(self._h, idx)

    def enable_peer_collectives(self, dist):
        """Switch the CG-loop exchanges from NCCL to the NVLink peer-memory kernels: all-gather the CUDA-IPC
        handles of the ranks' mailboxes over `dist` (torch.distributed) and map them."""
        import torch
        if self.nranks == 1:
            return False
        buf = (C.c_char * 64)()
        _chk(lib().exahost_comm_handle(self._h, buf))
        mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(self.nranks)]
        dist.all_gather(allh, mine)
        blob = b"".join(bytes(t.cpu().tolist()) for t in allh)
        ok = torch.ones(1, dtype=torch.int32, device="cuda")
        try:
            _chk(lib().exahost_set_peers(self._h, C.create_string_buffer(blob, len(blob))))
        except HostError as e:   # no peer access to a neighbour (topology / IPC restrictions)
            ok.zero_()
            self.peer_error = str(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)        # all ranks or none: the protocols must match
        if int(ok.item()) == 0:
            _chk(lib().exahost_set_peers(self._h, None))
            dist.barrier()
            return False
        dist.barrier()
        return True

    def kernel_timing(self, enable=True):
        lib().exahost_kernel_timing(self._h, int(enable))

    def kernel_time(self, which, reset=False):
        tot, cnt = C.c_double(0), C.c_long(0)
        lib().exahost_kernel_time(self._h, {"grad_mult": 0, "model_setup": 1}[which], C.byref(tot), C.byref(cnt), int(reset))
        return tot.value, cnt.value

    def set_deterministic(self, on=True):
        """Owner-computes scatter (exab200_set_deterministic): bitwise reproducible operator, residual and diagonal."""
        _chk(lib().exahost_set_tuning(self._h, 1, 97 if on else 98))

    def set_tuning(self, ctas_per_sm, variant):
        _chk(lib().exahost_set_tuning(self._h, ctas_per_sm, variant))

    def run(self, dts, bcs, extras=False):
        """The reference's time loop (src/mechanics_driver.cpp:837-907). bcs: list of (step, ids, comps, vals)."""
        hist = []
        for ti, dt in enumerate(dts, start=1):
            changed = False
            for b in bcs:
                if b[0] == ti:
                    self.set_bcs(b[1], b[2], b[3], b[4] if len(b) > 4 else None)
                    changed = True
            hist.append(self.step(float(dt), bc_changed=changed))
            if extras:
                hist[-1].update(self.extra_avgs())
        return hist

    def run_auto(self, auto_time, bcs):
        """The reference's time loop with Time.Auto (src/mechanics_driver.cpp:212,837-967): at most
        ceil(t_final / dt_min) steps, stops at the last step."""
        ctl = np.array([auto_time["dt_start"], 0.0, auto_time["dt_min"], auto_time["dt_scale"], auto_time["t_final"], 0.0])
        hist = []
        for ti in range(1, int(np.ceil(auto_time["t_final"] / auto_time["dt_min"])) + 1):
            changed = False
            for b in bcs:
                if b[0] == ti:
                    self.set_bcs(b[1], b[2], b[3], b[4] if len(b) > 4 else None)
                    changed = True
            hist.append(self.step_auto(ctl, bc_changed=changed))
            if ctl[5] != 0.0:
                break
        return hist
