"""ctypes binding of the exab200 C ABI (include/exab200.h).

PyTorch is used only as the owner of device memory and streams: every array crossing the ABI is
a CUDA float64/int32 tensor whose ``data_ptr()`` is handed to the library.  There is no CPU
fallback: if ``libexab200.so`` is missing or fails to load this module raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# EXAB200_LIBDIR: alternative build directory (tuning experiments only)
LIB_PATH = os.path.join(os.environ.get("EXAB200_LIBDIR", os.path.join(_HERE, "lib")), "libexab200.so")

FCC, BCC, HCP = 0, 1, 2
POWERVOCE, POWERVOCENL, MTSDD = 0, 1, 2
PA, EA = 0, 1
INTEG_FULL, INTEG_BBAR = 0, 1

# every symbol include/exab200.h declares
SYMBOLS = [
    "exab200_last_error", "exab200_version", "exab200_create", "exab200_destroy",
    "exab200_num_state_vars", "exab200_set_essential_mask", "exab200_hist_init",
    "exab200_setup_jacobians", "exab200_model_setup", "exab200_model_setup_evec",
    "exab200_failed_points", "exab200_residual_evec", "exab200_residual", "exab200_grad_setup",
    "exab200_grad_mult_evec", "exab200_grad_mult", "exab200_grad_mult_ex", "exab200_grad_mult_halo_supported",
    "exab200_grad_mult_halo", "exab200_set_deterministic", "exab200_grad_diag_evec", "exab200_grad_diag",
    "exab200_ea_assemble", "exab200_ea_mult_evec", "exab200_vol_sum", "exab200_calc_dp",
    "exab200_grad_calc", "exab200_launch_count", "exab200_set_tuning", "exab200_set_tangent_format",
]


class Config(C.Structure):
    _fields_ = [("xtal", C.c_int), ("slip", C.c_int), ("nprops", C.c_int), ("props", C.POINTER(C.c_double)),
                ("temp_k", C.c_double), ("nelems", C.c_long), ("nnodes", C.c_long), ("e2n", C.POINTER(C.c_int)),
                ("assembly", C.c_int), ("integ", C.c_int), ("device", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("exab200: %s not built -- run `python -c 'import __graft_entry__ as g; g.build()'`"
                               % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.exab200_last_error.restype = C.c_char_p
        _lib.exab200_launch_count.restype = C.c_long
    return _lib


class Exab200Error(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise Exab200Error(lib().exab200_last_error().decode())


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "device tensors must be contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Context:
    """Owns one exab200_ctx.  Mirrors the construction of an ECMechXtalModel + the operator's
    element scratch (src/mechanics_ecmech.hpp:126-245, src/mechanics_operator.cpp:227-262)."""

    def __init__(self, xtal, slip, props, temp_k, nelems, nnodes, e2n=None, assembly=PA, integ=INTEG_FULL,
                 device=0):
        props = np.ascontiguousarray(props, dtype=np.float64)
        self._props = props
        cfg = Config()
        cfg.xtal, cfg.slip, cfg.nprops = xtal, slip, props.size
        cfg.props = props.ctypes.data_as(C.POINTER(C.c_double))
        cfg.temp_k = temp_k
        cfg.nelems, cfg.nnodes = nelems, nnodes
        if e2n is not None:
            e2n = np.ascontiguousarray(e2n, dtype=np.int32)
            assert e2n.size == 8 * nelems
            cfg.e2n = e2n.ctypes.data_as(C.POINTER(C.c_int))
        self._e2n = e2n
        cfg.assembly, cfg.integ, cfg.device = assembly, integ, device
        self.nelems, self.nnodes = nelems, nnodes
        self.assembly, self.integ = assembly, integ
        h = C.c_void_p()
        _chk(lib().exab200_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.nstatev = lib().exab200_num_state_vars(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().exab200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- thin 1:1 wrappers ------------------------------------------------------------------
    def set_essential_mask(self, mask):
        if mask is None:
            _chk(lib().exab200_set_essential_mask(self._h, None))
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            assert m.size == self.nnodes
            _chk(lib().exab200_set_essential_mask(self._h, m.ctypes.data_as(C.c_void_p)))

    def hist_init(self, hist):
        _chk(lib().exab200_hist_init(self._h, _ptr(hist), _stream()))

    def setup_jacobians(self, xbeg, vel, dt, jac):
        _chk(lib().exab200_setup_jacobians(self._h, _ptr(xbeg), _ptr(vel), C.c_double(dt), _ptr(jac), _stream()))

    def model_setup(self, dt, jac, vel_L, stress0, hist0, stress1, hist1, matgrad):
        _chk(lib().exab200_model_setup(self._h, C.c_double(dt), _ptr(jac), _ptr(vel_L), _ptr(stress0), _ptr(hist0),
                                       _ptr(stress1), _ptr(hist1), _ptr(matgrad), _stream()))

    def model_setup_evec(self, dt, jac, vel_E, stress0, hist0, stress1, hist1, matgrad):
        _chk(lib().exab200_model_setup_evec(self._h, C.c_double(dt), _ptr(jac), _ptr(vel_E), _ptr(stress0),
                                            _ptr(hist0), _ptr(stress1), _ptr(hist1), _ptr(matgrad), _stream()))

    def failed_points(self):
        out = C.c_int(0)
        _chk(lib().exab200_failed_points(self._h, _stream(), C.byref(out)))
        return out.value

    def residual_evec(self, jac, stress, yE):
        _chk(lib().exab200_residual_evec(self._h, _ptr(jac), _ptr(stress), _ptr(yE), _stream()))

    def residual(self, jac, stress, yL):
        _chk(lib().exab200_residual(self._h, _ptr(jac), _ptr(stress), _ptr(yL), _stream()))

    def grad_setup(self, dt, matgrad, jac):
        self._keep = (matgrad, jac)
        _chk(lib().exab200_grad_setup(self._h, C.c_double(dt), _ptr(matgrad), _ptr(jac), _stream()))

    def grad_mult_evec(self, xE, yE):
        _chk(lib().exab200_grad_mult_evec(self._h, _ptr(xE), _ptr(yE), _stream()))

    def grad_mult(self, xL, yL, local_action=False):
        _chk(lib().exab200_grad_mult(self._h, _ptr(xL), _ptr(yL), int(local_action), _stream()))

    def grad_mult_ex(self, xL, yL, flags=0, dot_accum=None):
        _chk(lib().exab200_grad_mult_ex(self._h, _ptr(xL), _ptr(yL), int(flags), _ptr(dot_accum), _stream()))

    def grad_diag_evec(self, dE):
        _chk(lib().exab200_grad_diag_evec(self._h, _ptr(dE), _stream()))

    def grad_diag(self, dL):
        _chk(lib().exab200_grad_diag(self._h, _ptr(dL), _stream()))

    def ea_assemble(self, dt, matgrad, jac, emat):
        _chk(lib().exab200_ea_assemble(self._h, C.c_double(dt), _ptr(matgrad), _ptr(jac), _ptr(emat), _stream()))

    def ea_mult_evec(self, emat, xE, yE):
        _chk(lib().exab200_ea_mult_evec(self._h, _ptr(emat), _ptr(xE), _ptr(yE), _stream()))

    def vol_sum(self, jac, qf, vdim, out):
        _chk(lib().exab200_vol_sum(self._h, _ptr(jac), _ptr(qf), vdim, _ptr(out), _stream()))

    def calc_dp(self, hist, dp):
        _chk(lib().exab200_calc_dp(self._h, _ptr(hist), _ptr(dp), _stream()))

    def grad_calc(self, jac, field_L, grad):
        _chk(lib().exab200_grad_calc(self._h, _ptr(jac), _ptr(field_L), _ptr(grad), _stream()))

    def launch_count(self):
        return lib().exab200_launch_count(self._h)

    def set_deterministic(self, on=True):
        _chk(lib().exab200_set_deterministic(self._h, int(on)))

    def set_tangent_format(self, fmt):
        _chk(lib().exab200_set_tangent_format(self._h, int(fmt)))

    def set_tuning(self, ctas_per_sm, variant=0):
        _chk(lib().exab200_set_tuning(self._h, ctas_per_sm, variant))
