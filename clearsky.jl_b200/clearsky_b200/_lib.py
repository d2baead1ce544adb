"""ctypes binding of libclearsky_b200.so -- the same call sequence the Julia wrapper issues via ccall.

There is no CPU fallback: if the shared library is missing this module raises at import of the first
symbol, and every compute entry point fails with CS_ERR_CUDA when no B200 is present.
"""
import ctypes as C
import os

import numpy as np

CS_OK, CS_ERR_CUDA, CS_ERR_ARG, CS_ERR_NOMEM, CS_ERR_DOMAIN = 0, 1, 2, 3, 4
CS_DOPPLER, CS_LORENTZ, CS_VOIGT, CS_PHCO2 = 0, 1, 2, 3
CS_FARFIELD_DIRECT, CS_FARFIELD_EXPANSION = 0, 1
FARFIELD_MODES = {"direct": CS_FARFIELD_DIRECT, "expansion": CS_FARFIELD_EXPANSION}
CS_MAXCHEB = 16
CS_NTIMERS = 8
TIMER_NAMES = ("prep", "linesum", "table_fit", "table_eval", "cia", "rt", "reduce", "total")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLEARSKY_B200_LIB",
                          os.path.join(os.path.dirname(_HERE), "lib", "libclearsky_b200.so"))

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p

# name -> argtypes; every function returns int32 status except cs_last_error / cs_version
SIGNATURES = {
    "cs_device_count": [C.POINTER(C.c_int32)],
    "cs_ctx_create": [C.c_int32, C.POINTER(_vp)],
    "cs_ctx_create_on_stream": [C.c_int32, _vp, C.POINTER(_vp)],
    "cs_ctx_free": [_vp],
    "cs_ctx_synchronize": [_vp],
    "cs_ctx_timers": [_vp, _dp],
    "cs_ctx_timers_total": [_vp, _dp],
    "cs_ctx_launches": [_vp, _i64p],
    "cs_ctx_set_farfield": [_vp, C.c_int32],
    "cs_ctx_get_farfield": [_vp, C.POINTER(C.c_int32)],
    "cs_ctx_set_tau_floor": [_vp, C.c_double],
    "cs_fp64_peak": [_vp, C.c_int32, _dp],
    "cs_lines_upload": [_vp, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int16), C.c_int32,
                        C.POINTER(C.c_int32), _dp, C.POINTER(C.c_uint8), C.POINTER(_vp)],
    "cs_lines_free": [_vp],
    "cs_lines_set_grid_range": [_vp, C.c_double, C.c_double],
    "cs_line_params": [_vp, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp],
    "cs_xsec": [_vp, C.c_int32, C.c_int64, _dp, C.c_int64, _dp, _dp, _dp, C.c_double, _dp],
    "cs_count_evals": [_vp, C.c_int64, _dp, C.c_double, _i64p],
    "cs_bake": [_vp, C.c_int32, C.c_int64, _dp, C.c_int32, _dp, C.c_int32, _dp, _dp, C.c_double, C.c_int32,
                C.POINTER(_vp)],
    "cs_table_from_block": [_vp, C.c_int64, C.c_int32, _dp, C.c_int32, _dp, _dp, C.POINTER(_vp)],
    "cs_table_eval": [_vp, C.c_int64, _dp, _dp, _dp],
    "cs_table_block": [_vp, _dp],
    "cs_table_info": [_vp, _i64p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i64p],
    "cs_table_free": [_vp],
    "cs_cia_upload": [_vp, C.c_int32, _i64p, _i64p, _dp, _dp, _dp, C.c_int32, _i64p, _dp, _dp, C.c_int32,
                      C.c_int32, C.POINTER(_vp)],
    "cs_cia_free": [_vp],
    "cs_accel_from_sigma": [_vp, _dp, C.POINTER(_vp)],
    "cs_accel_upload": [_vp, C.c_int64, C.c_int64, _dp, _dp, C.POINTER(_vp)],
    "cs_accel_free": [_vp],
    "cs_sigma_create": [_vp, C.c_int64, _dp, C.c_int64, C.POINTER(_vp)],
    "cs_sigma_zero": [_vp],
    "cs_sigma_free": [_vp],
    "cs_sigma_add_table": [_vp, _vp, _dp, _dp, _dp],
    "cs_sigma_add_lines": [_vp, _vp, C.c_int32, _dp, _dp, _dp, C.c_double],
    "cs_sigma_add_cia": [_vp, _vp, _dp, _dp, _dp, _dp],
    "cs_sigma_add_accel": [_vp, _vp, _dp],
    "cs_sigma_add_host": [_vp, _dp],
    "cs_sigma_add_gray": [_vp, C.c_double, C.c_double],
    "cs_sigma_read": [_vp, _dp],
    "cs_fluxes": [_vp, C.c_int64, _dp, C.c_int32, _dp, _dp, _dp, C.c_double, _dp, _dp, C.c_double, C.c_int32,
                  _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp],
    "cs_fluxes_device": [_vp, C.c_int64, _dp, C.c_int32, _dp, _dp, _dp, C.c_double, _dp, _dp, C.c_double,
                         C.c_int32, _dp, _dp, _dp, _vp],
    "cs_fluxes_batch": [_vp, C.c_int64, _dp, C.c_int32, _dp, _dp, C.c_int64, _dp, C.c_double, _dp, _dp, C.c_double,
                        C.c_int32, _dp, _dp, _dp, _dp],
    "cs_opticaldepth": [_vp, C.c_int64, _dp, C.c_int32, _dp, _dp, C.c_double, C.c_double, _dp],
    "cs_par_parse": [_vp, C.c_int64, C.c_char_p, C.c_int32, C.c_int64, C.POINTER(C.c_int16), C.POINTER(C.c_int16),
                     _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_uint8)],
    "cs_par_read": [_vp, C.c_int64, C.c_char_p, C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int32,
                    C.POINTER(C.c_int16), C.c_int64, C.POINTER(C.c_int16), C.POINTER(C.c_int16), _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp,
                    _i64p, _i64p, _i64p],
    "cs_group_create": [C.c_int32, C.POINTER(C.c_int32), C.POINTER(_vp)],
    "cs_group_free": [_vp],
    "cs_group_size": [_vp, C.POINTER(C.c_int32)],
    "cs_group_ctx": [_vp, C.c_int32, C.POINTER(_vp)],
    "cs_group_buffer": [_vp, C.c_int32, C.c_int64, C.POINTER(_vp)],
    "cs_group_allreduce_sum": [_vp, C.c_int64],
    "cs_group_read": [_vp, C.c_int32, C.c_int64, _dp],
    "cs_group_rcm_step": [_vp, C.POINTER(_vp), C.c_double, C.c_int64],
    "cs_rcm_create": [_vp, C.c_int64, _dp, _dp, _dp, _dp, C.c_double, C.c_int64, _dp, C.c_int32, _dp, _dp, C.c_double, _dp, _dp,
                      C.c_double, C.c_int32, _dp, _dp, _dp, C.POINTER(_vp)],
    "cs_rcm_free": [_vp],
    "cs_rcm_step": [_vp, C.c_double, C.c_int64],
    "cs_rcm_set_temperature": [_vp, _dp],
    "cs_rcm_state": [_vp, _dp, _dp, _dp, _dp, _dp, _dp],
    "cs_rcm_info": [_vp, _i64p, _i64p, _i64p],
    "cs_rcm_ctx": [_vp, C.POINTER(_vp)],
    "cs_rcm_enqueue_fluxes": [_vp, _vp],
    "cs_rcm_enqueue_update": [_vp, _vp, C.c_double],
    "cs_rcm_flux_buffer": [_vp, C.POINTER(_vp)],
    "cs_rcm_peer_mailbox": [_vp, C.c_int32, C.POINTER(_vp), C.POINTER(C.c_int64)],
    "cs_rcm_peer_connect": [_vp, C.c_int32, C.c_int32, C.POINTER(_vp)],
    "cs_rcm_enqueue_step_peer": [_vp, C.c_double],
    "cs_rcm_peer_status": [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32)],
    "cs_ipc_export": [_vp, C.POINTER(C.c_uint8)],
    "cs_ipc_open": [_vp, C.POINTER(C.c_uint8), C.POINTER(_vp)],
    "cs_ipc_close": [_vp, _vp],
}

_lib = None


class ClearSkyError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[libclearsky_b200 status {code}] {msg}")
        self.code = code


def lib():
    """load the shared library (once); fail loudly if it has not been built"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python clearsky.jl_b200/build.py` "
                "(there is no CPU fallback for the CUDA path)")
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int32
        L.cs_last_error.restype = C.c_char_p
        L.cs_last_error.argtypes = []
        L.cs_version.restype = C.c_int32
        L.cs_version.argtypes = []
        _lib = L
    return _lib


def check(rc):
    if rc != CS_OK:
        msg = lib().cs_last_error().decode("utf-8", "replace")
        raise ClearSkyError(rc, msg)


def f64(a):
    """contiguous float64 numpy array (no copy when already so)"""
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    """double* of a contiguous float64 array, or NULL"""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def i64ptr(a):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_i64p)


class Context:
    """one CUDA device + stream (cs_ctx)"""

    def __init__(self, device=0, stream=None, borrowed=None):
        self.h = _vp()
        self._borrowed = borrowed is not None
        if borrowed is not None:
            self.h = borrowed          # owned by a cs_group
        elif stream is None:
            check(lib().cs_ctx_create(int(device), C.byref(self.h)))
        else:
            check(lib().cs_ctx_create_on_stream(int(device), _vp(int(stream)), C.byref(self.h)))
        self.device = int(device)

    def timers(self):
        t = np.zeros(CS_NTIMERS)
        check(lib().cs_ctx_timers(self.h, ptr(t)))
        return dict(zip(TIMER_NAMES, t.tolist()))

    def timers_total(self):
        """the same timers accumulated since the context was created (take differences around a region)"""
        t = np.zeros(CS_NTIMERS)
        check(lib().cs_ctx_timers_total(self.h, ptr(t)))
        return dict(zip(TIMER_NAMES, t.tolist()))

    def launches(self):
        n = C.c_int64(0)
        check(lib().cs_ctx_launches(self.h, C.byref(n)))
        return n.value

    def synchronize(self):
        check(lib().cs_ctx_synchronize(self.h))

    def set_farfield(self, mode):
        """far-wing treatment of the Voigt / Lorentz line sum: "direct" (every pair, like surf!) or "expansion"
        (local Taylor expansion of well-separated far-wing lines, truncation < 3e-11; see clearsky_b200.h)"""
        check(lib().cs_ctx_set_farfield(self.h, FARFIELD_MODES[mode] if isinstance(mode, str) else int(mode)))

    def set_tau_floor(self, τmin):
        """floor on the vertical layer optical depth in the flux kernel (reference: 1e-6, discretized.jl:174)"""
        check(lib().cs_ctx_set_tau_floor(self.h, float(τmin)))

    def get_farfield(self):
        m = C.c_int32(0)
        check(lib().cs_ctx_get_farfield(self.h, C.byref(m)))
        return {v: k for k, v in FARFIELD_MODES.items()}[m.value]

    def fp64_peak(self, iters=20000):
        v = C.c_double(0)
        check(lib().cs_fp64_peak(self.h, int(iters), C.byref(v)))
        return v.value

    def close(self):
        if self.h and not self._borrowed:
            lib().cs_ctx_free(self.h)
        self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=None):
    """process-wide context per device; device defaults to CLEARSKY_B200_DEVICE or LOCAL_RANK or 0"""
    if device is None:
        device = int(os.environ.get("CLEARSKY_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def device_count():
    n = C.c_int32(0)
    rc = lib().cs_device_count(C.byref(n))
    return n.value if rc == CS_OK else 0
