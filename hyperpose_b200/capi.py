"""ctypes binding of the C ABI declared in include/hyperpose_b200.h.

This is the only way Python (tests, bench.py, smoke) reaches the product: through the same
`extern "C"` entry points the C++ `hyperpose::parser::paf` / `hyperpose::dnn::tensorrt`
wrappers call.  The shared library must have been built in-tree (hyperpose_b200/build.py);
there is no fallback of any kind -- a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libhyperpose_b200.so")

N_PARTS, N_PAIRS = 18, 19
HP_OK, HP_ERR_ARG, HP_ERR_CAPACITY, HP_ERR_UNSUPPORTED, HP_ERR_CUDA, HP_ERR_BATCH = 0, -1, -2, -3, -4, -5

PART_DT = np.dtype({"names": ["has_value", "x", "y", "score"], "formats": ["<i4", "<f4", "<f4", "<f4"]})
HUMAN_DT = np.dtype([("parts", PART_DT, (N_PARTS,)), ("score", "<f4")])
PEAK_DT = np.dtype([("part_id", "<i4"), ("x", "<i4"), ("y", "<i4"), ("score", "<f4"), ("id", "<i4")])
CONN_DT = np.dtype([("cid1", "<i4"), ("cid2", "<i4"), ("score", "<f4")])
assert HUMAN_DT.itemsize == 292

# every symbol include/hyperpose_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "hp_last_error", "hp_device_count", "hp_version",
    "hp_paf_create", "hp_paf_destroy", "hp_paf_set_conf_thresh", "hp_paf_set_paf_thresh", "hp_paf_set_capacity",
    "hp_paf_process_host", "hp_paf_process_host_batched", "hp_paf_process_device", "hp_paf_fetch",
    "hp_paf_debug_peaks", "hp_paf_debug_connections", "hp_paf_launch_count",
]


class HyperposeError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"hyperpose_b200 status {status}: {msg}")
        self.status = status


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -m hyperpose_b200.build` "
                                    "(there is no CPU / PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.hp_last_error.restype = C.c_char_p
        L.hp_version.restype = C.c_char_p
        L.hp_device_count.restype = C.c_int
        L.hp_paf_create.argtypes = [C.POINTER(vp), C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
        L.hp_paf_destroy.argtypes = [vp]
        L.hp_paf_destroy.restype = None
        L.hp_paf_set_conf_thresh.argtypes = [vp, C.c_float]
        L.hp_paf_set_paf_thresh.argtypes = [vp, C.c_float]
        L.hp_paf_set_capacity.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.hp_paf_process_host.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, ip]
        L.hp_paf_process_host_batched.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, ip]
        L.hp_paf_process_device.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.hp_paf_fetch.argtypes = [vp, vp, C.c_int, ip, C.c_int]
        L.hp_paf_debug_peaks.argtypes = [vp, C.c_int, vp, C.c_int, ip]
        L.hp_paf_debug_connections.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, ip]
        L.hp_paf_launch_count.argtypes = [vp]
        L.hp_paf_launch_count.restype = C.c_longlong
        _lib = L
    return _lib


def check(rc):
    if rc != HP_OK:
        raise HyperposeError(rc, lib().hp_last_error().decode())


class PafParser:
    """Mirror of hyperpose::parser::paf (include/hyperpose/operator/parser/paf.hpp:17-93):
    same constructor arguments, process(conf, paf) -> humans, set_*_thresh."""

    def __init__(self, conf_thresh: float = 0.05, paf_thresh: float = 0.05, resolution_size=(-1, -1), device: int = 0):
        self._h = C.c_void_p()
        check(lib().hp_paf_create(C.byref(self._h), conf_thresh, paf_thresh, resolution_size[0], resolution_size[1], device))

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.hp_paf_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_conf_thresh(self, t):
        check(lib().hp_paf_set_conf_thresh(self._h, t))

    def set_paf_thresh(self, t):
        check(lib().hp_paf_set_paf_thresh(self._h, t))

    def set_capacity(self, peaks_per_part=0, candidates_per_limb=0, humans=0):
        check(lib().hp_paf_set_capacity(self._h, peaks_per_part, candidates_per_limb, humans))

    def process(self, conf: np.ndarray, paf: np.ndarray, cap: int = 128) -> np.ndarray:
        """One frame, host tensors conf[C,H,W], paf[2L,H,W] -> structured array of HUMAN_DT."""
        conf = np.ascontiguousarray(conf, np.float32)
        paf = np.ascontiguousarray(paf, np.float32)
        if conf.ndim != 3 or paf.ndim != 3:
            raise HyperposeError(HP_ERR_ARG, "Input of PAF::PROCESS didn't meet requirements: [conf, paf], tensor.dims() == 3")
        out = np.zeros(cap, HUMAN_DT)
        n = C.c_int(0)
        check(lib().hp_paf_process_host(self._h, conf.ctypes.data, paf.ctypes.data, conf.shape[0], paf.shape[0],
                                        conf.shape[1], conf.shape[2], out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    def process_batch(self, conf: np.ndarray, paf: np.ndarray, cap: int = 128):
        """N frames conf[N,C,H,W], paf[N,2L,H,W] -> list of N structured arrays."""
        conf = np.ascontiguousarray(conf, np.float32)
        paf = np.ascontiguousarray(paf, np.float32)
        N = conf.shape[0]
        out = np.zeros((N, cap), HUMAN_DT)
        n = (C.c_int * N)()
        check(lib().hp_paf_process_host_batched(self._h, conf.ctypes.data, paf.ctypes.data, N, conf.shape[1], paf.shape[1],
                                                conf.shape[2], conf.shape[3], out.ctypes.data, cap, n))
        return [out[i, :n[i]].copy() for i in range(N)]

    def process_device(self, d_conf_ptr: int, d_paf_ptr: int, N, c_conf, c_paf, H, W, stream: int = 0):
        check(lib().hp_paf_process_device(self._h, d_conf_ptr, d_paf_ptr, N, c_conf, c_paf, H, W, stream))

    def fetch(self, N: int, cap: int = 128):
        out = np.zeros((N, cap), HUMAN_DT)
        n = (C.c_int * N)()
        check(lib().hp_paf_fetch(self._h, out.ctypes.data, cap, n, N))
        return [out[i, :n[i]].copy() for i in range(N)]

    def debug_peaks(self, frame: int = 0, cap: int = 1 << 16) -> np.ndarray:
        out = np.zeros(cap, PEAK_DT)
        n = C.c_int(0)
        check(lib().hp_paf_debug_peaks(self._h, frame, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    def debug_connections(self, frame: int, pair_id: int, cap: int = 4096) -> np.ndarray:
        out = np.zeros(cap, CONN_DT)
        n = C.c_int(0)
        check(lib().hp_paf_debug_connections(self._h, frame, pair_id, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    @property
    def launch_count(self) -> int:
        return int(lib().hp_paf_launch_count(self._h))
