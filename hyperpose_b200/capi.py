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
    "hp_paf_debug_peaks", "hp_paf_debug_connections", "hp_paf_launch_count", "hp_paf_copy_results_device", "hp_paf_debug_timing",
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
        L.hp_paf_copy_results_device.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
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

    def debug_timing(self, N: int):
        """HPB_PAF_TIMING=1: (cta[N,19,4] ns stamps: start, ordered, candidates, matched; asm[N,6]: assembly start, end | path in the low
        two bits, then -- component-parallel path only -- staged, labelled, grouped, lanes done)"""
        out = np.zeros(N * (N_PAIRS * 4 + 6), np.uint64)
        lib().hp_paf_debug_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        check(lib().hp_paf_debug_timing(self._h, out.ctypes.data, N))
        return out[:N * N_PAIRS * 4].reshape(N, N_PAIRS, 4), out[N * N_PAIRS * 4:].reshape(N, 6)

    def copy_results_device(self, d_humans_ptr: int, d_counts_ptr: int, N: int, cap: int, stream: int = 0):
        check(lib().hp_paf_copy_results_device(self._h, d_humans_ptr, d_counts_ptr, N, cap, stream))


def handoff_stats() -> dict:
    """counters of the device-resident engine -> parser hand-off (csrc/handoff.h)"""
    L = lib()
    if not getattr(L, "_engine_bound", False):
        _bind_engine(L)
        L._engine_bound = True
    v = [C.c_longlong() for _ in range(4)]
    check(L.hp_handoff_stats(*[C.byref(x) for x in v]))
    return dict(zip(("published", "hits", "batch_parses", "misses"), [x.value for x in v]))


def handoff_enable(on: bool):
    L = lib()
    if not getattr(L, "_engine_bound", False):
        _bind_engine(L)
        L._engine_bound = True
    check(L.hp_handoff_enable(1 if on else 0))


# ---------------------------------------------------------------------------------------------
# DNN engine
# ---------------------------------------------------------------------------------------------
EXPORTS += [
    "hp_engine_create", "hp_engine_destroy", "hp_engine_info", "hp_engine_infer_u8_host", "hp_engine_infer_u8_device",
    "hp_engine_infer_f32_host", "hp_engine_outputs", "hp_engine_read_outputs_host", "hp_engine_copy_outputs_device", "hp_engine_sync",
    "hp_engine_launch_count", "hp_engine_debug_read_buffer", "hp_engine_debug_write_buffer", "hp_engine_debug_run_ops",
    "hp_pose_run_u8_host", "hp_engine_stage_frame_u8", "hp_engine_infer_staged", "hp_engine_debug_read_frames", "hp_engine_set_output_override", "hp_engine_set_profiling", "hp_engine_get_profile",
    "hp_engine_read_outputs_frames", "hp_engine_head_type", "hp_handoff_enable", "hp_handoff_stats",
    "hp_pose_submit_u8_host", "hp_pose_collect", "hp_pose_stats", "hp_paf_prepare", "hp_paf_state", "hp_paf_copy_results_host_async",
    "hp_paf_grow_capacity", "hp_pool_create", "hp_pool_destroy", "hp_pool_size", "hp_pool_set_capacity", "hp_pool_run_u8_host",
    "hp_pool_set_output_override", "hp_pool_launch_count", "hp_default_device", "hp_handoff_device_of",
    "hp_engine_create_ex", "hp_engine_dtype", "hp_pose_submit_u8_device",
]


def _bind_engine(L):
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    L.hp_engine_create.argtypes = [C.POINTER(vp), vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
    L.hp_engine_create_ex.argtypes = [C.POINTER(vp), vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
    L.hp_engine_dtype.argtypes = [vp]
    L.hp_engine_destroy.argtypes = [vp]
    L.hp_engine_destroy.restype = None
    L.hp_engine_info.argtypes = [vp, ip, ip, ip, ip, ip, ip, ip, C.POINTER(C.c_double)]
    L.hp_engine_infer_u8_host.argtypes = [vp, vp, C.c_int]
    L.hp_engine_infer_u8_device.argtypes = [vp, vp, C.c_int, vp]
    L.hp_engine_infer_f32_host.argtypes = [vp, vp, C.c_int]
    L.hp_engine_outputs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.hp_engine_read_outputs_host.argtypes = [vp, vp, vp, C.c_int]
    L.hp_engine_sync.argtypes = [vp]
    L.hp_engine_copy_outputs_device.argtypes = [vp, vp, vp, C.c_int, vp]
    L.hp_engine_launch_count.argtypes = [vp]
    L.hp_engine_launch_count.restype = C.c_longlong
    L.hp_engine_debug_read_buffer.argtypes = [vp, C.c_int, vp, C.c_int, ip, ip, ip]
    L.hp_engine_debug_write_buffer.argtypes = [vp, C.c_int, vp, C.c_int]
    L.hp_engine_debug_run_ops.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.hp_pose_run_u8_host.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, ip]
    L.hp_engine_stage_frame_u8.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int]
    L.hp_engine_infer_staged.argtypes = [vp, C.c_int]
    L.hp_engine_debug_read_frames.argtypes = [vp, vp, C.c_int]
    L.hp_engine_set_output_override.argtypes = [vp, vp, vp]
    L.hp_engine_set_profiling.argtypes = [vp, C.c_int]
    L.hp_engine_get_profile.argtypes = [vp, vp, vp, vp, C.c_int, ip, C.POINTER(C.c_longlong)]
    L.hp_engine_read_outputs_frames.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.c_int, C.c_int]
    L.hp_engine_head_type.argtypes = [vp]
    L.hp_handoff_enable.argtypes = [C.c_int]
    L.hp_handoff_stats.argtypes = [C.POINTER(C.c_longlong)] * 4
    L.hp_pose_submit_u8_host.argtypes = [vp, vp, vp, C.c_int, ip]
    L.hp_pose_submit_u8_device.argtypes = [vp, vp, vp, C.c_int, ip]
    L.hp_pose_collect.argtypes = [vp, C.c_int, vp, C.c_int, ip]
    L.hp_pose_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.hp_pool_create.argtypes = [C.POINTER(vp), ip, C.c_int, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_float, C.c_float]
    L.hp_pool_destroy.argtypes = [vp]
    L.hp_pool_destroy.restype = None
    L.hp_pool_size.argtypes = [vp]
    L.hp_pool_set_capacity.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.hp_pool_run_u8_host.argtypes = [vp, vp, C.c_int, vp, C.c_int, ip]
    L.hp_pool_set_output_override.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.hp_pool_launch_count.argtypes = [vp]
    L.hp_pool_launch_count.restype = C.c_longlong
    L.hp_default_device.restype = C.c_int
    L.hp_handoff_device_of.argtypes = [vp]


class Engine:
    """Mirror of hyperpose::dnn::tensorrt (include/hyperpose/operator/dnn/tensorrt.hpp:33-141):
    Engine(model_pack, input_size=(w, h), max_batch_size, factor=1/255, flip_rgb=True); inference(frames)."""

    def __init__(self, pack: bytes, input_size, max_batch_size: int = 8, factor: float = 1.0 / 255, flip_rgb: bool = True,
                 device: int = 0, dtype: str = "f16"):
        """dtype: "f16" (= data_type::kHALF) or "tf32" (= data_type::kFLOAT of the reference ctor, tensorrt.hpp:14-22)"""
        L = lib()
        if not getattr(L, "_engine_bound", False):
            _bind_engine(L)
            L._engine_bound = True
        self._h = C.c_void_p()
        self._pack = pack
        self.dtype = dtype
        check(L.hp_engine_create_ex(C.byref(self._h), pack, len(pack), int(input_size[0]), int(input_size[1]), max_batch_size,
                                    factor, 1 if flip_rgb else 0, device, {"f16": 0, "tf32": 1}[dtype]))
        v = [C.c_int() for _ in range(7)]
        fl = C.c_double()
        check(L.hp_engine_info(self._h, *[C.byref(x) for x in v], C.byref(fl)))
        (self.in_w, self.in_h, self.max_batch, self.c_conf, self.c_paf, self.out_h, self.out_w) = [x.value for x in v]
        self.flops_per_frame = fl.value

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.hp_engine_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def max_batch_size(self):
        return self.max_batch

    def input_size(self):
        return (self.in_w, self.in_h)

    def infer_u8(self, frames: np.ndarray):
        """frames u8[N,in_h,in_w,3] (BGR, already network-sized).  Asynchronous; outputs stay on the device."""
        frames = np.ascontiguousarray(frames, np.uint8)
        assert frames.ndim == 4 and frames.shape[1:] == (self.in_h, self.in_w, 3), frames.shape
        check(lib().hp_engine_infer_u8_host(self._h, frames.ctypes.data, frames.shape[0]))
        self._last_n = frames.shape[0]

    def stage_frame(self, slot: int, frame: np.ndarray, keep_ratio: bool = False):
        """one u8 HWC3 frame of ANY size -> GPU resize (cv::resize / non_scaling_resize) into batch slot `slot`"""
        frame = np.ascontiguousarray(frame, np.uint8)
        assert frame.ndim == 3 and frame.shape[2] == 3
        check(lib().hp_engine_stage_frame_u8(self._h, slot, frame.ctypes.data, frame.shape[0], frame.shape[1], 1 if keep_ratio else 0))

    def infer_staged(self, n: int):
        check(lib().hp_engine_infer_staged(self._h, n))
        self._last_n = n

    def debug_read_frames(self, n: int) -> np.ndarray:
        out = np.empty((n, self.in_h, self.in_w, 3), np.uint8)
        check(lib().hp_engine_debug_read_frames(self._h, out.ctypes.data, n))
        return out

    def infer_u8_device(self, d_ptr: int, n: int, stream: int = 0):
        check(lib().hp_engine_infer_u8_device(self._h, d_ptr, n, stream))
        self._last_n = n

    def infer_f32(self, nchw: np.ndarray):
        nchw = np.ascontiguousarray(nchw, np.float32)
        check(lib().hp_engine_infer_f32_host(self._h, nchw.ctypes.data, nchw.shape[0]))
        self._last_n = nchw.shape[0]

    def inference(self, frames: np.ndarray):
        """tensorrt::inference(std::vector<cv::Mat>): returns per image [conf[C,h,w], paf[2L,h,w]] host tensors
        (ordered by name: conf < paf, src/tensorrt.cpp:405)."""
        self.infer_u8(frames)
        conf, paf = self.read_outputs(frames.shape[0])
        return [[conf[i], paf[i]] for i in range(frames.shape[0])]

    def read_outputs(self, n: int):
        conf = np.empty((n, self.c_conf, self.out_h, self.out_w), np.float32)
        paf = np.empty((n, self.c_paf, self.out_h, self.out_w), np.float32)
        check(lib().hp_engine_read_outputs_host(self._h, conf.ctypes.data, paf.ctypes.data, n))
        return conf, paf

    def read_outputs_frames(self, n: int, publish: bool = False):
        """tensorrt::inference's return value: per image its own host buffers [conf_i, paf_i] (the feature_map_t storage,
        src/tensorrt.cpp:398-431).  publish=True (what the C++ drop-in does, where feature_map_t is read-only) registers them
        for the device-resident hand-off (handoff.h); a look-up compares every byte, so editing the arrays afterwards is safe."""
        if self.head_type == 1:
            sa, sb = (17, 5, self.out_h, self.out_w), (19, 9, self.out_h, self.out_w)
        else:
            sa, sb = (self.c_conf, self.out_h, self.out_w), (self.c_paf, self.out_h, self.out_w)
        a = [np.empty(sa, np.float32) for _ in range(n)]
        b = [np.empty(sb, np.float32) for _ in range(n)]
        pa = (C.c_void_p * n)(*[x.ctypes.data for x in a])
        pb = (C.c_void_p * n)(*[x.ctypes.data for x in b])
        check(lib().hp_engine_read_outputs_frames(self._h, pa, pb, n, 1 if publish else 0))
        return [[a[i], b[i]] for i in range(n)]

    @property
    def head_type(self) -> int:
        return int(lib().hp_engine_head_type(self._h))

    def device_outputs(self):
        a, b, s = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().hp_engine_outputs(self._h, C.byref(a), C.byref(b), C.byref(s)))
        return a.value, b.value, (s.value or 0)

    def sync(self):
        check(lib().hp_engine_sync(self._h))

    def copy_outputs_device(self, d_conf_ptr: int, d_paf_ptr: int, n: int, stream: int = 0):
        check(lib().hp_engine_copy_outputs_device(self._h, d_conf_ptr, d_paf_ptr, n, stream))

    @property
    def launch_count(self) -> int:
        return int(lib().hp_engine_launch_count(self._h))

    def debug_read_buffer(self, buf: int, n: int) -> np.ndarray:
        H, W, Cc = C.c_int(), C.c_int(), C.c_int()
        check(lib().hp_engine_debug_read_buffer(self._h, buf, None, n, C.byref(H), C.byref(W), C.byref(Cc)))
        out = np.empty((n, H.value, W.value, Cc.value), np.float32 if self.dtype == "tf32" else np.float16)
        check(lib().hp_engine_debug_read_buffer(self._h, buf, out.ctypes.data, n, C.byref(H), C.byref(W), C.byref(Cc)))
        return out

    def debug_write_buffer(self, buf: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr, np.float32 if self.dtype == "tf32" else np.float16)
        check(lib().hp_engine_debug_write_buffer(self._h, buf, arr.ctypes.data, arr.shape[0]))

    def debug_run_ops(self, first: int, last: int, n: int):
        check(lib().hp_engine_debug_run_ops(self._h, first, last, n))

    def set_output_override(self, d_conf_ptr: int, d_paf_ptr: int):
        check(lib().hp_engine_set_output_override(self._h, d_conf_ptr, d_paf_ptr))

    def set_profiling(self, enable: bool):
        check(lib().hp_engine_set_profiling(self._h, 1 if enable else 0))

    def get_profile(self):
        """-> (ms_per_op[n], op_type[n], flops_per_frame_per_op[n], runs)"""
        cap = 1024
        ms = np.zeros(cap, np.float64); ty = np.zeros(cap, np.int32); fl = np.zeros(cap, np.float64)
        n, runs = C.c_int(), C.c_longlong()
        check(lib().hp_engine_get_profile(self._h, ms.ctypes.data, ty.ctypes.data, fl.ctypes.data, cap, C.byref(n), C.byref(runs)))
        return ms[:n.value], ty[:n.value], fl[:n.value], runs.value

    def run_pose(self, parser: "PafParser", frames: np.ndarray, cap: int = 128):
        """hp_pose_run_u8_host: frames in, humans out (list of N structured arrays)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        N = frames.shape[0]
        out = np.zeros((N, cap), HUMAN_DT)
        n = (C.c_int * N)()
        check(lib().hp_pose_run_u8_host(self._h, parser._h, frames.ctypes.data, N, out.ctypes.data, cap, n))
        return [out[i, :n[i]].copy() for i in range(N)]


    def submit_pose(self, parser: "PafParser", frames: np.ndarray) -> int:
        """hp_pose_submit_u8_host: enqueue one batch (H2D on the copy stream, graph replay, record D2H); returns the ticket.
        `frames` must stay alive until collect_pose when it is page-locked memory (DMA reads it directly)."""
        assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"]
        t = C.c_int(-1)
        if isinstance(parser, PifPafParser):   # OpenPifPaf pack: decoder on its own stream underneath the next batch's convs
            lib().hp_pose_submit_pifpaf_u8_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
            check(lib().hp_pose_submit_pifpaf_u8_host(self._h, parser._h, frames.ctypes.data, frames.shape[0], C.byref(t)))
        else:
            check(lib().hp_pose_submit_u8_host(self._h, parser._h, frames.ctypes.data, frames.shape[0], C.byref(t)))
        self._ticket_n = getattr(self, "_ticket_n", {})
        self._ticket_n[t.value] = frames.shape[0]
        return t.value

    def submit_pose_device(self, parser: "PafParser", d_frames_ptr: int, n: int) -> int:
        """hp_pose_submit_u8_device: the frames are already in device memory"""
        t = C.c_int(-1)
        if isinstance(parser, PifPafParser):
            lib().hp_pose_submit_pifpaf_u8_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
            check(lib().hp_pose_submit_pifpaf_u8_device(self._h, parser._h, d_frames_ptr, n, C.byref(t)))
        else:
            check(lib().hp_pose_submit_u8_device(self._h, parser._h, d_frames_ptr, n, C.byref(t)))
        self._ticket_n = getattr(self, "_ticket_n", {})
        self._ticket_n[t.value] = n
        return t.value

    def collect_pose(self, ticket: int, cap: int = 128):
        N = self._ticket_n[ticket]
        out = np.zeros((N, cap), HUMAN_DT)
        n = (C.c_int * N)()
        check(lib().hp_pose_collect(self._h, ticket, out.ctypes.data, cap, n))
        return [out[i, :n[i]].copy() for i in range(N)]

    def pose_stats(self):
        a, b = C.c_longlong(), C.c_longlong()
        check(lib().hp_pose_stats(self._h, C.byref(a), C.byref(b)))
        return {"graph_captures": a.value, "graph_launches": b.value}


class Pool:
    """hp_pool_*: one engine + parser + host thread per GPU inside this process; frames shard in blocks of max_batch
    (SURVEY 8e), humans come back in frame order."""

    def __init__(self, pack: bytes, input_size, max_batch_size: int, devices=None, factor: float = 1.0 / 255, flip_rgb: bool = True,
                 conf_thresh: float = 0.05, paf_thresh: float = 0.05):
        L = lib()
        if not getattr(L, "_engine_bound", False):
            _bind_engine(L)
            L._engine_bound = True
        n = len(devices) if devices is not None else L.hp_device_count()
        devs = (C.c_int * n)(*(devices if devices is not None else range(n)))
        self._h = C.c_void_p()
        check(L.hp_pool_create(C.byref(self._h), devs, n, pack, len(pack), int(input_size[0]), int(input_size[1]), max_batch_size,
                               factor, 1 if flip_rgb else 0, conf_thresh, paf_thresh))
        self.n_gpus, self.in_w, self.in_h, self.max_batch = n, int(input_size[0]), int(input_size[1]), max_batch_size

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.hp_pool_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_capacity(self, peaks_per_part=0, candidates_per_limb=0, humans=0):
        check(lib().hp_pool_set_capacity(self._h, peaks_per_part, candidates_per_limb, humans))

    def set_output_override(self, d_conf_ptrs, d_paf_ptrs):
        n = self.n_gpus
        a = (C.c_void_p * n)(*d_conf_ptrs)
        b = (C.c_void_p * n)(*d_paf_ptrs)
        check(lib().hp_pool_set_output_override(self._h, a, b))

    def run(self, frames: np.ndarray, cap: int = 128):
        """frames u8[N_total, in_h, in_w, 3] (any N_total) -> list of N_total HUMAN_DT arrays, frame order"""
        assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"] and frames.shape[1:] == (self.in_h, self.in_w, 3)
        N = frames.shape[0]
        out = np.zeros((N, cap), HUMAN_DT)
        n = (C.c_int * N)()
        check(lib().hp_pool_run_u8_host(self._h, frames.ctypes.data, N, out.ctypes.data, cap, n))
        return [out[i, :n[i]].copy() for i in range(N)]

    @property
    def launch_count(self) -> int:
        return int(lib().hp_pool_launch_count(self._h))


# ---------------------------------------------------------------------------------------------
# OpenPifPaf decoder
# ---------------------------------------------------------------------------------------------
EXPORTS += ["hp_pifpaf_create", "hp_pifpaf_destroy", "hp_pifpaf_process_host", "hp_pifpaf_process_device", "hp_pifpaf_fetch",
            "hp_pifpaf_launch_count", "hp_pifpaf_debug_counts", "hp_pifpaf_debug_hr",
            "hp_pifpaf_pipeline_info", "hp_pifpaf_copy_results_host_async", "hp_pifpaf_grow_capacity",
            "hp_pose_submit_pifpaf_u8_host", "hp_pose_submit_pifpaf_u8_device"]


class PifPafParser:
    """Mirror of hyperpose::parser::pifpaf (include/hyperpose/operator/parser/pifpaf.hpp:8-26): PifPafParser(h, w, thresh)."""

    def __init__(self, net_h: int, net_w: int, thresh: float = 0.1, device: int = 0):
        L = lib()
        vp, ip = C.c_void_p, C.POINTER(C.c_int)
        L.hp_pifpaf_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_float, C.c_int]
        L.hp_pifpaf_destroy.argtypes = [vp]
        L.hp_pifpaf_destroy.restype = None
        L.hp_pifpaf_process_host.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, ip]
        L.hp_pifpaf_process_device.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
        L.hp_pifpaf_fetch.argtypes = [vp, vp, C.c_int, ip, C.c_int]
        self._h = C.c_void_p()
        check(L.hp_pifpaf_create(C.byref(self._h), net_h, net_w, thresh, device))

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.hp_pifpaf_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process_batch(self, pif: np.ndarray, paf: np.ndarray, cap: int = 128):
        """pif[N,17,5,h,w], paf[N,19,9,h,w] host tensors -> list of N HUMAN_DT arrays"""
        pif = np.ascontiguousarray(pif, np.float32)
        paf = np.ascontiguousarray(paf, np.float32)
        N, _, _, h, w = pif.shape
        out = np.zeros((N, cap), HUMAN_DT)
        n = (C.c_int * N)()
        check(lib().hp_pifpaf_process_host(self._h, pif.ctypes.data, paf.ctypes.data, N, h, w, out.ctypes.data, cap, n))
        return [out[i, :n[i]].copy() for i in range(N)]

    def process(self, pif: np.ndarray, paf: np.ndarray, cap: int = 128):
        return self.process_batch(pif[None], paf[None], cap)[0]

    def process_device(self, d_pif_ptr: int, d_paf_ptr: int, N: int, h: int, w: int, stream: int = 0):
        check(lib().hp_pifpaf_process_device(self._h, d_pif_ptr, d_paf_ptr, N, h, w, stream))

    def fetch(self, N: int, cap: int = 128):
        out = np.zeros((N, cap), HUMAN_DT)
        n = (C.c_int * N)()
        check(lib().hp_pifpaf_fetch(self._h, out.ctypes.data, cap, n, N))
        return [out[i, :n[i]].copy() for i in range(N)]

    @property
    def launch_count(self) -> int:
        lib().hp_pifpaf_launch_count.argtypes = [C.c_void_p]
        lib().hp_pifpaf_launch_count.restype = C.c_longlong
        return int(lib().hp_pifpaf_launch_count(self._h))

    def debug_hr(self, frame: int, field: int, h: int, w: int) -> np.ndarray:
        out = np.zeros(((h - 1) * 8 + 1, (w - 1) * 8 + 1), np.float32)
        lib().hp_pifpaf_debug_hr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        check(lib().hp_pifpaf_debug_hr(self._h, frame, field, out.ctypes.data))
        return out

    def debug_counts(self, frame: int = 0):
        out = (C.c_int * 7)()
        lib().hp_pifpaf_debug_counts.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        check(lib().hp_pifpaf_debug_counts(self._h, frame, out))
        return dict(zip(["seeds", "anns", "kept", "flags", "nms_h", "nms_w", "caf_entries"], list(out)))


EXPORTS += ["hp_ppn_create", "hp_ppn_destroy", "hp_ppn_set_point_thresh", "hp_ppn_set_limb_thresh", "hp_ppn_set_nms_thresh",
            "hp_ppn_process_host", "hp_ppn_process_device", "hp_ppn_fetch", "hp_ppn_launch_count"]


class PoseProposalParser:
    """Mirror of hyperpose::parser::pose_proposal (include/hyperpose/operator/parser/proposal_network.hpp:18-80):
    PoseProposalParser((net_w, net_h), point_thresh=0.10, limb_thresh=0.05, nms_thresh=0.3)."""

    def __init__(self, net_resolution, point_thresh: float = 0.10, limb_thresh: float = 0.05, nms_thresh: float = 0.3, device: int = 0):
        L = lib()
        vp, ip, ci, cf = C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_float
        L.hp_ppn_create.argtypes = [C.POINTER(vp), ci, ci, cf, cf, cf, ci]
        L.hp_ppn_destroy.argtypes = [vp]
        L.hp_ppn_destroy.restype = None
        for f in (L.hp_ppn_set_point_thresh, L.hp_ppn_set_limb_thresh, L.hp_ppn_set_nms_thresh):
            f.argtypes = [vp, cf]
        L.hp_ppn_process_host.argtypes = [vp] + [vp] * 6 + [ci] * 7 + [vp, ci, ip]
        L.hp_ppn_process_device.argtypes = [vp] + [vp] * 6 + [ci] * 7 + [vp]
        L.hp_ppn_fetch.argtypes = [vp, vp, ci, ip, ci]
        L.hp_ppn_launch_count.argtypes = [vp]
        L.hp_ppn_launch_count.restype = C.c_longlong
        self._h = C.c_void_p()
        check(L.hp_ppn_create(C.byref(self._h), int(net_resolution[0]), int(net_resolution[1]), point_thresh, limb_thresh, nms_thresh, device))

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.hp_ppn_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_point_thresh(self, t: float):
        check(lib().hp_ppn_set_point_thresh(self._h, t))

    def set_limb_thresh(self, t: float):
        check(lib().hp_ppn_set_limb_thresh(self._h, t))

    def set_nms_thresh(self, t: float):
        check(lib().hp_ppn_set_nms_thresh(self._h, t))

    def process_batch(self, conf_point, x, y, w, h, edge, cap: int = 256):
        """host tensors [N,K,gh,gw] x5 and edge [N,E,nh,nw,gh,gw] -> list of N HUMAN_DT arrays"""
        a = [np.ascontiguousarray(t, np.float32) for t in (conf_point, x, y, w, h, edge)]
        N, K, gh, gw = a[0].shape
        E, nh, nw = a[5].shape[1:4]
        out = np.zeros((N, cap), HUMAN_DT)
        n = (C.c_int * N)()
        check(lib().hp_ppn_process_host(self._h, *[t.ctypes.data for t in a], N, K, gh, gw, E, nh, nw, out.ctypes.data, cap, n))
        return [out[i, :n[i]].copy() for i in range(N)]

    def process(self, conf_point, conf_iou, x, y, w, h, edge, cap: int = 256):
        """pose_proposal::process(conf_point, conf_iou, x, y, w, h, edge): one frame; conf_iou is ignored like in the
        reference (src/pose_proposal.cpp:74) apart from its leading dimension"""
        K = np.asarray(conf_iou).shape[0]
        return self.process_batch(*[np.asarray(t)[None, :K] for t in (conf_point, x, y, w, h)], np.asarray(edge)[None], cap=cap)[0]

    @property
    def launch_count(self) -> int:
        return int(lib().hp_ppn_launch_count(self._h))
