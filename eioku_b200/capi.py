"""ctypes binding of libesd.so (include/esd.h).  Fails loudly when the CUDA library is
missing or no GPU is present -- there is no CPU fallback in this package."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESD_LIB") or os.path.join(_HERE, "libesd.so")  # ESD_LIB: tuning experiments only

ESD_DET_CONTENT, ESD_DET_ADAPTIVE, ESD_DET_HIST, ESD_DET_THRESHOLD, ESD_DET_HASH = 1, 2, 4, 8, 16
ESD_ABI_VERSION = 4
ESD_THRESH_FLOOR, ESD_THRESH_CEILING = 0, 1
ESD_FILTER_MERGE, ESD_FILTER_SUPPRESS = 0, 1
ESD_DOWNSCALE_FLOAT, ESD_DOWNSCALE_INT = 0, 1
ESD_SPLIT_AUTO, ESD_SPLIT_STRIPS, ESD_SPLIT_CHUNKS = 0, 1, 2
ESD_FMT_BGR24, ESD_FMT_NV12, ESD_FMT_I420 = 0, 1, 2
ESD_SCORE_CONTENT_VAL, ESD_SCORE_ADAPTIVE_VAL, ESD_SCORE_HIST_DIFF, ESD_SCORE_AVERAGE_RGB, ESD_SCORE_HASH_DIST, ESD_SCORE_ADAPTIVE_RATIO = range(6)
# the score array each detector's decision pass consumes (esd_decide_arrays / esd_decide_device)
DECISION_SCORE_KIND = {ESD_DET_CONTENT: ESD_SCORE_CONTENT_VAL, ESD_DET_ADAPTIVE: ESD_SCORE_ADAPTIVE_VAL, ESD_DET_HIST: ESD_SCORE_HIST_DIFF,
                       ESD_DET_THRESHOLD: ESD_SCORE_AVERAGE_RGB, ESD_DET_HASH: ESD_SCORE_HASH_DIST}

# every symbol include/esd.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = (
    "esd_abi_version", "esd_strerror", "esd_last_error", "esd_device_count", "esd_config_default",
    "esd_create", "esd_destroy", "esd_reset", "esd_get_geometry", "esd_get_touched_rows",
    "esd_push_frames", "esd_push_nv12", "esd_push_i420", "esd_push_rows", "esd_ingest_open", "esd_ingest_push_host", "esd_ingest_close", "esd_ingest_set_gather",
    "esd_ingest_stats", "esd_ingest_wait_copied", "esd_decide_device", "esd_copy_scores_device", "esd_synchronize", "esd_join", "esd_frames_pushed", "esd_read_scores", "esd_read_edge_counts", "esd_read_average_rgb",
    "esd_read_hash", "esd_read_hash_margin", "esd_debug_read_hash_input", "esd_process_frame_host",
    "esd_post_process", "esd_get_cuts",
    "esd_decide_arrays", "esd_debug_read_prev", "esd_debug_guard_selftest", "esd_set_timing", "esd_kernel_time", "esd_kernel_launches",
)


class EsdConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("detectors", C.c_int32),
        ("src_width", C.c_int32), ("src_height", C.c_int32),
        ("dst_width", C.c_int32), ("dst_height", C.c_int32),
        ("downscale_mode", C.c_int32), ("src_format", C.c_int32),
        ("content_threshold", C.c_double), ("content_weights", C.c_double * 4),
        ("content_weight_div", C.c_double),
        ("content_min_scene_len", C.c_int32), ("content_filter_mode", C.c_int32),
        ("adaptive_threshold", C.c_double), ("adaptive_min_content_val", C.c_double),
        ("adaptive_weights", C.c_double * 4), ("adaptive_weight_div", C.c_double),
        ("adaptive_window_width", C.c_int32), ("adaptive_min_scene_len", C.c_int32),
        ("hist_threshold", C.c_double), ("hist_bins", C.c_int32), ("hist_min_scene_len", C.c_int32),
        ("thresh_threshold", C.c_double), ("thresh_fade_bias", C.c_double),
        ("thresh_min_scene_len", C.c_int32), ("thresh_add_final_scene", C.c_int32),
        ("thresh_method", C.c_int32), ("edge_kernel_size", C.c_int32),
        ("rows_per_group", C.c_int32), ("pipeline_stages", C.c_int32),
        ("split_mode", C.c_int32), ("ctas_per_sm", C.c_int32),
        ("rows_per_stage", C.c_int32), ("reserved1", C.c_int32),
        ("max_cuts", C.c_int64), ("initial_capacity", C.c_int64),
        ("hash_threshold", C.c_double), ("hash_size", C.c_int32), ("hash_lowpass", C.c_int32),
        ("hash_min_scene_len", C.c_int32), ("reserved2", C.c_int32),
    ]


class EsdGeometry(C.Structure):
    _fields_ = [
        ("dst_width", C.c_int32), ("dst_height", C.c_int32),
        ("n_touched_rows", C.c_int32), ("row_bytes", C.c_int32),
        ("alg_bytes_per_frame", C.c_int64), ("compact_frame_bytes", C.c_int64),
        ("lane_stride", C.c_int32), ("lane_stride_taps", C.c_int32),
    ]


class EsdError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str = ""):
        self.status = status
        super().__init__(f"libesd: {what} (status {status}){': ' + detail if detail else ''}")


_lib = None


def load_library(path: Optional[str] = None):
    """dlopen libesd.so and declare prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -m eioku_b200.build` (nvcc, sm_100a). "
            "eioku_b200 has no CPU fallback.")
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.esd_abi_version.restype = C.c_int
    L.esd_strerror.restype = C.c_char_p
    L.esd_strerror.argtypes = [C.c_int]
    L.esd_last_error.restype = C.c_char_p
    L.esd_last_error.argtypes = [vp]
    L.esd_device_count.restype = C.c_int
    L.esd_config_default.restype = None
    L.esd_config_default.argtypes = [C.POINTER(EsdConfig)]
    L.esd_create.argtypes = [C.POINTER(vp), C.POINTER(EsdConfig), C.c_int]
    L.esd_destroy.restype = None
    L.esd_destroy.argtypes = [vp]
    L.esd_reset.argtypes = [vp]
    L.esd_get_geometry.argtypes = [vp, C.POINTER(EsdGeometry)]
    L.esd_get_touched_rows.argtypes = [vp, vp, i32]
    L.esd_push_frames.argtypes = [vp, vp, i64, i64, i64, i64, vp]
    L.esd_push_rows.argtypes = [vp, vp, i64, i64, vp]
    L.esd_push_nv12.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp]
    L.esd_push_i420.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, vp]
    L.esd_ingest_open.argtypes = [vp, i32, i32]
    L.esd_ingest_push_host.argtypes = [vp, vp, i64, i64, i64, i64]
    L.esd_ingest_close.argtypes = [vp]
    L.esd_ingest_set_gather.argtypes = [vp, i32]
    L.esd_ingest_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.esd_synchronize.argtypes = [vp]
    L.esd_join.argtypes = [vp, vp]
    L.esd_frames_pushed.restype = i64
    L.esd_frames_pushed.argtypes = [vp]
    L.esd_read_scores.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp]
    L.esd_get_cuts.argtypes = [vp, i32, i64, vp, i64, C.POINTER(i64), C.POINTER(i64)]
    L.esd_read_average_rgb.argtypes = [vp, i64, i64, vp]
    L.esd_read_edge_counts.argtypes = [vp, i64, i64, vp]
    L.esd_read_hash.argtypes = [vp, i64, i64, vp, vp]
    L.esd_read_hash_margin.argtypes = [vp, i64, i64, vp]
    L.esd_process_frame_host.argtypes = [vp, vp, i64, i64, i32, i64, vp, i64, C.POINTER(i64), C.POINTER(i64)]
    L.esd_debug_read_hash_input.argtypes = [vp, i64, vp, i64]
    L.esd_post_process.argtypes = [vp, i32, i64, vp, i64, C.POINTER(i64)]
    L.esd_decide_arrays.argtypes = [vp, i32, i64, i64, vp, vp, vp, i64, C.POINTER(i64)]
    L.esd_decide_device.argtypes = [vp, i32, i64, i64, vp, vp, vp, i64, C.POINTER(i64), vp]
    L.esd_copy_scores_device.argtypes = [vp, i32, i64, i64, vp, i32, vp]
    L.esd_ingest_wait_copied.argtypes = [vp]
    L.esd_debug_read_prev.argtypes = [vp, vp, i64]
    L.esd_debug_guard_selftest.argtypes = [C.c_int32]
    L.esd_set_timing.argtypes = [vp, i32]
    L.esd_kernel_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    L.esd_kernel_launches.restype = i64
    L.esd_kernel_launches.argtypes = [vp]
    if L.esd_abi_version() != ESD_ABI_VERSION:
        raise ImportError("libesd.so ABI version mismatch")
    if path == LIB_PATH:
        _lib = L
    return L


def default_config() -> EsdConfig:
    cfg = EsdConfig()
    load_library().esd_config_default(C.byref(cfg))
    return cfg


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EsdContext:
    """Owns one esd_ctx (one device, one video stream)."""

    def __init__(self, cfg: EsdConfig, device: int = 0):
        self._L = load_library()
        self._h = C.c_void_p()
        cfg.struct_size = C.sizeof(EsdConfig)
        rc = self._L.esd_create(C.byref(self._h), C.byref(cfg), int(device))
        if rc != 0:
            detail = (self._L.esd_last_error(None) or b"").decode()
            self._h = C.c_void_p()
            if rc == -1 and "window_width" in detail:
                raise ValueError(detail)
            raise EsdError(rc, "esd_create", detail)
        self.cfg = cfg
        self.device = int(device)
        g = EsdGeometry()
        self._check(self._L.esd_get_geometry(self._h, C.byref(g)), "esd_get_geometry")
        self.geometry = g
        self._pf_buf = np.empty(64, np.int64)

    # -- plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            detail = (self._L.esd_last_error(self._h) or b"").decode()
            raise EsdError(rc, what, detail)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.esd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- geometry
    @property
    def dst_size(self):
        return self.geometry.dst_width, self.geometry.dst_height

    @property
    def alg_bytes_per_frame(self) -> int:
        return int(self.geometry.alg_bytes_per_frame)

    def touched_rows(self) -> np.ndarray:
        rows = np.empty(self.geometry.n_touched_rows, np.int32)
        self._check(self._L.esd_get_touched_rows(self._h, _np_ptr(rows), rows.size), "esd_get_touched_rows")
        return rows

    # -- scoring
    def reset(self):
        self._check(self._L.esd_reset(self._h), "esd_reset")

    def push_device(self, data_ptr: int, n: int, frame_stride: int, pitch: int, first_frame_num: int, stream: int = 0):
        self._check(self._L.esd_push_frames(self._h, C.c_void_p(data_ptr), n, frame_stride, pitch, first_frame_num,
                                            C.c_void_p(stream)), "esd_push_frames")

    def push_tensor(self, frames, first_frame_num: int, stream: Optional[int] = None):
        """frames: torch.uint8 CUDA tensor [N,H,W,3] (or [H,W,3]); rows must be dense (stride 3,1 on W,C)."""
        import torch

        if frames.dim() == 3:
            frames = frames.unsqueeze(0)
        if frames.dtype != torch.uint8 or not frames.is_cuda:
            raise ValueError("push_tensor needs a CUDA uint8 tensor")
        if frames.shape[3] != 3 or frames.stride(3) != 1 or frames.stride(2) != 3:
            frames = frames.contiguous()
        if frames.device.index != self.device:
            raise ValueError(f"tensor on cuda:{frames.device.index}, context on cuda:{self.device}")
        n, h, w, _ = frames.shape
        if (w, h) != (self.cfg.src_width, self.cfg.src_height):
            raise ValueError(f"frame size {w}x{h} != configured {self.cfg.src_width}x{self.cfg.src_height}")
        if stream is None:
            stream = torch.cuda.current_stream(frames.device).cuda_stream
        fs = frames.stride(0) if n > 1 else frames.stride(1) * h
        self.push_device(frames.data_ptr(), n, fs, frames.stride(1), first_frame_num, stream)

    def push_nv12_device(self, y_ptr: int, uv_ptr: int, n: int, frame_stride: int, pitch: int, first_frame_num: int, stream: int = 0):
        self._check(self._L.esd_push_nv12(self._h, C.c_void_p(y_ptr), C.c_void_p(uv_ptr), n, frame_stride, pitch, first_frame_num,
                                          C.c_void_p(stream)), "esd_push_nv12")

    def push_i420_device(self, y_ptr: int, u_ptr: int, v_ptr: int, n: int, frame_stride: int, pitch_y: int, pitch_uv: int,
                         first_frame_num: int, stream: int = 0):
        """Planar YUV 4:2:0 frames on the device, the three planes of frame 0 given separately (an AVFrame's data[0..2])."""
        self._check(self._L.esd_push_i420(self._h, C.c_void_p(y_ptr), C.c_void_p(u_ptr), C.c_void_p(v_ptr), n, frame_stride, pitch_y,
                                          pitch_uv, first_frame_num, C.c_void_p(stream)), "esd_push_i420")

    def _check_nv12_shape(self, n_rows: int, w: int):
        W, H = self.cfg.src_width, self.cfg.src_height
        if self.cfg.src_format not in (ESD_FMT_NV12, ESD_FMT_I420):
            raise ValueError("context was not created with src_format = ESD_FMT_NV12 / ESD_FMT_I420")
        if (w, n_rows) != (W, H * 3 // 2):
            raise ValueError(f"YUV 4:2:0 frames must be [N, {H * 3 // 2}, {W}] (Y plane, then the chroma plane(s)), got [.., {n_rows}, {w}]")

    def push_nv12_tensor(self, frames, first_frame_num: int, stream: Optional[int] = None):
        """frames: torch.uint8 CUDA tensor [N, H*3/2, W] -- contiguous NV12 frames (rows dense, any row pitch), or, on an I420
        context, contiguous I420 frames (the array cv2.cvtColor(COLOR_YUV2BGR_I420) takes: rows must be W bytes apart)."""
        import torch

        if frames.dim() == 2:
            frames = frames.unsqueeze(0)
        if frames.dtype != torch.uint8 or not frames.is_cuda:
            raise ValueError("push_nv12_tensor needs a CUDA uint8 tensor")
        if frames.stride(2) != 1 or (self.cfg.src_format == ESD_FMT_I420 and frames.stride(1) != frames.shape[2]):
            frames = frames.contiguous()
        n, rows, w = frames.shape
        self._check_nv12_shape(rows, w)
        if stream is None:
            stream = torch.cuda.current_stream(frames.device).cuda_stream
        fs = frames.stride(0) if n > 1 else frames.stride(1) * rows
        self.push_device(frames.data_ptr(), n, fs, frames.stride(1), first_frame_num, stream)

    def ingest_push_nv12_numpy(self, frames: np.ndarray, first_frame_num: int):
        """Host NV12 frames [N, H*3/2, W] through the ingest ring (touched Y and UV rows only cross PCIe)."""
        if frames.ndim == 2:
            frames = frames[None]
        if frames.dtype != np.uint8 or frames.strides[2] != 1 or (self.cfg.src_format == ESD_FMT_I420 and frames.strides[1] != frames.shape[2]):
            frames = np.ascontiguousarray(frames, np.uint8)
        n, rows, w = frames.shape
        self._check_nv12_shape(rows, w)
        fs = frames.strides[0] if n > 1 else frames.strides[1] * rows
        self.ingest_push_host(frames.ctypes.data, n, fs, frames.strides[1], first_frame_num)

    def push_rows_device(self, data_ptr: int, n: int, first_frame_num: int, stream: int = 0):
        self._check(self._L.esd_push_rows(self._h, C.c_void_p(data_ptr), n, first_frame_num, C.c_void_p(stream)),
                    "esd_push_rows")

    # -- ingest ring (host frames)
    def ingest_open(self, n_slots: int = 3, frames_per_slot: int = 64):
        self._check(self._L.esd_ingest_open(self._h, n_slots, frames_per_slot), "esd_ingest_open")

    def ingest_push_host(self, data_ptr: int, n: int, frame_stride: int, pitch: int, first_frame_num: int):
        self._check(self._L.esd_ingest_push_host(self._h, C.c_void_p(data_ptr), n, frame_stride, pitch,
                                                 first_frame_num), "esd_ingest_push_host")

    def ingest_push_numpy(self, frames: np.ndarray, first_frame_num: int):
        if frames.ndim == 3:
            frames = frames[None]
        if frames.dtype != np.uint8 or frames.shape[3] != 3 or frames.strides[3] != 1 or frames.strides[2] != 3:
            frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w, _ = frames.shape
        if (w, h) != (self.cfg.src_width, self.cfg.src_height):
            raise ValueError(f"frame size {w}x{h} != configured {self.cfg.src_width}x{self.cfg.src_height}")
        fs = frames.strides[0] if n > 1 else frames.strides[1] * h
        self.ingest_push_host(frames.ctypes.data, n, fs, frames.strides[1], first_frame_num)

    def ingest_set_gather(self, n_threads: int):
        """n_threads > 0: host threads gather only the tap bytes of touched rows before the H2D copy."""
        self._check(self._L.esd_ingest_set_gather(self._h, int(n_threads)), "esd_ingest_set_gather")

    def ingest_close(self):
        self._check(self._L.esd_ingest_close(self._h), "esd_ingest_close")

    def ingest_stats(self):
        b, c = C.c_int64(), C.c_int64()
        self._check(self._L.esd_ingest_stats(self._h, C.byref(b), C.byref(c)), "esd_ingest_stats")
        return int(b.value), int(c.value)

    # -- results
    def synchronize(self):
        self._check(self._L.esd_synchronize(self._h), "esd_synchronize")

    def join(self, stream: int = 0):
        """Make `stream` wait for the finalize/decision tails enqueued so far (device-side)."""
        self._check(self._L.esd_join(self._h, C.c_void_p(stream)), "esd_join")

    @property
    def frames_pushed(self) -> int:
        return int(self._L.esd_frames_pushed(self._h))

    def read_scores(self, from_frame: int, n: int, want: Sequence[str] = ()):
        """-> dict with any of sums3, content_val, adaptive_val, adaptive_ratio, hist, hist_diff."""
        has_content = bool(self.cfg.detectors & (ESD_DET_CONTENT | ESD_DET_ADAPTIVE | ESD_DET_THRESHOLD | ESD_DET_HASH))
        has_hist = bool(self.cfg.detectors & ESD_DET_HIST)
        if not want:
            want = (["sums3", "content_val", "adaptive_val", "adaptive_ratio"] if has_content else []) + \
                   (["hist", "hist_diff"] if has_hist else [])
        out = {}
        if "sums3" in want: out["sums3"] = np.empty((n, 3), np.uint64)
        if "content_val" in want: out["content_val"] = np.empty(n, np.float64)
        if "adaptive_val" in want: out["adaptive_val"] = np.empty(n, np.float64)
        if "adaptive_ratio" in want: out["adaptive_ratio"] = np.empty(n, np.float64)
        if "hist" in want: out["hist"] = np.empty((n, self.cfg.hist_bins), np.uint32)
        if "hist_diff" in want: out["hist_diff"] = np.empty(n, np.float64)
        self._check(self._L.esd_read_scores(
            self._h, from_frame, n, _np_ptr(out.get("sums3")), _np_ptr(out.get("content_val")),
            _np_ptr(out.get("adaptive_val")), _np_ptr(out.get("adaptive_ratio")), _np_ptr(out.get("hist")),
            _np_ptr(out.get("hist_diff"))), "esd_read_scores")
        return out

    def read_edge_counts(self, from_frame: int, n: int) -> np.ndarray:
        out = np.empty(n, np.uint32)
        self._check(self._L.esd_read_edge_counts(self._h, from_frame, n, _np_ptr(out)), "esd_read_edge_counts")
        return out

    def read_average_rgb(self, from_frame: int, n: int) -> np.ndarray:
        out = np.empty(n, np.float64)
        self._check(self._L.esd_read_average_rgb(self._h, from_frame, n, _np_ptr(out)), "esd_read_average_rgb")
        return out

    def read_hash(self, from_frame: int, n: int):
        """-> (bits bool [n, size, size], hash_dist float64 [n]; NaN for the first frame of the video)."""
        size = int(self.cfg.hash_size)
        words = (size * size + 31) // 32
        raw = np.empty((n, words), np.uint32)
        dist = np.empty(n, np.float64)
        self._check(self._L.esd_read_hash(self._h, from_frame, n, _np_ptr(raw), _np_ptr(dist)), "esd_read_hash")
        bits = np.unpackbits(raw.view(np.uint8).reshape(n, words * 4), axis=1, bitorder="little")[:, :size * size]
        return bits.reshape(n, size, size).astype(bool), dist

    def read_hash_margin(self, from_frame: int, n: int) -> np.ndarray:
        """float32 [n]: per frame the smallest |DCT coefficient - median| (a hash bit is only defined above cv2.dct's noise, ~4e-6)."""
        out = np.empty(n, np.float32)
        self._check(self._L.esd_read_hash_margin(self._h, from_frame, n, _np_ptr(out)), "esd_read_hash_margin")
        return out

    def debug_hash_input(self, frame: int) -> np.ndarray:
        """uint8 [S, S] INTER_AREA thumbnail the hash of `frame` was computed from (most recent push only; test hook)."""
        s = int(self.cfg.hash_size) * int(self.cfg.hash_lowpass)
        out = np.empty((s, s), np.uint8)
        self._check(self._L.esd_debug_read_hash_input(self._h, frame, _np_ptr(out), out.size), "esd_debug_read_hash_input")
        return out

    def post_process(self, detector: int, last_frame_num: int):
        buf = np.empty(4, np.int64)
        nc = C.c_int64()
        self._check(self._L.esd_post_process(self._h, detector, last_frame_num, _np_ptr(buf), buf.size, C.byref(nc)),
                    "esd_post_process")
        return buf[: nc.value].tolist()

    def get_cuts(self, detector: int, from_index: int = 0):
        """-> (list of cut frame numbers from `from_index` on, total cuts emitted)."""
        cap = 4096
        while True:
            buf = np.empty(cap, np.int64)
            nw, nt = C.c_int64(), C.c_int64()
            rc = self._L.esd_get_cuts(self._h, detector, from_index, _np_ptr(buf), cap, C.byref(nw), C.byref(nt))
            if rc == -6 and nt.value - from_index > cap:  # ESD_ERR_CAPACITY: retry with a larger buffer
                cap = int(nt.value - from_index)
                continue
            self._check(rc, "esd_get_cuts")
            return buf[: nw.value].tolist(), int(nt.value)

    def process_frame_host(self, frame: np.ndarray, frame_num: int, detector: int, from_index: int = 0):
        """One host frame [H,W,3] uint8 (dense pixels, any row pitch) through the per-frame fast path.
        -> (cuts of `detector` from `from_index` on, total cuts emitted)."""
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3 or frame.strides[2] != 1 or frame.strides[1] != 3:
            frame = np.ascontiguousarray(frame, np.uint8)
        h, w, _ = frame.shape
        if (w, h) != (self.cfg.src_width, self.cfg.src_height):
            raise ValueError(f"frame size {w}x{h} != configured {self.cfg.src_width}x{self.cfg.src_height}")
        buf = self._pf_buf
        nw, nt = C.c_int64(), C.c_int64()
        rc = self._L.esd_process_frame_host(self._h, frame.ctypes.data, frame.strides[0], frame_num, detector, from_index,
                                            buf.ctypes.data, buf.size, C.byref(nw), C.byref(nt))
        if rc == -6 and nt.value - from_index > buf.size:  # more cuts pending than the small buffer holds
            return self.get_cuts(detector, from_index)
        self._check(rc, "esd_process_frame_host")
        return buf[: nw.value].tolist(), int(nt.value)

    def decide_arrays(self, detector: int, first_frame_num: int, scores: np.ndarray):
        scores = np.ascontiguousarray(scores, np.float64)
        n = scores.size
        ratio = np.empty(n, np.float64)
        cap = max(16, n)
        cuts = np.empty(cap, np.int64)
        nc = C.c_int64()
        self._check(self._L.esd_decide_arrays(self._h, detector, first_frame_num, n, _np_ptr(scores), _np_ptr(ratio),
                                              _np_ptr(cuts), cap, C.byref(nc)), "esd_decide_arrays")
        return cuts[: nc.value].tolist(), ratio

    def decide_device(self, detector: int, first_frame_num: int, d_scores_ptr: int, n: int, stream: int = 0, d_ratio_ptr: int = 0):
        """Decision pass over float64 scores already resident on this ctx's device -> list of cut frame numbers."""
        cap = max(16, min(int(n), int(self.cfg.max_cuts) if self.cfg.max_cuts > 0 else 65536))
        cuts = np.empty(cap, np.int64)
        nc = C.c_int64()
        self._check(self._L.esd_decide_device(self._h, detector, first_frame_num, n, C.c_void_p(d_scores_ptr),
                                              C.c_void_p(d_ratio_ptr) if d_ratio_ptr else None, _np_ptr(cuts), cap, C.byref(nc),
                                              C.c_void_p(stream)), "esd_decide_device")
        return cuts[: nc.value].tolist()

    def copy_scores_device(self, kind: int, from_frame: int, n: int, dst_ptr: int, dst_device: int = -1, stream: int = 0):
        """Async device-to-device (peer) copy of one per-frame score array into `dst_ptr` on `dst_device`."""
        self._check(self._L.esd_copy_scores_device(self._h, kind, from_frame, n, C.c_void_p(dst_ptr), int(dst_device),
                                                   C.c_void_p(stream)), "esd_copy_scores_device")

    def ingest_wait_copied(self):
        """Wait until the H2D copies of every ingest push so far have read the caller's (pinned) frames."""
        self._check(self._L.esd_ingest_wait_copied(self._h), "esd_ingest_wait_copied")

    def debug_last_hsv(self) -> np.ndarray:
        """uint8 [dst_h, dst_w, 3] HSV of the last pushed frame at detector resolution (test hook)."""
        w, h = self.dst_size
        buf = np.empty(w * h, np.uint32)
        self._check(self._L.esd_debug_read_prev(self._h, _np_ptr(buf), buf.size), "esd_debug_read_prev")
        out = np.empty((h, w, 3), np.uint8)
        b = buf.reshape(h, w)
        out[..., 0] = b & 255
        out[..., 1] = (b >> 8) & 255
        out[..., 2] = (b >> 16) & 255
        return out

    # -- instrumentation
    def set_timing(self, enable: bool = True):
        self._check(self._L.esd_set_timing(self._h, 1 if enable else 0), "esd_set_timing")

    def kernel_time(self):
        ms, n = C.c_double(), C.c_int64()
        self._check(self._L.esd_kernel_time(self._h, C.byref(ms), C.byref(n)), "esd_kernel_time")
        return float(ms.value), int(n.value)

    @property
    def kernel_launches(self) -> int:
        return int(self._L.esd_kernel_launches(self._h))
