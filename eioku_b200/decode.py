"""Compressed video -> device-resident frames: ctypes binding of libesd_decode.so.

Stands where ``cv2.VideoCapture`` stands in the reference's decode loops
(/root/reference/ml-service/src/services/model_manager.py:237-263) and where the ``ffmpeg -i`` child of the shipped scene task
stands (:736-755): ``MjpegVideo(path)`` is a video source for :class:`eioku_b200.scene_manager.SceneManager` whose batches
are CUDA tensors decoded on the GPU -- by this library's own baseline-JPEG kernels (bit-identical to cv2.imdecode) or, for
streams those do not cover, by nvJPEG -- so the decoded frames never exist in host memory.  Motion-JPEG in AVI only (NVDEC is
closed to this container, include/esd_decode.h); anything else raises and the caller falls back to its host decoder.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libesd_decode.so")
ESD_DECODE_ABI_VERSION = 1
ESD_JPEG_AUTO, ESD_JPEG_DEFAULT, ESD_JPEG_GPU_HYBRID, ESD_JPEG_HARDWARE, ESD_JPEG_NATIVE = 0, 1, 2, 3, 4
# native = this library's own kernels (bit-identical to cv2.imdecode); the others are nvJPEG back ends
BACKEND_NAMES = {ESD_JPEG_DEFAULT: "default", ESD_JPEG_GPU_HYBRID: "gpu_hybrid", ESD_JPEG_HARDWARE: "hardware", ESD_JPEG_NATIVE: "native"}
EXPORTED_SYMBOLS = ("esd_decode_abi_version", "esd_mjpeg_last_error", "esd_mjpeg_open", "esd_mjpeg_get_info", "esd_mjpeg_seek",
                    "esd_mjpeg_read", "esd_mjpeg_close")


class MjpegInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fps_num", C.c_int32), ("fps_den", C.c_int32),
                ("n_frames", C.c_int64), ("compressed_bytes", C.c_int64), ("backend", C.c_int32), ("hw_engines", C.c_int32),
                ("batch_frames", C.c_int32), ("reserved", C.c_int32)]


class DecodeError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str = ""):
        self.status = status
        super().__init__(f"libesd_decode: {what} (status {status}){': ' + detail if detail else ''}")


_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python -m eioku_b200.build` (nvcc, sm_100a, -lnvjpeg)")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.esd_decode_abi_version.restype = C.c_int
        L.esd_mjpeg_last_error.restype = C.c_char_p
        L.esd_mjpeg_last_error.argtypes = [vp]
        L.esd_mjpeg_open.argtypes = [C.POINTER(vp), C.c_char_p, C.c_int, i32, i32]
        L.esd_mjpeg_get_info.argtypes = [vp, C.POINTER(MjpegInfo)]
        L.esd_mjpeg_seek.argtypes = [vp, i64]
        L.esd_mjpeg_read.argtypes = [vp, i64, vp, C.POINTER(vp), C.POINTER(i64)]
        L.esd_mjpeg_close.restype = None
        L.esd_mjpeg_close.argtypes = [vp]
        if L.esd_decode_abi_version() != ESD_DECODE_ABI_VERSION:
            raise ImportError("libesd_decode.so ABI version mismatch")
        _lib = L
    return _lib


class _DevicePtr:
    """CUDA array interface over library-owned device memory (no copy; the library keeps the buffer alive)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}


class MjpegVideo:
    """A Motion-JPEG AVI file decoded on the GPU, batch by batch.  Implements the video-source protocol of SceneManager
    (``frame_size``, ``frame_rate``, ``start_frame``, ``read_batch``); batches are uint8 CUDA tensors [n, H, W, 3] (BGR) that
    alias the decoder's double-buffered output -- valid until the second next ``read_batch``."""

    pixel_format = "bgr24"

    def __init__(self, path: str, device: int = 0, batch_frames: int = 64, backend: int = ESD_JPEG_AUTO,
                 first_frame: int = 0, end_frame: Optional[int] = None):
        self._L = load_library()
        self._h = C.c_void_p()
        rc = self._L.esd_mjpeg_open(C.byref(self._h), os.fsencode(path), int(device), int(batch_frames), int(backend))
        if rc != 0:
            detail = (self._L.esd_mjpeg_last_error(None) or b"").decode()
            self._h = C.c_void_p()
            raise DecodeError(rc, "esd_mjpeg_open", detail)
        self.device = int(device)
        info = MjpegInfo()
        self._check(self._L.esd_mjpeg_get_info(self._h, C.byref(info)), "esd_mjpeg_get_info")
        self.info = info
        self.n_frames = int(info.n_frames)
        self.frame_size: Tuple[int, int] = (int(info.width), int(info.height))
        self.frame_rate = float(info.fps_num) / float(max(1, info.fps_den))
        self.backend = BACKEND_NAMES.get(int(info.backend), "?")
        # a frame range [first_frame, end_frame): frame-range sharding decodes each shard's load range independently
        self.start_frame = int(first_frame)
        self._end = self.n_frames if end_frame is None else min(int(end_frame), self.n_frames)
        if first_frame:
            self.seek(first_frame)
        self._pos = int(first_frame)

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise DecodeError(rc, what, (self._L.esd_mjpeg_last_error(self._h) or b"").decode())

    def seek(self, frame: int):
        self._check(self._L.esd_mjpeg_seek(self._h, int(frame)), "esd_mjpeg_seek")
        self._pos = int(frame)

    def read_batch(self, n: int = 0):
        """Next frames as a CUDA tensor (decode enqueued on torch's current stream of the device), None at the end."""
        import torch

        left = self._end - self._pos
        if left <= 0:
            return None
        want = left if n <= 0 else min(left, int(n))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ptr, got = C.c_void_p(), C.c_int64()
        self._check(self._L.esd_mjpeg_read(self._h, want, C.c_void_p(stream), C.byref(ptr), C.byref(got)), "esd_mjpeg_read")
        if got.value == 0:
            return None
        self._pos += int(got.value)
        w, h = self.frame_size
        return torch.as_tensor(_DevicePtr(ptr.value, (int(got.value), h, w, 3)), device=f"cuda:{self.device}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.esd_mjpeg_close(self._h)
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


class TeeVideo:
    """Wraps a GPU video source and keeps a device-side copy of every batch it hands out -- the decoded surface, for callers
    that need the frames afterwards (and for parity checks: the oracle runs on exactly what was scored)."""

    def __init__(self, src):
        self.src, self.kept = src, []
        self.frame_size, self.frame_rate = src.frame_size, src.frame_rate
        self.start_frame, self.pixel_format = src.start_frame, src.pixel_format

    def read_batch(self, n: int = 0):
        b = self.src.read_batch(n)
        if b is not None:
            self.kept.append(b.clone())  # enqueued on the stream the decode ran on, before the ring slot is reused
        return b

    def frames(self):
        import torch

        return torch.cat(self.kept)


def is_mjpeg_avi(path: str) -> bool:
    """Cheap sniff: RIFF/AVI whose first video stream header names a Motion-JPEG fourcc."""
    try:
        with open(path, "rb") as f:
            head = f.read(4096)
    except OSError:
        return False
    if head[:4] != b"RIFF" or head[8:12] != b"AVI ":
        return False
    i = head.find(b"strh")
    return i >= 0 and head[i + 8:i + 12] == b"vids" and head[i + 12:i + 16].upper() in (b"MJPG", b"AVI1", b"JPEG", b"IJPG")


class SeekError(RuntimeError):
    """A capture did not land on the frame it was asked to seek to (or a range came up short)."""


class CaptureRangeVideo:
    """Frames [first_frame, end_frame) of a file through a cv2.VideoCapture of its own: the reference's host decode loop
    (`cap = cv2.VideoCapture(path)` ... `cap.read()`, model_manager.py:237-263) for every codec OpenCV's FFmpeg build reads,
    restricted to a frame range so that SEVERAL of them can decode one file at once (service: `decode_workers`; one capture per
    frame-range shard, all shards scored on one GPU and merged by the usual global decision pass).  Frames are decoded straight
    into one of two reused batch buffers by a thread of the source's own, so decoding the next batch overlaps whatever the caller
    does with this one (cv2 and libesd both release the GIL); a batch stays valid until the next read_batch.  Host batches go
    through the ingest ring, whose host-side copy of pageable memory is complete when the push returns.

    A seek that does not land exactly breaks parity silently, so it is checked twice: the capture must report the requested
    position, and the frames listed in `watch` (the first frames of the other ranges) are fingerprinted as they pass, so that the
    caller can compare every range's first frame with the same frame as decoded by the range that owns it."""

    pixel_format = "bgr24"
    frames_stable = False

    def __init__(self, path: str, first_frame: int = 0, end_frame: Optional[int] = None, batch_frames: int = 64, watch=(),
                 threads: int = 0, prefetch: bool = True, until_eof: bool = False, frame_count: Optional[int] = None):
        """until_eof: ignore the container's frame count and read until the decoder stops (whole-file sources: some containers
        under-report CAP_PROP_FRAME_COUNT).  frame_count: the frame count the caller works with instead of the capture's own claim
        (all ranges of one job must agree on it).  A range that ends at that count tries to grab one more frame when it gets
        there and reports it in `trailing_frames` -- a container that under-reports its length must not lose its tail silently."""
        import queue
        import threading

        import cv2
        import numpy as np

        self._cv2, self._np = cv2, np
        if threads > 0 and hasattr(cv2, "CAP_PROP_N_THREADS"):
            # several captures at once: FFmpeg's own frame threads would otherwise take every core for each of them
            self._cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG, [cv2.CAP_PROP_N_THREADS, int(threads)])
            if not self._cap.isOpened():
                self._cap = cv2.VideoCapture(path)
        else:
            self._cap = cv2.VideoCapture(path)
        if not self._cap.isOpened():
            raise RuntimeError(f"Failed to open video: {path}")
        self.n_frames = int(self._cap.get(cv2.CAP_PROP_FRAME_COUNT)) if frame_count is None else int(frame_count)
        self.trailing_frames = None   # set once the claimed end of the file is reached: did the decoder have more?
        self._until_eof = bool(until_eof)
        self.frame_size = (int(self._cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(self._cap.get(cv2.CAP_PROP_FRAME_HEIGHT)))
        self.frame_rate = float(self._cap.get(cv2.CAP_PROP_FPS) or 30.0)
        self.start_frame = int(first_frame)
        self._end = (1 << 62) if until_eof else (self.n_frames if end_frame is None else min(int(end_frame), self.n_frames))
        self._pos = self.start_frame
        if self.start_frame:
            self._cap.set(cv2.CAP_PROP_POS_FRAMES, self.start_frame)
            got = int(self._cap.get(cv2.CAP_PROP_POS_FRAMES))
            if got != self.start_frame:
                self._cap.release()
                raise SeekError(f"{path}: seek to frame {self.start_frame} landed on {got}")
        self._batch = max(1, int(batch_frames))
        self._watch = {int(f) for f in watch} | {self.start_frame}
        self.digests = {}   # frame index -> fingerprint, for the watched frames this range decoded
        shape = (self._batch, self.frame_size[1], self.frame_size[0], 3)
        self._held = None
        self._thread = None
        if prefetch:
            self._free, self._ready = queue.Queue(), queue.Queue()
            for _ in range(2):
                self._free.put(np.empty(shape, np.uint8))
            self._thread = threading.Thread(target=self._pump, name="esd-capture", daemon=True)
            self._thread.start()
        else:
            self._buf = np.empty(shape, np.uint8)

    @staticmethod
    def _fingerprint(frame) -> int:
        import zlib

        return zlib.crc32(frame[::4].tobytes())   # every fourth row: 1.5 MB of a 1080p frame

    def _decode_into(self, buf) -> int:
        k = min(self._batch, self._end - self._pos)
        if k <= 0 and not self._until_eof and self._end >= self.n_frames and self.trailing_frames is None:
            self.trailing_frames = bool(self._cap.grab())
        got = 0
        for i in range(max(0, k)):
            ok, frame = self._cap.read(buf[i])
            if not ok:
                break
            if frame is not buf[i] and not self._np.shares_memory(frame, buf[i]):
                buf[i][...] = frame
            if self._pos + i in self._watch:
                self.digests[self._pos + i] = self._fingerprint(buf[i])
            got += 1
        self._pos += got
        return got

    def _pump(self):
        try:
            while True:
                buf = self._free.get()
                if buf is None:
                    return
                got = self._decode_into(buf)
                self._ready.put((buf, got))
                if got == 0:
                    return
        except BaseException as e:  # noqa: BLE001 - handed to the reading thread
            self._ready.put((e, -1))

    def read_batch(self, n: int = 0):
        if self._thread is None:
            got = self._decode_into(self._buf)
            return self._buf[:got] if got else None
        if self._held is not None:        # the caller is done with the previous batch
            self._free.put(self._held)
            self._held = None
        if self._ready is None:
            return None
        buf, got = self._ready.get()
        if got < 0:
            self._ready = None
            raise buf
        if got == 0:
            self._ready = None
            return None
        self._held = buf
        return buf[:got]

    def close(self):
        if self._thread is not None:
            self._free.put(None)
            self._thread.join(timeout=30)
            self._thread = None
        if self._cap is not None:
            self._cap.release()
            self._cap = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
