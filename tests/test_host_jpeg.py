"""CPU: the arithmetic of libesd_decode.so's own baseline-JPEG decoder (eioku_b200/csrc/jpeg_core.h + jpeg_parse.h: Huffman
decode, libjpeg's ISLOW IDCT, fancy h2v2 upsampling, JFIF colour conversion) compiled with g++ and compared with
cv2.imdecode -- libjpeg-turbo with its defaults -- BIT FOR BIT.  The CUDA kernels include the same headers, so what runs on
the GPU is pinned to a real reference implementation here, without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = tmp_path_factory.mktemp("jpeg") / "jpeg_shim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), os.path.join(ROOT, "tests", "jpeg_shim.cpp")])
    L = C.CDLL(str(so))
    L.shim_jpeg_info.argtypes = [C.c_void_p, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p]
    L.shim_jpeg_decode.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_char_p, C.c_int]
    return L


def decode(L, jpg: bytes, flat: int = 0):
    buf = np.frombuffer(jpg, np.uint8)
    w, h, err = C.c_int(), C.c_int(), C.create_string_buffer(256)
    if L.shim_jpeg_info(buf.ctypes.data, buf.size, C.byref(w), C.byref(h), err) != 0:
        raise ValueError(err.value.decode())
    out = np.empty((h.value, w.value, 3), np.uint8)
    rc = L.shim_jpeg_decode(buf.ctypes.data, buf.size, out.ctypes.data, err, flat)
    if rc == -2:
        return None  # restart markers: the flat decoder declines
    assert rc == 0, err.value
    return out


def _images():
    rng = np.random.default_rng(5)
    yield "noise", rng.integers(0, 256, (72, 96, 3), dtype=np.uint8)
    g = np.zeros((128, 160, 3), np.uint8)
    g[..., 0] = np.arange(160)[None, :] * 255 // 159
    g[..., 1] = np.arange(128)[:, None] * 255 // 127
    g[..., 2] = 255 - g[..., 0]
    yield "gradient", g
    yield "odd_size", rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)          # not a multiple of 16 or of 2
    yield "tiny", rng.integers(0, 256, (1, 1, 3), dtype=np.uint8)
    yield "flat_saturated", np.full((40, 40, 3), (255, 0, 255), np.uint8)
    smooth = cv2.resize(rng.integers(0, 256, (9, 16, 3), dtype=np.uint8), (640, 360), interpolation=cv2.INTER_CUBIC)
    yield "smooth_360p", smooth
    yield "smooth_noise", np.clip(smooth.astype(int) + rng.integers(-20, 21, smooth.shape), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("quality", [100, 95, 75, 30, 5])
def test_decoder_arithmetic_is_bit_exact_with_cv2_imdecode(shim, quality):
    for name, img in _images():
        for extra in ([], [cv2.IMWRITE_JPEG_OPTIMIZE, 1], [cv2.IMWRITE_JPEG_RST_INTERVAL, 3]):
            ok, jpg = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420] + extra)
            assert ok
            want = cv2.imdecode(jpg, cv2.IMREAD_COLOR)
            got = decode(shim, jpg.tobytes())
            assert got.shape == want.shape, (name, quality)
            assert np.array_equal(got, want), (name, quality, extra, int(np.abs(got.astype(int) - want.astype(int)).max()))
            flat = decode(shim, jpg.tobytes(), flat=1)   # the GPU fast path's form: unstuffed scan, one flat loop
            if extra and extra[0] == cv2.IMWRITE_JPEG_RST_INTERVAL:
                assert flat is None   # a restart interval is declared: those pictures take the general decoder
            else:
                assert flat is not None and np.array_equal(flat, want), (name, quality, extra, "flat")


@pytest.mark.parametrize("quality", [100, 90, 60, 20])
def test_many_threads_per_picture_reach_the_sequential_decoding(shim, quality):
    """The self-synchronising scheme (jpeg_core.h: decode_span / decode_scan_parallel_host, what the GPU runs with a thread per
    sub-sequence): speculative starts, rounds until no start state changes, prefix sums, one writing pass.  The shim compares the
    coefficients with the flat loop's and the picture with cv2.imdecode; thread budgets from 2 to 1024 change the sub-sequence
    length (>= 512 bits), never the result."""
    shim.shim_last_rounds.restype = C.c_int
    rng = np.random.default_rng(quality)
    big = cv2.resize(rng.integers(0, 256, (27, 48, 3), dtype=np.uint8), (1920, 1080), interpolation=cv2.INTER_CUBIC)
    big = np.clip(big.astype(int) + rng.integers(-6, 7, big.shape), 0, 255).astype(np.uint8)
    worst = 0
    for name, img in list(_images()) + [("1080p", big)]:
        for extra in ([], [cv2.IMWRITE_JPEG_OPTIMIZE, 1]):
            ok, jpg = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420] + extra)
            want = cv2.imdecode(jpg, cv2.IMREAD_COLOR)
            for budget in ((2, 7, 64, 1024) if name != "1080p" else (1024, 300)):
                got = decode(shim, jpg.tobytes(), flat=budget)
                assert got is not None and np.array_equal(got, want), (name, quality, extra, budget)
                worst = max(worst, shim.shim_last_rounds())
                # and with the sparse hand-off to the IDCT (entry list + block offsets instead of the cleared dense array)
                got = decode(shim, jpg.tobytes(), flat=-budget)
                assert got is not None and np.array_equal(got, want), (name, quality, extra, budget, "sparse")
    # correct states spread at least one sub-sequence per round; in practice a handful of rounds, far below the thread count
    assert 1 <= worst <= 24, worst


def test_mjpeg_avi_pictures_written_by_cv2_decode_bit_exactly(shim, tmp_path):
    """The pictures ffmpeg's mjpeg encoder puts into an AVI (what cv2.VideoWriter('MJPG') produces and bench.py feeds)."""
    import struct

    rng = np.random.default_rng(9)
    frames = [cv2.resize(rng.integers(0, 256, (6, 8, 3), dtype=np.uint8), (320, 184), interpolation=cv2.INTER_CUBIC) for _ in range(4)]
    p = str(tmp_path / "c.avi")
    w = cv2.VideoWriter(p, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (320, 184))
    for f in frames:
        w.write(f)
    w.release()
    b = open(p, "rb").read()
    i = b.find(b"movi") + 4
    n = 0
    while i + 8 <= len(b) and b[i:i + 4] == b"00dc":
        sz = struct.unpack("<I", b[i + 4:i + 8])[0]
        jpg = b[i + 8:i + 8 + sz]
        want = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(decode(shim, jpg), want), n
        assert np.array_equal(decode(shim, jpg, flat=1), want), n
        i += 8 + sz + (sz & 1)
        n += 1
    assert n == 4


def test_unsupported_streams_are_rejected_not_misdecoded(shim):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (32, 32, 3), dtype=np.uint8)
    for params, what in (([cv2.IMWRITE_JPEG_PROGRESSIVE, 1], "progressive"),
                         ([cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444], "4:2:0"),
                         ([cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422], "4:2:0")):
        ok, jpg = cv2.imencode(".jpg", img, params)
        with pytest.raises(ValueError) as e:
            decode(shim, jpg.tobytes())
        assert what in str(e.value)
    ok, jpg = cv2.imencode(".jpg", img[..., 0])
    with pytest.raises(ValueError):
        decode(shim, jpg.tobytes())
    with pytest.raises(ValueError):
        decode(shim, b"\xff\xd8\xff\xd9")


def test_damaged_scans_never_leave_their_buffers(tmp_path):
    """tests/jpeg_fuzz.cpp under AddressSanitizer + UBSan: the decoding code the device runs (flat loop, decode_span in its three
    modes, expand_block), fed with bit-flipped / overwritten / truncated / random scans behind a real picture's tables, in buffers
    of exactly the library's sizes.  Out-of-bounds reads are what neither the parity tests nor the red-zone allocator can see."""
    exe = str(tmp_path / "jpeg_fuzz")
    build = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe,
                            os.path.join(ROOT, "tests", "jpeg_fuzz.cpp")], capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("this toolchain has no AddressSanitizer runtime")
    assert build.returncode == 0, build.stderr[-2000:]
    rng = np.random.default_rng(3)
    smooth = cv2.resize(rng.integers(0, 256, (9, 16, 3), dtype=np.uint8), (320, 184), interpolation=cv2.INTER_CUBIC)
    pics = {"smooth_q80_opt": (np.clip(smooth.astype(int) + rng.integers(-10, 11, smooth.shape), 0, 255).astype(np.uint8), 80, [cv2.IMWRITE_JPEG_OPTIMIZE, 1]),
            "noise_q100": (rng.integers(0, 256, (72, 96, 3), dtype=np.uint8), 100, []),
            "flat_q30": (np.full((40, 40, 3), (255, 0, 255), np.uint8), 30, [])}
    for k, (name, (img, q, extra)) in enumerate(pics.items()):
        p = str(tmp_path / f"{name}.jpg")
        assert cv2.imwrite(p, img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420] + extra)
        run = subprocess.run([exe, p, str(11 + k), "400"], capture_output=True, text=True, timeout=600)
        assert run.returncode == 0 and "fuzz ok: 400 damaged scans" in run.stdout, (name, run.stdout[-300:], run.stderr[-3000:])
