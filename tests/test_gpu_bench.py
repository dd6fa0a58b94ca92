"""GPU: bench.py's own line keeps the contract (short run)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_bench_line_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu",
                          "--frames-per-step", "512", "--e2e-frames", "128", "--compressed-frames", "48", "--compressed-passes", "1", "--decode-sessions", "2"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["steps"] == 5 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["dtype"] == "u8"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.2 < r["frac"] < 1.3
    assert d["gpu_launches"] >= 5 * 3
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] in (128 * 1658880, 128 * 288 * 1536) and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]
    assert d["value"] > 5e5
    # sustained timing: >= 10 segments of exactly K steps, value from the median segment
    t = d["timing"]
    assert t["segments"] >= 10 and t["steps_per_segment"] == 5 and t["segment_ms_min"] <= t["segment_ms_median"] <= t["segment_ms_max"]
    assert abs(d["ms_per_step"] - t["segment_ms_median"] / 5) < 1e-9
    # parity gate against the committed cv2 golden of the clip that was timed
    assert d["parity"]["checked"] is True and d["parity"]["bit_exact"] is True and d["parity"]["frames"] >= 512
    assert set(e["modes"]) == {"gather", "dma_rows", "pageable_gather", "pageable_rows"} and e["mode"] in e["modes"]
    assert e["value"] == max(m["value"] for m in e["modes"].values())
    c = d["e2e_compressed"]
    assert c["value"] > 0 and c["bit_exact_on_decoded_surface"] is True and c["sessions_per_gpu"] == 2 and c["file_frames"] == 48
    assert 0 < c["h2d_bytes_per_frame"] < 1920 * 1080 * 3 / 4
    # both arms print the same config object and honour steps / warm-up
    ref = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5", "--warmup", "3",
                          "--frames-per-step", "512", "--ref-sample", "6", "--ref-reps", "1"], capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0, ref.stderr[-2000:]
    r = json.loads([ln for ln in ref.stdout.splitlines() if ln.startswith("{")][-1])
    assert r["config"] == d["config"] and r["steps"] == d["steps"] and r["warmup"] == d["warmup"]
    assert r["metric"] == d["metric"] and r["unit"] == d["unit"] and r["higher_is_better"] == d["higher_is_better"]
