"""CPU suite: the JSON line bench.py prints keeps the driver's contract (checked on the reference arm, which
needs no GPU; the GPU arm's line is checked by tests/test_gpu_bench.py on the box)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def _run(args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=600, env={**os.environ, **(env or {})})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-sample", "6", "--ref-reps", "1", "--compressed-frames", "4",
              "--compressed-cpu-passes", "1"])
    assert d["e2e_compressed"]["value"] > 0 and d["e2e_compressed"]["file_frames"] == 4
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "frames/sec (1080p ContentDetector)" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["dtype"] == "u8" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_honours_steps_and_warmup_and_never_loads_the_product():
    """the reference arm must not clamp --steps/--warmup, must not load libesd.so (its frames come
    from the CPU twin of the clip generator) and prints the same `config` object as the GPU arm."""
    code = (
        "import sys, runpy, io, json, contextlib\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '4', '--warmup', '3', '--ref-sample', '6', '--ref-reps', '1', '--no-compressed']\n"
        "buf = io.StringIO()\n"
        "with contextlib.redirect_stdout(buf):\n"
        "    try:\n"
        "        runpy.run_path(%r, run_name='__main__')\n"
        "    except SystemExit as e:\n"
        "        assert not e.code, e.code\n"
        "assert not any(m == 'eioku_b200' or m.startswith('eioku_b200.') for m in sys.modules), 'reference arm imported the product'\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libesd.so' not in maps and 'libesd_synth.so' not in maps, 'reference arm mapped a product library'\n"
        "print(buf.getvalue())\n") % os.path.join(ROOT, "bench.py")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][0])
    assert d["steps"] == 4 and d["warmup"] == 3
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.bench_config(1, 2048)


def test_reference_arm_other_ranks_exit_quietly():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_on_cpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_a_failing_compressed_leg_does_not_cost_the_run_its_line(monkeypatch):
    """bench.compressed_leg: any failure inside the leg comes back as {"error": ...} (so the headline line is still printed), the
    ranks leave together (the failure flag goes through all_max), and only a parity mismatch -- SystemExit -- stays fatal."""
    import bench

    calls = []

    def all_max(x):
        calls.append(x)
        return x

    def boom(*a, **k):
        raise RuntimeError("no Motion-JPEG writer here")

    monkeypatch.setattr(bench, "_compressed_leg_body", boom)
    out = bench.compressed_leg(0, 0, 1, None, "cuda:0", None, lambda: None, all_max)
    assert "no Motion-JPEG writer here" in out["error"] and calls == [1.0]

    def late(local, rank, world, dist, dev, args, barrier, all_max_, agree):
        agree("set-up")
        agree("timed region", "decoder said no")
        raise AssertionError("not reached")

    calls.clear()
    monkeypatch.setattr(bench, "_compressed_leg_body", late)
    out = bench.compressed_leg(0, 0, 1, None, "cuda:0", None, lambda: None, all_max)
    assert out["error"].startswith("timed region: decoder said no") and calls == [0.0, 1.0]

    def other_rank_failed(local, rank, world, dist, dev, args, barrier, all_max_, agree):
        agree("set-up")
        return {"value": 1.0}

    monkeypatch.setattr(bench, "_compressed_leg_body", other_rank_failed)
    assert "another rank failed" in bench.compressed_leg(0, 0, 2, None, "cuda:0", None, lambda: None, lambda x: 1.0)["error"]
    assert bench.compressed_leg(0, 0, 1, None, "cuda:0", None, lambda: None, all_max) == {"value": 1.0}

    def gate(*a, **k):
        raise SystemExit("bench.py: PARITY GATE FAILED on the decoded surface")

    monkeypatch.setattr(bench, "_compressed_leg_body", gate)
    import pytest as _pytest

    with _pytest.raises(SystemExit):
        bench.compressed_leg(0, 0, 1, None, "cuda:0", None, lambda: None, all_max)
