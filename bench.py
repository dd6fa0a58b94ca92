#!/usr/bin/env python
"""bench.py -- frames/s of the scene-scoring hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours:      one process per GPU (torchrun for N > 1), device-resident synthetic 1080p frames of the
           config-2 clip (ContentDetector(27, 15)), one step = one 2048-frame batch through
           esd_push_frames (fused TMA kernel + finalize + decision).  No collective on the data path:
           ranks score independent frame ranges (weak scaling).  The timed region is a run of SEGMENTS of
           exactly K steps each, back to back, long enough (>= 1 s, >= 10 segments) that the number is the
           sustained one; `value` comes from the median segment (max over ranks per segment).  After the timed
           region every rank checks its resident frames bit-for-bit against the committed cv2 golden (parity
           gate: a mismatch exits non-zero and prints no number).  Prints ONE JSON line on rank 0.
reference: the reference's CPU path (PySceneDetect logic on cv2, one process per host core) on a
           bounded sample of the same workload.  Loads nothing of eioku_b200.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# the product sets the same default on import (eioku_b200/__init__.py explains it); here as well so that it holds whichever
# module touches CUDA first
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

W, H, FPS = 1920, 1080, 30
DST = (256, 144)  # PySceneDetect >= 0.6.2 auto-downscale of a 1920-wide frame (factor 7.5)
TOUCHED_ROW_BYTES = 288 * W * 3  # source bytes of one frame the 2x2 taps read
TAP_BYTES = 288 * 6 * DST[0]     # the same after the host tap gather
SEED = 1002  # config 2 clip
METRIC = "frames/sec (1080p ContentDetector)"
UNIT = "frames/s"
WORKLOAD = ("configs[1]: ContentDetector(threshold=27,min_scene_len=15) on the synthetic 1920x1080 30fps clip "
            "(seed 1002), auto-downscale 256x144, scored in device-resident batches")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def bench_config(n_gpus: int, frames_per_step: int) -> dict:
    """The `config` object of the JSON line -- identical in both arms (the reference arm times a bounded sample of it)."""
    return {"workload": WORKLOAD, "frames_per_step": frames_per_step, "per_gpu_batch_bytes": frames_per_step * W * H * 3,
            "l2": "inputs larger than L2 (12.7 GB batch re-read every step, evict-first)", "parallelism": f"frame-range x{n_gpus}",
            "dst": list(DST)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons (B200_PROFILING.md recipe); every sample is stamped on receipt so that only
    the ones taken inside the timed region are summarised."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t_begin: float, t_end: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summarise(rows):
            sm, mx, reasons, power = [], [], set(), []
            for _, ln in rows:
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
                except ValueError:
                    continue
                for name, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons, power

        # a sample printed at time t describes the interval before t: keep those received inside the region
        inside = [r for r in self.lines if t_begin + 0.03 <= r[0] <= t_end + 0.03]
        sm, mx, reasons, power = summarise(inside)
        where = "timed region"
        if not sm:  # a region shorter than the sampling period: fall back to everything seen under load
            sm, mx, reasons, power = summarise(self.lines)
            where = "warm-up + timed region (no sample fell inside the timed region)"
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "sampled_during": where, "sm_mhz_min": float(min(sm)), "power_w_max": max(power)}


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so the pinned host frames of the e2e leg are
    first-touched on the NUMA node that owns the GPU's PCIe root (matters when 8 ranks pull 50 GB/s each)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_sample_frames(n: int, start: int = 0) -> np.ndarray:
    """n frames of the config-2 clip on the host, from the CPU twin of the clip generator (oracle/esd_oracle.c over
    synthclip/synth_core.h).  Never touches the product library or a GPU."""
    from oracle import c_oracle
    from synthclip import schedule

    sch = schedule.build_schedule(SEED, start + n)
    return c_oracle.synth_frames(SEED, W, H, sch.descs[start:start + n])


# ---------------------------------------------------------------------------------------------------
def bench_reference(args):
    """The reference arm: PySceneDetect's detector logic on real cv2 (oracle/psd_cv2.py -- the package itself is not
    installable offline), one process per host core, frames already decoded in RAM.  Honours --steps / --warmup: the
    passes per step shrink so that the whole run stays within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_baseline

    cores = cpu_baseline.available_cores()
    sample = args.ref_sample
    frames = cpu_sample_frames(sample)
    with cpu_baseline.Runner(frames, "content", cores=cores) as runner:
        cal = runner.step(1)  # calibration pass (also faults the sample in): seconds per pass
        budget_s = args.ref_budget_s / max(1, args.steps + args.warmup)
        reps = args.ref_reps if args.ref_reps > 0 else int(max(1, min(40, budget_s / max(cal["seconds"], 1e-3))))
        times, total, res = [], 0, cal
        for i in range(args.warmup + args.steps):
            res = runner.step(reps)
            if i >= args.warmup:
                times.append(res["seconds"])
                total += res["frames_total"]
    dt = sum(times)
    value = total / dt
    # the compressed-file leg of the same arm: cv2.VideoCapture decode (the reference's decode loop, model_manager.py:237-263)
    # + PySceneDetect logic on cv2, one process per core, each decoding and scoring the whole file -- the counterpart of the GPU
    # arm's `e2e_compressed`.  The file is the same Motion-JPEG AVI (frames from the CPU twin of the clip generator).
    comp = None
    if not args.no_compressed:
        path = None
        try:  # this leg must never cost the arm its headline line
            import cv2

            n = args.compressed_frames
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
            path = os.path.join(shm, f"esd_bench_mjpeg_ref_{os.getpid()}.avi")
            wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), float(FPS), (W, H))
            if not wr.isOpened():
                raise RuntimeError("this OpenCV build cannot write Motion-JPEG")
            for a in range(0, n, 32):
                for f in cpu_sample_frames(min(32, n - a), a):
                    wr.write(f)
            wr.release()
            with cpu_baseline.Runner(video_path=path, cores=cores) as runner:
                runner.step(1)
                r = runner.step(args.compressed_cpu_passes)
            comp = {"value": r["frames_per_s"], "unit": UNIT, "cores": r["cores"], "seconds": r["seconds"], "file_frames": n,
                    "file_bytes": os.path.getsize(path),
                    "what": "cv2.VideoCapture decode + PySceneDetect logic on cv2, one process per core, each decoding and scoring the whole file"}
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"bench.py: compressed leg of the reference arm skipped ({type(e).__name__}: {e})\n")
            comp = {"error": f"{type(e).__name__}: {e}"[:300], "unit": UNIT}
        finally:
            if path and os.path.exists(path):
                os.remove(path)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args.gpus, args.frames_per_step),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "cpu_model": cpu_model(), "kind": "port",
                         "sample": f"each step: {cores} processes x {sample} frames x {reps} passes of the config-2 clip (bounded sample of the "
                                   f"workload), PySceneDetect logic restated over {res['backend']} (scenedetect itself is not installable "
                                   f"offline), frames in RAM from the CPU twin of the clip generator"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if comp is not None:
        line["e2e_compressed"] = comp
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
def parity_gate(capi, cfg, local, clip, first, rank):
    """Every rank pushes its resident frames through a fresh context and compares integer sums and float64 scores with
    the committed cv2 golden of the very clip that was timed (rank 0 also the complete cut list of the clip's head)."""
    import torch

    out = {"checked": False, "frames": 0, "bit_exact": None}
    full_p = os.path.join(GOLDEN, "clip_c2_1080p_full.npz")
    head_p = os.path.join(GOLDEN, "clip_c2_1080p_head.npz")
    stream = torch.cuda.current_stream().cuda_stream
    ok = True
    nb = int(clip.shape[0])
    if os.path.exists(full_p):
        g = np.load(full_p)
        m = max(0, min(nb, int(g["n_frames"]) - first))
        if m > 1:
            with capi.EsdContext(cfg, local) as ctx:
                ctx.push_tensor(clip[:m], first, stream)
                sc = ctx.read_scores(first, m, ["sums3", "content_val"])
            # the first resident frame has no predecessor on this rank (frame 0 of the clip has none at all: 0 == golden)
            a = 0 if first == 0 else 1
            ok = ok and np.array_equal(sc["sums3"][a:], g["sums3"][first + a:first + m])
            ok = ok and np.array_equal(sc["content_val"][a:].view(np.uint64), g["content_val"][first + a:first + m].view(np.uint64))
            out["frames"] += m - a
            out["checked"] = True
            out["against"] = "tests/golden/clip_c2_1080p_full.npz (oracle/psd_cv2.py on real cv2)"
    if rank == 0 and os.path.exists(head_p):
        g = np.load(head_p)
        m = int(g["n_frames"])
        if m <= nb:
            with capi.EsdContext(cfg, local) as ctx:
                ctx.push_tensor(clip[:m], 0, stream)
                cuts, _ = ctx.get_cuts(capi.ESD_DET_CONTENT)
                sc = ctx.read_scores(0, m, ["sums3", "content_val"])
            ok = ok and cuts == g["cuts_content"].tolist()
            ok = ok and np.array_equal(sc["sums3"], g["sums3"]) and np.array_equal(sc["content_val"].view(np.uint64), g["content_val"].view(np.uint64))
            out["checked"] = True
            out["cuts_checked"] = len(cuts)
            out["frames"] = max(out["frames"], m)
    out["bit_exact"] = bool(ok) if out["checked"] else None
    return out


def e2e_leg(capi, cfg, local, host_np, mode, gthreads, steps, ring, barrier, all_max, world):
    """One end-to-end mode through the public ingest API: host frames -> ring -> H2D -> scoring -> cuts + scores D2H."""
    NE = host_np.shape[0]
    gather = mode.endswith("gather")
    ectx = capi.EsdContext(cfg, local)
    try:
        slots, per = ring if ring else ((4, 128) if gather else (3, 256))
        ectx.ingest_open(slots, per)
        ectx.ingest_set_gather(gthreads if gather else 0)

        def step(p):
            ectx.ingest_push_numpy(host_np, p)
            cuts_e, _ = ectx.get_cuts(capi.ESD_DET_CONTENT, 0)
            sc = ectx.read_scores(p, NE, ["content_val"])
            return len(cuts_e), sc["content_val"].nbytes

        p = 0
        for _ in range(2):
            step(p); p += NE
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(steps):
            ncut, nb = step(p); p += NE
            d2h = nb + 8 * ncut + 128
        ectx.synchronize()
        dt = all_max(time.perf_counter() - t0)
        h2d = ectx.ingest_stats()[0] // (steps + 2)
        ectx.ingest_close()
    finally:
        ectx.close()
    fps = steps * NE * world / dt
    # host DRAM traffic per frame (estimate): what the CPU and the DMA engine read and write in host memory
    dram = {"dma_rows": TOUCHED_ROW_BYTES, "gather": TOUCHED_ROW_BYTES + 2 * TAP_BYTES, "pageable_gather": TOUCHED_ROW_BYTES + 2 * TAP_BYTES,
            "pageable_rows": 3 * TOUCHED_ROW_BYTES}.get(mode, TOUCHED_ROW_BYTES)
    return {"value": fps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "pcie_GBps": fps * h2d / NE / 1e9,
            "host_dram_bytes_per_frame_est": dram, "host_dram_GBps_est": fps * dram / 1e9}


def config3_leg(capi, synth, local, rank, world, dist, dev, args):
    """BASELINE config 3 on the real ranks: AdaptiveDetector(3.0, window_width=2) on the 18 000-frame config-2 clip,
    frame ranges + (w+1)/w halo per rank, owned float64 scores all-gathered as plain tensors, ONE global decision pass on
    rank 0; timed as a whole (scoring + gather + decision), warm-up passes first, checked against the cv2 golden."""
    import torch

    from eioku_b200 import multi
    from eioku_b200.detectors import AdaptiveDetector
    from eioku_b200.scene_manager import SceneManager

    n_total, ww = args.config3_frames, 2
    sm = SceneManager(device=local, tuning={"initial_capacity": n_total + 64})
    sm.add_detector(AdaptiveDetector(adaptive_threshold=3.0, window_width=ww))
    ctx = sm.make_context(W, H)
    job = multi.DistributedShard(ctx, [capi.ESD_DET_ADAPTIVE], n_total, ww, rank, world, local)
    sh = job.shard
    n_load = sh.load_end - sh.load_start
    sch = synth.build_schedule(SEED, n_total)
    frames = torch.empty((n_load, H, W, 3), dtype=torch.uint8, device=dev)
    synth.fill(frames, SEED, sch.descs[sh.load_start:sh.load_end], chunk=256)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    batch = 2048

    def one_pass(want_scores=False):
        ctx.reset()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        pos = sh.load_start
        for a in range(0, n_load, batch):
            m = min(batch, n_load - a)
            ctx.push_tensor(frames[a:a + m], pos, stream)
            pos += m
        cuts, merged = job.finish(stream, want_scores)  # rank 0 returns once the decision pass has finished
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), 1000 * (time.perf_counter() - t0), cuts, merged

    one_pass()  # warm-up: lazy module load, plan + scratch allocation, NCCL channel set-up
    one_pass()
    runs = [one_pass(want_scores=(i == 0)) for i in range(5)]
    cuts, merged = runs[0][2], runs[0][3]
    t = torch.tensor([[r[0], r[1]] for r in runs], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev = float(t[:, 0].median())
    ms_wall = float(t[:, 1].median())

    # single-GPU time for the same number of frames, measured in the same run on rank 0: its resident shard re-scored
    # until n_total frames have gone through one context, then the same decision pass (kernel time is content-independent)
    t1_ms = None
    if rank == 0:
        reps = []
        buf = torch.empty(n_total, dtype=torch.float64, device=dev)
        for _ in range(4):
            ctx.reset()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            done = 0
            while done < n_total:
                m = min(n_load, n_total - done, batch)
                ctx.push_tensor(frames[:m], done, stream)
                done += m
            ctx.copy_scores_device(capi.ESD_SCORE_ADAPTIVE_VAL, 0, n_total, buf.data_ptr(), -1, stream)
            ctx.decide_device(capi.ESD_DET_ADAPTIVE, 0, buf.data_ptr(), n_total, stream)
            e1.record()
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1))
        t1_ms = float(np.median(reps[1:]))
    out = None
    if rank == 0:
        parity = None
        gp = os.path.join(GOLDEN, "clip_c2_1080p_full.npz")
        mine = cuts[capi.ESD_DET_ADAPTIVE]
        if os.path.exists(gp) and n_total == 18000:
            g = np.load(gp)
            av = merged[0].cpu().numpy()
            cuts2, ratio = ctx.decide_arrays(capi.ESD_DET_ADAPTIVE, 0, av)  # host-array flavour of the same pass: gives the ratios
            parity = bool(mine == g["cuts_adaptive"].tolist() and cuts2 == mine and
                          np.array_equal(np.nan_to_num(ratio, nan=-1).view(np.uint64), np.nan_to_num(g["adaptive_ratio"], nan=-1).view(np.uint64)) and
                          np.array_equal(av.view(np.uint64), g["content_val"].view(np.uint64)))
        out = {"what": "configs[2]: AdaptiveDetector(3.0, window_width=2), 18 000-frame 1080p clip, frame-range shards + (w+1)/w halo, "
                       "owned scores all-gathered as float64 tensors, one global decision pass on rank 0",
               "n_gpus": world, "frames": n_total, "frames_loaded_rank0": int(n_load),
               "ms_total": ms_dev, "ms_total_wall": ms_wall, "value": n_total / (ms_dev / 1000.0), "unit": UNIT,
               "t1_ms_same_run_rank0": t1_ms, "strong_scaling_eff": (t1_ms / (world * ms_dev)) if t1_ms else None,
               "cuts": len(mine), "bit_exact_vs_golden": parity,
               "timed": "median of 5 passes after 2 warm-up passes; a pass = barrier, push the shard, D2D copy of the owned scores, all-gather, "
                        "global decision on rank 0; CUDA events, max over ranks"}
    del frames
    ctx.close()
    torch.cuda.empty_cache()
    return out


def compressed_leg(local, rank, world, dist, dev, args, barrier, all_max):
    """The compressed-file leg must never cost the run its headline: any failure in it (no MJPG writer in this OpenCV build, no
    room in /dev/shm, a decoder error ...) is reported as `{"error": ...}` in `e2e_compressed` instead of killing the process --
    on every rank alike, because the leg contains collectives: at two points the ranks agree (all-reduce of a failure flag) whether
    all of them are still in, and leave together if not.  Only a parity mismatch stays fatal (SystemExit)."""

    class _LegFailed(Exception):
        pass

    state = {"mine": None}

    def agree(where, err=None):
        if err is not None and state["mine"] is None:
            state["mine"] = f"{where}: {err}"
        if all_max(1.0 if state["mine"] is not None else 0.0) > 0.0:
            raise _LegFailed(state["mine"] or f"{where}: another rank failed")

    try:
        try:
            return _compressed_leg_body(local, rank, world, dist, dev, args, barrier, all_max, agree)
        except _LegFailed:
            raise
        except SystemExit:
            raise
        except Exception as e:  # noqa: BLE001 - a failure before the first agreement point: tell the others, then leave
            if state["mine"] is None:
                state["mine"] = f"{type(e).__name__}: {e}"
            try:
                all_max(1.0)
            except Exception:  # noqa: BLE001
                pass
            raise _LegFailed(state["mine"]) from e
    except _LegFailed as e:
        sys.stderr.write(f"bench.py: compressed leg skipped ({e})\n")
        return {"error": str(e)[:300], "unit": UNIT}


def _compressed_leg_body(local, rank, world, dist, dev, args, barrier, all_max, agree):
    """End to end from a COMPRESSED file: a 1080p Motion-JPEG AVI of the clip's head in /dev/shm -> nvJPEG on
    the GPU (libesd_decode) -> scoring -> cuts; `sessions` decoder sessions per GPU, each decoding and scoring the whole file
    (the shape of the CPU arm: one process per core, each decoding and scoring the whole file with cv2.VideoCapture + the
    PySceneDetect logic).  Decoded frames never exist in host memory; only the compressed bytes cross PCIe."""
    import threading

    import cv2
    import torch

    import synthclip as synth
    from eioku_b200 import decode
    from eioku_b200.detectors import ContentDetector
    from eioku_b200.scene_manager import SceneManager

    n = args.compressed_frames
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(shm, f"esd_bench_mjpeg_{os.getpid()}.avi")
    sch = synth.build_schedule(SEED, n)
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), float(FPS), (W, H))
    if not wr.isOpened():
        raise RuntimeError("this OpenCV build cannot write Motion-JPEG")
    for a in range(0, n, 64):
        t = torch.empty((min(64, n - a), H, W, 3), dtype=torch.uint8, device=dev)
        synth.fill(t, SEED, sch.descs[a:a + 64])
        for f in t.cpu().numpy():
            wr.write(f)
    wr.release()
    size = os.path.getsize(path)
    # a decoder session is one host thread that spends its time inside libesd_decode / libesd (GIL released): ~30 us of host work
    # per picture, so the session count is chosen for pictures in flight on the GPU, not by the core count
    # (a node's cores are shared by its ranks: at least 2, at most 4 sessions per rank)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 8
    # (with a block of threads per picture the GPU side saturates at 3-4 sessions: profiles/r02_decode_parallel.md)
    sessions = args.decode_sessions if args.decode_sessions > 0 else max(2, min(4, cores // max(1, world)))
    out = None
    try:
        # parity on the decoded surface: the frames one session scored, downloaded, through the oracle's integer chain
        from oracle import c_oracle

        sm = SceneManager(device=local, batch_frames=64)
        sm.add_detector(ContentDetector())
        with decode.MjpegVideo(path, device=local, batch_frames=64) as v:
            backend = v.backend
            tee = decode.TeeVideo(v)
            sm.detect_scenes(tee, collect_scores=True)
            m = min(n, 96)
            surf = tee.frames()[:m].cpu().numpy()
        want, _, _ = c_oracle.score_frames(surf, DST[0], DST[1])
        bit_exact = bool(np.array_equal(sm.scores["sums3"][:m].astype(np.int64), want))
        cuts_ref = sm.get_cut_list()
        sm.close()
        if not bit_exact:
            raise SystemExit("bench.py: PARITY GATE FAILED on the decoded surface")

        agree("set-up")          # every rank has its file, its parity check and the session count: none goes on alone
        frames_done = [0] * sessions
        errs = []
        gate = threading.Barrier(sessions + 1)
        passes = args.compressed_passes

        def work(i):
            try:
                with torch.cuda.device(local):
                    st = torch.cuda.Stream(device=local)
                    with torch.cuda.stream(st):
                        smi = SceneManager(device=local, batch_frames=args.decode_batch, tuning={"reserved2": 24})
                        smi.add_detector(ContentDetector())
                        vi = decode.MjpegVideo(path, device=local, batch_frames=args.decode_batch)
                        smi.detect_scenes(vi, reuse_context=True)  # warm-up pass
                        gate.wait()
                        for _ in range(passes):
                            vi.seek(0)
                            frames_done[i] += smi.detect_scenes(vi, reuse_context=True)
                            if smi.get_cut_list() != cuts_ref:
                                raise RuntimeError("cut list changed between passes")
                        st.synchronize()
                        gate.wait()
                        vi.close()
                        smi.close()
            except BaseException as e:  # noqa: BLE001
                errs.append(repr(e))
                gate.abort()

        th = [threading.Thread(target=work, args=(i,)) for i in range(sessions)]
        for t in th:
            t.start()
        try:
            gate.wait()
            barrier()
            t0 = time.perf_counter()
            gate.wait()
            dt = time.perf_counter() - t0
        except threading.BrokenBarrierError:
            dt = float("nan")
        for t in th:
            t.join()
        agree("timed region", errs[0] if errs else None)
        dt = all_max(dt)
        value = world * sum(frames_done) / dt
        out = {"value": value, "unit": UNIT, "n_gpus": world,
               "decoder": ("libesd_decode's own baseline-JPEG kernels (bit-identical to cv2.imdecode)" if backend == "native" else f"nvJPEG ({backend}) via libesd_decode") + ", Motion-JPEG AVI",
               "decode_batch": args.decode_batch,
               "sessions_per_gpu": sessions, "file_frames": n, "file_bytes": size, "h2d_bytes_per_frame": size // n,
               "passes_per_session": passes, "seconds": dt, "bit_exact_on_decoded_surface": bit_exact, "cuts": len(cuts_ref),
               "note": "compressed file (page cache) -> compressed bytes H2D -> GPU decode -> scoring -> cuts D2H; decoded frames never visit host "
                       "memory.  NVDEC is closed to this container (profiles/r02_nvdec_caps.log), so the codec is MJPEG"}
        if world == 1 and not args.no_cpu:
            from oracle import cpu_baseline

            with cpu_baseline.Runner(video_path=path) as runner:
                runner.step(1)
                r = runner.step(args.compressed_cpu_passes)
            out["cpu_arm"] = {"value": r["frames_per_s"], "unit": UNIT, "cores": r["cores"], "seconds": r["seconds"],
                              "what": "cv2.VideoCapture decode (the reference's decode loop, model_manager.py:237-263) + PySceneDetect logic on cv2, "
                                      "one process per core, each decoding and scoring the whole file"}
            out["ratio_vs_cpu_arm"] = value / r["frames_per_s"]
    finally:
        try:
            os.remove(path)
        except OSError:
            pass
    return out


def bench_ours(args):
    import torch

    import synthclip as synth
    from eioku_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; eioku_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist  # plumbing only: barrier + max-reduce of the timing (+ the config-3 score gather)

        dist.init_process_group("nccl", device_id=torch.device(dev))
    NB = args.frames_per_step
    K = args.steps

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- synthetic clip batch, resident in HBM (rank r scores its own frame range of the clip)
    first = rank * NB
    sch = synth.build_schedule(SEED, first + NB)
    clip = torch.empty((NB, H, W, 3), dtype=torch.uint8, device=dev)
    synth.fill(clip, SEED, sch.descs[first:first + NB])
    torch.cuda.synchronize()

    cfg = capi.default_config()
    cfg.detectors = capi.ESD_DET_CONTENT
    cfg.src_width, cfg.src_height = W, H
    for kv in args.tune:
        k, v = kv.split("=")
        setattr(cfg, k, int(v))
    stream = torch.cuda.current_stream().cuda_stream

    # ---- how many K-step segments: at least args.min_segments and at least args.min_seconds of device time
    cfg.initial_capacity = 64 * NB
    with capi.EsdContext(cfg, local) as pctx:
        pos = 0
        for _ in range(3):
            pctx.push_tensor(clip, pos, stream); pos += NB
        pctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            pctx.push_tensor(clip, pos, stream); pos += NB
        pctx.join(stream)
        e1.record()
        torch.cuda.synchronize()
        est_step_ms = all_max(e0.elapsed_time(e1) / 8.0)
    segments = int(max(args.min_segments, math.ceil(args.min_seconds * 1000.0 / max(1e-3, est_step_ms * K))))
    segments = min(segments, max(args.min_segments, int(args.max_seconds * 1000.0 / max(1e-3, est_step_ms * K))))
    total_steps = args.warmup + segments * K
    cfg.initial_capacity = (total_steps + 64) * NB
    cfg.max_cuts = max(65536, 64 * (total_steps + 64))  # the repeated clip yields ~12 cuts per batch
    ctx = capi.EsdContext(cfg, local)
    geo = ctx.geometry

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    pos = 0
    for i in range(args.warmup):
        ctx.push_tensor(clip, pos, stream)
        pos += NB
        if i % 16 == 15:
            ctx.synchronize()
    ctx.synchronize()
    ctx.set_timing(True)
    launches0 = ctx.kernel_launches
    barrier()
    t_begin = time.time()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(segments + 1)]
    evs[0].record()
    for s in range(segments):
        for _ in range(K):
            ctx.push_tensor(clip, pos, stream)
            pos += NB
        ctx.join(stream)  # the last batch's finalize/decision tail is part of the segment
        evs[s + 1].record()
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    seg_ms = torch.tensor([evs[s].elapsed_time(evs[s + 1]) for s in range(segments)], dtype=torch.float64, device=dev)
    fused_ms, fused_n = ctx.kernel_time()
    ctx.set_timing(False)
    launches = ctx.kernel_launches - launches0
    _, n_cuts = ctx.get_cuts(capi.ESD_DET_CONTENT)
    if dist is not None:
        dist.all_reduce(seg_ms, op=dist.ReduceOp.MAX)
    fused_ms_max = all_max(fused_ms)
    seg = seg_ms.cpu().numpy()
    ms_seg = float(np.median(seg))
    value = K * NB * world / (ms_seg / 1000.0)
    ctx.close()
    cfg.initial_capacity = 8 * NB
    cfg.max_cuts = 65536

    # ---- parity gate on the frames that were just timed
    parity = parity_gate(capi, cfg, local, clip, first, rank)
    pt = torch.tensor([0.0 if parity["bit_exact"] is False else 1.0, float(parity["frames"])], dtype=torch.float64, device=dev)
    if dist is not None:
        pmin = pt.clone()
        dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(pt, op=dist.ReduceOp.SUM)
        if parity["checked"]:
            parity["bit_exact"] = bool(pmin[0] > 0.5)
        parity["frames"] = int(pt[1])
        parity["ranks_checked"] = world
    if parity["bit_exact"] is False:
        if rank == 0:
            sys.stderr.write("bench.py: PARITY GATE FAILED -- scores differ from tests/golden; no number reported\n")
        if dist is not None:
            dist.destroy_process_group()
        return 3

    # ---- end to end through the public API with HOST frames; every mode at every N, `e2e` = the best
    NE = args.e2e_frames
    host_pinned = clip[:NE].cpu().pin_memory()
    cores = len(os.sched_getaffinity(0)) if world == 1 else max(1, (os.cpu_count() or 1) // world)
    gthreads = args.e2e_gather_threads if args.e2e_gather_threads >= 0 else max(1, cores)
    e2e_steps = max(2, min(args.e2e_steps, 12))
    ring = tuple(int(v) for v in args.e2e_ring.split("x")) if args.e2e_ring else None
    modes = {}
    for mode in [m for m in args.e2e_modes.split(",") if m]:
        src = host_pinned.numpy() if not mode.startswith("pageable") else np.array(host_pinned.numpy(), copy=True)
        modes[mode] = e2e_leg(capi, cfg, local, src, mode, gthreads, e2e_steps, ring, barrier, all_max, world)
        modes[mode]["source"] = "pageable numpy array" if mode.startswith("pageable") else "pinned host memory"
        if mode.endswith("gather"):
            modes[mode]["host_gather_threads_per_rank"] = gthreads
        del src
    best = max(modes, key=lambda m: modes[m]["value"]) if modes else None
    del host_pinned

    # ---- end to end from a compressed file: GPU decode -> scoring (frames never in host memory)
    comp = None
    if not args.no_compressed:
        comp = compressed_leg(local, rank, world, dist, dev, args, barrier, all_max)

    # ---- config 3 (adaptive, halo shards, one global decision) on the real ranks
    c3 = None
    if args.config3 == "on" or (args.config3 == "auto" and world > 1):
        del clip
        clip = None
        torch.cuda.empty_cache()
        c3 = config3_leg(capi, synth, local, rank, world, dist, dev, args)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (fused) kernel, measured live with CUDA events on its stream
    peak, peak_src = measured_peak_gbs()
    alg_bytes_launch = NB * int(geo.alg_bytes_per_frame)
    avg_kernel_ms = fused_ms_max / max(1, fused_n)
    achieved = alg_bytes_launch / (avg_kernel_ms / 1000.0) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "fused_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- CPU baseline on this box's host cores (bounded sample; rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        if clip is not None:
            del clip
        torch.cuda.empty_cache()
        from oracle import cpu_baseline

        r = cpu_baseline.run(cpu_sample_frames(args.cpu_sample), "content", reps=args.cpu_reps)
        cpu = {"value": r["frames_per_s"], "unit": UNIT, "cores": r["cores"], "cpu_model": cpu_model(), "kind": "port",
               "sample": f"{args.cpu_sample} frames of the same clip x {args.cpu_reps} passes per process, one process per core, "
                         f"PySceneDetect logic over {r['backend']} ({r['seconds']:.1f} s)",
               "per_core": r["per_core_frames_per_s"]}

    e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if best:
        e2e = {"value": modes[best]["value"], "unit": UNIT, "h2d_bytes_per_step": modes[best]["h2d_bytes_per_step"],
               "d2h_bytes_per_step": modes[best]["d2h_bytes_per_step"], "frames_per_step": NE, "mode": best, "modes": modes,
               "note": "host frames -> esd_ingest_push_host -> H2D -> scoring -> cuts + scores D2H, every step.  dma_rows: the touched rows "
                       "DMA'd straight from pinned memory; gather: host threads copy only the tap bytes into the pinned ring; pageable_*: the "
                       "same from a plain numpy array.  Limiter: host DRAM + PCIe, not a kernel (frames are born in host memory)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_seg / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": bench_config(world, NB),
        "timing": {"segments": segments, "steps_per_segment": K, "timed_region_s": float(seg.sum() / 1000.0), "segment_ms_median": ms_seg,
                   "segment_ms_min": float(seg.min()), "segment_ms_max": float(seg.max()),
                   "method": "CUDA events around back-to-back segments of exactly K steps; per segment the max over ranks; value = frames of one "
                             "segment / median segment time"},
        "parity": parity,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_frame": int(geo.alg_bytes_per_frame),
                     "kernel": "fused_score_kernel", "avg_kernel_ms": avg_kernel_ms, "kernel_launches_timed": int(fused_n),
                     "kernel_share_of_step": fused_ms_max / float(seg.sum()),
                     "equivalent_ingest_GBps_not_roofline": value / world * W * H * 3 / 1e9},
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "cuts_found": int(n_cuts),
        "clocks": clocks,
    }
    if comp is not None:
        line["e2e_compressed"] = comp
    if c3 is not None:
        line["config3"] = c3
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=2048)
    ap.add_argument("--min-seconds", type=float, default=1.0, help="minimum device time of the timed region (segments of K steps are repeated)")
    ap.add_argument("--max-seconds", type=float, default=20.0, help="cap of the timed region (never fewer than --min-segments segments)")
    ap.add_argument("--min-segments", type=int, default=10)
    ap.add_argument("--e2e-frames", type=int, default=1024)
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--e2e-modes", default="gather,dma_rows,pageable_gather,pageable_rows",
                    help="comma list of end-to-end ingest modes to run (the best becomes `e2e`)")
    ap.add_argument("--cpu-sample", type=int, default=192)
    ap.add_argument("--cpu-reps", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-gather-threads", type=int, default=-1, help="host gather threads per rank (-1 = this rank's share of the host cores)")
    ap.add_argument("--e2e-ring", default="", help="ingest ring of the e2e leg as SLOTSxFRAMES (default 4x128 with gather, 3x256 DMA)")
    ap.add_argument("--no-compressed", action="store_true", help="skip the compressed-file end-to-end leg")
    ap.add_argument("--compressed-frames", type=int, default=1024, help="length of the Motion-JPEG file (four decode batches per pass: the decoder keeps two batches in flight per session)")
    ap.add_argument("--compressed-passes", type=int, default=20, help="passes over the file per session inside the timed region (~1 s at 80 k frames/s)")
    ap.add_argument("--compressed-cpu-passes", type=int, default=2)
    ap.add_argument("--decode-sessions", type=int, default=0, help="GPU decoder sessions per rank (0 = up to 4, fewer when the ranks share few cores)")
    ap.add_argument("--decode-batch", type=int, default=64, help="pictures per decode batch and session")
    ap.add_argument("--config3", default="auto", choices=["auto", "on", "off"], help="run BASELINE config 3 on the ranks (auto: when N > 1)")
    ap.add_argument("--config3-frames", type=int, default=18000)
    ap.add_argument("--ref-sample", type=int, default=192, help="--impl reference: frames per process per pass")
    ap.add_argument("--ref-reps", type=int, default=0, help="--impl reference: passes per process per step (0 = fit --ref-budget-s)")
    ap.add_argument("--ref-budget-s", type=float, default=100.0, help="--impl reference: wall-clock budget of all steps together")
    ap.add_argument("--tune", action="append", default=[], help="esd_config field=value (e.g. rows_per_group=2)")
    args = ap.parse_args()
    if args.warmup < 3:
        sys.stderr.write("bench.py: --warmup raised to 3 (timing rule: at least 3 warm-up steps)\n")
        args.warmup = 3
    if args.impl == "reference":
        return bench_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return bench_ours(args)


if __name__ == "__main__":
    sys.exit(main())
