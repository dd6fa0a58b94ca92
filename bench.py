#!/usr/bin/env python
"""bench.py -- frames/s of the scene-scoring hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours:      one process per GPU (torchrun for N > 1), device-resident synthetic 1080p frames of the
           config-2 clip (ContentDetector(27, 15)), one step = one 2048-frame batch through
           esd_push_frames (fused TMA kernel + finalize + decision).  No collective on the data path:
           ranks score independent frame ranges (weak scaling).  Prints ONE JSON line on rank 0.
reference: the reference's CPU path (PySceneDetect logic on cv2, one process per host core) on a
           bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

W, H, FPS = 1920, 1080, 30
SEED = 1002  # config 2 clip
METRIC = "frames/sec (1080p ContentDetector)"
UNIT = "frames/s"
WORKLOAD = ("configs[1]: ContentDetector(threshold=27,min_scene_len=15) on the synthetic 1920x1080 30fps clip "
            "(seed 1002), auto-downscale 256x144, scored in device-resident batches")


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
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
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


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


def host_sample_frames(n: int, start: int = 0) -> np.ndarray:
    """n frames of the config-2 clip on the host (GPU generator if there is one, else the CPU twin)."""
    import synthclip as synth

    sch = synth.build_schedule(SEED, start + n)
    descs = sch.descs[start:start + n]
    try:
        import torch

        if torch.cuda.is_available():
            from eioku_b200 import capi

            out = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda:0")
            synth.fill(out, SEED, descs)
            return out.cpu().numpy()
    except Exception:
        pass
    from oracle import c_oracle

    return c_oracle.synth_frames(SEED, W, H, descs)


def run_cpu_baseline(sample_frames: int, reps: int):
    from oracle import cpu_baseline

    frames = host_sample_frames(sample_frames)
    r = cpu_baseline.run(frames, "content", reps=reps)
    return r


# ---------------------------------------------------------------------------------------------------
def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_baseline

    cores = cpu_baseline.available_cores()
    sample, reps = args.ref_sample, args.ref_reps
    frames = host_sample_frames(sample)
    times = []
    total = 0
    res = None
    # a CPU step is ~3 s of wall clock (16 cores); keep the whole arm within a few minutes whatever K/W the caller passes
    args.steps = min(args.steps, 12)
    args.warmup = min(args.warmup, 2)
    for i in range(args.warmup + args.steps):
        res = cpu_baseline.run(frames, "content", cores=cores, reps=reps)
        if i >= args.warmup:
            times.append(res["seconds"])
            total += res["frames_total"]
    dt = sum(times)
    value = total / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": f"{cores} processes x {sample} frames x {reps} passes each (bounded sample of the clip)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "cpu_model": cpu_model(), "kind": "port",
                         "sample": f"{sample} frames x {reps} passes of the config-2 clip per process per step, PySceneDetect logic restated "
                                   f"over {res['backend']} (scenedetect itself is not installable offline), frames in RAM"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def bench_ours(args):
    import torch

    from eioku_b200 import capi
    import synthclip as synth

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
        import torch.distributed as dist  # plumbing only: barrier + max-reduce of the timing

        dist.init_process_group("nccl", device_id=torch.device(dev))
    NB = args.frames_per_step

    # ---- synthetic clip batch, resident in HBM (rank r scores its own frame range of the clip)
    first = rank * NB
    sch = synth.build_schedule(SEED, first + NB)
    clip = torch.empty((NB, H, W, 3), dtype=torch.uint8, device=dev)
    synth.fill(clip, SEED, sch.descs[first:first + NB])
    torch.cuda.synchronize()

    cfg = capi.default_config()
    cfg.detectors = capi.ESD_DET_CONTENT
    cfg.src_width, cfg.src_height = W, H
    cfg.initial_capacity = (args.steps + args.warmup + 2100) * NB
    cfg.max_cuts = max(65536, 64 * (args.steps + args.warmup + 2100))  # the repeated clip yields ~12 cuts per batch
    for kv in args.tune:
        k, v = kv.split("=")
        setattr(cfg, k, int(v))
    ctx = capi.EsdContext(cfg, local)
    geo = ctx.geometry
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    pos = 0
    # the clock sampler starts before the warm-up: nvidia-smi needs ~100 ms to deliver its first sample and a
    # short timed region would otherwise end before it; every sample is taken under load (warm-up + timed steps)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_w = time.perf_counter()
    nw = 0
    while nw < args.warmup or (rank == 0 and time.perf_counter() - t_w < 0.4 and nw < 2000):
        ctx.push_tensor(clip, pos, stream)
        pos += NB
        nw += 1
        if nw % 16 == 0:
            ctx.synchronize()
    ctx.synchronize()
    ctx.set_timing(True)
    launches0 = ctx.kernel_launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ctx.push_tensor(clip, pos, stream)
        pos += NB
    ctx.join(stream)  # the last batch's finalize/decision tail is part of the step
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    fused_ms, fused_n = ctx.kernel_time()
    ctx.set_timing(False)
    launches = ctx.kernel_launches - launches0
    cuts, n_cuts = ctx.get_cuts(capi.ESD_DET_CONTENT)
    t = torch.tensor([ms, fused_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, fused_ms_max = float(t[0]), float(t[1])
    frames_total = args.steps * NB * world
    value = frames_total / (ms_max / 1000.0)

    # ---- end to end through the public API with HOST frames: pinned host -> ingest ring (touched rows only
    #      cross PCIe) -> scoring -> cut list + scores read back.  Every rank does it; aggregate reported.
    NE = args.e2e_frames
    host = clip[:NE].cpu().pin_memory()
    host_np = host.numpy()
    ectx = capi.EsdContext(cfg, local)
    # e2e ingest mode: host threads gather the tap bytes of the touched rows (442 KB/frame over PCIe instead of
    # 1.66 MB); 0 = plain DMA of the touched rows.  Default: this rank's share of the host cores.
    gthreads = args.e2e_gather_threads
    if gthreads < 0:
        try:
            share = len(os.sched_getaffinity(0)) if world == 1 else (os.cpu_count() or 1) // world
        except Exception:
            share = (os.cpu_count() or 1) // world
        # measured (profiles/r01_pcie.log): gather beats the 28-30 k frames/s of plain DMA from ~10 threads up
        gthreads = share if share >= 10 else 0
    if args.e2e_ring:
        ectx.ingest_open(*[int(v) for v in args.e2e_ring.split("x")])
    else:
        ectx.ingest_open(4, 128) if gthreads else ectx.ingest_open(3, 256)
    ectx.ingest_set_gather(gthreads)
    e2e_steps = max(2, min(args.steps, 6))

    def e2e_step(p):
        ectx.ingest_push_numpy(host_np, p)
        cuts_e, _ = ectx.get_cuts(capi.ESD_DET_CONTENT, 0)
        sc = ectx.read_scores(p, NE, ["content_val"])
        return len(cuts_e), sc["content_val"].nbytes

    p = 0
    for _ in range(2):
        e2e_step(p); p += NE
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(e2e_steps):
        ncut, nb = e2e_step(p); p += NE
        d2h = nb + 8 * ncut + 128
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = e2e_steps * NE * world / float(te[0])
    h2d_per_step = ectx.ingest_stats()[0] // (e2e_steps + 2)  # bytes that actually crossed PCIe per step
    ectx.ingest_close()
    ectx.close()

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
        del clip
        torch.cuda.empty_cache()
        r = run_cpu_baseline(args.cpu_sample, args.cpu_reps)
        cpu = {"value": r["frames_per_s"], "unit": UNIT, "cores": r["cores"], "cpu_model": cpu_model(), "kind": "port",
               "sample": f"{args.cpu_sample} frames of the same clip x {args.cpu_reps} passes per process, one process per core, "
                         f"PySceneDetect logic over {r['backend']} ({r['seconds']:.1f} s)",
               "per_core": r["per_core_frames_per_s"]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": NB, "per_gpu_batch_bytes": NB * W * H * 3,
                   "l2": "inputs larger than L2 (12.7 GB batch re-read every step, evict-first)", "parallelism": f"frame-range x{world}",
                   "dst": [geo.dst_width, geo.dst_height], "cuts_found": n_cuts},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_frame": int(geo.alg_bytes_per_frame),
                     "kernel": "fused_score_kernel", "avg_kernel_ms": avg_kernel_ms, "kernel_share_of_step": fused_ms_max / ms_max,
                     "equivalent_ingest_GBps_not_roofline": value / world * W * H * 3 / 1e9},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_per_step, "d2h_bytes_per_step": d2h,
                "frames_per_step": NE, "host_gather_threads": gthreads,
                "note": ("pinned host frames -> host threads gather the tap bytes of the touched rows -> pinned ring -> H2D -> "
                         "scoring -> cuts+scores D2H") if gthreads else
                        "pinned host frames -> touched-rows-only H2D ring -> scoring -> cuts+scores D2H"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=2048)
    ap.add_argument("--e2e-frames", type=int, default=1024)
    ap.add_argument("--cpu-sample", type=int, default=192)
    ap.add_argument("--cpu-reps", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-gather-threads", type=int, default=-1, help="host gather threads of the e2e leg (-1 = this rank's share of cores, 0 = DMA rows)")
    ap.add_argument("--e2e-ring", default="", help="ingest ring of the e2e leg as SLOTSxFRAMES (default 4x128 with gather, 3x256 DMA)")
    ap.add_argument("--ref-sample", type=int, default=192, help="--impl reference: frames per process per pass")
    ap.add_argument("--ref-reps", type=int, default=40, help="--impl reference: passes per process per step (same as the cpu_baseline leg)")
    ap.add_argument("--tune", action="append", default=[], help="esd_config field=value (e.g. rows_per_group=2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps
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
