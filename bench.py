#!/usr/bin/env python
"""Headline benchmark: decoded Mpixel/s for a batch of 1080p key frames through -yuvf (recon + loop filter -> I420).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload yuvf|yuv|ppm] [--batch B]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)
    python bench.py --impl reference ...                                              (reference CPU arm)

A step = one pass of the hot path over one batch (default 1024 frames per GPU: 256 each of the noise / rgbgrad /
checker / diag 1080p inputs under bench_data/, made offline by the reference's gen_ppm + encoder --q 75 --loopfilter).

  value      whole-job Mpixel/s with the parsed frames already resident in HBM (kernel stage only)
  e2e        same metric through the public batch API with HOST (pinned) inputs and outputs: H2D of the ten arrays
             of every frame + kernels + D2H of the I420 bytes inside the timed region
  roofline   algorithmic bytes (820 B/macroblock in + tight I420 out) / device-timed kernel duration,
             against the measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the reference's own m06+m07 on the box's host cores (oracle/cpu_baseline.py), bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FRAMES_1080P = ["noise_1920x1080_q75.webp", "rgbgrad_1920x1080_q75.webp", "checker_1920x1080_q75.webp", "diag_1920x1080_q75.webp"]
METRIC = "decoded Mpixel/s (1080p -yuvf batch)"


def algorithmic_bytes(w, h, workload):
    mb = ((w + 15) // 16) * ((h + 15) // 16)
    if workload == "ppm":
        return 820 * mb + 3 * w * h
    return 820 * mb + w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)


class ClockSampler:
    """nvidia-smi clocks and throttle reasons under load (profiling recipe's clocks line). Started before the warm-up;
    samples whose timestamp falls inside the timed region are preferred, else all samples since the warm-up began
    (the GPU runs the same kernel back to back through both)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc, self.t0, self.t1 = gpu_index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_timed(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.05)
        self.proc.terminate()
        self.t.join(timeout=2)
        ok = [(ts, r) for ts, r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        inside = [(ts, r) for ts, r in ok if self.t0 is not None and self.t0 <= ts <= self.t1]
        use, window = (inside, "timed region") if inside else (ok, "warm-up + timed region")
        if not use:
            return None
        sm = [float(r[1]) for _, r in use]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for _, r in use for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(use[0][1][2]), "power_w_max": max(float(r[3]) for _, r in use),
                "samples": len(sm), "window": window, "reasons": reasons}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, kernel_version):
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        t = json.loads(p.read_text()).get(workload)
        return t.get(str(kernel_version)) if isinstance(t, dict) else t
    return None


def files_for(args):
    return [str(ROOT / "bench_data" / f) for f in FRAMES_1080P]


# The contract is ONE JSON line on stdout. Libraries print there too (NCCL announces its version on the first collective
# when NCCL_DEBUG says so), so the process's stdout is pointed at stderr for the whole run and the line goes out through a
# saved descriptor.
_JSON_FD = None


def claim_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# ------------------------------------------------------------------------------------------------ reference arm
def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, str(ROOT / "oracle"))
    import cpu_baseline
    files = files_for(args)
    # bounded sample: a "step" is ~2 s of decoding on every core at once; calibration pass = warm-up
    seconds = min(60.0, max(6.0, 2.0 * args.steps))
    t0 = time.perf_counter()
    r = cpu_baseline.run(files, mode=args.workload, procs=None, seconds=seconds)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": seconds * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int16", "data": "synthetic (gen_ppm patterns through the reference encoder, --q 75 --loopfilter)",
        "config": {"workload": f"1920x1080 key frames, -{args.workload}, mix noise/rgbgrad/checker/diag in equal parts",
                   "note": "reference m06+m07 CPU code on all host threads; one process per core; bounded sample",
                   "wall_s": round(wall, 1)},
        "cpu_baseline": {"value": r["value"], "unit": "Mpixel/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    # CPU baseline first (rank 0, N=1 only): no GPU work competes for the host cores while it runs
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out = subprocess.run([sys.executable, str(ROOT / "oracle" / "cpu_baseline.py"), *files_for(args), "--mode", args.workload,
                              "--seconds", str(args.cpu_seconds)], capture_output=True, text=True)
        if out.returncode == 0:
            r = json.loads(out.stdout.strip().splitlines()[-1])
            cpu = {"value": r["value"], "unit": "Mpixel/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        else:
            cpu = {"value": None, "unit": "Mpixel/s", "cores": 0, "kind": "unavailable", "sample": out.stderr[-300:]}

    import webp_decoder_b200 as W
    from webp_decoder_b200 import parse as P

    # a non-default torch stream: its handle is what the library launches on, so torch.cuda.Event sees our kernels
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = W.Context(local, stream.cuda_stream)
    kernel_version = args.kernel or int(os.environ.get("VP8_GPU_KERNEL", "3") or 3)
    ctx.set_kernel(kernel_version)
    if args.warps or args.images_per_sm:
        ctx.set_tuning(args.warps, args.images_per_sm)
    if world > 1 and not os.environ.get("VP8_GPU_HOST_THREADS"):
        # the ranks of one node share its cores: each takes its share for the host side of the end-to-end call
        ctx.set_transport(True, max(1, (os.cpu_count() or 1) // world))

    # ---- inputs: parse the distinct frames once (host threads, pinned arenas), replicate to the batch size
    files = files_for(args)
    pf = P.parse_batch([Path(f).read_bytes() for f in files], pinned=True)
    nd = pf.n
    order = [i % nd for i in range(args.batch)]  # interleaved mix
    kfs = [pf.kfs[i] for i in order]
    frs = [pf.frames[i] for i in order]
    w, h = pf.kfs[0].width, pf.kfs[0].height
    px_step = sum(pf.kfs[i].width * pf.kfs[i].height for i in order)
    alg_bytes = sum(algorithmic_bytes(pf.kfs[i].width, pf.kfs[i].height, args.workload) for i in order)
    filtered = args.workload != "yuv"

    def step_resident(b):
        ctx.run(b, filtered, W.TIGHT)
        if args.workload == "ppm":
            ctx.rgb(b)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- host front end (m01 + m02 + m05 equivalent): what feeding the GPU from .webp bytes costs on this box's cores
    host_fe = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        datas = [Path(f).read_bytes() for f in files] * 16
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        tmp = P.parse_batch(datas, threads=threads)
        dt = time.perf_counter() - t0
        host_fe = {"value": sum(tmp.kfs[i].width * tmp.kfs[i].height for i in range(tmp.n)) / dt / 1e6, "unit": "Mpixel/s",
                   "threads": threads, "frames": tmp.n,
                   "note": "vp8_parse_batch (container + header + bool/token decode), one image per host thread; serial per frame, "
                           "so an end-to-end run that starts from .webp bytes is bound by this, not by the GPU stage"}
        tmp.free()

    # ---- (1) kernel stage, inputs resident in HBM
    batch = ctx.upload(kfs, frs)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step_resident(batch)
    ctx.kernel_time()
    sync_all()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident(batch)
    e1.record(stream)
    sync_all()
    sampler.mark_timed(tw0, time.perf_counter())
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = ctx.launches - l0
    kern_ms, kern_n = ctx.kernel_time()
    cfg = ctx.last_launch_config()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    job_ms = float(t.item())
    value = world * px_step * args.steps / (job_ms / 1e3) / 1e6

    # spot-check the resident result against the reference decoder's digests (outside the timed region)
    parity = None
    dg_path = ROOT / "bench_data" / "digests.json"
    if args.workload != "ppm" and dg_path.exists():
        import hashlib
        dg = json.loads(dg_path.read_text())
        buf, offs, sizes = ctx.download_i420(batch)
        key = "yuvf" if filtered else "yuv"
        idx = sorted(set([0, 1, 2, 3, args.batch - 1, args.batch // 2]) & set(range(args.batch)))
        parity = all(hashlib.sha256(buf[int(offs[i]):int(offs[i]) + int(sizes[i])]).hexdigest() == dg[Path(files[order[i]]).name][key] for i in idx)
        del buf
    batch.free()

    # ---- (2) end to end through the batch API: pinned host inputs -> H2D -> kernels -> D2H into pinned host memory
    e2e = None
    if not args.no_e2e:
        host_out = W.PinnedBuffer(ctx.decode_bytes(kfs, ppm=(args.workload == "ppm")))

        def step_e2e():
            # one public call: chunked upload -> kernels -> download, overlapped on the library's internal streams
            ctx.decode_into(kfs, frs, host_out.array, filtered=filtered, ppm=(args.workload == "ppm"), chunk=args.chunk)

        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        sync_all()
        h0, d0 = ctx.h2d_bytes, ctx.d2h_bytes
        n_e2e = max(1, min(args.steps, args.e2e_steps))
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e2.record(stream)
        for _ in range(n_e2e):
            step_e2e()
        e3.record(stream)
        sync_all()
        wall_ms = (time.perf_counter() - t0) * 1e3
        t2 = torch.tensor([max(wall_ms, e2.elapsed_time(e3))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": world * px_step * n_e2e / (float(t2.item()) / 1e3) / 1e6, "unit": "Mpixel/s",
               "h2d_bytes_per_step": (ctx.h2d_bytes - h0) // n_e2e, "d2h_bytes_per_step": (ctx.d2h_bytes - d0) // n_e2e,
               "steps": n_e2e, "ms_per_step": float(t2.item()) / n_e2e,
               "api": "vp8_gpu_decode_i420" if args.workload != "ppm" else "vp8_gpu_decode_ppm",
               "note": "host side = already-parsed frames in pinned memory; token decode (m03/m05) not included"}
        host_out.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = kern_ms / max(kern_n, 1)
        achieved = alg_bytes / (k_ms / 1e3) / 1e9 if kern_n else None
        line = {
            "metric": METRIC if args.workload == "yuvf" else f"decoded Mpixel/s (1080p -{args.workload} batch)",
            "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": job_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int16", "data": "synthetic (gen_ppm patterns through the reference encoder, --q 75 --loopfilter; no published dataset)",
            "config": {"workload": f"{args.batch} x {w}x{h} key frames per GPU, -{args.workload} "
                                   f"({'recon + loop filter' if filtered else 'recon only'}"
                                   f"{' + fancy-upsampled RGB' if args.workload == 'ppm' else ''}), q75 --loopfilter level 8, "
                                   "mix noise/rgbgrad/checker/diag in equal parts, interleaved",
                       "images_per_s": value * 1e6 / (w * h), "batch_per_gpu": args.batch, "distinct_frames": nd,
                       "cache": f"inputs {sum(820 * pf.frames[i].mb_total for i in order) / 1e9:.2f} GB per step, far larger than the 126 MB L2",
                       "launch": cfg, "parity_spot_check_vs_reference_digests": parity},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": ncu_traffic(args.workload, kernel_version), "peak_source": peak_src,
                         "kernel": {2: "vp8_mb_pairs", 3: "vp8_mb_lockstep"}[kernel_version],
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes, "launches_timed": kern_n},
            "cpu_baseline": cpu,
            "host_front_end": host_fe,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        emit(line)
    pf.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="yuvf", choices=["yuvf", "yuv", "ppm"])
    ap.add_argument("--batch", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--kernel", type=int, default=0, help="wavefront kernel: 1 warp per macroblock, 2 half-warp per macroblock, 3 = 2 with several images per CTA in lockstep (0 = library default)")
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--images-per-sm", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--chunk", type=int, default=0, help="frames per pipeline chunk of the end-to-end call (0 = library default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (args.impl == "b200" and args.gpus > 1 and world == 1):  # (the relaunching parent just passes its children's output on)
        claim_stdout()
    if args.impl == "reference":
        return reference_arm(args)
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", str(Path(__file__).resolve())] + sys.argv[1:]
        return subprocess.call(cmd)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
