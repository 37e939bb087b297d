#!/usr/bin/env python
"""Headline benchmark: clip-frames/s of DragPoser's per-frame latent optimisation at fixed
100 iterations, 4096 synthetic clips per GPU, 6 trackers (BASELINE.json metric / configs[3]).

A "step" is one frame of the hot path for every clip of the batch: temporal-predictor
target (window 0 => every frame) + 100 x {decode, FK, tracker loss, adjoint, decoder
backward, latent Adam} + frame epilogue.  `value` has the tracker streams resident in
HBM; `e2e` goes through the host-buffer API (H2D of the step's targets and D2H of the
poses inside the timed region).  `--impl reference` times the CPU oracle port of the
reference loop (the reference is Python and cannot travel to the GPU box) on all host
cores.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden", "model_dancedb.npz")
FLOP_PER_CLIP_ITER = 35520.0      # decoder fwd + bwd-data GEMMs, folded 24->40->60->92
HBM_BYTES_PER_CLIP_FRAME = 1300.0  # algorithmic state + tracker + result bytes
MAX_ITER = 100


KERNEL_NAMES = {
    1: "dp_frame_simt_kernel (persistent per-frame loop; fp32 CUDA-core decoder)",
    3: "dp_frame_tc16_kernel (persistent per-frame loop; decoder GEMMs on tcgen05, fp16x2 split, weights in tensor memory)",
}
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the ncu --set full capture whose
# summary is committed under profiles/ (keyed by decoder path, clips per GPU, tracker config); null for other configurations
NCU_DRAM_BYTES_PER_LAUNCH = {(3, 4096, "6"): 2955008}


def fixed_opts(cfg):
    return dict(stop_eps_pos=-1.0, stop_eps_rot=-1.0, max_iter=MAX_ITER, min_loss_incr=-float("inf"), learning_rate=1e-2,
                lambda_rot=1, lambda_temporal=cfg.lambda_temporal, temporal_future_window=cfg.temporal_future_window,
                joint_adjustment_indices=cfg.joint_adjustment, joint_adjustment_weight=cfg.joint_adjustment_weight)


def get_config(name):
    from dragposer_b200 import synthetic

    return synthetic.config_3_trackers() if name == "3" else synthetic.config_6_trackers()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device, n_gpus=1):
        self.rows, self.proc = [], None
        self.device = str(device) if n_gpus == 1 else ",".join(str(i) for i in range(n_gpus))  # every GPU of the job

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    """One single-threaded worker = one clip, like the reference (B = 1, CPU, fp32)."""
    idx, n_warm, n_timed, tracker_cfg = args
    import torch

    torch.set_num_threads(1)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dragposer_port as port
    from dragposer_b200 import model, synthetic

    cfg = get_config(tracker_cfg)
    pm = model.load_folded_npz(GOLDEN)
    npz = np.load(GOLDEN)
    wl = synthetic.make_workload(pm, npz["offsets"], cfg, 1, n_warm + n_timed, first_clip=idx)
    drag = port.PortDragPose(port.PortWeights(npz), model.random_temporal_state(2222))
    drag.set_initial_state(wl["latent0"], np.zeros((1, 3)), [[1.0, 0, 0, 0]], np.zeros((1, 6)))
    o = fixed_opts(cfg)
    kw = dict(stop_eps_pos=o["stop_eps_pos"], stop_eps_rot=o["stop_eps_rot"], max_iter=o["max_iter"], min_loss_incr=o["min_loss_incr"],
              learning_rate=o["learning_rate"], lambda_rot=1.0, lambda_temporal=o["lambda_temporal"],
              temporal_future_window=o["temporal_future_window"], joint_adjustment=o["joint_adjustment_indices"],
              joint_adjustment_weight=o["joint_adjustment_weight"])
    t0 = None
    for t in range(n_warm + n_timed):
        if t == n_warm:
            t0 = time.perf_counter()
        drag.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], **kw)
    return time.perf_counter() - t0


# ----------------------------------------------------------------------------- single-frame latency (second half of the metric)
LAT_POS = (-2.6648, 0.9977, 3.7518)
LAT_ROT = (0.6381, 0.0078, -0.7698, 0.0110)


def dll_latency(calls=1000):
    """p50 / p90 wall-clock of drag_pose through the DragPoserDLL C ABI, B = 1, host buffers in and out:
    Unity parameters (MaxIter 5 / 10, lr 0.01, window 16) and the offline 100 iterations; targets of DragPoserDLL/main.cpp."""
    import ctypes as C
    import tempfile

    from dragposer_b200 import build, export_model

    class F3(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    class Q4(C.Structure):
        _fields_ = [("w", C.c_float), ("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    class F2(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float)]

    lib = C.CDLL(build.build_all()[1])
    V = C.c_void_p
    lib.init_drag_poser.restype = V
    lib.set_reference_skeleton.argtypes = [V, C.c_char_p]
    lib.load_models.argtypes = [V, C.c_char_p]
    lib.set_mask_and_weights.argtypes = [V, C.POINTER(C.c_float), C.POINTER(F2)]
    lib.init_drag_model.argtypes = [V, F3, Q4]
    lib.set_optim_params.argtypes = [V, C.c_float, C.c_float, C.c_int, C.c_float]
    lib.set_lambdas.argtypes = [V, C.c_float, C.c_float, C.c_int]
    lib.set_global_pos.argtypes = [V, F3]
    lib.drag_pose.argtypes = [V, C.c_int, C.POINTER(F3), C.POINTER(Q4), C.POINTER(Q4), C.POINTER(F3)]
    lib.destroy_drag_poser.argtypes = [V]
    lib.dp_last_status.argtypes = [V]
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_rundrag.npz"))
    out = {"boundary": "DragPoserDLL C ABI drag_pose, B = 1, 6 trackers, host buffers in/out, window 16", "calls": calls,
           "early_stop": "stopEpsPos 1e-4, stopEpsRot 1e-2 as in the reference session (the loop ends at MaxIter or at the thresholds)"}
    with tempfile.TemporaryDirectory() as d:
        export_model.export(GOLDEN, os.path.join(d, "model.dpm"), allow_random_temporal=True)
        for max_iter in (5, 10, 100):
            h = lib.init_drag_poser()
            lib.set_reference_skeleton(h, os.path.join(ROOT, "tests", "golden", "skeleton22.bvh").encode())
            lib.load_models(h, d.encode())
            mask = (C.c_float * 22)(*g["mask"].tolist())
            weights = (F2 * 22)(*[F2(*w) for w in g["weights"].tolist()])
            lib.init_drag_model(h, F3(*LAT_POS), Q4(*LAT_ROT))
            if lib.dp_last_status(h) != 0:
                raise SystemExit("bench.py: DragPoserDLL session failed to start")
            T = g["tgt_pos"].shape[0]
            pos = [(F3 * 6)(*[F3(*p) for p in g["tgt_pos"][t].tolist()]) for t in range(T)]
            rot = (Q4 * 6)(*[Q4(*q) for q in g["tgt_quat"].tolist()])
            res, gp = (Q4 * 22)(), (F3 * 1)()
            n = calls if max_iter < 100 else max(100, calls // 5)
            times = []
            for i in range(n + 20):  # the per-frame call sequence of DragPoser.cs:137-173
                lib.set_mask_and_weights(h, mask, weights)
                lib.set_optim_params(h, 0.01 * 0.01, 0.01, max_iter, 0.01)
                lib.set_lambdas(h, 1, 0.02, 16)
                t0 = time.perf_counter()
                lib.drag_pose(h, 6, pos[i % T], rot, res, gp)
                times.append(time.perf_counter() - t0)
                lib.set_global_pos(h, gp[0])
            lib.destroy_drag_poser(h)
            t = np.sort(np.array(times[20:])) * 1e3
            out[f"max_iter_{max_iter}"] = {"p50_ms": float(t[len(t) // 2]), "p90_ms": float(t[int(len(t) * 0.9)])}
    out["p50_ms"] = out["max_iter_5"]["p50_ms"]
    return out


def port_latency(calls=30):
    """The reference arm of the latency metric: the oracle port's DragPose.run, B = 1, one core, same parameters."""
    torch_threads = 1
    import torch

    torch.set_num_threads(torch_threads)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dragposer_port as port
    from dragposer_b200 import model, synthetic

    cfg = synthetic.config_6_trackers()
    pm = model.load_folded_npz(GOLDEN)
    npz = np.load(GOLDEN)
    out = {"boundary": "oracle port of DragPose.run, B = 1, 6 trackers, one core, window 16", "calls": calls,
           "early_stop": "stopEpsPos 1e-4, stopEpsRot 1e-2 (the loop ends at MaxIter or at the thresholds)"}
    for max_iter in (5, 10):
        wl = synthetic.make_workload(pm, npz["offsets"], cfg, 1, calls + 3)
        drag = port.PortDragPose(port.PortWeights(npz), model.random_temporal_state(2222))
        drag.set_initial_state(wl["latent0"], np.zeros((1, 3)), [[1.0, 0, 0, 0]], np.zeros((1, 6)))
        times = []
        for t in range(calls + 3):
            t0 = time.perf_counter()
            drag.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], stop_eps_pos=1e-4, stop_eps_rot=1e-2, max_iter=max_iter,
                     min_loss_incr=1e-5, learning_rate=1e-2, lambda_rot=1.0, lambda_temporal=0.02, temporal_future_window=16,
                     joint_adjustment=None, joint_adjustment_weight=0.0)
            times.append(time.perf_counter() - t0)
        t = np.sort(np.array(times[3:])) * 1e3
        out[f"max_iter_{max_iter}"] = {"p50_ms": float(t[len(t) // 2]), "p90_ms": float(t[int(len(t) * 0.9)])}
    out["p50_ms"] = out["max_iter_5"]["p50_ms"]
    return out


def cpu_reference(n_warm, n_timed, tracker_cfg, cores=None):
    """All host cores, one clip per core; returns (clip-frames/s, cores, sample text)."""
    import multiprocessing as mp

    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        times = pool.map(_cpu_worker, [(i, n_warm, n_timed, tracker_cfg) for i in range(cores)])
    value = cores * n_timed / max(times)
    sample = (f"{cores} clips x {n_timed} frames (one single-threaded process per core, B=1 like the reference), "
              f"{MAX_ITER} fixed iterations, {tracker_cfg} trackers, predictor every frame, after {n_warm} warm-up frames")
    return value, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = get_config(args.trackers)
    n_warm, n_timed = max(1, min(args.warmup, 1)), max(1, min(args.steps, 8))
    value, cores, sample = cpu_reference(n_warm, n_timed, args.trackers)
    line = {
        "impl": "reference", "metric": "clip-frames/sec (fixed 100 opt iters)", "value": value, "unit": "clip-frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cores / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, args.clips * args.gpus),
        "cpu_baseline": {"value": value, "unit": "clip-frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clip-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_latency:
        line["latency"] = port_latency()
    print(json.dumps(line))


def workload_config(args, cfg, clips_total):
    return {"workload": f"{args.trackers}-tracker synthetic streams (SURVEY 8d), model_dancedb weights, seed-2222 random-init predictor, "
                        f"window {cfg.temporal_future_window}, fixed {MAX_ITER} iterations",
            "clips_per_gpu": args.clips, "clips_total": clips_total, "trackers": int(args.trackers), "max_iter": MAX_ITER,
            "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"clip-shard x{args.gpus}"}


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from dragposer_b200 import dist as dpdist
    from dragposer_b200 import model, synthetic
    from dragposer_b200.engine import BatchedDragPose, RunOptions

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("DP_BENCH_REVERSE_DEVICES"):  # diagnostic: is a slow rank a slow GPU or a slow process?
        local = world - 1 - local
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = get_config(args.trackers)
    pm = model.load_folded_npz(GOLDEN)
    offsets = np.load(GOLDEN)["offsets"]
    tm = model.temporal_from_state(model.random_temporal_state(2222))
    B, K, W = args.clips, args.steps, args.warmup
    n_total = B * world
    T = W + K
    wl = synthetic.make_workload(pm, offsets, cfg, B, T, first_clip=rank * B, variable_mask=(args.trackers == "3"))
    eng = BatchedDragPose(pm, offsets, tm, B, device=local)
    eng.set_initial_state(wl["latent0"], np.zeros((B, 3)), np.tile([[1.0, 0, 0, 0]], (B, 1)), np.zeros((B, 6)))
    opts = RunOptions(decoder_path=args.decoder_path, **fixed_opts(cfg))
    E = wl["tgt_pos"].shape[2]
    variable = "n_ee" in wl
    d_tp = torch.from_numpy(wl["tgt_pos"]).to(dev)
    d_tr = torch.from_numpy(wl["tgt_rot"]).to(dev)
    if variable:
        d_j, d_w = torch.from_numpy(wl["joints_tb"]).to(dev), torch.from_numpy(wl["weights_tb"]).to(dev)
        d_ne = torch.from_numpy(wl["n_ee"]).to(dev)
    else:
        d_j = torch.from_numpy(wl["joints"].astype(np.int32)).to(dev)
        d_w = torch.from_numpy(wl["weights"]).to(dev)
        d_ne = None
    d_pose = torch.empty((T, B, 88), dtype=torch.float32, device=dev)
    d_gpos = torch.empty((T, B, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    work_stream = torch.cuda.Stream(device=dev)  # everything timed runs (and is timed) on this stream
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream

    def step(t):
        eng.run_frames_device(1, d_tp[t], d_tr[t], d_j[t] if variable else d_j, d_w[t] if variable else d_w, d_pose[t], d_gpos[t],
                              n_ee=d_ne[t] if variable else None, shared=not variable, ee_stride=E, stream=stream, options=opts)

    for t in range(W):
        step(t)
    if world > 1:  # warm the collective too (NCCL connects its channels on the first call of a shape)
        dpdist.gather_frames(d_pose[W:W + K], d_gpos[W:W + K], n_total)
    clocks = ClockSampler(local, world)
    if rank == 0:
        clocks.start()  # before the barrier: every rank must enter the timed region at the same moment
    eng.set_profiling(True)
    l0 = eng.launch_count()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    evs = []
    for k in range(K):
        flush.zero_()  # L2 flush between timed steps (outside the event pair)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(W + k)
        b.record()
        evs.append((a, b))
    if world > 1:  # the only collective of the job: ONE gather of the K frames' result rows, in rank order, inside the timed region
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dpdist.gather_frames(d_pose[W:W + K], d_gpos[W:W + K], n_total)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = eng.launch_count() - l0
    ms_pred, ms_frame, nprof = eng.profile()
    eng.set_profiling(False)
    clk = clocks.stop() if rank == 0 else None
    ms_by_rank, parts_by_rank = [ms / K], None
    if world > 1:
        dist.barrier()
        every = [torch.zeros(3, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(every, torch.tensor([ms, ms_frame / max(nprof, 1), ms_pred / max(nprof, 1)], dtype=torch.float64, device=dev))
        ms_by_rank = [float(t[0].item()) / K for t in every]
        parts_by_rank = [[round(float(t[1].item()), 4), round(float(t[2].item()), 4)] for t in every]
        ms = max(float(t[0].item()) for t in every)  # the job is as fast as its slowest rank
    value = n_total * K / (ms * 1e-3)

    # ---- e2e: HOST arrays through the public API (BatchedDragPose.run_frames), every frame's staging + H2D + D2H inside the
    # timed region; the call double-buffers them so they overlap the kernels of the neighbouring frames
    h_tp, h_tr = wl["tgt_pos"], wl["tgt_rot"]
    run_kw = dict(options=opts)

    host_out = (np.zeros((K, B, 88), np.float32), np.zeros((K, B, 3), np.float32))  # result arrays of the caller, reused

    def run_host(t0, t1):
        out = host_out if t1 - t0 == K else None
        if variable:
            return eng.run_frames(h_tp[t0:t1], h_tr[t0:t1], wl["joints_tb"][t0:t1], wl["weights_tb"][t0:t1], n_ee=wl["n_ee"][t0:t1],
                                  out=out, **run_kw)
        return eng.run_frames(h_tp[t0:t1], h_tr[t0:t1], wl["joints"], wl["weights"], out=out, **run_kw)

    run_host(0, min(W, 2))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    poses, gposes = run_host(W, W + K)
    last = (poses[-1], gposes[-1])
    if world > 1:  # gather of the last frame's poses across the ranks (NCCL)
        dpdist.gather_results(torch.from_numpy(last[0]).to(dev), torch.from_numpy(last[1]).to(dev), n_total)
        torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = n_total * K / e2e_s
    h2d = B * E * (3 + 9) * 4 + (B * E * (1 + 2) * 4 + B * 4 if variable else E * 12)
    d2h = B * (88 + 3) * 4

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        tf32_peak = (peaks.get("bf16_tflops_sustained") or 1400.0) / 2.0
        hbm_peak = peaks.get("hbm_gbs") or 6650.0
        which = "of measured (bf16 sustained / 2 = TF32)" if peaks else "of fallback"
        frame_ms = ms_frame / max(nprof, 1)
        achieved_tf = B * MAX_ITER * FLOP_PER_CLIP_ITER / (frame_ms * 1e-3) / 1e12
        achieved_gbs = B * HBM_BYTES_PER_CLIP_FRAME / (frame_ms * 1e-3) / 1e9
        line = {
            "metric": "clip-frames/sec (fixed 100 opt iters)", "value": value, "unit": "clip-frames/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, cfg, n_total),
            "roofline": {"bound": "tensor", "kernel": KERNEL_NAMES[eng.last_decoder_path()],
                         "achieved": achieved_tf, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved_tf / tf32_peak,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get((eng.last_decoder_path(), B, args.trackers)), "peak_source": which, "kernel_ms_per_launch": frame_ms,
                         "predictor_ms_per_step": ms_pred / max(nprof, 1),
                         "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak},
                         "note": "dependency-latency bound path (SURVEY 8d): both fractions are small by construction"},
            "e2e": {"value": e2e_value, "unit": "clip-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clk,
        }
        if world > 1:
            line["ms_per_step_by_rank"] = [round(v, 4) for v in ms_by_rank]
            line["frame_kernel_and_predictor_ms_by_rank"] = parts_by_rank
        if world == 1 and not args.no_latency:
            eng.close()  # the DLL session opens its own engine
            line["latency"] = dll_latency()
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_reference(1, 6, args.trackers)
            line["cpu_baseline"] = {"value": v, "unit": "clip-frames/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=4096, help="clips per GPU (weak scaling)")
    ap.add_argument("--trackers", default="6", choices=["6", "3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the B = 1 C-ABI latency measurement")
    ap.add_argument("--decoder-path", type=int, default=0, choices=[0, 1, 3], help="0 auto, 1 fp32 CUDA-core decoder, 3 tcgen05 fp16x2")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
