#!/usr/bin/env python
"""bench.py — the hot path's headline metric on B200.

Workload (BASELINE.json configs[1]): DC operating point of a Mos1 differential-pair amplifier
(N = 9 unknowns, 8 devices), 8192 Monte-Carlo instances per GPU (vt0/kp per transistor, g per load resistor). One
"step" = one batched dcop of all instances from a cold start (x = 0, fresh device state).

Metric: batched Newton iterations / second = sum over instances of the iterations that reached the linear solve,
divided by device time (CUDA events on the launch stream, max over ranks). Multi-GPU: the batch shards by instance
(weak scaling: 8192 instances per rank, no data-path collective; NCCL only gathers iteration counts and status flags).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference` times the reference algorithm's CPU restatement (oracle/, C++; the Rust original cannot be built in
this image) on the host cores for the same workload.
"""
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

B_PER_GPU = 8192
METRIC = "batched_newton_iters_per_sec"
UNIT = "newton_iters/s"
WORKLOAD = "C2: Mos1 diff-pair dcop x 8192 Monte-Carlo instances per GPU (N=9, 8 devices)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes_per_iteration(n, nnz_a, nnz_lu, per_inst_cols):
    """SURVEY.md §8(d) B_iter for this circuit, real arithmetic (w = 8): 2 Mos1 + 2 R + 1 I + 3 V.
    B_eval counts terminals gathered, per-instance parameter columns and Mos1 state (9 read + 9 written)."""
    w = 8
    b_eval = 2 * (48 + 8 * (9 + 9)) + 2 * 16 + 8 * per_inst_cols
    b_asm = w * (nnz_a + n)
    b_res = w * (nnz_a + 3 * n)
    b_lu = 2 * w * nnz_lu
    b_solve = w * (nnz_lu + 3 * n)
    b_conv = w * (2 * n + 2 * n)
    return {"eval": b_eval, "asm": b_asm, "res": b_res, "lu": b_lu, "solve": b_solve, "conv": b_conv,
            "total": b_eval + b_asm + b_res + b_lu + b_solve + b_conv}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def run_reference(args):
    """The reference arm: the CPU restatement of the reference algorithm on all host cores, same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import circuits as cc
    from oracle import pyoracle as po
    po.build()
    B = B_PER_GPU
    cores = host_cores()
    ck, ovr = cc.diffpair(), cc.diffpair_mc(B)
    oc = po.Circuit(ck.to_text())
    secs, iters = [], 0
    for k in range(args.warmup + args.steps):
        r = oc.batch(0, B, overrides=ovr, nthreads=cores, want_x=False)
        if k >= args.warmup:
            secs.append(r["seconds"])
            iters = int(r["iters"].sum())
    total = float(np.sum(secs))
    value = iters * args.steps / total
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "instances": B, "newton_iters_per_step": iters, "timed": "Solver::solve only (solvers pre-built)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"all {B} instances per step, {args.steps} steps; C++ restatement of the Rust reference (oracle/)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def aggregate(dist, world, device, times, iters, status, tran_ms):
    """Every collective of the measurement in ONE place, executed by EVERY rank in the same order (a collective that only
    some ranks reach hangs the job): MAX over ranks of the timed regions, the gather of per-instance iteration counts and
    status flags (the path's only data exchange), and the transient metric's time — dropped on all ranks if
    any rank could not measure it (tran_ms None). Returns (times, total_iters, all_ok, tran_ms). Runs on gloo/CPU in
    tests/test_shard.py."""
    import torch
    from spice21_b200.shard import gather_instances
    if dist is None or world == 1:
        return list(times), int(np.sum(iters)), bool(np.all(np.asarray(status) == 0)), tran_ms
    t = torch.tensor(list(times) + [tran_ms if tran_ms is not None else 0.0], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = torch.tensor([0.0 if tran_ms is None else 1.0], dtype=torch.float64, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    n = len(iters)
    it_all = gather_instances(np.asarray(iters), n * world, device=device)
    st_all = gather_instances(np.asarray(status), n * world, device=device)
    vals = t.tolist()
    return vals[:-1], int(it_all.sum()), bool(np.all(st_all == 0)), (vals[-1] if ok.item() > 0.5 else None)


KERNEL_NAMES = {
    "hybrid": "s21::k_hyb<double, dcop> (hybrid cooperative Newton kernel, kernels/hybrid.cu)",
    "jit-team": "k_jit (run-time specialised team kernel: 8/16 lanes per instance, rows in registers; host/jit_team.hpp)",
    "jit-thread": "k_jit (run-time specialised kernel, one thread per instance; host/jit.hpp)",
    "coop": "s21::k_coop<double, dcop> (cooperative Newton kernel, kernels/coop.cu)",
    "direct": "s21::k_dcop (one thread per instance, kernels/newton.cu)",
}


def run_ours(args):
    import torch
    import circuits as cc
    import spice21_b200 as s21

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if s21.cuda_device_count() < 1 or not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        import datetime
        # a collective that cannot complete fails after two minutes instead of hanging the job
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    B = B_PER_GPU
    ck = cc.diffpair()
    ovr = cc.diffpair_mc(B, first_instance=rank * B)  # each rank owns its own Monte-Carlo samples
    c = ck.to_s21().elaborate()
    batch = s21.Batch(c, B, device=local)
    stream = torch.cuda.Stream()  # a real (non-legacy) stream: the library launches on it, the events below time it
    torch.cuda.set_stream(stream)
    batch.set_stream(stream.cuda_stream)
    for k, v in ovr.items():
        batch.override(k, v)
    h2d = batch.sync_params(force_upload=True)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def step_device():
        batch.reset()
        batch.dcop_device()

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    launches_per_step = batch.stats()["launches"]  # kernels one reset + dcop_device enqueues (read() adds a layout kernel, untimed here)
    x, status, iters = batch.read()
    assert np.all(status == 0), "non-converged instances in the benchmark batch"
    iters_per_step = int(iters.sum())
    st = batch.stats()
    kname = batch.kernel_name()

    sampler = ClockSampler(local)
    sampler.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()  # evict the batch from L2 between timed steps (not timed)
        ev[k][0].record(stream)
        batch.reset()
        kev[k][0].record(stream)
        batch.dcop_device()
        kev[k][1].record(stream)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    step_ms = sum(a.elapsed_time(b) for a, b in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b in kev)

    # end to end through the public API with host buffers: H2D of the per-instance parameter pool from pinned memory,
    # reset, solve, D2H of x / status / iteration counts — every step.
    e2e_steps = max(args.steps, 5)
    for _ in range(2):
        batch.sync_params(force_upload=True); batch.reset(); batch.dcop_view()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        batch.sync_params(force_upload=True)
        batch.reset()
        x, status, iters = batch.dcop_view()  # results in the library's pinned host buffer (x[B][N], status[B], iters[B])
        e2e_check = float(x[-1, 0]) + int(iters[-1])  # the host reads the step's result
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert np.all(status == 0) and int(iters.sum()) == iters_per_step and np.isfinite(e2e_check)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    d2h = B * c.n_vars * 8 + 3 * B * 4

    tran = tran_metric(s21, cc, local, stream, rank)

    (tot_ms, tot_kern_ms, tot_e2e), tot_iters, all_ok, tran_ms = aggregate(
        dist, world, "cuda", (step_ms, kern_ms, e2e_s), iters, status, tran["ms"] if tran is not None else None)
    assert all_ok, "non-converged instances on some rank"
    if tran is not None:
        tran = dict(tran, ms=tran_ms) if tran_ms is not None else None
    if rank == 0:
        peak, peak_src = measured_peak()
        per_inst_cols = (h2d // 8 - 0) // ((B + 31) // 32 * 32) if h2d else 0
        bi = algorithmic_bytes_per_iteration(st["n"], st["nnz_a"], st["nnz_lu"], per_inst_cols)
        kernel_ms_avg = tot_kern_ms / args.steps
        traffic, traffic_src = None, None  # DRAM bytes per launch of this kernel from the committed ncu --set full capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kname) or {}
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except (OSError, ValueError):
            pass
        achieved = bi["total"] * iters_per_step / (kernel_ms_avg * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": tot_iters * args.steps / (tot_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "instances_per_gpu": B, "newton_iters_per_step_per_gpu": iters_per_step,
                       "n": st["n"], "nnz_a": st["nnz_a"], "nnz_lu": st["nnz_lu"], "stamp_slots": st["stamps"],
                       "l2": "256 MiB flush write between timed steps (untimed)", "step": "reset (cold start) + batched dcop kernel"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "kernel": KERNEL_NAMES.get(kname, kname), "kernel_ms": kernel_ms_avg,
                         "algorithmic_bytes_per_iteration": bi},
            "e2e": {"value": tot_iters * e2e_steps / tot_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * tot_e2e / e2e_steps, "steps": e2e_steps,
                    "path": "s21_batch_sync_params(force: H2D of the parameter pool from pinned memory) + s21_batch_reset + s21_batch_dcop_view (D2H of x/status/iters into pinned host memory)"},
            "gpu_launches": args.steps * launches_per_step,
            "clocks": sampler.summary(),
        }
        if tran is not None:
            line["tran"] = {"metric": "tran_timepoints_per_sec", "value": world * tran["instances"] * tran["timepoints"] / (tran["ms"] * 1e-3),
                            "unit": "timepoints/s", "newton_iters_per_sec": world * tran["iters"] / (tran["ms"] * 1e-3),
                            "ms_per_transient": tran["ms"], "workload": tran["workload"], "kernel": tran["kernel"]}
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(ck, ovr, B)
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def tran_metric(s21, cc, local, stream, rank, B=B_PER_GPU):
    """Second half of BASELINE.json's metric: transient timepoints/s. Workload = configs[0]'s circuit (the reference's Mos1
    CMOS ring oscillator, tests.rs:889-912) as a supply sweep of B instances per GPU, 200 fixed Backward-Euler steps in one
    launch (OP, IC release and the whole time loop on the device); device time from the library's CUDA events."""
    try:
        ro = cc.cmos_ro3(cc.add_mos1_defaults)
        b = s21.Batch(ro.to_s21().elaborate(ic={"1": 0.0}), B, device=local)
        b.set_stream(stream.cuda_stream)
        b.override("V:v1:dc", np.linspace(0.9, 1.1, B) + 1e-4 * rank)
        save = np.array([0, 1, 2], dtype=np.int32)
        best, out = None, None
        for _ in range(3):
            b.reset()
            t, w, st, it = b.tran(1e-11, 2e-9, save=save)
            ms = b.stats()["device_ms"]
            if best is None or ms < best:
                best, out = ms, (len(t), int(it.sum()), bool(np.all(st == 0)))
        assert out[2], "non-converged instances in the transient batch"
        return {"ms": best, "instances": B, "timepoints": out[0] - 1, "iters": out[1], "kernel": b.kernel_name(),
                "workload": f"C1 circuit (Mos1 CMOS ring oscillator, N=7) x {B} supply-sweep instances per GPU, tstep 1e-11, {out[0] - 1} points"}
    except Exception as e:  # the headline metric must not depend on the secondary one
        print(f"[bench] transient metric skipped: {e!r}", file=sys.stderr)
        return None


def cpu_baseline(ck, ovr, B):
    from oracle import pyoracle as po
    po.build()
    cores = host_cores()
    oc = po.Circuit(ck.to_text())
    oc.batch(0, B, overrides=ovr, nthreads=cores, want_x=False)  # warm-up
    r = oc.batch(0, B, overrides=ovr, nthreads=cores, want_x=False)
    r1 = oc.batch(0, min(B, 2048), overrides={k: v[:2048] for k, v in ovr.items()}, nthreads=1, want_x=False)
    return {"value": float(r["iters"].sum() / r["seconds"]), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"all {B} instances, one pass, {cores} threads; Solver::solve only",
            "single_core_value": float(r1["iters"].sum() / r1["seconds"])}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
