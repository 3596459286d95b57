"""Sweep the cooperative kernel's launch geometry (instances per CTA x threads) on the bench workload (C2).
Prints device ms per batched dcop (CUDA events on the launch stream, L2 flushed between runs)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import circuits as cc  # noqa: E402
import spice21_b200 as s21  # noqa: E402

B = int(os.environ.get("SWEEP_B", "8192"))
ck, ovr = cc.diffpair(), cc.diffpair_mc(B)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ref = None
for kernel, gi, th in [("direct", 0, 0), ("hybrid", 0, 0)] + [("coop", g, t) for g in (32, 16) for t in (128, 256)]:
    os.environ["S21_KERNEL"] = kernel
    if gi:
        os.environ["S21_COOP_GI"], os.environ["S21_COOP_THREADS"] = str(gi), str(th)
    b = s21.Batch(ck.to_s21().elaborate(), B)
    b.set_stream(stream.cuda_stream)
    for k, v in ovr.items():
        b.override(k, v)
    ms = []
    for rep in range(8):
        flush.zero_()
        b.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        b.dcop_device()
        e1.record(stream)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    x, st, it = b.read()
    if ref is None:
        ref = (x, st, it)
    same = np.array_equal(x, ref[0]) and np.array_equal(st, ref[1]) and np.array_equal(it, ref[2])
    print(f"{kernel:6s} gi={gi:2d} threads={th:3d}  ms={np.median(ms[3:]):.4f}  iters={int(it.sum())}  bit_identical_to_direct={same}", flush=True)
