"""Config C4 on the GPU: BSIM4 ring oscillator transient x VDD/temperature sweep (tests/circuits.py::bsim4_ring).
usage: python scripts/run_c4.py [B] [n_stages] [n_points] [oracle_instances]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import circuits as cc  # noqa: E402
import spice21_b200 as s21  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n_stages = int(sys.argv[2]) if len(sys.argv) > 2 else 21
n_points = int(sys.argv[3]) if len(sys.argv) > 3 else 500
n_oracle = int(sys.argv[4]) if len(sys.argv) > 4 else 0
tstep = 1e-10
ic_every = int(os.environ.get("RUNC4_IC_EVERY", "0"))  # further initial conditions every so many stages (tests/circuits.py bsim4_ring)
ck, ic = cc.bsim4_ring(n_stages, ic_every=ic_every)
ovr = cc.c4_sweep(B)
c = ck.to_s21().elaborate(ic=ic)
save = [c.names.index("s1"), c.names.index(f"s{n_stages // 2}"), c.names.index("vsup")]
b = s21.Batch(c, B)
for k, v in ovr.items():
    b.override(k, v)
# experiments (profiles/r02E_c4_modes.txt): what bench.py does around the same batch
if os.environ.get("RUNC4_TORCH"):
    import torch
    torch.cuda.set_device(0)
    _stream = torch.cuda.Stream()
    torch.cuda.set_stream(_stream)
    if os.environ.get("RUNC4_TORCH") == "2":
        _flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    b.set_stream(_stream.cuda_stream)
for rep in range(int(os.environ.get("RUNC4_REPS", "2"))):
    if os.environ.get("RUNC4_FORCE"):
        b.sync_params(force_upload=True)
    b.reset()
    t0 = time.time()
    t, wave, status, iters = b.tran(tstep, n_points * tstep, save=save)
    wall = time.time() - t0
    st = b.stats()
    print(f"rep {rep}: B={B} stages={n_stages} N={len(c.names)} points={len(t)} ok={int(np.sum(status == 0))}/{B} iters={int(iters.sum())} "
          f"device_ms={st.get('device_ms'):.2f} wall_s={wall:.3f} timepoints/s={B * (len(t) - 1) / wall:.3e} iters/s={iters.sum() / wall:.3e}", flush=True)
print("stats", st)
if n_oracle:
    from oracle import pyoracle as oracle
    t0 = time.time()
    sub = {k: v[:n_oracle] for k, v in ovr.items()}
    o = oracle.Circuit(ck.to_text()).batch(1, n_oracle, overrides=sub, tstep=tstep, tstop=n_points * tstep, ic=ic, nthreads=16)
    print(f"oracle {n_oracle} instances: {time.time() - t0:.2f}s wall, solve seconds {o['seconds']:.2f}, iters {int(o['iters'].sum())}, ok {int(np.sum(o['status'] == 0))}")
    ow = o["x"][:, :, save]
    print("max |gpu - oracle| =", float(np.max(np.abs(wave[:n_oracle] - ow))), " iters equal:", float(np.mean(iters[:n_oracle] == o["iters"])))
