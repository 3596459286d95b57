import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, circuits as cc, spice21_b200 as s21
from oracle import pyoracle as po
for ns in (21, 41, 61):
    ck, ic = cc.inverter_array(1, ns)
    r = po.Circuit(ck.to_text()).batch(1, 1, tstep=1e-11, tstop=2e-11, ic=ic, max_points=1)
    print(ns, 'oracle tran(1pt)', r['status'], r['iters'])
    for kern in (None, 'direct', 'coop', 'hybrid'):
        if kern: os.environ['S21_KERNEL'] = kern
        else: os.environ.pop('S21_KERNEL', None)
        c = ck.to_s21().elaborate(ic=ic)
        b = s21.Batch(c, 2)
        x, st, it = b.dcop()
        n = {name: k for k, name in enumerate(c.names)}
        print('   gpu dcop', kern, b.kernel_name(), st, it, 'r0s0..4', np.round(x[0, [n[f"r0s{k}"] for k in range(5)]], 3), b.setup_stats()['repaired_instances'])
        t, w, stt, itt = s21.Batch(ck.to_s21().elaborate(ic=ic), 1).tran(1e-11, 2e-11)
        print('   gpu tran', stt, itt)
