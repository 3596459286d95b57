"""Regenerates tests/golden/*.npz from the reference's checked-in golden waveforms.

Source: /root/reference/spice21/resources/test_*_tran.json — the snapshots the reference's own transient tests compare
against with abs tol 1e-6 (spice21/src/tests.rs:786-788, 942-945, 960-963, 1003-1006, 1046-1049, 1065-1068, 1372-1375).
The JSON text is converted losslessly (float64) to compressed .npz so the fixtures travel with the repo; nothing is
resampled or rounded. Run from the repo root:  python tests/golden/make_golden.py
"""
import glob
import json
import os

import numpy as np

SRC = "/root/reference/spice21/resources"
DST = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    for path in sorted(glob.glob(os.path.join(SRC, "test_*_tran.json"))):
        with open(path) as f:
            d = json.load(f)
        name = os.path.splitext(os.path.basename(path))[0]
        np.savez_compressed(os.path.join(DST, name + ".npz"), **{k: np.asarray(v, dtype=np.float64) for k, v in d.items()})
        print(name, {k: len(v) for k, v in d.items()})
