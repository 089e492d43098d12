#!/usr/bin/env python
"""Resampling time when the weights have collapsed onto a few particles (one source owns millions of slots).

    python tools/degenerate_case.py [n]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slam_robot_simu_b200 as mcl  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 25
    rng = np.random.RandomState(1)
    pos = rng.normal(0, 1, (3, n)).astype(np.float32)
    for heavy in (1, 7, 1000):
        lw = np.full(n, -60.0, dtype=np.float32)
        idx = rng.randint(0, n, heavy)
        lw[idx] = rng.normal(0, 1, heavy).astype(np.float32)
        f = mcl.Filter(n_particles=n, landmarks=[[0.0, 0.0]])
        ts = []
        for rep in range(4):
            f.set_particles(pos[0], pos[1], pos[2], lw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f.resample(ofs=0.37 / n, force=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        anc = f.ancestors()[0]
        ok = np.all(np.diff(anc.astype(np.int64)) >= 0) and set(np.unique(anc)) <= set(idx.tolist())
        print(f"n={n} heavy={heavy}: resample (scan + gather) {min(ts):.3f} ms, ancestors consistent={ok}")
        del f


if __name__ == "__main__":
    main()
