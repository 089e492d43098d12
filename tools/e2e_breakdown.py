#!/usr/bin/env python
"""Host-side breakdown of one drop-in step (ParticleFilter.main_pf, noise='philox') at C1.

    python tools/e2e_breakdown.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import slam_robot_simu_b200 as mcl  # noqa: E402


def main():
    n, L = 1 << 20, 16
    lm = bench.landmarks_for(L)
    pf = mcl.ParticleFilter(100, n_particles=n, landmarks=lm, noise="philox", likelihood="moment", lazy_px=True,
                            force_resample=True)
    flt = pf.engine
    acc = {}

    def wrap(obj, name, label):
        fn = getattr(obj, name)

        def inner(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0
            return r
        setattr(obj, name, inner)

    wrap(flt, "step_async", "Filter.step_async (stage z, 3 launches, D2H: enqueue)")
    wrap(flt, "stats_end", "Filter.stats_end (wait for the GPU + derive)")
    for _ in range(50):
        pf.main_pf()
    acc.clear()
    K = 1000
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        pf.main_pf()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / K
    print(f"main_pf wall {wall * 1e6:.1f} us/step")
    tot = 0.0
    for k, v in acc.items():
        print(f"  {k:55s} {v / K * 1e6:7.1f} us")
        tot += v / K
    print(f"  {'rest of main_pf (truth + observation of the next frame, tuple)':55s} {(wall - tot) * 1e6:7.1f} us")


if __name__ == "__main__":
    main()
