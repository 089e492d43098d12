#!/usr/bin/env python
"""Phase timelines of k_scan and k_step (mcl_debug_timeline): where the fixed time of a small step goes.

    MCL_TIMELINE=1 python slam-robot_simu_b200/build.py --force     # stamps are compiled out by default
    python tools/timeline.py [--n 1048576] [--workload c1]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

STEP = ["start", "scalars", "loop", "flush", "partial"]
SCAN = ["start", "prologue", "tile-scan", "look-back", "counts", "filled", "prep-done", "-", "-", "-", "-", "entry"]


def show(title, t, names, t0):
    rel = (t - t0) / 1e3
    print(f"  {title}: {t.shape[0]} CTAs")
    for s, nm in enumerate(names):
        col = rel[:, s][(t[:, s] > 0) & (rel[:, s] > -50.0) & (rel[:, s] < 1000.0)]
        if col.size:
            print(f"    {nm:10s} min {col.min():7.2f}  median {np.median(col):7.2f}  max {col.max():7.2f}  (n={col.size})")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c1")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--back-to-back", action="store_true", help="the measured step follows another step of the SAME filter "
                    "without a synchronisation in between: shows whether k_scan's CTAs are resident before k_step has drained")
    ap.add_argument("--rotate", type=int, default=0, help="step this many other filters between the measured steps (what bench.py "
                    "does: data L2-cold, code L2-warm) instead of flushing the L2 with a fill (code cold as well)")
    a = ap.parse_args()
    import torch

    import slam_robot_simu_b200 as mcl

    wl = dict(bench.WORKLOADS[a.workload])
    if a.n:
        wl["n"] = a.n
    lm = bench.landmarks_for(wl["L"])
    f = mcl.Filter(n_particles=wl["n"], n_trials=wl["T"], landmarks=lm, likelihood="moment", max_landmarks=wl["L"])
    tr = bench.Truth(lm, seed=7, T=wl["T"])
    f.init_particles(tr.pose())
    cyc = torch.from_numpy(tr.cycle()).cuda()
    for k in range(30):
        f.step_dev(cyc[k], force_resample=True)
    f.debug_timeline()  # arm
    n_scan = wl["T"] * ((wl["n"] + 2047) // 2048 + 2)  # the api sizes the scan region for tpt + 2 CTAs
    others = []
    for r in range(a.rotate):
        g = mcl.Filter(n_particles=wl["n"], n_trials=wl["T"], landmarks=lm, likelihood="moment", max_landmarks=wl["L"], seed=100 + r)
        g.init_particles(tr.pose())
        others.append(g)
    for rep in range(3):
        junk = None
        if others:
            for k in range(4):
                for g in others:
                    g.step_dev(cyc[k], force_resample=True)
        else:
            junk = torch.empty(64 << 20, dtype=torch.float32, device="cuda").fill_(1.0)  # flush L2
        torch.cuda.synchronize()
        if a.back_to_back:
            f.step_dev(cyc[29], force_resample=True)
        f.step_dev(cyc[30 + rep], force_resample=True)
        t = f.debug_timeline().astype(np.int64)
        ts, tc = t[:-n_scan], t[-n_scan:]
        t0 = tc[:, 0][tc[:, 0] > 0].min()
        print(f"--- step {rep}: microseconds after the first k_scan CTA started")
        show("k_scan", tc, SCAN, t0)
        print("    tile 0 (merges and publishes): " + "  ".join(f"{nm} {(v - t0) / 1e3:.2f}" for nm, v in zip(SCAN + ["-"], tc[0]) if v > 0))
        prep_rows = np.nonzero(tc[:, 6] > 0)[0]  # (the scalars CTA is the only one that stamps "prep-done")
        prow = tc[prep_rows[0]] if prep_rows.size else tc[-1]
        print("    scalars CTA: " + "  ".join(f"{nm} {(v - t0) / 1e3:.2f}" for nm, v in zip(SCAN + ["-"], prow) if v > 0))
        show("k_step", ts, STEP, t0)
        del junk


if __name__ == "__main__":
    main()
