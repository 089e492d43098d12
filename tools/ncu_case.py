#!/usr/bin/env python
"""A short, fixed sequence of filter steps for `ncu` (never a timing source).

    python tools/ncu_case.py [--workload c1] [--steps 6] [--likelihood moment]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (workload table, landmark map, truth generator)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c1")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--likelihood", default="moment")
    ap.add_argument("--n", type=int, default=0)
    a = ap.parse_args()
    import torch

    import slam_robot_simu_b200 as mcl

    wl = dict(bench.WORKLOADS[a.workload])
    if a.n:
        wl["n"] = a.n
    lm = bench.landmarks_for(wl["L"])
    f = mcl.Filter(n_particles=wl["n"], n_trials=wl["T"], landmarks=lm, likelihood=a.likelihood, max_landmarks=wl["L"],
                   sigma_obs=[0.3, float(np.deg2rad(2.0))] if a.likelihood in ("rb", "rb32") else None)
    tr = bench.Truth(lm, seed=7, T=wl["T"], fan=(a.likelihood == "rb_fan"), rb=(a.likelihood in ("rb", "rb32")))
    f.init_particles(tr.pose())
    for _ in range(a.steps):
        tr.advance()
        f.step(tr.observe(), force_resample=True)
    torch.cuda.synchronize()
    st = f.stats()[0]
    print(f"ok ess={st.ess:.1f} max_idx={st.max_idx} launches={f.launch_count()}")


if __name__ == "__main__":
    main()
