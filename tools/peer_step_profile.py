#!/usr/bin/env python
"""Per-kernel durations of the device-driven sharded step on ONE GPU (a filter of one shard: every exchange is
satisfied by the shard's own pushes, so the kernels are exactly those of a multi-GPU step minus the waiting).

    python tools/peer_step_profile.py [--n 125000000] [--L 64] [--steps 20]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=125_000_000)
    ap.add_argument("--L", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--likelihood", default="moment")
    a = ap.parse_args()
    import torch

    import slam_robot_simu_b200 as mcl

    lm = bench.landmarks_for(a.L)
    tr = bench.Truth(lm, seed=77)
    cyc = tr.cycle()
    sf = mcl.LocalShardedFilter(a.n, lm, 1, likelihood=a.likelihood, seed=4711, max_landmarks=a.L)
    sf.init_particles(tr.pose()[0])
    f = sf.shards[0]
    for k in range(25):
        f.step_peer(cyc[k], force_resample=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.steps):
        f.step_peer(cyc[(25 + k) % 360], force_resample=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"n={a.n} L={a.L}: {e0.elapsed_time(e1) / a.steps:.4f} ms/step  ({a.n * a.steps / e0.elapsed_time(e1) * 1e3:.4g} particle-steps/s)")
    f.profile(True)
    acc = np.zeros(5)
    for k in range(a.steps):
        f.step_peer(cyc[(25 + a.steps + k) % 360], force_resample=True)
        acc += np.array(f.profile_read_peer())
    acc /= a.steps
    names = ["totals pass + exchange", "scalars + expansion + exchange", "(unused)", "k_step", "statistics exchange"]
    for nm, v in zip(names, acc):
        print(f"  {nm:26s} {v * 1e3:9.1f} us")
    f.profile(False)
    # the same through the four stages, each bracketed by a device synchronisation (host clock)
    import time

    L = mcl._lib
    acc = np.zeros(4)
    for k in range(a.steps):
        for i, stg in enumerate((L.MCL_STAGE_A, L.MCL_STAGE_B, L.MCL_STAGE_C, L.MCL_STAGE_D)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f.step_peer(cyc[(25 + 2 * a.steps + k) % 360], force_resample=True, stages=stg)
            torch.cuda.synchronize()
            acc[i] += time.perf_counter() - t0
    acc /= a.steps
    for nm, v in zip(["stage A (totals, push)", "stage B (plan, expansion, push)", "stage C (scalars, k_step, push)", "stage D (merge)"], acc):
        print(f"  {nm:34s} {v * 1e6:9.1f} us (host clock, synchronised)")
    st = f.stats()[0]
    print(f"  ess/n = {st.ess / a.n:.4f}  plan.error = {f.shard_plan().error}")


if __name__ == "__main__":
    main()
