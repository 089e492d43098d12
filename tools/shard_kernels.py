#!/usr/bin/env python
"""Device time of the two passes of the sharded resampling on ONE GPU (one shard of a two-shard filter):
mcl_shard_total (+ segment prefix) and mcl_shard_expand, by CUDA events.

    python tools/shard_kernels.py [particles_per_shard]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import slam_robot_simu_b200 as mcl  # noqa: E402


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
    lm = bench.landmarks_for(64)
    tr = bench.Truth(lm, seed=7)
    # a running filter's weights: a few unsharded steps on m particles, then the state moves into a shard handle
    f0 = mcl.Filter(n_particles=m, landmarks=lm, likelihood="moment", max_landmarks=64)
    f0.init_particles(tr.pose())
    for _ in range(6):
        tr.advance()
        f0.step(tr.observe(), force_resample=True)
    x, y, th, lw = f0.get_particles(device=True)
    st = f0.stats()[0]
    del f0
    f = mcl.Filter(n_particles=m, landmarks=lm, likelihood="moment", max_landmarks=64, n_global=2 * m, slot_base=0)
    f.set_particles(x[0], y[0], th[0], lw[0])
    g = mcl.mcl_stats()
    g.max_logw, g.sum_w, g.sum_w2, g.ess, g.log_norm = st.max_logw, 2 * st.sum_w, 2 * st.sum_w2, 2 * st.ess, st.log_norm + np.log(2.0)
    for k in range(3):
        g.est[k] = st.est[k]
    f.shard_set_global(g)
    tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    cap = m // 8
    ex = [torch.empty(cap, dtype=torch.float32, device="cuda") for _ in range(3)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_tot, t_exp, t_exp_lb = [], [], []
    for rep in range(6):
        ev[0].record()
        f.shard_total(tot)
        ev[1].record()
        f.shard_expand(0, 0.3 / (2 * m), *ex, want_range=False)
        ev[2].record()
        torch.cuda.synchronize()
        t_tot.append(ev[0].elapsed_time(ev[1]))
        t_exp.append(ev[1].elapsed_time(ev[2]))
        ev[0].record()
        f.shard_expand(0, 0.3 / (2 * m), *ex, want_range=False)  # no fresh totals: the look-back path
        ev[1].record()
        torch.cuda.synchronize()
        t_exp_lb.append(ev[0].elapsed_time(ev[1]))
    print(f"m={m}: shard_total {min(t_tot):.3f} ms, expand with tile prefixes {min(t_exp):.3f} ms, expand with look-back {min(t_exp_lb):.3f} ms")


if __name__ == "__main__":
    main()
