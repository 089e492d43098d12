#!/usr/bin/env python
"""Record what the drop-in ParticleFilter returns, frame by frame, on a GPU box.

    python tools/record_dropin_frames.py [out.npz]          # default gpurun_out/dropin_frames.npz

The reference's own animate() (particle_filter.py:248-328) cannot meet the GPU class in one place: /root/reference
exists only in the authoring container, which has no GPU.  So the 7-tuples main_pf() returns on the B200 are recorded
here (seed 0, numpy noise: the reference's RNG stream), committed as tests/golden/dropin_frames.npz, and
tests/test_host_logic.py replays them into the UNMODIFIED animate() in the authoring container.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import slam_robot_simu_b200 as mcl

    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "dropin_frames.npz")
    np.random.seed(0)
    pf = mcl.ParticleFilter(100)  # the reference's constructor call (particle_filter.py:339)
    rec = {}
    n_frames = 8
    for i in range(n_frames):
        LM, x_true, x_est, px, Q, idx, val = pf.main_pf()
        rec[f"f{i}_LM"], rec[f"f{i}_x_true"], rec[f"f{i}_x_est"] = np.array(LM), np.array(x_true), np.array(x_est)
        rec[f"f{i}_px"], rec[f"f{i}_Q"] = np.array(px, dtype=np.float64), np.array(Q)
        rec[f"f{i}_idx"], rec[f"f{i}_val"] = int(idx), float(val)
        rec[f"f{i}_types"] = np.array([type(idx).__name__, type(val).__name__, type(px).__name__])
    rec["n_frames"] = n_frames
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **rec)
    print("wrote", out_path, "last frame idx/val", idx, val)


if __name__ == "__main__":
    main()
