"""On-disk trajectory / particle dump of a filter run.

The reference keeps its run only on screen (the trails P1 / P2 of animate(), particle_filter.py:240-301).  `TrajectoryLog`
records, per step, what main_pf() returned -- ground truth, MAP estimate, arg-max index and weight -- plus the engine's
statistics (ESS, log-normaliser, weighted mean and covariance, resampled / underflow flags) and, optionally, a strided
sub-sample of the particle cloud, and writes one .npz: enough to re-draw the animation or to diff two runs offline.
Host-side bookkeeping only; nothing here computes on particles.
"""
from __future__ import annotations

import numpy as np


class TrajectoryLog:
    def __init__(self, particles_every=0, max_particles=4096):
        self.particles_every = int(particles_every)  # 0: never store particles; k: every k-th step
        self.max_particles = int(max_particles)
        self.rows = {k: [] for k in ("x_true", "x_est", "max_idx", "max_weight", "ess", "log_norm", "mean", "cov",
                                     "resampled", "underflow", "step")}
        self.clouds = {}

    def record(self, x_true, x_est, max_idx, max_weight, stats, px=None):
        """stats: an mcl_stats (or the dict of engine.stats_to_dict); px: (3, NP) array-like, read only when stored."""
        g = (lambda k: stats[k]) if isinstance(stats, dict) else (lambda k: getattr(stats, k))
        r = self.rows
        r["x_true"].append(np.asarray(x_true, dtype=np.float64).reshape(3))
        r["x_est"].append(np.asarray(x_est, dtype=np.float64).reshape(3))
        r["max_idx"].append(int(max_idx))
        r["max_weight"].append(float(max_weight))
        r["ess"].append(float(g("ess")))
        r["log_norm"].append(float(g("log_norm")))
        r["mean"].append(np.array(g("mean")[:], dtype=np.float64).reshape(3))
        r["cov"].append(np.array(g("cov")[:], dtype=np.float64).reshape(3, 3))
        r["resampled"].append(int(g("resampled")))
        r["underflow"].append(int(g("underflow")))
        r["step"].append(int(g("step")))
        k = len(r["step"]) - 1
        if px is not None and self.particles_every > 0 and k % self.particles_every == 0:
            a = np.asarray(px, dtype=np.float64)
            stride = max(1, a.shape[1] // self.max_particles)
            self.clouds[k] = a[:, ::stride].astype(np.float32)

    def __len__(self):
        return len(self.rows["step"])

    def save(self, path):
        out = {k: np.array(v) for k, v in self.rows.items()}
        out["cloud_steps"] = np.array(sorted(self.clouds), dtype=np.int64)
        for k, a in self.clouds.items():
            out[f"cloud_{k}"] = a
        np.savez_compressed(path, **out)

    @staticmethod
    def load(path):
        with np.load(path) as z:
            d = {k: z[k] for k in z.files}
        d["clouds"] = {int(k): d.pop(f"cloud_{int(k)}") for k in d["cloud_steps"]}
        return d
