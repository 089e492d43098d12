#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference in the authoring container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Needs /root/reference (absent on the GPU box -- the committed .npz files are what travels).
The reference is imported behind ``oracle/mpl_stub.py`` (matplotlib is not installed and
``mlab.bivariate_normal`` no longer exists upstream).  ``np.random`` is seeded with the legacy
MT19937 seeder; its streams are frozen by NEP 19, so a test can regenerate the same draws.

Files written
  pf_seed{S}.npz     per-step scalars for every step, and full particle/weight arrays at the
                     steps listed in FULL_STEPS (kept small on purpose)
  helpers.npz        limit_angle / world2robot / MotionModel known-answer vectors
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MCL_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import mpl_stub  # noqa: E402

mpl_stub.install()

import particle_filter as ref_pf  # noqa: E402  (the reference, unmodified)
import motion_model as ref_mm  # noqa: E402
from mylib import limit as ref_limit  # noqa: E402
from mylib import transform as ref_tf  # noqa: E402
from mylib import error_ellipse as ref_ee  # noqa: E402

FULL_STEPS = {0: [0, 1, 5, 6, 7, 10, 100, 359], 1: [0, 5, 6, 39], 2: [0, 4, 39]}
N_STEPS = {0: 360, 1: 40, 2: 40}


class Recorder:
    """Wraps np.random.rand / multivariate_normal to capture the reference's draws per step."""

    def __init__(self):
        self.rand = []
        self.mvn = []
        self._rand = np.random.rand
        self._mvn = np.random.multivariate_normal

    def __enter__(self):
        def rand(*a):
            r = self._rand(*a)
            self.rand.append(r)
            return r

        def mvn(mean, cov, size=None):
            r = self._mvn(mean, cov, size)
            self.mvn.append(np.array(r))
            return r

        np.random.rand = rand
        np.random.multivariate_normal = mvn
        return self

    def __exit__(self, *exc):
        np.random.rand = self._rand
        np.random.multivariate_normal = self._mvn


def run_seed(seed, n_steps, full_steps):
    np.random.seed(seed)
    pf = ref_pf.ParticleFilter(100)
    out = {k: [] for k in ("x_true", "x_est", "max_idx", "max_val", "ess_after", "resampled", "ofs", "z_l",
                           "px_mean", "pw_sum")}
    full = {}
    for k in range(n_steps):
        px_before = np.array(pf._ParticleFilter__px)
        pw_before = np.array(pf._ParticleFilter__pw)
        with Recorder() as rec:
            LM, x_true, x_est, px, Q, max_idx, max_val = pf.main_pf()
        pw = pf._ParticleFilter__pw
        resampled = len(rec.rand) == 1
        v = rec.mvn[0].T  # (3,NP), particle_filter.py:165
        w = rec.mvn[1]  # (L,2),  particle_filter.py:152
        z_l = ref_tf.world2robot(x_true, LM) + w
        out["x_true"].append(x_true[:, 0].copy())
        out["x_est"].append(x_est[:, 0].copy())
        out["max_idx"].append(int(max_idx))
        out["max_val"].append(float(max_val))
        out["ess_after"].append(float(np.reciprocal(pw @ pw.T)[0, 0]))
        out["resampled"].append(resampled)
        out["ofs"].append(rec.rand[0] * (1 / px.shape[1]) if resampled else np.nan)
        out["z_l"].append(z_l)
        out["px_mean"].append(px.mean(axis=1))
        out["pw_sum"].append(float(pw.sum()))
        if k in full_steps:
            full[f"s{k}_px_before"] = px_before
            full[f"s{k}_pw_before"] = pw_before
            full[f"s{k}_v"] = v
            full[f"s{k}_w"] = w
            full[f"s{k}_px"] = np.array(px)
            full[f"s{k}_pw"] = np.array(pw)
    data = {k: np.array(v) for k, v in out.items()}
    data.update(full)
    data["LM"] = LM
    data["Q"] = Q
    data["seed"] = seed
    data["full_steps"] = np.array(sorted(full_steps))
    return data


def helpers():
    rng = np.random.RandomState(7)
    ang = np.concatenate([
        np.array([3.5, -3.5, np.pi, -np.pi, 7.0, 10.0, -0.0, 0.0, np.nextafter(np.pi, 4.0),
                  -np.nextafter(np.pi, 4.0), 2 * np.pi, -2 * np.pi, 100.0, -100.0]),
        rng.uniform(-12, 12, 200)])
    lim = np.array([float(ref_limit.limit_angle(a)) for a in ang])
    poses = rng.uniform(-15, 15, (64, 3, 1))
    poses[:, 2] = rng.uniform(-4, 4, (64, 1))
    lm = rng.uniform(-20, 20, (7, 2))
    w2r = np.array([ref_tf.world2robot(p, lm) for p in poses])
    r2w = np.array([ref_tf.robot2world(p, r) for p, r in zip(poses, w2r)])
    # MotionModel known answers (motion_model.py:31-86), Appendix A of SURVEY.md
    mm = ref_mm.MotionModel(2.0, .1, .1, .1, .1, .1, .1)
    pose = np.array([[10.0], [0.0], [np.pi / 2]])
    v, w = 10 * np.deg2rad(10.0), np.deg2rad(10.0)
    np.random.seed(0)
    mm_noise = np.array([mm.moveWithNoise(pose, v, w)[:, 0] for _ in range(16)])
    np.random.seed(0)
    mm_normals = np.random.standard_normal((16, 3))  # normal(0,s) == s*standard_normal(): same stream
    mm_clean = mm.moveWithoutNoise(pose, v, w)[:, 0]
    # ErrorEllipse consumer (mylib/error_ellipse.py:39-68)
    ee = ref_ee.ErrorEllipse(99.0)
    covs = []
    ells = []
    for _ in range(8):
        a = rng.normal(size=(2, 2))
        c = a @ a.T + 0.1 * np.eye(2)
        covs.append(c)
        ells.append(ee.calc_error_ellipse(c))
    return dict(lim_in=ang, lim_out=lim, w2r_pose=poses, w2r_lm=lm, w2r_out=w2r, r2w_out=r2w,
                mm_pose=pose, mm_v=v, mm_w=w, mm_dt=2.0, mm_a=np.full(6, .1), mm_noise=mm_noise,
                mm_normals=mm_normals, mm_clean=mm_clean, ee_cov=np.array(covs), ee_out=np.array(ells),
                ee_chi_Q=ee.calc_chi(99.0, (np.diag([0.03, 0.03, np.deg2rad(2.0)]) ** 2)[0:2, 0:2]))


def main():
    for seed, n in N_STEPS.items():
        d = run_seed(seed, n, FULL_STEPS[seed])
        path = os.path.join(HERE, f"pf_seed{seed}.npz")
        np.savez_compressed(path, **d)
        print(path, os.path.getsize(path) // 1024, "KiB", "resamples:", int(d["resampled"].sum()),
              "first:", np.nonzero(d["resampled"])[0][:8])
    path = os.path.join(HERE, "helpers.npz")
    np.savez_compressed(path, **helpers())
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
