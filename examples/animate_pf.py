#!/usr/bin/env python
"""The reference's animation loop (particle_filter.py:248-346) driven by the B200 filter.

    python examples/animate_pf.py [--particles 1000000] [--landmarks 16] [--frames 360] [--save out.mp4]

The reference script needs ONE changed line to run on the GPU engine (import the
drop-in class); its `animate()` then works as it is, because `main_pf()` returns the same 7-tuple and `px`
answers `px[0]`, `px[1]`, `px[2]` (LazyParticles copies the particles off the GPU only when they are read).
This file is that loop restated for large filters: the scatter is drawn from a sub-sample (a million markers help
nobody), the error ellipse of the weighted covariance (`pose_mean_cov`, feeding mylib/error_ellipse.py:39-55 the
way extended_kalman_filter.py:259-260 does) is added, and matplotlib is imported here only -- the engine never
needs it.  Drawing is off the timed path by construction: `frame()` = one `main_pf()` + host plotting.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ellipse_points(cov2, chi2=9.21034, n=64):
    """Outline of the confidence ellipse of a 2x2 covariance (99 %: chi2 with 2 dof = 9.21; what
    ErrorEllipse.calc_error_ellipse + its caller draw, mylib/error_ellipse.py:39-55)."""
    w, v = np.linalg.eigh(cov2)
    w = np.maximum(w, 0.0)
    t = np.linspace(0.0, 2.0 * np.pi, n)
    return v @ (np.sqrt(chi2 * w)[:, None] * np.stack([np.cos(t), np.sin(t)]))


class Animator:
    def __init__(self, pf, period_ms, plt, max_scatter=20000, seed=0):
        self.pf, self.period_ms, self.plt = pf, period_ms, plt
        self.max_scatter = int(max_scatter)
        self.rng = np.random.RandomState(seed)
        self.truth, self.est = [], []
        self.time_s = 0.0

    def frame(self, i=0):
        plt = self.plt
        self.time_s += self.period_ms / 1000
        lm, x_true, x_est, px, Q, w_idx, w_val = self.pf.main_pf()  # particle_filter.py:256
        n = px.shape[1]
        if n > self.max_scatter:  # lazy D2H of all of px, then a sub-sample for drawing
            pick = self.rng.choice(n, self.max_scatter, replace=False)
            sx, sy, sth = px[0][pick], px[1][pick], px[2][pick]
        else:
            sx, sy, sth = px[0], px[1], px[2]
        mean, cov = self.pf.pose_mean_cov()
        plt.cla()
        ax1 = plt.subplot2grid((1, 2), (0, 0))
        ax2 = plt.subplot2grid((1, 2), (0, 1))
        for ax in (ax1, ax2):
            ax.scatter(lm[:, 0], lm[:, 1], s=100, c="yellow", marker="*", alpha=0.5, edgecolors="orange")  # :264-267
            ax.scatter(sx, sy, c="cyan", marker="o", alpha=0.5)  # :278-279
        ax2.quiver(sx, sy, np.cos(sth), np.sin(sth), color="cyan", units="inches", scale=6.0, width=0.01)  # :280-281
        self.truth.append(x_true[0:2, :])
        self.est.append(x_est[0:2, :])
        for hist, col, cur in ((self.truth, "red", x_true), (self.est, "blue", x_est)):  # :284-301
            a, b = np.array(np.concatenate(hist, axis=1))
            for ax in (ax1, ax2):
                ax.plot(a, b, c=col, linewidth=1.0, linestyle="-")
                ax.scatter(cur[0], cur[1], c=col, marker="o", alpha=0.5)
        e = ellipse_points(cov[0:2, 0:2])
        ax2.plot(mean[0] + e[0], mean[1] + e[1], c="magenta", linewidth=1.0)
        ax2.annotate("Maximuim Likelihood Estimate:\n[Index]:{0}\n[Weight]:{1:.3g}".format(w_idx, w_val),  # :304
                     xy=(x_est[0, 0], x_est[1, 0]), xycoords="data", xytext=(0.55, 0.9), textcoords="axes fraction")
        half = 3.0 * np.sqrt(9.21034 * max(Q[0, 0], Q[1, 1]))  # ee.calc_chi(99, Q[0:2, 0:2]) * 3 (:322)
        ax2.set_xlim(x_true[0][0] - half, x_true[0][0] + half)
        ax2.set_ylim(x_true[1][0] - half, x_true[1][0] + half)
        ax1.set_title("Localization by PF (B200)")
        ax2.set_title("Zoom")
        return dict(n_drawn=len(sx), time_s=self.time_s, w_idx=w_idx, w_val=w_val, mean=mean, cov=cov)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=1000)
    ap.add_argument("--landmarks", type=int, default=0, help="0 = the reference's five landmarks")
    ap.add_argument("--frames", type=int, default=360)
    ap.add_argument("--save", default="")
    a = ap.parse_args()
    import matplotlib.pyplot as plt
    from matplotlib import animation

    import slam_robot_simu_b200 as mcl

    lm = None if a.landmarks == 0 else np.random.default_rng(1234).uniform(-20, 20, (a.landmarks, 2))
    period_ms = 100
    pf = mcl.ParticleFilter(period_ms, n_particles=a.particles, landmarks=lm, noise="philox" if a.particles > 100000 else "numpy",
                            likelihood="moment", lazy_px=True)
    fig = plt.figure(figsize=(18, 9))
    anim = Animator(pf, period_ms, plt)
    ani = animation.FuncAnimation(fig, anim.frame, frames=a.frames, blit=False, interval=period_ms, repeat=False)
    if a.save:
        ani.save(a.save, bitrate=6000)
    else:
        plt.show()


if __name__ == "__main__":
    main()
