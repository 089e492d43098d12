"""Drop-in `ParticleFilter` with the reference's constructor and `main_pf()` contract.

Mirrors takuyani/SLAM-Robot_Simu particle_filter.py:18-237: same constants (:30-84), same step order
(truth -> resample -> predict -> observe -> weight -> MAP estimate, :86-119), same 7-tuple return
`(LM, x_true, x_est, px, Q, max_idx, max_val)`, and -- with ``noise="numpy"`` -- the same
consumption of the legacy ``np.random`` stream (rand() iff resampling -> multivariate_normal(Q, NP)
-> multivariate_normal(R, L)), so a seeded run sees the reference's exact draws.

What stays on the host is what the reference does once per step on a single pose: the ground-truth
motion and the L-landmark observation.  Everything over the particles (resampling scan + gather,
motion, likelihood, normalisation, ESS, arg-max, mean/covariance) runs in libmcl_b200.so on the GPU.
There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from .engine import Filter

_BASE_ANG = np.pi / 2.0  # mylib/transform.py:12


def limit_angle(angle_in):
    """mylib/limit.py:11-26 for one host-side scalar (the ground-truth yaw)."""
    angle_out = np.absolute(angle_in)
    while angle_out > np.pi:
        angle_out -= np.pi * 2
    if angle_in < 0:
        angle_out *= -1
    return angle_out


def world2robot(origin, world):
    """mylib/transform.py:14-36 for the single true pose; origin (3,1), world (L,2) -> (L,2)."""
    yaw = _BASE_ANG - origin[2, 0]
    diff = world - origin.T[0, 0:2]
    rot = np.array([[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]])
    return (rot @ diff.T).T


class LazyParticles:
    """`px` for large filters: a (3, NP) float64 array fetched from the GPU only when it is read."""

    def __init__(self, flt):
        self._flt = flt
        self._cache = None
        self.shape = (3, flt.n)
        self.dtype = np.dtype(np.float64)
        # The reference returns the array of THIS step; a lazy view can only serve it while the filter has not moved on.
        self._stamp = getattr(flt, "_version", None)

    def _fetch(self):
        if self._cache is None:
            if self._stamp is not None and getattr(self._flt, "_version", None) != self._stamp:
                raise RuntimeError("px of an earlier main_pf() read after the filter advanced: read it (np.asarray(px)) "
                                   "before the next step, or construct ParticleFilter(..., lazy_px=False)")
            x, y, th, _ = self._flt.get_particles(want_logw=False)
            self._cache = np.stack([x[0], y[0], th[0]]).astype(np.float64)
        return self._cache

    def __array__(self, dtype=None, copy=None):
        a = self._fetch()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        return self._fetch()[idx]

    def __len__(self):
        return 3


class ParticleFilter:
    """ParticleFilter(period_ms) -- particle_filter.py:21.  Keyword extras select the larger configs."""

    def __init__(self, period_ms, *, n_particles=1000, landmarks=None, noise="numpy", likelihood="pair",
                 lazy_px=False, force_resample=False, device=0, seed=42, ess_div=100.0, trajectory=None):
        if noise not in ("numpy", "philox"):
            raise ValueError("noise must be 'numpy' (reference RNG stream) or 'philox' (on-device)")
        self.__DT_s = period_ms / 1000  # :30
        self.__NP = int(n_particles)  # :31
        self.__NP_RECIP = 1 / self.__NP  # :32
        self.__ESS_TH = self.__NP / ess_div  # :33
        if landmarks is None:
            landmarks = [[5.0, 5.0], [2.0, -3.0], [-3.0, 4.0], [-5.0, -1.0], [0.0, 0.0]]  # :39-43
        self.__LM = np.array(landmarks, dtype=np.float64)
        self.__RADIUS_m = 10.0  # :46
        self.__YAW_RATE_rps = np.deg2rad(10.0)  # :47
        self.__VEL_mps = self.__RADIUS_m * self.__YAW_RATE_rps  # :48
        self.__Q = np.diag([0.03, 0.03, np.deg2rad(2.0)]) ** 2  # :62-65
        self.__R = np.diag([0.3, 0.3]) ** 2  # :68-70
        self.__x_true = np.array([[self.__RADIUS_m], [0.0], [np.deg2rad(90.0)]])  # :74-79
        # np.random.multivariate_normal(0, C, n) is standard_normal(n * d).reshape(n, d) @ (sqrt(s)[:, None] * v)
        # with (u, s, v) = svd(C) recomputed on every call (numpy mtrand, legacy generator).  The factor is
        # constant, so it is computed once; the draws below consume the np.random stream identically.
        self.__Q_factor = self.__mvn_factor(self.__Q)
        self.__R_factor = self.__mvn_factor(self.__R)
        self.__noise = noise
        self.__lazy = bool(lazy_px)
        self.__force = bool(force_resample)
        self.__sensor = None
        sigma_obs = [np.sqrt(self.__R[0][0]), np.sqrt(self.__R[1][1])]
        if likelihood == "rb_fan":
            # the robot carries the fan-shaped scan sensor of graph_based_slam.py:73-167 with the
            # noise its Robot class sets (:604); the filter scores (distance, direction) of the detected landmarks
            from .scan_sensor import ScanSensor

            self.__sensor = ScanSensor(15.0, np.deg2rad(80.0), self.__LM)  # :901-902
            self.__sensor.setNoiseParam(5, 2, 2)
            sigma_obs = [self.__sensor.r_dist, self.__sensor.dir_sigma]
        self.__flt = Filter(n_particles=self.__NP, n_trials=1, period_ms=period_ms, landmarks=self.__LM,
                            likelihood=likelihood, motion="a", device=device, seed=seed,
                            ess_threshold=self.__ESS_TH,
                            q_std=np.sqrt(np.diag(self.__Q)), sigma_obs=sigma_obs,
                            vel=self.__VEL_mps, yaw_rate=self.__YAW_RATE_rps, dt=self.__DT_s)
        self.__flt.init_particles(self.__x_true[:, 0])  # :81-84
        self.__ess = float(self.__NP)  # 1 / sum((1/NP)^2)
        self.__ahead = None
        self.last_stats = None
        self.trajectory = trajectory  # a TrajectoryLog (or None): every main_pf() is recorded into it

    @staticmethod
    def __mvn_factor(cov):
        _, s, v = np.linalg.svd(cov)
        return np.sqrt(s)[:, None] * v

    @staticmethod
    def __mvn(factor, n):
        """np.random.multivariate_normal(zeros, cov, n), same stream, without the per-call SVD."""
        d = factor.shape[0]
        return np.dot(np.random.standard_normal(n * d).reshape(n, d), factor)

    # -- host-side pieces the reference also runs once per step on one pose ---------------------
    def __f(self, x):
        """particle_filter.py:121-142 for the (3,1) true pose."""
        yaw = x[2, 0]
        x_next = np.array([[x[0, 0] + self.__VEL_mps * (self.__DT_s * np.cos(yaw))],
                           [x[1, 0] + self.__VEL_mps * (self.__DT_s * np.sin(yaw))],
                           [x[2, 0] + self.__YAW_RATE_rps * self.__DT_s]])
        x_next[2, 0] = limit_angle(x_next[2, 0])
        return x_next

    def __observation(self, x_true):
        """particle_filter.py:144-154 (or ScanSensor.scan, graph_based_slam.py:128-167, in the fan-sensor mode)."""
        if self.__sensor is not None:
            return self.__sensor.scan(x_true)[0]
        z_l = world2robot(x_true, self.__LM)
        w = self.__mvn(self.__R_factor, self.__LM.shape[0])  # multivariate_normal([0, 0], R, L)
        return z_l + w

    @property
    def engine(self):
        return self.__flt

    def main_pf(self):
        """particle_filter.py:86-119."""
        flt = self.__flt
        if self.__ahead is not None:  # prepared while the GPU ran the previous step (see below)
            self.__x_true, z_ahead = self.__ahead
            self.__ahead = None
        else:
            self.__x_true = self.__f(self.__x_true)  # :100
            z_ahead = None

        # :104/:210-214 -- the gate is taken on the previous step's weights; the uniform offset is
        # drawn first (and only when resampling) so the np.random stream matches the reference
        do_resample = self.__force or (self.__ess < self.__ESS_TH)
        ofs = None
        noise = None
        if self.__noise == "numpy":
            if do_resample:
                ofs = np.random.rand() * self.__NP_RECIP  # :214
            v = self.__mvn(self.__Q_factor, self.__NP).T  # multivariate_normal([0, 0, 0], Q, NP).T (:165)
            noise = np.ascontiguousarray(v, dtype=np.float32)[:, None, :]
        z_l = z_ahead if z_ahead is not None else self.__observation(self.__x_true)  # :110

        # one crossing of the C ABI: stage z_l, enqueue k_scan + k_step + the statistics merge, start the read-back
        flt.step_async(np.ascontiguousarray(z_l, dtype=np.float64), ofs=ofs, noise=noise, force_resample=do_resample,
                       no_resample=not do_resample)
        if self.__noise == "philox":
            # The kernels are in flight: simulate the NEXT frame's ground truth and observation now (:100, :110 of
            # the next call).  Neither depends on the filter, and with on-device noise the observation draw is the
            # only consumer of np.random, so the stream is consumed in the same order, one frame early.
            x_next = self.__f(self.__x_true)
            self.__ahead = (x_next, self.__observation(x_next))
        st = flt.stats_end()[0]
        self.last_stats = st
        self.__ess = st.ess

        x_est = np.array([[st.est[0]], [st.est[1]], [st.est[2]]])  # :117
        if self.__lazy:
            px = LazyParticles(flt)
        else:
            x, y, th, _ = flt.get_particles(want_logw=False)
            px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
        if self.trajectory is not None:
            self.trajectory.record(self.__x_true, x_est, st.max_idx, st.max_weight, st, px)
        return self.__LM, self.__x_true, x_est, px, self.__Q, int(st.max_idx), float(st.max_weight)

    # -- extension a-14: weighted mean and covariance that feed mylib/error_ellipse.py:39-55 ------
    def pose_mean_cov(self, exact=False):
        """(mean (3,), cov (3,3)); ErrorEllipse.calc_error_ellipse takes cov[0:2, 0:2]."""
        if exact:
            self.__flt.pose_moments()
            st = self.__flt.stats()[0]
        else:
            st = self.last_stats if self.last_stats is not None else self.__flt.stats()[0]
        return np.array(st.mean[:]), np.array(st.cov[:]).reshape(3, 3)
