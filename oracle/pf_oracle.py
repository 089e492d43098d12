"""TEST INFRASTRUCTURE ONLY -- CPU (numpy, fp64) restatement of the reference particle-filter step.

This module is the *checker* for the CUDA path.  It may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.
The product package (``slam-robot_simu_b200``) never imports it and has no CPU fallback.

Every function cites the reference lines it restates (paths relative to /root/reference).

Parity status: the reference ships no tests or golden vectors, so this
restatement is pinned against outputs of the *unmodified reference class* run in the authoring
container behind ``oracle/mpl_stub.py``; the generator is ``tests/golden/make_golden.py`` and
the vectors live in ``tests/golden/*.npz`` (checked by ``tests/test_oracle_golden.py``).
The range/bearing likelihoods (``loglik_rb``, ``loglik_rb_fan``) have no counterpart in the reference: parity
unpinned for them; the scan sensor they consume (``scan_sensor``) is pinned (``tests/golden/scan_sensor.npz``).
"""
from __future__ import annotations

import math

import numpy as np

TWO_POW_52 = 4503599627370496.0
BASE_ANG = np.pi / 2.0  # mylib/transform.py:12


# --------------------------------------------------------------------------------------
# mylib/limit.py:11-26
# --------------------------------------------------------------------------------------
def limit_angle(angle_in):
    """Literal scalar restatement of mylib/limit.py:19-26."""
    angle_out = np.absolute(angle_in)
    while angle_out > np.pi:
        angle_out -= np.pi * 2
    if angle_in < 0:
        angle_out *= -1
    return angle_out


def limit_angle_vec(a):
    """Vectorised mylib/limit.py:19-26 (same sequence of fp64 operations per element)."""
    a = np.asarray(a, dtype=np.float64)
    r = np.absolute(a)
    while True:
        m = r > np.pi
        if not m.any():
            break
        r = np.where(m, r - np.pi * 2, r)
    return np.where(a < 0, r * -1, r)


# --------------------------------------------------------------------------------------
# mylib/transform.py:14-36
# --------------------------------------------------------------------------------------
def world2robot(origin, world):
    """Literal restatement of mylib/transform.py:31-36; origin (3,1), world (L,2) -> (L,2)."""
    yaw = BASE_ANG - origin[2, 0]
    diff = world - origin.T[0, 0:2]
    rot = np.array([[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]])
    return (rot @ diff.T).T


def world2robot_batch(px, world):
    """world2robot for every column of px (3,N) -> (N,L,2); same formula, broadcast."""
    yaw = BASE_ANG - px[2]  # (N,)
    c, s = np.cos(yaw), np.sin(yaw)
    dx = world[None, :, 0] - px[0][:, None]  # (N,L)
    dy = world[None, :, 1] - px[1][:, None]
    rx = c[:, None] * dx - s[:, None] * dy
    ry = s[:, None] * dx + c[:, None] * dy
    return np.stack([rx, ry], axis=-1)


# --------------------------------------------------------------------------------------
# matplotlib.mlab.bivariate_normal (matplotlib <= 3.0), call site particle_filter.py:191
# --------------------------------------------------------------------------------------
def bivariate_normal(X, Y, sigmax=1.0, sigmay=1.0, mux=0.0, muy=0.0, sigmaxy=0.0):
    Xmu = X - mux
    Ymu = Y - muy
    rho = sigmaxy / (sigmax * sigmay)
    z = Xmu ** 2 / sigmax ** 2 + Ymu ** 2 / sigmay ** 2 - 2 * rho * Xmu * Ymu / (sigmax * sigmay)
    denom = 2 * np.pi * sigmax * sigmay * np.sqrt(1 - rho ** 2)
    return np.exp(-z / (2 * (1 - rho ** 2))) / denom


# --------------------------------------------------------------------------------------
# constants of ParticleFilter.__init__, particle_filter.py:30-84
# --------------------------------------------------------------------------------------
class PFParams:
    """The literals of particle_filter.py:30-84, overridable for the larger configs."""

    def __init__(self, period_ms=100, n_particles=1000, landmarks=None, ess_div=100.0):
        self.DT_s = period_ms / 1000  # :30
        self.NP = int(n_particles)  # :31
        self.NP_RECIP = 1 / self.NP  # :32
        self.ESS_TH = self.NP / ess_div  # :33
        if landmarks is None:
            landmarks = [[5.0, 5.0], [2.0, -3.0], [-3.0, 4.0], [-5.0, -1.0], [0.0, 0.0]]  # :39-43
        self.LM = np.array(landmarks, dtype=np.float64)
        self.RADIUS_m = 10.0  # :46
        self.YAW_RATE_rps = np.deg2rad(10.0)  # :47
        self.VEL_mps = self.RADIUS_m * self.YAW_RATE_rps  # :48
        self.A = np.eye(3)  # :51-53
        self.B = np.diag([self.VEL_mps, self.VEL_mps, self.YAW_RATE_rps])  # :56-58
        self.Q = np.diag([0.03, 0.03, np.deg2rad(2.0)]) ** 2  # :62-65
        self.R = np.diag([0.3, 0.3]) ** 2  # :68-70
        self.x0 = np.array([[self.RADIUS_m], [0.0], [np.deg2rad(90.0)]])  # :74-79


# --------------------------------------------------------------------------------------
# particle_filter.py:121-142  (__f)
# --------------------------------------------------------------------------------------
def f_motion(x, P: PFParams):
    """x (3,N) -> x_next (3,N); restates particle_filter.py:129-142 operation by operation."""
    yaw = x[2, :]
    a = P.DT_s * np.cos(yaw)
    b = P.DT_s * np.sin(yaw)
    c = np.full_like(a, P.DT_s)
    u = np.array([a, b, c])
    x_next = (P.A @ x) + (P.B @ u)
    x_next[2, :] = limit_angle_vec(x_next[2, :])  # :139-140 (scalar loop in the reference)
    return x_next


def predict(px, v, P: PFParams):
    """particle_filter.py:165-166 with the noise draw ``v`` (3,N) injected."""
    return f_motion(px, P) + v


def observation(x_true, w, P: PFParams):
    """particle_filter.py:151-153 with the noise draw ``w`` (L,2) injected."""
    return world2robot(x_true, P.LM) + w


# --------------------------------------------------------------------------------------
# particle_filter.py:170-198  (__likelihood) and :226-237 (__normalize)
# --------------------------------------------------------------------------------------
def normalize(pw, NP_RECIP):
    """particle_filter.py:234-237."""
    with np.errstate(invalid="ignore", divide="ignore"):
        pw_sum = np.sum(pw)
        nor_pw = pw / pw_sum
    nor_pw[np.isnan(nor_pw)] = NP_RECIP
    return nor_pw


def likelihood_literal(px_est, pw, z_l, P: PFParams):
    """The per-particle Python loop of particle_filter.py:179-198, verbatim in structure."""
    sigma_xx = np.sqrt(P.R[0][0])
    sigma_yy = np.sqrt(P.R[1][1])
    sigma_xy = np.sqrt(P.R[0][1])
    NP = px_est.shape[1]
    bn = np.zeros(NP)
    for i in range(NP):
        px = np.array([px_est[:, i]]).T
        pz_l = world2robot(px, P.LM)
        diff_pz = pz_l - z_l
        dx = diff_pz[:, 0]
        dy = diff_pz[:, 1]
        bnlm = bivariate_normal(dx, dy, sigma_xx, sigma_yy, 0.0, 0.0, sigma_xy)
        bn[i] = bnlm.prod()
    pw_update = pw * bn
    return normalize(pw_update, 1 / NP)


def likelihood_product(px_est, z_l, LM, sigma_x, sigma_y, chunk=1 << 16):
    """bn[i] of particle_filter.py:185-192 for all particles at once (linear domain, fp64)."""
    N = px_est.shape[1]
    out = np.empty(N)
    for s in range(0, N, chunk):
        e = world2robot_batch(px_est[:, s:s + chunk], LM) - z_l[None]
        out[s:s + chunk] = bivariate_normal(e[..., 0], e[..., 1], sigma_x, sigma_y).prod(axis=1)
    return out


def loglik_cartesian(px_est, z_l, LM, sigma_x, sigma_y, chunk=1 << 16):
    """log of bn[i] (particle_filter.py:185-192) evaluated in the log domain, fp64.

    log pdf = -(dx^2/sx^2 + dy^2/sy^2)/2 - log(2 pi sx sy)  (rho = 0 because
    sigma_xy = sqrt(R[0][1]) = 0, particle_filter.py:181).
    """
    N = px_est.shape[1]
    L = LM.shape[0]
    out = np.empty(N)
    for s in range(0, N, chunk):
        e = world2robot_batch(px_est[:, s:s + chunk], LM) - z_l[None]
        q = (e[..., 0] / sigma_x) ** 2 + (e[..., 1] / sigma_y) ** 2
        out[s:s + chunk] = -0.5 * q.sum(axis=1)
    return out - L * math.log(2 * math.pi * sigma_x * sigma_y)


def update_weights_log(logw_prev, loglik):
    """Log-domain form of particle_filter.py:194-196: returns (normalised weights, log-normaliser)."""
    lw = logw_prev + loglik
    m = np.max(lw)
    s = np.sum(np.exp(lw - m))
    logz = m + math.log(s)
    return np.exp(lw - logz), logz


def ess(pw):
    """particle_filter.py:210."""
    pw = np.asarray(pw, dtype=np.float64).reshape(1, -1)
    return float(np.reciprocal(pw @ pw.T)[0, 0])


# --------------------------------------------------------------------------------------
# particle_filter.py:200-224  (__resampling)
# --------------------------------------------------------------------------------------
def comb(N, ofs):
    """resample_id of particle_filter.py:213-215 (``ofs`` already multiplied by 1/N)."""
    base_id = np.arange(0.0, 1.0, 1 / N)
    if base_id.shape[0] != N:  # arange length is ceil((stop-start)/step); see SURVEY Appendix A
        base_id = np.arange(N) * (1 / N)
    return base_id + ofs


def resample_indices_literal(pw, ofs):
    """The sequential walk of particle_filter.py:212-221, returning the ancestor of every slot.

    Raises IndexError exactly where the reference would (comb value above pw_cum[-1]).
    """
    pw = np.asarray(pw, dtype=np.float64).reshape(-1)
    N = pw.shape[0]
    pw_cum = np.cumsum(pw)
    resample_id = comb(N, ofs)
    anc = np.empty(N, dtype=np.int64)
    idx = 0
    for i in range(N):
        while resample_id[i] > pw_cum[idx]:
            idx += 1
        anc[i] = idx
    return anc


def resample_indices(pw, ofs, clamp=True):
    """Vectorised equivalent of the walk: first idx with pw_cum[idx] >= resample_id[i].

    np.searchsorted(side='left') returns exactly that for a non-decreasing cumulative sum.
    ``clamp`` applies the one documented divergence (ancestor N-1 where the reference would
    run off the end of pw_cum).
    """
    pw = np.asarray(pw, dtype=np.float64).reshape(-1)
    N = pw.shape[0]
    pw_cum = np.cumsum(pw)
    anc = np.searchsorted(pw_cum, comb(N, ofs), side="left")
    if clamp:
        anc = np.minimum(anc, N - 1)
    return anc.astype(np.int64)


def quantize_weights(w):
    """Normalised weights -> integers on the 2^-52 grid (round to nearest even, like cvt.rn)."""
    return np.rint(np.asarray(w, dtype=np.float64) * TWO_POW_52).astype(np.uint64)


def dequantize_weights(q):
    """Exact: every q < 2^53 is representable, and so is q * 2^-52."""
    return np.asarray(q, dtype=np.uint64).astype(np.float64) / TWO_POW_52


def resampling(px, pw, ofs_draw, P: PFParams):
    """particle_filter.py:210-224; ``ofs_draw`` is a callable returning np.random.rand()."""
    e = ess(pw)
    if e < P.ESS_TH:
        ofs = ofs_draw() * P.NP_RECIP
        anc = resample_indices(pw, ofs, clamp=False)
        px = px[:, anc]
        pw = np.full((1, P.NP), P.NP_RECIP)
        return px, pw, anc, ofs
    return px, pw, None, None


def map_estimate(px, pw):
    """particle_filter.py:115-117."""
    pw = np.asarray(pw).reshape(-1)
    max_val = np.max(pw)
    max_idx = int(np.argmax(pw))
    x_est = np.array(px[:, max_idx], ndmin=2).T
    return max_val, max_idx, x_est


# --------------------------------------------------------------------------------------
# extension a-14: weighted (circular) mean and 3x3 covariance feeding mylib/error_ellipse.py:39-55
# --------------------------------------------------------------------------------------
def weighted_moments(px, w):
    """mean (3,) with circular yaw mean, covariance (3,3) with the yaw residual wrapped."""
    w = np.asarray(w, dtype=np.float64).reshape(-1)
    w = w / w.sum()
    mx = np.sum(w * px[0])
    my = np.sum(w * px[1])
    mth = math.atan2(np.sum(w * np.sin(px[2])), np.sum(w * np.cos(px[2])))
    d = np.stack([px[0] - mx, px[1] - my, limit_angle_vec(px[2] - mth)])
    cov = (d * w[None]) @ d.T
    return np.array([mx, my, mth]), cov


# --------------------------------------------------------------------------------------
# extension a-15: MotionModel.moveWithNoise, motion_model.py:40-56, batched
# --------------------------------------------------------------------------------------
def motion_b(pose, v, w, dt, noise_a, normals):
    """pose (3,N); normals (3,N) standard normal draws in the order vel, omg, gam (:46-48).

    Keeps the reference quirk that ``sigma ** 2`` is passed as the *standard deviation*.
    """
    a1, a2, a3, a4, a5, a6 = noise_a
    velSq, omgSq = v ** 2, w ** 2
    velSigma = a1 * velSq + a2 * omgSq
    omgSigma = a3 * velSq + a4 * omgSq
    gamSigma = a5 * velSq + a6 * omgSq
    velHat = v + normals[0] * velSigma ** 2
    omgHat = w + normals[1] * omgSigma ** 2
    gamHat = normals[2] * gamSigma ** 2
    a = velHat / omgHat
    b = omgHat * dt
    pt = pose[2]
    pxs = pose[0] - (a * np.sin(pt)) + (a * np.sin(pt + b))
    pys = pose[1] + (a * np.cos(pt)) - (a * np.cos(pt + b))
    pts = limit_angle_vec(pt + (omgHat + gamHat) * dt)
    return np.stack([pxs, pys, pts])


# --------------------------------------------------------------------------------------
# extension a-16: range/bearing likelihood (sensor model graph_based_slam.py:150-167). UNPINNED.
# --------------------------------------------------------------------------------------
def rb_observation(x_true, LM):
    """Noise-free (range, bearing) of every landmark in the robot frame (graph_based_slam.py:150-153)."""
    p = world2robot(x_true, LM)
    return np.stack([np.hypot(p[:, 0], p[:, 1]), np.arctan2(p[:, 1], p[:, 0])], axis=1)


def loglik_rb(px_est, z_rb, LM, sigma_d, sigma_phi, chunk=1 << 15):
    """-1/2 sum_l [(d_obs-d_i)^2/sd^2 + limit_angle(phi_obs-phi_i)^2/sphi^2] - L log(2 pi sd sphi)."""
    N = px_est.shape[1]
    L = LM.shape[0]
    out = np.empty(N)
    for s in range(0, N, chunk):
        p = world2robot_batch(px_est[:, s:s + chunk], LM)
        d = np.hypot(p[..., 0], p[..., 1])
        phi = np.arctan2(p[..., 1], p[..., 0])
        ed = (z_rb[None, :, 0] - d) / sigma_d
        ephi = limit_angle_vec(z_rb[None, :, 1] - phi) / sigma_phi
        out[s:s + chunk] = -0.5 * (ed ** 2 + ephi ** 2).sum(axis=1)
    return out - L * math.log(2 * math.pi * sigma_d * sigma_phi)


# --------------------------------------------------------------------------------------
# the fan-shaped scan sensor (graph_based_slam.py:128-167) and the range/bearing
# likelihood with its visibility gate.  The SENSOR is pinned (tests/golden/scan_sensor.npz is
# written by the unmodified ScanSensor.scan); the LIKELIHOOD has no reference counterpart --
# the reference uses these observations in a graph optimiser, not a particle filter -- and is
# defined here: parity unpinned.
# --------------------------------------------------------------------------------------
def fan_visible(robot_lm, scan_range, scan_angle):
    """graph_based_slam.py:150-160: inside the scan range and inside the fan of +-scan_angle about the
    heading (the robot looks along +y of its own frame, mylib/transform.py:12)."""
    dist = np.linalg.norm(robot_lm, axis=-1)
    scan_rad = BASE_ANG - scan_angle
    return (dist <= scan_range) & (robot_lm[..., 1] >= np.absolute(robot_lm[..., 0]) * np.tan(scan_rad))


def scan_sensor(pose, LM, scan_range, scan_angle, r_dist, dir_sigma, orient_sigma, rng=np.random):
    """ScanSensor.scan (graph_based_slam.py:128-167).  Returns (ids, noisy (n,3), clean (n,3)) with columns
    (distance, direction, orientation); draws np.random.normal in the reference's order (:163-165)."""
    p = world2robot(pose, LM)
    dist = np.linalg.norm(p, axis=1)
    dirs = np.arctan2(p[:, 1], p[:, 0])
    orient = np.ones(p.shape[0]) * (BASE_ANG - pose[2, 0])
    flg = fan_visible(p, scan_range, scan_angle)
    ids, noisy, clean = [], [], []
    for i in np.nonzero(flg)[0]:
        d = rng.normal(dist[i], dist[i] * r_dist)
        a = limit_angle(rng.normal(dirs[i], dir_sigma))
        o = limit_angle(rng.normal(orient[i], orient_sigma))
        ids.append(int(i))
        noisy.append((d, a, o))
        clean.append((dist[i], dirs[i], orient[i]))
    return np.array(ids, dtype=np.int64), np.array(noisy).reshape(-1, 3), np.array(clean).reshape(-1, 3)


def rb_fan_table(ids, noisy, L):
    """(L, 2) observation table of the engine: (distance, direction) per landmark, NaN where not detected."""
    z = np.full((L, 2), np.nan)
    if len(ids):
        z[ids, 0] = noisy[:, 0]
        z[ids, 1] = noisy[:, 1]
    return z


def loglik_rb_fan(px_est, z_rb, LM, scan_range, scan_angle, r_dist, dir_sigma, miss_chi2, chunk=1 << 14):
    """Sum over the DETECTED landmarks l (finite z_rb[l, 0]) of
         -1/2 [ (d_obs - d_i)^2 / (d_obs r_dist)^2 + limit_angle(phi_obs - phi_i)^2 / dir_sigma^2 ]
    when the landmark lies inside particle i's own fan (fan_visible from the particle's pose), and of
    -1/2 miss_chi2 when it does not: the particle predicts it cannot see what was seen.  The distance noise
    follows the sensor, sigma_d = distance * R_Dist (:163).  No normalisation constant (it is the same for
    every particle)."""
    N = px_est.shape[1]
    obs = np.nonzero(np.isfinite(z_rb[:, 0]))[0]
    out = np.zeros(N)
    if obs.size == 0:
        return out
    lm, zo = LM[obs], z_rb[obs]
    sd = zo[:, 0] * r_dist
    for s in range(0, N, chunk):
        p = world2robot_batch(px_est[:, s:s + chunk], lm)
        d = np.sqrt(p[..., 0] ** 2 + p[..., 1] ** 2)
        phi = np.arctan2(p[..., 1], p[..., 0])
        ed = (zo[None, :, 0] - d) / sd[None, :]
        ephi = limit_angle_vec(zo[None, :, 1] - phi) / dir_sigma
        vis = fan_visible(p, scan_range, scan_angle)
        out[s:s + chunk] = -0.5 * np.where(vis, ed ** 2 + ephi ** 2, miss_chi2).sum(axis=1)
    return out


# --------------------------------------------------------------------------------------
# particle_filter.py:86-119  (main_pf): the whole step, drawing from np.random in the
# reference's order (rand if resampling -> multivariate_normal(Q,NP) -> multivariate_normal(R,L))
# --------------------------------------------------------------------------------------
class OracleFilter:
    """Vectorised restatement of ParticleFilter (particle_filter.py:18-237), linear-domain weights."""

    def __init__(self, period_ms=100, n_particles=1000, landmarks=None, rng=None):
        self.P = PFParams(period_ms, n_particles, landmarks)
        self.rng = np.random if rng is None else rng
        P = self.P
        self.x_true = P.x0.copy()
        self.px = np.tile(P.x0, (1, P.NP))  # :81
        self.pw = np.full((1, P.NP), P.NP_RECIP)  # :83-84
        self.last = {}

    def main_pf(self, literal_likelihood=False):
        P = self.P
        self.x_true = f_motion(self.x_true, P)  # :100
        ess_before = ess(self.pw)
        self.px, self.pw, anc, ofs = resampling(self.px, self.pw, self.rng.rand, P)  # :104
        v = self.rng.multivariate_normal([0.0, 0.0, 0.0], P.Q, P.NP).T  # :165
        self.px = predict(self.px, v, P)  # :166
        w = self.rng.multivariate_normal([0.0, 0.0], P.R, P.LM.shape[0])  # :152
        z_l = observation(self.x_true, w, P)  # :151-153
        if literal_likelihood:
            self.pw = likelihood_literal(self.px, self.pw, z_l, P)
        else:
            sx, sy = np.sqrt(P.R[0][0]), np.sqrt(P.R[1][1])
            bn = likelihood_product(self.px, z_l, P.LM, sx, sy)
            self.pw = normalize(self.pw * bn, P.NP_RECIP)  # :194-196
        max_val, max_idx, x_est = map_estimate(self.px, self.pw)  # :115-117
        self.last = dict(ess_before=ess_before, anc=anc, ofs=ofs, v=v, w=w, z_l=z_l)
        return P.LM, self.x_true, x_est, self.px, P.Q, max_idx, max_val
