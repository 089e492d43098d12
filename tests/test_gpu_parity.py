"""GPU parity tests: the CUDA path (through the ctypes C ABI) against the CPU oracle.

Everything here needs a B200 (``-m gpu``).  The oracle (oracle/pf_oracle.py) is only the checker.
Tolerances: resampling indices bit-exact; poses / weights / covariances 1e-5 relative (fp32 state
against the fp64 oracle, BASELINE.json north_star).
"""
import os

import numpy as np
import pytest

from oracle import pf_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def mcl():
    import slam_robot_simu_b200 as m

    return m


def _golden(golden_dir, seed=0):
    return np.load(os.path.join(golden_dir, f"pf_seed{seed}.npz"))


def _logw32(pw):
    with np.errstate(divide="ignore"):
        return np.log(np.asarray(pw, dtype=np.float64).reshape(-1)).astype(np.float32)


def _weights_from_logw32(lw32):
    w = np.exp(lw32.astype(np.float64))
    return w / w.sum()


def _assert_weights_close(got, want, rtol=RTOL, rtol_tail=None):
    """north_star's 1e-5 relative on every weight down to 1e-12 of the largest (measured at C1, tools/parity_margins.py:
    1e-6 on the weights above 1e-4 of the largest, 2.8e-6 down to 1e-12, 6.3e-6 down to 1e-20 -- the fp32 storage of
    the log-weights costs ~1e-7 per unit of log-weight).  The fp32 pair kernels pass `rtol_tail` for the weights
    between 1e-12 and 1e-4 of the largest.  Below 1e-12 of the largest a weight is under the resolution of the
    resampler's 2^-52 grid and the MUFU path eventually flushes it to zero: absolute floor."""
    got, want = np.asarray(got).reshape(-1), np.asarray(want).reshape(-1)
    heavy = want > want.max() * 1e-4
    np.testing.assert_allclose(got[heavy], want[heavy], rtol=rtol)
    big = want > want.max() * 1e-12
    np.testing.assert_allclose(got[big], want[big], rtol=rtol if rtol_tail is None else rtol_tail)
    np.testing.assert_allclose(got[~big], want[~big], rtol=1e-3, atol=1e-25)


def _oracle_predict_weight(px32, w_prev, v32, z_l, P, sx=0.3, sy=0.3):
    """fp64 oracle on the fp32 inputs the GPU saw; likelihood evaluated at the GPU's stored pose."""
    px_pred = O.predict(px32.astype(np.float64), v32.astype(np.float64), P)
    return px_pred


@pytest.mark.parametrize("likelihood", ["pair", "moment", "pair32"])
@pytest.mark.parametrize("seed,k", [(0, 0), (0, 1), (0, 5), (0, 7), (0, 100), (0, 359), (1, 39), (2, 39)])
def test_predict_weight_stage_parity(mcl, golden_dir, likelihood, seed, k):
    """__predict + __likelihood + __normalize + MAP (particle_filter.py:156-198, :115-117) on the
    reference's own arrays at the stored steps (no resampling at these steps)."""
    g = _golden(golden_dir, seed)
    if g["resampled"][k]:
        pytest.skip("resampling step: covered by test_full_step_parity")
    P = O.PFParams(100)
    px0 = g[f"s{k}_px_before"].astype(np.float32)
    lw0 = _logw32(g[f"s{k}_pw_before"])
    v32 = g[f"s{k}_v"].astype(np.float32)
    z_l = g["z_l"][k]
    f = mcl.Filter(n_particles=1000, landmarks=P.LM, likelihood=likelihood)
    f.set_particles(px0[0], px0[1], px0[2], lw0)
    f.predict_weight(z_l[None], noise=v32[:, None, :])
    x, y, th, lw = f.get_particles()
    got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
    # poses: fp64 oracle on the same fp32 inputs
    want_px = O.predict(px0.astype(np.float64), v32.astype(np.float64), P)
    np.testing.assert_allclose(got_px, want_px, rtol=RTOL, atol=1e-6)
    # and against the reference's own fp64 arrays (inputs rounded to fp32: ~1e-6 absolute)
    np.testing.assert_allclose(got_px, g[f"s{k}_px"], rtol=RTOL, atol=5e-6)
    # weights: oracle likelihood at the stored fp32 poses times the stored previous weights
    w_prev = _weights_from_logw32(lw0)
    ll = O.loglik_cartesian(got_px, z_l, P.LM, 0.3, 0.3)
    with np.errstate(divide="ignore"):
        want_w, _ = O.update_weights_log(np.log(w_prev), ll)
    got_w = f.weights()[0]
    _assert_weights_close(got_w, want_w, rtol_tail=3e-5 if likelihood == "pair32" else None)
    assert abs(got_w.sum() - 1.0) < 1e-6
    st = f.stats()[0]
    assert st.ess == pytest.approx(O.ess(want_w), rel=RTOL)
    # the reference's own weights (fp64 poses): looser, fp32 pose rounding moves log-weights by ~1e-4
    _assert_weights_close(got_w, g[f"s{k}_pw"], rtol=2e-3)
    # MAP estimate: arg-max of the oracle weights unless the top two are closer than the tolerance
    mi = int(np.argmax(want_w))
    assert st.max_idx == mi or want_w[st.max_idx] >= want_w[mi] * (1 - RTOL)
    assert st.max_idx == int(g["max_idx"][k])
    assert st.max_weight == pytest.approx(want_w[st.max_idx], rel=RTOL)
    np.testing.assert_allclose(np.array(st.est[:]), got_px[:, st.max_idx], rtol=0, atol=0)
    assert st.resampled == 0


@pytest.mark.parametrize("seed,k", [(0, 6), (0, 10), (1, 5), (2, 4)])
def test_full_step_parity_with_resampling(mcl, golden_dir, seed, k):
    """main_pf (:86-119) at steps where the reference resampled: indices bit-exact on the GPU's own
    quantised weights, and identical to the reference's fp64 walk except at razor-edge thresholds."""
    g = _golden(golden_dir, seed)
    assert g["resampled"][k]
    P = O.PFParams(100)
    px0 = g[f"s{k}_px_before"].astype(np.float32)
    lw0 = _logw32(g[f"s{k}_pw_before"])
    v32 = g[f"s{k}_v"].astype(np.float32)
    ofs = float(g["ofs"][k])
    z_l = g["z_l"][k]
    f = mcl.Filter(n_particles=1000, landmarks=P.LM)
    f.set_particles(px0[0], px0[1], px0[2], lw0)
    st0 = f.stats()[0]
    assert st0.ess < P.ESS_TH  # the gate of :211 fires on the device too
    q = f.qweights()[0]
    f.step(z_l[None], ofs=ofs, noise=v32[:, None, :])  # ESS-gated, not forced
    anc = f.ancestors()[0].astype(np.int64)
    st = f.stats()[0]
    assert st.resampled == 1
    # bit-exact against the literal reference walk fed the same (quantised) weights
    wq = O.dequantize_weights(q)
    np.testing.assert_array_equal(anc, O.resample_indices_literal(wq, ofs))
    # against the reference's walk on its un-quantised fp64 weights
    a_ref = O.resample_indices(g[f"s{k}_pw_before"], ofs, clamp=False)
    assert (anc != a_ref).mean() <= 0.003
    # poses after gather + predict
    x, y, th, _ = f.get_particles()
    got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
    want_px = O.predict(px0.astype(np.float64)[:, anc], v32.astype(np.float64), P)
    np.testing.assert_allclose(got_px, want_px, rtol=RTOL, atol=1e-6)
    same = anc == a_ref
    np.testing.assert_allclose(got_px[:, same], g[f"s{k}_px"][:, same], rtol=RTOL, atol=5e-6)
    # weights restart from 1/NP (:222)
    ll = O.loglik_cartesian(got_px, z_l, P.LM, 0.3, 0.3)
    want_w, _ = O.update_weights_log(np.full(1000, -np.log(1000.0)), ll)
    _assert_weights_close(f.weights()[0], want_w)
    assert st.max_idx == int(np.argmax(want_w))


@pytest.mark.parametrize("n", [1000, 4097, 16384, 1 << 20, 1000003])
@pytest.mark.parametrize("kind", ["random", "peaked", "degenerate"])
def test_resampling_indices_bit_exact(mcl, n, kind):
    """__resampling (:200-224): ancestors equal the reference walk on the same weights, bit for bit."""
    rng = np.random.RandomState(n % 9973 + len(kind))
    if kind == "random":
        lw = rng.normal(0, 1.5, n)
    elif kind == "peaked":
        lw = -0.5 * ((np.arange(n) - 0.37 * n) / (0.001 * n + 3)) ** 2 + rng.normal(0, 0.1, n)
    else:
        lw = np.full(n, -60.0)
        lw[rng.randint(0, n, 7)] = rng.normal(0, 1, 7)
    lw = (lw - lw.max()).astype(np.float32)
    pos = rng.normal(0, 1, (3, n)).astype(np.float32)
    f = mcl.Filter(n_particles=n, landmarks=[[0.0, 0.0]])
    f.set_particles(pos[0], pos[1], pos[2], lw)
    q = f.qweights()[0]
    w = f.weights()[0]
    want_w = _weights_from_logw32(lw)
    _assert_weights_close(w, want_w)
    assert np.array_equal(q, O.quantize_weights(w))  # the scan sees exactly rint(w * 2^52)
    ofs = rng.rand() * (1 / n)
    f.resample(ofs=ofs, force=True)
    anc = f.ancestors()[0].astype(np.int64)
    want = O.resample_indices(O.dequantize_weights(q), ofs, clamp=True)
    np.testing.assert_array_equal(anc, want)
    if n <= 4097:
        wq = O.dequantize_weights(q)
        try:
            lit = O.resample_indices_literal(wq, ofs)
            np.testing.assert_array_equal(anc, lit)
        except IndexError:
            pass  # the reference walk ran off the end of pw_cum; ours clamps
    # size-independent properties: monotone ancestors, offspring counts within 1 of N w
    assert np.all(np.diff(anc) >= 0)
    counts = np.bincount(anc, minlength=n)
    assert np.all(np.abs(counts - n * w) < 1.0 + 1e-6 * n * w + 1e-3)
    # the gather moved the poses and reset the weights to 1/NP (:216-222)
    x, y, th, lw1 = f.get_particles()
    np.testing.assert_array_equal(x[0], pos[0][anc])
    np.testing.assert_array_equal(th[0], pos[2][anc])
    assert np.all(lw1[0] == lw1[0][0])  # weights reset to 1/NP (:222): all stored log-weights equal
    np.testing.assert_allclose(f.weights()[0], 1.0 / n, rtol=1e-6)
    assert f.stats()[0].ess == pytest.approx(n)


@pytest.mark.parametrize("n", [1000, 5000, 70001])
@pytest.mark.parametrize("first", ["normal", "zero"])
def test_resampling_with_a_zero_comb_offset(mcl, n, first):
    """rand() may return exactly 0: the tooth u_0 = 0 then sits ON the lower end of particle 0's cumulative weight, and
    the reference walk (strict >, :219) gives slot 0 to particle 0 -- even when that particle's weight is 0."""
    rng = np.random.RandomState(n)
    lw = rng.normal(0, 1.5, n)
    if first == "zero":
        lw[0] = -np.inf
    lw = (lw - lw[np.isfinite(lw)].max()).astype(np.float32)
    pos = rng.normal(0, 1, (3, n)).astype(np.float32)
    f = mcl.Filter(n_particles=n, landmarks=[[0.0, 0.0]])
    f.set_particles(pos[0], pos[1], pos[2], lw)
    q = f.qweights()[0]
    f.resample(ofs=0.0, force=True)
    anc = f.ancestors()[0].astype(np.int64)
    wq = O.dequantize_weights(q)
    np.testing.assert_array_equal(anc, O.resample_indices(wq, 0.0, clamp=True))
    assert anc[0] == 0
    if n <= 5000:
        try:
            np.testing.assert_array_equal(anc, O.resample_indices_literal(wq, 0.0))
        except IndexError:
            pass  # the reference walk ran off the end of pw_cum; ours clamps


def test_scan_descriptors_survive_a_skipped_resampling(mcl):
    """More than 1024 scan tiles (look-back through prefixes), device ESS gate: resample -> gate holds -> resample.
    The launch whose gate held must still clear the descriptor buffer the third launch scans with."""
    n = 4_500_003
    rng = np.random.RandomState(5)
    pos = rng.normal(0, 1, (3, n)).astype(np.float32)
    f = mcl.Filter(n_particles=n, landmarks=[[0.0, 0.0]])
    for rnd, peaked in enumerate([True, False, True]):
        if peaked:
            lw = -0.5 * ((np.arange(n) - (0.3 + 0.2 * rnd) * n) / (0.0005 * n)) ** 2 + rng.normal(0, 0.5, n)
        else:
            lw = rng.normal(0, 0.1, n)
        lw = (lw - lw.max()).astype(np.float32)
        f.set_particles(pos[0], pos[1], pos[2], lw)
        q = f.qweights()[0]
        ess = f.stats()[0].ess
        ofs = rng.rand() * (1 / n)
        f.resample(ofs=ofs)  # gate on the device: ess < n / 100
        assert (ess < n / 100) == peaked
        if peaked:
            anc = f.ancestors()[0].astype(np.int64)
            want = O.resample_indices(O.dequantize_weights(q), ofs, clamp=True)
            np.testing.assert_array_equal(anc, want)
            x, _, _, _ = f.get_particles()
            np.testing.assert_array_equal(x[0], pos[0][anc])
        else:
            x, _, _, lw1 = f.get_particles()
            np.testing.assert_array_equal(x[0], pos[0])
            np.testing.assert_array_equal(lw1[0], lw)


def test_ess_gate_skips_resampling(mcl):
    n = 2048
    rng = np.random.RandomState(5)
    pos = rng.normal(0, 1, (3, n)).astype(np.float32)
    lw = rng.normal(0, 0.1, n).astype(np.float32)  # ESS ~ n: far above n/100
    f = mcl.Filter(n_particles=n, landmarks=[[0.0, 0.0]])
    f.set_particles(pos[0], pos[1], pos[2], lw)
    f.resample(ofs=0.3 / n)
    x, y, th, lw1 = f.get_particles()
    np.testing.assert_array_equal(x[0], pos[0])
    np.testing.assert_array_equal(lw1[0], lw)


def test_weighted_mean_and_covariance(mcl, golden_dir):
    """Extension a-14: moments that feed ErrorEllipse.calc_error_ellipse (mylib/error_ellipse.py:39-55)."""
    g = _golden(golden_dir, 0)
    P = O.PFParams(100)
    for k in (5, 100, 359):
        px = g[f"s{k}_px"].astype(np.float32)
        lw = _logw32(g[f"s{k}_pw"])
        f = mcl.Filter(n_particles=1000, landmarks=P.LM)
        f.set_particles(px[0], px[1], px[2], lw)
        w = _weights_from_logw32(lw)
        mean, cov = O.weighted_moments(px.astype(np.float64), w)
        st = f.stats()[0]
        np.testing.assert_allclose(np.array(st.mean[:]), mean, rtol=RTOL, atol=1e-7)
        got = np.array(st.cov[:]).reshape(3, 3)
        scale = np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
        np.testing.assert_allclose(got, cov, rtol=RTOL, atol=RTOL * scale.min())
        f.pose_moments()  # exact two-pass variant
        got2 = np.array(f.stats()[0].cov[:]).reshape(3, 3)
        np.testing.assert_allclose(got2, cov, rtol=RTOL, atol=1e-6 * scale.min())


def test_yaw_seam_mean(mcl):
    """Particles straddling +-pi: circular mean and wrapped residuals (mylib/limit.py:11-26)."""
    n = 4096
    rng = np.random.RandomState(11)
    th = O.limit_angle_vec(np.pi + rng.normal(0, 0.05, n)).astype(np.float32)
    x = rng.normal(2, 0.1, n).astype(np.float32)
    y = rng.normal(-3, 0.2, n).astype(np.float32)
    lw = rng.normal(0, 0.5, n).astype(np.float32)
    f = mcl.Filter(n_particles=n, landmarks=[[0.0, 0.0]])
    f.set_particles(x, y, th, lw)
    mean, cov = O.weighted_moments(np.stack([x, y, th]).astype(np.float64), _weights_from_logw32(lw))
    st = f.stats()[0]
    assert abs(O.limit_angle(st.mean[2] - mean[2])) < 1e-6
    scale = np.sqrt(np.outer(np.diag(cov), np.diag(cov)))  # (the cross terms are ~0: absolute floor relative to the variances)
    np.testing.assert_allclose(np.array(st.cov[:]).reshape(3, 3), cov, rtol=RTOL, atol=RTOL * scale.min())


def test_dropin_seeded_trajectory(mcl, golden_dir):
    """ParticleFilter(100).main_pf() under np.random.seed(0): same RNG consumption as the reference
    (x_true and the resampling schedule depend on it), estimates tracking the reference run."""
    g = _golden(golden_dir, 0)
    np.random.seed(0)
    pf = mcl.ParticleFilter(100)
    n_steps = 120
    resampled, err_gpu, err_ref = [], [], []
    for k in range(n_steps):
        LM, x_true, x_est, px, Q, max_idx, max_val = pf.main_pf()
        assert LM.shape == (5, 2) and x_true.shape == (3, 1) and x_est.shape == (3, 1)
        assert px.shape == (3, 1000) and px.dtype == np.float64 and Q.shape == (3, 3)
        np.testing.assert_allclose(x_true[:, 0], g["x_true"][k], atol=1e-12)
        resampled.append(pf.last_stats.resampled)
        if k < 6:  # before the first resampling fp32 rounding has not yet changed any decision
            assert max_idx == g["max_idx"][k]
            assert max_val == pytest.approx(g["max_val"][k], rel=2e-3)
            np.testing.assert_allclose(x_est[:, 0], g["x_est"][k], atol=1e-5)
            np.testing.assert_allclose(px.mean(axis=1), g["px_mean"][k], atol=1e-5)
            assert pf.last_stats.ess == pytest.approx(g["ess_after"][k], rel=2e-3)
        err_gpu.append(np.hypot(*(x_est[:2, 0] - x_true[:2, 0])))
        err_ref.append(np.hypot(*(g["x_est"][k][:2] - g["x_true"][k][:2])))
    # identical schedule while the fp32 and fp64 particle sets have not yet diverged (a free-running
    # filter amplifies rounding through resampling); afterwards the same resampling rate
    assert resampled[:48] == [int(b) for b in g["resampled"][:48]]
    assert abs(sum(resampled) - int(g["resampled"][:n_steps].sum())) <= 4
    assert np.mean(err_gpu) < 1.5 * np.mean(err_ref) + 0.05


@pytest.mark.parametrize("T,n", [(5, 3000), (96, 30001), (700, 4100)])
def test_trials_batch_equals_single_filters(mcl, T, n):
    """n_trials independent filters in one handle == the same filters one by one (bit-exact).  (5, 3000): a handful of
    tiles, the 256 x 8 scan geometry; (96, 30001): 1440 tiles shared by persistent CTAs of the 128 x 16 geometry, several
    CTAs per trial (tickets, look-back); (700, 4100): one CTA per trial carrying the prefix in a register."""
    Lm = 7
    rng = np.random.RandomState(2)
    lm = rng.uniform(-10, 10, (Lm, 2))
    pos = rng.normal(0, 0.3, (3, T, n)).astype(np.float32) + np.array([3.0, 1.0, 0.5], dtype=np.float32)[:, None, None]
    lw = rng.normal(0, 1.0, (T, n)).astype(np.float32)
    noise = (rng.normal(0, 1, (3, T, n)) * np.array([0.03, 0.03, 0.035])[:, None, None]).astype(np.float32)
    z = rng.normal(0, 3, (T, Lm, 2))
    ofs = rng.rand(T) / n
    fb = mcl.Filter(n_particles=n, n_trials=T, landmarks=lm)
    fb.set_particles(pos[0], pos[1], pos[2], lw)
    fb.step(z, ofs=ofs, noise=noise, force_resample=True)
    xb, yb, tb, lb = fb.get_particles()
    ancb = fb.ancestors()
    sb = fb.stats()
    for t in (range(T) if T <= 8 else list(range(0, T, max(1, T // 6))) + [T - 1]):
        f1 = mcl.Filter(n_particles=n, landmarks=lm)
        f1.set_particles(pos[0, t], pos[1, t], pos[2, t], lw[t])
        f1.step(z[t][None], ofs=ofs[t], noise=noise[:, t:t + 1], force_resample=True)
        x1, y1, t1, l1 = f1.get_particles()
        np.testing.assert_array_equal(ancb[t], f1.ancestors()[0])
        np.testing.assert_array_equal(xb[t], x1[0])
        np.testing.assert_array_equal(tb[t], t1[0])
        np.testing.assert_array_equal(lb[t], l1[0])
        s1 = f1.stats()[0]
        assert sb[t].max_idx == s1.max_idx and sb[t].ess == s1.ess and sb[t].log_norm == s1.log_norm


@pytest.mark.parametrize("likelihood", ["pair", "moment", "pair32"])
def test_c1_size_step_parity(mcl, likelihood):
    """C1 shape (2^20 particles x 16 landmarks), forced resampling, injected noise: full step vs oracle."""
    n, Lm = 1 << 20, 16
    rng = np.random.RandomState(1234)
    lm = np.random.default_rng(1234).uniform(-20, 20, (Lm, 2))
    P = O.PFParams(100, n_particles=n, landmarks=lm)
    truth = np.array([[10.0], [0.0], [np.pi / 2]])
    pos = (truth + rng.normal(0, 1, (3, n)) * np.array([[0.08], [0.08], [0.03]])).astype(np.float32)
    z = O.world2robot(truth, lm) + rng.normal(0, 0.3, (Lm, 2))
    # weights of a previous step
    ll0 = O.loglik_cartesian(pos.astype(np.float64), z, lm, 0.3, 0.3)
    lw = (ll0 - ll0.max()).astype(np.float32)
    noise = (rng.normal(0, 1, (3, n)) * np.sqrt(np.diag(P.Q))[:, None]).astype(np.float32)
    truth1 = O.f_motion(truth, P)
    z1 = O.world2robot(truth1, lm) + rng.normal(0, 0.3, (Lm, 2))
    ofs = rng.rand() / n
    f = mcl.Filter(n_particles=n, landmarks=lm, likelihood=likelihood)
    f.set_particles(pos[0], pos[1], pos[2], lw)
    q = f.qweights()[0]
    f.step(z1[None], ofs=ofs, noise=noise[:, None, :], force_resample=True)
    anc = f.ancestors()[0].astype(np.int64)
    np.testing.assert_array_equal(anc, O.resample_indices(O.dequantize_weights(q), ofs))
    x, y, th, _ = f.get_particles()
    got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
    want_px = O.predict(pos.astype(np.float64)[:, anc], noise.astype(np.float64), P)
    np.testing.assert_allclose(got_px, want_px, rtol=RTOL, atol=1e-6)
    ll = O.loglik_cartesian(got_px, z1, lm, 0.3, 0.3)
    want_w, logz = O.update_weights_log(np.full(n, -np.log(float(n))), ll)
    got_w = f.weights()[0]
    _assert_weights_close(got_w, want_w, rtol_tail=3e-5 if likelihood == "pair32" else None)
    st = f.stats()[0]
    assert st.ess == pytest.approx(O.ess(want_w), rel=RTOL)
    mi = int(np.argmax(want_w))
    assert st.max_idx == mi or want_w[st.max_idx] >= want_w[mi] * (1 - RTOL)
    mean, cov = O.weighted_moments(got_px, want_w)
    np.testing.assert_allclose(np.array(st.mean[:]), mean, rtol=RTOL, atol=1e-6)
    # the FUSED single-pass covariance (fp32 group sums folded into fp64, about the reference pose): measured 1.2e-6
    np.testing.assert_allclose(np.array(st.cov[:]).reshape(3, 3), cov, rtol=RTOL, atol=1e-9)


@pytest.mark.parametrize("scenario", ["running", "offset"])
def test_likelihood_64_and_1024_landmarks(mcl, scenario):
    """L = 64 (C3) and L = 1024 (C4): pair kernel == moment kernel == fp32 pair kernel == log-domain oracle.

    "running": the observation is taken where the cloud is after its motion step, as in a running filter -- the
    reference pose (the previous MAP estimate pushed through the motion) then sits at the likelihood peak and every
    kernel meets 1e-5.  "offset": the observation is taken BEFORE the motion step, so the whole cloud sits 0.17 m
    off and the useful particles score e^+80 (L = 64) to e^+760 (L = 1024) above the reference pose.  Log-weights are
    STORED in fp32 relative to the reference pose, and an fp32 ulp at 760 is 6e-5: outside the regime the relative
    storage is designed for, every kernel is asked 5e-5 / 1e-4 instead of 1e-5."""
    rng = np.random.RandomState(99)
    for Lm, n in ((64, 1 << 16), (1024, 1 << 13)):
        lm = rng.uniform(-20, 20, (Lm, 2))
        truth = np.array([[3.0], [-2.0], [0.7]])
        spread = 0.3 / np.sqrt(Lm)
        pos = (truth + rng.normal(0, 1, (3, n)) * np.array([[spread], [spread], [spread / 8]])).astype(np.float32)
        # the engine keeps log-weights relative to the likelihood of a reference pose (in a running filter
        # the previous MAP estimate; after set_particles, particle 0): put a good particle there, as the
        # filter would, so the fp32 log-weights of the heavy particles are small numbers
        pos[:, 0] = truth[:, 0].astype(np.float32)
        P = O.PFParams(100, n_particles=n, landmarks=lm)
        seen_from = O.f_motion(truth, P) if scenario == "running" else truth
        z = O.world2robot(seen_from, lm) + rng.normal(0, 0.3, (Lm, 2))
        noise = np.zeros((3, 1, n), dtype=np.float32)
        res, mirrored, flags = {}, {}, {}
        for lik in ("pair", "moment", "pair32"):
            for rule in (False, True):
                f = mcl.Filter(n_particles=n, landmarks=lm, likelihood=lik, max_landmarks=Lm, underflow_rule=rule)
                f.set_particles(pos[0], pos[1], pos[2], None)
                f.predict_weight(z[None], noise=noise)
                x, y, th, _ = f.get_particles()
                got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
                (mirrored if rule else res)[lik] = f.weights()[0]
                flags[lik, rule] = f.stats()[0].underflow
        # pure log-domain weights (underflow_rule = 0) against the log-domain oracle
        ll = O.loglik_cartesian(got_px, z, lm, 0.3, 0.3)
        want_w, _ = O.update_weights_log(np.full(n, -np.log(float(n))), ll)
        if scenario == "running":
            _assert_weights_close(res["pair"], want_w)
            _assert_weights_close(res["moment"], want_w)
            _assert_weights_close(res["pair32"], want_w, rtol_tail=3e-5)
        else:  # stored fp32 log-weights of +80 / +760: half an ulp is 4e-6 / 3e-5, for every kernel
            for lik in ("pair", "moment", "pair32"):
                _assert_weights_close(res[lik], want_w, rtol=5e-5, rtol_tail=1e-4)
        # the default mirrors the reference's fp64 PRODUCT (:192-196, :234-237): with 1024 landmarks 0.17 m off, every
        # product of pdfs underflows, the reference's weights are NaN -> 1/NP, and so are the engine's
        lit_w = O.normalize(np.full(n, 1.0 / n) * O.likelihood_product(got_px, z, lm, 0.3, 0.3), 1.0 / n)
        ref_uniform = bool(np.all(lit_w == 1.0 / n))
        assert ref_uniform == (scenario == "offset" and Lm == 1024)
        for lik in ("pair", "moment", "pair32"):
            assert flags[lik, False] == 0 and flags[lik, True] == int(ref_uniform)
            if ref_uniform:
                np.testing.assert_allclose(mirrored[lik], 1.0 / n, rtol=1e-6)
            else:
                np.testing.assert_array_equal(mirrored[lik], res[lik])


def test_motion_b_and_rb_likelihood(mcl, golden_dir):
    """Extensions a-15 (MotionModel.moveWithNoise, motion_model.py:31-62) and a-16 (range/bearing)."""
    n, Lm = 8192, 9
    rng = np.random.RandomState(17)
    lm = rng.uniform(-8, 8, (Lm, 2))
    truth = np.array([[1.0], [2.0], [0.4]])
    pos = (truth + rng.normal(0, 1, (3, n)) * np.array([[0.1], [0.1], [0.03]])).astype(np.float32)
    normals = rng.normal(0, 1, (3, n)).astype(np.float32)
    v, w, dt = 10 * np.deg2rad(10.0), np.deg2rad(10.0), 2.0
    z_rb = O.rb_observation(truth, lm) + rng.normal(0, 1, (Lm, 2)) * np.array([0.1, 0.03])
    f = mcl.Filter(n_particles=n, landmarks=lm, likelihood="rb", motion="b", sigma_obs=[0.1, 0.03], dt=dt, vel=v,
                   yaw_rate=w, motion_b_alpha=[0.1] * 6, max_landmarks=16)
    f.set_particles(pos[0], pos[1], pos[2], None)
    f.predict_weight(z_rb[None], noise=normals[:, None, :])
    x, y, th, _ = f.get_particles()
    got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
    want_px = O.motion_b(pos.astype(np.float64), v, w, dt, [0.1] * 6, normals.astype(np.float64))
    np.testing.assert_allclose(got_px, want_px, rtol=RTOL, atol=2e-6)
    ll = O.loglik_rb(got_px, z_rb, lm, 0.1, 0.03)
    want_w, _ = O.update_weights_log(np.full(n, -np.log(float(n))), ll)
    _assert_weights_close(f.weights()[0], want_w)


@pytest.mark.parametrize("seam", [False, True])
@pytest.mark.parametrize("Lm,n", [(9, 8192), (64, 1 << 15), (1024, 1 << 13)])
def test_rb32_all_landmark_likelihood_matches_oracle(mcl, Lm, n, seam):
    """BASELINE configs[4] as named: EVERY range/bearing landmark scored, fp32 + SFU (MCL_LIK_RB_F32), against the fp64
    restatement loglik_rb (sensor model graph_based_slam.py:150-167, wrap mylib/limit.py:19-26; parity unpinned: the
    reference has no range/bearing likelihood) and against the fp64 kernel (MCL_LIK_RB).  The kernel works about the
    reference pose, so its error is relative to |log w - log w_ref|: 2e-5 on the weights that carry mass.
    seam: one landmark's bearing residual sits at +-pi, so limit_angle flips its sign for part of the cloud.  fp32
    resolves an angle near pi to 2.4e-7 rad and the squared residual jumps by 2 pi (2 e) / sigma^2 ~ 4e4 per radian
    across the seam: only the wrap DECISION can be checked there (1e-2; a wrong turn is a factor e^+-1000)."""
    rng = np.random.RandomState(1000 + Lm)
    lm = rng.uniform(-20, 20, (Lm, 2))
    truth = np.array([[3.0], [-2.0], [0.7]])
    P = O.PFParams(100, n_particles=n, landmarks=lm)
    spread = 0.2 / np.sqrt(Lm)
    pos = (truth + rng.normal(0, 1, (3, n)) * np.array([[spread], [spread], [spread / 8]])).astype(np.float32)
    pos[:, 0] = truth[:, 0].astype(np.float32)
    moved = O.f_motion(truth, P)
    z = O.rb_observation(moved, lm) + rng.normal(0, 1, (Lm, 2)) * np.array([0.1, 0.03])
    z[:, 1] = O.limit_angle_vec(z[:, 1])
    if seam:
        z[0, 1] = O.limit_angle_vec(np.array([O.rb_observation(moved, lm)[0, 1] + np.pi - 2e-4]))[0]
    noise = np.zeros((3, 1, n), dtype=np.float32)
    res = {}
    for lik in ("rb", "rb32"):
        f = mcl.Filter(n_particles=n, landmarks=lm, likelihood=lik, sigma_obs=[0.1, 0.03], max_landmarks=Lm)
        f.set_particles(pos[0], pos[1], pos[2], None)
        f.predict_weight(z[None], noise=noise)
        x, y, th, _ = f.get_particles()
        got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
        res[lik] = f.weights()[0]
    ll = O.loglik_rb(got_px, z, lm, 0.1, 0.03)
    want_w, _ = O.update_weights_log(np.full(n, -np.log(float(n))), ll)
    _assert_weights_close(res["rb"], want_w)
    heavy = want_w > want_w.max() * 1e-4
    big = want_w > want_w.max() * 1e-12
    print(f"L={Lm} seam={seam}: rb32 max rel err heavy {np.abs(res['rb32'][heavy] / want_w[heavy] - 1).max():.2e}, "
          f"down to 1e-12 {np.abs(res['rb32'][big] / want_w[big] - 1).max():.2e}")
    np.testing.assert_allclose(res["rb32"][heavy], want_w[heavy], rtol=1e-2 if seam else 1e-5)
    np.testing.assert_allclose(res["rb32"][big], want_w[big], rtol=3e-2 if seam else 2e-5)


def test_rb32_far_particles_take_the_general_forms(mcl):
    """A cloud metres and tens of degrees away from the reference pose leaves the small-angle series: the kernel falls
    back to its minimax atan2 and to limit_angle's +-2 pi turn per pair.  The weights are e^-10^4 there, so the
    LOG-weights are compared, relative to particle 0 (the reference-pose particle): 2e-5 of the difference (measured 3.3e-6)."""
    Lm, n = 64, 8192
    rng = np.random.RandomState(77)
    lm = rng.uniform(-20, 20, (Lm, 2))
    truth = np.array([[3.0], [-2.0], [0.7]])
    P = O.PFParams(100, n_particles=n, landmarks=lm)
    pos = (truth + rng.normal(0, 1, (3, n)) * np.array([[1.5], [1.5], [0.6]])).astype(np.float32)
    pos[:, 0] = truth[:, 0].astype(np.float32)
    moved = O.f_motion(truth, P)
    z = O.rb_observation(moved, lm) + rng.normal(0, 1, (Lm, 2)) * np.array([0.1, 0.03])
    z[:, 1] = O.limit_angle_vec(z[:, 1])
    noise = np.zeros((3, 1, n), dtype=np.float32)
    f = mcl.Filter(n_particles=n, landmarks=lm, likelihood="rb32", sigma_obs=[0.1, 0.03], max_landmarks=Lm,
                   underflow_rule=False)
    f.set_particles(pos[0], pos[1], pos[2], None)
    f.predict_weight(z[None], noise=noise)
    x, y, th, lw = f.get_particles()
    got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
    ll = O.loglik_rb(got_px, z, lm, 0.1, 0.03)
    got, want = lw[0].astype(np.float64) - float(lw[0][0]), ll - ll[0]
    assert want.min() < -1e4  # the cloud really is far out
    err = np.abs(got - want) / (np.abs(want) + 1.0)
    print(f"rb32 far cloud: log-weight differences down to {want.min():.3g}, max relative error {err.max():.2e}")
    assert err.max() < 2e-5


def test_rb32_landmark_capacity(mcl):
    """MCL_LIK_RB_F32 stages 64 bytes per landmark in shared memory: 2048 landmarks run, 2049 are refused."""
    rng = np.random.RandomState(5)
    n = 4096
    lm = rng.uniform(-30, 30, (2049, 2))
    f = mcl.Filter(n_particles=n, landmarks=lm[:2048], likelihood="rb32", sigma_obs=[0.1, 0.03], max_landmarks=2049,
                   underflow_rule=False)
    g = mcl.Filter(n_particles=n, landmarks=lm[:2048], likelihood="rb", sigma_obs=[0.1, 0.03], max_landmarks=2049,
                   underflow_rule=False)
    truth = np.array([[1.0], [2.0], [-0.4]])
    pos = (truth + rng.normal(0, 1, (3, n)) * np.array([[0.004], [0.004], [0.0005]])).astype(np.float32)
    z = O.rb_observation(O.f_motion(truth, O.PFParams(100)), lm[:2048])[None]
    noise = np.zeros((3, 1, n), dtype=np.float32)
    for flt in (f, g):
        flt.set_particles(pos[0], pos[1], pos[2], None)
        flt.predict_weight(z, noise=noise)
    w32, w64 = f.weights()[0], g.weights()[0]
    heavy = w64 > w64.max() * 1e-4
    np.testing.assert_allclose(w32[heavy], w64[heavy], rtol=2e-5)
    with pytest.raises(RuntimeError, match="2048"):
        f.set_landmarks(lm)


def test_steps_captured_in_a_cuda_graph(mcl):
    """mcl_step_dev touches no host memory and synchronises nothing: a run of steps can be captured in a CUDA graph
    (programmatic-launch edges included) and launched as one unit.  The graph bakes in the buffer parity and the
    step counters of the steps it holds, so it stands for THOSE steps: launched once it must leave the filter in
    exactly the state the same calls leave it in eagerly."""
    import torch
    n, Lm, K = 50000, 5, 6
    lm = np.random.default_rng(3).uniform(-15, 15, (Lm, 2))
    rng = np.random.RandomState(8)
    z = torch.from_numpy(rng.normal(0, 3, (K, 1, Lm, 2))).cuda()

    def fresh():
        f = mcl.Filter(n_particles=n, landmarks=lm, seed=11)
        f.init_particles([1.0, 2.0, 0.3])
        return f

    eager = fresh()
    for k in range(K):
        eager.step_dev(z[k], force_resample=(k % 2 == 0))
    want = eager.get_particles()
    want_stats = eager.stats()[0]

    graphed = fresh()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for k in range(K):
                graphed.step_dev(z[k], force_resample=(k % 2 == 0))
    g.replay()
    torch.cuda.synchronize()
    got = graphed.get_particles()
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    got_stats = graphed.stats()[0]
    assert got_stats.max_idx == want_stats.max_idx and got_stats.ess == want_stats.ess


def test_philox_noise_statistics_and_determinism(mcl):
    """Device noise cannot be bit-compared with MT19937: check moments, independence of launch
    geometry (trial batching / particle offset) and run-to-run determinism."""
    n = 1 << 18
    lm = [[5.0, 5.0], [2.0, -3.0], [-3.0, 4.0]]
    z = np.zeros((1, 3, 2))

    def run(seed, **kw):
        f = mcl.Filter(n_particles=n, landmarks=lm, seed=seed, **kw)
        f.init_particles([0.0, 0.0, 0.0])
        f.step(z, no_resample=True)
        x, y, th, _ = f.get_particles()
        return x[0].astype(np.float64), y[0].astype(np.float64), th[0].astype(np.float64)

    x, y, th = run(7)
    x2, y2, th2 = run(7)
    assert np.array_equal(x, x2) and np.array_equal(th, th2)
    x3, _, _ = run(8)
    assert not np.array_equal(x, x3)
    P = O.PFParams(100)
    det = O.f_motion(np.zeros((3, 1)), P)[:, 0]
    for arr, mu, sd in ((x, det[0], 0.03), (y, det[1], 0.03), (th, det[2], np.deg2rad(2.0))):
        r = (arr - mu) / sd
        assert abs(r.mean()) < 5 / np.sqrt(n)
        assert abs(r.std() - 1) < 5 / np.sqrt(2 * n)
        assert abs((r ** 3).mean()) < 0.03 and abs((r ** 4).mean() - 3) < 0.08
    assert abs(np.corrcoef(x - det[0], y - det[1])[0, 1]) < 0.01
    assert abs(np.corrcoef(x - det[0], th - det[2])[0, 1]) < 0.01
    # a shard holding the second half of a 2n filter draws the noise of global slots [n, 2n)
    ff = mcl.Filter(n_particles=2 * n, landmarks=lm, seed=7)
    ff.init_particles([0.0, 0.0, 0.0])
    ff.step(z, no_resample=True)
    x_full = ff.get_particles()[0][0]
    fs = mcl.Filter(n_particles=n, landmarks=lm, seed=7, n_global=2 * n, slot_base=n)
    fs.init_particles([0.0, 0.0, 0.0])
    fs.step(z, no_resample=True)
    assert np.array_equal(fs.get_particles()[0][0], x_full[n:])
    assert np.array_equal(x_full[:n].astype(np.float64), x)


def test_error_paths(mcl):
    f = mcl.Filter(n_particles=1000, max_landmarks=4)
    with pytest.raises(mcl.MclError):
        f.step(np.zeros((1, 0, 2)))  # no landmarks set
    with pytest.raises(mcl.MclError):
        f.set_landmarks(np.zeros((9, 2)))  # more than max_landmarks
    f.set_landmarks(np.zeros((2, 2)))
    with pytest.raises(mcl.MclError):
        f.step(np.zeros((1, 2, 2)))  # particles not initialised


@pytest.mark.parametrize("G", [2, 4, 8])
@pytest.mark.parametrize("kind", ["spread", "left_heavy"])
def test_sharded_step_matches_unsharded(mcl, G, kind):
    """G shards (G handles on this one GPU, the all-to-all done by hand) take one full step --
    global resampling, predict, weight -- and end with exactly the particles of the unsharded filter."""
    _sharded_step_case(mcl, G, kind, 1 << 16)


def test_sharded_step_with_segmented_tile_prefixes(mcl):
    """The same at a size where mcl_shard_total's CTAs each sum SEVERAL scan tiles (more than 1024 tiles per
    shard: 4.2 M particles), so the expansion starts its tiles from segment prefix + tile prefix."""
    _sharded_step_case(mcl, 2, "spread", 1 << 23)


def _sharded_step_case(mcl, G, kind, n):
    import importlib

    import torch

    sh = importlib.import_module("slam-robot_simu_b200.sharded")
    Lm = 6
    m = n // G
    rng = np.random.RandomState(G + len(kind))
    lm = rng.uniform(-10, 10, (Lm, 2))
    truth = np.array([[2.0], [1.0], [0.4]])
    pos = (rng.normal(0, 1, (3, n)) * np.array([[0.2], [0.2], [0.05]]) + truth).astype(np.float32)
    if kind == "spread":
        lw = (-0.5 * ((np.arange(n) - 0.3 * n) / (0.05 * n)) ** 2 + rng.normal(0, 1.0, n)).astype(np.float32)
    else:  # nearly all the weight in the first shard: it fills slots of every other rank
        lw = (np.where(np.arange(n) < m // 3, 0.0, -30.0) + rng.normal(0, 1.0, n)).astype(np.float32)
    ofs = rng.rand() / n
    noise = (rng.normal(0, 1, (3, 1, n)) * np.array([0.03, 0.03, 0.035])[:, None, None]).astype(np.float32)
    z = O.world2robot(truth, lm) + rng.normal(0, 0.3, (Lm, 2))
    # unsharded
    f1 = mcl.Filter(n_particles=n, landmarks=lm)
    f1.set_particles(pos[0], pos[1], pos[2], lw)
    s1 = f1.stats()[0]
    q1 = f1.qweights()[0]
    f1.step(z[None], ofs=ofs, noise=noise, force_resample=True)
    x1, y1, t1, l1 = f1.get_particles()
    anc1 = f1.ancestors()[0]
    s1b = f1.stats()[0]
    # sharded
    fs = []
    for g in range(G):
        f = mcl.Filter(n_particles=m, landmarks=lm, n_global=n, slot_base=g * m)
        sl = slice(g * m, (g + 1) * m)
        f.set_particles(pos[0, sl], pos[1, sl], pos[2, sl], lw[sl])
        fs.append(f)

    def merged_stats():
        recs = []
        for f in fs:
            d = mcl.stats_to_dict(f.stats()[0])
            d["best_logw"] = d["max_logw"]
            recs.append(d)
        return sh.merge_stats(recs, [g * m for g in range(G)])

    merged = merged_stats()
    assert merged["max_idx"] == s1.max_idx
    assert merged["ess"] == pytest.approx(s1.ess, rel=1e-6)
    assert merged["log_norm"] == pytest.approx(s1.log_norm, abs=1e-7)
    np.testing.assert_allclose(merged["mean"], np.array(s1.mean[:]), rtol=1e-6)
    np.testing.assert_allclose(merged["cov"], np.array(s1.cov[:]).reshape(3, 3), rtol=1e-4, atol=1e-9)
    # the device merge (mcl_shard_merge, what ShardedParticleFilter uses) against its host twin
    recs_dev = torch.cat([f.stats_dev_bytes().clone() for f in fs])
    fs[0].shard_merge(recs_dev, G)
    dm = mcl.stats_to_dict(fs[0].stats()[0])
    assert dm["max_idx"] == merged["max_idx"]
    np.testing.assert_allclose(dm["est"], merged["est"], rtol=0, atol=0)
    for key in ("ess", "log_norm", "sum_w", "sum_w2", "max_logw", "max_weight"):
        assert dm[key] == pytest.approx(merged[key], rel=1e-12, abs=1e-300), key
    np.testing.assert_allclose(dm["mean"], merged["mean"], rtol=1e-12)
    np.testing.assert_allclose(dm["cov"], merged["cov"], rtol=1e-9, atol=1e-15)
    g_stats = mcl.mcl_stats()
    # install the UNSHARDED statistics so the weight map is bit-identical (the merged ones agree to ~1e-9,
    # which very rarely moves the fp32 scale by one ulp)
    g_stats.max_logw, g_stats.sum_w, g_stats.sum_w2 = s1.max_logw, s1.sum_w, s1.sum_w2
    g_stats.log_norm, g_stats.ess = s1.log_norm, s1.ess
    for k in range(3):
        g_stats.est[k] = s1.est[k]
    totals = []
    tot = torch.zeros(1, dtype=torch.int64, device="cuda")
    for g, f in enumerate(fs):
        f.shard_set_global(g_stats)
        f.shard_total(tot)
        totals.append(int(tot.item()))
        assert np.array_equal(f.qweights()[0], q1[g * m:(g + 1) * m])
    assert sum(totals) == int(q1.sum())
    ranges = sh.shard_ranges(totals, n, ofs)
    sends = []
    for g, f in enumerate(fs):
        ex = [torch.full((n,), np.nan, dtype=torch.float32, device="cuda") for _ in range(3)]
        first, count = f.shard_expand(sum(totals[:g]), ofs, *ex)
        assert (first, count) == ranges[g] or (count == 0 and ranges[g][1] == 0)
        mine = np.nonzero((anc1 >= g * m) & (anc1 < (g + 1) * m))[0]
        assert count == mine.size and (count == 0 or first == mine[0])
        sends.append([e.cpu().numpy() for e in ex])
    # the all-to-all, by hand
    plans = [sh.plan_offrank(ranges, n, G, r) for r in range(G)]
    for r, f in enumerate(fs):
        recv = [[], [], []]
        for s_rank in range(G):
            cnt = plans[r][1][s_rank]
            if cnt == 0:
                continue
            off = sum(plans[s_rank][0][:r])  # where rank s_rank's send buffer holds the slots for r
            for d in range(3):
                recv[d].append(sends[s_rank][d][off:off + cnt])
        n_pre, n_suf = plans[r][2], plans[r][3]
        rt = [torch.from_numpy(np.concatenate(v + [np.zeros(1, np.float32)])).cuda() for v in recv]
        f.shard_step(z[None], rt[0], rt[1], rt[2], n_pre, n_suf, noise=noise[:, :, r * m:(r + 1) * m])
        x, y, th, lwn = f.get_particles()
        sl = slice(r * m, (r + 1) * m)
        np.testing.assert_array_equal(x[0], x1[0][sl])
        np.testing.assert_array_equal(y[0], y1[0][sl])
        np.testing.assert_array_equal(th[0], t1[0][sl])
        np.testing.assert_array_equal(lwn[0], l1[0][sl])
    after = merged_stats()
    assert after["max_idx"] == s1b.max_idx and after["ess"] == pytest.approx(s1b.ess, rel=1e-6)
    # capacity too small is reported, not overrun
    heavy = int(np.argmax([sum(p[0]) for p in plans]))
    if sum(plans[heavy][0]) > 8:
        ex = [torch.empty(8, dtype=torch.float32, device="cuda") for _ in range(3)]
        with pytest.raises(mcl.MclError):
            fs[heavy].shard_expand(sum(totals[:heavy]), ofs, *ex)


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_sharded_filter_two_gpus(transport):
    """The real thing: ShardedParticleFilter over 2 GPUs (one process each) == the unsharded filter (Philox noise).
    "peer": device-driven exchange through CUDA IPC peer memory; "nccl": the host-driven fallback."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2); bench.py --gpus N repeats this check in-line (sharded_filter.parity)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29641" if transport == "nccl" else "29643", os.path.join(root, "tests", "_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, MASTER_ADDR="127.0.0.1", MCL_TRANSPORT=transport))
    assert r.returncode == 0 and f"{transport} worker ok" in r.stdout, r.stdout[-3000:] + r.stderr[-4000:]


# ---------------------------------------------------------------------------------------------
# edge cases: tiny and ragged sizes, extreme landmark counts, mixed gates in a trial batch
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 5, 127, 129, 2047, 2049])
def test_tiny_and_ragged_particle_counts(mcl, n):
    """Sizes around the vector width (4), the partition unit (128) and the scan tile (2048)."""
    rng = np.random.RandomState(n)
    lm = rng.uniform(-5, 5, (3, 2))
    P = O.PFParams(100, n_particles=n, landmarks=lm)
    truth = np.array([[1.0], [-1.0], [0.3]])
    pos = (truth + rng.normal(0, 0.1, (3, n))).astype(np.float32)
    lw = rng.normal(0, 1, n).astype(np.float32)
    noise = (rng.normal(0, 1, (3, n)) * np.sqrt(np.diag(P.Q))[:, None]).astype(np.float32)
    z = O.world2robot(truth, lm) + rng.normal(0, 0.3, (3, 2))
    ofs = rng.rand() / n
    f = mcl.Filter(n_particles=n, landmarks=lm)
    f.set_particles(pos[0], pos[1], pos[2], lw)
    q = f.qweights()[0]
    f.step(z[None], ofs=ofs, noise=noise[:, None, :], force_resample=True)
    anc = f.ancestors()[0].astype(np.int64)
    np.testing.assert_array_equal(anc, O.resample_indices(O.dequantize_weights(q), ofs))
    x, y, th, _ = f.get_particles()
    got = np.stack([x[0], y[0], th[0]]).astype(np.float64)
    np.testing.assert_allclose(got, O.predict(pos.astype(np.float64)[:, anc], noise.astype(np.float64), P), rtol=RTOL, atol=1e-6)
    ll = O.loglik_cartesian(got, z, lm, 0.3, 0.3)
    want_w, _ = O.update_weights_log(np.full(n, -np.log(float(n))), ll)
    _assert_weights_close(f.weights()[0], want_w)
    st = f.stats()[0]
    assert st.max_idx == int(np.argmax(want_w)) or want_w[st.max_idx] >= want_w.max() * (1 - RTOL)
    assert st.ess == pytest.approx(O.ess(want_w), rel=RTOL)


def test_single_and_maximum_landmark_counts(mcl):
    """L = 1 and L = 4096 (the capacity of the shared-memory landmark table)."""
    rng = np.random.RandomState(4)
    n = 4096
    truth = np.array([[0.5], [0.25], [-0.8]])
    for Lm in (1, 4096):
        lm = rng.uniform(-20, 20, (Lm, 2))
        spread = 0.3 / np.sqrt(Lm)
        pos = (truth + rng.normal(0, 1, (3, n)) * np.array([[spread], [spread], [spread / 10]])).astype(np.float32)
        pos[:, 0] = truth[:, 0].astype(np.float32)
        z = O.world2robot(truth, lm) + rng.normal(0, 0.3, (Lm, 2))
        out = {}
        for lik in ("pair", "moment"):
            # (4096 landmarks seen 0.17 m off: the reference's fp64 product of pdfs is 0 for every particle and its
            #  weights reset to 1/NP -- test_likelihood_64_and_1024_landmarks covers that; here the log-domain weights)
            f = mcl.Filter(n_particles=n, landmarks=lm, likelihood=lik, max_landmarks=max(Lm, 8), underflow_rule=False)
            f.set_particles(pos[0], pos[1], pos[2], None)
            f.predict_weight(z[None], noise=np.zeros((3, 1, n), dtype=np.float32))
            x, y, th, _ = f.get_particles()
            out[lik] = f.weights()[0]
        got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
        ll = O.loglik_cartesian(got_px, z, lm, 0.3, 0.3)
        want_w, _ = O.update_weights_log(np.full(n, -np.log(float(n))), ll)
        _assert_weights_close(out["pair"], want_w)
        _assert_weights_close(out["moment"], want_w)
    with pytest.raises(mcl.MclError):
        mcl.Filter(n_particles=16, max_landmarks=4097)
    f = mcl.Filter(n_particles=64, landmarks=[[0.0, 0.0]])
    f.init_particles([0.0, 0.0, 0.0])
    with pytest.raises(mcl.MclError):  # the comb offset is rand() * NP_RECIP (:214): [0, 1/N)
        f.resample(ofs=1.5 / 64, force=True)


def test_anisotropic_observation_noise(mcl):
    """sigma_x != sigma_y selects the 13-moment kernel; same weights as the pair kernel and the oracle."""
    rng = np.random.RandomState(8)
    n, Lm = 8192, 12
    lm = rng.uniform(-15, 15, (Lm, 2))
    truth = np.array([[2.0], [3.0], [1.9]])
    pos = (truth + rng.normal(0, 1, (3, n)) * np.array([[0.1], [0.1], [0.02]])).astype(np.float32)
    pos[:, 0] = truth[:, 0].astype(np.float32)
    z = O.world2robot(truth, lm) + rng.normal(0, 1, (Lm, 2)) * np.array([0.2, 0.5])
    res = {}
    for lik in ("pair", "moment"):
        f = mcl.Filter(n_particles=n, landmarks=lm, likelihood=lik, sigma_obs=[0.2, 0.5], max_landmarks=16)
        f.set_particles(pos[0], pos[1], pos[2], None)
        f.predict_weight(z[None], noise=np.zeros((3, 1, n), dtype=np.float32))
        x, y, th, _ = f.get_particles()
        res[lik] = f.weights()[0]
    got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
    ll = O.loglik_cartesian(got_px, z, lm, 0.2, 0.5)
    want_w, _ = O.update_weights_log(np.full(n, -np.log(float(n))), ll)
    _assert_weights_close(res["pair"], want_w)
    _assert_weights_close(res["moment"], want_w)


def test_trial_batch_with_mixed_resampling_gates(mcl):
    """ESS-gated step on a batch where some trials resample and some do not (particle_filter.py:211);
    ragged n (not a multiple of the 128-particle stride unit)."""
    T, n, Lm = 6, 1500, 5
    rng = np.random.RandomState(21)
    lm = rng.uniform(-8, 8, (Lm, 2))
    truth = np.array([[1.0], [1.0], [0.2]])
    P = O.PFParams(100, n_particles=n, landmarks=lm)
    pos = (truth[:, None, :] + rng.normal(0, 0.05, (3, T, n))).astype(np.float32)
    lw = np.zeros((T, n), dtype=np.float32)
    for t in range(T):
        if t % 2 == 0:  # degenerate weights: ESS ~ 3 << n / 100
            lw[t] = -40.0
            lw[t, rng.randint(0, n, 3)] = 0.0
        else:  # healthy weights: ESS ~ n
            lw[t] = rng.normal(0, 0.05, n)
    noise = (rng.normal(0, 1, (3, T, n)) * np.sqrt(np.diag(P.Q))[:, None, None]).astype(np.float32)
    z = np.stack([O.world2robot(truth, lm) + rng.normal(0, 0.3, (Lm, 2)) for _ in range(T)])
    ofs = rng.rand(T) / n
    f = mcl.Filter(n_particles=n, n_trials=T, landmarks=lm)
    f.set_particles(pos[0], pos[1], pos[2], lw)
    ess0 = [s.ess for s in f.stats()]
    q = f.qweights()
    f.step(z, ofs=ofs, noise=noise)  # gated, not forced
    st = f.stats()
    anc = f.ancestors()
    x, y, th, _ = f.get_particles()
    w = f.weights()
    for t in range(T):
        fired = ess0[t] < n / 100.0
        assert fired == (t % 2 == 0) and st[t].resampled == int(fired)
        src = pos[:, t].astype(np.float64)
        if fired:
            a = anc[t].astype(np.int64)
            np.testing.assert_array_equal(a, O.resample_indices(O.dequantize_weights(q[t]), ofs[t]))
            src = src[:, a]
            prev = np.full(n, -np.log(float(n)))
        else:
            prev = np.log(_weights_from_logw32(lw[t]))
        got = np.stack([x[t], y[t], th[t]]).astype(np.float64)
        np.testing.assert_allclose(got, O.predict(src, noise[:, t].astype(np.float64), P), rtol=RTOL, atol=1e-6)
        ll = O.loglik_cartesian(got, z[t], lm, 0.3, 0.3)
        want_w, _ = O.update_weights_log(prev, ll)
        _assert_weights_close(w[t], want_w)
        assert st[t].ess == pytest.approx(O.ess(want_w), rel=RTOL)


def test_yaw_wrap_at_the_seam_matches_limit_angle(mcl):
    """Particles whose yaw + w dt lands on both sides of pi: the wrap decision is the fp64 one of mylib/limit.py."""
    n = 4096
    P = O.PFParams(100, n_particles=n)
    rng = np.random.RandomState(12)
    wdt = P.YAW_RATE_rps * P.DT_s
    th = (np.pi - wdt + rng.uniform(-3e-6, 3e-6, n)).astype(np.float32)
    th[::2] *= -1  # and the mirrored seam; also exercises the sign rule for negative angles
    th[1::4] = (-np.pi - wdt + rng.uniform(-3e-6, 3e-6, n // 4)).astype(np.float32)
    pos = np.stack([rng.normal(0, 1, n), rng.normal(0, 1, n), th]).astype(np.float32)
    f = mcl.Filter(n_particles=n, landmarks=P.LM)
    f.set_particles(pos[0], pos[1], pos[2], None)
    f.predict_weight(np.zeros((1, 5, 2)), noise=np.zeros((3, 1, n), dtype=np.float32))
    got = f.get_particles()[2][0].astype(np.float64)
    want = O.predict(pos.astype(np.float64), np.zeros((3, n)), P)[2]
    assert np.all(np.abs(got - want) < 1e-6)  # same side of the seam everywhere
    assert (want > 3.0).any() and (want < -3.0).any()


def _fan_case(rng, n, Lm, spread):
    lm = rng.uniform(-14, 14, (Lm, 2))
    truth = np.array([[1.5], [-2.0], [0.6]])
    pos = (truth + rng.normal(0, 1, (3, n)) * np.array([[spread], [spread], [0.25 * spread]])).astype(np.float32)
    return lm, truth, pos


@pytest.mark.gpu
@pytest.mark.parametrize("Lm,spread", [(24, 0.3), (200, 2.0)])
def test_rb_fan_likelihood_matches_oracle(mcl, Lm, spread):
    """the fan-sensor likelihood (fp32 per pair, detected landmarks only, visibility gate) against
    the fp64 oracle on the same stored poses.  The gate is a step function, so particles that sit within 1e-4 of a
    gate boundary for some landmark are excluded (fp32 may legitimately take the other side)."""
    rng = np.random.RandomState(Lm)
    n = 4096
    lm, truth, pos = _fan_case(rng, n, Lm, spread)
    rng_obs = np.random.RandomState(Lm + 1)
    ids, noisy, _ = O.scan_sensor(truth, lm, 15.0, np.deg2rad(80.0), 0.05, np.deg2rad(2.0), np.deg2rad(2.0), rng=rng_obs)
    assert 2 <= ids.size < Lm
    z = O.rb_fan_table(ids, noisy, Lm)
    f = mcl.Filter(n_particles=n, landmarks=lm, likelihood="rb_fan", max_landmarks=Lm, q_std=[1e-9, 1e-9, 1e-9])
    f.set_particles(pos[0], pos[1], pos[2], None)
    f.predict_weight(z[None], noise=np.zeros((3, 1, n), np.float32))
    x, y, th, lw = f.get_particles()
    got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
    want = O.loglik_rb_fan(got_px, z, lm, 15.0, np.deg2rad(80.0), 0.05, np.deg2rad(2.0), 18.0)
    # robust particles: every detected landmark clearly inside or clearly outside the particle's fan
    p = O.world2robot_batch(got_px, lm[ids])
    d = np.sqrt(p[..., 0] ** 2 + p[..., 1] ** 2)
    m_range = np.abs(d - 15.0)
    m_ang = np.abs(p[..., 1] - np.abs(p[..., 0]) * np.tan(np.pi / 2 - np.deg2rad(80.0)))
    robust = (np.minimum(m_range, m_ang) > 1e-3).all(axis=1)
    assert robust.mean() > 0.95
    got = lw[0].astype(np.float64)
    # stored log-weights are relative to one per-step constant: compare differences to the best particle
    k = int(np.argmax(np.where(robust, want, -np.inf)))
    dg, dw = got - got[k], want - want[k]
    tol = 2e-3 + 2e-6 * np.abs(dw)  # fp32 residuals, MUFU sqrt, atan2f; chi^2 up to ~1e4 per particle
    bad = robust & (np.abs(dg - dw) > tol)
    assert not bad.any(), (int(bad.sum()), float(np.abs(dg - dw)[robust].max()))
    st = f.stats()[0]
    assert st.max_idx == int(np.argmax(got))


@pytest.mark.gpu
def test_rb_fan_gate_and_empty_scan(mcl):
    """Nothing detected -> the step leaves the weights uniform; a particle that looks away from every detected
    landmark pays exactly fan_miss_chi2 / 2 per landmark; a full step with resampling runs in this mode."""
    rng = np.random.RandomState(9)
    n, Lm = 2048, 32
    lm, truth, pos = _fan_case(rng, n, Lm, 0.05)
    f = mcl.Filter(n_particles=n, landmarks=lm, likelihood="rb_fan", max_landmarks=Lm, fan_miss_chi2=8.0,
                   q_std=[1e-9, 1e-9, 1e-9], vel=0.0, yaw_rate=0.0)
    f.set_particles(pos[0], pos[1], pos[2], None)
    f.predict_weight(np.full((1, Lm, 2), np.nan), noise=np.zeros((3, 1, n), np.float32))
    np.testing.assert_allclose(f.weights()[0], 1.0 / n, rtol=1e-6)
    # exact observation from the truth; half of the particles turned around
    ids, _, clean = O.scan_sensor(truth, lm, 15.0, np.deg2rad(80.0), 0.05, 0.03, 0.03, rng=np.random.RandomState(1))
    z = O.rb_fan_table(ids, clean, Lm)
    pos2 = np.repeat(truth.astype(np.float32), n, axis=1)
    pos2[2, n // 2:] += np.float32(np.pi)
    f.set_particles(pos2[0], pos2[1], pos2[2], None)
    f.predict_weight(z[None], noise=np.zeros((3, 1, n), np.float32))
    _, _, _, lw = f.get_particles()
    front, back = lw[0][: n // 2].astype(np.float64), lw[0][n // 2:].astype(np.float64)
    assert np.ptp(front) < 1e-3 and np.ptp(back) < 1e-6
    assert front.mean() - back.mean() == pytest.approx(0.5 * 8.0 * ids.size, abs=2e-3)
    # a normal step (resampling forced) in this mode
    f2 = mcl.Filter(n_particles=n, landmarks=lm, likelihood="rb_fan", max_landmarks=Lm)
    f2.init_particles(truth[:, 0])
    sens = mcl.ScanSensor(15.0, np.deg2rad(80.0), lm)
    sens.setNoiseParam(5, 2, 2)
    np.random.seed(3)
    for _ in range(5):
        zt, _ = sens.scan(truth)
        f2.step(zt[None], force_resample=True)
    st = f2.stats()[0]
    assert np.isfinite(st.ess) and 1.0 <= st.ess <= n
    assert np.hypot(st.est[0] - truth[0, 0], st.est[1] - truth[1, 0]) < 3.0


class _AxStub:
    def __init__(self, log):
        self.log = log

    def __getattr__(self, name):
        def call(*a, **k):
            self.log.append((name, a, k))
        return call


class _PltStub(_AxStub):
    def subplot2grid(self, *a, **k):
        return _AxStub(self.log)


@pytest.mark.gpu
def test_animation_caller_consumes_the_dropin(mcl):
    """the reference's animation loop (particle_filter.py:248-328), restated in
    examples/animate_pf.py, runs frame by frame against the GPU class: the 7-tuple, px read lazily and
    sub-sampled for drawing, the covariance ellipse.  (matplotlib is not needed: a recording stub stands in.)"""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("animate_pf", os.path.join(root, "examples", "animate_pf.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    np.random.seed(0)
    pf = mcl.ParticleFilter(100, n_particles=50000, noise="philox", likelihood="moment", lazy_px=True)
    log = []
    anim = mod.Animator(pf, 100, _PltStub(log), max_scatter=4000)
    for i in range(3):
        out = anim.frame(i)
    assert out["n_drawn"] == 4000 and out["time_s"] == pytest.approx(0.3)
    assert 0 <= out["w_idx"] < 50000 and 0.0 < out["w_val"] <= 1.0
    assert np.all(np.linalg.eigvalsh(out["cov"][0:2, 0:2]) > 0)
    names = [c[0] for c in log]
    assert names.count("scatter") >= 3 * 8 and "quiver" in names and "annotate" in names
    e = mod.ellipse_points(np.diag([4.0, 1.0]), chi2=1.0)
    assert np.abs(e[0]).max() == pytest.approx(2.0, rel=1e-3) and np.abs(e[1]).max() == pytest.approx(1.0, rel=1e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("trials", [1, 3])
def test_checkpoint_restores_a_bit_identical_filter(mcl, tmp_path, trials):
    """Filter.save / Filter.load -- the restored filter continues exactly like the saved one
    (same Philox counters, same weight map, same reference pose), gated resampling included."""
    n, Lm = 20000, 7
    rng = np.random.RandomState(4)
    lm = rng.uniform(-10, 10, (Lm, 2))
    tr = __import__("bench").Truth(lm, seed=2, T=trials)
    f = mcl.Filter(n_particles=n, n_trials=trials, landmarks=lm, likelihood="moment", seed=99)
    f.init_particles(tr.pose())
    zs = []
    for _ in range(12):
        tr.advance()
        zs.append(tr.observe())
    for k in range(6):
        f.step(zs[k])  # ESS-gated: some of these steps resample, some do not
    path = str(tmp_path / "ckpt.npz")
    f.save(path)
    g = mcl.Filter.load(path)
    assert g.n == n and g.T == trials and g.n_landmarks == Lm
    for k in range(6, 12):
        f.step(zs[k])
        g.step(zs[k])
    for a, b in zip(f.get_particles(), g.get_particles()):
        np.testing.assert_array_equal(a, b)
    sa, sb = f.stats(), g.stats()
    for t in range(trials):
        assert sa[t].max_idx == sb[t].max_idx and sa[t].ess == sb[t].ess and sa[t].step == sb[t].step == 12
        assert sa[t].resampled == sb[t].resampled


# ---------------------------------------------------------------------------------------------
# device-driven sharded filter (peer memory): G shards on this one GPU, no host in the loop
# ---------------------------------------------------------------------------------------------
def _peer_case(mcl, G, kind, n, steps=1):
    Lm = 6
    m = n // G
    rng = np.random.RandomState(100 + G + len(kind))
    lm = rng.uniform(-10, 10, (Lm, 2))
    truth = np.array([[2.0], [1.0], [0.4]])
    pos = (rng.normal(0, 1, (3, n)) * np.array([[0.2], [0.2], [0.05]]) + truth).astype(np.float32)
    if kind == "spread":
        lw = (-0.5 * ((np.arange(n) - 0.3 * n) / (0.05 * n)) ** 2 + rng.normal(0, 1.0, n)).astype(np.float32)
    elif kind == "left_heavy":  # nearly all the weight in the first shard: it fills slots of every other rank
        lw = (np.where(np.arange(n) < m // 3, 0.0, -30.0) + rng.normal(0, 1.0, n)).astype(np.float32)
    else:  # right_heavy: the last shard fills everybody's slots (suffix path of every lower rank)
        lw = (np.where(np.arange(n) >= n - m // 3, 0.0, -30.0) + rng.normal(0, 1.0, n)).astype(np.float32)
    sf = mcl.LocalShardedFilter(n, lm, G, likelihood="pair")
    sf.set_particles(pos[0], pos[1], pos[2], lw)
    f1 = mcl.Filter(n_particles=n, landmarks=lm, likelihood="pair")
    f1.set_particles(pos[0], pos[1], pos[2], lw, skip_stats=True)
    for k in range(steps):
        # the unsharded filter takes the MERGED statistics of the shards, so both quantise with the same weight map
        # (its own merge groups the fp32 partial sums differently: equal to ~1e-9, not bitwise)
        g_stats = sf.shards[0].stats()[0]
        for f in sf.shards[1:]:
            o = f.stats()[0]
            assert (o.max_logw, o.sum_w, o.sum_w2, o.ess, o.max_idx) == (g_stats.max_logw, g_stats.sum_w, g_stats.sum_w2, g_stats.ess, g_stats.max_idx)
        f1.restore_stats(g_stats, k)
        q = np.concatenate([f.qweights()[0] for f in sf.shards])
        assert np.array_equal(q, f1.qweights()[0])
        ofs = rng.rand() / n
        noise = (rng.normal(0, 1, (3, 1, n)) * np.array([0.03, 0.03, 0.035])[:, None, None]).astype(np.float32)
        z = O.world2robot(truth, lm) + rng.normal(0, 0.3, (Lm, 2))
        force = (k % 3 != 1)
        sf.step(z[None], ofs=ofs, noise=noise, force_resample=force, no_resample=not force)
        f1.step(z[None], ofs=ofs, noise=noise, force_resample=force, no_resample=not force)
        for f in sf.shards:
            pl = f.shard_plan()
            assert pl.error == 0 and pl.resampled == int(force)
        x1, y1, t1, l1 = f1.get_particles()
        xs, ys, ts, ls = sf.particles()
        if force:
            anc1 = f1.ancestors()[0].astype(np.int64)
            want = O.resample_indices_literal(O.dequantize_weights(q), ofs)
            assert np.array_equal(anc1, want)
            # the plan every shard derived on the device against the host twin of round 1
            sh = importlib.import_module("slam-robot_simu_b200.sharded")
            totals = [int(q[g * m:(g + 1) * m].sum()) for g in range(G)]
            ranges = sh.shard_ranges(totals, n, ofs)
            for g, f in enumerate(sf.shards):
                pl = f.shard_plan()
                assert [int(v) for v in pl.totals[:G]] == totals
                assert (pl.first_slot, pl.n_slots) == ranges[g]
                assert pl.k_offset == sum(totals[:g])
        np.testing.assert_array_equal(xs, x1[0])
        np.testing.assert_array_equal(ys, y1[0])
        np.testing.assert_array_equal(ts, t1[0])
        np.testing.assert_array_equal(ls, l1[0])
        s1 = f1.stats()[0]
        sg = sf.stats()
        assert sg["max_idx"] == s1.max_idx and sg["ess"] == pytest.approx(s1.ess, rel=1e-6)
        assert sg["log_norm"] == pytest.approx(s1.log_norm, abs=1e-7)


import importlib  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("G", [2, 4, 8])
@pytest.mark.parametrize("kind", ["spread", "left_heavy", "right_heavy"])
def test_peer_sharded_step_matches_unsharded(mcl, G, kind):
    """SURVEY 8(e), device-driven: statistics, totals and off-rank poses travel through peer memory, gate / k_offset /
    slot ranges are derived on the device -- the particles equal the unsharded filter's bit for bit."""
    _peer_case(mcl, G, kind, 1 << 16, steps=4)


@pytest.mark.gpu
def test_peer_sharded_step_with_segmented_tile_prefixes(mcl):
    _peer_case(mcl, 2, "spread", 1 << 23, steps=2)


@pytest.mark.gpu
def test_peer_sharded_philox_run_with_device_gate(mcl):
    """A free-running sharded filter (Philox noise, ESS gate taken on the device, no host decision anywhere) against
    the unsharded filter fed the shards' merged statistics before every step: identical particles, identical gates."""
    n, G, Lm = 1 << 17, 4, 16
    bench = importlib.import_module("bench")
    lm = bench.landmarks_for(Lm)
    truth = bench.Truth(lm, seed=3)
    sf = mcl.LocalShardedFilter(n, lm, G, likelihood="moment", seed=11)
    sf.init_particles(truth.pose()[0])
    ref = mcl.Filter(n_particles=n, landmarks=lm, likelihood="moment", seed=11, max_landmarks=Lm)
    ref.init_particles(truth.pose()[0])
    fired = 0
    for k in range(14):
        truth.advance()
        z = truth.observe()
        if k > 0:
            ref.restore_stats(sf.shards[0].stats()[0], k)
        sf.step(z)
        ref.step(z)
        x1, y1, t1, l1 = ref.get_particles()
        xs, ys, ts, ls = sf.particles()
        np.testing.assert_array_equal(xs, x1[0])
        np.testing.assert_array_equal(ts, t1[0])
        np.testing.assert_array_equal(ls, l1[0])
        r1 = ref.stats()[0].resampled
        assert all(f.shard_plan().resampled == r1 for f in sf.shards)
        fired += r1
    assert 0 < fired < 14  # the gate both fired and held


# ---------------------------------------------------------------------------------------------
# the reference's degenerate-weight rule (particle_filter.py:234-237)
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("likelihood", ["pair", "moment", "pair32"])
def test_underflow_resets_the_weights_like_the_reference(mcl, golden_dir, likelihood):
    """Golden run of the unmodified reference with the robot teleported 40 m away for steps 4 and 5: every pw * bn
    underflows, the reference's weights become NaN and are reset to 1/NP (max_idx 0, max_val 1/NP).  The engine's
    log-domain weights never underflow, so it detects the event (largest log(pw * bn) < log 2^-1075) and mirrors the
    reset.  Stage by stage on the reference's arrays, then free-running through 3 -> 7 with the state carried over."""
    g = np.load(os.path.join(golden_dir, "pf_underflow.npz"))
    P = O.PFParams(100)
    n = P.NP

    def check(f, k, got_px, w_prev):
        st = f.stats()[0]
        got_w = f.weights()[0]
        if k in (4, 5):
            assert st.underflow == 1 and st.max_idx == 0 == int(g[f"s{k}_max_idx"])
            np.testing.assert_allclose(got_w, 1.0 / n, rtol=1e-6)
            assert np.all(g[f"s{k}_pw"] == 1.0 / n)
            assert st.max_weight == pytest.approx(1.0 / n, rel=1e-12) and st.ess == pytest.approx(n)
            np.testing.assert_array_equal(np.array(st.est[:]), got_px[:, 0])
        else:
            assert st.underflow == 0
            ll = O.loglik_cartesian(got_px, g[f"s{k}_z_l"], P.LM, 0.3, 0.3)
            with np.errstate(divide="ignore"):
                want_w, _ = O.update_weights_log(np.log(w_prev), ll)
            _assert_weights_close(got_w, want_w, rtol_tail=3e-5 if likelihood == "pair32" else None)
            assert st.max_idx == int(np.argmax(want_w))
        return got_w

    # stage by stage: every step starts from the reference's own arrays
    for k in range(int(g["n_steps"])):
        px0 = g[f"s{k}_px_before"].astype(np.float32)
        lw0 = _logw32(g[f"s{k}_pw_before"])
        f = mcl.Filter(n_particles=n, landmarks=P.LM, likelihood=likelihood)
        f.set_particles(px0[0], px0[1], px0[2], lw0)
        f.predict_weight(g[f"s{k}_z_l"][None], noise=g[f"s{k}_v"].astype(np.float32)[:, None, :])
        x, y, th, _ = f.get_particles()
        got_w = check(f, k, np.stack([x[0], y[0], th[0]]).astype(np.float64), _weights_from_logw32(lw0))
        if k not in (4, 5):
            _assert_weights_close(got_w, g[f"s{k}_pw"], rtol=2e-3)  # the reference's fp64 poses: fp32 rounding of the inputs

    # free-running: the uniform reset must carry into the next step's weights (pw_before = 1/NP) and into the ESS gate
    px0 = g["s3_px_before"].astype(np.float32)
    f = mcl.Filter(n_particles=n, landmarks=P.LM, likelihood=likelihood)
    f.set_particles(px0[0], px0[1], px0[2], _logw32(g["s3_pw_before"]))
    w_prev = _weights_from_logw32(_logw32(g["s3_pw_before"]))
    for k in range(3, 8):
        f.step(g[f"s{k}_z_l"][None], ofs=0.5 / n, noise=g[f"s{k}_v"].astype(np.float32)[:, None, :])  # device ESS gate
        st = f.stats()[0]
        assert st.resampled == 0  # as in the reference run: ESS = NP after a reset
        x, y, th, _ = f.get_particles()
        w_prev = check(f, k, np.stack([x[0], y[0], th[0]]).astype(np.float64), w_prev)


@pytest.mark.gpu
def test_underflow_rule_in_the_sharded_filter(mcl):
    """The same rule on the GLOBAL weights of a device-driven sharded filter: every shard reports the reset, the
    arg-max is global particle 0, and the forced resampling that follows treats the weights as all-equal -- particles
    stay bit-identical to the unsharded filter's."""
    n, G = 1 << 14, 4
    P = O.PFParams(100)
    bench = importlib.import_module("bench")
    truth = bench.Truth(P.LM, seed=9)
    far = bench.Truth(P.LM, seed=9)
    far.x = far.x + np.array([[40.0, 40.0, 0.0]])
    sf = mcl.LocalShardedFilter(n, P.LM, G, likelihood="pair", seed=5)
    sf.init_particles(truth.pose()[0])
    ref = mcl.Filter(n_particles=n, landmarks=P.LM, likelihood="pair", seed=5)
    ref.init_particles(truth.pose()[0])
    for k, (src, force) in enumerate([(truth, False), (far, False), (far, True), (truth, True), (truth, False)]):
        truth.advance()
        far.advance()
        z = src.observe()
        if k > 0:
            ref.restore_stats(sf.shards[0].stats()[0], k)
        g = sf.step(z, force_resample=force, no_resample=not force)
        ref.step(z, force_resample=force, no_resample=not force)
        s1 = ref.stats()[0]
        assert g["underflow"] == s1.underflow == (1 if src is far else 0)
        assert all(f.stats()[0].underflow == s1.underflow for f in sf.shards)
        assert g["max_idx"] == s1.max_idx and g["ess"] == pytest.approx(s1.ess, rel=1e-6)
        if src is far:
            assert g["max_idx"] == 0 and g["ess"] == pytest.approx(n)
            np.testing.assert_allclose(np.concatenate([f.weights()[0] for f in sf.shards]), 1.0 / n, rtol=1e-6)
        x1, y1, t1, l1 = ref.get_particles()
        xs, ys, ts, ls = sf.particles()
        np.testing.assert_array_equal(xs, x1[0])
        np.testing.assert_array_equal(ts, t1[0])
        np.testing.assert_array_equal(ls, l1[0])


@pytest.mark.gpu
def test_sharded_checkpoint_restores_a_bit_identical_filter(mcl, tmp_path):
    """SURVEY 8(f)-4 for the sharded filter: one file per shard; the restored filter (fresh handles, fresh peer
    mappings) continues exactly like the saved one -- device gate, Philox counters and weight map included."""
    n, G, Lm = 1 << 16, 4, 16
    bench = importlib.import_module("bench")
    lm = bench.landmarks_for(Lm)
    truth = bench.Truth(lm, seed=21)
    a = mcl.LocalShardedFilter(n, lm, G, likelihood="moment", seed=77)
    a.init_particles(truth.pose()[0])
    zs = []
    for _ in range(10):
        truth.advance()
        zs.append(truth.observe())
    for k in range(4):
        a.step(zs[k])
    a.save(str(tmp_path / "ck"))
    want = []
    for k in range(4, 10):
        a.step(zs[k])
        want.append((a.particles(), a.stats()))
    b = mcl.LocalShardedFilter(n, lm, G, likelihood="moment", seed=77)
    b.restore(str(tmp_path / "ck"))
    fired = 0
    for k in range(4, 10):
        b.step(zs[k])
        (xa, ya, ta, la), sa = want[k - 4]
        xb, yb, tb, lb = b.particles()
        np.testing.assert_array_equal(xb, xa)
        np.testing.assert_array_equal(tb, ta)
        np.testing.assert_array_equal(lb, la)
        sb = b.stats()
        assert (sb["ess"], sb["max_idx"], sb["resampled"], sb["step"]) == (sa["ess"], sa["max_idx"], sa["resampled"], sa["step"])
        fired += sb["resampled"]
    assert fired > 0  # the restored gate was exercised


@pytest.mark.gpu
def test_trajectory_log_round_trip(mcl, tmp_path):
    """the on-disk trajectory / particle dump holds what main_pf() returned, step by step."""
    np.random.seed(0)
    log = mcl.TrajectoryLog(particles_every=2, max_particles=256)
    pf = mcl.ParticleFilter(100, trajectory=log)
    outs = [pf.main_pf() for _ in range(5)]
    log.save(str(tmp_path / "run.npz"))
    d = mcl.TrajectoryLog.load(str(tmp_path / "run.npz"))
    assert len(log) == 5 and list(d["step"]) == [1, 2, 3, 4, 5] and sorted(d["clouds"]) == [0, 2, 4]
    for k, (LM, x_true, x_est, px, Q, idx, val) in enumerate(outs):
        np.testing.assert_array_equal(d["x_true"][k], x_true[:, 0])
        np.testing.assert_array_equal(d["x_est"][k], x_est[:, 0])
        assert d["max_idx"][k] == idx and d["max_weight"][k] == val
    assert d["clouds"][4].shape == (3, 334) and np.all(np.linalg.eigvalsh(d["cov"][4][0:2, 0:2]) > 0)


@pytest.mark.gpu
def test_slots_past_the_final_cumulative_weight_are_bounded(mcl):
    """The one intentional divergence: the integer weights sum to 2^52 (1 +- 6e-8) because the
    scale 2^52 / sum is an fp32 number, and when they fall short the comb teeth past the total go to the LAST particle
    (the reference's walk would raise IndexError there).  At 2^27 particles that is at most N 2^-23 = 16 slots, and
    exactly the number the shortfall of the integer total predicts."""
    import torch

    n = 1 << 27
    g = torch.Generator(device="cuda").manual_seed(5)
    lw = torch.randn(n, generator=g, device="cuda", dtype=torch.float32) * 1.5
    zero = torch.zeros(n, device="cuda", dtype=torch.float32)
    f = mcl.Filter(n_particles=n, landmarks=[[0.0, 0.0]])
    f.set_particles(zero, zero, zero, lw)
    for trial in range(6):
        if trial:  # another spread -> another sum, another rounding of the fp32 scale (the total lands on either side of 2^52)
            f.set_particles(zero, zero, zero, lw * (1.0 + 0.21 * trial))
        q = f.qweights()[0]
        total = int(q.sum(dtype=np.uint64))
        ofs = (0.11 + 0.13 * trial) / n  # rand() * NP_RECIP: in [0, 1/N)
        f.resample(ofs=ofs, force=True)
        anc = f.ancestors()[0]
        assert np.all(np.diff(anc.astype(np.int64)) >= 0)
        # teeth u_i = i / N + ofs beyond total * 2^-52 have no owner in the reference: count them from the integers
        short = (1 << 52) - total
        tail = int((anc == n - 1).sum())
        fair = q[-1] * n / float(1 << 52)  # the last particle's own share of the comb
        dumped = tail - fair
        assert dumped <= n * 2.0 ** -23 + 2, (trial, tail, fair, short)
        # tooth i is unowned when i / N + ofs > total 2^-52: about N short 2^-52 - N ofs of them (ofs N < 1)
        assert abs(dumped - max(0.0, short * n / float(1 << 52) - ofs * n)) <= 2.0, (trial, dumped, short)
        print(f"trial {trial}: total - 2^52 = {-short}, slots of the last particle {tail} (its share {fair:.2f}), dumped {dumped:.2f}")
