import sys, numpy as np
sys.path.insert(0,'/root/repo')
import slam_robot_simu_b200 as mcl
from oracle import pf_oracle as O
for likelihood in ("pair","moment","pair32"):
    n, Lm = 1 << 20, 16
    rng = np.random.RandomState(1234)
    lm = np.random.default_rng(1234).uniform(-20, 20, (Lm, 2))
    P = O.PFParams(100, n_particles=n, landmarks=lm)
    truth = np.array([[10.0], [0.0], [np.pi / 2]])
    pos = (truth + rng.normal(0, 1, (3, n)) * np.array([[0.08], [0.08], [0.03]])).astype(np.float32)
    z = O.world2robot(truth, lm) + rng.normal(0, 0.3, (Lm, 2))
    ll0 = O.loglik_cartesian(pos.astype(np.float64), z, lm, 0.3, 0.3)
    lw = (ll0 - ll0.max()).astype(np.float32)
    noise = (rng.normal(0, 1, (3, n)) * np.sqrt(np.diag(P.Q))[:, None]).astype(np.float32)
    truth1 = O.f_motion(truth, P)
    z1 = O.world2robot(truth1, lm) + rng.normal(0, 0.3, (Lm, 2))
    ofs = rng.rand() / n
    f = mcl.Filter(n_particles=n, landmarks=lm, likelihood=likelihood)
    f.set_particles(pos[0], pos[1], pos[2], lw)
    f.step(z1[None], ofs=ofs, noise=noise[:, None, :], force_resample=True)
    x, y, th, lwg = f.get_particles()
    got_px = np.stack([x[0], y[0], th[0]]).astype(np.float64)
    ll = O.loglik_cartesian(got_px, z1, lm, 0.3, 0.3)
    want_w, logz = O.update_weights_log(np.full(n, -np.log(float(n))), ll)
    got_w = f.weights()[0]
    st = f.stats()[0]
    for lo,hi in ((1e-4,1),(1e-8,1e-4),(1e-12,1e-8),(1e-20,1e-12)):
        m=(want_w>want_w.max()*lo)&(want_w<=want_w.max()*hi)
        if m.any(): print(likelihood, f"weights in ({lo:g},{hi:g}]·max: n={m.sum()} max rel err {np.abs(got_w[m]/want_w[m]-1).max():.2e}")
    # stored log-weight error (relative to the best particle)
    d=(lwg[0].astype(np.float64)-lwg[0].max())-(ll-ll.max()); m=want_w>want_w.max()*1e-12
    print(likelihood,"stored lw abs err (down to 1e-12):",np.abs(d[m]).max())
    mean, cov = O.weighted_moments(got_px, want_w)
    c=np.array(st.cov[:]).reshape(3,3)
    print(likelihood,"mean rel/abs err",np.abs(np.array(st.mean[:])-mean).max(),"cov rel err",np.abs(c/cov-1).max(), "ess rel", abs(st.ess/O.ess(want_w)-1))
    f.pose_moments(); st2=f.stats()[0]; c2=np.array(st2.cov[:]).reshape(3,3)
    print(likelihood,"two-pass cov rel err",np.abs(c2/cov-1).max())
