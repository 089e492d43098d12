"""CPU tests (no GPU): the C ABI loads and exports everything include/mcl.h declares, fails loudly
without a device, the host-side mirror of the reference interface, and the sharded filter's host logic
(including a world_size-2 gloo run)."""
import ctypes as C
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import pf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    sys.path.insert(0, ROOT)
    bld = importlib.util.spec_from_file_location("mcl_build", os.path.join(ROOT, "slam-robot_simu_b200", "build.py"))
    mod = importlib.util.module_from_spec(bld)
    bld.loader.exec_module(mod)
    mod.build()  # no-op when libmcl_b200.so is up to date; nvcc cross-compiles without a GPU
    return importlib.import_module("slam-robot_simu_b200")


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "mcl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.load()
    names = _declared_functions()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/mcl.h but not exported by libmcl_b200.so"
        assert n in pkg._lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(pkg._lib.PROTOTYPES) == set(names)
    assert lib.mcl_abi_version() == pkg._lib.MCL_ABI_VERSION
    assert lib.mcl_status_string(0) == b"ok" and lib.mcl_status_string(-5) == b"no sm_100 CUDA device"


def test_struct_layout_matches_the_header(pkg, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "mcl.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(mcl_config),sizeof(mcl_stats),offsetof(mcl_stats,raw),offsetof(mcl_config,seed),'
                   'offsetof(mcl_config,fan_miss_chi2));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    L = pkg._lib
    assert got == [C.sizeof(L.mcl_config), C.sizeof(L.mcl_stats), L.mcl_stats.raw.offset, L.mcl_config.seed.offset,
                   L.mcl_config.fan_miss_chi2.offset]


def test_config_defaults_are_the_reference_literals(pkg):
    cfg = pkg.default_config(100.0)
    P = O.PFParams(100)
    assert cfg.struct_size == C.sizeof(pkg.mcl_config)
    assert cfg.n_particles == P.NP and cfg.dt == P.DT_s
    assert cfg.vel == pytest.approx(P.VEL_mps, rel=1e-15) and cfg.yaw_rate == pytest.approx(P.YAW_RATE_rps, rel=1e-15)
    np.testing.assert_allclose(np.array(cfg.q_std[:]) ** 2, np.diag(P.Q), rtol=1e-15)
    np.testing.assert_allclose(np.array(cfg.sigma_obs[:]) ** 2, np.diag(P.R), rtol=1e-15)
    assert cfg.ess_threshold == P.ESS_TH


def test_no_device_fails_loudly_and_never_falls_back(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = pkg._lib.load()
    cfg = pkg.default_config(100.0)
    need = C.c_size_t(0)
    assert lib.mcl_workspace_bytes(C.byref(cfg), C.byref(need)) == 0 and need.value > 1000 * 36
    h = C.c_void_p()
    assert lib.mcl_create(C.byref(cfg), None, 0, C.byref(h)) == pkg._lib.MCL_ERR_NODEVICE
    assert not h.value and b"no CPU fallback" in lib.mcl_last_error(None)
    with pytest.raises(pkg.MclError):
        pkg.Filter(n_particles=1000)
    with pytest.raises(pkg.MclError):
        pkg.ParticleFilter(100)
    # argument validation happens before any device work
    cfg.struct_size = 8
    assert lib.mcl_workspace_bytes(C.byref(cfg), C.byref(need)) == pkg._lib.MCL_ERR_ARG
    cfg = pkg.default_config(100.0)
    cfg.n_trials = 2
    cfg.n_global = 4000  # sharding and trial batching are exclusive
    assert lib.mcl_workspace_bytes(C.byref(cfg), C.byref(need)) == pkg._lib.MCL_ERR_ARG
    for fn in ("mcl_destroy", ):
        assert getattr(lib, fn)(None) == 0
    assert lib.mcl_step(None, None, None, None, 0, None) == pkg._lib.MCL_ERR_ARG


def test_workspace_scales_with_the_configs(pkg):
    lib = pkg._lib.load()
    for n, T, per_particle in ((1 << 20, 1, 36), (16384, 4096, 36), (125_000_000, 1, 36)):
        cfg = pkg.default_config(100.0)
        cfg.n_particles, cfg.n_trials, cfg.n_global = n, T, n
        need = C.c_size_t(0)
        assert lib.mcl_workspace_bytes(C.byref(cfg), C.byref(need)) == 0
        assert per_particle * n * T <= need.value <= per_particle * n * T * 1.02 + (8 << 20)


def test_host_mirror_of_the_reference_helpers(pkg, golden_dir):
    pf = importlib.import_module("slam-robot_simu_b200.particle_filter")
    h = np.load(os.path.join(golden_dir, "helpers.npz"))
    got = np.array([float(pf.limit_angle(a)) for a in h["lim_in"]])
    assert np.array_equal(got, h["lim_out"])
    for p, want in zip(h["w2r_pose"], h["w2r_out"]):
        np.testing.assert_allclose(pf.world2robot(p, h["w2r_lm"]), want, rtol=0, atol=1e-12)


def test_scan_sensor_mirror_against_reference_golden(pkg, golden_dir):
    """The product's host-side ScanSensor (scan_sensor.py) against the unmodified reference class
    (graph_based_slam.py:73-167; tests/golden/make_golden_scan.py): same detections, same draws."""
    g = np.load(os.path.join(golden_dir, "scan_sensor.npz"))
    for c in range(int(g["n_cases"])):
        lm = g[f"c{c}_lm"]
        sens = pkg.ScanSensor(float(g["scan_range"]), float(g["scan_angle"]), lm)
        sens.setNoiseParam(5, 2, 2)
        np.random.seed(int(g[f"c{c}_seed"]))
        z, ids = sens.scan(g[f"c{c}_pose"])
        np.testing.assert_array_equal(ids, g[f"c{c}_ids"])
        np.testing.assert_array_equal(z[ids], g[f"c{c}_noisy"][:, :2])
        assert np.isnan(z[:, 0]).sum() == lm.shape[0] - ids.size
        z0, _ = sens.scan(g[f"c{c}_pose"], noise=False)
        np.testing.assert_array_equal(z0[ids], g[f"c{c}_clean"][:, :2])
    cfg = pkg.default_config(100.0)
    assert cfg.fan_range == 15.0 and cfg.fan_half_angle == pytest.approx(np.deg2rad(80.0), rel=1e-15)  # :901-902


def test_philox_known_answers(pkg):
    sh = importlib.import_module("slam-robot_simu_b200.sharded")
    # Random123 known-answer vectors for philox4x32-10
    assert sh.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert sh.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert sh.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = [sh.philox_ofs(42, k, 1000) for k in range(2000)]
    assert 0.0 <= min(u) and max(u) < 1e-3 and abs(np.mean(u) * 1000 - 0.5) < 0.03


def test_plan_exchange_properties(pkg):
    sh = importlib.import_module("slam-robot_simu_b200.sharded")
    rng = np.random.RandomState(0)
    for world in (2, 4, 8):
        n = 1024 * world
        m = n // world
        for _ in range(50):
            cuts = np.sort(rng.randint(0, n + 1, world - 1))
            bounds = np.concatenate([[0], cuts, [n]])
            ranges = [(int(a), int(b - a)) for a, b in zip(bounds[:-1], bounds[1:])]
            sends, recvs = zip(*(sh.plan_exchange(ranges, n, world, r) for r in range(world)))
            S, R = np.array(sends), np.array(recvs)
            assert np.array_equal(S, R.T)  # what s sends to g is what g expects from s
            assert np.all(R.sum(axis=1) == m) and np.array_equal(S.sum(axis=1), [c for _, c in ranges])
        with pytest.raises(RuntimeError):
            sh.plan_exchange([(0, 10)] + [(10, 0)] * (world - 1), n, world, 0)


def test_sharded_host_logic_under_gloo_world_size_2():
    env = dict(os.environ, OMP_NUM_THREADS="1", MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "_gloo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "gloo worker ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_bench_fan_observations_follow_the_sensor_gate():
    """bench.Truth(fan=True) -- the synthetic input of the C4 fan-sensor line -- detects exactly the landmarks the
    reference's gate admits (oracle.fan_visible, graph_based_slam.py:155-160) and keeps (distance, direction)."""
    sys.path.insert(0, ROOT)
    import bench

    lm = bench.landmarks_for(256)
    tr = bench.Truth(lm, seed=3, T=2, fan=True)
    for _ in range(5):
        tr.advance()
    z = tr.observe()
    assert z.shape == (2, 256, 2)
    for t in range(2):
        pose = tr.pose()[t].reshape(3, 1)
        p = O.world2robot(pose, lm)
        vis = O.fan_visible(p, 15.0, np.deg2rad(80.0))
        assert np.array_equal(np.isfinite(z[t, :, 0]), vis) and 0 < vis.sum() < 256
        d = np.hypot(p[vis, 0], p[vis, 1])
        assert np.all(np.abs(z[t, vis, 0] / d - 1.0) < 0.3)  # 5 % distance noise
        assert np.all(np.abs(O.limit_angle_vec(z[t, vis, 1] - np.arctan2(p[vis, 1], p[vis, 0]))) < 0.25)  # 2 deg noise


def test_reference_animate_consumes_the_dropin_frames(pkg):
    """INTEGRATION.md promises "one changed import": the reference's own animate() (particle_filter.py:248-328), imported
    UNMODIFIED behind oracle/mpl_stub.py with a recording pyplot, is driven frame by frame with what the drop-in class
    returned on a B200 (tests/golden/dropin_frames.npz, recorded by tools/record_dropin_frames.py; the reference's
    sources and a GPU never share a machine here).  px arrives as the drop-in's LazyParticles, as it does for large
    filters.  The same frames are checked against the reference's own seeded run (tests/golden/pf_seed0.npz)."""
    ref_root = os.environ.get("MCL_REFERENCE", "/root/reference")
    frames_path = os.path.join(ROOT, "tests", "golden", "dropin_frames.npz")
    if not os.path.isdir(ref_root):
        pytest.skip("the reference's sources are not on this machine")
    if not os.path.exists(frames_path):
        pytest.skip("tests/golden/dropin_frames.npz has not been recorded yet")
    from oracle import mpl_stub

    mpl_stub.install()
    sys.path.insert(0, ref_root)
    try:
        import particle_filter as ref_pf  # the reference module, unmodified
    finally:
        sys.path.remove(ref_root)

    log = []

    class Ax:
        def __getattr__(self, name):
            def call(*a, **k):
                log.append((name, a, k))
            return call

    class Plt(Ax):
        def subplot2grid(self, *a, **k):
            return Ax()

    ref_pf.plt = Plt()
    fr = np.load(frames_path)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "pf_seed0.npz"))

    class FakeEngine:  # what LazyParticles asks of a Filter
        def __init__(self, px):
            self.n, self._px = px.shape[1], px.astype(np.float32)

        def get_particles(self, want_logw=True):
            return [self._px[0][None], self._px[1][None], self._px[2][None], None]

    class Replay:
        def __init__(self):
            self.i = 0

        def main_pf(self):
            i = self.i
            self.i += 1
            px = pkg.LazyParticles(FakeEngine(fr[f"f{i}_px"]))
            return (fr[f"f{i}_LM"], fr[f"f{i}_x_true"], fr[f"f{i}_x_est"], px, fr[f"f{i}_Q"], int(fr[f"f{i}_idx"]), float(fr[f"f{i}_val"]))

    replay = Replay()
    for i in range(int(fr["n_frames"])):
        log.clear()
        ref_pf.animate(i, replay, 100)
        names = [c[0] for c in log]
        assert names.count("scatter") == 8 and names.count("quiver") == 3 and "annotate" in names and "set_xlim" in names
        cloud = [c for c in log if c[0] == "scatter" and c[2].get("c") == "cyan"]
        np.testing.assert_allclose(cloud[0][1][0], fr[f"f{i}_px"][0], rtol=1e-6)  # every particle was handed to the plot
        txt = [c for c in log if c[0] == "annotate"][0][1][0]
        assert f"[Index]:{int(fr[f'f{i}_idx'])}" in txt
        # the drop-in's frame against the reference's own seeded run: same RNG stream, same truth, same MAP particle
        np.testing.assert_allclose(fr[f"f{i}_x_true"][:, 0], gold["x_true"][i], rtol=0, atol=1e-12)
        assert int(fr[f"f{i}_idx"]) == int(gold["max_idx"][i])
        np.testing.assert_allclose(fr[f"f{i}_x_est"][:, 0], gold["x_est"][i], rtol=1e-5, atol=1e-5)
        assert float(fr[f"f{i}_val"]) == pytest.approx(float(gold["max_val"][i]), rel=2e-3)
