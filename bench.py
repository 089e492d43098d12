#!/usr/bin/env python
"""bench.py -- particle-steps/s of the particle-filter step (predict + weight + resample).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4] [--impl b200|reference]

One "step" is one full pass of the hot path over one filter: systematic resampling (forced every
step: scan + gather) -> motion prediction -> landmark likelihood -> normaliser / ESS / arg-max /
mean+covariance.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024
BYTES_STEP_CANONICAL = 64  # fp32 SoA state, predict+weight 32 B + scan 4 B + gather 28 B
BYTES_K_STEP = 32          # k_step as built: ancestor 4 B + gathered pose 12 B in, pose + log-weight 16 B out
BYTES_K_SCAN = 8           # k_scan: log-weight 4 B in, ancestor 4 B out
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_step launch, per particle, from `ncu --set full` captures of the
# workload named by the key (tools/ncu_case.py; summaries under profiles/).  Workloads without a capture report null.
NCU_TRAFFIC = {
    ("c1", "moment"): ((16.838912e6 + 0.014080e6) / 1048576,
                       "ncu --set full of k_step at C1 itself, 2^20 x 16 (profiles/r02_e_ncu_c1_k_scan_k_step_summary.csv): 16.8 MB read; the "
                       "16.8 MB written stay in the 126 MB L2 (write-back) -- ncu replays start cache-cold, as the rotated replicas do"),
    ("c3", "moment"): ((520.255744e6 + 499.668480e6) / 33554432,
                       "ncu --set full of k_step at 32M particles x 64 landmarks (profiles/r02_d_ncu_32M_x64_k_step_moment_summary.csv), per particle x this launch"),
    ("c3", "pair32"): ((521.427968e6 + 505.885952e6) / 33554432,
                       "ncu --set full of k_step<pair32> at 32M particles x 64 landmarks (profiles/r02_c_ncu_32M_x64_k_step_pair32_summary.csv), per particle x this launch"),
    ("c4", "rb32"): ((195.028992e6 + 254.922752e6) / 16777216,
                     "ncu --set full of k_step<rb32> at C4 itself, 16.8M x 1024 (profiles/r02_h_ncu_c4_rb32_k_step_summary.csv)"),
}
HAVE_RB32 = True
# The LITERAL reference class (unmodified /root/reference/particle_filter.py behind oracle/mpl_stub.py, its per-particle
# Python loop :185-192) at its own default size.  /root/reference does not exist on the GPU box, so this is a constant
# measured in the authoring container (Intel Xeon, 1 core, 60 steps after 5 warm-up steps), not a live measurement.
REFERENCE_LITERAL_C0 = {"value": 3.03e4, "unit": "particle-steps/s", "ms_per_step": 33.0, "config": "1000 particles x 5 landmarks",
                        "where": "authoring container, 1 core of an Intel Xeon; NOT measured on the GPU box (the reference's sources "
                                 "do not travel there)", "kind": "reference (literal class), recorded constant"}
PAIR_LOOPS = {
    "pair": {"arithmetic": "fp64, 8 operations per pair", "bound": "FP64 pipe"},
    "pair32": {"arithmetic": "fp32 about the reference pose, 8 operations per pair, two particles per FFMA2 / FADD2", "bound": "FP32 pipe"},
    "rb": {"arithmetic": "fp64, libm sqrt / atan2, all landmarks scored", "bound": "FP64 pipe"},
    "rb32": {"arithmetic": "fp32 two particles per FFMA2/FMUL2/FADD2 + MUFU (sqrt, 2 rcp), series atan2, ALL landmarks scored",
             "bound": "issue (FMA pipe 54 %, XU 51 %, issue slots 63 %: profiles/r02_h_ncu_c4_rb32_k_step_summary.csv)"},
    "rb_fan": {"arithmetic": "fp32 + MUFU (rcp, sqrt), branch-free atan2, detected landmarks only", "bound": "FP32/SFU issue"},
}

WORKLOADS = {
    # name: (particles per filter, landmarks, trials per GPU, description)
    "c1": dict(key="c1", n=1 << 20, L=16, T=1, desc="C1 single filter 2^20 particles x 16 landmarks"),
    "c2": dict(key="c2", n=16384, L=16, T=4096, desc="C2 4096 independent trials x 16384 particles x 16 landmarks"),
    "c3": dict(key="c3", n=125_000_000, L=64, T=1, desc="C3 one 125M-particle shard x 64 landmarks (1B over 8 GPUs)"),
    "c4": dict(key="c4", n=16 << 20, L=1024, T=1, desc="C4 16M particles x 1024 landmarks (FP-bound likelihood)"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def landmarks_for(L):
    """the built-in 5 landmarks for L == 5, else uniform in [-20, 20]^2, default_rng(1234)."""
    if L == 5:
        return np.array([[5.0, 5.0], [2.0, -3.0], [-3.0, 4.0], [-5.0, -1.0], [0.0, 0.0]])
    return np.random.default_rng(1234).uniform(-20.0, 20.0, (L, 2))


class Truth:
    """The reference's ground-truth circle (particle_filter.py:46-48, :74-79, :121-142) and its
    observation model (:144-154) for `T` independent robots, host side (numpy, vectorised over T).

    The yaw rate is 10 deg/s and dt = 0.1 s, so the trajectory closes after exactly 360 steps:
    `cycle()` returns the observations of one revolution, which the benchmark replays."""

    PERIOD = 360

    def __init__(self, lm, seed, T=1, dt=0.1, fan=False, rb=False):
        self.rb = bool(rb)    # (distance, direction) of ALL landmarks, no visibility gate (MCL_LIK_RB; C4 as named)
        self.fan = bool(fan)  # observations of the fan-shaped scan sensor (graph_based_slam.py:128-167) instead
        self.lm = np.asarray(lm, dtype=np.float64)
        self.rng = np.random.RandomState(seed)
        self.dt = dt
        self.w = np.deg2rad(10.0)
        self.v = 10.0 * self.w
        phase = 0.05 * np.arange(T)
        th = np.pi / 2 + phase
        self.x = np.stack([10.0 * np.cos(phase), 10.0 * np.sin(phase), (th + np.pi) % (2 * np.pi) - np.pi], axis=1)

    def pose(self):
        return self.x.copy()

    def advance(self):
        x, y, th = self.x[:, 0], self.x[:, 1], self.x[:, 2]
        th2 = th + self.w * self.dt
        wrap = np.abs(th2) > np.pi
        th2 = np.where(wrap, (np.abs(th2) - 2 * np.pi) * np.sign(th2), th2)
        self.x = np.stack([x + self.v * self.dt * np.cos(th), y + self.v * self.dt * np.sin(th), th2], axis=1)

    def observe(self):
        """(T, L, 2): world2robot(x_true, LM) + N(0, 0.3^2)."""
        x, y, th = self.x[:, 0:1], self.x[:, 1:2], self.x[:, 2:3]
        s, c = np.sin(th), np.cos(th)
        dx, dy = self.lm[None, :, 0] - x, self.lm[None, :, 1] - y
        z = np.stack([s * dx - c * dy, c * dx + s * dy], axis=2)
        if self.rb:
            # range / bearing of graph_based_slam.py:150-153 with constant noise (MCL_LIK_RB: sigma_d = 0.3 m, sigma_phi = 2 deg)
            d = np.hypot(z[..., 0], z[..., 1])
            phi = np.arctan2(z[..., 1], z[..., 0])
            dn = d + self.rng.normal(0.0, 0.3, d.shape)
            pn = phi + self.rng.normal(0.0, np.deg2rad(2.0), phi.shape)
            return np.stack([dn, (pn + np.pi) % (2 * np.pi) - np.pi], axis=2)
        if self.fan:
            # (distance, direction) of the landmarks inside the fan (15 m, +-80 deg), noise as the reference's Robot
            # sets it (5 % of the distance, 2 deg; graph_based_slam.py:604); NaN rows where nothing is detected
            d = np.hypot(z[..., 0], z[..., 1])
            phi = np.arctan2(z[..., 1], z[..., 0])
            seen = (d <= 15.0) & (z[..., 1] >= np.abs(z[..., 0]) * np.tan(np.pi / 2 - np.deg2rad(80.0)))
            dn = self.rng.normal(d, 0.05 * d)
            pn = phi + self.rng.normal(0.0, np.deg2rad(2.0), phi.shape)
            pn = (pn + np.pi) % (2 * np.pi) - np.pi
            return np.where(seen[..., None], np.stack([dn, pn], axis=2), np.nan)
        return z + self.rng.normal(0.0, 0.3, z.shape)

    def cycle(self):
        """Observations of steps 1..360 (the pose returns to the start: Euler drift is ~1e-13)."""
        start = self.x.copy()
        out = []
        for _ in range(self.PERIOD):
            self.advance()
            out.append(self.observe())
        self.x = start
        return np.stack(out)  # (360, T, L, 2)


class ClockSampler(threading.Thread):
    """NVML SM clock + throttle reasons sampled while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.active = threading.Event()
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        self._names = names
        while not self._stop_evt.is_set():
            if self.active.is_set():
                self.sample_once()
            time.sleep(0.002)

    def sample_once(self):
        """One reading (also usable without the thread: a host-driven loop must not share the driver with NVML)."""
        if self.nv is None:
            return
        nv = self.nv
        names = getattr(self, "_names", None) or {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                                                  0x4: "sw_power_cap"}
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ======================================================================================================
# reference arm / CPU baseline: the numpy oracle (a port of the reference's step; the reference itself is
# a per-particle Python loop that cannot finish these shapes) on the host cores
# ======================================================================================================
_SH = {}  # arrays shared with the worker processes (anonymous shared mappings, inherited through fork)


def _shared(shape, dtype):
    import multiprocessing as mp

    n = int(np.prod(shape))
    raw = mp.RawArray("b", n * np.dtype(dtype).itemsize)
    return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape)


def _cpu_slice(args):
    """One worker's share of a step, entirely inside the shared arrays (nothing is pickled but five scalars):
    gather through the ancestors (:216-221), noise draw + motion (:165-166), log-likelihood (:185-192)."""
    from oracle import pf_oracle as O

    a, b, src, seed, step = args
    S = _SH
    P = S["P"]
    px = S["px"][src][:, S["anc"][a:b]]
    rng = np.random.RandomState((seed * 1000003 + step * 7919 + a) % (2 ** 32))
    v = rng.multivariate_normal([0.0, 0.0, 0.0], P.Q, b - a).T
    px2 = O.predict(px, v, P)
    S["px"][1 - src][:, a:b] = px2
    S["ll"][a:b] = O.loglik_cartesian(px2, S["z"], S["lm"], 0.3, 0.3)
    return 0


def cpu_step_runner(n, L, procs):
    """Returns step() -> seconds for ONE oracle step over n particles x L landmarks on `procs` processes.

    Order as particle_filter.py:86-119: resample (cumsum + comb search on one core, gather in the workers) ->
    predict (numpy noise draw + motion) -> weight (log-domain likelihood; normalisation on one core) -> arg-max.
    Particles, ancestors, observations and log-likelihoods live in shared memory the forked workers inherit, so a
    step moves no particle data between processes."""
    from oracle import pf_oracle as O

    lm = landmarks_for(L)
    truth = Truth(lm, seed=5)
    rng = np.random.RandomState(11)
    P = O.PFParams(100.0, n_particles=n, landmarks=lm)
    S = _SH
    S.clear()
    S.update(P=P, lm=lm, px=[_shared((3, n), np.float64), _shared((3, n), np.float64)], anc=_shared((n,), np.int64),
             ll=_shared((n,), np.float64), z=_shared((L, 2), np.float64))
    S["px"][0][:] = np.tile(P.x0, (1, n)) + rng.normal(0, 0.05, (3, n))
    state = dict(src=0, pw=np.full(n, 1.0 / n), k=0)
    pool = None
    if procs > 1:
        import multiprocessing as mp

        pool = mp.get_context("fork").Pool(procs)  # after the shared arrays exist: the children inherit them
    cuts = np.linspace(0, n, max(procs, 1) * 4 + 1).astype(int)  # 4 slices per process: load balance

    def step():
        t0 = time.perf_counter()
        truth.advance()
        ofs = rng.rand() / n
        S["anc"][:] = O.resample_indices(state["pw"], ofs)
        S["z"][:] = truth.observe()[0]
        jobs = [(int(a), int(b), state["src"], 11, state["k"]) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        if pool is None:
            for j in jobs:
                _cpu_slice(j)
        else:
            pool.map(_cpu_slice, jobs, chunksize=1)
        state["src"] = 1 - state["src"]
        state["k"] += 1
        w, _ = O.update_weights_log(np.full(n, -math.log(n)), S["ll"])
        state["pw"] = w
        O.map_estimate(S["px"][state["src"]], w)
        return time.perf_counter() - t0

    def close():
        if pool is not None:
            pool.terminate()

    return step, close


def cpu_baseline(wl, budget_s=12.0, procs=1):
    """Bounded sample: the same shape with fewer particles when one full step exceeds the budget."""
    n = wl["n"] * wl["T"]
    n_sample = min(n, 1 << 20)
    step, close = cpu_step_runner(n_sample, wl["L"], procs)
    step()  # warm-up (page faults, pool start)
    times = []
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end and len(times) < 50:
        times.append(step())
    close()
    t = float(np.median(times))
    return {"literal_reference_c0": REFERENCE_LITERAL_C0,
            "value": n_sample / t, "unit": "particle-steps/s", "cores": procs, "kind": "port",
            "sample": f"{len(times)} steps of {n_sample} particles x {wl['L']} landmarks (numpy fp64 oracle, "
                      f"{procs} process{'es' if procs > 1 else ''}), median {t * 1e3:.1f} ms/step"}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = max(1, (os.cpu_count() or 1))
    procs = min(procs, 32)
    n = wl["n"] * wl["T"]
    n_sample = min(n, 1 << 20)
    step, close = cpu_step_runner(n_sample, wl["L"], procs)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    k = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    dt = time.perf_counter() - t0
    close()
    value = n_sample * k / dt
    sample = (f"{k} steps of {n_sample} particles x {wl['L']} landmarks per step (bounded sample of {wl['desc']}), "
              f"numpy fp64 oracle port over {procs} processes")
    line = {"impl": "reference", "metric": "particle-steps/s (predict+weight+resample)", "value": value,
            "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": k, "warmup": args.warmup,
            "ms_per_step": dt / k * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "landmarks": wl["L"], "particles_per_step": n_sample},
            "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ======================================================================================================
# the B200 arm
# ======================================================================================================
def _read_clocks(sampler, n=3):
    """A few NVML readings taken while the GPU still works through what has been enqueued (never from a polling
    thread inside the timed region: NVML queries share the driver with the launches)."""
    for _ in range(n):
        sampler.sample_once()


def sharded_parity_check(mcl, torch, dist, dev, transport, n=1 << 18, L=16, steps=8):
    """A small sharded filter over the LIVE process group against the unsharded filter on rank 0: particles and
    log-weights must be equal bit for bit after every step (mixed resampling / non-resampling steps, Philox noise).
    The unsharded filter takes the shards' merged statistics before each step, so both quantise the weights with the
    same map (its own merge groups the fp32 partial sums differently: equal to ~1e-9, not bitwise).
    -> (ok, text).  Also used by tests/_nccl_worker.py."""
    world, rank = dist.get_world_size(), dist.get_rank()
    lm = landmarks_for(L)
    truth = Truth(lm, seed=3)
    sf = mcl.ShardedParticleFilter(n, lm, likelihood="moment", seed=11, transport=transport, max_landmarks=L)
    sf.init_particles(truth.pose()[0])
    ref = None
    if rank == 0:
        ref = mcl.Filter(n_particles=n, landmarks=lm, likelihood="moment", seed=11, max_landmarks=L, device=dev.index)
        ref.init_particles(truth.pose()[0])
    ok, bad = True, ""
    for k in range(steps):
        truth.advance()
        z = truth.observe()
        force = k % 3 != 1
        if rank == 0 and k > 0:
            ref.restore_stats(sf.flt.stats()[0], k)
        sf.step(z, force_resample=force)
        g = sf.drain() if sf.transport == "peer" else sf.global_stats
        x, y, th, lw = sf.local_particles(device=True)
        mine = torch.stack([x, y, th, lw])
        allp = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        if rank == 0:
            ref.step(z, force_resample=force, no_resample=not force)
            rx, ry, rt, rl = ref.get_particles()
            got = torch.cat(allp, dim=1).cpu().numpy()
            st = ref.stats()[0]
            same = all(np.array_equal(got[i], a[0]) for i, a in enumerate((rx, ry, rt, rl)))
            close = abs(g["ess"] / st.ess - 1) < 1e-6 and g["max_idx"] == st.max_idx and abs(g["log_norm"] - st.log_norm) < 1e-7
            if not (same and close):
                ok = False
                bad = bad or f"step {k}: particles equal={same} statistics close={close}"
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    ok = bool(int(flag.item()))
    del sf
    return ok, (f"bit-identical to the unsharded filter over {steps} steps ({n} particles, {world} shards, transport {transport})"
                if ok else f"MISMATCH: {bad}")


def measure_sharded(mcl, torch, dist, dev, n_per_gpu, L, likelihood, steps, warmup, transport="auto", min_seconds=0.5):
    """One filter of world * n_per_gpu particles sharded over the ranks, forced global resampling every step.
    transport "peer": device-driven exchange through peer memory (no NCCL, no host synchronisation in a step);
    "nccl": the host-driven step.  Timed on the device (CUDA events, blocks of `steps` steps, median block, MAX over
    ranks) and, around the same steps, by the host clock (the end-to-end number)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n_global = n_per_gpu * world
    lm = landmarks_for(L)
    truth = Truth(lm, seed=77)  # the same observations on every rank
    cyc = truth.cycle()
    sf = mcl.ShardedParticleFilter(n_global, lm, likelihood=likelihood, seed=4711, max_landmarks=L, transport=transport, lag=1)
    sf.init_particles(truth.pose()[0])
    k = 0

    def run(count):
        nonlocal k
        for _ in range(count):
            sf.step(cyc[k % Truth.PERIOD], force_resample=True)
            k += 1

    # untimed: at least `warmup` steps AND at least 0.6 s of them (clocks of all GPUs leave their idle state)
    W = max(warmup, 3)
    t_warm = time.perf_counter()
    run(W)
    sf.drain()
    torch.cuda.synchronize()
    per_step = torch.tensor([(time.perf_counter() - t_warm) / W], device=dev, dtype=torch.float64)
    dist.all_reduce(per_step, op=dist.ReduceOp.MAX)  # every rank must take the SAME number of steps
    per_step = max(float(per_step.item()), 1e-4)
    run(int(min(600, max(0.0, 0.6 / per_step))))
    sf.drain()
    n_blocks = int(min(400, max(1, math.ceil(min_seconds / (per_step * steps)))))
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index)
    launches0 = sf.flt.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_blocks)]
    for e0, e1 in evs:
        e0.record()
        run(steps)
        e1.record()
    _read_clocks(sampler)  # (the GPU is still working through the queue; NVML queries take milliseconds)
    g = sf.drain()
    torch.cuda.synchronize()
    launches = (sf.flt.launch_count() - launches0) // n_blocks  # per block of `steps` steps, as `steps` is reported
    # end to end: the same number of steps again, by the host clock alone (no NVML query inside the region)
    dist.barrier()
    t0 = time.perf_counter()
    run(steps * n_blocks)
    g = sf.drain()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dist.barrier()
    block_ms = float(np.median([e0.elapsed_time(e1) for e0, e1 in evs]))
    t = torch.tensor([block_ms, wall], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    block_ms, wall = float(t[0].item()), float(t[1].item())
    # per-kernel durations of a few more steps (CUDA events inside the library, same stream)
    kern = None
    moved = None
    if sf.transport == "peer":
        sf.flt.profile(True)
        acc = np.zeros(5)
        for _ in range(8):
            run(1)
            acc += np.array(sf.flt.profile_read_peer())
        sf.drain()
        sf.flt.profile(False)
        acc /= 8
        kern = {"totals_pass+exchange": acc[0], "scalars+expansion+exchange": acc[1], "unused": acc[2], "k_step": acc[3], "statistics_exchange": acc[4]}
        pl = sf.plan()
        moved = int(sf.m - (pl.win_hi - pl.win_lo))
    mv = torch.tensor([float(moved or 0)], device=dev, dtype=torch.float64)
    dist.all_reduce(mv, op=dist.ReduceOp.MAX)
    out = {"workload": f"one filter of {n_global} particles x {L} landmarks over {world} GPUs, global systematic resampling",
           "value": n_global * steps / (block_ms * 1e-3), "unit": "particle-steps/s", "ms_per_step": block_ms / steps, "steps": steps,
           "blocks": n_blocks, "timed_seconds": block_ms * n_blocks * 1e-3,
           "e2e_value": n_global * steps * n_blocks / wall, "e2e_ms_per_step": wall / (steps * n_blocks) * 1e3,
           "ess_over_n": g["ess"] / n_global, "gpu_launches_rank0": int(launches), "clocks_rank0": sampler.summary(),
           "transport": sf.transport, "transport_note": sf.transport_note, "kernel_ms_rank0": kern,
           "max_slots_received_from_peers_last_step": float(mv.item()),
           "exchange": ("statistics record (320 B), u64 weight total and off-rank poses written by the producing kernels into the "
                        "peers' memory (CUDA IPC mappings, NVLink); step-numbered flags; gate / k_offset / slot ranges derived on "
                        "the device; no NCCL call and no host synchronisation inside a step; statistics read back one step late"
                        if sf.transport == "peer" else
                        "allgather(mcl_stats) -> device merge; allgather(u64 total); D2H + host planning; one batched NCCL P2P group")}
    del sf
    torch.cuda.empty_cache()
    return out


def run_b200(args, wl):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import slam_robot_simu_b200 as mcl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = peaks()
    K, W = args.steps, max(args.warmup, 3)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_run(w_, K_, W_, spin_s, likelihood, min_seconds=0.25, keep=False, trials=None):
        """Blocks of K_ timed steps of workload w_ with device-resident inputs (at least min_seconds of them; the
        MEDIAN block is reported, MAX over ranks); returns the numbers of the JSON line."""
        n_, L_, T_ = w_["n"], w_["L"], (trials if trials is not None else w_["T"])
        lm_ = landmarks_for(L_)
        state_bytes = 2 * 16 * n_ * T_ + 4 * n_ * T_
        replicas = max(1, min(64, math.ceil(2.2 * L2_BYTES / state_bytes)))  # rotate filters: every step is L2-cold
        filters, zcyc_host, zcyc = [], [], []
        for r in range(replicas):
            f = mcl.Filter(n_particles=n_, n_trials=T_, landmarks=lm_, likelihood=likelihood, device=local,
                           seed=42 + 1000 * rank + r, max_landmarks=L_,
                           sigma_obs=[0.3, float(np.deg2rad(2.0))] if likelihood in ("rb", "rb32") else None)
            tr = Truth(lm_, seed=7 + 100 * rank + r * 17, T=T_, fan=(likelihood == "rb_fan"), rb=(likelihood in ("rb", "rb32")))
            f.init_particles(tr.pose())
            cyc = tr.cycle()  # (360, T, L, 2): replica r at filter step k sees cyc[k % 360]
            filters.append(f)
            zcyc_host.append(cyc)
            zcyc.append(torch.from_numpy(cyc).to(dev))
        torch.cuda.synchronize()
        counts = [0] * replicas  # steps taken by each replica

        def step_dev(r):
            filters[r].step_dev(zcyc[r][counts[r] % Truth.PERIOD], force_resample=True)
            counts[r] += 1

        # untimed: 20 spreading steps per replica, then spin until the clocks are at their loaded state
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(replicas):
            for _ in range(20):
                step_dev(r)
        e1.record()
        torch.cuda.synchronize()
        est_ms = max(e0.elapsed_time(e1) / (20 * replicas), 1e-3)
        t_spin = time.perf_counter()
        i = 0
        while time.perf_counter() - t_spin < spin_s:
            step_dev(i % replicas)
            i += 1
            if i % 64 == 0:
                torch.cuda.synchronize()
        for i in range(W_):
            step_dev(i % replicas)
        # the step time at the loaded state sizes the timed region (the first, cold steps above overestimate it)
        e0.record()
        for i in range(2 * K_):
            step_dev(i % replicas)
        e1.record()
        torch.cuda.synchronize()
        est_ms = min(est_ms, max(e0.elapsed_time(e1) / (2 * K_), 1e-3))
        n_blocks = int(min(4000, max(1, math.ceil(1.15 * min_seconds * 1e3 / (est_ms * K_)))))
        if world > 1:  # the same number of blocks on every rank
            nb = torch.tensor([n_blocks], device=dev)
            dist.all_reduce(nb, op=dist.ReduceOp.MAX)
            n_blocks = int(nb.item())
        sampler = ClockSampler(local)
        launches0 = sum(f.launch_count() for f in filters)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_blocks)]
        sync_all()
        i = W_
        for b0, b1 in evs:
            b0.record()
            for _ in range(K_):
                step_dev(i % replicas)
                i += 1
            b1.record()
        _read_clocks(sampler)  # all launches are enqueued; the GPU is still working through them
        sync_all()
        blocks = [b0.elapsed_time(b1) for b0, b1 in evs]
        ms_ = float(np.median(blocks))
        launches_ = (sum(f.launch_count() for f in filters) - launches0) // n_blocks
        if world > 1:
            t = torch.tensor([ms_], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        value_ = float(n_) * T_ * K_ * world / (ms_ * 1e-3)

        # per-kernel durations, measured live with CUDA events inside the library (same steps, L2-cold)
        kt = np.zeros(3)
        kn = int(min(200, max(8, 0.05e3 / est_ms)))
        for f in filters:
            f.profile(True)
        for j in range(kn):
            r = (i + j) % replicas
            step_dev(r)
            kt += np.array(filters[r].profile_read())
        ess_frac_ = float(np.mean([st.ess for st in filters[0].stats()])) / n_
        for f in filters:
            f.profile(False)
        kt /= kn
        ach = BYTES_K_STEP * n_ * T_ / (kt[2] * 1e-3) / 1e9
        tr_bpp, tr_src = NCU_TRAFFIC.get((w_.get("key"), likelihood), (None, "no ncu --set full capture of this workload"))
        roof = {"bound": "hbm", "kernel": "k_step (gather + predict + weight + statistics)", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": None if tr_bpp is None else tr_bpp * n_ * T_,
                "traffic_source": tr_src, "peak_source": peak_src, "algorithmic_bytes_per_particle": BYTES_K_STEP,
                "kernel_ms": {"k_scan": kt[1], "k_step": kt[2]},
                "scan_achieved_GBps": BYTES_K_SCAN * n_ * T_ / (kt[1] * 1e-3) / 1e9,
                "scan_frac": BYTES_K_SCAN * n_ * T_ / (kt[1] * 1e-3) / 1e9 / peak,
                "step_bytes_moved_per_particle": BYTES_K_STEP + BYTES_K_SCAN,
                "step_moved_frac": (BYTES_K_STEP + BYTES_K_SCAN) * value_ / world / 1e9 / peak,
                "step_canonical_bytes": BYTES_STEP_CANONICAL,
                "step_canonical_frac": BYTES_STEP_CANONICAL * value_ / world / 1e9 / peak}
        out = dict(value=value_, ms=ms_, launches=launches_, roofline=roof, ess=ess_frac_, replicas=replicas, blocks=n_blocks,
                   timed_seconds=float(np.sum(blocks)) * 1e-3, state_bytes=state_bytes, clocks=sampler.summary(),
                   landmarks_seen=float(np.mean([np.isfinite(c[..., 0]).sum(axis=-1).mean() for c in zcyc_host])))
        if keep:
            out.update(filters=filters, zcyc_host=zcyc_host, counts=counts)
        else:
            del filters, zcyc
            torch.cuda.empty_cache()
        return out

    def extra_line(w_, lr, K_):
        return {"workload": w_["desc"], "value": lr["value"], "unit": "particle-steps/s", "ms_per_step": lr["ms"] / K_,
                "steps": K_, "blocks": lr["blocks"], "ess_over_n": lr["ess"], "roofline": lr["roofline"], "clocks": lr["clocks"]}

    def pair_loop(w_, lr, likelihood, K_):
        # pairs_per_s: from k_step timed alone (CUDA events around single launches, the GPU idle in between: burst
        # clocks); pairs_per_s_sustained: from the back-to-back timed steps (a likelihood-bound kernel on the whole chip
        # runs into the power cap -- see this entry's `clocks`)
        k_ms = lr["roofline"]["kernel_ms"]["k_step"]
        seen = lr["landmarks_seen"] if likelihood == "rb_fan" else float(w_["L"])
        return {"landmarks_scored_per_step": seen, "pairs_per_s": w_["n"] * w_["T"] * seen / (k_ms * 1e-3),
                "pairs_per_s_sustained": w_["n"] * w_["T"] * seen / (lr["ms"] / K_ * 1e-3), **PAIR_LOOPS[likelihood]}

    # ------------------------------------------------------------------------------------------------------
    # N > 1: the headline is ONE filter sharded over the N GPUs (BASELINE.json configs[3]); extras: one independent
    # C1 filter per rank (no data-path collective), C2's trials split by rank, and the same 125 M shard run alone
    # ------------------------------------------------------------------------------------------------------
    if world > 1:
        c3 = WORKLOADS["c3"]
        n_shard = args.n if args.n else c3["n"]
        parity_ok, parity_txt = sharded_parity_check(mcl, torch, dist, dev, args.transport)
        sh = measure_sharded(mcl, torch, dist, dev, n_shard, c3["L"], args.likelihood, K, W, transport=args.transport)
        sh["parity"] = parity_txt
        extras = {}
        if not args.no_extras:
            lr = timed_run(c3, 12, 3, 0.1, args.likelihood, min_seconds=0.1)
            extras["one_gpu_shard_alone"] = extra_line(c3, lr, 12)
            lr = timed_run(WORKLOADS["c1"], max(K, 20), W, 0.2, args.likelihood)
            extras["replicas_c1"] = dict(extra_line(WORKLOADS["c1"], lr, max(K, 20)),
                                         parallelism=f"{world} ranks, one independent C1 filter each (no data-path collective)")
            c2 = WORKLOADS["c2"]
            lr = timed_run(c2, 12, 3, 0.1, args.likelihood, min_seconds=0.1, trials=c2["T"] // world)
            extras["c2_trials_split_by_rank"] = dict(extra_line(c2, lr, 12), scaling="strong",
                                                     parallelism=f"{c2['T']} trials, {c2['T'] // world} per rank, no collective")
        if rank == 0:
            frac = BYTES_STEP_CANONICAL * sh["value"] / world / 1e9 / peak
            kern = sh.get("kernel_ms_rank0") or {}
            k_step_ms = kern.get("k_step")
            line = {"metric": "particle-steps/s (predict+weight+resample)", "value": sh["value"], "unit": "particle-steps/s",
                    "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": sh["ms_per_step"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f32 state / f64 likelihood / u64 scan", "data": "synthetic",
                    "config": {"workload": sh["workload"], "likelihood": args.likelihood, "noise": "philox4x32-10 on device",
                               "resampling": "forced every step, global", "ess_over_n": sh["ess_over_n"],
                               "l2": f"{36 * n_shard >> 20} MiB of state per GPU > 126 MiB L2", "parallelism": f"particle-sharded x{world}",
                               "transport": sh["transport"], "timing": f"median of {sh['blocks']} blocks of {K} steps, CUDA events, MAX over ranks",
                               "note": "N = 1 runs C1 (one 2^20-particle filter); the weak-scaling base of THIS line is extras.one_gpu_shard_alone"},
                    "roofline": {"bound": "hbm", "kernel": "k_step (gather + predict + weight + statistics), rank 0, inside the sharded step",
                                 "achieved": None if not k_step_ms else BYTES_K_STEP * n_shard / (k_step_ms * 1e-3) / 1e9, "peak": peak,
                                 "unit": "GB/s", "frac": None if not k_step_ms else BYTES_K_STEP * n_shard / (k_step_ms * 1e-3) / 1e9 / peak,
                                 "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_particle": BYTES_K_STEP,
                                 "kernel_ms": kern, "step_canonical_bytes": BYTES_STEP_CANONICAL, "step_canonical_frac": frac,
                                 "note": "step_canonical_frac = whole step per GPU at SURVEY 8(d)'s canonical 64 B per particle-step"},
                    "cpu_baseline": None,
                    "e2e": {"value": sh["e2e_value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 16 * c3["L"],
                            "d2h_bytes_per_step": C.sizeof(mcl.mcl_stats), "ms_per_step": sh["e2e_ms_per_step"],
                            "api": "ShardedParticleFilter(..., lag=1).step(z_host, force_resample=True)",
                            "note": "host clock around the same steps: observation upload every step, merged statistics read back every "
                                    "step through a two-deep pinned ring (consumed one step late)"},
                    "sharded_filter": sh, "extras": extras, "gpu_launches": sh["gpu_launches_rank0"] * world, "clocks": sh["clocks_rank0"]}
            if not parity_ok:
                line["error"] = "sharded filter differs from the unsharded one: " + parity_txt
            print(json.dumps(line), flush=True)
        dist.destroy_process_group()
        return

    # ------------------------------------------------------------------------------------------------------
    # N = 1
    # ------------------------------------------------------------------------------------------------------
    n, L, T = wl["n"], wl["L"], wl["T"]
    lm = landmarks_for(L)
    main_run = timed_run(wl, K, W, 0.3, args.likelihood, keep=True)
    value, ms, launches, roofline, ess_frac = (main_run[k] for k in ("value", "ms", "launches", "roofline", "ess"))
    replicas, state_bytes, filters, zcyc_host, counts = (main_run[k] for k in ("replicas", "state_bytes", "filters", "zcyc_host", "counts"))
    clocks = main_run["clocks"]
    n_blocks_main, timed_s_main, seen_main = main_run["blocks"], main_run["timed_seconds"], main_run["landmarks_seen"]

    # ---- end to end through the reference-facing API: host observation in, host estimate out, every step ----
    e2e = None
    if T == 1 and args.likelihood != "rb":
        sync_all()
        pf = mcl.ParticleFilter(100, n_particles=n, landmarks=lm, noise="philox", likelihood=args.likelihood,
                                lazy_px=True, force_resample=True, device=local, seed=4242 + rank)
        for _ in range(max(W, 20)):
            pf.main_pf()
        ke = int(min(20000, max(K, 0.25 / max(60e-6, 1.7 * ms / K * 1e-3))))  # about 0.25 s of calls
        sync_all()
        t0 = time.perf_counter()
        for _ in range(ke):
            out = pf.main_pf()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": float(n) * ke / dt, "unit": "particle-steps/s", "h2d_bytes_per_step": 16 * L,
               "d2h_bytes_per_step": C.sizeof(mcl.mcl_stats), "steps": ke, "ms_per_step": dt / ke * 1e3,
               "api": "ParticleFilter(period_ms, noise='philox', lazy_px=True, force_resample=True).main_pf()",
               "note": "host truth motion + observation draw, z_l read by k_scan from the pinned staging slot over PCIe (no copy operation), k_scan + k_step + k_merge, the mcl_stats block of THIS step written by k_merge into pinned host memory "
                       "(no lag); px stays on the device until read (LazyParticles), so the timing excludes the 12 B/particle "
                       "px copy the reference's return value implies"}
        del pf
    else:
        sync_all()
        f = filters[0]
        ke = min(K, 200)
        t0 = time.perf_counter()
        for i in range(ke):
            f.step(zcyc_host[0][counts[0] % Truth.PERIOD], force_resample=True)
            counts[0] += 1
            st = f.stats()
        dt = time.perf_counter() - t0
        e2e = {"value": float(n) * T * ke / dt, "unit": "particle-steps/s", "h2d_bytes_per_step": 16 * L * T,
               "d2h_bytes_per_step": C.sizeof(mcl.mcl_stats) * T, "steps": ke, "api": "Filter.step(z_host) + Filter.stats()"}
    del filters, main_run
    torch.cuda.empty_cache()

    extras = {}
    large = None
    if args.workload == "c1" and not args.no_extras:
        # the per-GPU share of the metric's target configuration (1B particles x 64 landmarks over 8 GPUs): the same
        # kernels at a size where launch and tail latencies no longer dominate
        c3 = WORKLOADS["c3"]
        lr = timed_run(c3, 12, 3, 0.1, args.likelihood, min_seconds=0.1)
        large = extra_line(c3, lr, 12)
        if args.likelihood == "moment":
            # the GENERAL likelihood -- every particle against every landmark (particle_filter.py:185-192) -- on the
            # same two shapes; the headline runs its exact O(1)-per-particle moment form
            for lik, key, k_ in (("pair32", "c1", 40), ("pair32", "c3", 8), ("pair", "c3", 6)):
                w_ = WORKLOADS[key]
                lr = timed_run(w_, k_, 3, 0.1, lik, min_seconds=0.1)
                extras[f"{lik}_{key}"] = dict(extra_line(w_, lr, k_), likelihood=lik, pair_loop=pair_loop(w_, lr, lik, k_))
            c2 = WORKLOADS["c2"]
            lr = timed_run(c2, 12, 3, 0.1, "moment", min_seconds=0.1)
            extras["c2"] = extra_line(c2, lr, 12)
            if HAVE_RB32:
                c4 = WORKLOADS["c4"]
                lr = timed_run(c4, 4, 3, 0.0, "rb32", min_seconds=0.4)  # (25 ms per step: four blocks of four)
                extras["c4_rb_all_landmarks"] = dict(extra_line(c4, lr, 4), likelihood="rb32", pair_loop=pair_loop(c4, lr, "rb32", 4))
    cpu = cpu_baseline(wl, procs=max(1, min(32, os.cpu_count() or 1))) if not args.no_cpu else None
    line = {"metric": "particle-steps/s (predict+weight+resample)", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 state / f64 likelihood / u64 scan", "data": "synthetic",
            "config": {"workload": wl["desc"], "particles": n, "landmarks": L, "trials": T, "likelihood": args.likelihood,
                       "same_algorithm_as_reference_arm": args.likelihood != "moment",
                       "algorithm_note": ("the moment likelihood is an exact algebraic identity of the per-pair sum (SURVEY 7), O(1) per "
                                          "particle; the reference arm scores every (particle, landmark) pair -- the like-for-like GPU "
                                          "number is extras.pair32_c1") if args.likelihood == "moment" else "per-pair likelihood on both arms",
                       "noise": "philox4x32-10 on device", "resampling": "forced every step",
                       "ess_over_n": ess_frac,
                       "l2": f"rotating {replicas} independent filter replicas ({replicas * state_bytes >> 20} MiB > 126 MiB L2)",
                       "timing": f"median of {n_blocks_main} blocks of {K} steps (CUDA events), {timed_s_main:.3f} s timed",
                       "parallelism": "single GPU"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "large_filter": large, "extras": extras,
            "gpu_launches": int(launches), "clocks": clocks}
    if args.likelihood in PAIR_LOOPS:
        seen = seen_main if args.likelihood == "rb_fan" else float(L)
        line["pair_loop"] = {"landmarks_scored_per_step": seen, "pairs_per_s": n * T * seen / (roofline["kernel_ms"]["k_step"] * 1e-3),
                             **PAIR_LOOPS[args.likelihood]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--likelihood", default="moment", choices=["pair", "pair32", "moment", "rb", "rb32", "rb_fan"])
    ap.add_argument("--transport", default="auto", choices=["auto", "peer", "nccl"], help="N > 1: exchange of the sharded filter")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra workloads of the default line")
    ap.add_argument("--n", type=int, default=0, help="override the particle count of the workload (experiments)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.n:
        wl["n"] = args.n
        wl["desc"] += f" [particle count overridden: {args.n}]"
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
