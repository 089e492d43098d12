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
NCU_TRAFFIC_K_STEP_BYTES_PER_PARTICLE = (528.388608e6 + 501.015552e6) / 33554432  # profiles/r01_r_ncu_32M_x64_k_step_summary.csv

WORKLOADS = {
    # name: (particles per filter, landmarks, trials per GPU, description)
    "c1": dict(n=1 << 20, L=16, T=1, desc="C1 single filter 2^20 particles x 16 landmarks"),
    "c2": dict(n=16384, L=16, T=4096, desc="C2 4096 independent trials x 16384 particles x 16 landmarks"),
    "c3": dict(n=125_000_000, L=64, T=1, desc="C3 one 125M-particle shard x 64 landmarks (1B over 8 GPUs)"),
    "c4": dict(n=16 << 20, L=1024, T=1, desc="C4 16M particles x 1024 landmarks (FP-bound likelihood)"),
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

    def __init__(self, lm, seed, T=1, dt=0.1, fan=False):
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
def _cpu_chunk(args):
    from oracle import pf_oracle as O

    px, v, z, lm, dt = args
    P = O.PFParams(dt * 1000.0, n_particles=px.shape[1], landmarks=lm)
    px2 = O.predict(px, v, P)
    ll = O.loglik_cartesian(px2, z, lm, 0.3, 0.3)
    return px2, ll


def cpu_step_runner(n, L, procs):
    """Returns step() -> seconds for ONE oracle step over n particles x L landmarks, `procs` processes.

    Order as particle_filter.py:86-119: resample (cumsum + comb search + gather) -> predict (numpy
    noise draw + motion) -> weight (log-domain likelihood, normalise) -> arg-max."""
    from oracle import pf_oracle as O

    lm = landmarks_for(L)
    truth = Truth(lm, seed=5)
    rng = np.random.RandomState(11)
    P = O.PFParams(100.0, n_particles=n, landmarks=lm)
    state = dict(px=np.tile(P.x0, (1, n)) + rng.normal(0, 0.05, (3, n)), pw=np.full(n, 1.0 / n))
    pool = None
    if procs > 1:
        import multiprocessing as mp

        pool = mp.get_context("fork").Pool(procs)

    def step():
        t0 = time.perf_counter()
        truth.advance()
        ofs = rng.rand() / n
        anc = O.resample_indices(state["pw"], ofs)
        px = state["px"][:, anc]
        v = rng.multivariate_normal([0.0, 0.0, 0.0], P.Q, n).T
        z = truth.observe()[0]
        if pool is None:
            px2, ll = _cpu_chunk((px, v, z, lm, 0.1))
        else:
            cuts = np.linspace(0, n, procs + 1).astype(int)
            parts = pool.map(_cpu_chunk, [(px[:, a:b], v[:, a:b], z, lm, 0.1) for a, b in zip(cuts[:-1], cuts[1:])])
            px2 = np.concatenate([p[0] for p in parts], axis=1)
            ll = np.concatenate([p[1] for p in parts])
        w, _ = O.update_weights_log(np.full(n, -math.log(n)), ll)
        state["px"], state["pw"] = px2, w
        O.map_estimate(px2, w)
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
    return {"value": n_sample / t, "unit": "particle-steps/s", "cores": procs, "kind": "port",
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
def measure_sharded(mcl, torch, dist, dev, n_per_gpu, L, likelihood, steps, warmup):
    """One filter of world * n_per_gpu particles sharded over the ranks: allgather of statistics and
    weight totals + variable-count all-to-all of the resampled poses (NCCL), forced resampling."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n_global = n_per_gpu * world
    lm = landmarks_for(L)
    truth = Truth(lm, seed=77)  # the same observations on every rank
    cyc = truth.cycle()
    sf = mcl.ShardedParticleFilter(n_global, lm, likelihood=likelihood, seed=4711, max_landmarks=L)
    sf.init_particles(truth.pose()[0])
    k = 0
    # untimed: at least `warmup` steps AND at least 0.6 s of them -- a sharded step is ~2 ms, so a handful of steps
    # ends before the clocks of all GPUs have left their idle state (W is a minimum, not a maximum)
    t_warm = time.perf_counter()
    for _ in range(max(warmup, 3)):
        sf.step(cyc[k % Truth.PERIOD], force_resample=True)
        k += 1
    torch.cuda.synchronize()
    # the steps contain collectives, so every rank must take the SAME number of extra steps: agree on it first
    per_step = torch.tensor([(time.perf_counter() - t_warm) / max(warmup, 3)], device=dev, dtype=torch.float64)
    dist.all_reduce(per_step, op=dist.ReduceOp.MAX)
    extra = int(min(400, max(0.0, 0.6 / max(float(per_step.item()), 1e-4) - max(warmup, 3))))
    for _ in range(extra):
        sf.step(cyc[k % Truth.PERIOD], force_resample=True)
        k += 1
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    # The sharded step is host-driven, and NVML queries share the driver with the launches (a polling thread
    # slowed the step 2.5x): the clocks are read once right before and once right after the timed steps, while the
    # GPU is still loaded by them.
    sampler = ClockSampler(dev.index)
    sampler.sample_once()
    launches0 = sf.flt.launch_count()
    t0 = time.perf_counter()
    moved = 0
    for _ in range(steps):
        g = sf.step(cyc[k % Truth.PERIOD], force_resample=True)
        k += 1
        first, count = sf.last_ranges[rank]
        lo, hi = rank * sf.m, (rank + 1) * sf.m
        moved += count - max(0, min(first + count, hi) - max(first, lo))
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sampler.sample_once()
    launches = sf.flt.launch_count() - launches0
    t = torch.tensor([dt, float(moved)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0].item())
    out = {"workload": f"one filter of {n_global} particles x {L} landmarks over {world} GPUs, global systematic resampling",
           "value": n_global * steps / dt, "unit": "particle-steps/s", "ms_per_step": dt / steps * 1e3, "steps": steps,
           "ess_over_n": g["ess"] / n_global, "gpu_launches_rank0": int(launches), "clocks_rank0": sampler.summary(),
           "max_slots_sent_off_rank_per_step": float(t[1].item()) / steps,
           "collectives": "allgather(mcl_stats 312 B) -> device merge; allgather(u64 weight total); one batched NCCL P2P group for the fp32 poses of off-rank slots"}
    del sf
    torch.cuda.empty_cache()
    return out


def run_b200(args, wl):
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

    n, L, T = wl["n"], wl["L"], wl["T"]
    lm = landmarks_for(L)
    if world > 1 and args.workload == "c3":
        sh = measure_sharded(mcl, torch, dist, dev, n, L, args.likelihood, args.steps, args.warmup)
        if rank == 0:
            peak, peak_src = peaks()
            line = {"metric": "particle-steps/s (predict+weight+resample)", "value": sh["value"], "unit": "particle-steps/s",
                    "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": sh["ms_per_step"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f32 state / f64 likelihood / u64 scan", "data": "synthetic",
                    "config": {"workload": sh["workload"], "likelihood": args.likelihood, "noise": "philox4x32-10 on device",
                               "resampling": "forced every step, global", "ess_over_n": sh["ess_over_n"],
                               "l2": f"{36 * n >> 20} MiB of state per GPU > 126 MiB L2", "parallelism": f"particle-sharded x{world}"},
                    "roofline": {"bound": "hbm", "achieved": BYTES_STEP_CANONICAL * sh["value"] / world / 1e9, "peak": peak,
                                 "unit": "GB/s", "frac": BYTES_STEP_CANONICAL * sh["value"] / world / 1e9 / peak, "traffic": None,
                                 "peak_source": peak_src, "note": "whole step per GPU at the canonical 64 B per particle-step"},
                    "cpu_baseline": None, "e2e": {"value": sh["value"], "unit": "particle-steps/s",
                                                  "h2d_bytes_per_step": 16 * L + 304, "d2h_bytes_per_step": 304 + 8 + 16,
                                                  "note": "the sharded step is host-mediated: its timing already includes the "
                                                          "observation upload and the statistics read-back of every step"},
                    "sharded_filter": sh, "gpu_launches": sh["gpu_launches_rank0"] * world, "clocks": sh["clocks_rank0"]}
            print(json.dumps(line), flush=True)
        dist.destroy_process_group()
        return
    K, W = args.steps, max(args.warmup, 3)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_run(w_, K_, W_, spin_s):
        """K_ timed steps of workload w_ with device-resident inputs; returns the numbers of the JSON line."""
        n_, L_, T_ = w_["n"], w_["L"], w_["T"]
        lm_ = landmarks_for(L_)
        state_bytes = 2 * 16 * n_ * T_ + 4 * n_ * T_
        replicas = max(1, min(64, math.ceil(2.2 * L2_BYTES / state_bytes)))  # rotate filters: every step is L2-cold
        filters, zcyc_host, zcyc = [], [], []
        for r in range(replicas):
            f = mcl.Filter(n_particles=n_, n_trials=T_, landmarks=lm_, likelihood=args.likelihood, device=local,
                           seed=42 + 1000 * rank + r, max_landmarks=L_)
            tr = Truth(lm_, seed=7 + 100 * rank + r * 17, T=T_, fan=(args.likelihood == "rb_fan"))
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

        # untimed: 20 spreading steps per replica, then spin until the clocks are at their loaded
        # state (W is a minimum, not a maximum)
        for r in range(replicas):
            for _ in range(20):
                step_dev(r)
        t_spin = time.perf_counter()
        i = 0
        while time.perf_counter() - t_spin < spin_s:
            step_dev(i % replicas)
            i += 1
            if i % 64 == 0:
                torch.cuda.synchronize()
        for i in range(W_):
            step_dev(i % replicas)

        sampler = ClockSampler(local)
        sampler.start()
        launches0 = sum(f.launch_count() for f in filters)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        sampler.active.set()
        e0.record()
        for i in range(K_):
            step_dev((W_ + i) % replicas)
        e1.record()
        sync_all()
        sampler.active.clear()
        sampler.stop()
        ms_ = e0.elapsed_time(e1)
        launches_ = sum(f.launch_count() for f in filters) - launches0
        if world > 1:
            t = torch.tensor([ms_], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        value_ = float(n_) * T_ * K_ * world / (ms_ * 1e-3)

        # per-kernel durations, measured live with CUDA events inside the library (same steps, L2-cold)
        kt = np.zeros(3)
        kn = min(K_, 200)
        for f in filters:
            f.profile(True)
        for i in range(kn):
            r = (W_ + i) % replicas
            step_dev(r)
            kt += np.array(filters[r].profile_read())
        ess_frac_ = float(np.mean([st.ess for st in filters[0].stats()])) / n_
        for f in filters:
            f.profile(False)
        kt /= kn
        peak, peak_src = peaks()
        ach = BYTES_K_STEP * n_ * T_ / (kt[2] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_step (gather + predict + weight + statistics)", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": NCU_TRAFFIC_K_STEP_BYTES_PER_PARTICLE * n_ * T_,
                "traffic_source": "ncu --set full, 32M particles x 64 landmarks: dram__bytes_read + write of k_step per "
                                  "particle (profiles/r01_r_ncu_32M_x64_k_step_summary.csv), scaled to this launch",
                "peak_source": peak_src, "algorithmic_bytes_per_particle": BYTES_K_STEP,
                "kernel_ms": {"k_scan": kt[1], "k_step": kt[2]},
                "scan_achieved_GBps": BYTES_K_SCAN * n_ * T_ / (kt[1] * 1e-3) / 1e9,
                "step_bytes_moved_per_particle": BYTES_K_STEP + BYTES_K_SCAN,
                "step_moved_frac": (BYTES_K_STEP + BYTES_K_SCAN) * value_ / world / 1e9 / peak,
                "step_canonical_bytes": BYTES_STEP_CANONICAL,
                "step_canonical_frac": BYTES_STEP_CANONICAL * value_ / world / 1e9 / peak}
        return dict(value=value_, ms=ms_, launches=launches_, roofline=roof, ess=ess_frac_, replicas=replicas,
                    state_bytes=state_bytes, clocks=sampler.summary(), filters=filters, zcyc_host=zcyc_host, counts=counts)

    main_run = timed_run(wl, K, W, 0.3)
    value, ms, launches, roofline, ess_frac = (main_run[k] for k in ("value", "ms", "launches", "roofline", "ess"))
    replicas, state_bytes, filters, zcyc_host, counts = (main_run[k] for k in ("replicas", "state_bytes", "filters", "zcyc_host", "counts"))
    clocks = main_run["clocks"]

    # ---- end to end through the reference-facing API: host observation in, host estimate out, every step ----
    e2e = None
    if T == 1:
        sync_all()
        pf = mcl.ParticleFilter(100, n_particles=n, landmarks=lm, noise="philox", likelihood=args.likelihood,
                                lazy_px=True, force_resample=True, device=local, seed=4242 + rank)
        for _ in range(max(W, 20)):
            pf.main_pf()
        ke = min(K, 500)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(ke):
            out = pf.main_pf()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        import ctypes as C

        e2e = {"value": float(n) * ke * world / dt, "unit": "particle-steps/s", "h2d_bytes_per_step": 16 * L,
               "d2h_bytes_per_step": C.sizeof(mcl.mcl_stats), "steps": ke,
               "api": "ParticleFilter(period_ms, noise='philox', lazy_px=True, force_resample=True).main_pf()",
               "note": "host truth motion + observation draw, H2D of z_l, k_scan + k_step + k_merge, D2H of the mcl_stats block; px stays "
                       "on the device until read (LazyParticles)"}
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
        import ctypes as C

        e2e = {"value": float(n) * T * ke * world / dt, "unit": "particle-steps/s", "h2d_bytes_per_step": 16 * L * T,
               "d2h_bytes_per_step": C.sizeof(mcl.mcl_stats) * T, "steps": ke, "api": "Filter.step(z_host) + Filter.stats()"}
    sharded = None
    if world > 1 and not args.no_sharded:
        del filters, main_run
        torch.cuda.empty_cache()
        c3 = WORKLOADS["c3"]
        sharded = measure_sharded(mcl, torch, dist, dev, c3["n"], c3["L"], args.likelihood, 8, 3)
    large = None
    if world == 1 and args.workload == "c1" and not args.no_large:
        # the per-GPU share of the metric's target configuration (1B particles x 64 landmarks over 8 GPUs): the
        # same kernels at a size where launch and tail latencies no longer dominate
        del filters, main_run
        torch.cuda.empty_cache()
        c3 = WORKLOADS["c3"]
        lr = timed_run(c3, 12, 3, 0.1)
        large = {"workload": c3["desc"], "value": lr["value"], "unit": "particle-steps/s", "ms_per_step": lr["ms"] / 12,
                 "steps": 12, "ess_over_n": lr["ess"], "roofline": lr["roofline"]}
        del lr
        torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = cpu_baseline(wl) if world == 1 and not args.no_cpu else None
    line = {"metric": "particle-steps/s (predict+weight+resample)", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 state / f64 likelihood / u64 scan", "data": "synthetic",
            "config": {"workload": wl["desc"], "particles": n, "landmarks": L, "trials": T, "likelihood": args.likelihood,
                       "noise": "philox4x32-10 on device", "resampling": "forced every step",
                       "ess_over_n": ess_frac,
                       "l2": f"rotating {replicas} independent filter replicas ({replicas * state_bytes >> 20} MiB > 126 MiB L2)",
                       "parallelism": "single GPU" if world == 1 else f"{world} ranks, one independent filter each (no data-path collective)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "large_filter": large, "sharded_filter": sharded,
            "gpu_launches": int(launches), "clocks": clocks}
    if args.likelihood in ("pair", "rb_fan"):
        # the likelihood-bound configurations (C4): what the pair loop sustains.  rb_fan scores only the landmarks
        # the fan sensor detected (NaN rows of the observation are skipped); pair scores all L in fp64.
        seen = float(np.mean([np.isfinite(c[..., 0]).sum(axis=-1).mean() for c in zcyc_host])) if args.likelihood == "rb_fan" else float(L)
        k_ms = roofline["kernel_ms"]["k_step"]
        line["pair_loop"] = {"landmarks_scored_per_step": seen, "pairs_per_s": n * T * seen / (k_ms * 1e-3),
                             "arithmetic": "fp32 + MUFU (rcp, sqrt), branch-free atan2" if args.likelihood == "rb_fan" else "fp64, 8 operations per pair",
                             "bound": "FP32/SFU issue" if args.likelihood == "rb_fan" else "FP64 pipe"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--likelihood", default="moment", choices=["pair", "moment", "rb_fan"])
    ap.add_argument("--n", type=int, default=0, help="override the particle count of the workload (experiments)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the particle-sharded C3 measurement")
    ap.add_argument("--no-large", action="store_true", help="N = 1, c1: skip the extra C3-shard measurement")
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
