"""`Filter`: a thin object over one mcl_filter handle (include/mcl.h).

torch is used for exactly three things here: allocating the device workspace / noise / result
buffers, naming the CUDA stream, and (in sharded.py) torch.distributed.  All filter arithmetic
runs in libmcl_b200.so; nothing in this file computes on particles.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L


def _torch():
    import torch  # deferred: importing torch costs seconds and the ctypes layer does not need it

    return torch


def default_config(period_ms=100.0):
    """The literals of ParticleFilter.__init__ (particle_filter.py:30-84) as an mcl_config."""
    cfg = L.mcl_config()
    L.check(L.load().mcl_config_default(C.byref(cfg), float(period_ms)), "mcl_config_default")
    return cfg


class Filter:
    """One handle = `n_trials` independent filters of `n_particles` particles on one GPU."""

    def __init__(self, n_particles=1000, n_trials=1, period_ms=100.0, landmarks=None, likelihood="pair",
                 motion="a", device=0, seed=42, ess_threshold=None, q_std=None, sigma_obs=None,
                 max_landmarks=None, n_global=None, slot_base=0, motion_b_alpha=None, vel=None, yaw_rate=None,
                 dt=None, fan_range=None, fan_half_angle=None, fan_miss_chi2=None, underflow_rule=True):
        torch = _torch()
        self.lib = L.load()
        cfg = default_config(period_ms)
        cfg.device = int(device)
        cfg.n_particles = int(n_particles)
        cfg.n_trials = int(n_trials)
        cfg.n_global = int(n_global if n_global is not None else n_particles)
        cfg.slot_base = int(slot_base)
        cfg.seed = int(seed)
        cfg.likelihood = {"pair": L.MCL_LIK_PAIR, "moment": L.MCL_LIK_MOMENT, "rb": L.MCL_LIK_RB,
                          "rb_fan": L.MCL_LIK_RB_FAN, "pair32": L.MCL_LIK_PAIR_F32, "rb32": L.MCL_LIK_RB_F32}[likelihood]
        if likelihood == "rb_fan" and sigma_obs is None:
            # ScanSensor noise as the reference's Robot sets it (graph_based_slam.py:604): 5 % of the distance, 2 deg
            sigma_obs = [0.05, np.deg2rad(2.0)]
        for name, val in (("fan_range", fan_range), ("fan_half_angle", fan_half_angle), ("fan_miss_chi2", fan_miss_chi2)):
            if val is not None:
                setattr(cfg, name, float(val))
        cfg.underflow_rule = 1 if underflow_rule else 0  # particle_filter.py:234-237 (NaN weights -> 1/NP)
        cfg.motion = {"a": L.MCL_MOTION_A, "b": L.MCL_MOTION_B}[motion]
        # ESS_TH = NP / 100 (particle_filter.py:33)
        cfg.ess_threshold = float(ess_threshold) if ess_threshold is not None else cfg.n_global / 100.0
        if q_std is not None:
            cfg.q_std[:] = [float(v) for v in q_std]
        if sigma_obs is not None:
            cfg.sigma_obs[:] = [float(v) for v in sigma_obs]
        if motion_b_alpha is not None:
            cfg.motion_b_alpha[:] = [float(v) for v in motion_b_alpha]
        if vel is not None:
            cfg.vel = float(vel)
        if yaw_rate is not None:
            cfg.yaw_rate = float(yaw_rate)
        if dt is not None:
            cfg.dt = float(dt)
        lm = None if landmarks is None else np.ascontiguousarray(landmarks, dtype=np.float64).reshape(-1, 2)
        cfg.max_landmarks = int(max_landmarks if max_landmarks is not None else max(1, 0 if lm is None else len(lm), 8))
        self._create(cfg, lm)

    def _create(self, cfg, lm):
        torch = _torch()
        self.lib = L.load()
        self.cfg = cfg
        self.n = cfg.n_particles
        self.T = cfg.n_trials
        self.device = torch.device("cuda", cfg.device)
        need = C.c_size_t(0)
        L.check(self.lib.mcl_workspace_bytes(C.byref(cfg), C.byref(need)), "mcl_workspace_bytes")
        if not torch.cuda.is_available():
            raise L.MclError(L.MCL_ERR_NODEVICE, "Filter", "no CUDA device: the engine has no CPU fallback")
        self.workspace = torch.empty(need.value + 256, dtype=torch.uint8, device=self.device)
        base = (self.workspace.data_ptr() + 255) // 256 * 256
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.mcl_create(C.byref(cfg), C.c_void_p(base), need.value, C.byref(self._h)), "mcl_create")
        self.n_landmarks = 0
        if lm is not None:
            self.set_landmarks(lm)

    # ------------------------------------------------------------------ checkpoint
    def save(self, path):
        """Everything a bit-identical continuation needs, as one .npz: config, landmarks, particles with their
        log-weights, the statistics block (weight map, gate, MAP pose) and the step counter."""
        x, y, th, lw = self.get_particles()
        st = (L.mcl_stats * self.T)(*self.stats())
        np.savez(path, config=np.frombuffer(bytes(self.cfg), dtype=np.uint8), landmarks=self.landmarks,
                 x=x, y=y, yaw=th, logw=lw, stats=np.frombuffer(bytes(st), dtype=np.uint8), step=np.int64(st[0].step),
                 abi=np.int64(L.MCL_ABI_VERSION))

    @classmethod
    def load(cls, path, device=None):
        with np.load(path) as z:
            if int(z["abi"]) != L.MCL_ABI_VERSION:
                raise ValueError(f"{path}: written with ABI {int(z['abi'])}, this library is ABI {L.MCL_ABI_VERSION}")
            cfg = L.mcl_config.from_buffer_copy(z["config"].tobytes())
            if device is not None:
                cfg.device = int(device)
            self = cls.__new__(cls)
            self._create(cfg, np.ascontiguousarray(z["landmarks"], dtype=np.float64))
            self.set_particles(z["x"], z["y"], z["yaw"], z["logw"], skip_stats=True)
            st = (L.mcl_stats * self.T).from_buffer_copy(z["stats"].tobytes())
            self._check(self.lib.mcl_restore(self._h, st, int(z["step"]), self._stream()), "mcl_restore")
        return self

    def restore_stats(self, stats, step):
        """Install a statistics block (list of mcl_stats, or one) and the step counter: mcl_restore."""
        if isinstance(stats, L.mcl_stats):
            stats = [stats]
        st = (L.mcl_stats * self.T)(*stats)
        self._check(self.lib.mcl_restore(self._h, st, int(step), self._stream()), "mcl_restore")

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _check(self, status, where):
        L.check(status, where, self._h)

    _version = 0  # bumped by every call that changes the particles (LazyParticles refuses to serve an older state)

    def _changed(self):
        self._version += 1

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mcl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _f64(a, shape):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.size != math.prod(shape):
            raise ValueError(f"expected {shape}, got {a.shape}")
        return a

    # ------------------------------------------------------------------ setup
    def set_landmarks(self, lm_xy):
        lm = np.ascontiguousarray(lm_xy, dtype=np.float64).reshape(-1, 2)
        self._check(self.lib.mcl_set_landmarks(self._h, lm.ctypes.data, lm.shape[0], self._stream()), "mcl_set_landmarks")
        self.n_landmarks = lm.shape[0]
        self.landmarks = lm

    def init_particles(self, pose):
        """px = repmat(x_true, 1, NP), pw = 1/NP (particle_filter.py:81-84); pose (3,) or (T,3)."""
        p = np.ascontiguousarray(np.broadcast_to(np.asarray(pose, dtype=np.float64).reshape(-1, 3), (self.T, 3)))
        self._check(self.lib.mcl_init_particles(self._h, p.ctypes.data, self._stream()), "mcl_init_particles")
        self._changed()

    def set_particles(self, x, y, yaw, logw=None, skip_stats=False):
        """Arrays (T,n)/(n,) -- numpy (host) or torch CUDA tensors (device), fp32."""
        torch = _torch()
        on_dev = isinstance(x, torch.Tensor)
        if on_dev:
            arrs = [None if a is None else a.to(device=self.device, dtype=torch.float32).contiguous() for a in (x, y, yaw, logw)]
            ptrs = [None if a is None else C.c_void_p(a.data_ptr()) for a in arrs]
        else:
            arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, yaw, logw)]
            ptrs = [None if a is None else C.c_void_p(a.ctypes.data) for a in arrs]
        for a in arrs:
            if a is not None and (a.numel() if on_dev else a.size) != self.T * self.n:
                raise ValueError("particle arrays must hold n_trials * n_particles elements")
        mode = (1 if on_dev else 0) | (2 if skip_stats else 0)
        self._check(self.lib.mcl_set_particles(self._h, *ptrs, mode, self._stream()), "mcl_set_particles")
        self._changed()
        self._keep = arrs  # device sources must outlive the async copies

    def get_particles(self, want_logw=True, device=False):
        """-> x, y, yaw, logw as (T,n) fp32 numpy arrays (or CUDA tensors with device=True)."""
        torch = _torch()
        shape = (self.T, self.n)
        if device:
            out = [torch.empty(shape, dtype=torch.float32, device=self.device) for _ in range(4 if want_logw else 3)]
            ptrs = [C.c_void_p(a.data_ptr()) for a in out]
        else:
            out = [np.empty(shape, dtype=np.float32) for _ in range(4 if want_logw else 3)]
            ptrs = [C.c_void_p(a.ctypes.data) for a in out]
        if not want_logw:
            ptrs.append(None)
        self._check(self.lib.mcl_get_particles(self._h, *ptrs, 1 if device else 0, self._stream()), "mcl_get_particles")
        return out if want_logw else out + [None]

    # ------------------------------------------------------------------ the step and its stages
    def _noise_ptr(self, noise):
        if noise is None:
            return None
        torch = _torch()
        if not isinstance(noise, torch.Tensor):
            noise = torch.from_numpy(np.ascontiguousarray(noise, dtype=np.float32)).to(self.device, non_blocking=False)
        noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
        if noise.numel() != 3 * self.T * self.n:
            raise ValueError("noise must be (3, n_trials, n_particles)")
        self._noise = noise
        return C.c_void_p(noise.data_ptr())

    def _flags(self, noise, force, no_resample):
        f = 0
        if noise is None:
            f |= L.MCL_PHILOX_NOISE
        if force:
            f |= L.MCL_FORCE_RESAMPLE
        if no_resample:
            f |= L.MCL_NO_RESAMPLE
        return f

    def resample(self, ofs=None, force=False):
        """__resampling (particle_filter.py:200-224); ofs = rand() * NP_RECIP per trial, None = Philox."""
        flags = (L.MCL_FORCE_RESAMPLE if force else 0) | (L.MCL_PHILOX_NOISE if ofs is None else 0)
        o = None if ofs is None else self._f64(np.broadcast_to(np.asarray(ofs, dtype=np.float64).reshape(-1), (self.T,)), (self.T,))
        self._check(self.lib.mcl_resample(self._h, None if o is None else o.ctypes.data, flags, self._stream()), "mcl_resample")
        self._changed()

    def predict_weight(self, z, noise=None):
        """__predict + __likelihood + __normalize + MAP estimate (particle_filter.py:156-198, :115-117)."""
        zz = self._f64(z, (self.T, self.n_landmarks, 2))
        self._check(self.lib.mcl_predict_weight(self._h, zz.ctypes.data, self._noise_ptr(noise), self._flags(noise, False, True),
                                                self._stream()), "mcl_predict_weight")
        self._changed()

    def step(self, z, ofs=None, noise=None, force_resample=False, no_resample=False):
        """main_pf minus the ground-truth line (particle_filter.py:104-117)."""
        zz = self._f64(z, (self.T, self.n_landmarks, 2))
        o = None if ofs is None else self._f64(np.broadcast_to(np.asarray(ofs, dtype=np.float64).reshape(-1), (self.T,)), (self.T,))
        flags = self._flags(noise, force_resample, no_resample)
        if ofs is None and not no_resample:
            flags |= L.MCL_PHILOX_NOISE
        self._check(self.lib.mcl_step(self._h, zz.ctypes.data, None if o is None else o.ctypes.data, self._noise_ptr(noise),
                                      flags, self._stream()), "mcl_step")
        self._changed()

    def step_async(self, z, ofs=None, noise=None, force_resample=False, no_resample=False):
        """step() + stats_begin() in one crossing of the C ABI (mcl_step_async); collect with stats_end().
        `z` must already be a C-contiguous float64 array of (T, L, 2) elements -- no conversion on this path."""
        flags = self._flags(noise, force_resample, no_resample)
        if ofs is None and not no_resample:
            flags |= L.MCL_PHILOX_NOISE
        o = None if ofs is None else (C.c_double * self.T)(*([float(ofs)] * self.T if np.isscalar(ofs) else [float(v) for v in ofs]))
        self._check(self.lib.mcl_step_async(self._h, z.ctypes.data, o, self._noise_ptr(noise), flags, self._stream_cached()),
                    "mcl_step_async")
        self._changed()

    def _stream_cached(self):
        """The CUDA stream handle of the first call (a step every ~50 us cannot afford torch.cuda.current_stream())."""
        s = getattr(self, "_stream_handle", None)
        if s is None:
            s = self._stream_handle = self._stream()
        return s

    def step_dev(self, z_dev, ofs_dev=None, noise=None, force_resample=False, no_resample=False):
        """Same with z (T,L,2) / ofs (T,) fp64 CUDA tensors: no host traffic, graph-capturable."""
        flags = self._flags(noise, force_resample, no_resample)
        if ofs_dev is None and not no_resample:
            flags |= L.MCL_PHILOX_NOISE
        self._check(self.lib.mcl_step_dev(self._h, C.c_void_p(z_dev.data_ptr()),
                                          None if ofs_dev is None else C.c_void_p(ofs_dev.data_ptr()),
                                          self._noise_ptr(noise), flags, self._stream()), "mcl_step_dev")
        self._changed()

    def pose_moments(self):
        self._check(self.lib.mcl_pose_moments(self._h, self._stream()), "mcl_pose_moments")

    # ------------------------------------------------------------------ results
    def stats(self):
        arr = (L.mcl_stats * self.T)()
        self._check(self.lib.mcl_get_stats(self._h, arr, self._stream()), "mcl_get_stats")
        return list(arr)

    def stats_begin(self):
        """First half of stats(): enqueue the merge and the copy into a pinned buffer of the handle, return at once.
        The host is free until stats_end() -- ParticleFilter.main_pf simulates the next frame's observation there."""
        self._check(self.lib.mcl_get_stats_begin(self._h, self._stream()), "mcl_get_stats_begin")

    def stats_end(self):
        arr = (L.mcl_stats * self.T)()
        self._check(self.lib.mcl_get_stats_end(self._h, arr), "mcl_get_stats_end")
        return list(arr)

    def stats_dev_bytes(self):
        """The [n_trials] mcl_stats block in device memory as a uint8 CUDA tensor (a view of the workspace).
        Sums, ESS, MAP index and pose are current; log_norm / mean / cov are derived on the host by stats()."""
        self._check(self.lib.mcl_sync_stats(self._h, self._stream()), "mcl_sync_stats")
        ptr = C.c_void_p()
        self._check(self.lib.mcl_stats_dev(self._h, C.byref(ptr)), "mcl_stats_dev")
        off = ptr.value - self.workspace.data_ptr()
        return self.workspace[off:off + self.T * C.sizeof(L.mcl_stats)]

    def weights(self):
        w = np.empty((self.T, self.n), dtype=np.float64)
        self._check(self.lib.mcl_get_weights(self._h, w.ctypes.data, self._stream()), "mcl_get_weights")
        return w

    def qweights(self):
        q = np.empty((self.T, self.n), dtype=np.uint64)
        self._check(self.lib.mcl_get_qweights(self._h, q.ctypes.data, self._stream()), "mcl_get_qweights")
        return q

    def ancestors(self):
        a = np.empty((self.T, self.n), dtype=np.uint32)
        self._check(self.lib.mcl_get_ancestors(self._h, a.ctypes.data, self._stream()), "mcl_get_ancestors")
        return a

    # ------------------------------------------------------------------ measurement hooks
    def profile(self, enable=True):
        self._check(self.lib.mcl_profile(self._h, 1 if enable else 0), "mcl_profile")

    def profile_read(self):
        """(prep, scan, step) kernel milliseconds of the last step (synchronises)."""
        ms = (C.c_double * 3)()
        self._check(self.lib.mcl_profile_read(self._h, ms), "mcl_profile_read")
        return ms[0], ms[1], ms[2]

    def profile_read_peer(self):
        """(totals + exchange, scalars + expansion + exchange, 0, k_step, statistics exchange) ms of the last peer step."""
        ms = (C.c_double * 5)()
        self._check(self.lib.mcl_profile_read_peer(self._h, ms), "mcl_profile_read_peer")
        return tuple(ms)

    def debug_timeline(self):
        """Arms (first call) / reads the k_step phase timeline: (n_ctas, 8) uint64 nanosecond stamps."""
        n = C.c_int32(0)
        self._check(self.lib.mcl_debug_timeline(self._h, None, 0, C.byref(n)), "mcl_debug_timeline")
        out = np.zeros((n.value, 12), dtype=np.uint64)
        self._check(self.lib.mcl_debug_timeline(self._h, out.ctypes.data, out.size, C.byref(n)), "mcl_debug_timeline")
        return out

    def launch_count(self):
        return int(self.lib.mcl_launch_count(self._h))

    # ------------------------------------------------------------------ particle-sharded filter
    def shard_total(self, total_dev):
        self._check(self.lib.mcl_shard_total(self._h, C.c_void_p(total_dev.data_ptr()), self._stream()), "mcl_shard_total")

    def shard_expand(self, k_offset, ofs, send_x, send_y, send_yaw, want_range=True):
        """-> (first_slot, n_slots) of the slots this shard fills (None, None with want_range=False: no sync)."""
        cap = min(send_x.numel(), send_y.numel(), send_yaw.numel())
        first, count = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.mcl_shard_expand(self._h, int(k_offset), float(ofs), C.c_void_p(send_x.data_ptr()),
                                              C.c_void_p(send_y.data_ptr()), C.c_void_p(send_yaw.data_ptr()), cap,
                                              C.byref(first) if want_range else None, C.byref(count) if want_range else None,
                                              self._stream()), "mcl_shard_expand")
        return (first.value, count.value) if want_range else (None, None)

    def shard_step(self, z, recv_x, recv_y, recv_yaw, n_prefix, n_suffix, noise=None):
        zz = self._f64(z, (1, self.n_landmarks, 2))
        ptr = lambda a: None if a is None else C.c_void_p(a.data_ptr())  # noqa: E731
        self._check(self.lib.mcl_shard_step(self._h, zz.ctypes.data, self._noise_ptr(noise), ptr(recv_x), ptr(recv_y),
                                            ptr(recv_yaw), int(n_prefix), int(n_suffix), self._flags(noise, False, False),
                                            self._stream()), "mcl_shard_step")
        self._changed()

    def shard_merge(self, records, n_shards):
        """records: CUDA uint8 tensor holding the n_shards allgathered mcl_stats blocks."""
        self._check(self.lib.mcl_shard_merge(self._h, C.c_void_p(records.data_ptr()), int(n_shards), self._stream()),
                    "mcl_shard_merge")

    # ------------------------------------------------------------------ device-driven sharded filter (peer memory)
    def peer_alloc(self, rank, n_shards):
        """Allocate this shard's peer-visible region; -> its 64-byte CUDA IPC handle."""
        buf = (C.c_ubyte * L.MCL_IPC_HANDLE_BYTES)()
        self._check(self.lib.mcl_shard_peer_alloc(self._h, int(rank), int(n_shards), buf), "mcl_shard_peer_alloc")
        return bytes(buf)

    def peer_connect_ipc(self, handles):
        """handles: the n_shards IPC handles in rank order (other processes)."""
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self.lib.mcl_shard_peer_connect(self._h, buf, None), "mcl_shard_peer_connect")

    def peer_connect_local(self, filters):
        """filters: the n_shards Filter objects of this process, in rank order."""
        arr = (C.c_void_p * len(filters))(*[f._h for f in filters])
        self._check(self.lib.mcl_shard_peer_connect(self._h, None, arr), "mcl_shard_peer_connect")

    def step_peer(self, z, ofs=None, noise=None, force_resample=False, no_resample=False, stages=L.MCL_STAGE_ALL):
        zz = self._f64(z, (1, self.n_landmarks, 2))
        o = None if ofs is None else (C.c_double * 1)(float(ofs))
        self._check(self.lib.mcl_shard_step_peer(self._h, zz.ctypes.data, o, self._noise_ptr(noise),
                                                 self._flags(noise, force_resample, no_resample), int(stages), self._stream()),
                    "mcl_shard_step_peer")
        self._changed()

    def sync_peer(self, stages=L.MCL_STAGE_C | L.MCL_STAGE_D):
        self._check(self.lib.mcl_shard_sync_peer(self._h, int(stages), self._stream()), "mcl_shard_sync_peer")

    def shard_plan(self):
        """The plan of the last device-driven step (synchronises)."""
        pl = L.mcl_shard_plan()
        self._check(self.lib.mcl_shard_get_plan(self._h, C.byref(pl), self._stream()), "mcl_shard_get_plan")
        return pl

    def shard_set_global(self, merged: L.mcl_stats):
        self._check(self.lib.mcl_shard_set_global(self._h, C.byref(merged), self._stream()), "mcl_shard_set_global")


def stats_to_dict(s: L.mcl_stats):
    return dict(log_norm=s.log_norm, ess=s.ess, max_weight=s.max_weight, max_idx=int(s.max_idx),
                est=np.array(s.est[:]), mean=np.array(s.mean[:]), cov=np.array(s.cov[:]).reshape(3, 3),
                sum_w=s.sum_w, sum_w2=s.sum_w2, max_logw=s.max_logw, raw=np.array(s.raw[:]), ref=np.array(s.ref[:]),
                resampled=int(s.resampled), step=int(s.step), underflow=int(s.underflow))


def log_uniform(n_global):
    return -math.log(float(n_global))
