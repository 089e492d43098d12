"""One particle filter sharded over the ranks of a torch.distributed group (one process per GPU).

Rank g owns the contiguous global particles / output slots [g * m, (g + 1) * m), m = n_global / G.
Predict + weight is local.  Two exchanges couple the shards:

  * statistics -- every rank allgathers its 312-byte mcl_stats record straight from device memory and merges
    the records on the device (mcl_shard_merge; `merge_stats` below is its host twin, used by the CPU
    tests): global max, normaliser, ESS, arg-max, weighted mean / covariance.  All shards then quantise
    their weights with the same map and take the same resampling decision;
  * resampling -- a first pass (mcl_shard_total) sums the shard's integer weights, tile by tile; the allgather
    of the totals gives every shard the cumulative offset of its first particle; because the weights are
    integers on a 2^-52 grid the global scan is exact, and each shard computes, with GLOBAL slot numbers, the
    resampled pose of every slot its own particles fill (mcl_shard_expand, every scan tile starting from the
    prefixes of the first pass).  Ancestors are monotone, so that is one contiguous slot range per shard
    (`shard_ranges` derives all of them from the totals).  Slots inside a shard's own range keep the local
    ancestor/gather path; only the slots owned by other ranks are materialised as poses and travel in one
    batched group of NCCL point-to-point transfers (`plan_offrank`, `exchange_offrank`).  mcl_shard_step then
    predicts + weights [received prefix | locally gathered middle | received suffix].
    The result is bit-identical to the unsharded filter, and to the reference walk of :212-221.

Two transports carry those exchanges:

  * ``transport="peer"`` (default when CUDA IPC works): DEVICE-DRIVEN.  Every shard owns a mailbox and a receive
    staging area its peers can write (cudaMalloc + CUDA IPC, mapped once at construction).  The kernels that produce
    the statistics record, the weight total and the off-rank poses write them straight into the peers' memory over
    NVLink and raise a step-numbered flag; three single-CTA kernels wait on the own mailbox and derive gate, comb
    offset, k_offset and slot ranges on the device.  A step is one C call that only enqueues kernels
    (mcl_shard_step_peer): no NCCL, no host synchronisation; the merged statistics are read back asynchronously
    and may be consumed one step late (``lag=1``).
  * ``transport="nccl"``: the host-driven step of round 1 (allgather of records and totals, host planning, batched
    NCCL point-to-point), kept as the fallback and as the cross-check of the peer path.  It synchronises with the
    host once per step; when the statistics say the next step will resample, its expansion is started right away.

The pure-host pieces (`merge_stats`, `plan_exchange`, `exchange`, `philox_ofs`) import neither CUDA
nor the C library and are what the world_size-2 gloo tests exercise on CPU.
"""
from __future__ import annotations

import math

import numpy as np

_TWO_PI = 2.0 * math.pi


# ---------------------------------------------------------------------------------------------
# host logic
# ---------------------------------------------------------------------------------------------
def limit_angle(a):
    """mylib/limit.py:11-26."""
    r = abs(a)
    while r > math.pi:
        r -= _TWO_PI
    return -r if a < 0 else r


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al. 2011) on Python ints; the device twin is mcl::philox4x32_10."""
    c = [int(v) & 0xFFFFFFFF for v in counter]
    k = [int(v) & 0xFFFFFFFF for v in key]
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def philox_ofs(seed, step, n_global, trial=0):
    """The comb offset rand() * NP_RECIP (particle_filter.py:214) the device draws when no offset is
    injected (mcl::comb_offset): every rank computes the same value, and it equals the unsharded one."""
    r = philox4x32_10((0xFFFFFFFF, 0, step, trial), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u = (float(r[0]) * 4294967296.0 + float(r[1])) * (1.0 / 18446744073709551616.0)
    return u * (1.0 / n_global)


def merge_stats(records, slot_bases):
    """Merge per-shard statistics (dicts as returned by engine.stats_to_dict) into the global record.

    Shards share the reference pose of the moment sums (it derives from the previous GLOBAL estimate),
    so their un-normalised sums add after rescaling to the global maximum."""
    M = max(r["max_logw"] for r in records)
    scale = [math.exp(r["max_logw"] - M) if r["sum_w"] > 0 else 0.0 for r in records]
    s1 = sum(r["sum_w"] * s for r, s in zip(records, scale))
    s2 = sum(r["sum_w2"] * s * s for r, s in zip(records, scale))
    ref = np.asarray(records[0]["ref"], dtype=np.float64)
    raw = np.zeros(12)
    for r, s in zip(records, scale):
        a = np.asarray(r["raw"], dtype=np.float64) * s
        # re-centre the sums of a shard that used another reference pose (only right after
        # set_local_particles; in a running filter all shards derive it from the global estimate)
        d = np.asarray(r["ref"], dtype=np.float64) - ref
        d[2] = limit_angle(d[2])
        if np.any(d != 0.0):
            e = r["sum_w"] * s
            fx, fy, ft = a[0], a[1], a[2]
            a[5] += 2 * d[0] * fx + d[0] * d[0] * e
            a[6] += d[0] * fy + d[1] * fx + d[0] * d[1] * e
            a[7] += 2 * d[1] * fy + d[1] * d[1] * e
            a[8] += d[0] * ft + d[2] * fx + d[0] * d[2] * e
            a[9] += d[1] * ft + d[2] * fy + d[1] * d[2] * e
            a[10] += 2 * d[2] * ft + d[2] * d[2] * e
            a[0] += d[0] * e
            a[1] += d[1] * e
            a[2] += d[2] * e
        raw += a
    # arg-max: largest stored log-weight; ties go to the lowest global index (np.argmax, :116)
    best = None
    for g, r in enumerate(records):
        v = r["best_logw"]
        if best is None or v > best[0]:
            best = (v, g)
    bv, bg = best
    inv = 1.0 / s1
    mdx, mdy, mdt = raw[0] * inv, raw[1] * inv, raw[2] * inv
    mth = math.atan2(raw[3], raw[4])
    dl = limit_angle(mth - ref[2])
    cov = np.empty((3, 3))
    cov[0, 0] = raw[5] * inv - mdx * mdx
    cov[0, 1] = cov[1, 0] = raw[6] * inv - mdx * mdy
    cov[1, 1] = raw[7] * inv - mdy * mdy
    cov[0, 2] = cov[2, 0] = raw[8] * inv - mdx * mdt
    cov[1, 2] = cov[2, 1] = raw[9] * inv - mdy * mdt
    cov[2, 2] = raw[10] * inv - 2.0 * dl * mdt + dl * dl
    raw[11] = bv  # stored log-weight of the MAP particle
    return dict(max_logw=M, sum_w=s1, sum_w2=s2, log_norm=M + math.log(s1), ess=s1 * s1 / s2,
                max_weight=math.exp(bv - M) / s1, max_idx=int(slot_bases[bg] + records[bg]["max_idx"]),
                est=np.asarray(records[bg]["est"], dtype=np.float64), mean=np.array([ref[0] + mdx, ref[1] + mdy, mth]),
                cov=cov, raw=raw, ref=ref, owner=bg)


def count_le(K, n_global, ofs):
    """Number of comb teeth u_i = fl(fl(i * (1/N)) + ofs) <= K * 2^-52 (host twin of mcl::count_le; plain
    Python floats are IEEE doubles and CPython never fuses the multiply and the add)."""
    nd = float(n_global)
    delta = 1.0 / nd
    c = float(K) * 2.0 ** -52
    g = float(math.floor((c - ofs) * nd))

    def comb(i):
        m = i * delta
        return m + ofs

    r = g
    if comb(g + 1.0) <= c:
        r = g + 1.0
    elif comb(g) > c:
        r = g - 1.0
    return int(min(max(r + 1.0, 0.0), nd))


def shard_ranges(totals, n_global, ofs):
    """(first_slot, n_slots) of every shard from the allgathered integer weight totals: shard g fills the
    slots whose tooth falls in (K_g, K_g + total_g]; the last shard also takes any slots past the final
    cumulative weight.  Every rank computes the same list, so no second exchange is needed."""
    out, k = [], 0
    for g, t in enumerate(totals):
        first = count_le(k, n_global, ofs)
        k += int(t)
        end = n_global if g == len(totals) - 1 else count_le(k, n_global, ofs)
        out.append((first, end - first))
    return out


def plan_offrank(ranges, n_global, world, rank):
    """Split sizes when a shard keeps the slots of its own range and only off-rank slots travel.
    Returns (send, recv, n_prefix, n_suffix): send[g]/recv[s] as plan_exchange with the diagonal zeroed;
    the received buffer is [prefix | suffix] of my output range."""
    send, recv = plan_exchange(ranges, n_global, world, rank)
    send[rank] = 0
    recv[rank] = 0
    return send, recv, sum(recv[:rank]), sum(recv[rank + 1:])


def plan_exchange(ranges, n_global, world, rank):
    """Split sizes of the all-to-all.  ranges[s] = (first_slot, n_slots) of the slots shard s fills.

    Returns (send_counts, recv_counts): send_counts[g] slots of MY expanded range go to owner g, in slot
    order; recv_counts[s] slots arrive from shard s -- concatenated in rank order they are exactly my
    m = n_global / world output slots, because the shards' ranges tile [0, n_global) in rank order."""
    m = n_global // world
    first, count = ranges[rank]
    send = [max(0, min(first + count, (g + 1) * m) - max(first, g * m)) for g in range(world)]
    lo, hi = rank * m, (rank + 1) * m
    recv = [max(0, min(f + c, hi) - max(f, lo)) for f, c in ranges]
    if sum(recv) != m:
        raise RuntimeError(f"rank {rank}: shards fill {sum(recv)} of my {m} slots -- ranges {ranges} do not tile the comb")
    return send, recv


def exchange(arrays, first, count, n_global, group=None):
    """Variable-count all-to-all of the expanded poses.  `arrays`: tensors whose first `count` elements
    are the resampled values of global slots [first, first + count).  Returns my m output slots of each."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = arrays[0].device
    mine = torch.tensor([first, count], dtype=torch.int64, device=dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    ranges = [tuple(int(v) for v in t.tolist()) for t in gathered]
    send, recv = plan_exchange(ranges, n_global, world, rank)
    m = n_global // world
    out = []
    for a in arrays:
        o = torch.empty(m, dtype=a.dtype, device=dev)
        dist.all_to_all_single(o, a[:count].contiguous(), recv, send, group=group)
        out.append(o)
    return out, ranges


def exchange_offrank(send_arrays, ranges, n_global, group=None):
    """All-to-all of the off-rank slots only.  `send_arrays`: [slots below my range | slots above it] as
    written by mcl_shard_expand.  Returns (recv_arrays, n_prefix, n_suffix).

    All arrays travel in ONE batched group of point-to-point transfers (one NCCL launch), and only between
    the pairs of ranks that actually exchange slots -- in a running filter those are the neighbours."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    send, recv, n_prefix, n_suffix = plan_offrank(ranges, n_global, world, rank)
    n_recv = sum(recv)
    out = [torch.empty(max(n_recv, 1), dtype=a.dtype, device=a.device) for a in send_arrays]
    ops, so, ro = [], 0, 0
    for g in range(world):
        peer = g if group is None else dist.get_global_rank(group, g)
        for a, o in zip(send_arrays, out):  # the same order on both sides of a pair: matched one by one
            if send[g]:
                ops.append(dist.P2POp(dist.isend, a[so:so + send[g]], peer, group))
            if recv[g]:
                ops.append(dist.P2POp(dist.irecv, o[ro:ro + recv[g]], peer, group))
        so += send[g]
        ro += recv[g]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out, n_prefix, n_suffix


# ---------------------------------------------------------------------------------------------
# the sharded filter (needs CUDA + libmcl_b200.so)
# ---------------------------------------------------------------------------------------------
class ShardedParticleFilter:
    """n_global particles, contiguous shards of n_global / world_size per rank."""

    def __init__(self, n_global, landmarks, period_ms=100.0, likelihood="moment", seed=42, group=None, device=None,
                 ess_div=100.0, expand_factor=0.125, speculate=True, transport="auto", lag=0, **cfg):
        import torch
        import torch.distributed as dist

        from . import _lib as L
        from .engine import Filter, stats_to_dict

        self._L, self._stats_to_dict = L, stats_to_dict
        self.dist, self.torch, self.group = dist, torch, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if n_global % self.world or (n_global // self.world) % 4:
            raise ValueError("n_global must be a multiple of 4 * world_size")
        self.n_global, self.m = int(n_global), int(n_global) // self.world
        self.seed = int(seed)
        self.ess_threshold = self.n_global / ess_div  # ESS_TH = NP / 100 (particle_filter.py:33)
        dev_index = torch.cuda.current_device() if device is None else int(device)
        self.flt = Filter(n_particles=self.m, n_trials=1, period_ms=period_ms, landmarks=landmarks, likelihood=likelihood,
                          device=dev_index, seed=seed, n_global=self.n_global, slot_base=self.rank * self.m,
                          ess_threshold=self.ess_threshold, **cfg)
        self.dev = self.flt.device
        self.slot_bases = [g * self.m for g in range(self.world)]
        # only slots owned by OTHER ranks are materialised as poses: the send buffers hold `cap` slots and
        # grow when a step needs more (worst case n_global - m, when one shard holds all the weight)
        self.cap = min(self.n_global, int(self.m * expand_factor) + 1024)
        self.ex = None  # allocated by the first host-driven expansion (the peer transport never needs them)
        self.total_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._totals_buf = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
        self._totals_host = torch.zeros(self.world, dtype=torch.int64, pin_memory=True)
        self.speculate = bool(speculate)
        self._spec = None
        self._force_last = False
        self._gather_buf = None
        self._totals = None
        self._ofs_ahead = None
        self._stats_dtype = np.dtype(L.mcl_stats)
        self.step_count = 0
        self.global_stats = None
        self.last_ranges = None
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError("transport must be 'auto', 'peer' or 'nccl'")
        self.lag = int(lag)
        self.transport, self.transport_note = self._connect_peer_memory(transport)
        if self.transport == "nccl":
            self._connect_peers()

    def _connect_peer_memory(self, want):
        """Map every shard's peer-visible region into this process (CUDA IPC).  All ranks must agree on the
        outcome, so the success flags are all-reduced; 'auto' falls back to NCCL when any rank fails."""
        t, dist = self.torch, self.dist
        if want == "nccl":
            return "nccl", "requested"
        ok, why = 1, ""
        try:
            handle = self.flt.peer_alloc(self.rank, self.world)
        except Exception as e:  # noqa: BLE001
            ok, why, handle = 0, f"peer_alloc: {e}", bytes(self._L.MCL_IPC_HANDLE_BYTES)
        mine = t.tensor(list(handle), dtype=t.uint8, device=self.dev)
        allh = t.empty(self.world * mine.numel(), dtype=t.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        blob = bytes(allh.cpu().tolist())
        handles = [blob[g * len(handle):(g + 1) * len(handle)] for g in range(self.world)]
        if ok:
            try:
                self.flt.peer_connect_ipc(handles)
            except Exception as e:  # noqa: BLE001
                ok, why = 0, f"peer_connect: {e}"
        flag = t.tensor([ok], dtype=t.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            return "peer", "CUDA IPC peer mappings"
        if want == "peer":
            raise RuntimeError(f"transport='peer' is not available on rank {self.rank}: {why or 'another rank failed'}")
        return "nccl", f"fallback: {why or 'another rank could not map peer memory'}"

    def _connect_peers(self):
        """NCCL sets up a point-to-point connection the first time a pair of ranks exchanges data (milliseconds).
        Which pairs exchange slots depends on the weights, so every pair is connected once, here."""
        t, dist = self.torch, self.dist
        if self.world == 1:
            return
        sbuf = t.zeros(self.world, dtype=t.float32, device=self.dev)
        rbuf = t.empty(self.world, dtype=t.float32, device=self.dev)
        ops = []
        for g in range(self.world):
            if g == self.rank:
                continue
            peer = g if self.group is None else dist.get_global_rank(self.group, g)
            ops.append(dist.P2POp(dist.isend, sbuf[g:g + 1], peer, self.group))
            ops.append(dist.P2POp(dist.irecv, rbuf[g:g + 1], peer, self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        t.cuda.current_stream().synchronize()

    def _alloc_expand(self):
        t = self.torch
        self.ex = [t.empty(self.cap, dtype=t.float32, device=self.dev) for _ in range(3)]

    # -- statistics ---------------------------------------------------------------------------
    def _uniform_stats(self, est):
        s = self._L.mcl_stats()
        s.log_norm = math.log(self.n_global)
        s.ess = float(self.n_global)
        s.max_weight = 1.0 / self.n_global
        s.max_logw = 0.0
        s.sum_w = float(self.n_global)
        s.sum_w2 = float(self.n_global)
        s.derived = 1
        for k in range(3):
            s.est[k] = float(est[k])
            s.mean[k] = float(est[k])
            s.ref[k] = float(est[k])
        return s

    def _merge(self):
        """Statistics of the state just written, with no host round trip in between: allgather of the shards'
        mcl_stats straight from device memory -> merge on the device (mcl_shard_merge, twin of merge_stats) ->
        this shard's total of the quantised weights under the merged weight map -> allgather of the totals.
        The totals are what the NEXT step's resampling starts from; computing them here (0.1 ms, wasted only
        when that step does not resample) leaves one host synchronisation per step -- the read-back below."""
        t = self.torch
        mine = self.flt.stats_dev_bytes()  # uint8 view of this shard's record (runs the pending merge kernel)
        if self._gather_buf is None:
            self._gather_buf = t.empty(self.world * mine.numel(), dtype=t.uint8, device=self.dev)
        self.dist.all_gather_into_tensor(self._gather_buf, mine, group=self.group)
        self.flt.shard_merge(self._gather_buf, self.world)
        self._enqueue_totals()
        merged = self._stats_to_dict(self.flt.stats()[0])  # device-to-host + synchronise (the totals ride along)
        self._totals = [int(v) for v in self._totals_host.tolist()]
        self.global_stats = merged
        self._spec = None
        if self.speculate and (self._force_last or merged["ess"] < self.ess_threshold) and self._ofs_ahead is not None \
                and self._ofs_ahead[0] == self.step_count:
            # The next step will resample (its gate is this ESS), and everything its resampling needs is known:
            # start it now, so that the host work between two steps hides behind the scan instead of idling the GPU.
            ofs = self._ofs_ahead[1]
            self._spec = (self.step_count, ofs) + self._expand(ofs)
        return merged

    def _enqueue_totals(self):
        self.flt.shard_total(self.total_dev)
        self.dist.all_gather_into_tensor(self._totals_buf, self.total_dev, group=self.group)
        self._totals_host.copy_(self._totals_buf, non_blocking=True)  # pinned: complete at the next stream sync

    # -- API ------------------------------------------------------------------------------------
    def init_particles(self, pose):
        pose = np.asarray(pose, dtype=np.float64).reshape(3)
        self.flt.init_particles(pose)
        self._pending = 0
        if self.transport == "peer":
            # the freshly initialised statistics already describe the GLOBAL uniform filter (ESS = n_global)
            self.global_stats = dict(ess=float(self.n_global), est=pose.copy(), log_norm=math.log(self.n_global))
            return
        self.flt.shard_set_global(self._uniform_stats(pose))
        self._enqueue_totals()
        self.torch.cuda.current_stream().synchronize()
        self._totals = [int(v) for v in self._totals_host.tolist()]
        self._spec = None
        self.global_stats = dict(ess=float(self.n_global), est=pose.copy(), log_norm=math.log(self.n_global))

    def set_local_particles(self, x, y, yaw, logw=None):
        """This rank's m particles (numpy or CUDA tensors); all ranks then merge their statistics."""
        self.flt.set_particles(x, y, yaw, logw)
        if self.transport == "peer":
            self.flt.sync_peer()
            self.global_stats = self._stats_to_dict(self.flt.stats()[0])
            return self.global_stats
        return self._merge()

    def _expand(self, ofs):
        """Global systematic resampling (particle_filter.py:212-222): local slots keep the ancestor path,
        off-rank slots are exchanged as poses.  Returns (recv arrays, n_prefix, n_suffix, ranges)."""
        totals = self._totals  # allgathered at the end of the previous step (_merge)
        ranges = shard_ranges(totals, self.n_global, ofs)
        send, _, _, _ = plan_offrank(ranges, self.n_global, self.world, self.rank)
        if sum(send) > self.cap:
            self.cap = min(self.n_global, max(2 * self.cap, sum(send) + 1024))
            self.ex = None
        if self.ex is None:
            self._alloc_expand()
        self.flt.shard_expand(sum(totals[:self.rank]), ofs, *self.ex, want_range=False)
        recv, n_prefix, n_suffix = exchange_offrank(self.ex, ranges, self.n_global, self.group)
        self.last_ranges = ranges
        return recv, n_prefix, n_suffix

    def step(self, z, ofs=None, noise=None, force_resample=False):
        """main_pf minus the ground-truth line: resample (gated on the global ESS) -> predict -> weight ->
        merged statistics.  `noise`: this rank's (3, 1, m) slice of the draw, or None for Philox."""
        z = np.asarray(z, dtype=np.float64).reshape(1, -1, 2)
        if self.transport == "peer":
            return self._step_peer(z, ofs, noise, force_resample)
        resampled = force_resample or (self.global_stats["ess"] < self.ess_threshold)
        if resampled:
            if ofs is None:
                ahead = self._ofs_ahead
                ofs = ahead[1] if ahead is not None and ahead[0] == self.step_count else philox_ofs(self.seed, self.step_count, self.n_global)
            spec = self._spec
            if spec is not None and spec[0] == self.step_count and spec[1] == ofs:
                recv, n_prefix, n_suffix = spec[2:]  # started at the end of the previous step
            else:
                recv, n_prefix, n_suffix = self._expand(ofs)
            self.flt.shard_step(z, recv[0], recv[1], recv[2], n_prefix, n_suffix, noise=noise)
            self._keep = recv
        else:
            self.flt.predict_weight(z, noise=noise)
        self._spec = None
        self._force_last = bool(force_resample)
        self.step_count += 1
        # host work of the next step that does not need this step's results, done while the GPU is busy
        self._ofs_ahead = (self.step_count, philox_ofs(self.seed, self.step_count, self.n_global))
        merged = self._merge()
        merged["resampled"] = int(resampled)
        return merged

    def _step_peer(self, z, ofs, noise, force_resample):
        """One C call enqueues the whole step; the gate is taken on the device.  The merged statistics come back
        through a two-deep ring of pinned buffers: with lag = 1 the record returned is the one of the PREVIOUS step
        (None after the first), and the host never waits for the step it has just enqueued."""
        self.flt.step_peer(z, ofs=ofs, noise=noise, force_resample=force_resample)
        self.flt.stats_begin()
        self._pending = getattr(self, "_pending", 0) + 1
        self.step_count += 1
        merged = None
        while self._pending > self.lag:
            merged = self._stats_to_dict(self.flt.stats_end()[0])
            self._pending -= 1
        if merged is not None:
            self.global_stats = merged
        return merged

    def drain(self):
        """Wait for the statistics of every enqueued step; -> the newest record."""
        merged = None
        while getattr(self, "_pending", 0) > 0:
            merged = self._stats_to_dict(self.flt.stats_end()[0])
            self._pending -= 1
        if merged is not None:
            self.global_stats = merged
        return self.global_stats

    def plan(self):
        """Device-side plan of the last peer step (synchronises): slot range, window, gate, error flag."""
        pl = self.flt.shard_plan()
        if pl.error:
            raise RuntimeError(f"rank {self.rank}: a wait on a peer shard timed out (plan.error)")
        return pl

    def local_particles(self, device=False):
        x, y, th, lw = self.flt.get_particles(device=device)
        return x[0], y[0], th[0], lw[0]

    # -- checkpoint: one file per rank -------------------------------------------------------------
    def save(self, prefix):
        """Every rank writes `<prefix>.rank<g>.npz`: its shard (Filter.save: particles, log-weights, the GLOBAL
        statistics block its handle holds, step counter, config).  Call on all ranks, between steps."""
        self.drain()
        self.flt.save(f"{prefix}.rank{self.rank}.npz")

    def restore(self, prefix):
        """Load the shard written by save() into this (already constructed and connected) filter: the particles, the
        global statistics and the step counter.  The filter then continues bit-identically to the one that was saved."""
        with np.load(f"{prefix}.rank{self.rank}.npz") as z:
            cfg = self._L.mcl_config.from_buffer_copy(z["config"].tobytes())
            if (cfg.n_global, cfg.slot_base, cfg.n_particles) != (self.n_global, self.rank * self.m, self.m):
                raise ValueError(f"{prefix}: checkpoint of another sharding ({cfg.n_global} particles, base {cfg.slot_base})")
            self.flt.set_particles(z["x"], z["y"], z["yaw"], z["logw"], skip_stats=True)
            st = self._L.mcl_stats.from_buffer_copy(z["stats"].tobytes())
            self.flt.restore_stats(st, int(z["step"]))
            self.step_count = int(z["step"])
        self.global_stats = self._stats_to_dict(self.flt.stats()[0])
        self._pending = 0
        self._spec = None
        if self.transport == "nccl":
            self._enqueue_totals()
            self.torch.cuda.current_stream().synchronize()
            self._totals = [int(v) for v in self._totals_host.tolist()]
        return self.global_stats


class LocalShardedFilter:
    """The same sharded filter driven by ONE process: shard g is a Filter on devices[g] (several GPUs of the box,
    or -- as the tests do -- several shards on one GPU).  The peers' regions are plain pointers (peer access is
    enabled between devices); each step runs the four stages of mcl_shard_step_peer for all shards in turn, so every
    "push" is enqueued before the matching "wait" and a single stream never spins on itself."""

    def __init__(self, n_global, landmarks, n_shards, devices=None, period_ms=100.0, likelihood="moment", seed=42,
                 ess_div=100.0, **cfg):
        from . import _lib as L
        from .engine import Filter, stats_to_dict

        self._L, self._stats_to_dict = L, stats_to_dict
        if n_global % n_shards or (n_global // n_shards) % 4:
            raise ValueError("n_global must be a multiple of 4 * n_shards")
        self.G, self.n_global, self.m = int(n_shards), int(n_global), int(n_global) // int(n_shards)
        devices = [0] * self.G if devices is None else [int(d) for d in devices]
        self.shards = [Filter(n_particles=self.m, n_trials=1, period_ms=period_ms, landmarks=landmarks, likelihood=likelihood,
                              device=devices[g], seed=seed, n_global=self.n_global, slot_base=g * self.m,
                              ess_threshold=self.n_global / ess_div, **cfg) for g in range(self.G)]
        for g, f in enumerate(self.shards):
            f.peer_alloc(g, self.G)
        for f in self.shards:
            f.peer_connect_local(self.shards)

    def init_particles(self, pose):
        for f in self.shards:
            f.init_particles(pose)

    def set_particles(self, x, y, yaw, logw=None):
        """Global arrays of n_global elements; shard g takes [g m, (g + 1) m)."""
        for g, f in enumerate(self.shards):
            sl = slice(g * self.m, (g + 1) * self.m)
            f.set_particles(x[sl], y[sl], yaw[sl], None if logw is None else logw[sl])
        for stage in (self._L.MCL_STAGE_C, self._L.MCL_STAGE_D):
            for f in self.shards:
                f.sync_peer(stage)

    def step(self, z, ofs=None, noise=None, force_resample=False, no_resample=False):
        """noise: (3, 1, n_global) or None (Philox).  -> the merged statistics (of shard 0; all shards agree)."""
        L = self._L
        for stage in (L.MCL_STAGE_A, L.MCL_STAGE_B, L.MCL_STAGE_C, L.MCL_STAGE_D):
            for g, f in enumerate(self.shards):
                nz = None if noise is None else noise[:, :, g * self.m:(g + 1) * self.m]
                f.step_peer(z, ofs=ofs, noise=nz if stage == L.MCL_STAGE_C else None, force_resample=force_resample,
                            no_resample=no_resample, stages=stage)
        return self.stats()

    def stats(self, shard=0):
        return self._stats_to_dict(self.shards[shard].stats()[0])

    def save(self, prefix):
        """One file per shard, `<prefix>.rank<g>.npz` (the layout ShardedParticleFilter.save writes)."""
        for g, f in enumerate(self.shards):
            f.save(f"{prefix}.rank{g}.npz")

    def restore(self, prefix):
        for g, f in enumerate(self.shards):
            with np.load(f"{prefix}.rank{g}.npz") as z:
                f.set_particles(z["x"], z["y"], z["yaw"], z["logw"], skip_stats=True)
                f.restore_stats(self._L.mcl_stats.from_buffer_copy(z["stats"].tobytes()), int(z["step"]))

    def particles(self):
        parts = [f.get_particles() for f in self.shards]
        return tuple(np.concatenate([p[k][0] for p in parts]) for k in range(4))
