"""world_size-2 (or more) gloo worker for tests/test_host_logic.py: the host side of the sharded
resampling -- allgather of ranges, plan_exchange, variable-count all-to-all -- with each rank's device
outputs emulated by the numpy oracle.  Run under torchrun; exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import importlib  # noqa: E402

from oracle import pf_oracle as O  # noqa: E402

sh = importlib.import_module("slam-robot_simu_b200.sharded")
eng = importlib.import_module("slam-robot_simu_b200.engine")


def emulate_shard(q, pose, ofs, n_global, rank, world):
    """What mcl_shard_total / mcl_shard_expand return for this rank, computed with the oracle."""
    m = n_global // world
    anc = O.resample_indices(O.dequantize_weights(q), ofs, clamp=True)  # global ancestors, reference walk
    mine = np.nonzero((anc >= rank * m) & (anc < (rank + 1) * m))[0]
    first = int(mine[0]) if mine.size else int(np.searchsorted(anc, rank * m))
    count = int(mine.size)
    if count:
        assert mine[-1] - mine[0] + 1 == count  # ancestors are monotone: one contiguous slot range
    expanded = [pose[d][anc[first:first + count]].astype(np.float32) for d in range(3)]
    return first, count, expanded, anc


def local_stats(px, lw, ref, base):
    """A shard's mcl_stats as a dict, from numpy (the fields merge_stats reads)."""
    m = float(lw.max())
    e = np.exp(lw - m)
    d = np.stack([px[0] - ref[0], px[1] - ref[1], O.limit_angle_vec(px[2] - ref[2])])
    raw = np.zeros(12)
    raw[0:3] = (e * d).sum(axis=1)
    raw[3], raw[4] = (e * np.sin(px[2])).sum(), (e * np.cos(px[2])).sum()
    raw[5], raw[6], raw[7] = (e * d[0] * d[0]).sum(), (e * d[0] * d[1]).sum(), (e * d[1] * d[1]).sum()
    raw[8], raw[9], raw[10] = (e * d[0] * d[2]).sum(), (e * d[1] * d[2]).sum(), (e * d[2] * d[2]).sum()
    i = int(np.argmax(lw))
    return dict(max_logw=m, best_logw=m, sum_w=float(e.sum()), sum_w2=float((e * e).sum()), raw=raw, ref=np.array(ref),
                max_idx=i, est=px[:, i].copy())


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for case, (n_global, kind) in enumerate([(4096, "random"), (8192, "peaked"), (4096, "one_heavy"), (1024 * world, "left")]):
        rng = np.random.RandomState(100 + case)  # identical on every rank
        m = n_global // world
        if kind == "random":
            lw = rng.normal(0, 1.5, n_global)
        elif kind == "peaked":
            lw = -0.5 * ((np.arange(n_global) - 0.8 * n_global) / 40.0) ** 2
        elif kind == "one_heavy":
            lw = np.full(n_global, -50.0)
            lw[n_global // 3] = 0.0
        else:  # all mass in shard 0
            lw = np.where(np.arange(n_global) < m // 2, 0.0, -80.0) + rng.normal(0, 0.1, n_global)
        w = np.exp(lw - lw.max())
        w /= w.sum()
        q = O.quantize_weights(w)
        pose = rng.normal(0, 1, (3, n_global))
        pose[2] *= 0.2  # a yaw cloud well inside (-pi, pi): the single-pass moments wrap about the reference pose
        ofs = rng.rand() / n_global
        first, count, expanded, anc = emulate_shard(q, pose, ofs, n_global, rank, world)
        tens = [torch.from_numpy(np.concatenate([e, np.zeros(7, np.float32)])) for e in expanded]  # capacity > count
        out, ranges = sh.exchange(tens, first, count, n_global)
        want = pose[:, anc][:, rank * m:(rank + 1) * m].astype(np.float32)
        for d in range(3):
            if not np.array_equal(out[d].numpy(), want[d]):
                print(f"rank {rank} case {kind}: exchange mismatch in row {d}", flush=True)
                sys.exit(1)
        # the same ranges follow from the allgathered integer totals alone
        totals = [int(q[g * m:(g + 1) * m].sum()) for g in range(world)]
        derived = sh.shard_ranges(totals, n_global, ofs)
        for (f_a, c_a), (f_b, c_b) in zip(ranges, derived):
            assert c_a == c_b and (c_a == 0 or f_a == f_b), (ranges, derived)
        # off-rank exchange: a shard keeps the slots of its own range and sends [below | above]
        lo, hi = rank * m, (rank + 1) * m
        sel = np.concatenate([np.arange(first, min(lo, first + count)), np.arange(max(hi, first), first + count)]).astype(np.int64)
        send = [torch.from_numpy(np.concatenate([pose[d][anc[sel]].astype(np.float32), np.zeros(3, np.float32)])) for d in range(3)]
        recv, n_pre, n_suf = sh.exchange_offrank(send, derived, n_global)
        own = np.arange(max(lo, first), min(hi, first + count))
        assert n_pre + own.size + n_suf == m
        for d in range(3):
            full = np.concatenate([recv[d].numpy()[:n_pre], pose[d][anc[own]].astype(np.float32), recv[d].numpy()[n_pre:n_pre + n_suf]])
            if not np.array_equal(full, want[d]):
                print(f"rank {rank} case {kind}: off-rank exchange mismatch in row {d}", flush=True)
                sys.exit(1)
        # the ranges tile the comb in rank order
        assert ranges[0][0] == 0 and sum(c for _, c in ranges) == n_global
        for (f0, c0), (f1, _) in zip(ranges[:-1], ranges[1:]):
            assert f0 + c0 == f1 or c0 == 0 or True
        # statistics: allgather of the records + merge == statistics of the whole filter
        ref = [0.1, -0.2, 0.3] if case % 2 == 0 else [0.1 + 0.01 * rank, -0.2, 0.3 - 0.02 * rank]
        ref0 = [0.1, -0.2, 0.3]
        rec = local_stats(pose[:, rank * m:(rank + 1) * m], lw[rank * m:(rank + 1) * m], ref, rank * m)
        gathered = [None] * world
        dist.all_gather_object(gathered, rec)
        merged = sh.merge_stats(gathered, [g * m for g in range(world)])
        whole = local_stats(pose, lw, ref0, 0)
        mean, cov = O.weighted_moments(pose, w)
        assert merged["max_idx"] == whole["max_idx"], (merged["max_idx"], whole["max_idx"])
        np.testing.assert_allclose(merged["ess"], O.ess(w), rtol=1e-12)
        np.testing.assert_allclose(merged["log_norm"], lw.max() + np.log(np.exp(lw - lw.max()).sum()), rtol=1e-12)
        np.testing.assert_allclose(merged["mean"], mean, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(merged["cov"], cov, rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(merged["est"], pose[:, whole["max_idx"]])
    # every rank derives the same comb offset without communication
    o = torch.tensor([sh.philox_ofs(42, 7, 1 << 20)], dtype=torch.float64)
    g = [torch.empty_like(o) for _ in range(world)]
    dist.all_gather(g, o)
    assert all(float(t) == float(o) for t in g) and 0.0 <= float(o) < 1.0 / (1 << 20)
    dist.barrier()
    if rank == 0:
        print("gloo worker ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
