#!/usr/bin/env python
"""Where a sharded step spends its host time (diagnostic; run under torchrun with >= 2 GPUs).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_timeline.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import slam_robot_simu_b200 as mcl  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    n = int(os.environ.get("N_PER_GPU", 125_000_000)) * world
    lm = bench.landmarks_for(64)
    tr = bench.Truth(lm, seed=7)
    sf = mcl.ShardedParticleFilter(n, lm, likelihood="moment", seed=5)
    sf.init_particles(tr.pose()[0])
    cyc = tr.cycle()
    marks = {}
    sh = sys.modules[type(sf).__module__]

    def wrap(obj, name, label):
        fn = getattr(obj, name)

        def inner(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            marks[label] = marks.get(label, 0.0) + time.perf_counter() - t0
            return r
        setattr(obj, name, inner)

    wrap(sh, "shard_ranges", "host shard_ranges")
    wrap(sh, "plan_offrank", "host plan_offrank (x2)")
    wrap(sh, "philox_ofs", "host philox_ofs")
    wrap(sh, "exchange_offrank", "exchange_offrank (enqueue)")
    wrap(sf.flt, "shard_expand", "shard_expand (enqueue)")
    wrap(sf.flt, "shard_step", "shard_step (enqueue)")
    wrap(sf.flt, "stats_dev_bytes", "sync_stats (enqueue)")
    wrap(sf.flt, "shard_merge", "shard_merge (enqueue)")
    wrap(sf.flt, "shard_total", "shard_total (enqueue)")
    wrap(sf.flt, "stats", "stats() = D2H + WAIT for the GPU")
    wrap(sf.dist, "all_gather_into_tensor", "allgathers (enqueue)")
    # GPU-side phases: CUDA events on the current stream after each enqueue
    phases = ["expand (reset + scan + reset)", "exchange (NCCL p2p group)", "shard_step (prep + k_step)",
              "k_merge", "allgather stats", "k_shard_merge", "k_shard_total", "allgather totals"]
    evs = []

    def stamp(obj, name):
        fn = getattr(obj, name)

        def inner(*a, **k):
            r = fn(*a, **k)
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append(e)
            return r
        setattr(obj, name, inner)

    if os.environ.get("GPU_PHASES", "1") == "1":
        stamp(sf.flt, "shard_expand")
        stamp(sh, "exchange_offrank")
        stamp(sf.flt, "shard_step")
        stamp(sf.flt, "stats_dev_bytes")
        stamp(sf.dist, "all_gather_into_tensor")
        stamp(sf.flt, "shard_merge")
        stamp(sf.flt, "shard_total")
    for k in range(5):
        sf.step(cyc[k % bench.Truth.PERIOD], force_resample=True)
    marks.clear()
    evs.clear()
    K = 20
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for k in range(5, 5 + K):  # the observation sequence continues: the filter stays in its tracking regime
        sf.step(cyc[k % bench.Truth.PERIOD], force_resample=True)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / K
    if rank == 0 and evs:
        per = len(phases)
        assert len(evs) == per * K, (len(evs), per, K)
        acc = [0.0] * per
        gap = 0.0
        for i in range(K):
            for j in range(1, per):
                acc[j] += evs[i * per + j - 1].elapsed_time(evs[i * per + j])
            if i > 0:
                gap += evs[i * per - 1].elapsed_time(evs[i * per])  # last stamp of step i-1 -> end of expand of step i
        for j in range(1, per):
            print(f"  GPU {phases[j]:36s} {acc[j] / K * 1e3:8.1f} us")
        print(f"  GPU {'end of step -> end of next expand':36s} {gap / (K - 1) * 1e3:8.1f} us  (idle + " + phases[0] + ")")
    if rank == 0:
        print(f"wall per step {wall * 1e3:.3f} ms")
        tot = 0.0
        for name, v in marks.items():
            print(f"  {name:40s} {v / K * 1e6:8.1f} us")
            tot += v / K
        print(f"  {'(unattributed python)':40s} {(wall - tot) * 1e6:8.1f} us")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
