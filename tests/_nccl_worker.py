"""2+ GPU worker: ShardedParticleFilter -- transport from MCL_TRANSPORT: "peer" (device-driven, CUDA IPC peer memory)
or "nccl" (host-driven allgather + point-to-point) -- against the unsharded Filter run on rank 0 with the same Philox
seed: identical particles after every step, identical statistics.  Prints "<transport> worker ok"."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import slam_robot_simu_b200 as mcl  # noqa: E402

ShardedParticleFilter = mcl.ShardedParticleFilter


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    n, L = 1 << 18, 16
    m = n // world
    lm = bench.landmarks_for(L)
    truth = bench.Truth(lm, seed=3)
    transport = os.environ.get("MCL_TRANSPORT", "nccl")
    sf = ShardedParticleFilter(n, lm, likelihood="moment", seed=11, transport=transport)
    assert sf.transport == transport, (sf.transport, sf.transport_note)
    sf.init_particles(truth.pose()[0])
    ref = None
    if rank == 0:
        ref = mcl.Filter(n_particles=n, landmarks=lm, likelihood="moment", seed=11, max_landmarks=L)
        ref.init_particles(truth.pose()[0])
    ok = True
    for k in range(12):
        truth.advance()
        z = truth.observe()
        force = k % 3 != 1  # mix resampling and non-resampling steps
        if rank == 0 and k > 0:
            # the unsharded filter quantises with the shards' MERGED statistics (its own merge groups the fp32 partial
            # sums differently: equal to ~1e-9, which very rarely moves the fp32 scale of the weight map by one ulp)
            ref.restore_stats(sf.flt.stats()[0], k)
        g = sf.step(z, force_resample=force)
        x, y, th, lw = sf.local_particles()
        parts = [torch.from_numpy(np.stack([x, y, th, lw])).cuda()]
        allp = [torch.empty_like(parts[0]) for _ in range(world)]
        dist.all_gather(allp, parts[0])
        if rank == 0:
            ref.step(z, force_resample=force, no_resample=not force)
            rx, ry, rt, rl = ref.get_particles()
            got = torch.cat(allp, dim=1).cpu().numpy()
            st = ref.stats()[0]
            same = np.array_equal(got[0], rx[0]) and np.array_equal(got[1], ry[0]) and np.array_equal(got[2], rt[0])
            # log-weights are stored relative to a per-step constant that both runs derive identically
            same = same and np.array_equal(got[3], rl[0])
            # the fp32 group sums are grouped per shard: statistics agree to fp32-sum accuracy, not bitwise
            close = abs(g["ess"] / st.ess - 1) < 1e-6 and g["max_idx"] == st.max_idx and abs(g["log_norm"] - st.log_norm) < 1e-7
            if not (same and close):
                print(f"step {k}: particles equal={same} stats close={close} ess {g['ess']} vs {st.ess} idx {g['max_idx']} vs {st.max_idx}",
                      flush=True)
                ok = False
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    if rank == 0 and ok:
        print(f"{transport} worker ok", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
