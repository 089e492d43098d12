"""2+ GPU worker: ShardedParticleFilter -- transport from MCL_TRANSPORT: "peer" (device-driven, CUDA IPC peer memory)
or "nccl" (host-driven allgather + point-to-point) -- against the unsharded Filter run on rank 0 with the same Philox
seed: identical particles after every step, identical statistics (bench.sharded_parity_check, the check bench.py
--gpus N repeats in-line).  Prints "<transport> worker ok"."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import slam_robot_simu_b200 as mcl  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    transport = os.environ.get("MCL_TRANSPORT", "nccl")
    ok, text = bench.sharded_parity_check(mcl, torch, dist, dev, transport, n=1 << 18, L=16, steps=12)
    if rank == 0:
        print(text, flush=True)
        if ok:
            print(f"{transport} worker ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
