"""Multi-GPU check of reference-set-sharded Chamfer (BASELINE config 5) over NCCL.  Launch with

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/multi_gpu/run_sharded_nccl.py [n_points]

Every rank holds all queries and 1/N of the reference cloud; one all_reduce(MIN) of packed int64 keys per direction
yields the global (distance, lowest argmin).  Rank 0 compares with the unsharded kernel (bit-exact) and prints one
JSON line with device timings (max over ranks)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from pointdae_b200 import ops, sharded, synth


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    xyz2 = torch.from_numpy(synth.adversarial(synth.clouds(1, n, seed=5), seed=5, n_small=0, n_dup=200)).to(dev)
    xyz1 = torch.from_numpy(synth.prediction(synth.clouds(1, n, seed=5), seed=5)).to(dev)
    lo, hi = sharded.shard_bounds(n, world, rank)
    local_refs = xyz2[:, lo:hi].contiguous()

    def run():
        return sharded.chamfer_forward_sharded(xyz1, local_refs, lo)

    for _ in range(3):
        run()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        d1, d2l, i1, i2l = run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)

    fd1, fd2, fi1, fi2 = ops.chamfer_forward(xyz1, xyz2)
    ok = torch.equal(d1, fd1) and torch.equal(i1, fi1) and torch.equal(d2l, fd2[:, lo:hi]) and torch.equal(i2l, fi2[:, lo:hi])
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    e0.record()
    for _ in range(reps):
        ops.chamfer_forward(xyz1, xyz2)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"test": "ref-set-sharded chamfer over NCCL", "n_points": n, "world": world,
                          "bit_exact_vs_unsharded": bool(flag.item()), "sharded_ms": float(t.item()),
                          "unsharded_1gpu_ms": e0.elapsed_time(e1) / reps}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
