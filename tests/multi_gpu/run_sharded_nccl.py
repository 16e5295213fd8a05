"""Multi-GPU check of reference-set-sharded Chamfer (forward + backward) and kNN (BASELINE config 5) over NCCL.  Launch with

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
    cham_unsharded_ms = e0.elapsed_time(e1) / reps

    # ---- backward: masked local backward + one all_reduce(SUM) of the replicated cloud's gradient -------------------
    g1 = torch.full_like(fd1, 1.0 / fd1.numel())
    g2 = torch.full_like(fd2, 1.0 / fd2.numel())
    gx1, gx2l = sharded.chamfer_backward_sharded(xyz1, local_refs, lo, i1, i2l, g1, g2[:, lo:hi].contiguous())
    w1, w2 = ops.chamfer_backward(xyz1, xyz2, fi1, fi2, g1, g2)
    ok_b = (torch.allclose(gx1, w1, rtol=1e-5, atol=1e-6 * float(w1.abs().max()))
            and torch.allclose(gx2l, w2[:, lo:hi], rtol=1e-5, atol=1e-6 * float(w2.abs().max())))
    flag_b = torch.tensor([1 if ok_b else 0], device=dev)
    dist.all_reduce(flag_b, op=dist.ReduceOp.MIN)

    # ---- kNN with the reference set sharded (BASELINE config 5: FPS 2048 centres, k = 64) ------------------------------
    q, k = 2048, 64
    centers = ops.fps_gather(xyz2, q)[1]

    def run_knn():
        return sharded.knn_sharded(local_refs, centers, k, lo)

    for _ in range(3):
        run_knn()
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        kd, ki = run_knn()
    e1.record()
    torch.cuda.synchronize()
    tk = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    wd, wi = ops.knn_points(xyz2, centers, k)
    flag_k = torch.tensor([1 if (torch.equal(kd, wd) and torch.equal(ki, wi)) else 0], device=dev)
    dist.all_reduce(flag_k, op=dist.ReduceOp.MIN)
    e0.record()
    for _ in range(reps):
        ops.knn_points(xyz2, centers, k)
    e1.record()
    torch.cuda.synchronize()
    good = bool(flag.item()) and bool(flag_b.item()) and bool(flag_k.item())
    if rank == 0:
        print(json.dumps({"test": "ref-set-sharded chamfer + kNN over NCCL", "n_points": n, "world": world,
                          "chamfer_bit_exact_vs_unsharded": bool(flag.item()), "chamfer_sharded_ms": float(t.item()),
                          "chamfer_unsharded_1gpu_ms": cham_unsharded_ms,
                          "chamfer_backward_matches_unsharded": bool(flag_b.item()),
                          "knn": {"queries": q, "k": k, "bit_exact_vs_unsharded": bool(flag_k.item()),
                                  "sharded_ms": float(tk.item()), "unsharded_1gpu_ms": e0.elapsed_time(e1) / reps}}),
              flush=True)
    dist.destroy_process_group()
    sys.exit(0 if good else 1)


if __name__ == "__main__":
    main()
