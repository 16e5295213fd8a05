"""world_size-2 gloo test (CPU) of the multi-GPU host logic: reference-set-sharded Chamfer with the
packed-key MIN all-reduce, driven by the oracle for the per-rank compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cpu as oracle
from pointdae_b200 import sharded, synth


def _oracle_keys(queries, refs, ref_offset):
    q, r = queries.numpy(), refs.numpy()
    b, nq, _ = q.shape
    if r.shape[1] == 0:
        return torch.full((b, nq), 0x7fffffffffffffff, dtype=torch.int64)
    d, _, i, _ = oracle.chamfer_fwd(q, r)
    keys = (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | (i.astype(np.uint64) + np.uint64(ref_offset))
    return torch.from_numpy(keys.view(np.int64).copy())


def _oracle_unpack(keys):
    k = keys.numpy().view(np.uint64)
    d = (k >> np.uint64(32)).astype(np.uint32).view(np.float32)
    i = (k & np.uint64(0xffffffff)).astype(np.int64).astype(np.int32)
    return torch.from_numpy(d.copy()), torch.from_numpy(i.copy())


def _worker(rank, world, port, m_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x1 = torch.from_numpy(synth.clouds(2, 300, seed=1))
    x2 = torch.from_numpy(synth.adversarial(synth.clouds(2, m_total, seed=2), seed=2, n_small=0, n_dup=40))
    lo, hi = sharded.shard_bounds(m_total, world, rank)
    d1, d2l, i1, i2l = sharded.chamfer_forward_sharded(x1, x2[:, lo:hi].contiguous(), lo, keys_fn=_oracle_keys,
                                                       unpack_fn=_oracle_unpack)
    wd1, wd2, wi1, wi2 = oracle.chamfer_fwd(x1.numpy(), x2.numpy())
    ok = (np.array_equal(d1.numpy(), wd1) and np.array_equal(i1.numpy(), wi1)
          and np.array_equal(d2l.numpy(), wd2[:, lo:hi]) and np.array_equal(i2l.numpy(), wi2[:, lo:hi]))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("m_total", [513, 1])
def test_reference_set_sharded_chamfer_gloo_world2(m_total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m_total, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1
