"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: reference-set-sharded Chamfer forward (packed-key MIN
all-reduce) and backward (masked local backward + SUM all-reduce), and reference-set-sharded kNN (all-gather of the
per-rank candidate keys + W-way merge), all driven by the oracle for the per-rank compute."""
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


def _oracle_backward(x1, x2, i1, i2, g1, g2):
    a, b = oracle.chamfer_bwd(x1.numpy(), x2.numpy(), i1.numpy(), i2.numpy(), g1.numpy(), g2.numpy())
    return torch.from_numpy(a), torch.from_numpy(b)


def _oracle_knn_keys(ref_local, query, k, ref_offset):
    r, q = ref_local.numpy(), query.numpy()
    b, nq, _ = q.shape
    keys = np.full((b, nq, k), 0xffffffffffffffff, dtype=np.uint64)
    kk = min(k, r.shape[1])
    if kk:
        d, i = oracle.knn(r, q, kk)  # Euclidean distances; the keys need the squared ones
        d2 = np.empty_like(d)
        for bi in range(b):  # squared distance, KNN_CUDA order: fma over dims from the x product
            diff = r[bi][i[bi]] - q[bi][:, None, :]
            acc = (diff[..., 0] * diff[..., 0]).astype(np.float32)
            for c in range(1, diff.shape[-1]):
                acc = np.float32(1) * (diff[..., c].astype(np.float64) * diff[..., c].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
            d2[bi] = acc
        keys[:, :, :kk] = (d2.view(np.uint32).astype(np.uint64) << np.uint64(32)) | (i.astype(np.uint64) + np.uint64(ref_offset))
    return torch.from_numpy(keys.view(np.int64).copy())


def _numpy_merge(gathered, transpose_out):
    g = gathered.numpy().view(np.uint64)  # (W,B,Q,k)
    w, b, q, k = g.shape
    allk = np.sort(np.moveaxis(g, 0, 2).reshape(b, q, w * k), axis=-1)[:, :, :k]
    d = np.sqrt((allk >> np.uint64(32)).astype(np.uint32).view(np.float32))
    i = (allk & np.uint64(0xffffffff)).astype(np.int64)
    if transpose_out:
        d, i = d.transpose(0, 2, 1), i.transpose(0, 2, 1)
    return torch.from_numpy(np.ascontiguousarray(d)), torch.from_numpy(np.ascontiguousarray(i))


def _worker_bwd_knn(rank, world, port, m_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x1 = torch.from_numpy(synth.clouds(2, 200, seed=3))
    x2 = torch.from_numpy(synth.adversarial(synth.clouds(2, m_total, seed=4), seed=4, n_small=0, n_dup=min(20, m_total // 4)))
    lo, hi = sharded.shard_bounds(m_total, world, rank)
    x2l = x2[:, lo:hi].contiguous()
    wd1, wd2, wi1, wi2 = oracle.chamfer_fwd(x1.numpy(), x2.numpy())
    rng = np.random.default_rng(9)
    g1 = rng.uniform(0.5, 1.5, wd1.shape).astype(np.float32)
    g2 = rng.uniform(0.5, 1.5, wd2.shape).astype(np.float32)
    gx1, gx2l = sharded.chamfer_backward_sharded(x1, x2l, lo, torch.from_numpy(wi1), torch.from_numpy(wi2[:, lo:hi].copy()),
                                                 torch.from_numpy(g1), torch.from_numpy(g2[:, lo:hi].copy()),
                                                 backward_fn=_oracle_backward)
    wg1, wg2 = oracle.chamfer_bwd(x1.numpy(), x2.numpy(), wi1, wi2, g1, g2)
    ok = (np.allclose(gx1.numpy(), wg1, rtol=1e-5, atol=1e-6 * np.abs(wg1).max())
          and np.allclose(gx2l.numpy(), wg2[:, lo:hi], rtol=1e-5, atol=1e-6 * np.abs(wg2).max()))
    # the all-gather form (uneven slices are padded to the largest one): same gradients
    hx1, hx2l = sharded.chamfer_backward_gathered(x1, x2l, lo, m_total, torch.from_numpy(wi1), torch.from_numpy(wi2[:, lo:hi].copy()),
                                                  torch.from_numpy(g1), torch.from_numpy(g2[:, lo:hi].copy()),
                                                  backward_fn=_oracle_backward)
    ok = ok and (np.allclose(hx1.numpy(), wg1, rtol=1e-5, atol=1e-6 * np.abs(wg1).max())
                 and np.allclose(hx2l.numpy(), wg2[:, lo:hi], rtol=1e-5, atol=1e-6 * np.abs(wg2).max()))
    # kNN: 7 queries, k = 5, reference cloud x2 sharded
    k = min(5, m_total)
    query = x1[:, :7].contiguous()
    d, i = sharded.knn_sharded(x2l, query, k, lo, keys_fn=_oracle_knn_keys, merge_fn=_numpy_merge)
    wd, wi = oracle.knn(x2.numpy(), query.numpy(), k)
    ok = ok and np.array_equal(i.numpy(), wi) and np.array_equal(d.numpy(), wd)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("m_total", [301, 6])
def test_sharded_backward_and_knn_gloo_world2(m_total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker_bwd_knn, args=(r, 2, port, m_total, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1
