"""Column-split units of the symmetric Chamfer forward (every 512-row block's sweep cut into column chunks whose row
minima merge through packed (distance, index) keys): results must not depend on the chunk count, on the phased entry
point, or on the workspace size the caller offers -- and must equal the oracle (chamfer.cu:15-171 restated)."""
import numpy as np
import pytest
import torch

from oracle import cpu as oracle
from pointdae_b200 import _native, ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture
def split_hook():
    L = _native.lib()
    old = L.pdae_tune_chamfer_split(-1)
    yield L.pdae_tune_chamfer_split
    L.pdae_tune_chamfer_split(old)


@pytest.mark.parametrize("b,n,m", [(1, 2048, 2048), (3, 1500, 700), (2, 513, 1025), (5, 300, 2049), (2, 4096, 4095),
                                   (7, 1024, 1024), (1, 20000, 3000)])
def test_every_chunk_count_equals_the_oracle(split_hook, b, n, m):
    a, c = synth.clouds(b, n, seed=n + b), synth.clouds(b, m, seed=m + 7)
    a[:, 5] = c[:, 3]          # exact zero distance
    c[:, 11] = c[:, 10]        # duplicate reference point: the lower index must win in every chunking
    if m > 700:
        c[:, 600] = c[:, 10]   # ... also across chunk / tile boundaries
        a[:, 9] = a[:, 8]
    want = oracle.chamfer_fwd(a, c)
    ta, tc = cu(a), cu(c)
    for nc in (1, 2, 3, 5, 16, 0):
        split_hook(nc)
        got = ops.chamfer_forward(ta, tc)
        for g, w, name in zip(got, want, ("dist1", "dist2", "idx1", "idx2")):
            np.testing.assert_array_equal(g.cpu().numpy(), w, err_msg="%s with %d chunks" % (name, nc))


def test_headline_shape_split_vs_unsplit_and_context_manager(split_hook):
    a = cu(synth.prediction(synth.clouds(128, 2048, seed=1), seed=1))
    c = cu(synth.clouds(128, 2048, seed=1))
    split_hook(1)
    want = ops.chamfer_forward(a, c)
    for nc in (0, 2, 4):
        split_hook(nc)
        got = ops.chamfer_forward(a, c)
        assert all(torch.equal(g, w) for g, w in zip(got, want))
    split_hook(2)
    with ops.chamfer_column_split(False):  # column-key-only workspace: the library cannot split
        got = ops.chamfer_forward(a, c)
    assert all(torch.equal(g, w) for g, w in zip(got, want))
    ev = torch.cuda.Event()
    got = ops.chamfer_forward(a, c, scan_done=ev)  # phased entry point: separate unpack launch, same values
    assert all(torch.equal(g, w) for g, w in zip(got, want))


def test_split_forward_feeds_the_fused_loss_and_backward(split_hook):
    from pointdae_b200 import chamfer_dist
    xyz = synth.clouds(4, 3000, seed=3)
    pred = synth.prediction(xyz, seed=3)
    outs = []
    for nc in (1, 3):
        split_hook(nc)
        p = cu(pred).requires_grad_(True)
        loss = chamfer_dist.ChamferDistanceL1()(p, cu(xyz))
        loss.backward()
        outs.append((loss.detach().clone(), p.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-6, atol=1e-9)  # atomics: order-free sums


@pytest.mark.parametrize("n,m,lo,hi", [(3000, 5000, 1000, 3600), (100000, 30000, 0, 12500), (700, 2000, 1999, 2000)])
def test_sharded_entry_points_merge_their_chunks_into_the_returned_keys(split_hook, n, m, lo, hi):
    """pdae_chamfer_sharded_f32 / pdae_chamfer_min_keys_u64: the packed row keys a rank hands to the MIN all-reduce are
    themselves the merge target of its column-split units (global indices, identity-filled first)."""
    a = cu(synth.prediction(synth.clouds(1, max(n, m), seed=n), seed=n)[:, :n])
    c = cu(synth.clouds(1, max(n, m), seed=n)[:, :m])
    c[:, lo + 1 if lo + 1 < hi else lo] = c[:, lo]  # duplicate inside the slice
    sl = c[:, lo:hi].contiguous()
    split_hook(1)
    want = ops.chamfer_sharded_local(a, sl, lo)
    want_keys = ops.chamfer_min_keys(a, sl, lo)
    assert torch.equal(want[0], want_keys)
    d, i = ops.chamfer_unpack_keys(want[0])
    assert int(i.min()) >= lo and int(i.max()) < hi
    for nc in (2, 3, 7, 0):
        split_hook(nc)
        got = ops.chamfer_sharded_local(a, sl, lo)
        assert all(torch.equal(g, w) for g, w in zip(got, want)), nc
        assert torch.equal(ops.chamfer_min_keys(a, sl, lo), want_keys), nc
