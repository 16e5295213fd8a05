"""`pdae_step_f32` (csrc/step.cu): the hot path's step as one native call -- Chamfer forward, the patchifier as a
programmatic dependent launch, fused mean loss and backward on library-owned streams.  Same bits as the separate entry
points (which the other tests pin to the oracle), repeated calls on rotating inputs, inside a CUDA graph, and on shapes
that take the two-launch patchifier / the FP32-pipe forward."""
import numpy as np
import pytest
import torch

from oracle import cpu as oracle
from pointdae_b200 import _native, ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _separate(c, p, g, m, gone):
    fps_idx, center, nb, _ = ops.fps_group(c, g, m)
    d1, d2, i1, i2 = ops.chamfer_forward(p, c)
    loss3 = ops.chamfer_mean_loss(d1, d2)
    gp, gc = ops.chamfer_loss_backward(p, c, i1, i2, d1, d2, gone, 1.0, 1.0)
    return dict(fps_idx=fps_idx, center=center, neighborhood=nb, dist1=d1, dist2=d2, idx1=i1, idx2=i2, loss3=loss3, gpred=gp, gcloud=gc)


def _same(o, want, grads_exact=True):
    for k in ("fps_idx", "center", "neighborhood", "dist1", "dist2", "idx1", "idx2", "loss3"):
        assert torch.equal(getattr(o, k), want[k]), k
    for k in ("gpred", "gcloud"):  # the scatter adds float atomics in any order (like the reference's backward)
        got, ref = getattr(o, k), want[k]
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5 * float(ref.abs().max())), k


@pytest.mark.parametrize("b,n,g,m", [(128, 2048, 64, 32), (64, 1024, 64, 32), (3, 700, 20, 16), (2, 300, 8, 8), (2, 4096, 32, 32)])
def test_step_equals_the_separate_entry_points(b, n, g, m):
    xyz = synth.clouds(b, n, seed=n + b)
    c, p = cu(xyz), cu(synth.prediction(xyz, seed=b))
    gone = torch.full((1,), 0.5, device=DEV)
    want = _separate(c, p, g, m, gone)
    o = ops.hot_step(c, p, g, m, gone)
    torch.cuda.synchronize()
    _same(o, want)


def test_step_matches_the_oracle_and_is_repeatable_on_rotating_inputs():
    b, n, g, m = 48, 1024, 64, 32
    gone = torch.ones(1, device=DEV)
    bufs = ops.StepBuffers(b, n, g, m, torch.device(DEV))
    clouds = [synth.adversarial(synth.clouds(b, n, seed=s), seed=s) for s in (1, 2, 3)]
    preds = [synth.prediction(x, seed=7) for x in clouds]
    C, P = [cu(x) for x in clouds], [cu(x) for x in preds]
    for rep in range(2):
        for k in range(3):
            o = ops.hot_step(C[k], P[k], g, m, gone, buffers=bufs)
            assert o is bufs
            want_nb, want_c, _, want_fps = oracle.group(clouds[k], g, m)
            wd1, wd2, wi1, wi2 = oracle.chamfer_fwd(preds[k], clouds[k])
            np.testing.assert_array_equal(o.neighborhood.cpu().numpy(), want_nb)
            np.testing.assert_array_equal(o.center.cpu().numpy(), want_c)
            np.testing.assert_array_equal(o.fps_idx.cpu().numpy(), want_fps)
            np.testing.assert_array_equal(o.idx1.cpu().numpy(), wi1)
            np.testing.assert_array_equal(o.dist2.cpu().numpy(), wd2)
            gd = np.full(wd1.shape, 1.0 / wd1.size, dtype=np.float32)
            wg1, wg2 = oracle.chamfer_bwd(preds[k], clouds[k], wi1, wi2, gd, gd)
            assert np.allclose(o.gpred.cpu().numpy(), wg1, rtol=1e-5, atol=1e-5 * np.abs(wg1).max())
            assert np.allclose(o.gcloud.cpu().numpy(), wg2, rtol=1e-5, atol=1e-5 * np.abs(wg2).max())
            assert abs(float(o.loss3[0]) - (wd1.astype(np.float64).mean() + wd2.astype(np.float64).mean())) < 1e-6


def test_step_is_capturable_and_replays():
    b, n, g, m = 64, 2048, 64, 32
    xyz = synth.clouds(b, n, seed=9)
    c, p = cu(xyz), cu(synth.prediction(xyz, seed=9))
    gone = torch.ones(1, device=DEV)
    want = _separate(c, p, g, m, gone)
    bufs = ops.StepBuffers(b, n, g, m, torch.device(DEV))
    ops.hot_step(c, p, g, m, gone, buffers=bufs)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ops.hot_step(c, p, g, m, gone, buffers=bufs)
    for k in ("neighborhood", "dist1", "gpred", "loss3"):
        getattr(bufs, k).zero_()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    _same(bufs, want)


def test_two_host_threads_on_two_streams_share_the_helper_streams_safely():
    """the library-owned streams / events are per device: calls from two host threads (the reference's nn.DataParallel
    runs one thread per replica) must each wait on their own forward"""
    import threading
    b, n, g, m = 64, 2048, 64, 32
    gone = torch.ones(1, device=DEV)
    data, want, errs = [], [], []
    for t in range(2):
        xyz = synth.clouds(b, n, seed=40 + t)
        c, p = cu(xyz), cu(synth.prediction(xyz, seed=t))
        data.append((c, p))
        want.append(_separate(c, p, g, m, gone))
    torch.cuda.synchronize()

    def work(t):
        try:
            st = torch.cuda.Stream()
            bufs = ops.StepBuffers(b, n, g, m, torch.device(DEV))
            with torch.cuda.stream(st):
                for _ in range(20):
                    ops.hot_step(data[t][0], data[t][1], g, m, gone, buffers=bufs)
            st.synchronize()
            _same(bufs, want[t])
        except Exception as e:  # surfaced in the main thread
            errs.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs


def test_step_rejects_bad_arguments():
    L = _native.lib()
    assert L.pdae_step_f32(None, None, 2, 600, 8, 4, *([None] * 11), None, 0, None) != 0
    assert L.pdae_step_f32(None, None, 0, 600, 8, 4, *([None] * 11), None, 0, None) == 0
    c = cu(synth.clouds(2, 600, seed=1))
    with pytest.raises(RuntimeError):
        ops.hot_step(c, c[:, :500].contiguous(), 8, 4, torch.ones(1, device=DEV))
