"""Multi-GPU forms of the path (one process per GPU, torch.distributed for the plumbing).

* batch sharding (the default): every op is independent per cloud, so rank r simply
  owns clouds [lo, hi) = shard_bounds(B, world, r); there is NO data-path collective.
* reference-set sharding for scene-scale clouds (BASELINE config 5): each rank holds a slice of the
  reference cloud, scans all queries against it and emits packed keys
  (float_bits(min d) << 32 | global argmin); ONE all-reduce(MIN) over int64 per direction yields
  the global (distance, lowest argmin); the key is order-preserving because d >= 0.
  The backward needs one all-reduce(SUM) of the gradient of the replicated cloud (the sharded cloud's gradient is
  final on the rank that owns the slice).
* reference-set sharding for kNN: each rank emits its k best candidates per query as ascending packed keys, the
  lists are all-gathered (W*Q*k*8 bytes) and merged W-way on every rank -- same (distance, lower index) order as
  the single-GPU kernel, so the result is bit-identical.

The collectives are torch.distributed calls (NCCL over NVLink on the GPU box, gloo in the CPU
tests); the per-rank compute is pdae_chamfer_min_keys_u64 / pdae_chamfer_unpack_keys.  `keys_fn` /
`unpack_fn` exist so the CPU (gloo) tests can drive this host logic with the oracle.
"""
import torch
import torch.distributed as dist


def exchange_name(exchange=None):
    """which exchange the sharded forward uses (bench / test reports)"""
    if exchange is None:
        return "torch.distributed all_reduce(MIN) on packed int64 keys (NCCL over NVLink)"
    return exchange.name()


class PeerExchange:
    """Symmetric buffers of one rank for the fused exchange of reference-set-sharded Chamfer (csrc/exchange.cu): packed row
    keys | distances | indices of `rows` rows in ONE symmetric allocation (torch.distributed._symmetric_memory: every
    rank's copy is mapped into every rank), the peer pointers, and -- when the allocation has one -- the NVSwitch
    multicast address.  Collective: every rank of `group` constructs it with the same `rows`."""

    def __init__(self, rows, device, group=None, use_multimem=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        self.rows = int(rows)
        self.buf = symm_mem.empty(self.rows * 16, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.world, self.rank = int(self.hdl.world_size), int(self.hdl.rank)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        arr = ctypes.c_void_p * self.world
        self.keys_ptrs = arr(*base)
        self.dist_ptrs = arr(*[p + self.rows * 8 for p in base])
        self.idx_ptrs = arr(*[p + self.rows * 12 for p in base])
        self.keys = self.buf[: self.rows * 8].view(torch.int64)
        self.dist = self.buf[self.rows * 8: self.rows * 12].view(torch.float32)
        self.idx = self.buf[self.rows * 12: self.rows * 16].view(torch.int32)
        mc = 0
        try:
            mc = int(self.hdl.multicast_ptr or 0)
        except Exception:
            mc = 0
        self.mc = mc
        self.use_multimem = bool(mc) if use_multimem is None else (bool(use_multimem) and bool(mc))

    def name(self):
        return ("one kernel over NVLink peer memory (symmetric buffers): " +
                ("in-switch reduction, multimem.ld_reduce.min.u64 + multimem.st (NVLS)" if self.use_multimem
                 else "peer loads + min + peer stores"))

    def run(self):
        """keys of every rank -> (dist, idx) of all rows on every rank; returns views of this rank's result buffers
        (valid until the next run)."""
        from . import _native, ops
        lo, hi = shard_bounds(self.rows, self.world, self.rank)
        L = _native.lib()
        self.hdl.barrier(channel=0)  # every rank's keys are complete
        with torch.cuda.device(self.buf.device):
            st = torch.cuda.current_stream().cuda_stream
            if self.use_multimem:
                rc = L.pdae_chamfer_exchange_keys_multimem(self.mc, self.mc + self.rows * 8, self.mc + self.rows * 12, lo, hi, st)
            else:
                rc = L.pdae_chamfer_exchange_keys_peer(self.keys_ptrs, self.dist_ptrs, self.idx_ptrs, self.world, lo, hi, st)
        _native.check(rc, "pdae_chamfer_exchange_keys")
        self.hdl.barrier(channel=1)  # every rank's share has landed everywhere
        return self.dist, self.idx


def make_exchange(rows, device, group=None, use_multimem=None):
    """PeerExchange, or None where symmetric memory is not available (then the NCCL all-reduce is used)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and torch.device(device).type == "cuda"):
        return None
    try:
        return PeerExchange(rows, device, group, use_multimem)
    except Exception:
        return None


def shard_bounds(n, world, rank):
    """contiguous, balanced partition of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def _default_keys_fn(queries, refs, ref_offset):
    from . import ops
    return ops.chamfer_min_keys(queries, refs, ref_offset)


def _default_unpack_fn(keys):
    from . import ops
    return ops.chamfer_unpack_keys(keys)


def chamfer_direction_sharded(queries, refs_local, ref_offset, group=None, keys_fn=None, unpack_fn=None):
    """min / argmin of every query (B,Nq,3) over a reference cloud whose slice [ref_offset,
    ref_offset + refs_local.size(1)) lives on this rank.  Returns (dist (B,Nq) f32, idx (B,Nq) int32 global)."""
    keys_fn = keys_fn or _default_keys_fn
    unpack_fn = unpack_fn or _default_unpack_fn
    keys = keys_fn(queries, refs_local, ref_offset)  # int64 (B,Nq)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return unpack_fn(keys)


def chamfer_forward_sharded(xyz1, xyz2_local, xyz2_offset, group=None, keys_fn=None, unpack_fn=None, exchange=None):
    """Chamfer forward with cloud 2 (the big reference / target cloud) sharded along its points.

    xyz1 (B,N,3) is replicated; rank r holds xyz2[:, off:off+m_local].  Returns dist1, idx1 for all of
    xyz1 (global indices into xyz2) and dist2_local, idx2_local for this rank's slice of xyz2 (its
    nearest neighbours in the replicated xyz1 need no exchange).  On CUDA the rank's whole share is ONE pass
    (pdae_chamfer_sharded_f32: every local pair evaluated once); `keys_fn` / `unpack_fn` let the CPU tests drive
    the same host logic with the oracle."""
    if keys_fn is None and unpack_fn is None and exchange is not None:
        # fused exchange over NVLink peer memory: the local pass writes its keys into the symmetric buffer
        from . import ops
        b, n = xyz1.shape[:2]
        if exchange.rows != b * n:
            raise RuntimeError("exchange was built for %d rows, got %d" % (exchange.rows, b * n))
        _, dist2_local, idx2_local = ops.chamfer_sharded_local(xyz1, xyz2_local, xyz2_offset, out_keys=exchange.keys.view(b, n))
        dist1, idx1 = exchange.run()
        return dist1.view(b, n), dist2_local, idx1.view(b, n), idx2_local
    if keys_fn is None and unpack_fn is None:
        from . import ops
        keys, dist2_local, idx2_local = ops.chamfer_sharded_local(xyz1, xyz2_local, xyz2_offset)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
        dist1, idx1 = ops.chamfer_unpack_keys(keys)
        return dist1, dist2_local, idx1, idx2_local
    dist1, idx1 = chamfer_direction_sharded(xyz1, xyz2_local, xyz2_offset, group, keys_fn, unpack_fn)
    keys_fn = keys_fn or _default_keys_fn
    unpack_fn = unpack_fn or _default_unpack_fn
    dist2_local, idx2_local = unpack_fn(keys_fn(xyz2_local, xyz1, 0))
    return dist1, dist2_local, idx1, idx2_local


def chamfer_backward_gathered(xyz1, xyz2_local, xyz2_offset, m_total, idx1, idx2_local, grad_dist1, grad_dist2_local,
                              group=None, backward_fn=None):
    """Backward of chamfer_forward_sharded for clouds whose backward is cheap next to a collective (N = 100 000: 30 us on
    one GPU): ONE all-gather of every rank's packed (slice of cloud 2 | its match indices | its upstream gradient), then
    every rank runs the ordinary backward kernels on the whole pair and keeps its slice of grad_xyz2 -- no masking
    passes, no all-reduce of a gradient.  Slices must follow shard_bounds(m_total, world, rank).
    Returns (grad_xyz1 (B,N,3) complete on every rank, grad_xyz2_local (B,m_local,3))."""
    if backward_fn is None:
        from . import ops
        backward_fn = ops.chamfer_backward
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    b, ml = xyz2_local.shape[:2]
    if world == 1:
        return backward_fn(xyz1, xyz2_local, idx1.to(torch.int32), idx2_local, grad_dist1, grad_dist2_local)
    ml_max = (int(m_total) + world - 1) // world
    pack = torch.zeros((b, ml_max, 5), dtype=torch.float32, device=xyz2_local.device)
    pack[:, :ml, :3] = xyz2_local
    pack[:, :ml, 3] = idx2_local.view(torch.float32) if idx2_local.dtype == torch.int32 else idx2_local.to(torch.int32).view(torch.float32)
    pack[:, :ml, 4] = grad_dist2_local
    gathered = torch.empty((world,) + tuple(pack.shape), dtype=torch.float32, device=pack.device)
    dist.all_gather_into_tensor(gathered.view(world * b, ml_max, 5), pack, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(m_total, world, r)
        parts.append(gathered[r, :, : hi - lo])
    full = torch.cat(parts, dim=1)  # (B, M, 5)
    xyz2 = full[:, :, :3].contiguous()
    idx2 = full[:, :, 3].contiguous().view(torch.int32)
    gd2 = full[:, :, 4].contiguous()
    gx1, gx2 = backward_fn(xyz1, xyz2, idx1.to(torch.int32).contiguous(), idx2, grad_dist1, gd2)
    return gx1, gx2[:, xyz2_offset:xyz2_offset + ml].contiguous()


def chamfer_backward_sharded(xyz1, xyz2_local, xyz2_offset, idx1, idx2_local, grad_dist1, grad_dist2_local, group=None,
                             backward_fn=None):
    """Backward of chamfer_forward_sharded.  idx1 holds GLOBAL indices into xyz2; a pair (a_j, b_idx1[j]) can only be
    differentiated by the rank that owns b, so every rank runs the ordinary backward kernels on its slice with the
    foreign pairs masked out (zero upstream gradient), which yields the final gradient of its slice of xyz2 and a
    partial gradient of the replicated xyz1; ONE all-reduce(SUM) completes the latter.
    Returns (grad_xyz1 (B,N,3) complete on every rank, grad_xyz2_local (B,m_local,3))."""
    if backward_fn is None:
        from . import ops
        backward_fn = ops.chamfer_backward
    m_local = xyz2_local.size(1)
    local = (idx1 >= xyz2_offset) & (idx1 < xyz2_offset + m_local)
    idx1_local = torch.where(local, idx1 - xyz2_offset, torch.zeros_like(idx1)).to(torch.int32)
    gd1 = torch.where(local, grad_dist1, torch.zeros_like(grad_dist1))
    if m_local == 0:
        gx1, gx2_local = torch.zeros_like(xyz1), torch.zeros_like(xyz2_local)
    else:
        gx1, gx2_local = backward_fn(xyz1, xyz2_local, idx1_local, idx2_local, gd1, grad_dist2_local)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(gx1, op=dist.ReduceOp.SUM, group=group)
    return gx1, gx2_local


def knn_sharded(ref_local, query, k, ref_offset, group=None, transpose_out=False, keys_fn=None, merge_fn=None):
    """knn_cuda.KNN over a reference cloud sharded along its points: ref_local (B,r_local,D) = points
    [ref_offset, ref_offset + r_local) of the cloud, query (B,Q,D) replicated.  Returns (dist, idx) identical to the
    unsharded KNN(k, transpose_mode=True) (or the (B,k,Q) layout of transpose_mode=False with transpose_out)."""
    if keys_fn is None:
        from . import ops
        keys_fn = ops.knn_keys
    if merge_fn is None:
        from . import ops
        merge_fn = ops.knn_merge_keys
    keys = keys_fn(ref_local, query, k, ref_offset)  # (B,Q,k) int64
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
        flat = torch.empty((world * keys.size(0),) + tuple(keys.shape[1:]), dtype=keys.dtype, device=keys.device)
        dist.all_gather_into_tensor(flat, keys, group=group)  # rank-major concatenation along dim 0
        gathered = flat.view((world,) + tuple(keys.shape))
    else:
        gathered = keys.unsqueeze(0)
    return merge_fn(gathered.contiguous(), transpose_out)
