"""Multi-GPU forms of the path (one process per GPU, torch.distributed for the plumbing).

* batch sharding (the default): every op is independent per cloud, so rank r simply
  owns clouds [lo, hi) = shard_bounds(B, world, r); there is NO data-path collective.
* reference-set sharding for scene-scale clouds (BASELINE config 5): each rank holds a slice of the
  reference cloud, scans all queries against it and emits packed keys
  (float_bits(min d) << 32 | global argmin); ONE all-reduce(MIN) over int64 per direction yields
  the global (distance, lowest argmin); the key is order-preserving because d >= 0.
  The backward then needs one all-reduce(SUM) of the scattered gradient of the sharded cloud.

The collectives are torch.distributed calls (NCCL over NVLink on the GPU box, gloo in the CPU
tests); the per-rank compute is pdae_chamfer_min_keys_u64 / pdae_chamfer_unpack_keys.  `keys_fn` /
`unpack_fn` exist so the CPU (gloo) tests can drive this host logic with the oracle.
"""
import torch
import torch.distributed as dist


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


def chamfer_forward_sharded(xyz1, xyz2_local, xyz2_offset, group=None, keys_fn=None, unpack_fn=None):
    """Chamfer forward with cloud 2 (the big reference / target cloud) sharded along its points.

    xyz1 (B,N,3) is replicated; rank r holds xyz2[:, off:off+m_local].  Returns dist1, idx1 for all of
    xyz1 (global indices into xyz2) and dist2_local, idx2_local for this rank's slice of xyz2 (its
    nearest neighbours in the replicated xyz1 need no exchange).  On CUDA the rank's whole share is ONE pass
    (pdae_chamfer_sharded_f32: every local pair evaluated once); `keys_fn` / `unpack_fn` let the CPU tests drive
    the same host logic with the oracle."""
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
