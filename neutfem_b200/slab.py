"""
z-slab decomposition of the k-eff hot path across the GPUs of one box: one process per GPU (torchrun), slabs of
whole z planes, `torch.distributed` only for plumbing (rendezvous, broadcasting the NCCL id, gathering results);
the data path (CG scalars, interface values of the z-direction line solves) goes through the NCCL communicator
owned by libneutfem_b200.so. The reference is single-process; see DESIGN.md "z-slabs".
"""
from __future__ import annotations

import numpy as np


def partition_planes(nz: int, nranks: int):
    """Contiguous, balanced split of nz planes into nranks slabs [(z0, z1), ...]; every rank gets >= 1 plane."""
    if nranks > nz:
        raise ValueError(f"cannot split {nz} planes over {nranks} ranks")
    base, rem = divmod(nz, nranks)
    out, z = [], 0
    for r in range(nranks):
        n = base + (1 if r < rem else 0)
        out.append((z, z + n))
        z += n
    return out


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from `src` with torch.distributed (any backend)."""
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


class SlabSolver:
    """One rank's share of a z-slab solve. XS / flux arrays passed in or returned are LOCAL (planes [z0, z1))."""

    def __init__(self, rt_order, p_order, ng, x_breaks, y_breaks, z_breaks, rank, world, device, bcast=broadcast_bytes):
        from . import cabi
        self.cabi = cabi
        zb = np.asarray(z_breaks, dtype=np.float64)
        self.rank, self.world = int(rank), int(world)
        self.z0, self.z1 = partition_planes(zb.size - 1, world)[rank]
        self.ctx = cabi.Context(rt_order, p_order, ng, x_breaks, y_breaks, zb, device=device,
                                slab=(self.z0, self.z1, rank, world))
        if world > 1:
            uid = cabi.comm_unique_id() if rank == 0 else None
            uid = bcast(uid, 0)
            self.ctx.comm_init(uid, rank, world)

    def local_planes(self, a, ng_axes=1):
        """Slice a global array shaped [ng(,ng), nz, ny, nx] (or flat) to this rank's planes, flattened."""
        c = self.ctx
        nzg = a.size // (c.ng ** ng_axes * c.nx * c.ny)
        v = np.asarray(a).reshape([c.ng] * ng_axes + [nzg, c.ny, c.nx])
        return np.ascontiguousarray(v[..., self.z0:self.z1, :, :]).ravel()

    def close(self):
        self.ctx.close()
