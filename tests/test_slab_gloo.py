"""CPU, world_size 2 (gloo): the z-slab substructuring used by the multi-GPU path reproduces the global line solve,
and the plane partition / plumbing helpers behave. (The CUDA slab kernels are covered by tests/test_gpu_slab.py on 2 GPUs.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from slab_model import local_matrix, local_rhs, rank_finish, rank_publish, reduced_solve
        from neutfem_b200.slab import partition_planes
        rng = np.random.default_rng(4)
        nz = 13
        c = rng.uniform(0.2, 3.0, nz)            # f_z / D per cell along one z line
        x0 = rng.uniform(0.5, 1.5, nz)
        alpha, off = 0.25, -1.0 / 12.0           # RT1
        bc_lo, bc_hi = 2.0 * 1.3 * 4 / 0.7, 0.0  # Dirichlet below, nothing above
        z0, z1 = partition_planes(nz, world)[rank]
        A = local_matrix(c[z0:z1], alpha, off, bc_lo if rank == 0 else 0.0, bc_hi if rank == world - 1 else 0.0)
        T = local_rhs(x0[z0:z1])
        v, E, spikes = rank_publish(A, T, rank > 0, rank < world - 1)
        mine = torch.tensor([E[0], E[1], E[2], v[0], v[-1]], dtype=torch.float64)
        allv = [torch.zeros(5, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allv, mine)
        E_all = [tuple(t[:3].tolist()) for t in allv]
        vG_all = [tuple(t[3:].tolist()) for t in allv]
        JG = reduced_solve(E_all, vG_all)
        J = rank_finish(rank, world, v, E, spikes, JG)
        # global reference
        Ag = local_matrix(c, alpha, off, bc_lo, bc_hi)
        Jg = np.linalg.solve(Ag, local_rhs(x0))
        err = float(np.max(np.abs(J - Jg[z0:z1 + 1])) / np.max(np.abs(Jg)))
        # x^T (B A^-1 B^T) x: local formula sum z^2/m + v_Gamma . lambda summed over ranks == T^T J globally
        t = torch.tensor([float(T @ J)], dtype=torch.float64)
        dist.all_reduce(t)
        quad_err = abs(t.item() - float(local_rhs(x0) @ Jg)) / abs(float(local_rhs(x0) @ Jg))
        q.put((rank, err, quad_err))
    finally:
        dist.destroy_process_group()


def test_substructured_line_solve_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, qerr in res:
        assert err < 1e-12 and qerr < 1e-12, (rank, err, qerr)


def test_substructuring_many_ranks_single_process():
    from slab_model import local_matrix, local_rhs, rank_finish, rank_publish, reduced_solve
    from neutfem_b200.slab import partition_planes
    rng = np.random.default_rng(1)
    for nz, P, (alpha, off) in [(40, 8, (2 / 3, 1 / 3)), (9, 4, (0.25, -1 / 12)), (8, 8, (2 / 15, 1 / 30)), (7, 3, (2 / 3, 1 / 3))]:
        c = rng.uniform(0.1, 5.0, nz); x0 = rng.uniform(-1, 1, nz)
        parts = partition_planes(nz, P)
        assert parts[0][0] == 0 and parts[-1][1] == nz and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        pub = []
        for r, (z0, z1) in enumerate(parts):
            A = local_matrix(c[z0:z1], alpha, off, 3.0 if r == 0 else 0.0, 1.5 if r == P - 1 else 0.0)
            pub.append(rank_publish(A, local_rhs(x0[z0:z1]), r > 0, r < P - 1))
        JG = reduced_solve([p[1] for p in pub], [(p[0][0], p[0][-1]) for p in pub])
        Jg = np.linalg.solve(local_matrix(c, alpha, off, 3.0, 1.5), local_rhs(x0))
        for r, (z0, z1) in enumerate(parts):
            J = rank_finish(r, P, pub[r][0], pub[r][1], pub[r][2], JG)
            assert np.max(np.abs(J - Jg[z0:z1 + 1])) < 1e-12 * max(1.0, np.max(np.abs(Jg)))


def test_neighbour_mode_equals_full_interface_solve_for_thick_slabs():
    """Slabs of >= 30 planes: the coupling between a slab's two interfaces is below 1e-20 and the neighbour-exchange formulas
    (v_n of the rank below, v_0 of the rank above only) reproduce the global line solve to rounding; thin slabs do not qualify."""
    from slab_model import coupling, local_matrix, local_rhs, neighbour_multipliers, rank_publish
    from neutfem_b200.slab import partition_planes
    rng = np.random.default_rng(2)
    for nz, P, (alpha, off), thick in [(4 * 34, 4, (0.25, -1 / 12), True), (3 * 45 + 1, 3, (2 / 3, 1 / 3), True), (4 * 30, 4, (2 / 15, 1 / 30), True),
                                       (4 * 6, 4, (0.25, -1 / 12), False)]:
        c = rng.uniform(0.1, 5.0, nz); x0 = rng.uniform(-1, 1, nz)
        parts = partition_planes(nz, P)
        pub = []
        for r, (z0, z1) in enumerate(parts):
            A = local_matrix(c[z0:z1], alpha, off, 3.0 if r == 0 else 0.0, 1.5 if r == P - 1 else 0.0)
            pub.append(rank_publish(A, local_rhs(x0[z0:z1]), r > 0, r < P - 1))
        cmax = max(coupling(p[1]) for p in pub)
        assert (cmax < 1e-20) == thick
        if not thick:
            continue
        Jg = np.linalg.solve(local_matrix(c, alpha, off, 3.0, 1.5), local_rhs(x0))
        for r, (z0, z1) in enumerate(parts):
            v, E, (s0, sn) = pub[r]
            below = (pub[r - 1][1][2], pub[r - 1][0][-1]) if r > 0 else None
            above = (pub[r + 1][1][0], pub[r + 1][0][0]) if r < P - 1 else None
            lam0, lamn = neighbour_multipliers(r, P, E, (v[0], v[-1]), below, above)
            J = v + s0 * lam0 + sn * lamn
            assert np.max(np.abs(J - Jg[z0:z1 + 1])) < 1e-13 * max(1.0, np.max(np.abs(Jg)))
            # the part of column 0 of the local inverse that the kernels stop loading really is below the threshold
            cut = next((f for f in range(s0.size) if np.all(np.abs(s0[f:]) < 1e-22 * abs(s0[0]))), s0.size)
            assert cut < s0.size


def _nb_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from slab_model import local_matrix, local_rhs, neighbour_multipliers, rank_publish
        from neutfem_b200.slab import partition_planes
        rng = np.random.default_rng(6)
        nz = 2 * 36 + 1
        c = rng.uniform(0.2, 3.0, nz); x0 = rng.uniform(0.5, 1.5, nz)
        alpha, off = 0.25, -1.0 / 12.0
        z0, z1 = partition_planes(nz, world)[rank]
        A = local_matrix(c[z0:z1], alpha, off, 5.0 if rank == 0 else 0.0, 0.0)
        v, E, (s0, sn) = rank_publish(A, local_rhs(x0[z0:z1]), rank > 0, rank < world - 1)
        # neighbour exchange only: (G_nn, v_n) goes up, (G_00, v_0) goes down
        up = torch.tensor([E[2], v[-1]], dtype=torch.float64); down = torch.tensor([E[0], v[0]], dtype=torch.float64)
        got = torch.zeros(2, dtype=torch.float64)
        if rank == 0:
            dist.send(up, dst=1); dist.recv(got, src=1)
            lam0, lamn = neighbour_multipliers(0, world, E, (v[0], v[-1]), None, tuple(got.tolist()))
        else:
            dist.recv(got, src=0); dist.send(down, dst=0)
            lam0, lamn = neighbour_multipliers(1, world, E, (v[0], v[-1]), tuple(got.tolist()), None)
        J = v + s0 * lam0 + sn * lamn
        Jg = np.linalg.solve(local_matrix(c, alpha, off, 5.0, 0.0), local_rhs(x0))
        q.put((rank, float(np.max(np.abs(J - Jg[z0:z1 + 1])) / np.max(np.abs(Jg)))))
    finally:
        dist.destroy_process_group()


def test_neighbour_exchange_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + ((os.getpid() + 777) % 2000)
    procs = [ctx.Process(target=_nb_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-13, (rank, err)
