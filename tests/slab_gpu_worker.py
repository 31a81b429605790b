"""Worker of tests/test_gpu_slab.py: one process per GPU, compares the z-slab path with a single-GPU context."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import torch
    import torch.distributed as dist
    from helpers import random_problem, relerr
    from neutfem_b200 import cabi
    from neutfem_b200.slab import SlabSolver

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    ok = True
    msgs = []
    for (n, rt, pp) in [((9, 7, 11), 1, 1), ((6, 5, 2 * world), 2, 2), ((33, 4, 3 * world + 1), 0, 0), ((5, 6, 7), 2, 1),
                        ((16, 9, 3 * world + 2), 1, 1), ((264, 6, 2 * world), 1, 0), ((12, 40, 2 * world + 1), 2, 2),    # even nx: x-row / y-column kernels
                        ((96, 96, 12 * world + 2), 1, 1),    # mid-size, UNEVEN slabs: every collective-affecting decision must be rank-invariant
                        ((16, 10, 32 * world + 1), 1, 1), ((8, 6, 40 * world), 0, 0)]:   # THICK slabs: neighbour-exchange mode of the interface solve
        p = random_problem(17, 3, n, ng=2, bc="mixed")
        p["NSF"] *= 3.0
        nx, ny, nz = n
        # single-GPU reference on every rank (same device)
        full = cabi.Context(rt, pp, 2, p["xb"], p["yb"], p["zb"], device=rank)
        for a, t, v in p["bcs"]:
            full.set_bc(a, t, v)
        full.upload_xs(D=p["D"], SigR=p["SigR"], NSF=p["NSF"], Chi=p["Chi"], SigS=p["SigS"])
        full.build()
        s = SlabSolver(rt, pp, 2, p["xb"], p["yb"], p["zb"], rank, world, rank)
        c = s.ctx
        for a, t, v in p["bcs"]:
            c.set_bc(a, t, v)
        c.upload_xs(D=s.local_planes(p["D"]), SigR=s.local_planes(p["SigR"]), NSF=s.local_planes(p["NSF"]),
                    Chi=s.local_planes(p["Chi"]), SigS=s.local_planes(p["SigS"], 2))
        c.build()
        nl = c.n_phi_loc
        if nx % 2 == 0:
            kt = c.time_kernels(0, 1, True)
            thick = nz >= 32 * world
            msgs.append(f"rank{rank} n={n} RT{rt}P{pp} path {int(kt['path'])} neighbour mode {int(kt['slab_neighbour_mode'])} coupling {kt['slab_coupling']:.1e}")
            ok &= kt["path"] == 5.0 and int(kt["slab_neighbour_mode"]) == (1 if thick else 0)
        x = np.random.default_rng(5).uniform(0.5, 1.5, full.n_Phi)
        lo, hi = s.z0 * nx * ny * nl, s.z1 * nx * ny * nl
        for g in range(2):
            y_ref = full.schur_apply(g, x)
            y = c.schur_apply(g, x[lo:hi])
            e = relerr(y, y_ref[lo:hi])
            msgs.append(f"rank{rank} n={n} RT{rt}P{pp} g{g} apply err {e:.2e}")
            ok &= e < 1e-12
        for mode in (cabi.MODE_PARITY, cabi.MODE_FAST):
            for ctx in (full, c):
                ctx.reset_flux()
                ctx.set_solver(solver_type=cabi.BICGSTAB, tol_keff=1e-10, tol_flux=1e-10, max_outer=400, max_inner=4000, mode=mode)
            k_ref, st_ref = full.solve_keff(False)
            k, st = c.solve_keff(False)
            phi_ref = full.get_flux().reshape(2, -1)[:, lo:hi].ravel()
            e_phi = relerr(c.get_flux(), phi_ref)
            msgs.append(f"rank{rank} n={n} RT{rt}P{pp} mode{mode}: k {k:.12f} vs {k_ref:.12f}, flux err {e_phi:.2e}, outer {st['outer_iterations']}/{st_ref['outer_iterations']} cg {st['cg_iterations']}/{st_ref['cg_iterations']}")
            ok &= abs(k - k_ref) / k_ref < 1e-9 and e_phi < 1e-7
        full.close(); s.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("\n".join(msgs))
    print(f"SLAB_WORKER rank {rank} {'OK' if ok else 'FAIL'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
