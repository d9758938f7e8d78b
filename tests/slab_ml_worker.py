"""Worker of the multilevel-preconditioner slab test (one process per GPU, launched by torchrun from tests/test_slab.py):
the additive multilevel line preconditioner ('mlj') in slab mode against the single-GPU solve.  With slab boundaries at multiples of
16 planes the hierarchy is the one of the whole mesh (same aggregates, same line blocks, ONE top column for the whole device whose
residual and blocks are summed over the ranks), so the iteration counts must agree with the single-GPU run, not just the result."""
import os
import pickle
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plask_b200 import configs as cf  # noqa: E402
from plask_b200.solvers import Shockley3D, Static3D  # noqa: E402


def allgather_bytes(b):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, b)
    return out


def host_checks(rank, world):
    """no GPU: the aligned partition covers the axis, every boundary is a multiple of 16 planes, the local problems are slices of
    the global one (what the multilevel preconditioner relies on: no lateral aggregate straddles two slabs)"""
    n = (32 * world + 5, 20, 24)
    p = cf.config_B(n)
    q, own_lo, own_hi, (lo, hi) = cf.slab_problem(p, rank, world, align=16)
    ranges = allgather_bytes((lo, hi, own_lo, own_hi))
    owned = [(l + a, l + b) for (l, h, a, b) in ranges]
    assert owned[0][0] == 0 and owned[-1][1] == n[0]
    assert all(owned[r][1] == owned[r + 1][0] for r in range(world - 1))
    assert all(owned[r][1] % 16 == 0 for r in range(world - 1)), owned
    assert own_lo == (1 if rank > 0 else 0) and (hi - lo) - own_hi == (1 if rank < world - 1 else 0)
    eg = np.broadcast_to(p.elem_index_grid(), tuple(k - 1 for k in n))
    leg = np.broadcast_to(q.elem_index_grid(), tuple(k - 1 for k in q.n))
    assert np.array_equal(q.elem_mat[leg], p.elem_mat[eg][lo:hi - 1]) and np.array_equal(q.heat[leg], p.heat[eg][lo:hi - 1])
    # aggregate rows as the library counts them: (k + koff) >> 2 with koff = 16 - own_lo; halo planes fall into rows of their own
    koff = 16 - own_lo
    rows_owned = {(k + koff) >> 2 for k in range(own_lo, own_hi)}
    halos = [k for k in (own_lo - 1, own_hi) if 0 <= k < hi - lo]
    assert all(((k + koff) >> 2) not in rows_owned for k in halos)
    assert all(((k + koff) >> 4) not in {(j + koff) >> 4 for j in range(own_lo, own_hi)} for k in halos)
    # Shockley problem cut along a lateral axis although its own order has the vertical axis major
    pc = cf.config_C((16 * world + 7, 22, 52))
    axis = cf.slab_axis(pc, need_vertical_inside=True)
    assert axis in (0, 1)
    qc, c_lo, c_hi, (clo, chi) = cf.slab_problem(pc, rank, world, axis=axis, align=16)
    assert (clo + c_hi) % 16 == 0 or rank == world - 1
    if rank == 0:
        print("slab multilevel host logic ok")


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    host_checks(rank, world)
    if "--host-only" in sys.argv:
        dist.barrier()
        dist.destroy_process_group()
        return
    import torch
    local = rank % max(torch.cuda.device_count(), 1)

    def run(cls, name, prob, slab, dev, loops, setup=None):
        s = cls(name)
        s.device = dev
        s.problem = prob
        s.slab = slab
        s.iterative.preconditioner = "mlj"
        s.iterative.maxerr = 1e-11
        s.iterative.maxit = 50000
        if setup:
            setup(s)
        s.compute(loops)
        return s

    # ---- Static3D, config B: 32 planes per rank but the last (which gets 37), vertical axis minor
    n = (32 * world + 5, 20, 24)
    p = cf.config_B(n)
    q, own_lo, own_hi, _ = cf.slab_problem(p, rank, world, align=16)
    slab = dict(rank=rank, nranks=world, own_lo=own_lo, own_hi=own_hi, allgather=allgather_bytes)
    s = run(Static3D, f"mlslab{rank}", q, slab, local, 0)
    parts = allgather_bytes((cf.slab_field_owned(q, s.outTemperature(), own_lo, own_hi), s.stats))
    s.invalidate()
    assert all(x[1]["lin_iters"] == parts[0][1]["lin_iters"] for x in parts)
    if rank == 0:
        T = np.concatenate([x[0] for x in parts], axis=0).ravel()
        one = run(Static3D, "mlsingle", p, None, 0, 0)
        d = float(np.abs(T - one.outTemperature()).max())
        print(f"slab x{world} multilevel: PCG iterations {parts[0][1]['lin_iters']} (single GPU {one.stats['lin_iters']}), "
              f"max|Tml_slab - Tml_single| = {d:.3e} K")
        assert parts[0][1]["outer_loops"] == one.stats["outer_loops"]
        assert abs(parts[0][1]["lin_iters"] - one.stats["lin_iters"]) <= 0.03 * one.stats["lin_iters"] + 3
        assert d <= 1e-6
        one.invalidate()
    dist.barrier()

    # ---- Shockley3D, config C (junction, contacts): lateral axis 0 cut, 16 planes on the first ranks
    pc = cf.config_C((16 * world + 7, 22, 52), order="012")
    qc, c_lo, c_hi, _ = cf.slab_problem(pc, rank, world, align=16)
    slabc = dict(rank=rank, nranks=world, own_lo=c_lo, own_hi=c_hi, allgather=allgather_bytes)

    def shockley_setup(prob):
        def f(e):
            e.beta, e.js, e.maxerr = prob.beta, prob.js, prob.maxerr
        return f
    e = run(Shockley3D, f"mlshock{rank}", qc, slabc, local, 5, shockley_setup(pc))
    partsV = allgather_bytes((cf.slab_field_owned(qc, e.outVoltage(), c_lo, c_hi), e.stats))
    e.invalidate()
    if rank == 0:
        V = np.concatenate([x[0] for x in partsV], axis=0).ravel()
        onev = run(Shockley3D, "mlshock_single", pc, None, 0, 5, shockley_setup(pc))
        dv = float(np.abs(V - onev.outVoltage()).max())
        print(f"slab x{world} multilevel Shockley3D: PCG iterations {partsV[0][1]['lin_iters']} (single GPU {onev.stats['lin_iters']}), "
              f"max|Vml_slab - Vml_single| = {dv:.3e} V")
        assert abs(partsV[0][1]["lin_iters"] - onev.stats["lin_iters"]) <= 0.03 * onev.stats["lin_iters"] + 5
        assert dv <= 1e-7
        onev.invalidate()
    dist.barrier()

    # ---- a partition that is not aligned is refused with a clear message
    qb, b_lo, b_hi, _ = cf.slab_problem(p, rank, world)          # 32 w + 5 planes split evenly: not a multiple of 16
    if (b_hi - b_lo) % 16 != 0:
        bad = Static3D(f"mlbad{rank}")
        bad.device = local
        bad.problem = qb
        bad.slab = dict(rank=rank, nranks=world, own_lo=b_lo, own_hi=b_hi, allgather=allgather_bytes)
        bad.iterative.preconditioner = "mlj"
        failed = False
        try:
            bad.compute(1)
        except Exception as err:       # the refusal is collective: every rank raises, also the ones whose own slab is fine
            failed = "multiple of 16" in str(err)
        flags = allgather_bytes(failed)
        assert all(flags), flags
        bad.invalidate()
    if rank == 0:
        print("slab multilevel ok")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
