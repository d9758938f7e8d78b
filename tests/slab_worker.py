"""Worker of the slab-mode (multi-GPU) tests, one process per rank:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/slab_worker.py [--host-only]

--host-only (CPU, gloo): checks the host-side slab logic — partition, halo planes, local Dirichlet lists, blob
all-gather plumbing — without touching a GPU.
Default (needs one GPU per rank): solves a small config-B problem in slab mode through the C ABI and compares it on
rank 0 with the single-GPU solve of the global problem and with the oracle's Cholesky.
"""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from plask_b200 import configs as cf  # noqa: E402


def allgather_bytes(b):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, b)
    return out


def main():
    host_only = "--host-only" in sys.argv
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = (8 * world + 3, 14, 12)                       # major axis = axis 0, uneven split
    p = cf.config_B(n)
    assert cf.ORDERS[p.order][0] == 0
    q, own_lo, own_hi, (lo, hi) = cf.slab_problem(p, rank, world)

    # ---- host logic: partition covers the axis, halos overlap the neighbours' owned planes
    ranges = allgather_bytes((lo, hi, own_lo, own_hi))
    owned = [(l + a, l + b) for (l, h, a, b) in ranges]
    assert owned[0][0] == 0 and owned[-1][1] == n[0]
    assert all(owned[r][1] == owned[r + 1][0] for r in range(world - 1))
    assert own_lo == (1 if rank > 0 else 0) and (hi - lo) - own_hi == (1 if rank < world - 1 else 0)
    # local arrays are the slices of the global ones
    eg = np.broadcast_to(p.elem_index_grid(), tuple(k - 1 for k in n))
    leg = np.broadcast_to(q.elem_index_grid(), tuple(k - 1 for k in q.n))
    assert np.array_equal(q.elem_mat[leg], p.elem_mat[eg][lo:hi - 1])
    assert np.array_equal(q.heat[leg], p.heat[eg][lo:hi - 1])
    ng = np.broadcast_to(p.node_index_grid(), n)
    lng = np.broadcast_to(q.node_index_grid(), q.n)
    gfix = np.zeros(p.N, bool); gfix[p.bc_nodes] = True
    lfix = np.zeros(q.N, bool); lfix[q.bc_nodes] = True
    assert np.array_equal(lfix[lng], gfix[ng][lo:hi])
    # direct construction of the slab (what bench.py does at scale) gives the same local problem
    d = cf.config_B(n, rows0=(lo, hi))
    assert np.array_equal(d.elem_mat, q.elem_mat) and np.array_equal(d.heat, q.heat)
    assert np.array_equal(np.sort(d.bc_nodes), np.sort(q.bc_nodes))
    # Shockley problem: junction / role arrays are cut the same way and the local junction description is consistent
    pc = cf.config_C((8 * world + 5, 22, 52), order="012")
    qc, c_lo, c_hi, (clo, chi) = cf.slab_problem(pc, rank, world)
    egc = np.broadcast_to(pc.elem_index_grid(), tuple(k - 1 for k in pc.n))
    legc = np.broadcast_to(qc.elem_index_grid(), tuple(k - 1 for k in qc.n))
    assert np.array_equal(qc.elem_junc[legc], pc.elem_junc[egc][clo:chi - 1])
    assert np.array_equal(qc.noheat[legc], pc.noheat[egc][clo:chi - 1])
    dc = cf.config_C(pc.n, order="012", rows0=(clo, chi))      # direct construction (tools/run_configs.py at scale)
    assert np.array_equal(dc.elem_mat, qc.elem_mat) and np.array_equal(dc.elem_junc, qc.elem_junc)
    ia, ib = np.argsort(dc.bc_nodes), np.argsort(qc.bc_nodes)
    assert np.array_equal(dc.bc_nodes[ia], qc.bc_nodes[ib]) and np.array_equal(dc.bc_values[ia], qc.bc_values[ib])
    acts_g, _ = cf.setup_active(pc)
    acts_l, ncol_l = cf.setup_active(qc)
    if ncol_l:
        assert acts_l[0]["bottom"] == acts_g[0]["bottom"] and acts_l[0]["top"] == acts_g[0]["top"]
        assert acts_l[0]["back"] + clo >= acts_g[0]["back"] and acts_l[0]["front"] + clo <= acts_g[0]["front"]
    # node lists of boundary conditions are localised like the Dirichlet list
    ln, keep = cf.slab_local_nodes(p, q, lo, hi, p.bc_nodes)
    assert np.array_equal(ln.astype(np.uintp), q.bc_nodes) and keep.sum() == q.bc_nodes.size
    # a mesh in the generator's optimal order with the vertical axis major (rectilinear3d.cpp:74-85): the host cuts a LATERAL axis
    # instead (cf.slab_axis) and numbers the local meshes with that axis major — junctions and vertical lines stay inside the slabs
    pv = cf.config_C((6 * world + 5, 14, 52), order="201")
    assert cf.ORDERS[pv.order][0] == 2 and cf.slab_axis(pv) == 0
    qv, v_lo, v_hi, (vlo, vhi) = cf.slab_problem(pv, rank, world, axis=cf.slab_axis(pv))
    assert qv.order == "021" and qv.n == (vhi - vlo, 14, 52)
    egv = np.broadcast_to(pv.elem_index_grid(), tuple(k - 1 for k in pv.n))
    legv = np.broadcast_to(qv.elem_index_grid(), tuple(k - 1 for k in qv.n))
    assert np.array_equal(qv.elem_junc[legv], pv.elem_junc[egv][vlo:vhi - 1])
    assert np.array_equal(qv.elem_mat[legv], pv.elem_mat[egv][vlo:vhi - 1])
    gfv = np.zeros(pv.N); gfv[pv.bc_nodes] = pv.bc_values + 7.
    lfv = np.zeros(qv.N); lfv[qv.bc_nodes] = qv.bc_values + 7.
    assert np.array_equal(lfv[np.broadcast_to(qv.node_index_grid(), qv.n)], gfv[np.broadcast_to(pv.node_index_grid(), pv.n)][vlo:vhi])
    own = allgather_bytes(cf.slab_field_owned(qv, lfv, v_lo, v_hi))
    assert np.array_equal(cf.slab_assemble(pv, qv.order, own), gfv)
    blobs = allgather_bytes(bytes([rank]) * 64)
    assert [b[0] for b in blobs] == list(range(world))
    if host_only:
        dist.barrier()
        if rank == 0:
            print("slab host logic ok, world", world)
        dist.destroy_process_group()
        return

    # ---- device: collective solve in slab mode
    import torch
    from plask_b200.solvers import Static3D
    local = int(os.environ.get("LOCAL_RANK", rank))
    assert torch.cuda.device_count() > local, "one GPU per rank is needed"
    s = Static3D(f"slab{rank}")
    s.device = local
    s.problem = q
    s.slab = dict(rank=rank, nranks=world, own_lo=own_lo, own_hi=own_hi, allgather=allgather_bytes)
    s.iterative.maxerr = 1e-11
    s.iterative.maxit = 50000
    err = s.compute(0)
    T_loc = s.outTemperature()
    parts = allgather_bytes((cf.slab_field_owned(q, T_loc, own_lo, own_hi), s.stats, err))
    # halo planes returned by get_field agree with the neighbours' owned planes
    planes = T_loc.reshape(q.n[0], q.n[1], q.n[2])
    if rank > 0:
        assert np.array_equal(planes[0], parts[rank - 1][0][-1])
    if rank < world - 1:
        assert np.array_equal(planes[-1], parts[rank + 1][0][0])
    stats = [x[1] for x in parts]
    assert all(st["lin_iters"] == stats[0]["lin_iters"] and st["outer_loops"] == stats[0]["outer_loops"] for st in stats)
    assert all(x[2] == parts[0][2] for x in parts)
    # provider on a foreign mesh in slab mode (collective): every rank interpolates on its local mesh, the host keeps the points
    # whose coordinate along the slab axis lies in the rank's owned interval
    tgt_axes = [np.concatenate([[a[0] - 0.3], 0.5 * (a[:-1] + a[1:]) + 0.01, [a[-1] + 0.2]]) for a in p.axes]
    Ti_loc = s.outTemperature((tgt_axes, "102")).reshape(len(tgt_axes[1]), len(tgt_axes[0]), len(tgt_axes[2]))   # order 102: axis 1 slowest
    x_lo = -np.inf if rank == 0 else q.axes[0][own_lo]
    x_hi = np.inf if rank == world - 1 else q.axes[0][own_hi]
    mine = (tgt_axes[0] >= x_lo) & (tgt_axes[0] < x_hi)
    partsI = allgather_bytes((Ti_loc[:, mine, :], mine))
    s.invalidate()
    if rank == 0:
        from helpers import oracle_thermal
        T = np.concatenate([x[0] for x in parts], axis=0).ravel()      # order 012: plain C order of (n0, n1, n2)
        one = Static3D("single")
        one.device = 0
        one.problem = p
        one.iterative.maxerr = 1e-11
        one.iterative.maxit = 50000
        one.compute(0)
        T1 = one.outTemperature()
        o = oracle_thermal(p, algorithm="cholesky")
        o.compute(0)
        Ti = np.empty((len(tgt_axes[1]), len(tgt_axes[0]), len(tgt_axes[2])))
        cover = np.zeros(len(tgt_axes[0]), dtype=int)
        for vals, m in partsI:
            Ti[:, m, :] = vals
            cover += m
        assert np.all(cover == 1)
        Ti1 = one.outTemperature((tgt_axes, "102")).reshape(Ti.shape)
        di = float(np.abs(Ti - Ti1).max())
        print(f"slab x{world}: outTemperature on a foreign mesh, max|Ti_slab - Ti_single| = {di:.3e} K")
        assert di <= 1e-6
        d1 = float(np.abs(T - T1).max())
        d2 = float(np.abs(T - o.temperatures).max())
        print(f"slab x{world}: loops {stats[0]['outer_loops']}, PCG iterations {stats[0]['lin_iters']} (single GPU "
              f"{one.stats['lin_iters']}), max|T_slab - T_single| = {d1:.3e} K, max|T_slab - T_cholesky| = {d2:.3e} K")
        assert stats[0]["outer_loops"] == one.stats["outer_loops"] == len(o.history)
        assert abs(stats[0]["lin_iters"] - one.stats["lin_iters"]) <= 0.02 * one.stats["lin_iters"] + 2
        assert d1 <= 1e-6 and d2 <= 1e-3
        one.invalidate()
    dist.barrier()

    # ---- line-Jacobi preconditioner in slab mode (order 012: the vertical lines run along the minor axis, inside every slab)
    sl = Static3D(f"slabline{rank}")
    sl.device = local
    sl.problem = q
    sl.slab = dict(rank=rank, nranks=world, own_lo=own_lo, own_hi=own_hi, allgather=allgather_bytes)
    sl.iterative.preconditioner = "ljac"
    sl.iterative.maxerr = 1e-11
    sl.iterative.maxit = 50000
    sl.compute(0)
    partsL = allgather_bytes((cf.slab_field_owned(q, sl.outTemperature(), own_lo, own_hi), sl.stats))
    assert all(x[1]["lin_iters"] == partsL[0][1]["lin_iters"] for x in partsL)
    sl.invalidate()
    if rank == 0:
        TL = np.concatenate([x[0] for x in partsL], axis=0).ravel()
        d3 = float(np.abs(TL - T).max())
        print(f"slab x{world} line-Jacobi: PCG iterations {partsL[0][1]['lin_iters']} (Jacobi {stats[0]['lin_iters']}), "
              f"max|Tl_slab - T_slab| = {d3:.3e} K")
        assert partsL[0][1]["outer_loops"] == stats[0]["outer_loops"]
        assert partsL[0][1]["lin_iters"] < stats[0]["lin_iters"]
        assert d3 <= 1e-6
    dist.barrier()

    # ---- the same with the vertical axis as the MEDIUM axis of the mesh (order 021): the 'auto' layout stores it fastest
    # (pfem_set_layout), the slab axis stays the slowest one
    p21 = cf.config_B(n, order="021")
    q21, o_lo, o_hi, _ = cf.slab_problem(p21, rank, world)
    s21 = Static3D(f"slab021_{rank}")
    s21.device = local
    s21.problem = q21
    s21.slab = dict(rank=rank, nranks=world, own_lo=o_lo, own_hi=o_hi, allgather=allgather_bytes)
    s21.iterative.preconditioner = "ljac"
    s21.iterative.maxerr = 1e-11
    s21.iterative.maxit = 50000
    s21.compute(0)
    parts21 = allgather_bytes((cf.slab_field_owned(q21, s21.outTemperature(), o_lo, o_hi), s21.stats))
    s21.invalidate()
    if rank == 0:
        T21 = np.concatenate([x[0] for x in parts21], axis=0).ravel()
        one21 = Static3D("single021")
        one21.device = 0
        one21.problem = p21
        one21.iterative.preconditioner = "ljac"
        one21.iterative.maxerr = 1e-11
        one21.iterative.maxit = 50000
        one21.compute(0)
        d4 = float(np.abs(T21 - one21.outTemperature()).max())
        print(f"slab x{world} order 021 line-Jacobi (vertical-minor layout): PCG iterations {parts21[0][1]['lin_iters']} "
              f"(single GPU {one21.stats['lin_iters']}), max|T021_slab - T021_single| = {d4:.3e} K")
        assert parts21[0][1]["outer_loops"] == one21.stats["outer_loops"]
        assert d4 <= 1e-6
        one21.invalidate()
    dist.barrier()

    # ---- masked mesh (empty-elements="exclude") in slab mode: the air around the mesa is dropped on every rank
    sm = Static3D(f"slabmask{rank}")
    sm.device = local
    sm.problem = q
    sm.slab = dict(rank=rank, nranks=world, own_lo=own_lo, own_hi=own_hi, allgather=allgather_bytes)
    sm.empty_elements = "exclude"
    sm.iterative.maxerr = 1e-11
    sm.iterative.maxit = 50000
    sm.compute(0)
    partsM = allgather_bytes((cf.slab_field_owned(q, sm.outTemperature(), own_lo, own_hi), sm.stats))
    sm.invalidate()
    if rank == 0:
        TM = np.concatenate([x[0] for x in partsM], axis=0).ravel()
        onem = Static3D("singlemask")
        onem.device = 0
        onem.problem = p
        onem.empty_elements = "exclude"
        onem.iterative.maxerr = 1e-11
        onem.iterative.maxit = 50000
        onem.compute(0)
        d5 = float(np.abs(TM - onem.outTemperature()).max())
        print(f"slab x{world} masked mesh: PCG iterations {partsM[0][1]['lin_iters']} (single GPU {onem.stats['lin_iters']}), "
              f"max|Tm_slab - Tm_single| = {d5:.3e} K, nodes outside the masked mesh {int((~onem.masked_nodes()).sum())}")
        assert partsM[0][1]["outer_loops"] == onem.stats["outer_loops"]
        assert d5 <= 1e-6
        onem.invalidate()
    dist.barrier()

    # ---- boundary conditions of the 2nd / 3rd kind and radiation in slab mode (corrected form; conditions on the two
    # end planes of the slab axis live on one rank only, the others cross every slab)
    from helpers import face_nodes
    gconds = dict(convection=[(face_nodes(p, 2, -1), 4.0e4, 310.), (face_nodes(p, 0, 0), 9.0e4, 295.)],
                  heatflux=[(face_nodes(p, 0, -1), -3.0e5)], radiation=[(face_nodes(p, 1, -1), 0.85, 285.)])

    def localise(conds):
        out = {}
        for kind, lst in conds.items():
            out[kind] = []
            for c in lst:
                ln, keep = cf.slab_local_nodes(p, q, lo, hi, c[0])
                out[kind].append((ln,) + tuple(c[1:]))
        return out

    def with_conditions(sv, conds):
        sv.heatflux_boundary, sv.convection_boundary, sv.radiation_boundary = conds["heatflux"], conds["convection"], conds["radiation"]
        sv.boundary_verbatim = False
        sv.iterative.maxerr = 1e-11
        sv.iterative.maxit = 50000

    sb = Static3D(f"slabbc{rank}")
    sb.device = local
    sb.problem = q
    sb.slab = dict(rank=rank, nranks=world, own_lo=own_lo, own_hi=own_hi, allgather=allgather_bytes)
    with_conditions(sb, localise(gconds))
    sb.compute(0)
    partsB = allgather_bytes((cf.slab_field_owned(q, sb.outTemperature(), own_lo, own_hi), sb.stats))
    sb.invalidate()
    if rank == 0:
        from oracle import oracle as orc
        from helpers import oracle_thermal
        T = np.concatenate([x[0] for x in partsB], axis=0).ravel()
        o = oracle_thermal(p, algorithm="cholesky", boundaries=orc.BoundaryTerms(p.N, **gconds), quirk=False)
        o.compute(0)
        one = Static3D("singlebc")
        one.device = 0
        one.problem = p
        with_conditions(one, gconds)
        one.compute(0)
        d1 = float(np.abs(T - one.outTemperature()).max())
        d2 = float(np.abs(T - o.temperatures).max())
        print(f"slab x{world} with boundary terms: loops {partsB[0][1]['outer_loops']}, PCG iterations {partsB[0][1]['lin_iters']} "
              f"(single GPU {one.stats['lin_iters']}), max|Tb_slab - Tb_single| = {d1:.3e} K, max|Tb_slab - Tb_cholesky| = {d2:.3e} K")
        assert partsB[0][1]["outer_loops"] == one.stats["outer_loops"] == len(o.history)
        assert d1 <= 1e-6 and d2 <= 1e-3
        one.invalidate()
    dist.barrier()

    # ---- Shockley3D in slab mode: fixed number of loops, against the single-GPU solve and the oracle's Cholesky
    from plask_b200.solvers import Shockley3D
    LOOPS = 6

    def shockley(name, prob, slab=None, dev=local):
        e = Shockley3D(name)
        e.device = dev
        e.problem = prob
        e.slab = slab
        e.beta, e.js, e.maxerr = prob.beta, prob.js, prob.maxerr
        e.iterative.maxerr = 1e-12
        e.iterative.maxit = 50000
        e.compute(LOOPS)
        return e

    e = shockley(f"eslab{rank}", qc, dict(rank=rank, nranks=world, own_lo=c_lo, own_hi=c_hi, allgather=allgather_bytes))
    V_loc = e.outVoltage()
    heat_loc = e.outHeat()
    partsV = allgather_bytes((cf.slab_field_owned(qc, V_loc, c_lo, c_hi), e.stats, e.maxcur))
    assert all(x[1]["err"] == partsV[0][1]["err"] and x[1]["maxval"] == partsV[0][1]["maxval"] for x in partsV)
    assert all(tuple(x[2]) == tuple(partsV[0][2]) for x in partsV)
    e.invalidate()
    if rank == 0:
        from helpers import oracle_shockley
        V = np.concatenate([x[0] for x in partsV], axis=0).ravel()
        one = shockley("esingle", pc, None, 0)
        V1 = one.outVoltage()
        oe = oracle_shockley(pc, algorithm="cholesky")
        oe.compute(LOOPS)
        d1 = float(np.abs(V - V1).max())
        d2 = float(np.abs(V - oe.potential).max())
        st = partsV[0][1]
        print(f"slab x{world} Shockley: {LOOPS} loops, err {st['err']:.6g}% (single {one.stats['err']:.6g}%, oracle {oe.history[-1]['err']:.6g}%), "
              f"max|V_slab - V_single| = {d1:.3e} V, max|V_slab - V_cholesky| = {d2:.3e} V")
        assert d1 <= 1e-9 and d2 <= 1e-6
        assert abs(st["err"] - one.stats["err"]) <= 1e-6 * max(1., abs(one.stats["err"]))
        assert abs(st["maxval"] - one.stats["maxval"]) <= 1e-9 * abs(one.stats["maxval"])
        assert np.allclose(partsV[0][2], one.maxcur, rtol=1e-7, atol=1e-12)
        one.invalidate()
    dist.barrier()

    # ---- Shockley3D on a mesh whose own order has the VERTICAL axis major (order 201), line-Jacobi: lateral cut chosen by the host
    ev = Shockley3D(f"evslab{rank}")
    ev.device = local
    ev.problem = qv
    ev.slab = dict(rank=rank, nranks=world, own_lo=v_lo, own_hi=v_hi, allgather=allgather_bytes)
    ev.beta, ev.js, ev.maxerr = qv.beta, qv.js, qv.maxerr
    ev.iterative.preconditioner = "ljac"
    ev.iterative.maxerr = 1e-12
    ev.iterative.maxit = 50000
    ev.compute(LOOPS)
    partsW = allgather_bytes((cf.slab_field_owned(qv, ev.outVoltage(), v_lo, v_hi), ev.stats))
    ev.invalidate()
    if rank == 0:
        Vg = cf.slab_assemble(pv, qv.order, [x[0] for x in partsW])      # back in the numbering of the global mesh (order 201)
        onev = Shockley3D("evsingle")
        onev.device = 0
        onev.problem = pv
        onev.beta, onev.js, onev.maxerr = pv.beta, pv.js, pv.maxerr
        onev.iterative.preconditioner = "ljac"
        onev.iterative.maxerr = 1e-12
        onev.iterative.maxit = 50000
        onev.compute(LOOPS)
        d6 = float(np.abs(Vg - onev.outVoltage()).max())
        print(f"slab x{world} Shockley, mesh order 201 (vertical major) cut along axis 0, line-Jacobi: PCG iterations {partsW[0][1]['lin_iters']} "
              f"(single GPU {onev.stats['lin_iters']}), max|V201_slab - V201_single| = {d6:.3e} V")
        assert d6 <= 1e-9
        assert abs(partsW[0][1]["err"] - onev.stats["err"]) <= 1e-6 * max(1., abs(onev.stats["err"]))
        onev.invalidate()
    dist.barrier()

    # ---- Dynamic3D in slab mode: 6 Crank-Nicolson steps with a rebuild of k(T), cp(T) every 2 steps, line-Jacobi
    from plask_b200.solvers import Dynamic3D

    def dynamic(name, prob, slab=None, dev=local):
        d = Dynamic3D(name)
        d.device = dev
        d.problem = prob
        d.slab = slab
        d.timestep, d.rebuildfreq, d.logfreq = 20., 2, 1
        d.iterative.preconditioner = "ljac"
        d.iterative.maxerr = 1e-12
        d.iterative.maxit = 50000
        d.compute(100.)
        return d

    ph = cf.config_B(n)
    ph.heat = ph.heat * 200.
    qh, h_lo, h_hi, _ = cf.slab_problem(ph, rank, world)
    dd = dynamic(f"dslab{rank}", qh, dict(rank=rank, nranks=world, own_lo=h_lo, own_hi=h_hi, allgather=allgather_bytes))
    partsD = allgather_bytes((cf.slab_field_owned(qh, dd.outTemperature(), h_lo, h_hi), dd.stats, dd.maxT))
    assert all(x[2] == partsD[0][2] and x[1]["outer_loops"] == 6 for x in partsD)
    dd.invalidate()
    if rank == 0:
        Td = np.concatenate([x[0] for x in partsD], axis=0).ravel()
        oned = dynamic("dsingle", ph, None, 0)
        d7 = float(np.abs(Td - oned.outTemperature()).max())
        print(f"slab x{world} Dynamic3D: 6 steps, PCG iterations {partsD[0][1]['lin_iters']} (single GPU {oned.stats['lin_iters']}), max T {partsD[0][2]:.3f} K, "
              f"max|Td_slab - Td_single| = {d7:.3e} K")
        assert oned.maxT - 300. > 5. and abs(oned.maxT - partsD[0][2]) <= 1e-6 and d7 <= 1e-6
        oned.invalidate()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
