"""Property tests (hypothesis) of the host-side index logic: slab partition of the major axis, iteration-order strides
(rectilinear3d.cpp:20-32) and the localisation of node lists used for Dirichlet and boundary conditions in slab mode."""
import numpy as np
from hypothesis import given, settings, strategies as st

from plask_b200 import configs as cf


@settings(max_examples=200, deadline=None)
@given(st.integers(min_value=1, max_value=8), st.integers(min_value=0, max_value=400))
def test_slab_partition_covers_the_axis_once(nranks, extra):
    nK = nranks + extra                       # at least one owned plane per rank
    owned = []
    for r in range(nranks):
        lo, hi, own_lo, own_hi = cf.slab_local(nK, r, nranks)
        assert 0 <= lo < hi <= nK and 0 <= own_lo < own_hi <= hi - lo
        assert own_lo == (1 if r > 0 else 0) and (hi - lo) - own_hi == (1 if r < nranks - 1 else 0)   # one halo plane per neighbour
        owned.append((lo + own_lo, lo + own_hi))
    assert owned[0][0] == 0 and owned[-1][1] == nK
    assert all(owned[r][1] == owned[r + 1][0] for r in range(nranks - 1))
    sizes = [b - a for a, b in owned]
    assert max(sizes) - min(sizes) <= 1       # balanced


@settings(max_examples=100, deadline=None)
@given(st.tuples(st.integers(2, 9), st.integers(2, 9), st.integers(2, 9)), st.sampled_from(sorted(cf.ORDERS)))
def test_strides_are_the_iteration_order(n, order):
    ns, es = cf.strides_for(n, order)
    major, medium, minor = cf.ORDERS[order]
    assert ns[minor] == 1 and ns[medium] == n[minor] and ns[major] == n[minor] * n[medium]
    assert es[minor] == 1 and es[medium] == n[minor] - 1 and es[major] == (n[minor] - 1) * (n[medium] - 1)
    idx = sorted(i0 * ns[0] + i1 * ns[1] + i2 * ns[2] for i0 in range(n[0]) for i1 in range(n[1]) for i2 in range(n[2]))
    assert idx == list(range(n[0] * n[1] * n[2]))                     # a bijection onto 0..N-1
    # setOptimalIterationOrder (rectilinear3d.cpp:74-85): the largest axis is the slowest
    opt = cf.optimal_order(n)
    mj, md, mn = cf.ORDERS[opt]
    assert n[mn] <= n[md] <= n[mj]


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 4), st.integers(0, 6), st.sampled_from(["012", "021"]), st.integers(0, 2 ** 31 - 1))
def test_slab_local_nodes_is_the_restriction(nranks, extra, order, seed):
    n = (nranks + 2 + extra, 4, 5)
    p = cf.config_A(n, order=order)
    rng = np.random.default_rng(seed)
    nodes = rng.choice(p.N, size=min(p.N, 40), replace=False)
    ng = np.broadcast_to(p.node_index_grid(), p.n)
    seen = np.zeros(p.N, dtype=int)
    for r in range(nranks):
        q, own_lo, own_hi, (lo, hi) = cf.slab_problem(p, r, nranks)
        ln, keep = cf.slab_local_nodes(p, q, lo, hi, nodes)
        lng = np.broadcast_to(q.node_index_grid(), q.n)
        # the local node with local coordinates (i0, i1, i2) is the global node (lo + i0, i1, i2)
        back = {int(lng[i0, i1, i2]): int(ng[lo + i0, i1, i2]) for i0 in range(q.n[0]) for i1 in range(q.n[1]) for i2 in range(q.n[2])}
        assert [back[int(v)] for v in ln] == [int(v) for v in nodes[keep]]
        owned_global = {int(ng[lo + i0, i1, i2]) for i0 in range(own_lo, own_hi) for i1 in range(q.n[1]) for i2 in range(q.n[2])}
        for v in nodes[keep]:
            if int(v) in owned_global:
                seen[int(v)] += 1
    assert np.all(seen[nodes] == 1)           # every listed node is owned by exactly one rank


@settings(max_examples=200, deadline=None)
@given(st.integers(min_value=1, max_value=8), st.integers(min_value=0, max_value=600), st.sampled_from([4, 16]))
def test_aligned_slab_partition(nranks, extra, align):
    """slab_range(align): what the multilevel preconditioner asks for — the boundaries between ranks are multiples of `align`
    planes, the ranges cover the axis once, the blocks of `align` planes are balanced"""
    nK = align * (nranks - 1) + 1 + extra     # enough planes for one block per rank
    owned = [cf.slab_range(nK, r, nranks, align) for r in range(nranks)]
    assert owned[0][0] == 0 and owned[-1][1] == nK
    assert all(owned[r][1] == owned[r + 1][0] for r in range(nranks - 1))
    assert all(owned[r][1] % align == 0 for r in range(nranks - 1))
    assert all(b > a for a, b in owned)
    blocks = [-(-(b - a) // align) for a, b in owned]
    assert max(blocks) - min(blocks) <= 1
    # aggregate rows of the library: no 4x4 / 16x16 aggregate row holds planes of two ranks
    if align == 16:
        for r in range(nranks):
            lo, hi, own_lo, own_hi = cf.slab_local(nK, r, nranks, align)
            koff = 16 - own_lo
            for sh in (2, 4):
                mine = {(k + koff) >> sh for k in range(own_lo, own_hi)}
                halo = {(k + koff) >> sh for k in (own_lo - 1, own_hi) if 0 <= k < hi - lo}
                assert not (mine & halo)
