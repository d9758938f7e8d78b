"""DeviceFem — thin object wrapper over one pfem_ctx (one CUDA device, one mesh).

This is the level at which the PLaSK solver plugin would talk to the library (see
INTEGRATION.md); `plask_b200.solvers` builds the Static3D / Shockley3D mirrors on top of it.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def _dp(a):
    return a.ctypes.data_as(L.c_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class DeviceFem:
    def __init__(self, device=0):
        self.lib = L.load()
        self.ctx = L._vp()
        rc = self.lib.pfem_create(C.byref(self.ctx), int(device))
        if rc != 0:
            self.ctx = None
            L.check(None, rc)
        self.N = self.E = 0
        self.ncol = 0

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.pfem_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        return L.check(self.ctx, rc)

    # ---- problem description
    def set_layout(self, layout):
        """internal layout, before set_mesh: L.LAYOUT_ABI (default) or L.LAYOUT_VERTICAL_MINOR"""
        self._ck(self.lib.pfem_set_layout(self.ctx, int(layout)))

    def set_mesh(self, axes, strides):
        ax = [_f64(a) for a in axes]
        n = (L.c_sz * 3)(*[len(a) for a in ax])
        s = (L.c_sz * 3)(*[int(v) for v in strides])
        self._ck(self.lib.pfem_set_mesh(self.ctx, n, _dp(ax[0]), _dp(ax[1]), _dp(ax[2]), s))
        self.n = tuple(len(a) for a in ax)
        self.N = self.n[0] * self.n[1] * self.n[2]
        self.E = (self.n[0] - 1) * (self.n[1] - 1) * (self.n[2] - 1)

    def set_materials(self, elem_mat, T0, dT, c_lat, c_vert):
        m = np.ascontiguousarray(elem_mat, dtype=np.uint32)
        a, b = _f64(c_lat), _f64(c_vert)
        assert m.size == self.E and a.shape == b.shape and a.ndim == 2
        self._ck(self.lib.pfem_set_materials(self.ctx, m.ctypes.data_as(L._u32p), a.shape[0], float(T0), float(dT),
                                             a.shape[1], _dp(a), _dp(b)))

    def set_axis_weight(self, axis, w):
        """element weights along a physical axis (cylindrical 2-D embedding: midpoint radii); None removes them"""
        if w is None:
            self._ck(self.lib.pfem_set_axis_weight(self.ctx, int(axis), None))
            return
        a = _f64(w)
        assert a.ndim == 1 and a.size == self.n[axis] - 1
        self._ck(self.lib.pfem_set_axis_weight(self.ctx, int(axis), _dp(a)))

    def set_capacity(self, cp_dens):
        """cp(T)*dens(T) [J/(m^3 K)] per material id on the temperature grid of set_materials (Dynamic3D)."""
        a = _f64(cp_dens)
        assert a.ndim == 2
        self._ck(self.lib.pfem_set_capacity(self.ctx, a.shape[0], a.shape[1], _dp(a)))

    def set_dirichlet(self, nodes, values):
        n = np.ascontiguousarray(nodes, dtype=np.uintp)
        v = _f64(values)
        assert n.size == v.size
        self._ck(self.lib.pfem_set_dirichlet(self.ctx, n.size, n.ctypes.data_as(L._szp), _dp(v)))

    def set_source(self, heat):
        if heat is None:
            self._ck(self.lib.pfem_set_source(self.ctx, None))
        else:
            h = _f64(heat)
            assert h.size == self.E
            self._ck(self.lib.pfem_set_source(self.ctx, _dp(h)))

    def set_boundary(self, heatflux=(), convection=(), radiation=(), verbatim=True, mode2d=0):
        """heatflux_boundary / convection_boundary / radiation_boundary (therm3d.hpp:79-82) as lists of conditions in
        definition order: (nodes, q) / (nodes, coeff, ambient) / (nodes, emissivity, ambient).  Densified like
        BoundaryConditionsWithMesh::getValue: the FIRST condition naming a node wins.  mode2d = 1 / 2: the edge conditions of the 2-D
        solvers (therm2d.cpp, Cartesian / cylindrical) on the one-layer embedding; nodes are those of plane 0."""
        def dense(conds, nval):
            if not conds:
                return None, [None] * nval
            has = np.zeros(self.N, dtype=np.uint8)
            vals = [np.zeros(self.N) for _ in range(nval)]
            for cond in conds:
                nodes = np.asarray(cond[0], dtype=np.int64)
                new = nodes[has[nodes] == 0]
                for k in range(nval):
                    vals[k][new] = cond[1 + k]
                has[new] = 1
            return has, vals
        hf, (qf,) = dense(list(heatflux), 1)
        hc, (cc, ca) = dense(list(convection), 2)
        hr, (re, ra) = dense(list(radiation), 2)
        if hf is None and hc is None and hr is None:
            self._ck(self.lib.pfem_set_boundary(self.ctx, None))
            return
        b = L.Boundary()
        u8 = lambda a: a.ctypes.data_as(L._u8p) if a is not None else None
        dp = lambda a: _dp(a) if a is not None else None
        b.has_flux, b.flux = u8(hf), dp(qf)
        b.has_conv, b.conv_coeff, b.conv_ambient = u8(hc), dp(cc), dp(ca)
        b.has_rad, b.rad_emissivity, b.rad_ambient = u8(hr), dp(re), dp(ra)
        b.verbatim = int(bool(verbatim))
        b.mode2d = int(mode2d)
        self._ck(self.lib.pfem_set_boundary(self.ctx, C.byref(b)))

    def set_field(self, x):
        if np.isscalar(x):
            self._ck(self.lib.pfem_fill_field(self.ctx, float(x)))
        else:
            x = _f64(x)
            assert x.size == self.N
            self._ck(self.lib.pfem_set_field(self.ctx, _dp(x)))

    def set_elem_temperature(self, Te):
        if np.isscalar(Te):
            self._ck(self.lib.pfem_set_elem_temperature(self.ctx, None, float(Te)))
        else:
            t = _f64(Te)
            assert t.size == self.E
            self._ck(self.lib.pfem_set_elem_temperature(self.ctx, _dp(t), 0.))

    def set_junctions(self, junctions, elem_junc, elem_role, pcond, ncond, junc_cond, beta_col, js_col, stable=False):
        nj = len(junctions)
        arr = (L.Junction * max(nj, 1))()
        for k, j in enumerate(junctions):
            for name in ("bottom", "top", "left", "right", "back", "front", "ld", "offset", "height"):
                setattr(arr[k], name, getattr(j, name) if not isinstance(j, dict) else j[name])
        ej = np.ascontiguousarray(elem_junc, dtype=np.uint32)
        er = None if elem_role is None else np.ascontiguousarray(elem_role, dtype=np.uint8)
        jc, bc, jsn = _f64(junc_cond), _f64(beta_col), _f64(js_col)
        self.ncol = bc.size
        assert jc.size == 2 * self.ncol and jsn.size == self.ncol
        self._ck(self.lib.pfem_set_junctions(self.ctx, nj, arr, ej.ctypes.data_as(L._u32p),
                                             er.ctypes.data_as(L._u8p) if er is not None else None, float(pcond),
                                             float(ncond), self.ncol, _dp(jc), _dp(bc), _dp(jsn), int(bool(stable))))

    # ---- field exchange with another context on the same device (ThermoElectric meta loop)
    def set_noheat(self, noheat):
        nh = None if noheat is None else np.ascontiguousarray(noheat, dtype=np.uint8)
        self._ck(self.lib.pfem_set_noheat(self.ctx, nh.ctypes.data_as(L._u8p) if nh is not None else None))

    def take_temperature_from(self, thermal):
        """self (electrical): T_elem <- thermal's temperatures at my element midpoints"""
        self._ck(self.lib.pfem_transfer_temperature(self.ctx, thermal.ctx))

    def take_heat_from(self, electrical):
        """self (thermal): load vector <- electrical's Joule heat at my element midpoints"""
        rc = self.lib.pfem_transfer_heat(self.ctx, electrical.ctx)
        if rc < 0:   # the failing step recorded its message in the context it was working on
            detail = self.lib.pfem_last_error(electrical.ctx)
            L.check(electrical.ctx if detail else self.ctx, rc)
        return rc

    # ---- slab mode (one DeviceFem per GPU / process; see include/plaskfem_cuda.h)
    def slab_configure(self, rank, nranks, own_lo, own_hi):
        self._ck(self.lib.pfem_slab_configure(self.ctx, int(rank), int(nranks), int(own_lo), int(own_hi)))
        self.slab = (int(rank), int(nranks), int(own_lo), int(own_hi))

    def slab_export(self):
        buf = C.create_string_buffer(self.lib.pfem_slab_blob_size())
        self._ck(self.lib.pfem_slab_export(self.ctx, C.cast(buf, L._vp)))
        return bytes(buf.raw)

    def slab_connect(self, blobs):
        n = self.lib.pfem_slab_blob_size()
        assert all(len(b) == n for b in blobs)
        buf = C.create_string_buffer(b"".join(blobs), n * len(blobs))
        self._ck(self.lib.pfem_slab_connect(self.ctx, C.cast(buf, L._vp)))

    # ---- solve
    def opts(self, **kw):
        o = L.Opts()
        self.lib.pfem_default_opts(C.byref(o))
        for k, v in kw.items():
            if not hasattr(o, k):
                raise BadInput(f"unknown option {k}")
            setattr(o, k, v)
        return o

    def solve_thermal(self, **kw):
        o, st = self.opts(**kw), L.Stats()
        rc = self._ck(self.lib.pfem_solve_thermal(self.ctx, C.byref(o), C.byref(st)))
        return rc, st.as_dict()

    def solve_shockley(self, **kw):
        o, st = self.opts(**kw), L.Stats()
        rc = self._ck(self.lib.pfem_solve_shockley(self.ctx, C.byref(o), C.byref(st)))
        return rc, st.as_dict()

    def solve_dynamic(self, time, timestep, methodparam=0.5, lumping=True, rebuildfreq=0, log=False, **kw):
        """DynamicThermalFem3DSolver::compute(time) (femT3d.cpp:258-305, corrected update); returns (rc, stats[, maxT per step])."""
        o, st, d = self.opts(**kw), L.Stats(), L.Dynamic()
        d.time, d.timestep, d.methodparam, d.lumping, d.rebuildfreq = float(time), float(timestep), float(methodparam), int(bool(lumping)), int(rebuildfreq)
        buf = None
        if log:
            buf = np.zeros(int((time + timestep / 2.) / timestep) + 2)
            d.maxT_log, d.maxT_log_len = _dp(buf), buf.size
        rc = self._ck(self.lib.pfem_solve_dynamic(self.ctx, C.byref(o), C.byref(d), C.byref(st)))
        out = st.as_dict()
        if log:
            out["maxT_log"] = buf[:out["outer_loops"]].copy()
        return rc, out

    def solve_linear(self, **kw):
        o, st = self.opts(**kw), L.Stats()
        rc = self._ck(self.lib.pfem_solve_linear(self.ctx, C.byref(o), C.byref(st)))
        return rc, st.as_dict()

    def bench_pcg(self, iters, split_timing=False, **kw):
        o = self.opts(**kw)
        ms, a, u, n = C.c_double(0), C.c_double(0), C.c_double(0), C.c_longlong(0)
        self._ck(self.lib.pfem_bench_pcg(self.ctx, C.byref(o), int(iters), int(split_timing), C.byref(ms), C.byref(a),
                                         C.byref(u), C.byref(n)))
        return dict(ms=ms.value, apply_ms=a.value, update_ms=u.value, launches=n.value)

    # ---- results
    def get_field(self):
        x = np.empty(self.N)
        self._ck(self.lib.pfem_get_field(self.ctx, _dp(x)))
        return x

    def interpolate_field(self, axes, strides):
        """the field at the tensor-product points of a foreign rectilinear mesh (linear interpolation on the device)"""
        ax = [_f64(a) for a in axes]
        n = (L.c_sz * 3)(*[len(a) for a in ax])
        s = (L.c_sz * 3)(*[int(v) for v in strides])
        out = np.empty(len(ax[0]) * len(ax[1]) * len(ax[2]))
        self._ck(self.lib.pfem_interpolate_field(self.ctx, n, _dp(ax[0]), _dp(ax[1]), _dp(ax[2]), s, _dp(out)))
        return out

    def get_elem(self, what, noheat=None):
        nc = {L.ELEM_COND: 2, L.ELEM_CURRENT: 3, L.ELEM_HEAT: 1, L.ELEM_FLUX: 3}[what]
        out = np.empty((self.E, nc))
        nh = None if noheat is None else np.ascontiguousarray(noheat, dtype=np.uint8)
        self._ck(self.lib.pfem_get_elem(self.ctx, what, nh.ctypes.data_as(L._u8p) if nh is not None else None, _dp(out)))
        return out[:, 0] if nc == 1 else out

    def get_elem_temperature(self, elems):
        """T_elem at the given elements (element order of the mesh)"""
        e = np.ascontiguousarray(elems, dtype=np.uintp)
        out = np.empty(e.size)
        self._ck(self.lib.pfem_get_elem_temperature(self.ctx, e.size, e.ctypes.data_as(L._szp), _dp(out)))
        return out

    def get_junction_cond(self):
        out = np.empty((self.ncol, 2))
        self._ck(self.lib.pfem_get_junction_cond(self.ctx, _dp(out)))
        return out

    # ---- building blocks (tests / bench)
    def update_conductivity_thermal(self):
        self._ck(self.lib.pfem_update_conductivity_thermal(self.ctx))

    def update_conductivity_shockley(self):
        self._ck(self.lib.pfem_update_conductivity_shockley(self.ctx))

    def set_conductivity(self, cond):
        c = _f64(cond)
        assert c.size == 2 * self.E
        self._ck(self.lib.pfem_set_conductivity(self.ctx, _dp(c)))

    def apply(self, p, variant=3):
        p = _f64(p)
        q = np.empty(self.N)
        self._ck(self.lib.pfem_apply(self.ctx, _dp(p), _dp(q), int(variant)))
        return q

    def info(self, what):
        v = C.c_double(0)
        self._ck(self.lib.pfem_get_info(self.ctx, int(what), C.byref(v)))
        return v.value

    def apply_precond(self, r, **kw):
        """z = M^-1 r of the preconditioner `precond` built from the current conductivities (parity tests)"""
        r = _f64(r)
        z = np.empty(self.N)
        o = self.opts(**kw)
        self._ck(self.lib.pfem_apply_precond(self.ctx, C.byref(o), _dp(r), _dp(z)))
        return z

    def get_rhs(self):
        b = np.empty(self.N)
        self._ck(self.lib.pfem_get_rhs(self.ctx, _dp(b)))
        return b

    def get_diag(self):
        d = np.empty(self.N)
        self._ck(self.lib.pfem_get_diag(self.ctx, _dp(d)))
        return d
