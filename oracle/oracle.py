"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/liboracle.so`` (the CPU restatement of the reference hot path, see
``oracle_core.hh``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product package
``multigridmc_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_long)


def build(force=False):
    """Compile liboracle.so with the Makefile next to this file (g++ only, no dependencies)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force and os.path.exists(so):
        os.remove(so)
    # (make rebuilds only when a source is newer than the library; without the sources -- never the case in this
    #  repository -- the prebuilt library is used as it is)
    if os.path.exists(os.path.join(_HERE, "oracle_core.cc")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


class MGParams(C.Structure):
    _fields_ = [
        ("nlevel", C.c_int),
        ("smoother", C.c_int),
        ("coarse_solver", C.c_int),
        ("npresmooth", C.c_int),
        ("npostsmooth", C.c_int),
        ("ncoarsesmooth", C.c_int),
        ("cycle", C.c_int),
        ("coarse_scaling", C.c_double),
        ("omega", C.c_double),
    ]


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp = C.c_void_p
    sig = {
        "orc_last_error": (C.c_char_p, []),
        "orc_rng_create": (vp, [C.c_int, C.c_uint64]),
        "orc_rng_destroy": (None, [vp]),
        "orc_rng_normal": (C.c_int, [vp, C.c_long, c_dp]),
        "orc_rng_uniform": (C.c_int, [vp, C.c_long, c_dp]),
        "orc_lattice_nvertex": (C.c_long, [C.c_int, c_ip]),
        "orc_lattice_ncell": (C.c_long, [C.c_int, c_ip]),
        "orc_lattice_vertex_e2l": (C.c_long, [C.c_int, c_ip, c_ip]),
        "orc_lattice_vertex_l2e": (C.c_int, [C.c_int, c_ip, C.c_long, c_ip]),
        "orc_lattice_cell_e2l": (C.c_long, [C.c_int, c_ip, c_ip]),
        "orc_lattice_cell_l2e": (C.c_int, [C.c_int, c_ip, C.c_long, c_ip]),
        "orc_lattice_shift_vertexidx": (C.c_long, [C.c_int, c_ip, C.c_long, c_ip]),
        "orc_lattice_shift_cellidx": (C.c_long, [C.c_int, c_ip, C.c_long, c_ip]),
        "orc_lattice_corner_vertex": (C.c_long, [C.c_int, c_ip, C.c_long, c_ip]),
        "orc_lattice_fine_vertex_idx": (C.c_long, [C.c_int, c_ip, C.c_long]),
        "orc_lattice_vertex_coordinates": (C.c_int, [C.c_int, c_ip, C.c_long, c_dp]),
        "orc_lattice_coarsen": (C.c_int, [C.c_int, c_ip, c_ip]),
        "orc_op_create_prior": (vp, [C.c_int, c_ip, C.c_int, C.c_int, C.c_double, C.c_double]),
        "orc_op_create_measured": (vp, [vp, C.c_int, c_dp, c_dp, C.c_double, C.c_double, C.c_int, C.c_double]),
        "orc_op_create_test1d": (vp, [C.c_int]),
        "orc_op_destroy": (None, [vp]),
        "orc_op_ndof": (C.c_long, [vp]),
        "orc_op_m_lowrank": (C.c_int, [vp]),
        "orc_op_lattice": (C.c_int, [vp, c_ip, c_ip]),
        "orc_op_apply": (C.c_int, [vp, c_dp, c_dp]),
        "orc_op_nnz": (C.c_long, [vp]),
        "orc_op_get_csr": (C.c_int, [vp, c_lp, c_ip, c_dp]),
        "orc_op_B_nnz": (C.c_long, [vp]),
        "orc_op_get_B": (C.c_int, [vp, c_lp, c_ip, c_dp, c_dp]),
        "orc_op_precision": (C.c_int, [vp, c_dp]),
        "orc_op_covariance": (C.c_int, [vp, c_dp]),
        "orc_op_mean": (C.c_int, [vp, c_dp, c_dp, c_dp]),
        "orc_op_observed_mean_and_variance": (C.c_int, [vp, c_dp, c_dp, c_dp, c_dp, c_dp]),
        "orc_measurement_vector": (C.c_int, [vp, c_dp, C.c_double, c_dp]),
        "orc_hier_create": (vp, [vp, C.c_int, C.c_int]),
        "orc_hier_destroy": (None, [vp]),
        "orc_hier_op": (vp, [vp, C.c_int]),
        "orc_hier_ncolours": (C.c_int, [vp, C.c_int]),
        "orc_hier_order": (C.c_int, [vp, C.c_int, c_lp]),
        "orc_hier_restrict": (C.c_int, [vp, C.c_int, c_dp, c_dp]),
        "orc_hier_prolongate_add": (C.c_int, [vp, C.c_int, C.c_double, c_dp, c_dp]),
        "orc_smoother_create": (vp, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]),
        "orc_sampler_create": (vp, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, vp, C.c_uint64]),
        "orc_mgmc_create": (vp, [vp, C.POINTER(MGParams), vp, C.c_uint64]),
        "orc_mgprec_create": (vp, [vp, C.POINTER(MGParams)]),
        "orc_cholesky_solver_create": (vp, [vp, C.c_int]),
        "orc_obj_destroy": (None, [vp]),
        "orc_obj_apply": (C.c_int, [vp, c_dp, c_dp]),
        "orc_obj_set_philox_position": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32]),
        "orc_sampler_run": (C.c_int, [vp, c_dp, c_dp, c_dp, C.c_long, c_dp]),
        "orc_sampler_moments": (C.c_int, [vp, c_dp, c_dp, C.c_long, C.c_long, c_dp, c_dp]),
        "orc_loop_solve": (C.c_int, [vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_ip, c_ip, c_ip]),
        "orc_tau_int": (C.c_double, [c_dp, C.c_long, C.c_int]),
        "orc_philox_normal_pair": (None, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, c_dp, c_dp]),
        "orc_philox_raw": (None, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


class OracleError(RuntimeError):
    pass


def _chk(status):
    if status != 0:
        raise OracleError(lib().orc_last_error().decode())


def _ptr(h):
    if not h:
        raise OracleError(lib().orc_last_error().decode())
    return h


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


PDE = {"shiftedlaplace_fd": 0, "squared_shiftedlaplace_fd": 1, "shiftedlaplace_fem": 2}


class StdRng:
    """std::mt19937_64 (bits=64) or std::mt19937 (bits=32) with persistent distributions."""

    def __init__(self, seed, bits=64):
        self.h = _ptr(lib().orc_rng_create(bits, seed))

    def normal(self, n):
        out = np.empty(n)
        lib().orc_rng_normal(self.h, n, out.ctypes.data_as(c_dp))
        return out

    def uniform(self, n):
        out = np.empty(n)
        lib().orc_rng_uniform(self.h, n, out.ctypes.data_as(c_dp))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_rng_destroy(self.h)


class Lattice:
    def __init__(self, *n):
        self.dim = len(n)
        self.n, self._n = _i(list(n))

    @property
    def Nvertex(self):
        return lib().orc_lattice_nvertex(self.dim, self._n)

    @property
    def Ncell(self):
        return lib().orc_lattice_ncell(self.dim, self._n)

    def vertexidx_euclidean2linear(self, idx):
        a, p = _i(idx)
        return lib().orc_lattice_vertex_e2l(self.dim, self._n, p)

    def vertexidx_linear2euclidean(self, ell):
        out = np.zeros(3, dtype=np.int32)
        lib().orc_lattice_vertex_l2e(self.dim, self._n, ell, out.ctypes.data_as(c_ip))
        return out[: self.dim].copy()

    def cellidx_euclidean2linear(self, idx):
        a, p = _i(idx)
        return lib().orc_lattice_cell_e2l(self.dim, self._n, p)

    def cellidx_linear2euclidean(self, ell):
        out = np.zeros(3, dtype=np.int32)
        lib().orc_lattice_cell_l2e(self.dim, self._n, ell, out.ctypes.data_as(c_ip))
        return out[: self.dim].copy()

    def shift_vertexidx(self, ell, shift):
        a, p = _i(shift)
        return lib().orc_lattice_shift_vertexidx(self.dim, self._n, ell, p)

    def shift_cellidx(self, ell, shift):
        a, p = _i(shift)
        return lib().orc_lattice_shift_cellidx(self.dim, self._n, ell, p)

    def corner_is_internal_vertex(self, cell, corner):
        a, p = _i(corner)
        return lib().orc_lattice_corner_vertex(self.dim, self._n, cell, p)

    def fine_vertex_idx(self, ell):
        return lib().orc_lattice_fine_vertex_idx(self.dim, self._n, ell)

    def vertex_coordinates(self, ell):
        out = np.zeros(3)
        lib().orc_lattice_vertex_coordinates(self.dim, self._n, ell, out.ctypes.data_as(c_dp))
        return out[: self.dim].copy()

    def get_coarse_lattice(self):
        out = np.zeros(3, dtype=np.int32)
        _chk(lib().orc_lattice_coarsen(self.dim, self._n, out.ctypes.data_as(c_ip)))
        return Lattice(*[int(v) for v in out[: self.dim]])


class Operator:
    """LinearOperator A = A0 + B Sigma^-1 B^T (owning or borrowed handle)."""

    def __init__(self, h, owner=True, keep=None):
        self.h = _ptr(h)
        self._owner = owner
        self._keep = keep

    @classmethod
    def prior(cls, n, pde="shiftedlaplace_fd", Lambda=None, Lambda_min=None, Lambda_max=None):
        n_, p = _i(list(n))
        if Lambda is not None:
            return cls(lib().orc_op_create_prior(len(n), p, PDE[pde], 0, float(Lambda), 0.0))
        return cls(lib().orc_op_create_prior(len(n), p, PDE[pde], 1, float(Lambda_min), float(Lambda_max)))

    @classmethod
    def test1d(cls, lowrank):
        return cls(lib().orc_op_create_test1d(int(lowrank)))

    def measured(self, locations, variance, variance_scaling=1.0, radius=0.0, measure_global=False, variance_global=0.0):
        loc, lp = _d(np.asarray(locations, dtype=np.float64).ravel())
        var, vp = _d(variance)
        return Operator(
            lib().orc_op_create_measured(self.h, len(var), lp, vp, variance_scaling, radius, int(measure_global), variance_global)
        )

    def __del__(self):
        if getattr(self, "_owner", False) and getattr(self, "h", None):
            lib().orc_op_destroy(self.h)

    @property
    def ndof(self):
        return lib().orc_op_ndof(self.h)

    @property
    def m_lowrank(self):
        return lib().orc_op_m_lowrank(self.h)

    @property
    def shape(self):
        dim = C.c_int()
        n = np.zeros(3, dtype=np.int32)
        lib().orc_op_lattice(self.h, C.byref(dim), n.ctypes.data_as(c_ip))
        return tuple(int(v) for v in n[: dim.value])

    def apply(self, x):
        x, xp = _d(x)
        y = np.empty(self.ndof)
        _chk(lib().orc_op_apply(self.h, xp, y.ctypes.data_as(c_dp)))
        return y

    def csr(self):
        import scipy.sparse as sp

        n, nnz = self.ndof, lib().orc_op_nnz(self.h)
        rowptr = np.empty(n + 1, dtype=np.int64)
        col = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz)
        lib().orc_op_get_csr(self.h, rowptr.ctypes.data_as(c_lp), col.ctypes.data_as(c_ip), val.ctypes.data_as(c_dp))
        return sp.csr_matrix((val, col, rowptr), shape=(n, n))

    def B(self):
        """(rows, cols, vals, sigma) COO triplets of B and the diagonal of Sigma."""
        nnz, m = lib().orc_op_B_nnz(self.h), self.m_lowrank
        rows = np.empty(nnz, dtype=np.int64)
        cols = np.empty(nnz, dtype=np.int32)
        vals = np.empty(nnz)
        sigma = np.empty(m)
        lib().orc_op_get_B(self.h, rows.ctypes.data_as(c_lp), cols.ctypes.data_as(c_ip), vals.ctypes.data_as(c_dp), sigma.ctypes.data_as(c_dp))
        return rows, cols, vals, sigma

    def precision(self):
        n = self.ndof
        out = np.empty((n, n))
        _chk(lib().orc_op_precision(self.h, out.ctypes.data_as(c_dp)))
        return out

    def covariance(self):
        n = self.ndof
        out = np.empty((n, n))
        _chk(lib().orc_op_covariance(self.h, out.ctypes.data_as(c_dp)))
        return out

    def mean(self, xbar, y):
        xbar, xp = _d(xbar)
        y, yp = _d(y)
        out = np.empty(self.ndof)
        _chk(lib().orc_op_mean(self.h, xp, yp, out.ctypes.data_as(c_dp)))
        return out

    def observed_mean_and_variance(self, xbar, y, b_obs):
        xbar, xp = _d(xbar)
        y, yp = _d(y)
        b, bp = _d(b_obs)
        m, v = C.c_double(), C.c_double()
        _chk(lib().orc_op_observed_mean_and_variance(self.h, xp, yp, bp, C.byref(m), C.byref(v)))
        return m.value, v.value

    def measurement_vector(self, x0, radius):
        x0, p = _d(x0)
        out = np.empty(self.ndof)
        _chk(lib().orc_measurement_vector(self.h, p, radius, out.ctypes.data_as(c_dp)))
        return out


class _Obj:
    def __init__(self, h, keep):
        self.h = _ptr(h)
        self._keep = keep

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_obj_destroy(self.h)

    def apply(self, b, x):
        """apply(b, x) -> new x (x is copied; in/out semantics for smoothers and samplers)."""
        b, bp = _d(b)
        x = np.array(x, dtype=np.float64, copy=True)
        _chk(lib().orc_obj_apply(self.h, bp, x.ctypes.data_as(c_dp)))
        return x

    def set_philox_position(self, sample, chain=0, sweep_counter=0):
        _chk(lib().orc_obj_set_philox_position(self.h, sample, chain, sweep_counter))

    def run(self, f, x, b_obs, nsamples):
        f, fp = _d(f)
        b, bp = _d(b_obs)
        x = np.array(x, dtype=np.float64, copy=True)
        series = np.empty(nsamples)
        _chk(lib().orc_sampler_run(self.h, fp, x.ctypes.data_as(c_dp), bp, nsamples, series.ctypes.data_as(c_dp)))
        return x, series

    def moments(self, f, x, nwarmup, nsamples):
        f, fp = _d(f)
        x = np.array(x, dtype=np.float64, copy=True)
        n = x.size
        Ex, Exx = np.empty(n), np.empty((n, n))
        _chk(lib().orc_sampler_moments(self.h, fp, x.ctypes.data_as(c_dp), nwarmup, nsamples, Ex.ctypes.data_as(c_dp), Exx.ctypes.data_as(c_dp)))
        return Ex, Exx


LEX, COLOUR = 0, 1
FORWARD, BACKWARD = 1, 2


class Hierarchy:
    def __init__(self, op, nlevel, ordering=LEX):
        self.op = op
        self.nlevel = nlevel
        self.h = _ptr(lib().orc_hier_create(op.h, nlevel, ordering))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_hier_destroy(self.h)

    def level_op(self, level):
        return Operator(lib().orc_hier_op(self.h, level), owner=False, keep=self)

    def ncolours(self, level):
        return lib().orc_hier_ncolours(self.h, level)

    def order(self, level):
        out = np.empty(self.level_op(level).ndof, dtype=np.int64)
        lib().orc_hier_order(self.h, level, out.ctypes.data_as(c_lp))
        return out

    def restrict(self, level, x):
        x, xp = _d(x)
        out = np.empty(self.level_op(level + 1).ndof)
        lib().orc_hier_restrict(self.h, level, xp, out.ctypes.data_as(c_dp))
        return out

    def prolongate_add(self, level, alpha, xc, x):
        xc, cp = _d(xc)
        x = np.array(x, dtype=np.float64, copy=True)
        lib().orc_hier_prolongate_add(self.h, level, alpha, cp, x.ctypes.data_as(c_dp))
        return x

    def smoother(self, level, kind, omega, nsmooth=1, direction=FORWARD):
        k = {"SOR": 0, "SSOR": 1}[kind]
        return _Obj(lib().orc_smoother_create(self.h, level, k, omega, nsmooth, direction), self)

    def sampler(self, level, kind, omega=1.0, nsmooth=1, direction=FORWARD, rng=None, philox_seed=0):
        k = {"SOR": 0, "SSOR": 1, "Cholesky": 2}[kind]
        return _Obj(lib().orc_sampler_create(self.h, level, k, omega, nsmooth, direction, rng.h if rng else None, philox_seed), (self, rng))

    def cholesky_solver(self, level):
        return _Obj(lib().orc_cholesky_solver_create(self.h, level), self)

    def _params(self, smoother="SSOR", coarse_solver="Cholesky", npresmooth=1, npostsmooth=1, ncoarsesmooth=1, cycle=1, coarse_scaling=1.0, omega=1.0):
        return MGParams(self.nlevel, {"SOR": 0, "SSOR": 1}[smoother], {"SSOR": 0, "Cholesky": 1}[coarse_solver], npresmooth, npostsmooth, ncoarsesmooth, cycle, coarse_scaling, omega)

    def mgmc(self, rng=None, philox_seed=0, **kw):
        p = self._params(**kw)
        return _Obj(lib().orc_mgmc_create(self.h, C.byref(p), rng.h if rng else None, philox_seed), (self, rng))

    def preconditioner(self, **kw):
        p = self._params(**kw)
        return _Obj(lib().orc_mgprec_create(self.h, C.byref(p)), self)


def loop_solve(op, prec, b, rtol=1e-12, atol=1e-15, maxiter=100, verbose=0):
    """LoopSolver::apply (loop_solver.cc:9-53) -> (x, history, niter, converged)."""
    b, bp = _d(b)
    x = np.zeros(op.ndof)
    hist = np.zeros(maxiter)
    nh, ni, cv = C.c_int(), C.c_int(), C.c_int()
    _chk(lib().orc_loop_solve(op.h, prec.h, rtol, atol, maxiter, verbose, bp, x.ctypes.data_as(c_dp), hist.ctypes.data_as(c_dp), C.byref(nh), C.byref(ni), C.byref(cv)))
    return x, hist[: nh.value].copy(), ni.value, bool(cv.value)


def tau_int(series, window):
    s, p = _d(series)
    return lib().orc_tau_int(p, s.size, window)


def philox_normal_pair(seed, c0, c1, c2, c3):
    z0, z1 = C.c_double(), C.c_double()
    lib().orc_philox_normal_pair(seed, c0, c1, c2, c3, C.byref(z0), C.byref(z1))
    return z0.value, z1.value


def philox_raw(seed, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().orc_philox_raw(seed, c0, c1, c2, c3, out)
    return [int(v) for v in out]
