"""ctypes binding of include/mgmc_b200.h (the same stub a reference maintainer would write, see
INTEGRATION.md).  Fails loudly when the CUDA library is missing -- there is no fallback path."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.environ.get("MGMC_LIB", os.path.join(_CSRC, "libmgmc_b200.so"))
_LIB = None

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]
SOURCES = ["mgmc_b200.cu"]
HEADERS = ["fused.cuh", "tail.cuh", "noise_ahead.cuh", "varcoef.cuh", "lattice3d.cuh", "rowfuse.cuh", "kernels.cuh", "philox.cuh", "normal_tables.inc", "setup.hh", "../../include/mgmc_b200.h"]


def build(force=False, verbose=False):
    """nvcc-compile the CUDA library in-tree (cross-compiles for sm_100a without a GPU)."""
    srcs = [os.path.join(_CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(_CSRC, h) for h in HEADERS]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return _SO
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", _SO] + srcs
    subprocess.check_call(cmd, cwd=_CSRC)
    return _SO


class MgmcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mgmc error {code}: {msg}")
        self.code = code


class Desc(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("pde_model", C.c_int), ("Lambda", C.c_double),
        ("m_lowrank", C.c_int), ("B_nnz", C.c_int64),
        ("B_rows", C.POINTER(C.c_int64)), ("B_cols", C.POINTER(C.c_int32)),
        ("B_vals", C.POINTER(C.c_double)), ("Sigma", C.POINTER(C.c_double)),
        ("nlevel", C.c_int), ("smoother", C.c_int), ("coarse_solver", C.c_int),
        ("npresmooth", C.c_int), ("npostsmooth", C.c_int), ("ncoarsesmooth", C.c_int),
        ("cycle", C.c_int), ("coarse_scaling", C.c_double), ("omega", C.c_double),
        ("seed", C.c_uint64), ("device", C.c_int), ("nchains", C.c_int), ("first_chain", C.c_int),
        ("strip_rank", C.c_int), ("strip_nranks", C.c_int),
        ("kappa_sq", C.POINTER(C.c_double)),
    ]


c_dp = C.POINTER(C.c_double)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_SO):
        raise ImportError(
            f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the MGMC path is CUDA-only, there is no CPU fallback)")
    L = C.CDLL(_SO)
    vp, i, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    ip = C.POINTER(C.c_int)
    sig = {
        "mgmc_last_error": (C.c_char_p, []),
        "mgmc_create": (i, [C.POINTER(Desc), C.POINTER(vp)]),
        "mgmc_destroy": (None, [vp]),
        "mgmc_level_info": (i, [vp, i, ip, ip, C.POINTER(i64), ip]),
        "mgmc_level_nz": (i, [vp, i, ip]),
        "mgmc_host_stencil3": (i, [C.POINTER(Desc), i, c_dp, ip]),
        "mgmc_get_stencil": (i, [vp, i, c_dp]),
        "mgmc_host_stencil": (i, [C.POINTER(Desc), i, c_dp, ip]),
        "mgmc_host_coefficients": (i, [C.POINTER(Desc), i, c_dp, ip]),
        "mgmc_op_apply": (i, [vp, i, c_dp, c_dp]),
        "mgmc_restrict": (i, [vp, i, c_dp, c_dp]),
        "mgmc_prolongate_add": (i, [vp, i, dbl, c_dp, c_dp]),
        "mgmc_residual_restrict": (i, [vp, i, c_dp, c_dp, c_dp]),
        "mgmc_smoother_apply": (i, [vp, i, i, i, dbl, i, c_dp, c_dp]),
        "mgmc_sampler_apply": (i, [vp, i, i, i, dbl, i, c_dp, c_dp]),
        "mgmc_coarse_solve": (i, [vp, c_dp, c_dp]),
        "mgmc_coarse_sample": (i, [vp, c_dp, c_dp]),
        "mgmc_sampler_mgmc_apply": (i, [vp, c_dp, c_dp]),
        "mgmc_mgprec_apply": (i, [vp, c_dp, c_dp]),
        "mgmc_loop_solve": (i, [vp, c_dp, c_dp, dbl, dbl, i, c_dp, ip, ip, ip]),
        "mgmc_set_philox_position": (i, [vp, C.c_uint32, C.c_uint32]),
        "mgmc_set_rhs": (i, [vp, c_dp]),
        "mgmc_set_state": (i, [vp, c_dp]),
        "mgmc_get_state": (i, [vp, c_dp]),
        "mgmc_set_qoi": (i, [vp, i64, C.POINTER(i64), c_dp]),
        "mgmc_sample": (i, [vp, i64, c_dp]),
        "mgmc_sample_moments": (i, [vp, i64, c_dp, c_dp]),
        "mgmc_sample_timed": (i, [vp, i64, c_dp, c_dp]),
        "mgmc_strip_partition": (i, [C.POINTER(Desc), i, i, ip, ip, ip]),
        "mgmc_plan_passes": (i, [i, i, ip, i, ip, i, i, i, i, ip, ip, ip]),
        "mgmc_strip_handle_bytes": (i, []),
        "mgmc_strip_export": (i, [vp, vp]),
        "mgmc_strip_connect": (i, [vp, vp]),
        "mgmc_strip_error": (i, [vp]),
        "mgmc_launch_count": (i64, [vp]),
        "mgmc_profile_cycle": (i, [vp, i, i, C.c_char_p, c_dp, C.POINTER(i64), c_dp, ip]),
        "mgmc_cycle_model": (i, [vp, c_dp, c_dp]),
        "mgmc_tail_stamps": (i, [vp, i, ip, c_dp, ip]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError = header / library mismatch
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


EXPORTS = [
    "mgmc_last_error", "mgmc_create", "mgmc_destroy", "mgmc_level_info", "mgmc_get_stencil", "mgmc_host_stencil",
    "mgmc_op_apply", "mgmc_restrict", "mgmc_prolongate_add", "mgmc_residual_restrict", "mgmc_smoother_apply",
    "mgmc_sampler_apply", "mgmc_coarse_solve", "mgmc_coarse_sample", "mgmc_sampler_mgmc_apply", "mgmc_mgprec_apply",
    "mgmc_loop_solve", "mgmc_set_philox_position", "mgmc_set_rhs", "mgmc_set_state", "mgmc_get_state", "mgmc_set_qoi",
    "mgmc_sample", "mgmc_sample_moments", "mgmc_sample_timed", "mgmc_launch_count", "mgmc_profile_cycle",
    "mgmc_cycle_model", "mgmc_strip_partition", "mgmc_strip_handle_bytes", "mgmc_strip_export", "mgmc_strip_connect",
    "mgmc_strip_error", "mgmc_plan_passes", "mgmc_tail_stamps", "mgmc_host_coefficients", "mgmc_level_nz", "mgmc_host_stencil3",
]


def _chk(status):
    if status != 0:
        raise MgmcError(status, lib().mgmc_last_error().decode())


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


PDE = {"shiftedlaplace_fd": 0, "squared_shiftedlaplace_fd": 1, "shiftedlaplace_fem": 2}
SMOOTHER = {"SOR": 0, "SSOR": 1}
COARSE = {"SSOR": 0, "Cholesky": 1}
FORWARD, BACKWARD = 1, 2


def make_desc(nx, ny, nlevel, pde="shiftedlaplace_fd", Lambda=0.2, B=None, smoother="SSOR", coarse_solver="Cholesky",
              npresmooth=1, npostsmooth=1, ncoarsesmooth=1, cycle=1, coarse_scaling=1.0, omega=1.0, seed=5418513,
              device=0, nchains=1, first_chain=0, strip_rank=0, strip_nranks=0, kappa_sq=None, nz=None):
    """B = (rows, cols, vals, sigma) COO triplets of the measurement matrix (lexicographic rows).
    kappa_sq = kappa^2 at every interior vertex (lexicographic, (nx-1)*(ny-1) values) for a correlation length
    that varies in space (`periodic_kappa_sq`); None: constant, 1 / Lambda^2.
    nz = cells in z of a 3d lattice (Lattice3d; host vectors lexicographic with x fastest, z slowest); None: 2d."""
    d = Desc()
    d.dim, d.nx, d.ny, d.nz = (2, nx, ny, 1) if nz is None else (3, nx, ny, nz)
    d.pde_model, d.Lambda = PDE[pde], Lambda
    keep = []
    if kappa_sq is not None:
        ks = np.ascontiguousarray(kappa_sq, dtype=np.float64).ravel()
        if ks.size != (nx - 1) * (ny - 1):
            raise ValueError("kappa_sq needs (nx-1)*(ny-1) entries")
        keep.append(ks)
        d.kappa_sq = ks.ctypes.data_as(c_dp)
    if B is not None:
        rows = np.ascontiguousarray(B[0], dtype=np.int64)
        cols = np.ascontiguousarray(B[1], dtype=np.int32)
        vals = np.ascontiguousarray(B[2], dtype=np.float64)
        sigma = np.ascontiguousarray(B[3], dtype=np.float64)
        d.m_lowrank, d.B_nnz = len(sigma), len(vals)
        d.B_rows = rows.ctypes.data_as(C.POINTER(C.c_int64))
        d.B_cols = cols.ctypes.data_as(C.POINTER(C.c_int32))
        d.B_vals = vals.ctypes.data_as(c_dp)
        d.Sigma = sigma.ctypes.data_as(c_dp)
        keep += [rows, cols, vals, sigma]
    d.nlevel, d.smoother, d.coarse_solver = nlevel, SMOOTHER[smoother], COARSE[coarse_solver]
    d.npresmooth, d.npostsmooth, d.ncoarsesmooth = npresmooth, npostsmooth, ncoarsesmooth
    d.cycle, d.coarse_scaling, d.omega = cycle, coarse_scaling, omega
    d.seed, d.device, d.nchains, d.first_chain = seed, device, nchains, first_chain
    d.strip_rank, d.strip_nranks = strip_rank, strip_nranks
    d._keep = keep
    return d


def plan_passes(ncolours, colours, fix_after=(), omega_is_one=True, restrict_behind=False, lr_mx=0, lr_my=0):
    """Pass plan of one fused launch (host-only): (modes, margins[npass][4], input halo[4]) -- include/mgmc_b200.h."""
    n = len(colours)
    col = (C.c_int * max(n, 1))(*colours)
    fx = (C.c_int * max(len(fix_after), 1))(*fix_after)
    mode = (C.c_int * max(n, 1))()
    marg = (C.c_int * max(4 * n, 1))()
    halo = (C.c_int * 4)()
    _chk(lib().mgmc_plan_passes(ncolours, n, col, len(fix_after), fx, int(omega_is_one), int(restrict_behind), lr_mx, lr_my, mode, marg, halo))
    return list(mode)[:n], [list(marg)[4 * s:4 * s + 4] for s in range(n)], list(halo)


def strip_partition(desc, level, rank):
    """(row_lo, row_hi, distributed) of `rank` on `level` (host-only)."""
    lo, hi, dist = C.c_int(), C.c_int(), C.c_int()
    _chk(lib().mgmc_strip_partition(C.byref(desc), level, rank, C.byref(lo), C.byref(hi), C.byref(dist)))
    return lo.value, hi.value, bool(dist.value)


def periodic_kappa_sq(nx, ny, Lambda_min, Lambda_max):
    """kappa^2 of PeriodicCorrelationLengthModel (correlationlength_model.hh:83-113) at the interior vertices of an
    nx x ny lattice, lexicographic: Lambda(x) = Lambda_1 + Lambda_2 cos(pi x_1) cos(pi x_2)."""
    L1, L2 = 0.5 * (Lambda_max + Lambda_min), 0.5 * (Lambda_max - Lambda_min)
    cx = np.cos(np.pi * (np.arange(1, nx) / nx))
    cy = np.cos(np.pi * (np.arange(1, ny) / ny))
    lam = (L2 * cx)[None, :] * cy[:, None] + L1
    out = 1.0 / (lam * lam)
    return out.ravel()


def host_coefficients(desc, level):
    """Per-vertex operator of `level` (desc.kappa_sq given) as (9, ny_l + 1, nx_l + 1) planes, host-only."""
    nx, ny = desc.nx >> level, desc.ny >> level
    out = np.zeros((9, ny + 1, nx + 1))
    nc = C.c_int()
    _chk(lib().mgmc_host_coefficients(C.byref(desc), level, out.ctypes.data_as(c_dp), C.byref(nc)))
    return out, nc.value


def host_stencil3(desc, level):
    """Uniform radius-1 stencil of `level` of a 3d hierarchy as a (3, 3, 3) array [dk + 1, dj + 1, di + 1], host-only."""
    out = np.zeros((3, 3, 3))
    nc = C.c_int()
    _chk(lib().mgmc_host_stencil3(C.byref(desc), level, out.ctypes.data_as(c_dp), C.byref(nc)))
    return out, nc.value


def host_stencil(desc, level):
    out = np.zeros((9, 5, 5))
    nc = C.c_int()
    _chk(lib().mgmc_host_stencil(C.byref(desc), level, out.ctypes.data_as(c_dp), C.byref(nc)))
    return out, nc.value


class Context:
    """Owns an ``mgmc_ctx`` (operator hierarchy resident on one B200)."""

    def __init__(self, nx, ny, nlevel, **kw):
        self.desc = make_desc(nx, ny, nlevel, **kw)
        self.nchains = self.desc.nchains
        h = C.c_void_p()
        _chk(lib().mgmc_create(C.byref(self.desc), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().mgmc_destroy(self.h)
            self.h = None

    __del__ = close

    def level_info(self, level):
        nx, ny, nc, nd = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        _chk(lib().mgmc_level_info(self.h, level, C.byref(nx), C.byref(ny), C.byref(nd), C.byref(nc)))
        return nx.value, ny.value, nd.value, nc.value

    def ndof(self, level=0):
        return self.level_info(level)[2]

    def stencil(self, level):
        out = np.zeros((9, 5, 5))
        _chk(lib().mgmc_get_stencil(self.h, level, out.ctypes.data_as(c_dp)))
        return out

    def _out(self, level):
        return np.empty(self.ndof(level) * self.nchains)

    def op_apply(self, level, x):
        x, xp = _d(x)
        y = self._out(level)
        _chk(lib().mgmc_op_apply(self.h, level, xp, y.ctypes.data_as(c_dp)))
        return y

    def restrict(self, level, x):
        x, xp = _d(x)
        y = self._out(level + 1)
        _chk(lib().mgmc_restrict(self.h, level, xp, y.ctypes.data_as(c_dp)))
        return y

    def prolongate_add(self, level, alpha, xc, x):
        xc, cp = _d(xc)
        x = np.array(x, dtype=np.float64, copy=True)
        _chk(lib().mgmc_prolongate_add(self.h, level, alpha, cp, x.ctypes.data_as(c_dp)))
        return x

    def residual_restrict(self, level, f, x):
        f, fp = _d(f)
        x, xp = _d(x)
        y = self._out(level + 1)
        _chk(lib().mgmc_residual_restrict(self.h, level, fp, xp, y.ctypes.data_as(c_dp)))
        return y

    def smoother_apply(self, level, kind, b, x, omega=1.0, nsmooth=1, direction=FORWARD):
        b, bp = _d(b)
        x = np.array(x, dtype=np.float64, copy=True)
        _chk(lib().mgmc_smoother_apply(self.h, level, SMOOTHER[kind], direction, omega, nsmooth, bp, x.ctypes.data_as(c_dp)))
        return x

    def sampler_apply(self, level, kind, f, x, omega=1.0, nsmooth=1, direction=FORWARD):
        f, fp = _d(f)
        x = np.array(x, dtype=np.float64, copy=True)
        _chk(lib().mgmc_sampler_apply(self.h, level, SMOOTHER[kind], direction, omega, nsmooth, fp, x.ctypes.data_as(c_dp)))
        return x

    def coarse_solve(self, b):
        b, bp = _d(b)
        x = np.empty_like(b)
        _chk(lib().mgmc_coarse_solve(self.h, bp, x.ctypes.data_as(c_dp)))
        return x

    def coarse_sample(self, f):
        f, fp = _d(f)
        x = np.empty_like(f)
        _chk(lib().mgmc_coarse_sample(self.h, fp, x.ctypes.data_as(c_dp)))
        return x

    def mgmc_apply(self, f, x):
        f, fp = _d(f)
        x = np.array(x, dtype=np.float64, copy=True)
        _chk(lib().mgmc_sampler_mgmc_apply(self.h, fp, x.ctypes.data_as(c_dp)))
        return x

    def mgprec_apply(self, b):
        b, bp = _d(b)
        x = np.empty_like(b)
        _chk(lib().mgmc_mgprec_apply(self.h, bp, x.ctypes.data_as(c_dp)))
        return x

    def loop_solve(self, b, rtol=1e-12, atol=1e-15, maxiter=100):
        b, bp = _d(b)
        x = np.empty_like(b)
        hist = np.zeros(maxiter)
        nh, ni, cv = C.c_int(), C.c_int(), C.c_int()
        _chk(lib().mgmc_loop_solve(self.h, bp, x.ctypes.data_as(c_dp), rtol, atol, maxiter, hist.ctypes.data_as(c_dp), C.byref(nh), C.byref(ni), C.byref(cv)))
        return x, hist[: nh.value].copy(), ni.value, bool(cv.value)

    def set_philox_position(self, sample, sweep_counter=0):
        _chk(lib().mgmc_set_philox_position(self.h, sample, sweep_counter))

    def set_rhs(self, f):
        f, fp = _d(f)
        _chk(lib().mgmc_set_rhs(self.h, fp))

    def set_state(self, x):
        x, xp = _d(x)
        _chk(lib().mgmc_set_state(self.h, xp))

    def get_state(self):
        x = self._out(0)
        _chk(lib().mgmc_get_state(self.h, x.ctypes.data_as(c_dp)))
        return x

    def set_qoi(self, idx, val):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        val, vp = _d(val)
        _chk(lib().mgmc_set_qoi(self.h, len(idx), idx.ctypes.data_as(C.POINTER(C.c_int64)), vp))

    def sample(self, nsamples, series=True):
        out = np.empty(nsamples * self.nchains) if series else None
        _chk(lib().mgmc_sample(self.h, nsamples, out.ctypes.data_as(c_dp) if series else None))
        return out.reshape(nsamples, self.nchains) if series else None

    def sample_timed(self, nsamples, series=False):
        out = np.empty(nsamples * self.nchains) if series else None
        ms = C.c_double()
        _chk(lib().mgmc_sample_timed(self.h, nsamples, out.ctypes.data_as(c_dp) if series else None, C.byref(ms)))
        return ms.value, (out.reshape(nsamples, self.nchains) if series else None)

    def sample_moments(self, nsamples):
        n = self.ndof(0)
        mean, second = np.empty(n), np.empty(n)
        _chk(lib().mgmc_sample_moments(self.h, nsamples, mean.ctypes.data_as(c_dp), second.ctypes.data_as(c_dp)))
        return mean, second

    # ---- row-strip decomposition (one process per GPU) ----
    def strip_export(self):
        buf = C.create_string_buffer(lib().mgmc_strip_handle_bytes())
        _chk(lib().mgmc_strip_export(self.h, buf))
        return buf.raw

    def strip_connect(self, blobs):
        """blobs: list of the handle blobs of all ranks, ordered by rank."""
        data = b"".join(blobs)
        _chk(lib().mgmc_strip_connect(self.h, C.c_char_p(data)))

    def strip_error(self):
        return lib().mgmc_strip_error(self.h)

    def launch_count(self):
        return lib().mgmc_launch_count(self.h)

    def profile_cycle(self, nsamples=3, nslots_max=256):
        names = C.create_string_buffer(nslots_max * 64)
        ms = np.zeros(nslots_max)
        launches = np.zeros(nslots_max, dtype=np.int64)
        byts = np.zeros(nslots_max)
        n = C.c_int()
        _chk(lib().mgmc_profile_cycle(self.h, nsamples, nslots_max, names, ms.ctypes.data_as(c_dp), launches.ctypes.data_as(C.POINTER(C.c_int64)),
                                      byts.ctypes.data_as(c_dp), C.byref(n)))
        out = []
        for k in range(n.value):
            out.append((names.raw[k * 64:(k + 1) * 64].split(b"\0")[0].decode(), float(ms[k]), int(launches[k]), float(byts[k])))
        return out

    def tail_stamps(self, nmax=256):
        """[(kind, us)] per phase of the last persistent tail launch (needs MGMC_TAIL_STAMPS=1)"""
        kinds = (C.c_int * nmax)()
        us = np.zeros(nmax)
        n = C.c_int()
        _chk(lib().mgmc_tail_stamps(self.h, nmax, kinds, us.ctypes.data_as(c_dp), C.byref(n)))
        return [(int(kinds[k]), float(us[k])) for k in range(n.value)]

    def cycle_model(self):
        b, u = C.c_double(), C.c_double()
        _chk(lib().mgmc_cycle_model(self.h, C.byref(b), C.byref(u)))
        return b.value, u.value
