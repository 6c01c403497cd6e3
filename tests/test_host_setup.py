"""CPU tests of the product's host-side setup algebra (no GPU needed) against the oracle, and that
the C-ABI library loads and exports every symbol that include/mgmc_b200.h declares."""
import os
import re

import numpy as np
import pytest

import multigridmc_b200 as m
from multigridmc_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mgmc_b200.h")).read()
    declared = set(re.findall(r"\b(mgmc_[a-z0-9_]+)\s*\(", header))
    L = m.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/mgmc_b200.h but not exported"
    assert declared == set(capi.EXPORTS)


def test_no_cpu_fallback_without_gpu():
    """mgmc_create must fail loudly (MGMC_ERR_CUDA) when there is no device."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(m.MgmcError) as e:
        m.Context(16, 16, 2)
    assert e.value.code == -3


def _stencil_matrix(st, nx, ny):
    """Dense matrix of a 9-class stencil set on an (nx, ny)-cell lattice."""
    w, h = nx - 1, ny - 1
    A = np.zeros((w * h, w * h))
    cls = lambda i, n: 0 if i == 1 else (2 if i == n - 1 else 1)
    for j in range(1, ny):
        for i in range(1, nx):
            c = cls(i, nx) + 3 * cls(j, ny)
            for dj in range(-2, 3):
                for di in range(-2, 3):
                    ii, jj = i + di, j + dj
                    if 1 <= ii < nx and 1 <= jj < ny:
                        A[(j - 1) * w + i - 1, (jj - 1) * w + ii - 1] = st[c, dj + 2, di + 2]
    return A


@pytest.mark.parametrize("pde,n,nlevel", [
    ("shiftedlaplace_fd", (32, 32), 4),
    ("shiftedlaplace_fd", (32, 16), 3),
    ("squared_shiftedlaplace_fd", (32, 32), 3),
    ("squared_shiftedlaplace_fd", (64, 32), 4),
])
def test_galerkin_stencils_match_oracle_triple_product(oracle, pde, n, nlevel):
    """LinearOperator::coarsen (linear_operator.cc:10-23): the matrix-free stencil algebra must
    reproduce R A R^T of the oracle on every level, entry by entry (1e-12 relative)."""
    op = oracle.Operator.prior(n, pde, Lambda=0.2)
    H = oracle.Hierarchy(op, nlevel)
    desc = capi.make_desc(n[0], n[1], nlevel, pde=pde, Lambda=0.2)
    nx, ny = n
    for level in range(nlevel):
        st, nc = m.host_stencil(desc, level)
        A_ref = H.level_op(level).csr().toarray()
        A = _stencil_matrix(st, nx, ny)
        assert np.abs(A - A_ref).max() <= 1e-12 * np.abs(A_ref).max(), f"level {level}"
        assert nc == H.ncolours(level)
        nx, ny = nx // 2, ny // 2


def test_create_rejects_what_the_reference_rejects():
    # lattice2d.hh:198-213: odd extent / no interior vertex -> exit(-1) in the reference
    out = np.zeros((9, 5, 5))
    for nx, nlevel in ((6, 3), (4, 3)):
        desc = capi.make_desc(nx, nx, nlevel)
        with pytest.raises(m.MgmcError) as e:
            m.host_stencil(desc, 0)
        assert e.value.code == -1


def _planes_matrix(P, nx, ny):
    """Dense matrix of a per-vertex radius-1 operator given as (9, ny + 1, nx + 1) planes."""
    w, h = nx - 1, ny - 1
    A = np.zeros((w * h, w * h))
    for j in range(1, ny):
        for i in range(1, nx):
            for dj in range(-1, 2):
                for di in range(-1, 2):
                    ii, jj = i + di, j + dj
                    v = P[(dj + 1) * 3 + di + 1, j, i]
                    if 1 <= ii < nx and 1 <= jj < ny:
                        A[(j - 1) * w + i - 1, (jj - 1) * w + ii - 1] = v
                    else:
                        assert v == 0.0  # nothing points at the boundary
    return A


@pytest.mark.parametrize("n,nlevel", [((32, 32), 4), ((48, 16), 3), ((16, 64), 4)])
def test_periodic_kappa_operators_match_oracle_triple_product(oracle, n, nlevel):
    """PeriodicCorrelationLengthModel (correlationlength_model.hh:83-113) in ShiftedLaplaceFDOperator
    (shiftedlaplace_fd_operator.cc:33-56) and LinearOperator::coarsen (linear_operator.cc:10-23): the per-vertex
    coefficient planes of every level reproduce the oracle's matrices entry by entry (1e-12 relative), with the same
    number of sweep colours."""
    Lmin, Lmax = 0.12, 0.37
    op = oracle.Operator.prior(n, "shiftedlaplace_fd", Lambda_min=Lmin, Lambda_max=Lmax)
    H = oracle.Hierarchy(op, nlevel)
    desc = capi.make_desc(n[0], n[1], nlevel, kappa_sq=m.periodic_kappa_sq(n[0], n[1], Lmin, Lmax))
    nx, ny = n
    for level in range(nlevel):
        P, nc = m.host_coefficients(desc, level)
        A_ref = H.level_op(level).csr().toarray()
        A = _planes_matrix(P, nx, ny)
        assert np.abs(A - A_ref).max() <= 1e-12 * np.abs(A_ref).max(), f"level {level}"
        assert np.abs(A - A.T).max() <= 1e-13 * np.abs(A).max()
        assert nc == H.ncolours(level)
        nx, ny = nx // 2, ny // 2
    # the constant model through the same path reproduces the stencil algebra
    desc_c = capi.make_desc(32, 32, 3, kappa_sq=np.full(31 * 31, 1.0 / 0.2 ** 2))
    desc_s = capi.make_desc(32, 32, 3, Lambda=0.2)
    for level in range(3):
        P, nc = m.host_coefficients(desc_c, level)
        st, nc_s = m.host_stencil(desc_s, level)
        nl = 32 >> level
        assert nc == nc_s
        assert np.abs(_planes_matrix(P, nl, nl) - _stencil_matrix(st, nl, nl)).max() <= 1e-12 * np.abs(st).max()


def test_variable_kappa_rejections():
    ks = m.periodic_kappa_sq(16, 16, 0.1, 0.3)
    with pytest.raises(m.MgmcError) as e:  # only the shifted Laplacian carries a variable correlation length on the device
        m.host_coefficients(capi.make_desc(16, 16, 2, pde="squared_shiftedlaplace_fd", kappa_sq=ks), 0)
    assert e.value.code == -2
    with pytest.raises(m.MgmcError) as e:
        m.host_coefficients(capi.make_desc(16, 16, 2), 0)  # constant coefficients: mgmc_host_stencil
    assert e.value.code == -1
    bad = ks.copy()
    bad[3] = -1.0
    with pytest.raises(m.MgmcError) as e:
        m.host_coefficients(capi.make_desc(16, 16, 2, kappa_sq=bad), 0)
    assert e.value.code == -1
