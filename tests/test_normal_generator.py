"""The table-driven normal generator of the CUDA path (csrc/philox.cuh, host build of the same code) against the
oracle's libm Box-Muller on the same Philox counters (oracle_core.hh Philox::normal_pair)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "multigridmc_b200", "csrc")

SRC = r'''
#include "philox.cuh"
extern "C" void host_normal_pairs(unsigned long long seed, int n, const unsigned *ctr, double *z) {
  const mgmc::PhiloxKeys K = mgmc::philox_round_keys(seed);
  for (int i = 0; i < n; ++i)
    mgmc::normal_pair(K, ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], mgmc::kNormalConstsHost, mgmc::kNormalTabHost, z[2 * i], z[2 * i + 1]);
}
extern "C" void host_box_muller(int n, const unsigned long long *ab, double *z) {
  for (int i = 0; i < n; ++i) mgmc::box_muller(ab[2 * i], ab[2 * i + 1], mgmc::kNormalConstsHost, mgmc::kNormalTabHost, z[2 * i], z[2 * i + 1]);
}
'''


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    d = tmp_path_factory.mktemp("normgen")
    src = d / "normgen.cc"
    src.write_text(SRC)
    so = d / "libnormgen.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", CSRC, str(src), "-o", str(so)])
    return C.CDLL(str(so))


def test_normal_pair_matches_oracle_libm(hostlib, oracle):
    rng = np.random.default_rng(7)
    n, seed = 20000, 5418513
    ctr = rng.integers(0, 2**32, size=(n, 4), dtype=np.uint32)
    z = np.zeros((n, 2))
    hostlib.host_normal_pairs(C.c_ulonglong(seed), n, ctr.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p))
    ref = np.array([oracle.philox_normal_pair(seed, *map(int, c)) for c in ctr])
    # both are a few ulp from the exact value; the error of z = r cos(t) is relative to r
    scale = np.maximum(np.hypot(ref[:, 0], ref[:, 1]), 1e-300)[:, None]
    assert np.max(np.abs(z - ref) / scale) < 4e-15


def test_box_muller_edge_words(hostlib):
    """u -> 0, u -> 1, every log cell and every angle cell: compared with numpy in extended precision"""
    cases = []
    for k in list(range(64)) + [2**52 - 1 - i for i in range(64)] + [(j << 47) + d for j in range(32) for d in (0, 1, 2**47 - 1)]:
        for kb in (0, 2**52 - 1, 2**51, 12345678901234):
            cases.append(((k << 12) | 0xABC, (kb << 12) | 0x123))
    for kb in [(c << 47) + d for c in range(32) for d in (0, 2**46, 2**47 - 1)]:
        cases.append(((0x8000000000000 << 12), kb << 12))
    ab = np.array(cases, dtype=np.uint64)
    z = np.zeros((len(ab), 2))
    hostlib.host_box_muller(len(ab), ab.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p))
    ld = np.longdouble
    u1 = ((ab[:, 0] >> np.uint64(12)).astype(ld) + ld(0.5)) / ld(2**52)
    u2 = ((ab[:, 1] >> np.uint64(12)).astype(ld) + ld(0.5)) / ld(2**52)
    r = np.sqrt(ld(-2) * np.log(u1))
    t = ld(2) * np.pi.astype(ld) if hasattr(np.pi, "astype") else ld(2) * ld("3.14159265358979323846264338327950288")
    ref = np.stack([r * np.cos(t * u2), r * np.sin(t * u2)], axis=1)
    err = np.abs(z.astype(ld) - ref) / r[:, None]
    assert float(np.max(err)) < 1e-15
