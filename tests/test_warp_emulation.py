"""CPU: the host side of osc_warp.cuh -- the emulated 32-lane warp osc_core3.cuh runs on in
tests/host_core -- against plain numpy: shuffles, reductions, ballot, and the mma.m8n8k4
fragment layout (lane 4g+t holds A[g][t], B[t][g], D[g][2t], D[g][2t+1])."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r"""
#include "%s/operational-space-control_b200/csrc/osc_warp.cuh"
using namespace osc;
extern "C" {
void xchg16(const double* in, double* out) { Var<double> a, b; for (int l = 0; l < 32; ++l) a[l] = in[l]; Warp::xchg16(b, a); for (int l = 0; l < 32; ++l) out[l] = b[l]; }
void group4(const double* in, int r, double* out) { Var<double> a, b; for (int l = 0; l < 32; ++l) a[l] = in[l]; Warp::group4(b, a, r); for (int l = 0; l < 32; ++l) out[l] = b[l]; }
double wsum(const double* in) { Var<double> a; for (int l = 0; l < 32; ++l) a[l] = in[l]; return Warp::sum(a); }
double wmax(const double* in) { Var<double> a; for (int l = 0; l < 32; ++l) a[l] = in[l]; return Warp::max(a); }
unsigned ballot(const int* in) { Var<bool> a; for (int l = 0; l < 32; ++l) a[l] = in[l] != 0; return Warp::ballot(a); }
void max16(const double* in /*[16][32]*/, double* out) { Var<double> m[16]; double sc[16]; for (int q = 0; q < 16; ++q) for (int l = 0; l < 32; ++l) m[q][l] = in[32 * q + l]; Warp::max16(m, out, sc, 0); }
void max8(const double* in /*[8][32]*/, double* out) { Var<double> m[8]; double sc[8]; for (int q = 0; q < 8; ++q) for (int l = 0; l < 32; ++l) m[q][l] = in[32 * q + l]; Warp::maxn<8>(m, out, sc, 0); }
void mma(const double* a, const double* b, double* d0, double* d1) {
  Var<double> A, B, D0, D1;
  for (int l = 0; l < 32; ++l) { A[l] = a[l]; B[l] = b[l]; D0[l] = d0[l]; D1[l] = d1[l]; }
  Warp::mma884(D0, D1, A, B);
  for (int l = 0; l < 32; ++l) { d0[l] = D0[l]; d1[l] = D1[l]; }
}
}
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("warp")
    src = os.path.join(d, "w.cpp")
    open(src, "w").write(SRC % ROOT)
    so = os.path.join(d, "libw.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    L = C.CDLL(so)
    L.wsum.restype = C.c_double
    L.wmax.restype = C.c_double
    L.ballot.restype = C.c_uint
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_shuffles_and_reductions(lib):
    rng = np.random.default_rng(0)
    v = rng.standard_normal(32)
    out = np.zeros(32)
    lib.xchg16(_p(v), _p(out))
    assert np.array_equal(out, v[np.arange(32) ^ 16])
    for r in range(4):
        lib.group4(_p(v), r, _p(out))
        assert np.array_equal(out, v[(np.arange(32) & ~3) | r])
    assert abs(lib.wsum(_p(v)) - v.sum()) < 1e-13
    a = np.abs(v)
    assert lib.wmax(_p(a)) == a.max()
    bits = (rng.random(32) < 0.5).astype(np.int32)
    assert lib.ballot(_p(bits)) == sum(int(b) << i for i, b in enumerate(bits))
    m = np.abs(rng.standard_normal((16, 32)))
    o16 = np.zeros(16)
    lib.max16(_p(m), _p(o16))
    assert np.array_equal(o16, m.max(axis=1))
    o8 = np.zeros(8)
    lib.max8(_p(m), _p(o8))
    assert np.array_equal(o8, m[:8].max(axis=1))


def test_mma_fragment_layout(lib):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((8, 4))
    B = rng.standard_normal((4, 8))
    D = rng.standard_normal((8, 8))
    lanes = np.arange(32)
    g, t = lanes >> 2, lanes & 3
    a = np.ascontiguousarray(A[g, t])
    b = np.ascontiguousarray(B[t, g])
    d0 = np.ascontiguousarray(D[g, 2 * t])
    d1 = np.ascontiguousarray(D[g, 2 * t + 1])
    lib.mma(_p(a), _p(b), _p(d0), _p(d1))
    R = D + A @ B
    np.testing.assert_allclose(d0, R[g, 2 * t], rtol=0, atol=1e-14)
    np.testing.assert_allclose(d1, R[g, 2 * t + 1], rtol=0, atol=1e-14)
