"""CPU: the generated `autogen_functions.cc` (tools/gen_presets.py) -- closed-form bodies behind
the C interface of the CasADi file the reference compiles -- against the
oracle's QP assembly at q = 0, and against itself for q != 0 (the functions are affine /
constant in q: f(q) = H q + f(0), beq(q) = beq(0) - Aeq q, bineq(q) = -Aineq q)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "operational-space-control_b200")


def _lib(robot, tmp_path):
    src = os.path.join(PKG, robot, "autogen", "autogen_functions.cc")
    so = os.path.join(tmp_path, f"libautogen_{robot}.so")
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-o", so, src], check=True)
    return C.CDLL(so)


def _call(L, name, args, out_size):
    dp = C.POINTER(C.c_double)
    arr = (dp * len(args))(*[a.ctypes.data_as(dp) if a is not None else None for a in args])
    out = np.full(out_size, np.nan)
    res = (dp * 1)(out.ctypes.data_as(dp))
    fn = getattr(L, name)
    fn.restype = C.c_int
    assert fn(arr, res, None, None, 0) == 0
    return out


@pytest.mark.parametrize("robot,config", [("walter_sr", "standing"), ("walter_sr_wheels", "stairs"),
                                          ("unitree_go2", "go2_standing")])
def test_generated_casadi_interface_matches_oracle(oracle, tmp_path, robot, config):
    import osc_b200 as ob
    spec = ob.load_preset(robot)
    L = _lib(robot, str(tmp_path))
    nv, nu, nz, n, ns = spec.nv, spec.nu, 3 * spec.nc, spec.n, spec.ns
    inp = ob.synth.make_inputs(spec, 4, config)
    rng = np.random.default_rng(3)
    for e in range(4):
        M, Cv, J, bias, tg, mask = (np.ascontiguousarray(inp[k][e]) for k in
                                    ("M", "C", "J", "bias", "targets", "mask"))
        Ho, fo, Ao, lo, uo = oracle.build_qp(spec, M, Cv, J, bias, tg, mask)
        # the reference's transformMatrix copies (:517-522): row-major -> column-major
        Mc, Jc = np.asfortranarray(M).ravel("F"), None
        Jc = np.ascontiguousarray(J[3 * ns - nz:3 * ns, :].T)     # contact_jacobian nv x nz
        Jcc, Jt, tc = Jc.ravel("F"), J.ravel("F"), tg.ravel("F")
        q0 = np.zeros(n)
        Aeq = _call(L, "Aeq", [q0, Mc, Cv, Jcc], nv * n).reshape(n, nv).T
        beq = _call(L, "beq", [q0, Mc, Cv, Jcc], nv)
        Ain = _call(L, "Aineq", [q0], 4 * spec.nc * n).reshape(n, 4 * spec.nc).T
        bin_ = _call(L, "bineq", [q0], 4 * spec.nc)
        H = _call(L, "H", [q0, tc, Jt, bias], n * n).reshape(n, n).T
        f = _call(L, "f", [q0, tc, Jt, bias], n)
        np.testing.assert_array_equal(Aeq, Ao[:nv])
        np.testing.assert_array_equal(Ain, Ao[nv:nv + 4 * spec.nc])
        np.testing.assert_array_equal(beq, lo[:nv])
        np.testing.assert_array_equal(beq, uo[:nv])
        np.testing.assert_array_equal(bin_, uo[nv:nv + 4 * spec.nc])
        sc = np.abs(Ho).max()
        np.testing.assert_allclose(H, Ho, rtol=1e-13, atol=1e-13 * sc)
        np.testing.assert_allclose(f, fo, rtol=1e-12, atol=1e-12 * np.abs(fo).max())
        # q != 0
        q = rng.standard_normal(n)
        np.testing.assert_allclose(_call(L, "f", [q, tc, Jt, bias], n), H @ q + f,
                                   rtol=1e-11, atol=1e-9 * np.abs(f).max())
        np.testing.assert_allclose(_call(L, "beq", [q, Mc, Cv, Jcc], nv), beq - Aeq @ q,
                                   rtol=1e-11, atol=1e-10)
        np.testing.assert_allclose(_call(L, "bineq", [q], 4 * spec.nc), -Ain @ q,
                                   rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(_call(L, "H", [q, tc, Jt, bias], n * n).reshape(n, n).T, H)
    # the life-cycle shims the reference's evaluate_function calls (walter_sr/utilities.h:43-77)
    for fn in ("beq", "Aeq", "bineq", "Aineq", "H", "f"):
        getattr(L, fn + "_incref")()
        assert getattr(L, fn + "_checkout")() == 0
        getattr(L, fn + "_release")(0)
        getattr(L, fn + "_decref")()
