"""CPU: the C-ABI library loads, exports every symbol include/osc_b200.h declares, validates
arguments, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu


def _declared():
    src = open(os.path.join(ROOT, "include", "osc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(osc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_entry_point():
    from osc_b200 import capi
    L = capi.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/osc_b200.h but not exported"
    assert set(capi.EXPORTS) <= set(names)


def test_default_settings_are_osqp_063_defaults():
    from osc_b200 import capi
    s = capi.default_settings()
    assert (s.rho, s.sigma, s.alpha, s.eps_abs, s.eps_rel) == (0.1, 1e-6, 1.6, 1e-3, 1e-3)
    assert (s.scaling, s.adaptive_rho, s.adaptive_rho_interval, s.max_iter, s.check_termination,
            s.warm_start, s.adaptive_rho_tolerance) == (10, 1, 0, 4000, 25, 1, 5.0)


def test_invalid_arguments_are_rejected():
    import osc_b200 as ob
    from osc_b200 import capi
    L = capi.load()
    h = C.c_void_p()
    spec = capi.c_spec(ob.load_preset("walter_sr"))
    assert L.osc_create(None, None, 4, 0, C.byref(h)) == -1
    assert L.osc_create(C.byref(spec), None, 0, 0, C.byref(h)) == -1
    spec.nv = 15  # not a compiled shape
    assert L.osc_create(C.byref(spec), None, 4, 0, C.byref(h)) == -1
    assert b"unsupported robot shape" in L.osc_last_error(None)
    assert L.osc_step(None, None) == -1 and L.osc_destroy(None) == -1


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_gpu():
    import osc_b200 as ob
    from osc_b200 import capi
    with pytest.raises(capi.OscError, match="no CPU fallback"):
        capi.BatchedOSC(ob.load_preset("unitree_go2"), 8)


def test_presets_reproduce_reference_constants():
    import osc_b200 as ob
    w = ob.load_preset("walter_sr")
    g = ob.load_preset("unitree_go2")
    # SURVEY.md 8 row a17 (autogen_defines.h values)
    assert (w.nq, w.nv, w.nu, w.ns, w.nc, w.nz, w.n, w.m, w.s) == (15, 14, 8, 17, 8, 24, 46, 92, 102)
    assert (g.nq, g.nv, g.nu, g.ns, g.nc, g.nz, g.n, g.m, g.s) == (19, 18, 12, 5, 4, 12, 42, 76, 30)
    assert w.algorithmic_bytes == 14864 and g.algorithmic_bytes == 7664
    assert w.mu == g.mu == 0.8 and w.w_reg == 1e-4 and w.w_torque == 1e-4
    hdr = open(os.path.join(ROOT, "operational-space-control_b200", "walter_sr", "autogen",
                            "autogen_defines.h")).read()
    for name, val in (("design_vector_size", 46), ("Aeq_sz", 644), ("Aineq_sz", 1472),
                      ("H_sz", 2116), ("z_idx", 46), ("u_idx", 22)):
        assert re.search(rf"constexpr int {name} = {val};", hdr)
