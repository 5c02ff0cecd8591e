"""The C-ABI library loads without a GPU, exports every symbol include/nbt.h declares, plans on
the host, and FAILS LOUDLY (no CPU fallback) when asked to compute without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nbody_ai_b200 import _abi, engine, workloads
from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "nbt.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nbt_[a-z0-9_]+)\s*\(", txt)))


def test_exports_every_declared_symbol(nbt):
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(nbt, s), "libnbt.so does not export %s" % s
    assert nbt.nbt_abi_version() == _abi.NBT_ABI_VERSION


def test_struct_layout_matches_header(nbt):
    # nbt_run rejects a config whose struct_size disagrees with the C side
    cfg = _abi.make_config(1, 3, 10.0, 11, substeps=1)
    off = np.zeros(3, dtype=np.int64)
    par = workloads.c1_params(1)[:, :3].copy()
    inp = _abi.nbt_inputs()
    inp.params = _abi.ptr(par)
    assert nbt.nbt_plan_tt(C.byref(cfg), C.byref(inp), _abi.ptr(off, C.c_int64)) > 0
    cfg.struct_size += 4
    assert nbt.nbt_plan_tt(C.byref(cfg), C.byref(inp), _abi.ptr(off, C.c_int64)) == -2
    assert b"struct_size" in nbt.nbt_last_error()


def test_plan_helpers(nbt):
    par = workloads.c2_params(50, 2, seed=3)
    plan = engine.Plan(par, 365.0, 8760, flags=_abi.F_MEANS | _abi.F_LS_OC)
    B, Np = plan.B, plan.Np
    cap = np.diff(plan.tt_offsets)
    want = np.minimum(np.ceil(1.05 * 365.0 / par[:, 1:, 1]) + 3, 8760).astype(np.int64).ravel()
    assert np.array_equal(cap, want) and plan.cfg.tt_cap_max == cap.max()
    # launch order: a permutation sorted by decreasing substeps
    if plan.order is not None:
        assert sorted(plan.order) == list(range(B))
        assert np.all(np.diff(plan.substeps[plan.order]) <= 0)
    # O-C periodogram capacity = astropy grid for the largest possible transit count
    lcap = np.diff(plan.ls_offsets)
    for c, l in zip(cap[:20], lcap[:20]):
        assert l == (1 + int(np.round((0.5 - 1 / 180.) / (1.0 / (c - 1.0) / 10))) if c >= 3 else 0)
    # default scheme (AUTO): three drift stages per SBABC3 step, two per C4 step (negative counts)
    k = plan.substeps
    assert plan.n_steps == (int(k[k > 0].sum()) * 3 - int(k[k < 0].sum()) * 2) * 8759
    # sub-step rule: P_min / dt_sub >= 57.5 for SBABC3; C4 only at one step per output and P_min >= 250 steps
    step = 365.0 / 8759
    pmin = par[:, 1:, 1].min(1)
    assert np.all(pmin[k > 0] / (step / k[k > 0]) >= 57.5 - 1e-9)
    assert np.all(k[k < 0] == -1) and np.all(pmin[k < 0] >= 250 * step) and (k < 0).any()
    sb = engine.Plan(par, 365.0, 8760, scheme="sbabc3")
    assert np.array_equal(sb.substeps, np.abs(k)) and sb.n_steps == int(sb.substeps.sum()) * 8759 * 3


def test_invalid_arguments(nbt):
    cfg = _abi.make_config(1, 1, 10.0, 11, substeps=1)
    assert nbt.nbt_count_steps(C.byref(cfg), None) == -1          # n_bodies < 2
    cfg = _abi.make_config(1, 3, 10.0, 1, substeps=1)
    assert nbt.nbt_count_steps(C.byref(cfg), None) == -1          # n_outputs < 2
    # NBT_SCHEME_AUTO: every launch slot inside cfg.alt_systems must hold a system the planner marked
    par = workloads.c2_params(40, 0, seed=1)
    plan = engine.Plan(par, 365.0, 8760)
    assert 0 < plan.cfg.alt_systems < plan.B
    buf = engine.Buffers(plan)
    inp, out = buf.as_structs()
    plan.cfg.alt_systems += 1
    assert nbt.nbt_run(C.byref(plan.cfg), C.byref(inp), C.byref(out), None) == -13
    assert b"alt_systems" in nbt.nbt_last_error()
    with pytest.raises(ValueError):
        _abi.objects_to_params([{'m': 1.0}, {'m': 1e-3}])          # planet without a period
    with pytest.raises(ValueError):
        _abi.objects_to_params([{'m': 1.0}, {'m': 1e-3, 'P': 3.0, 'a': 0.04}])  # both P and a (rebound.add raises too)
    with pytest.raises(ValueError):
        _abi.objects_to_params([{'m': 1.0}, {'m': 1e-3, 'P': 3.0, 'l': 0.1}])   # unsupported key
    # a semi-major axis instead of a period: the Jacobi mu of rebound.add
    p = _abi.objects_to_params([{'m': 1.0}, {'m': 1e-3, 'a': 0.05}, {'m': 2e-3, 'a': 0.1}])
    assert abs(p[1, 1] - 2 * np.pi * np.sqrt(0.05 ** 3 / (_abi.NBT_G * 1.001))) < 1e-14
    assert abs(p[2, 1] - 2 * np.pi * np.sqrt(0.1 ** 3 / (_abi.NBT_G * 1.003))) < 1e-14
    # per-system grids come in pairs and exclude the [n_outputs][B] products; likelihood needs its arrays
    plan = engine.Plan(par, 365.0, 8760)
    buf = engine.Buffers(plan)
    inp, out = buf.as_structs()
    nd = np.full(plan.B, 100.0)
    inp.n_days_sys = _abi.ptr(nd)
    assert nbt.nbt_run(C.byref(plan.cfg), C.byref(inp), C.byref(out), None) == -14
    plan.cfg.flags |= _abi.F_LOGLIKE
    inp.n_days_sys = _abi.ptr(None)
    assert nbt.nbt_run(C.byref(plan.cfg), C.byref(inp), C.byref(out), None) == -15
    plan.cfg.flags = _abi.F_LS_PEAKS
    assert nbt.nbt_run(C.byref(plan.cfg), C.byref(inp), C.byref(out), None) == -16


def test_per_system_grids_plan(nbt):
    """Every draw of the training-set generator has its own grid (generate_simulations.py:45-50): the planner sizes
    capacities, sub-steps and step counts per system."""
    par = workloads.random_systems(64, 2, seed=5)
    nd = par[:, 1, 1] * 50
    no = np.round(nd * 24).astype(np.int32)
    plan = engine.Plan(par, nd, no, flags=_abi.F_PLANET1_ONLY)
    assert plan.cfg.n_outputs == no.max() and plan.cfg.n_days == nd.max()
    cap = np.diff(plan.tt_offsets).reshape(-1, 2)
    assert np.array_equal(cap[:, 0], np.minimum(np.ceil(1.05 * nd / par[:, 1, 1]) + 3, no)) and np.all(cap[:, 1] == 1)
    k = plan.substeps
    step = nd / (no - 1)
    assert np.array_equal(np.abs(k), np.maximum(np.ceil(step * 57.5 / par[:, 1:, 1].min(1)), 1))
    assert plan.n_steps == int((np.where(k > 0, 3 * k, -2 * k) * (no - 1)).sum())
    assert np.array_equal(plan.times_of(3), np.linspace(0, nd[3], no[3]))
    # uniform arrays are the scalar plan
    uni = engine.Plan(par, np.full(64, 100.0), np.full(64, 2401), flags=_abi.F_PLANET1_ONLY)
    sca = engine.Plan(par, 100.0, 2401, flags=_abi.F_PLANET1_ONLY)
    assert np.array_equal(uni.tt_offsets, sca.tt_offsets) and np.array_equal(uni.substeps, sca.substeps) and uni.n_steps == sca.n_steps


def test_uncalibrated_envelope(nbt):
    par = workloads.c1_params(1)[0]
    assert nbt.nbt_uncalibrated(_abi.ptr(par), 4) == 0
    q = par.copy(); q[2, 2] = 0.2
    assert nbt.nbt_uncalibrated(_abi.ptr(q), 4) == 1                 # e > 0.1
    q = par.copy(); q[2, 0] = 0.02 * q[0, 0]
    assert nbt.nbt_uncalibrated(_abi.ptr(q), 4) == 1                 # heavy planet
    q = par.copy(); q[2, 1] = 4.2
    assert nbt.nbt_uncalibrated(_abi.ptr(q), 4) == 1                 # pair inside 4 mutual Hill radii
    rs = workloads.random_systems(2000, 2, seed=9)
    plan = engine.Plan(rs, 365.0, 8760)
    frac = plan.uncalibrated().mean()
    assert 0.03 < frac < 0.25


def test_objects_to_params_defaults():
    p = _abi.objects_to_params(workloads.c1_objects())
    assert p.shape == (4, 7) and p[1, _abi.P_P] == 3.2888 and p[1, _abi.P_INC] == 3.14159 / 2
    assert np.all(p[:, _abi.P_OMEGA_NODE] == 0) and np.all(p[:, _abi.P_F] == 0)


@pytest.mark.skipif(_abi.load().nbt_device_count() > 0, reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(nbt):
    """Without a CUDA device every compute entry returns an error; the Python layer raises."""
    par = workloads.c1_params(2)
    plan = engine.Plan(par, 10.0, 241, flags=_abi.F_MEANS)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        engine.run(plan)
    from nbody_ai_b200 import simulation, tools
    with pytest.raises(RuntimeError, match="no CUDA device"):
        simulation.generate(workloads.c1_objects(), 10, 241)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        tools.lomb_scargle(np.arange(10.0), np.arange(10.0) ** 2)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        tools.maxavg(np.arange(5.0))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        tools.TTV(np.arange(5.0), np.arange(5.0))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under nbody_ai_b200/ may reference it."""
    pkg = os.path.join(ROOT, "nbody_ai_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libnbt_oracle" not in txt, fn


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No libnbt.so -> the loader raises; nothing falls back to the CPU."""
    monkeypatch.setattr(_abi, "_LIB", None)
    monkeypatch.setattr(_abi, "_LIB_PATH", str(tmp_path / "libnbt.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _abi.load()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.Plan(workloads.c1_params(1), 10.0, 241)


def test_oracle_is_only_used_by_test_infrastructure():
    """oracle/ may be imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs only."""
    allowed = {os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")}
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "gpurun_out", "__pycache__", "oracle", "tests", ".pytest_cache")]
        for fn in files:
            if fn.endswith(".py"):
                path = os.path.join(dirpath, fn)
                txt = open(path).read()
                if "import oracle" in txt or "from oracle" in txt:
                    assert path in allowed, path
