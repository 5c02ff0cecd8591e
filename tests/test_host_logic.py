"""Host-side mirror of the reference interfaces: parameter mapping, planning and dict layouts that
need no GPU (everything numeric on the path is CUDA and is covered by the -m gpu tests)."""
import numpy as np
import pytest

from nbody_ai_b200 import _abi, engine, nested, simulation, tools, workloads
from nbody_ai_b200.ttv_fit import NbodyLikelihood
from conftest import golden


def test_constants_and_scalar_helpers_match_reference():
    """nbody/tools.py:11-59: constants, sa, hill_sphere against the reference's own outputs."""
    ref = golden("ref_functions.npz")
    assert tools.G == 0.00029591220828559104 and tools.msun == 1.989e30 and tools.mearth == 5.972e24
    assert np.allclose([tools.sa(1.12, 3.2888), tools.sa(0.8, 20.0)], ref["sa"], rtol=1e-15)
    objs = workloads.c1_objects()
    hill = [tools.hill_sphere(objs, i=i) for i in (1, 2, 3)]
    assert np.allclose(hill, ref["hill"], rtol=1e-10)
    assert abs(tools.m2r_star(1.0) - 1.0) < 1e-15 and abs(tools.m2r_star(0.855) - 0.89) < 1e-12
    with pytest.raises(ValueError):
        tools.m2r_star(50.0)


def test_vectorised_priors_agree_with_scalar_hill_sphere():
    par = workloads.random_systems(40, 2, seed=4)
    for b in range(40):
        objs = [{'m': par[b, 0, 0]}, {'m': par[b, 1, 0], 'P': par[b, 1, 1]}, {'m': par[b, 2, 0], 'P': par[b, 2, 1]}]
        a1, a2 = tools.sa(par[b, 0, 0], par[b, 1, 1]), tools.sa(par[b, 0, 0], par[b, 2, 1])
        h1, h2 = tools.hill_sphere(objs, 1), tools.hill_sphere(objs, 2)
        assert a1 + h1 <= a2 - h2                       # the reference's acceptance rule (simulation.py:87)
        hb = workloads.hill_sphere_batch(par[b:b + 1, 0, 0], par[b:b + 1, 1, 0], par[b:b + 1, 1, 1])[0]
        assert abs(hb / h1 - 1) < 1e-9
    assert np.all((par[:, 1, 1] >= 0.8) & (par[:, 1, 1] <= 20) & (par[:, 0, 0] >= 0.75) & (par[:, 0, 0] <= 1.15))
    ratio = par[:, 2, 0] / par[:, 1, 0]
    assert np.all((ratio >= 0.1) & (ratio <= 10)) and np.all(par[:, 1:, 3] == np.pi / 2)


def test_workload_shapes():
    assert workloads.c1_params(3).shape == (3, 4, 7)
    c2 = workloads.c2_params(5, 6, seed=0)
    assert c2.shape == (35, 3, 7)
    assert np.array_equal(c2[0:7, :, :5], np.repeat(c2[0:1, :, :5], 7, axis=0))      # variants differ in omega_2 only
    assert len(set(c2[0:7, 2, 5])) == 7
    c3, meta = workloads.c3_params(10)
    assert c3.shape == (10, 3, 7) and meta["noutputs"] == meta["ndays"] * 48
    c4, meta4 = workloads.c4_params(4)
    assert c4.shape == (16, 3, 7) and meta4["noutputs"] == 34080 and np.all(c4[:, 1, 1] == 1.42)
    assert workloads.c5_params(20).shape == (20, 5, 7)


def test_plan_flags_and_views(nbt):
    par = workloads.c1_params(2)
    plan = engine.Plan(par, 365, 8760, flags=_abi.F_MEANS | _abi.F_LS_OC | _abi.F_RV | _abi.F_LS_RV | _abi.F_ORBIT)
    buf = engine.Buffers(plan)
    assert buf.rv.shape == (8759, 2) and buf.orbit_xz.shape == (2190, 2, 3, 2) and plan.rv_nfreq == 1816
    assert buf.series is None and len(buf.tt) == plan.tt_offsets[-1]
    assert plan.cfg.lanes_systems == 2                    # small 3-planet batch -> one planet per lane
    assert np.array_equal(plan.times, np.linspace(0, 365, 8760))
    f, p = buf.ls_rv_of(0)
    assert len(f) == 1816 and abs(f[0] - 1 / 365) < 1e-18
    p1 = engine.Plan(par, 60, 2881, flags=_abi.F_PLANET1_ONLY)
    assert list(np.diff(p1.tt_offsets)) == [23, 1, 1, 23, 1, 1]
    with pytest.raises(ValueError):
        engine.Plan(par[:, :, :6], 10, 241)
    with pytest.raises(ValueError):
        engine.Plan(par, 10, 1)


def test_sampler_parameter_mapping():
    objects = [{'m': 1.22}, {'m': 0.01, 'P': 0.94145, 'inc': np.pi / 2}, {'m': 1e-4, 'P': 2.0, 'e': 0.0, 'omega': 0.0, 'inc': np.pi / 2}]
    bounds = [[0.1, 0.2], {'P': [0.94, 0.95]}, {'P': [1.9, 2.4], 'm': [1e-5, 4e-4], 'omega': [-np.pi, np.pi]}]
    out = nested.apply_params(objects, bounds, [0.15, 0.9412, 2.1, 2e-4, 1.0])
    assert out[1]['P'] == 0.9412 and out[2]['P'] == 2.1 and out[2]['m'] == 2e-4 and out[2]['omega'] == 1.0
    assert objects[2]['P'] == 2.0                          # the input list is not mutated
    assert nested.nlfit_ndays(np.arange(56.0), objects) == int(np.round(1.5 * 56 * 0.94145))
    data = [{}, {'Tc': 0.3 + 0.94145 * np.array([0, 1, 2, 5, 9.0]), 'Tc_err': np.full(5, 1e-3)}, {}]
    L = NbodyLikelihood(data, objects, [{}, {'P': [0.94, 0.95]}, {'m': [1e-5, 4e-4]}])
    assert list(L.orbit) == [0, 1, 2, 5, 9] and abs(L.sim_time - (9 * 0.94145 + 5 * 0.94145)) < 1e-12
    th = L.prior_transform(np.array([0.5, 0.5]))
    assert np.allclose(th, [0.945, 2.05e-4])
    par = L.params_for(np.array([[0.941, 3e-4], [0.949, 1e-5]]))
    assert par.shape == (2, 3, 7) and par[0, 1, 1] == 0.941 and par[1, 2, 0] == 1e-5 and par[0, 2, 1] == 2.0


def test_randomize_draws_match_reference_rng():
    """simulation.randomize() consumes np.random exactly like the reference (fixture from the reference's own
    randomize() under np.random.seed(1234)); host-only, no GPU."""
    ref = golden("ref_functions.npz")["randomize_seed1234"]
    np.random.seed(1234)
    for row in ref[:6]:
        o = simulation.randomize()
        got = [o[0]['m'], o[1]['m'], o[1]['P'], o[2]['m'], o[2]['P'], o[2]['e'], o[2]['omega'], o[1]['inc'], o[2]['inc']]
        assert np.allclose(got, row, rtol=1e-12, atol=0)


def _plan_by_helpers(lib, par, n_days, n_outputs, flags, scheme, sort, nds=None, nos=None, substeps=None):
    """The plan assembled from the separate nbt_plan_* helpers (what engine.Plan did before nbt_plan existed)."""
    import ctypes as C
    B, N = par.shape[:2]
    Np = N - 1
    cfg = _abi.make_config(B, N, n_days, n_outputs, substeps=0, scheme=scheme, flags=flags)
    inp = _abi.nbt_inputs()
    inp.params = _abi.ptr(par)
    inp.n_days_sys, inp.n_outputs_sys = _abi.ptr(nds), _abi.ptr(nos, C.c_int32)
    if substeps is None:
        sub = np.zeros(B, dtype=np.int32)
        assert lib.nbt_plan_substeps(C.byref(cfg), C.byref(inp), _abi.ptr(sub, C.c_int32)) == 0
    else:
        sub = np.ascontiguousarray(substeps, dtype=np.int32)
    inp.substeps = _abi.ptr(sub, C.c_int32)
    marked = scheme == _abi.SCHEME_AUTO and bool((sub < 0).any())
    order = np.arange(B, dtype=np.int32)
    if sort and (B > 32 or marked):
        assert lib.nbt_plan_order(C.byref(cfg), C.byref(inp), _abi.ptr(order, C.c_int32)) >= 0
    elif marked:
        order = np.concatenate([np.flatnonzero(sub >= 0), np.flatnonzero(sub < 0)]).astype(np.int32)
    tto, lso = np.zeros(B * Np + 1, dtype=np.int64), np.zeros(B * Np + 1, dtype=np.int64)
    cap = lib.nbt_plan_tt(C.byref(cfg), C.byref(inp), _abi.ptr(tto, C.c_int64))
    assert lib.nbt_plan_ls(C.byref(cfg), _abi.ptr(tto, C.c_int64), _abi.ptr(lso, C.c_int64)) == 0
    return dict(substeps=sub, order=order, tt_offsets=tto, ls_offsets=lso, cap=cap, n_steps=lib.nbt_count_steps(C.byref(cfg), C.byref(inp)),
                alt=int((sub < 0).sum()) if scheme == _abi.SCHEME_AUTO else 0)


@pytest.mark.parametrize("threads", [1, 0])
def test_one_call_planner_matches_the_helpers(nbt, threads):
    """nbt_plan (one threaded pass, counting sort) == nbt_plan_substeps + _order + _tt + _ls + nbt_count_steps, bit for
    bit: uniform and per-system grids, every scheme family, sorted and unsorted, given sub-steps, tiny batches."""
    cases = [
        (workloads.c2_params(3000, 6, seed=3), 365.0, 8760, _abi.F_MEANS | _abi.F_LS_OC, _abi.SCHEME_AUTO, True),
        (workloads.c2_params(3000, 6, seed=3), 365.0, 8760, _abi.F_MEANS | _abi.F_LS_OC, _abi.SCHEME_AUTO, False),
        (workloads.c5_params(6000, seed=5), 1095.0, 26280, _abi.F_MEANS, _abi.SCHEME_AUTO, True),
        (workloads.random_systems(5000, 3, seed=6), 200.0, 1201, _abi.F_PLANET1_ONLY | _abi.F_LS_OC, _abi.SCHEME_SBABC3, True),
        (workloads.random_systems(20, 2, seed=7), 365.0, 8760, _abi.F_MEANS, _abi.SCHEME_AUTO, True),
        (workloads.random_systems(20, 2, seed=7), 100.0, 201, _abi.F_MEANS, _abi.SCHEME_M64, True),
    ]
    for par, nd, no, fl, sch, sort in cases:
        want = _plan_by_helpers(nbt, par, nd, no, fl, sch, sort)
        plan = engine.Plan(par, nd, no, flags=fl, scheme=sch, sort=sort, threads=threads)
        assert np.array_equal(plan.substeps, want["substeps"]) and np.array_equal(plan.tt_offsets, want["tt_offsets"])
        if fl & _abi.F_LS_OC:
            assert np.array_equal(plan.ls_offsets, want["ls_offsets"])
        got_order = np.arange(plan.B, dtype=np.int32) if plan.order is None else plan.order
        assert np.array_equal(got_order, want["order"])
        assert plan.n_steps == want["n_steps"] and plan.cfg.tt_cap_max == want["cap"] and plan.cfg.alt_systems == want["alt"]
    # per-system grids (generate_simulations.py:45-50) and caller-given sub-steps
    par = workloads.random_systems(4000, 2, seed=8)
    nds = par[:, 1, 1] * 60
    nos = np.rint(par[:, 1, 1] * 60 * 24).astype(np.int32)
    want = _plan_by_helpers(nbt, par, float(nds.max()), int(nos.max()), _abi.F_PLANET1_ONLY, _abi.SCHEME_AUTO, True, nds=nds, nos=nos)
    plan = engine.Plan(par, nds, nos, flags=_abi.F_PLANET1_ONLY, threads=threads)
    assert np.array_equal(plan.substeps, want["substeps"]) and np.array_equal(plan.order, want["order"])
    assert np.array_equal(plan.tt_offsets, want["tt_offsets"]) and plan.n_steps == want["n_steps"]
    given = np.where(np.arange(4000) % 3 == 0, 2, 1).astype(np.int32)
    want = _plan_by_helpers(nbt, par, 365.0, 8760, _abi.F_MEANS, _abi.SCHEME_SBABC3, True, substeps=given)
    plan = engine.Plan(par, 365.0, 8760, substeps=given, scheme="sbabc3", threads=threads)
    assert np.array_equal(plan.order, want["order"]) and plan.n_steps == want["n_steps"]
    uni = engine.Plan(par, 365.0, 8760, substeps=2, scheme="sbabc3", threads=threads)
    assert uni.substeps is None and uni.n_steps == 4000 * 8759 * 2 * 3


def test_bench_cpu_arm_plans_without_the_cuda_library(nbt):
    """bench.py --impl reference sizes its CSR buffers with numpy (HostPlan) so that the CPU arm never loads libnbt.so;
    the layout must be the library's."""
    import bench
    for par, nd, no, fl in ((workloads.c2_params(300, 6, seed=0), 365.0, 8760, _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC),
                            (workloads.c5_params(200), 1095.0, 26280, _abi.F_MEANS),
                            (workloads.c3_params(100)[0], 79, 3792, _abi.F_PLANET1_ONLY | _abi.F_LS_OC)):
        hp, lp = bench.HostPlan(par, nd, no, fl), engine.Plan(par, nd, no, flags=fl)
        assert np.array_equal(hp.tt_offsets, lp.tt_offsets) and hp.cfg.tt_cap_max == lp.cfg.tt_cap_max
        if fl & _abi.F_LS_OC:
            assert np.array_equal(hp.ls_offsets, lp.ls_offsets)
    from nbody_ai_b200 import dataset
    par = dataset.draw_block(20, 2, np.random.default_rng(2)).reshape(-1, 3, 7)
    nd, no = dataset.grids(par, 50)
    hp, lp = bench.HostPlan(par, nd, no, _abi.F_PLANET1_ONLY), engine.Plan(par, nd, no, flags=_abi.F_PLANET1_ONLY)
    assert np.array_equal(hp.tt_offsets, lp.tt_offsets) and np.array_equal(hp.n_outputs_sys, lp.n_outputs_sys)
    assert bench.f_step(2) == 579 and bench.f_step(3) == 926 and bench.f_step(4) == 1298


def test_bench_reference_arm_line(built):
    """`bench.py --impl reference` (the driver's CPU arm): one JSON line on stdout with the contract's keys, the same
    `config` as the GPU arm's, and no libnbt.so in the process."""
    import json
    import subprocess
    import sys
    import bench
    from conftest import ROOT
    code = ("import sys, json; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-sample', '8'];"
            "import bench; bench.main(); print('LIBNBT' if 'libnbt.so' in open('/proc/self/maps').read() else 'CLEAN', file=sys.stderr)")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and "CLEAN" in r.stderr
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    args = type("A", (), dict(draws=10000, omegas=6))()
    assert d["config"] == bench.base_config(args, 1) and d["value"] > 0


def test_fits_layout_of_the_grid(tmp_path):
    """grid.write_fits against the layout simulations/simulation_grid.py:219-295 writes with astropy: a primary HDU
    (TTV amplitude [min]) + three IMAGE extensions (O-C periodicity 1st / 2nd [epochs], RV amplitude [m/s]), BUNIT,
    COMMENT, the linear axes, the objects, EXTNAME; periodicities and RV masked to NaN where the amplitude is <= 1 min
    (:174-179); plain FITS: 2880-byte blocks of 80-character cards, big-endian doubles."""
    from nbody_ai_b200 import grid
    npts = 5
    obj = workloads.c4_objects()
    mass_grid, period_grid = workloads.c4_grids(npts)
    obj[2]['m'], obj[2]['P'] = mass_grid[-1][-1], period_grid[-1][-1]
    rng = np.random.default_rng(0)
    grids = {"ttv_grid": rng.uniform(0, 3, (npts, npts)) / 1440, "per_epoch_grid": rng.uniform(2, 400, (4, npts, npts)),
             "rv_grid": rng.uniform(1, 100, (npts, npts))}
    k = workloads.MSUN_ASTROPY / workloads.MEARTH_ASTROPY
    path = str(tmp_path / "ttv_grid.fits")
    grid.write_fits(path, grids, mass_grid, period_grid, obj, k)
    raw = open(path, "rb").read()
    assert len(raw) % 2880 == 0 and raw[:30] == b"SIMPLE  =                    T"
    hdus = grid.read_fits(path)
    assert [h["EXTNAME"] for h, _ in hdus] == ["TTV_AMPLITUDE", "TTV_PERIODICITY1", "TTV_PERIODICITY2", "RV_AMPLITUDE"]
    assert [h["BUNIT"] for h, _ in hdus] == ["minutes", "epochs", "epochs", "m/s"]
    assert "XTENSION" not in hdus[0][0] and all(h["XTENSION"] == "IMAGE" for h, _ in hdus[1:])
    amp = grids["ttv_grid"] * 24 * 60                      # the reference's expression (:220)
    mask = amp > 1
    assert np.array_equal(hdus[0][1], amp)
    for (h, a), want in zip(hdus[1:], (grids["per_epoch_grid"][0], grids["per_epoch_grid"][1], grids["rv_grid"])):
        assert np.array_equal(a[mask], want[mask]) and np.isnan(a[~mask]).all() and (~mask).any() and mask.any()
    for h, a in hdus:
        assert h["BITPIX"] == -64 and h["NAXIS"] == 2 and h["NAXIS1"] == npts and h["NAXIS2"] == npts and a.shape == (npts, npts)
        assert h["CRPIX1"] == 1 and h["CRPIX2"] == 1
        assert abs(h["CRVAL1"] - 0.5) < 1e-12 and abs(h["CDELT1"] - (325 - 0.5) / npts) < 1e-12          # mass axis [Mearth]
        assert abs(h["CRVAL2"] - 1.41) < 1e-12 and abs(h["CDELT2"] - (2.41 - 1.41) / npts) < 1e-12       # period ratio
        assert h["STARMASS"] == 1.0 and h["P1_PER"] == 1.42 and h["P1_OMEGA"] == 0.01 and abs(h["P2_OMEGA"] - np.pi / 2) < 1e-15
        assert h["P2_MASS"] == obj[2]['m'] and h["P2_PER"] == obj[2]['P'] and h["P1_ECC"] == 0.0
    assert hdus[0][0]["COMMENT"] == ["TTV amplitude in minutes"] and hdus[3][0]["COMMENT"] == ["Radial velocity amplitude in m/s"]


def test_make_sin_and_weighted_periodogram_host_branches():
    """The two off-path branches of the reference's tools that stay on the host: make_sin (nbody/tools.py:78-95) and the
    dy-weighted floating-mean periodogram behind lomb_scargle(dy=ndarray) — against direct formulas / scipy."""
    from scipy.signal import lombscargle
    t = np.linspace(0, 30, 200)
    pars = np.array([[0.2, 1.5, 0.3], [0.05, 0.7, 1.1]])
    want = 1.5 * np.sin(t * 2 * np.pi * 0.2 + 0.3) + 0.7 * np.sin(t * 2 * np.pi * 0.05 + 1.1)
    assert np.allclose(tools.make_sin(t, pars), want, rtol=0, atol=1e-14)
    assert np.allclose(tools.make_sin(t, pars.ravel()), want, rtol=0, atol=1e-14)
    assert np.allclose(tools.make_sin(t, pars[:, 1:], freqs=[0.2, 0.05]), want, rtol=0, atol=1e-14)
    rng = np.random.default_rng(3)
    y = want + rng.normal(0, 0.1, len(t))
    freq = 1 / 365 + np.arange(300) / 30 / 10
    # equal weights: the weighted form reduces to the floating-mean generalised LS that scipy computes
    p = tools._weighted_ls_host(t, y, np.full(len(t), 0.1), freq)
    ps = lombscargle(t, y, 2 * np.pi * freq, floating_mean=True, normalize=True)
    assert np.abs(p - ps).max() < 1e-10 and abs(freq[np.argmax(p)] - 0.2) < 0.01
    # unequal weights change it, and down-weighting the noisy half moves it towards the clean signal's periodogram
    dy = np.where(np.arange(len(t)) % 2 == 0, 0.05, 5.0)
    pw = tools._weighted_ls_host(t, y, dy, freq)
    assert np.abs(pw - p).max() > 1e-3 and abs(freq[np.argmax(pw)] - 0.2) < 0.01
