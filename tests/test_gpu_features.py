"""GPU tests of the round-2 additions, through the C ABI: per-system output grids and the batched training-set
generator, the fused on-device likelihoods, the periodogram peak reduction, the full element series (omega, M),
refined transit roots against true roots of the oracle trajectory, the IAS15 kernel and the verified mode."""
import numpy as np
import pytest

import oracle
from nbody_ai_b200 import _abi, dataset, engine, nested, workloads
from conftest import TOL_OC_MIN, TOL_TT_S, assert_parity_or_chaotic, compare_products, parity_report
from test_verified import HARD

pytestmark = pytest.mark.gpu
FULL = _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC


def oracle_run(plan):
    buf = engine.Buffers(plan)
    oracle.run(plan.cfg, plan.params, buf)
    return buf


def test_series_omega_and_M_vs_oracle(gpu_lib):
    """NBT_F_SERIES: every column of sim_data (x, y, z, P, a, e, inc, Omega, omega, M) against the oracle on eccentric
    orbits — omega and M in [0, 2 pi) like REBOUND's orbit conversion; both integrator kernels."""
    P = workloads.c1_params(2)
    P[:, 1:, 2] = [[0.05, 0.02, 0.08], [0.03, 0.06, 0.01]]
    P[:, 1:, 5] = [[1.0, 4.0, 5.5], [0.3, 2.9, 6.1]]
    P[:, 1:, 6] = [[0.5, 3.0, 2.0], [4.4, 0.1, 1.1]]
    for lanes in (0, 2):
        plan = engine.Plan(P, 40, 961, flags=_abi.F_SERIES | _abi.F_MEANS | _abi.F_MEAN_OMEGA, lanes=lanes)
        got, want = engine.run(plan), oracle_run(plan)
        g, w = got.series, want.series
        for j in range(3):
            c = 3 + 10 * j
            assert np.abs(g[:, :, c:c + 3] - w[:, :, c:c + 3]).max() < 1e-10
            assert np.abs(g[:, :, c + 3] / w[:, :, c + 3] - 1).max() < 1e-9 and np.abs(g[:, :, c + 4] / w[:, :, c + 4] - 1).max() < 1e-9
            assert np.abs(g[:, :, c + 5] - w[:, :, c + 5]).max() < 1e-9 and np.abs(g[:, :, c + 6] - w[:, :, c + 6]).max() < 1e-9
            for col in (7, 8, 9):          # Omega, omega, M: as angles (a sample next to the cut may sit on either side)
                d = np.abs(np.angle(np.exp(1j * (g[:, :, c + col] - w[:, :, c + col]))))
                assert d.max() < 1e-6, (j, col, d.max())
            for col in (8, 9):             # omega, M wrapped to [0, 2 pi)
                assert g[:, :, c + col].min() >= 0.0 and g[:, :, c + col].max() < 2 * np.pi
            assert (w[:, :, c + 8] > np.pi).any()
        dom = np.abs(got.mean_omega - want.mean_omega)
        assert np.all((dom < 1e-6) | (np.abs(dom - 2 * np.pi / 961) < 1e-6))     # a sample on the other side of the cut


def test_refined_roots_vs_true_roots(gpu_lib):
    """NBT_TT_REFINED (north_star item 2: Newton on the step's own interpolant) against the TRUE roots of
    x_planet - x_star on the oracle's IAS15 trajectory (bisection/Newton with re-integration): the refined mode is
    two orders of magnitude closer to them than the reference-compatible grid-linear mode."""
    par = np.concatenate([workloads.c1_params(1), workloads.random_systems(6, 3, seed=5)[[0, 2, 4]]])
    lin = engine.run(engine.Plan(par, 120, 2881, flags=_abi.F_MEANS))
    ref = engine.run(engine.Plan(par, 120, 2881, flags=_abi.F_MEANS, tt_mode=_abi.TT_REFINED))
    worst_lin = worst_ref = 0.0
    for b in range(len(par)):
        for j in range(3):
            true = oracle.refined_transit_times(par[b], 120, 2881, j)
            assert len(true) == lin.n_transits(b, j) == ref.n_transits(b, j)
            if not len(true):
                continue
            worst_lin = max(worst_lin, np.abs(lin.tt_of(b, j) - true).max() * 86400)
            worst_ref = max(worst_ref, np.abs(ref.tt_of(b, j) - true).max() * 86400)
    assert worst_ref < 0.02 and worst_lin > 10 * worst_ref, (worst_lin, worst_ref)


def test_per_system_grids_vs_oracle(gpu_lib):
    """generate_simulations.py:45-50: every draw on its own grid (Ndays = P1 * periods, Noutputs = round(P1 * periods *
    24)), one launch; planet-1 transit times / ephemeris / amplitude against the oracle run on the same grids."""
    blk = dataset.draw_block(64, 0, np.random.default_rng(11))
    par = blk.reshape(64, 3, 7)
    nd, no = dataset.grids(par, 60)
    assert len(np.unique(no)) > 50
    plan = engine.Plan(par, nd, no, flags=_abi.F_PLANET1_ONLY | _abi.F_MEANS)
    got, want = engine.run_verified(plan), oracle_run(plan)
    assert_parity_or_chaotic(got, want, plan, max_excluded_frac=0.1)
    # the same systems one at a time on scalar grids: same bits as inside the batch
    for b in (0, 17, 63):
        one = engine.run(engine.Plan(par[b:b + 1], float(nd[b]), int(no[b]), substeps=plan.substeps[b:b + 1], flags=plan.cfg.flags))
        fast = engine.run(plan)
        assert np.array_equal(one.tt_of(0, 0), fast.tt_of(b, 0)) and one.mean_e[0] == fast.mean_e[2 * b]


def test_training_set_batched_vs_oracle(gpu_lib, tmp_path):
    """The batched generator (one launch per block of candidates + omega variants): rejection rule and y values
    against the oracle for every row, reference pickle layout."""
    import pickle
    f = str(tmp_path / "ttv.pkl")
    st = {}
    X, y = dataset.generate_training_set(samples=24, omegas=2, periods=40, file=f, rng=np.random.default_rng(3), stats=st,
                                         verified=True)
    assert len(X) == len(y) == 24 * 3 and st["launches"] <= 2 and st["simulations"] == st["candidates"] * 3
    Xl, yl = pickle.load(open(f, 'rb'))
    assert len(Xl) == len(X) and all(len(r) == 7 for r in X)
    par = np.zeros((len(X), 3, 7))
    for i, r in enumerate(X):
        par[i, 0, 0], par[i, 1, 1], par[i, 2, 1], par[i, 2, 5], par[i, 2, 2] = r[0], r[1], r[3], r[5], r[6]
        par[i, 1, 0], par[i, 2, 0] = r[2] * dataset.mearth / dataset.msun, r[4] * dataset.mearth / dataset.msun
        par[i, 1:, 3] = np.pi / 2
    nd, no = dataset.grids(par, 40)
    plan = engine.Plan(par, nd, no, flags=_abi.F_PLANET1_ONLY)
    want = oracle_run(plan)
    bad = 0
    for i in range(len(X)):
        tc = want.tt_of(i, 0)[:40]
        if i % 3 == 0:      # accepted draws pass the reference's rule in the oracle too
            assert want.n_tt[2 * i] >= 40 and want.ttv_max[2 * i] * 1440 <= 180 * 1.001
        if len(tc) != len(y[i]) or np.abs(tc - y[i]).max() * 86400 >= TOL_TT_S:
            bad += 1
    assert bad <= 2          # chaotic rows only
    for g in range(0, len(X), 3):
        assert X[g][:5] == X[g + 1][:5] == X[g + 2][:5] and X[g][5] != X[g + 1][5]


def test_fused_nested_likelihood(gpu_lib):
    """NBT_F_LOGLIKE: nlfit's myloglike (nested.py:85-95) chained behind the forward model on the device — only
    loglike[B] comes back — against the two-call form and the oracle, both losses, penalty branch."""
    objects = [{'m': 1.22}, {'m': 10.4 * 0.000954, 'P': 0.94145, 'inc': np.pi / 2, 'e': 0, 'omega': 0},
               {'m': 3e-4, 'P': 2.1, 'inc': np.pi / 2, 'e': 0.02, 'omega': 1.0}]
    bounds = [[0.0, 1.0], {'P': [0.9414, 0.9415]}, {'P': [1.9, 2.4], 'm': [1.5e-5, 4.2e-4], 'e': [0, 0.1], 'omega': [0, 2 * np.pi]}]
    rng = np.random.default_rng(0)
    ep, ttv0, tt0 = nested.get_ttv(objects, ndays=79)
    xx = np.arange(0, 56, 2, dtype=float)
    yy = tt0[xx.astype(int)] + rng.normal(0, 0.5 / 1440, len(xx))
    yerr = np.full(len(xx), 0.5 / 1440)
    B = 96
    cubes = np.column_stack([rng.normal(tt0[0], 1e-4, B), rng.normal(0.94145, 1e-6, B), rng.uniform(0.9414, 0.9415, B),
                             rng.uniform(1.9, 2.4, B), rng.uniform(1.5e-5, 4.2e-4, B), rng.uniform(0, 0.1, B), rng.uniform(0, 2 * np.pi, B)])
    for loss in ("linear", "soft_l1"):
        L = nested.NestedLikelihood(xx, yy, yerr, objects, bounds, myloss=loss)
        assert L.ndays == nested.nlfit_ndays(xx, objects) == 78      # round(1.5 * 55 * 0.94145)
        got = L.loglike_batch(cubes)
        fwd = nested.get_ttv_batch(L.params_for(cubes), ndays=L.ndays)
        two = nested.loglike_batch(fwd, cubes[:, 0], cubes[:, 1], xx, yy, yerr, loss=loss)
        assert np.array_equal(got, two)
        for i in (0, 5, 50, 95):
            _, ttv, _ = fwd[i]
            want = oracle.loglike(ttv, cubes[i, 0], cubes[i, 1], xx, yy, yerr, loss=loss)
            assert abs(got[i] - want) <= 1e-9 * abs(want) + 1e-12
    # epochs beyond the integration: the reference's except branch
    L2 = nested.NestedLikelihood(np.append(xx, 20.0), np.append(yy, 20.0), np.append(yerr, 1e-3), objects, bounds)
    L2.ndays = 10
    ll = L2.loglike_batch(cubes[:8])
    fwd = nested.get_ttv_batch(L2.params_for(cubes[:8]), ndays=10)
    for i in range(8):
        want = oracle.loglike(fwd[i][1], cubes[i, 0], cubes[i, 1], L2.xx, L2.yy, L2.yerr)
        assert abs(ll[i] - want) <= 1e-9 * abs(want)


def test_fused_ttvfit_likelihood(gpu_lib):
    from nbody_ai_b200.ttv_fit import NbodyLikelihood
    from nbody_ai_b200.tools import mearth, msun
    truth = [{'m': 1.12}, {'m': 80 * mearth / msun, 'P': 3.2888, 'inc': np.pi / 2, 'e': 0, 'omega': 0},
             {'m': 40 * mearth / msun, 'P': 7.0, 'inc': np.pi / 2, 'e': 0, 'omega': 0}]
    t, ser, _, _ = oracle.integrate_series(_abi.objects_to_params(truth), 80, 80 * 24 + 1)
    tc1 = oracle.transit_times(ser[:, 3], ser[:, 0], t)[2:20]
    rng = np.random.default_rng(9)
    data = [{}, {'Tc': tc1 + rng.normal(0, 2e-4, 18), 'Tc_err': np.full(18, 2e-4)}, {}]
    bounds = [{}, {'P': [3.28, 3.30]}, {'P': [6.9, 7.1], 'm': [20 * mearth / msun, 60 * mearth / msun], 'omega': [-np.pi, np.pi]}]
    L = NbodyLikelihood(data, truth, bounds)
    thetas = np.array([L.prior_transform(u) for u in rng.uniform(size=(64, 4))])
    assert np.array_equal(L.loglike_batch(thetas), L.loglike_batch(thetas, fused=False))


def test_periodogram_peaks_on_device(gpu_lib):
    """NBT_F_LS_PEAKS: the three highest local maxima of every O-C periodogram, reduced on the device, against numpy
    on the full power arrays; and with ls_oc_power = NULL the power arrays are never copied back."""
    par = workloads.c2_params(40, 2, seed=8)
    plan = engine.Plan(par, 365.0, 8760, flags=FULL | _abi.F_LS_PEAKS)
    full = engine.run(plan)
    lean = engine.run(plan, engine.Buffers(plan, skip=("ls_oc_power", "ls_oc_nfreq")))
    assert lean.ls_oc_power is None and np.array_equal(lean.ls_oc_peaks, full.ls_oc_peaks, equal_nan=True)
    assert np.array_equal(lean.tt, full.tt, equal_nan=True) and np.array_equal(lean.ttv_max, full.ttv_max)
    checked = 0
    for b in range(plan.B):
        for j in range(2):
            pk = full.ls_oc_peaks[b * 2 + j]
            if full.n_transits(b, j) < 3:
                assert np.isnan(pk).all()
                continue
            f, p = full.ls_oc_of(b, j)
            k = np.flatnonzero((p[1:-1] > p[:-2]) & (p[1:-1] >= p[2:])) + 1
            k = k[np.argsort(-p[k], kind="stable")][:3]
            for r in range(3):
                if r < len(k):
                    assert pk[r, 1] == p[k[r]] and abs(pk[r, 0] - f[k[r]]) < 1e-15
                else:
                    assert np.isnan(pk[r]).all()
            checked += 1
    assert checked > 60


def test_ias15_kernel_vs_oracle(gpu_lib):
    """NBT_SCHEME_IAS15 on the GPU: the reference's own integrator, same algorithm as the oracle in the product's
    arithmetic — agreement far inside the tolerances, on the README system and on systems no fixed step resolves."""
    par = np.concatenate([workloads.c1_params(1)[:, :3], workloads.random_systems(12000, 2, seed=1001)[HARD]])
    plan = engine.Plan(par, 365.0, 8760, scheme="ias15", flags=FULL | _abi.F_RV)
    got, want = engine.run(plan), oracle_run(plan)
    w = compare_products(got, want, plan, tol_tt=1e-2, tol_oc=1e-5, tol_ae=1e-9)
    assert np.abs(got.rv - want.rv).max() < 1e-7 * np.abs(want.rv).max()
    for b in range(plan.B):
        for j in range(2):
            if got.n_transits(b, j) >= 3:
                assert np.abs(got.ls_oc_of(b, j)[1] - want.ls_oc_of(b, j)[1]).max() < 1e-5


def test_verified_mode_on_hard_systems(gpu_lib):
    """engine.run_verified on the GPU: systems flagged NBT_S_UNCALIBRATED are re-run at twice the resolution and, where
    that does not settle them, with IAS15; afterwards every non-chaotic system is inside the tolerances."""
    allp = workloads.random_systems(12000, 2, seed=1001)
    idx = np.concatenate([np.arange(0, 600, 3), np.array(HARD)])
    plan = engine.Plan(allp[idx], 365.0, 8760, flags=FULL)
    want = oracle_run(plan)
    fast = engine.run(plan)
    wf, _, _ = parity_report(fast, want, plan)
    flagged = (fast.status & _abi.S_UNCALIBRATED) != 0
    assert np.all(wf[~flagged] < 1.0) and not np.all(wf[-len(HARD):] < 1.0)
    rep = {}
    got = engine.run_verified(plan, report=rep)
    assert rep["flagged"] == int(flagged.sum()) and rep["ias15"] >= 2 and rep["accepted_doubling"] >= 1
    excluded, worst = assert_parity_or_chaotic(got, want, plan, max_excluded_frac=0.05)
    wv, _, _ = parity_report(got, want, plan)
    assert np.all(wv[-len(HARD):] < 1.0)
    for b in np.flatnonzero(~flagged)[:50]:
        assert np.array_equal(got.tt_of(b, 0), fast.tt_of(b, 0))


def test_broken_systems_read_nan(gpu_lib):
    """A system that breaks up stops at the next output; the optional series it no longer writes read NaN (never a
    previous call's data), in host-pointer and device-pointer mode alike."""
    import torch
    par = workloads.c1_params(4)
    par[1, 2, 1] = par[1, 1, 1] * 1.02            # two Jupiter-class planets on almost the same orbit: collision course
    par[1, 2, 0] = par[1, 1, 0] = 5e-3
    plan = engine.Plan(par, 200.0, 4801, flags=_abi.F_MEANS | _abi.F_RV | _abi.F_ORBIT | _abi.F_SERIES, lanes=0)
    dbuf = engine.DeviceBuffers(plan)
    for t in (dbuf.rv, dbuf.orbit_xz, dbuf.series):
        t.fill_(123.0)
    engine.run_device(plan, dbuf)
    torch.cuda.synchronize()
    h = dbuf.to_host()
    if h.status[1] & _abi.S_NONFINITE:
        assert np.isnan(h.rv[-1, 1]) and np.isnan(h.series[-1, 1]).all() and np.isnan(h.orbit_xz[-1, 1]).all()
        assert not np.isnan(h.rv[:, 0]).any() and not (h.rv == 123.0).any() and not (h.series == 123.0).any()


def test_whfast_preset_vs_oracle(gpu_lib):
    """generate(objects, ..., integrator='whfast') (nbody/simulation.py:39-42): the k_integrate_whfast kernel — WHFast at
    REBOUND's dt = 1e-3 d, symplectic corrector 5, safe_mode 0 — against the oracle's restatement of the same preset
    (tight) and against the oracle's IAS15 (the parity tolerances), through the batched ABI and the drop-in."""
    from nbody_ai_b200 import simulation
    par = np.concatenate([workloads.c1_params(1), workloads.random_systems(6, 3, seed=5)])
    plan = engine.Plan(par, 60.0, 1441, scheme="whfast", flags=FULL)
    got, want = engine.run(plan), oracle_run(plan)
    compare_products(got, want, plan, tol_tt=1e-4, tol_oc=1e-6, tol_ae=1e-9)
    ias_plan = engine.Plan(par, 60.0, 1441, scheme="ias15", flags=FULL)
    compare_products(got, oracle_run(ias_plan), ias_plan)
    sim_data = simulation.generate(workloads.c1_objects(), 30, 721, integrator='whfast')
    _, ser, _ = oracle.integrate_series_whfast(workloads.c1_params(1)[0], 30, 721)
    for j in range(3):
        assert np.abs(sim_data['pdata'][j]['x'] - ser[:, 3 + 10 * j]).max() < 1e-11
        assert np.abs(sim_data['pdata'][j]['e'] - ser[:, 8 + 10 * j]).max() < 1e-9
    assert np.abs(sim_data['star']['x'] - ser[:, 0]).max() < 1e-13


def test_solar_system_amplitude_ranges(gpu_lib):
    """solarsystem_nbody.py: the eight planets for 1000 years at 1-hour cadence, one planet per lane
    (k_integrate_lanes<8>) — the TTV amplitudes the script's comment quotes (:32-38: Mercury 0-1, Venus 3-5, Earth 4-6,
    Mars 20-40, Jupiter 500-1000, Saturn 1000-1500 min; read off a plot, every planet started at f = omega = 0, so an
    order-of-magnitude known answer only: the inner four within ~30 %, the giants — a few dozen transits of a 900-year
    term — within a factor of a few)."""
    from nbody_ai_b200.tools import mearth, msun
    objects = [{'m': 1.},
               {'m': 0.0553 * mearth / msun, 'P': 87.969, 'inc': 3.14159 / 2, 'e': 0.205},
               {'m': 0.815 * mearth / msun, 'P': 224.701, 'inc': 3.14159 / 2, 'e': 0.007},
               {'m': mearth / msun, 'P': 365.24, 'inc': 3.14159 / 2, 'e': 0.017},
               {'m': 0.107 * mearth / msun, 'P': 686.980, 'inc': 3.14159 / 2, 'e': 0.094},
               {'m': 317.83 * mearth / msun, 'P': 4332.589, 'inc': 3.14159 / 2, 'e': 0.049},
               {'m': 95.16 * mearth / msun, 'P': 10759.22, 'inc': 3.14159 / 2, 'e': 0.056},
               {'m': 14.54 * mearth / msun, 'P': 30588.740, 'inc': 3.14159 / 2, 'e': 0.0457},
               {'m': 17.15 * mearth / msun, 'P': 59799.9, 'inc': 3.14159 / 2, 'e': 0.0113}]
    par = _abi.objects_to_params(objects)[None]
    plan = engine.Plan(par, 1000 * 365, 1000 * 365 * 24, flags=0)          # solarsystem_nbody.py:24
    assert plan.cfg.lanes_systems == 1
    got = engine.run(plan)
    assert not got.status[0] & (_abi.S_NONFINITE | _abi.S_UNBOUND | _abi.S_TT_OVERFLOW)
    amp = got.ttv_max[:8] * 1440
    print("solar-system TTV max-avg [min]:", np.round(amp, 2), "transits", got.n_tt[:8])
    assert list(got.n_tt[:4]) == [4149, 1625, 1000, 532] or np.all(np.abs(got.n_tt[:4] - [4149, 1625, 1000, 532]) <= 1)
    for j, (lo, hi) in enumerate([(0.5, 2.0), (3.0, 6.0), (4.0, 8.0), (20.0, 55.0), (300.0, 3000.0), (800.0, 8000.0)]):
        assert lo < amp[j] < hi, (j, amp[j])


def test_pipeline_matches_sequential_runs(gpu_lib):
    """engine.Pipeline (two nbt_run calls in flight from two host threads, own streams and pinned pools): every batch's
    products are bit-identical to a plain engine.run of the same batch, also for ragged sizes and mixed flags."""
    batches = [workloads.c2_params(30 + 7 * k, 2, seed=40 + k) for k in range(5)]
    pipe = engine.Pipeline(depth=2)
    try:
        for start in (0, 2):                      # results stay valid for `depth` submissions: check them pairwise
            futs = [pipe.submit(batches[k], 120.0, 2881, flags=FULL) for k in range(start, min(start + 2, 5))]
            for k, f in zip(range(start, start + 2), futs):
                plan, buf = f.result()
                want = engine.run(engine.Plan(batches[k], 120.0, 2881, flags=FULL))
                for name in ("tt", "ttv", "n_tt", "period", "t0", "ttv_max", "mean_a", "mean_e", "mean_omega", "status", "ls_oc_power"):
                    assert np.array_equal(getattr(buf, name), getattr(want, name), equal_nan=True), (k, name)
        plan, buf = pipe.submit(batches[4], 60.0, 1441, flags=_abi.F_PLANET1_ONLY).result()
        want = engine.run(engine.Plan(batches[4], 60.0, 1441, flags=_abi.F_PLANET1_ONLY))
        assert np.array_equal(buf.tt, want.tt, equal_nan=True) and np.array_equal(buf.n_tt, want.n_tt)
    finally:
        pipe.close()
