// hostsim.cpp — TEST INFRASTRUCTURE.  Compiles nbody_ai_b200/csrc/nbt_core.cuh for the host so
// that the integrator's arithmetic (the exact functions the CUDA kernel calls per thread) can be
// checked against the oracle in the GPU-less container (pytest -m "not gpu") and used to
// calibrate the sub-step rule.  Never loaded by the product package.
#include <cstring>
#include <vector>

#include "../../nbody_ai_b200/csrc/nbt_core.cuh"
#include "../../nbody_ai_b200/csrc/nbt_ias15.cuh"
#include "../../nbody_ai_b200/csrc/nbt_whfast.cuh"

template <int NP>
static void run_np(const nbt_run_args& A, int first, int count) {
    const bool series = (A.flags & NBT_F_SERIES) != 0 && A.series != nullptr;
    for (int g = first; g < first + count; g++) {
        // both drift variants are exercised: even launch slots run the latency (branch-free) one
        if (series) nbt_simulate_system<NP, true, false>(A, A.scheme, A.order ? A.order[g] : g);
        else if (g % 2 == 0) nbt_simulate_system<NP, false, true>(A, A.scheme, A.order ? A.order[g] : g);
        else nbt_simulate_system<NP, false, false>(A, A.scheme, A.order ? A.order[g] : g);
    }
}

extern "C" int hostsim_run(const nbt_config* cfg, const nbt_inputs* in, nbt_outputs* out) {
    nbt_run_args A;
    std::memset(&A, 0, sizeof A);
    A.params = in->params; A.substeps = (cfg->substeps > 0) ? nullptr : in->substeps; A.order = in->order;
    A.tt_offsets = in->tt_offsets;
    A.n_days_sys = in->n_days_sys; A.n_outputs_sys = in->n_outputs_sys;
    A.tt = out->tt; A.n_tt = out->n_tt; A.mean_a = out->mean_a; A.mean_e = out->mean_e;
    A.mean_inc = out->mean_inc; A.mean_omega = out->mean_omega; A.status = out->status;
    A.rv = out->rv; A.orbit_xz = out->orbit_xz; A.series = out->series;
    A.B = cfg->n_systems; A.n_outputs = cfg->n_outputs; A.substeps_all = cfg->substeps;
    A.tt_mode = cfg->tt_mode; A.flags = cfg->flags; A.orbit_stride = cfg->orbit_stride > 0 ? cfg->orbit_stride : 4;
    A.n_days = cfg->n_days;
    if (cfg->scheme == NBT_SCHEME_IAS15) {   // the reference's own integrator, one system at a time (k_integrate_ias15)
        nbt_ias15_tables T;
        nbt_ias15_make_tables(T);
        for (int b = 0; b < A.B; b++)
            switch (cfg->n_bodies - 1) {
            case 1: nbt_simulate_system_ias15<1>(A, T, b); break;
            case 2: nbt_simulate_system_ias15<2>(A, T, b); break;
            case 3: nbt_simulate_system_ias15<3>(A, T, b); break;
            case 4: nbt_simulate_system_ias15<4>(A, T, b); break;
            default: return -1;
            }
        return 0;
    }
    if (cfg->scheme == NBT_SCHEME_WHFAST) {   // the reference's 'whfast' preset (k_integrate_whfast)
        for (int b = 0; b < A.B; b++)
            switch (cfg->n_bodies - 1) {
            case 1: nbt_simulate_system_whfast<1>(A, b); break;
            case 2: nbt_simulate_system_whfast<2>(A, b); break;
            case 3: nbt_simulate_system_whfast<3>(A, b); break;
            case 4: nbt_simulate_system_whfast<4>(A, b); break;
            default: return -1;
            }
        return 0;
    }
    // NBT_SCHEME_AUTO as in enqueue_path: the trailing cfg.alt_systems launch slots take C4, the rest SBABC3
    const bool automatic = cfg->scheme == NBT_SCHEME_AUTO;
    int nalt = automatic ? cfg->alt_systems : 0;
    if (nalt < 0) nalt = 0;
    if (nalt > A.B) nalt = A.B;
    for (int pass = 0; pass < 2; pass++) {
        const int first = pass == 0 ? 0 : A.B - nalt, count = pass == 0 ? A.B - nalt : nalt;
        A.scheme = nbt_make_scheme(automatic ? (pass == 0 ? NBT_SCHEME_SBABC3 : NBT_SCHEME_C4) : cfg->scheme);
        switch (cfg->n_bodies - 1) {
        case 1: run_np<1>(A, first, count); break;
        case 2: run_np<2>(A, first, count); break;
        case 3: run_np<3>(A, first, count); break;
        case 4: run_np<4>(A, first, count); break;
        case 5: run_np<5>(A, first, count); break;
        case 6: run_np<6>(A, first, count); break;
        case 7: run_np<7>(A, first, count); break;
        case 8: run_np<8>(A, first, count); break;
        default: return -1;
        }
    }
    return 0;
}

// custom scheme (for calibration sweeps): a[S], b[S+1]
extern "C" int hostsim_run_scheme(const nbt_config* cfg, const nbt_inputs* in, nbt_outputs* out, int S,
                                  const double* a, const double* b, const double* g) {
    nbt_run_args A;
    std::memset(&A, 0, sizeof A);
    A.params = in->params; A.substeps = (cfg->substeps > 0) ? nullptr : in->substeps; A.order = in->order;
    A.tt_offsets = in->tt_offsets;
    A.n_days_sys = in->n_days_sys; A.n_outputs_sys = in->n_outputs_sys;
    A.tt = out->tt; A.n_tt = out->n_tt; A.mean_a = out->mean_a; A.mean_e = out->mean_e;
    A.mean_inc = out->mean_inc; A.mean_omega = out->mean_omega; A.status = out->status;
    A.rv = out->rv; A.orbit_xz = out->orbit_xz; A.series = out->series;
    A.B = cfg->n_systems; A.n_outputs = cfg->n_outputs; A.substeps_all = cfg->substeps;
    A.tt_mode = cfg->tt_mode; A.flags = cfg->flags; A.orbit_stride = cfg->orbit_stride > 0 ? cfg->orbit_stride : 4;
    A.n_days = cfg->n_days;
    A.scheme.S = S;
    for (int i = 0; i < NBT_MAX_STAGES; i++) A.scheme.a[i] = i < S ? a[i] : 0.0;
    for (int i = 0; i <= NBT_MAX_STAGES; i++) A.scheme.b[i] = i <= S ? b[i] : 0.0;
    for (int i = 0; i <= NBT_MAX_STAGES; i++) A.scheme.g[i] = (g && i <= S) ? g[i] : 0.0;
    switch (cfg->n_bodies - 1) {
    case 1: run_np<1>(A, 0, A.B); break;
    case 2: run_np<2>(A, 0, A.B); break;
    case 3: run_np<3>(A, 0, A.B); break;
    case 4: run_np<4>(A, 0, A.B); break;
    default: return -1;
    }
    return 0;
}

// single Kepler drift, for unit tests against the closed-form two-body solution
extern "C" int hostsim_kepler(double mu, double h, double* state6) {
    int status = 0;
    double r, ir;
    nbt_radius(state6[0], state6[1], state6[2], r, ir);
    nbt_kepler_drift<true>(mu, h, state6[0], state6[1], state6[2], state6[3], state6[4], state6[5], r, ir, status);
    return status;
}

// nbt_acos over an array, for the accuracy test
extern "C" void hostsim_acos(const double* x, double* y, int n) {
    for (int i = 0; i < n; i++) y[i] = nbt_acos(x[i]);
}
