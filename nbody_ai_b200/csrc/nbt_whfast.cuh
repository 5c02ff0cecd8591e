// nbt_whfast.cuh — NBT_SCHEME_WHFAST: the integrator='whfast' preset of nbody/simulation.py:39-42
//     sim.integrator = "whfast"; sim.ri_whfast.corrector = 5; sim.ri_whfast.safe_mode = 0     (dt: REBOUND's default 1e-3 d)
// with REBOUND's semantics [3P] (Rein & Tamayo 2015; symplectic correctors: Wisdom, Holman & Touma 1996), in the
// product's own arithmetic — the Jacobi state, Kepler drift and interaction accelerations of nbt_core.cuh:
//   * drift-kick-drift Wisdom-Holman map in Jacobi coordinates; with safe_mode = 0 the half drifts of consecutive steps
//     merge and the state is synchronised (half drift, inverse corrector) only at the end of a sim.integrate(t);
//   * the corrector of order 5 is applied when a synchronised state starts stepping:  C = Z(-2a,-b51) Z(-a,-b52)
//     Z(a,b52) Z(2a,b51),  Z(a,b) = D(a) K(-b) D(-2a) K(b) D(a),  a = sqrt(7/40) dt,  b51 = -beta dt/6,
//     b52 = 5 beta dt/6,  beta = 1/(48 sqrt(7/40)),  and inverted at synchronisation;
//   * exact_finish_time = 1: the step that would overshoot an output time is replaced, after a synchronisation, by
//     one of the remaining length; dt returns to the last full step afterwards.
// One system per thread, the general sampler of nbt_ias15.cuh.  Not a throughput path: ~42 steps per hourly output.
#pragma once
#include "nbt_core.cuh"
#include "nbt_ias15.cuh"

template <int NP>
struct nbt_whfast {
    nbt_state<NP> s;
    nbt_consts<NP> c;
    double t, dt, dt_last_done;
    bool synced;
    long long n_steps, max_steps;
    int status;
};

// (out of line: a step of the corrector is 20 drifts and 8 kicks; inlined, the NP = 8 kernel was 240 000 SASS lines)
template <int NP>
NBT_HD_COLD void nbt_wh_drift(nbt_whfast<NP>& W, double h) {
    for (int i = 0; i < NP; i++)
        nbt_kepler_drift<true>(W.c.GM[i], h, W.s.x[i], W.s.y[i], W.s.z[i], W.s.vx[i], W.s.vy[i], W.s.vz[i], W.s.r[i], W.s.ir[i], W.status);
}

template <int NP>
NBT_HD_COLD void nbt_wh_kick(nbt_whfast<NP>& W, double h) {
    double ax[NP], ay[NP], az[NP];
    nbt_accel<NP>(W.s, W.c, ax, ay, az);
    for (int i = 0; i < NP; i++) {
        W.s.vx[i] = fma(h, ax[i], W.s.vx[i]); W.s.vy[i] = fma(h, ay[i], W.s.vy[i]); W.s.vz[i] = fma(h, az[i], W.s.vz[i]);
    }
}

template <int NP>
NBT_HD void nbt_wh_Z(nbt_whfast<NP>& W, double a, double b) {
    nbt_wh_drift(W, a); nbt_wh_kick(W, -b); nbt_wh_drift(W, -2.0 * a); nbt_wh_kick(W, b); nbt_wh_drift(W, a);
}

template <int NP>
NBT_HD_COLD void nbt_wh_corrector(nbt_whfast<NP>& W, double inv) {
    const double a1 = 0.41833001326703777398908601289259374469640768464934, a2 = 2.0 * a1;
    const double b51 = -0.0083001986759332891664501193034244790614369381874871;
    const double b52 = 0.041500993379666445832250596517122395307184690937435;
    const double dt = W.dt;
    nbt_wh_Z(W, -a2 * dt, -inv * b51 * dt);
    nbt_wh_Z(W, -a1 * dt, -inv * b52 * dt);
    nbt_wh_Z(W, a1 * dt, inv * b52 * dt);
    nbt_wh_Z(W, a2 * dt, inv * b51 * dt);
}

template <int NP>
NBT_HD void nbt_wh_synchronize(nbt_whfast<NP>& W) {
    if (W.synced) return;
    nbt_wh_drift(W, 0.5 * W.dt);
    nbt_wh_corrector(W, -1.0);
    W.synced = true;
}

template <int NP>
NBT_HD void nbt_wh_step(nbt_whfast<NP>& W) {
    if (W.synced) { nbt_wh_corrector(W, 1.0); nbt_wh_drift(W, 0.5 * W.dt); }
    else nbt_wh_drift(W, W.dt);
    W.t += 0.5 * W.dt;
    nbt_wh_kick(W, W.dt);
    W.synced = false;
    W.t += 0.5 * W.dt;
    W.dt_last_done = W.dt;
    W.n_steps++;
}

// sim.integrate(tmax) with exact_finish_time = 1
template <int NP>
NBT_HD void nbt_wh_integrate_to(nbt_whfast<NP>& W, double tmax) {
    double last_full_dt = W.dt;
    bool last_step = false;
    for (;;) {
        if (W.t + W.dt >= tmax) {               // the next step would overshoot
            if (W.t == tmax) break;
            if (last_step) {
                double tscale = 1e-12 * fabs(tmax);
                if (tscale < 1e-200) tscale = 1e-12;
                if (fabs(W.t - tmax) < tscale) break;
                nbt_wh_synchronize(W);
                W.dt = tmax - W.t;
            } else {
                last_step = true;
                nbt_wh_synchronize(W);
                if (W.dt_last_done != 0.0) last_full_dt = W.dt_last_done;
                W.dt = tmax - W.t;
            }
        }
        if (W.n_steps >= W.max_steps || (W.status & (NBT_S_KEPLER | NBT_S_NONFINITE))) { W.status |= NBT_S_NONFINITE; break; }
        nbt_wh_step(W);
    }
    nbt_wh_synchronize(W);
    W.dt = last_full_dt;
}

// Whole-system driver of NBT_SCHEME_WHFAST (the form of nbt_simulate_system for the reference's 'whfast' preset)
template <int NP>
NBT_HD void nbt_simulate_system_whfast(const nbt_run_args& A, int sys) {
    const double* par = A.params + (size_t)sys * (NP + 1) * NBT_NPAR;
    nbt_whfast<NP> W;
    nbt_setup<NP>(par, W.s, W.c);
    W.status = nbt_uncalibrated_impl(par, NP + 1) ? NBT_S_UNCALIBRATED : 0;
    const int NO = (A.n_outputs_sys != nullptr) ? A.n_outputs_sys[sys] : A.n_outputs;
    const double n_days = (A.n_days_sys != nullptr) ? A.n_days_sys[sys] : A.n_days;
    const double step = n_days / (double)(NO - 1);
    W.t = 0.0; W.dt = 1e-3; W.dt_last_done = 0.0; W.synced = true; W.n_steps = 0;      // REBOUND's default dt, in days
    W.max_steps = (long long)(4.0 * (n_days / W.dt + NO)) + 1000;
    nbt_sampler<NP> Q;
    nbt_sampler_init<NP>(Q);
    for (int o = 0; o < NO; o++) {
        const double t = (o == NO - 1) ? n_days : (double)o * step;
        nbt_wh_integrate_to<NP>(W, t);
        if (!nbt_sample<NP>(A, W.c, W.s, sys, o, NO, n_days, t, Q, W.status)) break;
    }
    nbt_sampler_finish<NP>(A, Q, sys, NO, W.status);
}
