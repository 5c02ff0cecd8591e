// nbt_ias15.cuh — IAS15 for one system per thread (NBT_SCHEME_IAS15), and the shared per-output sampling of the
// "reference integrator" kernels.
//
// IAS15 (Rein & Spiegel 2015, MNRAS 446, 1424) is what the reference integrates with: no call site of
// nbody/simulation.py:27 passes integrator=, so rebound.Simulation() keeps its default.  The product's hot path uses
// fixed-step Wisdom-Holman compositions instead, which cannot follow a close encounter and whose
// truncation error a chaotic pair amplifies; the handful of systems where step doubling shows that
// (engine.run_verified) are re-integrated here, with the reference's own algorithm: 15th-order Gauss-Radau
// predictor-corrector in inertial barycentric coordinates, adaptive step from the last series term (epsilon = 1e-9),
// compensated summation, the step shortened to land on every output time (exact_finish_time).
//
// One thread per system, working set (b, e, g series: 3 x 7 x 3N doubles) in local memory: latency-bound by
// construction and ~30x the cost of a composition run.  Host-compilable (tests/hostsim) like nbt_core.cuh.
#pragma once
#include "nbt_core.cuh"

struct nbt_ias15_tables {
    double h[8];        // Gauss-Radau spacings
    double rr[8][8];    // 1 / (h[j] - h[k])
    double cc[7][7];    // Newton basis -> monomial:  b_k = sum_{j>=k} cc[j][k] g_j
    double dd[7][7];    // monomial -> Newton:        g_k = sum_{j>=k} dd[j][k] b_j
};

// host-side construction (long double, so that the tables are correctly rounded)
inline void nbt_ias15_make_tables(nbt_ias15_tables& T) {
    static const double H[8] = {0.0, 0.0562625605369221464656521910318, 0.180240691736892364987579942780,
                                0.352624717113169637373907769648, 0.547153626330555383001448554766,
                                0.734210177215410531523210605558, 0.885320946839095768090359771030,
                                0.977520613561287501891174488626};
    for (int j = 0; j < 8; j++) {
        T.h[j] = H[j];
        for (int k = 0; k < 8; k++) T.rr[j][k] = (k < j) ? 1.0 / (H[j] - H[k]) : 0.0;
    }
    long double c[7][7] = {}, d[7][7] = {};
    c[0][0] = 1.0L;
    for (int j = 1; j < 7; j++)        // s (s - h1) ... (s - h_j) = sum_k c[j][k] s^(k+1)
        for (int k = 0; k <= j; k++)
            c[j][k] = ((k > 0) ? c[j - 1][k - 1] : 0.0L) - (long double)H[j] * ((k < j) ? c[j - 1][k] : 0.0L);
    for (int col = 0; col < 7; col++) {   // inverse of the unit-triangular map
        long double g[7];
        for (int k = 6; k >= 0; k--) {
            long double rhs = (k == col) ? 1.0L : 0.0L;
            for (int j = k + 1; j < 7; j++) rhs -= c[j][k] * g[j];
            g[k] = rhs;
        }
        for (int k = 0; k < 7; k++) d[col][k] = g[k];
    }
    for (int j = 0; j < 7; j++)
        for (int k = 0; k < 7; k++) { T.cc[j][k] = (double)c[j][k]; T.dd[j][k] = (double)d[j][k]; }
}

template <int NB>
struct nbt_ias15 {
    double m[NB];
    double x[3 * NB], v[3 * NB], a0[3 * NB];
    double b[7][3 * NB], e[7][3 * NB], g[7][3 * NB];
    double csx[3 * NB], csv[3 * NB];
    double t, dt, dt_last;
    int have_b;
    long long n_steps, n_reject, max_steps;
    double dt_floor;   // give up below this adaptive step (0 = never): see nbt_simulate_system_ias15
    bool gave_up;
};

// direct-summation Newtonian accelerations (REBOUND gravity "basic")
template <int NB>
NBT_HD void nbt_ias15_accel(const double* m, const double* x, double* a) {
    for (int i = 0; i < 3 * NB; i++) a[i] = 0.0;
    for (int i = 0; i < NB; i++)
        for (int j = i + 1; j < NB; j++) {
            const double dx = x[3 * j] - x[3 * i], dy = x[3 * j + 1] - x[3 * i + 1], dz = x[3 * j + 2] - x[3 * i + 2];
            const double r2 = dx * dx + dy * dy + dz * dz;
            const double pre = NBT_G / (r2 * sqrt(r2));
            const double pi_ = pre * m[j], pj = pre * m[i];
            a[3 * i] += pi_ * dx; a[3 * i + 1] += pi_ * dy; a[3 * i + 2] += pi_ * dz;
            a[3 * j] -= pj * dx;  a[3 * j + 1] -= pj * dy;  a[3 * j + 2] -= pj * dz;
        }
}

NBT_HD void nbt_add_cs(double& sum, double& carry, double inc) {
    const double y = inc - carry;
    const double t = sum + y;
    carry = (t - sum) - y;
    sum = t;
}

// set-up: rebound.add(**obj) in Jacobi fashion + move_to_com(), inertial coordinates (nbody/simulation.py:44-47)
template <int NB>
NBT_HD void nbt_ias15_setup(const double* par, nbt_ias15<NB>& I) {
    double M = 0.0, cx[3] = {0, 0, 0}, cv[3] = {0, 0, 0};
    for (int i = 0; i < NB; i++) {
        const double* p = par + i * NBT_NPAR;
        const double mi = p[NBT_P_M];
        I.m[i] = mi;
        if (i == 0) {
            for (int k = 0; k < 3; k++) I.x[k] = I.v[k] = 0.0;
        } else {
            const double mu = NBT_G * (M + mi);
            const double P = p[NBT_P_P], ec = p[NBT_P_E];
            const double a = cbrt(P * P * mu / (4.0 * NBT_PI * NBT_PI));
            const double cf = cos(p[NBT_P_F]), sf = sin(p[NBT_P_F]);
            const double r = a * (1.0 - ec * ec) / (1.0 + ec * cf);
            const double v0 = sqrt(mu / (a * (1.0 - ec * ec)));
            const double cO = cos(p[NBT_P_OMEGA_NODE]), sO = sin(p[NBT_P_OMEGA_NODE]);
            const double co = cos(p[NBT_P_OMEGA_PERI]), so = sin(p[NBT_P_OMEGA_PERI]);
            const double ci = cos(p[NBT_P_INC]), si = sin(p[NBT_P_INC]);
            const double cwf = co * cf - so * sf, swf = so * cf + co * sf;
            const double vp = -(swf + ec * so), vq = (cwf + ec * co);
            const double r3[3] = {r * (cO * cwf - sO * swf * ci), r * (sO * cwf + cO * swf * ci), r * (swf * si)};
            const double v3[3] = {v0 * (cO * vp - sO * vq * ci), v0 * (sO * vp + cO * vq * ci), v0 * (vq * si)};
            for (int k = 0; k < 3; k++) { I.x[3 * i + k] = cx[k] + r3[k]; I.v[3 * i + k] = cv[k] + v3[k]; }
        }
        for (int k = 0; k < 3; k++) {
            cx[k] = (cx[k] * M + I.x[3 * i + k] * mi) / (M + mi);
            cv[k] = (cv[k] * M + I.v[3 * i + k] * mi) / (M + mi);
        }
        M += mi;
    }
    for (int i = 0; i < NB; i++)
        for (int k = 0; k < 3; k++) { I.x[3 * i + k] -= cx[k]; I.v[3 * i + k] -= cv[k]; }
    for (int i = 0; i < 3 * NB; i++) {
        I.csx[i] = I.csv[i] = 0.0;
        for (int k = 0; k < 7; k++) I.b[k][i] = I.e[k][i] = I.g[k][i] = 0.0;
    }
    I.t = 0.0; I.dt = 1e-3; I.dt_last = 0.0;   // REBOUND's default initial dt
    I.have_b = 0; I.n_steps = 0; I.n_reject = 0; I.gave_up = false;
    I.max_steps = 2000000000LL;
    I.dt_floor = 0.0;
}

// one step of size dt; true if accepted (state advanced), false if rejected (I.dt shrunk)
template <int NB>
NBT_HD bool nbt_ias15_step(nbt_ias15<NB>& I, const nbt_ias15_tables& T, double dt) {
    constexpr int n3 = 3 * NB;
    const double eps_b = 1e-9, safety = 0.25;
    double xs[n3], as[n3];
    nbt_ias15_accel<NB>(I.m, I.x, I.a0);
    if (I.have_b && dt / I.dt_last > 20.0) {
        // a very large step-size increase (the step after a short landing step): do not extrapolate the old series
        for (int i = 0; i < n3; i++)
            for (int k = 0; k < 7; k++) I.b[k][i] = I.e[k][i] = 0.0;
    } else if (I.have_b) {
        const double q = dt / I.dt_last, q2 = q * q, q3 = q2 * q, q4 = q2 * q2, q5 = q4 * q, q6 = q3 * q3, q7 = q6 * q;
        for (int i = 0; i < n3; i++) {
            const double b0 = I.b[0][i], b1 = I.b[1][i], b2 = I.b[2][i], b3 = I.b[3][i], b4 = I.b[4][i], b5 = I.b[5][i], b6 = I.b[6][i];
            double be[7];
            for (int k = 0; k < 7; k++) be[k] = I.b[k][i] - I.e[k][i];
            I.e[0][i] = q * (b6 * 7.0 + b5 * 6.0 + b4 * 5.0 + b3 * 4.0 + b2 * 3.0 + b1 * 2.0 + b0);
            I.e[1][i] = q2 * (b6 * 21.0 + b5 * 15.0 + b4 * 10.0 + b3 * 6.0 + b2 * 3.0 + b1);
            I.e[2][i] = q3 * (b6 * 35.0 + b5 * 20.0 + b4 * 10.0 + b3 * 4.0 + b2);
            I.e[3][i] = q4 * (b6 * 35.0 + b5 * 15.0 + b4 * 5.0 + b3);
            I.e[4][i] = q5 * (b6 * 21.0 + b5 * 6.0 + b4);
            I.e[5][i] = q6 * (b6 * 7.0 + b5);
            I.e[6][i] = q7 * b6;
            for (int k = 0; k < 7; k++) I.b[k][i] = I.e[k][i] + be[k];
        }
    }
    for (int i = 0; i < n3; i++)
        for (int k = 0; k < 7; k++) {
            double gk = 0.0;
            for (int j = k; j < 7; j++) gk += T.dd[j][k] * I.b[j][i];
            I.g[k][i] = gk;
        }
    double pc_err = 1e300, pc_err_last = 2.0;
    int iter = 0;
    while (true) {
        if (pc_err < 1e-16) break;
        if (iter > 2 && pc_err_last <= pc_err) break;   // stopped converging (round-off floor)
        if (iter >= 12) break;
        pc_err_last = pc_err;
        iter++;
        double max_a = 0.0, max_db6 = 0.0;
        for (int n = 1; n < 8; n++) {
            const double s = T.h[n];
            for (int i = 0; i < n3; i++) {
                double p = I.b[6][i] * (1.0 / 72.0);
                p = p * s + I.b[5][i] * (1.0 / 56.0);
                p = p * s + I.b[4][i] * (1.0 / 42.0);
                p = p * s + I.b[3][i] * (1.0 / 30.0);
                p = p * s + I.b[2][i] * (1.0 / 20.0);
                p = p * s + I.b[1][i] * (1.0 / 12.0);
                p = p * s + I.b[0][i] * (1.0 / 6.0);
                p = p * s + I.a0[i] * 0.5;
                xs[i] = I.x[i] - I.csx[i] + s * dt * (I.v[i] + s * dt * p);
            }
            nbt_ias15_accel<NB>(I.m, xs, as);
            for (int i = 0; i < n3; i++) {
                double gk = (as[i] - I.a0[i]) * T.rr[n][0];
                for (int k = 1; k < n; k++) gk = (gk - I.g[k - 1][i]) * T.rr[n][k];
                const double dg = gk - I.g[n - 1][i];
                I.g[n - 1][i] = gk;
                for (int k = 0; k < n; k++) I.b[k][i] += T.cc[n - 1][k] * dg;
                if (n == 7) {
                    const double aa = fabs(as[i]), d6 = fabs(dg);
                    if (aa > max_a) max_a = aa;
                    if (d6 > max_db6) max_db6 = d6;
                }
            }
        }
        pc_err = max_db6 / max_a;
    }
    // step-size control on the magnitude of the last series term
    double max_a = 0.0, max_b6 = 0.0;
    for (int i = 0; i < n3; i++) {
        const double aa = fabs(I.a0[i]), bb = fabs(I.b[6][i]);
        if (aa > max_a) max_a = aa;
        if (bb > max_b6) max_b6 = bb;
    }
    const double err = max_b6 / max_a;
    double dt_new = (err > 0.0 && isfinite(err)) ? dt * pow(eps_b / err, 1.0 / 7.0) : dt / safety;
    if (fabs(dt_new / dt) < safety) {   // reject and retry with the smaller step
        I.dt = dt_new;
        I.n_reject++;
        if (I.have_b) {
            for (int i = 0; i < n3; i++)
                for (int k = 0; k < 7; k++) I.b[k][i] = I.e[k][i] = 0.0;
            I.have_b = 0;
        }
        return false;
    }
    if (dt_new / dt > 1.0 / safety) dt_new = dt / safety;
    for (int i = 0; i < n3; i++) {   // advance with compensated summation
        const double b0 = I.b[0][i], b1 = I.b[1][i], b2 = I.b[2][i], b3 = I.b[3][i], b4 = I.b[4][i], b5 = I.b[5][i], b6 = I.b[6][i];
        const double dx2 = (b6 / 72.0 + b5 / 56.0 + b4 / 42.0 + b3 / 30.0 + b2 / 20.0 + b1 / 12.0 + b0 / 6.0 + I.a0[i] / 2.0) * dt * dt;
        nbt_add_cs(I.x[i], I.csx[i], dx2);
        nbt_add_cs(I.x[i], I.csx[i], I.v[i] * dt);
        const double dv = (b6 / 8.0 + b5 / 7.0 + b4 / 6.0 + b3 / 5.0 + b2 / 4.0 + b1 / 3.0 + b0 / 2.0 + I.a0[i]) * dt;
        nbt_add_cs(I.v[i], I.csv[i], dv);
    }
    I.t += dt;
    I.dt_last = dt;
    I.dt = dt_new;
    I.have_b = 1;
    I.n_steps++;
    return true;
}

// sim.integrate(t_end) with exact_finish_time = 1: the last step is shortened to land on t_end
template <int NB>
NBT_HD void nbt_ias15_integrate_to(nbt_ias15<NB>& I, const nbt_ias15_tables& T, double t_end) {
    while (I.t < t_end) {
        // step budget: a collision drives the adaptive step towards zero; give up, the caller flags the system
        if (I.n_steps + I.n_reject > I.max_steps) { I.gave_up = true; return; }
        const double remaining = t_end - I.t;
        if (remaining <= 1e-14 * fabs(t_end)) { I.t = t_end; break; }
        double dt = I.dt;
        const double keep = I.dt;
        bool shortened = false;
        if (dt >= remaining) { dt = remaining; shortened = true; }
        if (!(dt > 0.0) || !isfinite(dt) || (!shortened && dt < I.dt_floor)) { I.gave_up = true; return; }
        if (nbt_ias15_step<NB>(I, T, dt) && shortened) { I.t = t_end; I.dt = keep; }
    }
}

// ---------------------------------------------------------------------------------------------
// Per-output sampling shared by the reference-integrator kernels: transit detection on the linspace grid
// (nbody/simulation.py:163-170), RV, orbit samples, series rows and element sums — the same products
// nbt_simulate_system forms in flight — from a Jacobi state.  Not the hot path: general nbt_elements only.
// ---------------------------------------------------------------------------------------------
template <int NP>
struct nbt_sampler {
    double dx_prev[NP], dvx_prev[NP], sum_a[NP], sum_e[NP], sum_i[NP], sum_o[NP];
    int cnt[NP];
    double xs_prev, t_prev;
};

template <int NP>
NBT_HD void nbt_sampler_init(nbt_sampler<NP>& Q) {
    for (int j = 0; j < NP; j++) { Q.dx_prev[j] = Q.dvx_prev[j] = Q.sum_a[j] = Q.sum_e[j] = Q.sum_i[j] = Q.sum_o[j] = 0.0; Q.cnt[j] = 0; }
    Q.xs_prev = Q.t_prev = 0.0;
}

// returns false if the state is broken (the caller stops)
template <int NP>
NBT_HD bool nbt_sample(const nbt_run_args& A, const nbt_consts<NP>& c, nbt_state<NP>& s, int sys, int o, int NO, double n_days,
                       double t, nbt_sampler<NP>& Q, int& status) {
    const int B = A.B, W = 3 + 10 * NP;
    const bool want_means = (A.flags & NBT_F_MEANS) != 0;
    const bool want_rv = (A.flags & NBT_F_RV) != 0 && A.rv != nullptr;
    const bool want_orbit = (A.flags & NBT_F_ORBIT) != 0 && A.orbit_xz != nullptr;
    const bool want_series = (A.flags & NBT_F_SERIES) != 0 && A.series != nullptr;
    const int n_tt_planets = (A.flags & NBT_F_PLANET1_ONLY) ? 1 : NP;
    const double rv_scale = 1.496e11 * ((double)NO / (n_days * 24.0 * 60.0 * 60.0));
    for (int j = 0; j < NP; j++) nbt_radius(s.x[j], s.y[j], s.z[j], s.r[j], s.ir[j]);
    double dx[NP], dvx[NP], qy[NP], qz[NP];
    double Sx = 0.0, Svx = 0.0, Ty = 0.0, Tz = 0.0;
    for (int j = 0; j < NP; j++) {
        dx[j] = s.x[j] + Sx; dvx[j] = s.vx[j] + Svx; qy[j] = s.y[j] + Ty; qz[j] = s.z[j] + Tz;
        Sx = fma(c.mr[j], s.x[j], Sx); Svx = fma(c.mr[j], s.vx[j], Svx);
        Ty = fma(c.mr[j], s.y[j], Ty); Tz = fma(c.mr[j], s.z[j], Tz);
    }
    const double xs = -Sx;
    if (!isfinite(Sx + Svx) || (status & (NBT_S_KEPLER | NBT_S_NONFINITE))) {
        status |= NBT_S_NONFINITE;
        nbt_fill_broken(want_rv ? A.rv : nullptr, want_orbit ? A.orbit_xz : nullptr, want_series ? A.series : nullptr, B, NP, A.orbit_stride, sys, o, NO);
        return false;
    }
    if (o > 0) {
        for (int j = 0; j < n_tt_planets; j++)
            if (Q.dx_prev[j] >= 0.0 && dx[j] <= 0.0) {
                const double t0 = nbt_crossing_time(A.tt_mode, Q.t_prev, t, Q.dx_prev[j], dx[j], Q.dvx_prev[j], dvx[j]);
                const int64_t off = A.tt_offsets[(size_t)sys * NP + j];
                const int64_t cap = A.tt_offsets[(size_t)sys * NP + j + 1] - off;
                if (Q.cnt[j] < cap) A.tt[off + Q.cnt[j]] = t0; else status |= NBT_S_TT_OVERFLOW;
                Q.cnt[j]++;
            }
        if (want_rv) A.rv[(size_t)(o - 1) * B + sys] = (xs - Q.xs_prev) * rv_scale;
    }
    for (int j = 0; j < NP; j++) { Q.dx_prev[j] = dx[j]; Q.dvx_prev[j] = dvx[j]; }
    Q.xs_prev = xs; Q.t_prev = t;
    if (want_means || want_series) {
        double* row = want_series ? A.series + ((size_t)o * B + sys) * W : nullptr;
        if (row) { row[0] = xs; row[1] = -Ty; row[2] = -Tz; }
        for (int j = 0; j < NP; j++) {
            const nbt_orbit ob = want_series ? nbt_elements<2>(c.GM[j], s.x[j], s.y[j], s.z[j], s.vx[j], s.vy[j], s.vz[j], s.r[j], s.ir[j])
                                             : nbt_elements<1>(c.GM[j], s.x[j], s.y[j], s.z[j], s.vx[j], s.vy[j], s.vz[j], s.r[j], s.ir[j]);
            Q.sum_a[j] += ob.a; Q.sum_e[j] += ob.e; Q.sum_i[j] += ob.inc; Q.sum_o[j] += ob.omega;
            if (!(ob.a > 0.0)) status |= NBT_S_UNBOUND;
            if (row) {
                double* p = row + 3 + 10 * j;
                p[0] = dx[j] + xs; p[1] = qy[j] - Ty; p[2] = qz[j] - Tz;
                p[3] = ob.P; p[4] = ob.a; p[5] = ob.e; p[6] = ob.inc; p[7] = ob.Omega; p[8] = ob.omega; p[9] = ob.M;
            }
        }
    }
    if (want_orbit && (o % A.orbit_stride) == 0) {
        const int q = o / A.orbit_stride;
        for (int j = 0; j < NP; j++) {
            double* dst = A.orbit_xz + (((size_t)q * B + sys) * NP + j) * 2;
            dst[0] = dx[j] + xs; dst[1] = qz[j] - Tz;
        }
    }
    return true;
}

template <int NP>
NBT_HD void nbt_sampler_finish(const nbt_run_args& A, const nbt_sampler<NP>& Q, int sys, int NO, int status) {
    const bool want_omega = (A.flags & (NBT_F_MEAN_OMEGA | NBT_F_SERIES)) != 0;
    for (int j = 0; j < NP; j++) {
        const size_t sp = (size_t)sys * NP + j;
        A.n_tt[sp] = Q.cnt[j];
        if (A.flags & (NBT_F_MEANS | NBT_F_SERIES)) {
            const double inv = 1.0 / (double)NO;
            if (A.mean_a) A.mean_a[sp] = Q.sum_a[j] * inv;
            if (A.mean_e) A.mean_e[sp] = Q.sum_e[j] * inv;
            if (A.mean_inc) A.mean_inc[sp] = Q.sum_i[j] * inv;
            if (A.mean_omega && want_omega) A.mean_omega[sp] = Q.sum_o[j] * inv;
        }
    }
    A.status[sys] = status;
}

// inertial barycentric coordinates -> Jacobi state (primary of body i = centre of mass of bodies 0..i-1)
template <int NP>
NBT_HD void nbt_inertial_to_jacobi(const double* m, const double* x, const double* v, nbt_state<NP>& s) {
    double M = m[0], cx[3] = {x[0], x[1], x[2]}, cv[3] = {v[0], v[1], v[2]};
    for (int i = 1; i <= NP; i++) {
        s.x[i - 1] = x[3 * i] - cx[0]; s.y[i - 1] = x[3 * i + 1] - cx[1]; s.z[i - 1] = x[3 * i + 2] - cx[2];
        s.vx[i - 1] = v[3 * i] - cv[0]; s.vy[i - 1] = v[3 * i + 1] - cv[1]; s.vz[i - 1] = v[3 * i + 2] - cv[2];
        for (int k = 0; k < 3; k++) {
            cx[k] = (cx[k] * M + x[3 * i + k] * m[i]) / (M + m[i]);
            cv[k] = (cv[k] * M + v[3 * i + k] * m[i]) / (M + m[i]);
        }
        M += m[i];
    }
}

// Whole-system driver of NBT_SCHEME_IAS15 (the form of nbt_simulate_system for the reference's own integrator)
template <int NP>
NBT_HD void nbt_simulate_system_ias15(const nbt_run_args& A, const nbt_ias15_tables& T, int sys) {
    const double* par = A.params + (size_t)sys * (NP + 1) * NBT_NPAR;
    nbt_ias15<NP + 1> I;
    nbt_ias15_setup<NP + 1>(par, I);
    nbt_state<NP> s;
    nbt_consts<NP> c;
    nbt_setup<NP>(par, s, c);      // constants (G M_i, m_i / M_i); the state is overwritten from the IAS15 coordinates
    int status = nbt_uncalibrated_impl(par, NP + 1) ? NBT_S_UNCALIBRATED : 0;
    const int NO = (A.n_outputs_sys != nullptr) ? A.n_outputs_sys[sys] : A.n_outputs;
    const double n_days = (A.n_days_sys != nullptr) ? A.n_days_sys[sys] : A.n_days;
    const double step = n_days / (double)(NO - 1);
    {   // step budget: one step per output plus ~25 per innermost orbit is what a regular system needs, and a scattering
        // event a few hundred more; a captured pair (two planets orbiting each other) keeps the adaptive step small for
        // good — the run is given up (NBT_S_NONFINITE) at 2 per output + 100 per innermost orbit.  (REBOUND has no budget
        // and would crawl on; such a system is chaotic and excluded from any comparison.)  The step floor only guards
        // against a collapse to zero: 1e-9 P_min is two point masses a few km apart.
        double pmin = INFINITY;
        for (int i = 1; i <= NP; i++) {
            const double P = par[i * NBT_NPAR + NBT_P_P];
            if (P > 0.0 && P < pmin) pmin = P;
        }
        I.max_steps = 2LL * NO + 10000LL + (isfinite(pmin) ? (long long)(100.0 * n_days / pmin) : 0LL);
        I.dt_floor = isfinite(pmin) ? 1e-9 * pmin : 0.0;
    }
    nbt_sampler<NP> Q;
    nbt_sampler_init<NP>(Q);
    for (int o = 0; o < NO; o++) {
        const double t = (o == NO - 1) ? n_days : (double)o * step;
        nbt_ias15_integrate_to<NP + 1>(I, T, t);
        if (I.gave_up) status |= NBT_S_NONFINITE;
        nbt_inertial_to_jacobi<NP>(I.m, I.x, I.v, s);
        if (!nbt_sample<NP>(A, c, s, sys, o, NO, n_days, t, Q, status)) break;
    }
    nbt_sampler_finish<NP>(A, Q, sys, NO, status);
}
