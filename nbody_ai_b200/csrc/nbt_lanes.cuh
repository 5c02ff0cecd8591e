// nbt_lanes.cuh — "one planet per lane" form of the integrator (device only).
//
// A system is spread over NP consecutive lanes of a warp (32/NP systems per warp): lane j holds
// the Jacobi coordinate of planet j.  The Kepler drifts (the bulk of a stage) run in parallel across
// lanes and the interaction kick exchanges positions / pair accelerations with warp shuffles
// across the bodies of a group.  Total work is ~20 % higher than in k_integrate (pairs and the
// Jacobi recurrences are evaluated redundantly) and a lane has a single dependency chain (ILP 1),
// so this form pays only where one thread per system is issue-bound or spills (measured, see
// nbt_plan_lanes): Np >= 5 always, Np >= 3 for small batches and single systems.
//
// Same arithmetic per body as nbt_core.cuh (nbt_kepler_drift, nbt_elements, nbt_rsqrt); sums over
// bodies are taken in the same order, so results agree with k_integrate to round-off.
#pragma once
#include "nbt_core.cuh"

#if defined(__CUDACC__)

__device__ __forceinline__ double nbt_shfl(unsigned mask, double v, int src) { return __shfl_sync(mask, v, src); }

// accelerations of the Jacobi coordinate held by this lane (interaction Hamiltonian of nbt_accel)
template <int NP>
__device__ __forceinline__ void nbt_accel_lane(unsigned gmask, int base, int j, double x, double y, double z,
                                               double irj, const double (&mr)[NP], const double (&Gm)[NP + 1],
                                               double GMj, double& ax, double& ay, double& az) {
    if (NP == 1) { ax = ay = az = 0.0; return; }
    // every lane gathers the Jacobi positions of the whole group and rebuilds the heliocentric ones
    double qx[NP], qy[NP], qz[NP];
    {
        double Sx = 0.0, Sy = 0.0, Sz = 0.0;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            const double xi = nbt_shfl(gmask, x, base + i), yi = nbt_shfl(gmask, y, base + i), zi = nbt_shfl(gmask, z, base + i);
            qx[i] = xi + Sx; qy[i] = yi + Sy; qz[i] = zi + Sz;
            Sx = fma(mr[i], xi, Sx); Sy = fma(mr[i], yi, Sy); Sz = fma(mr[i], zi, Sz);
        }
    }
    double mqx = 0.0, mqy = 0.0, mqz = 0.0; // own heliocentric position
#pragma unroll
    for (int i = 0; i < NP; i++)
        if (i == j) { mqx = qx[i]; mqy = qy[i]; mqz = qz[i]; }
    // star <-> this planet (the star-planet-1 pair is the Kepler part and is skipped)
    double px = 0.0, py = 0.0, pz = 0.0;    // inertial acceleration of this planet
    double cx = 0.0, cy = 0.0, cz = 0.0;    // this planet's pull on the star
    {
        const double r2 = fma(mqx, mqx, fma(mqy, mqy, mqz * mqz));
        const double ir = nbt_rsqrt(r2);
        const double ir3 = ir * ir * ir;
        double Gmj = 0.0;
#pragma unroll
        for (int i = 0; i < NP; i++) if (i == j) Gmj = Gm[i + 1];
        const double fs = (j >= 1) ? Gmj * ir3 : 0.0, fp = (j >= 1) ? -Gm[0] * ir3 : 0.0;
        cx = fs * mqx; cy = fs * mqy; cz = fs * mqz;
        px = fp * mqx; py = fp * mqy; pz = fp * mqz;
    }
    // planet <-> planet pairs seen from this lane
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const double dx = qx[i] - mqx, dy = qy[i] - mqy, dz = qz[i] - mqz;
        const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
        const double ir = nbt_rsqrt((i == j) ? 1.0 : r2);
        const double f = (i == j) ? 0.0 : Gm[i + 1] * ir * ir * ir;
        px = fma(f, dx, px); py = fma(f, dy, py); pz = fma(f, dz, pz);
    }
    // inertial -> Jacobi: a'_i = p_i - A_{i-1},  A_i = A_{i-1} + (m_i/M_i) a'_i,  A_{-1} = a_star
    double Ax = 0.0, Ay = 0.0, Az = 0.0;
#pragma unroll
    for (int i = 1; i < NP; i++) { // a_star = sum of the planets' pulls (same order as nbt_accel)
        Ax += nbt_shfl(gmask, cx, base + i); Ay += nbt_shfl(gmask, cy, base + i); Az += nbt_shfl(gmask, cz, base + i);
    }
    double jx = 0.0, jy = 0.0, jz = 0.0;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const double pix = nbt_shfl(gmask, px, base + i), piy = nbt_shfl(gmask, py, base + i), piz = nbt_shfl(gmask, pz, base + i);
        const double tx = pix - Ax, ty = piy - Ay, tz = piz - Az;
        if (i == j) { jx = tx; jy = ty; jz = tz; }
        Ax = fma(mr[i], tx, Ax); Ay = fma(mr[i], ty, Ay); Az = fma(mr[i], tz, Az);
    }
    if (j >= 1) { // + G M_j r'_j / r'_j^3   (1/r' carried by the drift)
        const double f = GMj * irj * irj * irj;
        jx = fma(f, x, jx); jy = fma(f, y, jy); jz = fma(f, z, jz);
    }
    ax = jx; ay = jy; az = jz;
}

// acceleration of a (possibly force-gradient) kick, see nbt_accel_grad
template <int NP>
__device__ __forceinline__ void nbt_accel_lane_grad(unsigned gmask, int base, int j, double x, double y, double z,
                                                    double irj, const double (&mr)[NP], const double (&Gm)[NP + 1],
                                                    double GMj, double gd, double& ax, double& ay, double& az) {
    nbt_accel_lane<NP>(gmask, base, j, x, y, z, irj, mr, Gm, GMj, ax, ay, az);
    if (gd != 0.0) {
        const double tx = fma(gd, ax, x), ty = fma(gd, ay, y), tz = fma(gd, az, z);
        double tr, tir;
        nbt_radius(tx, ty, tz, tr, tir);
        nbt_accel_lane<NP>(gmask, base, j, tx, ty, tz, tir, mr, Gm, GMj, ax, ay, az);
    }
}

// One lane of one system: integrate + sample + detect transits (the lanes form of nbt_simulate_system).
template <int NP, bool SERIES>
__device__ void nbt_simulate_lane(const nbt_run_args& A, const nbt_scheme& sc, int sys, int j, int base, unsigned gmask) {
    const double* par = A.params + (size_t)sys * (NP + 1) * NBT_NPAR;
    double mr[NP], Gm[NP + 1];
    double GMj = 0.0;
    double x, y, z, vx, vy, vz;
    {   // nbt_setup for this lane's planet (nbody/simulation.py:45-47)
        double M = par[NBT_P_M];
        Gm[0] = NBT_G * M;
        double mu_j = 0.0;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            const double m = par[(i + 1) * NBT_NPAR + NBT_P_M];
            M += m;
            mr[i] = m / M; Gm[i + 1] = NBT_G * m;
            if (i == j) mu_j = NBT_G * M;
        }
        GMj = mu_j;
        const double* p = par + (j + 1) * NBT_NPAR;
        const double P = p[NBT_P_P], e = p[NBT_P_E];
        const double a = cbrt(P * P * mu_j / (4.0 * NBT_PI * NBT_PI));
        const double cf = cos(p[NBT_P_F]), sf = sin(p[NBT_P_F]);
        const double r = a * (1.0 - e * e) / (1.0 + e * cf);
        const double v0 = sqrt(mu_j / (a * (1.0 - e * e)));
        const double cO = cos(p[NBT_P_OMEGA_NODE]), sO = sin(p[NBT_P_OMEGA_NODE]);
        const double co = cos(p[NBT_P_OMEGA_PERI]), so = sin(p[NBT_P_OMEGA_PERI]);
        const double ci = cos(p[NBT_P_INC]), si = sin(p[NBT_P_INC]);
        const double cwf = co * cf - so * sf, swf = so * cf + co * sf;
        x = r * (cO * cwf - sO * swf * ci);
        y = r * (sO * cwf + cO * swf * ci);
        z = r * (swf * si);
        const double vp = -(swf + e * so), vq = (cwf + e * co);
        vx = v0 * (cO * vp - sO * vq * ci);
        vy = v0 * (sO * vp + cO * vq * ci);
        vz = v0 * (vq * si);
    }
    int status = (j == 0 && nbt_uncalibrated_impl(par, NP + 1)) ? NBT_S_UNCALIBRATED : 0;
    const int B = A.B;
    const int NO = (A.n_outputs_sys != nullptr) ? A.n_outputs_sys[sys] : A.n_outputs;
    const double n_days = (A.n_days_sys != nullptr) ? A.n_days_sys[sys] : A.n_days;
    const int kraw = (A.substeps != nullptr) ? A.substeps[sys] : A.substeps_all;
    const int k = kraw < 0 ? -kraw : kraw;
    const double step = n_days / (double)(NO - 1);
    const double dt = step / (double)k;
    const bool want_means = (A.flags & NBT_F_MEANS) != 0;
    const bool want_omega = (A.flags & NBT_F_MEAN_OMEGA) != 0;
    const bool want_rv = (A.flags & NBT_F_RV) != 0 && A.rv != nullptr;
    const bool want_orbit = (A.flags & NBT_F_ORBIT) != 0 && A.orbit_xz != nullptr;
    const bool want_series = SERIES;
    const bool tt_on = !((A.flags & NBT_F_PLANET1_ONLY) && j > 0);
    const double rv_scale = 1.496e11 * ((double)NO / (n_days * 24.0 * 60.0 * 60.0));
    const int W = 3 + 10 * NP;

    double rj, irj;
    nbt_radius(x, y, z, rj, irj);
    double ax, ay, az;
    nbt_accel_lane_grad<NP>(gmask, base, j, x, y, z, irj, mr, Gm, GMj, sc.g[0] * dt * dt, ax, ay, az);

    double dx_prev = 0.0, dvx_prev = 0.0, xs_prev = 0.0, t_prev = 0.0;
    double sum_a = 0.0, sum_e = 0.0, sum_i = 0.0, sum_o = 0.0;
    int cnt = 0;
    const int64_t off = A.tt_offsets[(size_t)sys * NP + j];
    const int64_t cap = A.tt_offsets[(size_t)sys * NP + j + 1] - off;

    for (int o = 0; o < NO; o++) {
        if (o > 0) {
            for (int sub = 0; sub < k; sub++) {
                {
                    const double kb = sc.b[0] * dt;
                    vx = fma(kb, ax, vx); vy = fma(kb, ay, vy); vz = fma(kb, az, vz);
                }
                for (int st = 0; st < sc.S; st++) {
                    nbt_kepler_drift<(NP <= 2)>(GMj, sc.a[st] * dt, x, y, z, vx, vy, vz, rj, irj, status);
                    nbt_accel_lane_grad<NP>(gmask, base, j, x, y, z, irj, mr, Gm, GMj, sc.g[st + 1] * dt * dt, ax, ay, az);
                    const double kb = sc.b[st + 1] * dt;
                    vx = fma(kb, ax, vx); vy = fma(kb, ay, vy); vz = fma(kb, az, vz);
                }
            }
        }
        const double t = (o == NO - 1) ? n_days : (double)o * step;
        if (o > 0) nbt_radius(x, y, z, rj, irj);   // refresh the carried radius from the coordinates
        // heliocentric x, vx of this planet and the star's barycentric position (group prefix sums)
        double Sx = 0.0, Svx = 0.0, Sy = 0.0, Sz = 0.0, dx = 0.0, dvx = 0.0, hy = 0.0, hz = 0.0;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            const double xi = nbt_shfl(gmask, x, base + i), vxi = nbt_shfl(gmask, vx, base + i);
            if (i == j) { dx = xi + Sx; dvx = vxi + Svx; }
            Sx = fma(mr[i], xi, Sx); Svx = fma(mr[i], vxi, Svx);
            if (want_orbit || want_series) {
                const double yi = nbt_shfl(gmask, y, base + i), zi = nbt_shfl(gmask, z, base + i);
                if (i == j) { hy = yi + Sy; hz = zi + Sz; }
                Sy = fma(mr[i], yi, Sy); Sz = fma(mr[i], zi, Sz);
            }
        }
        const double xs = -Sx;
        const bool bad = !isfinite(Sx + Svx) || (status & (NBT_S_KEPLER | NBT_S_NONFINITE));
        if (__any_sync(gmask, bad)) {                                      // whole group stops together
            status |= NBT_S_NONFINITE;
            if (j == 0) nbt_fill_broken(want_rv ? A.rv : nullptr, want_orbit ? A.orbit_xz : nullptr, want_series ? A.series : nullptr, B, NP, A.orbit_stride, sys, o, NO);
            break;
        }
        if (o > 0) {
            if (tt_on && dx_prev >= 0.0 && dx <= 0.0) {                    // simulation.py:168
                const double t0 = nbt_crossing_time(A.tt_mode, t_prev, t, dx_prev, dx, dvx_prev, dvx);  // tools.py:61-65
                if (cnt < cap) A.tt[off + cnt] = t0; else status |= NBT_S_TT_OVERFLOW;
                cnt++;
            }
            if (want_rv && j == 0) A.rv[(size_t)(o - 1) * B + sys] = (xs - xs_prev) * rv_scale;
        }
        dx_prev = dx; dvx_prev = dvx; xs_prev = xs; t_prev = t;

        if (want_series) {
            const nbt_orbit ob = nbt_elements<2>(GMj, x, y, z, vx, vy, vz, rj, irj);
            {
                double* row = A.series + ((size_t)o * B + sys) * W;
                if (j == 0) { row[0] = xs; row[1] = -Sy; row[2] = -Sz; }
                double* p = row + 3 + 10 * j;
                p[0] = dx + xs; p[1] = hy - Sy; p[2] = hz - Sz;
                p[3] = ob.P; p[4] = ob.a; p[5] = ob.e; p[6] = ob.inc; p[7] = ob.Omega; p[8] = ob.omega; p[9] = ob.M;
            }
            sum_a += ob.a; sum_e += ob.e; sum_i += ob.inc; sum_o += ob.omega;
            if (!(ob.a > 0.0)) status |= NBT_S_UNBOUND;
        } else if (want_means) {
            if (want_omega) {
                const nbt_orbit ob = nbt_elements<1>(GMj, x, y, z, vx, vy, vz, rj, irj);
                sum_a += ob.a; sum_e += ob.e; sum_i += ob.inc; sum_o += ob.omega;
                if (!(ob.a > 0.0)) status |= NBT_S_UNBOUND;
            } else {
                const nbt_orbit ob = nbt_elements<0>(GMj, x, y, z, vx, vy, vz, rj, irj);
                sum_a += ob.a; sum_e += ob.e; sum_i += ob.inc;
                if (!(ob.a > 0.0)) status |= NBT_S_UNBOUND;
            }
        }
        if (want_orbit && (o % A.orbit_stride) == 0) {            // simulation.py:227-228
            double* dst = A.orbit_xz + (((size_t)(o / A.orbit_stride) * B + sys) * NP + j) * 2;
            dst[0] = dx + xs; dst[1] = hz - Sz;
        }
    }
    if (!isfinite(x + vx + y + vy + z + vz)) status |= NBT_S_NONFINITE;
    // group-wide status
    int gs = status;
#pragma unroll
    for (int i = 0; i < NP; i++) gs |= __shfl_sync(gmask, status, base + i);
    {
        const size_t sp = (size_t)sys * NP + j;
        A.n_tt[sp] = cnt;
        if (want_means || want_series) {
            const double inv = 1.0 / (double)NO;
            if (A.mean_a) A.mean_a[sp] = sum_a * inv;
            if (A.mean_e) A.mean_e[sp] = sum_e * inv;
            if (A.mean_inc) A.mean_inc[sp] = sum_i * inv;
            if (A.mean_omega && (want_omega || want_series)) A.mean_omega[sp] = sum_o * inv;
        }
        if (j == 0) A.status[sys] = gs;
    }
}

#endif // __CUDACC__
