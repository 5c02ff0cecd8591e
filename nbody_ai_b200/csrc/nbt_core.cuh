// nbt_core.cuh — per-system FP64 math of the batched TTV engine (sm_100a device code).
//
// One system = star + NP planets in JACOBI coordinates, all state in registers of one thread.
// The integrator is a symmetric BAB composition of the Wisdom-Holman map
//     K(b0 dt) D(a1 dt) K(b1 dt) ... D(aS dt) K(bS dt)
// with D = exact Kepler drift of every Jacobi coordinate (universal variables, Stumpff series)
// and K = kick from the interaction Hamiltonian (direct pairs minus the star-planet-1 pair,
// plus the Jacobi indirect term).  Replaces what the reference gets from REBOUND through
// sim.integrate(time) in nbody/simulation.py:141-150.
//
// Every function is __host__ __device__ so that tests/hostsim can run the *same* arithmetic on
// the CPU for numerics checks in the GPU-less container.  That host build is test
// infrastructure only; the product library (nbt_capi.cu) has no CPU path.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/nbt.h"

#if defined(__CUDACC__)
#define NBT_HD __host__ __device__ __forceinline__
#define NBT_HD_COLD __host__ __device__ __noinline__
#else
#define NBT_HD inline
#define NBT_HD_COLD inline
#endif

#define NBT_PI 3.14159265358979323846264338327950288
#define NBT_MAX_STAGES 8
#ifndef NBT_DRIFT_GROUP
#define NBT_DRIFT_GROUP(NP) ((NP) == 3 ? 3 : 2)   /* planets whose branch-free Kepler solves share a basic block */
#endif
#ifndef NBT_FAST_ELEMENTS
#define NBT_FAST_ELEMENTS(NP, LAT) ((NP) >= 2)   /* which kernels sample the elements through nbt_elements_fast */
#endif
#define NBT_C4_GRAD (1.0 / 24.0)  /* displacement coefficient of the gradient kick of NBT_SCHEME_C4 */

// ---------------------------------------------------------------------------------------------
// reciprocal / reciprocal square root: MUFU seed (rsqrt.approx.f64 / rcp.approx.f64, ~2^-20) +
// one cubic (Halley) correction in DFMA -> ~2^-60, i.e. correctly rounded to within 1 ulp.
// This is what keeps FP64 div/sqrt (10+ instructions each in libdevice) off the hot loop.
// ---------------------------------------------------------------------------------------------
NBT_HD double nbt_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);           // 1 - x y^2
    double p = fma(0.375, e, 0.5);            // 1/2 + 3/8 e
    return fma(y * e, p, y);                  // y (1 + e/2 + 3/8 e^2)
#else
    return 1.0 / sqrt(x);
#endif
}

NBT_HD double nbt_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);               // 1 - x y
    double p = fma(e, e, e);                  // e + e^2
    return fma(y, p, y);                      // y (1 + e + e^2)
#else
    return 1.0 / x;
#endif
}

// refine a reciprocal from a nearby guess y0 ~ 1/x (relative error <~ 1e-5): one cubic step
NBT_HD double nbt_rcp_from(double x, double y0) {
    const double e = fma(-x, y0, 1.0);
    return fma(y0, fma(e, e, e), y0);
}

// x^(-3/2) straight from the MUFU seed: y^3 (1 + 3/2 e + 15/8 e^2), e = 1 - x y^2 — one instruction less than
// nbt_rsqrt followed by two multiplications, and it is what every pair force needs
NBT_HD double nbt_rsqrt3(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double y2 = y * y;
    const double e = fma(-x, y2, 1.0);
    const double y3 = y2 * y;
    return fma(y3 * e, fma(1.875, e, 1.5), y3);
#else
    const double y = 1.0 / sqrt(x);
    return y * y * y;
#endif
}

// ---------------------------------------------------------------------------------------------
// Literal constants of the hot loop.  In device code they are read from the constant bank: ptxas folds a
// __constant__ double into the DFMA / DMUL as a c[bank][offset] operand, whereas a 64-bit literal is rebuilt
// in a register pair by two IMAD.MOV per use (ncu, round 1: 10 % of the warp instructions of k_integrate<2>
// were such moves) and costs a third register operand (DFMA with three distinct register operands issues
// every 2.68 instead of 2 cycles, profiles/README.md).
// ---------------------------------------------------------------------------------------------
#define NBT_KLIST(X)                                                                                  \
    X(c2_1, -1.0 / 24.0) X(c2_2, 1.0 / 720.0) X(c2_3, -1.0 / 40320.0) X(c2_4, 1.0 / 3628800.0)        \
    X(c2_5, -1.0 / 479001600.0) X(c2_6, 1.0 / 87178291200.0)                                          \
    X(c3_0, 1.0 / 6.0) X(c3_1, -1.0 / 120.0) X(c3_2, 1.0 / 5040.0) X(c3_3, -1.0 / 362880.0)           \
    X(c3_4, 1.0 / 39916800.0) X(c3_5, -1.0 / 6227020800.0) X(c3_6, 1.0 / 1307674368000.0)             \
    X(m6, -1.0 / 6.0) X(i24, 1.0 / 24.0) X(i3, 1.0 / 3.0)
#if defined(__CUDACC__)
#define NBT_KDEV(name, val) __constant__ double nbt_kd_##name = (val);
NBT_KLIST(NBT_KDEV)
#undef NBT_KDEV
#endif
#define NBT_KHOST(name, val) static const double nbt_kh_##name = (val);
NBT_KLIST(NBT_KHOST)
#undef NBT_KHOST
#if defined(__CUDA_ARCH__)
#define NBT_K(name) nbt_kd_##name
#else
#define NBT_K(name) nbt_kh_##name
#endif

// ---------------------------------------------------------------------------------------------
// Kepler drift of one Jacobi coordinate by h (either sign) under mu = G M_i.
//   universal Kepler equation  F(X) = r0 X + eta G2 + zeta G3 - h = 0,   G_k = X^k c_k(beta X^2)
//   beta = 2 mu / r0 - v0^2,  eta = r0 . v0,  zeta = mu - beta r0
// Solve: X0 = cubic series reversion at X = 0 (no Stumpff needed), then ONE quartically
// convergent step (same reversion, now with the full F, F', F'', F''' at X0) and a 3rd-order
// Taylor shift of G1..G3 to the corrected X.  If the correction is not small the (warp-
// divergent, rare) safe loop below iterates to convergence and flags NBT_S_KEPLER on failure.
// ---------------------------------------------------------------------------------------------
template <bool ESTRIN = true>
NBT_HD void nbt_stumpff(double z, double& c2, double& c3) {
    // c2 = 1/2! - z/4! + z^2/6! - ...,  c3 = 1/3! - z/5! + z^2/7! - ...   (|z| <~ 0.5)
    if (ESTRIN) {
        // Estrin's scheme: dependency depth 3 instead of Horner's 6 (the latency variant: few warps resident)
        const double z2 = z * z, z4 = z2 * z2;
        {
            const double p01 = fma(z, NBT_K(c2_1), 0.5), p23 = fma(z, NBT_K(c2_3), NBT_K(c2_2)),
                         p45 = fma(z, NBT_K(c2_5), NBT_K(c2_4));
            c2 = fma(z4, fma(z2, NBT_K(c2_6), p45), fma(z2, p23, p01));
        }
        {
            const double p01 = fma(z, NBT_K(c3_1), NBT_K(c3_0)), p23 = fma(z, NBT_K(c3_3), NBT_K(c3_2)),
                         p45 = fma(z, NBT_K(c3_5), NBT_K(c3_4));
            c3 = fma(z4, fma(z2, NBT_K(c3_6), p45), fma(z2, p23, p01));
        }
    } else {
        // Horner: the same twelve FP64 instructions, but one constant-bank operand per FMA and no coefficient
        // held in a register (Estrin's p23 / p45 need one each): the register-starved kernels (Np >= 3, the
        // throughput variant) interleave enough chains of their own
        c2 = fma(fma(fma(fma(fma(fma(z, NBT_K(c2_6), NBT_K(c2_5)), z, NBT_K(c2_4)), z, NBT_K(c2_3)), z, NBT_K(c2_2)), z, NBT_K(c2_1)), z, 0.5);
        c3 = fma(fma(fma(fma(fma(fma(z, NBT_K(c3_6), NBT_K(c3_5)), z, NBT_K(c3_4)), z, NBT_K(c3_3)), z, NBT_K(c3_2)), z, NBT_K(c3_1)), z, NBT_K(c3_0));
    }
}

// general-|z| Stumpff via argument quartering (only used by the safe path)
NBT_HD void nbt_stumpff_big(double z, double& c0, double& c1, double& c2, double& c3) {
    int n = 0;
    while (fabs(z) > 0.1 && n < 40) { z *= 0.25; n++; }
    nbt_stumpff(z, c2, c3);
    c1 = 1.0 - z * c3;
    c0 = 1.0 - z * c2;
    for (; n > 0; n--) {
        c3 = 0.25 * (c2 + c0 * c3);
        c2 = 0.5 * c1 * c1;
        c1 = c0 * c1;
        c0 = 2.0 * c0 * c0 - 1.0;
    }
}

NBT_HD void nbt_kepler_safe(double mu, double h, double r0, double eta, double beta, double zeta,
                            double& X, double& G1, double& G2, double& G3, double& r, int& status) {
    // Newton on the monotone F (F' = r > 0), safeguarded by a bracket that every iterate tightens;
    // a step that leaves the bracket is replaced by its midpoint.  Converged when the step drops
    // to round-off (a few ulp of X) or F vanishes.
    (void)mu;
    double Xc = X;
    double lo = -INFINITY, hi = INFINITY;
    bool ok = false;
    if (!(r0 > 0.0) || !isfinite(r0 + eta + beta + X)) { // broken state (collision / overflow): do not iterate
        status |= NBT_S_KEPLER | NBT_S_NONFINITE;
        G1 = G2 = G3 = r = NAN;
        return;
    }
    for (int it = 0; it < 80; it++) {
        double c0, c1, c2, c3;
        const double X2 = Xc * Xc;
        nbt_stumpff_big(beta * X2, c0, c1, c2, c3);
        G1 = Xc * c1; G2 = X2 * c2; G3 = X2 * Xc * c3;
        const double F = fma(r0, Xc, fma(eta, G2, zeta * G3)) - h;
        r = fma(eta, G1, fma(zeta, G2, r0));
        if (F > 0.0) hi = Xc; else lo = Xc;
        if (F == 0.0) { ok = true; break; }
        double Xn = Xc - F / r;
        if (!(r > 0.0) || !(Xn > lo && Xn < hi)) {
            if (isfinite(lo) && isfinite(hi)) Xn = 0.5 * (lo + hi);
            else Xn = (F > 0.0) ? Xc - fmax(fabs(Xc), fabs(h) / r0) : Xc + fmax(fabs(Xc), fabs(h) / r0);
        }
        const double d = Xn - Xc;
        Xc = Xn;
        if (fabs(d) <= 8.0 * 2.220446049250313e-16 * fabs(Xc)) { ok = true; break; }
    }
    double c0, c1, c2, c3;
    const double X2 = Xc * Xc;
    nbt_stumpff_big(beta * X2, c0, c1, c2, c3);
    G1 = Xc * c1; G2 = X2 * c2; G3 = X2 * Xc * c3;
    r = fma(eta, G1, fma(zeta, G2, r0));
    X = Xc;
    if (!ok || !(r > 0.0)) status |= NBT_S_KEPLER;
}

// r0 / ir0 = |r| and 1/|r| of the incoming position, carried from the end of the previous drift
// (a kick does not move positions) and refreshed from x, y, z at every output sample; on return they
// hold the new radius and its reciprocal.  This removes a rsqrt chain per planet per stage here and
// another in the indirect term of the kick.
// Middle path between the one-shot solve of nbt_kepler_drift and the bracketed fallback: a few more
// quartic steps with the polynomial Stumpff functions (valid while |beta X^2| < 0.5).  Eccentric,
// strongly perturbed orbits (the Hill-unstable systems of a random batch) miss the one-shot tolerance
// near pericentre; they converge here in 1-2 rounds at ~2x the cost of a normal drift instead of the
// ~10x of the bracketed Newton with argument-quartered Stumpff functions.
NBT_HD void nbt_kepler_mid(double mu, double h, double r0, double eta, double beta, double zeta,
                           double& X, double& G1, double& G2, double& G3, double& r, int& status) {
    const double be = beta * eta;
    bool ok = false;
    for (int it = 0; it < 5; it++) {
        const double X2 = X * X, zz = beta * X2;
        if (!(zz < 0.5 && zz > -0.5)) break;
        double c2, c3;
        nbt_stumpff(zz, c2, c3);
        const double G0 = fma(-zz, c2, 1.0);
        G1 = X * fma(-zz, c3, 1.0); G2 = X2 * c2; G3 = X2 * X * c3;
        const double F = fma(zeta, G3, fma(eta, G2, fma(r0, X, -h)));
        r = fma(eta, G1, fma(zeta, G2, r0));
        if (!(r > 0.0)) break;
        const double Fpp = fma(eta, G0, zeta * G1), Fppp = fma(zeta, G0, -be * G1);
        const double ir = 1.0 / r;
        const double u = -F * ir, a2 = 0.5 * ir * Fpp, a3 = (1.0 / 6.0) * ir * Fppp;
        const double d = u * fma(u, fma(u, fma(2.0 * a2, a2, -a3), -a2), 1.0);
        if (fabs(d) <= 1e-13 * fabs(X)) { ok = true; break; }   // G's are those of the converged X
        X += d;
    }
    if (!ok) nbt_kepler_safe(mu, h, r0, eta, beta, zeta, X, G1, G2, G3, r, status);
}

// The one-shot solve in two phases, no branch inside either.
//   nbt_kepler_solve: Kepler's equation — first guess, Stumpff functions, one quartic correction d, Taylor shift of
//                     G1..G3 and r to X + d, Gauss f and g (in `m`); returns whether the convergence test held.
//   nbt_kepler_apply: the new position / velocity / radius, written in place (12 FMAs).
// Between the two the caller redoes the rare failures out of line (middle path / fallback, nbt_kepler_finish)
// and makes their `m` the identity update (nbt_kepler_identity), so that nothing is selected or copied on the
// hot path: the drifts of a group of planets are two basic blocks whose dependency chains interleave.
// Written for few FP64 issue slots — a lone warp already fills the FP64 pipe of its SM sub-partition —:
// scalings by literals ride on FMAs, the 1/r0-scaled series coefficients are formed once, literals come from the
// constant bank.
struct nbt_kepler_mid_t { double fm1, g, fd, gdm1, r, ir; };   // Gauss f - 1, g, fdot, gdot - 1; new radius and its reciprocal

template <bool ESTRIN = true>
NBT_HD bool nbt_kepler_solve(double mu, double h, double x, double y, double z, double vx, double vy, double vz,
                             double r0, double ir0, nbt_kepler_mid_t& m) {
    const double s = h * ir0, s2 = s * s;
    const double mi0 = -mu * ir0;
    const double v02 = fma(vx, vx, fma(vy, vy, vz * vz));
    const double eta = fma(x, vx, fma(y, vy, z * vz));
    const double beta = fma(-2.0, mi0, -v02);
    const double zeta = fma(-beta, r0, mu);
    const double be = beta * eta;
    // X0: reversion of h = r0 X + eta X^2/2 + zeta X^3/6 - beta eta X^4/24 ...
    //   X = s - A s^2 + (2A^2 - B) s^3 + (-5A^3 + 5AB - C) s^4,  A = eta/2r0, B = zeta/6r0, C = -beta eta/24r0
    const double e1 = eta * ir0, z1 = zeta * ir0, b1 = beta * e1;
    const double e2 = e1 * e1, z6 = NBT_K(m6) * z1, he = 0.5 * e1;
    const double k3 = fma(0.5, e2, z6);
    const double k4 = fma(he, fma(-5.0, z6, -1.25 * e2), NBT_K(i24) * b1);
    const double X = fma(s2, fma(s, fma(s, k4, k3), -he), s);

    const double X2 = X * X, X3 = X2 * X;
    const double zz = beta * X2;
    const double base = fma(r0, X, -h);
    double c2, c3;
    nbt_stumpff<ESTRIN>(zz, c2, c3);
    const double G0 = fma(-zz, c2, 1.0);
    const double G1 = X * fma(-zz, c3, 1.0), G2 = X2 * c2, G3 = X3 * c3;
    const double F = fma(zeta, G3, fma(eta, G2, base));
    const double r1 = fma(eta, G1, fma(zeta, G2, r0));          // dF/dX
    const double Fpp = fma(eta, G0, zeta * G1);                 // d2F/dX2
    const double Fppp = fma(zeta, G0, -be * G1);                // d3F/dX3
    const double ir1 = nbt_rcp(r1);
    // quartic correction d = u (1 - a2 u + (2 a2^2 - a3) u^2), u = -F/r1, a2 = Fpp/(2 r1), a3 = Fppp/(6 r1)
    const double u = -F * ir1;
    const double w = ir1 * Fpp, v6 = ir1 * Fppp, hw = 0.5 * w;
    const double d = u * fma(u, fma(u, fma(hw, w, NBT_K(m6) * v6), -hw), 1.0);
    // shift G1..G3 and r from X to X + d (Horner in d):  dG_k/dX = G_{k-1},  dG0/dX = -beta G1,
    // dr/dX = Fpp, d2r/dX2 = Fppp, d3r/dX3 = -beta Fpp
    const double hd = 0.5 * d, td = NBT_K(i3) * d;
    const double bG1 = beta * G1, bG0 = beta * G0, bF = beta * Fpp;
    const double nG3 = fma(d, fma(hd, fma(td, G0, G1), G2), G3);
    const double nG2 = fma(d, fma(hd, fma(-td, bG1, G0), G1), G2);
    const double nG1 = fma(d, fma(-hd, fma(td, bG0, bG1), G0), G1);
    const double r = fma(d, fma(hd, fma(-td, bF, Fppp), Fpp), r1);
    const double ir = nbt_rcp_from(r, ir1);
    // Gauss f, g in the round-off-friendly "minus one" form
    m.fm1 = mi0 * nG2;
    m.g = fma(-mu, nG3, h);
    m.fd = (mi0 * nG1) * ir;
    m.gdm1 = (-mu * nG2) * ir;
    m.r = r; m.ir = ir;
    return fabs(d) <= 2e-4 * fabs(X) && zz < 0.5 && zz > -0.5;
}

NBT_HD void nbt_kepler_apply(const nbt_kepler_mid_t& m, double& x, double& y, double& z, double& vx, double& vy,
                             double& vz, double& r_out, double& ir_out) {
    const double nx = fma(m.fm1, x, fma(m.g, vx, x)), ny = fma(m.fm1, y, fma(m.g, vy, y)), nz = fma(m.fm1, z, fma(m.g, vz, z));
    vx = fma(m.fd, x, fma(m.gdm1, vx, vx));
    vy = fma(m.fd, y, fma(m.gdm1, vy, vy));
    vz = fma(m.fd, z, fma(m.gdm1, vz, vz));
    x = nx; y = ny; z = nz;
    r_out = m.r; ir_out = m.ir;
}

// `m` of a planet whose drift is already done (by nbt_kepler_finish): nbt_kepler_apply leaves its state alone
NBT_HD void nbt_kepler_identity(double r, double ir, nbt_kepler_mid_t& m) {
    m.fm1 = 0.0; m.g = 0.0; m.fd = 0.0; m.gdm1 = 0.0; m.r = r; m.ir = ir;
}

// slow continuation of a drift whose one-shot solve missed its tolerance (out of line: never taken in normal
// steps): the middle path, then the bracketed Newton.  Everything goes in and out BY VALUE: a reference to the
// caller's state would pin the whole state array to local memory on the hot path (ptxas stores it before every
// potential call).
struct nbt_body { double x, y, z, vx, vy, vz, r, ir; int status; };

NBT_HD_COLD nbt_body nbt_kepler_finish(double mu, double h, double x, double y, double z, double vx, double vy,
                                       double vz, double r0, double ir0) {
    nbt_body o;
    o.status = 0;
    const double mi0 = -mu * ir0;
    const double v02 = fma(vx, vx, fma(vy, vy, vz * vz));
    const double eta = fma(x, vx, fma(y, vy, z * vz));
    const double beta = fma(-2.0, mi0, -v02);
    const double zeta = fma(-beta, r0, mu);
    // first guess: the cubic reversion of the one-shot solve
    const double s = h * ir0;
    const double A = 0.5 * eta * ir0, Bc = zeta * ir0 / 6.0;
    double X = s * (1.0 + s * (-A + s * (2.0 * A * A - Bc)));
    double G1 = 0.0, G2 = 0.0, G3 = 0.0, r = r0;
    nbt_kepler_mid(mu, h, r0, eta, beta, zeta, X, G1, G2, G3, r, o.status);
    const double ir = 1.0 / r;
    const double fm1 = mi0 * G2;
    const double g = fma(-mu, G3, h);
    const double fd = (mi0 * G1) * ir;
    const double gdm1 = (-mu * G2) * ir;
    o.x = fma(fm1, x, fma(g, vx, x)); o.y = fma(fm1, y, fma(g, vy, y)); o.z = fma(fm1, z, fma(g, vz, z));
    o.vx = fma(fd, x, fma(gdm1, vx, vx)); o.vy = fma(fd, y, fma(gdm1, vy, vy)); o.vz = fma(fd, z, fma(gdm1, vz, vz));
    o.r = r; o.ir = ir;
    return o;
}

// r0 / ir0 = |r| and 1/|r| of the incoming position, carried from the end of the previous drift (a kick
// does not move positions) and refreshed from x, y, z at every output sample; on return they hold the new
// radius and its reciprocal.  Branchy form (the throughput variant at Np <= 2, the lanes kernel): one-shot
// solve, else middle path + fallback out of line.
template <bool ESTRIN>   // the polynomial scheme of the caller's kernel family (Np <= 2: Estrin), so that the branchy and the
                         // branch-free form of one Np give bit-identical results
NBT_HD void nbt_kepler_drift(double mu, double h, double& x, double& y, double& z, double& vx,
                             double& vy, double& vz, double& r0io, double& ir0io, int& status) {
    nbt_kepler_mid_t m;
    if (nbt_kepler_solve<ESTRIN>(mu, h, x, y, z, vx, vy, vz, r0io, ir0io, m)) {
        nbt_kepler_apply(m, x, y, z, vx, vy, vz, r0io, ir0io);
    } else {
        const nbt_body o = nbt_kepler_finish(mu, h, x, y, z, vx, vy, vz, r0io, ir0io);
        x = o.x; y = o.y; z = o.z; vx = o.vx; vy = o.vy; vz = o.vz; r0io = o.r; ir0io = o.ir; status |= o.status;
    }
}

// |r| and 1/|r| from the coordinates (start of a run and at every output sample)
NBT_HD void nbt_radius(double x, double y, double z, double& r, double& ir) {
    const double r2 = fma(x, x, fma(y, y, z * z));
    ir = nbt_rsqrt(r2);
    r = r2 * ir;
}

// ---------------------------------------------------------------------------------------------
// System state (structure of small fixed arrays -> registers after full unrolling)
// ---------------------------------------------------------------------------------------------
template <int NP>
struct nbt_state {
    double x[NP], y[NP], z[NP], vx[NP], vy[NP], vz[NP]; // Jacobi coordinates of planets 1..NP
    double r[NP], ir[NP];                               // carried |r'| and 1/|r'| (see nbt_kepler_drift)
};

template <int NP>
struct nbt_consts {
    double GM[NP];     // G * M_i, M_i = m_0 + ... + m_i          (Kepler mu of Jacobi coordinate i)
    double mr[NP];     // m_i / M_i                               (Jacobi <-> heliocentric shifts)
    double Gm[NP + 1]; // G * m_k for k = 0..NP                   (pair forces)
    double k02, k12;   // NP == 2: G m_0 M_2 / M_1, G m_1 M_2 / M_1 (closed form of nbt_accel)
};

// Jacobi accelerations from the interaction Hamiltonian
//   H_int = sum_{i>=2} G m_i M_{i-1} / r'_i  -  sum_{i>=2} G m_0 m_i / |q_i|  -  sum_{1<=i<j} G m_i m_j / |q_i - q_j|
// (q = heliocentric position; the star-planet-1 pair is the i=1 Kepler term and cancels).
template <int NP>
NBT_HD void nbt_accel(const nbt_state<NP>& s, const nbt_consts<NP>& c, double (&ax)[NP],
                      double (&ay)[NP], double (&az)[NP]) {
    if (NP == 1) { ax[0] = ay[0] = az[0] = 0.0; return; }
    if (NP == 2) {
        // two planets, everything substituted (21 DMUL + 22 DFMA + 3 DADD instead of 31 + 19 + 12 in the general form):
        //   q2 = r'_2 + (m1/M1) r'_1,  d = q2 - r'_1
        //   a'_1 = G m2 ( d/|d|^3 - q2/|q2|^3 )
        //   a'_2 = G M2 r'_2/|r'_2|^3 - (G m0 M2/M1) q2/|q2|^3 - (G m1 M2/M1) d/|d|^3
        const double qx = fma(c.mr[0], s.x[0], s.x[1]), qy = fma(c.mr[0], s.y[0], s.y[1]), qz = fma(c.mr[0], s.z[0], s.z[1]);
        const double dx = qx - s.x[0], dy = qy - s.y[0], dz = qz - s.z[0];
        const double w02 = nbt_rsqrt3(fma(qx, qx, fma(qy, qy, qz * qz)));
        const double w12 = nbt_rsqrt3(fma(dx, dx, fma(dy, dy, dz * dz)));
        const double ir = s.ir[1];
        const double c2 = c.GM[1] * ir * (ir * ir);
        const double t12 = c.Gm[2] * w12, t02 = c.Gm[2] * w02;
        const double u02 = c.k02 * w02, u12 = c.k12 * w12;
        ax[0] = fma(t12, dx, -t02 * qx); ay[0] = fma(t12, dy, -t02 * qy); az[0] = fma(t12, dz, -t02 * qz);
        ax[1] = fma(-u12, dx, fma(-u02, qx, c2 * s.x[1]));
        ay[1] = fma(-u12, dy, fma(-u02, qy, c2 * s.y[1]));
        az[1] = fma(-u12, dz, fma(-u02, qz, c2 * s.z[1]));
        return;
    }
    double qx[NP], qy[NP], qz[NP];
    {   // heliocentric positions: q_i = r'_i + S_{i-1},  S_i = S_{i-1} + (m_i/M_i) r'_i
        qx[0] = s.x[0]; qy[0] = s.y[0]; qz[0] = s.z[0];
        double Sx = c.mr[0] * s.x[0], Sy = c.mr[0] * s.y[0], Sz = c.mr[0] * s.z[0];
#pragma unroll
        for (int i = 1; i < NP; i++) {
            qx[i] = s.x[i] + Sx; qy[i] = s.y[i] + Sy; qz[i] = s.z[i] + Sz;
            if (i + 1 < NP) { Sx = fma(c.mr[i], s.x[i], Sx); Sy = fma(c.mr[i], s.y[i], Sy); Sz = fma(c.mr[i], s.z[i], Sz); }
        }
    }
    double a0x = 0.0, a0y = 0.0, a0z = 0.0; // star
    double px[NP], py[NP], pz[NP];          // planets (inertial accelerations, included pairs only)
    px[0] = py[0] = pz[0] = 0.0;
#pragma unroll
    for (int i = 1; i < NP; i++) { // star - planet i (i >= 2 in 1-based numbering)
        const double w = nbt_rsqrt3(fma(qx[i], qx[i], fma(qy[i], qy[i], qz[i] * qz[i])));
        const double fs = c.Gm[i + 1] * w, fp = -c.Gm[0] * w;
        a0x = fma(fs, qx[i], a0x); a0y = fma(fs, qy[i], a0y); a0z = fma(fs, qz[i], a0z);
        px[i] = fp * qx[i]; py[i] = fp * qy[i]; pz[i] = fp * qz[i];
    }
#pragma unroll
    for (int i = 0; i < NP; i++)
#pragma unroll
        for (int j = i + 1; j < NP; j++) {
            const double dx = qx[j] - qx[i], dy = qy[j] - qy[i], dz = qz[j] - qz[i];
            const double w = nbt_rsqrt3(fma(dx, dx, fma(dy, dy, dz * dz)));
            const double fi = c.Gm[j + 1] * w, fj = -c.Gm[i + 1] * w;
            px[i] = fma(fi, dx, px[i]); py[i] = fma(fi, dy, py[i]); pz[i] = fma(fi, dz, pz[i]);
            px[j] = fma(fj, dx, px[j]); py[j] = fma(fj, dy, py[j]); pz[j] = fma(fj, dz, pz[j]);
        }
    // inertial -> Jacobi: a'_i = a_i - A_{i-1},  A_i = A_{i-1} + (m_i/M_i) a'_i,  A_0 = a_0
    double Ax = a0x, Ay = a0y, Az = a0z;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        double jx = px[i] - Ax, jy = py[i] - Ay, jz = pz[i] - Az;
        if (i + 1 < NP) { Ax = fma(c.mr[i], jx, Ax); Ay = fma(c.mr[i], jy, Ay); Az = fma(c.mr[i], jz, Az); }
        if (i >= 1) { // + G M_i r'_i / r'_i^3
            const double ir = s.ir[i];
            const double f = (c.GM[i] * ir) * (ir * ir);
            jx = fma(f, s.x[i], jx); jy = fma(f, s.y[i], jy); jz = fma(f, s.z[i], jz);
        }
        ax[i] = jx; ay[i] = jy; az[i] = jz;
    }
}

// acceleration of a (possibly force-gradient) kick: a'(q') for gd == 0, else a'(q' + gd a'(q'));
// gd = g dt^2 is uniform over the launch
template <int NP>
NBT_HD void nbt_accel_grad(const nbt_state<NP>& s, const nbt_consts<NP>& c, double gd, double (&ax)[NP],
                           double (&ay)[NP], double (&az)[NP]) {
    nbt_accel<NP>(s, c, ax, ay, az);
    if (gd != 0.0) {
        nbt_state<NP> t;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            t.x[i] = fma(gd, ax[i], s.x[i]); t.y[i] = fma(gd, ay[i], s.y[i]); t.z[i] = fma(gd, az[i], s.z[i]);
            if (i >= 1) nbt_radius(t.x[i], t.y[i], t.z[i], t.r[i], t.ir[i]);
        }
        nbt_accel<NP>(t, c, ax, ay, az);
    }
}

// composition coefficients: K(b[0]) D(a[0]) K(b[1]) ... D(a[S-1]) K(b[S]).
// g[i] != 0 makes kick i a force-gradient kick (g[0] == g[S]: the two end kicks of consecutive steps
// share one evaluation, which is why (ax,ay,az) carries the kick acceleration): it is evaluated
// at the Jacobi positions displaced by g[i] dt^2 a'(q'), which adds g[i] dt^2 (a'.grad) a' — the
// commutator {B,{A,B}} = sum_i m'_i |a'_i|^2 of the interaction with the kinetic part of the Kepler
// terms — without second derivatives (error of the displacement trick: O(dt^4 eps^3)).
struct nbt_scheme {
    int S;
    double a[NBT_MAX_STAGES];
    double b[NBT_MAX_STAGES + 1];
    double g[NBT_MAX_STAGES + 1];
};

NBT_HD nbt_scheme nbt_make_scheme(int id) {
    nbt_scheme s;
    for (int i = 0; i < NBT_MAX_STAGES; i++) s.a[i] = 0.0;
    for (int i = 0; i <= NBT_MAX_STAGES; i++) s.b[i] = s.g[i] = 0.0;
    switch (id) {
    case NBT_SCHEME_C4: { // Simpson kicks (1/6, 2/3, 1/6) around two half drifts, gradient term in the middle
        // kick (Chin 1997, algorithm 4A, applied to the Wisdom-Holman splitting): both dt^2 terms cancel,
        // O(eps dt^4 + eps^2 dt^4) with two Kepler drifts and three force evaluations per step
        s.S = 2;
        s.a[0] = 0.5; s.a[1] = 0.5;
        s.b[0] = 1.0 / 6.0; s.b[1] = 2.0 / 3.0; s.b[2] = 1.0 / 6.0;
        s.g[1] = NBT_C4_GRAD;
    } break;
    case NBT_SCHEME_SBABC3: { // Laskar & Robutel (2001) SBAB_3 = 4-point Lobatto, O(eps dt^6 + eps^2 dt^2), with
        // their corrector c dt^3 {{A,B},B}, c = (13 - 5 sqrt 5)/288, folded into the end kicks as a
        // force-gradient term g = c / b[0]: O(eps dt^6 + eps^2 dt^4) with three drifts and four force
        // evaluations per step (the end kick's pair is shared by consecutive steps)
        const double s5 = 2.2360679774997896964091737;
        s.S = 3;
        s.a[0] = 0.5 * (1.0 - 1.0 / s5); s.a[1] = 1.0 / s5; s.a[2] = s.a[0];
        s.b[0] = 1.0 / 12.0; s.b[1] = 5.0 / 12.0; s.b[2] = 5.0 / 12.0; s.b[3] = 1.0 / 12.0;
        s.g[0] = s.g[3] = (13.0 - 5.0 * s5) / 24.0;
    } break;
    case NBT_SCHEME_Y4: { // Yoshida (1990) triple jump of kick-drift-kick
        const double w1 = 1.351207191959657634047688, w0 = -1.702414383919315268095376;
        s.S = 3;
        s.a[0] = w1; s.a[1] = w0; s.a[2] = w1;
        s.b[0] = 0.5 * w1; s.b[1] = 0.5 * (w0 + w1); s.b[2] = 0.5 * (w0 + w1); s.b[3] = 0.5 * w1;
    } break;
    case NBT_SCHEME_SABA4: { // Laskar & Robutel (2001) SBAB_4 = 5-point Lobatto: O(eps dt^8 + eps^2 dt^2)
        const double c1 = 0.1726731646460114281008538, c2 = 0.3273268353539885718991462;
        s.S = 4;
        s.a[0] = c1; s.a[1] = c2; s.a[2] = c2; s.a[3] = c1;
        s.b[0] = 1.0 / 20.0; s.b[1] = 49.0 / 180.0; s.b[2] = 16.0 / 45.0; s.b[3] = 49.0 / 180.0; s.b[4] = 1.0 / 20.0;
    } break;
    case NBT_SCHEME_M64: { // order (6,4): O(eps dt^6 + eps^2 dt^4); tools/derive_schemes.py
        const double t1 = -0.0437514219173741137407351, t2 = 0.5437514219173741137407351;
        s.S = 4;
        s.a[0] = t1; s.a[1] = t2; s.a[2] = t2; s.a[3] = t1;
        s.b[0] = 0.5316386245813511790852625; s.b[1] = -0.3086019704406066393237719;
        s.b[2] = 0.5539266917185109204770189; s.b[3] = -0.3086019704406066393237719;
        s.b[4] = 0.5316386245813511790852625;
    } break;
    default: // NBT_SCHEME_WH2 kick-drift-kick leapfrog
        s.S = 1; s.a[0] = 1.0; s.b[0] = 0.5; s.b[1] = 0.5;
    }
    return s;
}

// ---------------------------------------------------------------------------------------------
// set-up: rebound.Simulation.add(**obj) in Jacobi fashion (nbody/simulation.py:45-47).  The
// orbital elements of body i are relative to the centre of mass of bodies 0..i-1 with
// mu = G M_i, so they ARE the Jacobi coordinates; move_to_com() only fixes the barycentre,
// which the Jacobi set does not carry.
// ---------------------------------------------------------------------------------------------
template <int NP>
NBT_HD void nbt_setup(const double* par, nbt_state<NP>& s, nbt_consts<NP>& c) {
    double M = par[NBT_P_M];
    c.Gm[0] = NBT_G * M;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const double* p = par + (i + 1) * NBT_NPAR;
        const double m = p[NBT_P_M];
        M += m;
        const double mu = NBT_G * M;
        c.GM[i] = mu; c.mr[i] = m / M; c.Gm[i + 1] = NBT_G * m;
        const double P = p[NBT_P_P], e = p[NBT_P_E];
        const double a = cbrt(P * P * mu / (4.0 * NBT_PI * NBT_PI));
        const double cf = cos(p[NBT_P_F]), sf = sin(p[NBT_P_F]);
        const double r = a * (1.0 - e * e) / (1.0 + e * cf);
        const double v0 = sqrt(mu / (a * (1.0 - e * e)));
        const double cO = cos(p[NBT_P_OMEGA_NODE]), sO = sin(p[NBT_P_OMEGA_NODE]);
        const double co = cos(p[NBT_P_OMEGA_PERI]), so = sin(p[NBT_P_OMEGA_PERI]);
        const double ci = cos(p[NBT_P_INC]), si = sin(p[NBT_P_INC]);
        const double cwf = co * cf - so * sf, swf = so * cf + co * sf;
        s.x[i] = r * (cO * cwf - sO * swf * ci);
        s.y[i] = r * (sO * cwf + cO * swf * ci);
        s.z[i] = r * (swf * si);
        const double vp = -(swf + e * so), vq = (cwf + e * co);
        s.vx[i] = v0 * (cO * vp - sO * vq * ci);
        s.vy[i] = v0 * (sO * vp + cO * vq * ci);
        s.vz[i] = v0 * (vq * si);
        nbt_radius(s.x[i], s.y[i], s.z[i], s.r[i], s.ir[i]);
    }
    c.k02 = c.k12 = 0.0;
    if (NP == 2) { const double q = c.GM[1] / c.GM[0]; c.k02 = c.Gm[0] * q; c.k12 = c.Gm[1] * q; }
}

// one composition step of size dt; (ax,ay,az) holds the acceleration at the current positions
// on entry and on exit (shared between the last kick of a step and the first of the next)
// LAT: the latency variant (Np <= 2) — branch-free drifts fused into one basic block; ~20 % faster for a
// lone warp (small batches, the long systems of a heterogeneous one) at ~40 more registers, i.e. 4
// instead of 6 resident CTAs per SM and ~6 % less steady-state throughput (profiles/README.md)
template <int NP, bool LAT>
NBT_HD void nbt_step(nbt_state<NP>& s, const nbt_consts<NP>& c, const nbt_scheme& sc, double dt,
                     double (&ax)[NP], double (&ay)[NP], double (&az)[NP], int& status) {
    {
        const double kb = sc.b[0] * dt;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            s.vx[i] = fma(kb, ax[i], s.vx[i]); s.vy[i] = fma(kb, ay[i], s.vy[i]); s.vz[i] = fma(kb, az[i], s.vz[i]);
        }
    }
    for (int st = 0; st < sc.S; st++) {
        const double h = sc.a[st] * dt;
        if (NP >= 3 || LAT) {
            // branch-free drifts, NBT_DRIFT_GROUP planets at a time: solve (one basic block), the rare failures redone out
            // of line, apply (one basic block) — see nbt_kepler_solve
            constexpr int GR = (NP <= 2) ? NP : NBT_DRIFT_GROUP(NP);
#pragma unroll
            for (int i0 = 0; i0 < NP; i0 += GR) {
                nbt_kepler_mid_t m[GR];
                bool ok[GR];
                bool all_ok = true;
#pragma unroll
                for (int q = 0; q < GR; q++) {
                    ok[q] = true;
                    if (i0 + q < NP) {
                        const int i = i0 + q;
                        ok[q] = nbt_kepler_solve<(NP <= 2)>(c.GM[i], h, s.x[i], s.y[i], s.z[i], s.vx[i], s.vy[i], s.vz[i], s.r[i], s.ir[i], m[q]);
                    }
                    all_ok = all_ok && ok[q];
                }
                if (!all_ok) {
#pragma unroll
                    for (int q = 0; q < GR; q++)
                        if (i0 + q < NP && !ok[q]) {
                            const int i = i0 + q;
                            const nbt_body o = nbt_kepler_finish(c.GM[i], h, s.x[i], s.y[i], s.z[i], s.vx[i], s.vy[i], s.vz[i], s.r[i], s.ir[i]);
                            s.x[i] = o.x; s.y[i] = o.y; s.z[i] = o.z; s.vx[i] = o.vx; s.vy[i] = o.vy; s.vz[i] = o.vz;
                            s.r[i] = o.r; s.ir[i] = o.ir; status |= o.status;
                            nbt_kepler_identity(o.r, o.ir, m[q]);
                        }
                }
#pragma unroll
                for (int q = 0; q < GR; q++)
                    if (i0 + q < NP) {
                        const int i = i0 + q;
                        nbt_kepler_apply(m[q], s.x[i], s.y[i], s.z[i], s.vx[i], s.vy[i], s.vz[i], s.r[i], s.ir[i]);
                    }
            }
        } else {
#pragma unroll
            for (int i = 0; i < NP; i++)
                nbt_kepler_drift<(NP <= 2)>(c.GM[i], h, s.x[i], s.y[i], s.z[i], s.vx[i], s.vy[i], s.vz[i], s.r[i], s.ir[i], status);
        }
        nbt_accel_grad<NP>(s, c, sc.g[st + 1] * dt * dt, ax, ay, az);
        const double kb = sc.b[st + 1] * dt;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            s.vx[i] = fma(kb, ax[i], s.vx[i]); s.vy[i] = fma(kb, ay[i], s.vy[i]); s.vz[i] = fma(kb, az[i], s.vz[i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// acos without branches.  libdevice's acos takes different code paths for |x| < 0.5 and |x| >= 0.5;
// arguments of pericentre are uniform on the circle, so the lanes of a warp always split and execute
// both.  Here both ranges share one evaluation of the classic rational approximation
//   asin(s) = s + s z p(z)/q(z),  z = s^2,  |s| <= 0.5   (coefficients of the FreeBSD msun e_asin.c)
// with  z = x^2, s = x                     acos(x) = pi/2 - asin(x)             for |x| <  0.5
//       z = (1-|x|)/2, s = sqrt(z)         acos(x) = 2 asin(s) | pi - 2 asin(s)  for |x| >= 0.5
// selected per lane.  Max error vs libm over [-1, 1]: 3e-16 (tests/test_hostsim.py).
// ---------------------------------------------------------------------------------------------
NBT_HD double nbt_acos(double x) {
    const double ax = fabs(x);
    const bool big = ax >= 0.5;
    const double zb = 0.5 * (1.0 - ax);
    const double z = big ? zb : x * x;
    const double sb = (zb > 0.0) ? zb * nbt_rsqrt(zb) : 0.0;         // sqrt(zb)
    const double sv = big ? sb : x;
    const double p = z * fma(z, fma(z, fma(z, fma(z, fma(z, 3.47933107596021167570e-05, 7.91534994289814532176e-04),
                     -4.00555345006794114027e-02), 2.01212532134862925881e-01), -3.25565818622400915405e-01),
                     1.66666666666666657415e-01);
    const double q = fma(z, fma(z, fma(z, fma(z, 7.70381505559019352791e-02, -6.88283971605453293030e-01),
                     2.02094576023350569471e+00), -2.40339491173441421878e+00), 1.0);
    const double as = fma(sv, p * nbt_rcp(q), sv);                     // asin(sv)
    const double small = 0.5 * NBT_PI - as;
    const double large = (x > 0.0) ? 2.0 * as : NBT_PI - 2.0 * as;
    return big ? large : small;
}

// ---------------------------------------------------------------------------------------------
// Jacobi osculating elements with REBOUND's conventions [3P] (primary = COM of lower-index
// bodies, mu = G M_i): what ps[j].P/a/e/inc/Omega/omega/M return at nbody/simulation.py:150.
// ---------------------------------------------------------------------------------------------
// REBOUND moves f, l, M, theta and omega (not Omega, pomega) into [0, 2 pi) at the end of its orbit conversion
// (reb_tools_mod2pi: fmod, then + 2 pi if negative) [3P]
NBT_HD double nbt_mod2pi(double f) {
    f = fmod(f, 2.0 * NBT_PI);
    return (f < 0.0) ? f + 2.0 * NBT_PI : f;
}

NBT_HD double nbt_acos2(double num, double den, double disamb) {
    const double cth = num * nbt_rcp(den);
    double val;
    if (cth > -1.0 && cth < 1.0) val = nbt_acos(cth);
    else val = (cth <= -1.0) ? NBT_PI : 0.0;
    return (disamb < 0.0) ? -val : val;
}

struct nbt_orbit { double P, a, e, inc, Omega, omega, M; };

// MODE 0: a, e, inc (the analyze() means); 1: + omega; 2: the full set P, Omega, omega, M.
// Reciprocals / square roots go through the MUFU-seeded helpers (<= 1 ulp); the acos calls are
// libdevice's.
// r, ir: |r| and 1/|r| of (x, y, z) — the caller has just refreshed them (nbt_radius).
template <int MODE>
NBT_HD_COLD nbt_orbit nbt_elements(double mu, double x, double y, double z, double vx, double vy, double vz,
                                   double r, double ir) {
    nbt_orbit o;
    const double hx = y * vz - z * vy, hy = z * vx - x * vz, hz = x * vy - y * vx;
    const double h2 = fma(hx, hx, fma(hy, hy, hz * hz));
    const double v2 = fma(vx, vx, fma(vy, vy, vz * vz));
    const double ih = nbt_rsqrt(h2);
    const double vc2 = mu * ir;
    o.a = -mu * nbt_rcp(v2 - 2.0 * vc2);
    const double rv = fma(x, vx, fma(y, vy, z * vz));
    const double vd2 = v2 - vc2;
    const double imu = nbt_rcp(mu);
    const double ex = imu * (vd2 * x - rv * vx), ey = imu * (vd2 * y - rv * vy), ez = imu * (vd2 * z - rv * vz);
    const double e2 = fma(ex, ex, fma(ey, ey, ez * ez));
    o.e = (e2 > 0.0) ? e2 * nbt_rsqrt(e2) : 0.0;
    {
        const double ci = hz * ih;
        if (fabs(ci) < 0.03) {
            // near edge-on (every transiting system): acos(x) = pi/2 - asin(x), asin by its Maclaurin
            // series; the first neglected term x^11 * 63/2816 is < 4e-19 for |x| < 0.03
            const double c2 = ci * ci;
            const double p = fma(c2, fma(c2, fma(c2, fma(c2, 35.0 / 1152.0, 15.0 / 336.0), 3.0 / 40.0), 1.0 / 6.0), 1.0);
            o.inc = fma(-ci, p, 0.5 * NBT_PI);
        } else {
            o.inc = (ci > -1.0 && ci < 1.0) ? nbt_acos(ci) : ((ci <= -1.0) ? NBT_PI : 0.0);
        }
        if (!(h2 > 0.0)) o.inc = NAN;
    }
    o.P = o.Omega = o.omega = o.M = 0.0;
    if (MODE >= 1) {
        const double nx = -hy, ny = hx, n2 = nx * nx + ny * ny;
        double nn = 0.0;
        if (MODE >= 2 || o.inc < 1e-8 || o.inc > NBT_PI - 1e-8) nn = (n2 > 0.0) ? n2 * nbt_rsqrt(n2) : 0.0;
        if (o.inc < 1e-8 || o.inc > NBT_PI - 1e-8) {
            const double Omega = nbt_acos2(nx, nn, ny);
            const double pomega = nbt_acos2(ex, o.e, ey);
            o.omega = (o.inc < 0.5 * NBT_PI) ? pomega - Omega : Omega - pomega;
        } else {
            // cos(omega) = n.e / (|n| |e|): one rsqrt of the product instead of a sqrt and a reciprocal
            const double den2 = n2 * e2;
            const double cth = (den2 > 0.0) ? (nx * ex + ny * ey) * nbt_rsqrt(den2) : 2.0;
            const double val = (cth > -1.0 && cth < 1.0) ? nbt_acos(cth) : ((cth <= -1.0) ? NBT_PI : 0.0);
            o.omega = (ez < 0.0) ? -val : val;
        }
        if (o.e <= 1e-8) o.omega = 0.0;
        o.omega = nbt_mod2pi(o.omega);
        if (MODE >= 2) {
            const double vr = rv * ir;
            const double nmm = o.a / fabs(o.a) * sqrt(fabs(mu / (o.a * o.a * o.a)));
            o.P = 2.0 * NBT_PI / nmm;
            o.Omega = nbt_acos2(nx, nn, ny);
            if (o.e < 1.0) {
                const double ea = nbt_acos2(1.0 - r / o.a, o.e, vr);
                o.M = ea - o.e * sin(ea);
            } else {
                double ea = acosh((1.0 - r / o.a) / o.e);
                if (vr < 0.0) ea = -ea;
                o.M = o.e * sinh(ea) - ea;
            }
            o.M = nbt_mod2pi(o.M);
        }
    }
    return o;
}

// Branch-free a, e, inc (and omega) for the common case — a near edge-on, non-circular orbit — with the same
// arithmetic as nbt_elements; returns whether that case holds (else the caller redoes the sample through
// nbt_elements).  With no branch inside, the element chains of two planets interleave (the sampling work is
// a quarter of the bench step and otherwise runs at ILP 1).
template <bool OMEGA>
NBT_HD bool nbt_elements_fast(double mu, double x, double y, double z, double vx, double vy, double vz,
                              double ir, double& a, double& e, double& inc, double& omega) {
    const double hx = y * vz - z * vy, hy = z * vx - x * vz, hz = x * vy - y * vx;
    const double h2 = fma(hx, hx, fma(hy, hy, hz * hz));
    const double v2 = fma(vx, vx, fma(vy, vy, vz * vz));
    const double ih = nbt_rsqrt(h2);
    const double vc2 = mu * ir;
    a = -mu * nbt_rcp(v2 - 2.0 * vc2);
    const double rv = fma(x, vx, fma(y, vy, z * vz));
    const double vd2 = v2 - vc2;
    const double imu = nbt_rcp(mu);
    const double ex = imu * (vd2 * x - rv * vx), ey = imu * (vd2 * y - rv * vy), ez = imu * (vd2 * z - rv * vz);
    const double e2 = fma(ex, ex, fma(ey, ey, ez * ez));
    e = e2 * nbt_rsqrt(e2);
    const double ci = hz * ih;
    const double c2 = ci * ci;
    const double p = fma(c2, fma(c2, fma(c2, fma(c2, 35.0 / 1152.0, 15.0 / 336.0), 3.0 / 40.0), 1.0 / 6.0), 1.0);
    inc = fma(-ci, p, 0.5 * NBT_PI);
    bool ok = fabs(ci) < 0.03 && h2 > 0.0 && e2 > 1e-16;
    omega = 0.0;
    if (OMEGA) {
        const double nx = -hy, ny = hx, n2 = nx * nx + ny * ny;
        const double den2 = n2 * e2;
        const double cth = (nx * ex + ny * ey) * nbt_rsqrt(den2);
        ok = ok && den2 > 0.0 && cth > -1.0 && cth < 1.0;
        const double val = nbt_acos(cth);
        omega = (ez < 0.0) ? 2.0 * NBT_PI - val : val;   // in [0, 2 pi) like nbt_mod2pi (val > 0 here: |cth| < 1)
    }
    return ok;
}

// Mid-transit time inside one output interval.  GRID_LINEAR: the reference's find_zero
// (nbody/tools.py:61-65).  REFINED: Newton on the cubic Hermite interpolant of dx over the interval.
// Out of line: it runs once per transit, and inlining it per planet bloats the sampling loop.
NBT_HD_COLD double nbt_crossing_time(int tt_mode, double t_prev, double t, double dx_prev, double dx,
                                     double dvx_prev, double dvx) {
    if (tt_mode == NBT_TT_GRID_LINEAR) {
        const double m = (dx - dx_prev) / (t - t_prev);
        return -dx_prev / m + t_prev;
    }
    const double hh = t - t_prev;
    const double p0 = dx_prev, p1 = dx, m0 = dvx_prev * hh, m1 = dvx * hh;
    double u = p0 / (p0 - p1);
    if (!(u >= 0.0 && u <= 1.0)) u = 0.5;
    for (int it = 0; it < 6; it++) {
        const double u2 = u * u, u3 = u2 * u;
        const double f = (2 * u3 - 3 * u2 + 1) * p0 + (u3 - 2 * u2 + u) * m0 + (-2 * u3 + 3 * u2) * p1 + (u3 - u2) * m1;
        const double fp = (6 * u2 - 6 * u) * p0 + (3 * u2 - 4 * u + 1) * m0 + (-6 * u2 + 6 * u) * p1 + (3 * u2 - 2 * u) * m1;
        u -= f / fp;
    }
    return t_prev + u * hh;
}

// ---------------------------------------------------------------------------------------------
// Whole-system driver: integrate + sample on np.linspace(0, Ndays, Noutputs) + transit detection
// (nbody/simulation.py:131-170 fused; trajectories never reach memory unless asked for).
// ---------------------------------------------------------------------------------------------
struct nbt_run_args {
    // inputs
    const double* params;      // [B][NP+1][NBT_NPAR]
    const int32_t* substeps;   // [B] or nullptr
    const int32_t* order;      // [B] launch order (systems sorted by substeps) or nullptr
    const int64_t* tt_offsets; // [B*NP+1]
    const double* n_days_sys;      // [B] per-system grids, or nullptr (uniform n_days / n_outputs)
    const int32_t* n_outputs_sys;  // [B]
    // outputs
    double* tt; int32_t* n_tt;
    double* mean_a; double* mean_e; double* mean_inc; double* mean_omega;
    int32_t* status;
    double* rv; double* orbit_xz; double* series;
    // config
    int32_t B, n_outputs, substeps_all, tt_mode, flags, orbit_stride;
    double n_days;
    nbt_scheme scheme;      // composition of the systems with positive sub-step counts
    nbt_scheme scheme_alt;  // NBT_SCHEME_AUTO: C4, for the trailing launch-order slots of a launch (whole blocks)
};

// NaN in every optional-series sample from output o on of a system that stopped (the buffers are caller-owned and
// reused: whatever they held before must not pass for a trajectory)
// (pointers by value: a reference to the kernel's argument block would force a copy of it into local memory)
NBT_HD_COLD void nbt_fill_broken(double* rv, double* orbit_xz, double* series, int B, int NP, int orbit_stride, int sys, int o0, int NO) {
    const int W = 3 + 10 * NP;
    for (int o = o0; o < NO; o++) {
        if (rv && o > 0) rv[(size_t)(o - 1) * B + sys] = NAN;
        if (orbit_xz && (o % orbit_stride) == 0) {
            double* dst = orbit_xz + (((size_t)(o / orbit_stride) * B + sys) * NP) * 2;
            for (int q = 0; q < 2 * NP; q++) dst[q] = NAN;
        }
        if (series) {
            double* row = series + ((size_t)o * B + sys) * W;
            for (int q = 0; q < W; q++) row[q] = NAN;
        }
    }
}

// Is a system outside the envelope the sub-step rule was calibrated on?  A planet pair closer than
// 4 mutual Hill radii ((a_j - a_i) / R_H, R_H = ((m_i + m_j) / 3 M*)^(1/3) (a_i + a_j) / 2, a ~ P^(2/3)), an
// eccentricity above 0.1, or a planet heavier than 0.02 M* (the reference's priors reach 0.018; the round-2 sweep has
// 1115 systems with a planet above 0.013 M* and no close pair, all within 0.08 of a tolerance).  par: [n_bodies][NBT_NPAR].
NBT_HD_COLD bool nbt_uncalibrated_impl(const double* par, int N) {
    const double Ms = par[NBT_P_M];
    bool out = false;
    for (int j = 1; j < N; j++) {
        const double Pj = par[j * NBT_NPAR + NBT_P_P], mj = par[j * NBT_NPAR + NBT_P_M];
        if (par[j * NBT_NPAR + NBT_P_E] > 0.1 || mj > 0.02 * Ms) out = true;
        for (int i = 1; i < j; i++) {
            const double Pi = par[i * NBT_NPAR + NBT_P_P];
            if (!(Pi > 0.0) || !(Pj > 0.0) || !(Ms > 0.0)) continue;
            const double ai = cbrt(Pi * Pi), aj = cbrt(Pj * Pj);
            const double rh = cbrt((par[i * NBT_NPAR + NBT_P_M] + mj) / (3.0 * Ms)) * 0.5 * (ai + aj);
            if (fabs(aj - ai) < 4.0 * rh) out = true;
        }
    }
    return out;
}

// SERIES = the NBT_F_SERIES variant (full element set per sample, generate() compatibility); it is a
// separate instantiation so that its large code stays out of the instruction caches of the batched path
template <int NP, bool SERIES, bool LAT = false>
NBT_HD void nbt_simulate_system(const nbt_run_args& A, const nbt_scheme& sc, int sys) {
    nbt_state<NP> s;
    nbt_consts<NP> c;
    const double* par = A.params + (size_t)sys * (NP + 1) * NBT_NPAR;
    nbt_setup<NP>(par, s, c);
    int status = nbt_uncalibrated_impl(par, NP + 1) ? NBT_S_UNCALIBRATED : 0;
    const int B = A.B;
    const int NO = (A.n_outputs_sys != nullptr) ? A.n_outputs_sys[sys] : A.n_outputs;     // generate_simulations.py:45-50
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
    const int n_tt_planets = (A.flags & NBT_F_PLANET1_ONLY) ? 1 : NP;
    const double rv_scale = 1.496e11 * ((double)NO / (n_days * 24.0 * 60.0 * 60.0)); // simulation.py:157,186
    const int W = 3 + 10 * NP;

    double ax[NP], ay[NP], az[NP];
    nbt_accel_grad<NP>(s, c, sc.g[0] * dt * dt, ax, ay, az);

    double dx_prev[NP], dvx_prev[NP];
    double sum_a[NP], sum_e[NP], sum_i[NP], sum_o[NP];
    int cnt[NP];
    double xs_prev = 0.0, t_prev = 0.0;
#pragma unroll
    for (int j = 0; j < NP; j++) { sum_a[j] = sum_e[j] = sum_i[j] = sum_o[j] = 0.0; cnt[j] = 0; dx_prev[j] = dvx_prev[j] = 0.0; }

    for (int o = 0; o < NO; o++) {
        if (o > 0) {
            for (int sub = 0; sub < k; sub++) nbt_step<NP, LAT>(s, c, sc, dt, ax, ay, az, status);
        }
        const double t = (o == NO - 1) ? n_days : (double)o * step;
        if (o > 0) { // refresh the carried radii from the coordinates (bounds their round-off drift)
#pragma unroll
            for (int j = 0; j < NP; j++) nbt_radius(s.x[j], s.y[j], s.z[j], s.r[j], s.ir[j]);
        }
        // heliocentric x (= x_planet - x_star) and star barycentric position
        double dx[NP], dvx[NP];
        double Sx = 0.0, Svx = 0.0;
#pragma unroll
        for (int j = 0; j < NP; j++) {
            dx[j] = s.x[j] + Sx; dvx[j] = s.vx[j] + Svx;
            Sx = fma(c.mr[j], s.x[j], Sx); Svx = fma(c.mr[j], s.vx[j], Svx);
        }
        const double xs = -Sx;
        // a system whose state broke (collision, overflow, failed Kepler solve) stops here: its
        // lane idles instead of dragging the warp through the slow fallback for the rest of the run
        if (!isfinite(Sx + Svx) || (status & (NBT_S_KEPLER | NBT_S_NONFINITE))) {
            status |= NBT_S_NONFINITE;
            if (want_rv || want_orbit || want_series)   // the samples that will never come read NaN
                nbt_fill_broken(want_rv ? A.rv : nullptr, want_orbit ? A.orbit_xz : nullptr, want_series ? A.series : nullptr, B, NP, A.orbit_stride, sys, o, NO);
            break;
        }
        if (o > 0) {
#pragma unroll
            for (int j = 0; j < NP; j++) {
                if (j < n_tt_planets && dx_prev[j] >= 0.0 && dx[j] <= 0.0) { // simulation.py:168
                    const double t0 = nbt_crossing_time(A.tt_mode, t_prev, t, dx_prev[j], dx[j], dvx_prev[j], dvx[j]);
                    const int64_t off = A.tt_offsets[(size_t)sys * NP + j];
                    const int64_t cap = A.tt_offsets[(size_t)sys * NP + j + 1] - off;
                    if (cnt[j] < cap) A.tt[off + cnt[j]] = t0; else status |= NBT_S_TT_OVERFLOW;
                    cnt[j]++;
                }
            }
            if (want_rv) A.rv[(size_t)(o - 1) * B + sys] = (xs - xs_prev) * rv_scale;             // simulation.py:186
        }
#pragma unroll
        for (int j = 0; j < NP; j++) { dx_prev[j] = dx[j]; dvx_prev[j] = dvx[j]; }
        xs_prev = xs; t_prev = t;

        if (NBT_FAST_ELEMENTS(NP, LAT) && want_means && !want_series) {
            // the element chains of two planets at a time in one basic block (nbt_elements_fast), rare cases redone
#pragma unroll
            for (int j0 = 0; j0 < NP; j0 += 2) {
                double ea[2], ee[2], ei[2], eo[2];
                bool ok[2] = {true, true};
                if (want_omega) {
#pragma unroll
                    for (int q = 0; q < 2; q++)
                        if (j0 + q < NP) {
                            const int j = j0 + q;
                            ok[q] = nbt_elements_fast<true>(c.GM[j], s.x[j], s.y[j], s.z[j], s.vx[j], s.vy[j], s.vz[j], s.ir[j], ea[q], ee[q], ei[q], eo[q]);
                        }
                } else {
#pragma unroll
                    for (int q = 0; q < 2; q++)
                        if (j0 + q < NP) {
                            const int j = j0 + q;
                            ok[q] = nbt_elements_fast<false>(c.GM[j], s.x[j], s.y[j], s.z[j], s.vx[j], s.vy[j], s.vz[j], s.ir[j], ea[q], ee[q], ei[q], eo[q]);
                        }
                }
                if (!(ok[0] && ok[1])) {
#pragma unroll
                    for (int q = 0; q < 2; q++)
                        if (j0 + q < NP && !ok[q]) {
                            const int j = j0 + q;
                            const nbt_orbit ob = want_omega ? nbt_elements<1>(c.GM[j], s.x[j], s.y[j], s.z[j], s.vx[j], s.vy[j], s.vz[j], s.r[j], s.ir[j])
                                                            : nbt_elements<0>(c.GM[j], s.x[j], s.y[j], s.z[j], s.vx[j], s.vy[j], s.vz[j], s.r[j], s.ir[j]);
                            ea[q] = ob.a; ee[q] = ob.e; ei[q] = ob.inc; eo[q] = ob.omega;
                        }
                }
#pragma unroll
                for (int q = 0; q < 2; q++)
                    if (j0 + q < NP) {
                        const int j = j0 + q;
                        sum_a[j] += ea[q]; sum_e[j] += ee[q]; sum_i[j] += ei[q]; sum_o[j] += eo[q];
                        if (!(ea[q] > 0.0)) status |= NBT_S_UNBOUND;
                    }
            }
        } else if (want_means && !want_series) {
            if (want_omega) {
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    const nbt_orbit ob = nbt_elements<1>(c.GM[j], s.x[j], s.y[j], s.z[j], s.vx[j], s.vy[j], s.vz[j], s.r[j], s.ir[j]);
                    sum_a[j] += ob.a; sum_e[j] += ob.e; sum_i[j] += ob.inc; sum_o[j] += ob.omega;
                    if (!(ob.a > 0.0)) status |= NBT_S_UNBOUND;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    const nbt_orbit ob = nbt_elements<0>(c.GM[j], s.x[j], s.y[j], s.z[j], s.vx[j], s.vy[j], s.vz[j], s.r[j], s.ir[j]);
                    sum_a[j] += ob.a; sum_e[j] += ob.e; sum_i[j] += ob.inc;
                    if (!(ob.a > 0.0)) status |= NBT_S_UNBOUND;
                }
            }
        }
        if (want_orbit || want_series) {
            // barycentric positions: r_0 = -S_NP, r_j = r_0 + q_j
            double Ty = 0.0, Tz = 0.0, qy[NP], qz[NP];
#pragma unroll
            for (int j = 0; j < NP; j++) {
                qy[j] = s.y[j] + Ty; qz[j] = s.z[j] + Tz;
                Ty = fma(c.mr[j], s.y[j], Ty); Tz = fma(c.mr[j], s.z[j], Tz);
            }
            if (want_orbit && (o % A.orbit_stride) == 0) {                                       // simulation.py:227-228
                const int q = o / A.orbit_stride;
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    double* dst = A.orbit_xz + (((size_t)q * B + sys) * NP + j) * 2;
                    dst[0] = dx[j] + xs; dst[1] = qz[j] - Tz;
                }
            }
            if (want_series) {                                                                   // simulation.py:141-150
                double* row = A.series + ((size_t)o * B + sys) * W;
                row[0] = xs; row[1] = -Ty; row[2] = -Tz;
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    const nbt_orbit ob = nbt_elements<2>(c.GM[j], s.x[j], s.y[j], s.z[j], s.vx[j], s.vy[j], s.vz[j], s.r[j], s.ir[j]);
                    double* p = row + 3 + 10 * j;
                    p[0] = dx[j] + xs; p[1] = qy[j] - Ty; p[2] = qz[j] - Tz;
                    p[3] = ob.P; p[4] = ob.a; p[5] = ob.e; p[6] = ob.inc; p[7] = ob.Omega; p[8] = ob.omega; p[9] = ob.M;
                    sum_a[j] += ob.a; sum_e[j] += ob.e; sum_i[j] += ob.inc; sum_o[j] += ob.omega;
                    if (!(ob.a > 0.0)) status |= NBT_S_UNBOUND;
                }
            }
        }
    }
    bool finite = true;
#pragma unroll
    for (int j = 0; j < NP; j++) {
        const size_t sp = (size_t)sys * NP + j;
        finite = finite && isfinite(s.x[j] + s.vx[j] + s.y[j] + s.vy[j] + s.z[j] + s.vz[j]);
        A.n_tt[sp] = cnt[j];
        if (want_means || want_series) {
            const double inv = 1.0 / (double)NO;
            if (A.mean_a) A.mean_a[sp] = sum_a[j] * inv;
            if (A.mean_e) A.mean_e[sp] = sum_e[j] * inv;
            if (A.mean_inc) A.mean_inc[sp] = sum_i[j] * inv;
            if (A.mean_omega && (want_omega || want_series)) A.mean_omega[sp] = sum_o[j] * inv;
        }
    }
    if (!finite) status |= NBT_S_NONFINITE;
    A.status[sys] = status;
}
