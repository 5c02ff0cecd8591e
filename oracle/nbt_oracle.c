/*
 * nbt_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the reference's generate()/integrate()/analyze() path
 * (pearsonkyle/Nbody-AI).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (nbody_ai_b200 + libnbt.so) never
 * does and has no CPU fallback.
 *
 * PARITY PINNING.  The arithmetic of the reference path lives in third-party packages that
 * are neither vendored nor pinned nor installable here: `rebound` (nbody/simulation.py:8,
 * a 3.2x release is implied by sim.ri_tes at :35-38) and `astropy` LombScargle
 * (nbody/tools.py:4).  The reference ships no executable test.  This oracle is therefore
 * pinned (tests/test_oracle.py) against
 *   G1  the numeric table + periodogram peaks of figures/report_simulation.png (2-3 digits),
 *   G2  the 30 transit times of sim_data.txt (noise-limited, sigma 0.5 min),
 *   an independent scipy DOP853 (rtol 1e-13) integration of the same initial conditions,
 *   scipy.signal.lombscargle(floating_mean=True, normalize=True) for the periodogram.
 * In the strict sense (no runnable reference, no bit-level vectors) parity is UNPINNED beyond
 * those artefacts; DESIGN.md says the same.
 *
 * What is restated, in reference order:
 *   nbody/simulation.py:27-52    generate(): units day/AU/Msun, rebound.add(**obj) per body
 *                                (Jacobi primaries, a from P), move_to_com()
 *   nbody/simulation.py:33-42    integrator='whfast' preset: WHFast, corrector 5, safe_mode 0, dt = 1e-3 d
 *   nbody/simulation.py:131-160  integrate(): default REBOUND integrator (IAS15, adaptive
 *                                Gauss-Radau 15, Rein & Spiegel 2015) landing on
 *                                np.linspace(0, Ndays, Noutputs); star x,y,z; planet x,y,z and
 *                                Jacobi osculating P,a,e,inc,Omega,omega,M
 *   nbody/simulation.py:163-170  transit_times();  nbody/tools.py:61-65 find_zero()
 *   nbody/simulation.py:172-177  TTV() (2-parameter least squares)
 *   nbody/simulation.py:179-231  analyze()
 *   nbody/tools.py:22            maxavg();  nbody/tools.py:97-104 lomb_scargle()
 *   nbody/nested.py:85-95        myloglike()
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#include "../include/nbt.h"

#define MAXN NBT_MAX_BODIES
#define PI 3.14159265358979323846264338327950288

/* ------------------------------------------------------------------------------------------ */
/* rebound.Simulation.add(m=, P=, e=, inc=, Omega=, omega=, f=) + move_to_com()                 */
/* nbody/simulation.py:44-47.  [3P] REBOUND adds a body given by orbital elements relative to  */
/* the centre of mass of the bodies already present (Jacobi), with mu = G (M_interior + m).    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int n;
    double m[MAXN];
    double x[3 * MAXN], v[3 * MAXN];
} sys_t;

static void elements_to_state(double mu, double a, double e, double inc, double Om, double om,
                              double f, double* r3, double* v3) {
    /* Murray & Dermott rotation of the orbital-plane vector */
    double cf = cos(f), sf = sin(f);
    double r = a * (1.0 - e * e) / (1.0 + e * cf);
    double v0 = sqrt(mu / (a * (1.0 - e * e)));
    double cO = cos(Om), sO = sin(Om), co = cos(om), so = sin(om), ci = cos(inc), si = sin(inc);
    /* cos(om+f), sin(om+f) kept as products so rounding follows the element definition */
    double cwf = co * cf - so * sf, swf = so * cf + co * sf;
    r3[0] = r * (cO * cwf - sO * swf * ci);
    r3[1] = r * (sO * cwf + cO * swf * ci);
    r3[2] = r * (swf * si);
    /* velocity: v0 * [ -(sin(om+f) + e sin om) , (cos(om+f) + e cos om) ] rotated */
    double vp = -(swf + e * so), vq = (cwf + e * co);
    v3[0] = v0 * (cO * vp - sO * vq * ci);
    v3[1] = v0 * (sO * vp + cO * vq * ci);
    v3[2] = v0 * (vq * si);
}

static void build_system(const double* params, int n, sys_t* s) {
    s->n = n;
    double M = 0, cx[3] = {0, 0, 0}, cv[3] = {0, 0, 0}; /* running COM of bodies present */
    for (int i = 0; i < n; i++) {
        const double* p = params + (size_t)i * NBT_NPAR;
        double m = p[NBT_P_M];
        s->m[i] = m;
        if (i == 0) { /* {'m': M}: at the origin, at rest */
            for (int k = 0; k < 3; k++) s->x[k] = s->v[k] = 0.0;
        } else {
            double mu = NBT_G * (M + m);
            double P = p[NBT_P_P];
            double a = cbrt(P * P * mu / (4.0 * PI * PI));
            double r3[3], v3[3];
            elements_to_state(mu, a, p[NBT_P_E], p[NBT_P_INC], p[NBT_P_OMEGA_NODE],
                              p[NBT_P_OMEGA_PERI], p[NBT_P_F], r3, v3);
            for (int k = 0; k < 3; k++) {
                s->x[3 * i + k] = cx[k] + r3[k];
                s->v[3 * i + k] = cv[k] + v3[k];
            }
        }
        for (int k = 0; k < 3; k++) {
            cx[k] = (cx[k] * M + s->x[3 * i + k] * m) / (M + m);
            cv[k] = (cv[k] * M + s->v[3 * i + k] * m) / (M + m);
        }
        M += m;
    }
    /* move_to_com() */
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            s->x[3 * i + k] -= cx[k];
            s->v[3 * i + k] -= cv[k];
        }
}

/* ------------------------------------------------------------------------------------------ */
/* Direct-summation Newtonian accelerations (REBOUND gravity "basic", no softening)            */
/* ------------------------------------------------------------------------------------------ */
static void accel(const sys_t* s, const double* x, double* a) {
    int n = s->n;
    for (int i = 0; i < 3 * n; i++) a[i] = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            double dx = x[3 * j] - x[3 * i], dy = x[3 * j + 1] - x[3 * i + 1],
                   dz = x[3 * j + 2] - x[3 * i + 2];
            double r2 = dx * dx + dy * dy + dz * dz;
            double pre = NBT_G / (r2 * sqrt(r2));
            double pi_ = pre * s->m[j], pj = pre * s->m[i];
            a[3 * i] += pi_ * dx; a[3 * i + 1] += pi_ * dy; a[3 * i + 2] += pi_ * dz;
            a[3 * j] -= pj * dx;  a[3 * j + 1] -= pj * dy;  a[3 * j + 2] -= pj * dz;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* IAS15: 15th-order Gauss-Radau predictor-corrector with adaptive step (Rein & Spiegel 2015,  */
/* MNRAS 446, 1424 — the published algorithm behind rebound's default integrator; restated     */
/* from the paper, not from REBOUND's source, which is not available here).                    */
/*   a(s) = a0 + b0 s + b1 s^2 + ... + b6 s^7,  s in [0,1], sampled at the Radau nodes H[1..7] */
/* ------------------------------------------------------------------------------------------ */
static const double H[8] = {0.0,
                            0.0562625605369221464656521910318,
                            0.180240691736892364987579942780,
                            0.352624717113169637373907769648,
                            0.547153626330555383001448554766,
                            0.734210177215410531523210605558,
                            0.885320946839095768090359771030,
                            0.977520613561287501891174488626};
static double RR[8][8]; /* 1/(H[j]-H[k]) */
static double CC[7][7]; /* Newton basis -> monomial: b_k = sum_{j>=k} CC[j][k] g_j */
static double DD[7][7]; /* monomial -> Newton: g_k = sum_{j>=k} DD[j][k] b_j */
static int tables_ready = 0;

static void init_tables(void) {
    if (tables_ready) return;
    for (int j = 1; j < 8; j++)
        for (int k = 0; k < j; k++) RR[j][k] = 1.0 / (H[j] - H[k]);
    /* s (s-h1) ... (s-h_j) = sum_k CC[j][k] s^{k+1}  (long double to keep the tables exact) */
    long double c[7][7];
    memset(c, 0, sizeof c);
    c[0][0] = 1.0L;
    for (int j = 1; j < 7; j++) {
        for (int k = 0; k <= j; k++) {
            long double up = (k > 0) ? c[j - 1][k - 1] : 0.0L;
            long double same = (k < j) ? c[j - 1][k] : 0.0L;
            c[j][k] = up - (long double)H[j] * same;
        }
    }
    /* inverse of the (unit upper-triangular in (j>=k)) map: solve sum_{j>=k} c[j][k] g_j = b_k */
    long double d[7][7];
    memset(d, 0, sizeof d);
    for (int col = 0; col < 7; col++) { /* g as a function of unit b_col */
        long double g[7];
        for (int k = 6; k >= 0; k--) {
            long double rhs = (k == col) ? 1.0L : 0.0L;
            for (int j = k + 1; j < 7; j++) rhs -= c[j][k] * g[j];
            g[k] = rhs; /* c[k][k] == 1 */
        }
        for (int k = 0; k < 7; k++) d[col][k] = g[k];
    }
    for (int j = 0; j < 7; j++)
        for (int k = 0; k < 7; k++) {
            CC[j][k] = (double)c[j][k];
            DD[j][k] = (double)d[j][k];
        }
    tables_ready = 1;
}

typedef struct {
    sys_t s;
    double t, dt, dt_last;
    double a0[3 * MAXN];
    double b[7][3 * MAXN], e[7][3 * MAXN], g[7][3 * MAXN];
    double csx[3 * MAXN], csv[3 * MAXN]; /* compensated-summation carries */
    int have_b, gave_up;
    long n_steps, n_force, n_reject, max_steps;
} ias15_t;

static void ias15_init(ias15_t* I, const sys_t* s) {
    init_tables();
    memset(I, 0, sizeof *I);
    I->s = *s;
    I->dt = 1e-3; /* REBOUND's default initial dt */
    I->max_steps = 2000000000L;
}

static inline void add_cs(double* sum, double* carry, double inc) {
    double y = inc - *carry;
    double t = *sum + y;
    *carry = (t - *sum) - y;
    *sum = t;
}

/* try one step of size dt; returns 1 if accepted (state advanced), 0 if rejected (I->dt shrunk) */
static int ias15_step(ias15_t* I, double dt) {
    const int n3 = 3 * I->s.n;
    double* x0 = I->s.x; double* v0 = I->s.v; double* a0 = I->a0;
    double xs[3 * MAXN], as[3 * MAXN];
    const double eps_b = 1e-9, safety = 0.25;

    accel(&I->s, x0, a0); I->n_force++;

    /* predict b for this step from the previous step's b (ratio q), plus the last correction */
    if (I->have_b && dt / I->dt_last > 20.0) {
        /* a very large step-size increase (e.g. the step after a short landing step): do not
         * extrapolate the old series by q^7, start the predictor-corrector from zero */
        for (int i = 0; i < n3; i++)
            for (int k = 0; k < 7; k++) I->b[k][i] = I->e[k][i] = 0.0;
    } else if (I->have_b) {
        double q = dt / I->dt_last, q2 = q * q, q3 = q2 * q, q4 = q2 * q2, q5 = q4 * q, q6 = q3 * q3,
               q7 = q6 * q;
        for (int i = 0; i < n3; i++) {
            double b0 = I->b[0][i], b1 = I->b[1][i], b2 = I->b[2][i], b3 = I->b[3][i],
                   b4 = I->b[4][i], b5 = I->b[5][i], b6 = I->b[6][i];
            double be[7];
            for (int k = 0; k < 7; k++) be[k] = I->b[k][i] - I->e[k][i];
            I->e[0][i] = q * (b6 * 7.0 + b5 * 6.0 + b4 * 5.0 + b3 * 4.0 + b2 * 3.0 + b1 * 2.0 + b0);
            I->e[1][i] = q2 * (b6 * 21.0 + b5 * 15.0 + b4 * 10.0 + b3 * 6.0 + b2 * 3.0 + b1);
            I->e[2][i] = q3 * (b6 * 35.0 + b5 * 20.0 + b4 * 10.0 + b3 * 4.0 + b2);
            I->e[3][i] = q4 * (b6 * 35.0 + b5 * 15.0 + b4 * 5.0 + b3);
            I->e[4][i] = q5 * (b6 * 21.0 + b5 * 6.0 + b4);
            I->e[5][i] = q6 * (b6 * 7.0 + b5);
            I->e[6][i] = q7 * b6;
            for (int k = 0; k < 7; k++) I->b[k][i] = I->e[k][i] + be[k];
        }
    }
    for (int i = 0; i < n3; i++)
        for (int k = 0; k < 7; k++) {
            double gk = 0.0;
            for (int j = k; j < 7; j++) gk += DD[j][k] * I->b[j][i];
            I->g[k][i] = gk;
        }

    double pc_err = 1e300, pc_err_last = 2.0;
    int iter = 0;
    while (1) {
        if (pc_err < 1e-16) break;
        if (iter > 2 && pc_err_last <= pc_err) break; /* stopped converging (round-off floor) */
        if (iter >= 12) break;
        pc_err_last = pc_err;
        iter++;
        double max_a = 0.0, max_db6 = 0.0;
        for (int n = 1; n < 8; n++) {
            double s = H[n];
            for (int i = 0; i < n3; i++) {
                double p = I->b[6][i] * (1.0 / 72.0);
                p = p * s + I->b[5][i] * (1.0 / 56.0);
                p = p * s + I->b[4][i] * (1.0 / 42.0);
                p = p * s + I->b[3][i] * (1.0 / 30.0);
                p = p * s + I->b[2][i] * (1.0 / 20.0);
                p = p * s + I->b[1][i] * (1.0 / 12.0);
                p = p * s + I->b[0][i] * (1.0 / 6.0);
                p = p * s + a0[i] * 0.5;
                xs[i] = x0[i] - I->csx[i] + s * dt * (v0[i] + s * dt * p);
            }
            accel(&I->s, xs, as); I->n_force++;
            for (int i = 0; i < n3; i++) {
                double gk = (as[i] - a0[i]) * RR[n][0];
                for (int k = 1; k < n; k++) gk = (gk - I->g[k - 1][i]) * RR[n][k];
                double dg = gk - I->g[n - 1][i];
                I->g[n - 1][i] = gk;
                for (int k = 0; k < n; k++) I->b[k][i] += CC[n - 1][k] * dg;
                if (n == 7) {
                    double aa = fabs(as[i]), d6 = fabs(dg);
                    if (aa > max_a) max_a = aa;
                    if (d6 > max_db6) max_db6 = d6;
                }
            }
        }
        pc_err = max_db6 / max_a;
    }

    /* step-size control on the magnitude of the last series term */
    double max_a = 0.0, max_b6 = 0.0;
    for (int i = 0; i < n3; i++) {
        double aa = fabs(a0[i]), bb = fabs(I->b[6][i]);
        if (aa > max_a) max_a = aa;
        if (bb > max_b6) max_b6 = bb;
    }
    double err = max_b6 / max_a;
    double dt_new = (err > 0 && isfinite(err)) ? dt * pow(eps_b / err, 1.0 / 7.0) : dt / safety;
    if (fabs(dt_new / dt) < safety) { /* reject and retry with the smaller step */
        I->dt = dt_new;
        I->n_reject++;
        if (I->have_b) { /* undo prediction: restore b as of the end of the last step */
            for (int i = 0; i < n3; i++)
                for (int k = 0; k < 7; k++) I->b[k][i] = I->e[k][i] = 0.0;
            I->have_b = 0;
        }
        return 0;
    }
    if (dt_new / dt > 1.0 / safety) dt_new = dt / safety;

    /* advance with compensated summation */
    for (int i = 0; i < n3; i++) {
        double b0 = I->b[0][i], b1 = I->b[1][i], b2 = I->b[2][i], b3 = I->b[3][i], b4 = I->b[4][i],
               b5 = I->b[5][i], b6 = I->b[6][i];
        double dx2 = (b6 / 72.0 + b5 / 56.0 + b4 / 42.0 + b3 / 30.0 + b2 / 20.0 + b1 / 12.0 + b0 / 6.0 +
                      a0[i] / 2.0) * dt * dt;
        add_cs(&x0[i], &I->csx[i], dx2);
        add_cs(&x0[i], &I->csx[i], v0[i] * dt);
        double dv = (b6 / 8.0 + b5 / 7.0 + b4 / 6.0 + b3 / 5.0 + b2 / 4.0 + b1 / 3.0 + b0 / 2.0 + a0[i]) * dt;
        add_cs(&v0[i], &I->csv[i], dv);
    }
    /* e currently holds the prediction made for this step; keep (b - e) semantics for the next */
    I->t += dt;
    I->dt_last = dt;
    I->dt = dt_new;
    I->have_b = 1;
    I->n_steps++;
    return 1;
}

/* sim.integrate(t_end) with exact_finish_time=1: the last step is shortened to land on t_end */
static void ias15_integrate_to(ias15_t* I, double t_end) {
    while (I->t < t_end) {
        /* step budget: a close encounter drives the adaptive step towards zero (REBOUND would
         * crawl through it too); the oracle gives up and the caller marks the run non-finite */
        if (I->n_steps + I->n_reject > I->max_steps) { I->gave_up = 1; return; }
        double remaining = t_end - I->t;
        if (remaining <= 1e-14 * fabs(t_end)) { I->t = t_end; break; }
        double dt = I->dt, keep = I->dt;
        int shortened = 0;
        if (dt >= remaining) { dt = remaining; shortened = 1; }
        if (ias15_step(I, dt)) {
            if (shortened) { /* landed on the output; restore the un-shortened step suggestion */
                I->t = t_end;
                I->dt = keep;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* integrator='whfast' (nbody/simulation.py:39-42): sim.integrator = "whfast",                  */
/* ri_whfast.corrector = 5, ri_whfast.safe_mode = 0, dt left at REBOUND's default 1e-3 (days).  */
/* [3P] REBOUND's WHFast (Rein & Tamayo 2015), restated from the paper and the documented       */
/* behaviour of the settings: Wisdom-Holman drift-kick-drift in Jacobi coordinates — Kepler     */
/* drift of Jacobi coordinate i with mu = G (m_0 + ... + m_i), interaction kick from the         */
/* direct-sum accelerations without the star-planet-1 pair, transformed to Jacobi               */
/* accelerations, plus G eta_i r'_i / r'_i^3 for i >= 2 —; with safe_mode = 0 consecutive half  */
/* drifts are merged and the state is only "synchronised" (half drift + inverse corrector) at   */
/* the end of every sim.integrate(t); the symplectic corrector of order 5 (Wisdom, Holman &     */
/* Touma 1996: a_k = k sqrt(7/40), b_51 = -beta/6, b_52 = 5 beta/6, beta = 1/(48 sqrt(7/40)))   */
/* is applied when a synchronised state starts stepping and inverted at synchronisation;        */
/* exact_finish_time = 1: the step that would overshoot the requested time is replaced by one   */
/* of the remaining length (after a synchronisation), and dt returns to 1e-3 afterwards.        */
/* corrector_order = 0 switches the corrector off (tests: its effect on the error).             */
/* ------------------------------------------------------------------------------------------ */
static void sample_row(const sys_t* s, double* row);

typedef struct {
    int n, corrector, synced;
    double m[MAXN], eta[MAXN];          /* eta_i = m_0 + ... + m_i */
    double xj[3 * MAXN], vj[3 * MAXN];  /* Jacobi coordinates; index 0 = centre of mass */
    double t, dt, dt_last_done;
    long n_steps, max_steps;   /* budget: REBOUND's dt bookkeeping can collapse dt when an output interval is shorter than dt */
    int gave_up;
} whfast_t;

static void stumpff_cs(double z, double* c0, double* c1, double* c2, double* c3) {
    int n = 0;
    while (fabs(z) > 0.1) { z *= 0.25; n++; }
    /* c3 = sum (-z)^k / (2k+3)!, c2 = sum (-z)^k / (2k+2)! */
    double a3 = 1.0 / 6.0, a2 = 0.5, s3 = a3, s2 = a2;
    for (int k = 1; k < 12; k++) {
        a3 *= -z / ((2.0 * k + 2.0) * (2.0 * k + 3.0));
        a2 *= -z / ((2.0 * k + 1.0) * (2.0 * k + 2.0));
        s3 += a3; s2 += a2;
    }
    double C3 = s3, C2 = s2, C1 = 1.0 - z * C3, C0 = 1.0 - z * C2;
    for (; n > 0; n--) {
        C3 = (C2 + C0 * C3) * 0.25;
        C2 = C1 * C1 * 0.5;
        C1 = C0 * C1;
        C0 = 2.0 * C0 * C0 - 1.0;
    }
    *c0 = C0; *c1 = C1; *c2 = C2; *c3 = C3;
}

/* two-body advance by h (either sign) in universal variables: solve r0 G1 + eta G2 + mu G3 = h for X,
 * G_k = X^k c_k(beta X^2), beta = 2 mu / r0 - v0^2; Newton with a bisection safeguard, to round-off */
static int kepler_advance(double mu, double h, double* x, double* v) {
    const double r0 = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const double v2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double eta = x[0] * v[0] + x[1] * v[1] + x[2] * v[2];
    const double beta = 2.0 * mu / r0 - v2;
    double X = h / r0, lo = -INFINITY, hi = INFINITY;
    double G1 = 0, G2 = 0, G3 = 0, c0, c1, c2, c3, r = r0;
    int ok = 0;
    for (int it = 0; it < 60; it++) {
        stumpff_cs(beta * X * X, &c0, &c1, &c2, &c3);
        G1 = X * c1; G2 = X * X * c2; G3 = X * X * X * c3;
        const double F = r0 * G1 + eta * G2 + mu * G3 - h;
        r = r0 * c0 + eta * G1 + mu * G2;          /* dF/dX = new radius > 0: F is monotonic */
        if (F > 0) hi = X; else lo = X;
        double Xn = X - F / r;
        if (!(Xn > lo && Xn < hi)) Xn = (isfinite(lo) && isfinite(hi)) ? 0.5 * (lo + hi) : (F > 0 ? X - fabs(X) - 1e-3 : X + fabs(X) + 1e-3);
        if (fabs(Xn - X) <= 1e-16 * fabs(Xn) || Xn == X) { X = Xn; ok = 1; break; }
        X = Xn;
    }
    stumpff_cs(beta * X * X, &c0, &c1, &c2, &c3);
    G1 = X * c1; G2 = X * X * c2; G3 = X * X * X * c3;
    r = r0 * c0 + eta * G1 + mu * G2;
    const double f = 1.0 - mu * G2 / r0, g = h - mu * G3, fd = -mu * G1 / (r * r0), gd = 1.0 - mu * G2 / r;
    for (int k = 0; k < 3; k++) {
        const double xn = f * x[k] + g * v[k], vn = fd * x[k] + gd * v[k];
        x[k] = xn; v[k] = vn;
    }
    return ok;
}

static void wh_from_inertial(whfast_t* W, const sys_t* s, int corrector) {
    W->n = s->n; W->corrector = corrector; W->synced = 1; W->t = 0.0; W->dt = 1e-3; W->dt_last_done = 0.0; W->n_steps = 0; W->max_steps = 0; W->gave_up = 0;
    double M = 0.0, cx[3] = {0, 0, 0}, cv[3] = {0, 0, 0};     /* mass-weighted sums of bodies 0..i-1 */
    for (int i = 0; i < s->n; i++) {
        W->m[i] = s->m[i];
        if (i > 0)
            for (int k = 0; k < 3; k++) { W->xj[3 * i + k] = s->x[3 * i + k] - cx[k] / M; W->vj[3 * i + k] = s->v[3 * i + k] - cv[k] / M; }
        for (int k = 0; k < 3; k++) { cx[k] += s->m[i] * s->x[3 * i + k]; cv[k] += s->m[i] * s->v[3 * i + k]; }
        M += s->m[i];
        W->eta[i] = M;
    }
    for (int k = 0; k < 3; k++) { W->xj[k] = cx[k] / M; W->vj[k] = cv[k] / M; }
}

/* Jacobi -> inertial (positions and velocities): from the outermost body inwards */
static void wh_to_inertial(const whfast_t* W, sys_t* s) {
    s->n = W->n;
    double cx[3], cv[3];      /* centre of mass of bodies 0..i */
    for (int k = 0; k < 3; k++) { cx[k] = W->xj[k]; cv[k] = W->vj[k]; }
    for (int i = W->n - 1; i > 0; i--) {
        const double fr = W->m[i] / W->eta[i];
        for (int k = 0; k < 3; k++) {
            /* r_i = R_{i-1} + r'_i and R_i = R_{i-1} + fr r'_i  =>  R_{i-1} = R_i - fr r'_i */
            const double Rx = cx[k] - fr * W->xj[3 * i + k], Rv = cv[k] - fr * W->vj[3 * i + k];
            s->x[3 * i + k] = Rx + W->xj[3 * i + k]; s->v[3 * i + k] = Rv + W->vj[3 * i + k];
            cx[k] = Rx; cv[k] = Rv;
        }
        s->m[i] = W->m[i];
    }
    s->m[0] = W->m[0];
    for (int k = 0; k < 3; k++) { s->x[k] = cx[k]; s->v[k] = cv[k]; }
}

static void wh_drift(whfast_t* W, double h) {
    for (int i = 1; i < W->n; i++) kepler_advance(NBT_G * W->eta[i], h, W->xj + 3 * i, W->vj + 3 * i);
    for (int k = 0; k < 3; k++) W->xj[k] += h * W->vj[k];      /* centre of mass */
}

static void wh_kick(whfast_t* W, double h) {
    sys_t s;
    wh_to_inertial(W, &s);
    const int n = W->n;
    double a[3 * MAXN];
    for (int i = 0; i < 3 * n; i++) a[i] = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            if (i == 0 && j == 1) continue;                   /* in the Kepler part of the splitting */
            double d[3], r2 = 0.0;
            for (int k = 0; k < 3; k++) { d[k] = s.x[3 * j + k] - s.x[3 * i + k]; r2 += d[k] * d[k]; }
            const double pre = NBT_G / (r2 * sqrt(r2));
            for (int k = 0; k < 3; k++) { a[3 * i + k] += pre * s.m[j] * d[k]; a[3 * j + k] -= pre * s.m[i] * d[k]; }
        }
    /* inertial -> Jacobi accelerations: a'_i = a_i - (sum_{k<i} m_k a_k) / eta_{i-1} */
    double sa[3] = {W->m[0] * a[0], W->m[0] * a[1], W->m[0] * a[2]};
    for (int i = 1; i < n; i++) {
        double r2 = 0.0;
        for (int k = 0; k < 3; k++) r2 += W->xj[3 * i + k] * W->xj[3 * i + k];
        const double pre = (i > 1) ? NBT_G * W->eta[i] / (r2 * sqrt(r2)) : 0.0;
        for (int k = 0; k < 3; k++) {
            const double aj = a[3 * i + k] - sa[k] / W->eta[i - 1];
            W->vj[3 * i + k] += h * (aj + pre * W->xj[3 * i + k]);
            sa[k] += W->m[i] * a[3 * i + k];
        }
    }
}

static void wh_Z(whfast_t* W, double a, double b) {
    wh_drift(W, a); wh_kick(W, -b); wh_drift(W, -2.0 * a); wh_kick(W, b); wh_drift(W, a);
}

static void wh_corrector(whfast_t* W, double inv) {
    if (W->corrector != 5) return;
    const double a1 = 0.41833001326703777398908601289259374469640768464934, a2 = 2.0 * a1;
    const double b51 = -0.0083001986759332891664501193034244790614369381874871;
    const double b52 = 0.041500993379666445832250596517122395307184690937435;
    const double dt = W->dt;
    wh_Z(W, -a2 * dt, -inv * b51 * dt);
    wh_Z(W, -a1 * dt, -inv * b52 * dt);
    wh_Z(W, a1 * dt, inv * b52 * dt);
    wh_Z(W, a2 * dt, inv * b51 * dt);
}

static void wh_synchronize(whfast_t* W) {
    if (W->synced) return;
    wh_drift(W, 0.5 * W->dt);
    wh_corrector(W, -1.0);
    W->synced = 1;
}

static void wh_step(whfast_t* W) {
    if (W->synced) { wh_corrector(W, 1.0); wh_drift(W, 0.5 * W->dt); }
    else wh_drift(W, W->dt);
    W->t += 0.5 * W->dt;
    wh_kick(W, W->dt);
    W->synced = 0;
    W->t += 0.5 * W->dt;
    W->dt_last_done = W->dt;
    W->n_steps++;
}

/* sim.integrate(tmax), exact_finish_time = 1 */
static void wh_integrate_to(whfast_t* W, double tmax) {
    double last_full_dt = W->dt;
    int last_step = 0;
    for (;;) {
        if (W->t + W->dt >= tmax) {           /* the next step would overshoot */
            if (W->t == tmax) break;
            if (last_step) {
                double tscale = 1e-12 * fabs(tmax);
                if (tscale < 1e-200) tscale = 1e-12;
                if (fabs(W->t - tmax) < tscale) break;
                wh_synchronize(W);
                W->dt = tmax - W->t;
            } else {
                last_step = 1;
                wh_synchronize(W);
                if (W->dt_last_done != 0.0) last_full_dt = W->dt_last_done;
                W->dt = tmax - W->t;
            }
        }
        if (W->max_steps > 0 && W->n_steps >= W->max_steps) { W->gave_up = 1; break; }
        wh_step(W);
    }
    wh_synchronize(W);
    W->dt = last_full_dt;
}

/* integrate() with integrator='whfast': series[n_outputs][3 + 10*Np]; dt0 = REBOUND's default 1e-3 unless > 0;
 * returns the number of steps taken */
long nbo_integrate_series_whfast(const double* params, int n_bodies, double n_days, int n_outputs, double dt0,
                                 int corrector, double* series) {
    sys_t s;
    build_system(params, n_bodies, &s);
    whfast_t W;
    wh_from_inertial(&W, &s, corrector);
    if (dt0 > 0.0) W.dt = dt0;
    const int Wd = 3 + 10 * (n_bodies - 1);
    const double step = n_days / (double)(n_outputs - 1);
    W.max_steps = (long)(4.0 * (n_days / W.dt + n_outputs)) + 1000;
    for (int o = 0; o < n_outputs; o++) {
        const double t = (o == n_outputs - 1) ? n_days : o * step;
        if (!W.gave_up) wh_integrate_to(&W, t);
        if (W.gave_up) { for (int k = 0; k < Wd; k++) series[(size_t)o * Wd + k] = NAN; continue; }
        wh_to_inertial(&W, &s);
        sample_row(&s, series + (size_t)o * Wd);
    }
    return W.n_steps;
}

/* ------------------------------------------------------------------------------------------ */
/* Jacobi osculating elements as rebound's particle.P/a/e/inc/Omega/omega/M return them [3P]:  */
/* primary = centre of mass of all lower-index bodies, mu = G (M_interior + m).                */
/* nbody/simulation.py:148-150.                                                                */
/* ------------------------------------------------------------------------------------------ */
static double acos2(double num, double den, double disamb) {
    double val, c = num / den;
    if (c > -1.0 && c < 1.0) val = acos(c);
    else val = (c <= -1.0) ? PI : 0.0;
    return (disamb < 0.0) ? -val : val;
}

typedef struct { double P, a, e, inc, Omega, omega, M; } orbit_t;

/* [3P] REBOUND (reb_tools_particle_to_orbit, 3.x) ends with "move some of the angles into [0,2pi) range": f, l, M,
 * theta and omega go through reb_tools_mod2pi (fmod, + 2 pi if negative); Omega and pomega do not. */
static double mod2pi(double f) {
    f = fmod(f, 2.0 * PI);
    return (f < 0.0) ? f + 2.0 * PI : f;
}

static orbit_t state_to_orbit(double mu, const double* d, const double* w) {
    const double MIN_INC = 1e-8, MIN_ECC = 1e-8;
    orbit_t o;
    double hx = d[1] * w[2] - d[2] * w[1], hy = d[2] * w[0] - d[0] * w[2], hz = d[0] * w[1] - d[1] * w[0];
    double h = sqrt(hx * hx + hy * hy + hz * hz);
    double v2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double vc2 = mu / r;
    o.a = -mu / (v2 - 2.0 * vc2);
    double vr = (d[0] * w[0] + d[1] * w[1] + d[2] * w[2]) / r;
    double rvr = r * vr, vd2 = v2 - vc2;
    double ex = (vd2 * d[0] - rvr * w[0]) / mu, ey = (vd2 * d[1] - rvr * w[1]) / mu,
           ez = (vd2 * d[2] - rvr * w[2]) / mu;
    o.e = sqrt(ex * ex + ey * ey + ez * ez);
    double nmm = o.a / fabs(o.a) * sqrt(fabs(mu / (o.a * o.a * o.a)));
    o.P = 2.0 * PI / nmm;
    o.inc = acos2(hz, h, 1.0);
    double nx = -hy, ny = hx, nn = sqrt(nx * nx + ny * ny);
    o.Omega = acos2(nx, nn, ny);
    double ea = 0.0, f;
    if (o.e < 1.0) {
        ea = acos2(1.0 - r / o.a, o.e, vr);
        o.M = ea - o.e * sin(ea);
    } else {
        ea = acosh((1.0 - r / o.a) / o.e);
        if (vr < 0.0) ea = -ea;
        o.M = o.e * sinh(ea) - ea;
    }
    if (o.inc < MIN_INC || o.inc > PI - MIN_INC) { /* planar: longitudes measured from x */
        double theta = acos2(d[0], r, d[1]);
        double pomega = acos2(ex, o.e, ey);
        if (o.inc < PI / 2.0) { o.omega = pomega - o.Omega; f = theta - pomega; }
        else { o.omega = o.Omega - pomega; f = pomega - theta; }
        if (o.e <= MIN_ECC) { o.omega = 0.0; }
    } else {
        double wpf = acos2(nx * d[0] + ny * d[1], nn * r, d[2]);
        o.omega = acos2(nx * ex + ny * ey, nn * o.e, ez);
        f = wpf - o.omega;
        if (o.e <= MIN_ECC) { o.omega = 0.0; }
    }
    (void)f;
    o.omega = mod2pi(o.omega);
    o.M = mod2pi(o.M);
    return o;
}

/* series row layout (NBT_F_SERIES): star x,y,z, then per planet x,y,z,P,a,e,inc,Omega,omega,M */
static void sample_row(const sys_t* s, double* row) {
    int n = s->n;
    row[0] = s->x[0]; row[1] = s->x[1]; row[2] = s->x[2];
    double M = s->m[0], cx[3], cv[3];
    for (int k = 0; k < 3; k++) { cx[k] = s->x[k]; cv[k] = s->v[k]; }
    for (int i = 1; i < n; i++) {
        double d[3], w[3];
        for (int k = 0; k < 3; k++) { d[k] = s->x[3 * i + k] - cx[k]; w[k] = s->v[3 * i + k] - cv[k]; }
        orbit_t o = state_to_orbit(NBT_G * (M + s->m[i]), d, w);
        double* p = row + 3 + 10 * (i - 1);
        p[0] = s->x[3 * i]; p[1] = s->x[3 * i + 1]; p[2] = s->x[3 * i + 2];
        p[3] = o.P; p[4] = o.a; p[5] = o.e; p[6] = o.inc; p[7] = o.Omega; p[8] = o.omega; p[9] = o.M;
        for (int k = 0; k < 3; k++) {
            cx[k] = (cx[k] * M + s->x[3 * i + k] * s->m[i]) / (M + s->m[i]);
            cv[k] = (cv[k] * M + s->v[3 * i + k] * s->m[i]) / (M + s->m[i]);
        }
        M += s->m[i];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* exported fine-grained pieces (used by tests to pin each stage separately)                   */
/* ------------------------------------------------------------------------------------------ */

/* state[n][7] = m, x, y, z, vx, vy, vz after add()+move_to_com() */
void nbo_initial_state(const double* params, int n_bodies, double* state) {
    sys_t s;
    build_system(params, n_bodies, &s);
    for (int i = 0; i < n_bodies; i++) {
        state[7 * i] = s.m[i];
        for (int k = 0; k < 3; k++) { state[7 * i + 1 + k] = s.x[3 * i + k]; state[7 * i + 4 + k] = s.v[3 * i + k]; }
    }
}

/* integrate(): series[n_outputs][3 + 10*Np]; returns number of IAS15 steps taken */
long nbo_integrate_series(const double* params, int n_bodies, double n_days, int n_outputs,
                          double* series, double* final_state) {
    sys_t s;
    build_system(params, n_bodies, &s);
    ias15_t* I = (ias15_t*)malloc(sizeof(ias15_t));
    ias15_init(I, &s);
    /* step budget: a pair on a collision course drives the adaptive step towards zero; REBOUND crawls through it (it has
     * no budget), so the allowance is generous — 200 steps per output where a regular system takes one or two — before
     * the run is given up and marked non-finite.  (The product's IAS15 kernel gives up much earlier, nbt_ias15.cuh.) */
    I->max_steps = 200L * n_outputs + 100000L;
    int W = 3 + 10 * (n_bodies - 1);
    double step = n_days / (double)(n_outputs - 1); /* np.linspace(0, Ndays, Noutputs) */
    for (int o = 0; o < n_outputs; o++) {
        double t = (o == n_outputs - 1) ? n_days : o * step;
        if (!I->gave_up) ias15_integrate_to(I, t);
        if (I->gave_up) { for (int k = 0; k < W; k++) series[(size_t)o * W + k] = NAN; continue; }
        sample_row(&I->s, series + (size_t)o * W);
    }
    if (final_state)
        for (int i = 0; i < n_bodies; i++)
            for (int k = 0; k < 3; k++) {
                final_state[6 * i + k] = I->s.x[3 * i + k];
                final_state[6 * i + 3 + k] = I->s.v[3 * i + k];
            }
    long ns = I->n_steps;
    free(I);
    return ns;
}

/* TRUE mid-transit times of one planet (0-based index j among the planets): the roots of x_planet(t) - x_star(t)
 * with falling sign on the IAS15 trajectory itself, bracketed on the same output grid as transit_times() and
 * refined by a safeguarded Newton iteration that re-integrates from the bracket's left sample to every trial
 * time.  This is what NBT_TT_REFINED approximates (north_star item 2); the reference itself never computes it.
 * Returns the number of crossings (out holds at most cap of them). */
int nbo_refined_transit_times(const double* params, int n_bodies, double n_days, int n_outputs, int j,
                              double* out, int cap) {
    sys_t s;
    build_system(params, n_bodies, &s);
    ias15_t* I = (ias15_t*)malloc(sizeof(ias15_t));
    ias15_t* P = (ias15_t*)malloc(sizeof(ias15_t));
    ias15_t* Q = (ias15_t*)malloc(sizeof(ias15_t));
    ias15_init(I, &s);
    I->max_steps = 400L * n_outputs + 100000L;
    double step = n_days / (double)(n_outputs - 1);
    int c = 0, ip = 3 * (j + 1);
    double prev = 0.0, t_prev = 0.0;
    for (int o = 0; o < n_outputs && !I->gave_up; o++) {
        double t = (o == n_outputs - 1) ? n_days : o * step;
        if (o > 0) memcpy(P, I, sizeof(ias15_t));
        ias15_integrate_to(I, t);
        double cur = I->s.x[ip] - I->s.x[0];
        if (o > 0 && prev >= 0.0 && cur <= 0.0) {
            double lo = t_prev, hi = t, flo = prev, fhi = cur;
            double tr = lo + (hi - lo) * flo / (flo - fhi);
            if (!(tr >= lo && tr <= hi)) tr = 0.5 * (lo + hi);
            for (int it = 0; it < 60; it++) {
                memcpy(Q, P, sizeof(ias15_t));
                ias15_integrate_to(Q, tr);
                double f = Q->s.x[ip] - Q->s.x[0], fp = Q->s.v[ip] - Q->s.v[0];
                if (f > 0.0) { lo = tr; flo = f; } else { hi = tr; fhi = f; }
                double tn = tr - f / fp;
                if (!(tn > lo && tn < hi)) tn = 0.5 * (lo + hi);
                double dtn = fabs(tn - tr);
                tr = tn;
                if (dtn < 1e-13 || f == 0.0) break;
            }
            if (c < cap) out[c] = tr;
            c++;
        }
        prev = cur; t_prev = t;
    }
    free(I); free(P); free(Q);
    return c;
}

/* transit_times(xp, xs, times): simulation.py:163-170 with find_zero of tools.py:61-65 */
int nbo_transit_times(const double* xp, const double* xs, const double* times, int n, double* tt, int cap) {
    int c = 0;
    double prev = xp[0] - xs[0];
    for (int i = 1; i < n; i++) {
        double cur = xp[i] - xs[i];
        if (prev >= 0.0 && cur <= 0.0) {
            double m = (cur - prev) / (times[i] - times[i - 1]);
            double t0 = -prev / m + times[i - 1];
            if (c < cap) tt[c] = t0;
            c++;
        }
        prev = cur;
    }
    return c;
}

/* TTV(epochs = arange(n), tt): simulation.py:172-177.  out: ttv[n]; returns slope m, intercept b */
void nbo_ttv_fit(const double* tt, int n, double* ttv, double* m_out, double* b_out) {
    /* least squares of tt on k = 0..n-1 (what np.linalg.lstsq solves), centred for stability */
    double kbar = 0.5 * (n - 1), tbar = 0;
    for (int k = 0; k < n; k++) tbar += tt[k];
    tbar /= n;
    double skk = 0, skt = 0;
    for (int k = 0; k < n; k++) { double dk = k - kbar; skk += dk * dk; skt += dk * (tt[k] - tbar); }
    double m = (n > 1) ? skt / skk : 0.0;
    double b = tbar - m * kbar;
    for (int k = 0; k < n; k++) ttv[k] = tt[k] - m * k - b;
    *m_out = m; *b_out = b;
}

static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* maxavg(x) = (percentile(|x|, 75) + max|x|) / 2, numpy 'linear' percentile.  tools.py:22 */
double nbo_maxavg(const double* x, int n) {
    if (n <= 0) return NAN;
    double* s = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) s[i] = fabs(x[i]);
    qsort(s, n, sizeof(double), cmp_double);
    double pos = 0.75 * (n - 1);
    int lo = (int)floor(pos);
    double frac = pos - lo;
    double p75 = (lo + 1 < n) ? s[lo] + (s[lo + 1] - s[lo]) * frac : s[lo];
    double r = 0.5 * (p75 + s[n - 1]);
    free(s);
    return r;
}

/* astropy autofrequency grid (tools.py:104) [3P]: df = 1.0 / baseline / samples_per_peak;
 * Nf = 1 + int(np.round((fmax - fmin) / df)) (np.round = round-half-even = rint);
 * freq = fmin + df * arange(Nf). */
double nbo_ls_df(double baseline, int spp) { return 1.0 / baseline / (double)spp; }
int64_t nbo_ls_nfreq(double baseline, double minfreq, double maxfreq, int spp) {
    double df = nbo_ls_df(baseline, spp);
    return 1 + (int64_t)rint((maxfreq - minfreq) / df);
}

/* Generalised (floating-mean) Lomb-Scargle, standard normalisation, unit weights:
 * what LombScargle(t, y).power(f, method='cython'|'slow') evaluates [3P]. */
void nbo_lomb_scargle(const double* t, const double* y, int n, double f0, double df, int64_t nf,
                      double* power) {
    double w = 1.0 / n, ybar = 0;
    for (int i = 0; i < n; i++) ybar += w * y[i];
    double YY = 0;
    for (int i = 0; i < n; i++) YY += w * (y[i] - ybar) * (y[i] - ybar);
    for (int64_t k = 0; k < nf; k++) {
        double om = 2.0 * PI * (f0 + k * df);
        double S = 0, C = 0, S2 = 0, C2 = 0;
        for (int i = 0; i < n; i++) {
            double sn = sin(om * t[i]), cs = cos(om * t[i]);
            S += w * sn; C += w * cs;
            S2 += 2.0 * w * sn * cs; C2 += w * (cs * cs - sn * sn);
        }
        S2 -= 2.0 * S * C;
        C2 -= (C * C - S * S);
        double omtau = 0.5 * atan2(S2, C2);
        double ct = cos(omtau), st = sin(omtau);
        double YC = 0, YS = 0, Cc = 0, Ss = 0, Ct = 0, St = 0;
        for (int i = 0; i < n; i++) {
            double sn = sin(om * t[i]), cs = cos(om * t[i]);
            double cst = cs * ct + sn * st, snt = sn * ct - cs * st; /* cos, sin of om(t - tau) */
            double yc = y[i] - ybar;
            YC += w * yc * cst; YS += w * yc * snt;
            Cc += w * cst * cst; Ss += w * snt * snt;
            Ct += w * cst; St += w * snt;
        }
        Cc -= Ct * Ct; Ss -= St * St;
        /* a vanishing normalisation (sine term at the Nyquist frequency of integer epochs) is
         * round-off / round-off in astropy; both oracle and product drop that term instead */
        power[k] = ((Cc > 1e-14 ? YC * YC / Cc : 0.0) + (Ss > 1e-14 ? YS * YS / Ss : 0.0)) / YY;
    }
}

/* nested.py:85-95 myloglike for one forward model */
static double loss_fn(int loss, double z) { /* nested.py:79-82 */
    return loss == NBT_LOSS_SOFT_L1 ? 2.0 * (sqrt(1.0 + z) - 1.0) : z;
}

double nbo_loglike_loss(const double* ttv, int n_ttv, double c0, double c1, const double* x,
                        const double* y, const double* yerr, int n_data, int loss) {
    int ok = 1;
    for (int i = 0; i < n_data; i++) if ((int)x[i] >= n_ttv || (int)x[i] < -n_ttv) ok = 0;
    double s = 0;
    if (ok) {
        for (int i = 0; i < n_data; i++) {
            int k = (int)x[i]; if (k < 0) k += n_ttv;
            double r = (y[i] - (c1 * x[i] + c0) - ttv[k]) / yerr[i];
            s += loss_fn(loss, r * r);
        }
        return -0.5 * s;
    }
    for (int i = 0; i < n_ttv && i < n_data; i++) {
        double r = (y[i] - (c1 * x[i] + c0) - ttv[i]) / yerr[i];
        s += loss_fn(loss, r * r);
    }
    return -10.0 * s;
}

double nbo_loglike(const double* ttv, int n_ttv, double c0, double c1, const double* x,
                   const double* y, const double* yerr, int n_data) {
    int ok = 1;
    for (int i = 0; i < n_data; i++) if ((int)x[i] >= n_ttv || (int)x[i] < -n_ttv) ok = 0;
    if (ok) {
        double s = 0;
        for (int i = 0; i < n_data; i++) {
            int k = (int)x[i]; if (k < 0) k += n_ttv;
            double r = (y[i] - (c1 * x[i] + c0) - ttv[k]) / yerr[i];
            s += r * r;
        }
        return -0.5 * s;
    }
    /* except-branch: -10 * sum over the ttv that exist, paired index-for-index with the data */
    double s = 0;
    for (int i = 0; i < n_ttv && i < n_data; i++) {
        double r = (y[i] - (c1 * x[i] + c0) - ttv[i]) / yerr[i];
        s += r * r;
    }
    /* (the reference raises IndexError inside the loop too if n_ttv > n_data; it never is) */
    return -10.0 * s;
}

/* ------------------------------------------------------------------------------------------ */
/* Batched driver with the product's output layout (include/nbt.h), for parity tests and as   */
/* the timed CPU baseline.  Pointers are host memory.  Returns 0.                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const nbt_config* cfg; const nbt_inputs* in; nbt_outputs* out;
    atomic_int next; atomic_llong steps;
} run_ctx;

static void run_one(const nbt_config* cfg, const nbt_inputs* in, nbt_outputs* out, int b, int64_t* steps) {
    const int B = cfg->n_systems, N = cfg->n_bodies, Np = N - 1;
    /* per-system grid (simulations/generate_simulations.py:45-50) or the uniform one */
    const int NO = in->n_outputs_sys ? in->n_outputs_sys[b] : cfg->n_outputs;
    const double n_days = in->n_days_sys ? in->n_days_sys[b] : cfg->n_days;
    const int W = 3 + 10 * Np;
    const int stride = cfg->orbit_stride > 0 ? cfg->orbit_stride : 4;
    int64_t total_steps = 0;
    {
        const double* par = in->params + (size_t)b * N * NBT_NPAR;
        double* series = (double*)malloc(sizeof(double) * (size_t)NO * W);
        double* times = (double*)malloc(sizeof(double) * NO);
        double* col_p = (double*)malloc(sizeof(double) * NO);
        double* col_s = (double*)malloc(sizeof(double) * NO);
        if (cfg->scheme == NBT_SCHEME_WHFAST)      /* generate(..., integrator='whfast'), nbody/simulation.py:39-42 */
            total_steps += nbo_integrate_series_whfast(par, N, n_days, NO, 0.0, 5, series);
        else
            total_steps += nbo_integrate_series(par, N, n_days, NO, series, NULL);
        double step = n_days / (double)(NO - 1);
        for (int o = 0; o < NO; o++) times[o] = (o == NO - 1) ? n_days : o * step;
        int status = 0;
        for (int o = 0; o < NO; o++) {
            col_s[o] = series[(size_t)o * W];
            for (int k = 0; k < W; k++) if (!isfinite(series[(size_t)o * W + k])) status |= NBT_S_NONFINITE;
        }
        /* RV = np.diff(star x) * 1.496e11 * dt, dt = Noutputs/(Ndays*86400)  (simulation.py:157,186) */
        if ((cfg->flags & NBT_F_RV) && out->rv) {
            double dtc = (double)NO / (n_days * 24.0 * 60.0 * 60.0);
            double* rvs = (double*)malloc(sizeof(double) * (NO - 1));
            for (int o = 0; o + 1 < NO; o++) {
                rvs[o] = (col_s[o + 1] - col_s[o]) * 1.496e11 * dtc;
                out->rv[(size_t)o * B + b] = rvs[o];
            }
            if (out->rv_max) out->rv_max[b] = nbo_maxavg(rvs, NO - 1);
            if ((cfg->flags & NBT_F_LS_RV) && out->ls_rv_power) {
                double T = times[NO - 1] - times[1];
                int64_t nf = nbo_ls_nfreq(T, cfg->ls_rv_minfreq, cfg->ls_rv_maxfreq, cfg->ls_samples_per_peak);
                nbo_lomb_scargle(times + 1, rvs, NO - 1, cfg->ls_rv_minfreq,
                                 nbo_ls_df(T, cfg->ls_samples_per_peak), nf, out->ls_rv_power + (size_t)b * nf);
            }
            free(rvs);
        }
        for (int j = 0; j < Np; j++) {
            size_t sp = (size_t)b * Np + j;
            double* row0 = series + 3 + 10 * j;
            if (cfg->flags & NBT_F_MEANS) {
                double sa = 0, se = 0, si = 0, so = 0;
                for (int o = 0; o < NO; o++) {
                    const double* p = row0 + (size_t)o * W;
                    sa += p[4]; se += p[5]; si += p[6]; so += p[8];
                    if (p[4] < 0) status |= NBT_S_UNBOUND;
                }
                out->mean_a[sp] = sa / NO; out->mean_e[sp] = se / NO; out->mean_inc[sp] = si / NO;
                if ((cfg->flags & NBT_F_MEAN_OMEGA) && out->mean_omega) out->mean_omega[sp] = so / NO;
            }
            if ((cfg->flags & NBT_F_ORBIT) && out->orbit_xz)
                for (int o = 0, q = 0; o < NO; o += stride, q++) {
                    const double* p = row0 + (size_t)o * W;
                    double* dst = out->orbit_xz + (((size_t)q * B + b) * Np + j) * 2;
                    dst[0] = p[0]; dst[1] = p[2];
                }
            if ((cfg->flags & NBT_F_PLANET1_ONLY) && j > 0) {
                out->n_tt[sp] = 0; out->period[sp] = par[(j + 1) * NBT_NPAR + NBT_P_P]; out->t0[sp] = 0;
                out->ttv_max[sp] = 0;
                continue;
            }
            for (int o = 0; o < NO; o++) col_p[o] = row0[(size_t)o * W];
            int64_t off = in->tt_offsets[sp], cap = in->tt_offsets[sp + 1] - off;
            int n = nbo_transit_times(col_p, col_s, times, NO, out->tt + off, (int)cap);
            out->n_tt[sp] = n;
            if (n > cap) { status |= NBT_S_TT_OVERFLOW; n = (int)cap; }
            if (n >= 3) {
                double m, c;
                nbo_ttv_fit(out->tt + off, n, out->ttv + off, &m, &c);
                out->period[sp] = m; out->t0[sp] = c;
                out->ttv_max[sp] = nbo_maxavg(out->ttv + off, n);
            } else { /* simulation.py:207-210 fallback: per = objects P, t0 = 0, ttv = [0] */
                status |= NBT_S_FEW_TRANSITS;
                out->period[sp] = par[(j + 1) * NBT_NPAR + NBT_P_P]; out->t0[sp] = 0.0;
                for (int k = 0; k < (n > 1 ? n : 1) && k < cap; k++) out->ttv[off + k] = 0.0;
                out->ttv_max[sp] = 0.0;
            }
            if ((cfg->flags & NBT_F_LS_OC) && out->ls_oc_power && in->ls_offsets) {
                int64_t loff = in->ls_offsets[sp], lcap = in->ls_offsets[sp + 1] - loff;
                int64_t nf = 0;
                if (n >= 3) {
                    double T = (double)(n - 1);
                    nf = nbo_ls_nfreq(T, cfg->ls_oc_minfreq, cfg->ls_oc_maxfreq, cfg->ls_samples_per_peak);
                    if (nf > lcap) nf = lcap;
                    double* ep = (double*)malloc(sizeof(double) * n);
                    for (int k = 0; k < n; k++) ep[k] = k;
                    nbo_lomb_scargle(ep, out->ttv + off, n, cfg->ls_oc_minfreq,
                                     nbo_ls_df(T, cfg->ls_samples_per_peak), nf, out->ls_oc_power + loff);
                    free(ep);
                }
                out->ls_oc_nfreq[sp] = (int32_t)nf;
            }
        }
        if ((cfg->flags & NBT_F_SERIES) && out->series)
            for (int o = 0; o < NO; o++)
                memcpy(out->series + ((size_t)o * B + b) * W, series + (size_t)o * W, sizeof(double) * W);
        out->status[b] = status;
        free(series); free(times); free(col_p); free(col_s);
    }
    *steps = total_steps;
}

static void* run_worker(void* arg) {
    run_ctx* c = (run_ctx*)arg;
    for (;;) {
        int b = atomic_fetch_add(&c->next, 1);
        if (b >= c->cfg->n_systems) break;
        int64_t st = 0;
        run_one(c->cfg, c->in, c->out, b, &st);
        atomic_fetch_add(&c->steps, (long long)st);
    }
    return NULL;
}

int nbo_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int nbo_run(const nbt_config* cfg, const nbt_inputs* in, nbt_outputs* out, int n_threads,
            int64_t* ias15_steps_out) {
    init_tables();
    run_ctx c;
    c.cfg = cfg; c.in = in; c.out = out;
    atomic_init(&c.next, 0); atomic_init(&c.steps, 0);
    if (n_threads <= 0) n_threads = nbo_max_threads();
    if (n_threads > cfg->n_systems) n_threads = cfg->n_systems > 0 ? cfg->n_systems : 1;
    if (n_threads <= 1) run_worker(&c);
    else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
        for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, run_worker, &c);
        for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
        free(th);
    }
    if (ias15_steps_out) *ias15_steps_out = (int64_t)atomic_load(&c.steps);
    return 0;
}
