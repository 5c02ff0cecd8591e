/*
 * nbt.h — C ABI of the B200-native batched N-body TTV engine.
 *
 * This is the drop-in boundary for the reference's generate()/integrate()/analyze()
 * path.  The reference (pearsonkyle/Nbody-AI) has no FFI of its own: its boundary is the
 * Python function surface of nbody/simulation.py, below which it calls REBOUND through
 * ctypes.  Each entry point below names the reference interface it replaces.
 *
 *   nbt_run            <- nbody/simulation.py:27-52   generate()   (setup, Jacobi add, move_to_com)
 *                         nbody/simulation.py:131-160 integrate()  (sampling on linspace grid)
 *                         nbody/simulation.py:163-177 transit_times() + TTV()
 *                         nbody/simulation.py:179-231 analyze()    (means, O-C, maxavg, RV, periodograms)
 *                         nbody/nested.py:44-47       get_ttv()    (ttvfast forward model)
 *   nbt_lomb_scargle   <- nbody/tools.py:97-104       lomb_scargle() -> astropy autopower
 *   nbt_chi2           <- nbody/nested.py:85-95       myloglike()  (chi^2 of the forward model)
 *   nbt_ttvfit_loglike <- ttv_fit.py:58-94            nbody_fitter.loglike()
 *   nbt_run + inputs.like (NBT_F_LOGLIKE): either likelihood fused behind the forward model on the launch
 *                         stream — only loglike[B] and status[B] have to leave the device
 *   inputs.n_days_sys / n_outputs_sys <- simulations/generate_simulations.py:45-50: every draw has its own grid
 *                         (Ndays = P1 * periods, Noutputs = round(P1 * periods * 24)), one launch for all of them
 *   nbt_plan_*         <- no reference analogue (the reference grows Python lists); they size
 *                         the caller-owned CSR buffers the kernels write into.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  Every buffer is caller-owned.
 *   - units day / AU / Msun, G = NBT_G (nbody/tools.py:18).
 *   - all functions return 0 on success, <0 for an invalid argument, >0 for a CUDA error
 *     (the cudaError_t value); nbt_last_error() gives the thread-local message.
 *   - pointers in nbt_inputs / nbt_outputs are HOST pointers unless cfg.flags has
 *     NBT_F_DEVICE_PTRS, in which case all of them are device pointers on cfg.device
 *     and the call only enqueues work on `stream` (no copies, no synchronisation).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef NBT_H
#define NBT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBT_ABI_VERSION 2
#define NBT_G 0.00029591220828559104 /* day, AU, Msun — nbody/tools.py:18 */
#define NBT_MAX_BODIES 9             /* star + 8 planets (solarsystem_nbody.py) */

/* columns of the per-body parameter row (what the reference passes to rebound.add(**obj)) */
enum { NBT_P_M = 0, NBT_P_P = 1, NBT_P_E = 2, NBT_P_INC = 3, NBT_P_OMEGA_NODE = 4,
       NBT_P_OMEGA_PERI = 5, NBT_P_F = 6, NBT_NPAR = 7 };

/* integrator scheme: symmetric compositions of the Wisdom-Holman (Jacobi) map */
enum { NBT_SCHEME_WH2 = 0,   /* kick-drift-kick leapfrog */
       NBT_SCHEME_Y4 = 1,    /* Yoshida 4th-order triple jump */
       NBT_SCHEME_SABA4 = 2, /* Laskar-Robutel SBAB/SABA-class, 4 stages */
       NBT_SCHEME_M64 = 3,   /* order (6,4) 4-stage composition */
       NBT_SCHEME_C4 = 4,    /* 2-stage force-gradient composition, O(eps dt^4 + eps^2 dt^4) */
       NBT_SCHEME_SBABC3 = 5, /* 3-stage Lobatto composition + force-gradient end kicks, order (6,4):
                                 the accuracy of M64 with three Kepler drifts per step (the default) */
       NBT_SCHEME_COUNT = 6,
       NBT_SCHEME_AUTO = 6,  /* SBABC3, except that systems resolved finely enough by the output grid itself
                                 (P_min >= 250 output steps) take one cheaper C4 step per output:
                                 nbt_plan_substeps marks them with negative counts in inputs.substeps,
                                 nbt_plan_order sorts them last and counts them for cfg.alt_systems.  A
                                 negative count outside those slots runs |count| SBABC3 steps. */
       NBT_SCHEME_IAS15 = 7, /* adaptive 15th-order Gauss-Radau (Rein & Spiegel 2015), one system per thread — what
                                 the reference itself integrates with (REBOUND's default, nbody/simulation.py:31).
                                 Far slower than the compositions and latency-bound; it is there for the few
                                 systems no fixed step resolves (close encounters, error-amplifying pairs):
                                 engine.run_verified routes them here.  substeps / order / alt_systems /
                                 lanes_systems are ignored. */
       NBT_SCHEME_WHFAST = 8 }; /* integrator='whfast' of nbody/simulation.py:39-42: REBOUND's WHFast (drift-kick-drift
                                 in Jacobi coordinates) at its default dt = 1e-3 d with the 5th-order symplectic
                                 corrector applied at the start and undone at the end of every sim.integrate(t)
                                 (safe_mode = 0), last step of an interval shortened to land on the output
                                 (exact_finish_time).  cfg.substeps = full steps per output interval is derived. */

/* transit-time mode */
enum { NBT_TT_GRID_LINEAR = 0, /* reference semantics: sign test + linear interpolation on the
                                  output grid, nbody/simulation.py:163-170, tools.py:61-65 */
       NBT_TT_REFINED = 1 };   /* Newton-refined root of x_p - x_s on the integrator's own step */

/* cfg.flags */
enum { NBT_F_MEANS = 1 << 0,       /* time means of Jacobi a, e, inc (analyze :216-217) */
       NBT_F_MEAN_OMEGA = 1 << 1,  /* also the mean of omega */
       NBT_F_RV = 1 << 2,          /* store RV signal [n_outputs-1][B] + rv_max */
       NBT_F_ORBIT = 1 << 3,       /* store x[::4], z[::4] per planet */
       NBT_F_SERIES = 1 << 4,      /* store the full sim_data series (generate() compatibility) */
       NBT_F_LS_OC = 1 << 5,       /* O-C periodogram per planet */
       NBT_F_LS_RV = 1 << 6,       /* RV periodogram per system (needs NBT_F_RV) */
       NBT_F_PLANET1_ONLY = 1 << 7,/* ttvfast: transit products for planet 1 only (analyze :181-184) */
       NBT_F_DEVICE_PTRS = 1 << 8, /* all buffers are device pointers; async on `stream` */
       NBT_F_TIMING = 1 << 9,      /* record CUDA events around each stage (read with nbt_timing) */
       NBT_F_LOGLIKE = 1 << 10,    /* evaluate inputs.like behind the forward model, on the same stream: outputs.loglike */
       NBT_F_LS_PEAKS = 1 << 11    /* reduce each O-C periodogram (NBT_F_LS_OC) to its NBT_LS_NPEAKS highest local maxima
                                      on the device: outputs.ls_oc_peaks; with ls_oc_power == NULL (host pointers) the
                                      full power arrays never leave the device */
};
#define NBT_LS_NPEAKS 3

/* per-system status bits (the reference signals these only indirectly) */
enum { NBT_S_NONFINITE = 1 << 0, NBT_S_UNBOUND = 1 << 1, NBT_S_KEPLER = 1 << 2,
       NBT_S_TT_OVERFLOW = 1 << 3, NBT_S_FEW_TRANSITS = 1 << 4,
       NBT_S_UNCALIBRATED = 1 << 5, /* outside the envelope the sub-step rule was calibrated on (a planet pair closer
                                       than 4 mutual Hill radii, e > 0.1, or m_planet / m_star > 0.02): the fixed-step
                                       result may miss the parity tolerances; engine.run_verified checks such systems
                                       by step doubling and re-integrates the ones that need it */
       NBT_S_UNRESOLVED = 1 << 6 }; /* set by engine.run_verified: step doubling did not converge and no fallback ran */

typedef struct nbt_config {
    int32_t struct_size; /* = sizeof(nbt_config), ABI guard */
    int32_t n_systems;   /* B */
    int32_t n_bodies;    /* star + planets, 2..NBT_MAX_BODIES */
    int32_t n_outputs;   /* Noutputs of np.linspace(0, Ndays, Noutputs) (the largest, with inputs.n_outputs_sys) */
    double n_days;       /* Ndays (the largest, with inputs.n_days_sys) */
    int32_t substeps;    /* composition steps per output interval; <=0: use inputs.substeps[] */
    int32_t scheme;      /* NBT_SCHEME_* */
    int32_t tt_mode;     /* NBT_TT_* */
    int32_t flags;       /* NBT_F_* */
    int32_t device;      /* CUDA device ordinal */
    int32_t orbit_stride;/* 4 in the reference (analyze :227-228) */
    double ls_oc_minfreq, ls_oc_maxfreq; /* 1/180, 1/2 (analyze :213) */
    double ls_rv_minfreq, ls_rv_maxfreq; /* 1/365, 1/2 (tools.py:97) */
    int32_t ls_samples_per_peak;         /* 10 (tools.py:104) */
    int32_t tt_cap_max;  /* largest per-planet CSR capacity (nbt_plan_tt returns it); 0 = derive
                            from host tt_offsets (required > 0 with NBT_F_DEVICE_PTRS) */
    int32_t lanes_systems; /* leading launch-order slots integrated by the one-planet-per-lane
                            kernel (nbt_plan_lanes); 0 = none, n_systems = all */
    int32_t alt_systems; /* NBT_SCHEME_AUTO: trailing launch-order slots that take C4 steps (nbt_plan_order
                            returns it; every one must hold a negative count); 0 = all SBABC3 */
} nbt_config;

/* Likelihood evaluated behind the forward model (NBT_F_LOGLIKE).  All arrays follow the pointer convention of the
 * call (host, or device with NBT_F_DEVICE_PTRS). */
enum { NBT_LIKE_NESTED = 0,   /* nbody/nested.py:85-95 myloglike: planet-1 O-C against (x, y, yerr), linear term c1*x + c0 */
       NBT_LIKE_TTVFIT = 1 }; /* ttv_fit.py:58-94: transit times of every planet with data, shifted and offset (see nbt_ttvfit_loglike) */
enum { NBT_LOSS_LINEAR = 0,   /* z */
       NBT_LOSS_SOFT_L1 = 1 };/* 2 (sqrt(1 + z) - 1)   (nbody/nested.py:79-82) */
typedef struct nbt_like {
    int32_t kind;            /* NBT_LIKE_* */
    int32_t loss;            /* NBT_LOSS_* (NESTED only) */
    int32_t n_data;          /* observed epochs (NESTED) / observed orbit numbers n_obs (TTVFIT) */
    int32_t reserved;
    const double* c0;        /* NESTED [B] intercept cube[0] */
    const double* c1;        /* NESTED [B] slope cube[1] */
    const double* x;         /* NESTED [n_data] epochs */
    const double* y;         /* NESTED [n_data] mid-transit times */
    const double* yerr;      /* NESTED [n_data] */
    const int32_t* has_data; /* TTVFIT [Np] */
    const double* tc;        /* TTVFIT [Np][n_data] */
    const double* tc_err;    /* TTVFIT [Np][n_data] */
    const int32_t* orbit;    /* TTVFIT [n_data] orbit numbers (ttv_fit.py:38-39) */
} nbt_like;

typedef struct nbt_inputs {
    const double* params;      /* [B][n_bodies][NBT_NPAR] */
    const int32_t* substeps;   /* [B] or NULL (then cfg.substeps applies to all) */
    const int32_t* order;      /* [B] launch order (nbt_plan_order: heavy systems first, equal trip
                                  counts within a warp, unstable systems grouped) or NULL */
    const int64_t* tt_offsets; /* [B*Np + 1] CSR capacity offsets into tt/ttv (nbt_plan_tt) */
    const int64_t* ls_offsets; /* [B*Np + 1] CSR capacity offsets into ls_oc_power, or NULL */
    /* per-system output grids np.linspace(0, n_days_sys[b], n_outputs_sys[b]) — both or neither; NULL = the uniform
     * grid of cfg.  Not combinable with the [n_outputs][B] products (NBT_F_RV, _ORBIT, _SERIES, _LS_RV). */
    const double* n_days_sys;     /* [B] or NULL */
    const int32_t* n_outputs_sys; /* [B] or NULL */
    const nbt_like* like;         /* NBT_F_LOGLIKE: always a HOST pointer to the struct (its arrays follow the call's convention) */
} nbt_inputs;

typedef struct nbt_outputs {
    /* transit products, CSR over (system, planet).  With host pointers any of tt, ttv, n_tt, period, t0, ttv_max may
     * be NULL: it is still computed on the device (the fit, the periodogram and the likelihood read it) but not
     * copied back. */
    double* tt;        /* mid-transit times [day] */
    double* ttv;       /* O-C [day] */
    int32_t* n_tt;     /* [B*Np] transits found (may exceed capacity -> NBT_S_TT_OVERFLOW) */
    double* period;    /* [B*Np] OLS slope, or objects P when n_tt < 3 */
    double* t0;        /* [B*Np] OLS intercept, or 0 */
    double* ttv_max;   /* [B*Np] maxavg(ttv) [day] (tools.py:22) */
    /* orbit means over all outputs (NBT_F_MEANS) */
    double* mean_a;    /* [B*Np] */
    double* mean_e;
    double* mean_inc;
    double* mean_omega;/* NBT_F_MEAN_OMEGA */
    int32_t* status;   /* [B] NBT_S_* */
    /* optional series, all time-major so that a warp's stores coalesce */
    double* rv;        /* [n_outputs-1][B]   NBT_F_RV */
    double* rv_max;    /* [B]                NBT_F_RV */
    double* orbit_xz;  /* [ceil(n_outputs/stride)][B][Np][2]  NBT_F_ORBIT */
    double* series;    /* [n_outputs][B][3 + 10*Np]  NBT_F_SERIES: star x,y,z then per planet
                          x,y,z,P,a,e,inc,Omega,omega,M (simulation.py:13-25) */
    /* periodograms */
    double* ls_oc_power; /* CSR by ls_offsets     NBT_F_LS_OC */
    int32_t* ls_oc_nfreq;/* [B*Np] */
    double* ls_rv_power; /* [B][nbt_ls_nfreq(n_outputs-1 samples)]  NBT_F_LS_RV */
    double* ls_oc_peaks; /* [B*Np][NBT_LS_NPEAKS][2] (frequency [1/epoch], power), highest first, NaN-padded   NBT_F_LS_PEAKS */
    double* loglike;     /* [B]   NBT_F_LOGLIKE */
} nbt_outputs;

int nbt_abi_version(void);
int nbt_device_count(void);
const char* nbt_last_error(void);

/* Number of frequencies of astropy's autofrequency grid for a series of n samples spanning
 * `baseline` (= t.max - t.min): df = 1/baseline/spp, Nf = 1 + rint((fmax - fmin)/df)
 * (np.round is round-half-even), freq_k = fmin + k*df.  tools.py:104. */
int64_t nbt_ls_nfreq(double baseline, double minfreq, double maxfreq, int samples_per_peak);

/* Host planning helpers.  `in` supplies params (and, when set, n_days_sys / n_outputs_sys, substeps, order); the other
 * fields of `in` are ignored.
 *
 * nbt_plan_tt: fill tt_offsets[B*Np+1] with CSR capacities min(ceil(1.05*Ndays/P)+3, Noutputs) per planet.
 * Returns the largest single capacity (>0), or <0 on error. */
int nbt_plan_tt(const nbt_config* cfg, const nbt_inputs* in, int64_t* tt_offsets);
/* Fill order[B]: permutation of 0..B-1 sorted by decreasing substeps (then decreasing n_outputs_sys), Hill-unstable
 * systems (mutual separation < 4 R_Hill, from params) first within equal substeps — spread evenly over the
 * group instead where substeps >= 2.
 * in->params or in->substeps may be NULL (that key is skipped).  Negative counts (NBT_SCHEME_AUTO) sort last;
 * returns how many there are (>= 0), which the caller stores in cfg.alt_systems, or < 0 on error. */
int nbt_plan_order(const nbt_config* cfg, const nbt_inputs* in, int32_t* order);
/* Number of leading launch-order slots for the one-planet-per-lane kernel; the caller stores it in
 * cfg.lanes_systems. */
int nbt_plan_lanes(const nbt_config* cfg, const nbt_inputs* in);
/* Fill ls_offsets[B*Np+1] from tt_offsets (capacity of the O-C periodogram). */
int nbt_plan_ls(const nbt_config* cfg, const int64_t* tt_offsets, int64_t* ls_offsets);

/* Recommended composition steps per output interval for each system: ceil(step * R / P_min) with the
 * scheme's R (57.5 for SBABC3 / M64).  NBT_SCHEME_AUTO: -1 for the systems that take one C4 step
 * per output instead (P_min >= 250 output steps, Hill separation >= 4). */
int nbt_plan_substeps(const nbt_config* cfg, const nbt_inputs* in, int32_t* substeps);

/* Number of system-steps (kick+drift stages) nbt_run will execute for this plan (in->substeps, grids). */
int64_t nbt_count_steps(const nbt_config* cfg, const nbt_inputs* in);

/* The whole plan in one call, on n_threads host threads (0 = all): what nbt_plan_substeps + nbt_plan_order +
 * nbt_plan_lanes + nbt_plan_tt + nbt_plan_ls + nbt_count_steps produce, bit for bit, in one pass over params.  Callers
 * whose parameters change from call to call (samplers: nbody/nested.py:85, ttv_fit.py:47; the training-set loop,
 * simulations/generate_simulations.py:41-110) plan every call anew, so planning time is end-to-end time.
 *   sub-steps: cfg->substeps > 0 = uniform; else in->substeps if set; else computed into substeps_out[B].
 *   sort_mode 1: order[B] = nbt_plan_order's launch order (when B > 32 or C4-marked systems exist); 0: index order,
 *   except that C4-marked systems (NBT_SCHEME_AUTO) still move to the end.  order may be NULL with sort_mode 0 only if
 *   the caller accepts index order (then marked systems run SBABC3 steps).
 *   ls_offsets: NULL unless NBT_F_LS_OC.
 *   summary[5] = { system-steps of the plan, largest CSR capacity (cfg.tt_cap_max), cfg.alt_systems,
 *                  1 if order[] is the identity (the caller may pass inputs.order = NULL), cfg.lanes_systems }. */
int nbt_plan(const nbt_config* cfg, const nbt_inputs* in, int32_t sort_mode, int32_t* substeps_out, int32_t* order,
             int64_t* tt_offsets, int64_t* ls_offsets, int64_t* summary, int32_t n_threads);

/* 1 if the system described by one parameter row block [n_bodies][NBT_NPAR] is outside the envelope the sub-step
 * rule was calibrated on (what nbt_run reports as NBT_S_UNCALIBRATED), else 0. */
int nbt_uncalibrated(const double* params_one, int n_bodies);

/* The hot path: set up, integrate, detect transits, fit ephemerides, reduce. */
int nbt_run(const nbt_config* cfg, const nbt_inputs* in, nbt_outputs* out, void* cuda_stream);

/* analyze() on a recorded sim_data series (nbody/simulation.py:179-231): same outputs as
 * nbt_run, but the front end scans `series` ([n_outputs][B][3+10*Np], the NBT_F_SERIES layout)
 * sampled at `times` [n_outputs] instead of integrating; `dt` is sim_data['dt'] (:157).
 * in->params supplies the objects' P (the `< 3 transits` fallback, :207-210). */
int nbt_analyze(const nbt_config* cfg, const double* series, const double* times, double dt,
                const nbt_inputs* in, nbt_outputs* out, void* cuda_stream);

/* Stand-alone batched forms of the reference's array functions (device_ptrs as below).
 * transit_times(xp, xs, times): nbody/simulation.py:163-170. xp, xs: [n_series][n]; times: [n] if
 * times_shared else [n_series][n]; tt: [n_series][cap]; n_tt: [n_series] (may exceed cap). */
int nbt_transit_times(const double* xp, const double* xs, const double* times, int64_t n_series, int64_t n,
                      double* tt, int32_t* n_tt, int64_t cap, int times_shared, int device, int device_ptrs,
                      void* cuda_stream);
/* find_zero(t1, dx1, t2, dx2), elementwise over n: nbody/tools.py:61-65. */
int nbt_find_zero(const double* t1, const double* dx1, const double* t2, const double* dx2, int64_t n,
                  double* t0, int device, int device_ptrs, void* cuda_stream);
/* TTV(epochs, tt): nbody/simulation.py:172-177. epochs (NULL -> 0..n-1), tt, ttv: [n_series][n];
 * slope (= period m), intercept (= t0 b): [n_series]. */
int nbt_ttv_fit(const double* epochs, const double* tt, int64_t n_series, int64_t n, double* ttv,
                double* slope, double* intercept, int device, int device_ptrs, void* cuda_stream);
/* maxavg(x) = (percentile(|x|,75) + max|x|)/2: nbody/tools.py:22. x: [n_series][n]; out: [n_series]. */
int nbt_maxavg(const double* x, int64_t n_series, int64_t n, double* out, int device, int device_ptrs,
               void* cuda_stream);

/* Generalised (floating-mean) Lomb-Scargle of `n_series` independent series.
 * t, y: [n_series][n] (t == NULL -> t = 0..n-1), power: [n_series][n_freq],
 * freq_k = f0 + k*df.  device_ptrs != 0: pointers are device memory. */
int nbt_lomb_scargle(const double* t, const double* y, int64_t n_series, int64_t n,
                     double f0, double df, int64_t n_freq, double* power,
                     int device, int device_ptrs, void* cuda_stream);

/* chi^2 of nested.py:85-95 for B forward models against one data vector:
 * loglike[b] = -0.5 * sum_i loss(((y_i - (c1[b]*x_i + c0[b]) - ttv[b][x_i]) / yerr_i)^2), loss = NBT_LOSS_*
 * (nested.py:79-82), with the reference's penalty branch when epoch x_i is missing. */
int nbt_chi2(const double* ttv, const int64_t* tt_offsets, const int32_t* n_tt, int32_t n_systems,
             int32_t n_planets, const double* c0, const double* c1, const double* x, const double* y,
             const double* yerr, int32_t n_data, int32_t loss, double* loglike, int device, int device_ptrs,
             void* cuda_stream);

/* Likelihood of the UltraNest fitter, ttv_fit.py:47-94, for B forward models: for every planet j with
 * has_data[j], Tc_sim = its transit times shifted so that planet 1's first transit is 0 and offset by
 * the mean residual at the observed orbit numbers orbit[n_obs] (ttv_fit.py:40);
 * loglike[b] = sum_j -0.5 sum_i ((tc[j][i] - Tc_sim[orbit[i]]) / tc_err[j][i])^2, and -1e6 for a planet
 * whose simulation lacks an observed orbit number.  tc, tc_err: [n_planets][n_obs]. */
int nbt_ttvfit_loglike(const double* tt, const int64_t* tt_offsets, const int32_t* n_tt, int32_t n_systems,
                       int32_t n_planets, const int32_t* has_data, const double* tc, const double* tc_err,
                       const int32_t* orbit, int32_t n_obs, double* loglike, int device, int device_ptrs,
                       void* cuda_stream);

/* Grid-cell post-processing of simulations/simulation_grid.py:33-69 for B systems (planet 1):
 * exact periodogram of the O-C on the autofrequency grid (minfreq, maxfreq, samples_per_peak),
 * scipy find_peaks(height=peak_height) sorted by height, top n_order (<= 8) peak periods
 * [B][n_order] (0-padded), least-squares coefficients [B][2+2*n_order] of Tc on
 * [1, n, sin(2 pi n/P_k), cos(2 pi n/P_k)...] (0-padded), softmax(ttv) = percentile(|ttv|,95)/2,
 * avg_err = mean|harmonic model - ttv|.  tt/ttv/tt_offsets/n_tt are nbt_run's outputs;
 * tt_cap_max = cfg.tt_cap_max.  Systems with < 3 transits get NaN. */
int nbt_grid_post(const double* tt, const double* ttv, const int64_t* tt_offsets, const int32_t* n_tt,
                  int32_t n_systems, int32_t n_planets, int32_t tt_cap_max, double minfreq, double maxfreq,
                  int32_t samples_per_peak, double peak_height, int32_t n_order, double* peak_periods,
                  double* coeffs, int32_t* n_peaks, double* softmax, double* avg_err, int device,
                  int device_ptrs, void* cuda_stream);

/* Device time [ms] of the stages of this thread's last nbt_run / nbt_analyze issued with
 * NBT_F_TIMING, from CUDA events on the launch stream: ms4 = {front end (k_integrate or
 * k_scan_series), k_fit, k_ls_oc, RV stages}.  Synchronises on the last event. */
int nbt_timing(double* ms4);

/* Diagnostic: measured FP64 DFMA throughput of `device` (8 independent chains per thread,
 * 8 CTAs of 256 threads per SM) — the roofline denominator bench.py reports against, since
 * MEASURED_PEAKS.json carries no FP64 figure. */
int nbt_fp64_peak(int device, int iters, double* tflops_out, double* ms_out);
/* (iters < 0 runs |iters| iterations of the variant whose three DFMA operands are all distinct
 * per-thread registers — the register-bandwidth-limited rate real code sees.) */

/* Diagnostic: dependent-issue latency [SM cycles] of a DFMA chain and of a MUFU.RSQ64H + DADD
 * chain on `device` (one warp) — explains the latency-bound regime of small batches. */
int nbt_fp64_latency(int device, double* dfma_cycles, double* rsqrt_dadd_cycles);

#ifdef __cplusplus
}
#endif
#endif /* NBT_H */
