// nbt_capi.cu — CUDA kernels (sm_100a) + the C ABI of include/nbt.h.
//
// Kernels (all FP64, no tensor cores: nothing on the path is a dense contraction)
//   k_integrate<NP>   one system per thread, state in registers; integrate + sample on the
//                     linspace grid + transit detection + Jacobi-element sums, fused
//                     (nbody/simulation.py:27-52, 131-170, 216-217)                FP64-pipe bound
//   k_fit             one warp per (system, planet): linear ephemeris, O-C, maxavg
//                     (nbody/simulation.py:172-177, 205-210; nbody/tools.py:22)    tiny, HBM/latency
//   k_ls_oc           one warp per (system, planet): floating-mean Lomb-Scargle of the O-C
//                     (nbody/simulation.py:213 -> nbody/tools.py:97-104)           FP64-pipe bound
//   k_transpose / k_rv_max / k_ls_series   RV products (nbody/simulation.py:186-188)
//   k_chi2            nested-sampling log-likelihood (nbody/nested.py:85-95)
//   k_integrate_whfast the reference's integrator='whfast' preset (WHFast, corrector 5; nbody/simulation.py:39-42)
//   k_integrate_ias15 the reference's own integrator (REBOUND's default IAS15) for the few systems no fixed step
//                     resolves (engine.run_verified)                                latency-bound, rare
//   k_ls_peaks        on-device reduction of the O-C periodograms to their highest peaks
//   k_dfma_peak       DFMA-chain microbenchmark: the measured FP64 roofline denominator
//
// There is no CPU path in this library: without a CUDA device every compute entry fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "nbt_core.cuh"
#include "nbt_ias15.cuh"
#include "nbt_lanes.cuh"
#include "nbt_whfast.cuh"

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int nbt_fail(int rc, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return rc;
}

#define NBT_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            rc = nbt_fail((int)e__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),    \
                          __FILE__, __LINE__);                                                  \
            goto done;                                                                          \
        }                                                                                       \
    } while (0)

// ---------------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// k-th and (k+1)-th smallest (0-based) of |x[0..n)| by bisection on the IEEE bit pattern (monotone for
// non-negative doubles): 63 warp-wide counting passes, no scratch memory, no sort.  The first 64 values
// live in registers (a transit series is ~37 long) and a count is one ballot + popc per 32 values; the
// (k+1)-th follows from the k-th in one more pass (a duplicate of it, or the smallest larger value).
__device__ double warp_kth_abs(const double* x, int n, int k, int lane, double& next) {
    const unsigned long long PAD = ~0ull;   // above every candidate
    const unsigned long long a0 = (lane < n) ? (unsigned long long)__double_as_longlong(fabs(x[lane])) : PAD;
    const unsigned long long a1 = (lane + 32 < n) ? (unsigned long long)__double_as_longlong(fabs(x[lane + 32])) : PAD;
    unsigned long long ans = 0ull;
    for (int bit = 62; bit >= 0; bit--) {
        const unsigned long long cand = ans | (1ull << bit);
        int c = __popc(__ballot_sync(0xffffffffu, a0 < cand)) + __popc(__ballot_sync(0xffffffffu, a1 < cand));
        for (int i0 = 64; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            c += __popc(__ballot_sync(0xffffffffu, i < n && (unsigned long long)__double_as_longlong(fabs(x[i])) < cand));
        }
        if (c <= k) ans = cand;
    }
    // (k+1)-th: values <= ans number at least k + 1; one more of them means a duplicate
    int le = __popc(__ballot_sync(0xffffffffu, a0 <= ans)) + __popc(__ballot_sync(0xffffffffu, a1 <= ans));
    unsigned long long up = min(a0 > ans ? a0 : PAD, a1 > ans ? a1 : PAD);
    for (int i0 = 64; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const unsigned long long v = (i < n) ? (unsigned long long)__double_as_longlong(fabs(x[i])) : PAD;
        le += __popc(__ballot_sync(0xffffffffu, v <= ans));
        up = min(up, v > ans ? v : PAD);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) up = min(up, __shfl_xor_sync(0xffffffffu, up, o));
    next = __longlong_as_double((long long)((le >= k + 2) ? ans : up));
    return __longlong_as_double((long long)ans);
}

// maxavg(x) = (percentile(|x|, 75, 'linear') + max|x|) / 2   — nbody/tools.py:22
__device__ double warp_maxavg(const double* x, int n, int lane) {
    double mx = 0.0;
    for (int i = lane; i < n; i += 32) mx = fmax(mx, fabs(x[i]));
    mx = warp_max(mx);
    const double pos = 0.75 * (double)(n - 1);
    const int lo = (int)floor(pos);
    const double frac = pos - (double)lo;
    double vhi;
    const double vlo = warp_kth_abs(x, n, lo, lane, vhi);
    const double p75 = (lo + 1 < n) ? vlo + (vhi - vlo) * frac : vlo;
    return 0.5 * (p75 + mx);
}

// ---------------------------------------------------------------------------------------------
// k_integrate: the hot kernel
// ---------------------------------------------------------------------------------------------
#ifndef NBT_BLOCK
#define NBT_BLOCK 64
#endif
#ifndef NBT_MINBLOCKS
#define NBT_MINBLOCKS 1
#ifndef NBT_LAT_MINB
#define NBT_LAT_MINB 1   /* resident CTAs per SM the latency variant (Np = 2) is compiled for: 1 = no register cap (216) */
#endif
#endif
#ifndef NBT_LAT_MAX_SYSTEMS
#define NBT_LAT_MAX_SYSTEMS (4 * 56832)   /* launches below this many systems take the latency variant (Np <= 2) */
#endif

// launch-order slots [first, first + count) of the batch, one system per thread.  The first main_count
// slots take A.scheme, the rest A.scheme_alt (NBT_SCHEME_AUTO), in separate blocks so that the scheme is
// uniform over a block; one launch keeps the dispatch order heaviest-first (two concurrent launches
// race for the SMs: the light one getting ahead cost 12 ms of the 72 ms bench launch).
// (the throughput variant for two planets sits at 6 resident CTAs per SM, 168 registers: keep it there)
template <int NP, bool SERIES, bool LAT>
__global__ void __launch_bounds__(NBT_BLOCK, (NP == 2 && !SERIES) ? (LAT ? NBT_LAT_MINB : 6) : NBT_MINBLOCKS)
k_integrate(const nbt_run_args A, int first, int count, int main_count) {
    const int main_blocks = (main_count + NBT_BLOCK - 1) / NBT_BLOCK;
    const bool alt = (int)blockIdx.x >= main_blocks;
    const int g = alt ? main_count + ((int)blockIdx.x - main_blocks) * NBT_BLOCK + (int)threadIdx.x
                      : (int)blockIdx.x * NBT_BLOCK + (int)threadIdx.x;
    if (g >= (alt ? count : main_count)) return;
    const int sys = (A.order != nullptr) ? A.order[first + g] : first + g;
    nbt_simulate_system<NP, SERIES, LAT>(A, alt ? A.scheme_alt : A.scheme, sys);
}

// launch-order slots [first, first + count), one planet per lane (32 / NP systems per warp)
#define NBT_LBLOCK 64
template <int NP, bool SERIES>
__global__ void __launch_bounds__(NBT_LBLOCK) k_integrate_lanes(const nbt_run_args A, int first, int count) {
    constexpr int SPW = 32 / NP;
    const int warp = (int)((blockIdx.x * (unsigned)NBT_LBLOCK + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    const int g = lane / NP, j = lane - g * NP;
    if (g >= SPW) return;                               // spare lanes (32 % NP)
    const int idx = warp * SPW + g;
    if (idx >= count) return;                           // whole group leaves together
    const int sys = (A.order != nullptr) ? A.order[first + idx] : first + idx;
    const int base = g * NP;
    const unsigned gmask = ((NP == 32) ? 0xffffffffu : ((1u << NP) - 1u)) << base;
    nbt_simulate_lane<NP, SERIES>(A, A.scheme, sys, j, base, gmask);
}

// ---------------------------------------------------------------------------------------------
// k_fit: linear ephemeris + O-C + maxavg, one warp per (system, planet)
// ---------------------------------------------------------------------------------------------
struct nbt_fit_args {
    const double* params;
    const int64_t* tt_offsets;
    const double* tt;
    const int32_t* n_tt;
    double* ttv; double* period; double* t0; double* ttv_max;
    int32_t* status;
    int32_t B, Np, flags;
};

__global__ void __launch_bounds__(128) k_fit(const nbt_fit_args F) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= F.B * F.Np) return;
    const int sys = w / F.Np, j = w % F.Np;
    const double Pj = F.params[((size_t)sys * (F.Np + 1) + j + 1) * NBT_NPAR + NBT_P_P];
    if ((F.flags & NBT_F_PLANET1_ONLY) && j > 0) {
        if (lane == 0) { F.period[w] = Pj; F.t0[w] = 0.0; F.ttv_max[w] = 0.0; }
        return;
    }
    const int64_t off = F.tt_offsets[w];
    const int cap = (int)(F.tt_offsets[w + 1] - off);
    const int n = min(F.n_tt[w], cap);
    const double* tt = F.tt + off;
    double* ttv = F.ttv + off;
    if (n >= 3) {
        // least squares of tt on k = 0..n-1, centred (what np.linalg.lstsq solves, simulation.py:172-177)
        double st = 0.0;
        for (int k = lane; k < n; k += 32) st += tt[k];
        const double tbar = warp_sum(st) / (double)n;
        const double kbar = 0.5 * (double)(n - 1);
        double skk = 0.0, skt = 0.0;
        for (int k = lane; k < n; k += 32) {
            const double dk = (double)k - kbar;
            skk = fma(dk, dk, skk);
            skt = fma(dk, tt[k] - tbar, skt);
        }
        skk = warp_sum(skk); skt = warp_sum(skt);
        const double m = skt / skk;
        const double b = tbar - m * kbar;
        for (int k = lane; k < n; k += 32) ttv[k] = tt[k] - m * (double)k - b;
        __syncwarp();
        const double mx = warp_maxavg(ttv, n, lane);
        if (lane == 0) { F.period[w] = m; F.t0[w] = b; F.ttv_max[w] = mx; }
    } else if (lane == 0) { // simulation.py:207-210: per = objects[j]['P'], t0 = 0, ttv = [0]
        atomicOr(&F.status[sys], NBT_S_FEW_TRANSITS);
        F.period[w] = Pj; F.t0[w] = 0.0; F.ttv_max[w] = 0.0;
        for (int k = 0; k < max(n, 1) && k < cap; k++) ttv[k] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// Floating-mean (generalised) Lomb-Scargle, standard normalisation, unit weights
// (astropy LombScargle(t, y).power(f) exact methods [3P]; nbody/tools.py:97-104).
// One pass, six sums; for uniformly sampled series the phasors advance by a complex rotation
// (2 DMUL + 2 DFMA) and are re-seeded exactly every NBT_LS_RESYNC samples (rotation round-off grows
// like 1e-16 per sample: 3e-14 after 256).
// ---------------------------------------------------------------------------------------------
#define NBT_LS_RESYNC 256
#define NBT_LS_TINY 1e-14

__device__ __forceinline__ double ls_finish(double C, double S, double C2, double S2, double YC, double YS,
                                            double invn, double YY) {
    C *= invn; S *= invn; C2 *= invn; S2 *= 2.0 * invn; YC *= invn; YS *= invn;
    const double S2p = S2 - 2.0 * S * C;
    const double C2p = C2 - (C * C - S * S);
    // cos / sin of tau = atan2(S2p, C2p) / 2 by half-angle formulas (tau in (-pi/2, pi/2], so cos tau >= 0)
    // instead of atan2 + sincos: the larger of cos^2, sin^2 through a square root, the other from sin 2 tau
    const double R2 = fma(S2p, S2p, C2p * C2p);
    const double iR = (R2 > 0.0) ? rsqrt(R2) : 0.0;
    const double c2t = (R2 > 0.0) ? C2p * iR : 1.0, s2t = S2p * iR;
    const bool fwd = c2t >= 0.0;
    const double big = sqrt(0.5 * (1.0 + fabs(c2t)));            // cos tau if fwd, |sin tau| otherwise
    const double small_ = 0.5 * s2t / big;                        // sin tau if fwd, cos tau * sign otherwise
    const double ct = fwd ? big : fabs(small_);
    const double st = fwd ? small_ : ((s2t >= 0.0) ? big : -big);
    const double YCt = YC * ct + YS * st, YSt = YS * ct - YC * st;
    const double Ct = C * ct + S * st, St = S * ct - C * st;
    const double rot = C2 * c2t + S2 * s2t;
    const double CC = 0.5 * (1.0 + rot) - Ct * Ct;
    const double SS = 0.5 * (1.0 - rot) - St * St;
    // A vanishing normalisation (e.g. the sine term at exactly the Nyquist frequency of integer
    // epochs) carries no power: the term is dropped instead of returning round-off / round-off.
    const double pc = (CC > NBT_LS_TINY) ? YCt * YCt / CC : 0.0;
    const double ps = (SS > NBT_LS_TINY) ? YSt * YSt / SS : 0.0;
    return (pc + ps) / YY;
}

// y: n samples at t_i = i (arbitrary origin/scale folded into nu = cycles per sample); every sum by summation
__device__ __noinline__ double ls_power_uniform_sum(const double* __restrict__ y, int n, double ybar, double YY, double nu) {
    double sd, cd;
    sincospi(2.0 * nu, &sd, &cd);
    // sum cos(2wt) = 2 sum cos^2 - n: one FMA per sample instead of two
    double C = 0, S = 0, CC = 0, S2 = 0, YC = 0, YS = 0;
    for (int i0 = 0; i0 < n; i0 += NBT_LS_RESYNC) {
        double s = 0.0, c = 1.0;
        if (i0 > 0) sincospi(2.0 * nu * (double)i0, &s, &c);
        const int i1 = min(n, i0 + NBT_LS_RESYNC);
        for (int i = i0; i < i1; i++) {
            const double yi = y[i] - ybar;
            C += c; S += s;
            CC = fma(c, c, CC);
            S2 = fma(s, c, S2);
            YC = fma(yi, c, YC); YS = fma(yi, s, YS);
            const double cn = fma(c, cd, -s * sd);
            s = fma(s, cd, c * sd);
            c = cn;
        }
    }
    return ls_finish(C, S, fma(2.0, CC, -(double)n), S2, YC, YS, 1.0 / (double)n, YY);
}

// The same with the data-independent sums in closed form: on t = 0..n-1
//   sum e^{i w t} = e^{i w (n-1)/2} sin(n w/2) / sin(w/2)      (w = 2 pi nu, and again for 2 w),
// so only sum y cos and sum y sin are left in the loop: 7 FP64 instructions per (sample, frequency)
// pair instead of 11.  Where a denominator is small (nu within 1.6e-4 of a multiple of 1/2: the exact
// Nyquist frequency of integer epochs, whose sine term must cancel to round-off) the sums are summed.
__device__ double ls_power_uniform(const double* __restrict__ y, int n, double ybar, double YY, double nu) {
    // two sincospi calls; the double angles and the angles n w / 2, n w by the addition formulas
    double sh, ch, sm, cm;
    const double dn = (double)n, dm = (double)(n - 1);
    sincospi(nu, &sh, &ch);               // w / 2
    sincospi(dm * nu, &sm, &cm);          // phase (n-1) w / 2
    const double sd = 2.0 * sh * ch, cd = fma(-2.0 * sh, sh, 1.0);        // w
    if (fabs(sd) < 1e-3 || fabs(sh) < 1e-3) return ls_power_uniform_sum(y, n, ybar, YY, nu);
    const double sm2 = 2.0 * sm * cm, cm2 = fma(-2.0 * sm, sm, 1.0);      // (n-1) w
    const double sn = fma(sm, ch, cm * sh), cn_ = fma(cm, ch, -sm * sh);   // n w / 2
    const double D = sn / sh, D2 = (2.0 * sn * cn_) / sd;
    const double C = cm * D, S = sm * D, C2 = cm2 * D2, S2 = 0.5 * (sm2 * D2);
    double YC = 0, YS = 0;
    for (int i0 = 0; i0 < n; i0 += NBT_LS_RESYNC) {
        double s = 0.0, c = 1.0;
        if (i0 > 0) sincospi(2.0 * nu * (double)i0, &s, &c);
        const int i1 = min(n, i0 + NBT_LS_RESYNC);
        for (int i = i0; i < i1; i++) {
            const double yi = y[i] - ybar;
            YC = fma(yi, c, YC); YS = fma(yi, s, YS);
            const double cn = fma(c, cd, -s * sd);
            s = fma(s, cd, c * sd);
            c = cn;
        }
    }
    return ls_finish(C, S, C2, S2, YC, YS, 1.0 / dn, YY);
}

__device__ double ls_power_times(const double* __restrict__ t, const double* __restrict__ y, int n, double ybar,
                                 double YY, double f) {
    double C = 0, S = 0, C2 = 0, S2 = 0, YC = 0, YS = 0;
    for (int i = 0; i < n; i++) {
        double s, c;
        sincospi(2.0 * (f * t[i]), &s, &c);
        const double yi = y[i] - ybar;
        C += c; S += s;
        C2 = fma(c, c, fma(-s, s, C2));
        S2 = fma(s, c, S2);
        YC = fma(yi, c, YC); YS = fma(yi, s, YS);
    }
    return ls_finish(C, S, C2, S2, YC, YS, 1.0 / (double)n, YY);
}

__device__ __forceinline__ void warp_mean_var(const double* __restrict__ y, int n, int lane, double& ybar, double& YY) {
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += y[i];
    ybar = warp_sum(s) / (double)n;
    double q = 0.0;
    for (int i = lane; i < n; i += 32) { const double d = y[i] - ybar; q = fma(d, d, q); }
    YY = warp_sum(q) / (double)n;
}

__device__ __forceinline__ double ls_df(double baseline, int spp) { return 1.0 / baseline / (double)spp; }
__device__ __forceinline__ long long ls_nfreq(double baseline, double fmin, double fmax, int spp) {
    return 1 + (long long)rint((fmax - fmin) / ls_df(baseline, spp));
}

struct nbt_lsoc_args {
    const int64_t* tt_offsets; const int64_t* ls_offsets;
    const double* ttv; const int32_t* n_tt;
    double* power; int32_t* nfreq;
    int32_t B, Np, flags, spp;
    int32_t w0, w1;   // (system, planet) series [w0, w1) of this launch
    int32_t nchunk;   // warps per series: warp c takes the frequency chunks c, c + nchunk, ... (NBT_LS_FCHUNK each)
    double fmin, fmax;
};

// A series of n transits has ~4.9 n frequencies and costs ~n^2: one warp per series leaves the longest
// series (456 transits, 2250 frequencies) as a tail of their own, so a series is cut into chunks of
// NBT_LS_FCHUNK frequencies that separate warps take (nchunk of them per series; the surplus warps of a
// short series exit at once).
#define NBT_LS_FCHUNK 256
__global__ void __launch_bounds__(128) k_ls_oc(const nbt_lsoc_args L) {
    const long long id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int w = L.w0 + (int)(id / L.nchunk), c0 = (int)(id % L.nchunk);
    const int lane = threadIdx.x & 31;
    if (w >= L.w1) return;
    const int j = w % L.Np;
    const int64_t off = L.tt_offsets[w];
    const int cap = (int)(L.tt_offsets[w + 1] - off);
    const int n = min(L.n_tt[w], cap);
    if (n < 3 || ((L.flags & NBT_F_PLANET1_ONLY) && j > 0)) {
        if (lane == 0 && c0 == 0) L.nfreq[w] = 0;
        return;
    }
    const int64_t loff = L.ls_offsets[w];
    const long long lcap = L.ls_offsets[w + 1] - loff;
    const double T = (double)(n - 1); // epochs = arange(n)
    long long nf = ls_nfreq(T, L.fmin, L.fmax, L.spp);
    if (nf > lcap) nf = lcap;
    if (lane == 0 && c0 == 0) L.nfreq[w] = (int32_t)nf;
    if ((long long)c0 * NBT_LS_FCHUNK >= nf) return;
    const double df = ls_df(T, L.spp);
    const double* y = L.ttv + off;
    double ybar, YY;
    warp_mean_var(y, n, lane, ybar, YY);
    for (long long k0 = (long long)c0 * NBT_LS_FCHUNK; k0 < nf; k0 += (long long)L.nchunk * NBT_LS_FCHUNK) {
        const long long k1 = min(nf, k0 + NBT_LS_FCHUNK);
        for (long long k = k0 + lane; k < k1; k += 32)
            L.power[loff + k] = ls_power_uniform(y, n, ybar, YY, L.fmin + (double)k * df);
    }
}

// NBT_F_LS_PEAKS: the NBT_LS_NPEAKS highest local maxima (p[k-1] < p[k] >= p[k+1], interior samples) of each O-C
// periodogram, as (frequency, power) pairs, highest first, NaN-padded.  One warp per (system, planet) series.  Callers
// that only read the peak of a periodogram (report()'s tables, simulation_grid) then need 48 bytes per series instead
// of ~5 n_tt doubles — the power arrays were 234 of the 340 MB a bench step copied back.
__global__ void __launch_bounds__(128) k_ls_peaks(const int64_t* __restrict__ tt_offsets, const int64_t* __restrict__ ls_offsets,
                                                  const int32_t* __restrict__ n_tt, const int32_t* __restrict__ nfreq,
                                                  const double* __restrict__ power, int nsp, int spp, double fmin,
                                                  double* __restrict__ peaks) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= nsp) return;
    double* out = peaks + (size_t)w * NBT_LS_NPEAKS * 2;
    const int nf = nfreq[w];
    const int cap = (int)(tt_offsets[w + 1] - tt_offsets[w]);
    const int n = min(n_tt[w], cap);
    if (nf < 3 || n < 3) {
        if (lane < NBT_LS_NPEAKS * 2) out[lane] = NAN;
        return;
    }
    const double* p = power + ls_offsets[w];
    const double df = ls_df((double)(n - 1), spp);
    // per-lane best NBT_LS_NPEAKS, then NBT_LS_NPEAKS rounds of warp arg-max
    double bp[NBT_LS_NPEAKS]; int bk[NBT_LS_NPEAKS];
#pragma unroll
    for (int q = 0; q < NBT_LS_NPEAKS; q++) { bp[q] = -1.0; bk[q] = -1; }
    for (int k = 1 + lane; k < nf - 1; k += 32) {
        const double v = p[k];
        if (p[k - 1] < v && v >= p[k + 1]) {
            double cv = v; int ck = k;
#pragma unroll
            for (int q = 0; q < NBT_LS_NPEAKS; q++)
                if (cv > bp[q]) { const double tv = bp[q]; const int tk = bk[q]; bp[q] = cv; bk[q] = ck; cv = tv; ck = tk; }
        }
    }
    for (int r = 0; r < NBT_LS_NPEAKS; r++) {
        double v = bp[0]; int k = bk[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int ok = __shfl_xor_sync(0xffffffffu, k, o);
            if (ov > v || (ov == v && ok >= 0 && (k < 0 || ok < k))) { v = ov; k = ok; }
        }
        if (k >= 0 && bk[0] == k) {   // the winning lane pops its head
#pragma unroll
            for (int q = 0; q + 1 < NBT_LS_NPEAKS; q++) { bp[q] = bp[q + 1]; bk[q] = bk[q + 1]; }
            bp[NBT_LS_NPEAKS - 1] = -1.0; bk[NBT_LS_NPEAKS - 1] = -1;
        }
        if (lane == 0) {
            out[2 * r] = (k >= 0) ? fmin + (double)k * df : NAN;
            out[2 * r + 1] = (k >= 0) ? v : NAN;
        }
    }
}

// generic batched periodogram: one warp per (series, chunk of 32 frequencies)
__global__ void __launch_bounds__(128) k_ls_series(const double* __restrict__ t, const double* __restrict__ y,
                                                   long long n_series, int n, double f0, double df,
                                                   long long n_freq, double tscale, double* __restrict__ power) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long chunks = (n_freq + 31) / 32;
    if (w >= n_series * chunks) return;
    const long long s = w / chunks, ch = w % chunks;
    const double* ys = y + s * (long long)n;
    double ybar, YY;
    warp_mean_var(ys, n, lane, ybar, YY);
    const long long k = ch * 32 + lane;
    if (k >= n_freq) return;
    const double f = f0 + (double)k * df;
    power[s * n_freq + k] = (t != nullptr) ? ls_power_times(t + s * (long long)n, ys, n, ybar, YY, f)
                                           : ls_power_uniform(ys, n, ybar, YY, f * tscale);
}

// [rows][cols] -> [cols][rows] through a padded shared tile (coalesced both ways)
__global__ void k_transpose(const double* __restrict__ in, double* __restrict__ out, int rows, int cols) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int rr = r0 + r, cc = c0 + threadIdx.x;
        if (rr < rows && cc < cols) tile[r][threadIdx.x] = in[(size_t)rr * cols + cc];
    }
    __syncthreads();
    for (int c = threadIdx.y; c < 32; c += blockDim.y) {
        const int cc = c0 + c, rr = r0 + threadIdx.x;
        if (rr < rows && cc < cols) out[(size_t)cc * rows + rr] = tile[threadIdx.x][c];
    }
}

__global__ void __launch_bounds__(128) k_rv_max(const double* __restrict__ rv_sm, int B, int n, double* __restrict__ rv_max) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= B) return;
    const double mx = warp_maxavg(rv_sm + (size_t)w * n, n, lane);
    if (lane == 0) rv_max[w] = mx;
}


// ---------------------------------------------------------------------------------------------
// Stand-alone forms of the reference's array functions (same arithmetic as the fused path)
// ---------------------------------------------------------------------------------------------
// transit_times(xp, xs, times): nbody/simulation.py:163-170 + tools.py:61-65, one thread per series
__global__ void k_transit_times(const double* __restrict__ xp, const double* __restrict__ xs,
                                const double* __restrict__ times, long long n_series, long long n,
                                double* __restrict__ tt, int32_t* __restrict__ n_tt, long long cap, int times_shared) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_series) return;
    const double* a = xp + s * n; const double* b = xs + s * n;
    const double* t = times_shared ? times : times + s * n;
    double prev = a[0] - b[0];
    int c = 0;
    for (long long i = 1; i < n; i++) {
        const double cur = a[i] - b[i];
        if (prev >= 0.0 && cur <= 0.0) {
            const double m = (cur - prev) / (t[i] - t[i - 1]);
            const double t0 = -prev / m + t[i - 1];
            if (c < cap) tt[s * cap + c] = t0;
            c++;
        }
        prev = cur;
    }
    n_tt[s] = c;
}

// find_zero(t1, dx1, t2, dx2), elementwise: nbody/tools.py:61-65
__global__ void k_find_zero(const double* __restrict__ t1, const double* __restrict__ dx1,
                            const double* __restrict__ t2, const double* __restrict__ dx2, long long n,
                            double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = (dx2[i] - dx1[i]) / (t2[i] - t1[i]);
    out[i] = -dx1[i] / m + t1[i];
}

// TTV(epochs, tt): least squares tt ~ b + m*epoch, one warp per series (nbody/simulation.py:172-177)
__global__ void __launch_bounds__(128) k_ttv_fit(const double* __restrict__ epochs, const double* __restrict__ tt,
                                                 long long n_series, int n, double* __restrict__ ttv,
                                                 double* __restrict__ slope, double* __restrict__ icpt) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_series) return;
    const double* t = tt + w * n;
    const double* ep = epochs ? epochs + w * n : nullptr;
    double st = 0.0, sk = 0.0;
    for (int k = lane; k < n; k += 32) { st += t[k]; sk += ep ? ep[k] : (double)k; }
    const double tbar = warp_sum(st) / (double)n, kbar = warp_sum(sk) / (double)n;
    double skk = 0.0, skt = 0.0;
    for (int k = lane; k < n; k += 32) {
        const double dk = (ep ? ep[k] : (double)k) - kbar;
        skk = fma(dk, dk, skk);
        skt = fma(dk, t[k] - tbar, skt);
    }
    skk = warp_sum(skk); skt = warp_sum(skt);
    const double m = (skk > 0.0) ? skt / skk : 0.0; // rank-deficient (n == 1): minimum-norm solution of lstsq
    const double b = tbar - m * kbar;
    for (int k = lane; k < n; k += 32) ttv[w * n + k] = t[k] - m * (ep ? ep[k] : (double)k) - b;
    if (lane == 0) { slope[w] = m; icpt[w] = b; }
}

__global__ void __launch_bounds__(128) k_maxavg(const double* __restrict__ x, long long n_series, int n,
                                                double* __restrict__ out) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_series) return;
    const double v = warp_maxavg(x + w * n, n, lane);
    if (lane == 0) out[w] = v;
}

// analyze() front end on a recorded sim_data series (nbody/simulation.py:179-231): one thread per
// (system, planet) scans the time-major series [n_outputs][B][3+10*Np] and produces what
// k_integrate produces in-flight: transit times, element means, RV, orbit samples.
struct nbt_scan_args {
    const double* series; const double* times; const int64_t* tt_offsets;
    double* tt; int32_t* n_tt; double* mean_a; double* mean_e; double* mean_inc; double* mean_omega;
    int32_t* status; double* rv; double* orbit_xz;
    int32_t B, Np, n_outputs, flags, orbit_stride;
    double rv_scale;
};

__global__ void __launch_bounds__(128) k_scan_series(const nbt_scan_args A) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)A.B * A.Np) return;
    const int sys = (int)(g / A.Np), j = (int)(g % A.Np);
    const int W = 3 + 10 * A.Np, NO = A.n_outputs;
    const size_t rowstride = (size_t)A.B * W;
    const double* base = A.series + (size_t)sys * W;
    const double* pj = base + 3 + 10 * j;
    const bool tt_on = !((A.flags & NBT_F_PLANET1_ONLY) && j > 0);
    const bool want_means = (A.flags & NBT_F_MEANS) != 0;
    const bool want_rv = (A.flags & NBT_F_RV) != 0 && A.rv != nullptr && j == 0;
    const bool want_orbit = (A.flags & NBT_F_ORBIT) != 0 && A.orbit_xz != nullptr;
    const int64_t off = A.tt_offsets[g];
    const int cap = (int)(A.tt_offsets[g + 1] - off);
    double sa = 0, se = 0, si = 0, so = 0;
    double prev = 0, xs_prev = 0, t_prev = 0;
    int cnt = 0, status = 0;
    for (int o = 0; o < NO; o++) {
        const double* star = base + (size_t)o * rowstride;
        const double* p = pj + (size_t)o * rowstride;
        const double xs = star[0], t = A.times[o];
        const double cur = p[0] - xs;
        if (o > 0) {
            if (tt_on && prev >= 0.0 && cur <= 0.0) {
                const double m = (cur - prev) / (t - t_prev);
                const double t0 = -prev / m + t_prev;
                if (cnt < cap) A.tt[off + cnt] = t0; else status |= NBT_S_TT_OVERFLOW;
                cnt++;
            }
            if (want_rv) A.rv[(size_t)(o - 1) * A.B + sys] = (xs - xs_prev) * A.rv_scale;
        }
        prev = cur; xs_prev = xs; t_prev = t;
        if (want_means) {
            sa += p[4]; se += p[5]; si += p[6]; so += p[8];
            if (!(p[4] > 0.0)) status |= NBT_S_UNBOUND;
            if (!isfinite(p[0] + p[1] + p[2] + p[4] + p[5])) status |= NBT_S_NONFINITE;
        }
        if (want_orbit && (o % A.orbit_stride) == 0) {
            double* dst = A.orbit_xz + (((size_t)(o / A.orbit_stride) * A.B + sys) * A.Np + j) * 2;
            dst[0] = p[0]; dst[1] = p[2];
        }
    }
    A.n_tt[g] = cnt;
    if (want_means) {
        const double inv = 1.0 / (double)NO;
        A.mean_a[g] = sa * inv; A.mean_e[g] = se * inv; A.mean_inc[g] = si * inv;
        if (A.mean_omega) A.mean_omega[g] = so * inv;
    }
    if (status) atomicOr(&A.status[sys], status);
}


// ---------------------------------------------------------------------------------------------
// k_grid_post: post-processing of one grid cell (simulations/simulation_grid.py:33-69), one block
// per system, planet 1 only:
//   exact periodogram of the O-C on astropy's autofrequency grid (spp, fmin, fmax from the caller)
//   -> scipy.signal.find_peaks(power, height) -> sort by height, keep n_order
//   -> least squares of Tc on [1, n, sin(2 pi n / P_k), cos(2 pi n / P_k) ...]   (np.linalg.lstsq)
//   -> softmax(ttv) = 0.5 * percentile(|ttv|, 95), avg_err = mean |model - ttv|
// Dynamic shared memory: max(n_freq, n_tt * (2 + 2 n_order)) doubles (power, then the basis).
// ---------------------------------------------------------------------------------------------
#define NBT_GRID_MAXORDER 8
#define NBT_GRID_MAXCOL (2 + 2 * NBT_GRID_MAXORDER)

struct nbt_grid_args {
    const double* tt; const double* ttv; const int64_t* tt_offsets; const int32_t* n_tt;
    double* peak_periods; double* coeffs; int32_t* n_peaks; double* softmax; double* avg_err;
    int32_t B, Np, spp, n_order, smem_doubles;
    double fmin, fmax, height;
};

__device__ double warp_percentile_abs(const double* x, int n, double q, int lane) {
    const double pos = q * (double)(n - 1);
    const int lo = (int)floor(pos);
    const double frac = pos - (double)lo;
    double vhi;
    const double vlo = warp_kth_abs(x, n, lo, lane, vhi);
    if (lo + 1 >= n) return vlo;
    return vlo + (vhi - vlo) * frac;
}

__global__ void __launch_bounds__(128) k_grid_post(const nbt_grid_args G) {
    extern __shared__ double sm[];
    __shared__ double s_red[4];
    __shared__ double s_G[NBT_GRID_MAXCOL * NBT_GRID_MAXCOL], s_rhs[NBT_GRID_MAXCOL], s_c[NBT_GRID_MAXCOL];
    __shared__ double s_pk_period[NBT_GRID_MAXORDER];
    __shared__ int s_npk;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t off = G.tt_offsets[(size_t)b * G.Np];
    const int cap = (int)(G.tt_offsets[(size_t)b * G.Np + 1] - off);
    const int n = min(G.n_tt[(size_t)b * G.Np], cap);
    const double* tt = G.tt + off;
    const double* ttv = G.ttv + off;
    const int m_max = 2 + 2 * G.n_order;
    double* out_c = G.coeffs + (size_t)b * m_max;
    double* out_p = G.peak_periods + (size_t)b * G.n_order;
    if (n < 3) { // too few transits: the reference's lstsq/LombScargle would not run either
        for (int i = tid; i < m_max; i += 128) out_c[i] = NAN;
        for (int i = tid; i < G.n_order; i += 128) out_p[i] = NAN;
        if (tid == 0) { G.n_peaks[b] = 0; G.softmax[b] = NAN; G.avg_err[b] = NAN; }
        return;
    }
    // 1. periodogram into shared memory
    const double T = (double)(n - 1);
    long long nf = ls_nfreq(T, G.fmin, G.fmax, G.spp);
    if (nf > G.smem_doubles) nf = G.smem_doubles;
    const double df = ls_df(T, G.spp);
    double ybar, YY;
    warp_mean_var(ttv, n, lane, ybar, YY);   // every warp computes the same mean/variance
    for (long long k = tid; k < nf; k += 128) sm[k] = ls_power_uniform(ttv, n, ybar, YY, G.fmin + (double)k * df);
    __syncthreads();
    // 2. find_peaks(height) + top n_order by height (scipy: strict rise, plateau midpoint, strict fall;
    //    argsort(heights)[::-1] -> among equal heights the later peak first)
    if (tid == 0) {
        int idx[NBT_GRID_MAXORDER]; double hgt[NBT_GRID_MAXORDER];
        int np = 0;
        long long i = 1;
        const long long imax = nf - 1;
        while (i < imax) {
            if (sm[i - 1] < sm[i]) {
                long long ahead = i + 1;
                while (ahead < imax && sm[ahead] == sm[i]) ahead++;
                if (sm[ahead] < sm[i]) {
                    const long long mid = (i + ahead - 1) / 2;
                    const double hv = sm[mid];
                    if (hv >= G.height) { // insert, keeping (height desc, index desc) order
                        int pos = np < G.n_order ? np : G.n_order;
                        while (pos > 0 && (hgt[pos - 1] < hv || hgt[pos - 1] == hv)) pos--;
                        if (pos < G.n_order) {
                            const int last = (np < G.n_order ? np : G.n_order - 1);
                            for (int q = last; q > pos; q--) { idx[q] = idx[q - 1]; hgt[q] = hgt[q - 1]; }
                            idx[pos] = (int)mid; hgt[pos] = hv;
                            if (np < G.n_order) np++;
                        }
                    }
                    i = ahead;
                }
            }
            i++;
        }
        s_npk = np;
        for (int q = 0; q < G.n_order; q++) {
            const double per = (q < np) ? 1.0 / (G.fmin + (double)idx[q] * df) : NAN;
            s_pk_period[q] = per;
            out_p[q] = (q < np) ? per : 0.0;   // per_epoch_grid is zero-initialised in the reference
        }
        G.n_peaks[b] = np;
    }
    __syncthreads();
    const int np = s_npk;
    const int m = 2 + 2 * np;
    // 3. basis in shared memory (column-major [m][n]); the epoch column is centred and scaled so that the
    //    normal equations stay well conditioned
    const double kbar = 0.5 * (double)(n - 1), ksc = 1.0 / fmax(kbar, 1.0);
    if ((long long)n * m <= G.smem_doubles) {
        for (int k = tid; k < n; k += 128) {
            sm[k] = 1.0;
            sm[n + k] = ((double)k - kbar) * ksc;
            for (int q = 0; q < np; q++) {
                double sn, cs;
                sincospi(2.0 * (double)k / s_pk_period[q], &sn, &cs);
                sm[(2 + 2 * q) * n + k] = sn;
                sm[(3 + 2 * q) * n + k] = cs;
            }
        }
        __syncthreads();
        for (int e = tid; e < m * m + m; e += 128) {
            double acc = 0.0;
            if (e < m * m) {
                const int a = e / m, c = e % m;
                if (c >= a) { for (int k = 0; k < n; k++) acc = fma(sm[a * n + k], sm[c * n + k], acc); s_G[a * m + c] = acc; }
            } else {
                const int a = e - m * m;
                for (int k = 0; k < n; k++) acc = fma(sm[a * n + k], tt[k], acc);
                s_rhs[a] = acc;
            }
        }
        __syncthreads();
        if (tid == 0) { // Cholesky of the (m x m) SPD normal matrix, upper triangle in s_G
            bool ok = true;
            for (int a = 0; a < m && ok; a++) {
                for (int c = a; c < m; c++) {
                    double v = s_G[a * m + c];
                    for (int q = 0; q < a; q++) v -= s_G[q * m + a] * s_G[q * m + c];
                    if (c == a) { if (!(v > 0.0)) { ok = false; break; } s_G[a * m + a] = sqrt(v); }
                    else s_G[a * m + c] = v / s_G[a * m + a];
                }
            }
            if (ok) {
                for (int a = 0; a < m; a++) { double v = s_rhs[a]; for (int q = 0; q < a; q++) v -= s_G[q * m + a] * s_c[q]; s_c[a] = v / s_G[a * m + a]; }
                for (int a = m - 1; a >= 0; a--) { double v = s_c[a]; for (int q = a + 1; q < m; q++) v -= s_G[a * m + q] * s_c[q]; s_c[a] = v / s_G[a * m + a]; }
            } else for (int a = 0; a < m; a++) s_c[a] = NAN;
        }
        __syncthreads();
        // 4. model - ttv; the harmonic part of the model is ypred - (c0 + c1 n)
        double err = 0.0;
        for (int k = tid; k < n; k += 128) {
            double h = 0.0;
            for (int a = 2; a < m; a++) h = fma(s_c[a], sm[a * n + k], h);
            err += fabs(h - ttv[k]);
        }
        err = warp_sum(err);
        if (lane == 0) s_red[wid] = err;
        __syncthreads();
        if (tid == 0) {
            G.avg_err[b] = (s_red[0] + s_red[1] + s_red[2] + s_red[3]) / (double)n;
            // back to the unscaled epoch column: c1' (k - kbar) ksc + c0' = (c1' ksc) k + (c0' - c1' ksc kbar)
            const double c1 = s_c[1] * ksc;
            out_c[0] = s_c[0] - c1 * kbar;
            out_c[1] = c1;
            for (int a = 2; a < m_max; a++) out_c[a] = (a < m) ? s_c[a] : 0.0;
        }
    } else if (tid == 0) {
        for (int a = 0; a < m_max; a++) out_c[a] = NAN;
        G.avg_err[b] = NAN;
    }
    if (wid == 1) { // softmax = percentile(|ttv|, 95) / 2   (simulation_grid.py:17)
        const double p95 = warp_percentile_abs(ttv, n, 0.95, lane);
        if (lane == 0) G.softmax[b] = 0.5 * p95;
    }
}

// ---------------------------------------------------------------------------------------------
// k_chi2: nbody/nested.py:85-95, one thread per forward model
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double nbt_loss(int loss, double z) {   // nested.py:79-82
    return loss == NBT_LOSS_SOFT_L1 ? 2.0 * (sqrt(1.0 + z) - 1.0) : z;
}

__global__ void k_chi2(const double* __restrict__ ttv, const int64_t* __restrict__ tt_offsets,
                       const int32_t* __restrict__ n_tt, int B, int Np, const double* __restrict__ c0,
                       const double* __restrict__ c1, const double* __restrict__ x,
                       const double* __restrict__ y, const double* __restrict__ yerr, int n_data, int loss,
                       double* __restrict__ loglike) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t off = tt_offsets[(size_t)b * Np];
    const int cap = (int)(tt_offsets[(size_t)b * Np + 1] - off);
    const int n0 = min(n_tt[(size_t)b * Np], cap);
    const double* tv = ttv + off;
    const double a0 = c0[b], a1 = c1[b];
    bool ok = true;
    for (int i = 0; i < n_data; i++) { const int k = (int)x[i]; if (k >= n0 || k < -n0) ok = false; }
    double s = 0.0;
    if (ok) {
        for (int i = 0; i < n_data; i++) {
            int k = (int)x[i]; if (k < 0) k += n0;
            const double r = (y[i] - (a1 * x[i] + a0) - tv[k]) / yerr[i];
            s += nbt_loss(loss, r * r);
        }
        loglike[b] = -0.5 * s;
    } else { // IndexError branch: -10 * sum over the ttv that exist, paired index-for-index
        for (int i = 0; i < n0 && i < n_data; i++) {
            const double r = (y[i] - (a1 * x[i] + a0) - tv[i]) / yerr[i];
            s += nbt_loss(loss, r * r);
        }
        loglike[b] = -10.0 * s;
    }
}


// ---------------------------------------------------------------------------------------------
// k_ttvfit: likelihood of ttv_fit.py:47-94 (UltraNest fitter), one thread per forward model.
// For every planet with data: Tc_sim = its transit times; shifted so that planet 1's first
// transit is 0; offset by the mean residual at the observed orbit numbers; chi2 accumulated.
// A missing orbit number (IndexError in the reference) gives that planet's term -1e6.
// ---------------------------------------------------------------------------------------------
__global__ void k_ttvfit(const double* __restrict__ tt, const int64_t* __restrict__ tt_offsets,
                         const int32_t* __restrict__ n_tt, int B, int Np, const int32_t* __restrict__ has_data,
                         const double* __restrict__ tc, const double* __restrict__ tc_err,
                         const int32_t* __restrict__ orbit, int n_obs, double* __restrict__ loglike) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double chi2 = 0.0, sim_shift = 0.0;
    for (int j = 0; j < Np; j++) {
        if (!has_data[j]) continue;
        const int64_t off = tt_offsets[(size_t)b * Np + j];
        const int cap = (int)(tt_offsets[(size_t)b * Np + j + 1] - off);
        const int n = min(n_tt[(size_t)b * Np + j], cap);
        const double* ts = tt + off;
        bool ok = n > 0;
        if (ok && j == 0) sim_shift = ts[0];                 // Tc_sim.min(): transit times are increasing
        for (int i = 0; i < n_obs && ok; i++) { const int k = orbit[i]; if (k >= n || k < -n) ok = false; }
        if (!ok) { chi2 += -1e6; continue; }
        const double* d = tc + (size_t)j * n_obs;
        const double* e = tc_err + (size_t)j * n_obs;
        double mean = 0.0;
        for (int i = 0; i < n_obs; i++) { int k = orbit[i]; if (k < 0) k += n; mean += d[i] - (ts[k] - sim_shift); }
        mean /= (double)n_obs;
        double s2 = 0.0;
        for (int i = 0; i < n_obs; i++) {
            int k = orbit[i]; if (k < 0) k += n;
            const double r = (d[i] - ((ts[k] - sim_shift) + mean)) / e[i];
            s2 = fma(r, r, s2);
        }
        chi2 += -0.5 * s2;
    }
    loglike[b] = chi2;
}

// ---------------------------------------------------------------------------------------------
// k_dfma_peak: 8 independent DFMA chains per thread — the measured FP64 roofline denominator
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
    double v0 = threadIdx.x * 1e-9, v1 = v0 + 1e-3, v2 = v0 + 2e-3, v3 = v0 + 3e-3, v4 = v0 + 4e-3,
           v5 = v0 + 5e-3, v6 = v0 + 6e-3, v7 = v0 + 7e-3;
    for (int i = 0; i < iters; i++) {
        v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
        v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
    }
    const double r = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
    if (r == 12345.678) out[0] = r; // never true; keeps the chains alive
}

// same, but all three DFMA operands are distinct per-thread registers (what real code looks like:
// six 32-bit register reads per instruction instead of two)
__global__ void __launch_bounds__(256) k_dfma_peak3(double* out, int iters, double a, double b) {
    double v0 = threadIdx.x * 1e-9, v1 = v0 + 1e-3, v2 = v0 + 2e-3, v3 = v0 + 3e-3, v4 = v0 + 4e-3,
           v5 = v0 + 5e-3, v6 = v0 + 6e-3, v7 = v0 + 7e-3;
    double a0 = a + v0, a1 = a + v1, a2 = a - v2, a3 = a - v3, b0 = b + v4, b1 = b - v5, b2 = b + v6, b3 = b - v7;
    for (int i = 0; i < iters; i++) {
        v0 = fma(v0, a0, b0); v1 = fma(v1, a1, b1); v2 = fma(v2, a2, b2); v3 = fma(v3, a3, b3);
        v4 = fma(v4, a1, b2); v5 = fma(v5, a2, b3); v6 = fma(v6, a3, b0); v7 = fma(v7, a0, b1);
    }
    const double r = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
    if (r == 12345.678) out[0] = r;
}

// one warp, one dependent DFMA / MUFU.RSQ64H chain: cycles per dependent instruction
__global__ void k_fp64_latency(double* out, long long* cycles, int n, double a, double b) {
    double v = threadIdx.x * 1e-9 + 1.0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) v = fma(v, a, b);
    long long t1 = clock64();
    double w = v;
#pragma unroll 16
    for (int i = 0; i < n; i++) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w)); w = y + 1.5; }
    long long t2 = clock64();
    if (threadIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t1; }
    out[threadIdx.x] = v + w;
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
/* Smallest mutual Hill separation (a_j - a_i) / R_H,ij over all planet pairs of one system,
 * R_H,ij = ((m_i + m_j) / 3 M*)^(1/3) (a_i + a_j) / 2, a ~ P^(2/3).  Pairs inside 2 sqrt(3) are
 * Hill-unstable: chaotic, mostly short-lived (every system the kernel flags as broken in the
 * C2 priors has Delta < 4). */
static double hill_separation(const double* par, int N) {
    const double Ms = par[NBT_P_M];
    double dmin = INFINITY;
    for (int j = 1; j < N; j++)
        for (int i = 1; i < j; i++) {
            const double Pi = par[i * NBT_NPAR + NBT_P_P], Pj = par[j * NBT_NPAR + NBT_P_P];
            if (!(Pi > 0.0) || !(Pj > 0.0) || !(Ms > 0.0)) continue;
            const double ai = cbrt(Pi * Pi), aj = cbrt(Pj * Pj);
            const double rh = cbrt((par[i * NBT_NPAR + NBT_P_M] + par[j * NBT_NPAR + NBT_P_M]) / (3.0 * Ms)) * 0.5 * (ai + aj);
            const double d = fabs(aj - ai) / rh;
            if (d < dmin) dmin = d;
        }
    return dmin;
}

static int scheme_stages(int id) {
    switch (id) {
    case NBT_SCHEME_WH2: return 1;
    case NBT_SCHEME_Y4: return 3;
    case NBT_SCHEME_SABA4: return 4;
    case NBT_SCHEME_M64: return 4;
    case NBT_SCHEME_C4: return 2;
    case NBT_SCHEME_SBABC3: return 3;
    case NBT_SCHEME_AUTO: return 3;   /* systems with substeps > 0; the marked ones (< 0) take C4's 2 */
    case NBT_SCHEME_IAS15: return 1;  /* adaptive: steps are not known in advance (counted as one per output) */
    case NBT_SCHEME_WHFAST: return 1;
    default: return -1;
    }
}

static int check_cfg(const nbt_config* cfg) {
    if (!cfg) return nbt_fail(-1, "cfg is NULL");
    if (cfg->struct_size != (int32_t)sizeof(nbt_config))
        return nbt_fail(-2, "nbt_config.struct_size %d != %d (ABI mismatch)", cfg->struct_size, (int)sizeof(nbt_config));
    if (cfg->n_systems < 0) return nbt_fail(-3, "n_systems < 0");
    if (cfg->n_bodies < 2 || cfg->n_bodies > NBT_MAX_BODIES) return nbt_fail(-4, "n_bodies must be in 2..%d", NBT_MAX_BODIES);
    if (cfg->n_outputs < 2) return nbt_fail(-5, "n_outputs must be >= 2");
    if (!(cfg->n_days > 0.0)) return nbt_fail(-6, "n_days must be > 0");
    if (scheme_stages(cfg->scheme) < 0) return nbt_fail(-7, "unknown scheme %d", cfg->scheme);
    if (cfg->tt_mode != NBT_TT_GRID_LINEAR && cfg->tt_mode != NBT_TT_REFINED) return nbt_fail(-8, "unknown tt_mode");
    return 0;
}

template <int NP>
static cudaError_t launch_integrate(const nbt_run_args& A, int first, int count, int main_count, bool lanes, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    const bool series = (A.flags & NBT_F_SERIES) != 0 && A.series != nullptr;
    if (lanes) {
        constexpr int SPW = 32 / NP;
        const int warps = (count + SPW - 1) / SPW;
        const int grid = (warps * 32 + NBT_LBLOCK - 1) / NBT_LBLOCK;
        if (series) k_integrate_lanes<NP, true><<<grid, NBT_LBLOCK, 0, st>>>(A, first, count);
        else k_integrate_lanes<NP, false><<<grid, NBT_LBLOCK, 0, st>>>(A, first, count);
    } else {
        if (main_count > count) main_count = count;
        if (main_count < 0) main_count = 0;
        const int grid = (main_count + NBT_BLOCK - 1) / NBT_BLOCK + (count - main_count + NBT_BLOCK - 1) / NBT_BLOCK;
        // Np <= 2 has a latency variant (branch-free drifts, more registers): it wins unless the launch is
        // many waves of resident threads deep (148 SMs x 6 CTAs x 64 threads = 56 832 per wave)
        const bool lat = NP <= 2 && count < NBT_LAT_MAX_SYSTEMS;
        if (series) k_integrate<NP, true, false><<<grid, NBT_BLOCK, 0, st>>>(A, first, count, main_count);
        else if (lat) k_integrate<NP, false, (NP <= 2)><<<grid, NBT_BLOCK, 0, st>>>(A, first, count, main_count);
        else k_integrate<NP, false, false><<<grid, NBT_BLOCK, 0, st>>>(A, first, count, main_count);
    }
    return cudaGetLastError();
}

static cudaError_t launch_integrate_np(int np, const nbt_run_args& A, int first, int count, int main_count, bool lanes, cudaStream_t st) {
    switch (np) {
    case 1: return launch_integrate<1>(A, first, count, main_count, false, st);
    case 2: return launch_integrate<2>(A, first, count, main_count, lanes, st);
    case 3: return launch_integrate<3>(A, first, count, main_count, lanes, st);
    case 4: return launch_integrate<4>(A, first, count, main_count, lanes, st);
    case 5: return launch_integrate<5>(A, first, count, main_count, lanes, st);
    case 6: return launch_integrate<6>(A, first, count, main_count, lanes, st);
    case 7: return launch_integrate<7>(A, first, count, main_count, lanes, st);
    case 8: return launch_integrate<8>(A, first, count, main_count, lanes, st);
    default: return cudaErrorInvalidValue;
    }
}

// NBT_SCHEME_IAS15: one system per thread, warp-sized blocks (a handful of systems: spread them over the SMs)
template <int NP>
__global__ void __launch_bounds__(32) k_integrate_ias15(const nbt_run_args A, const nbt_ias15_tables T) {
    const int sys = (int)(blockIdx.x * 32u + threadIdx.x);
    if (sys >= A.B) return;
    nbt_simulate_system_ias15<NP>(A, T, sys);
}

static const nbt_ias15_tables& ias15_tables() {
    static nbt_ias15_tables T;
    static std::once_flag once;
    std::call_once(once, [] { nbt_ias15_make_tables(T); });
    return T;
}

// NBT_SCHEME_WHFAST: the reference's integrator='whfast' preset (nbody/simulation.py:39-42), one system per thread
template <int NP>
__global__ void __launch_bounds__(32) k_integrate_whfast(const nbt_run_args A) {
    const int sys = (int)(blockIdx.x * 32u + threadIdx.x);
    if (sys >= A.B) return;
    nbt_simulate_system_whfast<NP>(A, sys);
}

static cudaError_t launch_special_np(int np, int scheme, const nbt_run_args& A, cudaStream_t st) {
    const int grid = (A.B + 31) / 32;
    if (scheme == NBT_SCHEME_WHFAST) {
        switch (np) {
        case 1: k_integrate_whfast<1><<<grid, 32, 0, st>>>(A); break;
        case 2: k_integrate_whfast<2><<<grid, 32, 0, st>>>(A); break;
        case 3: k_integrate_whfast<3><<<grid, 32, 0, st>>>(A); break;
        case 4: k_integrate_whfast<4><<<grid, 32, 0, st>>>(A); break;
        case 5: k_integrate_whfast<5><<<grid, 32, 0, st>>>(A); break;
        case 6: k_integrate_whfast<6><<<grid, 32, 0, st>>>(A); break;
        case 7: k_integrate_whfast<7><<<grid, 32, 0, st>>>(A); break;
        case 8: k_integrate_whfast<8><<<grid, 32, 0, st>>>(A); break;
        default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
    if (scheme != NBT_SCHEME_IAS15) return cudaErrorNotSupported;
    const nbt_ias15_tables& T = ias15_tables();
    switch (np) {
    case 1: k_integrate_ias15<1><<<grid, 32, 0, st>>>(A, T); break;
    case 2: k_integrate_ias15<2><<<grid, 32, 0, st>>>(A, T); break;
    case 3: k_integrate_ias15<3><<<grid, 32, 0, st>>>(A, T); break;
    case 4: k_integrate_ias15<4><<<grid, 32, 0, st>>>(A, T); break;
    case 5: k_integrate_ias15<5><<<grid, 32, 0, st>>>(A, T); break;
    case 6: k_integrate_ias15<6><<<grid, 32, 0, st>>>(A, T); break;
    case 7: k_integrate_ias15<7><<<grid, 32, 0, st>>>(A, T); break;
    case 8: k_integrate_ias15<8><<<grid, 32, 0, st>>>(A, T); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// side streams per device: the lanes-kernel launches of a call (the long systems) run concurrently with
// the one-thread-per-system launch; high priority, because the long systems should get their SMs first
// (per calling thread: concurrent nbt_run calls from several host threads — engine.Pipeline — must not queue their
// copies behind each other's events, nor wait for each other in cudaStreamSynchronize)
struct SideStream { cudaStream_t st[2] = {nullptr, nullptr}; };
static thread_local SideStream g_side[64];

static cudaError_t get_side(int dev, SideStream** out) {
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    SideStream& s = g_side[dev];
    for (int i = 0; i < 2; i++)
        if (!s.st[i]) {
            int least = 0, greatest = 0;
            cudaError_t e = cudaDeviceGetStreamPriorityRange(&least, &greatest);
            if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&s.st[i], cudaStreamNonBlocking, greatest);
            if (e != cudaSuccess) return e;
        }
    *out = &s;
    return cudaSuccess;
}

#define NBT_LS_CHUNKS 4   // host path: ranges in which the O-C periodograms are computed and copied back

static std::mutex g_pool_mu;
static bool g_pool_ready[64] = {false};

static void ensure_pool(int dev) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (dev < 0 || dev >= 64 || g_pool_ready[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX; // keep freed blocks: repeated calls (sampler callbacks) reuse them
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    g_pool_ready[dev] = true;
}

// Optional in-stream kernel timing (cfg.flags & NBT_F_TIMING): CUDA events recorded on the launch
// stream around each stage of the last nbt_run of this thread; read back with nbt_timing().
struct TimingState {
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool have = false; int device = -1;
};
static thread_local TimingState g_timing;

static void timing_mark(int i, bool on, cudaStream_t st) {
    if (!on) return;
    if (!g_timing.ev[i]) cudaEventCreate(&g_timing.ev[i]);
    cudaEventRecord(g_timing.ev[i], st);
}

// bump allocator over one stream-ordered allocation
struct Arena {
    char* base = nullptr; size_t size = 0, used = 0;
    size_t reserve(size_t bytes) { size_t o = size; size += (bytes + 255) & ~(size_t)255; return o; }
};

// Enqueue every kernel of the path on `st`.  All pointers are device pointers.
// scan_series/scan_times != NULL: analyze() mode — the front end scans a recorded series
// (k_scan_series) instead of integrating (k_integrate).
static cudaError_t enqueue_path(const nbt_config* cfg, const nbt_inputs* in, const nbt_outputs* out,
                                double* rv_sm_ws, cudaStream_t st, const double* scan_series = nullptr,
                                const double* scan_times = nullptr, double scan_dt = 0.0,
                                cudaEvent_t after_fit = nullptr, int n_ls_chunks = 0,
                                const int32_t* ls_bounds = nullptr, cudaEvent_t* ls_events = nullptr) {
    const int B = cfg->n_systems, Np = cfg->n_bodies - 1, NO = cfg->n_outputs;
    if (B == 0) return cudaSuccess;
    cudaError_t e;
    const bool tim = (cfg->flags & NBT_F_TIMING) != 0;
    if (tim) { g_timing.have = true; g_timing.device = cfg->device; }
    timing_mark(0, tim, st);
    if (scan_series == nullptr) {
        nbt_run_args A;
        memset(&A, 0, sizeof A);
        A.params = in->params; A.substeps = (cfg->substeps > 0) ? nullptr : in->substeps; A.order = in->order;
        A.tt_offsets = in->tt_offsets;
        A.n_days_sys = in->n_days_sys; A.n_outputs_sys = in->n_outputs_sys;
        A.tt = out->tt; A.n_tt = out->n_tt; A.mean_a = out->mean_a; A.mean_e = out->mean_e;
        A.mean_inc = out->mean_inc; A.mean_omega = out->mean_omega; A.status = out->status;
        A.rv = out->rv; A.orbit_xz = out->orbit_xz; A.series = out->series;
        A.B = B; A.n_outputs = NO; A.substeps_all = cfg->substeps; A.tt_mode = cfg->tt_mode; A.flags = cfg->flags;
        A.orbit_stride = cfg->orbit_stride > 0 ? cfg->orbit_stride : 4;
        A.n_days = cfg->n_days;
        const bool automatic = cfg->scheme == NBT_SCHEME_AUTO;
        if (cfg->scheme == NBT_SCHEME_IAS15 || cfg->scheme == NBT_SCHEME_WHFAST) {
            // the two integrators of the reference itself (REBOUND's default and its 'whfast' preset), one system
            // per thread, no launch order / lanes / scheme mix
            if ((e = launch_special_np(Np, cfg->scheme, A, st)) != cudaSuccess) return e;
        } else {
        A.scheme = nbt_make_scheme(automatic ? NBT_SCHEME_SBABC3 : cfg->scheme);
        int nl = cfg->lanes_systems < 0 ? 0 : cfg->lanes_systems;   // leading launch-order slots on the lanes kernel
        if (nl > B) nl = B;
        if (Np == 1) nl = 0;
        int na = (automatic && cfg->alt_systems > 0) ? cfg->alt_systems : 0;   // trailing slots on the C4 scheme
        if (na > B) na = B;
        A.scheme_alt = automatic ? nbt_make_scheme(NBT_SCHEME_C4) : A.scheme;
        // launch-order segments: [0, nl) on the lanes kernel (one launch per scheme: it takes A.scheme only),
        // [nl, B) on the one-thread-per-system kernel (one launch, the C4 slots in its trailing blocks)
        struct Seg { int first, count, main_count; bool lanes, alt; } seg[3];
        int ns = 0;
        const int la = nl < B - na ? nl : B - na;          // lanes slots on SBABC3
        if (la > 0) seg[ns++] = {0, la, la, true, false};
        if (nl > la) seg[ns++] = {la, nl - la, nl - la, true, true};
        if (B > nl) seg[ns++] = {nl, B - nl, (B - na > nl ? B - na : nl) - nl, false, false};
        // the last segment stays on `st`, the lanes launches fork to the side streams
        const int main_seg = ns - 1;
        SideStream* side = nullptr;
        cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
        if (ns > 1) {
            if ((e = get_side(cfg->device, &side)) != cudaSuccess) return e;
            // fork/join events are per call (concurrent callers on one device share only the side streams)
            if ((e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming)) != cudaSuccess) return e;
            e = cudaEventRecord(fork, st);
        } else e = cudaSuccess;
        int nside = 0;
        for (int i = 0; i < ns && e == cudaSuccess; i++) {
            nbt_run_args As = A;
            if (seg[i].alt) As.scheme = A.scheme_alt;
            if (i == main_seg) { e = launch_integrate_np(Np, As, seg[i].first, seg[i].count, seg[i].main_count, seg[i].lanes, st); continue; }
            cudaStream_t ss = side->st[nside];
            e = cudaStreamWaitEvent(ss, fork, 0);
            if (e == cudaSuccess) e = launch_integrate_np(Np, As, seg[i].first, seg[i].count, seg[i].main_count, seg[i].lanes, ss);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&join[nside], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(join[nside], ss);
            nside++;
        }
        for (int i = 0; i < nside && e == cudaSuccess; i++) e = cudaStreamWaitEvent(st, join[i], 0);   // after st's own launch
        if (fork) cudaEventDestroy(fork);   // destruction is deferred until the events complete
        for (int i = 0; i < 2; i++) if (join[i]) cudaEventDestroy(join[i]);
        if (e != cudaSuccess) return e;
        }
    } else {
        nbt_scan_args S;
        memset(&S, 0, sizeof S);
        S.series = scan_series; S.times = scan_times; S.tt_offsets = in->tt_offsets;
        S.tt = out->tt; S.n_tt = out->n_tt; S.mean_a = out->mean_a; S.mean_e = out->mean_e;
        S.mean_inc = out->mean_inc; S.mean_omega = out->mean_omega; S.status = out->status;
        S.rv = out->rv; S.orbit_xz = out->orbit_xz;
        S.B = B; S.Np = Np; S.n_outputs = NO; S.flags = cfg->flags;
        S.orbit_stride = cfg->orbit_stride > 0 ? cfg->orbit_stride : 4;
        S.rv_scale = 1.496e11 * scan_dt; // simulation.py:186
        if ((e = cudaMemsetAsync(out->status, 0, (size_t)B * 4, st)) != cudaSuccess) return e;
        const long long nt = (long long)B * Np;
        k_scan_series<<<(unsigned)((nt + 127) / 128), 128, 0, st>>>(S);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }

    timing_mark(1, tim, st);
    const long long nw = (long long)B * Np;
    const int fit_grid = (int)((nw * 32 + 127) / 128);
    nbt_fit_args F;
    F.params = in->params; F.tt_offsets = in->tt_offsets; F.tt = out->tt; F.n_tt = out->n_tt;
    F.ttv = out->ttv; F.period = out->period; F.t0 = out->t0; F.ttv_max = out->ttv_max; F.status = out->status;
    F.B = B; F.Np = Np; F.flags = cfg->flags;
    k_fit<<<fit_grid, 128, 0, st>>>(F);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((cfg->flags & NBT_F_LOGLIKE) && in->like && out->loglike) {
        // the sampler's likelihood behind the forward model, on the same stream: tt / ttv never leave the device
        const nbt_like* L = in->like;
        if (L->kind == NBT_LIKE_NESTED)
            k_chi2<<<(B + 127) / 128, 128, 0, st>>>(out->ttv, in->tt_offsets, out->n_tt, B, Np, L->c0, L->c1, L->x, L->y, L->yerr,
                                                     L->n_data, L->loss, out->loglike);
        else
            k_ttvfit<<<(B + 127) / 128, 128, 0, st>>>(out->tt, in->tt_offsets, out->n_tt, B, Np, L->has_data, L->tc, L->tc_err,
                                                       L->orbit, L->n_data, out->loglike);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    timing_mark(2, tim, st);
    // everything but the periodograms and RV reductions is final here: the host path starts copying
    // those products back on a second stream while k_ls_oc runs
    if (after_fit && (e = cudaEventRecord(after_fit, st)) != cudaSuccess) return e;

    if ((cfg->flags & NBT_F_LS_OC) && out->ls_oc_power && in->ls_offsets) {
        nbt_lsoc_args L;
        L.tt_offsets = in->tt_offsets; L.ls_offsets = in->ls_offsets; L.ttv = out->ttv; L.n_tt = out->n_tt;
        L.power = out->ls_oc_power; L.nfreq = out->ls_oc_nfreq; L.B = B; L.Np = Np; L.flags = cfg->flags;
        L.spp = cfg->ls_samples_per_peak; L.fmin = cfg->ls_oc_minfreq; L.fmax = cfg->ls_oc_maxfreq;
        // host path: a few launches over ranges of series, an event after each, so that the periodograms of one
        // range go back to the host while the next range is computed (ls_bounds[0..n_ls_chunks])
        // warps per series from the largest transit capacity (cfg.tt_cap_max; unknown = 1, the warp loops)
        {
            const long long nf_max = cfg->tt_cap_max >= 3
                ? nbt_ls_nfreq((double)(cfg->tt_cap_max - 1), cfg->ls_oc_minfreq, cfg->ls_oc_maxfreq, cfg->ls_samples_per_peak) : 0;
            long long nch = (nf_max + NBT_LS_FCHUNK - 1) / NBT_LS_FCHUNK;
            L.nchunk = (int32_t)std::min<long long>(std::max<long long>(nch, 1), 16);
        }
        const int32_t whole[2] = {0, (int32_t)nw};
        const int nc = (n_ls_chunks > 0 && ls_bounds && ls_events) ? n_ls_chunks : 1;
        const int32_t* bounds = nc == n_ls_chunks && ls_bounds ? ls_bounds : whole;
        // (odd ranges on a side stream: the ragged tail of one launch overlaps the next launch)
        const bool piped = nc > 1 && after_fit != nullptr;
        SideStream* side = nullptr;
        if (piped) {
            if ((e = get_side(cfg->device, &side)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(side->st[1], after_fit, 0)) != cudaSuccess) return e;
        }
        for (int c = 0; c < nc; c++) {
            cudaStream_t ls = (piped && (c & 1)) ? side->st[1] : st;
            L.w0 = bounds[c]; L.w1 = bounds[c + 1];
            if (L.w1 > L.w0) {
                k_ls_oc<<<(unsigned)(((long long)(L.w1 - L.w0) * L.nchunk * 32 + 127) / 128), 128, 0, ls>>>(L);
                if ((e = cudaGetLastError()) != cudaSuccess) return e;
            }
            if (nc > 1 && (e = cudaEventRecord(ls_events[c], ls)) != cudaSuccess) return e;
        }
        if (piped)
            for (int c = 1; c < nc; c += 2)
                if ((e = cudaStreamWaitEvent(st, ls_events[c], 0)) != cudaSuccess) return e;
        if ((cfg->flags & NBT_F_LS_PEAKS) && out->ls_oc_peaks) {
            k_ls_peaks<<<(unsigned)((nw * 32 + 127) / 128), 128, 0, st>>>(in->tt_offsets, in->ls_offsets, out->n_tt, out->ls_oc_nfreq,
                                                                           out->ls_oc_power, (int)nw, cfg->ls_samples_per_peak,
                                                                           cfg->ls_oc_minfreq, out->ls_oc_peaks);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
    }
    timing_mark(3, tim, st);
    if ((cfg->flags & NBT_F_RV) && out->rv && rv_sm_ws && (out->rv_max || ((cfg->flags & NBT_F_LS_RV) && out->ls_rv_power))) {
        const int n = NO - 1;
        dim3 tb(32, 8), tg((B + 31) / 32, (n + 31) / 32);
        k_transpose<<<tg, tb, 0, st>>>(out->rv, rv_sm_ws, n, B);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (out->rv_max) {
            k_rv_max<<<(int)(((long long)B * 32 + 127) / 128), 128, 0, st>>>(rv_sm_ws, B, n, out->rv_max);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
        if ((cfg->flags & NBT_F_LS_RV) && out->ls_rv_power) {
            // times[1:] is uniform with spacing `step`; the periodogram is invariant under a time
            // shift, so t_i = i*step and nu = f*step cycles per sample (simulation.py:186-188).
            // (analyze mode makes the same assumption: cfg.n_days / n_outputs describe m['times'].)
            const double step = cfg->n_days / (double)(NO - 1);
            const double T = cfg->n_days - 1.0 * step;
            const double df = 1.0 / T / (double)cfg->ls_samples_per_peak;
            const long long nf = nbt_ls_nfreq(T, cfg->ls_rv_minfreq, cfg->ls_rv_maxfreq, cfg->ls_samples_per_peak);
            const long long chunks = (nf + 31) / 32;
            const long long warps = (long long)B * chunks;
            k_ls_series<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(nullptr, rv_sm_ws, B, n, cfg->ls_rv_minfreq,
                                                                            df, nf, step, out->ls_rv_power);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
    }
    timing_mark(4, tim, st);
    return cudaSuccess;
}

extern "C" {

int nbt_abi_version(void) { return NBT_ABI_VERSION; }

int nbt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* nbt_last_error(void) { return g_err.c_str(); }

int64_t nbt_ls_nfreq(double baseline, double minfreq, double maxfreq, int spp) {
    if (!(baseline > 0.0) || spp <= 0) return 0;
    const double df = 1.0 / baseline / (double)spp;
    return 1 + (int64_t)rint((maxfreq - minfreq) / df);
}

// per-system output grid (inputs.n_days_sys / n_outputs_sys) or the uniform one of cfg
static inline double grid_days(const nbt_config* cfg, const nbt_inputs* in, int b) {
    return (in && in->n_days_sys) ? in->n_days_sys[b] : cfg->n_days;
}
static inline int grid_outputs(const nbt_config* cfg, const nbt_inputs* in, int b) {
    return (in && in->n_outputs_sys) ? in->n_outputs_sys[b] : cfg->n_outputs;
}
static int check_grids(const nbt_config* cfg, const nbt_inputs* in) {
    if (!in) return 0;
    if ((in->n_days_sys == nullptr) != (in->n_outputs_sys == nullptr))
        return nbt_fail(-14, "inputs.n_days_sys and inputs.n_outputs_sys go together");
    if (in->n_days_sys && (cfg->flags & (NBT_F_RV | NBT_F_ORBIT | NBT_F_SERIES | NBT_F_LS_RV)))
        return nbt_fail(-14, "per-system output grids cannot be combined with the [n_outputs][B] products (RV, orbit, series)");
    return 0;
}

int nbt_plan_tt(const nbt_config* cfg, const nbt_inputs* in, int64_t* tt_offsets) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!in || !in->params || !tt_offsets) return nbt_fail(-10, "nbt_plan_tt: NULL argument");
    if ((rc = check_grids(cfg, in)) != 0) return rc;
    const double* params = in->params;
    const int B = cfg->n_systems, N = cfg->n_bodies, Np = N - 1;
    int64_t off = 0, capmax = 1;
    for (int b = 0; b < B; b++) {
        const double nd = grid_days(cfg, in, b);
        const int no = grid_outputs(cfg, in, b);
        if (!(nd > 0.0) || no < 2) return nbt_fail(-14, "system %d: n_days must be > 0 and n_outputs >= 2", b);
        for (int j = 0; j < Np; j++) {
            tt_offsets[(size_t)b * Np + j] = off;
            const double P = params[((size_t)b * N + j + 1) * NBT_NPAR + NBT_P_P];
            int64_t cap = 3;
            if (P > 0.0 && std::isfinite(P)) {
                double c = ceil(1.05 * nd / P) + 3.0;
                if (c > (double)no) c = (double)no; // at most one transit per interval
                cap = (int64_t)c;
            }
            if ((cfg->flags & NBT_F_PLANET1_ONLY) && j > 0) cap = 1;
            off += cap;
            capmax = std::max(capmax, cap);
        }
    }
    tt_offsets[(size_t)B * Np] = off;
    if (capmax > INT32_MAX) return nbt_fail(-11, "tt capacity overflow");
    return (int)capmax;
}

int nbt_plan_order(const nbt_config* cfg, const nbt_inputs* in, int32_t* order) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!in || !order) return nbt_fail(-10, "nbt_plan_order: NULL argument");
    const double* params = in->params;
    const int32_t* substeps = in->substeps;
    const int32_t* nos = in->n_outputs_sys;
    const int B = cfg->n_systems, N = cfg->n_bodies;
    std::iota(order, order + B, 0);
    // key 1: sub-steps, heaviest first (blocks are dispatched in index order, so the long
    // systems start first and a warp's lanes run equal trip counts);
    // key 2: Hill-unstable systems first within a group, so that the lanes that may break (and stop) or
    // fall back to the slow Kepler path share warps instead of each stalling a healthy warp;
    // key 3 (per-system grids): longer grids first, so that the lanes of a warp finish together.
    std::vector<uint8_t> unstable(params ? B : 0);
    if (params)
        for (int b = 0; b < B; b++) unstable[b] = hill_separation(params + (size_t)b * N * NBT_NPAR, N) < 4.0 ? 1 : 0;
    // negative counts (NBT_SCHEME_AUTO: one C4 step per output) are the lightest systems: they sort last,
    // which also keeps the two schemes in separate warps but for one
    std::stable_sort(order, order + B, [&](int32_t a, int32_t b) {
        const int ka = substeps ? substeps[a] : 1, kb = substeps ? substeps[b] : 1;
        if (ka != kb) return ka > kb;
        if (params && unstable[a] != unstable[b]) return unstable[a] > unstable[b];
        if (nos && nos[a] != nos[b]) return nos[a] > nos[b];
        return false;
    });
    // The multi-step groups (k >= 2) are the critical path of a launch, and a warp's slow-path excursions
    // add up over its lanes (measured: 32 unstable k = 3 systems in one warp run 73 ms, each alone 51 ms, a
    // warp of stable ones 53 ms): there the unstable systems are spread evenly over the group's warps instead.
    if (params && substeps && !nos) {
        std::vector<int32_t> tmp;
        for (int g0 = 0; g0 < B;) {
            const int k = substeps[order[g0]];
            int g1 = g0;
            while (g1 < B && substeps[order[g1]] == k) g1++;
            int u = 0;
            while (g0 + u < g1 && unstable[order[g0 + u]]) u++;
            const int n = g1 - g0, s = n - u;
            if (k >= 2 && u > 0 && s > 0) {
                tmp.assign(order + g0, order + g1);
                int iu = 0, is = 0;
                for (int p = 0; p < n; p++) {
                    const bool take_u = (long long)(p + 1) * u / n > (long long)p * u / n;
                    order[g0 + p] = take_u ? tmp[iu++] : tmp[u + is++];
                }
            }
            g0 = g1;
        }
    }
    int nalt = 0;
    if (substeps)
        for (int b = 0; b < B; b++) nalt += substeps[b] < 0 ? 1 : 0;
    return nalt;
}

/* How many leading launch-order slots go to the one-planet-per-lane kernel (nbt_config.lanes_systems).
 * Measured on B200 (profiles/README.md): a lane's stage is latency-bound and no shorter than a
 * thread's for Np = 2, so the lanes kernel only pays where one thread per system is issue-bound or
 * spills: every system when Np >= 5 (k_integrate<5..8> spill), or when Np >= 3 and the batch leaves
 * most SM sub-partitions empty (B <= 4096: single-system latency improves 1.7x at Np = 3). */
int nbt_plan_lanes(const nbt_config* cfg, const nbt_inputs* in) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    (void)in;
    const int B = cfg->n_systems, Np = cfg->n_bodies - 1;
    if (Np <= 1 || B == 0) return 0;
    if (Np >= 5) return B;
    if (Np >= 3 && B <= 4096) return B;
    return 0;
}

int nbt_plan_ls(const nbt_config* cfg, const int64_t* tt_offsets, int64_t* ls_offsets) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!tt_offsets || !ls_offsets) return nbt_fail(-10, "nbt_plan_ls: NULL argument");
    const int64_t n = (int64_t)cfg->n_systems * (cfg->n_bodies - 1);
    int64_t off = 0;
    for (int64_t i = 0; i < n; i++) {
        ls_offsets[i] = off;
        const int64_t cap = tt_offsets[i + 1] - tt_offsets[i];
        off += (cap >= 3) ? nbt_ls_nfreq((double)(cap - 1), cfg->ls_oc_minfreq, cfg->ls_oc_maxfreq, cfg->ls_samples_per_peak) : 0;
    }
    ls_offsets[n] = off;
    return 0;
}

/* Sub-step rule, calibrated against the oracle (tests/calibrate_substeps.py).
 * The binding tolerance is the orbit-averaged e to 1e-8 relative; the composition error scales
 * with (dt_sub / P_innermost)^order, so k = ceil(step * R(scheme) / P_min).  Systems outside the calibrated
 * envelope (NBT_S_UNCALIBRATED: close pairs, e > 0.1, heavy planets) get the same rule — their error is set by
 * chaotic amplification, which no a-priori rule predicts; engine.run_verified checks them by step doubling. */
int nbt_plan_substeps(const nbt_config* cfg, const nbt_inputs* in, int32_t* substeps) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!in || !in->params || !substeps) return nbt_fail(-10, "nbt_plan_substeps: NULL argument");
    if ((rc = check_grids(cfg, in)) != 0) return rc;
    const double* params = in->params;
    /* P_min / dt_sub per scheme id (WH2, Y4, SABA4, M64, C4, SBABC3, AUTO, IAS15 (unused), WHFAST (dt = 1e-3 d)) */
    static const double R[9] = {8000.0, 400.0, 800.0, 57.5, 250.0, 57.5, 57.5, 0.0, 0.0};
    const int B = cfg->n_systems, N = cfg->n_bodies;
    const bool automatic = cfg->scheme == NBT_SCHEME_AUTO;
    for (int b = 0; b < B; b++) {
        const double* par = params + (size_t)b * N * NBT_NPAR;
        const double step = grid_days(cfg, in, b) / (double)(grid_outputs(cfg, in, b) - 1);
        if (cfg->scheme == NBT_SCHEME_WHFAST) {   /* full steps of REBOUND's default dt = 1e-3 d per output interval */
            double k = floor(step / 1e-3 * (1.0 + 1e-12));
            if (k > 1e9) k = 1e9;
            substeps[b] = (int32_t)k;
            continue;
        }
        if (cfg->scheme == NBT_SCHEME_IAS15) { substeps[b] = 1; continue; }
        double pmin = INFINITY;
        for (int j = 1; j < N; j++) {
            const double P = par[j * NBT_NPAR + NBT_P_P];
            if (P > 0.0 && P < pmin) pmin = P;
        }
        double k = std::isfinite(pmin) ? ceil(step * R[cfg->scheme] / pmin) : 1.0;
        if (!(k >= 1.0)) k = 1.0;
        if (k > 65536.0) k = 65536.0;
        /* AUTO: a C4 step costs ~0.7 of an SBABC3 step and needs 3.8x the resolution, so it only pays
         * where a single step per output interval is already fine enough for it; Hill-unstable systems
         * (which may break, and which nbt_plan_order groups by sub-step count) keep SBABC3 */
        if (automatic && k == 1.0 && std::isfinite(pmin) && step * R[NBT_SCHEME_C4] <= pmin &&
            hill_separation(par, N) >= 4.0) k = -1.0;
        substeps[b] = (int32_t)k;
    }
    return 0;
}

int64_t nbt_count_steps(const nbt_config* cfg, const nbt_inputs* in) {
    if (check_cfg(cfg)) return -1;
    const int32_t* substeps = in ? in->substeps : nullptr;
    const int64_t S = scheme_stages(cfg->scheme);
    const int64_t S_alt = cfg->scheme == NBT_SCHEME_AUTO ? scheme_stages(NBT_SCHEME_C4) : S;
    int64_t tot = 0;
    for (int b = 0; b < cfg->n_systems; b++) {
        const int64_t per = (int64_t)grid_outputs(cfg, in, b) - 1;
        int64_t k = (cfg->substeps > 0 || !substeps) ? std::max(cfg->substeps, 1) : substeps[b];
        tot += per * (k < 0 ? -k * S_alt : k * S);
    }
    return tot;
}

/* nbt_plan: everything a caller plans before nbt_run, in one pass over the parameter block on n_threads host threads
 * (the separate nbt_plan_* helpers above, same results bit for bit — tests/test_host_logic.py).  A sampler or the
 * training-set generator plans every call anew (parameters change from call to call), so this sits on the end-to-end
 * path: 70 000 two-planet systems take ~1.5 ms here against ~10 ms through the five helpers. */
}  // extern "C" (a template cannot have C linkage)
template <class F>
static void plan_parallel(int B, int n_threads, F&& body) {
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    T = std::max(1, std::min(T, std::min(32, (B + 4095) / 4096)));
    if (T == 1) { body(0, 0, B); return; }
    std::vector<std::thread> th;
    th.reserve(T);
    for (int t = 0; t < T; t++) {
        const int lo = (int)((int64_t)B * t / T), hi = (int)((int64_t)B * (t + 1) / T);
        th.emplace_back([&body, t, lo, hi] { body(t, lo, hi); });
    }
    for (auto& x : th) x.join();
}
extern "C" {

int nbt_plan(const nbt_config* cfg, const nbt_inputs* in, int32_t sort_mode, int32_t* substeps_out, int32_t* order,
             int64_t* tt_offsets, int64_t* ls_offsets, int64_t* summary, int32_t n_threads) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!in || !in->params || !tt_offsets || !summary) return nbt_fail(-10, "nbt_plan: NULL argument");
    if ((rc = check_grids(cfg, in)) != 0) return rc;
    const int B = cfg->n_systems, N = cfg->n_bodies, Np = N - 1;
    const double* params = in->params;
    const bool uniform = cfg->substeps > 0;
    const int32_t* sub_in = uniform ? nullptr : in->substeps;
    if (!uniform && !sub_in && !substeps_out) return nbt_fail(-10, "nbt_plan: substeps_out is required when neither cfg.substeps nor inputs.substeps is set");
    if (sort_mode && !order) return nbt_fail(-10, "nbt_plan: order is required when sort_mode != 0");
    for (int b = 0; b < B; b++) {
        if (!(grid_days(cfg, in, b) > 0.0) || grid_outputs(cfg, in, b) < 2)
            return nbt_fail(-14, "system %d: n_days must be > 0 and n_outputs >= 2", b);
        if (!in->n_days_sys) break;
    }
    // pass 1 (parallel): sub-step counts, Hill flags, CSR capacities (stored as counts at [i + 1]), step totals
    if (!uniform && !sub_in) {
        nbt_inputs tmp = *in;
        tmp.substeps = nullptr;
        std::vector<int> rcs(64, 0);
        plan_parallel(B, n_threads, [&](int t, int lo, int hi) {
            nbt_config c = *cfg;
            nbt_inputs i2 = tmp;
            c.n_systems = hi - lo;
            i2.params = params + (size_t)lo * N * NBT_NPAR;
            if (tmp.n_days_sys) { i2.n_days_sys = tmp.n_days_sys + lo; i2.n_outputs_sys = tmp.n_outputs_sys + lo; }
            rcs[t] = nbt_plan_substeps(&c, &i2, substeps_out + lo);
        });
        for (int r : rcs) if (r) return r;
        sub_in = substeps_out;
    }
    const bool want_ls = ls_offsets != nullptr;
    const bool need_hill = order != nullptr && sort_mode == 1;
    std::vector<uint8_t> unstable(need_hill ? B : 0);
    const int64_t S = scheme_stages(cfg->scheme);
    const int64_t S_alt = cfg->scheme == NBT_SCHEME_AUTO ? scheme_stages(NBT_SCHEME_C4) : S;
    std::vector<int64_t> steps_part(64, 0), nalt_part(64, 0), capmax_part(64, 1);
    tt_offsets[0] = 0;
    if (want_ls) ls_offsets[0] = 0;
    plan_parallel(B, n_threads, [&](int t, int lo, int hi) {
        int64_t steps = 0, nalt = 0, capmax = 1;
        for (int b = lo; b < hi; b++) {
            const double* par = params + (size_t)b * N * NBT_NPAR;
            const double nd = grid_days(cfg, in, b);
            const int no = grid_outputs(cfg, in, b);
            for (int j = 0; j < Np; j++) {
                const double P = par[(j + 1) * NBT_NPAR + NBT_P_P];
                int64_t cap = 3;
                if (P > 0.0 && std::isfinite(P)) {
                    double c = ceil(1.05 * nd / P) + 3.0;
                    if (c > (double)no) c = (double)no;
                    cap = (int64_t)c;
                }
                if ((cfg->flags & NBT_F_PLANET1_ONLY) && j > 0) cap = 1;
                tt_offsets[(size_t)b * Np + j + 1] = cap;
                capmax = std::max(capmax, cap);
                if (want_ls)
                    ls_offsets[(size_t)b * Np + j + 1] =
                        (cap >= 3) ? nbt_ls_nfreq((double)(cap - 1), cfg->ls_oc_minfreq, cfg->ls_oc_maxfreq, cfg->ls_samples_per_peak) : 0;
            }
            const int64_t k = sub_in ? sub_in[b] : std::max(cfg->substeps, 1);
            steps += ((int64_t)no - 1) * (k < 0 ? -k * S_alt : k * S);
            nalt += k < 0 ? 1 : 0;
            if (need_hill) unstable[b] = hill_separation(par, N) < 4.0 ? 1 : 0;
        }
        steps_part[t] = steps; nalt_part[t] = nalt; capmax_part[t] = capmax;
    });
    int64_t n_steps = 0, nalt = 0, capmax = 1;
    for (int t = 0; t < 64; t++) { n_steps += steps_part[t]; nalt += nalt_part[t]; capmax = std::max(capmax, capmax_part[t]); }
    if (capmax > INT32_MAX) return nbt_fail(-11, "tt capacity overflow");
    // pass 2: capacities -> offsets
    const int64_t nsp = (int64_t)B * Np;
    for (int64_t i = 0; i < nsp; i++) tt_offsets[i + 1] += tt_offsets[i];
    if (want_ls)
        for (int64_t i = 0; i < nsp; i++) ls_offsets[i + 1] += ls_offsets[i];
    // pass 3: launch order — the stable sort of nbt_plan_order as a counting sort over (sub-steps descending, unstable first)
    const bool marked = cfg->scheme == NBT_SCHEME_AUTO && nalt > 0;
    const bool adaptive = cfg->scheme == NBT_SCHEME_IAS15 || cfg->scheme == NBT_SCHEME_WHFAST;
    bool identity = true;
    if (order) {
        if (sort_mode == 1 && (B > 32 || marked) && !adaptive) {
            int kmin = INT32_MAX, kmax = INT32_MIN;
            for (int b = 0; b < B; b++) { const int k = sub_in ? sub_in[b] : 1; kmin = std::min(kmin, k); kmax = std::max(kmax, k); }
            if (B == 0) { kmin = kmax = 1; }
            const int64_t nb = ((int64_t)kmax - kmin + 1) * 2;
            const int32_t* nos = in->n_outputs_sys;
            if (nb <= (1 << 22)) {
                std::vector<int32_t> start((size_t)nb + 1, 0);
                auto bucket = [&](int b) { const int k = sub_in ? sub_in[b] : 1; return (size_t)(kmax - k) * 2 + (1 - unstable[b]); };
                for (int b = 0; b < B; b++) start[bucket(b) + 1]++;
                for (int64_t i = 0; i < nb; i++) start[i + 1] += start[i];
                std::vector<int32_t> pos(start.begin(), start.end() - 1);
                for (int b = 0; b < B; b++) order[pos[bucket(b)]++] = b;
                if (nos)
                    for (int64_t i = 0; i < nb; i++)
                        if (start[i + 1] - start[i] > 1)
                            std::stable_sort(order + start[i], order + start[i + 1], [&](int32_t a, int32_t c) { return nos[a] > nos[c]; });
            } else {   // pathological range of counts: the comparison sort
                std::iota(order, order + B, 0);
                std::stable_sort(order, order + B, [&](int32_t a, int32_t c) {
                    const int ka = sub_in ? sub_in[a] : 1, kc = sub_in ? sub_in[c] : 1;
                    if (ka != kc) return ka > kc;
                    if (unstable[a] != unstable[c]) return unstable[a] > unstable[c];
                    if (nos && nos[a] != nos[c]) return nos[a] > nos[c];
                    return false;
                });
            }
            if (sub_in && !nos) {   // spread the unstable systems of the multi-step groups over the group's warps (see nbt_plan_order)
                std::vector<int32_t> tmp;
                for (int g0 = 0; g0 < B;) {
                    const int k = sub_in[order[g0]];
                    int g1 = g0;
                    while (g1 < B && sub_in[order[g1]] == k) g1++;
                    int u = 0;
                    while (g0 + u < g1 && unstable[order[g0 + u]]) u++;
                    const int n = g1 - g0, s = n - u;
                    if (k >= 2 && u > 0 && s > 0) {
                        tmp.assign(order + g0, order + g1);
                        int iu = 0, is = 0;
                        for (int p = 0; p < n; p++) {
                            const bool take_u = (long long)(p + 1) * u / n > (long long)p * u / n;
                            order[g0 + p] = take_u ? tmp[iu++] : tmp[u + is++];
                        }
                    }
                    g0 = g1;
                }
            }
        } else if (marked) {   // unsorted: the C4 systems still go last (scheme uniform per block)
            int w = 0;
            for (int b = 0; b < B; b++) if (sub_in[b] >= 0) order[w++] = b;
            for (int b = 0; b < B; b++) if (sub_in[b] < 0) order[w++] = b;
        } else {
            std::iota(order, order + B, 0);
        }
        for (int b = 0; b < B && identity; b++) identity = order[b] == b;
    }
    nbt_config c2 = *cfg;
    const int lanes = nbt_plan_lanes(&c2, in);
    summary[0] = n_steps;
    summary[1] = capmax;
    summary[2] = cfg->scheme == NBT_SCHEME_AUTO ? nalt : 0;
    summary[3] = identity ? 1 : 0;
    summary[4] = lanes;
    return 0;
}

int nbt_uncalibrated(const double* params_one, int n_bodies) {
    if (!params_one || n_bodies < 2 || n_bodies > NBT_MAX_BODIES) return 0;
    return ::nbt_uncalibrated_impl(params_one, n_bodies) ? 1 : 0;
}

static int run_impl(const nbt_config* cfg, const nbt_inputs* in, nbt_outputs* out, void* cuda_stream,
                    const double* scan_series, const double* scan_times, double scan_dt) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!in || !out) return nbt_fail(-10, "nbt_run: NULL inputs/outputs");
    const bool scan = scan_series != nullptr;
    if (scan && !scan_times) return nbt_fail(-10, "nbt_analyze: times is required");
    if (!in->params || !in->tt_offsets) return nbt_fail(-10, "nbt_run: params and tt_offsets are required");
    if (!out->status) return nbt_fail(-10, "nbt_run: status is required");
    if ((cfg->flags & NBT_F_DEVICE_PTRS) && (!out->tt || !out->ttv || !out->n_tt || !out->period || !out->t0 || !out->ttv_max))
        return nbt_fail(-10, "nbt_run: with device pointers tt, ttv, n_tt, period, t0, ttv_max are required (the kernels write them)");
    if ((rc = check_grids(cfg, in)) != 0) return rc;
    if (scan && in->n_days_sys) return nbt_fail(-14, "nbt_analyze: per-system grids are not supported");
    if (cfg->flags & NBT_F_LOGLIKE) {
        const nbt_like* L = in->like;
        if (!L || !out->loglike) return nbt_fail(-15, "nbt_run: NBT_F_LOGLIKE needs inputs.like and outputs.loglike");
        if (L->n_data < 0 || (L->kind != NBT_LIKE_NESTED && L->kind != NBT_LIKE_TTVFIT))
            return nbt_fail(-15, "nbt_run: bad nbt_like (kind %d, n_data %d)", L->kind, L->n_data);
        if (L->kind == NBT_LIKE_NESTED && (!L->c0 || !L->c1 || (L->n_data > 0 && (!L->x || !L->y || !L->yerr))))
            return nbt_fail(-15, "nbt_run: NBT_LIKE_NESTED needs c0, c1, x, y, yerr");
        if (L->kind == NBT_LIKE_NESTED && L->loss != NBT_LOSS_LINEAR && L->loss != NBT_LOSS_SOFT_L1)
            return nbt_fail(-15, "nbt_run: unknown loss %d", L->loss);
        if (L->kind == NBT_LIKE_TTVFIT && (L->n_data < 1 || !L->has_data || !L->tc || !L->tc_err || !L->orbit))
            return nbt_fail(-15, "nbt_run: NBT_LIKE_TTVFIT needs has_data, tc, tc_err, orbit and n_data >= 1");
    }
    if ((cfg->flags & NBT_F_LS_PEAKS) && (!(cfg->flags & NBT_F_LS_OC) || !out->ls_oc_peaks || !in->ls_offsets))
        return nbt_fail(-16, "nbt_run: NBT_F_LS_PEAKS needs NBT_F_LS_OC, inputs.ls_offsets and outputs.ls_oc_peaks");
    if ((cfg->flags & NBT_F_LS_PEAKS) && (cfg->flags & NBT_F_DEVICE_PTRS) && (!out->ls_oc_power || !out->ls_oc_nfreq))
        return nbt_fail(-16, "nbt_run: with device pointers NBT_F_LS_PEAKS needs ls_oc_power and ls_oc_nfreq as scratch");
    if ((cfg->flags & NBT_F_MEANS) && (!out->mean_a || !out->mean_e || !out->mean_inc))
        return nbt_fail(-10, "nbt_run: NBT_F_MEANS needs mean_a, mean_e, mean_inc");
    if (!scan && cfg->substeps <= 0 && !in->substeps) return nbt_fail(-12, "nbt_run: cfg.substeps <= 0 needs inputs.substeps");
    if (!scan && cfg->scheme == NBT_SCHEME_AUTO && cfg->alt_systems > 0) {
        if (cfg->alt_systems > cfg->n_systems || cfg->substeps > 0 || !in->substeps)
            return nbt_fail(-13, "nbt_run: cfg.alt_systems needs inputs.substeps and <= n_systems slots");
        if (!(cfg->flags & NBT_F_DEVICE_PTRS))   // host arrays: every C4 slot must be one the planner marked
            for (int g = cfg->n_systems - cfg->alt_systems; g < cfg->n_systems; g++)
                if (in->substeps[in->order ? in->order[g] : g] >= 0)
                    return nbt_fail(-13, "nbt_run: launch slot %d is inside cfg.alt_systems but its sub-step count is not negative", g);
    }
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_run: no CUDA device (there is no CPU fallback)");
    const int B = cfg->n_systems, N = cfg->n_bodies, Np = N - 1, NO = cfg->n_outputs;
    if (B == 0) return 0;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : cudaStreamPerThread;
    char* arena = nullptr;
    double* rv_ws = nullptr;
    cudaEvent_t ev_fit = nullptr;
    cudaEvent_t ev_ls[NBT_LS_CHUNKS] = {};
    NBT_CUDA(cudaSetDevice(cfg->device));
    ensure_pool(cfg->device);
    {
        const bool need_rv_ws = (cfg->flags & NBT_F_RV) && out->rv && (out->rv_max || ((cfg->flags & NBT_F_LS_RV) && out->ls_rv_power));
        const size_t rv_elems = (size_t)(NO - 1) * B;
        if (cfg->flags & NBT_F_DEVICE_PTRS) {
            if (need_rv_ws) NBT_CUDA(cudaMallocAsync((void**)&rv_ws, rv_elems * 8, st));
            NBT_CUDA(enqueue_path(cfg, in, out, rv_ws, st, scan_series, scan_times, scan_dt));
            if (rv_ws) NBT_CUDA(cudaFreeAsync(rv_ws, st));
            rv_ws = nullptr;
            goto done;
        }
        // host pointers: stage through one stream-ordered arena
        const int64_t nsp = (int64_t)B * Np;
        const int64_t ntt = in->tt_offsets[nsp];
        const int64_t nls = (in->ls_offsets && (cfg->flags & NBT_F_LS_OC)) ? in->ls_offsets[nsp] : 0;
        const int W = 3 + 10 * Np;
        const int stride = cfg->orbit_stride > 0 ? cfg->orbit_stride : 4;
        const size_t nq = (size_t)(NO + stride - 1) / stride;
        const int64_t nf_rv = (cfg->flags & NBT_F_LS_RV)
                                  ? nbt_ls_nfreq(cfg->n_days - cfg->n_days / (double)(NO - 1), cfg->ls_rv_minfreq,
                                                 cfg->ls_rv_maxfreq, cfg->ls_samples_per_peak) : 0;
        Arena ar;
        const int fl0 = cfg->flags;
        struct Buf { size_t off, bytes; const void* h_in; void* h_out; };
        auto mk = [&](size_t bytes, const void* hin, void* hout, bool enabled) {
            Buf b{0, 0, nullptr, nullptr};
            if (enabled && bytes > 0) { b.off = ar.reserve(bytes); b.bytes = bytes; b.h_in = hin; b.h_out = hout; }
            return b;
        };
        const int fl = cfg->flags;
        const nbt_like* HL = (fl0 & NBT_F_LOGLIKE) ? in->like : nullptr;
        const bool nested = HL && HL->kind == NBT_LIKE_NESTED, ttvfit = HL && HL->kind == NBT_LIKE_TTVFIT;
        const size_t nd = HL ? (size_t)HL->n_data : 0;
        Buf b_params = mk((size_t)B * N * NBT_NPAR * 8, in->params, nullptr, true);
        Buf b_nds = mk((size_t)B * 8, in->n_days_sys, nullptr, !scan && in->n_days_sys != nullptr);
        Buf b_nos = mk((size_t)B * 4, in->n_outputs_sys, nullptr, !scan && in->n_outputs_sys != nullptr);
        Buf b_lc0 = mk((size_t)B * 8, HL ? HL->c0 : nullptr, nullptr, nested);
        Buf b_lc1 = mk((size_t)B * 8, HL ? HL->c1 : nullptr, nullptr, nested);
        Buf b_lx = mk(nd * 8, HL ? HL->x : nullptr, nullptr, nested);
        Buf b_ly = mk(nd * 8, HL ? HL->y : nullptr, nullptr, nested);
        Buf b_le = mk(nd * 8, HL ? HL->yerr : nullptr, nullptr, nested);
        Buf b_lhas = mk((size_t)Np * 4, HL ? HL->has_data : nullptr, nullptr, ttvfit);
        Buf b_ltc = mk((size_t)Np * nd * 8, HL ? HL->tc : nullptr, nullptr, ttvfit);
        Buf b_lte = mk((size_t)Np * nd * 8, HL ? HL->tc_err : nullptr, nullptr, ttvfit);
        Buf b_lorb = mk(nd * 4, HL ? HL->orbit : nullptr, nullptr, ttvfit);
        Buf b_ll = mk((size_t)B * 8, nullptr, out->loglike, HL != nullptr);
        Buf b_sub = mk((size_t)B * 4, in->substeps, nullptr, !scan && cfg->substeps <= 0);
        Buf b_order = mk((size_t)B * 4, in->order, nullptr, !scan && in->order != nullptr);
        Buf b_times = mk((size_t)NO * 8, scan_times, nullptr, scan);
        Buf b_tto = mk((size_t)(nsp + 1) * 8, in->tt_offsets, nullptr, true);
        Buf b_lso = mk((size_t)(nsp + 1) * 8, in->ls_offsets, nullptr, nls > 0);
        Buf b_tt = mk((size_t)ntt * 8, nullptr, out->tt, true);
        Buf b_ttv = mk((size_t)ntt * 8, nullptr, out->ttv, true);
        Buf b_ntt = mk((size_t)nsp * 4, nullptr, out->n_tt, true);
        Buf b_per = mk((size_t)nsp * 8, nullptr, out->period, true);
        Buf b_t0 = mk((size_t)nsp * 8, nullptr, out->t0, true);
        Buf b_max = mk((size_t)nsp * 8, nullptr, out->ttv_max, true);
        const bool means = (fl & (NBT_F_MEANS | NBT_F_SERIES)) != 0;
        Buf b_ma = mk((size_t)nsp * 8, nullptr, out->mean_a, means && out->mean_a);
        Buf b_me = mk((size_t)nsp * 8, nullptr, out->mean_e, means && out->mean_e);
        Buf b_mi = mk((size_t)nsp * 8, nullptr, out->mean_inc, means && out->mean_inc);
        Buf b_mo = mk((size_t)nsp * 8, nullptr, out->mean_omega, ((fl & NBT_F_MEAN_OMEGA) || (fl & NBT_F_SERIES)) && out->mean_omega);
        Buf b_stat = mk((size_t)B * 4, nullptr, out->status, true);
        Buf b_rv = mk(rv_elems * 8, nullptr, out->rv, (fl & NBT_F_RV) && out->rv);
        Buf b_rvmax = mk((size_t)B * 8, nullptr, out->rv_max, (fl & NBT_F_RV) && out->rv && out->rv_max);
        Buf b_rvws = mk(rv_elems * 8, nullptr, nullptr, need_rv_ws);
        Buf b_orb = mk(nq * B * Np * 2 * 8, nullptr, out->orbit_xz, (fl & NBT_F_ORBIT) && out->orbit_xz);
        Buf b_ser = scan ? mk((size_t)NO * B * W * 8, scan_series, nullptr, true)
                         : mk((size_t)NO * B * W * 8, nullptr, out->series, (fl & NBT_F_SERIES) && out->series);
        const bool peaks = (fl & NBT_F_LS_PEAKS) != 0;
        Buf b_lsp = mk((size_t)nls * 8, nullptr, out->ls_oc_power, nls > 0 && (out->ls_oc_power || peaks));
        Buf b_lsn = mk((size_t)nsp * 4, nullptr, out->ls_oc_nfreq, nls > 0 && (out->ls_oc_nfreq || peaks));
        Buf b_lspk = mk((size_t)nsp * NBT_LS_NPEAKS * 2 * 8, nullptr, out->ls_oc_peaks, nls > 0 && peaks);
        Buf b_lsrv = mk((size_t)B * nf_rv * 8, nullptr, out->ls_rv_power, need_rv_ws && (fl & NBT_F_LS_RV) && out->ls_rv_power);
        NBT_CUDA(cudaMallocAsync((void**)&arena, ar.size ? ar.size : 256, st));
        auto dp = [&](const Buf& b) -> void* { return b.bytes ? (void*)(arena + b.off) : nullptr; };
        Buf* ins[] = {&b_params, &b_sub, &b_order, &b_tto, &b_lso, &b_times, &b_ser, &b_nds, &b_nos,
                      &b_lc0, &b_lc1, &b_lx, &b_ly, &b_le, &b_lhas, &b_ltc, &b_lte, &b_lorb};
        for (Buf* b : ins)
            if (b->bytes && b->h_in) NBT_CUDA(cudaMemcpyAsync(dp(*b), b->h_in, b->bytes, cudaMemcpyHostToDevice, st));
        // NaN-fill the CSR transit buffers (all-ones bytes are a NaN): unused capacity reads as NaN
        if (b_tt.bytes) NBT_CUDA(cudaMemsetAsync(dp(b_tt), 0xFF, b_tt.bytes, st));
        if (b_ttv.bytes) NBT_CUDA(cudaMemsetAsync(dp(b_ttv), 0xFF, b_ttv.bytes, st));
        if (b_lsp.bytes) NBT_CUDA(cudaMemsetAsync(dp(b_lsp), 0xFF, b_lsp.bytes, st));
        nbt_inputs din;
        din.params = (const double*)dp(b_params); din.substeps = (const int32_t*)dp(b_sub);
        din.order = (const int32_t*)dp(b_order); din.tt_offsets = (const int64_t*)dp(b_tto);
        din.ls_offsets = (const int64_t*)dp(b_lso);
        din.n_days_sys = (const double*)dp(b_nds); din.n_outputs_sys = (const int32_t*)dp(b_nos);
        nbt_like dlike;
        memset(&dlike, 0, sizeof dlike);
        din.like = nullptr;
        if (HL) {
            dlike.kind = HL->kind; dlike.loss = HL->loss; dlike.n_data = HL->n_data;
            dlike.c0 = (const double*)dp(b_lc0); dlike.c1 = (const double*)dp(b_lc1); dlike.x = (const double*)dp(b_lx);
            dlike.y = (const double*)dp(b_ly); dlike.yerr = (const double*)dp(b_le); dlike.has_data = (const int32_t*)dp(b_lhas);
            dlike.tc = (const double*)dp(b_ltc); dlike.tc_err = (const double*)dp(b_lte); dlike.orbit = (const int32_t*)dp(b_lorb);
            din.like = &dlike;
        }
        nbt_outputs dout;
        memset(&dout, 0, sizeof dout);
        dout.tt = (double*)dp(b_tt); dout.ttv = (double*)dp(b_ttv); dout.n_tt = (int32_t*)dp(b_ntt);
        dout.period = (double*)dp(b_per); dout.t0 = (double*)dp(b_t0); dout.ttv_max = (double*)dp(b_max);
        dout.mean_a = (double*)dp(b_ma); dout.mean_e = (double*)dp(b_me); dout.mean_inc = (double*)dp(b_mi);
        dout.mean_omega = (double*)dp(b_mo); dout.status = (int32_t*)dp(b_stat);
        dout.rv = (double*)dp(b_rv); dout.rv_max = (double*)dp(b_rvmax); dout.orbit_xz = (double*)dp(b_orb);
        dout.series = (double*)dp(b_ser); dout.ls_oc_power = (double*)dp(b_lsp); dout.ls_oc_nfreq = (int32_t*)dp(b_lsn);
        dout.ls_rv_power = (double*)dp(b_lsrv);
        dout.ls_oc_peaks = (double*)dp(b_lspk); dout.loglike = (double*)dp(b_ll);
        SideStream* side = nullptr;
        NBT_CUDA(get_side(cfg->device, &side));
        NBT_CUDA(cudaEventCreateWithFlags(&ev_fit, cudaEventDisableTiming));
        // O-C periodograms (the largest product) in up to NBT_LS_CHUNKS ranges of equal capacity
        int n_ls = 0;
        int32_t ls_bounds[NBT_LS_CHUNKS + 1];
        if (b_lsp.bytes && b_lsp.h_out && nls >= (int64_t)NBT_LS_CHUNKS * (1 << 20) && nsp < INT32_MAX) {
            n_ls = NBT_LS_CHUNKS;
            ls_bounds[0] = 0; ls_bounds[n_ls] = (int32_t)nsp;
            for (int c = 1; c < n_ls; c++)
                ls_bounds[c] = (int32_t)(std::lower_bound(in->ls_offsets, in->ls_offsets + nsp, nls * c / n_ls) - in->ls_offsets);
            for (int c = 0; c < n_ls; c++) NBT_CUDA(cudaEventCreateWithFlags(&ev_ls[c], cudaEventDisableTiming));
        }
        nbt_config hcfg = *cfg;
        if (hcfg.tt_cap_max <= 0) {   // derive the largest per-planet capacity from the host offsets
            int64_t cm = 1;
            for (int64_t i = 0; i < nsp; i++) cm = std::max(cm, in->tt_offsets[i + 1] - in->tt_offsets[i]);
            hcfg.tt_cap_max = (int32_t)std::min<int64_t>(cm, INT32_MAX);
        }
        NBT_CUDA(enqueue_path(&hcfg, &din, &dout, (double*)dp(b_rvws), st, scan ? (const double*)dp(b_ser) : nullptr,
                              (const double*)dp(b_times), scan_dt, ev_fit, n_ls, ls_bounds, ev_ls));
        // products that are final after k_fit go back on the side stream, overlapping k_ls_oc / the RV stages
        Buf* early[] = {&b_ll, &b_tt, &b_ttv, &b_ntt, &b_per, &b_t0, &b_max, &b_ma, &b_me, &b_mi, &b_mo, &b_stat, &b_orb, &b_ser};
        Buf* late[] = {&b_rv, &b_rvmax, &b_lsp, &b_lsn, &b_lsrv, &b_lspk};
        NBT_CUDA(cudaStreamWaitEvent(side->st[0], ev_fit, 0));
        for (Buf* b : early)
            if (b->bytes && b->h_out) NBT_CUDA(cudaMemcpyAsync(b->h_out, dp(*b), b->bytes, cudaMemcpyDeviceToHost, side->st[0]));
        for (int c = 0; c < n_ls; c++) {   // periodogram ranges follow their launches on the side stream
            const int64_t lo = in->ls_offsets[ls_bounds[c]], hi = in->ls_offsets[ls_bounds[c + 1]];
            NBT_CUDA(cudaStreamWaitEvent(side->st[0], ev_ls[c], 0));
            if (hi > lo) NBT_CUDA(cudaMemcpyAsync((char*)b_lsp.h_out + lo * 8, (char*)dp(b_lsp) + lo * 8, (size_t)(hi - lo) * 8,
                                                  cudaMemcpyDeviceToHost, side->st[0]));
        }
        for (Buf* b : late)
            if (b->bytes && b->h_out && !(n_ls > 0 && b == &b_lsp))
                NBT_CUDA(cudaMemcpyAsync(b->h_out, dp(*b), b->bytes, cudaMemcpyDeviceToHost, st));
        NBT_CUDA(cudaStreamSynchronize(side->st[0]));
        NBT_CUDA(cudaStreamSynchronize(st));
    }
done:
    if (ev_fit) cudaEventDestroy(ev_fit);
    for (int c = 0; c < NBT_LS_CHUNKS; c++) if (ev_ls[c]) cudaEventDestroy(ev_ls[c]);
    if (rc != 0 && arena) cudaDeviceSynchronize();     // error path: nothing may still touch the arena
    if (arena) cudaFreeAsync(arena, st);
    if (rv_ws) cudaFreeAsync(rv_ws, st);
    if (rc != 0) cudaGetLastError();
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return rc;
}

int nbt_run(const nbt_config* cfg, const nbt_inputs* in, nbt_outputs* out, void* cuda_stream) {
    return run_impl(cfg, in, out, cuda_stream, nullptr, nullptr, 0.0);
}

int nbt_analyze(const nbt_config* cfg, const double* series, const double* times, double dt,
                const nbt_inputs* in, nbt_outputs* out, void* cuda_stream) {
    if (!series) return nbt_fail(-10, "nbt_analyze: series is required");
    return run_impl(cfg, in, out, cuda_stream, series, times, dt);
}

int nbt_lomb_scargle(const double* t, const double* y, int64_t n_series, int64_t n, double f0, double df,
                     int64_t n_freq, double* power, int device, int device_ptrs, void* cuda_stream) {
    int rc = 0;
    if (!y || !power) return nbt_fail(-10, "nbt_lomb_scargle: NULL y/power");
    if (n_series < 0 || n < 1 || n > INT32_MAX || n_freq < 0) return nbt_fail(-11, "nbt_lomb_scargle: bad sizes");
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_lomb_scargle: no CUDA device (there is no CPU fallback)");
    if (n_series == 0 || n_freq == 0) return 0;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : cudaStreamPerThread;
    double *dt = nullptr, *dy = nullptr, *dpw = nullptr;
    const size_t ny = (size_t)n_series * n * 8, npw = (size_t)n_series * n_freq * 8;
    NBT_CUDA(cudaSetDevice(device));
    ensure_pool(device);
    {
        const double *kt = t, *ky = y;
        double* kp = power;
        if (!device_ptrs) {
            NBT_CUDA(cudaMallocAsync((void**)&dy, ny, st));
            NBT_CUDA(cudaMallocAsync((void**)&dpw, npw, st));
            NBT_CUDA(cudaMemcpyAsync(dy, y, ny, cudaMemcpyHostToDevice, st));
            if (t) {
                NBT_CUDA(cudaMallocAsync((void**)&dt, ny, st));
                NBT_CUDA(cudaMemcpyAsync(dt, t, ny, cudaMemcpyHostToDevice, st));
            }
            kt = dt; ky = dy; kp = dpw;
        }
        const long long chunks = (n_freq + 31) / 32, warps = n_series * chunks;
        k_ls_series<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(kt, ky, n_series, (int)n, f0, df, n_freq, 1.0, kp);
        NBT_CUDA(cudaGetLastError());
        if (!device_ptrs) {
            NBT_CUDA(cudaMemcpyAsync(power, dpw, npw, cudaMemcpyDeviceToHost, st));
            NBT_CUDA(cudaStreamSynchronize(st));
        }
    }
done:
    if (dt) cudaFreeAsync(dt, st);
    if (dy) cudaFreeAsync(dy, st);
    if (dpw) cudaFreeAsync(dpw, st);
    if (rc != 0) cudaGetLastError();
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return rc;
}

int nbt_chi2(const double* ttv, const int64_t* tt_offsets, const int32_t* n_tt, int32_t n_systems,
             int32_t n_planets, const double* c0, const double* c1, const double* x, const double* y,
             const double* yerr, int32_t n_data, int32_t loss, double* loglike, int device, int device_ptrs,
             void* cuda_stream) {
    int rc = 0;
    if (!ttv || !tt_offsets || !n_tt || !c0 || !c1 || !x || !y || !yerr || !loglike)
        return nbt_fail(-10, "nbt_chi2: NULL argument");
    if (loss != NBT_LOSS_LINEAR && loss != NBT_LOSS_SOFT_L1) return nbt_fail(-11, "nbt_chi2: unknown loss %d", loss);
    if (n_systems < 0 || n_planets < 1 || n_data < 0) return nbt_fail(-11, "nbt_chi2: bad sizes");
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_chi2: no CUDA device (there is no CPU fallback)");
    if (n_systems == 0) return 0;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : cudaStreamPerThread;
    char* arena = nullptr;
    NBT_CUDA(cudaSetDevice(device));
    ensure_pool(device);
    {
        const int64_t nsp = (int64_t)n_systems * n_planets;
        const double *k_ttv = ttv, *k_c0 = c0, *k_c1 = c1, *k_x = x, *k_y = y, *k_e = yerr;
        const int64_t* k_off = tt_offsets; const int32_t* k_n = n_tt; double* k_ll = loglike;
        if (!device_ptrs) {
            const int64_t ntt = tt_offsets[nsp];
            Arena ar;
            const size_t o_ttv = ar.reserve((size_t)ntt * 8), o_off = ar.reserve((size_t)(nsp + 1) * 8),
                         o_n = ar.reserve((size_t)nsp * 4), o_c0 = ar.reserve((size_t)n_systems * 8),
                         o_c1 = ar.reserve((size_t)n_systems * 8), o_x = ar.reserve((size_t)n_data * 8),
                         o_y = ar.reserve((size_t)n_data * 8), o_e = ar.reserve((size_t)n_data * 8),
                         o_ll = ar.reserve((size_t)n_systems * 8);
            NBT_CUDA(cudaMallocAsync((void**)&arena, ar.size, st));
            NBT_CUDA(cudaMemcpyAsync(arena + o_ttv, ttv, (size_t)ntt * 8, cudaMemcpyHostToDevice, st));
            NBT_CUDA(cudaMemcpyAsync(arena + o_off, tt_offsets, (size_t)(nsp + 1) * 8, cudaMemcpyHostToDevice, st));
            NBT_CUDA(cudaMemcpyAsync(arena + o_n, n_tt, (size_t)nsp * 4, cudaMemcpyHostToDevice, st));
            NBT_CUDA(cudaMemcpyAsync(arena + o_c0, c0, (size_t)n_systems * 8, cudaMemcpyHostToDevice, st));
            NBT_CUDA(cudaMemcpyAsync(arena + o_c1, c1, (size_t)n_systems * 8, cudaMemcpyHostToDevice, st));
            if (n_data) {
                NBT_CUDA(cudaMemcpyAsync(arena + o_x, x, (size_t)n_data * 8, cudaMemcpyHostToDevice, st));
                NBT_CUDA(cudaMemcpyAsync(arena + o_y, y, (size_t)n_data * 8, cudaMemcpyHostToDevice, st));
                NBT_CUDA(cudaMemcpyAsync(arena + o_e, yerr, (size_t)n_data * 8, cudaMemcpyHostToDevice, st));
            }
            k_ttv = (double*)(arena + o_ttv); k_off = (int64_t*)(arena + o_off); k_n = (int32_t*)(arena + o_n);
            k_c0 = (double*)(arena + o_c0); k_c1 = (double*)(arena + o_c1); k_x = (double*)(arena + o_x);
            k_y = (double*)(arena + o_y); k_e = (double*)(arena + o_e); k_ll = (double*)(arena + o_ll);
        }
        k_chi2<<<(n_systems + 127) / 128, 128, 0, st>>>(k_ttv, k_off, k_n, n_systems, n_planets, k_c0, k_c1, k_x, k_y, k_e,
                                                       n_data, loss, k_ll);
        NBT_CUDA(cudaGetLastError());
        if (!device_ptrs) {
            NBT_CUDA(cudaMemcpyAsync(loglike, k_ll, (size_t)n_systems * 8, cudaMemcpyDeviceToHost, st));
            NBT_CUDA(cudaStreamSynchronize(st));
        }
    }
done:
    if (arena) cudaFreeAsync(arena, st);
    if (rc != 0) cudaGetLastError();
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return rc;
}


} // extern "C"

// Stages the host arrays of a small stand-alone entry point onto the device (or passes device
// pointers through), copies the outputs back and restores the caller's current device.
struct Stage {
    cudaStream_t st; bool dev_ptrs; int prev_dev = -1; int rc = 0;
    std::vector<void*> allocs;
    struct Out { void* h; void* d; size_t bytes; };
    std::vector<Out> outs;
    Stage(int device, int device_ptrs, void* stream) : st(stream ? (cudaStream_t)stream : cudaStreamPerThread), dev_ptrs(device_ptrs != 0) {
        cudaGetDevice(&prev_dev);
        check(cudaSetDevice(device), "cudaSetDevice");
        if (!rc) ensure_pool(device);
    }
    void check(cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == 0) rc = nbt_fail((int)e, "%s failed: %s", what, cudaGetErrorString(e));
    }
    template <class T> const T* in(const T* h, size_t count) {
        if (dev_ptrs || h == nullptr || rc) return h;
        void* d = nullptr;
        check(cudaMallocAsync(&d, std::max<size_t>(count * sizeof(T), 8), st), "cudaMallocAsync");
        if (rc) return nullptr;
        allocs.push_back(d);
        check(cudaMemcpyAsync(d, h, count * sizeof(T), cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(H2D)");
        return (const T*)d;
    }
    template <class T> T* out(T* h, size_t count) {
        if (dev_ptrs || h == nullptr || rc) return h;
        void* d = nullptr;
        check(cudaMallocAsync(&d, std::max<size_t>(count * sizeof(T), 8), st), "cudaMallocAsync");
        if (rc) return nullptr;
        allocs.push_back(d);
        outs.push_back({h, d, count * sizeof(T)});
        return (T*)d;
    }
    int finish() {
        if (!rc) check(cudaGetLastError(), "kernel launch");
        if (!dev_ptrs && !rc) {
            for (auto& o : outs) check(cudaMemcpyAsync(o.h, o.d, o.bytes, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(D2H)");
            check(cudaStreamSynchronize(st), "cudaStreamSynchronize");
        }
        for (void* d : allocs) cudaFreeAsync(d, st);
        if (rc) cudaGetLastError();
        if (prev_dev >= 0) cudaSetDevice(prev_dev);
        return rc;
    }
};

extern "C" {

int nbt_transit_times(const double* xp, const double* xs, const double* times, int64_t n_series, int64_t n,
                      double* tt, int32_t* n_tt, int64_t cap, int times_shared, int device, int device_ptrs,
                      void* cuda_stream) {
    if (!xp || !xs || !times || !tt || !n_tt) return nbt_fail(-10, "nbt_transit_times: NULL argument");
    if (n_series < 0 || n < 1 || cap < 0) return nbt_fail(-11, "nbt_transit_times: bad sizes");
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_transit_times: no CUDA device (there is no CPU fallback)");
    if (n_series == 0) return 0;
    Stage S(device, device_ptrs, cuda_stream);
    const double* dxp = S.in(xp, (size_t)n_series * n);
    const double* dxs = S.in(xs, (size_t)n_series * n);
    const double* dt = S.in(times, (size_t)(times_shared ? n : n_series * n));
    double* dtt = S.out(tt, (size_t)n_series * cap);
    int32_t* dn = S.out(n_tt, (size_t)n_series);
    if (!S.rc) k_transit_times<<<(unsigned)((n_series + 127) / 128), 128, 0, S.st>>>(dxp, dxs, dt, n_series, n, dtt, dn, cap, times_shared);
    return S.finish();
}

int nbt_find_zero(const double* t1, const double* dx1, const double* t2, const double* dx2, int64_t n,
                  double* t0, int device, int device_ptrs, void* cuda_stream) {
    if (!t1 || !dx1 || !t2 || !dx2 || !t0) return nbt_fail(-10, "nbt_find_zero: NULL argument");
    if (n < 0) return nbt_fail(-11, "nbt_find_zero: bad size");
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_find_zero: no CUDA device (there is no CPU fallback)");
    if (n == 0) return 0;
    Stage S(device, device_ptrs, cuda_stream);
    const double *a = S.in(t1, n), *b = S.in(dx1, n), *c = S.in(t2, n), *d = S.in(dx2, n);
    double* o = S.out(t0, n);
    if (!S.rc) k_find_zero<<<(unsigned)((n + 127) / 128), 128, 0, S.st>>>(a, b, c, d, n, o);
    return S.finish();
}

int nbt_ttv_fit(const double* epochs, const double* tt, int64_t n_series, int64_t n, double* ttv,
                double* slope, double* intercept, int device, int device_ptrs, void* cuda_stream) {
    if (!tt || !ttv || !slope || !intercept) return nbt_fail(-10, "nbt_ttv_fit: NULL argument");
    if (n_series < 0 || n < 1 || n > INT32_MAX) return nbt_fail(-11, "nbt_ttv_fit: bad sizes");
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_ttv_fit: no CUDA device (there is no CPU fallback)");
    if (n_series == 0) return 0;
    Stage S(device, device_ptrs, cuda_stream);
    const double* de = S.in(epochs, (size_t)n_series * n);
    const double* dt = S.in(tt, (size_t)n_series * n);
    double* dv = S.out(ttv, (size_t)n_series * n);
    double* dm = S.out(slope, (size_t)n_series);
    double* db = S.out(intercept, (size_t)n_series);
    if (!S.rc) k_ttv_fit<<<(unsigned)((n_series * 32 + 127) / 128), 128, 0, S.st>>>(de, dt, n_series, (int)n, dv, dm, db);
    return S.finish();
}

int nbt_maxavg(const double* x, int64_t n_series, int64_t n, double* out, int device, int device_ptrs,
               void* cuda_stream) {
    if (!x || !out) return nbt_fail(-10, "nbt_maxavg: NULL argument");
    if (n_series < 0 || n < 1 || n > INT32_MAX) return nbt_fail(-11, "nbt_maxavg: bad sizes");
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_maxavg: no CUDA device (there is no CPU fallback)");
    if (n_series == 0) return 0;
    Stage S(device, device_ptrs, cuda_stream);
    const double* dx = S.in(x, (size_t)n_series * n);
    double* dout = S.out(out, (size_t)n_series);
    if (!S.rc) k_maxavg<<<(unsigned)((n_series * 32 + 127) / 128), 128, 0, S.st>>>(dx, n_series, (int)n, dout);
    return S.finish();
}

int nbt_timing(double* ms4) {
    if (!ms4) return nbt_fail(-10, "nbt_timing: NULL argument");
    if (!g_timing.have) return nbt_fail(-13, "nbt_timing: no nbt_run with NBT_F_TIMING on this thread yet");
    for (int i = 0; i < 4; i++) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(g_timing.ev[i + 1]);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, g_timing.ev[i], g_timing.ev[i + 1]);
        if (e != cudaSuccess) { cudaGetLastError(); return nbt_fail((int)e, "nbt_timing: %s", cudaGetErrorString(e)); }
        ms4[i] = (double)ms;
    }
    return 0;
}


int nbt_grid_post(const double* tt, const double* ttv, const int64_t* tt_offsets, const int32_t* n_tt,
                  int32_t n_systems, int32_t n_planets, int32_t tt_cap_max, double minfreq, double maxfreq,
                  int32_t samples_per_peak, double peak_height, int32_t n_order, double* peak_periods,
                  double* coeffs, int32_t* n_peaks, double* softmax, double* avg_err, int device,
                  int device_ptrs, void* cuda_stream) {
    if (!tt || !ttv || !tt_offsets || !n_tt || !peak_periods || !coeffs || !n_peaks || !softmax || !avg_err)
        return nbt_fail(-10, "nbt_grid_post: NULL argument");
    if (n_systems < 0 || n_planets < 1 || n_order < 0 || n_order > NBT_GRID_MAXORDER || samples_per_peak < 1 || tt_cap_max < 1)
        return nbt_fail(-11, "nbt_grid_post: bad sizes (n_order <= %d)", NBT_GRID_MAXORDER);
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_grid_post: no CUDA device (there is no CPU fallback)");
    if (n_systems == 0) return 0;
    const int m = 2 + 2 * n_order;
    const long long nf_cap = nbt_ls_nfreq((double)(tt_cap_max - 1 > 1 ? tt_cap_max - 1 : 1), minfreq, maxfreq, samples_per_peak);
    long long need = std::max<long long>(nf_cap, (long long)tt_cap_max * m);
    if (need * 8 > 200 * 1024) return nbt_fail(-12, "nbt_grid_post: %lld doubles of shared memory needed (max 25600)", need);
    Stage S(device, device_ptrs, cuda_stream);
    const int64_t nsp = (int64_t)n_systems * n_planets;
    int64_t ntt = 0;
    if (!device_ptrs) ntt = tt_offsets[nsp];
    nbt_grid_args G;
    G.tt = S.in(tt, (size_t)ntt); G.ttv = S.in(ttv, (size_t)ntt);
    G.tt_offsets = S.in(tt_offsets, (size_t)nsp + 1); G.n_tt = S.in(n_tt, (size_t)nsp);
    G.peak_periods = S.out(peak_periods, (size_t)n_systems * n_order); G.coeffs = S.out(coeffs, (size_t)n_systems * m);
    G.n_peaks = S.out(n_peaks, (size_t)n_systems); G.softmax = S.out(softmax, (size_t)n_systems);
    G.avg_err = S.out(avg_err, (size_t)n_systems);
    G.B = n_systems; G.Np = n_planets; G.spp = samples_per_peak; G.n_order = n_order; G.smem_doubles = (int32_t)need;
    G.fmin = minfreq; G.fmax = maxfreq; G.height = peak_height;
    if (!S.rc) {
        S.check(cudaFuncSetAttribute(k_grid_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(need * 8)), "cudaFuncSetAttribute");
        if (!S.rc) k_grid_post<<<n_systems, 128, (size_t)need * 8, S.st>>>(G);
    }
    return S.finish();
}


int nbt_ttvfit_loglike(const double* tt, const int64_t* tt_offsets, const int32_t* n_tt, int32_t n_systems,
                       int32_t n_planets, const int32_t* has_data, const double* tc, const double* tc_err,
                       const int32_t* orbit, int32_t n_obs, double* loglike, int device, int device_ptrs,
                       void* cuda_stream) {
    if (!tt || !tt_offsets || !n_tt || !has_data || !tc || !tc_err || !orbit || !loglike)
        return nbt_fail(-10, "nbt_ttvfit_loglike: NULL argument");
    if (n_systems < 0 || n_planets < 1 || n_obs < 1) return nbt_fail(-11, "nbt_ttvfit_loglike: bad sizes");
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_ttvfit_loglike: no CUDA device (there is no CPU fallback)");
    if (n_systems == 0) return 0;
    Stage S(device, device_ptrs, cuda_stream);
    const int64_t nsp = (int64_t)n_systems * n_planets;
    const int64_t ntt = device_ptrs ? 0 : tt_offsets[nsp];
    const double* d_tt = S.in(tt, (size_t)ntt);
    const int64_t* d_off = S.in(tt_offsets, (size_t)nsp + 1);
    const int32_t* d_n = S.in(n_tt, (size_t)nsp);
    const int32_t* d_has = S.in(has_data, (size_t)n_planets);
    const double* d_tc = S.in(tc, (size_t)n_planets * n_obs);
    const double* d_err = S.in(tc_err, (size_t)n_planets * n_obs);
    const int32_t* d_orb = S.in(orbit, (size_t)n_obs);
    double* d_ll = S.out(loglike, (size_t)n_systems);
    if (!S.rc) k_ttvfit<<<(n_systems + 127) / 128, 128, 0, S.st>>>(d_tt, d_off, d_n, n_systems, n_planets, d_has, d_tc, d_err, d_orb, n_obs, d_ll);
    return S.finish();
}

int nbt_fp64_latency(int device, double* dfma_cycles, double* rsqrt_dadd_cycles) {
    int rc = 0;
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_fp64_latency: no CUDA device");
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    double* d = nullptr; long long* c = nullptr;
    const int n = 1 << 14;
    NBT_CUDA(cudaSetDevice(device));
    {
        NBT_CUDA(cudaMalloc((void**)&d, 32 * 8));
        NBT_CUDA(cudaMalloc((void**)&c, 16));
        k_fp64_latency<<<1, 32>>>(d, c, n, 0.999999, 1e-7);
        k_fp64_latency<<<1, 32>>>(d, c, n, 0.999999, 1e-7);
        NBT_CUDA(cudaGetLastError());
        long long h[2] = {0, 0};
        NBT_CUDA(cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost));
        if (dfma_cycles) *dfma_cycles = (double)h[0] / n;
        if (rsqrt_dadd_cycles) *rsqrt_dadd_cycles = (double)h[1] / n;
    }
done:
    if (d) cudaFree(d);
    if (c) cudaFree(c);
    if (rc != 0) cudaGetLastError();
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return rc;
}

int nbt_fp64_peak(int device, int iters, double* tflops_out, double* ms_out) {
    int rc = 0;
    if (nbt_device_count() <= 0) return nbt_fail(-100, "nbt_fp64_peak: no CUDA device");
    if (iters == 0) iters = 1 << 16;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double* d = nullptr;
    NBT_CUDA(cudaSetDevice(device));
    {
        cudaDeviceProp prop;
        NBT_CUDA(cudaGetDeviceProperties(&prop, device));
        const int blocks = prop.multiProcessorCount * 8, threads = 256;
        NBT_CUDA(cudaMalloc((void**)&d, 8));
        NBT_CUDA(cudaEventCreate(&e0));
        NBT_CUDA(cudaEventCreate(&e1));
        const bool three = iters < 0;   // negative iters: the three-register-operand variant
        if (three) iters = -iters;
        if (three) k_dfma_peak3<<<blocks, threads>>>(d, iters / 8, 0.999999, 1e-7);
        else k_dfma_peak<<<blocks, threads>>>(d, iters / 8, 0.999999, 1e-7); // warm-up
        NBT_CUDA(cudaGetLastError());
        NBT_CUDA(cudaEventRecord(e0));
        if (three) k_dfma_peak3<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
        else k_dfma_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
        NBT_CUDA(cudaEventRecord(e1));
        NBT_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        NBT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = (double)blocks * threads * 8.0 * (double)iters * 2.0;
        if (tflops_out) *tflops_out = flops / ((double)ms * 1e-3) * 1e-12;
        if (ms_out) *ms_out = (double)ms;
    }
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (d) cudaFree(d);
    if (rc != 0) cudaGetLastError();
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return rc;
}

} // extern "C"
