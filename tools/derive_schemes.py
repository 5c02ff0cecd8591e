"""Derive / verify the coefficients of the symmetric BAB compositions used by the integrator.

A scheme is  K(b0) D(a1) K(b1) D(a2) ... D(aS) K(bS)  for H = H_Kepler + eps*H_int
(K = interaction kick, D = Kepler drift).  With kick "times" tau_i = a1+...+ai and c = tau-1/2:
  O(eps dt^k) error terms vanish  <=>  sum_i b_i c_i^j = int_{-1/2}^{1/2} c^j dc   for j < k
  the O(eps^2 dt^2) term {{A,B},B} vanishes  <=>  sum_{i<j} b_i b_j (tau_j - tau_i) = 1/6
(McLachlan 1995; Laskar & Robutel 2001).  Solved here with mpmath to 30 digits.
"""
import mpmath as mp

mp.mp.dps = 40


def taus(a):
    t, out = mp.mpf(0), [mp.mpf(0)]
    for ai in a:
        t += ai
        out.append(t)
    return out


def conditions(a, b):
    tau = taus(a)
    c = [t - mp.mpf(1) / 2 for t in tau]
    out = {"sum_a": sum(a), "sum_b": sum(b)}
    for j in (2, 4, 6, 8):
        out["m%d" % j] = sum(bi * ci ** j for bi, ci in zip(b, c)) - mp.mpf(1) / ((j + 1) * 2 ** j)
    e2 = mp.mpf(0)
    for i in range(len(b)):
        for j in range(i + 1, len(b)):
            e2 += b[i] * b[j] * (tau[j] - tau[i])
    out["eps2dt2"] = e2 - mp.mpf(1) / 6
    return out


def show(name, a, b):
    print("==", name)
    print(" a =", [mp.nstr(x, 25) for x in a])
    print(" b =", [mp.nstr(x, 25) for x in b])
    for k, v in conditions(a, b).items():
        print("   %-8s % .3e" % (k, float(v - (1 if k.startswith("sum") else 0))))


# leapfrog
show("WH2", [mp.mpf(1)], [mp.mpf(1) / 2] * 2)
# Yoshida 4 as BAB
w1 = 1 / (2 - mp.cbrt(2)); w0 = 1 - 2 * w1
show("Y4", [w1, w0, w1], [w1 / 2, (w0 + w1) / 2, (w0 + w1) / 2, w1 / 2])
# Lobatto SBAB4 (Laskar-Robutel): nodes 0, 1/2-sqrt(21)/14, 1/2, ..., weights 1/20, 49/180, 16/45
s21 = mp.sqrt(21)
show("SBAB4", [mp.mpf(1) / 2 - s21 / 14, s21 / 14, s21 / 14, mp.mpf(1) / 2 - s21 / 14],
     [mp.mpf(1) / 20, mp.mpf(49) / 180, mp.mpf(16) / 45, mp.mpf(49) / 180, mp.mpf(1) / 20])


def solve_bab4():
    """S=4: unknowns b0,b1,b2,tau1 ; equations sum_b, m2, m4, eps2dt2  -> order (6,4)."""
    def F(b0, b1, b2, t1):
        a = [t1, mp.mpf(1) / 2 - t1, mp.mpf(1) / 2 - t1, t1]
        b = [b0, b1, b2, b1, b0]
        c = conditions(a, b)
        return [c["sum_b"] - 1, c["m2"], c["m4"], c["eps2dt2"]]
    sols = []
    for t1 in [0.05, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, -0.1, 0.6]:
        for b0 in [0.02, 0.05, 0.1, 0.2, -0.05]:
            try:
                s = mp.findroot(F, (b0, 0.3, 0.3, t1), tol=1e-30, maxsteps=200)
            except Exception:
                continue
            key = tuple(round(float(v), 8) for v in s)
            if key not in [k for k, _ in sols]:
                sols.append((key, s))
    return [s for _, s in sols]


for s in solve_bab4():
    b0, b1, b2, t1 = s
    a = [t1, mp.mpf(1) / 2 - t1, mp.mpf(1) / 2 - t1, t1]
    b = [b0, b1, b2, b1, b0]
    show("BAB(6,4) S=4 candidate  max|coef|=%.3f" % max(abs(float(v)) for v in a + b), a, b)


def solve_bab3():
    """S=3: unknowns b0,b1,tau1; equations sum_b, m2, eps2dt2 -> order (4,4)."""
    def F(b0, b1, t1):
        a = [t1, 1 - 2 * t1, t1]
        b = [b0, b1, b1, b0]
        c = conditions(a, b)
        return [c["sum_b"] - 1, c["m2"], c["eps2dt2"]]
    sols = []
    for t1 in [0.1, 0.2, 0.3, 0.4, 0.6, 0.8, 1.2, 1.35]:
        for b0 in [0.05, 0.1, 0.2, 0.5, 0.7]:
            try:
                s = mp.findroot(F, (b0, 0.5 - b0, t1), tol=1e-30, maxsteps=200)
            except Exception:
                continue
            key = tuple(round(float(v), 8) for v in s)
            if key not in [k for k, _ in sols]:
                sols.append((key, s))
    return [s for _, s in sols]


for s in solve_bab3():
    b0, b1, t1 = s
    show("BAB(4,4) S=3 candidate", [t1, 1 - 2 * t1, t1], [b0, b1, b1, b0])
