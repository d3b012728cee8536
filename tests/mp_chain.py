"""60-digit reference values (mpmath) for the per-CpG Ising chain summaries — an independent check that does not share
code or formulation with either the oracle or the CUDA kernels.  TEST INFRASTRUCTURE."""
import mpmath as mp
import numpy as np

mp.mp.dps = 60
X = (-1, 1)


def chain(phi, rho_n, d_n):
    a, b, c = (mp.mpf(float(x)) for x in phi)
    rho = [mp.mpf(float(x)) for x in rho_n]
    dn = [mp.mpf(float(x)) for x in d_n]
    N = len(rho)
    al = [a + b * rho[n] for n in range(N)]
    be = [c / dn[n] for n in range(N - 1)]
    A = [[mp.mpf(1), mp.mpf(1)]]                       # everything left of n (excl. site n's own potential)
    for n in range(1, N):
        F = [A[n - 1][k] * mp.e ** (al[n - 1] * X[k]) for k in range(2)]
        A.append([sum(F[k] * mp.e ** (be[n - 1] * X[k] * X[kk]) for k in range(2)) for kk in range(2)])
    B = [None] * N
    B[N - 1] = [mp.mpf(1), mp.mpf(1)]
    for n in range(N - 2, -1, -1):
        C = [B[n + 1][k] * mp.e ** (al[n + 1] * X[k]) for k in range(2)]
        B[n] = [sum(C[kk] * mp.e ** (be[n] * X[k] * X[kk]) for kk in range(2)) for k in range(2)]
    Z = sum(A[N - 1][k] * mp.e ** (al[N - 1] * X[k]) for k in range(2))
    ex, exx = [], []
    for n in range(N):
        pm = [A[n][k] * mp.e ** (al[n] * X[k]) * B[n][k] for k in range(2)]
        ex.append((pm[1] - pm[0]) / (pm[0] + pm[1]))
    for n in range(N - 1):
        num = den = mp.mpf(0)
        for k in range(2):
            for kk in range(2):
                w = A[n][k] * mp.e ** (al[n] * X[k]) * mp.e ** (be[n] * X[k] * X[kk]) * mp.e ** (al[n + 1] * X[kk]) * B[n + 1][kk]
                num += X[k] * X[kk] * w
                den += w
        exx.append(num / den)
    return dict(al=al, be=be, A=A, B=B, logZ=mp.log(Z), ex=ex, exx=exx, N=N)


def _log_g(ch, p, q):
    lg1 = [mp.mpf(0)] * 2 if p == 1 else [mp.log(ch["A"][p - 1][k]) for k in range(2)]
    lg2 = [mp.mpf(0)] * 2 if q == ch["N"] else [mp.log(ch["B"][q - 1][k]) for k in range(2)]
    return lg1, lg2


def cross_entropy(ch, mix, p, q):
    """H(θ, θ̃) over sub-region [p,q] (statistical_summaries.jl:427-459 in exact arithmetic); ch is mix for the entropy."""
    lg1, lg2 = _log_g(mix, p, q)
    pp, pq = (1 + ch["ex"][p - 1]) / 2, (1 + ch["ex"][q - 1]) / 2
    e1 = pp * lg1[1] + (1 - pp) * lg1[0]
    e2 = pq * lg2[1] + (1 - pq) * lg2[0]
    return (mix["logZ"] - e1 - e2 - sum(mix["al"][n - 1] * ch["ex"][n - 1] for n in range(p, q + 1))
            - sum(mix["be"][n - 1] * ch["exx"][n - 1] for n in range(p, q)))


def stat_sums(phi, rho_n, d_n, sub_p, sub_q):
    ch = chain(phi, rho_n, d_n)
    K = len(sub_p)
    mml, nme = np.full(K, np.nan), np.full(K, np.nan)
    for k in range(K):
        p, q = int(sub_p[k]), int(sub_q[k])
        if p == 0:
            continue
        Nk = q - p + 1
        mml[k] = float((Nk + sum(ch["ex"][p - 1:q])) / (2 * Nk))
        nme[k] = float(cross_entropy(ch, ch, p, q) / (Nk * mp.log(2)))
    return dict(logZ=float(ch["logZ"]), ex=np.array([float(v) for v in ch["ex"]]), exx=np.array([float(v) for v in ch["exx"]]), mml=mml,
                nme=nme, chain=ch)


def cmd(phi1, phi2, rho_n, d_n, sub_p, sub_q):
    c1, c2 = chain(phi1, rho_n, d_n), chain(phi2, rho_n, d_n)
    cm = chain(0.5 * (np.asarray(phi1) + np.asarray(phi2)), rho_n, d_n)
    out = np.full(len(sub_p), np.nan)
    for k in range(len(sub_p)):
        p, q = int(sub_p[k]), int(sub_q[k])
        if p == 0:
            continue
        H = cross_entropy(c1, c1, p, q) + cross_entropy(c2, c2, p, q)
        CE = cross_entropy(c1, cm, p, q) + cross_entropy(c2, cm, p, q)
        out[k] = float(1 - H / CE)
    return out
