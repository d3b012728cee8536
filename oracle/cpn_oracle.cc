// cpn_oracle.cc — CPU ORACLE.  TEST INFRASTRUCTURE ONLY — NOT PRODUCT CODE.
//
// A line-faithful C++ restatement (double precision, same operation order, same
// O(M·L²) structure) of the per-region Ising estimation hot path of
// jordiabante/CpelNano.jl.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / `--impl reference` leg may load this library; the
// product (cpelnano.jl_b200/) never links, imports or calls it.
//
// PARITY STATUS
//   * pinned against the reference's own fixtures (tests/test_oracle_*.py):
//       test/runtests.jl:9-76 known answers, the docstring known answers of
//       src/transfer_matrix.jl, and the golden outputs under examples/
//       (theta/ex/exx/mml/nme files, group-comparison tests).
//   * "parity unpinned" for the M-step: NLsolve 4.4.1 / NLSolversBase 7.7.1 /
//       FiniteDiff 2.8.0 / Calculus 0.5.1 (Manifest.toml) are third-party
//       packages whose source is NOT under /root/reference and Julia is not
//       installed, so their published algorithms (trust-region dogleg with
//       autoscaling, central finite differences) are restated here from their
//       documentation.  No reference test pins EM output.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/src/).  Build: see oracle/Makefile (-ffp-contract=off so the
// compiler never fuses a*b+c; the reference is plain IEEE double arithmetic).

#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

const double INF = std::numeric_limits<double>::infinity();
const double QNAN = std::numeric_limits<double>::quiet_NaN();
const double LOG2 = std::log(2.0);                 // CpelNano.jl:27
const double CBRT_EPS = std::cbrt(2.220446049250313e-16);  // cbrt(eps(Float64))

// ---- 2x2 log-domain algebra --------------------------------------------------
// State convention: index 1 <-> x=-1, index 2 <-> x=+1; M[x_l][x_{l+1}].
struct V2 { double v1, v2; };
struct M2 { double a11, a12, a21, a22; };

// Julia maximum/minimum of a 2-vector propagate NaN.
inline double jmax(double a, double b) { if (a != a || b != b) return QNAN; return a > b ? a : b; }
inline double jmin(double a, double b) { if (a != a || b != b) return QNAN; return a < b ? a : b; }

// transfer_matrix.jl:508-517 — log_sum_exp on a 2-vector: only max and min are used.
inline double log_sum_exp2(double a, double b) {
    double x_max = jmax(a, b), x_min = jmin(a, b);
    return std::log(1.0 + std::exp(x_min - x_max)) + x_max;
}
// transfer_matrix.jl:529-543
inline M2 log_mat_mult(const M2& A, const M2& B) {
    M2 r;
    r.a11 = log_sum_exp2(A.a11 + B.a11, A.a12 + B.a21);
    r.a12 = log_sum_exp2(A.a11 + B.a12, A.a12 + B.a22);
    r.a21 = log_sum_exp2(A.a21 + B.a11, A.a22 + B.a21);
    r.a22 = log_sum_exp2(A.a21 + B.a12, A.a22 + B.a22);
    return r;
}
// transfer_matrix.jl:555-567
inline V2 log_vec_mat_mult(const V2& u, const M2& W) {
    V2 r;
    r.v1 = log_sum_exp2(u.v1 + W.a11, u.v2 + W.a21);
    r.v2 = log_sum_exp2(u.v1 + W.a12, u.v2 + W.a22);
    return r;
}
// transfer_matrix.jl:579-584
inline double log_vec_vec_mult(const V2& a, const V2& b) { return log_sum_exp2(a.v1 + b.v1, a.v2 + b.v2); }
// transfer_matrix.jl:596-613 — left-to-right product of a list of log matrices.
inline M2 mult_log_mats(const M2* Ms, size_t n) {
    if (n == 1) return Ms[0];
    M2 acc = log_mat_mult(Ms[0], Ms[1]);
    for (size_t i = 2; i < n; ++i) acc = log_mat_mult(acc, Ms[i]);
    return acc;
}
inline M2 mult_log_mats(const std::vector<M2>& Ms) { return mult_log_mats(Ms.data(), Ms.size()); }
inline M2 madd(const M2& A, const M2& B) { return M2{A.a11 + B.a11, A.a12 + B.a12, A.a21 + B.a21, A.a22 + B.a22}; }

// CpelNano.jl:45-47
const M2 logD1 = {std::log(1.0), -INF, -INF, std::log(3.0)};
const M2 logD2 = {std::log(3.0), std::log(1.0), std::log(1.0), std::log(3.0)};
const M2 logD3 = {std::log(2.0), -INF, -INF, std::log(3.0)};

// ---- link function -------------------------------------------------------------
// link_functions.jl:16-25: α_l = N_l (a + b ρ_l), β_l = c / d_l
void get_ab_from_phi(const double* phi, int64_t L, const double* Nl, const double* rho, const double* dl,
                     std::vector<double>& al, std::vector<double>& be) {
    al.resize(L); be.resize(L > 0 ? L - 1 : 0);
    for (int64_t l = 0; l < L; ++l) al[l] = Nl[l] * (phi[0] + phi[1] * rho[l]);
    for (int64_t l = 0; l + 1 < L; ++l) be[l] = phi[2] / dl[l];
}

// ---- marginal chain -------------------------------------------------------------
// transfer_matrix.jl:1147-1155
inline V2 get_log_u(double a) { double aux = 0.5 * a; return V2{-aux, aux}; }
// transfer_matrix.jl:1170-1182
inline M2 get_log_W(double a1, double a2, double b) {
    M2 W;
    W.a11 = -0.5 * (a1 + a2) + b;
    W.a21 = 0.5 * (a1 - a2) - b;
    W.a12 = 0.5 * (a2 - a1) - b;
    W.a22 = 0.5 * (a1 + a2) + b;
    return W;
}
// transfer_matrix.jl:1199-1208
inline double get_log_Z(const V2& u1, const V2& uL, const std::vector<M2>& Ws) {
    M2 aux = mult_log_mats(Ws);
    V2 v = log_vec_mat_mult(u1, aux);
    return log_vec_vec_mult(v, uL);
}
void build_log_Ws(const std::vector<double>& al, const std::vector<double>& be, std::vector<M2>& Ws) {
    Ws.resize(be.size());
    for (size_t l = 0; l < be.size(); ++l) Ws[l] = get_log_W(al[l], al[l + 1], be[l]);
}
// log Z(ϕ) through the link: the closure f of get_∇logZ, transfer_matrix.jl:1228-1241
double log_Z_of_phi(const double* phi, int64_t L, const double* Nl, const double* rho, const double* dl) {
    std::vector<double> al, be; std::vector<M2> Ws;
    get_ab_from_phi(phi, L, Nl, rho, dl, al, be);
    build_log_Ws(al, be, Ws);
    return get_log_Z(get_log_u(al[0]), get_log_u(al[L - 1]), Ws);
}
// transfer_matrix.jl:1225-1246 with Calculus.gradient(f,x) (Calculus 0.5.1, central rule):
//   ε_k = cbrt(eps)·max(1,|x_k|);  g_k = (f(x+ε_k e_k) − f(x−ε_k e_k)) / (ε_k + ε_k),
//   perturbing the argument vector in place and restoring it.
void get_grad_logZ(double* phi /*perturbed in place and restored*/, int64_t L, const double* Nl, const double* rho,
                   const double* dl, double* g, int64_t* n_logZ) {
    for (int i = 0; i < 3; ++i) {
        double epsilon = CBRT_EPS * std::max(1.0, std::fabs(phi[i]));
        double oldx = phi[i];
        phi[i] = oldx + epsilon;
        double fp = log_Z_of_phi(phi, L, Nl, rho, dl);
        phi[i] = oldx - epsilon;
        double fm = log_Z_of_phi(phi, L, Nl, rho, dl);
        phi[i] = oldx;
        g[i] = (fp - fm) / (epsilon + epsilon);
    }
    if (n_logZ) *n_logZ += 6;
}
// Insert-a-matrix expectation machinery shared by get_E_X_log (:1273), get_Ec_X_log (:743),
// get_rs_exp_log_g1! (statistical_summaries.jl:226-229): product of
// vcat(Ws[1:(l-1)],[D],Ws[l:end]) with the end cases of the reference.
double log_chain_with_insert(const V2& u1, const V2& uL, const std::vector<M2>& Ws, size_t l /*1-based*/, const M2& D,
                             std::vector<M2>& tmp) {
    size_t nW = Ws.size();
    tmp.clear();
    if (l == 1) { tmp.push_back(D); tmp.insert(tmp.end(), Ws.begin(), Ws.end()); }
    else if (l == nW + 1) { tmp.insert(tmp.end(), Ws.begin(), Ws.end()); tmp.push_back(D); }
    else { tmp.insert(tmp.end(), Ws.begin(), Ws.begin() + (l - 1)); tmp.push_back(D); tmp.insert(tmp.end(), Ws.begin() + (l - 1), Ws.end()); }
    M2 P = mult_log_mats(tmp);
    return log_vec_vec_mult(log_vec_mat_mult(u1, P), uL);
}
// transfer_matrix.jl:1273-1296 (marginal) and :743-766 (conditional): E[X_l] = exp(· − logZ) − 2
void get_E_X_log(const V2& u1, const V2& uL, const std::vector<M2>& Ws, double logZ, double* ex) {
    std::vector<M2> tmp;
    for (size_t l = 1; l <= Ws.size() + 1; ++l) ex[l - 1] = std::exp(log_chain_with_insert(u1, uL, Ws, l, logD1, tmp) - logZ) - 2.0;
}
// transfer_matrix.jl:1322-1349 (marginal) and :898-925 (conditional): W_l + logD2 in place of W_l
void get_E_XX_log(const V2& u1, const V2& uL, const std::vector<M2>& Ws, double logZ, double* exx) {
    std::vector<M2> tmp;
    for (size_t l = 0; l < Ws.size(); ++l) {
        tmp = Ws;
        tmp[l] = madd(Ws[l], logD2);
        M2 P = mult_log_mats(tmp);
        exx[l] = std::exp(log_vec_vec_mult(log_vec_mat_mult(u1, P), uL) - logZ) - 2.0;
    }
}

// ---- conditional chain ----------------------------------------------------------
// transfer_matrix.jl:632-643.  A call is (obs, log_pyx_u, log_pyx_m); obs == !isnan(log_pyx_u) in the flat layout.
inline double get_gc(bool x, double a, bool obs, double lu, double lm) {
    double g = x ? a : -a;
    if (obs) g += x ? lm : lu;
    return 0.5 * g;
}
// transfer_matrix.jl:658-663
inline V2 get_log_uc(double a, bool obs, double lu, double lm) { return V2{get_gc(false, a, obs, lu, lm), get_gc(true, a, obs, lu, lm)}; }
// transfer_matrix.jl:679-691
inline M2 get_log_Wc(double a1, double a2, double b, bool o1, double lu1, double lm1, bool o2, double lu2, double lm2) {
    M2 W;
    W.a11 = get_gc(false, a1, o1, lu1, lm1) + get_gc(false, a2, o2, lu2, lm2) + b;
    W.a21 = get_gc(true, a1, o1, lu1, lm1) + get_gc(false, a2, o2, lu2, lm2) - b;
    W.a12 = get_gc(false, a1, o1, lu1, lm1) + get_gc(true, a2, o2, lu2, lm2) - b;
    W.a22 = get_gc(true, a1, o1, lu1, lm1) + get_gc(true, a2, o2, lu2, lm2) + b;
    return W;
}
// transfer_matrix.jl:944-965 (get_cond_exs_log) incl. get_log_Zc :709-714
void get_cond_exs_log(const std::vector<double>& al, const std::vector<double>& be, const double* lu, const double* lm,
                      double* ex, double* exx, double* logZc_out) {
    size_t L = al.size();
    auto obs = [&](size_t l) { return !(lu[l] != lu[l]); };
    V2 u1 = get_log_uc(al[0], obs(0), lu[0], lm[0]);
    V2 uL = get_log_uc(al[L - 1], obs(L - 1), lu[L - 1], lm[L - 1]);
    std::vector<M2> Ws(be.size());
    for (size_t l = 0; l < be.size(); ++l)
        Ws[l] = get_log_Wc(al[l], al[l + 1], be[l], obs(l), lu[l], lm[l], obs(l + 1), lu[l + 1], lm[l + 1]);
    double logZc = log_vec_vec_mult(log_vec_mat_mult(u1, mult_log_mats(Ws)), uL);
    get_E_X_log(u1, uL, Ws, logZc, ex);
    get_E_XX_log(u1, uL, Ws, logZc, exx);
    if (logZc_out) *logZc_out = logZc;
}
// transfer_matrix.jl:843-870 — log p(y_m) with diag(p(y|−1),p(y|+1)) inserted in the prior chain.
double get_logpy_lgtrck(const V2& u1, const V2& uL, const std::vector<M2>& Ws, double logZ, const double* lu, const double* lm) {
    size_t nW = Ws.size();
    auto obs = [&](size_t l) { return !(lu[l] != lu[l]); };
    M2 mats = {0.0, 0.0, 0.0, 0.0};
    for (size_t l = 0; l < nW; ++l) {
        if (obs(l)) {
            M2 D = {lu[l], -INF, -INF, lm[l]};
            M2 DW = log_mat_mult(D, Ws[l]);
            mats = (l == 0) ? DW : log_mat_mult(mats, DW);
        } else {
            mats = (l == 0) ? Ws[l] : log_mat_mult(mats, Ws[l]);
        }
    }
    if (obs(nW)) {
        M2 D = {lu[nW], -INF, -INF, lm[nW]};
        mats = log_mat_mult(mats, D);
    }
    V2 v = log_vec_mat_mult(u1, mats);
    return log_vec_vec_mult(v, uL) - logZ;
}

// ---- region data for EM (group scale) ------------------------------------------
struct EmRegion {
    int64_t L, M;
    const double *Nl, *rho, *dl;   // [L],[L],[L-1]
    const double *lu, *lm;         // [M*L] read-major, NaN = unobserved
};
struct EmCounters { int64_t em_iters = 0, solver_iters = 0, logZ_evals = 0, estep_passes = 0; };

// em_algorithm.jl:99-113 — sums in read order, left-to-right dot products.
void e_step(const EmRegion& rg, const double* phi, double* ES) {
    std::vector<double> al, be;
    get_ab_from_phi(phi, rg.L, rg.Nl, rg.rho, rg.dl, al, be);
    std::vector<double> ex(rg.L), exx(rg.L - 1);
    ES[0] = ES[1] = ES[2] = 0.0;
    for (int64_t m = 0; m < rg.M; ++m) {
        get_cond_exs_log(al, be, rg.lu + m * rg.L, rg.lm + m * rg.L, ex.data(), exx.data(), nullptr);
        double d1 = 0.0, d2 = 0.0, d3 = 0.0;
        for (int64_t l = 0; l < rg.L; ++l) d1 += rg.Nl[l] * ex[l];
        for (int64_t l = 0; l < rg.L; ++l) d2 += (rg.Nl[l] * rg.rho[l]) * ex[l];
        for (int64_t l = 0; l + 1 < rg.L; ++l) d3 += (1.0 / rg.dl[l]) * exx[l];
        ES[0] += d1; ES[1] += d2; ES[2] += d3;
    }
}

// ---- restated NLsolve.nlsolve(f!, x0; iterations=20, ftol=1e-3) -------------------
// Defaults: method=:trust_region, xtol=0.0, factor=1.0, autoscale=true, autodiff=:central.
struct MStepProblem {
    const EmRegion* rg; const double* ES; EmCounters* cnt;
    // f!(U,ϕ): em_algorithm.jl:152-157
    void residual(double* phi, double* U) const {
        double g[3];
        get_grad_logZ(phi, rg->L, rg->Nl, rg->rho, rg->dl, g, cnt ? &cnt->logZ_evals : nullptr);
        U[0] = g[0] - ES[0] / (double)rg->M;
        U[1] = g[1] - ES[1] / (double)rg->M;
        U[2] = g[2] - ES[2] / (double)rg->M;
    }
    // NLSolversBase fj_finitediff!: f!(F,x) then FiniteDiff.finite_difference_jacobian!(J,f!,x,cache) (central):
    //   ε_k = max(cbrt(eps)·|x_k|, cbrt(eps)); J[:,k] = (F(x+ε_k e_k) − F(x−ε_k e_k)) / (2ε_k)
    void value_jacobian(double* x, double* F, double J[3][3]) const {
        residual(x, F);
        jacobian_only(x, J);
    }
    void jacobian_only(double* x, double J[3][3]) const {
        double x1[3] = {x[0], x[1], x[2]}, f1[3], f0[3];
        for (int k = 0; k < 3; ++k) {
            double x_save = x[k];
            double epsilon = std::max(CBRT_EPS * std::fabs(x_save), CBRT_EPS);
            x1[k] = x_save + epsilon; residual(x1, f1);
            x[k] = x_save - epsilon;  residual(x, f0);
            for (int i = 0; i < 3; ++i) J[i][k] = (f1[i] - f0[i]) / (2 * epsilon);
            x1[k] = x_save; x[k] = x_save;
        }
    }
};
inline double wdot(const double* wx, const double* x, const double* wy, const double* y) {
    double out = 0.0; for (int i = 0; i < 3; ++i) out += wx[i] * x[i] * wy[i] * y[i]; return out;
}
inline double wnorm(const double* w, const double* x) { return std::sqrt(wdot(w, x, w, x)); }
inline double sumabs2(const double* x) { return x[0] * x[0] + x[1] * x[1] + x[2] * x[2]; }
inline double norminf(const double* x) { return std::max(std::fabs(x[0]), std::max(std::fabs(x[1]), std::fabs(x[2]))); }
inline bool anynan(const double* x) { return x[0] != x[0] || x[1] != x[1] || x[2] != x[2]; }
inline bool allfinite(const double* x) { return std::isfinite(x[0]) && std::isfinite(x[1]) && std::isfinite(x[2]); }

// J \ r by LU with partial pivoting; returns false when a pivot is exactly zero (SingularException).
bool lu_solve3(const double Jin[3][3], const double* r, double* out) {
    double A[3][3]; double b[3] = {r[0], r[1], r[2]};
    std::memcpy(A, Jin, sizeof(A));
    for (int k = 0; k < 3; ++k) {
        int piv = k; double best = std::fabs(A[k][k]);
        for (int i = k + 1; i < 3; ++i) if (std::fabs(A[i][k]) > best) { best = std::fabs(A[i][k]); piv = i; }
        if (A[piv][k] == 0.0) return false;
        if (piv != k) { for (int j = 0; j < 3; ++j) std::swap(A[k][j], A[piv][j]); std::swap(b[k], b[piv]); }
        for (int i = k + 1; i < 3; ++i) {
            double m = A[i][k] / A[k][k];
            A[i][k] = m;
            for (int j = k + 1; j < 3; ++j) A[i][j] -= m * A[k][j];
            b[i] -= m * b[k];
        }
    }
    for (int i = 2; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < 3; ++j) s -= A[i][j] * out[j];
        out[i] = s / A[i][i];
    }
    return true;
}
// Moore–Penrose solve for the singular fallback of NLsolve's dogleg!: p = pinv(J) r with singular
// values ≤ eps() dropped.  Built from the Jacobi eigen-decomposition of JᵀJ.
void pinv_solve3(const double J[3][3], const double* r, double* out) {
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] = 0; for (int k = 0; k < 3; ++k) A[i][j] += J[k][i] * J[k][j]; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            if (A[p][q] == 0.0) continue;
            double th = (A[q][q] - A[p][p]) / (2 * A[p][q]);
            double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
            double c = 1 / std::sqrt(t * t + 1), s = t * c;
            for (int k = 0; k < 3; ++k) { double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
            for (int k = 0; k < 3; ++k) { double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
            for (int k = 0; k < 3; ++k) { double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
        }
    }
    double Jtr[3]; for (int i = 0; i < 3; ++i) { Jtr[i] = 0; for (int k = 0; k < 3; ++k) Jtr[i] += J[k][i] * r[k]; }
    out[0] = out[1] = out[2] = 0.0;
    for (int e = 0; e < 3; ++e) {
        double lam = A[e][e];                       // σ² of J
        if (!(lam > 0) || std::sqrt(lam) <= 2.220446049250313e-16) continue;
        double c = 0; for (int i = 0; i < 3; ++i) c += V[i][e] * Jtr[i];
        for (int i = 0; i < 3; ++i) out[i] += V[i][e] * c / lam;
    }
}
inline void matvec3(const double J[3][3], const double* v, double* out) {
    for (int i = 0; i < 3; ++i) out[i] = J[i][0] * v[0] + J[i][1] * v[1] + J[i][2] * v[2];
}
// NLsolve trust_region.jl dogleg!
void dogleg(double* p, const double* r, const double* d, const double J[3][3], double delta) {
    double p_i[3], p_c[3];
    if (!lu_solve3(J, r, p_i)) pinv_solve3(J, r, p_i);
    for (int i = 0; i < 3; ++i) p_i[i] = -p_i[i];
    if (wnorm(d, p_i) <= delta) { for (int i = 0; i < 3; ++i) p[i] = p_i[i]; return; }
    double g[3];
    for (int j = 0; j < 3; ++j) g[j] = J[0][j] * r[0] + J[1][j] * r[1] + J[2][j] * r[2];
    for (int j = 0; j < 3; ++j) g[j] = g[j] / (d[j] * d[j]);
    double Jg[3]; matvec3(J, g, Jg);
    double wg = wnorm(d, g);
    double coef = -(wg * wg) / sumabs2(Jg);
    for (int i = 0; i < 3; ++i) p_c[i] = coef * g[i];
    if (wnorm(d, p_c) >= delta) {
        double s = -delta / wnorm(d, g);
        for (int i = 0; i < 3; ++i) p[i] = g[i] * s;
    } else {
        double p_diff[3]; for (int i = 0; i < 3; ++i) p_diff[i] = p_i[i] - p_c[i];
        double b = 2 * wdot(d, p_c, d, p_diff);
        double wpd = wnorm(d, p_diff); double a = wpd * wpd;
        double wpc = wnorm(d, p_c);
        double tau = (-b + std::sqrt(b * b - 4 * a * (wpc * wpc - delta * delta))) / (2 * a);
        for (int i = 0; i < 3; ++i) p[i] = p_c[i] + tau * p_diff[i];
    }
}
// NLsolve trust_region_: returns converged(sol); x0 is copied, result in x.
// `threw` mirrors check_isfinite on the initial residual (IsFiniteException → caught by em_iter's try/catch).
bool nlsolve_trust_region(const MStepProblem& pb, const double* x0, double* x, double ftol, int iterations, bool* threw) {
    const double xtol = 0.0, factor = 1.0, eta = 1e-4;
    *threw = false;
    double r[3], J[3][3], fval[3], d[3], p[3], xold[3] = {QNAN, QNAN, QNAN};
    for (int i = 0; i < 3; ++i) x[i] = x0[i];
    pb.value_jacobian(x, fval, J);
    for (int i = 0; i < 3; ++i) r[i] = fval[i];
    if (!allfinite(r)) { *threw = true; return false; }
    int it = 0;
    bool x_converged = false;                     // xtol passed as NaN in the initial check
    bool f_converged = norminf(fval) <= ftol;
    bool stopped = anynan(x) || anynan(fval);
    bool converged = x_converged || f_converged;
    if (converged) return true;
    for (int j = 0; j < 3; ++j) {
        d[j] = std::sqrt(J[0][j] * J[0][j] + J[1][j] * J[1][j] + J[2][j] * J[2][j]);
        if (d[j] == 0.0) d[j] = 1.0;
    }
    double delta = factor * wnorm(d, x);
    if (delta == 0.0) delta = factor;
    while (!stopped && !converged && it < iterations) {
        ++it;
        if (pb.cnt) pb.cnt->solver_iters++;
        dogleg(p, r, d, J, delta);
        for (int i = 0; i < 3; ++i) xold[i] = x[i];
        for (int i = 0; i < 3; ++i) x[i] += p[i];
        pb.residual(x, fval);
        double r_predict[3]; matvec3(J, p, r_predict);
        for (int i = 0; i < 3; ++i) r_predict[i] += r[i];
        double rho = (sumabs2(r) - sumabs2(fval)) / (sumabs2(r) - sumabs2(r_predict));
        if (rho > eta) {
            for (int i = 0; i < 3; ++i) r[i] = fval[i];
            double Fdummy[3];
            pb.value_jacobian(x, Fdummy, J);      // jacobian!(df,x) re-evaluates F once (j_finitediff! → fj_finitediff!)
            for (int j = 0; j < 3; ++j) {
                double nj = std::sqrt(J[0][j] * J[0][j] + J[1][j] * J[1][j] + J[2][j] * J[2][j]);
                d[j] = std::max(0.1 * d[j], nj);
            }
            double dx[3] = {x[0] - xold[0], x[1] - xold[1], x[2] - xold[2]};
            x_converged = norminf(dx) <= xtol;
            f_converged = norminf(r) <= ftol;
            converged = x_converged || f_converged;
        } else {
            for (int i = 0; i < 3; ++i) x[i] -= p[i];
            x_converged = false; converged = false;
        }
        if (rho < 0.1) delta = delta / 2;
        else if (rho >= 0.9) delta = 2 * wnorm(d, p);
        else if (rho >= 0.5) delta = std::max(delta, 2 * wnorm(d, p));
        stopped = anynan(x) || anynan(fval);
    }
    return x_converged || f_converged;
}

// em_algorithm.jl:132-183
void em_iter(const EmRegion& rg, const double* phi_init, const double* box, double* phi_out, EmCounters* cnt) {
    double ES[3];
    e_step(rg, phi_init, ES);
    if (cnt) cnt->estep_passes++;
    MStepProblem pb{&rg, ES, cnt};
    double sol[3]; bool threw = false;
    bool conv = nlsolve_trust_region(pb, phi_init, sol, 1e-3, 20, &threw);
    if (threw || !conv) { for (int i = 0; i < 3; ++i) phi_out[i] = phi_init[i]; return; }
    for (int i = 0; i < 3; ++i) phi_out[i] = std::max(-box[i], std::min(sol[i], box[i]));
}
// em_algorithm.jl:263-288
double comp_ave_log_py(const EmRegion& rg, const double* phi) {
    std::vector<double> al, be; std::vector<M2> Ws;
    get_ab_from_phi(phi, rg.L, rg.Nl, rg.rho, rg.dl, al, be);
    V2 u1 = get_log_u(al[0]), uL = get_log_u(al[rg.L - 1]);
    build_log_Ws(al, be, Ws);
    double logZ = get_log_Z(u1, uL, Ws);
    double log_py = 0.0;
    for (int64_t m = 0; m < rg.M; ++m) log_py += get_logpy_lgtrck(u1, uL, Ws, logZ, rg.lu + m * rg.L, rg.lm + m * rg.L);
    return log_py / (double)rg.M;
}
// em_algorithm.jl:18-23
inline bool conv_check(const double* a, const double* b) {
    double s = 0.0; for (int i = 0; i < 3; ++i) { double d = a[i] - b[i]; s += d * d; }
    return std::sqrt(s) <= 3.0e-2;
}
// em_algorithm.jl:304-350
void em_inst(const EmRegion& rg, const double* phi0, int64_t max_em_iters, const double* box, double* phi_out, double* Q, EmCounters* cnt) {
    int64_t i = 1; bool conv = false, hit = false;
    double pt[3] = {phi0[0], phi0[1], phi0[2]}, ptp1[3];
    while (!conv && !hit) {
        em_iter(rg, pt, box, ptp1, cnt);
        if (cnt) cnt->em_iters++;
        conv = conv_check(pt, ptp1);
        hit = i >= max_em_iters;
        for (int k = 0; k < 3; ++k) pt[k] = ptp1[k];
        i += 1;
    }
    conv = conv && i > 2;
    *Q = conv ? comp_ave_log_py(rg, pt) : -INF;
    for (int k = 0; k < 3; ++k) phi_out[k] = pt[k];
}
// em_algorithm.jl:371-390 with pick_ϕ :35-53 (strict >, first maximum wins); starts are inputs (host RNG).
// Returns index of the picked start or -1.
int get_phihat(const EmRegion& rg, const double* phi0s, int64_t n_init, int64_t max_em_iters, const double* box,
               double* phi_hat, double* Qhat, double* phi_starts /*[n_init*3] or null*/, double* Q_starts /*[n_init] or null*/, EmCounters* cnt) {
    double best = -INF; int bi = -1; double bphi[3] = {QNAN, QNAN, QNAN};
    for (int64_t s = 0; s < n_init; ++s) {
        double ph[3], Q;
        em_inst(rg, phi0s + 3 * s, max_em_iters, box, ph, &Q, cnt);
        if (phi_starts) for (int k = 0; k < 3; ++k) phi_starts[3 * s + k] = ph[k];
        if (Q_starts) Q_starts[s] = Q;
        if (Q > best) { best = Q; bi = (int)s; for (int k = 0; k < 3; ++k) bphi[k] = ph[k]; }
    }
    *Qhat = best;
    for (int k = 0; k < 3; ++k) phi_hat[k] = (best > -INF) ? bphi[k] : QNAN;
    return (best > -INF) ? bi : -1;
}

// ---- analysis sub-regions (integers, bit-exact) -----------------------------------
// genome_analysis.jl:472-538
int64_t nls_reg_info(int64_t chrst, int64_t chrend, int64_t max_size_subreg, const int64_t* cpg_pos, int64_t N,
                     int64_t* sub_st, int64_t* sub_end, int64_t* sub_p, int64_t* sub_q) {
    int64_t reg_size = (chrend - chrst) + 1;
    int64_t k = (int64_t)std::ceil((double)reg_size / (double)max_size_subreg);
    if (k < 1) return 0;
    int64_t base_size = (int64_t)std::floor((double)reg_size / (double)k);
    std::vector<int64_t> sz(k, base_size);
    int64_t rem = ((reg_size % k) + k) % k;
    int64_t i = 0;
    while (rem > 0) { sz[i] += 1; rem -= 1; i = (i == k - 1) ? 0 : i + 1; }
    std::vector<int64_t> bnds(k + 1);
    bnds[0] = chrst;
    int64_t acc = 0;
    for (int64_t j = 0; j < k; ++j) { acc += sz[j]; bnds[j + 1] = acc + chrst - 1; }
    if (!sub_st) return k;
    for (int64_t j = 0; j < k; ++j) {
        sub_st[j] = bnds[j]; sub_end[j] = bnds[j + 1];
        int64_t fst = 0, lst = 0;
        for (int64_t n = 0; n < N; ++n) if (bnds[j] <= cpg_pos[n] && cpg_pos[n] < bnds[j + 1]) { fst = n + 1; break; }
        for (int64_t n = N - 1; n >= 0; --n) if (bnds[j] <= cpg_pos[n] && cpg_pos[n] < bnds[j + 1]) { lst = n + 1; break; }
        if (j < k - 1) { sub_p[j] = fst; sub_q[j] = lst; }
        else { sub_p[j] = fst; sub_q[j] = N; }                // last sub-region always ends at N (:524-526)
    }
    return k;
}

// ---- statistical summaries of a fitted per-CpG chain ------------------------------
struct ChainModel {            // after rscle_grp_mod! (nanopore.jl:269-280): N_l = 1, ρ_l = ρ_n, d_l = d_n, L = N
    int64_t N;
    std::vector<double> al, be;
    V2 u1, uN; std::vector<M2> Ws;
    double logZ;
    std::vector<double> ex, exx;
};
// statistical_summaries.jl:43-83 + transfer_matrix.jl:1361-1372
void build_chain(const double* phi, int64_t N, const double* rho_n, const double* d_n, ChainModel& cm, bool with_exps) {
    std::vector<double> ones(N, 1.0);
    cm.N = N;
    get_ab_from_phi(phi, N, ones.data(), rho_n, d_n, cm.al, cm.be);
    cm.u1 = get_log_u(cm.al[0]); cm.uN = get_log_u(cm.al[N - 1]);
    build_log_Ws(cm.al, cm.be, cm.Ws);
    cm.logZ = get_log_Z(cm.u1, cm.uN, cm.Ws);
    if (with_exps) {
        cm.ex.resize(N); cm.exx.resize(N - 1);
        get_E_X_log(cm.u1, cm.uN, cm.Ws, cm.logZ, cm.ex.data());
        get_E_XX_log(cm.u1, cm.uN, cm.Ws, cm.logZ, cm.exx.data());
    }
}
// statistical_summaries.jl:94-122
double comp_log_g_p(const ChainModel& cm, int64_t p, double xp) {
    if (p == 1) return 0.0;
    const std::vector<double>& a = cm.al; const std::vector<double>& b = cm.be;
    double aux = (a[p - 2] + xp * b[p - 2]) / 2.0;              // αs[p-1], βs[p-1] (1-based)
    V2 uNp = {-aux, +aux};
    if (p == 2) return log_vec_vec_mult(uNp, uNp);
    M2 Wp;                                                       // uses αs[p-2], βs[p-2] (1-based)
    Wp.a11 = -a[p - 3] / 2.0 + b[p - 3] - aux; Wp.a12 = -a[p - 3] / 2.0 - b[p - 3] + aux;
    Wp.a21 = +a[p - 3] / 2.0 - b[p - 3] - aux; Wp.a22 = +a[p - 3] / 2.0 + b[p - 3] + aux;
    if (p == 3) return log_vec_vec_mult(log_vec_mat_mult(cm.u1, Wp), uNp);
    std::vector<M2> tmp(cm.Ws.begin(), cm.Ws.begin() + (p - 3));
    tmp.push_back(Wp);
    return log_vec_vec_mult(log_vec_mat_mult(cm.u1, mult_log_mats(tmp)), uNp);
}
// statistical_summaries.jl:133-161
double comp_log_g_q(const ChainModel& cm, int64_t q, double xq) {
    int64_t L = cm.N;
    if (q == L) return 0.0;
    const std::vector<double>& a = cm.al; const std::vector<double>& b = cm.be;
    double aux = (a[q] + xq * b[q - 1]) / 2.0;                   // αs[q+1], βs[q] (1-based)
    V2 u1p = {-aux, +aux};
    if (q == L - 1) return log_vec_vec_mult(u1p, u1p);
    M2 Wp;                                                       // βs[q+1], αs[q+2] (1-based)
    Wp.a11 = -aux + b[q] - a[q + 1] / 2.0; Wp.a12 = -aux - b[q] + a[q + 1] / 2.0;
    Wp.a21 = +aux - b[q] - a[q + 1] / 2.0; Wp.a22 = +aux + b[q] + a[q + 1] / 2.0;
    if (q == L - 2) return log_vec_vec_mult(log_vec_mat_mult(u1p, Wp), cm.uN);
    std::vector<M2> tmp; tmp.push_back(Wp);
    tmp.insert(tmp.end(), cm.Ws.begin() + (q + 1), cm.Ws.end()); // log_Ws[(q+2):end] (1-based)
    return log_vec_vec_mult(log_vec_mat_mult(u1p, mult_log_mats(tmp)), cm.uN);
}
// statistical_summaries.jl:226-229 / :298-301 — E[log g(X_s)] under model `cm` with log g values (gm,gp).
double exp_log_g_at(const ChainModel& cm, int64_t site /*1-based*/, double gm, double gp, std::vector<M2>& tmp) {
    M2 logD = {std::log(gm), std::log(0.0), std::log(0.0), std::log(gp)};
    return std::exp(log_chain_with_insert(cm.u1, cm.uN, cm.Ws, (size_t)site, logD, tmp) - cm.logZ);
}
struct StatSums {
    std::vector<double> log_g;            // [4K] pm,pp,qm,qp
    std::vector<double> e_log_g1, e_log_g2, mml, nme;
};
// statistical_summaries.jl:172-200, :212-236, :284-308, :355-415, :470-500
void stat_sums(const ChainModel& cm, int64_t K, const int64_t* sub_p, const int64_t* sub_q, StatSums& ss) {
    int64_t N = cm.N;
    ss.log_g.assign(4 * K, QNAN); ss.e_log_g1.assign(K, QNAN); ss.e_log_g2.assign(K, QNAN);
    ss.mml.assign(K, QNAN); ss.nme.assign(K, QNAN);
    std::vector<M2> tmp;
    for (int64_t k = 0; k < K; ++k) {
        int64_t p = sub_p[k], q = sub_q[k];
        if (p != 0 && q != 0) {
            ss.log_g[4 * k + 0] = comp_log_g_p(cm, p, -1.0);
            ss.log_g[4 * k + 1] = comp_log_g_p(cm, p, +1.0);
            ss.log_g[4 * k + 2] = comp_log_g_q(cm, q, -1.0);
            ss.log_g[4 * k + 3] = comp_log_g_q(cm, q, +1.0);
        }
    }
    for (int64_t k = 0; k < K; ++k) {
        int64_t p = sub_p[k], q = sub_q[k];
        if (p != 0) ss.e_log_g1[k] = (p == 1) ? 0.0 : exp_log_g_at(cm, p, ss.log_g[4 * k + 0], ss.log_g[4 * k + 1], tmp);
        if (q != 0) ss.e_log_g2[k] = (q == N) ? 0.0 : exp_log_g_at(cm, q, ss.log_g[4 * k + 2], ss.log_g[4 * k + 3], tmp);
    }
    for (int64_t k = 0; k < K; ++k) {
        int64_t p = sub_p[k], q = sub_q[k];
        if (p == 0) continue;                 // (also covers the reference's out-of-bounds 0:N case → NaN)
        double Nk = (double)(q - p + 1);
        double s = 0.0; for (int64_t n = p; n <= q; ++n) s += cm.ex[n - 1];
        ss.mml[k] = 0.5 / Nk * (Nk + s);
        double dot1 = 0.0; for (int64_t n = p; n <= q; ++n) dot1 += cm.al[n - 1] * cm.ex[n - 1];
        double nme = cm.logZ - ss.e_log_g1[k] - ss.e_log_g2[k] - dot1;
        if (p != q) { double dot2 = 0.0; for (int64_t n = p; n <= q - 1; ++n) dot2 += cm.be[n - 1] * cm.exx[n - 1]; nme += -dot2; }
        else nme += 0.0;
        nme /= (double)(q - p + 1) * LOG2;
        ss.nme[k] = nme;
    }
}
// statistical_summaries.jl:427-459 with :248-272, :320-344
void comp_crs_ent(const ChainModel& cm, const ChainModel& mix, const std::vector<double>& mix_log_g, int64_t K,
                  const int64_t* sub_p, const int64_t* sub_q, std::vector<double>& out) {
    out.assign(K, QNAN);
    std::vector<M2> tmp;
    for (int64_t k = 0; k < K; ++k) {
        int64_t p = sub_p[k], q = sub_q[k];
        if (p == 0) continue;
        double g1 = (p == 1) ? 0.0 : exp_log_g_at(cm, p, mix_log_g[4 * k + 0], mix_log_g[4 * k + 1], tmp);
        double g2 = (q == cm.N) ? 0.0 : exp_log_g_at(cm, q, mix_log_g[4 * k + 2], mix_log_g[4 * k + 3], tmp);
        double dot1 = 0.0; for (int64_t n = p; n <= q; ++n) dot1 += mix.al[n - 1] * cm.ex[n - 1];
        double ce = mix.logZ - dot1 - g1 - g2;
        if (p != q) { double dot2 = 0.0; for (int64_t n = p; n <= q - 1; ++n) dot2 += mix.be[n - 1] * cm.exx[n - 1]; ce += -dot2; }
        else ce += 0.0;
        out[k] = ce;
    }
}
// statistical_summaries.jl:512-570
void comp_cmd(const double* phi1, const double* phi2, const double* nme1, const double* nme2, int64_t N, const double* rho_n,
              const double* d_n, int64_t K, const int64_t* sub_p, const int64_t* sub_q, double* cmd) {
    double phim[3]; for (int i = 0; i < 3; ++i) phim[i] = 0.5 * (phi1[i] + phi2[i]);
    ChainModel m1, m2, mix;
    build_chain(phi1, N, rho_n, d_n, m1, true);
    build_chain(phi2, N, rho_n, d_n, m2, true);
    build_chain(phim, N, rho_n, d_n, mix, true);
    std::vector<double> mix_log_g(4 * K, QNAN);
    for (int64_t k = 0; k < K; ++k) {
        int64_t p = sub_p[k], q = sub_q[k];
        if (p != 0 && q != 0) {
            mix_log_g[4 * k + 0] = comp_log_g_p(mix, p, -1.0); mix_log_g[4 * k + 1] = comp_log_g_p(mix, p, +1.0);
            mix_log_g[4 * k + 2] = comp_log_g_q(mix, q, -1.0); mix_log_g[4 * k + 3] = comp_log_g_q(mix, q, +1.0);
        }
    }
    std::vector<double> c1, c2;
    comp_crs_ent(m1, mix, mix_log_g, K, sub_p, sub_q, c1);
    comp_crs_ent(m2, mix, mix_log_g, K, sub_p, sub_q, c2);
    for (int64_t k = 0; k < K; ++k) {
        cmd[k] = QNAN;
        int64_t p = sub_p[k], q = sub_q[k];
        if (p == 0) continue;
        double ent1 = nme1[k] * (double)(q - p + 1) * LOG2;
        double ent2 = nme2[k] * (double)(q - p + 1) * LOG2;
        if (c1[k] == 0.0 && c2[k] == 0.0) continue;
        cmd[k] = 1.0 - (ent1 + ent2) / (c1[k] + c2[k]);
    }
}

// Combinatorics.combinations(1:n,k): lexicographic order.
bool next_comb(std::vector<int32_t>& c, int32_t n) {
    int32_t k = (int32_t)c.size();
    int32_t i = k - 1;
    while (i >= 0 && c[i] == n - k + i + 1) --i;
    if (i < 0) return false;
    ++c[i];
    for (int32_t j = i + 1; j < k; ++j) c[j] = c[j - 1] + 1;
    return true;
}

}  // namespace

// ==================================================================================
// C API (flat arrays; tests bind it with ctypes through oracle/oracle_py.py)
// ==================================================================================
extern "C" {

double orc_log_sum_exp3(double a, double b, double c) {   // docstring KAT transfer_matrix.jl:504-505 (3-vector: max & min only)
    double mx = jmax(jmax(a, b), c), mn = jmin(jmin(a, b), c);
    return std::log(1.0 + std::exp(mn - mx)) + mx;
}
double orc_get_gc(int x, double a, double lu, double lm) { return get_gc(x != 0, a, !(lu != lu), lu, lm); }
void orc_get_log_Wc(double a1, double a2, double b, double lu1, double lm1, double lu2, double lm2, double* W4 /*a11,a12,a21,a22*/) {
    M2 W = get_log_Wc(a1, a2, b, !(lu1 != lu1), lu1, lm1, !(lu2 != lu2), lu2, lm2);
    W4[0] = W.a11; W4[1] = W.a12; W4[2] = W.a21; W4[3] = W.a22;
}
// log Z of the marginal chain from α, β (transfer_matrix.jl:1199)
double orc_log_Z(int64_t L, const double* al, const double* be) {
    std::vector<double> a(al, al + L), b(be, be + (L - 1)); std::vector<M2> Ws;
    build_log_Ws(a, b, Ws);
    return get_log_Z(get_log_u(a[0]), get_log_u(a[L - 1]), Ws);
}
// E[X], E[XX] of the marginal chain by matrix insertion (transfer_matrix.jl:1273, :1322)
void orc_marg_exs_log(int64_t L, const double* al, const double* be, double* logZ, double* ex, double* exx) {
    std::vector<double> a(al, al + L), b(be, be + (L - 1)); std::vector<M2> Ws;
    build_log_Ws(a, b, Ws);
    V2 u1 = get_log_u(a[0]), uL = get_log_u(a[L - 1]);
    *logZ = get_log_Z(u1, uL, Ws);
    get_E_X_log(u1, uL, Ws, *logZ, ex);
    get_E_XX_log(u1, uL, Ws, *logZ, exx);
}
// P(x_l=+1) via logD3 (transfer_matrix.jl:1399-1422)
void orc_marg_px_log(int64_t L, const double* al, const double* be, double* px) {
    std::vector<double> a(al, al + L), b(be, be + (L - 1)); std::vector<M2> Ws, tmp;
    build_log_Ws(a, b, Ws);
    V2 u1 = get_log_u(a[0]), uL = get_log_u(a[L - 1]);
    double logZ = get_log_Z(u1, uL, Ws);
    for (int64_t l = 1; l <= L; ++l) px[l - 1] = std::exp(log_chain_with_insert(u1, uL, Ws, (size_t)l, logD3, tmp) - logZ) - 2.0;
}
// conditional expectations of one read (transfer_matrix.jl:944)
void orc_cond_exs_log(int64_t L, const double* al, const double* be, const double* lu, const double* lm, double* logZc, double* ex, double* exx) {
    std::vector<double> a(al, al + L), b(be, be + (L - 1));
    get_cond_exs_log(a, b, lu, lm, ex, exx, logZc);
}
// log p(y_m) (transfer_matrix.jl:843)
double orc_logpy(int64_t L, const double* al, const double* be, const double* lu, const double* lm) {
    std::vector<double> a(al, al + L), b(be, be + (L - 1)); std::vector<M2> Ws;
    build_log_Ws(a, b, Ws);
    V2 u1 = get_log_u(a[0]), uL = get_log_u(a[L - 1]);
    return get_logpy_lgtrck(u1, uL, Ws, get_log_Z(u1, uL, Ws), lu, lm);
}
void orc_get_ab_from_phi(const double* phi, int64_t L, const double* Nl, const double* rho, const double* dl, double* al, double* be) {
    std::vector<double> a, b; get_ab_from_phi(phi, L, Nl, rho, dl, a, b);
    std::copy(a.begin(), a.end(), al); std::copy(b.begin(), b.end(), be);
}
void orc_grad_logZ(const double* phi, int64_t L, const double* Nl, const double* rho, const double* dl, double* g) {
    double x[3] = {phi[0], phi[1], phi[2]};
    get_grad_logZ(x, L, Nl, rho, dl, g, nullptr);
}
int orc_conv_check(const double* a, const double* b, int64_t n) {   // em_algorithm.jl:18 (any length, docstring KAT :12-15)
    double s = 0.0; for (int64_t i = 0; i < n; ++i) { double d = a[i] - b[i]; s += d * d; }
    return std::sqrt(s) <= 3.0e-2;
}
void orc_e_step(int64_t L, int64_t M, const double* Nl, const double* rho, const double* dl, const double* lu, const double* lm,
                const double* phi, double* ES) {
    EmRegion rg{L, M, Nl, rho, dl, lu, lm};
    e_step(rg, phi, ES);
}
// One M-step (NLsolve restatement) from phi_init given ES; returns 1 if converged (em_algorithm.jl:152-178 without the clamp).
int orc_m_step(int64_t L, int64_t M, const double* Nl, const double* rho, const double* dl, const double* ES, const double* phi_init,
               double* sol, int64_t* counters /*[4] or null*/) {
    EmRegion rg{L, M, Nl, rho, dl, nullptr, nullptr};
    EmCounters c; MStepProblem pb{&rg, ES, &c};
    bool threw = false;
    bool conv = nlsolve_trust_region(pb, phi_init, sol, 1e-3, 20, &threw);
    if (counters) { counters[0] = c.em_iters; counters[1] = c.solver_iters; counters[2] = c.logZ_evals; counters[3] = c.estep_passes; }
    return (conv && !threw) ? 1 : 0;
}
void orc_em_iter(int64_t L, int64_t M, const double* Nl, const double* rho, const double* dl, const double* lu, const double* lm,
                 const double* phi_init, const double* box, double* phi_out) {
    EmRegion rg{L, M, Nl, rho, dl, lu, lm};
    em_iter(rg, phi_init, box, phi_out, nullptr);
}
double orc_ave_log_py(int64_t L, int64_t M, const double* Nl, const double* rho, const double* dl, const double* lu, const double* lm, const double* phi) {
    EmRegion rg{L, M, Nl, rho, dl, lu, lm};
    return comp_ave_log_py(rg, phi);
}
// Multi-start EM for a CSR batch of regions — same layout as cpn_fit_batch (include/cpelnano_b200.h).
// n_threads > 1 distributes regions dynamically (the analogue of Distributed.pmap, nanopore.jl:406).
int orc_fit_batch(int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l, const int64_t* read_off,
                  const int64_t* obs_off, const double* log_pyx_u, const double* log_pyx_m, const double* phi0, int32_t n_init,
                  int32_t max_em_iters, double amax, double bmax, double cmax, double* phi_hat, double* Q, int32_t* status,
                  int64_t* counters, double* phi_starts, double* Q_starts, int32_t n_threads) {
    double box[3] = {amax, bmax, cmax};
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            int64_t r = next.fetch_add(1);
            if (r >= R) break;
            int64_t L = grp_off[r + 1] - grp_off[r], M = read_off[r + 1] - read_off[r];
            EmRegion rg{L, M, Nl + grp_off[r], rho_l + grp_off[r], d_l + (grp_off[r] - r), log_pyx_u + obs_off[r], log_pyx_m + obs_off[r]};
            EmCounters c;
            int pick = get_phihat(rg, phi0 + (size_t)r * n_init * 3, n_init, max_em_iters, box, phi_hat + 3 * r, Q + r,
                                  phi_starts ? phi_starts + (size_t)r * n_init * 3 : nullptr, Q_starts ? Q_starts + (size_t)r * n_init : nullptr, &c);
            if (status) status[r] = pick;
            if (counters) { counters[4 * r] = c.em_iters; counters[4 * r + 1] = c.solver_iters; counters[4 * r + 2] = c.logZ_evals; counters[4 * r + 3] = c.estep_passes; }
        }
    };
    int nt = n_threads > 0 ? n_threads : 1;
    if (nt == 1) worker();
    else { std::vector<std::thread> th; for (int t = 0; t < nt; ++t) th.emplace_back(worker); for (auto& t : th) t.join(); }
    return 0;
}
int64_t orc_nls_reg_info(int64_t chrst, int64_t chrend, int64_t max_size_subreg, const int64_t* cpg_pos, int64_t N, int64_t* sub_st,
                         int64_t* sub_end, int64_t* sub_p, int64_t* sub_q) {
    return nls_reg_info(chrst, chrend, max_size_subreg, cpg_pos, N, sub_st, sub_end, sub_p, sub_q);
}
// get_stat_sums! after rscle_grp_mod! for one fitted region (statistical_summaries.jl:470)
void orc_stat_sums(const double* phi, int64_t N, const double* rho_n, const double* d_n, int64_t K, const int64_t* sub_p, const int64_t* sub_q,
                   double* logZ, double* ex, double* exx, double* log_g, double* e_log_g1, double* e_log_g2, double* mml, double* nme) {
    ChainModel cm; build_chain(phi, N, rho_n, d_n, cm, true);
    StatSums ss; stat_sums(cm, K, sub_p, sub_q, ss);
    *logZ = cm.logZ;
    std::copy(cm.ex.begin(), cm.ex.end(), ex); std::copy(cm.exx.begin(), cm.exx.end(), exx);
    std::copy(ss.log_g.begin(), ss.log_g.end(), log_g);
    std::copy(ss.e_log_g1.begin(), ss.e_log_g1.end(), e_log_g1); std::copy(ss.e_log_g2.begin(), ss.e_log_g2.end(), e_log_g2);
    std::copy(ss.mml.begin(), ss.mml.end(), mml); std::copy(ss.nme.begin(), ss.nme.end(), nme);
}
void orc_comp_cmd(const double* phi1, const double* phi2, const double* nme1, const double* nme2, int64_t N, const double* rho_n,
                  const double* d_n, int64_t K, const int64_t* sub_p, const int64_t* sub_q, double* cmd) {
    comp_cmd(phi1, phi2, nme1, nme2, N, rho_n, d_n, K, sub_p, sub_q, cmd);
}

// Unmatched group test sweep (grp_perm_tests.jl:200-291) for one region given statistic tables.
//   mml,nme: [S][K] (S = n1+n2, group 1 first);  cmd_tbl: [S*S][K], entry (i*S+j) valid for i<j (0-based);
//   labelings: [P][n1] ascending 1-based sample ids.  Outputs T[3K] (mml,nme,cmd), cnt[3K], p[3K].
void orc_unmat_test(int32_t n1, int32_t n2, int64_t K, const double* mml, const double* nme, const double* cmd_tbl,
                    const int32_t* labelings, int64_t P, int32_t exact, double* T, int64_t* cnt, double* pval) {
    int32_t S = n1 + n2;
    auto obs_stat = [&](const double* tab, int64_t k) {      // comp_unmat_stat_mml :118-127: sum(vectors)/n, left to right
        double s1 = tab[0 * K + k]; for (int32_t s = 1; s < n1; ++s) s1 = s1 + tab[(int64_t)s * K + k];
        double s2 = tab[(int64_t)n1 * K + k]; for (int32_t s = n1 + 1; s < S; ++s) s2 = s2 + tab[(int64_t)s * K + k];
        return s1 / (double)n1 - s2 / (double)n2;
    };
    for (int64_t k = 0; k < K; ++k) {
        T[k] = obs_stat(mml, k); T[K + k] = obs_stat(nme, k);
        // comp_unmat_stat_cmd :13-26: sum over (s1 outer, s2 inner) then / (n1*n2)
        double acc = 0.0; bool first = true;
        for (int32_t s1 = 0; s1 < n1; ++s1) for (int32_t s2 = 0; s2 < n2; ++s2) {
            double v = cmd_tbl[((int64_t)s1 * S + (n1 + s2)) * K + k];
            if (first) { acc = v; first = false; } else acc = acc + v;
        }
        T[2 * K + k] = acc / (double)(n1 * n2);
        cnt[k] = cnt[K + k] = cnt[2 * K + k] = 0;
        pval[k] = pval[K + k] = pval[2 * K + k] = QNAN;
    }
    if (P <= 0) return;
    std::vector<int32_t> g2(n2);
    for (int64_t pi = 0; pi < P; ++pi) {
        const int32_t* g1 = labelings + pi * n1;
        int32_t a = 0, c = 0;
        for (int32_t s = 1; s <= S; ++s) { if (a < n1 && g1[a] == s) ++a; else g2[c++] = s; }
        for (int64_t k = 0; k < K; ++k) {
            if (T[k] != T[k]) continue;                           // isnan(tmml_obs[k]) && continue
            auto perm_stat = [&](const double* tab) {
                double s1 = tab[(int64_t)(g1[0] - 1) * K + k]; for (int32_t s = 1; s < n1; ++s) s1 = s1 + tab[(int64_t)(g1[s] - 1) * K + k];
                double s2 = tab[(int64_t)(g2[0] - 1) * K + k]; for (int32_t s = 1; s < n2; ++s) s2 = s2 + tab[(int64_t)(g2[s] - 1) * K + k];
                return s1 / (double)n1 - s2 / (double)n2;
            };
            double tm = perm_stat(mml), tn = perm_stat(nme);
            double tc = 0.0;                                      // comp_unmat_perm_stat_cmd :86-107
            for (int32_t i = 0; i < n1; ++i) for (int32_t j = 0; j < n2; ++j) {
                int32_t ii = g1[i] - 1, jj = g2[j] - 1;
                int32_t lo = std::min(ii, jj), hi = std::max(ii, jj);
                tc = tc + cmd_tbl[((int64_t)lo * S + hi) * K + k];
            }
            tc = tc / (double)(n1 * n2);
            if (std::fabs(tm) >= std::fabs(T[k])) cnt[k]++;
            if (std::fabs(tn) >= std::fabs(T[K + k])) cnt[K + k]++;
            if (tc >= T[2 * K + k]) cnt[2 * K + k]++;
        }
    }
    for (int64_t k = 0; k < K; ++k) {
        if (T[k] != T[k]) continue;
        for (int t = 0; t < 3; ++t)
            pval[t * K + k] = exact ? (double)cnt[t * K + k] / (double)P : (1.0 + (double)cnt[t * K + k]) / (1.0 + (double)P);
    }
}
// All C(n,k) labelings in the order of Combinatorics.combinations (grp_perm_tests.jl:233); returns count.
int64_t orc_all_combinations(int32_t n, int32_t k, int32_t* out /*[count*k] or null*/) {
    std::vector<int32_t> c(k); for (int32_t i = 0; i < k; ++i) c[i] = i + 1;
    int64_t cnt = 0;
    do { if (out) std::copy(c.begin(), c.end(), out + cnt * k); ++cnt; } while (next_comb(c, n));
    return cnt;
}
// Matched group test sweep (grp_perm_tests.jl:432-499): mml1,mml2,nme1,nme2 [n][K]; js[P] sign words.
void orc_mat_test(int32_t n, int64_t K, const double* mml1, const double* mml2, const double* nme1, const double* nme2,
                  const int64_t* js, int64_t P, int32_t exact, double* T /*[2K]*/, int64_t* cnt /*[2K]*/, double* pval /*[2K]*/) {
    std::vector<double> dm((size_t)n * K), dn((size_t)n * K);
    for (int32_t s = 0; s < n; ++s) for (int64_t k = 0; k < K; ++k) {
        dm[(size_t)s * K + k] = mml1[(size_t)s * K + k] - mml2[(size_t)s * K + k];
        dn[(size_t)s * K + k] = nme1[(size_t)s * K + k] - nme2[(size_t)s * K + k];
    }
    for (int64_t k = 0; k < K; ++k) {
        double a = dm[k], b = dn[k];
        for (int32_t s = 1; s < n; ++s) { a = a + dm[(size_t)s * K + k]; b = b + dn[(size_t)s * K + k]; }
        T[k] = a / (double)n; T[K + k] = b / (double)n;
        cnt[k] = cnt[K + k] = 0; pval[k] = pval[K + k] = QNAN;
    }
    if (P <= 0) return;
    for (int64_t pi = 0; pi < P; ++pi) {
        for (int64_t k = 0; k < K; ++k) {
            if (T[k] != T[k]) continue;
            int64_t jk = js[pi]; double a = 0.0, b = 0.0;      // comp_mat_j_stat :388-405
            for (int32_t s = 0; s < n; ++s) {
                bool bit = (jk & 1) != 0;
                a += bit ? dm[(size_t)s * K + k] : -dm[(size_t)s * K + k];
                b += bit ? dn[(size_t)s * K + k] : -dn[(size_t)s * K + k];
                jk >>= 1;
            }
            a = a / (double)n; b = b / (double)n;
            if (std::fabs(a) >= std::fabs(T[k])) cnt[k]++;
            if (std::fabs(b) >= std::fabs(T[K + k])) cnt[K + k]++;
        }
    }
    for (int64_t k = 0; k < K; ++k) {
        if (T[k] != T[k]) continue;
        for (int t = 0; t < 2; ++t)
            pval[t * K + k] = exact ? (double)cnt[t * K + k] / (double)P : (1.0 + (double)cnt[t * K + k]) / (1.0 + (double)P);
    }
}
// Two-sample p-value counting (two_samp_perm_tests.jl:230-258): T_obs[3K], T_null[P][3][K] → cnt, n_valid, p.
void orc_two_samp_pvals(int64_t K, int64_t P, const double* T_obs, const double* T_null, int32_t exact, int64_t* cnt, int64_t* n_valid, double* pval) {
    for (int64_t k = 0; k < K; ++k) {
        for (int t = 0; t < 3; ++t) { cnt[t * K + k] = 0; n_valid[t * K + k] = 0; pval[t * K + k] = QNAN; }
        if (T_obs[k] != T_obs[k]) continue;
        for (int64_t pi = 0; pi < P; ++pi) {
            for (int t = 0; t < 3; ++t) {
                double v = T_null[(pi * 3 + t) * K + k];
                if (v != v) continue;
                n_valid[t * K + k]++;
                if (t < 2) { if (std::fabs(v) >= std::fabs(T_obs[t * K + k])) cnt[t * K + k]++; }
                else if (v >= T_obs[t * K + k]) cnt[t * K + k]++;
            }
        }
        if (!(n_valid[k] > 20)) continue;                     // length(tmml_perms_k) > 20 || continue
        for (int t = 0; t < 3; ++t)
            pval[t * K + k] = exact ? (double)cnt[t * K + k] / (double)n_valid[t * K + k] : (1.0 + (double)cnt[t * K + k]) / (1.0 + (double)n_valid[t * K + k]);
    }
}

// ---- counter-based stream of EM starts and read-label permutations ----------------------
// Independent restatement of the stream the library defines in cpelnano.jl_b200/csrc/cpn_philox.h, from the published
// Philox-4x32-10 algorithm (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; known-answer
// vectors of the Random123 distribution are checked in tests/test_oracle_kats.py).  The reference itself draws from Julia's
// global MersenneTwister (gen_ϕ0s em_algorithm.jl:65-78; sort!(sample(1:n, n1; replace=false)) two_samp_perm_tests.jl:211-214):
// that stream cannot be reproduced without Julia and is not claimed.
static void philox_round(uint32_t ctr[4], const uint32_t key[2]) {
    const uint64_t prod_a = 0xD2511F53ull * ctr[0];
    const uint64_t prod_b = 0xCD9E8D57ull * ctr[2];
    const uint32_t hi_a = (uint32_t)(prod_a >> 32), lo_a = (uint32_t)(prod_a & 0xffffffffull);
    const uint32_t hi_b = (uint32_t)(prod_b >> 32), lo_b = (uint32_t)(prod_b & 0xffffffffull);
    const uint32_t out[4] = {hi_b ^ ctr[1] ^ key[0], lo_b, hi_a ^ ctr[3] ^ key[1], lo_a};
    for (int i = 0; i < 4; ++i) ctr[i] = out[i];
}
void orc_philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out4) {
    uint32_t ctr[4] = {c0, c1, c2, c3};
    uint32_t key[2] = {(uint32_t)(seed & 0xffffffffull), (uint32_t)(seed >> 32)};
    philox_round(ctr, key);
    for (int r = 1; r < 10; ++r) {
        key[0] += 0x9E3779B9u; key[1] += 0xBB67AE85u;            // Weyl sequence key schedule
        philox_round(ctr, key);
    }
    for (int i = 0; i < 4; ++i) out4[i] = ctr[i];
}
static uint32_t scale_u32(uint32_t x, uint32_t m) { return (uint32_t)(((uint64_t)x * m) / 4294967296ull); }   // ⌊x·m / 2^32⌋ ∈ 0..m−1
// start `start` of unit (region, perm, half): the grid of gen_ϕ0s (em_algorithm.jl:72): a ∈ −5:0.01:5, b ∈ −50:0.01:50, c ∈ 0:0.01:10
void orc_gen_phi0(uint64_t seed, uint32_t region, uint32_t perm, uint32_t half, uint32_t start, double* phi0) {
    uint32_t w[4];
    orc_philox4x32_10(seed, region, perm, half * 65536u + start, 0u, w);
    phi0[0] = ((double)scale_u32(w[0], 1001u) - 500.0) / 100.0;
    phi0[1] = ((double)scale_u32(w[1], 10001u) - 5000.0) / 100.0;
    phi0[2] = (double)scale_u32(w[2], 1001u) / 100.0;
}
// ascending 1-based indices of sample 1 for permutation `perm` of `region` (n pooled reads, n1 in sample 1): selection sampling
void orc_gen_perm(uint64_t seed, uint32_t region, uint32_t perm, int32_t n, int32_t n1, int32_t* ids) {
    int32_t need = n1, k = 0;
    uint32_t w[4] = {0, 0, 0, 0};
    for (int32_t q = 0; q < n; ++q) {
        if (q % 4 == 0) orc_philox4x32_10(seed, region, perm, (uint32_t)(q / 4), 1u, w);
        if ((int32_t)scale_u32(w[q % 4], (uint32_t)(n - q)) < need) { ids[k++] = q + 1; --need; }
    }
}

// ---- legacy non-log primitives: only to pin test/runtests.jl:9-25,46-59 --------------
static void mm2(const double* A, const double* B, double* C) {
    C[0] = A[0] * B[0] + A[1] * B[2]; C[1] = A[0] * B[1] + A[1] * B[3];
    C[2] = A[2] * B[0] + A[3] * B[2]; C[3] = A[2] * B[1] + A[3] * B[3];
}
// get_u :282, get_W :305, get_Z :338, get_E_X :372, get_E_XX :419 (marginal, lu==null) and the conditional twins :19-174
void orc_nonlog_chain(int64_t L, const double* al, const double* be, const double* lu, const double* lm, double* Z, double* ex, double* exx) {
    std::vector<double> W(4 * (L - 1));
    double u1[2], uL[2];
    if (!lu) {
        double a1 = std::exp(0.5 * al[0]), aL = std::exp(0.5 * al[L - 1]);
        u1[0] = 1.0 / a1; u1[1] = a1; uL[0] = 1.0 / aL; uL[1] = aL;
        for (int64_t l = 0; l + 1 < L; ++l) {
            double e1 = std::exp(0.5 * al[l]), e2 = std::exp(0.5 * al[l + 1]), eb = std::exp(be[l]);
            W[4 * l + 0] = eb / (e1 * e2); W[4 * l + 2] = e1 / (e2 * eb); W[4 * l + 1] = e2 / (e1 * eb); W[4 * l + 3] = e1 * e2 * eb;
        }
    } else {
        auto g = [&](bool x, int64_t l) { return std::exp(get_gc(x, al[l], !(lu[l] != lu[l]), lu[l], lm[l])); };
        u1[0] = g(false, 0); u1[1] = g(true, 0); uL[0] = g(false, L - 1); uL[1] = g(true, L - 1);
        for (int64_t l = 0; l + 1 < L; ++l) {
            double f1 = g(false, l), t1 = g(true, l), f2 = g(false, l + 1), t2 = g(true, l + 1), eb = std::exp(be[l]);
            W[4 * l + 0] = f1 * f2 * eb; W[4 * l + 2] = t1 * f2 / eb; W[4 * l + 1] = f1 * t2 / eb; W[4 * l + 3] = t1 * t2 * eb;
        }
    }
    auto prod = [&](int64_t a, int64_t b, double* P) {   // prod(Ws[a..b)) left to right; identity if empty
        P[0] = 1; P[1] = 0; P[2] = 0; P[3] = 1; bool first = true;
        for (int64_t l = a; l < b; ++l) { if (first) { std::memcpy(P, &W[4 * l], 32); first = false; } else { double T[4]; mm2(P, &W[4 * l], T); std::memcpy(P, T, 32); } }
    };
    auto bil = [&](const double* P) { return (u1[0] * P[0] + u1[1] * P[2]) * uL[0] + (u1[0] * P[1] + u1[1] * P[3]) * uL[1]; };
    double P[4]; prod(0, L - 1, P); *Z = bil(P);
    const double DEX[4] = {-1, 0, 0, 1};
    for (int64_t l = 0; l < L; ++l) {
        double A[4], B[4], T1[4], T2[4];
        prod(0, l, A); prod(l, L - 1, B);
        mm2(A, DEX, T1); mm2(T1, B, T2);
        ex[l] = bil(T2) / *Z;
    }
    for (int64_t l = 0; l + 1 < L; ++l) {
        double A[4], B[4], Wd[4] = {W[4 * l], -W[4 * l + 1], -W[4 * l + 2], W[4 * l + 3]}, T1[4], T2[4];
        prod(0, l, A); prod(l + 1, L - 1, B);
        mm2(A, Wd, T1); mm2(T1, B, T2);
        exx[l] = bil(T2) / *Z;
    }
}

// ---- WGBS E-step: em_iter_wgbs (em_algorithm.jl:412-424) = get_ess ∘ get_cond_exs (transfer_matrix.jl:191-217) ---------------
// get_cond_exs is the NON-log conditional chain: get_uc :19, get_Wc :43-66, the matrices scaled by 1/max entry (:202:
// UniformScaling(1.0/maximum(maximum.(Ws))) * Ws), then get_Zc :73, get_Ec_X :102-126, get_Ec_XX :150-174 with D_EX = [-1 0;0 1] and
// Ws[l] .* D_EXX.  WGBS calls are hard (log p ∈ {0, −50}, CpelNano.jl:32-33).  Same left-to-right products as orc_nonlog_chain.
void orc_e_step_wgbs(int64_t L, int64_t M, const double* Nl, const double* rho, const double* dl, const double* lu, const double* lm,
                     const double* phi, double* ES) {
    std::vector<double> al, be;
    get_ab_from_phi(phi, L, Nl, rho, dl, al, be);
    std::vector<double> W(4 * (L - 1)), ex(L), exx(L - 1);
    ES[0] = ES[1] = ES[2] = 0.0;
    for (int64_t m = 0; m < M; ++m) {
        const double *u = lu + m * L, *v = lm + m * L;
        auto g = [&](bool x, int64_t l) { return std::exp(get_gc(x, al[l], !(u[l] != u[l]), u[l], v[l])); };
        const double u1[2] = {g(false, 0), g(true, 0)}, uL[2] = {g(false, L - 1), g(true, L - 1)};
        double wmax = -std::numeric_limits<double>::infinity();
        for (int64_t l = 0; l + 1 < L; ++l) {
            double f1 = g(false, l), t1 = g(true, l), f2 = g(false, l + 1), t2 = g(true, l + 1), eb = std::exp(be[l]);
            W[4 * l + 0] = f1 * f2 * eb; W[4 * l + 2] = t1 * f2 / eb; W[4 * l + 1] = f1 * t2 / eb; W[4 * l + 3] = t1 * t2 * eb;
            for (int k = 0; k < 4; ++k) wmax = std::max(wmax, W[4 * l + k]);
        }
        const double sc = 1.0 / wmax;
        for (auto& w : W) w = sc * w;
        auto prod = [&](int64_t a, int64_t b, double* P) {
            P[0] = 1; P[1] = 0; P[2] = 0; P[3] = 1; bool first = true;
            for (int64_t l = a; l < b; ++l) { if (first) { std::memcpy(P, &W[4 * l], 32); first = false; } else { double T[4]; mm2(P, &W[4 * l], T); std::memcpy(P, T, 32); } }
        };
        auto bil = [&](const double* P) { return (u1[0] * P[0] + u1[1] * P[2]) * uL[0] + (u1[0] * P[1] + u1[1] * P[3]) * uL[1]; };
        double P[4]; prod(0, L - 1, P);
        const double Zc = bil(P);
        const double DEX[4] = {-1, 0, 0, 1};
        for (int64_t l = 0; l < L; ++l) {
            double A[4], B[4], T1[4], T2[4];
            prod(0, l, A); prod(l, L - 1, B);
            mm2(A, DEX, T1); mm2(T1, B, T2);
            ex[l] = bil(T2) / Zc;
        }
        for (int64_t l = 0; l + 1 < L; ++l) {
            double A[4], B[4], Wd[4] = {W[4 * l], -W[4 * l + 1], -W[4 * l + 2], W[4 * l + 3]}, T1[4], T2[4];
            prod(0, l, A); prod(l + 1, L - 1, B);
            mm2(A, Wd, T1); mm2(T1, B, T2);
            exx[l] = bil(T2) / Zc;
        }
        double d1 = 0.0, d2 = 0.0, d3 = 0.0;                 // get_ess, em_algorithm.jl:99-113
        for (int64_t l = 0; l < L; ++l) d1 += Nl[l] * ex[l];
        for (int64_t l = 0; l < L; ++l) d2 += (Nl[l] * rho[l]) * ex[l];
        for (int64_t l = 0; l + 1 < L; ++l) d3 += (1.0 / dl[l]) * exx[l];
        ES[0] += d1; ES[1] += d2; ES[2] += d3;
    }
}

}  // extern "C"
