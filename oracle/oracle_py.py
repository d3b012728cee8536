"""ctypes binding of the CPU oracle (oracle/cpn_oracle.cc).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg —
never by the product package.  Function names mirror the reference's Julia names (ϕ → phi, ! dropped).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcpn_oracle.so")
_lib = None

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "cpn_oracle.cc")
    if force or not os.path.exists(_LIB_PATH) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_log_sum_exp3.restype = C.c_double
        _lib.orc_log_sum_exp3.argtypes = [C.c_double] * 3
        _lib.orc_get_gc.restype = C.c_double
        _lib.orc_get_gc.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        _lib.orc_log_Z.restype = C.c_double
        _lib.orc_logpy.restype = C.c_double
        _lib.orc_ave_log_py.restype = C.c_double
        _lib.orc_nls_reg_info.restype = C.c_int64
        _lib.orc_all_combinations.restype = C.c_int64
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


NaN = float("nan")


def log_sum_exp(x):
    x = list(x)
    assert len(x) in (2, 3)
    if len(x) == 2:
        x = [x[0], x[1], x[0]]
    return lib().orc_log_sum_exp3(*x)


def get_gc(x, alpha, call):
    lu, lm = call if call is not None else (NaN, NaN)
    return lib().orc_get_gc(int(bool(x)), alpha, lu, lm)


def get_log_Wc(a1, a2, b, call1, call2):
    W = np.zeros(4)
    l1 = call1 if call1 is not None else (NaN, NaN)
    l2 = call2 if call2 is not None else (NaN, NaN)
    lib().orc_get_log_Wc(C.c_double(a1), C.c_double(a2), C.c_double(b), C.c_double(l1[0]), C.c_double(l1[1]), C.c_double(l2[0]), C.c_double(l2[1]), _p(W))
    return W.reshape(2, 2)


def get_log_Z(al, be):
    al, be = _f(al), _f(be)
    return lib().orc_log_Z(C.c_int64(len(al)), _p(al), _p(be))


def marg_exs_log(al, be):
    al, be = _f(al), _f(be)
    L = len(al)
    logZ = C.c_double()
    ex, exx = np.zeros(L), np.zeros(L - 1)
    lib().orc_marg_exs_log(C.c_int64(L), _p(al), _p(be), C.byref(logZ), _p(ex), _p(exx))
    return logZ.value, ex, exx


def marg_px_log(al, be):
    al, be = _f(al), _f(be)
    px = np.zeros(len(al))
    lib().orc_marg_px_log(C.c_int64(len(al)), _p(al), _p(be), _p(px))
    return px


def get_cond_exs_log(al, be, lu, lm):
    al, be, lu, lm = _f(al), _f(be), _f(lu), _f(lm)
    L = len(al)
    logZc = C.c_double()
    ex, exx = np.zeros(L), np.zeros(L - 1)
    lib().orc_cond_exs_log(C.c_int64(L), _p(al), _p(be), _p(lu), _p(lm), C.byref(logZc), _p(ex), _p(exx))
    return logZc.value, ex, exx


def get_logpy(al, be, lu, lm):
    al, be, lu, lm = _f(al), _f(be), _f(lu), _f(lm)
    return lib().orc_logpy(C.c_int64(len(al)), _p(al), _p(be), _p(lu), _p(lm))


def get_ab_from_phi(phi, Nl, rho, dl):
    phi, Nl, rho, dl = _f(phi), _f(Nl), _f(rho), _f(dl)
    L = len(Nl)
    al, be = np.zeros(L), np.zeros(L - 1)
    lib().orc_get_ab_from_phi(_p(phi), C.c_int64(L), _p(Nl), _p(rho), _p(dl), _p(al), _p(be))
    return al, be


def get_grad_logZ(phi, Nl, rho, dl):
    phi, Nl, rho, dl = _f(phi), _f(Nl), _f(rho), _f(dl)
    g = np.zeros(3)
    lib().orc_grad_logZ(_p(phi), C.c_int64(len(Nl)), _p(Nl), _p(rho), _p(dl), _p(g))
    return g


def conv_check(a, b):
    a, b = _f(a), _f(b)
    return bool(lib().orc_conv_check(_p(a), _p(b), C.c_int64(len(a))))


def e_step(Nl, rho, dl, lu, lm, phi):
    """get_ess(map(get_cond_exs_log)) — em_algorithm.jl:142-146.  lu, lm: [M, L] (NaN = unobserved)."""
    Nl, rho, dl, lu, lm, phi = _f(Nl), _f(rho), _f(dl), _f(lu), _f(lm), _f(phi)
    M, L = lu.shape
    ES = np.zeros(3)
    lib().orc_e_step(C.c_int64(L), C.c_int64(M), _p(Nl), _p(rho), _p(dl), _p(lu), _p(lm), _p(phi), _p(ES))
    return ES


def e_step_wgbs(Nl, rho, dl, lu, lm, phi):
    """em_iter_wgbs's E-step: get_ess(map(get_cond_exs)) — the non-log chain of transfer_matrix.jl:191-217, em_algorithm.jl:421-424."""
    Nl, rho, dl, lu, lm, phi = _f(Nl), _f(rho), _f(dl), _f(lu), _f(lm), _f(phi)
    M, L = lu.shape
    ES = np.zeros(3)
    lib().orc_e_step_wgbs(C.c_int64(L), C.c_int64(M), _p(Nl), _p(rho), _p(dl), _p(lu), _p(lm), _p(phi), _p(ES))
    return ES


def m_step(Nl, rho, dl, M, ES, phi_init):
    Nl, rho, dl, ES, phi_init = _f(Nl), _f(rho), _f(dl), _f(ES), _f(phi_init)
    sol = np.zeros(3)
    cnt = np.zeros(4, dtype=np.int64)
    ok = lib().orc_m_step(C.c_int64(len(Nl)), C.c_int64(M), _p(Nl), _p(rho), _p(dl), _p(ES), _p(phi_init), _p(sol), _p(cnt))
    return bool(ok), sol, cnt


def em_iter(Nl, rho, dl, lu, lm, phi_init, box=(20.0, 150.0, 50.0)):
    Nl, rho, dl, lu, lm, phi_init, box = _f(Nl), _f(rho), _f(dl), _f(lu), _f(lm), _f(phi_init), _f(box)
    M, L = lu.shape
    out = np.zeros(3)
    lib().orc_em_iter(C.c_int64(L), C.c_int64(M), _p(Nl), _p(rho), _p(dl), _p(lu), _p(lm), _p(phi_init), _p(box), _p(out))
    return out


def comp_ave_log_py(Nl, rho, dl, lu, lm, phi):
    Nl, rho, dl, lu, lm, phi = _f(Nl), _f(rho), _f(dl), _f(lu), _f(lm), _f(phi)
    M, L = lu.shape
    return lib().orc_ave_log_py(C.c_int64(L), C.c_int64(M), _p(Nl), _p(rho), _p(dl), _p(lu), _p(lm), _p(phi))


def fit_batch(grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, phi0, n_init, max_em_iters=20,
              box=(20.0, 150.0, 50.0), n_threads=1):
    """get_ϕhat! over a CSR batch — same layout as cpn_fit_batch.  Returns dict."""
    grp_off = np.ascontiguousarray(grp_off, dtype=np.int64)
    read_off = np.ascontiguousarray(read_off, dtype=np.int64)
    R = len(grp_off) - 1
    Ls, Ms = np.diff(grp_off), np.diff(read_off)
    obs_off = np.concatenate([[0], np.cumsum(Ls * Ms)]).astype(np.int64)
    Nl, rho_l, d_l, lu, lm, phi0 = _f(Nl), _f(rho_l), _f(d_l), _f(log_pyx_u), _f(log_pyx_m), _f(phi0)
    assert phi0.size == R * n_init * 3
    phi_hat, Q = np.zeros((R, 3)), np.zeros(R)
    status = np.zeros(R, dtype=np.int32)
    counters = np.zeros((R, 4), dtype=np.int64)
    phi_s, Q_s = np.zeros((R, n_init, 3)), np.zeros((R, n_init))
    lib().orc_fit_batch(C.c_int64(R), _p(grp_off), _p(Nl), _p(rho_l), _p(d_l), _p(read_off), _p(obs_off), _p(lu), _p(lm), _p(phi0),
                        C.c_int32(n_init), C.c_int32(max_em_iters), C.c_double(box[0]), C.c_double(box[1]), C.c_double(box[2]),
                        _p(phi_hat), _p(Q), _p(status), _p(counters), _p(phi_s), _p(Q_s), C.c_int32(n_threads))
    return dict(phi_hat=phi_hat, Q=Q, pick=status, counters=counters, phi_starts=phi_s, Q_starts=Q_s)


def get_nls_reg_info(chrst, chrend, max_size_subreg, cpg_pos):
    cpg_pos = np.ascontiguousarray(cpg_pos, dtype=np.int64)
    K = lib().orc_nls_reg_info(C.c_int64(chrst), C.c_int64(chrend), C.c_int64(max_size_subreg), _p(cpg_pos), C.c_int64(len(cpg_pos)),
                               None, None, None, None)
    st, en, p, q = (np.zeros(K, dtype=np.int64) for _ in range(4))
    lib().orc_nls_reg_info(C.c_int64(chrst), C.c_int64(chrend), C.c_int64(max_size_subreg), _p(cpg_pos), C.c_int64(len(cpg_pos)),
                           _p(st), _p(en), _p(p), _p(q))
    return st, en, p, q


def get_stat_sums(phi, rho_n, d_n, sub_p, sub_q):
    phi, rho_n, d_n = _f(phi), _f(rho_n), _f(d_n)
    sub_p = np.ascontiguousarray(sub_p, dtype=np.int64); sub_q = np.ascontiguousarray(sub_q, dtype=np.int64)
    N, K = len(rho_n), len(sub_p)
    logZ = C.c_double()
    ex, exx = np.zeros(N), np.zeros(N - 1)
    log_g = np.zeros(4 * K)
    g1, g2, mml, nme = np.zeros(K), np.zeros(K), np.zeros(K), np.zeros(K)
    lib().orc_stat_sums(_p(phi), C.c_int64(N), _p(rho_n), _p(d_n), C.c_int64(K), _p(sub_p), _p(sub_q), C.byref(logZ), _p(ex), _p(exx),
                        _p(log_g), _p(g1), _p(g2), _p(mml), _p(nme))
    return dict(logZ=logZ.value, ex=ex, exx=exx, log_g=log_g.reshape(K, 4), e_log_g1=g1, e_log_g2=g2, mml=mml, nme=nme)


def comp_cmd(phi1, phi2, nme1, nme2, rho_n, d_n, sub_p, sub_q):
    phi1, phi2, nme1, nme2, rho_n, d_n = _f(phi1), _f(phi2), _f(nme1), _f(nme2), _f(rho_n), _f(d_n)
    sub_p = np.ascontiguousarray(sub_p, dtype=np.int64); sub_q = np.ascontiguousarray(sub_q, dtype=np.int64)
    K = len(sub_p)
    cmd = np.zeros(K)
    lib().orc_comp_cmd(_p(phi1), _p(phi2), _p(nme1), _p(nme2), C.c_int64(len(rho_n)), _p(rho_n), _p(d_n), C.c_int64(K), _p(sub_p), _p(sub_q), _p(cmd))
    return cmd


def all_combinations(n, k):
    cnt = lib().orc_all_combinations(C.c_int32(n), C.c_int32(k), None)
    out = np.zeros((cnt, k), dtype=np.int32)
    lib().orc_all_combinations(C.c_int32(n), C.c_int32(k), _p(out))
    return out


def unmat_est_reg_test(n1, n2, mml, nme, cmd_tbl, labelings, exact):
    """mml,nme [S,K]; cmd_tbl [S,S,K] (upper triangle valid); labelings [P,n1] 1-based ascending."""
    mml, nme, cmd_tbl = _f(mml), _f(nme), _f(cmd_tbl)
    labelings = np.ascontiguousarray(labelings, dtype=np.int32)
    K = mml.shape[1]
    P = labelings.shape[0]
    T, p = np.zeros(3 * K), np.zeros(3 * K)
    cnt = np.zeros(3 * K, dtype=np.int64)
    lib().orc_unmat_test(C.c_int32(n1), C.c_int32(n2), C.c_int64(K), _p(mml), _p(nme), _p(cmd_tbl), _p(labelings), C.c_int64(P), C.c_int32(int(exact)),
                         _p(T), _p(cnt), _p(p))
    return T.reshape(3, K), cnt.reshape(3, K), p.reshape(3, K)


def mat_est_reg_test(mml1, mml2, nme1, nme2, js, exact):
    mml1, mml2, nme1, nme2 = _f(mml1), _f(mml2), _f(nme1), _f(nme2)
    js = np.ascontiguousarray(js, dtype=np.int64)
    n, K = mml1.shape
    T, p = np.zeros(2 * K), np.zeros(2 * K)
    cnt = np.zeros(2 * K, dtype=np.int64)
    lib().orc_mat_test(C.c_int32(n), C.c_int64(K), _p(mml1), _p(mml2), _p(nme1), _p(nme2), _p(js), C.c_int64(len(js)), C.c_int32(int(exact)),
                       _p(T), _p(cnt), _p(p))
    return T.reshape(2, K), cnt.reshape(2, K), p.reshape(2, K)


def two_samp_pvals(T_obs, T_null, exact):
    """T_obs [3,K]; T_null [P,3,K]."""
    T_obs, T_null = _f(T_obs), _f(T_null)
    P, _, K = T_null.shape
    cnt = np.zeros(3 * K, dtype=np.int64); nv = np.zeros(3 * K, dtype=np.int64)
    p = np.zeros(3 * K)
    lib().orc_two_samp_pvals(C.c_int64(K), C.c_int64(P), _p(T_obs), _p(T_null), C.c_int32(int(exact)), _p(cnt), _p(nv), _p(p))
    return cnt.reshape(3, K), nv.reshape(3, K), p.reshape(3, K)


def nonlog_chain(al, be, lu=None, lm=None):
    al, be = _f(al), _f(be)
    L = len(al)
    Z = C.c_double()
    ex, exx = np.zeros(L), np.zeros(L - 1)
    if lu is not None:
        lu, lm = _f(lu), _f(lm)
    lib().orc_nonlog_chain(C.c_int64(L), _p(al), _p(be), _p(lu), _p(lm), C.byref(Z), _p(ex), _p(exx))
    return Z.value, ex, exx


# ---- counter-based stream of starts / permutations (restatement of csrc/cpn_philox.h's specification) -----------------------
NO_PERM = 0xFFFFFFFF


def philox4x32_10(seed, c0, c1, c2, c3):
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(C.c_uint64(seed), C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), _p(out))
    return out


def gen_phi0(seed, region, perm, half, start):
    out = np.zeros(3)
    lib().orc_gen_phi0(C.c_uint64(seed), C.c_uint32(region), C.c_uint32(perm), C.c_uint32(half), C.c_uint32(start), _p(out))
    return out


def gen_phi0s(seed, regions, n_init, perm=NO_PERM, half=0):
    """[len(regions), n_init, 3] starts of plain fits (perm = NO_PERM) or of one (perm, half) refit per region."""
    return np.array([[gen_phi0(seed, int(r), perm, half, s) for s in range(n_init)] for r in regions])


def gen_perm(seed, region, perm, n, n1):
    ids = np.zeros(n1, dtype=np.int32)
    lib().orc_gen_perm(C.c_uint64(seed), C.c_uint32(region), C.c_uint32(perm), C.c_int32(n), C.c_int32(n1), _p(ids))
    return ids
