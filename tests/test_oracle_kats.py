"""Pins the CPU oracle against the reference's own known answers (no GPU needed).

Sources: test/runtests.jl:9-76 and the docstring examples of src/transfer_matrix.jl, src/em_algorithm.jl,
src/link_functions.jl."""
import math

import numpy as np


def test_log_sum_exp_docstring(orc):
    # transfer_matrix.jl:504-505 — only max and min enter
    assert orc.log_sum_exp([1.0e-2, 1.0e-5, 1.0e-10]) == 0.6981596805576123


def test_gamma_and_Wc_docstrings(orc):
    call = (-50.0, -45.0)
    assert orc.get_gc(False, 0.0, call) == -25.0           # transfer_matrix.jl:625-629
    assert orc.get_gc(True, 0.0, call) == -22.5
    W = orc.get_log_Wc(0.0, 0.0, 0.0, call, call)          # :672-676
    assert np.array_equal(W, np.array([[-50.0, -47.5], [-47.5, -45.0]]))


def test_runtests_nonlog_marginal(orc):
    # test/runtests.jl:9-25
    L = 10
    Z, ex, exx = orc.nonlog_chain(np.zeros(L), np.zeros(L - 1))
    assert Z == 1024.0
    assert np.array_equal(ex, np.zeros(10)) and np.array_equal(exx, np.zeros(9))


def test_runtests_log_marginal(orc):
    # test/runtests.jl:27-41 — exact == on log Z
    L = 10
    logZ, ex, exx = orc.marg_exs_log(np.zeros(L), np.zeros(L - 1))
    assert logZ == math.log(1024)
    assert logZ == 6.931471805599453                        # transfer_matrix.jl:1191-1196
    assert np.allclose(ex, 0, atol=1e-5) and np.allclose(exx, 0, atol=1e-5)


def test_runtests_nonlog_conditional(orc):
    # test/runtests.jl:45-59
    L = 10
    Zc, ex, exx = orc.nonlog_chain(np.zeros(L), np.zeros(L - 1), np.full(L, -50.0), np.full(L, -50.0))
    assert Zc == 7.295566240503076e-215
    assert np.array_equal(ex, np.zeros(10)) and np.array_equal(exx, np.zeros(9))


def test_runtests_log_conditional(orc):
    # test/runtests.jl:61-76 — exact == on log Zc
    L = 10
    logZc, ex, exx = orc.get_cond_exs_log(np.zeros(L), np.zeros(L - 1), np.full(L, -100.0), np.full(L, -100.0))
    assert logZc == -993.0685281944009
    assert np.allclose(ex, 0, atol=1e-5) and np.allclose(exx, 0, atol=1e-5)


def test_cond_exs_docstrings(orc):
    # transfer_matrix.jl:935-941
    _, ex, exx = orc.get_cond_exs_log(np.zeros(2), np.zeros(1), np.full(2, -50.0), np.full(2, -45.0))
    assert np.allclose(ex, 0.9866142981513866, atol=1e-12) and np.allclose(exx, 0.9734077733168003, atol=1e-12)
    assert abs(ex[0] - math.tanh(2.5)) < 1e-12
    _, ex, exx = orc.get_cond_exs_log(np.zeros(2), np.zeros(1), np.full(2, -200.0), np.full(2, -200.0))
    assert np.all(ex == 3.552713678800501e-15) and np.all(exx == 3.552713678800501e-15)
    # :724-740 and :879-895: the characteristic noise values of the exp(·)−2 formulation
    L = 10
    _, ex, _ = orc.get_cond_exs_log(np.zeros(L), np.zeros(L - 1), np.full(L, -45.0), np.full(L, -45.0))
    assert np.all(ex == 3.552713678800501e-15)
    _, _, exx = orc.get_cond_exs_log(np.zeros(L), np.zeros(L - 1), np.full(L, -100.0), np.full(L, -100.0))
    assert np.allclose(exx[:5], -1.099120794378905e-13, atol=1e-20) and np.allclose(exx[5:], 1.1723955140041653e-13, atol=1e-20)


def test_grad_logZ_and_link_docstrings(orc):
    Nl, rho, dl = np.full(10, 5.0), np.full(10, 0.1), np.full(9, 10.0)
    assert np.array_equal(orc.get_grad_logZ([0.0, 0.0, 0.0], Nl, rho, dl), np.zeros(3))    # transfer_matrix.jl:1217-1222
    al, be = orc.get_ab_from_phi([0.0, 0.0, 0.0], Nl, rho, dl)                               # link_functions.jl:10-13
    assert np.array_equal(al, np.zeros(10)) and np.array_equal(be, np.zeros(9))
    al, be = orc.get_ab_from_phi([1.0, 2.0, 3.0], Nl, rho, dl)
    assert np.allclose(al, 5.0 * (1.0 + 2.0 * 0.1)) and np.allclose(be, 0.3)


def test_conv_check_docstring(orc):
    # em_algorithm.jl:18-23: ‖ϕ1−ϕ2‖₂ ≤ 3e-2.  (The docstring at :12-15 claims `true` for ten components of 0.01,
    # but √(10·1e-4) = 0.0316 > 0.03: the example predates the current threshold and is not executed by any test.)
    assert orc.conv_check(np.zeros(10), 0.1 + np.zeros(10)) is False
    assert orc.conv_check(np.zeros(10), 0.01 + np.zeros(10)) is False
    assert orc.conv_check(np.zeros(3), 0.01 + np.zeros(3)) is True
    assert orc.conv_check(np.zeros(3), np.array([0.03, 0.0, 0.0])) is True
    assert orc.conv_check(np.zeros(3), np.array([0.03, 0.001, 0.0])) is False


def test_marg_px(orc):
    # transfer_matrix.jl:1381-1397
    assert np.allclose(orc.marg_px_log(np.zeros(10), np.zeros(9)), 0.5, atol=1e-12)


def test_log_vs_nonlog_agree(orc):
    # calls/GeneralTesting/transfer_matrix_tests.jl:6-36 — the two formulations agree on a random chain
    rng = np.random.default_rng(0)
    L = 12
    al, be = rng.normal(0, 0.5, L), rng.normal(0, 0.5, L - 1)
    logZ, ex, exx = orc.marg_exs_log(al, be)
    Z, ex2, exx2 = orc.nonlog_chain(al, be)
    assert abs(logZ - math.log(Z)) < 1e-12
    assert np.allclose(ex, ex2, atol=1e-12) and np.allclose(exx, exx2, atol=1e-12)
    # brute force
    import itertools
    xs = np.array(list(itertools.product([-1, 1], repeat=L)))
    w = np.exp(xs @ al + (xs[:, :-1] * xs[:, 1:]) @ be)
    assert abs(logZ - math.log(w.sum())) < 1e-12
    assert np.allclose(ex, (w[:, None] * xs).sum(0) / w.sum(), atol=1e-12)


def test_logpy_matches_bruteforce(orc):
    rng = np.random.default_rng(1)
    L = 8
    al, be = rng.normal(0, 0.7, L), rng.normal(0, 0.7, L - 1)
    lu, lm = rng.normal(-3, 1, L), rng.normal(-3, 1, L)
    lu[2] = lm[2] = np.nan
    import itertools
    xs = np.array(list(itertools.product([-1, 1], repeat=L)))
    logw = xs @ al + (xs[:, :-1] * xs[:, 1:]) @ be
    ll = np.where(xs == 1, np.nan_to_num(lm), np.nan_to_num(lu)).sum(1)
    ref = np.log(np.exp(logw + ll).sum()) - np.log(np.exp(logw).sum())
    assert abs(orc.get_logpy(al, be, lu, lm) - ref) < 1e-12


def test_philox_known_answers_and_product_stream(orc, cpn):
    """Philox-4x32-10 known-answer vectors of the Random123 distribution (kat_vectors: zero, all-ones, digits of pi) pin the
    oracle's restatement; the library's host-side evaluation of the stream (cpn_rng_*, the same header the device kernels
    include) must then agree with the oracle's independent restatement on starts and permutations."""
    assert [hex(x) for x in orc.philox4x32_10(0, 0, 0, 0, 0)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in orc.philox4x32_10(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)] == \
        ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in orc.philox4x32_10((0x299f31d0 << 32) | 0xa4093822, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    lib = cpn.Library()
    rng = np.random.default_rng(1)
    for _ in range(200):
        seed, region, perm = int(rng.integers(0, 2**63)), int(rng.integers(0, 3_000_000)), int(rng.integers(0, 1000))
        a = cpn.capi.rng_phi0(lib, seed, region, 10)
        assert np.array_equal(a, orc.gen_phi0s(seed, [region], 10)[0])
        assert np.all((a[:, 0] >= -5) & (a[:, 0] <= 5) & (a[:, 1] >= -50) & (a[:, 1] <= 50) & (a[:, 2] >= 0) & (a[:, 2] <= 10))
        assert np.allclose(a * 100, np.rint(a * 100), atol=1e-9)                                    # on the 0.01 grid of gen_ϕ0s
        h = int(rng.integers(0, 2))
        assert np.array_equal(cpn.capi.rng_phi0(lib, seed, region, 4, perm, h), np.array([orc.gen_phi0(seed, region, perm, h, s) for s in range(4)]))
        n = int(rng.integers(2, 130)); n1 = int(rng.integers(1, n))
        ids = cpn.capi.rng_perm(lib, seed, region, perm, n, n1)
        assert np.array_equal(ids, orc.gen_perm(seed, region, perm, n, n1))
        assert len(ids) == n1 and np.all(np.diff(ids) > 0) and ids[0] >= 1 and ids[-1] <= n
    # uniformity of the subset draw: every read is in sample 1 with probability n1/n
    hits = np.zeros(12)
    for i in range(4000):
        hits[cpn.capi.rng_perm(lib, 5, 0, i, 12, 5) - 1] += 1
    assert np.all(np.abs(hits / 4000 - 5 / 12) < 0.04)
