"""ctypes binding of libcpelnano_b200.so — one Python function per symbol of include/cpelnano_b200.h.

Arrays are numpy (host) arrays unless the function name ends in `_device`, in which case raw device
addresses (ints, e.g. torch.Tensor.data_ptr()) are passed.  No computation happens in Python.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MSTEP_FD, MSTEP_ANALYTIC = 0, 1
FIT_NONE, FIT_NUMERICAL, FIT_FILTERED = -1, -2, -16        # status words (include/cpelnano_b200.h)

# every exported symbol of include/cpelnano_b200.h (tests check the library exports exactly these)
SYMBOLS = ["cpn_create", "cpn_create_multi", "cpn_num_devices", "cpn_destroy", "cpn_last_error", "cpn_version", "cpn_last_device_ms", "cpn_last_launches", "cpn_last_work",
           "cpn_measure_fp64_peak", "cpn_set_seed", "cpn_rng_phi0", "cpn_rng_perm", "cpn_set_profile", "cpn_get_profile", "cpn_analyze_batch", "cpn_analyze_batch_device", "cpn_fit_batch", "cpn_fit_batch_device", "cpn_pack_emissions", "cpn_fit_batch_packed", "cpn_analyze_batch_packed", "cpn_wgbs_pack", "cpn_analyze_wgbs_batch", "cpn_estep_batch", "cpn_estep_batch_starts", "cpn_mstep_batch",
           "cpn_stat_sums_batch", "cpn_nls_reg_info", "cpn_cmd_batch", "cpn_grp_unmat_sweep", "cpn_grp_mat_sweep", "cpn_two_samp_pvals", "cpn_two_samp_null_batch", "cpn_two_samp_test_batch", "cpn_grp_test_batch", "cpn_grp_test_batch_device", "cpn_region_filter",
           "cpn_calls_open", "cpn_calls_close", "cpn_calls_open_error", "cpn_calls_error", "cpn_calls_num_lines", "cpn_calls_num_reads", "cpn_calls_num_chr",
           "cpn_calls_chr_name", "cpn_calls_chr_lines", "cpn_calls_read_name", "cpn_calls_count_reads", "cpn_calls_fill",
           "cpn_cpg_scan", "cpn_chr_part", "cpn_grp_info_count", "cpn_grp_info_fill", "cpn_format_f64", "cpn_write_rows", "cpn_write_theta", "cpn_nls_reg_info_batch", "cpn_calls_coverage", "cpn_fasta_open", "cpn_fasta_close", "cpn_fasta_open_error",
           "cpn_fasta_num_records", "cpn_fasta_name", "cpn_fasta_length", "cpn_fasta_seq", "cpn_add_qvalues", "cpn_theta_open", "cpn_theta_close", "cpn_theta_open_error",
           "cpn_theta_num_rows", "cpn_theta_num_chr", "cpn_theta_chr_name", "cpn_theta_copy",
           "cpn_bam_open", "cpn_bam_close", "cpn_bam_open_error", "cpn_bam_error", "cpn_bam_num_records", "cpn_bam_num_refs", "cpn_bam_ref_name", "cpn_bam_ref_len",
           "cpn_bam_record", "cpn_bam_count_reads", "cpn_bam_fill_calls"]


class CpnError(RuntimeError):
    pass


def lib_path():
    """In-tree library; CPN_LIB overrides it (kernel-tuning builds under build/variants/)."""
    return os.environ.get("CPN_LIB") or os.path.join(_HERE, "libcpelnano_b200.so")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


class Library:
    """Loads the shared library (fails loudly if it has not been built) and owns one cpn_ctx per device."""

    def __init__(self, path=None):
        path = path or lib_path()
        if not os.path.exists(path):
            raise CpnError(f"{path} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback.")
        self.dll = C.CDLL(path)
        d = self.dll
        d.cpn_version.restype = C.c_char_p
        d.cpn_last_error.restype = C.c_char_p
        d.cpn_last_error.argtypes = [C.c_void_p]
        d.cpn_last_device_ms.restype = C.c_double
        d.cpn_last_device_ms.argtypes = [C.c_void_p]
        d.cpn_last_launches.restype = C.c_int64
        d.cpn_last_launches.argtypes = [C.c_void_p]
        d.cpn_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        d.cpn_create_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        d.cpn_destroy.argtypes = [C.c_void_p]
        d.cpn_nls_reg_info.restype = C.c_int64
        for f in ("cpn_calls_open_error", "cpn_calls_error", "cpn_calls_chr_name", "cpn_calls_read_name"):
            getattr(d, f).restype = C.c_char_p
        for f in ("cpn_calls_num_lines", "cpn_calls_num_reads", "cpn_calls_num_chr", "cpn_calls_chr_lines"):
            getattr(d, f).restype = C.c_int64
        d.cpn_cpg_scan.restype = d.cpn_chr_part.restype = d.cpn_format_f64.restype = C.c_int64
        d.cpn_calls_open.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]
        d.cpn_calls_close.argtypes = [C.c_void_p]
        d.cpn_calls_error.argtypes = [C.c_void_p]
        d.cpn_calls_num_lines.argtypes = d.cpn_calls_num_reads.argtypes = d.cpn_calls_num_chr.argtypes = [C.c_void_p]
        d.cpn_calls_chr_name.argtypes = d.cpn_calls_chr_lines.argtypes = d.cpn_calls_read_name.argtypes = [C.c_void_p, C.c_int64]
        self._ctx = {}

    def version(self):
        return self.dll.cpn_version().decode()

    def ctx(self, device=0):
        """The context of one device (int) or ONE context spanning several devices (tuple / list of ints: cpn_create_multi — the
        host-pointer batch entry points then shard every call over those GPUs from this single process)."""
        if isinstance(device, (list, tuple)):
            device = tuple(int(d) for d in device)
        if device not in self._ctx:
            h = C.c_void_p()
            if isinstance(device, tuple):
                ids = (C.c_int * len(device))(*device)
                rc = self.dll.cpn_create_multi(ids, C.c_int(len(device)), C.byref(h))
            else:
                rc = self.dll.cpn_create(int(device), C.byref(h))
            if rc != 0 or not h.value:
                raise CpnError(f"cpn_create(device={device}) failed with code {rc}: no usable sm_100 GPU (there is no CPU fallback)")
            self._ctx[device] = h
        return self._ctx[device]

    def close(self):
        for h in self._ctx.values():
            self.dll.cpn_destroy(h)
        self._ctx = {}

    def _check(self, rc, ctx):
        if rc != 0:
            raise CpnError(f"libcpelnano_b200 error {rc}: {self.dll.cpn_last_error(ctx).decode()}")

    def last_device_ms(self, device=0):
        return self.dll.cpn_last_device_ms(self.ctx(device))

    def last_launches(self, device=0):
        return self.dll.cpn_last_launches(self.ctx(device))

    def last_work(self, device=0):
        """(E-step read·groups, M-step sweep·groups, score read·groups, EM instances) of the last fit call (cpn_last_work)."""
        w = np.zeros(4)
        ctx = self.ctx(device)
        self._check(self.dll.cpn_last_work(ctx, _ptr(w)), ctx)
        return w

    def measure_fp64_peak(self, device=0):
        v = C.c_double()
        ctx = self.ctx(device)
        self._check(self.dll.cpn_measure_fp64_peak(ctx, C.byref(v)), ctx)
        return v.value

    def set_seed(self, seed, device=0):
        """Seed of the on-device start / permutation stream used wherever phi0 / perm_ids are None (cpn_set_seed)."""
        ctx = self.ctx(device)
        self._check(self.dll.cpn_set_seed(ctx, C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)), ctx)

    KERNEL_CLASSES = ("prep", "estep", "mstep", "score", "pick", "statsum", "cmd", "sweep")

    def set_profile(self, enable, device=0):
        ctx = self.ctx(device)
        self._check(self.dll.cpn_set_profile(ctx, C.c_int(int(enable))), ctx)

    def get_profile(self, device=0):
        ms = np.zeros(8)
        n = np.zeros(8, dtype=np.int64)
        ctx = self.ctx(device)
        self._check(self.dll.cpn_get_profile(ctx, _ptr(ms), _ptr(n)), ctx)
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    def pack_emissions(self, grp_off, read_off, log_pyx_u, log_pyx_m, n_threads=0, out=None):
        """cpn_pack_emissions: (llr [Σ M·L], read_base [Σ M]) from the two log-likelihood tables (host, no CUDA)."""
        grp_off, read_off = _i64(grp_off), _i64(read_off)
        lu, lm = _f64(log_pyx_u), _f64(log_pyx_m)
        R = len(grp_off) - 1
        llr, base = out if out is not None else (np.empty(lu.size), np.empty(int(read_off[-1])))
        rc = self.dll.cpn_pack_emissions(C.c_int64(R), _ptr(grp_off), _ptr(read_off), _ptr(lu), _ptr(lm), _ptr(llr), _ptr(base), C.c_int32(n_threads))
        if rc != 0:
            raise CpnError(f"cpn_pack_emissions failed with code {rc} (null pointer, negative extent, or an observed cell without a finite log p(y|+1))")
        return llr, base

    def fit_batch_packed(self, grp_off, Nl, rho_l, d_l, read_off, llr, read_base, phi0, n_init, max_em_iters=20,
                         box=(20.0, 150.0, 50.0), mstep_mode=MSTEP_ANALYTIC, per_start=False, device=0):
        """cpn_fit_batch_packed: cpn_fit_batch on packed emissions (pack_emissions)."""
        grp_off, read_off = _i64(grp_off), _i64(read_off)
        Nl, rho_l, d_l, llr, read_base = _f64(Nl), _f64(rho_l), _f64(d_l), _f64(llr), _f64(read_base)
        R = len(grp_off) - 1
        if phi0 is not None:
            phi0 = _f64(phi0)
            assert phi0.size == R * n_init * 3, "phi0 must be [R, n_init, 3]"
        phi_hat, Q = np.empty((R, 3)), np.empty(R)
        status = np.empty(R, dtype=np.int32)
        counters = np.empty((R, 4), dtype=np.int64)
        ps = np.empty((R, n_init, 3)) if per_start else None
        qs = np.empty((R, n_init)) if per_start else None
        ctx = self.ctx(device)
        rc = self.dll.cpn_fit_batch_packed(ctx, C.c_int64(R), _ptr(grp_off), _ptr(Nl), _ptr(rho_l), _ptr(d_l), _ptr(read_off), _ptr(llr),
                                           _ptr(read_base), _ptr(phi0), C.c_int32(n_init), C.c_int32(max_em_iters), C.c_double(box[0]),
                                           C.c_double(box[1]), C.c_double(box[2]), C.c_int32(mstep_mode), _ptr(phi_hat), _ptr(Q), _ptr(status),
                                           _ptr(counters), _ptr(ps), _ptr(qs))
        self._check(rc, ctx)
        out = dict(phi_hat=phi_hat, Q=Q, pick=status, counters=counters)
        if per_start:
            out.update(phi_starts=ps, Q_starts=qs)
        return out

    def wgbs_pack(self, cpg_off, read_off, calls, n_threads=0):
        """cpn_wgbs_pack: WGBS calls (int8 −1/0/+1 per read and CpG) → packed hard emissions (llr, read_base)."""
        cpg_off, read_off = _i64(cpg_off), _i64(read_off)
        calls = np.ascontiguousarray(calls, dtype=np.int8)
        R = len(cpg_off) - 1
        llr, base = np.empty(calls.size), np.empty(int(read_off[-1]))
        rc = self.dll.cpn_wgbs_pack(C.c_int64(R), _ptr(cpg_off), _ptr(read_off), _ptr(calls), _ptr(llr), _ptr(base), C.c_int32(n_threads))
        if rc != 0:
            raise CpnError(f"cpn_wgbs_pack failed with code {rc} (calls must be -1, 0 or +1)")
        return llr, base

    def analyze_wgbs_batch(self, cpg_off, rho_n, d_n, read_off, calls, phi0, n_init, sub_off, sub_p, sub_q, max_em_iters=20,
                           box=(20.0, 150.0, 50.0), mstep_mode=MSTEP_ANALYTIC, device=0):
        """cpn_analyze_wgbs_batch: batched pmap_analyze_reg_bam (get_ϕhat_wgbs! + get_stat_sums!) from per-CpG calls."""
        cpg_off, read_off, sub_off, sub_p, sub_q = (_i64(a) for a in (cpg_off, read_off, sub_off, sub_p, sub_q))
        rho_n, d_n = _f64(rho_n), _f64(d_n)
        calls = np.ascontiguousarray(calls, dtype=np.int8)
        R = len(cpg_off) - 1
        if phi0 is not None:
            phi0 = _f64(phi0)
            assert phi0.size == R * n_init * 3
        sN, sK = int(cpg_off[-1]), int(sub_off[-1])
        out = dict(phi_hat=np.empty((R, 3)), Q=np.empty(R), pick=np.empty(R, dtype=np.int32), counters=np.empty((R, 4), dtype=np.int64),
                   logZ=np.empty(R), ex=np.empty(sN), exx=np.empty(sN - R), log_g=np.empty((sK, 4)), e_log_g1=np.empty(sK),
                   e_log_g2=np.empty(sK), mml=np.empty(sK), nme=np.empty(sK))
        ctx = self.ctx(device)
        rc = self.dll.cpn_analyze_wgbs_batch(ctx, C.c_int64(R), _ptr(cpg_off), _ptr(rho_n), _ptr(d_n), _ptr(read_off), _ptr(calls), _ptr(phi0),
                                             C.c_int32(n_init), C.c_int32(max_em_iters), C.c_double(box[0]), C.c_double(box[1]), C.c_double(box[2]),
                                             C.c_int32(mstep_mode), _ptr(sub_off), _ptr(sub_p), _ptr(sub_q), _ptr(out["phi_hat"]), _ptr(out["Q"]),
                                             _ptr(out["pick"]), _ptr(out["counters"]), _ptr(out["logZ"]), _ptr(out["ex"]), _ptr(out["exx"]),
                                             _ptr(out["log_g"]), _ptr(out["e_log_g1"]), _ptr(out["e_log_g2"]), _ptr(out["mml"]), _ptr(out["nme"]))
        self._check(rc, ctx)
        return out

    def analyze_batch_raw(self, R, grp_off, Nl, rho_l, d_l, read_off, lu, lm, phi0, n_init, max_em_iters, box, mstep_mode, cpg_off, rho_n,
                          d_n, sub_off, sub_p, sub_q, phi_hat, Q, status, counters, logZ, ex, exx, log_g, e1, e2, mml, nme, on_device, device=0,
                          packed=False):
        """Pointer-level cpn_analyze_batch[_device] (ints = raw addresses, or numpy arrays).  packed=True: cpn_analyze_batch_packed —
        `lu` holds the log-likelihood ratios and `lm` the per-read base terms of pack_emissions (host pointers only)."""
        ctx = self.ctx(device)
        assert not (packed and on_device), "the packed entry point takes host pointers"
        fn = self.dll.cpn_analyze_batch_packed if packed else (self.dll.cpn_analyze_batch_device if on_device else self.dll.cpn_analyze_batch)
        rc = fn(ctx, C.c_int64(R), _ptr(grp_off), _ptr(Nl), _ptr(rho_l), _ptr(d_l), _ptr(read_off), _ptr(lu), _ptr(lm), _ptr(phi0),
                C.c_int32(n_init), C.c_int32(max_em_iters), C.c_double(box[0]), C.c_double(box[1]), C.c_double(box[2]), C.c_int32(mstep_mode),
                _ptr(cpg_off), _ptr(rho_n), _ptr(d_n), _ptr(sub_off), _ptr(sub_p), _ptr(sub_q), _ptr(phi_hat), _ptr(Q), _ptr(status),
                _ptr(counters), _ptr(logZ), _ptr(ex), _ptr(exx), _ptr(log_g), _ptr(e1), _ptr(e2), _ptr(mml), _ptr(nme))
        self._check(rc, ctx)

    def analyze_batch(self, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, phi0, n_init, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q,
                      max_em_iters=20, box=(20.0, 150.0, 50.0), mstep_mode=MSTEP_ANALYTIC, device=0):
        grp_off, read_off, cpg_off, sub_off, sub_p, sub_q = (_i64(a) for a in (grp_off, read_off, cpg_off, sub_off, sub_p, sub_q))
        Nl, rho_l, d_l, lu, lm, phi0, rho_n, d_n = (_f64(a) for a in (Nl, rho_l, d_l, log_pyx_u, log_pyx_m, phi0, rho_n, d_n))
        R = len(grp_off) - 1
        sN, sK = int(cpg_off[-1]), int(sub_off[-1])
        out = dict(phi_hat=np.empty((R, 3)), Q=np.empty(R), pick=np.empty(R, dtype=np.int32), counters=np.empty((R, 4), dtype=np.int64),
                   logZ=np.empty(R), ex=np.empty(sN), exx=np.empty(sN - R), log_g=np.empty((sK, 4)), e_log_g1=np.empty(sK),
                   e_log_g2=np.empty(sK), mml=np.empty(sK), nme=np.empty(sK))
        self.analyze_batch_raw(R, grp_off, Nl, rho_l, d_l, read_off, lu, lm, phi0, n_init, max_em_iters, box, mstep_mode, cpg_off, rho_n, d_n,
                               sub_off, sub_p, sub_q, out["phi_hat"], out["Q"], out["pick"], out["counters"], out["logZ"], out["ex"], out["exx"],
                               out["log_g"], out["e_log_g1"], out["e_log_g2"], out["mml"], out["nme"], False, device)
        return out

    # ---- EM -------------------------------------------------------------------------------------
    def fit_batch(self, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, phi0, n_init, max_em_iters=20,
                  box=(20.0, 150.0, 50.0), mstep_mode=MSTEP_ANALYTIC, per_start=False, device=0):
        grp_off, read_off = _i64(grp_off), _i64(read_off)
        Nl, rho_l, d_l, lu, lm = _f64(Nl), _f64(rho_l), _f64(d_l), _f64(log_pyx_u), _f64(log_pyx_m)
        R = len(grp_off) - 1
        if phi0 is not None:                       # None: starts drawn on the device from the context's seed (set_seed)
            phi0 = _f64(phi0)
            assert phi0.size == R * n_init * 3, "phi0 must be [R, n_init, 3]"
        phi_hat, Q = np.empty((R, 3)), np.empty(R)
        status = np.empty(R, dtype=np.int32)
        counters = np.empty((R, 4), dtype=np.int64)
        ps = np.empty((R, n_init, 3)) if per_start else None
        qs = np.empty((R, n_init)) if per_start else None
        ctx = self.ctx(device)
        rc = self.dll.cpn_fit_batch(ctx, C.c_int64(R), _ptr(grp_off), _ptr(Nl), _ptr(rho_l), _ptr(d_l), _ptr(read_off), _ptr(lu), _ptr(lm),
                                    _ptr(phi0), C.c_int32(n_init), C.c_int32(max_em_iters), C.c_double(box[0]), C.c_double(box[1]),
                                    C.c_double(box[2]), C.c_int32(mstep_mode), _ptr(phi_hat), _ptr(Q), _ptr(status), _ptr(counters),
                                    _ptr(ps), _ptr(qs))
        self._check(rc, ctx)
        out = dict(phi_hat=phi_hat, Q=Q, pick=status, counters=counters)
        if per_start:
            out.update(phi_starts=ps, Q_starts=qs)
        return out

    def fit_batch_raw(self, R, grp_off, Nl, rho_l, d_l, read_off, lu, lm, phi0, n_init, max_em_iters, box, mstep_mode, phi_hat, Q,
                      status, counters, on_device, device=0):
        """Pointer-level call (ints or numpy arrays) used by bench.py with pinned host or device buffers."""
        ctx = self.ctx(device)
        fn = self.dll.cpn_fit_batch_device if on_device else self.dll.cpn_fit_batch
        rc = fn(ctx, C.c_int64(R), _ptr(grp_off), _ptr(Nl), _ptr(rho_l), _ptr(d_l), _ptr(read_off), _ptr(lu), _ptr(lm), _ptr(phi0),
                C.c_int32(n_init), C.c_int32(max_em_iters), C.c_double(box[0]), C.c_double(box[1]), C.c_double(box[2]), C.c_int32(mstep_mode),
                _ptr(phi_hat), _ptr(Q), _ptr(status), _ptr(counters), None, None)
        self._check(rc, ctx)

    def estep_batch(self, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, phi, with_logpy=True, device=0):
        grp_off, read_off = _i64(grp_off), _i64(read_off)
        Nl, rho_l, d_l, lu, lm, phi = _f64(Nl), _f64(rho_l), _f64(d_l), _f64(log_pyx_u), _f64(log_pyx_m), _f64(phi)
        R = len(grp_off) - 1
        ES = np.empty((R, 3))
        lp = np.empty(R) if with_logpy else None
        ctx = self.ctx(device)
        rc = self.dll.cpn_estep_batch(ctx, C.c_int64(R), _ptr(grp_off), _ptr(Nl), _ptr(rho_l), _ptr(d_l), _ptr(read_off), _ptr(lu), _ptr(lm),
                                      _ptr(phi), _ptr(ES), _ptr(lp))
        self._check(rc, ctx)
        return ES, lp

    def estep_batch_starts(self, grp_off, Nl, rho_l, d_l, read_off, log_pyx_u, log_pyx_m, phi, device=0):
        """phi [R, n_init, 3] → ES [R, n_init, 3] through the paired E-step (k_make_pairs → k_estep_pair)."""
        grp_off, read_off = _i64(grp_off), _i64(read_off)
        Nl, rho_l, d_l, lu, lm, phi = _f64(Nl), _f64(rho_l), _f64(d_l), _f64(log_pyx_u), _f64(log_pyx_m), _f64(phi)
        R = len(grp_off) - 1
        assert phi.ndim == 3 and phi.shape[0] == R and phi.shape[2] == 3
        n_init = phi.shape[1]
        ES = np.empty((R, n_init, 3))
        ctx = self.ctx(device)
        rc = self.dll.cpn_estep_batch_starts(ctx, C.c_int64(R), _ptr(grp_off), _ptr(Nl), _ptr(rho_l), _ptr(d_l), _ptr(read_off), _ptr(lu),
                                             _ptr(lm), C.c_int32(n_init), _ptr(phi), _ptr(ES))
        self._check(rc, ctx)
        return ES

    def mstep_batch(self, grp_off, Nl, rho_l, d_l, n_reads, ES, phi_init, mstep_mode=MSTEP_ANALYTIC, device=0):
        grp_off, n_reads = _i64(grp_off), _i64(n_reads)
        Nl, rho_l, d_l, ES, phi_init = _f64(Nl), _f64(rho_l), _f64(d_l), _f64(ES), _f64(phi_init)
        R = len(grp_off) - 1
        sol = np.empty((R, 3))
        conv = np.empty(R, dtype=np.int32)
        cnt = np.empty((R, 2), dtype=np.int64)
        ctx = self.ctx(device)
        rc = self.dll.cpn_mstep_batch(ctx, C.c_int64(R), _ptr(grp_off), _ptr(Nl), _ptr(rho_l), _ptr(d_l), _ptr(n_reads), _ptr(ES), _ptr(phi_init),
                                      C.c_int32(mstep_mode), _ptr(sol), _ptr(conv), _ptr(cnt))
        self._check(rc, ctx)
        return sol, conv.astype(bool), cnt

    # ---- summaries ------------------------------------------------------------------------------
    def nls_reg_info(self, chrst, chrend, max_size_subreg, cpg_pos):
        cpg_pos = _i64(cpg_pos)
        N = len(cpg_pos)
        K = self.dll.cpn_nls_reg_info(C.c_int64(chrst), C.c_int64(chrend), C.c_int64(max_size_subreg), _ptr(cpg_pos), C.c_int64(N), None, None, None, None)
        if K < 0:
            raise CpnError("cpn_nls_reg_info: bad arguments")
        st, en, p, q = (np.empty(K, dtype=np.int64) for _ in range(4))
        self.dll.cpn_nls_reg_info(C.c_int64(chrst), C.c_int64(chrend), C.c_int64(max_size_subreg), _ptr(cpg_pos), C.c_int64(N), _ptr(st), _ptr(en), _ptr(p), _ptr(q))
        return st, en, p, q

    def stat_sums_batch(self, phi, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, reg_of=None, device=0):
        phi, rho_n, d_n = _f64(phi), _f64(rho_n), _f64(d_n)
        cpg_off, sub_off, sub_p, sub_q = _i64(cpg_off), _i64(sub_off), _i64(sub_p), _i64(sub_q)
        R = len(cpg_off) - 1
        T = phi.size // 3
        if reg_of is not None:
            reg_of = _i64(reg_of)
            ex_off = np.concatenate([[0], np.cumsum(np.diff(cpg_off)[reg_of])]).astype(np.int64)
            k_off = np.concatenate([[0], np.cumsum(np.diff(sub_off)[reg_of])]).astype(np.int64)
        else:
            ex_off, k_off = cpg_off, sub_off
        sN, sK = int(ex_off[-1]), int(k_off[-1])
        out = dict(logZ=np.empty(T), ex=np.empty(sN), exx=np.empty(sN - T), log_g=np.empty((sK, 4)), e_log_g1=np.empty(sK),
                   e_log_g2=np.empty(sK), mml=np.empty(sK), nme=np.empty(sK), ex_off=ex_off, k_off=k_off)
        ctx = self.ctx(device)
        rc = self.dll.cpn_stat_sums_batch(ctx, C.c_int64(T), _ptr(phi), _ptr(reg_of), _ptr(ex_off if reg_of is not None else None),
                                          _ptr(k_off if reg_of is not None else None), C.c_int64(R), _ptr(cpg_off), _ptr(rho_n), _ptr(d_n),
                                          _ptr(sub_off), _ptr(sub_p), _ptr(sub_q), _ptr(out["logZ"]), _ptr(out["ex"]), _ptr(out["exx"]),
                                          _ptr(out["log_g"]), _ptr(out["e_log_g1"]), _ptr(out["e_log_g2"]), _ptr(out["mml"]), _ptr(out["nme"]))
        self._check(rc, ctx)
        return out

    def cmd_batch(self, pair_a, pair_b, phi_tbl, reg_of, ex_off, ex_tbl, exx_tbl, nme_off, nme_tbl, cpg_off, rho_n, d_n, sub_off, sub_p,
                  sub_q, device=0):
        pair_a, pair_b, reg_of, ex_off, nme_off = _i64(pair_a), _i64(pair_b), _i64(reg_of), _i64(ex_off), _i64(nme_off)
        phi_tbl, ex_tbl, exx_tbl, nme_tbl, rho_n, d_n = _f64(phi_tbl), _f64(ex_tbl), _f64(exx_tbl), _f64(nme_tbl), _f64(rho_n), _f64(d_n)
        cpg_off, sub_off, sub_p, sub_q = _i64(cpg_off), _i64(sub_off), _i64(sub_p), _i64(sub_q)
        P, T, R = len(pair_a), len(reg_of), len(cpg_off) - 1
        Ks = np.diff(sub_off)[reg_of[pair_a]] if P else np.zeros(0, dtype=np.int64)
        cmd_off = np.concatenate([[0], np.cumsum(Ks)]).astype(np.int64)
        cmd = np.empty(int(cmd_off[-1]))
        ctx = self.ctx(device)
        rc = self.dll.cpn_cmd_batch(ctx, C.c_int64(P), _ptr(pair_a), _ptr(pair_b), C.c_int64(T), _ptr(phi_tbl), _ptr(reg_of), _ptr(ex_off),
                                    _ptr(ex_tbl), _ptr(exx_tbl), _ptr(nme_off), _ptr(nme_tbl), C.c_int64(R), _ptr(cpg_off), _ptr(rho_n), _ptr(d_n),
                                    _ptr(sub_off), _ptr(sub_p), _ptr(sub_q), _ptr(cmd_off), _ptr(cmd))
        self._check(rc, ctx)
        return cmd, cmd_off

    # ---- sweeps ---------------------------------------------------------------------------------
    def grp_unmat_sweep(self, n1, n2, tbl_off, mml, nme, cmd_tbl, labelings, exact, device=0):
        tbl_off = _i64(tbl_off)
        mml, nme, cmd_tbl = _f64(mml), _f64(nme), _f64(cmd_tbl)
        labelings = _i32(labelings)
        R, sK = len(tbl_off) - 1, int(tbl_off[-1])
        P = labelings.shape[0] if labelings.ndim == 2 else labelings.size // max(n1, 1)
        T, p = np.empty(3 * sK), np.empty(3 * sK)
        cnt = np.empty(3 * sK, dtype=np.int64)
        ctx = self.ctx(device)
        rc = self.dll.cpn_grp_unmat_sweep(ctx, C.c_int64(R), C.c_int32(n1), C.c_int32(n2), _ptr(tbl_off), _ptr(mml), _ptr(nme), _ptr(cmd_tbl),
                                          _ptr(labelings), C.c_int64(P), C.c_int32(int(exact)), _ptr(T), _ptr(cnt), _ptr(p))
        self._check(rc, ctx)
        return T, cnt, p

    def grp_mat_sweep(self, n, tbl_off, mml1, mml2, nme1, nme2, js, exact, device=0):
        tbl_off, js = _i64(tbl_off), _i64(js)
        mml1, mml2, nme1, nme2 = _f64(mml1), _f64(mml2), _f64(nme1), _f64(nme2)
        R, sK = len(tbl_off) - 1, int(tbl_off[-1])
        T, p = np.empty(2 * sK), np.empty(2 * sK)
        cnt = np.empty(2 * sK, dtype=np.int64)
        ctx = self.ctx(device)
        rc = self.dll.cpn_grp_mat_sweep(ctx, C.c_int64(R), C.c_int32(n), _ptr(tbl_off), _ptr(mml1), _ptr(mml2), _ptr(nme1), _ptr(nme2), _ptr(js),
                                        C.c_int64(len(js)), C.c_int32(int(exact)), _ptr(T), _ptr(cnt), _ptr(p))
        self._check(rc, ctx)
        return T, cnt, p

    def grp_test_batch(self, n1, n2, matched, phi, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, labelings, exact, device=0):
        """cpn_grp_test_batch: phi [R, S, 3] → unmatched (T, cnt, p) [ΣK·3]; matched (T, cnt, p) [ΣK·2] and tcmd [ΣK]."""
        phi, rho_n, d_n = _f64(phi), _f64(rho_n), _f64(d_n)
        cpg_off, sub_off, sub_p, sub_q = _i64(cpg_off), _i64(sub_off), _i64(sub_p), _i64(sub_q)
        labelings = _i64(labelings) if matched else _i32(labelings)
        R, sK = len(cpg_off) - 1, int(sub_off[-1])
        P = len(labelings) if matched else (labelings.shape[0] if labelings.ndim == 2 else labelings.size // max(n1, 1))
        ns = 2 if matched else 3
        T, p, cnt = np.empty(ns * sK), np.empty(ns * sK), np.empty(ns * sK, dtype=np.int64)
        tcmd = np.empty(sK)
        ctx = self.ctx(device)
        rc = self.dll.cpn_grp_test_batch(ctx, C.c_int64(R), C.c_int32(n1), C.c_int32(n2), C.c_int32(int(matched)), _ptr(phi), _ptr(cpg_off),
                                         _ptr(rho_n), _ptr(d_n), _ptr(sub_off), _ptr(sub_p), _ptr(sub_q), _ptr(labelings), C.c_int64(P),
                                         C.c_int32(int(exact)), _ptr(T), _ptr(cnt), _ptr(p), _ptr(tcmd))
        self._check(rc, ctx)
        return (T, cnt, p, tcmd) if matched else (T, cnt, p)

    def grp_test_batch_raw(self, R, n1, n2, matched, phi, cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, labelings, P, exact, T, cnt, pval, tcmd,
                           on_device, device=0):
        """Pointer-level cpn_grp_test_batch[_device]: ints = raw addresses (pinned host or device), numpy arrays = host; cpg_off,
        sub_off and labelings are always host numpy arrays."""
        ctx = self.ctx(device)
        fn = self.dll.cpn_grp_test_batch_device if on_device else self.dll.cpn_grp_test_batch
        rc = fn(ctx, C.c_int64(R), C.c_int32(n1), C.c_int32(n2), C.c_int32(int(matched)), _ptr(phi), _ptr(cpg_off), _ptr(rho_n), _ptr(d_n),
                _ptr(sub_off), _ptr(sub_p), _ptr(sub_q), _ptr(labelings), C.c_int64(P), C.c_int32(int(exact)), _ptr(T), _ptr(cnt), _ptr(pval), _ptr(tcmd))
        self._check(rc, ctx)

    def two_samp_null_batch(self, grp_off, Nl, rho_l, d_l, read_off, lu, lm, perm_off, n1, perm_ids, phi0, n_init, cpg_off, rho_n, d_n,
                            sub_off, sub_p, sub_q, max_em_iters=20, box=(20.0, 150.0, 50.0), mstep_mode=MSTEP_ANALYTIC, device=0):
        """cpn_two_samp_null_batch: returns (T_null flat [Σ P_r·3·K_r], fit_status [Σ P_r·2])."""
        grp_off, read_off, perm_off, n1 = _i64(grp_off), _i64(read_off), _i64(perm_off), _i64(n1)
        Nl, rho_l, d_l, lu, lm = _f64(Nl), _f64(rho_l), _f64(d_l), _f64(lu), _f64(lm)
        phi0 = None if phi0 is None else _f64(phi0)
        perm_ids = None if perm_ids is None else _i32(perm_ids)
        cpg_off, sub_off, sub_p, sub_q, rho_n, d_n = _i64(cpg_off), _i64(sub_off), _i64(sub_p), _i64(sub_q), _f64(rho_n), _f64(d_n)
        R = len(grp_off) - 1
        P_r = np.diff(perm_off)
        T = np.empty(int((3 * P_r * np.diff(sub_off)).sum()))
        status = np.empty(int(2 * perm_off[-1]), dtype=np.int32)
        ctx = self.ctx(device)
        rc = self.dll.cpn_two_samp_null_batch(ctx, C.c_int64(R), _ptr(grp_off), _ptr(Nl), _ptr(rho_l), _ptr(d_l), _ptr(read_off), _ptr(lu), _ptr(lm),
                                              _ptr(perm_off), _ptr(n1), _ptr(perm_ids), _ptr(phi0), C.c_int32(n_init), C.c_int32(max_em_iters),
                                              C.c_double(box[0]), C.c_double(box[1]), C.c_double(box[2]), C.c_int32(mstep_mode), _ptr(cpg_off),
                                              _ptr(rho_n), _ptr(d_n), _ptr(sub_off), _ptr(sub_p), _ptr(sub_q), _ptr(T), _ptr(status))
        self._check(rc, ctx)
        return T, status

    def two_samp_test_batch(self, grp_off, Nl, rho_l, d_l, read_off, lu, lm, n1, phi_s1, phi_s2, perm_off, perm_ids, exact, phi0, n_init,
                            cpg_off, rho_n, d_n, sub_off, sub_p, sub_q, max_em_iters=20, box=(20.0, 150.0, 50.0), mstep_mode=MSTEP_ANALYTIC,
                            want_null=False, device=0):
        """cpn_two_samp_test_batch: the whole two-sample test of R regions.  perm_ids / phi0 may be None (drawn on the device from
        the context's seed).  Returns dict(T, cnt, n_valid, pval [ΣK·3], optionally T_null and fit_status)."""
        grp_off, read_off, perm_off, n1 = _i64(grp_off), _i64(read_off), _i64(perm_off), _i64(n1)
        Nl, rho_l, d_l, lu, lm = _f64(Nl), _f64(rho_l), _f64(d_l), _f64(lu), _f64(lm)
        phi_s1, phi_s2 = _f64(phi_s1), _f64(phi_s2)
        perm_ids = None if perm_ids is None else _i32(perm_ids)
        phi0 = None if phi0 is None else _f64(phi0)
        exact = None if exact is None else _i32(exact)
        cpg_off, sub_off, sub_p, sub_q, rho_n, d_n = _i64(cpg_off), _i64(sub_off), _i64(sub_p), _i64(sub_q), _f64(rho_n), _f64(d_n)
        R, sK = len(grp_off) - 1, int(sub_off[-1])
        out = dict(T=np.empty(3 * sK), cnt=np.empty(3 * sK, dtype=np.int64), n_valid=np.empty(3 * sK, dtype=np.int64), pval=np.empty(3 * sK))
        T_null = status = None
        if want_null:
            T_null = np.empty(int((3 * np.diff(perm_off) * np.diff(sub_off)).sum()))
            status = np.empty(int(2 * perm_off[-1]), dtype=np.int32)
            out.update(T_null=T_null, fit_status=status)
        ctx = self.ctx(device)
        rc = self.dll.cpn_two_samp_test_batch(ctx, C.c_int64(R), _ptr(grp_off), _ptr(Nl), _ptr(rho_l), _ptr(d_l), _ptr(read_off), _ptr(lu), _ptr(lm),
                                              _ptr(n1), _ptr(phi_s1), _ptr(phi_s2), _ptr(perm_off), _ptr(perm_ids), _ptr(exact), _ptr(phi0),
                                              C.c_int32(n_init), C.c_int32(max_em_iters), C.c_double(box[0]), C.c_double(box[1]),
                                              C.c_double(box[2]), C.c_int32(mstep_mode), _ptr(cpg_off), _ptr(rho_n), _ptr(d_n), _ptr(sub_off),
                                              _ptr(sub_p), _ptr(sub_q), _ptr(out["T"]), _ptr(out["cnt"]), _ptr(out["n_valid"]), _ptr(out["pval"]),
                                              _ptr(T_null), _ptr(status))
        self._check(rc, ctx)
        return out

    def two_samp_pvals(self, tbl_off, P, T_obs, T_null, exact, device=0):
        tbl_off = _i64(tbl_off)
        T_obs, T_null = _f64(T_obs), _f64(T_null)
        R, sK = len(tbl_off) - 1, int(tbl_off[-1])
        cnt, nv = np.empty(3 * sK, dtype=np.int64), np.empty(3 * sK, dtype=np.int64)
        p = np.empty(3 * sK)
        ctx = self.ctx(device)
        rc = self.dll.cpn_two_samp_pvals(ctx, C.c_int64(R), _ptr(tbl_off), C.c_int64(P), _ptr(T_obs), _ptr(T_null), C.c_int32(int(exact)), _ptr(cnt),
                                         _ptr(nv), _ptr(p))
        self._check(rc, ctx)
        return cnt, nv, p


def rng_phi0(lib, seed, region, n_init, perm=-1, half=0):
    """The starts the device draws for one EM unit when phi0 is None (cpn_rng_phi0)."""
    out = np.empty((n_init, 3))
    if lib.dll.cpn_rng_phi0(C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.c_int64(region), C.c_int64(perm), C.c_int32(half), C.c_int32(n_init), _ptr(out)) != 0:
        raise CpnError("cpn_rng_phi0: bad arguments")
    return out


def rng_perm(lib, seed, region, perm, n, n1):
    """The ascending 1-based sample-1 indices the device draws for a permutation when perm_ids is None (cpn_rng_perm)."""
    out = np.empty(n1, dtype=np.int32)
    if lib.dll.cpn_rng_perm(C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.c_int64(region), C.c_int64(perm), C.c_int32(n), C.c_int32(n1), _ptr(out)) != 0:
        raise CpnError("cpn_rng_perm: bad arguments")
    return out


def region_filter(lib, n_cpg, grp_off, read_off, log_pyx_u, min_cov, n_threads=0):
    """Status words of pmap_analyze_reg's data filters (cpn_region_filter): 0 = fit the region, else FIT_FILTERED − reasons."""
    n_cpg, grp_off, read_off, lu = _i64(n_cpg), _i64(grp_off), _i64(read_off), _f64(log_pyx_u)
    st = np.empty(len(n_cpg), dtype=np.int32)
    if lib.dll.cpn_region_filter(C.c_int64(len(n_cpg)), _ptr(n_cpg), _ptr(grp_off), _ptr(read_off), _ptr(lu), C.c_double(min_cov),
                                 C.c_int32(n_threads), _ptr(st)) != 0:
        raise CpnError("cpn_region_filter: bad arguments")
    return st


def calls_coverage(lib, grp_off, read_off, log_pyx_u, n_threads=0):
    """(n_obs, n_grp_obs) per region of an emission batch: numerators of get_ave_depth / perc_gprs_obs (cpn_calls_coverage)."""
    grp_off, read_off, lu = _i64(grp_off), _i64(read_off), _f64(log_pyx_u)
    R = len(grp_off) - 1
    a, b = np.empty(R, dtype=np.int64), np.empty(R, dtype=np.int64)
    if lib.dll.cpn_calls_coverage(C.c_int64(R), _ptr(grp_off), _ptr(read_off), _ptr(lu), C.c_int32(n_threads), _ptr(a), _ptr(b)) != 0:
        raise CpnError("cpn_calls_coverage: bad arguments")
    return a, b


class FastaHandle:
    """cpn_fasta: records of a FASTA file (plain / .gz) read by the native loader; sequences are zero-copy uint8 views that keep
    the handle alive."""

    def __init__(self, lib, path):
        d = lib.dll
        d.cpn_fasta_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        d.cpn_fasta_close.argtypes = [C.c_void_p]
        d.cpn_fasta_open_error.restype = C.c_char_p
        for f in ("cpn_fasta_num_records", "cpn_fasta_length"):
            getattr(d, f).restype = C.c_int64
        d.cpn_fasta_num_records.argtypes = [C.c_void_p]
        d.cpn_fasta_name.restype = C.c_char_p
        d.cpn_fasta_seq.restype = C.c_void_p
        d.cpn_fasta_name.argtypes = d.cpn_fasta_length.argtypes = d.cpn_fasta_seq.argtypes = [C.c_void_p, C.c_int64]
        h = C.c_void_p()
        rc = d.cpn_fasta_open(str(path).encode(), C.byref(h))
        if rc != 0 or not h.value:
            raise CpnError(f"cpn_fasta_open failed with code {rc}: {d.cpn_fasta_open_error().decode()}")
        self.lib, self.h = lib, h
        self.records = {}
        for i in range(int(d.cpn_fasta_num_records(h))):
            n = int(d.cpn_fasta_length(h, i))
            ptr = d.cpn_fasta_seq(h, i)
            arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)) if n > 0 else np.zeros(0, dtype=np.uint8)
            arr.flags.writeable = False
            self.records[d.cpn_fasta_name(h, i).decode()] = arr

    def close(self):
        if self.h is not None:
            self.records = {}
            self.lib.dll.cpn_fasta_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- genome features (host only) ------------------------------------------------------------------------------------------
def cpg_scan(lib, seq, n_threads=0):
    """1-based positions of every CG in a uint8 sequence array (cpn_cpg_scan)."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    n = lib.dll.cpn_cpg_scan(_ptr(seq), C.c_int64(len(seq)), C.c_int32(n_threads), None, C.c_int64(0))
    out = np.empty(max(int(n), 0), dtype=np.int64)
    if n > 0:
        lib.dll.cpn_cpg_scan(_ptr(seq), C.c_int64(len(seq)), C.c_int32(n_threads), _ptr(out), C.c_int64(len(out)))
    return out


def chr_part(lib, cg, chr_size, reg_st, reg_end, size_est_reg=3000, min_grp_dist=10):
    cg = _i64(cg)
    a = (_ptr(cg), C.c_int64(len(cg)), C.c_int64(chr_size), C.c_int64(reg_st), C.c_int64(reg_end), C.c_int64(size_est_reg), C.c_int64(min_grp_dist))
    k = int(lib.dll.cpn_chr_part(*a, None, None, C.c_int64(0)))
    st, en = np.empty(k, dtype=np.int64), np.empty(k, dtype=np.int64)
    if k > 0:
        lib.dll.cpn_chr_part(*a, _ptr(st), _ptr(en), C.c_int64(k))
    return st, en


def grp_info_counts(lib, cg, chrst, chrend, min_grp_dist, N, L, n_threads=0):
    """cpn_grp_info_count into caller arrays N, L (int64 [R])."""
    cg, chrst, chrend = _i64(cg), _i64(chrst), _i64(chrend)
    rc = lib.dll.cpn_grp_info_count(_ptr(cg), C.c_int64(len(cg)), C.c_int64(len(chrst)), _ptr(chrst), _ptr(chrend), C.c_int64(min_grp_dist),
                                    C.c_int32(n_threads), _ptr(N), _ptr(L))
    if rc != 0:
        raise CpnError(f"cpn_grp_info_count error {rc}")


def grp_info_batch(lib, cg, chr_size, chrst, chrend, min_grp_dist=10, n_threads=0):
    """get_grp_info! for R regions: dict of CSR arrays (cpg_off, grp_off, cpg_pos, rho_n, d_n, Nl, rho_l, d_l, grp_st, grp_end)."""
    cg, chrst, chrend = _i64(cg), _i64(chrst), _i64(chrend)
    R = len(chrst)
    N, L = np.empty(R, dtype=np.int64), np.empty(R, dtype=np.int64)
    rc = lib.dll.cpn_grp_info_count(_ptr(cg), C.c_int64(len(cg)), C.c_int64(R), _ptr(chrst), _ptr(chrend), C.c_int64(min_grp_dist), C.c_int32(n_threads),
                                    _ptr(N), _ptr(L))
    if rc != 0:
        raise CpnError(f"cpn_grp_info_count error {rc}")
    cpg_off = np.concatenate([[0], np.cumsum(N)]).astype(np.int64); grp_off = np.concatenate([[0], np.cumsum(L)]).astype(np.int64)
    sN, sL = int(cpg_off[-1]), int(grp_off[-1])
    o = dict(cpg_off=cpg_off, grp_off=grp_off, cpg_pos=np.empty(sN, dtype=np.int64), rho_n=np.empty(sN), d_n=np.empty(int(np.maximum(N - 1, 0).sum())),
             Nl=np.empty(sL), rho_l=np.empty(sL), d_l=np.empty(int(np.maximum(L - 1, 0).sum())), grp_st=np.empty(sL, dtype=np.int64),
             grp_end=np.empty(sL, dtype=np.int64))
    rc = lib.dll.cpn_grp_info_fill(_ptr(cg), C.c_int64(len(cg)), C.c_int64(chr_size), C.c_int64(R), _ptr(chrst), _ptr(chrend), C.c_int64(min_grp_dist),
                                   _ptr(cpg_off), _ptr(grp_off), C.c_int32(n_threads), _ptr(o["cpg_pos"]), _ptr(o["rho_n"]), _ptr(o["d_n"]), _ptr(o["Nl"]),
                                   _ptr(o["rho_l"]), _ptr(o["d_l"]), _ptr(o["grp_st"]), _ptr(o["grp_end"]))
    if rc != 0:
        raise CpnError(f"cpn_grp_info_fill error {rc}")
    o["dn_off"] = np.concatenate([[0], np.cumsum(np.maximum(N - 1, 0))]).astype(np.int64)
    o["dl_off"] = np.concatenate([[0], np.cumsum(np.maximum(L - 1, 0))]).astype(np.int64)
    return o


def nls_reg_info_batch(lib, chrst, chrend, max_size_subreg, cpg_off, cpg_pos, n_threads=0):
    """get_nls_reg_info! for R regions → (sub_off [R+1], sub_st, sub_end, sub_p, sub_q)."""
    chrst, chrend, cpg_off, cpg_pos = _i64(chrst), _i64(chrend), _i64(cpg_off), _i64(cpg_pos)
    R = len(chrst)
    K = np.empty(R, dtype=np.int64)
    a = (C.c_int64(R), _ptr(chrst), _ptr(chrend), C.c_int64(max_size_subreg), _ptr(cpg_off), _ptr(cpg_pos), C.c_int32(n_threads))
    if lib.dll.cpn_nls_reg_info_batch(*a, _ptr(K), None, None, None, None, None) != 0:
        raise CpnError("cpn_nls_reg_info_batch: bad arguments")
    sub_off = np.concatenate([[0], np.cumsum(K)]).astype(np.int64)
    st, en, p, q = (np.empty(int(sub_off[-1]), dtype=np.int64) for _ in range(4))
    if lib.dll.cpn_nls_reg_info_batch(*a, None, _ptr(sub_off), _ptr(st), _ptr(en), _ptr(p), _ptr(q)) != 0:
        raise CpnError("cpn_nls_reg_info_batch: bad arguments")
    return sub_off, st, en, p, q


# ---- output text (host only) ------------------------------------------------------------------------------------------------
def format_f64(lib, x, sep=",", digits=-1):
    """Julia's text of the doubles in x, joined by sep (cpn_format_f64)."""
    x = _f64(np.atleast_1d(x))
    n = lib.dll.cpn_format_f64(_ptr(x), C.c_int64(len(x)), C.c_char(sep.encode()), C.c_int32(digits), None, C.c_int64(0))
    buf = C.create_string_buffer(int(n) + 1)
    lib.dll.cpn_format_f64(_ptr(x), C.c_int64(len(x)), C.c_char(sep.encode()), C.c_int32(digits), buf, C.c_int64(n))
    return buf.raw[:n].decode()


def write_rows(lib, path, chrom, a, b, v, digits=-1, clamp0=False, w=None):
    a, b, v = _i64(a), _i64(b), _f64(v)
    w = None if w is None else _f64(w)
    rc = lib.dll.cpn_write_rows(str(path).encode(), chrom.encode(), C.c_int64(len(v)), _ptr(a), _ptr(b), _ptr(v), C.c_int32(digits), C.c_int32(int(clamp0)),
                                _ptr(w))
    if rc != 0:
        raise CpnError(f"cpn_write_rows({path}) failed with code {rc}")


def read_theta(lib, path):
    """{chr: (chrst[n], chrend[n], phi[n, 3])} of a model file, rows in file order (cpn_theta_*)."""
    d = lib.dll
    d.cpn_theta_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    d.cpn_theta_close.argtypes = [C.c_void_p]
    d.cpn_theta_open_error.restype = C.c_char_p
    d.cpn_theta_num_rows.restype = d.cpn_theta_num_chr.restype = C.c_int64
    d.cpn_theta_num_rows.argtypes = d.cpn_theta_num_chr.argtypes = [C.c_void_p]
    d.cpn_theta_chr_name.restype = C.c_char_p
    d.cpn_theta_chr_name.argtypes = [C.c_void_p, C.c_int64]
    d.cpn_theta_copy.argtypes = [C.c_void_p] * 5
    h = C.c_void_p()
    rc = d.cpn_theta_open(str(path).encode(), C.byref(h))
    if rc != 0 or not h.value:
        raise CpnError(f"cpn_theta_open failed with code {rc}: {d.cpn_theta_open_error().decode()}")
    try:
        n = int(d.cpn_theta_num_rows(h))
        names = [d.cpn_theta_chr_name(h, i).decode() for i in range(int(d.cpn_theta_num_chr(h)))]
        cid, st, en, phi = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64), np.empty((n, 3))
        if d.cpn_theta_copy(h, _ptr(cid), _ptr(st), _ptr(en), _ptr(phi)) != 0:
            raise CpnError("cpn_theta_copy failed")
    finally:
        d.cpn_theta_close(h)
    out = {}
    for i, name in enumerate(names):
        m = cid == i
        out[name] = (st[m], en[m], phi[m])
    return out


def add_qvalues(lib, path):
    rc = lib.dll.cpn_add_qvalues(str(path).encode())
    if rc != 0:
        raise CpnError(f"cpn_add_qvalues({path}) failed with code {rc}")


def write_theta(lib, path, chrom, chrst, chrend, phi, grp_off, Nl, rho_l, d_l):
    chrst, chrend, grp_off, phi, rho_l, d_l = _i64(chrst), _i64(chrend), _i64(grp_off), _f64(phi), _f64(rho_l), _f64(d_l)
    Nl = None if Nl is None else _f64(Nl)
    rc = lib.dll.cpn_write_theta(str(path).encode(), chrom.encode(), C.c_int64(len(chrst)), _ptr(chrst), _ptr(chrend), _ptr(phi), _ptr(grp_off), _ptr(Nl),
                                 _ptr(rho_l), _ptr(d_l))
    if rc != 0:
        raise CpnError(f"cpn_write_theta({path}) failed with code {rc}")


class CallsHandle:
    """cpn_calls: a nanopolish calls file parsed once by the native loader (csrc/cpn_calls.cpp).  Host-only — works without a GPU."""

    def __init__(self, lib, path, n_threads=0):
        self.lib = lib
        h = C.c_void_p()
        rc = lib.dll.cpn_calls_open(str(path).encode(), C.c_int32(n_threads), C.byref(h))
        if rc != 0 or not h.value:
            raise CpnError(f"cpn_calls_open failed with code {rc}: {lib.dll.cpn_calls_open_error().decode()}")
        self.h = h
        d = lib.dll
        self.n_lines = int(d.cpn_calls_num_lines(h))
        self.n_reads = int(d.cpn_calls_num_reads(h))
        self.chromosomes = {d.cpn_calls_chr_name(h, C.c_int64(i)).decode(): int(d.cpn_calls_chr_lines(h, C.c_int64(i)))
                            for i in range(int(d.cpn_calls_num_chr(h)))}

    def close(self):
        if self.h is not None:
            self.lib.dll.cpn_calls_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def read_name(self, read_id):
        return self.lib.dll.cpn_calls_read_name(self.h, C.c_int64(int(read_id))).decode()

    def count_reads(self, chrom, chrst, chrend, n_threads=0):
        chrst, chrend = _i64(chrst), _i64(chrend)
        n = np.empty(len(chrst), dtype=np.int64)
        rc = self.lib.dll.cpn_calls_count_reads(self.h, chrom.encode(), C.c_int64(len(chrst)), _ptr(chrst), _ptr(chrend), C.c_int32(n_threads), _ptr(n))
        if rc != 0:
            raise CpnError(f"cpn_calls_count_reads error {rc}: {self.lib.dll.cpn_calls_error(self.h).decode()}")
        return n

    def fill(self, chrom, chrst, chrend, grp_off, grp_st, grp_end, mod_nse=True, read_off=None, n_threads=0, want_ids=False):
        """→ (read_off [R+1], log_pyx_u, log_pyx_m[, read_ids]) in the CSR layout of cpn_fit_batch."""
        chrst, chrend, grp_off, grp_st, grp_end = _i64(chrst), _i64(chrend), _i64(grp_off), _i64(grp_st), _i64(grp_end)
        R = len(chrst)
        if read_off is None:
            read_off = np.concatenate([[0], np.cumsum(self.count_reads(chrom, chrst, chrend, n_threads))]).astype(np.int64)
        read_off = _i64(read_off)
        n_obs = int(np.sum(np.diff(read_off) * np.diff(grp_off))) if R else 0
        lu, lm = np.empty(n_obs), np.empty(n_obs)
        ids = np.empty(int(read_off[-1]) if R else 0, dtype=np.int64)
        rc = self.lib.dll.cpn_calls_fill(self.h, chrom.encode(), C.c_int64(R), _ptr(chrst), _ptr(chrend), _ptr(grp_off), _ptr(grp_st), _ptr(grp_end),
                                         C.c_int32(int(bool(mod_nse))), _ptr(read_off), C.c_int32(n_threads), _ptr(lu), _ptr(lm),
                                         _ptr(ids) if want_ids else None)
        if rc != 0:
            raise CpnError(f"cpn_calls_fill error {rc}: {self.lib.dll.cpn_calls_error(self.h).decode()}")
        return (read_off, lu, lm, ids) if want_ids else (read_off, lu, lm)


class BamHandle:
    """cpn_bam: a BAM file (Bismark / Arioc XM tags) inflated and indexed once by the native reader (csrc/cpn_bam.cpp).
    Host-only — works without a GPU."""

    def __init__(self, lib, path, n_threads=0):
        self.lib = lib
        d = lib.dll
        d.cpn_bam_open.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]
        d.cpn_bam_close.argtypes = [C.c_void_p]
        d.cpn_bam_open_error.restype = d.cpn_bam_error.restype = d.cpn_bam_ref_name.restype = C.c_char_p
        d.cpn_bam_error.argtypes = d.cpn_bam_num_records.argtypes = d.cpn_bam_num_refs.argtypes = [C.c_void_p]
        d.cpn_bam_num_records.restype = d.cpn_bam_num_refs.restype = d.cpn_bam_ref_len.restype = C.c_int64
        d.cpn_bam_ref_name.argtypes = d.cpn_bam_ref_len.argtypes = [C.c_void_p, C.c_int64]
        h = C.c_void_p()
        rc = d.cpn_bam_open(str(path).encode(), C.c_int32(n_threads), C.byref(h))
        if rc != 0 or not h.value:
            raise CpnError(f"cpn_bam_open failed with code {rc}: {d.cpn_bam_open_error().decode()}")
        self.h = h
        self.n_records = int(d.cpn_bam_num_records(h))
        self.refs = {d.cpn_bam_ref_name(h, C.c_int64(i)).decode(): int(d.cpn_bam_ref_len(h, C.c_int64(i))) for i in range(int(d.cpn_bam_num_refs(h)))}
        self.ref_names = list(self.refs)

    def close(self):
        if self.h is not None:
            self.lib.dll.cpn_bam_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def record(self, i):
        """Fields of record i (file order) as the reference's BAM accessors give them: dict(ref, pos, rightpos, flag, mapq, has_xm,
        has_xs, qname, xm)."""
        ref, flag, mapq, hxm, hxs = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        pos, right, ql, xl = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        qn, xm = C.c_void_p(), C.c_void_p()
        rc = self.lib.dll.cpn_bam_record(self.h, C.c_int64(int(i)), C.byref(ref), C.byref(pos), C.byref(right), C.byref(flag), C.byref(mapq),
                                         C.byref(hxm), C.byref(hxs), C.byref(qn), C.byref(ql), C.byref(xm), C.byref(xl))
        if rc != 0:
            raise CpnError(f"cpn_bam_record({i}) failed with code {rc}")
        return dict(ref=ref.value, pos=pos.value, rightpos=right.value, flag=flag.value, mapq=mapq.value, has_xm=bool(hxm.value),
                    has_xs=bool(hxs.value), qname=C.string_at(qn.value, ql.value).decode(),
                    xm=C.string_at(xm.value, xl.value).decode() if hxm.value and xl.value else "")

    def _args(self, chrom, roi_st, roi_end, cpg_off, cpg_pos, pe, trim):
        roi_st, roi_end, cpg_off, cpg_pos = _i64(roi_st), _i64(roi_end), _i64(cpg_off), _i64(cpg_pos)
        trim = _i64(list(trim))
        assert trim.size == 4 and len(cpg_off) == len(roi_st) + 1
        return roi_st, roi_end, cpg_off, cpg_pos, trim

    def count_reads(self, chrom, roi_st, roi_end, cpg_off, cpg_pos, pe=False, trim=(0, 0, 0, 0), n_threads=0):
        roi_st, roi_end, cpg_off, cpg_pos, trim = self._args(chrom, roi_st, roi_end, cpg_off, cpg_pos, pe, trim)
        n = np.zeros(len(roi_st), dtype=np.int64)
        rc = self.lib.dll.cpn_bam_count_reads(self.h, chrom.encode(), C.c_int64(len(roi_st)), _ptr(roi_st), _ptr(roi_end), _ptr(cpg_off), _ptr(cpg_pos),
                                              C.c_int32(int(bool(pe))), _ptr(trim), C.c_int32(n_threads), _ptr(n))
        if rc != 0:
            raise CpnError(f"cpn_bam_count_reads error {rc}: {self.lib.dll.cpn_bam_error(self.h).decode()}")
        return n

    def calls(self, chrom, roi_st, roi_end, cpg_off, cpg_pos, pe=False, trim=(0, 0, 0, 0), n_threads=0):
        """get_calls_bam for every region of a chromosome → (read_off [R+1], calls int8 flat, read-major [m_r × N_r] per region)."""
        roi_st, roi_end, cpg_off, cpg_pos, trim = self._args(chrom, roi_st, roi_end, cpg_off, cpg_pos, pe, trim)
        n = self.count_reads(chrom, roi_st, roi_end, cpg_off, cpg_pos, pe, trim, n_threads)
        read_off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
        cell_off = np.concatenate([[0], np.cumsum(n * np.diff(cpg_off))]).astype(np.int64)
        out = np.zeros(int(cell_off[-1]), dtype=np.int8)
        if len(roi_st):
            rc = self.lib.dll.cpn_bam_fill_calls(self.h, chrom.encode(), C.c_int64(len(roi_st)), _ptr(roi_st), _ptr(roi_end), _ptr(cpg_off), _ptr(cpg_pos),
                                                 C.c_int32(int(bool(pe))), _ptr(trim), C.c_int32(n_threads), _ptr(cell_off), _ptr(out))
            if rc != 0:
                raise CpnError(f"cpn_bam_fill_calls error {rc}: {self.lib.dll.cpn_bam_error(self.h).decode()}")
        return read_off, out
