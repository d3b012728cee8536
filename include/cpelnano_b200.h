/* cpelnano_b200.h — C ABI of libcpelnano_b200.so (hand-written sm_100a CUDA kernels).
 *
 * Drop-in boundary for the per-region Ising estimation hot path of jordiabante/CpelNano.jl.
 * The reference has no FFI; the boundary is cut at the Julia functions listed below, and each entry
 * point here is what a `ccall` shim for that function binds (julia/CpelNanoB200.jl).
 * Plain C: pointers and sizes only.  All host buffers are caller-owned; the library never keeps a host
 * pointer after a call returns.  Return value 0 = ok, non-zero = error (text via cpn_last_error).
 * Per-region numerical failures are NOT errors: they come back as status codes / NaN, mirroring the
 * reference's "log and skip" behaviour (src/em_algorithm.jl:160-172,385).
 *
 * Conventions
 *   - state index: x=-1 "unmethylated", x=+1 "methylated"
 *   - CSR batches: `grp_off[R+1]` indexes per-group arrays (Nl, rho_l); d_l (L_r-1 entries per region)
 *     starts at grp_off[r]-r.  `cpg_off[R+1]` / d_n likewise at CpG scale.
 *   - emissions are read-major per region: element (m,l) of region r at obs_off[r] + m*L_r + l with
 *     obs_off[r] = sum_{r'<r} M_r' L_r'; NaN in log_pyx_u marks "group not observed in this read"
 *     (MethCallCpgGrp.obs == false, src/structures.jl:121-128).
 *   - there is NO CPU fallback: every compute entry point fails with CPN_ERR_CUDA if no sm_100 device.
 */
#ifndef CPELNANO_B200_H
#define CPELNANO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cpn_ctx cpn_ctx;

enum {
    CPN_OK = 0,
    CPN_ERR_ARG = 1,      /* bad argument (null pointer, negative size, L<2, ...) */
    CPN_ERR_CUDA = 2,     /* CUDA runtime error or no usable device */
    CPN_ERR_NOMEM = 3
};

/* M-step evaluation mode (src/em_algorithm.jl:152-163 solves ∇logZ(ϕ) = ES/M with NLsolve):
 *   CPN_MSTEP_ANALYTIC  gradient and Jacobian of log Z by forward-mode recursion along the chain (default)
 *   CPN_MSTEP_FD        the reference's central finite differences of log Z (Calculus.gradient inside a
 *                       FiniteDiff central Jacobian), same step sizes and evaluation order            */
enum { CPN_MSTEP_FD = 0, CPN_MSTEP_ANALYTIC = 1 };

/* Per-region status word of the EM fit (`status` outputs; SURVEY.md §5 failure handling).  The reference "logs and skips": a
 * region without a fit gets rs.ϕhat = [] (src/em_algorithm.jl:385), rs.proc stays false and the writers skip it.
 *   >= 0                   index (0-based) of the picked start: ok
 *   CPN_FIT_NONE           every start ended with Q = -Inf (stopped in its first iteration or hit max_em_iters, :341-344)
 *   CPN_FIT_NUMERICAL      as CPN_FIT_NONE, and at least one M-step of the region met a non-finite residual or iterate (the
 *                          exception the reference catches and logs, :160-172)
 *   <= CPN_FIT_FILTERED    the region never reached the EM: it failed a data filter of pmap_analyze_reg (src/nanopore.jl:309,
 *                          322-326).  Written by cpn_region_filter as CPN_FIT_FILTERED - (OR of the CPN_FILTER_* reasons).      */
enum { CPN_FIT_NONE = -1, CPN_FIT_NUMERICAL = -2, CPN_FIT_FILTERED = -16 };
enum { CPN_FILTER_FEW_CPGS = 1 /* N <= 9, :309 */, CPN_FILTER_NO_READS = 2 /* rs.m == 0 */, CPN_FILTER_LOW_DEPTH = 4 /* get_ave_depth <
       min_cov, :322 */, CPN_FILTER_FEW_GROUPS_OBS = 8 /* perc_gprs_obs < 0.66, :323 */, CPN_FILTER_ONE_GROUP = 16 /* L <= 1, :326 */ };

/* ---- context ------------------------------------------------------------------------------- */
int  cpn_create(int device_id, cpn_ctx** out);
/* One context spanning several GPUs of the box, for a single host process (the reference is one master process,
 * src/nanopore.jl:362-416).  The host-pointer entry points cpn_fit_batch, cpn_analyze_batch, cpn_grp_test_batch and
 * cpn_two_samp_test_batch then shard the regions of a call over the devices — contiguous blocks balanced by the cost proxy
 * reads x groups, one host thread per device, each device copying its results straight into the caller's arrays at
 * the shard's offset: no collective, outputs in input order, bit-identical to the single-device call.  Every other entry point
 * runs on device_ids[0].  cpn_last_device_ms = the slowest device's time.                                                       */
int  cpn_create_multi(const int* device_ids, int n_dev, cpn_ctx** out);
int  cpn_num_devices(const cpn_ctx* ctx);
void cpn_destroy(cpn_ctx* ctx);
const char* cpn_last_error(const cpn_ctx* ctx);           /* NUL-terminated, library-owned */
const char* cpn_version(void);
/* CUDA-event time (ms) of the device work of the last call on this ctx, and the number of kernels
 * it launched (for bench.py's gpu_launches). */
double  cpn_last_device_ms(const cpn_ctx* ctx);
int64_t cpn_last_launches(const cpn_ctx* ctx);
/* Algorithmic work of the EM fits of the last call, for roofline accounting (bench.py): work[0] = sum over EM instances of
 * E-step passes x reads x groups, work[1] = sum of derivative sweeps (analytic mode) or log Z evaluations (FD mode) x groups,
 * work[2] = reads x groups of the final log p(y) pass over converged instances, work[3] = number of EM instances.            */
int cpn_last_work(const cpn_ctx* ctx, double* work /*[4]*/);
/* Measured FP64 FMA peak of this device (dependent-chain-free DFMA loop), TFLOP/s; and a DRAM copy
 * bandwidth figure GB/s.  Used as roofline denominators when MEASURED_PEAKS.json has no FP64 entry. */
int cpn_measure_fp64_peak(cpn_ctx* ctx, double* tflops);
/* Seed of the counter-based stream (Philox-4x32-10, csrc/cpn_philox.h) that replaces the host RNG wherever an entry point is
 * given phi0 == NULL (EM starts, gen_ϕ0s src/em_algorithm.jl:65-78) or perm_ids == NULL (read-label permutations,
 * src/two_samp_perm_tests.jl:211-214).  Starts are keyed by (seed, position of the region in the call's batch, start), so a
 * result does not depend on how the library chunks or shards the batch.  Default seed 0.                                     */
int cpn_set_seed(cpn_ctx* ctx, uint64_t seed);
/* The same stream evaluated on the host (no GPU needed), for callers that want the arrays the device would draw:
 *   cpn_rng_phi0   the n_init starts [n_init*3] of one EM unit: region = position in the batch; perm < 0 for the plain fit, else the
 *                  permutation index with half = 0 / 1 for the two refits of that permutation
 *   cpn_rng_perm   the n1 ascending 1-based indices of sample 1 for permutation `perm` of `region` with n pooled reads           */
int cpn_rng_phi0(uint64_t seed, int64_t region, int64_t perm, int32_t half, int32_t n_init, double* phi0);
int cpn_rng_perm(uint64_t seed, int64_t region, int64_t perm, int32_t n, int32_t n1, int32_t* ids);
/* Optional per-launch profiler: when enabled, every kernel launch is bracketed by CUDA events on the
 * launching stream and its time accumulated per kernel class (8 classes: 0 prep, 1 E-step, 2 M-step,
 * 3 score, 4 pick, 5 summaries, 6 CMD, 7 sweeps).  cpn_set_profile resets the accumulators.          */
int cpn_set_profile(cpn_ctx* ctx, int enable);
int cpn_get_profile(const cpn_ctx* ctx, double* ms /*[8]*/, int64_t* launches /*[8]*/);

/* ---- (b)+(c) multi-start EM: get_ϕhat!(rs, config)  src/em_algorithm.jl:371-390 -------------------
 * For every region r: run n_init EM instances (em_inst, :304-350) from the given starts phi0 (the host
 * RNG owns gen_ϕ0s, :65-78), keep the first start with maximal Q = mean_m log p(y_m; ϕ) (pick_ϕ, :35-53).
 * Inputs are HOST pointers (copied to the device inside the call).
 *   phi0        [R * n_init * 3], or NULL: drawn on the device from the context's seed (cpn_set_seed)
 *   phi_hat     [R * 3]   NaN when no start converged
 *   Q           [R]       -Inf when no start converged
 *   status      [R]       index of the picked start (0-based), or CPN_FIT_NONE / CPN_FIT_NUMERICAL (status word above)
 *   counters    [R * 4]   (em_iters, solver_iters, logZ_evals, estep_passes) summed over starts; may be NULL
 *   phi_starts  [R * n_init * 3], Q_starts [R * n_init]  per-start results; may be NULL              */
int cpn_fit_batch(cpn_ctx* ctx, int64_t R,
                  const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                  const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m,
                  const double* phi0, int32_t n_init, int32_t max_em_iters,
                  double amax, double bmax, double cmax, int32_t mstep_mode,
                  double* phi_hat, double* Q, int32_t* status, int64_t* counters,
                  double* phi_starts, double* Q_starts);

/* Same computation with every buffer already resident in device memory (HBM) of ctx's device; used
 * to measure device-only throughput and by callers that keep packed batches on the GPU.            */
int cpn_fit_batch_device(cpn_ctx* ctx, int64_t R,
                  const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                  const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m,
                  const double* phi0, int32_t n_init, int32_t max_em_iters,
                  double amax, double bmax, double cmax, int32_t mstep_mode,
                  double* phi_hat, double* Q, int32_t* status, int64_t* counters,
                  double* phi_starts, double* Q_starts);

/* ---- batched pmap_analyze_reg: fit → rscle_grp_mod! → get_stat_sums!  src/nanopore.jl:331-343 -------------
 * One call per chromosome replaces the `pmap` at src/nanopore.jl:406.  Inputs of cpn_fit_batch plus the per-CpG
 * features and sub-regions of cpn_stat_sums_batch (model r ↔ region r).  ϕ̂ goes from the EM kernels to the
 * summaries kernel in device memory.  Regions without a converged start get NaN summaries (host: proc=false).
 * The _device variant takes device pointers for every buffer.                                              */
int cpn_analyze_batch(cpn_ctx* ctx, int64_t R,
                  const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                  const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m,
                  const double* phi0, int32_t n_init, int32_t max_em_iters,
                  double amax, double bmax, double cmax, int32_t mstep_mode,
                  const int64_t* cpg_off, const double* rho_n, const double* d_n,
                  const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                  double* phi_hat, double* Q, int32_t* status, int64_t* counters,
                  double* logZ, double* ex, double* exx, double* log_g,
                  double* e_log_g1, double* e_log_g2, double* mml, double* nme);
int cpn_analyze_batch_device(cpn_ctx* ctx, int64_t R,
                  const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                  const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m,
                  const double* phi0, int32_t n_init, int32_t max_em_iters,
                  double amax, double bmax, double cmax, int32_t mstep_mode,
                  const int64_t* cpg_off, const double* rho_n, const double* d_n,
                  const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                  double* phi_hat, double* Q, int32_t* status, int64_t* counters,
                  double* logZ, double* ex, double* exx, double* log_g,
                  double* e_log_g1, double* e_log_g2, double* mml, double* nme);

/* ---- packed emissions -------------------------------------------------------------------------------------------
 * The emissions of a read enter the model only through the log-likelihood ratio of every observed group and through the
 * read's Σ_obs log p(y|−1) (transfer_matrix.jl:632-643, :843-870: γ_l(x) and log p(y_m) − SURVEY.md A.3), so the two
 * tables of structures.jl:121-128 can be handed over as
 *   llr        [Σ_r M_r·L_r]  log_pyx_m − log_pyx_u per (read, group), NaN = unobserved (same order as log_pyx_*)
 *   read_base  [Σ_r M_r]      Σ over the observed groups of the read of log_pyx_u, added left to right
 * — half the bytes on PCIe.  cpn_pack_emissions (host, multi-threaded, no CUDA) computes both from the two tables with the
 * arithmetic of the device's own packing kernel, so the *_packed entry points return bit-identical results to their
 * two-table twins (as long as no |LLR| exceeds the clamp at 300).  All other arguments as in cpn_fit_batch /
 * cpn_analyze_batch (host pointers).                                                                                   */
int cpn_pack_emissions(int64_t R, const int64_t* grp_off, const int64_t* read_off, const double* log_pyx_u,
                       const double* log_pyx_m, double* llr, double* read_base, int32_t n_threads);
int cpn_fit_batch_packed(cpn_ctx* ctx, int64_t R,
                  const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                  const int64_t* read_off, const double* llr, const double* read_base,
                  const double* phi0, int32_t n_init, int32_t max_em_iters,
                  double amax, double bmax, double cmax, int32_t mstep_mode,
                  double* phi_hat, double* Q, int32_t* status, int64_t* counters,
                  double* phi_starts, double* Q_starts);
int cpn_analyze_batch_packed(cpn_ctx* ctx, int64_t R,
                  const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                  const int64_t* read_off, const double* llr, const double* read_base,
                  const double* phi0, int32_t n_init, int32_t max_em_iters,
                  double amax, double bmax, double cmax, int32_t mstep_mode,
                  const int64_t* cpg_off, const double* rho_n, const double* d_n,
                  const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                  double* phi_hat, double* Q, int32_t* status, int64_t* counters,
                  double* logZ, double* ex, double* exx, double* log_g,
                  double* e_log_g1, double* e_log_g2, double* mml, double* nme);

/* ---- WGBS front-end: pmap_analyze_reg_bam  src/wgbs.jl:335-389, get_ϕhat_wgbs!  src/em_algorithm.jl:556-575 ---------------
 * Input = what get_calls_bam (wgbs.jl:222-323; native reader: cpn_bam_* below) produces: per read and
 * CpG site a call x ∈ {-1 unmethylated, +1 methylated, 0 not covered}, int8, [Σ_r M_r·N_r] read-major per region.  The calls
 * are hard emissions (log p(y|x) ∈ {0, -50}, CpelNano.jl:32-33), the regions are modelled at CpG resolution (every CpG its own
 * group), and em_iter_wgbs (em_algorithm.jl:412-470) is em_iter on those emissions: the call packs them (cpn_wgbs_pack, one byte
 * per cell in, host) and runs the same kernels as cpn_analyze_batch_packed.  Outputs as cpn_analyze_batch.                    */
int cpn_wgbs_pack(int64_t R, const int64_t* cpg_off, const int64_t* read_off, const int8_t* calls,
                  double* llr, double* read_base, int32_t n_threads);
int cpn_analyze_wgbs_batch(cpn_ctx* ctx, int64_t R,
                  const int64_t* cpg_off, const double* rho_n, const double* d_n,
                  const int64_t* read_off, const int8_t* calls,
                  const double* phi0, int32_t n_init, int32_t max_em_iters,
                  double amax, double bmax, double cmax, int32_t mstep_mode,
                  const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                  double* phi_hat, double* Q, int32_t* status, int64_t* counters,
                  double* logZ, double* ex, double* exx, double* log_g,
                  double* e_log_g1, double* e_log_g2, double* mml, double* nme);

/* Building blocks of em_iter exposed for parity tests (host pointers):
 *   cpn_estep_batch   ES[r] = get_ess(map(get_cond_exs_log))  src/em_algorithm.jl:142-146 at phi[r]
 *                     and ave_logpy[r] = comp_ave_log_py (:263-288) at the same phi (may be NULL)
 *   cpn_mstep_batch   one NLsolve-equivalent solve from phi_init[r] given ES[r] (:152-178, before clamp);
 *                     converged[r] = converged(sol)                                                  */
int cpn_estep_batch(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l,
                    const double* d_l, const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m,
                    const double* phi /*[R*3]*/, double* ES /*[R*3]*/, double* ave_logpy /*[R] or NULL*/);
/*   cpn_estep_batch_starts   the same statistics for n_init parameter vectors per region (phi [R*n_init*3], start-major within
 *                     a region), evaluated the way an EM iteration of cpn_fit_batch evaluates them: the starts of a region are
 *                     paired (two per warp, sharing the region's emission rows) — get_ϕhat!'s inner E-step over all starts,
 *                     src/em_algorithm.jl:142-146 inside :374-383                                                   */
int cpn_estep_batch_starts(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l,
                           const double* d_l, const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m,
                           int32_t n_init, const double* phi /*[R*n_init*3]*/, double* ES /*[R*n_init*3]*/);
int cpn_mstep_batch(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l,
                    const double* d_l, const int64_t* n_reads /*[R]*/, const double* ES /*[R*3]*/,
                    const double* phi_init /*[R*3]*/, int32_t mstep_mode, double* sol /*[R*3]*/,
                    int32_t* converged /*[R]*/, int64_t* counters /*[R*2] solver_iters, evals; or NULL*/);

/* ---- (a)+(d) summaries of fitted per-CpG chains: get_stat_sums!(rs, config) -------------------------
 * src/statistical_summaries.jl:470-500 after rscle_grp_mod! (src/nanopore.jl:269-280): N_l=1, ρ_l=ρ_n,
 * d_l=d_n.  Sub-regions come from cpn_nls_reg_info (1-based CpG index ranges, 0 = empty).
 * A table of T fitted models is summarised in one call; model t lives on genomic region reg_of[t] (several
 * samples may share a region's features).  reg_of == NULL means T == R and model r ↔ region r, in which case
 * ex_off/k_off may be NULL too (they default to cpg_off/sub_off).
 *   phi [T*3]; reg_of [T]; ex_off [T+1] offsets of each model's E[X] block (E[XX] block at ex_off[t]-t);
 *   k_off [T+1] offsets of each model's per-sub-region outputs;
 *   cpg_off [R+1]; rho_n [ΣN]; d_n [ΣN - R]; sub_off [R+1]; sub_p, sub_q [ΣK]
 *   logZ [T]; ex [ex_off[T]]; exx [ex_off[T]-T]; log_g [k_off[T]*4] (pm,pp,qm,qp); e_log_g1,e_log_g2,mml,nme [k_off[T]] */
int cpn_stat_sums_batch(cpn_ctx* ctx, int64_t T, const double* phi, const int64_t* reg_of, const int64_t* ex_off, const int64_t* k_off,
                        int64_t R, const int64_t* cpg_off, const double* rho_n, const double* d_n,
                        const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                        double* logZ, double* ex, double* exx, double* log_g,
                        double* e_log_g1, double* e_log_g2, double* mml, double* nme);

/* get_nls_reg_info!(rs, config)  src/genome_analysis.jl:472-538 — integer, host-only, bit-exact.
 * Call with sub_st == NULL to obtain K only.  Returns K (number of analysis sub-regions) or -1.      */
int64_t cpn_nls_reg_info(int64_t chrst, int64_t chrend, int64_t max_size_subreg, const int64_t* cpg_pos, int64_t N,
                         int64_t* sub_st, int64_t* sub_end, int64_t* sub_p, int64_t* sub_q);

/* comp_cmd(rs1, rs2)  src/statistical_summaries.jl:512-570 for P pairs of fitted models.
 * Models live in a table of T models (phi_tbl [T*3]; ex_tbl/exx_tbl/nme_tbl are the outputs of
 * cpn_stat_sums_batch for those models, laid out with the per-model offsets below); each model t belongs to
 * genomic region reg_of[t] (features rho_n/d_n/sub_* indexed by region as in cpn_stat_sums_batch).
 *   pair_a, pair_b [P] model indices (both must share the same region); cmd [Σ_pairs K_region] at cmd_off[P+1] */
int cpn_cmd_batch(cpn_ctx* ctx, int64_t P, const int64_t* pair_a, const int64_t* pair_b,
                  int64_t T, const double* phi_tbl, const int64_t* reg_of,
                  const int64_t* ex_off /*[T+1] into ex_tbl; exx at ex_off[t]-t*/, const double* ex_tbl, const double* exx_tbl,
                  const int64_t* nme_off /*[T+1] into nme_tbl*/, const double* nme_tbl,
                  int64_t R, const int64_t* cpg_off, const double* rho_n, const double* d_n,
                  const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q,
                  const int64_t* cmd_off /*[P+1]*/, double* cmd);

/* ---- (e) permutation sweeps ---------------------------------------------------------------------------
 * unmat_est_reg_test / mat_est_reg_test  src/grp_perm_tests.jl:200-291 / :432-499 for R regions that all
 * have n1+n2 (unmatched) or n pairs (matched) fitted samples.  Statistic tables are inputs so that the
 * integer counts are a pure function of them (accumulation order = the reference's, no FMA contraction).
 *   tbl_off [R+1] offsets (in sub-regions) into the tables; K_r = tbl_off[r+1]-tbl_off[r]
 *   mml,nme [Σ_r S*K_r] sample-major per region (group 1 samples first); cmd_tbl [Σ_r S*S*K_r], entry
 *   (i*S+j)*K_r+k valid for i<j;  labelings [P*n1] ascending 1-based ids shared by all regions.
 *   Outputs per region and sub-region: T [ΣK*3] (tmml,tnme,tcmd), cnt [ΣK*3], pval [ΣK*3] (NaN if P==0)  */
int cpn_grp_unmat_sweep(cpn_ctx* ctx, int64_t R, int32_t n1, int32_t n2, const int64_t* tbl_off,
                        const double* mml, const double* nme, const double* cmd_tbl,
                        const int32_t* labelings, int64_t P, int32_t exact,
                        double* T, int64_t* cnt, double* pval);
/* matched: mml1,mml2,nme1,nme2 [Σ_r n*K_r]; js [P] sign words (bit s set = keep sign of pair s);
 * outputs T [ΣK*2] (tmml,tnme), cnt [ΣK*2], pval [ΣK*2].                                                 */
int cpn_grp_mat_sweep(cpn_ctx* ctx, int64_t R, int32_t n, const int64_t* tbl_off,
                      const double* mml1, const double* mml2, const double* nme1, const double* nme2,
                      const int64_t* js, int64_t P, int32_t exact,
                      double* T, int64_t* cnt, double* pval);
/* group comparison, all on the device  src/grp_perm_tests.jl:200-291 (unmatched) / :432-499 (matched) over the pmap of
 * :674: from the fitted parameters of the S = n1+n2 samples of every region to (T, counts, p-values) in one call —
 * get_stat_sums! of the R*S models, comp_cmd of every pair (all i<j unmatched; (s, n1+s) matched) and the labeling sweep.
 * Only phi goes in and only the statistics come out; E[X], E[XX], MML, NME and the CMD table never leave the GPU.
 *   phi [R][S][3]      group 1 first; every sample must have a model (the caller drops regions that do not, :646-655)
 *   cpg_off ... sub_q  per-CpG tables as in cpn_stat_sums_batch
 *   matched == 0:      labelings int32 [P][n1] ascending 1-based ids of group 1 (combinations); outputs T, cnt, pval
 *                      [sumK*3] as cpn_grp_unmat_sweep; tcmd unused
 *   matched != 0:      n1 == n2; labelings int64 js[P] sign words; outputs T, cnt, pval [sumK*2] as cpn_grp_mat_sweep and
 *                      tcmd [sumK] = sum of the pair CMDs / n (:308-314)                                               */
int cpn_grp_test_batch(cpn_ctx* ctx, int64_t R, int32_t n1, int32_t n2, int32_t matched, const double* phi,
                       const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off,
                       const int64_t* sub_p, const int64_t* sub_q, const void* labelings, int64_t P, int32_t exact,
                       double* T, int64_t* cnt, double* pval, double* tcmd);
/* The same with phi, rho_n, d_n, sub_p, sub_q and the outputs resident in device memory; cpg_off, sub_off and the labelings stay
 * host arrays (they size the launches).  Both variants process the regions in chunks of CPN_GRP_CHUNK_REGIONS (default 65536)
 * regions, so the workspace (about 25 KB per region at 6 vs 6) does not grow with the batch.                                     */
int cpn_grp_test_batch_device(cpn_ctx* ctx, int64_t R, int32_t n1, int32_t n2, int32_t matched, const double* phi,
                              const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off,
                              const int64_t* sub_p, const int64_t* sub_q, const void* labelings, int64_t P, int32_t exact,
                              double* T, int64_t* cnt, double* pval, double* tcmd);
/* two-sample null statistics, all on the device  src/two_samp_perm_tests.jl:41-87 (pmap_two_samp_null_stats) over the
 * pmap of :220: for every region r and every permutation i of its pooled reads, refit both halves (multi-start EM),
 * summarise both fits (rscle_grp_mod! + get_stat_sums!) and return Tmml = mml1-mml2, Tnme = nme1-nme2, Tcmd = comp_cmd.
 * The pooled reads cross PCIe once: the 2*P_r refits of a region are VIEWS of its packed emissions (a read-index row per
 * lane), so nothing is materialised per permutation.
 *   grp_off ... log_pyx_m  as in cpn_fit_batch; the reads of region r are vcat(calls_s1, calls_s2) (:54)
 *   perm_off [R+1]         permutations per region;  n1 [R] reads of sample 1
 *   perm_ids               for region r, P_r rows of n1[r] ascending 1-based indices into the pooled reads = sample 1 (:57);
 *                          sample 2 is the rest, in order (deleteat!, :58-59); or NULL: drawn on the device (cpn_set_seed)
 *   phi0 [sum P_r*2*n_init*3]  EM starts, ordered region, permutation, half, start; or NULL: drawn on the device (cpn_set_seed)
 *   cpg_off ... sub_q      per-CpG tables as in cpn_stat_sums_batch
 *   T_null [sum_r P_r*3*K_r]   region-major, then permutation, statistic (mml, nme, cmd), sub-region: the layout
 *                          cpn_two_samp_pvals reads; NaN for a permutation when either refit has no optimum (:65)
 *   fit_status [sum P_r*2] optional: picked start per refit, -1 = none                                              */
int cpn_two_samp_null_batch(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l,
                            const double* d_l, const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m,
                            const int64_t* perm_off, const int64_t* n1, const int32_t* perm_ids, const double* phi0,
                            int32_t n_init, int32_t max_em_iters, double amax, double bmax, double cmax, int32_t mstep_mode,
                            const int64_t* cpg_off, const double* rho_n, const double* d_n, const int64_t* sub_off,
                            const int64_t* sub_p, const int64_t* sub_q, double* T_null, int32_t* fit_status);
/* The whole two-sample test of a batch of regions in one call  src/two_samp_perm_tests.jl:157-273 (pmap_diff_two_samp_comp) over the
 * serial map of :445: observed statistics from the two fitted models (:165-167: get_stat_sums! of both, comp_cmd), null statistics
 * from the refits of every permutation (cpn_two_samp_null_batch's computation), NaN filter, ">20 valid" rule and counting (:230-258).
 * Only the pooled reads and the two models' ϕ̂ go in and only (T, cnt, n_valid, p) come out: T_null (216 KB per region at 1000
 * permutations) never crosses PCIe unless asked for.  Regions are processed in chunks of bounded workspace
 * (CPN_TWO_SAMP_MAX_VIEWS refits per chunk, default 2^20), so the batch may be a whole chromosome.
 *   phi_s1, phi_s2 [R*3]   ϕ̂ of the two samples (from their model files)
 *   perm_ids               as in cpn_two_samp_null_batch, or NULL: perm_off[r+1]-perm_off[r] index sets per region are drawn on the
 *                          device from the context's seed (ascending, uniform over the C(n, n1) subsets: the law of
 *                          sort!(sample(1:n, n1; replace=false)), :211-214)
 *   exact [R] or NULL      1: p = cnt/n_valid (all C(n,n1) splits were enumerated, :254), 0: p = (1+cnt)/(1+n_valid) (:256)
 *   phi0                   as in cpn_two_samp_null_batch, or NULL: drawn on the device from the context's seed
 *   T_obs, cnt, n_valid, pval [sumK*3]  region-major, (tmml, tnme, tcmd) blocks of K_r; p is NaN for empty sub-regions and where
 *                          at most 20 permutations gave a finite statistic
 *   T_null, fit_status     optional (NULL: not returned), layouts of cpn_two_samp_null_batch                                         */
int cpn_two_samp_test_batch(cpn_ctx* ctx, int64_t R, const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l,
                            const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m, const int64_t* n1,
                            const double* phi_s1, const double* phi_s2, const int64_t* perm_off, const int32_t* perm_ids,
                            const int32_t* exact, const double* phi0, int32_t n_init, int32_t max_em_iters, double amax, double bmax,
                            double cmax, int32_t mstep_mode, const int64_t* cpg_off, const double* rho_n, const double* d_n,
                            const int64_t* sub_off, const int64_t* sub_p, const int64_t* sub_q, double* T_obs, int64_t* cnt,
                            int64_t* n_valid, double* pval, double* T_null, int32_t* fit_status);
/* two-sample p-value counting  src/two_samp_perm_tests.jl:230-258: T_obs [ΣK*3] region-major (tmml,tnme,tcmd
 * blocks of K_r), T_null [Σ_r P*3*K_r] permutation-major per region; NaN nulls are dropped, >20 valid needed. */
int cpn_two_samp_pvals(cpn_ctx* ctx, int64_t R, const int64_t* tbl_off, int64_t P, const double* T_obs,
                       const double* T_null, int32_t exact, int64_t* cnt, int64_t* n_valid, double* pval);

/* ---- calls ingestion (host only, no GPU needed) ---------------------------------------------------------------------
 * Replaces get_dic_nanopolish (src/input_output.jl:536-570: the TSV is re-read from the top for every region) and
 * get_calls_nanopolish! (src/nanopore.jl:172-230: linear findfirst over the CpG groups per call) — i.e. get_calls!
 * (src/nanopore.jl:241-257) for ALL regions of a chromosome in two calls.  cpn_calls_open parses a nanopolish
 * call-methylation TSV (plain or .gz; first line = header; columns 1 chr, 3 start, 4 end, 5 read name, 7 log_lik_methylated,
 * 8 log_lik_unmethylated) once, with n_threads threads (0 = all), into columns per chromosome.  Decimal text is converted
 * correctly rounded, as parse(Float64, ·).  The handle is read-only afterwards and may be shared by threads.
 *   cpn_calls_count_reads  n_reads[r] = reads with at least one line on `chr` with start <= chrend[r] and end >= chrst[r]
 *                          (every such read becomes a row, rs.m, even if none of its lines matches a group)
 *   cpn_calls_fill         read_off [R+1] = exclusive cumulative sum of n_reads; grp_off/grp_st/grp_end: CSR of the regions'
 *                          CpG-group intervals (grp_int = [first CpG - 5, last CpG + 5], src/genome_analysis.jl:63-111);
 *                          writes log_pyx_u / log_pyx_m in the layout cpn_fit_batch takes (read-major [m_r x L_r] per region,
 *                          NaN = not observed); a line's group is the one with grp_st == start+1-5 and grp_end == end+1+5;
 *                          later lines overwrite earlier ones; mod_nse == 0 stores (0, -50) for the likelier state
 *                          (src/CpelNano.jl:32-33).  Rows are in order of first appearance in the file (the reference: Julia
 *                          Dict order).  read_ids (optional) [sum m_r]: index of each row's read for cpn_calls_read_name.
 * Errors: cpn_calls_open returns non-zero and leaves the text in cpn_calls_open_error() (thread-local).               */
typedef struct cpn_calls cpn_calls;
int  cpn_calls_open(const char* path, int32_t n_threads, cpn_calls** out);
void cpn_calls_close(cpn_calls* calls);
const char* cpn_calls_open_error(void);
const char* cpn_calls_error(const cpn_calls* calls);
int64_t cpn_calls_num_lines(const cpn_calls* calls);
int64_t cpn_calls_num_reads(const cpn_calls* calls);
int64_t cpn_calls_num_chr(const cpn_calls* calls);
const char* cpn_calls_chr_name(const cpn_calls* calls, int64_t i);
int64_t cpn_calls_chr_lines(const cpn_calls* calls, int64_t i);
const char* cpn_calls_read_name(const cpn_calls* calls, int64_t read_id);
int cpn_calls_count_reads(const cpn_calls* calls, const char* chr, int64_t R, const int64_t* chrst, const int64_t* chrend,
                          int32_t n_threads, int64_t* n_reads);
/*   cpn_calls_coverage    numerators of get_ave_depth / perc_gprs_obs (src/structures.jl:200-203) for every region of an emission batch:
 *                          n_obs[r] observed (read, group) cells, n_grp_obs[r] groups observed in at least one read               */
int cpn_calls_coverage(int64_t R, const int64_t* grp_off, const int64_t* read_off, const double* log_pyx_u, int32_t n_threads,
                       int64_t* n_obs, int64_t* n_grp_obs);
/*   cpn_region_filter     the data filters of pmap_analyze_reg (src/nanopore.jl:309, 322-326) as status words: 0 = the region goes on to
 *                          the EM, else CPN_FIT_FILTERED - (OR of CPN_FILTER_* reasons); n_cpg[r] = CpG sites of region r                 */
int cpn_region_filter(int64_t R, const int64_t* n_cpg, const int64_t* grp_off, const int64_t* read_off, const double* log_pyx_u,
                      double min_cov, int32_t n_threads, int32_t* status);
int cpn_calls_fill(const cpn_calls* calls, const char* chr, int64_t R, const int64_t* chrst, const int64_t* chrend,
                   const int64_t* grp_off, const int64_t* grp_st, const int64_t* grp_end, int32_t mod_nse,
                   const int64_t* read_off, int32_t n_threads, double* log_pyx_u, double* log_pyx_m, int64_t* read_ids);

/* ---- genome features (host only, no GPU needed) -----------------------------------------------------------------------
 * get_grp_info! (src/genome_analysis.jl:226-277: regex scan of the region, get_grps_from_cgs :63-111, get_ρn :122-140 — a regex
 * scan of a <= 1 kb window per CpG —, get_ρl/get_dn/get_dl :150-216) and get_chr_part (:327-458) for ALL regions of a chromosome,
 * from one index of the chromosome's CG dinucleotides.
 *   cpn_cpg_scan        1-based positions of the C of every CG (case-insensitive) in seq[0..n); returns the count; call with
 *                       cpg_pos = NULL (or cap too small) to size the output
 *   cpn_chr_part        estimation regions of [reg_st, reg_end] (whole chromosome: 1, chr_size; BED interval: its ends):
 *                       windows of size_est_reg whose right boundary is moved off the CpG group containing it; returns the
 *                       number of regions (call with NULL outputs to size)
 *   cpn_grp_info_count  N[r] CpG sites and L[r] CpG groups of region r (L = 0 when N <= 1: the reference stops there)
 *   cpn_grp_info_fill   cpg_off / grp_off = exclusive cumulative sums of N / L; writes cpg_pos, rho_n [sum N], Nl, rho_l, grp_st,
 *                       grp_end [sum L], d_n [sum max(N-1,0)], d_l [sum max(L-1,0)] — for batches of regions with N > 1 these
 *                       are exactly the CSR arrays (d_* at off[r]-r) of cpn_fit_batch / cpn_stat_sums_batch.               */
/*   cpn_nls_reg_info_batch  cpn_nls_reg_info for R regions (threaded): with sub_st == NULL writes K[r]; otherwise sub_off is the
 *                       exclusive cumulative sum of K and the four sub-region tables are filled                              */
int cpn_nls_reg_info_batch(int64_t R, const int64_t* chrst, const int64_t* chrend, int64_t max_size_subreg, const int64_t* cpg_off,
                           const int64_t* cpg_pos, int32_t n_threads, int64_t* K, const int64_t* sub_off, int64_t* sub_st, int64_t* sub_end,
                           int64_t* sub_p, int64_t* sub_q);
/*   cpn_fasta_*         FASTA file (plain or .gz) held in memory, line breaks removed (get_genome_info / get_chr_fa_rec,
 *                       src/genome_analysis.jl:19-61,283-312); name = identifier up to the first blank; cpn_fasta_seq returns a
 *                       library-owned pointer to cpn_fasta_length bytes, valid until cpn_fasta_close                          */
typedef struct cpn_fasta cpn_fasta;
int  cpn_fasta_open(const char* path, cpn_fasta** out);
void cpn_fasta_close(cpn_fasta* fa);
const char* cpn_fasta_open_error(void);
int64_t cpn_fasta_num_records(const cpn_fasta* fa);
const char* cpn_fasta_name(const cpn_fasta* fa, int64_t i);
int64_t cpn_fasta_length(const cpn_fasta* fa, int64_t i);
const char* cpn_fasta_seq(const cpn_fasta* fa, int64_t i);
int64_t cpn_cpg_scan(const char* seq, int64_t n, int32_t n_threads, int64_t* cpg_pos, int64_t cap);
int64_t cpn_chr_part(const int64_t* cg, int64_t ncg, int64_t chr_size, int64_t reg_st, int64_t reg_end, int64_t size_est_reg,
                     int64_t min_grp_dist, int64_t* out_st, int64_t* out_end, int64_t cap);
int cpn_grp_info_count(const int64_t* cg, int64_t ncg, int64_t R, const int64_t* chrst, const int64_t* chrend, int64_t min_grp_dist,
                       int32_t n_threads, int64_t* N, int64_t* L);
int cpn_grp_info_fill(const int64_t* cg, int64_t ncg, int64_t chr_size, int64_t R, const int64_t* chrst, const int64_t* chrend,
                      int64_t min_grp_dist, const int64_t* cpg_off, const int64_t* grp_off, int32_t n_threads, int64_t* cpg_pos,
                      double* rho_n, double* d_n, double* Nl, double* rho_l, double* d_l, int64_t* grp_st, int64_t* grp_end);

/* ---- output text (host only, no GPU needed) -----------------------------------------------------------------------------
 * The reference's writers (src/input_output.jl:95-462) print with Julia string interpolation; these entry points produce the
 * same bytes: Float64 as Julia's shortest round-trip text ("0.0001", "1.0e-5", "1.234567e6", "NaN"), round(x, digits=d) as
 * round-half-even of x*10^d divided back.
 *   cpn_format_f64  n values joined by sep (digits < 0: no rounding); returns the byte count (call with out = NULL to size)
 *   cpn_write_rows  appends "chr \t a[i] \t b[i] \t v[i] [\t w[i]] \n" for every i with non-NaN v[i] (write_output_ex/_exx: digits -1;
 *                   _mml/_nme and _tmml/_tnme: digits 4, w = p-values; _tcmd: digits 4, clamp0 = 1: max(0, v) first)
 *   cpn_write_theta appends write_output_ϕ rows: chr, chrst, chrend, "a,b,c", α list, β list with α = (a + b*rho_l)*Nl and
 *                   β = c/d_l (get_αβ_from_ϕ, src/link_functions.jl:16-25; Nl = NULL means all ones, the per-CpG chain)      */
int64_t cpn_format_f64(const double* x, int64_t n, char sep, int32_t digits, char* out, int64_t cap);
int cpn_write_rows(const char* path, const char* chr, int64_t n, const int64_t* a, const int64_t* b, const double* v, int32_t digits,
                   int32_t clamp0, const double* w);
/*   cpn_add_qvalues  add_qvalues_to_path (src/multiple_hypothesis.jl:14-49): appends the Benjamini-Hochberg q-value of the p column
 *                   (NaN where p is NaN) to every row of a tmml / tnme / tcmd file, re-printing the float fields as writedlm does     */
int cpn_add_qvalues(const char* path);
/*   cpn_theta_*      model (theta) files as read_mod_file_chr uses them (src/input_output.jl:581-630): chromosome, chrst, chrend and
 *                   phi = (a, b, c) of every line, in file order; the alpha / beta lists that may follow are skipped                 */
typedef struct cpn_theta cpn_theta;
int  cpn_theta_open(const char* path, cpn_theta** out);
void cpn_theta_close(cpn_theta* th);
const char* cpn_theta_open_error(void);
int64_t cpn_theta_num_rows(const cpn_theta* th);
int64_t cpn_theta_num_chr(const cpn_theta* th);
const char* cpn_theta_chr_name(const cpn_theta* th, int64_t i);
int cpn_theta_copy(const cpn_theta* th, int32_t* chr_id, int64_t* chrst, int64_t* chrend, double* phi);
int cpn_write_theta(const char* path, const char* chr, int64_t R, const int64_t* chrst, const int64_t* chrend, const double* phi,
                    const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l);

/*   cpn_bam_*        get_calls_bam (src/wgbs.jl:222-323, with clean_records :93-118, order_bams :66-78, get_align_strand :23-55,
 *                   try_olaps :129-139) for ALL estimation regions of a chromosome in two calls.  cpn_bam_open inflates a BAM file
 *                   once (BGZF blocks in parallel, n_threads <= 0: all cores) and indexes its records by reference and position;
 *                   no .bai is needed.  Host-only.
 *   cpn_bam_record        fields of record i in file order, as the reference's accessors return them: position and rightposition
 *                         1-based (BAM.position, BAM.rightposition), FLAG, MAPQ, presence of the XM / XS tags, template name and the
 *                         XM:Z string (pointers into the handle, not NUL-terminated: use the lengths); any output may be NULL
 *   cpn_bam_count_reads   n_reads[r] = templates of `chr` with at least one call on a CpG site of region r: records overlapping
 *                         [roi_st[r], min(roi_end[r], chr_size - 151)], mapped, with XM, without XS, FLAG in {0,16,83,99,147,163},
 *                         MAPQ > 39; single end (pe = 0): FLAG 0 / 16 only; paired end: names with exactly two such records,
 *                         R1 = the mate that starts first, FLAG pairs (99,147) (83,163) (147,99) (163,83).  cpg_off [R+1] /
 *                         cpg_pos: CSR of the regions' CpG positions (1-based, ascending).  trim[4] = (R1 5', R1 3', R2 5', R2 3')
 *                         bases removed from the XM strings (config.trim).  A chromosome the BAM header does not know gives zeros.
 *   cpn_bam_fill_calls    calls [sum_r n_reads[r] * N_r] int8, read-major per region (-1 unmethylated 'z', +1 methylated 'Z', 0 not
 *                         covered): the layout cpn_wgbs_pack / cpn_analyze_wgbs_batch take; cell_off[r] = sum_{q<r} n_reads[q] * N_q.
 *                         Rows follow the order of the reference's templates (first appearance among the region's records).
 * Errors: cpn_bam_open returns non-zero and leaves the text in cpn_bam_open_error() (thread-local); an unexpected FLAG (the
 * reference prints a message and exits, wgbs.jl:31-33, 45-47) makes the region calls return CPN_ERR_ARG with the text in
 * cpn_bam_error().  One difference: a mate R2 that ends before R1's string does contributes nothing (the reference throws).       */
typedef struct cpn_bam cpn_bam;
int  cpn_bam_open(const char* path, int32_t n_threads, cpn_bam** out);
void cpn_bam_close(cpn_bam* bam);
const char* cpn_bam_open_error(void);
const char* cpn_bam_error(const cpn_bam* bam);
int64_t cpn_bam_num_records(const cpn_bam* bam);
int64_t cpn_bam_num_refs(const cpn_bam* bam);
const char* cpn_bam_ref_name(const cpn_bam* bam, int64_t i);
int64_t cpn_bam_ref_len(const cpn_bam* bam, int64_t i);
int cpn_bam_record(const cpn_bam* bam, int64_t i, int32_t* ref, int64_t* pos, int64_t* rightpos, int32_t* flag, int32_t* mapq,
                   int32_t* has_xm, int32_t* has_xs, const char** qname, int64_t* qname_len, const char** xm, int64_t* xm_len);
int cpn_bam_count_reads(const cpn_bam* bam, const char* chr, int64_t R, const int64_t* roi_st, const int64_t* roi_end,
                        const int64_t* cpg_off, const int64_t* cpg_pos, int32_t pe, const int64_t* trim, int32_t n_threads,
                        int64_t* n_reads);
int cpn_bam_fill_calls(const cpn_bam* bam, const char* chr, int64_t R, const int64_t* roi_st, const int64_t* roi_end,
                       const int64_t* cpg_off, const int64_t* cpg_pos, int32_t pe, const int64_t* trim, int32_t n_threads,
                       const int64_t* cell_off, int8_t* calls);

#ifdef __cplusplus
}
#endif
#endif /* CPELNANO_B200_H */
