#!/usr/bin/env python3
"""bench.py — the estimation hot path on B200 next to the reference algorithm on the host, one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload chr22|targeted|whole-genome|two-sample|group] [--impl reference]

Workloads = the configs of BASELINE.json; the default, and the one `metric` is quoted on, is chr22.

  chr22         configs[1]: 1e5 estimation regions PER GPU (weak scaling) cycling through the 137 real 3 kb CpG layouts of the bundled
                hg38 window (L≈44 groups, N≈63 CpGs), 30 simulated nanopore reads each (cpel_samp_ont), 10 EM starts, ≤ 20 EM
                iterations, 350 bp sub-regions.  Step = one cpn_analyze_batch call = the batched pmap_analyze_reg
                (src/nanopore.jl:292-352: get_ϕhat! → rscle_grp_mod! → get_stat_sums!).
  targeted      configs[0]: the 33 regions of the bundled targeted example (tests/golden/targeted_example.npz: the reference's own
                CpG layouts and fitted ϕ̂, used as the truth the reads are simulated from — the raw calls file is not in the
                checkout), examples/targeted_example.jl's config (5 starts, ≤ 50 EM iterations, 351 bp sub-regions).
  whole-genome  configs[2]: 3e6 regions IN TOTAL (strong scaling), sharded over the ranks by cost (sharding.shard_bounds); the
                end-to-end step ends with the final result gather on rank 0 (NCCL over NVLink, then one device→host copy).
  two-sample    configs[3]: region pairs × 1000 permutations of 30+30 pooled reads → 2000 ten-start refits, summaries, CMD and
                p-value counting per region in ONE cpn_two_samp_test_batch call; EM starts and permutations are drawn on the
                device from a seed (nothing random crosses PCIe).  Default 2000 pairs per GPU (weak).
  group         configs[4]: 6 vs 6 fitted models per region, 3e6 regions in total (strong): summaries of the 12 models, CMD of
                all 66 pairs, all 924 labelings (cpn_grp_test_batch); the matched variant (64 sign patterns) is timed beside it.

JSON keys (driver contract): value = whole-job regions/s with inputs resident in HBM (device time, max over ranks); e2e = the
same through the host-pointer C ABI with pinned host buffers, host↔device copies (and, for strong scaling, the gather) inside the
timed region; roofline = dominant kernel class timed per launch with CUDA events on the launching stream against the FP64 DFMA
rate measured on this GPU in this run (MEASURED_PEAKS.json has no FP64 entry) — the path is FP64-bound, the HBM view is reported
alongside; cpu_baseline = the oracle (C++ restatement of the reference's algorithm — Julia is not installed, the reference itself
cannot run) on all host cores on a bounded sample, with the output-level parity of that sample.  --impl reference times that CPU
path alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "regions/sec fitted (EM+MML/NME)"
UNIT = "regions/s"
C_E = 48.0          # algorithmic FP64 flop per (read, group) of one E-step pass
C_M = 200.0         # per group of one gradient+Jacobian sweep of log Z (160 + 2 exp × 20), DESIGN.md §4
C_Y = 12.0          # per (read, group) of the final log p(y) pass
C_S, C_G = 60.0, 30.0   # summaries: per CpG, per (sub-region, CpG)  
C_CMD = 90.0        # per CpG of one mixture chain of comp_cmd (k_cmd_lane: 2 exp × 20 + forward message 8 + 2×2 product 24 + offsets/rescale 18)
HBM_PEAK = 6551.7   # GB/s, MEASURED_PEAKS.json
BOX = (20.0, 150.0, 50.0)
CPU_REGIONS_PER_THREAD = 16   # CPU legs of the fit workloads: regions per host thread and step (one region ≈ 1 s of one core; 16 per
                              # thread keeps the end-of-step imbalance of the dynamic schedule near 1/16 for both the cpu_baseline
                              # block and the --impl reference arm, which therefore time the same sample the same way)
WORKLOADS = ("chr22", "targeted", "whole-genome", "two-sample", "group")


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# kernel behind each timing class of cpn_get_profile (the E-step of a fit is the paired kernel; plain regions go through its
# resident work-queue form, views of the two-sample test through k_estep_pair<views>)
KERNEL_NAME = {"estep": "k_estep_pair_ws", "mstep": "k_mstep_ws"}


def ncu_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture of this workload
    (profiles/r02_ncu_traffic.json, written from `ncu --set full` reports); None when no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
            t = json.load(f)
        e = t.get(workload, {}).get(kernel)
        return e
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        self.proc = p
        for line in p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])
            if self.stop_flag.is_set():
                break
        try:
            p.kill()
        except Exception:
            pass

    def summary(self):
        self.stop_flag.set()
        time.sleep(0.15)
        try:
            self.proc.kill()
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ======================================================================================================================
# workload construction (host, numpy) — no product library involved, so the reference arm never maps libcpelnano_b200.so
# ======================================================================================================================
def sub_regions(orc_or_lib, nb, size):
    """get_nls_reg_info! per region; only a few hundred distinct layouts exist, so each is computed once.  `orc_or_lib`: anything
    with .get_nls_reg_info / .nls_reg_info (the oracle in the reference arm, the library otherwise — both are bit-exact)."""
    f = getattr(orc_or_lib, "get_nls_reg_info", None) or orc_or_lib.nls_reg_info
    cache, sp, sq, ks = {}, [], [], []
    for r in range(nb.R):
        key = (int(nb.chrst[r]), int(nb.chrend[r]), int(nb.cpg_off[r + 1] - nb.cpg_off[r]))
        if key not in cache:
            cache[key] = f(key[0], key[1], size, nb.cpg_pos[nb.cpg_off[r]:nb.cpg_off[r + 1]])
        _, _, p, q = cache[key]
        sp.append(p); sq.append(q); ks.append(len(p))
    sub_off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int64)
    return sub_off, np.concatenate(sp).astype(np.int64), np.concatenate(sq).astype(np.int64)


def build_workload(cpn, lib, R, M, n_init, seed, subreg=350, layouts=None, phi_true=None):
    nb = cpn.simulations.make_nano_batch(R, M=M, seed=seed, layouts=layouts, phi_true=phi_true)
    phi0 = cpn.simulations.gen_phi0s(np.random.default_rng(seed + 17), R, n_init)
    sub_off, sub_p, sub_q = sub_regions(lib, nb, subreg)
    arrays = dict(grp_off=nb.grp_off, Nl=nb.Nl, rho_l=nb.rho_l, d_l=nb.d_l, read_off=nb.read_off, lu=nb.log_pyx_u, lm=nb.log_pyx_m,
                  phi0=np.ascontiguousarray(phi0.reshape(-1)), cpg_off=nb.cpg_off, rho_n=nb.rho_n, d_n=nb.d_n, sub_off=sub_off,
                  sub_p=sub_p, sub_q=sub_q)
    return nb, arrays


def targeted_layouts():
    z = np.load(os.path.join(ROOT, "tests", "golden", "targeted_example.npz"))
    lay = {k: z[k] for k in ("chrst", "chrend", "cpg_off", "grp_off", "cpg_pos", "rho_n", "d_n", "Nl", "rho_l", "d_l")}
    return lay, z["phi"].astype(np.float64)


# CSR families of the fit workloads: data array → (offset array, 1 if the array has one entry fewer per region)
FAMILY = {"Nl": ("grp_off", 0), "rho_l": ("grp_off", 0), "d_l": ("grp_off", 1), "lu": ("obs_off", 0), "lm": ("obs_off", 0),
          "rho_n": ("cpg_off", 0), "d_n": ("cpg_off", 1), "sub_p": ("sub_off", 0), "sub_q": ("sub_off", 0)}


def tile_shard(arrays, obs_off, base_R, n_init, r0, r1):
    """Arrays of the global regions [r0, r1) of a batch that repeats a base batch of base_R regions (global region g ↔ base
    region g mod base_R).  Regions are independent, so repetition changes neither the per-region work nor the memory traffic —
    it only avoids minutes of numpy generating 3·10⁶ distinct regions."""
    offs = dict(grp_off=arrays["grp_off"], read_off=arrays["read_off"], cpg_off=arrays["cpg_off"], sub_off=arrays["sub_off"], obs_off=obs_off)
    segs = []
    g = r0
    while g < r1:
        a = g % base_R
        b = min(base_R, a + (r1 - g))
        segs.append((a, b))
        g += b - a
    out = {}
    for k, (ok, shrink) in FAMILY.items():
        o = offs[ok]
        out[k] = np.concatenate([arrays[k][o[a] - a * shrink:o[b] - b * shrink] for a, b in segs])
    out["phi0"] = np.concatenate([arrays["phi0"][a * n_init * 3:b * n_init * 3] for a, b in segs])
    for ok in ("grp_off", "read_off", "cpg_off", "sub_off"):
        o = offs[ok]
        out[ok] = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(o[a:b + 1]) for a, b in segs]))]).astype(np.int64)
    return out


def out_shapes(R, arrays):
    sN, sK = int(arrays["cpg_off"][-1]), int(arrays["sub_off"][-1])
    return dict(phi_hat=(3 * R, np.float64), Q=(R, np.float64), status=(R, np.int32), counters=(4 * R, np.int64), logZ=(R, np.float64),
                ex=(sN, np.float64), exx=(sN - R, np.float64), log_g=(4 * sK, np.float64), e1=(sK, np.float64), e2=(sK, np.float64),
                mml=(sK, np.float64), nme=(sK, np.float64))


IN_ORDER = ("grp_off", "Nl", "rho_l", "d_l", "read_off", "lu", "lm", "phi0")
IN2_ORDER = ("cpg_off", "rho_n", "d_n", "sub_off", "sub_p", "sub_q")
OUT_ORDER = ("phi_hat", "Q", "status", "counters", "logZ", "ex", "exx", "log_g", "e1", "e2", "mml", "nme")


def run_step(lib, R, ins, outs, n_init, max_iters, mode, on_device, device, packed=False):
    lib.analyze_batch_raw(R, *(ins[k] for k in IN_ORDER), n_init, max_iters, BOX, mode, *(ins[k] for k in IN2_ORDER),
                          *(outs[k] for k in OUT_ORDER), on_device, device, packed=packed)


# ======================================================================================================================
# CPU legs (oracle): the reference algorithm on the host cores
# ======================================================================================================================
def cpu_fit_step(orc, nb, arrays, idx, n_init, max_iters, threads, perturb=False):
    """The reference algorithm for the regions `idx`: multi-start EM (all threads, one region per task — the analogue of pmap)
    followed by the per-CpG summaries.  Returns (seconds, fit, (mml, nme) rows)."""
    import parity_report as pr
    sub = nb.subset(idx)
    phi0 = arrays["phi0"].reshape(nb.R, n_init, 3)[idx]
    so = arrays["sub_off"]
    ks = np.concatenate([np.arange(so[r], so[r + 1]) for r in idx]) if len(idx) else np.zeros(0, dtype=np.int64)
    sub_off = np.concatenate([[0], np.cumsum([so[r + 1] - so[r] for r in idx])]).astype(np.int64)
    t0 = time.perf_counter()
    fit = pr.oracle_fit(orc, sub, phi0, n_init, max_iters, perturb=perturb, n_threads=threads)
    rows = pr.oracle_rows(orc, sub, sub_off, arrays["sub_p"][ks], arrays["sub_q"][ks], fit["phi_hat"], fit["pick"])
    return time.perf_counter() - t0, fit, rows, sub_off


def cpu_two_sample_step(orc, w, n_perm, threads):
    """pmap_two_samp_null_stats composed on the oracle for the first region pair and n_perm permutations (2·n_perm ten-start
    refits as one all-thread batch, summaries, CMD); returns seconds."""
    nb, n1, n_init = w["nb"], w["n1"], w["n_init"]
    rg = nb.region(0)
    M = rg["M"]
    g_off, r_off, Nl, rho, dl, lu, lm, p0 = [0], [0], [], [], [], [], [], []
    for i in range(n_perm):
        idx = orc.gen_perm(w["seed"], 0, i, M, n1) - 1
        rest = np.setdiff1d(np.arange(M), idx)
        for h, sel in enumerate((idx, rest)):
            g_off.append(g_off[-1] + rg["L"]); r_off.append(r_off[-1] + len(sel))
            Nl.append(rg["Nl"]); rho.append(rg["rho_l"]); dl.append(rg["d_l"])
            lu.append(rg["lu"][sel].reshape(-1)); lm.append(rg["lm"][sel].reshape(-1))
            p0.append([orc.gen_phi0(w["seed"], 0, i, h, s) for s in range(n_init)])
    cat = np.concatenate
    t0 = time.perf_counter()
    f = orc.fit_batch(np.array(g_off), cat(Nl), cat(rho), cat(dl), np.array(r_off), cat(lu), cat(lm), np.array(p0), n_init, w["max_iters"],
                      n_threads=threads)
    a, b = nb.cpg_off[0], nb.cpg_off[1]
    p, q = w["sub_p"][:w["sub_off"][1]], w["sub_q"][:w["sub_off"][1]]
    for i in range(n_perm):
        if f["pick"][2 * i] < 0 or f["pick"][2 * i + 1] < 0:
            continue
        s1, s2 = (orc.get_stat_sums(f["phi_hat"][2 * i + h], nb.rho_n[a:b], nb.d_n[a:b - 1], p, q) for h in (0, 1))
        orc.comp_cmd(f["phi_hat"][2 * i], f["phi_hat"][2 * i + 1], s1["nme"], s2["nme"], nb.rho_n[a:b], nb.d_n[a:b - 1], p, q)
    return time.perf_counter() - t0


def cpu_group_step(orc, w, regions, threads):
    """unmat_est_reg_test on the oracle for `regions` regions: get_stat_sums! of the 12 models, comp_cmd of the 66 pairs, the
    924-labeling sweep; one region per task on `threads` threads.  Returns seconds."""
    from concurrent.futures import ThreadPoolExecutor
    S, n1, n2 = w["S"], w["n1"], w["n2"]
    labs = w["labelings"]

    def one(r):
        a, b = w["cpg_off"][r], w["cpg_off"][r + 1]
        k0, k1 = w["sub_off"][r], w["sub_off"][r + 1]
        rho, dn, p, q = w["rho_n"][a:b], w["d_n"][a - r:b - r - 1], w["sub_p"][k0:k1], w["sub_q"][k0:k1]
        ss = [orc.get_stat_sums(w["phi"][r, s], rho, dn, p, q) for s in range(S)]
        K = k1 - k0
        tbl = np.full((S, S, K), np.nan)
        for i in range(S):
            for j in range(i + 1, S):
                tbl[i, j] = orc.comp_cmd(w["phi"][r, i], w["phi"][r, j], ss[i]["nme"], ss[j]["nme"], rho, dn, p, q)
        orc.unmat_est_reg_test(n1, n2, np.stack([s["mml"] for s in ss]), np.stack([s["nme"] for s in ss]), tbl, labs, True)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:          # ctypes releases the GIL inside the oracle calls
        list(ex.map(one, range(regions)))
    return time.perf_counter() - t0


# ======================================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="chr22", choices=WORKLOADS)
    ap.add_argument("--regions", type=int, default=0, help="chr22 / two-sample: regions per GPU; whole-genome / group: regions in total (0 = the config's size)")
    ap.add_argument("--reads", type=int, default=30)
    ap.add_argument("--n-init", type=int, default=0)
    ap.add_argument("--max-em-iters", type=int, default=0)
    ap.add_argument("--perms", type=int, default=1000)
    ap.add_argument("--mstep-mode", default="analytic", choices=["analytic", "fd"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="regions in the CPU sample (0 = sized for ≈15 s)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    wl = args.workload
    dflt_regions = {"chr22": 100_000, "targeted": 33, "whole-genome": 3_000_000, "two-sample": 2000, "group": 3_000_000}[wl]
    args.regions = args.regions or dflt_regions
    args.n_init = args.n_init or (5 if wl == "targeted" else 10)
    args.max_em_iters = args.max_em_iters or (50 if wl == "targeted" else 20)
    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # rank 0 prints exactly one JSON line on stdout: whatever NCCL logs (its version banner appears from NCCL_DEBUG=VERSION up,
        # WARN included) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import cpelnano_b200 as cpn
    lib = cpn.Library()
    lib.ctx(local_rank)
    env = dict(args=args, rank=rank, local_rank=local_rank, world=world, torch=torch, dist=dist, cpn=cpn, lib=lib)
    line = {"chr22": bench_fit, "targeted": bench_fit, "whole-genome": bench_fit, "two-sample": bench_two_sample, "group": bench_group}[wl](env)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def base_line(args, world, config, scaling):
    return {"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config}


def fit_config(args, world):
    wl = args.workload
    names = {"chr22": f"synthetic chr22-scale: {args.regions} regions x {args.reads} reads per GPU, real hg38 CpG layouts (L~44, N~63), "
                      f"{args.n_init} EM starts, <={args.max_em_iters} EM iters, EM fit + MML/NME",
             "targeted": f"targeted example: the {args.regions} regions of examples/targeted_example (reference layouts and fitted phi as simulation truth), "
                         f"{args.reads} simulated reads, {args.n_init} EM starts, <={args.max_em_iters} EM iters, 351 bp sub-regions",
             "whole-genome": f"synthetic whole-genome: {args.regions} regions x {args.reads} reads IN TOTAL (base batch of 1e5 simulated regions repeated), "
                             f"{args.n_init} EM starts, <={args.max_em_iters} EM iters, EM fit + MML/NME, sharded by cost, results gathered on rank 0"}
    return {"workload": names[wl], "regions": args.regions, "regions_are": "per GPU" if wl == "chr22" else "total", "reads_per_region": args.reads,
            "n_init": args.n_init, "max_em_iters": args.max_em_iters,
            "parallelism": f"regions sharded over {world} GPU(s), no data-path collective" + ("" if wl == "chr22" else "; final result gather on rank 0"),
            "l2": "inputs (≈10 KB of packed emissions per region) are larger than the 126 MB L2" if wl != "targeted" else
                  "33 regions (0.4 MB): L2-resident, launch-latency-bound by construction"}


# ----------------------------------------------------------------------------------------------------------------------
# chr22 / targeted / whole-genome: cpn_analyze_batch
# ----------------------------------------------------------------------------------------------------------------------
def bench_fit(env):
    args, rank, local_rank, world, torch, dist, cpn, lib = (env[k] for k in ("args", "rank", "local_rank", "world", "torch", "dist", "cpn", "lib"))
    wl = args.workload
    strong = wl != "chr22"
    mode = cpn.MSTEP_ANALYTIC if args.mstep_mode == "analytic" else cpn.MSTEP_FD
    n_init, max_iters = args.n_init, args.max_em_iters
    config = fit_config(args, world)
    # ---- this rank's regions
    if wl == "chr22":
        nb, arrays = build_workload(cpn, lib, args.regions, args.reads, n_init, cpn.simulations.MASTER_SEED + 1000 * rank)
        R, base_R, r0 = args.regions, args.regions, 0
        mine = arrays
    else:
        if wl == "targeted":
            lay, phi_true = targeted_layouts()
            nb, arrays = build_workload(cpn, lib, len(phi_true), args.reads, n_init, cpn.simulations.MASTER_SEED, subreg=351, layouts=lay,
                                        phi_true=phi_true)
            base_R = nb.R
        else:
            base_R = min(100_000, args.regions)
            nb, arrays = build_workload(cpn, lib, base_R, args.reads, n_init, cpn.simulations.MASTER_SEED)
        total = args.regions
        if wl == "whole-genome":          # never push the box towards its memory limit: inputs + outputs are pinned host memory
            import psutil
            per_region = 25_500 * 2.2
            avail = psutil.virtual_memory().available
            cap = int(0.7 * avail / per_region) * world
            if total > cap:
                total = max(base_R, cap // base_R * base_R)
                config["regions_reduced_to_fit_host_memory"] = total
        cost = cpn.sharding.region_costs(arrays["grp_off"], arrays["read_off"], n_init)
        reps = -(-total // base_R)
        bounds = cpn.sharding.shard_bounds(np.tile(cost, reps)[:total], world)
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        R = r1 - r0
        mine = tile_shard(arrays, nb.obs_off, base_R, n_init, r0, r1) if (total != base_R or world > 1) else arrays
        config["regions"] = total
        config["shard_bounds"] = [int(b) for b in bounds]
    shapes = out_shapes(R, mine)
    tdt = lambda dt: torch.float64 if dt == np.float64 else (torch.int32 if dt == np.int32 else torch.int64)      # noqa: E731
    host_in = {}
    for k in IN_ORDER + IN2_ORDER:          # one array at a time: the unpinned copy is dropped before the next is pinned
        host_in[k] = torch.from_numpy(np.ascontiguousarray(mine[k])).pin_memory()
    offs_np = {k: np.ascontiguousarray(mine[k]) for k in ("grp_off", "read_off", "cpg_off", "sub_off")}
    del mine
    # all outputs of this rank live in ONE pinned buffer (the gather sends it as a whole); the library writes into views of it
    fields, o = {}, 0
    for k in OUT_ORDER:
        n, dt = shapes[k]
        fields[k] = (o, n, dt)
        o += (n * np.dtype(dt).itemsize + 15) & ~15
    out_bytes = o
    host_out_buf = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
    host_out = {k: host_out_buf[o:o + n * np.dtype(dt).itemsize].view(tdt(dt)) for k, (o, n, dt) in fields.items()}
    h2d_two_tables = int(sum(v.numel() * v.element_size() for v in host_in.values()))
    d2h = int(sum(v.numel() * v.element_size() for v in host_out.values()))
    p_host_in = {k: v.data_ptr() for k, v in host_in.items()}
    p_host_out = {k: v.data_ptr() for k, v in host_out.items()}
    # the end-to-end leg hands the emissions over packed (cpn_pack_emissions: one LLR per cell + one base per read — what the calls
    # loader / the Julia shim produce when they flatten the reads): same results, half the PCIe bytes
    host_llr = torch.empty(host_in["lu"].numel(), dtype=torch.float64).pin_memory()
    host_base = torch.empty(int(offs_np["read_off"][-1]), dtype=torch.float64).pin_memory()
    lib.pack_emissions(offs_np["grp_off"], offs_np["read_off"], host_in["lu"].numpy(), host_in["lm"].numpy(), out=(host_llr.numpy(), host_base.numpy()))
    p_host_packed = dict(p_host_in, lu=host_llr.data_ptr(), lm=host_base.data_ptr())
    h2d = h2d_two_tables - host_in["lm"].numel() * 8 + host_base.numel() * 8

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = lib.measure_fp64_peak(local_rank)

    # ---- device-resident leg: inputs in HBM, processed in slices of at most 250 000 regions (bounded workspace)
    dev_in = {k: v.cuda(non_blocking=True) for k, v in host_in.items()}
    dev_out = {k: torch.empty(shapes[k][0], dtype=tdt(shapes[k][1]), device="cuda") for k in OUT_ORDER}
    SL = 250_000
    slices = []
    for a in range(0, R, SL):
        b = min(R, a + SL)
        g0, o0, n0, k0 = (int(offs_np[k][a]) for k in ("grp_off", "grp_off", "cpg_off", "sub_off"))
        o0 = int(np.sum(np.diff(offs_np["grp_off"][:a + 1]) * np.diff(offs_np["read_off"][:a + 1]))) if a else 0
        so = {k: torch.from_numpy(offs_np[k][a:b + 1] - offs_np[k][a]).cuda() for k in offs_np}
        e = 8
        ins = {"grp_off": so["grp_off"].data_ptr(), "read_off": so["read_off"].data_ptr(), "cpg_off": so["cpg_off"].data_ptr(),
               "sub_off": so["sub_off"].data_ptr(),
               "Nl": dev_in["Nl"].data_ptr() + e * g0, "rho_l": dev_in["rho_l"].data_ptr() + e * g0, "d_l": dev_in["d_l"].data_ptr() + e * (g0 - a),
               "lu": dev_in["lu"].data_ptr() + e * o0, "lm": dev_in["lm"].data_ptr() + e * o0, "phi0": dev_in["phi0"].data_ptr() + e * a * n_init * 3,
               "rho_n": dev_in["rho_n"].data_ptr() + e * n0, "d_n": dev_in["d_n"].data_ptr() + e * (n0 - a),
               "sub_p": dev_in["sub_p"].data_ptr() + e * k0, "sub_q": dev_in["sub_q"].data_ptr() + e * k0}
        outs = {"phi_hat": dev_out["phi_hat"].data_ptr() + e * 3 * a, "Q": dev_out["Q"].data_ptr() + e * a, "status": dev_out["status"].data_ptr() + 4 * a,
                "counters": dev_out["counters"].data_ptr() + e * 4 * a, "logZ": dev_out["logZ"].data_ptr() + e * a,
                "ex": dev_out["ex"].data_ptr() + e * n0, "exx": dev_out["exx"].data_ptr() + e * (n0 - a), "log_g": dev_out["log_g"].data_ptr() + e * 4 * k0,
                "e1": dev_out["e1"].data_ptr() + e * k0, "e2": dev_out["e2"].data_ptr() + e * k0, "mml": dev_out["mml"].data_ptr() + e * k0,
                "nme": dev_out["nme"].data_ptr() + e * k0}
        slices.append((b - a, ins, outs, so))
    torch.cuda.synchronize()

    def resident_step():
        ms, launches, work = 0.0, 0, np.zeros(4)
        for (n, ins, outs, _) in slices:
            run_step(lib, n, ins, outs, n_init, max_iters, mode, True, local_rank)
            ms += lib.last_device_ms(local_rank); launches += lib.last_launches(local_rank); work += lib.last_work(local_rank)
        return ms, launches, work

    for _ in range(args.warmup):
        resident_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.set_profile(True, local_rank)
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, work = 0.0, 0, np.zeros(4)
    for _ in range(args.steps):
        ms, ln, wk = resident_step()
        dev_ms += ms; launches += ln; work += wk
    barrier()
    t_dev = time.perf_counter() - t0
    prof = lib.get_profile(local_rank)
    lib.set_profile(False, local_rank)
    clocks = sampler.summary()
    del dev_in, dev_out, slices
    torch.cuda.empty_cache()

    # ---- end-to-end leg: host-pointer ABI with pinned buffers (+ the result gather when the job is sharded)
    gather = cpn.sharding.ResultGather(out_bytes, torch.device("cuda", local_rank)) if (strong and world > 1) else None

    def e2e_step(packed=True):
        run_step(lib, R, p_host_packed if packed else p_host_in, p_host_out, n_init, max_iters, mode, False, local_rank, packed=packed)
        if gather is not None:
            gather.run(host_out_buf)

    # the two-table entry point (cpn_analyze_batch) timed beside it, and its outputs kept to show the two routes agree bit for bit
    for _ in range(args.warmup):
        e2e_step(False)
    keep = {k: host_out[k].clone() for k in ("phi_hat", "mml", "nme")}
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step(False)
    barrier()
    t_e2e_two = time.perf_counter() - t0
    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        e2e_step()
        e2e_steps.append(round((time.perf_counter() - t1) * 1e3, 2))
    barrier()
    t_e2e = time.perf_counter() - t0
    gb = gather.bytes_moved() if gather is not None else (0, 0)
    same_bits = all(torch.equal(torch.nan_to_num(keep[k], nan=-7.0), torch.nan_to_num(host_out[k], nan=-7.0)) for k in keep)
    red = torch.tensor([t_dev, t_e2e, dev_ms, t_e2e_two], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(R), float(h2d + gb[0]), float(d2h + gb[1]), float(launches)] + list(work), dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_dev, t_e2e, dev_ms, t_e2e_two = (float(x) for x in red.cpu())
    R_all, h2d_all, d2h_all, launches_all = (float(x) for x in tot.cpu()[:4])
    if rank != 0:
        return None
    line = base_line(args, world, config, "strong" if strong else "weak")
    counters = host_out["counters"].numpy().reshape(R, 4)
    pick = host_out["status"].numpy()
    Ls, Ms = np.diff(offs_np["grp_off"]).astype(np.float64), np.diff(offs_np["read_off"]).astype(np.float64)
    Ns, Ks = np.diff(offs_np["cpg_off"]).astype(np.float64), np.diff(offs_np["sub_off"]).astype(np.float64)
    fl_e, fl_m, fl_y = C_E * work[0], C_M * work[1], C_Y * work[2]          # rank 0's launches over all timed steps (cpn_last_work)
    fl_s = float(np.sum(C_S * Ns + C_G * Ks * Ns)) * args.steps
    dom = "estep" if prof["estep"][0] >= prof["mstep"][0] else "mstep"
    ms_dom, n_dom = prof[dom]
    achieved = (fl_e if dom == "estep" else fl_m) / (ms_dom * 1e-3) / 1e12 if ms_dom > 0 else 0.0
    bytes_region = 16 * Ms * Ls + 24 * Ls + 16 * Ns + 24 * n_init + 40 + 8 * (2 * Ns - 1) + 16 * Ks + 40
    hbm_ach = float(bytes_region.sum()) / (dev_ms / args.steps * 1e-3) / 1e9
    hbm_peak = load_peaks().get("hbm_gbs", HBM_PEAK)
    tr = ncu_traffic("k_" + dom, wl)
    line.update({"value": R_all * args.steps / t_dev, "ms_per_step": dev_ms / args.steps, "wall_ms_per_step": 1e3 * t_dev / args.steps,
                 "e2e": {"value": R_all * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                         "ms_steps_rank0": e2e_steps, "gather": "NCCL gather of every rank's result buffer to rank 0 + device→host copy, inside the timed step"
                         if gather is not None else None,
                         "entry": "cpn_analyze_batch_packed (emissions as one LLR per (read, group) + one base per read, cpn_pack_emissions)",
                         "two_table_entry": {"entry": "cpn_analyze_batch", "value": R_all * args.steps / t_e2e_two,
                                             "h2d_bytes_per_step": int(h2d_two_tables * world), "bit_identical_outputs_rank0": bool(same_bits)}},
                 "gpu_launches": int(launches_all), "clocks": clocks, "mstep": {"mode": args.mstep_mode},
                 "roofline": {"bound": "fp64", "kernel": KERNEL_NAME[dom], "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": achieved / fp64_peak if fp64_peak else None,
                              "traffic": tr["bytes_per_launch"] if tr else None, "traffic_note": tr["note"] if tr else "no ncu capture committed for this workload",
                              "peak_source": "DFMA loop (8 independent chains per thread, 128x8 threads per SM, 5 launches of >= 10 ms, best) measured on this "
                                             "GPU in this run; MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                              "launches": int(n_dom), "ms_per_launch": ms_dom / max(n_dom, 1),
                              "flop_model": {"estep": "48 flop/(read,group)/pass", "mstep": "200 flop/group/sweep", "source": "cpn_last_work counters"},
                              "whole_step": {"flop": (fl_e + fl_m + fl_y + fl_s) / args.steps,
                                             "achieved": (fl_e + fl_m + fl_y + fl_s) / (dev_ms * 1e-3) / 1e12,
                                             "frac": (fl_e + fl_m + fl_y + fl_s) / (dev_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None},
                              "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak}},
                 "kernel_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in prof.items() if v[1]},
                 "fit_stats": {"converged_regions": float(np.mean(pick >= 0)), "numerical_status_regions": float(np.mean(pick == -2)),
                               "em_iters_per_region": float(counters[:, 0].mean()), "solver_iters_per_region": float(counters[:, 1].mean()),
                               "sweeps_per_region": float(counters[:, 2].mean())}})
    if world == 1 and not args.no_cpu:
        import oracle_py as orc
        import parity_report as pr
        threads = os.cpu_count() or 1
        n_s = min(args.cpu_sample or CPU_REGIONS_PER_THREAD * threads, nb.R)      # the same sample rule as the --impl reference arm
        idx = np.arange(n_s)
        tc, fit_c, rows_c, so_c = cpu_fit_step(orc, nb, arrays, idx, n_init, max_iters, threads)
        line["cpu_baseline"] = {"value": n_s / tc, "unit": UNIT, "cores": threads, "kind": "port", "mstep": "finite differences of finite differences (the reference's NLsolve path)",
                                "sample": f"first {n_s} regions of the same workload, {tc:.1f} s (oracle = C++ restatement of the reference's "
                                          "O(M·L²) algorithm; the Julia reference cannot run here)"}
        # output-level parity on the sample (the tests assert it on 2000 regions): GPU vs oracle next to the oracle vs itself under a
        # one-ulp change of its inputs — rows of the 4-dp MML/NME files that differ, fit status flips, picked-start flips
        k1 = int(so_c[-1])
        got = pr.compare(pick[:n_s], host_out["mml"].numpy()[:k1], host_out["nme"].numpy()[:k1], fit_c["pick"], rows_c[0], rows_c[1], so_c)
        _, fit_p, rows_p, _ = cpu_fit_step(orc, nb, arrays, idx, n_init, max_iters, threads, perturb=True)
        slf = pr.compare(fit_p["pick"], rows_p[0], rows_p[1], fit_c["pick"], rows_c[0], rows_c[1], so_c)
        ok = (fit_c["pick"] >= 0) & (pick[:n_s] >= 0)
        dq = np.abs(fit_c["Q"][ok] - host_out["Q"].numpy()[:n_s][ok]) / np.abs(fit_c["Q"][ok]) if ok.any() else np.zeros(1)
        line["cpu_baseline"]["parity_on_sample"] = {"gpu_vs_oracle": got, "oracle_vs_one_ulp_perturbed_oracle": slf, "median_rel_dQ": float(np.median(dq)),
                                                    "note": "flip rates of 4-dp MML/NME rows, fit status and picked start; the reference algorithm's own "
                                                            "reproducibility is the yardstick (oracle/parity_report.py, tests/test_em_output_parity.py)"}
    return line


# ----------------------------------------------------------------------------------------------------------------------
# two-sample: cpn_two_samp_test_batch, everything random drawn on the device
# ----------------------------------------------------------------------------------------------------------------------
def two_sample_workload(cpn, sub_lib, R, n1, n_init, max_iters, seed):
    """sample 2 shares ϕ with sample 1 for 90 % of the regions and has a shifted ϕ (mean level: sign flip
    of a, b) for 10 %; 30 + 30 reads pooled per region, sample 1 first."""
    rng = np.random.default_rng(seed)
    nb1 = cpn.simulations.make_nano_batch(R, M=n1, seed=seed)
    phi2 = nb1.phi_true.copy()
    alt = rng.random(R) < 0.1
    phi2[alt, :2] *= -1.0
    nb2 = cpn.simulations.make_nano_batch(R, M=n1, seed=seed + 1, phi_true=phi2)
    Ls = np.diff(nb1.grp_off)
    obs = np.concatenate([[0], np.cumsum(2 * n1 * Ls)]).astype(np.int64)
    lu, lm = np.empty(obs[-1]), np.empty(obs[-1])
    for r in range(R):
        a, b, h = obs[r], obs[r + 1], n1 * Ls[r]
        lu[a:a + h] = nb1.log_pyx_u[nb1.obs_off[r]:nb1.obs_off[r + 1]]; lu[a + h:b] = nb2.log_pyx_u[nb2.obs_off[r]:nb2.obs_off[r + 1]]
        lm[a:a + h] = nb1.log_pyx_m[nb1.obs_off[r]:nb1.obs_off[r + 1]]; lm[a + h:b] = nb2.log_pyx_m[nb2.obs_off[r]:nb2.obs_off[r + 1]]
    nb = cpn.simulations.NanoBatch(grp_off=nb1.grp_off, cpg_off=nb1.cpg_off, read_off=(np.arange(R + 1) * 2 * n1).astype(np.int64), obs_off=obs, Nl=nb1.Nl,
                                   rho_l=nb1.rho_l, d_l=nb1.d_l, rho_n=nb1.rho_n, d_n=nb1.d_n, cpg_pos=nb1.cpg_pos, log_pyx_u=lu, log_pyx_m=lm,
                                   chrst=nb1.chrst, chrend=nb1.chrend, phi_true=nb1.phi_true)
    sub_off, sub_p, sub_q = sub_regions(sub_lib, nb, 350)
    return dict(nb=nb, nb1=nb1, nb2=nb2, phi2_true=phi2, sub_off=sub_off, sub_p=sub_p, sub_q=sub_q, n1=n1, n_init=n_init, max_iters=max_iters,
                seed=seed, alt=alt)


def two_sample_config(args, world):
    return {"workload": f"two-sample comparison: {args.regions} region pairs per GPU x {args.perms} permutations of {args.reads}+{args.reads} pooled reads, "
                        f"{args.n_init} EM starts per refit, <={args.max_em_iters} EM iters (cpn_two_samp_test_batch; starts and permutations drawn on the device)",
            "regions_per_gpu": args.regions, "perms": args.perms, "reads_per_sample": args.reads, "n_init": args.n_init, "max_em_iters": args.max_em_iters,
            "parallelism": f"region pairs sharded over {world} GPU(s), no data-path collective",
            "l2": "each region's pooled emissions (21 KB) are re-read by its 2000 refits from L1/L2 by design; a step touches "
                  f"{args.regions} x 21 KB of emissions and ≈ {args.regions * args.perms * 2 * args.n_init * 215 / 1e9:.1f} GB of EM state"}


def bench_two_sample(env):
    args, rank, local_rank, world, torch, dist, cpn, lib = (env[k] for k in ("args", "rank", "local_rank", "world", "torch", "dist", "cpn", "lib"))
    R, P, n1, n_init, max_iters = args.regions, args.perms, args.reads, args.n_init, args.max_em_iters
    mode = cpn.MSTEP_ANALYTIC if args.mstep_mode == "analytic" else cpn.MSTEP_FD
    seed = cpn.simulations.MASTER_SEED + 1000 * rank
    w = two_sample_workload(cpn, lib, R, n1, n_init, max_iters, seed)
    nb = w["nb"]
    # the two observed models: ϕ̂ of each sample's own fit (what the model files of diff_two_samp_comp hold); unfitted → the truth
    fits = []
    for b_, truth in ((w["nb1"], nb.phi_true), (w["nb2"], w["phi2_true"])):
        f = lib.fit_batch(b_.grp_off, b_.Nl, b_.rho_l, b_.d_l, b_.read_off, b_.log_pyx_u, b_.log_pyx_m, None, n_init, max_iters, mstep_mode=mode,
                          device=local_rank)
        fits.append(np.where((f["pick"] >= 0)[:, None], f["phi_hat"], truth))
    perm_off = (np.arange(R + 1) * P).astype(np.int64)
    n1v = np.full(R, n1, dtype=np.int64)
    lib.set_seed(seed, local_rank)
    fp64_peak = lib.measure_fp64_peak(local_rank)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()      # noqa: E731
    ins = {k: pin(v) for k, v in dict(Nl=nb.Nl, rho_l=nb.rho_l, d_l=nb.d_l, lu=nb.log_pyx_u, lm=nb.log_pyx_m, phi1=fits[0], phi2=fits[1], rho_n=nb.rho_n,
                                      d_n=nb.d_n, sub_p=w["sub_p"], sub_q=w["sub_q"]).items()}
    sK = int(w["sub_off"][-1])
    outs = {"T": torch.empty(3 * sK, dtype=torch.float64).pin_memory(), "cnt": torch.empty(3 * sK, dtype=torch.int64).pin_memory(),
            "nv": torch.empty(3 * sK, dtype=torch.int64).pin_memory(), "p": torch.empty(3 * sK, dtype=torch.float64).pin_memory()}
    h2d = int(sum(v.numel() * v.element_size() for v in ins.values())) + 8 * 6 * (R + 1)
    d2h = int(sum(v.numel() * v.element_size() for v in outs.values()))
    res = {}

    def step():
        res["out"] = lib.two_samp_test_batch(nb.grp_off, ins["Nl"].numpy(), ins["rho_l"].numpy(), ins["d_l"].numpy(), nb.read_off, ins["lu"].numpy(),
                                             ins["lm"].numpy(), n1v, ins["phi1"].numpy(), ins["phi2"].numpy(), perm_off, None, None, None, n_init,
                                             nb.cpg_off, ins["rho_n"].numpy(), ins["d_n"].numpy(), w["sub_off"], ins["sub_p"].numpy(), ins["sub_q"].numpy(),
                                             max_iters, BOX, mode, device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.set_profile(True, local_rank)
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, work, steps_ms = 0.0, 0, np.zeros(4), []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        step()
        steps_ms.append(round((time.perf_counter() - t1) * 1e3, 1))
        dev_ms += lib.last_device_ms(local_rank); launches += lib.last_launches(local_rank); work += lib.last_work(local_rank)
    barrier()
    t_e2e = time.perf_counter() - t0
    prof = lib.get_profile(local_rank)
    lib.set_profile(False, local_rank)
    clocks = sampler.summary()
    red = torch.tensor([t_e2e, dev_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_e2e, dev_ms = (float(x) for x in red.cpu())
    if rank != 0:
        return None
    line = base_line(args, world, two_sample_config(args, world), "weak")
    line["metric"] = "region pairs/sec tested (two-sample permutation test: 2·P multi-start EM refits + MML/NME/CMD + p-value counts per region)"
    fl_e, fl_m = C_E * work[0], C_M * work[1]
    dom = "estep" if prof["estep"][0] >= prof["mstep"][0] else "mstep"
    ms_dom, n_dom = prof[dom]
    achieved = (fl_e if dom == "estep" else fl_m) / (ms_dom * 1e-3) / 1e12 if ms_dom > 0 else 0.0
    pv = res["out"]["pval"]
    tr = ncu_traffic("k_" + dom, "two-sample")
    line.update({"value": world * R * args.steps / (dev_ms * 1e-3), "ms_per_step": dev_ms / args.steps,
                 "value_note": "CUDA-event time of the same host-pointer call (the one-off upload of 42 KB per region is < 0.1 % of a step): "
                               "the inputs of this workload are the pooled reads, everything else is generated and consumed on the device",
                 "e2e": {"value": world * R * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                         "ms_steps_rank0": steps_ms},
                 "refits_per_s": world * R * 2 * P * args.steps / (dev_ms * 1e-3), "em_instances_per_step": float(work[3] / args.steps),
                 "gpu_launches": int(tot.item()), "clocks": clocks, "mstep": {"mode": args.mstep_mode},
                 "roofline": {"bound": "fp64", "kernel": "k_estep_pair<views>" if dom == "estep" else KERNEL_NAME[dom], "achieved": achieved, "peak": fp64_peak,
                              "unit": "TFLOP/s",
                              "frac": achieved / fp64_peak if fp64_peak else None, "traffic": tr["bytes_per_launch"] if tr else None,
                              "traffic_note": tr["note"] if tr else "no ncu capture committed for this workload",
                              "peak_source": "DFMA loop measured on this GPU in this run", "launches": int(n_dom), "ms_per_launch": ms_dom / max(n_dom, 1),
                              "flop_model": {"estep": "48 flop/(read,group)/pass", "mstep": "200 flop/group/sweep", "source": "cpn_last_work counters"},
                              "whole_step": {"flop": (fl_e + fl_m) / args.steps, "achieved": (fl_e + fl_m) / (dev_ms * 1e-3) / 1e12}},
                 "kernel_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in prof.items() if v[1]},
                 "test_stats": {"frac_subregions_with_pvalue": float(np.mean(np.isfinite(pv))),
                                "frac_p_lt_0.05_null_regions": float(np.nanmean(pv.reshape(-1)[np.repeat(np.repeat(~w["alt"], np.diff(w["sub_off"])), 3)] < 0.05))
                                if np.isfinite(pv).any() else None}})
    if world == 1 and not args.no_cpu:
        import oracle_py as orc
        threads = os.cpu_count() or 1
        n_perm = max(4, threads // 2)
        t1 = cpu_two_sample_step(orc, w, n_perm, threads)
        n_perm = int(max(n_perm, min(P, round(n_perm * 15.0 / max(t1, 1e-3)))))
        tc = cpu_two_sample_step(orc, w, n_perm, threads)
        line["cpu_baseline"] = {"value": (n_perm / P) / tc, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n_perm} of the {P} permutations of the first region pair ({2 * n_perm} ten-start refits + summaries + CMD), {tc:.1f} s; "
                                          "a region pair costs P permutations, so pairs/s = (n_perm / P) / seconds"}
    return line


# ----------------------------------------------------------------------------------------------------------------------
# group 6 vs 6: cpn_grp_test_batch
# ----------------------------------------------------------------------------------------------------------------------
def group_workload(cpn, sub_lib, R, r0, r1, n1, n2, seed):
    """Regions [r0, r1) of the group workload (region g uses CpG layout g mod 137): per-CpG tables, sub-regions, and the fitted
    parameters of the S = n1+n2 samples: ϕ_s = ϕ_group + noise as in gen_grp_comp_data (simulations.jl:608,631); group 2 is
    shifted in 10 % of the regions."""
    lay = cpn.simulations.load_layouts()
    nlay = len(lay["cpg_off"]) - 1
    N = np.diff(lay["cpg_off"])

    class L:      # the 137 layouts as a NanoBatch-like object for sub_regions()
        R = nlay
        chrst, chrend, cpg_off, cpg_pos = lay["chrst"], lay["chrend"], lay["cpg_off"], lay["cpg_pos"]
    so, sp, sq = sub_regions(sub_lib, L, 350)
    which = np.arange(r0, r1) % nlay
    n = r1 - r0

    def tile(vals, off, shrink):
        # layouts repeat with period nlay: head (partial cycle), full cycles, tail
        parts = []
        g = r0
        while g < r1:
            a = g % nlay
            b = min(nlay, a + (r1 - g))
            if a == 0 and b == nlay and (r1 - g) >= 2 * nlay:
                reps = (r1 - g) // nlay
                parts.append(np.tile(vals[off[0] - 0:off[nlay] - nlay * shrink], reps))
                g += reps * nlay
                continue
            parts.append(vals[off[a] - a * shrink:off[b] - b * shrink])
            g += b - a
        return np.concatenate(parts)

    cpg_off = np.concatenate([[0], np.cumsum(N[which])]).astype(np.int64)
    sub_off = np.concatenate([[0], np.cumsum(np.diff(so)[which])]).astype(np.int64)
    rng = np.random.default_rng(seed)
    S = n1 + n2
    base = np.stack([rng.uniform(-1, 1, n), rng.uniform(-20, 20, n), rng.uniform(0, 5, n)], axis=1)
    shift = np.where(rng.random(n)[:, None] < 0.1, np.array([[0.5, 10.0, 1.0]]), 0.0)
    noise = np.stack([rng.uniform(-.1, .1, (n, S)), rng.uniform(-.5, .5, (n, S)), rng.uniform(0, .2, (n, S))], axis=2)
    phi = base[:, None, :] + noise
    phi[:, n1:, :] += shift[:, None, :]
    from itertools import combinations
    labs = np.array(list(combinations(range(1, S + 1), n1)), dtype=np.int32)
    return dict(phi=np.ascontiguousarray(phi), cpg_off=cpg_off, rho_n=tile(lay["rho_n"], lay["cpg_off"], 0), d_n=tile(lay["d_n"], lay["cpg_off"], 1),
                sub_off=sub_off, sub_p=tile(sp, so, 0), sub_q=tile(sq, so, 0), labelings=labs, S=S, n1=n1, n2=n2, N=N[which])


def group_config(args, world):
    return {"workload": f"group comparison 6 vs 6: {args.regions} regions IN TOTAL x 12 fitted samples, summaries of the 12 models + CMD of all 66 pairs + "
                        "all 924 labelings (unmatched); matched variant (64 sign patterns) timed beside it; real hg38 CpG layouts (N~63, 9 sub-regions)",
            "regions": args.regions, "regions_are": "total", "n1": 6, "n2": 6, "labelings": 924,
            "parallelism": f"regions sharded over {world} GPU(s), no data-path collective; results gathered on rank 0",
            "l2": "inputs (1.5 KB per region) and the on-device tables (≈ 25 KB per region) are far larger than the 126 MB L2"}


def bench_group(env):
    args, rank, local_rank, world, torch, dist, cpn, lib = (env[k] for k in ("args", "rank", "local_rank", "world", "torch", "dist", "cpn", "lib"))
    total = args.regions
    bounds = np.linspace(0, total, world + 1).astype(np.int64)          # regions cost the same to within the CpG count: equal blocks
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    R = r1 - r0
    w = group_workload(cpn, lib, total, r0, r1, 6, 6, cpn.simulations.MASTER_SEED + 7)
    S, n1, n2, labs = w["S"], w["n1"], w["n2"], w["labelings"]
    js = np.arange(2 ** n1, dtype=np.int64)
    sK = int(w["sub_off"][-1])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()      # noqa: E731
    hin = {k: pin(w[k]) for k in ("phi", "rho_n", "d_n", "sub_p", "sub_q")}
    # outputs of the unmatched test in one pinned buffer (gathered as a whole): T [3sK] f64, cnt [3sK] i64, p [3sK] f64
    out_buf = torch.empty(3 * 3 * sK * 8, dtype=torch.uint8).pin_memory()
    hT, hcnt, hp = (out_buf[i * 24 * sK:(i + 1) * 24 * sK].view(dt) for i, dt in enumerate((torch.float64, torch.int64, torch.float64)))
    mT, mcnt, mp, mtc = (torch.empty(n, dtype=dt).pin_memory() for n, dt in ((2 * sK, torch.float64), (2 * sK, torch.int64), (2 * sK, torch.float64),
                                                                              (sK, torch.float64)))
    h2d = int(sum(v.numel() * v.element_size() for v in hin.values())) + 16 * (R + 1) + labs.nbytes
    d2h = out_buf.numel()
    din = {k: v.cuda() for k, v in hin.items()}
    dT, dp = torch.empty(3 * sK, dtype=torch.float64, device="cuda"), torch.empty(3 * sK, dtype=torch.float64, device="cuda")
    dcnt = torch.empty(3 * sK, dtype=torch.int64, device="cuda")
    dtc = torch.empty(sK, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    fp64_peak = lib.measure_fp64_peak(local_rank)

    def call(matched, on_device):
        src = din if on_device else hin
        T, c, p = ((dT, dcnt, dp) if on_device else (hT, hcnt, hp)) if not matched else ((dT, dcnt, dp) if on_device else (mT, mcnt, mp))
        tc = dtc if on_device else mtc
        lib.grp_test_batch_raw(R, n1, n2, matched, src["phi"].data_ptr(), w["cpg_off"], src["rho_n"].data_ptr(), src["d_n"].data_ptr(), w["sub_off"],
                               src["sub_p"].data_ptr(), src["sub_q"].data_ptr(), js if matched else labs, len(js) if matched else len(labs), True,
                               T.data_ptr(), c.data_ptr(), p.data_ptr(), tc.data_ptr(), on_device, local_rank)
        return lib.last_device_ms(local_rank), lib.last_launches(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        call(False, True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.set_profile(True, local_rank)
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    for _ in range(args.steps):
        ms, ln = call(False, True)
        dev_ms += ms; launches += ln
    barrier()
    t_dev = time.perf_counter() - t0
    prof = lib.get_profile(local_rank)
    lib.set_profile(False, local_rank)
    clocks = sampler.summary()
    call(True, True)
    mat_ms = sum(call(True, True)[0] for _ in range(args.steps)) / args.steps
    del din
    torch.cuda.empty_cache()
    gather = cpn.sharding.ResultGather(out_buf.numel(), torch.device("cuda", local_rank)) if world > 1 else None

    def e2e_step():
        call(False, False)
        if gather is not None:
            gather.run(out_buf)

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    steps_ms = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        e2e_step()
        steps_ms.append(round((time.perf_counter() - t1) * 1e3, 2))
    barrier()
    t_e2e = time.perf_counter() - t0
    gb = gather.bytes_moved() if gather is not None else (0, 0)
    red = torch.tensor([t_dev, t_e2e, dev_ms, mat_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(h2d + gb[0]), float(d2h + gb[1]), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_dev, t_e2e, dev_ms, mat_ms = (float(x) for x in red.cpu())
    if rank != 0:
        return None
    config = group_config(args, world)
    config["shard_bounds"] = [int(b) for b in bounds]
    line = base_line(args, world, config, "strong")
    line["metric"] = "regions/sec tested (group comparison 6 vs 6: summaries of 12 models + 66 CMDs + 924-labeling sweep per region)"
    npair = S * (S - 1) // 2
    fl_cmd = C_CMD * npair * float(w["N"].sum()) * args.steps
    K = np.diff(w["sub_off"]).astype(np.float64)
    adds_sweep = float(np.sum(924.0 * K * (2 * S + n1 * n2))) * args.steps
    ms = {k: v[0] for k, v in prof.items()}
    dom = max(("cmd", "sweep", "statsum"), key=lambda k: ms.get(k, 0.0))
    if dom == "cmd":
        ach, unit, peak, note = fl_cmd / (ms["cmd"] * 1e-3) / 1e12, "TFLOP/s", fp64_peak, f"{C_CMD:.0f} flop per CpG of each of the 66 mixture chains"
    elif dom == "sweep":
        ach, unit, peak, note = adds_sweep / (ms["sweep"] * 1e-3) / 1e12, "TFLOP/s", fp64_peak / 2, ("(2S + n1·n2) = 60 correctly rounded FP64 adds per (labeling, "
                                                                                                      "sub-region), in the reference's order; peak = DADD rate = half the DFMA flop rate")
    else:
        ach, unit, peak, note = float(np.sum(C_S * w["N"] * S + C_G * K * w["N"] * S)) * args.steps / (ms["statsum"] * 1e-3) / 1e12, "TFLOP/s", fp64_peak, "60·N + 30·K·N flop per model"
    tr = ncu_traffic("k_" + dom, "group")
    line.update({"value": total * args.steps / t_dev, "ms_per_step": dev_ms / args.steps, "wall_ms_per_step": 1e3 * t_dev / args.steps,
                 "e2e": {"value": total * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
                         "ms_steps_rank0": steps_ms, "gather": "NCCL gather of (T, cnt, p) to rank 0 + device→host copy, inside the timed step" if gather else None},
                 "matched_variant": {"ms_per_step": mat_ms, "regions_per_s": total / (mat_ms * 1e-3) if mat_ms else None, "sign_patterns": 64},
                 "gpu_launches": int(tot[2].item()), "clocks": clocks,
                 "roofline": {"bound": "fp64", "kernel": {"cmd": "k_cmd_lane", "sweep": "k_unmat_sweep", "statsum": "k_stat_sums"}[dom], "achieved": ach, "peak": peak,
                              "unit": unit, "frac": ach / peak if peak else None, "traffic": tr["bytes_per_launch"] if tr else None,
                              "traffic_note": tr["note"] if tr else "no ncu capture committed for this workload", "flop_model": note,
                              "peak_source": "DFMA loop measured on this GPU in this run",
                              "per_class": {"cmd_tflops": fl_cmd / (ms["cmd"] * 1e-3) / 1e12 if ms.get("cmd") else None,
                                            "sweep_tadds": adds_sweep / (ms["sweep"] * 1e-3) / 1e12 if ms.get("sweep") else None}},
                 "kernel_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in prof.items() if v[1]},
                 "test_stats": {"frac_p_lt_0.05": float(np.nanmean(hp.numpy() < 0.05))}})
    if world == 1 and not args.no_cpu:
        import oracle_py as orc
        threads = os.cpu_count() or 1
        t1 = cpu_group_step(orc, w, min(R, threads), threads)
        n_s = int(min(R, max(8 * threads, threads * round(12.0 / max(t1, 1e-3)))))
        tc = cpu_group_step(orc, w, n_s, threads)
        line["cpu_baseline"] = {"value": n_s / tc, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"first {n_s} regions: get_stat_sums! x 12, comp_cmd x 66, unmat_est_reg_test with 924 labelings per region, {tc:.1f} s"}
    return line


# ======================================================================================================================
# reference arm: the reference's own algorithm (oracle) on the host cores, rank 0 only, no product library mapped
# ======================================================================================================================
def reference_arm(args):
    import importlib.util
    import oracle_py as orc

    def load(name):      # the host-side workload generator of the package, loaded file by file: the package __init__ would map the CUDA library
        spec = importlib.util.spec_from_file_location(f"_cpn_{name}", os.path.join(ROOT, "cpelnano.jl_b200", f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m

    class Cpn:
        simulations = load("simulations")
        sharding = load("sharding")
    cpn = Cpn
    threads = os.cpu_count() or 1
    wl = args.workload
    world = args.gpus
    n_init, max_iters = args.n_init, args.max_em_iters
    if wl in ("chr22", "targeted", "whole-genome"):
        config = fit_config(args, world)
        if wl == "targeted":
            lay, phi_true = targeted_layouts()
            nb, arrays = build_workload(cpn, orc, len(phi_true), args.reads, n_init, cpn.simulations.MASTER_SEED, subreg=351, layouts=lay, phi_true=phi_true)
            n_s = nb.R
        else:
            n_s = args.cpu_sample or CPU_REGIONS_PER_THREAD * threads
            nb, arrays = build_workload(cpn, orc, max(n_s, 64), args.reads, n_init, cpn.simulations.MASTER_SEED)
        idx = np.arange(n_s)
        run = lambda n: cpu_fit_step(orc, nb, arrays, idx[:n], n_init, max_iters, threads)[0]      # noqa: E731
        unit_per_step, sample = n_s, f"{n_s} regions of the same workload per step, one region per task on {threads} threads"
        metric = METRIC
    elif wl == "two-sample":
        config = two_sample_config(args, world)
        w = two_sample_workload(cpn, orc, 4, args.reads, n_init, max_iters, cpn.simulations.MASTER_SEED)
        n_perm = args.cpu_sample or max(8, 2 * threads)
        run = lambda n: cpu_two_sample_step(orc, w, max(2, n), threads)      # noqa: E731
        n_s, unit_per_step = n_perm, n_perm / args.perms
        sample = f"{n_perm} of the {args.perms} permutations of one region pair per step ({2 * n_perm} ten-start refits + summaries + CMD)"
        metric = "region pairs/sec tested (two-sample permutation test: 2·P multi-start EM refits + MML/NME/CMD + p-value counts per region)"
    else:
        config = group_config(args, world)
        n_s = args.cpu_sample or 16 * threads
        w = group_workload(cpn, orc, n_s, 0, n_s, 6, 6, cpn.simulations.MASTER_SEED + 7)
        run = lambda n: cpu_group_step(orc, w, n, threads)      # noqa: E731
        unit_per_step, sample = n_s, f"{n_s} regions per step: get_stat_sums! x 12, comp_cmd x 66, unmat_est_reg_test with 924 labelings"
        metric = "regions/sec tested (group comparison 6 vs 6: summaries of 12 models + 66 CMDs + 924-labeling sweep per region)"
    for _ in range(min(args.warmup, 1)):
        run(min(n_s, threads))
    t = 0.0
    for _ in range(args.steps):
        t += run(n_s)
    v = unit_per_step * args.steps / t
    line = {"impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak" if wl in ("chr22", "two-sample") else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "mstep": {"mode": "fd: finite differences of finite differences (the reference's NLsolve path)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": sample + " (oracle = C++ restatement of the reference algorithm; Julia absent)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
