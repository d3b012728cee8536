#!/usr/bin/env python3
"""Sweep the chunk-pipeline knobs of the host-pointer path on one workload, in one process:
CPN_CHUNKS x CPN_CHUNK_RATIO x CPN_WORKERS, median wall time of cpn_analyze_batch with pinned host buffers.
Usage: python tools/e2e_sweep.py [--regions 100000] [--reps 5] > gpurun_out/e2e_sweep.json
"""
import argparse
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=100_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--chunks", default="4,6,8,12")
    ap.add_argument("--ratios", default="1.0,1.3,1.6")
    ap.add_argument("--workers", default="2,3")
    ap.add_argument("--two-tables", action="store_true", help="cpn_analyze_batch instead of the packed entry bench.py's e2e leg uses")
    args = ap.parse_args()
    import torch
    import cpelnano_b200 as cpn
    lib = cpn.Library()
    lib.ctx(0)
    R = args.regions
    nb, arrays = bench.build_workload(cpn, lib, R, 30, 10, cpn.simulations.MASTER_SEED)
    shapes = bench.out_shapes(R, arrays)
    host_in = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in arrays.items()}
    host_out = {k: torch.empty(n, dtype=torch.float64 if dt == np.float64 else (torch.int32 if dt == np.int32 else torch.int64)).pin_memory()
                for k, (n, dt) in shapes.items()}
    pin = {k: v.data_ptr() for k, v in host_in.items()}
    packed = not args.two_tables
    if packed:
        host_llr = torch.empty(host_in["lu"].numel(), dtype=torch.float64).pin_memory()
        host_base = torch.empty(int(arrays["read_off"][-1]), dtype=torch.float64).pin_memory()
        lib.pack_emissions(arrays["grp_off"], arrays["read_off"], host_in["lu"].numpy(), host_in["lm"].numpy(), out=(host_llr.numpy(), host_base.numpy()))
        pin = dict(pin, lu=host_llr.data_ptr(), lm=host_base.data_ptr())
    pout = {k: v.data_ptr() for k, v in host_out.items()}

    def call():
        torch.cuda.synchronize()
        t = time.perf_counter()
        bench.run_step(lib, R, pin, pout, 10, 20, cpn.MSTEP_ANALYTIC, False, 0, packed=packed)
        return (time.perf_counter() - t) * 1e3

    ref = None
    rows = []
    combos = [(None, None, None)] + list(itertools.product([int(c) for c in args.chunks.split(",")], [float(r) for r in args.ratios.split(",")],
                                                           [int(w) for w in args.workers.split(",")]))
    for (c, r, w) in combos:
        for k, v in (("CPN_CHUNKS", c), ("CPN_CHUNK_RATIO", r), ("CPN_WORKERS", w)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        for _ in range(2):
            call()
        ts = [call() for _ in range(args.reps)]
        phi = host_out["phi_hat"].numpy().copy()
        if ref is None:
            ref = phi
        same = bool(np.array_equal(phi, ref, equal_nan=True))
        rows.append({"chunks": c, "ratio": r, "workers": w, "median_ms": float(np.median(ts)), "min_ms": float(min(ts)), "identical": same})
        print(json.dumps(rows[-1]), flush=True)
    best = min(rows, key=lambda x: x["median_ms"])
    print(json.dumps({"best": best, "default": rows[0]}))


if __name__ == "__main__":
    main()
