#!/bin/bash
# Builder-side GPU session (run under gpurun): stages selected by name, every stage logs into gpurun_out/<tag>_*.
#   tools/gpu_round.sh <tag> stage [stage ...]
# stages: tests bench ref group twosample targeted wg launches ncu_fit ncu_group ncu_two
# A number printed by a run under ncu is never a bench value; the ncu stages only run after the plain command exited 0.
set -u
TAG=$1; shift
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
for st in "$@"; do
  t0=$(date +%s)
  case $st in
    tests)     timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "tests rc=$?" ;;
    smoke)     timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    bench)     timeout 900 python bench.py > $O/${TAG}_bench_chr22.json 2> $O/${TAG}_bench_chr22.err; echo "bench rc=$?" ;;
    benchq)    timeout 600 python bench.py --no-cpu > $O/${TAG}_bench_chr22_nocpu.json 2> $O/${TAG}_bench_chr22_nocpu.err; echo "benchq rc=$?" ;;
    ref)       timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "ref rc=$?" ;;
    group)     timeout 900 python bench.py --workload group > $O/${TAG}_bench_group.json 2> $O/${TAG}_bench_group.err; echo "group rc=$?" ;;
    twosample) timeout 1500 python bench.py --workload two-sample --steps 2 --warmup 1 > $O/${TAG}_bench_two_sample.json 2> $O/${TAG}_bench_two_sample.err; echo "twosample rc=$?" ;;
    targeted)  timeout 600 python bench.py --workload targeted > $O/${TAG}_bench_targeted.json 2> $O/${TAG}_bench_targeted.err; echo "targeted rc=$?" ;;
    wg)        timeout 1500 python bench.py --workload whole-genome --steps 2 --warmup 1 --no-cpu > $O/${TAG}_bench_whole_genome.json 2> $O/${TAG}_bench_whole_genome.err; echo "wg rc=$?" ;;
    launches)  timeout 900 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/${TAG}_launches_chr22.csv \
                 python bench.py --no-cpu --steps 1 --warmup 1 > $O/${TAG}_launches_chr22.out 2>&1; echo "launches rc=$?" ;;
    launches_group) timeout 900 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/${TAG}_launches_group.csv \
                 python bench.py --workload group --regions 200000 --no-cpu --steps 1 --warmup 1 > $O/${TAG}_launches_group.out 2>&1; echo "launches_group rc=$?" ;;
    ncu_fit)   # first full-size E-step / M-step launch of a 100 000-region device-resident step
               timeout 900 $NCU --set full --import-source on -k regex:'k_estep|k_mstep_ws' -c 2 -o $O/${TAG}_ncu_fit -f \
                 python bench.py --no-cpu --steps 1 --warmup 1 > $O/${TAG}_ncu_fit.out 2>&1; echo "ncu_fit rc=$?" ;;
    ncu_fit_mid) # the 4th EM iteration (active list already compacted once or twice)
               timeout 900 $NCU --set full --import-source on -k regex:'k_estep|k_mstep_ws' -s 6 -c 2 -o $O/${TAG}_ncu_fit_mid -f \
                 python bench.py --no-cpu --steps 1 --warmup 1 > $O/${TAG}_ncu_fit_mid.out 2>&1; echo "ncu_fit_mid rc=$?" ;;
    ncu_group) timeout 900 $NCU --set full --import-source on -k regex:'k_stat_sums|k_cmd|k_unmat_sweep|k_mat_sweep' -c 4 -o $O/${TAG}_ncu_group -f \
                 python bench.py --workload group --regions 200000 --no-cpu --steps 1 --warmup 1 > $O/${TAG}_ncu_group.out 2>&1; echo "ncu_group rc=$?" ;;
    ncu_two)   timeout 900 $NCU --set full --import-source on -k regex:'k_two_samp_pvals|k_two_samp_views|k_score' -c 3 -o $O/${TAG}_ncu_two -f \
                 python bench.py --workload two-sample --regions 64 --no-cpu --steps 1 --warmup 1 > $O/${TAG}_ncu_two.out 2>&1; echo "ncu_two rc=$?" ;;
    parity)    timeout 180 python -m pytest tests/test_gpu_parity.py tests/test_pipeline_gpu.py -m gpu -x -q > $O/${TAG}_parity.log 2>&1; echo "parity rc=$?"; tail -3 $O/${TAG}_parity.log ;;
    ab)        bash tools/ab.sh 100000 $AB_LIBS > $O/${TAG}_ab.txt 2>&1; cat $O/${TAG}_ab.txt ;;
    trace)     CPN_TRACE=1 timeout 600 python bench.py --no-cpu --steps 1 --warmup 2 > $O/${TAG}_trace.json 2> $O/${TAG}_trace.err; echo "trace rc=$?" ;;
    multi)     # NG GPUs, the driver's launch line; one JSON line per workload
               TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
               for wl in chr22 whole-genome group two-sample; do
                 extra=""; [ "$wl" = "two-sample" ] && extra="--steps 2 --warmup 1"; [ "$wl" = "whole-genome" ] && extra="--steps 2 --warmup 1"
                 timeout 900 $TR bench.py --gpus $NG --workload $wl --no-cpu $extra > $O/${TAG}_bench_${wl}_${NG}gpu.json 2> $O/${TAG}_bench_${wl}_${NG}gpu.err; echo "multi $wl rc=$?"
               done ;;
    tests2)    timeout 900 python -m pytest tests -m gpu -x -q -k "device or multi or second" > $O/${TAG}_pytest_2gpu.log 2>&1; echo "tests2 rc=$?"; tail -3 $O/${TAG}_pytest_2gpu.log ;;
    *) echo "unknown stage $st" ;;
  esac
  echo "  stage $st: $(( $(date +%s) - t0 )) s"
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
