#!/bin/bash
# Kernel-tuning build: tools/build_variant.sh <name> "<extra nvcc defines>"  →  build/variants/<name>.so (select it with CPN_LIB).
# Only cpn_em.cu is recompiled with the defines; the other objects are the in-tree ones (run `make -C cpelnano.jl_b200/csrc` first).
set -e
NAME=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/cpelnano.jl_b200/csrc
mkdir -p $ROOT/build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $EXTRA -Xptxas -v -c $C/cpn_em.cu -o $ROOT/build/variants/${NAME}_em.o 2>&1 | grep -A2 "k_estepE\|k_estep_pairI\|k_estep_pair_wsE\|k_mstep_wsE" | grep "Used\|spill" || true
/usr/local/cuda/bin/nvcc $ARCH -shared -o $ROOT/build/variants/$NAME.so $ROOT/build/variants/${NAME}_em.o $C/cpn_ctx.o $C/cpn_summaries.o $C/cpn_sweeps.o $C/cpn_calls.o $C/cpn_genome.o $C/cpn_text.o $C/cpn_bam.o -lcudart -lz -lpthread
rm -f $ROOT/build/variants/${NAME}_em.o
echo "built build/variants/$NAME.so"
