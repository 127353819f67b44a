#!/bin/bash
# Compile-time ablation of k_tc's epilogue (timing only; results are wrong for TC_ABL != 0).
# Builds one library per switch and times bench.py's k_tc with each (run on the GPU box):
#   TC_ABL bits: 1 no agg store, 2 no pair barrier / exchange, 4 no ypart store,
#                8 no phase-C store + per-block-sum loads, 16 no mode store + phase-A loads
#                (also lets the compiler drop 7/8 of the digit recombination)
set -e
cd "$(dirname "$0")/.."
mkdir -p microbench/abl
for v in 0 1 2 4 8 16 31; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
       -DTC_ABL=$v -o microbench/abl/lib_abl$v.so sdrterm_b200/csrc/sdrb_api.cu
  echo -n "TC_ABL=$v k_tc ms: "
  SDRB_LIB=$PWD/microbench/abl/lib_abl$v.so python bench.py --steps 10 --warmup 3 --e2e-chunks 0 \
      --simo-chunks 0 --no-cpu 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.read())['roofline']['kernel_ms']['k_tc'])"
done
