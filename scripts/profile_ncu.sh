#!/bin/bash
# ncu captures for profiles/ (run under gpurun, 1 GPU). $1 = tag (e.g. r01)
TAG=${1:-r01}
mkdir -p gpurun_out
# (1) every launch of bench steps with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${TAG}.stdout 2>&1
# (2) full-set capture of the library's kernels on one layer fwd+bwd (second iteration = warm)
ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|attn_|ln_bwd|ln_fwd|colsum' -s 22 -c 22 \
    -o gpurun_out/prof_layer_${TAG} python scripts/run_one_layer.py 2 256 > gpurun_out/prof_layer_${TAG}.stdout 2>&1
ls -la gpurun_out/
