#!/bin/bash
# ncu evidence for profiles/ (run under gpurun on ONE B200):
#   bash profiles/ncu_capture.sh <tag>
# Every ncu run is preceded by the same command without ncu, which must exit 0.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CFG5="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
CFG3="python bench.py --workload cfg3 --steps 6 --warmup 50 --no-cpu-baseline"
CFG1="ssr-pre-0.4.0_b200/lib/perf_static_convolver 50"
FULL="--set full --clock-control none --import-source on"

$CFG5 > $OUT/plain_cfg5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_${TAG}.csv $CFG5 > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"

$CFG5 > $OUT/plain_cfg5.log 2>&1 && \
ncu $FULL -k regex:"mac_(tma_)?kernel" -s 2 -c 2 -o $OUT/mac_${TAG} $CFG5 > $OUT/ncu_mac.log 2>&1
echo "mac full rc=$?"

# block-loop FFT kernels (filter loading uses the batch kernels, not these)
$CFG5 > $OUT/plain_cfg5.log 2>&1 && \
ncu $FULL -k regex:fwd_fft_kernel -s 2 -c 2 -o $OUT/fwd_${TAG} $CFG5 > $OUT/ncu_fwd.log 2>&1
echo "fwd full rc=$?"

$CFG5 > $OUT/plain_cfg5.log 2>&1 && \
ncu $FULL -k regex:inv_fft_kernel -s 2 -c 2 -o $OUT/inv_${TAG} $CFG5 > $OUT/ncu_inv.log 2>&1
echo "inv full rc=$?"

# BRS, head rotating every block: generated-terms MAC (old + new filter state), the
# partial reduction and the three-group inverse kernel
$CFG3 > $OUT/plain_cfg3.log 2>&1 && \
ncu $FULL -k regex:"mac_kernel|dyn_reduce_kernel|inv_fft_kernel" -s 150 -c 3 \
    -o $OUT/cfg3_${TAG} $CFG3 > $OUT/ncu_cfg3.log 2>&1
echo "cfg3 full rc=$?"

# Level 1 fused kernel (cfg1: StaticConvolver, 8 partitions)
$CFG1 > $OUT/plain_cfg1.log 2>&1 && \
ncu $FULL -k regex:level1_fused_kernel -s 20 -c 2 -o $OUT/level1_${TAG} $CFG1 > $OUT/ncu_level1.log 2>&1
echo "level1 full rc=$?"
ls -la $OUT | tail -20
