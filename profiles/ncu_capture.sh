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

$CFG5 > $OUT/plain_cfg5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_${TAG}.csv $CFG5 > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"

$CFG5 > $OUT/plain_cfg5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mac_kernel -s 2 -c 2 \
    -o $OUT/mac_${TAG} $CFG5 > $OUT/ncu_mac.log 2>&1
echo "mac full rc=$?"

# block-loop FFT kernels: skip the 128 filter-load launches of fwd_fft_kernel
$CFG5 > $OUT/plain_cfg5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwd_fft_kernel -s 130 -c 2 \
    -o $OUT/fwd_${TAG} $CFG5 > $OUT/ncu_fwd.log 2>&1
echo "fwd full rc=$?"

$CFG5 > $OUT/plain_cfg5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:inv_fft_kernel -s 2 -c 2 \
    -o $OUT/inv_${TAG} $CFG5 > $OUT/ncu_inv.log 2>&1
echo "inv full rc=$?"

# filter-load FFT (24064 transforms per launch)
$CFG5 > $OUT/plain_cfg5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwd_fft_kernel -s 10 -c 1 \
    -o $OUT/fwdload_${TAG} $CFG5 > $OUT/ncu_fwdload.log 2>&1
echo "fwd load full rc=$?"

# BRS, head rotating every block (MT = 2 kernel, old + new filter tables)
$CFG3 > $OUT/plain_cfg3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mac_kernel -s 52 -c 2 \
    -o $OUT/mac_cfg3_${TAG} $CFG3 > $OUT/ncu_mac_cfg3.log 2>&1
echo "mac cfg3 full rc=$?"
ls -la $OUT | tail -20
