#!/bin/bash
# ncu evidence for profiles/ (round 2; run under gpurun on ONE B200):
#   bash profiles/ncu_capture_r2.sh
# Every ncu run is preceded by the same command without ncu, which must exit 0.  The
# sharded runs' MAC launches are captured on one GPU rendering the per-GPU shard shape
# (--outputs M/G: the same plan, grid and bytes as each rank of the N-GPU run).
set -u
OUT=gpurun_out
mkdir -p $OUT
BASE="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
FULL="--set full --clock-control none --import-source on"

run_pair() {   # tag, kernel regex, skip, count, command...
  local tag=$1 regex=$2 skip=$3 count=$4; shift 4
  "$@" > $OUT/plain_$tag.log 2>&1 || { echo "$tag: plain run failed"; tail -3 $OUT/plain_$tag.log; return; }
  ncu $FULL -k regex:"$regex" -s $skip -c $count -o $OUT/r2_$tag "$@" > $OUT/ncu_$tag.log 2>&1
  echo "$tag rc=$?"
}

# launch list of the default bench command (all kernels, serialised, cold caches)
$BASE > $OUT/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/r2_launches_cfg5_1gpu.csv $BASE > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"

run_pair mac_cfg5_gpus1 "mac_tma_kernel" 2 2 $BASE
run_pair mac_cfg5_gpus2 "mac_tma_kernel" 2 2 $BASE --outputs 256
run_pair mac_cfg5_gpus4 "mac_tma_kernel" 2 2 $BASE --outputs 128
run_pair mac_cfg5_gpus8 "mac_tma_kernel" 2 2 $BASE --outputs 64
run_pair mac_cfg4_gpus1 "mac_tma_kernel" 2 2 $BASE --workload cfg4
run_pair mac_cfg4_gpus8 "mac_tma_kernel" 2 2 $BASE --workload cfg4 --outputs 8
run_pair fft_cfg5 "fwd_fft_overlap_kernel|inv_fft_kernel" 4 4 $BASE --outputs 64
run_pair cfg3 "mac_kernel|dyn_reduce_kernel|inv_fft_kernel|fwd_fft_overlap_kernel" 160 4 python bench.py --workload cfg3 --steps 6 --warmup 50 --no-cpu-baseline --no-parity
run_pair cfg2 "mac_kernel|dyn_reduce_kernel|inv_fft_kernel|fwd_fft_overlap_kernel" 160 4 python bench.py --workload cfg2 --steps 6 --warmup 50 --no-cpu-baseline --no-parity
# summaries are made here (gpurun brings back at most 64 MiB): raw CSV pages + the markdown
# table of profiles/ncu_summary.py; only the two headline reports travel as .ncu-rep
for rep in $OUT/r2_*.ncu-rep; do
  ncu -i $rep --page raw --csv > ${rep%.ncu-rep}.raw.csv 2> /dev/null
done
python profiles/ncu_summary.py $OUT/r2_*.ncu-rep > $OUT/r2_ncu_kernels.md 2>&1
ncu -i $OUT/r2_mac_cfg5_gpus8.ncu-rep --page source --csv > $OUT/r2_mac_cfg5_gpus8.source.csv 2> /dev/null
mkdir -p $OUT/keep && mv $OUT/r2_mac_cfg5_gpus1.ncu-rep $OUT/r2_cfg3.ncu-rep $OUT/keep/ && rm -f $OUT/r2_*.ncu-rep && mv $OUT/keep/* $OUT/ && rmdir $OUT/keep
ls -la $OUT | tail -30
du -sh $OUT
