#!/bin/bash
# ncu evidence for profiles/ (round 2, second half: mac_tma_kernel with the CTA-head / prefetching producer,
# mac_dyn_tma_kernel, inv_dyn_cluster_kernel, ctl_upload_kernel, the non-uniform renderer's launches).
#   bash profiles/ncu_capture_r2b.sh        (under gpurun on ONE B200)
# Every ncu run is preceded by the same command without ncu, which must exit 0.  Sharded runs are captured
# on one GPU rendering the per-GPU shard shape (--outputs M/G: same plan, grid and bytes as every rank).
set -u
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
DEADLINE=${NCU_DEADLINE_S:-840}     # captures that would start later than this are skipped (GPU budget)
late() { local now=$(date +%s); [ $((now - T0)) -gt $DEADLINE ]; }
# NCU_ONLY="tag tag ...": capture only these (tags of run_pair; the launch lists are "launches", "launches_cfg3", "launches_nu")
want() { [ -z "${NCU_ONLY:-}" ] || [[ " $NCU_ONLY " == *" $1 "* ]]; }
BASE="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-sustained"
FULL="--set full --clock-control none --import-source on"
DYN="mac_dyn_tma_kernel|inv_dyn_cluster_kernel|fwd_fft_overlap_kernel|ctl_upload_kernel"

run_pair() {   # tag, kernel regex, skip, count, command...
  local tag=$1 regex=$2 skip=$3 count=$4; shift 4
  if ! want $tag; then return; fi
  if late; then echo "$tag: skipped (deadline)"; return; fi
  "$@" > $OUT/plain_$tag.log 2>&1 || { echo "$tag: plain run failed"; tail -3 $OUT/plain_$tag.log; return; }
  ncu $FULL -k regex:"$regex" -s $skip -c $count -o $OUT/r2b_$tag "$@" > $OUT/ncu_$tag.log 2>&1
  echo "$tag rc=$? t=$(( $(date +%s) - T0 ))s"
}

# launch lists (all kernels, serialised, cold caches): default bench command, cfg3, non-uniform renderer
if want launches; then $BASE > $OUT/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/r2b_launches_cfg5_1gpu.csv $BASE > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"; fi
C3="python bench.py --workload cfg3 --steps 6 --warmup 50 --no-cpu-baseline --no-parity --no-sustained"
run_pair mac_cfg5_gpus1 "mac_tma_kernel" 2 2 $BASE
run_pair cfg3 "$DYN" 200 8 $C3
run_pair mac_cfg5_gpus8 "mac_tma_kernel" 2 2 $BASE --outputs 64
run_pair mac_cfg4_gpus1 "mac_tma_kernel" 2 2 $BASE --workload cfg4
run_pair cfg2 "$DYN" 200 8 python bench.py --workload cfg2 --steps 6 --warmup 50 --no-cpu-baseline --no-parity --no-sustained
if want launches_cfg3 && ! late; then $C3 > $OUT/plain_launches_cfg3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file $OUT/r2b_launches_cfg3_1gpu.csv $C3 > $OUT/ncu_launches_cfg3.log 2>&1; fi
echo "launch list cfg3 rc=$?"
NU="$BASE --outputs 64 --nonuniform 1x4,2x2,4x10"
if want launches_nu && ! late; then $NU > $OUT/plain_launches_nu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file $OUT/r2b_launches_nu_shard64.csv $NU > $OUT/ncu_launches_nu.log 2>&1; fi
echo "launch list nu rc=$?"

run_pair fft_cfg5 "fwd_fft_overlap_kernel|inv_fft_kernel" 4 4 $BASE --outputs 64
run_pair mac_cfg5_gpus2 "mac_tma_kernel" 2 2 $BASE --outputs 256
run_pair mac_cfg5_gpus4 "mac_tma_kernel" 2 2 $BASE --outputs 128
run_pair mac_cfg4_gpus8 "mac_tma_kernel" 2 2 $BASE --workload cfg4 --outputs 8
for rep in $OUT/r2b_*.ncu-rep; do
  ncu -i $rep --page raw --csv > ${rep%.ncu-rep}.raw.csv 2> /dev/null
done
python profiles/ncu_summary.py $OUT/r2b_*.ncu-rep > $OUT/r2b_ncu_kernels_raw.md 2>&1
[ -f $OUT/r2b_cfg3.ncu-rep ] && ncu -i $OUT/r2b_cfg3.ncu-rep --page source --csv > $OUT/r2b_cfg3.source.csv 2> /dev/null
[ -f $OUT/r2b_mac_cfg5_gpus8.ncu-rep ] && ncu -i $OUT/r2b_mac_cfg5_gpus8.ncu-rep --page source --csv > $OUT/r2b_mac_cfg5_gpus8.source.csv 2> /dev/null
mkdir -p $OUT/keep; mv $OUT/r2b_cfg3.ncu-rep $OUT/r2b_mac_cfg5_gpus8.ncu-rep $OUT/keep/ 2> /dev/null; rm -f $OUT/r2b_*.ncu-rep; mv $OUT/keep/* $OUT/ 2> /dev/null; rmdir $OUT/keep
ls -la $OUT | tail -40
du -sh $OUT
