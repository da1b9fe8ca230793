# One 8-GPU lease: parity-carrying bench lines for cfg5 and cfg4 at N = 1/2/4/8, the gather /
# host-gather / sharded-cfg4 GPU tests on 8 real devices, and the A/Bs of the exchange step.
set -x
O=gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_gather.py tests/test_gpu_hostgather.py tests/test_gpu_fullsize.py -x -q -m gpu \
  -k "gather or hostgather or cfg4" > $O/r2_8gpu_tests.log 2>&1; tail -3 $O/r2_8gpu_tests.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
S="--steps 1000 --warmup 100"
$T --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 $S > $O/r2_8gpu_cfg5_n8.log 2>&1
$T --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 $S --no-host-direct > $O/r2_8gpu_cfg5_n8_nohostdirect.log 2>&1
# N = 2 and N = 4 side by side on disjoint GPUs (HBM-bound per GPU; the ranks of one run share nothing with the other)
CUDA_VISIBLE_DEVICES=0,1 $T --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 $S > $O/r2_8gpu_cfg5_n2.log 2>&1 &
CUDA_VISIBLE_DEVICES=2,3,4,5 $T --nproc-per-node 4 --master-port 29604 bench.py --gpus 4 $S > $O/r2_8gpu_cfg5_n4.log 2>&1 &
CUDA_VISIBLE_DEVICES=6 python bench.py --gpus 1 $S --no-cpu-baseline > $O/r2_8gpu_cfg5_n1.log 2>&1 &
wait
C="--workload cfg4 --steps 2000 --warmup 200"
$T --nproc-per-node 8 --master-port 29605 bench.py --gpus 8 $C > $O/r2_8gpu_cfg4_n8.log 2>&1
$T --nproc-per-node 4 --master-port 29606 bench.py --gpus 4 $C > $O/r2_8gpu_cfg4_n4.log 2>&1
$T --nproc-per-node 2 --master-port 29607 bench.py --gpus 2 $C > $O/r2_8gpu_cfg4_n2.log 2>&1
python bench.py --gpus 1 $C --no-cpu-baseline > $O/r2_8gpu_cfg4_n1.log 2>&1
$T --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 $S --gather nccl > $O/r2_8gpu_cfg5_n8_nccl.log 2>&1
SSR_B200_FWD_OVERLAP=0 $T --nproc-per-node 8 --master-port 29609 bench.py --gpus 8 $S > $O/r2_8gpu_cfg5_n8_nooverlap.log 2>&1
python scripts/summarize.py $O/r2_8gpu_cfg*.log
grep -h "Error\|error\|Traceback\|PARITY" $O/r2_8gpu_cfg*.log | head
