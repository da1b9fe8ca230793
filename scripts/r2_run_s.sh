# two GPUs: the multi-GPU tests and the sharded bench lines with the round's final kernels
set -x
O=gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_gather.py tests/test_gpu_hostgather.py tests/test_gpu_fullsize.py -x -q -m gpu \
  -k "gather or hostgather or cfg4" > $O/r2_s_tests.log 2>&1; tail -3 $O/r2_s_tests.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
S="--steps 1000 --warmup 100"
$T --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 $S > $O/r2_s_cfg5_n2.log 2>&1
$T --nproc-per-node 2 --master-port 29604 bench.py --gpus 2 $S --nonuniform 1x4,2x2,4x10 > $O/r2_s_nu3_cfg5_n2.log 2>&1
$T --nproc-per-node 2 --master-port 29605 bench.py --gpus 2 --workload cfg4 --steps 2000 --warmup 200 > $O/r2_s_cfg4_n2.log 2>&1
python scripts/summarize.py $O/r2_s_cfg*.log
grep -h "^{" $O/r2_s_nu3_cfg5_n2.log | cut -c1-400
grep -h "Error\|error\|Traceback\|PARITY" $O/r2_s_*.log | head
