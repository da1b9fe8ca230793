set -x
python -m pytest tests/test_gpu_gather.py tests/test_gpu_hostgather.py -x -q -m gpu > gpurun_out/r2_f2_tests.log 2>&1; tail -3 gpurun_out/r2_f2_tests.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 50 > gpurun_out/r2_f2_cfg5_n2.log 2>&1
$T --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 50 --no-host-direct > gpurun_out/r2_f2_cfg5_n2_nohostdirect.log 2>&1
$T --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --workload cfg4 --steps 1000 --warmup 100 > gpurun_out/r2_f2_cfg4_n2.log 2>&1
python scripts/summarize.py gpurun_out/r2_f2_*.log
grep -h "Error\|error\|Traceback\|PARITY" gpurun_out/r2_f2_c*.log | head
