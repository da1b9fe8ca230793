# two GPUs, final tree: the multi-GPU tests and one sharded cfg5 line
O=gpurun_out/r2c_e; mkdir -p $O
python -m pytest tests/test_gpu_gather.py tests/test_gpu_hostgather.py tests/test_gpu_fullsize.py -x -q -m gpu \
  -k "gather or hostgather or cfg4" > $O/pytest_2gpu.log 2>&1; tail -2 $O/pytest_2gpu.log
timeout 120 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29631 bench.py --gpus 2 --steps 20 --warmup 5 > $O/cfg5_n2.log 2>&1; echo "cfg5 n2 rc=$?"
grep -h "^{" $O/cfg5_n2.log > $O/cfg5_n2_driver_style.json; python profiles/summarize.py $O/*.json | cut -c1-200
