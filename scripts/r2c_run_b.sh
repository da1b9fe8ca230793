# two GPUs: the multi-GPU tests, then sharded bench lines with and without the block overlap (SSR_B200_BLOCK_OVERLAP)
O=gpurun_out/r2c_b; mkdir -p $O
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_gather.py tests/test_gpu_hostgather.py tests/test_gpu_fullsize.py -x -q -m gpu \
  -k "gather or hostgather or cfg4" > $O/pytest_gather_hostgather_cfg4_on_2_gpus.log 2>&1; tail -3 $O/pytest_gather_hostgather_cfg4_on_2_gpus.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29610
for ov in 1 0; do
  export SSR_B200_BLOCK_OVERLAP=$ov
  $T --nproc-per-node 2 --master-port $((port++)) bench.py --gpus 2 --steps 600 --warmup 100 --no-sustained > $O/cfg5_n2_ov$ov.log 2>&1; echo "cfg5 n2 ov=$ov rc=$?"
  $T --nproc-per-node 2 --master-port $((port++)) bench.py --gpus 2 --workload cfg4 --steps 2000 --warmup 200 --no-sustained > $O/cfg4_n2_ov$ov.log 2>&1; echo "cfg4 n2 ov=$ov rc=$?"
done
for f in $O/cfg*_n2_ov*.log; do grep -h "^{" $f > ${f%.log}.json; done
python profiles/summarize.py $O/*.json | cut -c1-220
grep -h "Error\|error\|Traceback\|PARITY" $O/*.log | head
