# eight GPUs, short: cfg5 and cfg4 bench lines with the block overlap (parity inside every line)
O=gpurun_out/r2c_c; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
date +%s > $O/t0
timeout 90 $T --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 600 --warmup 100 --no-sustained > $O/cfg5_n8.log 2>&1; echo "cfg5 n8 rc=$? $(( $(date +%s) - $(cat $O/t0) ))s"
timeout 70 $T --nproc-per-node 8 --master-port 29622 bench.py --gpus 8 --workload cfg4 --steps 2000 --warmup 200 --no-sustained > $O/cfg4_n8.log 2>&1; echo "cfg4 n8 rc=$? $(( $(date +%s) - $(cat $O/t0) ))s"
for f in $O/cfg*_n8.log; do grep -h "^{" $f > ${f%.log}.json; done
python profiles/summarize.py $O/*.json | cut -c1-220
grep -h "Error\|error\|Traceback\|PARITY" $O/*.log | head -5
