# round 2, call 7 (4 GPUs): SAM writer (pipelined), result fetch timing, then host phase times of the one-process 4-GPU call
O=gpurun_out/r2g; mkdir -p $O
python -m pytest tests/test_gpu_sam.py tests/test_cpp_host_mirror.py -m gpu -x -q > $O/pytest_sam.log 2>&1; tail -3 $O/pytest_sam.log
python tools/sam_throughput.py 20 > $O/sam_throughput.json 2> $O/sam.err; cat $O/sam_throughput.json
PCS_TIMING=1 python bench.py --no-cpu-baseline --steps 3 > $O/bench_n1.json 2> $O/bench_n1.err
grep "result fetch\|simulate_result" $O/bench_n1.err | tail -8
PCS_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err
grep "pcs host" $O/bench_n4.err | grep -v "plan:" | tail -24
python -c "import json; d=json.load(open('$O/bench_n4.json')); print(d['e2e'])"
