# round 2, call 8 (2 GPUs): full suite after the planner change, SAM writer thread, replicate over NVLink
O=gpurun_out/r2h; mkdir -p $O
(time python -m pytest tests -m gpu -x -q --durations=5) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -10 $O/pytest_gpu.log
python tools/sam_throughput.py 20 > $O/sam_throughput.json 2> $O/sam.err; cat $O/sam_throughput.json
PCS_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
grep "pcs host" $O/bench_n2.err | grep -v "plan:" | tail -12
python -c "import json; d=json.load(open('$O/bench_n2.json')); print(d['value'], d['ms_per_step'], d['detail']['sampler_kernel_ms_per_rank']); print(d['e2e'])"
PCS_TIMING=1 python bench.py --no-cpu-baseline --steps 5 > $O/bench_n1.json 2> $O/bench_n1.err
python -c "import json; d=json.load(open('$O/bench_n1.json')); print(d['value'], d['ms_per_step']); print(d['e2e'])"
grep "plan + launch\|simulate (plan" $O/bench_n1.err | tail -4
