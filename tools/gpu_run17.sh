# round 2, call 17 (2 GPUs): step with the coverage gather on a side stream; error codes straight from the block (N=1 check)
O=gpurun_out/r2q; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact or recount or shards" > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
for s in constant random; do
  python bench.py --no-cpu-baseline --no-e2e --steps 5 --sequencer $s > $O/bench_$s.json 2> $O/bench_$s.err
  python -c "import json; d=json.load(open('$O/bench_$s.json')); print('$s', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/c3_n2.json 2> $O/c3_n2.err
python -c "import json; d=json.load(open('$O/c3_n2.json')); print('n2', d['value'], d['ms_per_step'], d['detail']['sampler_kernel_ms_per_rank'], d['checks']); print(json.dumps(d['e2e']))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --exchange nccl > $O/c3_n2_nccl.json 2> $O/c3_n2_nccl.err
python -c "import json; d=json.load(open('$O/c3_n2_nccl.json')); print('n2 nccl', d['value'], d['ms_per_step'])"
