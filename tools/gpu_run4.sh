# round 2, call 4 (2 GPUs): new single-GPU tests, then the 2-rank bench (async step, one-process multi-GPU e2e)
O=gpurun_out/r2d; mkdir -p $O
(time python -m pytest tests/test_gpu_result.py tests/test_gpu_genomes.py tests/test_gpu_shared_tables.py tests/test_gpu_parity.py -m gpu -x -q -k "result or genomes or shared or contexts or sample_by_sample or ipc" --durations=5) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
PCS_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
cat $O/bench_n2.json
PCS_BALANCE=templates python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $O/bench_n2_templates.json 2> $O/bench_n2_templates.err
python -c "import json; d=json.load(open('$O/bench_n2_templates.json')); print('balance=templates', d['ms_per_step'], d['detail']['sampler_kernel_ms_per_rank'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --exchange nccl > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err
python -c "import json; d=json.load(open('$O/bench_n2_nccl.json')); print('nccl', d['ms_per_step'])"
tail -5 $O/bench_n2.err
