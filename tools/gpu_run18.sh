# round 2, call 18 (8 GPUs): C3 and C4 with the thinned kernel and the side-stream coverage gather
O=gpurun_out/r2r; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > $O/c3_n8.json 2> $O/c3_n8.err
python -c "import json; d=json.load(open('$O/c3_n8.json')); print('c3 n8', d['value'], d['ms_per_step'], d['detail']['sampler_kernel_ms_per_rank'], d['checks']); print(json.dumps(d['e2e']))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload C4 > $O/c4_n8.json 2> $O/c4_n8.err
python -c "import json; d=json.load(open('$O/c4_n8.json')); print('c4 n8', d['value'], d['ms_per_step'], d['detail']['sampler_kernel_ms_per_rank'], d['checks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/c3_n4.json 2> $O/c3_n4.err
python -c "import json; d=json.load(open('$O/c3_n4.json')); print('c3 n4', d['value'], d['ms_per_step'], d['detail']['sampler_kernel_ms_per_rank'], d['checks'])"
