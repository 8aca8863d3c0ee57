# round 2, call 21 (2 GPUs): pcs_simulate_multi sends the tables home over every device's link
O=gpurun_out/r2u; mkdir -p $O
python -m pytest tests/test_gpu_shared_tables.py tests/test_gpu_result.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export PCS_MULTI_LINKS=1; else unset PCS_MULTI_LINKS; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$v bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $O/c3_n2_links$v.json 2> $O/c3_n2_links$v.err
  python -c "import json; d=json.load(open('$O/c3_n2_links$v.json')); e=d['e2e']; print('one link only = $v: value', round(d['value']), 'cold', round(e['ms_per_step'],2), 'resident', round(e['forest_resident']['ms_per_step'],2))"
done
