# round 2, call 19: instance table built on the device: parity, then cold e2e against PCS_DEVICE_INSTANCES=0
O=gpurun_out/r2s; mkdir -p $O
(time python -m pytest tests -m gpu -x -q --durations=3) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
for v in 1 0 1 0; do
  PCS_DEVICE_INSTANCES=$v python bench.py --no-cpu-baseline --steps 3 > $O/bench_di$v.json 2> $O/bench_di$v.err
  python -c "import json; d=json.load(open('$O/bench_di$v.json')); e=d['e2e']; print('device instances $v: cold', round(e['ms_per_step'],2), 'resident', round(e['forest_resident']['ms_per_step'],2), 'h2d', e['h2d_bytes_per_step'], 'api', round(e['api_ms'],1), 'first', round(e['api']['first_call_ms'],1))"
done
PCS_TIMING=1 python bench.py --no-cpu-baseline --steps 3 > $O/bench_timing.json 2> $O/bench_timing.err
grep -n "flatten\]\|upload\|groups" $O/bench_timing.err | sed -n 30,50p
