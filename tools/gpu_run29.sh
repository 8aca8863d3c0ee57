# round 2: PCS_ASYNC_UPLOAD (plan slices on their own stream, coverage gather on the copy stream): parity, then e2e A/B
O=gpurun_out/r2ae; mkdir -p $O
PCS_ASYNC_UPLOAD=1 python -m pytest tests/test_gpu_result.py -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
for v in 1 0 1 0; do
  PCS_ASYNC_UPLOAD=$v python bench.py --no-cpu-baseline --steps 3 > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "import json; d=json.load(open('$O/bench_$v.json')); e=d['e2e']; print('async $v: cold', round(e['ms_per_step'],2), 'resident', round(e['forest_resident']['ms_per_step'],2))"
done
