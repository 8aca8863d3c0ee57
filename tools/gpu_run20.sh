# round 2, call 20: geometry kept by the forest, last sample launched in chromosome groups: parity, e2e A/B
O=gpurun_out/r2t; mkdir -p $O
(time python -m pytest tests -m gpu -x -q --durations=3) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
for v in split nosplit split nosplit; do
  if [ $v = nosplit ]; then export PCS_NO_SPLIT=1; else unset PCS_NO_SPLIT; fi
  python bench.py --no-cpu-baseline --steps 3 > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "import json; d=json.load(open('$O/bench_$v.json')); e=d['e2e']; print('$v: cold', round(e['ms_per_step'],2), 'resident', round(e['forest_resident']['ms_per_step'],2), 'api', round(e['api_ms'],1))"
done
unset PCS_NO_SPLIT
PCS_TIMING=1 python bench.py --no-cpu-baseline --steps 3 > $O/bench_timing.json 2> $O/bench_timing.err
grep -n "simulate (plan\|GPU: first\|tables copied\|plan + launch" $O/bench_timing.err | sed -n 8,20p
