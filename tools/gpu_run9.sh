# round 2, call 9: thinned tiles (single-end): parity first, then speed against PCS_THIN=0
O=gpurun_out/r2i; mkdir -p $O
(time python -m pytest tests -m gpu -x -q --durations=5) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
PCS_THIN=0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_closed_form.py -m gpu -x -q > $O/pytest_gpu_nothin.log 2>&1; tail -2 $O/pytest_gpu_nothin.log
for t in 1 0; do
  PCS_THIN=$t python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench_thin$t.json 2> $O/bench_thin$t.err
  python -c "import json; d=json.load(open('$O/bench_thin$t.json')); print('thin $t', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
done
for s in constant random; do
  python bench.py --no-cpu-baseline --no-e2e --steps 5 --sequencer $s > $O/bench_$s.json 2> $O/bench_$s.err
  python -c "import json; d=json.load(open('$O/bench_$s.json')); print('$s', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
for c in 4 5 6; do
  PCS_MIN_CTAS=$c python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench_ctas$c.json 2> $O/bench_ctas$c.err
  python -c "import json; d=json.load(open('$O/bench_ctas$c.json')); print('ctas $c', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
PCS_TIMING=1 python bench.py --no-cpu-baseline --steps 5 > $O/bench.json 2> $O/bench.err; python -c "import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step']); print(d['e2e'])"
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> $O/ncu_c3.err
