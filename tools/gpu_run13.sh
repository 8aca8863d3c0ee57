# round 2, call 13: error models -- two bits of the pair's error block inside the walk, the rest through the queue
O=gpurun_out/r2m; mkdir -p $O
(time python -m pytest tests/test_gpu_parity.py tests/test_gpu_closed_form.py tests/test_gpu_sam.py -m gpu -x -q --durations=3) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
for s in constant random; do
  python bench.py --no-cpu-baseline --no-e2e --steps 5 --sequencer $s > $O/bench_$s.json 2> $O/bench_$s.err
  python -c "import json; d=json.load(open('$O/bench_$s.json')); print('$s', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
for c in 3 5; do
  PCS_MIN_CTAS=$c python bench.py --no-cpu-baseline --no-e2e --steps 5 --sequencer constant > $O/bench_constant_ctas$c.json 2> $O/bench_constant_ctas$c.err
  python -c "import json; d=json.load(open('$O/bench_constant_ctas$c.json')); print('constant ctas $c', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
python bench.py --no-cpu-baseline --no-e2e --steps 5 --insert-size 300 --sequencer constant > $O/bench_paired_constant.json 2> $O/bench_paired_constant.err
python -c "import json; d=json.load(open('$O/bench_paired_constant.json')); print('paired constant', d['ms_per_step'], d['roofline']['kernel_ms'])"
for w in C2 C5; do
  python bench.py --no-cpu-baseline --no-e2e --steps 3 --workload $w > $O/bench_$w.json 2> $O/bench_$w.err
  python -c "import json; d=json.load(open('$O/bench_$w.json')); print('$w', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
done
PCS_TIMING=1 python bench.py --no-cpu-baseline --steps 3 > $O/bench.json 2> $O/bench.err; python -c "import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e']))"
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3_constant -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sequencer constant > /dev/null 2> $O/ncu_c3c.err
ls -la $O
