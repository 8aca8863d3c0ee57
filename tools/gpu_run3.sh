# round 2, call 3: the one-locus-per-visit walk: parity first, then speed at 3/4/5/6 resident CTAs
O=gpurun_out/r2c; mkdir -p $O
(time python -m pytest tests -m gpu -x -q --durations=5) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
for c in 3 4 5 6; do
  PCS_MIN_CTAS=$c python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench_ctas$c.json 2> $O/bench_ctas$c.err
  python - <<PY
import json
d=json.load(open("$O/bench_ctas$c.json")); print("ctas $c", d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"])
PY
done
for s in constant random; do
  python bench.py --no-cpu-baseline --no-e2e --steps 5 --sequencer $s > $O/bench_$s.json 2> $O/bench_$s.err
  python -c "import json; d=json.load(open('$O/bench_$s.json')); print('$s', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
python bench.py --no-cpu-baseline --no-e2e --steps 5 --insert-size 300 > $O/bench_paired.json 2> $O/bench_paired.err
python -c "import json; d=json.load(open('$O/bench_paired.json')); print('paired', d['ms_per_step'], d['roofline']['kernel_ms'])"
python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; cat $O/bench.json
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> $O/ncu_c3.err
