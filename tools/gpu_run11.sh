# round 2, call 11: the default bench line as the driver runs it (thinned tiles), host laps, fresh errorless capture
O=gpurun_out/r2k; mkdir -p $O
python bench.py > $O/bench.json 2> $O/bench.err; python -c "import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e'])); print(d['cpu_baseline'])"
PCS_TIMING=1 python bench.py --no-cpu-baseline --steps 3 > $O/bench_timing.json 2> $O/bench_timing.err
for s in constant random; do
  python bench.py --no-cpu-baseline --no-e2e --steps 5 --sequencer $s > $O/bench_$s.json 2> $O/bench_$s.err
  python -c "import json; d=json.load(open('$O/bench_$s.json')); print('$s', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
python bench.py --no-cpu-baseline --no-e2e --steps 5 --insert-size 300 > $O/bench_paired.json 2> $O/bench_paired.err
python -c "import json; d=json.load(open('$O/bench_paired.json')); print('paired', d['ms_per_step'], d['roofline']['kernel_ms'])"
for w in C2 C4 C5; do
  python bench.py --no-cpu-baseline --no-e2e --steps 3 --workload $w > $O/bench_$w.json 2> $O/bench_$w.err
  python -c "import json; d=json.load(open('$O/bench_$w.json')); print('$w', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
done
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> $O/ncu_c3.err
ls -la $O
