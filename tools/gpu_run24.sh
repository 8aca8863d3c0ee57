# round 2, final call: the records of the round (profiles/r02_v5_*)
O=gpurun_out/r2x; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
(time python -m pytest tests -m gpu -x -q --durations=5) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
python bench.py > $O/bench.json 2> $O/bench.err; python -c "import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['issue']); print(json.dumps(d['e2e'])); print(d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> $O/ncu_c3.err
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3_constant -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sequencer constant > /dev/null 2> $O/ncu_c3c.err
(time python bench.py --impl reference --steps 1 --warmup 0) > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 600 $O/bench_reference.json
ls -la $O
