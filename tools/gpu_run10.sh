# round 2, call 10: full suite with thinned tiles, ncu of the constant-quality error model
O=gpurun_out/r2j; mkdir -p $O
(time python -m pytest tests -m gpu -q --durations=5) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
python bench.py --no-cpu-baseline --no-e2e --steps 5 > $O/bench.json 2> $O/bench.err; python -c "import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3_constant -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sequencer constant > /dev/null 2> $O/ncu_c3c.err
ls -la $O
