# round 2, call 2: tests of the pipelined call + device result, bench with host phase times, fresh ncu captures
O=gpurun_out/r2b; mkdir -p $O
(time python -m pytest tests -m gpu -x -q --durations=8) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
PCS_TIMING=1 python bench.py --cpu-reads 5e7 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
cat $O/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> $O/ncu_c3.err
ls -la $O
