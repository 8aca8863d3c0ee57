# round 2: racecheck again (warp barrier between the queue's reads and its put-backs), parity of the error models, fresh capture of the constant-quality kernel
O=gpurun_out/r2ab; mkdir -p $O
timeout 400 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $O/racecheck.log 2>&1; tail -3 $O/racecheck.log
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck.log 2>&1; tail -2 $O/memcheck.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_closed_form.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3_constant -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sequencer constant > /dev/null 2> $O/ncu_c3c.err
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> $O/ncu_c3.err
python bench.py --no-cpu-baseline --no-e2e --steps 5 --sequencer constant 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('constant', d['ms_per_step'])"
ls $O
