# round 2, last call: the new side-stream gather test, and the bench line with the final ncu counters (stale: false)
O=gpurun_out/r2y; mkdir -p $O
python -m pytest tests/test_gpu_shared_tables.py -m gpu -x -q > $O/pytest_shared.log 2>&1; tail -3 $O/pytest_shared.log
python bench.py > $O/bench.json 2> $O/bench.err; python -c "import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['issue']['frac'], d['roofline']['issue']['stale']); print(json.dumps(d['e2e'])[:400]); print(d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
