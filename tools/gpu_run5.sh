# round 2, call 5: full GPU suite (general closed form, tracks, genomes), bench with the faster flatten, SAM throughput
O=gpurun_out/r2e; mkdir -p $O
(time python -m pytest tests -m gpu -x -q --durations=8) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -14 $O/pytest_gpu.log
PCS_TIMING=1 python bench.py --cpu-reads 3e7 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench.json")); print(d["value"], d["ms_per_step"]); print(json.dumps(d["e2e"], indent=1)); print(d["cpu_baseline"])
PY
grep "pcs flatten" $O/bench.err | head -9
python tools/sam_throughput.py 20 > $O/sam_throughput.json 2> $O/sam.err; cat $O/sam_throughput.json
