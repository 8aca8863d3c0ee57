mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a/smi.txt 2>&1
(time python -m pytest tests -m gpu -x -q --durations=15) > gpurun_out/r2a/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a/smoke.log 2>&1
PCS_TIMING=1 python bench.py > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
tail -5 gpurun_out/r2a/pytest_gpu.log; cat gpurun_out/r2a/smoke.log | tail -2; cat gpurun_out/r2a/bench.json
