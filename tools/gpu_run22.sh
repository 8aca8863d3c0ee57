# round 2, call 22: huge pages for the caller's output tables (first-touch cost of the copy-out)
O=gpurun_out/r2v; mkdir -p $O
cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag; nproc; free -g | head -2
for v in 1 0 1 0; do
  PCS_HUGE_OUT=$v python bench.py --no-cpu-baseline --steps 3 > $O/bench_huge$v.json 2> $O/bench_huge$v.err
  python -c "import json; d=json.load(open('$O/bench_huge$v.json')); e=d['e2e']; print('huge $v: cold', round(e['ms_per_step'],2), 'resident', round(e['forest_resident']['ms_per_step'],2), 'result vaf', round(e['device_result']['with_vaf']['ms_per_step'],1), 'no vaf', round(e['device_result']['without_vaf']['ms_per_step'],1), 'api', round(e['api_ms'],1))"
done
python -m pytest tests/test_gpu_result.py tests/test_gpu_shared_tables.py -m gpu -x -q 2>&1 | tail -2
