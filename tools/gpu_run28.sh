# round 2: host code built for x86-64-v3 against the plain build, e2e in one box
O=gpurun_out/r2ac; mkdir -p $O
grep -m1 "model name" /proc/cpuinfo; grep -m1 -o "avx2" /proc/cpuinfo
for v in avx noavx avx noavx; do
  lib=$PWD/process_b200/libpcs_seq.so; [ $v = noavx ] && lib=$PWD/process_b200/libpcs_seq_noavx.so
  PCS_LIB=$lib PCS_TIMING=1 python bench.py --no-cpu-baseline --steps 3 > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "import json; d=json.load(open('$O/bench_$v.json')); e=d['e2e']; print('$v: cold', round(e['ms_per_step'],2), 'resident', round(e['forest_resident']['ms_per_step'],2))"
  grep "flatten_forest " $O/bench_$v.err | tail -3 | tr '\n' ' '; echo
done
