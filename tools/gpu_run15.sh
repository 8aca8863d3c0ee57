# round 2, call 15: rolled thin loop as the default: full parity, A/B against the round's first thinned kernel in one box, profile
O=gpurun_out/r2o; mkdir -p $O
(time python -m pytest tests -m gpu -x -q --durations=3) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
run() { n=$1; lib=$2; shift 2
  PCS_LIB=$lib python bench.py --no-cpu-baseline --no-e2e --steps 5 "$@" > $O/$n.json 2> $O/$n.err
  python -c "import json; d=json.load(open('$O/$n.json')); print('$n', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"
}
for rep in 1 2; do
run old_errorless_$rep $PWD/process_b200/libpcs_seq_old.so
run new_errorless_$rep $PWD/process_b200/libpcs_seq.so
done
PCS_MIN_CTAS=4 run new_errorless_ctas4 $PWD/process_b200/libpcs_seq.so
PCS_MIN_CTAS=6 run new_errorless_ctas6 $PWD/process_b200/libpcs_seq.so
run new_paired $PWD/process_b200/libpcs_seq.so --insert-size 300
run old_paired $PWD/process_b200/libpcs_seq_old.so --insert-size 300
run new_paired_constant $PWD/process_b200/libpcs_seq.so --insert-size 300 --sequencer constant
run old_paired_constant $PWD/process_b200/libpcs_seq_old.so --insert-size 300 --sequencer constant
run new_C5 $PWD/process_b200/libpcs_seq.so --workload C5
run new_C2 $PWD/process_b200/libpcs_seq.so --workload C2
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3_constant -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sequencer constant > /dev/null 2> $O/ncu_c3c.err
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> $O/ncu_c3.err
ls $O | head -50
