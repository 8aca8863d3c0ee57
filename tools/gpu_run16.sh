# round 2, call 16: policy-selected carried_sid shape, unroll by model: parity, A/B against the round's first thinned kernel, e2e laps
O=gpurun_out/r2p; mkdir -p $O
(time python -m pytest tests -m gpu -x -q --durations=3) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
run() { n=$1; lib=$2; shift 2
  PCS_LIB=$lib python bench.py --no-cpu-baseline --no-e2e --steps 5 "$@" > $O/$n.json 2> $O/$n.err
  python -c "import json; d=json.load(open('$O/$n.json')); print('$n', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"
}
NEW=$PWD/process_b200/libpcs_seq.so; OLD=$PWD/process_b200/libpcs_seq_old.so
run old_errorless $OLD
run new_errorless $NEW
run new_constant $NEW --sequencer constant
run new_random $NEW --sequencer random
run new_paired_constant $NEW --insert-size 300 --sequencer constant
run new_C5 $NEW --workload C5
run new_C4 $NEW --workload C4
PCS_TIMING=1 python bench.py --no-cpu-baseline --steps 3 > $O/bench.json 2> $O/bench.err; python -c "import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e']))"
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3_constant -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sequencer constant > /dev/null 2> $O/ncu_c3c.err
ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 -o $O/prof_c3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> $O/ncu_c3.err
