# round 2, call 14: thin-loop variants (unroll / invariants in shared memory) x sequencer models
O=gpurun_out/r2n; mkdir -p $O
run() { # name lib args...
  n=$1; lib=$2; shift 2
  PCS_LIB=$lib python bench.py --no-cpu-baseline --no-e2e --steps 5 "$@" > $O/$n.json 2> $O/$n.err
  python -c "import json; d=json.load(open('$O/$n.json')); print('$n', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"
}
for v in a b c d; do
  lib=$PWD/process_b200/libpcs_seq_$v.so; [ $v = a ] && lib=$PWD/process_b200/libpcs_seq.so
  run ${v}_errorless $lib
  run ${v}_constant $lib --sequencer constant
  PCS_MIN_CTAS=5 run ${v}_constant_ctas5 $lib --sequencer constant
  PCS_MIN_CTAS=3 run ${v}_constant_ctas3 $lib --sequencer constant
  run ${v}_random $lib --sequencer random
done
