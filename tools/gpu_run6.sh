# round 2, call 6 (8 GPUs): scaling of C3 at 8 and 4 ranks, and C4 (the 8-GPU config of BASELINE.json) at 8 ranks
O=gpurun_out/r2f; mkdir -p $O
run() { # name, nproc, port, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $2 --steps 20 --warmup 5 ${@:4} > $O/$1.json 2> $O/$1.err
  echo "$1 rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$O/$1.json"))
    e=d.get("e2e") or {}
    print("$1", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "kernel/rank", [round(x,2) for x in d["detail"]["sampler_kernel_ms_per_rank"]], "e2e ms", e.get("ms_per_step"), "resident", (e.get("forest_resident") or {}).get("ms_per_step"), d["checks"])
except Exception as ex:
    print("$1 no line:", ex)
PY
}
run c3_n8 8 29521
run c3_n4 4 29522
PCS_BALANCE=templates run c3_n8_templates 8 29523 --no-e2e
run c4_n8 8 29524 --workload C4
tail -3 $O/c4_n8.err
