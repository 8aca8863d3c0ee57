# round 2, call 23 (8 GPUs): pcs_simulate_multi, tables home over one link or over every device's
O=gpurun_out/r2w; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
for v in all 1 4 all 1; do
  if [ $v = all ]; then unset PCS_MULTI_LINKS; else export PCS_MULTI_LINKS=$v; fi
  PCS_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > $O/c3_n8_links_$v.json 2> $O/c3_n8_links_$v.err
  python -c "import json; d=json.load(open('$O/c3_n8_links_$v.json')); e=d['e2e']; print('links $v: value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'cold', round(e['ms_per_step'],2), 'resident', round(e['forest_resident']['ms_per_step'],2))"
done
grep -n "simulate_multi\|flatten_forest \|replic" $O/c3_n8_links_all.err | tail -12
