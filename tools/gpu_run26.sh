# round 2: compute-sanitizer on smoke() (thinned tiles, error block + queue, device-built instances)
O=gpurun_out/r2aa; mkdir -p $O
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck.log 2>&1; tail -3 $O/memcheck.log
timeout 400 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $O/racecheck.log 2>&1; tail -3 $O/racecheck.log
