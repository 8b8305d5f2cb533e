mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
