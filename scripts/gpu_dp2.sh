mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err
echo "dp2 exit $?"; tail -c 1500 gpurun_out/bench_dp2.json; tail -5 gpurun_out/bench_dp2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dp_parity.py > gpurun_out/dp_parity.log 2>&1
echo "parity exit $?"; tail -5 gpurun_out/dp_parity.log
