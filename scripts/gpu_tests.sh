# full GPU suite + smoke (what the driver runs at round end)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -8 | tee gpurun_out/r2_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
