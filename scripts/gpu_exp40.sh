# which of the two added fences matter (16 = no per-thread gpu fence before the release, 1024 = no proxy fence before p_consumed), and their cost
for d in 16 1040; do echo "== dbg=$d"; GR_TC_DBG=$d RUNS=40 timeout 250 python scripts/determinism_check.py 2>&1 | tail -4; done
for d in 0 16 1024 1040; do echo "== perf dbg=$d"; GR_TC_DBG=$d SHAPES=256x500,256x300 timeout 200 python scripts/lstm_perf.py 2>&1 | tail -4; done
