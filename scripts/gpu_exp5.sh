for d in 1 5 9 13 4 8 12; do echo "=== dbg=$d"; GR_TC_DBG=$d timeout 120 python scripts/trace_tc.py 2>&1 | grep "B256\|mma_committed\|step length"; done
