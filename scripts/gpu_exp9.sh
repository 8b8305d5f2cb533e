timeout 100 python scripts/trace_a32.py fwd
timeout 100 python scripts/trace_a32.py dw
