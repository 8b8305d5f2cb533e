timeout 100 python scripts/trace_a32.py ffwd
GR_A32_NVG=1 timeout 100 python scripts/trace_a32.py ffwd
