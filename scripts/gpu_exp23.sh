for g in 148 111 74 37; do echo "grid $g"; GR_A32_GRID=$g timeout 100 python scripts/trace_a32.py fwd | grep "ms\|cycles per"; done
