timeout 200 python scripts/ctc_dbg.py 2>&1 | tail -40
