echo "=== normal"; timeout 120 python scripts/trace_tc.py 2>&1 | grep -v "globaltimer\|slowest"
echo "=== no h loads"; GR_TC_DBG=1 timeout 120 python scripts/trace_tc.py 2>&1 | grep -v "globaltimer\|slowest"
echo "=== no MMAs"; GR_TC_DBG=2 timeout 120 python scripts/trace_tc.py 2>&1 | grep -v "globaltimer\|slowest"
echo "=== neither"; GR_TC_DBG=3 timeout 120 python scripts/trace_tc.py 2>&1 | grep -v "globaltimer\|slowest"
