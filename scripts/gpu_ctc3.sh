for v in 0 1; do
GR_CTC_LSE3=$v timeout 200 python - <<'PY'
import sys, os; sys.path.insert(0, ".")
import torch, bench
print("lse3", os.environ.get("GR_CTC_LSE3"), bench.ctc_microbench(torch.device("cuda:0"), 6549.4)["ms"])
PY
done
GR_CTC_LSE3=1 timeout 300 python -m pytest tests/test_gpu_ctc.py -q -m gpu --timeout 150 2>&1 | tail -3
