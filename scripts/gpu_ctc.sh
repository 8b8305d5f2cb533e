timeout 500 python -m pytest tests/test_gpu_ctc.py tests/test_golden.py tests/test_gpu_models.py -q -m gpu --timeout 150 2>&1 | tail -4
for impl in v5 v4; do
GR_CTC_IMPL=$impl timeout 200 python - <<'PY'
import sys, os; sys.path.insert(0, ".")
import torch, bench
r = bench.ctc_microbench(torch.device("cuda:0"), 6549.4)
print(os.environ.get("GR_CTC_IMPL"), r["ms"], r["roofline"]["frac"])
PY
done
