timeout 200 python -m pytest tests/test_gpu_ctc.py tests/test_golden.py -q -m gpu --timeout 100 2>&1 | tail -15
timeout 100 python - <<'PY'
import sys; sys.path.insert(0, ".")
import torch, bench
print(bench.ctc_microbench(torch.device("cuda:0"), 6549.4))
PY
