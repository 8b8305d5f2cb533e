timeout 250 python -m pytest tests/test_gpu_decode.py tests/test_golden.py -q -m gpu --timeout 60 -x 2>&1 | tail -3
timeout 100 python - <<'PY'
import sys; sys.path.insert(0, ".")
import torch, bench
d = bench.decode_microbench(torch.device("cuda:0"), 6549.4)
print(d)
PY
