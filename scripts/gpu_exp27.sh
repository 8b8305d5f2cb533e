timeout 250 python -m pytest tests/test_gpu_models.py tests/test_gpu_gemm.py -q -m gpu --timeout 60 -x 2>&1 | tail -3
GR_TOWER_STREAMS=0 timeout 100 python scripts/step_breakdown.py 2>&1 | tail -22
timeout 200 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-ctc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'loss', d['loss_mean'])"
