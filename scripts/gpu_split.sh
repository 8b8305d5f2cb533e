timeout 300 python -m pytest tests/test_gpu_models.py -q -m gpu --timeout 150 2>&1 | tail -4
for s in 0 1; do
echo "== GR_PIPELINE=$s"
GR_PIPELINE=$s timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/pipe_$s.json 2> gpurun_out/pipe_$s.err
tail -3 gpurun_out/pipe_$s.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/pipe_$s.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss_mean'])"
done
