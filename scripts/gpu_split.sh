for s in 0 -1; do
echo "== GR_TOWER_PRIO=$s"
GR_TOWER_PRIO=$s timeout 300 python bench.py --steps 6 --warmup 3 --skip-cpu --skip-ctc > gpurun_out/prio_$s.json 2> gpurun_out/prio_$s.err
tail -3 gpurun_out/prio_$s.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/prio_$s.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['loss_mean'])"
done
