for s in 1 2 4 5 6 13; do
  python bench.py --steps 20 --warmup 3 --no-cpu --streams $s 2>/dev/null > gpurun_out/sweep_$s.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/sweep_$s.json').read().strip().splitlines()[-1]); print('streams', $s, d['ms_per_step'])"
done
