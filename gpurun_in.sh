for proj in codec dct; do python bench.py --family jpeg --projection $proj --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$proj', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['launch_ms'], d['roofline']['share_of_step'])"; done
for proj in codec; do python bench.py --family webp --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('webp', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['launch_ms'], d['roofline']['share_of_step'])"; done
