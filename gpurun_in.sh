python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 12 --warmup 3 > gpurun_out/bench_r1_short.json 2> gpurun_out/bench_r1_short.err; python -c "
import json
d=json.load(open('gpurun_out/bench_r1_short.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['share_of_step'], d['roofline']['exp_pipe'])"
