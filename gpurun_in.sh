python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 12 --warmup 3 > gpurun_out/bench_r1_short.json 2> gpurun_out/bench_r1_short.err; tail -c 1500 gpurun_out/bench_r1_short.json
