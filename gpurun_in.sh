python -m pytest tests/test_gpu_ops.py -m gpu -q -k "full_resolution" 2>&1 | tail -15
python bench.py > gpurun_out/bench_r1_full.json 2> gpurun_out/bench_r1_full.err; tail -c 400 gpurun_out/bench_r1_full.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; tail -c 600 gpurun_out/bench_r1_ref.json
