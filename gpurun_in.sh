python bench.py --batch 16 --micro-batches 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mb1_plain.json 2> gpurun_out/bench_mb1_plain.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --batch 16 --micro-batches 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -c 600 gpurun_out/bench_mb1_plain.json; wc -l gpurun_out/launches_r1b.csv
