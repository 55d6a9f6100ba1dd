timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "prescaled" 2>&1 | tail -3
for sel in 0 2 3 4 10 11 12; do python tools/attn_bench.py 8 8 65536 4 $(( (sel+1)*256 )) 1; done
for sel in 0 2 3 11; do python tools/attn_bench.py 16 4 65536 4 $(( (sel+1)*256 )) 1; done
