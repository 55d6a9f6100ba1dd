python -m pytest tests/test_gpu_sampler.py tests/test_gpu_ops.py -m gpu -q -k "dct or large_logits" 2>&1 | tail -15
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from ddpm_image_restoration_b200 import ops
x=torch.rand(64,3,256,256,device='cuda')*255
for _ in range(3): ops.jpeg_dct_project(x,10)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.jpeg_dct_project(x,10)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/20
print(f"jpeg_dct_project 64x3x256x256: {ms*1e3:.1f} us, {x.numel()*8/ms/1e6:.0f} GB/s")
PY
