"""Times the batched Jacobi SVD low-rank kernel (config 3: 64 images x 3 planes of 256x256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpm_image_restoration_b200 as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
x = torch.randn(B, 3, H, H, device="cuda") * 0.3 + torch.linspace(-1, 1, H, device="cuda").view(1, 1, 1, H)
for kr in (0.9, 0.6):
    P.svd_structure_preservation(x, kr); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = P.svd_structure_preservation(x, kr); e1.record(); torch.cuda.synchronize()
    print(f"svd_lowrank {B * 3} planes of {H}x{H}, k_ratio {kr}: {e0.elapsed_time(e1):.2f} ms")
