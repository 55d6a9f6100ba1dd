"""One op call between cudaProfilerStart/Stop, after warm-up, so that `ncu --profile-from-start off --set full` captures exactly
the kernels of that call (tuning / roofline evidence):

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/x \\
        python tools/profile_call.py attention HD HEADS L B [SCALE]      # ddpmir_attention_prescaled_f16, all its tiers
        python tools/profile_call.py conv3x3|gemm B H W CIN N             # tcgen05 implicit GEMM with fp32 residual + fp32 out
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddpm_image_restoration_b200 import ops

kind = sys.argv[1]
torch.manual_seed(0)
if kind == "attention":
    hd, heads, L, B = (int(v) for v in sys.argv[2:6])
    scale = float(sys.argv[6]) if len(sys.argv) > 6 else 0.35
    C = hd * heads
    qkv = torch.randn(B, L, 3 * C, device="cuda") * scale
    qkv[..., :C] *= 1.4426950408889634 / hd ** 0.5
    qkv = qkv.to(ops.qkv_dtype_for_attention(L, hd))
    run = lambda: ops.attention_prescaled(qkv, heads)
else:
    B, H, W, Cin, N = (int(v) for v in sys.argv[2:7])
    taps = 9 if kind == "conv3x3" else 1
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, taps * Cin, device="cuda") / (taps * Cin) ** 0.5).to(torch.bfloat16)
    bias, res = torch.randn(N, device="cuda"), torch.randn(B, H, W, N, device="cuda")
    fn = ops.conv3x3 if taps == 9 else ops.gemm
    run = lambda: fn(x, w, N, ops.IMPL_TENSOR, out_dtype=torch.float32, bias=bias, res=res)
for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one call of", sys.argv[1:])
