"""Times one GEMM-shaped op through the C ABI (tuning / roofline evidence):
   python tools/op_bench.py conv3x3|gemm B H W Cin N [variant]     (variant: 0 auto, 1 tiled tcgen05, 2 streaming tcgen05)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddpm_image_restoration_b200 import ops, _lib

kind = sys.argv[1]
B, H, W, Cin, N = (int(v) for v in sys.argv[2:7])
variant = int(sys.argv[7]) if len(sys.argv) > 7 else 0
_lib.lib().ddpmir_igemm_set_variant(variant)
taps = 9 if kind == "conv3x3" else 1
x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, taps * Cin, device="cuda") / (taps * Cin) ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
res = torch.randn(B, H, W, N, device="cuda")
fn = ops.conv3x3 if taps == 9 else ops.gemm
run = lambda: fn(x, w, N, ops.IMPL_TENSOR, out_dtype=torch.float32, bias=bias, res=res)
for _ in range(3):
    run()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
n, tot = 10, 0.0
for _ in range(n):
    flush.zero_()                                  # evict L2 between timed launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
ms = tot / n
M = B * H * W
flops = 2.0 * M * N * taps * Cin
bytes_ = M * Cin * 2 + N * taps * Cin * 2 + M * N * 4 * 2      # A once + W + fp32 residual in + fp32 out
print(f"{kind} B={B} {H}x{W} Cin={Cin} N={N} variant={variant}: {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s  "
      f"{bytes_ / ms / 1e6:.0f} GB/s (algorithmic bytes {bytes_ / 1e6:.1f} MB)")
