"""Per-kernel-class time of one training step (tuning aid): python tools/train_profile.py [B] [res] [precision]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpm_image_restoration_b200 as P
from ddpm_image_restoration_b200 import ops, ops_train, _lib
from ddpm_image_restoration_b200.training import Trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
res = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
m = P.WebPDiffusionModel().cuda().set_precision(prec)
tr = Trainer(m)
x0 = torch.rand(B, 3, res, res, device="cuda") * 2 - 1
xt = (x0 + 0.1 * torch.randn_like(x0)).clamp(-1, 1)
t = torch.rand(B, device="cuda")
for _ in range(2):
    tr.train_step(xt, t, x0)
torch.cuda.synchronize()
# wrap every ops / ops_train function with CUDA events
times = collections.defaultdict(lambda: [0, 0.0])
events = []
def wrap(mod, name):
    fn = getattr(mod, name)
    def w(*a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = fn(*a, **k); e.record()
        events.append((f"{mod.__name__.split('.')[-1]}.{name}", s, e))
        return r
    setattr(mod, name, w)
for mod, names in ((ops, ["conv3x3", "gemm", "groupnorm_stats", "groupnorm_apply", "block_transform", "maxpool2", "upsample2_concat", "conv_input",
                          "out_conv_tanh", "linear_rows", "lincomb", "cast_bf16", "mse", "ssim", "freq_loss_terms"]),
                   (ops_train, ["wgrad", "colsum", "groupnorm_backward", "gate_backward", "lrelu_mask_backward", "dropout", "maxpool2_backward",
                                "upsample2_concat_backward", "attention_train_forward", "attention_backward", "conv_input_backward",
                                "out_conv_tanh_backward", "frequency_aware_loss_backward", "linear_rows_backward", "adamw_step", "sumsq"])):
    for n in names:
        wrap(mod, n)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); tr.train_step(xt, t, x0); e1.record(); torch.cuda.synchronize()
for name, s, e in events:
    times[name][0] += 1; times[name][1] += s.elapsed_time(e)
tot = sum(v[1] for v in times.values())
print(f"train step B={B} {res}x{res} {prec}: {e0.elapsed_time(e1):.1f} ms total, {tot:.1f} ms inside wrapped ops")
for k, (n, ms) in sorted(times.items(), key=lambda kv: -kv[1][1]):
    print(f"  {ms:9.3f} ms  x{n:4d}  {k}")
