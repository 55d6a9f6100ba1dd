import torch


def nhwc(x, dtype=torch.float32):
    """NCHW cpu tensor -> NHWC cuda tensor of dtype."""
    return x.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()


def nchw(x):
    """NHWC cuda tensor -> NCHW cpu fp32."""
    return x.float().permute(0, 3, 1, 2).contiguous().cpu()


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def pack3(w, dtype):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().to(dtype).cuda()


def pack1(w, dtype):
    return w.reshape(w.shape[0], -1).contiguous().to(dtype).cuda()


def bf16_round(x):
    return x.to(torch.bfloat16).float()
