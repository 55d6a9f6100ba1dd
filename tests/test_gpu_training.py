"""Training step (webp_training.py:476-537) on the GPU kernels against torch.autograd over the restated oracle:
loss value, every parameter gradient, and the parameters after clip_grad_norm_ + AdamW steps (dropout disabled so the
comparison is deterministic; the dropout kernel has its own test)."""
import pytest
import torch

from oracle import restated as R
from oracle import weights as W
from util import rel

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def make(fam="webp", precision="fp32"):
    import ddpm_image_restoration_b200 as P
    from ddpm_image_restoration_b200.training import Trainer
    sd = W.make_state_dict(fam, 0)
    m = {"webp": P.WebPDiffusionModel, "jpeg": P.JPEGDiffusionModel, "avif": P.AVIFDiffusionModel}[fam]()
    m.load_state_dict(sd)
    m = m.cuda().set_precision(precision)
    return sd, m, Trainer(m, dropout=0.0)


def batch(hw=32, b=2):
    x0 = W.synthetic_images(b, hw, hw, seed=77)
    xt = R.codec_roundtrip(x0, 10, "webp")
    t = torch.tensor([37.0, 81.0][:b]) / 100.0
    return x0, xt, t


@pytest.mark.parametrize("fam", ["webp", "jpeg", "avif"])
def test_gradients_match_autograd_fp32(fam):
    sd, m, tr = make(fam)
    x0, xt, t = batch()
    loss_ref, grads_ref = R.training_step_reference(sd, xt, t, x0, fam)
    loss = tr.forward_backward(xt.cuda(), t.cuda(), x0.cuda())
    assert abs(float(loss) - float(loss_ref)) < 2e-3 * abs(float(loss_ref))
    # parameters the forward never uses (AVIF inverse_weights) have no autograd gradient; ours stay exactly zero
    assert set(grads_ref) == set(tr.grads) - tr.unused
    assert all(k.endswith("inverse_weights") and not tr.grads[k].any() for k in tr.unused)
    assert (fam == "avif") == bool(tr.unused)
    grads_ref = dict(grads_ref, **{k: torch.zeros_like(tr.grads[k]).cpu() for k in tr.unused})
    worst = max((rel(tr.grads[k].cpu(), g), k) for k, g in grads_ref.items())
    flat_ref = torch.cat([grads_ref[k].flatten() for k in tr.grads])
    print(f"{fam}: loss {float(loss):.6f} vs {float(loss_ref):.6f}; global grad rel-L2 {rel(tr.flat_grad.cpu(), flat_ref):.2e}; worst tensor {worst}")
    # the loss' phase term has a 1/|P| gradient (ill-conditioned at weak Fourier coefficients): 5e-3 overall
    assert rel(tr.flat_grad.cpu(), flat_ref) < 5e-3
    assert worst[0] < 5e-2


def test_optimizer_steps_match_torch():
    sd, m, tr = make("webp")
    x0, xt, t = batch()
    ref = {k: v.clone() for k, v in sd.items()}
    names = [k for k in tr.params]
    plist = [ref[k].requires_grad_() for k in names]
    opt = torch.optim.AdamW(plist, lr=2e-4, weight_decay=1e-5, betas=(0.9, 0.99))
    for step in range(2):
        cur = {k: v.detach() for k, v in ref.items()}
        _, grads = R.training_step_reference(cur, xt, t, x0, "webp")
        for k, p in zip(names, plist):
            p.grad = grads[k]
        torch.nn.utils.clip_grad_norm_(plist, 1.0)
        opt.step()
        tr.train_step(xt.cuda(), t.cuda(), x0.cuda())
    got = torch.cat([p.detach().flatten().cpu() for p in tr.params.values()])
    want = torch.cat([p.detach().flatten() for p in plist])
    start = torch.cat([sd[k].flatten() for k in names])
    # compare the UPDATE (2 Adam steps move every weight by ~ +-4e-4)
    assert rel(got - start, want - start) < 2e-2
    # the inference path sees the updated weights
    with torch.no_grad():
        out = m.eval()(xt.cuda(), t.cuda()).cpu()
    cur = {k: v.detach() for k, v in ref.items()}
    assert rel(out, R.unet_forward(cur, xt, t, None, "webp")) < 1e-3


def test_bf16_gradients_are_aligned():
    sd, m, tr = make("webp", "bf16")
    x0, xt, t = batch()
    _, grads_ref = R.training_step_reference(sd, xt, t, x0, "webp")
    tr.forward_backward(xt.cuda(), t.cuda(), x0.cuda())
    flat_ref = torch.cat([grads_ref[k].flatten() for k in tr.grads])
    g = tr.flat_grad.cpu()
    g, flat_ref = g.double(), flat_ref.double()
    cos = float((g * flat_ref).sum() / (g.norm() * flat_ref.norm()))
    print(f"bf16 gradient cosine vs fp32 autograd: {cos:.5f}, rel-L2 {rel(g, flat_ref):.3e}")
    assert cos > 0.99


def test_dropout_changes_the_step_but_is_reproducible():
    sd, m, tr = make("webp")
    tr.dropout_p = 0.1
    x0, xt, t = batch()
    l1 = float(tr.forward_backward(xt.cuda(), t.cuda(), x0.cuda(), dropout_seed=5)); g1 = tr.flat_grad.clone()
    l2 = float(tr.forward_backward(xt.cuda(), t.cuda(), x0.cuda(), dropout_seed=5)); g2 = tr.flat_grad.clone()
    l3 = float(tr.forward_backward(xt.cuda(), t.cuda(), x0.cuda(), dropout_seed=6))
    # same seed -> same mask (only the order of fp32 atomic sums differs); another seed -> another loss
    assert abs(l1 - l2) < 1e-5 * abs(l1) and abs(l1 - l3) > 1e-4 * abs(l1)
    assert rel(g1, g2) < 1e-4


def test_avif_training_step_matches_torch():
    """train_epoch_ddrm_avif's step (avif.py:562-577): two clip + AdamW(lr 1.5e-4) steps against torch.optim.AdamW fed with the
    oracle's autograd gradients; inverse_weights (no gradient in torch) must not move, not even by weight decay."""
    from ddpm_image_restoration_b200.training import train_epoch_ddrm_avif
    sd, m, tr = make("avif")
    assert tr.lr == 1.5e-4
    x0, xt, t = batch()
    ref = {k: v.clone() for k, v in sd.items()}
    names = [k for k in tr.params]
    plist = [ref[k].requires_grad_() for k in names]
    opt = torch.optim.AdamW(plist, lr=1.5e-4, weight_decay=1e-5, betas=(0.9, 0.99))
    for step in range(2):
        cur = {k: v.detach() for k, v in ref.items()}
        _, grads = R.training_step_reference(cur, xt, t, x0, "avif")
        for k, p in zip(names, plist):
            p.grad = grads.get(k)
        torch.nn.utils.clip_grad_norm_(plist, 1.0)
        opt.step()
        if step == 0:
            tr.train_step(xt.cuda(), t.cuda(), x0.cuda())
        else:
            train_epoch_ddrm_avif(tr, [(x0.cuda(), xt.cuda(), (t * 100).cuda())])
    used = [k for k in names if k not in tr.unused]
    got = torch.cat([tr.params[k].detach().flatten().cpu() for k in used])
    want = torch.cat([ref[k].detach().flatten() for k in used])
    start = torch.cat([sd[k].flatten() for k in used])
    assert rel(got - start, want - start) < 2e-2
    for k in tr.unused:
        assert torch.equal(tr.params[k].detach().cpu(), sd[k])
    assert tr.lr < 1.5e-4          # the epoch loop stepped the cosine schedule


def test_avif_bf16_gradients_are_aligned():
    sd, m, tr = make("avif", "bf16")
    x0, xt, t = batch()
    _, grads_ref = R.training_step_reference(sd, xt, t, x0, "avif")
    tr.forward_backward(xt.cuda(), t.cuda(), x0.cuda())
    keys = [k for k in tr.grads if k not in tr.unused]
    g = torch.cat([tr.grads[k].flatten() for k in keys]).cpu().double()
    flat_ref = torch.cat([grads_ref[k].flatten() for k in keys]).double()
    cos = float((g * flat_ref).sum() / (g.norm() * flat_ref.norm()))
    print(f"avif bf16 gradient cosine vs fp32 autograd: {cos:.5f}, rel-L2 {rel(g, flat_ref):.3e}")
    assert cos > 0.99
