"""Training row (BASELINE.json configs[4]) on the GPU: the fused kernels of csrc/train.cu against the
PyTorch expressions of the reference modules, and one full rate-distortion step against the reference's own CPU step
(tests/golden/train_step.npz, produced by oracle/make_golden.py train from the unmodified reference modules)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "train_step.npz"))


def _stf():
    from compressai.zoo import models
    from oracle import stf_ref, weights

    m = models["stf"]()
    m.load_state_dict(weights.seeded_state_dict(stf_ref.template_state_dict(), seed=0, stress=True), strict=False)
    return m.cuda().train()


def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_fused_gaussian_stage_matches_the_reference_expressions():
    """icm_gc_train_forward / backward vs GaussianConditional._likelihood + LowerBound + ste_round under autograd
    (entropy_models.py:626-659, bound_ops.py:21-62, ops.py:20-34), including scales below the 0.11 bound, likelihoods at the
    1e-9 bound, and a slice taken out of a wider latent (row stride != row length)."""
    from compressai.entropy_models import GaussianConditional
    from compressai.models._train import _gaussian_torch, gaussian_train
    from compressai.ops import ste_round

    gc = GaussianConditional(None).cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    B, C, H, W = 3, 32, 8, 12
    full = (torch.randn(B, 96, H, W, device="cuda", generator=g) * 6).requires_grad_()
    mu = (torch.randn(B, C, H, W, device="cuda", generator=g) * 2).requires_grad_()
    sc = torch.exp(torch.rand(B, C, H, W, device="cuda", generator=g) * 7 - 3.5)  # 0.03 .. 33: some below the bound
    sc[0, 0, 0, :4] = torch.tensor([0.11, 0.1099999, 0.1100001, 0.05], device="cuda")
    sc = sc.requires_grad_()
    noise = torch.rand(B, C, H, W, device="cuda", generator=g) - 0.5
    with torch.no_grad():
        full[1, 40:44] += 90.0  # far tails: likelihood clamps at 1e-9
    w_l = torch.randn(B, C, H, W, device="cuda", generator=g)  # mixed-sign upstream gradients exercise both pass-through rules
    w_h = torch.randn(B, C, H, W, device="cuda", generator=g)

    def run(fused):
        for t in (full, mu, sc):
            t.grad = None
        ys = full[:, 32:64]
        if fused:
            lik, y_hat = gaussian_train(gc, ys, mu, sc, noise)
        else:
            lik, y_hat = _gaussian_torch(gc, ys, mu, sc, noise), ste_round(ys - mu) + mu
        ((torch.log(lik) * w_l).sum() + (y_hat * w_h).sum()).backward()
        return lik.detach(), y_hat.detach(), full.grad.clone(), mu.grad.clone(), sc.grad.clone()

    a, b = run(True), run(False)
    assert torch.equal(a[1], b[1])                                      # y_hat: exact
    assert torch.allclose(a[0], b[0], rtol=1e-3, atol=1e-12)            # likelihoods within 1e-3 relative (north_star)
    assert (a[0] == 1e-9).any() and (b[0] == 1e-9).any()
    for k, name in ((2, "d/dy"), (3, "d/dmu"), (4, "d/dscale")):
        ref = b[k]
        err = (a[k] - ref).abs()
        tol = 2e-3 * ref.abs() + 1e-4 * ref.abs().max()
        assert bool((err <= tol).all()), f"{name}: max err {float(err.max()):.3g} at scale {float(ref.abs().max()):.3g}"
    assert torch.equal(a[4] == 0, b[4] == 0) or ((a[4] == 0) ^ (b[4] == 0)).float().mean() < 1e-3  # same pass-through pattern


@pytest.mark.parametrize("C,rows", [(48, 1000), (96, 777), (192, 513), (384, 130), (768, 67)])
def test_fused_layernorm_forward_and_backward_match_torch(C, rows):
    """icm_layernorm_train_forward / backward vs F.layer_norm under autograd (fp32 and bf16 output / incoming gradient)."""
    import torch.nn.functional as F

    from compressai.models._train import _LayerNormTrain

    g = torch.Generator(device="cuda").manual_seed(C)
    x = (torch.randn(rows, C, device="cuda", generator=g) * 3 + 1).requires_grad_()
    w = (1 + 0.2 * torch.randn(C, device="cuda", generator=g)).requires_grad_()
    b = (0.2 * torch.randn(C, device="cuda", generator=g)).requires_grad_()
    up = torch.randn(rows, C, device="cuda", generator=g)
    ref = F.layer_norm(x, (C,), w, b)
    ref.backward(up)
    want = (ref.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone())
    for bf16 in (False, True):
        x.grad = w.grad = b.grad = None
        y = _LayerNormTrain.apply(x, w, b, bf16)
        assert y.dtype == (torch.bfloat16 if bf16 else torch.float32)
        y.backward(up.to(y.dtype))
        tol = 2e-2 if bf16 else 2e-5
        assert torch.allclose(y.float(), want[0], rtol=tol, atol=tol)
        gt = 2e-2 if bf16 else 1e-4   # a bf16 incoming gradient carries 8 mantissa bits
        assert torch.allclose(x.grad, want[1], rtol=gt, atol=gt * float(want[1].abs().max()))
        assert torch.allclose(w.grad, want[2], rtol=gt, atol=gt * float(want[2].abs().max()))
        assert torch.allclose(b.grad, want[3], rtol=gt, atol=gt * float(want[3].abs().max()))


def test_flat_adam_and_clipping_match_torch():
    """FlatAdam (icm_grad_sumsq + icm_clip_coef + icm_adam_step) vs torch.optim.Adam + clip_grad_norm_ over three steps,
    with a pre-scale as after a SUM all-reduce over 4 ranks."""
    from compressai.training import FlatAdam

    torch.manual_seed(5)
    shapes = [(7,), (33, 5), (4, 3, 3, 3), (1,), (130,)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt = FlatAdam(ours, lr=1e-3)
    topt = torch.optim.Adam(ref, lr=1e-3)
    for step in range(3):
        opt.zero_grad()
        topt.zero_grad()
        gs = [torch.randn(s, device="cuda") * (50.0 if step == 0 else 0.01) for s in shapes]  # clipped on step 0 only
        for p, r, g in zip(ours, ref, gs):
            p.grad.copy_(g * 4.0)   # "sum over 4 ranks"
            r.grad = g.clone()
        v0 = [p._version for p in ours]
        opt.step(clip_max_norm=1.0, grad_pre_scale=0.25)
        norm = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        topt.step()
        assert abs(float(opt.grad_norm()) - float(norm)) <= 1e-4 * float(norm)
        assert all(p._version > v for p, v in zip(ours, v0))  # packed-weight caches key on the version counters
        for p, r in zip(ours, ref):
            assert torch.allclose(p.detach(), r.detach(), rtol=1e-5, atol=1e-6)


def test_training_step_matches_the_reference_step(gold):
    """One Trainer.step on the fixture's batch with the reference's draws replayed: loss terms within 1e-3 relative of the
    reference's CPU step (the row's bar), gradient norm likewise, probed gradients close, and the probed parameters after the
    step equal to the reference's (first Adam step: p - lr * g / (|g| + eps))."""
    from compressai.models._train import ReplayRng
    from compressai.training import Trainer
    from oracle import weights

    _strict_fp32()
    m = _stf()
    m.train_rng = ReplayRng()
    named = dict(m.named_parameters())
    tr = Trainer(m, lmbda=800.0, learning_rate=1e-5, aux_learning_rate=1e-4, clip_max_norm=1.0)
    x = weights.seeded_image((2, 3, 128, 128), seed=41).cuda()
    grads = {}
    probes = [k[5:] for k in gold.files if k.startswith("grad/")]
    def probe(n):
        def hook(p):
            if n not in grads:
                grads[n] = p.grad.detach().clone()
        return hook

    hooks = [named[n].register_post_accumulate_grad_hook(probe(n)) for n in probes]
    for step in range(2):
        torch.manual_seed(4242 + step)
        crit = tr.step(x)
        for k, g in (("loss", f"loss{step}"), ("bpp_loss", f"bpp{step}"), ("mse_loss", f"mse{step}"), ("aux_loss", f"aux{step}")):
            ref = float(gold[g])
            assert abs(crit[k].item() - ref) <= 1e-3 * abs(ref), (step, k, crit[k].item(), ref)
        assert abs(float(tr.optimizer.grad_norm()) - float(gold[f"norm{step}"])) <= 2e-3 * float(gold[f"norm{step}"])
        if step == 0:
            for h in hooks:
                h.remove()
            for n in probes:
                ref = torch.from_numpy(gold["grad/" + n])
                got = grads[n].cpu()
                assert float((got - ref).abs().max()) <= 2e-2 * float(ref.abs().max()) + 1e-7, n
                after, before = torch.from_numpy(gold["after/" + n]), torch.from_numpy(gold["before/" + n])
                lr = 1e-4 if n.endswith(".quantiles") else 1e-5
                d_ref, d_got = (after - before) / lr, (named[n].detach().cpu() - before) / lr
                sure = ref.abs() > 1e-3 * ref.abs().max()          # where the sign of the gradient is not in doubt
                if n.endswith(".quantiles"):
                    sure = torch.ones_like(sure)
                assert float((d_got - d_ref).abs()[sure].max()) <= 2e-2, n


def test_fused_and_unfused_training_forward_agree():
    from compressai.models._train import ReplayRng
    from oracle import weights

    _strict_fp32()
    m = _stf()
    m.train_rng = ReplayRng()
    x = weights.seeded_image((1, 3, 64, 128), seed=8).cuda()
    outs = []
    for fused in (True, False):
        m.train_fused = fused
        torch.manual_seed(77)
        o = m(x)
        outs.append(o)
    # the fused LayerNorm differs from torch's by fp32 rounding (~1e-6); after 12 blocks a few y values sit on the other side of a
    # rounding boundary, so y_hat (and with it a patch of x_hat, and that element's likelihood) may differ for a handful of
    # elements: compare the bulk
    ly0, ly1 = outs[0]["likelihoods"]["y"], outs[1]["likelihoods"]["y"]
    close = (ly0 - ly1).abs() <= 1e-3 * ly1.abs() + 1e-12
    assert float(close.float().mean()) >= 0.995
    d = (outs[0]["x_hat"] - outs[1]["x_hat"]).abs()
    assert float(d.mean()) <= 2e-4 and float((d > 5e-3).float().mean()) <= 0.01, (float(d.mean()), float(d.max()))


def test_eval_forward_still_runs_the_inference_kernels_after_training_steps():
    """Parameters move under the packed-weight cache (FlatAdam re-homes and updates them in place): the inference path must
    see the new values."""
    from compressai import _native
    from compressai.training import Trainer
    from oracle import weights

    m = _stf()
    x = weights.seeded_image((1, 3, 64, 64), seed=2).cuda()
    m.eval()
    before = m(x)["x_hat"].clone()
    m.train()
    tr = Trainer(m, lmbda=800.0, learning_rate=1e-3)
    tr.step(x)
    m.eval()
    n0 = _native.launch_count()
    after = m(x)["x_hat"]
    assert _native.launch_count() - n0 > 100
    assert not torch.equal(before, after)
    m.update(force=True)
    c = m.compress(x)
    d = m.decompress(c["strings"], c["shape"])
    assert torch.equal(d["x_hat"], m(x)["x_hat"].clamp(0, 1))


def test_cuda_graph_step_equals_eager_step():
    """Trainer(cuda_graph=True): the captured step replays forward, backward, clipping and both Adams.  With the randomness
    switched off (a TrainRng that returns zero noise and keep-masks of ones, drawn on the device so that it is capturable) the
    replayed trajectory must equal the eager one -- same kernels in the same order -- and the three warm-up steps the capture
    runs must leave no trace in the parameters, the Adam moments or the step counter."""
    from compressai.models._train import TrainRng
    from compressai.training import Trainer
    from oracle import weights

    class NoRandomness(TrainRng):
        def uniform(self, like):
            return torch.zeros_like(like)

        def keep_mask(self, like, keep):
            return like.new_ones((like.shape[0], 1, 1))

    _strict_fp32()
    x = weights.seeded_image((2, 3, 64, 64), seed=5).cuda()
    traj = []
    for graph in (False, True):
        m = _stf()
        m.train_rng = NoRandomness()
        tr = Trainer(m, lmbda=800.0, learning_rate=1e-4, cuda_graph=graph)
        losses = []
        for step in range(3):
            crit = tr.step(x)
            losses.append((float(crit["loss"].item()), float(crit["aux_loss"].item())))
        traj.append((losses, float(tr.optimizer._state[0].item()), tr.optimizer.param.detach().clone(), tr.aux_optimizer.param.detach().clone()))
    (le, te, pe, qe), (lg, tg, pg, qg) = traj
    assert te == tg == 3.0
    for (a, b), (c, d) in zip(le, lg):
        assert abs(a - c) <= 1e-4 * abs(a) and abs(b - d) <= 1e-4 * abs(b), (le, lg)
    assert le[2][0] < le[0][0]  # it trains
    # library backward kernels reduce with atomics, so gradients that are ~0 can change sign between two runs and Adam's first
    # steps move such an element by +-lr either way: compare the bulk, not the worst element
    d = (pe - pg).abs()
    print("graph vs eager: mean |dp| %.3g, fraction > 1e-5: %.4f, max %.3g" % (float(d.mean()), float((d > 1e-5).float().mean()), float(d.max())))
    assert float(d.mean()) <= 2e-6 and float((d > 1e-5).float().mean()) <= 0.02, (float(d.mean()), float((d > 1e-5).float().mean()))
    assert float((qe - qg).abs().mean()) <= 1e-5  # a quantile whose gradient is ~0 may take its +-lr step either way
