"""CPU tests of the host logic around the CUDA path:

* E9 -- the autograd halves of `LowerBound` / `ste_round` / `NonNegativeParametrizer`
  (/root/reference/compressai/ops/bound_ops.py:21-62, ops/ops.py:20-34, ops/parametrizers.py:23-49): the expected
  values below restate the reference rule; when /root/reference is present (the build container) the reference's own
  modules are imported unmodified and compared too.
* (f1) the evaluation driver against the reference's OWN `inference()` (compressai/utils/eval_model/__main__.py:96-139):
  tests/golden/eval_small.npz was recorded by executing that function's unmodified source on the reference STF
  (oracle/make_golden.py eval_golden) for a 100x150 image; here the driver is fed the reference's strings and padded
  reconstruction through a stand-in model, so its pad / crop / psnr / bpp arithmetic must reproduce the reference's
  numbers exactly, without a GPU.
"""
import os

import numpy as np
import pytest
import torch


def test_lower_bound_forward_and_gradient_rule():
    from compressai.ops import LowerBound

    lb = LowerBound(0.11)
    x = torch.tensor([-1.0, 0.05, 0.11, 0.2, 3.0, 0.1, 0.1], requires_grad=True)
    g = torch.tensor([1.0, -2.0, 0.5, 0.7, -1.0, 0.0, -0.0])
    y = lb(x)
    assert torch.equal(y.detach(), torch.tensor([0.11, 0.11, 0.11, 0.2, 3.0, 0.11, 0.11]))
    y.backward(g)
    # pass-through where x >= bound OR the gradient would move x towards the bound (grad < 0); zero otherwise
    assert torch.equal(x.grad, torch.tensor([0.0, -2.0, 0.5, 0.7, -1.0, 0.0, 0.0]))
    assert "bound" in dict(lb.named_buffers()) and lb.bound.shape == (1,)  # state_dict entry of the reference


def test_ste_round_is_round_with_identity_gradient():
    from compressai.ops import ste_round

    x = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 0.49999997, 1e6 + 0.5, -3.2], requires_grad=True)
    y = ste_round(x)
    assert torch.equal(y.detach(), torch.round(x.detach()))  # half-to-even, like torch.round
    g = torch.arange(1.0, 9.0)
    y.backward(g)
    assert torch.equal(x.grad, g)
    # the reference's formulation round(x) - x.detach() + x gives the same forward values on these inputs
    xr = x.detach()
    assert torch.equal(torch.round(xr) - xr + xr, y.detach())


def test_non_negative_parametrizer_values_and_gradient():
    from compressai.ops import NonNegativeParametrizer

    p = NonNegativeParametrizer(minimum=1e-6)
    pedestal = (2.0 ** -18) ** 2
    v = torch.tensor([0.0, 1e-4, 0.3, 2.0], requires_grad=True)
    out = p(v)
    bound = (1e-6 + pedestal) ** 0.5
    want = torch.maximum(v.detach(), torch.tensor(bound)) ** 2 - pedestal
    assert torch.allclose(out.detach(), want, rtol=0, atol=1e-12)
    out.sum().backward()
    # d/dv (max(v, b)^2): 2 v above the bound; below it the LowerBound rule blocks a positive gradient
    assert torch.allclose(v.grad, torch.tensor([0.0, 0.0, 0.6, 4.0]))
    assert torch.allclose(p.init(torch.tensor([0.25])), torch.sqrt(torch.tensor([0.25 + pedestal])))


@pytest.mark.skipif(not os.path.isdir("/root/reference/compressai/ops"), reason="reference tree not on this machine")
def test_ops_equal_the_reference_modules_on_random_inputs():
    import importlib.util

    def load(name):
        spec = importlib.util.spec_from_file_location("ref_" + name, f"/root/reference/compressai/ops/{name}.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    ref_bound, ref_ops = load("bound_ops"), load("ops")
    from compressai.ops import LowerBound, ste_round

    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(4096, generator=g) * 0.3
    grad = torch.randn(4096, generator=g)
    for bound in (0.0, 0.11, 1e-9):
        a, b = x0.clone().requires_grad_(), x0.clone().requires_grad_()
        ya, yb = LowerBound(bound)(a), ref_bound.LowerBound(bound)(b)
        ya.backward(grad); yb.backward(grad)
        assert torch.equal(ya, yb) and torch.equal(a.grad, b.grad)
    a, b = (x0 * 20).clone().requires_grad_(), (x0 * 20).clone().requires_grad_()
    ya, yb = ste_round(a), ref_ops.ste_round(b)
    ya.backward(grad); yb.backward(grad)
    assert torch.equal(ya, yb) and torch.equal(a.grad, b.grad)


# ------------------------------------------------------------------------------------------------ eval driver
class _ReplayModel:
    """Stands in for a codec: returns what the reference model returned when the fixture was recorded."""

    def __init__(self, gold):
        self.gold, self.seen = gold, {}

    def compress(self, x_padded):
        self.seen["x_padded"] = x_padded.clone()
        return {"strings": [[self.gold["y_string"].tobytes()], [self.gold["z_string"].tobytes()]], "shape": torch.Size(self.gold["shape"].tolist())}

    def decompress(self, strings, shape):
        assert list(shape) == self.gold["shape"].tolist()
        return {"x_hat": torch.from_numpy(self.gold["x_hat_padded"].copy())}


def test_eval_driver_reproduces_the_reference_inference_numbers(golden_dir):
    from compressai.utils import eval_model
    from oracle import weights

    gold = np.load(os.path.join(golden_dir, "eval_small.npz"))
    x = weights.seeded_image((3, 100, 150), seed=21)
    m = _ReplayModel(gold)
    rv = eval_model.inference(m, x)
    # padding: centred zeros to 128 x 192 (eval_model/__main__.py:103-115)
    xp = m.seen["x_padded"]
    assert list(xp.shape) == gold["x_padded_shape"].tolist() == [1, 3, 128, 192]
    assert torch.equal(xp[:, :, 14:114, 21:171], x.unsqueeze(0)) and float(xp.sum()) == pytest.approx(float(x.sum()), rel=1e-6)
    assert float(xp[:, :, :14].abs().sum()) == 0.0 and float(xp[:, :, :, :21].abs().sum()) == 0.0
    # crop + metrics (:126-139): bpp over the UNPADDED pixels, psnr against the unpadded image
    assert torch.equal(rv["x_hat"], torch.from_numpy(gold["x_hat"]))
    assert rv["bpp"] == float(gold["bpp"]) == (gold["y_string"].size + gold["z_string"].size) * 8.0 / (100 * 150)
    assert rv["psnr"] == pytest.approx(float(gold["psnr"]), abs=1e-9)
