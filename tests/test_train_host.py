"""Training row on the CPU: the training-mode forward's transforms + randomness against the reference's
own modules (tests/golden/train_step.npz, oracle/make_golden.py train), the loss arithmetic of train.py:53-74, and the
bucketed gradient all-reduce with a world-size-2 gloo group.  The fused CUDA pieces are covered by tests/test_gpu_train.py."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "train_step.npz"))


def _model():
    from compressai.zoo import models
    from oracle import stf_ref, weights

    m = models["stf"]()
    m.load_state_dict(weights.seeded_state_dict(stf_ref.template_state_dict(), seed=0, stress=True), strict=False)
    return m.train()


def test_training_forward_and_gradients_equal_the_reference_on_cpu(gold):
    """Same weights, same batch, the reference's draws replayed (DropPath masks and noise in its order and shapes): the
    loss terms agree to 1e-6 relative and the probed gradients to fp32 rounding.  The Gaussian stage is evaluated with the
    PyTorch expressions here (fused=False); on a GPU the fused kernels are compared with exactly these."""
    from compressai.models._train import ReplayRng, stf_train_forward
    from compressai.training import RateDistortionLoss
    from oracle import weights

    m = _model()
    x = weights.seeded_image((2, 3, 128, 128), seed=41)
    torch.manual_seed(4242)
    out = stf_train_forward(m, x, rng=ReplayRng(), fused=False)
    crit = RateDistortionLoss(800.0)(out, x)
    for k, g in (("loss", "loss0"), ("bpp_loss", "bpp0"), ("mse_loss", "mse0")):
        assert abs(crit[k].item() - float(gold[g])) <= 1e-6 * abs(float(gold[g])), (k, crit[k].item(), float(gold[g]))
    assert abs(torch.log(out["likelihoods"]["y"]).sum().item() - float(gold["y_lik_logsum0"])) <= 1e-6 * abs(float(gold["y_lik_logsum0"]))
    crit["loss"].backward()
    named = dict(m.named_parameters())
    for key in gold.files:
        if key.startswith("grad/"):
            ref = gold[key]
            got = named[key[5:]].grad.numpy()
            assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-12) + 1e-9, key
    # (the fixture's aux loss is taken after the main optimizer step, train.py:210-213: compared in tests/test_gpu_train.py)


def test_fused_training_path_refuses_cpu_tensors():
    from compressai._native import NativeError
    from oracle import weights

    m = _model()
    with pytest.raises(NativeError):
        m(weights.seeded_image((1, 3, 64, 64), seed=1))  # default: fused kernels, CUDA only


def test_rate_distortion_loss_formula():
    from compressai.training import RateDistortionLoss

    x = torch.rand(2, 3, 8, 8)
    out = {"x_hat": torch.rand(2, 3, 8, 8), "likelihoods": {"y": torch.rand(2, 4, 2, 2) * 0.9 + 0.05, "z": torch.rand(2, 2, 1, 1) * 0.9 + 0.05}}
    c = RateDistortionLoss(0.01)(out, x)
    bpp = sum(float(torch.log(l).sum()) for l in out["likelihoods"].values()) / (-math.log(2) * 2 * 8 * 8)
    mse = float(((out["x_hat"] - x) ** 2).mean())
    assert abs(c["bpp_loss"].item() - bpp) < 1e-6 and abs(c["mse_loss"].item() - mse) < 1e-7
    assert abs(c["loss"].item() - (0.01 * mse + bpp)) < 1e-6


def _bucket_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from compressai.training import GradientBuckets

        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
        unused = torch.nn.Parameter(torch.ones(5))  # never receives a gradient: its bucket must still be reduced by finish()
        params = list(net.parameters()) + [unused]
        order, off = [], 0
        for p in reversed(params):
            order.append((p, off, p.numel()))
            off += (p.numel() + 3) // 4 * 4
        flat = torch.zeros(off)
        for p, o, n in order:
            p.grad = flat[o:o + n].view_as(p)
        import copy

        twin = copy.deepcopy(net)  # same weights, no hooks: this rank's own gradients (the bucketed buffer is reduced in place,
        gb = GradientBuckets(flat, order, bucket_bytes=1024)  # possibly while backward is still running)
        assert len(gb.buckets) >= 2 and gb.buckets[0][0] == 0 and gb.buckets[-1][1] == off
        results = []
        for step in range(2):  # twice: the arrival counters must re-arm
            flat.zero_()
            g = torch.Generator().manual_seed(100 * step + rank)
            x = torch.randn(8, 16, generator=g)
            own = torch.autograd.grad(twin(x).square().sum(), list(twin.parameters()))
            local = torch.zeros_like(flat)
            by_param = dict(zip(net.parameters(), own))
            for p, o, n in order:
                if p in by_param:
                    local[o:o + n] = by_param[p].reshape(-1)
            net(x).square().sum().backward()
            gb.finish()
            results.append((local, flat.clone()))
        gb.remove()
        q.put((rank, [(a.numpy(), b.numpy()) for a, b in results]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_bucketed_gradient_allreduce_with_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step in range(2):
        total = res[0][step][0] + res[1][step][0]
        assert np.abs(total).max() > 0
        for r in (0, 1):
            assert np.allclose(res[r][step][1], total, rtol=1e-6, atol=1e-7)
        assert not np.allclose(res[0][step][0], res[1][step][0])
