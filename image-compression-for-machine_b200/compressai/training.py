"""Rate-distortion training step for the codec (BASELINE.json configs[4]).

Mirrors the reference's training script (train.py): `RateDistortionLoss` (:44-76, with the model's "x_hat" key -- the
shipped script reads "decompressedImage", which STF does not emit), `configure_optimizers`
(:105-169: every parameter but `*.quantiles` in the main Adam, the quantiles in the auxiliary one) and the body of
`train_one_epoch` (:172-214) as `train_step`.  What is B200-specific:

  * both optimizers keep their parameters, gradients and Adam moments in ONE flat fp32 buffer each (`FlatAdam`);
    gradient-norm clipping and the Adam update are two passes of csrc/train.cu over those buffers (28 B per
    parameter, HBM-bound) with the clip coefficient left on the device -- no host synchronisation inside a step;
  * data parallelism is one process per GPU; the flat gradient buffer is cut into buckets in reverse registration
    order and each bucket is all-reduced (NCCL over NVLink) as soon as autograd has produced its last gradient, so the
    399 MB of gradients travel while the rest of backward runs (`GradientBuckets`).  The averaging is folded into the
    Adam pass (grad scale 1 / world).
"""
import math

import torch
import torch.nn as nn

from compressai._native import NativeError, check, lib, stream_ptr


class RateDistortionLoss(nn.Module):
    """loss = lmbda * MSE(x_hat, x) + bpp,  bpp = sum over likelihood tensors of sum(log p) / (-ln 2 * N*H*W)  (train.py:53-74)."""

    def __init__(self, lmbda=1e-2):
        super().__init__()
        self.lmbda = lmbda

    def forward(self, output, target):
        N, _, H, W = target.shape
        denom = -math.log(2) * N * H * W
        out = {"bpp_loss": sum(torch.log(l).sum() / denom for l in output["likelihoods"].values()),
               "mse_loss": torch.nn.functional.mse_loss(output["x_hat"], target)}
        out["loss"] = self.lmbda * out["mse_loss"] + out["bpp_loss"]
        return out


class FlatAdam:
    """torch.optim.Adam (defaults: betas (0.9, 0.999), eps 1e-8, no weight decay) over parameters re-homed into one flat buffer.

    After construction every parameter's `.data` and `.grad` are views into `self.param` / `self.grad`, in REVERSE order of
    the given list (roughly the order in which backward produces the gradients), so that buckets of the gradient buffer
    complete early and contiguously."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = list(params)
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise NativeError("FlatAdam runs on CUDA only (csrc/train.cu); move the model to a CUDA device first")
        self.lr, self.betas, self.eps, self.t = float(lr), betas, float(eps), 0
        order = list(reversed(self.params))
        self.offsets, n = {}, 0
        for p in order:
            self.offsets[id(p)] = n
            n += (p.numel() + 3) // 4 * 4  # 16-byte aligned views
        self.n = n
        self.param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.param)
        self.exp_avg = torch.zeros_like(self.param)
        self.exp_avg_sq = torch.zeros_like(self.param)
        self._scratch = torch.zeros(4, dtype=torch.float32, device=dev)  # [sumsq, coef, norm, -]
        self._state = torch.zeros(4, dtype=torch.float32, device=dev)    # [t, 1 - b1^t, 1 / sqrt(1 - b2^t), -]: advanced on the device
        with torch.no_grad():
            for p in self.params:
                if p.dtype != torch.float32:
                    raise NativeError("FlatAdam: fp32 master parameters expected")
                o = self.offsets[id(p)]
                view = self.param[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[o:o + p.numel()].view_as(p)

    def order(self):
        """Parameters in buffer order with their (offset, numel)."""
        return [(p, self.offsets[id(p)], p.numel()) for p in reversed(self.params)]

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:  # a hook-free re-binding in case autograd replaced a .grad (it accumulates in place when one exists)
            o = self.offsets[id(p)]
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view_as(p)

    def step(self, clip_max_norm=0.0, grad_pre_scale=1.0):
        """One Adam update.  clip_max_norm > 0: clip_grad_norm_ over this optimizer's gradients first (train.py:208-209).
        grad_pre_scale: factor the buffer's gradients still need (1 / world after a SUM all-reduce)."""
        L, st = lib(), stream_ptr()
        self.t += 1
        s = self._scratch
        coef_ptr = None
        if clip_max_norm and clip_max_norm > 0:
            s[:1].zero_()
            check(L.icm_grad_sumsq(self.grad.data_ptr(), self.n, s.data_ptr(), st), "icm_grad_sumsq")
            check(L.icm_clip_coef(s.data_ptr(), float(clip_max_norm), float(grad_pre_scale), s.data_ptr() + 4, s.data_ptr() + 8, st), "icm_clip_coef")
            coef_ptr, host_scale = s.data_ptr() + 4, 1.0
        else:
            host_scale = float(grad_pre_scale)
        check(L.icm_adam_step(self.param.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.n,
                              self.lr, self.betas[0], self.betas[1], self.eps, 0, self._state.data_ptr(), coef_ptr, host_scale, st), "icm_adam_step")
        self.touch()

    def touch(self):
        """The parameters were updated in place behind autograd's back: bump their version counters (the inference engine's
        packed-weight cache keys on them).  Host-only; called after every step, eager or replayed from a CUDA graph."""
        for p in self.params:
            torch.autograd.graph.increment_version(p)

    def snapshot(self):
        return (self.param.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self._state.clone(), self.t)

    def restore(self, snap):
        self.param.copy_(snap[0]); self.exp_avg.copy_(snap[1]); self.exp_avg_sq.copy_(snap[2]); self._state.copy_(snap[3])
        self.t = snap[4]

    def grad_norm(self):
        """The (pre-scaled) gradient norm of the last clipped step (device scalar)."""
        return self._scratch[2]


def configure_optimizers(net, learning_rate=1e-5, aux_learning_rate=1e-4):
    """train.py:105-169 as it is meant for this codec: main Adam over every parameter whose name does not end in
    ".quantiles", auxiliary Adam over the quantiles (the shipped filter `'human' in name` matches no STF parameter)."""
    named = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
    main = [p for n, p in named if not n.endswith(".quantiles")]
    aux = [p for n, p in named if n.endswith(".quantiles")]
    return FlatAdam(main, learning_rate), FlatAdam(aux, aux_learning_rate)


class GradientBuckets:
    """Bucketed, overlapped all-reduce of a flat gradient buffer (one process per GPU, torch.distributed).

    `order` = [(parameter, offset, numel)] in buffer order; the buffer is cut into contiguous buckets of about
    `bucket_bytes`; a post-accumulate-grad hook per parameter counts arrivals and launches the bucket's all-reduce
    (SUM; the 1 / world lands in the optimizer pass) the moment its last gradient exists.  `finish()` launches whatever
    is left (parameters that received no gradient) and waits for every bucket."""

    def __init__(self, flat_grad, order, bucket_bytes=32 << 20, group=None):
        import torch.distributed as dist

        self.dist, self.group, self.flat = dist, group, flat_grad
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets, self._of = [], {}
        lo, count = 0, 0
        limit = max(1, bucket_bytes // 4)
        end = 0
        for p, off, n in order:
            end = off + (n + 3) // 4 * 4
            self._of[id(p)] = len(self.buckets)
            count += 1
            if end - lo >= limit:
                self.buckets.append([lo, end, count])
                lo, count = end, 0
        if count:
            self.buckets.append([lo, end, count])
        self._pending = [b[2] for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._work = []
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p, _, _ in order] if self.world > 1 else []

    def _launch(self, b):
        lo, hi, _ = self.buckets[b]
        self._launched[b] = True
        self._work.append(self.dist.all_reduce(self.flat[lo:hi], op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _hook(self, p):
        b = self._of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0 and not self._launched[b]:
            self._launch(b)

    def finish(self):
        if self.world > 1:
            for b in range(len(self.buckets)):
                if not self._launched[b]:
                    self._launch(b)
            for w in self._work:
                w.wait()
        self._work = []
        self._pending = [b[2] for b in self.buckets]
        self._launched = [False] * len(self.buckets)

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


class Trainer:
    """State of the training loop of train.py:172-214 for one process (= one GPU)."""

    def __init__(self, net, lmbda=800.0, learning_rate=1e-5, aux_learning_rate=1e-4, clip_max_norm=1.0, bucket_bytes=32 << 20, autocast=None,
                 cuda_graph=False):
        import torch.distributed as dist

        # cuda_graph: the whole step (forward, backward, bucketed all-reduces, clipping, both Adams) is captured once per
        # batch shape and replayed -- a step is ~5 000 small launches, 90 ms of host time against 20-40 ms of GPU time.
        self.cuda_graph = bool(cuda_graph)
        self._graph = self._graph_shape = self._static_x = self._static_out = None
        self.net = net
        self.criterion = RateDistortionLoss(lmbda)
        self.optimizer, self.aux_optimizer = configure_optimizers(net, learning_rate, aux_learning_rate)
        self.clip_max_norm = clip_max_norm
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.buckets = GradientBuckets(self.optimizer.grad, self.optimizer.order(), bucket_bytes)
        self.autocast = autocast  # None / torch.bfloat16: dtype of the transforms' GEMMs (parameters stay fp32)

    def step(self, x):
        """optimizer.zero_grad .. aux_optimizer.step of train.py:196-214 on one batch shard; returns the loss terms
        (device scalars; nothing is read back here)."""
        if not self.cuda_graph:
            return self._step(x)
        if self._graph is None or self._graph_shape != tuple(x.shape):
            self._capture(x)
        self._static_x.copy_(x, non_blocking=True)
        self._graph.replay()
        self.optimizer.t += 1
        self.aux_optimizer.t += 1
        self.optimizer.touch()
        self.aux_optimizer.touch()
        return self._static_out

    def _capture(self, x):
        """Warm up on a side stream (allocator, cuDNN plans, NCCL), roll the optimizer state back, capture one step."""
        self._static_x = x.clone()
        snaps = (self.optimizer.snapshot(), self.aux_optimizer.snapshot())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._step(self._static_x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_out = self._step(self._static_x)
        self.optimizer.restore(snaps[0])
        self.aux_optimizer.restore(snaps[1])
        self._graph_shape = tuple(x.shape)

    def _step(self, x):
        import torch.distributed as dist

        net = self.net
        self.optimizer.zero_grad()
        self.aux_optimizer.zero_grad()
        if self.autocast is not None:
            with torch.autocast("cuda", dtype=self.autocast):
                out = net(x)
        else:
            out = net(x)
        out["x_hat"] = out["x_hat"].float()
        crit = self.criterion(out, x)
        crit["loss"].backward()
        self.buckets.finish()
        self.optimizer.step(self.clip_max_norm, grad_pre_scale=1.0 / self.world)
        aux = net.aux_loss()
        aux.backward()
        if self.world > 1:  # 576 values: one tiny all-reduce
            dist.all_reduce(self.aux_optimizer.grad, op=dist.ReduceOp.SUM)
        self.aux_optimizer.step(0.0, grad_pre_scale=1.0 / self.world)
        crit["aux_loss"] = aux.detach()
        return crit


def train_step(trainer, x):
    return trainer.step(x)
