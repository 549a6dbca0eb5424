"""Hyper-prior + channel-conditional slice loop + rANS shared by the STF and WACNN codecs.

Both reference models run the same control flow after their analysis transform (models/stf.py:596-645,671-785;
models/cnn.py:141-189,210-332): h_a -> EntropyBottleneck -> h_mean_s / h_scale_s -> per slice { cc_mean, cc_scale
conv stacks -> build_indexes / quantise (or likelihood, or rANS decode) -> LRP conv stack } -> g_s.  They differ
in the latent width (384 vs 320), the number of 32-channel slices (12 vs 10), the number of support slices
(6 vs 5) and in g_a / g_s, which the sub-classes provide as `_analysis` / `_synthesis`.

Layout: activations are channels-last.  The growing concatenations of the context model are two persistent
support buffers [B*h*w, M + 32*(max_support(+1))] whose channel slots are written in place by the kernels.
"""
import torch

from compressai import ans
from compressai._native import ACT_HALF_TANH, NULL_VIEW, OUT_BF16, OUT_F32, NativeError, View, check, lib, stream_ptr, view_bcp

from ._engine import Engine
from .base import CompressionModel
from .utils import update_registered_buffers

SLICE = 32  # channels per slice in both reference models


class ChannelContextCodec(CompressionModel):
    latent_channels = 384   # M
    hyper_channels = 192    # channels of z
    num_slices = 12
    max_support_slices = 6
    scale_table_fn = None   # set by the sub-class module (get_scale_table)

    # ---------------------------------------------------------------------------------- reference API
    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = type(self).scale_table_fn()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        return updated

    def load_state_dict(self, state_dict, strict=False):
        update_registered_buffers(self.gaussian_conditional, "gaussian_conditional",
                                  ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        return super().load_state_dict(state_dict, strict=strict)

    @classmethod
    def from_state_dict(cls, state_dict):
        net = cls()
        net.load_state_dict(state_dict)
        return net

    def _check_input(self, x):
        if not x.is_cuda:
            raise NativeError(f"{type(self).__name__} runs on CUDA only (no CPU fallback); move the model and input to a CUDA device")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected an image batch [B, 3, H, W]")
        if x.shape[2] % 64 or x.shape[3] % 64:
            raise ValueError("H and W must be multiples of 64 (pad like the reference eval does, eval_model/__main__.py:103-115)")

    # Grouped launches (icm_conv2d_grouped).  Stacks that the reference's data flow leaves independent share a launch:
    # h_mean_s || h_scale_s, cc_mean[i] || cc_scale[i], and -- because only the FIRST max_support_slices decoded slices are
    # ever used as support (stf.py:612, cnn.py:162) -- every stack of the slices i >= max_support_slices.  Those slices are
    # then quantised / decoded / LRP-corrected in one step as well (2 * (num_slices - max_support) conv stacks, one entropy
    # launch, one rANS step and one grouped LRP stack instead of num_slices - max_support rounds of each).  Per output
    # element the arithmetic is the ungrouped kernel's, so symbols, indexes and bit-strings are unchanged.
    grouped = True

    def _tail_grouped(self):
        """Slices >= max_support in one step: needs the LRP stacks' [support | y_hat_i] split on a 64-channel chunk boundary."""
        return self.grouped and self.num_slices > self.max_support_slices + 1 and (self.latent_channels + SLICE * self.max_support_slices) % 64 == 0

    def _hyper_synthesis(self, z_hat_bf16, B, zh, zw):
        """h_mean_s / h_scale_s (stf.py:604-605) written straight into channels [0, M) of the support buffers:
        sup[0] = mean support, sup[1] = scale support, [2, B*P, pitch] with one 32-channel slot per slice behind the M
        hyper-prior channels (all num_slices slots when the tail slices are grouped, else max_support + 1)."""
        e = self._engine
        P = 16 * zh * zw
        M = self.latent_channels
        slots = self.num_slices if self._tail_grouped() else self.max_support_slices + 1
        sup = torch.zeros((2, B * P, M + SLICE * slots), dtype=torch.bfloat16, device=z_hat_bf16.device)
        if self.grouped:
            e.conv_stack_group(z_hat_bf16, B, zh, zw, [self.h_mean_s, self.h_scale_s], B, [0, 0], final_dtype=OUT_BF16,
                               final_out=sup, final_out_group_stride=sup.shape[1] * sup.shape[2])
        else:
            e.conv_stack(z_hat_bf16, B, zh, zw, self.h_mean_s, final_dtype=OUT_BF16, final_out=sup[0])
            e.conv_stack(z_hat_bf16, B, zh, zw, self.h_scale_s, final_dtype=OUT_BF16, final_out=sup[1])
        return sup[0], sup[1]

    def _slice_loop(self, mode, B, h, w, mean_sup, scale_sup, y=None, decoder=None, y_lik=None):
        """The channel-conditional slice loop (stf.py:611-631, 703-726, 754-776).

        mode "compress": returns (y_hat, symbols, indexes) with symbols/indexes int32 [B, M*P] in stream order;
        mode "forward":  fills y_lik (NCHW fp32) and returns (y_hat, None, None);
        mode "decompress": symbols come from `decoder`."""
        e, gc = self._engine, self.gaussian_conditional
        dev = mean_sup.device
        P = h * w
        L = lib()
        st = stream_ptr()
        M, Z, S, N = self.latent_channels, SLICE, self.max_support_slices, self.num_slices
        pitch = mean_sup.shape[-1]
        sup = None
        if self.grouped:  # both planes of the [2, B*P, pitch] buffer _hyper_synthesis made
            assert scale_sup.data_ptr() == mean_sup.data_ptr() + B * P * pitch * 2 and scale_sup.shape[-1] == pitch
            sup = torch.as_strided(mean_sup, (2 * B * P, pitch), (pitch, 1))
        tail = self._tail_grouped() and pitch == M + Z * N
        y_hat = torch.empty((B * P, M), dtype=torch.float32, device=dev)
        table = gc.scale_table_device(dev)
        sym = idx = None
        if mode == "compress":
            sym = torch.empty((B, M * P), dtype=torch.int32, device=dev)
            idx = torch.empty_like(sym)
        elif mode == "decompress":
            gct = gc.device_tables()
        # steps: (first slice, number of slices)
        steps = [(i, 1) for i in range(S if tail else N)] + ([(S, N - S)] if tail else [])
        for i, n in steps:
            k = min(i, S)
            C = Z * n
            if n > 1:    # every cc stack of the slices >= S: same support, 2n groups
                musc = torch.empty((B * P, 2 * C), dtype=torch.float32, device=dev)
                e.conv_stack_group(sup, B, h, w, list(self.cc_mean_transforms[i:]) + list(self.cc_scale_transforms[i:]), 2 * B,
                                   [0] * n + [B] * n, final_out=musc, final_out_group_stride=Z, cin=M + Z * k)
                v_mu, v_sc = view_bcp(musc, B, C, P), view_bcp(musc, B, C, P, C)
            elif self.grouped:
                musc, _, _ = e.conv_stack_group(sup, B, h, w, [self.cc_mean_transforms[i], self.cc_scale_transforms[i]], 2 * B, [0, B],
                                                cin=M + Z * k)
                v_mu, v_sc = view_bcp(musc[0], B, C, P), view_bcp(musc[1], B, C, P)
            else:
                mu, _, _ = e.conv_stack(mean_sup, B, h, w, self.cc_mean_transforms[i])
                sc, _, _ = e.conv_stack(scale_sup, B, h, w, self.cc_scale_transforms[i])
                v_mu, v_sc = view_bcp(mu, B, C, P), view_bcp(sc, B, C, P)
            v_hat = view_bcp(y_hat, B, C, P, Z * i)
            # pre-LRP ŷ_i, input of lrp_transforms[i]: slot i when every slice has one, else the slot behind the support slices
            v_slot = view_bcp(mean_sup, B, C, P, M + Z * (i if tail else k))
            if mode == "compress":
                check(L.icm_gc_quantize_index(view_bcp(y, B, C, P, Z * i), v_mu, v_sc, B, C, P, table.data_ptr(), table.numel(),
                                              gc._scale_bound_f, sym.data_ptr(), idx.data_ptr(), M * P, Z * P * i,
                                              v_hat, v_slot, NULL_VIEW, st), "icm_gc_quantize_index")
            elif mode == "forward":
                v_lik = View(y_lik.data_ptr() + Z * i * P * 4, M * P, P, 1)
                check(L.icm_gc_likelihood(view_bcp(y, B, C, P, Z * i), v_mu, v_sc, B, C, P, gc._scale_bound_f,
                                          gc.likelihood_bound if gc.use_likelihood_bound else 0.0, v_hat, v_lik, v_slot, NULL_VIEW, st),
                      "icm_gc_likelihood")
            else:
                idx_s = torch.empty((B, C * P), dtype=torch.int32, device=dev)
                sym_s = torch.empty_like(idx_s)
                check(L.icm_gc_build_indexes(v_sc, B, C, P, table.data_ptr(), table.numel(), gc._scale_bound_f,
                                             idx_s.data_ptr(), C * P, 0, st), "icm_gc_build_indexes")
                decoder.decode_step(gct, idx_s, out=sym_s)
                check(L.icm_gc_dequantize(sym_s.data_ptr(), C * P, 0, v_mu, B, C, P, v_hat, v_slot, NULL_VIEW, st), "icm_gc_dequantize")
            if n > 1:
                lrp = torch.empty((B * P, C), dtype=torch.float32, device=dev)
                e.conv_stack_group(mean_sup, B, h, w, list(self.lrp_transforms[i:]), B, [0] * n, tail_channels=[M + Z * (i + g) for g in range(n)],
                                   final_act=ACT_HALF_TANH, final_out=lrp, final_out_group_stride=Z, cin=M + Z * (k + 1))
            else:
                tc = [M + Z * i] if (tail and i >= S) else None
                if tc is not None:
                    lrp, _, _ = e.conv_stack_group(mean_sup, B, h, w, [self.lrp_transforms[i]], B, [0], tail_channels=tc,
                                                   final_act=ACT_HALF_TANH, cin=M + Z * (k + 1))
                    lrp = lrp[0]
                else:
                    lrp, _, _ = e.conv_stack(mean_sup, B, h, w, self.lrp_transforms[i], final_act=ACT_HALF_TANH, cin=M + Z * (k + 1))
            keep = i < S and n == 1  # only the first max_support_slices decoded slices are ever used as support (stf.py:612, cnn.py:162)
            check(L.icm_add_lrp(v_hat, view_bcp(lrp, B, C, P), B, C, P,
                                v_slot if keep else NULL_VIEW, view_bcp(scale_sup, B, C, P, M + Z * i) if keep else NULL_VIEW, st),
                  "icm_add_lrp")
        return y_hat, sym, idx

    def _hyper_analysis(self, y, B, h, w):
        e = self._engine
        z, zh, zw = e.conv_stack(e.cast_bf16(y), B, h, w, self.h_a)  # fp32 [B*zh*zw, 192]
        return z, zh, zw

    # ---------------------------------------------------------------------------------- forward / codec
    train_rng = None     # a compressai.models._train.TrainRng; None = draws on the device
    train_fused = True   # fused Gaussian-stage kernels (csrc/train.cu) in the training-mode forward

    def forward(self, x):
        """stf.py:582-645: {"x_hat" (unclamped), "likelihoods": {"y", "z"}} in NCHW.  Eval mode: the CUDA inference kernels,
        no autograd.  Training mode (noise quantisation, DropPath, gradients): models/_train.py."""
        if self.training:
            return self._train_forward(x)
        with torch.no_grad():
            return self._eval_forward(x)

    def _train_forward(self, x):
        raise NativeError(f"{type(self).__name__}: the training-mode forward is built for the STF codec only (BASELINE.json configs[4]); call .eval()")

    def _eval_forward(self, x):
        self._check_input(x)
        eb = self.entropy_bottleneck
        B = x.shape[0]
        y, h, w = self._analysis(x)
        z, zh, zw = self._hyper_analysis(y, B, h, w)
        Pz, P, Zc = zh * zw, h * w, self.hyper_channels
        z_hat = torch.empty((B * Pz, Zc), dtype=torch.bfloat16, device=x.device)
        z_lik = torch.empty((B, Zc, zh, zw), dtype=torch.float32, device=x.device)
        check(lib().icm_eb_process(1, view_bcp(z, B, Zc, Pz), B, Zc, Pz, eb.packed_params().data_ptr(),
                                   eb.likelihood_bound if eb.use_likelihood_bound else 0.0, None, None, NULL_VIEW,
                                   view_bcp(z_hat, B, Zc, Pz), View(z_lik.data_ptr(), Zc * Pz, Pz, 1), stream_ptr()), "icm_eb_process")
        mean_sup, scale_sup = self._hyper_synthesis(z_hat, B, zh, zw)
        y_lik = torch.empty((B, self.latent_channels, h, w), dtype=torch.float32, device=x.device)
        y_hat, _, _ = self._slice_loop("forward", B, h, w, mean_sup, scale_sup, y=y, y_lik=y_lik)
        x_hat = self._synthesis(y_hat, B, h, w, clamp=False)
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}}

    # ---------------------------------------------------------------------------------- micro-batching
    # The two rANS coders are latency-bound (one warp per image stream), the transforms are throughput-bound.
    # A batch is therefore split into `micro_batches` parts that run on their own CUDA streams, so that the
    # coder of one part overlaps the convolutions of another (the conv kernel's dynamic tile scheduler
    # tolerates SMs that are held by coder CTAs).  Results are identical to the unsplit run (kernels are batch-invariant).
    micro_batches = 2
    micro_batch_min = 16  # only split batches at least this large

    def _part_ranges(self, B):
        n = self.micro_batches if B >= self.micro_batch_min else 1
        n = max(1, min(n, B))
        base, extra = divmod(B, n)
        out, lo = [], 0
        for i in range(n):
            hi = lo + base + (1 if i < extra else 0)
            out.append((lo, hi))
            lo = hi
        return out

    def _run_parts(self, fns, coder_streams):
        """Run the callables on side streams (one each), joined back into the current stream."""
        if len(fns) == 1:
            return [fns[0]()]
        cur = torch.cuda.current_stream()
        pool = getattr(self, "_side_streams", None)
        if pool is None or len(pool) < len(fns) or pool[0].device != cur.device:
            pool = [torch.cuda.Stream(device=cur.device) for _ in fns]
            self._side_streams = pool
        outs = []
        for st, fn in zip(pool, fns):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs.append(fn())
        for st in pool[:len(fns)]:
            cur.wait_stream(st)
        return outs

    @staticmethod
    def _hand_over(t):
        """A tensor produced on a side stream is consumed on the current stream from now on."""
        if isinstance(t, torch.Tensor):
            t.record_stream(torch.cuda.current_stream())
        return t

    def _compress_transforms(self, x):
        """The throughput-bound half of compress(): analysis, hyper-prior, slice loop.  Returns the symbol / index planes of y and
        z in stream order (int32 [B, n]) and the z shape; asynchronous, no host logic that depends on device data."""
        eb = self.entropy_bottleneck
        B = x.shape[0]
        y, h, w = self._analysis(x)
        z, zh, zw = self._hyper_analysis(y, B, h, w)
        Pz, Zc = zh * zw, self.hyper_channels
        z_sym = torch.empty((B, Zc * Pz), dtype=torch.int32, device=x.device)
        z_idx = torch.empty_like(z_sym)
        z_hat = torch.empty((B * Pz, Zc), dtype=torch.bfloat16, device=x.device)
        check(lib().icm_eb_process(0, view_bcp(z, B, Zc, Pz), B, Zc, Pz, eb.packed_params().data_ptr(), 0.0,
                                   z_sym.data_ptr(), z_idx.data_ptr(), NULL_VIEW, view_bcp(z_hat, B, Zc, Pz), NULL_VIEW, stream_ptr()),
              "icm_eb_process")
        mean_sup, scale_sup = self._hyper_synthesis(z_hat, B, zh, zw)
        _, sym, idx = self._slice_loop("compress", B, h, w, mean_sup, scale_sup, y=y)
        return sym, idx, z_sym, z_idx, zh, zw

    def _compress_encode(self, sym, idx, z_sym, z_idx, worst_case=False):
        """The latency-bound half: both rANS encoders, device-resident streams ((packed uint8, sizes int32[B+1]) each)."""
        z_str = ans.encode_streams(self.entropy_bottleneck.device_tables(), z_sym, z_idx, return_device="async", worst_case=worst_case)
        y_str = ans.encode_streams(self.gaussian_conditional.device_tables(), sym, idx, return_device="async", worst_case=worst_case)
        return y_str, z_str

    def _compress_part(self, x, phase=None, worst_case=False):
        """compress() of one micro-batch, fully asynchronous: device-resident streams.  `phase("begin"/"end")`
        brackets the throughput-bound part (transforms + slice loop); the rANS encoders run after "end"."""
        if phase:
            phase("begin")
        sym, idx, z_sym, z_idx, zh, zw = self._compress_transforms(x)
        if phase:
            phase("end")
        y_str, z_str = self._compress_encode(sym, idx, z_sym, z_idx, worst_case)
        return {"y": y_str, "z": z_str, "shape": (zh, zw), "retry": (sym, idx, z_sym, z_idx)}

    @torch.no_grad()
    def compress(self, x, device_strings=False):
        """stf.py:671-732.  Returns {"strings": [[y_0..y_{B-1}], [z_0..z_{B-1}]], "shape": (h/4, w/4)}.

        device_strings=True keeps the streams on the GPU: "strings" is then [[(packed_y, sizes_y), ...], [(packed_z,
        sizes_z), ...]] with one (uint8, int32) CUDA tensor pair per micro-batch and no host synchronisation;
        decompress() accepts that form as is."""
        self._check_input(x)
        B = x.shape[0]
        parts = self._part_ranges(B)
        outs = self._run_parts([(lambda lo=lo, hi=hi: self._compress_part(x[lo:hi])) for lo, hi in parts], coder_streams=parts[0][1] - parts[0][0])
        shape = torch.Size(outs[0]["shape"])
        for o in outs:
            for k in ("y", "z"):
                self._hand_over(o[k][0]); self._hand_over(o[k][1])
        if device_strings:
            return {"strings": [[o["y"] for o in outs], [o["z"] for o in outs]], "shape": shape}
        gc_t, eb_t = self.gaussian_conditional.device_tables(), self.entropy_bottleneck.device_tables()
        y_strings, z_strings = [], []
        for o in outs:
            sym, idx, z_sym, z_idx = o["retry"]
            y_strings += ans.strings_to_host(*o["y"], retry=lambda: ans.encode_streams(gc_t, sym, idx))
            z_strings += ans.strings_to_host(*o["z"], retry=lambda: ans.encode_streams(eb_t, z_sym, z_idx))
        return {"strings": [y_strings, z_strings], "shape": shape}

    def _decompress_part(self, y_str, z_str, B, zh, zw, on_device, decoders=None):
        y_hat, decs = self._decode_part(y_str, z_str, B, zh, zw, on_device, decoders)
        return self._synthesis(y_hat, B, 4 * zh, 4 * zw, clamp=True), decs

    def _decode_part(self, y_str, z_str, B, zh, zw, on_device, decoders=None):
        """The latency-bound half of decompress(): z decode, hyper-synthesis, the slice loop around the y decoder."""
        eb = self.entropy_bottleneck
        dev = eb.quantiles.device
        Pz, Zc = zh * zw, self.hyper_channels
        h, w = 4 * zh, 4 * zw
        zdec = decoders[0] if decoders is not None else ans.acquire_decoder(B)
        zdec.set_streams_device(*z_str) if on_device else zdec.set_streams(z_str)
        z_idx = torch.arange(Zc, dtype=torch.int32, device=dev).repeat_interleave(Pz).repeat(B, 1)
        z_sym = zdec.decode_step(eb.device_tables(), z_idx)
        z_hat = torch.empty((B * Pz, Zc), dtype=torch.bfloat16, device=dev)
        check(lib().icm_eb_process(2, NULL_VIEW, B, Zc, Pz, eb.packed_params().data_ptr(), 0.0, z_sym.data_ptr(), None,
                                   NULL_VIEW, view_bcp(z_hat, B, Zc, Pz), NULL_VIEW, stream_ptr()), "icm_eb_process")
        mean_sup, scale_sup = self._hyper_synthesis(z_hat, B, zh, zw)
        ydec = decoders[1] if decoders is not None else ans.acquire_decoder(B)
        ydec.set_streams_device(*y_str) if on_device else ydec.set_streams(y_str)
        y_hat, _, _ = self._slice_loop("decompress", B, h, w, mean_sup, scale_sup, decoder=ydec)
        return y_hat, (zdec, ydec)

    @torch.no_grad()
    def decompress(self, strings, shape):
        """stf.py:734-785 for any batch size: strings = [[y strings], [z strings]], one of each per image
        (or the device-resident form produced by compress(..., device_strings=True))."""
        assert isinstance(strings, list) and len(strings) == 2
        eb = self.entropy_bottleneck
        dev = eb.quantiles.device
        if dev.type != "cuda":
            raise NativeError(f"{type(self).__name__} runs on CUDA only (no CPU fallback)")
        y_strings, z_strings = strings
        zh, zw = int(shape[0]), int(shape[1])
        on_device = len(z_strings) > 0 and isinstance(z_strings[0], tuple)
        if on_device:
            jobs = [(y, z, z[1].numel() - 1) for y, z in zip(y_strings, z_strings)]
        else:
            if len(y_strings) != len(z_strings):
                raise ValueError("need one y-string and one z-string per image")
            jobs = [(y_strings[lo:hi], z_strings[lo:hi], hi - lo) for lo, hi in self._part_ranges(len(z_strings))]
        outs = self._run_parts([(lambda y=y, z=z, n=n: self._decompress_part(y, z, n, zh, zw, on_device)) for y, z, n in jobs],
                               coder_streams=jobs[0][2])
        xs = [self._hand_over(o[0]) for o in outs]
        try:
            if not on_device:
                for _, decs in outs:
                    for d in decs:
                        d.check_status()
            else:
                # device-resident streams: the encoder's per-stream status (capacity / bad index; it became the decoder's
                # status when the streams were handed over) and the decoders' own are folded into ONE word on the device
                # and read once -- a failed encode must not decode zeros into a plausible-looking x_hat
                flag = torch.zeros(1, dtype=torch.int32, device=dev)
                for _, decs in outs:
                    for d in decs:
                        d.fold_status(flag)
                ans.raise_for_status(int(flag.item()), f"{type(self).__name__}.decompress")
        finally:
            for _, decs in outs:
                for d in decs:
                    ans.release_decoder(d)
        return {"x_hat": xs[0] if len(xs) == 1 else torch.cat(xs, 0)}
