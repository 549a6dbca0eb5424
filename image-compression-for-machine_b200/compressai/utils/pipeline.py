"""Stream-pipelined codec loop for throughput serving.

The two rANS coders are latency-bound (one warp per image stream, ~35 ms to encode and ~60 ms to decode a
768x512 image whatever the batch size) while the transforms are throughput-bound.  A batch is therefore cut
into jobs of `part` images; each job runs compress -> decompress on one stream of a small pool, jobs of the
same and of following batches are spread round-robin over the pool, and the GPU overlaps the coder of one
job with the convolutions of the others.  The per-image results are bit-identical to model.compress /
model.decompress on the whole batch (every kernel is batch-invariant).
"""
import torch

from compressai._native import check, lib


class RoundTripPipeline:
    def __init__(self, model, n_streams=12, part=16, conv_sm_limit=None, decoder_streams_per_cta=4):
        self.model = model
        self.decoder_streams_per_cta = int(decoder_streams_per_cta)
        # The decoder's CTAs (4 streams each, ~155 KB of shared memory) cannot share an SM with a persistent
        # conv CTA (~200 KB); the conv grid leaves them room.  None = 148 - streams * ceil(part / 4).
        self.conv_sm_limit = conv_sm_limit
        self.n_streams = int(n_streams)
        self.part = int(part)
        self._streams = None
        self._decoders = {}
        self._pinned = {}

    def _setup(self, device):
        if self._streams is None or self._streams[0].device != device:
            self._streams = [torch.cuda.Stream(device=device) for _ in range(self.n_streams)]
            self._decoders = {}

    def _decoder_pair(self, slot, n):
        from compressai import ans

        key = (slot, n)
        if key not in self._decoders:
            self._decoders[key] = (ans.StreamDecoder(n), ans.StreamDecoder(n))
        return self._decoders[key]

    def _pin(self, key, shape, dtype):
        t = self._pinned.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype).pin_memory()
            self._pinned[key] = t
        return t

    @torch.no_grad()
    def roundtrip(self, batches, host_io=False, out_host=None):
        """compress + decompress every batch of `batches` ([B,3,H,W] CUDA tensors, or pinned host tensors with
        host_io=True).  Returns (x_hats, strings): x_hats per batch (CUDA, or written into `out_host`), strings per
        batch as [[y bytes...], [z bytes...]] when host_io else None."""
        m = self.model
        dev = m.entropy_bottleneck.quantiles.device
        self._setup(dev)
        cur = torch.cuda.current_stream()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        limit = self.conv_sm_limit
        if limit is None:
            per = max(1, self.decoder_streams_per_cta)
            limit = max(sms // 2, sms - self.n_streams * ((self.part + per - 1) // per))
        check(lib().icm_set_conv_sm_limit(int(limit)), "icm_set_conv_sm_limit")
        check(lib().icm_set_decoder_streams_per_cta(self.decoder_streams_per_cta), "icm_set_decoder_streams_per_cta")
        for st in self._streams:
            st.wait_stream(cur)
        results, pending, job = [], [], 0
        try:
            for bi, x in enumerate(batches):
                B = x.shape[0]
                m._check_input(x if not host_io else x[:1].to(dev))
                parts = []
                for lo in range(0, B, self.part):
                    hi = min(B, lo + self.part)
                    slot = job % self.n_streams
                    st = self._streams[slot]
                    job += 1
                    with torch.cuda.stream(st):
                        xd = x[lo:hi].to(dev, non_blocking=True) if host_io else x[lo:hi]
                        c = m._compress_part(xd)
                        zh, zw = c["shape"]
                        y_str, z_str = c["y"], c["z"]
                        if host_io:  # streams leave for the host and come back, like bytes handed to a decoder
                            hy = self._pin(("y", bi, lo), y_str[0].shape, torch.uint8)
                            hz = self._pin(("z", bi, lo), z_str[0].shape, torch.uint8)
                            hsy = self._pin(("sy", bi, lo), y_str[1].shape, torch.int32)
                            hsz = self._pin(("sz", bi, lo), z_str[1].shape, torch.int32)
                            hy.copy_(y_str[0], non_blocking=True); hz.copy_(z_str[0], non_blocking=True)
                            hsy.copy_(y_str[1], non_blocking=True); hsz.copy_(z_str[1], non_blocking=True)
                            y_str = (hy.to(dev, non_blocking=True), hsy.to(dev, non_blocking=True))
                            z_str = (hz.to(dev, non_blocking=True), hsz.to(dev, non_blocking=True))
                            pending.append((bi, lo, hi, hy, hsy, hz, hsz))
                        x_hat, _ = m._decompress_part(y_str, z_str, hi - lo, zh, zw, True, decoders=self._decoder_pair(slot, hi - lo))
                        if out_host is not None:
                            out_host[bi][lo:hi].copy_(x_hat, non_blocking=True)
                        parts.append(x_hat)
                results.append(parts)
            for st in self._streams:
                cur.wait_stream(st)
        finally:
            check(lib().icm_set_conv_sm_limit(0), "icm_set_conv_sm_limit")
            check(lib().icm_set_decoder_streams_per_cta(0), "icm_set_decoder_streams_per_cta")
        x_hats = None
        if out_host is None:
            x_hats = []
            for parts in results:
                for t in parts:
                    t.record_stream(cur)
                x_hats.append(parts[0] if len(parts) == 1 else torch.cat(parts, 0))
        strings = None
        if host_io:
            cur.synchronize()
            strings = [[[], []] for _ in batches]
            for bi, lo, hi, hy, hsy, hz, hsz in pending:
                for packed, sizes, dst in ((hy, hsy, strings[bi][0]), (hz, hsz, strings[bi][1])):
                    hs = sizes.tolist()
                    if min(hs[:-1]) < 0:
                        raise RuntimeError("rANS encoder reported an error status (buffer capacity or bad index)")
                    raw, o = packed.numpy(), 0
                    for v in hs[:-1]:
                        dst.append(raw[o:o + v].tobytes())
                        o += v
        return x_hats, strings
