"""Stream-pipelined codec loop for throughput serving.

The two rANS coders are latency-bound (one warp per image stream, ~35 ms to encode and ~60 ms to decode a
768x512 image whatever the batch size) while the transforms are throughput-bound.  A batch is therefore cut
into jobs of `part` images; each job runs compress -> decompress on one stream of a small pool, jobs of the
same and of following batches are spread round-robin over the pool, and the GPU overlaps the coder of one
job with the convolutions of the others.  The per-image results are bit-identical to model.compress /
model.decompress on the whole batch (every kernel is batch-invariant).
"""
import os
import warnings

import torch

from compressai._native import check, lib

# The pipeline needs one hardware work queue per stream; CUDA's default is 8 and streams beyond that alias, which
# serialises a 5 ms rANS step of one job with the convolutions of another.  The variable is read when the CUDA
# context is created, so it only helps if this module is imported before the first CUDA call.
if "CUDA_DEVICE_MAX_CONNECTIONS" not in os.environ:
    if torch.cuda.is_initialized():
        warnings.warn("compressai.utils.pipeline imported after CUDA initialisation: set CUDA_DEVICE_MAX_CONNECTIONS=32 "
                      "in the environment, or use at most 8 pipeline streams", stacklevel=2)
    else:
        os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = "32"


class RoundTripPipeline:
    """lag / chains: the throughput-bound phases (compress transforms "C", synthesis "S") of all jobs are ordered
    into `chains` token chains C_t, S_{t-lag}, C_{t+1}, S_{t-lag+1}, ... with CUDA events, so that at most `chains`
    jobs compete for the tensor cores at any time while the rANS phases (encode after C, the decode loop before S) of
    up to `lag` other jobs run beside them.  Without the chain all jobs start in lockstep, reach their coders
    together and leave the GPU idle (measured run-to-run spread 400-530 images/s)."""

    def __init__(self, model, n_streams=12, part=32, conv_sm_limit=None, decoder_streams_per_cta=8, lag=8, chains=2, decode_priority=False):
        self.model = model
        self.decoder_streams_per_cta = int(decoder_streams_per_cta)
        # SMs the persistent conv / MLP kernels may occupy.  0 = all: the conv kernel's tile scheduler is dynamic, so a CTA
        # that finds its SM held by a decoder CTA (~190 KB of shared memory: they cannot share an SM) simply starts late
        # and finds no tiles left.  None = the round-1 rule for a static scheduler: 148 - (jobs in their decode loop) *
        # ceil(part / streams per CTA).
        self.conv_sm_limit = conv_sm_limit
        self.n_streams = int(n_streams)
        self.part = int(part)
        self.lag = max(0, min(int(lag), self.n_streams - 1))
        self.chains = max(0, int(chains))
        # decode_priority: the latency-bound decode loop of a job (12 x {two conv stacks, a decoder step, a conv stack}, small
        # kernels) runs on a high-priority twin of the job's stream, so that its CTAs are placed ahead of the queued CTAs of
        # other jobs' throughput-bound kernels
        self.decode_priority = bool(decode_priority)
        self._hi = None
        self._streams = None
        self._decoders = {}
        self._pinned = {}
        self._tokens = {}

    def _setup(self, device):
        if self._streams is None or self._streams[0].device != device:
            self._streams = [torch.cuda.Stream(device=device) for _ in range(self.n_streams)]
            self._hi = [torch.cuda.Stream(device=device, priority=-1) for _ in range(self.n_streams)] if self.decode_priority else None
            self._decoders = {}

    def _decoder_pair(self, slot, n):
        from compressai import ans

        key = (slot, n)
        if key not in self._decoders:
            self._decoders[key] = (ans.StreamDecoder(n), ans.StreamDecoder(n))
        return self._decoders[key]

    def _pin(self, shape, dtype):
        """A pinned staging buffer from the pool (returned by _unpin once its job's strings have been built).  The first request
        for a shape allocates one buffer per job that can be in flight (cudaHostAlloc takes ~35 ms for a 40 MB buffer: it must
        never happen in the middle of a run)."""
        key = (tuple(shape), dtype)
        free = self._pinned.get(key)
        if free is None:
            free = self._pinned[key] = [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(self.n_streams + 2)]
        return free.pop() if free else torch.empty(shape, dtype=dtype).pin_memory()

    def _unpin(self, *bufs):
        for t in bufs:
            self._pinned[(tuple(t.shape), t.dtype)].append(t)

    def _phase(self, chain):
        """begin: wait for the chain's token; end: pass it on."""
        if self.chains == 0:
            return None

        def hook(what):
            st = torch.cuda.current_stream()
            if what == "begin":
                ev = self._tokens.get(chain)
                if ev is not None:
                    st.wait_event(ev)
            else:
                ev = torch.cuda.Event()
                ev.record(st)
                self._tokens[chain] = ev
        return hook

    @torch.no_grad()
    def roundtrip(self, batches, host_io=False, out_host=None, keep_outputs=True):
        """See _roundtrip.  The encoders write into buffers sized for 16 bit/symbol; if a stream outgrows that (the
        folded device status says ICM_ERR_CAPACITY) the whole call is repeated with worst-case buffers."""
        from compressai import ans

        try:
            return self._roundtrip(batches, host_io, out_host, keep_outputs, worst_case=False)
        except ans.CapacityError:
            torch.cuda.synchronize()
            return self._roundtrip(batches, host_io, out_host, keep_outputs, worst_case=True)

    def _roundtrip(self, batches, host_io, out_host, keep_outputs, worst_case):
        """compress + decompress every batch of `batches` ([B,3,H,W] CUDA tensors, or pinned host tensors with
        host_io=True).  Returns (x_hats, strings): x_hats per batch (CUDA, or written into `out_host`), strings per
        batch as [[y bytes...], [z bytes...]] when host_io else None.  keep_outputs=False drops each job's x_hat as soon as
        it has been produced (or copied to `out_host`), so its memory is reused by the next job of that stream."""
        m = self.model
        dev = m.entropy_bottleneck.quantiles.device
        self._setup(dev)
        cur = torch.cuda.current_stream()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        limit = self.conv_sm_limit
        if limit is None:
            per = max(1, self.decoder_streams_per_cta)
            decoding = self.n_streams if self.chains == 0 else min(self.n_streams, self.lag + 1)
            limit = max(sms // 2, sms - decoding * ((self.part + per - 1) // per))
        check(lib().icm_set_conv_sm_limit(int(limit)), "icm_set_conv_sm_limit")
        check(lib().icm_set_decoder_streams_per_cta(self.decoder_streams_per_cta), "icm_set_decoder_streams_per_cta")
        flag = torch.zeros(1, dtype=torch.int32, device=dev)  # min over every encoder size / decoder status of the call
        for st in self._streams:
            st.wait_stream(cur)
            flag.record_stream(st)
        for st in self._hi or ():
            flag.record_stream(st)
        self._tokens = {}
        jobs = []
        for bi, x in enumerate(batches):
            m._check_input(x if not host_io else x[:1].to(dev))
            for lo in range(0, x.shape[0], self.part):
                jobs.append((bi, lo, min(x.shape[0], lo + self.part)))
        results = [[] for _ in batches]
        pending, decoded = [], {}
        from compressai import ans

        lag = self.lag if self.chains else 0

        def front(t):  # C phase + encoders + decode loop of job t
            bi, lo, hi = jobs[t]
            slot = t % self.n_streams
            x = batches[bi]
            with torch.cuda.stream(self._streams[slot]):
                xd = x[lo:hi].to(dev, non_blocking=True) if host_io else x[lo:hi]
                c = m._compress_part(xd, phase=self._phase(t % self.chains) if self.chains else None, worst_case=worst_case)
                zh, zw = c["shape"]
                y_str, z_str = c["y"], c["z"]
                if host_io:  # streams leave for the host and come back, like bytes handed to a decoder
                    hy, hz = self._pin(y_str[0].shape, torch.uint8), self._pin(z_str[0].shape, torch.uint8)
                    hsy, hsz = self._pin(y_str[1].shape, torch.int32), self._pin(z_str[1].shape, torch.int32)
                    hy.copy_(y_str[0], non_blocking=True); hz.copy_(z_str[0], non_blocking=True)
                    hsy.copy_(y_str[1], non_blocking=True); hsz.copy_(z_str[1], non_blocking=True)
                    y_str = (hy.to(dev, non_blocking=True), hsy.to(dev, non_blocking=True))
                    z_str = (hz.to(dev, non_blocking=True), hsz.to(dev, non_blocking=True))
                    ev = torch.cuda.Event()
                    ev.record(self._streams[slot])  # both directions of the staging buffers are done after this
                    pending.append((ev, bi, hy, hsy, hz, hsz))
                decs = self._decoder_pair(slot, hi - lo)
                if self._hi is None:
                    y_hat, _ = m._decode_part(y_str, z_str, hi - lo, zh, zw, True, decoders=decs)
                    for d in decs:  # encoder errors travel with the streams into the decoder status
                        d.fold_status(flag)
                else:
                    hs, ns = self._hi[slot], self._streams[slot]
                    hs.wait_stream(ns)
                    for tns in (*y_str, *z_str):
                        tns.record_stream(hs)
                    with torch.cuda.stream(hs):
                        y_hat, _ = m._decode_part(y_str, z_str, hi - lo, zh, zw, True, decoders=decs)
                        for d in decs:
                            d.fold_status(flag)
                    y_hat.record_stream(ns)
                    ns.wait_stream(hs)
                decoded[t] = (y_hat, zh, zw)

        def back(t):  # S phase of job t
            bi, lo, hi = jobs[t]
            y_hat, zh, zw = decoded.pop(t)
            with torch.cuda.stream(self._streams[t % self.n_streams]):
                hook = self._phase(t % self.chains) if self.chains else None
                if hook:
                    hook("begin")
                x_hat = m._synthesis(y_hat, hi - lo, 4 * zh, 4 * zw, clamp=True)
                if hook:
                    hook("end")
                if out_host is not None:
                    out_host[bi][lo:hi].copy_(x_hat, non_blocking=True)
                if keep_outputs and out_host is None:
                    results[bi].append(x_hat)

        strings = [[[], []] for _ in batches] if host_io else None

        def drain(block):
            """Build the Python byte strings of finished jobs (in job order) while the GPU works on later ones."""
            while pending and (block or pending[0][0].query()):
                ev, bi, hy, hsy, hz, hsz = pending.pop(0)
                ev.synchronize()
                for packed, sizes, dst in ((hy, hsy, strings[bi][0]), (hz, hsz, strings[bi][1])):
                    hs = sizes.tolist()
                    if min(hs[:-1]) < 0:
                        ans.raise_for_status(min(hs[:-1]), "RoundTripPipeline (encoder)")
                    raw, o = packed.numpy(), 0
                    for v in hs[:-1]:
                        dst.append(raw[o:o + v].tobytes())
                        o += v
                self._unpin(hy, hsy, hz, hsz)

        try:
            for t in range(len(jobs) + lag):
                if t < len(jobs):
                    front(t)
                if t - lag >= 0:
                    back(t - lag)
                if host_io:
                    drain(False)
            for st in self._streams:
                cur.wait_stream(st)
            if host_io:
                drain(True)
            # one host read for the whole call: a failed encode (capacity, bad index) must not pass as a decoded image
            ans.raise_for_status(int(flag.item()), "RoundTripPipeline")
        finally:
            check(lib().icm_set_conv_sm_limit(0), "icm_set_conv_sm_limit")
            check(lib().icm_set_decoder_streams_per_cta(0), "icm_set_decoder_streams_per_cta")
        x_hats = None
        if out_host is None and keep_outputs:
            x_hats = []
            for parts in results:
                for t in parts:
                    t.record_stream(cur)
                x_hats.append(parts[0] if len(parts) == 1 else torch.cat(parts, 0))
        if host_io:
            cur.synchronize()  # x_hat copies into out_host have landed
        return x_hats, strings
