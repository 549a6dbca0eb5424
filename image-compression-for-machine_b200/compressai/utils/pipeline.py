"""Stream-pipelined codec loop for throughput serving.

The two rANS coders are latency-bound (one warp per image stream, ~37-44 ms to encode and ~48 ms to decode a
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

    def __init__(self, model, n_streams=12, part=32, conv_sm_limit=None, decoder_streams_per_cta=8, lag=8, chains=2, decode_priority=False,
                 cuda_graphs=False):
        self.model = model
        # cuda_graphs: a job's ~340 launches are captured once per (stream slot, job shape) as four CUDA graphs -- compress
        # transforms, encoders, decode loop, synthesis (the event-chain hand-overs and the host copies stay between them) -- and
        # replayed: ~8 ms of host time per job become ~0.3 ms, which is what small jobs (8 images per GPU) and the host-facing
        # path are bound by.  The first job of a slot runs eagerly (packs weights, sizes the allocator); results are identical.
        self.cuda_graphs = bool(cuda_graphs)
        self._job_graphs = {}
        self._eager_runs = {}
        self._flag = None
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
        # decode_priority: the latency-bound decode loop of a job (7 x {grouped conv stacks, a decoder step, a grouped conv stack}, small
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
            self._job_graphs, self._eager_runs = {}, {}
            self._flag = torch.zeros(1, dtype=torch.int32, device=device)  # persistent: captured graphs fold their statuses into it

    def _decoder_pair(self, slot, n):
        from compressai import ans

        key = (slot, n)
        if key not in self._decoders:
            self._decoders[key] = (ans.StreamDecoder(n), ans.StreamDecoder(n))
        return self._decoders[key]

    def release_graphs(self):
        """Drop every captured job graph (and with it the graphs' private memory pools: ~2-4 GB per stream slot at 32 images of
        768x512); the next jobs run eagerly once and are captured again."""
        if self._job_graphs:
            torch.cuda.synchronize()  # no replay may still be running on the pools that are about to be freed
        self._job_graphs, self._eager_runs = {}, {}

    def _may_capture(self):
        """Graph pools are private memory: capture only while at least a third of the device memory is free."""
        free, total = torch.cuda.mem_get_info(self._flag.device)
        return free * 3 >= total

    def _capture_job(self, slot, n, x_shape):
        """Capture the four graphs of one job on its slot's streams; static tensors link them (shared memory pool).  The decode
        graph reads the streams where the encode graph left them: with host I/O they travel to pinned host memory and back
        into the same device buffers, so one set of graphs serves both modes."""
        import types

        m = self.model
        dev = self._flag.device
        ns = self._streams[slot]
        hs = self._hi[slot] if self._hi is not None else ns
        jg = types.SimpleNamespace()
        jg.x = torch.zeros(x_shape, dtype=torch.float32, device=dev)
        pool = torch.cuda.graph_pool_handle()
        jg.g_t, jg.g_e, jg.g_d, jg.g_s = (torch.cuda.CUDAGraph() for _ in range(4))
        n0 = lib().icm_launch_count()
        with torch.cuda.graph(jg.g_t, pool=pool, stream=ns):
            jg.sym, jg.idx, jg.z_sym, jg.z_idx, jg.zh, jg.zw = m._compress_transforms(jg.x)
        with torch.cuda.graph(jg.g_e, pool=pool, stream=ns):
            jg.y_str, jg.z_str = m._compress_encode(jg.sym, jg.idx, jg.z_sym, jg.z_idx, False)
        jg.y_in, jg.z_in = jg.y_str, jg.z_str
        decs = self._decoder_pair(slot, n)
        with torch.cuda.graph(jg.g_d, pool=pool, stream=hs):
            jg.y_hat, _ = m._decode_part(jg.y_in, jg.z_in, n, jg.zh, jg.zw, True, decoders=decs)
            for d in decs:
                d.fold_status(self._flag)
        with torch.cuda.graph(jg.g_s, pool=pool, stream=ns):
            jg.x_hat = m._synthesis(jg.y_hat, n, 4 * jg.zh, 4 * jg.zw, clamp=True)
        # kernels of the library recorded in the four graphs: reported on every replay so that icm_launch_count keeps counting
        # kernels that ran (the capture pass itself is counted once although it only records)
        jg.kernels = int(lib().icm_launch_count() - n0)
        return jg

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
        flag = self._flag  # min over every encoder size / decoder status of the call
        flag.zero_()
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

        def front_graph(t, jg):  # the same job replayed from its captured graphs
            bi, lo, hi = jobs[t]
            slot = t % self.n_streams
            ns = self._streams[slot]
            hs = self._hi[slot] if self._hi is not None else ns
            with torch.cuda.stream(ns):
                jg.x.copy_(batches[bi][lo:hi], non_blocking=True)
                hook = self._phase(t % self.chains) if self.chains else None
                if hook:
                    hook("begin")
                jg.g_t.replay()
                if hook:
                    hook("end")
                jg.g_e.replay()
                if host_io:
                    hy, hz = self._pin(jg.y_str[0].shape, torch.uint8), self._pin(jg.z_str[0].shape, torch.uint8)
                    hsy, hsz = self._pin(jg.y_str[1].shape, torch.int32), self._pin(jg.z_str[1].shape, torch.int32)
                    for h_, d_ in ((hy, jg.y_str[0]), (hz, jg.z_str[0]), (hsy, jg.y_str[1]), (hsz, jg.z_str[1])):
                        h_.copy_(d_, non_blocking=True)
                    for d_, h_ in ((jg.y_in[0], hy), (jg.z_in[0], hz), (jg.y_in[1], hsy), (jg.z_in[1], hsz)):
                        d_.copy_(h_, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(ns)
                    pending.append((ev, bi, hy, hsy, hz, hsz))
                if hs is not ns:
                    hs.wait_stream(ns)
                    with torch.cuda.stream(hs):
                        jg.g_d.replay()
                    ns.wait_stream(hs)
                else:
                    jg.g_d.replay()
            decoded[t] = (jg, jg.zh, jg.zw)
            lib().icm_note_graph_launches(jg.kernels)

        def front(t):  # C phase + encoders + decode loop of job t
            bi, lo, hi = jobs[t]
            slot = t % self.n_streams
            x = batches[bi]
            if self.cuda_graphs and not worst_case:
                key = (slot, hi - lo, tuple(x.shape[1:]))
                jg = self._job_graphs.get(key)
                if jg is None and self._eager_runs.get(key, 0) >= 1 and self._may_capture():
                    try:
                        jg = self._job_graphs[key] = self._capture_job(slot, hi - lo, (hi - lo,) + tuple(x.shape[1:]))
                    except RuntimeError as e:  # e.g. out of memory inside the capture: keep serving, eagerly
                        warnings.warn(f"RoundTripPipeline: CUDA-graph capture failed ({str(e)[:120]}); continuing without graphs")
                        self.cuda_graphs = False
                        self._job_graphs = {}
                        torch.cuda.synchronize()
                        jg = None
                if jg is not None:
                    return front_graph(t, jg)
                self._eager_runs[key] = self._eager_runs.get(key, 0) + 1
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
                if isinstance(y_hat, torch.Tensor):
                    x_hat = m._synthesis(y_hat, hi - lo, 4 * zh, 4 * zw, clamp=True)
                else:  # a captured job: its synthesis graph writes the job's static x_hat
                    y_hat.g_s.replay()
                    x_hat = y_hat.x_hat if (out_host is not None or not keep_outputs) else y_hat.x_hat.clone()
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
