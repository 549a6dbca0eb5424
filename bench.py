#!/usr/bin/env python
"""bench.py -- STF 768x512 compress+decompress throughput on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = compress + decompress of one batch of B synthetic 3x768x512 images per GPU (BASELINE.json
configs[2] on each GPU; the batch shards across GPUs with no collective).  --scaling weak (default): B images per
GPU whatever N; --scaling strong: B images in total, B / N per GPU (configs[2] as written: 64 images sharded over
1/2/4/8 GPUs).  Prints ONE JSON line:

  value         images/s, whole job, inputs and bit-streams resident in HBM (CUDA events, max over ranks)
  e2e           the same metric through the public API with HOST buffers: pinned-host image -> H2D ->
                model.compress -> Python `bytes` strings -> model.decompress(strings) -> x_hat -> D2H
  roofline      dominant kernel family of a step (per-call CUDA events in a separate instrumented step)
  cpu_baseline  the pinned CPU oracle (oracle/stf_ref.py + oracle/rans_oracle.c: the restatement of the
                reference's path) on this box's host cores, on a bounded sample of the same workload
  stress        the same pipeline with the survey's rate-raising weights (many CDF tables, ~26 % escapes): the
                default-initialised weights are the coder's easiest case (every y index is 0)
  latency_b1    BASELINE configs[1]: one 3x768x512 image, compress + decompress through the plain API, one at a time
  --impl reference   times only that CPU path (rank 0), same metric / config, and prints its own line.
"""
import os as _os

# one hardware work queue per pipeline stream: with the default of 8, streams alias and a 5 ms rANS step of one
# job falsely serialises the convolutions of another (measured: 16 streams ran at half the speed of 8)
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import argparse
import gc
import json
import math
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200"))
sys.path.insert(1, REPO)

import torch  # noqa: E402

H_IMG, W_IMG = 768, 512
SYMBOLS_PER_IMAGE = 589824 + 18432  # y + z
FLOPS_PER_IMAGE = 609.9e9           # conv + linear + bmm, compress + decompress
METRIC = "STF 768x512 compress+decompress images/s"


WEIGHTS = {"default": "PyTorch default init under torch.manual_seed(0) (the reference constructor's effective init, SURVEY.md 8d)",
           "stress": "default init + the survey's rate-raising tweak (SURVEY.md 8d: many CDF tables, ~26% escape symbols)"}


def make_model(device, weights="default", arch="stf"):
    """Reference-architecture STF (or WACNN), random weights (no checkpoint ships with the reference)."""
    from compressai.zoo import models

    torch.manual_seed(0)
    m = models[arch]()
    if weights == "stress" and arch == "stf":
        with torch.no_grad():
            m.layers[2].downsample.reduction.weight.mul_(8.0)
            ramp = torch.exp(torch.linspace(math.log(0.05), math.log(30.0), 32))
            for stack in m.cc_scale_transforms:
                stack[8].bias.copy_(ramp)
    m.update(force=True)
    return m.to(device).eval()


def make_images(batch, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand((batch, 3, H_IMG, W_IMG), generator=g)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md clocks line), read in-process through
    NVML every 100 ms.  (A polling `nvidia-smi -lms` child process was measurably intrusive: the pipelined step
    time became bimodal, 120 vs 175 ms, with it running.)  Falls back to one nvidia-smi query per sample."""

    NAMES = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period, self.rows, self._stop_ev = index, period, [], threading.Event()
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].strip().isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:  # noqa: BLE001
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            return mhz, self.max_mhz, [name for name, bit in self.NAMES if mask & bit]
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().splitlines()[0]
        c = [v.strip() for v in out.split(",")]
        return float(c[0]), float(c[1]), [name for (name, _), v in zip(self.NAMES, c[2:6]) if v.lower().startswith("active")]

    def run(self):
        while not self._stop_ev.is_set():
            try:
                self.rows.append(self._sample())
            except Exception:  # noqa: BLE001
                pass
            self._stop_ev.wait(self.period if self.nvml is not None else 0.5)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=3)
        sm = sorted(r[0] for r in self.rows)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[1] for r in self.rows), default=None),
                "reasons": sorted({n for r in self.rows for n in r[2]}), "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def pipe_capacity_bytes(B):
    """Bytes of bit-stream buffers that cross PCIe per step in the e2e path (fixed-capacity y and z buffers)."""
    return B * ((2 * 589824 + 64) + (2 * 18432 + 64) + 8)


def cpu_reference_step(sd, x, tabs):
    """One compress + decompress of the images in x on the host: the CPU restatement of the reference path."""
    from oracle import stf_ref

    c = stf_ref.compress(sd, x, gc_tab=tabs[0], eb_tab=tabs[1])
    d = stf_ref.decompress(sd, c["strings"], c["shape"], gc_tab=tabs[0], eb_tab=tabs[1])
    return c, d


def reference_state_dict(weights="default"):
    """STF parameters for the CPU arm without importing the repo's package (no native library is loaded by that arm):
    the oracle's own template + seeded values scaled like PyTorch's default initialisers (oracle/weights.py)."""
    from oracle import stf_ref
    from oracle import weights as oracle_weights

    return oracle_weights.seeded_state_dict(stf_ref.template_state_dict(), seed=0, stress=(weights == "stress"))


def cpu_baseline(model_sd, n_images, steps=1, warmup=0):
    from oracle import entropy, stf_ref

    sd = {k: v.detach().float().cpu() for k, v in model_sd.items()}
    tabs = (entropy.gc_tables(), entropy.eb_tables(stf_ref.eb_params(sd)))
    x = make_images(n_images, seed=0)
    for _ in range(warmup):
        cpu_reference_step(sd, x, tabs)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sd, x, tabs)
    dt = (time.perf_counter() - t0) / steps
    return n_images / dt, dt


def train_bench(dev, rank, world, dist, steps=6, warmup=3, total_batch=16, size=256):
    """BASELINE.json configs[4]: one rate-distortion training step of STF on `total_batch` 3 x size x size crops, sharded over
    the ranks (strong scaling: 16 / N per GPU), Adam 1e-5 / 1e-4, clip 1.0, gradient all-reduce over NCCL in buckets that
    overlap backward (compressai/training.py).  Transforms forward/backward: PyTorch autograd (library GEMMs/convolutions);
    Gaussian stage, clipping and both Adams: csrc/train.cu.  Timed with CUDA events, max over ranks."""
    from compressai.training import Trainer
    from compressai.zoo import models
    from compressai.utils.sharding import max_over_ranks

    per = max(1, total_batch // world)
    out = {"workload": f"stf training step, {total_batch}x3x{size}x{size} crops in total = {per} per GPU (BASELINE.json configs[4]), "
                       f"RateDistortionLoss lambda 800, Adam 1e-5 + aux Adam 1e-4, clip_grad_norm 1.0", "n_gpus": world, "steps": steps, "warmup": warmup}
    for label, autocast, graph in (("fp32", None, False), ("fp32_cuda_graph", None, True), ("bf16_autocast_cuda_graph", torch.bfloat16, True)):
        torch.manual_seed(0)
        net = models["stf"]().to(dev).train()
        tr = Trainer(net, lmbda=800.0, autocast=autocast, cuda_graph=graph)
        g = torch.Generator().manual_seed(77 + rank)
        x = torch.rand((per, 3, size, size), generator=g).pin_memory()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        loss = None
        for it in range(warmup + steps):
            if it == warmup:
                if dist is not None:
                    dist.barrier()
                torch.cuda.synchronize()
                e0.record()
            crit = tr.step(x.to(dev, non_blocking=True))   # host batch -> device inside the step
            loss = crit["loss"]
        e1.record()
        loss_v = float(loss.item())                          # the step's result read back
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1), device=dev) / steps
        # the flat Adam pass alone (28 B per parameter: p, g, m, v read; p, m, v written), against the HBM roofline
        opt = tr.optimizer
        ea.record()
        for _ in range(5):
            opt.step(1.0, 1.0 / world)
        eb.record()
        torch.cuda.synchronize()
        adam_ms = ea.elapsed_time(eb) / 5
        out[label] = {"ms_per_step": round(ms, 3), "images_per_s": round(total_batch / (ms / 1e3), 2), "loss_last": round(loss_v, 4),
                      "adam_clip_ms": round(adam_ms, 4), "adam_clip_gb_s": round((28 + 4) * opt.n / (adam_ms / 1e3) / 1e9, 1)}
        if label == "fp32":
            out["parameters"] = int(opt.n + tr.aux_optimizer.n)
            out["allreduce_bytes_per_step"] = int(4 * opt.n) if world > 1 else 0
            out["gradient_buckets"] = len(tr.buckets.buckets)
            out["precision"] = ("fp32 parameters / gradients / Adam state; fp32: PyTorch's default math modes like the reference's train.py "
                                "(TF32 cuDNN convolutions, fp32 matmuls); bf16_autocast: transforms' GEMMs and convolutions in bf16, residual stream, "
                                "entropy stage and loss in fp32")
        tr.buckets.remove()
        del tr, net
        gc.collect()
        torch.cuda.empty_cache()
    out["timing"] = "CUDA events around K steps incl. the H2D copy of each batch shard, max over ranks; loss read back after the last step"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--weights", default="default", choices=sorted(WEIGHTS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=-1, help="CUDA streams of the codec pipeline (-1 = automatic: ~384 images in flight, 16..32)")
    ap.add_argument("--part", type=int, default=32, help="images per pipeline job")
    ap.add_argument("--dec-per-cta", type=int, default=8, help="rANS decoder streams per CTA in the pipeline (1, 2, 4, 8, 16)")
    ap.add_argument("--lag", type=int, default=-1, help="pipeline: synthesis of job t is ordered after the compress transforms of job t+lag (-1 = automatic)")
    ap.add_argument("--chains", type=int, default=-1, help="pipeline: jobs allowed in a throughput-bound phase at once (0 = unordered; -1 = automatic: 2, or 4 for jobs of <= 8 images whose kernels do not fill the GPU)")
    ap.add_argument("--conv-sms", type=int, default=0, help="cap on SMs used by the conv kernel (0 = all: its tile scheduler is dynamic)")
    ap.add_argument("--decode-priority", type=int, default=-1, help="pipeline: run each job's decode loop on a high-priority stream (-1 = automatic)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --batch images per GPU; strong: --batch images in total")
    ap.add_argument("--graphs", type=int, default=1, help="pipeline: replay each job from CUDA graphs captured per (stream slot, job shape)")
    ap.add_argument("--no-stress", action="store_true", help="skip the stress-weights sub-record")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling sub-record (64 images in total)")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-image latency sub-record")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step sub-record (BASELINE.json configs[4])")
    ap.add_argument("--train-only", action="store_true", help="print only the training-step record (development aid)")
    ap.add_argument("--cpu-images", type=int, default=2, help="images per step of the CPU arm / cpu_baseline sample")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    b_gpu = args.batch // world if args.scaling == "strong" and args.batch % world == 0 else args.batch
    # Pipeline depth: the two coders are latency-bound (~35 ms + ~50 ms per job whatever its size), so throughput needs
    # ~400 images in flight; small per-GPU batches (strong scaling: 8 images per GPU at N = 8) therefore need more,
    # smaller jobs in flight.  The streams (and their high-priority twins) must fit the 32 hardware queues.
    def pipe_params(b):
        part = max(1, min(args.part, b))
        streams = args.streams if args.streams >= 0 else max(16, min(32, -(-384 // part)))
        prio = args.decode_priority if args.decode_priority >= 0 else (1 if streams <= 16 else 0)
        lag = args.lag if args.lag >= 0 else (streams - 4 if streams <= 16 else streams - 3)  # sweep of round 2: 16 / 12 gave 878 images/s, 12 / 8 gave 850
        return streams, part, lag, prio

    auto_chains = args.chains < 0
    chains_for = lambda part: (4 if part <= 8 else 2) if auto_chains else args.chains
    auto = (args.streams, args.lag, args.decode_priority)
    args.streams, _, args.lag, args.decode_priority = pipe_params(b_gpu)
    args.chains = chains_for(max(1, min(args.part, b_gpu)))
    if args.scaling == "strong":
        if args.batch % world:
            raise SystemExit(f"--scaling strong: --batch {args.batch} is not a multiple of the {world} GPUs")
        total_images, args.batch = args.batch, args.batch // world
        workload = (f"stf 3x{H_IMG}x{W_IMG}, {total_images} images per step in total = {args.batch} per GPU "
                    f"(BASELINE.json configs[2] as written; batch sharded, no collective)")
    else:
        workload = f"stf 3x{H_IMG}x{W_IMG}, {args.batch} images per GPU per step (BASELINE.json configs[2] on every GPU; batch sharded, no collective)"
    config = {"workload": workload, "images_per_gpu": args.batch, "height": H_IMG, "width": W_IMG,
              "weights": WEIGHTS[args.weights], "l2": "inputs larger than L2 (302 MB per batch at B=64)",
              "pipeline": f"{args.streams} CUDA streams x jobs of {args.part} images (compress -> decompress per job), K steps in flight; "
                          f"transform phases ordered in {args.chains} event chains, synthesis lagging {args.lag} jobs"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        # torchrun exports OMP_NUM_THREADS=1; the reference arm gets every host core this process may run on
        torch.set_num_threads(len(os.sched_getaffinity(0)))
        n_cpu = max(1, args.cpu_images)
        warm = max(1, min(args.warmup, 1))
        ips, dt = cpu_baseline(reference_state_dict(args.weights), n_cpu, steps=max(1, args.steps), warmup=warm)
        cores = torch.get_num_threads()
        sample = (f"{n_cpu} images (3x768x512) compress+decompress per timed step (a bounded sample of the {args.batch}-image workload; "
                  f"images/s = {n_cpu} / step time), {warm} untimed warm-up step, oracle/stf_ref.py fp32 + oracle/rans_oracle.c; "
                  f"parameters from oracle/weights.py (the repo's package and native library are not loaded by this arm)")
        config["cpu_sample_images_per_step"] = n_cpu
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this benchmark has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    from compressai import _native
    from compressai.utils.sharding import max_over_ranks

    if args.train_only:
        rec = train_bench(dev, rank, world, dist, steps=args.steps, warmup=max(3, args.warmup))
        if rank == 0:
            print(json.dumps({"train_step": rec}))
        if dist is not None:
            dist.destroy_process_group()
        return 0
    model = make_model(dev, args.weights)
    B = args.batch
    x_host = make_images(B, seed=rank).pin_memory()
    x_dev = x_host.to(dev)
    out_host = torch.empty((B, 3, H_IMG, W_IMG), dtype=torch.float32).pin_memory()

    from compressai.utils.pipeline import RoundTripPipeline

    def make_pipe(mdl):
        return RoundTripPipeline(mdl, n_streams=args.streams, part=min(args.part, B), conv_sm_limit=(args.conv_sms if args.conv_sms >= 0 else None),
                                 decoder_streams_per_cta=args.dec_per_cta, lag=args.lag, chains=args.chains, decode_priority=bool(args.decode_priority),
                                 cuda_graphs=bool(args.graphs))

    pipe = make_pipe(model)
    out_bufs = [out_host, torch.empty_like(out_host).pin_memory()] if not args.no_e2e else []

    def run_device(steps, p=None):
        """K steps = K batches through the stream pipeline, images and bit-streams resident in HBM."""
        return (p or pipe).roundtrip([x_dev] * steps, keep_outputs=False)  # a server hands x_hat on and releases it

    def run_e2e(steps):
        """K steps through the host-facing path: pinned images -> H2D -> compress -> streams to the host and back
        -> decompress -> x_hat -> pinned host buffers; returns the Python byte strings of every image."""
        return pipe.roundtrip([x_host] * steps, host_io=True, out_host=[out_bufs[i % 2] for i in range(steps)])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, wall=False):
        gc.collect()
        gc.disable()  # no collector pauses on the enqueueing thread inside the timed region
        try:
            return _timed(fn, steps, wall)
        finally:
            gc.enable()

    def _timed(fn, steps, wall=False):
        barrier()
        l0 = _native.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        r = fn(steps)
        e1.record()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 if wall else e0.elapsed_time(e1)
        launches = _native.launch_count() - l0
        ms = max_over_ranks(ms, device=dev)
        barrier()
        return ms, launches, r

    # at least 3 untimed steps, and enough of them for every pipeline stream to have run a job (each stream owns a
    # slice of the caching allocator and a decoder pair: their first use must not fall into the timed region)
    jobs_per_step = -(-B // min(args.part, B))
    warm_steps = max(args.warmup, 3, -(-args.streams // jobs_per_step))
    if args.graphs:  # every stream slot runs one job eagerly, captures its graphs on the second, replays from the third
        warm_steps = max(warm_steps, 2 * -(-args.streams // jobs_per_step) + 1)
    run_device(warm_steps)
    torch.cuda.synchronize()
    # the device-resident pipeline (no per-job host read) must give exactly what the host-string API gives
    check_n = min(B, 2)
    xh_dev, _ = pipe.roundtrip([x_dev[:check_n]], keep_outputs=True)
    c_host = model.compress(x_dev[:check_n])
    xh_host = model.decompress(c_host["strings"], c_host["shape"])["x_hat"]
    torch.cuda.synchronize()
    if not torch.equal(xh_dev[0], xh_host):
        raise SystemExit("bench: the device-resident pipeline and the host-string API disagree on x_hat")
    del xh_dev, xh_host, c_host
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ms, launches, _ = timed(run_device, args.steps)
    clocks = sampler.stop() if sampler else None
    n_img = B * world * args.steps
    value = n_img / (ms / 1e3)
    e2e = None
    str_bytes = 0
    if not args.no_e2e:
        run_e2e(warm_steps)  # fills the pool of pinned staging buffers: none is allocated inside the timed region
        ms_e2e, _, (_, strings) = timed(run_e2e, args.steps, wall=True)
        str_bytes = sum(len(s) for grp in strings[0] for s in grp)
        e2e = {"value": round(n_img / (ms_e2e / 1e3), 3), "unit": "images/s",
               "h2d_bytes_per_step": x_host.numel() * 4 + pipe_capacity_bytes(B),
               "d2h_bytes_per_step": out_host.numel() * 4 + pipe_capacity_bytes(B), "ms_per_step": round(ms_e2e / args.steps, 3),
               "timing": "wall clock around K pipelined steps incl. creation of the Python byte strings"}

    pipe.release_graphs()  # their private memory pools go back to the allocator before further pipelines are built
    torch.cuda.empty_cache()
    # the same pipeline with the survey's stress weights: many CDF tables in use, ~26 % escape symbols (the coder's hard case)
    stress = None
    if not args.no_stress and args.weights != "stress":
        m2 = make_model(dev, "stress")
        p2 = make_pipe(m2)
        run_device(warm_steps, p2)
        torch.cuda.synchronize()
        ms2, _, _ = timed(lambda k: run_device(k, p2), args.steps)
        c2 = m2.compress(x_dev[:min(B, 4)])
        nb2 = sum(len(s_) for s_ in c2["strings"][0]) / min(B, 4)
        stress = {"value": round(n_img / (ms2 / 1e3), 3), "unit": "images/s", "ms_per_step": round(ms2 / args.steps, 3),
                  "weights": WEIGHTS["stress"], "y_bytes_per_image": round(nb2, 1), "timing": "as `value`: device-resident pipeline, CUDA events"}
        del m2, p2, c2
        torch.cuda.empty_cache()
    # BASELINE configs[2] as written: the SAME 64 images sharded over the N GPUs (64 / N each), next to the weak-scaling
    # headline (64 per GPU).  At N = 8 a GPU holds 8 images per step and the coders' latency needs many small jobs in flight.
    strong = None
    if world > 1 and args.scaling == "weak" and B % world == 0 and not args.no_strong:
        bs = B // world
        resolved = (args.streams, args.lag, args.decode_priority)
        args.streams, args.lag, args.decode_priority = auto
        st3, part3, lag3, prio3 = pipe_params(bs)
        args.streams, args.lag, args.decode_priority = resolved
        p3 = RoundTripPipeline(model, n_streams=st3, part=part3, conv_sm_limit=(args.conv_sms if args.conv_sms >= 0 else None),
                               decoder_streams_per_cta=args.dec_per_cta, lag=lag3, chains=chains_for(part3), decode_priority=bool(prio3),
                               cuda_graphs=bool(args.graphs))
        xs = x_dev[:bs].contiguous()
        steps3 = args.steps * world  # the same number of images per GPU as the weak run
        p3.roundtrip([xs] * max(3, (2 if args.graphs else 1) * -(-st3 // max(1, -(-bs // part3))) + 1), keep_outputs=False)
        torch.cuda.synchronize()
        ms3, _, _ = timed(lambda k: p3.roundtrip([xs] * k, keep_outputs=False), steps3)
        strong = {"value": round(B * steps3 / (ms3 / 1e3), 3), "unit": "images/s", "images_total_per_step": B, "images_per_gpu": bs,
                  "steps": steps3, "ms_per_step": round(ms3 / steps3, 3), "scaling": "strong",
                  "pipeline": f"{st3} streams x jobs of {part3} images, {chains_for(part3)} event chains, synthesis lagging {lag3} jobs", "timing": "as `value`"}
        del p3, xs
        torch.cuda.empty_cache()
    # BASELINE configs[1]: one image at a time through the plain API (host strings), median of 7
    latency = None
    if rank == 0 and not args.no_latency:
        x1 = x_dev[:1]
        ts = []
        for it in range(9):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c1 = model.compress(x1)
            d1 = model.decompress(c1["strings"], c1["shape"])
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        ts = sorted(ts[2:])
        latency = {"ms": round(ts[len(ts) // 2], 2), "images_per_s": round(1e3 / ts[len(ts) // 2], 2),
                   "workload": "stf 1x3x768x512 compress + decompress (BASELINE.json configs[1]), plain API, host byte strings, one image at a time",
                   "timing": "wall clock incl. host enqueue and synchronisation, median of 7 after 2 warm-up runs"}
        del c1, d1
    train = None
    if not args.no_train and 16 % world == 0:
        train = train_bench(dev, rank, world, dist)
    # instrumented step: per-entry-point CUDA events -> dominant kernel family and its roofline
    roofline, families = None, None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        model.micro_batches = 1
        torch.cuda.synchronize()
        # cudaProfilerStart/Stop bracket exactly this step: `ncu --profile-from-start off ... python bench.py` then
        # lists the launches the roofline below is computed from (profiles/README.md), not the pipelined warm-up.
        torch.cuda.cudart().cudaProfilerStart()
        with _native.Profile() as prof:  # plain API, one stream: clean per-kernel times
            c = model.compress(x_dev, device_strings=True)
            model.decompress(c["strings"], c["shape"])
            summ = prof.summary()
        torch.cuda.cudart().cudaProfilerStop()
        total_ms = sum(v[1] for v in summ.values())
        families = {k: {"calls": v[0], "ms": round(v[1], 3), "share": round(v[1] / total_ms, 4)} for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1])}
        # conv_igemm_kernel is launched by icm_conv2d and icm_conv2d_grouped (several same-shaped convolutions per launch)
        conv_parts = [summ.get(k, (0, 0.0, 0.0)) for k in ("icm_conv2d", "icm_conv2d_grouped")]
        calls, conv_ms, conv_flops = (sum(p[i] for p in conv_parts) for i in range(3))
        dec_ms = summ.get("icm_rans_decoder_step", (0, 0.0, 0.0))[1]
        enc_ms = summ.get("icm_rans_encode_batch", (0, 0.0, 0.0))[1]
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        traffic, traffic_src = None, None
        try:  # DRAM bytes per conv launch from the committed `ncu --set full` capture of this workload (B = 64 only)
            tj = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))["conv_igemm_kernel"]
            if tj.get("images_per_step") == B and tj.get("workload") == "stf":
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:  # noqa: BLE001
            pass
        ach = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms else 0.0
        roofline = {"kernel": "conv_igemm_kernel (icm_conv2d + icm_conv2d_grouped: all convolutions and linears)", "bound": "tensor", "achieved": round(ach, 2),
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": round(ach / peak_tf, 4), "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400 (of fallback)",
                    "launches_per_step": calls, "avg_launch_us": round(conv_ms * 1e3 / max(calls, 1), 2),
                    "algorithmic_flops_per_step": conv_flops, "share_of_step": round(conv_ms / total_ms, 4),
                    "rans": {"decode_msym_s": round(B * 589824 / (dec_ms / 1e3) / 1e6, 1) if dec_ms else None,
                             "encode_msym_s": round(B * SYMBOLS_PER_IMAGE / (enc_ms / 1e3) / 1e6, 1) if enc_ms else None,
                             "streams_in_flight": B}}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
        n_cpu = max(1, args.cpu_images)
        ips, dt = cpu_baseline(model.state_dict(), n_cpu, steps=2, warmup=1)
        cpu = {"value": round(ips, 4), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{n_cpu} images (3x768x512) compress+decompress per step, 1 warm-up + 2 timed steps ({dt:.1f} s each), "
                         f"same weights as the GPU arm, oracle/stf_ref.py fp32 + oracle/rans_oracle.c"}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": round(value, 3), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warm_steps,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16 operands / fp32 accumulate (transforms); fp32 (entropy models); u64 (rANS)", "data": "synthetic",
            "config": config, "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches) * world, "roofline": roofline, "cpu_baseline": cpu, "stress": stress, "strong_scaling": strong, "latency_b1": latency, "train_step": train,
            "families": families,
            "msym_per_s": round(value * SYMBOLS_PER_IMAGE / 1e6, 2),
            "bytes_per_image": round(str_bytes / B, 1),
        }))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
