#!/usr/bin/env python
"""bench.py -- encode+decode throughput of the chunk-parallel range coder.

A "step" is one pass of the whole hot path over one batch that is already
resident in HBM: histogram -> (NCCL all-reduce of the K counts when N > 1) ->
cum_freq model -> encode (+ compaction into one stream) -> decode.
Workload at N=1: BASELINE.json configs[1] -- 1 GiB of Zipf(1.1) bytes, 256
symbols, static global frequency table, 64 KiB chunks.  For N > 1 every rank
codes its own 1 GiB shard of one global stream ("weak" scaling) under the
table built from the all-reduced counts.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference` times the reference's algorithm on the host cores.  The
reference is a Rust crate and this image has no Rust toolchain, so that arm
runs the C oracle port (oracle/rc_oracle.c), one chunk per thread.
"""
import argparse
import json
import os
import socket
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json's metric, verbatim; what `value` is exactly: uncompressed input bytes / time of one step, a step
# being histogram -> table -> encode -> decode of the batch (every byte encoded then decoded, bit-exact)
METRIC = "encode & decode GB/s (input bytes) at 1/2/4/8 B200, bit-exact to reference"
METRIC_DEFINITION = ("uncompressed input bytes per second through one round trip: histogram + table + encode + "
                     "decode of the batch; encode_gbs / decode_gbs are the two phases alone")
UNIT = "GB/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--bytes", type=int, default=1 << 30, help="input bytes per GPU")
    p.add_argument("--chunk", type=int, default=65536, help="symbols per chunk")
    p.add_argument("--alphabet", type=int, default=256)
    p.add_argument("--zipf", type=float, default=1.1)
    p.add_argument("--mode", default="static", choices=["static", "adaptive"],
                   help="static: one global table (all-reduced counts); adaptive: one table per chunk, mixed entropy")
    p.add_argument("--seed", type=lambda s: int(s, 0), default=None)
    p.add_argument("--e2e-steps", type=int, default=8)
    p.add_argument("--e2e-depth", type=int, default=2,
                   help="host-buffer calls of each kind (encode / decode) in flight in the end-to-end leg")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-parity", action="store_true")
    p.add_argument("--trace-steps", action="store_true", help="per-step phase times of this rank on stderr")
    p.add_argument("--enc-threads", type=int, default=0)
    p.add_argument("--dec-threads", type=int, default=0)
    p.add_argument("--restart", type=int, default=-1,
                   help="restart points every this many symbols of a chunk (several decoder lanes per chunk); "
                        "0 = none, -1 = chunk / 16 (static byte table), / 8 (per-chunk tables; / 16 from 128 Ki symbols), / 4 (u16 symbols)")
    a = p.parse_args()
    if a.alphabet > 256 and a.chunk == 65536:
        a.chunk = 32768  # 64 KiB chunks of u16 symbols
    if a.restart < 0:
        # lanes per chunk for the decoder: 16 for byte chunks of one static table (measured best: 19 warps per SM in
        # three even waves), 8 for per-chunk tables, 4 for the 4096-symbol alphabet (profiles/r02b_sweep.jsonl)
        # (per-chunk tables at 256 KiB chunks: 16 lanes decode in 5.8 ms against 7.8 with 8, profiles/r02c_sweep.jsonl)
        parts = (16 if a.chunk >= 131072 else 8) if a.mode == "adaptive" else (4 if a.alphabet > 256 else 16)
        a.restart = a.chunk // parts if a.chunk % (64 * parts) == 0 else (a.chunk // 4 if a.chunk % 256 == 0 else 0)
    if a.seed is None:
        a.seed = 0x5EED0002 if a.mode == "adaptive" else (0x5EED0003 if a.alphabet > 256 else 0x5EED0001)
    return a


S_CYCLE = (0.0, 0.25, 0.5, 0.8, 1.1, 1.5, 2.0, 3.0, 5.0)  # chunk j is drawn from Zipf(S_CYCLE[j % 9])


def workload_name(a):
    sb = 1 if a.alphabet <= 256 else 2
    kib = a.chunk * sb // 1024
    if a.mode == "adaptive":
        which = "BASELINE.json configs[2]" if a.alphabet == 256 and a.bytes == 1 << 30 else "custom shape"
        return (f"{a.bytes / 2**30:g} GiB/GPU mixed-entropy bytes (Zipf s cycling {S_CYCLE[0]}..{S_CYCLE[-1]} per chunk), "
                f"{a.alphabet} symbols, adaptive per-chunk histogram + table, {kib} KiB chunks ({which})")
    std = a.alphabet == 256 and a.zipf == 1.1 and a.chunk == 65536
    which = ("BASELINE.json configs[1]" if std and a.bytes == 1 << 30 else
             "BASELINE.json configs[4] shard" if std and a.bytes == 8 << 30 else
             "BASELINE.json configs[3]" if a.alphabet == 4096 and a.zipf == 1.1 and a.bytes == 1 << 30 else "custom shape")
    return (f"{a.bytes / 2**30:g} GiB/GPU synthetic Zipf(s={a.zipf}) {'bytes' if sb == 1 else 'u16 symbols'}, "
            f"{a.alphabet}-symbol static global freq table, {kib} KiB chunks ({which})")


def thresholds(mod, a):
    """Generator thresholds of the workload: one table, or S_CYCLE's tables cycling per chunk."""
    import numpy as np

    if a.mode == "adaptive":
        return np.stack([mod.zipf_thresholds(a.alphabet, s) for s in S_CYCLE])
    return mod.zipf_thresholds(a.alphabet, a.zipf)


# ----------------------------------------------------------------- clocks
def nvml_handle(index):
    """NVML handle of torch's cuda:<index> (by UUID, so CUDA_VISIBLE_DEVICES does not matter)."""
    import pynvml
    import torch

    pynvml.nvmlInit()
    uuid = str(torch.cuda.get_device_properties(index).uuid)
    if not uuid.startswith("GPU-"):
        uuid = "GPU-" + uuid
    try:
        return pynvml.nvmlDeviceGetHandleByUUID(uuid)
    except Exception:
        return pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML from a thread of this process while the timed
    region runs (every ~2 ms; `nvidia-smi -lms` starts too slowly for a region of tens of ms)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.recording = False  # NVML is initialised (start) before warm-up; samples count from arm() on
        self.thread = None
        self.max_mhz = None
        self.error = None

    def arm(self):
        self.recording = True

    def _sample(self, pynvml, h):
        mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        if not self.recording:
            return
        self.samples.append(mhz)
        for bit, name in ((pynvml.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                          (pynvml.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                          (pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                          (pynvml.nvmlClocksEventReasonHwPowerBrakeSlowdown, "hw_power_brake_slowdown"),
                          (pynvml.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")):
            if r & bit:
                self.reasons.add(name)

    def _run(self, pynvml, h):
        while not self.stop_flag.is_set():
            try:
                self._sample(pynvml, h)
            except Exception as e:  # keep what we have
                self.error = repr(e)
                return
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml

            h = nvml_handle(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._run, args=(pynvml, h), daemon=True)
            self.thread.start()
        except Exception as e:
            self.error = repr(e)

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=5)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "samples": len(self.samples), "reasons": sorted(self.reasons)}
        if self.error:
            out["error"] = self.error
        return out


def bind_near_gpu(index):
    """Run this rank on the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end leg are allocated on the GPU's NUMA node.  Returns the previous affinity (to restore)."""
    try:
        import pynvml

        h = nvml_handle(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, x in enumerate(words) for b in range(64) if (x >> b) & 1}
        old = os.sched_getaffinity(0)
        cpus &= old
        if cpus:
            os.sched_setaffinity(0, cpus)
        return old
    except Exception:
        return None


# -------------------------------------------------------- reference arm (CPU)
class CpuPass:
    """One pass of the reference's algorithm (oracle port) over a batch on the host cores: histogram
    (all threads) -> table -> encode -> decode, one chunk per thread at a time, into preallocated
    buffers -- the timed window holds the algorithm only."""

    def __init__(self, oracle, a, syms, threads):
        self.o, self.a, self.syms, self.threads = oracle, a, syms, threads
        self.buf = oracle.RoundTripBuffers(syms.size, a.chunk, syms.dtype.itemsize)

    def model(self):
        import numpy as np

        o, a = self.o, self.a
        if a.mode == "adaptive":  # per-chunk histograms: one chunk per thread as well
            from concurrent.futures import ThreadPoolExecutor

            n_chunks = self.buf.n_chunks
            c = np.zeros((n_chunks, a.alphabet), dtype=np.uint32)

            def one(t):
                for j in range(t, n_chunks, self.threads):
                    c[j] = o.histogram(self.syms[j * a.chunk:(j + 1) * a.chunk], a.alphabet).astype(np.uint32)

            with ThreadPoolExecutor(self.threads) as ex:
                list(ex.map(one, range(self.threads)))
            cs = np.cumsum(c, axis=1, dtype=np.uint64)
            cum = np.zeros_like(c)
            cum[:, 1:] = cs[:, :-1].astype(np.uint32)
            return c, cum, cs[:, -1].astype(np.uint32)
        c, _ = o.normalise(o.histogram_mt(self.syms, a.alphabet, self.threads))
        cum, total = o.calc_cum(c)
        return c, cum, total

    def run(self):
        """(seconds histogram+table, seconds encode, seconds decode)"""
        t0 = time.perf_counter()
        c, cum, total = self.model()
        t1 = time.perf_counter()
        self.buf.encode(self.syms, c, cum, total, self.threads)
        t2 = time.perf_counter()
        back = self.buf.decode(c, cum, total, self.threads)
        t3 = time.perf_counter()
        assert (back == self.syms).all()
        return t1 - t0, t2 - t1, t3 - t2, (c, cum, total)


def calibrate_sample(oracle, a, threads, target_s):
    """Pick a chunk-aligned sample of the workload that takes ~target_s seconds per pass."""
    sb = 1 if a.alphabet <= 256 else 2
    n_syms = a.bytes // sb
    thr = thresholds(oracle, a)
    gen_chunk = a.chunk if a.mode == "adaptive" else 0
    probe_n = min(n_syms, max(a.chunk, threads * a.chunk * 2))
    syms = oracle.generate(probe_n, a.alphabet, a.seed, thr, sym_bytes=sb, chunk_syms=gen_chunk)
    th, te, td, _ = CpuPass(oracle, a, syms, threads).run()
    rate = probe_n / (th + te + td)
    n = int(rate * target_s) // a.chunk * a.chunk
    n = max(a.chunk * threads, min(n, n_syms))
    return n, thr, gen_chunk, sb


def run_reference(a):
    """`--impl reference`: the reference's CPU path on the host cores (C oracle port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # rank 0 alone runs the CPU arm
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as oracle

    threads = oracle.hardware_threads()
    n, thr, gen_chunk, sb = calibrate_sample(oracle, a, threads, target_s=2.0)
    syms = oracle.generate(n, a.alphabet, a.seed, thr, sym_bytes=sb, chunk_syms=gen_chunk)
    job = CpuPass(oracle, a, syms, threads)
    times = []
    for i in range(a.warmup + a.steps):
        th, te, td, _ = job.run()
        if i >= a.warmup:
            times.append((th + te + td, th, te, td))
    tot = sum(t[0] for t in times)
    nb = n * sb
    value = nb * len(times) / tot / 1e9
    enc = nb * len(times) / sum(t[2] for t in times) / 1e9
    dec = nb * len(times) / sum(t[3] for t in times) / 1e9
    sample = (f"first {nb / 2**20:.0f} MiB of the workload per step ({n // a.chunk} chunks), histogram+table+encode+decode, "
              f"preallocated buffers")
    line = {
        "impl": "reference", "metric": METRIC, "metric_definition": METRIC_DEFINITION, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": tot / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "chunk_syms": a.chunk, "alphabet": a.alphabet, "zipf_s": a.zipf,
                   "mode": a.mode, "bytes_per_gpu": a.bytes, "host": socket.gethostname()},
        "encode_gbs": enc, "decode_gbs": dec,
        "phase_ms": {"histogram_table": sum(t[1] for t in times) / len(times) * 1e3,
                     "encode": sum(t[2] for t in times) / len(times) * 1e3,
                     "decode": sum(t[3] for t in times) / len(times) * 1e3},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the Rust crate (no Rust toolchain in this image), "
                                 "one chunk per thread at a time, histogram on all threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ our arm
def measure_bus(torch, dev, barrier, nbytes=256 << 20, reps=3):
    """Pinned-memory copy rates of this rank's GPU with every rank copying at the same time: host->device
    alone, device->host alone, and both directions at once (the ceiling of the end-to-end leg)."""
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def wall(fn):
        best = 1e9
        for _ in range(reps):
            barrier()
            t = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t)
        return best

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    h2d = nbytes / wall(lambda: d_a.copy_(h_a, non_blocking=True)) / 1e9
    d2h = nbytes / wall(lambda: h_b.copy_(d_b, non_blocking=True)) / 1e9
    duplex = 2 * nbytes / wall(both) / 1e9
    return h2d, d2h, duplex


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import range_coder_rust_b200 as rcb
    from range_coder_rust_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: range_coder_rust_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op="max"):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, "max")

    ctx = rcb.Context(local_rank)
    # torch.distributed is plumbing (rendezvous, barriers, max-over-ranks); the exchange step ON the path
    # -- the all-reduce of the count table -- is issued by the library itself over its own communicator
    comm = sharding.init_comm(ctx)
    if a.enc_threads or a.dec_threads:
        ctx.set_block_threads(a.enc_threads, a.dec_threads)
    adaptive = a.mode == "adaptive"
    K, n, chunk = a.alphabet, a.bytes, a.chunk
    sym_bytes = 1 if K <= 256 else 2
    n_syms = n // sym_bytes
    n_chunks = (n_syms + chunk - 1) // chunk
    thr = thresholds(rcb, a)
    # this rank's shard of one global counter-based stream, generated on the device
    d_syms = ctx.generate(n_syms, K, a.seed, thr, sym_bytes=sym_bytes, first=rank * n_syms,
                          chunk_syms=chunk if adaptive else 0)

    def build_counts(out=None):
        if adaptive:  # one histogram per chunk: no exchange between GPUs at all
            return ctx.histogram(d_syms, K, chunk_syms=chunk, out=out)
        counts = ctx.histogram(d_syms, K, out=out)
        if comm is not None:
            ctx.allreduce_counts(counts, comm)  # the path's only exchange: K u64 counts over NVLink
        return counts

    # first pass (untimed): build the model once to size the buffers
    counts = build_counts()
    model = ctx.model_from_counts(counts)
    cap = ctx.encode_bound(model, n_syms, sym_bytes, chunk) + 16
    d_stream = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_offsets = torch.empty(n_chunks + 1, dtype=torch.int64, device=dev)
    d_back = torch.empty_like(d_syms)
    # restart points: the encoder records its state every a.restart symbols, the decoder runs that many more lanes
    d_restart = ctx.restart_points(n_chunks, chunk, a.restart) if a.restart else None
    rs_kw = {"restart_syms": a.restart, "restart": d_restart} if d_restart is not None else {}

    ev = {k: [] for k in ("start", "hist", "model", "enc", "dec")}
    kern = {"encode_kernel": [], "scan": [], "gather": [], "decode_kernel": []}

    def step(timed):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record()
        ctx.histogram(d_syms, K, chunk_syms=chunk if adaptive else 0, out=counts)
        e[1].record()
        if comm is not None and not adaptive:
            ctx.allreduce_counts(counts, comm)
        ctx.model_from_counts(counts, model=model)
        e[2].record()
        ctx.encode_chunks(d_syms, chunk, model, out=d_stream, offsets=d_offsets, sync=False, **rs_kw)
        e[3].record()
        ctx.decode_chunks(d_stream, d_offsets, n_syms, chunk, model, sym_bytes=sym_bytes, out=d_back, sync=False,
                          **rs_kw)
        e[4].record()
        if timed:
            for k, x in zip(("start", "hist", "model", "enc", "dec"), e):
                ev[k].append(x)
            t = ctx.timings()  # synchronises; per-kernel CUDA events recorded inside the library
            for k in kern:
                kern[k].append(t[k])

    ctx.enable_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()  # NVML start-up happens here, outside the timed region (it can stall other ranks' launches)
    for _ in range(a.warmup):
        step(False)
    barrier()
    sampler.arm()
    launches0 = ctx.launch_count
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(a.steps):
        step(True)
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    total_ms = max_over_ranks(t_start.elapsed_time(t_end))
    nbytes = ctx.encode_result()
    ctx.decode_result()
    assert torch.equal(d_back, d_syms), "round trip failed"

    # the same stream through the plain decoder (one lane per chunk, no side information), outside the timed
    # region: what a decoder that was handed the reference's bytes alone achieves
    plain = None
    if d_restart is not None:
        d_back.zero_()
        ts = []
        for _ in range(3):
            ctx.decode_chunks(d_stream, d_offsets, n_syms, chunk, model, sym_bytes=sym_bytes, out=d_back, sync=False)
            ts.append(ctx.timings()["decode_kernel"])
        ctx.decode_result()
        assert torch.equal(d_back, d_syms), "round trip (plain decoder) failed"
        plain = {"decode_kernel_ms": max_over_ranks(min(ts)),
                 "what": "rcb_decode_chunks on the same stream: one lane per chunk, no restart points (untimed leg)"}

    def phase(a_, b_):
        return sum(x.elapsed_time(y) for x, y in zip(ev[a_], ev[b_])) / len(ev[a_])

    if a.trace_steps:
        for i in range(len(ev["start"])):
            print(f"[rank {rank}] step {i}: " + " ".join(
                f"{nm}={ev[p0][i].elapsed_time(ev[p1][i]):.3f}" for nm, p0, p1 in
                (("hist", "start", "hist"), ("model", "hist", "model"), ("enc", "model", "enc"), ("dec", "enc", "dec"))),
                file=sys.stderr, flush=True)

    ms = {"histogram": phase("start", "hist"), "model": phase("hist", "model"), "encode": phase("model", "enc"),
          "decode": phase("enc", "dec")}
    ms = {k: max_over_ranks(v) for k, v in ms.items()}
    kavg = {k: max_over_ranks(sum(v) / len(v)) for k, v in kern.items()}
    ms_per_step = total_ms / a.steps
    value = world * n / (ms_per_step * 1e-3) / 1e9
    ratio = nbytes / n

    # ---- roofline of the dominant kernel.  HBM figure as the contract defines it: algorithmic bytes
    # (input + code: encode reads N and writes C, decode reads C and writes N) / measured duration.
    # The coder kernels are bound by one lane's instruction stream, not by HBM, so the line also carries the
    # issue-slot view: counters from the tracked ncu capture (profiles/issue.json), ceiling computed live.
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = "decode_kernel" if kavg["decode_kernel"] >= kavg["encode_kernel"] else "encode_kernel"
    alg_bytes = n + nbytes
    achieved = alg_bytes / (kavg[dom] * 1e-3) / 1e9
    tag = "static" if not adaptive and K <= 256 else ("adaptive" if adaptive else "k4096")
    if d_restart is not None:
        tag += "_restart"  # the tracked captures of the restart-point build (profiles/traffic.json, issue.json)

    def tracked(name):
        path = os.path.join(ROOT, "profiles", name)
        return json.load(open(path)) if os.path.exists(path) else {}

    tr = tracked("traffic.json").get(tag, {})
    # the captures are of the 1 GiB batch (restart points: 4 lanes per chunk)
    traffic = tr.get(dom) if n == 1 << 30 else None
    issue = None
    cap_issue = tracked("issue.json").get(tag, {}).get(dom)
    if cap_issue:
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
        sym_rate = n_syms / (kavg[dom] * 1e-3)
        ceiling = 4 * sm_count * mhz * 1e6 * 32 / cap_issue["inst_per_symbol"]  # symbols/s with every scheduler busy
        lanes_per_chunk = (chunk + a.restart - 1) // a.restart if (d_restart is not None and dom == "decode_kernel") else 1
        warps = (n_chunks * lanes_per_chunk + 31) // 32
        issue = dict(cap_issue)
        issue.update({"symbols_per_s": sym_rate, "issue_ceiling_symbols_per_s": ceiling,
                      "frac_of_issue_ceiling": sym_rate / ceiling,
                      "warps_per_scheduler": warps / (4 * sm_count),
                      "cycles_per_symbol_per_lane": kavg[dom] * 1e-3 * mhz * 1e6 / (min(chunk, n_syms) / lanes_per_chunk)})
    roofline = {"bound": "issue", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": tr.get("source") if traffic else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kavg[dom], "kernels_ms": kavg,
                "issue": issue,
                "note": "achieved/peak/frac are the HBM figures the contract asks for; the kernel is a per-lane "
                        "sequential coder bound by instruction issue and latency (see `issue` and DESIGN.md)"}

    # ---- end to end through the host-buffer C ABI (pinned host memory, copies inside the timed region)
    e2e = None
    if not a.no_e2e:
        import queue
        import threading as th_

        old_affinity = bind_near_gpu(local_rank)
        h_syms = torch.empty(n_syms, dtype=d_syms.dtype, pin_memory=True)
        h_syms.copy_(d_syms)
        a_syms = h_syms.numpy()
        if sym_bytes == 2:
            a_syms = a_syms.view(np.uint16)

        # "One ctx per GPU per thread" (include/rcb200.h): a.e2e_depth encoder threads and as many decoder threads,
        # each with its own context, so that the encodes of later steps run while earlier steps are decoded and
        # both directions of the link carry data all the time (one call alone keeps the bus ~65 % busy, two of
        # each kind ~83 %; three were measured slower).  Every step still copies its input host->device and its
        # results device->host inside its two calls; a step's decode starts when its own encode has returned (it
        # needs that call's offsets and byte count).
        depth = max(1, a.e2e_depth)
        c_, cum_, total_, _f = (counts, None, None, None) if adaptive else model.tables()

        def make_ctx():
            cx = rcb.Context(local_rank, stream=torch.cuda.Stream(dev))
            return cx, (cx.model_from_counts(counts) if adaptive else cx.model_from_tables(c_, cum_, total_))

        encs = [(ctx, model)] + [make_ctx() for _ in range(depth - 1)]
        decs = [make_ctx() for _ in range(depth)]
        n_buf = 2 * depth  # code-stream buffers: free again when their step has been decoded
        rs_e2e = a.restart if d_restart is not None else 0
        rs_words = d_restart.numel() if d_restart is not None else 0
        h_streams = [torch.empty(cap, dtype=torch.uint8, pin_memory=True) for _ in range(n_buf)]
        h_rs = [torch.zeros(max(1, rs_words), dtype=torch.int64, pin_memory=True) for _ in range(n_buf)]
        h_backs = [torch.empty(n_syms, dtype=d_syms.dtype, pin_memory=True) for _ in range(depth)]
        streams = [t.numpy() for t in h_streams]
        rs_np = [t.numpy().view(np.uint64) for t in h_rs]
        backs = [t.numpy().view(np.uint16) if sym_bytes == 2 else t.numpy() for t in h_backs]
        result = {}

        def enc_host(cx, m_, slot):
            r = cx.encode_host(a_syms, chunk, m_, out_np=streams[slot], restart_syms=rs_e2e,
                               restart_np=rs_np[slot] if rs_e2e else None)
            return r[1], r[2]  # offsets, bytes

        def dec_host(cx, m_, slot, offs, nb, out_np):
            cx.decode_host(streams[slot][:nb], offs, n_syms, chunk, m_, sym_bytes=sym_bytes, out_np=out_np,
                           restart_syms=rs_e2e, restart_np=rs_np[slot] if rs_e2e else None)

        def run_steps(k_steps):
            q, free = queue.Queue(), queue.Queue()
            for b_ in range(n_buf):
                free.put(b_)
            nxt, lock = [0], th_.Lock()

            def decoder(j):
                cx, m_ = decs[j]
                with torch.cuda.device(dev):
                    while True:
                        item = q.get()
                        if item is None:
                            return
                        slot, offs, nb = item
                        dec_host(cx, m_, slot, offs, nb, backs[j])
                        result["nb"], result["offs"] = nb, offs
                        free.put(slot)

            def encoder(j):
                cx, m_ = encs[j]
                with torch.cuda.device(dev):
                    while True:
                        with lock:
                            i = nxt[0]
                            nxt[0] += 1
                        if i >= k_steps:
                            return
                        slot = free.get()
                        offs, nb = enc_host(cx, m_, slot)
                        q.put((slot, offs, nb))

            dts = [th_.Thread(target=decoder, args=(j,)) for j in range(depth)]
            ets = [th_.Thread(target=encoder, args=(j,)) for j in range(depth)]
            for t in dts + ets:
                t.start()
            for t in ets:
                t.join()
            for _ in dts:
                q.put(None)
            for t in dts:
                t.join()

        run_steps(n_buf)  # warm-up: sizes the device scratch of every context
        barrier()
        t0 = time.perf_counter()
        run_steps(a.e2e_steps)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        nb, offs = result["nb"], result["offs"]
        for b_ in backs:
            assert np.array_equal(b_, a_syms)
        # the same two calls back to back on one context (no overlap between steps), for reference
        t1 = time.perf_counter()
        offs1, nb1 = enc_host(ctx, model, 0)
        dec_host(ctx, model, 0, offs1, nb1, backs[0])
        dt_serial = max_over_ranks(time.perf_counter() - t1)
        barrier()
        off_bytes = (n_chunks + 1) * 8 + rs_words * 8  # offsets + restart points
        # the bus ceiling on this box, every rank copying at once (aggregate over ranks)
        h2d, d2h, duplex = measure_bus(torch, dev, barrier)
        bus = {"h2d_gbs": reduce_ranks(h2d, "sum"), "d2h_gbs": reduce_ranks(d2h, "sum"),
               "duplex_gbs": reduce_ranks(duplex, "sum"), "ranks": world,
               "how": "256 MiB pinned copies, all ranks at the same time, best of 3; duplex = both directions at once"}
        moved = world * (2 * n + 2 * nb + 2 * off_bytes)  # bytes over the bus per step, both directions, all ranks
        e2e = {"value": world * n * a.e2e_steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(world * (n + nb + off_bytes)),
               "d2h_bytes_per_step": int(world * (nb + off_bytes + n)),
               "bytes_are": "aggregate over all ranks, like `value`",
               "steps": a.e2e_steps, "ms_per_step": dt / a.e2e_steps * 1e3,
               "serial_ms_per_step": dt_serial * 1e3, "serial_value": world * n / dt_serial / 1e9,
               "pipelining": f"{depth} encode and {depth} decode calls in flight ({2 * depth} contexts, one host thread "
                             "each): later steps are encoded while earlier ones are decoded; serial_* = the two calls "
                             "back to back on one context",
               "bus": bus, "bus_gbs": moved * a.e2e_steps / dt / 1e9,
               "bus_frac": moved * a.e2e_steps / dt / 1e9 / bus["duplex_gbs"],
               "api": ("rcb_encode_host_restart + rcb_decode_host_restart" if rs_e2e else
                       "rcb_encode_host + rcb_decode_host") + " (C ABI, pinned host buffers)"}
        if old_affinity:
            os.sched_setaffinity(0, old_affinity)
        for cx, m_ in encs[1:] + decs:
            m_.close()
            cx.close()
        del h_syms, h_streams, h_backs, h_rs, streams, rs_np, backs

    # ---- parity on EVERY rank at every N: a deterministic >= 1 % subset of this rank's chunks (every 64th,
    # starting at 5) re-encoded by the oracle under this rank's table and compared byte for byte
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as oracle  # checker only

    host_threads = max(1, oracle.hardware_threads() // world)
    offs_np = d_offsets.cpu().numpy().astype(np.uint64)
    parity = None
    if not a.no_parity:
        picked = np.arange(5 % max(1, n_chunks), n_chunks - (1 if n_syms % chunk else 0), 64)
        idx = torch.from_numpy(picked).to(dev)
        full = d_syms[: (n_syms // chunk) * chunk].view(-1, chunk)
        sample = full[idx].reshape(-1).cpu().numpy()
        if sym_bytes == 2:
            sample = sample.view(np.uint16)
        if adaptive:
            pc = counts[idx].cpu().numpy().view(np.uint32)
            cs = np.cumsum(pc, axis=1, dtype=np.uint64)
            pcum = np.zeros_like(pc)
            pcum[:, 1:] = cs[:, :-1].astype(np.uint32)
            tabs = (pc, pcum, cs[:, -1].astype(np.uint32))
        else:
            c, cum, total, _ = model.tables()
            tabs = (c, cum, total)
        ref_stream, ref_offsets = oracle.encode_chunks(sample, chunk, *tabs, threads=host_threads)
        ok = True
        for k, i in enumerate(picked):
            got = d_stream[int(offs_np[i]):int(offs_np[i + 1])].cpu().numpy()
            if not np.array_equal(got, ref_stream[int(ref_offsets[k]):int(ref_offsets[k + 1])]):
                ok = False
                print(f"[rank {rank}] chunk {i} differs from the oracle", file=sys.stderr, flush=True)
                break
        all_ok = reduce_ranks(1.0 if ok else 0.0, "min") == 1.0
        parity = {"chunks_checked": int(reduce_ranks(float(picked.size), "sum")), "bit_exact": bool(all_ok),
                  "ranks": world, "chunks_per_rank": n_chunks,
                  "subset": "every 64th chunk of every rank's shard (from chunk 5), oracle re-encode, bytes compared",
                  "round_trip_all_chunks": True}

    # ---- CPU baseline (rank 0, N=1): oracle port on a bounded sample of the same batch; compares ALL its chunks
    cpu = None
    if not a.no_cpu and rank == 0 and world == 1:
        threads = oracle.hardware_threads()
        probe = d_syms[: min(n_syms, threads * chunk * 2)].cpu().numpy()
        if sym_bytes == 2:
            probe = probe.view(np.uint16)
        th, te, td, _ = CpuPass(oracle, a, probe, threads).run()
        rate = probe.size / (th + te + td)
        ns = max(chunk * threads, min(int(rate * 8.0) // chunk * chunk, n_syms // chunk * chunk))
        sample = d_syms[:ns].cpu().numpy()
        if sym_bytes == 2:
            sample = sample.view(np.uint16)
        job = CpuPass(oracle, a, sample, threads)
        if adaptive:
            th, te, td, tabs = job.run()
        else:  # the batch's table (all N symbols), as the GPU used it: histogram timed on the sample
            t0 = time.perf_counter()
            oracle.histogram_mt(sample, K, threads)
            th = time.perf_counter() - t0
            c, cum, total, _ = model.tables()
            t1 = time.perf_counter()
            job.buf.encode(sample, c, cum, total, threads)
            t2 = time.perf_counter()
            back = job.buf.decode(c, cum, total, threads)
            t3 = time.perf_counter()
            assert (back == sample).all()
            te, td = t2 - t1, t3 - t2
        nb_s = ns * sym_bytes
        cpu = {"value": nb_s / (th + te + td) / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {nb_s / 2**20:.0f} MiB ({ns // chunk} chunks) of the same batch, histogram+encode+decode",
               "encode_gbs": nb_s / te / 1e9, "decode_gbs": nb_s / td / 1e9, "histogram_ms": th * 1e3,
               "note": "C restatement of the Rust crate (no Rust toolchain in this image), one chunk per thread"}
        ref_stream, ref_offsets = job.buf.stream()
        k = ns // chunk
        got = d_stream[: int(offs_np[k])].cpu().numpy()
        ok = bool(np.array_equal(offs_np[: k + 1], ref_offsets) and np.array_equal(got, ref_stream))
        if parity is not None:
            parity["cpu_baseline_chunks_checked"] = int(k)
            parity["bit_exact"] = bool(parity["bit_exact"] and ok)
        assert ok, "GPU stream differs from the oracle on the CPU baseline's chunks"
    if parity is not None:
        assert parity["bit_exact"], "GPU stream differs from the oracle on the sampled chunks"

    if rank == 0:
        if adaptive:
            par = f"chunks sharded over {world} GPU(s); per-chunk tables, no exchange" if world > 1 else "single GPU"
        else:
            par = (f"chunks sharded over {world} GPU(s); one all-reduce of {K} u64 counts (rcb_allreduce_counts, NCCL "
                   f"{comm.nccl_version})") if world > 1 else "single GPU"
        line = {
            "metric": METRIC, "metric_definition": METRIC_DEFINITION, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(a), "chunk_syms": chunk, "alphabet": K, "zipf_s": a.zipf, "mode": a.mode,
                       "bytes_per_gpu": n, "n_chunks_per_gpu": n_chunks, "compressed_over_input": ratio,
                       "restart_syms": a.restart if d_restart is not None else 0,
                       "decoder_lanes_per_chunk": (chunk + a.restart - 1) // a.restart if d_restart is not None else 1,
                       "restart_bytes_over_input": (d_restart.numel() * 8 / n) if d_restart is not None else 0.0,
                       "l2": "inputs larger than L2 (1 GiB batch + 0.72 GiB stream per GPU vs 126 MB); no flush",
                       "parallelism": par},
            "encode_gbs": world * n / (ms["encode"] * 1e-3) / 1e9,
            "decode_gbs": world * n / (ms["decode"] * 1e-3) / 1e9,
            "phase_ms": ms,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "parity": parity, "plain_decoder": plain,
        }
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)
    if a.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch one rank per GPU the way the driver does
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
