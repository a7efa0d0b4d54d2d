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
    p.add_argument("--seed", type=lambda s: int(s, 0), default=0x5EED0001)
    p.add_argument("--e2e-steps", type=int, default=2)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--trace-steps", action="store_true", help="per-step phase times of this rank on stderr")
    p.add_argument("--enc-threads", type=int, default=0)
    p.add_argument("--dec-threads", type=int, default=0)
    return p.parse_args()


def workload_name(a):
    std = a.alphabet == 256 and a.zipf == 1.1 and a.chunk == 65536
    which = ("BASELINE.json configs[1]" if std and a.bytes == 1 << 30 else
             "BASELINE.json configs[4] shard" if std and a.bytes == 8 << 30 else "custom shape")
    return (f"{a.bytes / 2**30:g} GiB/GPU synthetic Zipf(s={a.zipf}) bytes, {a.alphabet}-symbol static global "
            f"freq table, {a.chunk // 1024} KiB chunks ({which})")


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
def cpu_roundtrip(oracle, syms, chunk, c, cum, total, threads):
    """One pass of the reference's algorithm (oracle port) over `syms`; returns
    (encode seconds, decode seconds, stream, offsets)."""
    t0 = time.perf_counter()
    stream, offsets = oracle.encode_chunks(syms, chunk, c, cum, total, threads=threads)
    t1 = time.perf_counter()
    dec, _ = oracle.decode_chunks(stream, offsets, syms.size, chunk, c, cum, total, threads=threads)
    t2 = time.perf_counter()
    assert (dec == syms).all()
    return t1 - t0, t2 - t1, stream, offsets


def calibrate_sample(oracle, a, threads, target_s):
    """Pick a chunk-aligned sample of the workload that takes ~target_s seconds per pass."""
    import numpy as np

    thr = oracle.zipf_thresholds(a.alphabet, a.zipf)
    probe_n = min(a.bytes, max(a.chunk, threads * a.chunk * 2))
    syms = oracle.generate(probe_n, a.alphabet, a.seed, thr)
    c, cum, total = oracle.model_from_symbols(syms, a.alphabet)
    te, td, _, _ = cpu_roundtrip(oracle, syms, a.chunk, c, cum, total, threads)
    rate = probe_n / (te + td)
    n = int(rate * target_s) // a.chunk * a.chunk
    n = max(a.chunk * threads, min(n, a.bytes))
    return n, thr


def run_reference(a):
    """`--impl reference`: the reference's CPU path on the host cores (C oracle port)."""
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # rank 0 alone runs the CPU arm
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as oracle

    threads = oracle.hardware_threads()
    n, thr = calibrate_sample(oracle, a, threads, target_s=2.0)
    syms = oracle.generate(n, a.alphabet, a.seed, thr)
    c, cum, total = oracle.model_from_symbols(syms, a.alphabet)
    times = []
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        c, cum, total = oracle.model_from_symbols(syms, a.alphabet)
        te, td, _, _ = cpu_roundtrip(oracle, syms, a.chunk, c, cum, total, threads)
        t = time.perf_counter() - t0
        if i >= a.warmup:
            times.append((t, te, td))
    tot = sum(t for t, _, _ in times)
    value = n * len(times) / tot / 1e9
    enc = n * len(times) / sum(te for _, te, _ in times) / 1e9
    dec = n * len(times) / sum(td for _, _, td in times) / 1e9
    sample = f"first {n / 2**20:.0f} MiB of the workload per step ({n // a.chunk} chunks), histogram+encode+decode"
    line = {
        "impl": "reference", "metric": METRIC, "metric_definition": METRIC_DEFINITION, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": tot / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "chunk_syms": a.chunk, "alphabet": a.alphabet, "zipf_s": a.zipf,
                   "bytes_per_gpu": a.bytes, "host": socket.gethostname()},
        "encode_gbs": enc, "decode_gbs": dec,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the Rust crate (no Rust toolchain in this image), "
                                 "one chunk per thread at a time"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ our arm
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import range_coder_rust_b200 as rcb

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

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = rcb.Context(local_rank)
    if a.enc_threads or a.dec_threads:
        ctx.set_block_threads(a.enc_threads, a.dec_threads)
    K, n, chunk = a.alphabet, a.bytes, a.chunk
    sym_bytes = 1 if K <= 256 else 2
    n_syms = n // sym_bytes
    n_chunks = (n_syms + chunk - 1) // chunk
    thr = rcb.zipf_thresholds(K, a.zipf)
    # this rank's shard of one global counter-based stream, generated on the device
    d_syms = ctx.generate(n_syms, K, a.seed, thr, sym_bytes=sym_bytes, first=rank * n_syms)

    counts = torch.empty(K, dtype=torch.int64, device=dev)
    model = None
    # first pass (untimed): build the model once to size the buffers
    ctx.histogram(d_syms, K, out=counts)
    if world > 1:
        dist.all_reduce(counts)
    model = ctx.model_from_counts(counts)
    cap = ctx.encode_bound(model, n_syms, sym_bytes, chunk) + 16
    d_stream = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_offsets = torch.empty(n_chunks + 1, dtype=torch.int64, device=dev)
    d_back = torch.empty_like(d_syms)

    ev = {k: [] for k in ("start", "hist", "model", "enc", "dec")}
    kern = {"encode_kernel": [], "scan": [], "gather": [], "decode_kernel": []}

    def step(timed):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record()
        ctx.histogram(d_syms, K, out=counts)
        e[1].record()
        if world > 1:
            dist.all_reduce(counts)  # the path's only exchange: K u64 counts over NVLink
        ctx.model_from_counts(counts, model=model)
        e[2].record()
        ctx.encode_chunks(d_syms, chunk, model, out=d_stream, offsets=d_offsets, sync=False)
        e[3].record()
        ctx.decode_chunks(d_stream, d_offsets, n_syms, chunk, model, sym_bytes=sym_bytes, out=d_back, sync=False)
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

    # roofline of the dominant kernel: algorithmic bytes (input + code) / measured duration
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = "decode_kernel" if kavg["decode_kernel"] >= kavg["encode_kernel"] else "encode_kernel"
    alg_bytes = n + nbytes  # encode: read N, write C; decode: read C, write N
    achieved = alg_bytes / (kavg[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kavg[dom],
                "kernels_ms": kavg,
                "note": "per-lane sequential coder: latency/issue-bound, see DESIGN.md for the issue ceiling"}

    # ---- end to end through the host-buffer C ABI (pinned host memory, copies inside the timed region)
    e2e = None
    if not a.no_e2e:
        old_affinity = bind_near_gpu(local_rank)
        h_syms = torch.empty(n_syms, dtype=d_syms.dtype, pin_memory=True)
        h_syms.copy_(d_syms)
        h_stream = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        h_back = torch.empty(n_syms, dtype=d_syms.dtype, pin_memory=True)
        a_syms, a_stream, a_back = h_syms.numpy(), h_stream.numpy(), h_back.numpy()
        if sym_bytes == 2:
            a_syms, a_back = a_syms.view(np.uint16), a_back.view(np.uint16)

        def e2e_step():
            _, offs, nb = ctx.encode_host(a_syms, chunk, model, out_np=a_stream)
            ctx.decode_host(a_stream[:nb], offs, n_syms, chunk, model, sym_bytes=sym_bytes, out_np=a_back)
            return nb, offs

        e2e_step()  # warm-up: sizes the device scratch
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            nb, offs = e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        assert np.array_equal(a_back, a_syms)
        off_bytes = (n_chunks + 1) * 8
        e2e = {"value": world * n * a.e2e_steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(n + nb + off_bytes), "d2h_bytes_per_step": int(nb + off_bytes + n),
               "steps": a.e2e_steps, "ms_per_step": dt / a.e2e_steps * 1e3,
               "api": "rcb_encode_host + rcb_decode_host (C ABI, pinned host buffers)"}
        if old_affinity:
            os.sched_setaffinity(0, old_affinity)

    # ---- CPU baseline (rank 0, N=1): oracle port on a bounded sample of the same workload; doubles as parity check
    cpu = None
    parity = None
    if not a.no_cpu and rank == 0 and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_bind as oracle

        threads = oracle.hardware_threads()
        c, cum, total, _ = model.tables()
        probe = d_syms[: min(n_syms, threads * chunk * 2)].cpu().numpy()
        te, td, _, _ = cpu_roundtrip(oracle, probe, chunk, c, cum, total, threads)
        rate = probe.size / (te + td)
        ns = max(chunk * threads, min(int(rate * 8.0) // chunk * chunk, n_syms))
        sample = d_syms[:ns].cpu().numpy()
        te, td, ref_stream, ref_offsets = cpu_roundtrip(oracle, sample, chunk, c, cum, total, threads)
        cpu = {"value": ns / (te + td) / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {ns / 2**20:.0f} MiB ({ns // chunk} chunks) of the same batch, encode+decode",
               "encode_gbs": ns / te / 1e9, "decode_gbs": ns / td / 1e9,
               "note": "C restatement of the Rust crate (no Rust toolchain in this image), one chunk per thread"}
        k = ns // chunk
        got_off = d_offsets[: k + 1].cpu().numpy().astype(np.uint64)
        got = d_stream[: int(got_off[-1])].cpu().numpy()
        ok = bool(np.array_equal(got_off, ref_offsets) and np.array_equal(got, ref_stream))
        parity = {"chunks_checked": int(k), "bit_exact": ok, "round_trip_all_chunks": True}
        assert ok, "GPU stream differs from the oracle on the sampled chunks"

    if rank == 0:
        line = {
            "metric": METRIC, "metric_definition": METRIC_DEFINITION, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(a), "chunk_syms": chunk, "alphabet": K, "zipf_s": a.zipf,
                       "bytes_per_gpu": n, "n_chunks_per_gpu": n_chunks, "compressed_over_input": ratio,
                       "l2": "inputs larger than L2 (1 GiB batch + 0.72 GiB stream per GPU vs 126 MB); no flush",
                       "parallelism": f"chunks sharded over {world} GPU(s); one all-reduce of {K} u64 counts"
                       if world > 1 else "single GPU"},
            "encode_gbs": world * n / (ms["encode"] * 1e-3) / 1e9,
            "decode_gbs": world * n / (ms["decode"] * 1e-3) / 1e9,
            "phase_ms": ms,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "parity": parity,
        }
        print(json.dumps(line))
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
