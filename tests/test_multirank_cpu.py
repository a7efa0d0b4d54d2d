"""N > 1 host logic on CPU: two `gloo` ranks shard one stream by chunks, all-reduce
their symbol counts, and must derive the same global table.  The
coding kernels need a GPU; here the per-rank histograms come from numpy and the
table rule from the oracle (checker), which is what every rank's GPU computes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, chunk, K, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as oracle
    from range_coder_rust_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    thr = oracle.zipf_thresholds(K, 1.1)
    lo, hi = sharding.shard_symbols(n, chunk, rank, world)
    # each rank generates only its shard of the one global counter-based stream
    local = oracle.generate(hi - lo, K, 0x5EED0001, thr, first=lo, threads=1)
    counts = torch.from_numpy(np.bincount(local, minlength=K).astype(np.int64))
    sharding.allreduce_counts(counts)
    c, _ = oracle.normalise(counts.numpy().astype(np.uint64))
    cum, total = oracle.calc_cum(c)
    # every chunk belongs to exactly one rank; compressed sizes give global stream offsets
    stream, offsets = oracle.encode_chunks(local, chunk, c, cum, total, threads=1)
    start, whole = sharding.global_offsets(stream.size)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, c=c, cum=cum, total=total, start=start,
             whole=whole, size=stream.size, stream=stream, offsets=offsets)
    dist.destroy_process_group()


def test_two_ranks_share_one_table_and_tile_the_stream(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as oracle

    world, n, chunk, K = 2, 10 * 4096 + 777, 4096, 256
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, n, chunk, K, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(world)]
    # shards tile [0, n) at chunk boundaries
    assert r[0]["lo"] == 0 and r[0]["hi"] == r[1]["lo"] and r[1]["hi"] == n and r[0]["hi"] % chunk == 0
    # identical tables on both ranks == the table of the whole stream
    thr = oracle.zipf_thresholds(K, 1.1)
    whole = oracle.generate(n, K, 0x5EED0001, thr)
    c, cum, total = oracle.model_from_symbols(whole, K)
    for x in r:
        assert x["total"] == total and np.array_equal(x["c"], c) and np.array_equal(x["cum"], cum)
    # per-rank segments concatenate to the single-process stream
    ref_stream, ref_offsets = oracle.encode_chunks(whole, chunk, c, cum, total)
    assert r[0]["start"] == 0 and r[1]["start"] == r[0]["size"]
    assert r[0]["whole"] == r[1]["whole"] == ref_stream.size
    assert np.array_equal(np.concatenate([r[0]["stream"], r[1]["stream"]]), ref_stream)
    k0 = int(r[0]["hi"]) // chunk
    assert np.array_equal(r[0]["offsets"], ref_offsets[:k0 + 1])
    assert np.array_equal(r[1]["offsets"] + np.uint64(r[1]["start"]), ref_offsets[k0:])


def test_shard_chunks_partition():
    from range_coder_rust_b200 import sharding

    for n_chunks in (0, 1, 7, 16384, 131071):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_chunks(n_chunks, rk, world) for rk in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_chunks
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
