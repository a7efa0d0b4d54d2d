// rcb_encode.cuh -- K3: chunk-parallel encode (src/encoder.rs:24-46 over
// src/range_coder.rs:53-135), one independent reference `Encoder` per lane.
//
// With 64 KiB chunks a 1 GiB batch is only 16384 lanes (512 warps on 592 warp
// schedulers), so run time = symbols per chunk x the latency of one lane's
// per-symbol instruction stream (in-order issue, one warp per scheduler).  The
// loop is organised to keep everything except the (lower, range) recurrence
// off that stream's critical path and free of branches:
//   - symbols arrive as 16-byte vectors, one vector of lookahead in registers,
//     1 KiB bulk L2 prefetches ahead of that (a lane walks its own chunk, so a
//     lone 16-byte load would be a random 32-byte DRAM sector);
//   - the {cum, c} entries of the next 32-bit word of symbols are fetched from
//     shared memory while the current word is coded (ping-pong eA / eB);
//   - byte emission is two funnel shifts plus a predicated 32-bit store, issued
//     one symbol late so it overlaps the next symbol's multiply chain;
//   - FUSED (power-of-two total >= 2^24, consistent table): range/total and the
//     renormalisation shift are one shift (rcb_core.cuh: fused_step), and the
//     rare events that need the reference's literal loops (loop 2, n1 >= 3) are
//     handled per 32-bit word: the word is coded speculatively without any
//     branch, and if a lane saw such an event it restores its checkpoint and
//     re-codes the word on the exact out-of-line path (stores are idempotent).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcb_core.cuh"

namespace rcb {

struct EncodeArgs {
    const void* syms;
    uint64_t n_syms;
    uint64_t chunk_syms;
    uint64_t n_chunks;
    const uint2* tabs;      // [n_models][K]
    const ModelHdr* hdrs;   // [n_models]
    uint32_t K;
    uint32_t per_chunk;     // 1: model index = chunk
    uint8_t* staging;       // [n_chunks][pitch]
    uint64_t pitch;
    uint32_t* lens;         // [n_chunks]
    uint32_t* status;       // [n_chunks]
};

struct RowStore {
    uint8_t* row;
    __device__ __forceinline__ void word_if(bool p, uint32_t pos, uint32_t w) const {
        asm volatile(
            "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.b32 [%1], %2;\n\t}"
            :
            : "r"((uint32_t)p), "l"(row + pos), "r"(w)
            : "memory");
    }
    __device__ __forceinline__ void byte(uint32_t pos, uint32_t b) const { row[pos] = (uint8_t)b; }
};

__device__ __forceinline__ uint4 ldg_stream_v4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Bulk L2 prefetch of the next `bytes` (multiple of 16) of a lane's row.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int SPW>
struct EncEntries {
    uint2 e[SPW];
};

// Lane state of the fused loop; travels by value through the exact re-code path.
struct EncWordState {
    uint64_t lo, rpt;
    uint32_t pend, nb, pos;  // byte sink
    uint32_t em_hi, em_sh;   // deferred emission of the previous symbol
    uint32_t err;
};

// Exact re-code of one word of symbols from a checkpoint (rare): fused arithmetic where it
// applies, the reference's literal loops (renorm_slow) where it does not.  Out of line so the
// unrolled hot loop stays small for the instruction cache.
template <int SPW>
__device__ __noinline__ EncWordState enc_word_exact(EncWordState s, EncEntries<SPW> en, FusedParams fp,
                                                    uint8_t* row, uint32_t cap) {
    RowStore rs{row};
    EncSink<RowStore, true> sink(rs, cap);
    sink.pend = s.pend;
    sink.nb = s.nb;
    sink.pos = s.pos;
#pragma unroll 1
    for (int b = 0; b < SPW; b++) {
        uint64_t nlo, rgp, nrpt;
        uint32_t sh;
        const bool ok = fused_step(s.lo, s.rpt, en.e[b].x, en.e[b].y, fp, nlo, rgp, nrpt, sh);
        sink.put(s.em_hi, s.em_sh);
        if (ok) {
            s.em_hi = hi32(nlo);
            s.em_sh = sh;
            s.lo = nlo << sh;
            s.rpt = nrpt;
        } else {
            s.em_sh = 0;
            uint64_t lo = nlo, rg = rgp;
            renorm_slow<false>(lo, rg, sink, s.err);
            s.lo = lo;
            s.rpt = rg >> fp.s;
        }
    }
    s.pend = sink.pend;
    s.nb = sink.nb;
    s.pos = sink.pos;
    return s;
}

// SHARED: one table for all chunks, staged in shared memory.
// !SHARED: one table per chunk, read through L1/L2 from global memory.
template <typename SYM, bool SHARED, bool POW2, bool CHECKED, bool RANGECHK, bool FUSED>
__global__ void __launch_bounds__(256, 1) encode_kernel(EncodeArgs a) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    uint2* s_tab = reinterpret_cast<uint2*>(s_raw);
    __shared__ ModelHdr s_hdr;
    if (SHARED) {
        for (uint32_t i = threadIdx.x; i < a.K; i += blockDim.x) s_tab[i] = a.tabs[i];
        if (threadIdx.x == 0) s_hdr = a.hdrs[0];
        __syncthreads();
    }
    const uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk >= a.n_chunks) return;
    const uint64_t first = chunk * a.chunk_syms;
    const uint64_t cnt = (a.n_syms - first < a.chunk_syms) ? (a.n_syms - first) : a.chunk_syms;
    const SYM* src = reinterpret_cast<const SYM*>(a.syms) + first;

    const uint2* tab = SHARED ? s_tab : a.tabs + chunk * a.K;
    DivParams div;
    bool pow2 = POW2;
    if (SHARED) {
        div = s_hdr.div;
    } else {
        ModelHdr h = a.hdrs[chunk];
        div = h.div;
        pow2 = (h.flags & MODEL_POW2) != 0;
    }
    const uint32_t K = a.K;
    const uint32_t cap = (uint32_t)a.pitch;

    uint64_t lo = 0, rg = ~0ull;  // src/range_coder.rs:13-20
    uint32_t err = 0;
    RowStore rs{a.staging + chunk * a.pitch};
    EncSink<RowStore, true> sink(rs, cap);

    constexpr int SPW = 4 / sizeof(SYM);  // symbols per 32-bit word
    using Entries = EncEntries<SPW>;
    auto lookup = [&](uint32_t w) -> Entries {
        Entries r;
#pragma unroll
        for (int b = 0; b < SPW; b++) {
            uint32_t s = sizeof(SYM) == 1 ? ((w >> (8 * b)) & 0xFFu) : ((w >> (16 * b)) & 0xFFFFu);
            if (RANGECHK && s >= K) {
                if (!err) err = ST_SYMBOL_RANGE;
                s = 0;
            }
            r.e[b] = tab[s];
        }
        return r;
    };
    auto generic_symbol = [&](uint2 e) {
        if (SHARED) {
            update_symbol<POW2, CHECKED>(lo, rg, e.x, e.y, div, sink, err);
        } else {
            if (pow2)
                update_symbol<true, CHECKED>(lo, rg, e.x, e.y, div, sink, err);
            else
                update_symbol<false, CHECKED>(lo, rg, e.x, e.y, div, sink, err);
        }
    };

    constexpr uint32_t PER = 16 / sizeof(SYM);
    constexpr uint64_t PF_VECS = 64;  // L2 prefetch granule: 64 vectors = 1 KiB of the lane's row
    uint64_t done = 0;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const uint4* v = reinterpret_cast<const uint4*>(src);
        const uint64_t nvec = cnt / PER;
        auto prefetch_block = [&](uint64_t i) {  // at the first vector of a granule: fetch the next one
            if ((i & (PF_VECS - 1)) == 0) {
                const uint64_t ahead = i + PF_VECS;
                if (ahead < nvec) {
                    const uint64_t left = (nvec - ahead) * 16;
                    prefetch_l2_bulk(v + ahead, (uint32_t)(left < PF_VECS * 16 ? left : PF_VECS * 16));
                }
            }
        };
        if (nvec) {
            prefetch_l2_bulk(v, (uint32_t)(nvec < PF_VECS ? nvec * 16 : PF_VECS * 16));
            uint4 cur = ldg_stream_v4(v);
            Entries eA = lookup(cur.x);
            uint64_t i = 0;
            if constexpr (FUSED) {
                // fast sink: no capacity test per store; room for a whole vector is checked once
                // per vector (16 symbols x at most 15 bytes + the deferred emission < 320 bytes)
                EncSink<RowStore, false> fs(rs, cap);
                const FusedParams fp{div.shift, div.shift - 24u, 1u << (48u - div.shift)};
                uint64_t rpt = rg >> fp.s;
                uint32_t em_hi = 0, em_sh = 0;  // previous symbol's bytes, emitted one symbol late
                auto code = [&](const Entries& en) {
                    const EncWordState chk{lo, rpt, fs.pend, fs.nb, fs.pos, em_hi, em_sh, err};
                    bool bad = false;
#pragma unroll
                    for (int b = 0; b < SPW; b++) {  // speculative: straight-line, no branch
                        uint64_t nlo, rgp, nrpt;
                        uint32_t sh;
                        const bool ok = fused_step(lo, rpt, en.e[b].x, en.e[b].y, fp, nlo, rgp, nrpt, sh);
                        fs.put(em_hi, em_sh);
                        em_hi = hi32(nlo);
                        em_sh = sh;
                        lo = nlo << sh;
                        rpt = nrpt;
                        bad |= !ok;
                    }
                    if (RCB_UNLIKELY(bad)) {  // restore the checkpoint and re-code the word exactly
                        const EncWordState r = enc_word_exact<SPW>(chk, en, fp, rs.row, cap);
                        lo = r.lo;
                        rpt = r.rpt;
                        fs.pend = r.pend;
                        fs.nb = r.nb;
                        fs.pos = r.pos;
                        em_hi = r.em_hi;
                        em_sh = r.em_sh;
                        err = r.err;
                    }
                };
#pragma unroll 1
                for (; i < nvec; i++) {
                    if (fs.pos + 320u > cap) break;  // finish this chunk on the capacity-checked path
                    prefetch_block(i);
                    const uint4 nxt = (i + 1 < nvec) ? ldg_stream_v4(v + i + 1) : make_uint4(0, 0, 0, 0);
                    Entries eB = lookup(cur.y);
                    code(eA);
                    eA = lookup(cur.z);
                    code(eB);
                    eB = lookup(cur.w);
                    code(eA);
                    eA = lookup(nxt.x);  // word 0 of the next vector (zeros past the end: entry 0, unused)
                    code(eB);
                    cur = nxt;
                }
                sink.pend = fs.pend;
                sink.nb = fs.nb;
                sink.pos = fs.pos;
                sink.put(em_hi, em_sh);
                rg = rpt << fp.s;  // the dropped low s bits never influence range / total (the next use)
                done = i * PER;    // anything left runs through the scalar, capacity-checked loop below
            } else {
                auto code = [&](const Entries& en) {
#pragma unroll
                    for (int b = 0; b < SPW; b++) generic_symbol(en.e[b]);
                };
#pragma unroll 1
                for (; i < nvec; i++) {
                    prefetch_block(i);
                    const uint4 nxt = (i + 1 < nvec) ? ldg_stream_v4(v + i + 1) : make_uint4(0, 0, 0, 0);
                    Entries eB = lookup(cur.y);
                    code(eA);
                    eA = lookup(cur.z);
                    code(eB);
                    eB = lookup(cur.w);
                    code(eA);
                    eA = lookup(nxt.x);
                    code(eB);
                    cur = nxt;
                }
                done = nvec * PER;
            }
        }
    }
#pragma unroll 1
    for (uint64_t i = done; i < cnt; i++) {
        uint32_t s = (uint32_t)src[i];
        if (RANGECHK && s >= K) {
            if (!err) err = ST_SYMBOL_RANGE;
            s = 0;
        }
        generic_symbol(tab[s]);
    }

    uint32_t len = sink.finish(lo);  // src/encoder.rs:40-46
    if (!err && sink.overflowed()) err = ST_OUT_CAPACITY;
    a.lens[chunk] = len;
    a.status[chunk] = err;
}

}  // namespace rcb
