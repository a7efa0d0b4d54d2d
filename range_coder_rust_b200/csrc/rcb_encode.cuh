// rcb_encode.cuh -- K3: chunk-parallel encode (src/encoder.rs:24-46 over
// src/range_coder.rs:53-135), one independent reference `Encoder` per lane.
//
// With 64 KiB chunks a 1 GiB batch is only 16384 lanes (512 warps on 592 warp
// schedulers), so run time = symbols per chunk x the latency of one lane's
// per-symbol dependency chain.  The loop below is therefore organised to keep
// everything except the (lower, range) recurrence off that chain:
//   - symbols arrive as 16-byte vectors, one vector of lookahead in registers;
//   - the {cum, c} entries of the next 32-bit word of symbols are fetched from
//     shared memory while the current word is coded (ping-pong eA / eB);
//   - byte emission is a funnel shift plus a predicated 32-bit store.
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
    __device__ __forceinline__ void word(uint32_t pos, uint32_t w) const {
        *reinterpret_cast<uint32_t*>(row + pos) = w;
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

// SHARED: one table for all chunks, staged in shared memory.
// !SHARED: one table per chunk, read through L1/L2 from global memory.
template <typename SYM, bool SHARED, bool POW2, bool CHECKED, bool RANGECHK>
__global__ void __launch_bounds__(256) encode_kernel(EncodeArgs a) {
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

    uint64_t lo = 0, rg = ~0ull;  // src/range_coder.rs:13-20
    uint32_t err = 0;
    RowStore rs{a.staging + chunk * a.pitch};
    EncSink<RowStore> sink(rs, (uint32_t)a.pitch);

    constexpr int SPW = 4 / sizeof(SYM);  // symbols per 32-bit word
    struct Entries {
        uint2 e[SPW];
    };
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
    auto code = [&](const Entries& en) {
#pragma unroll
        for (int b = 0; b < SPW; b++) {
            if (SHARED) {
                update_symbol<POW2, CHECKED>(lo, rg, en.e[b].x, en.e[b].y, div, sink, err);
            } else {
                if (pow2)
                    update_symbol<true, CHECKED>(lo, rg, en.e[b].x, en.e[b].y, div, sink, err);
                else
                    update_symbol<false, CHECKED>(lo, rg, en.e[b].x, en.e[b].y, div, sink, err);
            }
        }
    };

    constexpr uint32_t PER = 16 / sizeof(SYM);
    uint64_t done = 0;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const uint4* v = reinterpret_cast<const uint4*>(src);
        const uint64_t nvec = cnt / PER;
        if (nvec) {
            uint4 cur = ldg_stream_v4(v);
            Entries eA = lookup(cur.x);
#pragma unroll 1
            for (uint64_t i = 0; i < nvec; i++) {
                // one vector of lookahead hides the global-load latency behind 16 bytes of work
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
        }
        done = nvec * PER;
    }
#pragma unroll 1
    for (uint64_t i = done; i < cnt; i++) {
        uint32_t s = (uint32_t)src[i];
        if (RANGECHK && s >= K) {
            if (!err) err = ST_SYMBOL_RANGE;
            s = 0;
        }
        const uint2 e = tab[s];
        if (pow2)
            update_symbol<true, CHECKED>(lo, rg, e.x, e.y, div, sink, err);
        else
            update_symbol<false, CHECKED>(lo, rg, e.x, e.y, div, sink, err);
    }

    uint32_t len = sink.finish(lo);  // src/encoder.rs:40-46
    if (!err && sink.overflowed()) err = ST_OUT_CAPACITY;
    a.lens[chunk] = len;
    a.status[chunk] = err;
}

}  // namespace rcb
