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
//   - the table entries of the next 32-bit word of symbols are fetched from
//     shared memory while the current word is coded (ping-pong eA / eB);
//   - byte emission is two funnel shifts plus a predicated 32-bit store, issued
//     one symbol late so it overlaps the next symbol's multiply chain;
//   - FUSED (consistent table): the lane carries rpt = range / total and the
//     renormalisation shift folds into the division (rcb_core.cuh: fused_step);
//     the rare events that need the reference's literal loops (loop 2, n1 >= 3)
//     are handled per 32-bit word: the word is coded speculatively without any
//     branch, and if a lane saw such an event it restores its checkpoint and
//     re-codes the word on the exact out-of-line path (stores are idempotent).
//
// Table placement (TABLE):
//   TAB_SHARED   one {cum, c} table for all chunks, in shared memory
//   TAB_LANE     one table per chunk, each lane's cum[K+1] in its own shared-memory row
//                (adaptive per-chunk models; needs regular tables and (K+1)*4 bytes per lane)
//   TAB_GLOBAL   one table per chunk, read through L1/L2 (any table, any K)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cuda.h>  // CUtensorMap (type only; the driver entry point is resolved at run time)

#include <type_traits>

#include "rcb_core.cuh"

namespace rcb {

// ---------------------------------------------------------------------------
// TMA input (encode_tma_kernel): the symbols are a 2-D tensor [n_chunks][chunk_bytes] of bytes and a
// warp's 32 lanes consume vector i of their 32 consecutive rows at the same time, so ONE
// cp.async.bulk.tensor.2d per warp fetches a {64 bytes x 32 rows} box -- four vectors for every lane -- into
// a 2 KiB stage, completion on an mbarrier; TMA_STAGES stages per warp.  The box lands with the 64-byte
// swizzle (16-byte chunk index ^= row/2 mod 4), which makes the lanes' 16-byte reads conflict-free.
// ---------------------------------------------------------------------------
constexpr uint32_t TMA_STAGES = 4;
constexpr uint32_t TMA_BOX_BYTES = 64;
constexpr uint32_t TMA_STAGE_BYTES = TMA_BOX_BYTES * 32;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}"
        :
        : "r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, uint32_t x, uint32_t y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(tmap), "r"(bar), "r"(x), "r"(y)
        : "memory");
}

// The hot loop's step can take its shift from one count-leading-zeros + the exact loop-2 test (rcb_core.cuh:
// TPUT) instead of the compare/select tree: 3 instructions fewer per symbol, but the find sits on the
// (lower, range) recurrence and the encoder runs one warp per scheduler at configs[1] -- measured 3.35 -> 3.53 ms.
// Off; the decoder's fused loop uses it (its chain hides it behind the table load).
#ifndef RCB_ENC_CLZ
#define RCB_ENC_CLZ 0
#endif
constexpr bool ENC_CLZ = RCB_ENC_CLZ != 0;

enum : int { TAB_SHARED = 0, TAB_LANE = 1, TAB_GLOBAL = 2 };
// FM_GENCS: general total with the divide-free step (each table entry carries cs = floor(c * 2^64 / total),
// rcb_core.cuh: fused_step_cs); TAB_SHARED only.  FM_GEN keeps the multiply-high reciprocal.
// FM_GENM2: general total >= 2^25 with the table-wide reciprocal (fused_step_m2); plain {cum, c} table.
enum : int { FM_GENERIC = -1, FM_BIG = FUSE_BIG, FM_POW2 = FUSE_POW2, FM_GEN = FUSE_GEN, FM_LANE = 3, FM_GENCS = 4,
             FM_GENM2 = 5 };

struct EncodeArgs {
    const void* syms;
    uint64_t n_syms;
    uint64_t chunk_syms;
    uint64_t n_chunks;
    const uint2* tabs;      // [n_models][K]
    const ModelHdr* hdrs;   // [n_models]
    const uint2* tab_cs;    // [K] floor(c * 2^64 / total) as {lo, hi} (shared model, FM_GENCS)
    uint32_t K;
    uint32_t lanes_per_block;  // chunks per block (<= blockDim.x)
    uint8_t* staging;       // [n_chunks][pitch]
    uint64_t pitch;
    uint32_t* lens;         // [n_chunks]
    uint32_t* status;       // [n_chunks]
    // restart points (rcb_core.cuh: Restart): record r of a chunk = state in front of symbol (r+1)*restart_syms
    Restart* restart;       // [n_chunks][restart_per_chunk] or nullptr
    uint64_t restart_syms;  // multiple of 64
    uint32_t restart_per_chunk;  // ceil(chunk_syms / restart_syms) - 1
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

// Symbol input of the fused loops: a private ring of 16-byte pieces per lane in shared memory, filled
// by cp.async (LDGSTS) several vectors ahead.  No register is the destination of a global load, so
// nothing on the symbol path -- and no CALL into the exact re-code path -- ever waits on the memory
// system (the register scoreboard is per warp; a load in flight stalls the warp at every call
// boundary).  Each piece carries a 256-byte L2 prefetch hint: no separate prefetch instruction.
constexpr uint32_t ENC_RING_PIECES = 8;
constexpr uint32_t ENC_RING_STRIDE = ENC_RING_PIECES * 16 + 16;  // 144: 16-byte aligned, spreads banks
constexpr uint32_t ENC_RING_AHEAD = 6;                           // pieces requested ahead of the one in use

__device__ __forceinline__ void enc_ring_issue(bool p, uint32_t saddr, const void* g) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q cp.async.cg.shared.global.L2::256B [%0], [%1], 16;\n\t}"
        :
        : "r"(saddr), "l"(g), "r"((uint32_t)p)
        : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t saddr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
    return r;
}

template <int SPW>
struct EncEntries {
    uint2 e[SPW];  // {cum, c}
};
template <int SPW>
struct EncEntries4 {
    uint4 e[SPW];  // {cum, c, cs lo, cs hi}
};
__device__ __forceinline__ uint4 lds_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return r;
}

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
template <int SPW, int MODE>
__device__ __noinline__ EncWordState enc_word_exact(EncWordState s, EncEntries<SPW> en, FusedParams fp,
                                                    uint8_t* row, uint32_t cap) {
    RowStore rs{row};
    EncSink<RowStore, false> sink(rs, cap);  // the caller checked room for a whole vector (see the guard)
    sink.pend = s.pend;
    sink.nb = s.nb;
    sink.pos = s.pos;
#pragma unroll
    for (int b = 0; b < SPW; b++) {
        uint64_t nlo, rgp, nrpt;
        uint32_t sh;
        const bool ok = fused_step<MODE>(s.lo, s.rpt, en.e[b].x, en.e[b].y, fp, nlo, rgp, nrpt, sh);
        sink.put(s.em_hi, s.em_sh);
        // fast results unconditionally; only the symbol that needs the literal loops branches
        s.em_hi = hi32(nlo);
        s.em_sh = sh;
        s.lo = nlo << sh;
        s.rpt = nrpt;
        if (RCB_UNLIKELY(!ok)) {
            s.em_sh = 0;
            uint64_t lo = nlo, rg = rgp;
            renorm_slow<false>(lo, rg, sink, s.err);
            s.lo = lo;
            s.rpt = fused_rpt<MODE>(rg, fp);
        }
    }
    s.pend = sink.pend;
    s.nb = sink.nb;
    s.pos = sink.pos;
    return s;
}

template <typename SYM, int TABLE, int FMODE, bool CHECKED, bool RANGECHK, bool TMA_IN>
__device__ __forceinline__ void encode_body(const EncodeArgs& a, const CUtensorMap* tmap) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ ModelHdr s_hdr;
    static_assert(!TMA_IN || (TABLE == TAB_SHARED && FMODE != FM_GENERIC), "TMA input: shared table, fused loop");
    const uint32_t K = a.K;
    const uint32_t L = a.lanes_per_block;
    const uint64_t block_first = (uint64_t)blockIdx.x * L;
    // shared layout: table (TAB_SHARED: uint2[K]; TAB_LANE: u32[L][K+1]) | input rings[blockDim.x] (fused loops)
    constexpr bool CS = FMODE == FM_GENCS;
    constexpr bool M2 = FMODE == FM_GENM2;
    static_assert(!CS || TABLE == TAB_SHARED, "FM_GENCS needs the shared table");
    const uint32_t ring_off = TABLE == TAB_SHARED ? ((K * (CS ? 16u : 8u) + 15u) & ~15u)
                              : TABLE == TAB_LANE ? ((L * (K + 1u) * 4u + 15u) & ~15u)
                                                  : 0u;
    if (TABLE == TAB_SHARED) {
        if (CS) {
            uint4* s_tab4 = reinterpret_cast<uint4*>(s_raw);
            for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) {
                const uint2 t = a.tabs[i], r = a.tab_cs[i];
                s_tab4[i] = make_uint4(t.x, t.y, r.x, r.y);
            }
        } else {
            uint2* s_tab = reinterpret_cast<uint2*>(s_raw);
            for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) s_tab[i] = a.tabs[i];
        }
        if (threadIdx.x == 0) s_hdr = a.hdrs[0];
        if (TMA_IN) {  // one "full" barrier per stage per warp
            const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_raw);
            const uint32_t stages = (sbase + ring_off + 1023u) & ~1023u;
            const uint32_t bars = stages + (blockDim.x >> 5) * TMA_STAGES * TMA_STAGE_BYTES;
            if (threadIdx.x < (blockDim.x >> 5) * TMA_STAGES) mbar_init(bars + threadIdx.x * 8u, 1u);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // inits visible to the async proxy
        }
        __syncthreads();
    }
    if (TABLE == TAB_LANE) {
        // each lane's cum[0..K] (cum[K] = total) in its own row of K+1 words (odd pitch for K = 256:
        // lanes reading the same symbol hit different banks); filled cooperatively, coalesced
        uint32_t* rows = reinterpret_cast<uint32_t*>(s_raw);
        const uint64_t left = a.n_chunks - block_first;
        const uint32_t lanes = left < L ? (uint32_t)left : L;
        const uint32_t pitch = K + 1;
        for (uint32_t l = 0; l < lanes; l++) {
            const uint2* t = a.tabs + (block_first + l) * K;
            for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) rows[l * pitch + i] = t[i].x;
            if (threadIdx.x == 0) rows[l * pitch + K] = a.hdrs[block_first + l].div.total;
        }
        __syncthreads();
    }
    if (threadIdx.x >= L) return;
    const uint64_t chunk = block_first + threadIdx.x;
    if (chunk >= a.n_chunks) return;
    const uint64_t first = chunk * a.chunk_syms;
    const uint64_t cnt = (a.n_syms - first < a.chunk_syms) ? (a.n_syms - first) : a.chunk_syms;
    const SYM* src = reinterpret_cast<const SYM*>(a.syms) + first;

    const uint2* tab = TABLE == TAB_SHARED ? reinterpret_cast<const uint2*>(s_raw) : a.tabs + chunk * K;
    const uint4* tab4 = reinterpret_cast<const uint4*>(s_raw);  // FM_GENCS
    const uint32_t* row = reinterpret_cast<const uint32_t*>(s_raw) + (size_t)threadIdx.x * (K + 1);
    DivParams div;
    bool pow2;
    if (TABLE == TAB_SHARED) {
        div = s_hdr.div;
        pow2 = (s_hdr.flags & MODEL_POW2) != 0;
    } else {
        ModelHdr h = a.hdrs[chunk];
        div = h.div;
        pow2 = (h.flags & MODEL_POW2) != 0;
    }
    const uint32_t cap = (uint32_t)a.pitch;

    uint64_t lo = 0, rg = ~0ull;  // src/range_coder.rs:13-20
    uint32_t err = 0;
    RowStore rs{a.staging + chunk * a.pitch};
    EncSink<RowStore, true> sink(rs, cap);
    // restart points of this chunk: rp_n written so far, the next one is due in front of symbol (rp_n+1)*restart_syms
    Restart* const rpts = a.restart ? a.restart + chunk * a.restart_per_chunk : nullptr;
    uint32_t rp_n = 0;
    // a record holds range rounded down to a multiple of total_freq (what the fused loops carry; canonical, so
    // that every kernel flavour -- and the checker -- writes the same bytes)
    auto canon_range = [&](uint64_t r) -> uint64_t {
        return (pow2 ? range_par_total<true>(r, div) : range_par_total<false>(r, div)) * (uint64_t)div.total;
    };

    constexpr int SPW = 4 / sizeof(SYM);  // symbols per 32-bit word
    using Entries = typename std::conditional<CS, EncEntries4<SPW>, EncEntries<SPW>>::type;
    auto entry = [&](uint32_t s) -> uint2 {
        if (RANGECHK && s >= K) {
            if (!err) err = ST_SYMBOL_RANGE;
            s = 0;
        }
        if (TABLE == TAB_LANE) {
            const uint32_t cum = row[s];
            return make_uint2(cum, row[s + 1] - cum);  // regular table: c = cum[s+1] - cum[s]
        }
        if (CS) {
            const uint4 t = lds_u4(tab4 + s);
            return make_uint2(t.x, t.y);
        }
        return tab[s];
    };
    auto lookup = [&](uint32_t w) -> Entries {
        Entries r;
#pragma unroll
        for (int b = 0; b < SPW; b++) {
            uint32_t sy = sizeof(SYM) == 1 ? ((w >> (8 * b)) & 0xFFu) : ((w >> (16 * b)) & 0xFFFFu);
            if constexpr (CS) {
                if (RANGECHK && sy >= K) {
                    if (!err) err = ST_SYMBOL_RANGE;
                    sy = 0;
                }
                r.e[b] = lds_u4(tab4 + sy);
            } else {
                r.e[b] = entry(sy);
            }
        }
        return r;
    };
    auto generic_symbol = [&](uint2 e) {
        if (pow2)
            update_symbol<true, CHECKED>(lo, rg, e.x, e.y, div, sink, err);
        else
            update_symbol<false, CHECKED>(lo, rg, e.x, e.y, div, sink, err);
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
            uint64_t i = 0;
            if constexpr (FMODE != FM_GENERIC) {
                // ring prologue: pieces 0 .. AHEAD-1, one commit group each
                const uint32_t ring =
                    (uint32_t)__cvta_generic_to_shared(s_raw + ring_off + (size_t)threadIdx.x * ENC_RING_STRIDE);
                // TMA input: this warp's stages and barriers, the lane's swizzled slot inside a stage
                const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
                const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_raw);
                const uint32_t stages0 = (sbase + ring_off + 1023u) & ~1023u;
                const uint32_t my_stages = stages0 + warp * TMA_STAGES * TMA_STAGE_BYTES;
                const uint32_t my_bars = stages0 + (blockDim.x >> 5) * TMA_STAGES * TMA_STAGE_BYTES + warp * TMA_STAGES * 8u;
                const uint32_t row0 = (uint32_t)(chunk - lane);              // first chunk (tensor row) of this warp
                const uint32_t ncol = (uint32_t)((nvec + 3) >> 2);           // 64-byte columns of the row
                const uint32_t lane_off = lane * TMA_BOX_BYTES, lane_swz = (lane >> 1) & 3u;
                const unsigned wmask = TMA_IN ? __activemask() : 0u;
                uint32_t issued = 0;  // columns requested so far (lane 0 issues; every lane counts)
                auto tma_issue = [&](uint32_t col) {
                    const uint32_t st = col & (TMA_STAGES - 1);
                    if (lane == 0) {
                        mbar_expect_tx(my_bars + st * 8u, TMA_STAGE_BYTES);
                        tma_load_2d(my_stages + st * TMA_STAGE_BYTES, tmap, my_bars + st * 8u, col * TMA_BOX_BYTES, row0);
                    }
                };
                auto tma_wait = [&](uint32_t col) {
                    mbar_wait(my_bars + (col & (TMA_STAGES - 1)) * 8u, (col / TMA_STAGES) & 1u);
                };
                auto tma_vec = [&](uint64_t vi) -> uint4 {  // vector vi of this lane's row
                    const uint32_t col = (uint32_t)(vi >> 2), k = (uint32_t)vi & 3u;
                    return lds_v4(my_stages + (col & (TMA_STAGES - 1)) * TMA_STAGE_BYTES + lane_off + ((k ^ lane_swz) << 4));
                };
                uint4 cur;
                if constexpr (TMA_IN) {
                    for (; issued < TMA_STAGES && issued < ncol; issued++) tma_issue(issued);
                    tma_wait(0);
                    cur = tma_vec(0);
                } else {
#pragma unroll
                    for (uint32_t q = 0; q < ENC_RING_AHEAD; q++) {
                        enc_ring_issue(q < nvec, ring + q * 16, v + q);
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    }
                    asm volatile("cp.async.wait_group %0;" ::"n"(ENC_RING_AHEAD - 1) : "memory");
                    cur = lds_v4(ring);
                }
                Entries eA = lookup(cur.x);
                // fast sink: no capacity test per store; room for a whole vector is checked once
                // per vector (16 symbols x at most 15 bytes + the deferred emission < 320 bytes)
                EncSink<RowStore, false> fs(rs, cap);
                const FusedParams fp = make_fused(div);
                const Recip2 k2 = M2 ? make_recip2(div.total) : Recip2{0ull, 0u};
                uint32_t em_hi = 0, em_sh = 0;  // previous symbol's bytes, emitted one symbol late
                uint64_t rpt;
                // one flavour of the word coder per division mode (FM_LANE picks per lane at run time)
                auto run = [&](auto mode_tag) {
                    constexpr int MODE = decltype(mode_tag)::value;
                    rpt = fused_rpt<MODE>(rg, fp);
                    auto code = [&](const Entries& en) {
                        const EncWordState chk{lo, rpt, fs.pend, fs.nb, fs.pos, em_hi, em_sh, err};
                        bool bad = false;
#pragma unroll
                        for (int b = 0; b < SPW; b++) {  // speculative: straight-line, no branch
                            uint64_t nlo, rgp, nrpt;
                            uint32_t sh;
                            bool ok;
                            if constexpr (CS)
                                ok = fused_step_cs<ENC_CLZ>(lo, rpt, en.e[b].x, en.e[b].y,
                                                            ((uint64_t)en.e[b].w << 32) | en.e[b].z, nlo, rgp, nrpt, sh);
                            else if constexpr (M2)
                                ok = fused_step_m2<ENC_CLZ>(lo, rpt, en.e[b].x, en.e[b].y, k2, nlo, rgp, nrpt, sh);
                            else
                                ok = fused_step<MODE, ENC_CLZ>(lo, rpt, en.e[b].x, en.e[b].y, fp, nlo, rgp, nrpt, sh);
                            fs.put(em_hi, em_sh);
                            em_hi = hi32(nlo);
                            em_sh = sh;
                            lo = nlo << sh;
                            rpt = nrpt;
                            bad |= !ok;
                        }
                        if (RCB_UNLIKELY(bad)) {  // restore the checkpoint and re-code the word exactly
                            EncEntries<SPW> en2;
#pragma unroll
                            for (int b = 0; b < SPW; b++) en2.e[b] = make_uint2(en.e[b].x, en.e[b].y);
                            const EncWordState r = enc_word_exact<SPW, MODE>(chk, en2, fp, rs.row, cap);
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
                    // vectors up to the next restart point (or the end); the hot loop itself is unchanged
                    const uint64_t rvec = rpts ? a.restart_syms / PER : nvec;
                    uint64_t lim = rvec < nvec ? rvec : nvec;
                    for (;;) {
                    bool full = false;
#pragma unroll 1
                    for (; i < lim; i++) {
                        uint4 nxt;
                        if constexpr (TMA_IN) {
                            // the warp leaves together (lane 0 issues the loads the others wait for)
                            if (__any_sync(wmask, fs.pos + 320u > cap)) {
                                full = true;
                                break;
                            }
                            const uint32_t k = (uint32_t)i & 3u, col = (uint32_t)(i >> 2);
                            if (k == 0 && col >= 1) {
                                // every vector of column col-1 has been consumed (the last one was `cur` of the
                                // previous trip): its stage takes column col-1+STAGES
                                __syncwarp(wmask);
                                if (issued < ncol) tma_issue(issued);
                                issued += issued < ncol ? 1u : 0u;
                            }
                            if (k == 3 && i + 1 < nvec) tma_wait(col + 1);  // vector i+1 opens the next column
                            nxt = i + 1 < nvec ? tma_vec(i + 1) : make_uint4(0u, 0u, 0u, 0u);
                        } else {
                        if (fs.pos + 320u > cap) {  // finish this chunk on the capacity-checked path
                            full = true;
                            break;
                        }
                        // request piece i+AHEAD (its slot held piece i-2), retire all but the newest AHEAD-1
                        // groups: pieces <= i+1 have landed
                        const uint64_t q = i + ENC_RING_AHEAD;
                        enc_ring_issue(q < nvec, ring + ((uint32_t)q & (ENC_RING_PIECES - 1)) * 16, v + q);
                        asm volatile("cp.async.commit_group;" ::: "memory");
                        asm volatile("cp.async.wait_group %0;" ::"n"(ENC_RING_AHEAD - 1) : "memory");
                        nxt = lds_v4(ring + (((uint32_t)i + 1u) & (ENC_RING_PIECES - 1)) * 16);
                        }
                        Entries eB = lookup(cur.y);
                        code(eA);
                        eA = lookup(cur.z);
                        code(eB);
                        eB = lookup(cur.w);
                        code(eA);
                        eA = lookup(i + 1 < nvec ? nxt.x : 0u);  // word 0 of the next vector (unused past the end)
                        code(eB);
                        cur = nxt;
                    }
                    if (full || i >= nvec) break;
                    // restart point in front of vector i: lower as it stands, range as rpt * total, and the
                    // bytes emitted so far = stored + pending in the sink + the previous symbol's deferred ones
                    rpts[rp_n++] = Restart{lo, MODE == FUSE_GEN ? rpt * (uint64_t)div.total : rpt << fp.s,
                                           fs.pos + (fs.nb >> 3) + (em_sh >> 3), 0u};
                    lim = nvec - i > rvec ? i + rvec : nvec;
                    }
                    // back to the generic (lower, range) form: range = rpt * total keeps range / total
                    // == rpt, the only way `range` is used before the next update
                    rg = MODE == FUSE_GEN ? rpt * (uint64_t)div.total : rpt << fp.s;
                };
                if constexpr (FMODE == FM_LANE) {
                    if (pow2) run(std::integral_constant<int, FUSE_POW2>{});
                    else run(std::integral_constant<int, FUSE_GEN>{});
                } else if constexpr (CS || M2) {
                    run(std::integral_constant<int, FUSE_GEN>{});
                } else {
                    run(std::integral_constant<int, FMODE>{});
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                if constexpr (TMA_IN) {
                    // loads still in flight (an early exit): they must land before the block's shared memory
                    // can go to another block.
                    uint32_t seen = (uint32_t)(i >> 2) + 1u;  // columns 0 .. i/4 were waited for
                    seen = seen < ncol ? seen : ncol;
                    for (uint32_t cidx = seen; cidx < issued; cidx++) tma_wait(cidx);
                }
                sink.pend = fs.pend;
                sink.nb = fs.nb;
                sink.pos = fs.pos;
                sink.put(em_hi, em_sh);
                done = i * PER;  // anything left runs through the scalar, capacity-checked loop below
            } else {
                prefetch_l2_bulk(v, (uint32_t)(nvec < PF_VECS ? nvec * 16 : PF_VECS * 16));
                uint4 cur = ldg_stream_v4(v);
                Entries eA = lookup(cur.x);
                auto code = [&](const Entries& en) {
#pragma unroll
                    for (int b = 0; b < SPW; b++) generic_symbol(en.e[b]);
                };
#pragma unroll 1
                for (; i < nvec; i++) {
                    prefetch_block(i);
                    if (rpts && i * PER == (uint64_t)(rp_n + 1u) * a.restart_syms)
                        rpts[rp_n++] = Restart{lo, canon_range(rg), sink.pos + (sink.nb >> 3), 0u};
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
        if (rpts && i == (uint64_t)(rp_n + 1u) * a.restart_syms)
            rpts[rp_n++] = Restart{lo, canon_range(rg), sink.pos + (sink.nb >> 3), 0u};
        generic_symbol(entry((uint32_t)src[i]));
    }
    if (rpts)  // a ragged last chunk has fewer restart points: the rest are marked absent (range 0)
        for (; rp_n < a.restart_per_chunk; rp_n++) rpts[rp_n] = Restart{0ull, 0ull, 0u, 0u};

    uint32_t len = sink.finish(lo);  // src/encoder.rs:40-46
    if (!err && sink.overflowed()) err = ST_OUT_CAPACITY;
    a.lens[chunk] = len;
    a.status[chunk] = err;
}

template <typename SYM, int TABLE, int FMODE, bool CHECKED, bool RANGECHK>
__global__ void __launch_bounds__(512, 1) encode_kernel(EncodeArgs a) {
    encode_body<SYM, TABLE, FMODE, CHECKED, RANGECHK, false>(a, nullptr);
}

// Same coder, symbols staged by TMA (shared table, whole chunks only: the tensor has n_chunks full rows).
template <typename SYM, int FMODE, bool RANGECHK>
__global__ void __launch_bounds__(512, 1) encode_tma_kernel(EncodeArgs a, const __grid_constant__ CUtensorMap tmap) {
    encode_body<SYM, TAB_SHARED, FMODE, false, RANGECHK, true>(a, &tmap);
}

}  // namespace rcb
