// rcb_decode.cuh -- K5: chunk-parallel decode (src/decoder.rs:14-54 with
// examples/sample_impl.rs:27-45 as the built-in find_index), one independent
// reference `Decoder` per lane, same chunk layout as the encoder.
//
// Shared static model: the symbol comes from a bucketed cum_freq -> symbol
// table in shared memory (one 16-byte entry per bucket holds two candidate
// symbols and their cumulative bounds), verified exactly; the exact search over
// cum_freq is the out-of-line fallback.  Per-chunk models use the exact search
// on their own table.  Like the encoder, run time is one lane's per-symbol
// instruction stream, so the hot step is branch-free: funnel-shift window
// refill with a predicated load three words ahead, select trees instead of
// find-leading-one, (FUSED) range/total folded into the renormalisation shift,
// and rare events handled per 32-bit word of output by checkpoint + exact
// re-decode (see rcb_encode.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcb_core.cuh"

namespace rcb {

struct DecodeArgs {
    const uint8_t* stream;
    const uint64_t* offsets;  // [n_chunks+1]
    uint64_t n_syms;
    uint64_t chunk_syms;
    uint64_t n_chunks;
    const uint2* tabs;
    const ModelHdr* hdrs;
    const LutEntry* lut;      // shared model only
    uint32_t K;
    uint32_t per_chunk;
    void* out;
    uint32_t* status;
};

// Sequential reader of one chunk's bytes: aligned 32-bit loads, three words of
// lookahead in registers (the load issued at one refill is consumed three
// refills later, ~15 symbols, so its latency never reaches the coder).  Loads
// are clamped to the last word of the stream so a corrupt stream cannot make a
// lane read outside the caller's buffer.
struct GlobalFetch {
    const uint32_t* base;  // word that holds the chunk's first byte
    uint32_t idx;          // words consumed so far; w0 is word idx
    uint32_t last;         // last readable word (from base)
    uint32_t w0, w1, w2;
    __device__ __forceinline__ uint32_t load(uint32_t i) const { return __ldg(base + (i < last ? i : last)); }
    __device__ __forceinline__ void init(const uint8_t* stream, uint64_t off, uint64_t total_bytes) {
        const uint64_t word0 = off >> 2;
        base = reinterpret_cast<const uint32_t*>(stream) + word0;
        // the stream is readable up to the next multiple of 16 bytes (rcb200.h)
        const uint64_t lastw = (((total_bytes + 15) >> 4) << 2) - 1 - word0;
        last = lastw > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)lastw;
        idx = 0;
        w0 = load(0);
        w1 = load(1);
        w2 = load(2);
    }
    __device__ __forceinline__ uint32_t peek_be32() const { return bswap32(w0); }
    __device__ __forceinline__ void advance_if(bool p) {
        idx += p ? 1u : 0u;
        w0 = p ? w1 : w0;
        w1 = p ? w2 : w1;
        const uint32_t j = idx + 2u;
        const uint32_t* q = base + (j < last ? j : last);
        asm volatile(
            "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q ld.global.nc.b32 %0, [%2];\n\t}"
            : "+r"(w2)
            : "r"((uint32_t)p), "l"(q)
            : "memory");
    }
    __device__ __forceinline__ uint32_t words_fetched() const { return idx; }
};

__device__ __forceinline__ void prefetch_l2_bulk_dec(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ LutEntry lds_lut(uint32_t saddr) {
    LutEntry e;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(e.cumA), "=r"(e.cumB), "=r"(e.cumC), "=r"(e.syms)
                 : "r"(saddr));
    return e;
}

// Lane state that travels by value through the exact (out-of-line) paths.
struct DecLaneState {
    uint64_t lo, rg;
    uint32_t dh, dl, wh, wl, cnt;
    const uint32_t* base;
    uint32_t idx, last, w0, w1, w2;
    uint32_t err, syms;  // syms: symbols decoded by the call, packed like the output word
};

// One symbol on the exact path: the reference's search in the product domain plus the
// generic renormalisation (literal loops when needed).
template <bool CHECKED>
__device__ __forceinline__ uint32_t dec_symbol_exact(uint64_t& lo, uint64_t& rg, DecSink<GlobalFetch>& sink,
                                                     uint32_t& err, const uint2* tab, uint32_t K,
                                                     const DivParams& div, bool pow2) {
    const uint64_t rpt = pow2 ? range_par_total<true>(rg, div) : range_par_total<false>(rg, div);
    const uint64_t d = sink.data() - lo;  // examples/sample_impl.rs:29
    const uint32_t sym = find_index_exact(d, rpt, K, [&](uint32_t i) { return tab[i].x; });
    const uint2 e = tab[sym];
    uint64_t nlo = lo + rpt * (uint64_t)e.x;  // src/decoder.rs:42-50
    uint64_t rgn = rpt * (uint64_t)e.y;
    if (CHECKED && nlo < lo) {
        if (!err) err = ST_LOWER_OVERFLOW;
        nlo = 0;
        rgn = ~0ull;
    }
    lo = nlo;
    rg = rgn;
    renorm<CHECKED>(lo, rg, sink, err);  // consumes the same number of bytes (:52)
    return sym;
}

// n_syms symbols (<= 4) decoded exactly from a checkpoint; symbols packed sym_bits apart.
// lut_saddr != 0 (FUSED callers): each symbol first tries the table; a verified symbol whose
// renormalisation needs the literal loops (loop 2, the common reason to be here) skips the search.
template <bool CHECKED>
__device__ __noinline__ DecLaneState dec_exact(DecLaneState s, const uint2* tab, uint32_t K, DivParams div,
                                               uint32_t pow2, uint32_t n_syms, uint32_t sym_bits,
                                               uint32_t lut_saddr, float lut_scale) {
    GlobalFetch gf;
    gf.base = s.base;
    gf.idx = s.idx;
    gf.last = s.last;
    gf.w0 = s.w0;
    gf.w1 = s.w1;
    gf.w2 = s.w2;
    DecSink<GlobalFetch> sink(gf);
    sink.dh = s.dh;
    sink.dl = s.dl;
    sink.wh = s.wh;
    sink.wl = s.wl;
    sink.cnt = s.cnt;
    const FusedParams fp{div.shift, div.shift - 24u, 1u << (48u - div.shift)};
    uint32_t acc = 0;
#pragma unroll 1
    for (uint32_t b = 0; b < n_syms; b++) {
        uint32_t sym;
        bool resolved = false;
        if (lut_saddr) {
            const uint64_t rpt = s.rg >> fp.s;
            const uint32_t off = lut_offset16(sink.dh - hi32(s.lo), lut_rinv16(hi32(s.rg), lut_scale));
            const FusedDec r = fused_decode_step(s.lo, rpt, sink.data(), lds_lut(lut_saddr + off), fp);
            if (r.inside) {
                resolved = true;
                sym = r.sym;
                s.lo = r.nlo;
                s.rg = r.rgp;
                renorm<CHECKED>(s.lo, s.rg, sink, s.err);
            }
        }
        if (!resolved) sym = dec_symbol_exact<CHECKED>(s.lo, s.rg, sink, s.err, tab, K, div, pow2 != 0);
        acc |= sym << (sym_bits * b);
    }
    s.syms = acc;
    s.dh = sink.dh;
    s.dl = sink.dl;
    s.wh = sink.wh;
    s.wl = sink.wl;
    s.cnt = sink.cnt;
    s.idx = sink.f.idx;
    s.w0 = sink.f.w0;
    s.w1 = sink.f.w1;
    s.w2 = sink.f.w2;
    return s;
}

template <typename SYM, bool SHARED, bool POW2, bool CHECKED, bool FUSED>
__global__ void __launch_bounds__(256) decode_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ ModelHdr s_hdr;
    // shared layout: LutEntry[nb or 4096] | uint2[K]
    LutEntry* s_lut = reinterpret_cast<LutEntry*>(s_raw);
    uint2* s_tab = nullptr;
    if (SHARED) {
        if (threadIdx.x == 0) s_hdr = a.hdrs[0];
        __syncthreads();
        const uint32_t nb = (s_hdr.flags & MODEL_REGULAR) ? s_hdr.nb : 0u;
        const uint32_t nb_pad = FUSED ? 4096u : nb;  // FUSED indexes any of 4096 entries
        s_tab = reinterpret_cast<uint2*>(s_raw + (size_t)nb_pad * sizeof(LutEntry));
        const uint4* gl = reinterpret_cast<const uint4*>(a.lut);
        uint4* sl = reinterpret_cast<uint4*>(s_lut);
        const uint32_t total = s_hdr.div.total;
        for (uint32_t i = threadIdx.x; i < nb_pad; i += blockDim.x)
            sl[i] = i < nb ? gl[i] : make_uint4(total, total, total, 0u);  // empty interval: never verifies
        for (uint32_t i = threadIdx.x; i < a.K; i += blockDim.x) s_tab[i] = a.tabs[i];
        __syncthreads();
    }
    const uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk >= a.n_chunks) return;
    const uint64_t first = chunk * a.chunk_syms;
    const uint64_t cnt = (a.n_syms - first < a.chunk_syms) ? (a.n_syms - first) : a.chunk_syms;
    SYM* dst = reinterpret_cast<SYM*>(a.out) + first;

    const uint2* tab = SHARED ? s_tab : a.tabs + chunk * a.K;
    ModelHdr hdr = SHARED ? s_hdr : a.hdrs[chunk];
    const DivParams div = hdr.div;
    const bool pow2 = SHARED ? POW2 : ((hdr.flags & MODEL_POW2) != 0);
    const bool use_lut = SHARED && (hdr.flags & MODEL_REGULAR);
    const float lut_scale = hdr.lut_scale;
    const float max_bucket = (float)(hdr.nb ? hdr.nb - 1 : 0);
    const uint32_t K = a.K;

    const uint64_t off0 = a.offsets[chunk], off1 = a.offsets[chunk + 1];
    const uint64_t total_bytes = a.offsets[a.n_chunks];
    const uint32_t skip = (uint32_t)(off0 & 3u);
    GlobalFetch gf;
    gf.init(a.stream, off0, total_bytes);
    DecSink<GlobalFetch> sink(gf);
    // L2 prefetch of this lane's code bytes in 1 KiB granules, two granules ahead
    constexpr uint32_t PF_WORDS = 256;
    const uint8_t* pf_base = a.stream + (off0 & ~15ull);
    const uint64_t pf_avail = ((total_bytes + 15) & ~15ull) - (off0 & ~15ull);
    uint32_t pf_next = 0;  // next granule (in words from base) to request
    auto prefetch_to = [&](uint32_t upto_words) {
        while (pf_next < upto_words) {
            const uint64_t o = (uint64_t)pf_next * 4;
            if (o < pf_avail) {
                const uint64_t left = pf_avail - o;
                prefetch_l2_bulk_dec(pf_base + o, (uint32_t)(left < PF_WORDS * 4 ? left : PF_WORDS * 4));
            }
            pf_next += PF_WORDS;
        }
    };
    prefetch_to(2 * PF_WORDS);
    sink.prime(skip);  // src/decoder.rs:14-23

    uint64_t lo = 0, rg = ~0ull;
    uint32_t err = 0;

    constexpr uint32_t PER = 4 / sizeof(SYM);  // symbols per 32-bit store
    constexpr uint32_t SYM_BITS = 8 * sizeof(SYM);
    auto snapshot = [&]() -> DecLaneState {
        return DecLaneState{lo, rg, sink.dh, sink.dl, sink.wh, sink.wl, sink.cnt, sink.f.base,
                            sink.f.idx, sink.f.last, sink.f.w0, sink.f.w1, sink.f.w2, err, 0u};
    };
    auto restore = [&](const DecLaneState& st) {
        lo = st.lo;
        rg = st.rg;
        sink.dh = st.dh;
        sink.dl = st.dl;
        sink.wh = st.wh;
        sink.wl = st.wl;
        sink.cnt = st.cnt;
        sink.f.idx = st.idx;
        sink.f.w0 = st.w0;
        sink.f.w1 = st.w1;
        sink.f.w2 = st.w2;
        err = st.err;
    };

    uint64_t done = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 3u) == 0;
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst);
    const uint64_t nw = aligned ? cnt / PER : 0;

    if constexpr (FUSED) {
        const FusedParams fp{div.shift, div.shift - 24u, 1u << (48u - div.shift)};
        uint64_t rpt = rg >> fp.s;
        float rinv16 = lut_rinv16(hi32(rg), lut_scale);
        const uint32_t lut_saddr = (uint32_t)__cvta_generic_to_shared(s_lut);
#pragma unroll 1
        for (uint64_t i = 0; i < nw; i++) {
            if (sink.f.idx + PF_WORDS >= pf_next) prefetch_to(sink.f.idx + 2 * PF_WORDS);
            rg = rpt << fp.s;  // checkpoint in the generic form (low s bits never matter)
            const DecLaneState chk = snapshot();
            uint32_t acc = 0;
            bool bad = false;
#pragma unroll
            for (uint32_t b = 0; b < PER; b++) {  // speculative: straight-line, no branch
                const uint64_t data = sink.data();
                const uint32_t off = lut_offset16(sink.dh - hi32(lo), rinv16);
                const LutEntry e = lds_lut(lut_saddr + off);
                const FusedDec r = fused_decode_step(lo, rpt, data, e, fp);
                sink.put(0u, r.sh);
                lo = r.nlo << r.sh;
                rpt = r.nrpt;
                rinv16 = lut_rinv16(hi32(r.rgp << r.sh), lut_scale);
                acc |= r.sym << (SYM_BITS * b);
                bad |= !r.ok;
            }
            if (RCB_UNLIKELY(bad)) {  // restore the checkpoint and decode the word exactly
                const DecLaneState r = dec_exact<CHECKED>(chk, tab, K, div, 1u, PER, SYM_BITS, lut_saddr, lut_scale);
                restore(r);
                acc = r.syms;
                rpt = rg >> fp.s;
                rinv16 = lut_rinv16(hi32(rg), lut_scale);
            }
            dw[i] = acc;
        }
        done = nw * PER;
        rg = rpt << fp.s;
    }

    auto step = [&]() -> uint32_t {
        if (use_lut) {
            const uint64_t rpt = pow2 ? range_par_total<true>(rg, div) : range_par_total<false>(rg, div);
            const uint64_t d = sink.data() - lo;  // examples/sample_impl.rs:29
            const uint32_t b = lut_bucket(d, rg, lut_scale, max_bucket);
            const LutEntry e = s_lut[b];
            uint32_t sym;
            uint64_t P, rgn;
            if (RCB_LIKELY(lut_resolve(e, d, rpt, sym, P, rgn))) {
                uint64_t nlo = lo + P;  // src/decoder.rs:42-50
                if (CHECKED && nlo < lo) {
                    if (!err) err = ST_LOWER_OVERFLOW;
                    nlo = 0;
                    rgn = ~0ull;
                }
                lo = nlo;
                rg = rgn;
                renorm<CHECKED>(lo, rg, sink, err);
                return sym;
            }
            // table miss (rare): exact search, out of line
            const DecLaneState r = dec_exact<CHECKED>(snapshot(), tab, K, div, pow2 ? 1u : 0u, 1u, 0u, 0u, 0.0f);
            restore(r);
            return r.syms;
        }
        return dec_symbol_exact<CHECKED>(lo, rg, sink, err, tab, K, div, pow2);
    };

    if (!FUSED) {
#pragma unroll 1
        for (uint64_t i = 0; i < nw; i++) {
            if (sink.f.idx + PF_WORDS >= pf_next) prefetch_to(sink.f.idx + 2 * PF_WORDS);
            uint32_t acc = 0;
#pragma unroll
            for (uint32_t b = 0; b < PER; b++) acc |= step() << (SYM_BITS * b);
            dw[i] = acc;
        }
        done = nw * PER;
    }
#pragma unroll 1
    for (uint64_t i = done; i < cnt; i++) dst[i] = (SYM)step();

    const uint32_t used = sink.used(sink.f.words_fetched(), skip);
    if (!err && (uint64_t)used > off1 - off0) err = ST_TRUNCATED;  // src/decoder.rs:33
    a.status[chunk] = err;
}

}  // namespace rcb
