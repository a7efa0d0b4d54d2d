// rcb_decode.cuh -- K5: chunk-parallel decode (src/decoder.rs:14-54 with
// examples/sample_impl.rs:27-45 as the built-in find_index), one independent
// reference `Decoder` per lane, same chunk layout as the encoder.
//
// Shared static model: the symbol comes from a bucketed cum_freq -> symbol
// table in shared memory (one 16-byte entry per bucket holds two candidate
// symbols and their cumulative bounds), verified in the product domain; the
// exact search over cum_freq is the fallback.  Per-chunk models use the exact
// search on their own table.
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

// Sequential reader of one chunk's bytes: aligned 32-bit loads with one word of
// lookahead in a register.  Loads are clamped to the last word of the stream so
// a corrupt stream cannot make a lane read outside the caller's buffer.
struct GlobalFetch {
    const uint32_t* base;  // word that holds the chunk's first byte
    uint32_t idx;          // index (from base) of the word held in nextw
    uint32_t last;         // last readable word (from base)
    uint32_t nextw;
    __device__ __forceinline__ void init(const uint8_t* stream, uint64_t off, uint64_t total_bytes) {
        const uint64_t w0 = off >> 2;
        base = reinterpret_cast<const uint32_t*>(stream) + w0;
        // one padded word past the end of the stream is readable (rcb200.h)
        const uint64_t lastw = ((total_bytes + 3) >> 2) - w0;
        last = lastw > 0xFFFFFFFEull ? 0xFFFFFFFEu : (uint32_t)lastw;
        idx = 0;
        nextw = __ldg(base);
    }
    __device__ __forceinline__ uint32_t next_be32() {
        uint32_t r = bswap32(nextw);
        idx++;
        nextw = __ldg(base + (idx < last ? idx : last));
        return r;
    }
    __device__ __forceinline__ uint32_t words_fetched() const { return idx; }
};

template <typename SYM, bool SHARED, bool POW2, bool CHECKED>
__global__ void __launch_bounds__(256) decode_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ ModelHdr s_hdr;
    // shared layout: LutEntry[nb] | uint2[K]
    LutEntry* s_lut = reinterpret_cast<LutEntry*>(s_raw);
    uint2* s_tab = nullptr;
    if (SHARED) {
        if (threadIdx.x == 0) s_hdr = a.hdrs[0];
        __syncthreads();
        const uint32_t nb = (s_hdr.flags & MODEL_REGULAR) ? s_hdr.nb : 0u;
        s_tab = reinterpret_cast<uint2*>(s_raw + (size_t)nb * sizeof(LutEntry));
        const uint4* gl = reinterpret_cast<const uint4*>(a.lut);
        uint4* sl = reinterpret_cast<uint4*>(s_lut);
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) sl[i] = gl[i];
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
    const uint32_t skip = (uint32_t)(off0 & 3u);
    GlobalFetch gf;
    gf.init(a.stream, off0, a.offsets[a.n_chunks]);
    DecSink<GlobalFetch> sink(gf);
    sink.prime(skip);  // src/decoder.rs:14-23

    uint64_t lo = 0, rg = ~0ull;
    uint32_t err = 0;

    auto step = [&]() -> uint32_t {
        uint64_t rpt = pow2 ? range_par_total<true>(rg, div) : range_par_total<false>(rg, div);
        uint64_t d = sink.data() - lo;  // examples/sample_impl.rs:29
        uint32_t sym;
        uint64_t P, rgn;
        bool ok = false;
        if (use_lut) {
            uint32_t b = lut_bucket(d, rg, lut_scale, max_bucket);
            LutEntry e = s_lut[b];
            ok = lut_resolve(e, d, rpt, sym, P, rgn);
        }
        if (RCB_UNLIKELY(!ok)) {
            sym = find_index_exact(d, rpt, K, [&](uint32_t i) { return tab[i].x; });
            uint2 e = tab[sym];
            P = rpt * (uint64_t)e.x;
            rgn = rpt * (uint64_t)e.y;
        }
        // param_update with the symbol's (c, cum): src/decoder.rs:42-50
        uint64_t nlo = lo + P;
        if (CHECKED && nlo < lo) {
            if (!err) err = ST_LOWER_OVERFLOW;
            nlo = 0;
            rgn = ~0ull;
        }
        lo = nlo;
        rg = rgn;
        renorm<CHECKED>(lo, rg, sink, err);  // consumes the same number of bytes (:52)
        return sym;
    };

    constexpr uint32_t PER = 4 / sizeof(SYM);  // symbols per 32-bit store
    uint64_t done = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
        uint32_t* dw = reinterpret_cast<uint32_t*>(dst);
        const uint64_t nw = cnt / PER;
#pragma unroll 1
        for (uint64_t i = 0; i < nw; i++) {
            uint32_t acc = 0;
#pragma unroll
            for (uint32_t b = 0; b < PER; b++) acc |= step() << (8 * sizeof(SYM) * b);
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
