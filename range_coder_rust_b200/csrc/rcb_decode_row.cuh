// rcb_decode_row.cuh -- K5b: decode for tables that do not fit the fat-LUT kernel of
// rcb_decode.cuh: per-chunk (adaptive) models, large alphabets (K = 4096), and shared
// tables whose total is not a power of two >= 2^24.
//
// Same reference semantics (src/decoder.rs:14-54, examples/sample_impl.rs:27-45), same
// lane state and code-byte ring.  The table is the cum_freq row cum[0..K] (cum[K] = total)
// in shared memory -- one row per lane (TAB_LANE) or one for the block (TAB_SHARED) -- plus a
// thin bucket -> symbol LUT (u8 / u16 entries) built in the kernel prologue:
//   bucket b = floor(nb * (data - lower) / range)   (float estimate from the high words)
//   s0 = lut[b];  candidates s0 and s0+1 with bounds cum[s0], cum[s0+1], cum[s0+2]
// then the exact product-domain verification of fused_decode_step; a miss (several
// symbols in one bucket, estimate off, loop 2, ...) takes the out-of-line exact search.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "rcb_core.cuh"
#include "rcb_decode.cuh"
#include "rcb_encode.cuh"  // TAB_*, FM_*

namespace rcb {

constexpr uint32_t ROW_PAD = 5;  // row entries K .. K+4 hold total (odd row pitch for even K: lanes on distinct banks)

struct DecodeRowArgs {
    const uint8_t* stream;
    const uint64_t* offsets;  // [n_chunks+1]
    uint64_t n_syms;
    uint64_t chunk_syms;
    uint64_t n_chunks;
    const uint2* tabs;        // [n_models][K]
    const ModelHdr* hdrs;     // [n_models]
    uint32_t K;
    uint32_t lanes_per_block;
    uint32_t nb;              // LUT buckets (per lane or per block)
    void* out;
    uint32_t* status;
    DecSegment seg;           // rcb_decode.cuh
    const Restart* restart;   // restart points (rcb_decode.cuh: DecodeArgs)
    uint64_t restart_syms;
    uint32_t parts;           // lanes per chunk; lanes_per_block is a multiple of it
    uint32_t pf_words;        // start-up L2 prefetch per lane (rcb_decode.cuh: DecodeArgs)
};

// Literal renormalisation loops for a symbol that is already chosen (s.lo / s.rg hold lower' and
// range'): src/range_coder.rs:83-89 through the exact-path byte source.  Rare, out of line.
__device__ __noinline__ DecLaneState dec_renorm_exact(DecLaneState s) {
    GlobalFetch gf(s.base, s.rd, s.last, s.ring, s.ring_hi > RING_PIECES * 4 ? s.ring_hi - RING_PIECES * 4 : 0u,
                   s.ring_hi);
    DecSink<GlobalFetch> sink(gf);
    sink.dh = s.dh;
    sink.dl = s.dl;
    sink.wh = s.wh;
    sink.wl = s.wl;
    sink.cnt = s.cnt;
    renorm_slow<false>(s.lo, s.rg, sink, s.err);
    s.dh = sink.dh;
    s.dl = sink.dl;
    s.wh = sink.wh;
    s.wl = sink.wl;
    s.cnt = sink.cnt;
    s.rd = sink.f.idx;
    return s;
}

template <typename SYM, int TABLE, int FMODE, typename LUT_T>
__global__ void __launch_bounds__(512, 1) decode_row_kernel(DecodeRowArgs a) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ ModelHdr s_hdr;
    const uint32_t K = a.K, L = a.lanes_per_block, nb = a.nb;
    const uint32_t parts = a.parts;                              // lanes per chunk (restart points), 1 = whole chunks
    const uint32_t CB = L / parts;                               // chunks per block
    const uint64_t block_first = (uint64_t)blockIdx.x * CB;      // first chunk of this block
    const uint32_t row_words = K + ROW_PAD;  // cum[0..K-1], then total ROW_PAD times: candidates past K never verify
    const uint32_t n_rows = TABLE == TAB_LANE ? CB : 1u;         // the lanes of one chunk share its row and LUT
    // shared layout: rings[blockDim.x][RING_STRIDE] | rows[n_rows][K+ROW_PAD] u32 | luts[n_rows][nb] LUT_T
    uint8_t* s_ring = s_raw;
    uint32_t* s_rows = reinterpret_cast<uint32_t*>(s_raw + (size_t)blockDim.x * RING_STRIDE);
    LUT_T* s_luts = reinterpret_cast<LUT_T*>(s_rows + (size_t)n_rows * row_words);

    // ---- prologue: cum rows (coalesced) and bucket LUTs
    {
        const uint64_t left = a.n_chunks - block_first;
        const uint32_t lanes = TABLE == TAB_LANE ? (left < CB ? (uint32_t)left : CB) : 1u;  // rows to build
        for (uint32_t l = 0; l < lanes; l++) {
            const uint64_t model = TABLE == TAB_LANE ? block_first + l : 0;
            const uint2* t = a.tabs + model * K;
            for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) s_rows[l * row_words + i] = t[i].x;
            if (threadIdx.x < ROW_PAD) s_rows[l * row_words + K + threadIdx.x] = a.hdrs[model].div.total;
        }
        if (TABLE == TAB_SHARED && threadIdx.x == 0) s_hdr = a.hdrs[0];
        __syncthreads();
        if (TABLE == TAB_SHARED) {
            // entry b: symbol at the point 1/8 bucket below floor(b * total / nb)
            const uint64_t total = s_rows[K];
            const uint64_t margin = total / nb / 8 + 1;
            const bool nb_pow2 = (nb & (nb - 1)) == 0;
            const uint32_t nb_shift = 31u - (uint32_t)__clz((int)nb);
            for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
                uint64_t v0 = nb_pow2 ? ((uint64_t)b * total) >> nb_shift : (uint64_t)b * total / nb;
                v0 = v0 > margin ? v0 - margin : 0;
                uint32_t lft = 0, rgt = K - 1;  // symbol whose interval contains v0
                while (lft < rgt) {
                    const uint32_t mid = (lft + rgt) >> 1;
                    if ((uint64_t)s_rows[mid + 1] <= v0) lft = mid + 1; else rgt = mid;
                }
                s_luts[b] = (LUT_T)lft;
            }
        } else if (threadIdx.x < lanes * parts) {
            // the `parts` lanes of a chunk build its LUT together: each walks its share of the buckets and the
            // row beside them (both monotone) from the symbol a binary search finds for its first bucket.
            // (Dropping the zero-frequency symbols from the row was measured: the position -> symbol map costs
            // more LUT buckets than the compaction saves in scans.)
            const uint32_t my = threadIdx.x / parts, sub = threadIdx.x - my * parts;
            const uint32_t* row = s_rows + my * row_words;
            LUT_T* lut = s_luts + (size_t)my * nb;
            const uint64_t total = row[K];
            const uint64_t margin = total / nb / 8 + 1;
            const uint64_t step_q = total / nb, step_r = total % nb;  // floor(b*total/nb), incrementally
            const uint32_t b_lo = (uint32_t)((uint64_t)nb * sub / parts), b_hi = (uint32_t)((uint64_t)nb * (sub + 1) / parts);
            uint64_t q = (uint64_t)b_lo * total / nb, r = (uint64_t)b_lo * total % nb;
            uint32_t s = 0;
            if (b_lo) {  // symbol whose interval contains the first bucket's point
                const uint64_t v0 = q > margin ? q - margin : 0;
                uint32_t lft = 0, rgt = K - 1;
                while (lft < rgt) {
                    const uint32_t mid = (lft + rgt) >> 1;
                    if ((uint64_t)row[mid + 1] <= v0) lft = mid + 1; else rgt = mid;
                }
                s = lft;
            }
            for (uint32_t b = b_lo; b < b_hi; b++) {
                const uint64_t v0 = q > margin ? q - margin : 0;
                while (s < K - 1 && (uint64_t)row[s + 1] <= v0) s++;
                lut[b] = (LUT_T)s;
                q += step_q;
                r += step_r;
                if (r >= nb) {
                    r -= nb;
                    q++;
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x >= L) return;
    const LanePart lp = lane_part(block_first * parts + threadIdx.x, parts, a.n_chunks, a.n_syms, a.chunk_syms,
                                  a.restart_syms);
    if (!lp.has) return;
    const uint64_t chunk = lp.chunk, first = lp.first, chunk_cnt = lp.chunk_cnt;
    const uint32_t my_row = (uint32_t)(chunk - block_first);
    const bool seg_load = a.seg.state && a.seg.load, seg_save = a.seg.state && a.seg.save;
    uint64_t seg_begin = a.seg.state ? (a.seg.first < chunk_cnt ? a.seg.first : chunk_cnt) : 0;
    uint64_t seg_end =
        seg_save ? (a.seg.first + a.seg.syms < chunk_cnt ? a.seg.first + a.seg.syms : chunk_cnt) : chunk_cnt;
    if (parts > 1) {
        seg_begin = (uint64_t)lp.part * a.restart_syms;
        seg_end = seg_begin + a.restart_syms < chunk_cnt ? seg_begin + a.restart_syms : chunk_cnt;
    }
    const uint64_t cnt_all = seg_end - seg_begin;
    SYM* dst = reinterpret_cast<SYM*>(a.out) + first + seg_begin;

    const uint32_t* row = s_rows + (TABLE == TAB_LANE ? (size_t)my_row * row_words : 0);
    const LUT_T* lut = s_luts + (TABLE == TAB_LANE ? (size_t)my_row * nb : 0);
    const ModelHdr hdr = TABLE == TAB_SHARED ? s_hdr : a.hdrs[chunk];
    const DivParams div = hdr.div;
    const bool pow2 = (hdr.flags & MODEL_POW2) != 0;

    const uint64_t total_bytes = a.offsets[a.n_chunks];
    uint64_t off0 = a.offsets[chunk], off1 = a.offsets[chunk + 1];
    // caller-supplied offsets are validated per lane (see decode_kernel): a bad lane decodes nothing
    bool offsets_ok = off0 <= off1 && off1 - off0 >= 8 && off1 <= total_bytes;
    Restart rp{0ull, ~0ull, 0u, 0u};  // entry state: RangeCoder::new, or a restart point (see decode_kernel)
    bool restart_ok = true;
    if (lp.part) {
        rp = a.restart[chunk * (parts - 1u) + (lp.part - 1u)];
        restart_ok = offsets_ok && rp.rg != 0 && (uint64_t)rp.pos + 8 <= off1 - off0;
        if (!restart_ok) {
            offsets_ok = false;
            rp.pos = 0;
        }
    }
    if (!offsets_ok) off0 = off1 = 0;
    const uint64_t start = off0 + rp.pos;
    const uint64_t pb = start & ~15ull;
    const uint64_t readable = ((total_bytes + 15) & ~15ull) - pb;
    const uint32_t skip = (uint32_t)(start & 3u);
    const uint32_t rd0 = (uint32_t)((start - pb) >> 2);

    RingFill fill;
    fill.pbase = a.stream + pb;
    fill.wr = 0;
    fill.npieces = !offsets_ok ? 0u : readable > (0xFFFFFFF0ull << 4) ? 0xFFFFFFF0u : (uint32_t)(readable >> 4);
    const uint64_t cnt = offsets_ok ? cnt_all : 0;
    RingFetch rf;
    rf.ring = (uint32_t)__cvta_generic_to_shared(s_ring + (size_t)threadIdx.x * RING_STRIDE);
    rf.rd = seg_load ? a.seg.state[chunk].rd : rd0;
    rf.cur = 0;
    const uint32_t last_word = fill.npieces ? fill.npieces * 4 - 1 : 0u;

    constexpr uint32_t PF_WORDS = 256;
    uint32_t pf_next = rf.rd & ~3u;
    auto prefetch_to = [&](uint32_t upto_words) {
        while (pf_next < upto_words) {
            const uint64_t o = (uint64_t)pf_next * 4;
            if (offsets_ok && o < readable) {
                const uint64_t left = readable - o;
                const uint32_t g = (upto_words - pf_next < PF_WORDS ? upto_words - pf_next : PF_WORDS) * 4;
                prefetch_l2_bulk_dec(fill.pbase + o, (uint32_t)(left < g ? left : g));
            }
            pf_next += PF_WORDS;
        }
    };
    prefetch_to(pf_next + a.pf_words);
    fill.resync(rf);
    DecSink<RingFetch> sink(rf);

    uint64_t lo = rp.lo, rg = rp.rg;
    uint32_t err = 0;
    if (seg_load) {
        const DecResume st = a.seg.state[chunk];
        lo = st.lo;
        rg = st.rg;
        sink.dh = st.dh;
        sink.dl = st.dl;
        sink.wh = st.wh;
        sink.wl = st.wl;
        sink.cnt = st.cnt;
        err = st.err;
    } else {
        sink.prime(skip);  // src/decoder.rs:14-23
    }
    if (!offsets_ok) err = restart_ok ? ST_TRUNCATED : ST_RESTART;  // src/decoder.rs:33
    constexpr uint32_t PER = 4 / sizeof(SYM);
    constexpr uint32_t SYM_BITS = 8 * sizeof(SYM);
    const FusedParams fp = make_fused(div);
    const float fnb = (float)nb;
    const uint32_t nbm1 = nb - 1;

    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 3u) == 0;
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst);
    const uint64_t nw = aligned ? cnt / PER : 0;
    uint64_t done = 0;

    auto run = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        uint64_t rpt = fused_rpt<MODE>(rg, fp);
        float rinv = fast_rcp((float)hi32(rg)) * fnb;
        auto step = [&]() -> uint32_t {
            const uint64_t data = sink.data();
            // bucket estimate: nb * (data - lower) / range from the high words (>= 0)
            const float bf = (float)(sink.dh - hi32(lo)) * rinv;
            uint32_t b = __float_as_uint(__fadd_rz(bf, 8388608.0f)) & 0x7FFFFFu;
            b = b < nbm1 ? b : nbm1;
            uint32_t j = lut[b];  // first candidate symbol
            uint64_t nlo, up;
            if constexpr (TABLE == TAB_LANE) {
                // per-lane tables have coarse LUTs (a few hundred buckets): four candidates j .. j+3 with
                // bounds row[j .. j+4] (monotone: three compares pick one; more than four symbols in a
                // bucket get scanned below)
                const uint32_t* r4 = row + j;
                const uint32_t c0 = r4[0], c1 = r4[1], c2 = r4[2], c3 = r4[3], c4 = r4[4];
                const bool g1 = data >= mad64x32(rpt, c1, lo);
                const bool g2 = data >= mad64x32(rpt, c2, lo);
                const bool g3 = data >= mad64x32(rpt, c3, lo);
                const uint32_t cl = g2 ? (g3 ? c3 : c2) : (g1 ? c1 : c0);
                const uint32_t cu = g2 ? (g3 ? c4 : c3) : (g1 ? c2 : c1);
                j += (g1 ? 1u : 0u) + (g2 ? 1u : 0u) + (g3 ? 1u : 0u);
                nlo = mad64x32(rpt, cl, lo);
                up = mad64x32(rpt, cu, lo);
            } else {
                // block-wide tables afford a fine LUT (bucket <= smallest frequency): two candidates
                const uint64_t loA = mad64x32(rpt, row[j], lo);
                const uint64_t loB = mad64x32(rpt, row[j + 1u], lo);
                const uint64_t loC = mad64x32(rpt, row[j + 2u], lo);
                const bool takeB = data >= loB;
                nlo = takeB ? loB : loA;
                up = takeB ? loC : loB;
                j += takeB ? 1u : 0u;
            }
            if (RCB_UNLIKELY(!((data - nlo) < (up - nlo)))) {  // not lower' <= data < upper'
                // more symbols in this bucket than candidates (or the estimate was off): short scan in the
                // product domain -- the reference's search result is the largest s with
                // lower + rpt * cum[s] <= data, clamped to K-1 (examples/sample_impl.rs:33-44)
                j = j < K - 1u ? j : K - 1u;
#pragma unroll 1
                while (j > 0u && mad64x32(rpt, row[j], lo) > data) j--;
#pragma unroll 1
                while (j < K - 1u && mad64x32(rpt, row[j + 1u], lo) <= data) j++;
                nlo = mad64x32(rpt, row[j], lo);
                up = mad64x32(rpt, row[j + 1u], lo);
            }
            const uint32_t sym = j;
            const FusedRenorm r = fused_renorm<MODE>(nlo, up, fp);
            if (RCB_LIKELY(r.ok)) {
                sink.put(0u, r.sh);
                lo = nlo << r.sh;
                rpt = r.nrpt;
                rinv = fast_rcp((float)hi32(r.rgp << r.sh)) * fnb;
                return sym;
            }
            // the symbol is known, its renormalisation needs the literal loops (out of line)
            fill.drain();  // everything requested has landed: the exact path reads the ring
            const DecLaneState st{nlo, r.rgp, sink.dh, sink.dl, sink.wh, sink.wl, sink.cnt,
                                  reinterpret_cast<const uint32_t*>(fill.pbase), sink.f.rd, last_word, err, 0u,
                                  sink.f.ring, fill.wr * 4};
            const DecLaneState x = dec_renorm_exact(st);
            lo = x.lo;
            rg = x.rg;
            sink.dh = x.dh;
            sink.dl = x.dl;
            sink.wh = x.wh;
            sink.wl = x.wl;
            sink.cnt = x.cnt;
            sink.f.rd = x.rd;
            err = x.err;
            fill.after_exact(sink.f);
            rpt = fused_rpt<MODE>(rg, fp);
            rinv = fast_rcp((float)hi32(rg)) * fnb;
            return sym;
        };
#pragma unroll 1
        for (uint64_t i = 0; i < nw; i++) {
            fill.round(sink.f);
            uint32_t acc = 0;
#pragma unroll
            for (uint32_t k = 0; k < PER; k++) acc |= step() << (SYM_BITS * k);
            dw[i] = acc;
        }
        done = nw * PER;
#pragma unroll 1
        for (uint64_t i = done; i < cnt; i++) {
            fill.round(sink.f);
            dst[i] = (SYM)step();
        }
        rg = MODE == FUSE_GEN ? rpt * (uint64_t)div.total : rpt << fp.s;  // generic form (segment hand-over)
    };
    if constexpr (FMODE == FM_LANE) {
        if (pow2) run(std::integral_constant<int, FUSE_POW2>{});
        else run(std::integral_constant<int, FUSE_GEN>{});
    } else {
        run(std::integral_constant<int, FMODE>{});
    }

    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (seg_save) {  // the chunk continues in the next launch
        a.seg.state[chunk] = DecResume{lo, rg, sink.dh, sink.dl, sink.wh, sink.wl, sink.cnt, sink.f.rd, err, 0u};
        return;
    }
    const uint32_t used = sink.used(sink.f.rd - rd0, skip);
    if (parts > 1) {  // see decode_kernel: zeroed status, errors only; arrival state checked against the next record
        if (!err) {
            if (seg_end == chunk_cnt) {
                if ((uint64_t)rp.pos + used > off1 - off0) err = ST_TRUNCATED;
            } else {
                const Restart nx = a.restart[chunk * (parts - 1u) + lp.part];
                if (nx.lo != lo || nx.pos != rp.pos + used - 8u) err = ST_RESTART;
            }
        }
        if (err) atomicMax(a.status + chunk, err);
        return;
    }
    if (!err && (uint64_t)used > off1 - off0) err = ST_TRUNCATED;  // src/decoder.rs:33
    a.status[chunk] = err;
}

}  // namespace rcb
