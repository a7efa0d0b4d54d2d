// hostcore.cpp -- TEST INFRASTRUCTURE ONLY (never loaded by the package).
//
// Compiles the host instantiation of range_coder_rust_b200/csrc/rcb_core.cuh with
// g++ so the closed-form renormalisation, the byte sinks, the reciprocal division
// and the table-driven symbol lookup can be compared with the oracle on a machine
// without a GPU.  The loops below mirror encode_kernel / decode_kernel lane by lane.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "../../range_coder_rust_b200/csrc/rcb_core.cuh"

using namespace rcb;

namespace {

struct HostStore {
    uint8_t* row;
    void word_if(bool p, uint32_t pos, uint32_t w) const {
        if (p) memcpy(row + pos, &w, 4);
    }
    void byte(uint32_t pos, uint32_t b) const { row[pos] = (uint8_t)b; }
};

struct HostFetch {
    const uint8_t* base;  // aligned-down start
    uint64_t pos;         // next word offset from base
    const uint8_t* end;
    uint32_t* calls;
    uint32_t peek_be32() const {
        uint32_t w = 0;
        for (int i = 0; i < 4; i++) {
            const uint8_t* p = base + pos + i;
            uint32_t b = p < end ? *p : 0u;
            w = (w << 8) | b;
        }
        return w;
    }
    void advance_if(bool p) {
        if (p) {
            ++*calls;
            pos += 4;
        }
    }
};

uint32_t load_sym(const uint8_t* syms, uint64_t i, int sb) {
    if (sb == 1) return syms[i];
    return (uint32_t)syms[2 * i] | ((uint32_t)syms[2 * i + 1] << 8);
}
void store_sym(uint8_t* out, uint64_t i, int sb, uint32_t v) {
    if (sb == 1) {
        out[i] = (uint8_t)v;
    } else {
        out[2 * i] = (uint8_t)v;
        out[2 * i + 1] = (uint8_t)(v >> 8);
    }
}

// mirrors finalize_models_kernel
void build_header(uint32_t K, const uint32_t* c, const uint32_t* cum, uint32_t total, uint32_t lut_cap,
                  ModelHdr& h, std::vector<LutEntry>& lut) {
    bool pow2;
    make_div_params(total, &h.div, &pow2);
    uint32_t bad = 0;
    for (uint32_t i = 0; i < K; i++) {
        uint64_t end = (uint64_t)cum[i] + c[i];
        if (end > total) bad |= 1;
        if (i + 1 < K ? (uint64_t)cum[i + 1] != end : end != total) bad |= 2;
    }
    h.flags = (pow2 ? MODEL_POW2 : 0) | ((bad & 1) ? 0 : MODEL_CONSISTENT) | ((bad & 3) ? 0 : MODEL_REGULAR);
    h.nb = total < lut_cap ? total : lut_cap;
    h.wshift = 0;
    h.lut_scale = (float)h.nb;
    h.K = K;
    lut.clear();
    if (!(h.flags & MODEL_REGULAR)) return;
    lut.resize(h.nb);
    for (uint32_t b = 0; b < h.nb; b++) {
        uint64_t v0 = (uint64_t)b * total / h.nb;
        const uint64_t margin = (total / h.nb) >> 3;
        v0 = v0 > margin ? v0 - margin : 0;
        uint32_t left = 0, right = K - 1;
        while (left < right) {
            uint32_t mid = (left + right) >> 1;
            if ((uint64_t)cum[mid + 1] <= v0) left = mid + 1; else right = mid;
        }
        uint32_t A = left, B = A + 1;
        while (B < K && c[B] == 0) B++;
        LutEntry e;
        e.cumA = cum[A];
        e.cumB = cum[A] + c[A];
        e.cumC = e.cumB + (B < K ? c[B] : 0);
        e.syms = A | ((B & 0xFFFFu) << 16);
        lut[b] = e;
    }
}

}  // namespace

extern "C" int64_t hc_encode(const uint8_t* syms, uint64_t n, int sym_bytes, uint32_t K, const uint32_t* c,
                             const uint32_t* cum, uint32_t total, uint8_t* out, uint32_t cap, int checked,
                             uint32_t* status) {
    DivParams div;
    bool pow2;
    if (!make_div_params(total, &div, &pow2)) return -1;
    uint64_t lo = 0, rg = ~0ull;
    uint32_t err = 0;
    HostStore hs{out};
    EncSink<HostStore> sink(hs, cap);
    // fused paths (mirror encode_kernel's FM_BIG / FM_POW2 / FM_GEN instantiations), selected like
    // the host dispatch does: any consistent table, flavour by the total
    bool consistent = true;
    for (uint32_t i = 0; i < K; i++)
        if ((uint64_t)cum[i] + c[i] > total) consistent = false;
    if (consistent && !checked) {
        const FusedParams fp = make_fused(div);
        auto run = [&](auto tag) -> int64_t {
            constexpr int MODE = decltype(tag)::value;
            uint64_t rpt = fused_rpt<MODE>(rg, fp);
            for (uint64_t i = 0; i < n; i++) {
                uint32_t s = load_sym(syms, i, sym_bytes);
                if (s >= K) {
                    if (!err) err = ST_SYMBOL_RANGE;
                    s = 0;
                }
                uint64_t nlo, rgp, nrpt;
                uint32_t sh;
                if (fused_step<MODE>(lo, rpt, cum[s], c[s], fp, nlo, rgp, nrpt, sh)) {
                    sink.put((uint32_t)(nlo >> 32), sh);
                    lo = nlo << sh;
                    rpt = nrpt;
                } else {
                    lo = nlo;
                    rg = rgp;
                    renorm_slow<false>(lo, rg, sink, err);
                    rpt = fused_rpt<MODE>(rg, fp);
                }
            }
            uint32_t len = sink.finish(lo);
            if (!err && sink.overflowed()) err = ST_OUT_CAPACITY;
            *status = err;
            return (int64_t)len;
        };
        if (pow2 && div.shift >= 24) return run(std::integral_constant<int, FUSE_BIG>{});
        if (pow2) return run(std::integral_constant<int, FUSE_POW2>{});
        bool full = false;  // a symbol with c == total has no reciprocal constant (plan_encode: FM_GEN)
        for (uint32_t i = 0; i < K; i++) full |= c[i] == total;
        if (recip2_ok(total)) {  // table-wide reciprocal (mirrors encode_kernel's FM_GENM2 instantiation)
            const Recip2 k2 = make_recip2(total);
            uint64_t rpt = fused_rpt<FUSE_GEN>(rg, fp);
            for (uint64_t i = 0; i < n; i++) {
                uint32_t s = load_sym(syms, i, sym_bytes);
                if (s >= K) {
                    if (!err) err = ST_SYMBOL_RANGE;
                    s = 0;
                }
                uint64_t nlo, rgp, nrpt;
                uint32_t sh;
                if (fused_step_m2(lo, rpt, cum[s], c[s], k2, nlo, rgp, nrpt, sh)) {
                    if (nrpt != (rgp << sh) / total) return -2;
                    sink.put((uint32_t)(nlo >> 32), sh);
                    lo = nlo << sh;
                    rpt = nrpt;
                } else {
                    lo = nlo;
                    rg = rgp;
                    renorm_slow<false>(lo, rg, sink, err);
                    rpt = fused_rpt<FUSE_GEN>(rg, fp);
                }
            }
            uint32_t len = sink.finish(lo);
            if (!err && sink.overflowed()) err = ST_OUT_CAPACITY;
            *status = err;
            return (int64_t)len;
        }
        if (full) return run(std::integral_constant<int, FUSE_GEN>{});
        // divide-free general total (mirrors encode_kernel's FM_GENCS instantiation)
        std::vector<uint64_t> cs(K);
        for (uint32_t i = 0; i < K; i++) cs[i] = recip_of_freq(c[i], total);
        uint64_t rpt = fused_rpt<FUSE_GEN>(rg, fp);
        for (uint64_t i = 0; i < n; i++) {
            uint32_t s = load_sym(syms, i, sym_bytes);
            if (s >= K) {
                if (!err) err = ST_SYMBOL_RANGE;
                s = 0;
            }
            uint64_t nlo, rgp, nrpt;
            uint32_t sh;
            if (fused_step_cs(lo, rpt, cum[s], c[s], cs[s], nlo, rgp, nrpt, sh)) {
                if (nrpt != (rgp << sh) / total) return -2;  // the exactness test must never pass a wrong quotient
                sink.put((uint32_t)(nlo >> 32), sh);
                lo = nlo << sh;
                rpt = nrpt;
            } else {
                lo = nlo;
                rg = rgp;
                renorm_slow<false>(lo, rg, sink, err);
                rpt = fused_rpt<FUSE_GEN>(rg, fp);
            }
        }
        uint32_t len = sink.finish(lo);
        if (!err && sink.overflowed()) err = ST_OUT_CAPACITY;
        *status = err;
        return (int64_t)len;
    }
    for (uint64_t i = 0; i < n; i++) {
        uint32_t s = load_sym(syms, i, sym_bytes);
        if (s >= K) {
            if (!err) err = ST_SYMBOL_RANGE;
            s = 0;
        }
        if (pow2) {
            if (checked) update_symbol<true, true>(lo, rg, cum[s], c[s], div, sink, err);
            else update_symbol<true, false>(lo, rg, cum[s], c[s], div, sink, err);
        } else {
            if (checked) update_symbol<false, true>(lo, rg, cum[s], c[s], div, sink, err);
            else update_symbol<false, false>(lo, rg, cum[s], c[s], div, sink, err);
        }
    }
    uint32_t len = sink.finish(lo);
    if (!err && sink.overflowed()) err = ST_OUT_CAPACITY;
    *status = err;
    return (int64_t)len;
}

// stream/off0/off1: chunk bytes are stream[off0..off1); stream_len bounds the reads.
extern "C" int64_t hc_decode(const uint8_t* stream, uint64_t off0, uint64_t off1, uint64_t stream_len,
                             uint64_t n_syms, int sym_bytes, uint32_t K, const uint32_t* c, const uint32_t* cum,
                             uint32_t total, uint8_t* out, int use_lut, int checked, uint32_t lut_cap,
                             uint32_t* status, uint64_t* n_fallback) {
    ModelHdr h;
    std::vector<LutEntry> lut;
    if (total == 0) return -1;
    build_header(K, c, cum, total, lut_cap, h, lut);
    const bool pow2 = (h.flags & MODEL_POW2) != 0;
    const bool lut_ok = use_lut && (h.flags & MODEL_REGULAR);
    const float max_bucket = (float)(h.nb ? h.nb - 1 : 0);
    // emulate an unaligned chunk start inside an aligned buffer
    uint64_t al = off0 & ~3ull;
    uint32_t calls = 0;
    const uint32_t skip = (uint32_t)(off0 & 3u);
    HostFetch hf{stream + al, 0, stream + stream_len, &calls};
    DecSink<HostFetch> sink(hf);
    sink.prime(skip);
    uint64_t lo = 0, rg = ~0ull, fallbacks = 0;
    uint32_t err = 0;
    // fused fat-LUT path (mirrors decode_kernel's FUSE_BIG / FUSE_POW2 / FUSE_GEN instantiations)
    bool full = false;
    for (uint32_t i = 0; i < K; i++) full |= c[i] == total;
    if (lut_ok && total >= 2 && !(!pow2 && full) && (h.flags & MODEL_CONSISTENT) && !checked && lut_cap == 4096) {
        const int mode = pow2 ? (h.div.shift >= 24 ? FUSE_BIG : FUSE_POW2) : FUSE_GEN;
        FusedParams fp = make_fused(h.div);
        std::vector<LutEntry> pad(4096);
        for (uint32_t b = 0; b < 4096; b++) {
            if (b < h.nb) pad[b] = lut[b];
            else pad[b] = LutEntry{total, total, total, 0};
        }
        auto rpt_of = [&](uint64_t range) {
            return mode == FUSE_GEN ? fused_rpt<FUSE_GEN>(range, fp) : (range >> fp.s);
        };
        auto range_of = [&](uint64_t r) { return mode == FUSE_GEN ? r * (uint64_t)total : r << fp.s; };
        uint64_t rpt = rpt_of(rg);
        // shift-free estimate, as in decode_kernel: q = 1/float(rpt >> sr), rc = const / c of the candidate
        const uint32_t sr = mode == FUSE_GEN ? 0u : fused_sr(fp);
        auto q_of = [&](uint64_t r) { return mode == FUSE_GEN ? lut_q_gen(r) : lut_q(r, sr); };
        float q = q_of(rpt);
        float bf = lut_bf16_init(sink.data() - lo, rg, h.lut_scale);
        for (uint64_t i = 0; i < n_syms; i++) {
            const uint64_t data = sink.data();
            const uint32_t off = lut_offset16(bf);
            const LutEntry e = pad[off >> 4];
            const float rcA = lut_rc16(e.cumB - e.cumA, h.lut_scale, sr);
            const float rcB = lut_rc16(e.cumC - e.cumB, h.lut_scale, sr);
            FusedDec r;
            if (mode == FUSE_BIG) r = fused_decode_step<FUSE_BIG>(lo, rpt, data, e, fp);
            else if (mode == FUSE_POW2) r = fused_decode_step<FUSE_POW2>(lo, rpt, data, e, fp);
            else {
                // decode_kernel: table-wide constant for totals >= 2^25, per-candidate constants below
                if (recip2_ok(total)) r = fused_decode_step_m2(lo, rpt, data, e, make_recip2(total));
                else r = fused_decode_step_cs(lo, rpt, data, e, recip_of_freq(e.cumB - e.cumA, total),
                                              recip_of_freq(e.cumC - e.cumB, total));
                if (r.ok && r.nrpt != (r.rgp << r.sh) / total) return -2;  // exactness test passed a wrong quotient
            }
            uint32_t sym;
            if (r.ok) {
                sym = r.sym;
                bf = u64_to_float(data - r.nlo) * (q * (r.takeB ? rcB : rcA));
                sink.put(0, r.sh);
                lo = r.nlo << r.sh;
                rpt = r.nrpt;
                q = q_of(rpt);
            } else {
                fallbacks++;
                sym = find_index_exact(data - lo, rpt, K, [&](uint32_t j) { return cum[j]; });
                lo = lo + rpt * (uint64_t)cum[sym];
                rg = rpt * (uint64_t)c[sym];
                renorm<false>(lo, rg, sink, err);
                rpt = rpt_of(rg);
                q = q_of(rpt);
                bf = lut_bf16_init(sink.data() - lo, range_of(rpt), h.lut_scale);
            }
            store_sym(out, i, sym_bytes, sym);
        }
        const uint32_t used = sink.used(calls, skip);
        if (!err && (uint64_t)used > off1 - off0) err = ST_TRUNCATED;
        *status = err;
        if (n_fallback) *n_fallback = fallbacks;
        return (int64_t)used;
    }
    for (uint64_t i = 0; i < n_syms; i++) {
        uint64_t rpt = pow2 ? range_par_total<true>(rg, h.div) : range_par_total<false>(rg, h.div);
        uint64_t d = sink.data() - lo;
        uint32_t sym = 0;
        uint64_t P = 0, rgn = 0;
        bool ok = false;
        if (lut_ok) {
            uint32_t b = lut_bucket(d, rg, h.lut_scale, max_bucket);
            ok = lut_resolve(lut[b], d, rpt, sym, P, rgn);
        }
        if (!ok) {
            fallbacks++;
            sym = find_index_exact(d, rpt, K, [&](uint32_t j) { return cum[j]; });
            P = rpt * (uint64_t)cum[sym];
            rgn = rpt * (uint64_t)c[sym];
        }
        uint64_t nlo = lo + P;
        if (checked && nlo < lo) {
            if (!err) err = ST_LOWER_OVERFLOW;
            nlo = 0;
            rgn = ~0ull;
        }
        lo = nlo;
        rg = rgn;
        if (checked) renorm<true>(lo, rg, sink, err);
        else renorm<false>(lo, rg, sink, err);
        store_sym(out, i, sym_bytes, sym);
    }
    const uint32_t used = sink.used(calls, skip);
    if (!err && (uint64_t)used > off1 - off0) err = ST_TRUNCATED;
    *status = err;
    if (n_fallback) *n_fallback = fallbacks;
    return (int64_t)used;
}

// exhaustive check helper for the reciprocal: returns the number of mismatches
extern "C" uint64_t hc_check_division(const uint64_t* ranges, uint64_t n, uint32_t total) {
    DivParams div;
    bool pow2;
    if (!make_div_params(total, &div, &pow2)) return ~0ull;
    uint64_t bad = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t q = pow2 ? range_par_total<true>(ranges[i], div) : range_par_total<false>(ranges[i], div);
        if (q != ranges[i] / total) bad++;
    }
    return bad;
}

// Divide-free rpt_next (fused_rpt_cs): for every (rpt, c, sh) with (rpt * c) << sh < 2^64, a passing
// exactness test must come with the exact quotient.  Returns the number of wrong quotients that passed
// (must be 0) and, in *n_inexact, how many cases the test sent to the exact path.
extern "C" uint64_t hc_check_cs(const uint64_t* rpts, const uint32_t* cs_c, const uint32_t* shs, uint64_t n,
                                uint32_t total, uint64_t* n_inexact) {
    uint64_t wrong = 0, inexact = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t rpt = rpts[i];
        const uint32_t c = cs_c[i], sh = shs[i];
        const unsigned __int128 x = ((unsigned __int128)rpt * c) << sh;
        if (c == 0 || c >= total || (x >> 64) != 0) continue;  // outside the step's preconditions
        uint64_t nrpt;
        if (fused_rpt_cs(rpt, recip_of_freq(c, total), sh, nrpt)) {
            if (nrpt != (uint64_t)x / total) wrong++;
        } else {
            inexact++;
        }
    }
    if (n_inexact) *n_inexact = inexact;
    return wrong;
}

// fused_rpt_m2 (table-wide constant, totals >= 2^25): same contract as hc_check_cs, on range' directly.
extern "C" uint64_t hc_check_m2(const uint64_t* rgps, const uint32_t* shs, uint64_t n, uint32_t total,
                                uint64_t* n_inexact) {
    if (!recip2_ok(total)) return ~0ull;
    const Recip2 k = make_recip2(total);
    uint64_t wrong = 0, inexact = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t rgp = rgps[i];
        const uint32_t sh = shs[i];
        if ((((unsigned __int128)rgp) << sh) >> 64) continue;  // outside the step's preconditions
        uint64_t nrpt;
        if (fused_rpt_m2(rgp, sh, k, nrpt)) {
            if (nrpt != (rgp << sh) / total) wrong++;
        } else {
            inexact++;
        }
    }
    if (n_inexact) *n_inexact = inexact;
    return wrong;
}
