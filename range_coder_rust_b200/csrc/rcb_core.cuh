// rcb_core.cuh -- per-symbol range-coder arithmetic shared by every kernel.
//
// Semantics restated from the reference crate (paths under /root/reference):
//   src/range_coder.rs:53-92   param_update  (range, lower, loop 1, loop 2)
//   src/range_coder.rs:95-100  left_shift
//   src/range_coder.rs:110-116 no_carry_expansion      ("loop 1")
//   src/range_coder.rs:126-135 range_reduction_expansion ("loop 2")
//   src/encoder.rs:40-46       finish (8 bytes of lower_bound, MSB first)
//   src/decoder.rs:31-54       shift_left_buffer / decode
//   examples/sample_impl.rs:27-45 find_index (rfreq + binary search)
//
// Design (not a port): one coder state per GPU lane means the per-symbol
// dependency chain IS the run time, so the hot path is written as straight-line
// 32-bit code: the byte loops become a closed form (n1 = clz(lower ^ upper) / 8
// whole bytes at once, literal loops kept as a rare slow path), range/total is
// a shift or a multiply-high reciprocal, byte emission and input refill are
// funnel shifts with predicated stores/loads instead of branches, and the
// decoder never divides: it classifies `data - lower` in the product domain
// (cum * rpt <= d  <=>  cum <= d / rpt for integers).
//
// Everything here is __host__ __device__ so the same code can be compiled by
// g++ into a test-only harness (tests/hostcore) and compared with the oracle
// without a GPU.  The product never runs the host instantiation.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RCB_HD __host__ __device__ __forceinline__
#else
#define RCB_HD inline
#endif

#define RCB_LIKELY(x) (__builtin_expect(!!(x), 1))
#define RCB_UNLIKELY(x) (__builtin_expect(!!(x), 0))

namespace rcb {

constexpr uint64_t TOP8 = 1ull << 56;   // src/range_coder.rs:23
constexpr uint64_t TOP16 = 1ull << 48;  // src/range_coder.rs:24

// per-chunk status word (0 = ok); values mirror include/rcb200.h
enum : uint32_t {
    ST_OK = 0,
    ST_ZERO_FREQ = 1,        // range became 0: the reference never returns (range_coder.rs:83-85)
    ST_LOWER_OVERFLOW = 2,   // range_coder.rs:68-81
    ST_UPPER_OVERFLOW = 3,   // range_coder.rs:111,138-146
    ST_SYMBOL_RANGE = 4,     // symbol >= K (sample_impl.rs:19 would panic)
    ST_OUT_CAPACITY = 5,     // staging row too small; the reported length is the needed size
    ST_TRUNCATED = 6,        // decoder.rs:33 pop_front on an empty buffer
    ST_RESTART = 7           // a restart point does not fit its chunk, or the lane before it ended elsewhere
};

// ---------------------------------------------------------------------------
// Restart point (build-defined side information, absent from the reference): the coder state in
// front of symbol r * restart_syms of a chunk, as the ENCODER saw it -- lower_bound and range
// (src/range_coder.rs:9-11) and the number of code bytes the chunk's Encoder had emitted by then.
// The reference decoder mirrors the encoder's (lower_bound, range) exactly (src/decoder.rs:42-52)
// and its `data` is the 8 code bytes from that position on (src/decoder.rs:31-35), so a decoder
// lane can enter the chunk there: the chunk's bytes stay the reference's bytes, and one chunk is
// decoded by several lanes.  `rg` may be the lane's range rounded down to a multiple of total_freq
// (the only use of range before the next update is range / total_freq, src/range_coder.rs:62).
// ---------------------------------------------------------------------------
struct Restart {
    uint64_t lo, rg;
    uint32_t pos;  // code bytes emitted before the symbol = offset of the decoder's 8-byte window
    uint32_t pad;
};

RCB_HD uint64_t umul64hi(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

RCB_HD uint32_t clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)x);
#else
    return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}

// high word of (hi:lo) << sh, sh in 0..31
RCB_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
#endif
}

// low word of (hi:lo) >> sh, sh in 0..31
RCB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? ((lo >> sh) | (hi << (32u - sh))) : lo;
#endif
}

// 32-bit halves of a 64-bit value (explicit, so the compiler keeps 32-bit compares 32-bit)
RCB_HD uint32_t hi32(uint64_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
    return hi;
#else
    return (uint32_t)(x >> 32);
#endif
}
RCB_HD uint32_t lo32(uint64_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
    return lo;
#else
    return (uint32_t)x;
#endif
}

// a * b + c (low 64 bits), a 64-bit, b 32-bit: two multiply-adds (the second one carries the high
// word), not the three instructions (wide multiply-add, multiply, add) the compiler emits on its own.
RCB_HD uint64_t mad64x32(uint64_t a, uint32_t b, uint64_t c) {
#if defined(__CUDA_ARCH__)
    const uint64_t t = (uint64_t)lo32(a) * (uint64_t)b + c;  // IMAD.WIDE with the 64-bit addend
    uint32_t h;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(h) : "r"(hi32(a)), "r"(b), "r"(hi32(t)));
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo32(t)), "r"(h));
    return r;
#else
    return a * (uint64_t)b + c;
#endif
}

RCB_HD uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0u, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}

// ---------------------------------------------------------------------------
// range / total_freq  (src/range_coder.rs:38-40), bit-exact without a divide.
//   power-of-two total : shift.
//   otherwise          : q = mulhi(range, floor(2^64/total)) is floor(range/total)
//                        or one less; one compare fixes it (error < 1 because
//                        range < 2^64 and 2^64 - magic*total < total).
// ---------------------------------------------------------------------------
struct DivParams {
    uint64_t magic;   // floor(2^64 / total), only used when !pow2
    uint32_t total;
    uint32_t shift;   // log2(total) when total is a power of two
};

template <bool POW2>
RCB_HD uint64_t range_par_total(uint64_t range, const DivParams& p) {
    if (POW2) return range >> p.shift;
    uint64_t q = umul64hi(range, p.magic);
    uint64_t rem = range - q * (uint64_t)p.total;
    return q + (rem >= (uint64_t)p.total ? 1u : 0u);
}

// host-side helper: fill DivParams for a total (total >= 1)
inline bool make_div_params(uint32_t total, DivParams* p, bool* pow2) {
    if (total == 0) return false;
    p->total = total;
    *pow2 = (total & (total - 1)) == 0;
    p->shift = 0;
    p->magic = 0;
    if (*pow2) {
        uint32_t s = 0;
        while ((1u << s) != total) s++;
        p->shift = s;
    } else {
        p->magic = (uint64_t)((((unsigned __int128)1) << 64) / total);
    }
    return true;
}

// ---------------------------------------------------------------------------
// Renormalisation after range/lower have been updated for a symbol.
// Sink::put(lower_hi32, sh) receives the top sh/8 bytes (sh in {0,8,16,24}) of
// lower_bound; Sink::put_byte(b) one byte (slow path).  Order of bytes equals
// the reference's emission order (src/encoder.rs:35).
//
// Fast path: loop 1 emits n1 = clz(lower ^ (lower+range)) / 8 bytes (equal top
// bytes imply range < 2^56 after each shift, so the shifted sum never wraps),
// and loop 2 does not fire when range << 8*n1 >= 2^48.  Anything else (n1 >= 4,
// loop 2, range == 0, overflow) takes the literal loops below.
// ---------------------------------------------------------------------------
template <bool CHECKED, class Sink>
RCB_HD void renorm_slow(uint64_t& lo, uint64_t& rg, Sink& sink, uint32_t& err) {
    if (rg == 0) {  // reference: infinite loop (zero-frequency symbol)
        if (!err) err = ST_ZERO_FREQ;
        lo = 0;
        rg = ~0ull;
        return;
    }
    // loop 1: src/range_coder.rs:83-85,110-116
    for (;;) {
        uint64_t up = lo + rg;
        if (CHECKED && up < lo) {  // upper_bound().unwrap() panics
            if (!err) err = ST_UPPER_OVERFLOW;
            lo = 0;
            rg = ~0ull;
            return;
        }
        if ((lo ^ up) >= TOP8) break;
        sink.put_byte((uint32_t)(lo >> 56));
        rg <<= 8;
        lo <<= 8;
    }
    // loop 2: src/range_coder.rs:87-89,126-135 (entered only after loop 1 stopped)
    while (rg < TOP16) {
        rg = ~lo & (TOP16 - 1);
        sink.put_byte((uint32_t)(lo >> 56));
        rg <<= 8;
        lo <<= 8;
    }
}

template <bool CHECKED, class Sink>
RCB_HD void renorm(uint64_t& lo, uint64_t& rg, Sink& sink, uint32_t& err) {
    uint64_t up = lo + rg;
    uint32_t xh = (uint32_t)((lo ^ up) >> 32);
    uint32_t sh = clz32(xh) & 24u;  // 8 * n1 for n1 in 0..3 (xh == 0 -> 32 & 24 == 0, caught below)
    uint64_t rg2 = rg << sh;
    bool fast = (xh != 0) && (rg2 >= TOP16);
    if (CHECKED) fast = fast && (up >= lo);
    if (RCB_LIKELY(fast)) {
        sink.put((uint32_t)(lo >> 32), sh);
        lo <<= sh;
        rg = rg2;
    } else {
        renorm_slow<CHECKED>(lo, rg, sink, err);
    }
}

// One symbol through param_update (src/range_coder.rs:53-92).
template <bool POW2, bool CHECKED, class Sink>
RCB_HD void update_symbol(uint64_t& lo, uint64_t& rg, uint32_t cum, uint32_t c,
                          const DivParams& p, Sink& sink, uint32_t& err) {
    uint64_t rpt = range_par_total<POW2>(rg, p);  // :62
    uint64_t add = rpt * (uint64_t)cum;           // :70
    rg = rpt * (uint64_t)c;                       // :65
    uint64_t nlo = lo + add;
    if (CHECKED && nlo < lo) {                    // :74-80
        if (!err) err = ST_LOWER_OVERFLOW;
        nlo = 0;
        rg = ~0ull;
    }
    lo = nlo;
    renorm<CHECKED>(lo, rg, sink, err);
}

// ---------------------------------------------------------------------------
// Encoder byte sink.  `pend` holds the nb (< 32, multiple of 8) pending bits in
// its low end (bits above nb are stale and never read).  New bytes enter with
// one funnel shift; when 32 or more bits are pending the oldest 32 are stored
// as one word -- branch-free: the store is predicated and the counters are
// updated with selects, so a symbol's emission is straight-line code.
// Store::word_if(p, pos, w) writes 4 bytes at byte offset pos when p
// (pos % 4 == 0, memory order = emission order); Store::byte(pos, b) one byte.
// pos keeps counting past `cap`, so the final length is the needed capacity.
// CHECK_CAP = false drops the capacity test (the caller guarantees room).
// ---------------------------------------------------------------------------
template <class Store, bool CHECK_CAP = true>
struct EncSink {
    uint32_t pend = 0;
    uint32_t nb = 0;
    uint32_t pos = 0;
    uint32_t cap;
    Store st;

    RCB_HD EncSink(Store s, uint32_t cap_) : cap(cap_), st(s) {}

    RCB_HD void push(uint32_t over, uint32_t merged, uint32_t sh) {
        // (over:merged) is the 64-bit value (old pend << sh) | new bits
        pend = merged;
        const uint32_t nb2 = nb + sh;
        const bool fl = nb2 >= 32u;
        const uint32_t w = funnel_r(merged, over, nb2);  // bits [nb2-32, nb2); nb2 & 31 == nb2 - 32
        const bool room = CHECK_CAP ? (pos + 4u <= cap) : true;
        st.word_if(fl && room, pos, bswap32(w));
        pos += fl ? 4u : 0u;
        nb = nb2 & 31u;
    }
    RCB_HD void put(uint32_t lo_hi, uint32_t sh) {
        push(funnel_l(pend, 0u, sh), funnel_l(lo_hi, pend, sh), sh);
    }
    RCB_HD void put_byte(uint32_t b) { push(pend >> 24, (pend << 8) | (b & 0xFFu), 8u); }
    // src/encoder.rs:40-46: 8 x left_shift, then drain the pending bytes.
    RCB_HD uint32_t finish(uint64_t lo) {
        for (int i = 0; i < 8; i++) {
            put_byte((uint32_t)(lo >> 56));
            lo <<= 8;
        }
        while (nb) {
            uint32_t b = (pend >> (nb - 8u)) & 0xFFu;
            if (pos < cap) st.byte(pos, b);
            pos += 1u;
            nb -= 8u;
        }
        return pos;
    }
    RCB_HD bool overflowed() const { return pos > cap; }
};

// ---------------------------------------------------------------------------
// Fused fast step for a power-of-two total 2^s with 24 <= s <= 31 and a
// CONSISTENT table (no overflow is reachable, src/range_coder.rs:68-81,138-146).
// The coder carries rpt = range >> s instead of range: after the update
//   range' = rpt * c,  lower' = lower + rpt * cum,  upper' = lower + rpt * (cum + c)
// loop 1 shifts both left by sh = 8 * n1 and the next symbol needs
//   rpt_next = (range' << sh) >> s = range' >> (s - sh)          (range' << sh is exact)
// so the renormalisation shift and the division collapse into one right shift.
// n1 comes from three compares on the high word of lower' ^ upper' (find-leading-one
// costs 19 cycles of latency on sm_100a, compare + select 9), and loop 2 does not
// fire iff range' << sh >= 2^48 iff rpt_next >= 2^(48-s).
// Returns false when the literal loops are needed (n1 >= 4 or loop 2); then the
// caller renormalises (nlo, rgp) with renorm_slow.  hi/sh describe the bytes to
// emit: the top sh/8 bytes of lower'.
// ---------------------------------------------------------------------------
// Three flavours of the same step (the lane always carries rpt = range / total):
//   FUSE_BIG   total = 2^s, 24 <= s <= 31: rpt_next = range' >> (s - sh)      (one shift)
//   FUSE_POW2  total = 2^s, any s        : rpt_next = (range' << sh) >> s     (two shifts)
//   FUSE_GEN   any total                 : rpt_next = (range' << sh) / total  (multiply-high reciprocal)
enum : int { FUSE_BIG = 0, FUSE_POW2 = 1, FUSE_GEN = 2 };
// decode_kernel's fused flavours additionally tell the two divide-free forms of FUSE_GEN apart
enum : int { FUSE_GEN_M2 = 3 };

struct FusedParams {
    uint32_t s;      // log2(total) (power-of-two totals)
    uint32_t k0;     // s - 24 (FUSE_BIG): shift when n1 == 3
    DivParams div;   // FUSE_GEN
};

RCB_HD FusedParams make_fused(const DivParams& div) {
    FusedParams fp;
    fp.s = div.shift;
    fp.k0 = div.shift >= 24u ? div.shift - 24u : 0u;
    fp.div = div;
    return fp;
}

// rpt of a lane that holds `range` (start of a chunk, or after the exact path)
template <int MODE>
RCB_HD uint64_t fused_rpt(uint64_t range, const FusedParams& fp) {
    return MODE == FUSE_GEN ? range_par_total<false>(range, fp.div) : (range >> fp.s);
}

// TPUT = true: the same step tuned for instruction count instead of latency (several warps per scheduler:
// restart points).  The shift comes from one count-leading-zeros (sh = clz(xh) & 24; equal high words --
// n1 >= 4 -- give sh = 0) and "loop 2 stays idle" is tested on the shifted range itself, range' << sh >= 2^48,
// which is exact for n1 <= 3 and fails for n1 >= 4 (range' < 2^32): 3-4 integer instructions instead of 8.
RCB_HD uint32_t tput_shift(uint32_t xh) {
#if defined(__CUDA_ARCH__)
    uint32_t f;  // position of the leading one, 0xFFFFFFFF for 0: clz & 24 == ~f & 24 (one find + one logic op)
    asm("bfind.u32 %0, %1;" : "=r"(f) : "r"(xh));
    return ~f & 24u;
#else
    return clz32(xh) & 24u;
#endif
}
RCB_HD bool tput_loop2_idle(uint64_t rgp, uint32_t sh) {
    return funnel_l(lo32(rgp), hi32(rgp), sh) >= (1u << 16);  // hi32(range' << sh) >= 2^16
}

template <int MODE = FUSE_BIG, bool TPUT = false>
RCB_HD bool fused_step(uint64_t lo, uint64_t rpt, uint32_t cum, uint32_t c, const FusedParams& fp,
                       uint64_t& nlo, uint64_t& rgp, uint64_t& nrpt, uint32_t& sh) {
    nlo = mad64x32(rpt, cum, lo);
    const uint64_t up = mad64x32(rpt, cum + c, lo);
    rgp = mad64x32(rpt, c, 0ull);
    const uint32_t xh = hi32(nlo) ^ hi32(up);
    if (TPUT) {
        sh = tput_shift(xh);
        if (MODE == FUSE_BIG) {
            nrpt = rgp >> ((fp.k0 + 24u) - sh);
            return nrpt >= (uint64_t)(1u << (24u - fp.k0));  // range' << sh >= 2^48 (see fused_decode_step)
        }
        const uint64_t y = rgp << sh;
        nrpt = fused_rpt<MODE>(y, fp);
        return hi32(y) >= (1u << 16);
    }
    const bool p1 = xh < (1u << 24), p2 = xh < (1u << 16), p3 = xh < (1u << 8);
    if (MODE == FUSE_BIG) {
        const uint32_t k = p2 ? (p3 ? fp.k0 : fp.k0 + 8u) : (p1 ? fp.k0 + 16u : fp.k0 + 24u);
        sh = (fp.k0 + 24u) - k;
        nrpt = rgp >> k;
    } else {
        sh = p2 ? (p3 ? 24u : 16u) : (p1 ? 8u : 0u);
        nrpt = fused_rpt<MODE>(rgp << sh, fp);
    }
    // loop 2 stays idle iff range' << sh >= 2^48 iff range' >= 2^(48-sh): a test on the high word
    // of range' against 2^16 / 2^8 / 1 (exact for n1 <= 2; n1 == 3 and n1 >= 4, where the high
    // word is 0, conservatively take the literal loops).  Ready as early as the shift amount.
    const uint32_t need = p2 ? 1u : (p1 ? (1u << 8) : (1u << 16));
    return hi32(rgp) >= need;
}

// ---------------------------------------------------------------------------
// FUSE_GEN without a divide on the dependency chain.  The fused general-total step needs
//   rpt_next = floor((range' << sh) / total),   range' = rpt * c      (src/range_coder.rs:62,65)
// and a 64-bit divide by a run-time constant costs ~60 cycles of latency after range' and sh are known
// (shift, multiply-high by the reciprocal, remainder fix-up) -- 2/3 of the whole fused step.  Instead each
// symbol carries cs = floor(c * 2^64 / total) (c < total), so that with Y = rpt << sh (< 2^64 because
// Y * c = range' << sh):
//   Y * c / total = Y * cs / 2^64 + Y * delta / 2^64,   0 <= delta < 1
//   q' = floor(Y * cs / 2^64) = bits [64-sh, 128-sh) of the 128-bit product rpt * cs
// The product starts from rpt alone (it runs beside the lower/upper multiply-adds) and sh only enters
// in a final funnel shift.  q' is the exact quotient unless frac(Y * cs / 2^64) + Y / 2^64 reaches 1:
// with G the top 32 bits of that fraction and hy = hi32(Y) a sufficient test is G + hy + 2 <= 2^32 --
// it fails for ~2^-30 of the symbols of a real table; those take the exact path (word re-code), so the
// result never depends on the approximation.
// ---------------------------------------------------------------------------
RCB_HD uint64_t recip_of_freq(uint32_t c, uint32_t total) {  // floor(c * 2^64 / total), c < total
    return (uint64_t)((((unsigned __int128)c) << 64) / total);
}

RCB_HD bool fused_rpt_cs(uint64_t rpt, uint64_t cs, uint32_t sh, uint64_t& nrpt) {
    const uint32_t rl = lo32(rpt), rh = hi32(rpt), cl = lo32(cs), ch = hi32(cs);
#if defined(__CUDA_ARCH__)
    // hi64 through the compiler's multiply-high (wide multiply-adds chained by carry predicates, 7
    // instructions); the middle limb is recomputed in wrapping 32-bit arithmetic (2 multiply-adds + 1 add)
    const uint64_t ph = __umul64hi(rpt, cs);  // bits 64..127 of rpt * cs
    const uint64_t p0 = (uint64_t)rl * cl;
    const uint32_t pm = hi32(p0) + rl * ch + rh * cl, pl = lo32(p0);  // bits 32..63, 0..31
#else
    const uint64_t p0 = (uint64_t)rl * cl;
    const uint64_t t1 = (uint64_t)rl * ch + hi32(p0);
    const uint64_t t2 = (uint64_t)rh * cl + lo32(t1);
    const uint64_t ph = (uint64_t)rh * ch + hi32(t1) + hi32(t2);  // bits 64..127 of rpt * cs
    const uint32_t pm = lo32(t2), pl = lo32(p0);                   // bits 32..63, 0..31
#endif
    const uint32_t q_lo = funnel_l(pm, lo32(ph), sh), q_hi = funnel_l(lo32(ph), hi32(ph), sh);
    nrpt = ((uint64_t)q_hi << 32) | q_lo;
    const uint32_t G = funnel_l(pl, pm, sh);   // top 32 bits of frac((rpt << sh) * cs / 2^64)
    const uint32_t hy = funnel_l(rl, rh, sh);  // hi32(rpt << sh)
    return (uint64_t)G + hy <= 0xFFFFFFFEull;
}

// The same idea with ONE table-wide constant, for totals >= 2^25: with l = floor(log2 total) and
// m2 = floor(2^(64+l) / total) (63 < log2 m2 < 64 for a total that is not a power of two)
//   (range' << sh) / total = (range' * m2 / 2^64) / 2^(l-sh) + e,   0 <= e < 2^-l
// so q' = hi64(range' * m2) >> (l - sh) is the exact quotient unless the l-sh bits shifted out are all
// ones (then the exact path decides).  No per-symbol constant and no extra table lookup -- the decoder's
// candidates would need a third shared-memory array for theirs -- at the price of starting from range'
// instead of rpt (a longer chain: the encoder keeps the cs form).  Flagged share ~ 2^-(l-sh): 1e-5 of the
// symbols for l = 30, 3e-4 for l = 25.
struct Recip2 {
    uint64_t m2;  // floor(2^(64+l) / total)
    uint32_t l;   // floor(log2 total), >= 25
};
RCB_HD bool recip2_ok(uint32_t total) { return total >= (1u << 25) && (total & (total - 1u)) != 0u; }
RCB_HD Recip2 make_recip2(uint32_t total) {
    Recip2 r;
    r.l = 31u - clz32(total);
    r.m2 = (uint64_t)((((unsigned __int128)1) << (64u + r.l)) / total);
    return r;
}
RCB_HD bool fused_rpt_m2(uint64_t rgp, uint32_t sh, const Recip2& k, uint64_t& nrpt) {
    const uint64_t ph = umul64hi(rgp, k.m2);
    const uint32_t D = k.l - sh;  // 1 .. 31
    nrpt = ph >> D;
    return ((~lo32(ph)) << (32u - D)) != 0u;  // the D bits shifted out are not all ones
}

// fused_step with the divide-free rpt_next (encoder, FUSE_GEN tables that carry cs)
template <bool TPUT = false>
RCB_HD bool fused_step_cs(uint64_t lo, uint64_t rpt, uint32_t cum, uint32_t c, uint64_t cs,
                          uint64_t& nlo, uint64_t& rgp, uint64_t& nrpt, uint32_t& sh) {
    nlo = mad64x32(rpt, cum, lo);
    const uint64_t up = mad64x32(rpt, cum + c, lo);
    rgp = mad64x32(rpt, c, 0ull);
    const uint32_t xh = hi32(nlo) ^ hi32(up);
    if (TPUT) {
        sh = tput_shift(xh);
        const bool exact = fused_rpt_cs(rpt, cs, sh, nrpt);
        return exact & tput_loop2_idle(rgp, sh);
    }
    const bool p1 = xh < (1u << 24), p2 = xh < (1u << 16), p3 = xh < (1u << 8);
    sh = p2 ? (p3 ? 24u : 16u) : (p1 ? 8u : 0u);
    const bool exact = fused_rpt_cs(rpt, cs, sh, nrpt);
    const uint32_t need = p2 ? 1u : (p1 ? (1u << 8) : (1u << 16));  // see fused_step
    return exact & (hi32(rgp) >= need);
}

// ... and with the table-wide constant (totals >= 2^25)
template <bool TPUT = false>
RCB_HD bool fused_step_m2(uint64_t lo, uint64_t rpt, uint32_t cum, uint32_t c, const Recip2& k,
                          uint64_t& nlo, uint64_t& rgp, uint64_t& nrpt, uint32_t& sh) {
    nlo = mad64x32(rpt, cum, lo);
    const uint64_t up = mad64x32(rpt, cum + c, lo);
    rgp = mad64x32(rpt, c, 0ull);
    const uint32_t xh = hi32(nlo) ^ hi32(up);
    if (TPUT) {
        sh = tput_shift(xh);
        const bool exact = fused_rpt_m2(rgp, sh, k, nrpt);
        return exact & tput_loop2_idle(rgp, sh);
    }
    const bool p1 = xh < (1u << 24), p2 = xh < (1u << 16), p3 = xh < (1u << 8);
    sh = p2 ? (p3 ? 24u : 16u) : (p1 ? 8u : 0u);
    const bool exact = fused_rpt_m2(rgp, sh, k, nrpt);
    const uint32_t need = p2 ? 1u : (p1 ? (1u << 8) : (1u << 16));  // see fused_step
    return exact & (hi32(rgp) >= need);
}

// ---------------------------------------------------------------------------
// Decoder input window (src/decoder.rs:9,31-35).  (dh:dl) is `data`, aligned
// with lower_bound; (wh:wl) holds the following bytes left-aligned with `cnt`
// valid bits, refilled 32 bits at a time from Fetch (peek_be32(): the next
// big-endian word of the stream; advance_if(p): consume it when p).
// Invariant between symbols: cnt >= 32.
// ---------------------------------------------------------------------------
template <class Fetch>
struct DecSink {
    uint32_t dh = 0, dl = 0;
    uint32_t wh = 0, wl = 0;
    uint32_t cnt = 0;
    Fetch f;

    RCB_HD explicit DecSink(Fetch fetch) : f(fetch) {}

    RCB_HD uint64_t data() const { return ((uint64_t)dh << 32) | dl; }

    RCB_HD void refill() {
        // branch-free: when cnt < 32 (then wl == 0) append the next big-endian word
        const bool need = cnt < 32u;
        const uint32_t be = f.peek_be32();
        wl = need ? funnel_r(0u, be, cnt) : wl;  // low word of (be:0) >> cnt
        wh = need ? (wh | (be >> (cnt & 31u))) : wh;
        cnt += need ? 32u : 0u;
        f.advance_if(need);
    }
    // Decoder::new: data = first 8 bytes big-endian (decoder.rs:21).  The first
    // fetched word holds skip_bytes bytes that precede the chunk.
    RCB_HD void prime(uint32_t skip_bytes) {
        uint32_t first = f.peek_be32();
        f.advance_if(true);
        wh = skip_bytes ? (first << (8u * skip_bytes)) : first;
        wl = 0;
        cnt = 32u - 8u * skip_bytes;
        refill();
        for (int i = 0; i < 8; i++) put_byte(0);
    }
    RCB_HD void put(uint32_t /*lo_hi*/, uint32_t sh) {
        dh = funnel_l(dl, dh, sh);
        dl = funnel_l(wh, dl, sh);
        wh = funnel_l(wl, wh, sh);
        wl <<= sh;
        cnt -= sh;
        refill();
    }
    RCB_HD void put_byte(uint32_t /*b*/) { put(0u, 8u); }
    // bytes shifted into data so far = fetched - skipped - still buffered
    RCB_HD uint32_t used(uint32_t words_fetched, uint32_t skip_bytes) const {
        return 4u * words_fetched - skip_bytes - (cnt >> 3);
    }
};

// ---------------------------------------------------------------------------
// Symbol lookup (examples/sample_impl.rs:27-45) in the product domain.
// Reference: rfreq = d / rpt; binary search for the smallest `left` with
// cum[left+1] > rfreq.  cum[i] <= d / rpt  <=>  cum[i] * rpt <= d, and
// cum[i] * rpt <= total * rpt <= range cannot wrap for a validated table.
// CumAt(i) returns cum[i] for 1 <= i <= K-1.
// ---------------------------------------------------------------------------
template <class CumAt>
RCB_HD uint32_t find_index_exact(uint64_t d, uint64_t rpt, uint32_t K, CumAt cum_at) {
    uint32_t left = 0, right = K - 1;
    while (left < right) {
        uint32_t mid = (left + right) >> 1;
        uint64_t prod = rpt * (uint64_t)cum_at(mid + 1);
        if (prod <= d)
            left = mid + 1;
        else
            right = mid;
    }
    return left;
}

// ---------------------------------------------------------------------------
// Table-driven lookup for a REGULAR table (cum[i+1] == cum[i] + c[i]).
// The rfreq axis [0,total) is cut into nb = min(total, 4096) buckets of equal (fractional) width
// total / nb.  Entry b describes the symbol A whose interval contains floor(b * total / nb) and the next
// non-zero symbol B:  [cumA,cumB) -> A, [cumB,cumC) -> B.
// (Entry b is built for the point 1/8 bucket below that start, which absorbs the
// error of the estimate.)  The bucket comes from a float estimate of
// d/rpt ~= d*total/range taken from the high words only (range >= 2^48, so each
// high word carries >= 16 significant bits: the estimate is within 1/16 bucket
// above and 3/16 below the truth), then the choice is verified exactly in the
// product domain; any miss (estimate off, more than
// one boundary in the bucket, clamp-to-K-1 case, garbage stream) falls back to
// find_index_exact, so the result never depends on float rounding.
// ---------------------------------------------------------------------------
struct LutEntry {
    uint32_t cumA, cumB, cumC;
    uint32_t syms;  // A | B << 16
};

struct ModelHdr {
    DivParams div;       // 16 bytes
    uint32_t flags;      // RCB_MODEL_* (include/rcb200.h)
    uint32_t min_c;      // smallest non-zero c_freq (staging bound)
    uint32_t nb;         // LUT buckets (shared model only, else 0)
    uint32_t wshift;     // unused (0); kept for the 48-byte layout
    float lut_scale;     // number of buckets as a float: bucket = rfreq * lut_scale / total
    uint32_t K;
    uint32_t pad0, pad1;
};

// MODEL_FULLC (internal): some symbol has c == total, for which floor(c * 2^64 / total) does not fit 64 bits
enum : uint32_t { MODEL_POW2 = 1, MODEL_CONSISTENT = 2, MODEL_REGULAR = 4, MODEL_FULLC = 8 };

RCB_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

RCB_HD uint32_t lut_bucket(uint64_t d, uint64_t rg, float scale, float max_bucket) {
    float fd = (float)(uint32_t)(d >> 32);
    float fr = (float)(uint32_t)(rg >> 32);  // >= 2^16 for a live coder state
    float bf = fd * (fast_rcp(fr) * scale);  // >= 0; the LUT itself carries the safety margin
    bf = bf > max_bucket ? max_bucket : bf;  // also tames inf/NaN from a dead state
    return (uint32_t)bf;
}

// Returns true when the entry resolves the symbol; then sym, P = rpt*cum[sym]
// and rgn = rpt*c[sym] are set.
RCB_HD bool lut_resolve(const LutEntry& e, uint64_t d, uint64_t rpt, uint32_t& sym,
                        uint64_t& P, uint64_t& rgn) {
    uint64_t PA = rpt * (uint64_t)e.cumA;
    uint64_t PB = rpt * (uint64_t)e.cumB;
    uint64_t PC = rpt * (uint64_t)e.cumC;
    bool takeB = d >= PB;
    P = takeB ? PB : PA;
    uint64_t PN = takeB ? PC : PB;
    sym = takeB ? (e.syms >> 16) : (e.syms & 0xFFFFu);
    rgn = PN - P;
    return (d - P) < rgn;  // P <= d < PN (unsigned wrap makes d < P fail)
}


// ---------------------------------------------------------------------------
// Fused decode step (same preconditions as fused_step, plus a REGULAR table):
// the candidates' new lower bounds lower + rpt * cum{A,B,C} come straight out of
// multiply-adds, `data >= lower + rpt*cumB` picks the candidate, and
// lower' <= data < upper' is the exact verification (d - P < rpt*c in disguise).
// The byte offset of the entry comes from the shift-free estimate below.
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Shift-free bucket estimate for the next symbol (decode hot loop).  After symbol n (interval
// [nlo, nlo + rpt_n*c_n) chosen, data unshifted) the next rfreq is
//   (data' - lower') / rpt_{n+1}  ~=  (data - nlo) * total / (rpt_n * c_n)
// because both sides of the division are shifted by the same bytes.  So the estimate needs neither
// the renormalisation shift nor a reciprocal of the new range: with q_n = 1/float(rpt_n >> sr)
// (computed while the table load of symbol n is in flight) and rc = 16*lut_scale*2^-sr / c_n stored
// in a second shared-memory array (one float pair per 16-byte entry),
//   bf = float(data - nlo) * (q_n * rc)        = byte offset of the next entry, fractional.
// Error: rpt_n >> sr keeps >= 16 bits (estimate <= 1/16 bucket high), floor in rpt_{n+1} (<= 1/32 low):
// inside the 1/8-bucket margin the table is built with; the choice is verified exactly anyway.
// ---------------------------------------------------------------------------
RCB_HD float u64_to_float(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn((float)hi32(x), 4294967296.0f, (float)lo32(x));
#else
    return (float)hi32(x) * 4294967296.0f + (float)lo32(x);
#endif
}
RCB_HD uint32_t fused_sr(const FusedParams& fp) { return fp.s < 32u ? 32u - fp.s : 0u; }
// general totals: the estimate uses the whole rpt (no pre-shift), so rc carries no 2^-sr
RCB_HD float lut_q_gen(uint64_t rpt) { return fast_rcp(u64_to_float(rpt)); }
RCB_HD float lut_rc16(uint32_t c, float lut_scale, uint32_t sr) {
    return c ? (16.0f * lut_scale / (float)(1u << sr)) / (float)c : 0.0f;
}
RCB_HD float lut_q(uint64_t rpt, uint32_t sr) { return fast_rcp((float)(uint32_t)(rpt >> sr)); }
// estimate from scratch (loop entry, after the exact path): d = data - lower, range = rpt << s
RCB_HD float lut_bf16_init(uint64_t d, uint64_t rg, float lut_scale) {
    return u64_to_float(d) * (fast_rcp(u64_to_float(rg)) * (16.0f * lut_scale));
}
RCB_HD uint32_t lut_offset16(float bf) {  // byte offset of a 16-byte entry, 4096 entries
#if defined(__CUDA_ARCH__)
    // round-toward-zero add of 2^23 leaves floor(bf) in the mantissa; no clamp: a huge estimate
    // (garbage stream) lands on some in-bounds entry that fails verification (exact fallback)
    return __float_as_uint(__fadd_rz(bf, 8388608.0f)) & 0xFFF0u;
#else
    if (!(bf < 65535.0f)) return 0xFFF0u;
    return (uint32_t)bf & 0xFFF0u;
#endif
}

struct FusedDec {
    uint64_t nlo, rgp, nrpt;
    uint32_t sym, sh;
    bool takeB;   // second candidate of the entry chosen
    bool inside;  // symbol verified: lower' <= data < upper'
    bool ok;      // ... and the fast renormalisation applies
};

template <int MODE = FUSE_BIG, bool TPUT = false>
RCB_HD FusedDec fused_decode_step(uint64_t lo, uint64_t rpt, uint64_t data, const LutEntry& e,
                                  const FusedParams& fp) {
    FusedDec r;
    const uint64_t loA = mad64x32(rpt, e.cumA, lo);
    const uint64_t loB = mad64x32(rpt, e.cumB, lo);
    const uint64_t loC = mad64x32(rpt, e.cumC, lo);
    const bool takeB = data >= loB;
    r.nlo = takeB ? loB : loA;
    const uint64_t up = takeB ? loC : loB;
    r.sym = takeB ? (e.syms >> 16) : (e.syms & 0xFFFFu);
    // lower' <= data < upper'  <=>  data - lower' < range' (unsigned: data < lower' wraps to a huge value)
    const bool inside = (data - r.nlo) < (up - r.nlo);
    r.rgp = up - r.nlo;
    const uint32_t xh = hi32(r.nlo) ^ hi32(up);
    if (TPUT) {
        r.sh = tput_shift(xh);
        r.takeB = takeB;
        r.inside = inside;
        if (MODE == FUSE_BIG) {
            // rpt_next = range' >> (s - sh); range' << sh >= 2^48  <=>  rpt_next >= 2^(48 - s)  (s >= 24)
            r.nrpt = r.rgp >> ((fp.k0 + 24u) - r.sh);
            r.ok = inside & (r.nrpt >= (uint64_t)(1u << (24u - fp.k0)));
        } else {
            const uint64_t y = r.rgp << r.sh;
            r.nrpt = fused_rpt<MODE>(y, fp);
            r.ok = inside & (hi32(y) >= (1u << 16));
        }
        return r;
    }
    const bool p1 = xh < (1u << 24), p2 = xh < (1u << 16), p3 = xh < (1u << 8);
    if (MODE == FUSE_BIG) {
        const uint32_t k = p2 ? (p3 ? fp.k0 : fp.k0 + 8u) : (p1 ? fp.k0 + 16u : fp.k0 + 24u);
        r.sh = (fp.k0 + 24u) - k;
        r.nrpt = r.rgp >> k;
    } else {
        r.sh = p2 ? (p3 ? 24u : 16u) : (p1 ? 8u : 0u);
        r.nrpt = fused_rpt<MODE>(r.rgp << r.sh, fp);
    }
    const uint32_t need = p2 ? 1u : (p1 ? (1u << 8) : (1u << 16));  // see fused_step
    r.takeB = takeB;
    r.inside = inside;
    r.ok = inside & (hi32(r.rgp) >= need);
    return r;
}


// The same step for a general total with divide-free rpt_next: (csA, csB) are recip_of_freq of the
// entry's two candidates.
template <bool TPUT = false>
RCB_HD FusedDec fused_decode_step_cs(uint64_t lo, uint64_t rpt, uint64_t data, const LutEntry& e, uint64_t csA,
                                     uint64_t csB) {
    FusedDec r;
    const uint64_t loA = mad64x32(rpt, e.cumA, lo);
    const uint64_t loB = mad64x32(rpt, e.cumB, lo);
    const uint64_t loC = mad64x32(rpt, e.cumC, lo);
    const bool takeB = data >= loB;
    r.nlo = takeB ? loB : loA;
    const uint64_t up = takeB ? loC : loB;
    r.sym = takeB ? (e.syms >> 16) : (e.syms & 0xFFFFu);
    const bool inside = (data - r.nlo) < (up - r.nlo);
    r.rgp = up - r.nlo;
    const uint32_t xh = hi32(r.nlo) ^ hi32(up);
    if (TPUT) {
        r.sh = tput_shift(xh);
        const bool exact = fused_rpt_cs(rpt, takeB ? csB : csA, r.sh, r.nrpt);
        r.takeB = takeB;
        r.inside = inside;
        r.ok = inside & exact & tput_loop2_idle(r.rgp, r.sh);
        return r;
    }
    const bool p1 = xh < (1u << 24), p2 = xh < (1u << 16), p3 = xh < (1u << 8);
    r.sh = p2 ? (p3 ? 24u : 16u) : (p1 ? 8u : 0u);
    const bool exact = fused_rpt_cs(rpt, takeB ? csB : csA, r.sh, r.nrpt);
    const uint32_t need = p2 ? 1u : (p1 ? (1u << 8) : (1u << 16));
    r.takeB = takeB;
    r.inside = inside;
    r.ok = inside & exact & (hi32(r.rgp) >= need);
    return r;
}


// ... and with the table-wide constant (totals >= 2^25): no per-candidate constants to look up
template <bool TPUT = false>
RCB_HD FusedDec fused_decode_step_m2(uint64_t lo, uint64_t rpt, uint64_t data, const LutEntry& e, const Recip2& k) {
    FusedDec r;
    const uint64_t loA = mad64x32(rpt, e.cumA, lo);
    const uint64_t loB = mad64x32(rpt, e.cumB, lo);
    const uint64_t loC = mad64x32(rpt, e.cumC, lo);
    const bool takeB = data >= loB;
    r.nlo = takeB ? loB : loA;
    const uint64_t up = takeB ? loC : loB;
    r.sym = takeB ? (e.syms >> 16) : (e.syms & 0xFFFFu);
    const bool inside = (data - r.nlo) < (up - r.nlo);
    r.rgp = up - r.nlo;
    const uint32_t xh = hi32(r.nlo) ^ hi32(up);
    if (TPUT) {
        r.sh = tput_shift(xh);
        const bool exact = fused_rpt_m2(r.rgp, r.sh, k, r.nrpt);
        r.takeB = takeB;
        r.inside = inside;
        r.ok = inside & exact & tput_loop2_idle(r.rgp, r.sh);
        return r;
    }
    const bool p1 = xh < (1u << 24), p2 = xh < (1u << 16), p3 = xh < (1u << 8);
    r.sh = p2 ? (p3 ? 24u : 16u) : (p1 ? 8u : 0u);
    const bool exact = fused_rpt_m2(r.rgp, r.sh, k, r.nrpt);
    const uint32_t need = p2 ? 1u : (p1 ? (1u << 8) : (1u << 16));
    r.takeB = takeB;
    r.inside = inside;
    r.ok = inside & exact & (hi32(r.rgp) >= need);
    return r;
}

// Renormalisation half of the fused decode step, for callers that pick the symbol themselves
// (row kernel: candidates from a thin LUT, then a short scan): from lower' and upper' of the
// chosen symbol to the shift, the next rpt and the "fast renormalisation applies" flag.
struct FusedRenorm {
    uint64_t rgp, nrpt;
    uint32_t sh;
    bool ok;
};

template <int MODE>
RCB_HD FusedRenorm fused_renorm(uint64_t nlo, uint64_t up, const FusedParams& fp) {
    FusedRenorm r;
    r.rgp = up - nlo;
    const uint32_t xh = hi32(nlo) ^ hi32(up);
    const bool p1 = xh < (1u << 24), p2 = xh < (1u << 16), p3 = xh < (1u << 8);
    if (MODE == FUSE_BIG) {
        const uint32_t k = p2 ? (p3 ? fp.k0 : fp.k0 + 8u) : (p1 ? fp.k0 + 16u : fp.k0 + 24u);
        r.sh = (fp.k0 + 24u) - k;
        r.nrpt = r.rgp >> k;
    } else {
        r.sh = p2 ? (p3 ? 24u : 16u) : (p1 ? 8u : 0u);
        r.nrpt = fused_rpt<MODE>(r.rgp << r.sh, fp);
    }
    const uint32_t need = p2 ? 1u : (p1 ? (1u << 8) : (1u << 16));  // see fused_step
    r.ok = hi32(r.rgp) >= need;
    return r;
}

}  // namespace rcb
