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
// refill, select trees instead of find-leading-one, (FUSED) range/total folded
// into the renormalisation shift, and rare events handled per 32-bit word of
// output by checkpoint + exact re-decode (see rcb_encode.cuh).
//
// Code bytes reach a lane through a private 128-byte ring in shared memory
// filled by cp.async (LDGSTS) in 16-byte pieces, issued warp-synchronously once
// per output word.  Lanes consume their streams at data-dependent rates, and the
// register scoreboard is per warp: any global load into a register that a later
// (even predicated-off) instruction reads stalls the whole warp for the load's
// latency.  cp.async has no destination register, so the only loads on the
// symbol path are shared-memory loads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcb_core.cuh"

namespace rcb {

// Lane state between two segment launches of one chunk (host-buffer pipeline: a chunk is decoded in
// a few launches so that the device->host copy of its first symbols starts before its last ones exist).
struct DecResume {
    uint64_t lo, rg;
    uint32_t dh, dl, wh, wl, cnt, rd, err, pad;
};
struct DecSegment {
    uint64_t first;    // first symbol of every chunk decoded by this launch
    uint64_t syms;     // symbols per chunk in this launch (used when save != 0; else: to the chunk's end)
    DecResume* state;  // [n_chunks]; nullptr: the launch decodes whole chunks
    uint32_t load;     // continue from state (not the first segment)
    uint32_t save;     // store state at the end (not the last segment)
};

struct DecodeArgs {
    const uint8_t* stream;
    const uint64_t* offsets;  // [n_chunks+1]
    uint64_t n_syms;
    uint64_t chunk_syms;
    uint64_t n_chunks;
    const uint2* tabs;
    const ModelHdr* hdrs;
    const LutEntry* lut;      // shared model only
    const uint4* lut_cs;      // shared model: {csA lo, csA hi, csB lo, csB hi} per LUT entry (FUSE_GEN)
    uint32_t K;
    uint32_t per_chunk;
    void* out;
    uint32_t* status;
    DecSegment seg;
    // restart points (rcb_core.cuh: Restart): `parts` lanes per chunk, lane p > 0 enters at record p-1
    const Restart* restart;  // [n_chunks][parts - 1]
    uint64_t restart_syms;
    uint32_t parts;          // 1: one lane per chunk (no restart points)
    // start-up L2 prefetch per lane, in 32-bit words (multiple of 64 = 256 bytes; 0: none).  The host sizes it so
    // that the lanes resident at one time prefetch no more than a fraction of L2 (2 KiB per lane for one lane
    // per chunk; 262144 lanes of a few KB each would push 0.5 GB through a 126 MB L2 and read it twice)
    uint32_t pf_words;
};

// Which part of which chunk a lane decodes, and from which state (restart points).
struct LanePart {
    uint64_t chunk, first, chunk_cnt;
    uint32_t part;
    bool has;  // the lane has symbols to decode
};
__device__ __forceinline__ LanePart lane_part(uint64_t lane_id, uint32_t parts, uint64_t n_chunks, uint64_t n_syms,
                                              uint64_t chunk_syms, uint64_t restart_syms) {
    LanePart l;
    l.chunk = parts > 1 ? lane_id / parts : lane_id;
    l.part = parts > 1 ? (uint32_t)(lane_id - l.chunk * parts) : 0u;
    l.first = l.chunk * chunk_syms;
    l.has = l.chunk < n_chunks;
    l.chunk_cnt = 0;
    if (l.has) {
        l.chunk_cnt = (n_syms - l.first < chunk_syms) ? (n_syms - l.first) : chunk_syms;
        l.has = l.part == 0 || (uint64_t)l.part * restart_syms < l.chunk_cnt;  // ragged last chunk: fewer parts
    }
    return l;
}

constexpr uint32_t RING_PIECES = 8;                   // 16-byte pieces per lane
constexpr uint32_t RING_STRIDE = RING_PIECES * 16 + 16;  // 144 bytes: 16-byte aligned, spreads banks

// ---------------------------------------------------------------------------
// Fetch for the hot loops: words come from the lane's shared-memory ring.
// Positions are absolute word / piece indices from `pbase` (the 16-byte aligned
// address at or below the chunk's first byte).
// ---------------------------------------------------------------------------
struct RingFetch {
    uint32_t ring;   // shared-space address of this lane's ring
    uint32_t rd;     // words consumed; `cur` holds word rd
    uint32_t cur;
    __device__ __forceinline__ uint32_t slot(uint32_t w) const { return ring + (w & (RING_PIECES * 4 - 1)) * 4; }
    __device__ __forceinline__ uint32_t peek_be32() const { return bswap32(cur); }
    __device__ __forceinline__ void advance_if(bool p) {
        rd += p ? 1u : 0u;
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q ld.shared.b32 %0, [%2];\n\t}"
                     : "+r"(cur)
                     : "r"((uint32_t)p), "r"(slot(rd))
                     : "memory");
    }
    __device__ __forceinline__ void reload() {
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(cur) : "r"(slot(rd)) : "memory");
    }
};

// Ring producer side (per lane): pieces [.., wr) have been requested.
// MIRROR: a piece that goes to slot 0 is also copied behind slot 7 (the 16 bytes of padding of RING_STRIDE), so
// that 16 contiguous bytes can be read from ANY word of the ring without wrapping the address (position window
// of the fused loop).
template <bool MIRROR>
struct RingFillT {
    const uint8_t* pbase;  // 16-byte aligned
    uint32_t wr;           // pieces requested so far
    uint32_t npieces;      // pieces that exist in the readable stream
    __device__ __forceinline__ void issue_if(bool p, uint32_t ring) {
        const uint32_t saddr = ring + (wr & (RING_PIECES - 1)) * 16;
        const uint8_t* g = pbase + (uint64_t)wr * 16;
        asm volatile(
            "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q cp.async.cg.shared.global.L2::256B [%0], [%1], 16;\n\t}"
            :
            : "r"(saddr), "l"(g), "r"((uint32_t)p)
            : "memory");
        if (MIRROR) {
            const bool m = p & ((wr & (RING_PIECES - 1)) == 0u);
            asm volatile(
                "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q cp.async.cg.shared.global.L2::256B [%0], [%1], 16;\n\t}"
                :
                : "r"(ring + RING_PIECES * 16), "l"(g), "r"((uint32_t)m)
                : "memory");
        }
        wr += p ? 1u : 0u;
    }
    // Position-window form of round1(): `pos` = first byte (from pbase) the lane has not shifted into data yet.
    __device__ __forceinline__ void round1_pos(uint32_t pos, uint32_t ring) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        const uint32_t occ = wr * 16 - pos;
        issue_if((occ <= RING_PIECES * 16 - 16) & (wr < npieces), ring);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // Once per output word, all lanes together: retire older groups, top the ring up (<= 2 pieces).
    __device__ __forceinline__ void round(RingFetch& f) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        uint32_t occ = wr * 16 - f.rd * 4;  // bytes requested and not yet consumed
        // No "ring ran dry" check is needed: wait_group 1 leaves only the newest group (<= 32 bytes, the
        // pieces furthest ahead) in flight; a word consumes <= 12 bytes on the fast path and the ring is
        // topped up to > 112 bytes every round, so the completed part never drops below ~64 bytes.  The
        // exact path (which may consume more) drains the ring first and re-attaches or resyncs after.
        const bool p1 = (occ <= RING_PIECES * 16 - 16) & (wr < npieces);
        issue_if(p1, f.ring);
        occ += p1 ? 16u : 0u;
        const bool p2 = (occ <= RING_PIECES * 16 - 16) & (wr < npieces);
        issue_if(p2, f.ring);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // One piece per call: enough for the fused word loop (a clean word consumes <= 12 bytes, an unclean
    // one refills the ring completely before it is re-decoded).
    __device__ __forceinline__ void round1(RingFetch& f) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        const uint32_t occ = wr * 16 - f.rd * 4;
        issue_if((occ <= RING_PIECES * 16 - 16) & (wr < npieces), f.ring);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // Before the exact out-of-line path: everything requested has landed, so that path can read the ring.
    __device__ __forceinline__ void drain() {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // After it (drain() was called before): the ring still holds every requested piece.
    __device__ __forceinline__ void after_exact(RingFetch& f) {
        if (RCB_LIKELY(f.rd * 4 < wr * 16))
            f.reload();
        else
            resync(f);
    }
    // Before an unclean word is re-decoded from its checkpoint (f.rd restored by the caller).
    __device__ __forceinline__ void redo_ready(RingFetch& f) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (RCB_LIKELY((int32_t)(wr * 16 - f.rd * 4) >= 64 || wr >= npieces))
            f.reload();
        else
            resync(f);
    }
    // (Re)start after the read position moved arbitrarily (initial fill, exact path past the ring).
    __device__ __forceinline__ void resync(RingFetch& f) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (f.rd * 4 >= wr * 16) wr = f.rd >> 2;  // consumed past everything requested
#pragma unroll 1
        // signed: right after a restart the read position sits up to 12 bytes inside piece wr
        while (((int32_t)(wr * 16 - f.rd * 4) <= (int32_t)(RING_PIECES * 16 - 16)) & (wr < npieces))
            issue_if(true, f.ring);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        f.reload();
    }
};
using RingFill = RingFillT<false>;

// ---------------------------------------------------------------------------
// Fetch for the exact out-of-line path: the current word is cached; a new word comes from the
// lane's ring while it is inside the requested window [ring_lo, ring_hi) (the caller drained the
// ring first), else from global memory (clamped to the readable stream).
// ---------------------------------------------------------------------------
struct GlobalFetch {
    const uint32_t* base;  // pbase as words
    uint32_t idx;          // words consumed; `cur` holds word idx
    uint32_t last;         // last readable word
    uint32_t ring;         // shared-space address of the lane's ring (0: none)
    uint32_t ring_lo, ring_hi;
    uint32_t cur;
    __device__ __forceinline__ GlobalFetch(const uint32_t* b, uint32_t i, uint32_t l, uint32_t r = 0, uint32_t rlo = 0,
                                           uint32_t rhi = 0)
        : base(b), idx(i), last(l), ring(r), ring_lo(rlo), ring_hi(rhi) {
        load();
    }
    __device__ __forceinline__ void load() {
        if (ring && idx >= ring_lo && idx < ring_hi)
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(cur) : "r"(ring + (idx & (RING_PIECES * 4 - 1)) * 4) : "memory");
        else
            cur = __ldg(base + (idx < last ? idx : last));
    }
    __device__ __forceinline__ uint32_t peek_be32() const { return bswap32(cur); }
    __device__ __forceinline__ void advance_if(bool p) {
        if (p) {
            idx++;
            load();
        }
    }
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

__device__ __forceinline__ float2 lds_f2(uint32_t saddr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
    return v;
}

// Lane state that travels by value through the exact (out-of-line) path.
struct DecLaneState {
    uint64_t lo, rg;
    uint32_t dh, dl, wh, wl, cnt;
    const uint32_t* base;
    uint32_t rd, last;
    uint32_t err, syms;  // syms: symbols decoded by the call, packed like the output word
    uint32_t ring, ring_hi;  // drained ring: words [ring_hi - 32, ring_hi) are in shared memory (ring 0: none)
};

// One symbol on the exact path: the reference's search in the product domain plus the
// generic renormalisation (literal loops when needed).
template <bool CHECKED, class Sink>
__device__ __forceinline__ uint32_t dec_symbol_exact(uint64_t& lo, uint64_t& rg, Sink& sink, uint32_t& err,
                                                     const uint2* tab, uint32_t K, const DivParams& div,
                                                     bool pow2) {
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

// n_syms symbols (<= 4) decoded exactly from a checkpoint (the reference's search + the generic
// renormalisation), symbols packed sym_bits apart.  Out of line: the generic kernels' table-miss path.
template <bool CHECKED>
__device__ __noinline__ DecLaneState dec_exact(DecLaneState s, const uint2* tab, uint32_t K, DivParams div,
                                               uint32_t pow2, uint32_t n_syms, uint32_t sym_bits) {
    GlobalFetch gf(s.base, s.rd, s.last, s.ring, s.ring_hi > RING_PIECES * 4 ? s.ring_hi - RING_PIECES * 4 : 0u,
                   s.ring_hi);
    DecSink<GlobalFetch> sink(gf);
    sink.dh = s.dh;
    sink.dl = s.dl;
    sink.wh = s.wh;
    sink.wl = s.wl;
    sink.cnt = s.cnt;
    uint32_t acc = 0;
#pragma unroll 1
    for (uint32_t b = 0; b < n_syms; b++) {
        const uint32_t sym = dec_symbol_exact<CHECKED>(s.lo, s.rg, sink, s.err, tab, K, div, pow2 != 0);
        acc |= sym << (sym_bits * b);
    }
    s.syms = acc;
    s.dh = sink.dh;
    s.dl = sink.dl;
    s.wh = sink.wh;
    s.wl = sink.wl;
    s.cnt = sink.cnt;
    s.rd = sink.f.idx;
    return s;
}

// FMODE: -1 = generic per-symbol loop (bucket LUT + exact search); FUSE_BIG / FUSE_POW2 / FUSE_GEN = the fused
// word-speculative loop over the fat LUT for a power-of-two total >= 2^24 / any power of two / any total
// (FUSE_GEN: divide-free rpt_next from the candidates' reciprocal constants, rcb_core.cuh).
// The fused loops address the code bytes by position: `data` is carried by funnel shifts, the bytes that
// follow it are re-read from the lane's ring once per output word (four 32-bit loads + three byte permutes)
// and shifted along without any refill bookkeeping (14 shifts + ~8 other integer instructions per word where
// a counted refill per symbol took ~56).
// TP (fused loops only) = tuned for throughput instead of latency: with several warps per scheduler (restart
// points, many chunks) the loop is bound by the integer pipe and by shared-memory / LSU wavefronts, not by the
// dependency chain -- so the bucket estimate takes its one reciprocal AFTER the symbol is known (no second
// table: -6 wavefronts per symbol) and output words leave 16 bytes at a time (-24 wavefronts per word).
template <typename SYM, bool SHARED, bool POW2, bool CHECKED, int FMODE, bool TP = false>
__global__ void __launch_bounds__(TP ? 640 : 512, 1) decode_kernel(DecodeArgs a) {
    constexpr bool FUSED = FMODE >= 0;
    constexpr bool WIN = FUSED;  // mirrored ring
    static_assert(!TP || FUSED, "TP belongs to the fused loops");
    constexpr bool CSM = FMODE == FUSE_GEN;     // general total, per-candidate reciprocal constants (third array)
    constexpr bool M2M = FMODE == FUSE_GEN_M2;  // general total >= 2^25, one table-wide constant
    constexpr bool GENM = CSM || M2M;
    constexpr int MODE = !FUSED ? FUSE_BIG : (GENM ? FUSE_GEN : FMODE);
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ ModelHdr s_hdr;
    // shared layout: rings[blockDim.x][RING_STRIDE] | LutEntry[nb or 4096] | (FUSED) float2 rc[4096]
    //                | (FUSE_GEN) uint4 cs[4096] | uint2[K]
    // (entries and reciprocals in separate arrays: 16- and 8-byte strides spread random lookups over all
    // banks; one 32-byte record per bucket doubled the bank conflicts once several warps share an SM)
    uint8_t* s_ring = s_raw;
    LutEntry* s_lut = reinterpret_cast<LutEntry*>(s_raw + (size_t)blockDim.x * RING_STRIDE);
    uint2* s_tab = nullptr;
    if (SHARED) {
        if (threadIdx.x == 0) s_hdr = a.hdrs[0];
        __syncthreads();
        const uint32_t nb = (s_hdr.flags & MODEL_REGULAR) ? s_hdr.nb : 0u;
        const uint32_t nb_pad = FUSED ? 4096u : nb;  // FUSED indexes any of 4096 entries
        float2* s_rc = reinterpret_cast<float2*>(s_lut + nb_pad);
        uint4* s_cs = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(s_rc) + (FUSED ? 4096u * sizeof(float2) : 0u));
        s_tab = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(s_cs) + (CSM ? 4096u * sizeof(uint4) : 0u));
        const uint4* gl = reinterpret_cast<const uint4*>(a.lut);
        uint4* sl = reinterpret_cast<uint4*>(s_lut);
        const uint32_t total = s_hdr.div.total;
        const uint32_t sr = GENM ? 0u : fused_sr(make_fused(s_hdr.div));
        for (uint32_t i = threadIdx.x; i < nb_pad; i += blockDim.x) {
            const uint4 e = i < nb ? gl[i] : make_uint4(total, total, total, 0u);  // empty interval: never verifies
            sl[i] = e;
            if (FUSED && !TP)  // reciprocals of the candidates' frequencies for the shift-free estimate (rcb_core.cuh)
                s_rc[i] = make_float2(lut_rc16(e.y - e.x, s_hdr.lut_scale, sr), lut_rc16(e.z - e.y, s_hdr.lut_scale, sr));
            if (CSM) s_cs[i] = i < nb ? a.lut_cs[i] : make_uint4(0u, 0u, 0u, 0u);
        }
        for (uint32_t i = threadIdx.x; i < a.K; i += blockDim.x) s_tab[i] = a.tabs[i];
        __syncthreads();
    }
    const uint32_t parts = a.parts;
    const LanePart lp = lane_part((uint64_t)blockIdx.x * blockDim.x + threadIdx.x, parts, a.n_chunks, a.n_syms,
                                  a.chunk_syms, a.restart_syms);
    const unsigned live = __ballot_sync(0xFFFFFFFFu, lp.has);  // lanes of this warp that have symbols to decode
    if (!lp.has) return;
    const uint64_t chunk = lp.chunk, first = lp.first, chunk_cnt = lp.chunk_cnt;
    // this lane's part of the chunk: the whole chunk, one of `parts` pieces between restart points, or (host
    // pipeline) this launch's segment
    const bool seg_load = a.seg.state && a.seg.load, seg_save = a.seg.state && a.seg.save;
    uint64_t seg_begin = a.seg.state ? (a.seg.first < chunk_cnt ? a.seg.first : chunk_cnt) : 0;
    uint64_t seg_end =
        seg_save ? (a.seg.first + a.seg.syms < chunk_cnt ? a.seg.first + a.seg.syms : chunk_cnt) : chunk_cnt;
    if (parts > 1) {
        seg_begin = (uint64_t)lp.part * a.restart_syms;
        seg_end = seg_begin + a.restart_syms < chunk_cnt ? seg_begin + a.restart_syms : chunk_cnt;
    }
    uint64_t cnt = seg_end - seg_begin;
    SYM* dst = reinterpret_cast<SYM*>(a.out) + first + seg_begin;

    const uint2* tab = SHARED ? s_tab : a.tabs + chunk * a.K;
    ModelHdr hdr = SHARED ? s_hdr : a.hdrs[chunk];
    const DivParams div = hdr.div;
    const bool pow2 = SHARED ? POW2 : ((hdr.flags & MODEL_POW2) != 0);
    const bool use_lut = SHARED && (hdr.flags & MODEL_REGULAR);
    const float lut_scale = hdr.lut_scale;
    const float max_bucket = (float)(hdr.nb ? hdr.nb - 1 : 0);
    const uint32_t K = a.K;

    const uint64_t total_bytes = a.offsets[a.n_chunks];
    uint64_t off0 = a.offsets[chunk], off1 = a.offsets[chunk + 1];
    // Offsets come from the caller (a corrupt frame, a damaged index): a chunk is at least the 8 bytes
    // Decoder::new pops (src/decoder.rs:14-23) and lies inside the stream.  A lane whose offsets break
    // that decodes nothing and reports ST_TRUNCATED; it stays in the warp (the loops below vote), reads
    // no global memory (no piece is requested, no symbol decoded) and writes nothing.
    bool offsets_ok = off0 <= off1 && off1 - off0 >= 8 && off1 <= total_bytes;
    // a lane that enters at a restart point starts its 8-byte window `pos` bytes into the chunk; the record is
    // caller data like the offsets: it must leave room for the window, else the lane decodes nothing (ST_RESTART)
    Restart rp{0ull, ~0ull, 0u, 0u};
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
    const uint64_t start = off0 + rp.pos;                            // first byte of the lane's window
    const uint64_t pb = start & ~15ull;                              // piece base (byte offset)
    const uint64_t readable = ((total_bytes + 15) & ~15ull) - pb;    // rcb200.h: readable to the next 16
    const uint32_t skip = (uint32_t)(start & 3u);
    const uint32_t rd0 = (uint32_t)((start - pb) >> 2);

    RingFillT<WIN> fill;
    fill.pbase = a.stream + pb;
    fill.wr = 0;
    fill.npieces = !offsets_ok ? 0u : readable > (0xFFFFFFF0ull << 4) ? 0xFFFFFFF0u : (uint32_t)(readable >> 4);
    if (!offsets_ok) cnt = 0;
    RingFetch rf;
    rf.ring = (uint32_t)__cvta_generic_to_shared(s_ring + (size_t)threadIdx.x * RING_STRIDE);
    rf.rd = seg_load ? a.seg.state[chunk].rd : rd0;
    if (WIN && seg_load) {
        // the position window re-reads the bytes a resumed lane still held in registers: fill the ring from
        // the word of the first byte that is not in data yet
        const DecResume st = a.seg.state[chunk];
        rf.rd = (st.rd * 4u - (st.cnt >> 3)) >> 2;
    }
    rf.cur = 0;
    const uint32_t last_word = fill.npieces ? fill.npieces * 4 - 1 : 0u;

    // L2 prefetch of the next 2 KiB of this lane's code bytes; afterwards every ring piece carries a
    // 256-byte L2 prefetch hint (no per-word prefetch branch in the loops)
    constexpr uint32_t PF_WORDS = 256;
    uint32_t pf_next = rf.rd & ~3u;  // next granule (in words from pbase) to request
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
    fill.resync(rf);  // fill the ring from the read position, wait, load the current word
    DecSink<RingFetch> sink(rf);

    uint64_t lo = rp.lo, rg = rp.rg;  // RangeCoder::new (src/range_coder.rs:13-20) unless a restart point says otherwise
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
        sink.f.rd = st.rd;  // with (wh, wl, cnt): the position st.rd * 4 - cnt / 8 the fused loop starts from
        err = st.err;
    } else {
        sink.prime(skip);  // src/decoder.rs:14-23
    }
    if (!offsets_ok) err = restart_ok ? ST_TRUNCATED : ST_RESTART;  // src/decoder.rs:33: pop_front on an empty buffer

    constexpr uint32_t PER = 4 / sizeof(SYM);  // symbols per 32-bit store
    constexpr uint32_t SYM_BITS = 8 * sizeof(SYM);
    auto snapshot = [&]() -> DecLaneState {
        return DecLaneState{lo, rg, sink.dh, sink.dl, sink.wh, sink.wl, sink.cnt,
                            reinterpret_cast<const uint32_t*>(fill.pbase), sink.f.rd, last_word, err, 0u,
                            sink.f.ring, fill.wr * 4};
    };
    auto restore = [&](const DecLaneState& st) {  // after the exact path: adopt its state, re-attach the ring
        lo = st.lo;
        rg = st.rg;
        sink.dh = st.dh;
        sink.dl = st.dl;
        sink.wh = st.wh;
        sink.wl = st.wl;
        sink.cnt = st.cnt;
        sink.f.rd = st.rd;
        err = st.err;
        fill.after_exact(sink.f);
    };

    uint64_t done = 0;
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst);
    // TP stores 16 bytes at a time (hosts pick it only for 16-byte aligned parts; a lane that is not decodes its
    // symbols one by one below)
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & (TP ? 15u : 3u)) == 0;
    const uint64_t nw = aligned ? cnt / PER : 0;

    if constexpr (FUSED) {
        const FusedParams fp = make_fused(div);
        uint64_t rpt = fused_rpt<MODE>(rg, fp);
        const uint32_t sr = GENM ? 0u : fused_sr(fp);
        auto q_of = [&](uint64_t r) -> float { return GENM ? lut_q_gen(r) : lut_q(r, sr); };  // 1 / float(rpt >> sr)
        auto range_of = [&](uint64_t r) -> uint64_t { return GENM ? r * (uint64_t)div.total : r << fp.s; };
        const Recip2 k2 = M2M ? make_recip2(div.total) : Recip2{0ull, 0u};
        float q = q_of(rpt);
        float bf = lut_bf16_init(sink.data() - lo, rg, lut_scale);     // byte offset of the next entry
        const uint32_t lut_saddr = (uint32_t)__cvta_generic_to_shared(s_lut);
        const uint32_t rc_saddr = lut_saddr + 4096u * (uint32_t)sizeof(LutEntry);
        const uint32_t cs_saddr = rc_saddr + 4096u * (uint32_t)sizeof(float2);
        const uint32_t stage_saddr = rc_saddr;  // TP has no reciprocal table: its 32 KiB hold the staged output words

        // ---- position window.  (dh:dl) = data, pos = first byte (from pbase) not yet shifted into it.
        uint32_t dh = sink.dh, dl = sink.dl;
        uint32_t pos = sink.f.rd * 4u - (sink.cnt >> 3);
        const uint32_t ring = sink.f.ring;
        // the generic sink (exact re-decode, tail symbols) re-attached at `pos`: like prime() without the 8
        // bytes that are already in data; the word at pos / 4 has landed (callers make sure)
        auto attach = [&]() {
            sink.dh = dh;
            sink.dl = dl;
            sink.f.rd = pos >> 2;
            sink.f.reload();
            const uint32_t skipb = pos & 3u;
            const uint32_t first = sink.f.peek_be32();
            sink.f.advance_if(true);
            sink.wh = first << (8u * skipb);
            sink.wl = 0;
            sink.cnt = 32u - 8u * skipb;
            sink.refill();
        };
        auto detach = [&]() {
            dh = sink.dh;
            dl = sink.dl;
            pos = sink.f.rd * 4u - (sink.cnt >> 3);
        };
        struct WordChk {  // lane state at the start of a word: what the re-decode restarts from
            uint64_t lo, rpt;
            uint32_t dh, dl, pos;
        };
        // Four symbols, straight-line and speculative; `bad` = some symbol needs the exact path.
        auto decode_word = [&](uint32_t& acc, bool& bad) {
            acc = 0;
            bad = false;
            // the 3 * PER bytes that follow data, big-endian, from the (mirrored) ring: no address wrap
            const uint32_t a0 = ring + (pos & (RING_PIECES * 16u - 4u));
            const uint32_t sel = (pos & 3u) * 0x1111u + 0x0123u;  // bytes o .. o+3 of a word pair, reversed
            uint32_t x0, x1, x2, x3 = 0;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x0) : "r"(a0) : "memory");
            asm volatile("ld.shared.b32 %0, [%1+4];" : "=r"(x1) : "r"(a0) : "memory");
            asm volatile("ld.shared.b32 %0, [%1+8];" : "=r"(x2) : "r"(a0) : "memory");
            if (PER > 2) asm volatile("ld.shared.b32 %0, [%1+12];" : "=r"(x3) : "r"(a0) : "memory");
            uint32_t w0 = __byte_perm(x0, x1, sel), w1 = __byte_perm(x1, x2, sel);
            uint32_t w2 = PER > 2 ? __byte_perm(x2, x3, sel) : 0u;
            uint32_t tot = 0;
#pragma unroll
            for (uint32_t b = 0; b < PER; b++) {
                const uint64_t data = ((uint64_t)dh << 32) | dl;
                const uint32_t off = lut_offset16(bf);
                const LutEntry e = lds_lut(lut_saddr + off);
                float2 rc = make_float2(0.f, 0.f);
                if (!TP) rc = lds_f2(rc_saddr + (off >> 1));
                FusedDec r;
                if constexpr (CSM) {
                    const LutEntry k = lds_lut(cs_saddr + off);  // {csA lo, csA hi, csB lo, csB hi}
                    r = fused_decode_step_cs<true>(lo, rpt, data, e, ((uint64_t)k.cumB << 32) | k.cumA,
                                                   ((uint64_t)k.syms << 32) | k.cumC);
                } else if constexpr (M2M) {
                    r = fused_decode_step_m2<true>(lo, rpt, data, e, k2);
                } else {
                    r = fused_decode_step<MODE, true>(lo, rpt, data, e, fp);
                }
                if (TP) {
                    // next entry from the unshifted residue and the unshifted new range (both sides of
                    // (data' - lower') / range' carry the same shift): one reciprocal after the symbol is
                    // known, no second table -- shared-memory wavefronts, not latency, bound this flavour
                    bf = lut_bf16_init(data - r.nlo, r.rgp, lut_scale);
                } else {
                    // ... with the reciprocal taken BEFORE the symbol is known (1 / rpt while the table load is
                    // in flight, const / c from a second table): shortest dependency chain (rcb_core.cuh)
                    bf = u64_to_float(data - r.nlo) * (q * (r.takeB ? rc.y : rc.x));
                }
                // shift the window: a symbol takes <= 3 bytes on the fast path, so of the bytes behind data
                // only 3 * (symbols still to come) matter -- one register fewer per symbol
                dh = funnel_l(dl, dh, r.sh);
                dl = funnel_l(w0, dl, r.sh);
                if (b + 1 < PER) w0 = funnel_l(w1, w0, r.sh);
                if (b + 2 < PER) w1 = funnel_l(w2, w1, r.sh);
                if (b + 3 < PER) w2 <<= r.sh;
                tot += r.sh;
                lo = r.nlo << r.sh;
                rpt = r.nrpt;
                if (!TP) q = q_of(rpt);
                acc |= r.sym << (SYM_BITS * b);
                bad |= !r.ok;
            }
            pos += tot >> 3;
        };
        // The word was not clean for this lane: back to the checkpoint (register moves) and through the
        // word one symbol at a time on the generic sink -- the table step where it verifies, the reference's
        // literal loops / search where it does not.
        auto redo_word = [&](const WordChk& chk) -> uint32_t {
            lo = chk.lo;
            rpt = chk.rpt;
            dh = chk.dh;
            dl = chk.dl;
            pos = chk.pos;
            sink.f.rd = pos >> 2;
            // Everything requested has landed (the newest piece was issued a whole word ago: no wait in
            // practice).  The re-decode reads at most 4 x 14 bytes beyond the checkpoint, so with >= 64 bytes
            // in the ring it needs no new piece -- and no round trip to memory; otherwise (a run of unclean
            // words, the start of a chunk) fill the ring to the brim first.
            fill.redo_ready(sink.f);
            attach();
            uint32_t acc = 0;
#pragma unroll 1
            for (uint32_t b = 0; b < PER; b++) {
                const uint64_t data = sink.data();
                const uint32_t off = lut_offset16(lut_bf16_init(data - lo, range_of(rpt), lut_scale));
                const FusedDec r = fused_decode_step<MODE>(lo, rpt, data, lds_lut(lut_saddr + off), fp);
                uint32_t sym = r.sym;
                if (RCB_LIKELY(r.ok)) {
                    sink.put(0u, r.sh);
                    lo = r.nlo << r.sh;
                    rpt = r.nrpt;
                } else {
                    uint64_t l2 = r.nlo, g2 = r.rgp;
                    if (!r.inside) {  // table miss: the reference's search (examples/sample_impl.rs:27-45)
                        sym = find_index_exact(data - lo, rpt, K, [&](uint32_t j) { return tab[j].x; });
                        const uint2 t = tab[sym];
                        l2 = lo + rpt * (uint64_t)t.x;  // src/decoder.rs:42-50
                        g2 = rpt * (uint64_t)t.y;
                    }
                    renorm_slow<CHECKED>(l2, g2, sink, err);  // src/range_coder.rs:83-89, bytes from the ring
                    lo = l2;
                    rpt = fused_rpt<MODE>(g2, fp);
                }
                acc |= sym << (SYM_BITS * b);
            }
            rg = range_of(rpt);
            q = q_of(rpt);
            bf = lut_bf16_init(sink.data() - lo, rg, lut_scale);
            detach();
            // the next clean word reads 16 bytes from the word of `pos`: top the ring up if the exact path ate
            // into that margin (the per-word round adds at most one piece)
            if ((int32_t)(fill.wr * 16u - pos) < 48 && fill.wr < fill.npieces) {
                sink.f.rd = pos >> 2;
                fill.resync(sink.f);
            }
            return acc;
        };
        // Output (TP): a lane's 32-bit words go to global memory 16 bytes at a time.  Every lane writes its own
        // part of the output, so a warp's 4-byte stores are 32 separate sectors = 32 wavefronts of the LSU data
        // pipe per word; three words wait in shared memory ([slot][thread]: conflict-free) and leave with the
        // fourth as one 16-byte store.  (hosts launch TP kernels only when every part starts 16-byte aligned)
        const uint32_t stage = stage_saddr + threadIdx.x * 4u, stage_pitch = blockDim.x * 4u;
        auto put_word = [&](uint64_t w, uint32_t acc) {
            if (TP) {
                const uint32_t slot = (uint32_t)w & 3u;
                asm volatile(
                    "{\n\t.reg .pred q;\n\t.reg .b32 a, b, c;\n\t"
                    "setp.eq.b32 q, %0, 3;\n\t"
                    "@!q st.shared.b32 [%1], %2;\n\t"
                    "@q ld.shared.b32 a, [%3];\n\t"
                    "@q ld.shared.b32 b, [%4];\n\t"
                    "@q ld.shared.b32 c, [%5];\n\t"
                    "@q st.global.v4.b32 [%6], {a, b, c, %2};\n\t}"
                    :
                    : "r"(slot), "r"(stage + slot * stage_pitch), "r"(acc), "r"(stage), "r"(stage + stage_pitch),
                      "r"(stage + 2u * stage_pitch), "l"(dw + (w & ~3ull))
                    : "memory");
            } else {
                dw[w] = acc;
            }
        };
        // a corrected word: into its slot while the group is still staged, else straight to global memory
        auto fix_word = [&](uint64_t w, uint32_t acc) {
            if (TP && ((uint32_t)w & 3u) != 3u)
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(stage + ((uint32_t)w & 3u) * stage_pitch), "r"(acc) : "memory");
            else
                dw[w] = acc;
        };
        // Main loop: the words every live lane of the warp has (all of them, except in a warp holding the
        // ragged last chunk).  Its only branch in the common case is the back edge, and that branch is
        // warp-uniform (vote): a taken branch costs ~35 cycles with one warp per scheduler, and a per-lane
        // exit would park the lane at the reconvergence point until its whole warp left the loop.
        const uint32_t nw_warp = __reduce_min_sync(live, (uint32_t)(nw < 0xFFFFFFFFull ? nw : 0xFFFFFFFFull));
        uint64_t i = 0;
        while (i < nw_warp) {
            WordChk chk;
            uint32_t acc;
            bool bad, leave;
#pragma unroll 1
            do {
                fill.round1_pos(pos, ring);
                chk = WordChk{lo, rpt, dh, dl, pos};
                decode_word(acc, bad);
                put_word(i, acc);  // speculative as well: rewritten below when the word was not clean
                ++i;
                leave = __any_sync(live, bad) | (i >= nw_warp);
            } while (!leave);
            if (bad) fix_word(i - 1, redo_word(chk));
        }
        if (TP) {  // words of an unfinished group
#pragma unroll 1
            for (uint64_t w = i & ~3ull; w < i; w++) {
                uint32_t v;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(stage + ((uint32_t)w & 3u) * stage_pitch) : "memory");
                dw[w] = v;
            }
        }
#pragma unroll 1
        for (; i < nw; i++) {  // ragged warp only
            fill.round1_pos(pos, ring);
            const WordChk chk{lo, rpt, dh, dl, pos};
            uint32_t acc;
            bool bad;
            decode_word(acc, bad);
            if (RCB_UNLIKELY(bad)) acc = redo_word(chk);
            dw[i] = acc;
        }
        // back to the generic sink for the tail symbols and the final accounting
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        attach();
        done = nw * PER;
        rg = range_of(rpt);
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
            fill.drain();
            const DecLaneState r = dec_exact<CHECKED>(snapshot(), tab, K, div, pow2 ? 1u : 0u, 1u, 0u);
            restore(r);
            return r.syms;
        }
        return dec_symbol_exact<CHECKED>(lo, rg, sink, err, tab, K, div, pow2);
    };

    if (!FUSED) {
#pragma unroll 1
        for (uint64_t i = 0; i < nw; i++) {
            fill.round(sink.f);
            uint32_t acc = 0;
#pragma unroll
            for (uint32_t b = 0; b < PER; b++) acc |= step() << (SYM_BITS * b);
            dw[i] = acc;
        }
        done = nw * PER;
    }
#pragma unroll 1
    for (uint64_t i = done; i < cnt; i++) {
        fill.round(sink.f);
        dst[i] = (SYM)step();
    }

    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (seg_save) {  // the chunk continues in the next launch
        a.seg.state[chunk] = DecResume{lo, rg, sink.dh, sink.dl, sink.wh, sink.wl, sink.cnt, sink.f.rd, err, 0u};
        return;
    }
    const uint32_t used = sink.used(sink.f.rd - rd0, skip);  // bytes shifted into data by this lane (8 of them priming)
    if (parts > 1) {
        // several lanes per chunk: the status word was zeroed by the host, a lane reports only an error.  A lane
        // that ends at a restart point must have arrived at that record's state (a damaged stream or record
        // shows up here); the chunk's last lane checks the length like a whole-chunk lane.
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
