// rcb_adaptive.cuh -- SURVEY 8 f4: chunk-parallel coding under a table that changes after every symbol.
//
// The reference takes `&T: PModel` on every call (src/encoder.rs:24, src/decoder.rs:38), so its caller may
// update the table between symbols; it ships no such model.  The build defines the textbook one (DESIGN.md
// section 5):
//     c[i] = 1 for i < K;  after coding symbol s with the table as it was:  c[s] += inc;
//     when sum(c) would pass `limit`: c[i] = (c[i] + 1) >> 1 for all i;  cum = exclusive prefix sum, total = sum
// and every chunk restarts both the coder and the table.  Per symbol the coder needs c[s], cum[s] and total
// of a table that just changed, so a dense cum[] (O(K) to maintain) is replaced by a Fenwick tree per lane in
// shared memory: prefix sum, point update and the search "largest s with cum[s] <= rfreq"
// (examples/sample_impl.rs:27-45) are log2(K) dependent 2-byte loads each.  total is not constant, so
// range / total (src/range_coder.rs:38-40) is a true 64-bit division here; the decoder gets rfreq =
// (data - lower) / rpt from a float estimate corrected by exact products.
//
// First correct path, not tuned: one lane per chunk, u16 counts (limit + inc <= 65535), K <= 4096.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcb_core.cuh"
#include "rcb_decode.cuh"
#include "rcb_encode.cuh"

namespace rcb {

struct AdaptiveArgs {
    uint32_t K, inc, limit;
    uint32_t lanes_per_block;
    uint32_t pitch_words;  // 32-bit words per lane: counts u16[K] then tree u16[K+1], odd word pitch
    uint64_t n_syms, chunk_syms, n_chunks;
    // encode
    const void* syms;
    uint8_t* staging;
    uint64_t pitch;
    uint32_t* lens;
    // decode
    const uint8_t* stream;
    const uint64_t* offsets;
    void* out;
    uint32_t* status;
};

// One lane's table: counts cnt[0..K) and a 1-based Fenwick tree over them, both u16, in shared memory.
struct AdaptiveTable {
    uint16_t* cnt;
    uint16_t* tree;  // tree[1..K]
    uint32_t K, top;  // top = largest power of two <= K
    uint32_t total, inc, limit;

    __device__ __forceinline__ void init(uint16_t* base, uint32_t K_, uint32_t inc_, uint32_t limit_) {
        cnt = base;
        tree = base + K_;  // tree[0] unused
        K = K_;
        inc = inc_;
        limit = limit_;
        top = 1u << (31u - (uint32_t)__clz((int)K_));
        for (uint32_t i = 0; i < K; i++) cnt[i] = 1;
        rebuild();
    }
    __device__ __forceinline__ void rebuild() {
        uint32_t sum = 0;
        for (uint32_t i = 1; i <= K; i++) {
            const uint32_t v = cnt[i - 1];
            tree[i] = (uint16_t)v;
            sum += v;
        }
        for (uint32_t i = 1; i <= K; i++) {
            const uint32_t j = i + (i & (0u - i));
            if (j <= K) tree[j] = (uint16_t)(tree[j] + tree[i]);
        }
        total = sum;
    }
    // cum_freq(s): sum of the counts below s
    __device__ __forceinline__ uint32_t prefix(uint32_t s) const {
        uint32_t sum = 0;
        for (uint32_t i = s; i > 0; i -= i & (0u - i)) sum += tree[i];
        return sum;
    }
    // the caller's update after symbol s was coded
    __device__ __forceinline__ void update(uint32_t s) {
        if (total + inc > limit) {
            cnt[s] = (uint16_t)(cnt[s] + inc);
            for (uint32_t i = 0; i < K; i++) cnt[i] = (uint16_t)((cnt[i] + 1u) >> 1);
            rebuild();
            return;
        }
        cnt[s] = (uint16_t)(cnt[s] + inc);
        for (uint32_t i = s + 1; i <= K; i += i & (0u - i)) tree[i] = (uint16_t)(tree[i] + inc);
        total += inc;
    }
    // examples/sample_impl.rs:27-45: the largest s with cum[s] <= rfreq, clamped to K-1; cum[s] returned too
    __device__ __forceinline__ uint32_t find(uint32_t rfreq, uint32_t& cum) const {
        uint32_t pos = 0, rem = rfreq;
        for (uint32_t step = top; step > 0; step >>= 1) {
            const uint32_t np = pos + step;
            if (np <= K) {
                const uint32_t t = tree[np];
                if (t <= rem) {
                    pos = np;
                    rem -= t;
                }
            }
        }
        if (pos >= K) {  // rfreq >= total: the reference's search stops at the last symbol
            pos = K - 1;
            cum = prefix(pos);
            return pos;
        }
        cum = rfreq - rem;
        return pos;
    }
};

template <typename SYM>
__global__ void __launch_bounds__(512, 1) adaptive_encode_kernel(AdaptiveArgs a) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    if (threadIdx.x >= a.lanes_per_block) return;
    const uint64_t chunk = (uint64_t)blockIdx.x * a.lanes_per_block + threadIdx.x;
    if (chunk >= a.n_chunks) return;
    const uint64_t first = chunk * a.chunk_syms;
    const uint64_t cnt = (a.n_syms - first < a.chunk_syms) ? (a.n_syms - first) : a.chunk_syms;
    const SYM* src = reinterpret_cast<const SYM*>(a.syms) + first;
    AdaptiveTable t;
    t.init(reinterpret_cast<uint16_t*>(s_raw) + (size_t)threadIdx.x * a.pitch_words * 2, a.K, a.inc, a.limit);

    uint64_t lo = 0, rg = ~0ull;  // src/range_coder.rs:13-20
    uint32_t err = 0;
    RowStore rs{a.staging + chunk * a.pitch};
    EncSink<RowStore, true> sink(rs, (uint32_t)a.pitch);
    auto code = [&](uint32_t s) {
        if (s >= a.K) {
            if (!err) err = ST_SYMBOL_RANGE;
            s = 0;
        }
        const uint32_t c = t.cnt[s], cum = t.prefix(s);
        const uint64_t rpt = rg / (uint64_t)t.total;  // src/range_coder.rs:62 (total changes per symbol)
        rg = rpt * (uint64_t)c;                        // :65
        lo = lo + rpt * (uint64_t)cum;                 // :68-81 (cum + c <= total: cannot wrap)
        renorm<false>(lo, rg, sink, err);              // :83-89
        t.update(s);
    };
    constexpr uint32_t PER = 16 / sizeof(SYM);
    uint64_t done = 0;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const uint4* v = reinterpret_cast<const uint4*>(src);
        const uint64_t nvec = cnt / PER;
        uint4 cur = nvec ? ldg_stream_v4(v) : make_uint4(0, 0, 0, 0);
        for (uint64_t i = 0; i < nvec; i++) {
            const uint4 nxt = i + 1 < nvec ? ldg_stream_v4(v + i + 1) : make_uint4(0, 0, 0, 0);
            const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (sizeof(SYM) == 1) {
#pragma unroll
                    for (int b = 0; b < 4; b++) code((w[j] >> (8 * b)) & 0xFFu);
                } else {
#pragma unroll
                    for (int b = 0; b < 2; b++) code((w[j] >> (16 * b)) & 0xFFFFu);
                }
            }
            cur = nxt;
        }
        done = nvec * PER;
    }
    for (uint64_t i = done; i < cnt; i++) code((uint32_t)src[i]);
    const uint32_t len = sink.finish(lo);  // src/encoder.rs:40-46
    if (!err && sink.overflowed()) err = ST_OUT_CAPACITY;
    a.lens[chunk] = len;
    a.status[chunk] = err;
}

template <typename SYM>
__global__ void __launch_bounds__(512, 1) adaptive_decode_kernel(AdaptiveArgs a) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    // shared layout: rings[blockDim.x][RING_STRIDE] | tables[lanes][pitch_words]
    uint8_t* s_ring = s_raw;
    uint16_t* s_tabs = reinterpret_cast<uint16_t*>(s_raw + (size_t)blockDim.x * RING_STRIDE);
    if (threadIdx.x >= a.lanes_per_block) return;
    const uint64_t chunk = (uint64_t)blockIdx.x * a.lanes_per_block + threadIdx.x;
    if (chunk >= a.n_chunks) return;
    const uint64_t first = chunk * a.chunk_syms;
    uint64_t cnt = (a.n_syms - first < a.chunk_syms) ? (a.n_syms - first) : a.chunk_syms;
    SYM* dst = reinterpret_cast<SYM*>(a.out) + first;
    AdaptiveTable t;
    t.init(s_tabs + (size_t)threadIdx.x * a.pitch_words * 2, a.K, a.inc, a.limit);

    const uint64_t total_bytes = a.offsets[a.n_chunks];
    uint64_t off0 = a.offsets[chunk], off1 = a.offsets[chunk + 1];
    const bool offsets_ok = off0 <= off1 && off1 - off0 >= 8 && off1 <= total_bytes;  // see decode_kernel
    if (!offsets_ok) {
        off0 = off1 = 0;
        cnt = 0;
    }
    const uint64_t pb = off0 & ~15ull;
    const uint64_t readable = ((total_bytes + 15) & ~15ull) - pb;
    const uint32_t skip = (uint32_t)(off0 & 3u);
    const uint32_t rd0 = (uint32_t)((off0 - pb) >> 2);
    RingFill fill;
    fill.pbase = a.stream + pb;
    fill.wr = 0;
    fill.npieces = !offsets_ok ? 0u : readable > (0xFFFFFFF0ull << 4) ? 0xFFFFFFF0u : (uint32_t)(readable >> 4);
    RingFetch rf;
    rf.ring = (uint32_t)__cvta_generic_to_shared(s_ring + (size_t)threadIdx.x * RING_STRIDE);
    rf.rd = rd0;
    rf.cur = 0;
    fill.resync(rf);
    DecSink<RingFetch> sink(rf);
    sink.prime(skip);  // src/decoder.rs:14-23

    uint64_t lo = 0, rg = ~0ull;
    uint32_t err = offsets_ok ? 0u : (uint32_t)ST_TRUNCATED;
    auto step = [&]() -> uint32_t {
        const uint32_t total = t.total;
        const uint64_t rpt = rg / (uint64_t)total;  // src/range_coder.rs:62
        const uint64_t d = sink.data() - lo;        // examples/sample_impl.rs:29
        // rfreq = d / rpt, needed only up to `total`: float estimate, then exact in the product domain
        float est = u64_to_float(d) * fast_rcp(u64_to_float(rpt));
        uint32_t r = est >= (float)total ? total : (uint32_t)est;
#pragma unroll 1
        while (r > 0 && rpt * (uint64_t)r > d) r--;
#pragma unroll 1
        while (r < total && rpt * (uint64_t)(r + 1u) <= d) r++;
        uint32_t cum;
        const uint32_t sym = t.find(r, cum);
        const uint32_t c = t.cnt[sym];
        lo = lo + rpt * (uint64_t)cum;  // src/decoder.rs:42-50
        rg = rpt * (uint64_t)c;
        renorm<false>(lo, rg, sink, err);  // consumes as many bytes as the encoder emitted (:52)
        t.update(sym);
        return sym;
    };
    constexpr uint32_t PER = 4 / sizeof(SYM);
    constexpr uint32_t SYM_BITS = 8 * sizeof(SYM);
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 3u) == 0;
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst);
    const uint64_t nw = aligned ? cnt / PER : 0;
#pragma unroll 1
    for (uint64_t i = 0; i < nw; i++) {
        fill.round(sink.f);
        uint32_t acc = 0;
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) acc |= step() << (SYM_BITS * k);
        dw[i] = acc;
    }
#pragma unroll 1
    for (uint64_t i = nw * PER; i < cnt; i++) {
        fill.round(sink.f);
        dst[i] = (SYM)step();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const uint32_t used = sink.used(sink.f.rd - rd0, skip);
    if (!err && (uint64_t)used > off1 - off0) err = ST_TRUNCATED;  // src/decoder.rs:33
    a.status[chunk] = err;
}

}  // namespace rcb
