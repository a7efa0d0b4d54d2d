// rcb_stream.cuh -- single-lane "continued" coder: the reference's per-symbol
// API (Encoder::encode / Decoder::decode, src/encoder.rs:24-37, src/decoder.rs:38-54)
// keeps its state between calls; these kernels take that state in, code a slice
// of symbols on one GPU lane and hand the state back.  They exist so the host
// mirror of the reference API (include/rcb200.hpp) computes everything --
// including encode()'s return value, the number of bytes a symbol produced -- on
// the device; throughput comes from the chunk-parallel kernels, not from here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcb_core.cuh"

namespace rcb {

struct StreamState {      // mirrors rcb_stream_state (include/rcb200.h)
    uint64_t lower_bound; // src/range_coder.rs:9
    uint64_t range;       // src/range_coder.rs:11
    uint64_t data;        // src/decoder.rs:9
    uint64_t consumed;    // decoder: code bytes shifted into `data` so far (0 = not primed)
    uint32_t status;      // RCB_ST_* of the last call
    uint32_t pad;
};

struct ByteSinkEnc {
    uint8_t* out;
    uint64_t pos, cap;
    __device__ void put(uint32_t lo_hi, uint32_t sh) {
        for (uint32_t b = 0; b < sh; b += 8) put_byte(lo_hi >> (24 - b));
    }
    __device__ void put_byte(uint32_t b) {
        if (pos < cap) out[pos] = (uint8_t)b;
        pos++;
    }
};

struct ByteSinkDec {
    const uint8_t* code;  // device copy of code[base .. len): only the unread tail travels per call
    uint64_t len, pos;
    uint64_t data;
    bool truncated;
    uint64_t base;
    __device__ void put(uint32_t, uint32_t sh) {
        for (uint32_t b = 0; b < sh; b += 8) put_byte(0);
    }
    __device__ void put_byte(uint32_t) {  // src/decoder.rs:31-35
        uint32_t v = 0;
        if (pos < len) v = code[pos - base]; else truncated = true;
        pos++;
        data = (data << 8) | v;
    }
};

// Encoder::encode for n symbols (+ finish when `finish` != 0).  per_symbol (may be null)
// receives the byte count each symbol produced -- encode()'s return value.
__global__ void encode_stream_kernel(StreamState* st, const uint8_t* syms, uint64_t n, int sym_bytes,
                                     const uint2* tab, const ModelHdr* hdr, uint32_t K, uint8_t* out,
                                     uint64_t cap, uint64_t* n_out, uint32_t* per_symbol, int finish) {
    if (threadIdx.x || blockIdx.x) return;
    uint64_t lo = st->lower_bound, rg = st->range;
    ModelHdr h;
    h.flags = 0;
    if (n) h = *hdr;  // finish alone needs no model
    const bool pow2 = (h.flags & MODEL_POW2) != 0;
    ByteSinkEnc sink{out, 0, cap};
    uint32_t err = 0;
    for (uint64_t i = 0; i < n && !err; i++) {
        uint32_t s = sym_bytes == 1 ? syms[i] : (uint32_t)syms[2 * i] | ((uint32_t)syms[2 * i + 1] << 8);
        if (s >= K) {
            err = ST_SYMBOL_RANGE;
            break;
        }
        const uint2 e = tab[s];
        const uint64_t before = sink.pos;
        if (pow2) update_symbol<true, true>(lo, rg, e.x, e.y, h.div, sink, err);
        else update_symbol<false, true>(lo, rg, e.x, e.y, h.div, sink, err);
        if (per_symbol) per_symbol[i] = (uint32_t)(sink.pos - before);
    }
    if (finish && !err) {  // src/encoder.rs:40-46
        for (int i = 0; i < 8; i++) {
            sink.put_byte((uint32_t)(lo >> 56));
            lo <<= 8;
            rg <<= 8;
        }
    }
    if (!err && sink.pos > cap) err = ST_OUT_CAPACITY;
    st->lower_bound = lo;
    st->range = rg;
    st->status = err;
    *n_out = sink.pos;
}

// Decoder::new (first call: consumed == 0) + Decoder::decode for n symbols.
__global__ void decode_stream_kernel(StreamState* st, const uint8_t* code, uint64_t code_base, uint64_t len,
                                     uint64_t n, int sym_bytes, const uint2* tab, const ModelHdr* hdr, uint32_t K,
                                     uint8_t* out_syms) {
    if (threadIdx.x || blockIdx.x) return;
    uint64_t lo = st->lower_bound, rg = st->range;
    const ModelHdr h = *hdr;
    const bool pow2 = (h.flags & MODEL_POW2) != 0;
    ByteSinkDec sink{code, len, st->consumed, st->data, false, code_base};
    uint32_t err = 0;
    if (sink.pos == 0)  // src/decoder.rs:14-23
        for (int i = 0; i < 8; i++) sink.put_byte(0);
    for (uint64_t i = 0; i < n && !err && !sink.truncated; i++) {
        const uint64_t rpt = pow2 ? range_par_total<true>(rg, h.div) : range_par_total<false>(rg, h.div);
        const uint64_t d = sink.data - lo;  // examples/sample_impl.rs:29
        const uint32_t sym = find_index_exact(d, rpt, K, [&](uint32_t j) { return tab[j].x; });
        const uint2 e = tab[sym];
        const uint64_t nlo = lo + rpt * (uint64_t)e.x;
        if (nlo < lo) {
            err = ST_LOWER_OVERFLOW;
            break;
        }
        lo = nlo;
        rg = rpt * (uint64_t)e.y;
        renorm<true>(lo, rg, sink, err);
        if (sym_bytes == 1) {
            out_syms[i] = (uint8_t)sym;
        } else {
            out_syms[2 * i] = (uint8_t)sym;
            out_syms[2 * i + 1] = (uint8_t)(sym >> 8);
        }
    }
    if (!err && sink.truncated) err = ST_TRUNCATED;  // src/decoder.rs:33
    st->lower_bound = lo;
    st->range = rg;
    st->data = sink.data;
    st->consumed = sink.pos;
    st->status = err;
}

}  // namespace rcb
