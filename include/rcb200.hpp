// rcb200.hpp -- host-side mirror of the reference crate's public API
// (diegodox/range_coder_rust, src/lib.rs:1-13) on top of the C ABI in rcb200.h.
//
// The reference is Rust; this image has no Rust toolchain, so the host side is
// C++ with the same names, argument meaning and error behaviour:
//   range_coder::PModel            src/pmodel.rs:4-41
//   range_coder::RangeCoder        src/range_coder.rs:7-40,138-146
//   range_coder::Encoder           src/encoder.rs:7-55
//   range_coder::Decoder           src/decoder.rs:6-55
//   range_coder::error::RangeCoderError  src/error.rs:3-13
// plus range_coder::gpu::{encode_chunks, decode_chunks}: the bulk entry points a
// caller switches its per-symbol loops to (INTEGRATION.md shows the Rust binding).
//
// Every arithmetic step runs on the GPU (rcb_encode_stream / rcb_decode_stream /
// rcb_*_chunks); nothing here codes symbols on the CPU.  Where the reference
// panics (unwrap on an Err, divide by zero, pop_front on an empty buffer) these
// types throw; where it would never return (a coded symbol with c_freq == 0)
// they throw RangeCoderPanic("zero frequency").
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <deque>
#include <stdexcept>
#include <string>
#include <vector>

#include "rcb200.h"

namespace range_coder {

namespace error {
// src/error.rs:3-13
struct RangeCoderError : std::runtime_error {
    enum Kind { LowerBoundOverflow, UpperBoundOverflow } kind;
    uint64_t lower_bound, add_val, range;
    RangeCoderError(Kind k, uint64_t lo, uint64_t add, uint64_t rg)
        : std::runtime_error(k == LowerBoundOverflow
                                 ? "Overflow happend while lower_bound uppdating " + std::to_string(lo) + " + " +
                                       std::to_string(add) + " , " + std::to_string(rg)
                                 : "Overflow happend when calc upper_bound " + std::to_string(lo) + " + " +
                                       std::to_string(rg)),
          kind(k), lower_bound(lo), add_val(add), range(rg) {}
};
}  // namespace error

// what a Rust panic (unwrap, divide by zero, index out of bounds) becomes here
struct RangeCoderPanic : std::runtime_error {
    int code;  // rcb_error
    RangeCoderPanic(int c, const std::string& where) : std::runtime_error(where + ": " + rcb_strerror(c)), code(c) {}
};

inline void check(int rc, const char* where, const rcb_stream_state* st = nullptr) {
    if (rc == RCB_OK) return;
    if (st && rc == RCB_ERR_LOWER_OVERFLOW)
        throw error::RangeCoderError(error::RangeCoderError::LowerBoundOverflow, st->lower_bound, 0, st->range);
    if (st && rc == RCB_ERR_UPPER_OVERFLOW)
        throw error::RangeCoderError(error::RangeCoderError::UpperBoundOverflow, st->lower_bound, 0, st->range);
    throw RangeCoderPanic(rc, where);
}

// One GPU context per thread (the reference's types are single-threaded owned values).
class Context {
public:
    explicit Context(int device = 0) { check(rcb_ctx_create(device, nullptr, &h_), "rcb_ctx_create"); }
    ~Context() { rcb_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    rcb_ctx* handle() const { return h_; }
    static Context& thread_default() {
        thread_local Context ctx(0);
        return ctx;
    }

private:
    rcb_ctx* h_ = nullptr;
};

class Decoder;

// src/pmodel.rs:4-41.  The trait has no alphabet-size method; a GPU kernel needs one to
// snapshot the tables, so alphabet_count() (examples/sample_impl.rs:55-57) is part of the mirror.
class PModel {
public:
    virtual ~PModel() = default;
    virtual uint32_t c_freq(size_t index) const = 0;
    virtual uint32_t cum_freq(size_t index) const = 0;
    virtual uint32_t total_freq() const = 0;
    virtual size_t alphabet_count() const = 0;
    // Not consulted by the GPU path: the lookup of examples/sample_impl.rs:27-45 is built in.
    virtual size_t find_index(const Decoder&) const { throw std::logic_error("find_index runs on the GPU"); }
    // src/pmodel.rs:14-40 (f64 diagnostic, never used by the coder)
    double ideal_code_length(size_t index) const {
        double p = (double)c_freq(index);
        if (p == 0.0) throw std::domain_error("code length is undefind when probability is zero");
        return (std::log((double)total_freq()) - std::log(p)) / std::log(2.0);
    }
};

// Dense device snapshot of a PModel (the table boundary of SURVEY 8 b2).
class ModelSnapshot {
public:
    ModelSnapshot(Context& ctx, const PModel& pm) : ctx_(ctx), K_((uint32_t)pm.alphabet_count()) {
        std::vector<uint32_t> c(K_), cum(K_);
        for (uint32_t i = 0; i < K_; i++) {
            c[i] = pm.c_freq(i);
            cum[i] = pm.cum_freq(i);
        }
        uint32_t total = pm.total_freq();
        check(rcb_model_create(ctx.handle(), K_, 1, &m_), "rcb_model_create");
        int rc = rcb_model_from_tables(ctx.handle(), m_, c.data(), cum.data(), &total);
        if (rc) {
            rcb_model_destroy(m_);
            check(rc, "rcb_model_from_tables");  // total_freq == 0: the reference divides by zero
        }
    }
    ~ModelSnapshot() { rcb_model_destroy(m_); }
    ModelSnapshot(const ModelSnapshot&) = delete;
    ModelSnapshot& operator=(const ModelSnapshot&) = delete;
    const rcb_model* handle() const { return m_; }
    uint32_t alphabet_count() const { return K_; }
    int sym_bytes() const { return K_ <= 256 ? 1 : 2; }

private:
    Context& ctx_;
    uint32_t K_;
    rcb_model* m_ = nullptr;
};

// src/range_coder.rs:7-40,138-146 (state + getters; the update runs on the GPU)
class RangeCoder {
public:
    RangeCoder() { rcb_stream_state_init(&st_); }
    uint64_t lower_bound() const { return st_.lower_bound; }
    uint64_t range() const { return st_.range; }
    uint64_t range_par_total(uint32_t total_freq) const {
        if (total_freq == 0) throw RangeCoderPanic(RCB_ERR_ZERO_TOTAL, "range_par_total");
        return st_.range / (uint64_t)total_freq;
    }
    uint64_t upper_bound() const {
        uint64_t u = st_.lower_bound + st_.range;
        if (u < st_.lower_bound)
            throw error::RangeCoderError(error::RangeCoderError::UpperBoundOverflow, st_.lower_bound, 0, st_.range);
        return u;
    }
    rcb_stream_state& state() { return st_; }
    const rcb_stream_state& state() const { return st_; }

private:
    rcb_stream_state st_;
};

// src/encoder.rs:7-55
class Encoder {
public:
    RangeCoder range_coder;  // pub field in the reference (src/encoder.rs:8)
    explicit Encoder(Context& ctx = Context::thread_default()) : ctx_(ctx) {}
    const std::deque<uint8_t>& peek_code() const { return code_; }
    // one symbol; returns the number of bytes it produced (computed on the GPU)
    uint32_t encode(const PModel& pmodel, size_t index) {
        ModelSnapshot snap(ctx_, pmodel);
        return encode(snap, index);
    }
    uint32_t encode(const ModelSnapshot& snap, size_t index) {
        uint8_t sym[2] = {(uint8_t)index, (uint8_t)(index >> 8)};
        if (index >= snap.alphabet_count()) throw RangeCoderPanic(RCB_ERR_SYMBOL_OUT_OF_RANGE, "Encoder::encode");
        uint8_t out[16];
        uint64_t n = 0;
        check(rcb_encode_stream(ctx_.handle(), &range_coder.state(), sym, 1, snap.sym_bytes(), snap.handle(), out,
                                sizeof out, &n, nullptr, 0),
              "Encoder::encode", &range_coder.state());
        code_.insert(code_.end(), out, out + n);
        return (uint32_t)n;
    }
    // a slice of symbols in one call (same bytes as calling encode() once per symbol)
    template <class Sym>
    void encode_slice(const ModelSnapshot& snap, const Sym* syms, size_t n) {
        static_assert(sizeof(Sym) == 1 || sizeof(Sym) == 2, "u8 or u16 symbols");
        if ((int)sizeof(Sym) != snap.sym_bytes()) throw std::invalid_argument("symbol width does not match K");
        std::vector<uint8_t> out(n * 8 + 64);
        uint64_t produced = 0;
        check(rcb_encode_stream(ctx_.handle(), &range_coder.state(), syms, n, (int)sizeof(Sym), snap.handle(),
                                out.data(), out.size(), &produced, nullptr, 0),
              "Encoder::encode_slice", &range_coder.state());
        code_.insert(code_.end(), out.begin(), out.begin() + produced);
    }
    // src/encoder.rs:40-46: consumes the encoder
    std::deque<uint8_t> finish() {
        uint8_t out[8];
        uint64_t n = 0;
        check(rcb_encode_stream(ctx_.handle(), &range_coder.state(), nullptr, 0, 1, nullptr, out, sizeof out, &n,
                                nullptr, 1),
              "Encoder::finish");
        code_.insert(code_.end(), out, out + n);
        return std::move(code_);
    }

private:
    Context& ctx_;
    std::deque<uint8_t> code_;
};

// src/decoder.rs:6-55
class Decoder {
public:
    template <class Bytes>
    explicit Decoder(const Bytes& code, Context& ctx = Context::thread_default())
        : ctx_(ctx), buffer_(code.begin(), code.end()) {
        if (buffer_.size() < 8) throw RangeCoderPanic(RCB_ERR_TRUNCATED_STREAM, "Decoder::new");  // decoder.rs:33
    }
    const RangeCoder& range_coder() const { return rc_; }
    uint64_t data() const { return rc_.state().data; }
    size_t decode(const PModel& pmodel) {
        ModelSnapshot snap(ctx_, pmodel);
        return decode(snap);
    }
    size_t decode(const ModelSnapshot& snap) {
        uint8_t sym[2] = {0, 0};
        check(rcb_decode_stream(ctx_.handle(), &rc_.state(), buffer_.data(), buffer_.size(), 1, snap.sym_bytes(),
                                snap.handle(), sym),
              "Decoder::decode", &rc_.state());
        return (size_t)sym[0] | ((size_t)sym[1] << 8);
    }
    template <class Sym>
    std::vector<Sym> decode_n(const ModelSnapshot& snap, size_t n) {
        std::vector<Sym> out(n);
        check(rcb_decode_stream(ctx_.handle(), &rc_.state(), buffer_.data(), buffer_.size(), n, (int)sizeof(Sym),
                                snap.handle(), out.data()),
              "Decoder::decode_n", &rc_.state());
        return out;
    }

private:
    Context& ctx_;
    RangeCoder rc_;
    std::vector<uint8_t> buffer_;
};

// Bulk entry points: every chunk of chunk_syms symbols is one independent Encoder run.
namespace gpu {
struct Encoded {
    std::vector<uint8_t> stream;     // concatenated Encoder::finish() outputs
    std::vector<uint64_t> offsets;   // chunk i = stream[offsets[i] .. offsets[i+1])
    // restart points (rcb200.h): the Encoder's state every restart_syms symbols of a chunk -- side information
    // that lets decode_chunks run several GPU lanes per chunk; the stream above is unchanged by it
    uint64_t restart_syms = 0;
    std::vector<rcb_restart_point> restart;
};
// restart points: 16 parts per chunk for chunks of >= 32 Ki symbols, else 4, when the part is a whole number of
// 64-symbol units (0: none)
inline uint64_t default_restart_syms(uint64_t chunk_syms) {
    if (chunk_syms >= 32768 && chunk_syms % 1024 == 0) return chunk_syms / 16;
    return chunk_syms % 256 == 0 ? chunk_syms / 4 : 0;
}
template <class Sym>
Encoded encode_chunks(Context& ctx, const ModelSnapshot& snap, const std::vector<Sym>& syms, uint64_t chunk_syms,
                      uint64_t restart_syms = ~0ull) {
    const uint64_t n = syms.size(), n_chunks = chunk_syms ? (n + chunk_syms - 1) / chunk_syms : 0;
    Encoded e;
    e.offsets.resize(n_chunks + 1);
    e.stream.resize(rcb_encode_bound(ctx.handle(), snap.handle(), n, (int)sizeof(Sym), chunk_syms) + 16);
    if (restart_syms == ~0ull) restart_syms = default_restart_syms(chunk_syms);
    const uint64_t per = rcb_restart_points_per_chunk(chunk_syms, restart_syms);
    e.restart_syms = per ? restart_syms : 0;
    e.restart.resize(n_chunks * per);
    uint64_t bytes = 0;
    check(rcb_encode_host_restart(ctx.handle(), syms.data(), n, (int)sizeof(Sym), chunk_syms, snap.handle(),
                                  e.stream.data(), e.stream.size(), e.offsets.data(), &bytes, e.restart_syms,
                                  e.restart.empty() ? nullptr : e.restart.data()),
          "gpu::encode_chunks");
    e.stream.resize(bytes);
    return e;
}
template <class Sym>
std::vector<Sym> decode_chunks(Context& ctx, const ModelSnapshot& snap, const Encoded& e, uint64_t n_syms,
                               uint64_t chunk_syms) {
    std::vector<Sym> out(n_syms);
    std::vector<uint8_t> padded(e.stream);
    padded.resize((padded.size() + 31) & ~size_t(15));
    check(rcb_decode_host_restart(ctx.handle(), padded.data(), e.offsets.data(), n_syms, (int)sizeof(Sym), chunk_syms,
                                  snap.handle(), out.data(), e.restart_syms,
                                  e.restart.empty() ? nullptr : e.restart.data()),
          "gpu::decode_chunks");
    return out;
}

// the table follows the symbols -- what a caller of the reference gets by updating its
// FreqTable between Encoder::encode calls (`&T` per call, src/encoder.rs:24): counts start at 1, the coded
// symbol gains `inc`, all counts are halved (rounding up) when the total would pass `limit`.  One
// independent Encoder + table per chunk.  (rcb_adaptive_encode_chunks works on device buffers.)
struct AdaptiveFreq {
    uint32_t K, inc = 24, limit = 60000;
};
template <class Sym>
Encoded adaptive_encode_chunks(Context& ctx, const AdaptiveFreq& m, const std::vector<Sym>& syms, uint64_t chunk_syms) {
    const rcb_adaptive_params p{m.K, m.inc, m.limit};
    const uint64_t n = syms.size(), n_chunks = (n + chunk_syms - 1) / chunk_syms;
    const uint64_t cap = rcb_adaptive_encode_bound(&p, n, chunk_syms);
    if (!cap) throw RangeCoderPanic(RCB_ERR_UNSUPPORTED, "gpu::adaptive_encode_chunks");
    void *d_syms = nullptr, *d_out = nullptr, *d_off = nullptr;
    check(rcb_device_alloc(ctx.handle(), n * sizeof(Sym), &d_syms), "rcb_device_alloc");
    check(rcb_device_alloc(ctx.handle(), cap, &d_out), "rcb_device_alloc");
    check(rcb_device_alloc(ctx.handle(), (n_chunks + 1) * 8, &d_off), "rcb_device_alloc");
    check(rcb_copy_to_device(ctx.handle(), d_syms, syms.data(), n * sizeof(Sym)), "rcb_copy_to_device");
    uint64_t bytes = 0;
    const int rc = rcb_adaptive_encode_chunks(ctx.handle(), d_syms, n, (int)sizeof(Sym), chunk_syms, &p, (uint8_t*)d_out,
                                              cap, (uint64_t*)d_off, nullptr, &bytes);
    Encoded e;
    if (rc == RCB_OK) {
        e.stream.resize(bytes);
        e.offsets.resize(n_chunks + 1);
        rcb_copy_to_host(ctx.handle(), e.stream.data(), d_out, bytes);
        rcb_copy_to_host(ctx.handle(), e.offsets.data(), d_off, (n_chunks + 1) * 8);
    }
    rcb_device_free(ctx.handle(), d_syms);
    rcb_device_free(ctx.handle(), d_out);
    rcb_device_free(ctx.handle(), d_off);
    check(rc, "gpu::adaptive_encode_chunks");
    return e;
}
template <class Sym>
std::vector<Sym> adaptive_decode_chunks(Context& ctx, const AdaptiveFreq& m, const Encoded& e, uint64_t n_syms,
                                        uint64_t chunk_syms) {
    const rcb_adaptive_params p{m.K, m.inc, m.limit};
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    const uint64_t padded = (e.stream.size() + 31) & ~uint64_t(15);
    void *d_st = nullptr, *d_off = nullptr, *d_out = nullptr;
    check(rcb_device_alloc(ctx.handle(), padded, &d_st), "rcb_device_alloc");
    check(rcb_device_alloc(ctx.handle(), (n_chunks + 1) * 8, &d_off), "rcb_device_alloc");
    check(rcb_device_alloc(ctx.handle(), n_syms * sizeof(Sym), &d_out), "rcb_device_alloc");
    check(rcb_copy_to_device(ctx.handle(), d_st, e.stream.data(), e.stream.size()), "rcb_copy_to_device");
    check(rcb_copy_to_device(ctx.handle(), d_off, e.offsets.data(), (n_chunks + 1) * 8), "rcb_copy_to_device");
    std::vector<Sym> out(n_syms);
    const int rc = rcb_adaptive_decode_chunks(ctx.handle(), (const uint8_t*)d_st, (const uint64_t*)d_off, n_syms,
                                              (int)sizeof(Sym), chunk_syms, &p, d_out, nullptr);
    if (rc == RCB_OK) rcb_copy_to_host(ctx.handle(), out.data(), d_out, n_syms * sizeof(Sym));
    rcb_device_free(ctx.handle(), d_st);
    rcb_device_free(ctx.handle(), d_off);
    rcb_device_free(ctx.handle(), d_out);
    check(rc, "gpu::adaptive_decode_chunks");
    return out;
}

// A static table shared by chunks that live on several GPUs, one host thread driving all
// of them: GPU g owns the chunks [n_chunks*g/G, n_chunks*(g+1)/G).  The frequency table is what the
// reference's caller builds with FreqTable::add_alphabet_freq over ALL the data + calc_cum
// (examples/sample_impl.rs:77-81): per-GPU histograms of the local chunks, ONE all-reduce of the K
// counts over NVLink (rcb_allreduce_counts_multi), then the same deterministic table on every GPU.
// No payload byte crosses GPUs; the per-GPU streams concatenate in chunk order.
class MultiGpu {
public:
    explicit MultiGpu(int n_gpus) {
        if (n_gpus < 1 || n_gpus > rcb_device_count()) throw RangeCoderPanic(RCB_ERR_INVALID_ARGUMENT, "MultiGpu");
        ctx_.resize(n_gpus, nullptr);
        comm_.resize(n_gpus, nullptr);
        model_.resize(n_gpus, nullptr);
        for (int g = 0; g < n_gpus; g++) check(rcb_ctx_create(g, nullptr, &ctx_[g]), "rcb_ctx_create");
        check(rcb_comm_init_all(ctx_.data(), n_gpus, comm_.data()), "rcb_comm_init_all");
    }
    ~MultiGpu() {
        for (auto m : model_) rcb_model_destroy(m);
        for (auto k : comm_) rcb_comm_destroy(k);
        for (auto c : ctx_) rcb_ctx_destroy(c);
    }
    MultiGpu(const MultiGpu&) = delete;
    MultiGpu& operator=(const MultiGpu&) = delete;
    int gpus() const { return (int)ctx_.size(); }

    // histogram (sharded) -> all-reduce -> table -> encode, every chunk on its own GPU
    template <class Sym>
    Encoded encode_chunks(const std::vector<Sym>& syms, uint32_t K, uint64_t chunk_syms) {
        const int G = gpus();
        const uint64_t n = syms.size(), n_chunks = (n + chunk_syms - 1) / chunk_syms;
        std::vector<void*> d_syms(G, nullptr), d_counts(G, nullptr);
        std::vector<uint64_t> first(G + 1);
        for (int g = 0; g <= G; g++) first[g] = std::min<uint64_t>(n, n_chunks * g / G * chunk_syms);
        for (int g = 0; g < G; g++) {
            const uint64_t cnt = first[g + 1] - first[g];
            check(rcb_device_alloc(ctx_[g], cnt * sizeof(Sym), &d_syms[g]), "rcb_device_alloc");
            check(rcb_device_alloc(ctx_[g], (uint64_t)K * 8, &d_counts[g]), "rcb_device_alloc");
            check(rcb_copy_to_device(ctx_[g], d_syms[g], syms.data() + first[g], cnt * sizeof(Sym)), "rcb_copy_to_device");
            check(rcb_histogram(ctx_[g], d_syms[g], cnt, (int)sizeof(Sym), K, 0, d_counts[g]), "rcb_histogram");
        }
        check(rcb_allreduce_counts_multi(ctx_.data(), comm_.data(), d_counts.data(), K, G), "rcb_allreduce_counts_multi");
        Encoded e;
        e.offsets.assign(1, 0);
        for (int g = 0; g < G; g++) {
            if (model_[g]) rcb_model_destroy(model_[g]);
            model_[g] = nullptr;
            check(rcb_model_create(ctx_[g], K, 1, &model_[g]), "rcb_model_create");
            check(rcb_model_from_counts(ctx_[g], model_[g], d_counts[g], 8), "rcb_model_from_counts");
            const uint64_t cnt = first[g + 1] - first[g], chunks = (cnt + chunk_syms - 1) / chunk_syms;
            const uint64_t cap = rcb_encode_bound(ctx_[g], model_[g], cnt, (int)sizeof(Sym), chunk_syms) + 16;
            void *d_out = nullptr, *d_off = nullptr;
            check(rcb_device_alloc(ctx_[g], cap, &d_out), "rcb_device_alloc");
            check(rcb_device_alloc(ctx_[g], (chunks + 1) * 8, &d_off), "rcb_device_alloc");
            uint64_t bytes = 0;
            check(rcb_encode_chunks(ctx_[g], d_syms[g], cnt, (int)sizeof(Sym), chunk_syms, model_[g], (uint8_t*)d_out,
                                    cap, (uint64_t*)d_off, nullptr, &bytes),
                  "rcb_encode_chunks");
            const uint64_t base = e.stream.size(), k0 = e.offsets.size() - 1;
            e.stream.resize(base + bytes);
            e.offsets.resize(k0 + chunks + 1);
            check(rcb_copy_to_host(ctx_[g], e.stream.data() + base, d_out, bytes), "rcb_copy_to_host");
            check(rcb_copy_to_host(ctx_[g], e.offsets.data() + k0, d_off, (chunks + 1) * 8), "rcb_copy_to_host");
            for (uint64_t i = 0; i <= chunks; i++) e.offsets[k0 + i] += base;  // local -> global offsets
            rcb_device_free(ctx_[g], d_out);
            rcb_device_free(ctx_[g], d_off);
            rcb_device_free(ctx_[g], d_syms[g]);
            rcb_device_free(ctx_[g], d_counts[g]);
        }
        return e;
    }

    // every GPU decodes its own chunks under the table of the last encode_chunks
    template <class Sym>
    std::vector<Sym> decode_chunks(const Encoded& e, uint64_t n_syms, uint64_t chunk_syms) {
        const int G = gpus();
        const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
        std::vector<Sym> out(n_syms);
        for (int g = 0; g < G; g++) {
            const uint64_t c0 = n_chunks * g / G, c1 = n_chunks * (g + 1) / G;
            const uint64_t s0 = std::min(n_syms, c0 * chunk_syms), s1 = std::min(n_syms, c1 * chunk_syms);
            if (s1 == s0) continue;
            std::vector<uint64_t> offs(e.offsets.begin() + c0, e.offsets.begin() + c1 + 1);
            const uint64_t base = offs[0];
            for (auto& o : offs) o -= base;
            std::vector<uint8_t> part(e.stream.begin() + base, e.stream.begin() + base + offs.back());
            part.resize((part.size() + 31) & ~size_t(15));
            check(rcb_decode_host(ctx_[g], part.data(), offs.data(), s1 - s0, (int)sizeof(Sym), chunk_syms, model_[g],
                                  out.data() + s0),
                  "rcb_decode_host");
        }
        return out;
    }
    const rcb_model* model(int g) const { return model_[g]; }
    rcb_ctx* context(int g) const { return ctx_[g]; }

private:
    std::vector<rcb_ctx*> ctx_;
    std::vector<rcb_comm*> comm_;
    std::vector<rcb_model*> model_;
};
}  // namespace gpu

}  // namespace range_coder
