/*
 * rc_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the range coder in diegodox/range_coder_rust
 * (crate `range_coder` v0.1.0).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` legs may load this library; the
 * product path (range_coder_rust_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors / known-answer tests
 * (no #[test], output bytes of examples/sample_impl.rs are printed, not asserted)
 * and no Rust toolchain exists in this environment, so the reference itself
 * cannot be run.  The oracle is pinned only by (1) the source semantics cited
 * per function below, (2) the hand-checkable `sample_impl` vector in
 * tests/golden/, (3) an independent pure-Python transliteration
 * (oracle/rc_pyref.py) that must agree byte for byte.  rust/pin_reference/ is
 * the one-command pin for anyone with cargo: it runs the unmodified crate over
 * the same golden vectors (it cannot run here).
 *
 * Citations are path:line under /root/reference/.
 */
#ifndef RC_ORACLE_H
#define RC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes (the reference panics / loops forever in these cases). */
enum {
    RCO_OK = 0,
    RCO_ERR_LOWER_OVERFLOW = -1, /* src/range_coder.rs:68-81  LowerBoundOverflow          */
    RCO_ERR_UPPER_OVERFLOW = -2, /* src/range_coder.rs:111,138-146 UpperBoundOverflow      */
    RCO_ERR_ZERO_TOTAL = -3,     /* src/range_coder.rs:39 divide by zero panic              */
    RCO_ERR_ZERO_FREQ = -4,      /* range==0 => loop 1 never ends in the reference (:83-85) */
    RCO_ERR_CAPACITY = -5,       /* caller's output buffer too small (oracle-only)          */
    RCO_ERR_TRUNCATED = -6,      /* src/decoder.rs:33 pop_front().unwrap() on empty buffer  */
    RCO_ERR_SYMBOL_RANGE = -7    /* examples/sample_impl.rs:19 Vec::get(i).unwrap() on i>=K */
};

/* src/range_coder.rs:7-12 */
typedef struct {
    uint64_t lower_bound;
    uint64_t range;
} rco_range_coder;

/* src/range_coder.rs:13-20,26-28 */
void rco_rc_new(rco_range_coder *rc);
/* src/range_coder.rs:38-40 (caller guarantees total_freq != 0) */
uint64_t rco_range_par_total(const rco_range_coder *rc, uint32_t total_freq);
/* src/range_coder.rs:138-146; returns RCO_OK or RCO_ERR_UPPER_OVERFLOW */
int rco_upper_bound(const rco_range_coder *rc, uint64_t *upper);
/* src/range_coder.rs:53-92.  out must hold >= 16 bytes; *n_out = bytes pushed. */
int rco_param_update(rco_range_coder *rc, uint32_t c_freq, uint32_t cum_freq,
                     uint32_t total_freq, uint8_t *out, int *n_out);

/* src/encoder.rs:24-46: a whole Encoder run (new, encode per symbol, finish).
 * syms: n symbols of sym_bytes (1 = u8, 2 = u16 LE, 4 = u32) each < K.
 * Returns the code length (>= 8) or a negative RCO_ERR_*. */
int64_t rco_encode(const void *syms, uint64_t n, int sym_bytes, uint32_t K,
                   const uint32_t *c, const uint32_t *cum, uint32_t total,
                   uint8_t *out, uint64_t cap);

/* The same run stopped after n symbols, before finish(): lower_bound and range of the Encoder's
 * RangeCoder (src/encoder.rs:7-11) and, as the return value, the bytes encode() has returned so
 * far -- what the GPU path's restart points record (include/rcb200.h: rcb_restart_point). */
int64_t rco_encode_state(const void *syms, uint64_t n, int sym_bytes, uint32_t K,
                         const uint32_t *c, const uint32_t *cum, uint32_t total,
                         uint64_t *lower, uint64_t *range);

/* src/decoder.rs:14-54 with examples/sample_impl.rs:27-45 as find_index.
 * Returns the number of code bytes consumed (8 + sum n) or a negative error. */
int64_t rco_decode(const uint8_t *code, uint64_t len, uint64_t n_syms,
                   int sym_bytes, uint32_t K, const uint32_t *c,
                   const uint32_t *cum, uint32_t total, void *out_syms);

/* examples/sample_impl.rs:58-60 (called in the loop at :78-80) */
void rco_histogram(const void *syms, uint64_t n, int sym_bytes, uint32_t K,
                   uint64_t *counts);
/* examples/sample_impl.rs:61-69: exclusive prefix sum, returns total.
 * (u32 wrapping exactly like release-mode Rust.) */
uint32_t rco_calc_cum(const uint32_t *c, uint32_t K, uint32_t *cum);
/* Build-defined extension: identity when the u64 counts sum to <= 2^32-1 (the
 * reference rule), otherwise rescale to total = 2^31 exactly: c' = c ? max(1, floor(c*2^31/sum)) : 0
 * and the difference 2^31 - sum(c') is added to the largest count (lowest index on ties).
 * Returns 1 if rescaled. */
int rco_normalise(const uint64_t *counts, uint32_t K, uint32_t *c);

/* Chunked drivers used as the CPU baseline ("one chunk per thread at a time").
 * Chunk i covers symbols [i*chunk_syms, min(n,(i+1)*chunk_syms)).
 * Model tables: shared (per_chunk_model=0: c[K], cum[K], total[1]) or one per
 * chunk (c[n_chunks][K], cum[n_chunks][K], total[n_chunks]).
 * out: n_chunks rows of out_pitch bytes; lens[i] = code length or error. */
int rco_encode_chunks(const void *syms, uint64_t n, int sym_bytes,
                      uint64_t chunk_syms, uint32_t K, const uint32_t *c,
                      const uint32_t *cum, const uint32_t *total,
                      int per_chunk_model, uint8_t *out, uint64_t out_pitch,
                      int64_t *lens, int n_threads);
/* stream + offsets[n_chunks+1] framing (chunk i = stream[offsets[i]..offsets[i+1])). */
int rco_decode_chunks(const uint8_t *stream, const uint64_t *offsets,
                      uint64_t n, int sym_bytes, uint64_t chunk_syms,
                      uint32_t K, const uint32_t *c, const uint32_t *cum,
                      const uint32_t *total, int per_chunk_model,
                      void *out_syms, int64_t *consumed, int n_threads);

/* ---- adaptive-per-symbol model (build-defined) -------------------
 * The reference takes `&T: PModel` on EVERY call (src/encoder.rs:24, src/decoder.rs:38), so a caller may
 * change the table between symbols; it ships no such model.  The build defines one -- the textbook
 * adaptive frequency count with periodic halving -- as the caller-side loop it would be in Rust:
 *     table: c[i] = 1 for i < K; cum = exclusive prefix sum; total = K            (FreqTable + calc_cum)
 *     for each symbol s:  encoder.encode(&table, s)        with the table as it is BEFORE the update
 *                         c[s] += inc; if sum(c) > limit { c[i] = (c[i] + 1) >> 1 for all i }; calc_cum()
 * and the mirror image in the decoder (decode, then the same update).  Requires 1 <= inc, K <= limit,
 * limit + inc <= 2^32 - 1.  Bytes = exactly what the reference's Encoder emits for that call sequence. */
int64_t rco_adaptive_encode(const void *syms, uint64_t n, int sym_bytes, uint32_t K, uint32_t inc,
                            uint32_t limit, uint8_t *out, uint64_t cap);
int64_t rco_adaptive_decode(const uint8_t *code, uint64_t len, uint64_t n_syms, int sym_bytes,
                            uint32_t K, uint32_t inc, uint32_t limit, void *out_syms);
/* chunked drivers (every chunk restarts coder AND table), same conventions as rco_encode_chunks */
int rco_adaptive_encode_chunks(const void *syms, uint64_t n, int sym_bytes, uint64_t chunk_syms,
                               uint32_t K, uint32_t inc, uint32_t limit, uint8_t *out,
                               uint64_t out_pitch, int64_t *lens, int n_threads);
int rco_adaptive_decode_chunks(const uint8_t *stream, const uint64_t *offsets, uint64_t n,
                               int sym_bytes, uint64_t chunk_syms, uint32_t K, uint32_t inc,
                               uint32_t limit, void *out_syms, int64_t *consumed, int n_threads);

/* Synthetic data (not part of the reference).
 * u_j = mix64(seed + j*GOLDEN); r = u_j>>32; table t = (j/chunk_syms) % n_tables;
 * symbol = #{ i in [0,K-1) : thr[t][i] <= r }.  thr rows hold K-1 u32 thresholds
 * floor(2^32*CDF(i)) computed once on the host in double. */
void rco_generate(void *out, uint64_t first, uint64_t n, int sym_bytes,
                  uint32_t K, uint64_t seed, const uint32_t *thr,
                  uint32_t n_tables, uint64_t chunk_syms, int n_threads);

int rco_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
