/*
 * rcb200.h -- C ABI of the B200-native chunk-parallel range coder.
 *
 * Drop-in boundary for the encode/decode hot path of the reference crate
 * diegodox/range_coder_rust (crate `range_coder` v0.1.0).  The reference is a
 * Rust library with a per-symbol API and no FFI of its own; these are the entry
 * points a Rust `extern "C"` block binds to replace the
 * per-symbol loops of its callers.  Each entry point cites the reference
 * interface it replaces (paths under /root/reference).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - every function returns RCB_OK (0) or a negative rcb_error; nothing unwinds.
 *   - pointers named d_* are device pointers on the ctx's device, 16-byte
 *     aligned; h_* are host pointers.  Caller owns all buffers; the ctx owns
 *     scratch (staging rows, lengths) and is not thread-safe (one ctx per GPU
 *     per thread -- the reference's Encoder/Decoder are single-threaded owned
 *     values too, src/encoder.rs:7-11, src/decoder.rs:6-12).
 *   - work is issued on the ctx's CUDA stream.  Entry points that return a
 *     host result synchronise that stream; the *_async forms do not.
 *   - a "chunk" is one independent reference `Encoder` run: chunk i covers
 *     symbols [i*chunk_syms, min(n_syms,(i+1)*chunk_syms)) and its bytes
 *     stream[offsets[i] .. offsets[i+1]) are exactly what
 *     `Encoder::new(); encode(..)*; finish()` (src/encoder.rs:13-46) returns
 *     for those symbols under that chunk's model.
 */
#ifndef RCB200_H
#define RCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCB_VERSION 100 /* 0.1.0, tracks the reference crate version (Cargo.toml:3) */

typedef enum rcb_error {
    RCB_OK = 0,
    RCB_ERR_INVALID_ARGUMENT = -1,
    RCB_ERR_CUDA = -2,             /* a CUDA runtime call failed; see rcb_last_cuda_error */
    RCB_ERR_ZERO_TOTAL = -3,       /* total_freq == 0: divide-by-zero panic, src/range_coder.rs:39 */
    RCB_ERR_ZERO_FREQ_SYMBOL = -4, /* c_freq == 0 for a coded symbol: the reference never returns
                                      (range == 0 keeps loop 1 alive, src/range_coder.rs:83-85) */
    RCB_ERR_LOWER_OVERFLOW = -5,   /* RangeCoderError::LowerBoundOverflow, src/error.rs:6-10 */
    RCB_ERR_UPPER_OVERFLOW = -6,   /* RangeCoderError::UpperBoundOverflow, src/error.rs:12 */
    RCB_ERR_SYMBOL_OUT_OF_RANGE = -7, /* index >= alphabet: Vec::get().unwrap() panic,
                                         examples/sample_impl.rs:19 */
    RCB_ERR_OUT_CAPACITY = -8,     /* caller's output buffer too small */
    RCB_ERR_TRUNCATED_STREAM = -9, /* Decoder ran out of bytes: pop_front().unwrap(),
                                      src/decoder.rs:33 */
    RCB_ERR_INVALID_MODEL = -10,   /* table has cum_freq > total_freq (outside the path) */
    RCB_ERR_UNSUPPORTED = -11,
    RCB_ERR_NO_DEVICE = -12,       /* no CUDA device: there is no CPU fallback */
    RCB_ERR_NCCL = -13,            /* an NCCL call failed or libnccl.so.2 is missing; see rcb_comm_last_error */
    RCB_ERR_RESTART_POINT = -14    /* a restart point does not match the code stream (see rcb_restart_point) */
} rcb_error;

/* per-chunk status words written to d_status (0 = ok) */
enum {
    RCB_ST_OK = 0,
    RCB_ST_ZERO_FREQ = 1,
    RCB_ST_LOWER_OVERFLOW = 2,
    RCB_ST_UPPER_OVERFLOW = 3,
    RCB_ST_SYMBOL_RANGE = 4,
    RCB_ST_OUT_CAPACITY = 5,
    RCB_ST_TRUNCATED = 6,
    RCB_ST_RESTART = 7
};

/* model flags reported by rcb_model_info */
enum {
    RCB_MODEL_POW2 = 1,       /* total_freq is a power of two: range/total is a shift */
    RCB_MODEL_CONSISTENT = 2, /* cum+c <= total for every symbol: overflow errors unreachable */
    RCB_MODEL_REGULAR = 4     /* cum[i+1] == cum[i]+c[i]: table-driven decode lookup is sound */
};

typedef struct rcb_ctx rcb_ctx;
typedef struct rcb_model rcb_model;

const char *rcb_strerror(int err);
int rcb_version(void);
/* last cudaError_t seen by this ctx (0 = cudaSuccess) and its string */
int rcb_last_cuda_error(const rcb_ctx *ctx, const char **msg);

/* ---- context --------------------------------------------------------------
 * Replaces nothing in the reference (it has no runtime); owns the stream and
 * scratch.  `stream` is a cudaStream_t (NULL = the default stream). */
int rcb_device_count(void); /* CUDA devices visible to this process (0 when there is none) */
int rcb_ctx_create(int device, void *stream, rcb_ctx **out);
int rcb_ctx_destroy(rcb_ctx *ctx);
int rcb_ctx_set_stream(rcb_ctx *ctx, void *stream);
int rcb_ctx_synchronize(rcb_ctx *ctx);
/* tuning knobs (0 = default): threads per block of the coder kernels */
int rcb_ctx_set_block_threads(rcb_ctx *ctx, int encode_threads, int decode_threads);
/* kernels launched by this ctx so far (bench.py's gpu_launches) */
uint64_t rcb_ctx_launch_count(const rcb_ctx *ctx);
/* Per-kernel CUDA-event timing on the launching stream (off by default).
 * rcb_ctx_get_timings synchronises and fills ms[0..n): [0] encode kernel,
 * [1] length scan, [2] gather (compaction), [3] decode kernel, [4] decode
 * status summary of the most recent encode / decode issued with timing on. */
int rcb_ctx_enable_timing(rcb_ctx *ctx, int on);
int rcb_ctx_get_timings(rcb_ctx *ctx, float *ms, int n);

/* ---- device memory for hosts that carry no CUDA binding of their own (a Rust
 * caller sharding over several GPUs needs device-resident shards for
 * rcb_histogram / rcb_encode_chunks).  16-byte aligned; the copies run on the
 * ctx's stream and synchronise it. */
int rcb_device_alloc(rcb_ctx *ctx, uint64_t bytes, void **d_out);
int rcb_device_free(rcb_ctx *ctx, void *d_ptr);
int rcb_copy_to_device(rcb_ctx *ctx, void *d_dst, const void *h_src, uint64_t bytes);
int rcb_copy_to_host(rcb_ctx *ctx, void *h_dst, const void *d_src, uint64_t bytes);

/* ---- frequency model ------------------------------------------------------
 * Dense snapshot of a `PModel` (src/pmodel.rs:4-13): c_freq(i), cum_freq(i)
 * for i < K and total_freq().  PModel has no alphabet-size method, so K is
 * explicit.  n_models == 1: one static model shared by all chunks;
 * n_models == n_chunks: one model per chunk. */
int rcb_model_create(rcb_ctx *ctx, uint32_t K, uint64_t n_models, rcb_model **out);
int rcb_model_destroy(rcb_model *m);

/* FreqTable::add_alphabet_freq in a loop (examples/sample_impl.rs:58-60,78-80).
 * chunk_syms == 0: one histogram of all n_syms into d_counts = uint64[K].
 * chunk_syms  > 0: one histogram per chunk into d_counts = uint32[n_chunks][K].
 * Symbols are u8 (sym_bytes 1) or u16 little-endian (sym_bytes 2); a symbol
 * >= K is not counted and makes the call return RCB_ERR_SYMBOL_OUT_OF_RANGE. */
int rcb_histogram(rcb_ctx *ctx, const void *d_syms, uint64_t n_syms, int sym_bytes,
                  uint32_t K, uint64_t chunk_syms, void *d_counts);

/* FreqTable::calc_cum (examples/sample_impl.rs:61-69): c = counts, cum =
 * exclusive prefix sum, total = sum.  count_bytes is 8 (uint64[n_models][K]) or
 * 4 (uint32[n_models][K]).  When a u64 sum exceeds 2^32-1 (total_freq is u32,
 * src/pmodel.rs:10, and the reference has no rule for that) the counts are
 * rescaled to total = 2^31: c' = c ? max(1, floor(c * 2^31 / sum)) : 0, the
 * rounding difference going to the largest count (build-defined
 * extension; identical in the oracle).  Synchronises. */
int rcb_model_from_counts(rcb_ctx *ctx, rcb_model *m, const void *d_counts, int count_bytes);

/* Snapshot of arbitrary PModel tables from host memory:
 * h_c, h_cum: uint32[n_models][K]; h_total: uint32[n_models].  Synchronises. */
int rcb_model_from_tables(rcb_ctx *ctx, rcb_model *m, const uint32_t *h_c,
                          const uint32_t *h_cum, const uint32_t *h_total);

/* Read model `index` back (any pointer may be NULL). */
int rcb_model_get_tables(rcb_ctx *ctx, const rcb_model *m, uint64_t index, uint32_t *h_c,
                         uint32_t *h_cum, uint32_t *h_total, uint32_t *h_flags);

/* ---- encode: replaces the caller's loop over Encoder::encode + finish ------
 * (examples/sample_impl.rs:92-98 -> src/encoder.rs:24-46 ->
 * src/range_coder.rs:53-135).  Output: d_out[0 .. d_offsets[n_chunks]) and
 * d_offsets = uint64[n_chunks+1]; d_status = uint32[n_chunks] (may be NULL).
 * h_out_bytes receives d_offsets[n_chunks].  Returns the first chunk error
 * mapped to rcb_error, or RCB_ERR_OUT_CAPACITY if out_cap is too small (then
 * *h_out_bytes is the size needed). */
int rcb_encode_chunks(rcb_ctx *ctx, const void *d_syms, uint64_t n_syms, int sym_bytes,
                      uint64_t chunk_syms, const rcb_model *m, uint8_t *d_out,
                      uint64_t out_cap, uint64_t *d_offsets, uint32_t *d_status,
                      uint64_t *h_out_bytes);
/* same, no synchronisation and no host result (errors stay in d_status);
 * call rcb_encode_result later to fetch them. */
int rcb_encode_chunks_async(rcb_ctx *ctx, const void *d_syms, uint64_t n_syms, int sym_bytes,
                            uint64_t chunk_syms, const rcb_model *m, uint8_t *d_out,
                            uint64_t out_cap, uint64_t *d_offsets, uint32_t *d_status);
int rcb_encode_result(rcb_ctx *ctx, uint64_t *h_out_bytes);
/* an out_cap that always suffices for this model / shape */
uint64_t rcb_encode_bound(rcb_ctx *ctx, const rcb_model *m, uint64_t n_syms, int sym_bytes,
                          uint64_t chunk_syms);

/* ---- decode: replaces Decoder::new + the loop over Decoder::decode ---------
 * (examples/sample_impl.rs:110-120 -> src/decoder.rs:14-54, with
 * examples/sample_impl.rs:27-45 as the built-in find_index).  The symbol count
 * is out-of-band exactly like in the reference. d_stream must be readable up
 * to the next multiple of 16 bytes past d_offsets[n_chunks]. */
int rcb_decode_chunks(rcb_ctx *ctx, const uint8_t *d_stream, const uint64_t *d_offsets,
                      uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                      const rcb_model *m, void *d_syms_out, uint32_t *d_status);
int rcb_decode_chunks_async(rcb_ctx *ctx, const uint8_t *d_stream, const uint64_t *d_offsets,
                            uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                            const rcb_model *m, void *d_syms_out, uint32_t *d_status);
int rcb_decode_result(rcb_ctx *ctx);

/* ---- restart points: several decoder lanes per chunk ------------------------
 * Not in the reference (its Decoder walks one stream front to back,
 * src/decoder.rs:38-54); side information next to the offsets, the code bytes
 * stay exactly the reference's.  The reference decoder keeps the encoder's
 * (lower_bound, range) (src/decoder.rs:42-52 runs the same param_update) and its
 * `data` is the 8 code bytes at the current position (src/decoder.rs:31-35), so
 * the state in front of symbol j of a chunk is (lower_bound, range, code bytes
 * emitted so far) -- which the ENCODER knows.  rcb_encode_chunks_restart records
 * it every restart_syms symbols; rcb_decode_chunks_restart then decodes each
 * chunk with ceil(chunk_syms / restart_syms) lanes instead of one (64 KiB chunks
 * are too few lanes to fill a B200).  A lane that arrives at
 * the next record in a different state reports RCB_ST_RESTART, so a damaged
 * record or stream is detected, not silently decoded.
 *   d_restart = rcb_restart_point[n_chunks][rcb_restart_points_per_chunk()],
 *   record r of chunk i = state in front of symbol (r + 1) * restart_syms of it
 *   (range == 0: absent, ragged last chunk).  restart_syms: multiple of 64, at
 *   most 64 parts per chunk; restart_syms == 0 or d_restart == NULL = the plain
 *   calls above.  `range` is stored rounded down to a multiple of total_freq (only
 *   range / total_freq is used before the next update, src/range_coder.rs:62). */
typedef struct rcb_restart_point {
    uint64_t lower_bound; /* RangeCoder::lower_bound, src/range_coder.rs:9  */
    uint64_t range;       /* RangeCoder::range (src/range_coder.rs:11) / total_freq * total_freq */
    uint32_t code_bytes;  /* bytes Encoder::encode has returned so far (src/encoder.rs:24-37) */
    uint32_t reserved;
} rcb_restart_point;
uint64_t rcb_restart_points_per_chunk(uint64_t chunk_syms, uint64_t restart_syms);
int rcb_encode_chunks_restart(rcb_ctx *ctx, const void *d_syms, uint64_t n_syms, int sym_bytes,
                              uint64_t chunk_syms, const rcb_model *m, uint8_t *d_out,
                              uint64_t out_cap, uint64_t *d_offsets, uint32_t *d_status,
                              uint64_t restart_syms, rcb_restart_point *d_restart,
                              uint64_t *h_out_bytes);
int rcb_encode_chunks_restart_async(rcb_ctx *ctx, const void *d_syms, uint64_t n_syms, int sym_bytes,
                                    uint64_t chunk_syms, const rcb_model *m, uint8_t *d_out,
                                    uint64_t out_cap, uint64_t *d_offsets, uint32_t *d_status,
                                    uint64_t restart_syms, rcb_restart_point *d_restart);
int rcb_decode_chunks_restart(rcb_ctx *ctx, const uint8_t *d_stream, const uint64_t *d_offsets,
                              uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                              const rcb_model *m, void *d_syms_out, uint32_t *d_status,
                              uint64_t restart_syms, const rcb_restart_point *d_restart);
int rcb_decode_chunks_restart_async(rcb_ctx *ctx, const uint8_t *d_stream, const uint64_t *d_offsets,
                                    uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                                    const rcb_model *m, void *d_syms_out, uint32_t *d_status,
                                    uint64_t restart_syms, const rcb_restart_point *d_restart);

/* ---- host-buffer convenience (H2D, encode/decode, D2H inside the call) -----
 * What a Rust `gpu::encode_chunks(&pmodel, K, &symbols, chunk_syms)` binds.
 * Large batches run as a pipeline of up to 8 slices of whole chunks (copies in
 * and out overlap the kernels; rcb_decode_host additionally decodes chunks of
 * >= 32 KiB in four resumable launches) -- results do not depend on that.
 * Pinned (page-locked) host buffers let the copies run asynchronously; pageable
 * ones work and are slower.  Exactly h_offsets[n_chunks] bytes of h_stream are
 * read (the padding the device reader wants is added on the device side).
 * RCB_TRACE=1 in the environment prints the pipeline's timeline on stderr. */
int rcb_encode_host(rcb_ctx *ctx, const void *h_syms, uint64_t n_syms, int sym_bytes,
                    uint64_t chunk_syms, const rcb_model *m, uint8_t *h_out, uint64_t out_cap,
                    uint64_t *h_offsets, uint64_t *h_out_bytes);
int rcb_decode_host(rcb_ctx *ctx, const uint8_t *h_stream, const uint64_t *h_offsets,
                    uint64_t n_syms, int sym_bytes, uint64_t chunk_syms, const rcb_model *m,
                    void *h_syms_out);
/* the same with restart points in host memory (h_restart = rcb_restart_point[n_chunks][per chunk],
 * written by the encoder, read by the decoder; see "restart points" below) */
int rcb_encode_host_restart(rcb_ctx *ctx, const void *h_syms, uint64_t n_syms, int sym_bytes,
                            uint64_t chunk_syms, const rcb_model *m, uint8_t *h_out, uint64_t out_cap,
                            uint64_t *h_offsets, uint64_t *h_out_bytes, uint64_t restart_syms,
                            rcb_restart_point *h_restart);
int rcb_decode_host_restart(rcb_ctx *ctx, const uint8_t *h_stream, const uint64_t *h_offsets,
                            uint64_t n_syms, int sym_bytes, uint64_t chunk_syms, const rcb_model *m,
                            void *h_syms_out, uint64_t restart_syms,
                            const rcb_restart_point *h_restart);

/* ---- continued single-stream coder: the reference's per-symbol API ----------
 * Encoder / Decoder keep (lower_bound, range[, data]) between calls
 * (src/encoder.rs:7-11, src/decoder.rs:6-12).  These entry points code a slice
 * of symbols on one GPU lane from a caller-held state and return the new state,
 * so a host mirror of Encoder::encode / Decoder::decode (include/rcb200.hpp)
 * needs no CPU arithmetic.  Host pointers; shared model (n_models == 1) only. */
typedef struct rcb_stream_state {
    uint64_t lower_bound; /* RangeCoder::lower_bound, src/range_coder.rs:9  */
    uint64_t range;       /* RangeCoder::range,       src/range_coder.rs:11 */
    uint64_t data;        /* Decoder::data,           src/decoder.rs:9      */
    uint64_t consumed;    /* Decoder: code bytes shifted into data (0 = Decoder::new not run yet) */
    uint32_t status;      /* RCB_ST_* of the last call */
    uint32_t pad;
} rcb_stream_state;
/* RangeCoder::new (src/range_coder.rs:13-20): lower 0, range u64::MAX */
void rcb_stream_state_init(rcb_stream_state *st);
/* Encoder::encode for n symbols (src/encoder.rs:24-37); finish != 0 appends
 * Encoder::finish (src/encoder.rs:40-46).  h_per_symbol (may be NULL) gets each
 * symbol's emitted byte count -- encode()'s return value.  m may be NULL when
 * n_syms == 0 (finish alone). */
int rcb_encode_stream(rcb_ctx *ctx, rcb_stream_state *st, const void *h_syms, uint64_t n_syms,
                      int sym_bytes, const rcb_model *m, uint8_t *h_out, uint64_t out_cap,
                      uint64_t *h_n_out, uint32_t *h_per_symbol, int finish);
/* Decoder::new on the first call, then Decoder::decode n times (src/decoder.rs:14-54). */
int rcb_decode_stream(rcb_ctx *ctx, rcb_stream_state *st, const uint8_t *h_code, uint64_t code_len,
                      uint64_t n_syms, int sym_bytes, const rcb_model *m, void *h_syms_out);

/* ---- framed container (the reference has no wire format:
 * Encoder::finish returns raw bytes and the symbol count / model travel out of
 * band, src/encoder.rs:40-46, examples/sample_impl.rs:113).  Little-endian:
 *   "RCB2" | version u32 | sym_bytes u32 | K u32 | model_mode u32 (0 shared, 1 per chunk)
 *   | chunk_syms u64 | n_syms u64 | n_chunks u64 | payload_bytes u64
 *   | model: shared    -> total u32, reserved u32, cum[K] u32, c[K] u32
 *            per chunk -> c[n_chunks][K] u32 (cum = exclusive scan, total = sum)
 *   | offsets u64[n_chunks+1]
 *   | version 2 only: restart points, rcb_restart_point[n_chunks][per chunk] (24-byte records,
 *     little-endian), restart_syms = 64 * the u32 at byte 20 of the header (0 in version 1)
 *   | payload (chunk i = payload[offsets[i]..offsets[i+1]), exactly the reference's finish()
 *     bytes for that chunk)
 * (a u32 of padding follows model_mode: byte 20, the restart field.)
 * Pure host-side byte layout; all coding stays in the entry points above. */
typedef struct rcb_frame_info {
    uint32_t version, sym_bytes, K, model_mode;
    uint64_t chunk_syms, n_syms, n_chunks, payload_bytes;
    uint64_t model_off, offsets_off, payload_off, frame_bytes;
    uint64_t restart_syms, restart_off; /* 0, 0: no restart section (version 1) */
} rcb_frame_info;
uint64_t rcb_frame_bound(uint32_t K, uint64_t n_chunks, int per_chunk, uint64_t payload_bytes);
uint64_t rcb_frame_bound_restart(uint32_t K, uint64_t n_chunks, int per_chunk, uint64_t payload_bytes,
                                 uint64_t chunk_syms, uint64_t restart_syms);
/* model tables are read back from the device; h_stream/h_offsets as produced by rcb_encode_host */
int rcb_frame_write(rcb_ctx *ctx, const rcb_model *m, int sym_bytes, uint64_t chunk_syms, uint64_t n_syms,
                    const uint8_t *h_stream, const uint64_t *h_offsets, uint8_t *h_frame,
                    uint64_t frame_cap, uint64_t *h_frame_bytes);
/* ... with the restart points rcb_encode_host_restart produced (version 2 frame) */
int rcb_frame_write_restart(rcb_ctx *ctx, const rcb_model *m, int sym_bytes, uint64_t chunk_syms,
                            uint64_t n_syms, const uint8_t *h_stream, const uint64_t *h_offsets,
                            uint64_t restart_syms, const rcb_restart_point *h_restart,
                            uint8_t *h_frame, uint64_t frame_cap, uint64_t *h_frame_bytes);
/* validates the header and section sizes against len */
int rcb_frame_parse(const uint8_t *h_frame, uint64_t len, rcb_frame_info *info);
/* device model from the frame's model section (caller destroys it) */
int rcb_frame_model(rcb_ctx *ctx, const uint8_t *h_frame, const rcb_frame_info *info, rcb_model **out);
/* encode_host + frame_write / parse + model + decode_host */
int rcb_frame_encode_host(rcb_ctx *ctx, const void *h_syms, uint64_t n_syms, int sym_bytes,
                          uint64_t chunk_syms, const rcb_model *m, uint8_t *h_frame, uint64_t frame_cap,
                          uint64_t *h_frame_bytes);
/* rcb_frame_decode_host reads either version and uses a restart section when the frame has one */
int rcb_frame_encode_host_restart(rcb_ctx *ctx, const void *h_syms, uint64_t n_syms, int sym_bytes,
                                  uint64_t chunk_syms, const rcb_model *m, uint64_t restart_syms,
                                  uint8_t *h_frame, uint64_t frame_cap, uint64_t *h_frame_bytes);
int rcb_frame_decode_host(rcb_ctx *ctx, const uint8_t *h_frame, uint64_t len, void *h_syms_out,
                          uint64_t out_cap_bytes, uint64_t *h_n_syms);

/* ---- adaptive-per-symbol model --------------------------------
 * The reference takes `&T: PModel` on every call (src/encoder.rs:24,
 * src/decoder.rs:38): its caller may change the table between symbols.  The
 * build defines one such model (DESIGN.md section 5 spells out the caller-side
 * loop it stands for): counts start at 1, the coded symbol
 * gains `inc` AFTER it was coded, all counts are halved (rounding up) when the
 * total would pass `limit`; cum_freq = exclusive prefix sum, total_freq = sum.
 * Every chunk restarts the coder and the table, so chunk i's bytes are what
 * the reference's Encoder emits for that call sequence on chunk i's symbols.
 * Device pointers, same conventions as rcb_encode_chunks / rcb_decode_chunks.
 * Limits of this first path: K <= 4096, K <= limit, limit + inc <= 65535. */
typedef struct rcb_adaptive_params {
    uint32_t K;     /* alphabet size */
    uint32_t inc;   /* added to the coded symbol's count (>= 1) */
    uint32_t limit; /* halve all counts when total + inc would exceed this */
} rcb_adaptive_params;
uint64_t rcb_adaptive_encode_bound(const rcb_adaptive_params *p, uint64_t n_syms, uint64_t chunk_syms);
int rcb_adaptive_encode_chunks(rcb_ctx *ctx, const void *d_syms, uint64_t n_syms, int sym_bytes,
                               uint64_t chunk_syms, const rcb_adaptive_params *p, uint8_t *d_out,
                               uint64_t out_cap, uint64_t *d_offsets, uint32_t *d_status,
                               uint64_t *h_out_bytes);
int rcb_adaptive_decode_chunks(rcb_ctx *ctx, const uint8_t *d_stream, const uint64_t *d_offsets,
                               uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                               const rcb_adaptive_params *p, void *d_syms_out, uint32_t *d_status);

/* ---- multi-GPU: the path's only exchange step -----------------
 * Chunks shard over GPUs with no data-path collective.  A static model shared by
 * chunks on several GPUs needs the sum of the per-GPU count tables -- in the
 * reference the caller owns that loop (FreqTable::add_alphabet_freq over all the
 * data, examples/sample_impl.rs:77-81) -- which is ONE ncclAllReduce(sum) of
 * uint64[K] over NVLink; every rank then runs the same deterministic
 * rcb_model_from_counts and holds an identical table.  NCCL is bound at run time
 * (dlopen libnccl.so.2, or $RCB_NCCL_LIB), so hosts that never shard need none.
 *   one process / thread per GPU:  rank 0 calls rcb_comm_unique_id and ships the
 *     128 bytes to the others out of band; every rank calls rcb_comm_init_rank.
 *   one thread driving all GPUs:   rcb_comm_init_all + rcb_allreduce_counts_multi. */
#define RCB_UNIQUE_ID_BYTES 128
typedef struct rcb_comm rcb_comm;
int rcb_comm_unique_id(uint8_t *id /*[RCB_UNIQUE_ID_BYTES]*/);
int rcb_comm_init_rank(rcb_ctx *ctx, const uint8_t *id, int n_ranks, int rank, rcb_comm **out);
int rcb_comm_init_all(rcb_ctx *const *ctxs, int n_ctx, rcb_comm **out /*[n_ctx]*/);
int rcb_comm_destroy(rcb_comm *comm);
int rcb_comm_info(const rcb_comm *comm, int *n_ranks, int *rank, int *nccl_version);
const char *rcb_comm_last_error(const rcb_comm *comm);
/* d_counts = uint64[K] as written by rcb_histogram(chunk_syms == 0): summed over
 * the communicator's ranks, in place, on the ctx's stream; does not synchronise
 * (rcb_model_from_counts on the same ctx is ordered after it). */
int rcb_allreduce_counts(rcb_ctx *ctx, rcb_comm *comm, void *d_counts, uint32_t K);
int rcb_allreduce_counts_multi(rcb_ctx *const *ctxs, rcb_comm *const *comms, void *const *d_counts,
                               uint32_t K, int n_ctx);

/* ---- synthetic data (benchmark inputs; not in the reference)
 * symbol j = #{ i < K-1 : thr[t][i] <= mix64(seed + j*GOLDEN) >> 32 },
 * t = (j / chunk_syms) % n_tables; h_thr = uint32[n_tables][K-1]. */
int rcb_generate(rcb_ctx *ctx, void *d_out, uint64_t first, uint64_t n, int sym_bytes,
                 uint32_t K, uint64_t seed, const uint32_t *h_thr, uint32_t n_tables,
                 uint64_t chunk_syms);

#ifdef __cplusplus
}
#endif
#endif /* RCB200_H */
