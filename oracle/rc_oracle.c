/*
 * rc_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See rc_oracle.h for scope, the "PARITY UNPINNED" statement and who may call this.
 *
 * Every function restates the cited lines of /root/reference; loops are kept as
 * loops (no closed forms, no reciprocal) so that the CUDA path's closed forms
 * are checked against an independent formulation.
 */
#include "rc_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define TOP8 (1ull << (64 - 8))   /* src/range_coder.rs:23 */
#define TOP16 (1ull << (64 - 16)) /* src/range_coder.rs:24 */

/* src/range_coder.rs:13-20 */
void rco_rc_new(rco_range_coder *rc) {
    rc->lower_bound = 0;
    rc->range = UINT64_MAX;
}

/* src/range_coder.rs:38-40 */
uint64_t rco_range_par_total(const rco_range_coder *rc, uint32_t total_freq) {
    return rc->range / (uint64_t)total_freq;
}

/* src/range_coder.rs:138-146 */
int rco_upper_bound(const rco_range_coder *rc, uint64_t *upper) {
    uint64_t u = rc->lower_bound + rc->range;
    if (u < rc->lower_bound) return RCO_ERR_UPPER_OVERFLOW;
    *upper = u;
    return RCO_OK;
}

/* src/range_coder.rs:95-100 */
static uint8_t left_shift(rco_range_coder *rc) {
    uint8_t tmp = (uint8_t)(rc->lower_bound >> (64 - 8));
    rc->range <<= 8;
    rc->lower_bound <<= 8;
    return tmp;
}

/* src/range_coder.rs:110-116: 1 = byte produced, 0 = none, <0 = error */
static int no_carry_expansion(rco_range_coder *rc, uint8_t *byte) {
    uint64_t upper;
    int e = rco_upper_bound(rc, &upper); /* .unwrap() => panic */
    if (e) return e;
    if ((rc->lower_bound ^ upper) < TOP8) {
        *byte = left_shift(rc);
        return 1;
    }
    return 0;
}

/* src/range_coder.rs:126-135 */
static int range_reduction_expansion(rco_range_coder *rc, uint8_t *byte) {
    if (rc->range < TOP16) {
        uint64_t range_new = ~rc->lower_bound & (TOP16 - 1);
        rc->range = range_new;
        *byte = left_shift(rc);
        return 1;
    }
    return 0;
}

/* src/range_coder.rs:53-92 */
int rco_param_update(rco_range_coder *rc, uint32_t c_freq, uint32_t cum_freq,
                     uint32_t total_freq, uint8_t *out, int *n_out) {
    int n = 0, r;
    uint8_t b = 0;
    *n_out = 0;
    if (total_freq == 0) return RCO_ERR_ZERO_TOTAL; /* :39 div-by-zero panic */
    uint64_t range_par_total = rco_range_par_total(rc, total_freq); /* :62 */
    rc->range = range_par_total * (uint64_t)c_freq;                 /* :65 */
    uint64_t add = range_par_total * (uint64_t)cum_freq;            /* :70 */
    uint64_t nl = rc->lower_bound + add;
    if (nl < rc->lower_bound) return RCO_ERR_LOWER_OVERFLOW; /* :74-80 */
    rc->lower_bound = nl;
    /* The reference never terminates loop 1 when range == 0 (c_freq == 0):
     * lower ^ upper == 0 < TOP8 forever (:83-85,110-112).  Report instead. */
    if (rc->range == 0) return RCO_ERR_ZERO_FREQ;
    while ((r = no_carry_expansion(rc, &b)) > 0) { /* :83-85 */
        if (n >= 16) return RCO_ERR_ZERO_FREQ;
        out[n++] = b;
    }
    if (r < 0) return r;
    while (range_reduction_expansion(rc, &b) > 0) { /* :87-89 */
        if (n >= 16) return RCO_ERR_ZERO_FREQ;
        out[n++] = b;
    }
    *n_out = n;
    return RCO_OK;
}

static inline uint32_t load_sym(const void *syms, uint64_t i, int sym_bytes) {
    if (sym_bytes == 1) return ((const uint8_t *)syms)[i];
    if (sym_bytes == 2) {
        const uint8_t *p = (const uint8_t *)syms + 2 * i; /* u16 little-endian */
        return (uint32_t)p[0] | ((uint32_t)p[1] << 8);
    }
    const uint8_t *p = (const uint8_t *)syms + 4 * i;
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) |
           ((uint32_t)p[3] << 24);
}

static inline void store_sym(void *syms, uint64_t i, int sym_bytes, uint32_t v) {
    if (sym_bytes == 1) {
        ((uint8_t *)syms)[i] = (uint8_t)v;
    } else if (sym_bytes == 2) {
        uint8_t *p = (uint8_t *)syms + 2 * i;
        p[0] = (uint8_t)v;
        p[1] = (uint8_t)(v >> 8);
    } else {
        uint8_t *p = (uint8_t *)syms + 4 * i;
        p[0] = (uint8_t)v;
        p[1] = (uint8_t)(v >> 8);
        p[2] = (uint8_t)(v >> 16);
        p[3] = (uint8_t)(v >> 24);
    }
}

/* src/encoder.rs:24-37 (encode), :40-46 (finish) */
int64_t rco_encode(const void *syms, uint64_t n, int sym_bytes, uint32_t K,
                   const uint32_t *c, const uint32_t *cum, uint32_t total,
                   uint8_t *out, uint64_t cap) {
    rco_range_coder rc;
    rco_rc_new(&rc); /* src/encoder.rs:48-54 */
    uint64_t len = 0;
    uint8_t tmp[16];
    for (uint64_t i = 0; i < n; i++) {
        uint32_t index = load_sym(syms, i, sym_bytes);
        if (index >= K) return RCO_ERR_SYMBOL_RANGE;
        int k;
        int e = rco_param_update(&rc, c[index], cum[index], total, tmp, &k); /* :27-33 */
        if (e) return e;
        if (len + (uint64_t)k > cap) return RCO_ERR_CAPACITY;
        memcpy(out + len, tmp, (size_t)k); /* :35 append in emission order */
        len += (uint64_t)k;
    }
    if (len + 8 > cap) return RCO_ERR_CAPACITY;
    for (int i = 0; i < 8; i++) out[len++] = left_shift(&rc); /* :42-44 */
    return (int64_t)len;
}

/* The Encoder's state after n symbols and before finish(): (lower_bound, range) of its RangeCoder
 * (src/encoder.rs:7-11, src/range_coder.rs:7-12) and the number of bytes encode() has returned so far
 * (src/encoder.rs:24-37).  Checker for the restart points of the GPU path (include/rcb200.h). */
int64_t rco_encode_state(const void *syms, uint64_t n, int sym_bytes, uint32_t K,
                         const uint32_t *c, const uint32_t *cum, uint32_t total,
                         uint64_t *lower, uint64_t *range) {
    rco_range_coder rc;
    rco_rc_new(&rc);
    uint64_t len = 0;
    uint8_t tmp[16];
    for (uint64_t i = 0; i < n; i++) {
        uint32_t index = load_sym(syms, i, sym_bytes);
        if (index >= K) return RCO_ERR_SYMBOL_RANGE;
        int k;
        int e = rco_param_update(&rc, c[index], cum[index], total, tmp, &k);
        if (e) return e;
        len += (uint64_t)k;
    }
    *lower = rc.lower_bound;
    *range = rc.range;
    return (int64_t)len;
}

/* src/decoder.rs:6-12 */
typedef struct {
    rco_range_coder range_coder;
    uint64_t data;
    const uint8_t *buffer;
    uint64_t pos, len;
} rco_decoder;

/* src/decoder.rs:31-35 */
static int shift_left_buffer(rco_decoder *d, int n) {
    for (int i = 0; i < n; i++) {
        if (d->pos >= d->len) return RCO_ERR_TRUNCATED; /* pop_front().unwrap() */
        d->data = (d->data << 8) | (uint64_t)d->buffer[d->pos++];
    }
    return RCO_OK;
}

/* examples/sample_impl.rs:27-45 */
static uint32_t find_index(const rco_decoder *d, uint32_t K, const uint32_t *cum,
                           uint32_t total) {
    uint64_t rfreq = (d->data - d->range_coder.lower_bound) /
                     rco_range_par_total(&d->range_coder, total); /* :29-30 */
    uint32_t left = 0, right = K - 1;                              /* :33-34 */
    while (left < right) {
        uint32_t mid = (left + right) / 2;
        uint32_t mid_cum = cum[mid + 1];
        if ((uint64_t)mid_cum <= rfreq)
            left = mid + 1;
        else
            right = mid;
    }
    return left;
}

/* src/decoder.rs:14-23 (new), :38-54 (decode) */
int64_t rco_decode(const uint8_t *code, uint64_t len, uint64_t n_syms,
                   int sym_bytes, uint32_t K, const uint32_t *c,
                   const uint32_t *cum, uint32_t total, void *out_syms) {
    rco_decoder d;
    rco_rc_new(&d.range_coder);
    d.data = 0;
    d.buffer = code;
    d.pos = 0;
    d.len = len;
    int e = shift_left_buffer(&d, 8); /* :21 */
    if (e) return e;
    if (n_syms && total == 0) return RCO_ERR_ZERO_TOTAL;
    uint8_t tmp[16];
    for (uint64_t i = 0; i < n_syms; i++) {
        /* range/total == 0 cannot happen (range >= 2^48 > total) */
        uint32_t idx = find_index(&d, K, cum, total); /* :40 */
        int n;
        e = rco_param_update(&d.range_coder, c[idx], cum[idx], total, tmp, &n); /* :42-50 */
        if (e) return e;
        e = shift_left_buffer(&d, n); /* :52 */
        if (e) return e;
        store_sym(out_syms, i, sym_bytes, idx);
    }
    return (int64_t)d.pos;
}

/* examples/sample_impl.rs:58-60, loop at :78-80 */
void rco_histogram(const void *syms, uint64_t n, int sym_bytes, uint32_t K,
                   uint64_t *counts) {
    memset(counts, 0, sizeof(uint64_t) * K);
    for (uint64_t i = 0; i < n; i++) {
        uint32_t s = load_sym(syms, i, sym_bytes);
        if (s < K) counts[s] += 1;
    }
}

/* examples/sample_impl.rs:61-69 */
uint32_t rco_calc_cum(const uint32_t *c, uint32_t K, uint32_t *cum) {
    uint32_t cum_total = 0;
    for (uint32_t i = 0; i < K; i++) {
        cum[i] = cum_total;
        cum_total += c[i]; /* u32 wrapping, as release-mode Rust */
    }
    return cum_total;
}

/* Build-defined extension; identity == the reference whenever it applies.
 * sum > 2^32-1: rescale to total = 2^31 exactly: c' = c ? max(1, floor(c*2^31/sum)) : 0, then
 * the difference 2^31 - sum(c') goes to the largest count (lowest index on ties).
 * Returns 1 when the counts were rescaled, 0 for the identity. */
int rco_normalise(const uint64_t *counts, uint32_t K, uint32_t *c) {
    unsigned __int128 sum = 0;
    uint64_t vmax = 0;
    uint32_t imax = 0;
    for (uint32_t i = 0; i < K; i++) {
        sum += counts[i];
        if (counts[i] > vmax) {
            vmax = counts[i];
            imax = i;
        }
    }
    if (sum <= 0xFFFFFFFFull) {
        for (uint32_t i = 0; i < K; i++) c[i] = (uint32_t)counts[i];
        return 0;
    }
    int64_t scaled = 0;
    for (uint32_t i = 0; i < K; i++) {
        uint32_t v = 0;
        if (counts[i]) {
            unsigned __int128 q = ((unsigned __int128)counts[i] << 31) / sum;
            v = (uint32_t)q;
            if (v == 0) v = 1;
        }
        c[i] = v;
        scaled += v;
    }
    c[imax] = (uint32_t)((int64_t)c[imax] + ((int64_t)1 << 31) - scaled);
    return 1;
}

/* ------------------------------------------------------------------------- */
/* f4: adaptive-per-symbol table (build-defined; see rc_oracle.h).  Written as the caller-side loop  */
/* of the reference API: the table is a FreqTable (examples/sample_impl.rs:10-70) whose counts the     */
/* caller bumps between Encoder::encode / Decoder::decode calls, calc_cum() after every change.        */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint32_t K, inc, limit, total;
    uint32_t *c, *cum;
} adaptive_table;

static int adaptive_init(adaptive_table *t, uint32_t K, uint32_t inc, uint32_t limit) {
    if (K == 0 || inc == 0 || limit < K || (uint64_t)limit + inc > 0xFFFFFFFFull) return -1;
    t->K = K;
    t->inc = inc;
    t->limit = limit;
    t->c = (uint32_t *)malloc(sizeof(uint32_t) * K);
    t->cum = (uint32_t *)malloc(sizeof(uint32_t) * K);
    if (!t->c || !t->cum) return -1;
    for (uint32_t i = 0; i < K; i++) t->c[i] = 1; /* every symbol codable from the start */
    t->total = rco_calc_cum(t->c, K, t->cum);
    return 0;
}
static void adaptive_free(adaptive_table *t) {
    free(t->c);
    free(t->cum);
}
/* after symbol s was coded with the old table */
static void adaptive_update(adaptive_table *t, uint32_t s) {
    t->c[s] += t->inc;
    if ((uint64_t)t->total + t->inc > t->limit)
        for (uint32_t i = 0; i < t->K; i++) t->c[i] = (t->c[i] + 1) >> 1; /* stays >= 1 */
    t->total = rco_calc_cum(t->c, t->K, t->cum); /* examples/sample_impl.rs:61-69 */
}

int64_t rco_adaptive_encode(const void *syms, uint64_t n, int sym_bytes, uint32_t K, uint32_t inc,
                            uint32_t limit, uint8_t *out, uint64_t cap) {
    adaptive_table t;
    if (adaptive_init(&t, K, inc, limit)) return RCO_ERR_ZERO_TOTAL;
    rco_range_coder rc;
    rco_rc_new(&rc);
    uint64_t len = 0;
    uint8_t tmp[16];
    int64_t ret = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t index = load_sym(syms, i, sym_bytes);
        if (index >= K) { ret = RCO_ERR_SYMBOL_RANGE; break; }
        int k;
        int e = rco_param_update(&rc, t.c[index], t.cum[index], t.total, tmp, &k); /* src/encoder.rs:27-33 */
        if (e) { ret = e; break; }
        if (len + (uint64_t)k > cap) { ret = RCO_ERR_CAPACITY; break; }
        memcpy(out + len, tmp, (size_t)k);
        len += (uint64_t)k;
        adaptive_update(&t, index);
    }
    adaptive_free(&t);
    if (ret) return ret;
    if (len + 8 > cap) return RCO_ERR_CAPACITY;
    for (int i = 0; i < 8; i++) out[len++] = left_shift(&rc); /* src/encoder.rs:42-44 */
    return (int64_t)len;
}

int64_t rco_adaptive_decode(const uint8_t *code, uint64_t len, uint64_t n_syms, int sym_bytes,
                            uint32_t K, uint32_t inc, uint32_t limit, void *out_syms) {
    adaptive_table t;
    if (adaptive_init(&t, K, inc, limit)) return RCO_ERR_ZERO_TOTAL;
    rco_decoder d;
    rco_rc_new(&d.range_coder);
    d.data = 0;
    d.buffer = code;
    d.pos = 0;
    d.len = len;
    int e = shift_left_buffer(&d, 8); /* src/decoder.rs:21 */
    uint8_t tmp[16];
    for (uint64_t i = 0; i < n_syms && !e; i++) {
        uint32_t idx = find_index(&d, K, t.cum, t.total); /* examples/sample_impl.rs:27-45 on the live table */
        int n;
        e = rco_param_update(&d.range_coder, t.c[idx], t.cum[idx], t.total, tmp, &n);
        if (e) break;
        e = shift_left_buffer(&d, n);
        if (e) break;
        store_sym(out_syms, i, sym_bytes, idx);
        adaptive_update(&t, idx);
    }
    adaptive_free(&t);
    return e ? e : (int64_t)d.pos;
}

/* ------------------------------------------------------------------------- */
/* chunked multi-thread drivers: "one chunk per thread at a time"            */
/* ------------------------------------------------------------------------- */

typedef struct {
    int decode;
    const uint8_t *syms_in;
    uint8_t *syms_out;
    uint64_t n;
    int sym_bytes;
    uint64_t chunk_syms, n_chunks;
    uint32_t K;
    const uint32_t *c, *cum, *total;
    int per_chunk_model;
    uint8_t *out;
    uint64_t out_pitch;
    const uint8_t *stream;
    const uint64_t *offsets;
    int64_t *res;
    atomic_ullong next;
    int adaptive; /* f4: table restarts with the chunk and follows the symbols */
    uint32_t inc, limit;
} chunk_job;

static void *chunk_worker(void *arg) {
    chunk_job *j = (chunk_job *)arg;
    for (;;) {
        uint64_t i = atomic_fetch_add(&j->next, 1);
        if (i >= j->n_chunks) break;
        uint64_t first = i * j->chunk_syms;
        uint64_t cnt = j->n - first < j->chunk_syms ? j->n - first : j->chunk_syms;
        if (j->adaptive) {
            if (!j->decode)
                j->res[i] = rco_adaptive_encode(j->syms_in + first * j->sym_bytes, cnt, j->sym_bytes, j->K,
                                                j->inc, j->limit, j->out + i * j->out_pitch, j->out_pitch);
            else
                j->res[i] = rco_adaptive_decode(j->stream + j->offsets[i], j->offsets[i + 1] - j->offsets[i],
                                                cnt, j->sym_bytes, j->K, j->inc, j->limit,
                                                j->syms_out + first * j->sym_bytes);
            continue;
        }
        const uint32_t *c = j->per_chunk_model ? j->c + i * j->K : j->c;
        const uint32_t *cum = j->per_chunk_model ? j->cum + i * j->K : j->cum;
        uint32_t total = j->per_chunk_model ? j->total[i] : j->total[0];
        if (!j->decode) {
            j->res[i] = rco_encode(j->syms_in + first * j->sym_bytes, cnt, j->sym_bytes,
                                   j->K, c, cum, total, j->out + i * j->out_pitch,
                                   j->out_pitch);
        } else {
            j->res[i] = rco_decode(j->stream + j->offsets[i],
                                   j->offsets[i + 1] - j->offsets[i], cnt, j->sym_bytes,
                                   j->K, c, cum, total,
                                   j->syms_out + first * j->sym_bytes);
        }
    }
    return NULL;
}

static int run_job(chunk_job *j, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > j->n_chunks) n_threads = (int)(j->n_chunks ? j->n_chunks : 1);
    atomic_init(&j->next, 0);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    int started = 0;
    for (int t = 1; t < n_threads; t++) {
        if (pthread_create(&th[started], NULL, chunk_worker, j) == 0) started++;
    }
    chunk_worker(j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    int bad = 0;
    for (uint64_t i = 0; i < j->n_chunks; i++)
        if (j->res[i] < 0) bad++;
    return bad;
}

int rco_encode_chunks(const void *syms, uint64_t n, int sym_bytes,
                      uint64_t chunk_syms, uint32_t K, const uint32_t *c,
                      const uint32_t *cum, const uint32_t *total,
                      int per_chunk_model, uint8_t *out, uint64_t out_pitch,
                      int64_t *lens, int n_threads) {
    chunk_job j;
    memset(&j, 0, sizeof j);
    j.decode = 0;
    j.syms_in = (const uint8_t *)syms;
    j.n = n;
    j.sym_bytes = sym_bytes;
    j.chunk_syms = chunk_syms;
    j.n_chunks = chunk_syms ? (n + chunk_syms - 1) / chunk_syms : 0;
    j.K = K;
    j.c = c;
    j.cum = cum;
    j.total = total;
    j.per_chunk_model = per_chunk_model;
    j.out = out;
    j.out_pitch = out_pitch;
    j.res = lens;
    return run_job(&j, n_threads);
}

int rco_decode_chunks(const uint8_t *stream, const uint64_t *offsets,
                      uint64_t n, int sym_bytes, uint64_t chunk_syms,
                      uint32_t K, const uint32_t *c, const uint32_t *cum,
                      const uint32_t *total, int per_chunk_model,
                      void *out_syms, int64_t *consumed, int n_threads) {
    chunk_job j;
    memset(&j, 0, sizeof j);
    j.decode = 1;
    j.syms_out = (uint8_t *)out_syms;
    j.n = n;
    j.sym_bytes = sym_bytes;
    j.chunk_syms = chunk_syms;
    j.n_chunks = chunk_syms ? (n + chunk_syms - 1) / chunk_syms : 0;
    j.K = K;
    j.c = c;
    j.cum = cum;
    j.total = total;
    j.per_chunk_model = per_chunk_model;
    j.stream = stream;
    j.offsets = offsets;
    j.res = consumed;
    return run_job(&j, n_threads);
}

/* ------------------------------------------------------------------------- */
/* synthetic data                                                            */
/* ------------------------------------------------------------------------- */

#define GOLDEN 0x9E3779B97F4A7C15ull

static inline uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct {
    void *out;
    uint64_t first, n;
    int sym_bytes;
    uint32_t K;
    uint64_t seed;
    const uint32_t *thr;
    uint32_t n_tables;
    uint64_t chunk_syms;
    uint64_t lo, hi;
} gen_job;

static void *gen_worker(void *arg) {
    gen_job *g = (gen_job *)arg;
    uint32_t nthr = g->K - 1;
    for (uint64_t i = g->lo; i < g->hi; i++) {
        uint64_t j = g->first + i;
        uint32_t r = (uint32_t)(mix64(g->seed + j * GOLDEN) >> 32);
        uint32_t t = g->n_tables > 1 ? (uint32_t)((j / g->chunk_syms) % g->n_tables) : 0;
        const uint32_t *th = g->thr + (uint64_t)t * nthr;
        /* symbol = number of thresholds <= r (thresholds are non-decreasing) */
        uint32_t lo = 0, hi = nthr;
        while (lo < hi) {
            uint32_t mid = (lo + hi) / 2;
            if (th[mid] <= r)
                lo = mid + 1;
            else
                hi = mid;
        }
        store_sym(g->out, i, g->sym_bytes, lo);
    }
    return NULL;
}

void rco_generate(void *out, uint64_t first, uint64_t n, int sym_bytes,
                  uint32_t K, uint64_t seed, const uint32_t *thr,
                  uint32_t n_tables, uint64_t chunk_syms, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    gen_job jobs[256];
    pthread_t th[256];
    uint64_t per = (n + n_threads - 1) / n_threads;
    int started = 0;
    for (int t = 0; t < n_threads; t++) {
        gen_job *g = &jobs[t];
        g->out = out;
        g->first = first;
        g->n = n;
        g->sym_bytes = sym_bytes;
        g->K = K;
        g->seed = seed;
        g->thr = thr;
        g->n_tables = n_tables ? n_tables : 1;
        g->chunk_syms = chunk_syms ? chunk_syms : 1;
        g->lo = (uint64_t)t * per < n ? (uint64_t)t * per : n;
        g->hi = g->lo + per < n ? g->lo + per : n;
    }
    for (int t = 1; t < n_threads; t++) {
        if (pthread_create(&th[started], NULL, gen_worker, &jobs[t]) == 0)
            started++;
        else
            gen_worker(&jobs[t]);
    }
    gen_worker(&jobs[0]);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

int rco_hardware_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int rco_adaptive_encode_chunks(const void *syms, uint64_t n, int sym_bytes, uint64_t chunk_syms,
                               uint32_t K, uint32_t inc, uint32_t limit, uint8_t *out,
                               uint64_t out_pitch, int64_t *lens, int n_threads) {
    chunk_job j;
    memset(&j, 0, sizeof j);
    j.adaptive = 1;
    j.inc = inc;
    j.limit = limit;
    j.syms_in = (const uint8_t *)syms;
    j.n = n;
    j.sym_bytes = sym_bytes;
    j.chunk_syms = chunk_syms;
    j.n_chunks = chunk_syms ? (n + chunk_syms - 1) / chunk_syms : 0;
    j.K = K;
    j.out = out;
    j.out_pitch = out_pitch;
    j.res = lens;
    return run_job(&j, n_threads);
}

int rco_adaptive_decode_chunks(const uint8_t *stream, const uint64_t *offsets, uint64_t n,
                               int sym_bytes, uint64_t chunk_syms, uint32_t K, uint32_t inc,
                               uint32_t limit, void *out_syms, int64_t *consumed, int n_threads) {
    chunk_job j;
    memset(&j, 0, sizeof j);
    j.adaptive = 1;
    j.decode = 1;
    j.inc = inc;
    j.limit = limit;
    j.syms_out = (uint8_t *)out_syms;
    j.n = n;
    j.sym_bytes = sym_bytes;
    j.chunk_syms = chunk_syms;
    j.n_chunks = chunk_syms ? (n + chunk_syms - 1) / chunk_syms : 0;
    j.K = K;
    j.stream = stream;
    j.offsets = offsets;
    j.res = consumed;
    return run_job(&j, n_threads);
}
