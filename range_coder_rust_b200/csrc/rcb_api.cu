// rcb_api.cu -- C ABI (include/rcb200.h) over the sm_100a kernels.
// There is no CPU implementation behind these entry points: without a CUDA
// device every call fails with RCB_ERR_NO_DEVICE / RCB_ERR_CUDA.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <new>

#include "../../include/rcb200.h"
#include "rcb_kernels.cuh"
#include "rcb_decode_row.cuh"
#include "rcb_stream.cuh"
#include "rcb_adaptive.cuh"

using namespace rcb;

static_assert(sizeof(ModelHdr) == 48, "ModelHdr layout");
static_assert(sizeof(LutEntry) == 16, "LutEntry layout");
static_assert((int)RCB_ST_TRUNCATED == (int)ST_TRUNCATED, "status codes");
static_assert((int)RCB_ST_RESTART == (int)ST_RESTART, "status codes");
static_assert(sizeof(rcb_restart_point) == sizeof(Restart) && sizeof(Restart) == 24, "restart point layout");
static_assert((int)RCB_MODEL_REGULAR == (int)MODEL_REGULAR, "model flags");

#define LUT_CAP 4096u       // buckets of the shared-model decode LUT (64 KiB of shared memory)
#define MAX_K_SHARED 4096u  // table + LUT must fit 227 KiB of shared memory
#define MAX_K 65536u
#define MAX_SLICES 16       // slices of one host batch in flight (rcb_encode_host / rcb_decode_host)
#define DEC_TP_THREADS 640  // lanes per block of the throughput flavour of the fused decoder (rcb_decode.cuh)

struct rcb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int enc_threads = 0, dec_threads = 0;  // 0 = choose from the lane count
    int sm_count = 148;
    uint8_t* staging = nullptr;
    size_t staging_bytes = 0;
    uint32_t* lens = nullptr;
    uint32_t* status = nullptr;
    size_t chunk_cap = 0;
    unsigned long long* d_summary = nullptr;  // [0..3] encode, [4..7] decode
    unsigned long long* h_summary = nullptr;  // pinned mirror
    uint32_t* d_words = nullptr;              // [0] min c, [1] model bad bits, [2] hist oob, [3] max total
    uint32_t* h_words = nullptr;              // pinned mirror
    void* h2d = nullptr;                      // device scratch of the host-buffer entry points
    size_t h2d_bytes = 0;
    cudaError_t last_err = cudaSuccess;
    uint64_t launches = 0;
    uint64_t pending_pitch = 0, pending_out_cap = 0, pending_n_chunks = 0;
    const uint64_t* pending_offsets = nullptr;
    // host-buffer pipeline: slices of one batch on their own streams (copies overlap the coder kernels)
    cudaStream_t slice_stream[MAX_SLICES] = {};
    // all host->device copies of a batch go through one stream and all device->host copies through
    // another, in slice order: copies issued on several streams share the bus and every slice would
    // arrive late; in order, slice k is complete after (k+1)/S of the transfer time
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    DecResume* d_resume = nullptr;  // lane states between the segment launches of rcb_decode_host
    size_t resume_cap = 0;
    cudaEvent_t in_ready[MAX_SLICES] = {};
    cudaEvent_t out_ready[MAX_SLICES] = {};
    unsigned long long* d_slice_summary = nullptr;  // [MAX_SLICES][8]
    unsigned long long* h_slice_summary = nullptr;  // pinned mirror
    bool timing = false;
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_enc = false, ev_dec = false;
};

#define EV(c, i)                                                   \
    do {                                                           \
        if ((c)->timing) cudaEventRecord((c)->ev[i], (c)->stream); \
    } while (0)

struct rcb_model {
    rcb_ctx* ctx = nullptr;
    uint32_t K = 0;
    uint64_t n_models = 0;
    uint2* d_tab = nullptr;
    uint32_t* d_total = nullptr;
    ModelHdr* d_hdr = nullptr;
    LutEntry* d_lut = nullptr;
    uint2* d_tab_cs = nullptr;  // shared model: floor(c * 2^64 / total) per symbol
    uint4* d_lut_cs = nullptr;  // ... and per LUT entry (both candidates)
    ModelHdr h_hdr0;       // header of model 0 (the shared model)
    uint32_t min_c = 0;    // smallest non-zero c over all models
    uint32_t max_total = 0;
    uint32_t bad_bits = 0; // 4: some model inconsistent, 8: some model irregular
    bool ready = false;
};

#define CK(ctx, call)                                  \
    do {                                               \
        cudaError_t e__ = (call);                      \
        if (e__ != cudaSuccess) {                      \
            (ctx)->last_err = e__;                     \
            return RCB_ERR_CUDA;                       \
        }                                              \
    } while (0)

// Every entry point runs on its ctx's device and leaves the calling thread's current device as it found
// it (a multi-GPU host -- torch, a Rust thread pool -- must not see its device change under it).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) {
            err = cudaSetDevice(dev);
            switched = err == cudaSuccess;
        }
    }
    ~DeviceGuard() {
        if (switched && prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define ON_DEVICE(ctx)                                 \
    DeviceGuard dg__((ctx)->device);                   \
    do {                                               \
        if (dg__.err != cudaSuccess) {                 \
            (ctx)->last_err = dg__.err;                \
            return RCB_ERR_CUDA;                       \
        }                                              \
    } while (0)

#define CK_LAUNCH(ctx)                                 \
    do {                                               \
        (ctx)->launches++;                             \
        cudaError_t e__ = cudaGetLastError();          \
        if (e__ != cudaSuccess) {                      \
            (ctx)->last_err = e__;                     \
            return RCB_ERR_CUDA;                       \
        }                                              \
    } while (0)

extern "C" const char* rcb_strerror(int err) {
    switch (err) {
        case RCB_OK: return "ok";
        case RCB_ERR_INVALID_ARGUMENT: return "invalid argument";
        case RCB_ERR_CUDA: return "CUDA error";
        case RCB_ERR_ZERO_TOTAL: return "total_freq is zero (reference: divide-by-zero panic)";
        case RCB_ERR_ZERO_FREQ_SYMBOL: return "coded symbol has c_freq == 0 (reference: never returns)";
        case RCB_ERR_LOWER_OVERFLOW: return "RangeCoderError::LowerBoundOverflow";
        case RCB_ERR_UPPER_OVERFLOW: return "RangeCoderError::UpperBoundOverflow";
        case RCB_ERR_SYMBOL_OUT_OF_RANGE: return "symbol index >= alphabet size";
        case RCB_ERR_OUT_CAPACITY: return "output buffer too small";
        case RCB_ERR_TRUNCATED_STREAM: return "code stream truncated (reference: pop_front panic)";
        case RCB_ERR_INVALID_MODEL: return "invalid model table (cum_freq > total_freq)";
        case RCB_ERR_UNSUPPORTED: return "unsupported configuration";
        case RCB_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        case RCB_ERR_NCCL: return "NCCL error (or libnccl.so.2 not found)";
        case RCB_ERR_RESTART_POINT: return "restart point inconsistent with the code stream";
        default: return "unknown error";
    }
}

extern "C" int rcb_version(void) { return RCB_VERSION; }

extern "C" int rcb_last_cuda_error(const rcb_ctx* ctx, const char** msg) {
    if (!ctx) return 0;
    if (msg) *msg = cudaGetErrorString(ctx->last_err);
    return (int)ctx->last_err;
}

static int status_to_error(uint32_t st) {
    switch (st) {
        case ST_OK: return RCB_OK;
        case ST_ZERO_FREQ: return RCB_ERR_ZERO_FREQ_SYMBOL;
        case ST_LOWER_OVERFLOW: return RCB_ERR_LOWER_OVERFLOW;
        case ST_UPPER_OVERFLOW: return RCB_ERR_UPPER_OVERFLOW;
        case ST_SYMBOL_RANGE: return RCB_ERR_SYMBOL_OUT_OF_RANGE;
        case ST_OUT_CAPACITY: return RCB_ERR_OUT_CAPACITY;
        case ST_TRUNCATED: return RCB_ERR_TRUNCATED_STREAM;
        case ST_RESTART: return RCB_ERR_RESTART_POINT;
        default: return RCB_ERR_INVALID_ARGUMENT;
    }
}

// ------------------------------------------------------------------ context
extern "C" int rcb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int rcb_ctx_create(int device, void* stream, rcb_ctx** out) {
    if (!out) return RCB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return RCB_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return RCB_ERR_INVALID_ARGUMENT;
    rcb_ctx* c = new (std::nothrow) rcb_ctx();
    if (!c) return RCB_ERR_INVALID_ARGUMENT;
    c->device = device;
    DeviceGuard dg(device);
    if (dg.err != cudaSuccess) {
        delete c;
        return RCB_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    c->stream = (cudaStream_t)stream;  // NULL = the default stream, as in the CUDA runtime
    bool ok = cudaMalloc(&c->d_summary, 8 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMallocHost(&c->h_summary, 8 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc(&c->d_words, 8 * sizeof(uint32_t)) == cudaSuccess &&
              cudaMallocHost(&c->h_words, 8 * sizeof(uint32_t)) == cudaSuccess;
    if (!ok) {
        rcb_ctx_destroy(c);
        return RCB_ERR_CUDA;
    }
    *out = c;
    return RCB_OK;
}

extern "C" int rcb_ctx_destroy(rcb_ctx* c) {
    if (!c) return RCB_OK;
    DeviceGuard dg(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->staging);
    cudaFree(c->lens);
    cudaFree(c->status);
    cudaFree(c->d_summary);
    cudaFreeHost(c->h_summary);
    cudaFree(c->d_words);
    cudaFreeHost(c->h_words);
    cudaFree(c->h2d);
    for (int i = 0; i < 7; i++)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < MAX_SLICES; i++) {
        if (c->slice_stream[i]) cudaStreamDestroy(c->slice_stream[i]);
        if (c->in_ready[i]) cudaEventDestroy(c->in_ready[i]);
        if (c->out_ready[i]) cudaEventDestroy(c->out_ready[i]);
    }
    cudaFree(c->d_resume);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    cudaFree(c->d_slice_summary);
    cudaFreeHost(c->h_slice_summary);
    delete c;
    return RCB_OK;
}

extern "C" int rcb_ctx_set_stream(rcb_ctx* c, void* stream) {
    if (!c) return RCB_ERR_INVALID_ARGUMENT;
    CK(c, cudaStreamSynchronize(c->stream));
    c->stream = (cudaStream_t)stream;
    return RCB_OK;
}

extern "C" int rcb_ctx_synchronize(rcb_ctx* c) {
    if (!c) return RCB_ERR_INVALID_ARGUMENT;
    CK(c, cudaStreamSynchronize(c->stream));
    return RCB_OK;
}

extern "C" int rcb_ctx_set_block_threads(rcb_ctx* c, int enc, int dec) {
    if (!c) return RCB_ERR_INVALID_ARGUMENT;
    auto okv = [](int v) { return v == 0 || (v >= 32 && v <= 512 && v % 32 == 0); };
    if (!okv(enc) || !okv(dec)) return RCB_ERR_INVALID_ARGUMENT;
    c->enc_threads = enc;
    c->dec_threads = dec;
    return RCB_OK;
}

extern "C" uint64_t rcb_ctx_launch_count(const rcb_ctx* c) { return c ? c->launches : 0; }

extern "C" int rcb_ctx_enable_timing(rcb_ctx* c, int on) {
    if (!c) return RCB_ERR_INVALID_ARGUMENT;
    ON_DEVICE(c);
    if (on && !c->ev[0])
        for (int i = 0; i < 7; i++) CK(c, cudaEventCreate(&c->ev[i]));
    c->timing = on != 0;
    c->ev_enc = c->ev_dec = false;
    return RCB_OK;
}

extern "C" int rcb_ctx_get_timings(rcb_ctx* c, float* ms, int n) {
    if (!c || !ms || n < 0) return RCB_ERR_INVALID_ARGUMENT;
    CK(c, cudaStreamSynchronize(c->stream));
    float v[5] = {0, 0, 0, 0, 0};
    if (c->ev_enc) {
        CK(c, cudaEventElapsedTime(&v[0], c->ev[0], c->ev[1]));
        CK(c, cudaEventElapsedTime(&v[1], c->ev[1], c->ev[2]));
        CK(c, cudaEventElapsedTime(&v[2], c->ev[2], c->ev[3]));
    }
    if (c->ev_dec) {
        CK(c, cudaEventElapsedTime(&v[3], c->ev[4], c->ev[5]));
        CK(c, cudaEventElapsedTime(&v[4], c->ev[5], c->ev[6]));
    }
    for (int i = 0; i < n && i < 5; i++) ms[i] = v[i];
    return RCB_OK;
}

// Threads per block of the coder kernels.  One lane per chunk: with few lanes (<= 128 per SM) small
// blocks spread the warps over all SMs (latency-bound regime, one warp per scheduler); with many
// lanes bigger blocks share one copy of the shared-memory tables (throughput regime).
// Lanes per block: the coder kernels keep one block per SM busy (the decoder's tables fill shared memory),
// so the lanes are spread evenly over as few full waves of SM-count blocks as 512-thread blocks allow --
// 16384 lanes: 128 threads, one wave; 65536 lanes: 448 threads, one wave (not 256 blocks of 256 = 1.7 waves).
static int pick_threads(const rcb_ctx* c, int user, uint64_t n_chunks, uint64_t max_threads = 512) {
    if (user) return user;
    const uint64_t sms = (uint64_t)(c->sm_count > 0 ? c->sm_count : 148);
    const uint64_t waves = (n_chunks + sms * max_threads - 1) / (sms * max_threads);
    const uint64_t per_block = (n_chunks + waves * sms - 1) / (waves ? waves * sms : 1);
    uint64_t t = (per_block + 31) / 32 * 32;
    if (t < 128) t = 128;
    if (t > max_threads) t = max_threads;
    return (int)t;
}

static int ensure_chunks(rcb_ctx* c, uint64_t n_chunks) {
    if (n_chunks <= c->chunk_cap) return RCB_OK;
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->lens);
    cudaFree(c->status);
    c->lens = nullptr;
    c->status = nullptr;
    c->chunk_cap = 0;
    CK(c, cudaMalloc(&c->lens, n_chunks * sizeof(uint32_t)));
    CK(c, cudaMalloc(&c->status, n_chunks * sizeof(uint32_t)));
    c->chunk_cap = n_chunks;
    return RCB_OK;
}

static int ensure_staging(rcb_ctx* c, size_t bytes) {
    if (bytes <= c->staging_bytes) return RCB_OK;
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->staging);
    c->staging = nullptr;
    c->staging_bytes = 0;
    CK(c, cudaMalloc(&c->staging, bytes));
    c->staging_bytes = bytes;
    return RCB_OK;
}

static int ensure_h2d(rcb_ctx* c, size_t bytes) {
    if (bytes <= c->h2d_bytes) return RCB_OK;
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->h2d);
    c->h2d = nullptr;
    c->h2d_bytes = 0;
    CK(c, cudaMalloc(&c->h2d, bytes));
    c->h2d_bytes = bytes;
    return RCB_OK;
}

// ------------------------------------------------------------ device memory
extern "C" int rcb_device_alloc(rcb_ctx* c, uint64_t bytes, void** d_out) {
    if (!c || !d_out) return RCB_ERR_INVALID_ARGUMENT;
    *d_out = nullptr;
    ON_DEVICE(c);
    CK(c, cudaMalloc(d_out, bytes ? (size_t)bytes : 16));
    return RCB_OK;
}

extern "C" int rcb_device_free(rcb_ctx* c, void* d_ptr) {
    if (!c) return RCB_ERR_INVALID_ARGUMENT;
    if (!d_ptr) return RCB_OK;
    ON_DEVICE(c);
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaFree(d_ptr));
    return RCB_OK;
}

extern "C" int rcb_copy_to_device(rcb_ctx* c, void* d_dst, const void* h_src, uint64_t bytes) {
    if (!c || (bytes && (!d_dst || !h_src))) return RCB_ERR_INVALID_ARGUMENT;
    ON_DEVICE(c);
    if (bytes) CK(c, cudaMemcpyAsync(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RCB_OK;
}

extern "C" int rcb_copy_to_host(rcb_ctx* c, void* h_dst, const void* d_src, uint64_t bytes) {
    if (!c || (bytes && (!h_dst || !d_src))) return RCB_ERR_INVALID_ARGUMENT;
    ON_DEVICE(c);
    if (bytes) CK(c, cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RCB_OK;
}

// -------------------------------------------------------------------- model
extern "C" int rcb_model_create(rcb_ctx* c, uint32_t K, uint64_t n_models, rcb_model** out) {
    if (!c || !out || K == 0 || K > MAX_K || n_models == 0) return RCB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (n_models == 1 && K > MAX_K_SHARED) return RCB_ERR_UNSUPPORTED;
    ON_DEVICE(c);
    rcb_model* m = new (std::nothrow) rcb_model();
    if (!m) return RCB_ERR_INVALID_ARGUMENT;
    m->ctx = c;
    m->K = K;
    m->n_models = n_models;
    bool ok = cudaMalloc(&m->d_tab, n_models * K * sizeof(uint2)) == cudaSuccess &&
              cudaMalloc(&m->d_total, n_models * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc(&m->d_hdr, n_models * sizeof(ModelHdr)) == cudaSuccess;
    if (ok && n_models == 1)
        ok = cudaMalloc(&m->d_lut, LUT_CAP * sizeof(LutEntry)) == cudaSuccess &&
             cudaMalloc(&m->d_tab_cs, (size_t)K * sizeof(uint2)) == cudaSuccess &&
             cudaMalloc(&m->d_lut_cs, LUT_CAP * sizeof(uint4)) == cudaSuccess;
    if (!ok) {
        c->last_err = cudaGetLastError();
        rcb_model_destroy(m);
        return RCB_ERR_CUDA;
    }
    *out = m;
    return RCB_OK;
}

extern "C" int rcb_model_destroy(rcb_model* m) {
    if (!m) return RCB_OK;
    DeviceGuard dg(m->ctx ? m->ctx->device : 0);
    if (m->ctx) cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->d_tab);
    cudaFree(m->d_total);
    cudaFree(m->d_hdr);
    cudaFree(m->d_lut);
    cudaFree(m->d_tab_cs);
    cudaFree(m->d_lut_cs);
    delete m;
    return RCB_OK;
}

template <typename SYM>
static int launch_hist(rcb_ctx* c, const void* d_syms, uint64_t n, uint32_t K, uint64_t chunk_syms,
                       void* d_counts) {
    const int threads = K <= 4096 ? 256 : 64;  // private copy of the K bins per warp
    const size_t smem = (size_t)(threads / 32) * K * sizeof(uint32_t);
    static int shared_env = -1;  // RCB_HIST_SHARED=0: the per-warp-copy kernels (kept for comparison)
    if (shared_env < 0) {
        const char* e = getenv("RCB_HIST_SHARED");
        shared_env = e ? atoi(e) : 1;
    }
    if (chunk_syms == 0 && shared_env && K <= 16384) {
        // one copy of the bins per block, REP words per bin: 32 KiB (bytes) .. 128 KiB per block
        CK(c, cudaMemsetAsync(d_counts, 0, (size_t)K * sizeof(unsigned long long), c->stream));
        uint32_t rep_log2 = 5;
        while (rep_log2 > 0 && ((size_t)K << rep_log2) > 32768) rep_log2--;
        const bool full = sizeof(SYM) == 1 ? K == 256 : K == 65536;
        const size_t smem_s = ((size_t)(K + (full ? 0 : 1)) << rep_log2) * sizeof(uint32_t);
        const int thr = 512;
        size_t per_sm = (size_t)200 * 1024 / (smem_s + 1024);
        if (per_sm > 4) per_sm = 4;  // 2048 threads per SM
        if (per_sm < 1) per_sm = 1;
        const uint64_t nvec = n / (16 / sizeof(SYM));
        const uint64_t want = (nvec + thr * 4 - 1) / (thr * 4);
        const uint64_t maxb = (uint64_t)c->sm_count * per_sm;
        const int blocks = (int)(want < 1 ? 1 : (want > maxb ? maxb : want));
        auto go = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
            kern<<<blocks, thr, smem_s, c->stream>>>((const SYM*)d_syms, n, K, rep_log2, (unsigned long long*)d_counts,
                                                     c->d_words + 2);
        };
        if (full) go(hist_global_shared_kernel<SYM, true, 4>);
        else go(hist_global_shared_kernel<SYM, false, 4>);
        CK_LAUNCH(c);
    } else if (chunk_syms == 0 && sizeof(SYM) == 1 && K <= 256) {
        // byte symbols: replicated bins (rcb_kernels.cuh), 64 KiB of bins per block, 3 blocks per SM
        CK(c, cudaMemsetAsync(d_counts, 0, (size_t)K * sizeof(unsigned long long), c->stream));
        static int rep_env = -1;
        if (rep_env < 0) {
            const char* e = getenv("RCB_HIST_REP");
            rep_env = e ? atoi(e) : 0;
        }
        const uint32_t rep = rep_env == 8 || rep_env == 16 || rep_env == 32 ? (uint32_t)rep_env : 16u;
        const int thr8 = rep == 8 ? 256 : (rep == 16 ? 128 : 64);  // 8 / 4 / 2 warps share the 64 KiB
        const size_t smem8 = (size_t)(thr8 / 32) * K * rep * sizeof(uint32_t);
        const uint64_t nvec = n / 16;
        const uint64_t want = (nvec + thr8 * 8 - 1) / (thr8 * 8);
        const uint64_t maxb = (uint64_t)c->sm_count * 3;
        const int blocks = (int)(want < 1 ? 1 : (want > maxb ? maxb : want));
        auto go = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8);
            kern<<<blocks, thr8, smem8, c->stream>>>((const uint8_t*)d_syms, n, K, (unsigned long long*)d_counts,
                                                     c->d_words + 2);
        };
        if (K == 256) {
            if (rep == 8) go(hist_global_u8_kernel<true, 8, 2>);
            else if (rep == 16) go(hist_global_u8_kernel<true, 16, 4>);
            else go(hist_global_u8_kernel<true, 32, 8>);
        } else {
            if (rep == 8) go(hist_global_u8_kernel<false, 8, 2>);
            else if (rep == 16) go(hist_global_u8_kernel<false, 16, 4>);
            else go(hist_global_u8_kernel<false, 32, 8>);
        }
        CK_LAUNCH(c);
    } else if (chunk_syms == 0) {
        CK(c, cudaMemsetAsync(d_counts, 0, (size_t)K * sizeof(unsigned long long), c->stream));
        auto kern = hist_global_kernel<SYM>;
        CK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uint64_t nvec = n / (16 / sizeof(SYM));
        uint64_t want = (nvec + threads * 8 - 1) / (threads * 8);
        uint64_t maxb = (uint64_t)c->sm_count * 8;
        int blocks = (int)(want < 1 ? 1 : (want > maxb ? maxb : want));
        kern<<<blocks, threads, smem, c->stream>>>((const SYM*)d_syms, n, K,
                                                    (unsigned long long*)d_counts, c->d_words + 2);
        CK_LAUNCH(c);
    } else {
        uint64_t n_chunks = (n + chunk_syms - 1) / chunk_syms;
        if (n_chunks == 0) return RCB_OK;
        if (n_chunks > 0x7FFFFFFFull) return RCB_ERR_UNSUPPORTED;
        // one copy of the bins per block, REP interleaved words per bin (up to 32: a bank per lane).  Every chunk
        // zeroes and merges its copy, so the copy grows with the chunk: at most a word per 4 symbols, 8 to 32 KiB
        // (K = 256: REP = 32 for 64 KiB chunks 0.29 ms per GiB, REP = 16 for 16 KiB chunks 0.39 ms; REP = 8 was
        // 0.36 / 0.43, REP = 32 at 16 KiB 0.43; K = 4096: REP = 2, 0.35 ms)
        uint64_t cap_words = chunk_syms / 4;
        cap_words = cap_words < 2048 ? 2048 : (cap_words > 8192 ? 8192 : cap_words);
        uint32_t rep_log2 = 5;
        while (rep_log2 > 0 && ((uint64_t)K << rep_log2) > cap_words) rep_log2--;
        const bool full = sizeof(SYM) == 1 ? K == 256 : K == 65536;
        const size_t smem_c = ((size_t)(K + (full ? 0 : 1)) << rep_log2) * sizeof(uint32_t);
        auto go = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
            kern<<<(unsigned)n_chunks, 256, smem_c, c->stream>>>((const SYM*)d_syms, n, chunk_syms, K, rep_log2,
                                                                  (uint32_t*)d_counts, c->d_words + 2);
        };
        if (full) go(hist_chunks_kernel<SYM, true>);
        else go(hist_chunks_kernel<SYM, false>);
        CK_LAUNCH(c);
    }
    return RCB_OK;
}

extern "C" int rcb_histogram(rcb_ctx* c, const void* d_syms, uint64_t n_syms, int sym_bytes, uint32_t K,
                             uint64_t chunk_syms, void* d_counts) {
    if (!c || !d_counts || (n_syms && !d_syms) || K == 0 || K > MAX_K) return RCB_ERR_INVALID_ARGUMENT;
    if (sym_bytes != 1 && sym_bytes != 2) return RCB_ERR_UNSUPPORTED;
    if (K > 16384) return RCB_ERR_UNSUPPORTED;  // per-warp private bins must fit shared memory
    if (reinterpret_cast<uintptr_t>(d_syms) & 15u) return RCB_ERR_INVALID_ARGUMENT;
    ON_DEVICE(c);
    CK(c, cudaMemsetAsync(c->d_words + 2, 0, sizeof(uint32_t), c->stream));
    int r = sym_bytes == 1 ? launch_hist<uint8_t>(c, d_syms, n_syms, K, chunk_syms, d_counts)
                           : launch_hist<uint16_t>(c, d_syms, n_syms, K, chunk_syms, d_counts);
    if (r) return r;
    CK(c, cudaMemcpyAsync(c->h_words + 2, c->d_words + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return c->h_words[2] ? RCB_ERR_SYMBOL_OUT_OF_RANGE : RCB_OK;
}

static int finalize_model(rcb_ctx* c, rcb_model* m) {
    // words: [0] min c (atomicMin), [1] bad bits, [2] histogram flag (untouched), [3] max total
    c->h_words[4] = 0xFFFFFFFFu;
    c->h_words[5] = 0u;
    c->h_words[7] = 0u;
    CK(c, cudaMemcpyAsync(c->d_words, c->h_words + 4, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(c->d_words + 3, c->h_words + 7, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    if (m->n_models > 0x7FFFFFFFull) return RCB_ERR_UNSUPPORTED;
    finalize_models_kernel<<<(unsigned)m->n_models, 256, 0, c->stream>>>(
        m->d_tab, m->d_total, m->K, m->d_hdr, m->n_models == 1 ? m->d_lut : nullptr, LUT_CAP,
        m->n_models == 1 ? m->d_tab_cs : nullptr, m->n_models == 1 ? m->d_lut_cs : nullptr, c->d_words);
    CK_LAUNCH(c);
    CK(c, cudaMemcpyAsync(c->h_words, c->d_words, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(&m->h_hdr0, m->d_hdr, sizeof(ModelHdr), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    m->min_c = c->h_words[0];
    uint32_t bad = c->h_words[1];
    m->ready = false;
    if (bad & 2u) return RCB_ERR_ZERO_TOTAL;
    if (bad & 1u) return RCB_ERR_INVALID_MODEL;
    // per-model flags are in the headers; aggregate what the launch dispatch needs
    m->bad_bits = bad & (4u | 8u);
    m->max_total = c->h_words[3];
    m->ready = true;
    return RCB_OK;
}

extern "C" int rcb_model_from_counts(rcb_ctx* c, rcb_model* m, const void* d_counts, int count_bytes) {
    if (!c || !m || !d_counts || m->ctx != c) return RCB_ERR_INVALID_ARGUMENT;
    if (count_bytes != 4 && count_bytes != 8) return RCB_ERR_INVALID_ARGUMENT;
    if (m->n_models > 0x7FFFFFFFull) return RCB_ERR_UNSUPPORTED;
    ON_DEVICE(c);
    if (count_bytes == 8)
        counts_to_tables_kernel<unsigned long long><<<(unsigned)m->n_models, 256, 0, c->stream>>>(
            (const unsigned long long*)d_counts, m->K, m->d_tab, m->d_total);
    else
        counts_to_tables_kernel<uint32_t><<<(unsigned)m->n_models, 256, 0, c->stream>>>(
            (const uint32_t*)d_counts, m->K, m->d_tab, m->d_total);
    CK_LAUNCH(c);
    return finalize_model(c, m);
}

extern "C" int rcb_model_from_tables(rcb_ctx* c, rcb_model* m, const uint32_t* h_c, const uint32_t* h_cum,
                                     const uint32_t* h_total) {
    if (!c || !m || !h_c || !h_cum || !h_total || m->ctx != c) return RCB_ERR_INVALID_ARGUMENT;
    ON_DEVICE(c);
    const size_t n = (size_t)m->n_models * m->K;
    uint2* tmp = (uint2*)malloc(n * sizeof(uint2));
    if (!tmp) return RCB_ERR_INVALID_ARGUMENT;
    for (size_t i = 0; i < n; i++) tmp[i] = make_uint2(h_cum[i], h_c[i]);
    cudaError_t e = cudaMemcpyAsync(m->d_tab, tmp, n * sizeof(uint2), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(m->d_total, h_total, m->n_models * sizeof(uint32_t), cudaMemcpyHostToDevice,
                            c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    free(tmp);
    if (e != cudaSuccess) {
        c->last_err = e;
        return RCB_ERR_CUDA;
    }
    return finalize_model(c, m);
}

extern "C" int rcb_model_get_tables(rcb_ctx* c, const rcb_model* m, uint64_t index, uint32_t* h_c,
                                    uint32_t* h_cum, uint32_t* h_total, uint32_t* h_flags) {
    if (!c || !m || index >= m->n_models) return RCB_ERR_INVALID_ARGUMENT;
    ON_DEVICE(c);
    uint2* tmp = (uint2*)malloc((size_t)m->K * sizeof(uint2));
    if (!tmp) return RCB_ERR_INVALID_ARGUMENT;
    ModelHdr h;
    cudaError_t e = cudaMemcpyAsync(tmp, m->d_tab + index * m->K, (size_t)m->K * sizeof(uint2),
                                    cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(&h, m->d_hdr + index, sizeof(ModelHdr), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        free(tmp);
        c->last_err = e;
        return RCB_ERR_CUDA;
    }
    for (uint32_t i = 0; i < m->K; i++) {
        if (h_cum) h_cum[i] = tmp[i].x;
        if (h_c) h_c[i] = tmp[i].y;
    }
    free(tmp);
    if (h_total) *h_total = h.div.total;
    if (h_flags) *h_flags = h.flags;
    return RCB_OK;
}

// ------------------------------------------------------------------- encode
static uint64_t staging_pitch(const rcb_model* m, uint64_t chunk_syms) {
    // bits per symbol <= log2(total / min c) (+ ~2^-15 from the truncating division);
    // loop-2 range truncation adds < 0.1 %: 1/128 margin + 64 bytes, and the sync entry
    // point retries with the exact need if a row still overflows.
    double total = (double)m->max_total, minc = (double)(m->min_c ? m->min_c : 1);
    double bits = log2(total / minc);
    if (bits < 0.0) bits = 0.0;
    double bytes = (double)chunk_syms * bits / 8.0;
    uint64_t p = (uint64_t)(bytes * (1.0 + 1.0 / 128.0)) + 64 + 8;
    return (p + 15) & ~15ull;
}

extern "C" uint64_t rcb_encode_bound(rcb_ctx* c, const rcb_model* m, uint64_t n_syms, int sym_bytes,
                                     uint64_t chunk_syms) {
    (void)sym_bytes;
    if (!c || !m || !m->ready || chunk_syms == 0) return 0;
    uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    return n_chunks * staging_pitch(m, chunk_syms) + 16;
}

// Launch geometry and kernel flavour of one encode call.
// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [n_rows][row_bytes] bytes, box {64 bytes, 32 rows}, 64-byte swizzle, 256-byte L2 promotion
static bool make_symbol_tensor_map(CUtensorMap* tm, const void* base, uint64_t row_bytes, uint64_t n_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || row_bytes % 16 || row_bytes >= (1ull << 32) || n_rows == 0 || n_rows >= (1ull << 32)) return false;
    const cuuint64_t dims[2] = {row_bytes, n_rows};
    const cuuint64_t strides[1] = {row_bytes};
    const cuuint32_t box[2] = {TMA_BOX_BYTES, 32};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Restart points of a call (rcb_core.cuh: Restart): every `syms` symbols of a chunk, per_chunk records per chunk.
struct RestartSpec {
    Restart* pts = nullptr;
    uint64_t syms = 0;
    uint32_t per_chunk = 0;
    bool on() const { return pts != nullptr && per_chunk > 0; }
};

// Validates (restart_syms, d_restart) of an entry point.  Off (whole chunks) when either is 0 or the chunk has
// a single part; restart_syms must be a multiple of 64 (whole input vectors and output words per part).
static int make_restart_spec(uint64_t chunk_syms, uint64_t restart_syms, const void* d_restart, RestartSpec* out) {
    *out = RestartSpec();
    if (!restart_syms || !d_restart) return RCB_OK;
    if (restart_syms % 64) return RCB_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(d_restart) & 7u) return RCB_ERR_INVALID_ARGUMENT;
    const uint64_t parts = (chunk_syms + restart_syms - 1) / restart_syms;
    if (parts > 64) return RCB_ERR_UNSUPPORTED;  // lanes per chunk (a block of the row kernel holds whole chunks)
    if (parts <= 1) return RCB_OK;
    out->pts = static_cast<Restart*>(const_cast<void*>(d_restart));
    out->syms = restart_syms;
    out->per_chunk = (uint32_t)(parts - 1);
    return RCB_OK;
}

extern "C" uint64_t rcb_restart_points_per_chunk(uint64_t chunk_syms, uint64_t restart_syms) {
    if (!chunk_syms || !restart_syms) return 0;
    return (chunk_syms + restart_syms - 1) / restart_syms - 1;
}

struct EncPlan {
    int table;      // TAB_*
    int fmode;      // FM_*
    bool checked;
    bool tma;       // symbols staged by TMA (encode_tma_kernel)
    int threads;
    uint32_t lanes; // chunks per block
    size_t smem;
};

static EncPlan plan_encode(const rcb_ctx* c, const rcb_model* m, uint64_t n_chunks) {
    EncPlan p;
    p.tma = false;
    const bool shared = m->n_models == 1;
    p.checked = (m->bad_bits & 4u) != 0;
    const bool regular = (m->bad_bits & 8u) == 0;
    p.threads = pick_threads(c, c->enc_threads, n_chunks);
    if (shared) {
        p.table = TAB_SHARED;
        const bool pow2 = (m->h_hdr0.flags & MODEL_POW2) != 0;
        // general totals: the divide-free step needs cs = floor(c * 2^64 / total) < 2^64, i.e. no c == total
        const bool gencs = !pow2 && !(m->h_hdr0.flags & MODEL_FULLC) && !getenv("RCB_NO_GENCS");
        const bool genm2 = !pow2 && recip2_ok(m->h_hdr0.div.total) && !getenv("RCB_NO_M2");  // table-wide reciprocal
        p.fmode = p.checked ? FM_GENERIC
                            : (pow2 ? (m->h_hdr0.div.shift >= 24 ? FM_BIG : FM_POW2)
                                    : (genm2 ? FM_GENM2 : (gencs ? FM_GENCS : FM_GEN)));
        p.lanes = (uint32_t)p.threads;
        p.smem = (((size_t)m->K * (p.fmode == FM_GENCS ? sizeof(uint4) : sizeof(uint2)) + 15) & ~(size_t)15);
        if (p.fmode != FM_GENERIC) p.smem += (size_t)p.threads * ENC_RING_STRIDE;  // per-lane input rings
        return p;
    }
    // per-chunk models: each lane's cum[K+1] in its own shared-memory row (+ its input ring) when it fits
    const size_t row = ((size_t)m->K + 1) * sizeof(uint32_t);
    const size_t budget = 216 * 1024;
    const uint32_t lmax = (uint32_t)(budget / (row + ENC_RING_STRIDE));
    if (!p.checked && regular && lmax >= 32) {
        p.table = TAB_LANE;
        p.fmode = FM_LANE;
        uint32_t L = lmax < 512u ? lmax : 512u;
        const uint64_t sms = (uint64_t)(c->sm_count > 0 ? c->sm_count : 148);
        const uint64_t even = (n_chunks + sms - 1) / sms;  // one wave, all SMs busy, when it fits
        if (even <= L) L = (uint32_t)(even ? even : 1);
        if (c->enc_threads && (uint32_t)c->enc_threads < L) L = (uint32_t)c->enc_threads;
        p.lanes = L;
        p.threads = (int)((L + 31) / 32 * 32);
        p.smem = (((size_t)L * row + 15) & ~(size_t)15) + (size_t)p.threads * ENC_RING_STRIDE;
        return p;
    }
    p.table = TAB_GLOBAL;
    p.fmode = FM_GENERIC;
    p.lanes = (uint32_t)p.threads;
    p.smem = 0;
    return p;
}

template <typename SYM, int TABLE, int FMODE, bool CHECKED>
static void launch_encode_rc(rcb_ctx* c, const EncodeArgs& a, const EncPlan& p, unsigned blocks, bool rangechk) {
    auto go = [&](auto kern) {
        if (p.smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        kern<<<blocks, p.threads, p.smem, c->stream>>>(a);
    };
    if (rangechk) go(encode_kernel<SYM, TABLE, FMODE, CHECKED, true>);
    else go(encode_kernel<SYM, TABLE, FMODE, CHECKED, false>);
}

template <typename SYM, int FMODE>
static void launch_encode_tma(rcb_ctx* c, const EncodeArgs& a, const EncPlan& p, unsigned blocks, bool rangechk,
                              const CUtensorMap& tm) {
    auto go = [&](auto kern) {
        if (p.smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        kern<<<blocks, p.threads, p.smem, c->stream>>>(a, tm);
    };
    if (rangechk) go(encode_tma_kernel<SYM, FMODE, true>);
    else go(encode_tma_kernel<SYM, FMODE, false>);
}

template <typename SYM>
static void launch_encode_variant(rcb_ctx* c, const rcb_model* m, const EncodeArgs& a, const EncPlan& p,
                                  unsigned blocks) {
    const bool rangechk = (uint64_t)m->K < (1ull << (8 * sizeof(SYM)));
    if (p.tma) {
        CUtensorMap tm;
        if (make_symbol_tensor_map(&tm, a.syms, a.chunk_syms * sizeof(SYM), a.n_chunks)) {
            if (p.fmode == FM_BIG) return launch_encode_tma<SYM, FM_BIG>(c, a, p, blocks, rangechk, tm);
            if (p.fmode == FM_GENM2) return launch_encode_tma<SYM, FM_GENM2>(c, a, p, blocks, rangechk, tm);
        }
    }
    if (p.table == TAB_SHARED) {
        switch (p.fmode) {
            case FM_BIG: launch_encode_rc<SYM, TAB_SHARED, FM_BIG, false>(c, a, p, blocks, rangechk); break;
            case FM_POW2: launch_encode_rc<SYM, TAB_SHARED, FM_POW2, false>(c, a, p, blocks, rangechk); break;
            case FM_GEN: launch_encode_rc<SYM, TAB_SHARED, FM_GEN, false>(c, a, p, blocks, rangechk); break;
            case FM_GENCS: launch_encode_rc<SYM, TAB_SHARED, FM_GENCS, false>(c, a, p, blocks, rangechk); break;
            case FM_GENM2: launch_encode_rc<SYM, TAB_SHARED, FM_GENM2, false>(c, a, p, blocks, rangechk); break;
            default: launch_encode_rc<SYM, TAB_SHARED, FM_GENERIC, true>(c, a, p, blocks, rangechk); break;
        }
    } else if (p.table == TAB_LANE) {
        launch_encode_rc<SYM, TAB_LANE, FM_LANE, false>(c, a, p, blocks, rangechk);
    } else {
        if (p.checked) launch_encode_rc<SYM, TAB_GLOBAL, FM_GENERIC, true>(c, a, p, blocks, rangechk);
        else launch_encode_rc<SYM, TAB_GLOBAL, FM_GENERIC, false>(c, a, p, blocks, rangechk);
    }
}

// Issue encode + length scan + gather for a run of chunks on the ctx's current stream.  All buffers are
// explicit so the host-buffer pipeline can run slices of one batch on several streams.
// model_first: index of the first chunk's model (per-chunk models).
static int encode_issue(rcb_ctx* c, const void* d_syms, uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                        const rcb_model* m, uint64_t model_first, uint8_t* staging, uint64_t pitch, uint32_t* lens,
                        uint32_t* status, uint8_t* d_out, uint64_t out_cap, uint64_t* d_offsets,
                        unsigned long long* d_summary, bool timed, const RestartSpec* rs = nullptr) {
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    const bool shared = m->n_models == 1;
    EncodeArgs a;
    a.restart = rs && rs->on() ? rs->pts : nullptr;
    a.restart_syms = rs && rs->on() ? rs->syms : 0;
    a.restart_per_chunk = rs && rs->on() ? rs->per_chunk : 0u;
    a.syms = d_syms;
    a.n_syms = n_syms;
    a.chunk_syms = chunk_syms;
    a.n_chunks = n_chunks;
    a.tabs = shared ? m->d_tab : m->d_tab + model_first * m->K;
    a.hdrs = shared ? m->d_hdr : m->d_hdr + model_first;
    a.tab_cs = m->d_tab_cs;
    a.K = m->K;
    a.staging = staging;
    a.pitch = pitch;
    a.lens = lens;
    a.status = status;
    EncPlan plan = plan_encode(c, m, n_chunks);
    // TMA staging of the symbols (RCB_ENC_TMA=1): shared table, every chunk whole, rows a multiple of 16 bytes
    if (getenv("RCB_ENC_TMA") && plan.table == TAB_SHARED && (plan.fmode == FM_BIG || plan.fmode == FM_GENM2) &&
        n_syms % chunk_syms == 0 && (chunk_syms * sym_bytes) % 16 == 0 && chunk_syms * sym_bytes >= 64) {
        plan.tma = true;
        const size_t tab = (((size_t)m->K * sizeof(uint2) + 15) & ~(size_t)15);
        const size_t warps = (size_t)plan.threads / 32;
        plan.smem = tab + 1024 + warps * TMA_STAGES * TMA_STAGE_BYTES + warps * TMA_STAGES * 8;
    }
    a.lanes_per_block = plan.lanes;
    const unsigned blocks = (unsigned)((n_chunks + plan.lanes - 1) / plan.lanes);
    if (sym_bytes == 1)
        launch_encode_variant<uint8_t>(c, m, a, plan, blocks);
    else
        launch_encode_variant<uint16_t>(c, m, a, plan, blocks);
    CK_LAUNCH(c);
    if (timed) EV(c, 1);
    scan_lengths_kernel<<<1, 1024, 0, c->stream>>>(lens, status, n_chunks, d_offsets, d_summary);
    CK_LAUNCH(c);
    if (timed) EV(c, 2);
    gather_kernel<<<(unsigned)n_chunks, 256, 0, c->stream>>>(staging, pitch, lens, d_offsets, d_out, out_cap);
    CK_LAUNCH(c);
    if (timed) EV(c, 3);
    return RCB_OK;
}

static int encode_launch(rcb_ctx* c, const void* d_syms, uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                         const rcb_model* m, uint8_t* d_out, uint64_t out_cap, uint64_t* d_offsets,
                         uint32_t* d_status, uint64_t pitch_override, uint64_t restart_syms = 0,
                         void* d_restart = nullptr) {
    if (!c || !m || !m->ready || m->ctx != c || !d_offsets || chunk_syms == 0) return RCB_ERR_INVALID_ARGUMENT;
    RestartSpec rs;
    if (int rr = make_restart_spec(chunk_syms, restart_syms, d_restart, &rs)) return rr;
    if (n_syms && !d_syms) return RCB_ERR_INVALID_ARGUMENT;
    if (sym_bytes != 1 && sym_bytes != 2) return RCB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(d_syms) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u) ||
        (reinterpret_cast<uintptr_t>(d_offsets) & 7u))
        return RCB_ERR_INVALID_ARGUMENT;
    if (chunk_syms > 0x40000000ull) return RCB_ERR_UNSUPPORTED;  // code length is tracked in 32 bits
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    if (m->n_models != 1 && m->n_models != n_chunks) return RCB_ERR_INVALID_ARGUMENT;
    if (n_chunks > 0x7FFFFFFFull) return RCB_ERR_UNSUPPORTED;
    ON_DEVICE(c);
    if (n_chunks == 0) {
        CK(c, cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t), c->stream));
        CK(c, cudaMemsetAsync(c->d_summary, 0, 4 * sizeof(unsigned long long), c->stream));
        c->pending_out_cap = out_cap;
        c->pending_offsets = d_offsets;
        c->pending_n_chunks = 0;
        return RCB_OK;
    }
    int r = ensure_chunks(c, n_chunks);
    if (r) return r;
    uint64_t pitch = pitch_override ? pitch_override : staging_pitch(m, chunk_syms);
    if (pitch > 0xFFFFFFF0ull) return RCB_ERR_UNSUPPORTED;
    r = ensure_staging(c, (size_t)(n_chunks * pitch + 64));
    if (r) return r;
    c->pending_pitch = pitch;
    c->pending_out_cap = out_cap;
    c->pending_offsets = d_offsets;
    c->pending_n_chunks = n_chunks;

    EV(c, 0);
    r = encode_issue(c, d_syms, n_syms, sym_bytes, chunk_syms, m, 0, c->staging, pitch, c->lens,
                     d_status ? d_status : c->status, d_out, out_cap, d_offsets, c->d_summary, true, &rs);
    if (r) return r;
    c->ev_enc = c->timing;
    return RCB_OK;
}

extern "C" int rcb_encode_chunks_async(rcb_ctx* c, const void* d_syms, uint64_t n_syms, int sym_bytes,
                                       uint64_t chunk_syms, const rcb_model* m, uint8_t* d_out,
                                       uint64_t out_cap, uint64_t* d_offsets, uint32_t* d_status) {
    return encode_launch(c, d_syms, n_syms, sym_bytes, chunk_syms, m, d_out, out_cap, d_offsets, d_status, 0);
}

// Fetch the summary of the last encode on this ctx.  Sets *need_pitch when a staging row overflowed.
static int encode_result(rcb_ctx* c, uint64_t* total_bytes, uint64_t* need_pitch) {
    const uint64_t* d_offsets = c->pending_offsets;
    const uint64_t n_chunks = c->pending_n_chunks;
    unsigned long long total = 0;
    CK(c, cudaMemcpyAsync(c->h_summary, c->d_summary, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                          c->stream));
    if (d_offsets)
        CK(c, cudaMemcpyAsync(c->h_summary + 4, d_offsets + n_chunks, sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (d_offsets) total = c->h_summary[4];
    if (total_bytes) *total_bytes = total;
    if (need_pitch) *need_pitch = 0;
    if (c->h_summary[0]) {
        uint32_t st = (uint32_t)c->h_summary[2];
        if (st == ST_OUT_CAPACITY && need_pitch) *need_pitch = ((c->h_summary[3] + 15) & ~15ull) + 16;
        return status_to_error(st);
    }
    if (d_offsets && total > c->pending_out_cap) return RCB_ERR_OUT_CAPACITY;
    return RCB_OK;
}

extern "C" int rcb_encode_chunks(rcb_ctx* c, const void* d_syms, uint64_t n_syms, int sym_bytes,
                                 uint64_t chunk_syms, const rcb_model* m, uint8_t* d_out, uint64_t out_cap,
                                 uint64_t* d_offsets, uint32_t* d_status, uint64_t* h_out_bytes) {
    int r = encode_launch(c, d_syms, n_syms, sym_bytes, chunk_syms, m, d_out, out_cap, d_offsets, d_status, 0);
    if (r) return r;
    uint64_t need = 0;
    r = encode_result(c, h_out_bytes, &need);
    if (r == RCB_ERR_OUT_CAPACITY && need) {
        // a staging row was too small for this data: rerun once with the exact need
        r = encode_launch(c, d_syms, n_syms, sym_bytes, chunk_syms, m, d_out, out_cap, d_offsets, d_status, need);
        if (r) return r;
        r = encode_result(c, h_out_bytes, nullptr);
    }
    return r;
}

extern "C" int rcb_encode_chunks_restart_async(rcb_ctx* c, const void* d_syms, uint64_t n_syms, int sym_bytes,
                                               uint64_t chunk_syms, const rcb_model* m, uint8_t* d_out,
                                               uint64_t out_cap, uint64_t* d_offsets, uint32_t* d_status,
                                               uint64_t restart_syms, rcb_restart_point* d_restart) {
    return encode_launch(c, d_syms, n_syms, sym_bytes, chunk_syms, m, d_out, out_cap, d_offsets, d_status, 0,
                         restart_syms, d_restart);
}

extern "C" int rcb_encode_chunks_restart(rcb_ctx* c, const void* d_syms, uint64_t n_syms, int sym_bytes,
                                         uint64_t chunk_syms, const rcb_model* m, uint8_t* d_out, uint64_t out_cap,
                                         uint64_t* d_offsets, uint32_t* d_status, uint64_t restart_syms,
                                         rcb_restart_point* d_restart, uint64_t* h_out_bytes) {
    int r = encode_launch(c, d_syms, n_syms, sym_bytes, chunk_syms, m, d_out, out_cap, d_offsets, d_status, 0,
                          restart_syms, d_restart);
    if (r) return r;
    uint64_t need = 0;
    r = encode_result(c, h_out_bytes, &need);
    if (r == RCB_ERR_OUT_CAPACITY && need) {  // staging row too small: rerun once (rewrites the restart points too)
        r = encode_launch(c, d_syms, n_syms, sym_bytes, chunk_syms, m, d_out, out_cap, d_offsets, d_status, need,
                          restart_syms, d_restart);
        if (r) return r;
        r = encode_result(c, h_out_bytes, nullptr);
    }
    return r;
}

extern "C" int rcb_encode_result(rcb_ctx* c, uint64_t* h_out_bytes) {
    if (!c || !c->pending_offsets) return RCB_ERR_INVALID_ARGUMENT;
    return encode_result(c, h_out_bytes, nullptr);
}

// ------------------------------------------------------------------- decode
// Launch geometry and kernel flavour of one decode call.
struct DecPlan {
    int kind;       // 0: fat-LUT fused kernel / generic (rcb_decode.cuh), 1: row kernel (rcb_decode_row.cuh)
    int table;      // TAB_*
    int fmode;      // FM_* (row kernel), FM_BIG / FM_GENERIC (kind 0)
    bool checked, pow2, lut16, m2;
    bool tp;        // fused loop tuned for throughput (rcb_decode.cuh: TP)
    int threads;
    uint32_t lanes, nb;
    size_t smem;
};

// parts: lanes per chunk (restart points; 1 = whole chunks).  p.lanes is a multiple of it for the row kernel.
static DecPlan plan_decode(const rcb_ctx* c, const rcb_model* m, uint64_t n_chunks, uint32_t parts = 1) {
    DecPlan p;
    memset(&p, 0, sizeof p);
    const bool shared = m->n_models == 1;
    p.checked = (m->bad_bits & 4u) != 0;
    const bool regular = (m->bad_bits & 8u) == 0;
    p.threads = pick_threads(c, c->dec_threads, n_chunks * parts);
    p.lanes = (uint32_t)p.threads;
    {
        // several warps per scheduler (restart points, many chunks): the throughput flavour of the fused loop
        const char* w = getenv("RCB_DEC_TP");
        const uint64_t sched = 4ull * (uint64_t)(c->sm_count > 0 ? c->sm_count : 148);
        p.tp = w ? atoi(w) != 0 : n_chunks * parts >= 2 * 32 * sched;
    }
    const size_t budget = 216 * 1024;
    const size_t row = ((size_t)m->K + ROW_PAD) * sizeof(uint32_t);  // rcb_decode_row.cuh
    p.lut16 = m->K > 256;
    const size_t lut_elem = p.lut16 ? 2 : 1;
    if (shared) {
        p.table = TAB_SHARED;
        p.pow2 = (m->h_hdr0.flags & MODEL_POW2) != 0;
        const uint32_t shift = m->h_hdr0.div.shift;
        const uint64_t total = m->h_hdr0.div.total;
        if (!p.checked && regular) {
            // fat LUT (two candidates per bucket of total / nb, nb <= 4096 buckets): only when no bucket can hold
            // two boundaries, i.e. every frequency >= one bucket + the 1/8 estimate margin.  Any total: a power
            // of two >= 2^24 folds the division into the renormalisation shift (FM_BIG), smaller powers of two
            // take two shifts (FM_POW2), everything else the divide-free step (FM_GEN here = FUSE_GEN with cs).
            const uint64_t nbk = m->h_hdr0.nb ? m->h_hdr0.nb : 1;
            const uint64_t bucket = (total + nbk - 1) / nbk;  // bucket width, rounded up
            bool fat_ok = total >= 2 && (uint64_t)m->min_c >= bucket + (bucket >> 3) + 1 && !getenv("RCB_NO_FAT");
            if (!p.pow2 && ((m->h_hdr0.flags & MODEL_FULLC) || getenv("RCB_NO_GENCS"))) fat_ok = false;
            if (fat_ok) {
                p.kind = 0;
                p.fmode = p.pow2 ? (shift >= 24 ? FM_BIG : FM_POW2) : FM_GEN;
                // the throughput flavour may run up to DEC_TP_THREADS lanes per block (20 warps per SM fit its registers)
                if (p.tp) p.threads = pick_threads(c, c->dec_threads, n_chunks * parts, DEC_TP_THREADS);
                // general totals >= 2^25: one table-wide reciprocal; smaller ones: per-candidate constants
                p.m2 = p.fmode == FM_GEN && recip2_ok((uint32_t)total) && !getenv("RCB_NO_M2");
                // FUSED kernel: candidates (16 bytes) + reciprocals of their frequencies (8 bytes) per bucket
                // (+ their 64-bit reciprocal constants, 16 bytes, for general totals < 2^25)
                const size_t fixed = (size_t)LUT_CAP * (sizeof(LutEntry) + 8 + (p.fmode == FM_GEN && !p.m2 ? 16 : 0)) +
                                     (size_t)m->K * sizeof(uint2);
                while (p.threads > 32 && (size_t)p.threads * RING_STRIDE + fixed > budget) p.threads >>= 1;
                p.lanes = (uint32_t)p.threads;
                p.smem = (size_t)p.threads * RING_STRIDE + fixed;
                return p;
            }
            // thin LUT over the cum row: smallest power of two with bucket width <= min c (<= 65536 buckets)
            uint64_t want = m->min_c ? (total + m->min_c - 1) / m->min_c : 65536;
            uint32_t nb = 256;
            while (nb < want && nb < 65536) nb <<= 1;
            while (nb > 256 && (size_t)p.threads * RING_STRIDE + row + (size_t)nb * lut_elem > budget) nb >>= 1;
            if ((size_t)p.threads * RING_STRIDE + row + (size_t)nb * lut_elem <= budget) {
                p.kind = 1;
                p.nb = nb;
                p.fmode = p.pow2 ? (shift >= 24 ? FM_BIG : FM_POW2) : FM_GEN;
                p.smem = (size_t)p.threads * RING_STRIDE + row + (size_t)nb * lut_elem;
                p.lanes = (uint32_t)p.threads / parts * parts;  // whole chunks per block
                return p;
            }
        }
        // inconsistent / irregular / oversized tables: generic kernel (bucket LUT when regular, exact search)
        p.kind = 0;
        p.fmode = FM_GENERIC;
        const uint32_t nbg = (m->h_hdr0.flags & MODEL_REGULAR) ? m->h_hdr0.nb : 0u;
        p.smem = (size_t)p.threads * RING_STRIDE + (size_t)nbg * sizeof(LutEntry) + (size_t)m->K * sizeof(uint2);
        return p;
    }
    // per-chunk models: each chunk owns a cum row + a thin LUT in shared memory (shared by its `parts` lanes,
    // each of which has its own code-byte ring) when they fit
    const size_t chunk_min = (size_t)parts * RING_STRIDE + row + 256 * lut_elem;
    const uint32_t cmax = (uint32_t)(budget / chunk_min);
    if (!p.checked && regular && (uint64_t)cmax * parts >= 32) {
        uint32_t CB = cmax * parts < 512u ? cmax : 512u / parts;  // chunks per block
        const uint64_t sms = (uint64_t)(c->sm_count > 0 ? c->sm_count : 148);
        // whole waves of SM-count blocks: the chunks are spread evenly over as few waves as the cap allows
        const uint64_t waves = (n_chunks + sms * CB - 1) / (sms * CB);
        const uint64_t even = (n_chunks + waves * sms - 1) / (waves ? waves * sms : 1);
        if (even <= CB) CB = (uint32_t)(even ? even : 1);
        if (c->dec_threads && (uint32_t)c->dec_threads < CB * parts) CB = (uint32_t)c->dec_threads / parts;
        if (CB == 0) CB = 1;
        const uint32_t L = CB * parts;
        p.threads = (int)((L + 31) / 32 * 32);
        const size_t fixed = (size_t)p.threads * RING_STRIDE + (size_t)CB * row;
        size_t per_row = (budget - fixed) / CB / lut_elem;
        uint32_t nb = (uint32_t)(per_row > 4096 ? 4096 : per_row) & ~15u;
        p.kind = 1;
        p.table = TAB_LANE;
        p.fmode = FM_LANE;
        p.lanes = L;
        p.nb = nb;
        p.smem = fixed + (size_t)CB * nb * lut_elem;
        return p;
    }
    p.kind = 0;
    p.table = TAB_GLOBAL;
    p.fmode = FM_GENERIC;
    p.smem = (size_t)p.threads * RING_STRIDE;
    return p;
}

template <typename SYM>
static void launch_decode_variant(rcb_ctx* c, const rcb_model* m, const DecodeArgs& a, const DecPlan& p,
                                  unsigned blocks) {
    auto go = [&](auto kern) {
        if (p.smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        kern<<<blocks, p.threads, p.smem, c->stream>>>(a);
    };
    (void)m;
    if (p.table == TAB_SHARED) {
        if (p.fmode == FM_BIG && p.tp) go(decode_kernel<SYM, true, true, false, FUSE_BIG, true>);
        else if (p.fmode == FM_POW2 && p.tp) go(decode_kernel<SYM, true, true, false, FUSE_POW2, true>);
        else if (p.fmode == FM_GEN && p.m2 && p.tp) go(decode_kernel<SYM, true, false, false, FUSE_GEN_M2, true>);
        else if (p.fmode == FM_GEN && p.tp) go(decode_kernel<SYM, true, false, false, FUSE_GEN, true>);
        else if (p.fmode == FM_BIG) go(decode_kernel<SYM, true, true, false, FUSE_BIG>);
        else if (p.fmode == FM_POW2) go(decode_kernel<SYM, true, true, false, FUSE_POW2>);
        else if (p.fmode == FM_GEN && p.m2) go(decode_kernel<SYM, true, false, false, FUSE_GEN_M2>);
        else if (p.fmode == FM_GEN) go(decode_kernel<SYM, true, false, false, FUSE_GEN>);
        else if (p.pow2) {
            if (p.checked) go(decode_kernel<SYM, true, true, true, -1>);
            else go(decode_kernel<SYM, true, true, false, -1>);
        } else {
            if (p.checked) go(decode_kernel<SYM, true, false, true, -1>);
            else go(decode_kernel<SYM, true, false, false, -1>);
        }
    } else {
        if (p.checked) go(decode_kernel<SYM, false, false, true, -1>);
        else go(decode_kernel<SYM, false, false, false, -1>);
    }
}

template <typename SYM, typename LUT_T>
static void launch_decode_row(rcb_ctx* c, const DecodeRowArgs& a, const DecPlan& p, unsigned blocks) {
    auto go = [&](auto kern) {
        if (p.smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        kern<<<blocks, p.threads, p.smem, c->stream>>>(a);
    };
    if (p.table == TAB_LANE) {
        go(decode_row_kernel<SYM, TAB_LANE, FM_LANE, LUT_T>);
    } else {
        switch (p.fmode) {
            case FM_BIG: go(decode_row_kernel<SYM, TAB_SHARED, FM_BIG, LUT_T>); break;
            case FM_POW2: go(decode_row_kernel<SYM, TAB_SHARED, FM_POW2, LUT_T>); break;
            default: go(decode_row_kernel<SYM, TAB_SHARED, FM_GEN, LUT_T>); break;
        }
    }
}

// Issue decode + status summary for a run of chunks on the ctx's current stream (explicit buffers, see
// encode_issue).  d_offsets points at the run's first entry; offsets stay relative to d_stream.
static int decode_issue(rcb_ctx* c, const uint8_t* d_stream, const uint64_t* d_offsets, uint64_t n_syms,
                        int sym_bytes, uint64_t chunk_syms, const rcb_model* m, uint64_t model_first,
                        void* d_syms_out, uint32_t* status, unsigned long long* d_summary, bool timed,
                        const DecSegment* seg = nullptr, const RestartSpec* rs = nullptr) {
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    const bool shared = m->n_models == 1;
    DecodeArgs a;
    const bool restart = rs && rs->on() && !seg;
    const uint32_t parts = restart ? rs->per_chunk + 1u : 1u;
    a.restart = restart ? rs->pts : nullptr;
    a.restart_syms = restart ? rs->syms : 0;
    a.parts = parts;
    // several lanes per chunk report errors only (atomicMax): the status words start at 0
    if (restart) CK(c, cudaMemsetAsync(status, 0, n_chunks * sizeof(uint32_t), c->stream));
    if (seg) a.seg = *seg;
    else a.seg = DecSegment{0, 0, nullptr, 0u, 0u};
    a.stream = d_stream;
    a.offsets = d_offsets;
    a.n_syms = n_syms;
    a.chunk_syms = chunk_syms;
    a.n_chunks = n_chunks;
    a.tabs = shared ? m->d_tab : m->d_tab + model_first * m->K;
    a.hdrs = shared ? m->d_hdr : m->d_hdr + model_first;
    a.lut = m->d_lut;
    a.lut_cs = m->d_lut_cs;
    a.K = m->K;
    a.per_chunk = !shared;
    a.out = d_syms_out;
    a.status = status;
    DecPlan plan = plan_decode(c, m, n_chunks, parts);
    // the throughput flavour stores 16 bytes at a time: every lane's part of the output must start 16-byte aligned
    if ((chunk_syms * (uint64_t)sym_bytes) % 16 || (seg && (seg->first * (uint64_t)sym_bytes) % 16)) plan.tp = false;
    const uint64_t n_lanes = n_chunks * parts;
    const unsigned blocks = (unsigned)((n_lanes + plan.lanes - 1) / plan.lanes);  // row kernel: lanes % parts == 0
    {
        // start-up L2 prefetch per lane: 2 KiB when few lanes are resident, less when they are many, so that one
        // wave of blocks prefetches at most 32 MiB (a quarter of L2); under 256 bytes the ring pieces' own
        // 256-byte hints are the prefetch
        const uint64_t wave = (uint64_t)c->sm_count * plan.lanes;  // lanes resident at one time (>= 1 block per SM)
        const uint64_t resident = n_lanes < wave ? n_lanes : wave;
        uint64_t bytes = resident ? ((uint64_t)32 << 20) / resident : 2048;
        static int pf_env = -2;  // RCB_DEC_PF=<bytes>: override (measurements)
        if (pf_env == -2) {
            const char* e = getenv("RCB_DEC_PF");
            pf_env = e ? atoi(e) : -1;
        }
        if (pf_env >= 0) bytes = (uint64_t)pf_env;
        if (bytes > 2048) bytes = 2048;
        a.pf_words = (uint32_t)(bytes / 256) * 64;
    }
    if (plan.kind == 0) {
        if (sym_bytes == 1)
            launch_decode_variant<uint8_t>(c, m, a, plan, blocks);
        else
            launch_decode_variant<uint16_t>(c, m, a, plan, blocks);
    } else {
        DecodeRowArgs ra;
        ra.stream = a.stream;
        ra.offsets = a.offsets;
        ra.n_syms = a.n_syms;
        ra.chunk_syms = a.chunk_syms;
        ra.n_chunks = a.n_chunks;
        ra.tabs = a.tabs;
        ra.hdrs = a.hdrs;
        ra.K = a.K;
        ra.lanes_per_block = plan.lanes;
        ra.nb = plan.nb;
        ra.out = a.out;
        ra.status = a.status;
        ra.seg = a.seg;
        ra.restart = a.restart;
        ra.restart_syms = a.restart_syms;
        ra.parts = a.parts;
        ra.pf_words = a.pf_words;
        if (sym_bytes == 1) {
            if (plan.lut16) launch_decode_row<uint8_t, uint16_t>(c, ra, plan, blocks);
            else launch_decode_row<uint8_t, uint8_t>(c, ra, plan, blocks);
        } else {
            if (plan.lut16) launch_decode_row<uint16_t, uint16_t>(c, ra, plan, blocks);
            else launch_decode_row<uint16_t, uint8_t>(c, ra, plan, blocks);
        }
    }
    CK_LAUNCH(c);
    if (timed) EV(c, 5);
    if (seg && seg->state && seg->save) return RCB_OK;  // statuses exist after the last segment only
    status_summary_kernel<<<1, 1024, 0, c->stream>>>(status, n_chunks, d_summary);
    CK_LAUNCH(c);
    if (timed) EV(c, 6);
    return RCB_OK;
}

static int decode_launch(rcb_ctx* c, const uint8_t* d_stream, const uint64_t* d_offsets, uint64_t n_syms,
                         int sym_bytes, uint64_t chunk_syms, const rcb_model* m, void* d_syms_out,
                         uint32_t* d_status, uint64_t restart_syms, const void* d_restart) {
    if (!c || !m || !m->ready || m->ctx != c || !d_offsets || chunk_syms == 0) return RCB_ERR_INVALID_ARGUMENT;
    RestartSpec rs;
    if (int rr = make_restart_spec(chunk_syms, restart_syms, d_restart, &rs)) return rr;
    if (n_syms && (!d_stream || !d_syms_out)) return RCB_ERR_INVALID_ARGUMENT;
    if (sym_bytes != 1 && sym_bytes != 2) return RCB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(d_stream) & 15u) || (reinterpret_cast<uintptr_t>(d_syms_out) & 15u) ||
        (reinterpret_cast<uintptr_t>(d_offsets) & 7u))
        return RCB_ERR_INVALID_ARGUMENT;
    if (chunk_syms > 0x40000000ull) return RCB_ERR_UNSUPPORTED;  // same limit as the encoder
    if (n_syms > (1ull << 62)) return RCB_ERR_INVALID_ARGUMENT;  // n_syms * sym_bytes must not wrap
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    if (m->n_models != 1 && m->n_models != n_chunks) return RCB_ERR_INVALID_ARGUMENT;
    if (n_chunks > 0x7FFFFFFFull) return RCB_ERR_UNSUPPORTED;
    ON_DEVICE(c);
    if (n_chunks == 0) {
        CK(c, cudaMemsetAsync(c->d_summary + 4, 0, 4 * sizeof(unsigned long long), c->stream));
        return RCB_OK;
    }
    int r = ensure_chunks(c, n_chunks);
    if (r) return r;
    EV(c, 4);
    r = decode_issue(c, d_stream, d_offsets, n_syms, sym_bytes, chunk_syms, m, 0, d_syms_out,
                     d_status ? d_status : c->status, c->d_summary + 4, true, nullptr, &rs);
    if (r) return r;
    c->ev_dec = c->timing;
    return RCB_OK;
}

extern "C" int rcb_decode_chunks_async(rcb_ctx* c, const uint8_t* d_stream, const uint64_t* d_offsets,
                                       uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                                       const rcb_model* m, void* d_syms_out, uint32_t* d_status) {
    return decode_launch(c, d_stream, d_offsets, n_syms, sym_bytes, chunk_syms, m, d_syms_out, d_status, 0, nullptr);
}

extern "C" int rcb_decode_chunks_restart_async(rcb_ctx* c, const uint8_t* d_stream, const uint64_t* d_offsets,
                                               uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                                               const rcb_model* m, void* d_syms_out, uint32_t* d_status,
                                               uint64_t restart_syms, const rcb_restart_point* d_restart) {
    return decode_launch(c, d_stream, d_offsets, n_syms, sym_bytes, chunk_syms, m, d_syms_out, d_status,
                         restart_syms, d_restart);
}

extern "C" int rcb_decode_chunks_restart(rcb_ctx* c, const uint8_t* d_stream, const uint64_t* d_offsets,
                                         uint64_t n_syms, int sym_bytes, uint64_t chunk_syms, const rcb_model* m,
                                         void* d_syms_out, uint32_t* d_status, uint64_t restart_syms,
                                         const rcb_restart_point* d_restart) {
    int r = decode_launch(c, d_stream, d_offsets, n_syms, sym_bytes, chunk_syms, m, d_syms_out, d_status,
                          restart_syms, d_restart);
    if (r) return r;
    return rcb_decode_result(c);
}

extern "C" int rcb_decode_result(rcb_ctx* c) {
    if (!c) return RCB_ERR_INVALID_ARGUMENT;
    CK(c, cudaMemcpyAsync(c->h_summary + 4, c->d_summary + 4, 4 * sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (c->h_summary[4]) return status_to_error((uint32_t)c->h_summary[6]);
    return RCB_OK;
}

extern "C" int rcb_decode_chunks(rcb_ctx* c, const uint8_t* d_stream, const uint64_t* d_offsets,
                                 uint64_t n_syms, int sym_bytes, uint64_t chunk_syms, const rcb_model* m,
                                 void* d_syms_out, uint32_t* d_status) {
    int r = rcb_decode_chunks_async(c, d_stream, d_offsets, n_syms, sym_bytes, chunk_syms, m, d_syms_out,
                                    d_status);
    if (r) return r;
    return rcb_decode_result(c);
}

// ------------------------------------------------------- host-buffer wrappers
// H2D, code, D2H inside the call.  Large batches are cut into up to 8 slices of whole chunks, each on
// its own stream: slice k's copy-in overlaps slice k-1's kernels and slice k-2's copy-out, and the
// slices' kernels run side by side (a slice of the lanes takes as long as all of them: the coder is
// latency-bound per lane, so the kernels must overlap rather than queue).
static int ensure_slices(rcb_ctx* c, int n) {
    for (int i = 0; i < n; i++) {
        if (!c->slice_stream[i]) CK(c, cudaStreamCreateWithFlags(&c->slice_stream[i], cudaStreamNonBlocking));
        if (!c->in_ready[i]) CK(c, cudaEventCreateWithFlags(&c->in_ready[i], cudaEventDisableTiming));
        if (!c->out_ready[i]) CK(c, cudaEventCreateWithFlags(&c->out_ready[i], cudaEventDisableTiming));
    }
    if (!c->h2d_stream) CK(c, cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    if (!c->d2h_stream) CK(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    if (!c->d_slice_summary) {
        CK(c, cudaMalloc(&c->d_slice_summary, MAX_SLICES * 8 * sizeof(unsigned long long)));
        CK(c, cudaMallocHost(&c->h_slice_summary, MAX_SLICES * 8 * sizeof(unsigned long long)));
    }
    return RCB_OK;
}

// Writes a slice's summary (4 words) and its byte count into (mapped, pinned) host memory.
__global__ void publish_slice_kernel(unsigned long long* h_dst, const unsigned long long* d_sum,
                                     const unsigned long long* d_bytes) {
    for (int i = 0; i < 4; i++) h_dst[i] = d_sum[i];
    h_dst[4] = *d_bytes;
    __threadfence_system();
}

// RCB_TRACE=1: host timestamps of the pipeline stages on stderr (debugging aid)
static bool trace_on() {
    static int v = -1;
    if (v < 0) v = getenv("RCB_TRACE") ? 1 : 0;
    return v == 1;
}
static double now_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int pick_slices(uint64_t n_chunks, uint64_t bytes) {
    // more slices = shorter pipeline fill and drain (the first copy-in and the last copy-out run alone);
    // a slice still has to be a few MiB so that per-copy overheads stay invisible
    static int cap = 0;
    if (!cap) {
        const char* e = getenv("RCB_MAX_SLICES");
        cap = e ? atoi(e) : 8;  // measured: 16 slices 54.9 ms per round trip against 51.8 with 8 (1 GiB batches)
        if (cap < 1 || cap > MAX_SLICES) cap = 8;
    }
    int s = cap;
    while (s > 1 && (n_chunks / s < 64 || bytes / s < (8u << 20))) s >>= 1;
    return s;
}

static int encode_host_impl(rcb_ctx* c, const void* h_syms, uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                            const rcb_model* m, uint8_t* h_out, uint64_t out_cap, uint64_t* h_offsets,
                            uint64_t* h_out_bytes, uint64_t restart_syms, rcb_restart_point* h_restart) {
    if (!c || !m || !m->ready || chunk_syms == 0 || !h_offsets || !h_out_bytes) return RCB_ERR_INVALID_ARGUMENT;
    if (sym_bytes != 1 && sym_bytes != 2) return RCB_ERR_UNSUPPORTED;
    if (n_syms && !h_syms) return RCB_ERR_INVALID_ARGUMENT;
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    if (m->n_models != 1 && m->n_models != n_chunks) return RCB_ERR_INVALID_ARGUMENT;
    const uint64_t in_bytes = (n_syms * sym_bytes + 15) & ~15ull;
    const uint64_t pitch = staging_pitch(m, chunk_syms);
    const uint64_t bound = n_chunks * pitch + 32;
    // restart points: recorded on the device next to the offsets, copied out slice by slice
    RestartSpec rs;
    if (int rr = make_restart_spec(chunk_syms, restart_syms, h_restart, &rs)) return rr;
    const uint64_t R = rs.per_chunk;
    const uint64_t rs_bytes = (n_chunks * R * sizeof(Restart) + 15) & ~15ull;
    ON_DEVICE(c);
    const int S = pick_slices(n_chunks, n_syms * sym_bytes);
    const uint64_t off_bytes = ((n_chunks + 1 + S) * sizeof(uint64_t) + 15) & ~15ull;
    int r = ensure_h2d(c, in_bytes + bound + off_bytes + rs_bytes + 64);
    if (r) return r;
    uint8_t* d_in = (uint8_t*)c->h2d;
    uint8_t* d_out = d_in + in_bytes;
    uint64_t* d_off = (uint64_t*)(d_out + ((bound + 15) & ~15ull));
    Restart* d_rs = R ? (Restart*)((uint8_t*)d_off + off_bytes) : nullptr;
    if (S == 1 || (chunk_syms * sym_bytes) % 16 != 0) {
        if (n_syms)
            CK(c, cudaMemcpyAsync(d_in, h_syms, n_syms * sym_bytes, cudaMemcpyHostToDevice, c->stream));
        uint64_t total = 0;
        r = rcb_encode_chunks_restart(c, d_in, n_syms, sym_bytes, chunk_syms, m, d_out, bound, d_off, nullptr,
                                      R ? restart_syms : 0, (rcb_restart_point*)d_rs, &total);
        *h_out_bytes = total;
        if (r) return r;
        if (total > out_cap) return RCB_ERR_OUT_CAPACITY;
        CK(c, cudaMemcpyAsync(h_offsets, d_off, (n_chunks + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                              c->stream));
        if (total) CK(c, cudaMemcpyAsync(h_out, d_out, total, cudaMemcpyDeviceToHost, c->stream));
        if (R && n_chunks)
            CK(c, cudaMemcpyAsync(h_restart, d_rs, n_chunks * R * sizeof(Restart), cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        return RCB_OK;
    }
    // ---- sliced pipeline
    r = ensure_slices(c, S);
    if (r) return r;
    r = ensure_chunks(c, n_chunks);
    if (r) return r;
    r = ensure_staging(c, (size_t)(n_chunks * pitch + 64));
    if (r) return r;
    CK(c, cudaStreamSynchronize(c->stream));
    cudaStream_t user_stream = c->stream;
    uint64_t first[MAX_SLICES + 1];
    for (int k = 0; k <= S; k++) first[k] = n_chunks * k / S;
    int rc = RCB_OK;
    for (int k = 0; k < S && rc == RCB_OK; k++) {
        const uint64_t c0 = first[k], c1 = first[k + 1], nk = c1 - c0;
        const uint64_t s0 = c0 * chunk_syms, s1 = c1 * chunk_syms < n_syms ? c1 * chunk_syms : n_syms;
        c->stream = c->slice_stream[k];
        cudaError_t e = cudaMemcpyAsync(d_in + s0 * sym_bytes, (const uint8_t*)h_syms + s0 * sym_bytes,
                                        (s1 - s0) * sym_bytes, cudaMemcpyHostToDevice, c->h2d_stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->in_ready[k], c->h2d_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->slice_stream[k], c->in_ready[k], 0);
        if (e != cudaSuccess) {
            c->last_err = e;
            rc = RCB_ERR_CUDA;
            break;
        }
        // slice-local offsets (k extra slots before it), output region at the slice's worst-case position
        RestartSpec rsk = rs;
        if (R) rsk.pts = d_rs + c0 * R;
        rc = encode_issue(c, d_in + s0 * sym_bytes, s1 - s0, sym_bytes, chunk_syms, m, c0, c->staging + c0 * pitch,
                          pitch, c->lens + c0, c->status + c0, d_out + c0 * pitch, nk * pitch, d_off + c0 + k,
                          c->d_slice_summary + 8 * k, false, &rsk);
        if (rc) break;
        // the slice's status summary and byte count go straight into pinned host memory from a one-thread
        // kernel: a tiny device->host copy would queue behind the bulk copies on the copy engines
        publish_slice_kernel<<<1, 1, 0, c->stream>>>(c->h_slice_summary + 8 * k, c->d_slice_summary + 8 * k,
                                                     (const unsigned long long*)(d_off + c0 + k + nk));
        c->launches++;
        e = cudaGetLastError();
        if (e != cudaSuccess) {
            c->last_err = e;
            rc = RCB_ERR_CUDA;
        }
    }
    uint64_t base = 0;
    uint64_t bases[MAX_SLICES];
    bool staging_overflow = false;
    const double t_issue = now_ms();
    if (trace_on()) fprintf(stderr, "[rcb] encode_host: %d slices issued\n", S);
    for (int k = 0; k < S && rc == RCB_OK; k++) {
        const uint64_t c0 = first[k], nk = first[k + 1] - c0;
        cudaError_t e = cudaStreamSynchronize(c->slice_stream[k]);
        if (trace_on()) fprintf(stderr, "[rcb]   slice %d coded at +%.2f ms\n", k, now_ms() - t_issue);
        if (e != cudaSuccess) {
            c->last_err = e;
            rc = RCB_ERR_CUDA;
            break;
        }
        const unsigned long long* sum = c->h_slice_summary + 8 * k;
        if (sum[0]) {
            if ((uint32_t)sum[2] == ST_OUT_CAPACITY) staging_overflow = true;
            else rc = status_to_error((uint32_t)sum[2]);
            break;
        }
        const uint64_t bytes = sum[4];
        bases[k] = base;
        if (base + bytes <= out_cap) {
            // slice k's kernels are done (synchronised above): its bytes leave on the in-order D2H stream
            e = cudaMemcpyAsync(h_out + base, d_out + c0 * pitch, bytes, cudaMemcpyDeviceToHost, c->d2h_stream);
            if (e == cudaSuccess)
                // the slice's last entry is the next slice's first: only the final slice copies it
                e = cudaMemcpyAsync(h_offsets + c0, d_off + c0 + k, (nk + (k == S - 1 ? 1 : 0)) * sizeof(uint64_t),
                                    cudaMemcpyDeviceToHost, c->d2h_stream);
            if (e == cudaSuccess && R)
                e = cudaMemcpyAsync(h_restart + c0 * R, d_rs + c0 * R, nk * R * sizeof(Restart), cudaMemcpyDeviceToHost,
                                    c->d2h_stream);
            if (e != cudaSuccess) {
                c->last_err = e;
                rc = RCB_ERR_CUDA;
                break;
            }
        }
        base += bytes;
    }
    for (int k = 0; k < S; k++) cudaStreamSynchronize(c->slice_stream[k]);
    cudaStreamSynchronize(c->h2d_stream);
    cudaStreamSynchronize(c->d2h_stream);
    if (trace_on()) fprintf(stderr, "[rcb]   all copies done at +%.2f ms\n", now_ms() - t_issue);
    c->stream = user_stream;
    if (staging_overflow) {
        // a staging row was too small for this data (rare): the one-shot path sizes it exactly and reruns
        if (n_syms)
            CK(c, cudaMemcpyAsync(d_in, h_syms, n_syms * sym_bytes, cudaMemcpyHostToDevice, c->stream));
        uint64_t total = 0;
        r = rcb_encode_chunks_restart(c, d_in, n_syms, sym_bytes, chunk_syms, m, d_out, bound, d_off, nullptr,
                                      R ? restart_syms : 0, (rcb_restart_point*)d_rs, &total);
        *h_out_bytes = total;
        if (r) return r;
        if (total > out_cap) return RCB_ERR_OUT_CAPACITY;
        CK(c, cudaMemcpyAsync(h_offsets, d_off, (n_chunks + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                              c->stream));
        if (total) CK(c, cudaMemcpyAsync(h_out, d_out, total, cudaMemcpyDeviceToHost, c->stream));
        if (R) CK(c, cudaMemcpyAsync(h_restart, d_rs, n_chunks * R * sizeof(Restart), cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        return RCB_OK;
    }
    if (rc) return rc;
    *h_out_bytes = base;
    if (base > out_cap) return RCB_ERR_OUT_CAPACITY;
    // slice-local offsets -> offsets into the concatenated stream (slice k's last entry is slice k+1's first)
    for (int k = S - 1; k >= 0; k--) {
        const uint64_t c0 = first[k], nk = first[k + 1] - c0;
        if (k == S - 1) h_offsets[c0 + nk] += bases[k];
        for (uint64_t i = 0; i < nk; i++) h_offsets[c0 + i] += bases[k];
    }
    return RCB_OK;
}

extern "C" int rcb_encode_host(rcb_ctx* c, const void* h_syms, uint64_t n_syms, int sym_bytes,
                               uint64_t chunk_syms, const rcb_model* m, uint8_t* h_out, uint64_t out_cap,
                               uint64_t* h_offsets, uint64_t* h_out_bytes) {
    return encode_host_impl(c, h_syms, n_syms, sym_bytes, chunk_syms, m, h_out, out_cap, h_offsets, h_out_bytes, 0,
                            nullptr);
}

extern "C" int rcb_encode_host_restart(rcb_ctx* c, const void* h_syms, uint64_t n_syms, int sym_bytes,
                                       uint64_t chunk_syms, const rcb_model* m, uint8_t* h_out, uint64_t out_cap,
                                       uint64_t* h_offsets, uint64_t* h_out_bytes, uint64_t restart_syms,
                                       rcb_restart_point* h_restart) {
    return encode_host_impl(c, h_syms, n_syms, sym_bytes, chunk_syms, m, h_out, out_cap, h_offsets, h_out_bytes,
                            restart_syms, h_restart);
}

static int decode_host_impl(rcb_ctx* c, const uint8_t* h_stream, const uint64_t* h_offsets, uint64_t n_syms,
                            int sym_bytes, uint64_t chunk_syms, const rcb_model* m, void* h_syms_out,
                            uint64_t restart_syms, const rcb_restart_point* h_restart) {
    if (!c || !m || !m->ready || chunk_syms == 0 || !h_offsets) return RCB_ERR_INVALID_ARGUMENT;
    if (sym_bytes != 1 && sym_bytes != 2) return RCB_ERR_UNSUPPORTED;
    if (n_syms && (!h_stream || !h_syms_out)) return RCB_ERR_INVALID_ARGUMENT;
    if (chunk_syms > 0x40000000ull) return RCB_ERR_UNSUPPORTED;
    if (n_syms > (1ull << 62)) return RCB_ERR_INVALID_ARGUMENT;
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    if (m->n_models != 1 && m->n_models != n_chunks) return RCB_ERR_INVALID_ARGUMENT;
    if (n_chunks > 0x7FFFFFFFull) return RCB_ERR_UNSUPPORTED;
    const uint64_t total = h_offsets[n_chunks];
    // the offsets size the copies below: they must be monotone and end at `total`.  (A chunk shorter
    // than the 8 bytes Decoder::new pops is reported per chunk by the kernels as a truncated stream.)
    for (uint64_t i = 0; i < n_chunks; i++)
        if (h_offsets[i] > h_offsets[i + 1]) return RCB_ERR_INVALID_ARGUMENT;
    const uint64_t st_bytes = (total + 47) & ~15ull;
    const uint64_t off_bytes = ((n_chunks + 1) * sizeof(uint64_t) + 15) & ~15ull;
    const uint64_t out_bytes = (n_syms * sym_bytes + 15) & ~15ull;
    RestartSpec rs;
    if (int rr = make_restart_spec(chunk_syms, restart_syms, h_restart, &rs)) return rr;
    const uint64_t R = rs.per_chunk;
    const uint64_t rs_bytes = (n_chunks * R * sizeof(Restart) + 15) & ~15ull;
    ON_DEVICE(c);
    int r = ensure_h2d(c, st_bytes + off_bytes + out_bytes + rs_bytes + 64);
    if (r) return r;
    uint8_t* d_st = (uint8_t*)c->h2d;
    uint64_t* d_off = (uint64_t*)(d_st + st_bytes);
    uint8_t* d_out = (uint8_t*)d_off + off_bytes;
    Restart* d_rs = R ? (Restart*)(d_out + out_bytes) : nullptr;
    if (R && n_chunks)  // a few bytes per chunk: one copy ahead of everything else
        CK(c, cudaMemcpyAsync(d_rs, h_restart, n_chunks * R * sizeof(Restart), cudaMemcpyHostToDevice, c->stream));
    const int S = pick_slices(n_chunks, n_syms * sym_bytes);
    if (S == 1 || (chunk_syms * sym_bytes) % 16 != 0) {
        if (total) CK(c, cudaMemcpyAsync(d_st, h_stream, total, cudaMemcpyHostToDevice, c->stream));
        CK(c, cudaMemcpyAsync(d_off, h_offsets, (n_chunks + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice,
                              c->stream));
        r = rcb_decode_chunks_restart(c, d_st, d_off, n_syms, sym_bytes, chunk_syms, m, d_out, nullptr,
                                      R ? restart_syms : 0, (const rcb_restart_point*)d_rs);
        if (r) return r;
        if (n_syms)
            CK(c, cudaMemcpyAsync(h_syms_out, d_out, n_syms * sym_bytes, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        return RCB_OK;
    }
    // ---- sliced pipeline
    r = ensure_slices(c, S);
    if (r) return r;
    r = ensure_chunks(c, n_chunks);
    if (r) return r;
    CK(c, cudaMemcpyAsync(d_off, h_offsets, (n_chunks + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    // segments per chunk: 4 when a chunk is big enough for the split to pay (>= 32 KiB, quarter a multiple
    // of 16 symbols so that every segment starts on a word of output)
    // (with restart points the parts of a chunk are decoded side by side instead: one launch per slice)
    const int P = (!R && chunk_syms * sym_bytes >= (32u << 10) && chunk_syms % 64 == 0) ? 4 : 1;
    const uint64_t seg_syms = chunk_syms / P;
    if (P > 1 && c->resume_cap < n_chunks) {
        cudaFree(c->d_resume);
        c->d_resume = nullptr;
        c->resume_cap = 0;
        CK(c, cudaMalloc(&c->d_resume, n_chunks * sizeof(DecResume)));
        c->resume_cap = n_chunks;
    }
    cudaStream_t user_stream = c->stream;
    int rc = RCB_OK;
    uint64_t first[MAX_SLICES + 1];
    for (int k = 0; k <= S; k++) first[k] = n_chunks * k / S;
    cudaEvent_t tr_start = nullptr, tr_k[MAX_SLICES][4], tr_c[MAX_SLICES][4], tr_in[MAX_SLICES];
    if (trace_on()) {
        cudaEventCreate(&tr_start);
        for (int k = 0; k < MAX_SLICES; k++) {
            cudaEventCreate(&tr_in[k]);
            for (int q = 0; q < 4; q++) {
                cudaEventCreate(&tr_k[k][q]);
                cudaEventCreate(&tr_c[k][q]);
            }
        }
        cudaEventRecord(tr_start, c->h2d_stream);
    }
    for (int k = 0; k < S && rc == RCB_OK; k++) {
        const uint64_t c0 = first[k], c1 = first[k + 1];
        const uint64_t s0 = c0 * chunk_syms, s1 = c1 * chunk_syms < n_syms ? c1 * chunk_syms : n_syms;
        const uint64_t b0 = h_offsets[c0], b1 = h_offsets[c1];
        if (b1 < b0 || b1 > total) {
            rc = RCB_ERR_INVALID_ARGUMENT;
            break;
        }
        c->stream = c->slice_stream[k];
        cudaError_t e = cudaSuccess;
        if (b1 > b0) e = cudaMemcpyAsync(d_st + b0, h_stream + b0, b1 - b0, cudaMemcpyHostToDevice, c->h2d_stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->in_ready[k], c->h2d_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->slice_stream[k], c->in_ready[k], 0);
        if (trace_on()) cudaEventRecord(tr_in[k], c->h2d_stream);
        if (e != cudaSuccess) {
            c->last_err = e;
            rc = RCB_ERR_CUDA;
            break;
        }
        // each chunk is decoded in P launches so that the copy-out of its first symbols starts after 1/P of
        // the (per-chunk, latency-bound) decode time instead of all of it
        const uint64_t chunk_bytes = chunk_syms * sym_bytes;
        const uint64_t full = (s1 - s0) / chunk_syms;            // complete chunks of this slice
        const uint64_t tail_syms = (s1 - s0) - full * chunk_syms;  // ragged last chunk (0: none)
        for (int ph = 0; ph < P && rc == RCB_OK; ph++) {
            const uint64_t f0 = seg_syms * ph;
            DecSegment seg{f0, seg_syms, P > 1 ? c->d_resume + c0 : nullptr, ph > 0 ? 1u : 0u, ph < P - 1 ? 1u : 0u};
            RestartSpec rsk = rs;
            if (R) rsk.pts = d_rs + c0 * R;
            rc = decode_issue(c, d_st, d_off + c0, s1 - s0, sym_bytes, chunk_syms, m, c0, d_out + s0 * sym_bytes,
                              c->status + c0, c->d_slice_summary + 8 * k, false, R ? nullptr : &seg, R ? &rsk : nullptr);
            if (rc) break;
            if (trace_on()) cudaEventRecord(tr_k[k][ph], c->stream);
            e = cudaEventRecord(c->out_ready[k], c->stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(c->d2h_stream, c->out_ready[k], 0);
            const uint64_t w_syms = ph < P - 1 ? seg_syms : chunk_syms - f0;  // this segment's width
            uint8_t* hdst = (uint8_t*)h_syms_out + (s0 + f0) * sym_bytes;
            const uint8_t* dsrc = d_out + (s0 + f0) * sym_bytes;
            if (e == cudaSuccess && full) {
                if (P == 1)
                    e = cudaMemcpyAsync(hdst, dsrc, full * chunk_bytes, cudaMemcpyDeviceToHost, c->d2h_stream);
                else
                    e = cudaMemcpy2DAsync(hdst, chunk_bytes, dsrc, chunk_bytes, w_syms * sym_bytes, full,
                                          cudaMemcpyDeviceToHost, c->d2h_stream);
            }
            if (e == cudaSuccess && tail_syms > f0) {  // the ragged chunk's part of this segment
                const uint64_t n_t = (tail_syms - f0 < w_syms ? tail_syms - f0 : w_syms) * sym_bytes;
                e = cudaMemcpyAsync(hdst + full * chunk_bytes, dsrc + full * chunk_bytes, n_t, cudaMemcpyDeviceToHost,
                                    c->d2h_stream);
            }
            if (trace_on()) cudaEventRecord(tr_c[k][ph], c->d2h_stream);
            if (e != cudaSuccess) {
                c->last_err = e;
                rc = RCB_ERR_CUDA;
            }
        }
        if (rc) break;
        // status summary straight into pinned host memory (see rcb_encode_host)
        publish_slice_kernel<<<1, 1, 0, c->stream>>>(c->h_slice_summary + 8 * k, c->d_slice_summary + 8 * k,
                                                     c->d_slice_summary + 8 * k);
        c->launches++;
        e = cudaGetLastError();
        if (e != cudaSuccess) {
            c->last_err = e;
            rc = RCB_ERR_CUDA;
        }
    }
    for (int k = 0; k < S + 2; k++) {
        cudaError_t e = cudaStreamSynchronize(k < S ? c->slice_stream[k] : (k == S ? c->h2d_stream : c->d2h_stream));
        if (e != cudaSuccess && rc == RCB_OK) {
            c->last_err = e;
            rc = RCB_ERR_CUDA;
        }
    }
    c->stream = user_stream;
    if (trace_on()) {
        for (int k = 0; k < S; k++) {
            float t_in = 0;
            cudaEventElapsedTime(&t_in, tr_start, tr_in[k]);
            fprintf(stderr, "[rcb] decode_host slice %d: in %.2f |", k, t_in);
            for (int q = 0; q < P; q++) {
                float a = 0, b = 0;
                cudaEventElapsedTime(&a, tr_start, tr_k[k][q]);
                cudaEventElapsedTime(&b, tr_start, tr_c[k][q]);
                fprintf(stderr, " seg%d kernel %.2f copied %.2f |", q, a, b);
            }
            fprintf(stderr, "\n");
        }
        cudaEventDestroy(tr_start);
        for (int k = 0; k < MAX_SLICES; k++) {
            cudaEventDestroy(tr_in[k]);
            for (int q = 0; q < 4; q++) {
                cudaEventDestroy(tr_k[k][q]);
                cudaEventDestroy(tr_c[k][q]);
            }
        }
    }
    if (rc) return rc;
    for (int k = 0; k < S; k++)
        if (c->h_slice_summary[8 * k]) return status_to_error((uint32_t)c->h_slice_summary[8 * k + 2]);
    return RCB_OK;
}

extern "C" int rcb_decode_host(rcb_ctx* c, const uint8_t* h_stream, const uint64_t* h_offsets,
                               uint64_t n_syms, int sym_bytes, uint64_t chunk_syms, const rcb_model* m,
                               void* h_syms_out) {
    return decode_host_impl(c, h_stream, h_offsets, n_syms, sym_bytes, chunk_syms, m, h_syms_out, 0, nullptr);
}

extern "C" int rcb_decode_host_restart(rcb_ctx* c, const uint8_t* h_stream, const uint64_t* h_offsets,
                                       uint64_t n_syms, int sym_bytes, uint64_t chunk_syms, const rcb_model* m,
                                       void* h_syms_out, uint64_t restart_syms, const rcb_restart_point* h_restart) {
    return decode_host_impl(c, h_stream, h_offsets, n_syms, sym_bytes, chunk_syms, m, h_syms_out, restart_syms,
                            h_restart);
}

// ----------------------------------------------------------- synthetic data
extern "C" int rcb_generate(rcb_ctx* c, void* d_out, uint64_t first, uint64_t n, int sym_bytes, uint32_t K,
                            uint64_t seed, const uint32_t* h_thr, uint32_t n_tables, uint64_t chunk_syms) {
    if (!c || !d_out || !h_thr || K < 1 || K > MAX_K || n_tables == 0) return RCB_ERR_INVALID_ARGUMENT;
    if (sym_bytes != 1 && sym_bytes != 2) return RCB_ERR_UNSUPPORTED;
    if (sym_bytes == 1 && K > 256) return RCB_ERR_INVALID_ARGUMENT;
    if (n_tables > 1 && chunk_syms == 0) return RCB_ERR_INVALID_ARGUMENT;
    const size_t thr_bytes = (size_t)n_tables * (K - 1) * sizeof(uint32_t);
    if (thr_bytes > 200 * 1024) return RCB_ERR_UNSUPPORTED;
    if (n == 0) return RCB_OK;
    ON_DEVICE(c);
    uint32_t* d_thr = nullptr;
    CK(c, cudaMalloc(&d_thr, thr_bytes ? thr_bytes : 4));
    cudaError_t e = cudaMemcpyAsync(d_thr, h_thr, thr_bytes, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        const int threads = 256;
        uint64_t groups = (n + (16 / sym_bytes) - 1) / (16 / sym_bytes);
        uint64_t want = (groups + threads - 1) / threads;
        uint64_t maxb = (uint64_t)c->sm_count * 16;
        unsigned blocks = (unsigned)(want > maxb ? maxb : want);
        if (sym_bytes == 1) {
            auto kern = generate_kernel<uint8_t>;
            if (thr_bytes > 48 * 1024)
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)thr_bytes);
            kern<<<blocks, threads, thr_bytes, c->stream>>>((uint8_t*)d_out, first, n, K, seed, d_thr, n_tables,
                                                             chunk_syms ? chunk_syms : 1);
        } else {
            auto kern = generate_kernel<uint16_t>;
            if (thr_bytes > 48 * 1024)
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)thr_bytes);
            kern<<<blocks, threads, thr_bytes, c->stream>>>((uint16_t*)d_out, first, n, K, seed, d_thr, n_tables,
                                                              chunk_syms ? chunk_syms : 1);
        }
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_thr);
    if (e != cudaSuccess) {
        c->last_err = e;
        return RCB_ERR_CUDA;
    }
    return RCB_OK;
}

// ------------------------------------------------ continued single-stream coder
static_assert(sizeof(rcb_stream_state) == sizeof(StreamState), "rcb_stream_state layout");

extern "C" void rcb_stream_state_init(rcb_stream_state* st) {
    if (!st) return;
    memset(st, 0, sizeof *st);
    st->range = ~0ull;  // src/range_coder.rs:13-20
}

extern "C" int rcb_encode_stream(rcb_ctx* c, rcb_stream_state* st, const void* h_syms, uint64_t n_syms,
                                 int sym_bytes, const rcb_model* m, uint8_t* h_out, uint64_t out_cap,
                                 uint64_t* h_n_out, uint32_t* h_per_symbol, int finish) {
    if (!c || !st || !h_n_out) return RCB_ERR_INVALID_ARGUMENT;
    if (n_syms && (!m || !m->ready || m->ctx != c)) return RCB_ERR_INVALID_ARGUMENT;  // finish alone: no model
    if (m && m->n_models != 1) return RCB_ERR_UNSUPPORTED;
    if (sym_bytes != 1 && sym_bytes != 2) return RCB_ERR_UNSUPPORTED;
    if ((n_syms && !h_syms) || (out_cap && !h_out)) return RCB_ERR_INVALID_ARGUMENT;
    ON_DEVICE(c);
    const size_t sym_b = (size_t)((n_syms * sym_bytes + 15) & ~15ull);
    const size_t out_b = (size_t)((out_cap + 15) & ~15ull);
    const size_t per_b = (size_t)((n_syms * 4 + 15) & ~15ull);
    int r = ensure_h2d(c, 64 + sym_b + out_b + per_b + 16);
    if (r) return r;
    uint8_t* base = (uint8_t*)c->h2d;
    StreamState* d_st = (StreamState*)base;
    uint64_t* d_n = (uint64_t*)(base + 48);
    uint8_t* d_syms = base + 64;
    uint8_t* d_out = d_syms + sym_b;
    uint32_t* d_per = (uint32_t*)(d_out + out_b);
    CK(c, cudaMemcpyAsync(d_st, st, sizeof(StreamState), cudaMemcpyHostToDevice, c->stream));
    if (n_syms) CK(c, cudaMemcpyAsync(d_syms, h_syms, n_syms * sym_bytes, cudaMemcpyHostToDevice, c->stream));
    encode_stream_kernel<<<1, 1, 0, c->stream>>>(d_st, d_syms, n_syms, sym_bytes, m ? m->d_tab : nullptr,
                                                  m ? m->d_hdr : nullptr, m ? m->K : 0u, d_out, out_cap, d_n,
                                                  h_per_symbol ? d_per : nullptr, finish);
    CK_LAUNCH(c);
    uint64_t n_out = 0;
    CK(c, cudaMemcpyAsync(st, d_st, sizeof(StreamState), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(&n_out, d_n, sizeof n_out, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    *h_n_out = n_out;
    if (st->status) return status_to_error(st->status);
    if (n_out) CK(c, cudaMemcpyAsync(h_out, d_out, n_out, cudaMemcpyDeviceToHost, c->stream));
    if (h_per_symbol && n_syms)
        CK(c, cudaMemcpyAsync(h_per_symbol, d_per, n_syms * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RCB_OK;
}

extern "C" int rcb_decode_stream(rcb_ctx* c, rcb_stream_state* st, const uint8_t* h_code, uint64_t code_len,
                                 uint64_t n_syms, int sym_bytes, const rcb_model* m, void* h_syms_out) {
    if (!c || !st || !m || !m->ready || m->ctx != c) return RCB_ERR_INVALID_ARGUMENT;
    if (m->n_models != 1) return RCB_ERR_UNSUPPORTED;
    if (sym_bytes != 1 && sym_bytes != 2) return RCB_ERR_UNSUPPORTED;
    if ((code_len && !h_code) || (n_syms && !h_syms_out)) return RCB_ERR_INVALID_ARGUMENT;
    ON_DEVICE(c);
    // only the unread tail of the code travels (the caller's per-symbol loop would otherwise copy the
    // whole stream once per symbol)
    const uint64_t code_base = st->consumed < code_len ? st->consumed : code_len;
    const size_t code_b = (size_t)((code_len - code_base + 15) & ~15ull);
    const size_t sym_b = (size_t)((n_syms * sym_bytes + 15) & ~15ull);
    int r = ensure_h2d(c, 64 + code_b + sym_b + 16);
    if (r) return r;
    uint8_t* base = (uint8_t*)c->h2d;
    StreamState* d_st = (StreamState*)base;
    uint8_t* d_code = base + 64;
    uint8_t* d_syms = d_code + code_b;
    CK(c, cudaMemcpyAsync(d_st, st, sizeof(StreamState), cudaMemcpyHostToDevice, c->stream));
    if (code_len > code_base)
        CK(c, cudaMemcpyAsync(d_code, h_code + code_base, code_len - code_base, cudaMemcpyHostToDevice, c->stream));
    decode_stream_kernel<<<1, 1, 0, c->stream>>>(d_st, d_code, code_base, code_len, n_syms, sym_bytes, m->d_tab,
                                                  m->d_hdr, m->K, d_syms);
    CK_LAUNCH(c);
    CK(c, cudaMemcpyAsync(st, d_st, sizeof(StreamState), cudaMemcpyDeviceToHost, c->stream));
    if (n_syms)
        CK(c, cudaMemcpyAsync(h_syms_out, d_syms, n_syms * sym_bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (st->status) return status_to_error(st->status);
    return RCB_OK;
}

// ------------------------------------------------------------ framed container
// Host-side byte layout only (see rcb200.h); no symbol is coded here.
namespace {
const uint64_t FRAME_HDR = 56;
inline uint64_t al8(uint64_t x) { return (x + 7) & ~7ull; }
inline uint64_t frame_model_bytes(uint32_t K, uint64_t n_chunks, int per_chunk) {
    return per_chunk ? al8(n_chunks * (uint64_t)K * 4) : al8(8 + (uint64_t)K * 8);
}
template <typename T>
inline void put_le(uint8_t* p, T v) { memcpy(p, &v, sizeof v); }  // x86-64 / aarch64 hosts are little-endian
template <typename T>
inline T get_le(const uint8_t* p) {
    T v;
    memcpy(&v, p, sizeof v);
    return v;
}
}  // namespace

extern "C" uint64_t rcb_frame_bound(uint32_t K, uint64_t n_chunks, int per_chunk, uint64_t payload_bytes) {
    return FRAME_HDR + frame_model_bytes(K, n_chunks, per_chunk) + (n_chunks + 1) * 8 + payload_bytes;
}

extern "C" uint64_t rcb_frame_bound_restart(uint32_t K, uint64_t n_chunks, int per_chunk, uint64_t payload_bytes,
                                            uint64_t chunk_syms, uint64_t restart_syms) {
    return rcb_frame_bound(K, n_chunks, per_chunk, payload_bytes) +
           n_chunks * rcb_restart_points_per_chunk(chunk_syms, restart_syms) * sizeof(Restart);
}

extern "C" int rcb_frame_write(rcb_ctx* c, const rcb_model* m, int sym_bytes, uint64_t chunk_syms, uint64_t n_syms,
                               const uint8_t* h_stream, const uint64_t* h_offsets, uint8_t* h_frame,
                               uint64_t frame_cap, uint64_t* h_frame_bytes) {
    return rcb_frame_write_restart(c, m, sym_bytes, chunk_syms, n_syms, h_stream, h_offsets, 0, nullptr, h_frame,
                                   frame_cap, h_frame_bytes);
}

extern "C" int rcb_frame_write_restart(rcb_ctx* c, const rcb_model* m, int sym_bytes, uint64_t chunk_syms,
                                       uint64_t n_syms, const uint8_t* h_stream, const uint64_t* h_offsets,
                                       uint64_t restart_syms, const rcb_restart_point* h_restart, uint8_t* h_frame,
                                       uint64_t frame_cap, uint64_t* h_frame_bytes) {
    if (!c || !m || !m->ready || !h_offsets || !h_frame || !h_frame_bytes || chunk_syms == 0)
        return RCB_ERR_INVALID_ARGUMENT;
    RestartSpec rs;
    if (int rr = make_restart_spec(chunk_syms, restart_syms, h_restart, &rs)) return rr;
    const uint64_t R = rs.per_chunk;
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    const int per_chunk = m->n_models != 1;
    if (per_chunk && m->n_models != n_chunks) return RCB_ERR_INVALID_ARGUMENT;
    if (per_chunk && (m->bad_bits & 8u)) return RCB_ERR_UNSUPPORTED;  // per-chunk section stores c only
    const uint64_t payload = h_offsets[n_chunks];
    if (payload && !h_stream) return RCB_ERR_INVALID_ARGUMENT;
    const uint64_t rs_bytes = n_chunks * R * sizeof(Restart);
    const uint64_t need = rcb_frame_bound(m->K, n_chunks, per_chunk, payload) + rs_bytes;
    *h_frame_bytes = need;
    if (need > frame_cap) return RCB_ERR_OUT_CAPACITY;
    uint8_t* p = h_frame;
    memcpy(p, "RCB2", 4);
    put_le<uint32_t>(p + 4, R ? 2 : 1);  // version 2 = a restart section follows the offsets
    put_le<uint32_t>(p + 8, (uint32_t)sym_bytes);
    put_le<uint32_t>(p + 12, m->K);
    put_le<uint32_t>(p + 16, (uint32_t)per_chunk);
    put_le<uint32_t>(p + 20, R ? (uint32_t)(restart_syms / 64) : 0);
    put_le<uint64_t>(p + 24, chunk_syms);
    put_le<uint64_t>(p + 32, n_syms);
    put_le<uint64_t>(p + 40, n_chunks);
    put_le<uint64_t>(p + 48, payload);
    uint8_t* ms = p + FRAME_HDR;
    ON_DEVICE(c);
    const size_t n_ent = (size_t)m->n_models * m->K;
    uint2* tmp = (uint2*)malloc(n_ent * sizeof(uint2));
    if (!tmp) return RCB_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaMemcpyAsync(tmp, m->d_tab, n_ent * sizeof(uint2), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        free(tmp);
        c->last_err = e;
        return RCB_ERR_CUDA;
    }
    memset(ms, 0, frame_model_bytes(m->K, n_chunks, per_chunk));
    if (!per_chunk) {
        put_le<uint32_t>(ms, m->h_hdr0.div.total);
        for (uint32_t i = 0; i < m->K; i++) {
            put_le<uint32_t>(ms + 8 + 4 * (uint64_t)i, tmp[i].x);
            put_le<uint32_t>(ms + 8 + 4 * ((uint64_t)m->K + i), tmp[i].y);
        }
    } else {
        for (size_t i = 0; i < n_ent; i++) put_le<uint32_t>(ms + 4 * i, tmp[i].y);
    }
    free(tmp);
    uint8_t* os = ms + frame_model_bytes(m->K, n_chunks, per_chunk);
    memcpy(os, h_offsets, (n_chunks + 1) * 8);
    uint8_t* rsec = os + (n_chunks + 1) * 8;
    if (rs_bytes) memcpy(rsec, h_restart, rs_bytes);  // 24-byte little-endian records, as in memory
    if (payload) memcpy(rsec + rs_bytes, h_stream, payload);
    return RCB_OK;
}

extern "C" int rcb_frame_parse(const uint8_t* h_frame, uint64_t len, rcb_frame_info* info) {
    if (!h_frame || !info || len < FRAME_HDR) return RCB_ERR_INVALID_ARGUMENT;
    if (memcmp(h_frame, "RCB2", 4) != 0) return RCB_ERR_INVALID_ARGUMENT;
    rcb_frame_info f;
    f.version = get_le<uint32_t>(h_frame + 4);
    f.sym_bytes = get_le<uint32_t>(h_frame + 8);
    f.K = get_le<uint32_t>(h_frame + 12);
    f.model_mode = get_le<uint32_t>(h_frame + 16);
    f.chunk_syms = get_le<uint64_t>(h_frame + 24);
    f.n_syms = get_le<uint64_t>(h_frame + 32);
    f.n_chunks = get_le<uint64_t>(h_frame + 40);
    f.payload_bytes = get_le<uint64_t>(h_frame + 48);
    const uint32_t restart_units = get_le<uint32_t>(h_frame + 20);
    if ((f.version != 1 && f.version != 2) || (f.sym_bytes != 1 && f.sym_bytes != 2) || f.K == 0 || f.K > MAX_K ||
        f.model_mode > 1 || f.chunk_syms == 0)
        return RCB_ERR_INVALID_ARGUMENT;
    // version 1 has no restart section (the field is reserved, 0); version 2 has one
    if ((f.version == 1) != (restart_units == 0)) return RCB_ERR_INVALID_ARGUMENT;
    f.restart_syms = (uint64_t)restart_units * 64;
    f.restart_off = 0;
    // untrusted header: bound everything before it enters any arithmetic (n_syms * sym_bytes and the
    // ceil-division below must not wrap; chunk_syms has the coder's own limit)
    if (f.n_syms > (1ull << 62) || f.chunk_syms > 0x40000000ull) return RCB_ERR_INVALID_ARGUMENT;
    if (f.n_chunks != (f.n_syms + f.chunk_syms - 1) / f.chunk_syms) return RCB_ERR_INVALID_ARGUMENT;
    if (f.n_chunks > 0x7FFFFFFFull || f.payload_bytes > len) return RCB_ERR_INVALID_ARGUMENT;
    f.model_off = FRAME_HDR;
    f.offsets_off = f.model_off + frame_model_bytes(f.K, f.n_chunks, (int)f.model_mode);
    f.payload_off = f.offsets_off + (f.n_chunks + 1) * 8;
    if (f.restart_syms) {
        const uint64_t per = rcb_restart_points_per_chunk(f.chunk_syms, f.restart_syms);
        if (per == 0 || per > 63) return RCB_ERR_INVALID_ARGUMENT;  // same limits as the coder entry points
        f.restart_off = f.payload_off;
        f.payload_off += f.n_chunks * per * sizeof(Restart);  // n_chunks < 2^31, per < 64: no wrap
    }
    f.frame_bytes = f.payload_off + f.payload_bytes;
    if (f.frame_bytes > len) return RCB_ERR_TRUNCATED_STREAM;
    // offsets: start at 0, every chunk holds at least the 8 bytes of Encoder::finish
    // (src/encoder.rs:40-46), end at payload_bytes
    uint64_t prev = 0;
    for (uint64_t i = 0; i <= f.n_chunks; i++) {
        const uint64_t o = get_le<uint64_t>(h_frame + f.offsets_off + 8 * i);
        if (o > f.payload_bytes || (i == 0 ? o != 0 : o < prev + 8)) return RCB_ERR_INVALID_ARGUMENT;
        prev = o;
    }
    if (prev != f.payload_bytes) return RCB_ERR_INVALID_ARGUMENT;
    *info = f;
    return RCB_OK;
}

extern "C" int rcb_frame_model(rcb_ctx* c, const uint8_t* h_frame, const rcb_frame_info* f, rcb_model** out) {
    if (!c || !h_frame || !f || !out) return RCB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    const uint64_t n_models = f->model_mode ? f->n_chunks : 1;
    if (n_models == 0) return RCB_ERR_INVALID_ARGUMENT;
    rcb_model* m = nullptr;
    int r = rcb_model_create(c, f->K, n_models, &m);
    if (r) return r;
    const uint8_t* ms = h_frame + f->model_off;
    const size_t n_ent = (size_t)n_models * f->K;
    uint32_t* cc = (uint32_t*)malloc(n_ent * 4);
    uint32_t* cum = (uint32_t*)malloc(n_ent * 4);
    uint32_t* tot = (uint32_t*)malloc(n_models * 4);
    if (!cc || !cum || !tot) {
        free(cc);
        free(cum);
        free(tot);
        rcb_model_destroy(m);
        return RCB_ERR_INVALID_ARGUMENT;
    }
    if (!f->model_mode) {
        tot[0] = get_le<uint32_t>(ms);
        for (uint32_t i = 0; i < f->K; i++) {
            cum[i] = get_le<uint32_t>(ms + 8 + 4 * (uint64_t)i);
            cc[i] = get_le<uint32_t>(ms + 8 + 4 * ((uint64_t)f->K + i));
        }
    } else {
        for (uint64_t j = 0; j < n_models; j++) {
            uint32_t run = 0;  // examples/sample_impl.rs:61-69
            for (uint32_t i = 0; i < f->K; i++) {
                const uint32_t v = get_le<uint32_t>(ms + 4 * (j * f->K + i));
                cc[j * f->K + i] = v;
                cum[j * f->K + i] = run;
                run += v;
            }
            tot[j] = run;
        }
    }
    r = rcb_model_from_tables(c, m, cc, cum, tot);
    free(cc);
    free(cum);
    free(tot);
    if (r) {
        rcb_model_destroy(m);
        return r;
    }
    *out = m;
    return RCB_OK;
}

extern "C" int rcb_frame_encode_host_restart(rcb_ctx* c, const void* h_syms, uint64_t n_syms, int sym_bytes,
                                             uint64_t chunk_syms, const rcb_model* m, uint64_t restart_syms,
                                             uint8_t* h_frame, uint64_t frame_cap, uint64_t* h_frame_bytes) {
    if (!c || !m || !m->ready || !h_frame || !h_frame_bytes || chunk_syms == 0) return RCB_ERR_INVALID_ARGUMENT;
    if (restart_syms % 64) return RCB_ERR_INVALID_ARGUMENT;
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    const uint64_t cap = rcb_encode_bound(c, m, n_syms, sym_bytes, chunk_syms) + 64;
    uint8_t* stream = (uint8_t*)malloc(cap);
    uint64_t* offs = (uint64_t*)malloc((n_chunks + 1) * 8);
    if (!stream || !offs) {
        free(stream);
        free(offs);
        return RCB_ERR_INVALID_ARGUMENT;
    }
    uint64_t bytes = 0;
    rcb_restart_point* pts = nullptr;
    const uint64_t R = rcb_restart_points_per_chunk(chunk_syms, restart_syms);
    if (R) {
        pts = (rcb_restart_point*)malloc((size_t)(n_chunks * R ? n_chunks * R : 1) * sizeof(rcb_restart_point));
        if (!pts) {
            free(stream);
            free(offs);
            return RCB_ERR_INVALID_ARGUMENT;
        }
    }
    int r = rcb_encode_host_restart(c, h_syms, n_syms, sym_bytes, chunk_syms, m, stream, cap, offs, &bytes,
                                    R ? restart_syms : 0, pts);
    if (r == RCB_OK)
        r = rcb_frame_write_restart(c, m, sym_bytes, chunk_syms, n_syms, stream, offs, R ? restart_syms : 0, pts,
                                    h_frame, frame_cap, h_frame_bytes);
    free(stream);
    free(offs);
    free(pts);
    return r;
}

extern "C" int rcb_frame_encode_host(rcb_ctx* c, const void* h_syms, uint64_t n_syms, int sym_bytes,
                                     uint64_t chunk_syms, const rcb_model* m, uint8_t* h_frame, uint64_t frame_cap,
                                     uint64_t* h_frame_bytes) {
    return rcb_frame_encode_host_restart(c, h_syms, n_syms, sym_bytes, chunk_syms, m, 0, h_frame, frame_cap,
                                         h_frame_bytes);
}

extern "C" int rcb_frame_decode_host(rcb_ctx* c, const uint8_t* h_frame, uint64_t len, void* h_syms_out,
                                     uint64_t out_cap_bytes, uint64_t* h_n_syms) {
    rcb_frame_info f;
    int r = rcb_frame_parse(h_frame, len, &f);
    if (r) return r;
    if (h_n_syms) *h_n_syms = f.n_syms;
    if (f.n_syms > out_cap_bytes / f.sym_bytes) return RCB_ERR_OUT_CAPACITY;  // no product: cannot wrap
    if (f.n_syms == 0) return RCB_OK;
    rcb_model* m = nullptr;
    r = rcb_frame_model(c, h_frame, &f, &m);
    if (r) return r;
    // payload padded to a multiple of 16 bytes for the device reader; offsets may be unaligned in the frame
    uint8_t* stream = (uint8_t*)malloc((size_t)f.payload_bytes + 64);
    uint64_t* offs = (uint64_t*)malloc((size_t)(f.n_chunks + 1) * 8);
    if (!stream || !offs) {
        free(stream);
        free(offs);
        rcb_model_destroy(m);
        return RCB_ERR_INVALID_ARGUMENT;
    }
    memcpy(stream, h_frame + f.payload_off, (size_t)f.payload_bytes);
    memset(stream + f.payload_bytes, 0, 64);
    memcpy(offs, h_frame + f.offsets_off, (size_t)(f.n_chunks + 1) * 8);
    rcb_restart_point* pts = nullptr;
    if (f.restart_syms) {  // records may be unaligned in the frame
        const size_t nb = (size_t)(f.payload_off - f.restart_off);
        pts = (rcb_restart_point*)malloc(nb ? nb : 1);
        if (pts) memcpy(pts, h_frame + f.restart_off, nb);
    }
    r = (f.restart_syms && !pts)
            ? RCB_ERR_INVALID_ARGUMENT
            : rcb_decode_host_restart(c, stream, offs, f.n_syms, (int)f.sym_bytes, f.chunk_syms, m, h_syms_out,
                                      f.restart_syms, pts);
    free(pts);
    free(stream);
    free(offs);
    rcb_model_destroy(m);
    return r;
}

// ------------------------------------------- adaptive-per-symbol model (f4)
static bool adaptive_params_ok(const rcb_adaptive_params* p) {
    return p && p->K >= 1 && p->K <= 4096 && p->inc >= 1 && p->limit >= p->K &&
           (uint64_t)p->limit + p->inc <= 65535ull;
}

static uint64_t adaptive_pitch(uint64_t chunk_syms) {
    // a symbol costs at most log2(total / 1) <= 16 bits, plus the loop-2 truncation (< 1 %) and the flush
    const uint64_t p = chunk_syms * 2 + chunk_syms / 32 + 96;
    return (p + 15) & ~15ull;
}

extern "C" uint64_t rcb_adaptive_encode_bound(const rcb_adaptive_params* p, uint64_t n_syms, uint64_t chunk_syms) {
    if (!adaptive_params_ok(p) || chunk_syms == 0) return 0;
    return (n_syms + chunk_syms - 1) / chunk_syms * adaptive_pitch(chunk_syms) + 16;
}

// lanes per block and the per-lane table pitch (u16 counts[K] + u16 tree[K+1], odd number of words)
static void adaptive_geometry(const rcb_ctx* c, const rcb_adaptive_params* p, uint64_t n_chunks, size_t ring_bytes,
                              AdaptiveArgs& a, int& threads, size_t& smem) {
    uint32_t words = (2 * p->K + 1 + 1) / 2;
    if ((words & 1u) == 0) words++;
    const size_t per_lane = (size_t)words * 4 + ring_bytes;
    uint32_t L = (uint32_t)((200 * 1024) / per_lane);
    if (L > 512) L = 512;
    const uint64_t sms = (uint64_t)(c->sm_count > 0 ? c->sm_count : 148);
    const uint64_t even = (n_chunks + sms - 1) / sms;  // one wave over all SMs when it fits
    if (even <= L) L = (uint32_t)(even ? even : 1);
    a.lanes_per_block = L;
    a.pitch_words = words;
    threads = (int)((L + 31) / 32 * 32);
    smem = (size_t)L * words * 4 + (size_t)threads * ring_bytes;
}

extern "C" int rcb_adaptive_encode_chunks(rcb_ctx* c, const void* d_syms, uint64_t n_syms, int sym_bytes,
                                          uint64_t chunk_syms, const rcb_adaptive_params* p, uint8_t* d_out,
                                          uint64_t out_cap, uint64_t* d_offsets, uint32_t* d_status,
                                          uint64_t* h_out_bytes) {
    if (!c || !d_offsets || chunk_syms == 0 || (n_syms && !d_syms)) return RCB_ERR_INVALID_ARGUMENT;
    if (sym_bytes != 1 && sym_bytes != 2) return RCB_ERR_UNSUPPORTED;
    if (!adaptive_params_ok(p) || (sym_bytes == 1 && p->K > 256)) return RCB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(d_syms) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u) ||
        (reinterpret_cast<uintptr_t>(d_offsets) & 7u))
        return RCB_ERR_INVALID_ARGUMENT;
    if (chunk_syms > 0x40000000ull) return RCB_ERR_UNSUPPORTED;
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    if (n_chunks > 0x7FFFFFFFull) return RCB_ERR_UNSUPPORTED;
    ON_DEVICE(c);
    if (n_chunks == 0) {
        CK(c, cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t), c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        if (h_out_bytes) *h_out_bytes = 0;
        return RCB_OK;
    }
    int r = ensure_chunks(c, n_chunks);
    if (r) return r;
    uint64_t pitch = adaptive_pitch(chunk_syms);
    for (int attempt = 0; attempt < 2; attempt++) {
        if (pitch > 0xFFFFFFF0ull) return RCB_ERR_UNSUPPORTED;
        r = ensure_staging(c, (size_t)(n_chunks * pitch + 64));
        if (r) return r;
        AdaptiveArgs a;
        memset(&a, 0, sizeof a);
        a.K = p->K;
        a.inc = p->inc;
        a.limit = p->limit;
        a.n_syms = n_syms;
        a.chunk_syms = chunk_syms;
        a.n_chunks = n_chunks;
        a.syms = d_syms;
        a.staging = c->staging;
        a.pitch = pitch;
        a.lens = c->lens;
        a.status = d_status ? d_status : c->status;
        int threads;
        size_t smem;
        adaptive_geometry(c, p, n_chunks, 0, a, threads, smem);
        const unsigned blocks = (unsigned)((n_chunks + a.lanes_per_block - 1) / a.lanes_per_block);
        auto go = [&](auto kern) {
            if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<blocks, threads, smem, c->stream>>>(a);
        };
        if (sym_bytes == 1) go(adaptive_encode_kernel<uint8_t>);
        else go(adaptive_encode_kernel<uint16_t>);
        CK_LAUNCH(c);
        scan_lengths_kernel<<<1, 1024, 0, c->stream>>>(c->lens, a.status, n_chunks, d_offsets, c->d_summary);
        CK_LAUNCH(c);
        gather_kernel<<<(unsigned)n_chunks, 256, 0, c->stream>>>(c->staging, pitch, c->lens, d_offsets, d_out, out_cap);
        CK_LAUNCH(c);
        c->pending_out_cap = out_cap;
        c->pending_offsets = d_offsets;
        c->pending_n_chunks = n_chunks;
        uint64_t need = 0;
        r = encode_result(c, h_out_bytes, &need);
        if (r == RCB_ERR_OUT_CAPACITY && need && attempt == 0) {
            pitch = need;  // a staging row was too small for this data: once more with the exact need
            continue;
        }
        return r;
    }
    return RCB_ERR_OUT_CAPACITY;
}

extern "C" int rcb_adaptive_decode_chunks(rcb_ctx* c, const uint8_t* d_stream, const uint64_t* d_offsets,
                                          uint64_t n_syms, int sym_bytes, uint64_t chunk_syms,
                                          const rcb_adaptive_params* p, void* d_syms_out, uint32_t* d_status) {
    if (!c || !d_offsets || chunk_syms == 0) return RCB_ERR_INVALID_ARGUMENT;
    if (n_syms && (!d_stream || !d_syms_out)) return RCB_ERR_INVALID_ARGUMENT;
    if (sym_bytes != 1 && sym_bytes != 2) return RCB_ERR_UNSUPPORTED;
    if (!adaptive_params_ok(p) || (sym_bytes == 1 && p->K > 256)) return RCB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(d_stream) & 15u) || (reinterpret_cast<uintptr_t>(d_syms_out) & 15u) ||
        (reinterpret_cast<uintptr_t>(d_offsets) & 7u))
        return RCB_ERR_INVALID_ARGUMENT;
    if (chunk_syms > 0x40000000ull || n_syms > (1ull << 62)) return RCB_ERR_UNSUPPORTED;
    const uint64_t n_chunks = (n_syms + chunk_syms - 1) / chunk_syms;
    if (n_chunks > 0x7FFFFFFFull) return RCB_ERR_UNSUPPORTED;
    ON_DEVICE(c);
    if (n_chunks == 0) return RCB_OK;
    int r = ensure_chunks(c, n_chunks);
    if (r) return r;
    AdaptiveArgs a;
    memset(&a, 0, sizeof a);
    a.K = p->K;
    a.inc = p->inc;
    a.limit = p->limit;
    a.n_syms = n_syms;
    a.chunk_syms = chunk_syms;
    a.n_chunks = n_chunks;
    a.stream = d_stream;
    a.offsets = d_offsets;
    a.out = d_syms_out;
    a.status = d_status ? d_status : c->status;
    int threads;
    size_t smem;
    adaptive_geometry(c, p, n_chunks, RING_STRIDE, a, threads, smem);
    const unsigned blocks = (unsigned)((n_chunks + a.lanes_per_block - 1) / a.lanes_per_block);
    auto go = [&](auto kern) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<blocks, threads, smem, c->stream>>>(a);
    };
    if (sym_bytes == 1) go(adaptive_decode_kernel<uint8_t>);
    else go(adaptive_decode_kernel<uint16_t>);
    CK_LAUNCH(c);
    status_summary_kernel<<<1, 1024, 0, c->stream>>>(a.status, n_chunks, c->d_summary + 4);
    CK_LAUNCH(c);
    return rcb_decode_result(c);
}

// ------------------------------------------------------------------ multi-GPU
#include "rcb_comm.cuh"
