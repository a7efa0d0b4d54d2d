// rcb_kernels.cuh -- sm_100a kernels of the chunk-parallel range coder.
//
//   K1 hist_global_shared_kernel / hist_chunks_kernel   examples/sample_impl.rs:58-60,78-80
//      (hist_global_kernel, hist_global_u8_kernel: the per-warp-copy forms, RCB_HIST_SHARED=0)
//   K2 counts_to_tables_kernel / finalize_models_kernel  examples/sample_impl.rs:61-69
//   K3 encode_kernel                              src/encoder.rs:24-46, src/range_coder.rs:53-135
//   K4 scan_lengths_kernel + gather_kernel        (no reference analogue: one VecDeque there)
//   K5 decode_kernel                              src/decoder.rs:14-54, examples/sample_impl.rs:27-45
//   generate_kernel                               synthetic benchmark inputs
//
// Layout in HBM
//   symbols   [n_chunks][chunk_syms] u8 / u16-LE, dense, chunk i at i*chunk_syms
//   tables    [n_models][K] uint2 {cum, c}; headers [n_models] ModelHdr; LUT [nb] (shared model)
//   staging   [n_chunks][pitch] bytes, pitch % 16 == 0 (ctx scratch)
//   lengths   [n_chunks] u32; offsets [n_chunks+1] u64; stream dense bytes
// One coder state (lower_bound, range) per lane; lane l of a block owns chunk
// blockIdx.x * blockDim.x + l from its first to its last symbol.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rcb_core.cuh"
#include "rcb_encode.cuh"
#include "rcb_decode.cuh"

namespace rcb {

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v += t;
    }
    return v;
}

__device__ __forceinline__ unsigned long long warp_incl_scan64(unsigned long long v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v += t;
    }
    return v;
}

// =============================================================== K1 histogram
// Global histogram.  Each warp owns a private copy of the K bins in shared
// memory (bins of one copy are K consecutive words), merged into the u64
// global table with one atomic per non-zero bin per block.
template <typename SYM>
__global__ void __launch_bounds__(256) hist_global_kernel(const SYM* __restrict__ syms, uint64_t n,
                                                          uint32_t K, unsigned long long* counts,
                                                          uint32_t* bad) {
    extern __shared__ uint32_t s_hist[];  // [warps][K]
    const uint32_t warps = blockDim.x >> 5, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < warps * K; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    uint32_t* h = s_hist + warp * K;
    constexpr uint32_t PER = 16 / sizeof(SYM);
    const uint64_t nvec = n / PER;
    const uint4* v = reinterpret_cast<const uint4*>(syms);
    uint32_t oob = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 q = ldg_stream_v4(v + i);
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (sizeof(SYM) == 1) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    uint32_t s = (w[j] >> (8 * b)) & 0xFFu;
                    if (s < K) atomicAdd(&h[s], 1u); else oob = 1;
                }
            } else {
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    uint32_t s = (w[j] >> (16 * b)) & 0xFFFFu;
                    if (s < K) atomicAdd(&h[s], 1u); else oob = 1;
                }
            }
        }
    }
    // tail symbols (n not a multiple of the vector width)
    if (blockIdx.x == 0) {
        for (uint64_t i = nvec * PER + threadIdx.x; i < n; i += blockDim.x) {
            uint32_t s = syms[i];
            if (s < K) atomicAdd(&h[s], 1u); else oob = 1;
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < K; b += blockDim.x) {
        unsigned long long t = 0;
        for (uint32_t w = 0; w < warps; w++) t += s_hist[w * K + b];
        if (t) atomicAdd(&counts[b], t);
    }
    if (oob) atomicOr(bad, 1u);
}

// Byte symbols (K <= 256): each warp keeps REP interleaved copies of its bins
// (word = bin * REP + (lane % REP)), so the lanes that hit the same hot bin in one
// instruction -- 6 of 32 for the top symbol of Zipf(1.1) -- spread over REP words and banks
// instead of serialising on one shared-memory atomic.  The kernel is bound by the wavefronts of
// those atomics: with REP copies, 32 / REP lanes share a bank group, and two of them collide
// whenever their bins agree modulo 32 / REP (REP = 8: ~2.1 wavefronts per instruction, 16: ~1.5,
// 32: one -- every lane owns a bank).  Fewer warps fit as REP grows (REP * K * 4 bytes per warp), so
// NLOAD 16-byte loads per thread are in flight to keep HBM busy.
constexpr uint32_t HIST_REP = 8;  // hist_chunks / default geometry
template <bool FULL, uint32_t REP, int NLOAD>  // FULL: K == 256, no range check needed for a byte
__global__ void __launch_bounds__(256) hist_global_u8_kernel(const uint8_t* __restrict__ syms, uint64_t n,
                                                             uint32_t K, unsigned long long* counts,
                                                             uint32_t* bad) {
    extern __shared__ uint32_t s_hist[];  // [warps][K][REP]
    const uint32_t warps = blockDim.x >> 5, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < warps * K * REP; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    uint32_t* h = s_hist + warp * K * REP + (threadIdx.x & (REP - 1));
    const uint64_t nvec = n / 16;
    const uint4* v = reinterpret_cast<const uint4*>(syms);
    uint32_t oob = 0;
    auto count_word = [&](uint32_t w) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint32_t s = (w >> (8 * b)) & 0xFFu;
            if (FULL || s < K) atomicAdd(&h[s * REP], 1u); else oob = 1;
        }
    };
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (NLOAD - 1) * stride < nvec; i += NLOAD * stride) {
        uint4 q[NLOAD];
#pragma unroll
        for (int k = 0; k < NLOAD; k++) q[k] = ldg_stream_v4(v + i + k * stride);
#pragma unroll
        for (int k = 0; k < NLOAD; k++) {
            count_word(q[k].x);
            count_word(q[k].y);
            count_word(q[k].z);
            count_word(q[k].w);
        }
    }
    for (; i < nvec; i += stride) {
        const uint4 q0 = ldg_stream_v4(v + i);
        count_word(q0.x);
        count_word(q0.y);
        count_word(q0.z);
        count_word(q0.w);
    }
    if (blockIdx.x == 0) {  // tail symbols (n not a multiple of the vector width)
        for (uint64_t j = nvec * 16 + threadIdx.x; j < n; j += blockDim.x) {
            const uint32_t s = syms[j];
            if (FULL || s < K) atomicAdd(&h[s * REP], 1u); else oob = 1;
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < K; b += blockDim.x) {
        unsigned long long t = 0;
        for (uint32_t w = 0; w < warps; w++)
#pragma unroll
            for (uint32_t r = 0; r < REP; r++) t += s_hist[(w * K + b) * REP + r];
        if (t) atomicAdd(&counts[b], t);
    }
    if (oob) atomicOr(bad, 1u);
}

// One copy of the bins per BLOCK, REP interleaved words per bin (word = bin * REP + lane % REP), shared by every
// warp of the block: shared-memory atomics only collide inside one instruction, so warps need no private copy --
// what a copy per warp bought, REP buys for 1 / warps of the shared memory.  Byte symbols: REP = 32 (every lane
// owns a bank: one wavefront per atomic whatever the data), 32 KiB per block, several 512-thread blocks per SM.
// u16 symbols: REP = 32768 / K words (K = 4096: 8), 128 KiB, one block per SM.  NLOAD 16-byte loads per thread
// are in flight (one load per thread left HBM at 0.77 TB/s for the 4096-symbol alphabet).
template <typename SYM, bool FULL, int NLOAD>  // FULL: every value of SYM is < K
__global__ void __launch_bounds__(512) hist_global_shared_kernel(const SYM* __restrict__ syms, uint64_t n, uint32_t K,
                                                                 uint32_t rep_log2, unsigned long long* counts,
                                                                 uint32_t* bad) {
    extern __shared__ uint32_t s_hist[];  // [K + 1][REP]: bin K collects the symbols >= K (no branch per symbol)
    const uint32_t REP = 1u << rep_log2;
    const uint32_t nbins = FULL ? K : K + 1;
    for (uint32_t i = threadIdx.x; i < (nbins << rep_log2); i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    uint32_t* h = s_hist + (threadIdx.x & (REP - 1));
    constexpr uint32_t PER = 16 / sizeof(SYM);
    const uint64_t nvec = n / PER;
    const uint4* v = reinterpret_cast<const uint4*>(syms);
    auto count_word = [&](uint32_t w) {
        if (sizeof(SYM) == 1) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t s = (w >> (8 * b)) & 0xFFu;
                atomicAdd(&h[(FULL ? s : min(s, K)) << rep_log2], 1u);
            }
        } else {
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const uint32_t s = (w >> (16 * b)) & 0xFFFFu;
                atomicAdd(&h[(FULL ? s : min(s, K)) << rep_log2], 1u);
            }
        }
    };
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (NLOAD - 1) * stride < nvec; i += NLOAD * stride) {
        uint4 q[NLOAD];
#pragma unroll
        for (int k = 0; k < NLOAD; k++) q[k] = ldg_stream_v4(v + i + k * stride);
#pragma unroll
        for (int k = 0; k < NLOAD; k++) {
            count_word(q[k].x);
            count_word(q[k].y);
            count_word(q[k].z);
            count_word(q[k].w);
        }
    }
    for (; i < nvec; i += stride) {
        const uint4 q0 = ldg_stream_v4(v + i);
        count_word(q0.x);
        count_word(q0.y);
        count_word(q0.z);
        count_word(q0.w);
    }
    if (blockIdx.x == 0) {  // tail symbols (n not a multiple of the vector width)
        for (uint64_t j = nvec * PER + threadIdx.x; j < n; j += blockDim.x) {
            const uint32_t s = syms[j];
            atomicAdd(&h[(FULL ? s : min(s, K)) << rep_log2], 1u);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < K; b += blockDim.x) {
        unsigned long long t = 0;
        // rotate the start by the bin so that the lanes of a warp read different banks
        for (uint32_t r = 0; r < REP; r++) t += s_hist[(b << rep_log2) + ((r + b) & (REP - 1))];
        if (t) atomicAdd(&counts[b], t);
    }
    if (!FULL && threadIdx.x < REP && s_hist[(K << rep_log2) + threadIdx.x]) atomicOr(bad, 1u);
}

// One block per chunk: u32 counts[n_chunks][K].  The block's warps share one copy of the bins with REP interleaved
// words per bin (word = bin * REP + lane % REP; REP = 8 at K = 256, the shared memory of a copy per warp): a chunk
// of one dominant symbol (Zipf(5): 96 %) serialises 4 lanes per atomic instead of 32.  The host picks REP.
// Branch-free like hist_global_shared_kernel (symbols >= K land in bin K; FULL: every value of SYM is < K): the
// kernel is bound by instruction issue (15.8 warp instructions per symbol with a range-check branch per symbol).
template <typename SYM, bool FULL>
__global__ void __launch_bounds__(256) hist_chunks_kernel(const SYM* __restrict__ syms, uint64_t n,
                                                          uint64_t chunk_syms, uint32_t K, uint32_t rep_log2,
                                                          uint32_t* counts, uint32_t* bad) {
    extern __shared__ uint32_t s_hist[];  // [K (+ 1)][REP]
    const uint32_t REP = 1u << rep_log2;
    const uint32_t nbins = FULL ? K : K + 1;
    const uint64_t chunk = blockIdx.x;
    const uint64_t first = chunk * chunk_syms;
    const uint64_t cnt = (n - first < chunk_syms) ? (n - first) : chunk_syms;
    for (uint32_t i = threadIdx.x; i < (nbins << rep_log2); i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    uint32_t* h = s_hist + (threadIdx.x & (REP - 1));
    const SYM* p = syms + first;
    constexpr uint32_t PER = 16 / sizeof(SYM);
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
    uint64_t done = 0;
    auto count_sym = [&](uint32_t s) { atomicAdd(&h[(FULL ? s : min(s, K)) << rep_log2], 1u); };
    if (vec_ok) {
        const uint64_t nvec = cnt / PER;
        const uint4* v = reinterpret_cast<const uint4*>(p);
        auto count_word = [&](uint32_t w) {
            if (sizeof(SYM) == 1) {
#pragma unroll
                for (int b = 0; b < 4; b++) count_sym((w >> (8 * b)) & 0xFFu);
            } else {
#pragma unroll
                for (int b = 0; b < 2; b++) count_sym((w >> (16 * b)) & 0xFFFFu);
            }
        };
        constexpr int NLOAD = 4;  // 16-byte loads in flight per thread
        uint64_t i = threadIdx.x;
        for (; i + (NLOAD - 1) * blockDim.x < nvec; i += NLOAD * blockDim.x) {
            uint4 q[NLOAD];
#pragma unroll
            for (int k = 0; k < NLOAD; k++) q[k] = ldg_stream_v4(v + i + k * blockDim.x);
#pragma unroll
            for (int k = 0; k < NLOAD; k++) {
                count_word(q[k].x);
                count_word(q[k].y);
                count_word(q[k].z);
                count_word(q[k].w);
            }
        }
        for (; i < nvec; i += blockDim.x) {
            const uint4 q0 = ldg_stream_v4(v + i);
            count_word(q0.x);
            count_word(q0.y);
            count_word(q0.z);
            count_word(q0.w);
        }
        done = nvec * PER;
    }
    for (uint64_t i = done + threadIdx.x; i < cnt; i += blockDim.x) count_sym(p[i]);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < K; b += blockDim.x) {
        uint32_t t = 0;
        for (uint32_t r = 0; r < REP; r++) t += s_hist[(b << rep_log2) + ((r + b) & (REP - 1))];
        counts[chunk * K + b] = t;
    }
    if (!FULL && threadIdx.x < REP && s_hist[(K << rep_log2) + threadIdx.x]) atomicOr(bad, 1u);
}

// ============================================================ K2 model build
// counts -> {cum, c} tables.  One block (256 threads) per model:
//   c   = normalise(counts)            (identity unless a u64 sum > 2^32-1)
//   cum = exclusive prefix sum of c    (warp-shuffle scan + carry across tiles)
//   total = sum c
// Normalisation (build-defined: total_freq is u32 in the reference, src/pmodel.rs:10, and it has
// no rule for larger sums): when sum > 2^32-1 the counts are rescaled to total = 2^31 exactly,
//   c'_i = 0 if c_i == 0 else max(1, floor(c_i * 2^31 / sum)),
// and the difference 2^31 - sum(c') (|.| <= K) is added to the symbol with the largest count
// (lowest index on ties).  A power-of-two total keeps range/total a shift.  Identical in the oracle.
__device__ __forceinline__ uint32_t scale_count(unsigned long long c, unsigned long long sum) {
    if (c == 0) return 0u;
    const unsigned __int128 q = ((unsigned __int128)c << 31) / sum;
    const uint32_t v = (uint32_t)q;  // < 2^31
    return v ? v : 1u;
}

template <typename CNT>
__global__ void __launch_bounds__(256) counts_to_tables_kernel(const CNT* __restrict__ counts,
                                                               uint32_t K, uint2* tabs,
                                                               uint32_t* totals) {
    __shared__ unsigned long long s_warp[8];
    __shared__ unsigned long long s_wmax[8];
    __shared__ uint32_t s_widx[8];
    __shared__ unsigned long long s_sum;
    __shared__ unsigned long long s_carry;
    __shared__ uint32_t s_imax;
    __shared__ long long s_delta;
    const uint64_t model = blockIdx.x;
    const CNT* cnt = counts + model * K;
    uint2* tab = tabs + model * K;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    // pass 1: sum (u64) and the position of the largest count (lowest index on ties)
    unsigned long long part = 0, vmax = 0;
    uint32_t imax = 0xFFFFFFFFu;
    for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) {
        const unsigned long long v = (unsigned long long)cnt[i];
        part += v;
        if (v > vmax || (v == vmax && i < imax)) {
            vmax = v;
            imax = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        part += __shfl_down_sync(0xffffffffu, part, o);
        const unsigned long long ov = __shfl_down_sync(0xffffffffu, vmax, o);
        const uint32_t oi = __shfl_down_sync(0xffffffffu, imax, o);
        if (ov > vmax || (ov == vmax && oi < imax)) {
            vmax = ov;
            imax = oi;
        }
    }
    if (lane == 0) {
        s_warp[warp] = part;
        s_wmax[warp] = vmax;
        s_widx[warp] = imax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long sum = 0, bv = 0;
        uint32_t bi = 0xFFFFFFFFu;
        for (uint32_t w = 0; w < nwarps; w++) {
            sum += s_warp[w];
            if (s_wmax[w] > bv || (s_wmax[w] == bv && s_widx[w] < bi)) {
                bv = s_wmax[w];
                bi = s_widx[w];
            }
        }
        s_sum = sum;
        s_imax = bi;
        s_carry = 0;
        s_delta = 0;
    }
    __syncthreads();
    const unsigned long long sum = s_sum;
    const bool rescale = sum > 0xFFFFFFFFull;

    // pass 2 (only when rescaling): sum of the scaled counts -> correction for the largest symbol
    if (rescale) {
        unsigned long long sp = 0;
        for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) sp += scale_count((unsigned long long)cnt[i], sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sp += __shfl_down_sync(0xffffffffu, sp, o);
        __syncthreads();
        if (lane == 0) s_warp[warp] = sp;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (uint32_t w = 0; w < nwarps; w++) t += s_warp[w];
            s_delta = (long long)(1ull << 31) - (long long)t;
        }
        __syncthreads();
    }
    const long long delta = s_delta;
    const uint32_t fix = s_imax;

    // pass 3: tiles of blockDim.x symbols, exclusive scan with running carry
    for (uint32_t base = 0; base < K; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        unsigned long long raw = i < K ? (unsigned long long)cnt[i] : 0ull;
        uint32_t c = (uint32_t)raw;
        if (rescale) {
            c = scale_count(raw, sum);
            if (i == fix) c = (uint32_t)((long long)c + delta);
        }
        uint32_t incl = warp_incl_scan(c, lane);
        __syncthreads();
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
        for (uint32_t w = 0; w < warp; w++) woff += (uint32_t)s_warp[w];
        uint32_t carry = (uint32_t)s_carry;
        uint32_t cum = carry + woff + incl - c;  // u32 wrapping like release-mode Rust
        if (i < K) tab[i] = make_uint2(cum, c);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = (unsigned long long)(cum + c);
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[model] = (uint32_t)s_carry;
}

// Validate a table, derive its header (reciprocal, flags, staging bound) and,
// for the shared model, the decode LUT.  One block per model.
__global__ void __launch_bounds__(256) finalize_models_kernel(const uint2* __restrict__ tabs,
                                                              const uint32_t* __restrict__ totals,
                                                              uint32_t K, ModelHdr* hdrs,
                                                              LutEntry* lut, uint32_t lut_cap,
                                                              uint2* tab_cs, uint4* lut_cs,
                                                              uint32_t* summary /*[0]=min c, [1]=bad bits, [3]=max total*/) {
    __shared__ uint32_t s_flags_bad;  // bit0 inconsistent, bit1 irregular, bit2 cum>total, bit3 c==total
    __shared__ uint32_t s_minc;
    const uint64_t model = blockIdx.x;
    const uint2* tab = tabs + model * K;
    const uint32_t total = totals[model];
    if (threadIdx.x == 0) {
        s_flags_bad = 0;
        s_minc = 0xFFFFFFFFu;
    }
    __syncthreads();
    uint32_t bad = 0, minc = 0xFFFFFFFFu;
    for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) {
        uint2 e = tab[i];
        unsigned long long end = (unsigned long long)e.x + e.y;
        if (end > total) bad |= 1u;
        if (e.x > total) bad |= 4u;
        if (i + 1 < K) {
            if ((unsigned long long)tab[i + 1].x != end) bad |= 2u;
        } else if (end != total) {
            // unused code space above the last symbol (legal for a PModel): the row kernels derive
            // c[K-1] as total - cum[K-1], so such a table is not a prefix-sum table either
            bad |= 2u;
        }
        if (e.y && e.y < minc) minc = e.y;
        if (e.y == total) bad |= 8u;
        // shared model: reciprocal constants of the divide-free general-total step (rcb_core.cuh)
        if (tab_cs && total && (total & (total - 1u))) {  // only general totals use them
            const uint64_t r = e.y < total ? recip_of_freq(e.y, total) : ~0ull;
            tab_cs[i] = make_uint2(lo32(r), hi32(r));
        }
    }
    if (bad) atomicOr(&s_flags_bad, bad);
    atomicMin(&s_minc, minc);
    __syncthreads();
    if (threadIdx.x == 0) {
        ModelHdr h;
        h.div.total = total;
        h.div.shift = 0;
        h.div.magic = 0;
        uint32_t flags = 0;
        if (total && (total & (total - 1)) == 0) {
            flags |= MODEL_POW2;
            h.div.shift = 31u - (uint32_t)__clz((int)total);
        } else if (total) {
            h.div.magic = 0xFFFFFFFFFFFFFFFFull / total;  // == floor(2^64/total): total does not divide 2^64
        }
        if (!(s_flags_bad & 1u)) flags |= MODEL_CONSISTENT;
        if (!(s_flags_bad & 3u)) flags |= MODEL_REGULAR;
        if (s_flags_bad & 8u) flags |= MODEL_FULLC;
        h.flags = flags;
        h.min_c = s_minc;
        h.K = K;
        // LUT geometry (used only when lut != nullptr): nb = min(total, cap) buckets of equal, in general
        // fractional, width total / nb; bucket b starts at floor(b * total / nb).  (For a power-of-two total
        // >= cap this is the shift geometry b << log2(total / cap).)
        h.nb = total < lut_cap ? total : lut_cap;
        h.wshift = 0u;
        h.lut_scale = (float)h.nb;
        h.pad0 = h.pad1 = 0;
        hdrs[model] = h;
        if (s_flags_bad & 4u) atomicOr(&summary[1], 1u);
        if (total == 0) atomicOr(&summary[1], 2u);
        if (s_flags_bad & 1u) atomicOr(&summary[1], 4u);   // some model inconsistent
        if (s_flags_bad & 3u) atomicOr(&summary[1], 8u);   // some model irregular
        atomicMin(&summary[0], s_minc);
        atomicMax(&summary[3], total);
    }
    if (lut == nullptr || total == 0) return;
    __syncthreads();
    // LUT (shared model only; blockIdx.x == 0).  Built only for REGULAR tables.
    const ModelHdr& H = hdrs[model];
    if (!(H.flags & MODEL_REGULAR)) return;
    for (uint32_t b = threadIdx.x; b < H.nb; b += blockDim.x) {
        // the point this entry is built for: 1/8 bucket below the bucket start (estimate margin)
        unsigned long long v0 = (unsigned long long)b * total / H.nb;  // < total
        const unsigned long long margin = (total / H.nb) >> 3;
        v0 = v0 > margin ? v0 - margin : 0ull;
        // A = number of i in [1,K-1] with cum[i] <= v0  (the reference's search result)
        uint32_t left = 0, right = K - 1;
        while (left < right) {
            uint32_t mid = (left + right) >> 1;
            if ((unsigned long long)tab[mid + 1].x <= v0) left = mid + 1; else right = mid;
        }
        uint32_t A = left;
        uint2 ea = tab[A];
        uint32_t cumA = ea.x, cumB = ea.x + ea.y;
        uint32_t B = A + 1;
        while (B < K && tab[B].y == 0) B++;
        uint32_t cumC = cumB;
        if (B < K) cumC = cumB + tab[B].y;
        LutEntry e;
        e.cumA = cumA;
        e.cumB = cumB;
        e.cumC = cumC;
        e.syms = A | ((B & 0xFFFFu) << 16);
        lut[b] = e;
        if (lut_cs && (total & (total - 1u))) {  // cs of the entry's two candidates (general totals)
            const uint32_t cA = cumB - cumA, cB = cumC - cumB;
            const uint64_t ra = cA < total ? recip_of_freq(cA, total) : ~0ull;
            const uint64_t rb = cB < total ? recip_of_freq(cB, total) : ~0ull;
            lut_cs[b] = make_uint4(lo32(ra), hi32(ra), lo32(rb), hi32(rb));
        }
    }
}

// ============================================================ K4 compaction
// Exclusive scan of the chunk lengths into u64 offsets, plus an error summary:
// summary[0] = number of chunks with status != 0, [1] = first such chunk,
// [2] = its status, [3] = max length (needed pitch after ST_OUT_CAPACITY).
__global__ void __launch_bounds__(1024) scan_lengths_kernel(const uint32_t* __restrict__ lens,
                                                            const uint32_t* __restrict__ status,
                                                            uint64_t n, uint64_t* offsets,
                                                            unsigned long long* summary) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_bad[32];
    __shared__ unsigned long long s_first[32];
    __shared__ uint32_t s_maxlen[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint64_t lo = (uint64_t)threadIdx.x * per < n ? (uint64_t)threadIdx.x * per : n;
    const uint64_t hi = lo + per < n ? lo + per : n;
    unsigned long long sum = 0, bad = 0, firstbad = ~0ull;
    uint32_t maxlen = 0;
    for (uint64_t i = lo; i < hi; i++) {
        uint32_t l = lens[i];
        sum += l;
        maxlen = l > maxlen ? l : maxlen;
        if (status[i]) {
            bad++;
            if (firstbad == ~0ull) firstbad = i;
        }
    }
    unsigned long long incl = warp_incl_scan64(sum, lane);
    unsigned long long wbad = bad, wfirst = firstbad;
    uint32_t wmax = maxlen;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wbad += __shfl_down_sync(0xffffffffu, wbad, o);
        unsigned long long f = __shfl_down_sync(0xffffffffu, wfirst, o);
        wfirst = f < wfirst ? f : wfirst;
        uint32_t m = __shfl_down_sync(0xffffffffu, wmax, o);
        wmax = m > wmax ? m : wmax;
    }
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) {
        s_bad[warp] = wbad;
        s_first[warp] = wfirst;
        s_maxlen[warp] = wmax;
    }
    __syncthreads();
    unsigned long long woff = 0;
    for (uint32_t w = 0; w < warp; w++) woff += s_warp[w];
    unsigned long long run = woff + incl - sum;
    for (uint64_t i = lo; i < hi; i++) {
        offsets[i] = run;
        run += lens[i];
    }
    if (threadIdx.x == blockDim.x - 1) offsets[n] = woff + incl;
    if (threadIdx.x == 0) {
        unsigned long long tb = 0, tf = ~0ull;
        uint32_t tm = 0;
        for (uint32_t w = 0; w < (blockDim.x >> 5); w++) {
            tb += s_bad[w];
            tf = s_first[w] < tf ? s_first[w] : tf;
            tm = s_maxlen[w] > tm ? s_maxlen[w] : tm;
        }
        summary[0] = tb;
        summary[1] = tf;
        summary[2] = tb ? status[tf] : 0;
        summary[3] = tm;
    }
}

// 16-byte word at byte offset `off` (any alignment) of a 16-byte aligned row.
__device__ __forceinline__ uint4 load_unaligned16(const uint8_t* row, uint64_t off) {
    const uint4* p = reinterpret_cast<const uint4*>(row + (off & ~15ull));
    uint4 a = p[0];
    uint32_t sh = (uint32_t)(off & 15u);
    if (sh == 0) return a;
    uint4 b = p[1];
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t q = sh >> 2, r = (sh & 3u) * 8u;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t x0 = 0, x1 = 0;
        // q is uniform per block: the compiler resolves this into 4 predicated cases
#pragma unroll
        for (int qq = 0; qq < 4; qq++)
            if ((uint32_t)qq == q) {
                x0 = w[qq + k];
                x1 = w[qq + k + 1];
            }
        o[k] = __funnelshift_r(x0, x1, r);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// One block per chunk: staging row -> stream[offsets[chunk] ...).  The
// destination is written with aligned 16-byte stores; the (constant per chunk)
// source misalignment is absorbed by funnel shifts.
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t* __restrict__ staging, uint64_t pitch,
                                                     const uint32_t* __restrict__ lens,
                                                     const uint64_t* __restrict__ offsets,
                                                     uint8_t* out, uint64_t out_cap) {
    const uint64_t chunk = blockIdx.x;
    const uint8_t* row = staging + chunk * pitch;
    const uint64_t off = offsets[chunk];
    uint64_t len = lens[chunk];
    if (len > pitch) len = pitch;           // ST_OUT_CAPACITY rows hold only `pitch` bytes
    if (off + len > out_cap) return;        // reported through h_out_bytes > out_cap
    uint8_t* dst = out + off;
    uint64_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
    if (head > len) head = len;
    for (uint64_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = row[i];
    const uint64_t nvec = (len - head) >> 4;
    uint4* dv = reinterpret_cast<uint4*>(dst + head);
    for (uint64_t v = threadIdx.x; v < nvec; v += blockDim.x)
        dv[v] = load_unaligned16(row, head + (v << 4));
    const uint64_t tail0 = head + (nvec << 4);
    for (uint64_t i = tail0 + threadIdx.x; i < len; i += blockDim.x) dst[i] = row[i];
}

// Reduce a status array into the same 4-word summary scan_lengths_kernel writes.
__global__ void __launch_bounds__(1024) status_summary_kernel(const uint32_t* __restrict__ status,
                                                              uint64_t n, unsigned long long* summary) {
    __shared__ unsigned long long s_bad[32];
    __shared__ unsigned long long s_first[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long bad = 0, firstbad = ~0ull;
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (status[i]) {
            bad++;
            if (i < firstbad) firstbad = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bad += __shfl_down_sync(0xffffffffu, bad, o);
        unsigned long long f = __shfl_down_sync(0xffffffffu, firstbad, o);
        firstbad = f < firstbad ? f : firstbad;
    }
    if (lane == 0) {
        s_bad[warp] = bad;
        s_first[warp] = firstbad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tb = 0, tf = ~0ull;
        for (uint32_t w = 0; w < (blockDim.x >> 5); w++) {
            tb += s_bad[w];
            tf = s_first[w] < tf ? s_first[w] : tf;
        }
        summary[0] = tb;
        summary[1] = tf;
        summary[2] = tb ? status[tf] : 0;
        summary[3] = 0;
    }
}

// ============================================================ synthetic data
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename SYM>
__global__ void __launch_bounds__(256) generate_kernel(SYM* out, uint64_t first, uint64_t n, uint32_t K,
                                                       uint64_t seed, const uint32_t* __restrict__ thr,
                                                       uint32_t n_tables, uint64_t chunk_syms) {
    extern __shared__ uint32_t s_thr[];  // [n_tables][K-1]
    const uint32_t nthr = K - 1;
    for (uint32_t i = threadIdx.x; i < n_tables * nthr; i += blockDim.x) s_thr[i] = thr[i];
    __syncthreads();
    constexpr uint32_t PER = 16 / sizeof(SYM);
    const uint64_t ngroups = (n + PER - 1) / PER;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups;
         g += (uint64_t)gridDim.x * blockDim.x) {
        __align__(16) SYM v[PER];
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) {
            uint64_t j = first + g * PER + k;
            uint32_t r = (uint32_t)(mix64(seed + j * 0x9E3779B97F4A7C15ull) >> 32);
            uint32_t t = n_tables > 1 ? (uint32_t)((j / chunk_syms) % n_tables) : 0u;
            const uint32_t* th = s_thr + t * nthr;
            uint32_t lo = 0, hi = nthr;
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (th[mid] <= r) lo = mid + 1; else hi = mid;
            }
            v[k] = (SYM)lo;
        }
        uint64_t base = g * PER;
        if (base + PER <= n && (reinterpret_cast<uintptr_t>(out + base) & 15u) == 0) {
            *reinterpret_cast<uint4*>(out + base) = *reinterpret_cast<uint4*>(v);
        } else {
            for (uint32_t k = 0; k < PER && base + k < n; k++) out[base + k] = v[k];
        }
    }
}

}  // namespace rcb
