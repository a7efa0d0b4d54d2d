// lat.cu -- dependent-issue latency of the integer ops on the coder's critical
// path, measured on one warp of one SM (the regime the 1 GiB / 64 KiB-chunk
// workload runs in: <= 1 warp per scheduler).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat tools/microbench/lat.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define N 4096

#define CHAIN(NAME, INIT, ...)                                                   \
    __global__ void NAME(uint32_t* out, long long* cyc, uint32_t seed) {         \
        uint32_t a = seed + threadIdx.x, b = seed * 3 + 1, c = seed ^ 0x55;      \
        uint64_t x = ((uint64_t)a << 32) | b, y = ((uint64_t)c << 20) | 12345;   \
        float f = (float)seed + 1.5f;                                            \
        INIT;                                                                    \
        long long t0 = clock64();                                                \
        _Pragma("unroll 16") for (int i = 0; i < N; i++) { __VA_ARGS__; }               \
        long long t1 = clock64();                                                \
        out[threadIdx.x] = a + b + c + (uint32_t)x + (uint32_t)(x >> 32) + (uint32_t)y + (uint32_t)f; \
        if (threadIdx.x == 0) *cyc = t1 - t0;                                    \
    }

extern __shared__ uint32_t smem[];

CHAIN(k_iadd, , a = a + b + i)
CHAIN(k_lop3, , a = (a ^ b) & (c | i))
CHAIN(k_shf, , a = __funnelshift_l(a, b, a & 7))
CHAIN(k_imad, , a = a * b + c)
CHAIN(k_imadwide, , x = (uint64_t)(uint32_t)x * (uint64_t)b + y)
CHAIN(k_mul64x32, , x = x * (uint64_t)b + y)
CHAIN(k_shl64, , x = (x << (b & 24)) | 1)
CHAIN(k_shr64, , x = (x >> (b & 31)) | 0x8000000000000000ull)
CHAIN(k_add64, , x = x + y + i)
CHAIN(k_flo, , a = (uint32_t)__clz((int)a) + b)
CHAIN(k_bfind, , { uint32_t r; asm volatile("bfind.u32 %0, %1;" : "=r"(r) : "r"(a)); a = r | b; })
CHAIN(k_setp_sel, , a = (a < b ? c : a) + 1)
CHAIN(k_popc, , a = __popc(a) + b)
CHAIN(k_prmt, , a = __byte_perm(a, b, 0x0123) + 1)
CHAIN(k_i2f_f2i, , a = (uint32_t)((float)a * 0.5f) + b)
CHAIN(k_i2f, , { f = (float)a; a = __float_as_uint(f) | 1; })
CHAIN(k_rcp, , { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f)); f = r + 1.0f; })
CHAIN(k_fmul, , f = f * 1.0001f)
CHAIN(k_f2i, , { a = (uint32_t)f; f = __uint_as_float(a | 0x3f800000); })
CHAIN(k_umul64hi, , x = __umul64hi(x, y) | 0x8000000000000000ull)
CHAIN(k_lds, { for (int j = threadIdx.x; j < 1024; j += 32) smem[j] = (j * 7 + 1) & 1023; __syncwarp(); a &= 1023; },
      a = smem[a])
CHAIN(k_lds128, { for (int j = threadIdx.x; j < 4096; j += 32) smem[j] = (j * 7 + 1) & 1023; __syncwarp(); a &= 1023; },
      { uint4 v = reinterpret_cast<uint4*>(smem)[a]; a = (v.x ^ v.w) & 1023; })
// whole fast-path recurrence of the pow2 encoder (fused shift), no emission
CHAIN(k_chain_fused, uint64_t lo = x; uint64_t rpt = y | (1ull << 33);,
      {
          uint32_t cum = b & 0xFFFF, cc = (c & 0xFFFF) | 1;
          uint64_t nlo = lo + rpt * cum;
          uint64_t up = lo + rpt * (cum + cc);
          uint64_t rgp = rpt * cc;
          uint32_t xh = (uint32_t)(nlo >> 32) ^ (uint32_t)(up >> 32);
          uint32_t fl;
          asm volatile("bfind.u32 %0, %1;" : "=r"(fl) : "r"(xh));
          uint32_t sh = ~fl & 24, k = (fl & 24) | 6;
          rpt = (rgp >> k) | (1ull << 33);
          lo = nlo << sh;
          x = lo;
      })

#define RUN(NAME, SMEM)                                                         \
    do {                                                                        \
        NAME<<<1, 32, SMEM>>>(d_out, d_cyc, 12345);                             \
        cudaDeviceSynchronize();                                                \
        NAME<<<1, 32, SMEM>>>(d_out, d_cyc, 12345);                             \
        cudaDeviceSynchronize();                                                \
        long long c;                                                            \
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);                       \
        printf("%-16s %7.2f cycles/iter\n", #NAME, (double)c / N);              \
    } while (0)

int main() {
    uint32_t* d_out;
    long long* d_cyc;
    cudaMalloc(&d_out, 4096);
    cudaMalloc(&d_cyc, 8);
    RUN(k_iadd, 0);
    RUN(k_lop3, 0);
    RUN(k_shf, 0);
    RUN(k_imad, 0);
    RUN(k_imadwide, 0);
    RUN(k_mul64x32, 0);
    RUN(k_shl64, 0);
    RUN(k_shr64, 0);
    RUN(k_add64, 0);
    RUN(k_flo, 0);
    RUN(k_bfind, 0);
    RUN(k_setp_sel, 0);
    RUN(k_popc, 0);
    RUN(k_prmt, 0);
    RUN(k_i2f_f2i, 0);
    RUN(k_i2f, 0);
    RUN(k_rcp, 0);
    RUN(k_fmul, 0);
    RUN(k_f2i, 0);
    RUN(k_umul64hi, 0);
    RUN(k_lds, 4096);
    RUN(k_lds128, 16384);
    RUN(k_chain_fused, 0);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
