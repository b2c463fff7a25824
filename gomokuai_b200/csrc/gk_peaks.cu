// gk_peaks.cu -- microbenchmark behind the integer-issue roofline ("measure the real INT32 IADD3/LOP3
// rate and clock under load ... and use that as denominator").  Three streams of independent register-only
// instructions, 16 chains per thread so that no dependency or latency limits the rate:
//   mode 0  LOP3 only            -> what the ALU pipe alone sustains (LOP3 / SHF / IADD3 / PRMT / ISETP share it)
//   mode 1  IMAD only            -> what the FMA pipe alone sustains (IMAD, IMAD.SHL, IMAD.MOV, IMAD.IADD)
//   mode 2  LOP3 and IMAD, 1 : 1 -> the issue ceiling of an integer kernel that balances the two pipes
// The result is warp instructions per second over the whole GPU (counted analytically; loop overhead is 3 in 259).
#include <cuda_runtime.h>

#include <cstdint>

#include "gk_kernels.h"

namespace gk {

namespace {

constexpr int kChains = 16, kUnroll = 16;        // 256 measured instructions per loop iteration

template <int kMode>
__global__ void __launch_bounds__(1024, 1) issue_peak_kernel(uint32_t* out, int iters, uint32_t k) {
    uint32_t a[kChains];
#pragma unroll
    for (int j = 0; j < kChains; ++j) a[j] = threadIdx.x * 2654435761u + j;
    const uint32_t k2 = k * 40503u + 1u;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
            for (int j = 0; j < kChains; ++j) {
                const bool alu = kMode == 0 || (kMode == 2 && (j & 1) == 0);
                // inline PTX so that every step is exactly one SASS instruction (checked with cuobjdump: 256 per iteration)
                if (alu) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(a[j]) : "r"(k), "r"(k2));    // one LOP3.LUT
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(k), "r"(k2));            // one IMAD
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < kChains; ++j) x ^= a[j];
    if (x == 0x12345678u) out[threadIdx.x] = x;                      // keeps the chains alive; practically never taken
}

}  // namespace

cudaError_t measure_issue_peak(int mode, int sm_count, int iters, double* warp_inst_per_s, cudaStream_t stream) {
    uint32_t* d = nullptr;
    cudaError_t err = cudaMalloc(&d, 1024 * sizeof(uint32_t));
    if (err != cudaSuccess) return err;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto launch = [&](int n) {
        if (mode == 0) issue_peak_kernel<0><<<sm_count, 1024, 0, stream>>>(d, n, 0x01000193u);
        else if (mode == 1) issue_peak_kernel<1><<<sm_count, 1024, 0, stream>>>(d, n, 0x01000193u);
        else issue_peak_kernel<2><<<sm_count, 1024, 0, stream>>>(d, n, 0x01000193u);
    };
    launch(iters / 8 + 1);                                           // warm-up (clocks, instruction cache)
    cudaEventRecord(e0, stream);
    launch(iters);
    cudaEventRecord(e1, stream);
    err = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err == cudaSuccess && ms > 0.f)
        *warp_inst_per_s = double(sm_count) * 32.0 /* warps per CTA */ * double(iters) * kChains * kUnroll / (double(ms) * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return err;
}

}  // namespace gk
