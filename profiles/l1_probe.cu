// profiles/l1_probe.cu -- micro-benchmark for the integer-search question: what does one warp-wide
// load REQUEST cost on B200 as a function of how many distinct 128-byte lines / 32-byte sectors its 32 lanes touch and of
// the access width?  Patterns are the ones the SAD core produces:
//   rowshift : 32 lanes in ONE picture row, 5 bytes apart (raster chunk: lane = candidate)              -> 2 lines
//   rows4    : 8 candidates x 4 lanes, the 4 lanes of a candidate in 4 consecutive rows (16x16 PU burst)  -> up to 32 lines
//   rows32   : every lane in its own row (worst case)                                                    -> 32 lines
// each with LDG.32 / LDG.64 / LDG.128 (aligned down), data L1-resident (64 KB footprint) or L2-resident (8 MB footprint).
// Reports SM cycles per request with 8 resident warps per SM issuing dependent-free loads.
// Build + run:  nvcc -arch=sm_100a -O3 -o /tmp/l1_probe profiles/l1_probe.cu && /tmp/l1_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int WIDTH>   // bytes per lane per load: 4, 8, 16
__global__ void __launch_bounds__(256) k_probe(const uint8_t* __restrict__ base, int pattern, int stride, int span, int iters, unsigned* sink, long long* cyc)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    size_t off;
    if (pattern == 0)      off = (size_t)(warp * 8) * stride + lane * 5;                       // one row, 5 B apart
    else if (pattern == 1) off = (size_t)(warp * 8 + (lane & 3)) * stride + (lane >> 2) * 5;   // 8 candidates x 4 rows
    else                   off = (size_t)lane * stride + warp * 16;                            // 32 rows
    off &= ~(size_t)(WIDTH - 1);
    unsigned acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++)
    {
        const size_t o = off + (size_t)((it * 16) % span) * stride;       // walk down the rows like a PU walk
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
            const uint8_t* p = base + o + (size_t)u * stride;
            if (WIDTH == 4)       acc += __ldg((const unsigned*)p);
            else if (WIDTH == 8)  { const uint2 v = __ldg((const uint2*)p); acc += v.x ^ v.y; }
            else                  { const uint4 v = __ldg((const uint4*)p); acc += v.x ^ v.y ^ v.z ^ v.w; }
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    const int stride = 4032, rows = 2320;
    uint8_t* d; unsigned* sink; long long* cyc;
    cudaMalloc(&d, (size_t)stride * rows + 4096); cudaMemset(d, 1, (size_t)stride * rows + 4096);
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const char* pn[3] = { "rowshift", "rows4", "rows32" };
    printf("pattern   width  footprint  cycles/request (one CTA of 8 warps per SM, %d SMs)\n", prop.multiProcessorCount);
    for (int pat = 0; pat < 3; pat++)
        for (int w = 4; w <= 16; w *= 2)
            for (int big = 0; big < 2; big++)
            {
                const int span = big ? 2000 : 16, iters = 4000;
                for (int rep = 0; rep < 2; rep++)
                {
                    if (w == 4)      k_probe<4><<<prop.multiProcessorCount, 256>>>(d, pat, stride, span, iters, sink, cyc);
                    else if (w == 8) k_probe<8><<<prop.multiProcessorCount, 256>>>(d, pat, stride, span, iters, sink, cyc);
                    else             k_probe<16><<<prop.multiProcessorCount, 256>>>(d, pat, stride, span, iters, sink, cyc);
                    cudaDeviceSynchronize();
                }
                long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
                // 8 warps per SM each issue iters * 8 requests; the SM serves them concurrently: cycles per request SM-wide
                printf("%-9s %5d  %-9s  %.2f\n", pn[pat], w, big ? "L2 (8 MB)" : "L1 (64 K)", (double)c / ((double)iters * 8 * 8));
            }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
