// x265_b200/csrc/transform.cuh -- transform / quant class.
// Semantics: /root/reference/source/common/dct.cpp:43-81 (DST4), :83-440 (partial butterflies),
// :442-610 (dct/idct drivers + shifts), :612-713 (dequant / quant / nquant).
// The butterflies are exact integer factorizations of the HEVC matrices, so the shared-memory
// matrix product below is bit-identical (|acc| <= 32*90*32768 < 2^31).
#pragma once
#include "common.cuh"
#include "transform_mma.cuh"
#include <mutex>

// host: regenerate the HEVC matrices from the 32-point basis and upload them
static const int16_t h_basis[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
                                     64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
static int h_basis_at(int a)
{
    a &= 127;
    if (a > 64) a = 128 - a;
    return a <= 32 ? h_basis[a] : -h_basis[64 - a];
}
// The host table is a function-local static built exactly once (C++11 thread-safe initialisation) and never written again;
// the upload into the device-global __constant__ copy happens once per DEVICE under a mutex (several host threads create
// contexts concurrently when the per-call table is plugged into the encoder's worker pool).
struct DctHostTable
{
    int8_t tab[4][32 * 32];
    DctHostTable()
    {
        memset(tab, 0, sizeof(tab));
        for (int l = 0; l < 4; l++)
        {
            int n = 4 << l, step = 32 / n;
            for (int k = 0; k < n; k++)
                for (int j = 0; j < n; j++)
                    tab[l][k * n + j] = (int8_t)h_basis_at(k * step * (2 * j + 1));
        }
    }
};
static int build_dct_tables(int device)
{
    static const DctHostTable host;
    static const int8_t dst4[16] = { 29, 55, 74, 84, 74, 74, 0, -74, 84, -29, -74, 55, 55, -84, 74, -29 };
    static std::mutex mtx;
    static bool uploaded[64] = { false };
    std::lock_guard<std::mutex> lock(mtx);
    if (device >= 0 && device < 64 && uploaded[device]) return 0;
    CU_CHECK(cudaMemcpyToSymbol(c_dct, host.tab, sizeof(host.tab)));
    CU_CHECK(cudaMemcpyToSymbol(d_dct, host.tab, sizeof(host.tab)));
    CU_CHECK(cudaMemcpyToSymbol(c_dst4, dst4, sizeof(dst4)));
    CU_CHECK(cudaDeviceSynchronize());
    if (device >= 0 && device < 64) uploaded[device] = true;
    return 0;
}

// One CTA (256 threads) handles TPB = max(1, 1024 / N^2) TUs per iteration.  Both passes are shared-memory
// matrix products; rows are padded to N+2 int16 so that the 32 lanes of a warp (one output column each)
// hit 32 different banks, and every thread produces 4 outputs per input element it reads.
// op: X265CU_DCT / IDCT / DST4 / IDST4
template <int DEPTH, int N>
__global__ void __launch_bounds__(256) k_transform(int op, const int16_t* __restrict__ src, int16_t* __restrict__ dst,
                                                   int stride, int64_t tu_pitch, int n)
{
    constexpr int NN = N * N, LG = (N == 4 ? 2 : N == 8 ? 3 : N == 16 ? 4 : 5);
    constexpr int TPB = (1024 / NN) < 1 ? 1 : (1024 / NN > 16 ? 16 : 1024 / NN);       // 32:1  16:4  8:16  4:16
    constexpr int PS = N + 2;                                                           // padded row stride
    constexpr int OPT = 4;                                                              // outputs per thread
    constexpr int TPT = NN / OPT;                                                       // threads per TU (N=4: 4)
    __shared__ int16_t s_in[TPB * N * PS];
    __shared__ int16_t s_mid[TPB * N * PS];
    __shared__ int8_t s_m[NN];
    const bool fwd = (op == X265CU_DCT || op == X265CU_DST4);
    const bool dstm = (op == X265CU_DST4 || op == X265CU_IDST4);
    for (int i = threadIdx.x; i < NN; i += blockDim.x) s_m[i] = dstm ? c_dst4[i] : d_dct[LG - 2][i];
    const int shift1 = fwd ? LG - 1 + (DEPTH - 8) : 7;
    const int shift2 = fwd ? LG + 6 : 12 - (DEPTH - 8);
    const int tl = threadIdx.x / TPT, tt = threadIdx.x % TPT;      // TU slot in the CTA, thread inside the TU
    const int col = tt & (N - 1), rb = tt >> LG;                   // output column, first output row (rows rb + (N/4)*m)
    constexpr int RSTEP = N / OPT;

    for (int base = blockIdx.x * TPB; base < n; base += gridDim.x * TPB)
    {
        const int cnt = min(TPB, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * NN; i += blockDim.x)
        {
            int t = i >> (2 * LG), e = i & (NN - 1), y = e >> LG, x = e & (N - 1);
            s_in[(t * N + y) * PS + x] = fwd ? src[(int64_t)(base + t) * tu_pitch + (int64_t)y * stride + x] : src[(int64_t)(base + t) * NN + e];
        }
        __syncthreads();
        const bool act = tl < cnt && threadIdx.x < TPB * TPT;
        const int16_t* in = s_in + tl * N * PS;
        int16_t* mid = s_mid + tl * N * PS;
        if (act)
        {
            int acc[OPT] = { 0, 0, 0, 0 };
            if (fwd)
            {   // mid[k][j] = sum_q M[k][q] * in[j][q];  j = col, k = rb + RSTEP*m
#pragma unroll 8
                for (int q = 0; q < N; q++)
                {
                    const int x = in[col * PS + q];
#pragma unroll
                    for (int m = 0; m < OPT; m++) acc[m] += (int)s_m[(rb + RSTEP * m) * N + q] * x;
                }
#pragma unroll
                for (int m = 0; m < OPT; m++) mid[(rb + RSTEP * m) * PS + col] = (int16_t)((acc[m] + (1 << (shift1 - 1))) >> shift1);
            }
            else
            {   // mid[j][i] = clip16(sum_q M[q][i] * in[q][j]);  j = col, i = rb + RSTEP*m
#pragma unroll 8
                for (int q = 0; q < N; q++)
                {
                    const int x = in[q * PS + col];
#pragma unroll
                    for (int m = 0; m < OPT; m++) acc[m] += (int)s_m[q * N + rb + RSTEP * m] * x;
                }
#pragma unroll
                for (int m = 0; m < OPT; m++) mid[col * PS + rb + RSTEP * m] = (int16_t)clip16((acc[m] + (1 << (shift1 - 1))) >> shift1);
            }
        }
        __syncthreads();
        if (act)
        {
            int acc[OPT] = { 0, 0, 0, 0 };
            if (fwd)
            {
#pragma unroll 8
                for (int q = 0; q < N; q++)
                {
                    const int x = mid[col * PS + q];
#pragma unroll
                    for (int m = 0; m < OPT; m++) acc[m] += (int)s_m[(rb + RSTEP * m) * N + q] * x;
                }
#pragma unroll
                for (int m = 0; m < OPT; m++)
                    dst[(int64_t)(base + tl) * NN + (rb + RSTEP * m) * N + col] = (int16_t)((acc[m] + (1 << (shift2 - 1))) >> shift2);
            }
            else
            {
#pragma unroll 8
                for (int q = 0; q < N; q++)
                {
                    const int x = mid[q * PS + col];
#pragma unroll
                    for (int m = 0; m < OPT; m++) acc[m] += (int)s_m[q * N + rb + RSTEP * m] * x;
                }
#pragma unroll
                for (int m = 0; m < OPT; m++)
                    dst[(int64_t)(base + tl) * tu_pitch + (int64_t)col * stride + rb + RSTEP * m] = (int16_t)clip16((acc[m] + (1 << (shift2 - 1))) >> shift2);
            }
        }
    }
}

// ---- 4x4 / 8x8, contiguous TUs: one THREAD per TU, everything in registers -------------------------------
// The even/odd symmetry of the DCT rows (M[k][N-1-q] = (-1)^k M[k][q]) halves the multiplies; the sums are the
// same integers as the matrix product (dct.cpp:83-440 are the same factorisation), the rounding shifts and
// the int16 truncation / clipping are the reference's (dct.cpp:442-610).  Matrix entries are compile-time
// indexed constant-bank operands.  A thread reads and writes its TU with 16-byte accesses.
template <int N>
__device__ __forceinline__ void dct_fwd1d(const int (&x)[N], int (&y)[N])
{
    constexpr int LG = N == 4 ? 2 : 3, H = N / 2;
    int sm[H], df[H];
#pragma unroll
    for (int q = 0; q < H; q++) { sm[q] = x[q] + x[N - 1 - q]; df[q] = x[q] - x[N - 1 - q]; }
#pragma unroll
    for (int k = 0; k < N; k++)
    {
        int acc = 0;
#pragma unroll
        for (int q = 0; q < H; q++) acc += (int)c_dct[LG - 2][k * N + q] * ((k & 1) ? df[q] : sm[q]);
        y[k] = acc;
    }
}
template <int N>
__device__ __forceinline__ void dct_inv1d(const int (&x)[N], int (&y)[N])
{
    constexpr int LG = N == 4 ? 2 : 3, H = N / 2;
#pragma unroll
    for (int i = 0; i < H; i++)
    {
        int e = 0, o = 0;
#pragma unroll
        for (int q = 0; q < N; q += 2) { e += (int)c_dct[LG - 2][q * N + i] * x[q]; o += (int)c_dct[LG - 2][(q + 1) * N + i] * x[q + 1]; }
        y[i] = e + o; y[N - 1 - i] = e - o;
    }
}

template <int DEPTH, int N, bool FWD>
__global__ void __launch_bounds__(128) k_transform_reg(const int16_t* __restrict__ src, int16_t* __restrict__ dst, int n)
{
    constexpr int NN = N * N, LG = N == 4 ? 2 : 3;
    constexpr int shift1 = FWD ? LG - 1 + (DEPTH - 8) : 7;
    constexpr int shift2 = FWD ? LG + 6 : 12 - (DEPTH - 8);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int a[N][N];
    {
        const uint4* sp = (const uint4*)(src + (size_t)t * NN);
#pragma unroll
        for (int i = 0; i < NN / 8; i++)
        {
            const uint4 v = __ldg(sp + i);
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                const int e = i * 8 + k * 2;
                a[e / N][e % N] = (int)(int16_t)(w[k] & 0xffffu);
                a[(e + 1) / N][(e + 1) % N] = (int)(int16_t)(w[k] >> 16);
            }
        }
    }
    int m[N][N], o[N][N];
    if (FWD)
    {
        // mid[k][j] = (sum_q M[k][q] in[j][q] + r) >> shift1, truncated to int16; dst[k][c] = (sum_q M[k][q] mid[c][q] + r) >> shift2
#pragma unroll
        for (int j = 0; j < N; j++)
        {
            int y[N];
            dct_fwd1d<N>(a[j], y);
#pragma unroll
            for (int k = 0; k < N; k++) m[k][j] = (int)(int16_t)((y[k] + (1 << (shift1 - 1))) >> shift1);
        }
#pragma unroll
        for (int c = 0; c < N; c++)
        {
            int y[N];
            dct_fwd1d<N>(m[c], y);
#pragma unroll
            for (int k = 0; k < N; k++) o[k][c] = (int)(int16_t)((y[k] + (1 << (shift2 - 1))) >> shift2);
        }
    }
    else
    {
        // mid[j][i] = clip16((sum_q M[q][i] in[q][j] + 64) >> 7); out[c][i] = clip16((sum_q M[q][i] mid[q][c] + r) >> shift2)
#pragma unroll
        for (int j = 0; j < N; j++)
        {
            int x[N], y[N];
#pragma unroll
            for (int q = 0; q < N; q++) x[q] = a[q][j];
            dct_inv1d<N>(x, y);
#pragma unroll
            for (int i = 0; i < N; i++) m[j][i] = clip16((y[i] + (1 << (shift1 - 1))) >> shift1);
        }
#pragma unroll
        for (int c = 0; c < N; c++)
        {
            int x[N], y[N];
#pragma unroll
            for (int q = 0; q < N; q++) x[q] = m[q][c];
            dct_inv1d<N>(x, y);
#pragma unroll
            for (int i = 0; i < N; i++) o[c][i] = clip16((y[i] + (1 << (shift2 - 1))) >> shift2);
        }
    }
    uint4* dp = (uint4*)(dst + (size_t)t * NN);
#pragma unroll
    for (int i = 0; i < NN / 8; i++)
    {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const int e = i * 8 + k * 2;
            w[k] = ((uint32_t)o[e / N][e % N] & 0xffffu) | ((uint32_t)o[(e + 1) / N][(e + 1) % N] << 16);
        }
        dp[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <int DEPTH>
static int launch_transform_reg(x265cu_ctx* ctx, int op, int N, const int16_t* src, int16_t* dst, int stride, int64_t tu_pitch, int n)
{
    if ((op != X265CU_DCT && op != X265CU_IDCT) || (N != 4 && N != 8) || stride != N || tu_pitch != (int64_t)N * N ||
        (((uintptr_t)src | (uintptr_t)dst) & 15))
        return 0;
    const int blocks = (n + 127) / 128;
    if (N == 4)
    {
        if (op == X265CU_DCT) k_transform_reg<DEPTH, 4, true><<<blocks, 128, 0, ctx->stream>>>(src, dst, n);
        else                  k_transform_reg<DEPTH, 4, false><<<blocks, 128, 0, ctx->stream>>>(src, dst, n);
    }
    else
    {
        if (op == X265CU_DCT) k_transform_reg<DEPTH, 8, true><<<blocks, 128, 0, ctx->stream>>>(src, dst, n);
        else                  k_transform_reg<DEPTH, 8, false><<<blocks, 128, 0, ctx->stream>>>(src, dst, n);
    }
    return 1;
}

template <int DEPTH>
static int launch_transform_d(x265cu_ctx* ctx, int op, int N, const int16_t* src, int16_t* dst, int stride, int64_t tu_pitch, int n)
{
    int tpb = 1024 / (N * N); if (tpb < 1) tpb = 1; if (tpb > 16) tpb = 16;
    int blocks = (n + tpb - 1) / tpb;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    switch (N)
    {
    case 4:  k_transform<DEPTH, 4><<<blocks, 256, 0, ctx->stream>>>(op, src, dst, stride, tu_pitch, n); break;
    case 8:  k_transform<DEPTH, 8><<<blocks, 256, 0, ctx->stream>>>(op, src, dst, stride, tu_pitch, n); break;
    case 16: k_transform<DEPTH, 16><<<blocks, 256, 0, ctx->stream>>>(op, src, dst, stride, tu_pitch, n); break;
    default: k_transform<DEPTH, 32><<<blocks, 256, 0, ctx->stream>>>(op, src, dst, stride, tu_pitch, n); break;
    }
    return 0;
}

static int launch_transform(x265cu_ctx* ctx, int depth, int op, int N, const int16_t* src, int16_t* dst, int stride, int64_t tu_pitch, int n)
{
    if (n <= 0) return 0;
    if (op == X265CU_DST4 || op == X265CU_IDST4) N = 4;
    if (N != 4 && N != 8 && N != 16 && N != 32) { x265cu_set_error("transform size", cudaErrorInvalidValue, __FILE__, __LINE__); return -1; }
    // 16x16 / 32x32 go to the tensor-core (IMMA) kernel when the operands are vector-load aligned
    int took = (depth == 8) ? launch_transform_mma<8>(ctx, op, N, src, dst, stride, tu_pitch, n)
                            : launch_transform_mma<10>(ctx, op, N, src, dst, stride, tu_pitch, n);
    if (!took) took = (depth == 8) ? launch_transform_reg<8>(ctx, op, N, src, dst, stride, tu_pitch, n)
                                   : launch_transform_reg<10>(ctx, op, N, src, dst, stride, tu_pitch, n);
    if (!took)
    {
        if (depth == 8) launch_transform_d<8>(ctx, op, N, src, dst, stride, tu_pitch, n);
        else            launch_transform_d<10>(ctx, op, N, src, dst, stride, tu_pitch, n);
    }
    CU_LAUNCH_CHECK(ctx);
    return 0;
}

// quant / nquant (dct.cpp:664-713): one warp per TU; numSig via ballot + popc.
__global__ void __launch_bounds__(256) k_quant(const int16_t* __restrict__ coef, const int32_t* __restrict__ qc, int32_t* __restrict__ deltaU,
                                               int16_t* __restrict__ qCoef, int qBits, int add, int numCoeff, int n, int nquant,
                                               uint32_t* __restrict__ numSig)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < n; t += gridDim.x * wpb)
    {
        const int64_t o = (int64_t)t * numCoeff;
        int cnt = 0;
        for (int i = lane; i < numCoeff; i += 32)      // numCoeff is a multiple of 16; tail lanes masked
        {
            int c = coef[o + i];
            int tmp = abs(c) * qc[i];
            int lvl = (tmp + add) >> qBits;
            if (deltaU) deltaU[o + i] = (tmp - (lvl << qBits)) >> (qBits - 8);
            cnt += (lvl != 0);
            if (c < 0) lvl = -lvl;
            int q = clip16(lvl);
            qCoef[o + i] = (int16_t)(nquant ? abs(q) : q);
        }
        cnt = warp_sum(cnt);
        if (lane == 0) numSig[t] = (uint32_t)cnt;
    }
}

__global__ void __launch_bounds__(256) k_dequant_normal(const int16_t* __restrict__ q, int16_t* __restrict__ coef, int64_t num, int scale, int shift)
{
    const int add = 1 << (shift - 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num; i += (int64_t)gridDim.x * blockDim.x)
        coef[i] = (int16_t)clip16((q[i] * scale + add) >> shift);
}

// dct.cpp:636-662
__global__ void __launch_bounds__(256) k_dequant_scaling(const int16_t* __restrict__ q, const int32_t* __restrict__ dq, int16_t* __restrict__ coef,
                                                         int numCoeff, int64_t total, int per, int shift)
{
    shift += 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    {
        int p = q[i] * dq[i % numCoeff];
        int v;
        if (shift > per) v = clip16((p + (1 << (shift - per - 1))) >> (shift - per));
        else             v = clip16((int)((unsigned)clip16(p) << (per - shift)));
        coef[i] = (int16_t)v;
    }
}

// denoiseDct (dct.cpp:744-755), one TU
__global__ void k_denoise(int16_t* __restrict__ dctCoef, uint32_t* __restrict__ resSum, const uint16_t* __restrict__ offset, int numCoeff)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < numCoeff; i += gridDim.x * blockDim.x)
    {
        int lvl = dctCoef[i];
        int sign = lvl >> 31;
        lvl = (lvl + sign) ^ sign;
        resSum[i] += (uint32_t)lvl;
        lvl -= offset[i];
        dctCoef[i] = (int16_t)(lvl < 0 ? 0 : (lvl ^ sign) - sign);
    }
}
