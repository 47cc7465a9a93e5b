// x265_b200/csrc/transform.cuh -- transform / quant class.
// Semantics: /root/reference/source/common/dct.cpp:43-81 (DST4), :83-440 (partial butterflies),
// :442-610 (dct/idct drivers + shifts), :612-713 (dequant / quant / nquant).
// The butterflies are exact integer factorizations of the HEVC matrices, so the shared-memory
// matrix product below is bit-identical (|acc| <= 32*90*32768 < 2^31).
#pragma once
#include "common.cuh"

// host: regenerate the HEVC matrices from the 32-point basis and upload them
static const int16_t h_basis[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
                                     64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
static int h_basis_at(int a)
{
    a &= 127;
    if (a > 64) a = 128 - a;
    return a <= 32 ? h_basis[a] : -h_basis[64 - a];
}
static int build_dct_tables()
{
    static int8_t tab[4][32 * 32];
    memset(tab, 0, sizeof(tab));
    for (int l = 0; l < 4; l++)
    {
        int n = 4 << l, step = 32 / n;
        for (int k = 0; k < n; k++)
            for (int j = 0; j < n; j++)
                tab[l][k * n + j] = (int8_t)h_basis_at(k * step * (2 * j + 1));
    }
    static const int8_t dst4[16] = { 29, 55, 74, 84, 74, 74, 0, -74, 84, -29, -74, 55, 55, -84, 74, -29 };
    CU_CHECK(cudaMemcpyToSymbol(c_dct, tab, sizeof(tab)));
    CU_CHECK(cudaMemcpyToSymbol(c_dst4, dst4, sizeof(dst4)));
    return 0;
}

// One CTA handles `tpb` TUs.  smem: in tile, mid tile (int16), matrix copy (int8).
// op: X265CU_DCT / IDCT / DST4 / IDST4
template <int DEPTH>
__global__ void __launch_bounds__(256) k_transform(int op, int N, const int16_t* __restrict__ src, int16_t* __restrict__ dst,
                                                   int stride, int64_t tu_pitch, int n, int tpb)
{
    extern __shared__ int16_t sm[];
    const int NN = N * N;
    int16_t* s_in = sm;                      // tpb * NN
    int16_t* s_mid = sm + tpb * NN;          // tpb * NN
    __shared__ int8_t s_m[32 * 32];
    const int lg = 31 - __clz(N);
    const bool fwd = (op == X265CU_DCT || op == X265CU_DST4);
    const bool dstm = (op == X265CU_DST4 || op == X265CU_IDST4);
    for (int i = threadIdx.x; i < NN; i += blockDim.x) s_m[i] = dstm ? c_dst4[i] : c_dct[lg - 2][i];
    const int shift1 = fwd ? lg - 1 + (DEPTH - 8) : 7;
    const int shift2 = fwd ? lg + 6 : 12 - (DEPTH - 8);

    for (int base = blockIdx.x * tpb; base < n; base += gridDim.x * tpb)
    {
        const int cnt = min(tpb, n - base);
        __syncthreads();
        // load: forward reads the strided residual, inverse reads contiguous coefficients
        for (int i = threadIdx.x; i < cnt * NN; i += blockDim.x)
        {
            int t = i / NN, e = i - t * NN;
            if (fwd) { int y = e >> lg, x = e & (N - 1); s_in[i] = src[(int64_t)(base + t) * tu_pitch + (int64_t)y * stride + x]; }
            else     s_in[i] = src[(int64_t)(base + t) * NN + e];
        }
        __syncthreads();
        // pass 1
        for (int i = threadIdx.x; i < cnt * NN; i += blockDim.x)
        {
            int t = i / NN, e = i - t * NN;
            const int16_t* in = s_in + t * NN;
            int acc = 0;
            if (fwd)
            {   // out[k*N + j] = sum_i M[k][i] * in[j*N + i]
                int k = e >> lg, jj = e & (N - 1);
                for (int q = 0; q < N; q++) acc += (int)s_m[k * N + q] * in[jj * N + q];
                s_mid[i] = (int16_t)((acc + (1 << (shift1 - 1))) >> shift1);
            }
            else
            {   // out[j*N + i2] = clip16(sum_k M[k][i2] * in[k*N + j])
                int jj = e >> lg, i2 = e & (N - 1);
                for (int q = 0; q < N; q++) acc += (int)s_m[q * N + i2] * in[q * N + jj];
                s_mid[i] = (int16_t)clip16((acc + (1 << (shift1 - 1))) >> shift1);
            }
        }
        __syncthreads();
        // pass 2
        for (int i = threadIdx.x; i < cnt * NN; i += blockDim.x)
        {
            int t = i / NN, e = i - t * NN;
            const int16_t* in = s_mid + t * NN;
            int acc = 0;
            if (fwd)
            {
                int k = e >> lg, jj = e & (N - 1);
                for (int q = 0; q < N; q++) acc += (int)s_m[k * N + q] * in[jj * N + q];
                dst[(int64_t)(base + t) * NN + e] = (int16_t)((acc + (1 << (shift2 - 1))) >> shift2);
            }
            else
            {
                int jj = e >> lg, i2 = e & (N - 1);
                for (int q = 0; q < N; q++) acc += (int)s_m[q * N + i2] * in[q * N + jj];
                dst[(int64_t)(base + t) * tu_pitch + (int64_t)jj * stride + i2] = (int16_t)clip16((acc + (1 << (shift2 - 1))) >> shift2);
            }
        }
    }
}

static int launch_transform(x265cu_ctx* ctx, int depth, int op, int N, const int16_t* src, int16_t* dst, int stride, int64_t tu_pitch, int n)
{
    if (n <= 0) return 0;
    if (op == X265CU_DST4 || op == X265CU_IDST4) N = 4;
    int tpb = 1024 / (N * N); if (tpb < 1) tpb = 1; if (tpb > 16) tpb = 16;
    int blocks = (n + tpb - 1) / tpb;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    size_t smem = (size_t)2 * tpb * N * N * sizeof(int16_t);
    if (depth == 8) k_transform<8><<<blocks, 256, smem, ctx->stream>>>(op, N, src, dst, stride, tu_pitch, n, tpb);
    else            k_transform<10><<<blocks, 256, smem, ctx->stream>>>(op, N, src, dst, stride, tu_pitch, n, tpb);
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
