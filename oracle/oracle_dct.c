/* oracle/oracle_dct.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Restates /root/reference/source/common/dct.cpp:43-742.  The reference evaluates the transforms
 * with partial butterflies (dct.cpp:83-440); integer arithmetic is exact and never overflows int32
 * (|acc| <= 32*90*32768 < 2^31), so the plain matrix product used here yields identical values.
 * The matrices are regenerated from the 32-point HEVC basis (pinned against constants.cpp:270-344
 * in tests/test_oracle_vs_ref.py).
 */
#include "oracle.h"
#include <stdlib.h>

/* magnitude of the 32-point basis at angle index a*pi/64, a = 0..32 (HEVC spec 8.6.4.2) */
static const int16_t k_basis[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
                                     64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
static int16_t g_mat[4][32 * 32];
static int g_mat_ready = 0;

static int basis_at(int a) /* cos-like lookup for angle index a (period 128) */
{
    a &= 127;
    if (a > 64) a = 128 - a;
    return a <= 32 ? k_basis[a] : -k_basis[64 - a];
}

static void build_matrices(void)
{
    for (int l = 0; l < 4; l++)
    {
        int n = 4 << l, step = 32 / n;         /* row k of the n-point matrix = row k*step of the 32-point */
        for (int k = 0; k < n; k++)
            for (int j = 0; j < n; j++)
                g_mat[l][k * n + j] = (int16_t)basis_at(k * step * (2 * j + 1));
    }
    g_mat_ready = 1;
}

const int16_t* orc_dct_matrix(int n)
{
    if (!g_mat_ready) build_matrices();
    switch (n) { case 4: return g_mat[0]; case 8: return g_mat[1]; case 16: return g_mat[2]; case 32: return g_mat[3]; }
    return NULL;
}

static inline int clip16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

/* one forward pass (dct.cpp:83-240,418-440): out[k*n + j] = (sum_i M[k][i]*in[j*n + i] + add) >> shift */
static void fwd_pass(const int16_t* M, const int16_t* in, int16_t* out, int n, int shift)
{
    int add = 1 << (shift - 1);
    for (int j = 0; j < n; j++)
        for (int k = 0; k < n; k++)
        {
            int acc = 0;
            for (int i = 0; i < n; i++) acc += (int)M[k * n + i] * in[j * n + i];
            out[k * n + j] = (int16_t)((acc + add) >> shift);     /* truncating cast, no clip (dct.cpp:113) */
        }
}

/* one inverse pass (dct.cpp:242-416): out[j*n + i] = clip16((sum_k M[k][i]*in[k*n + j] + add) >> shift) */
static void inv_pass(const int16_t* M, const int16_t* in, int16_t* out, int n, int shift)
{
    int add = 1 << (shift - 1);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++)
        {
            int acc = 0;
            for (int k = 0; k < n; k++) acc += (int)M[k * n + i] * in[k * n + j];
            out[j * n + i] = (int16_t)clip16((acc + add) >> shift);
        }
}

/* dct.cpp:459-525: shifts log2N-1+(depth-8), log2N+6 */
void orc_dct(const int16_t* src, int16_t* dst, intptr_t srcStride, int n)
{
    int16_t blk[32 * 32], tmp[32 * 32];
    const int16_t* M = orc_dct_matrix(n);
    int lg = ilog2(n);
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) blk[y * n + x] = src[y * srcStride + x];
    fwd_pass(M, blk, tmp, n, lg - 1 + (ORC_DEPTH - 8));
    fwd_pass(M, tmp, dst, n, lg + 6);
}

/* dct.cpp:544-610: shifts 7, 12-(depth-8) */
void orc_idct(const int16_t* src, int16_t* dst, intptr_t dstStride, int n)
{
    int16_t blk[32 * 32], tmp[32 * 32];
    const int16_t* M = orc_dct_matrix(n);
    inv_pass(M, src, tmp, n, 7);
    inv_pass(M, tmp, blk, n, 12 - (ORC_DEPTH - 8));
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) dst[y * dstStride + x] = blk[y * n + x];
}

/* 4x4 DST-VII matrix rows (HEVC spec 8.6.4.2; dct.cpp:43-81 is its fast form) */
static const int16_t k_dst4[4][4] = { { 29, 55, 74, 84 }, { 74, 74, 0, -74 }, { 84, -29, -74, 55 }, { 55, -84, 74, -29 } };

void orc_dst4(const int16_t* src, int16_t* dst, intptr_t srcStride)
{
    int16_t blk[16], tmp[16];
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) blk[y * 4 + x] = src[y * srcStride + x];
    fwd_pass(&k_dst4[0][0], blk, tmp, 4, 1 + (ORC_DEPTH - 8));
    fwd_pass(&k_dst4[0][0], tmp, dst, 4, 8);
}

void orc_idst4(const int16_t* src, int16_t* dst, intptr_t dstStride)
{
    int16_t blk[16], tmp[16];
    inv_pass(&k_dst4[0][0], src, tmp, 4, 7);
    inv_pass(&k_dst4[0][0], tmp, blk, 4, 12 - (ORC_DEPTH - 8));
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) dst[y * dstStride + x] = blk[y * 4 + x];
}

/* dct.cpp:664-686 */
uint32_t orc_quant(const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef, int qBits, int add, int numCoeff)
{
    uint32_t nsig = 0;
    for (int i = 0; i < numCoeff; i++)
    {
        int c = coef[i];
        int neg = c < 0;
        int t = abs(c) * quantCoeff[i];
        int lvl = (t + add) >> qBits;
        deltaU[i] = (t - (lvl << qBits)) >> (qBits - 8);
        nsig += (lvl != 0);
        if (neg) lvl = -lvl;
        qCoef[i] = (int16_t)clip16(lvl);
    }
    return nsig;
}

/* dct.cpp:688-713 */
uint32_t orc_nquant(const int16_t* coef, const int32_t* quantCoeff, int16_t* qCoef, int qBits, int add, int numCoeff)
{
    uint32_t nsig = 0;
    for (int i = 0; i < numCoeff; i++)
    {
        int c = coef[i];
        int neg = c < 0;
        int t = abs(c) * quantCoeff[i];
        int lvl = (t + add) >> qBits;
        nsig += (lvl != 0);
        if (neg) lvl = -lvl;
        qCoef[i] = (int16_t)abs(clip16(lvl));
    }
    return nsig;
}

/* dct.cpp:612-634 */
void orc_dequant_normal(const int16_t* q, int16_t* coef, int num, int scale, int shift)
{
    int add = 1 << (shift - 1);
    for (int i = 0; i < num; i++)
        coef[i] = (int16_t)clip16((q[i] * scale + add) >> shift);
}

/* dct.cpp:636-662 */
void orc_dequant_scaling(const int16_t* q, const int32_t* dq, int16_t* coef, int num, int per, int shift)
{
    shift += 4;
    if (shift > per)
    {
        int add = 1 << (shift - per - 1);
        for (int i = 0; i < num; i++)
            coef[i] = (int16_t)clip16((q[i] * dq[i] + add) >> (shift - per));
    }
    else
    {
        for (int i = 0; i < num; i++)
        {
            int v = clip16(q[i] * dq[i]);
            coef[i] = (int16_t)clip16((int)((unsigned)v << (per - shift)));
        }
    }
}

/* dct.cpp:744-755 */
void orc_denoise_dct(int16_t* dctCoef, uint32_t* resSum, const uint16_t* offset, int numCoeff)
{
    for (int i = 0; i < numCoeff; i++)
    {
        int lvl = dctCoef[i];
        int sign = lvl >> 31;
        lvl = (lvl + sign) ^ sign;
        resSum[i] += (uint32_t)lvl;
        lvl -= offset[i];
        dctCoef[i] = (int16_t)(lvl < 0 ? 0 : (lvl ^ sign) - sign);
    }
}
