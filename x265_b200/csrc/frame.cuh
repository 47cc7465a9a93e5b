// x265_b200/csrc/frame.cuh -- frame-level fused kernels of the CTU-analysis pipeline:
//   k_build_me_jobs : PU geometry table + predictor field -> motionEstimate jobs
//                     (Search::setSearchRange search.cpp:2724-2768, CUData::clipMv cudata.cpp:1915-1928)
//   k_cu_residual   : per TU, fully on-chip: best-ref select -> luma MC (predict.cpp:245-266) -> sub_ps
//                     -> dct -> quant -> dequant_normal -> idct / DC shortcut -> add_ps -> sse_pp
//                     (search.cpp:3178 estimateResidualQT; quant.cpp:397-470, :543-605 non-RDOQ path)
//   k_intra_search  : per CU (8/16/32): neighbours from the source plane, 1:2:1 filter, 35 predictions
//                     evaluated in registers, SA8D by an 8-lane shuffle Hadamard (search.cpp:1358-1444)
#pragma once
#include "common.cuh"
#include "interp.cuh"
#include "intra.cuh"

#include "geometry.h"        // PuDesc / CuDesc / TuDesc and the host-side table builder

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__global__ void __launch_bounds__(256) k_build_me_jobs(const PuDesc* __restrict__ pus, int n, const int16_t* __restrict__ field,
                                                       int fw, int fh, int width, int height, int stride,
                                                       int method, int subme, int merange, x265cu_me_job* __restrict__ jobs)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PuDesc d = pus[i];
    const int y = d.offset / stride, x = d.offset - y * stride;
    const int bx = x >> 4, by = y >> 4, r = d.ref;
    const int16_t* f  = field + ((size_t)(r * fh + by) * fw + bx) * 2;
    const int16_t* fr = field + ((size_t)(r * fh + by) * fw + clampi(bx + 1, 0, fw - 1)) * 2;
    const int16_t* fb = field + ((size_t)(r * fh + clampi(by + 1, 0, fh - 1)) * fw + bx) * 2;
    x265cu_me_job j;
    j.offset = d.offset; j.ref = d.ref; j.pw = d.pw; j.ph = d.ph;
    j.qmvp[0] = f[0]; j.qmvp[1] = f[1];
    j.mvc[0] = fr[0]; j.mvc[1] = fr[1]; j.mvc[2] = fb[0]; j.mvc[3] = fb[1];
    j.mvc[4] = j.mvc[5] = j.mvc[6] = j.mvc[7] = 0;
    j.numCand = 2; j.method = (int8_t)method; j.subme = (int8_t)subme; j.merange = (int8_t)merange;
    const int dist = merange << 2;
    const int xmax = (width + 8 - d.cuX - 1) << 2, xmin = -((64 + 8 + d.cuX - 1) << 2);
    const int ymax = (height + 8 - d.cuY - 1) << 2, ymin = -((64 + 8 + d.cuY - 1) << 2);
    int mnx = clampi(f[0] - dist, xmin, xmax) >> 2, mxx = clampi(f[0] + dist, xmin, xmax) >> 2;
    int mny = clampi(f[1] - dist, ymin, ymax) >> 2, mxy = clampi(f[1] + dist, ymin, ymax) >> 2;
    if (mxy < mny) mxy = mny;
    j.mvmin[0] = (int16_t)mnx; j.mvmin[1] = (int16_t)mny; j.mvmax[0] = (int16_t)mxx; j.mvmax[1] = (int16_t)mxy;
    jobs[i] = j;
}

// {cost, mvx, mvy, 0} int32 x4 -> {cost int32, mvx int16, mvy int16}: what the host mode decision reads back
__global__ void k_pack_me(const int32_t* __restrict__ me_out, int n, int2* __restrict__ packed)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 v = ((const int4*)me_out)[i];
    packed[i] = make_int2(v.x, (int)(((uint32_t)(uint16_t)(int16_t)v.y) | ((uint32_t)(uint16_t)(int16_t)v.z << 16)));
}

// ---------------------------------------------------------------------------------------------
template <typename P>
__global__ void __launch_bounds__(256) k_cu_residual(const P* __restrict__ fenc, const P* const* __restrict__ refs, int stride,
                                                     const CuDesc* __restrict__ cus, const TuDesc* __restrict__ tus, int ntu,
                                                     const int32_t* __restrict__ cu_jobs, int numRefs, const int32_t* __restrict__ me_out,
                                                     int qp, int16_t* __restrict__ coef, P* const* __restrict__ recon,
                                                     unsigned long long* __restrict__ cu_sse, uint32_t* __restrict__ cu_numsig, int32_t* __restrict__ cu_ref)
{
    constexpr int DEPTH = PixTraits<P>::depth;
    constexpr int maxv = PixTraits<P>::maxv;
    __shared__ int16_t s_win[39 * 40];     // (T+7)^2 source window
    __shared__ int16_t s_mid[39 * 32];     // hps(rowExt) intermediate
    __shared__ int16_t s_pred[32 * 32];
    __shared__ int16_t s_a[32 * 34];       // rows padded to 34 in the forward passes (bank-conflict-free column reads)
    __shared__ int16_t s_b[32 * 34];
    __shared__ int8_t  s_m[32 * 32];
    __shared__ int s_red[8];
    __shared__ int s_lvl0;                 // quantised DC level of the TU
    __shared__ unsigned long long s_red64[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int t = blockIdx.x; t < ntu; t += gridDim.x)
    {
        const TuDesc tu = tus[t];
        const CuDesc cu = cus[tu.cu];
        const int S = cu.size, T = S > 32 ? 32 : S, NN = T * T;
        const int lg = 31 - __clz(T);
        // best reference by ME cost, ties -> lowest index
        int best = 0, bcost = 0x7fffffff;
        for (int r = 0; r < numRefs; r++)
        {
            int c = me_out[(size_t)cu_jobs[tu.cu * numRefs + r] * 4];
            if (c < bcost) { bcost = c; best = r; }
        }
        const int32_t* mo = me_out + (size_t)cu_jobs[tu.cu * numRefs + best] * 4;
        const int qx = mo[1], qy = mo[2], xf = qx & 3, yf = qy & 3;
        const int px = cu.x + tu.tx, py = cu.y + tu.ty;
        const P* src = refs[best] + (size_t)py * stride + px + (qx >> 2) + (ptrdiff_t)(qy >> 2) * stride;
        __syncthreads();
        for (int i = tid; i < NN; i += blockDim.x) s_m[i] = d_dct[lg - 2][i];
        // ---- luma MC into s_pred ----
        if (!(xf | yf))
        {
            for (int i = tid; i < NN; i += blockDim.x) { int y = i >> lg, x = i & (T - 1); s_pred[i] = (int16_t)src[(ptrdiff_t)y * stride + x]; }
        }
        else
        {
            const int hl = xf ? 3 : 0, vt = yf ? 3 : 0;
            const int ww = T + (xf ? 7 : 0), wh = T + (yf ? 7 : 0);
            for (int i = tid; i < ww * wh; i += blockDim.x)
            {
                int r = i / ww, c = i - r * ww;
                s_win[r * 40 + c] = (int16_t)src[(ptrdiff_t)(r - vt) * stride + (c - hl)];
            }
            __syncthreads();
            const int16_t* cx = c_lumaFilter[xf];
            const int16_t* cy = c_lumaFilter[yf];
            if (!yf)
            {
                for (int i = tid; i < NN; i += blockDim.x)
                {
                    int y = i >> lg, x = i & (T - 1), sum = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) sum += (int)s_win[y * 40 + x + k] * cx[k];
                    s_pred[i] = (int16_t)interp_finish<DEPTH>(sum, 0);
                }
            }
            else if (!xf)
            {
                for (int i = tid; i < NN; i += blockDim.x)
                {
                    int y = i >> lg, x = i & (T - 1), sum = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) sum += (int)s_win[(y + k) * 40 + x] * cy[k];
                    s_pred[i] = (int16_t)interp_finish<DEPTH>(sum, 0);
                }
            }
            else
            {
                for (int i = tid; i < T * wh; i += blockDim.x)
                {
                    int r = i >> lg, x = i & (T - 1), sum = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) sum += (int)s_win[r * 40 + x + k] * cx[k];
                    s_mid[r * 32 + x] = (int16_t)interp_finish<DEPTH>(sum, 1);
                }
                __syncthreads();
                for (int i = tid; i < NN; i += blockDim.x)
                {
                    int y = i >> lg, x = i & (T - 1), sum = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) sum += (int)s_mid[(y + k) * 32 + x] * cy[k];
                    s_pred[i] = (int16_t)interp_finish<DEPTH>(sum, 2);
                }
            }
        }
        __syncthreads();
        // ---- residual ----
        const P* fe = fenc + (size_t)py * stride + px;
        for (int i = tid; i < NN; i += blockDim.x) { int y = i >> lg, x = i & (T - 1); s_a[y * 34 + x] = (int16_t)((int)fe[(size_t)y * stride + x] - (int)s_pred[i]); }
        __syncthreads();
        // ---- forward DCT: two passes (dct.cpp:83-240, 442-525) ----
        {
            const int sh1 = lg - 1 + (DEPTH - 8), sh2 = lg + 6;
            for (int i = tid; i < NN; i += blockDim.x)
            {
                int k = i >> lg, jj = i & (T - 1), acc = 0;
                for (int q = 0; q < T; q++) acc += (int)s_m[k * T + q] * s_a[jj * 34 + q];
                s_b[k * 34 + jj] = (int16_t)((acc + (1 << (sh1 - 1))) >> sh1);
            }
            __syncthreads();
            for (int i = tid; i < NN; i += blockDim.x)
            {
                int k = i >> lg, jj = i & (T - 1), acc = 0;
                for (int q = 0; q < T; q++) acc += (int)s_m[k * T + q] * s_b[jj * 34 + q];
                s_a[i] = (int16_t)((acc + (1 << (sh2 - 1))) >> sh2);           // coefficients back in linear k*T + j order
            }
            __syncthreads();
        }
        // ---- quant (dct.cpp:664-686; quant.cpp:411,465-466) + dequant (dct.cpp:612-634; quant.cpp:556,567) ----
        const int per = qp / 6, rem = qp - per * 6;
        const int quantScale = rem == 0 ? 26214 : rem == 1 ? 23302 : rem == 2 ? 20560 : rem == 3 ? 18396 : rem == 4 ? 16384 : 14564;
        const int invScale = (rem == 0 ? 40 : rem == 1 ? 45 : rem == 2 ? 51 : rem == 3 ? 57 : rem == 4 ? 64 : 72) << per;
        const int transformShift = 15 - DEPTH - lg;
        const int qbits = 14 + per + transformShift;
        const int add = 85 << (qbits - 9);
        const int dqshift = 20 - 14 - transformShift;
        int16_t* qout = coef + cu.coef_off + (size_t)((tu.ty / T) * (S / T) + (tu.tx / T)) * NN;
        int cnt = 0;
        for (int i = tid; i < NN; i += blockDim.x)
        {
            int c = s_a[i];
            int tmp = abs(c) * quantScale;
            int lvl = (tmp + add) >> qbits;
            cnt += (lvl != 0);
            if (c < 0) lvl = -lvl;
            int q = clip16(lvl);
            qout[i] = (int16_t)q;
            if (i == 0) s_lvl0 = q;
            s_b[i] = (int16_t)clip16((q * invScale + (1 << (dqshift - 1))) >> dqshift);      // dequantised
        }
        cnt = warp_sum(cnt);
        if (lane == 0) s_red[warp] = cnt;
        __syncthreads();
        int numSig = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) numSig += s_red[w];
        const int q0 = (int)s_b[0];        // dequantised DC (valid after the barrier)
        // the reference tests coeff[0] != 0 on the QUANTISED level (quant.cpp:588)
        const bool dcOnlyRef = (numSig == 1) && (s_lvl0 != 0);
        __syncthreads();
        // ---- inverse transform into s_a (residual') ----
        if (numSig == 0)
        {
            for (int i = tid; i < NN; i += blockDim.x) s_a[i] = 0;
        }
        else if (dcOnlyRef)
        {
            const int shift_2nd = 12 - (DEPTH - 8) - 3;
            const int dc = (((q0 + 1) >> 1) * 8 + (1 << (shift_2nd - 1))) >> shift_2nd;
            for (int i = tid; i < NN; i += blockDim.x) s_a[i] = (int16_t)dc;
        }
        else
        {
            const int sh1 = 7, sh2 = 12 - (DEPTH - 8);
            for (int i = tid; i < NN; i += blockDim.x)
            {
                int jj = i >> lg, i2 = i & (T - 1), acc = 0;
                for (int q = 0; q < T; q++) acc += (int)s_m[q * T + i2] * s_b[q * T + jj];
                s_a[i] = (int16_t)clip16((acc + (1 << (sh1 - 1))) >> sh1);
            }
            __syncthreads();
            for (int i = tid; i < NN; i += blockDim.x)
            {
                int jj = i >> lg, i2 = i & (T - 1), acc = 0;
                for (int q = 0; q < T; q++) acc += (int)s_m[q * T + i2] * s_a[q * T + jj];
                s_b[i] = (int16_t)clip16((acc + (1 << (sh2 - 1))) >> sh2);
            }
            __syncthreads();
            for (int i = tid; i < NN; i += blockDim.x) s_a[i] = s_b[i];
        }
        __syncthreads();
        // ---- reconstruction + distortion ----
        const int dIdx = S == 64 ? 0 : (S == 32 ? 1 : (S == 16 ? 2 : 3));
        P* rc = recon[dIdx] + (size_t)py * stride + px;
        unsigned long long sse = 0;
        for (int i = tid; i < NN; i += blockDim.x)
        {
            int y = i >> lg, x = i & (T - 1);
            int v = numSig ? clip3i(0, maxv, (int)s_pred[i] + (int)s_a[i]) : (int)s_pred[i];
            rc[(size_t)y * stride + x] = (P)v;
            int d = (int)fe[(size_t)y * stride + x] - v;
            sse += (unsigned)(d * d);
        }
        sse = warp_sum64(sse);
        if (lane == 0) s_red64[warp] = sse;
        __syncthreads();
        if (tid == 0)
        {
            unsigned long long tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += s_red64[w];
            atomicAdd(&cu_sse[tu.cu], tot);
            atomicAdd(&cu_numsig[tu.cu], (uint32_t)numSig);
            cu_ref[tu.cu] = best;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// closed-form intra prediction of one pixel (intrapred.cpp:53-204); nb/filt in shared memory
__device__ __forceinline__ int intra_pixel(const int16_t* __restrict__ nbs, int N, int lg, int mode, int bFilter, int dc, int y, int x, int maxv)
{
    const int N2 = 2 * N;
    if (mode == 0)
        return ((N - 1 - x) * nbs[N2 + 1 + y] + (N - 1 - y) * nbs[1 + x] + (x + 1) * nbs[1 + N] + (y + 1) * nbs[N2 + 1 + N] + N) >> (lg + 1);
    if (mode == 1)
    {
        if (bFilter)
        {
            if (x == 0 && y == 0) return (nbs[1] + nbs[N2 + 1] + 2 * dc + 2) >> 2;
            if (y == 0) return (nbs[1 + x] + 3 * dc + 2) >> 2;
            if (x == 0) return (nbs[N2 + 1 + y] + 3 * dc + 2) >> 2;
        }
        return dc;
    }
    const bool hor = mode < 18;
    const int angOff = hor ? 10 - mode : mode - 26;
    const int angle = c_angle[8 + angOff];
    const int mainBase = hor ? N2 + 1 : 1, sideBase = hor ? 1 : N2 + 1;
    const int r = hor ? x : y, c = hor ? y : x;           // vertical-family frame coordinates
    if (angle == 0)
    {
        int v = nbs[mainBase + c];
        if (bFilter && c == 0) v = clip3i(0, maxv, (int)(int16_t)(nbs[mainBase] + ((nbs[sideBase + r] - nbs[0]) >> 1)));
        return v;
    }
    const int pos = (r + 1) * angle, o = pos >> 5, f = pos & 31;
    const int inv = angle < 0 ? c_invAngle[-angOff - 1] : 0;
    auto ref_at = [&](int idx) -> int {
        if (idx >= 0) return nbs[mainBase + idx];
        if (idx == -1) return nbs[0];
        const int i = -idx - 2;                                    // projected side neighbours (intrapred.cpp:150-160)
        return nbs[sideBase - 1 + ((128 + (i + 1) * inv) >> 8)];
    };
    const int a = ref_at(o + c);
    if (!f) return a;
    return ((32 - f) * a + f * ref_at(o + c + 1) + 16) >> 5;
}

// 8x8 Hadamard abs-sum over an 8-lane group: each lane holds one row of differences
__device__ __forceinline__ int had8_group(int d[8], int lane)
{
    had4(d[0], d[1], d[2], d[3]); had4(d[4], d[5], d[6], d[7]);
#pragma unroll
    for (int k = 0; k < 4; k++) { int p = d[k], q = d[k + 4]; d[k] = p + q; d[k + 4] = p - q; }
#pragma unroll
    for (int s = 1; s < 8; s <<= 1)
    {
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            int v = __shfl_xor_sync(0xffffffffu, d[k], s);
            d[k] = (lane & s) ? (v - d[k]) : (d[k] + v);
        }
    }
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc += abs(d[k]);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    return acc;          // raw 8x8 sum, identical in the 8 lanes of the group
}

template <typename P>
__global__ void __launch_bounds__(256) k_intra_search(const P* __restrict__ fenc, int stride, const CuDesc* __restrict__ cus, int ncu,
                                                      uint32_t* __restrict__ intra_cost)
{
    constexpr int maxv = PixTraits<P>::maxv;
    __shared__ int16_t s_nb[129], s_filt[129];
    __shared__ uint32_t s_cost[36];
    __shared__ int s_dc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int grp = lane >> 3, row = lane & 7;
    for (int c = blockIdx.x; c < ncu; c += gridDim.x)
    {
        const CuDesc cu = cus[c];
        const int S = cu.size;
        uint32_t* out = intra_cost + (size_t)c * 36;
        if (S == 64) { if (tid < 36) out[tid] = 0; continue; }
        const int lg = 31 - __clz(S), N2 = 2 * S;
        const P* o = fenc + (size_t)cu.y * stride + cu.x;
        __syncthreads();
        for (int i = tid; i < 4 * S + 1; i += blockDim.x)
        {
            int v;
            if (i == 0) v = o[-stride - 1];
            else if (i <= N2) v = o[-stride + (i - 1)];
            else v = o[(ptrdiff_t)(i - N2 - 1) * stride - 1];
            s_nb[i] = (int16_t)v;
        }
        __syncthreads();
        for (int i = tid; i < 4 * S + 1; i += blockDim.x) s_filt[i] = (int16_t)intra_filter_elem(s_nb, i, S);
        if (tid == 0)
        {
            int sum = S;
            for (int i = 0; i < S; i++) sum += s_nb[1 + i] + s_nb[N2 + 1 + i];
            s_dc = sum / N2;
        }
        __syncthreads();
        const int dc = s_dc;
        const int bFilter = S <= 16;
        if (S == 8)
        {   // 4 modes per warp pass: lane group g -> mode, lane row -> block row
            for (int m0 = warp * 4; m0 < 35; m0 += nwarps * 4)
            {
                const int mode = min(m0 + grp, 34);
                const int16_t* nbs = intra_use_filtered(mode, S) ? s_filt : s_nb;
                int d[8];
#pragma unroll
                for (int x = 0; x < 8; x++)
                    d[x] = (int)o[(size_t)row * stride + x] - intra_pixel(nbs, S, lg, mode, bFilter, dc, row, x, maxv);
                int raw = had8_group(d, lane);
                if (row == 0 && m0 + grp < 35) s_cost[m0 + grp] = (uint32_t)((raw + 2) >> 2);
            }
        }
        else
        {   // one mode per warp pass; 4 tiles (one 16x16) per pass: group g -> 8x8 tile of the 16x16
            for (int mode = warp; mode < 35; mode += nwarps)
            {
                const int16_t* nbs = intra_use_filtered(mode, S) ? s_filt : s_nb;
                int total = 0;
                for (int by = 0; by < S; by += 16)
                    for (int bx = 0; bx < S; bx += 16)
                    {
                        const int y = by + (grp >> 1) * 8 + row, x0 = bx + (grp & 1) * 8;
                        int d[8];
#pragma unroll
                        for (int x = 0; x < 8; x++)
                            d[x] = (int)o[(size_t)y * stride + x0 + x] - intra_pixel(nbs, S, lg, mode, bFilter, dc, y, x0 + x, maxv);
                        int raw = had8_group(d, lane);
                        raw += __shfl_xor_sync(0xffffffffu, raw, 8);
                        raw += __shfl_xor_sync(0xffffffffu, raw, 16);
                        total += (raw + 2) >> 2;                       // sa8d_16x16 rounds once per 16x16 (pixel.cpp:341-351)
                    }
                if (lane == 0) s_cost[mode] = (uint32_t)total;
            }
        }
        __syncthreads();
        if (tid < 35) out[tid] = s_cost[tid];
        if (tid == 0)
        {
            uint32_t bc = 0xffffffffu; int bm = 0;
            for (int m = 0; m < 35; m++) if (s_cost[m] < bc) { bc = s_cost[m]; bm = m; }
            out[35] = (uint32_t)bm;
        }
    }
}
