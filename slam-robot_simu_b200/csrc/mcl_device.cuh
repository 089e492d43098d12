// Device-side building blocks of the MCL engine (sm_100a).  Included by mcl_kernels.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcl.h"

namespace mcl {

constexpr double kPi = 3.141592653589793238462643383279502884;
constexpr double kTwoPi = 6.283185307179586476925286766559005768;
constexpr double kTwo52 = 4503599627370496.0;
constexpr double kInvTwo52 = 1.0 / 4503599627370496.0;

constexpr int kStepThreads = 256;  // predict+weight block
constexpr int kPPT = 4;            // particles per thread per iteration (one float4 per array)
constexpr int kStepTile = kStepThreads * kPPT;

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// partial-reduction record (doubles)
//  0 m   1 s1   2 s2   3 dx 4 dy 5 dth 6 sin 7 cos 8 xx 9 xy 10 yy 11 xth 12 yth 13 thth
// 14 argmax value (stored fp32 logw, exact in a double)   15 argmax index (exact in a double)
constexpr int kNumSums = 13;  // entries 1..13 are sums that rescale with exp(m_b - M)
constexpr int kPart = 16;

// Per-trial, per-step scalars written by k_prep and read by every other kernel of the step.
struct PrepOut {
    double mom[13];        // landmark/observation moments for MCL_LIK_MOMENT
    double c_shift;        // log-likelihood of the reference pose, subtracted from every particle
    double ref[3];         // reference pose: moment sums are accumulated about it
    double log_norm_prev;  // normaliser of the stored log-weights
    double ofs;            // comb offset (already multiplied by 1/N)
    double lik_const;      // -L log(2 pi s0 s1)
    double base_lw;        // -log(n_global): normalised log-weight right after resampling
    int32_t do_resample;
    int32_t L;
};

// ---------------------------------------------------------------------------------------------
// mylib/limit.py:19-26 in fp64: r = |a|; while r > pi: r -= 2 pi; sign restored for a < 0.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double limit_angle(double a) {
    double r = fabs(a);
    if (r > 64.0) {  // outside the filter's regime: avoid a long (or, for inf, endless) subtraction loop
        if (!(r < 1.0e300)) return NAN;
        r = fmod(r, kTwoPi);
    }
    while (r > kPi) r = __dsub_rn(r, kTwoPi);
    return (a < 0.0) ? -r : r;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter-based: the noise of particle i at step k does not
// depend on launch geometry or on how the filter is sharded.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// two standard normals from two 32-bit words (Box-Muller, fp32; |n| <= 5.9)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = (static_cast<float>(a >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
    const float u2 = (static_cast<float>(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.28318530717958647692f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

__device__ __forceinline__ double philox_uniform01(uint4 r) {
    return (static_cast<double>(r.x) * 4294967296.0 + static_cast<double>(r.y)) * (1.0 / 18446744073709551616.0);
}

// ---------------------------------------------------------------------------------------------
// Fixed-point weight: rint(exp(logw - log_norm) * 2^52).  One definition shared by the scan, the
// shard-total and the export kernel so that all three see identical integers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long quantize_weight(float logw, double log_norm) {
    const double w = exp(__dsub_rn(static_cast<double>(logw), log_norm));
    return __double2ull_rn(__dmul_rn(w, kTwo52));
}

// u_i = fl(fl(i * (1/N)) + ofs): numpy rounds twice (np.arange then "+ ofs",
// particle_filter.py:213-215); the explicit _rn intrinsics stop nvcc contracting them into one FMA.
__device__ __forceinline__ double comb_value(long long i, double delta, double ofs) {
    return __dadd_rn(__dmul_rn(static_cast<double>(i), delta), ofs);
}

// Number of comb teeth u_i (0 <= i < n) with u_i <= K * 2^-52.  The walk of particle_filter.py:219
// advances while u_i > pw_cum[idx]; with integer cumulative weights K_j the ancestor of slot i is
// the first j with K_j * 2^-52 >= u_i, so source j owns slots [count_le(K_{j-1}), count_le(K_j)).
__device__ __forceinline__ long long count_le(unsigned long long K, double delta, double ofs, long long n,
                                              double n_as_double) {
    const double c = __dmul_rn(static_cast<double>(K), kInvTwo52);  // exact for K < 2^53
    double g = floor((c - ofs) * n_as_double);
    g = fmin(fmax(g, -1.0), n_as_double - 1.0);
    long long i = static_cast<long long>(g);
    while (i + 1 < n && comb_value(i + 1, delta, ofs) <= c) ++i;
    while (i >= 0 && comb_value(i, delta, ofs) > c) --i;
    return i + 1;
}

// ---------------------------------------------------------------------------------------------
// block-wide sum of NV doubles (256 threads); result valid in thread 0 .. (all threads after sync
// read smem[0..NV)).  smem must hold (blockDim/32) * NV doubles.
// ---------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) smem[warp * NV + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < nwarp; ++w) s += smem[w * NV + threadIdx.x];
        smem[threadIdx.x] = s;  // safe: each thread only overwrites its own column of row 0 after reading it
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = smem[k];
    __syncthreads();
}

__device__ __forceinline__ double block_max(double v, double *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double r = smem[0];
    for (int w = 1; w < nwarp; ++w) r = fmax(r, smem[w]);
    __syncthreads();
    return r;
}

}  // namespace mcl
