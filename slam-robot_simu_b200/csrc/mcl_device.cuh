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
constexpr float kLog2e = 1.4426950408889634f;

constexpr int kStepThreads = 256;  // predict+weight block
#ifndef MCL_STEP_CTAS_PER_SM
#define MCL_STEP_CTAS_PER_SM 3
#endif
constexpr int kStepCtasPerSm = MCL_STEP_CTAS_PER_SM;  // resident CTAs per SM k_step is compiled for
constexpr int kPPT = 4;            // particles per thread per iteration (one float4 per array)
constexpr int kStepTile = kStepThreads * kPPT;
constexpr int kStepUnit = 128;     // granularity of the static particle partition over blocks
constexpr int kFoldEvery = 16;     // fp32 pair kernels: landmarks per fp32 partial sum before it is folded into fp64
constexpr int kFlushEvery = 8;     // fp32 group sums are folded into the fp64 sums every 8 iterations

// The scan is bound by the latency of a tile's dependent steps, not by issue or bytes: what
// raises its throughput is more tiles in flight per SM.  128 threads x 16 sources keep the tile at 2048 particles
// and let EIGHT tiles reside on an SM (64 registers x 128 threads each) instead of four.
constexpr int kScanThreads = 128;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr int kScanWin = 32 * kScanItems;  // slots one warp fills per window of its private shared-memory strip

// partial-reduction record (doubles)
//  0 m   1 s1   2 s2   3 dx 4 dy 5 dth 6 sin 7 cos 8 xx 9 xy 10 yy 11 xth 12 yth 13 thth
// 14 argmax value (stored fp32 logw, exact in a double)   15 argmax index (exact in a double)
constexpr int kNumSums = 13;  // entries 1..13 are sums that rescale with exp(m_b - M)
constexpr int kPart = 16;

// Per-trial, per-step scalars written by k_prep and read by every other kernel of the step.
struct PrepOut {
    double mom[13];        // landmark/observation moments for the moment likelihood
    double c_shift;        // log-likelihood of the reference pose, subtracted from every particle
    double ref[3];         // reference pose: moment sums are accumulated about it
    double max_prev;       // max of the stored log-weights (they are kept relative to it)
    double ofs;            // comb offset (already multiplied by 1/N)
    double lik_const;      // -L log(2 pi s0 s1)
    float wq_m;            // weight map of the stored log-weights: w * 2^52 = ex2((lw - wq_m) log2e) * wq_scale
    float wq_scale;
    int32_t do_resample;
    int32_t L;
    double abs_shift;      // log(pw * bn of the reference, un-normalised) - stored log-weight, for THIS step's weights
    int32_t uniform;       // the stored weights underflowed in the reference's arithmetic: read them as all-equal
    int32_t pad;
};

// particle_filter.py:234-237: pw_update / sum is NaN when every product pw * bn underflowed to 0 in fp64, and NaN
// weights are replaced by 1/NP.  A product vanishes below half the smallest denormal, 2^-1075.
constexpr double kLogUnderflow = -745.1332191019412;
constexpr double kNoAbsShift = 1.0e300;  // no likelihood step defines the absolute scale (set_particles, RB models)
__host__ __device__ inline bool weights_underflowed(double max_logw, double sum_w, double abs_shift) {
    // (written so that NaN anywhere counts as underflow: the reference's isnan() test catches exactly that)
    return !(sum_w > 0.0 && sum_w < 1.0e300 && max_logw + abs_shift >= kLogUnderflow);
}

// ---------------------------------------------------------------------------------------------
// The transcendental part of mcl_stats from its raw sums: log-normaliser, weight of the MAP particle,
// circular mean and covariance about the mean.  Host side of mcl_get_stats, and k_cov.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline double limit_angle_hd(double a) {
    double r = fabs(a);
    if (!(r < 1.0e300)) return a;
    if (r > 64.0) r = fmod(r, kTwoPi);
    while (r > kPi) r -= kTwoPi;
    return (a < 0.0) ? -r : r;
}
__host__ __device__ inline void derive_stats(mcl_stats &s) {
    if (s.derived) return;
    const double s1 = s.sum_w;
    s.log_norm = s.max_logw + log(s1);
    s.max_weight = exp(s.raw[11] - s.max_logw) / s1;
    const double inv = 1.0 / s1;
    const double mdx = s.raw[0] * inv, mdy = s.raw[1] * inv, mdt = s.raw[2] * inv;
    const double mth = atan2(s.raw[3], s.raw[4]);       // circular mean of the yaw
    const double del = limit_angle_hd(mth - s.ref[2]);  // its offset from the reference pose of the sums
    s.mean[0] = s.ref[0] + mdx;
    s.mean[1] = s.ref[1] + mdy;
    s.mean[2] = mth;
    const double cxx = s.raw[5] * inv - mdx * mdx, cxy = s.raw[6] * inv - mdx * mdy, cyy = s.raw[7] * inv - mdy * mdy;
    const double cxt = s.raw[8] * inv - mdx * mdt, cyt = s.raw[9] * inv - mdy * mdt;
    const double ctt = s.raw[10] * inv - 2.0 * del * mdt + del * del;
    s.cov[0] = cxx; s.cov[1] = cxy; s.cov[2] = cxt;
    s.cov[3] = cxy; s.cov[4] = cyy; s.cov[5] = cyt;
    s.cov[6] = cxt; s.cov[7] = cyt; s.cov[8] = ctt;
    s.derived = 1;
}

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

// fp32 wrap of a residual into [-pi, pi] (statistics only; one turn is enough for |d| < 3 pi)
__device__ __forceinline__ float wrap_residual(float d) {
    if (d > 3.14159265f) d -= 6.28318531f;
    if (d < -3.14159265f) d += 6.28318531f;
    return d;
}

// optional phase timeline (mcl_debug_timeline): thread 0 of a CTA stamps the global nanosecond timer
constexpr int kDbgSlots = 12;
// (compiled in only with -DMCL_TIMELINE: MCL_TIMELINE=1 python slam-robot_simu_b200/build.py --force)
// Always compiled.  Used at ONE place in k_step (between the per-step constants and the particle loop) where
// this opaque, never-taken side effect makes ptxas allocate registers markedly better (76 instead of 124 bytes
// of spills under the 80-register cap; ~10 % of k_step at 125 M particles).  Measured, not understood.
__device__ __forceinline__ void dbg_stamp_real(unsigned long long *dbg, int slot) {
    if (dbg != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dbg[(static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * kDbgSlots + slot] = t;
    }
}
__device__ __forceinline__ void dbg_stamp(unsigned long long *dbg, int slot) {
#ifdef MCL_TIMELINE
    if (dbg != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dbg[(static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * kDbgSlots + slot] = t;
    }
#else
    (void)dbg;
    (void)slot;
#endif
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// atan2 in ~22 branch-free fp32 instructions (atan2f is ~60 with its special-case branches): octant reduction
// a = min/max in [0, 1] (MUFU.RCP), odd minimax polynomial of degree 15 (1.4e-7 rad evaluated in fp32),
// quadrant fix-up by selects.  (0, 0) -> 0.  Used per (particle, landmark) pair by the fan-sensor likelihood.
__device__ __forceinline__ float atan2_fast(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), 1.0e-30f), mn = fminf(ax, ay);
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(mx));
    const float a = mn * rc, s = a * a;
    float p = -4.054567019e-03f;
    p = fmaf(p, s, 2.186295735e-02f);
    p = fmaf(p, s, -5.591232636e-02f);
    p = fmaf(p, s, 9.642197334e-02f);
    p = fmaf(p, s, -1.390862957e-01f);
    p = fmaf(p, s, 1.994656566e-01f);
    p = fmaf(p, s, -3.332986079e-01f);
    p = fmaf(p, s, 9.999993356e-01f);
    float r = a * p;
    r = (ay > ax) ? 1.57079632679489662f - r : r;
    r = (x < 0.0f) ? 3.14159265358979324f - r : r;
    return copysignf(r, y);
}

// ---------------------------------------------------------------------------------------------
// Packed fp32 pairs (Blackwell FFMA2 / FADD2): one instruction, two lanes of arithmetic.  The pair loops of the
// fp32 likelihoods run two particles per register pair, which halves their issue slots.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// sin(d) and cos(d) - 1 of a SMALL angle with fp32 relative accuracy (what the shifted-origin likelihoods need:
// the absolute error of both is ~6e-8 |d|, not 6e-8).  |d| <= 0.5: odd / even Taylor polynomials (truncation
// below 2e-11); beyond that the particle is far from the reference pose, its weight is nil, and libm is used.
__device__ __forceinline__ void sincosm1_small(float d, float &sd, float &cm1) {
    if (fabsf(d) <= 0.5f) {
        const float d2 = d * d;
        float ps = 2.7557319e-6f;             //  1/9!
        ps = fmaf(ps, d2, -1.9841270e-4f);    // -1/7!
        ps = fmaf(ps, d2, 8.3333333e-3f);     //  1/5!
        ps = fmaf(ps, d2, -1.6666667e-1f);    // -1/3!
        sd = fmaf(d * d2, ps, d);
        float pc = -2.7557319e-7f;            // -1/10!
        pc = fmaf(pc, d2, 2.4801587e-5f);     //  1/8!
        pc = fmaf(pc, d2, -1.3888889e-3f);    // -1/6!
        pc = fmaf(pc, d2, 4.1666667e-2f);     //  1/4!
        pc = fmaf(pc, d2, -0.5f);
        cm1 = d2 * pc;
    } else {
        float c;
        sincosf(d, &sd, &c);
        cm1 = c - 1.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter-based: the noise of particle i at step k does not
// depend on launch geometry or on how the filter is sharded.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = static_cast<unsigned long long>(0xD2511F53u) * c.x;
        const unsigned long long p1 = static_cast<unsigned long long>(0xCD9E8D57u) * c.z;
        const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
        const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// The same with the ten round keys precomputed (host side, philox_round_keys): in a kernel they are operands
// from the parameter bank instead of 18 additions per call.
struct PhiloxKeys {
    uint32_t x[10], y[10];
};
__host__ __device__ inline PhiloxKeys philox_round_keys(unsigned long long seed) {
    PhiloxKeys rk;
    uint32_t kx = static_cast<uint32_t>(seed), ky = static_cast<uint32_t>(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        rk.x[r] = kx;
        rk.y[r] = ky;
        kx += 0x9E3779B9u;
        ky += 0xBB67AE85u;
    }
    return rk;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys &rk) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = static_cast<unsigned long long>(0xD2511F53u) * c.x;
        const unsigned long long p1 = static_cast<unsigned long long>(0xCD9E8D57u) * c.z;
        const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
        const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
        c = make_uint4(hi1 ^ c.y ^ rk.x[r], lo1, hi0 ^ c.w ^ rk.y[r], lo0);
    }
    return c;
}

// Two standard normals from two 32-bit words: Box-Muller in fp32 on the MUFU pipe.  The top 23 bits
// of each word become a float in [1,2) without touching the conversion unit; |n| <= 5.65.
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = 2.0f - __uint_as_float(0x3f800000u | (a >> 9));  // (0, 1]
    const float u2 = __uint_as_float(0x3f800000u | (b >> 9)) - 1.0f;  // [0, 1)
    const float r = sqrt_approx(-1.3862943611198906f * lg2_approx(u1));  // sqrt(-2 ln u1)
    float s, c;
    __sincosf(6.28318530717958647692f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

// the same from two 21-bit fields (k_step slices 256 Philox bits into twelve of them); |n| <= 5.4
__device__ __forceinline__ float2 box_muller21(uint32_t a, uint32_t b) {
    const float u1 = 2.0f - __uint_as_float(0x3f800000u | (a << 2));  // (0, 1]
    const float u2 = __uint_as_float(0x3f800000u | (b << 2)) - 1.0f;  // [0, 1)
    const float r = sqrt_approx(-1.3862943611198906f * lg2_approx(u1));
    float s, c;
    __sincosf(6.28318530717958647692f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

__device__ __forceinline__ double philox_uniform01(uint4 r) {
    return (static_cast<double>(r.x) * 4294967296.0 + static_cast<double>(r.y)) * (1.0 / 18446744073709551616.0);
}

// ---------------------------------------------------------------------------------------------
// The weight of a stored log-weight, on the 2^-52 grid.  ONE definition shared by the resampling
// scan, the shard totals and the export kernels, so all of them see identical integers:
//     q = rn_u64( ex2((lw - m) * log2e) * scale ),  scale = fl32(2^52 / sum_i exp(lw_i - m)),
//     w = q * 2^-52.
// fp32 + MUFU.EX2: relative error ~3e-7 against exp() in fp64, far inside the 1e-5 parity band;
// what must be exact -- the cumulative sums and the comparison with the comb -- is integer.
// ---------------------------------------------------------------------------------------------
struct WeightMap {
    float m, scale;
};
__device__ __forceinline__ float weight_scaled(float lw, WeightMap wm) {
    return __fmul_rn(ex2_approx(__fmul_rn(__fsub_rn(lw, wm.m), kLog2e)), wm.scale);
}
__device__ __forceinline__ unsigned long long quantize_weight(float lw, WeightMap wm) {
    return __float2ull_rn(weight_scaled(lw, wm));
}
// the same integer as a double (exact: rint of an fp32 value below 2^53)
__device__ __forceinline__ double quantize_weight_d(float lw, WeightMap wm) {
    return static_cast<double>(rintf(weight_scaled(lw, wm)));
}
__device__ __forceinline__ WeightMap make_weight_map(double max_logw, double sum_w) {
    WeightMap wm;
    wm.m = static_cast<float>(max_logw);
    wm.scale = static_cast<float>(kTwo52 / sum_w);
    return wm;
}

// u_i = fl(fl(i * (1/N)) + ofs): numpy rounds twice (np.arange then "+ ofs",
// particle_filter.py:213-215); the explicit _rn intrinsics stop nvcc contracting them into one FMA.
__device__ __forceinline__ double comb_value(double i, double delta, double ofs) {
    return __dadd_rn(__dmul_rn(i, delta), ofs);
}

// Number of comb teeth u_i (0 <= i < n) with u_i <= K * 2^-52, all in fp64 on integer-valued doubles
// (K < 2^53 and i < 2^32 are exact).  The walk of particle_filter.py:219 advances while
// u_i > pw_cum[idx]; with integer cumulative weights K_j the ancestor of slot i is the first j with
// K_j * 2^-52 >= u_i, so source j owns slots [count_le(K_{j-1}), count_le(K_j)).
// Branch-free: the real-arithmetic guess floor((c - ofs) N) is within one of the answer (its error is
// ~1e-15 N teeth), and the two neighbouring teeth are checked against the comb exactly as numpy rounds it.
__device__ __forceinline__ double count_le(double Kd, double delta, double ofs, double nd) {
    const double c = Kd * kInvTwo52;  // exact
    const double g = floor((c - ofs) * nd);
    const double g1 = g + 1.0;
    const double u1 = comb_value(g1, delta, ofs);
    const double u0 = comb_value(g, delta, ofs);
    double r = g1;                 // count if tooth g is the last one at or below c
    if (u1 <= c) r = g1 + 1.0;
    else if (u0 > c) r = g;
    return (r > nd) ? nd : r;      // r >= 0 always: c >= 0 and ofs < 1/N give g >= -1
}

// The same count when the real-arithmetic position a = (K 2^-52 - ofs) N is safely inside a tooth
// interval: one FMA, one floor.  The rounding of a (<= 1.2e-16 N) plus the deviation of numpy's doubly
// rounded comb from i/N + ofs (<= 2.3e-16 N) is below 8e-7 teeth for N < 2^31, so a fractional part in
// [2^-16, 1 - 2^-16] decides exactly; anything closer to an integer takes the exact path above.
struct CombFast {
    double a1, a0, nd;  // a = fma(K, a1, a0): a1 = N 2^-52 (exact), a0 = -fl(ofs N)
};
__device__ __forceinline__ CombFast make_comb_fast(double ofs, double nd) {
    CombFast f;
    f.a1 = nd * kInvTwo52;
    f.a0 = -(ofs * nd);
    f.nd = nd;
    return f;
}
// The fast count in integer registers: t = fma(K, a1, a0 + 1.5 2^36) is a rounded to the 2^-16 grid with the
// integer round(a 2^16) in its mantissa, so floor(a) and the distance to the nearest integer are bit fields --
// no floor, no fp64 compares, no conversion.  a0 + 1.5 2^36 is rounded once more (<= 2^-17), which puts the
// total error at 1.04 grid steps: fractions within 2 steps of an integer are reported through `ok` (it becomes
// 0) and the caller redoes those counts with count_le_fast.  a > -1, and a in (-1, 0) yields 0 (mod 2^32).
constexpr double kCombMagic = 103079215104.0;  // 1.5 * 2^36
__device__ __forceinline__ uint32_t count_le_bits(double Kd, double a1, double a0m, uint32_t &ok) {
    const double t = fma(Kd, a1, a0m);
    const uint32_t lo = static_cast<uint32_t>(__double2loint(t)), hi = static_cast<uint32_t>(__double2hiint(t));
    ok = min(ok, (lo + 2u) & 0xfffcu);
    return __funnelshift_r(lo, hi, 16) + 1u;
}
__device__ __forceinline__ double count_le_fast(double Kd, const CombFast &f, double delta, double ofs) {
    const double a = fma(Kd, f.a1, f.a0);
    const double g = floor(a);
    const double fr = a - g;
    if (fr < 0x1p-16 || fr > 1.0 - 0x1p-16) return count_le(Kd, delta, ofs, f.nd);
    const double r = g + 1.0;
    return (r > f.nd) ? f.nd : r;
}

// ---------------------------------------------------------------------------------------------
// sin/cos of an fp32 angle to ~1e-13 with 12 fp64 operations and ONE conversion:
//   k = rint(th * 32/pi) by the magic-number trick (no F2I), r_hi = th - k * HI exactly in fp32
//   (HI = pi/32 rounded to 17 bits), r = r_hi + corr[k] with corr[k] = -k (pi/32 - HI) tabulated,
//   sin r / cos r by short Taylor polynomials (|r| <= pi/64), rotated by the tabulated sin/cos(k pi/32).
// The table (256 entries of 32 B, k in [-128, 127]) is built on the host and read through L1.
// ---------------------------------------------------------------------------------------------
constexpr int kTrigEntries = 256;
constexpr float kTrigInvStep = 10.185916357881302f;  // 32 / pi
constexpr float kTrigStepHi = 0x1.921fp-4f;          // pi/32 to 17 significant bits
constexpr double kTrigStepLo = 6.752708724411249e-07;  // pi/32 - kTrigStepHi

struct __align__(32) TrigEntry {
    double s, c, corr, pad;
};

__device__ __forceinline__ void sincos_tab(float th, const TrigEntry *__restrict__ tab, double &s, double &c) {
    const float t = fmaf(th, kTrigInvStep, 12582912.0f);
    const int k = __float_as_int(t) - 0x4B400000;
    if (abs(k) > 127 || !(fabsf(th) < 16.0f)) {  // outside the table (or NaN): the libm path
        sincos(static_cast<double>(th), &s, &c);
        return;
    }
    const float kf = t - 12582912.0f;
    const float rh = fmaf(-kf, kTrigStepHi, th);  // exact: k has 7 bits, HI 17, and the result is small
    const double2 sc = __ldg(reinterpret_cast<const double2 *>(tab + (k + 128)));  // L1-resident 8 KB table
    const double r = static_cast<double>(rh) + __ldg(&tab[k + 128].corr);
    const double r2 = r * r;
    const double ps = fma(r2, 1.0 / 120.0, -1.0 / 6.0) * r2;
    const double sr = fma(r, ps, r);
    double pc = fma(r2, -1.0 / 720.0, 1.0 / 24.0);
    pc = fma(r2, pc, -0.5);
    const double cr = fma(r2, pc, 1.0);
    s = fma(sc.x, cr, sc.y * sr);
    c = fma(sc.y, cr, -(sc.x * sr));
}

// ---------------------------------------------------------------------------------------------
// Particle-sharded filter, device-driven exchange over peer memory (NVLink / NVSwitch).
// Every shard owns one PeerBox plus a receive staging area in memory its peers can write (cudaMalloc + CUDA IPC
// between processes, plain pointers inside one process).  A shard WRITES into its peers' boxes and POLLS only its
// own: record / total / "expansion done" flags carry the step number (never reset), data before flag with a
// system-scope fence in between.  Nothing on this path goes through the host or through NCCL.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxShards = 16;

struct PeerBox {
    mcl_stats rec[kMaxShards];              // rec[g]: statistics record of shard g for the step in flag_stats[g]
    unsigned long long total[kMaxShards];   // total[g]: integer weight total of shard g for the step in flag_total[g]
    unsigned int flag_stats[kMaxShards];
    unsigned int flag_total[kMaxShards];
    unsigned int flag_done[kMaxShards];     // shard g has written every pose it owes this shard (step number)
    unsigned int pad[16];
};

// Per-step plan of a shard's resampling, derived ON THE DEVICE from the exchanged totals (local memory): the public
// struct of include/mcl.h.
using ShardPlan = mcl_shard_plan;
static_assert(kMaxShards == MCL_MAX_SHARDS, "mcl.h and the kernels disagree on the shard limit");

// Where a shard reaches its peers (device memory, constant for the handle's lifetime).
struct PeerView {
    PeerBox *box[kMaxShards];                                    // box[me] is local
    float *rx[kMaxShards], *ry[kMaxShards], *rth[kMaxShards];    // receive staging of shard g, indexed by g's LOCAL slot
    int me, G;
    long long m;                                                 // particles (= output slots) per shard
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// the early-dependency waits of k_scan / k_step give up after this long and use cudaGridDependencySynchronize instead
constexpr unsigned long long kEarlyWaitNs = 5000000ull;
__device__ __forceinline__ unsigned long long global_timer_ns();
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Wait until *flag has reached `epoch` (flags only grow).  A peer that never arrives must not hang the GPU: after
// kPeerTimeoutNs the wait gives up and raises the plan's sticky error, which every later wait honours at once.
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ void peer_wait(const unsigned int *flag, unsigned int epoch, ShardPlan *plan) {
    if (*reinterpret_cast<volatile int *>(&plan->error)) return;
    const unsigned long long t0 = global_timer_ns();
    while (static_cast<int>(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(200);
        if (global_timer_ns() - t0 > kPeerTimeoutNs) {
            *reinterpret_cast<volatile int *>(&plan->error) = 1;
            return;
        }
    }
}
// whether the step resamples: host flags first, else the gate the previous statistics exchange left in the plan
__device__ __forceinline__ int plan_resample(uint32_t flags, const ShardPlan *plan) {
    if (flags & MCL_NO_RESAMPLE) return 0;
    if (flags & MCL_FORCE_RESAMPLE) return 1;
    return *reinterpret_cast<const volatile int *>(&plan->gate);
}

// ---------------------------------------------------------------------------------------------
// block-wide sum of NV doubles (256 threads); result valid in all threads.
// smem must hold (blockDim/32) * NV doubles.
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
