// Kernels of the MCL step (sm_100a).  One translation unit: included by mcl_api.cu.
//
// A filter step is TWO launches:
//   k_scan   tile CTAs run the decoupled look-back scan of the 64-bit fixed-point weights and write the ancestor
//            of every slot (trials that do not resample leave early); two extra CTAs per trial derive the per-step
//            scalars (resample gate, comb offset, reference pose, landmark table + moments, likelihood shift) and
//            merge the previous step's CTA partials into mcl_stats.
//   k_step   [gather through the ancestors] + predict + weight + statistics; every CTA leaves one 128-byte partial
//            record -- the records are merged where they are consumed (next k_scan, k_merge, k_xchg_stats).
// Small steps are chained by programmatic dependent launch, and the kernels' ends overlap: a CTA that is resident
// before its predecessor grid has drained waits for exactly what it reads -- k_scan's merging CTAs for the count of
// finished k_step CTAs, k_step's set-up for the scalars CTA's flag -- and for the whole grid only where it must.
// Stage / utility kernels: k_reduce (statistics only), k_gather (stand-alone resampling gather),
// k_cov (exact two-pass covariance), k_export_weights, k_init*, k_fill, k_reset_scan.
#pragma once
#include "mcl_device.cuh"

namespace mcl {

struct StateView {
    float *x, *y, *th, *lw;
};

// likelihood variants of k_step (template parameter)
constexpr int kLikPair = 0;       // per (particle, landmark) pair, landmarks staged in shared memory
constexpr int kLikMomentIso = 1;  // 8 per-step moments, sigma_x == sigma_y (the reference's R)
constexpr int kLikMomentGen = 2;  // 13 per-step moments, sigma_x != sigma_y
constexpr int kLikRB = 3;         // range / bearing residuals with limit_angle wrap
constexpr int kLikRBFan = 4;      // fan-shaped scan sensor: detected landmarks only, visibility gate, fp32 per pair
constexpr int kLikPair32 = 5;     // per pair in fp32 about the reference pose (shifted origin), two particles per FFMA2
constexpr int kLikRB32 = 6;       // range / bearing residuals of ALL landmarks in fp32 (MUFU sqrt / rcp, branch-free atan2)

__device__ __forceinline__ void pair_residual_c(double s, double c, double tx, double ty, double4 t, double &ex,
                                                double &ey) {
    // robot-frame landmark minus observation; mylib/transform.py:31-35 with yaw' = pi/2 - theta:
    //   rx = sin(th)*(lx-x) - cos(th)*(ly-y),  ry = cos(th)*(lx-x) + sin(th)*(ly-y)
    ex = fma(s, t.x, fma(-c, t.y, tx - t.z));
    ey = fma(c, t.x, fma(s, t.y, ty - t.w));
}

__device__ __forceinline__ void pair_residual_rb(double s, double c, double tx, double ty, double4 t, double &ed,
                                                 double &ephi) {
    const double rx = fma(s, t.x, fma(-c, t.y, tx));
    const double ry = fma(c, t.x, fma(s, t.y, ty));
    ed = t.z - sqrt(fma(rx, rx, ry * ry));
    ephi = limit_angle(t.w - atan2(ry, rx));
}

// The fan sensor (MCL_LIK_RB_FAN).  Table entry t = (lx, ly, 1 / (d_obs r_dist), phi_obs) of a DETECTED landmark;
// (rx, ry) its position in the particle's frame.  Returns the chi^2 of the pair: the two normalised residuals
//   (d_obs - d) / (d_obs r_dist) = 1/r_dist - d t.z        (distance noise grows with the distance, :163)
//   limit_angle(phi_obs - phi) / dir_sigma
// when the landmark is inside the particle's own fan (d <= range and ry >= |rx| tan(pi/2 - angle), :155-160),
// else `miss`.  Same expression in fp64 (reference pose, once per step) and in fp32 (every pair).
template <typename F>
__device__ __forceinline__ F fan_pair_chi2(F rx, F ry, F tz, F tw, F range2, F fan_tan, F inv_r, F inv_sphi, F miss);
template <>
__device__ __forceinline__ double fan_pair_chi2<double>(double rx, double ry, double tz, double tw, double range2,
                                                        double fan_tan, double inv_r, double inv_sphi, double miss) {
    const double d2 = fma(rx, rx, ry * ry);
    const bool vis = d2 <= range2 && ry >= fabs(rx) * fan_tan;
    const double ed = inv_r - sqrt(d2) * tz;
    const double ep = limit_angle(tw - atan2(ry, rx)) * inv_sphi;
    return vis ? fma(ed, ed, ep * ep) : miss;
}
template <>
__device__ __forceinline__ float fan_pair_chi2<float>(float rx, float ry, float tz, float tw, float range2, float fan_tan,
                                                      float inv_r, float inv_sphi, float miss) {
    const float d2 = fmaf(rx, rx, ry * ry);
    const bool vis = d2 <= range2 && ry >= fabsf(rx) * fan_tan;
    const float ed = fmaf(-sqrt_approx(d2), tz, inv_r);
    const float ep = wrap_residual(tw - atan2_fast(ry, rx)) * inv_sphi;
    return vis ? fmaf(ed, ed, ep * ep) : miss;
}

// -------------------------------------------------------------------------------------------------
// per-step scalars ("prep"), executed by block 0 of each trial inside k_scan
// -------------------------------------------------------------------------------------------------
struct PrepParams {
    const double *lm;       // [L][2]
    const double *z;        // [T][L][2]  (null: decision fields only, stage API)
    const double *ofs;      // [T] or null
    PrepOut *prep;          // [T]
    double4 *tab;           // [T][Lmax]
    int L, Lmax;
    uint32_t flags;
    int likelihood;
    double ess_threshold, vdt, wdt, inv_s0sq, inv_s1sq, lik_const, inv_n;
    double fan_range, fan_tan, fan_miss, r_dist, dir_sigma;  // MCL_LIK_RB_FAN
    double n_global;        // particles of the whole filter (the uniform reset of particle_filter.py:234-237)
    int underflow_rule;     // mcl_config::underflow_rule
    uint64_t seed;
    uint32_t step;
};

// resample gate and comb offset of trial t: every block of the trial evaluates the same expression
__device__ __forceinline__ int decide_resample(uint32_t flags, double ess, double ess_threshold) {
    if (flags & MCL_NO_RESAMPLE) return 0;
    if (flags & MCL_FORCE_RESAMPLE) return 1;
    return (ess < ess_threshold) ? 1 : 0;  // particle_filter.py:211
}
__device__ __forceinline__ double comb_offset(const PrepParams &p, int t) {
    if (p.ofs != nullptr) return p.ofs[t];
    if (p.flags & MCL_PHILOX_NOISE) {
        const uint4 r = philox4x32_10(make_uint4(0xFFFFFFFFu, 0u, p.step, static_cast<uint32_t>(t)),
                                      make_uint2(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32)));
        return __dmul_rn(philox_uniform01(r), p.inv_n);  // rand() * NP_RECIP, particle_filter.py:214
    }
    return 0.0;
}

// all kScanThreads threads of one block; red: >= 8 * 15 doubles of shared memory
struct PrepIn {
    double est[3];     // MAP pose of the previous step
    double max_logw;   // maximum and sum of the stored weights (weight map)
    double sum_w;
    int uniform;       // the stored weights underflowed (reference arithmetic): all-equal, sum_w = n_global
};
__device__ __noinline__ void prep_block(const PrepParams &p, const PrepIn &in, const TrigEntry *trig, int t,
                                        int do_resample, double ofs, double *red) {
    const int tid = threadIdx.x;
    PrepOut *po = p.prep + t;
    // ---- reference pose: previous MAP estimate pushed through the noise-free motion (__f) ----
    // (est is a stored fp32 pose; the table sincos keeps this once-per-step code short)
    double s0, c0;
    sincos_tab(static_cast<float>(in.est[2]), trig, s0, c0);
    // the kernels subtract the reference in fp32: make it exactly representable
    const double rx = static_cast<double>(static_cast<float>(in.est[0] + p.vdt * c0));
    const double ry = static_cast<double>(static_cast<float>(in.est[1] + p.vdt * s0));
    const double rth = static_cast<double>(static_cast<float>(limit_angle(in.est[2] + p.wdt)));

    // ---- landmark table, moments, likelihood of the reference pose ----
    double acc[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) acc[k] = 0.0;
    int L_out = p.L;
    if (p.z != nullptr && p.likelihood == MCL_LIK_RB_FAN) {
        // Detected landmarks (finite, positive distance) are compacted into the table -- the pair loop of k_step
        // then runs over them only -- and the reference pose is scored with the same expression in fp64.
        __shared__ int s_cnt[32];
        double s, c;
        sincos_tab(static_cast<float>(rth), trig, s, c);
        const int lane = tid & 31, warp = tid >> 5, nw = (static_cast<int>(blockDim.x) + 31) >> 5;
        const double range2 = p.fan_range * p.fan_range, inv_r = 1.0 / p.r_dist, inv_sphi = 1.0 / p.dir_sigma;
        int base = 0;
        for (int l0 = 0; l0 < p.L; l0 += blockDim.x) {
            const int l = l0 + tid;
            double zd = 0.0, zp = 0.0;
            bool seen = false;
            if (l < p.L) {
                zd = __ldcv(&p.z[(static_cast<long long>(t) * p.L + l) * 2]);
                zp = __ldcv(&p.z[(static_cast<long long>(t) * p.L + l) * 2 + 1]);
                seen = zd > 0.0 && zd < 1.0e300 && zp == zp;  // NaN fails every comparison
            }
            const unsigned int b = __ballot_sync(0xffffffffu, seen);
            if (lane == 0) s_cnt[warp] = __popc(b);
            __syncthreads();
            int pos = base, tot = 0;
        #pragma unroll
    for (int w = 0; w < nw; ++w) {
                if (w < warp) pos += s_cnt[w];
                tot += s_cnt[w];
            }
            pos += __popc(b & ((1u << lane) - 1u));
            if (seen) {
                const double lx = p.lm[2 * l], ly = p.lm[2 * l + 1];
                const double4 e = make_double4(lx, ly, 1.0 / (zd * p.r_dist), zp);
                p.tab[static_cast<long long>(t) * p.Lmax + pos] = e;
                const double dx = lx - rx, dy = ly - ry;
                acc[13] += fan_pair_chi2<double>(s * dx - c * dy, c * dx + s * dy, e.z, e.w, range2, p.fan_tan, inv_r,
                                                 inv_sphi, p.fan_miss);
            }
            base += tot;
            __syncthreads();
        }
        L_out = base;
        block_sum<15>(acc, red);
    } else if (p.z != nullptr) {
        double s, c;
        sincos_tab(static_cast<float>(rth), trig, s, c);
        const double tx = -(s * rx - c * ry), ty = -(c * rx + s * ry);
        for (int l = tid; l < p.L; l += blockDim.x) {
            const double lx = p.lm[2 * l], ly = p.lm[2 * l + 1];
            const double zx = __ldcv(&p.z[(static_cast<long long>(t) * p.L + l) * 2]);  // (z may be mapped host memory: never a cached copy)
            const double zy = __ldcv(&p.z[(static_cast<long long>(t) * p.L + l) * 2 + 1]);
            const double4 e = make_double4(lx, ly, zx, zy);
            if (p.likelihood != MCL_LIK_PAIR_F32 && p.likelihood != MCL_LIK_RB_F32) p.tab[static_cast<long long>(t) * p.Lmax + l] = e;
            acc[0] += lx;       acc[1] += ly;       acc[2] += zx;       acc[3] += zy;
            acc[4] += lx * lx;  acc[5] += ly * ly;  acc[6] += lx * ly;
            acc[7] += zx * zx;  acc[8] += zy * zy;
            acc[9] += lx * zx;  acc[10] += ly * zx; acc[11] += lx * zy; acc[12] += ly * zy;
            double e0, e1;
            if (p.likelihood == MCL_LIK_RB || p.likelihood == MCL_LIK_RB_F32) pair_residual_rb(s, c, tx, ty, e, e0, e1);
            else pair_residual_c(s, c, tx, ty, e, e0, e1);
            if (p.likelihood == MCL_LIK_RB_F32) {
                // fp32 range/bearing kernel about the reference pose: the landmark as the reference pose sees it
                // (g, |g|), and the reference pose's own residuals (e0 = z_d - |g|, e1 = wrap(z_phi - phi_g)).  A particle
                // changes the range by dr and the bearing by dphi, both small and computed with RELATIVE fp32 accuracy;
                // the kernel sums dr'(dr' - 2 e0') and dphi (dphi - 2 e1) -- differences of squares, primes = / sigma.
                const double gx = fma(s, lx, fma(-c, ly, tx)), gy = fma(c, lx, fma(s, ly, ty));
                const double g2 = fmax(fma(gx, gx, gy * gy), 1.0e-12), rho = sqrt(g2);
                float *slot = reinterpret_cast<float *>(&p.tab[static_cast<long long>(t) * p.Lmax + l]);
                slot[0] = static_cast<float>(gx);
                slot[1] = static_cast<float>(gy);
                slot[2] = static_cast<float>(rho * sqrt(p.inv_s0sq));        // |g| / sigma_d
                slot[3] = static_cast<float>(2.0 / g2);
                slot[4] = static_cast<float>(g2);
                slot[5] = static_cast<float>(2.0 * e0 * sqrt(p.inv_s0sq));   // 2 e_d(ref) / sigma_d
                slot[6] = static_cast<float>(e1);                            // e_phi(ref), wrapped
                slot[7] = static_cast<float>(2.0 * e1);                      // 2 e_phi(ref); 1 / sigma_phi^2 where the sums are folded
            }
            acc[13] = fma(e0, e0, acc[13]);
            acc[14] = fma(e1, e1, acc[14]);
            // fp32 pair kernel: the landmark as the reference pose sees it, g = M_r (l - p_r) = residual + observation,
            // and twice the residual u there -- a particle's residual is u + d with d small, and the kernel sums
            // d (d + 2 u) = |u + d|^2 - |u|^2, the difference to the reference pose's chi^2, never the squares
            if (p.likelihood == MCL_LIK_PAIR_F32)
                p.tab[static_cast<long long>(t) * p.Lmax + l] = make_double4(e0 + zx, e1 + zy, 2.0 * e0, 2.0 * e1);
        }
        block_sum<15>(acc, red);
    }
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 13; ++k) po->mom[k] = acc[k];
        if (p.likelihood == MCL_LIK_RB_FAN) po->c_shift = -0.5 * acc[13];  // residuals are already normalised
        else po->c_shift = (p.z != nullptr) ? (-0.5 * (acc[13] * p.inv_s0sq + acc[14] * p.inv_s1sq) + p.lik_const) : 0.0;
        po->ref[0] = rx;
        po->ref[1] = ry;
        po->ref[2] = rth;
        po->max_prev = in.max_logw;
        po->uniform = in.uniform;
        // absolute scale of the weights this step will store: log(pw_prev) = stored - log(sum) (or -log NP after a
        // resampling or a uniform reset), log(bn) = log-likelihood; the stored value is relative to c_shift
        const bool cartesian = p.likelihood == MCL_LIK_PAIR || p.likelihood == MCL_LIK_MOMENT || p.likelihood == MCL_LIK_PAIR_F32;
        po->abs_shift = (p.z != nullptr && cartesian && p.underflow_rule) ? po->c_shift - log(do_resample ? p.n_global : in.sum_w) : kNoAbsShift;
        po->lik_const = p.lik_const;
        const WeightMap wm = make_weight_map(in.max_logw, in.sum_w);
        po->wq_m = wm.m;
        po->wq_scale = wm.scale;
        po->L = L_out;
        po->do_resample = do_resample;
        po->ofs = ofs;
    }
}

// -------------------------------------------------------------------------------------------------
// Running statistics.  Every warp keeps ONE running maximum (warp-uniform, so rescaling is a
// uniform branch) and fp32 sums relative to it; every kFlushEvery iterations the sums are reduced
// across the warp and folded into fp64 sums held one-per-lane (lane k owns sum k).
//   sums: 0 s1  1 s2  2 dx 3 dy 4 dth 5 sin 6 cos 7 xx 8 xy 9 yy 10 xth 11 yth 12 thth
// sin/cos (5, 6) use the fp64 sin/cos of the likelihood and are kept per thread in fp64.
// -------------------------------------------------------------------------------------------------
struct WarpStats {
    float m;            // warp-uniform running max of the stored log-weights (finite sentinel)
    float g[11];        // fp32 group sums: s1 s2 dx dy dth xx xy yy xth yth thth
    double asin_, acos_;  // per-thread fp64 sums of e sin(th), e cos(th)
    double wacc;        // lane k < 11: fp64 sum k (order as g[])
    float best_v;
    uint32_t best_i;

    __device__ __forceinline__ void init() {
        m = -3.4028234e38f;
        best_v = -INFINITY;
        best_i = 0xffffffffu;
        asin_ = acos_ = wacc = 0.0;
#pragma unroll
        for (int k = 0; k < 11; ++k) g[k] = 0.f;
    }
    // called by ALL lanes of the warp with the largest log-weight of their current group (-inf if none)
    __device__ __forceinline__ void raise_max(float gm) {
        if (__any_sync(0xffffffffu, gm > m)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
            const float r = ex2_approx((m - gm) * kLog2e);  // 0 the first time
            const float r2 = r * r;
            const int lane = threadIdx.x & 31;
#pragma unroll
            for (int k = 0; k < 11; ++k) g[k] *= (k == 1) ? r2 : r;  // s2 carries exp(2 (lw - m))
            const double rd = r;
            asin_ *= rd;
            acos_ *= rd;
            wacc *= (lane == 1) ? static_cast<double>(r2) : rd;
            m = gm;
        }
    }
    __device__ __forceinline__ void add(float lw, float x, float y, float th, double s, double c, float rx, float ry,
                                        float rth, uint32_t idx) {
        if (lw > best_v) {  // strict: first occurrence wins (np.argmax, particle_filter.py:116)
            best_v = lw;
            best_i = idx;
        }
        const float e = ex2_approx((lw - m) * kLog2e);  // lw = -inf gives 0 (m is finite)
        const float dx = x - rx, dy = y - ry, dth = wrap_residual(th - rth);
        const float ex = e * dx, ey = e * dy, et = e * dth;
        g[0] += e;
        g[1] = fmaf(e, e, g[1]);
        g[2] += ex;
        g[3] += ey;
        g[4] += et;
        g[5] = fmaf(ex, dx, g[5]);
        g[6] = fmaf(ex, dy, g[6]);
        g[7] = fmaf(ey, dy, g[7]);
        g[8] = fmaf(ex, dth, g[8]);
        g[9] = fmaf(ey, dth, g[9]);
        g[10] = fmaf(et, dth, g[10]);
        const double ed = e;
        asin_ = fma(ed, s, asin_);
        acos_ = fma(ed, c, acos_);
    }
    // warp-wide: all lanes
    __device__ __forceinline__ void flush() {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            float v = g[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == k) wacc += static_cast<double>(v);
            g[k] = 0.f;
        }
    }
};

__device__ __forceinline__ uint32_t float_order_key(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// -------------------------------------------------------------------------------------------------
// CTA partials and their merges.
// Every k_step / k_reduce CTA leaves ONE 128-byte record {max, 13 sums relative to it, arg-max value, index}
// in partials[] and exits -- no fence, no ticket, no "last CTA".  The records are merged where they are
// consumed, each consumer reading only the columns it needs:
//   merge_light  (max, sum w, sum w^2)  every tile CTA of the next k_scan: weight map + ESS gate
//   merge_light_best (+ arg-max)        the scalars CTA of k_scan: MAP pose -> reference pose
//   merge_full   (all 14 columns)       the statistics CTA of k_scan, or k_merge on demand: mcl_stats
// All merges run the same code on the same records in the same order: every CTA derives bit-identical values.
// -------------------------------------------------------------------------------------------------
// rec_count: the last thing a producer CTA does is count itself as finished, releasing everything it stored (its poses
// and weights, its record, CTA 0's mcl_stats fields).  The NEXT kernel's merging CTAs (k_scan: resident 2 - 15 us early
// under programmatic dependent launch, tools/timeline.py --back-to-back) take the count as their dependency instead of
// the completion of the grid.
__device__ __forceinline__ void partial_write(WarpStats &ws, double *pp, unsigned long long *dbg, unsigned int *rec_count) {
    constexpr int NW = kStepThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double s_sum[NW][kNumSums];
    __shared__ float s_m[NW];
    __shared__ uint32_t s_bk[NW], s_bi[NW];

    ws.flush();
    dbg_stamp(dbg, 3);
    double as = ws.asin_, ac = ws.acos_;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        as += __shfl_xor_sync(0xffffffffu, as, o);
        ac += __shfl_xor_sync(0xffffffffu, ac, o);
    }
    // arg-max of the warp: largest value, then smallest index
    const uint32_t key = float_order_key(ws.best_v);
    const uint32_t wkey = __reduce_max_sync(0xffffffffu, key);
    const uint32_t widx = __reduce_min_sync(0xffffffffu, (key == wkey) ? ws.best_i : 0xffffffffu);
    // sums in the order of the partial record: s1 s2 dx dy dth sin cos xx xy yy xth yth thth
    if (lane < 5) s_sum[warp][lane] = ws.wacc;
    if (lane >= 5 && lane < 11) s_sum[warp][lane + 2] = ws.wacc;
    if (lane == 0) {
        s_sum[warp][5] = as;
        s_sum[warp][6] = ac;
        s_m[warp] = ws.m;
        s_bk[warp] = wkey;
        s_bi[warp] = widx;
    }
    __syncthreads();
    if (tid < kNumSums + 1) {
        float M = s_m[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) M = fmaxf(M, s_m[w]);
        if (tid < kNumSums) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float r = ex2_approx((s_m[w] - M) * kLog2e);
                acc = fma(s_sum[w][tid], static_cast<double>((tid == 1) ? r * r : r), acc);
            }
            pp[1 + tid] = acc;
        } else {
            uint32_t bk = s_bk[0], bi = s_bi[0];
#pragma unroll
            for (int w = 1; w < NW; ++w)
                if (s_bk[w] > bk || (s_bk[w] == bk && s_bi[w] < bi)) { bk = s_bk[w]; bi = s_bi[w]; }
            pp[0] = static_cast<double>(M);
            pp[14] = static_cast<double>(float_from_order_key(bk));
            pp[15] = static_cast<double>(bi);
        }
    }
    // (barrier, then ONE releasing reduction: the release is cumulative over what the barrier ordered before it --
    //  the semaphore idiom of CUTLASS; a fence in every thread cost the tail 0.5 us)
    if (rec_count != nullptr) {  // (null: the statistics merge follows at once -- mcl_step_async -- and nobody polls)
        __syncthreads();
        if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(rec_count) : "memory");
    }
    dbg_stamp(dbg, 4);
}

struct MergeScratch {  // shared memory of one merge (256 threads)
    float m[8];
    double a[8], b[8];
    uint32_t k[8], i[8];
};

// (max, sum w, sum w^2) over the nb partial records of trial t; all 256 threads, result in every thread
template <int NT>
__device__ __noinline__ void merge_light_raw(const double *pp, int nb, MergeScratch &sc, float &M, double &s1, double &s2);
// ... with the reference's underflow rule applied: `uf` and the uniform (0, N, N) when it fired
template <int NT>
__device__ __forceinline__ void merge_light(const double *pp, int nb, MergeScratch &sc, float &M, double &s1, double &s2,
                                            double abs_shift, double n_global, int &uf) {
    merge_light_raw<NT>(pp, nb, sc, M, s1, s2);
    uf = weights_underflowed(static_cast<double>(M), s1, abs_shift) ? 1 : 0;
    if (uf) {
        M = 0.0f;
        s1 = n_global;
        s2 = n_global;
    }
}
template <int NT>
__device__ __noinline__ void merge_light_raw(const double *pp, int nb, MergeScratch &sc, float &M, double &s1, double &s2) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nw = NT / 32;
    float mx = -3.4028234e38f;
    double a1 = 0.0, a2 = 0.0;
    constexpr int kB = 4;
    for (int b0 = tid; b0 < nb; b0 += NT * kB) {
        double m4[kB], p4[kB], q4[kB];
#pragma unroll
        for (int u = 0; u < kB; ++u) {  // clamped rows, masked below: the 12 loads are independent
            const double *r = pp + min(b0 + u * NT, nb - 1) * kPart;
            m4[u] = __ldcg(r);
            p4[u] = __ldcg(r + 1);
            q4[u] = __ldcg(r + 2);
        }
#pragma unroll
        for (int u = 0; u < kB; ++u) {
            if (b0 + u * NT < nb) {
                const float mb = static_cast<float>(m4[u]);
                if (mb > mx) {
                    const float r = ex2_approx((mx - mb) * kLog2e);
                    a1 *= static_cast<double>(r);
                    a2 *= static_cast<double>(r * r);
                    mx = mb;
                }
                const float r = ex2_approx((mb - mx) * kLog2e);
                a1 = fma(p4[u], static_cast<double>(r), a1);
                a2 = fma(q4[u], static_cast<double>(r * r), a2);
            }
        }
    }
    // warp, then CTA: rescale to the common maximum and add
    float wm = mx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    {
        const float r = ex2_approx((mx - wm) * kLog2e);
        a1 *= static_cast<double>(r);
        a2 *= static_cast<double>(r * r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    __syncthreads();  // the scratch may still be in use by a previous merge
    if (lane == 0) { sc.m[warp] = wm; sc.a[warp] = a1; sc.b[warp] = a2; }
    __syncthreads();
    M = sc.m[0];
#pragma unroll
    for (int w = 1; w < nw; ++w) M = fmaxf(M, sc.m[w]);
    s1 = 0.0;
    s2 = 0.0;
#pragma unroll
    for (int w = 0; w < nw; ++w) {
        const float r = ex2_approx((sc.m[w] - M) * kLog2e);
        s1 = fma(sc.a[w], static_cast<double>(r), s1);
        s2 = fma(sc.b[w], static_cast<double>(r * r), s2);
    }
}

// (max, sum w, sum w^2) AND the arg-max in one pass over the records: the scalars CTA of k_scan needs both, and k_step
// cannot start before that CTA is done -- a second pass over the (cold) records cost it ~2 us at C1
template <int NT>
__device__ __noinline__ void merge_light_best(const double *pp, int nb, MergeScratch &sc, float &M, double &s1, double &s2,
                                              float &best_v, uint32_t &best_i) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nw = NT / 32;
    float mx = -3.4028234e38f;
    double a1 = 0.0, a2 = 0.0;
    uint32_t bk = 0u, bi = 0xffffffffu;
    constexpr int kB = 2;
    for (int b0 = tid; b0 < nb; b0 += NT * kB) {
        double m4[kB], p4[kB], q4[kB], v4[kB], i4[kB];
#pragma unroll
        for (int u = 0; u < kB; ++u) {  // clamped rows, masked below: the loads are independent
            const double *r = pp + min(b0 + u * NT, nb - 1) * kPart;
            m4[u] = __ldcg(r);
            p4[u] = __ldcg(r + 1);
            q4[u] = __ldcg(r + 2);
            v4[u] = __ldcg(r + 14);
            i4[u] = __ldcg(r + 15);
        }
#pragma unroll
        for (int u = 0; u < kB; ++u) {
            if (b0 + u * NT < nb) {
                const float mb = static_cast<float>(m4[u]);
                if (mb > mx) {
                    const float r = ex2_approx((mx - mb) * kLog2e);
                    a1 *= static_cast<double>(r);
                    a2 *= static_cast<double>(r * r);
                    mx = mb;
                }
                const float r = ex2_approx((mb - mx) * kLog2e);
                a1 = fma(p4[u], static_cast<double>(r), a1);
                a2 = fma(q4[u], static_cast<double>(r * r), a2);
                const uint32_t k = float_order_key(static_cast<float>(v4[u]));
                const uint32_t i = static_cast<uint32_t>(i4[u]);
                if (k > bk || (k == bk && i < bi)) { bk = k; bi = i; }
            }
        }
    }
    float wm = mx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    {
        const float r = ex2_approx((mx - wm) * kLog2e);
        a1 *= static_cast<double>(r);
        a2 *= static_cast<double>(r * r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    const uint32_t wk = __reduce_max_sync(0xffffffffu, bk);
    const uint32_t wi = __reduce_min_sync(0xffffffffu, (bk == wk) ? bi : 0xffffffffu);
    __syncthreads();  // the scratch may still be in use by a previous merge
    if (lane == 0) { sc.m[warp] = wm; sc.a[warp] = a1; sc.b[warp] = a2; sc.k[warp] = wk; sc.i[warp] = wi; }
    __syncthreads();
    M = sc.m[0];
    bk = sc.k[0];
    bi = sc.i[0];
#pragma unroll
    for (int w = 1; w < nw; ++w) {
        M = fmaxf(M, sc.m[w]);
        if (sc.k[w] > bk || (sc.k[w] == bk && sc.i[w] < bi)) { bk = sc.k[w]; bi = sc.i[w]; }
    }
    s1 = 0.0;
    s2 = 0.0;
#pragma unroll
    for (int w = 0; w < nw; ++w) {
        const float r = ex2_approx((sc.m[w] - M) * kLog2e);
        s1 = fma(sc.a[w], static_cast<double>(r), s1);
        s2 = fma(sc.b[w], static_cast<double>(r * r), s2);
    }
    best_v = float_from_order_key(bk);
    best_i = bi;
}

// All columns -> mcl_stats (sums, ESS, MAP index and pose).  One 256-thread CTA per trial; runs with a cold
// instruction cache (once per step on one SM), hence small: rolled loops, no libm -- log-normaliser, mean and
// covariance are derived from the sums where they are consumed (derive_stats).
template <int NT>
__device__ __noinline__ void merge_full(const double *pp, int nb, mcl_stats *st, const float *ex, const float *ey,
                                        const float *eth, unsigned long long *dbg, double n_global, bool apply_rule) {
    constexpr int NW = kStepThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nw = NT / 32;
    __shared__ double s_part[NW][kPart];
    __shared__ float s_mx[NW];
    __shared__ uint32_t s_bk[NW], s_bi[NW];
    const int col = tid & (kPart - 1), row0 = tid >> 4;  // 16 threads read one 128-byte record
    constexpr int kRows = NT / kPart;  // 16 threads read one 128-byte record
    dbg_stamp(dbg, 7);
    // pass 1: the global maximum of the partial maxima (loads in explicit batches: one L2 round trip each)
    constexpr int kBatch = 8;
    float mx = -3.4028234e38f;
    for (int b0 = row0; b0 < nb; b0 += kRows * kBatch) {
        double m8[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) m8[u] = __ldcg(pp + min(b0 + u * kRows, nb - 1) * kPart);  // clamped: no branch
#pragma unroll
        for (int u = 0; u < kBatch; ++u) mx = fmaxf(mx, static_cast<float>(m8[u]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_mx[warp] = mx;
    __syncthreads();
    float MM = s_mx[0];
#pragma unroll
    for (int w = 1; w < nw; ++w) MM = fmaxf(MM, s_mx[w]);
    dbg_stamp(dbg, 8);
    // pass 2: column `col` of every record, rescaled to MM (columns 1..13 are sums; column 2 carries squares)
    double acc = 0.0;
    uint32_t bkey = 0u, bidx = 0xffffffffu;
    for (int bb = 0; bb < nb; bb += kRows * kBatch) {  // CTA-uniform trip count: the loop body shuffles
        const int b0 = bb + row0;
        double m8[kBatch], v8[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {  // clamped rows, masked below: the 16 loads of a batch are independent
            const int b = min(b0 + u * kRows, nb - 1);
            m8[u] = __ldcg(pp + b * kPart);
            v8[u] = __ldcg(pp + b * kPart + col);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const bool ok = b0 + u * kRows < nb;
            const float r = ok ? ex2_approx((static_cast<float>(m8[u]) - MM) * kLog2e) : 0.0f;  // empty CTA: sentinel -> 0
            acc = fma(v8[u], static_cast<double>((col == 2) ? r * r : r), acc);
            // arg-max: column 14 holds the value, column 15 (the next lane) the index; largest value, then
            // smallest index (np.argmax, particle_filter.py:116)
            const double idx_d = __shfl_down_sync(0xffffffffu, v8[u], 1);
            if (col == 14 && ok) {
                const uint32_t k = float_order_key(static_cast<float>(v8[u]));
                const uint32_t i = static_cast<uint32_t>(idx_d);
                if (k > bkey || (k == bkey && i < bidx)) { bkey = k; bidx = i; }
            }
        }
    }
    dbg_stamp(dbg, 9);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);  // lanes l and l + 16 hold the same column
    if (lane < kPart) s_part[warp][lane] = acc;
    {
        const uint32_t wk = __reduce_max_sync(0xffffffffu, bkey);
        const uint32_t wi = __reduce_min_sync(0xffffffffu, (bkey == wk) ? bidx : 0xffffffffu);
        if (lane == 0) { s_bk[warp] = wk; s_bi[warp] = wi; }
    }
    __syncthreads();
    if (tid >= 1 && tid <= kNumSums) {  // thread c sums column c over the warps
        double tot = 0.0;
    #pragma unroll
    for (int w = 0; w < nw; ++w) tot += s_part[w][tid];
        if (tid == 1) st->sum_w = tot;
        else if (tid == 2) st->sum_w2 = tot;
        else st->raw[tid - 3] = tot;  // dx dy dth sin cos xx xy yy xth yth thth
        if (tid <= 2) s_part[0][tid] = tot;
    }
    __syncthreads();
    dbg_stamp(dbg, 10);
    if (tid == 0) {
        uint32_t bk = s_bk[0], bj = s_bi[0];
    #pragma unroll
    for (int w = 1; w < nw; ++w)
            if (s_bk[w] > bk || (s_bk[w] == bk && s_bi[w] < bj)) { bk = s_bk[w]; bj = s_bi[w]; }
        const double s1 = s_part[0][1], s2 = s_part[0][2];
        // particle_filter.py:234-237 (a shard of a sharded filter leaves the decision to the merge of all shards)
        const bool uf = apply_rule && weights_underflowed(static_cast<double>(MM), s1, st->abs_shift);
        if (uf) {
            bj = 0u;  // np.argmax of all-equal weights (:116)
            st->sum_w = n_global;
            st->sum_w2 = n_global;
        }
        st->underflow = uf ? 1 : 0;
        st->max_logw = uf ? 0.0 : static_cast<double>(MM);
        st->ess = uf ? n_global : s1 * s1 / s2;  // 1 / sum (w/s1)^2, particle_filter.py:210
        st->raw[11] = uf ? 0.0 : static_cast<double>(float_from_order_key(bk));  // stored log-weight of the MAP particle
        st->max_idx = bj;
        st->est[0] = __ldcg(ex + bj);
        st->est[1] = __ldcg(ey + bj);
        st->est[2] = __ldcg(eth + bj);
        st->derived = 0;  // log_norm, max_weight, mean, cov: derive_stats()
        dbg_stamp(dbg, 6);
    }
}

// on-demand merge (mcl_get_stats and friends): one CTA per trial
struct MergeParams {
    const double *partials;
    int bpt;
    mcl_stats *stats;
    StateView cur;
    long long ld;
    double n_global;
    int apply_rule;
    // single filters read back through mcl_step_async / mcl_get_stats_begin: the block also lands in pinned host memory
    // mapped into the device (no copy-engine operation behind the kernel), then the flag takes the read-back's number
    uint32_t *host_out;
    unsigned int *host_flag;
    unsigned int host_seq;
};
__global__ void __launch_bounds__(kStepThreads) k_merge(const MergeParams p) {
    cudaGridDependencySynchronize();  // launched programmatically behind k_step: its launch latency is hidden
    const int t = blockIdx.x;
    const long long off = static_cast<long long>(t) * p.ld;
    merge_full<kStepThreads>(p.partials + static_cast<long long>(t) * p.bpt * kPart, p.bpt, p.stats + t, p.cur.x + off, p.cur.y + off,
               p.cur.th + off, nullptr, p.n_global, p.apply_rule != 0);
    if (p.host_out != nullptr) {  // (gridDim.x == 1)
        __syncthreads();
        constexpr int kWords = static_cast<int>(sizeof(mcl_stats) / sizeof(uint32_t));
        static_assert(kWords <= kStepThreads && sizeof(mcl_stats) % sizeof(uint32_t) == 0, "one word per thread");
        if (threadIdx.x < kWords) p.host_out[threadIdx.x] = __ldcg(reinterpret_cast<const uint32_t *>(p.stats) + threadIdx.x);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) st_release_sys(p.host_flag, p.host_seq);
    }
}

// -------------------------------------------------------------------------------------------------
// k_scan : per-step scalars + single-pass chained scan (decoupled look-back) over the fixed-point
// weights + ancestors.
// Tile descriptor = status (2 bits) << 62 | value, one 64-bit word, so status and value are read
// atomically.  Tiles are handed out by an atomic ticket so a tile's predecessors are always running.
// The look-back is block-wide (256 predecessors per round): when every tile of a 2^20-particle
// filter starts in the same wave, the prefix reaches the last tile in two rounds instead of sixteen.
// Descriptors are double-buffered: this launch clears the buffer the NEXT scan will use.
//
// Ancestors: every source's slot range [count_le(K_{j-1}), count_le(K_j)) comes from a branch-free
// fp64 evaluation; the block then fills ancestors by slot -- owners mark their first slot in a
// shared-memory window, an inclusive max-scan spreads the marks, stores are coalesced.
// -------------------------------------------------------------------------------------------------
struct ScanParams {
    PrepParams prep;
    unsigned long long *dbg;    // optional phase timeline [T][tpt + 1][kDbgSlots]
    int do_prep;                // 1: an extra block of each trial's row (the last one, or block 1: extras_front) only writes PrepOut
    int early_merge;            // one-wave scan of a single filter: the merging CTAs wait for rec_count, not for the grid
    int extras_front;           // the extra CTAs are blocks 1.. of the row (dispatched first, resident earliest), not the last
    unsigned int *prep_flag;    // [T] the scalars CTA releases this launch's epoch here once PrepOut and the table are written
    unsigned int prep_epoch;
    const unsigned int *rec_count;  // [T] producer CTAs finished so far
    unsigned int rec_target;    // ... and the count at which the producer of the partials is done
    const TrigEntry *trig;
    const float *lw;            // [T][ld]
    mcl_stats *stats;           // [T] statistics of the stored weights (weight map, ESS, estimate)
    // n_partials > 0: the statistics are still the CTA partials of the previous k_step / k_reduce: tiles merge
    // (max, sum, sum^2), the scalars CTA the arg-max, and one more extra CTA merges everything into stats[]
    const double *partials;     // [T][n_partials][kPart]
    int n_partials;
    uint4 *wm_pub;              // [T][2] {weight-map m, scale, resample gate, epoch}, {comb offset, epoch}: published by tile 0
    unsigned int epoch;         // scan launches of this handle so far (never 0: the words start as 0)
    StateView cur;              // the stored state (MAP pose lookup)
    unsigned long long *desc;   // [T][tpt] tile descriptors of this launch (all zero on entry)
    unsigned long long *desc_next;  // [T][tpt] cleared for the next launch (may be null)
    unsigned int *ready;        // [T][32] bit per tile of the direct range: aggregate published (zero on entry)
    unsigned int *ready_next;   // cleared for the next launch (may be null)
    unsigned int *tickets;      // [T][4]  (slot 0 = scan ticket)
    uint32_t *anc;              // [T][ld] (slot - slot_first)
    unsigned long long *total;  // optional [T] : inclusive total of the local weights
    // kScanExpand after mcl_shard_total: exclusive prefix of every tile's weight (that pass summed the tiles
    // anyway) -- tiles then need no look-back and no descriptors at all
    const unsigned long long *tile_prefix;  // [tpt] weight before the tile inside its segment
    const unsigned long long *seg_prefix;   // [segments] weight before the segment
    int tiles_per_seg;
    long long n, ld, n_global;  // local sources, stride, comb length
    long long j_global_base;    // global index of local source 0 (sharded)
    long long slot_first;       // anc is indexed by (global slot - slot_first)
    long long anc_cap;          // capacity of anc per trial
    unsigned long long k_offset;  // sum of the quantised weights of all lower-ranked shards
    double delta;               // fl(1 / n_global)
    int tpt;
    int use_override;           // 1: always resample, ofs from ofs_override (sharded resampling)
    int persistent;             // 1: fewer tile CTAs than tiles: a CTA keeps taking tickets
    double ofs_override;
    // kScanExpand (sharded filter): slots inside this shard's own output range [own_lo, own_hi) keep the
    // ancestor path (anc[slot - own_lo]); slots owned by other ranks get the resampled POSE written to
    // the send buffers: slots below own_lo at (slot - slot_first), slots from own_hi on at
    // send_hi_off + (slot - max(own_hi, slot_first)); anc_cap = capacity of the send buffers.
    long long own_lo, own_hi, send_hi_off, send_hi_base;
    StateView src;
    float *ox, *oy, *oth;
    // kScanExpand, device-driven (mcl_shard_step_peer): gate, comb offset and k_offset come from the plan the
    // totals exchange left in device memory, and off-rank poses go straight into their owner's receive staging
    const ShardPlan *plan;
    const PeerView *pv;
    uint32_t xflags;            // host flags of the step (MCL_FORCE_RESAMPLE / MCL_NO_RESAMPLE)
    int no_dep_sync;            // launched programmatically behind a kernel whose results it does not need
};

constexpr int kDirectTiles = 1024;  // tiles that sum all predecessors' aggregates directly (2M particles)
constexpr int kScanAncestors = 0;  // write the ancestor of every slot
constexpr int kScanExpand = 2;     // write the resampled pose of every slot (sender side of the sharded filter)

enum : unsigned long long {
    kDescAggregate = 1ull << 62,
    kDescPrefix = 2ull << 62,
    kDescValueMask = (1ull << 62) - 1ull
};

struct ScanShared {
    double ofs;
    float wq_m, wq_scale;
    int resample;
    int uniform;  // the stored weights underflowed in the reference's arithmetic (particle_filter.py:234-237): all-equal
    unsigned int tile;
};

// kScanExpand: where the resampled particle of one global slot goes (see ScanParams::own_lo)
__device__ __forceinline__ void expand_store(const ScanParams &p, uint32_t *anc, long long slot, uint32_t j) {
    if (slot >= p.own_lo && slot < p.own_hi) {
        anc[slot - p.own_lo] = j;
    } else if (p.plan != nullptr) {
        // the slot's owner and its index there (off-rank slots are ~1e-4 of all: the division is off the hot path)
        const uint32_t m = static_cast<uint32_t>(p.pv->m), s32 = static_cast<uint32_t>(slot);
        const uint32_t g = s32 / m, l = s32 - g * m;
        p.pv->rx[g][l] = p.src.x[j];
        p.pv->ry[g][l] = p.src.y[j];
        p.pv->rth[g][l] = p.src.th[j];
    } else {
        const long long o = (slot < p.own_lo) ? slot - p.slot_first : p.send_hi_off + (slot - p.send_hi_base);
        if (o >= 0 && o < p.anc_cap) {
            p.ox[o] = p.src.x[j];
            p.oy[o] = p.src.y[j];
            p.oth[o] = p.src.th[j];
        }
    }
}

// ITEMS sources per thread, 2048 / ITEMS threads per tile.  16: eight tiles resident per SM -- the geometry of large
// scans, which are bound by the latency of a tile at a fixed number of resident tiles (0.64 -> 0.49 ms at 125 M).
// 8: four tiles of 256 threads -- shorter per-tile latency, better while all tiles of a filter fit in ONE wave (C1).
#ifndef MCL_SCAN16_CTAS
#define MCL_SCAN16_CTAS 8
#endif
constexpr int scan_ctas_per_sm(int items) { return items == 16 ? MCL_SCAN16_CTAS : 4; }
template <int MODE, int ITEMS>
__global__ void __launch_bounds__(kScanTile / ITEMS, scan_ctas_per_sm(ITEMS)) k_scan(const __grid_constant__ ScanParams p) {
    // (the names of the namespace-scope defaults, re-bound to this instantiation's geometry)
    constexpr int kScanItems = ITEMS, kScanThreads = kScanTile / ITEMS, kScanWin = 32 * ITEMS;
    const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kScanThreads / 32;
    __shared__ ScanShared s_sc;
    __shared__ unsigned long long s_warp[NW];
    __shared__ int s_first[NW];
    __shared__ uint32_t s_slot[2];
    __shared__ __align__(16) uint32_t s_mark[kScanTile];  // also the reduction scratch of prep_block
    __shared__ uint32_t s_wmax[NW];

    // Programmatic dependent launch: the next kernel of the stream may be scheduled as soon as every CTA of
    // this grid has started; nothing is read or written before the previous kernel has completed.
    cudaTriggerProgrammaticLaunchCompletion();
    dbg_stamp(p.dbg, 11);  // entry, before any dependency is resolved (tools/timeline.py --back-to-back)
    // (the expansion of the device-driven sharded step is launched programmatically behind the per-step scalars
    //  kernel, on which it does not depend: it runs beside it)
    // per-step scalars are the job of one extra CTA per trial: they are only needed by k_step, and their
    // once-per-step code (cold instruction cache) must not delay a tile that every later tile waits for
    // (with the early dependency below the extra CTAs are blocks 1 and 2 of the row, so that they are dispatched -- and
    //  resident -- as early as tile 0; `lin` is a CTA's index among the tile CTAs)
    const int n_extra_ctas = p.do_prep ? (p.n_partials > 0 ? 2 : 1) : 0;
    const bool front = p.extras_front != 0;
    const bool prep_cta = p.do_prep && blockIdx.x == (front ? 1u : gridDim.x - 1);
    // ... and another extra CTA merges ALL columns of the partials into mcl_stats (host read-back, k_cov)
    const bool stats_cta = p.do_prep && p.n_partials > 0 && blockIdx.x == (front ? 2u : gridDim.x - 2);
    const unsigned int lin = (front && blockIdx.x != 0u) ? blockIdx.x - static_cast<unsigned int>(n_extra_ctas) : blockIdx.x;
    __shared__ MergeScratch s_merge;
    const long long soff = static_cast<long long>(t) * p.ld;
    const double *pp = p.partials + static_cast<long long>(t) * p.n_partials * kPart;
    // One-wave scan of a single filter.  Under programmatic dependent launch this grid's CTAs become resident as the
    // producer's CTAs retire -- tile 0 some 12 us before the producer grid has completed -- and then sit in
    // cudaGridDependencySynchronize; only after it did tile 0 start to merge the producer's records (cold: ~4 us) while
    // every other tile waited for its publication.  For tile 0 and the scalars CTA the dependency is therefore the
    // count of finished producer CTAs: each producer CTA releases all its stores when it counts itself, so at the
    // target everything this kernel reads is visible and no producer CTA still reads the ancestors.  ONE thread polls
    // (the SM is shared with producer CTAs still at work); tile 0 first runs the merge once on whatever the records
    // hold -- result discarded -- which brings its instructions and the records' lines in.
    const bool early_cta = p.early_merge && p.n_partials > 0 && !p.no_dep_sync && (prep_cta || (blockIdx.x == 0 && !p.persistent));
    if (early_cta) {
        __shared__ int s_done;
        if (tid == 0) s_done = (ld_acquire_gpu(p.rec_count + t) == p.rec_target) ? 1 : 0;
        __syncthreads();
        if (!s_done) {  // (CTA-uniform) there is time to spend
            float m_, v_;
            double a_, b_;
            uint32_t i_;
            if (prep_cta) merge_light_best<kScanThreads>(pp, p.n_partials, s_merge, m_, a_, b_, v_, i_);
            else merge_light_raw<kScanThreads>(pp, p.n_partials, s_merge, m_, a_, b_);
            if (tid == 0) {
                // (bounded: a count that never arrives falls back to the dependency every other CTA uses)
                const unsigned long long t0 = global_timer_ns();
                int ok = 1;
                while (ld_acquire_gpu(p.rec_count + t) != p.rec_target) {
                    __nanosleep(200);
                    if (global_timer_ns() - t0 > kEarlyWaitNs) { ok = 0; break; }
                }
                s_done = ok;
            }
            __syncthreads();
            if (!s_done) cudaGridDependencySynchronize();
        }
    } else if (!p.no_dep_sync) {
        cudaGridDependencySynchronize();
    }
    dbg_stamp(p.dbg, 0);
    if (stats_cta) {
        merge_full<kScanThreads>(pp, p.n_partials, p.stats + t, p.cur.x + soff, p.cur.y + soff, p.cur.th + soff, nullptr, p.prep.n_global,
                   p.n_global == p.n);
        return;
    }
    const mcl_stats *stp = p.stats + t;
    // One tile per CTA (single filters): tiles are numbered by blockIdx.x.  CTAs are dispatched in that order, so a
    // tile's predecessors are always running (the assumption of every single-pass scan) and no ticket round trip
    // precedes the first load.  Persistent CTAs (trial batches) take tickets.
    // (tiles whose prefixes are known -- after a totals pass -- do not depend on each other; striding over them
    //  statically instead of by ticket was measured 10 % slower: tiles differ in work, tickets balance it)
    const bool ordered = !p.persistent;
    // When the statistics still live in the CTA partials, tile 0 (certainly running) merges (max, sum, sum^2)
    // and publishes what every tile needs -- weight map, resample gate, comb offset -- as two aligned 16-byte
    // words tagged with this launch's epoch; the others poll them (one round trip) -- 500 CTAs each reading
    // all partials would hammer the same few L2 lines.
    uint4 *pub = p.wm_pub + 2 * t;  // word A {m, scale, gate, epoch}, word B {ofs lo, ofs hi, epoch, 0}
    if (tid == 0) {
        const bool may_poll = p.n_partials > 0 && !prep_cta;
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
        if (may_poll) {  // speculative first poll, in flight together with the ticket
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(pub));
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(pub + 1));
        }
        const unsigned int tk = prep_cta ? 0xffffffffu : (ordered ? lin : atomicAdd(&p.tickets[t * 4 + 0], 1u));
        s_sc.tile = tk;
        if (may_poll && tk != 0u) {
            while (a.w != p.epoch || b.z != p.epoch) {
                __nanosleep(32);
                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(pub));
                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(pub + 1));
            }
            s_sc.wq_m = __uint_as_float(a.x);
            s_sc.wq_scale = __uint_as_float(a.y);
            s_sc.resample = static_cast<int>(a.z);
            s_sc.uniform = static_cast<int>(b.w);
            s_sc.ofs = __hiloint2double(static_cast<int>(b.y), static_cast<int>(b.x));
        }
    }
    __syncthreads();
    const bool merger = p.n_partials > 0 && (prep_cta || s_sc.tile == 0u);
    double st_max = 0.0, st_sum = 1.0;
    float pre_bv = 0.f;
    uint32_t pre_bi = 0u;
    if (merger || p.n_partials == 0) {  // CTA-uniform: this CTA derives the scalars itself
        double st_ess;
        int st_uniform;
        if (merger) {
            float M;
            double s1, s2;
            // (a shard of a sharded filter never merges here: its statistics exchange has set n_partials = 0)
            if (prep_cta) {  // the scalars CTA also needs the arg-max: one pass over the records for both
                merge_light_best<kScanThreads>(pp, p.n_partials, s_merge, M, s1, s2, pre_bv, pre_bi);
                st_uniform = weights_underflowed(static_cast<double>(M), s1, stp->abs_shift) ? 1 : 0;
                if (st_uniform) {
                    M = 0.0f;
                    s1 = s2 = p.prep.n_global;
                }
            } else {
                merge_light<kScanThreads>(pp, p.n_partials, s_merge, M, s1, s2, stp->abs_shift, p.prep.n_global, st_uniform);
            }
            st_max = static_cast<double>(M);
            st_sum = s1;
            st_ess = s1 * s1 / s2;
        } else {
            st_max = stp->max_logw;
            st_sum = stp->sum_w;
            st_ess = stp->ess;
            st_uniform = stp->underflow;
        }
        if (tid == 0) {
            const bool planned = p.plan != nullptr;
            const int rs = planned ? plan_resample(p.xflags, p.plan)
                                   : (p.use_override ? 1 : decide_resample(p.prep.flags, st_ess, p.prep.ess_threshold));
            const WeightMap w0 = make_weight_map(st_max, st_sum);
            const double o = planned ? p.plan->ofs : (p.use_override ? p.ofs_override : (rs ? comb_offset(p.prep, t) : 0.0));
            s_sc.resample = rs;
            s_sc.uniform = st_uniform;
            s_sc.wq_m = w0.m;
            s_sc.wq_scale = w0.scale;
            s_sc.ofs = o;
            if (merger && !prep_cta) {
                const uint4 va = make_uint4(__float_as_uint(w0.m), __float_as_uint(w0.scale), static_cast<unsigned int>(rs), p.epoch);
                const uint4 vb = make_uint4(static_cast<unsigned int>(__double2loint(o)), static_cast<unsigned int>(__double2hiint(o)), p.epoch,
                                            static_cast<unsigned int>(st_uniform));
                asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(pub), "r"(va.x), "r"(va.y), "r"(va.z), "r"(va.w) : "memory");
                asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(pub + 1), "r"(vb.x), "r"(vb.y), "r"(vb.z), "r"(vb.w) : "memory");
            }
        }
        __syncthreads();
    }
    const int resample = s_sc.resample;
    const int uniform = s_sc.uniform;
    const double ofs = s_sc.ofs;
    WeightMap wm;
    wm.m = s_sc.wq_m;
    wm.scale = s_sc.wq_scale;
    dbg_stamp(p.dbg, 1);
    if (prep_cta) {
        if (p.ready_next != nullptr && tid < kDirectTiles / 32) p.ready_next[t * (kDirectTiles / 32) + tid] = 0u;
        PrepIn in;
        in.max_logw = st_max;
        in.sum_w = st_sum;
        in.uniform = uniform;
        if (p.n_partials > 0) {
            uint32_t bi = pre_bi;
            if (uniform) bi = 0u;  // np.argmax of all-equal weights (:116)
            in.est[0] = p.cur.x[soff + bi];
            in.est[1] = p.cur.y[soff + bi];
            in.est[2] = p.cur.th[soff + bi];
        } else {
            in.est[0] = stp->est[0];
            in.est[1] = stp->est[1];
            in.est[2] = stp->est[2];
        }
        prep_block(p.prep, in, p.trig, t, resample, ofs, reinterpret_cast<double *>(s_mark));
        dbg_stamp(p.dbg, 6);
        // k_step's CTAs (resident early under programmatic launch) read the scalars and stage the table before the
        // rest of this grid has finished: barrier, then one release that is cumulative over the CTA's stores
        __syncthreads();
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.prep_flag + t), "r"(p.prep_epoch) : "memory");
        return;
    }
    // The descriptors the NEXT scan launch starts from are cleared whether or not this trial resamples: a launch
    // that skipped the clear would leave the previous resampling's aggregates / prefixes in the buffer the launch
    // after it looks back through (tiles past the direct range accept any non-zero status at once).
    if (p.desc_next != nullptr && static_cast<int>(lin) * kScanThreads < p.tpt) {
        const int n_tile_ctas = static_cast<int>(gridDim.x) - n_extra_ctas;
        unsigned long long *dn = p.desc_next + static_cast<long long>(t) * p.tpt;
        for (long long i = static_cast<long long>(lin) * kScanThreads + tid; i < p.tpt;
             i += static_cast<long long>(n_tile_ctas) * kScanThreads)
            dn[i] = 0ull;
    }
    if (!resample) return;
    const unsigned long long k_offset = (MODE == kScanExpand && p.plan != nullptr) ? p.plan->k_offset : p.k_offset;
    const float *lw = p.lw + static_cast<long long>(t) * p.ld;
    // Persistent over tiles: a CTA keeps taking tickets, so the prologue above is paid once per CTA.
    __shared__ unsigned int s_next[2];  // two slots by iteration parity: one barrier per tile is enough
    // (requesting the next tile's weights a tile ahead -- cp.async into a double buffer -- was measured: no gain for
    //  the expansion, and its 16 KB of shared memory cost the look-back scan 6 %)
    const int n_tile_ctas_s = static_cast<int>(gridDim.x) - n_extra_ctas;
    unsigned long long solo_prefix = 0ull;
    int it = 0;
    for (long long tile = s_sc.tile; tile < p.tpt; tile = static_cast<long long>(s_next[it & 1]), ++it) {
    const long long j0 = tile * kScanTile + static_cast<long long>(tid) * kScanItems;

    // ---- load + quantise (16 consecutive items per thread, four 16-byte loads) ----
    // The integers rint(w 2^52) are kept as fp32 (exact: rintf of an fp32 value) -- 16 registers instead of 32 --
    // and widened where they are summed.
    float qf[kScanItems];
    if (j0 < p.n) {
        // the state arrays are padded to a multiple of 128 elements, so the vector loads stay in bounds
        float v[kScanItems];
#pragma unroll
        for (int c = 0; c < kScanItems / 4; ++c) {
            float4 a = *reinterpret_cast<const float4 *>(lw + j0 + 4 * c);
            if (uniform) a = make_float4(0.f, 0.f, 0.f, 0.f);  // CTA-uniform: weights reset to 1/NP (:234-237)
            v[4 * c] = a.x; v[4 * c + 1] = a.y; v[4 * c + 2] = a.z; v[4 * c + 3] = a.w;
        }
        if (j0 + kScanItems <= p.n) {
#pragma unroll
            for (int k = 0; k < kScanItems; ++k) qf[k] = rintf(weight_scaled(v[k], wm));
        } else {
#pragma unroll
            for (int k = 0; k < kScanItems; ++k) qf[k] = (j0 + k < p.n) ? rintf(weight_scaled(v[k], wm)) : 0.0f;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) qf[k] = 0.0f;
    }
    // the first mark window is cleared here, off the critical path: the barriers of the block scan order it
#pragma unroll
    for (int c = 0; c < kScanItems / 4; ++c)  // the thread's own 64 bytes (= its lane's strip of the warp's window)
        *reinterpret_cast<uint4 *>(&s_mark[tid * kScanItems + 4 * c]) = make_uint4(0u, 0u, 0u, 0u);
    double tsum_d = 0.0;  // sums of integers below 2^53 are exact in fp64
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) tsum_d += static_cast<double>(qf[k]);
    const unsigned long long tsum = static_cast<unsigned long long>(tsum_d);

    // ---- block exclusive scan of the thread sums ----
    unsigned long long incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_off = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const unsigned long long v = s_warp[w];
        if (w < warp) warp_off += v;
        tile_total += v;
    }
    const unsigned long long thread_excl = warp_off + incl - tsum;
    __syncthreads();  // s_warp is reused by the look-back

    dbg_stamp(p.dbg, 2);
    // ---- decoupled look-back, block-wide ----
    unsigned long long *desc = p.desc + static_cast<long long>(t) * p.tpt;
    unsigned long long excl = 0;
    // tiles of the direct range also raise a bit in ready[tile / 32] once their aggregate is visible
    unsigned int *rdy = p.ready + t * (kDirectTiles / 32);
    const bool have_prefix = (MODE == kScanExpand) && p.tile_prefix != nullptr;
    // A trial whose tiles are ALL taken by this one persistent CTA (trial batches: thousands of small filters) needs
    // no descriptors and no look-back: the tickets come in order, so the prefix is simply carried in a register.
    const bool solo = p.persistent && n_tile_ctas_s == 1;
    if (have_prefix) {
        excl = __ldg(p.seg_prefix + tile / p.tiles_per_seg) + __ldg(p.tile_prefix + tile);
    } else if (solo) {
        excl = solo_prefix;
        solo_prefix += tile_total;
    } else if (tile == 0) {
        if (tid == 0) {
            atomicExch(&desc[0], kDescAggregate | tile_total);  // == its prefix; see the direct range
            __threadfence();
            atomicOr(&rdy[0], 1u);
        }
    } else {
        if (tid == 0) {
            atomicExch(&desc[tile], kDescAggregate | tile_total);
            if (tile < kDirectTiles) {
                __threadfence();
                atomicOr(&rdy[tile >> 5], 1u << (tile & 31));
            }
        }
        long long look = tile - 1;  // window covers tiles [look-255, look]
        if (tile <= kDirectTiles) {
            // Small scans start all their tiles in one wave: waiting for a predecessor's PREFIX would
            // serialise them in rounds of 256.  Every aggregate is published at about the same time, so sum
            // ALL predecessors' aggregates instead (<= 4 independent loads per thread).  Waiting is done on the
            // ready bit masks -- one warp polls <= 32 words -- because 500 CTAs polling 4 KB of descriptors
            // each is what made this phase slow; the descriptors are then read exactly once.
            if (warp == 0) {
                const int gi = static_cast<int>(tile >> 5);  // lane g watches group g (g <= gi)
                const unsigned int need = (lane < gi) ? 0xffffffffu : ((lane == gi) ? ((1u << (tile & 31)) - 1u) : 0u);
                while (true) {
                    const unsigned int m = need ? *reinterpret_cast<volatile unsigned int *>(&rdy[lane]) : 0u;
                    if (__all_sync(0xffffffffu, (m & need) == need)) break;
                    __nanosleep(64);
                }
                __threadfence();
            }
            __syncthreads();
            unsigned long long d[kDirectTiles / kScanThreads];
#pragma unroll
            for (int r = 0; r < kDirectTiles / kScanThreads; ++r) {
                const long long idx = look - tid - r * kScanThreads;
                d[r] = (idx >= 0) ? *reinterpret_cast<volatile unsigned long long *>(&desc[idx]) : kDescPrefix;
            }
            unsigned long long contrib = 0;
#pragma unroll
            for (int r = 0; r < kDirectTiles / kScanThreads; ++r) {
                // an aggregate (status 1) or a prefix (status 2) of tile idx both END with idx's own total only
                // if we use aggregates; prefixes are converted back below
                contrib += d[r] & kDescValueMask;
            }
            // NOTE: tiles publish their aggregate first and overwrite it with the inclusive prefix later; a
            // prefix read here would double count.  Tiles in the direct range therefore never publish prefixes
            // (see below): every descriptor read above is an aggregate.
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            if (lane == 0) s_warp[warp] = contrib;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < NW; ++w) excl += s_warp[w];
            __syncthreads();
        } else
        // (measured and dropped: a round with two barriers instead of four, the 64-bit shuffle tree replaced by two
        //  hardware REDUX on 27-bit halves -- 15 % fewer look-back instructions, 0.5 % SLOWER: the scan is bound by
        //  the latency of its dependent steps, not by instruction issue)
        while (true) {
            // one round = the 256 descriptors before `look`, two per thread: position tid (nearest first) and tid + 128
            const long long idx0 = look - tid, idx1 = look - kScanThreads - tid;
            unsigned long long d0 = kDescPrefix, d1 = kDescPrefix;  // tiles before the first one: an empty prefix
            if (idx0 >= 0) d0 = *reinterpret_cast<volatile unsigned long long *>(&desc[idx0]);
            if (idx1 >= 0) d1 = *reinterpret_cast<volatile unsigned long long *>(&desc[idx1]);
            while (!__syncthreads_and((d0 >> 62) != 0ull && (d1 >> 62) != 0ull)) {
                if ((d0 >> 62) == 0ull) d0 = *reinterpret_cast<volatile unsigned long long *>(&desc[idx0]);
                if ((d1 >> 62) == 0ull) d1 = *reinterpret_cast<volatile unsigned long long *>(&desc[idx1]);
            }
            const unsigned int pre0 = __ballot_sync(0xffffffffu, (d0 >> 62) == 2ull);
            const unsigned int pre1 = __ballot_sync(0xffffffffu, (d1 >> 62) == 2ull);
            if (lane == 0)
                s_first[warp] = pre0 ? (warp * 32 + __ffs(pre0) - 1)
                                     : (pre1 ? (kScanThreads + warp * 32 + __ffs(pre1) - 1) : 0x7fffffff);
            __syncthreads();
            int first = 0x7fffffff;  // nearest predecessor (smallest position) that already holds a prefix
#pragma unroll
            for (int w = 0; w < NW; ++w) first = min(first, s_first[w]);
            unsigned long long contrib = ((tid <= first && idx0 >= 0) ? (d0 & kDescValueMask) : 0ull) +
                                         ((tid + kScanThreads <= first && idx1 >= 0) ? (d1 & kDescValueMask) : 0ull);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            if (lane == 0) s_warp[warp] = contrib;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < NW; ++w) excl += s_warp[w];
            __syncthreads();
            if (first != 0x7fffffff) break;
            look -= 2 * kScanThreads;
        }
        // tiles of the direct range keep their aggregate (later direct tiles sum aggregates); the LAST direct
        // tile and everything after it publish inclusive prefixes for the classic look-back
        if (tid == 0 && tile >= kDirectTiles) atomicExch(&desc[tile], kDescPrefix | (excl + tile_total));
    }
    if (tid == 0 && p.total != nullptr && tile == p.tpt - 1) p.total[t] = excl + tile_total;
    // The next ticket is taken only now: a ticket held by a CTA that has not started the tile would stall
    // every later tile's look-back.  Its round trip overlaps the ancestor phase below.
    if (tid == 0) s_next[it & 1] = p.persistent ? atomicAdd(&p.tickets[t * 4 + 0], 1u) : 0xffffffffu;
    dbg_stamp(p.dbg, 3);
    // ---- slot range of every source: u32 counts, non-owners carry rel = 2^31 (never inside a window) ----
    const double nd = static_cast<double>(p.n_global);
    const CombFast cf = make_comb_fast(ofs, nd);
    const double Kd0 = static_cast<double>(k_offset + excl + thread_excl);
    uint32_t rel[kScanItems];  // first slot of source k relative to the thread's first slot
    uint32_t first_slot, cur;
    // tiles that are full and do not hold the filter's last particle need no per-item checks (CTA-uniform)
    const bool plain_tile = (tile + 1) * kScanTile <= p.n && (tile + 1) * kScanTile < p.n_global - p.j_global_base;
    bool exact = !plain_tile;
    if (plain_tile) {
        const double a0m = cf.a0 + kCombMagic;
        uint32_t ok = 0xffffffffu;
        double Kd = Kd0;
        first_slot = count_le_bits(Kd, cf.a1, a0m, ok);
        cur = first_slot;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            Kd += static_cast<double>(qf[k]);
            const uint32_t end = count_le_bits(Kd, cf.a1, a0m, ok);
            rel[k] = (end != cur) ? cur - first_slot : 0x80000000u;
            cur = end;
        }
        // a count too close to a tooth, or past the end of the comb: this thread redoes its sources exactly
        exact = (ok == 0u) || cur > static_cast<uint32_t>(p.n_global);
    }
    if (exact) {
        // the last particle of the whole filter also takes every slot past the final cumulative weight
        // (the reference would raise IndexError there)
        const long long k_last64 = p.n_global - 1 - p.j_global_base - j0;
        const int k_last = (k_last64 >= 0 && k_last64 < kScanItems) ? static_cast<int>(k_last64) : -1;
        const int n_valid = static_cast<int>(min(static_cast<long long>(kScanItems), max(0ll, p.n - j0)));
        double Kd = Kd0;
        // the filter's very first source owns the slots from 0 on: with ofs == 0 the tooth u_0 = 0 is "<= K" for
        // K = 0, and the reference walk (strict >, particle_filter.py:219) leaves slot 0 with particle 0
        const double fs = (p.j_global_base + j0 == 0) ? 0.0 : count_le_fast(Kd, cf, p.delta, ofs);
        double c = fs;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            Kd += static_cast<double>(qf[k]);
            double end = count_le_fast(Kd, cf, p.delta, ofs);
            if (k == k_last) end = nd;
            if (k < n_valid && end > c) {
                rel[k] = static_cast<uint32_t>(static_cast<long long>(c - fs));
                c = end;
            } else {
                rel[k] = 0x80000000u;
            }
        }
        first_slot = static_cast<uint32_t>(static_cast<long long>(fs));
        cur = static_cast<uint32_t>(static_cast<long long>(c));
    }
    // ---- ancestors ----
    // Common case: every WARP fills the slots of its own 512 sources by itself -- marks, max-scan and stores inside
    // a 2 KB strip of shared memory that only this warp touches, ordered by __syncwarp: no block barrier at all.
    // A warp's slot range is [first_slot of lane 0, cur of its last lane holding sources); slot `first_slot` of any
    // thread is owned by the first source at or after it that owns slots at all, so every slot of the range has a
    // mark at or before it inside the warp (or is covered by the carry of the previous pass).  Marks hold the source
    // index itself; 0 doubles as "empty", which is right because only the very first source can be 0 and it owns
    // whatever precedes the first other mark.  (A strip of 1024 slots per warp, so that one pass of marks nearly
    // always suffices, was measured: k_scan 0.486 -> 0.665 ms at 125 M -- 18 KB of shared memory per CTA and the
    // lanes' strips 128 bytes apart.)  Tiles whose sources own far more slots than sources (collapsed
    // weights: one particle owning up to every slot) keep the block-cooperative fill below.
    const uint32_t anc_lo = (MODE == kScanAncestors) ? 0u : static_cast<uint32_t>(p.own_lo);
    uint32_t *anc = p.anc + static_cast<long long>(t) * p.ld;
    const long long n_in_tile = min(static_cast<long long>(kScanTile), p.n - tile * kScanTile);
    const int last_tid = static_cast<int>((n_in_tile - 1) / kScanItems);
    // eight consecutive slots from registers: one 32-byte store when the chunk lies inside [lo, hi) and in this
    // handle's ancestor array, else slot by slot (the two partial chunks of a range; off-rank slots of an expansion)
    auto store8 = [&](uint32_t slot0, const uint32_t (&v)[8], uint32_t lo, uint32_t hi, bool reg_store) {
        if (reg_store) {
            if (slot0 >= lo && slot0 + 8u <= hi) {
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(anc + (slot0 - anc_lo)), "r"(v[0]),
                             "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                             : "memory");
            } else {
                // (parking the partial chunks in shared memory and storing them once per warp was measured 6 % SLOWER:
                //  one more dependent shared-memory round trip at the end of every tile)
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (slot0 + k >= lo && slot0 + k < hi) anc[slot0 + k - anc_lo] = v[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (slot0 + k >= lo && slot0 + k < hi) expand_store(p, anc, static_cast<long long>(slot0) + k, v[k]);
        }
    };
    const bool big = static_cast<double>(tile_total) * cf.a1 > 3.0 * kScanTile;  // CTA-uniform
    if (!big) {
        uint32_t *wm = s_mark + warp * kScanWin;
        // (the range ends with the last lane that HOLDS sources: padding lanes of the filter's last tile sit at the
        //  final cumulative weight, which may fall short of the slots the last particle was given)
        const int last_lane = min(31, last_tid - warp * 32);
        const uint32_t wf = __shfl_sync(0xffffffffu, first_slot, 0);
        const uint32_t we = (last_lane >= 0) ? __shfl_sync(0xffffffffu, cur, max(last_lane, 0)) : wf;
        const uint32_t mis = (wf - anc_lo) & 7u;   // 32-byte stores: windows start where (slot - anc_lo) is a multiple of 8
        const bool aligned = mis <= wf;
        const uint32_t Wb = aligned ? wf - mis : wf;
        const uint32_t span = (we > wf) ? we - Wb : 0u;
        const uint32_t my_off = first_slot - Wb;
        const uint32_t jsrc0 = static_cast<uint32_t>(tile * kScanTile) + static_cast<uint32_t>(tid * kScanItems);
        uint32_t carry = 0u;
        for (uint32_t cb = 0; cb < span; cb += kScanWin) {
            if (cb > 0u) {  // (the first window was cleared before the block scan)
#pragma unroll
                for (int c = 0; c < kScanItems / 4; ++c)
                    *reinterpret_cast<uint4 *>(&wm[lane * kScanItems + 4 * c]) = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
            if (cur != first_slot) {
                const uint32_t wofs = my_off - cb;  // modulo 2^32 (n_global < 2^31: no aliasing)
#pragma unroll
                for (int k = 0; k < kScanItems; ++k) {
                    const uint32_t pos = wofs + rel[k];
                    if (pos < static_cast<uint32_t>(kScanWin)) wm[pos] = jsrc0 + k;
                }
            }
            __syncwarp();
            const bool reg_store = (MODE == kScanAncestors) ||
                                   (aligned && static_cast<long long>(Wb) + cb >= p.own_lo &&
                                    static_cast<long long>(Wb) + cb + kScanWin <= p.own_hi);
            // the window in chunks of 256 slots: eight consecutive slots per lane and chunk
#pragma unroll
            for (int half = 0; half < kScanWin / 256; ++half) {
                if (cb + half * 256u >= span) break;  // warp-uniform
                uint32_t v[8];
                {
                    const uint4 a = *reinterpret_cast<const uint4 *>(&wm[half * 256 + lane * 8]);
                    const uint4 b = *reinterpret_cast<const uint4 *>(&wm[half * 256 + lane * 8 + 4]);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                }
#pragma unroll
                for (int k = 1; k < 8; ++k) v[k] = max(v[k], v[k - 1]);
                uint32_t run = v[7];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) run = max(run, __shfl_up_sync(0xffffffffu, run, o));  // (lanes < o get their own value)
                uint32_t before = __shfl_up_sync(0xffffffffu, run, 1);
                before = (lane == 0) ? carry : max(before, carry);
                carry = max(carry, __shfl_sync(0xffffffffu, run, 31));
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = max(v[k], before);
                store8(Wb + cb + static_cast<uint32_t>(half * 256 + lane * 8), v, wf, we, reg_store);
            }
        }
    } else {
        // tile slot range [S0, S1): S0 from the first thread, S1 from the last thread holding sources
        if (tid == 0) s_slot[0] = first_slot;
        if (tid == last_tid) s_slot[1] = cur;
        __syncthreads();
        const uint32_t S0 = s_slot[0], S1 = s_slot[1];
        // the ancestor array is written with 32-byte stores: the windows start where (slot - first slot of anc[]) is a
        // multiple of 8 (anc[] starts at slot 0, or at own_lo in the sharded expansion)
        const uint32_t mis = (S0 - anc_lo) & 7u;
        const bool aligned = mis <= S0;
        const uint32_t W0 = aligned ? S0 - mis : S0;
        const uint32_t span = (S1 > S0) ? S1 - W0 : 0u;
        const uint32_t my_off = first_slot - W0;  // threads with sources: >= 0

        dbg_stamp(p.dbg, 4);
        uint32_t carry = 0;  // source (tile-local index + 1) that owns the slot just before the window
        const uint32_t jbase = static_cast<uint32_t>(tile * kScanTile) - 1u;
        for (uint32_t cb = 0; cb < span; cb += kScanTile) {
            // window of kScanTile slots starting at W0 + cb; every thread handles kScanItems consecutive slots of it
            // registers -> anc[] directly when the whole window belongs to this handle's ancestor array (always, unsharded)
            const bool reg_store = (MODE == kScanAncestors) ||
                                   (aligned && static_cast<long long>(W0) + cb >= p.own_lo &&
                                    static_cast<long long>(W0) + cb + kScanTile <= p.own_hi);
            const uint32_t slot0 = W0 + cb + static_cast<uint32_t>(tid * kScanItems);
            if (cb > 0u && span > 4u * kScanTile) {
                // Degenerate weights: one source owns many windows.  A window in which no source starts is a
                // constant fill with the carried owner -- one barrier and two stores per thread instead of a scan.
                bool starts = false;
                const uint32_t wofs = my_off - cb;
#pragma unroll
                for (int k = 0; k < kScanItems; ++k) starts = starts || (wofs + rel[k] < static_cast<uint32_t>(kScanTile));
                if (!__syncthreads_or(starts)) {
                    const uint32_t j = jbase + carry;
                    if (reg_store) {
                        const uint32_t c8[8] = {j, j, j, j, j, j, j, j};
                        store8(slot0, c8, S0, S1, true);
                        store8(slot0 + 8u, c8, S0, S1, true);
                    } else {
                        const int lim = static_cast<int>(min(static_cast<uint32_t>(kScanTile), span - cb));
#pragma unroll
                        for (int k = 0; k < kScanItems; ++k) {
                            const int pos = k * kScanThreads + tid;
                            const long long slot = static_cast<long long>(W0) + cb + pos;
                            if (pos < lim && slot >= S0) expand_store(p, anc, slot, j);
                        }
                    }
                    continue;
                }
            }
            if (cb > 0u) {  // (the first window was cleared before the block scan)
#pragma unroll
                for (int c = 0; c < kScanItems / 4; ++c)
                    *reinterpret_cast<uint4 *>(&s_mark[tid * kScanItems + 4 * c]) = make_uint4(0u, 0u, 0u, 0u);
                __syncthreads();
            }
            if (cur != first_slot) {
                // slot positions relative to the window, modulo 2^32 (n_global < 2^31: no aliasing)
                const uint32_t wofs = my_off - cb;
                const uint32_t id0 = static_cast<uint32_t>(tid * kScanItems + 1);
#pragma unroll
                for (int k = 0; k < kScanItems; ++k) {
                    const uint32_t pos = wofs + rel[k];
                    if (pos < static_cast<uint32_t>(kScanTile)) s_mark[pos] = id0 + k;
                }
            }
            __syncthreads();
            // inclusive max-scan over the window: first the maximum of the thread's own 16 entries ...
            uint32_t tmax = 0u;
#pragma unroll
            for (int c = 0; c < kScanItems / 4; ++c) {
                const uint4 a = *reinterpret_cast<const uint4 *>(&s_mark[tid * kScanItems + 4 * c]);
                tmax = max(max(tmax, max(a.x, a.y)), max(a.z, a.w));
            }
            uint32_t run = tmax;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) run = max(run, __shfl_up_sync(0xffffffffu, run, o));
            if (lane == 31) s_wmax[warp] = run;
            uint32_t before = __shfl_up_sync(0xffffffffu, run, 1);
            if (lane == 0) before = 0u;
            __syncthreads();
            uint32_t wbefore = carry, wall = carry;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t m = s_wmax[w];
                if (w < warp) wbefore = max(wbefore, m);
                wall = max(wall, m);
            }
            before = max(before, wbefore);
            carry = wall;
            // ... then the entries again, eight at a time, with the running maximum carried along
            uint32_t running = before;
#pragma unroll
            for (int g = 0; g < kScanItems / 8; ++g) {
                uint32_t v[8];
                {
                    const uint4 a = *reinterpret_cast<const uint4 *>(&s_mark[tid * kScanItems + 8 * g]);
                    const uint4 b = *reinterpret_cast<const uint4 *>(&s_mark[tid * kScanItems + 8 * g + 4]);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    running = max(running, v[k]);
                    v[k] = running;
                }
                if (reg_store) {
                    // slots of this handle's own ancestor array (unsharded: all of them): straight from registers
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] += jbase;
                    store8(slot0 + 8u * g, v, S0, S1, true);
                } else {
                    *reinterpret_cast<uint4 *>(&s_mark[tid * kScanItems + 8 * g]) = make_uint4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<uint4 *>(&s_mark[tid * kScanItems + 8 * g + 4]) = make_uint4(v[4], v[5], v[6], v[7]);
                }
            }
            if (!reg_store) {
                __syncthreads();
                // coalesced stores
                const int lim = static_cast<int>(min(static_cast<uint32_t>(kScanTile), span - cb));
#pragma unroll
                for (int k = 0; k < kScanItems; ++k) {
                    const int pos = k * kScanThreads + tid;
                    const long long slot = static_cast<long long>(W0) + cb + pos;
                    if (pos < lim && slot >= S0) expand_store(p, anc, slot, jbase + s_mark[pos]);
                }
            }
            __syncthreads();  // every read of the window is done before the next one is cleared
        }
    }
    __syncthreads();  // s_next (written at the top of this iteration) and the scratch are settled
    }  // tiles
    // poses written into peers' memory must be visible there before the "done" flag of the next kernel
    if (MODE == kScanExpand && p.plan != nullptr) __threadfence_system();
    dbg_stamp(p.dbg, 5);
}

// -------------------------------------------------------------------------------------------------
// k_step : [gather] + predict + weight + reductions, one launch per filter step.
// -------------------------------------------------------------------------------------------------
struct StepParams {
    StateView in, out;
    const uint32_t *anc;   // [T][ld]
    // sharded filter: output slots [0, win_lo) and [win_hi, n) were filled by other ranks; their resampled
    // poses arrive in recv_* ([prefix | suffix]).  Unsharded: win_lo = 0, win_hi = n.
    uint32_t win_lo, win_hi;
    const float *recv_x, *recv_y, *recv_th;
    // device-driven sharded step: the window comes from the plan in device memory and the receive staging is
    // indexed by the output slot itself (recv[i] instead of the packed [prefix | suffix])
    const ShardPlan *plan;
    const float *noise;    // [3][T][n] dense, or null
    const PrepOut *prep;   // [T]
    const double4 *tab;    // [T][Lmax]
    const TrigEntry *trig; // [kTrigEntries]
    double *partials;      // [T][bpt][kPart]
    unsigned int *rec_count;  // [T] producer CTAs finished so far (monotonic; see partial_write)
    const unsigned int *prep_flag;  // [T] epoch of the scalars in `prep` / `tab` (released by the scalars CTA of k_scan)
    unsigned int prep_epoch;
    int early_prep;           // read the scalars behind prep_flag, BEFORE the scan grid has completed
    unsigned int *tickets; // [T][4]
    mcl_stats *stats;      // [T]
    unsigned long long *dbg;  // optional phase timeline [T][bpt][kDbgSlots]
    uint32_t n;            // particles per trial
    uint32_t per;          // particles per block (multiple of kStepUnit)
    long long ld;
    long long slot_base;   // global index of local particle 0 (Philox counters)
    int T, bpt, Lmax;
    int motion;
    uint32_t step;
    uint64_t seed;
    PhiloxKeys rk;         // round keys of seed (Philox noise)
    float vdt_f, wdt_f;    // VEL*dt, YAW_RATE*dt
    double wdt;
    double vel, yaw_rate, dt;
    float q_std[3];
    double inv_s0sq, inv_s1sq;
    float inv_s0_f, inv_s1_f;   // 1 / sigma_obs (fp32 pair kernels)
    float fan_r2, fan_tan, fan_inv_r, fan_inv_sphi, fan_miss;  // MCL_LIK_RB_FAN
    double mb_sig[3];      // motion B: the three "sigma ** 2" scales (motion_model.py:43-48)
};

template <int LIK, bool PHILOX>
__global__ void __launch_bounds__(kStepThreads, kStepCtasPerSm) k_step(const __grid_constant__ StepParams p) {
    const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double4 *s_tab = reinterpret_cast<double4 *>(smem_raw);  // [L] (pair / RB only)
    __shared__ PrepOut s_prep;
    __shared__ uint32_t s_win[2];

    cudaTriggerProgrammaticLaunchCompletion();  // see k_scan
    // The per-step constants below need the scalars CTA of the scan only, the particle loop the whole scan: a CTA that
    // is resident before the scan has drained (programmatic launch) spends that time on its 0.75 us of set-up.
    if (p.early_prep) {
        __shared__ int s_flag_ok;
        if (tid == 0) {
            const unsigned long long t0 = global_timer_ns();
            int ok = 1;
            while (ld_acquire_gpu(p.prep_flag + t) != p.prep_epoch) {
                __nanosleep(200);
                if (global_timer_ns() - t0 > kEarlyWaitNs) { ok = 0; break; }  // (bounded: fall back to the grid dependency)
            }
            s_flag_ok = ok;
        }
        __syncthreads();
        if (!s_flag_ok) cudaGridDependencySynchronize();
    } else {
        cudaGridDependencySynchronize();
    }
    dbg_stamp(p.dbg, 0);
    if (tid < static_cast<int>(sizeof(PrepOut) / sizeof(double)))
        reinterpret_cast<double *>(&s_prep)[tid] = reinterpret_cast<const double *>(p.prep + t)[tid];
    if (tid == 64) {
        s_win[0] = (p.plan != nullptr) ? p.plan->win_lo : p.win_lo;
        s_win[1] = (p.plan != nullptr) ? p.plan->win_hi : p.win_hi;
    }
    __syncthreads();
    const int L = s_prep.L;
    if (LIK == kLikPair || LIK == kLikRB) {
        const double4 *g = p.tab + static_cast<long long>(t) * p.Lmax;
        for (int l = tid; l < L; l += kStepThreads) s_tab[l] = g[l];
        __syncthreads();
    }
    // pair32: per landmark {gx, gx, gy, gy, 2ux, 2ux, 2uy, 2uy} -- both halves of an FFMA2 operand see the same landmark
    float *s_tab8 = reinterpret_cast<float *>(smem_raw);
    if (LIK == kLikPair32) {
        const double4 *g = p.tab + static_cast<long long>(t) * p.Lmax;
        for (int l = tid; l < L; l += kStepThreads) {
            const double4 e = g[l];
            const float gx = static_cast<float>(e.x), gy = static_cast<float>(e.y),
                        ux = static_cast<float>(e.z * sqrt(p.inv_s0sq)), uy = static_cast<float>(e.w * sqrt(p.inv_s1sq));
            *reinterpret_cast<float4 *>(s_tab8 + 8 * l) = make_float4(gx, gx, gy, gy);
            *reinterpret_cast<float4 *>(s_tab8 + 8 * l + 4) = make_float4(ux, ux, uy, uy);  // (2 u / sigma, scaled below)
        }
        __syncthreads();
    }
    if (LIK == kLikRB32) {
        // the per-step scalars CTA left eight floats per landmark; staged with every value twice, so that both halves
        // of an FFMA2 operand (two particles) see the same landmark: 64 bytes per landmark
        const float4 *g = reinterpret_cast<const float4 *>(p.tab + static_cast<long long>(t) * p.Lmax);
        float4 *d = reinterpret_cast<float4 *>(smem_raw);
        for (int l = tid; l < L; l += kStepThreads) {
            const float4 a = g[2 * l], b = g[2 * l + 1];
            d[4 * l] = make_float4(a.x, a.x, a.y, a.y);
            d[4 * l + 1] = make_float4(a.z, a.z, a.w, a.w);
            d[4 * l + 2] = make_float4(b.x, b.x, b.y, b.y);
            d[4 * l + 3] = make_float4(b.z, b.z, b.w, b.w);
        }
        __syncthreads();
    }
    float4 *s_tabf = reinterpret_cast<float4 *>(smem_raw);  // [L detected] (fan sensor: fp32 per pair)
    if (LIK == kLikRBFan) {
        const double4 *g = p.tab + static_cast<long long>(t) * p.Lmax;
        for (int l = tid; l < L; l += kStepThreads) {
            const double4 e = g[l];
            s_tabf[l] = make_float4(static_cast<float>(e.x), static_cast<float>(e.y), static_cast<float>(e.z),
                                    static_cast<float>(e.w));
        }
        __syncthreads();
    }
    const int resample = s_prep.do_resample;
    const float ref_x = static_cast<float>(s_prep.ref[0]), ref_y = static_cast<float>(s_prep.ref[1]),
                ref_th = static_cast<float>(s_prep.ref[2]);
    float ref_s = 0.f, ref_c = 1.f;  // fp32 kernels about the reference pose: its sin / cos
    if (LIK == kLikPair32 || LIK == kLikRB32) {
        double rs, rc;
        sincos_tab(ref_th, p.trig, rs, rc);
        ref_s = static_cast<float>(rs);
        ref_c = static_cast<float>(rc);
    }
    const float max_prev = static_cast<float>(s_prep.max_prev);
    const double shift = s_prep.lik_const - s_prep.c_shift;  // per-step constant added to every log-likelihood
    // moment constants
    const double Ld = static_cast<double>(L);
    const double Slx = s_prep.mom[0], Sly = s_prep.mom[1], Szx = s_prep.mom[2], Szy = s_prep.mom[3];
    const double isoK0 = s_prep.mom[4] + s_prep.mom[5] + s_prep.mom[7] + s_prep.mom[8];  // S_ll + S_zz
    const double isoA = s_prep.mom[9] + s_prep.mom[12];                                  // sum zx lx + zy ly
    const double isoB = s_prep.mom[11] - s_prep.mom[10];                                 // sum zy lx - zx ly
    const double half0 = -0.5 * p.inv_s0sq, half1 = -0.5 * p.inv_s1sq;

    const long long off = static_cast<long long>(t) * p.ld;
    const float *ix = p.in.x + off, *iy = p.in.y + off, *ith = p.in.th + off, *ilw = p.in.lw + off;
    float *ox = p.out.x + off, *oy = p.out.y + off, *oth = p.out.th + off, *olw = p.out.lw + off;
    const uint32_t *anc = p.anc + off;

    // injected noise is the final draw v of particle_filter.py:165; Philox normals are unit and get sqrt(diag Q)
    const float nscale[3] = {PHILOX ? p.q_std[0] : 1.0f, PHILOX ? p.q_std[1] : 1.0f, PHILOX ? p.q_std[2] : 1.0f};

    WarpStats ws;
    ws.init();
    dbg_stamp_real(p.dbg, 1);  // (see mcl_device.cuh: kept in all builds for its effect on register allocation)

    // static contiguous partition of the trial's particles over the blocks (p.per is a multiple of 128,
    // so a warp's 128 particles are either all inside the block's range or all outside)
    const unsigned long long first64 = static_cast<unsigned long long>(blockIdx.x) * p.per;
    const uint32_t first = static_cast<uint32_t>(min(first64, static_cast<unsigned long long>(p.n)));
    const uint32_t last = static_cast<uint32_t>(min(first64 + p.per, static_cast<unsigned long long>(p.n)));

    // The gather is two dependent trips to memory -- ancestors, then the poses they point to -- and 6 warps per
    // scheduler do not cover both (ncu: 22 % of the stall samples sat on these two loads).  Each thread therefore
    // requests the ancestors of its NEXT iteration (one 16-byte cp.async into a slot only it reads back) while it works
    // on the current one: no registers held across the iteration, one trip left exposed (k_step -3.5 % at 125 M).
    __shared__ __align__(16) uint4 s_anc[2][kStepThreads];
    auto request_anc = [&](uint32_t b, int buf) {
        if (resample && b < last) {
            const unsigned int dst = static_cast<unsigned int>(__cvta_generic_to_shared(&s_anc[buf][tid]));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(anc + b) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (p.early_prep) cudaGridDependencySynchronize();  // ancestors, and the state buffer the scan's scalars CTA read
    request_anc(first + tid * kPPT, 0);
    int since_flush = 0, itn = 0;
    for (uint32_t wbase = first + (tid - lane) * kPPT; wbase < last; wbase += kStepTile, ++itn) {
        const uint32_t base = wbase + lane * kPPT;
        const bool active = base < last;  // false only in the last, partial unit of the trial
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const uint4 a4_pre = s_anc[itn & 1][tid];
        request_anc(base + kStepTile, (itn + 1) & 1);
        float x[kPPT], y[kPPT], th[kPPT], lwp[kPPT] = {0.f, 0.f, 0.f, 0.f};
        float lwf[kPPT] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        double s[kPPT], c[kPPT];
        if (active) {
            // ---------------- load (gather through the ancestors when this step resamples) ----------------
            if (resample) {
                const uint4 a4 = a4_pre;
                const uint32_t a[kPPT] = {a4.x, a4.y, a4.z, a4.w};
                const uint32_t win_lo = s_win[0], win_hi = s_win[1];
                if (base >= win_lo && base + kPPT <= win_hi) {
#pragma unroll
                    for (int k = 0; k < kPPT; ++k) {
                        const uint32_t j = (base + k < p.n) ? a[k] : 0u;
                        x[k] = ix[j];
                        y[k] = iy[j];
                        th[k] = ith[j];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kPPT; ++k) {
                        const uint32_t i = base + k;
                        if ((i >= win_lo && i < win_hi) || i >= p.n) {  // (padding lanes read particle 0)
                            const uint32_t j = (i < p.n) ? a[k] : 0u;
                            x[k] = ix[j];
                            y[k] = iy[j];
                            th[k] = ith[j];
                        } else {
                            const uint32_t r = (p.plan != nullptr || i < win_lo) ? i : i - win_hi + win_lo;
                            x[k] = p.recv_x[r];
                            y[k] = p.recv_y[r];
                            th[k] = p.recv_th[r];
                        }
                    }
                }
            } else {
                const float4 vx = *reinterpret_cast<const float4 *>(ix + base);
                const float4 vy = *reinterpret_cast<const float4 *>(iy + base);
                const float4 vt = *reinterpret_cast<const float4 *>(ith + base);
                const float4 vw = *reinterpret_cast<const float4 *>(ilw + base);
                x[0] = vx.x; x[1] = vx.y; x[2] = vx.z; x[3] = vx.w;
                y[0] = vy.x; y[1] = vy.y; y[2] = vy.z; y[3] = vy.w;
                th[0] = vt.x; th[1] = vt.y; th[2] = vt.z; th[3] = vt.w;
                // weights stay relative to their maximum: pw * bn (:194) in logs, minus a per-step constant
                lwp[0] = vw.x - max_prev; lwp[1] = vw.y - max_prev; lwp[2] = vw.z - max_prev; lwp[3] = vw.w - max_prev;
                if (s_prep.uniform) lwp[0] = lwp[1] = lwp[2] = lwp[3] = 0.0f;  // weights were reset to 1/NP (:234-237)
            }

            // ---------------- motion noise (unit normals; Motion A scales them by sqrt(diag Q)) ----------------
            float nz[kPPT][3];
            if (!PHILOX) {
                const long long plane = static_cast<long long>(p.T) * p.n;
                const float *nb = p.noise + static_cast<long long>(t) * p.n;
#pragma unroll
                for (int k = 0; k < kPPT; ++k) {
                    const bool ok = base + k < p.n;
#pragma unroll
                    for (int d = 0; d < 3; ++d) nz[k][d] = ok ? nb[d * plane + base + k] : 0.0f;
                }
            } else {
                // 2 Philox calls = 256 bits = twelve 21-bit uniforms = the 12 normals of this thread's 4 particles
                const unsigned long long grp = static_cast<unsigned long long>(p.slot_base + base) >> 2;
                const uint4 r0 = philox4x32_10(make_uint4(static_cast<uint32_t>(grp), static_cast<uint32_t>(grp >> 32) * 2u,
                                                          p.step, static_cast<uint32_t>(t)), p.rk);
                const uint4 r1 = philox4x32_10(make_uint4(static_cast<uint32_t>(grp), static_cast<uint32_t>(grp >> 32) * 2u + 1u,
                                                          p.step, static_cast<uint32_t>(t)), p.rk);
                const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                float nrm[12];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    // words 2q, 2q+1 -> three 21-bit fields: top 21 of each word, and 11 + 10 of their low bits
                    const uint32_t f0 = w[2 * q] >> 11, f1 = w[2 * q + 1] >> 11;
                    const uint32_t f2 = ((w[2 * q] & 0x7ffu) << 10) | (w[2 * q + 1] & 0x3ffu);
                    // fields pair up across the two halves: (q, q+2) share one Box-Muller for their third normal
                    const float2 a = box_muller21(f0, f1);
                    nrm[3 * q + 0] = a.x;
                    nrm[3 * q + 1] = a.y;
                    nrm[3 * q + 2] = __uint_as_float(f2);  // third field, paired below
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float2 a = box_muller21(__float_as_uint(nrm[3 * q + 2]), __float_as_uint(nrm[3 * (q + 2) + 2]));
                    nrm[3 * q + 2] = a.x;
                    nrm[3 * (q + 2) + 2] = a.y;
                }
#pragma unroll
                for (int k = 0; k < kPPT; ++k)
#pragma unroll
                    for (int d = 0; d < 3; ++d) nz[k][d] = nrm[3 * k + d];
            }

            // ---------------- predict ----------------
            if (p.motion == MCL_MOTION_A) {
#pragma unroll
                for (int k = 0; k < kPPT; ++k) {
                    // particle_filter.py:129-142 then "+ v" (:166): cos/sin of the PRE-update yaw; the wrap
                    // is applied before the noise, so the stored yaw may leave [-pi, pi]
                    float s0, c0;
                    __sincosf(th[k], &s0, &c0);
                    x[k] = fmaf(nz[k][0], nscale[0], fmaf(p.vdt_f, c0, x[k]));
                    y[k] = fmaf(nz[k][1], nscale[1], fmaf(p.vdt_f, s0, y[k]));
                    float tn = th[k] + p.wdt_f;
                    if (!(fabsf(tn) < 3.1415f))  // at or past the seam: wrap exactly like mylib/limit.py, in fp64
                        tn = static_cast<float>(limit_angle(static_cast<double>(th[k]) + p.wdt));
                    th[k] = fmaf(nz[k][2], nscale[2], tn);
                }
            } else {
#pragma unroll
                for (int k = 0; k < kPPT; ++k) {
                    // motion_model.py:46-56; sigma**2 used as the standard deviation (reference quirk)
                    const double thd = static_cast<double>(th[k]);
                    const double vh = p.vel + static_cast<double>(nz[k][0]) * p.mb_sig[0];
                    const double wh = p.yaw_rate + static_cast<double>(nz[k][1]) * p.mb_sig[1];
                    const double gh = static_cast<double>(nz[k][2]) * p.mb_sig[2];
                    const double a = vh / wh, b = wh * p.dt;
                    double s0, c0, s1, c1;
                    sincos(thd, &s0, &c0);
                    sincos(thd + b, &s1, &c1);
                    x[k] = static_cast<float>(static_cast<double>(x[k]) - a * s0 + a * s1);
                    y[k] = static_cast<float>(static_cast<double>(y[k]) + a * c0 - a * c1);
                    th[k] = static_cast<float>(limit_angle(thd + (wh + gh) * p.dt));
                }
            }

            // ---------------- weight: log-likelihood of the STORED (fp32) pose, fp64 arithmetic ----------------
            float fan_chi2[kPPT] = {0.f, 0.f, 0.f, 0.f};
            float dll32[kPPT] = {0.f, 0.f, 0.f, 0.f};  // fp32 kernels: the new stored log-weight (relative to the reference pose's)
            if (LIK == kLikPair32) {
                // The particle sits at the reference pose plus a small (dx, dy, dth).  With M(th) = [[s, -c], [c, s]]
                // and g_l = M(ref)(l - p_ref):  M(th)(l - p) = g_l + (cos dth - 1) g_l + sin dth J g_l - M(th)(dx, dy),
                // J(g) = (g_y, -g_x).  Everything after the first g_l is the small d_l the kernel works with; its
                // rounding error scales with |d_l|, not with the 20-m coordinates.
                // Components are scaled by 1 / sigma before they are squared, so ONE accumulator per particle holds the
                // chi^2 DIFFERENCE to the reference pose.  It runs in fp32 over blocks of kFoldEvery landmarks and is
                // folded into fp64 between blocks: a plain fp32 sum over 1024 landmarks was measured at 1.5e-5 of
                // log-weight on the particles that carry mass (+7 % time for the folding).
                const float is0 = p.inv_s0_f, is1 = p.inv_s1_f;
                float a0[kPPT], b0[kPPT], tx0[kPPT], a1[kPPT], nb1[kPPT], ty1[kPPT];
#pragma unroll
                for (int k = 0; k < kPPT; ++k) {
                    float sd, cm1;
                    sincosm1_small(wrap_residual(th[k] - ref_th), sd, cm1);
                    const float sf = fmaf(ref_s, cm1, fmaf(ref_c, sd, ref_s));   // sin(th) = s_r cos d + c_r sin d
                    const float cf = fmaf(ref_c, cm1, fmaf(-ref_s, sd, ref_c));  // cos(th) = c_r cos d - s_r sin d
                    const float ddx = x[k] - ref_x, ddy = y[k] - ref_y;
                    a0[k] = cm1 * is0;
                    b0[k] = sd * is0;
                    tx0[k] = -fmaf(sf, ddx, -cf * ddy) * is0;
                    a1[k] = cm1 * is1;
                    nb1[k] = -sd * is1;
                    ty1[k] = -fmaf(cf, ddx, sf * ddy) * is1;
                    s[k] = static_cast<double>(sf);
                    c[k] = static_cast<double>(cf);
                }
                f32x2 A0[2], B0[2], TX[2], A1[2], NB1[2], TY[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    A0[h] = pack2(a0[2 * h], a0[2 * h + 1]);
                    B0[h] = pack2(b0[2 * h], b0[2 * h + 1]);
                    TX[h] = pack2(tx0[2 * h], tx0[2 * h + 1]);
                    A1[h] = pack2(a1[2 * h], a1[2 * h + 1]);
                    NB1[h] = pack2(nb1[2 * h], nb1[2 * h + 1]);
                    TY[h] = pack2(ty1[2 * h], ty1[2 * h + 1]);
                }
                double dq[kPPT] = {0.0, 0.0, 0.0, 0.0};
                for (int l0 = 0; l0 < L; l0 += kFoldEvery) {
                    f32x2 Q[2] = {0ull, 0ull};
                    const int l1 = min(L, l0 + kFoldEvery);
#pragma unroll 2
                    for (int l = l0; l < l1; ++l) {
                        const ulonglong2 g = *reinterpret_cast<const ulonglong2 *>(s_tab8 + 8 * l);      // {gx gx}, {gy gy}
                        const ulonglong2 u = *reinterpret_cast<const ulonglong2 *>(s_tab8 + 8 * l + 4);  // {2ux/s0 ..}, {2uy/s1 ..}
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const f32x2 dx = fma2(A0[h], g.x, fma2(B0[h], g.y, TX[h]));
                            Q[h] = fma2(dx, add2(dx, u.x), Q[h]);
                            const f32x2 dy = fma2(A1[h], g.y, fma2(NB1[h], g.x, TY[h]));
                            Q[h] = fma2(dy, add2(dy, u.y), Q[h]);
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float q0, q1;
                        unpack2(Q[h], q0, q1);
                        dq[2 * h] += static_cast<double>(q0);
                        dq[2 * h + 1] += static_cast<double>(q1);
                    }
                }
#pragma unroll
                for (int k = 0; k < kPPT; ++k) dll32[k] = static_cast<float>(static_cast<double>(lwp[k]) - 0.5 * dq[k]);
            }
            if (LIK == kLikRB32) {
                float c1_[kPPT], b_[kPPT], ntx[kPPT], nty[kPPT];
#pragma unroll
                for (int k = 0; k < kPPT; ++k) {
                    float sd, cm1;
                    sincosm1_small(wrap_residual(th[k] - ref_th), sd, cm1);
                    const float sf = fmaf(ref_s, cm1, fmaf(ref_c, sd, ref_s));
                    const float cf = fmaf(ref_c, cm1, fmaf(-ref_s, sd, ref_c));
                    const float ddx = x[k] - ref_x, ddy = y[k] - ref_y;
                    c1_[k] = 1.0f + cm1;
                    b_[k] = sd;
                    ntx[k] = -fmaf(sf, ddx, -cf * ddy);
                    nty[k] = -fmaf(cf, ddx, sf * ddy);
                    s[k] = static_cast<double>(sf);
                    c[k] = static_cast<double>(cf);
                }
                // The particle sees the landmark at R g + T: the reference pose's view g turned by the yaw difference
                // delta (R = [[cos, sin], [-sin, cos]]) and shifted by T = -M(theta)(dx, dy).  With u = g.T, v = g x T:
                //   |R g + T|^2 - |g|^2 = |T|^2 + 2 (cos delta u - sin delta v)   (the turn changes no norm: no cancellation)
                //   g . (R g + T) = cos delta |g|^2 + u,   g x (R g + T) = -sin delta |g|^2 + v
                // Two particles per register pair (FFMA2 / FMUL2 / FADD2): 23 packed instructions, 3 MUFU and the clamp per
                // TWO pairs on the main path; limit_angle's +-2 pi turn only for a landmark whose reference residual lies
                // within 0.11 rad of the seam (a warp-uniform test) or a particle far from the reference pose.
                f32x2 C1[2], NB[2], TX[2], TY[2], NTX[2], T2h[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k0 = 2 * h, k1 = 2 * h + 1;
                    C1[h] = pack2(c1_[k0], c1_[k1]);
                    NB[h] = pack2(-b_[k0], -b_[k1]);
                    TX[h] = pack2(ntx[k0], ntx[k1]);
                    NTX[h] = pack2(-ntx[k0], -ntx[k1]);
                    TY[h] = pack2(nty[k0], nty[k1]);
                    T2h[h] = pack2(0.5f * fmaf(ntx[k0], ntx[k0], nty[k0] * nty[k0]), 0.5f * fmaf(ntx[k1], ntx[k1], nty[k1] * nty[k1]));
                }
                const f32x2 kOne = pack2(1.0f, 1.0f);
                const f32x2 kC3 = pack2(-1.0f / 3.0f, -1.0f / 3.0f), kC5 = pack2(0.2f, 0.2f), kC7 = pack2(-1.0f / 7.0f, -1.0f / 7.0f);
                const ulonglong2 *tab2 = reinterpret_cast<const ulonglong2 *>(smem_raw);
                double dq[kPPT] = {0.0, 0.0, 0.0, 0.0};
                for (int l0 = 0; l0 < L; l0 += kFoldEvery) {
                    f32x2 Qr[2] = {0ull, 0ull}, Qp[2] = {0ull, 0ull};  // range terms in sigma_d units, bearing terms in rad^2
                    const int l1 = min(L, l0 + kFoldEvery);
#pragma unroll 1
                    for (int l = l0; l < l1; ++l) {
                        const ulonglong2 t0 = tab2[4 * l];      // {gx gx}, {gy gy}
                        const ulonglong2 t1 = tab2[4 * l + 1];  // {|g|/sigma_d ..}, {2/|g|^2 ..}
                        const ulonglong2 t2q = tab2[4 * l + 2]; // {|g|^2 ..}, {2 e_d/sigma_d ..}
                        const ulonglong2 t3 = tab2[4 * l + 3];  // {e_phi ..}, {2 e_phi ..}
                        float ephi, ephi_;
                        unpack2(t3.x, ephi, ephi_);
                        const bool seam = fabsf(ephi) > 3.03f;  // |dphi| < 0.1003 on the main path: no turn possible elsewhere
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const f32x2 u = fma2(t0.x, TX[h], mul2(t0.y, TY[h]));
                            const f32x2 v = fma2(t0.x, TY[h], mul2(t0.y, NTX[h]));
                            const f32x2 den = fma2(C1[h], t2q.x, u);
                            const f32x2 crs = fma2(NB[h], t2q.x, v);
                            // range: |g| (sqrt(1 + eps) - 1) = |g| eps / (1 + sqrt(1 + eps)), eps = (|R g + T|^2 - |g|^2) / |g|^2.
                            // (Its series in eps with this closed form as the fallback measured SLOWER, 28.8 vs 26.4 ms at C4:
                            // some landmark is always at arm's length of some lane, so the warp pays both.  One MUFU.RCP of
                            // the product (1 + sqrt(1 + eps)) (|g|^2 + g.d) in place of the two below: 27.7 ms.  The loop is
                            // runs at FMA pipe 54 %, XU 51 %, issue 63 % (profiles/r02_h_ncu_c4_rb32_k_step_summary.csv): no one pipe
                            // to relieve.  ONE fallback branch per landmark behind both halves, so that they interleave: 25.7 ms.)
                            const f32x2 eps = mul2(fma2(C1[h], u, fma2(NB[h], v, T2h[h])), t1.y);
                            float o0, o1, r0, r1;
                            unpack2(add2(kOne, eps), o0, o1);
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(1.0f + sqrt_approx(fmaxf(o0, 0.0f))));
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(1.0f + sqrt_approx(fmaxf(o1, 0.0f))));
                            const f32x2 dr = mul2(mul2(t1.x, eps), pack2(r0, r1));
                            Qr[h] = fma2(dr, sub2(dr, t2q.y), Qr[h]);  // (e_d - dr)^2 - e_d^2, dr = (range change) / sigma_d
                            // bearing: the change is atan2(g x (R g + T), g . (R g + T)): a small angle for the particles that
                            // matter, where the series keeps its RELATIVE accuracy
                            float d0, d1;
                            unpack2(den, d0, d1);
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
                            const f32x2 tq = mul2(crs, pack2(r0, r1)), tt = mul2(tq, tq);
                            f32x2 dphi = fma2(mul2(tq, tt), fma2(tt, fma2(tt, kC7, kC5), kC3), tq);
                            float q0, q1;
                            unpack2(tt, q0, q1);
                            const bool ok0 = d0 > 0.0f && q0 < 0.01f, ok1 = d1 > 0.0f && q1 < 0.01f;
                            if (!(ok0 && ok1) || seam) {
                                float c0, c1, p0, p1;
                                unpack2(crs, c0, c1);
                                unpack2(dphi, p0, p1);
                                if (!ok0) p0 = atan2_fast(c0, d0);  // far from the reference pose
                                if (!ok1) p1 = atan2_fast(c1, d1);
                                // limit_angle (mylib/limit.py:19-26) as the +-2 pi turn of (new residual - old residual)
                                const float w0 = ephi - p0, w1 = ephi - p1;
                                p0 -= (w0 > 3.14159265f) ? -6.28318531f : ((w0 < -3.14159265f) ? 6.28318531f : 0.0f);
                                p1 -= (w1 > 3.14159265f) ? -6.28318531f : ((w1 < -3.14159265f) ? 6.28318531f : 0.0f);
                                dphi = pack2(p0, p1);
                            }
                            Qp[h] = fma2(dphi, sub2(dphi, t3.y), Qp[h]);  // (e_phi - dphi)^2 - e_phi^2, in rad^2
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float q0, q1, f0, f1;
                        unpack2(Qr[h], q0, q1);
                        unpack2(Qp[h], f0, f1);
                        dq[2 * h] += fma(static_cast<double>(f0), p.inv_s1sq, static_cast<double>(q0));
                        dq[2 * h + 1] += fma(static_cast<double>(f1), p.inv_s1sq, static_cast<double>(q1));
                    }
                }
#pragma unroll
                for (int k = 0; k < kPPT; ++k) dll32[k] = static_cast<float>(static_cast<double>(lwp[k]) - 0.5 * dq[k]);
            }
            if (LIK == kLikRBFan) {
                // the fan sensor is scored in fp32, one table entry against the thread's four particles at a time
                float sf[kPPT], cf[kPPT];
#pragma unroll
                for (int k = 0; k < kPPT; ++k) {
                    sincos_tab(th[k], p.trig, s[k], c[k]);
                    sf[k] = static_cast<float>(s[k]);
                    cf[k] = static_cast<float>(c[k]);
                }
#pragma unroll 2
                for (int l = 0; l < L; ++l) {
                    const float4 e = s_tabf[l];
#pragma unroll
                    for (int k = 0; k < kPPT; ++k) {
                        const float dx = e.x - x[k], dy = e.y - y[k];
                        fan_chi2[k] += fan_pair_chi2<float>(fmaf(sf[k], dx, -cf[k] * dy), fmaf(cf[k], dx, sf[k] * dy), e.z, e.w,
                                                            p.fan_r2, p.fan_tan, p.fan_inv_r, p.fan_inv_sphi, p.fan_miss);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kPPT; ++k) {
                if (LIK == kLikPair32 || LIK == kLikRB32) {  // already relative to the reference pose (and to lwp)
                    lwf[k] = dll32[k];
                    continue;
                }
                if (LIK != kLikRBFan) sincos_tab(th[k], p.trig, s[k], c[k]);
                const double xd = x[k], yd = y[k];
                double ll;
                if (LIK == kLikMomentIso) {
                    // sum_l |M(l - p) - z_l|^2 = S_ll + S_zz - 2 p.S_l + L |p|^2 - 2 (s A + c B) + 2 S_z.(M p),
                    // M = [[s, -c], [c, s]]  (rotations preserve the norm; exact identity, evaluated in fp64)
                    const double mpx = fma(s[k], xd, -(c[k] * yd)), mpy = fma(c[k], xd, s[k] * yd);
                    double Q = fma(Ld, fma(xd, xd, yd * yd), isoK0);
                    Q = fma(-2.0, fma(xd, Slx, yd * Sly), Q);
                    Q = fma(-2.0, fma(s[k], isoA, c[k] * isoB), Q);
                    Q = fma(2.0, fma(Szx, mpx, Szy * mpy), Q);
                    ll = half0 * Q;
                } else if (LIK == kLikMomentGen) {
                    // sum_l ex^2 = S(aa) + 2 tx S(a) + L tx^2 with a_l = s lx - c ly - zx, tx = -(s x - c y)
                    const double tx = -(s[k] * xd - c[k] * yd), ty = -(c[k] * xd + s[k] * yd);
                    const double ss = s[k] * s[k], cc = c[k] * c[k], sc = s[k] * c[k];
                    const double Sxx = s_prep.mom[4], Syy = s_prep.mom[5], Sxy = s_prep.mom[6];
                    const double Sa = s[k] * Slx - c[k] * Sly - Szx;
                    const double Saa = ss * Sxx + cc * Syy - 2.0 * sc * Sxy + s_prep.mom[7] -
                                       2.0 * s[k] * s_prep.mom[9] + 2.0 * c[k] * s_prep.mom[10];
                    const double Sb = c[k] * Slx + s[k] * Sly - Szy;
                    const double Sbb = cc * Sxx + ss * Syy + 2.0 * sc * Sxy + s_prep.mom[8] -
                                       2.0 * c[k] * s_prep.mom[11] - 2.0 * s[k] * s_prep.mom[12];
                    const double q0 = Saa + 2.0 * tx * Sa + Ld * tx * tx;
                    const double q1 = Sbb + 2.0 * ty * Sb + Ld * ty * ty;
                    ll = fma(half0, q0, half1 * q1);
                } else if (LIK == kLikRBFan) {
                    ll = -0.5 * static_cast<double>(fan_chi2[k]);
                } else {
                    const double tx = -(s[k] * xd - c[k] * yd), ty = -(c[k] * xd + s[k] * yd);
                    double q0 = 0.0, q1 = 0.0;
                    if (LIK == kLikPair) {
#pragma unroll 4
                        for (int l = 0; l < L; ++l) {
                            double e0, e1;
                            pair_residual_c(s[k], c[k], tx, ty, s_tab[l], e0, e1);
                            q0 = fma(e0, e0, q0);
                            q1 = fma(e1, e1, q1);
                        }
                    } else {
                        for (int l = 0; l < L; ++l) {
                            double e0, e1;
                            pair_residual_rb(s[k], c[k], tx, ty, s_tab[l], e0, e1);
                            q0 = fma(e0, e0, q0);
                            q1 = fma(e1, e1, q1);
                        }
                    }
                    ll = fma(half0, q0, half1 * q1);
                }
                lwf[k] = static_cast<float>(static_cast<double>(lwp[k]) + (ll + shift));
            }
            // ---------------- store ----------------
            *reinterpret_cast<float4 *>(ox + base) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4 *>(oy + base) = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4 *>(oth + base) = make_float4(th[0], th[1], th[2], th[3]);
            *reinterpret_cast<float4 *>(olw + base) = make_float4(lwf[0], lwf[1], lwf[2], lwf[3]);
#pragma unroll
            for (int k = 0; k < kPPT; ++k)
                if (base + k >= p.n) lwf[k] = -INFINITY;  // padding of the last unit: no statistics
        }

        // ---------------- running statistics of what was stored (all lanes take part in the votes) ----------------
        ws.raise_max(fmaxf(fmaxf(lwf[0], lwf[1]), fmaxf(lwf[2], lwf[3])));
        if (active) {
#pragma unroll
            for (int k = 0; k < kPPT; ++k)
                if (base + k < p.n) ws.add(lwf[k], x[k], y[k], th[k], s[k], c[k], ref_x, ref_y, ref_th, base + k);
        }
        if (++since_flush == kFlushEvery) {
            ws.flush();
            since_flush = 0;
        }
    }

    dbg_stamp(p.dbg, 2);
    if (blockIdx.x == 0 && tid == 0) {
        // what the merge of the partials cannot know; and re-arm the scan ticket (the scan has completed)
        mcl_stats *st = p.stats + t;
        st->ref[0] = s_prep.ref[0];
        st->ref[1] = s_prep.ref[1];
        st->ref[2] = s_prep.ref[2];
        st->resampled = resample;
        st->step = static_cast<int>(p.step + 1u);
        st->abs_shift = s_prep.abs_shift;
        p.tickets[t * 4 + 0] = 0u;
    }
    partial_write(ws, p.partials + (static_cast<long long>(t) * p.bpt + blockIdx.x) * kPart, p.dbg, p.rec_count ? p.rec_count + t : nullptr);
}

// -------------------------------------------------------------------------------------------------
// k_reduce : statistics of the stored state (after mcl_set_particles); nothing is written back.
// The reference pose of the moment sums is particle 0 (from_particle0) or the stored estimate.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStepThreads, 2) k_reduce(const StepParams p, int from_particle0) {
    const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const long long off = static_cast<long long>(t) * p.ld;
    const float *ix = p.in.x + off, *iy = p.in.y + off, *ith = p.in.th + off, *ilw = p.in.lw + off;
    float ref_x, ref_y, ref_th;
    if (from_particle0) {
        ref_x = ix[0];
        ref_y = iy[0];
        ref_th = ith[0];
    } else {
        ref_x = static_cast<float>(p.stats[t].est[0]);
        ref_y = static_cast<float>(p.stats[t].est[1]);
        ref_th = static_cast<float>(p.stats[t].est[2]);
    }
    WarpStats ws;
    ws.init();
    int since_flush = 0;
    const uint32_t stride = gridDim.x * kStepThreads;
    for (uint32_t wb = blockIdx.x * kStepThreads + (tid - lane); wb < p.n; wb += stride) {
        const uint32_t i = wb + lane;
        const bool active = i < p.n;
        float lw = -INFINITY, x = 0.f, y = 0.f, th = 0.f;
        if (active) {
            lw = ilw[i];
            x = ix[i];
            y = iy[i];
            th = ith[i];
        }
        ws.raise_max(lw);
        if (active) {
            double s, c;
            sincos(static_cast<double>(th), &s, &c);
            ws.add(lw, x, y, th, s, c, ref_x, ref_y, ref_th, i);
        }
        if (++since_flush == 32) {
            ws.flush();
            since_flush = 0;
        }
        if (wb + stride < wb) break;  // 32-bit wrap guard for n close to 2^32
    }
    if (blockIdx.x == 0 && tid == 0) {
        mcl_stats *st = p.stats + t;
        st->ref[0] = ref_x;
        st->ref[1] = ref_y;
        st->ref[2] = ref_th;
        st->abs_shift = kNoAbsShift;  // weights set by the caller: relative only, no likelihood step behind them
        p.tickets[t * 4 + 0] = 0u;
    }
    partial_write(ws, p.partials + (static_cast<long long>(t) * p.bpt + blockIdx.x) * kPart, nullptr, p.rec_count ? p.rec_count + t : nullptr);
}

// -------------------------------------------------------------------------------------------------
// k_gather : stand-alone __resampling (particle_filter.py:216-222): px[:, i] = px_temp[:, idx], pw = 1/NP
// -------------------------------------------------------------------------------------------------
struct GatherParams {
    StateView in, out;
    const uint32_t *anc;
    const PrepOut *prep;
    mcl_stats *stats;
    unsigned int *tickets;
    long long n, ld;
    double n_global;
};

__global__ void __launch_bounds__(256) k_gather(const GatherParams p) {
    const int t = blockIdx.y;
    const PrepOut *po = p.prep + t;
    const int rs = po->do_resample;
    const long long off = static_cast<long long>(t) * p.ld;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        if (rs) {
            const uint32_t j = p.anc[off + i];
            p.out.x[off + i] = p.in.x[off + j];
            p.out.y[off + i] = p.in.y[off + j];
            p.out.th[off + i] = p.in.th[off + j];
            p.out.lw[off + i] = 0.0f;
        } else {
            p.out.x[off + i] = p.in.x[off + i];
            p.out.y[off + i] = p.in.y[off + i];
            p.out.th[off + i] = p.in.th[off + i];
            p.out.lw[off + i] = p.in.lw[off + i];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.tickets[t * 4 + 0] = 0u;  // re-arm the scan ticket (no k_step follows in the stage API)
        mcl_stats *st = p.stats + t;  // (merged before this launch: ensure_stats)
        st->resampled = rs;
        if (rs) {  // weights reset to 1/NP (:222): stored log-weights 0, normaliser log NP
            st->ess = p.n_global;
            st->max_logw = 0.0;
            st->sum_w = p.n_global;
            st->sum_w2 = p.n_global;
            st->raw[11] = 0.0;
            st->derived = 0;
            st->underflow = 0;  // the log-weights were physically reset
            st->abs_shift = kNoAbsShift;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// k_cov : exact weighted covariance about the (already known) mean, yaw residual wrapped (two-pass)
// -------------------------------------------------------------------------------------------------
struct CovParams {
    StateView in;
    mcl_stats *stats;
    double *partials;       // [T][bpt][8]
    unsigned int *tickets;  // slot 2
    long long n, ld;
    int bpt;
};

__global__ void __launch_bounds__(256) k_cov(const CovParams p) {
    const int t = blockIdx.y, tid = threadIdx.x;
    __shared__ double s_red[8 * 8];
    __shared__ int s_last;
    __shared__ double s_mean[4];
    if (tid == 0) {
        mcl_stats d = p.stats[t];
        derive_stats(d);
        s_mean[0] = d.mean[0]; s_mean[1] = d.mean[1]; s_mean[2] = d.mean[2]; s_mean[3] = d.log_norm;
    }
    __syncthreads();
    const double mx = s_mean[0], my = s_mean[1], mth = s_mean[2], ln = s_mean[3];
    const long long off = static_cast<long long>(t) * p.ld;
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + tid; i < p.n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const double w = exp(static_cast<double>(p.in.lw[off + i]) - ln);
        const double dx = p.in.x[off + i] - mx, dy = p.in.y[off + i] - my;
        const double dt = limit_angle(static_cast<double>(p.in.th[off + i]) - mth);
        a[0] = fma(w * dx, dx, a[0]); a[1] = fma(w * dx, dy, a[1]); a[2] = fma(w * dy, dy, a[2]);
        a[3] = fma(w * dx, dt, a[3]); a[4] = fma(w * dy, dt, a[4]); a[5] = fma(w * dt, dt, a[5]);
    }
    block_sum<6>(a, s_red);
    if (tid == 0) {
        double *pp = p.partials + (static_cast<long long>(t) * p.bpt + blockIdx.x) * 8;
        for (int k = 0; k < 6; ++k) pp[k] = a[k];
        __threadfence();
        s_last = (atomicAdd(&p.tickets[t * 4 + 2], 1u) == gridDim.x - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double s[6] = {0, 0, 0, 0, 0, 0};
    const double *pp = p.partials + static_cast<long long>(t) * p.bpt * 8;
    for (int b = tid; b < static_cast<int>(gridDim.x); b += blockDim.x)
        for (int k = 0; k < 6; ++k) s[k] += __ldcg(pp + b * 8 + k);
    block_sum<6>(s, s_red);
    if (tid == 0) {
        mcl_stats *o = p.stats + t;
        o->mean[0] = mx; o->mean[1] = my; o->mean[2] = mth;
        o->log_norm = ln;
        o->max_weight = exp(o->raw[11] - o->max_logw) / o->sum_w;
        o->derived = 1;  // exact two-pass moments: mcl_get_stats must not overwrite them
        o->cov[0] = s[0]; o->cov[1] = s[1]; o->cov[2] = s[3];
        o->cov[3] = s[1]; o->cov[4] = s[2]; o->cov[5] = s[4];
        o->cov[6] = s[3]; o->cov[7] = s[4]; o->cov[8] = s[5];
        p.tickets[t * 4 + 2] = 0u;
    }
}

// -------------------------------------------------------------------------------------------------
// small utility kernels
// -------------------------------------------------------------------------------------------------
__global__ void k_init(StateView s, const double *pose, long long n, long long ld, float lw0) {
    const int t = blockIdx.y;
    const float x = static_cast<float>(pose[3 * t]), y = static_cast<float>(pose[3 * t + 1]),
                th = static_cast<float>(pose[3 * t + 2]);
    const long long off = static_cast<long long>(t) * ld;
    // pad elements beyond n are initialised too (they are read by vector loads, never used)
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < ld;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        s.x[off + i] = x; s.y[off + i] = y; s.th[off + i] = th; s.lw[off + i] = lw0;
    }
}

__global__ void k_init_stats(mcl_stats *stats, const double *pose, double n_global, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    mcl_stats z = {};  // stored log-weights are 0: w = exp(0 - log NP) = 1/NP (particle_filter.py:83-84)
    z.ess = n_global;
    z.max_idx = 0;
    z.sum_w = n_global; z.sum_w2 = n_global; z.max_logw = 0.0;
    z.abs_shift = kNoAbsShift;
    for (int k = 0; k < 3; ++k) { z.est[k] = pose[3 * t + k]; z.mean[k] = pose[3 * t + k]; z.ref[k] = pose[3 * t + k]; }
    stats[t] = z;
}

__global__ void k_fill(float *dst, long long ld, float value) {
    const int t = blockIdx.y;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < ld;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        dst[t * ld + i] = value;
}

// sharded resampling runs the scan outside a step: clear its tile descriptors and ticket (trial 0)
__global__ void k_reset_scan(unsigned long long *desc, unsigned int *ready, unsigned int *tickets, int tpt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < tpt) desc[i] = 0ull;
    if (i < kDirectTiles / 32) ready[i] = 0u;
    if (i == 0) tickets[0] = 0u;
}

// Sum of a shard's quantised weights under the (global) weight map of stats[0]: a plain streaming reduction.
// Integer adds are exact and order-independent, so one 64-bit atomicAdd per warp is deterministic.
// Sharded filter: merge the shards' mcl_stats records (allgathered, device memory) into the global statistics.
// Device twin of sharded.merge_stats: shards share the reference pose of the moment sums (it derives from the
// previous GLOBAL estimate), so their un-normalised sums add after rescaling to the global maximum; a shard
// that used another reference (right after mcl_set_particles) is re-centred.  One thread: G is the GPU count.
struct ShardMergeParams {
    const mcl_stats *rec;  // [G]
    int G;
    long long m;           // particles per shard: global index of shard g's particle i is g * m + i
    mcl_stats *out;        // this handle's statistics block (becomes the global one)
    int apply_rule;        // device-driven exchange: the records carry particle 0's pose, the underflow rule can be applied
};
__device__ __noinline__ void shard_merge_records(const ShardMergeParams p, bool derive) {
    double M = p.rec[0].max_logw;
    for (int g = 1; g < p.G; ++g) M = fmax(M, p.rec[g].max_logw);
    double s1 = 0.0, s2 = 0.0, raw[11];
    for (int c = 0; c < 11; ++c) raw[c] = 0.0;
    const double rx = p.rec[0].ref[0], ry = p.rec[0].ref[1], rt = p.rec[0].ref[2];
    int bg = 0;
    double bv = p.rec[0].max_logw;
    for (int g = 0; g < p.G; ++g) {
        const mcl_stats &r = p.rec[g];
        const double sc = (r.sum_w > 0.0) ? exp(r.max_logw - M) : 0.0;
        const double e = r.sum_w * sc;
        s1 += e;
        s2 += r.sum_w2 * sc * sc;
        double a[11];
        for (int c = 0; c < 11; ++c) a[c] = r.raw[c] * sc;
        const double dx = r.ref[0] - rx, dy = r.ref[1] - ry, dt = limit_angle_hd(r.ref[2] - rt);
        if (dx != 0.0 || dy != 0.0 || dt != 0.0) {
            const double fx = a[0], fy = a[1], ft = a[2];
            a[5] += 2.0 * dx * fx + dx * dx * e;
            a[6] += dx * fy + dy * fx + dx * dy * e;
            a[7] += 2.0 * dy * fy + dy * dy * e;
            a[8] += dx * ft + dt * fx + dx * dt * e;
            a[9] += dy * ft + dt * fy + dy * dt * e;
            a[10] += 2.0 * dt * ft + dt * dt * e;
            a[0] += dx * e;
            a[1] += dy * e;
            a[2] += dt * e;
        }
        for (int c = 0; c < 11; ++c) raw[c] += a[c];
        if (r.max_logw > bv) {  // ties go to the lowest global index (np.argmax, particle_filter.py:116)
            bv = r.max_logw;
            bg = g;
        }
    }
    const mcl_stats &b = p.rec[bg];
    mcl_stats o = p.rec[0];  // ref, resampled, step, abs_shift: identical on every shard
    o.max_logw = M;
    o.sum_w = s1;
    o.sum_w2 = s2;
    o.ess = s1 * s1 / s2;
    for (int c = 0; c < 11; ++c) o.raw[c] = raw[c];
    o.raw[11] = bv;
    o.max_idx = static_cast<uint32_t>(static_cast<long long>(bg) * p.m + b.max_idx);
    for (int c = 0; c < 3; ++c) o.est[c] = b.est[c];
    o.underflow = 0;
    if (p.apply_rule && weights_underflowed(M, s1, o.abs_shift)) {
        // particle_filter.py:234-237 on the GLOBAL weights: reset to 1/NP; arg-max = the first particle, whose pose
        // shard 0 left in its record's (otherwise host-derived) mean
        const double ng = static_cast<double>(p.m) * p.G;
        o.underflow = 1;
        o.max_logw = 0.0;
        o.sum_w = ng;
        o.sum_w2 = ng;
        o.ess = ng;
        o.raw[11] = 0.0;
        o.max_idx = 0;
        for (int c = 0; c < 3; ++c) o.est[c] = p.rec[0].mean[c];
    }
    o.derived = 0;
    if (derive) derive_stats(o);  // (libm, once per step on one thread: the device-driven step leaves it to the host)
    *p.out = o;
}
__global__ void k_shard_merge(const ShardMergeParams p) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    shard_merge_records(p, true);
}

// -------------------------------------------------------------------------------------------------
// Device-driven exchange of the sharded filter (mcl_shard_step_peer): three single-CTA kernels between the big ones.
// Each has a PUSH half (write into the peers' boxes, then raise the step number there) and a WAIT half (poll the own
// box until every peer's step number has arrived, then derive what the next kernel needs).  A real step launches
// both halves as one kernel; the one-GPU emulation of the tests launches all shards' PUSH halves before any WAIT
// half, on one stream, so nothing ever spins there.
// -------------------------------------------------------------------------------------------------
constexpr int kXchgPush = 1, kXchgWait = 2;

struct XchgParams {
    const PeerView *pv;
    ShardPlan *plan;
    mcl_stats *stats;           // this handle's statistics block (trial 0)
    uint32_t epoch;             // step number + 1 of the step being exchanged
    uint32_t flags;             // host flags of the step
    int mode;                   // kXchgPush | kXchgWait
    // k_xchg_stats
    const double *partials;     // CTA partials of the k_step / k_reduce just run (n_partials = 0: stats[] is current)
    int n_partials;
    StateView cur;
    double ess_threshold;
    // k_xchg_total
    unsigned long long *seg;    // [n_seg] segment totals of k_shard_total -> exclusive prefixes
    int n_seg;
    double ofs_host;            // injected comb offset (have_ofs) ...
    int have_ofs;               // ... else drawn with Philox exactly like the unsharded filter (comb_offset)
    uint64_t seed;
    uint32_t step;
    long long n_global;
    double delta;               // fl(1 / n_global)
};

// statistics: merge the own CTA partials -> push the 312-byte record to every shard -> wait for all records -> merge
// them (all shards compute bit-identical global statistics) -> resampling gate of the next step.
__global__ void __launch_bounds__(kStepThreads) k_xchg_stats(const XchgParams p) {
    const int tid = threadIdx.x;
    const PeerView &pv = *p.pv;
    constexpr int kWords = static_cast<int>(sizeof(mcl_stats) / sizeof(unsigned long long));
    static_assert(sizeof(mcl_stats) % sizeof(unsigned long long) == 0, "mcl_stats is copied in 8-byte words");
    if (p.mode & kXchgPush) {
        if (p.n_partials > 0)
            merge_full<kStepThreads>(p.partials, p.n_partials, p.stats, p.cur.x, p.cur.y, p.cur.th, nullptr, 0.0, false);
        if (tid < 3 && p.cur.x != nullptr) {  // this shard's first particle: the MAP pose of a filter reset to 1/NP
            const float *a = (tid == 0) ? p.cur.x : (tid == 1 ? p.cur.y : p.cur.th);
            p.stats->mean[tid] = static_cast<double>(a[0]);
        }
        __syncthreads();
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(p.stats);
        for (int i = tid; i < pv.G * kWords; i += blockDim.x) {
            const int g = i / kWords, w = i - g * kWords;
            reinterpret_cast<unsigned long long *>(&pv.box[g]->rec[pv.me])[w] = __ldcg(src + w);
        }
        __syncthreads();
        if (tid < pv.G) {
            __threadfence_system();
            st_release_sys(&pv.box[tid]->flag_stats[pv.me], p.epoch);
        }
    }
    if (p.mode & kXchgWait) {
        PeerBox *mine = pv.box[pv.me];
        if (tid < pv.G) peer_wait(&mine->flag_stats[tid], p.epoch, p.plan);
        __syncthreads();
        if (tid == 0) {
            ShardMergeParams mp;
            mp.rec = mine->rec;
            mp.G = pv.G;
            mp.m = pv.m;
            mp.out = p.stats;
            mp.apply_rule = 1;
            shard_merge_records(mp, false);
            p.plan->gate = (p.stats->ess < p.ess_threshold) ? 1 : 0;  // particle_filter.py:211, on the GLOBAL ESS
        }
    }
}

// totals: segment totals -> prefixes (the expansion starts every tile from them), the shard's total -> every shard;
// then, from all totals: k_offset, the comb offset, the slots this shard fills and the part of its own output
// range it gathers locally -- what the host used to derive after a device-to-host copy (sharded.shard_ranges).
__global__ void __launch_bounds__(1024) k_xchg_total(const XchgParams p) {
    __shared__ unsigned long long s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PeerView &pv = *p.pv;
    if (!plan_resample(p.flags, p.plan)) {
        if (tid == 0 && (p.mode & kXchgWait)) {
            p.plan->resampled = 0;
            p.plan->epoch = p.epoch;
        }
        return;
    }
    if (p.mode & kXchgPush) {
        const unsigned long long x = (tid < p.n_seg) ? p.seg[tid] : 0ull;
        unsigned long long incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        unsigned long long off = 0ull, total = 0ull;
        for (int w = 0; w < 32; ++w) {
            if (w < warp) off += s_w[w];
            total += s_w[w];
        }
        if (tid < p.n_seg) p.seg[tid] = off + incl - x;
        if (tid < pv.G) {
            pv.box[tid]->total[pv.me] = total;
            __threadfence_system();
            st_release_sys(&pv.box[tid]->flag_total[pv.me], p.epoch);
        }
    }
    if (p.mode & kXchgWait) {
        PeerBox *mine = pv.box[pv.me];
        if (tid < pv.G) peer_wait(&mine->flag_total[tid], p.epoch, p.plan);
        __syncthreads();
        if (tid == 0) {
            ShardPlan *pl = p.plan;
            double ofs = p.ofs_host;
            if (!p.have_ofs) {  // mcl::comb_offset of the unsharded filter: same seed and step -> same draw on every shard
                const uint4 r = philox4x32_10(make_uint4(0xFFFFFFFFu, 0u, p.step, 0u),
                                              make_uint2(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32)));
                ofs = __dmul_rn(philox_uniform01(r), p.delta);
            }
            const double nd = static_cast<double>(p.n_global);
            unsigned long long k = 0ull, k_me = 0ull;
            long long first = 0, end = 0;
            for (int g = 0; g < pv.G; ++g) {
                const unsigned long long tg = *reinterpret_cast<volatile unsigned long long *>(&mine->total[g]);
                pl->totals[g] = tg;
                if (g == pv.me) {
                    k_me = k;
                    first = static_cast<long long>(count_le(static_cast<double>(k), p.delta, ofs, nd));
                }
                k += tg;
                // the last shard also takes every slot past the final cumulative weight
                if (g == pv.me) end = (g == pv.G - 1) ? p.n_global : static_cast<long long>(count_le(static_cast<double>(k), p.delta, ofs, nd));
            }
            const long long lo = static_cast<long long>(pv.me) * pv.m;
            long long wl = first - lo, wh = end - lo;
            wl = wl < 0 ? 0 : (wl > pv.m ? pv.m : wl);
            wh = wh < 0 ? 0 : (wh > pv.m ? pv.m : wh);
            if (wh < wl) wh = wl;
            pl->k_offset = k_me;
            pl->ofs = ofs;
            pl->first_slot = first;
            pl->n_slots = end - first;
            pl->win_lo = static_cast<uint32_t>(wl);
            pl->win_hi = static_cast<uint32_t>(wh);
            pl->resampled = 1;
            pl->epoch = p.epoch;
        }
    }
}

// done: every pose this shard owed its peers is in their staging (the expansion ended with a system fence) ->
// raise the step number in every box; wait until every peer has done the same for this shard.
__global__ void __launch_bounds__(32) k_xchg_done(const XchgParams p) {
    const int tid = threadIdx.x;
    const PeerView &pv = *p.pv;
    if (!plan_resample(p.flags, p.plan)) return;
    if ((p.mode & kXchgPush) && tid < pv.G) {
        __threadfence_system();
        st_release_sys(&pv.box[tid]->flag_done[pv.me], p.epoch);
    }
    if ((p.mode & kXchgWait) && tid < pv.G) peer_wait(&pv.box[pv.me]->flag_done[tid], p.epoch, p.plan);
}

// Sharded filter, first pass of the global resampling: the shard's total of the quantised weights (atomicAdd into
// *total).  The tiles are those of k_scan; every CTA takes one contiguous segment of them and leaves, tile by
// tile, the weight that precedes the tile inside the segment (tile_prefix) and the segment's total (seg_total).
// k_tile_prefix turns the segment totals into prefixes, so the second pass (k_scan<kScanExpand>) starts every
// tile at seg_prefix[segment] + tile_prefix[tile] -- no look-back, no descriptors.
constexpr int kMaxSegments = 1024;
__global__ void __launch_bounds__(kScanThreads) k_shard_total(const float *lw, long long n, const mcl_stats *stats,
                                                              unsigned long long *total, unsigned long long *tile_prefix,
                                                              unsigned long long *seg_total, int tpt, int tiles_per_seg,
                                                              const ShardPlan *plan, uint32_t flags) {
    __shared__ WeightMap s_wm;
    __shared__ unsigned long long s_part[2][kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (plan != nullptr && !plan_resample(flags, plan)) return;  // device-driven step: gated on the device
    if (tid == 0) s_wm = make_weight_map(stats[0].max_logw, stats[0].sum_w);
    __syncthreads();
    const WeightMap wm = s_wm;
    const bool uniform = stats[0].underflow != 0;  // (global statistics: every shard takes the same branch)
    unsigned long long run = 0ull;  // thread 0: weight of the segment's tiles so far
    const long long t0 = static_cast<long long>(blockIdx.x) * tiles_per_seg;
    const long long t1 = min(static_cast<long long>(tpt), t0 + tiles_per_seg);
    // The loads of the NEXT tile are issued before the current one is reduced: a tile is 64 bytes per thread, and with
    // one tile in flight per CTA the pass ran at 3.9 TB/s, bound by the load -> reduce -> barrier chain of each tile.
    float4 cur4[kScanItems / 4], nxt4[kScanItems / 4];
    auto load_tile = [&](long long tile, float4 (&dst)[kScanItems / 4]) {
        const long long j0 = tile * kScanTile + static_cast<long long>(tid) * kScanItems;
#pragma unroll
        for (int c = 0; c < kScanItems / 4; ++c)  // the state arrays are padded to a multiple of 128 elements: in bounds
            dst[c] = (tile < t1 && j0 < n) ? *reinterpret_cast<const float4 *>(lw + j0 + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    load_tile(t0, cur4);
    int par = 0;
    for (long long tile = t0; tile < t1; ++tile, par ^= 1) {
        load_tile(tile + 1, nxt4);
        const long long j0 = tile * kScanTile + static_cast<long long>(tid) * kScanItems;
        unsigned long long acc = 0ull;
        if (j0 < n) {
#pragma unroll
            for (int c = 0; c < kScanItems / 4; ++c) {
                float4 a = cur4[c];
                if (uniform) a = make_float4(0.f, 0.f, 0.f, 0.f);
                const float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (j0 + 4 * c + k < n) acc += quantize_weight(v[k], wm);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_part[par][warp] = acc;
        __syncthreads();  // one barrier per tile: the two halves of s_part alternate
        if (tid == 0) {
            unsigned long long s = 0ull;
#pragma unroll
            for (int w = 0; w < kScanThreads / 32; ++w) s += s_part[par][w];
            tile_prefix[tile] = run;
            run += s;
        }
#pragma unroll
        for (int c = 0; c < kScanItems / 4; ++c) cur4[c] = nxt4[c];
    }
    if (tid == 0) {
        seg_total[blockIdx.x] = run;
        if (run != 0ull && total != nullptr) atomicAdd(total, run);
    }
}

// in-place exclusive prefix of v[0..n), n <= 1024: one CTA, one element per thread
__global__ void __launch_bounds__(1024) k_tile_prefix(unsigned long long *v, int n) {
    __shared__ unsigned long long s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long x = (tid < n) ? v[tid] : 0ull;
    unsigned long long incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned long long off = 0ull;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    if (tid < n) v[tid] = off + incl - x;
}

// normalised weights w = q 2^-52 (fp64) and the integers q the resampler scans
__global__ void k_export_weights(const float *lw, const mcl_stats *stats, double *w, unsigned long long *q, long long n,
                                 long long ld) {
    const int t = blockIdx.y;
    const WeightMap wm = make_weight_map(stats[t].max_logw, stats[t].sum_w);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float v = stats[t].underflow ? 0.0f : lw[t * ld + i];  // reset to 1/NP (:234-237)
        if (w) w[t * n + i] = static_cast<double>(weight_scaled(v, wm)) * kInvTwo52;
        if (q) q[t * n + i] = quantize_weight(v, wm);
    }
}

}  // namespace mcl
