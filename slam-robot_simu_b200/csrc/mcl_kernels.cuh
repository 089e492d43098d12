// Kernels of the MCL step (sm_100a).  One translation unit: included by mcl_api.cu.
//
//   k_prep            1 block / trial : per-step scalars (resample gate, comb offset, reference pose,
//                                       landmark table + moments, likelihood shift), resets tickets
//   k_scan_ancestors  decoupled look-back scan of 64-bit fixed-point weights -> ancestor of every slot
//   k_step            [gather] + predict + weight + block partials + last-block finalise (stats)
//   k_gather          stand-alone resampling gather (stage API)
//   k_cov             exact two-pass weighted covariance
//   k_export_*        weights / quantised weights for the host
#pragma once
#include "mcl_device.cuh"

namespace mcl {

struct StateView {
    float *x, *y, *th, *lw;
};

// -------------------------------------------------------------------------------------------------
// k_prep
// -------------------------------------------------------------------------------------------------
struct PrepParams {
    const double *lm;       // [L][2]
    const double *z;        // [T][L][2]  (may be null in reduce-only mode)
    const double *ofs;      // [T] or null
    const mcl_stats *stats; // [T] previous step
    PrepOut *prep;          // [T]
    double4 *tab;           // [T][Lmax]
    unsigned long long *desc;  // [T][tpt] scan tile descriptors
    unsigned int *tickets;  // [T][4]
    StateView cur;          // for ref-from-first-particle mode
    long long ld;
    int L, Lmax, tpt;
    uint32_t flags;
    int likelihood;
    int ref_from_particle0; // reduce-only after set_particles
    int reduce_only;        // no motion this pass: the reference pose is the stored estimate itself
    double ess_threshold, vdt, wdt, inv_s0sq, inv_s1sq, lik_const, base_lw, inv_n;
    uint64_t seed;
    uint32_t step;
};

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

__global__ void __launch_bounds__(128) k_prep(const PrepParams p) {
    const int t = blockIdx.x, tid = threadIdx.x;
    __shared__ double red[4 * 16];
    const mcl_stats st = p.stats[t];
    PrepOut *po = p.prep + t;

    // ---- reference pose: previous MAP estimate pushed through the noise-free motion (__f) ----
    double rx, ry, rth;
    if (p.ref_from_particle0) {
        rx = p.cur.x[t * p.ld];
        ry = p.cur.y[t * p.ld];
        rth = p.cur.th[t * p.ld];
    } else if (p.reduce_only) {
        rx = st.est[0];
        ry = st.est[1];
        rth = st.est[2];
    } else {
        double s0, c0;
        sincos(st.est[2], &s0, &c0);
        rx = st.est[0] + p.vdt * c0;
        ry = st.est[1] + p.vdt * s0;
        rth = limit_angle(st.est[2] + p.wdt);
    }

    // ---- landmark table, moments, likelihood of the reference pose ----
    double acc[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) acc[k] = 0.0;
    if (p.z != nullptr) {
        double s, c;
        sincos(rth, &s, &c);
        const double tx = -(s * rx - c * ry), ty = -(c * rx + s * ry);
        for (int l = tid; l < p.L; l += blockDim.x) {
            const double lx = p.lm[2 * l], ly = p.lm[2 * l + 1];
            const double zx = p.z[(static_cast<long long>(t) * p.L + l) * 2];
            const double zy = p.z[(static_cast<long long>(t) * p.L + l) * 2 + 1];
            const double4 e = make_double4(lx, ly, zx, zy);
            p.tab[static_cast<long long>(t) * p.Lmax + l] = e;
            acc[0] += lx;       acc[1] += ly;       acc[2] += zx;       acc[3] += zy;
            acc[4] += lx * lx;  acc[5] += ly * ly;  acc[6] += lx * ly;
            acc[7] += zx * zx;  acc[8] += zy * zy;
            acc[9] += lx * zx;  acc[10] += ly * zx; acc[11] += lx * zy; acc[12] += ly * zy;
            double e0, e1;
            if (p.likelihood == MCL_LIK_RB) pair_residual_rb(s, c, tx, ty, e, e0, e1);
            else pair_residual_c(s, c, tx, ty, e, e0, e1);
            acc[13] = fma(e0, e0, acc[13]);
            acc[14] = fma(e1, e1, acc[14]);
        }
    }
    block_sum<15>(acc, red);

    // ---- reset the scan tile descriptors and tickets of this trial ----
    for (int i = tid; i < p.tpt; i += blockDim.x) p.desc[static_cast<long long>(t) * p.tpt + i] = 0ull;
    if (tid < 4) p.tickets[t * 4 + tid] = 0u;

    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 13; ++k) po->mom[k] = acc[k];
        po->c_shift = (p.z != nullptr) ? (-0.5 * (acc[13] * p.inv_s0sq + acc[14] * p.inv_s1sq) + p.lik_const) : 0.0;
        po->ref[0] = rx;
        po->ref[1] = ry;
        po->ref[2] = rth;
        po->log_norm_prev = st.log_norm;
        po->lik_const = p.lik_const;
        po->base_lw = p.base_lw;
        po->L = p.L;
        int rs;
        if (p.flags & MCL_NO_RESAMPLE) rs = 0;
        else if (p.flags & MCL_FORCE_RESAMPLE) rs = 1;
        else rs = (st.ess < p.ess_threshold) ? 1 : 0;  // particle_filter.py:211
        po->do_resample = rs;
        double ofs = 0.0;
        if (p.ofs != nullptr) ofs = p.ofs[t];
        else if (p.flags & MCL_PHILOX_NOISE) {
            const uint4 r = philox4x32_10(make_uint4(0xFFFFFFFFu, 0u, p.step, static_cast<uint32_t>(t)),
                                          make_uint2(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32)));
            ofs = __dmul_rn(philox_uniform01(r), p.inv_n);  // rand() * NP_RECIP, particle_filter.py:214
        }
        po->ofs = ofs;
    }
}

// -------------------------------------------------------------------------------------------------
// k_scan_ancestors : single-pass chained scan (decoupled look-back) over uint64 fixed-point weights.
// Tile descriptor = status (2 bits) << 62 | value, one 64-bit word, so status and value are read
// atomically.  Tiles are handed out by an atomic ticket so a tile's predecessors are always running.
// -------------------------------------------------------------------------------------------------
struct ScanParams {
    const float *lw;            // [T][ld]
    const PrepOut *prep;        // [T]
    unsigned long long *desc;   // [T][tpt]
    unsigned int *tickets;      // [T][4]  (slot 0 = scan ticket)
    uint32_t *anc;              // [T][ld] (slot - slot_first)
    unsigned long long *total;  // optional [T] : inclusive total of the local weights
    long long n, ld, n_global;  // local sources, stride, comb length
    long long j_global_base;    // global index of local source 0 (sharded)
    long long slot_first;       // anc is indexed by (global slot - slot_first)
    long long anc_cap;          // capacity of anc per trial
    unsigned long long k_offset;  // sum of the quantised weights of all lower-ranked shards
    double delta;               // fl(1 / n_global)
    int tpt;
    int mode;                   // kScanAncestors / kScanTotals / kScanExpand
    int use_override;           // 1: take log_norm / ofs from the two fields below, always resample
    double log_norm_override, ofs_override;
    StateView src;              // kScanExpand: poses of the local sources ...
    float *ox, *oy, *oth;       // ... copied to out[slot - slot_first] for every slot they fill
};

constexpr int kScanAncestors = 0;  // write the ancestor of every slot
constexpr int kScanTotals = 1;     // only the inclusive total of the quantised weights
constexpr int kScanExpand = 2;     // write the resampled pose of every slot (sender side of the sharded filter)
constexpr int kShortRun = 16;      // offspring runs up to this length are written by the owning thread
constexpr int kRunQueue = 128;     // longer runs are queued and written by the whole block

enum : unsigned long long {
    kDescAggregate = 1ull << 62,
    kDescPrefix = 2ull << 62,
    kDescValueMask = (1ull << 62) - 1ull
};

template <int MODE>
__global__ void __launch_bounds__(kScanThreads) k_scan_ancestors(const ScanParams p) {
    const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PrepOut *po = p.prep + t;
    if (!p.use_override && !po->do_resample) return;
    const double log_norm = p.use_override ? p.log_norm_override : po->log_norm_prev;
    const double ofs = p.use_override ? p.ofs_override : po->ofs;

    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    __shared__ unsigned long long s_excl;
    if (tid == 0) s_tile = atomicAdd(&p.tickets[t * 4 + 0], 1u);
    __syncthreads();
    const long long tile = s_tile;
    const long long j0 = tile * kScanTile + static_cast<long long>(tid) * kScanItems;
    const float *lw = p.lw + static_cast<long long>(t) * p.ld;

    // ---- load + quantise (8 consecutive items per thread, two 16-byte loads) ----
    unsigned long long q[kScanItems];
    {
        float v[kScanItems];
        if (j0 + kScanItems <= p.n) {
            const float4 a = *reinterpret_cast<const float4 *>(lw + j0);
            const float4 b = *reinterpret_cast<const float4 *>(lw + j0 + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
            for (int k = 0; k < kScanItems; ++k) q[k] = quantize_weight(v[k], log_norm);
        } else {
#pragma unroll
            for (int k = 0; k < kScanItems; ++k) q[k] = (j0 + k < p.n) ? quantize_weight(lw[j0 + k], log_norm) : 0ull;
        }
    }
    unsigned long long tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) tsum += q[k];

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
    for (int w = 0; w < kScanThreads / 32; ++w) {
        const unsigned long long v = s_warp[w];
        if (w < warp) warp_off += v;
        tile_total += v;
    }
    const unsigned long long thread_excl = warp_off + incl - tsum;

    // ---- decoupled look-back (warp 0) ----
    unsigned long long *desc = p.desc + static_cast<long long>(t) * p.tpt;
    if (warp == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&desc[0], kDescPrefix | tile_total);
        } else {
            if (lane == 0) atomicExch(&desc[tile], kDescAggregate | tile_total);
            long long look = tile - 1;  // window covers tiles [look-31, look]
            while (true) {
                const long long idx = look - lane;
                unsigned long long d;
                do {
                    d = (idx >= 0) ? *reinterpret_cast<volatile unsigned long long *>(&desc[idx]) : kDescPrefix;
                } while (__any_sync(0xffffffffu, (d >> 62) == 0ull));
                const unsigned int is_prefix = __ballot_sync(0xffffffffu, (d >> 62) == 2ull);
                const int first = is_prefix ? (__ffs(is_prefix) - 1) : 32;  // nearest predecessor holding a prefix
                unsigned long long contrib = (lane <= first) ? (d & kDescValueMask) : 0ull;
                if (idx < 0) contrib = 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                excl += contrib;
                if (is_prefix) break;
                look -= 32;
            }
            if (lane == 0) atomicExch(&desc[tile], kDescPrefix | (excl + tile_total));
        }
        if (lane == 0) {
            s_excl = excl;
            if (p.total != nullptr && tile == p.tpt - 1) p.total[t] = excl + tile_total;
        }
    }
    __syncthreads();
    if (MODE == kScanTotals) return;

    // ---- ancestors: source j owns slots [count_le(K_{j-1}), count_le(K_j)) ----
    __shared__ long long s_run_lo[kRunQueue], s_run_hi[kRunQueue];
    __shared__ unsigned int s_run_src[kRunQueue];
    __shared__ unsigned int s_nrun;
    if (tid == 0) s_nrun = 0u;
    __syncthreads();
    unsigned long long K = p.k_offset + s_excl + thread_excl;
    const double nd = static_cast<double>(p.n_global);
    uint32_t *anc = p.anc + static_cast<long long>(t) * p.anc_cap;
    auto emit = [&](long long o, uint32_t j, float px, float py, float pth) {
        if (o >= 0 && o < p.anc_cap) {
            if (MODE == kScanExpand) {
                p.ox[o] = px;
                p.oy[o] = py;
                p.oth[o] = pth;
            } else {
                anc[o] = j;
            }
        }
    };
    if (j0 < p.n) {
        long long start = count_le(K, p.delta, ofs, p.n_global, nd);
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const long long j = j0 + k;
            if (j < p.n) {
                K += q[k];
                // the last particle of the whole filter also takes every slot past the final cumulative
                // weight (the reference would raise IndexError there)
                const long long end =
                    (p.j_global_base + j == p.n_global - 1) ? p.n_global : count_le(K, p.delta, ofs, p.n_global, nd);
                if (end > start) {
                    unsigned int slot = kRunQueue;
                    if (end - start > kShortRun) slot = atomicAdd(&s_nrun, 1u);
                    if (slot < kRunQueue) {
                        s_run_lo[slot] = start;
                        s_run_hi[slot] = end;
                        s_run_src[slot] = static_cast<uint32_t>(j);
                    } else {
                        float px = 0.f, py = 0.f, pth = 0.f;
                        if (MODE == kScanExpand) {
                            px = p.src.x[j];
                            py = p.src.y[j];
                            pth = p.src.th[j];
                        }
                        for (long long s = start; s < end; ++s) emit(s - p.slot_first, static_cast<uint32_t>(j), px, py, pth);
                    }
                }
                start = end;
            }
        }
    }
    __syncthreads();
    const unsigned int nrun = min(s_nrun, static_cast<unsigned int>(kRunQueue));
    for (unsigned int r = 0; r < nrun; ++r) {
        const long long lo = s_run_lo[r], hi = s_run_hi[r];
        const uint32_t j = s_run_src[r];
        float px = 0.f, py = 0.f, pth = 0.f;
        if (MODE == kScanExpand) {
            px = p.src.x[j];
            py = p.src.y[j];
            pth = p.src.th[j];
        }
        for (long long s = lo + tid; s < hi; s += kScanThreads) emit(s - p.slot_first, j, px, py, pth);
    }
}

// -------------------------------------------------------------------------------------------------
// k_step : [gather] + predict + weight + reductions.
// MODE 0 = full step; MODE 1 = reduce only (statistics of the current state, nothing written back)
// -------------------------------------------------------------------------------------------------
struct StepParams {
    StateView in, out;
    const uint32_t *anc;   // [T][ld]
    const float *noise;    // [3][T][n] dense, or null
    const PrepOut *prep;   // [T]
    const double4 *tab;    // [T][Lmax]
    double *partials;      // [T][bpt][kPart]
    unsigned int *tickets; // [T][4] (slot 1 = reduce ticket)
    mcl_stats *stats;      // [T]
    long long n, ld;
    long long slot_base;   // global index of local particle 0 (Philox counters)
    int T, bpt, Lmax;
    int likelihood, motion;
    uint32_t flags, step;
    uint64_t seed;
    double vdt, wdt;       // VEL*dt, YAW_RATE*dt
    double vel, yaw_rate, dt;
    double q_std[3];
    double inv_s0sq, inv_s1sq;
    double mb_sig[3];      // motion B: the three "sigma ** 2" scales (motion_model.py:43-48)
};

template <int MODE>
__global__ void __launch_bounds__(kStepThreads, 2) k_step(const StepParams p) {
    const int t = blockIdx.y, tid = threadIdx.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double4 *s_tab = reinterpret_cast<double4 *>(smem_raw);                    // [L]
    __shared__ double s_red[(kStepThreads / 32) * kPart];
    __shared__ double s_prep[24];
    __shared__ int s_last;

    const PrepOut *po = p.prep + t;
    const int L = po->L;
    if (MODE == 0 && p.likelihood != MCL_LIK_MOMENT) {
        const double4 *g = p.tab + static_cast<long long>(t) * p.Lmax;
        for (int l = tid; l < L; l += blockDim.x) s_tab[l] = g[l];
    }
    if (tid < 13) s_prep[tid] = po->mom[tid];
    if (tid == 13) s_prep[13] = po->c_shift;
    if (tid >= 14 && tid < 17) s_prep[tid] = po->ref[tid - 14];
    if (tid == 17) s_prep[17] = po->log_norm_prev;
    if (tid == 18) s_prep[18] = po->lik_const;
    if (tid == 19) s_prep[19] = po->base_lw;
    const int resample = (MODE == 0) ? po->do_resample : 0;
    __syncthreads();
    const double c_shift = s_prep[13], ref_x = s_prep[14], ref_y = s_prep[15], ref_th = s_prep[16];
    const double log_norm_prev = s_prep[17], lik_const = s_prep[18], base_lw = s_prep[19];

    const long long off = static_cast<long long>(t) * p.ld;
    const float *ix = p.in.x + off, *iy = p.in.y + off, *ith = p.in.th + off, *ilw = p.in.lw + off;
    float *ox = p.out.x + off, *oy = p.out.y + off, *oth = p.out.th + off, *olw = p.out.lw + off;
    const uint32_t *anc = p.anc + off;
    const uint2 key = make_uint2(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32));

    // running (max, sums) of this thread
    double m = -INFINITY;
    double acc[kNumSums];
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) acc[k] = 0.0;
    float best_v = -INFINITY;
    long long best_i = 0x7fffffffffffffffll;

    for (long long base = (static_cast<long long>(blockIdx.x) * kStepThreads + tid) * kPPT; base < p.n;
         base += static_cast<long long>(gridDim.x) * kStepTile) {
        float x[kPPT], y[kPPT], th[kPPT], lwp[kPPT] = {0.f, 0.f, 0.f, 0.f};
        // ---------------- load (gather through the ancestors when this step resamples) ----------------
        if (resample) {
            const uint4 a4 = *reinterpret_cast<const uint4 *>(anc + base);
            const uint32_t a[kPPT] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int k = 0; k < kPPT; ++k) {
                const uint32_t j = (base + k < p.n) ? a[k] : 0u;
                x[k] = ix[j];
                y[k] = iy[j];
                th[k] = ith[j];
            }
        } else {
            const float4 vx = *reinterpret_cast<const float4 *>(ix + base);
            const float4 vy = *reinterpret_cast<const float4 *>(iy + base);
            const float4 vt = *reinterpret_cast<const float4 *>(ith + base);
            const float4 vw = *reinterpret_cast<const float4 *>(ilw + base);
            x[0] = vx.x; x[1] = vx.y; x[2] = vx.z; x[3] = vx.w;
            y[0] = vy.x; y[1] = vy.y; y[2] = vy.z; y[3] = vy.w;
            th[0] = vt.x; th[1] = vt.y; th[2] = vt.z; th[3] = vt.w;
            lwp[0] = vw.x; lwp[1] = vw.y; lwp[2] = vw.z; lwp[3] = vw.w;
        }

        double lw_new[kPPT];
        if (MODE == 0) {
            // ---------------- motion noise ----------------
            float nz[kPPT][3];
            if (p.noise != nullptr) {
                const long long plane = static_cast<long long>(p.T) * p.n;
                const float *nb = p.noise + static_cast<long long>(t) * p.n;
#pragma unroll
                for (int k = 0; k < kPPT; ++k) {
                    const bool ok = base + k < p.n;
#pragma unroll
                    for (int d = 0; d < 3; ++d) nz[k][d] = ok ? nb[d * plane + base + k] : 0.0f;
                }
            } else {
                // 3 Philox calls give the 12 normals of this thread's 4 particles
                const unsigned long long grp = static_cast<unsigned long long>(p.slot_base + base) >> 2;
                float nrm[12];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(grp), static_cast<uint32_t>(grp >> 32) * 4u + c,
                                                             p.step, static_cast<uint32_t>(t)), key);
                    const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
                    nrm[4 * c + 0] = a.x; nrm[4 * c + 1] = a.y; nrm[4 * c + 2] = b.x; nrm[4 * c + 3] = b.y;
                }
#pragma unroll
                for (int k = 0; k < kPPT; ++k)
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float sc = (p.motion == MCL_MOTION_A) ? static_cast<float>(p.q_std[d]) : 1.0f;
                        nz[k][d] = nrm[3 * k + d] * sc;
                    }
            }

            // ---------------- predict ----------------
#pragma unroll
            for (int k = 0; k < kPPT; ++k) {
                const double thd = static_cast<double>(th[k]);
                double xn, yn, thn;
                if (p.motion == MCL_MOTION_A) {
                    // particle_filter.py:129-142 then "+ v" (:166): cos/sin of the PRE-update yaw; the wrap is
                    // applied before the noise, so the stored yaw may leave [-pi, pi]
                    float s0, c0;
                    sincosf(th[k], &s0, &c0);
                    xn = static_cast<double>(x[k]) + p.vdt * static_cast<double>(c0) + static_cast<double>(nz[k][0]);
                    yn = static_cast<double>(y[k]) + p.vdt * static_cast<double>(s0) + static_cast<double>(nz[k][1]);
                    thn = limit_angle(thd + p.wdt) + static_cast<double>(nz[k][2]);
                } else {
                    // motion_model.py:46-56; sigma**2 used as the standard deviation (reference quirk)
                    const double vh = p.vel + static_cast<double>(nz[k][0]) * p.mb_sig[0];
                    const double wh = p.yaw_rate + static_cast<double>(nz[k][1]) * p.mb_sig[1];
                    const double gh = static_cast<double>(nz[k][2]) * p.mb_sig[2];
                    const double a = vh / wh, b = wh * p.dt;
                    double s0, c0, s1, c1;
                    sincos(thd, &s0, &c0);
                    sincos(thd + b, &s1, &c1);
                    xn = static_cast<double>(x[k]) - a * s0 + a * s1;
                    yn = static_cast<double>(y[k]) + a * c0 - a * c1;
                    thn = limit_angle(thd + (wh + gh) * p.dt);
                }
                x[k] = static_cast<float>(xn);
                y[k] = static_cast<float>(yn);
                th[k] = static_cast<float>(thn);
            }

            // ---------------- weight: log-likelihood of the STORED (fp32) pose, fp64 arithmetic ----------------
            double s[kPPT], c[kPPT], tx[kPPT], ty[kPPT];
#pragma unroll
            for (int k = 0; k < kPPT; ++k) {
                sincos(static_cast<double>(th[k]), &s[k], &c[k]);
                const double xd = x[k], yd = y[k];
                tx[k] = -(s[k] * xd - c[k] * yd);
                ty[k] = -(c[k] * xd + s[k] * yd);
            }
            double q0[kPPT], q1[kPPT];
            if (p.likelihood == MCL_LIK_MOMENT) {
                // sum_l ex^2 = S(aa) - 2 P1 S(a) + L P1^2 with a_l = s lx - c ly - zx, P1 = s x - c y (= -tx)
                const double Slx = s_prep[0], Sly = s_prep[1], Szx = s_prep[2], Szy = s_prep[3];
                const double Sxx = s_prep[4], Syy = s_prep[5], Sxy = s_prep[6], Szxzx = s_prep[7], Szyzy = s_prep[8];
                const double Slxzx = s_prep[9], Slyzx = s_prep[10], Slxzy = s_prep[11], Slyzy = s_prep[12];
                const double Ld = static_cast<double>(L);
#pragma unroll
                for (int k = 0; k < kPPT; ++k) {
                    const double ss = s[k] * s[k], cc = c[k] * c[k], sc = s[k] * c[k];
                    const double Sa = s[k] * Slx - c[k] * Sly - Szx;
                    const double Saa = ss * Sxx + cc * Syy - 2.0 * sc * Sxy + Szxzx - 2.0 * s[k] * Slxzx + 2.0 * c[k] * Slyzx;
                    const double Sb = c[k] * Slx + s[k] * Sly - Szy;
                    const double Sbb = cc * Sxx + ss * Syy + 2.0 * sc * Sxy + Szyzy - 2.0 * c[k] * Slxzy - 2.0 * s[k] * Slyzy;
                    q0[k] = Saa + 2.0 * tx[k] * Sa + Ld * tx[k] * tx[k];
                    q1[k] = Sbb + 2.0 * ty[k] * Sb + Ld * ty[k] * ty[k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < kPPT; ++k) q0[k] = q1[k] = 0.0;
                if (p.likelihood == MCL_LIK_PAIR) {
#pragma unroll 4
                    for (int l = 0; l < L; ++l) {
                        const double4 tb = s_tab[l];
#pragma unroll
                        for (int k = 0; k < kPPT; ++k) {
                            double ex, ey;
                            pair_residual_c(s[k], c[k], tx[k], ty[k], tb, ex, ey);
                            q0[k] = fma(ex, ex, q0[k]);
                            q1[k] = fma(ey, ey, q1[k]);
                        }
                    }
                } else {
                    for (int l = 0; l < L; ++l) {
                        const double4 tb = s_tab[l];
#pragma unroll
                        for (int k = 0; k < kPPT; ++k) {
                            double ed, ephi;
                            pair_residual_rb(s[k], c[k], tx[k], ty[k], tb, ed, ephi);
                            q0[k] = fma(ed, ed, q0[k]);
                            q1[k] = fma(ephi, ephi, q1[k]);
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kPPT; ++k) {
                const double ll = -0.5 * (q0[k] * p.inv_s0sq + q1[k] * p.inv_s1sq) + lik_const;
                const double prev = resample ? base_lw : (static_cast<double>(lwp[k]) - log_norm_prev);
                lw_new[k] = prev + (ll - c_shift);  // pw * bn (:194), in logs, shifted by a per-step constant
            }
            // ---------------- store ----------------
            float lwf[kPPT];
#pragma unroll
            for (int k = 0; k < kPPT; ++k) lwf[k] = static_cast<float>(lw_new[k]);
            *reinterpret_cast<float4 *>(ox + base) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4 *>(oy + base) = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4 *>(oth + base) = make_float4(th[0], th[1], th[2], th[3]);
            *reinterpret_cast<float4 *>(olw + base) = make_float4(lwf[0], lwf[1], lwf[2], lwf[3]);
#pragma unroll
            for (int k = 0; k < kPPT; ++k) lw_new[k] = static_cast<double>(lwf[k]);  // statistics of what is stored
        } else {
#pragma unroll
            for (int k = 0; k < kPPT; ++k) lw_new[k] = static_cast<double>(lwp[k]);
        }

        // ---------------- running statistics (online max + rescaled sums) ----------------
#pragma unroll
        for (int k = 0; k < kPPT; ++k) {
            if (base + k >= p.n) continue;
            const double lw = lw_new[k];
            const float lwf = static_cast<float>(lw);
            if (lwf > best_v) {  // strict: first occurrence wins (np.argmax, particle_filter.py:116)
                best_v = lwf;
                best_i = base + k;
            }
            double e;
            if (lw > m) {
                const double r = exp(m - lw);  // exp(-inf) = 0 on the first particle
#pragma unroll
                for (int a = 0; a < kNumSums; ++a) acc[a] *= (a == 1) ? r * r : r;  // s2 carries exp(2 (lw - m))
                m = lw;
                e = 1.0;
            } else {
                e = exp(lw - m);
            }
            const double dx = static_cast<double>(x[k]) - ref_x, dy = static_cast<double>(y[k]) - ref_y;
            const double thd = static_cast<double>(th[k]);
            const double dth = limit_angle(thd - ref_th);
            double sn, cs;
            sincos(thd, &sn, &cs);
            acc[0] += e;
            acc[1] = fma(e, e, acc[1]);
            const double ex = e * dx, ey = e * dy, et = e * dth;
            acc[2] += ex;  acc[3] += ey;  acc[4] += et;
            acc[5] = fma(e, sn, acc[5]);
            acc[6] = fma(e, cs, acc[6]);
            acc[7] = fma(ex, dx, acc[7]);
            acc[8] = fma(ex, dy, acc[8]);
            acc[9] = fma(ey, dy, acc[9]);
            acc[10] = fma(ex, dth, acc[10]);
            acc[11] = fma(ey, dth, acc[11]);
            acc[12] = fma(et, dth, acc[12]);
        }
    }

    // ---------------- block partial ----------------
    const double M = block_max(m, s_red);
    {
        const double r = (m == -INFINITY) ? 0.0 : exp(m - M);
        // s2 carries exp(2 (lw - m)): rescale with r^2
#pragma unroll
        for (int a = 0; a < kNumSums; ++a) acc[a] *= (a == 1) ? r * r : r;
    }
    block_sum<kNumSums>(acc, s_red);
    // argmax across the block
    {
        float bv = best_v;
        long long bi = best_i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        __shared__ float s_bv[kStepThreads / 32];
        __shared__ long long s_bi[kStepThreads / 32];
        if ((tid & 31) == 0) { s_bv[tid >> 5] = bv; s_bi[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kStepThreads / 32; ++w)
                if (s_bv[w] > bv || (s_bv[w] == bv && s_bi[w] < bi)) { bv = s_bv[w]; bi = s_bi[w]; }
            double *pp = p.partials + (static_cast<long long>(t) * p.bpt + blockIdx.x) * kPart;
            pp[0] = M;
#pragma unroll
            for (int a = 0; a < kNumSums; ++a) pp[1 + a] = acc[a];
            pp[14] = static_cast<double>(bv);
            pp[15] = static_cast<double>(bi);
            __threadfence();
            s_last = (atomicAdd(&p.tickets[t * 4 + 1], 1u) == static_cast<unsigned int>(gridDim.x) - 1u);
        }
        __syncthreads();
    }
    if (!s_last) return;
    __threadfence();

    // ---------------- last block of the trial: merge partials -> stats ----------------
    const double *pp = p.partials + static_cast<long long>(t) * p.bpt * kPart;
    double mm = -INFINITY;
    for (int b = tid; b < static_cast<int>(gridDim.x); b += blockDim.x) mm = fmax(mm, __ldcg(pp + b * kPart));
    const double MM = block_max(mm, s_red);
    double tot[kNumSums];
#pragma unroll
    for (int a = 0; a < kNumSums; ++a) tot[a] = 0.0;
    float bv = -INFINITY;
    long long bi = 0x7fffffffffffffffll;
    for (int b = tid; b < static_cast<int>(gridDim.x); b += blockDim.x) {
        const double *q = pp + b * kPart;
        const double mb = __ldcg(q);
        const double r = (mb == -INFINITY) ? 0.0 : exp(mb - MM);
#pragma unroll
        for (int a = 0; a < kNumSums; ++a) tot[a] += __ldcg(q + 1 + a) * ((a == 1) ? r * r : r);
        const float v = static_cast<float>(__ldcg(q + 14));
        const long long i = static_cast<long long>(__ldcg(q + 15));
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
    block_sum<kNumSums>(tot, s_red);
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        __shared__ float s_bv2[kStepThreads / 32];
        __shared__ long long s_bi2[kStepThreads / 32];
        if ((tid & 31) == 0) { s_bv2[tid >> 5] = bv; s_bi2[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kStepThreads / 32; ++w)
                if (s_bv2[w] > bv || (s_bv2[w] == bv && s_bi2[w] < bi)) { bv = s_bv2[w]; bi = s_bi2[w]; }
            mcl_stats *st = p.stats + t;
            const double s1 = tot[0], s2 = tot[1];
            st->max_logw = MM;
            st->sum_w = s1;
            st->sum_w2 = s2;
            st->log_norm = MM + log(s1);
            st->ess = s1 * s1 / s2;  // 1 / sum (w/s1)^2, particle_filter.py:210
            st->max_weight = exp(static_cast<double>(bv) - MM) / s1;
            st->max_idx = bi;
            const float *ex = (MODE == 0) ? ox : ix, *ey = (MODE == 0) ? oy : iy, *eth = (MODE == 0) ? oth : ith;
            st->est[0] = __ldcg(ex + bi);
            st->est[1] = __ldcg(ey + bi);
            st->est[2] = __ldcg(eth + bi);
#pragma unroll
            for (int a = 0; a < 11; ++a) st->raw[a] = tot[2 + a];
            st->raw[11] = 0.0;
            st->ref[0] = ref_x; st->ref[1] = ref_y; st->ref[2] = ref_th;
            // mean / covariance about the reference pose (single pass)
            const double inv = 1.0 / s1;
            const double mdx = tot[2] * inv, mdy = tot[3] * inv, mdt = tot[4] * inv;
            const double mth = atan2(tot[5], tot[6]);            // circular mean of the yaw
            const double del = limit_angle(mth - ref_th);         // its offset from the reference
            st->mean[0] = ref_x + mdx;
            st->mean[1] = ref_y + mdy;
            st->mean[2] = mth;
            const double cxx = tot[7] * inv - mdx * mdx, cxy = tot[8] * inv - mdx * mdy, cyy = tot[9] * inv - mdy * mdy;
            const double cxt = tot[10] * inv - mdx * mdt, cyt = tot[11] * inv - mdy * mdt;
            const double ctt = tot[12] * inv - 2.0 * del * mdt + del * del;
            st->cov[0] = cxx; st->cov[1] = cxy; st->cov[2] = cxt;
            st->cov[3] = cxy; st->cov[4] = cyy; st->cov[5] = cyt;
            st->cov[6] = cxt; st->cov[7] = cyt; st->cov[8] = ctt;
            if (MODE == 0) {
                st->resampled = resample;
                st->step = static_cast<int32_t>(p.step + 1u);
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// k_gather : stand-alone __resampling (particle_filter.py:216-222): px[:, i] = px_temp[:, idx], pw = 1/NP
// -------------------------------------------------------------------------------------------------
struct GatherParams {
    StateView in, out;
    const uint32_t *anc;
    const PrepOut *prep;
    mcl_stats *stats;
    long long n, ld;
    double n_global;
};

__global__ void __launch_bounds__(256) k_gather(const GatherParams p) {
    const int t = blockIdx.y;
    const PrepOut *po = p.prep + t;
    const int rs = po->do_resample;
    const long long off = static_cast<long long>(t) * p.ld;
    const float base_lw = static_cast<float>(po->base_lw);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        if (rs) {
            const uint32_t j = p.anc[off + i];
            p.out.x[off + i] = p.in.x[off + j];
            p.out.y[off + i] = p.in.y[off + j];
            p.out.th[off + i] = p.in.th[off + j];
            p.out.lw[off + i] = base_lw;
        } else {
            p.out.x[off + i] = p.in.x[off + i];
            p.out.y[off + i] = p.in.y[off + i];
            p.out.th[off + i] = p.in.th[off + i];
            p.out.lw[off + i] = p.in.lw[off + i];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        mcl_stats *st = p.stats + t;
        st->resampled = rs;
        if (rs) {  // weights reset to 1/NP (:222)
            st->log_norm = 0.0;
            st->ess = p.n_global;
            st->max_weight = 1.0 / p.n_global;
            st->max_logw = po->base_lw;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// k_shard_expand : sender side of the sharded resampling: resampled poses of every global slot this
// shard's particles fill, written at (slot - slot_first).
// -------------------------------------------------------------------------------------------------
struct ExpandParams {
    StateView in;
    const uint32_t *anc;  // indexed by slot - slot_first
    float *ox, *oy, *oth;
    long long n_slots;
};

__global__ void __launch_bounds__(256) k_shard_expand(const ExpandParams p) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n_slots;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint32_t j = p.anc[i];
        p.ox[i] = p.in.x[j];
        p.oy[i] = p.in.y[j];
        p.oth[i] = p.in.th[j];
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
    const mcl_stats *st = p.stats + t;
    const double mx = st->mean[0], my = st->mean[1], mth = st->mean[2], ln = st->log_norm;
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
    mcl_stats z = {};
    z.log_norm = 0.0;
    z.ess = n_global;
    z.max_weight = 1.0 / n_global;
    z.max_idx = 0;
    z.sum_w = 1.0; z.sum_w2 = 1.0 / n_global; z.max_logw = -log(n_global);
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
__global__ void k_reset_scan(unsigned long long *desc, unsigned int *tickets, int tpt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < tpt) desc[i] = 0ull;
    if (i == 0) tickets[0] = 0u;
}

__global__ void k_export_weights(const float *lw, const mcl_stats *stats, double *w, unsigned long long *q, long long n,
                                 long long ld, int use_override, double log_norm_override) {
    const int t = blockIdx.y;
    const double ln = use_override ? log_norm_override : stats[t].log_norm;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float v = lw[t * ld + i];
        if (w) w[t * n + i] = exp(__dsub_rn(static_cast<double>(v), ln));
        if (q) q[t * n + i] = quantize_weight(v, ln);
    }
}

__global__ void k_philox_selftest(uint4 ctr, uint2 key, uint4 *out, float2 *bm) {
    const uint4 r = philox4x32_10(ctr, key);
    out[0] = r;
    bm[0] = box_muller(r.x, r.y);
}

}  // namespace mcl
