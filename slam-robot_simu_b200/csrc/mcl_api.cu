// libmcl_b200.so -- host side of the C ABI declared in include/mcl.h.
//
// Everything here is plumbing: argument checks, the workspace layout, kernel launches on the
// caller's stream.  No arithmetic of the filter happens on the host; there is no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "mcl_kernels.cuh"

using namespace mcl;

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Layout {
    long long n, ld;
    int T, Lmax, tpt, bpt, cov_bpt;
    uint32_t per;
    size_t off_state[2], off_anc, off_desc, off_tickets, off_partials, off_cov, off_stats, off_prep, off_tab, off_lm,
        off_z, off_ofs, off_pose, off_total, off_trig, off_pub, off_ready, off_seg, off_reccount, off_prepflag, total;
};

int make_layout(const mcl_config &c, int sm_count, Layout &L) {
    if (c.n_particles < 1 || c.n_particles > 0x7FFFFF00ll) return MCL_ERR_ARG;  // 32-bit particle indices
    if (c.n_trials < 1 || c.n_trials > 65535) return MCL_ERR_ARG;
    if (c.max_landmarks < 1 || c.max_landmarks > 4096) return MCL_ERR_ARG;
    L.n = c.n_particles;
    L.ld = static_cast<long long>(align_up(static_cast<size_t>(c.n_particles), 128));
    L.T = c.n_trials;
    L.Lmax = c.max_landmarks;
    L.tpt = static_cast<int>((L.n + kScanTile - 1) / kScanTile);
    // k_step is persistent: the resident CTAs of every SM share a trial's particles in contiguous ranges that are
    // multiples of 128 particles; with many trials every trial still gets at least one CTA per 8192 particles
    const long long units = (L.n + kStepUnit - 1) / kStepUnit;
    long long blocks = static_cast<long long>(sm_count) * kStepCtasPerSm / L.T;
    const long long min_blocks = (L.T > 1) ? (L.n + 8191) / 8192 : 1;
    if (blocks < min_blocks) blocks = min_blocks;
    if (blocks > units) blocks = units;
    if (blocks < 1) blocks = 1;
    L.per = static_cast<uint32_t>(((units + blocks - 1) / blocks) * kStepUnit);
    L.bpt = static_cast<int>((L.n + L.per - 1) / L.per);
    L.cov_bpt = L.bpt;
    size_t o = 0;
    const size_t state_bytes = static_cast<size_t>(4) * L.T * L.ld * sizeof(float);
    for (int b = 0; b < 2; ++b) {
        L.off_state[b] = o;
        o = align_up(o + state_bytes, kAlign);
    }
    L.off_anc = o;      o = align_up(o + static_cast<size_t>(L.T) * L.ld * sizeof(uint32_t), kAlign);
    L.off_desc = o;     o = align_up(o + static_cast<size_t>(2) * L.T * L.tpt * sizeof(unsigned long long), kAlign);
    L.off_tickets = o;  o = align_up(o + static_cast<size_t>(L.T) * 4 * sizeof(unsigned int), kAlign);
    L.off_partials = o; o = align_up(o + static_cast<size_t>(L.T) * L.bpt * kPart * sizeof(double), kAlign);
    L.off_cov = o;      o = align_up(o + static_cast<size_t>(L.T) * L.cov_bpt * 8 * sizeof(double), kAlign);
    L.off_stats = o;    o = align_up(o + static_cast<size_t>(L.T) * sizeof(mcl_stats), kAlign);
    L.off_prep = o;     o = align_up(o + static_cast<size_t>(L.T) * sizeof(PrepOut), kAlign);
    L.off_tab = o;      o = align_up(o + static_cast<size_t>(L.T) * L.Lmax * sizeof(double4), kAlign);
    L.off_lm = o;       o = align_up(o + static_cast<size_t>(L.Lmax) * 2 * sizeof(double), kAlign);
    L.off_z = o;        o = align_up(o + static_cast<size_t>(L.T) * L.Lmax * 2 * sizeof(double), kAlign);
    L.off_ofs = o;      o = align_up(o + static_cast<size_t>(L.T) * sizeof(double), kAlign);
    L.off_pose = o;     o = align_up(o + static_cast<size_t>(L.T) * 3 * sizeof(double), kAlign);
    L.off_total = o;    o = align_up(o + static_cast<size_t>(L.T) * sizeof(unsigned long long), kAlign);
    L.off_pub = o;      o = align_up(o + static_cast<size_t>(L.T) * 2 * sizeof(uint4), kAlign);
    L.off_ready = o;    o = align_up(o + static_cast<size_t>(2) * L.T * (kDirectTiles / 32) * sizeof(unsigned int), kAlign);
    L.off_trig = o;     o = align_up(o + static_cast<size_t>(kTrigEntries) * sizeof(TrigEntry), kAlign);
    L.off_seg = o;      o = align_up(o + static_cast<size_t>(kMaxSegments) * sizeof(unsigned long long), kAlign);
    L.off_reccount = o; o = align_up(o + static_cast<size_t>(L.T) * sizeof(unsigned int), kAlign);
    L.off_prepflag = o; o = align_up(o + static_cast<size_t>(L.T) * sizeof(unsigned int), kAlign);
    L.total = o;
    return MCL_OK;
}

int validate(const mcl_config *c) {
    if (c == nullptr || c->struct_size != sizeof(mcl_config)) return MCL_ERR_ARG;
    if (!(c->dt > 0.0) || !(c->sigma_obs[0] > 0.0) || !(c->sigma_obs[1] > 0.0)) return MCL_ERR_ARG;
    if (c->likelihood < MCL_LIK_PAIR || c->likelihood > MCL_LIK_RB_F32) return MCL_ERR_ARG;
    if (c->likelihood == MCL_LIK_RB_FAN &&
        !(c->fan_range > 0.0 && c->fan_half_angle > 0.0 && c->fan_half_angle < kPi / 2 && c->fan_miss_chi2 >= 0.0))
        return MCL_ERR_ARG;
    if (c->motion < MCL_MOTION_A || c->motion > MCL_MOTION_B) return MCL_ERR_ARG;
    if (c->n_global < c->n_particles || c->slot_base < 0 || c->slot_base + c->n_particles > c->n_global)
        return MCL_ERR_ARG;
    if (c->n_global != c->n_particles && c->n_trials != 1) return MCL_ERR_ARG;
    if (c->n_global > 0x7FFFFF00ll) return MCL_ERR_ARG;  // slots and ancestors are 32-bit
    return MCL_OK;
}

}  // namespace

struct mcl_filter {
    mcl_config cfg;
    Layout lay;
    int sm_count = 0;
    bool owns_ws = false;
    unsigned char *ws = nullptr;
    StateView state[2];
    uint32_t *anc = nullptr;
    unsigned long long *desc[2] = {nullptr, nullptr};
    int desc_cur = 0;  // descriptor buffer the next scan launch uses (all zero)
    unsigned int *tickets = nullptr;
    double *partials = nullptr, *cov_partials = nullptr;
    mcl_stats *stats = nullptr;
    PrepOut *prep = nullptr;
    double4 *tab = nullptr;
    double *lm = nullptr, *z = nullptr, *ofs = nullptr, *pose = nullptr;
    unsigned long long *total = nullptr;
    TrigEntry *trig = nullptr;
    unsigned long long *seg = nullptr;  // [kMaxSegments] weight totals of the tile segments of mcl_shard_total
    int tiles_per_seg = 1;
    uint4 *wm_pub = nullptr;
    unsigned int *ready[2] = {nullptr, nullptr};  // double-buffered like desc[]
    unsigned int scan_epoch = 0;
    int cur = 0;
    int L = 0;
    uint32_t step = 0;
    bool have_particles = false;
    bool stats_from_particle0 = false;  // next reduce-only pass takes its reference pose from particle 0
    // pinned landing buffers of mcl_get_stats_begin/_end: a ring of two, so the read-back of step k can still be
    // in flight while step k + 1 is enqueued (the host then never idles the GPU waiting for statistics)
    mcl_stats *stats_pinned[2] = {nullptr, nullptr};  // mapped: [T] blocks, then the flag word of the direct landing
    cudaEvent_t stats_event[2] = {nullptr, nullptr};
    unsigned int *rec_count = nullptr;  // [T] producer CTAs finished so far (partial_write)
    unsigned int rec_target = 0u;       // ... and the value once every launch enqueued so far has finished
    unsigned int *prep_flag = nullptr;  // [T] epoch of the scalars (released by the scalars CTA of k_scan)
    unsigned int prep_epoch = 0u;
    bool merge_follows = false;         // set around the step of mcl_step_async
    bool partials_counted = false;      // the launch that wrote `partials` counted its CTAs (a scan may wait on the count)
    unsigned int stats_expect[2] = {0u, 0u};  // != 0: k_merge lands the block itself and sets the flag to this number
    unsigned int stats_seq = 0u;
    uint64_t stats_begun = 0, stats_ended = 0;
    bool stats_dirty = false;           // the CTA partials of the last k_step / k_reduce are newer than stats[]
    bool tile_prefix_valid = false;     // desc[desc_cur] holds the tile prefixes of the CURRENT weights (mcl_shard_total)
    bool desc_stale = false;            // desc[desc_cur] holds tile sums, not the zeros a look-back scan starts from
    double delta = 0.0;                 // fl(1 / n_global), NP_RECIP (particle_filter.py:32)
    std::string err;
    uint64_t launches = 0;
    bool profiling = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // 0..3 mcl_step*, 0..4 + 5 the peer step
    unsigned long long *dbg = nullptr;  // optional k_step phase timeline (mcl_debug_timeline)
    // pinned staging ring of the per-step host inputs (z, ofs): cudaMemcpyAsync from pageable memory may wait for
    // the stream, which would put a host synchronisation into every step
    static constexpr int kStageSlots = 4;
    unsigned char *stage_pin[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool stage_used[kStageSlots] = {false, false, false, false};
    size_t stage_bytes = 0;
    int stage_next = 0;
    // device-driven sharded filter (mcl_shard_peer_*): the peer-visible region (mailbox + receive staging), the
    // peers' mapped regions, the plan the exchange kernels keep on the device
    unsigned char *peer_region = nullptr;
    size_t peer_region_bytes = 0;
    int peer_rank = -1, peer_G = 0;
    bool peer_connected = false;
    void *peer_ipc_opened[kMaxShards] = {};
    PeerView *peer_view = nullptr;  // device copy
    ShardPlan *plan = nullptr;      // device
    uint32_t xchg_epoch = 0;        // exchange rounds so far (never reset: the flags in the boxes only grow)
};

namespace {

thread_local std::string g_err;  // errors raised before a handle exists

int fail(mcl_filter *h, int code, const char *what) {
    if (h) h->err = what;
    else g_err = what;
    return code;
}

#define MCL_CUDA(h, call)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            char buf__[512];                                                                      \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                     __FILE__, __LINE__);                                                         \
            return fail(h, MCL_ERR_CUDA, buf__);                                                  \
        }                                                                                         \
    } while (0)

#define MCL_LAUNCH_CHECK(h, name)                                                          \
    do {                                                                                   \
        (h)->launches += 1;                                                                \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            char buf__[512];                                                               \
            snprintf(buf__, sizeof buf__, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
            return fail(h, MCL_ERR_CUDA, buf__);                                           \
        }                                                                                  \
    } while (0)

inline cudaStream_t S(void *stream) { return static_cast<cudaStream_t>(stream); }

int use_device(mcl_filter *h) {
    int d = -1;
    MCL_CUDA(h, cudaGetDevice(&d));
    if (d != h->cfg.device) MCL_CUDA(h, cudaSetDevice(h->cfg.device));
    return MCL_OK;
}

// host twin of mcl::count_le (same formula; no FMA contraction on the host build, and the comb's two
// roundings are forced through volatiles)
long long host_count_le(unsigned long long K, double delta, double ofs, long long n) {
    const double nd = static_cast<double>(n);
    const double c = static_cast<double>(K) * kInvTwo52;
    const double g = std::floor((c - ofs) * nd);
    auto comb = [&](double k) {
        volatile double m = k * delta;
        volatile double u = m + ofs;
        return static_cast<double>(u);
    };
    double r = g;
    if (comb(g + 1.0) <= c) r = g + 1.0;
    else if (comb(g) > c) r = g - 1.0;
    return static_cast<long long>(std::fmin(std::fmax(r + 1.0, 0.0), nd));
}

// ---- launch helpers ------------------------------------------------------------------------------
// The two kernels of a step are launched with programmatic stream serialization (PDL): the dependent
// grid is scheduled while the previous one drains, and waits in cudaGridDependencySynchronize().
// MCL_PDL=0 in the environment turns it off.
bool pdl_enabled() {
    static const bool on = [] {
        const char *e = std::getenv("MCL_PDL");
        return !(e != nullptr && e[0] == '0');
    }();
    return on;
}
// Programmatic launch hides ~3 us of launch latency per kernel: worth 10 % of a 32 us step at C1, nothing at 125 M
// particles, and at C4 (16.8 M x 1024 landmarks, k_step 33 ms) it was measured to COST 14 ms per step: the dependent
// k_step CTAs become resident early and sit on the registers the scan's remaining tiles need.  Small steps only.
bool pdl_for(const mcl_filter *h);

bool pdl_for(const mcl_filter *h) {
    return pdl_enabled() && !h->profiling && static_cast<long long>(h->lay.T) * h->lay.n <= (1ll << 23);
}

template <typename P>
void launch_pdl(void (*kernel)(P), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, const P &params) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, params);
}

void fill_prep(mcl_filter *h, PrepParams &p, const double *z_dev, const double *ofs_dev, uint32_t flags) {
    const mcl_config &c = h->cfg;
    p = PrepParams{};
    p.lm = h->lm;
    p.z = z_dev;
    p.ofs = ofs_dev;
    p.prep = h->prep;
    p.tab = h->tab;
    p.L = h->L;
    p.Lmax = h->lay.Lmax;
    p.flags = flags;
    p.likelihood = c.likelihood;
    p.ess_threshold = c.ess_threshold;
    p.vdt = c.vel * c.dt;
    p.wdt = c.yaw_rate * c.dt;
    p.inv_s0sq = 1.0 / (c.sigma_obs[0] * c.sigma_obs[0]);
    p.inv_s1sq = 1.0 / (c.sigma_obs[1] * c.sigma_obs[1]);
    p.lik_const = (c.likelihood == MCL_LIK_RB_FAN) ? 0.0 : -static_cast<double>(h->L) * std::log(2.0 * kPi * c.sigma_obs[0] * c.sigma_obs[1]);
    p.fan_range = c.fan_range;
    p.fan_tan = std::tan(kPi / 2 - c.fan_half_angle);  // ry >= |rx| tan(pi/2 - angle), graph_based_slam.py:156-158
    p.fan_miss = c.fan_miss_chi2;
    p.r_dist = c.sigma_obs[0];
    p.dir_sigma = c.sigma_obs[1];
    p.inv_n = h->delta;
    p.n_global = static_cast<double>(c.n_global);
    p.underflow_rule = c.underflow_rule;
    p.seed = c.seed;
    p.step = h->step;
}

void fill_scan(mcl_filter *h, ScanParams &p) {
    p = ScanParams{};
    p.prep_flag = h->prep_flag;
    p.prep_epoch = h->prep_epoch;
    p.lw = h->state[h->cur].lw;
    p.stats = h->stats;
    p.desc = h->desc[h->desc_cur];
    p.desc_next = nullptr;
    p.ready = h->ready[h->desc_cur];
    p.ready_next = nullptr;
    p.tickets = h->tickets;
    p.anc = h->anc;
    p.total = nullptr;
    p.n = h->lay.n;
    p.ld = h->lay.ld;
    p.n_global = h->cfg.n_global;
    p.j_global_base = h->cfg.slot_base;
    p.slot_first = 0;
    p.anc_cap = h->lay.ld;
    p.own_lo = 0;
    p.own_hi = h->cfg.n_global;
    p.k_offset = 0;
    p.delta = h->delta;
    p.tpt = h->lay.tpt;
    p.wm_pub = h->wm_pub;
    p.epoch = ++h->scan_epoch;
    if (p.epoch == 0u) p.epoch = ++h->scan_epoch;  // 0 is the initial value of the flags
}

// mcl_stats on the device is merged lazily from the CTA partials: by the statistics CTA of the next k_scan,
// or here when the host (or a stage call) needs it first.
int ensure_stats(mcl_filter *h, cudaStream_t st, uint32_t *host_out = nullptr, unsigned int *host_flag = nullptr,
                 unsigned int host_seq = 0u) {
    if (!h->stats_dirty) return MCL_OK;
    MergeParams m{};
    m.host_out = host_out;
    m.host_flag = host_flag;
    m.host_seq = host_seq;
    m.partials = h->partials;
    m.bpt = h->lay.bpt;
    m.stats = h->stats;
    m.cur = h->state[h->cur];
    m.ld = h->lay.ld;
    m.n_global = static_cast<double>(h->cfg.n_global);
    m.apply_rule = (h->cfg.n_global == h->cfg.n_particles) ? 1 : 0;  // a shard leaves the rule to the merge of all shards
    launch_pdl(k_merge, dim3(h->lay.T), dim3(kStepThreads), 0, st, pdl_for(h), m);
    MCL_LAUNCH_CHECK(h, "k_merge");
    h->stats_dirty = false;
    return MCL_OK;
}

// In-step (or stage) scan: per-step scalars by an extra CTA of each trial, then the scan of the trials that
// resample.  `full` = false launches the scalars only (MCL_NO_RESAMPLE).
int launch_step_scan(mcl_filter *h, const double *z_dev, const double *ofs_dev, uint32_t flags, bool full,
                     cudaStream_t st) {
    ScanParams sp;
    fill_scan(h, sp);
    fill_prep(h, sp.prep, z_dev, ofs_dev, flags);
    sp.do_prep = 1;
    h->prep_epoch += 1u;
    sp.prep_epoch = h->prep_epoch;
    sp.trig = h->trig;
    sp.dbg = (h->dbg != nullptr && full) ? h->dbg + static_cast<size_t>(h->lay.T) * h->lay.bpt * kDbgSlots : nullptr;
    if (full) {
        if (h->desc_stale) {  // tile sums of an mcl_shard_total that no expansion consumed
            k_reset_scan<<<(h->lay.tpt + 255) / 256, 256, 0, st>>>(sp.desc, sp.ready, h->tickets, h->lay.tpt);
            MCL_LAUNCH_CHECK(h, "k_reset_scan");
            h->desc_stale = false;
            h->tile_prefix_valid = false;
        }
        sp.desc_next = h->desc[h->desc_cur ^ 1];
        sp.ready_next = h->ready[h->desc_cur ^ 1];
        h->desc_cur ^= 1;
    }
    if (h->stats_dirty) {  // statistics still live in the CTA partials of the previous k_step / k_reduce
        sp.partials = h->partials;
        sp.n_partials = h->lay.bpt;
        sp.cur = h->state[h->cur];
    }
    // persistent tile CTAs (at most 4 per SM) + [the CTA that merges the partials into mcl_stats] + the CTA
    // that derives the per-step scalars
    // (one filter: one CTA per tile, the hardware scheduler keeps the look-back pipeline full; many trials:
    // a trial's few tiles share persistent CTAs so the per-CTA prologue is paid once)
    static const int persist_single = [] { const char *e = std::getenv("MCL_SCAN_PERSISTENT"); return e ? std::atoi(e) : 0; }();
    // geometry: all tiles in one wave of 4 per SM (a filter of up to ~1.2 M particles) -> 256 threads x 8 sources; beyond that the
    // scan is bound by per-tile latency at a fixed number of resident tiles -> 128 threads x 16, eight tiles per SM
    const bool wide = static_cast<long long>(h->lay.T) * h->lay.tpt > static_cast<long long>(h->sm_count) * scan_ctas_per_sm(8);
    const int per_sm = wide ? scan_ctas_per_sm(16) : scan_ctas_per_sm(8);
    const int tile_ctas = !full ? 0 : ((h->lay.T == 1 && persist_single <= 0) ? h->lay.tpt
                                       : std::min(h->lay.tpt, std::max(1, h->sm_count * std::max(persist_single, per_sm) / h->lay.T)));
    sp.persistent = (tile_ctas < h->lay.tpt) ? 1 : 0;
    sp.rec_count = h->rec_count;
    sp.rec_target = h->rec_target;
    sp.early_merge = (h->lay.T == 1 && !wide && !sp.persistent && sp.n_partials > 0 && h->partials_counted && pdl_for(h)) ? 1 : 0;
    sp.extras_front = sp.early_merge;
    dim3 grid(tile_ctas + (h->stats_dirty ? 2 : 1), h->lay.T);
    h->stats_dirty = false;
    if (wide) launch_pdl(k_scan<kScanAncestors, 16>, grid, dim3(kScanTile / 16), 0, st, pdl_for(h), sp);
    else launch_pdl(k_scan<kScanAncestors, 8>, grid, dim3(kScanTile / 8), 0, st, pdl_for(h), sp);
    MCL_LAUNCH_CHECK(h, "k_scan");
    return MCL_OK;
}

// Sharded resampling: the scan runs outside a step, with the weight map of the (global) statistics.
int launch_shard_scan(mcl_filter *h, ScanParams &sp, int mode, cudaStream_t st) {
    if (mode == kScanExpand && h->tile_prefix_valid) {
        sp.tile_prefix = sp.desc;  // written by mcl_shard_total for exactly these weights: no look-back
        sp.seg_prefix = h->seg;
        sp.tiles_per_seg = h->tiles_per_seg;
    } else {
        k_reset_scan<<<(h->lay.tpt + 255) / 256, 256, 0, st>>>(sp.desc, sp.ready, h->tickets, h->lay.tpt);
        MCL_LAUNCH_CHECK(h, "k_reset_scan");
    }
    h->tile_prefix_valid = false;
    h->desc_stale = false;  // the trailing k_reset_scan below leaves zeros
    dim3 grid(h->lay.tpt, 1);
    static const bool persistent_expand = [] { const char *e = std::getenv("MCL_EXPAND_PERSISTENT"); return e == nullptr || e[0] != '0'; }();
    if (sp.tile_prefix != nullptr && persistent_expand) {
        // tiles are independent once their prefixes are known: persistent CTAs (tickets) pay the kernel prologue once
        // per CTA instead of once per tile (expand 0.55 -> 0.50 ms at 125 M particles)
        sp.persistent = 1;
        grid.x = std::min(h->lay.tpt, h->sm_count * scan_ctas_per_sm(16));
    }
    k_scan<kScanExpand, 16><<<grid, kScanTile / 16, 0, st>>>(sp);
    MCL_LAUNCH_CHECK(h, "k_scan");
    k_reset_scan<<<(h->lay.tpt + 255) / 256, 256, 0, st>>>(sp.desc, sp.ready, h->tickets, h->lay.tpt);
    MCL_LAUNCH_CHECK(h, "k_reset_scan");
    return MCL_OK;
}

template <int LIK>
void launch_step_lik(const StepParams &p, dim3 grid, size_t smem, bool philox, cudaStream_t st, bool pdl) {
    if (philox) launch_pdl(k_step<LIK, true>, grid, dim3(kStepThreads), smem, st, pdl, p);
    else launch_pdl(k_step<LIK, false>, grid, dim3(kStepThreads), smem, st, pdl, p);
}

int step_variant(const mcl_config &c) {
    if (c.likelihood == MCL_LIK_PAIR) return kLikPair;
    if (c.likelihood == MCL_LIK_RB) return kLikRB;
    if (c.likelihood == MCL_LIK_RB_FAN) return kLikRBFan;
    if (c.likelihood == MCL_LIK_PAIR_F32) return kLikPair32;
    if (c.likelihood == MCL_LIK_RB_F32) return kLikRB32;
    return (c.sigma_obs[0] == c.sigma_obs[1]) ? kLikMomentIso : kLikMomentGen;
}

struct StepWindow {
    uint32_t lo, hi;
    const float *x, *y, *th;
};

int launch_step(mcl_filter *h, const float *noise_dev, uint32_t flags, bool reduce_only, cudaStream_t st,
                const StepWindow *win = nullptr, const ShardPlan *plan = nullptr) {
    const mcl_config &c = h->cfg;
    h->tile_prefix_valid = false;  // the weights (or their statistics) are about to change
    StepParams p{};
    p.in = h->state[h->cur];
    p.out = h->state[h->cur ^ 1];
    p.anc = h->anc;
    p.win_lo = win ? win->lo : 0u;
    p.win_hi = win ? win->hi : static_cast<uint32_t>(h->lay.n);
    p.recv_x = win ? win->x : nullptr;
    p.recv_y = win ? win->y : nullptr;
    p.recv_th = win ? win->th : nullptr;
    p.plan = plan;
    p.noise = noise_dev;
    p.prep = h->prep;
    p.tab = h->tab;
    p.trig = h->trig;
    p.partials = h->partials;
    // the CTAs count themselves only when a scan may wait on the count: not when k_merge is enqueued right behind this
    // launch (mcl_step_async), and not without programmatic launch (large filters, profiling) -- there the next kernel
    // starts after this grid has completed anyway
    h->partials_counted = !h->merge_follows && pdl_for(h);
    if (!h->partials_counted) {
        p.rec_count = nullptr;
    } else {
        p.rec_count = h->rec_count;
        h->rec_target += static_cast<unsigned int>(h->lay.bpt);  // (every counted k_step / k_reduce launch: bpt CTAs per trial)
    }
    p.tickets = h->tickets;
    p.stats = h->stats;
    p.dbg = h->dbg;
    p.n = static_cast<uint32_t>(h->lay.n);
    p.per = h->lay.per;
    p.ld = h->lay.ld;
    p.slot_base = c.slot_base;
    p.T = h->lay.T;
    p.bpt = h->lay.bpt;
    p.Lmax = h->lay.Lmax;
    p.motion = c.motion;
    p.step = h->step;
    p.seed = c.seed;
    p.rk = philox_round_keys(c.seed);
    p.vdt_f = static_cast<float>(c.vel * c.dt);
    p.wdt_f = static_cast<float>(c.yaw_rate * c.dt);
    p.wdt = c.yaw_rate * c.dt;
    p.vel = c.vel;
    p.yaw_rate = c.yaw_rate;
    p.dt = c.dt;
    for (int k = 0; k < 3; ++k) p.q_std[k] = static_cast<float>(c.q_std[k]);
    p.inv_s0sq = 1.0 / (c.sigma_obs[0] * c.sigma_obs[0]);
    p.inv_s1sq = 1.0 / (c.sigma_obs[1] * c.sigma_obs[1]);
    p.inv_s0_f = static_cast<float>(1.0 / c.sigma_obs[0]);
    p.inv_s1_f = static_cast<float>(1.0 / c.sigma_obs[1]);
    p.fan_r2 = static_cast<float>(c.fan_range * c.fan_range);
    p.fan_tan = static_cast<float>(std::tan(kPi / 2 - c.fan_half_angle));
    p.fan_inv_r = static_cast<float>(1.0 / c.sigma_obs[0]);
    p.fan_inv_sphi = static_cast<float>(1.0 / c.sigma_obs[1]);
    p.fan_miss = static_cast<float>(c.fan_miss_chi2);
    {
        // motion_model.py:43-48: sigma = a_i v^2 + a_j w^2 and np.random.normal(0, sigma ** 2)
        const double v2 = c.vel * c.vel, w2 = c.yaw_rate * c.yaw_rate;
        const double *a = c.motion_b_alpha;
        const double s0 = a[0] * v2 + a[1] * w2, s1 = a[2] * v2 + a[3] * w2, s2 = a[4] * v2 + a[5] * w2;
        p.mb_sig[0] = s0 * s0;
        p.mb_sig[1] = s1 * s1;
        p.mb_sig[2] = s2 * s2;
    }
    dim3 grid(h->lay.bpt, h->lay.T);
    if (reduce_only) {
        k_reduce<<<grid, kStepThreads, 0, st>>>(p, h->stats_from_particle0 ? 1 : 0);
        MCL_LAUNCH_CHECK(h, "k_reduce");
        h->stats_dirty = true;
        return MCL_OK;
    }
    const bool philox = (noise_dev == nullptr);
    const bool pdl = pdl_for(h);
    p.prep_flag = h->prep_flag;
    p.prep_epoch = h->prep_epoch;
    p.early_prep = (pdl && win == nullptr && plan == nullptr) ? 1 : 0;
    // shared-memory landmark table: 32 bytes per landmark; the fp32 range/bearing kernel stages every value twice (64)
    const size_t smem = static_cast<size_t>(h->L) * sizeof(double4) * (c.likelihood == MCL_LIK_RB_F32 ? 2 : 1);
    switch (step_variant(c)) {
        case kLikPair: launch_step_lik<kLikPair>(p, grid, smem, philox, st, pdl); break;
        case kLikRB: launch_step_lik<kLikRB>(p, grid, smem, philox, st, pdl); break;
        case kLikRBFan: launch_step_lik<kLikRBFan>(p, grid, smem, philox, st, pdl); break;
        case kLikPair32: launch_step_lik<kLikPair32>(p, grid, smem, philox, st, pdl); break;
        case kLikRB32: launch_step_lik<kLikRB32>(p, grid, smem, philox, st, pdl); break;
        case kLikMomentIso: launch_step_lik<kLikMomentIso>(p, grid, 0, philox, st, pdl); break;
        default: launch_step_lik<kLikMomentGen>(p, grid, 0, philox, st, pdl); break;
    }
    MCL_LAUNCH_CHECK(h, "k_step");
    h->stats_dirty = true;
    return MCL_OK;
}

int reduce_only_pass(mcl_filter *h, cudaStream_t st) {
    int rc = h->stats_from_particle0 ? MCL_OK : ensure_stats(h, st);
    if (rc != MCL_OK) return rc;
    rc = launch_step(h, nullptr, MCL_NO_RESAMPLE, true, st);
    h->stats_from_particle0 = false;
    return rc;
}

int copy2d(mcl_filter *h, void *dst, size_t dpitch_el, const void *src, size_t spitch_el, size_t el_bytes, size_t width_el,
           size_t rows, cudaMemcpyKind kind, cudaStream_t st) {
    MCL_CUDA(h, cudaMemcpy2DAsync(dst, dpitch_el * el_bytes, src, spitch_el * el_bytes, width_el * el_bytes, rows, kind, st));
    return MCL_OK;
}

int step_impl(mcl_filter *h, const double *z_dev, const double *ofs_dev, const float *noise_dev, uint32_t flags,
              cudaStream_t st) {
    if (h->L <= 0) return fail(h, MCL_ERR_STATE, "mcl_set_landmarks has not been called");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "particles are not initialised");
    if (noise_dev == nullptr && !(flags & MCL_PHILOX_NOISE))
        return fail(h, MCL_ERR_ARG, "noise_dev is NULL and MCL_PHILOX_NOISE is not set");
    if (h->cfg.n_global != h->cfg.n_particles && !(flags & MCL_NO_RESAMPLE))
        return fail(h, MCL_ERR_STATE, "a particle-sharded filter resamples through mcl_shard_total/mcl_shard_expand");
    const bool prof = h->profiling;
    if (prof) MCL_CUDA(h, cudaEventRecord(h->ev[0], st));
    if (prof) MCL_CUDA(h, cudaEventRecord(h->ev[1], st));
    int rc = launch_step_scan(h, z_dev, ofs_dev, flags, !(flags & MCL_NO_RESAMPLE), st);
    if (rc != MCL_OK) return rc;
    if (prof) MCL_CUDA(h, cudaEventRecord(h->ev[2], st));
    rc = launch_step(h, noise_dev, flags, false, st);
    if (rc != MCL_OK) return rc;
    if (prof) MCL_CUDA(h, cudaEventRecord(h->ev[3], st));
    h->cur ^= 1;
    h->step += 1;
    return MCL_OK;
}

// z_direct != nullptr: a small observation is not copied to the device at all -- *z_direct receives the device address
// of the pinned, mapped staging slot and the scalars CTA of k_scan reads the few hundred bytes over PCIe (one round trip
// of a CTA that has slack) instead of waiting for a copy-engine operation in front of the step.  The caller records
// stage_ev[*slot_out] behind the kernels that read the slot.
int stage_inputs(mcl_filter *h, const double *z_host, const double *ofs_host, cudaStream_t st, const double **z_direct = nullptr,
                 int *slot_out = nullptr) {
    if (z_direct) *z_direct = nullptr;
    if (z_host == nullptr && ofs_host == nullptr) return MCL_OK;
    if (ofs_host != nullptr) {
        // rand() * NP_RECIP (particle_filter.py:214) lies in [0, 1/N); the slot arithmetic relies on it
        for (int t = 0; t < h->lay.T; ++t)
            if (!(ofs_host[t] >= 0.0 && ofs_host[t] < h->delta))
                return fail(h, MCL_ERR_ARG, "comb offset outside [0, 1/N): pass rand() * (1 / n_global)");
    }
    const size_t zb = static_cast<size_t>(h->lay.T) * h->L * 2 * sizeof(double);
    const size_t ob = static_cast<size_t>(h->lay.T) * sizeof(double);
    const size_t need = static_cast<size_t>(h->lay.T) * h->lay.Lmax * 2 * sizeof(double) + ob;
    if (h->stage_bytes < need) {  // first use
        for (int i = 0; i < mcl_filter::kStageSlots; ++i) {
            if (h->stage_pin[i]) cudaFreeHost(h->stage_pin[i]);
            h->stage_pin[i] = nullptr;
            h->stage_used[i] = false;
            MCL_CUDA(h, cudaHostAlloc(reinterpret_cast<void **>(&h->stage_pin[i]), need, cudaHostAllocMapped));
            if (h->stage_ev[i] == nullptr) MCL_CUDA(h, cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming));
        }
        h->stage_bytes = need;
    }
    const int slot = h->stage_next;
    h->stage_next = (slot + 1) % mcl_filter::kStageSlots;
    if (h->stage_used[slot]) MCL_CUDA(h, cudaEventSynchronize(h->stage_ev[slot]));  // the copy that last read this slot
    unsigned char *pin = h->stage_pin[slot];
    bool direct = false;
    if (z_host != nullptr) {
        std::memcpy(pin, z_host, zb);
        if (z_direct != nullptr && slot_out != nullptr && zb <= 4096) {
            void *dev = nullptr;
            MCL_CUDA(h, cudaHostGetDevicePointer(&dev, pin, 0));
            *z_direct = static_cast<const double *>(dev);
            *slot_out = slot;
            direct = true;
        } else {
            MCL_CUDA(h, cudaMemcpyAsync(h->z, pin, zb, cudaMemcpyHostToDevice, st));
        }
    }
    if (ofs_host != nullptr) {
        std::memcpy(pin + need - ob, ofs_host, ob);
        MCL_CUDA(h, cudaMemcpyAsync(h->ofs, pin + need - ob, ob, cudaMemcpyHostToDevice, st));
    }
    if (!direct) MCL_CUDA(h, cudaEventRecord(h->stage_ev[slot], st));  // (direct: recorded by the caller, behind the step)
    h->stage_used[slot] = true;
    return MCL_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int mcl_abi_version(void) { return MCL_ABI_VERSION; }

const char *mcl_status_string(int status) {
    switch (status) {
        case MCL_OK: return "ok";
        case MCL_ERR_ARG: return "bad argument";
        case MCL_ERR_CUDA: return "CUDA error";
        case MCL_ERR_STATE: return "call sequence error";
        case MCL_ERR_NOMEM: return "out of memory / workspace too small";
        case MCL_ERR_NODEVICE: return "no sm_100 CUDA device";
        default: return "unknown status";
    }
}

const char *mcl_last_error(const mcl_filter *h) { return h ? h->err.c_str() : g_err.c_str(); }

int mcl_config_default(mcl_config *cfg, double period_ms) {
    if (cfg == nullptr) return MCL_ERR_ARG;
    std::memset(cfg, 0, sizeof *cfg);
    const double deg = kPi / 180.0;
    cfg->struct_size = sizeof(mcl_config);
    cfg->device = 0;
    cfg->n_particles = 1000;             // particle_filter.py:31
    cfg->n_trials = 1;
    cfg->max_landmarks = 64;
    cfg->dt = period_ms / 1000.0;        // :30
    cfg->yaw_rate = 10.0 * deg;          // :47
    cfg->vel = 10.0 * cfg->yaw_rate;     // :46,:48
    cfg->q_std[0] = 0.03;                // :62-65
    cfg->q_std[1] = 0.03;
    cfg->q_std[2] = 2.0 * deg;
    cfg->sigma_obs[0] = 0.3;             // :68-70
    cfg->sigma_obs[1] = 0.3;
    cfg->ess_threshold = 1000.0 / 100.0; // :33
    for (int k = 0; k < 6; ++k) cfg->motion_b_alpha[k] = 0.1;  // graph_based_slam.py:605
    cfg->likelihood = MCL_LIK_PAIR;
    cfg->motion = MCL_MOTION_A;
    cfg->seed = 42;
    cfg->n_global = cfg->n_particles;
    cfg->slot_base = 0;
    cfg->fan_range = 15.0;              // graph_based_slam.py:901
    cfg->fan_half_angle = 80.0 * deg;   // :902
    cfg->fan_miss_chi2 = 18.0;          // two 3-sigma residuals
    cfg->underflow_rule = 1;            // particle_filter.py:234-237
    return MCL_OK;
}

int mcl_workspace_bytes(const mcl_config *cfg, size_t *bytes) {
    if (bytes == nullptr) return MCL_ERR_ARG;
    int rc = validate(cfg);
    if (rc != MCL_OK) return fail(nullptr, rc, "invalid mcl_config");
    Layout L;
    // the layout depends on the SM count only through the number of k_step CTAs (partials: 128 B each); size for
    // the device the filter will run on when one is visible, else for the B200's 148 SMs
    int sms = 148;
    {
        int count = 0;
        cudaDeviceProp prop;
        if (cudaGetDeviceCount(&count) == cudaSuccess && cfg->device >= 0 && cfg->device < count &&
            cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess)
            sms = prop.multiProcessorCount;
        else
            cudaGetLastError();
    }
    rc = make_layout(*cfg, sms, L);
    if (rc != MCL_OK) return fail(nullptr, rc, "invalid sizes in mcl_config");
    *bytes = L.total;
    return MCL_OK;
}

int mcl_create(const mcl_config *cfg, void *workspace_dev, size_t workspace_bytes, mcl_filter **out) {
    if (out == nullptr) return MCL_ERR_ARG;
    *out = nullptr;
    int rc = validate(cfg);
    if (rc != MCL_OK) return fail(nullptr, rc, "invalid mcl_config");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(nullptr, MCL_ERR_NODEVICE, "no CUDA device is visible (this engine has no CPU fallback)");
    }
    if (cfg->device < 0 || cfg->device >= count) return fail(nullptr, MCL_ERR_ARG, "device ordinal out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, MCL_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return fail(nullptr, MCL_ERR_NODEVICE, "device is not compute capability 10.x (kernels are built for sm_100a only)");
    mcl_filter *h = new (std::nothrow) mcl_filter();
    if (h == nullptr) return fail(nullptr, MCL_ERR_NOMEM, "host allocation failed");
    h->cfg = *cfg;
    h->sm_count = prop.multiProcessorCount;
    rc = make_layout(*cfg, h->sm_count, h->lay);
    if (rc != MCL_OK) {
        delete h;
        return fail(nullptr, rc, "invalid sizes in mcl_config");
    }
    if (cudaSetDevice(cfg->device) != cudaSuccess) {
        delete h;
        return fail(nullptr, MCL_ERR_CUDA, "cudaSetDevice failed");
    }
    if (workspace_dev != nullptr) {
        if (workspace_bytes < h->lay.total || (reinterpret_cast<uintptr_t>(workspace_dev) % kAlign) != 0) {
            delete h;
            return fail(nullptr, MCL_ERR_NOMEM, "workspace too small or not 256-byte aligned");
        }
        h->ws = static_cast<unsigned char *>(workspace_dev);
    } else {
        void *p = nullptr;
        if (cudaMalloc(&p, h->lay.total) != cudaSuccess) {
            cudaGetLastError();
            delete h;
            return fail(nullptr, MCL_ERR_NOMEM, "cudaMalloc of the workspace failed");
        }
        h->ws = static_cast<unsigned char *>(p);
        h->owns_ws = true;
    }
    const Layout &L = h->lay;
    const size_t plane = static_cast<size_t>(L.T) * L.ld;
    for (int b = 0; b < 2; ++b) {
        float *base = reinterpret_cast<float *>(h->ws + L.off_state[b]);
        h->state[b].x = base;
        h->state[b].y = base + plane;
        h->state[b].th = base + 2 * plane;
        h->state[b].lw = base + 3 * plane;
    }
    h->anc = reinterpret_cast<uint32_t *>(h->ws + L.off_anc);
    h->desc[0] = reinterpret_cast<unsigned long long *>(h->ws + L.off_desc);
    h->desc[1] = h->desc[0] + static_cast<size_t>(L.T) * L.tpt;
    h->tickets = reinterpret_cast<unsigned int *>(h->ws + L.off_tickets);
    h->partials = reinterpret_cast<double *>(h->ws + L.off_partials);
    h->cov_partials = reinterpret_cast<double *>(h->ws + L.off_cov);
    h->stats = reinterpret_cast<mcl_stats *>(h->ws + L.off_stats);
    h->prep = reinterpret_cast<PrepOut *>(h->ws + L.off_prep);
    h->tab = reinterpret_cast<double4 *>(h->ws + L.off_tab);
    h->lm = reinterpret_cast<double *>(h->ws + L.off_lm);
    h->z = reinterpret_cast<double *>(h->ws + L.off_z);
    h->ofs = reinterpret_cast<double *>(h->ws + L.off_ofs);
    h->pose = reinterpret_cast<double *>(h->ws + L.off_pose);
    h->total = reinterpret_cast<unsigned long long *>(h->ws + L.off_total);
    h->trig = reinterpret_cast<TrigEntry *>(h->ws + L.off_trig);
    h->seg = reinterpret_cast<unsigned long long *>(h->ws + L.off_seg);
    h->rec_count = reinterpret_cast<unsigned int *>(h->ws + L.off_reccount);
    h->prep_flag = reinterpret_cast<unsigned int *>(h->ws + L.off_prepflag);
    h->wm_pub = reinterpret_cast<uint4 *>(h->ws + L.off_pub);
    h->ready[0] = reinterpret_cast<unsigned int *>(h->ws + L.off_ready);
    h->ready[1] = h->ready[0] + static_cast<size_t>(L.T) * (kDirectTiles / 32);
    h->delta = 1.0 / static_cast<double>(cfg->n_global);
    // the pair kernels stage up to 4096 landmarks (128 KiB) in dynamic shared memory
    const int max_dyn = 4096 * static_cast<int>(sizeof(double4));
    cudaFuncSetAttribute(k_step<kLikPair, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    cudaFuncSetAttribute(k_step<kLikPair, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    cudaFuncSetAttribute(k_step<kLikRB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    cudaFuncSetAttribute(k_step<kLikRB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    cudaFuncSetAttribute(k_step<kLikRBFan, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    cudaFuncSetAttribute(k_step<kLikRBFan, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    cudaFuncSetAttribute(k_step<kLikPair32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    cudaFuncSetAttribute(k_step<kLikPair32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    cudaFuncSetAttribute(k_step<kLikRB32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    cudaFuncSetAttribute(k_step<kLikRB32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    // sin/cos table of sincos_tab: entry k-128 holds sin, cos of k*pi/32 and the correction -k*(pi/32 - HI)
    {
        static TrigEntry tab[kTrigEntries];
        for (int i = 0; i < kTrigEntries; ++i) {
            const double k = static_cast<double>(i - 128);
            tab[i].s = std::sin(k * (kPi / 32.0));
            tab[i].c = std::cos(k * (kPi / 32.0));
            tab[i].corr = -k * (kPi / 32.0 - static_cast<double>(kTrigStepHi));
            tab[i].pad = 0.0;
        }
        if (cudaMemcpy(h->trig, tab, sizeof tab, cudaMemcpyHostToDevice) != cudaSuccess) {
            if (h->owns_ws) cudaFree(h->ws);
            delete h;
            return fail(nullptr, MCL_ERR_CUDA, "uploading the trig table failed");
        }
    }
    cudaError_t e = cudaMemset(h->ws + L.off_tickets, 0, static_cast<size_t>(L.T) * 4 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(h->ws + L.off_desc, 0, static_cast<size_t>(2) * L.T * L.tpt * sizeof(unsigned long long));
    if (e == cudaSuccess)
        e = cudaMemset(h->ws + L.off_ready, 0, static_cast<size_t>(2) * L.T * (kDirectTiles / 32) * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(h->ws + L.off_pub, 0, static_cast<size_t>(L.T) * 2 * sizeof(uint4));
    if (e == cudaSuccess) e = cudaMemset(h->ws + L.off_reccount, 0, static_cast<size_t>(L.T) * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(h->ws + L.off_prepflag, 0, static_cast<size_t>(L.T) * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(h->ws + L.off_stats, 0, static_cast<size_t>(L.T) * sizeof(mcl_stats));
    if (e != cudaSuccess) {
        if (h->owns_ws) cudaFree(h->ws);
        delete h;
        return fail(nullptr, MCL_ERR_CUDA, "initialising the workspace failed");
    }
    *out = h;
    return MCL_OK;
}

int mcl_destroy(mcl_filter *h) {
    if (h == nullptr) return MCL_OK;
    cudaSetDevice(h->cfg.device);
    for (cudaEvent_t &e : h->ev)
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (h->stats_event[i]) cudaEventDestroy(h->stats_event[i]);
        if (h->stats_pinned[i]) cudaFreeHost(h->stats_pinned[i]);
    }
    if (h->dbg) cudaFree(h->dbg);
    for (int i = 0; i < mcl_filter::kStageSlots; ++i) {
        if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
        if (h->stage_pin[i]) cudaFreeHost(h->stage_pin[i]);
    }
    for (void *p : h->peer_ipc_opened)
        if (p) cudaIpcCloseMemHandle(p);
    if (h->peer_view) cudaFree(h->peer_view);
    if (h->plan) cudaFree(h->plan);
    if (h->peer_region) cudaFree(h->peer_region);
    if (h->owns_ws && h->ws) cudaFree(h->ws);
    delete h;
    return MCL_OK;
}

int64_t mcl_stride(const mcl_filter *h) { return h ? h->lay.ld : 0; }

int mcl_set_landmarks(mcl_filter *h, const double *lm_xy_host, int32_t n_landmarks, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (lm_xy_host == nullptr || n_landmarks < 1 || n_landmarks > h->lay.Lmax)
        return fail(h, MCL_ERR_ARG, "landmark table is NULL or larger than max_landmarks");
    if (h->cfg.likelihood == MCL_LIK_RB_F32 && n_landmarks > 2048)
        return fail(h, MCL_ERR_ARG, "MCL_LIK_RB_F32 stages 64 bytes per landmark in shared memory: at most 2048 landmarks");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    MCL_CUDA(h, cudaMemcpyAsync(h->lm, lm_xy_host, static_cast<size_t>(n_landmarks) * 2 * sizeof(double),
                                cudaMemcpyHostToDevice, S(stream)));
    MCL_CUDA(h, cudaStreamSynchronize(S(stream)));  // the host array may be a temporary
    h->L = n_landmarks;
    return MCL_OK;
}

int mcl_init_particles(mcl_filter *h, const double *pose_host, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (pose_host == nullptr) return fail(h, MCL_ERR_ARG, "pose_host is NULL");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    cudaStream_t st = S(stream);
    const Layout &L = h->lay;
    MCL_CUDA(h, cudaMemcpyAsync(h->pose, pose_host, static_cast<size_t>(L.T) * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    MCL_CUDA(h, cudaStreamSynchronize(st));
    const float lw0 = 0.0f;  // stored log-weights are relative: uniform = all equal
    const int bx = static_cast<int>(std::min<long long>((L.ld + 255) / 256, 2048));
    h->cur = 0;
    k_init<<<dim3(bx, L.T), 256, 0, st>>>(h->state[0], h->pose, L.n, L.ld, lw0);
    MCL_LAUNCH_CHECK(h, "k_init");
    k_init<<<dim3(bx, L.T), 256, 0, st>>>(h->state[1], h->pose, L.n, L.ld, lw0);
    MCL_LAUNCH_CHECK(h, "k_init");
    k_init_stats<<<(L.T + 127) / 128, 128, 0, st>>>(h->stats, h->pose, static_cast<double>(h->cfg.n_global), L.T);
    MCL_LAUNCH_CHECK(h, "k_init_stats");
    h->have_particles = true;
    h->tile_prefix_valid = false;
    h->stats_dirty = false;
    h->step = 0;
    return MCL_OK;
}

int mcl_set_particles(mcl_filter *h, const float *x, const float *y, const float *yaw, const float *logw, int on_device,
                      void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (x == nullptr || y == nullptr || yaw == nullptr) return fail(h, MCL_ERR_ARG, "x, y and yaw must not be NULL");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    cudaStream_t st = S(stream);
    const Layout &L = h->lay;
    const bool skip_stats = (on_device & MCL_SET_SKIP_STATS) != 0;
    on_device &= MCL_SET_ON_DEVICE;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    StateView &s = h->state[h->cur];
    const float *src[4] = {x, y, yaw, logw};
    float *dst[4] = {s.x, s.y, s.th, s.lw};
    for (int a = 0; a < 4; ++a) {
        if (src[a] == nullptr) continue;
        rc = copy2d(h, dst[a], L.ld, src[a], L.n, sizeof(float), L.n, L.T, kind, st);
        if (rc != MCL_OK) return rc;
    }
    if (logw == nullptr) {
        const float lw0 = 0.0f;
        const int bx = static_cast<int>(std::min<long long>((L.ld + 255) / 256, 2048));
        k_fill<<<dim3(bx, L.T), 256, 0, st>>>(s.lw, L.ld, lw0);
        MCL_LAUNCH_CHECK(h, "k_fill");
    }
    if (!on_device) MCL_CUDA(h, cudaStreamSynchronize(st));  // pageable host sources may go away
    h->have_particles = true;
    h->tile_prefix_valid = false;
    if (skip_stats) return MCL_OK;  // the caller installs statistics itself (mcl_shard_set_global)
    h->stats_from_particle0 = true;
    return reduce_only_pass(h, st);
}

int mcl_get_particles(mcl_filter *h, float *x, float *y, float *yaw, float *logw, int on_device, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "particles are not initialised");
    cudaStream_t st = S(stream);
    const Layout &L = h->lay;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const StateView &s = h->state[h->cur];
    float *dst[4] = {x, y, yaw, logw};
    const float *src[4] = {s.x, s.y, s.th, s.lw};
    for (int a = 0; a < 4; ++a) {
        if (dst[a] == nullptr) continue;
        rc = copy2d(h, dst[a], L.n, src[a], L.ld, sizeof(float), L.n, L.T, kind, st);
        if (rc != MCL_OK) return rc;
    }
    if (!on_device) MCL_CUDA(h, cudaStreamSynchronize(st));
    return MCL_OK;
}

int mcl_resample(mcl_filter *h, const double *ofs_host, uint32_t flags, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "particles are not initialised");
    if (h->cfg.n_global != h->cfg.n_particles)
        return fail(h, MCL_ERR_STATE, "a particle-sharded filter resamples through mcl_shard_total/mcl_shard_expand");
    if (ofs_host == nullptr && !(flags & MCL_PHILOX_NOISE))
        return fail(h, MCL_ERR_ARG, "ofs_host is NULL and MCL_PHILOX_NOISE is not set");
    cudaStream_t st = S(stream);
    rc = stage_inputs(h, nullptr, ofs_host, st);
    if (rc != MCL_OK) return rc;
    rc = launch_step_scan(h, nullptr, ofs_host ? h->ofs : nullptr, flags & ~MCL_NO_RESAMPLE, true, st);
    if (rc != MCL_OK) return rc;
    GatherParams g{};
    g.in = h->state[h->cur];
    g.out = h->state[h->cur ^ 1];
    g.anc = h->anc;
    g.prep = h->prep;
    g.stats = h->stats;
    g.tickets = h->tickets;
    g.n = h->lay.n;
    g.ld = h->lay.ld;
    g.n_global = static_cast<double>(h->cfg.n_global);
    const int bx = static_cast<int>(std::min<long long>((h->lay.n + 255) / 256, static_cast<long long>(h->sm_count) * 16));
    k_gather<<<dim3(bx, h->lay.T), 256, 0, st>>>(g);
    MCL_LAUNCH_CHECK(h, "k_gather");
    h->cur ^= 1;
    return MCL_OK;  // statistics are merged (the scan's statistics CTA) and updated (k_gather): not dirty
}

int mcl_predict_weight(mcl_filter *h, const double *z_host, const float *noise_dev, uint32_t flags, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (z_host == nullptr) return fail(h, MCL_ERR_ARG, "z_host is NULL");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    cudaStream_t st = S(stream);
    rc = stage_inputs(h, z_host, nullptr, st);
    if (rc != MCL_OK) return rc;
    return step_impl(h, h->z, nullptr, noise_dev, (flags | MCL_NO_RESAMPLE) & ~MCL_FORCE_RESAMPLE, st);
}

int mcl_step(mcl_filter *h, const double *z_host, const double *ofs_host, const float *noise_dev, uint32_t flags,
             void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (z_host == nullptr) return fail(h, MCL_ERR_ARG, "z_host is NULL");
    if (ofs_host == nullptr && !(flags & (MCL_PHILOX_NOISE | MCL_NO_RESAMPLE)))
        return fail(h, MCL_ERR_ARG, "ofs_host is NULL and neither MCL_PHILOX_NOISE nor MCL_NO_RESAMPLE is set");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    cudaStream_t st = S(stream);
    const double *z_direct = nullptr;
    int slot = -1;
    rc = stage_inputs(h, z_host, ofs_host, st, &z_direct, &slot);
    if (rc != MCL_OK) return rc;
    rc = step_impl(h, z_direct ? z_direct : h->z, ofs_host ? h->ofs : nullptr, noise_dev, flags, st);
    if (z_direct != nullptr) {  // the slot is free again once the step that reads it has run
        const cudaError_t e = cudaEventRecord(h->stage_ev[slot], st);
        if (rc == MCL_OK && e != cudaSuccess) MCL_CUDA(h, e);
    }
    return rc;
}

int mcl_step_dev(mcl_filter *h, const double *z_dev, const double *ofs_dev, const float *noise_dev, uint32_t flags,
                 void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (z_dev == nullptr) return fail(h, MCL_ERR_ARG, "z_dev is NULL");
    if (ofs_dev == nullptr && !(flags & (MCL_PHILOX_NOISE | MCL_NO_RESAMPLE)))
        return fail(h, MCL_ERR_ARG, "ofs_dev is NULL and neither MCL_PHILOX_NOISE nor MCL_NO_RESAMPLE is set");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    return step_impl(h, z_dev, ofs_dev, noise_dev, flags, S(stream));
}

int mcl_step_async(mcl_filter *h, const double *z_host, const double *ofs_host, const float *noise_dev, uint32_t flags,
                   void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    h->merge_follows = true;
    int rc = mcl_step(h, z_host, ofs_host, noise_dev, flags, stream);
    h->merge_follows = false;
    if (rc != MCL_OK) return rc;
    return mcl_get_stats_begin(h, stream);
}

int mcl_pose_moments(mcl_filter *h, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "particles are not initialised");
    rc = ensure_stats(h, S(stream));
    if (rc != MCL_OK) return rc;
    CovParams p{};
    p.in = h->state[h->cur];
    p.stats = h->stats;
    p.partials = h->cov_partials;
    p.tickets = h->tickets;
    p.n = h->lay.n;
    p.ld = h->lay.ld;
    p.bpt = h->lay.cov_bpt;
    k_cov<<<dim3(h->lay.cov_bpt, h->lay.T), 256, 0, S(stream)>>>(p);
    MCL_LAUNCH_CHECK(h, "k_cov");
    return MCL_OK;
}

int mcl_get_stats(mcl_filter *h, mcl_stats *out_host, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (out_host == nullptr) return fail(h, MCL_ERR_ARG, "out_host is NULL");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    rc = ensure_stats(h, S(stream));
    if (rc != MCL_OK) return rc;
    MCL_CUDA(h, cudaMemcpyAsync(out_host, h->stats, static_cast<size_t>(h->lay.T) * sizeof(mcl_stats), cudaMemcpyDeviceToHost,
                                S(stream)));
    MCL_CUDA(h, cudaStreamSynchronize(S(stream)));
    for (int t = 0; t < h->lay.T; ++t) derive_stats(out_host[t]);  // log-normaliser, mean, covariance from the sums
    return MCL_OK;
}

int mcl_get_stats_begin(mcl_filter *h, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    if (h->stats_begun - h->stats_ended >= 2) return fail(h, MCL_ERR_STATE, "two statistics read-backs are already pending");
    const size_t bytes = static_cast<size_t>(h->lay.T) * sizeof(mcl_stats);
    if (h->stats_pinned[0] == nullptr) {  // first use: the landing buffers and the events of the read-backs
        for (int i = 0; i < 2; ++i) {
            MCL_CUDA(h, cudaHostAlloc(reinterpret_cast<void **>(&h->stats_pinned[i]), bytes + 64, cudaHostAllocMapped));
            std::memset(h->stats_pinned[i], 0, bytes + 64);
            MCL_CUDA(h, cudaEventCreateWithFlags(&h->stats_event[i], cudaEventDisableTiming));
        }
    }
    const int slot = static_cast<int>(h->stats_begun & 1u);
    // A single filter whose statistics still sit in the CTA partials: the merge kernel writes the block into the
    // mapped landing buffer itself and releases a flag -- no copy-engine operation behind it, and the host polls the
    // flag instead of an event (main_pf: ~6 us of a 55 us call).
    h->stats_expect[slot] = 0u;
    if (h->lay.T == 1 && h->stats_dirty) {
        void *dev = nullptr;
        MCL_CUDA(h, cudaHostGetDevicePointer(&dev, h->stats_pinned[slot], 0));
        h->stats_seq += 1u;
        if (h->stats_seq == 0u) h->stats_seq = 1u;
        rc = ensure_stats(h, S(stream), static_cast<uint32_t *>(dev),
                          reinterpret_cast<unsigned int *>(static_cast<unsigned char *>(dev) + bytes), h->stats_seq);
        if (rc != MCL_OK) return rc;
        h->stats_expect[slot] = h->stats_seq;
    } else {
        rc = ensure_stats(h, S(stream));
        if (rc != MCL_OK) return rc;
        MCL_CUDA(h, cudaMemcpyAsync(h->stats_pinned[slot], h->stats, bytes, cudaMemcpyDeviceToHost, S(stream)));
    }
    MCL_CUDA(h, cudaEventRecord(h->stats_event[slot], S(stream)));
    h->stats_begun += 1;
    return MCL_OK;
}

int mcl_get_stats_end(mcl_filter *h, mcl_stats *out_host) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (out_host == nullptr) return fail(h, MCL_ERR_ARG, "out_host is NULL");
    if (h->stats_begun == h->stats_ended) return fail(h, MCL_ERR_STATE, "mcl_get_stats_end without mcl_get_stats_begin");
    const int slot = static_cast<int>(h->stats_ended & 1u);  // the OLDEST pending read-back
    if (h->stats_expect[slot] != 0u) {
        const size_t bytes = static_cast<size_t>(h->lay.T) * sizeof(mcl_stats);
        const volatile unsigned int *flag =
            reinterpret_cast<const volatile unsigned int *>(reinterpret_cast<const unsigned char *>(h->stats_pinned[slot]) + bytes);
        unsigned int spins = 0u;
        while (*flag != h->stats_expect[slot]) {
            if ((++spins & 0x3ffu) == 0u) {  // every 1024 polls: has the kernel failed, or finished without the flag?
                const cudaError_t e = cudaEventQuery(h->stats_event[slot]);
                if (e == cudaSuccess) {
                    if (*flag == h->stats_expect[slot]) break;
                    h->stats_ended += 1;
                    return fail(h, MCL_ERR_CUDA, "the statistics kernel completed without landing its block");
                }
                if (e != cudaErrorNotReady) {
                    h->stats_ended += 1;
                    h->err = std::string("statistics read-back: ") + cudaGetErrorString(e);
                    return MCL_ERR_CUDA;
                }
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    } else {
        MCL_CUDA(h, cudaEventSynchronize(h->stats_event[slot]));
    }
    h->stats_ended += 1;
    for (int t = 0; t < h->lay.T; ++t) {
        out_host[t] = h->stats_pinned[slot][t];
        derive_stats(out_host[t]);  // log-normaliser, mean, covariance from the sums
    }
    return MCL_OK;
}

int mcl_stats_dev(mcl_filter *h, const mcl_stats **out_dev) {
    if (h == nullptr || out_dev == nullptr) return MCL_ERR_ARG;
    // the block is current only after mcl_sync_stats(h, stream) (or a stats read-back) on the caller's stream: a merge
    // launched from here would run on the default stream and could race with the step that is still in flight
    if (h->stats_dirty) return fail(h, MCL_ERR_STATE, "mcl_stats_dev: call mcl_sync_stats(h, stream) first (a merge is pending)");
    *out_dev = h->stats;
    return MCL_OK;
}

int mcl_sync_stats(mcl_filter *h, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    return ensure_stats(h, S(stream));
}

static int export_weights(mcl_filter *h, double *w_host, uint64_t *q_host, void *stream) {
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "particles are not initialised");
    cudaStream_t st = S(stream);
    rc = ensure_stats(h, st);
    if (rc != MCL_OK) return rc;
    const Layout &L = h->lay;
    // the idle state buffer (16 B per particle) is the staging area for the 8 B per particle export
    void *stage = h->state[h->cur ^ 1].x;
    const int bx = static_cast<int>(std::min<long long>((L.n + 255) / 256, static_cast<long long>(h->sm_count) * 16));
    k_export_weights<<<dim3(bx, L.T), 256, 0, st>>>(h->state[h->cur].lw, h->stats, w_host ? static_cast<double *>(stage) : nullptr,
                                                     q_host ? static_cast<unsigned long long *>(stage) : nullptr, L.n, L.ld);
    MCL_LAUNCH_CHECK(h, "k_export_weights");
    void *dst = w_host ? static_cast<void *>(w_host) : static_cast<void *>(q_host);
    MCL_CUDA(h, cudaMemcpyAsync(dst, stage, static_cast<size_t>(L.T) * L.n * 8, cudaMemcpyDeviceToHost, st));
    MCL_CUDA(h, cudaStreamSynchronize(st));
    return MCL_OK;
}

int mcl_get_weights(mcl_filter *h, double *w_host, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (w_host == nullptr) return fail(h, MCL_ERR_ARG, "w_host is NULL");
    return export_weights(h, w_host, nullptr, stream);
}

int mcl_get_qweights(mcl_filter *h, uint64_t *q_host, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (q_host == nullptr) return fail(h, MCL_ERR_ARG, "q_host is NULL");
    return export_weights(h, nullptr, q_host, stream);
}

int mcl_get_ancestors(mcl_filter *h, uint32_t *anc_host, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (anc_host == nullptr) return fail(h, MCL_ERR_ARG, "anc_host is NULL");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    const Layout &L = h->lay;
    rc = copy2d(h, anc_host, L.n, h->anc, L.ld, sizeof(uint32_t), L.n, L.T, cudaMemcpyDeviceToHost, S(stream));
    if (rc != MCL_OK) return rc;
    MCL_CUDA(h, cudaStreamSynchronize(S(stream)));
    return MCL_OK;
}

int mcl_profile(mcl_filter *h, int enable) {
    if (h == nullptr) return MCL_ERR_ARG;
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    if (enable && h->ev[0] == nullptr)
        for (cudaEvent_t &e : h->ev) MCL_CUDA(h, cudaEventCreate(&e));
    h->profiling = enable != 0;
    return MCL_OK;
}

int mcl_profile_read(mcl_filter *h, double *ms3_host) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (ms3_host == nullptr) return fail(h, MCL_ERR_ARG, "ms3_host is NULL");
    if (!h->profiling || h->ev[0] == nullptr) return fail(h, MCL_ERR_STATE, "mcl_profile(h, 1) has not been called");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    MCL_CUDA(h, cudaEventSynchronize(h->ev[3]));
    for (int k = 0; k < 3; ++k) {
        float ms = 0.f;
        MCL_CUDA(h, cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1]));
        ms3_host[k] = ms;
    }
    return MCL_OK;
}

int mcl_profile_read_peer(mcl_filter *h, double *ms5_host) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (ms5_host == nullptr) return fail(h, MCL_ERR_ARG, "ms5_host is NULL");
    if (!h->profiling || h->ev[0] == nullptr) return fail(h, MCL_ERR_STATE, "mcl_profile(h, 1) has not been called");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    MCL_CUDA(h, cudaEventSynchronize(h->ev[5]));
    for (int k = 0; k < 5; ++k) {
        float ms = 0.f;
        MCL_CUDA(h, cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1]));
        ms5_host[k] = ms;
    }
    return MCL_OK;
}

uint64_t mcl_launch_count(const mcl_filter *h) { return h ? h->launches : 0; }

int mcl_debug_timeline(mcl_filter *h, uint64_t *stamps_host, int64_t capacity, int32_t *n_ctas) {
    if (h == nullptr) return MCL_ERR_ARG;
#ifndef MCL_TIMELINE
    (void)stamps_host; (void)capacity; (void)n_ctas;
    return fail(h, MCL_ERR_STATE, "libmcl_b200.so was built without -DMCL_TIMELINE (MCL_TIMELINE=1 python build.py --force)");
#endif
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    const size_t n_step = static_cast<size_t>(h->lay.T) * h->lay.bpt * kDbgSlots;
    const size_t n_scan = static_cast<size_t>(h->lay.T) * (h->lay.tpt + 2) * kDbgSlots;
    const size_t n = n_step + n_scan;  // k_step stamps, then k_scan stamps
    if (h->dbg == nullptr) {  // first call arms the timeline; the next steps fill it
        void *p = nullptr;
        if (cudaMalloc(&p, n * sizeof(unsigned long long)) != cudaSuccess) return fail(h, MCL_ERR_NOMEM, "cudaMalloc failed");
        cudaMemset(p, 0, n * sizeof(unsigned long long));
        h->dbg = static_cast<unsigned long long *>(p);
    }
    if (n_ctas) *n_ctas = h->lay.T * (h->lay.bpt + h->lay.tpt + 2);
    if (stamps_host == nullptr) return MCL_OK;
    if (capacity < static_cast<int64_t>(n)) return fail(h, MCL_ERR_ARG, "stamps_host too small");
    MCL_CUDA(h, cudaDeviceSynchronize());
    MCL_CUDA(h, cudaMemcpy(stamps_host, h->dbg, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return MCL_OK;
}

// ---- particle-sharded single filter -----------------------------------------------------------
int mcl_shard_total(mcl_filter *h, uint64_t *total_dev, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (total_dev == nullptr) return fail(h, MCL_ERR_ARG, "total_dev is NULL");
    if (h->lay.T != 1) return fail(h, MCL_ERR_STATE, "sharding needs n_trials == 1");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    cudaStream_t st = S(stream);
    rc = ensure_stats(h, st);
    if (rc != MCL_OK) return rc;
    MCL_CUDA(h, cudaMemsetAsync(total_dev, 0, sizeof(uint64_t), st));
    // the tile sums go where the scan keeps its descriptors (unused when the tile prefixes are known); the
    // expansion clears them again
    // Every CTA sums one contiguous segment of tiles and leaves each tile's prefix WITHIN the segment; a one-CTA
    // kernel then turns the (at most 1024) segment totals into prefixes.
    unsigned long long *tiles = h->desc[h->desc_cur];
    const int blocks_max = std::min(kMaxSegments, h->sm_count * 8);
    h->tiles_per_seg = (h->lay.tpt + blocks_max - 1) / blocks_max;
    const int blocks = (h->lay.tpt + h->tiles_per_seg - 1) / h->tiles_per_seg;
    k_shard_total<<<blocks, kScanThreads, 0, st>>>(h->state[h->cur].lw, h->lay.n, h->stats,
                                                    reinterpret_cast<unsigned long long *>(total_dev), tiles, h->seg,
                                                    h->lay.tpt, h->tiles_per_seg, nullptr, 0u);
    MCL_LAUNCH_CHECK(h, "k_shard_total");
    k_tile_prefix<<<1, 1024, 0, st>>>(h->seg, blocks);
    MCL_LAUNCH_CHECK(h, "k_tile_prefix");
    h->tile_prefix_valid = true;
    h->desc_stale = true;
    return MCL_OK;
}

int mcl_shard_expand(mcl_filter *h, uint64_t k_offset, double ofs, float *send_x_dev, float *send_y_dev,
                     float *send_yaw_dev, int64_t cap, int64_t *first_slot, int64_t *n_slots, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (cap < 0 || (cap > 0 && (send_x_dev == nullptr || send_y_dev == nullptr || send_yaw_dev == nullptr)))
        return fail(h, MCL_ERR_ARG, "NULL send buffer or negative capacity");
    if ((first_slot == nullptr) != (n_slots == nullptr)) return fail(h, MCL_ERR_ARG, "first_slot and n_slots go together");
    if (h->lay.T != 1) return fail(h, MCL_ERR_STATE, "sharding needs n_trials == 1");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    cudaStream_t st = S(stream);
    const long long NG = h->cfg.n_global;
    const long long lo = h->cfg.slot_base, hi = lo + h->cfg.n_particles;
    const long long s0 = host_count_le(k_offset, h->delta, ofs, NG);
    ScanParams sp;
    fill_scan(h, sp);
    sp.total = h->total;
    sp.use_override = 1;
    sp.ofs_override = ofs;
    sp.k_offset = k_offset;
    sp.slot_first = s0;
    sp.anc_cap = cap;
    sp.own_lo = lo;
    sp.own_hi = hi;
    sp.send_hi_off = (lo > s0) ? lo - s0 : 0;   // size of the piece below own_lo when the range reaches own_lo
    sp.send_hi_base = (hi > s0) ? hi : s0;
    sp.src = h->state[h->cur];
    sp.ox = send_x_dev;
    sp.oy = send_y_dev;
    sp.oth = send_yaw_dev;
    rc = launch_shard_scan(h, sp, kScanExpand, st);
    if (rc != MCL_OK) return rc;
    if (first_slot == nullptr) return MCL_OK;  // the caller derives the range from the allgathered totals
    unsigned long long total = 0;
    MCL_CUDA(h, cudaMemcpyAsync(&total, h->total, sizeof total, cudaMemcpyDeviceToHost, st));
    MCL_CUDA(h, cudaStreamSynchronize(st));
    const bool last = (hi == NG);
    const long long s1 = last ? NG : host_count_le(k_offset + total, h->delta, ofs, NG);
    *first_slot = s0;
    *n_slots = s1 - s0;
    const long long off_rank = std::max(0ll, std::min(lo, s1) - s0) + std::max(0ll, s1 - std::max(hi, s0));
    if (off_rank > cap) return fail(h, MCL_ERR_NOMEM, "mcl_shard_expand: more off-rank slots than `cap`");
    return MCL_OK;
}

int mcl_shard_step(mcl_filter *h, const double *z_host, const float *noise_dev, const float *recv_x_dev,
                   const float *recv_y_dev, const float *recv_yaw_dev, int64_t n_prefix, int64_t n_suffix, uint32_t flags,
                   void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (z_host == nullptr) return fail(h, MCL_ERR_ARG, "z_host is NULL");
    if (h->lay.T != 1) return fail(h, MCL_ERR_STATE, "sharding needs n_trials == 1");
    if (n_prefix < 0 || n_suffix < 0 || n_prefix + n_suffix > h->lay.n) return fail(h, MCL_ERR_ARG, "bad prefix/suffix");
    if ((n_prefix + n_suffix > 0) && (recv_x_dev == nullptr || recv_y_dev == nullptr || recv_yaw_dev == nullptr))
        return fail(h, MCL_ERR_ARG, "received poses are NULL");
    if (h->L <= 0) return fail(h, MCL_ERR_STATE, "mcl_set_landmarks has not been called");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "particles are not initialised");
    if (noise_dev == nullptr && !(flags & MCL_PHILOX_NOISE))
        return fail(h, MCL_ERR_ARG, "noise_dev is NULL and MCL_PHILOX_NOISE is not set");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    cudaStream_t st = S(stream);
    rc = stage_inputs(h, z_host, nullptr, st);
    if (rc != MCL_OK) return rc;
    // scalars only; the gate is forced: the resampling itself was done by mcl_shard_expand + the exchange
    rc = launch_step_scan(h, h->z, nullptr, (flags | MCL_FORCE_RESAMPLE) & ~MCL_NO_RESAMPLE, false, st);
    if (rc != MCL_OK) return rc;
    StepWindow win;
    win.lo = static_cast<uint32_t>(n_prefix);
    win.hi = static_cast<uint32_t>(h->lay.n - n_suffix);
    win.x = recv_x_dev;
    win.y = recv_y_dev;
    win.th = recv_yaw_dev;
    rc = launch_step(h, noise_dev, flags, false, st, &win);
    if (rc != MCL_OK) return rc;
    h->cur ^= 1;
    h->step += 1;
    return MCL_OK;
}

int mcl_restore(mcl_filter *h, const mcl_stats *stats_host, uint32_t step, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (stats_host == nullptr) return fail(h, MCL_ERR_ARG, "stats_host is NULL");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "mcl_restore follows mcl_set_particles (MCL_SET_SKIP_STATS)");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    MCL_CUDA(h, cudaMemcpyAsync(h->stats, stats_host, static_cast<size_t>(h->lay.T) * sizeof(mcl_stats),
                                cudaMemcpyHostToDevice, S(stream)));
    MCL_CUDA(h, cudaStreamSynchronize(S(stream)));
    h->step = step;
    h->stats_from_particle0 = false;
    h->stats_dirty = false;
    h->tile_prefix_valid = false;
    if (h->plan != nullptr) {
        // a shard of a device-driven sharded filter: the resampling gate of the next step lives on the device
        // (the statistics installed here are the GLOBAL ones of the filter that was saved)
        const int32_t gate = (stats_host[0].ess < h->cfg.ess_threshold) ? 1 : 0;
        MCL_CUDA(h, cudaMemcpy(reinterpret_cast<unsigned char *>(h->plan) + offsetof(mcl_shard_plan, gate), &gate, sizeof gate,
                               cudaMemcpyHostToDevice));
    }
    return MCL_OK;
}

int mcl_shard_merge(mcl_filter *h, const mcl_stats *records_dev, int n_shards, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (records_dev == nullptr || n_shards < 1) return fail(h, MCL_ERR_ARG, "records_dev is NULL or n_shards < 1");
    if (h->lay.T != 1) return fail(h, MCL_ERR_STATE, "sharding needs n_trials == 1");
    if (static_cast<long long>(n_shards) * h->cfg.n_particles != h->cfg.n_global)
        return fail(h, MCL_ERR_ARG, "n_shards * n_particles != n_global (shards must be equal)");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    ShardMergeParams mp{};
    mp.rec = records_dev;
    mp.G = n_shards;
    mp.m = h->cfg.n_particles;
    mp.out = h->stats;
    mp.apply_rule = 0;  // (host-driven path: the records do not carry particle 0's pose)
    k_shard_merge<<<1, 32, 0, S(stream)>>>(mp);
    MCL_LAUNCH_CHECK(h, "k_shard_merge");
    h->tile_prefix_valid = false;
    h->stats_from_particle0 = false;
    h->stats_dirty = false;
    return MCL_OK;
}

int mcl_shard_set_global(mcl_filter *h, const mcl_stats *merged_host, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (merged_host == nullptr) return fail(h, MCL_ERR_ARG, "merged_host is NULL");
    if (h->lay.T != 1) return fail(h, MCL_ERR_STATE, "sharding needs n_trials == 1");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    MCL_CUDA(h, cudaMemcpyAsync(h->stats, merged_host, sizeof(mcl_stats), cudaMemcpyHostToDevice, S(stream)));
    MCL_CUDA(h, cudaStreamSynchronize(S(stream)));
    h->tile_prefix_valid = false;
    h->stats_from_particle0 = false;
    h->stats_dirty = false;
    return MCL_OK;
}


// ---- device-driven sharded filter (peer memory) ---------------------------------------------------
namespace {
struct PeerRegion {
    size_t off_box, off_rx, off_ry, off_rth, total;
};
PeerRegion peer_region_layout(long long ld) {
    PeerRegion r;
    size_t o = 0;
    r.off_box = o; o = align_up(o + sizeof(PeerBox), kAlign);
    r.off_rx = o;  o = align_up(o + static_cast<size_t>(ld) * sizeof(float), kAlign);
    r.off_ry = o;  o = align_up(o + static_cast<size_t>(ld) * sizeof(float), kAlign);
    r.off_rth = o; o = align_up(o + static_cast<size_t>(ld) * sizeof(float), kAlign);
    r.total = o;
    return r;
}
int peer_ready(mcl_filter *h) {
    if (h->peer_region == nullptr || !h->peer_connected)
        return fail(h, MCL_ERR_STATE, "mcl_shard_peer_alloc / mcl_shard_peer_connect have not been called");
    return MCL_OK;
}
void fill_xchg(mcl_filter *h, XchgParams &x, uint32_t flags, int mode) {
    x = XchgParams{};
    x.pv = h->peer_view;
    x.plan = h->plan;
    x.stats = h->stats;
    x.epoch = h->xchg_epoch;
    x.flags = flags;
    x.mode = mode;
    x.ess_threshold = h->cfg.ess_threshold;
    x.seed = h->cfg.seed;
    x.step = h->step;
    x.n_global = h->cfg.n_global;
    x.delta = h->delta;
}
// statistics exchange: push = merge of the own partials + record to every box, wait = merge of all records
int launch_xchg_stats(mcl_filter *h, int mode, cudaStream_t st) {
    XchgParams x;
    fill_xchg(h, x, 0u, mode);
    x.cur = h->state[h->cur];
    if ((mode & kXchgPush) && h->stats_dirty) {
        x.partials = h->partials;
        x.n_partials = h->lay.bpt;
        h->stats_dirty = false;
    }
    k_xchg_stats<<<1, kStepThreads, 0, st>>>(x);
    MCL_LAUNCH_CHECK(h, "k_xchg_stats");
    h->stats_from_particle0 = false;
    h->tile_prefix_valid = false;
    return MCL_OK;
}
}  // namespace

int mcl_shard_peer_alloc(mcl_filter *h, int rank, int n_shards, unsigned char *ipc_handle_out) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (n_shards < 1 || n_shards > kMaxShards || rank < 0 || rank >= n_shards)
        return fail(h, MCL_ERR_ARG, "rank / n_shards out of range (at most MCL_MAX_SHARDS shards)");
    if (h->lay.T != 1) return fail(h, MCL_ERR_STATE, "sharding needs n_trials == 1");
    if (static_cast<long long>(n_shards) * h->cfg.n_particles != h->cfg.n_global ||
        static_cast<long long>(rank) * h->cfg.n_particles != h->cfg.slot_base)
        return fail(h, MCL_ERR_ARG, "shards must be equal: n_shards * n_particles == n_global, slot_base == rank * n_particles");
    if (h->peer_region != nullptr) return fail(h, MCL_ERR_STATE, "mcl_shard_peer_alloc was already called");
    static_assert(sizeof(cudaIpcMemHandle_t) == MCL_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    const PeerRegion r = peer_region_layout(h->lay.ld);
    void *p = nullptr;
    // cudaMalloc, not the caller's workspace: an IPC handle names a whole allocation
    if (cudaMalloc(&p, r.total) != cudaSuccess) {
        cudaGetLastError();
        return fail(h, MCL_ERR_NOMEM, "cudaMalloc of the peer region failed");
    }
    MCL_CUDA(h, cudaMemset(p, 0, r.total));
    h->peer_region = static_cast<unsigned char *>(p);
    h->peer_region_bytes = r.total;
    h->peer_rank = rank;
    h->peer_G = n_shards;
    void *q = nullptr;
    MCL_CUDA(h, cudaMalloc(&q, sizeof(PeerView)));
    h->peer_view = static_cast<PeerView *>(q);
    MCL_CUDA(h, cudaMalloc(&q, sizeof(ShardPlan)));
    MCL_CUDA(h, cudaMemset(q, 0, sizeof(ShardPlan)));
    h->plan = static_cast<ShardPlan *>(q);
    if (ipc_handle_out != nullptr) {
        cudaIpcMemHandle_t ih;
        MCL_CUDA(h, cudaIpcGetMemHandle(&ih, p));
        std::memcpy(ipc_handle_out, &ih, sizeof ih);
    }
    return MCL_OK;
}

int mcl_shard_peer_connect(mcl_filter *h, const unsigned char *ipc_handles, mcl_filter *const *local_peers) {
    if (h == nullptr) return MCL_ERR_ARG;
    if ((ipc_handles == nullptr) == (local_peers == nullptr))
        return fail(h, MCL_ERR_ARG, "exactly one of ipc_handles / local_peers must be given");
    if (h->peer_region == nullptr) return fail(h, MCL_ERR_STATE, "mcl_shard_peer_alloc has not been called");
    if (h->peer_connected) return fail(h, MCL_ERR_STATE, "mcl_shard_peer_connect was already called");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    const PeerRegion r = peer_region_layout(h->lay.ld);
    PeerView pv{};
    pv.me = h->peer_rank;
    pv.G = h->peer_G;
    pv.m = h->cfg.n_particles;
    for (int g = 0; g < h->peer_G; ++g) {
        unsigned char *base = nullptr;
        if (g == h->peer_rank) {
            base = h->peer_region;
        } else if (local_peers != nullptr) {
            mcl_filter *o = local_peers[g];
            if (o == nullptr || o->peer_region == nullptr || o->peer_G != h->peer_G || o->peer_rank != g ||
                o->cfg.n_particles != h->cfg.n_particles)
                return fail(h, MCL_ERR_ARG, "local_peers[g] is not shard g of this filter (mcl_shard_peer_alloc first)");
            if (o->cfg.device != h->cfg.device) {
                int can = 0;
                MCL_CUDA(h, cudaDeviceCanAccessPeer(&can, h->cfg.device, o->cfg.device));
                if (!can) return fail(h, MCL_ERR_NODEVICE, "no peer access between the shards' devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(o->cfg.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) MCL_CUDA(h, e);
                cudaGetLastError();
            }
            base = o->peer_region;
        } else {
            cudaIpcMemHandle_t ih;
            std::memcpy(&ih, ipc_handles + static_cast<size_t>(g) * MCL_IPC_HANDLE_BYTES, sizeof ih);
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                char buf[256];
                snprintf(buf, sizeof buf, "cudaIpcOpenMemHandle of shard %d failed: %s", g, cudaGetErrorString(e));
                return fail(h, MCL_ERR_CUDA, buf);
            }
            h->peer_ipc_opened[g] = p;
            base = static_cast<unsigned char *>(p);
        }
        pv.box[g] = reinterpret_cast<PeerBox *>(base + r.off_box);
        pv.rx[g] = reinterpret_cast<float *>(base + r.off_rx);
        pv.ry[g] = reinterpret_cast<float *>(base + r.off_ry);
        pv.rth[g] = reinterpret_cast<float *>(base + r.off_rth);
    }
    MCL_CUDA(h, cudaMemcpy(h->peer_view, &pv, sizeof pv, cudaMemcpyHostToDevice));
    h->peer_connected = true;
    return MCL_OK;
}

int mcl_shard_sync_peer(mcl_filter *h, uint32_t stages, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    int rc = peer_ready(h);
    if (rc != MCL_OK) return rc;
    rc = use_device(h);
    if (rc != MCL_OK) return rc;
    const bool push = (stages & MCL_STAGE_C) != 0, wait = (stages & MCL_STAGE_D) != 0;
    if (!push && !wait) return fail(h, MCL_ERR_ARG, "stages must contain MCL_STAGE_C and / or MCL_STAGE_D");
    if (push) h->xchg_epoch += 1;
    return launch_xchg_stats(h, (push ? kXchgPush : 0) | (wait ? kXchgWait : 0), S(stream));
}

int mcl_shard_step_peer(mcl_filter *h, const double *z_host, const double *ofs_host, const float *noise_dev,
                        uint32_t flags, uint32_t stages, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    int rc = peer_ready(h);
    if (rc != MCL_OK) return rc;
    if ((stages & MCL_STAGE_ALL) == 0 || (stages & ~MCL_STAGE_ALL) != 0) return fail(h, MCL_ERR_ARG, "bad stages");
    if (h->L <= 0) return fail(h, MCL_ERR_STATE, "mcl_set_landmarks has not been called");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "particles are not initialised");
    if ((stages & (MCL_STAGE_B | MCL_STAGE_C)) && z_host == nullptr) return fail(h, MCL_ERR_ARG, "z_host is NULL");
    if ((stages & MCL_STAGE_C) && noise_dev == nullptr && !(flags & MCL_PHILOX_NOISE))
        return fail(h, MCL_ERR_ARG, "noise_dev is NULL and MCL_PHILOX_NOISE is not set");
    if (ofs_host != nullptr && !(ofs_host[0] >= 0.0 && ofs_host[0] < h->delta))
        return fail(h, MCL_ERR_ARG, "comb offset outside [0, 1/N): pass rand() * (1 / n_global)");
    rc = use_device(h);
    if (rc != MCL_OK) return rc;
    cudaStream_t st = S(stream);
    const bool A = stages & MCL_STAGE_A, B = stages & MCL_STAGE_B, Cc = stages & MCL_STAGE_C, D = stages & MCL_STAGE_D;
    const bool may_resample = !(flags & MCL_NO_RESAMPLE);
    if (A) {
        if (h->stats_dirty) return fail(h, MCL_ERR_STATE, "statistics were not exchanged: mcl_shard_sync_peer first");
        h->xchg_epoch += 1;
    }
    const bool prof = h->profiling && stages == MCL_STAGE_ALL;
    if (prof) MCL_CUDA(h, cudaEventRecord(h->ev[0], st));
    if (may_resample && A) {
        // first pass over the weights under the GLOBAL map: tile prefixes inside contiguous segments + segment totals
        const int blocks_max = std::min(kMaxSegments, h->sm_count * 8);
        h->tiles_per_seg = (h->lay.tpt + blocks_max - 1) / blocks_max;
        const int blocks = (h->lay.tpt + h->tiles_per_seg - 1) / h->tiles_per_seg;
        k_shard_total<<<blocks, kScanThreads, 0, st>>>(h->state[h->cur].lw, h->lay.n, h->stats, nullptr, h->desc[h->desc_cur],
                                                        h->seg, h->lay.tpt, h->tiles_per_seg, h->plan, flags);
        MCL_LAUNCH_CHECK(h, "k_shard_total");
        h->desc_stale = true;
    }
    if (A || B) {  // (also when the host forbids resampling: the kernel then only records that in the plan)
        XchgParams x;
        fill_xchg(h, x, flags, (A ? kXchgPush : 0) | (B ? kXchgWait : 0));
        x.seg = h->seg;
        x.n_seg = (h->lay.tpt + h->tiles_per_seg - 1) / h->tiles_per_seg;
        x.have_ofs = ofs_host != nullptr;
        x.ofs_host = ofs_host ? ofs_host[0] : 0.0;
        k_xchg_total<<<1, 1024, 0, st>>>(x);
        MCL_LAUNCH_CHECK(h, "k_xchg_total");
    }
    if (prof) MCL_CUDA(h, cudaEventRecord(h->ev[1], st));
    if (B) {
        // per-step scalars (one CTA; the gate comes from the plan).  They depend on the statistics and the plan only,
        // so they are enqueued BEFORE the expansion, which is launched programmatically behind them and runs beside them.
        rc = stage_inputs(h, z_host, nullptr, st);
        if (rc != MCL_OK) return rc;
        ScanParams sp;
        fill_scan(h, sp);
        fill_prep(h, sp.prep, h->z, nullptr, flags);
        sp.do_prep = 1;
        h->prep_epoch += 1u;
        sp.prep_epoch = h->prep_epoch;
        sp.trig = h->trig;
        sp.plan = h->plan;
        sp.xflags = flags;
        launch_pdl(k_scan<kScanAncestors, 8>, dim3(1, 1), dim3(kScanTile / 8), 0, st, false, sp);
        MCL_LAUNCH_CHECK(h, "k_scan");
    }
    if (may_resample && B) {
        // second pass: global slot numbers, every tile starting from segment prefix + tile prefix (no look-back);
        // own slots -> ancestors, off-rank slots -> poses into the owner's staging
        ScanParams sp;
        fill_scan(h, sp);
        sp.use_override = 1;
        sp.own_lo = h->cfg.slot_base;
        sp.own_hi = h->cfg.slot_base + h->cfg.n_particles;
        sp.src = h->state[h->cur];
        sp.plan = h->plan;
        sp.pv = h->peer_view;
        sp.xflags = flags;
        sp.tile_prefix = sp.desc;
        sp.seg_prefix = h->seg;
        sp.tiles_per_seg = h->tiles_per_seg;
        sp.persistent = 1;
        sp.no_dep_sync = 1;
        dim3 grid(std::min(h->lay.tpt, h->sm_count * scan_ctas_per_sm(16)), 1);
        launch_pdl(k_scan<kScanExpand, 16>, grid, dim3(kScanTile / 16), 0, st, pdl_enabled(), sp);
        MCL_LAUNCH_CHECK(h, "k_scan");
    }
    if (may_resample && (B || Cc)) {
        XchgParams x;
        fill_xchg(h, x, flags, (B ? kXchgPush : 0) | (Cc ? kXchgWait : 0));
        k_xchg_done<<<1, 32, 0, st>>>(x);
        MCL_LAUNCH_CHECK(h, "k_xchg_done");
    }
    if (prof) MCL_CUDA(h, cudaEventRecord(h->ev[2], st));
    if (Cc) {
        StepWindow win;
        win.lo = 0u;
        win.hi = static_cast<uint32_t>(h->lay.n);
        const PeerRegion pr = peer_region_layout(h->lay.ld);
        win.x = reinterpret_cast<const float *>(h->peer_region + pr.off_rx);
        win.y = reinterpret_cast<const float *>(h->peer_region + pr.off_ry);
        win.th = reinterpret_cast<const float *>(h->peer_region + pr.off_rth);
        if (prof) MCL_CUDA(h, cudaEventRecord(h->ev[3], st));
        rc = launch_step(h, noise_dev, flags, false, st, &win, h->plan);
        if (rc != MCL_OK) return rc;
        if (prof) MCL_CUDA(h, cudaEventRecord(h->ev[4], st));
        h->cur ^= 1;
        h->step += 1;
    }
    if (Cc || D) {
        rc = launch_xchg_stats(h, (Cc ? kXchgPush : 0) | (D ? kXchgWait : 0), st);
        if (rc != MCL_OK) return rc;
    }
    if (prof) MCL_CUDA(h, cudaEventRecord(h->ev[5], st));
    return MCL_OK;
}

int mcl_shard_get_plan(mcl_filter *h, mcl_shard_plan *out_host, void *stream) {
    if (h == nullptr) return MCL_ERR_ARG;
    if (out_host == nullptr) return fail(h, MCL_ERR_ARG, "out_host is NULL");
    if (h->plan == nullptr) return fail(h, MCL_ERR_STATE, "mcl_shard_peer_alloc has not been called");
    int rc = use_device(h);
    if (rc != MCL_OK) return rc;
    MCL_CUDA(h, cudaMemcpyAsync(out_host, h->plan, sizeof(mcl_shard_plan), cudaMemcpyDeviceToHost, S(stream)));
    MCL_CUDA(h, cudaStreamSynchronize(S(stream)));
    return MCL_OK;
}

}  // extern "C"
