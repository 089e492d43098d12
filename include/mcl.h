/*
 * mcl.h -- C ABI of the B200 Monte-Carlo-localisation engine (libmcl_b200.so).
 *
 * The reference (takuyani/SLAM-Robot_Simu) has no FFI: its boundary for this path is the Python
 * class ParticleFilter(period_ms) / main_pf() (particle_filter.py:21, :86-119).  Each entry point
 * below names the reference method it replaces; the Python mirror that binds them with ctypes is
 * slam-robot_simu_b200/particle_filter.py, and INTEGRATION.md shows the stub a maintainer adds.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no C++/torch types cross this boundary.
 *   - every call returns MCL_OK (0) or a negative mcl_status; nothing throws across the ABI;
 *     mcl_last_error(h) gives the text of the last failure on that handle.
 *   - "dev" pointers are CUDA device pointers owned by the caller (e.g. torch tensor data_ptr());
 *     "host" pointers are ordinary process memory.  `stream` is a cudaStream_t passed as void*.
 *   - all work is enqueued on `stream` and is asynchronous unless the call says "synchronises".
 *   - one handle per (device, stream) user; calls on one handle are not thread-safe, different
 *     handles are independent.
 *   - there is no CPU fallback: creating a filter without a usable sm_100 device fails.
 *
 * Particle state is SoA fp32 (x, y, yaw, log-weight), `n_trials` independent filters of
 * `n_particles` each, trial t at offset t * mcl_stride(h).  Likelihood arithmetic is fp64 in
 * registers; resampling thresholds are fp64 and the cumulative weights are exact 64-bit integers
 * on a 2^-52 grid (bit-exact against the reference walk, particle_filter.py:212-221).
 */
#ifndef MCL_B200_H
#define MCL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCL_ABI_VERSION 3

typedef struct mcl_filter mcl_filter; /* opaque */

typedef enum mcl_status {
    MCL_OK = 0,
    MCL_ERR_ARG = -1,      /* bad argument / NULL pointer / size out of range          */
    MCL_ERR_CUDA = -2,     /* a CUDA runtime call or kernel launch failed                */
    MCL_ERR_STATE = -3,    /* call sequence error (e.g. no landmarks set)               */
    MCL_ERR_NOMEM = -4,    /* workspace too small / allocation failed                    */
    MCL_ERR_NODEVICE = -5  /* no CUDA device, or not compute capability 10.x             */
} mcl_status;

/* likelihood kernels (particle_filter.py:170-198 is the Cartesian model) */
typedef enum mcl_likelihood {
    MCL_LIK_PAIR = 0,   /* every particle x every landmark, fp64, landmarks staged in shared memory */
    MCL_LIK_MOMENT = 1, /* same result from 13 per-step landmark moments (exact identity), O(1)/particle */
    MCL_LIK_RB = 2,     /* range/bearing residuals with limit_angle wrap (extension; sensor model
                           graph_based_slam.py:150-167) */
    MCL_LIK_RB_FAN = 3, /* the fan-shaped scan sensor (ScanSensor, graph_based_slam.py:73-167): only DETECTED
                           landmarks are scored (z rows with a NaN distance are skipped), distance noise grows
                           with the distance (:163), and a landmark outside the particle's own fan (:155-160)
                           costs fan_miss_chi2 / 2 instead of its residuals.  fp32 per pair (MUFU sqrt, atan2f):
                           the likelihood-heavy kernel; sigma_obs = (R_Dist gain, direction sigma [rad]) */
    MCL_LIK_PAIR_F32 = 4, /* MCL_LIK_PAIR evaluated in fp32 ABOUT THE REFERENCE POSE (the previous MAP estimate pushed
                           through the motion): a particle's residual is the reference pose's residual u_l plus a
                           small d_l, and the kernel sums d_l (d_l + 2 u_l) -- the difference of the two chi^2, no
                           cancellation -- 8 fp32 operations per pair, two particles per FFMA2 */
    MCL_LIK_RB_F32 = 5  /* MCL_LIK_RB for ALL landmarks in fp32 about the reference pose (MUFU sqrt / rcp, series
                           atan2 of the small bearing change, two particles per FFMA2): the FP32 / SFU-bound kernel
                           of BASELINE configs[4].  64 bytes of shared memory per landmark: mcl_set_landmarks
                           returns MCL_ERR_ARG beyond 2048 landmarks */
} mcl_likelihood;

typedef enum mcl_motion {
    MCL_MOTION_A = 0, /* particle_filter.py:121-142 (__f) + additive N(0,Q) noise (:165-166)  */
    MCL_MOTION_B = 1  /* motion_model.py:31-62 (MotionModel.moveWithNoise), batched            */
} mcl_motion;

/* flags for mcl_step / mcl_resample */
#define MCL_FORCE_RESAMPLE 1u /* resample regardless of the ESS gate (benchmark mode)           */
#define MCL_NO_RESAMPLE 2u    /* skip the resampling stage                                      */
#define MCL_PHILOX_NOISE 4u   /* draw motion noise (and the comb offset when ofs==NULL) on device */

/* Constants of ParticleFilter.__init__ (particle_filter.py:30-84). mcl_config_default fills the
 * reference literals; every field can then be overridden. */
typedef struct mcl_config {
    uint32_t struct_size;   /* = sizeof(mcl_config), checked                               */
    int32_t device;         /* CUDA device ordinal                                         */
    int64_t n_particles;    /* NP, particles per trial held by THIS handle (:31)           */
    int32_t n_trials;       /* independent filters batched in one handle (>=1)             */
    int32_t max_landmarks;  /* capacity of the landmark table (<= 4096)                    */
    double dt;              /* DT_s = period_ms / 1000 (:30)                               */
    double vel;             /* VEL_mps (:48)                                               */
    double yaw_rate;        /* YAW_RATE_rps (:47)                                          */
    double q_std[3];        /* sqrt(diag(Q)): x [m], y [m], yaw [rad] (:62-65)             */
    double sigma_obs[2];    /* sqrt(R[0][0]), sqrt(R[1][1]) (:68-70, :179-180); in MCL_LIK_RB
                               these are (sigma_range, sigma_bearing)                      */
    double ess_threshold;   /* ESS_TH = NP / 100 (:33)                                     */
    double motion_b_alpha[6]; /* a1..a6 of MotionModel (motion_model.py:20-29)            */
    int32_t likelihood;     /* mcl_likelihood                                              */
    int32_t motion;         /* mcl_motion                                                  */
    uint64_t seed;          /* Philox key                                                  */
    /* particle sharding of ONE filter over several handles/GPUs (n_trials must be 1):     */
    int64_t n_global;       /* particles in the whole filter (== n_particles if unsharded) */
    int64_t slot_base;      /* global index of this handle's first particle                */
    /* MCL_LIK_RB_FAN (graph_based_slam.py:155-160, :901-902):                              */
    double fan_range;       /* scan range [m] (SCN_SENS_RANGE_m = 15)                      */
    double fan_half_angle;  /* half opening angle of the fan [rad] (SCN_SENS_ANGLE_rps = 80 deg) */
    double fan_miss_chi2;   /* chi^2 charged for a detected landmark outside the particle's fan */
    int32_t underflow_rule; /* 1 (default): mirror particle_filter.py:234-237 -- when every pw * bn would underflow to 0
                               in the reference's fp64 product the weights are reset to 1/NP (mcl_stats.underflow);
                               0: pure log-domain weights (never reset unless every stored log-weight is -inf / NaN).
                               With hundreds of landmarks the reference underflows on any un-resampled second step
                               (SURVEY.md section 0.1): long runs on such maps want 0. */
    int32_t reserved0;
} mcl_config;

/* Per-trial scalars of one step; the reference returns max_idx/max_val/x_est from main_pf
 * (particle_filter.py:115-119); mean/cov are the extension that feeds
 * mylib/error_ellipse.py:39-55 (ErrorEllipse.calc_error_ellipse(cov[0:2,0:2])). */
typedef struct mcl_stats {
    double log_norm;   /* log sum_i exp(logw_i): normalised weight = exp(logw_i - log_norm) */
    double ess;        /* 1 / sum w_i^2 (:210)                                              */
    double max_weight; /* max_val (:115)                                                    */
    int64_t max_idx;   /* max_idx, first occurrence (:116); local index when sharded        */
    double est[3];     /* x_est = px[:, max_idx] (:117)                                     */
    double mean[3];    /* weighted mean, circular for yaw                                   */
    double cov[9];     /* weighted 3x3 covariance, row-major, yaw residual wrapped           */
    double sum_w;      /* s1: sum exp(logw - max_logw)  (partials for sharded merge)        */
    double sum_w2;     /* s2: sum exp(2 (logw - max_logw))                                  */
    double max_logw;   /* max_i logw_i                                                      */
    double raw[12];    /* un-normalised sums about `ref`: dx dy dth sin cos xx xy yy xth yth thth,
                          then [11] = stored log-weight of the MAP particle                  */
    double ref[3];     /* reference pose the moment sums are taken about                    */
    int32_t resampled; /* 1 if this step's resampling stage fired                           */
    int32_t step;      /* number of mcl_step/mcl_predict_weight calls so far                */
    int32_t derived;   /* 1 once log_norm, max_weight, mean and cov have been derived from the sums:
                          the kernels write only the sums; mcl_get_stats derives the rest on the host */
    int32_t underflow; /* 1: the reference's degenerate-weight rule fired (particle_filter.py:234-237): every product
                          pw * bn underflowed to 0 in fp64 (or every stored log-weight is -inf / NaN), the sum is 0,
                          the quotient NaN, and the weights were reset to 1/NP.  ess = NP, max_idx = 0, est = particle
                          0, max_weight = 1/NP; the stored log-weights are to be read as all-equal (the next step,
                          the resampler and mcl_get_weights do so); mean / cov keep the log-domain weights. */
    double abs_shift;  /* log of the reference's UN-normalised weight pw * bn minus the stored log-weight (what the
                          underflow test adds to max_logw); 1e300 when no likelihood step defines it */
} mcl_stats;

int mcl_abi_version(void);
const char *mcl_status_string(int status);
const char *mcl_last_error(const mcl_filter *h);

/* ParticleFilter.__init__ literals (particle_filter.py:30-84) for `period_ms`. */
int mcl_config_default(mcl_config *cfg, double period_ms);

/* Bytes of device workspace a filter of this config needs (state double-buffers, ancestors,
 * scan tile descriptors, reduction partials). */
int mcl_workspace_bytes(const mcl_config *cfg, size_t *bytes);

/* Create a filter.  `workspace_dev` is caller-owned device memory of at least
 * mcl_workspace_bytes (256-byte aligned), e.g. a torch uint8 tensor; pass NULL to let the library
 * cudaMalloc it.  Replaces ParticleFilter.__init__ (:21-84) minus the state fill. */
int mcl_create(const mcl_config *cfg, void *workspace_dev, size_t workspace_bytes, mcl_filter **out);
int mcl_destroy(mcl_filter *h);

int64_t mcl_stride(const mcl_filter *h); /* elements between consecutive trials in the state arrays */

/* self.__LM (:39-43): host array [L][2] fp64, shared by all trials. */
int mcl_set_landmarks(mcl_filter *h, const double *lm_xy_host, int32_t n_landmarks, void *stream);

/* px = repmat(x_true, 1, NP); pw = 1/NP (:81-84).  pose_host is [n_trials][3]. */
int mcl_init_particles(mcl_filter *h, const double *pose_host, void *stream);

/* Overwrite / read the particle state.  Arrays are [n_trials][n_particles] fp32 (dense, no
 * stride); logw may be NULL (set: uniform; get: skipped).  `on_device` is a bit set: MCL_SET_ON_DEVICE
 * = the pointers are device pointers; MCL_SET_SKIP_STATS (set only) = do not recompute the statistics
 * (the sharded filter installs global ones with mcl_shard_set_global).  get synchronises when the
 * pointers are host pointers. */
#define MCL_SET_ON_DEVICE 1
#define MCL_SET_SKIP_STATS 2
int mcl_set_particles(mcl_filter *h, const float *x, const float *y, const float *yaw, const float *logw,
                      int on_device, void *stream);
int mcl_get_particles(mcl_filter *h, float *x, float *y, float *yaw, float *logw, int on_device, void *stream);

/* ---- stage calls (one per reference method) --------------------------------------------------
 * __resampling (:200-224).  ofs_host: [n_trials] comb offsets = rand() * NP_RECIP (:214), or NULL
 * with MCL_PHILOX_NOISE.  Fires per trial when ess < ess_threshold (:211) or MCL_FORCE_RESAMPLE. */
int mcl_resample(mcl_filter *h, const double *ofs_host, uint32_t flags, void *stream);

/* __predict (:156-168) + __likelihood/__normalize (:170-198, :226-237) + MAP estimate (:115-117).
 * z_host: [n_trials][L][2] fp64 observations z_l (:151-153) -- (range, bearing) in MCL_LIK_RB and
 * MCL_LIK_RB_FAN (there a NaN range marks a landmark the sensor did not detect).
 * noise_dev: device fp32 [3][n_trials][n_particles] = the draw v of :165 (rows x, y, yaw), or NULL
 * with MCL_PHILOX_NOISE.  For MCL_MOTION_B the three rows are standard normals (vel, omg, gam). */
int mcl_predict_weight(mcl_filter *h, const double *z_host, const float *noise_dev, uint32_t flags, void *stream);

/* main_pf (:86-119) without the ground-truth line: resample -> predict -> weight -> estimate. */
int mcl_step(mcl_filter *h, const double *z_host, const double *ofs_host, const float *noise_dev, uint32_t flags,
             void *stream);
/* Same, with z and ofs already resident on the device: no host traffic and no synchronisation, so a run of steps can
 * be captured in a CUDA graph (tests: test_steps_captured_in_a_cuda_graph).  A captured step carries the buffer parity
 * and the step / Philox counters of the call that was captured: a graph stands for THOSE steps, it is not a loop body. */
int mcl_step_dev(mcl_filter *h, const double *z_dev, const double *ofs_dev, const float *noise_dev, uint32_t flags,
                 void *stream);

/* mcl_step followed by mcl_get_stats_begin in ONE crossing of the ABI: the observation is staged through a pinned
 * ring (no host synchronisation; an observation of at most 4 KB is read by the step straight from that pinned,
 * mapped slot -- no copy operation), the two kernels and the statistics merge are enqueued, and the read-back of the
 * [n_trials] mcl_stats is started.  Collect it with mcl_get_stats_end -- the host is free in between (the drop-in
 * ParticleFilter.main_pf simulates the next frame's observation there). */
int mcl_step_async(mcl_filter *h, const double *z_host, const double *ofs_host, const float *noise_dev, uint32_t flags,
                   void *stream);

/* Recompute mean/cov with the exact two-pass definition (wrap about the circular mean). */
int mcl_pose_moments(mcl_filter *h, void *stream);

/* ---- results ---------------------------------------------------------------------------------
 * Copies [n_trials] mcl_stats to host memory and synchronises the stream. */
int mcl_get_stats(mcl_filter *h, mcl_stats *out_host, void *stream);
/* The same in two halves, so the host can work while the step runs: mcl_get_stats_begin enqueues the pending
 * merge, the copy into a pinned buffer of the handle and an event, and returns at once; mcl_get_stats_end waits
 * for that event only and fills out_host (log_norm / max_weight / mean / cov derived on the host as above).
 * A single filter (n_trials == 1) with a merge pending skips the copy: the merge kernel writes the block into the
 * pinned, mapped buffer itself and releases a sequence number that mcl_get_stats_end polls (errors surface through
 * the event). */
int mcl_get_stats_begin(mcl_filter *h, void *stream);
int mcl_get_stats_end(mcl_filter *h, mcl_stats *out_host);
/* Device address of the [n_trials] mcl_stats block (for callers that poll asynchronously).  On the device
 * only the sums, ess, max_idx and est are current (see `derived`). */
int mcl_stats_dev(mcl_filter *h, const mcl_stats **out_dev);
/* Enqueue, on `stream`, whatever merge is still pending so that the device block above is current. */
int mcl_sync_stats(mcl_filter *h, void *stream);
/* Normalised weights exp(logw - log_norm) as fp64 [n_trials][n_particles] (pw of :194-196). */
int mcl_get_weights(mcl_filter *h, double *w_host, void *stream);
/* The integers the resampler scans: rint(w * 2^52) as uint64 [n_trials][n_particles]. */
int mcl_get_qweights(mcl_filter *h, uint64_t *q_host, void *stream);
/* Ancestor of every slot from the last resampling stage, [n_trials][n_particles] uint32. */
int mcl_get_ancestors(mcl_filter *h, uint32_t *anc_host, void *stream);

/* ---- measurement hooks (bench.py) -----------------------------------------------------------------
 * mcl_profile(h, 1): every following mcl_step* call brackets its two kernels (k_scan, k_step) with CUDA
 * events on the caller's stream.  mcl_profile_read synchronises on the last step and returns that step's
 * kernel durations in milliseconds: ms[0] = 0 (historic: a separate prep kernel), ms[1] k_scan, ms[2] k_step.
 * mcl_launch_count: kernels launched through this handle since creation. */
int mcl_profile(mcl_filter *h, int enable);
int mcl_profile_read(mcl_filter *h, double *ms3_host);
/* The same for the last mcl_shard_step_peer(..., MCL_STAGE_ALL): ms[0] totals pass + totals exchange, ms[1] per-step
 * scalars and expansion (side by side) + "done" exchange (both include the wait for the slowest peer), ms[2] = 0
 * (historic: the scalars used to run here), ms[3] k_step, ms[4] statistics exchange. */
int mcl_profile_read_peer(mcl_filter *h, double *ms5_host);
uint64_t mcl_launch_count(const mcl_filter *h);
/* Debug aid (needs a library built with -DMCL_TIMELINE, else MCL_ERR_STATE): the first call (stamps_host may be NULL) arms a per-CTA phase timeline of k_step; later calls
 * copy [n_ctas][12] global-timer nanosecond stamps of the most recent step (synchronises the device;
 * n_ctas = k_step CTAs followed by k_scan CTAs):
 * 0 start, 1 scalars loaded, 2 loop done, 3 warp sums flushed, 4 partial written, 5 ticket taken,
 * 6 statistics written, 7..10 stages of the merge (last CTA only). */
int mcl_debug_timeline(mcl_filter *h, uint64_t *stamps_host, int64_t capacity, int32_t *n_ctas);

/* ---- checkpoint ------------------------------------------------------------------
 * The state of a filter is its particles (mcl_get_particles, log-weights included), its [n_trials] mcl_stats
 * (mcl_get_stats: weight map, ESS gate, MAP pose = next reference pose) and the step counter (mcl_stats.step; the
 * Philox counter).  mcl_restore installs the last two after mcl_set_particles(..., MCL_SET_SKIP_STATS): the
 * restored filter continues bit-identically to the one that was saved. */
int mcl_restore(mcl_filter *h, const mcl_stats *stats_host, uint32_t step, void *stream);

/* ---- particle-sharded single filter (n_trials == 1, one handle per GPU) -----------------------
 * The host exchanges the small records; see slam-robot_simu_b200/sharded.py.  Order per step:
 *   mcl_predict_weight on every shard -> allgather of mcl_stats -> mcl_shard_merge (or a host merge +
 *   mcl_shard_set_global);
 *   when resampling: mcl_shard_total -> allgather of the totals -> mcl_shard_expand -> all-to-all
 *   of the off-rank poses -> mcl_shard_step (instead of mcl_predict_weight).
 * mcl_shard_set_global: install the merged statistics (global max_logw / sum_w define the weight
 * map, est the reference pose, ess the resampling gate). */
int mcl_shard_set_global(mcl_filter *h, const mcl_stats *merged_host, void *stream);
/* mcl_shard_merge: the same without the host -- records_dev holds the n_shards mcl_stats blocks (device memory,
 * as allgathered from mcl_stats_dev after mcl_sync_stats; shard g owns particles [g n_particles, (g+1) n_particles));
 * one small kernel merges them into this handle's statistics (device twin of sharded.merge_stats).  Asynchronous. */
int mcl_shard_merge(mcl_filter *h, const mcl_stats *records_dev, int n_shards, void *stream);
/* mcl_shard_total: sum of this shard's quantised weights under the global weight map -> *total_dev. */
int mcl_shard_total(mcl_filter *h, uint64_t *total_dev, void *stream);
/* mcl_shard_expand: with k_offset = sum of the totals of lower ranks, resample this shard's particles
 * with GLOBAL slot numbers.  Slots inside the shard's own range [slot_base, slot_base + n_particles) keep
 * the ancestor path (an index into the local state, consumed by mcl_shard_step); slots owned by other
 * ranks get their resampled pose written to send_{x,y,yaw}_dev, packed in slot order: first the slots
 * below the own range, then the slots above it -- exactly the layout a rank-ordered all-to-all sends.
 * `cap` = capacity of the send buffers in slots.  first_slot/n_slots (host; may both be NULL, then the
 * call does not synchronise) return the filled range [first, first + n). */
int mcl_shard_expand(mcl_filter *h, uint64_t k_offset, double ofs, float *send_x_dev, float *send_y_dev,
                     float *send_yaw_dev, int64_t cap, int64_t *first_slot, int64_t *n_slots, void *stream);
/* mcl_shard_step: predict + weight of the resampled shard.  Output slots [0, n_prefix) and
 * [n_particles - n_suffix, n_particles) were filled by other ranks: their poses are recv_*[0 .. n_prefix +
 * n_suffix) ([prefix | suffix], what the all-to-all delivers in rank order); the slots between gather
 * through the ancestors of mcl_shard_expand.  Weights restart from 1/NP (particle_filter.py:222). */
int mcl_shard_step(mcl_filter *h, const double *z_host, const float *noise_dev, const float *recv_x_dev,
                   const float *recv_y_dev, const float *recv_yaw_dev, int64_t n_prefix, int64_t n_suffix, uint32_t flags,
                   void *stream);

/* ---- particle-sharded single filter, DEVICE-DRIVEN (peer memory over NVLink / NVSwitch) -------------------------
 * The same step as the calls above with no host in the loop: statistics records, weight totals and the resampled
 * poses of off-rank slots are written by the producing kernels straight into the peers' memory, flags carry the
 * step number, slot ranges are derived on the device, and nothing synchronises with the host.  Semantics are those
 * of particle_filter.py:210-224 on the GLOBAL weights; results are bit-identical to the host-driven calls above.
 *
 *   mcl_shard_peer_alloc    allocates this shard's peer-visible region (mailbox + receive staging of n_particles
 *                           poses) with cudaMalloc and returns its CUDA IPC handle (64 bytes; may be NULL).
 *   mcl_shard_peer_connect  maps the peers' regions: `ipc_handles` = the n_shards handles in rank order (other
 *                           processes, one per GPU), or `local_peers` = the n_shards handles of THIS process
 *                           (several devices, or several shards on one device as the tests do).  Exactly one of
 *                           the two is non-NULL.
 *   mcl_shard_step_peer     main_pf minus the ground-truth line for this shard: [totals -> exchange -> expansion ->
 *                           exchange] when the step resamples (MCL_FORCE_RESAMPLE, or the ESS gate the previous
 *                           statistics exchange left on the device), predict + weight, statistics exchange.
 *                           ofs_host: one comb offset (rand() * NP_RECIP) or NULL = the Philox draw of the
 *                           unsharded filter.  `stages`: MCL_STAGE_ALL for a real step.  The one-GPU emulation
 *                           calls the four stages separately, each for all shards in turn, so that every "push"
 *                           precedes the matching "wait" on a single stream.
 *   mcl_shard_sync_peer     the statistics exchange alone (after mcl_set_particles), stages MCL_STAGE_C | _D.
 *   mcl_shard_get_plan      synchronises and copies the plan of the last step (diagnostics, error flag).
 * Statistics: mcl_get_stats / _begin / _end return the GLOBAL record after a step. */
#define MCL_IPC_HANDLE_BYTES 64
#define MCL_MAX_SHARDS 16
#define MCL_STAGE_A 1u /* integer weight totals; push them                                   */
#define MCL_STAGE_B 2u /* wait for the totals; plan; per-step scalars (needs z_host) beside the expansion; push "done" */
#define MCL_STAGE_C 4u /* wait for "done"; predict + weight (needs noise_dev or Philox); push the statistics record */
#define MCL_STAGE_D 8u /* wait for the records; merge; resampling gate of the next step      */
#define MCL_STAGE_ALL 15u

typedef struct mcl_shard_plan {
    uint64_t k_offset;   /* sum of the integer weight totals of all lower-ranked shards                 */
    double ofs;          /* comb offset of the step (particle_filter.py:214)                            */
    int64_t first_slot;  /* global slots [first_slot, first_slot + n_slots) are filled by this shard    */
    int64_t n_slots;
    uint64_t totals[MCL_MAX_SHARDS];
    uint32_t win_lo, win_hi; /* own output slots [win_lo, win_hi) gather locally, the rest came from peers */
    int32_t gate;        /* 1: the merged ESS is below the threshold (:211) -> the next step resamples  */
    int32_t resampled;   /* whether the step `epoch` resampled                                          */
    int32_t error;       /* 1: a wait on a peer timed out (sticky)                                      */
    uint32_t epoch;
} mcl_shard_plan;

int mcl_shard_peer_alloc(mcl_filter *h, int rank, int n_shards, unsigned char *ipc_handle_out);
int mcl_shard_peer_connect(mcl_filter *h, const unsigned char *ipc_handles, mcl_filter *const *local_peers);
int mcl_shard_step_peer(mcl_filter *h, const double *z_host, const double *ofs_host, const float *noise_dev,
                        uint32_t flags, uint32_t stages, void *stream);
int mcl_shard_sync_peer(mcl_filter *h, uint32_t stages, void *stream);
int mcl_shard_get_plan(mcl_filter *h, mcl_shard_plan *out_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MCL_B200_H */
