// C-ABI of the hot path (include/mbavo.h): context, level storage, evaluation orchestration.
// Host side of evaluate_cost_hessian_gradient (src/ba_tracker/spline_update_step.cpp:97-349) re-designed:
//   reference: 2 blocking H2D + 5 launches each followed by cudaDeviceSynchronize + 1 blocking D2H per evaluation
//   here:      pose kernel (the spline state travels as its launch parameter: no H2D copy) + fused tracking kernel whose
//              last block stores [cost, g, triu(H)] straight into mapped pinned host memory and publishes a sequence
//              number the host spins on: no D2H copy, no stream synchronisation, 2 launches per evaluation.
#include "../../include/mbavo.h"
#include "mbavo_device.h"

#include <cuda.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace mbavo
{
    cudaError_t launch_pose_kernel(int K, const EvalStage &stage, int total_samples, int with_jacobian, float *samples,
                                   double *mid, int *seg_end, cudaStream_t stream, GnState *gn, int knots_from, bool dependent,
                                   int buf_select, int samples_stride, int mid_stride, int seg_end_stride, double *dbg = nullptr);
    cudaError_t launch_track_debug_kernel(int K, int NK, const TrackParams &prm, dim3 grid, size_t smem, cudaStream_t stream);
    size_t track_cost_tma_smem_bytes(int K, int N, int S, int TP, int box_w, int box_h);
    cudaError_t launch_track_cost_tma(int K, const TrackParams &prm, const void *tmap, int box_w, int box_h, int grid, size_t smem,
                                      double *cost_out, unsigned long long *counters, cudaStream_t stream);
    cudaError_t launch_track_kernel(int K, int NK, bool with_j, bool big, const TrackParams &prm, dim3 grid, size_t smem,
                                    cudaStream_t stream, int *query_occupancy, bool dependent);
    size_t track_kernel_smem_bytes(int K, int NK, bool with_j, bool big, int N, int S, int TP);
    size_t sweep_kernel_smem_bytes(int K, int NK, int N, int S, int TP);
    cudaError_t launch_sweep_kernel(int K, int NK, const SweepParams &sp, const EvalStage &stage, int num_sms, size_t smem,
                                    cudaStream_t stream, bool dependent, int *query_occupancy);
    cudaError_t launch_pack_kernel(const unsigned char *I, const float *dIxy, int H, int W, PairTexel *pair, unsigned int *quad,
                                   int *inexact, cudaStream_t stream);
    cudaError_t launch_pyr_down_kernel(const unsigned char *src, int Ws, unsigned char *dst, int Hd, int Wd, cudaStream_t stream);
    cudaError_t launch_select_kernels(const SelectParams &prm, int total_cells, cudaStream_t stream);
    cudaError_t launch_pyr_step_kernel(PyrStepParams &p, cudaStream_t stream);
    cudaError_t launch_pack_image_kernel(const unsigned char *I, int H, int W, PairTexel *pair, unsigned int *quad, float *grad,
                                         cudaStream_t stream);
} // namespace mbavo

using namespace mbavo;

namespace
{
    thread_local char g_err[512] = "";

    int fail(int code, const char *fmt, ...)
    {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(g_err, sizeof g_err, fmt, ap);
        va_end(ap);
        return code;
    }

#define CUDA_TRY(expr)                                                                                      \
    do                                                                                                      \
    {                                                                                                       \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            return fail(MBAVO_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

    struct LevelStore
    {
        bool set = false;
        bool owns = false;      // buffers below were allocated by us (MBAVO_MEM_HOST)
        LevelDev dev{};         // what the kernel sees
        // owned buffers
        unsigned char *ref_I = nullptr;
        float *ref_dIxy = nullptr;
        unsigned char *cur_I[kMaxFrames] = {};
        double *xy = nullptr, *z = nullptr;
        size_t cap_pix = 0, cap_pts = 0; // capacities of the owned buffers (pixels per image, points)
        int cap_frames = 0;
        // keyframe texels (LevelDev::ref_pair / ref_quad), rebuilt by every mbavo_set_level
        PairTexel *tex_pair = nullptr;
        unsigned int *tex_quad = nullptr;
        size_t cap_tex = 0;
        // device-built pyramid path (mbavo_set_keyframe_pyramid / mbavo_set_live_pyramid / mbavo_set_level_points)
        unsigned char *pyr_ref = nullptr, *pyr_cur[kMaxFrames] = {};
        float *pyr_grad = nullptr;
        double *pts_xy = nullptr, *pts_z = nullptr;
        size_t pyr_cap_ref = 0, pyr_cap_cur = 0, pyr_cap_grad = 0, pts_cap = 0;
        int pyr_cap_frames = 0;
        bool has_key = false, has_live = false, has_pts = false;
        // always owned
        int2 *pattern = nullptr;
        int pattern_shadow[2 * 128] = {}; // host copy of what `pattern` holds (mbavo_set_frame / set_points: skip unchanged uploads)
        int pattern_shadow_n = 0;
        unsigned char *flags = nullptr;
        double *patch_cost = nullptr;
        int num_bad = 0;
        int last_eval_frames = 0;
    };

} // namespace

struct mbavo_ctx
{
    int device = 0;
    mbavo_limits lim{};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // second stream of mbavo_set_frame (point uploads under the pyramid build / the sweep)
    cudaEvent_t copy_done = nullptr, images_done = nullptr;
    // MBAVO_UPLOAD_ASYNC: the context's stream has NOT yet been made to wait for the point copies of the last mbavo_set_frame.
    // A persistent sweep does not need that wait — every Hessian pass spins on its level's ready flag, which the copy stream
    // sets right behind the level's points — everything else joins the streams first (join_uploads).
    bool join_pending = false;
    int upload_levels = 0;
    unsigned int upload_epoch = 0;
    unsigned int *ready_dev = nullptr;    // [MBAVO_MAX_LEVELS] epoch of the points resident in each level
    unsigned int *ready_src = nullptr;    // pinned: the value the flag copies carry
    int num_sms = 148;

    double cap[kMaxFrames] = {}, exp_time[kMaxFrames] = {};
    int n_frames_times = 0;

    LevelStore levels[MBAVO_MAX_LEVELS];

    EvalStage stage{};       // spline state of the evaluation being issued (launch parameter of the pose kernel)
    unsigned long long *phase_times_dev = nullptr; // development (mbavo_debug_phase_times)
    int trace_row = 0;
    GnState *gn_state = nullptr;                   // device-resident Gauss-Newton sweep (mbavo_gn_sweep)
    double *kf_dev = nullptr, *kf_host = nullptr;  // keyframe statistics: sums + poses (device / pinned)
    // semi-dense point selection (mbavo_select_points): cell records, level-0 depth map, counts (device / pinned)
    int4 *sel_cells = nullptr;
    float *sel_depth = nullptr;
    size_t sel_cells_cap = 0, sel_depth_cap = 0;
    int *sel_count_dev = nullptr, *sel_count_host = nullptr;
    long long device_sweeps = 0;                   // sweeps completed on the device-resident path
    long long persistent_sweeps = 0;               // ... of which inside ONE launch (sweep_kernel)
    bool use_persistent = true;                    // MBAVO_NO_PERSISTENT=1: one launch per pass
    SweepCtl *sweep_ctl = nullptr;                 // pass barrier of the persistent sweep kernel
    unsigned long long *sweep_pass_times = nullptr; // globaltimer stamps of the last persistent sweep (SweepParams::pass_times)
    int sweep_last_levels = 0;
    unsigned int sweep_base = 0;                   // value of sweep_ctl->done after the sweeps issued so far
    bool shard_shares_device = false;              // a peer rank lives on this GPU: persistent grids could not be co-resident
    bool use_device_sweep = true;                  // MBAVO_NO_DEVICE_SWEEP=1: every evaluation returns to the host
    long long big_block_batches = 2000; // MBAVO_BIG_BLOCK_BATCHES: Hessian pass uses the big block shape from this many batches
    bool use_pdl = true;     // MBAVO_NO_PDL=1: tracking kernel fully serialised behind the pose kernel
    bool small_level_split = true; // MBAVO_NO_SMALL_SPLIT=1: lane = pixel on every level, whatever its size

    // point sharding (mbavo_shard_*): this rank's mailbox, the mapped mailboxes of all ranks, exchange counters
    Mailbox *mailbox = nullptr;
    bool mailbox_fresh = false;             // zeroed by mbavo_shard_export and not yet consumed by a connect
    unsigned char device_uuid[16] = {};     // of this context's GPU (peers on the same GPU are recognised by it)
    ShardParams shard{};                    // world <= 1: not sharded
    bool peer_is_ipc[kMaxShards] = {};      // opened with cudaIpcOpenMemHandle (to be closed)
    unsigned long long shard_seq = 0, aux_seq = 0;
    int points_global[MBAVO_MAX_LEVELS] = {};
    float *samples = nullptr;
    int samples_stride = 0;  // floats from record buffer A to record buffer B
    double *mid = nullptr;
    int *seg_end = nullptr;
    double *block_partials = nullptr;
    size_t block_partials_cap = 0;
    unsigned int *counter = nullptr;
    double *packed_dev = nullptr;                          // E_max doubles, device
    double *result_host = nullptr, *result_map = nullptr;  // E_max (value, sequence) pairs: mapped pinned memory (host / device view)
    std::vector<double> result_vals;                       // the values of the last blocking evaluation, compacted
    int result_len = 0;
    unsigned long long seq = 0;
    int *outlier_result_dev = nullptr, *outlier_result_host = nullptr;

    bool use_texels = true;  // MBAVO_NO_TEXELS=1: always gather ref_I / ref_dIxy directly
    int force_phases = 0;    // MBAVO_PHASES=n: override the exposure-phase split (development / tests)
    int phase_fast = 0;      // MBAVO_PHASE_FAST=1: phase-fastest lane order
    int *inexact_dev = nullptr, *inexact_host = nullptr;

    struct OccEntry
    {
        int K, NK, with_h, packed;
        size_t smem;
        int occ;
    };
    std::vector<OccEntry> occ_cache;

    long long launches = 0;
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = -1.f;
};

namespace
{
    struct DeviceGuard
    {
        int prev = -1;
        explicit DeviceGuard(int dev)
        {
            cudaGetDevice(&prev);
            if (prev != dev)
                cudaSetDevice(dev);
            else
                prev = -1;
        }
        ~DeviceGuard()
        {
            if (prev >= 0)
                cudaSetDevice(prev);
        }
    };

    void free_level(LevelStore &L)
    {
        if (L.owns)
        {
            cudaFree(L.ref_I);
            cudaFree(L.ref_dIxy);
            for (auto &c : L.cur_I)
            {
                cudaFree(c);
                c = nullptr;
            }
            cudaFree(L.xy);
            cudaFree(L.z);
        }
        L.ref_I = nullptr, L.ref_dIxy = nullptr, L.xy = nullptr, L.z = nullptr;
        L.cap_pix = L.cap_pts = 0, L.cap_frames = 0;
        L.owns = false;
    }

    // detectOutliersAndUploadToGpu (blur_aware_direct_tracker.cpp:650-698) on the device: one block, fixed-order tree sums.
    // result[0] = number flagged by this call (over all ranks when sharded), result[1] = 1 if a peer timed out.
    // Sharded: mean, variance and the flag count are all-reduced through the mailboxes (3 small exchanges, seq0 + 0..2).
    __global__ void outlier_kernel(const double *__restrict__ cost, int P, int stride, double k_sigma,
                                   unsigned char *__restrict__ flags, int *__restrict__ result, const ShardParams sh)
    {
        __shared__ double s_a[1024];
        __shared__ double s_b[1024];
        __shared__ double s_x[8];
        const int tid = threadIdx.x;
        bool ok = true;
        double sum = 0, cnt = 0;
        for (int i = tid; i < P; i += blockDim.x)
        {
            const double c = cost[(size_t)i * stride];
            if (c >= 1e-8)
                sum += c, cnt += 1;
        }
        s_a[tid] = sum, s_b[tid] = cnt;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1)
        {
            if (tid < o)
                s_a[tid] += s_a[tid + o], s_b[tid] += s_b[tid + o];
            __syncthreads();
        }
        if (tid == 0)
            s_x[0] = s_a[0], s_x[1] = s_b[0];
        __syncthreads();
        if (sh.world > 1)
            ok = shard_allreduce_small(sh, sh.seq, s_x, 2) && ok;
        const double n = s_x[1], mu = s_x[0] / n;
        __syncthreads();
        double var = 0;
        for (int i = tid; i < P; i += blockDim.x)
        {
            const double c = cost[(size_t)i * stride];
            if (c >= 1e-8)
                var += (c - mu) * (c - mu);
        }
        s_a[tid] = var;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1)
        {
            if (tid < o)
                s_a[tid] += s_a[tid + o];
            __syncthreads();
        }
        if (tid == 0)
            s_x[0] = s_a[0];
        __syncthreads();
        if (sh.world > 1)
            ok = shard_allreduce_small(sh, sh.seq + 1, s_x, 1) && ok;
        const double thr = k_sigma * (double)sqrtf((float)(s_x[0] / n)); // `max_chi_square_error * sqrtf(var)`, :687
        __syncthreads();
        double bad = 0;
        for (int i = tid; i < P; i += blockDim.x)
        {
            const double c = cost[(size_t)i * stride];
            if (fabs(c - mu) > thr)
            {
                flags[i] = 1;
                bad += 1;
            }
        }
        s_a[tid] = bad;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1)
        {
            if (tid < o)
                s_a[tid] += s_a[tid + o];
            __syncthreads();
        }
        if (tid == 0)
            s_x[0] = s_a[0];
        __syncthreads();
        if (sh.world > 1)
            ok = shard_allreduce_small(sh, sh.seq + 2, s_x, 1) && ok;
        if (tid == 0)
        {
            result[0] = (int)s_x[0];
            result[1] = ok ? 0 : 1;
        }
    }

    // isKeyframe statistics (blur_aware_direct_tracker.cpp:205-248): every host-map point is un-projected with its depth,
    // moved into the live camera at three poses T_cur2ref (capture time, -/+ half the exposure) and projected (pinhole,
    // CameraPinhole.cpp:79-110 without distortion).  out[0] = sum |flow|^2, out[1] = sum |blur kernel|^2 (fixed-order tree
    // sums, all-reduced through the mailboxes when sharded), out[2] = 1 if a peer timed out.
    __global__ void __launch_bounds__(256) keyframe_kernel(const char *__restrict__ xy, int xy_stride, int xy_offset, const double *__restrict__ z, int P,
                                    double fx, double fy, double cx, double cy, const double *__restrict__ poses_tq,
                                    double *__restrict__ out, const ShardParams sh)
    {
        __shared__ double s_a[256];
        __shared__ double s_b[256];
        __shared__ double s_x[8];
        __shared__ double Rt[3][12]; // per pose: R (row-major) and t
        const int tid = threadIdx.x;
        if (tid < 3)
        {
            const double *p = poses_tq + 7 * tid, qx = p[3], qy = p[4], qz = p[5], qw = p[6];
            double *R = Rt[tid];
            R[0] = qw * qw + qx * qx - qy * qy - qz * qz, R[1] = 2 * (qx * qy - qw * qz), R[2] = 2 * (qx * qz + qw * qy);
            R[3] = 2 * (qx * qy + qw * qz), R[4] = qw * qw - qx * qx + qy * qy - qz * qz, R[5] = 2 * (qy * qz - qw * qx);
            R[6] = 2 * (qx * qz - qw * qy), R[7] = 2 * (qy * qz + qw * qx), R[8] = qw * qw - qx * qx - qy * qy + qz * qz;
            R[9] = p[0], R[10] = p[1], R[11] = p[2];
        }
        __syncthreads();
        double flow = 0, kern = 0;
        for (int i = tid; i < P; i += blockDim.x)
        {
            const double *pt = reinterpret_cast<const double *>(xy + (size_t)i * xy_stride + xy_offset);
            const double x = pt[0], y = pt[1], d = z[i];
            const double Pr[3] = {d * ((x - cx) / fx), d * ((y - cy) / fy), d};
            double u[3], v[3];
            for (int k = 0; k < 3; ++k)
            {
                const double *R = Rt[k];
                const double a = Pr[0] - R[9], b = Pr[1] - R[10], c = Pr[2] - R[11];
                const double X = R[0] * a + R[3] * b + R[6] * c, Y = R[1] * a + R[4] * b + R[7] * c, Z = R[2] * a + R[5] * b + R[8] * c; // R^T (P - t)
                u[k] = fx * X / Z + cx, v[k] = fy * Y / Z + cy;
            }
            flow += (u[0] - x) * (u[0] - x) + (v[0] - y) * (v[0] - y);
            kern += (u[1] - u[2]) * (u[1] - u[2]) + (v[1] - v[2]) * (v[1] - v[2]);
        }
        s_a[tid] = flow, s_b[tid] = kern;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1)
        {
            if (tid < o)
                s_a[tid] += s_a[tid + o], s_b[tid] += s_b[tid + o];
            __syncthreads();
        }
        if (tid == 0)
            s_x[0] = s_a[0], s_x[1] = s_b[0];
        __syncthreads();
        bool ok = true;
        if (sh.world > 1)
            ok = shard_allreduce_small(sh, sh.seq, s_x, 2);
        if (tid == 0)
            out[0] = s_x[0], out[1] = s_x[1], out[2] = ok ? 0.0 : 1.0;
    }

    // Sample time and segment of (frame, sample): compute_virtual_camera_poses.cu:33 + SplineFunctor.h:13-19.  The device
    // recomputes the same expression with explicit round-to-nearest ops; volatile keeps the host from contracting.
    int segment_of_sample(double cap, double expo, int i, int N, double t0, double dt)
    {
        volatile double a = expo * 0.5;
        volatile double b = (double)i * expo;
        volatile double c = b / ((double)(N - 1) + 1e-8);
        volatile double t = (cap - a) + c;
        volatile double s = (t - t0) / dt;
        return (int)s;
    }

    struct EvalPlan
    {
        int K, NK, kmin, N, F, P, S, TP, PH, batches_per_frame;
        bool with_h;
        bool big; // Hessian pass: one block of 20 (16) warps per SM instead of two of 8
        dim3 grid;
        size_t smem;
        int E;
    };

    int plan_evaluation(mbavo_ctx *ctx, int level, const mbavo_spline *sp, bool with_h, EvalPlan &pl)
    {
        if (!ctx || level < 0 || level >= MBAVO_MAX_LEVELS || !sp)
            return fail(MBAVO_EINVAL, "bad context / level / spline");
        LevelStore &L = ctx->levels[level];
        if (!L.set)
            return fail(MBAVO_ENOTREADY, "level %d has not been set", level);
        if (ctx->n_frames_times < L.dev.F)
            return fail(MBAVO_ENOTREADY, "frame times set for %d frames, level has %d", ctx->n_frames_times, L.dev.F);
        if (sp->spline_deg_k != 2 && sp->spline_deg_k != 4)
            return fail(MBAVO_EINVAL, "spline_deg_k must be 2 or 4 (got %d)", sp->spline_deg_k);
        if (sp->num_ctrl_knots < sp->spline_deg_k || sp->num_ctrl_knots > 16 || sp->num_ctrl_knots > ctx->lim.max_num_ctrl_knots)
            return fail(MBAVO_ECAPACITY, "num_ctrl_knots %d outside [k, min(16, max_num_ctrl_knots)]", sp->num_ctrl_knots);
        if (!(sp->sample_dt > 0))
            return fail(MBAVO_EINVAL, "sample_dt must be positive");
        pl.K = sp->spline_deg_k;
        pl.N = L.dev.N, pl.F = L.dev.F, pl.P = L.dev.P, pl.S = L.dev.S;
        pl.with_h = with_h;

        EvalStage *st = &ctx->stage;
        int seg[kMaxFrames * 64];
        int lo = 1 << 30, hi = -(1 << 30);
        for (int f = 0; f < pl.F; ++f)
            for (int i = 0; i < pl.N; ++i)
            {
                const int idx = segment_of_sample(ctx->cap[f], ctx->exp_time[f], i, pl.N, sp->start_time, sp->sample_dt);
                if (idx < 0 || idx + pl.K > sp->num_ctrl_knots)
                    return fail(MBAVO_ERANGE, "frame %d sample %d lies in segment %d, outside the %d control knots", f, i, idx,
                                sp->num_ctrl_knots);
                seg[f * pl.N + i] = idx;
                lo = idx < lo ? idx : lo;
                hi = idx > hi ? idx : hi;
            }
        pl.kmin = lo;
        pl.NK = hi - lo + pl.K;
        for (int e = 0; e < pl.F * pl.N; ++e)
            st->seg_off[e] = (unsigned char)(seg[e] - lo);
        if (pl.NK > MBAVO_MAX_KNOT_WINDOW || (pl.K == 2 && pl.NK > 6) || (pl.K == 4 && pl.NK > 7))
            return fail(MBAVO_ECAPACITY, "exposure windows touch %d control knots; at most %d are supported for k=%d", pl.NK,
                        pl.K == 2 ? 6 : 7, pl.K);
        pl.E = with_h ? packed_len(pl.NK) : 1;
        pl.TP = 32 / pl.S > 0 ? 32 / pl.S : 1;
        pl.batches_per_frame = (pl.P + pl.TP - 1) / pl.TP;
        // lanes of a warp = PH exposure phases x 32/PH pixel slots.  PH = 1 (lane = pixel, samples walked sequentially in
        // registers) measured fastest on every BASELINE config (profiles/r1c_variants.md); MBAVO_PHASES overrides it.
        pl.PH = 1;
        if (ctx->force_phases > 0)
        {
            pl.PH = ctx->force_phases;
            while (pl.PH > pl.N)
                pl.PH >>= 1;
        }
        else if (ctx->small_level_split && pl.S == 8)
        {
            // A level with fewer 4-point batches than the GPU has warp slots is pure latency: every warp walks all N samples of its
            // one batch while most slots idle.  Such levels (coarse pyramid levels; every level of a frame sharded over many GPUs)
            // give each warp fewer points and split the exposure samples over the lanes instead — 2 points x 8 pixels x 2
            // phases, or 1 point x 8 pixels x 4 phases — so that the same work spreads over 2x / 4x the warps at half / a quarter
            // of the per-warp latency.  (For levels that fill the GPU the unsplit form is faster: profiles/r1_history.md.)
            const long long slots = (long long)ctx->num_sms * track_warps(true, 3, true); // warp slots of the big block shape
            const long long pts = (long long)pl.P * pl.F;
            if (pts <= slots && pl.N >= 8)
                pl.TP = 1, pl.PH = 4;
            else if (pts <= 2 * slots && pl.N >= 4)
                pl.TP = 2, pl.PH = 2;
            pl.batches_per_frame = (pl.P + pl.TP - 1) / pl.TP;
        }
        // block shape of the Hessian pass: big blocks once the batches occupy most SMs (measured cross-over ~2000 batches)
        pl.big = with_h && (long long)pl.batches_per_frame * pl.F >= ctx->big_block_batches;
        pl.smem = track_kernel_smem_bytes(pl.K, pl.NK, with_h, pl.big, pl.N, pl.S, pl.TP);
        if (pl.big && pl.smem > 200 * 1024)
        {
            pl.big = false;
            pl.smem = track_kernel_smem_bytes(pl.K, pl.NK, with_h, false, pl.N, pl.S, pl.TP);
        }
        if (pl.smem > 200 * 1024)
            return fail(MBAVO_ECAPACITY, "shared memory need %zu B exceeds 200 KiB (N=%d, S=%d, window=%d)", pl.smem, pl.N, pl.S,
                        pl.NK);
        int occ = 0;
        const int packed = L.dev.ref_pair != nullptr ? 1 : 0;
        for (const auto &c : ctx->occ_cache)
            if (c.K == pl.K && c.NK == pl.NK && c.with_h == (with_h ? 1 : 0) && c.packed == packed + 2 * (pl.big ? 1 : 0) && c.smem == pl.smem)
                occ = c.occ;
        if (occ == 0)
        {
            TrackParams query{};
            query.lv = L.dev; // selects the texel / direct-gather instantiation
            cudaError_t e = launch_track_kernel(pl.K, pl.NK, with_h, pl.big, query, dim3(), pl.smem, nullptr, &occ, false);
            if (e != cudaSuccess)
                return fail(MBAVO_ECUDA, "occupancy query failed: %s", cudaGetErrorString(e));
            ctx->occ_cache.push_back({pl.K, pl.NK, with_h ? 1 : 0, packed + 2 * (pl.big ? 1 : 0), pl.smem, occ});
        }
        const int wpb = track_warps(with_h, pl.NK, pl.big);
        int want = (pl.batches_per_frame + wpb - 1) / wpb;
        int cap_blocks = ctx->num_sms * occ / pl.F;
        if (cap_blocks < 1)
            cap_blocks = 1;
        pl.grid = dim3(want < cap_blocks ? want : cap_blocks, pl.F, 1);

        // spline state (launch parameter of the pose kernel)
        std::memcpy(st->knots_t, sp->knots_t, sizeof(double) * 3 * sp->num_ctrl_knots);
        std::memcpy(st->knots_R, sp->knots_R, sizeof(double) * 4 * sp->num_ctrl_knots);
        std::memcpy(st->cap, ctx->cap, sizeof(double) * pl.F);
        std::memcpy(st->exp_time, ctx->exp_time, sizeof(double) * pl.F);
        st->t0 = sp->start_time, st->dt = sp->sample_dt;
        st->n_knots = sp->num_ctrl_knots, st->K = pl.K, st->N = pl.N, st->F = pl.F, st->kmin = pl.kmin, st->NK = pl.NK;
        return MBAVO_OK;
    }

    // pose kernel -> tracking kernel on ctx->stream.  blocking: the tracking kernel also publishes the result into the
    // context's mapped pinned buffer under a fresh sequence number (wait_result picks it up).
    // Extras of a launch that belongs to a device-resident Gauss-Newton sweep
    struct SweepLaunch
    {
        GnParams gn{};          // gn.state == nullptr: plain evaluation
        int knots_from = 0;     // pose kernel: 0 launch parameter, 1 state->cur, 2 state->cand
        bool first = false;     // first launch of the sweep: its pose kernel is not launched programmatically
        bool skip_pose = false; // the records this evaluation needs are already in a buffer (BufSelect below)
        bool pose_with_j = false; // candidate records computed with Jacobians: the next level may stand on them
        int buf_select = kBufA; // record buffer the pose kernel writes and the tracking kernel reads
    };

    // make the context's stream wait for the point copies of an asynchronous mbavo_set_frame (see mbavo_ctx::join_pending)
    int join_uploads(mbavo_ctx *ctx)
    {
        if (ctx->join_pending)
        {
            CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
            ctx->join_pending = false;
        }
        return MBAVO_OK;
    }

    // before buffers of the context are replaced: nothing may be in flight on either stream
    int quiesce(mbavo_ctx *ctx)
    {
        if (ctx->join_pending)
        {
            CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
            ctx->join_pending = false;
        }
        if (cudaStreamQuery(ctx->stream) != cudaSuccess)
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return MBAVO_OK;
    }

    // per-block partials of `blocks` blocks with E elements each
    int ensure_block_partials(mbavo_ctx *ctx, size_t need)
    {
        if (need > ctx->block_partials_cap)
        {
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->block_partials);
            ctx->block_partials = nullptr;
            CUDA_TRY(cudaMalloc(&ctx->block_partials, need * sizeof(double)));
            ctx->block_partials_cap = need;
        }
        return MBAVO_OK;
    }

    // the launch parameter of one pass (one tracking-kernel launch, or one pass of the persistent sweep kernel); consumes one
    // result sequence number when blocking and one exchange sequence number when sharded
    void fill_track_params(mbavo_ctx *ctx, int level, const EvalPlan &pl, double *packed_dev_out, bool blocking, double inv_num_residuals,
                           double huber_a, const SweepLaunch *sweep, TrackParams &prm)
    {
        LevelStore &L = ctx->levels[level];
        prm = TrackParams{};
        prm.lv = L.dev;
        prm.samples = ctx->samples;
        prm.mid = ctx->mid;
        prm.seg_end = ctx->seg_end;
        prm.buf_select = sweep ? sweep->buf_select : (int)kBufA;
        prm.samples_stride = ctx->samples_stride, prm.mid_stride = kMidDoubles * kMaxFrames, prm.seg_end_stride = kMaxSegments * kMaxFrames;
        prm.inv_num_residuals = inv_num_residuals;
        prm.huber_a = (float)huber_a;
        prm.TP = pl.TP;
        prm.PH = pl.PH;
        prm.phase_fast = ctx->phase_fast;
        prm.batches_per_frame = pl.batches_per_frame;
        prm.block_partials = ctx->block_partials;
        prm.counter = ctx->counter;
        prm.packed_out = packed_dev_out;
        prm.phase_times = ctx->phase_times_dev;
        prm.trace_row = ctx->trace_row++;
        if (sweep)
            prm.gn = sweep->gn;
        if (ctx->shard.world > 1)
        {
            prm.shard = ctx->shard;
            prm.shard.seq = ++ctx->shard_seq;
        }
        if (blocking)
        {
            prm.host_out = reinterpret_cast<double2 *>(ctx->result_map);
            prm.seq = ++ctx->seq;
            ctx->result_len = pl.E;
        }
    }

    int run_evaluation(mbavo_ctx *ctx, int level, const EvalPlan &pl, double *packed_dev_out, bool blocking, double inv_num_residuals,
                       double huber_a, const SweepLaunch *sweep = nullptr)
    {
        LevelStore &L = ctx->levels[level];
        cudaStream_t s = ctx->stream;
        int rc = join_uploads(ctx);
        if (rc != MBAVO_OK)
            return rc;
        rc = ensure_block_partials(ctx, (size_t)pl.grid.x * pl.grid.y * (pl.E + 1)); // (rows padded to an even pitch)
        if (rc != MBAVO_OK)
            return rc;
        L.last_eval_frames = pl.F;
        const int buf_select = sweep ? sweep->buf_select : (int)kBufA;
        if (!(sweep && sweep->skip_pose))
        {
            ctx->launches += 1;
            CUDA_TRY(launch_pose_kernel(pl.K, ctx->stage, pl.N * pl.F, (pl.with_h || (sweep && sweep->pose_with_j)) ? 1 : 0, ctx->samples,
                                        ctx->mid, ctx->seg_end, s, sweep ? sweep->gn.state : nullptr, sweep ? sweep->knots_from : 0,
                                        sweep && !sweep->first && ctx->use_pdl && !ctx->shard_shares_device, buf_select, ctx->samples_stride,
                                        kMidDoubles * kMaxFrames,
                                        kMaxSegments * kMaxFrames));
        }
        ctx->launches += 1;
        TrackParams prm;
        fill_track_params(ctx, level, pl, packed_dev_out, blocking, inv_num_residuals, huber_a, sweep, prm);
        if (ctx->timing)
            CUDA_TRY(cudaEventRecord(ctx->ev0, s));
        // event timing brackets the tracking kernel alone, so it is then launched fully serialised
        // (ranks that time-slice one GPU run fully serialised: with programmatic dependent launches a sweep whose passes wait for
        // a peer PROCESS on the same device was observed to read the candidate's records early — tests only; one GPU per rank otherwise)
        CUDA_TRY(launch_track_kernel(pl.K, pl.NK, pl.with_h, pl.big, prm, pl.grid, pl.smem, s, nullptr,
                                     ctx->use_pdl && !ctx->timing && !ctx->shard_shares_device));
        if (ctx->timing)
            CUDA_TRY(cudaEventRecord(ctx->ev1, s));
        return MBAVO_OK;
    }

    // Spin until both self-validating words of every element of the result carry the tag of ctx->seq (a microsecond after
    // the tracking kernel's last block stores them; see TrackParams::host_out), collecting the values into
    // ctx->result_vals.  The stream is polled now and then so that a failed launch cannot hang the caller.
    int wait_published(mbavo_ctx *ctx, int first, int count, unsigned long long seq, double *vals, const char *what)
    {
        const volatile unsigned long long *words = reinterpret_cast<const volatile unsigned long long *>(ctx->result_host);
        int done = 0; // elements [first, first + done) have been seen with the right tag
        for (unsigned long long spins = 1; done < count; ++spins)
        {
            while (done < count && read_published(words + 2 * (first + done), seq, vals ? vals + done : nullptr))
                ++done;
            if (done < count && (spins & 0xfff) == 0)
            {
                cudaError_t e = cudaStreamQuery(ctx->stream);
                if (e != cudaSuccess && e != cudaErrorNotReady)
                    return fail(MBAVO_ECUDA, "%s failed: %s", what, cudaGetErrorString(e));
                if (e == cudaSuccess)
                {
                    // the stream is idle: every store of the kernel is visible by now
                    std::atomic_thread_fence(std::memory_order_acquire);
                    for (; done < count; ++done)
                        if (!read_published(words + 2 * (first + done), seq, vals ? vals + done : nullptr))
                            return fail(MBAVO_ECUDA, "%s finished without publishing its result", what);
                }
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        return MBAVO_OK;
    }

    int wait_result(mbavo_ctx *ctx) { return wait_published(ctx, 0, ctx->result_len, ctx->seq, ctx->result_vals.data(), "evaluation"); }

    // The coarse-to-fine sweep as TWO launches: the pose kernel for the knots the sweep starts from, then sweep_kernel, which runs
    // the Hessian pass and the cost pass of every level inside one resident grid (mbavo_device.h, SweepCtl).  Returns 1 when there
    // is no instantiation for this window (the caller runs the sweep pass by pass), < 0 on errors.
    int run_persistent_sweep(mbavo_ctx *ctx, int level_coarse, int nlev, int chain, const mbavo_spline *sp, double radius, double huber_a,
                             unsigned long long *seq_of_level)
    {
        cudaStream_t s = ctx->stream;
        EvalPlan plans[2 * MBAVO_MAX_LEVELS];
        size_t smem = 0, partials = 0;
        for (int li = 0; li < nlev; ++li)
            for (int pass = 0; pass < 2; ++pass)
            {
                EvalPlan &pl = plans[2 * li + pass];
                int rc = plan_evaluation(ctx, level_coarse - li, sp, pass == 0, pl);
                if (rc != MBAVO_OK)
                    return rc;
                if (pl.K != plans[0].K || pl.NK != plans[0].NK || pl.kmin != plans[0].kmin || pl.N != plans[0].N || pl.F != 1)
                    return 1;
                const size_t need = sweep_kernel_smem_bytes(pl.K, plans[0].NK, pl.N, pl.S, pl.TP);
                smem = need > smem ? need : smem;
                partials = (size_t)ctx->num_sms * (pl.E + 1) > partials ? (size_t)ctx->num_sms * (pl.E + 1) : partials; // (rows padded to an even pitch)
            }
        if (smem > 200 * 1024)
            return 1;
        // is there an instantiation for this window, and does one block per SM fit?  (asked once per shape: with_h = 2 marks the entries
        // of the sweep kernel in the occupancy cache)
        int occ = 0;
        cudaError_t e = cudaSuccess;
        for (const auto &c : ctx->occ_cache)
            if (c.K == plans[0].K && c.NK == plans[0].NK && c.with_h == 2 && c.smem == smem)
                occ = c.occ;
        if (occ == 0)
        {
            e = launch_sweep_kernel(plans[0].K, plans[0].NK, SweepParams{}, ctx->stage, ctx->num_sms, smem, s, false, &occ);
            if (e == cudaSuccess && occ >= 1)
                ctx->occ_cache.push_back({plans[0].K, plans[0].NK, 2, 0, smem, occ});
        }
        if (e == cudaErrorNotSupported || (e == cudaSuccess && occ < 1))
        {
            cudaGetLastError();
            return 1;
        }
        if (e != cudaSuccess)
            return fail(MBAVO_ECUDA, "sweep kernel: %s", cudaGetErrorString(e));
        if (!ctx->sweep_ctl)
        {
            CUDA_TRY(cudaMalloc(&ctx->sweep_ctl, sizeof(SweepCtl)));
            CUDA_TRY(cudaMemset(ctx->sweep_ctl, 0, sizeof(SweepCtl)));
            CUDA_TRY(cudaMalloc(&ctx->sweep_pass_times, sizeof(unsigned long long) * (1 + 2 * kMaxSweepLevels)));
            CUDA_TRY(cudaMemset(ctx->sweep_pass_times, 0, sizeof(unsigned long long) * (1 + 2 * kMaxSweepLevels)));
            ctx->sweep_base = 0;
        }
        int rc = ensure_block_partials(ctx, partials);
        if (rc != MBAVO_OK)
            return rc;
        // records of the starting knots (with Jacobians) into buffer A; initialises the sweep state (cur knots, status, cur_buf = 0).
        // Launched before the passes' parameters are put together: the kernel runs while the host does that.
        ctx->launches += 2;
        CUDA_TRY(launch_pose_kernel(plans[0].K, ctx->stage, plans[0].N, 1, ctx->samples, ctx->mid, ctx->seg_end, s, ctx->gn_state, 0, false, kBufA,
                                    ctx->samples_stride, kMidDoubles * kMaxFrames, kMaxSegments * kMaxFrames));
        static thread_local SweepParams prm; // ~10 KB: kept off the stack of the caller's thread
        prm.n_levels = nlev, prm.ctl = ctx->sweep_ctl, prm.base = ctx->sweep_base, prm.pass_times = ctx->sweep_pass_times;
        ctx->sweep_last_levels = nlev;
        for (int li = 0; li < nlev; ++li)
        {
            const int level = level_coarse - li;
            LevelStore &L = ctx->levels[level];
            L.last_eval_frames = 1;
            for (int pass = 0; pass < 2; ++pass)
            {
                const EvalPlan &pl = plans[2 * li + pass];
                const long long pts = ctx->shard.world > 1 ? ctx->points_global[level] : pl.P;
                if (ctx->shard.world > 1 && pts < pl.P)
                    return fail(MBAVO_ENOTREADY, "mbavo_shard_set_global_points has not been called for level %d", level);
                const long long nres = (pts - L.num_bad) * pl.F * pl.S;
                if (nres <= 0)
                    return fail(MBAVO_EINVAL, "level %d: no residuals left (%lld points, %d flagged as outliers)", level, pts, L.num_bad);
                SweepLaunch sw;
                sw.gn.state = ctx->gn_state;
                sw.gn.mode = pass == 0 ? 1 : 2;
                sw.gn.n_knots = sp->num_ctrl_knots, sw.gn.kmin = pl.kmin;
                sw.gn.chain = chain, sw.gn.slot = li, sw.gn.last = (li == nlev - 1) ? 1 : 0;
                sw.gn.radius = radius;
                // record buffers: the knots the sweep stands on / the candidate; the roles swap on the device at every commit
                sw.buf_select = pass == 0 ? (li == 0 ? kBufA : kBufCur) : kBufCand;
                fill_track_params(ctx, level, pl, ctx->packed_dev, true, 1.0 / (double)nres, huber_a, &sw, prm.pass[2 * li + pass]);
                if (pass == 0 && ctx->join_pending && level < ctx->upload_levels)
                {
                    // the level's points may still be on their way: the pass waits for the flag the copy stream sets behind them
                    prm.pass[2 * li].ready_flag = ctx->ready_dev + level;
                    prm.pass[2 * li].ready_epoch = ctx->upload_epoch;
                }
                if (pass == 1)
                    seq_of_level[li] = ctx->seq;
            }
        }
        if (ctx->timing)
            CUDA_TRY(cudaEventRecord(ctx->ev0, s));
        e = launch_sweep_kernel(plans[0].K, plans[0].NK, prm, ctx->stage, ctx->num_sms, smem, s, ctx->use_pdl && !ctx->timing, nullptr);
        if (e != cudaSuccess)
            return fail(MBAVO_ECUDA, "sweep kernel launch: %s", cudaGetErrorString(e));
        if (ctx->timing)
            CUDA_TRY(cudaEventRecord(ctx->ev1, s));
        ctx->sweep_base += 2u * (unsigned int)nlev;
        // levels of the upload this sweep does not visit are not covered by its flag waits
        if (ctx->join_pending && !(level_coarse - nlev + 1 == 0 && level_coarse + 1 >= ctx->upload_levels))
            return join_uploads(ctx);
        return MBAVO_OK;
    }
} // namespace

extern "C"
{
    const char *mbavo_last_error(void) { return g_err; }
    int mbavo_version(void) { return MBAVO_VERSION; }
    int mbavo_packed_len(int knot_window) { return packed_len(knot_window); }

    int mbavo_create(const mbavo_limits *lim, mbavo_ctx **out)
    {
        if (!lim || !out)
            return fail(MBAVO_EINVAL, "null argument");
        if (lim->max_num_frames < 1 || lim->max_num_frames > MBAVO_MAX_FRAMES)
            return fail(MBAVO_ECAPACITY, "max_num_frames must be in [1, %d]", MBAVO_MAX_FRAMES);
        if (lim->max_num_virtual_poses_per_frame < 1 || lim->max_num_virtual_poses_per_frame > 64)
            return fail(MBAVO_ECAPACITY, "max_num_virtual_poses_per_frame must be in [1, 64]");
        if (lim->max_num_keypoints < 1 || lim->max_patch_size < 1 || lim->max_patch_size > 128)
            return fail(MBAVO_ECAPACITY, "max_num_keypoints >= 1 and max_patch_size in [1, 128] required");
        if (lim->max_num_ctrl_knots < 2 || lim->max_num_ctrl_knots > 16)
            return fail(MBAVO_ECAPACITY, "max_num_ctrl_knots must be in [2, 16]");
        int dev = lim->device;
        if (dev < 0)
            CUDA_TRY(cudaGetDevice(&dev));
        int count = 0;
        CUDA_TRY(cudaGetDeviceCount(&count));
        if (dev >= count)
            return fail(MBAVO_EINVAL, "device %d out of range (%d devices)", dev, count);
        DeviceGuard guard(dev);

        mbavo_ctx *ctx = new mbavo_ctx();
        // any failure below releases what has been allocated so far (mbavo_destroy copes with a half-built context)
        struct Guard
        {
            mbavo_ctx *c;
            ~Guard()
            {
                if (c)
                    mbavo_destroy(c);
            }
        } undo{ctx};
        ctx->device = dev;
        ctx->lim = *lim;
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
        ctx->num_sms = prop.multiProcessorCount;
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        ctx->stream = ctx->own_stream;
        const size_t nsamp = (size_t)lim->max_num_frames * lim->max_num_virtual_poses_per_frame;
        // two record buffers each (A, B): see BufSelect
        ctx->samples_stride = (int)(nsamp * sample_rec_floats(4));
        CUDA_TRY(cudaMalloc(&ctx->samples, 2 * nsamp * sample_rec_floats(4) * sizeof(float)));
        CUDA_TRY(cudaMalloc(&ctx->mid, 2 * sizeof(double) * kMidDoubles * kMaxFrames));
        CUDA_TRY(cudaMalloc(&ctx->seg_end, 2 * sizeof(int) * kMaxSegments * kMaxFrames));
        CUDA_TRY(cudaMalloc(&ctx->counter, sizeof(unsigned int)));
        CUDA_TRY(cudaMemset(ctx->counter, 0, sizeof(unsigned int)));
        const int emax = packed_len(MBAVO_MAX_KNOT_WINDOW);
        CUDA_TRY(cudaMalloc(&ctx->packed_dev, sizeof(double) * emax));
        CUDA_TRY(cudaHostAlloc(&ctx->result_host, sizeof(double) * 2 * emax, cudaHostAllocMapped));
        std::memset(ctx->result_host, 0, sizeof(double) * 2 * emax);
        ctx->result_vals.assign(emax, 0.0);
        CUDA_TRY(cudaHostGetDevicePointer(&ctx->result_map, ctx->result_host, 0));
        CUDA_TRY(cudaMalloc(&ctx->outlier_result_dev, sizeof(int) * 4));
        CUDA_TRY(cudaMallocHost(&ctx->outlier_result_host, sizeof(int) * 4));
        CUDA_TRY(cudaEventCreate(&ctx->ev0));
        CUDA_TRY(cudaEventCreate(&ctx->ev1));
        CUDA_TRY(cudaMalloc(&ctx->inexact_dev, sizeof(int)));
        CUDA_TRY(cudaMallocHost(&ctx->inexact_host, sizeof(int)));
        const char *g = getenv("MBAVO_NO_PDL");
        ctx->use_pdl = !(g && g[0] == '1');
        g = getenv("MBAVO_NO_DEVICE_SWEEP");
        ctx->use_device_sweep = !(g && g[0] == '1');
        g = getenv("MBAVO_NO_SMALL_SPLIT");
        ctx->small_level_split = !(g && g[0] == '1');
        g = getenv("MBAVO_NO_PERSISTENT");
        ctx->use_persistent = !(g && g[0] == '1');
        g = getenv("MBAVO_BIG_BLOCK_BATCHES");
        if (g)
            ctx->big_block_batches = atoll(g);
        g = getenv("MBAVO_NO_TEXELS");
        ctx->use_texels = !(g && g[0] == '1');
        g = getenv("MBAVO_PHASE_FAST");
        ctx->phase_fast = (g && g[0] == '1') ? 1 : 0;
        g = getenv("MBAVO_PHASES");
        if (g)
        {
            const int ph = atoi(g);
            if (ph == 1 || ph == 2 || ph == 4 || ph == 8 || ph == 16 || ph == 32)
                ctx->force_phases = ph;
        }
        undo.c = nullptr;
        *out = ctx;
        return MBAVO_OK;
    }

    int mbavo_destroy(mbavo_ctx *ctx)
    {
        if (!ctx)
            return MBAVO_OK;
        DeviceGuard guard(ctx->device);
        if (ctx->copy_stream)
            cudaStreamSynchronize(ctx->copy_stream); // an asynchronous mbavo_set_frame may still be copying into the buffers freed below
        if (ctx->stream)
            cudaStreamSynchronize(ctx->stream);
        for (auto &L : ctx->levels)
        {
            free_level(L);
            cudaFree(L.pattern);
            cudaFree(L.flags);
            cudaFree(L.patch_cost);
            cudaFree(L.tex_pair);
            cudaFree(L.tex_quad);
            cudaFree(L.pyr_ref);
            cudaFree(L.pyr_grad);
            for (auto &c : L.pyr_cur)
                cudaFree(c);
            cudaFree(L.pts_xy);
            cudaFree(L.pts_z);
        }
        cudaFree(ctx->inexact_dev);
        cudaFreeHost(ctx->inexact_host);
        mbavo_shard_disconnect(ctx);
        cudaFree(ctx->mailbox);
        cudaFree(ctx->phase_times_dev);
        cudaFree(ctx->gn_state);
        cudaFree(ctx->sweep_ctl);
        cudaFree(ctx->sweep_pass_times);
        cudaFree(ctx->kf_dev);
        cudaFreeHost(ctx->kf_host);
        cudaFree(ctx->sel_cells);
        cudaFree(ctx->sel_depth);
        cudaFree(ctx->sel_count_dev);
        cudaFreeHost(ctx->sel_count_host);
        cudaFree(ctx->samples);
        cudaFree(ctx->mid);
        cudaFree(ctx->seg_end);
        cudaFree(ctx->block_partials);
        cudaFree(ctx->counter);
        cudaFree(ctx->packed_dev);
        cudaFreeHost(ctx->result_host);
        cudaFree(ctx->outlier_result_dev);
        cudaFreeHost(ctx->outlier_result_host);
        if (ctx->ev0)
            cudaEventDestroy(ctx->ev0);
        if (ctx->ev1)
            cudaEventDestroy(ctx->ev1);
        if (ctx->own_stream)
            cudaStreamDestroy(ctx->own_stream);
        if (ctx->copy_stream)
            cudaStreamDestroy(ctx->copy_stream);
        if (ctx->copy_done)
            cudaEventDestroy(ctx->copy_done);
        if (ctx->images_done)
            cudaEventDestroy(ctx->images_done);
        cudaFree(ctx->ready_dev);
        cudaFreeHost(ctx->ready_src);
        cudaGetLastError();
        delete ctx;
        return MBAVO_OK;
    }

    int mbavo_set_stream(mbavo_ctx *ctx, void *stream)
    {
        if (!ctx)
            return fail(MBAVO_EINVAL, "null context");
        DeviceGuard guard(ctx->device);
        {
            const int rcq = quiesce(ctx); // pending uploads belong to the old stream's order
            if (rcq != MBAVO_OK)
                return rcq;
        }
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
        return MBAVO_OK;
    }

    int mbavo_set_frame_times(mbavo_ctx *ctx, int n_frames, const double *cap_time, const double *exp_time)
    {
        if (!ctx || !cap_time || !exp_time)
            return fail(MBAVO_EINVAL, "null argument");
        if (n_frames < 1 || n_frames > ctx->lim.max_num_frames)
            return fail(MBAVO_ECAPACITY, "n_frames %d outside [1, %d]", n_frames, ctx->lim.max_num_frames);
        for (int f = 0; f < n_frames; ++f)
            ctx->cap[f] = cap_time[f], ctx->exp_time[f] = exp_time[f];
        ctx->n_frames_times = n_frames;
        return MBAVO_OK;
    }

    int mbavo_set_level(mbavo_ctx *ctx, int level, const mbavo_level *d)
    {
        if (!ctx || !d || level < 0 || level >= MBAVO_MAX_LEVELS)
            return fail(MBAVO_EINVAL, "bad context / level");
        if (d->mem != MBAVO_MEM_HOST && d->mem != MBAVO_MEM_DEVICE)
            return fail(MBAVO_EINVAL, "mem must be MBAVO_MEM_HOST or MBAVO_MEM_DEVICE");
        if (d->H < 2 || d->W < 2 || !d->ref_I || !d->ref_dIxy || !d->cur_I || !d->keypoint_xy || !d->keypoint_z || !d->pattern_xy)
            return fail(MBAVO_EINVAL, "level needs H,W >= 2 and non-null image / keypoint / pattern pointers");
        if (d->n_frames < 1 || d->n_frames > ctx->lim.max_num_frames)
            return fail(MBAVO_ECAPACITY, "n_frames %d outside [1, %d]", d->n_frames, ctx->lim.max_num_frames);
        if (d->num_keypoints < 1 || d->num_keypoints > ctx->lim.max_num_keypoints)
            return fail(MBAVO_ECAPACITY, "num_keypoints %d outside [1, %d]", d->num_keypoints, ctx->lim.max_num_keypoints);
        if (d->patch_size < 1 || d->patch_size > ctx->lim.max_patch_size)
            return fail(MBAVO_ECAPACITY, "patch_size %d outside [1, %d]", d->patch_size, ctx->lim.max_patch_size);
        if (d->num_virtual_poses < 1 || d->num_virtual_poses > ctx->lim.max_num_virtual_poses_per_frame)
            return fail(MBAVO_ECAPACITY, "num_virtual_poses %d outside [1, %d]", d->num_virtual_poses,
                        ctx->lim.max_num_virtual_poses_per_frame);
        if (d->keypoint_xy_stride < 16 || d->keypoint_xy_offset < 0 || d->keypoint_xy_offset + 16 > d->keypoint_xy_stride ||
            d->keypoint_xy_stride % 8 != 0 || d->keypoint_xy_offset % 8 != 0)
            return fail(MBAVO_EINVAL, "keypoint_xy stride/offset must describe two aligned doubles per record");
        if (!(d->fx > 0) || !(d->fy > 0))
            return fail(MBAVO_EINVAL, "fx, fy must be positive");
        DeviceGuard guard(ctx->device);
        LevelStore &L = ctx->levels[level];
        cudaStream_t s = ctx->stream;
        // nothing in flight may still read the buffers we are about to replace (only mbavo_evaluate_async leaves work in
        // flight; every other entry point returns with the stream idle)
        {
            const int rcq = quiesce(ctx);
            if (rcq != MBAVO_OK)
                return rcq;
        }
        const size_t npix = (size_t)d->H * d->W;
        const int P = d->num_keypoints, F = d->n_frames;

        if (!L.pattern)
        {
            CUDA_TRY(cudaMalloc(&L.pattern, sizeof(int2) * ctx->lim.max_patch_size));
            CUDA_TRY(cudaMalloc(&L.flags, ctx->lim.max_num_keypoints));
            CUDA_TRY(cudaMalloc(&L.patch_cost, sizeof(double) * (size_t)ctx->lim.max_num_keypoints * ctx->lim.max_num_frames));
        }
        if (d->mem == MBAVO_MEM_HOST)
        {
            if (!L.owns || L.cap_pix < npix || L.cap_pts < (size_t)P || L.cap_frames < F)
            {
                free_level(L);
                CUDA_TRY(cudaMalloc(&L.ref_I, npix));
                CUDA_TRY(cudaMalloc(&L.ref_dIxy, npix * 2 * sizeof(float)));
                for (int f = 0; f < F; ++f)
                    CUDA_TRY(cudaMalloc(&L.cur_I[f], npix));
                CUDA_TRY(cudaMalloc(&L.xy, sizeof(double) * 2 * P));
                CUDA_TRY(cudaMalloc(&L.z, sizeof(double) * P));
                L.owns = true, L.cap_pix = npix, L.cap_pts = P, L.cap_frames = F;
            }
            CUDA_TRY(cudaMemcpyAsync(L.ref_I, d->ref_I, npix, cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(L.ref_dIxy, d->ref_dIxy, npix * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
            for (int f = 0; f < F; ++f)
                CUDA_TRY(cudaMemcpyAsync(L.cur_I[f], d->cur_I[f], npix, cudaMemcpyHostToDevice, s));
            // compact the records to packed double2 on the way in
            CUDA_TRY(cudaMemcpy2DAsync(L.xy, 16, (const char *)d->keypoint_xy + d->keypoint_xy_offset, d->keypoint_xy_stride, 16, P,
                                       cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(L.z, d->keypoint_z, sizeof(double) * P, cudaMemcpyHostToDevice, s));
            L.dev.ref_I = L.ref_I, L.dev.ref_dIxy = reinterpret_cast<const float2 *>(L.ref_dIxy);
            for (int f = 0; f < F; ++f)
                L.dev.cur_I[f] = L.cur_I[f];
            L.dev.xy = reinterpret_cast<const char *>(L.xy), L.dev.xy_stride = 16, L.dev.xy_offset = 0;
            L.dev.z = L.z;
        }
        else
        {
            if (L.owns)
                free_level(L);
            L.dev.ref_I = d->ref_I, L.dev.ref_dIxy = reinterpret_cast<const float2 *>(d->ref_dIxy);
            for (int f = 0; f < F; ++f)
                L.dev.cur_I[f] = d->cur_I[f];
            L.dev.xy = reinterpret_cast<const char *>(d->keypoint_xy);
            L.dev.xy_stride = d->keypoint_xy_stride, L.dev.xy_offset = d->keypoint_xy_offset;
            L.dev.z = d->keypoint_z;
        }
        for (int f = F; f < kMaxFrames; ++f)
            L.dev.cur_I[f] = nullptr;
        // keyframe texels: built from the device copies on every call (the caller may have changed the image content
        // behind an unchanged pointer); kept only if the texels reproduce every gradient value exactly (LevelDev)
        L.dev.ref_pair = nullptr, L.dev.ref_quad = nullptr;
        bool texels_pending = false;
        if (ctx->use_texels && npix < (size_t)1 << 27)
        {
            if (L.cap_tex < npix)
            {
                cudaFree(L.tex_pair);
                cudaFree(L.tex_quad);
                L.tex_pair = nullptr, L.tex_quad = nullptr, L.cap_tex = 0;
                CUDA_TRY(cudaMalloc(&L.tex_pair, npix * sizeof(PairTexel)));
                CUDA_TRY(cudaMalloc(&L.tex_quad, npix * sizeof(unsigned int)));
                L.cap_tex = npix;
            }
            CUDA_TRY(cudaMemsetAsync(ctx->inexact_dev, 0, sizeof(int), s));
            CUDA_TRY(launch_pack_kernel(L.dev.ref_I, reinterpret_cast<const float *>(L.dev.ref_dIxy), d->H, d->W, L.tex_pair,
                                        L.tex_quad, ctx->inexact_dev, s));
            ctx->launches += 1;
            CUDA_TRY(cudaMemcpyAsync(ctx->inexact_host, ctx->inexact_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
            texels_pending = true; // the verdict is read after the one synchronisation at the end
        }
        CUDA_TRY(cudaMemcpyAsync(L.pattern, d->pattern_xy, sizeof(int2) * d->patch_size, cudaMemcpyDefault, s));
        L.pattern_shadow_n = 0;
        if (!d->ext_outlier_flags)
        {
            CUDA_TRY(cudaMemsetAsync(L.flags, 0, ctx->lim.max_num_keypoints, s)); // blur_aware_direct_tracker.cpp:600-601
            L.num_bad = 0;
        }
        L.dev.H = d->H, L.dev.W = d->W;
        L.dev.fx = d->fx, L.dev.fy = d->fy, L.dev.cx = d->cx, L.dev.cy = d->cy;
        L.dev.inv_fx = 1.0 / d->fx, L.dev.inv_fy = 1.0 / d->fy;
        L.dev.P = P, L.dev.S = d->patch_size, L.dev.N = d->num_virtual_poses, L.dev.F = F;
        L.dev.pattern = L.pattern;
        L.dev.flags = d->ext_outlier_flags ? d->ext_outlier_flags : L.flags;
        if (d->ext_patch_cost)
        {
            if (d->ext_patch_cost_stride < 1)
                return fail(MBAVO_EINVAL, "ext_patch_cost_stride must be >= 1");
            L.dev.patch_cost = d->ext_patch_cost, L.dev.patch_cost_stride = d->ext_patch_cost_stride;
        }
        else
            L.dev.patch_cost = L.patch_cost, L.dev.patch_cost_stride = 1;
        L.set = true;
        L.has_key = L.has_live = L.has_pts = false; // the level no longer comes from the device-built pyramid path
        CUDA_TRY(cudaStreamSynchronize(s)); // host buffers are only borrowed for the call
        if (texels_pending && *ctx->inexact_host == 0)
            L.dev.ref_pair = L.tex_pair, L.dev.ref_quad = L.tex_quad;
        return MBAVO_OK;
    }

    // ---- device-built pyramids ----------------------------------------------------------------
    static int ensure_level_scratch(mbavo_ctx *ctx, LevelStore &L)
    {
        if (!L.pattern)
        {
            CUDA_TRY(cudaMalloc(&L.pattern, sizeof(int2) * ctx->lim.max_patch_size));
            CUDA_TRY(cudaMalloc(&L.flags, ctx->lim.max_num_keypoints));
            CUDA_TRY(cudaMalloc(&L.patch_cost, sizeof(double) * (size_t)ctx->lim.max_num_keypoints * ctx->lim.max_num_frames));
        }
        return MBAVO_OK;
    }

    static void pyramid_level_ready(LevelStore &L)
    {
        L.set = L.has_key && L.has_live && L.has_pts;
    }

    // validation, allocation and the enqueue of the upload + pyramid / gradient / texel kernels on `s` (no synchronisation)
    // what: bit 0 = allocate + enqueue the level-0 copy on s_copy, bit 1 = enqueue the kernels on s (mbavo_set_frame issues the copies
    // of both images first, on its copy stream, and the kernels afterwards)
    static int enqueue_keyframe_pyramid(mbavo_ctx *ctx, int n_levels, int mem, const unsigned char *ref_I0, int H0, int W0, cudaStream_t s,
                                        int what = 3, cudaStream_t s_copy = nullptr)
    {
        if (!s_copy)
            s_copy = s;
        if (!ctx || !ref_I0 || n_levels < 1 || n_levels > MBAVO_MAX_LEVELS || (mem != MBAVO_MEM_HOST && mem != MBAVO_MEM_DEVICE))
            return fail(MBAVO_EINVAL, "bad arguments");
        if ((H0 >> (n_levels - 1)) < 2 || (W0 >> (n_levels - 1)) < 2)
            return fail(MBAVO_EINVAL, "%d levels of a %d x %d image leave less than 2 x 2 pixels", n_levels, H0, W0);
        for (int l = 0; l < n_levels; ++l)
        {
            LevelStore &L = ctx->levels[l];
            const int H = H0 / (1 << l), W = W0 / (1 << l); // ImagePyramid.h:71-72
            const size_t npix = (size_t)H * W;
            int rc = ensure_level_scratch(ctx, L);
            if (rc != MBAVO_OK)
                return rc;
            if (L.has_key && (L.dev.H != H || L.dev.W != W))
                L.has_live = false; // live images of another size are void
            if (L.pyr_cap_ref < npix)
            {
                cudaFree(L.pyr_ref);
                L.pyr_ref = nullptr, L.pyr_cap_ref = 0;
                CUDA_TRY(cudaMalloc(&L.pyr_ref, npix));
                L.pyr_cap_ref = npix;
            }
            if (ctx->use_texels && L.cap_tex < npix)
            {
                cudaFree(L.tex_pair);
                cudaFree(L.tex_quad);
                L.tex_pair = nullptr, L.tex_quad = nullptr, L.cap_tex = 0;
                CUDA_TRY(cudaMalloc(&L.tex_pair, npix * sizeof(PairTexel)));
                CUDA_TRY(cudaMalloc(&L.tex_quad, npix * sizeof(unsigned int)));
                L.cap_tex = npix;
            }
            if (!ctx->use_texels && L.pyr_cap_grad < npix)
            {
                cudaFree(L.pyr_grad);
                L.pyr_grad = nullptr, L.pyr_cap_grad = 0;
                CUDA_TRY(cudaMalloc(&L.pyr_grad, npix * 2 * sizeof(float)));
                L.pyr_cap_grad = npix;
            }
            if (l == 0 && (what & 1))
                CUDA_TRY(cudaMemcpyAsync(L.pyr_ref, ref_I0, npix, mem == MBAVO_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s_copy));
            if (what & 2)
            {
                if (l > 0)
                    CUDA_TRY(launch_pyr_down_kernel(ctx->levels[l - 1].pyr_ref, W0 / (1 << (l - 1)), L.pyr_ref, H, W, s));
                CUDA_TRY(launch_pack_image_kernel(L.pyr_ref, H, W, ctx->use_texels ? L.tex_pair : nullptr, ctx->use_texels ? L.tex_quad : nullptr,
                                                  ctx->use_texels ? nullptr : L.pyr_grad, s));
                ctx->launches += l == 0 ? 1 : 2;
            }
            if (L.owns)
                free_level(L); // buffers of an earlier mbavo_set_level
            L.dev.H = H, L.dev.W = W;
            L.dev.ref_I = L.pyr_ref;
            L.dev.ref_dIxy = reinterpret_cast<const float2 *>(ctx->use_texels ? nullptr : L.pyr_grad);
            L.dev.ref_pair = ctx->use_texels ? L.tex_pair : nullptr;
            L.dev.ref_quad = ctx->use_texels ? L.tex_quad : nullptr;
            L.has_key = true;
            pyramid_level_ready(L);
        }
        return MBAVO_OK;
    }

    int mbavo_set_keyframe_pyramid(mbavo_ctx *ctx, int n_levels, int mem, const unsigned char *ref_I0, int H0, int W0)
    {
        if (!ctx)
            return fail(MBAVO_EINVAL, "null context");
        DeviceGuard guard(ctx->device);
        cudaStream_t s = ctx->stream;
        {
            const int rcq = quiesce(ctx);
            if (rcq != MBAVO_OK)
                return rcq;
        }
        int rc = enqueue_keyframe_pyramid(ctx, n_levels, mem, ref_I0, H0, W0, s);
        cudaError_t e = cudaStreamSynchronize(s); // the host image is only borrowed for the call
        if (rc == MBAVO_OK && e != cudaSuccess)
            return fail(MBAVO_ECUDA, "upload failed: %s", cudaGetErrorString(e));
        return rc;
    }

    static int enqueue_live_pyramid(mbavo_ctx *ctx, int n_levels, int mem, const unsigned char *const *cur_I0, int n_frames, cudaStream_t s,
                                    int what = 3, cudaStream_t s_copy = nullptr)
    {
        if (!s_copy)
            s_copy = s;
        if (!ctx || !cur_I0 || n_levels < 1 || n_levels > MBAVO_MAX_LEVELS || (mem != MBAVO_MEM_HOST && mem != MBAVO_MEM_DEVICE))
            return fail(MBAVO_EINVAL, "bad arguments");
        if (n_frames < 1 || n_frames > ctx->lim.max_num_frames)
            return fail(MBAVO_ECAPACITY, "n_frames %d outside [1, %d]", n_frames, ctx->lim.max_num_frames);
        for (int l = 0; l < n_levels; ++l)
            if (!ctx->levels[l].has_key)
                return fail(MBAVO_ENOTREADY, "mbavo_set_keyframe_pyramid has not set level %d", l);
        for (int l = 0; l < n_levels; ++l)
        {
            LevelStore &L = ctx->levels[l];
            const size_t npix = (size_t)L.dev.H * L.dev.W;
            if (L.pyr_cap_cur < npix || L.pyr_cap_frames < n_frames)
            {
                for (auto &c : L.pyr_cur)
                {
                    cudaFree(c);
                    c = nullptr;
                }
                L.pyr_cap_cur = 0, L.pyr_cap_frames = 0;
                for (int f = 0; f < n_frames; ++f)
                    CUDA_TRY(cudaMalloc(&L.pyr_cur[f], npix));
                L.pyr_cap_cur = npix, L.pyr_cap_frames = n_frames;
            }
            for (int f = 0; f < n_frames; ++f)
            {
                if (l == 0)
                {
                    if (!cur_I0[f])
                        return fail(MBAVO_EINVAL, "null image");
                    if (what & 1)
                        CUDA_TRY(cudaMemcpyAsync(L.pyr_cur[f], cur_I0[f], npix, mem == MBAVO_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s_copy));
                }
                else if (what & 2)
                {
                    CUDA_TRY(launch_pyr_down_kernel(ctx->levels[l - 1].pyr_cur[f], ctx->levels[l - 1].dev.W, L.pyr_cur[f], L.dev.H, L.dev.W, s));
                    ctx->launches += 1;
                }
                L.dev.cur_I[f] = L.pyr_cur[f];
            }
            for (int f = n_frames; f < kMaxFrames; ++f)
                L.dev.cur_I[f] = nullptr;
            L.dev.F = n_frames;
            if (L.flags && (what & 2))
                CUDA_TRY(cudaMemsetAsync(L.flags, 0, ctx->lim.max_num_keypoints, s)); // tracker.cpp:600-601
            L.num_bad = 0;
            L.has_live = true;
            pyramid_level_ready(L);
        }
        return MBAVO_OK;
    }

    int mbavo_set_live_pyramid(mbavo_ctx *ctx, int n_levels, int mem, const unsigned char *const *cur_I0, int n_frames)
    {
        if (!ctx)
            return fail(MBAVO_EINVAL, "null context");
        DeviceGuard guard(ctx->device);
        cudaStream_t s = ctx->stream;
        {
            const int rcq = quiesce(ctx);
            if (rcq != MBAVO_OK)
                return rcq;
        }
        int rc = enqueue_live_pyramid(ctx, n_levels, mem, cur_I0, n_frames, s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (rc == MBAVO_OK && e != cudaSuccess)
            return fail(MBAVO_ECUDA, "upload failed: %s", cudaGetErrorString(e));
        return rc;
    }


    // enqueue the uploads of one level's points on the context's stream (no synchronisation)
    // s: stream of the copies; s_flags: stream of the outlier-flag memset.  They differ in mbavo_set_frame: a memset is a KERNEL, and
    // a kernel on the copy stream could not start while a persistent sweep kernel that waits for this very upload holds every SM.
    static int enqueue_level_points(mbavo_ctx *ctx, int level, const mbavo_level_points *d, cudaStream_t s, cudaStream_t s_flags = nullptr,
                                    bool clear_flags = true)
    {
        if (!s_flags)
            s_flags = s;
        if (!ctx || !d || level < 0 || level >= MBAVO_MAX_LEVELS)
            return fail(MBAVO_EINVAL, "bad context / level");
        if (d->mem != MBAVO_MEM_HOST && d->mem != MBAVO_MEM_DEVICE)
            return fail(MBAVO_EINVAL, "mem must be MBAVO_MEM_HOST or MBAVO_MEM_DEVICE");
        if (!d->keypoint_xy || !d->keypoint_z || !d->pattern_xy)
            return fail(MBAVO_EINVAL, "null keypoint / pattern pointers");
        if (d->num_keypoints < 1 || d->num_keypoints > ctx->lim.max_num_keypoints)
            return fail(MBAVO_ECAPACITY, "num_keypoints %d outside [1, %d]", d->num_keypoints, ctx->lim.max_num_keypoints);
        if (d->patch_size < 1 || d->patch_size > ctx->lim.max_patch_size)
            return fail(MBAVO_ECAPACITY, "patch_size %d outside [1, %d]", d->patch_size, ctx->lim.max_patch_size);
        if (d->num_virtual_poses < 1 || d->num_virtual_poses > ctx->lim.max_num_virtual_poses_per_frame)
            return fail(MBAVO_ECAPACITY, "num_virtual_poses %d outside [1, %d]", d->num_virtual_poses,
                        ctx->lim.max_num_virtual_poses_per_frame);
        if (d->keypoint_xy_stride < 16 || d->keypoint_xy_offset < 0 || d->keypoint_xy_offset + 16 > d->keypoint_xy_stride ||
            d->keypoint_xy_stride % 8 != 0 || d->keypoint_xy_offset % 8 != 0)
            return fail(MBAVO_EINVAL, "keypoint_xy stride/offset must describe two aligned doubles per record");
        if (!(d->fx > 0) || !(d->fy > 0))
            return fail(MBAVO_EINVAL, "fx, fy must be positive");
        LevelStore &L = ctx->levels[level];
        if (L.set && !L.has_key)
            return fail(MBAVO_EINVAL, "level %d was set by mbavo_set_level; points of such a level are replaced by mbavo_set_level", level);
        int rc = ensure_level_scratch(ctx, L);
        if (rc != MBAVO_OK)
            return rc;
        const int P = d->num_keypoints;
        if (d->mem == MBAVO_MEM_HOST)
        {
            if (L.pts_cap < (size_t)P)
            {
                cudaFree(L.pts_xy);
                cudaFree(L.pts_z);
                L.pts_xy = L.pts_z = nullptr, L.pts_cap = 0;
                CUDA_TRY(cudaMalloc(&L.pts_xy, sizeof(double) * 2 * P));
                CUDA_TRY(cudaMalloc(&L.pts_z, sizeof(double) * P));
                L.pts_cap = P;
            }
            if (d->keypoint_xy_stride == 16 && d->keypoint_xy_offset == 0)
                CUDA_TRY(cudaMemcpyAsync(L.pts_xy, d->keypoint_xy, sizeof(double) * 2 * P, cudaMemcpyHostToDevice, s));
            else // compact the records to packed double2 on the way in
                CUDA_TRY(cudaMemcpy2DAsync(L.pts_xy, 16, (const char *)d->keypoint_xy + d->keypoint_xy_offset, d->keypoint_xy_stride, 16, P,
                                           cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(L.pts_z, d->keypoint_z, sizeof(double) * P, cudaMemcpyHostToDevice, s));
            L.dev.xy = reinterpret_cast<const char *>(L.pts_xy), L.dev.xy_stride = 16, L.dev.xy_offset = 0;
            L.dev.z = L.pts_z;
        }
        else
        {
            L.dev.xy = reinterpret_cast<const char *>(d->keypoint_xy);
            L.dev.xy_stride = d->keypoint_xy_stride, L.dev.xy_offset = d->keypoint_xy_offset;
            L.dev.z = d->keypoint_z;
        }
        // the pattern is a few dozen bytes that rarely change: a host-side shadow of what the device holds saves the copy
        if (d->mem != MBAVO_MEM_HOST || L.pattern_shadow_n != d->patch_size ||
            std::memcmp(L.pattern_shadow, d->pattern_xy, sizeof(int2) * d->patch_size) != 0)
        {
            CUDA_TRY(cudaMemcpyAsync(L.pattern, d->pattern_xy, sizeof(int2) * d->patch_size, cudaMemcpyDefault, s));
            L.pattern_shadow_n = 0;
            if (d->mem == MBAVO_MEM_HOST && d->patch_size <= 128)
            {
                std::memcpy(L.pattern_shadow, d->pattern_xy, sizeof(int2) * d->patch_size);
                L.pattern_shadow_n = d->patch_size;
            }
        }
        if (clear_flags)
            CUDA_TRY(cudaMemsetAsync(L.flags, 0, P, s_flags));
        L.num_bad = 0;
        L.dev.fx = d->fx, L.dev.fy = d->fy, L.dev.cx = d->cx, L.dev.cy = d->cy;
        L.dev.inv_fx = 1.0 / d->fx, L.dev.inv_fy = 1.0 / d->fy;
        L.dev.P = P, L.dev.S = d->patch_size, L.dev.N = d->num_virtual_poses;
        L.dev.pattern = L.pattern;
        L.dev.flags = L.flags;
        L.dev.patch_cost = L.patch_cost, L.dev.patch_cost_stride = 1;
        L.has_pts = true;
        pyramid_level_ready(L);
        return MBAVO_OK;
    }


    int mbavo_set_level_points(mbavo_ctx *ctx, int level, const mbavo_level_points *d)
    {
        if (!ctx)
            return fail(MBAVO_EINVAL, "null context");
        DeviceGuard guard(ctx->device);
        {
            const int rcq = quiesce(ctx);
            if (rcq != MBAVO_OK)
                return rcq;
        }
        int rc = enqueue_level_points(ctx, level, d, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream); // host buffers are only borrowed for the call
        if (rc == MBAVO_OK && e != cudaSuccess)
            return fail(MBAVO_ECUDA, "upload failed: %s", cudaGetErrorString(e));
        return rc;
    }

    int mbavo_set_points_pyramid(mbavo_ctx *ctx, int n_levels, const mbavo_level_points *points)
    {
        if (!ctx || !points || n_levels < 1 || n_levels > MBAVO_MAX_LEVELS)
            return fail(MBAVO_EINVAL, "bad arguments");
        DeviceGuard guard(ctx->device);
        {
            const int rcq = quiesce(ctx);
            if (rcq != MBAVO_OK)
                return rcq;
        }
        int rc = MBAVO_OK;
        for (int l = 0; l < n_levels && rc == MBAVO_OK; ++l)
            rc = enqueue_level_points(ctx, l, points + l, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream); // one synchronisation for all levels
        if (rc == MBAVO_OK && e != cudaSuccess)
            return fail(MBAVO_ECUDA, "upload failed: %s", cudaGetErrorString(e));
        return rc;
    }

    // Keyframe + live frame + points of every level in ONE call: the level-0 images go up on the context's stream, followed by the
    // pyramid / gradient / texel kernels; the points of all levels go up on a second stream at the same time (the two copies share
    // the PCIe link, but the pyramid kernels run under the point copies); the context's stream then waits for the second one.
    // One synchronisation — or none with MBAVO_UPLOAD_ASYNC, in which case the host buffers must stay valid and unchanged until
    // the next blocking call on this context has returned (every evaluation / sweep entry point is one).
    int mbavo_set_frame(mbavo_ctx *ctx, int n_levels, int mem, const unsigned char *ref_I0, int H0, int W0,
                        const unsigned char *const *cur_I0, int n_frames, const mbavo_level_points *points, int flags)
    {
        if (!ctx || !points || (!ref_I0 && !cur_I0) || n_levels < 1 || n_levels > MBAVO_MAX_LEVELS)
            return fail(MBAVO_EINVAL, "bad arguments");
        DeviceGuard guard(ctx->device);
        cudaStream_t s = ctx->stream;
        if (!ctx->copy_stream)
        {
            CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->images_done, cudaEventDisableTiming));
            CUDA_TRY(cudaMalloc(&ctx->ready_dev, sizeof(unsigned int) * MBAVO_MAX_LEVELS));
            CUDA_TRY(cudaMemset(ctx->ready_dev, 0, sizeof(unsigned int) * MBAVO_MAX_LEVELS));
            CUDA_TRY(cudaMallocHost(&ctx->ready_src, sizeof(unsigned int)));
        }
        // nothing of an earlier frame may still be in flight on either stream (its buffers are about to be replaced)
        if (cudaStreamQuery(s) != cudaSuccess)
            CUDA_TRY(cudaStreamSynchronize(s));
        if (cudaStreamQuery(ctx->copy_stream) != cudaSuccess)
            CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
        ctx->join_pending = false;
        int rc = MBAVO_OK;
        for (int l = 0; l < n_levels && rc == MBAVO_OK; ++l)
            if (!ref_I0 && !ctx->levels[l].has_key)
                rc = fail(MBAVO_ENOTREADY, "level %d has no keyframe pyramid", l);
        // 1. every copy goes to the second stream, the images first (everything waits for them), and is issued BEFORE any kernel: the
        //    link starts working at once while the host is still enqueuing
        cudaStream_t cs = ctx->copy_stream;
        if (rc == MBAVO_OK && ref_I0)
            rc = enqueue_keyframe_pyramid(ctx, n_levels, mem, ref_I0, H0, W0, s, 1, cs);
        if (rc == MBAVO_OK && cur_I0)
            rc = enqueue_live_pyramid(ctx, n_levels, mem, cur_I0, n_frames, s, 1, cs);
        if (rc == MBAVO_OK)
        {
            cudaError_t e = cudaEventRecord(ctx->images_done, cs);
            if (e != cudaSuccess)
                rc = fail(MBAVO_ECUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
        }
        // 2. the points, coarse level first (the order a sweep needs them), behind the images on the same stream; behind each level's
        //    points goes its ready flag
        *ctx->ready_src = ++ctx->upload_epoch;
        for (int l = n_levels - 1; l >= 0 && rc == MBAVO_OK; --l)
        {
            rc = enqueue_level_points(ctx, l, points + l, ctx->copy_stream, s, false); // (the flags are cleared by the pyramid steps below)
            if (rc == MBAVO_OK)
            {
                cudaError_t e = cudaMemcpyAsync(ctx->ready_dev + l, ctx->ready_src, sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->copy_stream);
                if (e != cudaSuccess)
                    rc = fail(MBAVO_ECUDA, "ready flag: %s", cudaGetErrorString(e));
            }
        }
        if (rc == MBAVO_OK)
        {
            cudaError_t e = cudaEventRecord(ctx->copy_done, ctx->copy_stream);
            if (e != cudaSuccess)
                rc = fail(MBAVO_ECUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
        }
        // 3. the context's stream: pyramid / gradient / texel kernels as soon as the images are there (and the outlier-flag memsets)
        if (rc == MBAVO_OK)
        {
            cudaError_t e = cudaStreamWaitEvent(s, ctx->images_done, 0);
            if (e != cudaSuccess)
                rc = fail(MBAVO_ECUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(e));
        }
        // one launch per level: texels of the keyframe's level l, level l + 1 of the keyframe and of the live images, flags of level l
        for (int l = 0; l < n_levels && rc == MBAVO_OK; ++l)
        {
            LevelStore &L = ctx->levels[l];
            PyrStepParams ps{};
            int nj = 0;
            if (ref_I0)
            {
                PyrJob &j = ps.job[nj++];
                j.kind = 0, j.src = L.pyr_ref, j.Hs = L.dev.H, j.Ws = L.dev.W;
                j.pair = ctx->use_texels ? L.tex_pair : nullptr, j.quad = ctx->use_texels ? L.tex_quad : nullptr;
                j.grad = ctx->use_texels ? nullptr : reinterpret_cast<float2 *>(L.pyr_grad);
                if (l + 1 < n_levels)
                {
                    PyrJob &d = ps.job[nj++];
                    LevelStore &C = ctx->levels[l + 1];
                    d.kind = 1, d.src = L.pyr_ref, d.Hs = L.dev.H, d.Ws = L.dev.W, d.dst = C.pyr_ref, d.Hd = C.dev.H, d.Wd = C.dev.W;
                }
            }
            if (cur_I0 && l + 1 < n_levels)
                for (int f = 0; f < n_frames; ++f)
                {
                    PyrJob &d = ps.job[nj++];
                    LevelStore &C = ctx->levels[l + 1];
                    d.kind = 1, d.src = L.pyr_cur[f], d.Hs = L.dev.H, d.Ws = L.dev.W, d.dst = C.pyr_cur[f], d.Hd = C.dev.H, d.Wd = C.dev.W;
                }
            {
                PyrJob &z = ps.job[nj++];
                z.kind = 2, z.dst = L.flags, z.Wd = ctx->lim.max_num_keypoints; // tracker.cpp:600-601
            }
            ps.n_jobs = nj;
            cudaError_t e = launch_pyr_step_kernel(ps, s);
            if (e != cudaSuccess)
                rc = fail(MBAVO_ECUDA, "pyramid step: %s", cudaGetErrorString(e));
            ctx->launches += 1;
        }
        if (rc == MBAVO_OK && (flags & MBAVO_UPLOAD_ASYNC))
        {
            ctx->join_pending = true; // joined lazily: by the first entry point that needs the points on the context's stream
            ctx->upload_levels = n_levels;
            return MBAVO_OK;
        }
        cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream), e2 = cudaStreamSynchronize(s);
        if (rc == MBAVO_OK && (e1 != cudaSuccess || e2 != cudaSuccess))
            return fail(MBAVO_ECUDA, "upload failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        return rc;
    }

    int mbavo_select_points(mbavo_ctx *ctx, int n_levels, const mbavo_point_selection *sel, int *num_selected)
    {
        if (!ctx || !sel || !num_selected || n_levels < 1 || n_levels > MBAVO_MAX_LEVELS)
            return fail(MBAVO_EINVAL, "bad arguments");
        if (!sel->depth_z || !sel->pattern_xy || (sel->depth_mem != MBAVO_MEM_HOST && sel->depth_mem != MBAVO_MEM_DEVICE))
            return fail(MBAVO_EINVAL, "null depth map / pattern, or bad depth_mem");
        if (sel->cell_H < 1 || sel->cell_W < 1)
            return fail(MBAVO_EINVAL, "grid_selection cells must be positive (the tracker uses 30 x 30)");
        if (sel->patch_size < 1 || sel->patch_size > ctx->lim.max_patch_size)
            return fail(MBAVO_ECAPACITY, "patch_size %d outside [1, %d]", sel->patch_size, ctx->lim.max_patch_size);
        if (sel->num_virtual_poses < 1 || sel->num_virtual_poses > ctx->lim.max_num_virtual_poses_per_frame)
            return fail(MBAVO_ECAPACITY, "num_virtual_poses %d outside [1, %d]", sel->num_virtual_poses,
                        ctx->lim.max_num_virtual_poses_per_frame);
        if (!(sel->fx > 0) || !(sel->fy > 0))
            return fail(MBAVO_EINVAL, "fx, fy must be positive");
        for (int l = 0; l < n_levels; ++l)
            if (!ctx->levels[l].has_key)
                return fail(MBAVO_ENOTREADY, "level %d has no keyframe pyramid (mbavo_set_keyframe_pyramid comes first)", l);
        DeviceGuard guard(ctx->device);
        cudaStream_t s = ctx->stream;
        {
            const int rcq = quiesce(ctx);
            if (rcq != MBAVO_OK)
                return rcq;
        }
        const int H0 = ctx->levels[0].dev.H, W0 = ctx->levels[0].dev.W;
        SelectParams prm{};
        int total_cells = 0;
        const int cap = ctx->lim.max_num_keypoints;
        for (int l = 0; l < n_levels; ++l)
        {
            LevelStore &L = ctx->levels[l];
            SelectLevel &S = prm.lv[l];
            S.I = L.dev.ref_I, S.H = L.dev.H, S.W = L.dev.W;
            S.ch = (int)(sel->cell_H / std::pow(1.414, l)), S.cw = (int)(sel->cell_W / std::pow(1.414, l)); // FeatureDetectorBase.cpp:60-61
            if (S.ch < 1 || S.cw < 1)
                return fail(MBAVO_EINVAL, "cell %d x %d shrinks to nothing at level %d", sel->cell_H, sel->cell_W, l);
            if (S.H != H0 / (1 << l) || S.W != W0 / (1 << l))
                return fail(MBAVO_ENOTREADY, "level %d does not belong to the pyramid of level 0", l);
            S.nch = S.H / S.ch + 1, S.ncw = S.W / S.cw + 1; // :63-64
            S.cell_base = total_cells;
            total_cells += S.nch * S.ncw;
            if (L.pts_cap < (size_t)cap)
            {
                cudaFree(L.pts_xy);
                cudaFree(L.pts_z);
                L.pts_xy = L.pts_z = nullptr, L.pts_cap = 0;
                CUDA_TRY(cudaMalloc(&L.pts_xy, sizeof(double) * 2 * cap));
                CUDA_TRY(cudaMalloc(&L.pts_z, sizeof(double) * cap));
                L.pts_cap = cap;
            }
            S.xy = reinterpret_cast<double2 *>(L.pts_xy), S.z = L.pts_z;
            int rc = ensure_level_scratch(ctx, L);
            if (rc != MBAVO_OK)
                return rc;
        }
        if (ctx->sel_cells_cap < (size_t)total_cells)
        {
            cudaFree(ctx->sel_cells);
            ctx->sel_cells = nullptr, ctx->sel_cells_cap = 0;
            CUDA_TRY(cudaMalloc(&ctx->sel_cells, sizeof(int4) * total_cells));
            ctx->sel_cells_cap = total_cells;
        }
        if (!ctx->sel_count_dev)
        {
            CUDA_TRY(cudaMalloc(&ctx->sel_count_dev, sizeof(int) * MBAVO_MAX_LEVELS));
            CUDA_TRY(cudaMallocHost(&ctx->sel_count_host, sizeof(int) * MBAVO_MAX_LEVELS));
        }
        const float *depth = sel->depth_z;
        if (sel->depth_mem == MBAVO_MEM_HOST)
        {
            const size_t npix = (size_t)H0 * W0;
            if (ctx->sel_depth_cap < npix)
            {
                cudaFree(ctx->sel_depth);
                ctx->sel_depth = nullptr, ctx->sel_depth_cap = 0;
                CUDA_TRY(cudaMalloc(&ctx->sel_depth, sizeof(float) * npix));
                ctx->sel_depth_cap = npix;
            }
            CUDA_TRY(cudaMemcpyAsync(ctx->sel_depth, sel->depth_z, sizeof(float) * npix, cudaMemcpyHostToDevice, s));
            depth = ctx->sel_depth;
        }
        prm.n_levels = n_levels, prm.score_threshold = sel->score_threshold, prm.depth_z = depth, prm.W0 = W0, prm.capacity = cap;
        prm.cell_rec = ctx->sel_cells, prm.count = ctx->sel_count_dev;
        CUDA_TRY(launch_select_kernels(prm, total_cells, s));
        ctx->launches += 2;
        CUDA_TRY(cudaMemcpyAsync(ctx->sel_count_host, ctx->sel_count_dev, sizeof(int) * n_levels, cudaMemcpyDeviceToHost, s));
        for (int l = 0; l < n_levels; ++l)
        {
            LevelStore &L = ctx->levels[l];
            CUDA_TRY(cudaMemcpyAsync(L.pattern, sel->pattern_xy, sizeof(int2) * sel->patch_size, cudaMemcpyHostToDevice, s));
            L.pattern_shadow_n = 0;
            CUDA_TRY(cudaMemsetAsync(L.flags, 0, cap, s));
        }
        CUDA_TRY(cudaStreamSynchronize(s));
        for (int l = 0; l < n_levels; ++l)
        {
            num_selected[l] = ctx->sel_count_host[l];
            if (num_selected[l] > cap) // nothing is kept: the points of every level are void
            {
                for (int k = 0; k < n_levels; ++k)
                    ctx->levels[k].has_pts = false, pyramid_level_ready(ctx->levels[k]);
                return fail(MBAVO_ECAPACITY, "level %d: %d points selected, max_num_keypoints is %d", l, num_selected[l], cap);
            }
        }
        for (int l = 0; l < n_levels; ++l)
        {
            LevelStore &L = ctx->levels[l];
            const int n = num_selected[l];
            // a sharded context keeps its contiguous block of the selection (every rank selects the same points)
            int lo = 0, hi = n;
            if (ctx->shard.world > 1)
            {
                const int base = n / ctx->shard.world, rem = n % ctx->shard.world, r = ctx->shard.rank;
                lo = r * base + (r < rem ? r : rem), hi = lo + base + (r < rem ? 1 : 0);
                if (n < ctx->shard.world)
                    lo = hi = 0; // fewer points than ranks: the level is empty on EVERY rank (a collective needs them all)
                ctx->points_global[l] = n;
            }
            const double scale = (double)(1 << l); // tracker.cpp:765-771
            L.dev.xy = reinterpret_cast<const char *>(L.pts_xy + 2 * (size_t)lo), L.dev.xy_stride = 16, L.dev.xy_offset = 0;
            L.dev.z = L.pts_z + lo;
            L.num_bad = 0;
            L.dev.fx = sel->fx / scale, L.dev.fy = sel->fy / scale, L.dev.cx = sel->cx / scale, L.dev.cy = sel->cy / scale;
            L.dev.inv_fx = 1.0 / L.dev.fx, L.dev.inv_fy = 1.0 / L.dev.fy;
            L.dev.P = hi - lo, L.dev.S = sel->patch_size, L.dev.N = sel->num_virtual_poses;
            L.dev.pattern = L.pattern;
            L.dev.flags = L.flags;
            L.dev.patch_cost = L.patch_cost, L.dev.patch_cost_stride = 1;
            L.has_pts = hi > lo; // a level without points cannot be evaluated
            pyramid_level_ready(L);
        }
        return MBAVO_OK;
    }

    int mbavo_get_points(mbavo_ctx *ctx, int level, int capacity, double *xy, double *z, int *num_points)
    {
        if (!ctx || !num_points || level < 0 || level >= MBAVO_MAX_LEVELS)
            return fail(MBAVO_EINVAL, "bad arguments");
        LevelStore &L = ctx->levels[level];
        if (!L.has_pts && !(L.set && !L.has_key))
        {
            *num_points = 0;
            return L.has_key ? MBAVO_OK : fail(MBAVO_ENOTREADY, "level %d has no points", level);
        }
        const int P = L.dev.P;
        *num_points = P;
        if (!xy && !z)
            return MBAVO_OK;
        if (capacity < P)
            return fail(MBAVO_ECAPACITY, "level %d holds %d points, capacity is %d", level, P, capacity);
        DeviceGuard guard(ctx->device);
        {
            const int rcj = join_uploads(ctx);
            if (rcj != MBAVO_OK)
                return rcj;
        }
        cudaStream_t s = ctx->stream;
        if (xy)
            CUDA_TRY(cudaMemcpy2DAsync(xy, 16, L.dev.xy + L.dev.xy_offset, L.dev.xy_stride, 16, P, cudaMemcpyDeviceToHost, s));
        if (z)
            CUDA_TRY(cudaMemcpyAsync(z, L.dev.z, sizeof(double) * P, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        return MBAVO_OK;
    }

    int mbavo_set_live_images(mbavo_ctx *ctx, int level, int mem, const unsigned char *const *cur_I, int n_frames)
    {
        if (!ctx || level < 0 || level >= MBAVO_MAX_LEVELS || !ctx->levels[level].set || !cur_I)
            return fail(MBAVO_ENOTREADY, "level not set");
        LevelStore &L = ctx->levels[level];
        if (n_frames != L.dev.F)
            return fail(MBAVO_EINVAL, "n_frames %d differs from the level's %d", n_frames, L.dev.F);
        if ((mem == MBAVO_MEM_HOST) != L.owns)
            return fail(MBAVO_EINVAL, "mem must match the memory kind the level was set with");
        DeviceGuard guard(ctx->device);
        cudaStream_t s = ctx->stream;
        {
            const int rcq = quiesce(ctx);
            if (rcq != MBAVO_OK)
                return rcq;
        }
        const size_t npix = (size_t)L.dev.H * L.dev.W;
        for (int f = 0; f < n_frames; ++f)
        {
            if (!cur_I[f])
                return fail(MBAVO_EINVAL, "null image");
            if (L.owns)
                CUDA_TRY(cudaMemcpyAsync(L.cur_I[f], cur_I[f], npix, cudaMemcpyHostToDevice, s));
            else
                L.dev.cur_I[f] = cur_I[f];
        }
        CUDA_TRY(cudaMemsetAsync(const_cast<unsigned char *>(L.dev.flags), 0, L.dev.P, s)); // tracker.cpp:600-601
        L.num_bad = 0;
        CUDA_TRY(cudaStreamSynchronize(s));
        return MBAVO_OK;
    }

    int mbavo_set_outliers(mbavo_ctx *ctx, int level, const unsigned char *flags, int num_bad)
    {
        if (!ctx || level < 0 || level >= MBAVO_MAX_LEVELS || !ctx->levels[level].set)
            return fail(MBAVO_ENOTREADY, "level not set");
        DeviceGuard guard(ctx->device);
        {
            const int rcj = join_uploads(ctx);
            if (rcj != MBAVO_OK)
                return rcj;
        }
        LevelStore &L = ctx->levels[level];
        const int pts_all = ctx->shard.world > 1 && ctx->points_global[level] > 0 ? ctx->points_global[level] : L.dev.P;
        if (num_bad < 0 || num_bad >= pts_all)
            return fail(MBAVO_EINVAL, "num_bad_keypoints %d outside [0, %d)", num_bad, pts_all);
        unsigned char *dst = const_cast<unsigned char *>(L.dev.flags);
        if (flags)
            CUDA_TRY(cudaMemcpyAsync(dst, flags, L.dev.P, cudaMemcpyHostToDevice, ctx->stream));
        else
            CUDA_TRY(cudaMemsetAsync(dst, 0, L.dev.P, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        L.num_bad = flags ? num_bad : 0;
        return MBAVO_OK;
    }

    int mbavo_set_num_bad(mbavo_ctx *ctx, int level, int num_bad)
    {
        if (!ctx || level < 0 || level >= MBAVO_MAX_LEVELS || !ctx->levels[level].set)
            return fail(MBAVO_ENOTREADY, "level not set");
        const int pts_all2 = ctx->shard.world > 1 && ctx->points_global[level] > 0 ? ctx->points_global[level] : ctx->levels[level].dev.P;
        if (num_bad < 0 || num_bad >= pts_all2)
            return fail(MBAVO_EINVAL, "num_bad_keypoints %d outside [0, %d)", num_bad, pts_all2);
        ctx->levels[level].num_bad = num_bad;
        return MBAVO_OK;
    }

    int mbavo_unpack(const double *packed, int kmin, int NK, int n_knots, double *total_cost, double *H, double *g)
    {
        if (!packed || !total_cost || NK < 1 || kmin < 0 || kmin + NK > n_knots)
            return fail(MBAVO_EINVAL, "bad window [%d, %d) for %d knots", kmin, kmin + NK, n_knots);
        *total_cost = packed[0];
        if (!H)
            return MBAVO_OK;
        if (!g)
            return fail(MBAVO_EINVAL, "gradient must be given with hessian");
        // merge_hessian_gradient_cost.cpp:25-86 for a window of NK knots: t-block -> 3 (kmin + j), w-block -> 3 (n + kmin + j)
        const int n = 6 * NK + 1, Wd = 6 * n_knots;
        std::memset(H, 0, sizeof(double) * Wd * Wd);
        std::memset(g, 0, sizeof(double) * Wd);
        const int off_t = 3 * kmin, off_w = 3 * (n_knots + kmin) - 3 * NK;
        const double *ptr = packed + n;
        for (int j = 0; j < n - 1; ++j)
        {
            const int rr = j + (j < 3 * NK ? off_t : off_w);
            g[rr] += packed[j + 1];
            for (int c = j; c < n - 1; ++c, ++ptr)
            {
                const int cc = c + (c < 3 * NK ? off_t : off_w);
                H[(size_t)rr * Wd + cc] += *ptr;
                if (cc != rr)
                    H[(size_t)cc * Wd + rr] += *ptr;
            }
        }
        return MBAVO_OK;
    }

    int mbavo_evaluate(mbavo_ctx *ctx, int level, const mbavo_spline *sp, double huber_a, double *total_cost, double *H, double *g)
    {
        if (!total_cost)
            return fail(MBAVO_EINVAL, "total_cost is null");
        if ((H == nullptr) != (g == nullptr))
            return fail(MBAVO_EINVAL, "hessian and gradient must both be given or both be null");
        if (!ctx)
            return fail(MBAVO_EINVAL, "null context");
        DeviceGuard guard(ctx->device);
        EvalPlan pl;
        int rc = plan_evaluation(ctx, level, sp, H != nullptr, pl);
        if (rc != MBAVO_OK)
            return rc;
        LevelStore &L = ctx->levels[level];
        // spline_update_step.cpp:116; sharded: the point count and the outlier count are the global ones
        const long long pts = ctx->shard.world > 1 ? ctx->points_global[level] : pl.P;
        if (ctx->shard.world > 1 && pts < pl.P)
            return fail(MBAVO_ENOTREADY, "mbavo_shard_set_global_points has not been called for level %d", level);
        const long long nres = (pts - L.num_bad) * pl.F * pl.S;
        if (nres <= 0)
            return fail(MBAVO_EINVAL, "level %d: no residuals left (%lld points, %d flagged as outliers)", level, pts, L.num_bad);
        rc = run_evaluation(ctx, level, pl, ctx->packed_dev, true, 1.0 / (double)nres, huber_a);
        if (rc != MBAVO_OK)
            return rc;
        rc = wait_result(ctx);
        if (rc != MBAVO_OK)
            return rc;
        if (ctx->shard.world > 1 && ctx->result_vals[0] != ctx->result_vals[0])
            return fail(MBAVO_ENCCL, "sharded evaluation: a peer rank did not arrive within 4 s");
        if (ctx->timing)
        {
            CUDA_TRY(cudaEventSynchronize(ctx->ev1));
            cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
        }
        return mbavo_unpack(ctx->result_vals.data(), pl.kmin, pl.NK, sp->num_ctrl_knots, total_cost, H, g);
    }

    // One Hessian-pass evaluation whose per-stage intermediates come back to the host (SURVEY.md §8d stage gates): the fp64 pose
    // and spline Jacobian blocks of every exposure sample before they are rounded into the fp32 sample records, the fp64 patch
    // centres, and every pixel's raw residual and raw 1 x 6NK Jacobian row.  Runs the PRODUCT kernels with their debug stores
    // switched on (pose_kernel's dbg pointer, track_pass<DBG = true>), not a second implementation.
    int mbavo_debug_dump(mbavo_ctx *ctx, int level, const mbavo_spline *sp, double huber_a, mbavo_debug_out *out)
    {
        if (!ctx || !out)
            return fail(MBAVO_EINVAL, "null argument");
        DeviceGuard guard(ctx->device);
        EvalPlan pl;
        int rc = plan_evaluation(ctx, level, sp, true, pl);
        if (rc != MBAVO_OK)
            return rc;
        LevelStore &L = ctx->levels[level];
        if (ctx->shard.world > 1)
            return fail(MBAVO_EINVAL, "mbavo_debug_dump works on an unsharded context");
        const long long nres = (long long)(pl.P - L.num_bad) * pl.F * pl.S;
        if (nres <= 0)
            return fail(MBAVO_EINVAL, "no residuals left");
        cudaStream_t s = ctx->stream;
        rc = join_uploads(ctx);
        if (rc != MBAVO_OK)
            return rc;
        const size_t n_samp = (size_t)pl.N * pl.F, n_pts = (size_t)pl.P * pl.F, n_pix = n_pts * pl.S;
        double *d_pose = nullptr;
        double2 *d_centres = nullptr;
        float *d_r = nullptr, *d_J = nullptr;
        struct Free
        {
            void **p[4];
            ~Free()
            {
                for (auto q : p)
                    cudaFree(*q);
            }
        } undo{{(void **)&d_pose, (void **)&d_centres, (void **)&d_r, (void **)&d_J}};
        CUDA_TRY(cudaMalloc(&d_pose, sizeof(double) * n_samp * 47));
        CUDA_TRY(cudaMalloc(&d_centres, sizeof(double2) * n_pts));
        CUDA_TRY(cudaMalloc(&d_r, sizeof(float) * n_pix));
        CUDA_TRY(cudaMalloc(&d_J, sizeof(float) * n_pix * 6 * pl.NK));
        CUDA_TRY(cudaMemsetAsync(d_pose, 0, sizeof(double) * n_samp * 47, s));
        CUDA_TRY(cudaMemsetAsync(d_centres, 0, sizeof(double2) * n_pts, s));
        CUDA_TRY(cudaMemsetAsync(d_r, 0, sizeof(float) * n_pix, s));
        CUDA_TRY(cudaMemsetAsync(d_J, 0, sizeof(float) * n_pix * 6 * pl.NK, s));
        // small block shape, texel / direct variant as the level dictates
        pl.big = false;
        pl.smem = track_kernel_smem_bytes(pl.K, pl.NK, true, false, pl.N, pl.S, pl.TP);
        const int wpb = track_warps(true, pl.NK, false);
        int want = (pl.batches_per_frame + wpb - 1) / wpb;
        pl.grid = dim3(want < 4 * ctx->num_sms ? want : 4 * ctx->num_sms, pl.F, 1);
        rc = ensure_block_partials(ctx, (size_t)pl.grid.x * pl.grid.y * (pl.E + 1)); // (rows padded to an even pitch)
        if (rc != MBAVO_OK)
            return rc;
        CUDA_TRY(launch_pose_kernel(pl.K, ctx->stage, pl.N * pl.F, 1, ctx->samples, ctx->mid, ctx->seg_end, s, nullptr, 0, false, kBufA,
                                    ctx->samples_stride, kMidDoubles * kMaxFrames, kMaxSegments * kMaxFrames, d_pose));
        TrackParams prm;
        fill_track_params(ctx, level, pl, ctx->packed_dev, true, 1.0 / (double)nres, huber_a, nullptr, prm);
        prm.dbg_centres = d_centres, prm.dbg_r = d_r, prm.dbg_J = d_J;
        cudaError_t e = launch_track_debug_kernel(pl.K, pl.NK, prm, pl.grid, pl.smem, s);
        if (e == cudaErrorNotSupported)
        {
            cudaGetLastError();
            return fail(MBAVO_ECAPACITY, "mbavo_debug_dump is built for k=2 with 2 or 3 knots in the window and k=4 with 4 (texel path)");
        }
        if (e != cudaSuccess)
            return fail(MBAVO_ECUDA, "debug kernel: %s", cudaGetErrorString(e));
        ctx->launches += 2;
        L.last_eval_frames = pl.F;
        rc = wait_result(ctx);
        if (rc != MBAVO_OK)
            return rc;
        CUDA_TRY(cudaStreamSynchronize(s));
        out->kmin = pl.kmin, out->knot_window = pl.NK, out->spline_deg_k = pl.K;
        out->cost = ctx->result_vals[0];
        std::vector<double> hp(n_samp * 47);
        CUDA_TRY(cudaMemcpy(hp.data(), d_pose, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost));
        for (size_t g = 0; g < n_samp; ++g)
        {
            const double *o = hp.data() + g * 47;
            if (out->poses_tq)
                std::memcpy(out->poses_tq + 7 * g, o, sizeof(double) * 7);
            if (out->blend_weights)
                std::memcpy(out->blend_weights + pl.K * g, o + 7, sizeof(double) * pl.K);
            if (out->theta)
                std::memcpy(out->theta + 9 * pl.K * g, o + 11, sizeof(double) * 9 * pl.K);
            if (out->segment_start_knot)
                out->segment_start_knot[g] = pl.kmin + ctx->stage.seg_off[g];
        }
        if (out->patch_centres)
            CUDA_TRY(cudaMemcpy(out->patch_centres, d_centres, sizeof(double2) * n_pts, cudaMemcpyDeviceToHost));
        if (out->residuals)
            CUDA_TRY(cudaMemcpy(out->residuals, d_r, sizeof(float) * n_pix, cudaMemcpyDeviceToHost));
        if (out->jacobians)
            CUDA_TRY(cudaMemcpy(out->jacobians, d_J, sizeof(float) * n_pix * 6 * pl.NK, cudaMemcpyDeviceToHost));
        return MBAVO_OK;
    }

    // EXPERIMENT: one cost-only evaluation through the TMA-staged variant of the cost pass (csrc/track_cost_tma.cu) — per point a
    // box_w x box_h tile of the 8-bit keyframe fetched by cp.async.bulk.tensor.2d into shared memory, taps from the tile, global
    // fallback outside it.  Same arithmetic and result as mbavo_evaluate's cost-only branch; returns the kernel time (CUDA
    // events) and the fraction of samples served from the tiles so that it can be judged against the product kernel.
    int mbavo_debug_cost_tma(mbavo_ctx *ctx, int level, const mbavo_spline *sp, double huber_a, int box_w, int box_h, double *total_cost,
                             float *kernel_ms, double *tile_fraction)
    {
        if (!ctx || !total_cost)
            return fail(MBAVO_EINVAL, "null argument");
        DeviceGuard guard(ctx->device);
        EvalPlan pl;
        int rc = plan_evaluation(ctx, level, sp, false, pl);
        if (rc != MBAVO_OK)
            return rc;
        LevelStore &L = ctx->levels[level];
        pl.TP = 4, pl.PH = 1, pl.batches_per_frame = (pl.P + 3) / 4; // the variant's own mapping: lane = pixel, 4 points per warp batch
        if (ctx->shard.world > 1 || pl.F != 1 || pl.S != 8 || !L.dev.ref_quad)
            return fail(MBAVO_ECAPACITY, "the TMA variant is built for one frame, 8-pixel patterns, texel levels, unsharded contexts");
        if (box_w < 16 || box_w > 256 || box_w % 16 || box_h < 2 || box_h > 256 || (box_w * box_h) % 128 || L.dev.W % 16 ||
            ((uintptr_t)L.dev.ref_I & 15))
            return fail(MBAVO_EINVAL, "box must be 16k x h with 128 | box bytes, image width a multiple of 16, image 16-byte aligned");
        const size_t smem = track_cost_tma_smem_bytes(pl.K, pl.N, pl.S, pl.TP, box_w, box_h);
        if (smem > 220 * 1024)
            return fail(MBAVO_ECAPACITY, "box %d x %d needs %zu B of shared memory per block", box_w, box_h, smem);
        typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
            return fail(MBAVO_ECUDA, "cuTensorMapEncodeTiled is not available");
        alignas(64) CUtensorMap map;
        const cuuint64_t gdim[2] = {(cuuint64_t)L.dev.W, (cuuint64_t)L.dev.H}, gstride[1] = {(cuuint64_t)L.dev.W};
        const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, estr[2] = {1, 1};
        const CUresult cr = ((encode_fn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<unsigned char *>(L.dev.ref_I), gdim, gstride, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS)
            return fail(MBAVO_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
        cudaStream_t s = ctx->stream;
        rc = join_uploads(ctx);
        if (rc != MBAVO_OK)
            return rc;
        double *d_cost = nullptr;
        CUDA_TRY(cudaMalloc(&d_cost, sizeof(double) + 2 * sizeof(unsigned long long)));
        unsigned long long *d_cnt = reinterpret_cast<unsigned long long *>(d_cost + 1);
        CUDA_TRY(cudaMemsetAsync(d_cost, 0, sizeof(double) + 2 * sizeof(unsigned long long), s));
        CUDA_TRY(launch_pose_kernel(pl.K, ctx->stage, pl.N, 0, ctx->samples, ctx->mid, ctx->seg_end, s, nullptr, 0, false, kBufA, ctx->samples_stride,
                                    kMidDoubles * kMaxFrames, kMaxSegments * kMaxFrames));
        TrackParams prm;
        const long long nres = (long long)(pl.P - L.num_bad) * pl.F * pl.S;
        fill_track_params(ctx, level, pl, ctx->packed_dev, false, 1.0 / (double)nres, huber_a, nullptr, prm);
        int per_sm = (int)((220 * 1024) / smem);
        per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
        int grid = (pl.batches_per_frame + 7) / 8;
        grid = grid < ctx->num_sms * per_sm ? grid : ctx->num_sms * per_sm;
        CUDA_TRY(cudaEventRecord(ctx->ev0, s));
        cudaError_t e = launch_track_cost_tma(pl.K, prm, &map, box_w, box_h, grid, smem, d_cost, d_cnt, s);
        if (e != cudaSuccess)
        {
            cudaFree(d_cost);
            return fail(MBAVO_ECUDA, "TMA cost kernel: %s", cudaGetErrorString(e));
        }
        CUDA_TRY(cudaEventRecord(ctx->ev1, s));
        ctx->launches += 2;
        struct
        {
            double cost;
            unsigned long long cnt[2];
        } h;
        e = cudaMemcpyAsync(&h, d_cost, sizeof h, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(s);
        cudaFree(d_cost);
        if (e != cudaSuccess)
            return fail(MBAVO_ECUDA, "TMA cost kernel failed: %s", cudaGetErrorString(e));
        *total_cost = h.cost;
        if (kernel_ms)
            cudaEventElapsedTime(kernel_ms, ctx->ev0, ctx->ev1);
        if (tile_fraction)
            *tile_fraction = (h.cnt[0] + h.cnt[1]) ? (double)h.cnt[0] / (double)(h.cnt[0] + h.cnt[1]) : 0.0;
        return MBAVO_OK;
    }

    int mbavo_evaluate_async(mbavo_ctx *ctx, int level, const mbavo_spline *sp, double huber_a, int with_hessian,
                             long long num_residuals_global, double *packed_dev, int *kmin, int *knot_window)
    {
        if (!ctx || !packed_dev)
            return fail(MBAVO_EINVAL, "null argument");
        DeviceGuard guard(ctx->device);
        EvalPlan pl;
        int rc = plan_evaluation(ctx, level, sp, with_hessian != 0, pl);
        if (rc != MBAVO_OK)
            return rc;
        LevelStore &L = ctx->levels[level];
        const long long nres = num_residuals_global > 0 ? num_residuals_global : (long long)(pl.P - L.num_bad) * pl.F * pl.S;
        if (nres <= 0)
            return fail(MBAVO_EINVAL, "level %d: no residuals left (%d points, %d flagged as outliers)", level, pl.P, L.num_bad);
        if (kmin)
            *kmin = pl.kmin;
        if (knot_window)
            *knot_window = pl.NK;
        // the caller reduces the vectors itself (e.g. ncclAllReduce): no mailbox exchange for this launch
        const ShardParams keep = ctx->shard;
        ctx->shard.world = 0;
        rc = run_evaluation(ctx, level, pl, packed_dev, false, 1.0 / (double)nres, huber_a);
        ctx->shard = keep;
        return rc;
    }

    // mbavo_gn_sweep without host round trips between the evaluations: the kernels of all levels are enqueued at once
    // (pose -> Hessian pass [+ solve, candidate] -> pose at the candidate -> cost pass [+ record, commit]), chained by
    // programmatic dependent launches; the host waits once.  Returns 1 (not an error) when the caller has to fall back to
    // the evaluation-by-evaluation path: feature disabled, event timing on, or the device-side solve met normal equations
    // that are not safely positive definite / a negative model decrease.
    int mbavo_gn_sweep_device(mbavo_ctx *ctx, int level_coarse, int level_fine, int chain, int k, double t0, double dt, int n,
                              double *knots_t, double *knots_R, double radius, double huber_a, double *costs)
    {
        if (!ctx || !ctx->use_device_sweep || ctx->force_phases > 0)
            return 1;
        // event timing brackets ONE kernel: the persistent sweep kernel qualifies, a chain of per-pass launches does not
        const bool timing_needs_persistent = ctx->timing;
        const int nlev = level_coarse - level_fine + 1;
        if (nlev < 1 || nlev > MBAVO_MAX_LEVELS || n < 2 || n > 16)
            return 1;
        DeviceGuard guard(ctx->device);
        if (!ctx->gn_state)
        {
            CUDA_TRY(cudaMalloc(&ctx->gn_state, sizeof(GnState)));
            CUDA_TRY(cudaMemset(ctx->gn_state, 0, sizeof(GnState)));
        }
        mbavo_spline sp{k, t0, dt, n, knots_t, knots_R};
        unsigned long long seq_of_level[MBAVO_MAX_LEVELS] = {};
        // The sample records depend on the knots, the frame times and the number of exposure samples only — not on the level.
        // When every level of the sweep uses the same sample count, a level after the first finds the records of its knots in
        // one of the two buffers (the previous level's, or its committed candidate's) and runs no pose kernel of its own.
        // (point-sharded contexts keep one pose kernel per evaluation: the shared-record form has only been measured on one GPU)
        bool reuse = nlev > 1 && ctx->shard.world <= 1;
        for (int li = 1; li < nlev && reuse; ++li)
            reuse = ctx->levels[level_coarse - li].set && ctx->levels[level_coarse].set &&
                    ctx->levels[level_coarse - li].dev.N == ctx->levels[level_coarse].dev.N;
        // ---- the whole sweep inside ONE launch (sweep_kernel) when every level shares one exposure-sample count, one frame, the
        // texel path and a knot window the kernel is instantiated for; else pass by pass below
        bool persistent = reuse || (nlev == 1 && ctx->shard.world <= 1);
        if (ctx->shard.world > 1)
        {
            persistent = !ctx->shard_shares_device; // every rank needs its whole GPU
            for (int li = 1; li < nlev && persistent; ++li)
                persistent = ctx->levels[level_coarse - li].set && ctx->levels[level_coarse].set &&
                             ctx->levels[level_coarse - li].dev.N == ctx->levels[level_coarse].dev.N;
        }
        persistent = persistent && ctx->use_persistent;
        if (timing_needs_persistent && !persistent)
            return 1;
        for (int li = 0; li < nlev && persistent; ++li)
        {
            const LevelStore &L = ctx->levels[level_coarse - li];
            persistent = L.set && L.dev.F == 1 && L.dev.ref_pair != nullptr && L.dev.S == ctx->levels[level_coarse].dev.S;
        }
        if (persistent)
        {
            const int rcp = run_persistent_sweep(ctx, level_coarse, nlev, chain, &sp, radius, huber_a, seq_of_level);
            if (rcp < 0)
                return rcp;
            persistent = rcp == MBAVO_OK; // 1: no instantiation for this window, fall through
            if (timing_needs_persistent && !persistent)
                return 1;
        }
        for (int li = 0; li < nlev && !persistent; ++li)
        {
            const int level = level_coarse - li;
            LevelStore &L = ctx->levels[level];
            for (int pass = 0; pass < 2; ++pass)
            {
                EvalPlan pl{};
                int rc = plan_evaluation(ctx, level, &sp, pass == 0, pl);
                if (rc != MBAVO_OK)
                    return rc;
                const long long pts = ctx->shard.world > 1 ? ctx->points_global[level] : pl.P;
                if (ctx->shard.world > 1 && pts < pl.P)
                    return fail(MBAVO_ENOTREADY, "mbavo_shard_set_global_points has not been called for level %d", level);
                const long long nres = (pts - L.num_bad) * pl.F * pl.S;
                if (nres <= 0)
                    return fail(MBAVO_EINVAL, "level %d: no residuals left (%lld points, %d flagged as outliers)", level, pts, L.num_bad);
                SweepLaunch sw;
                sw.gn.state = ctx->gn_state;
                sw.gn.mode = pass == 0 ? 1 : 2;
                sw.gn.n_knots = n, sw.gn.kmin = pl.kmin;
                sw.gn.chain = chain, sw.gn.slot = li, sw.gn.last = (li == nlev - 1) ? 1 : 0;
                sw.gn.radius = radius;
                sw.knots_from = pass == 1 ? 2 : (li == 0 ? 0 : 1);
                sw.first = li == 0 && pass == 0;
                if (reuse)
                {
                    sw.skip_pose = pass == 0 && li > 0;
                    sw.pose_with_j = pass == 1 && li < nlev - 1;
                    if (chain) // the buffer roles swap whenever a candidate is committed: read the state on the device
                        sw.buf_select = pass == 0 ? (li == 0 ? kBufA : kBufCur) : kBufCand;
                    else
                        sw.buf_select = pass == 0 ? kBufA : kBufB;
                }
                rc = run_evaluation(ctx, level, pl, ctx->packed_dev, true, 1.0 / (double)nres, huber_a, &sw);
                if (rc != MBAVO_OK)
                    return rc;
                if (pass == 1)
                    seq_of_level[li] = ctx->seq;
            }
        }
        // wait for the last level's scalars and the final knots, then read the earlier levels' (published before them)
        double scal[4 * MBAVO_MAX_LEVELS], knots_out[7 * 16];
        int rcw = wait_published(ctx, 4 * (nlev - 1), 4, ctx->seq, scal + 4 * (nlev - 1), "sweep");
        if (rcw == MBAVO_OK)
            rcw = wait_published(ctx, 4 * MBAVO_MAX_LEVELS, 7 * n, ctx->seq, knots_out, "sweep");
        if (rcw != MBAVO_OK)
        {
            if (persistent && ctx->sweep_ctl)
            {
                // a block gave up (SweepCtl::abort) or the launch failed: drain the stream and re-arm the pass barrier
                if (ctx->copy_stream)
                    cudaStreamSynchronize(ctx->copy_stream);
                ctx->join_pending = false;
                cudaStreamSynchronize(ctx->stream);
                cudaMemset(ctx->sweep_ctl, 0, sizeof(SweepCtl));
                cudaMemset(ctx->counter, 0, sizeof(unsigned int));
                cudaGetLastError();
                ctx->sweep_base = 0;
            }
            return rcw;
        }
        bool fallback = false, peer_lost = false;
        const volatile unsigned long long *words = reinterpret_cast<const volatile unsigned long long *>(ctx->result_host);
        for (int li = 0; li < nlev; ++li)
        {
            for (int e = 0; e < 4; ++e)
                if (!read_published(words + 2 * (4 * li + e), seq_of_level[li], scal + 4 * li + e))
                    return fail(MBAVO_ECUDA, "sweep: level %d did not publish", level_coarse - li);
            const double c = scal[4 * li], cc = scal[4 * li + 1];
            if (scal[4 * li + 2] != 0.0)
                fallback = true;
            if (ctx->shard.world > 1 && (c != c || cc != cc))
                peer_lost = true;
            if (costs)
                costs[2 * li] = c, costs[2 * li + 1] = cc;
        }
        if (peer_lost)
            return fail(MBAVO_ENCCL, "sharded sweep: a peer rank did not arrive within 4 s");
        if (fallback)
        {
            fail(1, "device sweep declined: solve status %g %g %g %g ...", scal[2], nlev > 1 ? scal[6] : 0.0, nlev > 2 ? scal[10] : 0.0,
                 nlev > 3 ? scal[14] : 0.0);
            return 1;
        }
        ++ctx->device_sweeps;
        if (persistent)
        {
            ++ctx->persistent_sweeps;
            ctx->join_pending = false; // every level's Hessian pass saw its ready flag: the point copies have all landed
        }
        if (chain)
        {
            for (int e = 0; e < 3 * n; ++e)
                knots_t[e] = knots_out[e];
            for (int e = 0; e < 4 * n; ++e)
                knots_R[e] = knots_out[3 * n + e];
        }
        return MBAVO_OK;
    }

    int mbavo_keyframe_stats(mbavo_ctx *ctx, int level, const double *poses_tq, double *avg_flow, double *avg_kernel_len)
    {
        if (!ctx || !poses_tq || !avg_flow || !avg_kernel_len || level < 0 || level >= MBAVO_MAX_LEVELS || !ctx->levels[level].set)
            return fail(MBAVO_ENOTREADY, "level not set");
        DeviceGuard guard(ctx->device);
        {
            const int rcj = join_uploads(ctx);
            if (rcj != MBAVO_OK)
                return rcj;
        }
        LevelStore &L = ctx->levels[level];
        cudaStream_t s = ctx->stream;
        if (!ctx->kf_dev)
        {
            CUDA_TRY(cudaMalloc(&ctx->kf_dev, sizeof(double) * 32));
            CUDA_TRY(cudaMallocHost(&ctx->kf_host, sizeof(double) * 32));
        }
        std::memcpy(ctx->kf_host + 8, poses_tq, sizeof(double) * 21);
        CUDA_TRY(cudaMemcpyAsync(ctx->kf_dev + 8, ctx->kf_host + 8, sizeof(double) * 21, cudaMemcpyHostToDevice, s));
        ShardParams sh = ctx->shard;
        if (sh.world > 1)
            sh.seq = ++ctx->aux_seq;
        keyframe_kernel<<<1, 256, 0, s>>>(L.dev.xy, L.dev.xy_stride, L.dev.xy_offset, L.dev.z, L.dev.P, L.dev.fx, L.dev.fy, L.dev.cx,
                                           L.dev.cy, ctx->kf_dev + 8, ctx->kf_dev, sh);
        CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
        CUDA_TRY(cudaMemcpyAsync(ctx->kf_host, ctx->kf_dev, sizeof(double) * 3, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (ctx->kf_host[2] != 0.0)
            return fail(MBAVO_ENCCL, "sharded keyframe statistics: a peer rank did not arrive within 4 s");
        const double n = ctx->shard.world > 1 && ctx->points_global[level] > 0 ? ctx->points_global[level] : L.dev.P;
        *avg_flow = (double)sqrtf((float)(ctx->kf_host[0] / n));        // tracker.cpp:244-245: sqrtf of the mean
        *avg_kernel_len = (double)sqrtf((float)(ctx->kf_host[1] / n));
        return MBAVO_OK;
    }

    int mbavo_patch_costs(mbavo_ctx *ctx, int level, double *out)
    {
        if (!ctx || !out || level < 0 || level >= MBAVO_MAX_LEVELS || !ctx->levels[level].set)
            return fail(MBAVO_ENOTREADY, "level not set");
        DeviceGuard guard(ctx->device);
        {
            const int rcj = join_uploads(ctx);
            if (rcj != MBAVO_OK)
                return rcj;
        }
        LevelStore &L = ctx->levels[level];
        if (L.last_eval_frames == 0)
            return fail(MBAVO_ENOTREADY, "no evaluation has run on level %d", level);
        CUDA_TRY(cudaMemcpy2DAsync(out, sizeof(double), L.dev.patch_cost, sizeof(double) * L.dev.patch_cost_stride, sizeof(double),
                                   (size_t)L.dev.P * L.dev.F, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return MBAVO_OK;
    }

    int mbavo_detect_outliers(mbavo_ctx *ctx, int level, double k_sigma, int *num_bad)
    {
        if (!ctx || !num_bad || level < 0 || level >= MBAVO_MAX_LEVELS || !ctx->levels[level].set)
            return fail(MBAVO_ENOTREADY, "level not set");
        DeviceGuard guard(ctx->device);
        {
            const int rcj = join_uploads(ctx);
            if (rcj != MBAVO_OK)
                return rcj;
        }
        LevelStore &L = ctx->levels[level];
        if (L.last_eval_frames == 0)
            return fail(MBAVO_ENOTREADY, "no evaluation has run on level %d", level);
        // "only works for tracking one frame" (blur_aware_direct_tracker.cpp:641): statistics over frame 0's patches
        ShardParams sh = ctx->shard;
        if (sh.world > 1)
        {
            sh.seq = ctx->aux_seq + 1;
            ctx->aux_seq += 3;
        }
        outlier_kernel<<<1, 1024, 0, ctx->stream>>>(L.dev.patch_cost, L.dev.P, L.dev.patch_cost_stride, k_sigma,
                                                    const_cast<unsigned char *>(L.dev.flags), ctx->outlier_result_dev, sh);
        CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
        CUDA_TRY(cudaMemcpyAsync(ctx->outlier_result_host, ctx->outlier_result_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->outlier_result_host[1] != 0)
            return fail(MBAVO_ENCCL, "sharded outlier detection: a peer rank did not arrive within 4 s");
        L.num_bad = ctx->outlier_result_host[0];
        *num_bad = L.num_bad;
        return MBAVO_OK;
    }

    // ---- point sharding over the GPUs of one node ------------------------------------------------------------------
    // The mailbox is zeroed HERE, by mbavo_shard_export, i.e. before its address / IPC handle can reach a peer — never in
    // mbavo_shard_connect: a peer that connects first may already have written its first vector into this rank's mailbox
    // by the time this rank connects, and a later memset would wipe it.  The sequence numbers restart at 0 with every
    // export + connect round, so a connect requires a fresh export (mailbox_fresh).
    static int ensure_mailbox(mbavo_ctx *ctx)
    {
        if (!ctx->mailbox)
            CUDA_TRY(cudaMalloc(&ctx->mailbox, sizeof(Mailbox)));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaMemset(ctx->mailbox, 0, sizeof(Mailbox)));
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
        static_assert(sizeof(prop.uuid) == 16, "device UUID");
        CUDA_TRY(cudaMemcpy(ctx->mailbox->owner_uuid, &prop.uuid, 16, cudaMemcpyHostToDevice));
        std::memcpy(ctx->device_uuid, &prop.uuid, 16);
        CUDA_TRY(cudaDeviceSynchronize());
        ctx->mailbox_fresh = true;
        return MBAVO_OK;
    }

    int mbavo_shard_export(mbavo_ctx *ctx, void *handle_out, void **mailbox_ptr_out)
    {
        if (!ctx)
            return fail(MBAVO_EINVAL, "null context");
        DeviceGuard guard(ctx->device);
        int rc = ensure_mailbox(ctx);
        if (rc != MBAVO_OK)
            return rc;
        if (handle_out)
        {
            static_assert(sizeof(cudaIpcMemHandle_t) == MBAVO_IPC_HANDLE_BYTES, "IPC handle size");
            cudaIpcMemHandle_t h;
            CUDA_TRY(cudaIpcGetMemHandle(&h, ctx->mailbox));
            std::memcpy(handle_out, &h, sizeof h);
        }
        if (mailbox_ptr_out)
            *mailbox_ptr_out = ctx->mailbox;
        return MBAVO_OK;
    }

    int mbavo_shard_connect(mbavo_ctx *ctx, int world, int rank, const void *handles, void *const *mailbox_ptrs)
    {
        if (!ctx || world < 1 || world > kMaxShards || rank < 0 || rank >= world || (world > 1 && !handles && !mailbox_ptrs))
            return fail(MBAVO_EINVAL, "bad sharding arguments (world %d, rank %d, at most %d ranks)", world, rank, kMaxShards);
        DeviceGuard guard(ctx->device);
        if (!ctx->mailbox || !ctx->mailbox_fresh)
            return fail(MBAVO_ENOTREADY, "mbavo_shard_connect needs a fresh mbavo_shard_export of this context (it zeroes the mailbox "
                                         "before the handle is shared; sequence numbers restart with every connection)");
        ctx->mailbox_fresh = false;
        mbavo_shard_disconnect(ctx);
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ShardParams sh{};
        sh.world = world, sh.rank = rank;
        for (int r = 0; r < world; ++r)
        {
            if (r == rank)
                sh.peer[r] = ctx->mailbox;
            else if (mailbox_ptrs)
            {
                // same process: plain peer access (a no-op on the same device)
                cudaPointerAttributes at{};
                CUDA_TRY(cudaPointerGetAttributes(&at, mailbox_ptrs[r]));
                if (at.device != ctx->device)
                {
                    cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                        return fail(MBAVO_ECUDA, "peer access %d -> %d: %s", ctx->device, at.device, cudaGetErrorString(e));
                    cudaGetLastError();
                }
                sh.peer[r] = static_cast<Mailbox *>(mailbox_ptrs[r]);
            }
            else
            {
                cudaIpcMemHandle_t h;
                std::memcpy(&h, static_cast<const char *>(handles) + (size_t)r * MBAVO_IPC_HANDLE_BYTES, sizeof h);
                void *p = nullptr;
                CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
                sh.peer[r] = static_cast<Mailbox *>(p);
                ctx->peer_is_ipc[r] = true;

            }
        }
        // a peer on THIS GPU (another context of this process, or another process time-slicing the device)?  Its mailbox says
        // which GPU it lives on.  Such ranks cannot all keep a whole-GPU persistent grid resident, so their sweeps run pass by pass.
        for (int r = 0; r < world; ++r)
        {
            if (r == rank)
                continue;
            unsigned char peer_uuid[16];
            CUDA_TRY(cudaMemcpy(peer_uuid, sh.peer[r]->owner_uuid, 16, cudaMemcpyDeviceToHost));
            if (std::memcmp(peer_uuid, ctx->device_uuid, 16) == 0)
                ctx->shard_shares_device = true;
        }
        ctx->shard = sh;
        ctx->shard_seq = ctx->aux_seq = 0; // the sequence numbers restart with the connection (mailbox zeroed by the export)
        return MBAVO_OK;
    }

    int mbavo_shard_disconnect(mbavo_ctx *ctx)
    {
        if (!ctx)
            return MBAVO_OK;
        DeviceGuard guard(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        for (int r = 0; r < kMaxShards; ++r)
        {
            if (ctx->peer_is_ipc[r] && ctx->shard.peer[r])
                cudaIpcCloseMemHandle(ctx->shard.peer[r]);
            ctx->peer_is_ipc[r] = false;
        }
        ctx->shard = ShardParams{};
        ctx->shard_shares_device = false;
        return MBAVO_OK;
    }

    int mbavo_shard_set_global_points(mbavo_ctx *ctx, int level, int num_keypoints_global)
    {
        if (!ctx || level < 0 || level >= MBAVO_MAX_LEVELS || num_keypoints_global < 1)
            return fail(MBAVO_EINVAL, "bad level / point count");
        ctx->points_global[level] = num_keypoints_global;
        return MBAVO_OK;
    }

    // Development aid (not declared in mbavo.h): globaltimer stamps of the tracking kernel's phases of the last launch;
    // only MBAVO_PROFILE_PHASES builds of the kernel write them.  out: 64 rows (launches since the last call) x 16 stamps (ns).
    int mbavo_debug_phase_times(mbavo_ctx *ctx, unsigned long long *out)
    {
        if (!ctx || !out)
            return MBAVO_EINVAL;
        DeviceGuard guard(ctx->device);
        const size_t bytes = 64 * 16 * sizeof(unsigned long long);
        if (!ctx->phase_times_dev)
        {
            CUDA_TRY(cudaMalloc(&ctx->phase_times_dev, bytes));
            CUDA_TRY(cudaMemset(ctx->phase_times_dev, 0, bytes));
        }
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaMemcpy(out, ctx->phase_times_dev, bytes, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemset(ctx->phase_times_dev, 0, bytes));
        ctx->trace_row = 0; // re-arm: the next launch writes row 0
        return MBAVO_OK;
    }

    long long mbavo_kernel_launches(const mbavo_ctx *ctx) { return ctx ? ctx->launches : 0; }
    long long mbavo_device_sweeps(const mbavo_ctx *ctx) { return ctx ? ctx->device_sweeps : 0; }
    long long mbavo_persistent_sweeps(const mbavo_ctx *ctx) { return ctx ? ctx->persistent_sweeps : 0; }

    int mbavo_sweep_pass_times(mbavo_ctx *ctx, double *us_out, int capacity, int *num_passes)
    {
        if (!ctx || !us_out || !num_passes)
            return fail(MBAVO_EINVAL, "null argument");
        *num_passes = 0;
        if (!ctx->sweep_pass_times || ctx->sweep_last_levels < 1)
            return fail(MBAVO_ENOTREADY, "no persistent sweep has run on this context");
        const int n = 2 * ctx->sweep_last_levels;
        if (capacity < n)
            return fail(MBAVO_ECAPACITY, "%d passes, capacity %d", n, capacity);
        DeviceGuard guard(ctx->device);
        unsigned long long t[1 + 2 * kMaxSweepLevels];
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaMemcpy(t, ctx->sweep_pass_times, sizeof(unsigned long long) * (n + 1), cudaMemcpyDeviceToHost));
        for (int p = 0; p < n; ++p)
            us_out[p] = (double)(t[p + 1] - t[p]) * 1e-3;
        *num_passes = n;
        return MBAVO_OK;
    }

    int mbavo_level_uses_texels(const mbavo_ctx *ctx, int level)
    {
        if (!ctx || level < 0 || level >= MBAVO_MAX_LEVELS || !ctx->levels[level].set)
            return -1;
        return ctx->levels[level].dev.ref_pair != nullptr ? 1 : 0;
    }

    int mbavo_enable_kernel_timing(mbavo_ctx *ctx, int enable)
    {
        if (!ctx)
            return fail(MBAVO_EINVAL, "null context");
        ctx->timing = enable != 0;
        ctx->last_ms = -1.f;
        return MBAVO_OK;
    }

    float mbavo_last_kernel_ms(mbavo_ctx *ctx)
    {
        if (!ctx || !ctx->timing)
            return -1.f;
        // (also valid after mbavo_evaluate_async: waits for the tracking kernel of the last launch)
        DeviceGuard guard(ctx->device);
        float ms = -1.f;
        if (cudaEventSynchronize(ctx->ev1) != cudaSuccess || cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess)
        {
            cudaGetLastError();
            return ctx->last_ms;
        }
        return ms;
    }
}
