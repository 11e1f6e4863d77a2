// Internal device-side data layout shared by the pose kernel, the tracking kernels and the host API.
#ifndef MBAVO_DEVICE_H_
#define MBAVO_DEVICE_H_

#include <cuda_runtime.h>
#include <stdint.h>

namespace mbavo
{
    constexpr int kMaxFrames = 16;
    constexpr int kMaxKnotWindow = 8;  // NK
    constexpr int kMaxSegments = 7;    // NK - k + 1 for k = 2
    // Warps per block of the tracking kernel.  The cost-only pass runs 4 blocks of 8 warps per SM.  The Hessian pass has
    // two shapes: "big" = ONE block of 20 warps per SM at <= 96 registers (16 warps for knot windows >= 4, whose rows
    // need more shared memory) — more resident warps and half as many per-block partials for the last-block reduction,
    // best once there are enough batches to occupy the SMs — and "small" = two blocks of 8 warps per SM, which spreads
    // the few batches of a coarse pyramid level over more SMs (profiles/r1_history.md).
#ifndef MBAVO_BIG_WARPS
#define MBAVO_BIG_WARPS 20 // warps of a "big" block for knot windows <= 3 (16: 128 registers per thread instead of 96)
#endif
    __host__ __device__ constexpr int track_warps(bool with_j, int NK, bool big) { return (with_j && big) ? (NK <= 3 ? MBAVO_BIG_WARPS : 16) : 8; }

    // One exposure sample (virtual pose) as the tracking kernel consumes it: fp32, 16-byte aligned records laid out so
    // that every operand PAIR of the kernel's packed FFMA2 arithmetic is an aligned pair of one 128-bit shared load.
    // With Rm = R - I (rotation matrix of the pose quaternion minus identity, rounded after the subtraction,
    // compute_virtual_camera_poses.cu:102-109) and t the translation:
    //   [0..3]    (Rm00 Rm01) (Rm10 Rm11)                  row pairs of the upper-left block:  R^T g
    //   [4..7]    (Rm20 Rm21) Rm02 Rm12
    //   [8..11]   Rm22 tz (tx ty)
    //   then per knot j of the segment (10 floats each):
    //     wt_j                      translation blend weight                               (SplineFunctor.h:30-40, 74-91)
    //     Th00 Th10 Th20            first column of Theta_j
    //     (Th01 Th02) (Th11 Th12) (Th21 Th22)   remaining columns as row pairs
    //   Theta_j (3x3) = d theta / d w_j, theta the right perturbation of the pose rotation:
    //   dq/dw_j = L(q) [I/2; 0] Theta_j                                                    (SplineFunctor.h:178-213, 274-361)
    constexpr int kRecGeom = 12;
    __host__ __device__ constexpr int sample_rec_floats(int K) { return (kRecGeom + 10 * K + 3) / 4 * 4; }

    // per-frame fp64 data for the patch centre (compute_local_patches_xy.cu:26-49): R_r2c (9) and t_r2c (3)
    constexpr int kMidDoubles = 12;

    // Per-evaluation spline state, passed BY VALUE as the pose kernel's launch parameter (2.3 KB of the 4 KB parameter
    // space): no host -> device copy precedes an evaluation.
    struct EvalStage
    {
        double knots_t[3 * 16];
        double knots_R[4 * 16];
        double cap[kMaxFrames];
        double exp_time[kMaxFrames];
        double t0, dt;
        int n_knots, K, N, F, kmin, NK;
        unsigned char seg_off[kMaxFrames * 64]; // host-computed segment start knot of every sample minus kmin (authoritative)
    };

    // Keyframe texels, built once per mbavo_set_level / mbavo_set_frame (make_pair_texel, track_common.cu).  They carry the
    // values of ref_I / ref_dIxy bit for bit in a gather-friendly layout:
    //   patch texel (16 B per pixel): the 4 x 4 BYTE neighbourhood of (x, y) — rows y-1 .. y+2, one word per row, columns
    //                                 x-1 .. x+2 from the low byte up.  The middle 2 x 2 bytes are the intensities of the bilinear
    //                                 footprint (x, y) (x+1, y) (x, y+1) (x+1, y+1); each of the 8 bytes next to them is used by
    //                                 exactly one tap's gradient: 2 gx(x, y) = byte(x+1, y) - byte(x-1, y) and so on, which is
    //                                 Gradient.h's 0.5 * central difference when the bytes are the real neighbours — and the
    //                                 packer CHOOSES those 8 bytes so that the differences reproduce the given gradient image
    //                                 exactly (real neighbours in the interior, a copy of the opposite byte where Gradient.h
    //                                 writes the zero gradient of a border pixel).  A gradient image that no bytes reproduce
    //                                 (not a whole number of half grey levels, or out of a byte's reach) makes the level fall back
    //                                 to gathering ref_I / ref_dIxy directly.  The 4 corners are unused.
    //                                 — the whole footprint of a Hessian-pass sample (4 intensities, 4 gradients) in ONE
    //                                 128-bit load; the kernel converts a byte to 2^23 + b with one PRMT and takes differences
    //                                 (MBAVO_TEXEL = 0: the round-2 row pairs, 8 halves  gx gy (x,y) | gx gy (x+1,y) | I I | 0,
    //                                 two 128-bit loads per sample — kept for A/B runs, profiles/r2_history.md)
    //   quad texel (4 B per pixel):   bytes     I(x,y) I(x+1,y) I(x,y+1) I(x+1,y+1)
    //                                 — the whole footprint of a cost-only sample in one 32-bit load
    // x+1 / y+1 are clamped to the last column / row (those taps have weight 0 there).
#ifndef MBAVO_TEXEL
#define MBAVO_TEXEL 3 // 3: patch texel | 0: row pairs
#endif
    typedef uint4 PairTexel;
    struct LevelDev
    {
        const unsigned char *ref_I;
        const float2 *ref_dIxy;
        const PairTexel *ref_pair;    // nullptr: gradients the texels cannot hold -> the kernels gather ref_I / ref_dIxy directly
        const unsigned int *ref_quad;
        const unsigned char *cur_I[kMaxFrames];
        int H, W;
        double fx, fy, cx, cy;
        double inv_fx, inv_fy;
        const char *xy;     // records with two doubles at byte offset xy_offset, stride xy_stride
        int xy_stride, xy_offset;
        const double *z;
        int P, S, N, F;
        const int2 *pattern;
        const unsigned char *flags; // 1 = outlier
        double *patch_cost;         // [F * P * patch_cost_stride]
        int patch_cost_stride;
    };

    // ---- point sharding over the GPUs of one node: one-shot all-reduce through peer-mapped mailboxes ---------------
    // Every rank owns one Mailbox in its device memory; all ranks map all mailboxes (CUDA IPC across processes, plain peer
    // access inside a process) and WRITE their contribution into every rank's mailbox over NVLink: the last block of the
    // tracking kernel stores its packed vector into slot[parity][my_rank] of every peer as self-validating words, polls the
    // W slots of its OWN mailbox until every word carries the tag of this exchange and sums them in rank order (identical,
    // deterministic result on every rank).  Two parities suffice: a rank can only be one evaluation ahead of a peer (it
    // completes exchange k + 1 only after the peer has sent its k + 1 vector, i.e. after the peer finished reading k).
    constexpr int kMaxShards = 8;
    constexpr int kMailVec = (6 * kMaxKnotWindow + 1) * (6 * kMaxKnotWindow + 2) / 2; // packed_len(kMaxKnotWindow)
    // The packed vector travels as self-validating words (like TrackParams::host_out, and like NCCL's LL protocol): every
    // double is two 8-byte words (tag | low half) (tag | high half), tag = publish_tag(sequence number of the exchange), written
    // with one 16-byte store per element and polled by the receiver word by word — no system-scope fence between data and a
    // flag, no flag round trip: one NVLink traversal.  The small-vector exchanges of the outlier / keyframe statistics keep the
    // fence + flag form (they run once per accepted LM step).
    struct Mailbox
    {
        unsigned long long owner_uuid[2];           // UUID of the GPU the mailbox lives on (written at export, read by the peers at connect)
        unsigned long long aux_flag[2][kMaxShards]; // sequence number of the small vector in aux[parity][source rank]
        double aux[2][kMaxShards][8];
        ulonglong2 slot[2][kMaxShards][kMailVec];   // [parity][source rank][element] -> (tag | lo32, tag | hi32)
    };
    struct ShardParams
    {
        int world, rank;                // world <= 1: not sharded
        Mailbox *peer[kMaxShards];      // peer[rank] is this rank's own mailbox
        unsigned long long seq;         // sequence number of this exchange (same on every rank)
    };

#ifdef __CUDACC__
    // ---- system-scope accesses of the sharding mailboxes (peer memory over NVLink) ------------------------------
    __device__ __forceinline__ void st_sys(double *p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
    __device__ __forceinline__ double ld_sys(const double *p)
    {
        double v;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
    {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    }
    __device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
    {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p)
    {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ void st_release_gpu(unsigned int *p, unsigned int v)
    {
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
    __device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p)
    {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ unsigned long long global_timer_ns()
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        return t;
    }
    __device__ __forceinline__ void st_sys_v2(ulonglong2 *p, unsigned long long a, unsigned long long b)
    {
        asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
    }
    __device__ __forceinline__ ulonglong2 ld_sys_v2(const ulonglong2 *p)
    {
        ulonglong2 v;
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
        return v;
    }
    // spin until *flag == seq; gives up after ~4 s (a peer that never arrives must not hang the GPU)
    __device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long seq)
    {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(flag) != seq)
        {
            __nanosleep(64);
            if (global_timer_ns() - t0 > 4000000000ull)
                return false;
        }
        return true;
    }


    // All-reduce (sum) of n <= 8 doubles in shared memory `vals` over the ranks, executed by ALL threads of a one-block
    // kernel (blockDim >= world * n); on return vals holds the global sums on every rank.  false: a peer timed out.
    __device__ __forceinline__ bool shard_allreduce_small(const ShardParams &sh, unsigned long long seq, double *vals, int n)
    {
        __shared__ int ok_s;
        const int par = (int)(seq & 1ull), tid = threadIdx.x;
        if (tid == 0)
            ok_s = 1;
        if (tid < sh.world * n)
            st_sys(&sh.peer[tid / n]->aux[par][sh.rank][tid % n], vals[tid % n]);
        __threadfence_system();
        __syncthreads();
        if (tid < sh.world)
        {
            st_release_sys(&sh.peer[tid]->aux_flag[par][sh.rank], seq);
            if (!wait_flag(&sh.peer[sh.rank]->aux_flag[par][tid], seq))
                ok_s = 0;
        }
        __syncthreads();
        if (tid < n)
        {
            double s = 0.0;
            for (int r = 0; r < sh.world; ++r)
                s += ld_sys(&sh.peer[sh.rank]->aux[par][r][tid]);
            vals[tid] = s;
        }
        __syncthreads();
        return ok_s != 0;
    }
#endif

    // ---- device-resident Gauss-Newton sweep (mbavo_gn_sweep) -----------------------------------------------------------
    // The control knots of a sweep live in device memory: the Hessian-pass kernel's last block solves the damped normal
    // equations and forms the candidate knots, the candidate's pose kernel reads them from here, the cost-pass kernel's
    // last block records the candidate cost (and, when chaining, commits the candidate).  The host enqueues the kernels
    // of all levels at once and waits once.
    struct GnState
    {
        double cur_t[3 * 16], cur_R[4 * 16];   // knots the sweep currently stands on
        double cand_t[3 * 16], cand_R[4 * 16]; // candidate of the level in flight
        double step[6 * 16];                   // first LM step of the level in flight, full ordering [dt(3n), dw(3n)]
        double cost, cand_cost, model;
        int status;                            // 0 ok, 1 normal equations not safely positive definite, 2 model decrease < 0
        int cur_buf;                           // which of the two sample-record buffers holds the records of cur_t / cur_R
    };
    // Sample records of a sweep live in two buffers (A at the base pointers, B one stride further): one holds the records of
    // the knots the sweep stands on, the other those of the candidate, so that a level whose knots did not change (or became
    // the candidate) finds its records already there instead of running the pose kernel again.
    enum BufSelect : int
    {
        kBufA = 0,        // static: buffer A (plain evaluations; sweeps that never commit keep their knots in A)
        kBufCur = 1,      // dynamic: state->cur_buf
        kBufCand = 2,     // dynamic: 1 - state->cur_buf
        kBufB = 3         // static: buffer B
    };
    struct GnParams
    {
        GnState *state;     // nullptr: plain evaluation
        int mode;           // 1: Hessian pass -> solve + candidate;  2: cost pass at the candidate -> record / commit
        int n_knots, kmin;  // kmin: first knot of the packed window
        int chain;          // mode 2: commit the candidate when it lowered the cost
        int slot;           // mode 2: (cost, candidate cost, status, model) go to host_out[4 slot ..]; last != 0: knots follow
        int last;
        double radius;
    };

    struct TrackParams
    {
        LevelDev lv;
        const float *samples;     // [F * N * rec] sample records
        const double *mid;        // [F * kMidDoubles]
        const int *seg_end;       // [F * kMaxSegments]: one past the last sample index of every segment offset
        int buf_select;           // BufSelect: which record buffer this launch reads
        int samples_stride, mid_stride, seg_end_stride; // element offsets of buffer B
        double inv_num_residuals; // 1 / ((P - num_bad) F S)            spline_update_step.cpp:116-117
        float huber_a;
        int TP;                   // points per warp batch
        int PH;                   // exposure-sample phases per pixel: the 32 lanes of a warp are PH phases x 32/PH pixels
        int phase_fast;           // lane = slot * PH + phase instead of phase * (32/PH) + slot
        int batches_per_frame;
        double *block_partials;   // [gridDim.x * gridDim.y * E]
        unsigned int *counter;    // last-block-done ticket
        double *packed_out;       // [E] device memory
        // Blocking evaluations: the last block also stores the packed vector straight into mapped pinned host memory, every
        // element as two 8-byte words (tag | low half of the value bits) (tag | high half), tag = publish_tag(seq): EACH
        // 8-byte word validates itself (an aligned 8-byte access is single-copy atomic in the PTX memory model; a 16-byte
        // vector store is not guaranteed to be), so the host spins until both tags of all E elements show the evaluation's
        // tag — no system-scope fence, no separate flag, no D2H copy, no stream synchronisation, and no reliance on how
        // the 16-byte store crosses PCIe / C2C.
        double2 *host_out;                    // [E] word pairs in mapped pinned host memory, or nullptr
        unsigned long long seq;
        ShardParams shard;                    // point sharding: the vector published is the sum over all ranks
        GnParams gn;                          // device-resident Gauss-Newton sweep
        // persistent sweep after an asynchronous mbavo_set_frame: the level's points are valid once *ready_flag == ready_epoch (the
        // copy stream writes the flag right behind the points); nullptr: no wait
        const unsigned int *ready_flag;
        unsigned int ready_epoch;
        // mbavo_debug_dump (track_debug_kernel only): per-stage intermediates, device memory, each nullable
        double2 *dbg_centres;                 // [F * P] patch centre of every point (compute_local_patches_xy.cu:26-49)
        float *dbg_r;                         // [F * P * S] raw residual of every pixel
        float *dbg_J;                         // [F * P * S * 6 NK] raw Jacobian row of every pixel, ordered [t-block | w-block] of the window
        unsigned long long *phase_times;      // development: globaltimer stamps of kernel phases (MBAVO_PROFILE_PHASES builds),
        int trace_row;                        // 16 stamps per row; row = ordinal of the launch since the trace was armed
    };

    // ---- persistent sweep (one launch for the whole coarse-to-fine Gauss-Newton sweep) ---------------------------------------
    // sweep_kernel runs every pass of mbavo_gn_sweep — Hessian pass and cost pass of every level — inside ONE grid of one
    // block per SM.  The passes are separated by a ticket + flag barrier: every block announces its per-block partials with
    // the ticket of TrackParams::counter, the LAST block to arrive finishes the pass (sum of the partials, exchange, solve,
    // candidate knots AND the candidate's sample records; or record / commit) and then releases SweepCtl::done, which the
    // other blocks acquire before they read the records of the next pass.  `done` counts passes monotonically over the
    // life of the context, so nothing has to be reset between sweeps.
    struct SweepCtl
    {
        unsigned int done;  // passes completed so far
        int abort;          // a block gave up waiting (4 s): every block leaves the kernel
    };
    constexpr int kMaxSweepLevels = 8; // MBAVO_MAX_LEVELS of include/mbavo.h
    struct SweepParams
    {
        TrackParams pass[2 * kMaxSweepLevels]; // [2 li]: Hessian pass of level li (coarse first), [2 li + 1]: its cost pass
        int n_levels;
        SweepCtl *ctl;
        unsigned int base;                     // value of ctl->done when this sweep starts
        // globaltimer stamps (ns): [0] kernel entry (after the wait for the pose kernel), [1 + p] release of pass p — so that
        // pass p lasted stamps[1 + p] - stamps[p], everything included (wait, records, batches, reduction, solve, pose)
        unsigned long long *pass_times;
    };

    // ---- one pyramid step of mbavo_set_frame in ONE launch (track_common.cu) ----------------------------------------------------
    // Everything that reads the level-l images of a new frame: gradients + texels of the keyframe's level l (pack), the 2x2-box
    // level l + 1 of the keyframe and of every live image (down), and the reset of the level's outlier flags (zero) — jobs that
    // are independent of each other, so they share a launch (the host enqueues n_levels kernels instead of ~5 per level).
    struct PyrJob
    {
        int kind;                  // 0: pack, 1: down, 2: zero
        int block_begin;           // first block of the job in the launch (256 threads per block, one thread per output pixel / byte)
        const unsigned char *src;  // pack: the level image; down: the finer image
        int Hs, Ws;                // its size
        unsigned char *dst;        // down: the coarser image (Hd x Wd); zero: the bytes to clear
        int Hd, Wd;                // down: size of dst; zero: Wd = number of bytes
        PairTexel *pair;           // pack outputs (each nullable, see launch_pack_image_kernel)
        unsigned int *quad;
        float2 *grad;
    };
    constexpr int kMaxPyrJobs = 3 + kMaxFrames;
    struct PyrStepParams
    {
        PyrJob job[kMaxPyrJobs];
        int n_jobs;
    };

    // ---- semi-dense point selection (select_kernel.cu) -------------------------------------------------------------------
    constexpr int kSelectMaxLevels = 8; // MBAVO_MAX_LEVELS of include/mbavo.h
    struct SelectLevel
    {
        const unsigned char *I; // keyframe image of the level
        int H, W;
        int ch, cw;             // cell size at this level: (int)(cell / 1.414^level)   FeatureDetectorBase.cpp:60-61
        int nch, ncw;           // H / ch + 1, W / cw + 1                                :63-64
        int cell_base;          // first cell of the level in the concatenated cell list
        double2 *xy;            // selected points of the level, in cell order
        double *z;
    };
    struct SelectParams
    {
        SelectLevel lv[kSelectMaxLevels];
        int n_levels;
        float score_threshold;
        const float *depth_z;   // level-0 depth map, H0 x W0
        int W0;
        int capacity;           // points the xy / z arrays hold per level
        int4 *cell_rec;         // per cell: (x, y, bits of z, selected)
        int *count;             // [n_levels] points selected (may exceed capacity: the excess is not stored)
    };

    __host__ __device__ constexpr int packed_len(int NK) { return (6 * NK + 1) * (6 * NK + 2) / 2; }

    // ---- self-validating result words (TrackParams::host_out) ------------------------------------------------------------
    // 32-bit tag of a sequence number; never 0 (the buffer starts zeroed)
    __host__ __device__ inline unsigned long long publish_tag(unsigned long long seq) { return seq % 0xffffffffull + 1ull; }
#ifdef __CUDACC__
    __device__ __forceinline__ void publish_host(double2 *slot, double v, unsigned long long seq)
    {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v), tag = publish_tag(seq) << 32;
        // one 16-byte store instruction; correctness only needs each aligned 8-byte half to arrive whole
        *reinterpret_cast<ulonglong2 *>(slot) = make_ulonglong2(tag | (b & 0xffffffffull), tag | (b >> 32));
    }
#endif
    // host side: true when both words of `slot` carry the tag of `seq`; *v receives the value
    inline bool read_published(const volatile unsigned long long *slot, unsigned long long seq, double *v)
    {
        const unsigned long long w0 = slot[0], w1 = slot[1], tag = publish_tag(seq);
        if ((w0 >> 32) != tag || (w1 >> 32) != tag)
            return false;
        const unsigned long long b = (w0 & 0xffffffffull) | (w1 << 32);
        if (v)
            __builtin_memcpy(v, &b, sizeof b);
        return true;
    }
} // namespace mbavo

#endif
