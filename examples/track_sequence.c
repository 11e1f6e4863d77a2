/* A tracker loop on the C-ABI, in plain C: what a caller of BlurAwareDirectTracker::trackFrame does
 * (src/ba_tracker/blur_aware_direct_tracker.cpp:88-203), on a synthetic sequence.
 *
 *   keyframe        mbavo_set_keyframe_pyramid + mbavo_select_points   (tmpProcessKeyframe, :343-410)
 *   per frame       mbavo_track_frame                                  (:112-162, 200-203)
 *   keyframe test   mbavo_is_keyframe                                  (:251-262)
 *   new keyframe    mbavo_tracker_new_keyframe + the two calls above   (:183-199)
 *
 * The scene is a textured plane at depth 7.5 in front of the first camera; the camera moves with a constant twist; every frame
 * is rendered with mbavo_synthesize_blurred from the first keyframe (32 poses over the exposure).  Prints the pose error of
 * every frame against the trajectory that rendered it and exits non-zero if one exceeds 2e-3 of the scene depth / 1.5e-3 rad.
 *
 *   gcc -O2 -std=c99 examples/track_sequence.c -Iinclude -Lmba-vo_b200/lib -lmbavo_b200 -lm -Wl,-rpath,$PWD/mba-vo_b200/lib -o track_sequence
 */
#include "mbavo.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define H 240
#define W 320
#define LEVELS 3
#define DEPTH 7.5
#define CHECK(call)                                                                    \
    do                                                                                 \
    {                                                                                  \
        int rc_ = (call);                                                              \
        if (rc_ != MBAVO_OK)                                                           \
        {                                                                              \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, mbavo_last_error());         \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

static const double kFx = 160.0, kFy = 160.0, kCx = 160.0, kCy = 120.0;
static const double kTwist[6] = {1.6, -0.9, 0.5, 0.05, -0.12, 0.2}; /* per second: [translation, rotation] */
static const double kFrameDt = 0.1, kExposure = 0.06;
/* the residual pattern of the reference's test (test/test_blur_aware_tracker_modules.cpp:662-679) */
static const int kPattern[16] = {-2, -2, 2, -2, -1, -1, 1, -1, 0, 0, 0, 1, -2, 2, 2, 2};

static void gt_pose(double t, double *pose7) /* T_cur2key(t) = Exp(t * twist) */
{
    double tg[6];
    for (int i = 0; i < 6; ++i)
        tg[i] = kTwist[i] * t;
    mbavo_se3_exp(tg, pose7, pose7 + 3);
}

static int render(const unsigned char *key, double capture, int num_poses, unsigned char *out)
{
    double poses[32 * 7];
    for (int j = 0; j < num_poses; ++j)
        gt_pose(num_poses == 1 ? capture : capture - 0.5 * kExposure + j * kExposure / (num_poses - 1), poses + 7 * j);
    return mbavo_synthesize_blurred(-1, MBAVO_MEM_HOST, key, H, W, DEPTH, kFx, kFy, kCx, kCy, poses, num_poses, out);
}

static int new_keyframe(mbavo_ctx *ctx, const unsigned char *image, const float *depth, int *counts)
{
    mbavo_point_selection sel;
    CHECK(mbavo_set_keyframe_pyramid(ctx, LEVELS, MBAVO_MEM_HOST, image, H, W));
    sel.score_threshold = 4.0f, sel.cell_H = 6, sel.cell_W = 6;
    sel.depth_mem = MBAVO_MEM_HOST, sel.depth_z = depth;
    sel.fx = kFx, sel.fy = kFy, sel.cx = kCx, sel.cy = kCy;
    sel.pattern_xy = kPattern, sel.patch_size = 8, sel.num_virtual_poses = 16;
    CHECK(mbavo_select_points(ctx, LEVELS, &sel, counts));
    return 0;
}

int main(void)
{
    static unsigned char key[H * W], frame[H * W];
    static float depth[H * W];
    /* band-limited texture: a few sinusoids */
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            double v = 0.0;
            for (int k = 1; k <= 12; ++k)
                v += sin(0.11 * k * x * cos(0.7 * k) + 0.13 * k * y * sin(0.7 * k) + 1.3 * k) / k;
            key[y * W + x] = (unsigned char)(128.0 + 40.0 * v < 0 ? 0 : 128.0 + 40.0 * v > 255 ? 255 : 128.0 + 40.0 * v);
            depth[y * W + x] = (float)DEPTH;
        }

    mbavo_limits lim = {-1, 1, 16, 8192, 8, 2};
    mbavo_ctx *ctx = NULL;
    CHECK(mbavo_create(&lim, &ctx));
    int counts[LEVELS];
    if (new_keyframe(ctx, key, depth, counts))
        return 1;
    printf("keyframe 0: %d / %d / %d points selected on the GPU\n", counts[0], counts[1], counts[2]);

    mbavo_tracker trk;
    mbavo_lm_options opt;
    mbavo_lm_default_options(&opt);
    opt.huber_a = 10.0, opt.max_chi_square_error = 3.0;
    CHECK(mbavo_tracker_init(&trk, 2, kFrameDt, 0.0));

    int bad = 0;
    for (int i = 1; i <= 6; ++i)
    {
        const double cap = i * kFrameDt;
        mbavo_frame_result res;
        double want[7];
        CHECK(render(key, cap, 32, frame));
        CHECK(mbavo_track_frame(ctx, &trk, LEVELS, MBAVO_MEM_HOST, frame, cap, kExposure, &opt, &res));
        gt_pose(cap, want);
        double et = 0.0, dot = 0.0;
        for (int a = 0; a < 3; ++a)
            et += (res.t_cur2world[a] - want[a]) * (res.t_cur2world[a] - want[a]);
        for (int a = 0; a < 4; ++a)
            dot += res.q_cur2world[a] * want[3 + a];
        const double er = 2.0 * acos(fabs(dot) > 1.0 ? 1.0 : fabs(dot));
        const int kf = mbavo_is_keyframe(res.avg_flow, res.avg_kernel_len, 12.0, 40.0, 6.0);
        printf("frame %d: |dt| %.2e (%.1e of the depth)  rotation error %.2e rad  flow %.1f px  blur %.1f px  LM iterations %d/%d/%d%s\n", i,
               sqrt(et), sqrt(et) / DEPTH, er, res.avg_flow, res.avg_kernel_len, res.levels[2].num_iterations, res.levels[1].num_iterations,
               res.levels[0].num_iterations, kf ? "  -> new keyframe" : "");
        if (sqrt(et) > 2e-3 * DEPTH || er > 1.5e-3)
            bad = 1;
        if (kf)
        {
            /* the sharp frame at the capture time and the depth of the same plane seen from there: e_z . (R z r + t) = DEPTH */
            double q[4] = {want[3], want[4], want[5], want[6]};
            const double r20 = 2 * (q[0] * q[2] - q[3] * q[1]), r21 = 2 * (q[1] * q[2] + q[3] * q[0]), r22 = 1 - 2 * (q[0] * q[0] + q[1] * q[1]);
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x)
                    depth[y * W + x] = (float)((DEPTH - want[2]) / (r20 * (x - kCx) / kFx + r21 * (y - kCy) / kFy + r22));
            CHECK(render(key, cap, 1, frame));
            CHECK(mbavo_tracker_new_keyframe(&trk, cap));
            if (new_keyframe(ctx, frame, depth, counts))
                return 1;
            printf("keyframe at frame %d: %d / %d / %d points\n", i, counts[0], counts[1], counts[2]);
        }
    }
    printf("kernels launched: %lld\n", mbavo_kernel_launches(ctx));
    mbavo_destroy(ctx);
    return bad;
}
