// BlurAwareDirectTracker::trackFrame behind the C-ABI: what the reference does per blurred frame
// around the hot path (src/ba_tracker/blur_aware_direct_tracker.cpp:88-203), composed from the library's own entry points —
// nothing here computes on the CPU what the device path offers.
//   first frame / spline initialisation     :91-109    mbavo_tracker_init
//   blurred-frame pyramid + upload          :112-117   mbavo_set_live_pyramid
//   constant-velocity prediction            :119-145   mbavo_predict_spline
//   optimizeTrajectory (coarse -> fine)     :544-588   mbavo_optimize_level per level
//   isKeyframe statistics                   :205-248   mbavo_keyframe_stats
//   velocity from the neighbouring frames   :155-162   mbavo_frame_velocity
//   re-anchoring on a new keyframe          :186-199   mbavo_tracker_new_keyframe
//   keyframe decision                       :251-262   mbavo_is_keyframe
// The new keyframe's image / depth / point selection
// (mbavo_set_keyframe_pyramid, mbavo_select_points) stay with the caller, as in tmpProcessKeyframe.
#include "../../include/mbavo.h"

#include <cstring>

namespace
{
    void compose(const double *ta, const double *qa, const double *tb, const double *qb, double *t, double *q)
    {
        // Transformation a * b: rotation qa qb, translation qa tb + ta
        const double ux = 2.0 * (qa[1] * tb[2] - qa[2] * tb[1]), uy = 2.0 * (qa[2] * tb[0] - qa[0] * tb[2]), uz = 2.0 * (qa[0] * tb[1] - qa[1] * tb[0]);
        const double r[3] = {tb[0] + qa[3] * ux + (qa[1] * uz - qa[2] * uy), tb[1] + qa[3] * uy + (qa[2] * ux - qa[0] * uz),
                             tb[2] + qa[3] * uz + (qa[0] * uy - qa[1] * ux)};
        const double x = qa[3] * qb[0] + qa[0] * qb[3] + qa[1] * qb[2] - qa[2] * qb[1];
        const double y = qa[3] * qb[1] + qa[1] * qb[3] + qa[2] * qb[0] - qa[0] * qb[2];
        const double z = qa[3] * qb[2] + qa[2] * qb[3] + qa[0] * qb[1] - qa[1] * qb[0];
        const double w = qa[3] * qb[3] - qa[0] * qb[0] - qa[1] * qb[1] - qa[2] * qb[2];
        t[0] = r[0] + ta[0], t[1] = r[1] + ta[1], t[2] = r[2] + ta[2];
        q[0] = x, q[1] = y, q[2] = z, q[3] = w;
    }

    mbavo_spline spline_of(const mbavo_tracker *tr)
    {
        mbavo_spline sp;
        sp.spline_deg_k = tr->spline_deg_k, sp.start_time = tr->start_time, sp.sample_dt = tr->sample_dt;
        sp.num_ctrl_knots = tr->num_ctrl_knots, sp.knots_t = tr->knots_t, sp.knots_R = tr->knots_R;
        return sp;
    }
} // namespace

extern "C"
{
    int mbavo_tracker_init(mbavo_tracker *tr, int spline_deg_k, double sample_dt, double keyframe_capture_time)
    {
        if (!tr || spline_deg_k != 2 || !(sample_dt > 0)) // the tracker inserts two knots (:103-104): only k = 2 is consistent
            return MBAVO_EINVAL;
        std::memset(tr, 0, sizeof(*tr));
        tr->spline_deg_k = spline_deg_k, tr->num_ctrl_knots = 2;
        tr->sample_dt = sample_dt, tr->start_time = keyframe_capture_time;
        for (int i = 0; i < 2; ++i)
            tr->knots_R[4 * i + 3] = 1.0;
        tr->prev_q[3] = 1.0, tr->keyframe_q[3] = 1.0;
        tr->prev_timestamp = keyframe_capture_time;
        return MBAVO_OK;
    }

    int mbavo_track_frame(mbavo_ctx *ctx, mbavo_tracker *tr, int n_levels, int mem, const unsigned char *cur_I0, double capture_time,
                          double exposure_time, const mbavo_lm_options *opt, mbavo_frame_result *res)
    {
        if (!ctx || !tr || !cur_I0 || !opt || !res || n_levels < 1 || n_levels > MBAVO_MAX_LEVELS)
            return MBAVO_EINVAL;
        std::memset(res, 0, sizeof(*res));
        const double dt_frame = capture_time - tr->prev_timestamp; // :120
        if (!(dt_frame > 0) || !(exposure_time > 0))
            return MBAVO_EINVAL;
        // Everything fallible runs on a COPY of the tracker state; *tr is committed only when the whole frame has succeeded, so a
        // failed call (time-out of a peer rank, exposure outside the spline, a CUDA error) can be retried without the
        // constant-velocity prediction being applied twice.
        mbavo_tracker w = *tr;
        const unsigned char *frames[1] = {cur_I0};
        int rc = mbavo_set_live_pyramid(ctx, n_levels, mem, frames, 1); // :112-117
        if (rc != MBAVO_OK)
            return rc;
        w.start_time = capture_time - 0.5 * exposure_time;       // :144
        rc = mbavo_predict_spline(w.num_ctrl_knots, w.knots_t, w.knots_R, w.velocity, dt_frame); // :122-145
        if (rc != MBAVO_OK)
            return rc;
        rc = mbavo_set_frame_times(ctx, 1, &capture_time, &exposure_time); // uploadDataToGpu, :701-719
        if (rc != MBAVO_OK)
            return rc;
        mbavo_frame_result out;
        std::memset(&out, 0, sizeof out);
        for (int level = n_levels - 1; level >= 0; --level) // optimizeTrajectory, :571-575
        {
            rc = mbavo_optimize_level(ctx, level, w.spline_deg_k, w.start_time, w.sample_dt, w.num_ctrl_knots, w.knots_t, w.knots_R, opt,
                                      &out.levels[level]);
            if (rc == MBAVO_ENOTREADY && level > 0)
                continue; // a coarse level the selection left without points
            if (rc != MBAVO_OK)
                return rc;
            out.levels_run |= 1 << level;
        }
        const mbavo_spline sp = spline_of(&w);
        // isKeyframe statistics (:205-248): poses at the capture time and at -/+ half the exposure
        double poses[21];
        const double times[3] = {capture_time, capture_time - 0.5 * exposure_time, capture_time + 0.5 * exposure_time};
        for (int i = 0; i < 3; ++i)
        {
            rc = mbavo_spline_pose(&sp, times[i], poses + 7 * i, poses + 7 * i + 3);
            if (rc != MBAVO_OK)
                return rc;
        }
        rc = mbavo_keyframe_stats(ctx, 0, poses, &out.avg_flow, &out.avg_kernel_len);
        if (rc != MBAVO_OK)
            return rc;
        // velocity from the neighbouring frames (:155-162); poses[0..6] is the pose at the capture time
        rc = mbavo_frame_velocity(w.prev_t, w.prev_q, poses, poses + 3, dt_frame, w.velocity);
        if (rc != MBAVO_OK)
            return rc;
        std::memcpy(w.prev_t, poses, sizeof(double) * 3);
        std::memcpy(w.prev_q, poses + 3, sizeof(double) * 4);
        w.prev_timestamp = capture_time; // :200
        std::memcpy(out.t_cur2key, poses, sizeof(double) * 3);
        std::memcpy(out.q_cur2key, poses + 3, sizeof(double) * 4);
        compose(w.keyframe_t, w.keyframe_q, poses, poses + 3, out.t_cur2world, out.q_cur2world); // :203
        *tr = w;
        *res = out;
        return MBAVO_OK;
    }

    int mbavo_is_keyframe(double avg_flow, double avg_kernel_len, double max_flow_mag0, double max_flow_mag1, double max_blur_kernel_mag)
    {
        // blur_aware_direct_tracker.cpp:251-262: far enough from the keyframe and sharp enough, or simply too far
        if (avg_flow > max_flow_mag0 && avg_kernel_len < max_blur_kernel_mag)
            return 1;
        if (avg_flow > max_flow_mag1)
            return 1;
        return 0;
    }

    int mbavo_tracker_new_keyframe(mbavo_tracker *tr, double capture_time)
    {
        if (!tr)
            return MBAVO_EINVAL;
        const mbavo_spline sp = spline_of(tr);
        double t[3], q[4];
        int rc = mbavo_spline_pose(&sp, capture_time, t, q); // :191
        if (rc != MBAVO_OK)
            return rc;
        double kt[3], kq[4];
        compose(tr->keyframe_t, tr->keyframe_q, t, q, kt, kq); // :192
        const double zero[3] = {0, 0, 0}, ident[4] = {0, 0, 0, 1};
        double nt[3 * MBAVO_MAX_TRACKER_KNOTS], nR[4 * MBAVO_MAX_TRACKER_KNOTS];
        rc = mbavo_spline_transform_to(&sp, capture_time, zero, ident, nt, nR); // :194-196
        if (rc != MBAVO_OK)
            return rc;
        std::memcpy(tr->knots_t, nt, sizeof(double) * 3 * tr->num_ctrl_knots);
        std::memcpy(tr->knots_R, nR, sizeof(double) * 4 * tr->num_ctrl_knots);
        std::memcpy(tr->keyframe_t, kt, sizeof(kt));
        std::memcpy(tr->keyframe_q, kq, sizeof(kq));
        tr->prev_t[0] = tr->prev_t[1] = tr->prev_t[2] = 0.0; // mTprevB2W = Transformation() (:197)
        tr->prev_q[0] = tr->prev_q[1] = tr->prev_q[2] = 0.0, tr->prev_q[3] = 1.0;
        return MBAVO_OK;
    }
}
