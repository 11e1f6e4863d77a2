// TEST INFRASTRUCTURE — never linked into the product library.
//
// Drives the reference's own semi-dense point selection on the CPU: ImagePyramid<T>::computePyramid
// (src/core/measurements/ImagePyramid.h:59-99), compute_image_gradients with the magnitude image
// (src/core/image_proc/Gradient.h:17-75), FeatureDetectorSemiDense::detect
// (src/core/feature_detectors/FeatureDetectorSemiDense.cpp:16-59) and FeatureDetectorBase::gridSelection
// (FeatureDetectorBase.cpp:49-92), all compiled UNMODIFIED from /root/reference (the two .cpp files are handed to the
// compiler by oracle/Makefile; OpenCV's KeyPoint comes from oracle/cv_stub).  The depth look-up of
// BlurAwareDirectTracker::tmpProcessKeyframe (blur_aware_direct_tracker.cpp:389-409) sits in a translation unit that needs
// Eigen/OpenCV proper, so those few lines are restated at the end of this function.
#include "core/feature_detectors/FeatureDetectorSemiDense.h"
#include "core/image_proc/Gradient.h"
#include "core/measurements/ImagePyramid.h"

#include <cmath>
#include <vector>

using namespace SLAM::Core;

extern "C"
{
    // The reference's own pyramid and gradient loops (ImagePyramid<T>::computePyramid, ImagePyramid.h:59-99;
    // compute_image_gradients, Gradient.h:17-75): levels[] receives the n_levels images back to back (H0*W0, then
    // (H0/2)*(W0/2), ...), grads[] the interleaved (dx, dy) float images in the same order.
    int mbavo_refselect_pyramid(const unsigned char *I0, int H0, int W0, int n_levels, unsigned char *levels, float *grads)
    {
        Image<unsigned char> img(H0, W0, 1);
        img.copyFrom(const_cast<unsigned char *>(I0), H0, W0, 1);
        ImagePyramid<unsigned char> pyr;
        pyr.setNumOfPyramidLevels(n_levels);
        pyr.computePyramid(&img);
        size_t off = 0;
        for (int lv = 0; lv < n_levels; ++lv)
        {
            Image<unsigned char> *im = pyr.getImagePtr(lv);
            const int H = im->nHeight(), W = im->nWidth();
            Image<float> grad(H, W, 2);
            compute_image_gradients<unsigned char, float>(im, &grad);
            for (int i = 0; i < H * W; ++i)
            {
                levels[off + i] = im->getData()[i];
                grads[2 * (off + i)] = grad.getData()[2 * i];
                grads[2 * (off + i) + 1] = grad.getData()[2 * i + 1];
            }
            off += (size_t)H * W;
        }
        return 0;
    }

    // xy: n_levels x max_points x 2 doubles, z: n_levels x max_points doubles, count: n_levels ints (the number selected,
    // even where it exceeds max_points; only the first max_points are stored)
    int mbavo_refselect_points(const unsigned char *I0, int H0, int W0, int n_levels, float score_threshold, int cell_H, int cell_W,
                               const float *depth_z, int max_points, double *xy, double *z, int *count, float *mag_out /* nullable: level 0 */)
    {
        Image<unsigned char> img(H0, W0, 1);
        img.copyFrom(const_cast<unsigned char *>(I0), H0, W0, 1);
        ImagePyramid<unsigned char> pyr;
        pyr.setNumOfPyramidLevels(n_levels);
        pyr.computePyramid(&img);
        ImagePyramid<float> magPyr; // Frame::computeGradImagePyramid (src/core/measurements/Frame.cpp:125-152)
        magPyr.setNumOfPyramidLevels(n_levels);
        for (int lv = 0; lv < n_levels; ++lv)
        {
            Image<unsigned char> *im = pyr.getImagePtr(lv);
            const int H = im->nHeight(), W = im->nWidth();
            Image<float> grad(H, W, 2), mag(H, W, 1);
            compute_image_gradients<unsigned char, float>(im, &grad, &mag);
            magPyr.copyPyramidFrom(lv, mag.getData(), H, W, 1);
            if (lv == 0 && mag_out)
                for (int i = 0; i < H * W; ++i)
                    mag_out[i] = mag.getData()[i];
        }
        FeatureDetectorOptions opt; // blur_aware_direct_tracker.cpp:355-359
        opt.detector_type = ENUM_SEMIDENSE;
        opt.score_threshold = score_threshold;
        opt.grid_selection_cell_H = cell_H;
        opt.grid_selection_cell_W = cell_W;
        FeatureDetectorSemiDense det(opt);
        det.detect(&magPyr);
        for (int lv = 0; lv < n_levels; ++lv)
        {
            // blur_aware_direct_tracker.cpp:389-409
            double scale = pow(2, lv);
            std::vector<cv::KeyPoint> &kps = det.getFeaturePoints(lv);
            int n = 0;
            for (auto &kpt : kps)
            {
                int x = kpt.pt.x * scale + 0.5;
                int y = kpt.pt.y * scale + 0.5;
                float zz = depth_z[(size_t)y * W0 + x];
                if (zz < 1e-2)
                    continue;
                if (n < max_points)
                {
                    xy[((size_t)lv * max_points + n) * 2 + 0] = kpt.pt.x;
                    xy[((size_t)lv * max_points + n) * 2 + 1] = kpt.pt.y;
                    z[(size_t)lv * max_points + n] = zz;
                }
                ++n;
            }
            count[lv] = n;
        }
        return 0;
    }
}
