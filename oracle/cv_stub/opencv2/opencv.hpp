/* TEST INFRASTRUCTURE.  Stand-in for the OpenCV headers, which this image does not have: just enough of cv::Point2f,
 * cv::KeyPoint and cv::Mat for the reference's FeatureDetectorBase.{h,cpp} / FeatureDetectorSemiDense.{h,cpp} to compile
 * UNMODIFIED for oracle/_ref (the detector only stores pt.x, pt.y and response; OpenCV's KeyPoint value-initialises them
 * to 0, which gridSelection relies on).  Never part of the product library. */
#ifndef MBAVO_ORACLE_CV_STUB_H
#define MBAVO_ORACLE_CV_STUB_H
namespace cv
{
    struct Point2f
    {
        float x = 0.f, y = 0.f;
    };
    struct KeyPoint
    {
        Point2f pt;
        float size = 0.f, angle = -1.f, response = 0.f;
        int octave = 0, class_id = -1;
    };
    struct Mat
    {
    };
} // namespace cv
#endif
