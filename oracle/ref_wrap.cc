// TEST INFRASTRUCTURE ONLY (oracle). C entry points around the reference's own ORBextractor
// (compiled unmodified from /root/reference/src/ORBextractor.cc against oracle/shim by
// oracle/build_ref.sh). Used to validate orb_oracle.cc and as the "reference" CPU baseline.
#include <cstdint>
#include <cstring>
#include <vector>

#include "ORBextractor.h"  // the reference header, found through -I/root/reference/include

extern "C" {

void* ref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new ORB_SLAM2::ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_extractor_destroy(void* h) { delete (ORB_SLAM2::ORBextractor*)h; }

// runs operator(); returns the keypoint count and writes up to cap records
// (x, y, size, angle, response, octave) + 32-byte descriptors
int ref_extract(void* h, const uint8_t* img, int w, int ht, size_t stride, float* kps6, uint8_t* desc32, int cap) {
    ORB_SLAM2::ORBextractor& ex = *(ORB_SLAM2::ORBextractor*)h;
    cv::Mat image(ht, w, CV_8UC1, (void*)img, stride);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
    ex(image, cv::Mat(), kps, desc);
    const int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; ++i) {
        const cv::KeyPoint& k = kps[i];
        const float rec[6] = {k.pt.x, k.pt.y, k.size, k.angle, k.response, (float)k.octave};
        if (kps6) std::memcpy(kps6 + 6 * i, rec, sizeof(rec));
        if (desc32) std::memcpy(desc32 + 32 * i, desc.ptr(i), 32);
    }
    return n;
}

void ref_get_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2) {
    ORB_SLAM2::ORBextractor& ex = *(ORB_SLAM2::ORBextractor*)h;
    std::vector<float> a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(),
                       c = ex.GetScaleSigmaSquares(), d = ex.GetInverseScaleSigmaSquares();
    for (int i = 0; i < ex.GetLevels(); ++i) { scale[i] = a[i]; invScale[i] = b[i]; sigma2[i] = c[i]; invSigma2[i] = d[i]; }
}

void ref_level_dims(void* h, int l, int* w, int* ht) {
    ORB_SLAM2::ORBextractor& ex = *(ORB_SLAM2::ORBextractor*)h;
    *w = ex.mvImagePyramid[l].cols; *ht = ex.mvImagePyramid[l].rows;
}
// copies pyramid level l including `border` pixels of the frame around it (0..19)
void ref_level_image(void* h, int l, uint8_t* out, int border) {
    ORB_SLAM2::ORBextractor& ex = *(ORB_SLAM2::ORBextractor*)h;
    const cv::Mat& m = ex.mvImagePyramid[l];
    const int W = m.cols + 2 * border;
    for (int y = -border; y < m.rows + border; ++y)
        std::memcpy(out + (size_t)(y + border) * W, m.data + (ptrdiff_t)y * (ptrdiff_t)m.step - border, W);
}

}  // extern "C"
