// Drop-in replacement of the reference's include/ORBextractor.h (class ORB_SLAM2::ORBextractor,
// /root/reference/include/ORBextractor.h:45-111) over the C ABI of orb_b200.h. Same constructor,
// same operator(), same getters, same public mvImagePyramid, so Frame::ExtractORB
// (src/Frame.cc:247-253), Frame::ComputeStereoMatches (src/Frame.cc:473,563,580) and Tracking's ctor
// (src/Tracking.cc:119-125) compile unchanged against it. Header-only; link liborb_b200.so.
//
// Differences a maintainer should know:
//   * ExtractorNode (ORBextractor.h:32-43) is gone - the quadtree runs on the device;
//   * the protected helpers ComputePyramid / ComputeKeyPointsOctTree / DistributeOctTree are gone;
//   * one instance is bound to one CUDA device (ctor argument `device`, default 0) and, like the
//     reference, is not re-entrant; CUDA failures are fatal (std::runtime_error) - no CPU fallback.
#ifndef ORBSLAM2_B200_ORBEXTRACTOR_H
#define ORBSLAM2_B200_ORBEXTRACTOR_H

#include <cassert>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#if defined(__has_include)
#if __has_include(<opencv2/core/core.hpp>) && !defined(ORB_B200_NO_OPENCV)
#include <opencv2/core/core.hpp>
#else
#include "cv_compat.h"
#endif
#else
#include <opencv2/core/core.hpp>
#endif

#include "../orb_b200.h"

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures_, float scaleFactor_, int nlevels_, int iniThFAST_, int minThFAST_, int device = 0)
        : nfeatures(nfeatures_), scaleFactor(scaleFactor_), nlevels(nlevels_), iniThFAST(iniThFAST_), minThFAST(minThFAST_),
          device_(device) {
        // the scale tables do not depend on the image size: same float chain as the reference ctor (410-431)
        mvScaleFactor.assign(nlevels, 1.0f); mvLevelSigma2.assign(nlevels, 1.0f);
        for (int i = 1; i < nlevels; i++) {
            mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
        mvImagePyramid.resize(nlevels);
    }
    ~ORBextractor() { if (handle_) orbx_destroy(handle_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image. Mask is ignored (as in the reference).
    void operator()(cv::InputArray image_, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors_) {
        if (image_.empty()) return;  // src/ORBextractor.cc:1046-1047
        cv::Mat image = image_.getMat();
        assert(image.type() == CV_8UC1);
        ensure_handle(image.cols, image.rows);
        int n = 0;
        check(orbx_extract(handle_, image.data, (size_t)image.step, kps_.data(), desc_.data(), cap_, &n));
        if (n == 0) descriptors_.release();  // 1064-1065
        else {
            descriptors_.create(n, 32, CV_8U);  // 1068
            cv::Mat d = descriptors_.getMat();
            for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), &desc_[(size_t)i * 32], 32);
        }
        keypoints.clear();
        keypoints.reserve(n);
        for (int i = 0; i < n; ++i) {
            cv::KeyPoint k;
            k.pt.x = kps_[i].x; k.pt.y = kps_[i].y; k.size = kps_[i].size; k.angle = kps_[i].angle;
            k.response = kps_[i].response; k.octave = kps_[i].octave; k.class_id = -1;
            keypoints.push_back(k);
        }
        if (download_pyramid_) {
            // level 0 is the input itself (ComputePyramid copies it, 1107-1132): host copy; levels 1.. in one download
            std::vector<uint8_t*> dst(nlevels > 1 ? nlevels - 1 : 0);
            std::vector<size_t> strides(dst.size());
            for (int l = 0; l < nlevels; ++l) {
                int w = 0, h = 0;
                check(orbx_level_size(handle_, l, &w, &h));
                mvImagePyramid[l].create(h, w, CV_8UC1);
                if (l == 0) {
                    for (int y = 0; y < h; ++y) std::memcpy(mvImagePyramid[0].ptr(y), image.ptr(y), (size_t)w);
                } else {
                    dst[l - 1] = mvImagePyramid[l].data;
                    strides[l - 1] = (size_t)mvImagePyramid[l].step;
                }
            }
            if (nlevels > 1) check(orbx_pyramid_levels(handle_, 0, 1, nlevels - 1, dst.data(), strides.data()));
        }
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Host copy of the pyramid of the last frame (Frame::ComputeStereoMatches reads it). Filling it costs a
    // device->host copy of ~1.2 MB per 640x480 frame; monocular / RGB-D callers can switch it off.
    std::vector<cv::Mat> mvImagePyramid;
    void SetPyramidDownload(bool on) { download_pyramid_ = on; }
    orbx_handle NativeHandle() { return handle_; }  // device-resident pyramid / results via orb_b200.h

protected:
    void check(int rc) {
        if (rc != ORB_OK && rc != ORB_ECAPACITY) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error());
    }
    void ensure_handle(int w, int h) {
        if (handle_ && w == width_ && h == height_) return;
        if (handle_) { orbx_destroy(handle_); handle_ = nullptr; }
        orbx_config cfg{nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST};
        check(orbx_create(&cfg, device_, w, h, 1, &handle_));
        width_ = w; height_ = h;
        cap_ = orbx_max_keypoints(handle_);
        kps_.resize(cap_); desc_.resize((size_t)cap_ * 32);
    }

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

    int device_ = 0, width_ = 0, height_ = 0, cap_ = 0;
    bool download_pyramid_ = true;
    orbx_handle handle_ = nullptr;
    std::vector<orbx_keypoint> kps_;
    std::vector<uint8_t> desc_;
};

}  // namespace ORB_SLAM2

#endif
