// TEST INFRASTRUCTURE ONLY (oracle). Force-included (-include) when the reference's matcher and map data model are
// compiled unmodified for oracle/_ref/libref_slam.so (oracle/build_ref.sh):
//  * the standard headers come first, then `private` / `protected` are turned into `public` for everything that follows -
//    i.e. for the reference's own headers - so that oracle/ref_slam_wrap.cc can fill Frame / KeyFrame / MapPoint
//    members from arrays and call Frame::AssignFeaturesToGrid & co. directly. Access control does not change code
//    generation; no reference line is altered.
//  * include/Converter.h pulls Eigen and g2o (absent here); its include guard is pre-defined and the one member the
//    compiled files use (toDescriptorVector, src/Frame.cc:399, src/KeyFrame.cc:78: one row header per descriptor) is
//    supplied instead.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <limits.h>
#include <list>
#include <map>
#include <mutex>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "cvshim.h"

#define CONVERTER_H
namespace ORB_SLAM2 {
class Converter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& D) {
        std::vector<cv::Mat> v;
        v.reserve(D.rows);
        for (int j = 0; j < D.rows; ++j) v.push_back(D.row(j));
        return v;
    }
};
}  // namespace ORB_SLAM2

#define private public
#define protected public
