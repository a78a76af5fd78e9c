// Latency of the drop-in's real operating point (src/Frame.cc:247-253 calls the extractor on ONE image per frame): the C++
// facade's ORB_SLAM2::ORBextractor::operator() at batch 1 from host memory, with and without the host copy of the pyramid
// (mvImagePyramid, needed by Frame::ComputeStereoMatches only), and ORBmatcher::SearchForInitialization on a frame pair.
// Wall clock (steady_clock) around the calls, which block until the results are in host memory - what a SLAM caller sees.
//   single_frame <w> <h> <frames.raw (n frames)> <n> <nfeatures> <reps>  -> one JSON object on stdout
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "orbslam2_b200/FrameGrid.h"
#include "orbslam2_b200/ORBextractor.h"
#include "orbslam2_b200/ORBmatcher.h"

struct FrameLite {   // the members ORBmatcher::SearchForInitialization touches on the reference's Frame (include/Frame.h)
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;   // undistorted image bounds, static like Frame's (190-193): the facade gates on the device
    ORB_SLAM2::FrameGrid<cv::KeyPoint> grid;
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, int minLevel, int maxLevel) const {
        return grid.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    }
};
float FrameLite::mnMinX = 0.f, FrameLite::mnMaxX = 0.f, FrameLite::mnMinY = 0.f, FrameLite::mnMaxY = 0.f;
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static double median(std::vector<double> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const int w = atoi(argv[1]), h = atoi(argv[2]), n = atoi(argv[4]), nf = atoi(argv[5]), reps = atoi(argv[6]);
    std::vector<unsigned char> buf((size_t)w * h * n);
    FILE* f = fopen(argv[3], "rb");
    if (!f || fread(buf.data(), 1, buf.size(), f) != buf.size()) { fprintf(stderr, "cannot read %s\n", argv[3]); return 2; }
    fclose(f);
    ORB_SLAM2::ORBextractor ex(nf, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
    double t_pyr = 0, t_nopyr = 0;
    size_t nk = 0;
    for (int mode = 0; mode < 2; ++mode) {
        ex.SetPyramidDownload(mode == 0);
        std::vector<double> t;
        for (int r = 0; r < reps + 5; ++r) {
            cv::Mat im(h, w, CV_8UC1, buf.data() + (size_t)(r % n) * w * h, (size_t)w);
            const double t0 = now_us();
            ex(im, cv::Mat(), kps, desc);
            if (r >= 5) t.push_back(now_us() - t0);
            nk = kps.size();
        }
        (mode == 0 ? t_pyr : t_nopyr) = median(t);
    }
    // SearchForInitialization between frames 0 and 1 (mono initialisation uses 2 * nFeatures, src/Tracking.cc:124-125)
    ORB_SLAM2::ORBextractor ini(2 * nf, 1.2f, 8, 20, 7);
    ini.SetPyramidDownload(false);
    FrameLite F[2];
    FrameLite::mnMinX = 0.f; FrameLite::mnMinY = 0.f; FrameLite::mnMaxX = (float)w; FrameLite::mnMaxY = (float)h;
    for (int k = 0; k < 2; ++k) {
        cv::Mat im(h, w, CV_8UC1, buf.data() + (size_t)(k % n) * w * h, (size_t)w);
        ini(im, cv::Mat(), F[k].mvKeysUn, F[k].mDescriptors);
        F[k].grid.SetBounds(0.f, 0.f, (float)w, (float)h);
        F[k].grid.Assign(F[k].mvKeysUn);
    }
    ORB_SLAM2::ORBmatcher matcher(0.9f, true);
    std::vector<double> tm;
    int nm = 0;
    for (int r = 0; r < reps / 4 + 5; ++r) {
        std::vector<cv::Point2f> prev(F[0].mvKeysUn.size());
        for (size_t i = 0; i < prev.size(); ++i) prev[i] = F[0].mvKeysUn[i].pt;
        std::vector<int> m12;
        const double t0 = now_us();
        nm = matcher.SearchForInitialization(F[0], F[1], prev, m12, 100);
        if (r >= 5) tm.push_back(now_us() - t0);
    }
    printf("{\"extract_us_with_pyramid_download\": %.1f, \"extract_us\": %.1f, \"keypoints\": %zu, \"search_for_initialization_us\": %.1f, "
           "\"init_matches\": %d, \"init_keypoints\": %zu, \"reps\": %d}\n",
           t_pyr, t_nopyr, nk, median(tm), nm, F[0].mvKeysUn.size(), reps);
    return 0;
}
