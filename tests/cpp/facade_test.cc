// Test driver for the C++ facade (include/orbslam2_b200): extracts two raw 8-bit images with
// ORB_SLAM2::ORBextractor, runs ORBmatcher::SearchForInitialization on a Frame stand-in that has the
// reference Frame's members, and dumps everything for tests/test_gpu_facade.py to compare with the oracle.
//   facade_test <w> <h> <imgA.raw> <imgB.raw> <out.bin> [nfeatures]
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "orbslam2_b200/FrameGrid.h"
#include "orbslam2_b200/ORBextractor.h"
#include "orbslam2_b200/ORBmatcher.h"

struct FrameLite {  // the members ORBmatcher::SearchForInitialization touches on the reference's Frame
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    ORB_SLAM2::FrameGrid<cv::KeyPoint> grid;
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, int minLevel, int maxLevel) const {
        return grid.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    }
};

struct MapPointLite { bool bad; bool isBad() const { return bad; } };
struct KeyFrameLite {  // the members ORBmatcher::SearchByBoW(KF,KF) touches on the reference's KeyFrame
    std::vector<cv::KeyPoint> mvKeysUn;
    std::map<unsigned, std::vector<unsigned> > mFeatVec;  // DBoW2::FeatureVector is such a map
    std::vector<MapPointLite*> mps;
    cv::Mat mDescriptors;
    std::vector<MapPointLite*> GetMapPointMatches() { return mps; }
};

static std::vector<unsigned char> read_all(const char* path, size_t n) {
    std::vector<unsigned char> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const int w = atoi(argv[1]), h = atoi(argv[2]);
    const int nf = argc > 6 ? atoi(argv[6]) : 1000;
    std::vector<unsigned char> a = read_all(argv[3], (size_t)w * h), b = read_all(argv[4], (size_t)w * h);
    ORB_SLAM2::ORBextractor ex(2 * nf, 1.2f, 8, 20, 7);  // mono initialisation uses 2*nFeatures (src/Tracking.cc:124-125)
    FrameLite F[2];
    unsigned char* imgs[2] = {a.data(), b.data()};
    FILE* out = fopen(argv[5], "wb");
    for (int k = 0; k < 2; ++k) {
        cv::Mat im(h, w, CV_8UC1, imgs[k], (size_t)w);
        ex(im, cv::Mat(), F[k].mvKeysUn, F[k].mDescriptors);
        F[k].grid.SetBounds(0.f, 0.f, (float)w, (float)h);
        F[k].grid.Assign(F[k].mvKeysUn);
        const int n = (int)F[k].mvKeysUn.size();
        fwrite(&n, 4, 1, out);
        for (int i = 0; i < n; ++i) {
            const cv::KeyPoint& kp = F[k].mvKeysUn[i];
            const float rec[6] = {kp.pt.x, kp.pt.y, kp.size, kp.angle, kp.response, (float)kp.octave};
            fwrite(rec, 4, 6, out);
        }
        for (int i = 0; i < n; ++i) fwrite(F[k].mDescriptors.ptr(i), 1, 32, out);
        if (k == 0) {  // level 3 of the host pyramid copy, to check mvImagePyramid
            const cv::Mat& p = ex.mvImagePyramid[3];
            const int dims[2] = {p.cols, p.rows};
            fwrite(dims, 4, 2, out);
            for (int y = 0; y < p.rows; ++y) fwrite(p.ptr(y), 1, p.cols, out);
        }
    }
    std::vector<cv::Point2f> prev(F[0].mvKeysUn.size());
    for (size_t i = 0; i < prev.size(); ++i) prev[i] = F[0].mvKeysUn[i].pt;
    std::vector<int> m12;
    ORB_SLAM2::ORBmatcher matcher(0.9f, true);
    const int nm = matcher.SearchForInitialization(F[0], F[1], prev, m12, 100);
    const int n1 = (int)m12.size();
    fwrite(&nm, 4, 1, out);
    fwrite(&n1, 4, 1, out);
    fwrite(m12.data(), 4, n1, out);
    std::vector<int> bf;
    const int nb = matcher.SearchBruteForce(F[0].mDescriptors, F[1].mDescriptors, bf);
    fwrite(&nb, 4, 1, out);
    fwrite(bf.data(), 4, bf.size(), out);
    // SearchByBoW(KF,KF): feature vectors = a fixed hash of the keypoint position (nodes 0..39); map points on
    // 7 of 8 keypoints, every 13th one bad
    KeyFrameLite K[2];
    std::vector<MapPointLite> pool[2];
    for (int k = 0; k < 2; ++k) {
        K[k].mvKeysUn = F[k].mvKeysUn; K[k].mDescriptors = F[k].mDescriptors;
        const int n = (int)K[k].mvKeysUn.size();
        pool[k].resize(n);
        K[k].mps.assign(n, (MapPointLite*)NULL);
        for (int i = 0; i < n; ++i) {
            const cv::KeyPoint& kp = K[k].mvKeysUn[i];
            K[k].mFeatVec[(unsigned)(((int)(kp.pt.x / 80) + 8 * (int)(kp.pt.y / 96)) % 40)].push_back((unsigned)i);
            pool[k][i].bad = i % 13 == 5;
            if (i % 8 != 3) K[k].mps[i] = &pool[k][i];
        }
    }
    std::vector<MapPointLite*> m12bow;
    ORB_SLAM2::ORBmatcher bowMatcher(0.75f, true);
    const int nbow = bowMatcher.SearchByBoW<KeyFrameLite, MapPointLite>(&K[0], &K[1], m12bow);
    fwrite(&nbow, 4, 1, out);
    for (size_t i = 0; i < m12bow.size(); ++i) { const int v = m12bow[i] ? (int)(m12bow[i] - &pool[1][0]) : -1; fwrite(&v, 4, 1, out); }
    fclose(out);
    printf("bow matches %d; ", nbow);
    printf("facade ok: %d + %d keypoints, %d init matches, %d brute-force matches\n", (int)F[0].mvKeysUn.size(), (int)F[1].mvKeysUn.size(), nm, nb);
    return 0;
}
