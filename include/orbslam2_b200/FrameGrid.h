// The 64 x 48 keypoint lookup grid of the reference's Frame (FRAME_GRID_COLS/ROWS,
// /root/reference/include/Frame.h:37-38): AssignFeaturesToGrid (src/Frame.cc:230-245), PosInGrid
// (382-392) and GetFeaturesInArea (327-380). It defines the candidate ORDER (ix-major, then iy, then
// insertion) that every windowed search iterates in, hence the tie-breaks of the match sets.
#ifndef ORBSLAM2_B200_FRAMEGRID_H
#define ORBSLAM2_B200_FRAMEGRID_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

namespace ORB_SLAM2 {

template <class KeyPointT>
class FrameGrid {
public:
    static const int COLS = 64, ROWS = 48;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<size_t> mGrid[COLS][ROWS];
    const std::vector<KeyPointT>* keys = nullptr;

    void SetBounds(float minX, float minY, float maxX, float maxY) {
        mnMinX = minX; mnMinY = minY; mnMaxX = maxX; mnMaxY = maxY;
        mfGridElementWidthInv = static_cast<float>(COLS) / static_cast<float>(mnMaxX - mnMinX);
        mfGridElementHeightInv = static_cast<float>(ROWS) / static_cast<float>(mnMaxY - mnMinY);
    }
    void Assign(const std::vector<KeyPointT>& keysUn) {
        keys = &keysUn;
        for (int i = 0; i < COLS; i++) for (int j = 0; j < ROWS; j++) mGrid[i][j].clear();
        for (size_t i = 0; i < keysUn.size(); i++) {
            const int posX = (int)std::round((keysUn[i].pt.x - mnMinX) * mfGridElementWidthInv);
            const int posY = (int)std::round((keysUn[i].pt.y - mnMinY) * mfGridElementHeightInv);
            if (posX < 0 || posX >= COLS || posY < 0 || posY >= ROWS) continue;
            mGrid[posX][posY].push_back(i);
        }
    }
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const {
        std::vector<size_t> vIndices;
        const int nMinCellX = std::max(0, (int)std::floor((x - mnMinX - r) * mfGridElementWidthInv));
        if (nMinCellX >= COLS) return vIndices;
        const int nMaxCellX = std::min(COLS - 1, (int)std::ceil((x - mnMinX + r) * mfGridElementWidthInv));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = std::max(0, (int)std::floor((y - mnMinY - r) * mfGridElementHeightInv));
        if (nMinCellY >= ROWS) return vIndices;
        const int nMaxCellY = std::min(ROWS - 1, (int)std::ceil((y - mnMinY + r) * mfGridElementHeightInv));
        if (nMaxCellY < 0) return vIndices;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                for (size_t j = 0; j < mGrid[ix][iy].size(); j++) {
                    const KeyPointT& kp = (*keys)[mGrid[ix][iy][j]];
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    if (std::fabs(kp.pt.x - x) < r && std::fabs(kp.pt.y - y) < r) vIndices.push_back(mGrid[ix][iy][j]);
                }
        return vIndices;
    }
};

}  // namespace ORB_SLAM2
#endif
