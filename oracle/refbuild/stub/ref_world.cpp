/* STAND-INS — TEST INFRASTRUCTURE ONLY (see ref_world.h). */
#include "Frame.h"

namespace ORB_SLAM2 {

LightMapPoint::LightMapPoint(MapPoint* pMP) : mnMapPointId(pMP ? pMP->mnId : 0), mpCache(0), mp(pMP) {}
LightKeyFrame::LightKeyFrame(KeyFrame* pKF) : mnId(pKF ? pKF->mnId : 0), kf(pKF) {}

/* MapPoint.cc:484-514 */
int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) {
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pKF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
    return nScale;
}
int MapPoint::PredictScale(const float& currentDist, Frame* pF) {
    float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
    return nScale;
}

std::set<MapPoint*> KeyFrame::GetMapPoints() {
    std::set<MapPoint*> s;
    for (size_t i = 0; i < mvpMapPoints.size(); i++) {
        MapPoint* p = mvpMapPoints[i].getMapPoint();
        if (p && !p->isBad()) s.insert(p);
    }
    return s;
}
std::vector<MapPoint*> KeyFrame::GetMapPointMatches() {
    std::vector<MapPoint*> v(mvpMapPoints.size());
    for (size_t i = 0; i < v.size(); i++) v[i] = mvpMapPoints[i].getMapPoint();
    return v;
}
std::vector<size_t> KeyFrame::GetFeaturesInArea(const float& x, const float& y, const float& r) const {
    std::vector<size_t> v;
    for (size_t i = 0; i < mvKeysUn.size(); i++)
        if (fabs(mvKeysUn[i].pt.x - x) < r && fabs(mvKeysUn[i].pt.y - y) < r) v.push_back(i);
    return v;
}

/* Converter.cc:27-35 */
std::vector<cv::Mat> Converter::toDescriptorVector(const cv::Mat& Descriptors) {
    std::vector<cv::Mat> vDesc;
    vDesc.reserve(Descriptors.rows);
    for (int j = 0; j < Descriptors.rows; j++) vDesc.push_back(Descriptors.row(j));
    return vDesc;
}

} // namespace ORB_SLAM2
