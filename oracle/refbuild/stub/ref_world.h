/* STAND-INS for the reference classes that cannot be compiled here — TEST INFRASTRUCTURE ONLY (oracle/refbuild).
 *
 * MapPoint.h, KeyFrame.h, LightMapPoint.h, LightKeyFrame.h and Converter.h drag in ROS, Boost.Serialization and
 * Eigen/g2o, none of which exist in this image. They are replaced - through their own include guards, see the Makefile -
 * by plain-data classes that expose exactly the members ORBmatcher.cc, Frame.cc and PnPsolver.cc touch, with the
 * reference's names and signatures (corbslam_client/include/MapPoint.h:83-165, KeyFrame.h:101-274,
 * LightMapPoint.h:33-68, Converter.h:36). They hold data and return it; the only behaviour restated is
 * MapPoint::PredictScale (MapPoint.cc:484-514) and Converter::toDescriptorVector (Converter.cc:27-35).
 * Frame, ORBextractor, ORBmatcher, PnPsolver, ORBVocabulary / DBoW2 are the reference's REAL classes. */
#ifndef CORB_REF_WORLD_H
#define CORB_REF_WORLD_H
#ifdef __cplusplus
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include <opencv2/core/core.hpp>

#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

namespace ORB_SLAM2 {

class Frame;
class KeyFrame;
class Map;
class Cache;
class KeyFrameDatabase;
class MapPoint;

class Cache { /* Cache.h:107-108: the update queues Optimizer.cc:230,254 feed (recorded, nothing else) */
public:
    void addUpdateKeyframe(KeyFrame* pKF) { updatedKFs.push_back(pKF); }
    void addUpdateMapPoint(MapPoint* pMP) { updatedMPs.push_back(pMP); }
    std::vector<KeyFrame*> updatedKFs;
    std::vector<MapPoint*> updatedMPs;
};

class LightMapPoint { /* LightMapPoint.h:33-68; the cache lookup is replaced by the pointer itself */
public:
    LightMapPoint() : mnMapPointId(0), mpCache(0), mp(0) {}
    LightMapPoint(MapPoint* pMP);
    bool operator<(const LightMapPoint& o) const { return mnMapPointId < o.mnMapPointId; }
    bool operator==(const LightMapPoint& o) const { return mnMapPointId == o.mnMapPointId; }
    MapPoint* getMapPoint() const { return mp; }
    long unsigned int mnMapPointId;
    Cache* mpCache;
    MapPoint* mp;
};

class LightKeyFrame {
public:
    LightKeyFrame() : mnId(0), kf(0) {}
    LightKeyFrame(KeyFrame* pKF);
    bool operator<(const LightKeyFrame& o) const { return mnId < o.mnId; }
    KeyFrame* getKeyFrame() const { return kf; }
    long unsigned int mnId;
    KeyFrame* kf;
};

class MapPoint { /* MapPoint.h:83-165 */
public:
    MapPoint() : mnId(0), mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), mbTrackInView(false), mnTrackScaleLevel(0),
                 mTrackViewCos(0), mnLastFrameSeen(0), mnBAGlobalForKF(0), nObs(0), bad(false), fixed(false), mfMinDistance(0),
                 mfMaxDistance(0), cache(0), nNormalUpdates(0) {}
    /* members Optimizer::BundleAdjustment touches (MapPoint.h:81-176) */
    void SetWorldPos(const cv::Mat& Pos) { Pos.copyTo(mWorldPos); }
    bool getFixed() { return fixed; }
    std::map<KeyFrame*, size_t> GetObservations() { return obs; }
    void UpdateNormalAndDepth() { nNormalUpdates++; }
    Cache* getCache() { return cache; }
    cv::Mat mPosGBA;
    long unsigned int mnBAGlobalForKF;
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    int Observations() { return nObs; }
    bool isBad() { return bad; }
    void AddObservation(KeyFrame* pKF, size_t idx) { obs[pKF] = idx; nObs++; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return obs.count(pKF) ? (int)obs[pKF] : -1; }
    bool IsInKeyFrame(KeyFrame* pKF) { return obs.count(pKF) != 0; }
    void Replace(MapPoint*) {}
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; } /* MapPoint.cc:473-481 */
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    int PredictScale(const float& currentDist, Frame* pF);

    long unsigned int mnId;
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    long unsigned int mnLastFrameSeen;
    /* plain data behind the getters */
    cv::Mat mWorldPos, mNormalVector, mDescriptor;
    int nObs;
    bool bad, fixed;
    float mfMinDistance, mfMaxDistance;
    std::map<KeyFrame*, size_t> obs;
    Cache* cache;
    int nNormalUpdates;
};

class KeyFrame { /* KeyFrame.h:101-274 */
public:
    KeyFrame() : mnId(0), fx(0), fy(0), cx(0), cy(0), invfx(0), invfy(0), mbf(0), mb(0), mThDepth(0), N(0), mnScaleLevels(0),
                 mfScaleFactor(0), mfLogScaleFactor(0), mnMinX(0), mnMinY(0), mnMaxX(0), mnMaxY(0), mnBAGlobalForKF(0), mpCacher(0),
                 bad(false), fixed(false) {}
    /* members Optimizer::BundleAdjustment touches (KeyFrame.h:95-111,194,244-246,286) */
    void SetPose(const cv::Mat& Tcw_) { Tcw_.copyTo(Tcw); }
    cv::Mat GetPose() { return Tcw.clone(); }
    bool isBad() { return bad; }
    bool getFixed() { return fixed; }
    cv::Mat mTcwGBA;
    long unsigned int mnBAGlobalForKF;
    Cache* mpCacher;
    bool bad, fixed;
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
    cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = LightMapPoint(pMP); }
    std::set<MapPoint*> GetMapPoints();
    std::vector<MapPoint*> GetMapPointMatches();
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx].getMapPoint(); }
    /* not on the pinned paths (SearchByProjection(KeyFrame*, ...), Fuse): plain filter, KeyFrame.cc's grid not restated */
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
    bool IsInImage(const float& x, const float& y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }

    long unsigned int mnId;
    float fx, fy, cx, cy, invfx, invfy, mbf, mb, mThDepth;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    int mnScaleLevels;
    float mfScaleFactor, mfLogScaleFactor;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    float mnMinX, mnMinY, mnMaxX, mnMaxY;
    std::vector<LightMapPoint> mvpMapPoints;
    cv::Mat Tcw, Ow;
};

class Converter { /* Converter.h:36 */
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors);
};

} // namespace ORB_SLAM2
#endif
#endif
