// Z-up yaw/pitch/roll thin-lens camera (reference: src/core/Camera.{h,cpp}).
#pragma once
#include "Math.h"
#include "../../include/zillum_cuda.h"

namespace zillum {

class Camera {
public:
    Camera(Vec3f pos = Vec3f(0, 0, 0), Vec3f angle = Vec3f(90.0f, 0.0f, 0.0f)) : mPos(pos), mAngle(angle) { update(); }

    void move(Vec3f v) { mPos = mPos + v; }
    void setFOV(float fov) { mFOV = fov; if (mFOV > 90.0f) mFOV = 90.0f; if (mFOV < 0.1f) mFOV = 0.1f; }
    void lookAt(Vec3f focus) { setDir(focus - mPos); }
    void setDir(Vec3f dir);
    void setPos(Vec3f p) { mPos = p; }
    void setAngle(Vec3f angle) { mAngle = angle; update(); }
    void setAspect(float asp) { mAspect = asp; }
    void setLensRadius(float r) { mLensRadius = r; }
    void setFocalDist(float d) { mFocalDist = d; }

    Vec3f pos() const { return mPos; }
    Vec3f angle() const { return mAngle; }
    Vec3f front() const { return mFront; }
    Vec3f right() const { return mRight; }
    Vec3f up() const { return mUp; }
    float FOV() const { return mFOV; }
    float aspect() const { return mAspect; }
    float lensRadius() const { return mLensRadius; }
    float focalDist() const { return mFocalDist; }

    // The camera uniforms every integrator uploads (NaivePath.cpp:49-59).
    ZlCamera uniforms() const;

private:
    void update();
    Vec3f mPos, mAngle, mFront, mRight, mUp{0.0f, 0.0f, 1.0f};
    float mFOV = 45.0f, mAspect = 1.0f, mLensRadius = 0.0f, mFocalDist = 1.0f;
};

}  // namespace zillum
