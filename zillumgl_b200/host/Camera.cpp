#include "Camera.h"

namespace zillum {

void Camera::setDir(Vec3f dir) {   // Camera.cpp:86-94
    dir = normalize(dir);
    mAngle.y = std::asin(dir.z / length(dir)) * 57.295779513082320876798154814105f;
    float lxy = std::sqrt(dir.x * dir.x + dir.y * dir.y);
    mAngle.x = std::asin(dir.y / lxy) * 57.295779513082320876798154814105f;
    if (dir.x < 0) mAngle.x = 180.0f - mAngle.x;
    update();
}

void Camera::update() {            // Camera.cpp:149-162
    float x = std::sin(radians(mAngle.x)) * std::cos(radians(mAngle.y));
    float y = std::cos(radians(mAngle.x)) * std::cos(radians(mAngle.y));
    float z = std::sin(radians(mAngle.y));
    mFront = normalize(Vec3f(x, y, z));
    mRight = normalize(cross(mFront, Vec3f(0.0f, 0.0f, 1.0f)));
    // roll: the reference passes mAngle.z to glm::rotate un-converted (radians)
    mRight = normalize(rotation(mAngle.z, mFront) * mRight);
    mUp = normalize(cross(mRight, mFront));
}

ZlCamera Camera::uniforms() const {
    ZlCamera c;
    for (int i = 0; i < 3; i++) { c.F[i] = mFront[i]; c.R[i] = mRight[i]; c.U[i] = mUp[i]; c.pos[i] = mPos[i]; }
    Mat3f inv = inverse(Mat3f{mRight, mUp, mFront});
    const Vec3f* cols[3] = {&inv.c0, &inv.c1, &inv.c2};
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) c.matInv[3 * j + i] = (*cols[j])[i];
    c.tanFOV = std::tan(radians(mFOV * 0.5f));
    c.asp = mAspect;
    c.lensRadius = mLensRadius;
    c.focalDist = mFocalDist;
    return c;
}

}  // namespace zillum
