// Camera, host side: src/engine/graphics/camera.zig:5-43.
//
// The parity boundary is the 96-byte UniformData block (uvt_camera); the pitch/yaw →
// matrix helper restates zmath.matFromRollPitchYaw (zmath @6a0747fe, not vendored;
// DirectXMath convention, row vectors) and is off the parity path.
#include "uvt_host.h"

#include <cmath>
#include <cstring>

namespace {
constexpr float kPi = 3.14159265358979323846f;
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
}  // namespace

extern "C" {

void uvt_mat_from_pitch_yaw(float pitch, float yaw, float m[16]) {
    const float cp = std::cos(pitch), sp = std::sin(pitch);
    const float cy = std::cos(yaw), sy = std::sin(yaw);
    // roll = 0: rows (cy,0,-sy,0), (sp*sy,cp,sp*cy,0), (cp*sy,-sp,cp*cy,0), (0,0,0,1)
    const float r[16] = {cy, 0.0f, -sy, 0.0f,
                         sp * sy, cp, sp * cy, 0.0f,
                         cp * sy, -sp, cp * cy, 0.0f,
                         0.0f, 0.0f, 0.0f, 1.0f};
    std::memcpy(m, r, sizeof r);
}

void uvt_camera_init(uvt_camera_state *c) {
    std::memset(c, 0, sizeof *c);
    c->fov = kPi / 2.0f;  // camera.zig:6
    c->cam_mat[0] = c->cam_mat[5] = c->cam_mat[10] = c->cam_mat[15] = 1.0f;
}

void uvt_camera_rotate(uvt_camera_state *c, float pitch, float yaw) {
    c->pitch = clampf(c->pitch + pitch * 0.001f, -kPi / 2.0f, kPi / 2.0f);
    c->yaw = c->yaw + yaw * 0.001f;
    uvt_mat_from_pitch_yaw(c->pitch, c->yaw, c->cam_mat);
}

void uvt_camera_set_pos(uvt_camera_state *c, const float pos[4]) { std::memcpy(c->cam_pos, pos, sizeof(float) * 4); }

void uvt_camera_increment_fov(uvt_camera_state *c, float increment) {
    c->fov = clampf(c->fov + increment * 0.1f, 0.314f, 2.4f);
}

void uvt_camera_as_uniform_data(const uvt_camera_state *c, uvt_camera *out) {
    std::memset(out, 0, sizeof *out);
    std::memcpy(out->cam_pos, c->cam_pos, sizeof out->cam_pos);
    std::memcpy(out->cam_mat, c->cam_mat, sizeof out->cam_mat);
    out->fov = c->fov;
}

}  // extern "C"
