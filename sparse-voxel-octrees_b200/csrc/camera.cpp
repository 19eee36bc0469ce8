// Per-frame camera constants, computed on the host once per frame.
//
// Restates what renderBatch derives from the global matrix stacks before its
// loops (reference src/Main.cpp:149-163) together with the few Mat4 operations
// behind it: MatrixStack::get(INV_MODELVIEW_STACK) = model.pseudoInvert()*view
// (src/math/MatrixStack.cpp:71-73), Mat4::pseudoInvert (src/math/Mat4.cpp:59-65),
// the 4x4 product (:67-78), Mat4*Vec3 (:80-86), Mat4::rotXYZ (:114-125) and the
// orbit camera of renderLoop (src/Main.cpp:212-213, 244-245). About twenty
// scalars per frame; float operation order is the reference's so the rays that
// the kernels generate are bit-identical. Must be compiled without FMA
// contraction (-ffp-contract=off); libm's sinf/cosf/tanf/sqrtf are the same
// calls the reference makes through std::sin/cos/tan/sqrt on floats.
#include "camera.hpp"

#include <cmath>
#include <cstring>

namespace svo {

namespace {

const float kPi = float(3.14159265358979323846);

struct M4 {
    float a[16]; // row-major, a[row*4 + col]
};

M4 identity() {
    M4 m;
    for (int i = 0; i < 16; ++i) m.a[i] = (i%5 == 0) ? 1.0f : 0.0f;
    return m;
}

M4 translation(float x, float y, float z) {
    M4 m = identity();
    m.a[3] = x;
    m.a[7] = y;
    m.a[11] = z;
    return m;
}

M4 mul(const M4 &l, const M4 &r) {
    M4 out;
    for (int row = 0; row < 4; ++row)
        for (int col = 0; col < 4; ++col)
            out.a[row*4 + col] = l.a[row*4]*r.a[col] + l.a[row*4 + 1]*r.a[4 + col] +
                                 l.a[row*4 + 2]*r.a[8 + col] + l.a[row*4 + 3]*r.a[12 + col];
    return out;
}

// transpose of the rotation part times the negated translation
M4 pseudoInverse(const M4 &m) {
    M4 rot;
    for (int row = 0; row < 4; ++row)
        for (int col = 0; col < 4; ++col)
            rot.a[row*4 + col] = m.a[col*4 + row];
    rot.a[12] = rot.a[13] = rot.a[14] = 0.0f;
    return mul(rot, translation(-m.a[3], -m.a[7], -m.a[11]));
}

M4 rotationXYZ(float degX, float degY, float degZ) {
    float rx = degX*kPi/180.0f, ry = degY*kPi/180.0f, rz = degZ*kPi/180.0f;
    float cx = std::cos(rx), cy = std::cos(ry), cz = std::cos(rz);
    float sx = std::sin(rx), sy = std::sin(ry), sz = std::sin(rz);
    M4 m = identity();
    m.a[0] = cy*cz;  m.a[1] = -cx*sz + sx*sy*cz;  m.a[2]  =  sx*sz + cx*sy*cz;
    m.a[4] = cy*sz;  m.a[5] =  cx*cz + sx*sy*sz;  m.a[6]  = -sx*cz + cx*sy*sz;
    m.a[8] = -sy;    m.a[9] =  sx*cy;             m.a[10] =  cx*cy;
    return m;
}

// Mat4*Vec3 with w = 1
void transformPoint(const M4 &m, float x, float y, float z, float out[3]) {
    for (int row = 0; row < 3; ++row)
        out[row] = m.a[row*4]*x + m.a[row*4 + 1]*y + m.a[row*4 + 2]*z + m.a[row*4 + 3];
}

} // namespace

void orbitCamera(float pitchDeg, float yawDeg, float radius, svo_camera &out) {
    M4 model = mul(rotationXYZ(pitchDeg, 0.0f, 0.0f), rotationXYZ(0.0f, yawDeg, 0.0f));
    M4 view = translation(0.0f, 0.0f, -radius);
    std::memcpy(out.model, model.a, sizeof out.model);
    std::memcpy(out.view, view.a, sizeof out.view);
}

void frameConstants(const svo_camera &cam, const float center[3], int width, int height, int strips,
                    svo_frame_constants &out) {
    M4 model, view;
    std::memcpy(model.a, cam.model, sizeof model.a);
    std::memcpy(view.a, cam.view, sizeof view.a);
    M4 tform = mul(pseudoInverse(model), view);

    out.width = width;
    out.height = height;
    out.strips = strips;
    out.tile_size = 8;                                   // TileSize, Main.cpp:63

    float eye[3];
    transformPoint(tform, 0.0f, 0.0f, 0.0f, eye);
    for (int i = 0; i < 3; ++i) out.pos[i] = eye[i] + center[i] + 1.0f;   // Main.cpp:152

    tform.a[3] = tform.a[7] = tform.a[11] = 0.0f;        // Main.cpp:154

    out.a11 = tform.a[0]; out.a12 = tform.a[1];
    out.a21 = tform.a[4]; out.a22 = tform.a[5];
    out.a31 = tform.a[8]; out.a32 = tform.a[9];

    out.scale = 2.0f/width;                              // Main.cpp:156
    out.tile_scale = out.tile_size*out.scale;            // :157
    float planeDist = 1.0f/std::tan(kPi/6.0f);           // :158
    out.zx = planeDist*tform.a[2];                       // :159
    out.zy = planeDist*tform.a[6];
    out.zz = planeDist*tform.a[10];
    out.coarse_scale = 2.0f*out.tile_size/(planeDist*height);   // :160
    out.aspect = height/(float)width;                    // Main.cpp:62

    float l[3];
    transformPoint(tform, -1.0f, 1.0f, -1.0f, l);        // :163
    float invLen = 1.0f/std::sqrt(l[0]*l[0] + l[1]*l[1] + l[2]*l[2]);
    for (int i = 0; i < 3; ++i) out.light[i] = l[i]*invLen;
    out.beam_bias = 0.03f;                               // :197
}

// ---- the viewer's camera control (row f4) ------------------------------------------------------------
// Events.cpp's processEvent (:38-77) and read-and-clear speed getters (:110-124), followed by what renderLoop
// does once waitEvent returns something other than idle mouse motion (Main.cpp:229-252). Float operations in
// the reference's order (std::fmod / std::fabs / std::min / std::max on floats).

void viewerInit(svo_viewer_state &s) {
    std::memset(&s, 0, sizeof s);
    s.radius = 1.0f;                                                  // Main.cpp:207-209
    const M4 model = identity(), view = translation(0.0f, 0.0f, -s.radius);   // :212-213
    std::memcpy(s.camera.model, model.a, sizeof s.camera.model);
    std::memcpy(s.camera.view, view.a, sizeof s.camera.view);
}

int viewerFeed(svo_viewer_state &s, const svo_viewer_event &e) {
    if (s.quit) return SVO_VIEWER_QUIT;
    switch (e.type) {                                                 // Events.cpp:38-77
    case SVO_EVENT_MOUSE_MOTION: s.mouse_dx = e.dx; s.mouse_dy = e.dy; break;
    case SVO_EVENT_BUTTON_DOWN:
        if (e.code == SVO_BUTTON_LEFT) s.mouse_down[0] = 1;
        else if (e.code == SVO_BUTTON_RIGHT) s.mouse_down[1] = 1;
        break;
    case SVO_EVENT_BUTTON_UP:
        if (e.code == SVO_BUTTON_LEFT) s.mouse_down[0] = 0;
        else if (e.code == SVO_BUTTON_RIGHT) s.mouse_down[1] = 0;
        break;
    case SVO_EVENT_KEY_DOWN: if (e.code == SVO_KEY_ESCAPE) s.escape_down = 1; break;
    case SVO_EVENT_KEY_UP: if (e.code == SVO_KEY_ESCAPE) s.escape_down = 0; break;
    default: break;
    }
    if (e.type == SVO_EVENT_MOUSE_MOTION && !s.mouse_down[0] && !s.mouse_down[1]) return SVO_VIEWER_WAIT;   // Main.cpp:229
    if (s.escape_down) s.quit = 1;                                    // :231-234
    const float mx = float(s.mouse_dx), my = float(s.mouse_dy);       // :236-237 (the getters clear the speeds)
    s.mouse_dx = s.mouse_dy = 0;
    if (s.mouse_down[0] && (mx != 0 || my != 0)) {                    // :238-246
        s.pitch = std::fmod(s.pitch - my, 360.0f);
        s.yaw = std::fmod(s.yaw + (std::fabs(s.pitch) > 90.0f ? mx : -mx), 360.0f);
        if (s.pitch > 180.0f) s.pitch -= 360.0f;
        else if (s.pitch < -180.0f) s.pitch += 360.0f;
        const M4 model = mul(rotationXYZ(s.pitch, 0.0f, 0.0f), rotationXYZ(0.0f, s.yaw, 0.0f));
        std::memcpy(s.camera.model, model.a, sizeof s.camera.model);
        s.preview = 1;
    } else if (s.mouse_down[1] && my != 0) {                          // :247-251
        const float lo = 1.0f - my*0.01f;
        const float clampedLo = (lo < 0.5f) ? 0.5f : lo;              // std::max(lo, 0.5f)
        const float factor = (1.5f < clampedLo) ? 1.5f : clampedLo;   // std::min(.., 1.5f)
        s.radius *= factor;
        s.radius = (25.0f < s.radius) ? 25.0f : s.radius;             // std::min(radius, 25.0f)
        const M4 view = translation(0.0f, 0.0f, -s.radius);
        std::memcpy(s.camera.view, view.a, sizeof s.camera.view);
        s.preview = 1;
    } else {
        s.preview = 0;                                                // :252-254
    }
    return s.quit ? SVO_VIEWER_QUIT : SVO_VIEWER_FRAME;
}

} // namespace svo
