// Host-side per-frame camera constants (see camera.cpp).
#pragma once

#include "../../include/svo_b200.h"

namespace svo {

void orbitCamera(float pitchDeg, float yawDeg, float radius, svo_camera &out);
void frameConstants(const svo_camera &cam, const float center[3], int width, int height, int strips,
                    svo_frame_constants &out);

void viewerInit(svo_viewer_state &state);
int viewerFeed(svo_viewer_state &state, const svo_viewer_event &event);

} // namespace svo
