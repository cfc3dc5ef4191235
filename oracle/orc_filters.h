// orc_filters.h — CPU restatement of the reference's temporal reprojection and joint
// bilateral denoise/compose shaders.  TEST INFRASTRUCTURE ONLY (see orc_math.h).
//
//   assets/shaders/Process.ReProject.comp.slang:60-181   (calculateWeight :40-55)
//   assets/shaders/Process.DenoiseJBF.comp.slang:96-195  (JBF :70-94, EdgeDetect :55-68)
//   assets/shaders/common/Const_Func.slang:51-68, 84-126, 188-203 (ST2084, GT tonemap, YCoCg)
//
// Images are the reference's formats (src/Rendering/VulkanBaseRenderer.cpp:488-543):
// RGBA16F colour planes as 4 x uint16 half bit patterns per pixel, R32_UINT ids, RG32F motion.
// Loads outside an image return 0 (Vulkan robust image access), stores outside are dropped.
// lerp(a,b,t) is evaluated as a*(1-t) + b*t (the GLSL.std.450 FMix definition).
// PARITY UNPINNED: the reference ships no fixtures for these shaders and they cannot be
// compiled here; pinned by review only.
#pragma once
#include "../include/gknext_types.h"
#include "orc_math.h"

namespace orc {

void reproject(const GkUniformBufferObject& U, uint32_t W, uint32_t H, bool needClamp, bool needSpatio, const uint16_t* src,
               const uint16_t* history, const float* motion, const uint32_t* objId0, const uint32_t* objId1, const uint16_t* normal,
               uint16_t* out);

void denoiseJBF(const GkUniformBufferObject& U, uint32_t W, uint32_t H, const uint16_t* diffuse, const uint16_t* spec,
                const uint16_t* normal, const uint32_t* objId0, const uint32_t* objId1, const uint16_t* albedo, uint16_t* out);

} // namespace orc
