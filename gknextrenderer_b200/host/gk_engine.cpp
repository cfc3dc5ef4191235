// gk_engine.cpp — see gk_engine.h.
#include "gk_engine.h"

namespace gk {

static float HaltonSequence(int index, int base) // Engine.cpp:609-618
{
    float f = 1.0f, result = 0.0f;
    while (index > 0) {
        f = f / base;
        result = result + f * (index % base);
        index = index / base;
    }
    return result;
}

GkUniformBufferObject EngineMirror::GetUniformBufferObject(VkOffset2D offset, VkExtent2D extent) // Engine.cpp:660-773
{
    GkUniformBufferObject ubo{};
    const Assets::Camera renderCam = scene.GetRenderCamera();
    auto& env = scene.GetEnvSettings();
    mat4 ModelView = renderCam.ModelView;
    mat4 Projection = perspective(radians(renderCam.FieldOfView), extent.width / static_cast<float>(extent.height), 0.1f, 10000.0f);

    ubo.FastGather = userSettings.FastGather;
    ubo.FastInterpole = userSettings.FastInterpole;
    ubo.DebugDraw_Lighting = userSettings.DebugDraw_Lighting;
    ubo.DisableSpatialReuse = userSettings.DisableSpatialReuse;
    ubo.SuperResolution = userSettings.SuperResolution;
    Projection.m[1][1] *= -1;
    const mat4 ProjectionUnJit = Projection;

    if (userSettings.TAA) {
        const int n = userSettings.TemporalFrames;
        const int i = (int)(totalFrames % (uint32_t)n);
        const float jx = HaltonSequence(i + 1, 2) - 0.5f, jy = HaltonSequence(i + 1, 3) - 0.5f;
        Projection.m[2][0] = jx / static_cast<float>(extent.width) * 2.0f;
        Projection.m[2][1] = jy / static_cast<float>(extent.height) * 2.0f;
    }

    const mat4 MVI = inverse(ModelView), PI = inverse(Projection);
    const mat4 VP = Projection * ModelView, VPu = ProjectionUnJit * ModelView;
    memcpy(ubo.ModelView, ModelView.data(), 64);
    memcpy(ubo.Projection, Projection.data(), 64);
    memcpy(ubo.ModelViewInverse, MVI.data(), 64);
    memcpy(ubo.ProjectionInverse, PI.data(), 64);
    memcpy(ubo.ViewProjection, VP.data(), 64);
    memcpy(ubo.ViewProjectionUnJit, VPu.data(), 64);
    memcpy(ubo.PrevViewProjection, prevUBO_.TotalFrames != 0 ? prevUBO_.ViewProjection : ubo.ViewProjection, 64);
    memcpy(ubo.PrevViewProjectionUnJit, prevUBO_.TotalFrames != 0 ? prevUBO_.ViewProjectionUnJit : ubo.ViewProjectionUnJit, 64);

    ubo.ViewportRect[0] = (float)offset.x, ubo.ViewportRect[1] = (float)offset.y, ubo.ViewportRect[2] = (float)extent.width, ubo.ViewportRect[3] = (float)extent.height;
    // SunViewProjection feeds only the shadow-map illuminator (other renderers): left identity.
    mat4 I;
    memcpy(ubo.SunViewProjection, I.data(), 64);
    ubo.SelectedId = scene.GetSelectedId();

    ubo.Aperture = renderCam.Aperture;
    ubo.FocusDistance = renderCam.FocalDistance;

    ubo.SkyRotation = env.SkyRotation;
    ubo.MaxNumberOfBounces = userSettings.MaxNumberOfBounces;
    ubo.TotalFrames = totalFrames;
    ubo.NumberOfSamples = userSettings.NumberOfSamples;
    ubo.NumberOfBounces = userSettings.NumberOfBounces;
    ubo.AdaptiveSample = userSettings.AdaptiveSample;
    ubo.AdaptiveVariance = userSettings.AdaptiveVariance;
    ubo.AdaptiveSteps = userSettings.AdaptiveSteps;
    ubo.TAA = userSettings.TAA;
    ubo.RandomSeed = 0; // rand() in the reference; unused by the path tracer
    const vec3 sd = env.SunDirection();
    ubo.SunDirection[0] = sd.x, ubo.SunDirection[1] = sd.y, ubo.SunDirection[2] = sd.z, ubo.SunDirection[3] = 0;
    ubo.SunColor[0] = ubo.SunColor[1] = ubo.SunColor[2] = 1.f * env.SunIntensity, ubo.SunColor[3] = 0;
    ubo.SkyIntensity = env.SkyIntensity;
    ubo.SkyIdx = (uint32_t)env.SkyIdx;
    // The reference fills BackGroundColor with (0.4,0.6,1.0)*4*SkyIntensity and never reads it;
    // this backend reads it as the constant sky texel (before the min(10,.) clamp and
    // the SkyIntensity scale of SampleIBL), so the mirror stores the plain colour.
    ubo.BackGroundColor[0] = env.SkyColor.x, ubo.BackGroundColor[1] = env.SkyColor.y, ubo.BackGroundColor[2] = env.SkyColor.z, ubo.BackGroundColor[3] = 0;
    ubo.HasSky = env.HasSky;
    ubo.HasSun = env.HasSun && env.SunIntensity > 0;

    ubo.ShowHeatmap = userSettings.ShowVisualDebug;
    ubo.HeatmapScale = userSettings.HeatmapScale;
    ubo.UseCheckerBoard = userSettings.UseCheckerBoardRendering;
    ubo.TemporalFrames = progressiveRendering ? (1024 / userSettings.TemporalFrames) : userSettings.TemporalFrames;
    ubo.HDR = hdr;
    ubo.PaperWhiteNit = userSettings.PaperWhiteNit;
    ubo.LightCount = scene.GetLightCount();
    ubo.BFSigma = userSettings.DenoiseSigma;
    ubo.BFSigmaLum = userSettings.DenoiseSigmaLum;
    ubo.BFSigmaNormal = userSettings.DenoiseSigmaNormal;
    ubo.BFSize = userSettings.Denoiser ? userSettings.DenoiseSize : 0;
    ubo.ShowEdge = userSettings.ShowEdge;
    ubo.ProgressiveRender = progressiveRendering;

    prevUBO_ = ubo;
    return ubo;
}

void EngineMirror::GetScreenToWorldRay(vec2 locationSS, VkExtent2D extent, vec3& org, vec3& dir) const // Engine.cpp:464-479
{
    const vec2 uv(locationSS.x / extent.width * 2.0f - 1.0f, locationSS.y / extent.height * 2.0f - 1.0f);
    const mat4 MVI = mat4::from(prevUBO_.ModelViewInverse), PI = mat4::from(prevUBO_.ProjectionInverse);
    const vec4 origin = MVI * vec4(0, 0, 0, 1);
    const vec4 target = PI * vec4(uv.x, uv.y, 1, 1);
    const vec3 tn = normalize(vec3(target.x, target.y, target.z));
    const vec4 rd = MVI * vec4(tn, 0.0f);
    org = vec3(origin.x, origin.y, origin.z);
    dir = vec3(rd.x, rd.y, rd.z);
}

void CudaPathTracingRenderer::CreateSwapChain(const VkExtent2D& extent)
{
    DeleteSwapChain();
    GkConfig cfg{};
    cfg.device = device_;
    cfg.width = extent.width, cfg.height = extent.height;
    cfg.tileIndex = tileIndex_, cfg.tileCount = tileCount_, cfg.tileRows = tileRows_;
    cfg.flags = traceAllRows_ ? GK_CFG_TRACE_ALL_ROWS : GK_CFG_DEFAULT;
    check(gk_create(&cfg, &ctx_), "gk_create");
    extent_ = extent;
    instancesUploaded_ = false;
}

void CudaPathTracingRenderer::DeleteSwapChain()
{
    if (ctx_) gk_destroy(ctx_);
    ctx_ = nullptr;
}

void CudaPathTracingRenderer::OnPostLoadScene()
{
    check(gk_upload_scene(ctx_, &GetScene().Desc()), "gk_upload_scene");
    GetScene().MarkDirty();
    instancesUploaded_ = false;
}

void CudaPathTracingRenderer::BeforeNextFrame()
{
    // Scene::UpdateNodes -> AfterUpdateScene -> TLAS update (VulkanBaseRenderer.cpp:960-970)
    auto& scene = GetScene();
    if (scene.UpdateNodes() || !instancesUploaded_) {
        const auto& px = scene.GetNodeProxys();
        // The reference rebuilds its TLAS on every dirty frame (RayTraceBaseRenderer.cpp:216-228).
        // Here an update with an unchanged instance count may refit the existing tree; the backend
        // falls back to a rebuild by itself when the refitted tree got too loose.
        const bool refit = instancesUploaded_ && px.size() == lastInstanceCount_;
        if (refit && !scene.LastUpdateWasFull() && scene.ChangedProxies().size() * 4 < px.size()) {
            // few nodes changed (MarkNodeDirty): only their proxies travel, the rest of the array is on the device already
            const auto& changed = scene.ChangedProxies();
            sparseStaging_.resize(changed.size());
            for (size_t k = 0; k < changed.size(); ++k) sparseStaging_[k] = px[changed[k]];
            check(gk_update_instances_sparse(ctx_, changed.data(), sparseStaging_.data(), (uint32_t)changed.size(), 1), "gk_update_instances_sparse");
            instanceBytesUploaded_ += changed.size() * (sizeof(GkNodeProxy) + sizeof(uint32_t));
        } else {
            check(gk_update_instances(ctx_, px.data(), (uint32_t)px.size(), refit ? 1 : 0), "gk_update_instances");
            instanceBytesUploaded_ += px.size() * sizeof(GkNodeProxy);
        }
        updatesSinceRebuild_ = refit ? updatesSinceRebuild_ + 1 : 0;
        lastInstanceCount_ = px.size();
        instancesUploaded_ = true;
    }
}

void CudaPathTracingRenderer::Render(void*, uint32_t)
{
    GkUniformBufferObject ubo = baseRender_.GetUniformBufferObject({0, 0}, extent_);
    check(gk_set_ubo(ctx_, &ubo), "gk_set_ubo");
    check(gk_render_frame(ctx_), "gk_render_frame");
    baseRender_.totalFrames += 1;
}

} // namespace gk
