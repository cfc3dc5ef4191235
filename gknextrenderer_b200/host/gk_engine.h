// gk_engine.h — host-side mirror of the pieces of NextEngine / VulkanBaseRenderer that sit
// directly above the renderer boundary:
//
//   UserSettings ...................... src/Runtime/UserSettings.hpp:6-57 (defaults Engine.cpp:125-177, Options.cpp:10-35)
//   NextEngine::GetUniformBufferObject  src/Runtime/Engine.cpp:660-773
//   NextEngine::GetScreenToWorldRay .... src/Runtime/Engine.cpp:464-479
//   ERendererType / LogicRendererBase .. src/Rendering/VulkanBaseRenderer.hpp:48-55, 239-269
//   PathTracingRenderer pass sequence .. src/Rendering/PathTracing/PathTracingRenderer.cpp:97-231
//
// CudaPathTracingRenderer is the logic renderer a maintainer registers behind the switch; it
// only forwards to the C ABI in include/gknext_cuda.h.
#pragma once
#include "../../include/gknext_cuda.h"
#include "gk_assets.h"
#include <stdexcept>

namespace gk {

struct VkExtent2D {
    uint32_t width, height;
};
struct VkOffset2D {
    int32_t x, y;
};

struct UserSettings {
    int RendererType = 0;
    int32_t NumberOfSamples = 8;
    int32_t NumberOfBounces = 5;
    int32_t MaxNumberOfBounces = 10;
    bool AdaptiveSample = false;
    float AdaptiveVariance = 6.0f;
    int AdaptiveSteps = 4;
    bool TAA = true;
    bool FastGather = false;
    bool FastInterpole = false;
    bool DebugDraw_Lighting = false;
    bool DisableSpatialReuse = false;
    int SuperResolution = 1;
    bool ShowVisualDebug = false;
    float HeatmapScale = 1.0f;
    bool UseCheckerBoardRendering = false;
    int TemporalFrames = 16;
    bool Denoiser = false;
    float DenoiseSigma = 2.0f;
    float DenoiseSigmaLum = 3.0f;
    float DenoiseSigmaNormal = 0.005f;
    int DenoiseSize = 5;
    float PaperWhiteNit = 600.f;
    bool ShowEdge = false;
};

enum ERendererType {
    ERT_PathTracing,
    ERT_Hybrid,
    ERT_ModernDeferred,
    ERT_LegacyDeferred,
    ERT_VoxelTracing,
    ERT_CudaPathTracing, // the one value this backend adds
};

// The part of NextEngine a logic renderer can reach: scene, settings, per-frame UBO.
class EngineMirror {
public:
    Assets::Scene scene;
    UserSettings userSettings;
    bool progressiveRendering = false;
    bool hdr = false;
    uint32_t totalFrames = 0;

    GkUniformBufferObject GetUniformBufferObject(VkOffset2D offset, VkExtent2D extent);
    void GetScreenToWorldRay(vec2 locationSS, VkExtent2D extent, vec3& org, vec3& dir) const;
    const GkUniformBufferObject& PrevUBO() const { return prevUBO_; }

private:
    GkUniformBufferObject prevUBO_{};
};

class LogicRendererBase {
public:
    explicit LogicRendererBase(EngineMirror& base) : baseRender_(base) {}
    virtual ~LogicRendererBase() {}
    virtual void OnDeviceSet() {}
    virtual void CreateSwapChain(const VkExtent2D&) {}
    virtual void DeleteSwapChain() {}
    virtual void Render(void* /*VkCommandBuffer*/, uint32_t /*imageIndex*/) {}
    virtual void BeforeNextFrame() {}
    EngineMirror& baseRender_;
    Assets::Scene& GetScene() { return baseRender_.scene; }
};

class CudaPathTracingRenderer : public LogicRendererBase {
public:
    explicit CudaPathTracingRenderer(EngineMirror& base, int device = -1) : LogicRendererBase(base), device_(device) {}
    ~CudaPathTracingRenderer() override { DeleteSwapChain(); }
    void OnDeviceSet() override {}
    void CreateSwapChain(const VkExtent2D& extent) override;
    void DeleteSwapChain() override;
    void Render(void* commandBuffer, uint32_t imageIndex) override;
    void BeforeNextFrame() override;
    // scene load boundary: called where Scene::RebuildMeshBuffer runs (Scene.cpp:101)
    void OnPostLoadScene();
    GkContext* Context() { return ctx_; }
    uint64_t InstanceBytesUploaded() const { return instanceBytesUploaded_; } // host->device bytes of all instance updates so far
    VkExtent2D Extent() const { return extent_; }
    void SetTile(uint32_t index, uint32_t count, uint32_t rows) { tileIndex_ = index, tileCount_ = count, tileRows_ = rows; }
    void SetTraceAllRows(bool on) { traceAllRows_ = on; } // frame-sharded progressive rendering (GK_CFG_TRACE_ALL_ROWS)

private:
    static void check(GkStatus s, const char* what)
    {
        if (s != GK_OK) throw std::runtime_error(std::string(what) + ": " + gk_last_error());
    }
    GkContext* ctx_ = nullptr;
    VkExtent2D extent_{0, 0};
    int device_;
    uint32_t tileIndex_ = 0, tileCount_ = 1, tileRows_ = 16;
    bool traceAllRows_ = false;
    bool instancesUploaded_ = false;
    size_t lastInstanceCount_ = 0;
    std::vector<GkNodeProxy> sparseStaging_;
    uint64_t instanceBytesUploaded_ = 0;
    uint32_t updatesSinceRebuild_ = 0;
};

} // namespace gk
