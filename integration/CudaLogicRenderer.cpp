// CudaLogicRenderer — the shim that puts the CUDA backend (include/gknext_cuda.h, lib/libgknext_cuda.so) behind the
// reference's renderer switch.  This is the file a maintainer adds to the reference tree as
// src/Rendering/Cuda/CudaLogicRenderer.cpp (INTEGRATION.md lists the three one-line edits that register it).
// It is compiled here against stub headers (integration/stubs) by tests/test_integration_shim.py, so its syntax and its
// use of the C ABI are checked; it cannot run here (no Vulkan).  The present step is left as one hook,
// PresentDenoised(), because it needs the engine's Vulkan helpers.
#include "gknext_cuda.h"
#include "Rendering/VulkanBaseRenderer.hpp"
#include "Assets/Scene.hpp"
#include "Utilities/Exception.hpp"
#include <string>
#include <vector>

namespace Cuda {

static_assert(sizeof(Assets::Vertex) == sizeof(GkVertex), "Assets::Vertex <-> GkVertex");
static_assert(sizeof(Assets::Material) == sizeof(GkMaterial), "Assets::Material <-> GkMaterial");
static_assert(sizeof(Assets::LightObject) == sizeof(GkLightObject), "Assets::LightObject <-> GkLightObject");
static_assert(sizeof(Assets::NodeProxy) == sizeof(GkNodeProxy), "Assets::NodeProxy <-> GkNodeProxy");
static_assert(sizeof(Assets::UniformBufferObject) == sizeof(GkUniformBufferObject), "Assets::UniformBufferObject <-> GkUniformBufferObject");

static void Check(GkStatus s, const char* what) // C status -> the engine's exception (Utilities/Exception.hpp:10-16)
{
    if (s != GK_OK) Throw(std::runtime_error(std::string(what) + ": " + gk_last_error()));
}

// Hands rtDenoised (RGBA16F, host memory) to the engine's resolve / present path (VulkanBaseRenderer.cpp:1192-1209):
// either a staging buffer + vkCmdCopyBufferToImage, or - zero copy - the VkImage's memory imported once with
// cudaImportExternalMemory (the OIDN interop of PathTracingRenderer.cpp:247-285 is the pattern).  Implemented in the engine.
void PresentDenoised(VkCommandBuffer cmd, VkImage target, const void* rgba16f, size_t bytes);

class CudaLogicRenderer final : public Vulkan::LogicRendererBase {
public:
    using LogicRendererBase::LogicRendererBase;
    ~CudaLogicRenderer() override { DeleteSwapChain(); }

    void CreateSwapChain(const VkExtent2D& extent) override // VulkanBaseRenderer.hpp:246
    {
        GkConfig cfg{};
        cfg.device = -1, cfg.width = extent.width, cfg.height = extent.height, cfg.tileCount = 1;
        Check(gk_create(&cfg, &ctx_), "gk_create");
        extent_ = extent;
        sceneUploaded_ = instancesUploaded_ = false;
        host_.resize(gk_plane_bytes(ctx_, GK_PLANE_DENOISED));
    }
    void DeleteSwapChain() override // hpp:247
    {
        if (ctx_) gk_destroy(ctx_);
        ctx_ = nullptr;
    }

    // Called from Scene::RebuildMeshBuffer (Scene.cpp:101-270) while the CPU vertex arrays still exist (they are freed at
    // Scene.cpp:195), or lazily from the first BeforeNextFrame after a load.
    void UploadScene()
    {
        Assets::Scene& scene = baseRender_.GetScene();
        std::vector<GkModelDesc> models;
        for (const auto& m : scene.Models())
            models.push_back({reinterpret_cast<const GkVertex*>(m.CPUVertices().data()), m.CPUIndices().data(), (uint32_t)m.CPUVertices().size(),
                              (uint32_t)m.CPUIndices().size()});
        std::vector<GkMaterial> mats;
        for (const auto& fm : scene.Materials()) mats.push_back(reinterpret_cast<const GkMaterial&>(fm.gpuMaterial_));
        GkSceneDesc d{};
        d.models = models.data(), d.materials = mats.data(), d.lights = reinterpret_cast<const GkLightObject*>(scene.Lights().data());
        d.modelCount = (uint32_t)models.size(), d.materialCount = (uint32_t)mats.size(), d.lightCount = (uint32_t)scene.Lights().size();
        Check(gk_upload_scene(ctx_, &d), "gk_upload_scene");
        sceneUploaded_ = true, instancesUploaded_ = false;
    }

    void NotifySceneUpdated() { sceneChanged_ = true; } // from RayTraceBaseRenderer::AfterUpdateScene (RayTraceBaseRenderer.cpp:176-228)

    void BeforeNextFrame() override // hpp:249; runs after Scene::UpdateNodes (VulkanBaseRenderer.cpp:967-969)
    {
        if (!sceneUploaded_) UploadScene();
        auto& proxies = baseRender_.GetScene().GetNodeProxys(); // the 208-byte records Scene.cpp:464-511 just wrote
        if (sceneChanged_ || !instancesUploaded_) {
            const bool refit = instancesUploaded_ && proxies.size() == lastCount_;
            Check(gk_update_instances(ctx_, reinterpret_cast<const GkNodeProxy*>(proxies.data()), (uint32_t)proxies.size(), refit ? 1 : 0), "gk_update_instances");
            lastCount_ = proxies.size(), instancesUploaded_ = true, sceneChanged_ = false;
        }
    }

    void Render(VkCommandBuffer cmd, uint32_t imageIndex) override // hpp:248
    {
        (void)imageIndex;
        // the same 784-byte block the shaders read (VulkanBaseRenderer.cpp:375-383, Engine.cpp:660-773)
        const Assets::UniformBufferObject ubo = baseRender_.DelegateGetUniformBufferObject(VkOffset2D{0, 0}, extent_);
        Check(gk_set_ubo(ctx_, reinterpret_cast<const GkUniformBufferObject*>(&ubo)), "gk_set_ubo");
        Check(gk_render_frame(ctx_), "gk_render_frame");
        Check(gk_readback(ctx_, GK_PLANE_DENOISED, host_.data(), host_.size()), "gk_readback");
        PresentDenoised(cmd, baseRender_.rtDenoised->GetImage(), host_.data(), host_.size());
    }

    // NextEngine::RayCastGPU (Engine.cpp:647-653): the queued RayCastIO records, answered in place
    void RayCast(std::vector<GkRayCastIO>& io) { Check(gk_raycast_task(ctx_, io.data(), (uint32_t)io.size()), "gk_raycast_task"); }

private:
    GkContext* ctx_ = nullptr;
    VkExtent2D extent_{};
    bool sceneUploaded_ = false, instancesUploaded_ = false, sceneChanged_ = true;
    size_t lastCount_ = 0;
    std::vector<unsigned char> host_;
};

// VulkanBaseRenderer::RegisterLogicRenderer (VulkanBaseRenderer.cpp:1071-1094) gains:
//     case ERendererType::ERT_CudaPathTracing: logicRenderers_[type] = Cuda::MakeCudaLogicRenderer(*this); break;
std::unique_ptr<Vulkan::LogicRendererBase> MakeCudaLogicRenderer(Vulkan::VulkanBaseRenderer& base) { return std::make_unique<CudaLogicRenderer>(base); }

} // namespace Cuda
