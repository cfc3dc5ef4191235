// STUB of the declarations CudaLogicRenderer.cpp touches in the reference's src/Rendering/VulkanBaseRenderer.hpp
// (class LogicRendererBase :239-269, ERendererType :48-55, rtDenoised :164, DelegateGetUniformBufferObject :148,
// GetScene :97).  Only names, signatures and sizes: it exists so that the shim can be compiled and ABI-checked in an
// environment without Vulkan, glm or the engine (tests/test_integration_shim.py).  In the real tree the shim includes
// the real header and this directory is not on the include path.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include "Assets/Scene.hpp"

struct VkExtent2D { uint32_t width, height; };
struct VkOffset2D { int32_t x, y; };
typedef struct VkCommandBuffer_T* VkCommandBuffer;
typedef struct VkImage_T* VkImage;

namespace Vulkan {

class RenderImage {
public:
    VkImage GetImage() const { return nullptr; }
};

enum ERendererType { ERT_PathTracing, ERT_Hybrid, ERT_ModernDeferred, ERT_LegacyDeferred, ERT_VoxelTracing,
                     ERT_CudaPathTracing /* the one value the integration adds */ };

class VulkanBaseRenderer {
public:
    Assets::Scene& GetScene() { return scene_; }
    std::function<Assets::UniformBufferObject(VkOffset2D, VkExtent2D)> DelegateGetUniformBufferObject;
    std::unique_ptr<RenderImage> rtDenoised;
private:
    Assets::Scene scene_;
};

class LogicRendererBase {
public:
    LogicRendererBase(VulkanBaseRenderer& baseRender) : baseRender_(baseRender) {}
    virtual ~LogicRendererBase() {}
    virtual void OnDeviceSet() {}
    virtual void CreateSwapChain(const VkExtent2D& extent) { (void)extent; }
    virtual void DeleteSwapChain() {}
    virtual void Render(VkCommandBuffer commandBuffer, uint32_t imageIndex) { (void)commandBuffer, (void)imageIndex; }
    virtual void BeforeNextFrame() {}
    VulkanBaseRenderer& baseRender_;
    const Assets::Scene& GetScene() { return baseRender_.GetScene(); }
};

} // namespace Vulkan
