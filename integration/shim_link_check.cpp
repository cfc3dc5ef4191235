// Links the shim against libgknext_cuda.so (every C-ABI symbol it uses must resolve) with a trivial present hook.
#include "Rendering/VulkanBaseRenderer.hpp"
#include <cstdio>
namespace Cuda {
std::unique_ptr<Vulkan::LogicRendererBase> MakeCudaLogicRenderer(Vulkan::VulkanBaseRenderer& base);
void PresentDenoised(VkCommandBuffer, VkImage, const void*, size_t) {}
}
int main()
{
    Vulkan::VulkanBaseRenderer base;
    auto r = Cuda::MakeCudaLogicRenderer(base);
    std::puts(r ? "shim linked" : "null");
    return 0; // nothing is rendered here: CreateSwapChain would need a B200
}
