// STUB of the Assets:: declarations the shim touches (sizes and accessor names of src/Assets/Scene.hpp:52-105,
// Model.hpp:222-224, Material.hpp:40-81, Vertex.hpp, UniformBuffer.hpp).  See Rendering/VulkanBaseRenderer.hpp (stub).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace Assets {

struct Vertex { float Position[3], Normal[3], Tangent[4], TexCoord[2]; uint32_t MaterialIndex; }; // 52 B (Vertex.hpp)
struct Material { unsigned char bytes[64]; };       // 64-B GPU material (Material.hpp:40-75)
struct FMaterial final { std::string name_; uint32_t globalId_; Material gpuMaterial_; }; // Material.hpp:76-81
struct LightObject { unsigned char bytes[80]; };    // UniformBuffer.hpp
struct NodeProxy { unsigned char bytes[208]; };     // UniformBuffer.hpp, written by Scene.cpp:464-511
struct UniformBufferObject { unsigned char bytes[784]; }; // UniformBuffer.hpp; offsets asserted in gknext_types.h

class Model {
public:
    const std::vector<Vertex>& CPUVertices() const { return vertices_; }   // Model.hpp:222
    const std::vector<uint32_t>& CPUIndices() const { return indices_; }   // Model.hpp:224
private:
    std::vector<Vertex> vertices_;
    std::vector<uint32_t> indices_;
};

class Scene {
public:
    const std::vector<Model>& Models() const { return models_; }           // Scene.hpp:53
    std::vector<FMaterial>& Materials() { return materials_; }             // Scene.hpp:54
    const std::vector<LightObject>& Lights() const { return lights_; }     // Scene.hpp:56
    std::vector<NodeProxy>& GetNodeProxys() { return nodeProxys; }         // Scene.hpp:94
private:
    std::vector<Model> models_;
    std::vector<FMaterial> materials_;
    std::vector<LightObject> lights_;
    std::vector<NodeProxy> nodeProxys;
};

} // namespace Assets
