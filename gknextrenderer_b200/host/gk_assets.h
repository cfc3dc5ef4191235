// gk_assets.h — host-side mirror of the reference's scene API surface, just wide enough to
// feed the CUDA backend through include/gknext_cuda.h.  Same names and argument meaning as
// the reference so that a scene written against Assets::* reads the same here:
//
//   Assets::Material factories ......... src/Assets/Material.hpp:10-38
//   Assets::Model (+CreateBox/Sphere/CornellBox) src/Assets/Model.hpp, Model.cpp:929-1080
//   Assets::Node (TRS, 16 material slots, prev transform) src/Assets/Model.hpp:96-170, Model.cpp:1252-1296
//   Assets::Camera / EnvironmentSetting  src/Assets/Model.hpp:20-92
//   Assets::Scene (nodes/models/materials/lights, node proxies) src/Assets/Scene.hpp:31-207, Scene.cpp:464-511
//   CornellBox::Create ................. src/Assets/CornellBox.cpp:18-151
//   SceneList::CornellBox .............. src/Runtime/SceneList.cpp:186-247
//
// Out of scope: glTF loading, textures, animation tracks, physics bodies.
#pragma once
#include "../../include/gknext_types.h"
#include "gk_glm.h"
#include <array>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace gk::Assets {

using Vertex = GkVertex;
using LightObject = GkLightObject;
using NodeProxy = GkNodeProxy;

struct Material : GkMaterial {
    static Material make(vec3 d, float fuzz, float ri, uint32_t model, float metal, float ri2 = 0.f)
    {
        Material m{};
        m.Diffuse[0] = d.x, m.Diffuse[1] = d.y, m.Diffuse[2] = d.z, m.Diffuse[3] = 1.f;
        m.DiffuseTextureId = m.MRATextureId = m.NormalTextureId = -1;
        m.Fuzziness = fuzz, m.RefractionIndex = ri, m.MaterialModel = model, m.Metalness = metal, m.RefractionIndex2 = ri2;
        return m;
    }
    static Material Lambertian(vec3 d) { return make(d, 1.0f, 1.f, GK_MAT_LAMBERTIAN, 0); }
    static Material Metallic(vec3 d, float fuzz) { return make(d, fuzz, 1.45f, GK_MAT_METALLIC, 1); }
    static Material Mixture(vec3 d, float fuzz) { return make(d, fuzz, 1.45f, GK_MAT_MIXTURE, 0); }
    static Material Dielectric(float ri, float fuzz) { return make(vec3(1.f), fuzz, ri, GK_MAT_DIELECTRIC, 0, ri); }
    static Material Isotropic(vec3 d, float ri, float fuzz) { return make(d, fuzz, ri, GK_MAT_ISOTROPIC, 0); }
    static Material DiffuseLight(vec3 d) { return make(d, 0.0f, 0.0f, GK_MAT_DIFFUSE_LIGHT, 0); }
};

struct FMaterial {
    std::string name_;
    uint32_t globalId_;
    Material gpuMaterial_;
};

struct Camera {
    std::string name;
    mat4 ModelView;
    float FieldOfView = 40;
    float Aperture = 0;
    float FocalDistance = 10;
};

struct EnvironmentSetting {
    float ControlSpeed = 5.0f;
    bool GammaCorrection = true;
    bool HasSky = true;
    bool HasSun = false;
    int32_t SkyIdx = 0;
    float SunRotation = 0.5f;
    float SkyRotation = 0;
    float SkyIntensity = 100.0f;
    float SunIntensity = 500.0f;
    vec3 SkyColor = vec3(0.4f, 0.6f, 1.0f); // constant sky texel (this backend's stand-in for the HDR texture)
    std::vector<Camera> cameras;
    vec3 SunDirection() const
    {
        const float pi = 3.14159265358979323846f;
        return normalize(vec3(sinf(SunRotation * pi), 0.75f, cosf(SunRotation * pi)));
    }
};

class Model {
public:
    Model() = default;
    Model(std::vector<Vertex>&& v, std::vector<uint32_t>&& i, bool needGenTSpace = true);
    static Model CreateBox(const vec3& p0, const vec3& p1);
    static Model CreateSphere(const vec3& center, float radius);
    static uint32_t CreateCornellBox(float scale, std::vector<Model>& models, std::vector<FMaterial>& materials, std::vector<LightObject>& lights);
    // generalisations used by the procedural benchmark scenes (not in the reference)
    static Model CreateUVSphere(const vec3& center, float radius, int slices, int stacks);
    static Model CreateGridBox(const vec3& p0, const vec3& p1, int n, float bump, uint32_t seed);

    const std::vector<Vertex>& CPUVertices() const { return vertices_; }
    const std::vector<uint32_t>& CPUIndices() const { return indices_; }
    uint32_t NumberOfVertices() const { return (uint32_t)vertices_.size(); }
    uint32_t NumberOfIndices() const { return (uint32_t)indices_.size(); }
    vec3 GetLocalAABBMin() const { return aabbMin_; }
    vec3 GetLocalAABBMax() const { return aabbMax_; }
    uint32_t SectionCount() const; // Scene.cpp:138-148: slices of <=65535 triangles, at most 10
    void Append(const Model& other, const mat4& xf, uint32_t materialSlot);

private:
    std::vector<Vertex> vertices_;
    std::vector<uint32_t> indices_;
    vec3 aabbMin_, aabbMax_;
    void recalcBounds();
};

class Node {
public:
    static std::shared_ptr<Node> CreateNode(std::string name, vec3 translation, quat rotation, vec3 scale, uint32_t modelId, uint32_t instanceId, bool replace);
    Node(std::string name, vec3 translation, quat rotation, vec3 scale, uint32_t id, uint32_t instanceId, bool replace);
    void SetTranslation(vec3 t) { translation_ = t; }
    void SetRotation(quat r) { rotation_ = r; }
    void SetScale(vec3 s) { scaling_ = s; }
    vec3 Translation() const { return translation_; }
    void RecalcLocalTransform();
    void RecalcTransform(bool full = true);
    const mat4& WorldTransform() const { return transform_; }
    uint32_t GetModel() const { return modelId_; }
    const std::string& GetName() const { return name_; }
    void SetVisible(bool v) { visible_ = v; }
    bool IsVisible() const { return visible_; }
    bool IsDrawable() const { return modelId_ != (uint32_t)-1; }
    uint32_t GetInstanceId() const { return instanceId_; }
    bool TickVelocity(mat4& combinedTS);
    bool IsUnsettled() const { return prevTransform_ != transform_; } // the next tick changes combinedPrevTS again
    void SetMaterial(const std::vector<uint32_t>& m);
    const std::array<uint32_t, 16>& Materials() const { return materialIdx_; }
    NodeProxy GetNodeProxy() const;

private:
    std::string name_;
    vec3 translation_;
    quat rotation_;
    vec3 scaling_;
    mat4 localTransform_, transform_, prevTransform_;
    mat4 steadyCombined_;        // prev * inverse(cur) of a node at rest, kept while the transform is unchanged
    bool steadyValid_ = false;
    uint32_t modelId_, instanceId_;
    bool visible_ = true;
    std::array<uint32_t, 16> materialIdx_{};
};

// Storage of the per-frame node proxies: page-locked when a CUDA device is present (the reference
// writes them into a mapped device buffer, Scene.cpp:464-511), ordinary memory otherwise.
void* ProxyAlloc(size_t bytes);
void ProxyFree(void* p);
template <class T> struct ProxyAllocator {
    using value_type = T;
    ProxyAllocator() = default;
    template <class U> ProxyAllocator(const ProxyAllocator<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(ProxyAlloc(n * sizeof(T))); }
    void deallocate(T* p, size_t) { ProxyFree(p); }
    template <class U> bool operator==(const ProxyAllocator<U>&) const { return true; }
    template <class U> bool operator!=(const ProxyAllocator<U>&) const { return false; }
};
using ProxyVector = std::vector<NodeProxy, ProxyAllocator<NodeProxy>>;

class Scene {
public:
    std::vector<std::shared_ptr<Node>>& Nodes() { return nodes_; }
    std::vector<Model>& Models() { return models_; }
    std::vector<FMaterial>& Materials() { return materials_; }
    std::vector<LightObject>& Lights() { return lights_; }
    EnvironmentSetting& GetEnvSettings() { return envSettings_; }
    const Camera& GetRenderCamera() const { return envSettings_.cameras.at(cameraIdx_); }
    uint32_t GetLightCount() const { return (uint32_t)lights_.size(); }
    uint32_t GetSelectedId() const { return selectedId_; }
    void SetSelectedId(uint32_t id) { selectedId_ = id; }
    // Scene::UpdateNodesGpuDriven (Scene.cpp:464-511): one proxy per (drawable node, section).
    bool UpdateNodes();
    const ProxyVector& GetNodeProxys() const { return nodeProxys_; }
    void MarkDirty() { sceneDirty_ = true, fullDirty_ = true; }
    // A caller that knows WHICH node it changed says so: the next UpdateNodes then re-evaluates only the marked nodes and
    // the nodes still settling from the previous tick (their combinedPrevTS changes once more); every other proxy is what
    // the full loop of the reference would write again.  ChangedProxies() lists the records that call rewrote.
    void MarkNodeDirty(uint32_t node) { touched_.push_back(node), sceneDirty_ = true; }
    bool LastUpdateWasFull() const { return lastUpdateFull_; }
    const std::vector<uint32_t>& ChangedProxies() const { return changedProxies_; }
    std::vector<GkMaterial> GpuMaterials() const;

    // flat view for gk_upload_scene (pointers stay valid while the scene is unchanged)
    const GkSceneDesc& Desc();

private:
    std::vector<std::shared_ptr<Node>> nodes_;
    std::vector<Model> models_;
    std::vector<FMaterial> materials_;
    std::vector<LightObject> lights_;
    EnvironmentSetting envSettings_;
    uint32_t cameraIdx_ = 0, selectedId_ = (uint32_t)-1;
    bool sceneDirty_ = true, fullDirty_ = true, lastUpdateFull_ = true;
    std::vector<uint32_t> touched_, settling_, proxyOffset_, changedProxies_;
    ProxyVector nodeProxys_;
    std::vector<GkModelDesc> modelDescs_;
    std::vector<GkMaterial> gpuMaterials_;
    GkSceneDesc desc_{};
};

} // namespace gk::Assets

namespace gk::SceneList {
// Built-in and procedural scenes.  CornellBox follows src/Runtime/SceneList.cpp:186-247 (physics
// is ignored: nodes stay at their frame-0 transforms); the others are the synthetic benchmark
// scenes defined in SURVEY.md §8(d).
void CornellBox(Assets::Scene& scene);
void ProceduralRoom(Assets::Scene& scene, uint32_t targetTriangles = 1000000, uint32_t seed = 1234);
void ProceduralRoomUnique(Assets::Scene& scene, uint32_t targetTriangles, uint32_t seed = 1234); // every object its own model: no instancing
void BrickField(Assets::Scene& scene, uint32_t brickCount = 200000, uint32_t seed = 42);
void BrickFieldStep(Assets::Scene& scene, uint32_t frame, uint32_t seed = 42); // moves 1 % of the bricks
void InstancedCity(Assets::Scene& scene, uint32_t buildingVariants = 40, uint32_t gridSide = 100, uint32_t seed = 7, int facadeN = 46);
} // namespace gk::SceneList
