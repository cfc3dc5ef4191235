// gk_assets.cpp — see gk_assets.h.
#include "gk_assets.h"
#include "../../include/gknext_cuda.h"
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <thread>
#include <new>
#include <random>
#include <unordered_set>

namespace gk::Assets {

static Vertex V(vec3 p, vec3 n, vec2 uv, uint32_t mat)
{
    Vertex v{};
    v.Position[0] = p.x, v.Position[1] = p.y, v.Position[2] = p.z;
    v.Normal[0] = n.x, v.Normal[1] = n.y, v.Normal[2] = n.z;
    v.Tangent[0] = 1, v.Tangent[1] = 0, v.Tangent[2] = 0, v.Tangent[3] = 0;
    v.TexCoord[0] = uv.x, v.TexCoord[1] = uv.y;
    v.MaterialIndex = mat;
    return v;
}

Model::Model(std::vector<Vertex>&& v, std::vector<uint32_t>&& i, bool) : vertices_(std::move(v)), indices_(std::move(i)) { recalcBounds(); }

void Model::recalcBounds()
{
    aabbMin_ = vec3(1e30f), aabbMax_ = vec3(-1e30f);
    for (auto& v : vertices_) {
        vec3 p(v.Position[0], v.Position[1], v.Position[2]);
        aabbMin_ = vmin(aabbMin_, p), aabbMax_ = vmax(aabbMax_, p);
    }
}

uint32_t Model::SectionCount() const
{
    const size_t maxIdx = 65535 * 3;
    size_t s = (indices_.size() + maxIdx - 1) / maxIdx;
    return (uint32_t)std::min<size_t>(s, 10);
}

Model Model::CreateBox(const vec3& p0, const vec3& p1) // Model.cpp:949-999
{
    const vec3 nrm[6] = {{-1, 0, 0}, {1, 0, 0}, {0, 0, -1}, {0, 0, 1}, {0, -1, 0}, {0, 1, 0}};
    const vec3 q[6][4] = {
        {{p0.x, p0.y, p0.z}, {p0.x, p0.y, p1.z}, {p0.x, p1.y, p1.z}, {p0.x, p1.y, p0.z}},
        {{p1.x, p0.y, p1.z}, {p1.x, p0.y, p0.z}, {p1.x, p1.y, p0.z}, {p1.x, p1.y, p1.z}},
        {{p1.x, p0.y, p0.z}, {p0.x, p0.y, p0.z}, {p0.x, p1.y, p0.z}, {p1.x, p1.y, p0.z}},
        {{p0.x, p0.y, p1.z}, {p1.x, p0.y, p1.z}, {p1.x, p1.y, p1.z}, {p0.x, p1.y, p1.z}},
        {{p0.x, p0.y, p0.z}, {p1.x, p0.y, p0.z}, {p1.x, p0.y, p1.z}, {p0.x, p0.y, p1.z}},
        {{p1.x, p1.y, p0.z}, {p0.x, p1.y, p0.z}, {p0.x, p1.y, p1.z}, {p1.x, p1.y, p1.z}},
    };
    std::vector<Vertex> vs;
    std::vector<uint32_t> is;
    for (int f = 0; f < 6; ++f) {
        for (int k = 0; k < 4; ++k) vs.push_back(V(q[f][k], nrm[f], {0, 0}, 0));
        const uint32_t b = f * 4;
        for (uint32_t k : {0u, 1u, 2u, 0u, 2u, 3u}) is.push_back(b + k);
    }
    return Model(std::move(vs), std::move(is), true);
}

Model Model::CreateUVSphere(const vec3& center, float radius, int slices, int stacks) // Model.cpp:1001-1080 with free tessellation
{
    const float pi = 3.14159265358979323846f;
    std::vector<Vertex> vs;
    std::vector<uint32_t> is;
    const float j0d = pi / static_cast<float>(stacks);
    const float i0d = (pi + pi) / static_cast<float>(slices);
    float j0 = 0.f;
    for (int j = 0; j <= stacks; ++j) {
        const float v = radius * -std::sin(j0), z = radius * std::cos(j0);
        const float n0 = -std::sin(j0), n1 = std::cos(j0);
        float i0 = 0;
        for (int i = 0; i <= slices; ++i) {
            vec3 p(center.x + v * std::sin(i0), center.y + z, center.z + v * std::cos(i0));
            vec3 n(n0 * std::sin(i0), n1, n0 * std::cos(i0));
            vs.push_back(V(p, n, {static_cast<float>(i) / slices, static_cast<float>(j) / stacks}, 0));
            i0 += i0d;
        }
        j0 += j0d;
    }
    const int s1 = slices + 1;
    int r0 = 0, r1 = s1;
    for (int j = 0; j < stacks; ++j) {
        for (int i = 0; i < slices; ++i) {
            is.push_back(r0 + i), is.push_back(r1 + i), is.push_back(r1 + i + 1);
            is.push_back(r0 + i), is.push_back(r1 + i + 1), is.push_back(r0 + i + 1);
        }
        r0 += s1, r1 += s1;
    }
    return Model(std::move(vs), std::move(is), true);
}

Model Model::CreateSphere(const vec3& center, float radius) { return CreateUVSphere(center, radius, 32, 16); }

// A box whose faces are n x n quads with a seeded outward bump: a cheap stand-in for a
// detailed mesh of 12*n*n triangles.
Model Model::CreateGridBox(const vec3& p0, const vec3& p1, int n, float bump, uint32_t seed)
{
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::vector<Vertex> vs;
    std::vector<uint32_t> is;
    const vec3 c = (p0 + p1) * 0.5f, h = (p1 - p0) * 0.5f;
    // face frames: normal, u axis, v axis (u x v = normal)
    const vec3 N[6] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    const vec3 Ux[6] = {{0, 1, 0}, {0, 0, 1}, {0, 0, 1}, {1, 0, 0}, {1, 0, 0}, {0, 1, 0}};
    const vec3 Vx[6] = {{0, 0, 1}, {0, 1, 0}, {1, 0, 0}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}};
    for (int f = 0; f < 6; ++f) {
        const uint32_t base = (uint32_t)vs.size();
        for (int j = 0; j <= n; ++j)
            for (int i = 0; i <= n; ++i) {
                float u = -1.f + 2.f * i / n, v = -1.f + 2.f * j / n;
                const bool border = (i == 0 || j == 0 || i == n || j == n);
                float d = border ? 0.f : bump * U(rng);
                vec3 p = c + (N[f] * (1.f + d) + Ux[f] * u + Vx[f] * v) * h;
                vs.push_back(V(p, N[f], {float(i) / n, float(j) / n}, 0));
            }
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                uint32_t a = base + j * (n + 1) + i, b = a + 1, d2 = a + (n + 1), e = d2 + 1;
                is.push_back(a), is.push_back(b), is.push_back(e);
                is.push_back(a), is.push_back(e), is.push_back(d2);
            }
    }
    return Model(std::move(vs), std::move(is), true);
}

void Model::Append(const Model& o, const mat4& xf, uint32_t slot)
{
    const uint32_t base = (uint32_t)vertices_.size();
    for (auto v : o.vertices_) {
        vec4 p = xf * vec4(vec3(v.Position[0], v.Position[1], v.Position[2]), 1.f);
        vec4 n = xf * vec4(vec3(v.Normal[0], v.Normal[1], v.Normal[2]), 0.f);
        vec3 nn = normalize(vec3(n.x, n.y, n.z));
        v.Position[0] = p.x, v.Position[1] = p.y, v.Position[2] = p.z;
        v.Normal[0] = nn.x, v.Normal[1] = nn.y, v.Normal[2] = nn.z;
        v.MaterialIndex = slot;
        vertices_.push_back(v);
    }
    for (auto i : o.indices_) indices_.push_back(base + i);
    recalcBounds();
}

// CornellBox::Create, src/Assets/CornellBox.cpp:18-151
uint32_t Model::CreateCornellBox(float s, std::vector<Model>& models, std::vector<FMaterial>& materials, std::vector<LightObject>& lights)
{
    const uint32_t prev = (uint32_t)materials.size();
    materials.push_back({"", prev + 0, Material::Lambertian(vec3(0.65f, 0.05f, 0.05f))});
    materials.push_back({"", prev + 1, Material::Lambertian(vec3(0.12f, 0.45f, 0.15f))});
    materials.push_back({"", prev + 2, Material::Lambertian(vec3(0.73f, 0.73f, 0.73f))});
    materials.push_back({"", prev + 3, Material::DiffuseLight(vec3(2000.0f))});

    std::vector<Vertex> vs;
    std::vector<uint32_t> is;
    const vec3 off(s * 0.5, 0, -s * 0.5);
    const vec3 l0(0, 0, 0), l1(0, 0, -s), l2(0, s, -s), l3(0, s, 0);
    const vec3 r0(s, 0, 0), r1(s, 0, -s), r2(s, s, -s), r3(s, s, 0);
    const vec2 uv[4] = {{0, 1}, {1, 1}, {1, 0}, {0, 0}};
    auto quad = [&](vec3 a, vec3 b, vec3 c, vec3 d, vec3 n, uint32_t mat, bool flip) {
        const uint32_t i = (uint32_t)vs.size();
        vs.push_back(V(a - off, n, uv[0], mat)), vs.push_back(V(b - off, n, uv[1], mat)), vs.push_back(V(c - off, n, uv[2], mat)), vs.push_back(V(d - off, n, uv[3], mat));
        if (!flip) for (uint32_t k : {0u, 1u, 2u, 0u, 2u, 3u}) is.push_back(i + k);
        else for (uint32_t k : {2u, 1u, 0u, 3u, 2u, 0u}) is.push_back(i + k);
    };
    quad(l0, l1, l2, l3, vec3(1, 0, 0), 1, false);  // left, green
    quad(r0, r1, r2, r3, vec3(-1, 0, 0), 0, true);  // right, red
    quad(l1, r1, r2, l2, vec3(0, 0, 1), 2, false);  // back
    quad(l0, r0, r1, l1, vec3(0, 1, 0), 2, false);  // floor
    quad(l2, r2, r3, l3, vec3(0, -1, 0), 2, false); // ceiling
    {
        const float x0 = s * (163.0f / 555.0f), x1 = s * (393.0f / 555.0f);
        const float z0 = s * (-555.0f + 432.0f) / 555.0f, z1 = s * (-555.0f + 202.0f) / 555.0f;
        const float y1 = s * 0.999f;
        quad(vec3(x0, y1, z1), vec3(x1, y1, z1), vec3(x1, y1, z0), vec3(x0, y1, z0), vec3(0, -1, 0), 3, false);
        LightObject light{};
        auto set = [](float* d, vec3 v, float w) { d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = w; };
        set(light.p0, vec3(x0, y1, z1) - off, 1);
        set(light.p1, vec3(x0, y1, z0) - off, 1);
        set(light.p3, vec3(x1, y1, z1) - off, 1);
        set(light.normal_area, vec3(0, -1, 0), (x1 - x0) * (z0 - z1));
        light.lightMatIdx = prev + 3;
        lights.push_back(light);
    }
    models.push_back(Model(std::move(vs), std::move(is), true));
    return (uint32_t)models.size() - 1;
}

std::shared_ptr<Node> Node::CreateNode(std::string name, vec3 t, quat r, vec3 s, uint32_t modelId, uint32_t instanceId, bool replace)
{
    return std::make_shared<Node>(name, t, r, s, modelId, instanceId, replace);
}

Node::Node(std::string name, vec3 t, quat r, vec3 s, uint32_t id, uint32_t instanceId, bool replace)
    : name_(name), translation_(t), rotation_(r), scaling_(s), modelId_(id), instanceId_(instanceId), visible_(false)
{
    RecalcLocalTransform();
    RecalcTransform();
    prevTransform_ = replace ? transform_ : translate(vec3(0, -100, 0)); // Model.cpp:1351-1358
}

void Node::RecalcLocalTransform() { localTransform_ = translate(translation_) * mat4_cast(rotation_) * scale(scaling_); } // Model.cpp:1254
void Node::RecalcTransform(bool) { RecalcLocalTransform(); transform_ = localTransform_; }                               // no parent chains here

bool Node::TickVelocity(mat4& combinedTS) // Model.cpp:1279-1296
{
    // same arithmetic as the reference every time it is evaluated; a node at rest (prev == cur, both
    // unchanged since the last tick) re-uses the matrix it produced then
    const bool atRest = !(prevTransform_ != transform_);
    if (atRest && steadyValid_) {
        combinedTS = steadyCombined_;
    } else {
        combinedTS = prevTransform_ * inverse(transform_);
        steadyValid_ = atRest;
        if (atRest) steadyCombined_ = combinedTS;
    }
    prevTransform_ = transform_;
    vec4 p = combinedTS * vec4(0, 0, 0, 1);
    return (p.x * p.x + p.y * p.y + p.z * p.z) > 0.1f;
}

void Node::SetMaterial(const std::vector<uint32_t>& m)
{
    materialIdx_.fill(0);
    for (size_t i = 0; i < m.size() && i < 16; ++i) materialIdx_[i] = m[i];
}

static std::mutex gProxyMutex;
static std::unordered_set<void*> gPinned;

void* ProxyAlloc(size_t bytes)
{
    if (void* p = gk_host_alloc(bytes)) {
        std::lock_guard<std::mutex> lock(gProxyMutex);
        gPinned.insert(p);
        return p;
    }
    void* p = malloc(bytes);
    if (!p) throw std::bad_alloc();
    return p;
}

void ProxyFree(void* p)
{
    bool pinned;
    {
        std::lock_guard<std::mutex> lock(gProxyMutex);
        pinned = gPinned.erase(p) != 0;
    }
    if (pinned) gk_host_free(p);
    else free(p);
}

NodeProxy Node::GetNodeProxy() const // Model.cpp:1326-1342
{
    NodeProxy p{};
    p.instanceId = instanceId_;
    p.modelId = modelId_;
    memcpy(p.worldTS, transform_.data(), 64);
    p.visible = visible_ ? 1 : 0;
    for (int i = 0; i < 16; ++i) p.matId[i] = materialIdx_[i];
    return p;
}

bool Scene::UpdateNodes() // Scene.cpp:464-511
{
    if (nodes_.empty() || !sceneDirty_) return false;
    sceneDirty_ = false;
    if (!fullDirty_ && proxyOffset_.size() == nodes_.size() + 1) {
        // ---- incremental: only the marked nodes and the ones still settling (see MarkNodeDirty)
        std::vector<uint32_t> set(touched_);
        set.insert(set.end(), settling_.begin(), settling_.end());
        std::sort(set.begin(), set.end());
        set.erase(std::unique(set.begin(), set.end()), set.end());
        touched_.clear(), settling_.clear(), changedProxies_.clear();
        bool moved = false, structural = false;
        auto sections = [&](uint32_t i) {
            const auto& node = nodes_[i];
            return (node->IsDrawable() && node->GetModel() < models_.size()) ? models_[node->GetModel()].SectionCount() : 0u;
        };
        for (uint32_t i : set) // a node whose proxy count changed reshapes the list: full pass below, before anything ticks
            if (i < nodes_.size() && proxyOffset_[i + 1] - proxyOffset_[i] != sections(i)) structural = true;
        for (size_t k = 0; k < set.size(); ++k) {
            const uint32_t i = set[k];
            if (structural || i >= nodes_.size()) continue;
            if (k + 8 < set.size() && set[k + 8] < nodes_.size()) { // the nodes are scattered heap objects: hide the misses
                const char* ahead = reinterpret_cast<const char*>(nodes_[set[k + 8]].get());
                for (int line = 0; line < 8; ++line) __builtin_prefetch(ahead + 64 * line);
                __builtin_prefetch(&nodeProxys_[proxyOffset_[set[k + 8]]]);
            }
            auto& node = nodes_[i];
            const uint32_t want = sections(i);
            if (!node->IsDrawable()) continue;
            const bool unsettled = node->IsUnsettled();
            mat4 combined;
            if (node->TickVelocity(combined)) moved = true;
            if (unsettled) settling_.push_back(i);
            for (uint32_t section = 0; section < want; ++section) {
                NodeProxy& proxy = nodeProxys_[proxyOffset_[i] + section];
                proxy = node->GetNodeProxy();
                memcpy(proxy.combinedPrevTS, combined.data(), 64);
                proxy.modelId = node->GetModel() * 10 + section;
                proxy.nort = section == 0 ? 0 : 1;
                changedProxies_.push_back(proxyOffset_[i] + section);
            }
        }
        if (!structural) {
            lastUpdateFull_ = false;
            if (moved) sceneDirty_ = true; // Scene.cpp:509: a moving node keeps the scene dirty for the next frame
            return true;
        }
        settling_.clear();
    }
    fullDirty_ = false, lastUpdateFull_ = true;
    touched_.clear(), settling_.clear(), changedProxies_.clear();
    // The proxy of a node depends on that node only, so large scenes are filled by several threads:
    // pass 1 counts the proxies of every node (drawable, valid model, one per section), pass 2 writes
    // them at their prefix offsets.  The result is identical to the sequential loop of the reference.
    const size_t n = nodes_.size();
    std::vector<uint32_t> offset(n + 1, 0);
    for (size_t i = 0; i < n; ++i) {
        const auto& node = nodes_[i];
        uint32_t c = 0;
        if (node->IsDrawable() && node->GetModel() < models_.size()) c = models_[node->GetModel()].SectionCount();
        offset[i + 1] = offset[i] + c;
    }
    nodeProxys_.resize(offset[n]);
    proxyOffset_ = offset;
    std::atomic<bool> moved{false};
    std::mutex settleMutex;
    auto fill = [&](size_t lo, size_t hi) {
        std::vector<uint32_t> unsettledHere;
        for (size_t i = lo; i < hi; ++i) {
            auto& node = nodes_[i];
            if (!node->IsDrawable()) continue;
            if (node->IsUnsettled()) unsettledHere.push_back((uint32_t)i);
            mat4 combined;
            if (node->TickVelocity(combined)) moved.store(true, std::memory_order_relaxed);
            if (node->GetModel() >= models_.size()) continue;
            const Model& model = models_[node->GetModel()];
            for (uint32_t section = 0; section < model.SectionCount(); ++section) {
                NodeProxy& proxy = nodeProxys_[offset[i] + section];
                proxy = node->GetNodeProxy();
                memcpy(proxy.combinedPrevTS, combined.data(), 64);
                proxy.modelId = node->GetModel() * 10 + section;
                proxy.nort = section == 0 ? 0 : 1;
            }
        }
        if (!unsettledHere.empty()) {
            std::lock_guard<std::mutex> lock(settleMutex);
            settling_.insert(settling_.end(), unsettledHere.begin(), unsettledHere.end());
        }
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned threads = n < 16384 ? 1u : std::min(8u, hw);
    if (threads == 1) fill(0, n);
    else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < threads; ++t) pool.emplace_back(fill, n * t / threads, n * (t + 1) / threads);
        for (auto& th : pool) th.join();
    }
    if (moved.load()) sceneDirty_ = true; // Scene.cpp:509 (the next pass only has to settle the nodes that moved)
    return true;
}

std::vector<GkMaterial> Scene::GpuMaterials() const
{
    std::vector<GkMaterial> r;
    for (auto& m : materials_) r.push_back(m.gpuMaterial_);
    return r;
}

const GkSceneDesc& Scene::Desc()
{
    modelDescs_.clear();
    for (auto& m : models_) modelDescs_.push_back({m.CPUVertices().data(), m.CPUIndices().data(), m.NumberOfVertices(), m.NumberOfIndices()});
    gpuMaterials_ = GpuMaterials();
    desc_.models = modelDescs_.data();
    desc_.modelCount = (uint32_t)modelDescs_.size();
    desc_.materials = gpuMaterials_.data();
    desc_.materialCount = (uint32_t)gpuMaterials_.size();
    desc_.lights = lights_.data();
    desc_.lightCount = (uint32_t)lights_.size();
    return desc_;
}

} // namespace gk::Assets

namespace gk::SceneList {

using namespace gk::Assets;

static std::shared_ptr<Node> addNode(Scene& s, const std::string& name, vec3 t, quat r, vec3 sc, uint32_t model, std::vector<uint32_t> mats)
{
    auto n = Node::CreateNode(name, t, r, sc, model, (uint32_t)s.Nodes().size(), false);
    n->SetVisible(true);
    n->SetMaterial(mats);
    s.Nodes().push_back(n);
    return n;
}

void CornellBox(Scene& scene) // SceneList.cpp:186-247
{
    auto& env = scene.GetEnvSettings();
    auto& materials = scene.Materials();
    const uint32_t prev = (uint32_t)materials.size();
    Camera cam;
    cam.name = "Cam";
    cam.ModelView = lookAt(vec3(0, 2.78, 10.78), vec3(0, 2.78, 0), vec3(0, 1, 0));
    cam.FieldOfView = 40, cam.Aperture = 0, cam.FocalDistance = 10;
    env.cameras.push_back(cam);
    env.ControlSpeed = 200.0f, env.GammaCorrection = true, env.HasSky = false, env.HasSun = false;

    const uint32_t cbox = Model::CreateCornellBox(5.55f, scene.Models(), materials, scene.Lights());
    addNode(scene, "cbox", vec3(0, 0, 0), quat(1, 0, 0, 0), vec3(1, 1, 1), cbox, {prev + 0, prev + 1, prev + 2, prev + 3});
    const vec3 spherePos(1.30, 1.01 + 2.00 * 0.0, 0.80), boxPos(-1.30, 0, -0.80);
    materials.push_back({"cbox_white", prev + 4, Material::Lambertian(vec3(0.73f, 0.73f, 0.73f))});
    materials.push_back({"cball_white", prev + 5, Material::Mixture(vec3(0.73f, 0.73f, 0.73f), 0.01f)});
    scene.Models().push_back(Model::CreateBox(vec3(-0.80, 0, -0.80), vec3(0.80, 1.60, 0.80)));
    scene.Models().push_back(Model::CreateSphere(vec3(0, 0, 0), 1.0f));
    addNode(scene, "Sphere1", spherePos, quat(vec3(0, 0.5f, 0)), vec3(1, 1, 1), cbox + 2, {prev + 5});
    addNode(scene, "Box", boxPos, quat(vec3(0, 0.25f, 0)), vec3(1, 2, 1), cbox + 1, {prev + 4});
}

// ---- synthetic benchmark scenes ----

static uint32_t addMaterialPalette(Scene& scene, std::mt19937& rng, int count)
{
    std::uniform_real_distribution<float> U(0.f, 1.f);
    const uint32_t first = (uint32_t)scene.Materials().size();
    for (int i = 0; i < count; ++i) {
        const float pick = U(rng);
        const vec3 col(0.15f + 0.8f * U(rng), 0.15f + 0.8f * U(rng), 0.15f + 0.8f * U(rng));
        Material m;
        if (pick < 0.60f) m = Material::Lambertian(col);
        else if (pick < 0.80f) m = Material::Mixture(col, 0.5f * U(rng));
        else if (pick < 0.90f) m = Material::Metallic(col, 0.3f * U(rng));
        else if (pick < 0.95f) m = Material::Dielectric(1.5f, 0.0f);
        else m = Material::DiffuseLight(col * 20.f);
        scene.Materials().push_back({"", first + i, m});
    }
    return first;
}

void ProceduralRoom(Scene& scene, uint32_t targetTriangles, uint32_t seed)
{
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    auto& env = scene.GetEnvSettings();
    Camera cam;
    cam.name = "Cam";
    cam.ModelView = lookAt(vec3(-9.0, 3.2, 9.0), vec3(0.0, 1.2, 0.0), vec3(0, 1, 0));
    cam.FieldOfView = 40, cam.Aperture = 0, cam.FocalDistance = 10;
    env.cameras.push_back(cam);
    env.HasSky = true, env.HasSun = true, env.SkyIntensity = 1.0f, env.SunIntensity = 20.f, env.SunRotation = 0.35f;

    scene.Materials().push_back({"floor", 0, Material::Lambertian(vec3(0.7f, 0.7f, 0.7f))});
    scene.Materials().push_back({"wall", 1, Material::Lambertian(vec3(0.6f, 0.55f, 0.5f))});
    const uint32_t pal = addMaterialPalette(scene, rng, 64);

    // room shell: floor + four walls, open to the sky (20 x 4 x 20 m)
    Model shell = Model::CreateBox(vec3(-10, -0.2f, -10), vec3(10, 0, 10));
    // no two faces of the shell are coplanar (coplanar overlaps are distance ties by construction)
    Model wallZ = Model::CreateBox(vec3(-9.97f, -0.1f, -0.1f), vec3(9.97f, 4, 0.1f));
    Model wallX = Model::CreateBox(vec3(-0.1f, -0.13f, -10.23f), vec3(0.1f, 4.05f, 10.23f));
    shell.Append(wallZ, translate(vec3(0, 0, -10.1f)), 1);
    shell.Append(wallZ, translate(vec3(0, 0, 10.1f)), 1);
    shell.Append(wallX, translate(vec3(-10.1f, 0, 0)), 1);
    shell.Append(wallX, translate(vec3(10.1f, 0, 0)), 1);
    scene.Models().push_back(shell);
    addNode(scene, "shell", vec3(0, 0, 0), quat(1, 0, 0, 0), vec3(1, 1, 1), 0, {0, 1});

    const uint32_t base = (uint32_t)scene.Models().size();
    scene.Models().push_back(Model::CreateBox(vec3(-0.5f, -0.5f, -0.5f), vec3(0.5f, 0.5f, 0.5f)));      // 12
    scene.Models().push_back(Model::CreateSphere(vec3(0, 0, 0), 0.5f));                                  // 1024
    scene.Models().push_back(Model::CreateUVSphere(vec3(0, 0, 0), 0.5f, 48, 24));                        // 2304
    scene.Models().push_back(Model::CreateUVSphere(vec3(0, 0, 0), 0.5f, 16, 8));                         // 256
    scene.Models().push_back(Model::CreateGridBox(vec3(-0.5f, -0.5f, -0.5f), vec3(0.5f, 0.5f, 0.5f), 6, 0.08f, seed + 1)); // 432
    uint64_t tris = scene.Models()[0].NumberOfIndices() / 3;
    while (tris < targetTriangles) {
        const uint32_t m = base + (uint32_t)(U(rng) * 5.f) % 5;
        const vec3 t(-9.5f + 19.f * U(rng), 0.3f + 3.4f * U(rng), -9.5f + 19.f * U(rng));
        const quat r(vec3(6.2831853f * U(rng), 6.2831853f * U(rng), 6.2831853f * U(rng)));
        const float sc = 0.25f + 0.75f * U(rng);
        addNode(scene, "obj", t, r, vec3(sc, sc * (0.6f + 0.8f * U(rng)), sc), m, {pal + (uint32_t)(U(rng) * 64.f) % 64});
        tris += scene.Models()[m].NumberOfIndices() / 3;
    }
}

// The same room with UNIQUE geometry: every object is its own model (its own BLAS), nothing is instanced, so the
// 1 M triangles are 1 M distinct triangle records (~50 MB of geometry + BVH) instead of six meshes that live in L1.
// Kitchen / LivingRoom-class: the reference's gallery scenes are single-use meshes (assets/models, not in this tree).
void ProceduralRoomUnique(Scene& scene, uint32_t targetTriangles, uint32_t seed)
{
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    auto& env = scene.GetEnvSettings();
    Camera cam;
    cam.name = "Cam";
    cam.ModelView = lookAt(vec3(-9.0, 3.2, 9.0), vec3(0.0, 1.2, 0.0), vec3(0, 1, 0));
    cam.FieldOfView = 40, cam.Aperture = 0, cam.FocalDistance = 10;
    env.cameras.push_back(cam);
    env.HasSky = true, env.HasSun = true, env.SkyIntensity = 1.0f, env.SunIntensity = 20.f, env.SunRotation = 0.35f;

    scene.Materials().push_back({"floor", 0, Material::Lambertian(vec3(0.7f, 0.7f, 0.7f))});
    scene.Materials().push_back({"wall", 1, Material::Lambertian(vec3(0.6f, 0.55f, 0.5f))});
    const uint32_t pal = addMaterialPalette(scene, rng, 64);

    Model shell = Model::CreateBox(vec3(-10, -0.2f, -10), vec3(10, 0, 10));
    Model wallZ = Model::CreateBox(vec3(-9.97f, -0.1f, -0.1f), vec3(9.97f, 4, 0.1f));
    Model wallX = Model::CreateBox(vec3(-0.1f, -0.13f, -10.23f), vec3(0.1f, 4.05f, 10.23f));
    shell.Append(wallZ, translate(vec3(0, 0, -10.1f)), 1);
    shell.Append(wallZ, translate(vec3(0, 0, 10.1f)), 1);
    shell.Append(wallX, translate(vec3(-10.1f, 0, 0)), 1);
    shell.Append(wallX, translate(vec3(10.1f, 0, 0)), 1);
    scene.Models().push_back(shell);
    addNode(scene, "shell", vec3(0, 0, 0), quat(1, 0, 0, 0), vec3(1, 1, 1), 0, {0, 1});

    uint64_t tris = scene.Models()[0].NumberOfIndices() / 3;
    uint32_t k = 0;
    while (tris < targetTriangles) {
        // bumpy boxes (12 n^2 triangles, own random surface) and tessellated spheres of varying resolution
        Model m = (k % 3 == 2) ? Model::CreateUVSphere(vec3(0, 0, 0), 0.5f, 40 + (int)(U(rng) * 24.f), 20 + (int)(U(rng) * 12.f))
                               : Model::CreateGridBox(vec3(-0.5f, -0.5f, -0.5f), vec3(0.5f, 0.5f, 0.5f), 10 + (int)(U(rng) * 8.f), 0.08f, seed * 977u + k);
        const uint32_t id = (uint32_t)scene.Models().size();
        tris += m.NumberOfIndices() / 3;
        scene.Models().push_back(std::move(m));
        const vec3 t(-9.5f + 19.f * U(rng), 0.3f + 3.4f * U(rng), -9.5f + 19.f * U(rng));
        const quat r(vec3(6.2831853f * U(rng), 6.2831853f * U(rng), 6.2831853f * U(rng)));
        const float sc = 0.35f + 0.9f * U(rng);
        addNode(scene, "obj", t, r, vec3(sc, sc * (0.6f + 0.8f * U(rng)), sc), id, {pal + (uint32_t)(U(rng) * 64.f) % 64});
        ++k;
    }
}

static Model makeBrick(int nx, int nz)
{
    // LEGO-like brick on the reference's lattice (src/MagicaLego/MagicaLegoGameInstance.cpp:38-45):
    // 0.08 x 0.095 x 0.08 m per cell, one 12-gon stud per cell.
    const float cx = 0.08f, cy = 0.095f, cz = 0.08f;
    Model brick = Model::CreateBox(vec3(0, 0, 0), vec3(cx * nx, cy, cz * nz));
    std::vector<Vertex> vs;
    std::vector<uint32_t> is;
    const int seg = 12;
    const float r = 0.024f, hgt = 0.017f, pi = 3.14159265358979323846f;
    for (int ix = 0; ix < nx; ++ix)
        for (int iz = 0; iz < nz; ++iz) {
            const vec3 c(cx * (ix + 0.5f), cy, cz * (iz + 0.5f));
            const uint32_t b = (uint32_t)vs.size();
            for (int k = 0; k < seg; ++k) {
                const float a = 2.f * pi * k / seg;
                const vec3 n(cosf(a), 0, sinf(a));
                vs.push_back(V(c + n * r, n, {0, 0}, 0));
                vs.push_back(V(c + n * r + vec3(0, hgt, 0), n, {0, 0}, 0));
            }
            const uint32_t top = (uint32_t)vs.size();
            vs.push_back(V(c + vec3(0, hgt, 0), vec3(0, 1, 0), {0, 0}, 0));
            for (int k = 0; k < seg; ++k) {
                const uint32_t a0 = b + 2 * k, a1 = a0 + 1, b0 = b + 2 * ((k + 1) % seg), b1 = b0 + 1;
                is.push_back(a0), is.push_back(a1), is.push_back(b1);
                is.push_back(a0), is.push_back(b1), is.push_back(b0);
                is.push_back(top), is.push_back(b1), is.push_back(a1);
            }
        }
    Model studs(std::move(vs), std::move(is), true);
    brick.Append(studs, mat4(), 0);
    return brick;
}

void BrickField(Scene& scene, uint32_t brickCount, uint32_t seed)
{
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    auto& env = scene.GetEnvSettings();
    Camera cam;
    cam.name = "Cam";
    cam.ModelView = lookAt(vec3(-14.0, 9.0, 14.0), vec3(0.0, 0.5, 0.0), vec3(0, 1, 0));
    cam.FieldOfView = 40, cam.Aperture = 0, cam.FocalDistance = 10;
    env.cameras.push_back(cam);
    env.HasSky = true, env.HasSun = true, env.SkyIntensity = 1.0f, env.SunIntensity = 20.f, env.SunRotation = 0.3f;

    const uint32_t pal = addMaterialPalette(scene, rng, 32);
    const int dims[8][2] = {{1, 1}, {1, 2}, {1, 4}, {2, 2}, {2, 3}, {2, 4}, {1, 6}, {2, 6}};
    for (auto& d : dims) scene.Models().push_back(makeBrick(d[0], d[1]));
    scene.Models().push_back(Model::CreateBox(vec3(-25, -0.1f, -25), vec3(25, 0, 25)));
    scene.Materials().push_back({"ground", (uint32_t)scene.Materials().size(), Material::Lambertian(vec3(0.5f, 0.5f, 0.5f))});
    addNode(scene, "ground", vec3(0, 0, 0), quat(1, 0, 0, 0), vec3(1, 1, 1), 8, {(uint32_t)scene.Materials().size() - 1});
    const int G = 512; // lattice cells per side
    for (uint32_t i = 0; i < brickCount; ++i) {
        const uint32_t m = (uint32_t)(U(rng) * 8.f) % 8;
        const int gx = (int)(U(rng) * G) - G / 2, gz = (int)(U(rng) * G) - G / 2, gy = (int)(U(rng) * U(rng) * 24.f);
        addNode(scene, "brick", vec3(gx * 0.08f, gy * 0.095f, gz * 0.08f), quat(1, 0, 0, 0), vec3(1, 1, 1), m, {pal + (uint32_t)(U(rng) * 32.f) % 32});
    }
}

void BrickFieldStep(Scene& scene, uint32_t frame, uint32_t seed)
{
    std::mt19937 rng(seed * 7919u + frame);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    auto& nodes = scene.Nodes();
    const size_t n = nodes.size() - 1;
    const size_t moves = n / 100;
    const int G = 512;
    for (size_t k = 0; k < moves; ++k) {
        const size_t i = 1 + (size_t)(U(rng) * n) % n;
        const int gx = (int)(U(rng) * G) - G / 2, gz = (int)(U(rng) * G) - G / 2, gy = (int)(U(rng) * U(rng) * 24.f);
        nodes[i]->SetTranslation(vec3(gx * 0.08f, gy * 0.095f, gz * 0.08f));
        nodes[i]->RecalcTransform(true);
        scene.MarkNodeDirty((uint32_t)i);
    }
}

void InstancedCity(Scene& scene, uint32_t variants, uint32_t gridSide, uint32_t seed, int facadeN)
{
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    auto& env = scene.GetEnvSettings();
    Camera cam;
    cam.name = "Cam";
    const float span = gridSide * 12.f;
    cam.ModelView = lookAt(vec3(-0.45f * span, 0.18f * span, 0.45f * span), vec3(0.0, 10.0, 0.0), vec3(0, 1, 0));
    cam.FieldOfView = 40, cam.Aperture = 0, cam.FocalDistance = 10;
    env.cameras.push_back(cam);
    env.HasSky = true, env.HasSun = true, env.SkyIntensity = 1.0f, env.SunIntensity = 20.f, env.SunRotation = 0.3f;

    const uint32_t pal = addMaterialPalette(scene, rng, 64);
    for (uint32_t v = 0; v < variants; ++v) {
        const float w = 3.f + 2.f * U(rng), d = 3.f + 2.f * U(rng), h = 8.f + 40.f * U(rng) * U(rng);
        scene.Models().push_back(Model::CreateGridBox(vec3(-w, 0, -d), vec3(w, h, d), facadeN, 0.04f, seed * 131u + v)); // 12*n*n tris
    }
    scene.Models().push_back(Model::CreateBox(vec3(-span, -0.5f, -span), vec3(span, 0, span)));
    scene.Materials().push_back({"ground", (uint32_t)scene.Materials().size(), Material::Lambertian(vec3(0.4f, 0.4f, 0.42f))});
    addNode(scene, "ground", vec3(0, 0, 0), quat(1, 0, 0, 0), vec3(1, 1, 1), variants, {(uint32_t)scene.Materials().size() - 1});
    for (uint32_t gz = 0; gz < gridSide; ++gz)
        for (uint32_t gx = 0; gx < gridSide; ++gx) {
            const uint32_t m = (uint32_t)(U(rng) * variants) % variants;
            const vec3 t((gx - gridSide * 0.5f + 0.5f) * 12.f + 2.f * (U(rng) - 0.5f), 0, (gz - gridSide * 0.5f + 0.5f) * 12.f + 2.f * (U(rng) - 0.5f));
            addNode(scene, "bldg", t, quat(vec3(0, 6.2831853f * U(rng), 0)), vec3(1, 0.6f + 0.8f * U(rng), 1), m, {pal + (uint32_t)(U(rng) * 64.f) % 64});
        }
}

} // namespace gk::SceneList
