// gk_host_capi.cpp — flat C entry points over the host mirror, so that the Python test and
// benchmark harness (ctypes) can drive the same C++ objects a host application would:
// build a scene with Assets::*, fill the UBO with EngineMirror, and run frames through
// CudaPathTracingRenderer (LogicRendererBase's five virtuals).
#include "gk_engine.h"
#include <cstring>
#include <string>

using namespace gk;

static thread_local std::string g_err;

struct HostRenderer {
    CudaPathTracingRenderer r;
    HostRenderer(EngineMirror& e, int device) : r(e, device) {}
};

#define GKH_TRY(stmt)                 \
    try {                             \
        stmt;                         \
        return 0;                     \
    } catch (const std::exception& e) { \
        g_err = e.what();             \
        return -1;                    \
    }

extern "C" {

const char* gkh_last_error() { return g_err.c_str(); }

// scene names: "cornell", "room" / "roomu" (instanced / unique geometry; p0 = target triangles, p1 = seed), "bricks" (p0 = count, p1 = seed),
// "city" (p0 = variants, p1 = grid side, p2 = seed, p3 = facade subdivisions), "empty"
void* gkh_engine_create(const char* sceneName, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3)
{
    try {
        EngineMirror* e = new EngineMirror();
        const std::string n = sceneName ? sceneName : "";
        if (n == "cornell") SceneList::CornellBox(e->scene);
        else if (n == "room") SceneList::ProceduralRoom(e->scene, p0 ? p0 : 1000000u, p1 ? p1 : 1234u);
        else if (n == "roomu") SceneList::ProceduralRoomUnique(e->scene, p0 ? p0 : 1000000u, p1 ? p1 : 1234u);
        else if (n == "bricks") SceneList::BrickField(e->scene, p0 ? p0 : 200000u, p1 ? p1 : 42u);
        else if (n == "city") SceneList::InstancedCity(e->scene, p0 ? p0 : 40u, p1 ? p1 : 100u, p2 ? p2 : 7u, p3 ? (int)p3 : 46);
        else if (n == "empty") {
            Assets::Camera cam;
            cam.ModelView = lookAt(vec3(0, 0, 5), vec3(0, 0, 0), vec3(0, 1, 0));
            e->scene.GetEnvSettings().cameras.push_back(cam);
        } else {
            delete e;
            g_err = "unknown scene '" + n + "'";
            return nullptr;
        }
        return e;
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return nullptr;
    }
}
void gkh_engine_destroy(void* h) { delete (EngineMirror*)h; }

const GkSceneDesc* gkh_scene_desc(void* h) { return &((EngineMirror*)h)->scene.Desc(); }
uint64_t gkh_scene_triangles(void* h, int instanced)
{
    auto& s = ((EngineMirror*)h)->scene;
    uint64_t t = 0;
    if (!instanced) for (auto& m : s.Models()) t += m.NumberOfIndices() / 3;
    else for (auto& n : s.Nodes()) if (n->IsDrawable()) t += s.Models()[n->GetModel()].NumberOfIndices() / 3;
    return t;
}
// Scene::UpdateNodes; returns the proxy count (the list is rebuilt only when the scene is dirty)
uint32_t gkh_update_nodes(void* h)
{
    auto& s = ((EngineMirror*)h)->scene;
    s.UpdateNodes();
    return (uint32_t)s.GetNodeProxys().size();
}
const GkNodeProxy* gkh_node_proxies(void* h) { return ((EngineMirror*)h)->scene.GetNodeProxys().data(); }
// what the last gkh_update_nodes rewrote: returns -1 when it was a full pass, else the number of proxy indices (see Scene::MarkNodeDirty)
int64_t gkh_changed_proxies(void* h, const uint32_t** indices)
{
    auto& s = ((EngineMirror*)h)->scene;
    if (s.LastUpdateWasFull()) return -1;
    *indices = s.ChangedProxies().data();
    return (int64_t)s.ChangedProxies().size();
}
void gkh_mark_dirty(void* h) { ((EngineMirror*)h)->scene.MarkDirty(); }
void gkh_scene_step(void* h, uint32_t frame) { SceneList::BrickFieldStep(((EngineMirror*)h)->scene, frame); }
int gkh_set_node_translation(void* h, uint32_t node, float x, float y, float z)
{
    auto& s = ((EngineMirror*)h)->scene;
    if (node >= s.Nodes().size()) return -1;
    s.Nodes()[node]->SetTranslation(vec3(x, y, z));
    s.Nodes()[node]->RecalcTransform(true);
    s.MarkNodeDirty(node);
    return 0;
}

int gkh_set_setting(void* h, const char* name, double v)
{
    EngineMirror* e = (EngineMirror*)h;
    UserSettings& u = e->userSettings;
    auto& env = e->scene.GetEnvSettings();
    const std::string n = name;
    if (n == "NumberOfSamples") u.NumberOfSamples = (int)v;
    else if (n == "NumberOfBounces") u.NumberOfBounces = (int)v;
    else if (n == "MaxNumberOfBounces") u.MaxNumberOfBounces = (int)v;
    else if (n == "TAA") u.TAA = v != 0;
    else if (n == "FastGather") u.FastGather = v != 0;
    else if (n == "DisableSpatialReuse") u.DisableSpatialReuse = v != 0;
    else if (n == "DebugDraw_Lighting") u.DebugDraw_Lighting = v != 0;
    else if (n == "TemporalFrames") u.TemporalFrames = (int)v;
    else if (n == "Denoiser") u.Denoiser = v != 0;
    else if (n == "DenoiseSigma") u.DenoiseSigma = (float)v;
    else if (n == "DenoiseSigmaLum") u.DenoiseSigmaLum = (float)v;
    else if (n == "DenoiseSize") u.DenoiseSize = (int)v;
    else if (n == "PaperWhiteNit") u.PaperWhiteNit = (float)v;
    else if (n == "ProgressiveRender") e->progressiveRendering = v != 0;
    else if (n == "HDR") e->hdr = v != 0;
    else if (n == "TotalFrames") e->totalFrames = (uint32_t)v;
    else if (n == "HasSky") env.HasSky = v != 0;
    else if (n == "HasSun") env.HasSun = v != 0;
    else if (n == "SkyIntensity") env.SkyIntensity = (float)v;
    else if (n == "SunIntensity") env.SunIntensity = (float)v;
    else if (n == "SunRotation") env.SunRotation = (float)v;
    else if (n == "Aperture") env.cameras.at(0).Aperture = (float)v;
    else if (n == "FocalDistance") env.cameras.at(0).FocalDistance = (float)v;
    else if (n == "SelectedId") e->scene.SetSelectedId((uint32_t)v);
    else {
        g_err = "unknown setting '" + n + "'";
        return -1;
    }
    return 0;
}
int gkh_set_camera_lookat(void* h, const float* eye, const float* center, const float* up, float fov)
{
    auto& cam = ((EngineMirror*)h)->scene.GetEnvSettings().cameras.at(0);
    cam.ModelView = lookAt(vec3(eye[0], eye[1], eye[2]), vec3(center[0], center[1], center[2]), vec3(up[0], up[1], up[2]));
    cam.FieldOfView = fov;
    return 0;
}
void gkh_get_ubo(void* h, uint32_t w, uint32_t hgt, GkUniformBufferObject* out) { *out = ((EngineMirror*)h)->GetUniformBufferObject({0, 0}, {w, hgt}); }
void gkh_advance_frame(void* h) { ((EngineMirror*)h)->totalFrames += 1; }
void gkh_screen_ray(void* h, float x, float y, uint32_t w, uint32_t hgt, float* org, float* dir)
{
    vec3 o, d;
    ((EngineMirror*)h)->GetScreenToWorldRay(vec2(x, y), {w, hgt}, o, d);
    org[0] = o.x, org[1] = o.y, org[2] = o.z, dir[0] = d.x, dir[1] = d.y, dir[2] = d.z;
}

// ---- LogicRendererBase-shaped driver over the CUDA backend ----
void* gkh_renderer_create(void* engine, int device) { return new HostRenderer(*(EngineMirror*)engine, device); }
void gkh_renderer_destroy(void* r) { delete (HostRenderer*)r; }
int gkh_renderer_set_trace_all_rows(void* r, int on) { GKH_TRY(((HostRenderer*)r)->r.SetTraceAllRows(on != 0)) }
int gkh_renderer_set_tile(void* r, uint32_t index, uint32_t count, uint32_t rows) { GKH_TRY(((HostRenderer*)r)->r.SetTile(index, count, rows)) }
int gkh_renderer_create_swapchain(void* r, uint32_t w, uint32_t h) { GKH_TRY(((HostRenderer*)r)->r.CreateSwapChain({w, h})) }
int gkh_renderer_delete_swapchain(void* r) { GKH_TRY(((HostRenderer*)r)->r.DeleteSwapChain()) }
int gkh_renderer_post_load_scene(void* r) { GKH_TRY(((HostRenderer*)r)->r.OnPostLoadScene()) }
int gkh_renderer_before_next_frame(void* r) { GKH_TRY(((HostRenderer*)r)->r.BeforeNextFrame()) }
int gkh_renderer_render(void* r) { GKH_TRY(((HostRenderer*)r)->r.Render(nullptr, 0)) }
uint64_t gkh_renderer_instance_bytes_uploaded(void* r) { return ((HostRenderer*)r)->r.InstanceBytesUploaded(); }
void* gkh_renderer_context(void* r) { return ((HostRenderer*)r)->r.Context(); }

} // extern "C"
