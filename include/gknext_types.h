/*
 * gknext_types.h — plain-old-data wire formats shared by the reference's host
 * code and the CUDA backend.  Every struct here is a byte-for-byte mirror of a
 * struct the reference already shares between C++ and its shaders; sizes and
 * offsets are pinned by static asserts at the bottom.
 *
 * Reference layouts (paths relative to the gkNextRenderer checkout):
 *   UniformBufferObject  assets/shaders/common/BasicTypes.slang:21-83   (784 B)
 *   NodeProxy            assets/shaders/common/BasicTypes.slang:85-94   (208 B)
 *   ModelData            assets/shaders/common/BasicTypes.slang:96-110  ( 64 B)
 *   AmbientCube          assets/shaders/common/BasicTypes.slang:115-137 ( 56 B)
 *   VoxelData            assets/shaders/common/BasicTypes.slang:144-150 ( 16 B)
 *   LightObject          assets/shaders/common/BasicTypes.slang:170-181 ( 80 B)
 *   Material             src/Assets/Material.hpp:8-74                   ( 64 B)
 *   Vertex (CPU)         src/Assets/Vertex.hpp:9-26                     ( 52 B)
 *   GPUVertex            src/Assets/Vertex.hpp:32-48                    ( 24 B)
 *   RayCastResult        src/Assets/UniformBuffer.hpp:71-79             ( 48 B)
 *
 * Matrices are column-major float[16] exactly as glm::mat4 lays them out
 * (src/Utilities/Glm.hpp:3-5): element (row r, column c) is m[c*4 + r].
 * `bool` members of the shader structs are 32-bit (BasicTypes.slang:3).
 */
#ifndef GKNEXT_TYPES_H_
#define GKNEXT_TYPES_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__cplusplus)
#define GK_ALIGN(n) alignas(n)
#else
#define GK_ALIGN(n) _Alignas(n)
#endif

typedef struct GK_ALIGN(16) GkUniformBufferObject {
    float ModelView[16];
    float Projection[16];
    float ModelViewInverse[16];
    float ProjectionInverse[16];
    float ViewProjection[16];
    float PrevViewProjection[16];
    float ViewProjectionUnJit[16];
    float PrevViewProjectionUnJit[16];

    float ViewportRect[4];
    float SunDirection[4];
    float SunColor[4];
    float BackGroundColor[4]; /* unused by the reference; this backend reads it as the
                                 constant sky radiance when HasSky is set and no sky
                                 texture has been uploaded ("sky") */

    float SunViewProjection[16];

    float Aperture;
    float FocusDistance;
    float SkyRotation;
    float HeatmapScale;

    float PaperWhiteNit;
    float SkyIntensity;
    uint32_t SkyIdx;
    uint32_t TotalFrames;

    uint32_t MaxNumberOfBounces;
    uint32_t NumberOfSamples;
    uint32_t NumberOfBounces;
    uint32_t RandomSeed;

    uint32_t LightCount;
    uint32_t HasSky;
    uint32_t ShowHeatmap;
    uint32_t UseCheckerBoard;

    uint32_t TemporalFrames;
    uint32_t HasSun;
    uint32_t HDR;
    uint32_t AdaptiveSample;

    float AdaptiveVariance;
    uint32_t AdaptiveSteps;
    uint32_t TAA;
    uint32_t SelectedId;

    uint32_t ShowEdge;
    uint32_t ProgressiveRender;
    float BFSigma;
    float BFSigmaLum;

    float BFSigmaNormal;
    uint32_t BFSize;

    uint32_t FastGather;

    uint32_t FastInterpole;
    uint32_t DebugDraw_Lighting;
    uint32_t DisableSpatialReuse;
    uint32_t SuperResolution;
} GkUniformBufferObject;

typedef struct GK_ALIGN(16) GkNodeProxy {
    uint32_t instanceId;
    uint32_t modelId; /* model * 10 + section (src/Assets/Scene.cpp:494) */
    uint32_t visible;
    uint32_t nort;
    float worldTS[16];
    float combinedPrevTS[16];
    uint32_t matId[16];
} GkNodeProxy;

typedef struct GK_ALIGN(16) GkModelData {
    uint32_t indexOffset;
    uint32_t indexCount;
    uint32_t vertexOffset;
    uint32_t vertexCount;
    float localAabbMin[4];
    float localAabbMax[4];
    uint32_t modelType;
    uint32_t voxelDataIdx;
    uint32_t reorderOffset;
    uint32_t reserved2;
} GkModelData;

typedef struct GK_ALIGN(8) GkAmbientCube {
    uint32_t PosZ, NegZ, PosY, NegY, PosX, NegX;
    uint32_t PosZ_D, NegZ_D, PosY_D, NegY_D, PosX_D, NegX_D;
    uint32_t skyVisibility_pznzpyny;
    uint32_t skyVisibility_pxnxs0s1;
} GkAmbientCube;

typedef struct GK_ALIGN(16) GkVoxelData {
    uint32_t matId;
    uint32_t age;
    uint32_t distanceToSolid_gg_z01;
    uint32_t distanceToSolid_x01_y01;
} GkVoxelData;

typedef struct GK_ALIGN(16) GkLightObject {
    float p0[4];
    float p1[4];
    float p3[4];
    float normal_area[4];
    uint32_t lightMatIdx;
    uint32_t reserved1, reserved2, reserved3;
} GkLightObject;

enum GkMaterialModel {
    GK_MAT_LAMBERTIAN = 0,
    GK_MAT_METALLIC = 1,
    GK_MAT_DIELECTRIC = 2,
    GK_MAT_ISOTROPIC = 3,
    GK_MAT_DIFFUSE_LIGHT = 4,
    GK_MAT_MIXTURE = 5
};

typedef struct GK_ALIGN(16) GkMaterial {
    float Diffuse[4];
    int32_t DiffuseTextureId;
    int32_t MRATextureId;
    int32_t NormalTextureId;
    float Fuzziness;
    float RefractionIndex;
    uint32_t MaterialModel;
    float Metalness;
    float RefractionIndex2;
    float NormalTextureScale;
    float Reserverd2;
} GkMaterial;

/* CPU vertex, as held by Assets::Model before upload. */
typedef struct GkVertex {
    float Position[3];
    float Normal[3];
    float Tangent[4];
    float TexCoord[2];
    uint32_t MaterialIndex;
} GkVertex;

/* Shading vertex after Assets::MakeVertex: IEEE binary16 bit patterns. */
typedef struct GkGPUVertex {
    uint16_t posx, posy, posz, texcoordx;
    uint16_t normalx, normaly, normalz, texcoordy;
    uint16_t tangentx, tangenty, tangentz, tangentw; /* tangentw = (w>0 ? 2 : 0) << 8 | matIdx */
} GkGPUVertex;

typedef struct GkRayCastResult {
    float HitPoint[4];
    float Normal[4];
    float T;
    uint32_t InstanceId;
    uint32_t MaterialId;
    uint32_t Hitted;
} GkRayCastResult;

/* Input record of the batched ray cast (src/Assets/UniformBuffer.hpp:61-69). */
typedef struct GkRayCastIn {
    float Origin[4];
    float Direction[4];
    float TMin;
    float TMax;
    float Reversed0;
    float Reversed1;
} GkRayCastIn;

/* In-place record of the GPU ray-cast task: RayCastContext in assets/shaders/Task.RayCast.comp.slang, RayCastIO on the host
 * (src/Assets/UniformBuffer.hpp:81-85). */
typedef struct GkRayCastIO {
    GkRayCastIn Context;
    GkRayCastResult Result;
} GkRayCastIO;

/* One Assets::Model as it exists between Scene::Reload and Model::FreeMemory
 * (src/Assets/Scene.cpp:101-196): the backend copies during upload. */
typedef struct GkModelDesc {
    const GkVertex* vertices;
    const uint32_t* indices;
    uint32_t vertexCount;
    uint32_t indexCount; /* 3 per triangle */
} GkModelDesc;

typedef struct GkSceneDesc {
    const GkModelDesc* models;
    const GkMaterial* materials;
    const GkLightObject* lights;
    uint32_t modelCount;
    uint32_t materialCount;
    uint32_t lightCount;
    uint32_t reserved;
} GkSceneDesc;

/* Probe-grid constants (src/Assets/UniformBuffer.hpp:19-26; AmbientCube.slang:12-15). */
#define GK_CUBE_SIZE_XY 192
#define GK_CUBE_SIZE_Z 48
#define GK_CUBE_UNIT 0.25f

#ifdef __cplusplus
}
static_assert(sizeof(GkUniformBufferObject) == 784, "UBO must stay 784 bytes");
static_assert(sizeof(GkNodeProxy) == 208, "NodeProxy");
static_assert(sizeof(GkModelData) == 64, "ModelData");
static_assert(sizeof(GkAmbientCube) == 56, "AmbientCube");
static_assert(sizeof(GkVoxelData) == 16, "VoxelData");
static_assert(sizeof(GkLightObject) == 80, "LightObject");
static_assert(sizeof(GkMaterial) == 64, "Material");
static_assert(sizeof(GkVertex) == 52, "Vertex");
static_assert(sizeof(GkGPUVertex) == 24, "GPUVertex");
static_assert(sizeof(GkRayCastResult) == 48, "RayCastResult");
static_assert(sizeof(GkRayCastIn) == 48, "RayCastIn");
static_assert(sizeof(GkRayCastIO) == 96, "RayCastIO");
#endif

#endif /* GKNEXT_TYPES_H_ */
