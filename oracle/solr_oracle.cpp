// oracle/solr_oracle.cpp — TEST INFRASTRUCTURE.  CPU restatement of the reference's ray-propagation path.
//
// This file is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.  It restates, in plain scalar C++ (IEEE float, no
// FMA contraction: built with -ffp-contract=off), what the reference's CUDA engine computes per pixel,
// function by function, citing the reference lines it follows (paths relative to
// /root/reference/solr/engines/cuda/).  It is pinned against the reference itself: the reference's own
// device code compiled for the host (oracle/ref_build, libsolr_ref_cpu.so) must give bit-identical
// float buffers on the golden scenes (tests/test_oracle_vs_reference.py, tests/golden/).
//
// It also counts work in reference traversal order (box tests, primitive tests per type, rays, shade
// calls) — the algorithmic-flop numerator of the roofline.
//
// Deliberate, documented deviations from the reference's undefined behaviour:
//   * pixels are computed once each (the reference's 12x12 launch lets threads with x >= width alias the
//     next row, CudaRayTracer.cu:445-458 — a race, not a semantic);
//   * reads one or two floats past the end of the random table (CudaRayTracer.cu:475-477 at the last
//     pixel of a full-size frame) read 0;
//   * locals the reference leaves uninitialised (GI ray when no first hit, :115,:365) are zero.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/solr_b200_types.h"
#include "solr_oracle.h"

namespace
{
// ---- float3/float4 arithmetic as helper_math.h spells it (component-wise, left to right) ----
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 v3(const b200_float3& a) { return V3{a.x, a.y, a.z}; }
inline V4 v4(const b200_float4& a) { return V4{a.x, a.y, a.z, a.w}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float b) { return V3{a.x * b, a.y * b, a.z * b}; }
inline V3 operator*(float b, V3 a) { return V3{b * a.x, b * a.y, b * a.z}; }
inline V3 operator/(V3 a, float b) { return V3{a.x / b, a.y / b, a.z / b}; }
inline void operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline void operator*=(V3& a, float b) { a.x *= b; a.y *= b; a.z *= b; }
inline V4 operator+(V4 a, V4 b) { return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator-(V4 a, V4 b) { return V4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline V4 operator*(V4 a, V4 b) { return V4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline V4 operator*(V4 a, float b) { return V4{a.x * b, a.y * b, a.z * b, a.w * b}; }
inline V4 operator*(float b, V4 a) { return V4{b * a.x, b * a.y, b * a.z, b * a.w}; }
inline V4 operator/(V4 a, float b) { return V4{a.x / b, a.y / b, a.z / b, a.w / b}; }
inline void operator+=(V4& a, V4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
inline void operator-=(V4& a, V4 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w; }
inline void operator-=(V4& a, float b) { a.x -= b; a.y -= b; a.z -= b; a.w -= b; }
inline void operator*=(V4& a, float b) { a.x *= b; a.y *= b; a.z *= b; a.w *= b; }
inline void operator/=(V4& a, float b) { a.x /= b; a.y /= b; a.z /= b; a.w /= b; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }               // helper_math.h:1248
inline float length(V3 v) { return sqrtf(dot(v, v)); }                                    // :1291
inline V3 normalize(V3 v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }  // :62,:1309
inline float fmaxf_(float a, float b) { return a > b ? a : b; }
inline float fminf_(float a, float b) { return a < b ? a : b; }

struct RayOT { V3 origin; V3 direction; }; // reference "Ray": origin + look-at target
struct RayDir { V3 origin, direction, inv; int sx, sy, sz; };

// Optional log of the closest-hit rays (analysis only, tests/analysis_ray_order.py): 9 floats per ray —
// pixel, iteration, origin, look-at target, hit distance (or -1) — appended under an atomic counter.
static float* g_rayLog = nullptr;
static unsigned long long g_rayLogCapacity = 0;
static unsigned long long g_rayLogCount = 0;
static thread_local int g_currentPixel = 0;

struct Ctx
{
    const b200_SceneInfo& si;
    const b200_PostProcessingInfo& pp;
    const b200_BoundingBox* boxes; int nbBoxes;
    const b200_Primitive* prims; int nbPrims;
    const b200_Material* mats; int nbMats;
    const b200_LightInformation* lights; int lightInfoSize; int nbLamps;
    const unsigned char* tex;
    const float* randoms; int randomTableSize; // "MAX_BITMAP_SIZE" of the build being restated
    oracle_Counters* cnt;
    float rnd(long i) const { return (i >= 0 && i < randomTableSize) ? randoms[i] : 0.f; }
};

// VectorUtils.cuh:31-42
inline void saturateVector(V4& v)
{
    v.x = (v.x < 0.f) ? 0.f : v.x; v.y = (v.y < 0.f) ? 0.f : v.y; v.z = (v.z < 0.f) ? 0.f : v.z; v.w = (v.w < 0.f) ? 0.f : v.w;
    v.x = (v.x > 1.f) ? 1.f : v.x; v.y = (v.y > 1.f) ? 1.f : v.y; v.z = (v.z > 1.f) ? 1.f : v.z; v.w = (v.w > 1.f) ? 1.f : v.w;
}
// VectorUtils.cuh:45-52
inline V3 crossProduct(V3 b, V3 c) { return V3{b.y * c.z - b.z * c.y, b.z * c.x - b.x * c.z, b.x * c.y - b.y * c.x}; }
// VectorUtils.cuh:61-64
inline void vectorReflection(V3& r, V3 i, V3 n) { r = i - 2.f * dot(i, n) * n; }
// VectorUtils.cuh:73-87
inline void vectorRefraction(V3& refracted, V3 incident, float n1, V3 normal, float n2)
{
    refracted = incident;
    if (n2 != 0.f)
    {
        float eta = n1 / n2;
        float c1 = -dot(incident, normal);
        float cs2 = 1.f - eta * eta * (1.f - c1 * c1);
        if (cs2 >= 0.f)
            refracted = eta * incident + (eta * c1 - sqrtf(cs2)) * normal;
    }
}
// VectorUtils.cuh:92-95
inline V3 project(V3 A, V3 B) { return B * (dot(A, B) / dot(B, B)); }
// VectorUtils.cuh:104-142 (X, then Y, then Z Euler rotation about rotationCenter)
inline void vectorRotation(V3& v, V3 c, const float* angles)
{
    float cx = cosf(angles[0]), cy = cosf(angles[1]), cz = cosf(angles[2]);
    float sx = sinf(angles[0]), sy = sinf(angles[1]), sz = sinf(angles[2]);
    V3 vec = v3(v.x - c.x, v.y - c.y, v.z - c.z);
    V3 res = vec;
    res.y = vec.y * cx - vec.z * sx;
    res.z = vec.y * sx + vec.z * cx;
    vec = res;
    res.z = vec.z * cy - vec.x * sy;
    res.x = vec.z * sy + vec.x * cy;
    vec = res;
    res.x = vec.x * cz - vec.y * sz;
    res.y = vec.x * sz + vec.y * cz;
    v.x = res.x + c.x; v.y = res.y + c.y; v.z = res.z + c.z;
}

// GeometryIntersections.cuh:36-44
inline RayDir makeRayDir(V3 origin, V3 direction)
{
    RayDir r;
    r.origin = origin; r.direction = direction;
    r.inv.x = direction.x != 0.f ? 1.f / direction.x : 1.f;
    r.inv.y = direction.y != 0.f ? 1.f / direction.y : 1.f;
    r.inv.z = direction.z != 0.f ? 1.f / direction.z : 1.f;
    r.sx = r.inv.x < 0; r.sy = r.inv.y < 0; r.sz = r.inv.z < 0;
    return r;
}

// GeometryIntersections.cuh:52-79
inline bool boxIntersection(const b200_BoundingBox& box, const RayDir& ray, float t0, float t1)
{
    float tmin = (box.parameters[ray.sx].x - ray.origin.x) * ray.inv.x;
    float tmax = (box.parameters[1 - ray.sx].x - ray.origin.x) * ray.inv.x;
    float tymin = (box.parameters[ray.sy].y - ray.origin.y) * ray.inv.y;
    float tymax = (box.parameters[1 - ray.sy].y - ray.origin.y) * ray.inv.y;
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (box.parameters[ray.sz].z - ray.origin.z) * ray.inv.z;
    float tzmax = (box.parameters[1 - ray.sz].z - ray.origin.z) * ray.inv.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return ((tmin < t1) && (tmax > t0));
}

// ---------------------------------------------------------------- texture mapping (TextureMapping.cuh)
inline void normalMap(int index, const b200_Material& m, const unsigned char* t, V3& normal, float strength) // :30-40
{
    int i = m.textureOffset.y + index;
    unsigned char r = t[i], g = t[i + 1];
    normal.x -= strength * (r / 256.f - 0.5f);
    normal.y -= strength * (g / 256.f - 0.5f);
    normal.z = 0.f;
}
inline void bumpMap(int index, const b200_Material& m, const unsigned char* t, float& value) // :45-57
{
    int i = m.textureOffset.z + index;
    value = 10.f * (t[i] + t[i + 1] + t[i + 2]) / 768.f;
}
inline void specularMap(int index, const b200_Material& m, const unsigned char* t, V4& specular) // :62-73
{
    int i = m.textureOffset.w + index;
    specular.x = t[i] / 256.f;
    specular.y = 1000.f * t[i + 1] / 256.f;
    specular.z = t[i + 2] / 256.f;
}
inline void reflectionMap(int index, const b200_Material& m, const unsigned char* t, V4& attributes) // :78-87
{
    int i = m.advancedTextureOffset.x + index;
    attributes.x *= (t[i] + t[i + 1] + t[i + 2]) / 768.f;
}
inline void transparencyMap(int index, const b200_Material& m, const unsigned char* t, V4& attributes) // :92-102
{
    int i = m.advancedTextureOffset.y + index;
    attributes.y *= (t[i] + t[i + 1] + t[i + 2]) / 768.f;
}
inline void ambientOcclusionMap(int index, const b200_Material& m, const unsigned char* t, V4& adv) // :107-116
{
    int i = m.advancedTextureOffset.z + index;
    adv.x = (t[i] + t[i + 1] + t[i + 2]) / 768.f;
}
// :118-160
inline void juliaSet(const b200_Material& material, const b200_SceneInfo& si, float x, float y, V4& color)
{
    float W = (float)material.textureMapping.x, H = (float)material.textureMapping.y;
    float cRe = -0.7f + 0.4f * sinf(si.timestamp / 1500.f);
    float cIm = 0.27015f + 0.4f * cosf(si.timestamp / 2000.f);
    float newRe = 1.5f * (x - W / 2.f) / (0.5f * W);
    float newIm = (y - H / 2.f) / (0.5f * H);
    int n;
    float maxIterations = 40.f + si.pathTracingIteration;
    for (n = 0; n < maxIterations; n++)
    {
        float oldRe = newRe, oldIm = newIm;
        newRe = oldRe * oldRe - oldIm * oldIm + cRe;
        newIm = 2.f * oldRe * oldIm + cIm;
        if ((newRe * newRe + newIm * newIm) > 4.f) break;
    }
    color.x = 1.f - color.x * (n / maxIterations);
    color.y = 1.f - color.y * (n / maxIterations);
    color.z = 1.f - color.z * (n / maxIterations);
    color.w = 1.f - (n / maxIterations);
}
// :162-198 (Im_factor is a double in the reference)
inline void mandelbrotSet(const b200_Material& material, const b200_SceneInfo& si, float x, float y, V4& color)
{
    float W = (float)material.textureMapping.x, H = (float)material.textureMapping.y;
    float MinRe = -2.f, MaxRe = 1.f, MinIm = -1.2f;
    float MaxIm = MinIm + (MaxRe - MinRe) * H / W;
    float Re_factor = (MaxRe - MinRe) / (W - 1.f);
    double Im_factor = (MaxIm - MinIm) / (H - 1.f);
    float maxIterations = B200_NB_MAX_ITERATIONS + si.pathTracingIteration;
    float c_im = MaxIm - y * Im_factor;
    float c_re = MinRe + x * Re_factor;
    float Z_re = c_re, Z_im = c_im;
    bool isInside = true;
    unsigned n;
    for (n = 0; isInside && n < maxIterations; ++n)
    {
        float Z_re2 = Z_re * Z_re, Z_im2 = Z_im * Z_im;
        if (Z_re2 + Z_im2 > 4.f) isInside = false;
        Z_im = 2.f * Z_re * Z_im + c_im;
        Z_re = Z_re2 - Z_im2 + c_re;
    }
    color.x = 1.f - color.x * (n / maxIterations);
    color.y = 1.f - color.y * (n / maxIterations);
    color.z = 1.f - color.z * (n / maxIterations);
    color.w = 1.f - (n / maxIterations);
}
// shared tail of the three UV mappers (:246-282, :318-346, :412-445)
inline void fetchTexel(const Ctx& c, const b200_Material& material, int u, int v, V4& result, V3& normal, V4& specular,
                       V4& attributes, V4& adv)
{
    int A = (v * material.textureMapping.x + u) * material.textureMapping.w;
    int B = material.textureMapping.x * material.textureMapping.y * material.textureMapping.w;
    int index = A % B;
    int i = material.textureOffset.x + index;
    result.x = c.tex[i] / 256.f; result.y = c.tex[i + 1] / 256.f; result.z = c.tex[i + 2] / 256.f;
    float strength = 3.f;
    if (material.textureIds.z != B200_TEXTURE_NONE) bumpMap(index, material, c.tex, strength);
    if (material.textureIds.y != B200_TEXTURE_NONE) normalMap(index, material, c.tex, normal, strength);
    if (material.textureIds.w != B200_TEXTURE_NONE) specularMap(index, material, c.tex, specular);
    if (material.advancedTextureIds.x != B200_TEXTURE_NONE) reflectionMap(index, material, c.tex, attributes);
    if (material.advancedTextureIds.y != B200_TEXTURE_NONE) transparencyMap(index, material, c.tex, attributes);
    if (material.advancedTextureIds.z != B200_TEXTURE_NONE) ambientOcclusionMap(index, material, c.tex, adv);
}
// :205-286
inline V4 triangleUVMapping(const Ctx& c, const b200_Primitive& p, V3 areas, V3& normal, V4& specular, V4& attributes, V4& adv)
{
    const b200_Material& material = c.mats[p.materialId];
    V4 result = v4(material.color);
    float sum = areas.x + areas.y + areas.z;
    float Tx = (p.vt0.x * areas.x + p.vt1.x * areas.y + p.vt2.x * areas.z) / sum;
    float Ty = (p.vt0.y * areas.x + p.vt1.y * areas.y + p.vt2.y * areas.z) / sum;
    float mox = 0.f, moy = 0.f;
    if (material.attributes.y == 1)
    {
        mox = material.mappingOffset.x * c.si.timestamp;
        moy = material.mappingOffset.y * c.si.timestamp;
    }
    int u = Tx * material.textureMapping.x + mox;
    int v = Ty * material.textureMapping.y + moy;
    u = u % material.textureMapping.x;
    v = v % material.textureMapping.y;
    if (u >= 0 && u < material.textureMapping.x && v >= 0 && v < material.textureMapping.y)
    {
        switch (material.textureIds.x)
        {
        case B200_TEXTURE_MANDELBROT: mandelbrotSet(material, c.si, u, v, result); break;
        case B200_TEXTURE_JULIA: juliaSet(material, c.si, u, v, result); break;
        default: fetchTexel(c, material, u, v, result, normal, specular, attributes, adv);
        }
    }
    return result;
}
// :294-349
inline V4 sphereUVMapping(const Ctx& c, const b200_Primitive& p, V3 intersection, V3& normal, V4& specular, V4& attributes, V4& adv)
{
    const b200_Material& material = c.mats[p.materialId];
    V4 result = v4(material.color);
    V3 I = normalize(intersection - v3(p.p0));
    float U = ((atan2f(I.x, I.z) / 3.14159265358979323846f) + 1.f) * .5f;
    float V = (asinf(I.y) / 3.14159265358979323846f) + .5f;
    int u = material.textureMapping.x * (U * p.vt1.x);
    int v = material.textureMapping.y * (V * p.vt1.y);
    if (material.textureMapping.x != 0) u = u % material.textureMapping.x;
    if (material.textureMapping.y != 0) v = v % material.textureMapping.y;
    if (u >= 0 && u < material.textureMapping.x && v >= 0 && v < material.textureMapping.y)
        fetchTexel(c, material, u, v, result, normal, specular, attributes, adv);
    return result;
}
// :357-447 (non-Kinect branch)
inline V4 cubeMapping(const Ctx& c, const b200_Primitive& p, V3 intersection, V3& normal, V4& specular, V4& attributes, V4& adv)
{
    const b200_Material& material = c.mats[p.materialId];
    V4 result = v4(material.color);
    int u = ((p.type == B200_PT_CHECKBOARD) || (p.type == B200_PT_XZPLANE) || (p.type == B200_PT_XYPLANE))
                ? (intersection.x - p.p0.x + p.size.x) : (intersection.z - p.p0.z + p.size.z);
    int v = ((p.type == B200_PT_CHECKBOARD) || (p.type == B200_PT_XZPLANE))
                ? (intersection.z + p.p0.z + p.size.z) : (intersection.y - p.p0.y + p.size.y);
    if (material.textureMapping.x != 0) u = u % material.textureMapping.x;
    if (material.textureMapping.y != 0) v = v % material.textureMapping.y;
    if (u >= 0 && u < material.textureMapping.x && v >= 0 && v < material.textureMapping.x)
    {
        switch (material.textureIds.x)
        {
        case B200_TEXTURE_MANDELBROT: mandelbrotSet(material, c.si, u, v, result); break;
        case B200_TEXTURE_JULIA: juliaSet(material, c.si, u, v, result); break;
        default: fetchTexel(c, material, u, v, result, normal, specular, attributes, adv);
        }
    }
    return result;
}
// :449-456
inline bool wireFrameMapping(float x, float y, int width)
{
    int X = fabsf(x), Y = fabsf(y);
    return (X % 100 <= width) || (Y % 100 <= width);
}

// ---------------------------------------------------------------- primitive tests (GeometryIntersections.cuh)
// :87-151
inline V4 skyboxMapping(const Ctx& c, const RayOT& ray)
{
    const b200_Material& material = c.mats[c.si.skyboxMaterialId];
    V4 result = v4(material.color);
    V3 dir = normalize(ray.direction - ray.origin);
    float a = 2.f * dot(dir, dir);
    float b = 2.f * dot(ray.origin, dir);
    float cc = dot(ray.origin, ray.origin) - (c.si.skyboxRadius * c.si.skyboxRadius);
    float d = b * b - 2.f * a * cc;
    if (d <= 0.f || a == 0.f) return result;
    float r = sqrtf(d);
    float t1 = (-b - r) / a, t2 = (-b + r) / a;
    if (t1 <= c.si.geometryEpsilon && t2 <= c.si.geometryEpsilon) return result;
    float t = 0.f;
    if (t1 <= c.si.geometryEpsilon) t = t2;
    else if (t2 <= c.si.geometryEpsilon) t = t1;
    else t = (t1 < t2) ? t1 : t2;
    if (t < c.si.geometryEpsilon) return result;
    V3 intersection = normalize(ray.origin + t * dir);
    float U = ((atan2f(intersection.x, intersection.z) / 3.14159265358979323846f) + 1.f) * .5f;
    float V = (asinf(intersection.y) / 3.14159265358979323846f) + .5f;
    int u = int(material.textureMapping.x * U);
    int v = int(material.textureMapping.y * V);
    if (material.textureMapping.x != 0) u %= material.textureMapping.x;
    if (material.textureMapping.y != 0) v %= material.textureMapping.y;
    if (u >= 0 && u < material.textureMapping.x && v >= 0 && v < material.textureMapping.y)
    {
        int A = (v * material.textureMapping.x + u) * material.textureMapping.w;
        int B = material.textureMapping.x * material.textureMapping.y * material.textureMapping.w;
        int i = material.textureOffset.x + A % B;
        result.x = c.tex[i] / 256.f; result.y = c.tex[i + 1] / 256.f; result.z = c.tex[i + 2] / 256.f;
    }
    return result;
}
// :159-212
inline bool ellipsoidIntersection(const Ctx& c, const b200_Primitive& e, const RayDir& ray, V3& intersection, V3& normal, float& shadowIntensity)
{
    shadowIntensity = 1.f;
    V3 O_C = ray.origin - v3(e.p0);
    V3 dir = normalize(ray.direction);
    float a = ((dir.x * dir.x) / (e.size.x * e.size.x)) + ((dir.y * dir.y) / (e.size.y * e.size.y)) + ((dir.z * dir.z) / (e.size.z * e.size.z));
    float b = ((2.f * O_C.x * dir.x) / (e.size.x * e.size.x)) + ((2.f * O_C.y * dir.y) / (e.size.y * e.size.y)) + ((2.f * O_C.z * dir.z) / (e.size.z * e.size.z));
    float cc = ((O_C.x * O_C.x) / (e.size.x * e.size.x)) + ((O_C.y * O_C.y) / (e.size.y * e.size.y)) + ((O_C.z * O_C.z) / (e.size.z * e.size.z)) - 1.f;
    float d = ((b * b) - (4.f * a * cc));
    if (d < 0.f || a == 0.f || b == 0.f || cc == 0.f) return false;
    d = sqrtf(d);
    float t1 = (-b + d) / (2.f * a), t2 = (-b - d) / (2.f * a);
    if (t1 <= c.si.geometryEpsilon && t2 <= c.si.geometryEpsilon) return false;
    float t = 0.f;
    if (t1 <= c.si.geometryEpsilon) t = t2;
    else if (t2 <= c.si.geometryEpsilon) t = t1;
    else t = (t1 < t2) ? t1 : t2;
    if (t < c.si.geometryEpsilon) return false;
    intersection = ray.origin + t * dir;
    normal = intersection - v3(e.p0);
    normal.x = 2.f * normal.x / (e.size.x * e.size.x);
    normal.y = 2.f * normal.y / (e.size.y * e.size.y);
    normal.z = 2.f * normal.z / (e.size.z * e.size.z);
    normal = normalize(normal);
    return true;
}
// :220-284
inline bool sphereIntersection(const Ctx& c, const b200_Primitive& s, const RayDir& ray, V3& intersection, V3& normal, float& shadowIntensity)
{
    bool back = false;
    V3 O_C = ray.origin - v3(s.p0);
    V3 dir = normalize(ray.direction);
    float a = 2.f * dot(dir, dir);
    float b = 2.f * dot(O_C, dir);
    float cc = dot(O_C, O_C) - (s.size.x * s.size.x);
    float d = b * b - 2.f * a * cc;
    if (d <= 0.f || a == 0.f) return false;
    float r = sqrtf(d);
    float t1 = (-b - r) / a, t2 = (-b + r) / a;
    if (t1 <= c.si.geometryEpsilon && t2 <= c.si.geometryEpsilon) return false;
    float t = 0.f;
    if (t1 <= c.si.geometryEpsilon) { t = t2; back = true; }
    else if (t2 <= c.si.geometryEpsilon) t = t1;
    else t = (t1 < t2) ? t1 : t2;
    if (t < c.si.geometryEpsilon) return false;
    intersection = ray.origin + t * dir;
    if (c.mats[s.materialId].attributes.y == 0)
        normal = intersection - v3(s.p0);
    else
    {
        V3 nc;
        nc.x = s.p0.x + 0.008f * s.size.x * cosf(c.si.timestamp + intersection.x);
        nc.y = s.p0.y + 0.008f * s.size.y * sinf(c.si.timestamp + intersection.y);
        nc.z = s.p0.z + 0.008f * s.size.z * sinf(cosf(c.si.timestamp + intersection.z));
        normal = intersection - nc;
    }
    normal = normalize(normal);
    if (back) normal *= -1.f;
    r = dot(dir, normal);
    shadowIntensity = (c.mats[s.materialId].transparency != 0.f) ? (1.f - fabsf(r)) : 1.f;
    return true;
}
// :293-349 (cylinder) and :358-416 (cone: identical maths)
inline bool cylinderIntersection(const Ctx& c, const b200_Primitive& cy, const RayDir& ray, V3& intersection, V3& normal, float& shadowIntensity)
{
    V3 O_C = ray.origin - v3(cy.p0);
    V3 dir = ray.direction;
    V3 n1 = v3(cy.n1);
    V3 n = crossProduct(dir, n1);
    float ln = length(n);
    if ((ln < c.si.geometryEpsilon) && (ln > -c.si.geometryEpsilon)) return false;
    n = normalize(n);
    float d = fabsf(dot(O_C, n));
    if (d > cy.size.y) return false;
    V3 O = crossProduct(O_C, n1);
    float t = -dot(O, n) / ln;
    if (t < 0.f) return false;
    O = normalize(crossProduct(n, n1));
    float s = fabsf(sqrtf(cy.size.x * cy.size.x - d * d) / dot(dir, O));
    float t1 = t - s, t2 = t + s;
    intersection = ray.origin + t1 * dir;
    V3 HB1 = intersection - v3(cy.p0), HB2 = intersection - v3(cy.p1);
    float scale1 = dot(HB1, n1), scale2 = dot(HB2, n1);
    if (scale1 < c.si.geometryEpsilon || scale2 > c.si.geometryEpsilon)
    {
        intersection = ray.origin + t2 * dir;
        HB1 = intersection - v3(cy.p0); HB2 = intersection - v3(cy.p1);
        scale1 = dot(HB1, n1); scale2 = dot(HB2, n1);
        if (scale1 < c.si.geometryEpsilon || scale2 > c.si.geometryEpsilon) return false;
    }
    V3 V = intersection - v3(cy.p2);
    normal = V - project(V, n1);
    normal = normalize(normal);
    shadowIntensity = 1.f;
    return true;
}
// :424-567
inline bool planeIntersection(const Ctx& c, const b200_Primitive& p, const RayDir& ray, V3& intersection, V3& normal, float& shadowIntensity, bool reverse)
{
    bool collision = false;
    float reverted = reverse ? -1.f : 1.f;
    const b200_Material& mat = c.mats[p.materialId];
    normal = v3(p.n0);
    switch (p.type)
    {
    case B200_PT_MAGICCARPET:
    case B200_PT_CHECKBOARD:
    {
        intersection.y = p.p0.y;
        float y = ray.origin.y - p.p0.y;
        if (reverted * ray.direction.y < 0.f && reverted * ray.origin.y > reverted * p.p0.y)
        {
            intersection.x = ray.origin.x + y * ray.direction.x / -ray.direction.y;
            intersection.z = ray.origin.z + y * ray.direction.z / -ray.direction.y;
            collision = fabsf(intersection.x - p.p0.x) < p.size.x && fabsf(intersection.z - p.p0.z) < p.size.z;
        }
        break;
    }
    case B200_PT_XZPLANE:
    {
        float y = ray.origin.y - p.p0.y;
        if (reverted * ray.direction.y < 0.f && reverted * ray.origin.y > reverted * p.p0.y)
        {
            intersection.x = ray.origin.x + y * ray.direction.x / -ray.direction.y;
            intersection.y = p.p0.y;
            intersection.z = ray.origin.z + y * ray.direction.z / -ray.direction.y;
            collision = fabsf(intersection.x - p.p0.x) < p.size.x && fabsf(intersection.z - p.p0.z) < p.size.z;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(intersection.x, intersection.z, mat.attributes.w);
        }
        if (!collision && reverted * ray.direction.y > 0.f && reverted * ray.origin.y < reverted * p.p0.y)
        {
            normal = -normal;
            intersection.x = ray.origin.x + y * ray.direction.x / -ray.direction.y;
            intersection.y = p.p0.y;
            intersection.z = ray.origin.z + y * ray.direction.z / -ray.direction.y;
            collision = fabsf(intersection.x - p.p0.x) < p.size.x && fabsf(intersection.z - p.p0.z) < p.size.z;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(intersection.x, intersection.z, mat.attributes.w);
        }
        break;
    }
    case B200_PT_YZPLANE:
    {
        float x = ray.origin.x - p.p0.x;
        if (reverted * ray.direction.x < 0.f && reverted * ray.origin.x > reverted * p.p0.x)
        {
            intersection.x = p.p0.x;
            intersection.y = ray.origin.y + x * ray.direction.y / -ray.direction.x;
            intersection.z = ray.origin.z + x * ray.direction.z / -ray.direction.x;
            collision = fabsf(intersection.y - p.p0.y) < p.size.y && fabsf(intersection.z - p.p0.z) < p.size.z;
            if (mat.innerIllumination.x != 0.f)
                collision &= int(fabsf(intersection.z)) % 4000 < 2000 && int(fabsf(intersection.y)) % 4000 < 2000;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(intersection.y, intersection.z, mat.attributes.w);
        }
        if (!collision && reverted * ray.direction.x > 0.f && reverted * ray.origin.x < reverted * p.p0.x)
        {
            normal = -normal;
            intersection.x = p.p0.x;
            intersection.y = ray.origin.y + x * ray.direction.y / -ray.direction.x;
            intersection.z = ray.origin.z + x * ray.direction.z / -ray.direction.x;
            collision = fabsf(intersection.y - p.p0.y) < p.size.y && fabsf(intersection.z - p.p0.z) < p.size.z;
            if (mat.innerIllumination.x != 0.f)
                collision &= int(fabsf(intersection.z)) % 4000 < 2000 && int(fabsf(intersection.y)) % 4000 < 2000;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(intersection.y, intersection.z, mat.attributes.w);
        }
        break;
    }
    case B200_PT_XYPLANE:
    case B200_PT_CAMERA:
    {
        float z = ray.origin.z - p.p0.z;
        if (reverted * ray.direction.z < 0.f && reverted * ray.origin.z > reverted * p.p0.z)
        {
            intersection.z = p.p0.z;
            intersection.x = ray.origin.x + z * ray.direction.x / -ray.direction.z;
            intersection.y = ray.origin.y + z * ray.direction.y / -ray.direction.z;
            collision = fabsf(intersection.x - p.p0.x) < p.size.x && fabsf(intersection.y - p.p0.y) < p.size.y;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(intersection.x, intersection.y, mat.attributes.w);
        }
        if (!collision && reverted * ray.direction.z > 0.f && reverted * ray.origin.z < reverted * p.p0.z)
        {
            normal = -normal;
            intersection.z = p.p0.z;
            intersection.x = ray.origin.x + z * ray.direction.x / -ray.direction.z;
            intersection.y = ray.origin.y + z * ray.direction.y / -ray.direction.z;
            collision = fabsf(intersection.x - p.p0.x) < p.size.x && fabsf(intersection.y - p.p0.y) < p.size.y;
            if (mat.attributes.z == 2) collision &= wireFrameMapping(intersection.x, intersection.y, mat.attributes.w);
        }
        break;
    }
    }
    if (collision)
    {
        shadowIntensity = 1.f;
        V4 color = v4(mat.color);
        if (p.type == B200_PT_CAMERA || mat.textureIds.x != B200_TEXTURE_NONE)
        {
            V4 specular = {0.f, 0.f, 0.f, 0.f}, attributes = {0.f, 0.f, 0.f, 0.f}, adv = {0.f, 0.f, 0.f, 0.f};
            color = cubeMapping(c, p, intersection, normal, specular, attributes, adv);
            shadowIntensity = color.w;
        }
        if ((color.x + color.y + color.z) / 3.f >= c.si.transparentColor) collision = false;
    }
    return collision;
}
// :575-659
inline bool triangleIntersection(const Ctx& c, const b200_Primitive& tri, const RayDir& ray, V3& intersection, V3& normal, V3& areas, float& shadowIntensity, bool processingShadows)
{
    V3 p0 = v3(tri.p0), p1 = v3(tri.p1), p2 = v3(tri.p2);
    V3 E01 = p1 - p0, E03 = p2 - p0;
    V3 P = crossProduct(ray.direction, E03);
    float det = dot(E01, P);
    if (fabsf(det) < c.si.geometryEpsilon) return false;
    V3 T = ray.origin - p0;
    float a = dot(T, P) / det;
    if (a < 0.f || a > 1.f) return false;
    V3 Q = crossProduct(T, E01);
    float b = dot(ray.direction, Q) / det;
    if (b < 0.f || b > 1.f) return false;
    if ((a + b) > 1.f)
    {
        // E21 = p1 - p1 = 0 in the reference (:604) => det_ = 0 => always rejected (:607-608)
        V3 E23 = p0 - p1, E21 = p1 - p1;
        V3 P_ = crossProduct(ray.direction, E21);
        float det_ = dot(E23, P_);
        if (fabsf(det_) < c.si.geometryEpsilon) return false;
        V3 T_ = ray.origin - p2;
        float a_ = dot(T_, P_) / det_;
        if (a_ < 0.f) return false;
        V3 Q_ = crossProduct(T_, E23);
        float b_ = dot(ray.direction, Q_) / det_;
        if (b_ < 0.f) return false;
    }
    float t = dot(E03, Q) / det;
    if (t < 0) return false;
    intersection = ray.origin + t * ray.direction;
    V3 v0 = p0 - intersection, v1 = p1 - intersection, v2 = p2 - intersection;
    areas.x = 0.5f * length(crossProduct(v1, v2));
    areas.y = 0.5f * length(crossProduct(v0, v2));
    areas.z = 0.5f * length(crossProduct(v0, v1));
    normal = normalize((v3(tri.n0) * areas.x + v3(tri.n1) * areas.y + v3(tri.n2) * areas.z) / (areas.x + areas.y + areas.z));
    if (c.si.doubleSidedTriangles)
    {
        // dangling else of the reference (:643-647): the `else` binds to the inner `if`
        V3 N = normalize(ray.direction);
        if (processingShadows)
        {
            if (dot(N, normal) <= 0.f) return false;
            else if (dot(N, normal) >= 0.f) return false;
        }
    }
    V3 dir = normalize(ray.direction);
    float r = dot(dir, normal);
    if (r > 0.f) normal *= -1.f;
    shadowIntensity = 1.f;
    return true;
}

inline void countPrim(oracle_Counters* k, int type)
{
    switch (type)
    {
    case B200_PT_ENVIRONMENT: case B200_PT_SPHERE: k->sphere_tests++; break;
    case B200_PT_CYLINDER: k->cylinder_tests++; break;
    case B200_PT_CONE: k->cone_tests++; break;
    case B200_PT_ELLIPSOID: k->ellipsoid_tests++; break;
    case B200_PT_TRIANGLE: k->triangle_tests++; break;
    default: k->plane_tests++;
    }
}

// GeometryIntersections.cuh:667-772
bool intersectionWithPrimitives(const Ctx& c, const RayOT& ray, int iteration, int& closestPrimitive, V3& closestIntersection,
                                V3& closestNormal, V3& closestAreas, V4& colorBox, int currentMaterialId)
{
    c.cnt->rays++;
    bool intersections = false;
    float minDistance = (iteration < 2) ? c.si.viewDistance : c.si.viewDistance / (iteration + 1);
    RayDir r = makeRayDir(ray.origin, ray.direction - ray.origin);
    V3 intersection = {0.f, 0.f, 0.f}, normal = {0.f, 0.f, 0.f};
    bool i = false;
    float shadowIntensity = 0.f;
    int cptBoxes = 0;
    while (cptBoxes < c.nbBoxes)
    {
        const b200_BoundingBox& box = c.boxes[cptBoxes];
        c.cnt->box_tests++;
        if (boxIntersection(box, r, 0.f, minDistance))
        {
            if (c.si.renderBoxes != 0)
            {
                // the reference's device material array always has NB_MAX_MATERIALS slots; unset ones read as 0
                const int m = box.startIndex % B200_NB_MAX_MATERIALS;
                if (m < c.nbMats) colorBox += v4(c.mats[m].color) / 200.f;
                else colorBox += V4{0.f, 0.f, 0.f, 0.f} / 200.f;
            }
            else
            {
                for (int cpt = 0; cpt < box.nbPrimitives; ++cpt)
                {
                    const b200_Primitive& primitive = c.prims[box.startIndex + cpt];
                    const b200_Material& material = c.mats[primitive.materialId];
                    if (material.attributes.x == 0 || (material.attributes.x == 1 && currentMaterialId != primitive.materialId))
                    {
                        V3 areas = {0.f, 0.f, 0.f};
                        if (c.si.extendedGeometry)
                        {
                            i = false;
                            countPrim(c.cnt, primitive.type);
                            switch (primitive.type)
                            {
                            case B200_PT_ENVIRONMENT:
                            case B200_PT_SPHERE: i = sphereIntersection(c, primitive, r, intersection, normal, shadowIntensity); break;
                            case B200_PT_CYLINDER: i = cylinderIntersection(c, primitive, r, intersection, normal, shadowIntensity); break;
                            case B200_PT_CONE: i = cylinderIntersection(c, primitive, r, intersection, normal, shadowIntensity); break;
                            case B200_PT_ELLIPSOID: i = ellipsoidIntersection(c, primitive, r, intersection, normal, shadowIntensity); break;
                            case B200_PT_TRIANGLE: i = triangleIntersection(c, primitive, r, intersection, normal, areas, shadowIntensity, false); break;
                            default: i = planeIntersection(c, primitive, r, intersection, normal, shadowIntensity, false);
                            }
                        }
                        else
                        {
                            c.cnt->triangle_tests++;
                            i = triangleIntersection(c, primitive, r, intersection, normal, areas, shadowIntensity, false);
                        }
                        float distance = length(intersection - r.origin);
                        if (i && distance > c.si.geometryEpsilon && distance < minDistance)
                        {
                            minDistance = distance;
                            closestPrimitive = box.startIndex + cpt;
                            closestIntersection = intersection;
                            closestNormal = normal;
                            closestAreas = areas;
                            intersections = true;
                            c.cnt->accepted_hits++;
                        }
                    }
                }
            }
            ++cptBoxes;
        }
        else
            cptBoxes += box.indexForNextBox.x;
    }
    if (g_rayLog)
    {
        unsigned long long slot;
#pragma omp atomic capture
        slot = g_rayLogCount++;
        if (slot < g_rayLogCapacity)
        {
            float* w = g_rayLog + 9 * slot;
            w[0] = (float)g_currentPixel; w[1] = (float)iteration;
            w[2] = ray.origin.x; w[3] = ray.origin.y; w[4] = ray.origin.z;
            w[5] = ray.direction.x; w[6] = ray.direction.y; w[7] = ray.direction.z;
            w[8] = intersections ? minDistance : -1.f;
        }
    }
    return intersections;
}

// GeometryIntersections.cuh:798-908
float processShadows(const Ctx& c, V3 lampCenter, V3 origin, int lightId, int iteration, V4& color, int objectId)
{
    c.cnt->rays++; c.cnt->shadow_rays++;
    float result = 0.f;
    int cptBoxes = 0;
    color.x = 0.f; color.y = 0.f; color.z = 0.f;
    V3 dirv = lampCenter - origin;
    RayDir r = makeRayDir(origin + normalize(dirv) * c.si.rayEpsilon, dirv);
    const float minDistance = (iteration < 2) ? c.si.viewDistance : c.si.viewDistance / (iteration + 1);
    while (result < (c.si.shadowIntensity) && cptBoxes < c.nbBoxes)
    {
        const b200_BoundingBox& box = c.boxes[cptBoxes];
        c.cnt->box_tests++;
        if (boxIntersection(box, r, 0.f, minDistance))
        {
            int cpt = 0;
            while (result < c.si.shadowIntensity && cpt < box.nbPrimitives)
            {
                V3 intersection = {0.f, 0.f, 0.f}, normal = {0.f, 0.f, 0.f}, areas = {0.f, 0.f, 0.f};
                float shadowIntensity = 0.f;
                const b200_Primitive& primitive = c.prims[box.startIndex + cpt];
                if (primitive.index != lightId && primitive.index != objectId && c.mats[primitive.materialId].attributes.x == 0)
                {
                    bool hit = false;
                    if (c.si.extendedGeometry)
                    {
                        if (primitive.type != B200_PT_CAMERA) countPrim(c.cnt, primitive.type == B200_PT_ENVIRONMENT ? B200_PT_XYPLANE : primitive.type);
                        switch (primitive.type)
                        {
                        case B200_PT_SPHERE: hit = sphereIntersection(c, primitive, r, intersection, normal, shadowIntensity); break;
                        case B200_PT_ELLIPSOID: hit = ellipsoidIntersection(c, primitive, r, intersection, normal, shadowIntensity); break;
                        case B200_PT_CYLINDER: hit = cylinderIntersection(c, primitive, r, intersection, normal, shadowIntensity); break;
                        case B200_PT_CONE: hit = cylinderIntersection(c, primitive, r, intersection, normal, shadowIntensity); break;
                        case B200_PT_TRIANGLE: hit = triangleIntersection(c, primitive, r, intersection, normal, areas, shadowIntensity, true); break;
                        case B200_PT_CAMERA: hit = false; break;
                        default: hit = planeIntersection(c, primitive, r, intersection, normal, shadowIntensity, false); break;
                        }
                    }
                    else
                    {
                        c.cnt->triangle_tests++;
                        hit = triangleIntersection(c, primitive, r, intersection, normal, areas, shadowIntensity, true);
                    }
                    if (hit)
                    {
                        V3 O_I = intersection - r.origin;
                        V3 O_L = r.direction;
                        float l = length(O_I);
                        if (l > c.si.geometryEpsilon && l < length(O_L))
                        {
                            float ratio = shadowIntensity * c.si.shadowIntensity;
                            const b200_Material& m = c.mats[primitive.materialId];
                            if (m.transparency != 0.f)
                            {
                                O_L = normalize(O_L);
                                float a = fabsf(dot(O_L, normal));
                                float rr = (m.transparency == 0.f) ? 1.f : (1.f - m.transparency);
                                ratio *= rr * a;
                                color.x += ratio * (0.3f - 0.3f * m.color.x);
                                color.y += ratio * (0.3f - 0.3f * m.color.y);
                                color.z += ratio * (0.3f - 0.3f * m.color.z);
                            }
                            result += ratio;
                        }
                    }
                }
                ++cpt;
            }
            ++cptBoxes;
        }
        else
            cptBoxes += box.indexForNextBox.x;
    }
    result = fmaxf_(0.f, fminf_(result, c.si.shadowIntensity));
    return result;
}

// GeometryShaders.cuh:36-124
V4 intersectionShader(const Ctx& c, const b200_Primitive& p, V3& intersection, V3 areas, V3& normal, V4& specular, V4& attributes, V4& adv)
{
    const b200_Material& mat = c.mats[p.materialId];
    V4 col = v4(mat.color);
    col.w = 0.f;
    if (c.si.extendedGeometry)
    {
        switch (p.type)
        {
        case B200_PT_CONE: case B200_PT_CYLINDER: case B200_PT_ENVIRONMENT: case B200_PT_SPHERE: case B200_PT_ELLIPSOID:
            if (mat.textureIds.x != B200_TEXTURE_NONE) col = sphereUVMapping(c, p, intersection, normal, specular, attributes, adv);
            break;
        case B200_PT_CHECKBOARD:
            if (mat.textureIds.x != B200_TEXTURE_NONE)
                col = cubeMapping(c, p, intersection, normal, specular, attributes, adv);
            else
            {
                int x = c.si.viewDistance + ((intersection.x - p.p0.x) / p.size.x);
                int z = c.si.viewDistance + ((intersection.z - p.p0.z) / p.size.x);
                if (x % 2 == 0)
                {
                    if (z % 2 == 0) { col.x = 1.f - col.x; col.y = 1.f - col.y; col.z = 1.f - col.z; }
                }
                else
                {
                    if (z % 2 != 0) { col.x = 1.f - col.x; col.y = 1.f - col.y; col.z = 1.f - col.z; }
                }
            }
            break;
        case B200_PT_XYPLANE: case B200_PT_YZPLANE: case B200_PT_XZPLANE: case B200_PT_CAMERA:
            if (mat.textureIds.x != B200_TEXTURE_NONE) col = cubeMapping(c, p, intersection, normal, specular, attributes, adv);
            break;
        case B200_PT_TRIANGLE:
            if (mat.textureIds.x != B200_TEXTURE_NONE) col = triangleUVMapping(c, p, areas, normal, specular, attributes, adv);
            break;
        }
    }
    else if (mat.textureIds.x != B200_TEXTURE_NONE)
        col = triangleUVMapping(c, p, areas, normal, specular, attributes, adv);
    return col;
}

// GeometryIntersections.cuh:916-1080
V4 primitiveShader(const Ctx& c, int index, V3 origin, V3& normal, int objectId, V3& intersection, V3 areas, V4& closestColor,
                   int iteration, V4& refractionFromColor, float& shadowIntensity, V4& totalBlinn, V4& attributes)
{
    c.cnt->shade_calls++;
    const b200_Primitive& primitive = c.prims[objectId];
    const b200_Material& material = c.mats[primitive.materialId];
    V4 lampsColor = {0.f, 0.f, 0.f, 0.f};
    shadowIntensity = 0.f;
    V3 bumpNormal = {0.f, 0.f, 0.f};
    V4 adv = {0.f, 0.f, 0.f, 0.f};
    V4 specular = {material.specular.x, material.specular.y, material.specular.z, 0.f};
    V4 intersectionColor = intersectionShader(c, primitive, intersection, areas, bumpNormal, specular, attributes, adv);
    normal += bumpNormal;
    normal = normalize(normal);
    if (material.attributes.z == 1) return intersectionColor;
    if (c.si.graphicsLevel > B200_GL_NO_SHADING)
    {
        closestColor *= material.innerIllumination.x;
        for (int cpt = 0; cpt < c.lightInfoSize; ++cpt)
        {
            int cptLamp = (c.si.pathTracingIteration >= B200_NB_MAX_ITERATIONS) ? (c.si.pathTracingIteration % c.lightInfoSize) : 0;
            const b200_LightInformation& li = c.lights[cptLamp];
            if (li.primitiveId != primitive.index)
            {
                V3 center = v3(li.location);
                int t = (index + c.si.timestamp) % (c.randomTableSize - 3);
                const b200_Material& m = c.mats[li.materialId];
                if (c.si.pathTracingIteration >= B200_NB_MAX_ITERATIONS)
                {
                    float a = m.innerIllumination.y * 10.f * c.si.pathTracingIteration / c.si.maxPathTracingIterations;
                    center.x += c.rnd(t) * a; center.y += c.rnd(t + 1) * a; center.z += c.rnd(t + 2) * a;
                }
                V3 lightRay = center - intersection;
                float lightRayLength = length(lightRay);
                if (lightRayLength < m.innerIllumination.z)
                {
                    V4 shadowColor = {0.f, 0.f, 0.f, 0.f};
                    lightRay = normalize(lightRay);
                    float lambert = material.innerIllumination.x + dot(normal, lightRay);
                    if (lambert > 0.f && c.si.graphicsLevel > 3 && iteration < 4 && material.innerIllumination.x == 0.f)
                        shadowIntensity = processShadows(c, center, intersection, li.primitiveId, iteration, shadowColor, objectId);
                    if (c.si.graphicsLevel > B200_GL_NO_SHADING)
                    {
                        float photonEnergy = sqrtf(lightRayLength / m.innerIllumination.z);
                        photonEnergy = (photonEnergy > 1.f) ? 1.f : photonEnergy;
                        photonEnergy = (photonEnergy < 0.f) ? 0.f : photonEnergy;
                        lambert *= (lambert < 0.f) ? -c.mats[primitive.materialId].transparency : 1.f;
                        if (li.materialId != B200_MATERIAL_NONE)
                            lambert *= c.mats[li.materialId].innerIllumination.x;
                        else
                            lambert *= li.color.w;
                        if (material.innerIllumination.w != 0.f)
                            lambert *= (1.f + c.rnd(t) * material.innerIllumination.w * 100.f);
                        lambert *= (1.f - shadowIntensity);
                        lambert += c.si.backgroundColor.w;
                        lambert *= (1.f - photonEnergy);
                        lampsColor += lambert * v4(li.color) - shadowColor;
                        if (c.si.graphicsLevel > 1 && shadowIntensity < c.si.shadowIntensity)
                        {
                            V3 viewRay = normalize(intersection - origin);
                            V3 blinnDir = lightRay - viewRay;
                            float temp = sqrtf(dot(blinnDir, blinnDir));
                            if (temp != 0.f)
                            {
                                blinnDir = (1.f / temp) * blinnDir;
                                float blinnTerm = dot(blinnDir, normal);
                                blinnTerm = (blinnTerm < 0.f) ? 0.f : blinnTerm;
                                blinnTerm = specular.x * powf(blinnTerm, specular.y);
                                blinnTerm *= (1.f - photonEnergy);
                                totalBlinn += v4(li.color) * li.color.w * blinnTerm;
                                totalBlinn.w = specular.z;
                            }
                        }
                    }
                }
            }
            closestColor += intersectionColor * lampsColor;
            if (material.advancedTextureIds.z != B200_TEXTURE_NONE) closestColor *= adv.x;
            saturateVector(closestColor);
            refractionFromColor = intersectionColor;
            saturateVector(totalBlinn);
        }
    }
    else
        closestColor = intersectionColor;
    return closestColor;
}

// CudaRayTracer.cu:69-408
V4 launchRayTracing(const Ctx& c, int index, const RayOT& ray, float& depthOfField, b200_int4& id)
{
    const b200_SceneInfo& si = c.si;
    V4 intersectionColor = {0.f, 0.f, 0.f, 0.f};
    V3 closestIntersection = {0.f, 0.f, 0.f}, firstIntersection = {0.f, 0.f, 0.f}, normal = {0.f, 0.f, 0.f};
    int closestPrimitive = -1;
    bool carryon = true;
    RayOT rayOrigin = ray;
    float initialRefraction = 1.f;
    int iteration = 0;
    id.x = -1; id.z = 0; id.w = 0;
    int currentMaterialId = -2;
    float colorContributions[B200_NB_MAX_ITERATIONS + 1];
    V4 colors[B200_NB_MAX_ITERATIONS + 1];
    memset(colorContributions, 0, sizeof(colorContributions));
    memset(colors, 0, sizeof(colors));
    V4 recursiveBlinn = {0.f, 0.f, 0.f, 0.f};
    float shadowIntensity = 0.f;
    V4 refractionFromColor = {0.f, 0.f, 0.f, 0.f};
    V3 reflectedTarget = {0.f, 0.f, 0.f};
    V4 closestColor = {0.f, 0.f, 0.f, 0.f}, colorBox = {0.f, 0.f, 0.f, 0.f};
    V3 latestIntersection = ray.origin;
    float rayLength = 0.f;
    depthOfField = si.viewDistance;
    int reflectedRays = -1;
    RayOT reflectedRay = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    float reflectedRatio = 0.f;
    RayOT pathTracingRay = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    float pathTracingRatio = 0.f;
    V4 pathTracingColor = {0.f, 0.f, 0.f, 0.f};
    bool useGlobalIllumination = false;
    V4 rBlinn = {0.f, 0.f, 0.f, 0.f};
    int currentMaxIteration = (si.graphicsLevel < B200_GL_REFLECTIONS) ? 1 : si.nbRayIterations + si.pathTracingIteration;
    currentMaxIteration = (currentMaxIteration > B200_NB_MAX_ITERATIONS) ? B200_NB_MAX_ITERATIONS : currentMaxIteration;

    while (iteration < currentMaxIteration && rayLength < si.viewDistance && carryon)
    {
        V3 areas = {0.f, 0.f, 0.f};
        if (iteration == 0) c.cnt->primary_rays++;
        carryon = intersectionWithPrimitives(c, rayOrigin, iteration, closestPrimitive, closestIntersection, normal, areas, colorBox, currentMaterialId);
        if (carryon)
        {
            const b200_Primitive& prim = c.prims[closestPrimitive];
            const b200_Material& mat = c.mats[prim.materialId];
            currentMaterialId = prim.materialId;
            V4 attributes = {mat.reflection, mat.transparency, mat.refraction, mat.opacity};
            if (iteration == 0)
            {
                colors[iteration] = V4{0.f, 0.f, 0.f, 0.f};
                colorContributions[iteration] = 1.f;
                firstIntersection = closestIntersection;
                latestIntersection = closestIntersection;
                depthOfField = length(firstIntersection - ray.origin);
                if (mat.innerIllumination.x == 0.f && (si.advancedIllumination == B200_AI_BASIC || si.advancedIllumination == B200_AI_FULL))
                {
                    int t = (index + si.pathTracingIteration * 100 + si.timestamp) % (c.randomTableSize - 3);
                    pathTracingRay.origin = closestIntersection + normal * si.rayEpsilon;
                    pathTracingRay.direction.x = normal.x + 100.f * c.rnd(t);
                    pathTracingRay.direction.y = normal.y + 100.f * c.rnd(t + 1);
                    pathTracingRay.direction.z = normal.z + 100.f * c.rnd(t + 2);
                    float cos_theta = dot(normalize(pathTracingRay.direction), normal);
                    if (cos_theta < 0.f) pathTracingRay.direction = -pathTracingRay.direction;
                    pathTracingRay.direction += closestIntersection;
                    pathTracingRatio = (1.f - attributes.y) * fabsf(cos_theta);
                    useGlobalIllumination = true;
                }
                id.x = prim.index;
            }
            rBlinn.w = attributes.y;
            colors[iteration] = primitiveShader(c, index, rayOrigin.origin, normal, closestPrimitive, closestIntersection, areas, closestColor,
                                                iteration, refractionFromColor, shadowIntensity, rBlinn, attributes);
            id.z += mat.innerIllumination.x * 256;
            float segmentLength = length(closestIntersection - latestIntersection);
            latestIntersection = closestIntersection;
            float transparency = attributes.y;
            float a = 0.f;
            if (attributes.y != 0.f)
            {
                float refraction = attributes.z;
                if (initialRefraction == refraction)
                {
                    refraction = 1.f;
                    float len = segmentLength * (attributes.w * (1.f - transparency));
                    rayLength += len;
                    rayLength = (rayLength > si.viewDistance) ? si.viewDistance : rayLength;
                    a = (rayLength / si.viewDistance);
                    colors[iteration].x -= a; colors[iteration].y -= a; colors[iteration].z -= a;
                }
                V3 O_E = normalize(closestIntersection - rayOrigin.origin);
                vectorRefraction(reflectedTarget, O_E, refraction, normal, initialRefraction);
                colorContributions[iteration] = transparency - a;
                initialRefraction = refraction;
                if (reflectedRays == -1 && attributes.x != 0.f)
                {
                    vectorReflection(reflectedRay.direction, O_E, normal);
                    reflectedRay.origin = closestIntersection + reflectedRay.direction * si.rayEpsilon;
                    reflectedRay.direction = closestIntersection + reflectedRay.direction;
                    reflectedRatio = attributes.x;
                    reflectedRays = iteration;
                }
            }
            else if (attributes.x != 0.f)
            {
                V3 O_E = normalize(closestIntersection - rayOrigin.origin);
                vectorReflection(reflectedTarget, O_E, normal);
                colorContributions[iteration] = attributes.x;
            }
            else
            {
                carryon = false;
                colorContributions[iteration] = 1.f;
            }
            rBlinn /= (float)(iteration + 1);
            recursiveBlinn.x = (rBlinn.x > recursiveBlinn.x) ? rBlinn.x : recursiveBlinn.x;
            recursiveBlinn.y = (rBlinn.y > recursiveBlinn.y) ? rBlinn.y : recursiveBlinn.y;
            recursiveBlinn.z = (rBlinn.z > recursiveBlinn.z) ? rBlinn.z : recursiveBlinn.z;
            rayOrigin.origin = closestIntersection + reflectedTarget * si.rayEpsilon;
            rayOrigin.direction = closestIntersection + reflectedTarget;
            if (si.pathTracingIteration != 0 && mat.color.w != 0.f)
            {
                float ratio = mat.color.w;
                ratio *= (attributes.y == 0.f) ? 1000.f : 1.f;
                int rindex = (index + si.timestamp) % (c.randomTableSize - 3);
                rayOrigin.direction.x += c.rnd(rindex) * ratio;
                rayOrigin.direction.y += c.rnd(rindex + 1) * ratio;
                rayOrigin.direction.z += c.rnd(rindex + 2) * ratio;
            }
        }
        else
        {
            if (si.skyboxMaterialId != B200_MATERIAL_NONE)
            {
                colors[iteration] = skyboxMapping(c, rayOrigin);
                float rad = colors[iteration].x + colors[iteration].y + colors[iteration].z;
                id.z += (rad > 2.5f) ? rad * 256.f : 0.f;
            }
            else if (si.gradientBackground)
            {
                V3 up = {0.f, 1.f, 0.f};
                V3 dir = normalize(rayOrigin.direction - rayOrigin.origin);
                float angle = 0.5f - dot(up, dir);
                angle = (angle > 1.f) ? 1.f : angle;
                colors[iteration] = (1.f - angle) * v4(si.backgroundColor);
            }
            else
                colors[iteration] = v4(si.backgroundColor);
            colorContributions[iteration] = 1.f;
        }
        iteration++;
    }

    V3 areas = {0.f, 0.f, 0.f};
    if (si.graphicsLevel >= B200_GL_REFLECTIONS && reflectedRays != -1)
        if (intersectionWithPrimitives(c, reflectedRay, reflectedRays, closestPrimitive, closestIntersection, normal, areas, colorBox, currentMaterialId))
        {
            V4 attributes = {c.mats[c.prims[closestPrimitive].materialId].reflection, 0.f, 0.f, 0.f};
            V4 color = primitiveShader(c, index, reflectedRay.origin, normal, closestPrimitive, closestIntersection, areas, closestColor,
                                       reflectedRays, refractionFromColor, shadowIntensity, rBlinn, attributes);
            colors[reflectedRays] += color * reflectedRatio;
            id.w = shadowIntensity * 255;
        }

    bool test = true;
    if ((si.advancedIllumination == B200_AI_BASIC || si.advancedIllumination == B200_AI_FULL) && si.pathTracingIteration >= B200_NB_MAX_ITERATIONS)
    {
        if (useGlobalIllumination && si.advancedIllumination == B200_AI_FULL)
        {
            if (intersectionWithPrimitives(c, pathTracingRay, 30, closestPrimitive, closestIntersection, normal, areas, colorBox, B200_MATERIAL_NONE))
            {
                if (c.prims[closestPrimitive].materialId != B200_MATERIAL_NONE)
                {
                    const b200_Material& material = c.mats[c.prims[closestPrimitive].materialId];
                    if (material.innerIllumination.x == 0.f)
                    {
                        colors[0] = v4(material.color) * material.innerIllumination.x * pathTracingRatio;
                        test = false;
                    }
                    else
                        colors[0] = v4(material.color) * pathTracingRatio;
                }
                if (test)
                {
                    pathTracingRatio *= 0.1f; // STANDARD_LUNINANCE_STRENGTH, Consts.h:52
                    V4 attributes = {0.f, 0.f, 0.f, 0.f};
                    const b200_Material& material = c.mats[c.prims[closestPrimitive].materialId];
                    if (material.innerIllumination.x == 0.f)
                        colors[0] -= si.shadowIntensity;
                    else
                        pathTracingColor = primitiveShader(c, index, pathTracingRay.origin, normal, closestPrimitive, closestIntersection, areas,
                                                           closestColor, iteration, refractionFromColor, shadowIntensity, rBlinn, attributes);
                }
            }
            else if (si.skyboxMaterialId != B200_MATERIAL_NONE)
            {
                pathTracingColor = skyboxMapping(c, pathTracingRay);
                pathTracingRatio *= 0.2f; // SKYBOX_LUNINANCE_STRENGTH, Consts.h:53
            }
        }
        else if (si.skyboxMaterialId != B200_MATERIAL_NONE)
        {
            pathTracingColor = skyboxMapping(c, pathTracingRay);
            pathTracingRatio *= 0.2f;
        }
        if (test) colors[0] += pathTracingColor * pathTracingRatio;
    }

    if (test)
    {
        for (int i = iteration - 2; i >= 0; --i)
            colors[i] = colors[i] * (1.f - colorContributions[i]) + colors[i + 1] * colorContributions[i];
        intersectionColor = colors[0];
        intersectionColor += recursiveBlinn;
    }
    else
        intersectionColor = colors[0];

    float D1 = si.viewDistance * 0.95f;
    if (si.atmosphericEffect == B200_AE_FOG && depthOfField > D1)
    {
        float D2 = si.viewDistance * 0.05f;
        float a = depthOfField - D1;
        float b = 1.f - (a / D2);
        intersectionColor = intersectionColor * b + v4(si.backgroundColor) * (1.f - b);
    }
    id.y = iteration;
    intersectionColor -= colorBox;
    return intersectionColor;
}

// GeometryShaders.cuh:132-165
inline void makeColor(const b200_SceneInfo& si, V4 color, unsigned char* bitmap, int index)
{
    int mdc = index * B200_COLOR_DEPTH;
    color.x = (color.x > 1.f) ? 1.f : color.x; color.y = (color.y > 1.f) ? 1.f : color.y; color.z = (color.z > 1.f) ? 1.f : color.z;
    color.x = (color.x < 0.f) ? 0.f : color.x; color.y = (color.y < 0.f) ? 0.f : color.y; color.z = (color.z < 0.f) ? 0.f : color.z;
    if (si.frameBufferType == B200_FT_BGR)
    {
        int y = index / si.size.y, x = index % si.size.x;
        int i = ((y + 1) * si.size.y - x - 1) * B200_COLOR_DEPTH;
        bitmap[i] = (unsigned char)(color.z * 255.f);
        bitmap[i + 1] = (unsigned char)(color.y * 255.f);
        bitmap[i + 2] = (unsigned char)(color.x * 255.f);
    }
    else
    {
        bitmap[mdc] = (unsigned char)(color.x * 255.f);
        bitmap[mdc + 1] = (unsigned char)(color.y * 255.f);
        bitmap[mdc + 2] = (unsigned char)(color.z * 255.f);
    }
}

// CudaRayTracer.cu:437-563 (k_standardRenderer) and :840-926 (k_anaglyphRenderer), one pixel
void renderPixel(const Ctx& c, int x, int y, const float* eye, const float* target, const float* angles,
                 b200_PostProcessingBuffer* post, b200_int4* ids)
{
    const b200_SceneInfo& si = c.si;
    const int W = si.size.x, H = si.size.y;
    const int index = y * W + x;
    if (si.pathTracingIteration > ids[index].y && ids[index].w == 0 && si.pathTracingIteration > 0 &&
        si.pathTracingIteration <= B200_NB_MAX_ITERATIONS)
        return;
    c.cnt->pixels++;
    V3 origin = v3(eye[0], eye[1], eye[2]), direction = v3(target[0], target[1], target[2]);
    V3 rotationCenter = {0.f, 0.f, 0.f};
    if (si.cameraType == B200_CT_VR) rotationCenter = origin;
    float dof = 0.f;

    if (si.cameraType == B200_CT_ANAGLYPH)
    {
        float ratio = (float)W / (float)H;
        float stepx = ratio * angles[3] / (float)W, stepy = angles[3] / (float)H;
        V4 eyeColor[2];
        for (int e = 0; e < 2; ++e)
        {
            RayOT eyeRay;
            eyeRay.origin = v3(e == 0 ? origin.x - si.eyeSeparation : origin.x + si.eyeSeparation, origin.y, origin.z);
            eyeRay.direction.x = direction.x - stepx * (float)(x - (W / 2));
            eyeRay.direction.y = direction.y + stepy * (float)(y - (H / 2));
            eyeRay.direction.z = direction.z;
            vectorRotation(eyeRay.origin, rotationCenter, angles);
            vectorRotation(eyeRay.direction, rotationCenter, angles);
            eyeColor[e] = launchRayTracing(c, index, eyeRay, dof, ids[index]);
        }
        float r1 = eyeColor[0].x * 0.299f + eyeColor[0].y * 0.587f + eyeColor[0].z * 0.114f;
        float b1 = 0.f, g1 = 0.f, r2 = 0.f, g2 = eyeColor[1].y, b2 = eyeColor[1].z;
        if (si.pathTracingIteration == 0) post[index].colorInfo.w = dof;
        if (si.pathTracingIteration <= B200_NB_MAX_ITERATIONS)
        {
            post[index].colorInfo.x = r1 + r2; post[index].colorInfo.y = g1 + g2; post[index].colorInfo.z = b1 + b2;
        }
        else
        {
            post[index].colorInfo.x += r1 + r2; post[index].colorInfo.y += g1 + g2; post[index].colorInfo.z += b1 + b2;
        }
        return;
    }

    if (si.cameraType == B200_CT_VR)
    {
        // k_3DVisionRenderer, CudaRayTracer.cu:953-1043: side-by-side stereo, one ray tree per pixel.  The focus distance is
        // read from the accumulation buffer while the same launch may be writing it (iteration 0 only): a race in the
        // reference; here whatever the buffer holds when the pixel is reached.
        const float focus = fabsf(post[W / 2 * H / 2].colorInfo.w - origin.z);
        const float eyeSeparation = si.eyeSeparation * (direction.z / focus);
        dof = c.pp.param1;
        const int halfWidth = W / 2;
        const float ratio = (float)W / (float)H;
        const float stepx = ratio * angles[3] / (float)W, stepy = angles[3] / (float)H;
        RayOT eyeRay;
        if (x < halfWidth)
        {
            eyeRay.origin = v3(origin.x + eyeSeparation, origin.y, origin.z);
            eyeRay.direction.x = direction.x - stepx * (float)(x - (W / 2) + halfWidth / 2) + si.eyeSeparation;
        }
        else
        {
            eyeRay.origin = v3(origin.x - eyeSeparation, origin.y, origin.z);
            eyeRay.direction.x = direction.x - stepx * (float)(x - (W / 2) - halfWidth / 2) - si.eyeSeparation;
        }
        eyeRay.direction.y = direction.y + stepy * (float)(y - (H / 2));
        eyeRay.direction.z = direction.z;
        vectorRotation(eyeRay.origin, rotationCenter, angles);
        vectorRotation(eyeRay.direction, rotationCenter, angles);
        V4 color = launchRayTracing(c, index, eyeRay, dof, ids[index]);
        if (si.advancedIllumination == B200_AI_RANDOM)
        {
            int rindex = (index + si.timestamp) % c.randomTableSize;
            color += v4(si.backgroundColor) * c.rnd(rindex) * 5.f;
        }
        if (si.pathTracingIteration == 0) post[index].colorInfo.w = dof;
        if (si.pathTracingIteration <= B200_NB_MAX_ITERATIONS)
        {
            post[index].colorInfo.x = color.x; post[index].colorInfo.y = color.y; post[index].colorInfo.z = color.z;
        }
        else
        {
            post[index].colorInfo.x += color.x; post[index].colorInfo.y += color.y; post[index].colorInfo.z += color.z;
        }
        return;
    }

    if (si.cameraType == B200_CT_PANORAMIC)
    {
        // k_fishEyeRenderer, CudaRayTracer.cu:757-813: 360 degrees around the vertical axis across the image width
        RayOT ray = {origin, direction};
        if (si.pathTracingIteration >= B200_NB_MAX_ITERATIONS)
        {
            const int rindex = (index + si.timestamp) % (c.randomTableSize - 3);
            const float a = float(si.pathTracingIteration) / float(si.maxPathTracingIterations);
            ray.direction.x += c.rnd(rindex) * post[index].colorInfo.w * c.pp.param2 * a;
            ray.direction.y += c.rnd(rindex + 1) * post[index].colorInfo.w * c.pp.param2 * a;
            ray.direction.z += c.rnd(rindex + 2) * post[index].colorInfo.w * c.pp.param2 * a;
        }
        const float stepy = angles[3] / (float)H;
        ray.direction.y = ray.direction.y + stepy * (float)(y - (H / 2));
        const float stepx = 2.f * 3.14159265358979323846f / W;
        const float fishEyeAngles[4] = {0.f, angles[1] + stepx * (float)x, 0.f, 0.f};
        vectorRotation(ray.direction, ray.origin, fishEyeAngles);
        V4 color = {0.f, 0.f, 0.f, 0.f};
        color += launchRayTracing(c, index, ray, dof, ids[index]);
        if (si.pathTracingIteration == 0) post[index].colorInfo.w = dof;
        if (si.pathTracingIteration <= B200_NB_MAX_ITERATIONS)
        {
            post[index].colorInfo.x = color.x; post[index].colorInfo.y = color.y; post[index].colorInfo.z = color.z;
        }
        else
        {
            post[index].colorInfo.x += color.x; post[index].colorInfo.y += color.y; post[index].colorInfo.z += color.z;
        }
        return;
    }

    RayOT ray = {origin, direction};
    const float AA[4][2] = {{3.f, 5.f}, {5.f, -3.f}, {-3.f, -5.f}, {-5.f, 3.f}};
    bool antialiasingActivated = (si.cameraType == B200_CT_ANTIALIASED);
    // NATURAL_DEPTHOFFIELD is defined (Consts.h:54); precedence as written: index + (timestamp % (MAX-2))
    if (c.pp.type != B200_PPE_DEPTH_OF_FIELD && si.pathTracingIteration >= B200_NB_MAX_ITERATIONS)
    {
        float a = (c.pp.param1 / 20000.f);
        long rindex = (long)index + si.timestamp % (c.randomTableSize - 2);
        ray.origin.x += c.rnd(rindex) * post[index].colorInfo.w * a;
        ray.origin.y += c.rnd(rindex + 1) * post[index].colorInfo.w * a;
    }
    if (si.cameraType == B200_CT_ORTHOGRAPHIC)
    {
        ray.direction.x = ray.origin.z * 0.001f * (x - (W / 2));
        ray.direction.y = -ray.origin.z * 0.001f * (y - (H / 2));
        ray.origin.x = ray.direction.x;
        ray.origin.y = ray.direction.y;
    }
    else
    {
        float ratio = (float)W / (float)H;
        float stepx = ratio * angles[3] / (float)W, stepy = angles[3] / (float)H;
        ray.direction.x = ray.direction.x - stepx * (x - (W / 2));
        ray.direction.y = ray.direction.y + stepy * (y - (H / 2));
    }
    vectorRotation(ray.origin, rotationCenter, angles);
    vectorRotation(ray.direction, rotationCenter, angles);

    V4 color = {0.f, 0.f, 0.f, 0.f};
    RayOT r = ray;
    if (antialiasingActivated)
        for (int I = 0; I < 4; ++I)
        {
            r.origin.x += AA[I][0]; r.origin.y += AA[I][1];
            color += launchRayTracing(c, index, r, dof, ids[index]);
        }
    else if (si.pathTracingIteration >= B200_NB_MAX_ITERATIONS)
    {
        r.direction.x += AA[si.pathTracingIteration % 4][0];
        r.direction.y += AA[si.pathTracingIteration % 4][1];
    }
    color += launchRayTracing(c, index, r, dof, ids[index]);
    if (si.advancedIllumination == B200_AI_RANDOM)
    {
        int rindex = (index + si.timestamp) % c.randomTableSize;
        color += v4(si.backgroundColor) * c.rnd(rindex) * 5.f;
    }
    if (antialiasingActivated) color /= 5.f;
    if (si.pathTracingIteration == 0) post[index].colorInfo.w = dof;
    if (si.pathTracingIteration <= B200_NB_MAX_ITERATIONS)
    {
        post[index].colorInfo.x = color.x; post[index].colorInfo.y = color.y; post[index].colorInfo.z = color.z;
        post[index].sceneInfo.x = color.x; post[index].sceneInfo.y = color.y; post[index].sceneInfo.z = color.z;
    }
    else
    {
        post[index].sceneInfo.x = (ids[index].z > 0) ? fmaxf_(post[index].sceneInfo.x, color.x) : color.x;
        post[index].sceneInfo.y = (ids[index].z > 0) ? fmaxf_(post[index].sceneInfo.y, color.y) : color.y;
        post[index].sceneInfo.z = (ids[index].z > 0) ? fmaxf_(post[index].sceneInfo.z, color.z) : color.z;
        post[index].colorInfo.x += post[index].sceneInfo.x;
        post[index].colorInfo.y += post[index].sceneInfo.y;
        post[index].colorInfo.z += post[index].sceneInfo.z;
    }
}

inline void addCounters(oracle_Counters& a, const oracle_Counters& b)
{
    uint64_t* pa = reinterpret_cast<uint64_t*>(&a);
    const uint64_t* pb = reinterpret_cast<const uint64_t*>(&b);
    for (size_t i = 0; i < sizeof(oracle_Counters) / sizeof(uint64_t); ++i) pa[i] += pb[i];
}
} // namespace

extern "C" {

int oracle_abi_version() { return 1; }

// analysis: start / stop logging closest-hit rays (see g_rayLog); returns the number of rays logged so far
unsigned long long oracle_ray_log(float* buffer, unsigned long long capacityRays)
{
    const unsigned long long n = g_rayLogCount;
    g_rayLog = buffer; g_rayLogCapacity = capacityRays; g_rayLogCount = 0;
    return n;
}

// analysis: node visits of an ideal front-to-back walk of each ray through the engine's 4-wide tree (128-byte records: rows
// lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] refs[4]; ref >= 0 inner node, < 0 leaf / empty) — the inner nodes whose
// box the ray crosses before its hit (every crossed node for a miss).  A lower bound of what a culling walk visits, the same
// for any order in which rays are grouped into warps.
// stackAt (optional): entries on the walk's stack after `cutAfter` visits (0 if the walk is over by then), maxStack (optional): the
// deepest the stack gets — how much state a walk parked after that many node rounds would have to carry.
void oracle_walk_visits_ex(const float* nodes, int nbNodes, const float* rays9, unsigned long long nRays, unsigned int* visits,
                           int cutAfter, unsigned int* stackAt, unsigned int* maxStack);
void oracle_walk_visits(const float* nodes, int nbNodes, const float* rays9, unsigned long long nRays, unsigned int* visits)
{
    oracle_walk_visits_ex(nodes, nbNodes, rays9, nRays, visits, 0, nullptr, nullptr);
}
void oracle_walk_visits_ex(const float* nodes, int nbNodes, const float* rays9, unsigned long long nRays, unsigned int* visits,
                           int cutAfter, unsigned int* stackAt, unsigned int* maxStack)
{
#pragma omp parallel for schedule(dynamic, 1024)
    for (long long k = 0; k < (long long)nRays; ++k)
    {
        const float* w = rays9 + 9 * k;
        const float ox = w[2], oy = w[3], oz = w[4];
        const float dx = w[5] - ox, dy = w[6] - oy, dz = w[7] - oz;
        const float len = sqrtf(dx * dx + dy * dy + dz * dz);
        const float tLimit = (w[8] > 0.f && len > 0.f) ? w[8] / len : 3.0e38f;
        const float ix = dx != 0.f ? 1.f / dx : 1.f, iy = dy != 0.f ? 1.f / dy : 1.f, iz = dz != 0.f ? 1.f / dz : 1.f;
        int stack[256];
        int sp = 0;
        unsigned int count = 0, deepest = 0, atCut = 0;
        if (nbNodes > 0) stack[sp++] = 0;
        while (sp > 0)
        {
            if ((unsigned int)sp > deepest) deepest = (unsigned int)sp;
            if ((int)count == cutAfter) atCut = (unsigned int)sp;
            const int node = stack[--sp];
            ++count;
            const float* n = nodes + 32 * (size_t)node;
            for (int c = 0; c < 4; ++c)
            {
                int ref;
                memcpy(&ref, n + 24 + c, sizeof(int));
                if (ref < 0) continue;
                float t0x = (n[c] - ox) * ix, t1x = (n[12 + c] - ox) * ix;
                float t0y = (n[4 + c] - oy) * iy, t1y = (n[16 + c] - oy) * iy;
                float t0z = (n[8 + c] - oz) * iz, t1z = (n[20 + c] - oz) * iz;
                if (t0x > t1x) { const float t = t0x; t0x = t1x; t1x = t; }
                if (t0y > t1y) { const float t = t0y; t0y = t1y; t1y = t; }
                if (t0z > t1z) { const float t = t0z; t0z = t1z; t1z = t; }
                const float tmin = fmaxf(fmaxf(fmaxf(t0x, t0y), t0z), 0.f), tmax = fminf(fminf(fminf(t1x, t1y), t1z), tLimit);
                if (tmin <= tmax && sp < 256 && ref < nbNodes) stack[sp++] = ref;
            }
        }
        visits[k] = count;
        if (stackAt) stackAt[k] = atCut;
        if (maxStack) maxStack[k] = deepest;
    }
}

void oracle_render(const oracle_Scene* s, const b200_SceneInfo* sceneInfo, const b200_PostProcessingInfo* postInfo,
                   const float* eye, const float* target, const float* angles, b200_PostProcessingBuffer* post,
                   b200_int4* ids, unsigned char* bitmap, int rowBegin, int rowEnd, int rowStride, int nThreads,
                   oracle_Counters* counters)
{
    oracle_Counters total;
    memset(&total, 0, sizeof(total));
    const int W = sceneInfo->size.x, H = sceneInfo->size.y;
    if (rowEnd > H) rowEnd = H;
    if (rowStride < 1) rowStride = 1;
    const int nRows = (rowEnd - rowBegin + rowStride - 1) / rowStride;
#pragma omp parallel num_threads(nThreads > 0 ? nThreads : 1)
    {
        oracle_Counters local;
        memset(&local, 0, sizeof(local));
        Ctx c = {*sceneInfo, *postInfo, s->boxes, s->nbBoxes, s->primitives, s->nbPrimitives, s->materials, s->nbMaterials,
                 s->lightInformation, s->lightInformationSize, s->nbLamps, s->textures, s->randoms, s->randomTableSize, &local};
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < nRows; ++k)
        {
            const int y = rowBegin + k * rowStride;
            for (int x = 0; x < W; ++x)
            {
                g_currentPixel = y * W + x;
                renderPixel(c, x, y, eye, target, angles, post, ids);
                // k_default, CudaRayTracer.cu:1057-1073
                const int index = y * W + x;
                V4 localColor = v4(post[index].colorInfo);
                if (sceneInfo->pathTracingIteration > B200_NB_MAX_ITERATIONS)
                    localColor /= (float)(sceneInfo->pathTracingIteration - B200_NB_MAX_ITERATIONS + 1);
                makeColor(*sceneInfo, localColor, bitmap, index);
            }
        }
#pragma omp critical
        addCounters(total, local);
    }
    if (counters) *counters = total;
}

// Post-processing effects, CudaRayTracer.cu:1081-1358 (k_depthOfField, k_ambiantOcclusion, k_radiosity, k_filter,
// k_cartoon): each reads the float accumulation buffer (and ids.z for radiosity) of the whole frame and rewrites the
// RGB8 bitmap; run after the frame's ray pass when postInfo->type != ppe_none (cudaRender's second switch, :1853-1886).
void oracle_post_process(const oracle_Scene* s, const b200_SceneInfo* sceneInfo, const b200_PostProcessingInfo* postInfo,
                         const b200_PostProcessingBuffer* post, const b200_int4* ids, unsigned char* bitmap)
{
    const b200_SceneInfo& si = *sceneInfo;
    const b200_PostProcessingInfo& pp = *postInfo;
    const int W = si.size.x, H = si.size.y, wh = W * H;
    const int iter = si.pathTracingIteration;
    const float* randoms = s->randoms;
    if (pp.type == B200_PPE_NONE) return;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
        {
            const int index = y * W + x;
            V4 out = {0.f, 0.f, 0.f, 0.f};
            switch (pp.type)
            {
            case B200_PPE_DEPTH_OF_FIELD: // :1081-1119
            {
                const float depth = fabsf(post[index].colorInfo.w - pp.param1) / si.viewDistance;
                for (int i = 0; i < pp.param3; ++i)
                {
                    const int ix = i % wh, iy = (i + 1000) % wh;
                    const int xx = (int)(x + depth * randoms[ix] * pp.param2);
                    const int yy = (int)(y + depth * randoms[iy] * pp.param2);
                    if (xx >= 0 && xx < W && yy >= 0 && yy < H)
                    {
                        const int li = yy * W + xx;
                        if (li >= 0 && li < wh) out += v4(post[li].colorInfo);
                    }
                    else
                        out += v4(post[index].colorInfo);
                }
                out /= (float)pp.param3;
                if (iter > B200_NB_MAX_ITERATIONS) out /= (float)(iter - B200_NB_MAX_ITERATIONS + 1);
                break;
            }
            case B200_PPE_AMBIENT_OCCLUSION: // :1127-1180
            {
                out = v4(post[index].colorInfo);
                const float depth = out.w;
                float occ = 0.f, c = 0.f;
                int i = 0;
                for (int X = -16; X < 16; X += 2)
                    for (int Y = -16; Y < 16; Y += 2)
                    {
                        const int ix = i % wh, iy = (i + 100) % wh;
                        ++i;
                        c += 1.f;
                        const int xx = (int)(x + (X * pp.param2 * randoms[ix] / 10.f));
                        const int yy = (int)(y + (Y * pp.param2 * randoms[iy] / 10.f));
                        if (xx >= 0 && xx < W && yy >= 0 && yy < H)
                        {
                            if (post[yy * W + xx].colorInfo.w >= depth) occ += 1.f;
                        }
                        else
                            occ += 1.f;
                    }
                occ /= c;
                occ += 0.3f;
                if (occ < 1.f) { out.x *= occ; out.y *= occ; out.z *= occ; }
                if (iter > B200_NB_MAX_ITERATIONS) out /= (float)(iter - B200_NB_MAX_ITERATIONS + 1);
                saturateVector(out);
                break;
            }
            case B200_PPE_RADIOSITY: // :1188-1228
            {
                const int div = (iter > B200_NB_MAX_ITERATIONS) ? (iter - B200_NB_MAX_ITERATIONS + 1) : 1;
                for (int i = 0; i < pp.param3; ++i)
                {
                    const int ix = (i + iter) % wh, iy = (i + 100 + iter) % wh;
                    const int xx = (int)(x + randoms[ix] * pp.param2);
                    const int yy = (int)(y + randoms[iy] * pp.param2);
                    out += v4(post[index].colorInfo);
                    if (xx >= 0 && xx < W && yy >= 0 && yy < H)
                    {
                        const int li = yy * W + xx;
                        V4 light = v4(post[li].colorInfo);
                        const float k = (float)ids[li].z;
                        light.x = light.x * k / 256.f; light.y = light.y * k / 256.f; light.z = light.z * k / 256.f; light.w = light.w * k / 256.f;
                        out += light;
                    }
                }
                out /= (float)pp.param3;
                out /= (float)div;
                saturateVector(out);
                break;
            }
            case B200_PPE_FILTER: // :1236-1330
            {
                static const int fsize[6] = {3, 5, 3, 3, 5, 5};
                static const float factor[6] = {1.f, 1.f, 1.f, 1.f, 0.2f, 0.125f}, bias[6] = {128.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                static const float kern[6][5][5] = {
                    {{-1, -1, 0, 0, 0}, {-1, 0, 1, 0, 0}, {0, 1, 1, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},                          // emboss
                    {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {-1, -1, 2, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},                           // find edges
                    {{-1, -1, -1, 0, 0}, {-1, 9, -1, 0, 0}, {-1, -1, -1, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},                     // sharpen
                    {{0, 0.2f, 0, 0, 0}, {0.2f, 0.2f, 0.2f, 0, 0}, {0, 0.2f, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}},              // blur
                    {{1, 0, 0, 0, 0}, {0, 1, 0, 0, 0}, {0, 0, 1, 0, 0}, {0, 0, 0, 1, 0}, {0, 0, 0, 0, 1}},                             // motion blur
                    {{-1, -1, -1, -1, -1}, {-1, 2, 2, 2, -1}, {-1, 2, 8, 2, -1}, {-1, 2, 2, 2, -1}, {-1, -1, -1, -1, -1}}};           // subtle sharpen
                if ((unsigned int)pp.param3 < 6u)
                {
                    const int f = pp.param3, n = fsize[f];
                    V4 acc = {0.f, 0.f, 0.f, 0.f};
                    for (int fx = 0; fx < n; ++fx)
                        for (int fy = 0; fy < n; ++fy)
                        {
                            const int imx = (x - n / 2 + fx + W) % W, imy = (y - n / 2 + fy + H) % H;
                            V4 c = v4(post[imy * W + imx].colorInfo);
                            if (iter > B200_NB_MAX_ITERATIONS) c /= (float)(iter - B200_NB_MAX_ITERATIONS + 1);
                            acc.x += c.x * kern[f][fx][fy]; acc.y += c.y * kern[f][fx][fy]; acc.z += c.z * kern[f][fx][fy];
                        }
                    out.x += fminf(fmaxf(factor[f] * acc.x + bias[f] / 255.f, 0.f), 1.f);
                    out.y += fminf(fmaxf(factor[f] * acc.y + bias[f] / 255.f, 0.f), 1.f);
                    out.z += fminf(fmaxf(factor[f] * acc.z + bias[f] / 255.f, 0.f), 1.f);
                }
                saturateVector(out);
                break;
            }
            case B200_PPE_CARTOON: // :1338-1357
            {
                const float depth = si.viewDistance / fabsf(post[index].colorInfo.w - pp.param1);
                out = V4{depth, depth, depth, 0.f};
                saturateVector(out);
                break;
            }
            default: // k_default
                out = v4(post[index].colorInfo);
                if (iter > B200_NB_MAX_ITERATIONS) out /= (float)(iter - B200_NB_MAX_ITERATIONS + 1);
                break;
            }
            out.w = 1.f;
            makeColor(si, out, bitmap, index);
        }
}

// flops per unit of work, applied to counts taken in reference traversal order.
double oracle_algorithmic_flops(const oracle_Counters* k)
{
    return 15.0 * (double)k->rays + 12.0 * (double)k->box_tests + 35.0 * (double)k->sphere_tests +
           90.0 * (double)(k->cylinder_tests + k->cone_tests) + 47.0 * (double)k->triangle_tests +
           10.0 * (double)k->plane_tests + 60.0 * (double)k->ellipsoid_tests + 150.0 * (double)k->shade_calls;
}
}
