// shade.cuh — shading, texture lookup and the bounce loop of the ray-propagation path.
// Semantics: reference CUDA engine (paths relative to /root/reference/solr/engines/cuda/).
#pragma once
#include "trace.cuh"

#define SB_PI 3.14159265358979323846f

// ---------------------------------------------------------------- texture maps (TextureMapping.cuh:30-116)
SB_DEV void texMaps(const b200_Material& m, const int index, float3& normal, float4& specular, float4& attributes,
                    float4& adv)
{
    const unsigned char* t = cS.tex;
    // bump map (:45-57) only computes a local `value`; nothing observable.
    if (m.textureIds.y != B200_TEXTURE_NONE) // normal map :30-40
    {
        const int i = m.textureOffset.y + index;
        normal.x -= 3.f * (t[i] / 256.f - 0.5f);
        normal.y -= 3.f * (t[i + 1] / 256.f - 0.5f);
        normal.z = 0.f;
    }
    if (m.textureIds.w != B200_TEXTURE_NONE) // specular map :62-73
    {
        const int i = m.textureOffset.w + index;
        specular.x = t[i] / 256.f;
        specular.y = 1000.f * t[i + 1] / 256.f;
        specular.z = t[i + 2] / 256.f;
    }
    if (m.advancedTextureIds.x != B200_TEXTURE_NONE) // reflection map :78-87
    {
        const int i = m.advancedTextureOffset.x + index;
        attributes.x *= (t[i] + t[i + 1] + t[i + 2]) / 768.f;
    }
    if (m.advancedTextureIds.y != B200_TEXTURE_NONE) // transparency map :92-102
    {
        const int i = m.advancedTextureOffset.y + index;
        attributes.y *= (t[i] + t[i + 1] + t[i + 2]) / 768.f;
    }
    if (m.advancedTextureIds.z != B200_TEXTURE_NONE) // ambient occlusion map :107-116
    {
        const int i = m.advancedTextureOffset.z + index;
        adv.x = (t[i] + t[i + 1] + t[i + 2]) / 768.f;
    }
}

SB_DEV void fetchTexel(const b200_Material& m, const int u, const int v, float4& result, float3& normal,
                       float4& specular, float4& attributes, float4& adv)
{
    const int A = (v * m.textureMapping.x + u) * m.textureMapping.w;
    const int B = m.textureMapping.x * m.textureMapping.y * m.textureMapping.w;
    const int index = A % B;
    const int i = m.textureOffset.x + index;
    result.x = cS.tex[i] / 256.f; result.y = cS.tex[i + 1] / 256.f; result.z = cS.tex[i + 2] / 256.f;
    texMaps(m, index, normal, specular, attributes, adv);
}

// TextureMapping.cuh:118-160
__device__ __noinline__ void juliaSet(const b200_Material& m, const float x, const float y, float4& color)
{
    const float W = (float)m.textureMapping.x, H = (float)m.textureMapping.y;
    const float cRe = -0.7f + 0.4f * sinf(cSI.timestamp / 1500.f);
    const float cIm = 0.27015f + 0.4f * cosf(cSI.timestamp / 2000.f);
    float newRe = 1.5f * (x - W / 2.f) / (0.5f * W);
    float newIm = (y - H / 2.f) / (0.5f * H);
    int n;
    const float maxIterations = 40.f + cSI.pathTracingIteration;
    for (n = 0; n < maxIterations; n++)
    {
        const float oldRe = newRe, oldIm = newIm;
        newRe = oldRe * oldRe - oldIm * oldIm + cRe;
        newIm = 2.f * oldRe * oldIm + cIm;
        if ((newRe * newRe + newIm * newIm) > 4.f) break;
    }
    color.x = 1.f - color.x * (n / maxIterations);
    color.y = 1.f - color.y * (n / maxIterations);
    color.z = 1.f - color.z * (n / maxIterations);
    color.w = 1.f - (n / maxIterations);
}

// TextureMapping.cuh:162-198 (Im_factor is a double there)
__device__ __noinline__ void mandelbrotSet(const b200_Material& m, const float x, const float y, float4& color)
{
    const float W = (float)m.textureMapping.x, H = (float)m.textureMapping.y;
    const float MinRe = -2.f, MaxRe = 1.f, MinIm = -1.2f;
    const float MaxIm = MinIm + (MaxRe - MinRe) * H / W;
    const float Re_factor = (MaxRe - MinRe) / (W - 1.f);
    const double Im_factor = (MaxIm - MinIm) / (H - 1.f);
    const float maxIterations = B200_NB_MAX_ITERATIONS + cSI.pathTracingIteration;
    const float c_im = MaxIm - y * Im_factor;
    const float c_re = MinRe + x * Re_factor;
    float Z_re = c_re, Z_im = c_im;
    bool isInside = true;
    unsigned n;
    for (n = 0; isInside && n < maxIterations; ++n)
    {
        const float Z_re2 = Z_re * Z_re, Z_im2 = Z_im * Z_im;
        if (Z_re2 + Z_im2 > 4.f) isInside = false;
        Z_im = 2.f * Z_re * Z_im + c_im;
        Z_re = Z_re2 - Z_im2 + c_re;
    }
    color.x = 1.f - color.x * (n / maxIterations);
    color.y = 1.f - color.y * (n / maxIterations);
    color.z = 1.f - color.z * (n / maxIterations);
    color.w = 1.f - (n / maxIterations);
}

// TextureMapping.cuh:205-286
__device__ __noinline__ float4 triangleUVMapping(const b200_Primitive& p, const float3 areas,
                                                 float3& normal, float4& specular, float4& attributes, float4& adv)
{
    const b200_Material& m = cS.mats[p.materialId];
    float4 result = f4(m.color.x, m.color.y, m.color.z, m.color.w);
    const float sum = areas.x + areas.y + areas.z;
    const float Tx = (p.vt0.x * areas.x + p.vt1.x * areas.y + p.vt2.x * areas.z) / sum;
    const float Ty = (p.vt0.y * areas.x + p.vt1.y * areas.y + p.vt2.y * areas.z) / sum;
    float mox = 0.f, moy = 0.f;
    if (m.attributes.y == 1)
    {
        mox = m.mappingOffset.x * cSI.timestamp;
        moy = m.mappingOffset.y * cSI.timestamp;
    }
    int u = Tx * m.textureMapping.x + mox;
    int v = Ty * m.textureMapping.y + moy;
    u = u % m.textureMapping.x;
    v = v % m.textureMapping.y;
    if (u >= 0 && u < m.textureMapping.x && v >= 0 && v < m.textureMapping.y)
    {
        switch (m.textureIds.x)
        {
        case B200_TEXTURE_MANDELBROT: mandelbrotSet(m, u, v, result); break;
        case B200_TEXTURE_JULIA: juliaSet(m, u, v, result); break;
        default: fetchTexel(m, u, v, result, normal, specular, attributes, adv);
        }
    }
    return result;
}

// TextureMapping.cuh:294-349
__device__ __noinline__ float4 sphereUVMapping(const b200_Primitive& p, const float3 I, float3& normal, float4& specular,
                                               float4& attributes, float4& adv)
{
    const b200_Material& m = cS.mats[p.materialId];
    float4 result = f4(m.color.x, m.color.y, m.color.z, m.color.w);
    const float3 d = normalize(I - f3(p.p0.x, p.p0.y, p.p0.z));
    const float U = ((atan2f(d.x, d.z) / SB_PI) + 1.f) * .5f;
    const float V = (asinf(d.y) / SB_PI) + .5f;
    int u = m.textureMapping.x * (U * p.vt1.x);
    int v = m.textureMapping.y * (V * p.vt1.y);
    if (m.textureMapping.x != 0) u = u % m.textureMapping.x;
    if (m.textureMapping.y != 0) v = v % m.textureMapping.y;
    if (u >= 0 && u < m.textureMapping.x && v >= 0 && v < m.textureMapping.y)
        fetchTexel(m, u, v, result, normal, specular, attributes, adv);
    return result;
}

// TextureMapping.cuh:357-447 (non-Kinect branch)
__device__ __noinline__ float4 cubeMapping(const b200_Primitive& p, float3 I, float3& normal,
                                           float4& specular, float4& attributes, float4& adv)
{
    const b200_Material& m = cS.mats[p.materialId];
    float4 result = f4(m.color.x, m.color.y, m.color.z, m.color.w);
    int u = ((p.type == B200_PT_CHECKBOARD) || (p.type == B200_PT_XZPLANE) || (p.type == B200_PT_XYPLANE))
                ? (I.x - p.p0.x + p.size.x) : (I.z - p.p0.z + p.size.z);
    int v = ((p.type == B200_PT_CHECKBOARD) || (p.type == B200_PT_XZPLANE)) ? (I.z + p.p0.z + p.size.z) : (I.y - p.p0.y + p.size.y);
    if (m.textureMapping.x != 0) u = u % m.textureMapping.x;
    if (m.textureMapping.y != 0) v = v % m.textureMapping.y;
    if (u >= 0 && u < m.textureMapping.x && v >= 0 && v < m.textureMapping.x)
    {
        switch (m.textureIds.x)
        {
        case B200_TEXTURE_MANDELBROT: mandelbrotSet(m, u, v, result); break;
        case B200_TEXTURE_JULIA: juliaSet(m, u, v, result); break;
        default: fetchTexel(m, u, v, result, normal, specular, attributes, adv);
        }
    }
    return result;
}

// GeometryIntersections.cuh:87-151
__device__ __noinline__ float4 skyboxMapping(const float3 origin, const float3 target)
{
    const b200_Material& m = cS.mats[cSI.skyboxMaterialId];
    float4 result = f4(m.color.x, m.color.y, m.color.z, m.color.w);
    const float3 dir = normalize(target - origin);
    const float a = 2.f * dot(dir, dir);
    const float b = 2.f * dot(origin, dir);
    const float c = dot(origin, origin) - (cSI.skyboxRadius * cSI.skyboxRadius);
    const float d = b * b - 2.f * a * c;
    if (d <= 0.f || a == 0.f) return result;
    const float r = sqrtf(d);
    const float t1 = (-b - r) / a;
    const float t2 = (-b + r) / a;
    if (t1 <= cSI.geometryEpsilon && t2 <= cSI.geometryEpsilon) return result;
    float t = 0.f;
    if (t1 <= cSI.geometryEpsilon) t = t2;
    else if (t2 <= cSI.geometryEpsilon) t = t1;
    else t = (t1 < t2) ? t1 : t2;
    if (t < cSI.geometryEpsilon) return result;
    const float3 I = normalize(origin + t * dir);
    const float U = ((atan2f(I.x, I.z) / SB_PI) + 1.f) * .5f;
    const float V = (asinf(I.y) / SB_PI) + .5f;
    int u = int(m.textureMapping.x * U);
    int v = int(m.textureMapping.y * V);
    if (m.textureMapping.x != 0) u %= m.textureMapping.x;
    if (m.textureMapping.y != 0) v %= m.textureMapping.y;
    if (u >= 0 && u < m.textureMapping.x && v >= 0 && v < m.textureMapping.y)
    {
        const int A = (v * m.textureMapping.x + u) * m.textureMapping.w;
        const int B = m.textureMapping.x * m.textureMapping.y * m.textureMapping.w;
        const int i = m.textureOffset.x + A % B;
        result.x = cS.tex[i] / 256.f; result.y = cS.tex[i + 1] / 256.f; result.z = cS.tex[i + 2] / 256.f;
    }
    return result;
}

// GeometryShaders.cuh:36-124
SB_DEV float4 intersectionShader(const b200_Primitive& p, const b200_Material& m, const float3 I,
                                 const float3 areas, float3& normal, float4& specular, float4& attributes, float4& adv)
{
    float4 col = f4(m.color.x, m.color.y, m.color.z, 0.f);
    const bool textured = m.textureIds.x != B200_TEXTURE_NONE;
    if (cSI.extendedGeometry)
    {
        switch (p.type)
        {
        case B200_PT_CONE: case B200_PT_CYLINDER: case B200_PT_ENVIRONMENT: case B200_PT_SPHERE: case B200_PT_ELLIPSOID:
            if (textured) col = sphereUVMapping(p, I, normal, specular, attributes, adv);
            break;
        case B200_PT_CHECKBOARD:
            if (textured)
                col = cubeMapping(p, I, normal, specular, attributes, adv);
            else
            {
                const int x = cSI.viewDistance + ((I.x - p.p0.x) / p.size.x);
                const int z = cSI.viewDistance + ((I.z - p.p0.z) / p.size.x);
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
            if (textured) col = cubeMapping(p, I, normal, specular, attributes, adv);
            break;
        case B200_PT_TRIANGLE:
            if (textured) col = triangleUVMapping(p, areas, normal, specular, attributes, adv);
            break;
        }
    }
    else if (textured)
        col = triangleUVMapping(p, areas, normal, specular, attributes, adv);
    return col;
}

#ifdef SOLR_DEBUG_SHADOWCMP
__device__ float g_dbgRec[64 * 32];
__device__ int g_dbgN;
#endif
SB_DEV float rnd(const int i) { return __ldg(cS.randoms + i); }

// Shadow ray of one shaded hit: packet walk when the policy says so for this ray class, else per lane.
SB_DEV float4 traceShadow(const float3 center, const float3 I, const int lightId, const int iteration, const int objectId,
                          const bool need, const bool packet, Counters& cnt)
{
    float4 sh = f4(0.f, 0.f, 0.f, 0.f);
    if (need) cnt.rays++;
    if (packet)
    {
        if (__any_sync(FULL_MASK, need)) sh = shadowWalkPacket(center, I, lightId, iteration, objectId, need);
    }
#if UW_GROUP
    else if (cS.nbUWide > 0)
    {
        // every lane calls the group walk; |direction| = distance to the lamp; any-hit is exact when every caster is opaque
        const float3 d = center - I;
        const bool grouped = need && cS.opaqueShadows && dot(d, d) >= 1.0002f;
        const WalkOut w = groupWalk(grouped, UW_SHADOW, I + normalize(d) * cSI.rayEpsilon, d, iteration, 0, lightId, objectId);
        if (grouped) sh.w = w.shadow;
        if (need && (!grouped || w.shadow < 0.f)) sh = shadowWalkWide(center, I, lightId, iteration, objectId); // < 0: overflow in a degenerate tree
    }
#endif
    else if (need)
    {
        if (cS.nbUWide > 0)
        {
            // |direction| = distance to the lamp; any-hit is exact when every caster is opaque
            const float3 d = center - I;
            bool ordered = true;
            if (cS.opaqueShadows && dot(d, d) >= 1.0002f)
            {
                sh.w = unorderedWalk(UW_SHADOW, I + normalize(d) * cSI.rayEpsilon, d, iteration, 0, lightId, objectId).shadow;
                ordered = sh.w < 0.f; // stack overflow in a degenerate tree
#ifdef SOLR_DEBUG_SHADOWCMP
                if (!ordered)
                {
                    const float4 so = shadowWalkWide(center, I, lightId, iteration, objectId);
                    if (so.w != sh.w)
                    {
                        const int slot = atomicAdd(&g_dbgN, 1);
                        if (slot < 64)
                        {
                            float* o = g_dbgRec + 32 * slot;
                            Ray r; makeRay(r, I + normalize(d) * cSI.rayEpsilon, d);
                            o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = d.x; o[4] = d.y; o[5] = d.z; o[6] = so.w; o[7] = sh.w;
                            o[8] = (float)lightId; o[9] = (float)objectId;
                            int nb = 0;
                            const float lenOL = length(r.d);
                            for (int idx = 0; idx < cS.nbPrimitives && nb < 4; ++idx)
                            {
                                const int meta = cS.meta[idx];
                                const int origIndex = cS.prims[idx].index;
                                if (PM_FAST(meta) != 0 || origIndex == lightId || origIndex == objectId) continue;
                                float3 P; int fl; float ps;
                                if (!primitiveTest(idx, meta, r, P, fl, ps)) continue;
                                const float l = length(P - r.o);
                                if (!(l > cSI.geometryEpsilon && l < lenOL)) continue;
                                const int leaf = cS.primLeaf[idx];
                                float lt;
                                const bool lp = slabT(cS.leafRecs[2 * leaf], cS.leafRecs[2 * leaf + 1], r, cSI.viewDistance, lt);
                                o[10 + 5 * nb] = (float)idx; o[11 + 5 * nb] = l; o[12 + 5 * nb] = (float)PM_TYPE(meta); o[13 + 5 * nb] = lp ? 1.f : 0.f; o[14 + 5 * nb] = lt;
                                ++nb;
                            }
                            o[30] = (float)nb; o[31] = lenOL;
                        }
                    }
                }
#endif
            }
            if (ordered) sh = shadowWalkWide(center, I, lightId, iteration, objectId);
        }
        else
            sh = (cS.nbWide > 0) ? shadowWalkWide(center, I, lightId, iteration, objectId) : shadowWalk(center, I, lightId, iteration, objectId);
    }
    return sh;
}

// GeometryIntersections.cuh:916-1080.  Warp-uniform form: every lane of the warp calls it, `act` says whether
// this lane has a hit to shade; the shadow walk in the lamp loop is hoisted out of the per-lane conditions so
// that the 32 shadow rays of a tile can be walked as one packet.
SB_DEV float4 primitiveShader(const bool act, const int index, const float3 origin, float3& normal, const int objectIdIn, const float3 I,
                              const float3 areas, float4& closestColor, const int iteration, float& shadowIntensity, float4& totalBlinn,
                              float4& attributes, const bool packetShadow, Counters& cnt)
{
    const int objectId = objectIdIn;
    int primIndex = -1;
    float4 mInner = f4(0.f, 0.f, 0.f, 0.f);
    float matTransparency = 0.f;
    bool aoMapped = false;
    float4 lampsColor = f4(0.f, 0.f, 0.f, 0.f);
    float4 intersectionColor = f4(0.f, 0.f, 0.f, 0.f);
    float4 adv = f4(0.f, 0.f, 0.f, 0.f);
    float4 specular = f4(0.f, 0.f, 0.f, 0.f);
    bool lit = false; // this lane runs the lamp loop
    if (act)
    {
        const b200_Primitive& primitive = cS.prims[objectId];
        const b200_Material& material = cS.mats[primitive.materialId];
        primIndex = primitive.index;
        mInner = *reinterpret_cast<const float4*>(&material.innerIllumination);
        matTransparency = material.transparency;
        aoMapped = material.advancedTextureIds.z != B200_TEXTURE_NONE;
        specular = f4(material.specular.x, material.specular.y, material.specular.z, 0.f);
        shadowIntensity = 0.f;
        float3 bumpNormal = f3(0.f, 0.f, 0.f);
        intersectionColor = intersectionShader(primitive, material, I, areas, bumpNormal, specular, attributes, adv);
        normal += bumpNormal;
        normal = normalize(normal);
        lit = material.attributes.z != 1; // wireframe returns the constant colour (:947-951)
    }
    if (cSI.graphicsLevel > B200_GL_NO_SHADING)
    {
        if (lit) closestColor *= mInner.x;
        const int nbLights = cS.lightInfoSize;
        for (int cpt = 0; cpt < nbLights; ++cpt)
        {
            // the reference's lamp loop runs lightInformationSize passes over the SAME lamp (:956-960)
            const int cptLamp = (cSI.pathTracingIteration >= B200_NB_MAX_ITERATIONS) ? (cSI.pathTracingIteration % nbLights) : 0;
            const b200_LightInformation& li = cS.lights[cptLamp];
            const float4 liColor = *reinterpret_cast<const float4*>(&li.color);
            const int lightPrimId = li.primitiveId;
            const b200_Material& m = cS.mats[li.materialId];
            const float4 lInner = *reinterpret_cast<const float4*>(&m.innerIllumination);
            const int t = (index + cSI.timestamp) % (cS.randomTableSize - 3);
            float3 center = f3(li.location.x, li.location.y, li.location.z);
            float3 lightRay = f3(0.f, 0.f, 0.f);
            float lightRayLength = 0.f, lambert = 0.f;
            bool inRange = false, needShadow = false;
            if (lit && lightPrimId != primIndex)
            {
                if (cSI.pathTracingIteration >= B200_NB_MAX_ITERATIONS)
                {
                    const float a = lInner.y * 10.f * cSI.pathTracingIteration / cSI.maxPathTracingIterations;
                    center.x += rnd(t) * a; center.y += rnd(t + 1) * a; center.z += rnd(t + 2) * a;
                }
                lightRay = center - I;
                lightRayLength = length(lightRay);
                inRange = lightRayLength < lInner.z;
                if (inRange)
                {
                    lightRay = normalize(lightRay);
                    lambert = mInner.x + dot(normal, lightRay);
                    needShadow = lambert > 0.f && cSI.graphicsLevel > 3 && iteration < 4 && mInner.x == 0.f;
                }
            }
            const float4 sh = traceShadow(center, I, lightPrimId, iteration, objectId, needShadow, packetShadow, cnt);
            if (inRange)
            {
                float4 shadowColor = f4(0.f, 0.f, 0.f, 0.f);
                if (needShadow) { shadowColor.x = sh.x; shadowColor.y = sh.y; shadowColor.z = sh.z; shadowIntensity = sh.w; }
                float photonEnergy = sqrtf(lightRayLength / lInner.z);
                photonEnergy = (photonEnergy > 1.f) ? 1.f : photonEnergy;
                photonEnergy = (photonEnergy < 0.f) ? 0.f : photonEnergy;
                lambert *= (lambert < 0.f) ? -matTransparency : 1.f;
                if (li.materialId != B200_MATERIAL_NONE)
                    lambert *= lInner.x;
                else
                    lambert *= liColor.w;
                if (mInner.w != 0.f) lambert *= (1.f + rnd(t) * mInner.w * 100.f);
                lambert *= (1.f - shadowIntensity);
                lambert += cSI.backgroundColor.w;
                lambert *= (1.f - photonEnergy);
                lampsColor += lambert * liColor - shadowColor;
                if (cSI.graphicsLevel > 1 && shadowIntensity < cSI.shadowIntensity)
                {
                    const float3 viewRay = normalize(I - origin);
                    float3 blinnDir = lightRay - viewRay;
                    const float temp = sqrtf(dot(blinnDir, blinnDir));
                    if (temp != 0.f)
                    {
                        blinnDir = (1.f / temp) * blinnDir;
                        float blinnTerm = dot(blinnDir, normal);
                        blinnTerm = (blinnTerm < 0.f) ? 0.f : blinnTerm;
                        blinnTerm = specular.x * powf(blinnTerm, specular.y);
                        blinnTerm *= (1.f - photonEnergy);
                        totalBlinn += liColor * liColor.w * blinnTerm;
                        totalBlinn.w = specular.z;
                    }
                }
            }
            if (lit)
            {
                closestColor += intersectionColor * lampsColor;
                if (aoMapped) closestColor *= adv.x;
                saturate4(closestColor);
                saturate4(totalBlinn);
            }
        }
    }
    else if (lit)
        closestColor = intersectionColor;
    return (act && !lit) ? intersectionColor : closestColor;
}

SB_DEV void vectorReflection(float3& r, const float3 i, const float3 n) { r = i - 2.f * dot(i, n) * n; } // VectorUtils.cuh:61-64
SB_DEV void vectorRefraction(float3& refracted, const float3 incident, const float n1, const float3 normal, const float n2) // :73-87
{
    refracted = incident;
    if (n2 != 0.f)
    {
        const float eta = n1 / n2;
        const float c1 = -dot(incident, normal);
        const float cs2 = 1.f - eta * eta * (1.f - c1 * c1);
        if (cs2 >= 0.f) refracted = eta * incident + (eta * c1 - sqrtf(cs2)) * normal;
    }
}

// VectorUtils.cuh:104-142, with the six sin/cos hoisted to the host (they depend on the camera only).
struct Rotation { float cx, cy, cz, sx, sy, sz; };
// The sums of two products below are written with the rounding the reference's build has — first product fused into the
// sum, second rounded on its own (what nvcc emits for a*b + c*d) — as explicit intrinsics: inside a large kernel the
// compiler otherwise picks differently depending on what else uses the products, and a ray that differs in the last
// bit decides grazing hits differently (profiles/r01_history.md, "pinned rounding").
SB_DEV float mulAdd2(const float a, const float b, const float c, const float d) { return __fmaf_rn(a, b, __fmul_rn(c, d)); }  // a*b + c*d
SB_DEV float mulSub2(const float a, const float b, const float c, const float d) { return __fmaf_rn(a, b, -__fmul_rn(c, d)); } // a*b - c*d
// `stereo`: k_3DVisionRenderer's copy of the function fuses the other product in the first sum
SB_DEV void vectorRotation(float3& v, const float3 c, const Rotation& R, const bool stereo = false)
{
    float3 vec = f3(__fadd_rn(v.x, -c.x), __fadd_rn(v.y, -c.y), __fadd_rn(v.z, -c.z));
    float3 res = vec;
    // in the reference's build the differences fuse their first product, the sums their second (k_standardRenderer SASS)
    res.y = mulSub2(vec.y, R.cx, vec.z, R.sx);
    res.z = stereo ? mulAdd2(vec.y, R.sx, vec.z, R.cx) : mulAdd2(vec.z, R.cx, vec.y, R.sx);
    vec = res;
    res.z = mulSub2(vec.z, R.cy, vec.x, R.sy);
    res.x = mulAdd2(vec.x, R.cy, vec.z, R.sy);
    vec = res;
    res.x = mulSub2(vec.x, R.cz, vec.y, R.sz);
    res.y = mulAdd2(vec.y, R.cz, vec.x, R.sz);
    v.x = __fadd_rn(res.x, c.x); v.y = __fadd_rn(res.y, c.y); v.z = __fadd_rn(res.z, c.z);
}

// Closest-hit walk of one ray class: packet walk when the policy says so, else per lane.
SB_DEV Hit traceClosest(const float3 o, const float3 t, const int iteration, const int matId, const bool need, const bool packet,
                        float3& rayNd, Counters& cnt)
{
    Hit hit;
    hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
    if (need) { cnt.rays++; rayNd = normalize(t - o); }
    if (packet)
    {
        if (__any_sync(FULL_MASK, need)) hit = closestHitPacket(o, t, iteration, matId, need);
    }
#if UW_GROUP
    else if (cS.nbUWide > 0)
    {
        hit = closestHitGroup(o, t, iteration, matId, need); // every lane calls
    }
#endif
    else if (need)
    {
        if (cS.nbUWide > 0)
        {
            hit = closestHitOrderIndependent(o, t, iteration, matId);
        }
        else
            hit = (cS.nbWide > 0) ? closestHitWide(o, t, iteration, matId) : closestHit(o, t, iteration, matId);
    }
    return hit;
}

// ---------------------------------------------------------------------------------------------------
// CudaRayTracer.cu:69-408 — the bounce loop.  Its state is a struct and its body one function per stage
// so that two drivers can run it:
//   * launchRayTracing (below): one thread owns a pixel from the first ray to the last, the 32 lanes of a warp (one
//     8x4-pixel tile) step through the passes together (warp-uniform, so the walks can run as packets);
//   * the staged kernels of engine.cu: one launch per pass over a compacted queue of the paths that are still
//     alive, the state parked in global memory in between, so that every lane of every warp carries a ray.
// colors[] / colorContributions[] are the reference's 11-entry local arrays (:92-95), folded back to front at
// :382-385; the store is a template parameter (thread-local arrays, or one global array per pass).
// ---------------------------------------------------------------------------------------------------
struct PathState
{
    float3 curO, curT;
    float initialRefraction;
    int currentMaterialId;
    float4 closestColor;
    float shadowIntensity;
    float4 rBlinn, recursiveBlinn;
    float3 latestIntersection;
    float rayLength;
    float depthOfField;
    int reflectedRays; // pass of the first transparent + reflective hit, or -1
    float3 reflO, reflT;
    float reflectedRatio;
    bool carryon;
    int iteration;
    int idx, idz, idw; // PrimitiveXYIdBuffer x, z, w
    // global-illumination ray of the first hit (:163-175)
    float3 giO, giT;
    float pathTracingRatio;
    bool useGlobalIllumination;
    float4 colorBox;
};

struct LocalColors
{
    float4 c[B200_NB_MAX_ITERATIONS + 1];
    float k[B200_NB_MAX_ITERATIONS + 1];
    SB_DEV float4 color(const int i) const { return c[i]; }
    SB_DEV float contribution(const int i) const { return k[i]; }
    SB_DEV void setColor(const int i, const float4 v) { c[i] = v; }
    SB_DEV void setContribution(const int i, const float v) { k[i] = v; }
};

// one float4 / float array per pass, indexed by path slot
struct GlobalColors
{
    float4* c;
    float* k;
    size_t slot, stride;
    // read past L1 (ld.global.cg): in the fused driver the pass that wrote them may have run on another SM during this same launch
    SB_DEV float4 color(const int i) const { return ldOnce(c + (size_t)i * stride + slot); }
    SB_DEV float contribution(const int i) const { return ldOnce(k + (size_t)i * stride + slot); }
    SB_DEV void setColor(const int i, const float4 v) { stOnce(c + (size_t)i * stride + slot, v); }
    SB_DEV void setContribution(const int i, const float v) { stOnce(k + (size_t)i * stride + slot, v); }
};

SB_DEV int pathMaxIteration()
{
    int m = (cSI.graphicsLevel < B200_GL_REFLECTIONS) ? 1 : cSI.nbRayIterations + cSI.pathTracingIteration;
    return (m > B200_NB_MAX_ITERATIONS) ? B200_NB_MAX_ITERATIONS : m;
}

SB_DEV void pathInit(PathState& s, const float3 rayO, const float3 rayT)
{
    s.curO = rayO; s.curT = rayT;
    s.initialRefraction = 1.f;
    s.currentMaterialId = -2;
    s.closestColor = f4(0.f, 0.f, 0.f, 0.f);
    s.shadowIntensity = 0.f;
    s.rBlinn = f4(0.f, 0.f, 0.f, 0.f);
    s.recursiveBlinn = f4(0.f, 0.f, 0.f, 0.f);
    s.latestIntersection = rayO;
    s.rayLength = 0.f;
    s.depthOfField = cSI.viewDistance;
    s.reflectedRays = -1;
    s.reflO = f3(0.f, 0.f, 0.f); s.reflT = f3(0.f, 0.f, 0.f);
    s.reflectedRatio = 0.f;
    s.carryon = true;
    s.iteration = 0;
    s.idx = -1; s.idz = 0; s.idw = 0;
    s.giO = f3(0.f, 0.f, 0.f); s.giT = f3(0.f, 0.f, 0.f);
    s.pathTracingRatio = 0.f;
    s.useGlobalIllumination = false;
    s.colorBox = f4(0.f, 0.f, 0.f, 0.f);
}

// One pass of the loop at :125-294 for a lane whose own loop condition holds (act); lanes with act == false only keep the
// warp-synchronous walks company.  rayO is the primary origin (first-hit depth).  `given`: the closest hit, when the walk
// was done elsewhere.
template <class Colors>
SB_DEV void pathPass(PathState& s, Colors& C, const int pass, const bool act, const int index, const float3 rayO, const int packetMask,
                     Counters& cnt, const Hit* given = nullptr)
{
    const bool debugBoxes = cSI.renderBoxes != 0;
    float3 areas = f3(0.f, 0.f, 0.f);
    float3 normal = f3(0.f, 0.f, 0.f);
    float3 rayNd = f3(0.f, 0.f, 0.f); // normalize(target - origin) of the walk that produced `hit`
    Hit hit;
    hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
    bool found = false;
    if (debugBoxes)
    {
        if (act) { cnt.rays++; boxDebugWalk(s.curO, s.curT, s.iteration, s.colorBox); }
    }
    else
    {
        if (given)
        {
            // the walk ran before the path was loaded (engine.cu k_stage_pass, k_stage_fused)
            if (act) { cnt.rays++; rayNd = normalize(s.curT - s.curO); hit = *given; }
        }
        else
            hit = traceClosest(s.curO, s.curT, pass, s.currentMaterialId, act, (packetMask & (pass == 0 ? 1 : 2)) != 0, rayNd, cnt);
        found = act && hit.prim >= 0;
    }
    if (act) s.carryon = found;
    float4 attributes = f4(0.f, 0.f, 0.f, 0.f);
    float matInnerX = 0.f, matColorW = 0.f;
    if (found)
    {
        const int meta = __ldg(cS.meta + hit.prim);
        hitNormal(hit.prim, meta, hit.p, hit.flags, rayNd, normal, areas);
        const int matId = PM_MATERIAL(meta);
        const b200_Material& mat = cS.mats[matId];
        s.currentMaterialId = matId;
        const float4 rrto = *reinterpret_cast<const float4*>(&mat.reflection); // reflection, refraction, transparency, opacity
        attributes = f4(rrto.x, rrto.z, rrto.y, rrto.w);
        matInnerX = mat.innerIllumination.x;
        matColorW = mat.color.w;
        if (pass == 0)
        {
            C.setColor(0, f4(0.f, 0.f, 0.f, 0.f));
            C.setContribution(0, 1.f);
            s.latestIntersection = hit.p;
            s.depthOfField = length(hit.p - rayO);
            if (matInnerX == 0.f && (cSI.advancedIllumination == B200_AI_BASIC || cSI.advancedIllumination == B200_AI_FULL))
            {
                const int t = (index + cSI.pathTracingIteration * 100 + cSI.timestamp) % (cS.randomTableSize - 3);
                s.giO = hit.p + normal * cSI.rayEpsilon;
                s.giT.x = normal.x + 100.f * rnd(t);
                s.giT.y = normal.y + 100.f * rnd(t + 1);
                s.giT.z = normal.z + 100.f * rnd(t + 2);
                const float cos_theta = dot(normalize(s.giT), normal);
                if (cos_theta < 0.f) s.giT = -s.giT;
                s.giT += hit.p;
                s.pathTracingRatio = (1.f - attributes.y) * fabsf(cos_theta);
                s.useGlobalIllumination = true;
            }
            s.idx = __ldg(&cS.prims[hit.prim].index);
        }
        s.rBlinn.w = attributes.y;
    }
    const float4 shaded = primitiveShader(found, index, s.curO, normal, hit.prim, hit.p, areas, s.closestColor, pass, s.shadowIntensity,
                                          s.rBlinn, attributes, (packetMask & (pass == 0 ? 4 : 8)) != 0, cnt);
    if (found)
    {
        const float3 closestIntersection = hit.p;
        float4 colorOfPass = shaded;
        float contribution;
        float3 reflectedTarget = f3(0.f, 0.f, 0.f);
        s.idz += matInnerX * 256;
        const float segmentLength = length(closestIntersection - s.latestIntersection);
        s.latestIntersection = closestIntersection;
        const float transparency = attributes.y;
        float a = 0.f;
        if (attributes.y != 0.f)
        {
            float refraction = attributes.z;
            if (s.initialRefraction == refraction)
            {
                refraction = 1.f;
                const float len = segmentLength * (attributes.w * (1.f - transparency));
                s.rayLength += len;
                s.rayLength = (s.rayLength > cSI.viewDistance) ? cSI.viewDistance : s.rayLength;
                a = (s.rayLength / cSI.viewDistance);
                colorOfPass.x -= a; colorOfPass.y -= a; colorOfPass.z -= a;
            }
            const float3 O_E = normalize(closestIntersection - s.curO);
            vectorRefraction(reflectedTarget, O_E, refraction, normal, s.initialRefraction);
            contribution = transparency - a;
            s.initialRefraction = refraction;
            if (s.reflectedRays == -1 && attributes.x != 0.f)
            {
                float3 rd;
                vectorReflection(rd, O_E, normal);
                s.reflO = closestIntersection + rd * cSI.rayEpsilon;
                s.reflT = closestIntersection + rd;
                s.reflectedRatio = attributes.x;
                s.reflectedRays = pass;
            }
        }
        else if (attributes.x != 0.f)
        {
            const float3 O_E = normalize(closestIntersection - s.curO);
            vectorReflection(reflectedTarget, O_E, normal);
            contribution = attributes.x;
        }
        else
        {
            s.carryon = false;
            contribution = 1.f;
        }
        C.setColor(pass, colorOfPass);
        C.setContribution(pass, contribution);
        s.rBlinn /= (float)(pass + 1);
        s.recursiveBlinn.x = (s.rBlinn.x > s.recursiveBlinn.x) ? s.rBlinn.x : s.recursiveBlinn.x;
        s.recursiveBlinn.y = (s.rBlinn.y > s.recursiveBlinn.y) ? s.rBlinn.y : s.recursiveBlinn.y;
        s.recursiveBlinn.z = (s.rBlinn.z > s.recursiveBlinn.z) ? s.rBlinn.z : s.recursiveBlinn.z;
        s.curO = closestIntersection + reflectedTarget * cSI.rayEpsilon;
        s.curT = closestIntersection + reflectedTarget;
        if (cSI.pathTracingIteration != 0 && matColorW != 0.f)
        {
            float ratio = matColorW;
            ratio *= (attributes.y == 0.f) ? 1000.f : 1.f;
            const int rindex = (index + cSI.timestamp) % (cS.randomTableSize - 3);
            s.curT.x += rnd(rindex) * ratio;
            s.curT.y += rnd(rindex + 1) * ratio;
            s.curT.z += rnd(rindex + 2) * ratio;
        }
    }
    else if (act)
    {
        float4 c;
        if (cSI.skyboxMaterialId != B200_MATERIAL_NONE)
        {
            c = skyboxMapping(s.curO, s.curT);
            const float rad = c.x + c.y + c.z;
            s.idz += (rad > 2.5f) ? rad * 256.f : 0.f;
        }
        else if (cSI.gradientBackground)
        {
            const float3 up = f3(0.f, 1.f, 0.f);
            const float3 dir = normalize(s.curT - s.curO);
            float angle = 0.5f - dot(up, dir);
            angle = (angle > 1.f) ? 1.f : angle;
            c = (1.f - angle) * f4(cSI.backgroundColor.x, cSI.backgroundColor.y, cSI.backgroundColor.z, cSI.backgroundColor.w);
        }
        else
            c = f4(cSI.backgroundColor.x, cSI.backgroundColor.y, cSI.backgroundColor.z, cSI.backgroundColor.w);
        C.setColor(pass, c);
        C.setContribution(pass, 1.f);
    }
    if (act) s.iteration = pass + 1;
}

// extra reflected ray of the first transparent + reflective hit (:296-315)
template <class Colors>
SB_DEV void pathReflectedRay(PathState& s, Colors& C, const bool want, const int index, const int packetMask, Counters& cnt)
{
    const bool debugBoxes = cSI.renderBoxes != 0;
    float3 areas = f3(0.f, 0.f, 0.f);
    float3 normal = f3(0.f, 0.f, 0.f);
    float3 rayNd = f3(0.f, 0.f, 0.f);
    Hit hit;
    hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
    bool found = false;
    if (debugBoxes)
    {
        if (want) { cnt.rays++; boxDebugWalk(s.reflO, s.reflT, s.reflectedRays, s.colorBox); }
    }
    else
    {
        hit = traceClosest(s.reflO, s.reflT, s.reflectedRays, s.currentMaterialId, want, (packetMask & 2) != 0, rayNd, cnt);
        found = want && hit.prim >= 0;
    }
    float4 attributes = f4(0.f, 0.f, 0.f, 0.f);
    if (found)
    {
        const int meta = __ldg(cS.meta + hit.prim);
        hitNormal(hit.prim, meta, hit.p, hit.flags, rayNd, normal, areas);
        attributes.x = cS.mats[PM_MATERIAL(meta)].reflection;
    }
    const float4 color = primitiveShader(found, index, s.reflO, normal, hit.prim, hit.p, areas, s.closestColor, s.reflectedRays,
                                         s.shadowIntensity, s.rBlinn, attributes, (packetMask & 8) != 0, cnt);
    if (found)
    {
        C.setColor(s.reflectedRays, C.color(s.reflectedRays) + color * s.reflectedRatio);
        s.idw = s.shadowIntensity * 255;
    }
}

// back-to-front fold (:382-385), Blinn, fog (:391-400), box-debug tint
template <class Colors>
SB_DEV float4 pathFinish(const PathState& s, const Colors& C, const bool folded)
{
    float4 intersectionColor;
    if (folded)
    {
        // colors[i] = colors[i] (1 - k[i]) + colors[i + 1] k[i] for i = iteration - 2 .. 0, carried in a register
        float4 acc = (s.iteration > 0) ? C.color(s.iteration - 1) : f4(0.f, 0.f, 0.f, 0.f);
        for (int i = s.iteration - 2; i >= 0; --i)
        {
            const float k = C.contribution(i);
            acc = C.color(i) * (1.f - k) + acc * k;
        }
        intersectionColor = acc;
        intersectionColor += s.recursiveBlinn;
    }
    else
        intersectionColor = C.color(0);
    const float D1 = cSI.viewDistance * 0.95f;
    if (cSI.atmosphericEffect == B200_AE_FOG && s.depthOfField > D1)
    {
        const float D2 = cSI.viewDistance * 0.05f;
        const float a = s.depthOfField - D1;
        const float b = 1.f - (a / D2);
        intersectionColor = intersectionColor * b + f4(cSI.backgroundColor.x, cSI.backgroundColor.y, cSI.backgroundColor.z, cSI.backgroundColor.w) * (1.f - b);
    }
    intersectionColor -= s.colorBox;
    return intersectionColor;
}

// The whole ray tree of one pixel in one thread, warp-uniform: `valid` says whether this lane owns a pixel, and a lane
// whose ray tree has ended simply stops taking part (act == false) while the warp finishes.
SB_DEV float4 launchRayTracing(const bool valid, const int index, const float3 rayO, const float3 rayT, float& depthOfField, int4& id,
                               Counters& cnt)
{
    PathState s;
    pathInit(s, rayO, rayT);
    LocalColors C;
    // The reference zero-fills both arrays (:94-95); every entry read later (the fold over 0..iteration-1, the reflected-ray and
    // GI updates of entries already produced) is written by the pass that produced it, so only entry 0 needs a defined
    // value for lanes that own no pixel.
    C.c[0] = f4(0.f, 0.f, 0.f, 0.f); C.k[0] = 0.f;
    const int currentMaxIteration = pathMaxIteration();
    const bool debugBoxes = cSI.renderBoxes != 0;
    const int packetMask = cP.packetMask;

    for (int pass = 0; pass < currentMaxIteration; ++pass)
    {
        // a lane takes part in pass p iff its own loop condition (:125) holds; then iteration == p
        const bool act = valid && s.carryon && s.rayLength < cSI.viewDistance;
        if (!__any_sync(FULL_MASK, act)) break;
        pathPass(s, C, pass, act, index, rayO, packetMask, cnt);
    }

    {
        const bool want = valid && cSI.graphicsLevel >= B200_GL_REFLECTIONS && s.reflectedRays != -1;
        if (__any_sync(FULL_MASK, want)) pathReflectedRay(s, C, want, index, packetMask, cnt);
    }

    bool test = true;
    if ((cSI.advancedIllumination == B200_AI_BASIC || cSI.advancedIllumination == B200_AI_FULL) &&
        cSI.pathTracingIteration >= B200_NB_MAX_ITERATIONS)
    {
        // global-illumination ray (:317-378)
        const bool giFull = cSI.advancedIllumination == B200_AI_FULL;
        const bool want = valid && s.useGlobalIllumination && giFull;
        float3 areas = f3(0.f, 0.f, 0.f);
        float3 normal = f3(0.f, 0.f, 0.f);
        float3 rayNd = f3(0.f, 0.f, 0.f);
        Hit hit;
        hit.prim = -1; hit.p = f3(0.f, 0.f, 0.f); hit.flags = 0;
        bool giHit = false;
        float4 pathTracingColor = f4(0.f, 0.f, 0.f, 0.f);
        if (giFull)
        {
            if (debugBoxes)
            {
                if (want) { cnt.rays++; boxDebugWalk(s.giO, s.giT, 30, s.colorBox); }
            }
            else
            {
                hit = traceClosest(s.giO, s.giT, 30, B200_MATERIAL_NONE, want, (packetMask & 2) != 0, rayNd, cnt);
                giHit = want && hit.prim >= 0;
            }
        }
        bool shadeGi = false;
        float4 attributes = f4(0.f, 0.f, 0.f, 0.f);
        if (giHit)
        {
            const int meta = __ldg(cS.meta + hit.prim);
            hitNormal(hit.prim, meta, hit.p, hit.flags, rayNd, normal, areas);
            const b200_Material& material = cS.mats[PM_MATERIAL(meta)];
            const float4 mc = f4(material.color.x, material.color.y, material.color.z, material.color.w);
            if (cS.prims[hit.prim].materialId != B200_MATERIAL_NONE)
            {
                if (material.innerIllumination.x == 0.f)
                {
                    C.c[0] = mc * material.innerIllumination.x * s.pathTracingRatio;
                    test = false;
                }
                else
                    C.c[0] = mc * s.pathTracingRatio;
            }
            if (test)
            {
                s.pathTracingRatio *= 0.1f; // STANDARD_LUNINANCE_STRENGTH (Consts.h:52)
                if (material.innerIllumination.x == 0.f)
                    C.c[0] -= cSI.shadowIntensity;
                else
                    shadeGi = true;
            }
        }
        if (giFull)
        {
            const float4 c = primitiveShader(shadeGi, index, s.giO, normal, hit.prim, hit.p, areas, s.closestColor, s.iteration, s.shadowIntensity,
                                             s.rBlinn, attributes, (packetMask & 8) != 0, cnt);
            if (shadeGi) pathTracingColor = c;
        }
        if (valid && !giHit && cSI.skyboxMaterialId != B200_MATERIAL_NONE)
        {
            // no GI hit, or aiBasic (where the reference maps an uninitialised ray, :369-375; zero here)
            pathTracingColor = skyboxMapping(s.giO, s.giT);
            s.pathTracingRatio *= 0.2f; // SKYBOX_LUNINANCE_STRENGTH (Consts.h:53)
        }
        if (test) C.c[0] += pathTracingColor * s.pathTracingRatio;
    }

    const float4 intersectionColor = pathFinish(s, C, test);
    depthOfField = s.depthOfField;
    id.x = s.idx; id.y = s.iteration; id.z = s.idz; id.w = s.idw;
    return intersectionColor;
}
