// vec.cuh — float3/float4 arithmetic for the engine's device code.
// Component-wise, left-to-right, so expression shapes (and therefore nvcc's FMA contraction and the
// --use_fast_math lowering of / sqrtf rsqrtf) match what the reference engine computes with its vector
// helpers (/root/reference/solr/engines/cuda/helper_math.h:1248-1318: dot, length, normalize = v*rsqrtf(dot)).
#pragma once
#include <cuda_runtime.h>

#define SB_DEV __device__ __forceinline__

SB_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
SB_DEV float4 f4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
SB_DEV float3 xyz(const float4& a) { return make_float3(a.x, a.y, a.z); }
SB_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
SB_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
SB_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
SB_DEV float3 operator*(float3 a, float b) { return f3(a.x * b, a.y * b, a.z * b); }
SB_DEV float3 operator*(float b, float3 a) { return f3(b * a.x, b * a.y, b * a.z); }
SB_DEV float3 operator/(float3 a, float b) { return f3(a.x / b, a.y / b, a.z / b); }
SB_DEV void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
SB_DEV void operator*=(float3& a, float b) { a.x *= b; a.y *= b; a.z *= b; }
SB_DEV float4 operator+(float4 a, float4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
SB_DEV float4 operator-(float4 a, float4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
SB_DEV float4 operator*(float4 a, float4 b) { return f4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
SB_DEV float4 operator*(float4 a, float b) { return f4(a.x * b, a.y * b, a.z * b, a.w * b); }
SB_DEV float4 operator*(float b, float4 a) { return f4(b * a.x, b * a.y, b * a.z, b * a.w); }
SB_DEV float4 operator/(float4 a, float b) { return f4(a.x / b, a.y / b, a.z / b, a.w / b); }
SB_DEV void operator+=(float4& a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
SB_DEV void operator-=(float4& a, float4 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w; }
SB_DEV void operator-=(float4& a, float b) { a.x -= b; a.y -= b; a.z -= b; a.w -= b; }
SB_DEV void operator*=(float4& a, float b) { a.x *= b; a.y *= b; a.z *= b; a.w *= b; }
SB_DEV void operator/=(float4& a, float b) { a.x /= b; a.y /= b; a.z /= b; a.w /= b; }
// Pinned rounding.  nvcc contracts a*b + c*d as fma(a, b, c*d) — first product fused, second rounded on its own — unless
// something else also uses a product, in which case it rounds that product first: in this engine's large kernels the same
// dot product came out one ulp apart in two walks (the sphere test's a = 2 dot(dir, dir) when dir.x*dir.x was also wanted
// by the ellipsoid test), and the discriminant b*b - 2ac magnifies one ulp of `a` into a different answer for grazing rays.
// The reference's own build has the plain contraction everywhere (checked against its SASS and, pixel for pixel, against
// its frames), so the dot and cross products are spelled out with it.
SB_DEV float dot(float3 a, float3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y))); }
SB_DEV float length(float3 v) { return sqrtf(dot(v, v)); }
// dot(v, v) inside normalize: the reference's build rounds x*x on its own and fuses y, z (its SASS, e.g. the direction of every
// ray in sphereIntersection) — the other way round from the dot products above
SB_DEV float dotSelfForNormalize(float3 v) { return __fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, __fmul_rn(v.x, v.x))); }
SB_DEV float3 normalize(float3 v) { float invLen = rsqrtf(dotSelfForNormalize(v)); return v * invLen; }
SB_DEV float3 cross(float3 b, float3 c)
{
    return f3(__fmaf_rn(b.y, c.z, -__fmul_rn(b.z, c.y)), __fmaf_rn(b.z, c.x, -__fmul_rn(b.x, c.z)), __fmaf_rn(b.x, c.y, -__fmul_rn(b.y, c.x)));
}
SB_DEV void saturate4(float4& v)
{
    v.x = (v.x < 0.f) ? 0.f : v.x; v.y = (v.y < 0.f) ? 0.f : v.y; v.z = (v.z < 0.f) ? 0.f : v.z; v.w = (v.w < 0.f) ? 0.f : v.w;
    v.x = (v.x > 1.f) ? 1.f : v.x; v.y = (v.y > 1.f) ? 1.f : v.y; v.z = (v.z > 1.f) ? 1.f : v.z; v.w = (v.w > 1.f) ? 1.f : v.w;
}

// ---------------------------------------------------------------------------------------------------
// Accesses to data that is written once and read once (paths parked between passes, their colours, queue entries, the frame's
// buffers): no line in L1 (another SM may have written it during the same launch: fused stages) and evict-first in L2, so that a
// frame's hundreds of megabytes of them leave the L2 to the walk trees and the thread-local memory.  STREAM_HINTS: 0 plain
// (ld.cg / st), 1 parked paths, colours and queues, 2 the frame's buffers as well.
// ---------------------------------------------------------------------------------------------------
#ifndef STREAM_HINTS
#define STREAM_HINTS 1 // measured on config 2: 4.57 -> 4.50 ms (level 2: the same)
#endif
SB_DEV unsigned long long evictFirstPolicy()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
SB_DEV float ldOnce(const float* p)
{
#if STREAM_HINTS
    float v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(evictFirstPolicy()));
    return v;
#else
    return __ldcg(p);
#endif
}
SB_DEV float4 ldOnce(const float4* p)
{
#if STREAM_HINTS
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(evictFirstPolicy()));
    return v;
#else
    return __ldcg(p);
#endif
}
SB_DEV void stOnce(float* p, const float v)
{
#if STREAM_HINTS
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(evictFirstPolicy()) : "memory");
#else
    *p = v;
#endif
}
SB_DEV void stOnce(float4* p, const float4 v)
{
#if STREAM_HINTS
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(evictFirstPolicy()) : "memory");
#else
    *p = v;
#endif
}
