// oracle/ref_build/cpu_shim.h — TEST INFRASTRUCTURE, not product code.
//
// Lets g++ compile the reference's CUDA translation unit (solr/engines/cuda/CudaRayTracer.cu and
// the .cuh files it includes, read from /root/reference where they lie) as plain host C++, so the
// reference's OWN ray-propagation code runs on CPU cores.  The Makefile force-includes this file and
// rewrites only the `kernel<<<grid, block, shmem, stream>>>(` launch tokens (not valid C++) into
// SOLR_CPU_LAUNCH(grid, block) kernel(  — every other byte of the reference source is compiled as is.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <math.h>
#include <algorithm>
#include <cuda_runtime_api.h> // types only (float3, dim3, cudaStream_t ...); no libcudart is linked
#include <vector_types.h>
#include <vector_functions.h>

// CUDA built-ins a kernel body reads.  One set per OpenMP worker.
struct SolrCpuIdx { unsigned int x, y, z; };
extern thread_local SolrCpuIdx threadIdx;
extern thread_local SolrCpuIdx blockIdx;
extern thread_local SolrCpuIdx blockDim;

static inline bool solr_cpu_set_idx(long blk, long t, const dim3& g, const dim3& b)
{
    blockDim.x = b.x; blockDim.y = b.y; blockDim.z = 1;
    blockIdx.x = (unsigned)(blk % g.x); blockIdx.y = (unsigned)(blk / g.x); blockIdx.z = 0;
    threadIdx.x = (unsigned)(t % b.x); threadIdx.y = (unsigned)(t / b.x); threadIdx.z = 0;
    return true;
}

// A "launch" = every (block, thread) pair runs the kernel body once; blocks are spread over host threads.
#define SOLR_CPU_PRAGMA(x) _Pragma(#x)
#define SOLR_CPU_LAUNCH(G, B)                                                           \
    SOLR_CPU_PRAGMA(omp parallel for schedule(dynamic, 4))                              \
    for (long solr_blk = 0; solr_blk < (long)(G).x * (long)(G).y; ++solr_blk)           \
        for (long solr_t = 0; solr_t < (long)(B).x * (long)(B).y; ++solr_t)             \
            if (solr_cpu_set_idx(solr_blk, solr_t, (G), (B)))

// CUDA's device overload set has max/min(float, float); helper_math.h's host fallbacks only declare the
// int versions, which would silently truncate `max(0.f, min(result, shadowIntensity))`
// (GeometryIntersections.cuh:906) on a host build.  Same semantics as the device builtins (fmaxf/fminf).
static inline float max(float a, float b) { return a > b ? a : b; }
static inline float min(float a, float b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
