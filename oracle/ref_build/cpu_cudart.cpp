// oracle/ref_build/cpu_cudart.cpp — TEST INFRASTRUCTURE, not product code.
// Host-memory stand-ins for the handful of CUDA runtime entry points the reference's engine calls
// (CudaRayTracer.cu:1360-1672, CudaKernel.cpp:391-537), so the CPU build of the reference links
// without libcudart and without a GPU.  "Device" memory is host memory.
#include "cpu_shim.h"
#include <cstdio>

thread_local SolrCpuIdx threadIdx;
thread_local SolrCpuIdx blockIdx;
thread_local SolrCpuIdx blockDim;

extern "C" {
cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* c) { *c = 1; return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = 0; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceReset() { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "cpu-shim"; }
cudaError_t cudaDriverGetVersion(int* v) { *v = 0; return cudaSuccess; }
cudaError_t cudaRuntimeGetVersion(int* v) { *v = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) { memset(p, 0, sizeof(*p)); snprintf(p->name, sizeof(p->name), "host CPU (reference kernels via cpu_shim)"); return cudaSuccess; }
}
