// ============================================================================
// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// "CUDA on the host" shim: lets g++ compile the reference's UNMODIFIED VX3 device sources
// (/root/reference/src/VX3/*.cu, src/Utils/*.h, *.cuh) as plain C++, so that the reference's own
// VX3_Link::updateForces, VX3_Voxel::timeStep, VX3_Collision, VX3_MathTree::eval,
// VX3_VoxelyzeKernel::doTimeStep ... EXECUTE in this container (no GPU, and the device path does not
// build with nvcc >= 12 because of device-side cudaDeviceSynchronize).  oracle/Makefile force-includes
// this header (-include) and puts oracle/ref_vx3/stubs first on the include path; nothing of the
// reference is copied — the sources are compiled from where they lie, outputs go to oracle/_ref/.
//
// What the shim supplies:
//   * __device__/__host__/__global__ as empty attributes; threadIdx/blockIdx/blockDim/gridDim globals;
//   * kernel launches: the Makefile pipes the ONE file that has launches (VX3_VoxelyzeKernel.cu) through
//     sed, which rewrites `kernel<<<grid, block>>>(args)` to `vxhost_launch(kernel, grid, block, args)`
//     on the fly (stdin of g++; nothing is written to disk).  vxhost_launch runs every thread of the
//     grid SEQUENTIALLY, x-coordinate outermost — a legal CUDA schedule, and the canonical order
//     (first, second ascending) that SURVEY.md Appendix A.7 / the oracle define for the racy sweeps;
//   * cudaMalloc/cudaMemcpy/cudaFree on host memory ("device" objects are host objects), the atomics the
//     device containers use, cudaOccupancyMaxPotentialBlockSize, cuRAND (stubs/curand.h);
//   * printf capture (the history frames and "real_stepsize" line are device printf in the reference).
// ============================================================================
#pragma once
#include <cassert>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>

#define __device__
#define __host__
#define __global__
#define __inline__ inline
#define __forceinline__ inline

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline dim3 threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaLimit { cudaLimitMallocHeapSize, cudaLimitPrintfFifoSize, cudaLimitStackSize };
inline const char *cudaGetErrorString(cudaError_t) { return "host shim"; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)calloc(1, n ? n : 1); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaFree(void *) { return cudaSuccess; } // the reference leaks by design (VX3_MemoryCleaner.h); keep objects alive
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)1 << 34; return cudaSuccess; }
inline cudaError_t cudaDeviceGetLimit(size_t *v, cudaLimit) { *v = 1 << 20; return cudaSuccess; }
inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
template <class K> inline cudaError_t cudaOccupancyMaxPotentialBlockSize(int *minGrid, int *block, K, size_t = 0, int = 0) {
    *minGrid = 1;
    *block = 256;
    return cudaSuccess;
}

// CUDA math-library function the host libm lacks (VX3_MathTree.h:116)
inline double normcdf(double x) { return 0.5 * erfc(-x * M_SQRT1_2); }

inline int atomicCAS(int *p, int cmp, int val) { int old = *p; if (old == cmp) *p = val; return old; }
inline int atomicExch(int *p, int val) { int old = *p; *p = val; return old; }
inline int atomicAdd(int *p, int v) { int old = *p; *p += v; return old; }
inline double atomicAdd(double *p, double v) { double old = *p; *p += v; return old; }

// sequential execution of a grid: global x outermost, then y, then z (see header)
template <class... P, class... A> inline void vxhost_launch(void (*kernel)(P...), dim3 grid, dim3 block, A &&...args) {
    const dim3 sT = threadIdx, sB = blockIdx, sD = blockDim, sG = gridDim; // launches nest (kernels launching kernels)
    for (unsigned gx = 0; gx < grid.x * block.x; gx++)
        for (unsigned gy = 0; gy < grid.y * block.y; gy++)
            for (unsigned gz = 0; gz < grid.z * block.z; gz++) {
                blockDim = block;
                gridDim = grid;
                blockIdx = dim3(gx / block.x, gy / block.y, gz / block.z);
                threadIdx = dim3(gx % block.x, gy % block.y, gz % block.z);
                kernel(args...);
            }
    threadIdx = sT; blockIdx = sB; blockDim = sD; gridDim = sG;
}

// device printf -> capture buffer (vx3ref_take_output in the harness); dropped when capture is off
inline std::string &vxhost_out() { static std::string s; return s; }
inline bool &vxhost_capture() { static bool on = false; return on; }
inline int vxhost_printf(const char *fmt, ...) {
    if (!vxhost_capture()) return 0;
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (n > 0) vxhost_out().append(buf, (size_t)(n < (int)sizeof(buf) ? n : (int)sizeof(buf) - 1));
    return n;
}
#define printf vxhost_printf
