// Halo exchange for ONE body decomposed into slabs over the GPUs of a box (BASELINE config 5; not in the reference,
// whose only multi-GPU mode is file i -> device i % nDevices, src/Executables/vx3_node_worker.cu:88-93).
//
// Every rank holds the voxels of its slab plus GHOST copies (VX3_VOX_GHOST) of the neighbour slabs' face voxels, and every
// link with at least one owned end; a link that crosses a face is evaluated on both sides from identical inputs, so it
// needs no message of its own.  The only per-step traffic is the 64-byte pose record (position,
// orientation, next-step temperature, previousDt) of each face voxel, and it moves inside the step kernels' stream with no
// host involvement and no NCCL call:
//   k_halo_send  writes my face poses straight into the neighbour's receive buffer (peer memory opened through CUDA IPC,
//                i.e. plain stores over NVLink), then publishes the step number with a system-scope release;
//   k_halo_recv  spins (acquire, system scope) on the step number the neighbours published into MY flags, then moves the
//                received records into my ghost voxels' pose records.
// Receive buffers are double-buffered by step parity: a neighbour can be at most one step ahead of me (its step s+2 send
// needs my step s+1 send, which follows my step s receive), so the buffer it overwrites is never the one I still read.
//
// OVERLAP.  The exchange of step s runs on a second stream (Halo::stream2) behind an event recorded after the step's voxel
// pass, while the main stream already evaluates step s+1's INTERIOR links — the links with two owned ends, which read no ghost
// pose; the host-side partition stores them first (parallel.partition_slabs), so they are one tile range.  The main stream
// waits for the exchange (Halo::ev_halo) only before the FACE-link range.  Send + wait + scatter are hidden behind the largest
// kernel of the step.
#pragma once
#include <cuda_runtime.h>

#include "vx3_kernels.cuh"

namespace vx3 {

#define VX3_HALO_BLOCK 256
// a dead neighbour must not hang the GPU for ever: the wait gives up after this many milliseconds (VX3_HALO_TIMEOUT_MS, default
// 60 s — a rank can be held up for seconds by paging, JIT or a profiler), marks the batch failed and freezes it (dt = 0: every
// later step kernel of the stream is a no-op); the host sees VX3_ERR_CUDA at its next step / sync call
#define VX3_HALO_TIMEOUT_MS_DEFAULT 60000

struct HaloSide {                 // one neighbour
    int n_send = 0, n_recv = 0;
    int32_t *send_idx = nullptr;  // my face voxels (global voxel indices of this batch), in the order both sides agree on
    int32_t *recv_idx = nullptr;  // my ghost voxels fed by this neighbour, same order as its send list
    double *recv_buf = nullptr;   // MINE: [2 parities][n_recv][8] doubles, written by the neighbour
    unsigned int *recv_flag = nullptr; // MINE: [2 parities] step numbers (+1), written by the neighbour; 128-B apart
    double *peer_buf = nullptr;   // the neighbour's recv_buf for my side (IPC-mapped)
    unsigned int *peer_flag = nullptr;
    unsigned int *send_count = nullptr; // arrival counter of k_halo_send's CTAs
    bool peer_open = false;  // peer block mapped through CUDA IPC (another process)
    bool peer_local = false; // peer block belongs to a batch of this process
};

struct Halo {
    bool on = false;
    HaloSide side[2]; // 0 = lower neighbour, 1 = upper neighbour
    int *err = nullptr; // device flag: spin limit hit
    cudaStream_t stream2 = nullptr;               // the exchange runs here
    cudaEvent_t ev_step = nullptr, ev_halo = nullptr; // voxel pass of step s done / ghost poses of step s in place
    bool pending = false;                         // an exchange is in flight that the main stream has not waited for yet
    int face_tile0 = -1;                          // link tiles [0, face_tile0) hold no link with a ghost end (-1: not analysed yet)
    long long spin_cycles = 0;
};

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// my face poses -> the neighbour's receive buffer (peer stores), then the step number
__global__ void __launch_bounds__(VX3_HALO_BLOCK) k_halo_send(const double *__restrict__ pose, const int32_t *__restrict__ idx, int n, double *peer_buf,
                                                               unsigned int *peer_flag, unsigned int *count, unsigned int step1, int parity) {
    double2 *dst = reinterpret_cast<double2 *>(peer_buf + (size_t)parity * n * 8);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * n; i += gridDim.x * blockDim.x) { // one 16-byte quarter per thread: coalesced peer stores
        const int v = idx[i >> 2];
        dst[i] = reinterpret_cast<const double2 *>(pose + 8 * (size_t)v)[i & 3];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int arrived = atomicAdd(count, 1u);
        if (arrived == gridDim.x - 1) { // last CTA: every CTA's stores are fenced
            *count = 0u;
            st_release_sys(peer_flag + 32 * parity, step1);
        }
    }
}

// wait for the neighbour's step number, then ghost poses <- receive buffer
__global__ void __launch_bounds__(VX3_HALO_BLOCK) k_halo_recv(double *pose, const int32_t *__restrict__ idx, int n, const double *recv_buf,
                                                               const unsigned int *recv_flag, unsigned int step1, int parity, int *err, long long spin_cycles,
                                                               SimD *simd) {
    __shared__ int ok;
    if (threadIdx.x == 0) {
        ok = *reinterpret_cast<volatile int *>(err) == 0; // sticky: after one failed wait nothing is scattered any more
        const long long t0 = clock64();
        while (ok && ld_acquire_sys(recv_flag + 32 * parity) < step1) {
            if (clock64() - t0 > spin_cycles) {
                ok = 0;
                *err = 1;
                simd->err = VX3_ERR_CUDA;
                simd->dt = 0.0f; // freeze: doTimeStep(0) does nothing (VX3_VoxelyzeKernel.cu:240-241)
                __threadfence();
            }
        }
    }
    __syncthreads();
    if (!ok) return;
    const double2 *src = reinterpret_cast<const double2 *>(recv_buf + (size_t)parity * n * 8);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * n; i += gridDim.x * blockDim.x) {
        const int v = idx[i >> 2];
        reinterpret_cast<double2 *>(pose + 8 * (size_t)v)[i & 3] = __ldcv(src + i); // written by another device: never from a stale L1 line
    }
}

} // namespace vx3
