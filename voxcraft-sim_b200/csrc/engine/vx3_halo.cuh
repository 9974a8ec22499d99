// Halo exchange for ONE body decomposed into slabs over the GPUs of a box (BASELINE config 5; not in the reference,
// whose only multi-GPU mode is file i -> device i % nDevices, src/Executables/vx3_node_worker.cu:88-93).
//
// Every rank holds the voxels of its slab plus GHOST copies (VX3_VOX_GHOST) of the neighbour slabs' face voxels, and every
// link with at least one owned end; a link that crosses a face is evaluated on both sides from identical inputs, so it
// needs no message of its own.  The only per-step traffic is the 64-byte pose record (position,
// orientation, next-step temperature, previousDt) of each face voxel, and it moves inside the step kernels themselves, with no
// host involvement, no NCCL call and (on the default path) no kernel of its own:
//   SEND     k_voxels<.., HALO> stores a face voxel's new record into its own pose row AND into the neighbour's receive buffer
//            (peer memory opened through CUDA IPC: plain stores over NVLink).  Sending threads fence at system scope before
//            their CTA counts itself in; the last CTA of the pass to finish publishes the send number to both neighbours
//            and does the step's bookkeeping (tail_light) in a second warp.  (HaloOut, vx3_device.cuh.)
//   RECEIVE  k_links_deferred<.., HALO> runs the interior link tiles first (parallel.partition_slabs stores the links with two
//            owned ends first).  Before its first face tile a warp makes sure the neighbours' send number is in — one warp of
//            the grid polls the system-scope flags, the others a flag in local memory — and the face links then read their
//            ghost ends straight from the receive buffers (HaloIn::ghost_row): no copy into the ghost voxels' pose rows, and
//            the transfer has the whole interior range (97 % of the pass) to arrive.
//   The ghost rows of the pose array are brought up to date only when a stepping call returns (read-backs): k_halo_wait +
//   k_halo_recv below, once per call.
// Receive buffers are double-buffered by step parity: a neighbour can be at most one step ahead of me (its step s+2 send
// needs my step s+1 send, which follows my step-s+1 link pass, the last reader of its step-s records), so the buffer it overwrites
// is never the one I still read.  The send number lives in device memory (Halo::seq), not in a kernel argument, so a stretch of
// steps replays as a CUDA Graph like any other batch.
//
// MEASURED on the 4M-voxel body, 8 GPUs, same box back to back (us per step): stand-alone k_halo_send / k_halo_wait / k_halo_recv
// kernels in the step stream 215.3; receive inside the link pass 205.3; plus send inside the voxel pass 196.1 (2.04e10 voxel-steps/s,
// 6.6x one GPU).  At 2 GPUs the three differ by < 1 % (600 us steps hide a 20 us exchange either way).  The stand-alone kernels
// remain as the fallback: two slabs driven by one process (they could hold each other's CTA slots while they wait), slabs one
// voxel thick, batches with voxel removal, VX3_HALO_INKERNEL=0 / VX3_HALO_SENDFUSED=0.
// Earlier experiments, all slower (4 / 8 GPUs): the exchange on a second stream under the next link pass 462 / - against 382 / 213
// (the link pass is a persistent tile loop that fills every CTA slot, the other stream's kernels do not get on the SMs); the link pass
// as two launches around the receive 359 / 216 against 353 / 211; hundreds of CTAs polling a system-scope flag: 5x slower steps; a
// collect step inside the link pass (every warp copies a slice, face tiles wait for all): no gain, it is a grid-wide barrier.
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
    bool pending = false;                         // my poses are out, the neighbours' are not collected yet (host-side hint only)
    // device-side sequence numbers {sent, collected, send arrivals, receive arrivals}: the exchange kernels take the step number from here, not from
    // a kernel argument, so that a stretch of steps can be captured once in a CUDA Graph and replayed
    unsigned int *seq = nullptr;
    // the link pass reads the neighbours' poses from the receive buffers itself (HaloIn, vx3_device.cuh) unless a neighbour is a batch of this process
    // (two slabs sharing one device could hold each other's CTA slots while they wait) or VX3_HALO_INKERNEL=0
    bool inkernel = false;
    HaloIn *hin = nullptr;
    unsigned int *state = nullptr;
    int32_t *ghost_row = nullptr;
    // ... and the voxel pass sends: a face voxel's record goes to the neighbour as it is computed (HaloOut); needs disjoint face lists
    bool send_fused = false;
    HaloOut *hout = nullptr;
    unsigned int *out_count = nullptr;
    int send_blocks = 0;                          // CTAs of k_halo_send (0: this slab sends nothing)
    int face_tile0 = -1;                          // link tiles [0, face_tile0) hold no link with a ghost end (-1: not analysed yet)
    long long spin_cycles = 0;
};

// Both neighbours in one launch: the CTAs [0, nb[0]) serve side 0, the rest side 1; one 16-byte quarter of a pose record per
// thread and ONE quarter per thread (the grid covers the face), so a kernel is a single round of independent accesses.
struct HaloSendArgs {
    const int32_t *idx[2];
    double *peer_buf[2];
    unsigned int *peer_flag[2], *count[2];
    int n[2], nb[2];
};
struct HaloRecvArgs {
    const int32_t *idx[2];
    const double *recv_buf[2];
    const unsigned int *recv_flag[2];
    int n[2], nb[2];
};

// my face poses -> the neighbours' receive buffers (peer stores), then the step number.  tail >= 0: the last CTA also does the
// end-of-step bookkeeping of the (single) simulation, k_tail_light's work with check_stop = tail — one launch less per step
__global__ void __launch_bounds__(VX3_HALO_BLOCK) k_halo_send(Dev D, HaloSendArgs a, unsigned int *seq, int tail) {
    const double *__restrict__ pose = D.pose;
    if (tail >= 0 && blockIdx.x == gridDim.x - 1 && threadIdx.x == VX3_HALO_BLOCK - 1) tail_light(D, 0, tail);
    // this is send number seq[0] + 1 (every CTA reads it before the last one to arrive bumps it); neighbours run in lock step, so
    // both sides of a face count the same sends
    const unsigned int step1 = *reinterpret_cast<volatile unsigned int *>(seq) + 1u;
    const int parity = (int)((step1 - 1u) & 1u);
    const int sd = (int)blockIdx.x < a.nb[0] ? 0 : 1;
    const int blk = sd ? (int)blockIdx.x - a.nb[0] : (int)blockIdx.x;
    const int n = a.n[sd];
    double2 *dst = reinterpret_cast<double2 *>(a.peer_buf[sd] + (size_t)parity * n * 8);
    for (int i = blk * blockDim.x + threadIdx.x; i < 4 * n; i += a.nb[sd] * blockDim.x) { // coalesced peer stores
        const int v = a.idx[sd][i >> 2];
        dst[i] = reinterpret_cast<const double2 *>(pose + 8 * (size_t)v)[i & 3];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int arrived = atomicAdd(a.count[sd], 1u);
        if (arrived == (unsigned int)a.nb[sd] - 1) { // last CTA of this side: every CTA's stores are fenced
            *a.count[sd] = 0u;
            st_release_sys(a.peer_flag[sd] + 32 * parity, step1);
        }
        const unsigned int all = atomicAdd(seq + 2, 1u);
        if (all == (unsigned int)(a.nb[0] + a.nb[1]) - 1) { // last CTA of the launch: every CTA has read seq[0]
            seq[2] = 0u;
            __threadfence();
            seq[0] = step1;
        }
    }
}

// wait for the neighbours' step numbers: ONE thread per side polls the system-scope flag (hundreds of CTAs polling it, as a first
// version of the wide receive kernel did, slow the step down 5x: 1.9 ms instead of 0.36 ms on 4 GPUs)
__global__ void k_halo_wait(HaloRecvArgs a, const unsigned int *seq, int *err, long long spin_cycles, SimD *simd) {
    const int sd = threadIdx.x;
    if (sd > 1 || a.nb[sd] == 0) return;
    const unsigned int step1 = seq[0]; // my sends so far = the neighbours' send I need
    if (step1 == 0 || seq[1] == step1) return; // nothing sent yet / already collected
    const int parity = (int)((step1 - 1u) & 1u);
    if (*reinterpret_cast<volatile int *>(err) != 0) return; // sticky: after one failed wait nothing is scattered any more
    const long long t0 = clock64();
    while (ld_acquire_sys(a.recv_flag[sd] + 32 * parity) < step1) {
        if (clock64() - t0 > spin_cycles) {
            *err = 1;
            simd->err = VX3_ERR_CUDA;
            simd->dt = 0.0f; // freeze: doTimeStep(0) does nothing (VX3_VoxelyzeKernel.cu:240-241)
            __threadfence();
            return;
        }
    }
}

// ghost poses <- receive buffers (after k_halo_wait in the same stream)
__global__ void __launch_bounds__(VX3_HALO_BLOCK) k_halo_recv(double *pose, HaloRecvArgs a, unsigned int *seq, const int *err) {
    if (*err) return;
    const unsigned int step1 = seq[0];
    if (step1 == 0 || *reinterpret_cast<volatile unsigned int *>(seq + 1) == step1) return; // (the same answer in every CTA: seq[1] moves when all have read it)
    const int parity = (int)((step1 - 1u) & 1u);
    const int sd = (int)blockIdx.x < a.nb[0] ? 0 : 1;
    const int blk = sd ? (int)blockIdx.x - a.nb[0] : (int)blockIdx.x;
    const int n = a.n[sd];
    const double2 *src = reinterpret_cast<const double2 *>(a.recv_buf[sd] + (size_t)parity * n * 8);
    for (int i = blk * blockDim.x + threadIdx.x; i < 4 * n; i += a.nb[sd] * blockDim.x) {
        const int v = a.idx[sd][i >> 2];
        reinterpret_cast<double2 *>(pose + 8 * (size_t)v)[i & 3] = __ldcv(src + i); // written by another device: never from a stale L1 line
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int all = atomicAdd(seq + 3, 1u);
        if (all == (unsigned int)(a.nb[0] + a.nb[1]) - 1) { // last CTA: every CTA of the launch has read seq[1]
            seq[3] = 0u;
            seq[1] = step1;
        }
    }
}

} // namespace vx3
