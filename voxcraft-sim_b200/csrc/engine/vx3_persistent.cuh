// On-chip persistent step kernel for ONE small collision-free body (BASELINE config 2 class).
//
// The streaming path pays three dependent kernel launches per doTimeStep, which is launch/latency bound for a
// body whose whole state (a few MB) fits on chip.  Here one cooperative grid (one CTA per SM) runs many steps per
// launch: every thread permanently owns at most one link and one voxel whose private state (link history,
// strain, momenta, flags) stays in REGISTERS for the whole launch; only what the other phase needs crosses the
// chip through L2 — the voxel pose (56 B) and the link's end forces/moments (96 B) — in the same global arrays
// the streaming kernels use, so the two paths can alternate freely.  Two grid barriers per step separate the
// link phase from the voxel phase exactly like the reference's two child grids
// (src/VX3/VX3_VoxelyzeKernel.cu:259-269 gpu_update_links, :306-312 gpu_update_voxels); temperature / each voxel's
// temperature for the next step is computed between a barrier's arrive and its wait and travels in its pose record.
// The physics is the same code as the streaming kernels (vx3_physics.cuh).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vx3_kernels.cuh"

namespace vx3 {

#define VX3_PERSIST_MAX_BLOCK 256

struct PersistentPlan {
    bool ok = false;
    int grid = 0, block = 0;
    int links_per_cta = 0, vox_per_cta = 0;
    unsigned int *barrier = nullptr; // [0] arrival counter, [1] divergence flag, [4..] phase cycle counters (debug)
    bool timing = false;
    // point-to-point phase flags (replace the two grid-wide barriers): CTA c waits only for the CTAs that own the
    // voxels its links touch / the links its voxels touch
    bool p2p = false;
    unsigned int *flags = nullptr; // [2][grid][32]: link phases done, voxel phases done (one 128-B line per counter)
    int *deps = nullptr;           // [2][grid][VX3_PERSIST_MAX_DEPS]: producers of the link phase / of the voxel phase
    int *ndeps = nullptr;          // [2][grid]
};

#define VX3_PERSIST_MAX_DEPS 32
struct P2P {
    unsigned int *flags;
    const int *deps, *ndeps;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned int *p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double2 ldcg2(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }

// grid barrier split in two halves so independent work can run while the arrivals propagate
__device__ __forceinline__ void grid_arrive(unsigned int *counter) {
    __syncthreads(); // every thread's exchange stores are issued (and ordered before thread 0's release)
    if (threadIdx.x == 0) red_release_add(counter, 1u);
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Spin with relaxed loads (an acquire load would invalidate L1 on every poll), then ONE acquire fence.
__device__ __forceinline__ void grid_wait(const unsigned int *counter, unsigned int target) {
    if (threadIdx.x == 0) {
        while (ld_relaxed_u32(counter) < target) {}
#ifndef VX3_PERSIST_NO_ACQ_FENCE
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
    }
    __syncthreads();
}

__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// publish "this CTA finished its phase number `value`": every thread's exchange stores precede the release store
__device__ __forceinline__ void p2p_publish(unsigned int *flag, unsigned int value) {
    __syncthreads();
    if (threadIdx.x == 0) st_release_u32(flag, value);
}
// wait until every producer CTA of my next phase has published `target`; returns the divergence flag (uniform)
__device__ __forceinline__ unsigned int p2p_wait(const unsigned int *flags, int my_dep, unsigned int target, const unsigned int *divflag, int *s_div) {
    if (my_dep >= 0) {
        const unsigned int *f = flags + 32 * (size_t)my_dep;
        while (ld_acquire_u32(f) < target) {}
    }
    if (threadIdx.x == 0) *s_div = (int)ld_relaxed_u32(divflag);
    __syncthreads();
    return (unsigned int)*s_div;
}

__global__ void __launch_bounds__(VX3_PERSIST_MAX_BLOCK, 1)
k_persistent(Dev D, long long nsteps, int check_stop, int links_per_cta, int vox_per_cta, unsigned int *bar, int timing, P2P p2p) {
    __shared__ int s_stop;
    __shared__ int s_div;
    __shared__ SimC sS; // per-simulation constants on chip: global loads would miss L1 after every barrier's fence
    for (int i = threadIdx.x; i < (int)(sizeof(SimC) / 4); i += blockDim.x) reinterpret_cast<int *>(&sS)[i] = reinterpret_cast<const int *>(&D.simc[0])[i];
    __syncthreads();
    long long tk[5] = {0, 0, 0, 0, 0}, c0 = 0;
#define TICK(i)                                                                                                         \
    if (timing) {                                                                                                       \
        long long c_ = clock64();                                                                                       \
        tk[i] += c_ - c0;                                                                                               \
        c0 = c_;                                                                                                        \
    }
    const SimC &S = sS;
    SimD &dy = D.simd[0];
    if (dy.status != VX3_SIM_RUNNING) return; // uniform over the grid
    const float dtF = dy.dt;
    if (dtF == 0) return;
    const double dt = dtF;
    double t = dy.t;
    const unsigned int G = gridDim.x;
    const bool vary = S.vary_temp && S.temp_period > 0;

    // ---- my link ----
    int g = -1;
    if ((int)threadIdx.x < links_per_cta) {
        g = blockIdx.x * links_per_cta + threadIdx.x;
        if (g >= D.nlinkslots) g = -1;
    }
    LinkRegs L;
    int2 ends = make_int2(-1, -1);
    LinkMatC lm;
    float pdN = 0, pdP = 0, numN = 0, numP = 0;
    double szN = 0, szP = 0;
    float cteN = 0, cteP = 0;
    double onN = 0, onP = 0; // thermal_on_after
    bool fixN = false, fixP = false, intN = false, intP = false;
    if (g >= 0) {
        ends = D.lends[g];
        L.state = D.lstate[g];
        if (ends.x < 0 || (L.state & (LKS_DETACHED | LKS_REMOVED))) g = -1;
    }
    if (g >= 0) {
        const VoxMatC &mN = D.vmat_tab[D.vmat[ends.x]], &mP = D.vmat_tab[D.vmat[ends.y]];
        if (mN.fixed && mP.fixed) g = -1;
        else {
            const int axis = (L.state & LKS_AXIS_MASK) >> LKS_AXIS_SHIFT;
            lm = D.lmat_tab[D.lmat[g]];
            const double2 h0 = *D.lh(0, g), h1 = *D.lh(1, g), h2 = *D.lh(2, g), h3 = *D.lh(3, g), h4 = *D.lh(4, g);
            L.pos2 = V3(h0.x, h0.y, h1.x);
            L.angle1v = V3(h1.y, h2.x, h2.y);
            L.angle2v = V3(h3.x, h3.y, h4.x);
            L.rest = h4.y;
            const float4 sn = D.lstrain[g];
            L.strain = sn.x; L.maxStrain = sn.y; L.strainOffset = sn.z; L.stress = sn.w;
            const float2 ar = D.larea[g];
            L.area = ar.x; L.tsum = ar.y;
            L.state &= ~LKS_JUST_CREATED;
            pdN = unpack_pd(D.pose[8 * (size_t)ends.x + 7]); pdP = unpack_pd(D.pose[8 * (size_t)ends.y + 7]);
            numN = mN.dampMultNum; numP = mP.dampMultNum;
            szN = mN.size[axis]; szP = mP.size[axis];
            cteN = mN.alphaCTE; cteP = mP.alphaCTE;
            onN = mN.thermal_on_after; onP = mP.thermal_on_after;
            fixN = mN.fixed; fixP = mP.fixed;
            // does the voxel phase integrate this end (then its previousDt becomes dt)?
            intN = !mN.fixed && !(D.vflags[ends.x] & VXF_REMOVED);
            intP = !mP.fixed && !(D.vflags[ends.y] & VXF_REMOVED);
        }
    }
    // ---- my voxel ----
    int v = -1;
    if ((int)threadIdx.x < vox_per_cta) {
        v = blockIdx.x * vox_per_cta + threadIdx.x;
        if (v >= D.nvox) v = -1;
    }
    VoxRegs r;
    VoxMatC vm;
    float tempe = 0, tempe_next = 0; // this step's temperature / the next step's (published in the pose record)
    double phase = 0;
    int vl[6] = {-1, -1, -1, -1, -1, -1};
    const ExtC *px = nullptr;
    short ic[3] = {0, 0, 0};
    bool vthermal = false, vint = false;
    if (v >= 0) {
        vm = D.vmat_tab[D.vmat[v]];
        r.flags = D.vflags[v];
        phase = D.phase[v];
        vthermal = vary && !(r.flags & VXF_REMOVED) && !vm.fixed;
        vint = !(r.flags & VXF_REMOVED) && !vm.fixed;
        load_pose(D.pose, v, r.pos, r.orient);
        tempe_next = unpack_t(D.pose[8 * (size_t)v + 7]);
        tempe = D.tempe[v];
        const double2 m0 = *D.mo(0, v), m1 = *D.mo(1, v), m2 = *D.mo(2, v);
        r.linMom = V3(m0.x, m0.y, m1.x);
        r.angMom = V3(m1.y, m2.x, m2.y);
#pragma unroll
        for (int i = 0; i < 6; i++) vl[i] = D.vlinks[6 * (size_t)v + i];
        const int ext = D.vext[v];
        px = ext >= 0 ? &D.exts[ext] : nullptr;
        ic[0] = D.ixyz[3 * (size_t)v]; ic[1] = D.ixyz[3 * (size_t)v + 1]; ic[2] = D.ixyz[3 * (size_t)v + 2];
    }
    const bool fixedAll = px && (px->dof & 0x3F) == 0x3F;

    unsigned int phase_no = 0; // barriers passed
    long long done = 0;
    int status = VX3_SIM_RUNNING;
    // point-to-point mode: thread k polls the k-th producer CTA of each phase
    const bool use_p2p = p2p.flags != nullptr;
    int dep_link = -1, dep_vox = -1; // producer (voxel-phase CTA) for my link phase / (link-phase CTA) for my voxel phase
    unsigned int *my_lflag = nullptr, *my_vflag = nullptr;
    const unsigned int *lflags = nullptr, *vflags = nullptr;
    if (use_p2p) {
        lflags = p2p.flags;
        vflags = p2p.flags + 32 * (size_t)G;
        my_lflag = p2p.flags + 32 * (size_t)blockIdx.x;
        my_vflag = p2p.flags + 32 * (size_t)(G + blockIdx.x);
        if ((int)threadIdx.x < p2p.ndeps[blockIdx.x]) dep_link = p2p.deps[(size_t)blockIdx.x * VX3_PERSIST_MAX_DEPS + threadIdx.x];
        if ((int)threadIdx.x < p2p.ndeps[G + blockIdx.x]) dep_vox = p2p.deps[(size_t)(G + blockIdx.x) * VX3_PERSIST_MAX_DEPS + threadIdx.x];
    }

    if (timing) c0 = clock64();
    for (long long s = 0; s < nsteps; s++) {
        // ================= link phase (gpu_update_links) =================
        if (g >= 0) {
            V3 pN, pP;
            Q4 qN, qP;
            {
                const double *a = D.pose + 8 * (size_t)ends.x, *b = D.pose + 8 * (size_t)ends.y;
                const double2 a0 = ldcg2(a), a1 = ldcg2(a + 2), a2 = ldcg2(a + 4), a3 = ldcg2(a + 6);
                const double2 b0 = ldcg2(b), b1 = ldcg2(b + 2), b2 = ldcg2(b + 4), b3 = ldcg2(b + 6);
                pN = V3(a0.x, a0.y, a1.x); qN = Q4(a1.y, a2.x, a2.y, a3.x);
                pP = V3(b0.x, b0.y, b1.x); qP = Q4(b1.y, b2.x, b2.y, b3.x);
                // updateRestLength() with the ends' temperatures for this step (published by their voxel passes)
                if (vary && ((!fixN && !(onN > t)) || (!fixP && !(onP > t)))) {
                    const float tN = unpack_t(a3.y), tP = unpack_t(b3.y);
                    L.rest = 0.5 * (szN * (1 + tN * cteN) + szP * (1 + tP * cteP));
                }
            }
            LinkOut o;
            link_update_forces(L, lm, D.strain_pool, D.stress_pool, pN, qN, pP, qP, numN / pdN, numP / pdP, o);
            *D.lf(0, g) = make_double2(o.forceNeg.x, o.forceNeg.y);
            *D.lf(1, g) = make_double2(o.forceNeg.z, o.momentNeg.x);
            *D.lf(2, g) = make_double2(o.momentNeg.y, o.momentNeg.z);
            *D.lf(3, g) = make_double2(o.forcePos.x, o.forcePos.y);
            *D.lf(4, g) = make_double2(o.forcePos.z, o.momentPos.x);
            *D.lf(5, g) = make_double2(o.momentPos.y, o.momentPos.z);
            if (L.strain > 100) atomicExch(&bar[1], 1u);
            if (intN) pdN = dtF;
            if (intP) pdP = dtF;
        }
        TICK(0);
        if (use_p2p) p2p_publish(my_lflag, (unsigned int)(s + 1));
        else grid_arrive(bar);
        // --- while the arrivals propagate: the temperature the NEXT step starts with (gpu_update_temperature at t+dt) ---
        if (v >= 0) {
            tempe = tempe_next;
            if (vthermal && !(vm.thermal_on_after > t + dtF)) tempe_next = voxel_temperature(S, t + dtF, phase);
        }
        TICK(1);
        unsigned int div;
        if (use_p2p) div = p2p_wait(lflags, dep_vox, (unsigned int)(s + 1), &bar[1], &s_div);
        else {
            grid_wait(bar, ++phase_no * G);
            div = ld_relaxed_u32(&bar[1]);
        }
        TICK(2);
        if (div) { // a link diverged in this step: doTimeStep returns false before the voxel pass
            status = VX3_SIM_DIVERGED;
            done = s + 1;
            break;
        }
        // ================= voxel phase (gpu_update_voxels) =================
        if (v >= 0 && vint) {
            V3 F(0, 0, 0), M(0, 0, 0);
#pragma unroll
            for (int i = 0; i < 6; i++) {
                if (vl[i] >= 0) {
                    const int k0 = (i & 1) ? 3 : 0;
                    const double2 a = __ldcg(D.lf(k0, vl[i])), b = __ldcg(D.lf(k0 + 1, vl[i])), c = __ldcg(D.lf(k0 + 2, vl[i]));
                    F += V3(a.x, a.y, b.x);
                    M += V3(b.y, c.x, c.y);
                }
            }
            V3 ff(0, 0, 0);
            if (S.has_ff && !fixedAll) {
                double vars[9];
                prog_vars(S, dy, t, r.pos.x, r.pos.y, r.pos.z, vars);
                ff.x = eval_slot(D, S, VX3_PROG_FORCE_X, vars, 0.0);
                ff.y = eval_slot(D, S, VX3_PROG_FORCE_Y, vars, 0.0);
                ff.z = eval_slot(D, S, VX3_PROG_FORCE_Z, vars, 0.0);
            }
            voxel_time_step(r, vm, px, ic[0], ic[1], ic[2], tempe, F, M, V3(), V3(), ff, dt);
            if (S.has_attach_cond) {
                double vars[9];
                prog_vars(S, dy, t, r.pos.x, r.pos.y, r.pos.z, vars);
                bool all = true;
                for (int c = 0; c < 5 && all; c++) all = eval_slot(D, S, VX3_PROG_ATTACH_0 + c, vars, 1.0) > 0;
                if (all) r.flags |= VXF_ENABLE_ATTACH;
                else r.flags &= ~VXF_ENABLE_ATTACH;
            }
            store_pose(D.pose, v, r.pos, r.orient, tempe_next, dtF);
        }
        t += dtF; // currentTime += dt (:352)
        done = s + 1;
        TICK(3);
        if (use_p2p) p2p_publish(my_vflag, (unsigned int)(s + 1));
        else grid_arrive(bar);
        // --- while the arrivals propagate: the stop condition for the next step ---
        if (check_stop && threadIdx.x == 0) {
            s_stop = 0;
            if (S.prog_n[VX3_PROG_STOP] > 0) {
                double vars[9];
                prog_vars(S, dy, t, dy.com[0], dy.com[1], dy.com[2], vars);
                bool ok;
                s_stop = mt_eval<VX3_MAX_TOKENS>(D.tokens + S.prog_off[VX3_PROG_STOP], S.prog_n[VX3_PROG_STOP], vars, &ok) > 0;
            }
        }
        if (use_p2p) {
            if (p2p_wait(vflags, dep_link, (unsigned int)(s + 1), &bar[1], &s_div)) { // someone diverged meanwhile
                status = VX3_SIM_DIVERGED;
                break;
            }
        } else
            grid_wait(bar, ++phase_no * G);
        TICK(4);
        if (check_stop && s_stop) { // identical in every CTA: CoM, angle, ... only change on the streaming path's sampling steps
            status = VX3_SIM_STOPPED;
            break;
        }
    }

    if (use_p2p && status == VX3_SIM_DIVERGED) { // CTAs may leave at different steps: never let a neighbour wait for me
        __syncthreads();
        if (threadIdx.x == 0) {
            st_release_u32(my_lflag, 0xFFFFFFFFu);
            st_release_u32(my_vflag, 0xFFFFFFFFu);
        }
    }
    // ---- write the register-resident state back ----
    if (g >= 0) {
        *D.lh(0, g) = make_double2(L.pos2.x, L.pos2.y);
        *D.lh(1, g) = make_double2(L.pos2.z, L.angle1v.x);
        *D.lh(2, g) = make_double2(L.angle1v.y, L.angle1v.z);
        *D.lh(3, g) = make_double2(L.angle2v.x, L.angle2v.y);
        *D.lh(4, g) = make_double2(L.angle2v.z, L.rest);
        D.lstrain[g] = make_float4(L.strain, L.maxStrain, L.strainOffset, L.stress);
        D.lstate[g] = L.state;
    }
    if (v >= 0) {
        *D.mo(0, v) = make_double2(r.linMom.x, r.linMom.y);
        *D.mo(1, v) = make_double2(r.linMom.z, r.angMom.x);
        *D.mo(2, v) = make_double2(r.angMom.y, r.angMom.z);
        D.vflags[v] = r.flags;
        D.tempe[v] = tempe;
    }
    if (timing && threadIdx.x == 0) {
        unsigned long long *o = reinterpret_cast<unsigned long long *>(bar + 4) + 8 * blockIdx.x;
        for (int i = 0; i < 5; i++) o[i] = (unsigned long long)tk[i];
        o[5] = (unsigned long long)done;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        dy.t = t;
        dy.steps += done;
        dy.status = status;
        if (status == VX3_SIM_DIVERGED) dy.diverged = 1;
    }
}

// host side ---------------------------------------------------------------------------------------------------
// Decides whether the batch qualifies and sizes the launch; fills the producer lists of the point-to-point phase flags.
// The device arrays (barrier, flags, deps, ndeps) are slices of the batch arena, placed by the caller (vx3_engine.cu).
inline void persistent_plan(PersistentPlan &p, const std::vector<SimC> &simc, bool any_collide, bool any_detach, bool any_cilia,
                            const cudaDeviceProp &prop, const std::vector<int2> &lends, const std::vector<int32_t> &vlinks,
                            std::vector<int> &deps, std::vector<int> &ndeps) {
    p.ok = false;
    p.p2p = false;
    if (simc.size() != 1 || any_collide || any_detach || any_cilia) return;
    if (!prop.cooperativeLaunch) return;
    const int L = simc[0].lcap, V = simc[0].nvox;
    int G = prop.multiProcessorCount;
    const int most = L > V ? L : V;
    if ((most + 31) / 32 < G) G = (most + 31) / 32;
    if (G < 1) G = 1;
    const int lpc = (L + G - 1) / G, vpc = (V + G - 1) / G;
    int T = lpc > vpc ? lpc : vpc;
    T = (T + 31) / 32 * 32;
    if (T < 32) T = 32;
    if (T > VX3_PERSIST_MAX_BLOCK) return; // body too large for one item per thread: streaming path
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_persistent, T, 0) != cudaSuccess || nb < 1) {
        cudaGetLastError();
        return;
    }
    p.timing = getenv("VX3_PERSIST_TIMING") != nullptr;
    p.grid = G;
    p.block = T;
    p.links_per_cta = lpc;
    p.vox_per_cta = vpc;
    p.ok = true;
    // ---- producer lists for the point-to-point phase flags ----
    if (getenv("VX3_PERSIST_GLOBAL_BARRIER")) return;
    deps.assign(2 * (size_t)G * VX3_PERSIST_MAX_DEPS, -1);
    ndeps.assign(2 * (size_t)G, 0);
    auto add = [&](int which, int cta, int producer) -> bool {
        if (producer == cta) return true; // own phases are ordered by program order
        int *d = &deps[((size_t)which * G + cta) * VX3_PERSIST_MAX_DEPS];
        int &n = ndeps[(size_t)which * G + cta];
        for (int k = 0; k < n; k++)
            if (d[k] == producer) return true;
        if (n >= VX3_PERSIST_MAX_DEPS || n >= T) return false;
        d[n++] = producer;
        return true;
    };
    bool fits = true;
    for (int g = 0; g < L && fits; g++) { // link phase of CTA g/lpc reads the poses owned by the voxel CTAs of its ends
        const int2 e = lends[simc[0].loff + g];
        if (e.x < 0) continue;
        fits = add(0, g / lpc, (e.x - simc[0].voff) / vpc) && add(0, g / lpc, (e.y - simc[0].voff) / vpc);
    }
    for (int v = 0; v < V && fits; v++) // voxel phase of CTA v/vpc reads the forces owned by the link CTAs of its links
        for (int i = 0; i < 6 && fits; i++) {
            const int li = vlinks[6 * ((size_t)simc[0].voff + v) + i];
            if (li >= 0) fits = add(1, v / vpc, (li - simc[0].loff) / lpc);
        }
    if (!fits) return; // too many neighbours for one poll per thread: keep the grid barrier
    // symmetric closure: whoever reads my data must also be waited for before I overwrite it (WAR)
    for (int c = 0; c < G && fits; c++) {
        for (int k = 0; k < ndeps[c] && fits; k++) fits = add(1, deps[(size_t)c * VX3_PERSIST_MAX_DEPS + k], c);
        for (int k = 0; k < ndeps[G + c] && fits; k++) fits = add(0, deps[((size_t)G + c) * VX3_PERSIST_MAX_DEPS + k], c);
    }
    if (!fits) return;
    p.p2p = true;
}

inline int persistent_run(PersistentPlan &p, const Dev &D, cudaStream_t st, long long nsteps, bool check_stop, long long *launches) {
    if (!p.ok) return -1;
    if (cudaMemsetAsync(p.barrier, 0, 2 * sizeof(unsigned int), st) != cudaSuccess) return -1;
    // the arrival counter is 32-bit: 2 barriers per step, grid arrivals each
    const long long max_chunk = 0x7FFFFFFFll / (2ll * p.grid) - 4;
    while (nsteps > 0) {
        long long n = nsteps < max_chunk ? nsteps : max_chunk;
        int cs = check_stop ? 1 : 0;
        Dev d = D;
        int tm = p.timing ? 1 : 0;
        P2P pp;
        pp.flags = p.p2p ? p.flags : nullptr;
        pp.deps = p.deps;
        pp.ndeps = p.ndeps;
        if (p.p2p && cudaMemsetAsync(p.flags, 0, 2 * (size_t)p.grid * 32 * sizeof(unsigned int), st) != cudaSuccess) return -1;
        void *args[] = {(void *)&d, (void *)&n, (void *)&cs, (void *)&p.links_per_cta, (void *)&p.vox_per_cta, (void *)&p.barrier, (void *)&tm, (void *)&pp};
        if (cudaLaunchCooperativeKernel((const void *)k_persistent, dim3(p.grid), dim3(p.block), args, 0, st) != cudaSuccess) return -1;
        if (launches) (*launches)++;
        if (p.timing) {
            std::vector<unsigned long long> h(8 * (size_t)p.grid);
            cudaStreamSynchronize(st);
            cudaMemcpy(h.data(), p.barrier + 4, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            const char *nm[5] = {"link", "arrive1+temp", "wait1", "voxel", "arrive2+wait2"};
            fprintf(stderr, "[persist timing] %llu steps, cycles/step over %d CTAs (min / mean / max):", h[5], p.grid);
            for (int k = 0; k < 5; k++) {
                double mn = 1e30, mx = 0, sum = 0;
                for (int c = 0; c < p.grid; c++) {
                    const double x = (double)h[8 * (size_t)c + k] / (h[8 * (size_t)c + 5] ? (double)h[8 * (size_t)c + 5] : 1.0);
                    mn = x < mn ? x : mn;
                    mx = x > mx ? x : mx;
                    sum += x;
                }
                fprintf(stderr, "  %s %.0f/%.0f/%.0f", nm[k], mn, sum / p.grid, mx);
            }
            fprintf(stderr, "\n");
        }
        nsteps -= n;
        if (nsteps > 0 && cudaMemsetAsync(p.barrier, 0, 2 * sizeof(unsigned int), st) != cudaSuccess) return -1;
    }
    return 0;
}

} // namespace vx3
