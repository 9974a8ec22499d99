// On-chip persistent step kernel for ONE small collision-free body (BASELINE config 2 class).
//
// The streaming path pays three dependent kernel launches per doTimeStep, which is launch/latency bound for a body
// whose whole state (a few MB) fits on chip.  Here one cooperative grid (one CTA per SM) runs many steps per launch.
//
// Decomposition: the host cuts the body into one compact spatial block of voxels per CTA (recursive coordinate
// bisection of the lattice coordinates).  A CTA owns its block's voxels and evaluates EVERY link that touches one of
// them — a link across a block face is evaluated by both neighbours, from the same inputs with the same code, hence
// to the same bits — so the end forces a voxel gathers never leave the SM: they go through shared memory.  Every
// thread permanently owns at most one link and one voxel whose private state (link history, strain, momenta, flags)
// stays in REGISTERS for the whole launch.  The only data that crosses the chip is the voxel pose record (64 B), ONCE
// per step: poses are double-buffered by step parity (no write-after-read hazard, hence no second synchronisation),
// and a CTA starts step s as soon as the CTAs that own the far ends of its face links have published step s-1
// (point-to-point flags, release/acquire).  The reference orders the same two phases with two device-wide child-grid
// syncs per step (src/VX3/VX3_VoxelyzeKernel.cu:259-269 gpu_update_links, :306-312 gpu_update_voxels).
//
// Voxel phase in three roles: a voxel's step is three nearly independent latency chains — translation (force, floor,
// friction, position), rotation (moment, quaternion update: sqrt/sin/cos) and the next step's temperature (sin) — so three
// threads in three different warps run them concurrently (lanes [0,RO) translate, [RO,2RO) rotate, [2RO,3RO) temperature;
// RO = the CTA's voxel count rounded up to a warp).  They meet through shared memory (old orientation, this step's
// temperature, the "resting on the floor" decision that clears the angular momentum) and each stores its own part of
// the 64-byte pose record.  Temperature: each voxel's temperature for the next step travels in its pose record.
// The physics is the same code as the streaming kernels (vx3_physics.cuh), so the two paths alternate freely (the
// streaming path takes the CoM-sampling steps) and are bit-identical.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "vx3_kernels.cuh"

namespace vx3 {

#define VX3_PERSIST_MAX_BLOCK 256
#define VX3_PERSIST_MAX_DEPS 32

struct PersistentPlan {
    bool ok = false;
    int grid = 0, block = 0;
    bool timing = false;
    // arena slices (placed by vx3_engine.cu)
    unsigned int *ctl = nullptr;   // [0] divergence word, [1] exit counter, [4..] phase cycle counters (debug)
    unsigned int *flags = nullptr; // [grid][32]: steps completed by the CTA in this launch (one 128-B line per counter)
    int *lk_slot = nullptr;        // [grid][block]: global link slot of the lane (-1 none); bit 30 set = evaluated here for a neighbour's benefit only (not written back)
    int *vx_id = nullptr;          // [grid][block]: global voxel of the lane (-1 none)
    int *vx_lane = nullptr;        // [grid][block][6]: lane (in this CTA) of the voxel's link in each direction, -1 none
    int *deps = nullptr;           // [grid][VX3_PERSIST_MAX_DEPS]: CTAs whose poses this CTA's links read
    int *ndeps = nullptr;          // [grid]
    double *pose_alt = nullptr;    // [nvox][8]: odd-parity pose buffer
    int ro = 0;                    // voxel lanes per role
};
#define VX3_PERSIST_DUP (1 << 30)

struct PersistArgs {
    unsigned int *ctl, *flags;
    const int *lk_slot, *vx_id, *vx_lane, *deps, *ndeps;
    double *pose_alt;
    int ro; // role offset: voxel lanes per role (multiple of 32, 3 * ro <= block)
};
#define VX3_PERSIST_MAX_RO 96

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double2 ldcg2(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ void stcg2(double *p, double a, double b) { __stcg(reinterpret_cast<double2 *>(p), make_double2(a, b)); }

__global__ void __launch_bounds__(VX3_PERSIST_MAX_BLOCK, 1) k_persistent(Dev D, long long nsteps, int check_stop, int timing, PersistArgs A) {
    __shared__ int s_stop, s_div;
    __shared__ SimC sS;                                    // per-simulation constants on chip
    __shared__ double sF[VX3_PERSIST_MAX_BLOCK][13];       // end forces of the CTA's links: Fneg, Mneg, Fpos, Mpos (+1: a 12-double record stride maps every 4th lane to the same banks)
    const int T = blockDim.x, tid = threadIdx.x;
    for (int i = tid; i < (int)(sizeof(SimC) / 4); i += T) reinterpret_cast<int *>(&sS)[i] = reinterpret_cast<const int *>(&D.simc[0])[i];
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
    const double t0 = dy.t;
    double t = t0;
    const unsigned int G = gridDim.x;
    const bool vary = S.vary_temp && S.temp_period > 0;
    double *const P0 = D.pose, *const P1 = A.pose_alt;

    // ---- my link ----
    // Lane assignment: a warp executes the union of its lanes' branches, and the large-angle branch of orientLink
    // (VX3_Link.cu:111-125) costs ~50 % on top of the common path.  Typically a few links per CTA are in that regime, and
    // with the host's order nearly every warp holds one or two of them (profiles/: 85 % of the warps ran the branch
    // with 2 lanes active).  So each launch starts by sorting the CTA's links by regime: small-angle links fill the
    // lanes from 0 up, large-angle links from the last lane down — the regime is sticky (hysteresis), so most warps stay
    // uniform for the whole launch.  sMap translates the host's lane numbers (vx_lane) to the sorted ones.
    __shared__ int sSlot[VX3_PERSIST_MAX_BLOCK];
    __shared__ unsigned char sMap[VX3_PERSIST_MAX_BLOCK], sCls[VX3_PERSIST_MAX_BLOCK];
    {
        const int g0 = A.lk_slot[(size_t)blockIdx.x * T + tid];
        int cls = 2; // 0 small-angle, 1 large-angle, 2 no link
        if (g0 >= 0) cls = (D.lstate[g0 & ~VX3_PERSIST_DUP] & LKS_SMALL) ? 0 : 1;
        sCls[tid] = (unsigned char)cls;
        sSlot[tid] = -1;
        __syncthreads();
        if (tid == 0) {
            // Large-angle links go to the warp that has a scheduler to itself: warps are dealt to the SM's four schedulers
            // round-robin, so with nw warps the scheduler (nw % 4) hosts one warp fewer than schedulers 0..(nw % 4)-1 — its
            // first warp, index nw % 4, is the least contended place for the longest instruction stream.  Overflow goes
            // to the lanes from the top down; small-angle links fill the remaining lanes from 0 up.
            const int nw = T / 32, lw = nw % 4;
            int nl = 0;
            for (int k = 0; k < T; k++) nl += sCls[k] == 1;
            const int in_lw = lw < nw - 1 ? (nl < 32 ? nl : 32) : 0; // lanes [32*lw, 32*lw + in_lw), then (T - 1 - j) for the rest
            int li = 0, lo = 0;
            auto next_free = [&](int x) { // next lane not reserved for a large-angle link
                for (;; x++) {
                    if (x >= T) return x;
                    const bool res = (x >= 32 * lw && x < 32 * lw + in_lw) || (x > T - 1 - (nl - in_lw));
                    if (!res) return x;
                }
            };
            lo = next_free(0);
            for (int k = 0; k < T; k++) {
                if (sCls[k] == 1) {
                    sMap[k] = (unsigned char)(li < in_lw ? 32 * lw + li : T - 1 - (li - in_lw));
                    li++;
                } else if (sCls[k] == 0) {
                    sMap[k] = (unsigned char)lo;
                    lo = next_free(lo + 1);
                }
            }
            for (int k = 0; k < T; k++)
                if (sCls[k] == 2) { // empty lanes take what is left
                    sMap[k] = (unsigned char)lo;
                    lo = next_free(lo + 1);
                }
        }
        __syncthreads();
        sSlot[sMap[tid]] = g0;
        __syncthreads();
    }
    int g = sSlot[tid];
    const bool dup = g >= 0 && (g & VX3_PERSIST_DUP);
    if (g >= 0) g &= ~VX3_PERSIST_DUP;
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
            pdN = unpack_pd(P0[8 * (size_t)ends.x + 7]); pdP = unpack_pd(P0[8 * (size_t)ends.y + 7]);
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
    // a lane without a live link still owns a force record that voxels may index: keep it zero
#pragma unroll
    for (int k = 0; k < 12; k++) sF[tid][k] = 0.0;

    // ---- my voxel and my role in its step: 0 translate, 1 rotate, 2 temperature ----
    __shared__ double sOrient[2][VX3_PERSIST_MAX_RO][5]; // orientation by step parity: the rotate role writes the new one while the translate role reads the old
    __shared__ float sTemp[VX3_PERSIST_MAX_RO];          // this step's temperature (temperature role -> translate role)
    __shared__ int sZero[2][VX3_PERSIST_MAX_RO];         // translate role -> rotate role, by step parity: on the floor in static friction, clear angMom (VX3_Voxel.cu:259-264)
    const int role = tid / A.ro, vk = tid - role * A.ro;
    const int v = role < 3 ? A.vx_id[(size_t)blockIdx.x * T + vk] : -1;
    V3 pos, linMom, angMom;
    Q4 orient;
    int vflags = 0;
    VoxMatC vm;
    float tempe = 0, tempe_next = 0; // this step's temperature / the next step's (published in the pose record)
    float pd_cur = 0;
    double phase = 0;
    int vl[6] = {-1, -1, -1, -1, -1, -1}; // lane of the link in each direction
    const ExtC *px = nullptr;
    short ic[3] = {0, 0, 0};
    bool vthermal = false, vint = false;
    if (v >= 0) {
        vm = D.vmat_tab[D.vmat[v]];
        vflags = D.vflags[v];
        phase = D.phase[v];
        vthermal = vary && !(vflags & VXF_REMOVED) && !vm.fixed;
        vint = !(vflags & VXF_REMOVED) && !vm.fixed;
        load_pose(P0, v, pos, orient);
        const double tp = P0[8 * (size_t)v + 7];
        tempe_next = unpack_t(tp);
        pd_cur = unpack_pd(tp);
        tempe = D.tempe[v];
        const double2 m0 = *D.mo(0, v), m1 = *D.mo(1, v), m2 = *D.mo(2, v);
        linMom = V3(m0.x, m0.y, m1.x);
        angMom = V3(m1.y, m2.x, m2.y);
#pragma unroll
        for (int i = 0; i < 6; i++) {
            vl[i] = A.vx_lane[((size_t)blockIdx.x * T + vk) * 6 + i];
            if (vl[i] >= 0) vl[i] = sMap[vl[i]];
            if (D.vlinks[6 * (size_t)v + i] < 0) vl[i] = -1;
        }
        const int ext = D.vext[v];
        px = ext >= 0 ? &D.exts[ext] : nullptr;
        ic[0] = D.ixyz[3 * (size_t)v]; ic[1] = D.ixyz[3 * (size_t)v + 1]; ic[2] = D.ixyz[3 * (size_t)v + 2];
        if (role == 0) {
            // the odd-parity buffer starts as a copy: voxels that are never integrated keep their record in both
            store_pose(P1, v, pos, orient, tempe_next, pd_cur);
            sZero[0][vk] = 0;
        } else if (role == 1) {
            sOrient[0][vk][0] = orient.w; sOrient[0][vk][1] = orient.x; sOrient[0][vk][2] = orient.y; sOrient[0][vk][3] = orient.z;
        }
    }
    const bool fixedAll = px && (px->dof & 0x3F) == 0x3F;

    // point-to-point: thread k polls the k-th neighbour CTA
    int dep = -1;
    if (tid < A.ndeps[blockIdx.x]) dep = A.deps[(size_t)blockIdx.x * VX3_PERSIST_MAX_DEPS + tid];
    unsigned int *my_flag = A.flags + 32 * (size_t)blockIdx.x;
    const unsigned int *dep_flag = dep >= 0 ? A.flags + 32 * (size_t)dep : nullptr;
    unsigned int *divword = A.ctl; // 0 = no divergence; otherwise nsteps - s of the EARLIEST diverging step s (atomicMax)

    long long done = 0;
    int status = VX3_SIM_RUNNING;
    if (timing) c0 = clock64();
    long long s = 0;
    for (; s < nsteps; s++) {
        // ================= wait: the poses of step s are published by every CTA my links reach into =================
        const double *Pr = (s & 1) ? P1 : P0;
        double *Pw = (s & 1) ? P0 : P1;
        if (dep_flag) {
            // (the divergence word is looked at every 8th poll only: behind the acquire it would cost a second L2 round trip per poll)
            for (int spins = 0; ld_acquire_u32(dep_flag) < (unsigned int)s;)
                if ((++spins & 7) == 0 && ld_relaxed_u32(divword)) break; // the producer may have left: the simulation is over
        }
        if (tid == 0) s_div = (int)ld_relaxed_u32(divword);
        __syncthreads();
        if (s_div) { // somebody's link diverged at a step <= s: doTimeStep returned false there
            status = VX3_SIM_DIVERGED;
            break;
        }
        TICK(0);
        // ================= link phase (gpu_update_links) =================
        bool mydiv = false;
        if (g >= 0) {
            V3 pN, pP;
            Q4 qN, qP;
            {
                const double *a = Pr + 8 * (size_t)ends.x, *b = Pr + 8 * (size_t)ends.y;
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
            double *f = sF[tid];
            f[0] = o.forceNeg.x; f[1] = o.forceNeg.y; f[2] = o.forceNeg.z; f[3] = o.momentNeg.x; f[4] = o.momentNeg.y; f[5] = o.momentNeg.z;
            f[6] = o.forcePos.x; f[7] = o.forcePos.y; f[8] = o.forcePos.z; f[9] = o.momentPos.x; f[10] = o.momentPos.y; f[11] = o.momentPos.z;
            // a state read-back sees the end forces in global memory: only the last step's matter (any step's when a stop condition
            // may end the launch) — 18 KB of stores per CTA and step that the release fence of the publish would otherwise wait for
            if (!dup && (check_stop || s == nsteps - 1)) {
                *D.lf(0, g) = make_double2(o.forceNeg.x, o.forceNeg.y);
                *D.lf(1, g) = make_double2(o.forceNeg.z, o.momentNeg.x);
                *D.lf(2, g) = make_double2(o.momentNeg.y, o.momentNeg.z);
                *D.lf(3, g) = make_double2(o.forcePos.x, o.forcePos.y);
                *D.lf(4, g) = make_double2(o.forcePos.z, o.momentPos.x);
                *D.lf(5, g) = make_double2(o.momentPos.y, o.momentPos.z);
            }
            mydiv = L.strain > 100;
            if (intN) pdN = dtF;
            if (intP) pdP = dtF;
        }
        TICK(1);
        // --- this step's temperature for the translate role (floor penetration) ---
        if (v >= 0 && role == 2) {
            tempe = tempe_next;
            sTemp[vk] = tempe;
        }
        if (__syncthreads_or(mydiv)) { // a link of this CTA diverged in this step: doTimeStep returns false before the voxel pass
            if (tid == 0) atomicMax(divword, (unsigned int)(nsteps - s));
            status = VX3_SIM_DIVERGED;
            break;
        }
        TICK(2);
        // ================= voxel phase (gpu_update_voxels), three roles in three warps =================
        if (v >= 0 && vint) {
            double *w = Pw + 8 * (size_t)v;
            if (role == 0) { // ---- translate ----
                V3 F(0, 0, 0);
#pragma unroll
                for (int i = 0; i < 6; i++)
                    if (vl[i] >= 0) {
                        const double *f = sF[vl[i]] + ((i & 1) ? 6 : 0);
                        F += V3(f[0], f[1], f[2]);
                    }
                V3 ff(0, 0, 0);
                if (S.has_ff && !fixedAll) {
                    double vars[9];
                    prog_vars(S, dy, t, pos.x, pos.y, pos.z, vars);
                    ff.x = eval_slot(D, S, VX3_PROG_FORCE_X, vars, 0.0);
                    ff.y = eval_slot(D, S, VX3_PROG_FORCE_Y, vars, 0.0);
                    ff.z = eval_slot(D, S, VX3_PROG_FORCE_Z, vars, 0.0);
                }
                const double *oq = sOrient[s & 1][vk];
                const Q4 orient0(oq[0], oq[1], oq[2], oq[3]);
                tempe = sTemp[vk];
                voxel_step_translate(pos, linMom, vflags, orient0, vm, px, ic[0], ic[1], ic[2], tempe, F, V3(), V3(), ff, dt);
                sZero[(s + 1) & 1][vk] = voxel_step_join(pos, vflags, vm, px, tempe) ? 1 : 0;
                if (S.has_attach_cond) {
                    double vars[9];
                    prog_vars(S, dy, t, pos.x, pos.y, pos.z, vars);
                    bool all = true;
                    for (int c = 0; c < 5 && all; c++) all = eval_slot(D, S, VX3_PROG_ATTACH_0 + c, vars, 1.0) > 0;
                    if (all) vflags |= VXF_ENABLE_ATTACH;
                    else vflags &= ~VXF_ENABLE_ATTACH;
                }
                stcg2(w, pos.x, pos.y);
                __stcg(w + 2, pos.z);
            } else if (role == 1) { // ---- rotate ----
                if (sZero[s & 1][vk]) angMom = V3(0, 0, 0); // the previous step's join
                V3 M(0, 0, 0);
#pragma unroll
                for (int i = 0; i < 6; i++)
                    if (vl[i] >= 0) {
                        const double *f = sF[vl[i]] + ((i & 1) ? 9 : 3);
                        M += V3(f[0], f[1], f[2]);
                    }
                voxel_step_rotate(orient, angMom, vm, px, M, dt);
                double *nq = sOrient[(s + 1) & 1][vk];
                nq[0] = orient.w; nq[1] = orient.x; nq[2] = orient.y; nq[3] = orient.z;
                __stcg(w + 3, orient.w);
                stcg2(w + 4, orient.x, orient.y);
                __stcg(w + 6, orient.z);
            } else { // ---- the temperature the NEXT step starts with (gpu_update_temperature at t+dt) ----
                if (vthermal && !(vm.thermal_on_after > t + dtF)) tempe_next = voxel_temperature(S, t + dtF, phase);
                pd_cur = dtF;
                __stcg(w + 7, pack_tp(tempe_next, dtF));
            }
        }
        t += dtF; // currentTime += dt (:352)
        done = s + 1;
        TICK(3);
        // ================= publish step s+1 =================
        __syncthreads(); // every thread's pose stores are issued (and ordered before thread 0's release)
        if (tid == 0) {
            st_release_u32(my_flag, (unsigned int)(s + 1));
            if (check_stop) { // the stop condition for the next step: identical in every CTA (CoM, angle, ... only change on the streaming path's sampling steps)
                s_stop = 0;
                if (S.prog_n[VX3_PROG_STOP] > 0) {
                    double vars[9];
                    prog_vars(S, dy, t, dy.com[0], dy.com[1], dy.com[2], vars);
                    bool ok;
                    s_stop = mt_eval<VX3_MAX_TOKENS>(D.tokens + S.prog_off[VX3_PROG_STOP], S.prog_n[VX3_PROG_STOP], vars, &ok) > 0;
                }
            }
        }
        if (check_stop) {
            __syncthreads();
            if (s_stop) {
                status = VX3_SIM_STOPPED;
                break;
            }
        }
        TICK(4);
    }

    // ---- write the register-resident state back ----
    if (g >= 0 && !dup) {
        *D.lh(0, g) = make_double2(L.pos2.x, L.pos2.y);
        *D.lh(1, g) = make_double2(L.pos2.z, L.angle1v.x);
        *D.lh(2, g) = make_double2(L.angle1v.y, L.angle1v.z);
        *D.lh(3, g) = make_double2(L.angle2v.x, L.angle2v.y);
        *D.lh(4, g) = make_double2(L.angle2v.z, L.rest);
        D.lstrain[g] = make_float4(L.strain, L.maxStrain, L.strainOffset, L.stress);
        D.lstate[g] = L.state;
    }
    // (the translate role's last join decision is visible to the rotate role: every exit path passes a CTA barrier after it)
    if (v >= 0) {
        double *mo0 = reinterpret_cast<double *>(D.mo(0, v)), *mo1 = reinterpret_cast<double *>(D.mo(1, v)), *mo2 = reinterpret_cast<double *>(D.mo(2, v));
        if (role == 0) {
            mo0[0] = linMom.x; mo0[1] = linMom.y; mo1[0] = linMom.z;
            D.vflags[v] = vflags;
        } else if (role == 1) {
            if (vint && sZero[done & 1][vk]) angMom = V3(0, 0, 0);
            mo1[1] = angMom.x; mo2[0] = angMom.y; mo2[1] = angMom.z;
        } else
            D.tempe[v] = tempe;
    }
    if (timing && tid == 0) {
        unsigned long long *o = reinterpret_cast<unsigned long long *>(A.ctl + 4) + 8 * blockIdx.x;
        for (int i = 0; i < 5; i++) o[i] = (unsigned long long)tk[i];
        o[5] = (unsigned long long)done;
    }
    // ---- the last CTA to leave writes the simulation's scalars; every CTA then restores its voxels' records in the
    // batch's pose array (the even buffer) — only after ALL CTAs have left their loops, because a slower neighbour may
    // still be reading the even buffer ----
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int left = atomicAdd(A.ctl + 1, 1u);
        if (left == G - 1) {
            __threadfence();
            const unsigned int dv = ld_relaxed_u32(divword);
            if (dv) { // diverged in the link pass of step sd: CurStepCount was incremented, time was not (:241-281)
                const long long sd = nsteps - (long long)dv;
                double td = t0;
                for (long long i = 0; i < sd; i++) td += dtF;
                dy.t = td;
                dy.steps += sd + 1;
                dy.status = VX3_SIM_DIVERGED;
                dy.diverged = 1;
            } else {
                dy.t = t;
                dy.steps += done;
                dy.status = status;
            }
        }
        while (ld_acquire_u32(A.ctl + 1) < G) {}
    }
    __syncthreads();
    if (v >= 0) { // each role restores its part of the record in the batch's pose array (the even buffer)
        double *w = P0 + 8 * (size_t)v;
        if (role == 0) { w[0] = pos.x; w[1] = pos.y; w[2] = pos.z; }
        else if (role == 1) { w[3] = orient.w; w[4] = orient.x; w[5] = orient.y; w[6] = orient.z; }
        else w[7] = pack_tp(tempe_next, pd_cur);
    }
}

// host side ---------------------------------------------------------------------------------------------------
struct PersistentTables {
    std::vector<int> lk_slot, vx_id, vx_lane, deps, ndeps;
};

// The block partition and the lane tables are a pure function of the body's lattice and link topology, and a worker
// that evaluates one batch after another (or re-creates the same body: bench.py's end-to-end loop) sees the same
// topology again and again: the last few plans are kept, keyed by a 128-bit hash of (coordinates, link ends,
// voxel link slots, offsets, SM count).  A hit replaces the bisection + table build (~1.5 ms for a 20^3 body).
struct PersistentPlanCache {
    struct Entry {
        unsigned long long h0 = 0, h1 = 0;
        int grid = 0, block = 0, ro = 0;
        PersistentTables tb;
        unsigned long long stamp = 0;
    };
    std::mutex mu;
    std::vector<Entry> entries;
    unsigned long long clock = 0;
    static PersistentPlanCache &get() {
        static PersistentPlanCache c;
        return c;
    }
    static void mix(unsigned long long &h0, unsigned long long &h1, const void *data, size_t bytes) {
        const unsigned char *p = static_cast<const unsigned char *>(data);
        size_t i = 0;
        for (; i + 8 <= bytes; i += 8) {
            unsigned long long w;
            memcpy(&w, p + i, 8);
            h0 = (h0 ^ w) * 0x9E3779B97F4A7C15ull;
            h0 ^= h0 >> 29;
            h1 = (h1 + w) * 0xC2B2AE3D27D4EB4Full;
            h1 ^= h1 >> 31;
        }
        for (; i < bytes; i++) {
            h0 = (h0 ^ p[i]) * 0x100000001B3ull;
            h1 = (h1 + p[i]) * 0x9E3779B97F4A7C15ull;
        }
    }
};

// recursive coordinate bisection: voxels idx[lo, hi) go to CTAs [c0, c0 + nc).  The selection runs on packed 64-bit keys
// (coordinate along the cut axis, voxel index) in a scratch array: plain integer compares, no indirection — (coordinate,
// index) is a strict total order, so the partition is unique whatever nth_element does inside.
inline void persist_rcb(std::vector<int> &idx, int lo, int hi, int c0, int nc, const int16_t *ixyz, int voff, std::vector<int> &cta_of, std::vector<unsigned long long> &keys,
                        int depth = 0) {
    if (nc == 1) {
        for (int i = lo; i < hi; i++) cta_of[idx[i]] = c0;
        return;
    }
    int mn[3] = {1 << 30, 1 << 30, 1 << 30}, mx[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
    for (int i = lo; i < hi; i++) {
        const int16_t *c = ixyz + 3 * ((size_t)voff + idx[i]);
        for (int a = 0; a < 3; a++) {
            mn[a] = std::min(mn[a], (int)c[a]);
            mx[a] = std::max(mx[a], (int)c[a]);
        }
    }
    int ax = 0;
    for (int a = 1; a < 3; a++)
        if (mx[a] - mn[a] > mx[ax] - mn[ax]) ax = a;
    const int ncl = nc / 2;
    const int mid = lo + (int)((long long)(hi - lo) * ncl / nc);
    for (int i = lo; i < hi; i++) keys[i] = ((unsigned long long)(unsigned)(ixyz[3 * ((size_t)voff + idx[i]) + ax] + 32768) << 32) | (unsigned)idx[i];
    std::nth_element(keys.begin() + lo, keys.begin() + mid, keys.begin() + hi);
    for (int i = lo; i < hi; i++) idx[i] = (int)(unsigned)keys[i];
    if (depth < 2 && hi - lo > 16384) { // the two halves touch disjoint ranges of idx, keys and cta_of: run the top levels on 4 threads
        std::thread left([&]() { persist_rcb(idx, lo, mid, c0, ncl, ixyz, voff, cta_of, keys, depth + 1); });
        persist_rcb(idx, mid, hi, c0 + ncl, nc - ncl, ixyz, voff, cta_of, keys, depth + 1);
        left.join();
    } else {
        persist_rcb(idx, lo, mid, c0, ncl, ixyz, voff, cta_of, keys, depth + 1);
        persist_rcb(idx, mid, hi, c0 + ncl, nc - ncl, ixyz, voff, cta_of, keys, depth + 1);
    }
}

// Decides whether the batch qualifies, cuts the body into blocks and builds the per-CTA lane tables.  The device arrays
// are slices of the batch arena, placed by the caller (vx3_engine.cu).
inline void persistent_plan(PersistentPlan &p, const std::vector<SimC> &simc, bool any_collide, bool any_dynamic_topology, bool any_cilia,
                            const cudaDeviceProp &prop, const int2 *lends, const int32_t *vlinks, const int16_t *ixyz, PersistentTables &tb) {
    p.ok = false;
    if (simc.size() != 1 || any_collide || any_dynamic_topology || any_cilia) return;
    if (!prop.cooperativeLaunch) return;
    const int L = simc[0].lcap, V = simc[0].nvox, voff = simc[0].voff, loff = simc[0].loff;
    if (V < 1) return;
    int G = prop.multiProcessorCount;
    if ((long long)V > (long long)G * VX3_PERSIST_MAX_RO || (long long)L > (long long)G * VX3_PERSIST_MAX_BLOCK) return; // cannot fit: skip the partitioning
    if ((V + 15) / 16 < G) G = (V + 15) / 16; // at least ~16 voxels per CTA
    if (G < 1) G = 1;
    unsigned long long h0 = 0x243F6A8885A308D3ull, h1 = 0x13198A2E03707344ull;
    const bool use_cache = getenv("VX3_NO_PLAN_CACHE") == nullptr;
    if (use_cache) {
        const int hdr[6] = {V, L, voff, loff, G, (int)sizeof(int2)};
        PersistentPlanCache::mix(h0, h1, hdr, sizeof(hdr));
        PersistentPlanCache::mix(h0, h1, ixyz + 3 * (size_t)voff, 3 * sizeof(int16_t) * (size_t)V);
        PersistentPlanCache::mix(h0, h1, lends + loff, sizeof(int2) * (size_t)L);
        PersistentPlanCache::mix(h0, h1, vlinks + 6 * (size_t)voff, 6 * sizeof(int32_t) * (size_t)V);
        PersistentPlanCache &pc = PersistentPlanCache::get();
        std::lock_guard<std::mutex> lk(pc.mu);
        for (auto &e : pc.entries)
            if (e.h0 == h0 && e.h1 == h1) {
                e.stamp = ++pc.clock;
                tb = e.tb;
                p.timing = getenv("VX3_PERSIST_TIMING") != nullptr;
                p.grid = e.grid;
                p.block = e.block;
                p.ro = e.ro;
                p.ok = true;
                return;
            }
    }
    const bool lapt = getenv("VX3_CREATE_TIMING") != nullptr;
    auto lt0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!lapt) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[create timing]   plan: %-20s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - lt0).count());
        lt0 = now;
    };
    std::vector<int> idx(V), cta_of(V, 0);
    for (int i = 0; i < V; i++) idx[i] = i;
    std::vector<unsigned long long> keys(V);
    persist_rcb(idx, 0, V, 0, G, ixyz, voff, cta_of, keys);
    lap("bisection");
    std::vector<std::vector<int>> vox(G), lnk(G);
    for (int i = 0; i < V; i++) vox[cta_of[i]].push_back(i); // ascending voxel index within a CTA
    for (int l = 0; l < L; l++) {
        const int2 e = lends[(size_t)loff + l];
        if (e.x < 0) continue;
        const int ca = cta_of[e.x - voff], cb = cta_of[e.y - voff];
        lnk[ca].push_back(l);
        if (cb != ca) lnk[cb].push_back(l | VX3_PERSIST_DUP); // owner = the CTA of the negative end
    }
    int most = 1;
    for (int c = 0; c < G; c++) most = std::max(most, (int)std::max(vox[c].size(), lnk[c].size()));
    int mostv = 1;
    for (int c = 0; c < G; c++) mostv = std::max(mostv, (int)vox[c].size());
    const int ro = (mostv + 31) / 32 * 32; // voxel lanes per role (translate / rotate / temperature)
    int T = (std::max(most, 3 * ro) + 31) / 32 * 32;
    if (T > VX3_PERSIST_MAX_BLOCK || ro > VX3_PERSIST_MAX_RO) return; // blocks too large for one item per thread: streaming path
    lap("block lists");
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_persistent, T, 0) != cudaSuccess || nb < 1) {
        cudaGetLastError();
        return;
    }
    lap("occupancy query");
    tb.lk_slot.assign((size_t)G * T, -1);
    tb.vx_id.assign((size_t)G * T, -1);
    tb.vx_lane.assign((size_t)G * T * 6, -1);
    tb.deps.assign((size_t)G * VX3_PERSIST_MAX_DEPS, -1);
    tb.ndeps.assign(G, 0);
    std::vector<int> lane_of(L, -1);
    for (int c = 0; c < G; c++) {
        for (size_t k = 0; k < lnk[c].size(); k++) {
            const int l = lnk[c][k] & ~VX3_PERSIST_DUP;
            tb.lk_slot[(size_t)c * T + k] = (loff + l) | (lnk[c][k] & VX3_PERSIST_DUP);
            lane_of[l] = (int)k;
            const int2 e = lends[(size_t)loff + l];
            for (int other : {cta_of[e.x - voff], cta_of[e.y - voff]}) {
                if (other == c) continue;
                int *d = &tb.deps[(size_t)c * VX3_PERSIST_MAX_DEPS];
                int &n = tb.ndeps[c];
                bool have = false;
                for (int q = 0; q < n; q++) have |= d[q] == other;
                if (have) continue;
                if (n >= VX3_PERSIST_MAX_DEPS || n >= T) return; // too many neighbours for one poll per thread
                d[n++] = other;
            }
        }
        for (size_t k = 0; k < vox[c].size(); k++) {
            const int vi = vox[c][k];
            tb.vx_id[(size_t)c * T + k] = voff + vi;
            for (int dir = 0; dir < 6; dir++) {
                const int li = vlinks[6 * ((size_t)voff + vi) + dir];
                if (li < 0) continue;
                const int l = li - loff;
                if (l < 0 || l >= L || lane_of[l] < 0) return; // inconsistent adjacency: leave it to the streaming path
                tb.vx_lane[((size_t)c * T + k) * 6 + dir] = lane_of[l];
            }
        }
        for (size_t k = 0; k < lnk[c].size(); k++) lane_of[lnk[c][k] & ~VX3_PERSIST_DUP] = -1;
    }
    lap("lane tables");
    p.timing = getenv("VX3_PERSIST_TIMING") != nullptr;
    p.grid = G;
    p.block = T;
    p.ro = ro;
    p.ok = true;
    if (use_cache) {
        PersistentPlanCache &pc = PersistentPlanCache::get();
        std::lock_guard<std::mutex> lk(pc.mu);
        if (pc.entries.size() >= 4) { // evict the least recently used plan
            size_t old = 0;
            for (size_t i = 1; i < pc.entries.size(); i++)
                if (pc.entries[i].stamp < pc.entries[old].stamp) old = i;
            pc.entries.erase(pc.entries.begin() + old);
        }
        PersistentPlanCache::Entry e;
        e.h0 = h0; e.h1 = h1; e.grid = G; e.block = T; e.ro = ro; e.tb = tb; e.stamp = ++pc.clock;
        pc.entries.push_back(std::move(e));
    }
}

inline int persistent_run(PersistentPlan &p, const Dev &D, cudaStream_t st, long long nsteps, bool check_stop, long long *launches) {
    if (!p.ok) return -1;
    const long long max_chunk = 0x3FFFFFFFll; // the flags and the divergence word are 32-bit step counts
    while (nsteps > 0) {
        long long n = nsteps < max_chunk ? nsteps : max_chunk;
        int cs = check_stop ? 1 : 0;
        Dev d = D;
        int tm = p.timing ? 1 : 0;
        PersistArgs a{p.ctl, p.flags, p.lk_slot, p.vx_id, p.vx_lane, p.deps, p.ndeps, p.pose_alt, p.ro};
        if (cudaMemsetAsync(p.ctl, 0, 2 * sizeof(unsigned int), st) != cudaSuccess) return -1;
        if (cudaMemsetAsync(p.flags, 0, (size_t)p.grid * 32 * sizeof(unsigned int), st) != cudaSuccess) return -1;
        void *args[] = {(void *)&d, (void *)&n, (void *)&cs, (void *)&tm, (void *)&a};
        if (cudaLaunchCooperativeKernel((const void *)k_persistent, dim3(p.grid), dim3(p.block), args, 0, st) != cudaSuccess) return -1;
        if (launches) (*launches)++;
        if (p.timing) {
            std::vector<unsigned long long> h(8 * (size_t)p.grid);
            cudaStreamSynchronize(st);
            cudaMemcpy(h.data(), p.ctl + 4, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            const char *nm[5] = {"wait", "link", "temp+sync", "voxel", "publish"};
            fprintf(stderr, "[persist timing] %llu steps, block %d, cycles/step over %d CTAs (min / mean / max):", h[5], p.block, p.grid);
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
    }
    return 0;
}

} // namespace vx3
