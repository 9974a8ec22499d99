// Fused step kernel for batches whose link topology is fixed (no collisions / attach / detach / voxel removal / signals).
//
// The two-pass streaming path (k_links -> k_voxels, vx3_kernels.cuh) materialises every link's end forces in HBM: 96 B
// written by the link pass and read again by the voxel pass, 192 of the ~340 B a link moves per step, none of it
// algorithmic.  Here the host cuts every body into compact spatial blocks of <= VX3_FUSE_BV voxels
// (recursive coordinate bisection of the lattice coordinates, per simulation) and STORES the batch in that order: a
// block's voxels are a contiguous index range, so are the links whose two ends lie in the block (INTERIOR links), and
// a simulation's remaining links (across a block face, or to ghost voxels of a neighbour slab) follow as one range.
// The ABI keeps the model's numbering: uploads and read-backs go through the permutation (vx3_engine.cu).
// A CTA processes a block: it evaluates the interior links, passes their end forces through shared memory and integrates
// the block's voxels right after — those forces never reach HBM, and every array is read and written in full coalesced
// runs.  The face links stay on the two-pass route: a small pre-pass (k_links<.., LIST>) evaluates them from the
// step-start poses and writes their end forces to the global force array, from which the voxel phase gathers them.
// Because an interior link only reads poses of its own block, everything stays in place: no second pose buffer, no
// cross-CTA ordering inside the launch.
//
// Per block:   stage   every voxel's pose record (64 B, read once) and link slots -> shared memory
//              links   interior links: state from HBM, end poses and materials from shared memory, new state back to
//                      HBM, end forces -> shared memory [dir][comp][voxel]
//              voxels  momenta / flags from HBM, pose from shared memory, forces from shared memory (interior) or the
//                      global force array (face links), summed in direction order 0..5 like VX3_Voxel::force()/moment()
// The arithmetic is the code of the streaming kernels (vx3_physics.cuh) on the same inputs: bit-identical results
// (tests/test_gpu_fused.py).  The end forces of interior links are only written to HBM by the WRITE_LF instantiation,
// which the engine uses for the last step of every stepping call, so that a state read-back sees them.
#pragma once
#include <atomic>
#include <thread>
#include <vector>

#include "vx3_kernels.cuh"

namespace vx3 {

#ifndef VX3_FUSE_BV
#define VX3_FUSE_BV 120 // voxels per block (max; <= 255: local indices are bytes)
#endif
#ifndef VX3_FUSE_T
#define VX3_FUSE_T 128
#endif
#ifndef VX3_FUSE_MIN_CTAS
#define VX3_FUSE_MIN_CTAS 4
#endif
#ifndef VX3_FUSE_PREFETCH
#define VX3_FUSE_PREFETCH 1
#endif
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#define VX3_FUSE_INTERIOR (-2) // shared-memory slot marker: the link in this direction is interior to the block

struct FusedArgs {
    const int4 *blk; // [nblocks] {first voxel, first interior link, links | voxels << 16, simulation}
    int nblocks;
};

struct FusedSmem {
    double F[36][VX3_FUSE_BV]; // end forces by receiving voxel: [direction * 6 + {Fx Fy Fz Mx My Mz}][local voxel]
    double P[8][VX3_FUSE_BV];  // pose records of the block's voxels, component-major
    int vl[6][VX3_FUSE_BV];    // link slot per direction: >= 0 face link (forces in HBM), -1 none, VX3_FUSE_INTERIOR
    int vmat[VX3_FUSE_BV];
    VoxMatC vm[VX3_SM_VMATS];
    LinkMatC lm[VX3_SM_LMATS];
    int div;
};

template <bool SMTAB, bool WRITE_LF> __global__ void __launch_bounds__(VX3_FUSE_T, VX3_FUSE_MIN_CTAS) k_fused(Dev D, FusedArgs A) {
    extern __shared__ __align__(16) unsigned char fused_smem_raw[];
    FusedSmem &sm = *reinterpret_cast<FusedSmem *>(fused_smem_raw);
    const int tid = threadIdx.x;
    if (SMTAB) {
        for (int i = tid; i < D.n_vmats * (int)(sizeof(VoxMatC) / 4); i += VX3_FUSE_T) reinterpret_cast<int *>(sm.vm)[i] = reinterpret_cast<const int *>(D.vmat_tab)[i];
        for (int i = tid; i < D.n_lmats * (int)(sizeof(LinkMatC) / 4); i += VX3_FUSE_T) reinterpret_cast<int *>(sm.lm)[i] = reinterpret_cast<const int *>(D.lmat_tab)[i];
    }
    if (tid == 0) sm.div = 0;
    int4 bd = (int)blockIdx.x < A.nblocks ? __ldg(A.blk + blockIdx.x) : make_int4(0, 0, 0, 0);
    for (int blk = blockIdx.x; blk < A.nblocks; blk += gridDim.x) {
        const int4 d = bd;
        if (blk + (int)gridDim.x < A.nblocks) {
            bd = __ldg(A.blk + blk + gridDim.x);
#if VX3_FUSE_PREFETCH
            // Every phase below starts with a burst of loads right after a CTA barrier, so its memory latency is exposed
            // (ncu: 4.8 long-scoreboard stall cycles per issue without this).  The NEXT block's data is two contiguous
            // index ranges: ask L2 for all of it now (~400 lines, 3 instructions per thread), a whole block ahead of its use.
            const int pv = bd.x, pl = bd.y, pnl = bd.z & 0xFFFF, pnv = (bd.z >> 16) & 0xFFFF;
            for (int k = 2 * tid; k < pnv; k += 2 * VX3_FUSE_T) prefetch_l2(D.pose + 8 * (size_t)(pv + k)); // 128-B line = 2 records
            for (int k = 4 * tid; k < pnv; k += 4 * VX3_FUSE_T) prefetch_l2(D.vlinks + 6 * (size_t)(pv + k));
            for (int k = 8 * tid; k < pnv + 8; k += 8 * VX3_FUSE_T) { // blocked planes: 8 double2 per line
                const int v = min(pv + k, pv + pnv - 1);
                prefetch_l2(D.mo(0, v)); prefetch_l2(D.mo(1, v)); prefetch_l2(D.mo(2, v));
                prefetch_l2(D.vc4 + v);
            }
            for (int k = 16 * tid; k < pnv + 16; k += 16 * VX3_FUSE_T) {
                const int v = min(pv + k, pv + pnv - 1);
                prefetch_l2(D.phase + v);
                prefetch_l2(D.vflags + v);
            }
            for (int i = 8 * tid; i < pnl + 8; i += 8 * VX3_FUSE_T) {
                const int g = min(pl + i, pl + pnl - 1);
                if (pnl > 0) {
                    prefetch_l2(D.lh(0, g)); prefetch_l2(D.lh(1, g)); prefetch_l2(D.lh(2, g)); prefetch_l2(D.lh(3, g)); prefetch_l2(D.lh(4, g));
                    prefetch_l2(D.lstrain + g);
                    prefetch_l2(D.lc4 + g);
                }
            }
            for (int i = 16 * tid; i < pnl + 16; i += 16 * VX3_FUSE_T) {
                const int g = min(pl + i, pl + pnl - 1);
                if (pnl > 0) {
                    prefetch_l2(D.larea + g);
                    prefetch_l2(D.lstate + g);
                }
            }
#endif
        }
        const int nl = d.z & 0xFFFF, nv = (d.z >> 16) & 0xFFFF, sim = d.w;
        // ---- the simulation's hot scalars (uniform over the block) ----
        const int4 *hp = reinterpret_cast<const int4 *>(D.simd + sim);
        const int4 hot0 = ldv(hp), hot1 = ldv(hp + 1), hot2 = ldv(hp + 2);
        const double t = __hiloint2double(hot0.y, hot0.x);
        const int status = hot0.z;
        const float dtF = __int_as_float(hot1.x);
        const int hot_flags = hot1.y;
        if (status != VX3_SIM_RUNNING || dtF == 0) continue; // uniform: the whole block sits out
        const int vstart = d.x, lstart = d.y;
        int4 le = tid < nl ? __ldg(D.lc4 + lstart + tid) : make_int4(0, 0, 0, 0);
        __syncthreads(); // the previous block's voxel phase is done with the shared arrays (first pass: the tables are in)
        // ================= stage: pose records and link slots of the block's voxels =================
        for (int k = tid; k < nv; k += VX3_FUSE_T) {
            const int v = vstart + k;
            const int4 ve = __ldg(D.vc4 + v); // {material, simulation, external, -}
            const double2 *ps = reinterpret_cast<const double2 *>(D.pose + 8 * (size_t)v);
            const double2 a = ldv(ps), b = ldv(ps + 1), c = ldv(ps + 2), e = ldv(ps + 3);
            const int2 *vs = reinterpret_cast<const int2 *>(D.vlinks + 6 * (size_t)v);
            const int2 l0 = vs[0], l1 = vs[1], l2 = vs[2];
            sm.P[0][k] = a.x; sm.P[1][k] = a.y; sm.P[2][k] = b.x; sm.P[3][k] = b.y;
            sm.P[4][k] = c.x; sm.P[5][k] = c.y; sm.P[6][k] = e.x; sm.P[7][k] = e.y;
            sm.vmat[k] = ve.x;
            // the block's interior links are the slot range [lstart, lstart + nl)
            sm.vl[0][k] = (unsigned)(l0.x - lstart) < (unsigned)nl ? VX3_FUSE_INTERIOR : l0.x;
            sm.vl[1][k] = (unsigned)(l0.y - lstart) < (unsigned)nl ? VX3_FUSE_INTERIOR : l0.y;
            sm.vl[2][k] = (unsigned)(l1.x - lstart) < (unsigned)nl ? VX3_FUSE_INTERIOR : l1.x;
            sm.vl[3][k] = (unsigned)(l1.y - lstart) < (unsigned)nl ? VX3_FUSE_INTERIOR : l1.y;
            sm.vl[4][k] = (unsigned)(l2.x - lstart) < (unsigned)nl ? VX3_FUSE_INTERIOR : l2.x;
            sm.vl[5][k] = (unsigned)(l2.y - lstart) < (unsigned)nl ? VX3_FUSE_INTERIOR : l2.y;
        }
        __syncthreads();
        // ================= interior links (gpu_update_links, VX3_VoxelyzeKernel.cu:566-581) =================
        for (int i = tid; i < nl; i += VX3_FUSE_T) {
            const int4 en = le; // {vneg, vpos, material, simulation}
            if (i + VX3_FUSE_T < nl) le = __ldg(D.lc4 + lstart + i + VX3_FUSE_T);
            const int gc = lstart + i, lmi = en.z, iN = en.x - vstart, iP = en.y - vstart;
            const double2 h0 = ldv(D.lh(0, gc)), h1 = ldv(D.lh(1, gc)), h2 = ldv(D.lh(2, gc)), h3 = ldv(D.lh(3, gc)), h4 = ldv(D.lh(4, gc));
            const float4 sn = ldv(D.lstrain + gc);
            const float2 ar = ldv(D.larea + gc);
            LinkRegs L;
            L.state = ldv(D.lstate + gc);
            L.pos2 = V3(h0.x, h0.y, h1.x);
            L.angle1v = V3(h1.y, h2.x, h2.y);
            L.angle2v = V3(h3.x, h3.y, h4.x);
            L.rest = h4.y;
            const V3 pN(sm.P[0][iN], sm.P[1][iN], sm.P[2][iN]), pP(sm.P[0][iP], sm.P[1][iP], sm.P[2][iP]);
            const Q4 qN(sm.P[3][iN], sm.P[4][iN], sm.P[5][iN], sm.P[6][iN]), qP(sm.P[3][iP], sm.P[4][iP], sm.P[5][iP], sm.P[6][iP]);
            const double tpN = sm.P[7][iN], tpP = sm.P[7][iP];
            const float tN = unpack_t(tpN), pdN = unpack_pd(tpN), tP = unpack_t(tpP), pdP = unpack_pd(tpP);
            const int axis = (L.state & LKS_AXIS_MASK) >> LKS_AXIS_SHIFT;
            struct { double size, on_after; float cte, dmn; int fixed; } mN, mP;
            {
                const int vmN = sm.vmat[iN], vmP = sm.vmat[iP];
                const VoxMatC &a = SMTAB ? sm.vm[vmN] : D.vmat_tab[vmN], &b = SMTAB ? sm.vm[vmP] : D.vmat_tab[vmP];
                mN.size = a.size[axis]; mN.on_after = a.thermal_on_after; mN.cte = a.alphaCTE; mN.dmn = a.dampMultNum; mN.fixed = a.fixed;
                mP.size = b.size[axis]; mP.on_after = b.thermal_on_after; mP.cte = b.alphaCTE; mP.dmn = b.dampMultNum; mP.fixed = b.fixed;
            }
            LinkOut o;
            o.forceNeg = o.momentNeg = o.forcePos = o.momentPos = V3(0, 0, 0);
            const bool live = !(L.state & (LKS_DETACHED | LKS_REMOVED)) && !(mN.fixed && mP.fixed);
            if (live) {
                L.state &= ~LKS_JUST_CREATED;
                L.strain = sn.x; L.maxStrain = sn.y; L.strainOffset = sn.z; L.stress = sn.w;
                L.area = ar.x; L.tsum = ar.y;
                if (hot_flags & SHF_THERMAL) { // updateRestLength() from either end's setTemperature (VX3_Voxel.cu:107-113)
                    const bool actN = !mN.fixed && !(mN.on_after > t), actP = !mP.fixed && !(mP.on_after > t);
                    if (actN || actP) L.rest = 0.5 * (mN.size * (1 + tN * mN.cte) + mP.size * (1 + tP * mP.cte)); // VX3_Voxel.h:95-98
                }
                const float dmN = mN.dmn / pdN, dmP = mP.dmn / pdP; // dampingMultiplier() (VX3_Voxel.h:206-208)
                LinkMid mid;
                link_stage_a(L, pN, qN, pP, qP, mid);
                if (!mid.small) link_stage_large(mid.pos2, mid.angle1, mid.angle2, mid.angle1v, L.rest);
                const LinkMatC &lm = SMTAB ? sm.lm[lmi] : D.lmat_tab[lmi];
                link_stage_c(L, mid, lm, D.strain_pool, D.stress_pool, dmN, dmP, o);
                *D.lh(0, gc) = make_double2(L.pos2.x, L.pos2.y);
                *D.lh(1, gc) = make_double2(L.pos2.z, L.angle1v.x);
                *D.lh(2, gc) = make_double2(L.angle1v.y, L.angle1v.z);
                *D.lh(3, gc) = make_double2(L.angle2v.x, L.angle2v.y);
                *D.lh(4, gc) = make_double2(L.angle2v.z, L.rest);
                D.lstrain[gc] = make_float4(L.strain, L.maxStrain, L.strainOffset, L.stress);
                D.lstate[gc] = L.state;
                if (WRITE_LF) {
                    *D.lf(0, gc) = make_double2(o.forceNeg.x, o.forceNeg.y);
                    *D.lf(1, gc) = make_double2(o.forceNeg.z, o.momentNeg.x);
                    *D.lf(2, gc) = make_double2(o.momentNeg.y, o.momentNeg.z);
                    *D.lf(3, gc) = make_double2(o.forcePos.x, o.forcePos.y);
                    *D.lf(4, gc) = make_double2(o.forcePos.z, o.momentPos.x);
                    *D.lf(5, gc) = make_double2(o.momentPos.y, o.momentPos.z);
                }
                if (L.strain > 100) { // divergence (every link is checked, see k_links)
                    D.simd[sim].diverged = 1;
                    sm.div = 1;
                }
            }
            // the negative end holds this link in its slot 2*axis, the positive end in 2*axis+1 (checked by the host plan)
            double *fn = &sm.F[12 * axis][iN], *fp = &sm.F[12 * axis + 6][iP];
            fn[0] = o.forceNeg.x; fn[VX3_FUSE_BV] = o.forceNeg.y; fn[2 * VX3_FUSE_BV] = o.forceNeg.z;
            fn[3 * VX3_FUSE_BV] = o.momentNeg.x; fn[4 * VX3_FUSE_BV] = o.momentNeg.y; fn[5 * VX3_FUSE_BV] = o.momentNeg.z;
            fp[0] = o.forcePos.x; fp[VX3_FUSE_BV] = o.forcePos.y; fp[2 * VX3_FUSE_BV] = o.forcePos.z;
            fp[3 * VX3_FUSE_BV] = o.momentPos.x; fp[4 * VX3_FUSE_BV] = o.momentPos.y; fp[5 * VX3_FUSE_BV] = o.momentPos.z;
        }
        __syncthreads();
        // ================= voxels (gpu_update_voxels, :582-623 -> VX3_Voxel::timeStep) =================
        // doTimeStep returns before the voxel pass when a link has diverged (:273-281): seen here for the block's own links and
        // for everything flagged before (the face pre-pass, blocks that ran earlier); see DESIGN.md §5 "defined behaviour"
        if (sm.div | ldv(&D.simd[sim].diverged)) {
            if (tid == 0) sm.div = 0; // (ordered before the next block's link phase by the two barriers in between)
            continue;
        }
        const double temp_amp = __hiloint2double(hot1.w, hot1.z), temp_period = __hiloint2double(hot2.y, hot2.x);
        for (int k = tid; k < nv; k += VX3_FUSE_T) {
            const int v = vstart + k;
            const int4 ve = __ldg(D.vc4 + v);
            const int vmi = ve.x, ext = ve.z;
            const int s0 = sm.vl[0][k], s1 = sm.vl[1][k], s2 = sm.vl[2][k], s3 = sm.vl[3][k], s4 = sm.vl[4][k], s5 = sm.vl[5][k];
            // ---- every global load of this voxel, all independent ----
            const double2 m0 = ldv(D.mo(0, v)), m1 = ldv(D.mo(1, v)), m2 = ldv(D.mo(2, v));
            const double phase = ldv(D.phase + v);
            VoxRegs r;
            r.flags = ldv(D.vflags + v);
            VX3_LOAD_END_FORCE(0, s0)
            VX3_LOAD_END_FORCE(1, s1)
            VX3_LOAD_END_FORCE(2, s2)
            VX3_LOAD_END_FORCE(3, s3)
            VX3_LOAD_END_FORCE(4, s4)
            VX3_LOAD_END_FORCE(5, s5)
            r.pos = V3(sm.P[0][k], sm.P[1][k], sm.P[2][k]);
            r.orient = Q4(sm.P[3][k], sm.P[4][k], sm.P[5][k], sm.P[6][k]);
            const double tp = sm.P[7][k];
            const float tempe = unpack_t(tp); // this step's temperature (gpu_update_temperature at time t)
            const float pd_old = unpack_pd(tp);
            r.linMom = V3(m0.x, m0.y, m1.x);
            r.angMom = V3(m1.y, m2.x, m2.y);
            V3 F(0, 0, 0), M(0, 0, 0); // force()/moment() sum the links in direction order 0..5 (VX3_Voxel.cu:350-397)
#define VX3_FUSE_ADD(dir, slot)                                                                                         \
    if (slot == VX3_FUSE_INTERIOR) {                                                                                    \
        F += V3(sm.F[6 * dir][k], sm.F[6 * dir + 1][k], sm.F[6 * dir + 2][k]);                                          \
        M += V3(sm.F[6 * dir + 3][k], sm.F[6 * dir + 4][k], sm.F[6 * dir + 5][k]);                                      \
    } else                                                                                                              \
        VX3_ADD_END_FORCE(dir, slot >= 0)
            VX3_FUSE_ADD(0, s0)
            VX3_FUSE_ADD(1, s1)
            VX3_FUSE_ADD(2, s2)
            VX3_FUSE_ADD(3, s3)
            VX3_FUSE_ADD(4, s4)
            VX3_FUSE_ADD(5, s5)
#undef VX3_FUSE_ADD
            const VoxMatC &m = SMTAB ? sm.vm[vmi] : D.vmat_tab[vmi];
            const double dt = dtF;
            D.tempe[v] = tempe;
            const double tnext = t + dtF; // temperature the next step will start with, see pack_tp
            float tempe_next = tempe;
            if ((hot_flags & SHF_THERMAL) && !(r.flags & VXF_REMOVED) && !(m.thermal_on_after > tnext) && !m.fixed)
                tempe_next = voxel_temperature(temp_amp, temp_period, (hot_flags & SHF_EXPANSION) != 0, tnext, phase);
            if (r.flags & VX3_VOX_GHOST) continue; // a neighbour slab owns this voxel: its pose record arrives with the halo exchange
            if ((r.flags & VXF_REMOVED) || m.fixed) {
                if (tempe_next != tempe) D.pose[8 * (size_t)v + 7] = pack_tp(tempe_next, pd_old);
                continue;
            }
            V3 cil(0, 0, 0);
            if ((hot_flags & SHF_CILIA) && !(r.flags & VX3_VOX_SURFACE) && m.cilia != 0 && !(m.cilia_on_after > t)) { // gpu_update_cilia_force :846-859
                V3 cf = load3(D.base_cilia, v);
                if (hot_flags & SHF_SIGNALS) cf += D.sig[6 * (size_t)v] * load3(D.shift_cilia, v);
                cil = r.orient.RotateVec3D(cf) * m.cilia;
            }
            V3 ff(0, 0, 0);
            const ExtC *px = ext >= 0 ? &D.exts[ext] : nullptr;
            const bool fixedAll = px && (px->dof & 0x3F) == 0x3F;
            if ((hot_flags & SHF_FORCE_FIELD) && !fixedAll) {
                const SimC &S = D.simc[sim];
                const SimD &dy = D.simd[sim];
                double vars[9];
                prog_vars(S, dy, t, r.pos.x, r.pos.y, r.pos.z, vars);
                ff.x = eval_slot(D, S, VX3_PROG_FORCE_X, vars, 0.0);
                ff.y = eval_slot(D, S, VX3_PROG_FORCE_Y, vars, 0.0);
                ff.z = eval_slot(D, S, VX3_PROG_FORCE_Z, vars, 0.0);
            }
            int ix = 0, iy = 0, iz = 0;
            if (px) {
                const short *ic = D.ixyz + 3 * (size_t)v;
                ix = ic[0]; iy = ic[1]; iz = ic[2];
            }
            voxel_time_step(r, m, px, ix, iy, iz, tempe, F, M, V3(0, 0, 0), cil, ff, dt);
            if (hot_flags & SHF_ATTACH_COND) { // enableAttach = AND of the five attach conditions at the new position (:609-621)
                const SimC &S = D.simc[sim];
                const SimD &dy = D.simd[sim];
                double vars[9];
                prog_vars(S, dy, t, r.pos.x, r.pos.y, r.pos.z, vars);
                bool all = true;
                for (int c = 0; c < 5 && all; c++) all = eval_slot(D, S, VX3_PROG_ATTACH_0 + c, vars, 1.0) > 0;
                if (all) r.flags |= VXF_ENABLE_ATTACH;
                else r.flags &= ~VXF_ENABLE_ATTACH;
            }
            store_pose(D.pose, v, r.pos, r.orient, tempe_next, dtF);
            *D.mo(0, v) = make_double2(r.linMom.x, r.linMom.y);
            *D.mo(1, v) = make_double2(r.linMom.z, r.angMom.x);
            *D.mo(2, v) = make_double2(r.angMom.y, r.angMom.z);
            D.vflags[v] = r.flags;
        }
    }
}

// ------------------------------------------------------------------ host plan
struct FusedPlan {
    bool ok = false;
    int nblocks = 0, ninterior = 0, nface = 0;
    int grid = 1, face_tiles = 0, face_grid = 1;
    size_t smem = 0;
    bool smtab = true;
    // arena slices (placed by vx3_engine.cu)
    int4 *blk = nullptr;
    int *face_slot = nullptr; // links of the pre-pass (per simulation one contiguous slot range)
    int4 *face_c4 = nullptr;  // their {vneg, vpos, material, simulation} records (copy of lc4)
};

// Storage order of a batch: external (model) index -> device index, per simulation a permutation of its own range.
struct FusedLayout {
    std::vector<int> vperm, lperm; // [nvox], [nslots]
    std::vector<int4> blk;
    std::vector<int2> face_range;  // per simulation {first face slot, count}
    long long ninterior = 0, nface = 0;
};

template <class Fn> inline void fuse_parallel(size_t ntasks, int nthreads, Fn fn) {
    if (nthreads <= 1 || ntasks <= 1) {
        for (size_t i = 0; i < ntasks; i++) fn(i);
        return;
    }
    std::atomic<size_t> next{0};
    auto w = [&]() {
        for (size_t k; (k = next.fetch_add(1)) < ntasks;) fn(k);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads && (size_t)t < ntasks; t++) th.emplace_back(w);
    w();
    for (auto &x : th) x.join();
}

// recursive coordinate bisection of the voxels idx[lo, hi) (indices local to the simulation) into blocks [b0, b0 + nb).
// The x extent is discounted: voxels are numbered x-fastest, so blocks that are long in x keep longer runs of the model's order.
inline void fuse_rcb(int *idx, int lo, int hi, int b0, int nb, const int16_t *ix, const int16_t *iy, const int16_t *iz, int *blk_of, int par_depth) {
    if (nb == 1) {
        for (int i = lo; i < hi; i++) blk_of[idx[i]] = b0;
        return;
    }
    const int16_t *const co[3] = {ix, iy, iz};
    int mn[3] = {1 << 30, 1 << 30, 1 << 30}, mx[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
    for (int i = lo; i < hi; i++)
        for (int a = 0; a < 3; a++) {
            const int c = co[a][idx[i]];
            mn[a] = std::min(mn[a], c);
            mx[a] = std::max(mx[a], c);
        }
    const int ext[3] = {2 * (mx[0] - mn[0] + 1), 3 * (mx[1] - mn[1] + 1), 3 * (mx[2] - mn[2] + 1)};
    int ax = 2;
    if (ext[1] > ext[ax]) ax = 1;
    if (ext[0] > ext[ax]) ax = 0;
    const int16_t *const ca = co[ax];
    const int nbl = nb / 2;
    const int mid = lo + (int)((long long)(hi - lo) * nbl / nb);
    std::nth_element(idx + lo, idx + mid, idx + hi, [&](int p, int q) { // (coordinate, index): a strict total order
        const int cp = ca[p], cq = ca[q];
        return cp != cq ? cp < cq : p < q;
    });
    if (par_depth > 0 && hi - lo > 16384) {
        std::thread left([&]() { fuse_rcb(idx, lo, mid, b0, nbl, ix, iy, iz, blk_of, par_depth - 1); });
        fuse_rcb(idx, mid, hi, b0 + nbl, nb - nbl, ix, iy, iz, blk_of, par_depth - 1);
        left.join();
    } else {
        fuse_rcb(idx, lo, mid, b0, nbl, ix, iy, iz, blk_of, 0);
        fuse_rcb(idx, mid, hi, b0 + nbl, nb - nbl, ix, iy, iz, blk_of, 0);
    }
}

// Cuts every simulation into blocks and derives the storage order.  Pure host code on the models (also reachable
// through vx3_fused_plan_check, the CPU test hook).  Returns false when a model does not fit the block model — its
// adjacency is not the lattice form (a link sits in slot 2*axis of its negative end and 2*axis+1 of its positive end) —
// and the batch then keeps the model's order and the two-pass kernels.
inline bool fused_layout_build(const vx3_model_desc *models, int n, const std::vector<SimC> &simc, int bv, FusedLayout &L) {
    if (bv < 1 || bv > 255) return false;
    size_t nvox = 0, nslots = 0;
    for (int s = 0; s < n; s++) {
        if (simc[s].lcap != models[s].n_links) return false; // spare pool slots: dynamic topology
        nvox += simc[s].nvox;
        nslots += simc[s].lcap;
    }
    if (nvox == 0) return false;
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    const int nthreads = nvox + nslots > 100000 ? std::min(hw, 16) : 1;
    const int inner = n == 1 ? nthreads : 1; // one big body: parallel inside; many bodies: parallel over them
    std::vector<int> bbase(n + 1, 0);
    const int target = bv > 8 ? bv - 2 : bv; // the proportional splits may exceed the average by a voxel or two
    for (int s = 0; s < n; s++) bbase[s + 1] = bbase[s] + (simc[s].nvox + target - 1) / target;
    const int nb = bbase[n];
    L.vperm.assign(nvox, 0);
    L.lperm.assign(nslots, 0);
    L.blk.assign(nb, make_int4(0, 0, 0, 0));
    L.face_range.assign(n, make_int2(0, 0));
    std::atomic<int> bad{0};
    std::vector<long long> nint_s(n, 0), nface_s(n, 0);
    fuse_parallel(n, n == 1 ? 1 : nthreads, [&](size_t s) {
        const vx3_model_desc &m = models[s];
        const SimC &S = simc[s];
        const int V = S.nvox, NL = m.n_links, nbs = bbase[s + 1] - bbase[s];
        if (V == 0) return;
        // ---- adjacency must be the lattice form, both ways ----
        std::atomic<int> lbad{0};
        fuse_parallel((size_t)(V + 65535) / 65536, inner, [&](size_t c) {
            for (int v = (int)c * 65536; v < std::min(V, ((int)c + 1) * 65536); v++)
                for (int d = 0; d < 6; d++) {
                    const int l = m.vox_links[6 * (size_t)v + d];
                    if (l < 0) continue;
                    if (l >= NL || m.link_axis[l] != d / 2 || ((d & 1) ? m.link_vpos[l] : m.link_vneg[l]) != v) lbad = 1;
                }
        });
        fuse_parallel((size_t)(NL + 65535) / 65536, inner, [&](size_t c) {
            for (int l = (int)c * 65536; l < std::min(NL, ((int)c + 1) * 65536); l++) {
                const int a = m.link_axis[l], vn = m.link_vneg[l], vp = m.link_vpos[l];
                if (a < 0 || a > 2 || vn < 0 || vn >= V || vp < 0 || vp >= V || m.vox_links[6 * (size_t)vn + 2 * a] != l || m.vox_links[6 * (size_t)vp + 2 * a + 1] != l) lbad = 1;
            }
        });
        if (lbad) {
            bad = 1;
            return;
        }
        // ---- blocks ----
        std::vector<int> idx(V), blk_of(V), vcount(nbs, 0), vstart(nbs + 1, 0);
        for (int i = 0; i < V; i++) idx[i] = i;
        fuse_rcb(idx.data(), 0, V, 0, nbs, m.ix, m.iy, m.iz, blk_of.data(), inner > 1 ? 4 : 0);
        for (int i = 0; i < V; i++) vcount[blk_of[i]]++;
        for (int b = 0; b < nbs; b++) {
            if (vcount[b] > bv) bad = 1;
            vstart[b + 1] = vstart[b] + vcount[b];
        }
        if (bad) return;
        std::vector<int> order(V), cur(vstart.begin(), vstart.end() - 1); // device position -> model voxel, ascending model index within a block
        for (int i = 0; i < V; i++) {
            const int at = cur[blk_of[i]]++;
            order[at] = i;
            L.vperm[(size_t)S.voff + i] = S.voff + at;
        }
        // ---- links: every link is listed once, at its positive end (slot 2*axis+1), block by block ----
        std::vector<int> nint(nbs, 0), nfac(nbs, 0);
        const size_t nbch = (size_t)(nbs + 63) / 64;
        fuse_parallel(nbch, inner, [&](size_t c) {
            for (int b = (int)c * 64; b < std::min(nbs, ((int)c + 1) * 64); b++)
                for (int k = vstart[b]; k < vstart[b + 1]; k++)
                    for (int a = 0; a < 3; a++) {
                        const int l = m.vox_links[6 * (size_t)order[k] + 2 * a + 1];
                        if (l < 0) continue;
                        if (blk_of[m.link_vneg[l]] == b) nint[b]++;
                        else nfac[b]++;
                    }
        });
        std::vector<int> lstart(nbs + 1, 0), fstart(nbs + 1, 0);
        for (int b = 0; b < nbs; b++) {
            lstart[b + 1] = lstart[b] + nint[b];
            fstart[b + 1] = fstart[b] + nfac[b];
        }
        if (lstart[nbs] + fstart[nbs] != NL) { // (cannot happen after the adjacency check)
            bad = 1;
            return;
        }
        nint_s[s] = lstart[nbs];
        nface_s[s] = fstart[nbs];
        L.face_range[s] = make_int2(S.loff + lstart[nbs], fstart[nbs]);
        fuse_parallel(nbch, inner, [&](size_t c) {
            for (int b = (int)c * 64; b < std::min(nbs, ((int)c + 1) * 64); b++) {
                int ci = S.loff + lstart[b], cf = S.loff + lstart[nbs] + fstart[b];
                for (int k = vstart[b]; k < vstart[b + 1]; k++)
                    for (int a = 0; a < 3; a++) {
                        const int l = m.vox_links[6 * (size_t)order[k] + 2 * a + 1];
                        if (l < 0) continue;
                        L.lperm[(size_t)S.loff + l] = blk_of[m.link_vneg[l]] == b ? ci++ : cf++;
                    }
                if (nint[b] > 0xFFFF) bad = 1;
                L.blk[bbase[s] + b] = make_int4(S.voff + vstart[b], S.loff + lstart[b], nint[b] | (vcount[b] << 16), (int)s);
            }
        });
    });
    if (bad) return false;
    for (int s = 0; s < n; s++) {
        L.ninterior += nint_s[s];
        L.nface += nface_s[s];
    }
    return true;
}

} // namespace vx3
