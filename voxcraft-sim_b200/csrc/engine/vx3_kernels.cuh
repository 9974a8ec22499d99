// Streaming step kernels: one launch per phase over the concatenated SoA arrays of the whole batch.
// Phase order per step = VX3_VoxelyzeKernel::doTimeStep (src/VX3/VX3_VoxelyzeKernel.cu:237-359):
//   k_links -> [k_grid_build, k_contact] -> [k_resolve_detach] -> k_voxels -> [k_signals] -> [k_secondary]
//   -> [k_com_partial] -> k_tail
#pragma once
#include "vx3_physics.cuh"

namespace vx3 {

#define VX3_BLOCK 256
// Tile sizes (threads per CTA = items per tile) and occupancy targets of the two streaming hot kernels; tuned on
// B200, see DESIGN.md §4.  Both are persistent tile loops that carry the next item's indices in registers (below).
#ifndef VX3_LINK_T
#define VX3_LINK_T 128
#endif
#ifndef VX3_LINKS_MIN_CTAS
#define VX3_LINKS_MIN_CTAS 4
#endif
#ifndef VX3_VOX_T
// First statement of every kernel that may be launched programmatically (launch_pdl, vx3_engine.cu): wait until the preceding kernel
// of the stream has completed and its writes are visible, then allow the NEXT kernel's CTAs to be scheduled behind this grid's.
// Both are no-ops in a normal launch.
#define VX3_PDL_ENTRY()                                              \
    asm volatile("griddepcontrol.wait;" ::: "memory");               \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define VX3_VOX_T 128
#endif
#ifndef VX3_VOXELS_MIN_CTAS
#define VX3_VOXELS_MIN_CTAS 4
#endif

// The 8th double of the pose record carries two floats: the voxel's temperature AT THE CURRENT SIMULATION TIME (what
// gpu_update_temperature will set at the start of the next step — the voxel pass computes it once per voxel, the link
// pass reads it with the pose instead of evaluating sin() for both ends of every link) and the voxel's previousDt
// (VX3_Voxel.h:206-208 dampingMultiplier), so the link pass needs no second gather for it.
__device__ __forceinline__ double pack_tp(float tempe, float prevdt) { return __hiloint2double(__float_as_int(prevdt), __float_as_int(tempe)); }
__device__ __forceinline__ float unpack_t(double d) { return __int_as_float(__double2loint(d)); }
__device__ __forceinline__ float unpack_pd(double d) { return __int_as_float(__double2hiint(d)); }

__device__ __forceinline__ void load_pose(const double *__restrict__ pose, int v, V3 &p, Q4 &q) {
    const double2 *s = reinterpret_cast<const double2 *>(pose + 8 * (size_t)v);
    const double2 a = s[0], b = s[1], c = s[2], d = s[3];
    p = V3(a.x, a.y, b.x);
    q = Q4(b.y, c.x, c.y, d.x);
}
__device__ __forceinline__ V3 load_pos(const double *__restrict__ pose, int v) {
    const double2 *s = reinterpret_cast<const double2 *>(pose + 8 * (size_t)v);
    const double2 a = s[0];
    return V3(a.x, a.y, s[1].x);
}
// 256-bit accesses (sm_100: LDG/STG.E.ENL2.256) for the 64-byte pose record — two requests instead of four, in the link
// pass's end-pose gathers and the voxel pass's load and store: -2.8 % step time on config 3, -2 % on config 5 (A/B builds).
#ifndef VX3_POSE256
#define VX3_POSE256 1
#endif
__device__ __forceinline__ void st4(double *p, double a, double b, double c, double d);
__device__ __forceinline__ void store_pose(double *pose, int v, const V3 &p, const Q4 &q, float tempe_next, float prevdt) {
#if VX3_POSE256
    st4(pose + 8 * (size_t)v, p.x, p.y, p.z, q.w);
    st4(pose + 8 * (size_t)v + 4, q.x, q.y, q.z, pack_tp(tempe_next, prevdt));
    return;
#endif
    double2 *s = reinterpret_cast<double2 *>(pose + 8 * (size_t)v);
    s[0] = make_double2(p.x, p.y);
    s[1] = make_double2(p.z, q.w);
    s[2] = make_double2(q.x, q.y);
    s[3] = make_double2(q.z, pack_tp(tempe_next, prevdt));
}
__device__ __forceinline__ V3 load3(const double *__restrict__ a, size_t i) { return V3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
__device__ __forceinline__ void store3(double *a, size_t i, const V3 &v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }
__device__ __forceinline__ V3 load_linmom(const Dev &D, int v) {
    const double2 a = *D.mo(0, v);
    return V3(a.x, a.y, D.mo(1, v)->x);
}

// ------------------------------------------------------------------ links
// gpu_update_links (VX3_VoxelyzeKernel.cu:566-581) with the temperature-driven rest-length refresh of
// gpu_update_temperature (:625-650) folded in.
//
// What the pass is bound by was established by A/B builds on config 3 (DESIGN.md §4 has the table): its bytes
// (~250 MB per step there) move at HBM speed when the arithmetic is removed, so what is left is to overlap ~900-1500
// instructions per link with them at 16 warps per SM.  The structure that measured best:
//  * one memory level per link: a persistent tile loop in which every thread carries the constant index record
//    {ends, material, simulation} of its NEXT link in registers, so that every load of the current link — slot record,
//    both end poses, the simulation's hot scalars — is independent of every other.  They are issued back to back with
//    volatile loads: ptxas otherwise sinks each load behind the early-out tests next to its first use and turns one
//    exposed round trip per link into three dependent ones;
//  * blocked SoA records (idx_lh / idx_lf): coalesced 16-byte accesses, full-sector writes;
//  * the material tables live in shared memory, so the per-link constants cost an LDS, not a dependent global load.
// Tried and dropped, each measured slower: cp.async staging of the next links' records (the LSU write wavefronts of the
// gathers throttle the MIO queue), L1/L2 software prefetch, a class-sorted processing order that makes warps uniform in
// small/large-angle mode (−34 % instructions, but the indirection costs more in sector efficiency than it saves), and
// storing end forces by receiving voxel (streams for the voxel pass, but leaves partially written sectors -> ECC RMW), and
// (round 2) software pipelining in registers — every load of link i+1 issued at the top of iteration i, the index record two
// links ahead: 168 registers, 3 CTAs per SM instead of 4, and 20-26 % slower (config 3 110 vs 92 us per step, config 5 link pass
// 942 vs 746 us): the pass is bound by dependent fp64 issue at register-limited occupancy, and warps, not hidden loads, are
// what it is short of.
#define VX3_SM_VMATS 32
#define VX3_SM_LMATS 64
struct LinkSmem {
    VoxMatL vm[VX3_SM_VMATS];
    LinkMatC lm[VX3_SM_LMATS];
};
// Loads the compiler must not sink below a branch: the streaming kernels issue every load of an item back to back and
// only then look at the values (nvcc otherwise moves each load next to its first use, behind the early-out tests, which
// turns one memory round trip per item into three dependent ones — see profiles/).
#define VX3_LDQ ".volatile"
__device__ __forceinline__ double2 ldv(const double2 *p) {
    double2 r;
    asm volatile("ld" VX3_LDQ ".global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldv(const float4 *p) {
    float4 r;
    asm volatile("ld" VX3_LDQ ".global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int4 ldv(const int4 *p) {
    int4 r;
    asm volatile("ld" VX3_LDQ ".global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldv(const float2 *p) {
    float2 r;
    asm volatile("ld" VX3_LDQ ".global.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldv(const int *p) {
    int r;
    asm volatile("ld" VX3_LDQ ".global.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ double ldv(const double *p) {
    double r;
    asm volatile("ld" VX3_LDQ ".global.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
// predicated 16-byte load (keeps the zero it is given when the predicate is off)
__device__ __forceinline__ double2 ldv_if(const double2 *p, bool on) {
    double2 r = make_double2(0.0, 0.0);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\t@p ld" VX3_LDQ ".global.v2.f64 {%0, %1}, [%2];\n\t}" : "+d"(r.x), "+d"(r.y) : "l"(p), "r"((int)on));
    return r;
}

// Gathers of records another kernel wrote (end poses in the link pass, end forces in the voxel pass): the qualifier is a
// build-time choice so that L1-cached variants can be measured against the pinned volatile form (scripts/build_variants.sh).
#ifndef VX3_GATHER_LD
#define VX3_GATHER_LD "ld.volatile.global"
#endif
__device__ __forceinline__ double2 ldgat(const double2 *p) {
    double2 r;
    asm volatile(VX3_GATHER_LD ".v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldgat_if(const double2 *p, bool on) {
    double2 r = make_double2(0.0, 0.0);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\t@p " VX3_GATHER_LD ".v2.f64 {%0, %1}, [%2];\n\t}" : "+d"(r.x), "+d"(r.y) : "l"(p), "r"((int)on));
    return r;
}

struct double4v { double x, y, z, w; };
__device__ __forceinline__ double4v ldgat4(const double *p) {
    double4v r;
    asm volatile(VX3_GATHER_LD ".v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st4(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__device__ __forceinline__ int4 link_c4(const Dev &D, long long g) { return g < D.nlinkslots ? __ldg(D.lc4 + g) : make_int4(-1, -1, 0, 0); }

// The large-angle branch of orientLink costs ~540 instructions on top of ~780, and in an actuated body a few percent of
// the links are in that regime — spread so that most warps hold one or two of them and execute the whole branch for 2
// active lanes (profiles/r01_sass_regions_k_links_c3.txt: 35 % of this kernel's warp instructions ran with <= 3 lanes).
// k_links runs the branch in place; k_links_deferred (below) defers those links to dense passes.  Which is faster depends
// on the batch, so the engine times both on a batch's first streaming steps and keeps the faster (vx3_engine.cu,
// launch_links).  (A third variant, a CTA-wide shared-memory queue drained by one warp between two barriers, was never
// the fastest on any workload and was removed.)

// SMTAB: the batch's material tables fit the shared-memory copies (the normal case); otherwise they are read from global
// LIST: the pass runs over a list of link slots (list_slot[i], with their index records list_c4[i]) instead of all slots —
// the pre-pass of the fused step (vx3_fused.cuh), which evaluates only the links across block faces.
template <bool SMTAB, bool LIST = false>
__global__ void __launch_bounds__(VX3_LINK_T, VX3_LINKS_MIN_CTAS) k_links(Dev D, int ntiles, const int *__restrict__ list_slot = nullptr, const int4 *__restrict__ list_c4 = nullptr, int nlist = 0, int tile0 = 0) {
    VX3_PDL_ENTRY();
    __shared__ LinkSmem sm;
    const int tid = threadIdx.x;
    const long long G = gridDim.x;
    auto fetch = [&](long long i, int &slot) -> int4 { // index record and slot of item i (empty past the end)
        if (LIST) {
            if (i < nlist) {
                slot = __ldg(list_slot + i);
                return __ldg(list_c4 + i);
            }
            slot = 0;
            return make_int4(-1, -1, 0, 0);
        }
        slot = (int)i;
        return link_c4(D, i);
    };
    if (SMTAB) {
        for (int i = tid; i < D.n_vmats * (int)(sizeof(VoxMatL) / 4); i += VX3_LINK_T) reinterpret_cast<int *>(sm.vm)[i] = reinterpret_cast<const int *>(D.vmatl_tab)[i];
        for (int i = tid; i < D.n_lmats * (int)(sizeof(LinkMatC) / 4); i += VX3_LINK_T) reinterpret_cast<int *>(sm.lm)[i] = reinterpret_cast<const int *>(D.lmat_tab)[i];
    }
    __syncthreads();
    long long tile = (long long)tile0 + blockIdx.x; // tiles [tile0, ntiles): the whole array, or one of the two ranges of a slab batch (vx3_halo.cuh)
    int gnext;
    int4 c4 = fetch(tile * VX3_LINK_T + tid, gnext);
    for (; tile < ntiles; tile += G) {
        // ---- the next item's constant indices (consumed by the next iteration) ----
        const int gc = gnext;
        const int4 c4n = fetch((tile + G) * VX3_LINK_T + tid, gnext);
        const int4 c = c4;
        c4 = c4n;
        bool live = c.x >= 0; // empty pool slot / past the end
        LinkRegs L;
        LinkMid mid;
        float dmN = 0, dmP = 0;
        int state0 = 0, hotf = 0;
        mid.small = true;
        if (!live) continue;
        {
            // ---- every load of this link, all independent ----
            const double2 h0 = ldv(D.lh(0, gc)), h1 = ldv(D.lh(1, gc)), h2 = ldv(D.lh(2, gc)), h3 = ldv(D.lh(3, gc)), h4 = ldv(D.lh(4, gc));
            const float4 sn = ldv(D.lstrain + gc);
            const float2 ar = ldv(D.larea + gc);
            L.state = state0 = ldv(D.lstate + gc);
            const double2 *pa = reinterpret_cast<const double2 *>(D.pose + 8 * (size_t)c.x), *pb = reinterpret_cast<const double2 *>(D.pose + 8 * (size_t)c.y);
#if VX3_POSE256
            const double4v A0 = ldgat4(reinterpret_cast<const double *>(pa)), A1 = ldgat4(reinterpret_cast<const double *>(pa) + 4);
            const double4v B0 = ldgat4(reinterpret_cast<const double *>(pb)), B1 = ldgat4(reinterpret_cast<const double *>(pb) + 4);
            const double2 a0 = make_double2(A0.x, A0.y), a1 = make_double2(A0.z, A0.w), a2 = make_double2(A1.x, A1.y), a3 = make_double2(A1.z, A1.w);
            const double2 b0 = make_double2(B0.x, B0.y), b1 = make_double2(B0.z, B0.w), b2 = make_double2(B1.x, B1.y), b3 = make_double2(B1.z, B1.w);
#else
            const double2 a0 = ldgat(pa), a1 = ldgat(pa + 1), a2 = ldgat(pa + 2), a3 = ldgat(pa + 3);
            const double2 b0 = ldgat(pb), b1 = ldgat(pb + 1), b2 = ldgat(pb + 2), b3 = ldgat(pb + 3);
#endif
            const int vmN = ldv(D.vmat + c.x), vmP = ldv(D.vmat + c.y);
            const int4 *hp = reinterpret_cast<const int4 *>(D.simd + c.w);
            const int4 hot0 = ldv(hp), hot1 = ldv(hp + 1);
            L.pos2 = V3(h0.x, h0.y, h1.x);
            L.angle1v = V3(h1.y, h2.x, h2.y);
            L.angle2v = V3(h3.x, h3.y, h4.x);
            L.rest = h4.y;
            const V3 pN(a0.x, a0.y, a1.x), pP(b0.x, b0.y, b1.x);
            const Q4 qN(a1.y, a2.x, a2.y, a3.x), qP(b1.y, b2.x, b2.y, b3.x);
            const float tN = unpack_t(a3.y), pdN = unpack_pd(a3.y), tP = unpack_t(b3.y), pdP = unpack_pd(b3.y);
            const double t = __hiloint2double(hot0.y, hot0.x);
            const int status = hot0.z;
            const float dt = __int_as_float(hot1.x);
            const int hot_flags = hot1.y;
            hotf = hot_flags;
            const int axis = (L.state & LKS_AXIS_MASK) >> LKS_AXIS_SHIFT;
            // the few end-material values, read up front so that their latencies overlap
            struct { double size, on_after; float cte, dmn; int fixed; } mN, mP;
            {
                const VoxMatL &a = SMTAB ? sm.vm[vmN] : D.vmatl_tab[vmN], &b = SMTAB ? sm.vm[vmP] : D.vmatl_tab[vmP];
                mN.size = a.size[axis]; mN.on_after = a.thermal_on_after; mN.cte = a.alphaCTE; mN.dmn = a.dampMultNum; mN.fixed = a.fixed;
                mP.size = b.size[axis]; mP.on_after = b.thermal_on_after; mP.cte = b.alphaCTE; mP.dmn = b.dampMultNum; mP.fixed = b.fixed;
            }
            if (L.state & (LKS_DETACHED | LKS_REMOVED)) live = false;
            if (status != VX3_SIM_RUNNING || dt == 0 || (mN.fixed && mP.fixed)) live = false;
            if (!live) continue;
            {
                L.state &= ~LKS_JUST_CREATED;
                L.strain = sn.x; L.maxStrain = sn.y; L.strainOffset = sn.z; L.stress = sn.w;
                L.area = ar.x; L.tsum = ar.y;
                if (hot_flags & SHF_THERMAL) { // updateRestLength() from either end's setTemperature (VX3_Voxel.cu:107-113)
                    const bool actN = !mN.fixed && !(mN.on_after > t), actP = !mP.fixed && !(mP.on_after > t);
                    if (actN || actP) L.rest = 0.5 * (mN.size * (1 + tN * mN.cte) + mP.size * (1 + tP * mP.cte)); // VX3_Voxel.h:95-98
                }
                // dampingMultiplier() = 2*_sqrtMass*zetaInternal/previousDt (float)
                dmN = mN.dmn / pdN;
                dmP = mP.dmn / pdP;
                link_stage_a(L, pN, qN, pP, qP, mid);
            }
        }
        if (!mid.small) link_stage_large(mid.pos2, mid.angle1, mid.angle2, mid.angle1v, L.rest);
        const LinkMatC &lm = SMTAB ? sm.lm[c.z] : D.lmat_tab[c.z];
        LinkOut o;
        link_stage_c(L, mid, lm, D.strain_pool, D.stress_pool, dmN, dmP, o);
        *D.lh(0, gc) = make_double2(L.pos2.x, L.pos2.y);
        *D.lh(1, gc) = make_double2(L.pos2.z, L.angle1v.x);
        *D.lh(2, gc) = make_double2(L.angle1v.y, L.angle1v.z);
        *D.lh(3, gc) = make_double2(L.angle2v.x, L.angle2v.y);
        *D.lh(4, gc) = make_double2(L.angle2v.z, L.rest);
        D.lstrain[gc] = make_float4(L.strain, L.maxStrain, L.strainOffset, L.stress);
        // EnableDetach: a link past its failure strain goes on its simulation's failed-link list; k_resolve_detach takes it off the
        // voxels after this step's attach phase (gpu_update_detach, VX3_VoxelyzeKernel.cu:946-968, runs after updateAttach)
        if ((hotf & SHF_DETACH) && mat_failed(lm, L.maxStrain) && !(L.state & LKS_FAILED)) {
            const SimC &Sd = D.simc[c.w];
            const int k = atomicAdd(&D.simd[c.w].fail_count, 1);
            if (k < Sd.fail_cap) {
                D.fail_list[Sd.fail_off + k] = gc;
                L.state |= LKS_FAILED;
            }
        }
        if (L.state != state0) D.lstate[gc] = L.state; // (regime / velocity-valid / new-link bits rarely change)
        *D.lf(0, gc) = make_double2(o.forceNeg.x, o.forceNeg.y);
        *D.lf(1, gc) = make_double2(o.forceNeg.z, o.momentNeg.x);
        *D.lf(2, gc) = make_double2(o.momentNeg.y, o.momentNeg.z);
        *D.lf(3, gc) = make_double2(o.forcePos.x, o.forcePos.y);
        *D.lf(4, gc) = make_double2(o.forcePos.z, o.momentPos.x);
        *D.lf(5, gc) = make_double2(o.momentPos.y, o.momentPos.z);
        // divergence: the reference samples one random link per step (:273-280); every link is checked here
        if (L.strain > 100) D.simd[c.w].diverged = 1;
    }
}


// ---- halo exchange of a slab batch inside the link pass (vx3_halo.cuh has the protocol) ----
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed_sys(unsigned int *p, unsigned int v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_vol_u32(const unsigned int *p) { return *reinterpret_cast<const volatile unsigned int *>(p); }
__device__ __forceinline__ void halo_fail(SimD *simd, const HaloIn &h) { // a neighbour did not send in time: mark the batch failed and freeze it
    *h.err = 1;
    simd->err = VX3_ERR_CUDA;
    simd->dt = 0.0f; // doTimeStep(0) does nothing (VX3_VoxelyzeKernel.cu:240-241)
    __threadfence();
}
// A slab's link pass reads the poses of its ghost voxels straight from the receive buffers the neighbours wrote (no copy into the
// ghost voxels' pose rows, no kernel of its own): link tiles [0, face_tile0) touch no ghost; before its first later tile a warp
// makes sure the neighbours' send number is in.  ONE warp of the grid — the first to get there — waits on the neighbours' flags
// (system scope) and tells the others through a flag in local memory.  Returns the send number whose buffers are to be read
// (0: nothing has been sent yet, the ghost rows still hold the model's initial poses).
__device__ __noinline__ unsigned int halo_arrival_wait(const HaloIn *hin, SimD *simd, int lane) {
    const HaloIn &h = *hin;
    const unsigned int step1 = ld_vol_u32(h.seq), collected = ld_vol_u32(h.seq + 1);
    if (step1 == 0 || collected == step1) return step1; // (collected by k_halo_recv already: the buffers are complete)
    const int parity = (int)((step1 - 1u) & 1u);
    unsigned int *claim = h.state, *arrived = h.state + 32; // both hold send numbers and only grow
    if (lane == 0) {
        const long long t0 = clock64();
        if (ld_vol_u32(arrived) < step1) {
            if (atomicMax(claim, step1) < step1) { // mine to wait for
                for (int sd = 0; sd < 2; sd++) {
                    if (h.n[sd] == 0) continue;
                    while (ld_acquire_sys(h.flag[sd] + 32 * parity) < step1)
                        if (clock64() - t0 > h.spin_cycles) {
                            halo_fail(simd, h);
                            break;
                        }
                }
                __threadfence();
                *reinterpret_cast<volatile unsigned int *>(arrived) = step1;
            } else {
                while (ld_vol_u32(arrived) < step1)
                    if (clock64() - t0 > 2 * h.spin_cycles) break;
            }
        }
        __threadfence();
    }
    __syncwarp();
    return step1;
}
// where a face link's end pose is read: the voxel's own row, or its record in the receive buffer of send number step1
__device__ __forceinline__ const double2 *halo_pose_row(const HaloIn *hin, const double *pose, int v, unsigned int step1) {
    const int r = step1 ? __ldg(hin->ghost_row + v) : -1;
    if (r < 0) return reinterpret_cast<const double2 *>(pose + 8 * (size_t)v);
    const int sd = r >> 30, slot = r & 0x3FFFFFFF;
    return reinterpret_cast<const double2 *>(hin->buf[sd] + ((size_t)((step1 - 1u) & 1u) * hin->n[sd] + slot) * 8);
}

// Second variant of the link pass: a lane whose link turns out to need the large-angle branch
// does NOT process it — it pushes the link's slot number onto its warp's private queue (shared memory, warp-aggregated
// push) and idles for the rest of the iteration, so the warp runs the small-angle path only (~780 instructions instead
// of ~1200).  Whenever a warp's queue holds 32 entries (and once more at the end) the warp spends one iteration on a
// DENSE pass: every lane takes one deferred link, reloads its inputs (nothing has been stored for it yet, so they are
// unchanged) and runs the complete update, large-angle branch included, with all lanes busy.  The gathers of a dense pass
// are uncoalesced, but only the few percent of deferred links pay for that.  Same arithmetic on the same inputs: bit-identical
// to k_links.  No CTA barrier anywhere.
// HALO (slab batches): links of tiles >= HaloIn::face_tile0 may have a ghost end, whose pose comes from the receive buffers.
template <bool SMTAB, bool HALO = false> __global__ void __launch_bounds__(VX3_LINK_T, VX3_LINKS_MIN_CTAS) k_links_deferred(Dev D, int ntiles, int tile0 = 0) {
    VX3_PDL_ENTRY();
    __shared__ LinkSmem sm;
    __shared__ int sDef[VX3_LINK_T / 32][64];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long G = gridDim.x;
    if (SMTAB) {
        for (int i = tid; i < D.n_vmats * (int)(sizeof(VoxMatL) / 4); i += VX3_LINK_T) reinterpret_cast<int *>(sm.vm)[i] = reinterpret_cast<const int *>(D.vmatl_tab)[i];
        for (int i = tid; i < D.n_lmats * (int)(sizeof(LinkMatC) / 4); i += VX3_LINK_T) reinterpret_cast<int *>(sm.lm)[i] = reinterpret_cast<const int *>(D.lmat_tab)[i];
    }
    __syncthreads();
    long long tile = (long long)tile0 + blockIdx.x;
    const long long F = HALO ? D.hin->face_tile0 : 0;
    bool ghosts_in = false;
    unsigned int halo_step1 = 0; // the send number whose receive buffers hold the ghost poses (0: none yet)
    int4 c4 = link_c4(D, tile * VX3_LINK_T + tid);
    int nd = 0; // entries in this warp's queue (warp-uniform)
    for (;;) {
        const bool tiles_left = tile < ntiles; // uniform over the CTA
        bool dense;
        int gc;
        int4 c;
        if (nd >= 32 || (!tiles_left && nd > 0)) { // ---- dense pass over deferred links ----
            dense = true;
            const int take = nd < 32 ? nd : 32;
            gc = lane < take ? sDef[wid][nd - take + lane] : -1;
            nd -= take;
            c = gc >= 0 ? __ldg(D.lc4 + gc) : make_int4(-1, -1, 0, 0);
        } else if (tiles_left) { // ---- the next tile; the item after it is prefetched into registers ----
            dense = false;
            const int4 c4n = link_c4(D, (tile + G) * VX3_LINK_T + tid);
            if (HALO && tile >= F && !ghosts_in) {
                halo_step1 = halo_arrival_wait(D.hin, D.simd, lane);
                ghosts_in = true;
            }
            gc = (int)(tile * VX3_LINK_T + tid);
            c = c4;
            c4 = c4n;
            tile += G;
        } else
            break;
        bool live = c.x >= 0; // empty pool slot / past the end
        LinkRegs L;
        LinkMid mid;
        float dmN = 0, dmP = 0;
        int state0 = 0, hotf = 0;
        mid.small = true;
        if (live) {
            // ---- every load of this link, all independent ----
            const double2 h0 = ldv(D.lh(0, gc)), h1 = ldv(D.lh(1, gc)), h2 = ldv(D.lh(2, gc)), h3 = ldv(D.lh(3, gc)), h4 = ldv(D.lh(4, gc));
            const float4 sn = ldv(D.lstrain + gc);
            const float2 ar = ldv(D.larea + gc);
            L.state = state0 = ldv(D.lstate + gc);
            const double2 *pa = reinterpret_cast<const double2 *>(D.pose + 8 * (size_t)c.x), *pb = reinterpret_cast<const double2 *>(D.pose + 8 * (size_t)c.y);
            if (HALO && ghosts_in) { // a ghost end's pose is in the neighbour's receive buffer
                pa = halo_pose_row(D.hin, D.pose, c.x, halo_step1);
                pb = halo_pose_row(D.hin, D.pose, c.y, halo_step1);
            }
#if VX3_POSE256
            const double4v A0 = ldgat4(reinterpret_cast<const double *>(pa)), A1 = ldgat4(reinterpret_cast<const double *>(pa) + 4);
            const double4v B0 = ldgat4(reinterpret_cast<const double *>(pb)), B1 = ldgat4(reinterpret_cast<const double *>(pb) + 4);
            const double2 a0 = make_double2(A0.x, A0.y), a1 = make_double2(A0.z, A0.w), a2 = make_double2(A1.x, A1.y), a3 = make_double2(A1.z, A1.w);
            const double2 b0 = make_double2(B0.x, B0.y), b1 = make_double2(B0.z, B0.w), b2 = make_double2(B1.x, B1.y), b3 = make_double2(B1.z, B1.w);
#else
            const double2 a0 = ldgat(pa), a1 = ldgat(pa + 1), a2 = ldgat(pa + 2), a3 = ldgat(pa + 3);
            const double2 b0 = ldgat(pb), b1 = ldgat(pb + 1), b2 = ldgat(pb + 2), b3 = ldgat(pb + 3);
#endif
            const int vmN = ldv(D.vmat + c.x), vmP = ldv(D.vmat + c.y);
            const int4 *hp = reinterpret_cast<const int4 *>(D.simd + c.w);
            const int4 hot0 = ldv(hp), hot1 = ldv(hp + 1);
            L.pos2 = V3(h0.x, h0.y, h1.x);
            L.angle1v = V3(h1.y, h2.x, h2.y);
            L.angle2v = V3(h3.x, h3.y, h4.x);
            L.rest = h4.y;
            const V3 pN(a0.x, a0.y, a1.x), pP(b0.x, b0.y, b1.x);
            const Q4 qN(a1.y, a2.x, a2.y, a3.x), qP(b1.y, b2.x, b2.y, b3.x);
            const float tN = unpack_t(a3.y), pdN = unpack_pd(a3.y), tP = unpack_t(b3.y), pdP = unpack_pd(b3.y);
            const double t = __hiloint2double(hot0.y, hot0.x);
            const int status = hot0.z;
            const float dt = __int_as_float(hot1.x);
            const int hot_flags = hot1.y;
            hotf = hot_flags;
            const int axis = (L.state & LKS_AXIS_MASK) >> LKS_AXIS_SHIFT;
            struct { double size, on_after; float cte, dmn; int fixed; } mN, mP;
            {
                const VoxMatL &a = SMTAB ? sm.vm[vmN] : D.vmatl_tab[vmN], &b = SMTAB ? sm.vm[vmP] : D.vmatl_tab[vmP];
                mN.size = a.size[axis]; mN.on_after = a.thermal_on_after; mN.cte = a.alphaCTE; mN.dmn = a.dampMultNum; mN.fixed = a.fixed;
                mP.size = b.size[axis]; mP.on_after = b.thermal_on_after; mP.cte = b.alphaCTE; mP.dmn = b.dampMultNum; mP.fixed = b.fixed;
            }
            if (L.state & (LKS_DETACHED | LKS_REMOVED)) live = false;
            if (status != VX3_SIM_RUNNING || dt == 0 || (mN.fixed && mP.fixed)) live = false;
            if (live) {
                L.state &= ~LKS_JUST_CREATED;
                L.strain = sn.x; L.maxStrain = sn.y; L.strainOffset = sn.z; L.stress = sn.w;
                L.area = ar.x; L.tsum = ar.y;
                if (hot_flags & SHF_THERMAL) { // updateRestLength() from either end's setTemperature (VX3_Voxel.cu:107-113)
                    const bool actN = !mN.fixed && !(mN.on_after > t), actP = !mP.fixed && !(mP.on_after > t);
                    if (actN || actP) L.rest = 0.5 * (mN.size * (1 + tN * mN.cte) + mP.size * (1 + tP * mP.cte)); // VX3_Voxel.h:95-98
                }
                dmN = mN.dmn / pdN; // dampingMultiplier() = 2*_sqrtMass*zetaInternal/previousDt (float)
                dmP = mP.dmn / pdP;
                link_stage_a(L, pN, qN, pP, qP, mid);
            }
        }
        // ---- defer the links that need the large-angle branch (not in a dense pass: there they run it) ----
        const bool defer = live && !dense && !mid.small;
        const unsigned dm = __ballot_sync(0xFFFFFFFFu, defer);
        if (defer) sDef[wid][nd + __popc(dm & ((1u << lane) - 1u))] = gc;
        nd += __popc(dm);
        __syncwarp();
        if (!live || defer) continue;
        if (!mid.small) link_stage_large(mid.pos2, mid.angle1, mid.angle2, mid.angle1v, L.rest);
        const LinkMatC &lm = SMTAB ? sm.lm[c.z] : D.lmat_tab[c.z];
        LinkOut o;
        link_stage_c(L, mid, lm, D.strain_pool, D.stress_pool, dmN, dmP, o);
        *D.lh(0, gc) = make_double2(L.pos2.x, L.pos2.y);
        *D.lh(1, gc) = make_double2(L.pos2.z, L.angle1v.x);
        *D.lh(2, gc) = make_double2(L.angle1v.y, L.angle1v.z);
        *D.lh(3, gc) = make_double2(L.angle2v.x, L.angle2v.y);
        *D.lh(4, gc) = make_double2(L.angle2v.z, L.rest);
        D.lstrain[gc] = make_float4(L.strain, L.maxStrain, L.strainOffset, L.stress);
        // EnableDetach: a link past its failure strain goes on its simulation's failed-link list; k_resolve_detach takes it off the
        // voxels after this step's attach phase (gpu_update_detach, VX3_VoxelyzeKernel.cu:946-968, runs after updateAttach)
        if ((hotf & SHF_DETACH) && mat_failed(lm, L.maxStrain) && !(L.state & LKS_FAILED)) {
            const SimC &Sd = D.simc[c.w];
            const int k = atomicAdd(&D.simd[c.w].fail_count, 1);
            if (k < Sd.fail_cap) {
                D.fail_list[Sd.fail_off + k] = gc;
                L.state |= LKS_FAILED;
            }
        }
        if (L.state != state0) D.lstate[gc] = L.state; // (regime / velocity-valid / new-link bits rarely change)
        *D.lf(0, gc) = make_double2(o.forceNeg.x, o.forceNeg.y);
        *D.lf(1, gc) = make_double2(o.forceNeg.z, o.momentNeg.x);
        *D.lf(2, gc) = make_double2(o.momentNeg.y, o.momentNeg.z);
        *D.lf(3, gc) = make_double2(o.forcePos.x, o.forcePos.y);
        *D.lf(4, gc) = make_double2(o.forcePos.z, o.momentPos.x);
        *D.lf(5, gc) = make_double2(o.momentPos.y, o.momentPos.z);
        if (L.strain > 100) D.simd[c.w].diverged = 1;
    }
}

// ------------------------------------------------------------------ voxels
__device__ __forceinline__ void prog_vars(const SimC &S, const SimD &dy, double t, double x, double y, double z, double *vars) {
    vars[0] = x; vars[1] = y; vars[2] = z;
    vars[3] = dy.collision_count; vars[4] = t; vars[5] = dy.recent_angle; vars[6] = dy.target_closeness;
    vars[7] = dy.num_close_pairs; vars[8] = S.nvox;
}
__device__ __forceinline__ double eval_slot(const Dev &D, const SimC &S, int slot, const double *vars, double dflt) {
    if (S.prog_n[slot] <= 0) return dflt; // "tag absent": defined result (vx3_abi.h, vx3_program)
    bool ok;
    // the evaluator keeps one value per token (VX3_MathTree.h:52): programs of up to 128 tokens — all the reference's demos — run in
    // a 1 KB frame, longer ones (the reference allows 1024, VX3_VoxelyzeKernel.cuh:113-119) in the 8 KB one
    if (S.prog_n[slot] <= VX3_DEV_MAX_TOKENS) return mt_eval<VX3_DEV_MAX_TOKENS>(D.tokens + S.prog_off[slot], S.prog_n[slot], vars, &ok);
    return mt_eval<VX3_MAX_TOKENS>(D.tokens + S.prog_off[slot], S.prog_n[slot], vars, &ok);
}

// gpu_update_voxels (VX3_VoxelyzeKernel.cu:582-623) -> VX3_Voxel::timeStep
//
// Same idea as the link pass: a persistent tile loop in which the voxel's six link slots and constant indices
// {material, simulation, external} are fetched one item ahead into registers, so that the pose, momenta, hot scalars of
// the simulation and the end forces of exactly those directions whose slot holds a link are all issued together.
struct VoxSmem {
    VoxMatC vm[VX3_SM_VMATS];
};
struct VoxIdx {
    int2 l0, l1, l2; // vlinks[0..5]
    int4 c4;         // vmat, sim, ext
};
__device__ __forceinline__ VoxIdx vox_idx(const Dev &D, long long v) {
    VoxIdx r;
    r.l0 = r.l1 = r.l2 = make_int2(-1, -1);
    r.c4 = make_int4(0, 0, -1, 0);
    if (v < D.nvox) {
        const int2 *src = reinterpret_cast<const int2 *>(D.vlinks + 6 * (size_t)v);
        r.l0 = src[0]; r.l1 = src[1]; r.l2 = src[2];
        r.c4 = __ldg(D.vc4 + v);
    }
    return r;
}
// the voxel is the negative end of the links in its even (+axis) slots and the positive end of those in its odd slots
#define VX3_LOAD_END_FORCE(dir, slot)                                                                                   \
    const double2 fa##dir = ldgat_if(D.lf(3 * (dir & 1), slot), slot >= 0), fb##dir = ldgat_if(D.lf(3 * (dir & 1) + 1, slot), slot >= 0),                \
                  fc##dir = ldgat_if(D.lf(3 * (dir & 1) + 2, slot), slot >= 0);
#define VX3_ADD_END_FORCE(dir, cond)                                                                                    \
    if (cond) {                                                                                                         \
        F += V3(fa##dir.x, fa##dir.y, fb##dir.x);                                                                       \
        M += V3(fb##dir.y, fc##dir.x, fc##dir.y);                                                                       \
    }

__device__ void tail_light(const Dev &D, int sim, int check_stop);
// HALO (slab batches, vx3_halo.cuh): a face voxel's record also goes into the neighbour's receive buffer, and the last CTA to finish
// publishes the send number to both neighbours and, with tail >= 0, does the end-of-step bookkeeping (tail_light, check_stop = tail).
template <bool SMTAB, bool HALO = false> __global__ void __launch_bounds__(VX3_VOX_T, VX3_VOXELS_MIN_CTAS) k_voxels(Dev D, int ntiles, int tail = -1) {
    VX3_PDL_ENTRY();
    __shared__ VoxSmem sm;
    const int tid = threadIdx.x;
    const long long G = gridDim.x;
    // this pass is send number seq[0] + 1; every CTA reads the count before the last one to finish moves it
    const unsigned int send_parity = HALO ? (ld_vol_u32(D.hout->seq) & 1u) : 0u;
    auto peer_row = [&](int code) -> double * { // code = vc4.w - 1
        const HaloOut &o = *D.hout;
        const int sd = code >> 30, slot = code & 0x3FFFFFFF;
        return o.buf[sd] + ((size_t)send_parity * o.n[sd] + slot) * 8;
    };
    bool sent = false; // this thread has stored into a neighbour's buffer
    auto send_row_as_is = [&](int v, int code) { // a face voxel this pass does not move (fixed, removed, frozen simulation): its row as it stands
        sent = true;
        const double2 *src = reinterpret_cast<const double2 *>(D.pose + 8 * (size_t)v);
        double2 *dst = reinterpret_cast<double2 *>(peer_row(code));
#pragma unroll
        for (int k = 0; k < 4; k++) dst[k] = ldgat(src + k);
    };
    if (SMTAB) { // material table -> shared memory
        for (int i = tid; i < D.n_vmats * (int)(sizeof(VoxMatC) / 4); i += VX3_VOX_T) reinterpret_cast<int *>(sm.vm)[i] = reinterpret_cast<const int *>(D.vmat_tab)[i];
        __syncthreads();
    }
    long long tile = blockIdx.x;
    VoxIdx nx = vox_idx(D, tile * VX3_VOX_T + tid);
    for (; tile < ntiles; tile += G) {
        const VoxIdx id = nx;
        nx = vox_idx(D, (tile + G) * VX3_VOX_T + tid);
        const long long vv = tile * VX3_VOX_T + tid;
        if (vv >= D.nvox) continue;
        const int v = (int)vv;
        // ---- every load of this voxel, all independent ----
        const double2 *ps = reinterpret_cast<const double2 *>(D.pose + 8 * (size_t)v);
#if VX3_POSE256
        const double4v P0 = ldgat4(reinterpret_cast<const double *>(ps)), P1 = ldgat4(reinterpret_cast<const double *>(ps) + 4);
        const double2 a = make_double2(P0.x, P0.y), b = make_double2(P0.z, P0.w), c = make_double2(P1.x, P1.y), d = make_double2(P1.z, P1.w);
#else
        const double2 a = ldv(ps), b = ldv(ps + 1), c = ldv(ps + 2), d = ldv(ps + 3);
#endif
        const double2 m0 = ldv(D.mo(0, v)), m1 = ldv(D.mo(1, v)), m2 = ldv(D.mo(2, v));
        const int4 *hp = reinterpret_cast<const int4 *>(D.simd + id.c4.y);
        const int4 hot0 = ldv(hp), hot1 = ldv(hp + 1), hot2 = ldv(hp + 2);
        const double phase = ldv(D.phase + v);
        VoxRegs r;
        r.flags = ldv(D.vflags + v);
        const int flags0 = r.flags;
        VX3_LOAD_END_FORCE(0, id.l0.x)
        VX3_LOAD_END_FORCE(1, id.l0.y)
        VX3_LOAD_END_FORCE(2, id.l1.x)
        VX3_LOAD_END_FORCE(3, id.l1.y)
        VX3_LOAD_END_FORCE(4, id.l2.x)
        VX3_LOAD_END_FORCE(5, id.l2.y)
        r.pos = V3(a.x, a.y, b.x);
        r.orient = Q4(b.y, c.x, c.y, d.x);
        const float tempe = unpack_t(d.y); // this step's temperature (gpu_update_temperature at time t)
        const float pd_old = unpack_pd(d.y);
        r.linMom = V3(m0.x, m0.y, m1.x);
        r.angMom = V3(m1.y, m2.x, m2.y);
        const int vmi = id.c4.x, sim = id.c4.y, ext = id.c4.z;
        V3 F(0, 0, 0), M(0, 0, 0); // force()/moment() sum the links in direction order 0..5 (VX3_Voxel.cu:350-397)
        VX3_ADD_END_FORCE(0, id.l0.x >= 0)
        VX3_ADD_END_FORCE(1, id.l0.y >= 0)
        VX3_ADD_END_FORCE(2, id.l1.x >= 0)
        VX3_ADD_END_FORCE(3, id.l1.y >= 0)
        VX3_ADD_END_FORCE(4, id.l2.x >= 0)
        VX3_ADD_END_FORCE(5, id.l2.y >= 0)
        const double t = __hiloint2double(hot0.y, hot0.x);
        const int status = hot0.z, sdiverged = hot0.w;
        const float dtF = __int_as_float(hot1.x);
        const int hot_flags = hot1.y;
        const double temp_amp = __hiloint2double(hot1.w, hot1.z), temp_period = __hiloint2double(hot2.y, hot2.x);
        const VoxMatC &m = SMTAB ? sm.vm[vmi] : D.vmat_tab[vmi];
        if (status != VX3_SIM_RUNNING || sdiverged || dtF == 0) {
            if (HALO && id.c4.w) send_row_as_is(v, id.c4.w - 1);
            continue;
        }
        const double dt = dtF;
        D.tempe[v] = tempe;
        // temperature the next step will start with (time t + dt), see pack_tp
        const double tnext = t + dtF;
        float tempe_next = tempe;
        if ((hot_flags & SHF_THERMAL) && !(r.flags & VXF_REMOVED) && !(m.thermal_on_after > tnext) && !m.fixed)
            tempe_next = voxel_temperature(temp_amp, temp_period, (hot_flags & SHF_EXPANSION) != 0, tnext, phase);
        if (r.flags & VX3_VOX_GHOST) continue; // a neighbour slab owns this voxel: its pose record arrives with the halo exchange
        if ((r.flags & VXF_REMOVED) || m.fixed) {
            if (tempe_next != tempe) D.pose[8 * (size_t)v + 7] = pack_tp(tempe_next, pd_old);
            if (HALO && id.c4.w) send_row_as_is(v, id.c4.w - 1);
            continue;
        }
        V3 contact(0, 0, 0);
        if (D.contact) {
            contact = load3(D.contact, v);
            store3(D.contact, v, V3());
        }
        V3 cil(0, 0, 0);
        if ((hot_flags & SHF_CILIA) && !(r.flags & VX3_VOX_SURFACE) && m.cilia != 0 && !(m.cilia_on_after > t)) { // gpu_update_cilia_force :846-859
            V3 cf = load3(D.base_cilia, v);
            if (hot_flags & SHF_SIGNALS) cf += D.sig[6 * (size_t)v] * load3(D.shift_cilia, v); // baseCiliaForce + localSignal * shiftCiliaForce
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
        voxel_time_step(r, m, px, ix, iy, iz, tempe, F, M, contact, cil, ff, dt);
        // enableAttach = AND of the five attach conditions at the new position (:609-621)
        if (hot_flags & SHF_ATTACH_COND) {
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
        if (HALO && id.c4.w) {
            sent = true;
            store_pose(peer_row(id.c4.w - 1), 0, r.pos, r.orient, tempe_next, dtF);
        }
        *D.mo(0, v) = make_double2(r.linMom.x, r.linMom.y);
        *D.mo(1, v) = make_double2(r.linMom.z, r.angMom.x);
        *D.mo(2, v) = make_double2(r.angMom.y, r.angMom.z);
        if (r.flags != flags0) D.vflags[v] = r.flags; // (the friction / attach bits rarely change)
    }
    if (HALO) {
        // Every sending thread fences its peer stores (system scope) before its CTA counts itself in, so when the last CTA sees the full
        // count every record is visible to the neighbour: it needs one more fence (cumulativity; nothing of its own is outstanding,
        // so no NVLink round trip) and two plain system-scope stores — not two store-releases, each of which waits a round trip.
        // The bookkeeping of the step runs beside it in another warp.
        __shared__ int sLast;
        const HaloOut &o = *D.hout;
        if (sent) __threadfence_system();
        __syncthreads();
        if (tid == 0) sLast = atomicAdd(o.count, 1u) == (unsigned int)G - 1;
        __syncthreads();
        if (sLast) { // every CTA's records are out, every CTA has read seq and simd
            if (tid == 0) {
                *o.count = 0u;
                const unsigned int step1 = ld_vol_u32(o.seq) + 1u;
                __threadfence_system();
                for (int sd = 0; sd < 2; sd++)
                    if (o.n[sd] > 0) st_relaxed_sys(o.flag[sd] + 32 * (int)send_parity, step1);
                *reinterpret_cast<volatile unsigned int *>(o.seq) = step1;
            } else if (tid == 32 && tail >= 0)
                tail_light(D, 0, tail);
        }
    } else if (tail >= 0) {
        // plain steps of any other batch: the last CTA to finish does what k_tail_light would do in a launch of its own (every CTA
        // has consumed the simulations' scalars before it counts itself in; nothing between the voxel pass and the end of the step
        // reads them in the batches this is used for — no signals, no voxel removal)
        __shared__ int sLastT;
        __syncthreads();
        if (tid == 0) sLastT = atomicAdd(D.vox_count, 1u) == (unsigned int)G - 1;
        __syncthreads();
        if (sLastT) {
            if (tid == 0) *D.vox_count = 0u;
            for (int sim = tid; sim < D.nsims; sim += VX3_VOX_T) tail_light(D, sim, tail);
        }
    }
}

// ------------------------------------------------------------------ collision grid
__device__ __forceinline__ unsigned cell_hash(int sim, int cx, int cy, int cz) {
    return ((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u) ^ ((unsigned)cz * 83492791u) ^ ((unsigned)sim * 2654435761u);
}
__device__ __forceinline__ bool sim_collides(const SimC &S) { return S.enable_collision || S.enable_attach; }

// regenerateSurfaceVoxels (:495-513) + uniform-grid insert (replaces the O(S^2) sweep of gpu_update_attach :833-843).
// The grid is a hash table of buckets with VX3_CELL_SLOTS inline slots each (one 32-byte sector): a surface voxel takes
// slot atomicAdd(cell_cnt[bucket]) — no counting pass, no scan, no second pass; a voxel that finds the inline slots taken
// (many voxels squeezed into one cell, or two cells in one bucket) chains itself onto the bucket's overflow list.  The
// order inside a bucket is whatever the atomics produce — the contact phase sorts each voxel's partners by index before
// it accumulates, so the result does not depend on it.  cell_cnt / cell_ovf are zeroed by a memset node
// before this kernel.  Also writes each voxel's ContactRec and publishes this step's temperature (updateTemperature :219-235).
__global__ void __launch_bounds__(VX3_BLOCK) k_grid_build(Dev D) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= D.nvox) return;
    const int sim = D.vsim[v];
    const SimC &S = D.simc[sim];
    const SimD &dy = D.simd[sim];
    ContactRec rec;
    rec.cx = rec.cy = rec.cz = 0;
    rec.bucket = -1;
    rec.px = rec.py = rec.pz = rec.bs = 0;
    rec.sim = sim;
    rec.mat = 0;
    rec.fixed = 0;
    rec._pad = 0;
    // (the surface flags follow the links in EVERY running simulation of the batch, also one that has collisions off but loses
    // links to detach / removal while its batch mates collide; only the grid is for colliding simulations)
    if (dy.status == VX3_SIM_RUNNING && !dy.diverged && dy.dt != 0) {
        int flags = D.vflags[v];
        bool interior = true; // VX3_Voxel::updateSurface (VX3_Voxel.cu:515-524): the bit named SURFACE means interior
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const int li = D.vlinks[6 * (size_t)v + i];
            if (li < 0 || (D.lstate[li] & LKS_DETACHED)) interior = false;
        }
        const int nf = interior ? (flags | VX3_VOX_SURFACE) : (flags & ~VX3_VOX_SURFACE);
        if (nf != flags) D.vflags[v] = nf;
        const float tempe = sim_collides(S) ? unpack_t(D.pose[8 * (size_t)v + 7]) : 0.0f;
        if (sim_collides(S)) D.tempe[v] = tempe; // this step's temperature, for the contact phase
        if (sim_collides(S) && !interior && !(flags & VXF_REMOVED)) {
            const V3 p = load_pos(D.pose, v);
            const int mat = D.vmat[v];
            const VoxMatC &m = D.vmat_tab[mat];
            rec.cx = (int)floor(p.x * S.cell_inv);
            rec.cy = (int)floor(p.y * S.cell_inv);
            rec.cz = (int)floor(p.z * S.cell_inv);
            rec.bucket = (int)(cell_hash(sim, rec.cx, rec.cy, rec.cz) & (unsigned)D.hmask);
            rec.px = p.x; rec.py = p.y; rec.pz = p.z;
            rec.bs = base_size_average(m, tempe);
            rec.mat = mat;
            rec.fixed = m.fixed;
            const int slot = atomicAdd(&D.cell_cnt[rec.bucket], 1);
            if (slot < VX3_CELL_SLOTS) {
                CellItem it;
                it.x = (float)p.x; it.y = (float)p.y; it.z = (float)p.z;
                it.v = v;
                *reinterpret_cast<float4 *>(D.cell_items + VX3_CELL_SLOTS * (size_t)rec.bucket + slot) = make_float4(it.x, it.y, it.z, __int_as_float(it.v));
            } else
                D.cell_next[v] = atomicExch(&D.cell_ovf[rec.bucket], v + 1) - 1;
        }
    }
    D.crec[v] = rec;
}

// surface flags only (simulations without collisions but with detach): regenerateSurfaceVoxels (:495-513)
__global__ void __launch_bounds__(VX3_BLOCK) k_surface(Dev D) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= D.nvox) return;
    const SimD &dy = D.simd[D.vsim[v]];
    if (dy.status != VX3_SIM_RUNNING || dy.diverged || dy.dt == 0) return;
    const int flags = D.vflags[v];
    bool interior = true;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const int li = D.vlinks[6 * (size_t)v + i];
        if (li < 0 || (D.lstate[li] & LKS_DETACHED)) interior = false;
    }
    const int nf = interior ? (flags | VX3_VOX_SURFACE) : (flags & ~VX3_VOX_SURFACE);
    if (nf != flags) D.vflags[v] = nf;
}

// Connectivity filter in front of the depth-5 neighbour search.  uf[] is a union-find forest over the voxels: the host
// initialises it with the connected components of the model's link graph, k_resolve unions the two trees whenever it
// accepts an attach, and nothing ever splits a tree (detach / removal only remove links).  So the forest's components are
// a superset of the real ones: different roots => no path at all => not within 5 links, exactly; same root => ask the
// search.  Bodies of a pile that have not stuck together skip the search (up to ~1000 dependent path steps) entirely.
__device__ __forceinline__ int uf_find(const Dev &D, int x) {
    int p = D.uf[x];
    while (p != x) {
        x = p;
        p = D.uf[x];
    }
    return x;
}
__device__ __forceinline__ bool uf_disconnected(const Dev &D, int a, int b) { return D.uf && uf_find(D, a) != uf_find(D, b); }

// is_neighbor (VX3_VoxelyzeKernel.cu:651-680), iterative
__device__ bool is_neighbor_dfs(const Dev &D, int v1, int v2, int depth);
// The reference walks every non-reversing path of up to 5 links from voxel1 depth-first: true iff the two voxels are at most
// 5 links apart in the link graph (a shortest path never reverses).  For a pair that is NOT — two bodies of a pile glued
// together somewhere else, touching here: the common case once a pile has settled — that walk is ~4700 path steps of two
// dependent loads each, in one lane, every step (config 4 after 5000 steps: k_contact 503 us instead of 37).  Same answer from
// both ends: the ball of radius 2 around v2 (<= 37 voxels) against the ball of radius 3 around v1, on the adjacency table
// D.vnb (one 32-byte row per voxel; the voxels' link slots are kept symmetric by attach, detach and removal): 45 rows in
// 9 rounds of independent loads.  Without the table (no simulation of the batch attaches) the path walk stays.
__device__ __forceinline__ void load_nb_row(const Dev &D, int x, int r[6]) {
    const int4 a = *reinterpret_cast<const int4 *>(D.vnb + 8 * (size_t)x);
    const int2 c = *reinterpret_cast<const int2 *>(D.vnb + 8 * (size_t)x + 4);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = c.x; r[5] = c.y;
}
__device__ __forceinline__ unsigned long long nb_bloom(int x) { return 1ull << (((unsigned)x * 0x9E3779B1u) >> 26); }
__device__ __noinline__ bool within_five_links(const Dev &D, int v1, int v2) {
    if (v1 == v2) return true;
    int B[44], nB = 0; // v2's ball of radius 2 (with repeats; <= 37 while the slots are symmetric)
    int lvl[6], nl = 0;
    int r[6][6];
    load_nb_row(D, v2, r[0]);
    B[nB++] = v2;
#pragma unroll
    for (int i = 0; i < 6; i++)
        if (r[0][i] >= 0) {
            if (r[0][i] == v1) return true;
            lvl[nl++] = r[0][i];
            B[nB++] = r[0][i];
        }
#pragma unroll
    for (int k = 0; k < 6; k++)
        if (k < nl) load_nb_row(D, lvl[k], r[k]);
#pragma unroll
    for (int k = 0; k < 6; k++)
        if (k < nl)
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const int x = r[k][i];
                if (x >= 0 && x != v2) {
                    if (x == v1) return true;
                    B[nB++] = x;
                }
            }
    unsigned long long bloom = 0;
    for (int j = 0; j < nB; j++) bloom |= nb_bloom(B[j]);
    auto in_ball = [&](int x) {
        if (!(bloom & nb_bloom(x))) return false;
        for (int j = 0; j < nB; j++)
            if (B[j] == x) return true;
        return false;
    };
    // v1's side, level by level: distance 3, 4, 5 in total
    load_nb_row(D, v1, r[0]);
    nl = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
        if (r[0][i] >= 0) {
            if (in_ball(r[0][i])) return true;
            lvl[nl++] = r[0][i];
        }
    int A2[36], n2 = 0;
#pragma unroll
    for (int k = 0; k < 6; k++)
        if (k < nl) load_nb_row(D, lvl[k], r[k]);
#pragma unroll
    for (int k = 0; k < 6; k++)
        if (k < nl)
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const int x = r[k][i];
                if (x >= 0 && x != v1) {
                    if (in_ball(x)) return true;
                    A2[n2++] = x;
                }
            }
    for (int k0 = 0; k0 < n2; k0 += 6) {
#pragma unroll
        for (int k = 0; k < 6; k++)
            if (k0 + k < n2) load_nb_row(D, A2[k0 + k], r[k]);
#pragma unroll
        for (int k = 0; k < 6; k++)
            if (k0 + k < n2)
#pragma unroll
                for (int i = 0; i < 6; i++)
                    if (r[k][i] >= 0 && in_ball(r[k][i])) return true;
    }
    return false;
}
// The answer does not depend on the order the paths are tried in; the reference's depth-first order can walk hundreds of
// 5-link paths before it tries the 2-link one that ends the search, so the short paths are tried first (<= 36 nodes).
__device__ bool is_neighbor(const Dev &D, int v1, int v2, int depth) {
    if (depth == 5 && D.vnb) return within_five_links(D, v1, v2);
    if (depth > 2 && is_neighbor_dfs(D, v1, v2, 2)) return true;
    return is_neighbor_dfs(D, v1, v2, depth);
}
__device__ bool is_neighbor_dfs(const Dev &D, int v1, int v2, int depth) {
    if (v1 == v2) return true;
    if (depth <= 0) return false;
    int sv[6], sl[6], si[6];
    int d = 0;
    sv[0] = v1; sl[0] = -1; si[0] = 0;
    while (d >= 0) {
        if (si[d] >= 6) { d--; continue; }
        const int i = si[d]++;
        const int li = D.vlinks[6 * (size_t)sv[d] + i];
        if (li < 0 || li == sl[d]) continue;
        const int2 e = D.lends[li];
        const int other = (e.x == sv[d]) ? e.y : e.x;
        if (other == v2) return true;
        if (depth - (d + 1) <= 0) continue;
        d++;
        sv[d] = other; sl[d] = li; si[d] = 0;
    }
    return false;
}

// The contact phase's question "are hi and lo within five links?", with memory: the answer depends on the link graph only, the
// graph changes on a handful of steps (SimD::topo_epoch counts the changes), and a settled pile asks about the same touching
// pairs every step.  Eight entries per voxel (its row is written by the lanes of the one warp that leads the voxel's pairs; an entry
// is one 8-byte word, so a reader sees an old or a new entry, never a torn one, and either is valid for the epoch it names).
// Only k_contact uses it: the resolve phase changes the graph between its own queries.
__device__ __forceinline__ bool within_five_links_cached(const Dev &D, int sim, int hi, int lo) {
    if (!D.nbcache) return is_neighbor(D, hi, lo, 5);
    const int epoch = D.simd[sim].topo_epoch;
    int2 *row = D.nbcache + 8 * (size_t)hi;
    const int4 *r4 = reinterpret_cast<const int4 *>(row);
    const int4 q0 = r4[0], q1 = r4[1], q2 = r4[2], q3 = r4[3];
    const int px[8] = {q0.x, q0.z, q1.x, q1.z, q2.x, q2.z, q3.x, q3.z}, ev[8] = {q0.y, q0.w, q1.y, q1.w, q2.y, q2.w, q3.y, q3.w};
    int victim = -1;
    const int h0 = (int)(((unsigned)lo * 0x9E3779B1u) >> 29);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (px[k] == lo && (ev[k] >> 1) == epoch) return ev[k] & 1;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) { // a stale entry to replace, looked for from a partner-dependent start so that lanes spread out
        const int j = (h0 + k) & 7;
        if (victim < 0 && (ev[j] >> 1) != epoch) victim = j;
    }
    if (victim < 0) victim = h0;
    const bool nb = is_neighbor(D, hi, lo, 5);
    row[victim] = make_int2(lo, (epoch << 1) | (nb ? 1 : 0));
    return nb;
}

// test hook (vx3_batch_check_neighbor_search): both searches on pseudo-random pairs of one simulation's voxels
__global__ void k_check_neighbor_search(Dev D, int v0, int nv, int n_pairs, unsigned seed, int *out2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    unsigned h = seed * 0x9E3779B1u + (unsigned)i * 0x85EBCA6Bu;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    const int a = (int)(h % (unsigned)nv);
    unsigned g = h * 0xC2B2AE35u + 0x27D4EB2Fu;
    g ^= g >> 16;
    // every other pair from the same few dozen indices (a body of the pile), the rest from a wider window
    const int span = (i & 1) ? 64 : 600;
    int bq = a + (int)(g % (unsigned)(2 * span + 1)) - span;
    bq = bq < 0 ? 0 : (bq >= nv ? nv - 1 : bq);
    const bool fast = within_five_links(D, v0 + a, v0 + bq);
    const bool walk = is_neighbor_dfs(D, v0 + a, v0 + bq, 5);
    if (fast != walk) atomicAdd(out2, 1);
    if (walk) atomicAdd(out2 + 1, 1);
}

// VX3_Collision (VX3_Collision.cu:3-31): force stored on voxel1 (= the higher index of the pair)
__device__ __forceinline__ V3 pair_contact_force(const Dev &D, int hi, int lo, const VoxMatC &m1, const VoxMatC &m2) {
    const double penetrationStiff = 2.0f / (1.0f / m1.penStiff + 1.0f / m2.penStiff);
    const double dampingC = 0.5f * (m1.colDampT + m2.colDampT);
    const V3 p1 = load_pos(D.pose, hi), p2 = load_pos(D.pose, lo);
    const V3 offset = p2 - p1;
    const double NomDist = (base_size_average(m1, D.tempe[hi]) + base_size_average(m2, D.tempe[lo])) * VX3_COLLISION_ENVELOPE_RADIUS;
    const double RelDist = NomDist - offset.Length();
    if (RelDist > 0) {
        const V3 unit = offset.Normalized();
        const V3 vel1 = load_linmom(D, hi) * m1.massInverse, vel2 = load_linmom(D, lo) * m2.massInverse;
        const double relativeVelocity = vel1.Dot(unit) - vel2.Dot(unit);
        return unit * (penetrationStiff * RelDist + dampingC * relativeVelocity);
    }
    return V3(0, 0, 0);
}

// Is voxel u (found in one of the 27 cells around v's) a contact partner of v?  The envelope test of
// handle_collision_attachment (VX3_VoxelyzeKernel.cu:682-705); direct lattice neighbours are skipped (is_neighbor depth 1)
// unless the link was made this step (`fresh`: its contact force is added and taken back, :827-830).
struct ContactSelf {
    int v;
    ContactRec r;
    // first cut on the bucket's float positions: nothing farther than the largest possible envelope of this simulation (= the
    // grid's cell edge) plus the rounding of the two float positions can pass the exact test; voxels of other simulations that
    // hashed to the bucket fall outside [vlo, vhi)
    float fx, fy, fz, reach;
    int vlo, vhi;
};
// the 64-byte record in two 256-bit requests (sm_100 LDG.E.ENL2.256) instead of four 128-bit ones: a contact sweep looks at
// ~90 candidate records per surface voxel, and the phase is bound by the number of memory requests it issues
__device__ __forceinline__ ContactRec load_crec(const Dev &D, int v) {
    ContactRec r;
    unsigned long long a0, a1, a2, a3, b0, b1, b2, b3;
    const ContactRec *s = D.crec + v;
    asm("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(s));
    asm("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(b0), "=l"(b1), "=l"(b2), "=l"(b3) : "l"(reinterpret_cast<const char *>(s) + 32));
    r.cx = (int)(unsigned)a0; r.cy = (int)(a0 >> 32); r.cz = (int)(unsigned)a1; r.bucket = (int)(a1 >> 32);
    r.px = __longlong_as_double((long long)a2); r.py = __longlong_as_double((long long)a3);
    r.pz = __longlong_as_double((long long)b0); r.bs = __longlong_as_double((long long)b1);
    r.sim = (int)(unsigned)b2; r.mat = (int)(b2 >> 32); r.fixed = (int)(unsigned)b3; r._pad = 0;
    return r;
}
__device__ __forceinline__ void contact_self(const Dev &D, int v, ContactSelf &c) {
    c.v = v;
    c.r = load_crec(D, v);
    if (c.r.bucket >= 0) {
        const SimC &S = D.simc[c.r.sim];
        c.fx = (float)c.r.px; c.fy = (float)c.r.py; c.fz = (float)c.r.pz;
        const float edge = (float)(1.0 / S.cell_inv);
        // float rounding of a coordinate is <= 2^-24 |x|; both positions are within a few cell edges of each other
        c.reach = edge * 1.001f + 2.4e-7f * (fabsf(c.fx) + fabsf(c.fy) + fabsf(c.fz) + 6.0f * edge);
        c.vlo = S.voff;
        c.vhi = S.voff + S.nvox;
    }
}
// conservative: false only when the exact envelope test (contact_envelope) must fail as well
__device__ __forceinline__ bool contact_first_cut(const ContactSelf &c, const float4 it) {
    const int u = __float_as_int(it.w);
    if (u < c.vlo || u >= c.vhi || u == c.v) return false;
    const float dx = it.x - c.fx, dy = it.y - c.fy, dz = it.z - c.fz;
    return dx * dx + dy * dy + dz * dz <= c.reach * c.reach;
}
// One partner of v (sorted position irrelevant here): contact force on v from this pair, target hit, signal trigger and,
// for the pairs v leads (v is the higher index) when emit is set, the attach-candidate test (:729-812) on the step-start
// link graph with an atomic append of the candidate.
__device__ __forceinline__ V3 contact_partner(const Dev &D, const SimC &S, int v, int entry, bool emit, int &hits, bool &fire) {
    const int u = entry & 0x3FFFFFFF;
    const bool fresh = (entry >> 30) & 1;
    const int hi = v > u ? v : u, lo = v > u ? u : v;
    const VoxMatC &m1 = D.vmat_tab[D.vmat[hi]], &m2 = D.vmat_tab[D.vmat[lo]];
    V3 f(0, 0, 0);
    if (S.enable_collision) {
        f = pair_contact_force(D, hi, lo, m1, m2);
        if (v != hi) f = -f;
        if ((m1.is_target && !m2.is_target) || (m2.is_target && !m1.is_target)) {
            if (v == hi) hits++;
            if (!D.vmat_tab[D.vmat[v]].is_target) fire = true;
        }
    }
    if (!emit || v != hi || fresh) return f;
    const int fl = D.vflags[lo], fh = D.vflags[hi];
    if (!(fh & VXF_ENABLE_ATTACH) || !(fl & VXF_ENABLE_ATTACH)) return f;
    if (m1.fixed || m2.fixed) return f;
    if (D.vmat[hi] != D.vmat[lo]) return f;
    if (!m1.sticky) return f;
    V3 p1, p2;
    Q4 q1;
    load_pose(D.pose, hi, p1, q1);
    p2 = load_pos(D.pose, lo);
    const V3 e = p1 - p2;
    const V3 ea = q1.RotateVec3DInv(-e);
    const V3 fa = ea.Abs();
    int dir1, dir2, axis, rev = 0;
    if (fa.x >= fa.y && fa.x >= fa.z) {
        axis = 0;
        if (ea.x < 0) { dir1 = 1; dir2 = 0; rev = 1; } else { dir1 = 0; dir2 = 1; }
    } else if (fa.y >= fa.x && fa.y >= fa.z) {
        axis = 1;
        if (ea.y < 0) { dir1 = 3; dir2 = 2; rev = 1; } else { dir1 = 2; dir2 = 3; }
    } else {
        axis = 2;
        if (ea.z < 0) { dir1 = 5; dir2 = 4; rev = 1; } else { dir1 = 4; dir2 = 5; }
    }
    // slots only fill up during the attach phase, so an occupied slot now stays a rejection at this pair's turn
    if (D.vlinks[6 * (size_t)hi + dir1] >= 0 || D.vlinks[6 * (size_t)lo + dir2] >= 0) return f;
    if (!uf_disconnected(D, hi, lo) && within_five_links_cached(D, D.vsim[hi], hi, lo)) return f; // links are only added during the phase: true now stays true
    const int slot = atomicAdd(&D.simd[D.vsim[hi]].cand_count, 1);
    if (slot < S.cand_cap) {
        Cand cd;
        cd.key = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
        cd.info = dir1 | (dir2 << 3) | (axis << 6) | (rev << 8);
        cd._pad = 0;
        D.cands[S.cand_off + slot] = cd;
    }
    return f;
}

__device__ __forceinline__ void contact_fire_signal(const Dev &D, const SimD &dy, int v, int matv) { // receiveSignal(100, currentTime, force = true) (VX3_Voxel.cu:315-335)
    const SigMatC &sm = D.smat_tab[matv];
    double *sg = D.sig + 6 * (size_t)v;
    const double t = dy.t;
    sg[2] = t + sm.inactive_period;
    sg[0] = 100.0;
    double val = 100.0 * sm.value_decay;
    if (val < 0.1) val = 0;
    sg[4] = val;
    sg[5] = t;
}

// The contact phase of the step, one WARP per surface voxel: 27 lanes walk the 27 cells around the voxel's cell at once,
// the partners found are ranked by index, every lane evaluates one partner (contact force, target hit, attach-candidate
// test incl. the depth-5 neighbour search), and lane 0 adds the forces up in ascending partner index — the same sums in the
// canonical sequential pair order.
#define VX3_CONTACT_WARPS 4
#ifndef VX3_CONTACT_MIN_CTAS
#define VX3_CONTACT_MIN_CTAS 8 // 64 registers: the phase is a chain of dependent loads per warp, so resident warps are what counts (6 / 8 / 10 / 12 CTAs: 59.9 / 52.4 / 62.2 / 69.6 us on config 4)
#endif
#define VX3_MAX_SURVIVORS 192 // voxels that pass the float first cut and are not plain lattice neighbours, per surface voxel
// the exact envelope test of handle_collision_attachment (VX3_VoxelyzeKernel.cu:682-697) on two ContactRecs, voxel1 = higher index
__device__ __forceinline__ bool contact_envelope(const ContactRec &rv, int v, const ContactRec &ru, int u) {
    if (rv.fixed && ru.fixed) return false;
    const V3 pv(rv.px, rv.py, rv.pz), pu(ru.px, ru.py, ru.pz);
    const V3 diff = (v > u) ? (pv - pu) : (pu - pv);
    const double watch = ((v > u) ? (rv.bs + ru.bs) : (ru.bs + rv.bs)) * VX3_COLLISION_ENVELOPE_RADIUS;
    if (diff.x > watch || diff.x < -watch) return false;
    if (diff.y > watch || diff.y < -watch) return false;
    if (diff.z > watch || diff.z < -watch) return false;
    return !(diff.Length() > watch);
}

// one warp's scratch for contact_warp
struct ContactWarpSmem {
    int surv[VX3_MAX_SURVIVORS];
    unsigned char from[VX3_MAX_SURVIVORS]; // which of the 27 cells the survivor was found under
    int list[VX3_MAX_PARTNERS], sorted[VX3_MAX_PARTNERS];
    double force[VX3_MAX_PARTNERS][3];
    int cnt[2];
};
// emit: the step's contact phase (k_contact).  !emit: the re-evaluation of the two voxels of an attach the resolve phase has just
// accepted (the pair's contact force is added and taken back in its place in the sum, :827-830) — forces only: no attach
// candidates, no collision count, no signal.
__device__ __forceinline__ void contact_warp(const Dev &D, int v, bool emit, int lane, ContactWarpSmem &W) {
    ContactSelf cs;
    contact_self(D, v, cs);
    if (cs.r.bucket < 0) return;
    const SimC &S = D.simc[cs.r.sim];
    SimD &dy = D.simd[cs.r.sim];
    if (lane < 2) W.cnt[lane] = 0;
    // ---- my own links, one per lane: the far end, and whether the link was made in this step (its contact force is added and
    // taken back, :827-830) — a plain lattice neighbour is no contact partner (is_neighbor depth 1, :699-703) ----
    int my_vo = -1, my_fresh = 0;
    if (lane < 6) {
        const int li = D.vlinks[6 * (size_t)v + lane];
        if (li >= 0) {
            const int2 e = D.lends[li];
            my_vo = (e.x == v) ? e.y : e.x;
            my_fresh = (D.lstate[li] & LKS_JUST_CREATED) ? 1 : 0;
        }
    }
    int vo[6];
    unsigned freshmask = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        vo[i] = __shfl_sync(0xFFFFFFFFu, my_vo, i);
        freshmask |= (unsigned)__shfl_sync(0xFFFFFFFFu, my_fresh, i) << i;
    }
    __syncwarp();
    // ---- stage 1: 27 lanes walk the 27 cells; whoever passes the float first cut and is not a plain neighbour goes on the list ----
    auto consider = [&](int u) {
        int tag = u;
#pragma unroll
        for (int i = 0; i < 6; i++)
            if (vo[i] == u) {
                if (!((freshmask >> i) & 1)) return; // linked, and not in this step: never a partner
                tag = u | (1 << 30);
            }
        const int k = atomicAdd(&W.cnt[0], 1);
        if (k < VX3_MAX_SURVIVORS) {
            W.surv[k] = tag;
            W.from[k] = (unsigned char)lane;
        }
    };
    if (lane < 27) {
        const int dx = lane % 3 - 1, dy_ = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
        const int cx = cs.r.cx + dx, cy = cs.r.cy + dy_, cz = cs.r.cz + dz;
        const int b = (int)(cell_hash(cs.r.sim, cx, cy, cz) & (unsigned)D.hmask);
        const int nb = D.cell_cnt[b];
        if (nb > 0) {
            const float4 *items = reinterpret_cast<const float4 *>(D.cell_items + VX3_CELL_SLOTS * (size_t)b);
            const int m = nb < VX3_CELL_SLOTS ? nb : VX3_CELL_SLOTS;
            float4 it[VX3_CELL_SLOTS];
#pragma unroll
            for (int k = 0; k < VX3_CELL_SLOTS; k++)
                if (k < m) it[k] = __ldcg(items + k);
#pragma unroll
            for (int k = 0; k < VX3_CELL_SLOTS; k++)
                if (k < m && contact_first_cut(cs, it[k])) consider(__float_as_int(it[k].w));
            if (nb > VX3_CELL_SLOTS) // overflow chain: no float copy, straight to the exact test
                for (int u = D.cell_ovf[b] - 1; u >= 0; u = D.cell_next[u])
                    if (u != v) consider(u);
        }
    }
    __syncwarp();
    int ns = W.cnt[0];
    if (ns > VX3_MAX_SURVIVORS) {
        if (lane == 0) dy.err = VX3_ERR_CAPACITY;
        ns = VX3_MAX_SURVIVORS;
    }
    // ---- stage 2: one lane per survivor: its record, the exact test (same cell-independent arithmetic as the all-pairs sweep) ----
    for (int i = lane; i < ns; i += 32) {
        const int tag = W.surv[i], u = tag & 0x3FFFFFFF;
        const ContactRec ur = load_crec(D, u);
        // two of the 27 cells (or a cell of another simulation) may share a hash bucket: a voxel counts only under its own cell
        const int from = W.from[i];
        if (ur.cx != cs.r.cx + from % 3 - 1 || ur.cy != cs.r.cy + (from / 3) % 3 - 1 || ur.cz != cs.r.cz + from / 9 - 1 || ur.sim != cs.r.sim) continue;
        if (!contact_envelope(cs.r, v, ur, u)) continue;
        const int k = atomicAdd(&W.cnt[1], 1);
        if (k < VX3_MAX_PARTNERS) W.list[k] = tag;
    }
    __syncwarp();
    int n = W.cnt[1];
    if (n > VX3_MAX_PARTNERS) {
        if (lane == 0) dy.err = VX3_ERR_CAPACITY;
        n = VX3_MAX_PARTNERS;
    }
    if (n == 0) { // nobody near: the pending contact force is zero (the voxel pass cleared it)
        if (lane == 0 && S.enable_collision) store3(D.contact, v, V3(0, 0, 0));
        return;
    }
    // rank by partner index (indices are unique)
    for (int i = lane; i < n; i += 32) {
        const int key = W.list[i] & 0x3FFFFFFF;
        int r = 0;
        for (int j = 0; j < n; j++) r += (W.list[j] & 0x3FFFFFFF) < key;
        W.sorted[r] = W.list[i];
    }
    __syncwarp();
    int hits = 0;
    bool fire = false;
    for (int i = lane; i < n; i += 32) {
        const V3 f = contact_partner(D, S, v, W.sorted[i], emit, hits, fire);
        W.force[i][0] = f.x; W.force[i][1] = f.y; W.force[i][2] = f.z;
    }
    __syncwarp();
    hits = __reduce_add_sync(0xFFFFFFFFu, hits);
    fire = __any_sync(0xFFFFFFFFu, fire);
    if (lane == 0 && S.enable_collision) {
        V3 c(0, 0, 0);
        for (int i = 0; i < n; i++) {
            const V3 f(W.force[i][0], W.force[i][1], W.force[i][2]);
            c += f;
            if ((W.sorted[i] >> 30) & 1) c -= f;
        }
        store3(D.contact, v, c);
        if (emit && hits) atomicAdd(&dy.collision_count, hits);
        if (emit && fire && S.enable_signals) contact_fire_signal(D, dy, v, cs.r.mat);
    }
}

__global__ void __launch_bounds__(32 * VX3_CONTACT_WARPS, VX3_CONTACT_MIN_CTAS) k_contact(Dev D) {
    VX3_PDL_ENTRY();
    __shared__ ContactWarpSmem sm[VX3_CONTACT_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int v = blockIdx.x * VX3_CONTACT_WARPS + w;
    if (v >= D.nvox) return; // (whole warps leave together)
    contact_warp(D, v, true, lane, sm[w]);
}

// Attach resolution, then detach, for ONE simulation per CTA.
//
// Attach: sequential resolution of the simulation's candidates in canonical (first, second) order: a candidate is accepted
// only if both facing slots are still empty and the voxels are still not within 5 links at its turn.  Creates
// the link like VX3_Link's device ctor + reset() (VX3_Link.cu:31-70).  The candidates are sorted by key first: in shared memory
// up to 2048 of them, in place in the simulation's (power-of-two sized) region of the candidate array beyond that.
// Detach (gpu_update_detach, VX3_VoxelyzeKernel.cu:946-968): the links the link pass put on the failed list leave their voxels'
// slots — after the attach phase, as in the reference, so a slot freed now cannot be claimed in the same step.
#define VX3_RESOLVE_SM 2048
#ifndef VX3_RESOLVE_T
#define VX3_RESOLVE_T 256
#endif
// ---- transverse info of a link the attach phase creates (VX3_Link::reset -> updateTransverseInfo, VX3_Link.cu:58-70, 85-88) ----
// VX3 never refreshes a link's transverse area / strain sum while stepping (VX3_Link.cu:146-150, commented out), so they are frozen
// at creation: size^2 and 0 at import (every strain is zero), but for a link made by the attach phase the end voxels' CURRENT
// poissons strains when nu != 0 — through the voxel's cache (poissonsStrain(), VX3_Voxel.cu:470-476), which is invalidated at the
// end of every voxel time step (:266).  The cache lives in Dev::pcache with a validity stamp instead of a flag the voxel pass
// would have to write every step: -2 invalid; -1 valid since import (zeros) — for good on a voxel that never reaches the end of
// timeStep (all DOFs fixed), until the first step has been integrated otherwise; s >= 0: computed during step s.
__device__ __forceinline__ float link_axial_strain(const Dev &D, int li, bool positiveEnd) { // VX3_Link::axialStrain, VX3_Link.cu:72-74
    const int2 e = D.lends[li];
    const float strain = D.lstrain[li].x;
    const float strainRatio = D.vmat_tab[D.vmat[e.y]].E / D.vmat_tab[D.vmat[e.x]].E; // pVPos E / pVNeg E (reset(), :63)
    return positiveEnd ? 2.0f * strain * strainRatio / (1.0f + strainRatio) : 2.0f * strain / (1.0f + strainRatio);
}
__device__ void voxel_poissons_strain(const Dev &D, int v, int steps_now, float ps[3]) { // poissonsStrain() with strain(true), VX3_Voxel.cu:428-476
    const float4 c = D.pcache[v];
    const int stamp = __float_as_int(c.w);
    const int ext = D.vext[v];
    const bool fixedAll = ext >= 0 && (D.exts[ext].dof & 0x3F) == 0x3F;
    const bool valid = stamp == -1 ? (fixedAll || steps_now == 0) : stamp == steps_now;
    if (valid) {
        ps[0] = c.x; ps[1] = c.y; ps[2] = c.z;
        return;
    }
    float intStrRet[3] = {0.f, 0.f, 0.f};
    int numBondAxis[3] = {0, 0, 0};
    bool tension[3];
    for (int i = 0; i < 6; i++) {
        const int li = D.vlinks[6 * (size_t)v + i];
        if (li >= 0) {
            intStrRet[i >> 1] += link_axial_strain(D, li, (i & 1) != 0); // isNegative(direction): this voxel is that link's positive end
            numBondAxis[i >> 1]++;
        }
    }
    for (int i = 0; i < 3; i++) {
        if (numBondAxis[i] == 2) intStrRet[i] *= 0.5f;
        tension[i] = (numBondAxis[i] == 2) || (ext >= 0 && (numBondAxis[i] == 1 && ((D.exts[ext].dof & (1 << i)) || D.exts[ext].force[i] != 0)));
    }
    if (!(tension[0] && tension[1] && tension[2])) {
        float add = 0;
        for (int i = 0; i < 3; i++)
            if (tension[i]) add += intStrRet[i];
        const float value = powf(1.0f + add, -D.vmat_tab[D.vmat[v]].nu) - 1.0f;
        for (int i = 0; i < 3; i++)
            if (!tension[i]) intStrRet[i] = value;
    }
    ps[0] = intStrRet[0]; ps[1] = intStrRet[1]; ps[2] = intStrRet[2];
    D.pcache[v] = make_float4(ps[0], ps[1], ps[2], __int_as_float(fixedAll ? -1 : steps_now));
}
// {currentTransverseArea, currentTransverseStrainSum} of the new link g (already in both voxels' slots, strain 0)
__device__ float2 resolve_link_transverse(const Dev &D, int vneg, int vpos, int axis, int steps_now) {
    float area[2], sum[2];
    const int vs[2] = {vneg, vpos};
    for (int k = 0; k < 2; k++) {
        const VoxMatC &m = D.vmat_tab[D.vmat[vs[k]]];
        const float size = (float)m.nomSize;
        if (m.nu == 0 || !D.pcache) { // transverseArea / transverseStrainSum return early (VX3_Voxel.cu:479, 498)
            area[k] = size * size;
            sum[k] = 0.0f;
            continue;
        }
        float p[3];
        voxel_poissons_strain(D, vs[k], steps_now, p);
        const int a = (axis + 1) % 3, b = (axis + 2) % 3; // the two other axes, ascending: (y, z), (x, z), (x, y)
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        area[k] = (float)(size * size * (1 + (double)p[lo]) * (1 + (double)p[hi]));
        sum[k] = p[lo] + p[hi];
    }
    return make_float2(0.5f * (area[0] + area[1]), 0.5f * (sum[0] + sum[1]));
}

// (true: the link was made — the caller re-evaluates the two voxels' contact forces)
__device__ bool resolve_accept(const Dev &D, const SimC &S, SimD &dy, int sim, unsigned long long key, int info) {
    const int hi = (int)(key >> 32), lo = (int)(key & 0xFFFFFFFFu);
    const int dir1 = info & 7, dir2 = (info >> 3) & 7, axis = (info >> 6) & 3, rev = (info >> 8) & 1;
    if (D.vlinks[6 * (size_t)hi + dir1] >= 0 || D.vlinks[6 * (size_t)lo + dir2] >= 0) return false;
    if (!uf_disconnected(D, hi, lo) && is_neighbor(D, hi, lo, 5)) return false;
    if (dy.link_cnt >= S.lcap) { dy.err = VX3_ERR_CAPACITY; return false; }
    const VoxMatC &mh = D.vmat_tab[D.vmat[hi]];
    if (mh.self_lmat < 0) { dy.err = VX3_ERR_INVALID; return false; }
    const int g = S.loff + dy.link_cnt++;
    const int vneg = rev ? lo : hi, vpos = rev ? hi : lo; // pVNeg/pVPos of VX3_Link(voxelA, dirA, voxelB, dirB)
    D.vlinks[6 * (size_t)hi + dir1] = g;
    D.vlinks[6 * (size_t)lo + dir2] = g;
    if (D.vnb) {
        D.vnb[8 * (size_t)hi + dir1] = lo;
        D.vnb[8 * (size_t)lo + dir2] = hi;
    }
    D.lends[g] = make_int2(vneg, vpos);
    D.lmat[g] = mh.self_lmat;
    D.lc4[g] = make_int4(vneg, vpos, mh.self_lmat, sim);
    D.lstate[g] = (axis << LKS_AXIS_SHIFT) | LKS_SMALL | LKS_JUST_CREATED | (S.safety_guard << LKS_NEWLINK_SHIFT);
    const VoxMatC &mn = D.vmat_tab[D.vmat[vneg]], &mp = D.vmat_tab[D.vmat[vpos]];
    for (int k = 0; k < 4; k++) *D.lh(k, g) = make_double2(0.0, 0.0);
    *D.lh(4, g) = make_double2(0.0, 0.5 * (base_size_axis(mn, D.tempe[vneg], axis) + base_size_axis(mp, D.tempe[vpos], axis)));
    for (int k = 0; k < 6; k++) *D.lf(k, g) = make_double2(0.0, 0.0); // the new link's end forces start at zero (VX3_Link::reset)
    D.lstrain[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    D.larea[g] = resolve_link_transverse(D, vneg, vpos, axis, (int)dy.steps); // updateTransverseInfo(): frozen for the link's life
    dy.attach_events++;
    dy.topo_epoch++;
    if (D.uf) { // the two voxels' trees become one
        const int ra = uf_find(D, hi), rb = uf_find(D, lo);
        if (ra != rb) D.uf[ra > rb ? ra : rb] = ra > rb ? rb : ra;
    }
    __threadfence_block();
    return true;
}

__global__ void __launch_bounds__(VX3_RESOLVE_T) k_resolve_detach(Dev D) {
    VX3_PDL_ENTRY();
    __shared__ unsigned long long skey[VX3_RESOLVE_SM];
    __shared__ int sinfo[VX3_RESOLVE_SM];
    __shared__ ContactWarpSmem cw;
    const int sim = blockIdx.x;
    const SimC &S = D.simc[sim];
    SimD &dy = D.simd[sim];
    int n = S.cand_cap > 0 ? dy.cand_count : 0;
    const int nfail = S.fail_cap > 0 ? dy.fail_count : 0;
    if (n == 0 && nfail == 0) return;
    const bool running = dy.status == VX3_SIM_RUNNING && !dy.diverged && dy.dt != 0;
    if (threadIdx.x == 0) {
        if (n > dy.cand_peak) dy.cand_peak = n;
        if (nfail > dy.fail_peak) dy.fail_peak = nfail;
    }
    if (n > S.cand_cap) { // more simultaneous candidates than the region holds: this simulation reports it, the others go on
        if (threadIdx.x == 0) dy.err = VX3_ERR_CAPACITY;
        n = S.cand_cap;
    }
    if (n > 0 && running) {
        Cand *cands = D.cands + S.cand_off;
        int np2 = 1;
        while (np2 < n) np2 <<= 1;
        const bool in_sm = np2 <= VX3_RESOLVE_SM;
        if (in_sm) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                skey[i] = i < n ? cands[i].key : ~0ull;
                sinfo[i] = i < n ? cands[i].info : 0;
            }
        } else {
            for (int i = n + threadIdx.x; i < np2; i += blockDim.x) cands[i].key = ~0ull; // pad in place (np2 <= cand_cap)
        }
        __syncthreads();
        for (int k = 2; k <= np2; k <<= 1) // bitonic sort by (first, second)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj <= i) continue;
                    const bool up = (i & k) == 0;
                    if (in_sm) {
                        if ((skey[i] > skey[ixj]) == up) {
                            const unsigned long long tk = skey[i]; skey[i] = skey[ixj]; skey[ixj] = tk;
                            const int ti = sinfo[i]; sinfo[i] = sinfo[ixj]; sinfo[ixj] = ti;
                        }
                    } else if ((cands[i].key > cands[ixj].key) == up) {
                        const Cand t = cands[i]; cands[i] = cands[ixj]; cands[ixj] = t;
                    }
                }
                __syncthreads();
            }
        if (threadIdx.x < 32) { // one warp: lane 0 takes the candidates in order; an accepted pair's two voxels get their contact forces
                                // re-evaluated by the whole warp (the pair's force is added and taken back in its place, :827-830)
            const int lane = threadIdx.x;
            for (int c = 0; c < n; c++) {
                const unsigned long long key = in_sm ? skey[c] : cands[c].key;
                int made = 0;
                if (lane == 0) made = resolve_accept(D, S, dy, sim, key, in_sm ? sinfo[c] : cands[c].info) ? 1 : 0;
                made = __shfl_sync(0xFFFFFFFFu, made, 0);
                if (made && S.enable_collision) {
                    __syncwarp();
                    contact_warp(D, (int)(key >> 32), false, lane, cw);
                    __syncwarp();
                    contact_warp(D, (int)(key & 0xFFFFFFFFu), false, lane, cw);
                    __syncwarp();
                }
            }
        }
    }
    __syncthreads(); // detach after attach
    if (nfail > 0 && running) {
        const int m = nfail < S.fail_cap ? nfail : S.fail_cap;
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            const int g = D.fail_list[S.fail_off + k];
            const int st = D.lstate[g];
            if (st & (LKS_DETACHED | LKS_REMOVED)) continue;
            const int2 e = D.lends[g];
            D.lstate[g] = (st | LKS_DETACHED) & ~LKS_FAILED;
            for (int i = 0; i < 6; i++) {
                if (D.vlinks[6 * (size_t)e.x + i] == g) {
                    D.vlinks[6 * (size_t)e.x + i] = -1;
                    if (D.vnb) D.vnb[8 * (size_t)e.x + i] = -1;
                }
                if (D.vlinks[6 * (size_t)e.y + i] == g) {
                    D.vlinks[6 * (size_t)e.y + i] = -1;
                    if (D.vnb) D.vnb[8 * (size_t)e.y + i] = -1;
                }
            }
            atomicAdd(&dy.detach_events, 1);
            atomicAdd(&dy.topo_epoch, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        dy.cand_count = 0;
        dy.fail_count = 0;
    }
}

// SecondaryExperiment (VX3_VoxelyzeKernel.cu:336-350): removeVoxels (:365-399) per voxel — a voxel whose material's
// RemoveFromSimulationAfterThisManySeconds has passed is marked removed together with its links, and both ends'
// slots are cleared — and the one-time re-initialisation of the initial positions (saveInitialPosition).
// Runs after the voxel pass, before k_tail (which still holds the step's currentTime).
__global__ void __launch_bounds__(VX3_BLOCK) k_secondary(Dev D) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= D.nvox) return;
    const int sim = D.nsims == 1 ? 0 : D.vsim[v];
    const SimC &S = D.simc[sim];
    const SimD &dy = D.simd[sim];
    if (!S.secondary_experiment || dy.status != VX3_SIM_RUNNING || dy.diverged || dy.dt == 0) return;
    const double t = dy.t;
    const VoxMatC &m = D.vmat_tab[D.vmat[v]];
    const int flags = D.vflags[v];
    if (!(flags & VXF_REMOVED) && m.remove_after > 0 && m.remove_after < t) {
        D.vflags[v] = flags | VXF_REMOVED;
        // gpu_update_temperature skips removed voxels (:632-633): the temperature freezes at THIS step's value — the voxel pass
        // has already put the next step's into the pose record, take it back
        double *tp = D.pose + 8 * (size_t)v + 7;
        *tp = pack_tp(D.tempe[v], unpack_pd(*tp));
        for (int k = 0; k < 6; k++) {
            const int li = D.vlinks[6 * (size_t)v + k];
            if (li < 0) continue;
            D.lstate[li] |= LKS_REMOVED; // (a neighbour removed in the same step sets the same bit: benign)
            atomicAdd(&D.simd[sim].topo_epoch, 1);
            const int2 e = D.lends[li];
            const int nb = (e.x == v) ? e.y : e.x;
            for (int q = 0; q < 6; q++)
                if (D.vlinks[6 * (size_t)nb + q] == li) {
                    D.vlinks[6 * (size_t)nb + q] = -1;
                    if (D.vnb) D.vnb[8 * (size_t)nb + q] = -1;
                    break;
                }
            D.vlinks[6 * (size_t)v + k] = -1;
            if (D.vnb) D.vnb[8 * (size_t)v + k] = -1;
        }
    }
    if (!dy.initpos_reinitialized && S.reinit_after < t) store3(D.initpos, v, load_pos(D.pose, v)); // saveInitialPosition()
}

// ------------------------------------------------------------------ signals
// VX3_Voxel::propagateSignal / packMaker / localSignalDecay / receiveSignal (VX3_Voxel.cu:279-348), which the reference
// runs at the end of every voxel's timeStep with all voxels in parallel: a voxel writes its NEIGHBOURS' signal state, so
// the reference result depends on thread timing.  Defined behaviour here (and in the oracle): the voxels take their turns
// in ascending voxel index.  That order is reproduced in parallel:
//   1. sprop[i] = the value voxel i propagates AT ITS TURN.  It depends only on i's start state and on what its
//      LOWER-index neighbours propagate, so iterating sprop[i] = f(sprop[lower neighbours]) from any start converges to
//      the one assignment the sequential order produces (induction over the index); the loop runs until nothing changes
//      (one or two rounds unless many adjacent voxels are active at once).
//   2. every voxel then replays its own history of the step: receives from lower-index senders, its own
//      propagate / pacemaker / decay, receives from higher-index senders.
// One CTA per simulation (a simulation's voxels only talk to each other).
struct SigState {
    double ls, lsdt, inact, next, val, act;
};
__device__ __forceinline__ void sig_receive(SigState &s, const SigMatC &m, double value, double activeTime, bool force) { // :315-335
    if (!force && s.inact > activeTime) return;
    if (value < 0.1) return;
    s.inact = activeTime + m.inactive_period;
    s.ls = value;
    s.val = value * m.value_decay;
    if (s.val < 0.1) s.val = 0;
    s.act = activeTime;
}
// the ≤6 link neighbours of voxel v, ascending by voxel index (-1 = none, sorted last as INT_MAX)
__device__ __forceinline__ void sig_neighbours(const Dev &D, int v, int nb[6]) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int li = D.vlinks[6 * (size_t)v + k];
        int o = 0x7FFFFFFF;
        if (li >= 0) {
            const int2 e = D.lends[li];
            o = (e.x == v) ? e.y : e.x;
        }
        nb[k] = o;
    }
#pragma unroll
    for (int i = 1; i < 6; i++) { // insertion sort
        const int key = nb[i];
        int j = i - 1;
        while (j >= 0 && nb[j] > key) {
            nb[j + 1] = nb[j];
            j--;
        }
        nb[j + 1] = key;
    }
}
__device__ __forceinline__ bool sig_runs(const Dev &D, int v) { // does timeStep reach its end for this voxel (:162-174, :586-589)?
    if (D.vflags[v] & VXF_REMOVED) return false;
    if (D.vmat_tab[D.vmat[v]].fixed) return false;
    const int ext = D.vext[v];
    if (ext >= 0 && (D.exts[ext].dof & 0x3F) == 0x3F) return false;
    return true;
}
__global__ void __launch_bounds__(256) k_signals(Dev D) {
    __shared__ int s_any;
    const int sim = blockIdx.x;
    const SimC &S = D.simc[sim];
    const SimD &dy = D.simd[sim];
    if (!S.enable_signals || dy.status != VX3_SIM_RUNNING || dy.diverged || dy.dt == 0) return;
    const double t = dy.t;
    const int v0 = S.voff, v1 = S.voff + S.nvox;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    // ---- 1a. start value: what the voxel would propagate if nothing reached it before its turn ----
    bool any = false;
    for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        const double *sg = D.sig + 6 * (size_t)v;
        double pv = 0;
        if (sig_runs(D, v) && !(sg[5] > t) && !(sg[4] < 0.1)) pv = sg[4];
        D.sprop[v] = pv;
        any |= pv > 0;
    }
    if (any) s_any = 1;
    __syncthreads();
    const bool senders = s_any != 0;
    // ---- 1b. fixpoint over the lower-index senders ----
    if (senders) {
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_any = 0;
            __syncthreads();
            bool changed = false;
            for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
                if (!sig_runs(D, v)) continue;
                int nb[6];
                sig_neighbours(D, v, nb);
                const double *sg = D.sig + 6 * (size_t)v;
                SigState s{sg[0], sg[1], sg[2], sg[3], sg[4], sg[5]};
                const SigMatC &m = D.smat_tab[D.vmat[v]];
                int prev = -1;
                for (int k = 0; k < 6; k++) {
                    const int n = nb[k];
                    if (n >= v) break;
                    if (n == prev) continue;
                    prev = n;
                    const double pn = *(volatile double *)(D.sprop + n);
                    if (pn > 0) sig_receive(s, m, pn, t + D.smat_tab[D.vmat[n]].time_delay, false);
                }
                const double pv = (!(s.act > t) && !(s.val < 0.1)) ? s.val : 0.0;
                if (pv != *(volatile double *)(D.sprop + v)) {
                    D.sprop[v] = pv;
                    changed = true;
                }
            }
            if (changed) s_any = 1;
            __syncthreads();
            if (!s_any) break;
        }
    }
    // ---- 2. replay ----
    for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        double *sg = D.sig + 6 * (size_t)v;
        SigState s{sg[0], sg[1], sg[2], sg[3], sg[4], sg[5]};
        const SigMatC &m = D.smat_tab[D.vmat[v]];
        int nb[6];
        int k = 0, prev = -1;
        if (senders) {
            sig_neighbours(D, v, nb);
            for (; k < 6; k++) {
                const int n = nb[k];
                if (n >= v) break;
                if (n == prev) continue;
                prev = n;
                const double pn = D.sprop[n];
                if (pn > 0) sig_receive(s, m, pn, t + D.smat_tab[D.vmat[n]].time_delay, false);
            }
        }
        if (sig_runs(D, v)) {
            if (!(s.act > t) && !(s.val < 0.1)) { // propagateSignal (:336-362): the sends are replayed by the receivers
                s.val = 0;
                s.act = 0;
                s.inact = t + 2 * m.time_delay + m.inactive_period;
            }
            if (m.is_pacemaker && !(s.next > t)) { // packMaker (:304-313)
                sig_receive(s, m, 100.0, t, true);
                s.next = t + m.pacemaker_period;
            }
            if (!(s.lsdt > t)) { // localSignalDecay (:291-302)
                if (s.ls < 0.1) s.ls = 0;
                else {
                    s.ls = s.ls * 0.9;
                    s.lsdt = t + 0.01;
                }
            }
        }
        if (senders)
            for (; k < 6; k++) {
                const int n = nb[k];
                if (n == 0x7FFFFFFF) break;
                if (n == prev || n == v) continue;
                prev = n;
                const double pn = D.sprop[n];
                if (pn > 0) sig_receive(s, m, pn, t + D.smat_tab[D.vmat[n]].time_delay, false);
            }
        sg[0] = s.ls; sg[1] = s.lsdt; sg[2] = s.inact; sg[3] = s.next; sg[4] = s.val; sg[5] = s.act;
    }
}

// ------------------------------------------------------------------ reductions
// updateCurrentCenterOfMass (:477-493) stage 1 + the per-voxel sums of collectResults
// (VX3_SimulationManager.cu:455-466): one CTA per chunk of one simulation's voxels, fixed-order tree.
__global__ void __launch_bounds__(VX3_BLOCK) k_com_partial(Dev D) {
    __shared__ double sh[6][VX3_BLOCK];
    const Chunk ck = D.chunks[blockIdx.x];
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < ck.vcount; i += blockDim.x) {
        const int v = ck.vstart + i;
        const VoxMatC &m = D.vmat_tab[D.vmat[v]];
        if (!m.is_measured || (D.vflags[v] & VX3_VOX_GHOST)) continue;
        const V3 p = load_pos(D.pose, v);
        const double mass = m.mass;
        a[0] += p.x * mass; a[1] += p.y * mass; a[2] += p.z * mass; a[3] += mass;
        a[4] += p.Dist(load3(D.initpos, v));
        a[5] += 1.0;
    }
    for (int k = 0; k < 6; k++) sh[k][threadIdx.x] = a[k];
    __syncthreads();
    for (int off = VX3_BLOCK / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off)
            for (int k = 0; k < 6; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x < 6) D.com_part[6 * (size_t)blockIdx.x + threadIdx.x] = sh[threadIdx.x][0];
}

__device__ __forceinline__ void com_finalize(const Dev &D, const SimC &S, SimD &dy) {
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int c = 0; c < S.nchunks; c++)
        for (int k = 0; k < 6; k++) s[k] += D.com_part[6 * (size_t)(S.chunk_off + c) + k];
    if (s[3] == 0) { dy.com[0] = dy.com[1] = dy.com[2] = 0; }
    else {
        const double inv = 1.0 / s[3]; // Vec3D::operator/ multiplies by the reciprocal
        dy.com[0] = inv * s[0]; dy.com[1] = inv * s[1]; dy.com[2] = inv * s[2];
    }
    dy.total_dist = s[4];
    dy.n_measured = (int)s[5];
}

__device__ __forceinline__ bool stop_condition_met(const Dev &D, const SimC &S, const SimD &dy) { // :162-182
    if (S.prog_n[VX3_PROG_STOP] <= 0) return false;
    double vars[9];
    prog_vars(S, dy, dy.t, dy.com[0], dy.com[1], dy.com[2], vars);
    bool ok;
    return mt_eval<VX3_MAX_TOKENS>(D.tokens + S.prog_off[VX3_PROG_STOP], S.prog_n[VX3_PROG_STOP], vars, &ok) > 0;
}

// End of doTimeStep (:314-352), one CTA per simulation: CoM / angle / target-closeness sampling every
// int(TempPeriod/dt) steps, currentTime += dt, divergence and (run mode) the stop condition for the next step.
__global__ void __launch_bounds__(128) k_tail(Dev D, int com_ready, int check_stop) {
    __shared__ double red[128];
    __shared__ int redn[128];
    __shared__ int sample;
    const int sim = blockIdx.x;
    const SimC &S = D.simc[sim];
    SimD &dy = D.simd[sim];
    if (dy.status != VX3_SIM_RUNNING) return;
    const float dtF = dy.dt;
    if (threadIdx.x == 0) {
        sample = 0;
        dy.steps += 1;
        if (dtF != 0) {
            if (dy.diverged) dy.status = VX3_SIM_DIVERGED;
            else {
                const int CycleStep = (int)(S.temp_period / dtF);
                if (CycleStep > 0 && dy.steps % CycleStep == 0) {
                    if (!com_ready) dy.err = VX3_ERR_INVALID;
                    sample = 1;
                }
            }
        }
    }
    __syncthreads();
    if (dtF == 0 || dy.status != VX3_SIM_RUNNING) return;
    if (sample) {
        if (S.pair_radius != 0) { // computeTargetCloseness (:545-563)
            double acc = 0;
            int cnt = 0;
            const int nt = S.ntgt;
            for (int i = threadIdx.x; i < nt; i += blockDim.x) {
                const V3 pi = load_pos(D.pose, D.targets[S.tgt_off + i]);
                for (int j = i + 1; j < nt; j++) {
                    const double d = pi.Dist(load_pos(D.pose, D.targets[S.tgt_off + j]));
                    if (d < S.pair_radius) cnt++;
                    acc += 1 / d;
                }
            }
            red[threadIdx.x] = acc;
            redn[threadIdx.x] = cnt;
            __syncthreads();
            for (int off = 64; off > 0; off >>= 1) {
                if (threadIdx.x < off) {
                    red[threadIdx.x] += red[threadIdx.x + off];
                    redn[threadIdx.x] += redn[threadIdx.x + off];
                }
                __syncthreads();
            }
        }
        if (threadIdx.x == 0) {
            dy.angle_samples++;
            for (int k = 0; k < 3; k++) {
                dy.com_hist[0][k] = dy.com_hist[1][k];
                dy.com_hist[1][k] = dy.com[k];
            }
            com_finalize(D, S, dy);
            const V3 A(dy.com_hist[0][0], dy.com_hist[0][1], dy.com_hist[0][2]), B(dy.com_hist[1][0], dy.com_hist[1][1], dy.com_hist[1][2]),
                C(dy.com[0], dy.com[1], dy.com[2]);
            if (B == C || A == B || dy.angle_samples < 3) dy.recent_angle = 0;
            else dy.recent_angle = acos((B - A).Dot(C - B) / (B.Dist(A) * C.Dist(B)));
            if (S.pair_radius != 0) {
                dy.target_closeness = red[0];
                dy.num_close_pairs = redn[0];
            }
        }
    }
    if (threadIdx.x == 0) {
        if (S.secondary_experiment && !dy.initpos_reinitialized && S.reinit_after < dy.t) { // :344-348
            dy.initpos_reinitialized = 1;
            for (int k = 0; k < 3; k++) dy.com0[k] = dy.com[k]; // InitializeCenterOfMass()
        }
        dy.t += dtF;
        if (check_stop && stop_condition_met(D, S, dy)) dy.status = VX3_SIM_STOPPED;
    }
}

// k_tail for a step on which NO simulation of the batch samples its centre of mass (the host knows the cadence): one
// thread per simulation.  A simulation that nevertheless reaches a sampling step here reports VX3_ERR_INVALID.
// Everything the decision needs sits in the first 64 bytes of SimD and is loaded up front, in one round trip (as a chain of
// early-outs the same code was five dependent loads, 6.5 us per step for a kernel that does nothing most of the time).
__device__ void tail_light(const Dev &D, int sim, int check_stop) {
    SimD &dy = D.simd[sim];
    const int status = dy.status, diverged = dy.diverged, hot = dy.hot_flags;
    const float dtF = dy.dt;
    const double t = dy.t, period = dy.temp_period;
    const long long steps = dy.steps + 1;
    if (status != VX3_SIM_RUNNING) return;
    dy.steps = steps;
    if (dtF == 0) return;
    if (diverged) {
        dy.status = VX3_SIM_DIVERGED;
        return;
    }
    const int CycleStep = (int)(period / dtF);
    if (CycleStep > 0 && steps % CycleStep == 0) dy.err = VX3_ERR_INVALID;
    if (hot & SHF_SECONDARY) {
        const SimC &S = D.simc[sim];
        if (!dy.initpos_reinitialized && S.reinit_after < t) { // :344-348
            dy.initpos_reinitialized = 1;
            for (int k = 0; k < 3; k++) dy.com0[k] = dy.com[k]; // InitializeCenterOfMass()
        }
    }
    dy.t = t + dtF;
    if (check_stop && (hot & SHF_STOP_PROG) && stop_condition_met(D, D.simc[sim], dy)) dy.status = VX3_SIM_STOPPED;
}
__global__ void __launch_bounds__(128) k_tail_light(Dev D, int check_stop) {
    VX3_PDL_ENTRY();
    const int sim = blockIdx.x * blockDim.x + threadIdx.x;
    if (sim < D.nsims) tail_light(D, sim, check_stop);
}

// mode 0: initial state (saveInitialPosition is done on the host; InitializeCenterOfMass, VX3_SimulationManager.cu:54-55)
// mode 1: results (updateCurrentCenterOfMass + computeFitness, :116-117)   mode 2: stop check before the first step (:63)
__global__ void k_sim_update(Dev D, int mode) {
    const int sim = blockIdx.x * blockDim.x + threadIdx.x;
    if (sim >= D.nsims) return;
    const SimC &S = D.simc[sim];
    SimD &dy = D.simd[sim];
    if (mode == 2) {
        if (dy.status == VX3_SIM_RUNNING && stop_condition_met(D, S, dy)) dy.status = VX3_SIM_STOPPED;
        return;
    }
    com_finalize(D, S, dy);
    if (mode == 0) {
        for (int k = 0; k < 3; k++) dy.com0[k] = dy.com[k];
        return;
    }
    if (S.prog_n[VX3_PROG_FITNESS] <= 0) dy.fitness = 0;
    else { // computeFitness (:530-534)
        double vars[9];
        prog_vars(S, dy, dy.t, dy.com[0] - dy.com0[0], dy.com[1] - dy.com0[1], dy.com[2] - dy.com0[2], vars);
        bool ok;
        dy.fitness = mt_eval<VX3_MAX_TOKENS>(D.tokens + S.prog_off[VX3_PROG_FITNESS], S.prog_n[VX3_PROG_FITNESS], vars, &ok);
    }
}

// temperature at the current time into the pose records (batch creation): gpu_update_temperature of the first step
__global__ void __launch_bounds__(VX3_BLOCK) k_temp_init(Dev D) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= D.nvox) return;
    const int sim = D.vsim[v];
    const SimC &S = D.simc[sim];
    const VoxMatC &m = D.vmat_tab[D.vmat[v]];
    const double t = D.simd[sim].t;
    float tempe = D.tempe[v];
    if (thermal_active(S, m, D.vflags[v], t)) tempe = voxel_temperature(S, t, D.phase[v]);
    D.pose[8 * (size_t)v + 7] = pack_tp(tempe, 0.0f); // previousDt starts at 0 (VX3_Voxel.h:309)
}

__global__ void k_set_dt(Dev D, float dt) { // dt < 0: DtFrac * recommendedTimeStep() (:244-254)
    const int sim = blockIdx.x * blockDim.x + threadIdx.x;
    if (sim >= D.nsims) return;
    const SimC &S = D.simc[sim];
    if (dt < 0) {
        double od = S.optimal_dt;
        if (od < 1e-10) od = 1e-10;
        D.simd[sim].dt = (float)(S.dt_frac * od);
    } else
        D.simd[sim].dt = dt;
}

__global__ void k_step_cap(Dev D) {
    const int sim = blockIdx.x * blockDim.x + threadIdx.x;
    if (sim >= D.nsims) return;
    if (D.simd[sim].status == VX3_SIM_RUNNING) D.simd[sim].status = VX3_SIM_STEP_CAP;
}

} // namespace vx3
