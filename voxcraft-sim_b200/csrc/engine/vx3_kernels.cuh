// Streaming step kernels: one launch per phase over the concatenated SoA arrays of the whole batch.
// Phase order per step = VX3_VoxelyzeKernel::doTimeStep (src/VX3/VX3_VoxelyzeKernel.cu:237-359):
//   k_links -> [k_grid_count, k_grid_scan, k_grid_fill, k_contact, k_resolve] -> [k_detach] -> k_voxels
//   -> [k_com_partial] -> k_tail
#pragma once
#include "vx3_physics.cuh"

namespace vx3 {

#define VX3_BLOCK 256
// occupancy targets of the two streaming hot kernels (CTAs of VX3_BLOCK threads per SM); tuned on B200, see DESIGN.md §4
#ifndef VX3_LINKS_MIN_CTAS
#define VX3_LINKS_MIN_CTAS 2
#endif
#ifndef VX3_VOXELS_MIN_CTAS
#define VX3_VOXELS_MIN_CTAS 2
#endif

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
// The 8th double of the pose record carries the voxel's temperature AT THE CURRENT SIMULATION TIME (what
// gpu_update_temperature will set at the start of the next step): the voxel pass computes it once per voxel, the
// link pass reads it with the pose instead of evaluating sin() for both ends of every link.
__device__ __forceinline__ void load_pose_t(const double *__restrict__ pose, int v, V3 &p, Q4 &q, float &tempe) {
    const double2 *s = reinterpret_cast<const double2 *>(pose + 8 * (size_t)v);
    const double2 a = s[0], b = s[1], c = s[2], d = s[3];
    p = V3(a.x, a.y, b.x);
    q = Q4(b.y, c.x, c.y, d.x);
    tempe = (float)d.y;
}
__device__ __forceinline__ void store_pose(double *pose, int v, const V3 &p, const Q4 &q, float tempe_next) {
    double2 *s = reinterpret_cast<double2 *>(pose + 8 * (size_t)v);
    s[0] = make_double2(p.x, p.y);
    s[1] = make_double2(p.z, q.w);
    s[2] = make_double2(q.x, q.y);
    s[3] = make_double2(q.z, (double)tempe_next);
}
__device__ __forceinline__ V3 load3(const double *__restrict__ a, size_t i) { return V3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
__device__ __forceinline__ void store3(double *a, size_t i, const V3 &v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }

// ------------------------------------------------------------------ links
// gpu_update_links (VX3_VoxelyzeKernel.cu:566-581) with the temperature-driven rest-length refresh of
// gpu_update_temperature (:625-650) folded in.
__global__ void __launch_bounds__(VX3_BLOCK, VX3_LINKS_MIN_CTAS) k_links(Dev D) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= D.nlinkslots) return;
    // Loads are issued in dependency LEVELS, each level before any branch that could separate it from the next:
    // the kernel is latency bound (ncu: long-scoreboard stalls dominate), so the number of serialised DRAM round
    // trips matters more than a few loads wasted on skipped links.
    // ---- level 1: everything indexed by the link slot ----
    const int2 e = D.lends[g];
    LinkRegs L;
    L.state = D.lstate[g];
    const int lmi = D.lmat[g];
    const double *h = D.lhist + 9 * (size_t)g;
    L.pos2 = V3(h[0], h[1], h[2]);
    L.angle1v = V3(h[3], h[4], h[5]);
    L.angle2v = V3(h[6], h[7], h[8]);
    const float4 sn = D.lstrain[g];
    const float2 ar = D.larea[g];
    L.rest = D.lrest[g];
    if (e.x < 0) return;
    // ---- level 2: everything indexed by the two end voxels (+ the link material) ----
    V3 pN, pP;
    Q4 qN, qP;
    float tN, tP; // the ends' temperatures for this step (computed by the voxel pass of the previous step)
    load_pose_t(D.pose, e.x, pN, qN, tN);
    load_pose_t(D.pose, e.y, pP, qP, tP);
    const int vmN = D.vmat[e.x], vmP = D.vmat[e.y];
    const float pdN = D.prevdt[e.x], pdP = D.prevdt[e.y];
    const int sim = D.nsims == 1 ? 0 : D.vsim[e.x];
    const LinkMatC &lm = D.lmat_tab[lmi];
    // ---- level 3: small tables (L1/L2 resident) ----
    const SimC &S = D.simc[sim];
    SimD &dy = D.simd[sim];
    const VoxMatC &mN = D.vmat_tab[vmN], &mP = D.vmat_tab[vmP];
    const int status = dy.status;
    const float dt = dy.dt;
    const bool fixedBoth = mN.fixed && mP.fixed;
    const float numN = mN.dampMultNum, numP = mP.dampMultNum;
    if (L.state & (LKS_DETACHED | LKS_REMOVED)) return;
    if (status != VX3_SIM_RUNNING || dt == 0 || fixedBoth) return;
    const int axis = (L.state & LKS_AXIS_MASK) >> LKS_AXIS_SHIFT;
    L.state &= ~LKS_JUST_CREATED;
    L.strain = sn.x; L.maxStrain = sn.y; L.strainOffset = sn.z; L.stress = sn.w;
    L.area = ar.x; L.tsum = ar.y;
    if (S.vary_temp && S.temp_period > 0) { // updateRestLength() from either end's setTemperature (VX3_Voxel.cu:107-113)
        const double t = dy.t;
        if (thermal_active(S, mN, 0, t) || thermal_active(S, mP, 0, t)) {
            L.rest = 0.5 * (base_size_axis(mN, tN, axis) + base_size_axis(mP, tP, axis));
            D.lrest[g] = L.rest;
        }
    }
    // dampingMultiplier() = 2*_sqrtMass*zetaInternal/previousDt (float)
    const float dmN = numN / pdN, dmP = numP / pdP;
    LinkOut o;
    link_update_forces(L, lm, D.strain_pool, D.stress_pool, pN, qN, pP, qP, dmN, dmP, o);
    double *hw = D.lhist + 9 * (size_t)g;
    hw[0] = L.pos2.x; hw[1] = L.pos2.y; hw[2] = L.pos2.z;
    hw[3] = L.angle1v.x; hw[4] = L.angle1v.y; hw[5] = L.angle1v.z;
    hw[6] = L.angle2v.x; hw[7] = L.angle2v.y; hw[8] = L.angle2v.z;
    D.lstrain[g] = make_float4(L.strain, L.maxStrain, L.strainOffset, L.stress);
    D.lstate[g] = L.state;
    double2 *f = reinterpret_cast<double2 *>(D.lforce + 12 * (size_t)g);
    f[0] = make_double2(o.forceNeg.x, o.forceNeg.y);
    f[1] = make_double2(o.forceNeg.z, o.momentNeg.x);
    f[2] = make_double2(o.momentNeg.y, o.momentNeg.z);
    f[3] = make_double2(o.forcePos.x, o.forcePos.y);
    f[4] = make_double2(o.forcePos.z, o.momentPos.x);
    f[5] = make_double2(o.momentPos.y, o.momentPos.z);
    // divergence: the reference samples one random link per step (:273-280); every link is checked here
    if (L.strain > 100) dy.diverged = 1;
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
    return mt_eval<VX3_DEV_MAX_TOKENS>(D.tokens + S.prog_off[slot], S.prog_n[slot], vars, &ok);
}

// gpu_update_voxels (VX3_VoxelyzeKernel.cu:582-623) -> VX3_Voxel::timeStep
__global__ void __launch_bounds__(VX3_BLOCK, VX3_VOXELS_MIN_CTAS) k_voxels(Dev D) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= D.nvox) return;
    // loads in dependency levels (see k_links)
    // ---- level 1: everything indexed by the voxel ----
    VoxRegs r;
    float tempe;
    load_pose_t(D.pose, v, r.pos, r.orient, tempe); // this step's temperature (gpu_update_temperature at time t)
    r.flags = D.vflags[v];
    const int vmi = D.vmat[v];
    const int sim = D.nsims == 1 ? 0 : D.vsim[v];
    const double *mo = D.mom + 6 * (size_t)v;
    r.linMom = V3(mo[0], mo[1], mo[2]);
    r.angMom = V3(mo[3], mo[4], mo[5]);
    int vl[6];
#pragma unroll
    for (int i = 0; i < 6; i++) vl[i] = D.vlinks[6 * (size_t)v + i];
    const double phase = D.phase[v];
    // ---- level 2: link end forces, tables ----
    double2 fa[6], fb[6], fc[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        if (vl[i] >= 0) {
            const double2 *f = reinterpret_cast<const double2 *>(D.lforce + 12 * (size_t)vl[i] + ((i & 1) ? 6 : 0));
            fa[i] = f[0]; fb[i] = f[1]; fc[i] = f[2];
        }
    }
    const SimC &S = D.simc[sim];
    const SimD &dy = D.simd[sim];
    const VoxMatC &m = D.vmat_tab[vmi];
    if (dy.status != VX3_SIM_RUNNING || dy.diverged) return;
    const float dtF = dy.dt;
    if (dtF == 0) return;
    const double dt = dtF, t = dy.t;
    D.tempe[v] = tempe;
    // temperature the next step will start with (time t + dt), see store_pose
    const double tnext = t + dtF;
    const float tempe_next = thermal_active(S, m, r.flags, tnext) ? voxel_temperature(S, tnext, phase) : tempe;
    if ((r.flags & VXF_REMOVED) || m.fixed) {
        if (tempe_next != tempe) D.pose[8 * (size_t)v + 7] = (double)tempe_next;
        return;
    }
    D.prevdt[v] = (float)dt;
    V3 F(0, 0, 0), M(0, 0, 0);
#pragma unroll
    for (int i = 0; i < 6; i++) {
        if (vl[i] >= 0) {
            F += V3(fa[i].x, fa[i].y, fb[i].x);
            M += V3(fb[i].y, fc[i].x, fc[i].y);
        }
    }
    V3 contact(0, 0, 0);
    if (D.contact) {
        contact = load3(D.contact, v);
        store3(D.contact, v, V3());
    }
    V3 cil(0, 0, 0);
    if (S.enable_cilia && !(r.flags & VX3_VOX_SURFACE) && m.cilia != 0 && !(m.cilia_on_after > t)) { // gpu_update_cilia_force :846-859
        cil = r.orient.RotateVec3D(load3(D.base_cilia, v)) * m.cilia;                                 // localSignal = 0 (signals are off)
    }
    V3 ff(0, 0, 0);
    const int ext = D.vext[v];
    const ExtC *px = ext >= 0 ? &D.exts[ext] : nullptr;
    const bool fixedAll = px && (px->dof & 0x3F) == 0x3F;
    if (S.has_ff && !fixedAll) {
        double vars[9];
        prog_vars(S, dy, t, r.pos.x, r.pos.y, r.pos.z, vars);
        ff.x = eval_slot(D, S, VX3_PROG_FORCE_X, vars, 0.0);
        ff.y = eval_slot(D, S, VX3_PROG_FORCE_Y, vars, 0.0);
        ff.z = eval_slot(D, S, VX3_PROG_FORCE_Z, vars, 0.0);
    }
    const short *ic = D.ixyz + 3 * (size_t)v;
    voxel_time_step(r, m, px, ic[0], ic[1], ic[2], tempe, F, M, contact, cil, ff, dt);
    // enableAttach = AND of the five attach conditions at the new position (:609-621)
    if (S.has_attach_cond) {
        double vars[9];
        prog_vars(S, dy, t, r.pos.x, r.pos.y, r.pos.z, vars);
        bool all = true;
        for (int c = 0; c < 5 && all; c++) all = eval_slot(D, S, VX3_PROG_ATTACH_0 + c, vars, 1.0) > 0;
        if (all) r.flags |= VXF_ENABLE_ATTACH;
        else r.flags &= ~VXF_ENABLE_ATTACH;
    }
    store_pose(D.pose, v, r.pos, r.orient, tempe_next);
    double *mw = D.mom + 6 * (size_t)v;
    mw[0] = r.linMom.x; mw[1] = r.linMom.y; mw[2] = r.linMom.z;
    mw[3] = r.angMom.x; mw[4] = r.angMom.y; mw[5] = r.angMom.z;
    D.vflags[v] = r.flags;
}

// ------------------------------------------------------------------ collision grid
__device__ __forceinline__ unsigned cell_hash(int sim, int cx, int cy, int cz) {
    return ((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u) ^ ((unsigned)cz * 83492791u) ^ ((unsigned)sim * 2654435761u);
}
__device__ __forceinline__ bool sim_collides(const SimC &S) { return S.enable_collision || S.enable_attach; }

// regenerateSurfaceVoxels (:495-513) + uniform-grid insert (replaces the O(S^2) sweep of gpu_update_attach :833-843).
// Also publishes this step's temperature (what updateTemperature :219-235 set) for the contact phase.
__global__ void __launch_bounds__(VX3_BLOCK) k_grid_count(Dev D) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= D.nvox) return;
    const int sim = D.vsim[v];
    const SimC &S = D.simc[sim];
    const SimD &dy = D.simd[sim];
    int4 vc = make_int4(0, 0, 0, -1);
    if (dy.status == VX3_SIM_RUNNING && !dy.diverged && dy.dt != 0 && sim_collides(S)) {
        int flags = D.vflags[v];
        D.tempe[v] = (float)D.pose[8 * (size_t)v + 7]; // this step's temperature, for the contact phase
        bool interior = true; // VX3_Voxel::updateSurface (VX3_Voxel.cu:515-524): the bit named SURFACE means interior
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const int li = D.vlinks[6 * (size_t)v + i];
            if (li < 0 || (D.lstate[li] & LKS_DETACHED)) interior = false;
        }
        const int nf = interior ? (flags | VX3_VOX_SURFACE) : (flags & ~VX3_VOX_SURFACE);
        if (nf != flags) D.vflags[v] = nf;
        if (!interior && !(flags & VXF_REMOVED)) {
            const V3 p = load_pos(D.pose, v);
            vc.x = (int)floor(p.x * S.cell_inv);
            vc.y = (int)floor(p.y * S.cell_inv);
            vc.z = (int)floor(p.z * S.cell_inv);
            vc.w = (int)(cell_hash(sim, vc.x, vc.y, vc.z) & (unsigned)D.hmask);
            atomicAdd(&D.cell_cnt[vc.w], 1);
        }
    }
    D.vcell[v] = vc;
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

// exclusive scan of the bucket counts (single CTA, int4-vectorised), resets the counters for the next step
__global__ void __launch_bounds__(1024) k_grid_scan(Dev D) {
    __shared__ int part[1024];
    const int H = D.hmask + 1;            // power of two >= 1024
    const int per = H / 1024;             // buckets per thread (multiple of 4 when H >= 4096)
    const int b0 = threadIdx.x * per;
    int s = 0;
    if (per >= 4) {
        const int4 *c4 = reinterpret_cast<const int4 *>(D.cell_cnt + b0);
        for (int i = 0; i < per / 4; i++) {
            const int4 c = c4[i];
            s += c.x + c.y + c.z + c.w;
        }
    } else
        for (int i = 0; i < per; i++) s += D.cell_cnt[b0 + i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) { // Hillis-Steele inclusive scan
        int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    if (per >= 4) {
        int4 *c4 = reinterpret_cast<int4 *>(D.cell_cnt + b0);
        int4 *s4 = reinterpret_cast<int4 *>(D.cell_start + b0);
        int4 *u4 = reinterpret_cast<int4 *>(D.cell_cursor + b0);
        for (int i = 0; i < per / 4; i++) {
            const int4 c = c4[i];
            int4 o;
            o.x = run; o.y = o.x + c.x; o.z = o.y + c.y; o.w = o.z + c.z;
            run = o.w + c.w;
            s4[i] = o;
            u4[i] = o;
            c4[i] = make_int4(0, 0, 0, 0);
        }
    } else
        for (int i = 0; i < per; i++) {
            const int c = D.cell_cnt[b0 + i];
            D.cell_start[b0 + i] = run;
            D.cell_cursor[b0 + i] = run;
            D.cell_cnt[b0 + i] = 0;
            run += c;
        }
    if (threadIdx.x == 1023) D.cell_start[H] = part[1023];
    if (threadIdx.x == 0) *D.cand_count = 0;
}

__global__ void __launch_bounds__(VX3_BLOCK) k_grid_fill(Dev D) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= D.nvox) return;
    const int b = D.vcell[v].w;
    if (b < 0) return;
    D.cell_items[atomicAdd(&D.cell_cursor[b], 1)] = v;
}

// is_neighbor (VX3_VoxelyzeKernel.cu:651-680), iterative
__device__ bool is_neighbor(const Dev &D, int v1, int v2, int depth) {
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
        const V3 vel1 = load3(D.mom, 2 * (size_t)hi) * m1.massInverse, vel2 = load3(D.mom, 2 * (size_t)lo) * m2.massInverse;
        const double relativeVelocity = vel1.Dot(unit) - vel2.Dot(unit);
        return unit * (penetrationStiff * RelDist + dampingC * relativeVelocity);
    }
    return V3(0, 0, 0);
}

// Contact phase of one surface voxel: all partners inside the collision envelope
// (handle_collision_attachment, VX3_VoxelyzeKernel.cu:682-727), accumulated in ascending partner index = the
// canonical sequential pair order.  emit = also count target hits and emit attach candidates
// for the pairs this voxel leads (it is the higher index).
__device__ __noinline__ void contact_phase(const Dev &D, int v, bool emit) {
    const int sim = D.vsim[v];
    const SimC &S = D.simc[sim];
    SimD &dy = D.simd[sim];
    const int4 vc = D.vcell[v];
    const int matv = D.vmat[v];
    const VoxMatC &mv = D.vmat_tab[matv];
    const V3 pv = load_pos(D.pose, v);
    const double bsv = base_size_average(mv, D.tempe[v]);
    int vl[6], vo[6]; // own links and their other ends
#pragma unroll
    for (int i = 0; i < 6; i++) {
        vl[i] = D.vlinks[6 * (size_t)v + i];
        vo[i] = -1;
        if (vl[i] >= 0) {
            const int2 e = D.lends[vl[i]];
            vo[i] = (e.x == v) ? e.y : e.x;
        }
    }
    int partner[VX3_MAX_PARTNERS];
    int np = 0;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy_ = -1; dy_ <= 1; dy_++)
            for (int dx = -1; dx <= 1; dx++) {
                const int cx = vc.x + dx, cy = vc.y + dy_, cz = vc.z + dz;
                const int b = (int)(cell_hash(sim, cx, cy, cz) & (unsigned)D.hmask);
                const int s0 = D.cell_start[b], s1 = D.cell_start[b + 1];
                for (int k = s0; k < s1; k++) {
                    const int u = D.cell_items[k];
                    if (u == v) continue;
                    const int4 uc = D.vcell[u];
                    if (uc.x != cx || uc.y != cy || uc.z != cz || D.vsim[u] != sim) continue; // other cell hashed to this bucket
                    const VoxMatC &mu = D.vmat_tab[D.vmat[u]];
                    if (mv.fixed && mu.fixed) continue;
                    const V3 pu = load_pos(D.pose, u);
                    const V3 diff = (v > u) ? (pv - pu) : (pu - pv); // voxel1 - voxel2, voxel1 = higher index
                    const double bsu = base_size_average(mu, D.tempe[u]);
                    const double watch = ((v > u) ? (bsv + bsu) : (bsu + bsv)) * VX3_COLLISION_ENVELOPE_RADIUS;
                    if (diff.x > watch || diff.x < -watch) continue;
                    if (diff.y > watch || diff.y < -watch) continue;
                    if (diff.z > watch || diff.z < -watch) continue;
                    if (diff.Length() > watch) continue;
                    // direct lattice neighbours are skipped (is_neighbor depth 1) unless the link was made this step
                    bool linked = false, fresh = false;
#pragma unroll
                    for (int i = 0; i < 6; i++)
                        if (vo[i] == u) {
                            linked = true;
                            if (D.lstate[vl[i]] & LKS_JUST_CREATED) fresh = true;
                        }
                    if (linked && !fresh) continue;
                    if (np < VX3_MAX_PARTNERS) partner[np++] = fresh ? (u | (1 << 30)) : u;
                    else dy.err = VX3_ERR_CAPACITY;
                }
            }
    // ascending partner index (insertion sort; np is small)
    for (int i = 1; i < np; i++) {
        const int key = partner[i];
        int j = i - 1;
        while (j >= 0 && (partner[j] & 0x3FFFFFFF) > (key & 0x3FFFFFFF)) {
            partner[j + 1] = partner[j];
            j--;
        }
        partner[j + 1] = key;
    }
    V3 c(0, 0, 0);
    int hits = 0;
    for (int i = 0; i < np; i++) {
        const int u = partner[i] & 0x3FFFFFFF;
        const bool fresh = (partner[i] >> 30) & 1;
        const int hi = v > u ? v : u, lo = v > u ? u : v;
        const VoxMatC &m1 = D.vmat_tab[D.vmat[hi]], &m2 = D.vmat_tab[D.vmat[lo]];
        if (S.enable_collision) {
            V3 f = pair_contact_force(D, hi, lo, m1, m2);
            if (v != hi) f = -f;
            c += f;
            if (fresh) c -= f; // a link was created for this pair: its contact force is taken back (:827-830)
            if (v == hi && ((m1.is_target && !m2.is_target) || (m2.is_target && !m1.is_target))) hits++;
        }
        if (!emit || v != hi || fresh) continue;
        // ---- attach candidate test (:729-812) on the step-start link graph ----
        const int fl = D.vflags[lo], fh = D.vflags[hi];
        if (!(fh & VXF_ENABLE_ATTACH) || !(fl & VXF_ENABLE_ATTACH)) continue;
        if (m1.fixed || m2.fixed) continue;
        if (D.vmat[hi] != D.vmat[lo]) continue;
        if (!m1.sticky) continue;
        V3 p1, p2;
        Q4 q1, q2;
        load_pose(D.pose, hi, p1, q1);
        p2 = load_pos(D.pose, lo);
        const V3 e = p1 - p2;
        const V3 ea = q1.RotateVec3DInv(-e);
        const V3 f = ea.Abs();
        int dir1, dir2, axis, rev = 0;
        if (f.x >= f.y && f.x >= f.z) {
            axis = 0;
            if (ea.x < 0) { dir1 = 1; dir2 = 0; rev = 1; } else { dir1 = 0; dir2 = 1; }
        } else if (f.y >= f.x && f.y >= f.z) {
            axis = 1;
            if (ea.y < 0) { dir1 = 3; dir2 = 2; rev = 1; } else { dir1 = 2; dir2 = 3; }
        } else {
            axis = 2;
            if (ea.z < 0) { dir1 = 5; dir2 = 4; rev = 1; } else { dir1 = 4; dir2 = 5; }
        }
        // slots only fill up during the attach phase, so an occupied slot now stays a rejection at this pair's turn
        if (D.vlinks[6 * (size_t)hi + dir1] >= 0 || D.vlinks[6 * (size_t)lo + dir2] >= 0) continue;
        if (is_neighbor(D, hi, lo, 5)) continue; // links are only added during the phase: true now stays true
        const int slot = atomicAdd(D.cand_count, 1);
        if (slot < D.cand_cap) {
            Cand cd;
            cd.key = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
            cd.info = dir1 | (dir2 << 3) | (axis << 6) | (rev << 8);
            cd._pad = 0;
            D.cands[slot] = cd;
        }
    }
    if (S.enable_collision) {
        store3(D.contact, v, c);
        if (emit && hits) atomicAdd(&dy.collision_count, hits);
    }
}

__global__ void __launch_bounds__(128) k_contact(Dev D) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= D.nvox) return;
    if (D.vcell[v].w < 0) return;
    contact_phase(D, v, true);
}

// Sequential resolution of the attach candidates in canonical (first, second) order: a candidate is accepted
// only if both facing slots are still empty and the voxels are still not within 5 links at its turn
// (SURVEY.md A.7).  Creates the link like VX3_Link's device ctor + reset() (VX3_Link.cu:31-70).
__global__ void __launch_bounds__(1024) k_resolve(Dev D) {
    __shared__ unsigned long long skey[2048];
    __shared__ int sinfo[2048];
    int n = *D.cand_count;
    if (n == 0) return;
    if (n > D.cand_cap || n > 2048) {
        if (threadIdx.x == 0)
            for (int s = 0; s < D.nsims; s++) D.simd[s].err = VX3_ERR_CAPACITY;
        n = min(n, min(D.cand_cap, 2048));
    }
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        skey[i] = i < n ? D.cands[i].key : ~0ull;
        sinfo[i] = i < n ? D.cands[i].info : 0;
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) // bitonic sort
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    if ((skey[i] > skey[ixj]) == up) {
                        const unsigned long long tk = skey[i]; skey[i] = skey[ixj]; skey[ixj] = tk;
                        const int ti = sinfo[i]; sinfo[i] = sinfo[ixj]; sinfo[ixj] = ti;
                    }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x != 0) return;
    for (int c = 0; c < n; c++) {
        const int hi = (int)(skey[c] >> 32), lo = (int)(skey[c] & 0xFFFFFFFFu);
        const int info = sinfo[c];
        const int dir1 = info & 7, dir2 = (info >> 3) & 7, axis = (info >> 6) & 3, rev = (info >> 8) & 1;
        if (D.vlinks[6 * (size_t)hi + dir1] >= 0 || D.vlinks[6 * (size_t)lo + dir2] >= 0) continue;
        if (is_neighbor(D, hi, lo, 5)) continue;
        const int sim = D.vsim[hi];
        const SimC &S = D.simc[sim];
        SimD &dy = D.simd[sim];
        if (dy.link_cnt >= S.lcap) { dy.err = VX3_ERR_CAPACITY; continue; }
        const VoxMatC &mh = D.vmat_tab[D.vmat[hi]];
        if (mh.self_lmat < 0) { dy.err = VX3_ERR_INVALID; continue; }
        const int g = S.loff + dy.link_cnt++;
        const int vneg = rev ? lo : hi, vpos = rev ? hi : lo; // pVNeg/pVPos of VX3_Link(voxelA, dirA, voxelB, dirB)
        D.vlinks[6 * (size_t)hi + dir1] = g;
        D.vlinks[6 * (size_t)lo + dir2] = g;
        D.lends[g] = make_int2(vneg, vpos);
        D.lmat[g] = mh.self_lmat;
        D.lstate[g] = (axis << LKS_AXIS_SHIFT) | LKS_SMALL | LKS_JUST_CREATED | (S.safety_guard << LKS_NEWLINK_SHIFT);
        double *h = D.lhist + 9 * (size_t)g;
        for (int k = 0; k < 9; k++) h[k] = 0.0;
        double *f = D.lforce + 12 * (size_t)g;
        for (int k = 0; k < 12; k++) f[k] = 0.0;
        D.lstrain[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        const VoxMatC &mn = D.vmat_tab[D.vmat[vneg]], &mp = D.vmat_tab[D.vmat[vpos]];
        D.lrest[g] = 0.5 * (base_size_axis(mn, D.tempe[vneg], axis) + base_size_axis(mp, D.tempe[vpos], axis));
        const float sn = (float)mn.nomSize, sp = (float)mp.nomSize; // transverseArea() with zero strain
        D.larea[g] = make_float2(0.5f * (sn * sn + sp * sp), 0.0f);
        dy.attach_events++;
        __threadfence_block();
        if (S.enable_collision) { // take this pair's contact force back in sequence position (:827-830)
            contact_phase(D, hi, false);
            contact_phase(D, lo, false);
        }
    }
}

// gpu_update_detach (VX3_VoxelyzeKernel.cu:946-968)
__global__ void __launch_bounds__(VX3_BLOCK) k_detach(Dev D) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= D.nlinkslots) return;
    const int2 e = D.lends[g];
    if (e.x < 0) return;
    const int st = D.lstate[g];
    if (st & (LKS_DETACHED | LKS_REMOVED)) return;
    const int sim = D.vsim[e.x];
    const SimC &S = D.simc[sim];
    SimD &dy = D.simd[sim];
    if (!S.enable_detach || dy.status != VX3_SIM_RUNNING || dy.diverged || dy.dt == 0) return;
    const LinkMatC &lm = D.lmat_tab[D.lmat[g]];
    if (!mat_failed(lm, D.lstrain[g].y)) return;
    D.lstate[g] = st | LKS_DETACHED;
    for (int i = 0; i < 6; i++) {
        if (D.vlinks[6 * (size_t)e.x + i] == g) D.vlinks[6 * (size_t)e.x + i] = -1;
        if (D.vlinks[6 * (size_t)e.y + i] == g) D.vlinks[6 * (size_t)e.y + i] = -1;
    }
    atomicAdd(&dy.detach_events, 1);
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
        for (int k = 0; k < 6; k++) {
            const int li = D.vlinks[6 * (size_t)v + k];
            if (li < 0) continue;
            D.lstate[li] |= LKS_REMOVED; // (a neighbour removed in the same step sets the same bit: benign)
            const int2 e = D.lends[li];
            const int nb = (e.x == v) ? e.y : e.x;
            for (int q = 0; q < 6; q++)
                if (D.vlinks[6 * (size_t)nb + q] == li) {
                    D.vlinks[6 * (size_t)nb + q] = -1;
                    break;
                }
            D.vlinks[6 * (size_t)v + k] = -1;
        }
    }
    if (!dy.initpos_reinitialized && S.reinit_after < t) store3(D.initpos, v, load_pos(D.pose, v)); // saveInitialPosition()
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
        if (!m.is_measured) continue;
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
    D.pose[8 * (size_t)v + 7] = (double)tempe;
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
