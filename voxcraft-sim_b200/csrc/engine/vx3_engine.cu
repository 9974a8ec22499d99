// C ABI of the engine (include/vx3_abi.h): batch construction, the step/run drivers and read-back.
// Host orchestration only — all physics is in the kernels (vx3_kernels.cuh, vx3_persistent.cuh).
// There is NO CPU fallback: without a usable sm_100 device every entry point fails with VX3_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../../include/vx3_abi.h"
#include "../../../include/vx3_model.h"
#include "vx3_kernels.cuh"
#include "../host/vx3_materials.h"
#include "vx3_persistent.cuh"
#include "vx3_halo.cuh"
#include "vx3_fused.cuh"
#include "vx3_history.h"

using namespace vx3;

static inline float __int_as_float_host(int i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CK(call)                                                                                                        \
    do {                                                                                                                \
        cudaError_t e_ = (call);                                                                                        \
        if (e_ != cudaSuccess) return fail(VX3_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
    } while (0)

extern "C" const char *vx3_last_error(void) { return g_err.c_str(); }
extern "C" int vx3_abi_version(void) { return VX3_ABI_VERSION; }
extern "C" size_t vx3_abi_sizeof(const char *name) {
    if (!name) return 0;
#define SZ(T) if (!strcmp(name, #T)) return sizeof(T)
    SZ(vx3_token); SZ(vx3_program); SZ(vx3_voxel_material); SZ(vx3_link_material); SZ(vx3_external); SZ(vx3_sim_options);
    SZ(vx3_model_desc); SZ(vx3_result); SZ(vx3_state_view); SZ(vx3_run_opts); SZ(vx3_material_params); SZ(vx3_env_params);
#undef SZ
    return 0;
}

// ------------------------------------------------------------------ per-kernel timing (bench hook)
enum KernelClass { KC_LINKS = 0, KC_VOXELS, KC_GRID_BUILD, KC_CONTACT, KC_RESOLVE, KC_DETACH, KC_SURFACE, KC_SECONDARY, KC_SIGNALS, KC_COM, KC_TAIL, KC_PERSISTENT, KC_HALO, KC_FUSED, KC_LINKS_FACE, KC_COUNT };
static const char *const kKernelNames[KC_COUNT] = {"k_links", "k_voxels", "k_grid_build", "k_contact", "k_resolve",
                                                   "k_detach", "k_surface", "k_secondary", "k_signals", "k_com_partial", "k_tail", "k_persistent", "k_halo", "k_fused", "k_links_face"};
struct Profiler {
    bool on = false;
    std::vector<cudaEvent_t> ev; // pairs
    std::vector<int> cls;
    size_t used = 0;
    double ms[KC_COUNT] = {};
    long long cnt[KC_COUNT] = {};
    bool begin(int c, cudaStream_t st) {
        if (!on || used + 2 > ev.size()) return false;
        cls.push_back(c);
        cudaEventRecord(ev[used], st);
        return true;
    }
    void end(cudaStream_t st) {
        cudaEventRecord(ev[used + 1], st);
        used += 2;
    }
    void collect() { // after the stream has been synchronised
        for (size_t i = 0; i + 1 < used; i += 2) {
            float t = 0;
            if (cudaEventElapsedTime(&t, ev[i], ev[i + 1]) == cudaSuccess) {
                ms[cls[i / 2]] += t;
                cnt[cls[i / 2]]++;
            }
        }
        used = 0;
        cls.clear();
    }
};

// ------------------------------------------------------------------ device arena + resource cache
// One batch = ONE device allocation and ONE host->device copy: every array of the batch is a 256-byte aligned slice of
// an arena, staged in one pinned host buffer and uploaded with a single cudaMemcpyAsync (the reference pays one
// cudaMalloc + one blocking cudaMemcpy per voxel/link/material object, VX3_VoxelyzeKernel.cu:107-160).  Arenas, their
// pinned staging buffers, streams and events are kept in a small per-process cache when a batch is destroyed, so that a
// worker which evaluates one batch after another (vx3_node_worker) pays cudaMalloc/cudaFree/cudaMallocHost once.
struct Resources {
    int device = -1;
    char *d = nullptr; // device arena
    char *h = nullptr; // pinned staging
    size_t dcap = 0, hcap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
static std::mutex g_res_mu;
static std::vector<Resources> g_res_idle;
static const size_t kMaxIdlePerDevice = 2;

static void resources_free(Resources &r) {
    if (r.d) cudaFree(r.d);
    if (r.h) cudaFreeHost(r.h);
    if (r.ev0) cudaEventDestroy(r.ev0);
    if (r.ev1) cudaEventDestroy(r.ev1);
    if (r.stream) cudaStreamDestroy(r.stream);
    r = Resources();
}

// the current device must be `device`
static int resources_acquire(int device, size_t dbytes, size_t hbytes, Resources &out) {
    out = Resources();
    {
        std::lock_guard<std::mutex> lk(g_res_mu);
        int best = -1;
        for (size_t i = 0; i < g_res_idle.size(); i++) {
            const Resources &r = g_res_idle[i];
            if (r.device != device) continue;
            if (best < 0) best = (int)i;
            else { // prefer one that already fits, then the smallest such; otherwise the largest (least to regrow)
                const Resources &q = g_res_idle[best];
                const bool rf = r.dcap >= dbytes && r.hcap >= hbytes, qf = q.dcap >= dbytes && q.hcap >= hbytes;
                if ((rf && !qf) || (rf == qf && (rf ? r.dcap < q.dcap : r.dcap > q.dcap))) best = (int)i;
            }
        }
        if (best >= 0) {
            out = g_res_idle[best];
            g_res_idle.erase(g_res_idle.begin() + best);
        }
    }
    out.device = device;
    if (!out.stream) {
        CK(cudaStreamCreateWithFlags(&out.stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&out.ev0));
        CK(cudaEventCreate(&out.ev1));
    }
    if (out.dcap < dbytes) {
        if (out.d) cudaFree(out.d);
        out.d = nullptr;
        out.dcap = 0;
        const size_t want = (dbytes + (dbytes >> 3) + ((size_t)1 << 20)) & ~(((size_t)1 << 20) - 1); // headroom so similar batches reuse it
        cudaError_t e = cudaMalloc((void **)&out.d, want);
        if (e != cudaSuccess) {
            resources_free(out);
            return fail(VX3_ERR_CUDA, std::string("cudaMalloc (batch arena): ") + cudaGetErrorString(e));
        }
        out.dcap = want;
    }
    if (out.hcap < hbytes) {
        if (out.h) cudaFreeHost(out.h);
        out.h = nullptr;
        out.hcap = 0;
        const size_t want = (hbytes + (hbytes >> 3) + ((size_t)1 << 20)) & ~(((size_t)1 << 20) - 1);
        cudaError_t e = cudaMallocHost((void **)&out.h, want);
        if (e != cudaSuccess) {
            resources_free(out);
            return fail(VX3_ERR_CUDA, std::string("cudaMallocHost (staging): ") + cudaGetErrorString(e));
        }
        out.hcap = want;
    }
    return VX3_OK;
}

// grows the device arena of an acquired set (nothing has been uploaded into it yet)
static int resources_grow_device(Resources &r, size_t dbytes) {
    if (r.dcap >= dbytes) return VX3_OK;
    if (r.d) cudaFree(r.d);
    r.d = nullptr;
    r.dcap = 0;
    const size_t want = (dbytes + (dbytes >> 3) + ((size_t)1 << 20)) & ~(((size_t)1 << 20) - 1);
    cudaError_t e = cudaMalloc((void **)&r.d, want);
    if (e != cudaSuccess) return fail(VX3_ERR_CUDA, std::string("cudaMalloc (batch arena): ") + cudaGetErrorString(e));
    r.dcap = want;
    return VX3_OK;
}

static void resources_release(Resources &r) {
    if (!r.stream && !r.d && !r.h) return;
    std::vector<Resources> drop;
    {
        std::lock_guard<std::mutex> lk(g_res_mu);
        g_res_idle.push_back(r);
        size_t same = 0;
        for (const Resources &q : g_res_idle) same += q.device == r.device;
        while (same > kMaxIdlePerDevice) { // drop the smallest idle arena of this device
            int worst = -1;
            for (size_t i = 0; i < g_res_idle.size(); i++)
                if (g_res_idle[i].device == r.device && (worst < 0 || g_res_idle[i].dcap < g_res_idle[worst].dcap)) worst = (int)i;
            drop.push_back(g_res_idle[worst]);
            g_res_idle.erase(g_res_idle.begin() + worst);
            same--;
        }
    }
    r = Resources();
    for (Resources &q : drop) {
        cudaSetDevice(q.device);
        resources_free(q);
    }
}

// frees every cached arena / staging buffer / stream (idle ones only; live batches keep theirs)
extern "C" void vx3_engine_trim(void) {
    std::vector<Resources> all;
    {
        std::lock_guard<std::mutex> lk(g_res_mu);
        all.swap(g_res_idle);
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (Resources &q : all) {
        cudaSetDevice(q.device);
        resources_free(q);
    }
    cudaSetDevice(cur);
    {
        PersistentPlanCache &pc = PersistentPlanCache::get();
        std::lock_guard<std::mutex> lk(pc.mu);
        pc.entries.clear();
    }
}

// std::vector without the serial zero fill of resize(): the batch builder's threads write every element themselves
template <class T> struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    NoInitAlloc() = default;
    template <class U> NoInitAlloc(const NoInitAlloc<U> &) {}
    template <class U, class... A> void construct(U *p, A &&...a) {
        if constexpr (sizeof...(A) == 0) ::new ((void *)p) U; // default-init: no write for trivial types
        else ::new ((void *)p) U(std::forward<A>(a)...);
    }
};
template <class T> using RawVec = std::vector<T, NoInitAlloc<T>>;

// slices of the arena, planned first (sizes), then staged + uploaded in one go
struct ArenaPlan {
    struct Item {
        void **field;
        const void *src;
        size_t bytes, off;
    };
    std::vector<Item> up, zero;
    template <class T, class U, class Al> void upload(T **field, const std::vector<U, Al> &h) {
        static_assert(sizeof(T) == sizeof(U), "element size");
        if (h.empty()) zero.push_back(Item{(void **)field, nullptr, sizeof(T), 0});
        else up.push_back(Item{(void **)field, h.data(), h.size() * sizeof(U), 0});
    }
    template <class T> void zeroed(T **field, size_t n) { zero.push_back(Item{(void **)field, nullptr, std::max<size_t>(n, 1) * sizeof(T), 0}); }
    // arrays that were built in place in the staging buffer: same offset in the arena, nothing to copy
    std::vector<Item> placed;
    size_t placed_bytes = 0;
    template <class T> T *take(char *staging, T **field, size_t n) {
        const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
        placed.push_back(Item{(void **)field, nullptr, bytes, placed_bytes});
        T *p = reinterpret_cast<T *>(staging + placed_bytes);
        placed_bytes += (bytes + 255) & ~(size_t)255;
        return p;
    }
    size_t up_bytes = 0, total = 0;
    void layout() {
        size_t o = placed_bytes;
        for (Item &i : up) { i.off = o; o += (i.bytes + 255) & ~(size_t)255; }
        up_bytes = o;
        for (Item &i : zero) { i.off = o; o += (i.bytes + 255) & ~(size_t)255; }
        total = std::max<size_t>(o, 256);
    }
};

// ------------------------------------------------------------------ batch object
struct vx3_batch {
    Profiler prof;
    int device = 0;
    Resources res; // arena, pinned staging, stream, events (returned to the cache on destroy)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<void *> allocs; // allocations made after construction (halo buffers)
    Dev D;
    int nsims = 0;
    std::vector<SimC> simc;
    std::vector<std::string> names;
    std::vector<int> n_vmats; // per sim, for reporting
    std::vector<std::vector<int>> lmat_global; // per sim: local link-material index -> global
    std::vector<std::vector<int>> matid;       // per sim: local voxel-material index -> matid
    std::vector<std::vector<int>> vmat_local;  // per sim: voxel -> local material index
    std::vector<std::vector<float>> matcolor;  // per sim: r,g,b,a per material (0..1) for the history header
    std::vector<vx3_sim_options> opts;
    std::vector<std::vector<vx3_voxel_material>> h_vmats; // per sim host copy (data pointers cleared) for the history writer
    std::vector<LinkMatC> h_lmat_tab;
    bool any_collide = false, any_sticky = false, any_detach = false, any_secondary = false, any_signals = false;
    long long hsteps = 0; // doTimeStep calls issued so far (all running simulations advance together)
    std::vector<float> hdt; // per-sim dt in use
    double last_ms = 0;
    long long last_launches = 0;
    long long launches = 0;
    bool link_smtab = true, vox_smtab = true; // material tables fit the kernels' shared-memory copies
    int link_tiles = 0, vox_tiles = 0, link_grid = 1, vox_grid = 1; // persistent tile loops of the streaming kernels
    Halo halo;            // slab decomposition of one body over several GPUs (vx3_batch_halo_*)
    bool any_ghost = false;
    std::vector<double> model_dt; // per simulation: recommendedTimeStep() before the first step (VX3_SimulationManager.cu:56-58)
    bool any_sticky_poisson = false; // an attaching simulation has a sticky material with nu != 0: Dev::pcache
    bool pdl = false; // programmatic dependent launch between the step kernels (launch_pdl; VX3_PDL=1)
    bool tail_in_voxels = false; // plain steps of batches of <= 128 simulations: k_voxels' last CTA does k_tail_light's work (VX3_TAIL_FUSED=0: off)
    size_t vox_active = 0; // voxels [vox_active, nvox) are all ghosts (a slab model lists them last): the voxel pass leaves their tiles out
    int link_queue = -1; // link pass variant: -1 = still being timed (launch_links), 0 = in place, 1 = deferred dense passes
    int lq_trials = 0;
    double lq_ms[2] = {0, 0};
    cudaEvent_t lq_ev[2] = {nullptr, nullptr};
    PersistentPlan pplan; // on-chip path for a single small collision-free body
    bool use_persistent = true;
    FusedPlan fplan;      // fused link + voxel step over spatial blocks for fixed-topology batches (vx3_fused.cuh)
    bool use_fused = true;
    // CUDA-Graph stretches of the streaming path (advance): VX3_GRAPH_STEPS plain steps captured once, replayed while no
    // centre-of-mass sampling step falls inside; one executable graph per value of check_stop
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    long long graph_launches[2] = {0, 0};
    bool graph_failed = false, capturing = false;
    // launch-start snapshot of the arena for the on-chip persistent path (persist_guard): a divergence inside a persistent launch
    // is re-run on the streaming path, which stops every voxel at the reference's step
    size_t arena_bytes = 0;
    unsigned char *snap = nullptr;
    bool snap_valid = false;
    long long snap_hsteps = 0;
    // storage order of a fused batch: model (ABI) index -> device index and back, global indices; empty = identity
    std::vector<int> vperm, lperm, vinv, linv;
    int vdev(size_t ext) const { return vperm.empty() ? (int)ext : vperm[ext]; }
    int ldev(size_t ext) const { return lperm.empty() ? (int)ext : lperm[ext]; }
    int vext(size_t dev) const { return vinv.empty() ? (int)dev : vinv[dev]; }
    int lext(size_t dev) const { return linv.empty() ? (int)dev : linv[dev]; }

    template <class T> int alloc(T **p, size_t n, bool zero = true) {
        *p = nullptr;
        if (n == 0) n = 1;
        cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
        if (e != cudaSuccess) return fail(VX3_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        allocs.push_back(*p);
        if (zero) {
            e = cudaMemsetAsync(*p, 0, n * sizeof(T), stream);
            if (e != cudaSuccess) return fail(VX3_ERR_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e));
        }
        return VX3_OK;
    }
    template <class T> int upload(T **p, const std::vector<T> &h) {
        int rc = alloc(p, h.size(), h.empty());
        if (rc) return rc;
        if (!h.empty()) {
            cudaError_t e = cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) return fail(VX3_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e));
            e = cudaStreamSynchronize(stream); // h may be a temporary
            if (e != cudaSuccess) return fail(VX3_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
        }
        return VX3_OK;
    }
};

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static int setup_stream_kernels(vx3_batch *b, const cudaDeviceProp &prop);

// Device record of a link material (the strain/stress points in the device layout: syncVectors pushes a 0 in front of the
// host data, which itself starts with 0 — VX3_Material.cu:463-477).
static void device_linkmat(const vx3_link_material &in, LinkMatC &lm, std::vector<float> &sd, std::vector<float> &ss) {
    memset(&lm, 0, sizeof(lm));
    lm.E = in.m.E; lm.nu = in.m.nu; lm.eHat = in.m.eHat; lm.epsilonFail = in.m.epsilonFail;
    lm.a1 = in.a1; lm.a2 = in.a2; lm.b1 = in.b1; lm.b2 = in.b2; lm.b3 = in.b3;
    lm.sqA1 = in.sqA1; lm.sqA2xIp = in.sqA2xIp; lm.sqB1 = in.sqB1; lm.sqB2xFMp = in.sqB2xFMp; lm.sqB3xIp = in.sqB3xIp;
    lm.linear = in.m.linear;
    sd.assign(1, 0.0f);
    ss.assign(1, 0.0f);
    for (int k = 0; k < in.m.n_data; k++) {
        sd.push_back(in.m.strain_data ? in.m.strain_data[k] : 0.0f);
        ss.push_back(in.m.stress_data ? in.m.stress_data[k] : 0.0f);
    }
    while (sd.size() < 2) { sd.push_back(0.0f); ss.push_back(0.0f); }
}

static std::string blob(const void *p, size_t n) { return std::string((const char *)p, n); }

static int validate_model(const vx3_model_desc &m, int idx) {
    auto bad = [&](const char *what) { return fail(VX3_ERR_INVALID, "model " + std::to_string(idx) + ": " + what); };
    if (m.n_voxels <= 0) return bad("no voxels");
    if (m.n_links < 0 || m.n_voxel_mats <= 0) return bad("bad counts");
    if (!m.voxel_mats || !m.vox_mat || !m.pos || !m.vox_flags || !m.vox_links || !m.ix || !m.iy || !m.iz) return bad("missing voxel arrays");
    if (m.n_links > 0 && (!m.link_mats || !m.link_vneg || !m.link_vpos || !m.link_axis || !m.link_mat)) return bad("missing link arrays");
    for (int i = 0; i < m.n_voxels; i++)
        if (m.vox_mat[i] < 0 || m.vox_mat[i] >= m.n_voxel_mats) return bad("voxel material index out of range");
    for (int i = 0; i < m.n_links; i++) {
        if (m.link_vneg[i] < 0 || m.link_vneg[i] >= m.n_voxels || m.link_vpos[i] < 0 || m.link_vpos[i] >= m.n_voxels) return bad("link end out of range");
        if (m.link_mat[i] < 0 || m.link_mat[i] >= m.n_link_mats) return bad("link material index out of range");
        if (m.link_axis[i] < 0 || m.link_axis[i] > 2) return bad("link axis out of range");
    }
    for (int i = 0; i < 6 * m.n_voxels; i++)
        if (m.vox_links[i] >= m.n_links) return bad("voxel link slot out of range");
    // a link of axis a occupies direction slot 2a (+a) of its negative end and 2a+1 (-a) of its positive end
    // (CVX_Voxel::addLinkInfo, src/old/VX_Voxel.cpp; VX3_Link ctor VX3_Link.cu:31-56): the engine stores each end's
    // force by (voxel, direction)
    for (int i = 0; i < m.n_links; i++) {
        const int a = m.link_axis[i];
        if (m.vox_links[6 * m.link_vneg[i] + 2 * a] != i || m.vox_links[6 * m.link_vpos[i] + 2 * a + 1] != i)
            return bad("vox_links does not hold the link in slot 2*axis of its negative end and 2*axis+1 of its positive end");
    }
    for (int s = 0; s < VX3_PROG_COUNT; s++) {
        if (m.prog[s].n < 0 || m.prog[s].n > VX3_MAX_TOKENS) return bad("token program too long");
        if (m.prog[s].n > 0 && !m.prog[s].tok) return bad("token program pointer missing");
    }
    return VX3_OK;
}

static void run_com(vx3_batch *b, int mode) {
    k_com_partial<<<b->D.nchunks, VX3_BLOCK, 0, b->stream>>>(b->D);
    k_sim_update<<<cdiv(b->nsims, 128), 128, 0, b->stream>>>(b->D, mode);
    b->launches += 2;
}

extern "C" int vx3_batch_create(int device, const vx3_model_desc *models, int n, vx3_batch **out) {
    if (!out) return fail(VX3_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!models || n <= 0) return fail(VX3_ERR_INVALID, "no models");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(VX3_ERR_NO_DEVICE, "no CUDA device available (this engine has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(VX3_ERR_NO_DEVICE, "device index out of range");
    // cudaGetDeviceProperties is slow (it queries every attribute): once per device and process
    static std::mutex prop_mu;
    static std::map<int, cudaDeviceProp> prop_cache;
    cudaDeviceProp prop;
    {
        std::lock_guard<std::mutex> lk(prop_mu);
        auto it = prop_cache.find(device);
        if (it == prop_cache.end()) {
            cudaDeviceProp p;
            CK(cudaGetDeviceProperties(&p, device));
            it = prop_cache.emplace(device, p).first;
        }
        prop = it->second;
    }
    const bool timing = getenv("VX3_CREATE_TIMING") != nullptr;
    auto tp0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[create timing] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tp0).count());
        tp0 = now;
    };
    if (prop.major < 10) return fail(VX3_ERR_NO_DEVICE, "device is not sm_100 (kernels are built for sm_100a only)");
    {
        // the checks walk every index of every model: large batches are vetted by a few threads; the first failing model is
        // vetted again on this thread so that its message lands in this thread's error slot
        size_t items = 0;
        for (int i = 0; i < n; i++) items += (size_t)std::max(models[i].n_voxels, 0) + (size_t)std::max(models[i].n_links, 0);
        std::atomic<int> first_bad{n};
        const int nt = (n >= 16 && items > 200000) ? (int)std::min<size_t>({(size_t)std::max(1u, std::thread::hardware_concurrency()), (size_t)8, (size_t)n}) : 1;
        if (nt > 1) {
            std::atomic<int> next{0};
            auto worker = [&]() {
                for (int i; (i = next.fetch_add(1)) < n;)
                    if (validate_model(models[i], i)) {
                        int cur = first_bad.load();
                        while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {}
                    }
            };
            std::vector<std::thread> th;
            for (int k = 0; k < nt; k++) th.emplace_back(worker);
            for (auto &x : th) x.join();
            if (first_bad < n) return validate_model(models[first_bad], first_bad);
        } else
            for (int i = 0; i < n; i++) {
                int rc = validate_model(models[i], i);
                if (rc) return rc;
            }
    }
    lap("validate");
    CK(cudaSetDevice(device));
    vx3_batch *b = new vx3_batch();
    b->device = device;
    b->nsims = n;
    auto cleanup = [&](int rc) {
        vx3_batch_destroy(b);
        return rc;
    };

    // ---- global tables ----
    std::vector<VoxMatC> vmat_tab;
    std::vector<SigMatC> smat_tab; // same index as vmat_tab
    std::vector<LinkMatC> lmat_tab;
    std::vector<float> strain_pool, stress_pool;
    std::map<std::string, int> vmat_index, lmat_index;
    std::vector<vx3_token> tokens;
    std::vector<ExtC> exts;
    std::vector<Chunk> chunks;
    std::vector<int32_t> targets;
    std::vector<SimD> simd(n);
    b->simc.resize(n);
    b->names.resize(n);
    b->lmat_global.resize(n);
    b->matid.resize(n);
    b->vmat_local.resize(n);
    b->matcolor.resize(n);
    b->opts.resize(n);
    b->h_vmats.resize(n);
    b->hdt.assign(n, 0.0f);

    auto add_linkmat = [&](LinkMatC lm, const std::vector<float> &sd, const std::vector<float> &ss) -> int {
        lm.data_off = 0;
        lm.n_data = (int)sd.size();
        std::string key = blob(&lm, sizeof(lm)) + blob(sd.data(), sd.size() * 4) + blob(ss.data(), ss.size() * 4);
        auto it = lmat_index.find(key);
        if (it != lmat_index.end()) return it->second;
        lm.data_off = (int)strain_pool.size();
        strain_pool.insert(strain_pool.end(), sd.begin(), sd.end());
        stress_pool.insert(stress_pool.end(), ss.begin(), ss.end());
        lmat_tab.push_back(lm);
        lmat_index[key] = (int)lmat_tab.size() - 1;
        return (int)lmat_tab.size() - 1;
    };

    size_t nvox = 0, nslots = 0;
    for (int s = 0; s < n; s++) {
        const vx3_model_desc &m = models[s];
        bool sticky = false;
        for (int i = 0; i < m.n_voxel_mats; i++) sticky |= m.voxel_mats[i].sticky != 0;
        const bool collide = m.opt.enable_collision || m.opt.enable_attach;
        int lcap = m.n_links;
        if (collide && sticky) lcap = m.link_capacity > m.n_links ? m.link_capacity : m.n_links + 6 * m.n_voxels + 1024;
        SimC &S = b->simc[s];
        memset(&S, 0, sizeof(S));
        S.voff = (int)nvox;
        S.nvox = m.n_voxels;
        S.loff = (int)nslots;
        S.lcap = lcap;
        S.nhostlinks = m.n_links;
        nvox += m.n_voxels;
        nslots += lcap;
        if (nvox > 0x3FFFFFFF || nslots > 0x7FFFFFF0) return cleanup(fail(VX3_ERR_INVALID, "batch too large for 32-bit indices"));
        b->any_collide |= collide;
        b->any_sticky |= collide && sticky;
        b->any_detach |= m.opt.enable_detach != 0;
        b->any_secondary |= m.opt.secondary_experiment != 0;
        b->any_signals |= m.opt.enable_signals != 0;
    }

    const size_t VS = (nvox + 31) / 32 * 32, LS = (std::max<size_t>(nslots, 1) + 31) / 32 * 32;
    bool any_cilia = false;
    for (int s = 0; s < n; s++) any_cilia |= models[s].opt.enable_cilia != 0;
    // ---- the per-voxel / per-link arrays are built IN PLACE in the pinned staging buffer (no intermediate copies): acquire
    // it now, sized by an upper bound of everything that will be uploaded; the device arena may still grow later ----
    Dev &D = b->D;
    memset(&D, 0, sizeof(D));
    size_t big_bytes = nvox * (8 * 8 + 8 + 24 + 4 * 5 + 4 + 6 * 4 + 6 + 16 + (any_cilia ? 48 : 0)) + 3 * VS * 16 + nslots * (8 + 16 + 4 + 4 + 16 + 8) + 5 * LS * 16;
    size_t small_bound = (size_t)n * (sizeof(SimC) + sizeof(SimD) + 64) + nvox * 8 + 64 * 256 + (size_t)256 * 256 * 40 + ((size_t)1 << 20);
    for (int s = 0; s < n; s++) {
        const vx3_model_desc &m = models[s];
        small_bound += (size_t)(2 * m.n_voxel_mats + m.n_link_mats + 2) * (sizeof(VoxMatC) + sizeof(VoxMatL) + sizeof(SigMatC) + sizeof(LinkMatC) + 512);
        for (int i = 0; i < m.n_voxel_mats; i++) small_bound += 16 * (size_t)(m.voxel_mats[i].n_data + 4);
        for (int i = 0; i < m.n_link_mats; i++) small_bound += 16 * (size_t)(m.link_mats[i].m.n_data + 4);
        for (int p = 0; p < VX3_PROG_COUNT; p++) small_bound += sizeof(vx3_token) * (size_t)(m.prog[p].n + 1);
        small_bound += sizeof(ExtC) * (size_t)(m.n_externals + 1) + sizeof(Chunk) * (size_t)(m.n_voxels / 4096 + 2);
    }
    big_bytes += 40 * 256; // alignment of the slices
    // Fused step (vx3_fused.cuh): fixed link topology only, and OPT-IN (environment VX3_FUSED=1 at creation): it moves 40 % less
    // DRAM traffic than the two-pass kernels but is slower on B200 — the step is bound by fp64 instruction latency at 16
    // warps per SM, not by HBM, and the fused kernel's barriers expose more of it (profiles/r01_ncu_k_fused_*).
    int fuse_bv = VX3_FUSE_BV;
    if (const char *e = getenv("VX3_FUSE_BV")) fuse_bv = std::max(8, std::min(VX3_FUSE_BV, atoi(e))); // test hook: smaller blocks
    const char *fenv = getenv("VX3_FUSED");
    // (signals are excluded because k_signals defines its result by the voxels' index order, which the block layout permutes)
    const bool fuse_eligible = fenv && fenv[0] == '1' && !b->any_collide && !b->any_detach && !b->any_secondary && !b->any_signals && nslots > 0;
    FusedLayout lay;
    const bool permuted = fuse_eligible && fused_layout_build(models, n, b->simc, fuse_bv, lay);
    if (permuted) {
        small_bound += nslots * (4 + 16) + (nvox / 6 + (size_t)n + 16) * 16 + 8 * 256;
        b->vperm.swap(lay.vperm);
        b->lperm.swap(lay.lperm);
        b->vinv.resize(nvox);
        b->linv.resize(nslots);
        for (size_t v = 0; v < nvox; v++) b->vinv[b->vperm[v]] = (int)v;
        for (size_t g = 0; g < nslots; g++) b->linv[b->lperm[g]] = (int)g;
    }
    lap("storage order");
    {
        int rc0 = resources_acquire(device, big_bytes + (6 * LS * 16) + small_bound, big_bytes + small_bound, b->res);
        if (rc0) return cleanup(rc0);
    }
    b->stream = b->res.stream;
    b->ev0 = b->res.ev0;
    b->ev1 = b->res.ev1;
    ArenaPlan plan;
    char *const stg = b->res.h;
    double *pose = plan.take(stg, &D.pose, nvox * 8);
    double2 *mom2 = plan.take(stg, &D.mom2, 3 * VS);
    int32_t *vflags = plan.take(stg, &D.vflags, nvox);
#define TAKE(field, count) plan.take(stg, const_cast<std::remove_const<std::remove_pointer<decltype(D.field)>::type>::type **>(&D.field), count)
    int32_t *vmat = TAKE(vmat, nvox);
    int32_t *vsim = TAKE(vsim, nvox);
    double *phase = TAKE(phase, nvox);
    float *tempe = plan.take(stg, &D.tempe, nvox);
    int32_t *vlinks = plan.take(stg, &D.vlinks, nvox * 6);
    int32_t *vext = TAKE(vext, nvox);
    int16_t *ixyz = TAKE(ixyz, nvox * 3);
    double *initpos = plan.take(stg, &D.initpos, nvox * 3);
    double *base_cilia = nullptr, *shift_cilia = nullptr;
    if (any_cilia) {
        base_cilia = TAKE(base_cilia, nvox * 3);
        shift_cilia = TAKE(shift_cilia, nvox * 3);
    }
    int2 *lends = plan.take(stg, &D.lends, nslots);
    int32_t *lstate = plan.take(stg, &D.lstate, nslots);
    int32_t *lmat = plan.take(stg, &D.lmat, nslots);
    int4 *lc4 = plan.take(stg, &D.lc4, nslots);
    int4 *vc4 = TAKE(vc4, nvox);
    double2 *lh2 = plan.take(stg, &D.lh2, 5 * LS);
    float4 *lstrain = plan.take(stg, &D.lstrain, nslots);
    float2 *larea = plan.take(stg, &D.larea, nslots);
#undef TAKE
    if (plan.placed_bytes > big_bytes) return cleanup(fail(VX3_ERR_INVALID, "internal: staging bound too small"));
    for (size_t v = nvox; v < VS; v++) // padding of the blocked records
        for (int p = 0; p < 3; p++) mom2[idx_mo(p, v)] = make_double2(0.0, 0.0);
    for (size_t g = nslots; g < LS; g++)
        for (int p = 0; p < 5; p++) lh2[idx_lh(p, g)] = make_double2(0.0, 0.0);

    // Pass A (serial, small): materials, programs, externals, targets, CoM chunks — everything that appends to a table
    // shared by the batch.  Pass B (below, one thread per range of simulations): the per-voxel and per-link arrays.
    struct SimBuild {
        std::vector<int> vm_global;
        int ext_base = 0;
        double maxSize = 0, maxCte = 0;
    };
    std::vector<SimBuild> sb(n);
    for (int s = 0; s < n; s++) {
        const vx3_model_desc &m = models[s];
        SimC &S = b->simc[s];
        SimD &dy = simd[s];
        memset(&dy, 0, sizeof(dy));
        b->names[s] = std::string(m.name, strnlen(m.name, sizeof(m.name)));
        b->opts[s] = m.opt;
        // link materials of the model, then the (m,m) pair of every sticky material (attach)
        std::vector<int> &lg = b->lmat_global[s];
        lg.resize(m.n_link_mats);
        for (int i = 0; i < m.n_link_mats; i++) {
            LinkMatC lm;
            std::vector<float> sd, ss;
            device_linkmat(m.link_mats[i], lm, sd, ss);
            lg[i] = add_linkmat(lm, sd, ss);
        }
        std::vector<int> &vm_global = sb[s].vm_global;
        vm_global.resize(m.n_voxel_mats);
        b->matid[s].resize(m.n_voxel_mats);
        double &maxSize = sb[s].maxSize, &maxCte = sb[s].maxCte;
        for (int i = 0; i < m.n_voxel_mats; i++) {
            const vx3_voxel_material &in = m.voxel_mats[i];
            VoxMatC vm;
            memset(&vm, 0, sizeof(vm));
            vm.nomSize = in.nomSize;
            for (int k = 0; k < 3; k++) {
                vm.size[k] = in.nomSize * in.extScale[k];
                maxSize = std::max(maxSize, fabs(vm.size[k]));
            }
            maxCte = std::max(maxCte, (double)fabsf(in.alphaCTE));
            vm.cilia = in.cilia;
            vm.thermal_on_after = in.thermal_on_after_s;
            vm.cilia_on_after = in.cilia_on_after_s;
            vm.remove_after = in.remove_after_s;
            vm.alphaCTE = in.alphaCTE; vm.muStatic = in.muStatic; vm.muKinetic = in.muKinetic;
            vm.massInverse = in.massInverse; vm.mass = in.mass; vm.momentInertiaInverse = in.momentInertiaInverse;
            vm.globalDampT = in.zetaGlobal * in._2xSqMxExS;
            vm.globalDampR = in.zetaGlobal * in._2xSqIxExSxSxS;
            vm.colDampT = in.zetaCollision * in._2xSqMxExS;
            vm.penStiff = (float)(2 * in.E * in.nomSize);
            vm.gravityForce = -in.mass * 9.80665f * in.gravMult;
            vm.dampMultNum = 2 * in.sqrtMass * in.zetaInternal;
            vm.E = in.E;
            vm.nu = in.nu;
            if (in.sticky && in.nu != 0 && m.opt.enable_attach) b->any_sticky_poisson = true;
            vm.fixed = in.fixed != 0; vm.sticky = in.sticky != 0; vm.is_target = in.is_target != 0; vm.is_measured = in.is_measured != 0;
            vm.matid = in.matid;
            vm.self_lmat = -1;
            if (in.sticky && S.lcap > m.n_links) {
                int found = -1;
                for (int k = 0; k < m.n_link_mats && found < 0; k++)
                    if (m.link_mats[k].vox1_mat == i && m.link_mats[k].vox2_mat == i) found = lg[k];
                if (found < 0) { // VX3_MaterialLink(mat, mat) of an attach-created link (VX3_MaterialLink.cu:24-149): the same blend the
                                 // host builder applies to the model's own links (csrc/host/vx3_materials.cpp)
                    vx3_link_material self;
                    std::vector<std::vector<float>> pool;
                    pool.reserve(2);
                    vx3::link_constants(in, in, i, i, &self, &pool);
                    LinkMatC lm;
                    std::vector<float> sd, ss;
                    device_linkmat(self, lm, sd, ss);
                    found = add_linkmat(lm, sd, ss);
                }
                vm.self_lmat = found;
            }
            SigMatC sg;
            memset(&sg, 0, sizeof(sg));
            sg.pacemaker_period = in.pacemaker_period;
            sg.value_decay = in.signal_value_decay;
            sg.time_delay = in.signal_time_delay;
            sg.inactive_period = in.inactive_period;
            sg.is_pacemaker = in.is_pacemaker != 0;
            std::string key = blob(&vm, sizeof(vm));
            if (m.opt.enable_signals) key += blob(&sg, sizeof(sg));
            auto it = vmat_index.find(key);
            if (it == vmat_index.end()) {
                vmat_tab.push_back(vm);
                smat_tab.push_back(sg);
                it = vmat_index.emplace(key, (int)vmat_tab.size() - 1).first;
            }
            vm_global[i] = it->second;
            b->matid[s][i] = in.matid;
            b->h_vmats[s].push_back(in);
            b->h_vmats[s].back().strain_data = b->h_vmats[s].back().stress_data = nullptr;
            b->matcolor[s].push_back(in.r / 255.0f);
            b->matcolor[s].push_back(in.g / 255.0f);
            b->matcolor[s].push_back(in.b / 255.0f);
            b->matcolor[s].push_back(in.a / 255.0f);
        }
        // programs
        for (int p = 0; p < VX3_PROG_COUNT; p++) {
            S.prog_off[p] = (int)tokens.size();
            S.prog_n[p] = m.prog[p].n;
            for (int k = 0; k < m.prog[p].n; k++) tokens.push_back(m.prog[p].tok[k]);
        }
        S.has_ff = S.prog_n[VX3_PROG_FORCE_X] > 0 || S.prog_n[VX3_PROG_FORCE_Y] > 0 || S.prog_n[VX3_PROG_FORCE_Z] > 0;
        S.has_attach_cond = 0;
        for (int c = 0; c < 5; c++) S.has_attach_cond |= S.prog_n[VX3_PROG_ATTACH_0 + c] > 0;
        S.vary_temp = m.opt.vary_temp_enabled != 0;
        S.enable_expansion = m.opt.enable_expansion != 0;
        S.enable_collision = m.opt.enable_collision != 0;
        S.enable_attach = m.opt.enable_attach != 0;
        S.enable_detach = m.opt.enable_detach != 0;
        S.enable_cilia = m.opt.enable_cilia != 0;
        S.safety_guard = m.opt.safety_guard;
        S.secondary_experiment = m.opt.secondary_experiment != 0;
        S.enable_signals = m.opt.enable_signals != 0;
        S.reinit_after = m.opt.reinit_initial_position_after_s;
        S.temp_amp = m.opt.temp_amplitude;
        S.temp_period = m.opt.temp_period;
        S.vox_size = m.opt.vox_size;
        S.pair_radius = m.opt.max_dist_in_voxel_lengths_to_count_as_pair * m.opt.vox_size;
        S.dt_frac = m.opt.dt_frac;
        S.optimal_dt = vx3_model_recommended_dt(&m);
        S.dt_from_state = 0;
        for (int i = 0; i < m.n_links && !S.dt_from_state; i++) S.dt_from_state = m.link_mats[m.link_mat[i]].m.nu != 0.0f;
        b->model_dt.push_back(S.optimal_dt); // recommendedTimeStep() of the model as imported: what CUDA_Simulation prints before its loop
        if (S.dt_from_state) S.optimal_dt = vx3_model_first_step_dt(&m);
        dy.hot_flags = ((S.vary_temp && S.temp_period > 0) ? SHF_THERMAL : 0) | (S.enable_expansion ? SHF_EXPANSION : 0) | (S.enable_cilia ? SHF_CILIA : 0) |
                       (S.has_ff ? SHF_FORCE_FIELD : 0) | (S.has_attach_cond ? SHF_ATTACH_COND : 0) | (S.enable_signals ? SHF_SIGNALS : 0) |
                       (S.enable_detach ? SHF_DETACH : 0) | (S.secondary_experiment ? SHF_SECONDARY : 0) | (S.prog_n[VX3_PROG_STOP] > 0 ? SHF_STOP_PROG : 0);
        dy.temp_amp = S.temp_amp;
        dy.temp_period = S.temp_period;
        dy.topo_epoch = 1;
        b->vmat_local[s].assign(m.vox_mat, m.vox_mat + m.n_voxels);
        sb[s].ext_base = (int)exts.size();
        for (int i = 0; i < m.n_externals; i++) {
            const vx3_external &e = m.externals[i];
            ExtC x;
            memset(&x, 0, sizeof(x));
            x.dof = e.dof_fixed;
            for (int k = 0; k < 3; k++) { x.force[k] = e.force[k]; x.moment[k] = e.moment[k]; x.translation[k] = e.translation[k]; }
            for (int k = 0; k < 4; k++) x.rotq[k] = e.rotation_q[k];
            exts.push_back(x);
        }
        S.tgt_off = (int)targets.size();
        {
            bool any_target = false;
            for (int i = 0; i < m.n_voxel_mats; i++) any_target |= m.voxel_mats[i].is_target != 0;
            if (any_target)
                for (int i = 0; i < m.n_voxels; i++) {
                    if (m.vox_mat[i] < 0 || m.vox_mat[i] >= m.n_voxel_mats) continue; // validate_model has already vetted the indices
                    if (m.voxel_mats[m.vox_mat[i]].is_target) targets.push_back(b->vdev((size_t)S.voff + i)); // registerTargets
                }
        }
        S.ntgt = (int)targets.size() - S.tgt_off;
        // CoM chunks
        S.chunk_off = (int)chunks.size();
        const int CH = 4096;
        for (int c0 = 0; c0 < m.n_voxels; c0 += CH) chunks.push_back(Chunk{s, S.voff + c0, std::min(CH, m.n_voxels - c0), 0});
        S.nchunks = (int)chunks.size() - S.chunk_off;
        dy.status = VX3_SIM_RUNNING;
        dy.link_cnt = m.n_links;
        // collision grid cell: at least the largest collision envelope that can occur, 2 * 0.625 * baseSizeAverage at the
        // temperature that makes a voxel largest.  Temperatures stay in [-|A|, 0], or [-|A|, |A|] with EnableExpansion
        // (gpu_update_temperature, VX3_VoxelyzeKernel.cu:640-645), widened by whatever the model starts with; a material grows
        // with cte * T, so a positive CTE without EnableExpansion never exceeds the nominal size — the cell is then 1.25 voxels
        // instead of 1.5 and the 27-cell sweep sees 42 % less volume
        double t_lo = 0, t_hi = 0;
        if (m.opt.vary_temp_enabled && m.opt.temp_period > 0) {
            t_lo = -fabs(m.opt.temp_amplitude);
            t_hi = m.opt.enable_expansion ? fabs(m.opt.temp_amplitude) : 0.0;
        }
        if (m.temp && (m.opt.enable_collision || m.opt.enable_attach))
            for (int i = 0; i < m.n_voxels; i++) {
                t_lo = std::min(t_lo, (double)m.temp[i]);
                t_hi = std::max(t_hi, (double)m.temp[i]);
            }
        double grow = 1.0;
        for (int i = 0; i < m.n_voxel_mats; i++) {
            const double cte = m.voxel_mats[i].alphaCTE;
            grow = std::max(grow, std::max(1 + t_lo * cte, 1 + t_hi * cte));
        }
        const double cell = 2 * VX3_COLLISION_ENVELOPE_RADIUS * sb[s].maxSize * grow * (1 + 1e-5);
        S.cell_inv = cell > 0 ? 1.0 / cell : 1.0;
    }

    std::atomic<int> fill_err{0}, ghost_seen{0};
    // Pass B tasks: voxels [i0, i1) resp. link slots [i0, i1) of simulation s.  Every element of every array is written.
    auto fill_voxels = [&](int s, int i0, int i1) {
        const vx3_model_desc &m = models[s];
        const SimC &S = b->simc[s];
        const std::vector<int> &vm_global = sb[s].vm_global;
        const int vo = S.voff, ext_base = sb[s].ext_base;
        bool ghost = false;
        for (int i = i0; i < i1; i++) {
            const size_t g = (size_t)b->vdev((size_t)vo + i);
            pose[8 * g + 0] = m.pos[3 * i]; pose[8 * g + 1] = m.pos[3 * i + 1]; pose[8 * g + 2] = m.pos[3 * i + 2];
            if (m.orient) for (int k = 0; k < 4; k++) pose[8 * g + 3 + k] = m.orient[4 * i + k];
            else { pose[8 * g + 3] = 1.0; pose[8 * g + 4] = pose[8 * g + 5] = pose[8 * g + 6] = 0.0; }
            pose[8 * g + 7] = 0.0; // {temperature, previousDt}: k_temp_init
            for (int k = 0; k < 3; k++) initpos[3 * g + k] = m.pos[3 * i + k];
            const double l0 = m.lin_mom ? m.lin_mom[3 * i] : 0.0, l1 = m.lin_mom ? m.lin_mom[3 * i + 1] : 0.0, l2 = m.lin_mom ? m.lin_mom[3 * i + 2] : 0.0;
            const double a0 = m.ang_mom ? m.ang_mom[3 * i] : 0.0, a1 = m.ang_mom ? m.ang_mom[3 * i + 1] : 0.0, a2 = m.ang_mom ? m.ang_mom[3 * i + 2] : 0.0;
            mom2[idx_mo(0, g)] = make_double2(l0, l1);
            mom2[idx_mo(1, g)] = make_double2(l2, a0);
            mom2[idx_mo(2, g)] = make_double2(a1, a2);
            vflags[g] = (m.vox_flags[i] & VXF_BOOLSTATE_MASK) | VXF_ENABLE_ATTACH;
            ghost |= (m.vox_flags[i] & VX3_VOX_GHOST) != 0;
            vmat[g] = vm_global[m.vox_mat[i]];
            vsim[g] = s;
            phase[g] = m.phase_offset ? m.phase_offset[i] : 0.0;
            tempe[g] = m.temp ? m.temp[i] : 0.0f;
            for (int k = 0; k < 6; k++) {
                int li = m.vox_links[6 * i + k];
                vlinks[6 * g + k] = li >= 0 ? b->ldev((size_t)S.loff + li) : -1;
            }
            vext[g] = -1;
            if (m.vox_ext && m.vox_ext[i] >= 0) {
                if (m.vox_ext[i] >= m.n_externals) {
                    fill_err = 1;
                    return;
                }
                vext[g] = ext_base + m.vox_ext[i];
            }
            vc4[g] = make_int4(vmat[g], s, vext[g], 0);
            ixyz[3 * g] = m.ix[i]; ixyz[3 * g + 1] = m.iy[i]; ixyz[3 * g + 2] = m.iz[i];
            if (any_cilia) for (int k = 0; k < 3; k++) {
                base_cilia[3 * g + k] = m.base_cilia ? m.base_cilia[3 * i + k] : 0.0;
                shift_cilia[3 * g + k] = m.shift_cilia ? m.shift_cilia[3 * i + k] : 0.0;
            }
        }
        if (ghost) ghost_seen = 1;
    };
    auto fill_links = [&](int s, int i0, int i1) { // slots up to the simulation's capacity: the spare ones are marked empty
        const vx3_model_desc &m = models[s];
        const SimC &S = b->simc[s];
        const std::vector<int> &lg = b->lmat_global[s];
        const int vo = S.voff;
        for (int i = i0; i < i1; i++) {
            const size_t g = (size_t)b->ldev((size_t)S.loff + i);
            if (i >= m.n_links) { // spare pool slot (attach)
                lends[g] = make_int2(-1, -1);
                lc4[g] = make_int4(-1, -1, 0, 0);
                lstate[g] = 0;
                lmat[g] = 0;
                for (int p = 0; p < 5; p++) lh2[idx_lh(p, g)] = make_double2(0.0, 0.0);
                lstrain[g] = make_float4(0, 0, 0, 0);
                larea[g] = make_float2(0, 0);
                continue;
            }
            const int vn = m.link_vneg[i], vp = m.link_vpos[i], ax = m.link_axis[i];
            lends[g] = make_int2(b->vdev((size_t)vo + vn), b->vdev((size_t)vo + vp));
            lmat[g] = lg[m.link_mat[i]];
            lc4[g] = make_int4(lends[g].x, lends[g].y, lmat[g], s);
            int st = (ax << LKS_AXIS_SHIFT);
            if (!m.link_small_angle || m.link_small_angle[i]) st |= LKS_SMALL;
            if (m.link_flags && (m.link_flags[i] & VX3_LINK_LOCAL_VELOCITY_VALID)) st |= LKS_VALID;
            lstate[g] = st;
            double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int k = 0; k < 3; k++) {
                if (m.link_pos2) h[k] = m.link_pos2[3 * i + k];
                if (m.link_angle1v) h[3 + k] = m.link_angle1v[3 * i + k];
                if (m.link_angle2v) h[6 + k] = m.link_angle2v[3 * i + k];
            }
            const vx3_voxel_material &mn = m.voxel_mats[m.vox_mat[vn]], &mp = m.voxel_mats[m.vox_mat[vp]];
            const float tn = m.temp ? m.temp[vn] : 0.0f, tp = m.temp ? m.temp[vp] : 0.0f;
            // VX3_Link::reset() defaults (VX3_Link.cu:58-70) unless the model carries link state
            const double rest = m.link_rest_length ? m.link_rest_length[i]
                                                   : 0.5 * ((mn.nomSize * mn.extScale[ax]) * (1 + tn * mn.alphaCTE) + (mp.nomSize * mp.extScale[ax]) * (1 + tp * mp.alphaCTE));
            lh2[idx_lh(0, g)] = make_double2(h[0], h[1]);
            lh2[idx_lh(1, g)] = make_double2(h[2], h[3]);
            lh2[idx_lh(2, g)] = make_double2(h[4], h[5]);
            lh2[idx_lh(3, g)] = make_double2(h[6], h[7]);
            lh2[idx_lh(4, g)] = make_double2(h[8], rest);
            float4 sn = make_float4(0, 0, 0, 0);
            if (m.link_strain) sn.x = m.link_strain[i];
            if (m.link_max_strain) sn.y = m.link_max_strain[i];
            if (m.link_strain_offset) sn.z = m.link_strain_offset[i];
            if (m.link_stress) sn.w = m.link_stress[i];
            lstrain[g] = sn;
            const float a0 = (float)mn.nomSize, a1 = (float)mp.nomSize;
            larea[g] = make_float2(m.link_transverse_area ? m.link_transverse_area[i] : 0.5f * (a0 * a0 + a1 * a1),
                                   m.link_transverse_strain_sum ? m.link_transverse_strain_sum[i] : 0.0f);
        }
    };
    {
        // tasks of <= 64k items, pulled by a few threads (disjoint output ranges); small batches stay on this thread
        struct Task { int s, kind, i0, i1; };
        std::vector<Task> tasks;
        const int CH = 65536;
        for (int s = 0; s < n; s++) {
            for (int i = 0; i < models[s].n_voxels; i += CH) tasks.push_back(Task{s, 0, i, std::min(i + CH, models[s].n_voxels)});
            for (int i = 0; i < b->simc[s].lcap; i += CH) tasks.push_back(Task{s, 1, i, std::min(i + CH, b->simc[s].lcap)});
        }
        int nthreads = 1;
        if (nvox + nslots > 200000) nthreads = (int)std::min<size_t>({(size_t)std::max(1u, std::thread::hardware_concurrency()), (size_t)16, tasks.size()});
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            for (size_t k; (k = next.fetch_add(1)) < tasks.size();) {
                const Task &t = tasks[k];
                if (t.kind == 0) fill_voxels(t.s, t.i0, t.i1);
                else fill_links(t.s, t.i0, t.i1);
            }
        };
        if (nthreads <= 1) worker();
        else {
            std::vector<std::thread> th;
            for (int k = 0; k < nthreads; k++) th.emplace_back(worker);
            for (auto &x : th) x.join();
        }
        if (fill_err) return cleanup(fail(VX3_ERR_INVALID, "external index out of range"));
        b->any_ghost |= ghost_seen != 0;
        b->vox_active = nvox;
        if (b->any_ghost)
            while (b->vox_active > 0 && (vflags[b->vox_active - 1] & VX3_VOX_GHOST)) b->vox_active--;
    }

    lap("host model -> SoA arrays (in staging)");
    D.nsims = n;
    D.nvox = (int)nvox;
    D.nlinkslots = (int)nslots;
    D.nchunks = (int)chunks.size();
    D.vstride = (int)VS;
    D.lstride = (int)LS;
    // ---- plan the rest of the arena: the small tables follow the in-place arrays, zero-initialised slices come last ----
    std::vector<int32_t> uf_host, vnb_host;
    std::vector<float4> pcache_host;
#define UP(field, vec) plan.upload(const_cast<std::remove_const<std::remove_pointer<decltype(D.field)>::type>::type **>(&D.field), vec)
    UP(simc, b->simc);
    UP(simd, simd);
    UP(vmat_tab, vmat_tab);
    UP(lmat_tab, lmat_tab);
    std::vector<VoxMatL> vl(vmat_tab.size());
    for (size_t i = 0; i < vmat_tab.size(); i++) {
        memset(&vl[i], 0, sizeof(VoxMatL));
        for (int k = 0; k < 3; k++) vl[i].size[k] = vmat_tab[i].size[k];
        vl[i].thermal_on_after = vmat_tab[i].thermal_on_after;
        vl[i].alphaCTE = vmat_tab[i].alphaCTE;
        vl[i].dampMultNum = vmat_tab[i].dampMultNum;
        vl[i].fixed = vmat_tab[i].fixed;
    }
    UP(vmatl_tab, vl);
    D.n_vmats = (int)vmat_tab.size();
    D.n_lmats = (int)lmat_tab.size();
    b->h_lmat_tab = lmat_tab;
    UP(strain_pool, strain_pool);
    UP(stress_pool, stress_pool);
    UP(tokens, tokens);
    UP(exts, exts);
    UP(chunks, chunks);
    UP(targets, targets);
    if (b->any_signals) {
        UP(smat_tab, smat_tab);
        plan.zeroed(&D.sig, 6 * nvox);
        plan.zeroed(&D.sprop, nvox);
    }
#undef UP
    plan.zeroed(&D.lf2, 6 * LS);
    plan.zeroed(&D.com_part, chunks.size() * 6);
    plan.zeroed(&D.vox_count, 1);
    {
        const char *pe = getenv("VX3_PDL");
        b->pdl = pe && pe[0] == '1';
        const char *tf = getenv("VX3_TAIL_FUSED");
        // measured: config 4 (one simulation) 61.8 -> 59.4 us per step; config 3 (512 simulations: four rounds of dependent loads in one
        // CTA at the end of the pass) 91.4 -> 94.9 — so only where one round covers every simulation
        b->tail_in_voxels = !(tf && tf[0] == '0') && !b->any_signals && !b->any_secondary && n <= VX3_VOX_T;
    }
    if (b->any_collide) {
        int H = 1024; // buckets >= 2 x voxels: two occupied cells in one bucket are rare, a walk meets few foreign voxels
        while ((size_t)H < 2 * nvox && H < (1 << 24)) H <<= 1;
        if (const char *e = getenv("VX3_GRID_BUCKETS")) { // test hook: a tiny table forces shared buckets and overflow chains
            int h = atoi(e);
            if (h >= 1 && (h & (h - 1)) == 0) H = h;
        }
        D.hmask = H - 1;
        plan.zeroed(&D.contact, nvox * 3);
        plan.zeroed(&D.cell_cnt, 2 * (size_t)H); // counts, then overflow heads
        plan.zeroed(&D.cell_items, (size_t)H * VX3_CELL_SLOTS); // CellItem[H][8]: 128 bytes per bucket
        plan.zeroed(&D.cell_next, nvox);
        plan.zeroed(&D.crec, nvox);
        if (b->any_sticky) { // connected components of the models' link graphs (roots = smallest voxel index), for uf_find
            uf_host.resize(nvox);
            for (size_t v = 0; v < nvox; v++) uf_host[v] = (int32_t)v;
            auto find = [&](int x) {
                while (uf_host[x] != x) {
                    uf_host[x] = uf_host[uf_host[x]];
                    x = uf_host[x];
                }
                return x;
            };
            for (size_t g = 0; g < nslots; g++) {
                if (lends[g].x < 0) continue;
                const int ra = find(lends[g].x), rb = find(lends[g].y);
                if (ra != rb) uf_host[std::max(ra, rb)] = std::min(ra, rb);
            }
            for (size_t v = 0; v < nvox; v++) uf_host[v] = find((int)v);
            plan.upload(&D.uf, uf_host);
            vnb_host.assign(nvox * 8, -1);
            for (size_t v = 0; v < nvox; v++)
                for (int k = 0; k < 6; k++) {
                    const int li = vlinks[6 * v + k];
                    if (li >= 0) vnb_host[8 * v + k] = lends[li].x == (int)v ? lends[li].y : lends[li].x;
                }
            plan.upload(&D.vnb, vnb_host);
            plan.zeroed(&D.nbcache, nvox * 8);
            if (b->any_sticky_poisson) { // cached poissons strain per voxel: valid with zeros since import where the host computed it
                                         // (CVX_Link::reset of every model link calls it on both ends when nu != 0)
                pcache_host.assign(nvox, make_float4(0.f, 0.f, 0.f, __int_as_float_host(-2)));
                for (size_t v = 0; v < nvox; v++) {
                    bool linked = false;
                    for (int k = 0; k < 6; k++) linked |= vlinks[6 * v + k] >= 0;
                    if (linked && vmat_tab[vmat[v]].nu != 0) pcache_host[v].w = __int_as_float_host(-1);
                }
                plan.upload(&D.pcache, pcache_host);
            }
        }
    }
    {
        // per-simulation regions: attach candidates (a power-of-two capacity of at least 4 per voxel, so that the candidate sort
        // can pad in place) and the failed-link list of EnableDetach (every slot could fail at once)
        size_t ncand = 0, nfail = 0;
        for (int s = 0; s < n; s++) {
            SimC &S = b->simc[s];
            const vx3_model_desc &m = models[s];
            bool sticky = false;
            for (int i = 0; i < m.n_voxel_mats; i++) sticky |= m.voxel_mats[i].sticky != 0;
            S.cand_off = S.cand_cap = S.fail_off = S.fail_cap = 0;
            if ((m.opt.enable_collision || m.opt.enable_attach) && sticky && S.lcap > S.nhostlinks) {
                size_t cap = 256;
                while (cap < 4 * (size_t)S.nvox + 256) cap <<= 1;
                S.cand_off = (int)ncand;
                S.cand_cap = (int)cap;
                ncand += cap;
            }
            if (m.opt.enable_detach && S.lcap > 0) {
                S.fail_off = (int)nfail;
                S.fail_cap = S.lcap;
                nfail += (size_t)S.lcap;
            }
        }
        if (ncand > 0x7FFFFFF0 || nfail > 0x7FFFFFF0) return cleanup(fail(VX3_ERR_INVALID, "batch too large for 32-bit indices"));
        if (ncand) plan.zeroed(&D.cands, ncand);
        if (nfail) plan.zeroed(&D.fail_list, nfail);
    }
    // on-chip path for a single small collision-free body: its control words, flags, lane tables and the odd-parity pose
    // buffer are arena slices too
    PersistentTables ptab;
    persistent_plan(b->pplan, b->simc, b->any_collide, b->any_detach || b->any_secondary || b->any_ghost || b->any_signals, any_cilia, prop, lends, vlinks, ixyz, ptab);
    if (b->pplan.ok) {
        plan.zeroed(&b->pplan.ctl, 4 + 2 * 8 * (size_t)b->pplan.grid);
        plan.zeroed(&b->pplan.flags, (size_t)b->pplan.grid * 32);
        plan.zeroed(&b->pplan.pose_alt, 8 * nvox);
        plan.upload(&b->pplan.lk_slot, ptab.lk_slot);
        plan.upload(&b->pplan.vx_id, ptab.vx_id);
        plan.upload(&b->pplan.vx_lane, ptab.vx_lane);
        plan.upload(&b->pplan.deps, ptab.deps);
        plan.upload(&b->pplan.ndeps, ptab.ndeps);
    }
    std::vector<int> face_slot;
    std::vector<int4> face_c4;
    if (permuted) { // the face links of each simulation are one slot range after its blocks' interior links
        FusedPlan &f = b->fplan;
        f.ok = true;
        f.nblocks = (int)lay.blk.size();
        f.ninterior = (int)lay.ninterior;
        f.nface = (int)lay.nface;
        face_slot.reserve(f.nface);
        face_c4.reserve(f.nface);
        for (int s = 0; s < n; s++)
            for (int k = 0; k < lay.face_range[s].y; k++) {
                face_slot.push_back(lay.face_range[s].x + k);
                face_c4.push_back(lc4[(size_t)lay.face_range[s].x + k]);
            }
        plan.upload(&f.blk, lay.blk);
        plan.upload(&f.face_slot, face_slot);
        plan.upload(&f.face_c4, face_c4);
    }
    lap("persistent + fused plans");
    plan.layout();
    lap("arena layout");
    int rc;
    if (plan.up_bytes > b->res.hcap) return cleanup(fail(VX3_ERR_INVALID, "internal: staging bound too small"));
    // a batch on the persistent path keeps a launch-start copy of its arena (persist_guard) in the upper half of the same
    // allocation: it comes out of the per-process cache like the arena itself (no cudaMalloc / cudaFree per batch)
    const size_t arena_al = (plan.total + 255) / 256 * 256;
    if ((rc = resources_grow_device(b->res, b->pplan.ok ? 2 * arena_al : plan.total))) return cleanup(rc);
    b->arena_bytes = plan.total;
    b->snap = b->pplan.ok ? reinterpret_cast<unsigned char *>(b->res.d) + arena_al : nullptr;
    lap("grow arena");
    for (const ArenaPlan::Item &it : plan.placed) *it.field = b->res.d + it.off;
    for (const ArenaPlan::Item &it : plan.up) {
        *it.field = b->res.d + it.off;
        memcpy(b->res.h + it.off, it.src, it.bytes); // the small tables
    }
    for (const ArenaPlan::Item &it : plan.zero) *it.field = b->res.d + it.off;
    if (D.cell_cnt) D.cell_ovf = D.cell_cnt + ((size_t)D.hmask + 1);
    lap("stage into pinned memory");
    cudaError_t ce = cudaSuccess;
    if (plan.up_bytes) ce = cudaMemcpyAsync(b->res.d, b->res.h, plan.up_bytes, cudaMemcpyHostToDevice, b->stream);
    if (ce == cudaSuccess && plan.total > plan.up_bytes) ce = cudaMemsetAsync(b->res.d + plan.up_bytes, 0, plan.total - plan.up_bytes, b->stream);
    if (ce != cudaSuccess) return cleanup(fail(VX3_ERR_CUDA, std::string("batch upload: ") + cudaGetErrorString(ce)));
    if ((rc = setup_stream_kernels(b, prop))) return cleanup(rc);
    // device-side init at the top of CUDA_Simulation (VX3_SimulationManager.cu:20-24,54-55)
    run_com(b, 0);
    k_temp_init<<<cdiv(D.nvox, VX3_BLOCK), VX3_BLOCK, 0, b->stream>>>(D);
    k_set_dt<<<cdiv(n, 128), 128, 0, b->stream>>>(D, -1.0f);
    for (int s = 0; s < n; s++) {
        double od = b->simc[s].optimal_dt;
        if (od < 1e-10) od = 1e-10;
        b->hdt[s] = (float)(b->simc[s].dt_frac * od);
    }
    ce = cudaStreamSynchronize(b->stream);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    if (ce != cudaSuccess) return cleanup(fail(VX3_ERR_CUDA, std::string("batch init: ") + cudaGetErrorString(ce)));
    lap("upload + device init");
    *out = b;
    return VX3_OK;
}

extern "C" int vx3_batch_check_neighbor_search(vx3_batch *b, int sim, int n_pairs, unsigned seed, int *mismatches, int *positives) {
    if (!b || sim < 0 || sim >= b->nsims || n_pairs <= 0 || !mismatches || !positives) return fail(VX3_ERR_INVALID, "bad arguments");
    if (!b->D.vnb) return fail(VX3_ERR_INVALID, "the batch holds no adjacency table (no simulation in it attaches)");
    CK(cudaSetDevice(b->device));
    int *out = nullptr;
    int rc = b->alloc(&out, 2);
    if (rc) return rc;
    k_check_neighbor_search<<<cdiv(n_pairs, 128), 128, 0, b->stream>>>(b->D, b->simc[sim].voff, b->simc[sim].nvox, n_pairs, seed, out);
    int h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, out, sizeof(h), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    CK(cudaGetLastError());
    *mismatches = h[0];
    *positives = h[1];
    return VX3_OK;
}

static void graph_invalidate(vx3_batch *b);
extern "C" void vx3_batch_destroy(vx3_batch *b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (int sd = 0; sd < 2; sd++)
        if (b->halo.side[sd].peer_open) cudaIpcCloseMemHandle(b->halo.side[sd].peer_flag);
    graph_invalidate(b);
    for (auto &e : b->prof.ev) cudaEventDestroy(e);
    for (auto &e : b->lq_ev)
        if (e) cudaEventDestroy(e);
    for (void *p : b->allocs) cudaFree(p);
    resources_release(b->res); // arena, staging, stream and events go back to the cache
    delete b;
}

// ------------------------------------------------------------------ stepping
static int set_dt(vx3_batch *b, float dt) {
    k_set_dt<<<cdiv(b->nsims, 128), 128, 0, b->stream>>>(b->D, dt);
    b->launches++;
    for (int s = 0; s < b->nsims; s++) {
        if (dt < 0) {
            double od = b->simc[s].optimal_dt;
            if (od < 1e-10) od = 1e-10;
            b->hdt[s] = (float)(b->simc[s].dt_frac * od);
        } else
            b->hdt[s] = dt;
    }
    return VX3_OK;
}

// does any simulation sample its centre of mass at doTimeStep call number `step` (1-based CurStepCount)?
static bool com_step(const vx3_batch *b, long long step) {
    for (int s = 0; s < b->nsims; s++) {
        if (b->hdt[s] == 0) continue;
        const int cycle = (int)(b->simc[s].temp_period / b->hdt[s]);
        if (cycle > 0 && step % cycle == 0) return true;
    }
    return false;
}
// number of steps from hsteps until (and including) the next CoM sampling step; 0 = never
static long long next_com_step(const vx3_batch *b) {
    long long best = 0;
    for (int s = 0; s < b->nsims; s++) {
        if (b->hdt[s] == 0) continue;
        const int cycle = (int)(b->simc[s].temp_period / b->hdt[s]);
        if (cycle <= 0) continue;
        const long long nxt = (b->hsteps / cycle + 1) * cycle - b->hsteps;
        if (best == 0 || nxt < best) best = nxt;
    }
    return best;
}

// Programmatic dependent launch for the kernels that begin with VX3_PDL_ENTRY (vx3_kernels.cuh): the next kernel's CTAs are scheduled
// while the previous kernel drains and wait at their first instruction until it has completed and its writes are visible —
// what overlaps is the launch gap between two dependent kernels, nothing else.  Opt-in (VX3_PDL=1): measured no faster inside the
// replayed graphs (config 3 92.5 -> 93.5, config 4 59.6 -> 60.3 us per step; the early CTAs hold slots the draining kernel's tail could use).
template <class... KArgs, class... Args>
static inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    if (pdl) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
    }
    cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}
#define LAUNCH_P(cls, kern, grid, block, smem, ...)                                                                     \
    do {                                                                                                                \
        const bool p_ = b->prof.begin(cls, st);                                                                         \
        launch_pdl(kern, dim3(grid), dim3(block), smem, st, b->pdl && !p_, __VA_ARGS__);                                \
        if (p_) b->prof.end(st);                                                                                        \
        b->launches++;                                                                                                  \
    } while (0)

#define LAUNCH(cls, kern, grid, block, ...)                                                                            \
    do {                                                                                                                \
        const bool p_ = b->prof.begin(cls, st);                                                                         \
        kern<<<grid, block, 0, st>>>(__VA_ARGS__);                                                                      \
        if (p_) b->prof.end(st);                                                                                        \
        b->launches++;                                                                                                  \
    } while (0)

#define LAUNCH_SM(cls, kern, grid, block, smem, ...)                                                                   \
    do {                                                                                                                \
        const bool p_ = b->prof.begin(cls, st);                                                                         \
        kern<<<grid, block, smem, st>>>(__VA_ARGS__);                                                                   \
        if (p_) b->prof.end(st);                                                                                        \
        b->launches++;                                                                                                  \
    } while (0)

// persistent tile loops: one wave of CTAs (SMs x resident CTAs per SM), each striding over the tiles
static int setup_stream_kernels(vx3_batch *b, const cudaDeviceProp &prop) {
    b->link_smtab = b->D.n_vmats <= VX3_SM_VMATS && b->D.n_lmats <= VX3_SM_LMATS;
    b->vox_smtab = b->D.n_vmats <= VX3_SM_VMATS;
    const void *kl = b->link_smtab ? (const void *)k_links<true> : (const void *)k_links<false>;
    const void *kv = b->vox_smtab ? (const void *)k_voxels<true> : (const void *)k_voxels<false>;
    int nl = 0, nv = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nl, kl, VX3_LINK_T, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nv, kv, VX3_VOX_T, 0));
    if (nl < 1 || nv < 1) return fail(VX3_ERR_CUDA, "streaming kernels do not fit on this device");
    b->link_tiles = cdiv(b->D.nlinkslots, VX3_LINK_T);
    b->vox_tiles = std::max(1, cdiv((int)b->vox_active, VX3_VOX_T));
    b->link_grid = std::max(1, std::min(b->link_tiles, nl * prop.multiProcessorCount));
    b->vox_grid = std::max(1, std::min(b->vox_tiles, nv * prop.multiProcessorCount));
    FusedPlan &f = b->fplan;
    if (f.ok) { // fused step: dynamic shared memory above the 48 KB default, one wave of CTAs striding over the blocks
        f.smtab = b->link_smtab && b->vox_smtab;
        f.smem = sizeof(FusedSmem);
        const void *ks[4] = {(const void *)k_fused<true, false>, (const void *)k_fused<true, true>, (const void *)k_fused<false, false>, (const void *)k_fused<false, true>};
        int nf = 0;
        for (const void *k : ks) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f.smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nf, ks[f.smtab ? 0 : 2], VX3_FUSE_T, f.smem));
        if (nf < 1) f.ok = false;
        else {
            f.grid = std::max(1, std::min(f.nblocks, nf * prop.multiProcessorCount));
            f.face_tiles = cdiv(f.nface, VX3_LINK_T);
            f.face_grid = std::max(1, std::min(f.face_tiles, nl * prop.multiProcessorCount));
        }
    }
    return VX3_OK;
}

// The link pass has two bit-identical variants (k_links: the large-angle branch in place; k_links_deferred: those links
// deferred to dense per-warp passes); which is faster depends on the batch, so the first streaming steps of a batch time
// both with CUDA events (one warm-up round, then 3 trials each) and the batch keeps the faster.  VX3_LINK_QUEUE=0/1 pins it.
static void launch_links(vx3_batch *b, int tile0 = 0, int tile_end = -1, bool collect = false) {
    const Dev &D = b->D;
    cudaStream_t st = b->stream;
    if (tile_end < 0) tile_end = b->link_tiles;
    if (tile_end <= tile0) return;
    const int grid = std::max(1, std::min(tile_end - tile0, b->link_grid));
    int variant = b->link_queue;
    bool trial = false;
    if (variant < 0) {
        if (!b->lq_ev[0]) {
            const char *e = getenv("VX3_LINK_QUEUE");
            if (e && (e[0] == '0' || e[0] == '1')) b->link_queue = variant = e[0] - '0';
            else if (b->halo.on) b->link_queue = variant = 1; // slabs of one large body (measured faster there); and a host-side event wait
                                                              // inside a step could deadlock slabs that one thread queues in rounds
            else {
                cudaEventCreate(&b->lq_ev[0]);
                cudaEventCreate(&b->lq_ev[1]);
            }
        }
        if (variant < 0) {
            trial = true;
            variant = b->lq_trials & 1;
            cudaEventRecord(b->lq_ev[0], st);
        }
    }
    if (collect) { // slab batch: the face links read their ghost ends from the receive buffers (vx3_kernels.cuh, halo_arrival_wait)
        if (b->link_smtab) LAUNCH_P(KC_LINKS, (k_links_deferred<true, true>), grid, VX3_LINK_T, 0, D, tile_end, tile0);
        else LAUNCH_P(KC_LINKS, (k_links_deferred<false, true>), grid, VX3_LINK_T, 0, D, tile_end, tile0);
        return;
    }
    if (b->link_smtab) {
        if (variant) LAUNCH_P(KC_LINKS, (k_links_deferred<true, false>), grid, VX3_LINK_T, 0, D, tile_end, tile0);
        else LAUNCH_P(KC_LINKS, (k_links<true, false>), grid, VX3_LINK_T, 0, D, tile_end, nullptr, nullptr, 0, tile0);
    } else {
        if (variant) LAUNCH_P(KC_LINKS, (k_links_deferred<false, false>), grid, VX3_LINK_T, 0, D, tile_end, tile0);
        else LAUNCH_P(KC_LINKS, (k_links<false, false>), grid, VX3_LINK_T, 0, D, tile_end, nullptr, nullptr, 0, tile0);
    }
    if (trial) {
        cudaEventRecord(b->lq_ev[1], st);
        float ms = 0;
        if (cudaEventSynchronize(b->lq_ev[1]) == cudaSuccess && cudaEventElapsedTime(&ms, b->lq_ev[0], b->lq_ev[1]) == cudaSuccess && b->lq_trials >= 2)
            b->lq_ms[variant] += ms;
        if (++b->lq_trials >= 8) b->link_queue = b->lq_ms[1] < b->lq_ms[0] ? 1 : 0;
    }
}

template <class T> static int d2h(vx3_batch *b, std::vector<T> &h, const T *d, size_t off, size_t n);
// ghost poses of the previous step: wait for the neighbours' step numbers, scatter their records (k_halo_recv) — on the main
// stream, placed right before the first kernel that reads a ghost pose
static void halo_wait(vx3_batch *b) {
    Halo &H = b->halo;
    if (!H.on) return;
    if (!H.pending && !b->capturing) return; // (inside a captured stretch every step collects: the kernels themselves know whether there is anything new)
    H.pending = false;
    cudaStream_t st = b->stream;
    const Dev &D = b->D;
    HaloRecvArgs a;
    memset(&a, 0, sizeof(a));
    for (int sd = 0; sd < 2; sd++) {
        HaloSide &h = H.side[sd];
        if (h.n_recv > 0 && (h.peer_open || h.peer_local)) {
            a.idx[sd] = h.recv_idx; a.recv_buf[sd] = h.recv_buf; a.recv_flag[sd] = h.recv_flag;
            a.n[sd] = h.n_recv;
            a.nb[sd] = std::min(1024, cdiv(4 * h.n_recv, VX3_HALO_BLOCK));
        }
    }
    if (a.nb[0] + a.nb[1] > 0) {
        LAUNCH(KC_HALO, k_halo_wait, 1, 32, a, H.seq, H.err, H.spin_cycles, D.simd);
        LAUNCH(KC_HALO, k_halo_recv, a.nb[0] + a.nb[1], VX3_HALO_BLOCK, D.pose, a, H.seq, H.err);
    }
}
// CUDA loads a kernel's code at its first launch (lazy module loading), and that load can wait for the kernels already running
// on the device.  A slab's k_halo_wait spins until its neighbour has sent — if the neighbour is driven by this process and its
// first launch of some kernel has to load it, the two wait for each other until the spin limit (seen as a failed
// vx3_batch_sync in tests/test_decomposition.py with two slabs on one device).  So every kernel of the step path is loaded before
// the first exchange.
static void preload_step_kernels() {
    static std::once_flag once;
    std::call_once(once, [] {
        const void *ks[] = {(const void *)k_links<true>, (const void *)k_links<false>, (const void *)k_links_deferred<true>, (const void *)k_links_deferred<false>, (const void *)k_links_deferred<true, true>, (const void *)k_links_deferred<false, true>,
                            (const void *)k_links<true, true>, (const void *)k_links<false, true>, (const void *)k_voxels<true>, (const void *)k_voxels<false>, (const void *)k_voxels<true, true>, (const void *)k_voxels<false, true>,
                            (const void *)k_fused<true, false>, (const void *)k_fused<true, true>, (const void *)k_fused<false, false>, (const void *)k_fused<false, true>,
                            (const void *)k_tail_light, (const void *)k_tail, (const void *)k_com_partial, (const void *)k_sim_update, (const void *)k_set_dt,
                            (const void *)k_halo_send, (const void *)k_halo_wait, (const void *)k_halo_recv, (const void *)k_surface, (const void *)k_secondary,
                            (const void *)k_signals, (const void *)k_step_cap};
        for (const void *k : ks) {
            cudaFuncAttributes fa;
            if (cudaFuncGetAttributes(&fa, k) != cudaSuccess) cudaGetLastError();
        }
    });
}

// first step of a connected slab batch: the spin limit in cycles, and how many leading link tiles are free of ghost ends (the
// host-side partition stores the face links last)
static int halo_prepare(vx3_batch *b) {
    Halo &H = b->halo;
    H.send_blocks = 0;
    for (int sd = 0; sd < 2; sd++) {
        const HaloSide &h = H.side[sd];
        if (h.n_send > 0 && (h.peer_open || h.peer_local)) H.send_blocks += std::min(1024, cdiv(4 * h.n_send, VX3_HALO_BLOCK));
    }
    if (H.face_tile0 >= 0) return VX3_OK;
    double ms = VX3_HALO_TIMEOUT_MS_DEFAULT;
    if (const char *e = getenv("VX3_HALO_TIMEOUT_MS")) ms = std::max(1.0, atof(e));
    int khz = 1500000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, b->device);
    H.spin_cycles = (long long)(ms * (double)khz);
    H.face_tile0 = 0;
    const char *ik = getenv("VX3_HALO_INKERNEL");
    // (VX3_HALO_INKERNEL=2, a test hook: also between two small slabs of one process, where the stand-alone kernels are the default
    // because a waiting link pass of one slab could hold the CTA slots the other slab needs)
    const bool local = H.side[0].peer_local || H.side[1].peer_local;
    H.inkernel = !(ik && ik[0] == '0') && (!local || (ik && ik[0] == '2')) && !(b->use_fused && b->fplan.ok) && b->D.nlinkslots > 0;
    // two-range link pass (interior links before the receive) as two launches: opt-in — measured no faster than send-early /
    // receive-late alone (4 GPUs 359 vs 353, 8 GPUs 216 vs 211 us per step): the second link launch costs what the hidden wait saves
    const bool two_launch = !H.inkernel && b->D.nlinkslots > 0 && getenv("VX3_HALO_OVERLAP") && getenv("VX3_HALO_OVERLAP")[0] == '1';
    if (H.inkernel || two_launch) {
        std::vector<int2> ends;
        std::vector<int32_t> vflags;
        int rc;
        if ((rc = d2h(b, ends, b->D.lends, 0, (size_t)b->D.nlinkslots))) return rc;
        if ((rc = d2h(b, vflags, b->D.vflags, 0, (size_t)b->D.nvox))) return rc;
        CK(cudaStreamSynchronize(b->stream));
        int first_face = b->D.nlinkslots;
        for (int g = 0; g < b->D.nlinkslots; g++) {
            if (ends[g].x < 0) continue;
            if ((vflags[ends[g].x] | vflags[ends[g].y]) & VX3_VOX_GHOST) {
                first_face = g;
                break;
            }
        }
        H.face_tile0 = first_face / VX3_LINK_T; // the boundary tile goes with the face range
    }
    if (H.inkernel) {
        int rc;
        if (!H.state && (rc = b->alloc(&H.state, 128))) return rc;
        if (!H.hin && (rc = b->alloc(&H.hin, 1))) return rc;
        HaloIn hi;
        memset(&hi, 0, sizeof(hi));
        hi.face_tile0 = H.face_tile0;
        std::vector<int32_t> grow((size_t)b->D.nvox, -1);
        for (int sd = 0; sd < 2; sd++) {
            const HaloSide &h = H.side[sd];
            if (h.n_recv > 0 && (h.peer_open || h.peer_local)) {
                hi.n[sd] = h.n_recv; hi.idx[sd] = h.recv_idx; hi.buf[sd] = h.recv_buf; hi.flag[sd] = h.recv_flag;
                std::vector<int32_t> ri;
                if ((rc = d2h(b, ri, h.recv_idx, 0, (size_t)h.n_recv))) return rc;
                CK(cudaStreamSynchronize(b->stream));
                for (int k = 0; k < h.n_recv; k++) grow[ri[k]] = (sd << 30) | k;
            }
        }
        if (!H.ghost_row && (rc = b->upload(&H.ghost_row, grow))) return rc;
        hi.ghost_row = H.ghost_row;
        hi.seq = H.seq; hi.state = H.state; hi.err = H.err; hi.spin_cycles = H.spin_cycles;
        CK(cudaMemcpyAsync(H.hin, &hi, sizeof(hi), cudaMemcpyHostToDevice, b->stream));
        CK(cudaStreamSynchronize(b->stream));
        b->D.hin = H.hin;
        b->link_queue = 1; // the deferred variant carries the collect
        graph_invalidate(b);
        // ---- the send side in the voxel pass ----
        const char *sf = getenv("VX3_HALO_SENDFUSED");
        bool fuse_send = !(sf && sf[0] == '0') && !b->any_secondary && !b->any_signals && H.send_blocks > 0;
        std::vector<int4> vc4;
        if (fuse_send) {
            if ((rc = d2h(b, vc4, b->D.vc4, 0, (size_t)b->D.nvox))) return rc;
            CK(cudaStreamSynchronize(b->stream));
            for (auto &c : vc4) c.w = 0;
            for (int sd = 0; sd < 2 && fuse_send; sd++) {
                const HaloSide &h = H.side[sd];
                if (!(h.n_send > 0 && (h.peer_open || h.peer_local))) continue;
                std::vector<int32_t> si;
                if ((rc = d2h(b, si, h.send_idx, 0, (size_t)h.n_send))) return rc;
                CK(cudaStreamSynchronize(b->stream));
                for (int k = 0; k < h.n_send; k++) {
                    if (vc4[si[k]].w) { // a voxel on both faces (a slab one voxel thick): the stand-alone send kernel handles that
                        fuse_send = false;
                        break;
                    }
                    vc4[si[k]].w = ((sd << 30) | k) + 1;
                }
            }
        }
        if (fuse_send) {
            if (!H.out_count && (rc = b->alloc(&H.out_count, 1))) return rc;
            if (!H.hout && (rc = b->alloc(&H.hout, 1))) return rc;
            HaloOut ho;
            memset(&ho, 0, sizeof(ho));
            for (int sd = 0; sd < 2; sd++) {
                const HaloSide &h = H.side[sd];
                if (h.n_send > 0 && (h.peer_open || h.peer_local)) {
                    ho.buf[sd] = h.peer_buf; ho.flag[sd] = h.peer_flag; ho.n[sd] = h.n_send;
                }
            }
            ho.seq = H.seq; ho.count = H.out_count;
            CK(cudaMemcpyAsync(H.hout, &ho, sizeof(ho), cudaMemcpyHostToDevice, b->stream));
            CK(cudaMemcpyAsync(const_cast<int4 *>(b->D.vc4), vc4.data(), sizeof(int4) * vc4.size(), cudaMemcpyHostToDevice, b->stream));
            CK(cudaStreamSynchronize(b->stream));
            b->D.hout = H.hout;
            H.send_fused = true;
        }
    }
    return VX3_OK;
}

// last: the final step of a stepping call — the fused kernel then also writes the end forces of interior links to HBM,
// where a state read-back expects them
static void launch_step(vx3_batch *b, bool check_stop, bool last) {
    const Dev &D = b->D;
    cudaStream_t st = b->stream;
    const bool fused = b->use_fused && b->fplan.ok;
    if (fused) {
        halo_wait(b);
        const FusedPlan &f = b->fplan;
        if (f.nface > 0) {
            if (b->link_smtab) LAUNCH_SM(KC_LINKS_FACE, (k_links<true, true>), f.face_grid, VX3_LINK_T, 0, D, f.face_tiles, f.face_slot, f.face_c4, f.nface);
            else LAUNCH_SM(KC_LINKS_FACE, (k_links<false, true>), f.face_grid, VX3_LINK_T, 0, D, f.face_tiles, f.face_slot, f.face_c4, f.nface);
        }
        const FusedArgs a{f.blk, f.nblocks};
        if (f.smtab) {
            if (last) LAUNCH_SM(KC_FUSED, (k_fused<true, true>), f.grid, VX3_FUSE_T, f.smem, D, a);
            else LAUNCH_SM(KC_FUSED, (k_fused<true, false>), f.grid, VX3_FUSE_T, f.smem, D, a);
        } else {
            if (last) LAUNCH_SM(KC_FUSED, (k_fused<false, true>), f.grid, VX3_FUSE_T, f.smem, D, a);
            else LAUNCH_SM(KC_FUSED, (k_fused<false, false>), f.grid, VX3_FUSE_T, f.smem, D, a);
        }
    } else if (D.nlinkslots > 0) {
        if (b->halo.on && b->halo.inkernel) launch_links(b, 0, -1, true);
        else if (b->halo.on && b->halo.face_tile0 > 0) { // interior links first: they read no ghost pose, the exchange of the previous step may still be in flight
            launch_links(b, 0, b->halo.face_tile0);
            halo_wait(b);
            launch_links(b, b->halo.face_tile0, b->link_tiles);
        } else {
            halo_wait(b);
            launch_links(b);
        }
    }
    if (b->any_collide) {
        cudaMemsetAsync(D.cell_cnt, 0, 2 * sizeof(int32_t) * ((size_t)D.hmask + 1), st); // every bucket empty (counts and overflow heads)
        LAUNCH(KC_GRID_BUILD, k_grid_build, cdiv(D.nvox, VX3_BLOCK), VX3_BLOCK, D);
        LAUNCH_P(KC_CONTACT, k_contact, cdiv(D.nvox, VX3_CONTACT_WARPS), 32 * VX3_CONTACT_WARPS, 0, D);
    } else if (b->any_detach || b->any_secondary) { // keep the surface flags current (regenerateSurfaceVoxels after a detach / removal)
        LAUNCH(KC_SURFACE, k_surface, cdiv(D.nvox, VX3_BLOCK), VX3_BLOCK, D);
    }
    // attach resolution, then detach, one CTA per simulation (both usually find empty lists and leave at once)
    if ((b->any_sticky || b->any_detach) && D.nlinkslots > 0) LAUNCH_P(KC_RESOLVE, k_resolve_detach, b->nsims, VX3_RESOLVE_T, 0, D);
    const bool com = !b->capturing && com_step(b, b->hsteps + 1);
    int tail_in_voxels = -1; // >= 0: the voxel pass's last CTA does the end-of-step bookkeeping
    if (fused) {
    } else if (b->halo.on && b->halo.send_fused) { // the voxel pass sends the face poses and, on a plain step, does the end-of-step bookkeeping
        const int tail = com ? -1 : (check_stop ? 1 : 0);
        if (b->vox_smtab) LAUNCH_P(KC_VOXELS, (k_voxels<true, true>), b->vox_grid, VX3_VOX_T, 0, D, b->vox_tiles, tail);
        else LAUNCH_P(KC_VOXELS, (k_voxels<false, true>), b->vox_grid, VX3_VOX_T, 0, D, b->vox_tiles, tail);
    } else {
        tail_in_voxels = !com && b->tail_in_voxels ? (check_stop ? 1 : 0) : -1;
        if (b->vox_smtab) LAUNCH_P(KC_VOXELS, (k_voxels<true, false>), b->vox_grid, VX3_VOX_T, 0, D, b->vox_tiles, tail_in_voxels);
        else LAUNCH_P(KC_VOXELS, (k_voxels<false, false>), b->vox_grid, VX3_VOX_T, 0, D, b->vox_tiles, tail_in_voxels);
    }
    if (b->any_signals) LAUNCH(KC_SIGNALS, k_signals, b->nsims, 256, D); // end of timeStep (VX3_Voxel.cu:270-275), before removeVoxels
    if (b->any_secondary) LAUNCH(KC_SECONDARY, k_secondary, cdiv(D.nvox, VX3_BLOCK), VX3_BLOCK, D);
    if (com) LAUNCH(KC_COM, k_com_partial, D.nchunks, VX3_BLOCK, D);
    // end-of-step bookkeeping: k_tail on a sampling step; otherwise inside the voxel pass (its last CTA: slab batches that send from
    // the voxel pass, and any batch of <= 128 simulations), inside the stand-alone send kernel of a slab batch, or k_tail_light
    int tail_in_send = -1;
    const bool sent_by_voxel_pass = !fused && b->halo.on && b->halo.send_fused;
    if (com) LAUNCH(KC_TAIL, k_tail, b->nsims, 128, D, 1, check_stop ? 1 : 0);
    else if (sent_by_voxel_pass || tail_in_voxels >= 0) {
    } else if (b->halo.on && b->halo.send_blocks > 0) tail_in_send = check_stop ? 1 : 0;
    else LAUNCH_P(KC_TAIL, k_tail_light, cdiv(b->nsims, 128), 128, 0, D, check_stop ? 1 : 0);
    if (sent_by_voxel_pass) b->halo.pending = true;
    else if (b->halo.on) { // my face poses to the neighbour slabs (vx3_halo.cuh); their poses are collected before the next face-link pass
        Halo &H = b->halo;
        HaloSendArgs a;
        memset(&a, 0, sizeof(a));
        for (int sd = 0; sd < 2; sd++) {
            HaloSide &h = H.side[sd];
            if (h.n_send > 0 && (h.peer_open || h.peer_local)) {
                a.idx[sd] = h.send_idx; a.peer_buf[sd] = h.peer_buf; a.peer_flag[sd] = h.peer_flag; a.count[sd] = h.send_count;
                a.n[sd] = h.n_send;
                a.nb[sd] = std::min(1024, cdiv(4 * h.n_send, VX3_HALO_BLOCK));
            }
        }
        if (a.nb[0] + a.nb[1] > 0) LAUNCH(KC_HALO, k_halo_send, a.nb[0] + a.nb[1], VX3_HALO_BLOCK, D, a, H.seq, tail_in_send);
        H.pending = true;
    }
    b->hsteps++;
}

// ---- CUDA-Graph stretches ----
// The streaming path issues 3-6 kernels per doTimeStep; for small batches the step is launch-bound.  VX3_GRAPH_STEPS plain steps
// (no centre-of-mass sampling step among them, not the last step of a call) are captured once into a graph and replayed:
// one driver call per stretch instead of ~100.  Same kernels, same order, same arguments: bit-identical to per-step launches
// (tests/test_gpu_graph.py).  Off for profiled runs (events around every launch), while the link-pass variant is still being
// timed.  Slab batches qualify too: their exchange kernels take the step number from device memory.  VX3_GRAPH=0 disables it.
#define VX3_GRAPH_STEPS 32
static void graph_invalidate(vx3_batch *b) {
    for (int i = 0; i < 2; i++)
        if (b->graph_exec[i]) {
            cudaGraphExecDestroy(b->graph_exec[i]);
            b->graph_exec[i] = nullptr;
        }
}
static bool graph_eligible(const vx3_batch *b) {
    static const bool enabled = [] {
        const char *e = getenv("VX3_GRAPH");
        return !(e && e[0] == '0');
    }();
    if (!enabled || b->graph_failed || b->prof.on) return false;
    // two slabs driven by ONE host thread (the same-process test set-up) cannot run ahead of each other by a whole stretch: the
    // first one's receive would spin while the second one's graph is still being instantiated
    if (b->halo.on && (b->halo.side[0].peer_local || b->halo.side[1].peer_local)) return false;
    const bool fused = b->use_fused && b->fplan.ok;
    if (!fused && b->D.nlinkslots > 0 && b->link_queue < 0) return false; // launch_links is still timing its two variants
    return true;
}
static void launch_step(vx3_batch *b, bool check_stop, bool last);
static bool graph_ensure(vx3_batch *b, bool check_stop) {
    const int gi = check_stop ? 1 : 0;
    if (b->graph_exec[gi]) return true;
    const long long h0 = b->hsteps, l0 = b->launches;
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        b->graph_failed = true;
        return false;
    }
    b->capturing = true;
    for (int i = 0; i < VX3_GRAPH_STEPS; i++) launch_step(b, check_stop, false);
    b->capturing = false;
    const cudaError_t e1 = cudaStreamEndCapture(b->stream, &g);
    b->graph_launches[gi] = b->launches - l0;
    b->hsteps = h0;
    b->launches = l0;
    cudaError_t e2 = cudaErrorUnknown;
    if (e1 == cudaSuccess && g) e2 = cudaGraphInstantiate(&b->graph_exec[gi], g, 0);
    if (g) cudaGraphDestroy(g);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        cudaGetLastError();
        b->graph_exec[gi] = nullptr;
        b->graph_failed = true;
        return false;
    }
    return true;
}

// advance k steps; the on-chip persistent kernel takes the stretches between CoM sampling steps when the batch
// qualifies (single small collision-free body), the streaming kernels do the rest
static int advance(vx3_batch *b, long long k, bool check_stop) {
    if (b->any_ghost && !b->halo.on && k > 0)
        return fail(VX3_ERR_INVALID, "the batch holds ghost voxels of a decomposed body but its halo exchange is not connected (vx3_batch_halo_connect)");
    if (b->halo.on) {
        int rc = halo_prepare(b);
        if (rc) return rc;
    }
    while (k > 0) {
        if (b->use_persistent && b->pplan.ok) {
            long long nxt = next_com_step(b); // steps until the next sampling step (it must run on the streaming path)
            long long run = (nxt == 0) ? k : std::min(k, nxt - 1);
            if (run > 0) {
                const bool p_ = b->prof.begin(KC_PERSISTENT, b->stream);
                int rc = persistent_run(b->pplan, b->D, b->stream, run, check_stop, &b->launches);
                if (p_) b->prof.end(b->stream);
                if (rc) return fail(VX3_ERR_CUDA, "persistent kernel launch failed");
                b->hsteps += run;
                k -= run;
                continue;
            }
        }
        if (k > VX3_GRAPH_STEPS && graph_eligible(b)) { // a stretch of plain steps, never the last step of the call
            const long long nxt = next_com_step(b);
            if ((nxt == 0 || nxt > VX3_GRAPH_STEPS) && graph_ensure(b, check_stop)) {
                const int gi = check_stop ? 1 : 0;
                if (cudaGraphLaunch(b->graph_exec[gi], b->stream) != cudaSuccess) return fail(VX3_ERR_CUDA, "cudaGraphLaunch failed");
                b->hsteps += VX3_GRAPH_STEPS;
                b->launches += b->graph_launches[gi];
                b->halo.pending = b->halo.on; // the stretch ends with a send
                k -= VX3_GRAPH_STEPS;
                continue;
            }
        }
        launch_step(b, check_stop, k == 1);
        k--;
    }
    halo_wait(b); // the ghosts of the last step are in place when the call returns (read-backs, centre of mass)
    return VX3_OK;
}

static int check_device_errors(vx3_batch *b) {
    std::vector<SimD> h(b->nsims);
    CK(cudaMemcpyAsync(h.data(), b->D.simd, sizeof(SimD) * b->nsims, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    for (int s = 0; s < b->nsims; s++)
        if (h[s].err) return fail(h[s].err, "simulation " + std::to_string(s) + ": device-side capacity/consistency error (link pool, partner list or attach candidates)");
    if (b->halo.on) {
        int herr = 0;
        CK(cudaMemcpy(&herr, b->halo.err, sizeof(int), cudaMemcpyDeviceToHost));
        if (herr) return fail(VX3_ERR_CUDA, "halo exchange: a neighbour rank did not publish its face poses in time");
    }
    return VX3_OK;
}

// ---- divergence inside a persistent launch ----
// doTimeStep returns false BEFORE any voxel moves in the diverging step (VX3_VoxelyzeKernel.cu:273-281).  The persistent
// kernel's CTAs run up to one step apart, so when one of them sees a link diverge, others have already integrated that step:
// step count, time and status are exact, the voxel state is not.  The guard copies the batch's arena aside before a stepping
// call that will use the persistent path (one device-to-device copy of a few MB per call) and, if the call ends DIVERGED,
// puts it back and replays the stretch on the streaming path, which shares the physics code bit for bit and stops exactly
// like the reference.
static int persist_guard_begin(vx3_batch *b, long long k) {
    b->snap_valid = false;
    if (!(b->use_persistent && b->pplan.ok) || k <= 0 || b->arena_bytes == 0 || !b->snap) return VX3_OK;
    CK(cudaMemcpyAsync(b->snap, b->res.d, b->arena_bytes, cudaMemcpyDeviceToDevice, b->stream));
    b->snap_valid = true;
    b->snap_hsteps = b->hsteps;
    return VX3_OK;
}
static int fetch_simd(vx3_batch *b, std::vector<SimD> &h);
static int persist_guard_end(vx3_batch *b, bool check_stop) {
    if (!b->snap_valid) return VX3_OK;
    b->snap_valid = false;
    std::vector<SimD> h;
    int rc = fetch_simd(b, h);
    if (rc) return rc;
    if (h[0].status != VX3_SIM_DIVERGED) return VX3_OK;
    const long long steps_div = h[0].steps;
    const long long hsteps_end = b->hsteps;
    CK(cudaMemcpyAsync(b->res.d, b->snap, b->arena_bytes, cudaMemcpyDeviceToDevice, b->stream));
    rc = fetch_simd(b, h);
    if (rc) return rc;
    const long long replay = steps_div - h[0].steps; // doTimeStep calls up to and including the one that returned false
    b->hsteps = b->snap_hsteps;
    const bool up = b->use_persistent;
    b->use_persistent = false;
    rc = replay > 0 ? advance(b, replay, check_stop) : VX3_OK;
    b->use_persistent = up;
    b->hsteps = hsteps_end; // the host-side call counter keeps counting calls (finished simulations ignore them)
    if (rc) return rc;
    CK(cudaStreamSynchronize(b->stream));
    return VX3_OK;
}

extern "C" int vx3_batch_step_dt(vx3_batch *b, int64_t k, float dt) {
    if (!b || k < 0) return fail(VX3_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(b->device));
    const long long l0 = b->launches;
    set_dt(b, dt);
    CK(cudaEventRecord(b->ev0, b->stream));
    int rc = persist_guard_begin(b, k);
    if (rc) return rc;
    rc = advance(b, k, false);
    if (rc) return rc;
    CK(cudaEventRecord(b->ev1, b->stream));
    rc = persist_guard_end(b, false); // (after the timing event: the check reads the simulation's scalars back)
    if (rc) return rc;
    CK(cudaEventSynchronize(b->ev1));
    CK(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, b->ev0, b->ev1);
    b->last_ms = ms;
    b->last_launches = b->launches - l0;
    b->prof.collect();
    return check_device_errors(b);
}

extern "C" int vx3_batch_step(vx3_batch *b, int64_t k) { return vx3_batch_step_dt(b, k, -1.0f); }

// queue k steps without waiting for them (vx3_batch_sync waits): lets one host thread drive several batches whose
// step streams depend on each other (slabs of a decomposed body)
extern "C" int vx3_batch_step_async(vx3_batch *b, int64_t k, float dt) {
    if (!b || k < 0) return fail(VX3_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(b->device));
    set_dt(b, dt);
    int rc = advance(b, k, false);
    if (rc) return rc;
    CK(cudaGetLastError());
    return VX3_OK;
}

extern "C" int vx3_batch_sync(vx3_batch *b) {
    if (!b) return fail(VX3_ERR_INVALID, "batch is NULL");
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    if (b->halo.on) return check_device_errors(b); // a halo wait that ran into its spin limit must not pass silently
    return VX3_OK;
}

static int fetch_simd(vx3_batch *b, std::vector<SimD> &h) {
    h.resize(b->nsims);
    CK(cudaMemcpyAsync(h.data(), b->D.simd, sizeof(SimD) * b->nsims, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return VX3_OK;
}

extern "C" int vx3_batch_run(vx3_batch *b, const vx3_run_opts *opts, vx3_history_cb cb, void *user) {
    if (!b) return fail(VX3_ERR_INVALID, "batch is NULL");
    CK(cudaSetDevice(b->device));
    long long max_steps = (opts && opts->max_steps > 0) ? opts->max_steps : 1000000;
    int chunk = (opts && opts->steps_per_launch > 0) ? opts->steps_per_launch : 0;
    const bool history = opts && opts->emit_history && cb;
    const long long l0 = b->launches;
    set_dt(b, -1.0f);
    std::vector<SimD> h;
    int rc = fetch_simd(b, h);
    if (rc) return rc;
    // history frame cadence (VX3_SimulationManager.cu:56,71)
    std::vector<long long> frame_every(b->nsims, 0);
    HistoryWriter hw;
    // CurStepCount at the start of this run: the loop index j of CUDA_Simulation restarts at 0 with every run, the device
    // counter does not (earlier vx3_batch_step / _run calls)
    std::vector<long long> steps0(b->nsims);
    for (int s = 0; s < b->nsims; s++) {
        steps0[s] = h[s].steps;
        // the CoM sampling cadence is scheduled from the host's call counter and executed on the device's CurStepCount
        if (h[s].status == VX3_SIM_RUNNING && h[s].steps != b->hsteps)
            return fail(VX3_ERR_INVALID, "internal: host and device step counters of simulation " + std::to_string(s) + " disagree");
    }
    std::vector<char> announced_end(b->nsims, 0);
    char line[768];
    if (history) {
        // the stdout of CUDA_Simulation up to its step loop, per simulation (VX3_SimulationManager.cu:25, :40-50, :56-58)
        for (int s = 0; s < b->nsims; s++) {
            const vx3_sim_options &o = b->opts[s];
            std::string pre;
            snprintf(line, sizeof(line), "\033[0;32m%d) Simulation %d runs: %s.\n\033[0m", b->device, s, b->names[s].c_str());
            pre += line;
            // recommendedTimeStep() called BEFORE the loop (:56-58): the model as imported — the OptimalDt the steps use is evaluated in
            // the first doTimeStep and differs when a link material has nu != 0 (vx3_model_first_step_dt)
            const double rec = (size_t)s < b->model_dt.size() ? b->model_dt[(size_t)s] : b->simc[s].optimal_dt;
            const int real_stepsize = (int)(o.record_step_size / (10000 * rec * o.dt_frac)) + 1;
            if (o.record_step_size) {
                frame_every[s] = real_stepsize;
                pre += hw.header(b->matid[s], b->matcolor[s], o.vox_size);
            }
            // recommendedTimeStep() warns on a model without links every time it is called (VX3_VoxelyzeKernel.cu:189-191): twice
            // around this line, once more from the first doTimeStep (:245-247)
            if (b->simc[s].nhostlinks == 0) pre += "WARNING: No links.\nWARNING: No links.\n";
            snprintf(line, sizeof(line), "real_stepsize: %d ; recommendedTimeStep %f; d_v3->DtFrac %f . \n", real_stepsize, rec, o.dt_frac);
            pre += line;
            cb(user, s, pre.data(), pre.size());
        }
    }
    CK(cudaEventRecord(b->ev0, b->stream));
    k_sim_update<<<cdiv(b->nsims, 128), 128, 0, b->stream>>>(b->D, 2); // StopConditionMet() before the first step
    b->launches++;
    long long j = 0; // loop index of CUDA_Simulation (:62); all simulations start this run at j = 0
    bool first_chunk = true;
    if (chunk <= 0) {
        long long perstep = std::max<long long>(1, (long long)b->D.nvox + b->D.nlinkslots);
        chunk = (int)std::max<long long>(16, std::min<long long>(4096, 40000000 / perstep));
    }
    while (j < max_steps) {
        // frames are emitted at the top of loop iteration j when j % real_stepsize == 0, after that iteration's step (:70-114)
        long long todo = std::min<long long>(chunk, max_steps - j);
        if (history) {
            for (int s = 0; s < b->nsims; s++)
                if (frame_every[s] > 0) {
                    long long nf = frame_every[s] - (j % frame_every[s]); // steps until a frame step is completed
                    if (j % frame_every[s] == 0) nf = 1;
                    todo = std::min(todo, nf);
                }
        }
        rc = persist_guard_begin(b, todo);
        if (rc) return rc;
        rc = advance(b, todo, true);
        if (rc) return rc;
        rc = persist_guard_end(b, true);
        if (rc) return rc;
        j += todo;
        rc = fetch_simd(b, h);
        if (rc) return rc;
        bool running = false;
        for (int s = 0; s < b->nsims; s++) {
            if (h[s].err) return fail(h[s].err, "simulation " + std::to_string(s) + ": device-side capacity/consistency error");
            running |= h[s].status == VX3_SIM_RUNNING;
        }
        if (history && first_chunk) {
            for (int s = 0; s < b->nsims; s++)
                if (b->simc[s].nhostlinks == 0 && h[s].steps > steps0[s]) cb(user, s, "WARNING: No links.\n", 19);
        }
        first_chunk = false;
        if (history) {
            for (int s = 0; s < b->nsims; s++) {
                if (frame_every[s] <= 0) continue;
                // iteration jj = j-1 just completed its doTimeStep; the reference prints when jj % real_stepsize == 0
                // and the simulation executed that step (it breaks out of the loop before printing otherwise)
                const long long jj = j - 1;
                if (jj % frame_every[s] != 0) continue;
                if (h[s].steps != steps0[s] + j) continue; // stopped or diverged earlier
                if (h[s].status == VX3_SIM_DIVERGED) continue;
                std::string fr;
                rc = history_frame(b, s, jj, h[s].t, hw, fr);
                if (rc) return rc;
                cb(user, s, fr.data(), fr.size());
            }
        }
        if (history) { // "Diverged" is printed inside the loop, at the step that failed (:65-69)
            for (int s = 0; s < b->nsims; s++)
                if (h[s].status == VX3_SIM_DIVERGED && !announced_end[s]) {
                    announced_end[s] = 1;
                    snprintf(line, sizeof(line), "\033[1;31m\n\n%d) Simulation %d Diverged: %s.\n\033[0m", b->device, s, b->names[s].c_str());
                    cb(user, s, line, strlen(line));
                }
        }
        if (!running) break;
    }
    if (j >= max_steps) {
        k_step_cap<<<cdiv(b->nsims, 128), 128, 0, b->stream>>>(b->D);
        b->launches++;
    }
    run_com(b, 1); // updateCurrentCenterOfMass + computeFitness (:116-117)
    if (history) { // the closing line of CUDA_Simulation (:118-119)
        rc = fetch_simd(b, h);
        if (rc) return rc;
        for (int s = 0; s < b->nsims; s++) {
            snprintf(line, sizeof(line), "\033[0;34m%d) Simulation %d ends: %s Time: %f, angleSampleTimes: %d.\n\033[0m", b->device, s, b->names[s].c_str(),
                     h[s].t, h[s].angle_samples);
            cb(user, s, line, strlen(line));
        }
    }
    CK(cudaEventRecord(b->ev1, b->stream));
    CK(cudaEventSynchronize(b->ev1));
    CK(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, b->ev0, b->ev1);
    b->last_ms = ms;
    b->last_launches = b->launches - l0;
    b->prof.collect();
    return VX3_OK;
}

extern "C" int vx3_batch_set_profiling(vx3_batch *b, int on, int use_persistent) {
    if (!b) return fail(VX3_ERR_INVALID, "batch is NULL");
    CK(cudaSetDevice(b->device));
    b->use_persistent = use_persistent != 0;
    b->prof.on = on != 0;
    if (on && b->prof.ev.empty()) {
        b->prof.ev.resize(16384);
        for (auto &e : b->prof.ev) CK(cudaEventCreate(&e));
    }
    for (int k = 0; k < KC_COUNT; k++) {
        b->prof.ms[k] = 0;
        b->prof.cnt[k] = 0;
    }
    return VX3_OK;
}

extern "C" int vx3_batch_kernel_stats(vx3_batch *b, int index, char *name, int name_cap, double *total_ms, int64_t *launches) {
    if (!b) return fail(VX3_ERR_INVALID, "batch is NULL");
    if (index < 0 || index >= KC_COUNT) return 1; // end of list
    if (name && name_cap > 0) {
        strncpy(name, kKernelNames[index], name_cap - 1);
        name[name_cap - 1] = 0;
    }
    if (total_ms) *total_ms = b->prof.ms[index];
    if (launches) *launches = b->prof.cnt[index];
    return VX3_OK;
}

extern "C" int vx3_batch_set_fused(vx3_batch *b, int on) {
    if (!b) return fail(VX3_ERR_INVALID, "batch is NULL");
    b->use_fused = on != 0;
    graph_invalidate(b);
    return VX3_OK;
}

extern "C" int vx3_batch_fused_info(vx3_batch *b, int32_t *out4) {
    if (!b || !out4) return fail(VX3_ERR_INVALID, "bad arguments");
    out4[0] = b->use_fused && b->fplan.ok;
    out4[1] = b->fplan.nblocks;
    out4[2] = b->fplan.ninterior;
    out4[3] = b->fplan.nface;
    return VX3_OK;
}

// Host-only: the block layout of the models and its consistency counts (test hook, no device needed).
extern "C" int vx3_fused_plan_check(const vx3_model_desc *models, int n, int max_block_voxels, int64_t *out8) {
    if (!models || n <= 0 || !out8) return fail(VX3_ERR_INVALID, "bad arguments");
    for (int i = 0; i < n; i++) {
        int rc = validate_model(models[i], i);
        if (rc) return rc;
    }
    std::vector<SimC> simc(n);
    size_t nvox = 0, nslots = 0;
    for (int s = 0; s < n; s++) {
        memset(&simc[s], 0, sizeof(SimC));
        simc[s].voff = (int)nvox;
        simc[s].nvox = models[s].n_voxels;
        simc[s].loff = (int)nslots;
        simc[s].lcap = models[s].n_links;
        nvox += models[s].n_voxels;
        nslots += models[s].n_links;
    }
    FusedLayout L;
    const int bv = max_block_voxels > 0 ? std::min(max_block_voxels, VX3_FUSE_BV) : VX3_FUSE_BV;
    if (!fused_layout_build(models, n, simc, bv, L)) return fail(VX3_ERR_INVALID, "the models do not fit the block model");
    // permutations: every device index of a simulation's range is hit exactly once
    std::vector<int> vseen(nvox, 0), lseen(nslots, 0), blk_of_dev(nvox, -1);
    int64_t bad = 0, largest = 0;
    for (int s = 0; s < n; s++) {
        for (int i = 0; i < simc[s].nvox; i++) {
            const int d = L.vperm[(size_t)simc[s].voff + i];
            if (d < simc[s].voff || d >= simc[s].voff + simc[s].nvox) bad++;
            else vseen[d]++;
        }
        for (int i = 0; i < simc[s].lcap; i++) {
            const int d = L.lperm[(size_t)simc[s].loff + i];
            if (d < simc[s].loff || d >= simc[s].loff + simc[s].lcap) bad++;
            else lseen[d]++;
        }
    }
    // blocks: consecutive voxel ranges inside one simulation, within the size bound
    int vnext = 0;
    for (size_t b = 0; b < L.blk.size(); b++) {
        const int4 d = L.blk[b];
        const int nv = (d.z >> 16) & 0xFFFF;
        largest = std::max<int64_t>(largest, nv);
        if (d.x != vnext || nv > bv || d.x < simc[d.w].voff || d.x + nv > simc[d.w].voff + simc[d.w].nvox) bad++;
        vnext = d.x + nv;
        for (int k = 0; k < nv; k++) blk_of_dev[(size_t)d.x + k] = (int)b;
    }
    if (vnext != (int)nvox) bad++;
    // links: a slot in a block's interior range joins two voxels of that block; a slot in a face range joins two blocks
    std::vector<int> owner(nslots, -1); // block of an interior slot
    int64_t ninterior = 0;
    for (size_t b = 0; b < L.blk.size(); b++) {
        const int4 d = L.blk[b];
        for (int i = 0; i < (d.z & 0xFFFF); i++) owner[(size_t)d.y + i] = (int)b;
        ninterior += d.z & 0xFFFF;
    }
    int64_t nface = 0;
    for (int s = 0; s < n; s++) {
        const vx3_model_desc &m = models[s];
        nface += L.face_range[s].y;
        for (int i = 0; i < m.n_links; i++) {
            const int d = L.lperm[(size_t)simc[s].loff + i];
            const int bn = blk_of_dev[L.vperm[(size_t)simc[s].voff + m.link_vneg[i]]], bp = blk_of_dev[L.vperm[(size_t)simc[s].voff + m.link_vpos[i]]];
            const bool in_face = d >= L.face_range[s].x && d < L.face_range[s].x + L.face_range[s].y;
            if (owner[d] >= 0 ? (bn != owner[d] || bp != owner[d] || in_face) : (bn == bp || !in_face)) bad++;
        }
    }
    if (ninterior != L.ninterior || nface != L.nface) bad++;
    int64_t vok = 0, lok = 0;
    for (size_t v = 0; v < nvox; v++) vok += vseen[v] == 1;
    for (size_t g = 0; g < nslots; g++) lok += lseen[g] == 1;
    out8[0] = (int64_t)L.blk.size();
    out8[1] = ninterior;
    out8[2] = nface;
    out8[3] = (int64_t)nslots;
    out8[4] = (int64_t)nvox;
    out8[5] = largest;
    out8[6] = bad ? -bad : vok;
    out8[7] = lok;
    return VX3_OK;
}

extern "C" int vx3_batch_last_timing(vx3_batch *b, double *ms, int64_t *launches) {
    if (!b) return fail(VX3_ERR_INVALID, "batch is NULL");
    if (ms) *ms = b->last_ms;
    if (launches) *launches = b->last_launches;
    return VX3_OK;
}

extern "C" int vx3_batch_recommended_dt(vx3_batch *b, int sim, double *out) {
    if (!b || sim < 0 || sim >= b->nsims || !out) return fail(VX3_ERR_INVALID, "bad arguments");
    *out = b->simc[sim].optimal_dt;
    return VX3_OK;
}

// ------------------------------------------------------------------ read-back
template <class T> static int d2h(vx3_batch *b, std::vector<T> &h, const T *d, size_t off, size_t n) {
    h.resize(n);
    if (n == 0) return VX3_OK;
    CK(cudaMemcpyAsync(h.data(), d + off, n * sizeof(T), cudaMemcpyDeviceToHost, b->stream));
    return VX3_OK;
}

extern "C" int vx3_batch_state(vx3_batch *b, int sim, vx3_state_view *w) {
    if (!b || !w || sim < 0 || sim >= b->nsims) return fail(VX3_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(b->device));
    const SimC &S = b->simc[sim];
    std::vector<SimD> hd;
    int rc = fetch_simd(b, hd);
    if (rc) return rc;
    const int nv = S.nvox, nl = hd[sim].link_cnt;
    if (w->n_voxels < nv || w->n_links < nl) {
        w->n_voxels = nv;
        w->n_links = nl;
        return fail(VX3_ERR_INVALID, "state view buffers too small");
    }
    w->n_voxels = nv;
    w->n_links = nl;
    const Dev &D = b->D;
    std::vector<double> pose, contact;
    // blocked records: copy the 32-item blocks that cover the simulation's range; v0/l0 = first item of the first block
    const size_t v0 = (size_t)S.voff / 32 * 32, v1 = ((size_t)S.voff + nv + 31) / 32 * 32;
    const size_t l0 = (size_t)S.loff / 32 * 32, l1 = ((size_t)S.loff + nl + 31) / 32 * 32;
    std::vector<double2> mom2, lh2, lf2;
    std::vector<int32_t> vflags, vlinks, lstate, lmat;
    std::vector<float> tempe;
    std::vector<int2> lends;
    std::vector<float4> lstrain;
    if ((rc = d2h(b, pose, D.pose, 8 * (size_t)S.voff, 8 * (size_t)nv))) return rc;
    if ((rc = d2h(b, mom2, (const double2 *)D.mom2, 3 * v0, 3 * (v1 - v0)))) return rc;
    if ((rc = d2h(b, vflags, D.vflags, S.voff, nv))) return rc;
    if ((rc = d2h(b, tempe, D.tempe, S.voff, nv))) return rc;
    if ((rc = d2h(b, vlinks, D.vlinks, 6 * (size_t)S.voff, 6 * (size_t)nv))) return rc;
    if (D.contact && (rc = d2h(b, contact, D.contact, 3 * (size_t)S.voff, 3 * (size_t)nv))) return rc;
    std::vector<double> sig;
    if (D.sig && w->signal && (rc = d2h(b, sig, D.sig, 6 * (size_t)S.voff, 6 * (size_t)nv))) return rc;
    if ((rc = d2h(b, lends, D.lends, S.loff, nl))) return rc;
    if ((rc = d2h(b, lstate, D.lstate, S.loff, nl))) return rc;
    if ((rc = d2h(b, lmat, D.lmat, S.loff, nl))) return rc;
    if (nl > 0 && (rc = d2h(b, lh2, (const double2 *)D.lh2, 5 * l0, 5 * (l1 - l0)))) return rc;
    if (nl > 0 && (rc = d2h(b, lf2, (const double2 *)D.lf2, 6 * l0, 6 * (l1 - l0)))) return rc;
    if ((rc = d2h(b, lstrain, D.lstrain, S.loff, nl))) return rc;
    CK(cudaStreamSynchronize(b->stream));
    // the arrays are in the batch's storage order; the view is in the model's (vdev / ldev / vext / lext)
    for (int i = 0; i < nv; i++) {
        const size_t vd = (size_t)(b->vdev((size_t)S.voff + i) - S.voff);
        for (int k = 0; k < 3; k++) {
            if (w->pos) w->pos[3 * i + k] = pose[8 * vd + k];
            const size_t vb = (size_t)S.voff + vd - v0; // index relative to the copied blocks
            const double mo[6] = {mom2[idx_mo(0, vb)].x, mom2[idx_mo(0, vb)].y, mom2[idx_mo(1, vb)].x, mom2[idx_mo(1, vb)].y, mom2[idx_mo(2, vb)].x, mom2[idx_mo(2, vb)].y};
            if (w->lin_mom) w->lin_mom[3 * i + k] = mo[k];
            if (w->ang_mom) w->ang_mom[3 * i + k] = mo[3 + k];
            if (w->contact_force) w->contact_force[3 * i + k] = contact.empty() ? 0.0 : contact[3 * vd + k];
        }
        if (w->orient) for (int k = 0; k < 4; k++) w->orient[4 * i + k] = pose[8 * vd + 3 + k];
        if (w->vox_flags) w->vox_flags[i] = vflags[vd] & VXF_BOOLSTATE_MASK;
        if (w->temp) w->temp[i] = tempe[vd];
        if (w->signal) for (int k = 0; k < 6; k++) w->signal[6 * i + k] = sig.empty() ? 0.0 : sig[6 * vd + k];
        if (w->vox_links) for (int k = 0; k < 6; k++) w->vox_links[6 * i + k] = vlinks[6 * vd + k] >= 0 ? b->lext((size_t)vlinks[6 * vd + k]) - S.loff : -1;
    }
    // global link-material index -> the simulation's local index (attach-created materials follow the model's)
    std::map<int, int> lm_local;
    for (size_t i = 0; i < b->lmat_global[sim].size(); i++) lm_local.emplace(b->lmat_global[sim][i], (int)i);
    int next_local = (int)b->lmat_global[sim].size();
    for (int i = 0; i < nl; i++) {
        const size_t ld = (size_t)(b->ldev((size_t)S.loff + i) - S.loff);
        if (w->link_vneg) w->link_vneg[i] = lends[ld].x >= 0 ? b->vext((size_t)lends[ld].x) - S.voff : lends[ld].x - S.voff;
        if (w->link_vpos) w->link_vpos[i] = lends[ld].y >= 0 ? b->vext((size_t)lends[ld].y) - S.voff : lends[ld].y - S.voff;
        if (w->link_axis) w->link_axis[i] = (lstate[ld] & LKS_AXIS_MASK) >> LKS_AXIS_SHIFT;
        if (w->link_mat) {
            auto it = lm_local.find(lmat[ld]);
            if (it == lm_local.end()) it = lm_local.emplace(lmat[ld], next_local++).first;
            w->link_mat[i] = it->second;
        }
        const size_t lb = (size_t)S.loff + ld - l0;
        const double h[10] = {lh2[idx_lh(0, lb)].x, lh2[idx_lh(0, lb)].y, lh2[idx_lh(1, lb)].x, lh2[idx_lh(1, lb)].y, lh2[idx_lh(2, lb)].x, lh2[idx_lh(2, lb)].y,
                              lh2[idx_lh(3, lb)].x, lh2[idx_lh(3, lb)].y, lh2[idx_lh(4, lb)].x, lh2[idx_lh(4, lb)].y};
        double fn[6], fp[6];
        for (int p = 0; p < 3; p++) {
            const double2 a = lf2[idx_lf(p, lb)], c = lf2[idx_lf(3 + p, lb)];
            fn[2 * p] = a.x; fn[2 * p + 1] = a.y;
            fp[2 * p] = c.x; fp[2 * p + 1] = c.y;
        }
        for (int k = 0; k < 3; k++) {
            if (w->link_pos2) w->link_pos2[3 * i + k] = h[k];
            if (w->link_angle1v) w->link_angle1v[3 * i + k] = h[3 + k];
            if (w->link_angle2v) w->link_angle2v[3 * i + k] = h[6 + k];
            if (w->link_force_neg) w->link_force_neg[3 * i + k] = fn[k];
            if (w->link_moment_neg) w->link_moment_neg[3 * i + k] = fn[3 + k];
            if (w->link_force_pos) w->link_force_pos[3 * i + k] = fp[k];
            if (w->link_moment_pos) w->link_moment_pos[3 * i + k] = fp[3 + k];
        }
        if (w->link_strain) w->link_strain[i] = lstrain[ld].x;
        if (w->link_max_strain) w->link_max_strain[i] = lstrain[ld].y;
        if (w->link_strain_offset) w->link_strain_offset[i] = lstrain[ld].z;
        if (w->link_stress) w->link_stress[i] = lstrain[ld].w;
        if (w->link_flags) w->link_flags[i] = lstate[ld] & LKS_PUBLIC_MASK;
        if (w->link_rest_length) w->link_rest_length[i] = h[9];
    }
    return VX3_OK;
}

extern "C" int vx3_batch_results(vx3_batch *b, vx3_result *out) {
    if (!b || !out) return fail(VX3_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(b->device));
    run_com(b, 1);
    std::vector<SimD> h;
    int rc = fetch_simd(b, h);
    if (rc) return rc;
    for (int s = 0; s < b->nsims; s++) {
        vx3_result &r = out[s];
        memset(&r, 0, sizeof(r));
        strncpy(r.name, b->names[s].c_str(), sizeof(r.name) - 1);
        const SimD &d = h[s];
        r.status = d.status;
        r.num_voxel = b->simc[s].nvox;
        r.num_measured_voxel = d.n_measured;
        r.num_close_pairs = d.num_close_pairs;
        r.steps = d.steps;
        r.num_links = d.link_cnt;
        r.collision_count = d.collision_count;
        r.current_time = d.t;
        r.fitness_score = d.status == VX3_SIM_DIVERGED ? NAN : d.fitness;
        r.vox_size = b->simc[s].vox_size;
        for (int k = 0; k < 3; k++) {
            r.initial_com[k] = d.com0[k];
            r.current_com[k] = d.com[k];
        }
        r.total_distance_of_all_voxels = d.total_dist;
        r.recent_angle = d.recent_angle;
        r.target_closeness = d.target_closeness;
        r.dt = d.dt;
    }
    return VX3_OK;
}

extern "C" int vx3_batch_positions(vx3_batch *b, int sim, double *init_pos, double *pos, int32_t *mats) {
    if (!b || sim < -1 || sim >= b->nsims) return fail(VX3_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(b->device));
    // sim = -1: every simulation of the batch, concatenated in model order (one device->host copy for the whole batch)
    const int s0 = sim < 0 ? 0 : sim, s1 = sim < 0 ? b->nsims : sim + 1;
    const size_t voff = b->simc[s0].voff;
    size_t nv = 0;
    for (int s = s0; s < s1; s++) nv += b->simc[s].nvox;
    // device -> host through the batch's pinned staging buffer when it is large enough (idle after creation), else pageable
    const size_t pose_bytes = pos ? 8 * nv * sizeof(double) : 0, ip_bytes = init_pos ? 3 * nv * sizeof(double) : 0;
    std::vector<double> pose_v, ip_v;
    const double *pose_h = nullptr, *ip_h = nullptr;
    if (pose_bytes + ip_bytes <= b->res.hcap) {
        if (pos) {
            CK(cudaMemcpyAsync(b->res.h, b->D.pose + 8 * voff, pose_bytes, cudaMemcpyDeviceToHost, b->stream));
            pose_h = reinterpret_cast<const double *>(b->res.h);
        }
        if (init_pos) {
            CK(cudaMemcpyAsync(b->res.h + pose_bytes, b->D.initpos + 3 * voff, ip_bytes, cudaMemcpyDeviceToHost, b->stream));
            ip_h = reinterpret_cast<const double *>(b->res.h + pose_bytes);
        }
    } else {
        int rc;
        if (pos && (rc = d2h(b, pose_v, b->D.pose, 8 * voff, 8 * nv))) return rc;
        if (init_pos && (rc = d2h(b, ip_v, (const double *)b->D.initpos, 3 * voff, 3 * nv))) return rc;
        pose_h = pose_v.data();
        ip_h = ip_v.data();
    }
    CK(cudaStreamSynchronize(b->stream));
    if (pos)
        for (size_t i = 0; i < nv; i++) {
            const size_t vd = (size_t)b->vdev(voff + i) - voff; // storage order -> model order
            for (int k = 0; k < 3; k++) pos[3 * i + k] = pose_h[8 * vd + k];
        }
    if (init_pos) {
        if (b->vperm.empty()) memcpy(init_pos, ip_h, ip_bytes);
        else
            for (size_t i = 0; i < nv; i++) {
                const size_t vd = (size_t)b->vdev(voff + i) - voff;
                for (int k = 0; k < 3; k++) init_pos[3 * i + k] = ip_h[3 * vd + k];
            }
    }
    if (mats) {
        size_t o = 0;
        for (int s = s0; s < s1; s++)
            for (int i = 0; i < b->simc[s].nvox; i++) mats[o++] = b->matid[s][b->vmat_local[s][i]];
    }
    return VX3_OK;
}

// sortResults (VX3_SimulationManager.cu:472) with VX3_SimulationResult::compareFitnessScore
// (VX3_SimulationResult.h:26-33): fitness descending, NaN last.  Host-only.
extern "C" void vx3_sort_results(vx3_result *r, int n) {
    if (!r || n <= 1) return;
    std::stable_sort(r, r + n, [](const vx3_result &a, const vx3_result &b) {
        const bool an = std::isnan(a.fitness_score), bn = std::isnan(b.fitness_score);
        if (an) return false;
        if (bn) return true;
        return a.fitness_score > b.fitness_score;
    });
}

// ------------------------------------------------------------------ slab decomposition (vx3_halo.cuh)
// One receive block per side: [2 flags, 128 bytes apart][2 parities][n_recv][8] doubles, one cudaMalloc, one IPC handle.
static const size_t kHaloFlagBytes = 256;

extern "C" int vx3_batch_halo_setup(vx3_batch *b, int side, int n_send, const int32_t *send_vox, int n_recv, const int32_t *recv_vox) {
    if (!b || side < 0 || side > 1 || n_send < 0 || n_recv < 0) return fail(VX3_ERR_INVALID, "bad arguments");
    preload_step_kernels();
    if (b->nsims != 1) return fail(VX3_ERR_INVALID, "halo exchange applies to a batch of ONE decomposed body");
    if (b->any_collide) return fail(VX3_ERR_INVALID, "halo exchange: collisions / attach are not supported across slabs");
    CK(cudaSetDevice(b->device));
    HaloSide &h = b->halo.side[side];
    if (h.recv_flag) return fail(VX3_ERR_INVALID, "halo side already set up");
    const int nv = b->simc[0].nvox;
    std::vector<int32_t> si(send_vox, send_vox + n_send), ri(recv_vox, recv_vox + n_recv);
    for (int v : si) if (v < 0 || v >= nv) return fail(VX3_ERR_INVALID, "halo send index out of range");
    for (int v : ri) if (v < 0 || v >= nv) return fail(VX3_ERR_INVALID, "halo receive index out of range");
    for (int &v : si) v = b->vdev((size_t)v); // model index -> storage index (one simulation: voff = 0)
    for (int &v : ri) v = b->vdev((size_t)v);
    int rc;
    if ((rc = b->upload(&h.send_idx, si))) return rc;
    if ((rc = b->upload(&h.recv_idx, ri))) return rc;
    if ((rc = b->alloc(&h.send_count, 1))) return rc;
    unsigned char *blk = nullptr;
    if ((rc = b->alloc(&blk, kHaloFlagBytes + sizeof(double) * 16 * (size_t)std::max(n_recv, 1)))) return rc;
    h.recv_flag = reinterpret_cast<unsigned int *>(blk);
    h.recv_buf = reinterpret_cast<double *>(blk + kHaloFlagBytes);
    h.n_send = n_send;
    h.n_recv = n_recv;
    if (!b->halo.err && (rc = b->alloc(&b->halo.err, 1))) return rc;
    if (!b->halo.seq && (rc = b->alloc(&b->halo.seq, 4))) return rc;
    CK(cudaStreamSynchronize(b->stream));
    return VX3_OK;
}

extern "C" int vx3_batch_halo_export(vx3_batch *b, int side, void *handle64) {
    if (!b || side < 0 || side > 1 || !handle64) return fail(VX3_ERR_INVALID, "bad arguments");
    HaloSide &h = b->halo.side[side];
    if (!h.recv_flag) return fail(VX3_ERR_INVALID, "halo side not set up");
    CK(cudaSetDevice(b->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hd;
    CK(cudaIpcGetMemHandle(&hd, h.recv_flag));
    memcpy(handle64, &hd, 64);
    return VX3_OK;
}

// peer_handle64: what the neighbour on this side exported for ITS side facing me; peer_n_recv must equal my n_send
extern "C" int vx3_batch_halo_connect(vx3_batch *b, int side, const void *peer_handle64, int peer_n_recv) {
    if (!b || side < 0 || side > 1 || !peer_handle64) return fail(VX3_ERR_INVALID, "bad arguments");
    HaloSide &h = b->halo.side[side];
    if (!h.recv_flag) return fail(VX3_ERR_INVALID, "halo side not set up");
    if (peer_n_recv != h.n_send) return fail(VX3_ERR_INVALID, "halo: the neighbour expects a different number of face voxels than this rank sends");
    CK(cudaSetDevice(b->device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, peer_handle64, 64);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h.peer_flag = reinterpret_cast<unsigned int *>(p);
    h.peer_buf = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(p) + kHaloFlagBytes);
    h.peer_open = true;
    b->halo.on = true;
    b->use_persistent = false;
    graph_invalidate(b);
    return VX3_OK;
}

// same-process variant of halo_connect (one host process driving both slabs, e.g. two batches on one device in the tests)
extern "C" int vx3_batch_halo_connect_local(vx3_batch *b, int side, vx3_batch *peer) {
    if (!b || !peer || side < 0 || side > 1) return fail(VX3_ERR_INVALID, "bad arguments");
    HaloSide &h = b->halo.side[side];
    HaloSide &ph = peer->halo.side[1 - side];
    if (!h.recv_flag || !ph.recv_flag) return fail(VX3_ERR_INVALID, "halo side not set up");
    if (ph.n_recv != h.n_send) return fail(VX3_ERR_INVALID, "halo: the neighbour expects a different number of face voxels than this rank sends");
    if (b->device != peer->device) {
        CK(cudaSetDevice(b->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(VX3_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
    }
    h.peer_flag = ph.recv_flag;
    h.peer_buf = ph.recv_buf;
    h.peer_open = false; // nothing to close
    h.peer_local = true;
    b->halo.on = true;
    b->use_persistent = false;
    graph_invalidate(b);
    return VX3_OK;
}

extern "C" int vx3_batch_counters(vx3_batch *b, int sim, int64_t *out8) {
    if (!b || !out8 || sim < 0 || sim >= b->nsims) return fail(VX3_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(b->device));
    std::vector<SimD> h;
    int rc = fetch_simd(b, h);
    if (rc) return rc;
    const SimD &d = h[sim];
    const int64_t v[8] = {d.attach_events, d.detach_events, d.link_cnt, d.nsurface, d.cand_peak, d.fail_peak, 0, 0};
    for (int i = 0; i < 8; i++) out8[i] = v[i];
    return VX3_OK;
}

// raw centre-of-mass sums of one simulation over its OWNED voxels: sum m*x, m*y, m*z, sum m, sum |pos - initial pos|, count
// (a decomposed body's ranks add these up before dividing, updateCurrentCenterOfMass VX3_VoxelyzeKernel.cu:477-493)
extern "C" int vx3_batch_com_sums(vx3_batch *b, int sim, double *out6) {
    if (!b || sim < 0 || sim >= b->nsims || !out6) return fail(VX3_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(b->device));
    k_com_partial<<<b->D.nchunks, VX3_BLOCK, 0, b->stream>>>(b->D);
    b->launches++;
    const SimC &S = b->simc[sim];
    std::vector<double> part;
    int rc = d2h(b, part, (const double *)b->D.com_part, 6 * (size_t)S.chunk_off, 6 * (size_t)S.nchunks);
    if (rc) return rc;
    CK(cudaStreamSynchronize(b->stream));
    for (int k = 0; k < 6; k++) out6[k] = 0;
    for (int c = 0; c < S.nchunks; c++)
        for (int k = 0; k < 6; k++) out6[k] += part[6 * (size_t)c + k];
    return VX3_OK;
}

#include "vx3_history.inl"
