// ertb_cuda.cu -- C ABI (include/eradiate_b200.h) of the sm_100a path tracer:
// scene upload, parameter updates, render launches and the KAT entry points.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared
//        -Xcompiler -fPIC  (see __graft_entry__.build()).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include "ertb_kernel.cuh"
#include "ertb_kernel_pool.cuh"
#include "ertb_canopy.cuh"

// Host-side NVTX ranges (the B200 counterpart of the reference's MI_PROFILER_NVTX scopes):
// table commit, kernel launch and film read-back show up as named ranges in an nsys / ncu timeline.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ----------------------------------------------------------------------------
// error handling
// ----------------------------------------------------------------------------
static thread_local std::string g_error;

static int set_error(const std::string &msg) {
    g_error = msg;
    return 1;
}
#define CUDA_TRY(expr)                                                                    \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e)); \
    } while (0)

// ----------------------------------------------------------------------------
// scene object
// ----------------------------------------------------------------------------
struct HostPhase {
    int type = 0;
    float params[4] = { 0, 0, 0, 0 };
    std::vector<float> values, nodes;
    std::vector<float> mueller[5]; // tabulated_polarized: m12, m22, m33, m34, m44
};

struct HostSensor {
    ertb_sensor_desc desc;
    std::vector<double> directions; // normalised
    std::vector<double> origins;    // mradiancemeter
    double *d_origins = nullptr;    // device copy (3D kernel)
    float *d_table = nullptr;       // device primary-ray table
    double ray_offset = 0.0;
    int use_table = 0;
};

// Where one committed set of device tables lives: the scene's own slot (synchronous uploads on
// the default stream) or one slot of the batch pipeline (pinned staging buffer, private tables,
// private stream), so that the tables of context i+1 can be built and uploaded while context i
// is still rendering.
struct TableSlot {
    float *h_pinned = nullptr;
    size_t h_capacity = 0;
    float *d_blob = nullptr;
    size_t d_blob_capacity = 0;
    float *d_ocean = nullptr; // ocean_legacy transmittance tables and the parameters they were built for
    double ocean_key[3] = { -1.0, -1.0, -1.0 };
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool busy = false;
    bool async = false;
};

#define ERTB_BATCH_SLOTS 4
#define ERTB_BATCH_GRID_DIV 3
#define ERTB_MIN_WARP_PATHS 32ULL

struct BatchState {
    bool open = false;
    bool with_stats = false;
    int next = 0;
    std::vector<int> sensors;
    std::vector<size_t> offsets; // per item, in doubles, into d_accum
    size_t total = 0;
    double *d_accum = nullptr;
    size_t d_accum_capacity = 0;
    unsigned long long *d_counters = nullptr; // 16 per item: [0] work counter, [8..15] statistics
    size_t d_counters_capacity = 0;
    double t_begin = 0.0;
};

// ----------------------------------------------------------------------------
// canopy: host copy + BVH construction
// ----------------------------------------------------------------------------
struct HostLeafGroup {
    std::vector<float> disks;       // n x 7 leaves
    std::vector<float> trunk_disks; // trunk caps (diffuse)
    std::vector<float> cylinders;   // n x 7: p0, p1, radius (diffuse)
    float reflectance = 0.f, transmittance = 0.f, trunk_reflectance = 0.f;
    std::vector<float> triangles;   // n x 18: v0, v1, v2, shading normals n0, n1, n2 (mesh elements)
    std::vector<int> triangle_bsdf; // n: index into mesh_bsdfs
    std::vector<float> mesh_bsdfs;  // m x 2: bilambertian reflectance, transmittance
};

struct BvhBox {
    float lo[3], hi[3];
    void grow(const BvhBox &o) {
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], o.lo[k]); hi[k] = fmaxf(hi[k], o.hi[k]); }
    }
    static BvhBox empty() {
        BvhBox b;
        for (int k = 0; k < 3; ++k) { b.lo[k] = INFINITY; b.hi[k] = -INFINITY; }
        return b;
    }
};

// Binary BVH over `boxes` (binned SAH splits, median split for small nodes; leaves hold <= ERTB_BVH_LEAF
// primitives), then folded into
// the device layout where every node carries the boxes of its two children. `order` receives the
// primitive permutation the leaves index into (offset by `prim_base`); returns the root's index in
// `out` (several trees share one array).
struct BinNode { BvhBox box; int first, count, left; };

static int build_bvh(const std::vector<BvhBox> &boxes, std::vector<ErtbBvhNode> &out, std::vector<int> &order, int prim_base) {
    const int n = (int) boxes.size();
    order.resize(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::vector<BinNode> nodes;
    struct Item { int node, first, count; };
    std::vector<Item> stack;
    nodes.push_back(BinNode());
    stack.push_back({ 0, 0, n });
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        BvhBox bb = BvhBox::empty(), cb = BvhBox::empty();
        for (int i = it.first; i < it.first + it.count; ++i) {
            const BvhBox &b = boxes[order[i]];
            bb.grow(b);
            BvhBox c;
            for (int k = 0; k < 3; ++k) c.lo[k] = c.hi[k] = 0.5f * (b.lo[k] + b.hi[k]);
            cb.grow(c);
        }
        for (int k = 0; k < 3; ++k) { // pad by an ulp-scale margin: traversal compares in float
            float pad = 1e-6f * fmaxf(1.f, fmaxf(fabsf(bb.lo[k]), fabsf(bb.hi[k])));
            bb.lo[k] -= pad; bb.hi[k] += pad;
        }
        BinNode nd;
        nd.box = bb; nd.first = it.first; nd.count = it.count; nd.left = -1;
        int axis = 0;
        float ext = cb.hi[0] - cb.lo[0];
        for (int k = 1; k < 3; ++k) if (cb.hi[k] - cb.lo[k] > ext) { ext = cb.hi[k] - cb.lo[k]; axis = k; }
        if (it.count > ERTB_BVH_LEAF && ext > 0.f) {
            int mid = it.first + it.count / 2;
            bool split_done = false;
            if (it.count > 8) {
                // binned surface-area heuristic: 16 bins per axis over the centroid bounds; the split that
                // minimises area(left) * n_left + area(right) * n_right wins (fewer boxes pierced per ray)
                const int NBIN = 16;
                auto area = [](const BvhBox &b) {
                    float dx = fmaxf(b.hi[0] - b.lo[0], 0.f), dy = fmaxf(b.hi[1] - b.lo[1], 0.f), dz = fmaxf(b.hi[2] - b.lo[2], 0.f);
                    return dx * dy + dy * dz + dz * dx;
                };
                double best_cost = 1e300;
                int best_axis = -1, best_bin = -1;
                for (int ax = 0; ax < 3; ++ax) {
                    const float lo = cb.lo[ax], w = cb.hi[ax] - cb.lo[ax];
                    if (!(w > 0.f)) continue;
                    BvhBox bb_bin[NBIN];
                    int cnt[NBIN];
                    for (int k = 0; k < NBIN; ++k) { bb_bin[k] = BvhBox::empty(); cnt[k] = 0; }
                    for (int i = it.first; i < it.first + it.count; ++i) {
                        const BvhBox &b = boxes[order[i]];
                        int k = (int) (NBIN * (0.5f * (b.lo[ax] + b.hi[ax]) - lo) / w);
                        k = k < 0 ? 0 : (k >= NBIN ? NBIN - 1 : k);
                        bb_bin[k].grow(b);
                        cnt[k]++;
                    }
                    float right_area[NBIN];
                    int right_cnt[NBIN];
                    BvhBox acc = BvhBox::empty();
                    int c = 0;
                    for (int k = NBIN - 1; k > 0; --k) { acc.grow(bb_bin[k]); c += cnt[k]; right_area[k] = c ? area(acc) : 0.f; right_cnt[k] = c; }
                    acc = BvhBox::empty();
                    c = 0;
                    for (int k = 0; k < NBIN - 1; ++k) { // split between bin k and k + 1
                        acc.grow(bb_bin[k]); c += cnt[k];
                        if (c == 0 || right_cnt[k + 1] == 0) continue;
                        double cost = (double) area(acc) * c + (double) right_area[k + 1] * right_cnt[k + 1];
                        if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = k; }
                    }
                }
                if (best_axis >= 0) {
                    const float lo = cb.lo[best_axis], w = cb.hi[best_axis] - cb.lo[best_axis];
                    auto it_mid = std::partition(order.begin() + it.first, order.begin() + it.first + it.count, [&](int a) {
                        int k = (int) (16 * (0.5f * (boxes[a].lo[best_axis] + boxes[a].hi[best_axis]) - lo) / w);
                        k = k < 0 ? 0 : (k >= 16 ? 15 : k);
                        return k <= best_bin;
                    });
                    mid = (int) (it_mid - order.begin());
                    split_done = mid > it.first && mid < it.first + it.count;
                    if (!split_done) mid = it.first + it.count / 2;
                }
            }
            if (!split_done)
            std::nth_element(order.begin() + it.first, order.begin() + mid, order.begin() + it.first + it.count,
                             [&](int a, int b) { return boxes[a].lo[axis] + boxes[a].hi[axis] < boxes[b].lo[axis] + boxes[b].hi[axis]; });
            nd.left = (int) nodes.size();
            nd.count = 0;
            nodes.push_back(BinNode());
            nodes.push_back(BinNode());
            stack.push_back({ nd.left, it.first, mid - it.first });
            stack.push_back({ nd.left + 1, mid, it.first + it.count - mid });
        }
        nodes[it.node] = nd;
    }
    // fold: one wide node per inner binary node (a single-leaf tree gets an inner root with one empty child)
    std::vector<int> wide_of(nodes.size(), -1);
    const int root = (int) out.size();
    auto child = [&](const BinNode &c, float *lo, float *hi, int &ref, int &cnt) {
        for (int k = 0; k < 3; ++k) { lo[k] = c.box.lo[k]; hi[k] = c.box.hi[k]; }
        if (c.count > 0) { ref = prim_base + c.first; cnt = c.count; }
        else { ref = wide_of[&c - nodes.data()]; cnt = 0; }
    };
    for (size_t i = 0; i < nodes.size(); ++i)
        if (nodes[i].left >= 0) { wide_of[i] = (int) out.size(); out.push_back(ErtbBvhNode()); }
    if (nodes[0].left < 0) {
        ErtbBvhNode w;
        child(nodes[0], w.lo0, w.hi0, w.c0, w.n0);
        for (int k = 0; k < 3; ++k) w.lo1[k] = w.hi1[k] = 3e38f; // a point no segment reaches
        w.c1 = 0; w.n1 = 0;
        out.push_back(w);
        return root;
    }
    for (size_t i = 0; i < nodes.size(); ++i)
        if (nodes[i].left >= 0) {
            ErtbBvhNode &w = out[wide_of[i]];
            child(nodes[nodes[i].left], w.lo0, w.hi0, w.c0, w.n0);
            child(nodes[nodes[i].left + 1], w.lo1, w.hi1, w.c1, w.n1);
        }
    return root;
}

// largest dynamic shared-memory size requested so far, per (device, kernel function): see ERTB_OCC
static std::map<std::pair<int, const void *>, size_t> g_smem_attr;
static std::mutex g_smem_attr_mutex;

struct ertb_scene {
    int device = 0;
    int geometry = 0;
    double surface_z = 0, medium_bottom = 0, medium_top = 0;
    double bs_center[3] = { 0, 0, 0 };
    double bs_radius = 0;
    int has_medium = 0, n_layers = 0, homogeneous = 0;
    float scale = 1.f;
    std::vector<float> sigma_t, albedo, phase_weight;
    std::vector<int> band_starts; // banded majorant: first layer of every band chosen at the last commit
    cudaStream_t last_stream = nullptr; // stream of the last render through the main slot
    bool last_stream_valid = false;
    int n_phase = 0;
    HostPhase phase[ERTB_MAX_PHASE];
    int bsdf_type = 0;
    float bsdf_params[ERTB_MAX_BSDF_PARAMS];
    int has_patch = 0, patch_bsdf_type = 0; // CentralPatchSurface: second ground BSDF inside a rectangle
    float patch_bsdf_params[ERTB_MAX_BSDF_PARAMS];
    double patch_rect[4] = { 0, 0, 0, 0 };
    double emitter_dir[3];
    float irradiance = 1.f;
    int integrator = 0, rr_depth = 5;
    long long max_depth = -1;
    int polarized = 0, meridian_align = 0, phase_mis = 0;
    std::vector<HostSensor> sensors;

    // derived / device state
    bool dirty = true;
    ErtbParams base;            // everything but the per-launch fields
    std::vector<float> blob;    // host copy of the table blob
    TableSlot main;             // tables of the synchronous path
    TableSlot slots[ERTB_BATCH_SLOTS];
    BatchState batch;
    unsigned long long *d_counter = nullptr; // work counter + stats (1 + 8)
    double *d_accum = nullptr;
    size_t d_accum_capacity = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;
    std::map<std::pair<const void *, size_t>, int> occupancy; // per kernel instantiation and smem size
    int max_smem_optin = 0;
    double *d_gl = nullptr;     // Gauss-Legendre nodes/weights of the ocean transmittance quadrature
    double astro_diameter = 0.0;   // astroobject: angular diameter in degrees (0 = directional emitter)
    int hide_emitters = 0;         // integrator.cpp:29: primary rays do not see the disc
    float *d_bsdf_table = nullptr; // mqdiffuse: the measured table, [z][y][x] (static)
    int bsdf_table_res[3] = { 0, 0, 0 };
    // canopy (plane-parallel scenes): host description and the device BVH
    std::vector<HostLeafGroup> leaf_groups;
    std::vector<int> instance_group;
    std::vector<double> instance_offset;
    ErtbCanopy canopy;          // device pointers (zero-initialised: no canopy)
    void *d_canopy[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    std::vector<int> mesh_bsdf_base; // per group: first row of its mesh BSDFs in the table blob
    bool needs_3d = false;      // canopy or perspective sensor: rendered by ertb_canopy_kernel
    bool has_mesh = false;      // some group holds triangles: the MESH instances of that kernel
    bool small_groups = false;  // every group holds <= ERTB_VOTE_MAX_PRIMS primitives: the VOTE form of its BVH stage
};

static int build_canopy(ertb_scene *S) {
    memset(&S->canopy, 0, sizeof S->canopy);
    const int ninst = (int) S->instance_group.size();
    if (ninst == 0) return 0;
    const int ng = (int) S->leaf_groups.size();
    // world bounding box of every instance
    std::vector<BvhBox> gbox(ng, BvhBox::empty());
    std::vector<std::vector<BvhBox>> dboxes(ng);
    // primitives of a group, in this order: leaf disks (kind 0), trunk cap disks (kind 1), cylinders (kind 2)
    std::vector<std::vector<float>> prims(ng); // 8 floats each: device record, kind in the last slot
    std::vector<float> tris; // 16 floats per triangle (ErtbCanopy::tris), in input order; the record points here
    S->mesh_bsdf_base.assign(ng, 0);
    int n_mesh_bsdfs = 0;
    for (int g = 0; g < ng; ++g) {
        const HostLeafGroup &hg = S->leaf_groups[g];
        S->mesh_bsdf_base[g] = n_mesh_bsdfs;
        n_mesh_bsdfs += (int) hg.mesh_bsdfs.size() / 2;
        for (size_t i = 0; i < hg.triangles.size() / 18; ++i) {
            const float *q = &hg.triangles[18 * i];
            BvhBox b;
            for (int k = 0; k < 3; ++k) {
                b.lo[k] = fminf(q[k], fminf(q[3 + k], q[6 + k]));
                b.hi[k] = fmaxf(q[k], fmaxf(q[3 + k], q[6 + k]));
                const float pad = 1e-6f * fmaxf(1.f, fmaxf(fabsf(b.lo[k]), fabsf(b.hi[k]))); // axis-aligned faces: no empty box
                b.lo[k] -= pad; b.hi[k] += pad;
            }
            dboxes[g].push_back(b);
            gbox[g].grow(b);
            const int kind = 3, tri = (int) tris.size() / 16, mat = S->mesh_bsdf_base[g] + hg.triangle_bsdf[i];
            // record: (v0, triangle index) (e1 = v1 - v0, kind); tris: (e2 = v2 - v0, mesh BSDF row) n0 n1 n2
            float row[8] = { q[0], q[1], q[2], 0.f, q[3] - q[0], q[4] - q[1], q[5] - q[2], 0.f };
            memcpy(&row[3], &tri, sizeof(float));
            memcpy(&row[7], &kind, sizeof(float));
            prims[g].insert(prims[g].end(), row, row + 8);
            float ext[16] = { q[6] - q[0], q[7] - q[1], q[8] - q[2], 0.f, q[9], q[10], q[11], 0.f,
                              q[12], q[13], q[14], 0.f, q[15], q[16], q[17], 0.f };
            memcpy(&ext[3], &mat, sizeof(float));
            tris.insert(tris.end(), ext, ext + 16);
        }
        for (int pass = 0; pass < 2; ++pass) {
            const std::vector<float> &dk = pass == 0 ? hg.disks : hg.trunk_disks;
            for (size_t i = 0; i < dk.size() / 7; ++i) {
                const float *q = &dk[7 * i];
                BvhBox b;
                for (int k = 0; k < 3; ++k) {
                    float e = q[6] * sqrtf(fmaxf(1.f - q[3 + k] * q[3 + k], 0.f));
                    b.lo[k] = q[k] - e; b.hi[k] = q[k] + e;
                }
                dboxes[g].push_back(b);
                gbox[g].grow(b);
                int kind = pass;
                float row[8] = { q[0], q[1], q[2], q[6], q[3], q[4], q[5], 0.f };
                memcpy(&row[7], &kind, sizeof(float));
                prims[g].insert(prims[g].end(), row, row + 8);
            }
        }
        for (size_t i = 0; i < hg.cylinders.size() / 7; ++i) {
            const float *q = &hg.cylinders[7 * i];
            const float ax[3] = { q[3] - q[0], q[4] - q[1], q[5] - q[2] };
            const float L = sqrtf(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
            BvhBox b;
            for (int k = 0; k < 3; ++k) {
                float u = ax[k] / L, e = q[6] * sqrtf(fmaxf(1.f - u * u, 0.f));
                b.lo[k] = fminf(q[k], q[3 + k]) - e; b.hi[k] = fmaxf(q[k], q[3 + k]) + e;
            }
            dboxes[g].push_back(b);
            gbox[g].grow(b);
            int kind = 2;
            float row[8] = { q[0], q[1], q[2], q[6], ax[0], ax[1], ax[2], 0.f }; // p0, radius | axis vector p1 - p0
            memcpy(&row[7], &kind, sizeof(float));
            prims[g].insert(prims[g].end(), row, row + 8);
        }
    }
    double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int i = 0; i < ninst; ++i)
        for (int k = 0; k < 3; ++k) {
            lo[k] = fmin(lo[k], S->instance_offset[3 * i + k] + (double) gbox[S->instance_group[i]].lo[k]);
            hi[k] = fmax(hi[k], S->instance_offset[3 * i + k] + (double) gbox[S->instance_group[i]].hi[k]);
        }
    ErtbCanopy &C = S->canopy;
    for (int k = 0; k < 3; ++k) {
        double pad = 1e-4 * fmax(1.0, hi[k] - lo[k]);
        C.lo[k] = lo[k] - pad; C.hi[k] = hi[k] + pad;
        C.origin[k] = 0.5 * (lo[k] + hi[k]);
    }
    C.lo[2] = fmax(C.lo[2], S->surface_z); // leaves may dip below the ground plane; rays never do
    // bottom level: one tree per group, disks reordered into leaf order
    std::vector<ErtbBvhNode> blas;
    std::vector<int> blas_root(ng);
    S->small_groups = ng > 0;
    std::vector<float> disks;
    for (int g = 0; g < ng; ++g) {
        std::vector<int> order;
        if (dboxes[g].size() > (size_t) ERTB_VOTE_MAX_PRIMS) S->small_groups = false;
        blas_root[g] = build_bvh(dboxes[g], blas, order, (int) disks.size() / 8);
        for (int i : order) disks.insert(disks.end(), &prims[g][8 * (size_t) i], &prims[g][8 * (size_t) i] + 8);
    }
    if (const char *e = getenv("ERTB_CANOPY_VOTE")) S->small_groups = atoi(e) != 0; // developer knob (A/B runs)
    // top level over the instances (float coordinates relative to the canopy origin)
    std::vector<BvhBox> iboxes(ninst);
    for (int i = 0; i < ninst; ++i)
        for (int k = 0; k < 3; ++k) {
            float off = (float) (S->instance_offset[3 * i + k] - C.origin[k]);
            iboxes[i].lo[k] = off + gbox[S->instance_group[i]].lo[k];
            iboxes[i].hi[k] = off + gbox[S->instance_group[i]].hi[k];
        }
    std::vector<ErtbBvhNode> tlas;
    std::vector<int> iorder;
    build_bvh(iboxes, tlas, iorder, 0);
    std::vector<float> inst(4 * (size_t) ninst);
    for (int j = 0; j < ninst; ++j) {
        const int i = iorder[j];
        for (int k = 0; k < 3; ++k) inst[4 * j + k] = (float) (S->instance_offset[3 * i + k] - C.origin[k]);
        int g = S->instance_group[i];
        memcpy(&inst[4 * j + 3], &g, sizeof(float));
    }
    const void *src[6] = { tlas.data(), inst.data(), blas.data(), blas_root.data(), disks.data(), tris.data() };
    const size_t bytes[6] = { tlas.size() * sizeof(ErtbBvhNode), inst.size() * sizeof(float),
                              blas.size() * sizeof(ErtbBvhNode), blas_root.size() * sizeof(int),
                              disks.size() * sizeof(float), tris.size() * sizeof(float) };
    for (int k = 0; k < 6; ++k) {
        if (bytes[k] == 0) continue;
        CUDA_TRY(cudaMalloc(&S->d_canopy[k], bytes[k]));
        CUDA_TRY(cudaMemcpy(S->d_canopy[k], src[k], bytes[k], cudaMemcpyHostToDevice));
    }
    C.n_instances = ninst;
    C.tlas = (const ErtbBvhNode *) S->d_canopy[0];
    C.inst = (const float4 *) S->d_canopy[1];
    C.blas = (const ErtbBvhNode *) S->d_canopy[2];
    C.blas_root = (const int *) S->d_canopy[3];
    C.disks = (const float4 *) S->d_canopy[4];
    C.tris = (const float4 *) S->d_canopy[5];
    return 0;
}

static size_t align4(size_t n) { return (n + 3) & ~size_t(3); }
static double wall_ms();

// Banded majorant (ertb_kernel_pool.cuh): cut the layer stack into <= ERTB_MAX_BANDS contiguous bands
// minimising the expected cost of a vertical traverse in units of one loop trip,
//     sum over bands (band majorant x band thickness)  +  ERTB_BAND_PENALTY x (number of bands),
// i.e. tentative collisions plus one boundary crossing per band (a crossing is handled inside the trip
// that reaches it: one boundary distance, a fraction of a trip). O(n^2 x bands) dynamic programme.
// Bands are used when the predicted cost of a traverse, 1 (the trip that ends the segment) + the sum
// above, is below ERTB_BAND_MIN_GAIN x the same figure for the single global majorant.
#define ERTB_MAX_BANDS 8
#ifndef ERTB_BAND_PENALTY
#define ERTB_BAND_PENALTY 0.1
#endif
#ifndef ERTB_BAND_MIN_GAIN
#define ERTB_BAND_MIN_GAIN 0.5
#endif
// (developer knobs for A/B runs, read once: ERTB_BAND_PENALTY / ERTB_BAND_MIN_GAIN in the environment)
static double env_double(const char *name, double dflt, double lo, double hi) {
    const char *e = getenv(name);
    if (!e) return dflt;
    char *end = nullptr;
    double v = strtod(e, &end);
    return (end == e || !(v >= lo) || !(v <= hi)) ? dflt : v;
}
// developer tuning knobs (integers), clamped: a zero threshold would stall the persistent scheduler loop
static int env_int_clamped(const char *name, int dflt, int lo, int hi) {
    const char *e = getenv(name);
    if (!e) return dflt;
    char *end = nullptr;
    long v = strtol(e, &end, 10);
    if (end == e) return dflt;
    return (int) (v < lo ? lo : (v > hi ? hi : v));
}
static double band_penalty() { static const double v = env_double("ERTB_BAND_PENALTY", ERTB_BAND_PENALTY, 0.0, 100.0); return v; }
static double band_min_gain() { static const double v = env_double("ERTB_BAND_MIN_GAIN", ERTB_BAND_MIN_GAIN, 0.0, 10.0); return v; }

static std::vector<int> choose_bands(const std::vector<double> &sigma, double dz) {
    const int n = (int) sigma.size();
    // candidate cut points: the layer edges with the largest jumps of log(sigma_t) plus a uniform
    // comb, <= 96 in all, so that the programme costs microseconds (it runs at every parameter update)
    std::vector<int> cand;
    if (n <= 96) {
        for (int i = 0; i <= n; ++i) cand.push_back(i);
    } else {
        std::vector<std::pair<double, int>> jump;
        for (int i = 1; i < n; ++i) {
            double a = fmax(sigma[i - 1], 1e-300), b = fmax(sigma[i], 1e-300);
            jump.push_back({ -fabs(log(a / b)), i });
        }
        std::partial_sort(jump.begin(), jump.begin() + 48, jump.end());
        std::vector<char> mark(n + 1, 0);
        mark[0] = mark[n] = 1;
        for (int k = 0; k < 48; ++k) mark[jump[k].second] = 1;
        for (int k = 1; k < 48; ++k) mark[(int) ((long long) k * n / 48)] = 1;
        for (int i = 0; i <= n; ++i) if (mark[i]) cand.push_back(i);
    }
    const int m = (int) cand.size(); // cand[0] = 0, cand[m - 1] = n
    std::vector<double> segmax(m - 1, 0.0);
    for (int c = 0; c + 1 < m; ++c)
        for (int i = cand[c]; i < cand[c + 1]; ++i) segmax[c] = fmax(segmax[c], sigma[i]);
    const double INF = 1e300;
    std::vector<std::vector<double>> cost(ERTB_MAX_BANDS + 1, std::vector<double>(m, INF));
    std::vector<std::vector<int>> from(ERTB_MAX_BANDS + 1, std::vector<int>(m, -1));
    cost[0][0] = 0.0;
    for (int k = 1; k <= ERTB_MAX_BANDS; ++k)
        for (int j = 1; j < m; ++j) {
            double mx = 0.0;
            for (int i = j - 1; i >= 0; --i) { // band = layers [cand[i], cand[j])
                mx = fmax(mx, segmax[i]);
                if (cost[k - 1][i] >= INF) continue;
                double c = cost[k - 1][i] + mx * dz * (cand[j] - cand[i]) + band_penalty();
                if (c < cost[k][j]) { cost[k][j] = c; from[k][j] = i; }
            }
        }
    int best = 1;
    for (int k = 2; k <= ERTB_MAX_BANDS; ++k) if (cost[k][m - 1] < cost[best][m - 1] - 1e-12) best = k;
    if (1.0 + cost[best][m - 1] > band_min_gain() * (1.0 + cost[1][m - 1])) best = 1;
    std::vector<int> starts(best);
    for (int k = best, j = m - 1; k >= 1; --k) { j = from[k][j]; starts[k - 1] = cand[j]; }
    return starts; // first layer of every band (starts[0] = 0)
}

// distr_1d.h:548-600 compute_cdf_scalar: trapezoid CDF accumulated in double
static void build_tab_leaf(const HostPhase &hp, ErtbPhaseLeaf &L, std::vector<float> &blob) {
    const int n = (int) hp.values.size();
    const bool irregular = hp.type != ERTB_PHASE_TABULATED;
    L.n_nodes = n;
    L.inv_interval = (float) ((n - 1) / 2.0);
    L.off_pdf = (int) blob.size();
    blob.insert(blob.end(), hp.values.begin(), hp.values.end());
    blob.resize(align4(blob.size()), 0.f);
    L.off_cdf = (int) blob.size();
    double integral = 0.0, interval = 2.0 / (n - 1);
    int v0 = -1, v1 = -1;
    std::vector<float> cdf(n - 1);
    for (int i = 0; i < n - 1; ++i) {
        double w = irregular ? ((double) hp.nodes[i + 1] - (double) hp.nodes[i]) : interval;
        double value = 0.5 * w * ((double) hp.values[i] + (double) hp.values[i + 1]);
        integral += value;
        cdf[i] = (float) integral;
        if (value > 0.0) {
            if (v0 < 0) v0 = i;
            v1 = i;
        }
    }
    blob.insert(blob.end(), cdf.begin(), cdf.end());
    blob.resize(align4(blob.size()), 0.f);
    L.valid0 = v0;
    L.valid1 = v1;
    L.integral = cdf[v1];
    L.normalization = 1.f / L.integral;
    L.off_nodes = -1;
    if (irregular) {
        L.off_nodes = (int) blob.size();
        blob.insert(blob.end(), hp.nodes.begin(), hp.nodes.end());
        blob.resize(align4(blob.size()), 0.f);
    }
    L.off_mueller = -1;
    L.mueller_stride = 0;
    if (hp.type == ERTB_PHASE_TABULATED_POLARIZED) {
        L.off_mueller = (int) blob.size();
        L.mueller_stride = (int) align4((size_t) n);
        for (int k = 0; k < 5; ++k) {
            size_t start = blob.size();
            if ((int) hp.mueller[k].size() == n) blob.insert(blob.end(), hp.mueller[k].begin(), hp.mueller[k].end());
            blob.resize(start + L.mueller_stride, 0.f);
        }
    }
}

static int validate_tab(const HostPhase &hp) {
    const int n = (int) hp.values.size();
    if (n < 2) return set_error("ContinuousDistribution: needs at least two entries!");
    if (n > ERTB_MAX_PHASE_NODES) return set_error("tabulated phase function: too many nodes");
    bool mass = false;
    for (int i = 0; i < n; ++i) {
        if (!(hp.values[i] >= 0.f)) return set_error("ContinuousDistribution: entries must be non-negative!");
        if (hp.values[i] > 0.f) mass = true;
    }
    if (!mass) return set_error("ContinuousDistribution: no probability mass found!");
    if (hp.type != ERTB_PHASE_TABULATED) {
        if ((int) hp.nodes.size() != n) return set_error("'nodes' and 'values' must have the same length");
        for (int i = 0; i < n - 1; ++i)
            if (!(hp.nodes[i + 1] > hp.nodes[i]))
                return set_error("IrregularContinuousDistribution: node positions must be strictly increasing!");
        if (hp.nodes[0] != -1.f || hp.nodes[n - 1] != 1.f) return set_error("'nodes' bounds must be [-1, 1]");
    }
    return 0;
}

// Rebuild derived tables + base kernel parameters and upload them.
static int scene_commit(ertb_scene *S, TableSlot &T) {
    NvtxRange nvtx("ertb:scene_commit");
    CUDA_TRY(cudaSetDevice(S->device));
    ErtbParams &P = S->base;
    memset(&P, 0, sizeof P);
    const bool sph = S->geometry == ERTB_GEOM_SPHERICAL_SHELL;
    P.spherical = sph;
    P.R = (float) S->surface_z;
    P.Rd = S->surface_z;
    P.has_medium = S->has_medium;
    P.H = S->has_medium ? (float) (S->medium_top - S->surface_z) : 0.f;
    P.n_layers = S->has_medium ? S->n_layers : 1;
    P.h_off = (float) (S->surface_z - S->medium_bottom);
    P.inv_dz = S->has_medium ? (float) (S->n_layers / (S->medium_top - S->medium_bottom)) : 0.f;
    P.n_phase = S->n_phase;

    std::vector<float> &blob = S->blob;
    blob.clear();
    double majorant = 0.0;
    if (S->has_medium) {
        const int n = S->n_layers;
        float mx = S->sigma_t[0];
        for (int i = 1; i < n; ++i) mx = fmaxf(mx, S->sigma_t[i]);
        majorant = (double) S->scale * (double) mx; // heterogeneous.cpp:163 / homogeneous.cpp:157
        P.off_preal = (int) blob.size();
        for (int i = 0; i < n; ++i) {
            double st = (double) S->scale * (double) S->sigma_t[i];
            double pr = majorant > 0.0 ? st / majorant : 0.0;
            if (S->homogeneous) pr = majorant > 0.0 ? 1.0 : 0.0;
            blob.push_back((float) fmin(fmax(pr, 0.0), 1.0));
        }
        blob.resize(align4(blob.size()), 0.f);
        P.off_albedo = (int) blob.size();
        blob.insert(blob.end(), S->albedo.begin(), S->albedo.end());
        blob.resize(align4(blob.size()), 0.f);
        P.off_cumw = (int) blob.size();
        if (S->n_phase > 1) {
            std::vector<double> cum(n, 0.0);
            for (int k = 0; k < S->n_phase - 1; ++k)
                for (int i = 0; i < n; ++i) {
                    cum[i] += (double) S->phase_weight[(size_t) k * n + i];
                    blob.push_back((float) cum[i]);
                }
            blob.resize(align4(blob.size()), 0.f);
        }
        // banded majorant (pool kernel only; ERTB_MAJORANT=global keeps the reference's single majorant)
        P.n_bands = 1;
        {
            const char *e = getenv("ERTB_MAJORANT");
            const bool want = !(e && strcmp(e, "global") == 0) && !S->homogeneous && !S->needs_3d && n > 1 &&
                              S->integrator != ERTB_INTEGRATOR_PIECEWISE_VOLPATH && majorant > 0.0;
            if (want) {
                std::vector<double> sg(n);
                for (int i = 0; i < n; ++i) sg[i] = (double) S->scale * (double) S->sigma_t[i];
                const double dz = (S->medium_top - S->medium_bottom) / n;
                // The programme costs ~0.1 ms: spectral loops commit once per context, so (1) skip it when even
                // the ideal partition (every layer its own band, two boundary stops) cannot halve the trips,
                // (2) keep the previous cut while it still does (its majorants are re-derived below anyway).
                double tau = 0.0, mxg = 0.0;
                for (int i = 0; i < n; ++i) { tau += sg[i] * dz; mxg = fmax(mxg, sg[i]); }
                const double cost_global = 1.0 + mxg * dz * n + band_penalty();
                std::vector<int> starts(1, 0);
                if (1.0 + tau + 2.0 * band_penalty() < band_min_gain() * cost_global) {
                    bool reuse = false;
                    if (S->band_starts.size() > 1 && S->band_starts.back() < n) {
                        double c = 0.0;
                        for (size_t k = 0; k < S->band_starts.size(); ++k) {
                            const int a = S->band_starts[k], b = k + 1 < S->band_starts.size() ? S->band_starts[k + 1] : n;
                            double mx = 0.0;
                            for (int i = a; i < b; ++i) mx = fmax(mx, sg[i]);
                            c += mx * dz * (b - a) + band_penalty();
                        }
                        reuse = 1.0 + c < band_min_gain() * cost_global;
                    }
                    starts = reuse ? S->band_starts : choose_bands(sg, dz);
                }
                S->band_starts = starts;
                const int nb = (int) starts.size();
                if (nb > 1) {
                    P.n_bands = nb;
                    P.off_band_lo = (int) blob.size();
                    for (int k = 0; k < nb; ++k) // altitude above the ground of the band's lower boundary
                        blob.push_back((float) (S->medium_bottom + starts[k] * dz - S->surface_z));
                    blob.push_back((float) (S->medium_top - S->surface_z));
                    blob.resize(align4(blob.size()), 0.f);
                    P.off_band_ratio = (int) blob.size();
                    for (int k = 0; k < nb; ++k) {
                        double mx = 0.0;
                        for (int i = starts[k]; i < (k + 1 < nb ? starts[k + 1] : n); ++i) mx = fmax(mx, sg[i]);
                        // (a band of vacuum gets a token majorant: its flights overshoot the band at once)
                        blob.push_back((float) (majorant / fmax(mx, 1e-6 * majorant)));
                    }
                    blob.resize(align4(blob.size()), 0.f);
                    P.off_band_iratio = (int) blob.size();
                    for (int k = 0; k < nb; ++k) blob.push_back(1.f / blob[(size_t) P.off_band_ratio + k]);
                    blob.resize(align4(blob.size()), 0.f);
                }
            }
        }
        P.piecewise = S->integrator == ERTB_INTEGRATOR_PIECEWISE_VOLPATH;
        if (P.piecewise) {
            // piecewise.cpp:445-507 precompute_optical_thickness, folded into one table of the vertical
            // optical depth above each layer boundary (double accumulation, float storage)
            const double dz = (S->medium_top - S->medium_bottom) / n;
            P.dz = (float) dz;
            P.off_sigma = (int) blob.size();
            for (int i = 0; i < n; ++i) blob.push_back((float) ((double) S->scale * (double) S->sigma_t[i]));
            blob.resize(align4(blob.size()), 0.f);
            P.off_tau = (int) blob.size();
            std::vector<double> tt((size_t) n + 1, 0.0); // from boundary i up to the top
            for (int i = n - 1; i >= 0; --i) tt[i] = tt[i + 1] + (double) S->scale * (double) S->sigma_t[i] * dz;
            for (int i = 0; i <= n; ++i) blob.push_back((float) tt[i]);
            blob.resize(align4(blob.size()), 0.f);
            // optical depth above the ground level (the grid may start below the surface)
            double xg = fmin(fmax((S->surface_z - S->medium_bottom) / dz, 0.0), (double) n);
            int lg = (int) fmin(floor(xg), (double) (n - 1));
            P.tau_ground = (float) (tt[lg] - (double) S->scale * (double) S->sigma_t[lg] * (xg - lg) * dz);
        }
        for (int k = 0; k < S->n_phase; ++k) {
            ErtbPhaseLeaf &L = P.leaf[k];
            const HostPhase &hp = S->phase[k];
            L.type = hp.type;
            L.p0 = hp.params[0];
            L.off_nodes = -1;
            L.off_mueller = -1;
            if (hp.type == ERTB_PHASE_TABULATED || hp.type == ERTB_PHASE_TABULATED_IRREGULAR ||
                hp.type == ERTB_PHASE_TABULATED_POLARIZED) {
                if (validate_tab(hp)) return 1;
                build_tab_leaf(hp, L, blob);
            }
        }
    }
    P.majorant = (float) majorant;
    P.inv_majorant = majorant > 0.0 ? (float) (1.0 / majorant) : INFINITY;
    P.canopy = S->canopy;
    P.canopy.off_leaf_bsdf = (int) blob.size();
    for (const HostLeafGroup &g : S->leaf_groups) { // 4 floats per group: leaf r, leaf t, trunk rho, -
        blob.push_back(g.reflectance); blob.push_back(g.transmittance); blob.push_back(g.trunk_reflectance); blob.push_back(0.f);
    }
    blob.resize(align4(blob.size()), 0.f);
    P.canopy.off_mesh_bsdf = (int) blob.size(); // 2 floats per mesh BSDF, groups concatenated (mesh_bsdf_base)
    for (const HostLeafGroup &g : S->leaf_groups) blob.insert(blob.end(), g.mesh_bsdfs.begin(), g.mesh_bsdfs.end());
    blob.resize(align4(blob.size()), 0.f);
    P.canopy.patch_type = -1;
    if (S->has_patch) { // the patch BSDF travels with the tables (spectral updates, batch slots)
        P.canopy.patch_type = S->patch_bsdf_type;
        P.canopy.off_patch_bsdf = (int) blob.size();
        blob.insert(blob.end(), S->patch_bsdf_params, S->patch_bsdf_params + ERTB_MAX_BSDF_PARAMS);
        blob.resize(align4(blob.size()), 0.f);
        for (int k = 0; k < 4; ++k) P.canopy.patch_rect[k] = S->patch_rect[k];
    }

    const size_t blob_bytes = blob.size() * sizeof(float);
    if (blob_bytes > (size_t) S->max_smem_optin - 1024)
        return set_error("scene tables do not fit in one SM's shared memory");
    if (blob_bytes > T.d_blob_capacity) {
        if (T.d_blob) cudaFree(T.d_blob);
        T.d_blob = nullptr;
        T.d_blob_capacity = 0;
        CUDA_TRY(cudaMalloc(&T.d_blob, blob_bytes));
        T.d_blob_capacity = blob_bytes;
    }
    if (blob_bytes && T.async) {
        if (blob_bytes > T.h_capacity) {
            if (T.h_pinned) cudaFreeHost(T.h_pinned);
            T.h_pinned = nullptr;
            T.h_capacity = 0;
            CUDA_TRY(cudaMallocHost(&T.h_pinned, blob_bytes));
            T.h_capacity = blob_bytes;
        }
        memcpy(T.h_pinned, blob.data(), blob_bytes);
        CUDA_TRY(cudaMemcpyAsync(T.d_blob, T.h_pinned, blob_bytes, cudaMemcpyHostToDevice, T.stream));
    } else if (blob_bytes) {
        CUDA_TRY(cudaMemcpy(T.d_blob, blob.data(), blob_bytes, cudaMemcpyHostToDevice));
    }
    P.blob = T.d_blob;
    P.blob_bytes = (int) blob_bytes;

    P.bsdf_type = S->bsdf_type;
    memcpy(P.bsdf, S->bsdf_params, sizeof P.bsdf);
    P.ocean_tables = nullptr;
    if (S->bsdf_type == ERTB_BSDF_OCEAN_LEGACY) {
        // OceanBSDF::update() (ocean_legacy.cpp:313-372): scalars on the host, tables on the device
        double n_real, n_imag;
        ertb_ocean_host::derive(S->bsdf_params, P.bsdf, n_real, n_imag);
        const double ws = S->bsdf_params[1];
        if (!S->d_gl) {
            CUDA_TRY(cudaMalloc(&S->d_gl, 2 * ERTB_OC_RES * sizeof(double)));
            double gl[2 * ERTB_OC_RES];
            ertb_ocean_host::gauss_legendre(ERTB_OC_RES, gl, gl + ERTB_OC_RES);
            CUDA_TRY(cudaMemcpy(S->d_gl, gl, sizeof gl, cudaMemcpyHostToDevice));
        }
        if (!T.d_ocean) CUDA_TRY(cudaMalloc(&T.d_ocean, 2 * ERTB_OC_RES * ERTB_OC_RES * sizeof(float)));
        if (T.ocean_key[0] != n_real || T.ocean_key[1] != n_imag || T.ocean_key[2] != ws) {
            ertb_ocean_tables_kernel<<<(2 * ERTB_OC_RES * ERTB_OC_RES + 127) / 128, 128, 0, T.stream>>>(n_real, n_imag, ws, S->d_gl, T.d_ocean);
            CUDA_TRY(cudaGetLastError());
            // the main slot builds its tables on the legacy stream, which does not order against a caller's
            // non-blocking stream: finish them before any render can be queued behind this commit
            if (!T.async) CUDA_TRY(cudaStreamSynchronize(T.stream));
            T.ocean_key[0] = n_real; T.ocean_key[1] = n_imag; T.ocean_key[2] = ws;
        }
        P.ocean_tables = T.d_ocean;
    } else if (S->bsdf_type == ERTB_BSDF_MQDIFFUSE) {
        for (int i = 0; i < 3; ++i) P.bsdf[i] = (float) S->bsdf_table_res[i];
        P.ocean_tables = S->d_bsdf_table;
    } else if (S->bsdf_type == ERTB_BSDF_MEASURED_MONO) {
        P.ocean_tables = S->d_bsdf_table; // self-describing: sizes and offsets in its header
    } else if (S->bsdf_type >= ERTB_BSDF_OCEAN_MISHCHENKO) {
        ertb_ocean_host::derive_glint(S->bsdf_type, S->bsdf_params, P.bsdf);
    }
    double dn = sqrt(S->emitter_dir[0] * S->emitter_dir[0] + S->emitter_dir[1] * S->emitter_dir[1] +
                     S->emitter_dir[2] * S->emitter_dir[2]);
    for (int i = 0; i < 3; ++i) P.sun[i] = (float) (-S->emitter_dir[i] / dn);
    P.irradiance = S->irradiance;
    P.astro_omc = P.astro_sin2 = P.astro_radiance = 0.f;
    if (S->astro_diameter > 0.0) { // astroobject.cpp:75-80
        const double a = 0.5 * S->astro_diameter * M_PI / 180.0;
        const double omc = 2.0 * sin(0.5 * a) * sin(0.5 * a); // 1 - cos a without cancellation
        P.astro_omc = (float) omc;
        P.astro_sin2 = (float) (sin(a) * sin(a));
        // the direct view of the disc by unscattered primary rays (volpath.cpp:329-330 switches it off)
        P.astro_radiance = S->hide_emitters ? 0.f : (float) ((double) S->irradiance / (2.0 * M_PI * omc));
    }
    P.polarized = S->polarized;
    P.phase_mis = S->phase_mis;
    P.meridian_align = S->meridian_align;
    P.mis = S->integrator == ERTB_INTEGRATOR_VOLPATHMIS;
    P.rr_depth = (unsigned) S->rr_depth;
    P.max_depth = S->max_depth < 0 ? 0xffffffffu : (unsigned) S->max_depth;
    S->dirty = T.async; // a batch slot's tables are private to that batch item
    return 0;
}

static void slot_release(TableSlot &T) {
    if (T.h_pinned) cudaFreeHost(T.h_pinned);
    if (T.d_blob) cudaFree(T.d_blob);
    if (T.d_ocean) cudaFree(T.d_ocean);
    if (T.done) cudaEventDestroy(T.done);
    if (T.stream) cudaStreamDestroy(T.stream);
    T = TableSlot();
}

// mdistant.cpp:180-190: ray_offset default
static double sensor_ray_offset(const ertb_scene *S, const ertb_sensor_desc &sd) {
    if (sd.ray_offset >= 0.0) return sd.ray_offset;
    const double eps = 2.220446049250313e-16 * 0.5 * 1500.0; // math::RayEpsilon<double> (parity target)
    double rad = fmax(eps, S->bs_radius * (1.0 + eps));
    return sd.target_type == ERTB_TARGET_NONE ? rad : 2.0 * rad;
}

static int build_sensor(ertb_scene *S, HostSensor &hs) {
    const ertb_sensor_desc &sd = hs.desc;
    hs.ray_offset = sensor_ray_offset(S, sd);
    const int npix = sd.width * sd.height;
    const bool sph = S->geometry == ERTB_GEOM_SPHERICAL_SHELL;
    if (sd.type == ERTB_SENSOR_MRADIANCEMETER) {
        // mradiancemeter.cpp:147-172: explicit rays. Per pixel [n0 | d | class | h0]: class 1 = enters the
        // atmosphere at n0 (h0 = H), 2 = reaches the ground through vacuum at n0, 3 = starts INSIDE the
        // atmosphere at altitude h0 above n0's foot point, 0 = sees nothing
        if (sd.n_directions != sd.width || sd.height != 1) return set_error("Film size must be [n_radiancemeters, 1]");
        hs.use_table = 1;
        std::vector<float> table((size_t) npix * 8, 0.f);
        const double Rg = S->surface_z, Rt = S->has_medium ? S->medium_top : S->surface_z;
        for (int i = 0; i < npix; ++i) {
            const double *o = &hs.origins[3 * (size_t) i];
            const double dx = hs.directions[3 * i], dy = hs.directions[3 * i + 1], dz = hs.directions[3 * i + 2];
            float *t = &table[(size_t) i * 8];
            t[3] = (float) dx; t[4] = (float) dy; t[5] = (float) dz;
            t[2] = 1.f;
            if (sph) {
                const double oo = o[0] * o[0] + o[1] * o[1] + o[2] * o[2], r = sqrt(oo), b = o[0] * dx + o[1] * dy + o[2] * dz;
                if (r < Rg) continue;
                if (S->has_medium && r < Rt) {
                    if (!sd.in_medium) return set_error("mradiancemeter: an origin lies inside the atmosphere but the sensor has no medium");
                    t[0] = (float) (o[0] / r); t[1] = (float) (o[1] / r); t[2] = (float) (o[2] / r);
                    t[6] = 3.f; t[7] = (float) (r - Rg);
                    continue;
                }
                if (sd.in_medium) return set_error("mradiancemeter: the sensor has a medium but an origin lies outside the atmosphere");
                const double Rs = S->has_medium ? Rt : Rg;
                const double disc = b * b - (oo - Rs * Rs);
                if (disc < 0.0 || b > 0.0) continue;
                const double t0 = -b - sqrt(disc);
                t[0] = (float) ((o[0] + t0 * dx) / Rs); t[1] = (float) ((o[1] + t0 * dy) / Rs); t[2] = (float) ((o[2] + t0 * dz) / Rs);
                t[6] = S->has_medium ? 1.f : 2.f; t[7] = S->has_medium ? (float) (Rt - Rg) : 0.f;
            } else {
                const double h = o[2] - Rg, H = Rt - Rg;
                if (h < 0.0) continue;
                if (S->has_medium && h < H) {
                    if (!sd.in_medium) return set_error("mradiancemeter: an origin lies inside the atmosphere but the sensor has no medium");
                    t[6] = 3.f; t[7] = (float) h;
                } else {
                    if (sd.in_medium) return set_error("mradiancemeter: the sensor has a medium but an origin lies outside the atmosphere");
                    if (dz < 0.0) { t[6] = S->has_medium ? 1.f : 2.f; t[7] = (float) H; }
                }
            }
        }
        CUDA_TRY(cudaMalloc(&hs.d_table, table.size() * sizeof(float)));
        CUDA_TRY(cudaMemcpy(hs.d_table, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc(&hs.d_origins, hs.origins.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(hs.d_origins, hs.origins.data(), hs.origins.size() * sizeof(double), cudaMemcpyHostToDevice));
        return 0;
    }
    if (sd.type == ERTB_SENSOR_MDISTANT) {
        if (sd.n_directions != sd.width || sd.height != 1)
            return set_error("Film size must be [sensor_count, 1]");
        hs.use_table = (!sph) || sd.target_type == ERTB_TARGET_POINT;
        std::vector<float> table((size_t) npix * 8, 0.f);
        const double Rt = S->has_medium ? S->medium_top : S->surface_z;
        for (int i = 0; i < npix; ++i) {
            double dx = hs.directions[3 * i], dy = hs.directions[3 * i + 1], dz = hs.directions[3 * i + 2];
            float *t = &table[(size_t) i * 8];
            t[3] = (float) dx; t[4] = (float) dy; t[5] = (float) dz;
            t[6] = 1.f;
            t[2] = 1.f;
            if (sph && sd.target_type == ERTB_TARGET_POINT) {
                // same case analysis as primary_entry_sph() (origin = target - d * ray_offset)
                const double *T = sd.target;
                const double off = hs.ray_offset, Rg = S->surface_z;
                double o[3] = { T[0] - dx * off, T[1] - dy * off, T[2] - dz * off };
                double b = o[0] * dx + o[1] * dy + o[2] * dz;
                double oo = o[0] * o[0] + o[1] * o[1] + o[2] * o[2];
                double Rs = oo >= Rt * Rt ? Rt : Rg;
                double disc = b * b - (oo - Rs * Rs);
                if (oo < Rg * Rg || disc < 0.0 || b > 0.0) { t[6] = 0.f; continue; }
                double t0 = -b - sqrt(disc);
                t[0] = (float) ((o[0] + t0 * dx) / Rs);
                t[1] = (float) ((o[1] + t0 * dy) / Rs);
                t[2] = (float) ((o[2] + t0 * dz) / Rs);
                t[6] = oo >= Rt * Rt ? 1.f : 2.f;
            } else if (!sph) {
                double tz = sd.target_type == ERTB_TARGET_POINT ? sd.target[2]
                          : sd.target_type == ERTB_TARGET_NONE ? S->bs_center[2] : sd.target_to_world[11];
                double oz = tz - dz * hs.ray_offset - S->surface_z;
                double Hd = S->has_medium ? S->medium_top - S->surface_z : 0.0;
                t[6] = (!(dz < 0.0) || oz < 0.0) ? 0.f : (oz >= Hd ? 1.f : 2.f);
            }
        }
        CUDA_TRY(cudaMalloc(&hs.d_table, table.size() * sizeof(float)));
        CUDA_TRY(cudaMemcpy(hs.d_table, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

static void fill_sensor_params(const ertb_scene *S, const HostSensor &hs, ErtbSensor &o) {
    const ertb_sensor_desc &sd = hs.desc;
    memset(&o, 0, sizeof o);
    o.type = sd.type;
    o.width = sd.width;
    o.height = sd.height;
    o.target_type = sd.target_type;
    o.use_table = hs.use_table;
    o.table = hs.d_table;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o.to_world[3 * r + c] = (float) sd.to_world[4 * r + c];
    for (int i = 0; i < 3; ++i) o.target[i] = sd.target[i];
    for (int i = 0; i < 12; ++i) o.target_to_world[i] = sd.target_to_world[i];
    for (int i = 0; i < 3; ++i) o.bs_center[i] = S->bs_center[i];
    const double eps = 2.220446049250313e-16 * 0.5 * 1500.0;
    o.bs_radius = fmax(eps, S->bs_radius * (1.0 + eps));
    o.ray_offset = hs.ray_offset;
    o.flux_norm = (float) (2.0 * M_PI / (double) (sd.width * sd.height));
    o.origins = hs.d_origins;
    if (sd.type == ERTB_SENSOR_MRADIANCEMETER) o.in_medium = sd.in_medium;
    if (sd.type == ERTB_SENSOR_PERSPECTIVE) {
        o.cam_origin[0] = sd.to_world[3]; o.cam_origin[1] = sd.to_world[7]; o.cam_origin[2] = sd.to_world[11];
        o.tan_half_fov = (float) tan(0.5 * sd.x_fov_deg * M_PI / 180.0);
        o.aspect = (float) sd.width / (float) sd.height;
        o.near_clip = (float) sd.near_clip;
        o.far_clip = (float) sd.far_clip;
        o.in_medium = sd.in_medium;
    }
}

// ----------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------
// (the header declares every entry point extern "C"; definitions inherit the linkage)

int ertb_abi_version(void) { return ERTB_ABI_VERSION; }
const char *ertb_last_error(void) { return g_error.c_str(); }

int ertb_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error(std::string("no usable CUDA device (there is no CPU fallback): ") +
                  (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
        return -1;
    }
    return n;
}

void ertb_scene_destroy(ertb_scene *S) {
    if (!S) return;
    cudaSetDevice(S->device);
    for (auto &hs : S->sensors) {
        if (hs.d_table) cudaFree(hs.d_table);
        if (hs.d_origins) cudaFree(hs.d_origins);
    }
    slot_release(S->main);
    for (auto &T : S->slots) slot_release(T);
    if (S->batch.d_accum) cudaFree(S->batch.d_accum);
    if (S->batch.d_counters) cudaFree(S->batch.d_counters);
    if (S->d_gl) cudaFree(S->d_gl);
    if (S->d_bsdf_table) cudaFree(S->d_bsdf_table);
    for (void *q : S->d_canopy)
        if (q) cudaFree(q);
    if (S->d_counter) cudaFree(S->d_counter);
    if (S->d_accum) cudaFree(S->d_accum);
    if (S->ev0) cudaEventDestroy(S->ev0);
    if (S->ev1) cudaEventDestroy(S->ev1);
    delete S;
}

// measured_mono: the device indexes the table through the sizes / offsets of its header (ertb_measured.cuh), so
// every one of them is checked against the array length here; nullptr when the table is well formed
static const char *measured_table_error(const float *T, int n_floats) {
    const long n = n_floats;
    if (n < 32 || n >= (1 << 24)) return "measured_mono: invalid table size";
    auto whole = [&](int k, long &v) { v = (long) T[k]; return T[k] >= 0.f && T[k] < 16777216.f && (float) v == T[k]; };
    long h[28];
    for (int k = 0; k < 28; ++k)
        if (!whole(k, h[k])) return "measured_mono: malformed table header";
    const long n_phi = h[0], n_theta = h[1], slices = n_phi * n_theta;
    if (n_phi < 1 || n_theta < 1 || h[2] > 1 || h[3] > 1 || (h[4] != 0 && h[4] != 1 && h[4] != 2 && h[4] != 4))
        return "measured_mono: malformed table header";
    if (h[26] != (n_phi > 1 ? n_theta : 0) || h[27] != (n_theta > 1 ? 1 : 0)) return "measured_mono: malformed slice strides";
    auto fits = [&](long off, long count) { return off >= 32 && count >= 0 && off + count <= n; };
    auto grid = [&](int k) { return h[k] >= 2 && h[k + 1] >= 2; }; // w, h
    if (!grid(5) || !grid(8) || !grid(11) || !grid(16) || !grid(21)) return "Distribution2D(): input array resolution must be >= 2!";
    if (!fits(h[7], h[5] * h[6]) || !fits(h[10], h[8] * h[9]) || !fits(h[23], slices * h[21] * h[22]) ||
        !fits(h[24], n_phi) || !fits(h[25], n_theta))
        return "measured_mono: table offsets out of range";
    for (int k : { 11, 16 })
        if (!fits(h[k + 2], slices * h[k] * h[k + 1]) || !fits(h[k + 3], slices * (h[k + 1] - 1)) ||
            !fits(h[k + 4], slices * h[k + 1] * (h[k] - 1)))
            return "measured_mono: table offsets out of range";
    for (long k = 0; k < n; ++k)
        if (!std::isfinite(T[k])) return "measured_mono: non-finite table entry";
    return nullptr;
}

int ertb_scene_create(const ertb_scene_desc *D, int device, ertb_scene **out) {
    if (!D || !out) return set_error("null argument");
    *out = nullptr;
    if (D->abi_version != ERTB_ABI_VERSION) return set_error("ABI version mismatch");
    int ndev = ertb_device_count();
    if (ndev < 0) return 1;
    if (device < 0 || device >= ndev) return set_error("invalid CUDA device index");
    if (D->geometry != ERTB_GEOM_PLANE_PARALLEL && D->geometry != ERTB_GEOM_SPHERICAL_SHELL)
        return set_error("unsupported geometry");
    if (D->bsdf_type < 0 || D->bsdf_type > ERTB_BSDF_MEASURED_MONO) return set_error("unsupported BSDF type");
    if (D->bsdf_type == ERTB_BSDF_MEASURED_MONO) {
        if (!D->bsdf_table) return set_error("measured_mono: missing table");
        if (const char *why = measured_table_error(D->bsdf_table, D->bsdf_table_res[0])) return set_error(why);
    }
    if (D->bsdf_type == ERTB_BSDF_MQDIFFUSE) {
        if (!D->bsdf_table) return set_error("mqdiffuse: missing table");
        for (int i = 0; i < 3; ++i)
            if (D->bsdf_table_res[i] < 1 || D->bsdf_table_res[i] > 4096) return set_error("mqdiffuse: invalid table resolution");
    }
    if (D->bsdf_type == ERTB_BSDF_OCEAN_GRASP && D->bsdf_params[6] != 0.f)
        return set_error("ocean_grasp: only component=0 (full BRDF) is supported");
    if (D->bsdf_type == ERTB_BSDF_OCEAN_LEGACY && D->bsdf_params[6] != 0.f)
        return set_error("ocean_legacy: only component=0 (full BRDF) is supported");
    if (D->n_sensors < 1 || !D->sensors) return set_error("scene has no sensor");
    if (D->rr_depth <= 0) return set_error("\"rr_depth\" must be set to a value greater than zero!");
    if (D->has_medium) {
        if (D->n_layers < 1 || D->n_layers > ERTB_MAX_LAYERS) return set_error("invalid number of layers");
        if (!D->sigma_t || !D->albedo) return set_error("medium arrays missing");
        if (D->n_phase < 1 || D->n_phase > ERTB_MAX_PHASE) return set_error("invalid number of phase leaves");
        if (D->n_phase > 1 && !D->phase_weight) return set_error("phase_weight missing");
        if (!(D->medium_top > D->medium_bottom)) return set_error("invalid medium extent");
        if (D->medium_top < D->surface_z) return set_error("top of atmosphere below the surface");
    }
    ertb_scene *S = new ertb_scene();
    S->device = device;
    S->geometry = D->geometry;
    S->surface_z = D->surface_z;
    S->medium_bottom = D->medium_bottom;
    S->medium_top = D->medium_top;
    memcpy(S->bs_center, D->bsphere_center, sizeof S->bs_center);
    S->bs_radius = D->bsphere_radius;
    S->has_medium = D->has_medium;
    S->n_layers = D->n_layers;
    S->homogeneous = D->homogeneous;
    S->scale = D->sigma_t_scale;
    if (D->has_medium) {
        S->sigma_t.assign(D->sigma_t, D->sigma_t + D->n_layers);
        S->albedo.assign(D->albedo, D->albedo + D->n_layers);
        S->n_phase = D->n_phase;
        if (D->n_phase > 1)
            S->phase_weight.assign(D->phase_weight, D->phase_weight + (size_t) D->n_phase * D->n_layers);
        for (int k = 0; k < D->n_phase; ++k) {
            const ertb_phase_desc &pd = D->phase[k];
            HostPhase &hp = S->phase[k];
            hp.type = pd.type;
            memcpy(hp.params, pd.params, sizeof hp.params);
            if (pd.type == ERTB_PHASE_TABULATED || pd.type == ERTB_PHASE_TABULATED_IRREGULAR ||
                pd.type == ERTB_PHASE_TABULATED_POLARIZED) {
                if (!pd.values || pd.n_nodes < 2) { delete S; return set_error("tabulated phase: values missing"); }
                hp.values.assign(pd.values, pd.values + pd.n_nodes);
                if (pd.type != ERTB_PHASE_TABULATED) {
                    if (!pd.nodes) { delete S; return set_error("tabulated phase: nodes missing"); }
                    hp.nodes.assign(pd.nodes, pd.nodes + pd.n_nodes);
                }
                if (pd.type == ERTB_PHASE_TABULATED_POLARIZED)
                    for (int m = 0; m < 5; ++m)
                        if (pd.mueller[m]) hp.mueller[m].assign(pd.mueller[m], pd.mueller[m] + pd.n_nodes);
            } else if (pd.type < 0 || pd.type > ERTB_PHASE_TABULATED_POLARIZED) {
                delete S;
                return set_error("unsupported phase function type");
            }
        }
    }
    S->bsdf_type = D->bsdf_type;
    memcpy(S->bsdf_params, D->bsdf_params, sizeof S->bsdf_params);
    if (D->emitter_angular_diameter != 0.0) {
        if (!(D->emitter_angular_diameter > 0.0 && D->emitter_angular_diameter < 180.0)) {
            delete S;
            return set_error("Invalid angular diameter specified! (must be in ]0, 180[)");
        }
        S->astro_diameter = D->emitter_angular_diameter;
        S->hide_emitters = D->hide_emitters != 0;
    }
    memset(S->patch_bsdf_params, 0, sizeof S->patch_bsdf_params);
    if (D->has_patch) {
        if (D->geometry != ERTB_GEOM_PLANE_PARALLEL || D->polarized) {
            delete S;
            return set_error("CentralPatchSurface is supported in unpolarized plane-parallel scenes only");
        }
        if (D->patch_bsdf_type < ERTB_BSDF_DIFFUSE || D->patch_bsdf_type > ERTB_BSDF_HAPKE ||
            D->bsdf_type == ERTB_BSDF_OCEAN_LEGACY || D->bsdf_type >= ERTB_BSDF_OCEAN_MISHCHENKO) {
            delete S;
            return set_error("CentralPatchSurface: only the diffuse / rpv / rtls / hapke BSDFs can be blended");
        }
        S->has_patch = 1;
        S->patch_bsdf_type = D->patch_bsdf_type;
        memcpy(S->patch_bsdf_params, D->patch_bsdf_params, sizeof S->patch_bsdf_params);
        memcpy(S->patch_rect, D->patch_rect, sizeof S->patch_rect);
        S->needs_3d = true;
    }
    memcpy(S->emitter_dir, D->emitter_direction, sizeof S->emitter_dir);
    S->irradiance = D->irradiance;
    S->integrator = D->integrator;
    S->rr_depth = D->rr_depth;
    S->max_depth = D->max_depth;
    S->polarized = D->polarized != 0;
    S->phase_mis = D->phase_mis != 0 && D->has_medium && D->n_phase > 1;
    S->meridian_align = D->meridian_align != 0;
    if (S->polarized && D->integrator == ERTB_INTEGRATOR_VOLPATHMIS) {
        delete S;
        return set_error("This integrator currently does not support polarized mode!"); // volpathmis.cpp:130-132
    }

    if (D->integrator < ERTB_INTEGRATOR_VOLPATH || D->integrator > ERTB_INTEGRATOR_PIECEWISE_VOLPATH) {
        delete S;
        return set_error("unsupported integrator type");
    }
    if (D->integrator == ERTB_INTEGRATOR_PIECEWISE_VOLPATH && S->has_medium &&
        (S->homogeneous || S->geometry != ERTB_GEOM_PLANE_PARALLEL)) {
        // medium.cpp:99-118: only the piecewise medium (plane-parallel layer stack) has the *_real interface
        delete S;
        return set_error("Medium::sample_interaction_real(): not implemented! (piecewise_volpath needs a "
                         "plane-parallel piecewise medium)");
    }

    if (D->n_instances > 0) {
        if (D->geometry != ERTB_GEOM_PLANE_PARALLEL || S->polarized) {
            delete S;
            return set_error("explicit canopies are supported in unpolarized plane-parallel scenes only");
        }
        if (D->n_leaf_groups < 1 || !D->leaf_groups || !D->instance_group || !D->instance_offset) {
            delete S;
            return set_error("canopy arrays missing");
        }
        for (int g = 0; g < D->n_leaf_groups; ++g) {
            const ertb_leaf_group_desc &gd = D->leaf_groups[g];
            if (gd.n_disks < 0 || gd.n_triangles < 0 || (gd.n_disks > 0 && !gd.disks) ||
                gd.n_disks + gd.n_triangles + gd.n_cylinders + gd.n_trunk_disks < 1) { delete S; return set_error("leaf group without primitives"); }
            HostLeafGroup hg;
            if (gd.n_disks) hg.disks.assign(gd.disks, gd.disks + 7 * (size_t) gd.n_disks);
            if (gd.n_triangles) {
                if (!gd.triangles || !gd.triangle_bsdf || !gd.mesh_bsdfs || gd.n_mesh_bsdfs < 1 || gd.n_mesh_bsdfs > 65535) {
                    delete S; return set_error("mesh arrays missing");
                }
                S->has_mesh = true;
                hg.triangles.assign(gd.triangles, gd.triangles + 18 * (size_t) gd.n_triangles);
                hg.triangle_bsdf.assign(gd.triangle_bsdf, gd.triangle_bsdf + gd.n_triangles);
                hg.mesh_bsdfs.assign(gd.mesh_bsdfs, gd.mesh_bsdfs + 2 * (size_t) gd.n_mesh_bsdfs);
                for (float v : hg.triangles)
                    if (!std::isfinite(v)) { delete S; return set_error("mesh contains invalid vertex data"); }
                for (int m : hg.triangle_bsdf)
                    if (m < 0 || m >= gd.n_mesh_bsdfs) { delete S; return set_error("triangle refers to an unknown mesh BSDF"); }
            }
            hg.reflectance = gd.reflectance;
            hg.transmittance = gd.transmittance;
            if ((gd.n_trunk_disks > 0 && !gd.trunk_disks) || (gd.n_cylinders > 0 && !gd.cylinders) ||
                gd.n_trunk_disks < 0 || gd.n_cylinders < 0) { delete S; return set_error("trunk arrays missing"); }
            if (gd.n_trunk_disks) hg.trunk_disks.assign(gd.trunk_disks, gd.trunk_disks + 7 * (size_t) gd.n_trunk_disks);
            if (gd.n_cylinders) hg.cylinders.assign(gd.cylinders, gd.cylinders + 7 * (size_t) gd.n_cylinders);
            hg.trunk_reflectance = gd.trunk_reflectance;
            S->leaf_groups.push_back(hg);
        }
        for (int i = 0; i < D->n_instances; ++i) {
            if (D->instance_group[i] < 0 || D->instance_group[i] >= D->n_leaf_groups) {
                delete S;
                return set_error("instance refers to an unknown leaf group");
            }
            S->instance_group.push_back(D->instance_group[i]);
            for (int k = 0; k < 3; ++k) S->instance_offset.push_back(D->instance_offset[3 * i + k]);
        }
        S->needs_3d = true;
    }
    for (int i = 0; i < D->n_sensors; ++i)
        if (D->sensors[i].type == ERTB_SENSOR_PERSPECTIVE) {
            if (D->geometry != ERTB_GEOM_PLANE_PARALLEL || S->polarized) {
                delete S;
                return set_error("perspective sensors are supported in unpolarized plane-parallel scenes only");
            }
            S->needs_3d = true;
        }

    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete S; return set_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
    cudaDeviceGetAttribute(&S->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&S->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (cudaMalloc(&S->d_counter, 16 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaEventCreate(&S->ev0) != cudaSuccess || cudaEventCreate(&S->ev1) != cudaSuccess) {
        ertb_scene_destroy(S);
        return set_error("device allocation failed");
    }
    if (D->bsdf_type == ERTB_BSDF_MQDIFFUSE || D->bsdf_type == ERTB_BSDF_MEASURED_MONO) { // the measured table lives in global memory (L2-resident)
        size_t n = 1;
        for (int i = 0; i < 3; ++i) { S->bsdf_table_res[i] = D->bsdf_table_res[i]; n *= (size_t) D->bsdf_table_res[i]; }
        if (D->bsdf_type == ERTB_BSDF_MEASURED_MONO) n = (size_t) D->bsdf_table_res[0];
        if (cudaMalloc(&S->d_bsdf_table, n * sizeof(float)) != cudaSuccess ||
            cudaMemcpy(S->d_bsdf_table, D->bsdf_table, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            ertb_scene_destroy(S);
            return set_error("measured BSDF: table upload failed");
        }
    }
    for (int i = 0; i < D->n_sensors; ++i) {
        HostSensor hs;
        hs.desc = D->sensors[i];
        const ertb_sensor_desc &sd = hs.desc;
        if (sd.width < 1 || sd.height < 1) { ertb_scene_destroy(S); return set_error("invalid film size"); }
        if (sd.type == ERTB_SENSOR_MRADIANCEMETER && (!sd.origins || !sd.directions || sd.n_directions < 1)) {
            ertb_scene_destroy(S);
            return set_error("mradiancemeter: origins / directions missing");
        }
        if (sd.type == ERTB_SENSOR_MRADIANCEMETER) hs.origins.assign(sd.origins, sd.origins + 3 * (size_t) sd.n_directions);
        if (sd.type == ERTB_SENSOR_MDISTANT || sd.type == ERTB_SENSOR_MRADIANCEMETER) {
            if (!sd.directions || sd.n_directions < 1) { ertb_scene_destroy(S); return set_error("mdistant: directions missing"); }
            hs.directions.resize(3 * (size_t) sd.n_directions);
            for (int k = 0; k < sd.n_directions; ++k) {
                const double *v = sd.directions + 3 * k;
                double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
                if (!(n > 0.0)) { ertb_scene_destroy(S); return set_error("mdistant: zero-length direction"); }
                for (int c = 0; c < 3; ++c) hs.directions[3 * k + c] = v[c] / n;
            }
        } else if (sd.type == ERTB_SENSOR_PERSPECTIVE) {
            const double z = sd.to_world[11];
            const bool inside = D->has_medium && z > D->surface_z && z < D->medium_top;
            if (!(sd.x_fov_deg > 0.0 && sd.x_fov_deg < 180.0) || !(sd.near_clip > 0.0) || !(sd.far_clip > sd.near_clip)) {
                ertb_scene_destroy(S);
                return set_error("perspective: invalid field of view or clip planes");
            }
            if (z <= D->surface_z) { ertb_scene_destroy(S); return set_error("perspective: the camera is below the surface"); }
            if (inside != (sd.in_medium != 0)) {
                ertb_scene_destroy(S);
                return set_error("perspective: the sensor's medium does not match its position "
                                 "(experiments/_canopy_atmosphere.py:248-258 sets it for cameras inside the atmosphere)");
            }
        } else if (sd.type != ERTB_SENSOR_HDISTANT && sd.type != ERTB_SENSOR_DISTANTFLUX && sd.type != ERTB_SENSOR_MPDISTANT) {
            ertb_scene_destroy(S);
            return set_error("unsupported sensor type");
        }
        hs.desc.directions = nullptr;
        hs.desc.origins = nullptr;
        S->sensors.push_back(hs);
        if (build_sensor(S, S->sensors.back())) { ertb_scene_destroy(S); return 1; }
    }
    if (build_canopy(S)) { ertb_scene_destroy(S); return 1; }
    if (scene_commit(S, S->main)) { ertb_scene_destroy(S); return 1; }
    *out = S;
    return 0;
}

int ertb_scene_update(ertb_scene *S, int param, int index, const float *data, size_t count) {
    if (!S || !data) return set_error("null argument");
    auto need = [&](size_t n) -> int {
        return count == n ? 0 : set_error("ertb_scene_update: size mismatch for parameter " + std::to_string(param));
    };
    switch (param) {
        case ERTB_PARAM_SIGMA_T:
            if (!S->has_medium) return set_error("scene has no medium");
            if (need(S->n_layers)) return 1;
            S->sigma_t.assign(data, data + count);
            break;
        case ERTB_PARAM_ALBEDO:
            if (!S->has_medium) return set_error("scene has no medium");
            if (need(S->n_layers)) return 1;
            S->albedo.assign(data, data + count);
            break;
        case ERTB_PARAM_PHASE_WEIGHT:
            if (S->n_phase <= 1) break; // single leaf: weight is implicit
            if (need((size_t) S->n_phase * S->n_layers)) return 1;
            S->phase_weight.assign(data, data + count);
            break;
        case ERTB_PARAM_PHASE_VALUES:
            if (index < 0 || index >= S->n_phase) return set_error("invalid phase leaf index");
            if (need(S->phase[index].values.size())) return 1;
            S->phase[index].values.assign(data, data + count);
            break;
        case ERTB_PARAM_PHASE_PARAMS:
            if (index < 0 || index >= S->n_phase) return set_error("invalid phase leaf index");
            if (need(4)) return 1;
            memcpy(S->phase[index].params, data, 4 * sizeof(float));
            break;
        case ERTB_PARAM_PHASE_MUELLER: {
            int leaf = index / 5, k = index % 5;
            if (index < 0 || leaf >= S->n_phase || S->phase[leaf].type != ERTB_PHASE_TABULATED_POLARIZED)
                return set_error("invalid polarized phase leaf index");
            if (need(S->phase[leaf].values.size())) return 1;
            S->phase[leaf].mueller[k].assign(data, data + count);
            break;
        }
        case ERTB_PARAM_LEAF_BSDF:
            if (index < 0 || index >= (int) S->leaf_groups.size()) return set_error("invalid leaf group index");
            if (need(2)) return 1;
            S->leaf_groups[index].reflectance = data[0];
            S->leaf_groups[index].transmittance = data[1];
            break;
        case ERTB_PARAM_TRUNK_BSDF:
            if (index < 0 || index >= (int) S->leaf_groups.size()) return set_error("invalid leaf group index");
            if (need(1)) return 1;
            S->leaf_groups[index].trunk_reflectance = data[0];
            break;
        case ERTB_PARAM_MESH_BSDF: {
            const int g = index >> 16, m = index & 0xffff;
            if (index < 0 || g >= (int) S->leaf_groups.size() || 2 * (size_t) m + 1 >= S->leaf_groups[g].mesh_bsdfs.size())
                return set_error("invalid mesh BSDF index");
            if (need(2)) return 1;
            S->leaf_groups[g].mesh_bsdfs[2 * m] = data[0];
            S->leaf_groups[g].mesh_bsdfs[2 * m + 1] = data[1];
            break;
        }
        case ERTB_PARAM_PATCH_BSDF_PARAMS:
            if (!S->has_patch) return set_error("scene has no central patch");
            if (need(ERTB_MAX_BSDF_PARAMS)) return 1;
            memcpy(S->patch_bsdf_params, data, sizeof S->patch_bsdf_params);
            break;
        case ERTB_PARAM_BSDF_PARAMS:
            if (need(ERTB_MAX_BSDF_PARAMS)) return 1;
            memcpy(S->bsdf_params, data, sizeof S->bsdf_params);
            break;
        case ERTB_PARAM_IRRADIANCE:
            if (need(1)) return 1;
            S->irradiance = data[0];
            break;
        default:
            return set_error("unknown parameter id");
    }
    S->dirty = true;
    return 0;
}

int ertb_sensor_pixel_count(const ertb_scene *S, int sensor) {
    if (!S || sensor < 0 || sensor >= (int) S->sensors.size()) return -1;
    return S->sensors[sensor].desc.width * S->sensors[sensor].desc.height;
}

static int launch_render(ertb_scene *S, int sensor, uint64_t seed, uint64_t spp, uint64_t sample_offset,
                         double *accum_dev, unsigned long long *stats_dev, cudaStream_t stream,
                         bool with_stats, TableSlot *slot = nullptr, unsigned long long *counter_dev = nullptr) {
    if (!S) return set_error("null scene");
    if (sensor < 0 || sensor >= (int) S->sensors.size()) return set_error("invalid sensor index");
    if (spp == 0) return set_error("spp must be > 0");
    if (sample_offset + spp >= (1ULL << 40)) return set_error("sample index exceeds 2^40");
    CUDA_TRY(cudaSetDevice(S->device));
    NvtxRange nvtx("ertb:launch_render");
    if (slot) {
        if (scene_commit(S, *slot)) return 1;
    } else {
        // Outside the batch entry points a scene has ONE table slot and ONE work counter: renders of a scene are
        // serialised. A render queued on another stream than the previous one waits for it, and a commit (a
        // synchronous upload into the tables the previous render may still be reading) waits for the caller's stream.
        if (S->last_stream_valid && S->last_stream != stream) CUDA_TRY(cudaStreamSynchronize(S->last_stream));
        if (S->dirty) {
            CUDA_TRY(cudaStreamSynchronize(stream));
            if (scene_commit(S, S->main)) return 1;
        }
        S->last_stream = stream;
        S->last_stream_valid = true;
    }
    ErtbParams P = S->base;
    if (!counter_dev) counter_dev = S->d_counter;
    const HostSensor &hs = S->sensors[sensor];
    fill_sensor_params(S, hs, P.sensor);
    if (hs.desc.type == ERTB_SENSOR_MDISTANT || hs.desc.type == ERTB_SENSOR_MRADIANCEMETER) { P.sensor_up[0] = 0.f; P.sensor_up[1] = 1.f; P.sensor_up[2] = 0.f; }
    else { P.sensor_up[0] = (float) hs.desc.to_world[1]; P.sensor_up[1] = (float) hs.desc.to_world[5]; P.sensor_up[2] = (float) hs.desc.to_world[9]; }
    P.seed = seed;
    P.spp = spp;
    P.sample_offset = sample_offset;
    P.n_pixels = (unsigned) (hs.desc.width * hs.desc.height);
    P.tw = ERTB_TW;
    P.tw = env_int_clamped("ERTB_TW", P.tw, 1, 32); // tuning knob

    // persistent grid: as many CTAs as can be resident (queried, not assumed)
    const bool sph = P.spherical;
    // two execution models of the same estimator: warp-private shared-memory pools (default)
    // or one register-resident path per lane (ERTB_KERNEL=legacy, kept for A/B measurements)
    bool use_pool = true;
    if (const char *e = getenv("ERTB_KERNEL")) use_pool = strcmp(e, "legacy") != 0;
    int blocks_per_sm = 0;
    const bool pol = S->polarized != 0;
    const bool pw = S->base.piecewise != 0;
    const bool c3d = S->needs_3d; // canopy / perspective camera: the 3D kernel (ertb_canopy.cuh)
    if (pol || pw) use_pool = true; // the polarized and the piecewise paths exist in the pool kernel only
    if (c3d) use_pool = false;
    // general primary rays, the finite solar disc (astroobject) and the mixture weight of multiphase exist in the GEN
    // instances of the pool kernel (compiled with statistics on) and in the general instances of the 3D kernel
    const bool gen_needed = !c3d && (hs.desc.type == ERTB_SENSOR_MPDISTANT || hs.desc.type == ERTB_SENSOR_MRADIANCEMETER ||
                                     S->astro_diameter > 0.0 || S->phase_mis);
    if (gen_needed) use_pool = true;
    const bool bands = S->base.n_bands > 1;
    int block = c3d ? ERTB_CANOPY_BLOCK : (use_pool ? ertb_pool_block(pol, bands && !pw) : ERTB_BLOCK);
    size_t smem = use_pool ? ertb_pool_smem_bytes((size_t) S->base.blob_bytes, pol, bands && !pw) : (size_t) S->base.blob_bytes;
    // large tables: the pools of a full-size CTA do not fit next to them -- launch smaller CTAs of the same instance
    // (every warp owns its pool; the kernel takes the number of warps from blockDim)
    while (use_pool && smem > (size_t) S->max_smem_optin && block > 128) {
        block -= 64;
        smem = ertb_pool_smem_bytes((size_t) S->base.blob_bytes, pol, bands && !pw, block);
    }
    if (use_pool && smem > (size_t) S->max_smem_optin) { // huge tables: fall back to the register kernel
        if (pol || pw || gen_needed) return set_error("scene tables leave no shared memory for the path pools");
        use_pool = false;
        smem = (size_t) S->base.blob_bytes;
        block = ERTB_BLOCK;
    }
    // ... and so do the plugins added after SURVEY 8a (glint family, mqdiffuse): the lean instances keep the
    // instruction stream of the headline configurations free of them (the register and 3D kernels are general)
    const bool gen = gen_needed || (use_pool && !c3d && S->bsdf_type >= ERTB_BSDF_OCEAN_MISHCHENKO);
    // (attribute + occupancy are queried once per kernel instantiation and table size, then cached)
    // The dynamic shared-memory limit is an attribute of the FUNCTION (per device), shared by every scene and table
    // size in the process: it is only ever raised.  (A spectral loop whose contexts alternate between two blob sizes
    // used to lower it again on the second size and fail with "invalid argument" on the third launch, once the 448-thread
    // CTAs needed more than the 48 KB every function has by default.)
#define ERTB_OCC(KERNEL)                                                                              \
    do {                                                                                              \
        {                                                                                             \
            std::lock_guard<std::mutex> lock(g_smem_attr_mutex);                                      \
            size_t &cur = g_smem_attr[std::make_pair(S->device, (const void *) KERNEL)];              \
            if (smem > cur) {                                                                         \
                CUDA_TRY(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
                cur = smem;                                                                           \
            }                                                                                         \
        }                                                                                             \
        auto key = std::make_pair((const void *) KERNEL, smem);                                       \
        auto it = S->occupancy.find(key);                                                             \
        if (it != S->occupancy.end()) { blocks_per_sm = it->second; break; }                          \
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, KERNEL, block, smem)); \
        S->occupancy[key] = blocks_per_sm;                                                            \
    } while (0)
#define ERTB_POOL_VARIANT_B(MACRO, SPH_, POL_, PW_, B_)                                               \
    do {                                                                                              \
        if (gen) MACRO((ertb_render_pool_kernel<SPH_, true, POL_, PW_, true, B_, true>));             \
        else if (with_stats) MACRO((ertb_render_pool_kernel<SPH_, true, POL_, PW_, true, B_>));       \
        else if (coll) MACRO((ertb_render_pool_kernel<SPH_, false, POL_, PW_, true, B_>));            \
        else MACRO((ertb_render_pool_kernel<SPH_, false, POL_, PW_, false, B_>));                     \
    } while (0)
#define ERTB_POOL_VARIANT(MACRO, SPH_, POL_, PW_)                                                     \
    do {                                                                                              \
        if (!PW_ && bands) ERTB_POOL_VARIANT_B(MACRO, SPH_, POL_, false, true);                       \
        else ERTB_POOL_VARIANT_B(MACRO, SPH_, POL_, PW_, false);                                      \
    } while (0)
#define ERTB_CANOPY_VARIANT_V(MACRO, MESH_, VOTE_)                                                    \
    do {                                                                                              \
        if (pw) { if (with_stats) MACRO((ertb_canopy_kernel<true, true, MESH_, VOTE_>)); else MACRO((ertb_canopy_kernel<false, true, MESH_, VOTE_>)); } \
        else    { if (with_stats) MACRO((ertb_canopy_kernel<true, false, MESH_, VOTE_>)); else MACRO((ertb_canopy_kernel<false, false, MESH_, VOTE_>)); } \
    } while (0)
#define ERTB_CANOPY_VARIANT(MACRO, MESH_)                                                             \
    do {                                                                                              \
        if (S->small_groups) ERTB_CANOPY_VARIANT_V(MACRO, MESH_, true);                               \
        else ERTB_CANOPY_VARIANT_V(MACRO, MESH_, false);                                              \
    } while (0)
#define ERTB_DISPATCH(MACRO)                                                                          \
    do {                                                                                              \
        if (c3d) {                                                                                    \
            if (S->has_mesh || S->bsdf_type >= ERTB_BSDF_OCEAN_MISHCHENKO || S->astro_diameter > 0.0 || S->phase_mis) { /* the general instances */ \
                ERTB_CANOPY_VARIANT(MACRO, true);                                                     \
            } else ERTB_CANOPY_VARIANT(MACRO, false);                                                 \
        } else if (pw && pol) ERTB_POOL_VARIANT(MACRO, false, true, true);                            \
        else if (pw) ERTB_POOL_VARIANT(MACRO, false, false, true);                                    \
        else if (use_pool && pol) {                                                                   \
            if (sph) ERTB_POOL_VARIANT(MACRO, true, true, false);                                     \
            else ERTB_POOL_VARIANT(MACRO, false, true, false);                                        \
        } else if (use_pool) {                                                                        \
            if (sph) ERTB_POOL_VARIANT(MACRO, true, false, false);                                    \
            else ERTB_POOL_VARIANT(MACRO, false, false, false);                                       \
        } else {                                                                                      \
            if (sph) { if (with_stats) MACRO((ertb_render_kernel<true, true>)); else MACRO((ertb_render_kernel<true, false>)); } \
            else     { if (with_stats) MACRO((ertb_render_kernel<false, true>)); else MACRO((ertb_render_kernel<false, false>)); } \
        }                                                                                             \
    } while (0)
    // the occupancy query does not depend on the flush variant; the launch picks it from the chunk size
    bool coll = false;
    ERTB_DISPATCH(ERTB_OCC);
    if (blocks_per_sm < 1) return set_error("render kernel cannot be resident on this device");
    if (use_pool) {
        // (banded walks lose lanes at band boundaries: keep stepping longer, C3 +4 %; polarized records are
        // expensive to move: stay in the walk until few are left, C5 +2 %)
        P.tw = 32; P.twi = pol ? 4 : (bands ? 8 : 16);
        P.tw = env_int_clamped("ERTB_POOL_TW", P.tw, 1, ERTB_POOL_NS);
        P.twi = env_int_clamped("ERTB_POOL_TWI", P.twi, 1, 32);
    }
    // chunk: a few thousand paths, so that the queue hands out >> n_warps chunks
    const unsigned long long total = (unsigned long long) P.n_pixels * spp;
    const unsigned long long n_warps = (unsigned long long) S->sm_count * blocks_per_sm * (block / 32);
    unsigned long long chunk = total / (n_warps * 16ULL);
    if (chunk > 4096) chunk = 4096;
    if (chunk < 32) chunk = 32;
    if (chunk > spp) chunk = spp;
    P.chunk = (unsigned) chunk;
    // Film flush. Lanes switch pixel once or twice per chunk, and a per-lane flush is three float64 atomics on one of
    // 3 x n_pixels addresses: ~10^7 atomics per C2 launch. Round 1 used per-lane flushes for chunks of >= 512 paths
    // ("hidden behind the walk") and the warp-collective flush (lanes grouped by pixel, shuffle reduction, one lane
    // issues the atomics) below that. Measured in round 2 (profiles/r02k_film_flush.md): the per-lane variant's time
    // depends on WHERE the accumulators were allocated -- 3.49 ms with bench.py's torch buffer, 4.05 ms with the
    // library's own cudaMalloc'ed film behind ertb_render / mi_render, same kernel, same parameters -- while the
    // collective variant runs in 3.60 ms wherever they are. The pool kernel therefore always flushes collectively;
    // ERTB_FLUSH=lane keeps the other variant reachable for A/B runs.
    coll = use_pool;
    {
        static const char *e = getenv("ERTB_FLUSH"); // developer knob (lane | coll)
        if (e && strcmp(e, "coll") == 0) coll = true;
        if (e && strcmp(e, "lane") == 0) coll = chunk < 512;
    }
    if (coll && use_pool && !with_stats) { // same resources, but the attribute is per instantiation
        int bps = 0;
        std::swap(bps, blocks_per_sm);
        ERTB_DISPATCH(ERTB_OCC);
        if (blocks_per_sm < bps) bps = blocks_per_sm;
        blocks_per_sm = bps;
        if (blocks_per_sm < 1) return set_error("render kernel cannot be resident on this device");
    }
#undef ERTB_OCC
    P.chunks_per_pixel = (unsigned) ((spp + chunk - 1) / chunk);
    P.n_chunks = (unsigned long long) P.chunks_per_pixel * P.n_pixels;
    P.work_counter = counter_dev;
    P.accum = accum_dev;
    P.stats = stats_dev;
    CUDA_TRY(cudaMemsetAsync(counter_dev, 0, sizeof(unsigned long long), stream));

    // small renders: do not spread the paths over more warps than can keep their pools busy
    unsigned long long warp_paths = use_pool ? ERTB_MIN_WARP_PATHS : 32ULL;
    warp_paths = (unsigned long long) env_int_clamped("ERTB_MIN_WARP_PATHS", (int) warp_paths, 32, 1 << 30);
    if (warp_paths < 32) warp_paths = 32;
    unsigned long long want_warps = (total + warp_paths - 1) / warp_paths;
    unsigned long long want_blocks = (want_warps * 32ULL + block - 1) / block;
    unsigned long long grid = (unsigned long long) S->sm_count * blocks_per_sm;
    if (slot) {
        // batch items share the device: with a fraction of the resident CTAs each, consecutive
        // items run side by side and the drain of one overlaps the bulk of the next
        int div = ERTB_BATCH_GRID_DIV;
        div = env_int_clamped("ERTB_BATCH_GRID_DIV", div, 1, 64);
        const int n_items = (int) S->batch.sensors.size(); // a batch of one or two items has nothing to share with
        if (n_items < div) div = n_items > 0 ? n_items : 1;
        if (div > 1 && blocks_per_sm / div >= 1) grid = (unsigned long long) S->sm_count * (blocks_per_sm / div);
    }
    if (want_blocks < grid) grid = want_blocks ? want_blocks : 1;
    {
        static const bool dump = getenv("ERTB_DUMP_PARAMS") != nullptr; // developer aid
        if (dump) {
            unsigned long long h = 1469598103934665603ULL;
            const unsigned char *pb = (const unsigned char *) &P;
            for (size_t i = 0; i < sizeof P; ++i) { h ^= pb[i]; h *= 1099511628211ULL; }
            fprintf(stderr, "ertb launch: grid %llu block %d smem %zu tw %d twi %d chunk %u n_chunks %llu spp %llu off %llu seed %llu "
                            "npix %u stats %p with_stats %d coll %d bands %d use_table %d sensor_type %d params_hash %016llx\n",
                    grid, block, smem, P.tw, P.twi, P.chunk, P.n_chunks, (unsigned long long) P.spp,
                    (unsigned long long) P.sample_offset, (unsigned long long) P.seed, P.n_pixels, (void *) P.stats,
                    (int) with_stats, (int) coll, P.n_bands, P.sensor.use_table, P.sensor.type, h);
        }
    }
#define ERTB_LAUNCH(KERNEL) KERNEL<<<(unsigned) grid, block, smem, stream>>>(P)
    ERTB_DISPATCH(ERTB_LAUNCH);
#undef ERTB_LAUNCH
#undef ERTB_DISPATCH
#undef ERTB_POOL_VARIANT
#undef ERTB_POOL_VARIANT_B
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int ertb_render(ertb_scene *S, int sensor, uint64_t seed, uint64_t spp, uint64_t sample_offset,
                double *sum_wl, double *sum_l, double *sum_l2, ertb_render_stats *stats) {
    return ertb_render_stokes(S, sensor, seed, spp, sample_offset, sum_wl, sum_l, sum_l2, nullptr, stats);
}

int ertb_render_stokes(ertb_scene *S, int sensor, uint64_t seed, uint64_t spp, uint64_t sample_offset,
                       double *sum_wl, double *sum_l, double *sum_l2, double *sum_stokes,
                       ertb_render_stats *stats) {
    if (!S) return set_error("null scene");
    int npix = ertb_sensor_pixel_count(S, sensor);
    if (npix <= 0) return set_error("invalid sensor index");
    CUDA_TRY(cudaSetDevice(S->device));
    const int rows = S->polarized ? 7 : 3;
    size_t bytes = (size_t) rows * npix * sizeof(double);
    if (bytes > S->d_accum_capacity) {
        if (S->d_accum) cudaFree(S->d_accum);
        CUDA_TRY(cudaMalloc(&S->d_accum, bytes));
        S->d_accum_capacity = bytes;
    }
    static const bool timing = getenv("ERTB_TIMING") != nullptr; // developer aid: host time of each step of the call
    const double t0 = timing ? wall_ms() : 0.0;
    CUDA_TRY(cudaMemsetAsync(S->d_accum, 0, bytes, 0));
    unsigned long long *d_stats = S->d_counter + 8;
    CUDA_TRY(cudaMemsetAsync(d_stats, 0, 8 * sizeof(unsigned long long), 0));
    CUDA_TRY(cudaEventRecord(S->ev0, 0));
    const double t1 = timing ? wall_ms() : 0.0;
    if (launch_render(S, sensor, seed, spp, sample_offset, S->d_accum, d_stats, 0, stats != nullptr)) return 1;
    const double t2 = timing ? wall_ms() : 0.0;
    CUDA_TRY(cudaEventRecord(S->ev1, 0));
    std::vector<double> host((size_t) rows * npix);
    {
        NvtxRange nvtx("ertb:film_readback");
        CUDA_TRY(cudaMemcpy(host.data(), S->d_accum, bytes, cudaMemcpyDeviceToHost));
    }
    if (timing) {
        float kms = 0.f;
        cudaEventElapsedTime(&kms, S->ev0, S->ev1);
        fprintf(stderr, "ertb_render: memsets %.3f ms, commit + launch %.3f ms, wait + read-back %.3f ms, kernel (events) %.3f ms\n",
                t1 - t0, t2 - t1, wall_ms() - t2, kms);
    }
    if (sum_stokes) {
        if (S->polarized) memcpy(sum_stokes, host.data() + 3 * (size_t) npix, 4 * (size_t) npix * sizeof(double));
        else memset(sum_stokes, 0, 4 * (size_t) npix * sizeof(double));
    }
    if (sum_wl) memcpy(sum_wl, host.data(), npix * sizeof(double));
    if (sum_l) memcpy(sum_l, host.data() + npix, npix * sizeof(double));
    if (sum_l2) memcpy(sum_l2, host.data() + 2 * (size_t) npix, npix * sizeof(double));
    if (stats) {
        unsigned long long h[8];
        CUDA_TRY(cudaMemcpy(h, d_stats, sizeof h, cudaMemcpyDeviceToHost));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, S->ev0, S->ev1));
        memset(stats, 0, sizeof *stats);
        stats->n_paths = h[0];
        stats->trips_main = h[1];
        stats->trips_nee = h[2];
        stats->n_scatter = h[3];
        stats->n_surface = h[4];
        stats->device_ms = ms;
        stats->n_launches = 1;
        stats->n_bands = S->base.n_bands;
    }
    return 0;
}

int ertb_render_device(ertb_scene *S, int sensor, uint64_t seed, uint64_t spp, uint64_t sample_offset,
                       void *accum_dev, void *stats_dev, void *stream) {
    if (!accum_dev) return set_error("null accumulator buffer");
    return launch_render(S, sensor, seed, spp, sample_offset, (double *) accum_dev,
                         (unsigned long long *) stats_dev, (cudaStream_t) stream, stats_dev != nullptr);
}

// ----------------------------------------------------------------------------
// Pipelined multi-context rendering (the contexts x sensors loop of _render.py:433-468)
// ----------------------------------------------------------------------------
static double wall_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double) ts.tv_sec + 1e-6 * (double) ts.tv_nsec;
}

static int batch_drain(ertb_scene *S) {
    for (auto &T : S->slots)
        if (T.stream) {
            CUDA_TRY(cudaStreamSynchronize(T.stream));
            T.busy = false;
        }
    return 0;
}

int ertb_batch_begin(ertb_scene *S, int n_items, const int *sensors, int with_stats) {
    if (!S || !sensors) return set_error("null argument");
    if (n_items < 1) return set_error("ertb_batch_begin: empty batch");
    CUDA_TRY(cudaSetDevice(S->device));
    BatchState &B = S->batch;
    if (batch_drain(S)) return 1; // an abandoned batch may still be running
    B.open = false;
    const int rows = S->polarized ? 7 : 3;
    B.sensors.assign(sensors, sensors + n_items);
    B.offsets.resize(n_items);
    size_t total = 0;
    for (int i = 0; i < n_items; ++i) {
        int npix = ertb_sensor_pixel_count(S, sensors[i]);
        if (npix <= 0) return set_error("ertb_batch_begin: invalid sensor index");
        B.offsets[i] = total;
        total += (size_t) rows * npix;
    }
    B.total = total;
    if (total > B.d_accum_capacity) {
        if (B.d_accum) cudaFree(B.d_accum);
        B.d_accum = nullptr;
        B.d_accum_capacity = 0;
        CUDA_TRY(cudaMalloc(&B.d_accum, total * sizeof(double)));
        B.d_accum_capacity = total;
    }
    if ((size_t) n_items * 16 > B.d_counters_capacity) {
        if (B.d_counters) cudaFree(B.d_counters);
        B.d_counters = nullptr;
        B.d_counters_capacity = 0;
        CUDA_TRY(cudaMalloc(&B.d_counters, (size_t) n_items * 16 * sizeof(unsigned long long)));
        B.d_counters_capacity = (size_t) n_items * 16;
    }
    CUDA_TRY(cudaMemset(B.d_accum, 0, total * sizeof(double)));
    CUDA_TRY(cudaMemset(B.d_counters, 0, (size_t) n_items * 16 * sizeof(unsigned long long)));
    for (auto &T : S->slots)
        if (!T.stream) {
            CUDA_TRY(cudaStreamCreateWithFlags(&T.stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&T.done, cudaEventDisableTiming));
            T.async = true;
        }
    B.with_stats = with_stats != 0;
    B.next = 0;
    B.open = true;
    B.t_begin = wall_ms();
    return 0;
}

int ertb_batch_push(ertb_scene *S, int sensor, uint64_t seed, uint64_t spp, uint64_t sample_offset) {
    if (!S) return set_error("null scene");
    BatchState &B = S->batch;
    if (!B.open) return set_error("ertb_batch_push: no open batch");
    if (B.next >= (int) B.sensors.size()) return set_error("ertb_batch_push: more items than announced");
    const int i = B.next;
    if (sensor != B.sensors[i]) return set_error("ertb_batch_push: sensor differs from the announced one");
    CUDA_TRY(cudaSetDevice(S->device));
    TableSlot &T = S->slots[i % ERTB_BATCH_SLOTS];
    if (T.busy) { // the render that last used this slot's tables must have finished
        CUDA_TRY(cudaEventSynchronize(T.done));
        T.busy = false;
    }
    unsigned long long *ctr = B.d_counters + (size_t) i * 16;
    if (launch_render(S, sensor, seed, spp, sample_offset, B.d_accum + B.offsets[i],
                      B.with_stats ? ctr + 8 : nullptr, T.stream, B.with_stats, &T, ctr))
        return 1;
    CUDA_TRY(cudaEventRecord(T.done, T.stream));
    T.busy = true;
    B.next = i + 1;
    return 0;
}

int ertb_batch_end(ertb_scene *S, double *accum_out, size_t count, ertb_render_stats *stats, double *elapsed_ms) {
    if (!S) return set_error("null scene");
    BatchState &B = S->batch;
    if (!B.open) return set_error("ertb_batch_end: no open batch");
    CUDA_TRY(cudaSetDevice(S->device));
    B.open = false;
    if (batch_drain(S)) return 1;
    if (elapsed_ms) *elapsed_ms = wall_ms() - B.t_begin;
    if (B.next != (int) B.sensors.size()) return set_error("ertb_batch_end: fewer items pushed than announced");
    if (accum_out) {
        if (count != B.total) return set_error("ertb_batch_end: output size mismatch");
        CUDA_TRY(cudaMemcpy(accum_out, B.d_accum, B.total * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (stats) {
        std::vector<unsigned long long> h((size_t) B.next * 16);
        CUDA_TRY(cudaMemcpy(h.data(), B.d_counters, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (int i = 0; i < B.next; ++i) {
            const unsigned long long *c = &h[(size_t) i * 16 + 8];
            memset(&stats[i], 0, sizeof stats[i]);
            stats[i].n_paths = c[0];
            stats[i].trips_main = c[1];
            stats[i].trips_nee = c[2];
            stats[i].n_scatter = c[3];
            stats[i].n_surface = c[4];
            stats[i].n_launches = 1;
            stats[i].n_bands = S->base.n_bands;
        }
    }
    return 0;
}

// ----------------------------------------------------------------------------
// KAT kernels: the same device functions the render kernel uses, one thread per query
// ----------------------------------------------------------------------------
__global__ void kat_bsdf_eval_kernel(ErtbParams P, size_t n, const float *wi, const float *wo, float *out) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 a = mk3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), b = mk3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
    float ci = a.z, co = b.z;
    float v = 0.f;
    if (bsdf_is_local(P.bsdf_type)) v = lf_eval(P, a, b);
    else if (ci > 0.f && co > 0.f) v = bsdf_f(P, ci, co, cos_dphi(ci, co, dot3(a, b))) * co;
    out[i] = v;
}
__global__ void kat_bsdf_sample_kernel(ErtbParams P, size_t n, const float *wi, const float *u, float *wo, float *w) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 a = mk3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
    f3 o;
    float weight = 0.f;
    if (bsdf_is_local(P.bsdf_type)) {
        weight = lf_sample(P, a, u[3 * i], u[3 * i + 1], u[3 * i + 2], o);
    } else {
        o = cosine_hemisphere(u[3 * i + 1], u[3 * i + 2]);
        if (a.z > 0.f && o.z > 0.f) weight = bsdf_f(P, a.z, o.z, cos_dphi(a.z, o.z, dot3(a, o))) * ERTB_PI;
    }
    wo[3 * i] = o.x; wo[3 * i + 1] = o.y; wo[3 * i + 2] = o.z;
    w[i] = weight;
}
__global__ void kat_phase_eval_kernel(ErtbParams P, int leaf, size_t n, const float *c, float *out) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = leaf_eval(P.blob, P.leaf[leaf], -c[i]); // graphics -> physics cosine
}
__global__ void kat_phase_sample_kernel(ErtbParams P, int leaf, size_t n, const float *u, float *ct, float *w, float *pdf) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ww, pp;
    ct[i] = leaf_sample(P.blob, P.leaf[leaf], u[2 * i], ww, pp);
    w[i] = ww;
    pdf[i] = pp;
}

__global__ void kat_phase_mueller_kernel(ErtbParams P, int leaf, size_t n, const float *wi, const float *wo, float *M, float *pdf) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m[16], pp;
    leaf_mueller(P.blob, P.leaf[leaf], mk3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), mk3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), m, pp);
    for (int k = 0; k < 16; ++k) M[16 * i + k] = m[k];
    pdf[i] = pp;
}

__global__ void kat_bsdf_mueller_kernel(ErtbParams P, size_t n, const float *wi, const float *wo, float *M) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 a = mk3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), b = mk3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
    float m[16];
    // shading frame = the axes: the world implicit bases are then the local ones BSDF::eval refers to
    if (bsdf_is_mueller(P.bsdf_type)) {
        lf_eval_mueller(P, false, a, b, mk3(1.f, 0.f, 0.f), mk3(0.f, 1.f, 0.f), mk3(0.f, 0.f, 1.f), m);
    } else {
        for (int k = 0; k < 16; ++k) m[k] = 0.f;
        if (bsdf_is_local(P.bsdf_type)) m[0] = lf_eval(P, a, b);
        else if (a.z > 0.f && b.z > 0.f) m[0] = bsdf_f(P, a.z, b.z, cos_dphi(a.z, b.z, dot3(a, b))) * b.z;
    }
    for (int k = 0; k < 16; ++k) M[16 * i + k] = m[k];
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { return cudaMalloc(&p, n * sizeof(T)) == cudaSuccess ? 0 : 1; }
};

static int kat_prepare(ertb_scene *S) {
    if (!S) return set_error("null scene");
    CUDA_TRY(cudaSetDevice(S->device));
    if (S->dirty && scene_commit(S, S->main)) return 1;
    return 0;
}
#define KAT_GRID(n) (unsigned) (((n) + 127) / 128), 128

int ertb_kat_bsdf_eval(ertb_scene *S, size_t n, const float *wi, const float *wo, float *out) {
    if (kat_prepare(S)) return 1;
    DevBuf<float> a, b, o;
    if (a.alloc(3 * n) || b.alloc(3 * n) || o.alloc(n)) return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(a.p, wi, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(b.p, wo, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    kat_bsdf_eval_kernel<<<KAT_GRID(n)>>>(S->base, n, a.p, b.p, o.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, o.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
int ertb_kat_bsdf_sample(ertb_scene *S, size_t n, const float *wi, const float *u, float *wo, float *weight) {
    if (kat_prepare(S)) return 1;
    DevBuf<float> a, b, o, w;
    if (a.alloc(3 * n) || b.alloc(3 * n) || o.alloc(3 * n) || w.alloc(n)) return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(a.p, wi, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(b.p, u, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    kat_bsdf_sample_kernel<<<KAT_GRID(n)>>>(S->base, n, a.p, b.p, o.p, w.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(wo, o.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(weight, w.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
int ertb_kat_bsdf_mueller(ertb_scene *S, size_t n, const float *wi, const float *wo, float *mueller) {
    if (kat_prepare(S)) return 1;
    DevBuf<float> a, b, m;
    if (a.alloc(3 * n) || b.alloc(3 * n) || m.alloc(16 * n)) return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(a.p, wi, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(b.p, wo, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    kat_bsdf_mueller_kernel<<<KAT_GRID(n)>>>(S->base, n, a.p, b.p, m.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(mueller, m.p, 16 * n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
int ertb_kat_phase_eval(ertb_scene *S, int leaf, size_t n, const float *c, float *out) {
    if (kat_prepare(S)) return 1;
    if (leaf < 0 || leaf >= S->n_phase) return set_error("invalid phase leaf index");
    DevBuf<float> a, o;
    if (a.alloc(n) || o.alloc(n)) return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(a.p, c, n * sizeof(float), cudaMemcpyHostToDevice));
    kat_phase_eval_kernel<<<KAT_GRID(n)>>>(S->base, leaf, n, a.p, o.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, o.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
int ertb_kat_phase_sample(ertb_scene *S, int leaf, size_t n, const float *u, float *ct, float *weight, float *pdf) {
    if (kat_prepare(S)) return 1;
    if (leaf < 0 || leaf >= S->n_phase) return set_error("invalid phase leaf index");
    DevBuf<float> a, c, w, p;
    if (a.alloc(2 * n) || c.alloc(n) || w.alloc(n) || p.alloc(n)) return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(a.p, u, 2 * n * sizeof(float), cudaMemcpyHostToDevice));
    kat_phase_sample_kernel<<<KAT_GRID(n)>>>(S->base, leaf, n, a.p, c.p, w.p, p.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(ct, c.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(weight, w.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(pdf, p.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

int ertb_kat_phase_mueller(ertb_scene *S, int leaf, size_t n, const float *wi, const float *wo, float *mueller, float *pdf) {
    if (kat_prepare(S)) return 1;
    if (leaf < 0 || leaf >= S->n_phase) return set_error("invalid phase leaf index");
    DevBuf<float> a, b, m, p;
    if (a.alloc(3 * n) || b.alloc(3 * n) || m.alloc(16 * n) || p.alloc(n)) return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(a.p, wi, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(b.p, wo, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    kat_phase_mueller_kernel<<<KAT_GRID(n)>>>(S->base, leaf, n, a.p, b.p, m.p, p.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(mueller, m.p, 16 * n * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(pdf, p.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

// piecewise medium: analytic free flight / exact transmittance (ertb_piecewise.cuh)
__global__ void kat_piecewise_sample_kernel(ErtbParams P, size_t n, const float *z, const float *mu, const float *u,
                                            float *t, int *kind) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s, h;
    // (the render kernel draws u on the 2^-24 grid, where 1 - u is exact; arbitrary test inputs are not)
    kind[i] = pw_flight(P, P.blob, z[i], mu[i], (float) -log1p(-(double) u[i]), s, h);
    t[i] = s;
}
__global__ void kat_piecewise_tr_kernel(ErtbParams P, size_t n, const float *z, const float *mu, float *tr) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    tr[i] = mu[i] > 0.f ? pw_transmittance_up(P, P.blob, z[i], mu[i]) : 0.f;
}

int ertb_kat_piecewise_sample(ertb_scene *S, size_t n, const float *altitude, const float *mu, const float *u,
                              float *distance, int32_t *kind) {
    if (kat_prepare(S)) return 1;
    if (!S->base.piecewise) return set_error("not a piecewise scene");
    DevBuf<float> a, b, c, t;
    DevBuf<int> k;
    if (a.alloc(n) || b.alloc(n) || c.alloc(n) || t.alloc(n) || k.alloc(n)) return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(a.p, altitude, n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(b.p, mu, n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c.p, u, n * sizeof(float), cudaMemcpyHostToDevice));
    kat_piecewise_sample_kernel<<<KAT_GRID(n)>>>(S->base, n, a.p, b.p, c.p, t.p, k.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(distance, t.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(kind, k.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}
int ertb_kat_piecewise_transmittance(ertb_scene *S, size_t n, const float *altitude, const float *mu, float *tr) {
    if (kat_prepare(S)) return 1;
    if (!S->base.piecewise) return set_error("not a piecewise scene");
    DevBuf<float> a, b, t;
    if (a.alloc(n) || b.alloc(n) || t.alloc(n)) return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(a.p, altitude, n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(b.p, mu, n * sizeof(float), cudaMemcpyHostToDevice));
    kat_piecewise_tr_kernel<<<KAT_GRID(n)>>>(S->base, n, a.p, b.p, t.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(tr, t.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

// Sensor rays as the reference defines them (origin = target - d * ray_offset).
__global__ void kat_sensor_ray_kernel(ErtbParams P, double ray_offset, size_t n, const float *fs_,
                                      const float *as_, double *origin, double *dir, float *weight) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ErtbSensor &S = P.sensor;
    float fx = fs_[2 * i], fy = fs_[2 * i + 1], ax = as_[2 * i], ay = as_[2 * i + 1];
    f3 d, fs = mk3(1.f, 0.f, 0.f), ft = mk3(0.f, 1.f, 0.f);
    float w = 1.f;
    if (S.type == ERTB_SENSOR_PERSPECTIVE) { // same arithmetic as canopy_primary()
        f3 dc = normalize3(mk3((1.f - 2.f * fx) * S.tan_half_fov, (1.f - 2.f * fy) * S.tan_half_fov / S.aspect, 1.f));
        const float *M = S.to_world;
        d = normalize3(mk3(M[0] * dc.x + M[1] * dc.y + M[2] * dc.z, M[3] * dc.x + M[4] * dc.y + M[5] * dc.z,
                           M[6] * dc.x + M[7] * dc.y + M[8] * dc.z));
        float near_t = S.near_clip / dc.z;
        origin[3 * i] = S.cam_origin[0] + (double) (near_t * d.x);
        origin[3 * i + 1] = S.cam_origin[1] + (double) (near_t * d.y);
        origin[3 * i + 2] = S.cam_origin[2] + (double) (near_t * d.z);
        dir[3 * i] = d.x; dir[3 * i + 1] = d.y; dir[3 * i + 2] = d.z;
        weight[i] = 1.f;
        return;
    }
    if (S.type == ERTB_SENSOR_MRADIANCEMETER) {
        int idx = min((int) (fx * (float) S.width), S.width - 1);
        const float *t = S.table + 8 * (size_t) idx;
        for (int k = 0; k < 3; ++k) { origin[3 * i + k] = S.origins[3 * idx + k]; dir[3 * i + k] = t[3 + k]; }
        weight[i] = 1.f;
        return;
    }
    if (S.type == ERTB_SENSOR_MPDISTANT) {
        const float *M = S.to_world;
        d = normalize3(mk3(M[2], M[5], M[8]));
        fs = mk3(M[0], M[3], M[6]);
        ft = mk3(M[1], M[4], M[7]);
        ax = fx; ay = fy;
    } else if (S.type == ERTB_SENSOR_MDISTANT) {
        int idx = min((int) (fx * (float) S.width), S.width - 1);
        const float *t = S.table + 8 * (size_t) idx;
        d = mk3(t[3], t[4], t[5]);
        onb(d, fs, ft);
    } else {
        f3 hv = uniform_hemisphere(fx, fy);
        const float *M = S.to_world;
        d = mk3(-(M[0] * hv.x + M[1] * hv.y + M[2] * hv.z), -(M[3] * hv.x + M[4] * hv.y + M[5] * hv.z),
                -(M[6] * hv.x + M[7] * hv.y + M[8] * hv.z));
        fs = mk3(M[0], M[3], M[6]);
        ft = mk3(M[1], M[4], M[7]);
        if (S.type == ERTB_SENSOR_DISTANTFLUX) w = hv.z * S.flux_norm;
    }
    double tx, ty, tz;
    if (S.target_type == ERTB_TARGET_POINT) {
        tx = S.target[0]; ty = S.target[1]; tz = S.target[2];
    } else if (S.target_type == ERTB_TARGET_NONE) {
        float ox, oy;
        disk_concentric(ax, ay, ox, oy);
        tx = S.bs_center[0] + ((double) fs.x * ox + (double) ft.x * oy) * S.bs_radius;
        ty = S.bs_center[1] + ((double) fs.y * ox + (double) ft.y * oy) * S.bs_radius;
        tz = S.bs_center[2] + ((double) fs.z * ox + (double) ft.z * oy) * S.bs_radius;
    } else {
        float lx, ly;
        if (S.target_type == ERTB_TARGET_RECTANGLE) { lx = fmaf(2.f, ax, -1.f); ly = fmaf(2.f, ay, -1.f); }
        else disk_concentric(ax, ay, lx, ly);
        const double *T = S.target_to_world;
        tx = T[0] * lx + T[1] * ly + T[3];
        ty = T[4] * lx + T[5] * ly + T[7];
        tz = T[8] * lx + T[9] * ly + T[11];
    }
    origin[3 * i] = tx - (double) d.x * ray_offset;
    origin[3 * i + 1] = ty - (double) d.y * ray_offset;
    origin[3 * i + 2] = tz - (double) d.z * ray_offset;
    dir[3 * i] = d.x; dir[3 * i + 1] = d.y; dir[3 * i + 2] = d.z;
    weight[i] = w;
}

int ertb_kat_sensor_ray(ertb_scene *S, int sensor, size_t n, const float *film_sample,
                        const float *aperture_sample, double *origin, double *dir, float *weight) {
    if (kat_prepare(S)) return 1;
    if (sensor < 0 || sensor >= (int) S->sensors.size()) return set_error("invalid sensor index");
    ErtbParams P = S->base;
    fill_sensor_params(S, S->sensors[sensor], P.sensor);
    DevBuf<float> a, b, w;
    DevBuf<double> o, d;
    if (a.alloc(2 * n) || b.alloc(2 * n) || w.alloc(n) || o.alloc(3 * n) || d.alloc(3 * n))
        return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(a.p, film_sample, 2 * n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(b.p, aperture_sample, 2 * n * sizeof(float), cudaMemcpyHostToDevice));
    kat_sensor_ray_kernel<<<KAT_GRID(n)>>>(P, S->sensors[sensor].ray_offset, n, a.p, b.p, o.p, d.p, w.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(origin, o.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(dir, d.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(weight, w.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}


// ----------------------------------------------------------------------------
// canopy KATs: the BVH ray caster and the leaf BSDF of the 3D kernel, point-wise
// ----------------------------------------------------------------------------
__global__ void kat_canopy_intersect_kernel(ErtbParams P, size_t n, const double *o, const float *dir, const float *tmax,
                                            double *t, float *normal, int *group) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3] = { o[3 * i], o[3 * i + 1], o[3 * i + 2] };
    f3 d = normalize3(mk3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
    CanopyHit H;
    double th = canopy_nearest(P.canopy, p, d, (double) tmax[i], -1, -1, H);
    t[i] = th;
    group[i] = -1;
    normal[3 * i] = normal[3 * i + 1] = normal[3 * i + 2] = 0.f;
    if (H.inst >= 0) {
        float4 in = __ldg(P.canopy.inst + H.inst);
        group[i] = __float_as_int(in.w);
        double q[3] = { p[0] + th * (double) d.x, p[1] + th * (double) d.y, p[2] + th * (double) d.z };
        int kind, mat;
        f3 n = canopy_normal(P.canopy, q, H, kind, mat);
        normal[3 * i] = n.x; normal[3 * i + 1] = n.y; normal[3 * i + 2] = n.z;
    }
}

int ertb_kat_canopy_intersect(ertb_scene *S, size_t n, const double *origin, const float *dir, const float *tmax,
                              double *t, float *normal, int *group) {
    if (kat_prepare(S)) return 1;
    if (S->canopy.n_instances == 0) return set_error("scene has no canopy");
    DevBuf<double> o, tt;
    DevBuf<float> d, tm, nn;
    DevBuf<int> g;
    if (o.alloc(3 * n) || tt.alloc(n) || d.alloc(3 * n) || tm.alloc(n) || nn.alloc(3 * n) || g.alloc(n))
        return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(o.p, origin, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.p, dir, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(tm.p, tmax, n * sizeof(float), cudaMemcpyHostToDevice));
    kat_canopy_intersect_kernel<<<KAT_GRID(n)>>>(S->base, n, o.p, d.p, tm.p, tt.p, nn.p, g.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(t, tt.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(normal, nn.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(group, g.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

__global__ void kat_leaf_bsdf_kernel(float r, float tr, size_t n, const float *ci, const float *co, const float *u,
                                     float *out, float *wo) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (co) { out[i] = bilambertian_eval(r, tr, ci[i], co[i]); return; }
    f3 w;
    out[i] = bilambertian_sample(r, tr, ci[i], u[3 * i], u[3 * i + 1], u[3 * i + 2], w);
    wo[3 * i] = w.x; wo[3 * i + 1] = w.y; wo[3 * i + 2] = w.z;
}

static int kat_leaf_bsdf(ertb_scene *S, int group, size_t n, const float *ci, const float *co, const float *u,
                         float *out, float *wo) {
    if (kat_prepare(S)) return 1;
    if (group < 0 || group >= (int) S->leaf_groups.size()) return set_error("invalid leaf group index");
    DevBuf<float> a, b, c, o, w;
    if (a.alloc(n) || b.alloc(n) || c.alloc(3 * n) || o.alloc(n) || w.alloc(3 * n)) return set_error("cudaMalloc failed");
    CUDA_TRY(cudaMemcpy(a.p, ci, n * sizeof(float), cudaMemcpyHostToDevice));
    if (co) CUDA_TRY(cudaMemcpy(b.p, co, n * sizeof(float), cudaMemcpyHostToDevice));
    if (u) CUDA_TRY(cudaMemcpy(c.p, u, 3 * n * sizeof(float), cudaMemcpyHostToDevice));
    const HostLeafGroup &g = S->leaf_groups[group];
    kat_leaf_bsdf_kernel<<<KAT_GRID(n)>>>(g.reflectance, g.transmittance, n, a.p, co ? b.p : nullptr, c.p, o.p, w.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, o.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (wo) CUDA_TRY(cudaMemcpy(wo, w.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

int ertb_kat_leaf_bsdf_eval(ertb_scene *S, int group, size_t n, const float *cos_i, const float *cos_o, float *out) {
    if (!cos_i || !cos_o || !out) return set_error("null argument");
    return kat_leaf_bsdf(S, group, n, cos_i, cos_o, nullptr, out, nullptr);
}

int ertb_kat_leaf_bsdf_sample(ertb_scene *S, int group, size_t n, const float *cos_i, const float *u, float *wo,
                              float *weight) {
    if (!cos_i || !u || !wo || !weight) return set_error("null argument");
    return kat_leaf_bsdf(S, group, n, cos_i, nullptr, u, weight, wo);
}
