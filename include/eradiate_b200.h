/*
 * eradiate_b200.h -- C ABI of the B200-native Monte Carlo volumetric path
 * tracer that replaces the Mitsuba kernel behind eradiate.kernel.mi_render.
 *
 * Drop-in boundary (reference, read-only):
 *   src/eradiate/kernel/_render.py:186  mi_load_dict   -> ertb_scene_create
 *   src/eradiate/kernel/_render.py:212  mi_traverse    -> (host only; parameter
 *                                                          keys map to ertb_scene_update_*)
 *   src/eradiate/kernel/_render.py:440  parameters.update(...) -> ertb_scene_update_*
 *   src/eradiate/kernel/_render.py:459  mi.render(scene, sensor, seed, spp)
 *                                                      -> ertb_render / ertb_render_device
 *   src/eradiate/kernel/_render.py:466  mi.Bitmap(film.bitmap()) -> the three
 *                                       per-pixel accumulators written by ertb_render
 *
 * The scene arrives as a flat POD descriptor: the Python host
 * (eradiate_b200/kernel/_scene.py) walks the very same nested Mitsuba dict
 * Eradiate builds today and flattens it.  Every field cites the reference
 * plugin parameter it carries.  No torch / C++ types cross this boundary.
 *
 * All functions return 0 on success, non-zero on failure; the message is
 * available from ertb_last_error() (thread-local).  The host shim raises
 * RuntimeError, mirroring src/eradiate/experiments/_core.py:670-671.
 */
#ifndef ERADIATE_B200_H
#define ERADIATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERTB_ABI_VERSION 16
#define ERTB_MAX_PHASE 4       /* leaves of the flattened blendphase tree */
#define ERTB_MAX_BSDF_PARAMS 16
#define ERTB_MAX_LAYERS 4096   /* sigma_t + albedo + weights must fit one SM's shared memory */
#define ERTB_MAX_PHASE_NODES 2048

/* Scene geometry: the three 1D stencils Eradiate emits
 * (src/eradiate/scenes/geometry.py:176-266). */
enum ertb_geometry {
    ERTB_GEOM_PLANE_PARALLEL = 0, /* cube slab + (a)rectangle ground */
    ERTB_GEOM_SPHERICAL_SHELL = 1 /* sphere (TOA, null BSDF) + sphere (ground) */
};

/* ERP/bsdfs + MI/src/bsdfs emitted by eradiate's bsdf_factory
 * (src/eradiate/scenes/bsdfs/_core.py:13-29). */
enum ertb_bsdf_type {
    ERTB_BSDF_DIFFUSE = 0,      /* MI/src/bsdfs/diffuse.cpp:100-178; params: reflectance */
    ERTB_BSDF_RPV = 1,          /* ERP/bsdfs/rpv.cpp:99-193;   params: rho_0, k, g, rho_c */
    ERTB_BSDF_RTLS = 2,         /* ERP/bsdfs/rtls.cpp:86-293;  params: f_iso, f_vol, f_geo, h, r, b */
    ERTB_BSDF_HAPKE = 3,        /* ERP/bsdfs/hapke.cpp:93-381; params: w, b, c, theta(deg), B_0, h */
    ERTB_BSDF_OCEAN_LEGACY = 4, /* ERP/bsdfs/ocean_legacy.cpp; params: wavelength(nm), wind_speed,
                                   wind_direction(deg, North-left), chlorinity, pigmentation,
                                   shadowing(0/1), component (only 0) */
    ERTB_BSDF_BLACK = 5,        /* reflectance 0 (no surface contribution) */
    /* ERP/bsdfs/ocean_mishchenko.cpp:86-346 (sun glint only: isotropic Beckmann facets of the Cox-Munk
     * mean square slope, height-correlated Smith G, Mishchenko-Travis Fresnel matrix);
     * params: wind_speed, eta, k, ext_ior */
    ERTB_BSDF_OCEAN_MISHCHENKO = 6,
    /* ERP/bsdfs/ocean_grasp.cpp:107-539 (Frouin whitecaps + Lambertian water body + glint);
     * params: wavelength(nm), wind_speed, eta, k, ext_ior, water_body_reflectance, component (only 0) */
    ERTB_BSDF_OCEAN_GRASP = 7,
    /* ERP/bsdfs/maignan.cpp:83-254 (Maignan et al. 2009 polarized land reflectance; eval() carries no
     * cosine and sample() returns C*F without dividing by the pdf -- reproduced as is);
     * params: C, ndvi, refr_re, refr_im, ext_ior */
    ERTB_BSDF_MAIGNAN = 8,
    /* ERP/bsdfs/mqdiffuse.cpp:56-210 (measured quasi-diffuse: trilinear, clamped lookup of a
     * (cos_theta_o, phi_d, cos_theta_i) table, cosine-hemisphere sampling); no params, table in
     * `bsdf_table` / `bsdf_table_res` */
    ERTB_BSDF_MQDIFFUSE = 9,
    /* ERP/bsdfs/measured_mono.cpp:40-512 (RGL measured material, Dupuy & Jakob 2018, at one wavelength): five
     * Marginal2D<.., Continuous> interpolants (MI/include/mitsuba/core/distr_2d.h:868-1480) flattened by the host
     * into one float table, `bsdf_table` with `bsdf_table_res` = { number of floats, 1, 1 }; the layout (a 32-float
     * header of sizes / offsets, then the tables; `spectra` already blended at the scene wavelength) is the one
     * eradiate_b200/kernel/_measured.py documents and ertb_scene_create() validates; no params */
    ERTB_BSDF_MEASURED_MONO = 10
};

enum ertb_phase_type {
    ERTB_PHASE_ISOTROPIC = 0,  /* MI/src/phase/isotropic.cpp:39-60 */
    ERTB_PHASE_RAYLEIGH = 1,   /* MI/src/phase/rayleigh.cpp:61-110; params[0] = depolarization */
    ERTB_PHASE_HG = 2,         /* MI/src/phase/hg.cpp:64-100; params[0] = g */
    ERTB_PHASE_TABULATED = 3,  /* MI/src/phase/tabphase.cpp:77-124 (regular cos-theta grid) */
    ERTB_PHASE_TABULATED_IRREGULAR = 4, /* ERP/phase/tabphase_irregular.cpp:111-152 */
    ERTB_PHASE_RAYLEIGH_POLARIZED = 5,  /* ERP/phase/rayleigh_polarized.cpp:55-176; params[0] = depolarization */
    ERTB_PHASE_TABULATED_POLARIZED = 6  /* ERP/phase/tabphase_polarized.cpp:226-430: `values` = m11 on
                                           irregular `nodes`, `mueller` = m12, m22, m33, m34, m44 */
};

enum ertb_sensor_type {
    ERTB_SENSOR_MDISTANT = 0,   /* ERP/sensors/mdistant.cpp:192-242 */
    ERTB_SENSOR_HDISTANT = 1,   /* ERP/sensors/hdistant.cpp:232-275 */
    ERTB_SENSOR_DISTANTFLUX = 2, /* ERP/sensors/distantflux.cpp:148-195 */
    ERTB_SENSOR_PERSPECTIVE = 3, /* MI/src/sensors/perspective.cpp:200-236 (pinhole; canopy scenes) */
    /* ERP/sensors/mpdistant.cpp:214-262: ONE direction (to_world * z), the film sample picks the point of the
     * target the ray goes through (an image of the target region seen from infinity) */
    ERTB_SENSOR_MPDISTANT = 4,
    /* ERP/sensors/mradiancemeter.cpp:147-172: one (origin, direction) per film column; origins may lie inside
     * the atmosphere (`in_medium`), e.g. ground-based sky radiance */
    ERTB_SENSOR_MRADIANCEMETER = 5
};

enum ertb_target_type {
    ERTB_TARGET_NONE = 0,      /* bounding-sphere cross section */
    ERTB_TARGET_POINT = 1,
    ERTB_TARGET_RECTANGLE = 2, /* shape target: rectangle with a to_world */
    ERTB_TARGET_DISK = 3
};

enum ertb_integrator_type {
    ERTB_INTEGRATOR_VOLPATH = 0,   /* MI/src/integrators/volpath.cpp:93-572 */
    ERTB_INTEGRATOR_VOLPATHMIS = 1, /* MI/src/integrators/volpathmis.cpp:124-669 (mono: 1x1 weights) */
    /* ERP/integrators/piecewise_volpath.cpp:91-527 over ERP/media/piecewise.cpp:183-429: analytic
     * free-flight sampling and exact shadow-ray transmittance through the layer stack.  Plane-parallel
     * heterogeneous media only (src/eradiate/experiments/_helpers.py:127-165). */
    ERTB_INTEGRATOR_PIECEWISE_VOLPATH = 2
};

/* One leaf of the (flattened) phase-function tree. */
typedef struct ertb_phase_desc {
    int32_t type;          /* enum ertb_phase_type */
    int32_t n_nodes;       /* tabulated: number of entries of `values` (>= 2) */
    float params[4];       /* see enum */
    const float *values;   /* tabulated: pdf samples, physics convention, cos(theta) ascending */
    const float *nodes;    /* tabulated_irregular / _polarized: cos(theta) nodes in [-1,1]; else NULL */
    const float *mueller[5]; /* tabulated_polarized: m12, m22, m33, m34, m44 (n_nodes each; NULL = 0) */
} ertb_phase_desc;

typedef struct ertb_sensor_desc {
    int32_t type;          /* enum ertb_sensor_type */
    int32_t width, height; /* film size; mdistant: width = n_directions, height = 1 */
    int32_t n_directions;  /* mdistant */
    const double *directions; /* mdistant: 3*n, ray propagation directions (as in the plugin's
                                 `directions` string), need not be normalised */
    double to_world[16];   /* hdistant / distantflux: row-major 4x4 */
    int32_t target_type;   /* enum ertb_target_type */
    int32_t _pad0;
    double target[3];      /* point target */
    double target_to_world[16]; /* rectangle/disk target: maps [-1,1]^2 x {0} (unit disk) to world */
    double ray_offset;     /* < 0: derive from the scene bounding sphere (mdistant.cpp:180-190) */
    /* perspective (to_world = camera-to-world; film width x height): horizontal/vertical field of
     * view resolved to the X axis by the host (perspective.cpp parse_fov), clip planes
     * (near_clip 1e-2, far_clip 1e4 by default), and whether the camera sits inside the
     * atmosphere (the `medium` reference Eradiate adds, experiments/_canopy_atmosphere.py:248-258) */
    double x_fov_deg;
    double near_clip, far_clip;
    int32_t in_medium;     /* perspective / mradiancemeter: the sensor sits inside the atmosphere */
    int32_t _pad1;
    const double *origins; /* mradiancemeter: 3*n ray origins (directions: `directions`, n = width) */
} ertb_sensor_desc;

/* Explicit 3D canopies (src/eradiate/scenes/biosphere/_leaf_cloud.py:1150-1175,
 * _core.py:266-296): `shapegroup`s of `disk` leaves sharing one `bilambertian` BSDF
 * (ERP/bsdfs/bilambertian.cpp:60-215), optionally with a trunk, placed in the scene by `instance`s
 * whose to_world is a pure translation.  Plane-parallel scenes only; the leaves sit inside the atmosphere and do
 * not change the medium of a path (no medium interface).
 * disks: n_disks x 7 floats = centre xyz, unit normal xyz, radius (MI/src/shapes/disk.cpp:
 * to_world = look_at x uniform scale).  A group holds at least one primitive (n_disks may be 0 when it has
 * triangles). */
typedef struct ertb_leaf_group_desc {
    int32_t n_disks;
    float reflectance;     /* bilambertian `reflectance` (uniform) */
    float transmittance;   /* bilambertian `transmittance` (uniform) */
    int32_t _pad;
    const float *disks;
    /* AbstractTree (src/eradiate/scenes/biosphere/_tree.py:150-180): the group may also hold a trunk, i.e.
     * `cylinder`s (MI/src/shapes/cylinder.cpp:560-615, open tubes) and cap `disk`s sharing ONE one-sided
     * `diffuse` BSDF (MI/src/bsdfs/diffuse.cpp:100-178). */
    int32_t n_cylinders;
    int32_t n_trunk_disks;
    const float *cylinders;     /* n_cylinders x 7: p0 xyz, p1 xyz, radius */
    const float *trunk_disks;   /* n_trunk_disks x 7, same layout as `disks` */
    float trunk_reflectance;
    /* MeshTreeElement (src/eradiate/scenes/biosphere/_tree.py:285-470): `ply` / `obj` triangle meshes
     * (MI/src/shapes/ply.cpp, obj.cpp, MI/src/render/mesh.cpp), each with its own `bilambertian` BSDF.  The host
     * reads the files, applies `to_world` and computes the vertex normals (eradiate_b200/kernel/_mesh.py); here
     * they are plain triangles: 18 floats = v0, v1, v2, then the unit shading normals at the three vertices
     * (mesh.cpp:1500-1560: interpolated and renormalised at the hit; all three equal to the face normal when the
     * shape was loaded with `face_normals`).  Hit test: Moeller-Trumbore, MI/include/mitsuba/render/mesh.h:481-504. */
    int32_t n_triangles;
    int32_t n_mesh_bsdfs;
    int32_t _pad2;
    const float *triangles;        /* n_triangles x 18 */
    const int32_t *triangle_bsdf;  /* n_triangles: index into mesh_bsdfs */
    const float *mesh_bsdfs;       /* n_mesh_bsdfs x 2: bilambertian reflectance, transmittance */
} ertb_leaf_group_desc;

typedef struct ertb_scene_desc {
    int32_t abi_version;   /* must be ERTB_ABI_VERSION */
    int32_t geometry;      /* enum ertb_geometry */

    /* Geometry, metres (kernel length unit, src/eradiate/units.py).
     * plane-parallel : ground plane z = surface_z, slab top z = medium_top;
     *                  the volume grid spans [medium_bottom, medium_top].
     * spherical shell: ground sphere radius = surface_z (planet radius + ground
     *                  altitude), TOA sphere radius = medium_top; the radial grid spans
     *                  [medium_bottom, medium_top] (medium_bottom = rmin * medium_top,
     *                  ERP/volumes/sphericalcoords.cpp:102-123). */
    double surface_z;
    double medium_bottom;
    double medium_top;
    double bsphere_center[3]; /* scene bounding sphere (sensor ray_offset, emitter distance) */
    double bsphere_radius;

    /* Medium: MI/src/media/heterogeneous.cpp:155-201 (homogeneous.cpp = 1 layer). */
    int32_t has_medium;
    int32_t n_layers;
    const float *sigma_t;  /* n_layers, m^-1, float32 exactly as the reference stores it */
    const float *albedo;   /* n_layers */
    float sigma_t_scale;   /* heterogeneous `scale` */
    int32_t homogeneous;   /* 1: homogeneous.cpp semantics (majorant = sigma_t, no null collisions) */

    /* Phase function: flattened blendphase tree (MI/src/phase/blendphase.cpp:100-190).
     * phase_weight[i*n_layers + l] = probability of leaf i in layer l (rows sum to 1);
     * NULL when n_phase == 1. */
    int32_t n_phase;
    int32_t _pad1;
    ertb_phase_desc phase[ERTB_MAX_PHASE];
    const float *phase_weight;

    /* Surface */
    int32_t bsdf_type;     /* enum ertb_bsdf_type */
    int32_t _pad2;
    float bsdf_params[ERTB_MAX_BSDF_PARAMS];

    /* Emitter: MI/src/emitters/directional.cpp:171-201 */
    double emitter_direction[3]; /* direction light travels (to_world * (0,0,1)) */
    float irradiance;
    int32_t _pad3;

    /* Integrator: MI/src/render/integrator.cpp:562-566 defaults */
    int32_t integrator;    /* enum ertb_integrator_type */
    int32_t rr_depth;      /* default 5 */
    int64_t max_depth;     /* -1 = unbounded */
    /* Polarized (Stokes / Mueller) transport: the scalar_mono_polarized variant of the reference.
     * MI/src/integrators/stokes.cpp:97-168 rotates the Stokes vector to the meridian plane
     * (`meridian_align`) or to the sensor's x-axis before it is written to S0..S3. */
    int32_t polarized;
    int32_t meridian_align;

    int32_t n_sensors;
    int32_t _pad4;
    const ertb_sensor_desc *sensors;

    /* Canopy (all zero / NULL for 1D scenes) */
    int32_t n_leaf_groups;
    int32_t n_instances;
    const ertb_leaf_group_desc *leaf_groups;
    const int32_t *instance_group;  /* n_instances: index into leaf_groups */
    const double *instance_offset;  /* n_instances x 3: translation of the instance */

    /* CentralPatchSurface (src/eradiate/scenes/surface/_central_patch.py:185-215): a `blendbsdf` whose
     * weight is the 3x3 central-patch mask, i.e. `patch_bsdf` inside the axis-aligned rectangle
     * [cx - hx, cx + hx] x [cy - hy, cy + hy] of the ground plane and the scene BSDF (bsdf_type /
     * bsdf_params = `bsdf_0`, the background) outside. Plane-parallel scenes; land BSDFs only. */
    int32_t has_patch;
    int32_t patch_bsdf_type;        /* enum ertb_bsdf_type; not the ocean */
    float patch_bsdf_params[ERTB_MAX_BSDF_PARAMS];
    double patch_rect[4];           /* cx, cy, hx, hy (metres) */
    /* mqdiffuse: the plugin's VolumeGrid, data[z][y][x] with x = cos_theta_o, y = phi_d / 2 pi, z = cos_theta_i
     * (mqdiffuse.cpp:80-86); copied at scene creation, not updatable (the plugin exposes no parameter) */
    const float *bsdf_table;
    int32_t bsdf_table_res[3];      /* x, y, z; measured_mono: { number of floats, 1, 1 } */
    /* ERP/phase/multiphase.cpp:176-200 (`use_mis`, root node only): the weight of a sampled direction is the mixture
     * sum_j w_j value_j / sum_j w_j pdf_j over all leaves instead of the drawn leaf's own weight.  Only set when
     * that differs (a leaf whose value is not its pdf: depolarized Rayleigh, Mueller-valued phase functions). */
    int32_t phase_mis;
    /* ERP/emitters/astroobject.cpp:54-242: the light source is a uniform disc of this angular diameter (degrees,
     * in ]0, 180[) centred on -emitter_direction, radiating `irradiance` / solid angle; 0 = the delta
     * `directional` emitter.  1D scenes (no canopy / camera / central patch). */
    double emitter_angular_diameter;
    /* MI/src/render/integrator.cpp:29 `hide_emitters`: camera rays do not see the emitter directly (volpath.cpp:103,
     * :114, :329-330: no emitter hit at depth 0, and hits of rays that have not scattered yet are no longer counted
     * in full).  Only the astroobject disc can be seen at all, so the flag matters with it alone. */
    int32_t hide_emitters;
    int32_t _pad_tail;
} ertb_scene_desc;

/* Named updatable parameters (KernelSceneParameterMap keys resolve to these;
 * src/eradiate/scenes/atmosphere/_core.py:777-805, phase/_tabulated.py:273-281,
 * phase/_blend.py:282-305, bsdfs/_rpv.py, illumination/_directional.py). */
enum ertb_param {
    ERTB_PARAM_SIGMA_T = 0,      /* float[n_layers] */
    ERTB_PARAM_ALBEDO = 1,       /* float[n_layers] */
    ERTB_PARAM_PHASE_WEIGHT = 2, /* float[n_phase*n_layers] (leaf probabilities) */
    ERTB_PARAM_PHASE_VALUES = 3, /* index = leaf; float[n_nodes] */
    ERTB_PARAM_BSDF_PARAMS = 4,  /* float[ERTB_MAX_BSDF_PARAMS] */
    ERTB_PARAM_IRRADIANCE = 5,   /* float[1] */
    ERTB_PARAM_PHASE_PARAMS = 6, /* index = leaf; float[4] */
    ERTB_PARAM_PHASE_MUELLER = 7, /* index = leaf * 5 + k (k: m12, m22, m33, m34, m44); float[n_nodes] */
    ERTB_PARAM_LEAF_BSDF = 8,    /* index = leaf group; float[2]: reflectance, transmittance */
    ERTB_PARAM_PATCH_BSDF_PARAMS = 9, /* float[ERTB_MAX_BSDF_PARAMS]: the central patch's BSDF */
    ERTB_PARAM_TRUNK_BSDF = 10,       /* index = leaf group; float[1]: trunk reflectance */
    ERTB_PARAM_MESH_BSDF = 11         /* index = (leaf group << 16) | mesh BSDF; float[2]: reflectance, transmittance */
};

typedef struct ertb_render_stats {
    uint64_t n_paths;        /* samples traced by this call */
    uint64_t trips_main;     /* iterations of hot loop #2 (volpath.cpp:170-393) */
    uint64_t trips_nee;      /* iterations of hot loop #3 (volpath.cpp:454-551) */
    uint64_t n_scatter;      /* real medium collisions */
    uint64_t n_surface;      /* surface (non-null) interactions */
    double device_ms;        /* kernel time measured with CUDA events on the launch stream */
    int32_t n_launches;      /* kernels launched by this call */
    int32_t n_bands;         /* altitude bands of the majorant (1 = the reference's global majorant) */
} ertb_render_stats;

typedef struct ertb_scene ertb_scene; /* opaque */

/* Library / device probing.  ertb_device_count returns -1 (and sets the error)
 * when no CUDA device is usable: there is NO CPU fallback. */
int ertb_abi_version(void);
const char *ertb_last_error(void);
int ertb_device_count(void);

/* mi_load_dict equivalent: validates the descriptor, uploads tables to `device`. */
int ertb_scene_create(const ertb_scene_desc *desc, int device, ertb_scene **out);
void ertb_scene_destroy(ertb_scene *scene);

/* parameters.update(...) equivalent. */
int ertb_scene_update(ertb_scene *scene, int param, int index, const float *data, size_t count);

/* mi.render equivalent.  Renders samples [sample_offset, sample_offset+spp) of
 * every pixel of `sensor` (sample sharding across GPUs = disjoint offsets with
 * the same seed).  Outputs (host, n_pixels doubles each, row-major film):
 *   sum_wl : sum of ray_weight * L   (root bitmap * spp, integrator.cpp:489)
 *   sum_l  : sum of L                (moment.cpp:101, `nested`)
 *   sum_l2 : sum of L^2              (moment.cpp:104, `m2_nested`)
 * Any output pointer may be NULL. */
int ertb_render(ertb_scene *scene, int sensor, uint64_t seed, uint64_t spp,
                uint64_t sample_offset, double *sum_wl, double *sum_l, double *sum_l2,
                ertb_render_stats *stats);

/* Same, but accumulates into a caller-owned DEVICE buffer of 3*n_pixels doubles
 * ([sum_wl | sum_l | sum_l2], zeroed by the caller) on `stream` (a cudaStream_t,
 * NULL = default stream) and does not synchronise: used for the HBM-resident
 * throughput number and for the NCCL all-reduce of the accumulators.
 * `stats_dev` (optional) is a device buffer of 8 uint64 counters.
 * Ordering: outside the ertb_batch_* entry points a scene owns ONE table slot and ONE work counter, so its
 * renders are serialised by the library -- a launch on another stream than the previous one first waits for it,
 * and a launch that has to commit updated parameters first waits for `stream` (the upload is synchronous).
 * Any stream is therefore safe, including non-blocking ones; overlap of renders comes from ertb_batch_push. */
int ertb_render_device(ertb_scene *scene, int sensor, uint64_t seed, uint64_t spp,
                       uint64_t sample_offset, void *accum_dev, void *stats_dev,
                       void *stream);

/* Polarized scenes: same as ertb_render plus sum_stokes = [4][n_pixels] sums of
 * ray_weight * (S0, S1, S2, S3) in the output frame of the stokes integrator (may be NULL).
 * The device variant expects accum_dev to hold (3 + 4) * n_pixels doubles for polarized scenes. */
int ertb_render_stokes(ertb_scene *scene, int sensor, uint64_t seed, uint64_t spp,
                       uint64_t sample_offset, double *sum_wl, double *sum_l, double *sum_l2,
                       double *sum_stokes, ertb_render_stats *stats);

int ertb_sensor_pixel_count(const ertb_scene *scene, int sensor);

/* Pipelined version of the contexts x sensors loop of mi_render
 * (src/eradiate/kernel/_render.py:433-468).  The reference updates the scene
 * parameters, renders and copies the film back strictly one (context, sensor) at a time.
 * Here the caller announces the items of the loop, then for each item applies the parameter
 * updates (ertb_scene_update) and calls ertb_batch_push, which snapshots the current
 * parameters into private device tables (pinned staging buffer + async copy) and launches
 * the render on one of ERTB_BATCH_SLOTS streams WITHOUT synchronising: the host prepares
 * context i+1 while context i renders, and the tail of one render overlaps the head of the
 * next.  ertb_batch_end waits for everything and returns all accumulators with one copy:
 * item i occupies rows * n_pixels(sensors[i]) doubles ([sum_wl | sum_l | sum_l2 (| S0..S3)],
 * rows = 3, or 7 for polarized scenes), items concatenated in push order.  Results are the
 * same as n_items ertb_render calls with the same arguments.
 *   stats      : n_items entries or NULL (only filled when with_stats != 0; device_ms is 0)
 *   elapsed_ms : host wall-clock time from ertb_batch_begin to the end of the last render */
int ertb_batch_begin(ertb_scene *scene, int n_items, const int *sensors, int with_stats);
int ertb_batch_push(ertb_scene *scene, int sensor, uint64_t seed, uint64_t spp,
                    uint64_t sample_offset);
int ertb_batch_end(ertb_scene *scene, double *accum_out, size_t count,
                   ertb_render_stats *stats, double *elapsed_ms);

/* Known-answer-test entry points: evaluate the device implementations of the
 * plugins point-wise (one thread per query).  Host pointers.
 *   bsdf_eval  : wi, wo local-frame unit vectors (3*n); out = f * cos(theta_o) (n)
 *   bsdf_sample: wi (3*n), u (3*n: lobe-selection sample1, then sample2.x, sample2.y)
 *                -> wo (3*n), weight (n)
 *   phase_eval : leaf index, cos between wo and wi ("graphics" convention, n) -> out (n)
 *   phase_sample: leaf index, u (2*n) -> cos_theta of wo w.r.t. propagation dir (n), weight (n), pdf (n)
 *   sensor_ray : film sample (2*n) + aperture sample (2*n) -> origin (3*n) dir (3*n) weight (n)
 */
int ertb_kat_bsdf_eval(ertb_scene *scene, size_t n, const float *wi, const float *wo, float *out);
int ertb_kat_bsdf_sample(ertb_scene *scene, size_t n, const float *wi, const float *u,
                         float *wo, float *weight);
int ertb_kat_phase_eval(ertb_scene *scene, int leaf, size_t n, const float *cos_theta, float *out);
int ertb_kat_phase_sample(ertb_scene *scene, int leaf, size_t n, const float *u,
                          float *cos_theta, float *weight, float *pdf);
/* piecewise medium (ERP/media/piecewise.cpp:183-332 sample_interaction_real, :335-429
 * eval_transmittance_pdf_real; golden vectors in ERP/tests/media/test_piecewise.py), in the
 * kernel's reduced coordinates: altitude above the ground (n), vertical direction cosine (n).
 *   piecewise_sample: + uniform sample (n) -> flight distance (n), kind (n): 0 = collision at
 *                     that distance, 1 = reached the ground at that distance, 2 = left through the top
 *   piecewise_transmittance: -> exp(-optical depth) from the altitude to the top along mu > 0 */
int ertb_kat_piecewise_sample(ertb_scene *scene, size_t n, const float *altitude, const float *mu,
                              const float *u, float *distance, int32_t *kind);
int ertb_kat_piecewise_transmittance(ertb_scene *scene, size_t n, const float *altitude, const float *mu,
                                     float *transmittance);
/* phase_mueller: leaf, wi (3*n, = -propagation direction), wo (3*n) -> 4x4 Mueller matrices in the
 * implicit Stokes bases of the two directions (16*n, row-major) and pdf (n) */
int ertb_kat_phase_mueller(ertb_scene *scene, int leaf, size_t n, const float *wi, const float *wo,
                           float *mueller, float *pdf);
/* bsdf_mueller (polarized scenes): wi = si.wi, wo (3*n, local frame, z = normal) -> BSDF::eval as a 4x4
 * Mueller matrix in the implicit Stokes bases of -wo and wi (16*n, row-major), i.e. what the reference's
 * plugin tests read (ERP/tests/bsdfs/test_ocean_mishchenko.py:44-109, test_maignan.py:28-88) */
int ertb_kat_bsdf_mueller(ertb_scene *scene, size_t n, const float *wi, const float *wo, float *mueller);
int ertb_kat_sensor_ray(ertb_scene *scene, int sensor, size_t n, const float *film_sample,
                        const float *aperture_sample, double *origin, double *dir, float *weight);

/* Canopy KATs (3D kernel): the BVH ray caster and the leaf BSDF, point-wise. Host pointers.
 *   canopy_intersect: world-space origins (3*n doubles), directions (3*n floats, normalised on the
 *                     device), tmax (n) -> distance t (inf = miss), leaf normal (3*n), leaf group (n; -1)
 *   leaf_bsdf_eval  : bilambertian of `group`: cos of wi / wo with the leaf normal -> f * |cos_o|
 *   leaf_bsdf_sample: cos_i (n), u (3*n: lobe selection, then the 2D sample) -> local wo (3*n), weight */
int ertb_kat_canopy_intersect(ertb_scene *scene, size_t n, const double *origin, const float *dir,
                              const float *tmax, double *t, float *normal, int *group);
int ertb_kat_leaf_bsdf_eval(ertb_scene *scene, int group, size_t n, const float *cos_i,
                            const float *cos_o, float *out);
int ertb_kat_leaf_bsdf_sample(ertb_scene *scene, int group, size_t n, const float *cos_i,
                              const float *u, float *wo, float *weight);

#ifdef __cplusplus
}
#endif
#endif /* ERADIATE_B200_H */
