/*
 * ertb_oracle_canopy.h -- explicit 3D canopies for the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Disk leaves in instanced shape groups (src/eradiate/scenes/biosphere/_leaf_cloud.py:1150-1175,
 * _core.py:266-296), intersected as MI/src/shapes/disk.cpp:388-407 does, with the bilambertian
 * leaf BSDF of ERP/bsdfs/bilambertian.cpp:60-215.  The acceleration structure is a uniform 3D
 * grid walked with a DDA -- deliberately NOT the BVH of the CUDA library, so that the GPU/oracle
 * parity tests cross-check two independent ray casters.
 */
#ifndef ERTB_ORACLE_CANOPY_H
#define ERTB_ORACLE_CANOPY_H

#include "../include/eradiate_b200.h"

/* primitive kinds of a group: leaf disks first, then the trunk's cap disks, then its cylinders */
enum { CANOPY_LEAF = 0, CANOPY_TRUNK = 1, CANOPY_MESH = 2 };

typedef struct {
    int n_disks;         /* leaves + trunk disks */
    int n_leaf_disks;
    int n_cylinders;
    float *disks;        /* n_disks x 7 (owned copy: leaves, then trunk disks) */
    const float *cylinders; /* n_cylinders x 7: p0, p1, radius (borrowed) */
    int n_triangles;        /* mesh elements (MI/src/render/mesh.cpp): primitives after the cylinders */
    const float *triangles; /* n_triangles x 18: v0, v1, v2, shading normals n0, n1, n2 (borrowed) */
    const int *triangle_bsdf;
    const float *mesh_bsdfs; /* n x 2: bilambertian reflectance, transmittance */
    double trunk_reflectance;
    double lo[3], hi[3]; /* bounding box of the group (local coordinates) */
    int res[3];
    double cell[3];
    int *cell_start;     /* res[0]*res[1]*res[2] + 1 */
    int *cell_items;
    double reflectance, transmittance;
} canopy_group_t;

typedef struct {
    int n_groups, n_instances;
    canopy_group_t *groups;
    const int *instance_group;
    const double *instance_offset;
} canopy_t;

typedef struct {
    double t;        /* INFINITY: no hit */
    double p[3];     /* hit point re-projected onto the disk (disk.cpp:482-485) */
    double n[3];     /* disk normal (m_frame.n) / outward normal of the cylinder / face normal of the triangle */
    double sh_n[3];  /* shading normal: n, except on a mesh with vertex normals (interpolated, mesh.cpp:1500-1535) */
    int group;
    int kind;        /* CANOPY_LEAF (bilambertian), CANOPY_TRUNK (one-sided diffuse), CANOPY_MESH (bilambertian r, t below) */
    double mesh_r, mesh_t;
} canopy_hit_t;

int canopy_init(canopy_t *C, const ertb_scene_desc *d);
void canopy_free(canopy_t *C);
/* nearest leaf along o + t d, 0 <= t <= maxt */
canopy_hit_t canopy_intersect(const canopy_t *C, const double o[3], const double d[3], double maxt);

/* bilambertian.cpp: local frame, z = leaf normal.  eval returns f * |cos(theta_o)|. */
double bilambertian_eval(double r, double t, const double wi[3], const double wo[3]);
double bilambertian_pdf(double r, double t, const double wi[3], const double wo[3]);
/* returns the weight (value / pdf) and the sampled direction */
double bilambertian_sample(double r, double t, const double wi[3], double sample1, double u1, double u2,
                           double wo[3]);

#endif
