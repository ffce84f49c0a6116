// grid_bodies.h -- per-index bodies of the grid-build kernels (grid_build.cu): every geometric
// quantity of the reference GridBlock, with the reference's formulas and operation order so that
// the arrays are bit-identical to the ones GridBlock<T>::init_grid_block produces
// (/root/reference/src/grid/grid/grid.h:89-167).  Shared by the CUDA kernels and by the host
// emulation harness (tests/emu).
#pragma once

#include <stdint.h>

#include "scalar.h"

namespace ib {

// VTK / SU2 element codes (grid/grid_io.cpp: elem_type_from_vtk_type)
enum { VTK_TRI = 5, VTK_QUAD = 9, VTK_TETRA = 10, VTK_HEX = 12, VTK_WEDGE = 13, VTK_PYRAMID = 14 };

constexpr int64_t GB_NONE = INT64_MAX;   // "no cell attached yet" in face_left / face_right

struct GridBuild {
    int dim;
    int64_t n_vertices, n_cells, n_ghost, n_faces, n_records, n_bfaces;
    // inputs
    const double* pos;             // n_vertices x 3
    const int32_t* ctype;          // n_cells
    const int64_t* cverts;         // n_cells x 8, -1 padded
    const int64_t* fverts;         // n_faces x 4, -1 padded, first discoverer's order
    const int64_t* cf_off;         // n_cells + 1
    const int64_t* cf_face;        // n_records: face of every (cell, local face) record
    const int64_t* bf_face;        // n_bfaces boundary faces in (tag, marker) order
    // outputs
    double* f_centre;              // n_faces x 3
    double* f_area;
    double* f_tan1;
    double* f_tan2;
    double* f_norm;
    double* c_centroid;            // (n_cells + n_ghost) x 3
    double* c_volume;              // n_cells
    int64_t* f_left;
    int64_t* f_right;
    int32_t* cf_out;               // n_records
    int64_t* cf_nbr;               // n_records
};

IBIS_HD int gb_num_vertices(int t) {
    return t == VTK_TRI ? 3 : t == VTK_QUAD ? 4 : t == VTK_TETRA ? 4 : t == VTK_HEX ? 8 : t == VTK_WEDGE ? 6
         : t == VTK_PYRAMID ? 5 : 0;
}
IBIS_HD int gb_num_faces(int t) {
    return t == VTK_TRI ? 3 : t == VTK_QUAD ? 4 : t == VTK_TETRA ? 4 : t == VTK_HEX ? 6 : t == VTK_WEDGE ? 5
         : t == VTK_PYRAMID ? 5 : 0;
}

// area of the triangle (a, b, c): half the norm of the cross product (grid/geom.h:30-45)
IBIS_HD double gb_tri_area(const double* pos, int64_t a, int64_t b, int64_t c) {
    const double abx = pos[3 * b] - pos[3 * a], aby = pos[3 * b + 1] - pos[3 * a + 1], abz = pos[3 * b + 2] - pos[3 * a + 2];
    const double acx = pos[3 * c] - pos[3 * a], acy = pos[3 * c + 1] - pos[3 * a + 1], acz = pos[3 * c + 2] - pos[3 * a + 2];
    const double cx = aby * acz - abz * acy;
    const double cy = abz * acx - abx * acz;
    const double cz = abx * acy - aby * acx;
    return 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
}

// centre = vertex average, area by triangles (geom.h:30-62), tangents and normal from the first
// discoverer's vertex order (interface.h:187-250)
IBIS_HD void gb_face_geometry(const GridBuild& G, int64_t f) {
    const int64_t* v = G.fverts + 4 * f;
    const int nv = v[3] >= 0 ? 4 : v[2] >= 0 ? 3 : 2;
    const double* pos = G.pos;
    for (int d = 0; d < 3; d++) {
        double acc = 0.0;
        for (int k = 0; k < nv; k++) acc += pos[3 * v[k] + d];
        G.f_centre[3 * f + d] = acc / (double)nv;
    }
    const double d0 = pos[3 * v[1]] - pos[3 * v[0]], d1 = pos[3 * v[1] + 1] - pos[3 * v[0] + 1],
                 d2 = pos[3 * v[1] + 2] - pos[3 * v[0] + 2];
    const double len2 = d0 * d0 + d1 * d1 + d2 * d2;
    if (nv == 2) G.f_area[f] = sqrt(len2);
    else if (nv == 3) G.f_area[f] = gb_tri_area(pos, v[0], v[1], v[2]);
    else G.f_area[f] = gb_tri_area(pos, v[0], v[1], v[2]) + gb_tri_area(pos, v[0], v[2], v[3]);
    const double il = 1.0 / sqrt(len2);
    const double t1x = il * d0, t1y = il * d1, t1z = il * d2;
    double t2x = 0.0, t2y = 0.0, t2z = 1.0;
    if (nv > 2) {
        const double sx = pos[3 * v[2]] - pos[3 * v[0]], sy = pos[3 * v[2] + 1] - pos[3 * v[0] + 1],
                     sz = pos[3 * v[2] + 2] - pos[3 * v[0] + 2];
        const double proj = t1x * sx + t1y * sy + t1z * sz;
        const double ux = sx - proj * t1x, uy = sy - proj * t1y, uz = sz - proj * t1z;
        const double il2 = 1.0 / sqrt(ux * ux + uy * uy + uz * uz);
        t2x = il2 * ux;
        t2y = il2 * uy;
        t2z = il2 * uz;
    }
    G.f_tan1[3 * f] = t1x; G.f_tan1[3 * f + 1] = t1y; G.f_tan1[3 * f + 2] = t1z;
    G.f_tan2[3 * f] = t2x; G.f_tan2[3 * f + 1] = t2y; G.f_tan2[3 * f + 2] = t2z;
    G.f_norm[3 * f] = t1y * t2z - t1z * t2y;
    G.f_norm[3 * f + 1] = t1z * t2x - t1x * t2z;
    G.f_norm[3 * f + 2] = t1x * t2y - t1y * t2x;
}

// centroid of a valid cell = vertex average (cell.h:236-254)
IBIS_HD void gb_cell_centroid(const GridBuild& G, int64_t c) {
    const int nv = gb_num_vertices(G.ctype[c]);
    const int64_t* v = G.cverts + 8 * c;
    for (int d = 0; d < 3; d++) {
        double acc = 0.0;
        for (int k = 0; k < nv; k++) acc += G.pos[3 * v[k] + d];
        G.c_centroid[3 * c + d] = acc / (double)nv;
    }
}

// left / right attachment and outsigns of one cell: the cell is the LEFT cell of a face when
// (x_f - x_c) . n > 0 (grid.h:186-213).  A conforming face has one cell on each side, so the
// stores of different cells never collide.
IBIS_HD void gb_attach(const GridBuild& G, int64_t c) {
    for (int64_t r = G.cf_off[c]; r < G.cf_off[c + 1]; r++) {
        const int64_t f = G.cf_face[r];
        const double dx = G.f_centre[3 * f] - G.c_centroid[3 * c], dy = G.f_centre[3 * f + 1] - G.c_centroid[3 * c + 1],
                     dz = G.f_centre[3 * f + 2] - G.c_centroid[3 * c + 2];
        const double dot = dx * G.f_norm[3 * f] + dy * G.f_norm[3 * f + 1] + dz * G.f_norm[3 * f + 2];
        if (dot > 0.0) {
            G.cf_out[r] = 1;
            G.f_left[f] = c;
        } else {
            G.cf_out[r] = -1;
            G.f_right[f] = c;
        }
    }
}

// ghost cell of boundary face k: id n_cells + k, on the empty side (grid.h:414-433); its centroid is
// the valid cell's mirrored through the face centre (grid.h:364-403)
IBIS_HD void gb_ghost(const GridBuild& G, int64_t k) {
    const int64_t f = G.bf_face[k], ghost = G.n_cells + k;
    int64_t valid;
    if (G.f_left[f] == GB_NONE) {
        G.f_left[f] = ghost;
        valid = G.f_right[f];
    } else {
        G.f_right[f] = ghost;
        valid = G.f_left[f];
    }
    if (valid == GB_NONE || valid >= G.n_cells) return;     // reported by the host check
    for (int d = 0; d < 3; d++) {
        const double dxyz = G.f_centre[3 * f + d] - G.c_centroid[3 * valid + d];
        G.c_centroid[3 * ghost + d] = G.f_centre[3 * f + d] + dxyz;
    }
}

// volume (2-D: area by triangles, cell.h:263-280; 3-D: divergence theorem over the faces in
// cell-face order, cell.h:282-303, geom.h:67-83) and neighbours (cell.h:327) of one cell
IBIS_HD void gb_volume_neighbours(const GridBuild& G, int64_t c) {
    const int64_t r0 = G.cf_off[c], r1 = G.cf_off[c + 1];
    for (int64_t r = r0; r < r1; r++) {
        const int64_t f = G.cf_face[r];
        G.cf_nbr[r] = G.f_left[f] == c ? G.f_right[f] : G.f_left[f];
    }
    if (G.dim == 2) {
        const int64_t* v = G.cverts + 8 * c;
        G.c_volume[c] = G.ctype[c] == VTK_TRI ? gb_tri_area(G.pos, v[0], v[1], v[2])
                                              : gb_tri_area(G.pos, v[0], v[1], v[2]) + gb_tri_area(G.pos, v[0], v[2], v[3]);
        return;
    }
    double acc = 0.0;
    for (int64_t r = r0; r < r1; r++) {
        const int64_t f = G.cf_face[r];
        const double out = (double)G.cf_out[r];
        const double dotv = G.f_centre[3 * f] * out * G.f_norm[3 * f] + G.f_centre[3 * f + 1] * out * G.f_norm[3 * f + 1] +
                            G.f_centre[3 * f + 2] * out * G.f_norm[3 * f + 2];
        acc = acc + dotv * G.f_area[f];
    }
    G.c_volume[c] = acc / 3.0;
}

}  // namespace ib
