// box_grid.cpp -- closed-form construction of the GridBlock arrays of an nx x ny x nz hex box
// "in unstructured form" (BASELINE.json configs[4]), optionally of one
// z-slab of it with a two-ring halo for the domain-decomposed runs.
//
// It produces exactly what GridBlock<T>::init_grid_block (/root/reference/src/grid/grid/grid.h:
// 89-167) produces for the SU2 file of that box -- same face numbering (order of first discovery
// looping cells in order and faces in VTK order, grid.h:113-133, grid_io.cpp:298-348), same
// orientation (the first discoverer's vertex order, interface.h:187-250), same ghost numbering
// (alphabetical tags, marker order, grid.h:414-433) and the same floating-point formulas in the
// same order (geom.h:30-83, cell.h:236-303, grid.h:186-213,364-403) -- without the reference's
// string-keyed hash map (17 us per cell; 64 M cells would take ~20 min and > 100 GB).  The
// general-purpose mirror of the same algorithm is ibis_b200/grid.py; tests/test_box_grid.py checks
// the two (and through them the reference itself) agree bit for bit.
//
// Built with -ffp-contract=off: every operation is an individually rounded IEEE operation.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "ibis_b200_host.h"

namespace {

struct V3 {
    double x, y, z;
};
inline V3 sub(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }

// geom.h:30-62 through ibis_b200/grid.py:_tri_area
inline double tri_area(const V3& a, const V3& b, const V3& c) {
    const V3 ab = sub(b, a), ac = sub(c, a);
    const double cx = ab.y * ac.z - ab.z * ac.y;
    const double cy = ab.z * ac.x - ab.x * ac.z;
    const double cz = ab.x * ac.y - ab.y * ac.x;
    return 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
}

struct FaceGeom {
    double area;
    V3 n, t1, t2, c;
};

// one quad interface from its four vertices in the first discoverer's order
inline FaceGeom quad_geom(const V3 p[4]) {
    FaceGeom g;
    g.c.x = (((0.0 + p[0].x) + p[1].x) + p[2].x + p[3].x) / 4.0;
    g.c.y = (((0.0 + p[0].y) + p[1].y) + p[2].y + p[3].y) / 4.0;
    g.c.z = (((0.0 + p[0].z) + p[1].z) + p[2].z + p[3].z) / 4.0;
    g.area = tri_area(p[0], p[1], p[2]) + tri_area(p[0], p[2], p[3]);
    const V3 d01 = sub(p[1], p[0]);
    const double il = 1.0 / sqrt(d01.x * d01.x + d01.y * d01.y + d01.z * d01.z);
    g.t1 = {il * d01.x, il * d01.y, il * d01.z};
    const V3 t2s = sub(p[2], p[0]);
    const double proj = g.t1.x * t2s.x + g.t1.y * t2s.y + g.t1.z * t2s.z;
    const V3 t2 = {t2s.x - proj * g.t1.x, t2s.y - proj * g.t1.y, t2s.z - proj * g.t1.z};
    const double il2 = 1.0 / sqrt(t2.x * t2.x + t2.y * t2.y + t2.z * t2.z);
    g.t2 = {il2 * t2.x, il2 * t2.y, il2 * t2.z};
    g.n = {g.t1.y * g.t2.z - g.t1.z * g.t2.y, g.t1.z * g.t2.x - g.t1.x * g.t2.z,
           g.t1.x * g.t2.y - g.t1.y * g.t2.x};
    return g;
}

}  // namespace

struct ibis_b200_box {
    int nx, ny, nz;          // global
    double lx, ly, lz;
    int k_lo, k_hi;          // owned slab [k_lo, k_hi)
    ibis_b200_grid_desc desc;
    std::vector<int64_t> face_left, face_right, cf_off, cf_ids, cf_nbr, tag_off, tag_faces;
    std::vector<int32_t> cf_out;
    std::vector<double> f_area, f_norm, f_t1, f_t2, f_c, c_vol, c_cen;
    // halo exchange lists, peers: lower (rank-1) then upper (rank+1) when present
    int n_peers = 0;
    int32_t peer_dir[2] = {0, 0};
    std::vector<int64_t> send_off, send_cells, recv_off, recv_cells;
    // local cell id of global (i, j, k); -1 when not held
    std::vector<int> kslot;  // global k -> local layer slot, -1 if absent
    int n_layers = 0;
    int64_t cell_id(int i, int j, int k) const {
        const int s = kslot[k];
        return s < 0 ? -1 : (int64_t)i + (int64_t)nx * ((int64_t)j + (int64_t)ny * s);
    }
    V3 vertex(int i, int j, int k) const {            // su2.box_hex: (i * lx) / nx
        return {((double)i * lx) / (double)nx, ((double)j * ly) / (double)ny, ((double)k * lz) / (double)nz};
    }
};

namespace {

thread_local std::string g_box_error;

// VTK hex local vertex (grid_io.cpp / su2.box_hex): 0 (i,j,k) 1 (i+1,j,k) 2 (i+1,j+1,k) 3 (i,j+1,k), 4..7 at k+1
const int HEXV[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
// VTK hex faces (grid_io.cpp:298-348): bottom, south, top, north, west, east
const int HEXF[6][4] = {{0, 1, 2, 3}, {0, 1, 5, 4}, {4, 5, 6, 7}, {2, 3, 7, 6}, {0, 4, 7, 3}, {1, 5, 6, 2}};
// neighbour offset across each face
const int HEXN[6][3] = {{0, 0, -1}, {0, -1, 0}, {0, 0, 1}, {0, 1, 0}, {-1, 0, 0}, {1, 0, 0}};

// The cell whose VTK face defines the orientation of the interface between (i,j,k) and its
// neighbour across local face lf, in the GLOBAL discovery order (cells k-major, i-minor): the
// earlier of the two cells; on the boundary the cell itself.  Returns that cell and its local face.
inline void discoverer(const ibis_b200_box& b, int i, int j, int k, int lf, int& di, int& dj, int& dk, int& dlf) {
    const int ni = i + HEXN[lf][0], nj = j + HEXN[lf][1], nk = k + HEXN[lf][2];
    const bool inside = ni >= 0 && ni < b.nx && nj >= 0 && nj < b.ny && nk >= 0 && nk < b.nz;
    // lower faces (bottom, south, west) belong to the neighbour (as its top, north, east) if it exists
    if (inside && (lf == 0 || lf == 1 || lf == 4)) {
        di = ni; dj = nj; dk = nk;
        dlf = (lf == 0) ? 2 : (lf == 1) ? 3 : 5;
    } else {
        di = i; dj = j; dk = k; dlf = lf;
    }
}

int build(ibis_b200_box& b) {
    const int nx = b.nx, ny = b.ny, nz = b.nz;
    // ---- layers held locally: owned, ring-1 lower/upper, ring-2 lower/upper ----
    std::vector<int> layers;                       // global k of each local slot
    for (int k = b.k_lo; k < b.k_hi; k++) layers.push_back(k);
    const int n_owned_layers = (int)layers.size();
    const bool has_lo = b.k_lo > 0, has_hi = b.k_hi < nz;
    if (has_lo) layers.push_back(b.k_lo - 1);
    if (has_hi) layers.push_back(b.k_hi);
    const int n_face_layers = (int)layers.size();  // owned + ring 1: cells with a face list
    if (has_lo && b.k_lo - 2 >= 0) layers.push_back(b.k_lo - 2);
    if (has_hi && b.k_hi + 1 < nz) layers.push_back(b.k_hi + 1);
    b.n_layers = (int)layers.size();
    b.kslot.assign(nz, -1);
    for (int s = 0; s < b.n_layers; s++) b.kslot[layers[s]] = s;
    const int64_t plane = (int64_t)nx * ny;
    const int64_t n_owned = plane * n_owned_layers, n_fcells = plane * n_face_layers, n_cells = plane * b.n_layers;

    // ---- face numbering: order of first discovery over the local cells that carry faces ----
    // id of the interface on the low side of cell (i,j,k) along axis a: fid[a][...]; the high-side
    // interface of the last cell is stored one further.
    auto has_faces = [&](int k) { return k >= 0 && k < nz && b.kslot[k] >= 0 && b.kslot[k] < n_face_layers; };
    // per (slot, j, i): number of new faces of that cell -> prefix sum -> first new face id
    std::vector<int64_t> first_new((size_t)n_fcells + 1, 0);
    auto is_new = [&](int i, int j, int k, int lf) {
        // new unless the neighbour across carries faces and comes EARLIER in the local cell order
        const int ni = i + HEXN[lf][0], nj = j + HEXN[lf][1], nk = k + HEXN[lf][2];
        if (ni < 0 || ni >= nx || nj < 0 || nj >= ny || nk < 0 || nk >= nz) return true;
        if (!has_faces(nk)) return true;
        return b.cell_id(ni, nj, nk) > b.cell_id(i, j, k);
    };
#pragma omp parallel for schedule(static)
    for (int s = 0; s < n_face_layers; s++) {
        const int k = layers[s];
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                int cnt = 0;
                for (int lf = 0; lf < 6; lf++) cnt += is_new(i, j, k, lf) ? 1 : 0;
                first_new[(size_t)(i + (int64_t)nx * (j + (int64_t)ny * s)) + 1] = cnt;
            }
    }
    for (int64_t c = 0; c < n_fcells; c++) first_new[c + 1] += first_new[c];
    const int64_t n_faces = first_new[n_fcells];
    if (n_cells >= ((int64_t)1 << 31) || n_faces >= ((int64_t)1 << 31)) {
        g_box_error = "box too large for 32-bit ids on one GPU";
        return IBIS_B200_EINVAL;
    }
    // face id of local face lf of cell (i,j,k): own numbering if new, else the neighbour's
    auto face_id = [&](int i, int j, int k, int lf) -> int64_t {
        int ci = i, cj = j, ck = k, clf = lf;
        if (!is_new(i, j, k, lf)) {
            ci = i + HEXN[lf][0]; cj = j + HEXN[lf][1]; ck = k + HEXN[lf][2];
            clf = (lf == 0) ? 2 : (lf == 1) ? 3 : (lf == 2) ? 0 : (lf == 3) ? 1 : (lf == 4) ? 5 : 4;
        }
        const int64_t c = b.cell_id(ci, cj, ck);
        int64_t id = first_new[c];
        for (int q = 0; q < clf; q++) id += is_new(ci, cj, ck, q) ? 1 : 0;
        return id;
    };

    // ---- boundary tags, alphabetical: bottom east north south top west (su2.box_hex marker order) ----
    // local boundary faces are those of face-carrying cells; marker order is first index outer
    std::vector<int> fk;                           // global k of face-carrying layers, ascending
    for (int s = 0; s < n_face_layers; s++) fk.push_back(layers[s]);
    std::sort(fk.begin(), fk.end());
    b.tag_off.assign(7, 0);
    struct BFace { int i, j, k, lf; };
    std::vector<BFace> bfaces;
    auto end_tag = [&](int t) { b.tag_off[t + 1] = (int64_t)bfaces.size(); };
    if (has_faces(0) ) for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) bfaces.push_back({i, j, 0, 0});
    end_tag(0);                                                                         // bottom
    for (int j = 0; j < ny; j++) for (int k : fk) bfaces.push_back({nx - 1, j, k, 5});
    end_tag(1);                                                                         // east
    for (int i = 0; i < nx; i++) for (int k : fk) bfaces.push_back({i, ny - 1, k, 3});
    end_tag(2);                                                                         // north
    for (int i = 0; i < nx; i++) for (int k : fk) bfaces.push_back({i, 0, k, 1});
    end_tag(3);                                                                         // south
    if (has_faces(nz - 1)) for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) bfaces.push_back({i, j, nz - 1, 2});
    end_tag(4);                                                                         // top
    for (int j = 0; j < ny; j++) for (int k : fk) bfaces.push_back({0, j, k, 4});
    end_tag(5);                                                                         // west
    const int64_t n_ghost = (int64_t)bfaces.size();
    const int64_t n_total = n_cells + n_ghost;
    if (n_total >= ((int64_t)1 << 31)) {
        g_box_error = "box too large for 32-bit ids on one GPU";
        return IBIS_B200_EINVAL;
    }

    // ---- allocate ----
    b.face_left.assign(n_faces, -1);
    b.face_right.assign(n_faces, -1);
    b.f_area.resize(n_faces);
    b.f_norm.resize(3 * n_faces);
    b.f_t1.resize(3 * n_faces);
    b.f_t2.resize(3 * n_faces);
    b.f_c.resize(3 * n_faces);
    b.cf_off.resize(n_cells + 1);
    b.cf_ids.resize(6 * n_fcells);
    b.cf_out.resize(6 * n_fcells);
    b.cf_nbr.resize(6 * n_fcells);
    b.c_vol.assign(n_cells, 0.0);
    b.c_cen.assign(3 * n_total, 0.0);
    b.tag_faces.resize(n_ghost);
    for (int64_t c = 0; c <= n_cells; c++) b.cf_off[c] = 6 * std::min(c, n_fcells);

    // ---- cell centroids (cell.h:236-254): sequential vertex sum / 8 ----
#pragma omp parallel for schedule(static)
    for (int s = 0; s < b.n_layers; s++) {
        const int k = layers[s];
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                double sx = 0.0, sy = 0.0, sz = 0.0;
                for (int v = 0; v < 8; v++) {
                    const V3 p = b.vertex(i + HEXV[v][0], j + HEXV[v][1], k + HEXV[v][2]);
                    sx += p.x; sy += p.y; sz += p.z;
                }
                const int64_t c = b.cell_id(i, j, k);
                b.c_cen[3 * c + 0] = sx / 8.0;
                b.c_cen[3 * c + 1] = sy / 8.0;
                b.c_cen[3 * c + 2] = sz / 8.0;
            }
    }

    // ---- faces: geometry from the global first discoverer's vertex order; attach cells ----
#pragma omp parallel for schedule(static)
    for (int s = 0; s < n_face_layers; s++) {
        const int k = layers[s];
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                const int64_t c = b.cell_id(i, j, k);
                for (int lf = 0; lf < 6; lf++) {
                    const int64_t f = face_id(i, j, k, lf);
                    b.cf_ids[6 * c + lf] = f;
                    if (!is_new(i, j, k, lf)) continue;       // geometry written by the cell that numbers it
                    int di, dj, dk, dlf;
                    discoverer(b, i, j, k, lf, di, dj, dk, dlf);
                    V3 p[4];
                    for (int q = 0; q < 4; q++) {
                        const int v = HEXF[dlf][q];
                        p[q] = b.vertex(di + HEXV[v][0], dj + HEXV[v][1], dk + HEXV[v][2]);
                    }
                    const FaceGeom g = quad_geom(p);
                    b.f_area[f] = g.area;
                    b.f_norm[3 * f] = g.n.x; b.f_norm[3 * f + 1] = g.n.y; b.f_norm[3 * f + 2] = g.n.z;
                    b.f_t1[3 * f] = g.t1.x; b.f_t1[3 * f + 1] = g.t1.y; b.f_t1[3 * f + 2] = g.t1.z;
                    b.f_t2[3 * f] = g.t2.x; b.f_t2[3 * f + 1] = g.t2.y; b.f_t2[3 * f + 2] = g.t2.z;
                    b.f_c[3 * f] = g.c.x; b.f_c[3 * f + 1] = g.c.y; b.f_c[3 * f + 2] = g.c.z;
                }
            }
    }
    // left / right by the sign of (x_f - x_c).n (grid.h:186-213); every cell writes only its own side
#pragma omp parallel for schedule(static)
    for (int s = 0; s < n_face_layers; s++) {
        const int k = layers[s];
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                const int64_t c = b.cell_id(i, j, k);
                for (int lf = 0; lf < 6; lf++) {
                    const int64_t f = b.cf_ids[6 * c + lf];
                    const double dx = b.f_c[3 * f] - b.c_cen[3 * c], dy = b.f_c[3 * f + 1] - b.c_cen[3 * c + 1],
                                 dz = b.f_c[3 * f + 2] - b.c_cen[3 * c + 2];
                    const double dot = dx * b.f_norm[3 * f] + dy * b.f_norm[3 * f + 1] + dz * b.f_norm[3 * f + 2];
                    const bool left = dot > 0.0;
                    b.cf_out[6 * c + lf] = left ? 1 : -1;
                    // the neighbour takes the other side; halo ring-2 cells (no face list) are attached here
                    const int ni = i + HEXN[lf][0], nj = j + HEXN[lf][1], nk = k + HEXN[lf][2];
                    const bool inside = ni >= 0 && ni < nx && nj >= 0 && nj < ny && nk >= 0 && nk < nz;
                    const int64_t nb = inside ? b.cell_id(ni, nj, nk) : -1;
                    if (left) b.face_left[f] = c; else b.face_right[f] = c;
                    if (inside && !has_faces(nk)) {
                        if (left) b.face_right[f] = nb; else b.face_left[f] = nb;
                    }
                }
            }
    }
    // ghosts take the empty side (grid.h:221-230); ghost centroid mirrored through the face (grid.h:364-403)
    for (int64_t q = 0; q < n_ghost; q++) {
        const BFace& bf = bfaces[q];
        const int64_t c = b.cell_id(bf.i, bf.j, bf.k);
        const int64_t f = b.cf_ids[6 * c + bf.lf];
        const int64_t ghost = n_cells + q;
        b.tag_faces[q] = f;
        if (b.face_left[f] < 0) b.face_left[f] = ghost; else b.face_right[f] = ghost;
        for (int d = 0; d < 3; d++) {
            const double dxyz = b.f_c[3 * f + d] - b.c_cen[3 * c + d];
            b.c_cen[3 * ghost + d] = b.f_c[3 * f + d] + dxyz;
        }
    }
    // neighbours (grid.h:340-358) and volumes (geom.h:67-83): sequential sum in cell-face order / 3
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < n_fcells; c++) {
        double acc = 0.0;
        for (int lf = 0; lf < 6; lf++) {
            const int64_t f = b.cf_ids[6 * c + lf];
            b.cf_nbr[6 * c + lf] = (b.face_left[f] == c) ? b.face_right[f] : b.face_left[f];
            const double out = (double)b.cf_out[6 * c + lf];
            const double dotv = b.f_c[3 * f] * out * b.f_norm[3 * f] + b.f_c[3 * f + 1] * out * b.f_norm[3 * f + 1] +
                                b.f_c[3 * f + 2] * out * b.f_norm[3 * f + 2];
            acc = acc + dotv * b.f_area[f];
        }
        b.c_vol[c] = acc / 3.0;
    }
    // halo ring-2 cells: a positive placeholder volume (never used: no residual, no gradient there)
    for (int64_t c = n_fcells; c < n_cells; c++) b.c_vol[c] = 1.0;

    // ---- halo exchange lists ----
    b.n_peers = 0;
    b.send_off.assign(1, 0);
    b.recv_off.assign(1, 0);
    auto add_layer = [&](std::vector<int64_t>& dst, int k) {
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) dst.push_back(b.cell_id(i, j, k));
    };
    if (has_lo) {
        b.peer_dir[b.n_peers++] = -1;
        add_layer(b.send_cells, b.k_lo);                              // their upper ring 1
        if (b.k_lo + 1 < b.k_hi) add_layer(b.send_cells, b.k_lo + 1); // their upper ring 2
        add_layer(b.recv_cells, b.k_lo - 1);
        if (b.k_lo - 2 >= 0) add_layer(b.recv_cells, b.k_lo - 2);
        b.send_off.push_back((int64_t)b.send_cells.size());
        b.recv_off.push_back((int64_t)b.recv_cells.size());
    }
    if (has_hi) {
        b.peer_dir[b.n_peers++] = 1;
        add_layer(b.send_cells, b.k_hi - 1);
        if (b.k_hi - 2 >= b.k_lo) add_layer(b.send_cells, b.k_hi - 2);
        add_layer(b.recv_cells, b.k_hi);
        if (b.k_hi + 1 < nz) add_layer(b.recv_cells, b.k_hi + 1);
        b.send_off.push_back((int64_t)b.send_cells.size());
        b.recv_off.push_back((int64_t)b.recv_cells.size());
    }

    ibis_b200_grid_desc& d = b.desc;
    memset(&d, 0, sizeof d);
    d.dim = 3;
    d.n_tags = 6;
    d.n_cells = n_cells;
    d.n_ghost = n_ghost;
    d.n_faces = n_faces;
    d.face_left = b.face_left.data();
    d.face_right = b.face_right.data();
    d.face_area = b.f_area.data();
    d.face_norm = b.f_norm.data();
    d.face_tan1 = b.f_t1.data();
    d.face_tan2 = b.f_t2.data();
    d.face_centre = b.f_c.data();
    d.cell_face_offsets = b.cf_off.data();
    d.cell_face_ids = b.cf_ids.data();
    d.cell_outsigns = b.cf_out.data();
    d.cell_neighbours = b.cf_nbr.data();
    d.cell_volume = b.c_vol.data();
    d.cell_centroid = b.c_cen.data();
    d.tag_offsets = b.tag_off.data();
    d.tag_faces = b.tag_faces.data();
    d.n_owned = n_owned;
    return 0;
}

}  // namespace

extern "C" {

int ibis_b200_host_box_create(ibis_b200_box** out, int nx, int ny, int nz, double lx, double ly, double lz,
                              int k_lo, int k_hi) {
    if (!out || nx < 1 || ny < 1 || nz < 1 || k_lo < 0 || k_hi > nz || k_lo >= k_hi) {
        g_box_error = "bad box arguments";
        return IBIS_B200_EINVAL;
    }
    // a slab must be at least two layers thick when it has a neighbour (two-ring halo from ONE peer)
    if ((k_lo > 0 || k_hi < nz) && k_hi - k_lo < 2) {
        g_box_error = "a slab needs at least two cell layers";
        return IBIS_B200_EINVAL;
    }
    ibis_b200_box* b = new (std::nothrow) ibis_b200_box();
    if (!b) return IBIS_B200_ENOMEM;
    b->nx = nx; b->ny = ny; b->nz = nz;
    b->lx = lx; b->ly = ly; b->lz = lz;
    b->k_lo = k_lo; b->k_hi = k_hi;
    int rc;
    try {
        rc = build(*b);
    } catch (std::bad_alloc&) {
        g_box_error = "out of host memory";
        rc = IBIS_B200_ENOMEM;
    }
    if (rc != 0) {
        delete b;
        return rc;
    }
    *out = b;
    return 0;
}

const ibis_b200_grid_desc* ibis_b200_host_box_desc(const ibis_b200_box* b) { return &b->desc; }

int ibis_b200_host_box_halo(const ibis_b200_box* b, int* n_peers, int32_t peer_dir[2], const int64_t** send_offsets,
                            const int64_t** send_cells, const int64_t** recv_offsets, const int64_t** recv_cells) {
    *n_peers = b->n_peers;
    peer_dir[0] = b->peer_dir[0];
    peer_dir[1] = b->peer_dir[1];
    *send_offsets = b->send_off.data();
    *send_cells = b->send_cells.data();
    *recv_offsets = b->recv_off.data();
    *recv_cells = b->recv_cells.data();
    return 0;
}

void ibis_b200_host_box_free(ibis_b200_box* b) { delete b; }

const char* ibis_b200_host_box_last_error(void) { return g_box_error.c_str(); }

}  // extern "C"
