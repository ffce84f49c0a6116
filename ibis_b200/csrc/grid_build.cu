// grid_build.cu -- ibis_b200_grid_build: an SU2 mesh (plain arrays) -> the arrays of the reference
// GridBlock, in the reference's numbering (include/ibis_b200.h, "grid build").
//
// The reference discovers faces cell by cell through an unordered_map keyed on strings built from
// the sorted vertex ids (grid/interface.cpp:38-80; ~17 us per cell).  Here:
//   host    every (cell, local face) record gets the sorted vertex tuple as key; ONE parallel
//           lexicographic sort groups the records of a face; a face's id is the rank of its first
//           record among all first records (= the reference's order of first discovery), by a flag
//           array + prefix sum; boundary markers find their face by binary search in the sorted keys;
//   device  face centres / areas / tangents / normals, cell centroids, left-right attachment and
//           outsigns, ghost cells and their mirrored centroids, volumes and neighbours:
//           five kernels over faces / cells with the bodies of grid_bodies.h.
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>
#if defined(_OPENMP)
#include <parallel/algorithm>
#define IBIS_SORT __gnu_parallel::sort
#else
#define IBIS_SORT std::sort
#endif

#include "../../include/ibis_b200.h"
#include "backend.h"
#include "context.h"
#include "grid_bodies.h"

using namespace ib;

struct ibis_b200_grid {
    ibis_b200_grid_desc desc;
    std::vector<int64_t> face_left, face_right, cf_off, cf_face, cf_nbr, tag_off, tag_faces, tag_ghosts, fverts;
    std::vector<int32_t> cf_out;
    std::vector<double> f_area, f_norm, f_tan1, f_tan2, f_centre, c_volume, c_centroid;
};

namespace {

// VTK local face tables (grid/grid_io.cpp:298-348): local vertex tuples of every face of a cell type
struct FaceTable { int nf; int nv[6]; int v[6][4]; };
const FaceTable* face_table(int t) {
    static const FaceTable tri = {3, {2, 2, 2}, {{0, 1}, {1, 2}, {2, 0}}};
    static const FaceTable quad = {4, {2, 2, 2, 2}, {{0, 1}, {1, 2}, {2, 3}, {3, 0}}};
    static const FaceTable tet = {4, {3, 3, 3, 3}, {{0, 1, 2}, {0, 1, 3}, {1, 2, 3}, {0, 2, 3}}};
    static const FaceTable hex = {6, {4, 4, 4, 4, 4, 4},
                                  {{0, 1, 2, 3}, {0, 1, 5, 4}, {4, 5, 6, 7}, {2, 3, 7, 6}, {0, 4, 7, 3}, {1, 5, 6, 2}}};
    static const FaceTable wedge = {5, {3, 3, 4, 4, 4}, {{0, 1, 2}, {3, 5, 4}, {1, 4, 5, 2}, {0, 2, 5, 3}, {0, 3, 4, 1}}};
    static const FaceTable pyr = {5, {4, 3, 3, 3, 3}, {{0, 3, 2, 1}, {2, 3, 4}, {0, 4, 3}, {0, 1, 4}, {1, 2, 4}}};
    switch (t) {
        case VTK_TRI: return &tri;
        case VTK_QUAD: return &quad;
        case VTK_TETRA: return &tet;
        case VTK_HEX: return &hex;
        case VTK_WEDGE: return &wedge;
        case VTK_PYRAMID: return &pyr;
    }
    return nullptr;
}

struct Rec {
    int64_t k[4];    // sorted vertex ids, -1 padding first
    int64_t rec;     // record index = position in the reference's discovery loop
};
inline bool key_less(const int64_t* a, const int64_t* b) {
    for (int i = 0; i < 4; i++)
        if (a[i] != b[i]) return a[i] < b[i];
    return false;
}
inline bool key_eq(const int64_t* a, const int64_t* b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3]; }
inline void sort4(int64_t* k) {
    if (k[0] > k[1]) std::swap(k[0], k[1]);
    if (k[2] > k[3]) std::swap(k[2], k[3]);
    if (k[0] > k[2]) std::swap(k[0], k[2]);
    if (k[1] > k[3]) std::swap(k[1], k[3]);
    if (k[1] > k[2]) std::swap(k[1], k[2]);
}

enum { OP_FACE = 0, OP_CENTROID = 1, OP_ATTACH = 2, OP_GHOST = 3, OP_VOLUME = 4 };
template <int OP>
IBIS_HD void gb_apply(const GridBuild& G, int64_t i) {
    if (OP == OP_FACE) gb_face_geometry(G, i);
    else if (OP == OP_CENTROID) gb_cell_centroid(G, i);
    else if (OP == OP_ATTACH) gb_attach(G, i);
    else if (OP == OP_GHOST) gb_ghost(G, i);
    else gb_volume_neighbours(G, i);
}

#if defined(__CUDACC__) && !defined(IBIS_B200_EMU)
template <int OP>
__global__ void __launch_bounds__(256) k_grid_build(const GridBuild G, const int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gb_apply<OP>(G, i);
}
template <int OP>
int gb_launch(const GridBuild& G, int64_t n, gpu::stream_t s) {
    if (n == 0) return 0;
    k_grid_build<OP><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(G, n);
    return gpu::last_error();
}
#else
template <int OP>
int gb_launch(const GridBuild& G, int64_t n, gpu::stream_t) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) gb_apply<OP>(G, i);
    return 0;
}
#endif

// IBIS_B200_TIMING=1: wall time of the stages of ibis_b200_grid_build on stderr
struct Laps {
    bool on;
    double t0;
    static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
    Laps() : on(getenv("IBIS_B200_TIMING") != nullptr), t0(now()) {}
    void lap(const char* what) {
        if (!on) return;
        const double t = now();
        fprintf(stderr, "[ibis_b200 grid_build] %-34s %8.3f s\n", what, t - t0);
        t0 = t;
    }
};

// device arrays of one build, freed on every exit path
struct DevPool {
    std::vector<void*> ptrs;
    gpu::stream_t stream = nullptr;
    ~DevPool() {
        for (void* p : ptrs) gpu::dev_free(p);
        if (stream) gpu::stream_destroy(stream);
    }
    template <typename T>
    int alloc(T** d, size_t n) {
        void* p = nullptr;
        if (gpu::dev_malloc(&p, n * sizeof(T)) != 0) return IBIS_B200_ENOMEM;
        ptrs.push_back(p);
        *d = static_cast<T*>(p);
        return 0;
    }
    template <typename T>
    int upload(const T** d, const T* h, size_t n) {
        T* p = nullptr;
        if (int rc = alloc(&p, n)) return rc;
        if (n && gpu::h2d(p, h, n * sizeof(T), stream) != 0) return IBIS_B200_ECUDA;
        *d = p;
        return 0;
    }
};

}  // namespace

extern "C" {

int ibis_b200_grid_build(ibis_b200_grid** out, const ibis_b200_mesh_desc* m, int device) {
    if (!out || !m || !m->vertices || !m->cell_types || !m->cell_vertices || (m->n_tags > 0 && !m->marker_offsets))
        return fail_global(IBIS_B200_EINVAL, "null argument");
    if (m->dim != 2 && m->dim != 3) return fail_global(IBIS_B200_EINVAL, "grid dimension must be 2 or 3");
    const int64_t nc = m->n_cells;
    Laps laps;
    // ---- (cell, local face) records in the reference's loop order ----
    std::vector<int64_t> cf_off((size_t)nc + 1, 0);
    for (int64_t c = 0; c < nc; c++) {
        const FaceTable* ft = face_table(m->cell_types[c]);
        const bool flat = m->cell_types[c] == VTK_TRI || m->cell_types[c] == VTK_QUAD;
        if (!ft || flat != (m->dim == 2)) return fail_global(IBIS_B200_EINVAL, "unsupported cell type in mesh");
        cf_off[c + 1] = cf_off[c] + ft->nf;
    }
    const int64_t nr = cf_off[nc];
    std::vector<Rec> recs((size_t)nr);
    std::vector<int64_t> rec_verts((size_t)4 * nr);
    bool bad_vertex = false;
#pragma omp parallel for schedule(static) reduction(|| : bad_vertex)
    for (int64_t c = 0; c < nc; c++) {
        const FaceTable* ft = face_table(m->cell_types[c]);
        for (int j = 0; j < ft->nf; j++) {
            const int64_t r = cf_off[c] + j;
            int64_t* v = &rec_verts[(size_t)4 * r];
            for (int k = 0; k < 4; k++) {
                v[k] = k < ft->nv[j] ? m->cell_vertices[8 * c + ft->v[j][k]] : -1;
                if (k < ft->nv[j] && (v[k] < 0 || v[k] >= m->n_vertices)) bad_vertex = true;
                recs[r].k[k] = v[k];
            }
            sort4(recs[r].k);
            recs[r].rec = r;
        }
    }
    if (bad_vertex) return fail_global(IBIS_B200_EINVAL, "cell refers to a vertex outside the vertex list");
    laps.lap("cell-face records");
    // ---- one sort groups the records of every face ----
    IBIS_SORT(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) {
        return key_eq(a.k, b.k) ? a.rec < b.rec : key_less(a.k, b.k);
    });
    laps.lap("sort");
    // face id = rank of the face's first record among all first records
    std::vector<int64_t> first_rank((size_t)nr + 1, 0);      // flag, then exclusive prefix sum
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nr; i++)
        if (i == 0 || !key_eq(recs[i].k, recs[i - 1].k)) first_rank[recs[i].rec] = 1;
    int64_t nf = 0;
    for (int64_t r = 0; r < nr; r++) {
        const int64_t flag = first_rank[r];
        first_rank[r] = nf;
        nf += flag;
    }
    std::vector<int64_t> cf_face((size_t)nr), group_face((size_t)nr);   // group_face: face of sorted position i
    {
        // positions of the group starts, so that every sorted entry finds its group's first record
        std::vector<int64_t> start((size_t)nr);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nr; i++) start[i] = (i == 0 || !key_eq(recs[i].k, recs[i - 1].k)) ? i : -1;
        int64_t cur = 0;
        for (int64_t i = 0; i < nr; i++) {
            if (start[i] >= 0) cur = i;
            start[i] = cur;
        }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nr; i++) {
            group_face[i] = first_rank[recs[start[i]].rec];
            cf_face[recs[i].rec] = group_face[i];
        }
    }
    ibis_b200_grid* g = new (std::nothrow) ibis_b200_grid();
    if (!g) return fail_global(IBIS_B200_ENOMEM, "out of host memory");
    struct Guard { ibis_b200_grid* g; ~Guard() { delete g; } } guard{g};
    g->fverts.assign((size_t)4 * nf, -1);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nr; i++)
        if (i == 0 || !key_eq(recs[i].k, recs[i - 1].k))
            memcpy(&g->fverts[(size_t)4 * group_face[i]], &rec_verts[(size_t)4 * recs[i].rec], 4 * sizeof(int64_t));
    // ---- boundary markers: face by binary search, ghosts numbered per tag in marker order ----
    const int64_t n_bf = m->n_tags > 0 ? m->marker_offsets[m->n_tags] : 0;
    if (n_bf > 0 && !m->marker_vertices) return fail_global(IBIS_B200_EINVAL, "null argument");
    g->tag_off.assign(1, 0);
    if (m->n_tags > 0) g->tag_off.assign(m->marker_offsets, m->marker_offsets + m->n_tags + 1);
    g->tag_faces.assign((size_t)n_bf, 0);
    g->tag_ghosts.assign((size_t)n_bf, 0);
    bool unmatched = false;
#pragma omp parallel for schedule(static) reduction(|| : unmatched)
    for (int64_t k = 0; k < n_bf; k++) {
        Rec key;
        for (int q = 0; q < 4; q++) key.k[q] = m->marker_vertices[4 * k + q];
        sort4(key.k);
        auto it = std::lower_bound(recs.begin(), recs.end(), key, [](const Rec& a, const Rec& b) { return key_less(a.k, b.k); });
        if (it == recs.end() || !key_eq(it->k, key.k)) { unmatched = true; continue; }
        g->tag_faces[k] = group_face[it - recs.begin()];
        g->tag_ghosts[k] = nc + k;
    }
    if (unmatched) return fail_global(IBIS_B200_EINVAL, "boundary marker element does not match any cell face");
    std::vector<Rec>().swap(recs);
    std::vector<int64_t>().swap(rec_verts);
    std::vector<int64_t>().swap(group_face);

    laps.lap("face ids, markers");
    // ---- geometry on the device ----
    const int64_t nt = nc + n_bf;
    int prev_dev = -1;
    gpu::get_device(&prev_dev);
    if (device >= 0 && gpu::set_device(device) != 0) return fail_global(IBIS_B200_ECUDA, "cannot select the device");
    struct DevRestore { int d; ~DevRestore() { if (d >= 0) gpu::set_device(d); } } restore{device >= 0 ? prev_dev : -1};
    DevPool P;
    if (gpu::stream_create(&P.stream) != 0) return fail_global(IBIS_B200_ECUDA, "cannot create a stream");
    GridBuild G;
    memset(&G, 0, sizeof(G));
    G.dim = m->dim;
    G.n_vertices = m->n_vertices; G.n_cells = nc; G.n_ghost = n_bf; G.n_faces = nf; G.n_records = nr; G.n_bfaces = n_bf;
    int rc = 0;
#define TRY(e) do { if (!rc) rc = (e); } while (0)
    TRY(P.upload(&G.pos, m->vertices, (size_t)3 * m->n_vertices));
    TRY(P.upload(&G.ctype, m->cell_types, (size_t)nc));
    TRY(P.upload(&G.cverts, m->cell_vertices, (size_t)8 * nc));
    TRY(P.upload(&G.fverts, (const int64_t*)g->fverts.data(), (size_t)4 * nf));
    TRY(P.upload(&G.cf_off, (const int64_t*)cf_off.data(), (size_t)nc + 1));
    TRY(P.upload(&G.cf_face, (const int64_t*)cf_face.data(), (size_t)nr));
    TRY(P.upload(&G.bf_face, (const int64_t*)g->tag_faces.data(), (size_t)n_bf));
    TRY(P.alloc(&G.f_centre, (size_t)3 * nf)); TRY(P.alloc(&G.f_area, (size_t)nf));
    TRY(P.alloc(&G.f_tan1, (size_t)3 * nf)); TRY(P.alloc(&G.f_tan2, (size_t)3 * nf)); TRY(P.alloc(&G.f_norm, (size_t)3 * nf));
    TRY(P.alloc(&G.c_centroid, (size_t)3 * nt)); TRY(P.alloc(&G.c_volume, (size_t)nc));
    TRY(P.alloc(&G.f_left, (size_t)nf)); TRY(P.alloc(&G.f_right, (size_t)nf));
    TRY(P.alloc(&G.cf_out, (size_t)nr)); TRY(P.alloc(&G.cf_nbr, (size_t)nr));
    if (rc) return fail_global(rc, "grid build: device allocation / upload failed");
    // GB_NONE = 0x7fff...ff: every byte but the top one is 0xff, so fill on the host side of the copy
    {
        std::vector<int64_t> none((size_t)nf, GB_NONE);
        if (nf && (gpu::h2d(G.f_left, none.data(), (size_t)nf * 8, P.stream) || gpu::h2d(G.f_right, none.data(), (size_t)nf * 8, P.stream) ||
                   gpu::stream_sync(P.stream)))
            return fail_global(IBIS_B200_ECUDA, "grid build: upload failed");
    }
    if (gpu::memset_async(G.c_centroid, 0, (size_t)3 * nt * 8, P.stream)) return fail_global(IBIS_B200_ECUDA, "grid build: memset failed");
    if (laps.on) gpu::stream_sync(P.stream);
    laps.lap("upload");
    TRY(gb_launch<OP_FACE>(G, nf, P.stream));
    TRY(gb_launch<OP_CENTROID>(G, nc, P.stream));
    TRY(gb_launch<OP_ATTACH>(G, nc, P.stream));
    TRY(gb_launch<OP_GHOST>(G, n_bf, P.stream));
    TRY(gb_launch<OP_VOLUME>(G, nc, P.stream));
    if (rc) return fail_global(IBIS_B200_ECUDA, (std::string("grid build kernels: ") + gpu::error_string(rc)).c_str());
    if (laps.on) gpu::stream_sync(P.stream);
    laps.lap("geometry kernels (5 launches)");
    g->face_left.resize((size_t)nf); g->face_right.resize((size_t)nf);
    g->f_area.resize((size_t)nf); g->f_norm.resize((size_t)3 * nf); g->f_tan1.resize((size_t)3 * nf);
    g->f_tan2.resize((size_t)3 * nf); g->f_centre.resize((size_t)3 * nf);
    g->c_volume.resize((size_t)nc); g->c_centroid.resize((size_t)3 * nt);
    g->cf_out.resize((size_t)nr); g->cf_nbr.resize((size_t)nr);
#define DOWN(vec, ptr) TRY(vec.empty() ? 0 : gpu::d2h(vec.data(), ptr, vec.size() * sizeof(vec[0]), P.stream));
    DOWN(g->face_left, G.f_left) DOWN(g->face_right, G.f_right) DOWN(g->f_area, G.f_area) DOWN(g->f_norm, G.f_norm)
    DOWN(g->f_tan1, G.f_tan1) DOWN(g->f_tan2, G.f_tan2) DOWN(g->f_centre, G.f_centre) DOWN(g->c_volume, G.c_volume)
    DOWN(g->c_centroid, G.c_centroid) DOWN(g->cf_out, G.cf_out) DOWN(g->cf_nbr, G.cf_nbr)
    TRY(gpu::stream_sync(P.stream));
#undef DOWN
#undef TRY
    if (rc) return fail_global(IBIS_B200_ECUDA, (std::string("grid build download: ") + gpu::error_string(rc)).c_str());
    laps.lap("download");
    bool open_face = false;
#pragma omp parallel for schedule(static) reduction(|| : open_face)
    for (int64_t f = 0; f < nf; f++)
        if (g->face_left[f] == GB_NONE || g->face_right[f] == GB_NONE) open_face = true;
    if (open_face)
        return fail_global(IBIS_B200_EINVAL, "grid has faces without two cells: a boundary marker is missing or the mesh is non-conforming");
    g->cf_off.swap(cf_off);
    g->cf_face.swap(cf_face);
    ibis_b200_grid_desc& d = g->desc;
    memset(&d, 0, sizeof(d));
    d.dim = m->dim; d.n_tags = m->n_tags > 0 ? m->n_tags : 0;
    d.n_cells = nc; d.n_ghost = n_bf; d.n_faces = nf; d.n_owned = nc;
    d.face_left = g->face_left.data(); d.face_right = g->face_right.data();
    d.face_area = g->f_area.data(); d.face_norm = g->f_norm.data(); d.face_tan1 = g->f_tan1.data();
    d.face_tan2 = g->f_tan2.data(); d.face_centre = g->f_centre.data();
    d.cell_face_offsets = g->cf_off.data(); d.cell_face_ids = g->cf_face.data(); d.cell_outsigns = g->cf_out.data();
    d.cell_neighbours = g->cf_nbr.data(); d.cell_volume = g->c_volume.data(); d.cell_centroid = g->c_centroid.data();
    d.tag_offsets = g->tag_off.data(); d.tag_faces = g->tag_faces.data();
    guard.g = nullptr;
    *out = g;
    return 0;
}

const ibis_b200_grid_desc* ibis_b200_grid_desc_of(const ibis_b200_grid* g) { return g ? &g->desc : nullptr; }
const int64_t* ibis_b200_grid_face_vertices(const ibis_b200_grid* g) { return g ? g->fverts.data() : nullptr; }
const int64_t* ibis_b200_grid_tag_ghosts(const ibis_b200_grid* g) { return g ? g->tag_ghosts.data() : nullptr; }
int ibis_b200_grid_destroy(ibis_b200_grid* g) {
    delete g;
    return 0;
}

}  // extern "C"
