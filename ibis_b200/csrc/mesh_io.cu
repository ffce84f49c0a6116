// mesh_io.cu -- ibis_b200_mesh_read_su2: an SU2 text file -> ibis_b200_mesh_desc (host only).
//
// Replaces GridIO::GridIO(file) + GridIO::read_su2_grid (grid/grid_io.cpp:100-268): section headings
// NDIME / NPOIN / NELEM / NMARK (MARKER_TAG, MARKER_ELEMS), unknown lines ignored, the trailing index
// some writers put after a vertex or an element ignored, a tag that appears twice has its elements
// appended, tags come out alphabetical (the reference keeps them in a std::map).  The reference reads
// the file through getline + substr + stod/stoi one token at a time; here the file is read once, the
// line starts of a section are indexed, and the lines are parsed in parallel with strtod / strtoll
// (the same correctly rounded conversions, so coordinates are bit-identical).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/ibis_b200.h"
#include "context.h"
#include "grid_bodies.h"

using namespace ib;

struct ibis_b200_mesh {
    ibis_b200_mesh_desc desc;
    std::vector<double> vertices;
    std::vector<int32_t> cell_types, marker_types;
    std::vector<int64_t> cell_vertices, marker_offsets, marker_vertices;
    std::vector<std::string> tags;
};

namespace {

struct Lines {
    const char* buf;
    size_t size, pos = 0;
    Lines(const char* b, size_t n) : buf(b), size(n) {}
    // next line as [b, e), leading / trailing blanks trimmed; false at end of file
    bool next(const char*& b, const char*& e) {
        if (pos >= size) return false;
        const char* nl = (const char*)memchr(buf + pos, '\n', size - pos);
        const size_t end = nl ? (size_t)(nl - buf) : size;
        b = buf + pos;
        e = buf + end;
        pos = end + 1;
        while (b < e && (*b == ' ' || *b == '\t' || *b == '\r')) b++;
        while (e > b && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r')) e--;
        return true;
    }
};
bool starts_with(const char* b, const char* e, const char* prefix) {
    const size_t n = strlen(prefix);
    return (size_t)(e - b) >= n && memcmp(b, prefix, n) == 0;
}
std::string value_of(const char* b, const char* e) {      // NAME=VALUE -> VALUE, trimmed
    const char* eq = (const char*)memchr(b, '=', e - b);
    const char* v = eq ? eq + 1 : b;
    while (v < e && (*v == ' ' || *v == '\t')) v++;
    return std::string(v, e);
}
// element line "type v0 v1 ..." -> type, n vertex ids padded with -1 to `width`; false when malformed
// (strto* skip '\n' as white space: a token found past `end` belongs to the next line -- the line is short)
bool parse_element(const char* b, const char* end, int width, int32_t* type, int64_t* out) {
    char* q = nullptr;
    const long t = strtol(b, &q, 10);
    if (q == b || q > end) return false;
    const int n = t == 3 ? 2 : gb_num_vertices((int)t);      // 3 = line (boundary of a 2-D grid)
    if (n == 0 || n > width) return false;
    *type = (int32_t)t;
    for (int k = 0; k < width; k++) out[k] = -1;
    for (int k = 0; k < n; k++) {
        const char* s = q;
        out[k] = strtoll(s, &q, 10);
        if (q == s || q > end) return false;
    }
    return true;
}

}  // namespace

extern "C" {

int ibis_b200_mesh_read_su2(ibis_b200_mesh** out, const char* path) {
    if (!out || !path) return fail_global(IBIS_B200_EINVAL, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail_global(IBIS_B200_EINVAL, (std::string("Could not find ") + path).c_str());   // grid_io.cpp:103-106
    fseek(f, 0, SEEK_END);
    const long fsize = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)fsize + 1, '\0');
    const size_t got = fsize > 0 ? fread(buf.data(), 1, (size_t)fsize, f) : 0;
    fclose(f);
    if ((long)got != fsize) return fail_global(IBIS_B200_EINVAL, "short read of the grid file");
    ibis_b200_mesh* m = new (std::nothrow) ibis_b200_mesh();
    if (!m) return fail_global(IBIS_B200_ENOMEM, "out of host memory");
    struct Guard { ibis_b200_mesh* m; ~Guard() { delete m; } } guard{m};
    int dim = 0;
    std::map<std::string, std::pair<std::vector<int32_t>, std::vector<int64_t>>> markers;
    Lines L(buf.data(), (size_t)fsize);
    const char *b, *e;
    // the next n lines of the file, as start pointers (each line ends at '\n' or '\0' for strto*)
    // (a section cannot hold more lines than the file has bytes: a corrupt count fails here, not in the allocator)
    auto index_lines = [&](int64_t n, std::vector<const char*>& starts, std::vector<const char*>& ends) {
        if (n > (int64_t)fsize) return false;
        starts.resize((size_t)n);
        ends.resize((size_t)n);
        for (int64_t i = 0; i < n; i++) {
            if (!L.next(b, e)) return false;
            starts[i] = b;
            ends[i] = e;
        }
        return true;
    };
    try {
    while (L.next(b, e)) {
        if (starts_with(b, e, "NDIME")) {
            dim = atoi(value_of(b, e).c_str());
            if (dim != 2 && dim != 3)
                return fail_global(IBIS_B200_EINVAL, "Invalid number of dimensions in su2 file");   // grid_io.cpp:199-202
        } else if (starts_with(b, e, "NPOIN")) {
            const int64_t n = atoll(value_of(b, e).c_str());
            std::vector<const char*> starts, ends;
            if (dim == 0 || n < 0 || !index_lines(n, starts, ends)) return fail_global(IBIS_B200_EINVAL, "truncated NPOIN section");
            m->vertices.assign((size_t)3 * n, 0.0);
            bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
            for (int64_t i = 0; i < n; i++) {
                const char* s = starts[i];
                for (int d = 0; d < dim; d++) {
                    char* q = nullptr;
                    m->vertices[(size_t)3 * i + d] = strtod(s, &q);
                    if (q == s || q > ends[i]) bad = true;
                    s = q;
                }
            }
            if (bad) return fail_global(IBIS_B200_EINVAL, "malformed vertex line");
        } else if (starts_with(b, e, "NELEM")) {
            const int64_t n = atoll(value_of(b, e).c_str());
            std::vector<const char*> starts, ends;
            if (n < 0 || !index_lines(n, starts, ends)) return fail_global(IBIS_B200_EINVAL, "truncated NELEM section");
            m->cell_types.assign((size_t)n, 0);
            m->cell_vertices.assign((size_t)8 * n, -1);
            bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
            for (int64_t i = 0; i < n; i++)
                if (!parse_element(starts[i], ends[i], 8, &m->cell_types[i], &m->cell_vertices[(size_t)8 * i])) bad = true;
            if (bad) return fail_global(IBIS_B200_EINVAL, "malformed or unsupported element line");
        } else if (starts_with(b, e, "MARKER_TAG")) {
            const std::string tag = value_of(b, e);
            if (!L.next(b, e)) return fail_global(IBIS_B200_EINVAL, "truncated marker");
            const int64_t n = atoll(value_of(b, e).c_str());
            std::vector<const char*> starts, ends;
            if (n < 0 || !index_lines(n, starts, ends)) return fail_global(IBIS_B200_EINVAL, "truncated marker");
            auto& mk = markers[tag];                          // a repeated tag is appended to (grid_io.cpp:254-262)
            const size_t n0 = mk.first.size();
            mk.first.resize(n0 + (size_t)n);
            mk.second.resize(4 * (n0 + (size_t)n));
            bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
            for (int64_t i = 0; i < n; i++)
                if (!parse_element(starts[i], ends[i], 4, &mk.first[n0 + i], &mk.second[4 * (n0 + i)])) bad = true;
            if (bad) return fail_global(IBIS_B200_EINVAL, "malformed marker element line");
        }
    }
    if (dim == 0) return fail_global(IBIS_B200_EINVAL, "no NDIME line in the grid file");
    m->marker_offsets.assign(1, 0);
    for (auto& kv : markers) {                                // std::map: alphabetical
        m->tags.push_back(kv.first);
        m->marker_types.insert(m->marker_types.end(), kv.second.first.begin(), kv.second.first.end());
        m->marker_vertices.insert(m->marker_vertices.end(), kv.second.second.begin(), kv.second.second.end());
        m->marker_offsets.push_back((int64_t)m->marker_types.size());
    }
    ibis_b200_mesh_desc& d = m->desc;
    memset(&d, 0, sizeof(d));
    d.dim = dim;
    d.n_tags = (int32_t)m->tags.size();
    d.n_vertices = (int64_t)(m->vertices.size() / 3);
    d.n_cells = (int64_t)m->cell_types.size();
    d.vertices = m->vertices.data();
    d.cell_types = m->cell_types.data();
    d.cell_vertices = m->cell_vertices.data();
    d.marker_offsets = m->marker_offsets.data();
    d.marker_vertices = m->marker_vertices.data();
    guard.m = nullptr;
    *out = m;
    return 0;
    } catch (const std::bad_alloc&) {
        return fail_global(IBIS_B200_ENOMEM, "out of host memory while reading the grid file");
    } catch (const std::exception& ex) {
        return fail_global(IBIS_B200_EINVAL, ex.what());
    }
}

const ibis_b200_mesh_desc* ibis_b200_mesh_desc_of(const ibis_b200_mesh* m) { return m ? &m->desc : nullptr; }
const char* ibis_b200_mesh_tag(const ibis_b200_mesh* m, int i) {
    return (m && i >= 0 && i < (int)m->tags.size()) ? m->tags[i].c_str() : nullptr;
}
int ibis_b200_mesh_destroy(ibis_b200_mesh* m) {
    delete m;
    return 0;
}

}  // extern "C"
