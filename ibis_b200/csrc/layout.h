// layout.h -- internal renumbering of the grid handed over in reference numbering
// (plain host C++; included by abi.cu for the CUDA library and for the tests/emu build).
//
// Cells are reordered along a Morton (Z-order) curve of their centroids inside each class --
// owned | halo with faces (ring 1) | halo without faces (ring 2) | boundary ghosts -- so that
// consecutive ids form compact bricks: neighbour gathers of a thread block stay in L1/L2 and the
// tiled residual kernel shares most interface fluxes inside a block.  Interfaces are renumbered
// in order of first discovery over the new cell order, which makes the interfaces "homed" in a
// tile of cells one contiguous range.  None of this is visible through the ABI: the ref2int
// permutation is applied at upload / download, and the per-cell arithmetic (face order inside a
// cell, orientation of every interface) is untouched, so results do not depend on it.
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "ibis_b200.h"

namespace ib {

struct Renumbered {
    ibis_b200_grid_desc desc;                 // same content as the input, new numbering
    std::vector<int> ref2int, int2ref;        // cells incl. ghosts
    std::vector<int> face_ref2int;            // reference interface id -> internal id (both renumberings composed)
    std::vector<int64_t> face_left, face_right, cf_ids, cf_nbr, tag_faces;
    std::vector<int32_t> cf_out;
    std::vector<int64_t> cf_off;
    std::vector<double> f_area, f_norm, f_t1, f_t2, f_c, c_vol, c_cen;
    std::vector<int64_t> cell_first_face;     // per cell with faces: face counter before its discoveries
    int64_t n_grad = 0;                       // cells [0, n_grad) carry a face list
};

inline uint64_t spread3(uint64_t x) {        // 21 bits -> every third bit
    x &= 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}
inline uint64_t spread2(uint64_t x) {        // 31 bits -> every second bit
    x &= 0x7fffffff;
    x = (x | x << 16) & 0x0000ffff0000ffffULL;
    x = (x | x << 8) & 0x00ff00ff00ff00ffULL;
    x = (x | x << 4) & 0x0f0f0f0f0f0f0f0fULL;
    x = (x | x << 2) & 0x3333333333333333ULL;
    x = (x | x << 1) & 0x5555555555555555ULL;
    return x;
}

// stable LSD radix sort of idx[] by key[idx]
inline void radix_sort_by_key(const std::vector<uint64_t>& key, std::vector<int>& idx) {
    const size_t n = idx.size();
    if (n < 2) return;
    uint64_t maxk = 0;
    for (size_t i = 0; i < n; i++) maxk = std::max(maxk, key[idx[i]]);
    std::vector<int> tmp(n);
    for (int shift = 0; shift < 64 && (maxk >> shift) != 0; shift += 16) {
        std::vector<size_t> count(65537, 0);
        for (size_t i = 0; i < n; i++) count[((key[idx[i]] >> shift) & 0xffff) + 1]++;
        for (int b = 0; b < 65536; b++) count[b + 1] += count[b];
        for (size_t i = 0; i < n; i++) tmp[count[(key[idx[i]] >> shift) & 0xffff]++] = idx[i];
        idx.swap(tmp);
    }
}

// returns 0, or a negative code with *err set
inline int renumber_grid(const ibis_b200_grid_desc* gd, bool morton, Renumbered& R, const char** err) {
    const int64_t n_cells = gd->n_cells, n_ghost = gd->n_ghost, n_total = n_cells + n_ghost, n_faces = gd->n_faces;
    const int64_t n_owned = gd->n_owned > 0 ? gd->n_owned : n_cells;
    const int dim = gd->dim;
    // cells with a face list must form a prefix of [n_owned, n_cells)
    int64_t n_grad = n_cells;
    for (int64_t i = 0; i < n_cells; i++)
        if (gd->cell_face_offsets[i + 1] == gd->cell_face_offsets[i]) { n_grad = i; break; }
    for (int64_t i = n_grad; i < n_cells; i++)
        if (gd->cell_face_offsets[i + 1] != gd->cell_face_offsets[i]) {
            *err = "cells without a face list (outer halo ring) must come last among the cells";
            return IBIS_B200_EINVAL;
        }
    if (n_grad < n_owned) { *err = "owned cell without faces"; return IBIS_B200_EINVAL; }
    R.n_grad = n_grad;

    // ---- cell order ----
    R.int2ref.resize(n_total);
    for (int64_t i = 0; i < n_total; i++) R.int2ref[i] = (int)i;
    if (morton && n_cells > 1) {
        double lo[3] = {1e300, 1e300, 1e300}, vsum = 0.0;
        for (int64_t i = 0; i < n_cells; i++)
            for (int d = 0; d < 3; d++) lo[d] = std::min(lo[d], gd->cell_centroid[3 * i + d]);
        for (int64_t i = 0; i < n_grad; i++) vsum += fabs(gd->cell_volume[i]);
        double h = dim == 3 ? cbrt(vsum / (double)n_grad) : sqrt(vsum / (double)n_grad);
        if (!(h > 0.0) || !std::isfinite(h)) h = 1.0;
        std::vector<uint64_t> key(n_total, 0);
        const double qmax = dim == 3 ? 2097151.0 : 2147483647.0;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n_cells; i++) {
            uint64_t q[3];
            for (int d = 0; d < 3; d++) {
                double v = floor((gd->cell_centroid[3 * i + d] - lo[d]) / h);
                if (!(v > 0.0)) v = 0.0;
                if (v > qmax) v = qmax;
                q[d] = (uint64_t)v;
            }
            key[i] = dim == 3 ? (spread3(q[0]) | spread3(q[1]) << 1 | spread3(q[2]) << 2)
                              : (spread2(q[0]) | spread2(q[1]) << 1);
        }
        const int64_t bounds[4] = {0, n_owned, n_grad, n_cells};
        for (int cls = 0; cls < 3; cls++) {
            std::vector<int> idx(R.int2ref.begin() + bounds[cls], R.int2ref.begin() + bounds[cls + 1]);
            radix_sort_by_key(key, idx);
            std::copy(idx.begin(), idx.end(), R.int2ref.begin() + bounds[cls]);
        }
    }
    R.ref2int.resize(n_total);
    for (int64_t i = 0; i < n_cells; i++) R.ref2int[R.int2ref[i]] = (int)i;
    // ghosts follow their interior cell
    if (n_ghost > 0) {
        std::vector<uint64_t> gkey(n_total, 0);
        const int64_t n_bf = gd->n_tags > 0 ? gd->tag_offsets[gd->n_tags] : 0;
        for (int64_t k = 0; k < n_bf; k++) {
            const int64_t f = gd->tag_faces[k];
            const int64_t l = gd->face_left[f], r = gd->face_right[f];
            const int64_t ghost = l < n_cells ? r : l, valid = l < n_cells ? l : r;
            if (ghost >= n_cells && ghost < n_total && valid < n_cells) gkey[ghost] = (uint64_t)R.ref2int[valid];
        }
        std::vector<int> idx(R.int2ref.begin() + n_cells, R.int2ref.end());
        if (morton) radix_sort_by_key(gkey, idx);
        std::copy(idx.begin(), idx.end(), R.int2ref.begin() + n_cells);
        for (int64_t i = n_cells; i < n_total; i++) R.ref2int[R.int2ref[i]] = (int)i;
    }

    // ---- interfaces: numbered in order of first discovery over the new cell order ----
    std::vector<int64_t> fnew(n_faces, -1);
    R.cell_first_face.assign(n_grad + 1, 0);
    int64_t counter = 0;
    R.cf_off.assign(n_cells + 1, 0);
    for (int64_t c = 0; c < n_cells; c++) {
        const int64_t o = R.int2ref[c];
        R.cf_off[c + 1] = R.cf_off[c] + (gd->cell_face_offsets[o + 1] - gd->cell_face_offsets[o]);
    }
    for (int64_t c = 0; c < n_grad; c++) {
        const int64_t o = R.int2ref[c];
        R.cell_first_face[c] = counter;
        for (int64_t q = gd->cell_face_offsets[o]; q < gd->cell_face_offsets[o + 1]; q++) {
            const int64_t f = gd->cell_face_ids[q];
            if (f < 0 || f >= n_faces) { *err = "cell_face_ids out of range"; return IBIS_B200_EINVAL; }
            if (fnew[f] < 0) fnew[f] = counter++;
        }
    }
    R.cell_first_face[n_grad] = counter;
    for (int64_t f = 0; f < n_faces; f++)
        if (fnew[f] < 0) fnew[f] = counter++;       // interfaces no cell lists (not expected)
    R.face_ref2int.resize(n_faces);
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < n_faces; f++) R.face_ref2int[f] = (int)fnew[f];

    // ---- permuted copies ----
    const int64_t n_cf = R.cf_off[n_cells];
    R.face_left.resize(n_faces); R.face_right.resize(n_faces); R.f_area.resize(n_faces);
    R.f_norm.resize(3 * n_faces); R.f_t1.resize(3 * n_faces); R.f_t2.resize(3 * n_faces); R.f_c.resize(3 * n_faces);
    auto map_cell = [&](int64_t c) -> int64_t { return (c >= 0 && c < n_total) ? R.ref2int[c] : c; };
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < n_faces; f++) {
        const int64_t g = fnew[f];
        R.face_left[g] = map_cell(gd->face_left[f]);
        R.face_right[g] = map_cell(gd->face_right[f]);
        R.f_area[g] = gd->face_area[f];
        for (int d = 0; d < 3; d++) {
            R.f_norm[3 * g + d] = gd->face_norm[3 * f + d];
            R.f_t1[3 * g + d] = gd->face_tan1[3 * f + d];
            R.f_t2[3 * g + d] = gd->face_tan2[3 * f + d];
            R.f_c[3 * g + d] = gd->face_centre[3 * f + d];
        }
    }
    R.cf_ids.resize(n_cf); R.cf_nbr.resize(n_cf); R.cf_out.resize(n_cf);
    R.c_vol.resize(n_cells); R.c_cen.resize(3 * n_total);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < n_total; c++) {
        const int64_t o = R.int2ref[c];
        for (int d = 0; d < 3; d++) R.c_cen[3 * c + d] = gd->cell_centroid[3 * o + d];
        if (c >= n_cells) continue;
        R.c_vol[c] = gd->cell_volume[o];
        const int64_t src = gd->cell_face_offsets[o], dst = R.cf_off[c];
        const int64_t nf = gd->cell_face_offsets[o + 1] - src;
        for (int64_t j = 0; j < nf; j++) {
            R.cf_ids[dst + j] = fnew[gd->cell_face_ids[src + j]];
            R.cf_nbr[dst + j] = map_cell(gd->cell_neighbours[src + j]);
            R.cf_out[dst + j] = gd->cell_outsigns[src + j];
        }
    }
    const int64_t n_bf = gd->n_tags > 0 ? gd->tag_offsets[gd->n_tags] : 0;
    R.tag_faces.resize(n_bf);
    for (int64_t k = 0; k < n_bf; k++) R.tag_faces[k] = fnew[gd->tag_faces[k]];

    R.desc = *gd;
    R.desc.face_left = R.face_left.data();
    R.desc.face_right = R.face_right.data();
    R.desc.face_area = R.f_area.data();
    R.desc.face_norm = R.f_norm.data();
    R.desc.face_tan1 = R.f_t1.data();
    R.desc.face_tan2 = R.f_t2.data();
    R.desc.face_centre = R.f_c.data();
    R.desc.cell_face_offsets = R.cf_off.data();
    R.desc.cell_face_ids = R.cf_ids.data();
    R.desc.cell_outsigns = R.cf_out.data();
    R.desc.cell_neighbours = R.cf_nbr.data();
    R.desc.cell_volume = R.c_vol.data();
    R.desc.cell_centroid = R.c_cen.data();
    R.desc.tag_faces = R.tag_faces.data();
    R.desc.n_owned = n_owned;
    return 0;
}


// ---------------------------------------------------------------------------------------
// Tiles and the tile-owned interface numbering of the sides path.
//
// A tile is a run of at most T consecutive cells that does not straddle n_owned: tiles
// [0, nt_owned) cover the owned cells, [nt_owned, nt_all) the inner halo ring (cells with a face
// list).  Every interface is "homed" in the tile of the first cell (in cell order) that lists it.
// Inside the home range of tile b the interfaces are ordered in three sections,
//     [R-owned | both-owned | L-owned],
//   both-owned: the other cell is in tile b too, or is a boundary ghost (tile b writes both sides),
//   R-owned:    the other cell is elsewhere and tile b's cell is the RIGHT cell,
//   L-owned:    the other cell is elsewhere and tile b's cell is the LEFT cell,
// the outer two sorted by the tile of the other cell.  Consequences: the left sides tile b
// produces are the contiguous range [h0 + nR, h0 + nh), its right sides [h0, h0 + nR + n0), and
// the interfaces it shares with any one other tile are one contiguous range per orientation:
// every transfer of interface data of a tile, in either direction, is a handful of bulk copies.
// ---------------------------------------------------------------------------------------
struct TileEntry {        // a range of interfaces of tile t that are homed in another tile
    int f_start, count;
    int section;          // 0: the home tile owns the RIGHT side (t owns the left), 2: the reverse
    int slot_start;       // tile-local flux slot of the first interface (home slots come first)
    int stage_start;      // index of t's own side of the first interface in t's staging buffer
};
struct TileLayout {
    int T = 0, nt_owned = 0, nt_all = 0;
    std::vector<int> c0;          // nt_all + 1: cells of tile b = [c0[b], c0[b+1])
    std::vector<int> h0;          // nt_all + 1: home interfaces of tile b = [h0[b], h0[b+1])
    std::vector<int> nR, n0;      // sizes of the first two sections
    std::vector<int> ent_off;     // nt_all + 1 into entries
    std::vector<TileEntry> entries;
    std::vector<int> slot_other;  // per interface: its flux slot in the tile of the non-home cell, -1
    std::vector<int> home_tile;   // per interface
    int cap_slots = 1, cap_stage = 1;
    // staging layout of tile b: left sides first: [home left range | entries of section 0 ...],
    // then right sides: [home right range | entries of section 2 ...]
    int n_own_left(int b) const {
        int n = (h0[b + 1] - h0[b]) - nR[b];
        for (int e = ent_off[b]; e < ent_off[b + 1]; e++)
            if (entries[e].section == 0) n += entries[e].count;
        return n;
    }
};

inline int regroup_faces(Renumbered& R, int T, TileLayout& TL, const char** err) {
    const ibis_b200_grid_desc& d = R.desc;
    const int64_t n_cells = d.n_cells, n_faces = d.n_faces, n_grad = R.n_grad;
    const int64_t n_owned = d.n_owned > 0 ? d.n_owned : n_cells;
    TL.T = T;
    TL.nt_owned = (int)((n_owned + T - 1) / T);
    const int nt_halo = (int)((n_grad - n_owned + T - 1) / T);
    TL.nt_all = TL.nt_owned + nt_halo;
    const int nt = TL.nt_all;
    TL.c0.assign(nt + 1, 0);
    for (int b = 0; b <= TL.nt_owned; b++) TL.c0[b] = (int)std::min<int64_t>((int64_t)b * T, n_owned);
    for (int b = 0; b <= nt_halo; b++) TL.c0[TL.nt_owned + b] = (int)std::min<int64_t>(n_owned + (int64_t)b * T, n_grad);
    auto tile_of = [&](int64_t c) -> int {
        if (c < n_owned) return (int)(c / T);
        return TL.nt_owned + (int)((c - n_owned) / T);
    };
    const int BIG = 0x7fffffff;
    // home tile, section and sort key of every interface
    std::vector<uint64_t> key(n_faces);
    std::vector<int> home(n_faces, BIG), other_tile(n_faces, BIG), section(n_faces, 1);
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < n_faces; f++) {
        const int64_t l = R.face_left[f], r = R.face_right[f];
        const bool lv = l >= 0 && l < n_grad, rv = r >= 0 && r < n_grad;
        int hm = BIG, sec = 1, ot = BIG;
        if (lv || rv) {
            const bool home_is_left = lv && (!rv || l < r);
            const int64_t hc = home_is_left ? l : r, oc = home_is_left ? r : l;
            hm = tile_of(hc);
            const bool o_list = oc >= 0 && oc < n_grad;
            if (oc >= n_cells || (o_list && tile_of(oc) == hm)) {
                sec = 1;                                   // ghost, or both cells in the home tile
            } else {
                sec = home_is_left ? 2 : 0;
                ot = o_list ? tile_of(oc) : BIG - 1;       // outer halo ring: nobody produces that side
            }
        }
        home[f] = hm;
        section[f] = sec;
        other_tile[f] = ot;
    }
    std::vector<int> order(n_faces);
    for (int64_t f = 0; f < n_faces; f++) {
        order[f] = (int)f;
        const uint64_t hm = (uint64_t)std::min(home[f], 0xFFFFFF), ot = section[f] == 1 ? 0 : (uint64_t)std::min(other_tile[f], 0xFFFFFF);
        key[f] = hm << 34 | (uint64_t)section[f] << 32 | ot;
    }
    radix_sort_by_key(key, order);                         // stable: ties keep the discovery order
    std::vector<int64_t> fnew(n_faces);
    for (int64_t k = 0; k < n_faces; k++) fnew[order[k]] = k;
    // permute the interface arrays and everything that names an interface
    {
        std::vector<int64_t> fl(n_faces), fr(n_faces);
        std::vector<double> fa(n_faces), fn(3 * n_faces), ft1(3 * n_faces), ft2(3 * n_faces), fc(3 * n_faces);
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < n_faces; k++) {
            const int64_t f = order[k];
            fl[k] = R.face_left[f];
            fr[k] = R.face_right[f];
            fa[k] = R.f_area[f];
            for (int q = 0; q < 3; q++) {
                fn[3 * k + q] = R.f_norm[3 * f + q];
                ft1[3 * k + q] = R.f_t1[3 * f + q];
                ft2[3 * k + q] = R.f_t2[3 * f + q];
                fc[3 * k + q] = R.f_c[3 * f + q];
            }
        }
        R.face_left.swap(fl); R.face_right.swap(fr); R.f_area.swap(fa); R.f_norm.swap(fn);
        R.f_t1.swap(ft1); R.f_t2.swap(ft2); R.f_c.swap(fc);
    }
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)R.cf_ids.size(); q++) R.cf_ids[q] = fnew[R.cf_ids[q]];
    for (auto& f : R.tag_faces) f = fnew[f];
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < n_faces; f++) R.face_ref2int[f] = (int)fnew[R.face_ref2int[f]];
    R.desc.face_left = R.face_left.data(); R.desc.face_right = R.face_right.data();
    R.desc.face_area = R.f_area.data(); R.desc.face_norm = R.f_norm.data();
    R.desc.face_tan1 = R.f_t1.data(); R.desc.face_tan2 = R.f_t2.data(); R.desc.face_centre = R.f_c.data();
    R.desc.cell_face_ids = R.cf_ids.data(); R.desc.tag_faces = R.tag_faces.data();

    // per-tile section sizes, home ranges, entries
    TL.h0.assign(nt + 1, 0); TL.nR.assign(nt, 0); TL.n0.assign(nt, 0);
    TL.home_tile.assign(n_faces, BIG);
    std::vector<int> cnt(nt + 1, 0);
    for (int64_t k = 0; k < n_faces; k++) {
        const int f = order[k];
        TL.home_tile[k] = home[f];
        if (home[f] == BIG) continue;
        cnt[home[f]]++;
        if (section[f] == 0) TL.nR[home[f]]++;
        if (section[f] == 1) TL.n0[home[f]]++;
    }
    for (int b = 0; b < nt; b++) TL.h0[b + 1] = TL.h0[b] + cnt[b];
    // entries: maximal runs of equal (home, section != 1, other tile) in the new order
    std::vector<std::vector<TileEntry>> per_tile(nt);
    for (int64_t k = 0; k < n_faces;) {
        const int f = order[k];
        if (home[f] == BIG || section[f] == 1 || other_tile[f] >= nt) { k++; continue; }
        int64_t e = k + 1;
        while (e < n_faces && home[order[e]] == home[f] && section[order[e]] == section[f] &&
               other_tile[order[e]] == other_tile[f]) e++;
        TileEntry te;
        te.f_start = (int)k; te.count = (int)(e - k); te.section = section[f]; te.slot_start = 0; te.stage_start = 0;
        per_tile[other_tile[f]].push_back(te);
        k = e;
    }
    TL.ent_off.assign(nt + 1, 0);
    TL.slot_other.assign(n_faces, -1);
    for (int b = 0; b < nt; b++) {
        const int nh = TL.h0[b + 1] - TL.h0[b];
        int slot = nh;
        int stage_l = nh - TL.nR[b];                       // after the home left range
        int n_own_l = stage_l;
        for (auto& te : per_tile[b]) if (te.section == 0) n_own_l += te.count;
        int stage_r = n_own_l + TL.nR[b] + TL.n0[b];       // after the home right range
        for (auto& te : per_tile[b]) {
            te.slot_start = slot;
            for (int q = 0; q < te.count; q++) TL.slot_other[te.f_start + q] = slot + q;
            slot += te.count;
            if (te.section == 0) { te.stage_start = stage_l; stage_l += te.count; }
            else { te.stage_start = stage_r; stage_r += te.count; }
            TL.entries.push_back(te);
        }
        TL.ent_off[b + 1] = (int)TL.entries.size();
        TL.cap_slots = std::max(TL.cap_slots, slot);
        TL.cap_stage = std::max(TL.cap_stage, stage_r);
    }
    (void)err;
    return 0;
}

}  // namespace ib
