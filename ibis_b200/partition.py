"""Cell decomposition of a general unstructured GridBlock for the domain-decomposed runs
(no reference counterpart: the reference is a single block).

``partition_rcb`` assigns every valid cell to one of ``n_parts`` ranks by recursive coordinate
bisection of the cell centroids (balanced to within one cell, deterministic).  ``extract`` cuts the
sub-grid of one rank out of the global grid in the form ``ibis_b200_create`` expects:

    cells   [owned | halo ring 1 (face neighbours of owned cells; they keep their full face list so
             that their gradients, limiters and interface sides can be recomputed locally) |
             halo ring 2 (face neighbours of ring 1; values only, no face list) | boundary ghosts]
    faces   every interface of an owned or ring-1 cell, with the global orientation, frame and
            per-cell face order untouched -- which is what makes the per-cell arithmetic, hence the
            residual, independent of the decomposition, bit for bit
    ghosts  one per local boundary interface, numbered after the cells in (alphabetical tag, global
            marker order), the reference's convention (grid/grid.h:414-433)

together with the exchange lists for ``ibis_b200_halo_setup``: per peer rank the owned cells to send
and the halo cells to receive, both in ascending GLOBAL cell id so that the two sides agree without
talking to each other.  Host-side setup code: numpy, runs once."""
import numpy as np

from .grid import Grid


def partition_rcb(g, n_parts):
    """owner[n_cells]: rank of every valid cell.  Recursive coordinate bisection: split the longest
    axis of the current set at the cell-count median (ties broken by global id), recurse with the
    rank counts floor/ceil."""
    n = int(g.n_cells)
    cen = np.asarray(g.cell_centroid)[:n]
    owner = np.zeros(n, np.int32)

    def split(ids, r0, nr):
        if nr == 1:
            owner[ids] = r0
            return
        n_lo = nr // 2
        ext = cen[ids].max(axis=0) - cen[ids].min(axis=0)
        axis = int(np.argmax(ext[: int(g.dim)]))
        order = ids[np.lexsort((ids, cen[ids, axis]))]
        k = len(ids) * n_lo // nr
        split(order[:k], r0, n_lo)
        split(order[k:], r0 + n_lo, nr - n_lo)

    split(np.arange(n, dtype=np.int64), 0, int(n_parts))
    return owner


class SubGrid(Grid):
    """grid.Grid of one rank + maps back to the global grid + halo exchange lists."""


def extract(g, owner, rank):
    n_cells_g = int(g.n_cells)
    off = np.asarray(g.cell_face_offsets, np.int64)
    nbr = np.asarray(g.cell_neighbours, np.int64)
    fids = np.asarray(g.cell_face_ids, np.int64)
    outs = np.asarray(g.cell_outsigns)
    owner = np.asarray(owner)

    def valid_neighbours(cells):
        """global ids of the valid (non-ghost) face neighbours of `cells`"""
        if len(cells) == 0:
            return np.zeros(0, np.int64)
        idx = np.concatenate([np.arange(off[c], off[c + 1]) for c in cells]) if len(cells) < 64 else \
            _ranges(off[cells], off[cells + 1])
        nb = nbr[idx]
        return np.unique(nb[(nb >= 0) & (nb < n_cells_g)])

    owned = np.nonzero(owner == rank)[0].astype(np.int64)
    is_owned = np.zeros(n_cells_g, bool)
    is_owned[owned] = True
    ring1 = valid_neighbours(owned)
    ring1 = ring1[~is_owned[ring1]]
    in12 = is_owned.copy()
    in12[ring1] = True
    ring2 = valid_neighbours(ring1)
    ring2 = ring2[~in12[ring2]]
    cells = np.concatenate([owned, ring1, ring2])            # local -> global (valid cells)
    n_owned, n_face_cells, n_cells = len(owned), len(owned) + len(ring1), len(cells)
    g2l = np.full(n_cells_g + int(g.n_ghost), -1, np.int64)
    g2l[cells] = np.arange(n_cells)

    # interfaces of the cells that keep a face list, in ascending global id
    fc = cells[:n_face_cells]
    cf_idx = _ranges(off[fc], off[fc + 1])
    faces = np.unique(fids[cf_idx])
    f_g2l = np.full(int(g.n_faces), -1, np.int64)
    f_g2l[faces] = np.arange(len(faces))

    # local boundary interfaces and ghosts, in (alphabetical tag, global marker order)
    tags, n_ghost = {}, 0
    for name in sorted(g.tags):
        tf = np.asarray(g.tags[name]["faces"], np.int64)
        tg = np.asarray(g.tags[name]["ghosts"], np.int64)
        keep = f_g2l[tf] >= 0
        loc_ghost = n_cells + n_ghost + np.arange(int(keep.sum()))
        g2l[tg[keep]] = loc_ghost
        tags[name] = dict(faces=f_g2l[tf[keep]], ghosts=loc_ghost, global_ghosts=tg[keep])
        n_ghost += int(keep.sum())

    s = SubGrid()
    s.dim, s.n_cells, s.n_ghost, s.n_faces, s.n_owned = int(g.dim), n_cells, n_ghost, len(faces), n_owned
    s.tags = tags
    s.face_left = g2l[np.asarray(g.face_left, np.int64)[faces]]
    s.face_right = g2l[np.asarray(g.face_right, np.int64)[faces]]
    if (s.face_left < 0).any() or (s.face_right < 0).any():
        raise ValueError("partition: an interface of a ring-1 cell has a cell outside the two-ring halo")
    for k in ("face_area", "face_norm", "face_tan1", "face_tan2", "face_centre"):
        setattr(s, k, np.ascontiguousarray(np.asarray(getattr(g, k))[faces]))
    counts = np.zeros(n_cells, np.int64)
    counts[:n_face_cells] = off[fc + 1] - off[fc]
    s.cell_face_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    s.cell_face_ids = f_g2l[fids[cf_idx]]
    s.cell_outsigns = np.ascontiguousarray(outs[cf_idx])
    s.cell_neighbours = g2l[nbr[cf_idx]]
    if (s.cell_neighbours < 0).any():
        raise ValueError("partition: a neighbour of a ring-1 cell is outside the two-ring halo")
    s.cell_volume = np.ascontiguousarray(np.asarray(g.cell_volume)[cells])
    loc2glob_all = np.concatenate([cells] + [tags[t]["global_ghosts"] for t in sorted(tags)]).astype(np.int64)
    s.cell_centroid = np.ascontiguousarray(np.asarray(g.cell_centroid)[loc2glob_all])
    s.local_to_global = loc2glob_all                           # cells then ghosts
    s.face_local_to_global = faces

    # exchange lists: peers in ascending rank, cells in ascending global id on both sides
    halo = cells[n_owned:]
    halo_owner = owner[halo]
    peers = sorted(set(int(r) for r in halo_owner))
    s.peer_ranks = peers
    recv, recv_off = [], [0]
    for r in peers:
        ids = np.sort(halo[halo_owner == r])
        recv.append(g2l[ids])
        recv_off.append(recv_off[-1] + len(ids))
    s.recv_cells = np.concatenate(recv).astype(np.int64) if recv else np.zeros(0, np.int64)
    s.recv_offsets = np.asarray(recv_off, np.int64)
    s._owner, s._rank, s._g2l = owner, rank, g2l
    s.send_cells, s.send_offsets = None, None                  # filled by connect_lists (needs the peers' halos)
    return s


def connect_lists(subs):
    """Fill the send lists of every sub-grid from the receive lists of its peers (in a real run each
    rank does this after an all-gather of the halo ids; here all sub-grids are at hand)."""
    for s in subs:
        send, send_off = [], [0]
        for r in s.peer_ranks:
            o = subs[r]
            q = o.peer_ranks.index(s._rank)
            wanted_local_on_o = o.recv_cells[o.recv_offsets[q]:o.recv_offsets[q + 1]]
            wanted_global = o.local_to_global[wanted_local_on_o]
            send.append(s._g2l[wanted_global])
            send_off.append(send_off[-1] + len(wanted_global))
        s.send_cells = np.concatenate(send).astype(np.int64) if send else np.zeros(0, np.int64)
        s.send_offsets = np.asarray(send_off, np.int64)
        assert (s.send_cells >= 0).all() and (s.send_cells < s.n_owned).all()
    # symmetry: whoever I receive from also receives from me (two-ring halos are symmetric)
    for s in subs:
        for r in s.peer_ranks:
            assert s._rank in subs[r].peer_ranks
    return subs


def decompose(g, n_parts):
    """owner map + the sub-grids of all ranks with their exchange lists."""
    owner = partition_rcb(g, n_parts)
    return owner, connect_lists([extract(g, owner, r) for r in range(n_parts)])


def _ranges(starts, ends):
    """concatenate(arange(s, e) for s, e in zip(starts, ends)) without the Python loop"""
    starts, ends = np.asarray(starts, np.int64), np.asarray(ends, np.int64)
    lens = ends - starts
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    out = np.ones(total, np.int64)
    pos = np.concatenate([[0], np.cumsum(lens)[:-1]])
    nz = lens > 0
    out[pos[nz]] = starts[nz] - np.concatenate([[0], (ends[nz] - 1)[:-1]])
    out[0] = starts[nz][0]
    return np.cumsum(out)
