"""native_io.py -- the reference's native_binary flow snapshots as FILES, written and read through the snapshot
ring of the C ABI (include/ibis_b200.h: ibis_b200_snapshot_*): the device keeps stepping while a snapshot drains
to page-locked host memory on the download stream and a writer thread puts it on disk.

File layout = what NativeBinaryOutput<T>::write produces (io/native.cpp:186-259) and NativeBinaryInput<T>::read
consumes (native.cpp:280-339): `<plot_dir>/<NNNN>/{T,p,vx,vy[,vz]}`, each the raw float64 values of the valid cells in
reference numbering; `<plot_dir>/<NNNN>/meta_data.json` = {"time": t} dumped with indent 4; one line `NNNN` appended
to `<plot_dir>/flows`.  tests/test_flow_io.py compares the bytes with snapshots the reference itself wrote."""
import json
import os
import queue
import threading

import numpy as np

NAMES3 = ("T", "p", "vx", "vy", "vz")


def _meta_text(time):
    # nlohmann::json::dump(4) of {"time": t}: shortest round-trip representation of the double, like repr()
    t = float(time)
    txt = repr(t)
    if txt.endswith(".0") and "e" not in txt:
        pass                                    # nlohmann prints 85.0 as 85.0
    return "{\n    \"time\": " + txt + "\n}"


def write_arrays(plot_dir, index, arrays, time):
    """arrays: (4|5, n_cells) float64 in the order T p vx vy [vz] -> the reference's snapshot directory"""
    time_dir = f"{index:04d}"
    d = os.path.join(plot_dir, time_dir)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "meta_data.json"), "w") as f:
        f.write(_meta_text(time))
    with open(os.path.join(plot_dir, "flows"), "a") as f:
        f.write(time_dir + "\n")
    for q in range(arrays.shape[0]):
        with open(os.path.join(d, NAMES3[q]), "wb") as f:
            f.write(memoryview(np.ascontiguousarray(arrays[q])))


def read_arrays(plot_dir, index, n_cells, dim, out=None):
    """-> ((4|5, n_cells) float64, time); `out` (e.g. a page-locked array of the context) is filled in place"""
    d = os.path.join(plot_dir, f"{index:04d}")
    nv = 5 if dim == 3 else 4
    a = out if out is not None else np.empty((nv, n_cells))
    for q in range(nv):
        path = os.path.join(d, NAMES3[q])
        if os.path.getsize(path) != 8 * n_cells:
            raise ValueError(f"{path}: {os.path.getsize(path)} bytes, expected {8 * n_cells} (one float64 per valid cell)")
        with open(path, "rb") as f:
            f.readinto(memoryview(a[q]).cast("B"))
    with open(os.path.join(d, "meta_data.json")) as f:
        time = float(json.load(f)["time"])
    return a, time


class SnapshotWriter:
    """Writes snapshots of a FlowStates buffer without stopping the solver:

        w = SnapshotWriter(ctx, plot_dir)
        ... steps ...; w.write(fs, index, time); ... more steps (the download and the file write overlap them) ...
        w.close()

    Two page-locked host buffers alternate (the ring's two download slots): write() returns as soon as the pack
    kernel and the device-to-host copy are enqueued; a thread waits for the copy and writes the files."""

    def __init__(self, ctx, plot_dir):
        self.ctx, self.plot_dir = ctx, plot_dir
        nv = 5 if ctx.dim == 3 else 4
        self.host = [ctx.pinned_array((nv, ctx.n_cells)) for _ in range(2)]
        self.pending = [None, None]                    # per slot: a threading.Event set when its files are on disk
        self.turn = 0
        self.q = queue.Queue()
        self.error = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        os.makedirs(plot_dir, exist_ok=True)

    def _run(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            slot, index, time, done = job
            try:
                self.ctx.snapshot_download_end(slot)
                write_arrays(self.plot_dir, index, self.host[slot], time)
            except Exception as e:                       # noqa: BLE001  (surfaced by the next write() / close())
                self.error = e
            done.set()

    def write(self, fs, index, time):
        if self.error:
            raise self.error
        slot = self.turn
        self.turn ^= 1
        if self.pending[slot] is not None:
            self.pending[slot].wait()                    # the host buffer of this slot is on disk
        done = threading.Event()
        self.pending[slot] = done
        self.ctx.snapshot_download_begin(slot, fs, self.host[slot])
        self.q.put((slot, index, time, done))

    def close(self):
        self.q.put(None)
        self.thread.join()
        if self.error:
            raise self.error


def read_snapshot(ctx, fs, plot_dir, index, host=None):
    """NativeBinaryInput::read: files -> page-locked host arrays -> upload stream -> fs (+ update_thermo_from_pT).
    Returns the snapshot's time."""
    nv = 5 if ctx.dim == 3 else 4
    host = host if host is not None else ctx.pinned_array((nv, ctx.n_cells))
    _, time = read_arrays(plot_dir, index, ctx.n_cells, ctx.dim, out=host)
    ctx.snapshot_upload_begin(0, host)
    ctx.snapshot_upload_end(0, fs)
    ctx.sync()
    return time
