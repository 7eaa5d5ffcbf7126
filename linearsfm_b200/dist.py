"""Multi-GPU sharding of the merge tree: one process per GPU, torch.distributed for
the plumbing.

The joins of one tree level are independent (LinearSFMImp.cpp:1938-2033 touches only maps 2i, 2i+1),
so rank g owns the contiguous, power-of-two aligned leaf slice [g*c, (g+1)*c) and runs log2(c)
levels with NO communication.  Above that, level by level, the owner of map 2i+1 hands its whole map
to the owner of map 2i (one point-to-point transfer per pair and level), which joins them.  There
is no collective on the data path.  The re-base rule of the scheduler depends on the GLOBAL output
index (LinearSFMImp.cpp:1997), which is why every local call passes `first_index`.

`run_sharded` is backend-agnostic: the GPU backend is `api.Tree`; the CPU tests drive the same
schedule with an oracle-based backend over gloo.
"""
from __future__ import annotations

import numpy as np

from .localmap import LocalMap


def plan(num_maps: int, world: int):
    """Returns (chunk, local_levels, owners) : chunk = leaves per rank (power of two), local_levels =
    log2(chunk), owners = ranks holding a map after the local phase (in map order)."""
    per = -(-num_maps // world)
    chunk = 1
    while chunk < per:
        chunk *= 2
    nlocal = chunk.bit_length() - 1
    nonempty = -(-num_maps // chunk)
    return chunk, nlocal, list(range(nonempty))


def slice_of(num_maps: int, world: int, rank: int):
    chunk, _, _ = plan(num_maps, world)
    lo = min(rank * chunk, num_maps)
    hi = min(lo + chunk, num_maps)
    return lo, hi


_INT_FIELDS = ("stno", "Ui", "Uj", "photo", "feature", "FBlock")
_DBL_FIELDS = ("stVal", "U", "W", "V")


def pack_map(lm: LocalMap):
    """LocalMap -> (header int64[8], ints int32[...], doubles float64[...])"""
    head = np.array([lm.Ref, lm.FRef, lm.m, lm.n, lm.nU, lm.nW, 0, 0], np.int64)
    ints = np.concatenate([np.asarray(getattr(lm, f), np.int32).reshape(-1) for f in _INT_FIELDS])
    dbls = np.concatenate([np.asarray(getattr(lm, f), np.float64).reshape(-1) for f in _DBL_FIELDS])
    head[6], head[7] = ints.shape[0], dbls.shape[0]
    return head, ints, dbls


def unpack_map(head, ints, dbls) -> LocalMap:
    Ref, FRef, m, n, nU, nW = (int(x) for x in head[:6])
    r = 6 * m + 3 * n
    isz = dict(stno=r, Ui=nU, Uj=nU, photo=nW, feature=nW, FBlock=n)
    dsz = dict(stVal=r, U=36 * nU, W=18 * nW, V=9 * n)
    kw, o = {}, 0
    for f in _INT_FIELDS:
        kw[f] = ints[o:o + isz[f]].copy(); o += isz[f]
    o = 0
    for f in _DBL_FIELDS:
        kw[f] = dbls[o:o + dsz[f]].copy(); o += dsz[f]
    return LocalMap(Ref=Ref, FRef=FRef, m=m, n=n, **kw)


def send_map(lm: LocalMap, dst: int, device=None):
    import torch
    import torch.distributed as dist
    for a in pack_map(lm):
        t = torch.from_numpy(np.ascontiguousarray(a))
        if device is not None:
            t = t.to(device)
        if a.dtype == np.int64:
            dist.send(t, dst)
        elif t.numel() > 0:
            dist.send(t, dst)


def recv_map(src: int, device=None) -> LocalMap:
    import torch
    import torch.distributed as dist
    head = torch.zeros(8, dtype=torch.int64, device=device or "cpu")
    dist.recv(head, src)
    h = head.cpu().numpy()
    ints = torch.zeros(int(h[6]), dtype=torch.int32, device=device or "cpu")
    dbls = torch.zeros(int(h[7]), dtype=torch.float64, device=device or "cpu")
    if ints.numel() > 0:
        dist.recv(ints, src)
    if dbls.numel() > 0:
        dist.recv(dbls, src)
    return unpack_map(h, ints.cpu().numpy(), dbls.cpu().numpy())


def run_sharded(backend, num_maps: int, rank: int, world: int, device=None):
    """Run the sharded merge tree.  `backend` already holds this rank's leaf slice and offers
        solve_levels(first_index, levels)   run `levels` levels on the maps it holds
        count()                             number of maps it holds
        get(i) -> LocalMap                  download map i
        set(list[LocalMap])                 replace the held maps
        append(list[LocalMap])              add maps after the held ones
        finish()                            final re-base of the single root map
    Returns True on the rank that holds the final map (rank 0)."""
    chunk, nlocal, owners = plan(num_maps, world)
    lo, hi = slice_of(num_maps, world, rank)
    if hi > lo and nlocal > 0:
        backend.solve_levels(lo, nlocal)
    # cross-rank levels: map j of the current level lives on rank owners[j]
    while len(owners) > 1:
        nxt = []
        for i in range(0, len(owners), 2):
            if i + 1 < len(owners):
                a, b = owners[i], owners[i + 1]
                if rank == b:
                    if hasattr(backend, "send_first"):
                        backend.send_first(a, device)          # device-to-device (NCCL)
                    else:
                        send_map(backend.get(0), a, device)
                    backend.set([])
                elif rank == a:
                    if hasattr(backend, "recv_append"):
                        backend.recv_append(b, device)
                    else:
                        backend.append([recv_map(b, device)])
                    backend.solve_levels(i, 1)
            else:
                a = owners[i]
                if rank == a:
                    backend.solve_levels(i, 1)       # leftover: only the re-base rule applies
            nxt.append(owners[i])
        owners = nxt
    if rank == owners[0]:
        backend.finish()
        return True
    return False


class TreeBackend:
    """GPU backend: maps stay resident in HBM inside an api.Tree between levels."""

    def __init__(self, api, maps):
        self.api = api
        self.tree = api.Tree(maps)
        self.n = len(maps)

    def solve_levels(self, first_index, levels):
        self.tree.solve(first_index=first_index, max_levels=levels)
        self.n = self.tree.result_count()
        self.tree.adopt_result()

    def count(self):
        return self.n

    def get(self, i):
        # the held maps are the adopted result: download through a zero-level solve
        self.tree.solve(first_index=0, max_levels=0)
        lm = self.tree.download(i)
        self.tree.adopt_result()
        return lm

    def set(self, maps):
        if maps:
            self.tree.set_maps(maps)       # replaces the uploaded leaf set
        self.n = len(maps)                 # empty: the rank is done, keep its uploaded leaves resident

    def append(self, maps):
        self.tree.append_maps(maps)
        self.n += len(maps)

    def send_first(self, dst, device):
        """Hand the single held map to rank `dst` without leaving device memory: pack it into one
        contiguous buffer (D2D copies on our stream) and NCCL-send the raw bytes."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from ._lib import lib
        self.tree.solve(first_index=0, max_levels=0)          # input set -> result set (no work)
        s = self.tree.result_shape(0)
        nbytes = int(lib().lsfm_map_device_bytes(C.byref(s)))
        head = torch.tensor([s.Ref, s.FRef, s.m, s.n, s.nU, s.nW, nbytes, 0], dtype=torch.int64, device=device)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.tree.export_device(0, buf.data_ptr(), nbytes)
        dist.send(head, dst)
        dist.send(buf, dst)
        torch.cuda.synchronize(device)
        self.tree.adopt_result()

    def recv_append(self, src, device):
        import torch
        import torch.distributed as dist
        from ._lib import LsfmMap
        head = torch.zeros(8, dtype=torch.int64, device=device)
        dist.recv(head, src)
        h = [int(x) for x in head.cpu().tolist()]
        buf = torch.empty(h[6], dtype=torch.uint8, device=device)
        dist.recv(buf, src)
        torch.cuda.synchronize(device)
        shape = LsfmMap(Ref=h[0], FRef=h[1], m=h[2], n=h[3], nU=h[4], nW=h[5], r=6 * h[2] + 3 * h[3])
        self.tree.append_device(shape, buf.data_ptr(), h[6])
        self.n += 1

    def finish(self):
        self.tree.solve(first_index=0, max_levels=-1)
        self.n = self.tree.result_count()
