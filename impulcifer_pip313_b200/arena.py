"""Shared page-locked host arenas for the multi-GPU dispatcher ("results landing in one host arena in
set order").

One process per GPU is the launch model (``torch.distributed.run``), so "one host arena" is one POSIX shared-memory
segment that every rank maps and registers with its own CUDA context (``cudaHostRegister``): each rank DMAs the
recording sets it owns straight out of the shared INPUT arena and DMAs their cropped responses straight into its slots
of the shared RESULT arena.  The final result gather is therefore the D2H copy itself: no pickling, no inter-GPU
traffic, no collective; rank 0 (or any rank) reads every set in set order after the closing barrier.

This replaces the ``gather_object`` of pickled ``SetResult``s of round 1 (1.3 GB per rank per step through pickle).
"""
from __future__ import annotations

from multiprocessing import shared_memory

import numpy as np


class HostArena:
    """``nbytes`` of shared host memory, created by one rank (``create=True``) and attached by name by the others."""

    def __init__(self, name: str, nbytes: int, create: bool):
        self.name, self.nbytes, self.owner = name, int(nbytes), bool(create)
        if create:
            try:  # a crashed earlier run may have left the name behind
                old = shared_memory.SharedMemory(name=name)
                old.close()
                old.unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=max(self.nbytes, 1))
        else:
            self.shm = shared_memory.SharedMemory(name=name)
            try:  # Python < 3.13 registers attached segments with the resource tracker, which would unlink the
                from multiprocessing import resource_tracker  # owner's segment when this process exits
                resource_tracker.unregister(self.shm._name, 'shared_memory')
            except Exception:
                pass
        self._registered = False
        self._np = np.ndarray((self.nbytes,), dtype=np.uint8, buffer=self.shm.buf)

    def register(self) -> bool:
        """Page-lock the mapping for this process's CUDA context (async DMA at full PCIe rate).  No-op without CUDA."""
        import torch
        if self._registered or not torch.cuda.is_available() or self.nbytes == 0:
            return self._registered
        rc = torch.cuda.cudart().cudaHostRegister(self._np.ctypes.data, self.nbytes, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister({self.name}, {self.nbytes} B) failed with {rc}")
        self._registered = True
        return True

    def register_ranges(self, ranges) -> int:
        """Page-lock only the byte ranges ``[(offset, nbytes), ...]`` this process will DMA (the sets a rank owns):
        registering the whole arena in every rank pins it G times over and runs into the locked-memory limit
        (8 ranks x 23 GB failed on the 8-GPU box).  Ranges are widened to page boundaries and merged.  Returns the
        number of bytes registered."""
        import torch
        if not torch.cuda.is_available() or self.nbytes == 0:
            return 0
        page = 4096
        base = self._np.ctypes.data
        spans = sorted(((base + int(o)) // page * page, -(-(base + int(o) + int(n)) // page) * page) for o, n in ranges if n > 0)
        merged = []
        for a, b in spans:
            if merged and a <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], b)
            else:
                merged.append([a, b])
        end = -(-(base + self.nbytes) // page) * page
        total = 0
        self._ranges = getattr(self, "_ranges", [])
        for a, b in merged:
            b = min(b, end)
            rc = torch.cuda.cudart().cudaHostRegister(a, b - a, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister({self.name}, {b - a} B at +{a - base}) failed with {rc}")
            self._ranges.append(a)
            total += b - a
        return total

    def view(self, dtype, offset: int, shape):
        """torch tensor over ``shape`` elements of ``dtype`` starting ``offset`` bytes into the arena."""
        import torch
        count = int(np.prod(shape))
        arr = np.ndarray((count,), dtype=dtype, buffer=self.shm.buf, offset=int(offset))
        return torch.from_numpy(arr).view(*shape)

    def close(self):
        import torch
        if torch.cuda.is_available() and (self._registered or getattr(self, "_ranges", None)):
            torch.cuda.synchronize()
            if self._registered:
                torch.cuda.cudart().cudaHostUnregister(self._np.ctypes.data)
            for a in getattr(self, "_ranges", []):
                torch.cuda.cudart().cudaHostUnregister(a)
            self._registered, self._ranges = False, []
        self._np = None
        try:
            self.shm.close()
        except BufferError:  # tensor views still alive: the mapping goes with the process
            pass
        if self.owner:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass


class ResultArena:
    """Result slots of ``n_sets`` recording sets in set order inside a ``HostArena``:
    ``meta`` int64 [n_sets, 2 + 3 n_irs] = (tail_ind, n_irs, peaks.., crop.., knees..) and
    ``data`` float64 [n_sets, n_irs * max_tail] (set s keeps its ``n_irs x tail_ind`` matrix at the head of row s)."""

    def __init__(self, arena: HostArena, n_sets: int, n_irs: int, max_tail: int, offset: int = 0):
        self.n_sets, self.n_irs, self.max_tail = int(n_sets), int(n_irs), int(max_tail)
        self.meta_cols = 2 + 3 * self.n_irs
        meta_bytes = self.n_sets * self.meta_cols * 8
        self.meta = arena.view(np.int64, offset, (self.n_sets, self.meta_cols))
        self.data = arena.view(np.float64, offset + meta_bytes, (self.n_sets, self.n_irs * self.max_tail))
        self.data_offset = offset + meta_bytes
        self.slot_bytes = self.n_irs * self.max_tail * 8

    def slot_ranges(self, sets):
        """Byte ranges of the data slots of ``sets`` inside the arena (for ``HostArena.register_ranges``)."""
        return [(self.data_offset + int(s) * self.slot_bytes, self.slot_bytes) for s in sets]

    @staticmethod
    def nbytes(n_sets: int, n_irs: int, max_tail: int) -> int:
        return int(n_sets) * ((2 + 3 * int(n_irs)) + int(n_irs) * int(max_tail)) * 8

    def put_meta(self, s: int, result):
        row = self.meta[s]
        row[0], row[1] = int(result.tail_ind), len(result.names)
        n = self.n_irs
        row[2:2 + n] = torch_i64(result.peaks)
        row[2 + n:2 + 2 * n] = torch_i64(result.crop)
        row[2 + 2 * n:2 + 3 * n] = torch_i64(result.knees)

    def get(self, s: int):
        """(tail_ind, data [n_irs, tail_ind] view, peaks, crop, knees) of set ``s``."""
        row = self.meta[s]
        tail, n = int(row[0]), int(row[1])
        d = self.data[s, :n * tail].view(n, tail)
        k = self.n_irs
        return tail, d, row[2:2 + n], row[2 + k:2 + k + n], row[2 + 2 * k:2 + 2 * k + n]


def torch_i64(v):
    import torch
    return torch.as_tensor(np.asarray(v, dtype=np.int64))
