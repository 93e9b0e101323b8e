"""A minimal reader for the kind of HDF5 file RedClust.jl ships (data/example_datasets.h5, read by
/root/reference/src/example_data.jl:33-71 through HDF5.jl): superblock version 0, version-1 object headers, old-style
groups (symbol-table B-tree + local heap) and compact new-style groups (hard-link messages in the object header, which is
what HDF5.jl writes for these few entries), contiguous little-endian integer / IEEE float datasets.  No libhdf5, no
h5py (neither is available where this runs).  Anything else in a file raises NotImplementedError instead of guessing.
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.raw = f.read()
        if self.raw[:8] != _SIG:
            raise ValueError(f"{path} is not an HDF5 file")
        if self.raw[8] != 0:
            raise NotImplementedError(f"HDF5 superblock version {self.raw[8]} (only version 0 is supported)")
        if self.raw[13] != 8 or self.raw[14] != 8:
            raise NotImplementedError("only 8-byte offsets / lengths are supported")
        # root group symbol-table entry follows the four superblock addresses (base, free space, end of file, driver)
        self.root = self._entry(24 + 4 * 8)

    # -- low level ----------------------------------------------------------------------------------------------
    def _u(self, off, size):
        return int.from_bytes(self.raw[off:off + size], "little")

    def _entry(self, off):
        """Symbol-table entry: link-name offset, object-header address, cache type, scratch-pad."""
        name_off, ohdr, cache = self._u(off, 8), self._u(off + 8, 8), self._u(off + 16, 4)
        btree = heap = None
        if cache == 1:
            btree, heap = self._u(off + 24, 8), self._u(off + 32, 8)
        return dict(name_off=name_off, ohdr=ohdr, btree=btree, heap=heap)

    def _messages(self, addr):
        """(type, payload offset, size) of every message of a version-1 object header, continuations included."""
        if self.raw[addr] != 1:
            raise NotImplementedError(f"object header version {self.raw[addr]} (only version 1 is supported)")
        nmsg, size = self._u(addr + 2, 2), self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            off, left = blocks.pop(0)
            end = off + left
            while off + 8 <= end and len(out) < nmsg:
                mtype, msize = self._u(off, 2), self._u(off + 2, 2)
                payload = off + 8
                if mtype == 0x10:                                     # continuation: (offset, length) of another block
                    blocks.append((self._u(payload, 8), self._u(payload + 8, 8)))
                out.append((mtype, payload, msize))
                off = payload + msize
        return out

    def _link(self, off):
        """Version-1 link message -> (name, object-header address); hard links only."""
        if self.raw[off] != 1:
            raise NotImplementedError("link message version")
        flags = self.raw[off + 1]
        p = off + 2
        if flags & 0x08:
            if self.raw[p] != 0:
                raise NotImplementedError("soft / external links are not supported")
            p += 1
        if flags & 0x04:
            p += 8                                                    # creation order
        if flags & 0x10:
            p += 1                                                    # character set
        lsz = 1 << (flags & 3)
        ln = self._u(p, lsz); p += lsz
        name = self.raw[p:p + ln].decode(); p += ln
        return name, self._u(p, 8)

    def _heap_name(self, heap, off):
        if self.raw[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        data = self._u(heap + 24, 8)
        end = self.raw.index(b"\x00", data + off)
        return self.raw[data + off:end].decode()

    def _group_entries(self, btree, heap):
        """name -> symbol-table entry for an old-style group: walk the version-1 B-tree down to its SNOD leaves."""
        out = {}
        if self.raw[btree:btree + 4] != b"TREE" or self.raw[btree + 4] != 0:
            raise ValueError("bad group B-tree node")
        level, used = self.raw[btree + 5], self._u(btree + 6, 2)
        off = btree + 24                                              # past signature, type, level, count, two siblings
        for k in range(used):
            child = self._u(off + 8 + k * 16, 8)                      # key_k (8 bytes) child_k (8 bytes) key_k+1 ...
            if level > 0:
                out.update(self._group_entries(child, heap))
                continue
            if self.raw[child:child + 4] != b"SNOD":
                raise ValueError("bad symbol-table node")
            for e in range(self._u(child + 6, 2)):
                ent = self._entry(child + 8 + e * 40)
                out[self._heap_name(heap, ent["name_off"])] = ent
        return out

    # -- objects --------------------------------------------------------------------------------------------------
    def _open(self, entry):
        btree, heap = entry["btree"], entry["heap"]
        msgs = self._messages(entry["ohdr"])
        if btree is None:
            for mtype, off, _ in msgs:
                if mtype == 0x11:                                     # symbol-table message: this object is a group
                    btree, heap = self._u(off, 8), self._u(off + 8, 8)
        links, is_group = {}, btree is not None
        for mtype, off, _ in msgs:
            if mtype == 0x02:                                         # link info: a new-style group
                is_group = True
                if self._u(off + 2 + (8 if self.raw[off + 1] & 1 else 0), 8) != 0xFFFFFFFFFFFFFFFF:
                    raise NotImplementedError("dense link storage (fractal heap) is not supported")
            elif mtype == 0x06:                                       # link message
                is_group = True
                name, addr = self._link(off)
                links[name] = dict(name_off=0, ohdr=addr, btree=None, heap=None)
        if btree is not None:
            links.update(self._group_entries(btree, heap))
        if is_group:
            return H5Group(self, links)
        return self._dataset(msgs)

    def _dataset(self, msgs):
        dims = dtype = addr = nbytes = None
        for mtype, off, size in msgs:
            if mtype == 0x01:                                         # dataspace
                ver, rank, flags = self.raw[off], self.raw[off + 1], self.raw[off + 2]
                start = off + (8 if ver == 1 else 4)
                dims = [self._u(start + 8 * d, 8) for d in range(rank)]
            elif mtype == 0x03:                                       # datatype
                cls, bits0, tsize = self.raw[off] & 0x0f, self.raw[off + 1], self._u(off + 4, 4)
                if bits0 & 1:
                    raise NotImplementedError("big-endian data")
                if cls == 0:
                    dtype = np.dtype(f"<{'i' if bits0 & 8 else 'u'}{tsize}")
                elif cls == 1:
                    dtype = np.dtype(f"<f{tsize}")
                else:
                    raise NotImplementedError(f"datatype class {cls}")
            elif mtype == 0x08:                                       # data layout
                ver = self.raw[off]
                if ver == 3:
                    if self.raw[off + 1] != 1:
                        raise NotImplementedError("only contiguous datasets are supported")
                    addr, nbytes = self._u(off + 2, 8), self._u(off + 10, 8)
                else:
                    raise NotImplementedError(f"data layout message version {ver}")
        if dims is None or dtype is None or addr is None:
            raise ValueError("dataset without dataspace / datatype / layout")
        count = int(np.prod(dims)) if dims else 1
        if nbytes != count * dtype.itemsize:
            raise ValueError("dataset size does not match its dataspace")
        return np.frombuffer(self.raw, dtype, count, addr).reshape(dims).copy()

    def __getitem__(self, name):
        return self._open(self.root)[name]

    def keys(self):
        return self._open(self.root).keys()

    def close(self):
        self.raw = b""


class H5Group:
    def __init__(self, f, entries):
        self._f, self._entries = f, entries

    def keys(self):
        return sorted(self._entries)

    def __getitem__(self, name):
        """A nested group, or the dataset as a numpy array in HDF5 (row-major) dimension order -- a Julia array of size
        (a, b) is stored with dimensions (b, a), i.e. this returns its transpose."""
        if name not in self._entries:
            raise KeyError(name)
        return self._f._open(self._entries[name])
