"""A minimal pure-Python reader for the HDF5 subset the reference's bundled fixture `aflw2kmini.h5` uses (no h5py in this
image): superblock v0, 8-byte offsets, root group as a symbol table (B-tree v1 + local heap + SNOD), version-1 object
headers, chunked datasets with a single unfiltered chunk (layout v3, B-tree v1 chunk index), fixed-point / IEEE float
types and variable-length uint8 sequences in global heaps (GCOL).  Used by make_golden_aflw2kmini.py only."""
import struct

import numpy as np


class File:
    def __init__(self, path):
        self.d = open(path, "rb").read()
        d = self.d
        assert d[:8] == b"\x89HDF\r\n\x1a\n" and d[8] == 0 and d[13] == 8 and d[14] == 8, "only superblock v0 with 8-byte offsets"
        btree, heap = struct.unpack_from("<QQ", d, 56 + 24)
        self.datasets = {name: addr for name, addr in self._group(btree, heap)}

    # -- groups ---------------------------------------------------------------------------------------------
    def _heap_data(self, heap):
        d = self.d
        assert d[heap:heap + 4] == b"HEAP"
        size, free, addr = struct.unpack_from("<QQQ", d, heap + 8)
        return addr

    def _group(self, btree, heap):
        d = self.d
        names = self._heap_data(heap)
        assert d[btree:btree + 4] == b"TREE"
        node_type, level, used = struct.unpack_from("<BBH", d, btree + 4)
        assert node_type == 0
        p = btree + 8 + 16  # siblings
        children = []
        for i in range(used):
            key, child = struct.unpack_from("<QQ", d, p)
            p += 16
            children.append(child)
        for c in children:
            if level > 0:
                yield from self._group_node(c, names)
            else:
                yield from self._snod(c, names)

    def _group_node(self, addr, names):
        d = self.d
        node_type, level, used = struct.unpack_from("<BBH", d, addr + 4)
        p = addr + 8 + 16
        for i in range(used):
            key, child = struct.unpack_from("<QQ", d, p)
            p += 16
            if level > 0:
                yield from self._group_node(child, names)
            else:
                yield from self._snod(child, names)

    def _snod(self, addr, names):
        d = self.d
        assert d[addr:addr + 4] == b"SNOD"
        n, = struct.unpack_from("<H", d, addr + 6)
        p = addr + 8
        for i in range(n):
            name_off, obj = struct.unpack_from("<QQ", d, p)
            s = names + name_off
            name = d[s:d.index(b"\0", s)].decode()
            yield name, obj
            p += 40

    # -- object headers (version 1) -------------------------------------------------------------------------
    def _messages(self, addr):
        d = self.d
        assert d[addr] == 1
        nmsg, = struct.unpack_from("<H", d, addr + 2)
        hsize, = struct.unpack_from("<I", d, addr + 8)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", d, p)
                body = p + 8
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", d, body)
                    blocks.append((caddr, clen))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    def read(self, name):
        d = self.d
        shape = dtype = None
        chunk_btree = None
        for mtype, body, msize in self._messages(self.datasets[name]):
            if mtype == 0x01:  # dataspace
                ver, rank, flags = struct.unpack_from("<BBB", d, body)
                off = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<%dQ" % rank, d, off)
            elif mtype == 0x03:  # datatype
                cls = d[body] & 0x0F
                size, = struct.unpack_from("<I", d, body + 4)
                dtype = (cls, size)
            elif mtype == 0x08:  # layout
                ver, lclass = d[body], d[body + 1]
                assert ver == 3 and lclass == 2, "only chunked layout v3"
                rank = d[body + 2]
                chunk_btree, = struct.unpack_from("<Q", d, body + 3)
                chunk_dims = struct.unpack_from("<%dI" % rank, d, body + 11)
        assert shape is not None and dtype is not None and chunk_btree is not None, name
        # single chunk: first child of the chunk B-tree
        assert d[chunk_btree:chunk_btree + 4] == b"TREE" and d[chunk_btree + 4] == 1 and d[chunk_btree + 5] == 0
        used, = struct.unpack_from("<H", d, chunk_btree + 6)
        assert used == 1, "only single-chunk datasets"
        rank1 = len(chunk_dims)
        p = chunk_btree + 8 + 16
        csize, fmask = struct.unpack_from("<II", d, p)
        assert fmask == 0
        p += 8 + 8 * rank1
        data, = struct.unpack_from("<Q", d, p)
        cls, size = dtype
        n = int(np.prod(shape))
        if cls == 1:  # float
            return np.frombuffer(d, dtype={4: "<f4", 8: "<f8"}[size], count=n, offset=data).reshape(shape).copy()
        if cls == 0:  # fixed point
            return np.frombuffer(d, dtype={1: "u1", 2: "<i2", 4: "<i4", 8: "<i8"}[size], count=n, offset=data).reshape(shape).copy()
        if cls == 9:  # variable length: 16-byte references (length u32, heap address u64, index u32)
            out = []
            for i in range(n):
                length, haddr, idx = struct.unpack_from("<IQI", d, data + 16 * i)
                out.append(self._gcol(haddr, idx)[:length])
            return out
        raise NotImplementedError(f"datatype class {cls}")

    def _gcol(self, addr, index):
        d = self.d
        assert d[addr:addr + 4] == b"GCOL"
        size, = struct.unpack_from("<Q", d, addr + 8)
        p, end = addr + 16, addr + size
        while p + 16 <= end:
            idx, refs, _, osize = struct.unpack_from("<HHIQ", d, p)
            if idx == index:
                return np.frombuffer(d, dtype=np.uint8, count=osize, offset=p + 16).copy()
            if idx == 0:
                break
            p += 16 + ((osize + 7) & ~7)
        raise KeyError(index)
