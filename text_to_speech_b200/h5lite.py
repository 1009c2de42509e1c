"""A minimal read-only HDF5 parser -- enough for Keras `.weights.h5` files (h5py is not available here).

The reference restores a model with `model.load_weights('<name>.weights.h5')`
(`custom_train_objects/checkpoint_manager.py:169-216`); Keras 3 writes that file through h5py with the library's
default (earliest) format: superblock version 0, old-style groups (symbol-table message -> v1 B-tree + local heap),
one contiguous, uncompressed, little-endian dataset per variable. That subset of the HDF5 file-format specification
(version 1.1 / 3.0 documents: superblock v0-v3, object headers v1/v2, symbol-table and compact link-message groups,
dataspace v1/v2, fixed-point / floating-point datatypes, compact / contiguous layouts, chunked layout WITHOUT filters
through the v1 chunk B-tree) is what `H5File` reads. Anything else (dense link storage in fractal heaps, filters,
variable-length or compound types) raises `H5FormatError` with the name of the missing feature.

    with H5File(path) as f:
        for name, arr in f.datasets().items(): ...      # {'layers/conv1d/vars/0': ndarray, ...}
"""
from __future__ import annotations

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"


class H5FormatError(ValueError):
    pass


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        self.path = path
        self.base = self._find_superblock()
        self._parse_superblock()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.buf = b""

    # ---- low level ------------------------------------------------------------------------------------------------
    def _u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], "little")

    def _find_superblock(self):
        off = 0
        while off + 8 <= len(self.buf):              # the signature sits at 0 or at 512 * 2^k (user block)
            if self.buf[off:off + 8] == SIGNATURE:
                return off
            off = 512 if off == 0 else off * 2
        raise H5FormatError(f"{self.path}: not an HDF5 file (no superblock signature)")

    def _parse_superblock(self):
        b = self.base
        ver = self.buf[b + 8]
        if ver in (0, 1):
            self.so, self.sl = self.buf[b + 13], self.buf[b + 14]
            p = b + 24 + (4 if ver == 1 else 0)
            # base address, free-space address, end of file, driver info, then the root group's symbol-table entry
            p += 4 * self.so
            self.root_header = self._u(p + self.so, self.so) + self.base
        elif ver in (2, 3):
            self.so, self.sl = self.buf[b + 9], self.buf[b + 10]
            p = b + 12
            self.root_header = self._u(p + 3 * self.so, self.so) + self.base
        else:
            raise H5FormatError(f"unsupported superblock version {ver}")
        if self.so not in (4, 8) or self.sl not in (4, 8):
            raise H5FormatError(f"unsupported offset/length sizes {self.so}/{self.sl}")

    # ---- object headers ---------------------------------------------------------------------------------------------
    def _messages(self, addr):
        """[(type, flags, payload offset, payload size)] of the object header at `addr` (v1 or v2, continuations followed)."""
        out = []
        if self.buf[addr:addr + 4] == b"OHDR":
            flags = self.buf[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16                      # access / modification / change / birth times
            if flags & 0x10:
                p += 4                       # max compact / min dense attributes
            nsz = 1 << (flags & 3)
            chunk0 = self._u(p, nsz)
            p += nsz
            track = bool(flags & 0x04)
            blocks = [(p, chunk0)]
            while blocks:
                p, size = blocks.pop(0)
                end = p + size
                while p + 4 <= end:
                    mtype, msize, mflags = self.buf[p], self._u(p + 1, 2), self.buf[p + 3]
                    p += 4 + (2 if track else 0)
                    if p + msize > end + 4:
                        break
                    if mtype == 0x10:
                        o, n = self._u(p, self.so) + self.base, self._u(p + self.so, self.sl)
                        blocks.append((o + 4, n - 8))        # 'OCHK' signature in front, checksum behind
                    elif mtype != 0:
                        out.append((mtype, mflags, p, msize))
                    p += msize
            return out
        ver = self.buf[addr]
        if ver != 1:
            raise H5FormatError(f"unsupported object header version {ver} at {addr}")
        n_msgs, size = self._u(addr + 2, 2), self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        while blocks and len(out) < n_msgs + 64:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end:
                mtype, msize, mflags = self._u(p, 2), self._u(p + 2, 2), self.buf[p + 4]
                p += 8
                if mtype == 0x10:
                    blocks.append((self._u(p, self.so) + self.base, self._u(p + self.so, self.sl)))
                elif mtype != 0:
                    out.append((mtype, mflags, p, msize))
                p += msize
        return out

    # ---- groups ---------------------------------------------------------------------------------------------------
    def _heap_data(self, addr):
        if self.buf[addr:addr + 4] != b"HEAP":
            raise H5FormatError("bad local heap signature")
        return self._u(addr + 8 + 2 * self.sl, self.so) + self.base

    def _btree_group_entries(self, btree, heap_data, out):
        if self.buf[btree:btree + 4] != b"TREE":
            raise H5FormatError("bad B-tree signature")
        level, used = self.buf[btree + 5], self._u(btree + 6, 2)
        p = btree + 8 + 2 * self.so
        for i in range(used):
            child = self._u(p + self.sl + i * (self.sl + self.so), self.so) + self.base
            if level > 0:
                self._btree_group_entries(child, heap_data, out)
                continue
            if self.buf[child:child + 4] != b"SNOD":
                raise H5FormatError("bad symbol-table node signature")
            n = self._u(child + 6, 2)
            q = child + 8
            for _ in range(n):
                name_off, header = self._u(q, self.so), self._u(q + self.so, self.so) + self.base
                s = heap_data + name_off
                name = self.buf[s:self.buf.index(b"\0", s)].decode("utf-8")
                out.append((name, header))
                q += 2 * self.so + 4 + 4 + 16

    def _links(self, addr):
        """[(name, object header address)] of the group at `addr`, or None if the object is not a group."""
        links, is_group = [], False
        for mtype, _, p, size in self._messages(addr):
            if mtype == 0x11:                              # symbol table: v1 B-tree + local heap
                is_group = True
                self._btree_group_entries(self._u(p, self.so) + self.base, self._heap_data(self._u(p + self.so, self.so) + self.base), links)
            elif mtype == 0x06:                            # link message (compact new-style group)
                is_group = True
                flags = self.buf[p + 1]
                q = p + 2
                ltype = 0
                if flags & 0x08:
                    ltype = self.buf[q]
                    q += 1
                if flags & 0x04:
                    q += 8
                if flags & 0x10:
                    q += 1
                nsz = 1 << (flags & 3)
                nlen = self._u(q, nsz)
                q += nsz
                name = self.buf[q:q + nlen].decode("utf-8")
                q += nlen
                if ltype == 0:
                    links.append((name, self._u(q, self.so) + self.base))
            elif mtype == 0x02:                            # link info: dense storage lives in a fractal heap
                is_group = True
                flags = self.buf[p + 1]
                q = p + 2 + (8 if flags & 1 else 0)
                if self._u(q, self.so) != (1 << (8 * self.so)) - 1:
                    raise H5FormatError("dense link storage (fractal heap) is not supported; save with the library's default (earliest) format")
            elif mtype == 0x0A:
                is_group = True
        return links if is_group else None

    # ---- datasets -------------------------------------------------------------------------------------------------
    def _dtype(self, p):
        cls, ver = self.buf[p] & 0x0F, self.buf[p] >> 4
        bits0 = self.buf[p + 1]
        size = self._u(p + 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 1 and size in (2, 4, 8):
            return np.dtype(f"{order}f{size}")
        if cls == 0 and size in (1, 2, 4, 8):
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        raise H5FormatError(f"unsupported datatype class {cls} (version {ver}, size {size}): only fixed- and floating-point numbers")

    def _read_dataset(self, addr):
        shape = dtype = None
        layout = None
        for mtype, _, p, size in self._messages(addr):
            if mtype == 0x01:
                ver, rank = self.buf[p], self.buf[p + 1]
                q = p + (8 if ver == 1 else 4)
                shape = tuple(self._u(q + i * self.sl, self.sl) for i in range(rank))
            elif mtype == 0x03:
                dtype = self._dtype(p)
            elif mtype == 0x08:
                layout = p
            elif mtype == 0x0B:
                raise H5FormatError("filtered (compressed) datasets are not supported")
        if shape is None or dtype is None or layout is None:
            return None
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        p = layout
        ver = self.buf[p]
        if ver in (1, 2):                                  # old libraries: dimensionality, class, address, 4-byte dims
            rank1, cls = self.buf[p + 1], self.buf[p + 2]
            q = p + 8
            if cls == 0:
                q += 4 * rank1
                size = self._u(q, 4)
                raw = self.buf[q + 4:q + 4 + size]
                return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).astype(dtype.newbyteorder("="))
            a = self._u(q, self.so)
            if cls == 1:
                if a == (1 << (8 * self.so)) - 1:
                    return np.zeros(shape, dtype.newbyteorder("="))
                a += self.base
                return np.frombuffer(self.buf[a:a + n * dtype.itemsize], dtype=dtype, count=n).reshape(shape).astype(dtype.newbyteorder("="))
            if cls == 2:
                cdims = tuple(self._u(q + self.so + 4 * i, 4) for i in range(rank1 - 1))
                out = np.zeros(shape, dtype)
                self._read_chunks(a + self.base, rank1, cdims, out)
                return out.astype(dtype.newbyteorder("="))
            raise H5FormatError(f"unsupported data layout class {cls}")
        if ver != 3:
            raise H5FormatError(f"unsupported data layout message version {ver}")
        cls = self.buf[p + 1]
        if cls == 0:                                       # compact: the data sits in the header
            size = self._u(p + 2, 2)
            raw = self.buf[p + 4:p + 4 + size]
        elif cls == 1:                                     # contiguous
            a = self._u(p + 2, self.so)
            if a == (1 << (8 * self.so)) - 1:
                return np.zeros(shape, dtype.newbyteorder("="))    # never written: fill value 0
            a += self.base
            raw = self.buf[a:a + n * dtype.itemsize]
        elif cls == 2:                                     # chunked, unfiltered: v1 B-tree of raw chunks
            rank1 = self.buf[p + 2]
            btree = self._u(p + 3, self.so) + self.base
            cdims = tuple(self._u(p + 3 + self.so + 4 * i, 4) for i in range(rank1 - 1))
            out = np.zeros(shape, dtype)
            self._read_chunks(btree, rank1, cdims, out)
            return out.astype(dtype.newbyteorder("="))
        else:
            raise H5FormatError(f"unsupported data layout class {cls}")
        if len(raw) < n * dtype.itemsize:
            raise H5FormatError("dataset runs past the end of the file")
        return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).astype(dtype.newbyteorder("="))

    def _read_chunks(self, btree, rank1, cdims, out):
        if self.buf[btree:btree + 4] != b"TREE" or self.buf[btree + 4] != 1:
            raise H5FormatError("bad chunk B-tree")
        level, used = self.buf[btree + 5], self._u(btree + 6, 2)
        key = 8 + 8 * rank1
        p = btree + 8 + 2 * self.so
        for i in range(used):
            k = p + i * (key + self.so)
            csize, mask = self._u(k, 4), self._u(k + 4, 4)
            offs = tuple(self._u(k + 8 + 8 * d, 8) for d in range(rank1 - 1))
            child = self._u(k + key, self.so) + self.base
            if level > 0:
                self._read_chunks(child, rank1, cdims, out)
                continue
            if mask:
                raise H5FormatError("filtered chunks are not supported")
            chunk = np.frombuffer(self.buf[child:child + csize], dtype=out.dtype, count=int(np.prod(cdims))).reshape(cdims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, out.shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

    # ---- public ---------------------------------------------------------------------------------------------------
    def walk(self):
        """Yields (path, object header address, is_group) depth first, children in stored order."""
        seen = set()

        def rec(prefix, addr):
            if addr in seen:
                return
            seen.add(addr)
            links = self._links(addr)
            if links is None:
                yield prefix, addr, False
                return
            if prefix:
                yield prefix, addr, True
            for name, child in links:
                yield from rec(f"{prefix}/{name}" if prefix else name, child)

        yield from rec("", self.root_header)

    def datasets(self):
        out = {}
        for path, addr, is_group in self.walk():
            if not is_group:
                arr = self._read_dataset(addr)
                if arr is not None:
                    out[path] = arr
        return out


def read_h5_datasets(path):
    with H5File(path) as f:
        return f.datasets()
