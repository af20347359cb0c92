"""Pure-Python reader for MATLAB v7.3 ``.mat`` files (= HDF5 with a 512-byte user block), enough for the reference's
``RecordedData.mat`` (``fwi_script.py:18`` loads it with ``mat73``, which is not installable here):
superblock v0, v1 object headers, v1 group B-trees + local heaps, data layout v3 (compact / contiguous / chunked with
deflate + shuffle), fixed-point, floating-point and compound ``{real, imag}`` datatypes.  Host-side I/O only: NumPy +
the standard library (``struct``, ``zlib``); no device work.

    d = load_mat73("RecordedData.mat")     # dict of NumPy arrays, MATLAB orientation like mat73.loadmat
    d["REC_DATA"][tx, rx], d["x_circ"], d["y_circ"], d["f"], d["C"], d["x"], d["y"]
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_SIGNATURE = b"\x89HDF\r\n\x1a\n"


class Hdf5Mini:
    def __init__(self, path):
        self.d = open(path, "rb").read()
        self.sb = self.d.find(_SIGNATURE)
        if self.sb < 0:
            raise ValueError(f"{path}: no HDF5 signature (MATLAB v7.3 expected)")
        d, o = self.d, self.sb + 8
        ver = d[o]
        if ver != 0:
            raise NotImplementedError(f"HDF5 superblock version {ver}")
        o += 8
        o += 8  # group leaf / internal node K, consistency flags
        base, _fs, _eof, _drv = struct.unpack_from("<4Q", d, o)
        o += 32
        # a non-zero user block shifts every address by the superblock offset
        self.base = self.sb if (base == 0 and self.sb != 0) else base
        _lno, oh, cache_type, _ = struct.unpack_from("<QQII", d, o)
        o += 24
        self.root_btree, self.root_heap = struct.unpack_from("<QQ", d, o) if cache_type == 1 else (None, None)
        self.root_oh = oh

    def _abs(self, a):
        return a + self.base

    def _messages(self, addr):
        d, a = self.d, self._abs(addr)
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", d, a)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        blocks, out = [(a + 16, hsize)], []
        while blocks:
            p, ln = blocks.pop(0)
            end = p + ln
            while p + 8 <= end and len(out) < nmsg:
                t, sz, fl = struct.unpack_from("<HHB", d, p)
                p += 8
                body = d[p:p + sz]
                p += sz
                if t == 0x10:  # continuation
                    off, l = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self._abs(off), l))
                out.append((t, fl, body))
        return out

    def _heap_str(self, heap_addr, off):
        d, a = self.d, self._abs(heap_addr)
        assert d[a:a + 4] == b"HEAP"
        seg = struct.unpack_from("<Q", d, a + 24)[0]
        p = self._abs(seg) + off
        return d[p:d.index(b"\0", p)].decode()

    def entries(self):
        """{name: object-header address} of the root group."""
        d, res = self.d, {}

        def walk(addr):
            a = self._abs(addr)
            if d[a:a + 4] == b"TREE":
                n = struct.unpack_from("<H", d, a + 6)[0]
                p = a + 8 + 16
                for _ in range(n):
                    p += 8
                    walk(struct.unpack_from("<Q", d, p)[0])
                    p += 8
            elif d[a:a + 4] == b"SNOD":
                n = struct.unpack_from("<H", d, a + 6)[0]
                p = a + 8
                for _ in range(n):
                    lno, oh, _ct = struct.unpack_from("<QQI", d, p)
                    res[self._heap_str(self.root_heap, lno)] = oh
                    p += 40
            else:
                raise ValueError("unexpected group node")

        walk(self.root_btree)
        return res

    def _dtype(self, b, o=0):
        cv, b0, b1, _b2, size = struct.unpack_from("<BBBBI", b, o)
        cls, ver = cv & 15, cv >> 4
        o += 8
        end = "<" if not b0 & 1 else ">"
        if cls == 0:
            return np.dtype(end + ("i" if (b0 >> 3) & 1 else "u") + str(size)), o + 4
        if cls == 1:
            return np.dtype(end + "f" + str(size)), o + 12
        if cls == 6:  # compound (MATLAB complex = {real, imag})
            n = b0 | (b1 << 8)
            names, fmts, offs = [], [], []
            for _ in range(n):
                e = b.index(b"\0", o)
                names.append(b[o:e].decode())
                o = o + ((e - o + 1) + 7) // 8 * 8 if ver < 3 else e + 1
                if ver == 1:
                    off = struct.unpack_from("<I", b, o)[0]
                    o += 4 + 1 + 3 + 4 + 4 + 16
                elif ver == 2:
                    off = struct.unpack_from("<I", b, o)[0]
                    o += 4
                else:
                    nb = 1 if size < 256 else 2 if size < 65536 else 4
                    off = int.from_bytes(b[o:o + nb], "little")
                    o += nb
                dt, o = self._dtype(b, o)
                fmts.append(dt)
                offs.append(off)
            return np.dtype({"names": names, "formats": fmts, "offsets": offs, "itemsize": size}), o
        raise NotImplementedError(f"HDF5 datatype class {cls}")

    def read(self, oh):
        shape = dtype = layout = None
        filters = []
        for t, _fl, b in self._messages(oh):
            if t == 1:  # dataspace
                ver, rank = b[0], b[1]
                shape = struct.unpack_from("<%dQ" % rank, b, 8 if ver == 1 else 4)
            elif t == 3:
                dtype, _ = self._dtype(b)
            elif t == 8:  # layout v3
                if b[0] != 3:
                    raise NotImplementedError("data layout version")
                if b[1] == 1:
                    layout = ("contig",) + struct.unpack_from("<QQ", b, 2)
                elif b[1] == 2:
                    nd = b[2]
                    layout = ("chunked", struct.unpack_from("<Q", b, 3)[0], struct.unpack_from("<%dI" % nd, b, 11))
                else:
                    sz = struct.unpack_from("<H", b, 2)[0]
                    layout = ("compact", b[4:4 + sz])
            elif t == 0xB:  # filter pipeline
                ver, nf = b[0], b[1]
                o = 8 if ver == 1 else 2
                for _ in range(nf):
                    if ver == 1:
                        fid, nl, _ffl, ncv = struct.unpack_from("<HHHH", b, o)
                        o += 8 + (nl + 7) // 8 * 8
                        cv = struct.unpack_from("<%dI" % ncv, b, o)
                        o += 4 * ncv + (4 if ncv % 2 else 0)
                    else:
                        fid = struct.unpack_from("<H", b, o)[0]
                        o += 2
                        nl = 0
                        if fid >= 256:
                            nl = struct.unpack_from("<H", b, o)[0]
                            o += 2
                        _ffl, ncv = struct.unpack_from("<HH", b, o)
                        o += 4 + nl
                        cv = struct.unpack_from("<%dI" % ncv, b, o)
                        o += 4 * ncv
                    filters.append((fid, cv))
        if layout[0] == "contig":
            return np.frombuffer(self.d, dtype=dtype, count=int(np.prod(shape)), offset=self._abs(layout[1])).reshape(shape)
        if layout[0] == "compact":
            return np.frombuffer(layout[1], dtype=dtype).reshape(shape)
        bt, cd = layout[1], layout[2]
        cdims, nd = cd[:-1], len(shape)
        out = np.zeros(shape, dtype=dtype)
        d = self.d

        def walk(addr):
            a = self._abs(addr)
            assert d[a:a + 4] == b"TREE"
            _nt, lvl, n = struct.unpack_from("<BBH", d, a + 4)
            p = a + 24
            for _ in range(n):
                csz, fmask = struct.unpack_from("<II", d, p)
                offs = struct.unpack_from("<%dQ" % (nd + 1), d, p + 8)
                p += 8 + 8 * (nd + 1)
                child = struct.unpack_from("<Q", d, p)[0]
                p += 8
                if lvl > 0:
                    walk(child)
                    continue
                raw = d[self._abs(child):self._abs(child) + csz]
                for fi, (fid, cv) in reversed(list(enumerate(filters))):
                    if fmask >> fi & 1:
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:  # shuffle
                        es = cv[0]
                        arr = np.frombuffer(raw, dtype=np.uint8)
                        ne = len(arr) // es
                        raw = arr[:ne * es].reshape(es, ne).T.tobytes()
                    else:
                        raise NotImplementedError(f"HDF5 filter {fid}")
                ch = np.frombuffer(raw, dtype=dtype).reshape(cdims)
                sl = tuple(slice(o_, min(o_ + c, sh)) for o_, c, sh in zip(offs, cdims, shape))
                out[sl] = ch[tuple(slice(0, x.stop - x.start) for x in sl)]

        walk(bt)
        return out


def load_mat73(path):
    """All numeric variables of a MATLAB v7.3 file as NumPy arrays in MATLAB orientation (HDF5 stores the dimensions
    reversed, so every array is transposed; complex variables arrive as ``{real, imag}`` compounds)."""
    h = Hdf5Mini(path)
    out = {}
    for name, oh in h.entries().items():
        if name.startswith("#"):
            continue
        a = h.read(oh)
        if a.dtype.names and set(a.dtype.names) == {"real", "imag"}:
            a = a["real"] + 1j * a["imag"]
        out[name] = np.ascontiguousarray(a.T)
    return out
