"""Dependency-free reader / writer for the subset of HDF5 that Keras weight files (``save_weights(..., save_format='h5')``,
reference models/tfmodel.py:150-166) are made of — SURVEY 8f N4. Neither h5py nor libhdf5 is part of this stack.

Reader (``File``): super-block versions 0-3 (with or without a user block), object headers version 1 and 2 (continuation blocks),
old-style groups (symbol-table message -> version-1 B-tree -> symbol nodes -> local heap: what libhdf5 writes with the default
``libver='earliest'``, i.e. what h5py / Keras produce) and new-style groups with compact link messages; datasets with contiguous,
compact or chunked (version-1 B-tree index; deflate / shuffle / fletcher32 filters) layout; fixed-point, floating-point, fixed-length
string and variable-length string (global heap) datatypes; attribute messages version 1-3. Dense link / attribute storage (fractal
heaps), compound / array datatypes, external files and version-4 chunk indexes raise ``H5Error`` instead of guessing.

Writer (``write``): super-block version 0, version-1 object headers, symbol-table groups (leaf K = 4, internal K = 16: libhdf5's
defaults), contiguous little-endian datasets, version-1 attribute messages with fixed-length (null-padded) string or numeric values —
byte-for-byte the structures the reader walks in a libhdf5-written file.

Pinning: the reader is checked against a file written by the real HDF5 library (tests/golden/hdf5_matlab73_testdouble.mat, a MATLAB
7.3 file = HDF5 with a 512-byte user block, from SciPy's test data); the writer is checked against the reader. No h5py exists in this
image, so "libhdf5 opens what the writer produced" is NOT verified here — stated in DESIGN.md.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


# =========================================================================================================================== reader
class _Node(object):
    def __init__(self, f, name, addr):
        self._f, self.name, self._addr = f, name, addr
        self._msgs = f._object_header(addr)
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = {}
            for t, _, d in self._msgs:
                if t == 0x0015 and self._f._dense_attrs(d):
                    raise H5Error('{}: dense attribute storage (fractal heap) is not supported'.format(self.name))
                if t == 0x000C:
                    k, v = self._f._attribute(d)
                    self._attrs[k] = v
        return self._attrs


class Group(_Node):
    def __init__(self, f, name, addr):
        super().__init__(f, name, addr)
        self._links = None

    def _load(self):
        if self._links is not None:
            return self._links
        f, links = self._f, {}
        for t, _, d in self._msgs:
            if t == 0x0011:                                   # symbol table: B-tree + local heap
                btree, heap = f._unpack_addr(d, 0), f._unpack_addr(d, f.O)
                hdata = f._local_heap(heap)
                for off, addr in f._group_btree(btree):
                    end = hdata.index(b'\0', off)
                    links[hdata[off:end].decode('utf8')] = addr
            elif t == 0x0006:                                 # compact link message
                name, addr = f._link(d)
                if addr is not None:
                    links[name] = addr
            elif t == 0x0002:                                 # link info: a fractal heap address means dense storage
                flags = d[1]
                pos = 2 + (8 if flags & 1 else 0)
                if f._unpack_addr(d, pos) != UNDEF:
                    raise H5Error('{}: dense link storage (fractal heap) is not supported'.format(self.name))
        self._links = links
        return links

    def keys(self):
        return list(self._load().keys())

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split('/') if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            links = node._load()
            if part not in links:
                raise KeyError('{} (no member {!r} in {})'.format(path, part, node.name))
            node = node._f._open((node.name.rstrip('/') + '/' + part), links[part])
        return node

    def visit(self, fn, _prefix=''):
        for k in sorted(self._load()):
            child = self[k]
            fn(_prefix + k, child)
            if isinstance(child, Group):
                child.visit(fn, _prefix + k + '/')


class Dataset(_Node):
    def __init__(self, f, name, addr):
        super().__init__(f, name, addr)
        self.shape = self.dtype = None
        self._layout = self._filters = None
        self._vlen = False
        for t, _, d in self._msgs:
            if t == 0x0001:
                self.shape = f._dataspace(d)
            elif t == 0x0003:
                self.dtype, self._vlen = f._datatype(d)
            elif t == 0x0008:
                self._layout = d
            elif t == 0x000B:
                self._filters = f._filter_pipeline(d)
        if self.shape is None or self.dtype is None or self._layout is None:
            raise H5Error('{}: not a dataset (dataspace / datatype / layout message missing)'.format(name))

    def read(self):
        f, d = self._f, self._layout
        n = int(np.prod(self.shape)) if self.shape else 1
        esize = 16 if self._vlen else self.dtype.itemsize
        version = d[0]
        if version in (1, 2):
            rank, cls = d[1], d[2]
            pos = 8
            addr = None
            if cls != 0:
                addr = f._unpack_addr(d, pos)
                pos += f.O
            dims = struct.unpack_from('<{}I'.format(rank), d, pos)
            pos += 4 * rank
            if cls == 2:
                raw = self._chunked(addr, dims, esize)
            elif cls == 1:
                raw = b'\0' * (n * esize) if addr == UNDEF else f._bytes(addr, n * esize)
            else:
                size, = struct.unpack_from('<I', d, pos)
                raw = bytes(d[pos + 4:pos + 4 + size])
        elif version in (3, 4):
            cls = d[1]
            if cls == 0:
                size, = struct.unpack_from('<H', d, 2)
                raw = bytes(d[4:4 + size])
            elif cls == 1:
                addr, size = f._unpack_addr(d, 2), f._unpack_len(d, 2 + f.O)
                raw = b'\0' * (n * esize) if addr == UNDEF else f._bytes(addr, n * esize)
            elif cls == 2 and version == 3:
                rank = d[2]
                addr = f._unpack_addr(d, 3)
                dims = struct.unpack_from('<{}I'.format(rank), d, 3 + f.O)
                raw = self._chunked(addr, dims[:-1], esize)
            else:
                raise H5Error('{}: layout class {} of a version-{} layout message is not supported'.format(self.name, cls, version))
        else:
            raise H5Error('{}: layout message version {}'.format(self.name, version))
        if self._vlen:
            return f._vlen_strings(raw, self.shape)
        return np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape).copy()

    def __getitem__(self, key):
        return self.read()[key]

    def _chunked(self, btree, chunk, esize):
        f, rank = self._f, len(self.shape)
        chunk = tuple(int(c) for c in chunk[:rank])
        out = np.zeros(self.shape, dtype=np.dtype((np.void, esize)))
        if btree == UNDEF:
            return out.tobytes()
        for size, mask, offs, addr in f._chunk_btree(btree, rank):
            raw = f._bytes(addr, size)
            for i, (fid, cd) in reversed(list(enumerate(self._filters or []))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    a = np.frombuffer(raw, np.uint8)
                    k = len(a) // esize
                    raw = a[:k * esize].reshape(esize, k).T.tobytes() + a[k * esize:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise H5Error('{}: filter {} is not supported'.format(self.name, fid))
            block = np.frombuffer(raw, dtype=out.dtype, count=int(np.prod(chunk))).reshape(chunk)
            sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, self.shape))
            sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = block[sl_in]
        return out.tobytes()


class File(Group):
    """Read-only view of an HDF5 file: ``f['group/dataset'].read()``, ``f.attrs``, ``f.keys()``, ``f.visit(fn)``."""

    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self.buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, 'rb') as fh:
                self.buf = fh.read()
        self._cache = {}
        self._superblock()
        super().__init__(self, '/', self._root)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    # ---- low level
    def _bytes(self, addr, n):
        a = addr + self.base
        if addr == UNDEF or a < 0 or a + n > len(self.buf):
            raise H5Error('address {:#x}+{} outside the file ({} bytes)'.format(addr, n, len(self.buf)))
        return self.buf[a:a + n]

    def _uint(self, data, pos, size):
        return int.from_bytes(data[pos:pos + size], 'little')

    def _unpack_addr(self, data, pos):
        v = self._uint(data, pos, self.O)
        return UNDEF if v == (1 << (8 * self.O)) - 1 else v

    def _unpack_len(self, data, pos):
        return self._uint(data, pos, self.L)

    def _superblock(self):
        b, off = self.buf, 0
        while True:                                           # the signature sits at 0, 512, 1024, 2048 ... (user block)
            if b[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(b):
                raise H5Error('not an HDF5 file (signature not found)')
        version = b[off + 8]
        if version in (0, 1):
            self.O, self.L = b[off + 13], b[off + 14]
            pos = off + 24 + (4 if version == 1 else 0)
            self.base = 0
            base = self._unpack_addr(b, pos)
            self.base = base if base != UNDEF else 0
            pos += 4 * self.O                                 # base, free-space info, end of file, driver info
            self._root = self._unpack_addr(b, pos + self.O)   # root symbol-table entry: name offset, object header address
        elif version in (2, 3):
            self.O, self.L = b[off + 9], b[off + 10]
            self.base = 0
            base = self._unpack_addr(b, off + 12)
            self.base = base if base != UNDEF else 0
            self._root = self._unpack_addr(b, off + 12 + 3 * self.O)
        else:
            raise H5Error('super-block version {}'.format(version))
        if self.base == 0 and off:
            self.base = off                                   # some writers leave the base address 0 behind a user block

    def _open(self, name, addr):
        if addr not in self._cache:
            msgs = self._object_header(addr)
            kinds = {t for t, _, _ in msgs}
            self._cache[addr] = Dataset(self, name, addr) if 0x0008 in kinds else Group(self, name, addr)
        return self._cache[addr]

    def _object_header(self, addr):
        head = self._bytes(addr, 16)
        msgs = []
        if head[:4] == b'OHDR':
            flags = head[5]
            pos = 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
            csz = 1 << (flags & 3)
            head = self._bytes(addr, pos + csz)
            size0 = self._uint(head, pos, csz)
            blocks = [(addr + pos + csz, size0)]
            while blocks:
                a, n = blocks.pop(0)
                data = self._bytes(a, n)
                p = 0
                hs = 4 + (2 if flags & 4 else 0)
                while p + hs <= n:
                    t, size, mflags = data[p], self._uint(data, p + 1, 2), data[p + 3]
                    body = data[p + hs:p + hs + size]
                    p += hs + size
                    if t == 0x10:
                        ca, cl = self._unpack_addr(body, 0), self._unpack_len(body, self.O)
                        if self._bytes(ca, 4) != b'OCHK':
                            raise H5Error('object header continuation without OCHK signature')
                        blocks.append((ca + 4, cl - 8))
                    elif t != 0:
                        msgs.append((t, mflags, body))
            return msgs
        if head[0] != 1:
            raise H5Error('object header version {} at {:#x}'.format(head[0], addr))
        nmsgs, = struct.unpack_from('<H', head, 2)
        hsize, = struct.unpack_from('<I', head, 8)
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsgs + 64:
            a, n = blocks.pop(0)
            data = self._bytes(a, n)
            p = 0
            while p + 8 <= n:
                t, size, mflags = struct.unpack_from('<HHB', data, p)
                body = data[p + 8:p + 8 + size]
                p += 8 + size
                if t == 0x10:
                    blocks.append((self._unpack_addr(body, 0), self._unpack_len(body, self.O)))
                elif t != 0:
                    msgs.append((t, mflags, body))
        return msgs

    def _local_heap(self, addr):
        h = self._bytes(addr, 8 + 2 * self.L + self.O)
        if h[:4] != b'HEAP':
            raise H5Error('local heap signature missing at {:#x}'.format(addr))
        size = self._unpack_len(h, 8)
        return self._bytes(self._unpack_addr(h, 8 + 2 * self.L), size)

    def _group_btree(self, addr):
        """(heap offset of the name, object header address) of every link below a version-1 group B-tree node."""
        h = self._bytes(addr, 8 + 2 * self.O)
        if h[:4] == b'SNOD':
            n, = struct.unpack_from('<H', h, 6)
            es = 2 * self.O + 24
            data = self._bytes(addr + 8, n * es)
            return [(self._uint(data, i * es, self.O), self._unpack_addr(data, i * es + self.O)) for i in range(n)]
        if h[:4] != b'TREE' or h[4] != 0:
            raise H5Error('group B-tree node expected at {:#x}'.format(addr))
        used, = struct.unpack_from('<H', h, 6)
        data = self._bytes(addr + 8 + 2 * self.O, used * (self.L + self.O) + self.L)
        out = []
        for i in range(used):
            out.extend(self._group_btree(self._unpack_addr(data, i * (self.L + self.O) + self.L)))
        return out

    def _chunk_btree(self, addr, rank):
        h = self._bytes(addr, 8 + 2 * self.O)
        if h[:4] != b'TREE' or h[4] != 1:
            raise H5Error('chunk B-tree node expected at {:#x}'.format(addr))
        level, (used,) = h[5], struct.unpack_from('<H', h, 6)
        ks = 8 + 8 * (rank + 1)
        data = self._bytes(addr + 8 + 2 * self.O, used * (ks + self.O) + ks)
        out = []
        for i in range(used):
            p = i * (ks + self.O)
            size, mask = struct.unpack_from('<II', data, p)
            offs = struct.unpack_from('<{}Q'.format(rank), data, p + 8)
            child = self._unpack_addr(data, p + ks)
            if level:
                out.extend(self._chunk_btree(child, rank))
            else:
                out.append((size, mask, offs, child))
        return out

    def _link(self, d):
        flags = d[1]
        pos, ltype = 2, 0
        if flags & 8:
            ltype = d[pos]
            pos += 1
        if flags & 4:
            pos += 8
        if flags & 0x10:
            pos += 1
        ls = 1 << (flags & 3)
        n = self._uint(d, pos, ls)
        pos += ls
        name = bytes(d[pos:pos + n]).decode('utf8')
        pos += n
        return name, (self._unpack_addr(d, pos) if ltype == 0 else None)     # soft / external links are skipped

    def _dense_attrs(self, d):
        flags = d[1]
        pos = 2 + (2 if flags & 1 else 0)
        return self._unpack_addr(d, pos) != UNDEF

    # ---- message decoders
    def _dataspace(self, d):
        version, rank, flags = d[0], d[1], d[2]
        if version == 1:
            pos = 8
        elif version == 2:
            if d[3] == 2:                                      # null dataspace
                return (0,)
            pos = 4
        else:
            raise H5Error('dataspace message version {}'.format(version))
        return tuple(self._unpack_len(d, pos + i * self.L) for i in range(rank))

    def _datatype(self, d):
        """-> (numpy dtype, is variable-length string)."""
        cls, b0, b1 = d[0] & 0x0F, d[1], d[2]
        size, = struct.unpack_from('<I', d, 4)
        order = '>' if b0 & 1 else '<'
        if cls == 0:
            return np.dtype('{}{}{}'.format(order, 'i' if b0 & 8 else 'u', size)), False
        if cls == 1:
            if size not in (2, 4, 8):
                raise H5Error('floating-point type of {} bytes'.format(size))
            return np.dtype('{}f{}'.format(order, size)), False
        if cls == 3:
            return np.dtype('S{}'.format(size)), False
        if cls == 9:
            if (b0 & 0x0F) != 1:
                raise H5Error('variable-length sequences are not supported (only strings)')
            return np.dtype('O'), True
        if cls == 8:                                           # enumeration (h5py booleans): the base integer type follows
            return self._datatype(d[8:])
        if cls == 7:
            return np.dtype('<u{}'.format(size)), False        # object references: raw addresses
        raise H5Error('datatype class {} is not supported'.format(cls))

    def _filter_pipeline(self, d):
        version, n = d[0], d[1]
        pos = 8 if version == 1 else 2
        out = []
        for _ in range(n):
            fid, = struct.unpack_from('<H', d, pos)
            pos += 2
            nlen = 0
            if version == 1 or fid >= 256:
                nlen, = struct.unpack_from('<H', d, pos)
                pos += 2
            _, ncd = struct.unpack_from('<HH', d, pos)
            pos += 4
            if version == 1:
                nlen = (nlen + 7) // 8 * 8
            pos += nlen
            cd = struct.unpack_from('<{}I'.format(ncd), d, pos)
            pos += 4 * ncd
            if version == 1 and ncd % 2:
                pos += 4
            out.append((fid, cd))
        return out

    def _vlen_strings(self, raw, shape):
        n = int(np.prod(shape)) if shape else 1
        es = 4 + self.O + 4
        vals = []
        for i in range(n):
            ln, = struct.unpack_from('<I', raw, i * es)
            addr = self._unpack_addr(raw, i * es + 4)
            idx, = struct.unpack_from('<I', raw, i * es + 4 + self.O)
            vals.append(self._global_heap_object(addr, idx)[:ln].decode('utf8') if ln else '')
        if not shape:
            return vals[0]
        out = np.empty(n, dtype=object)
        out[:] = vals
        return out.reshape(shape)

    def _global_heap_object(self, addr, index):
        h = self._bytes(addr, 8 + self.L)
        if h[:4] != b'GCOL':
            raise H5Error('global heap signature missing at {:#x}'.format(addr))
        total = self._unpack_len(h, 8)
        data = self._bytes(addr, total)
        p = 8 + self.L
        while p + 8 + self.L <= total:
            idx, = struct.unpack_from('<H', data, p)
            size = self._unpack_len(data, p + 8)
            if idx == 0:
                break
            if idx == index:
                return bytes(data[p + 8 + self.L:p + 8 + self.L + size])
            p += 8 + self.L + (size + 7) // 8 * 8
        raise H5Error('global heap object {} not found'.format(index))

    def _attribute(self, d):
        version = d[0]
        nsz, tsz, ssz = struct.unpack_from('<HHH', d, 2)
        pos = 8 if version in (1, 2) else 9
        pad = (lambda n: (n + 7) // 8 * 8) if version == 1 else (lambda n: n)
        name = bytes(d[pos:pos + nsz]).split(b'\0')[0].decode('utf8')
        pos += pad(nsz)
        dt, vlen = self._datatype(d[pos:pos + tsz])
        pos += pad(tsz)
        shape = self._dataspace(d[pos:pos + ssz]) if ssz else ()
        pos += pad(ssz)
        raw = bytes(d[pos:])
        if vlen:
            return name, self._vlen_strings(raw, shape)
        n = int(np.prod(shape)) if shape else 1
        a = np.frombuffer(raw, dtype=dt, count=n).reshape(shape).copy()
        return name, (a[()] if not shape else a)


# =========================================================================================================================== writer
def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


def _msg(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack('<HHB3x', mtype, len(body), flags) + body


def _dataspace_msg(shape):
    return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(struct.pack('<Q', int(s)) for s in shape)


def _datatype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        exp, man = (8, 23) if dt.itemsize == 4 else (11, 52)
        return struct.pack('<BBBBI', 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize) + struct.pack(
            '<HHBBBBI', 0, 8 * dt.itemsize, man, exp, 0, man, (1 << (exp - 1)) - 1)
    if dt.kind in 'iu':
        return struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0, 0, 0, dt.itemsize) + struct.pack('<HH', 0, 8 * dt.itemsize)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x01, 0, 0, max(dt.itemsize, 1))      # null-padded ASCII (h5py's mapping of numpy 'S')
    raise H5Error('cannot store dtype {}'.format(dt))


def _attr_msg(name, value):
    a = np.asarray(value)
    if a.dtype.kind == 'U':
        a = np.char.encode(a, 'utf8')
    if a.dtype.kind == 'S' and a.dtype.itemsize == 0:
        a = a.astype('S1')
    if a.dtype.kind in 'fiu':
        a = a.astype(a.dtype.newbyteorder('<'))
    nm = name.encode('utf8') + b'\0'
    dt, ds = _datatype_msg(a.dtype), _dataspace_msg(a.shape)
    body = struct.pack('<BxHHH', 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + a.tobytes()
    if len(body) > 64000:
        raise H5Error('attribute {} is too large for a compact object-header message'.format(name))
    return _msg(0x000C, body)


class _Writer(object):
    LEAF_K, NODE_K = 4, 16

    def __init__(self):
        self.buf = bytearray(b'\0' * 96)                      # super-block written last

    def alloc(self, data):
        self.buf.extend(b'\0' * (-len(self.buf) % 8))
        addr = len(self.buf)
        self.buf.extend(data)
        return addr

    def header(self, msgs):
        body = b''.join(msgs)
        return self.alloc(struct.pack('<BxHII4x', 1, len(msgs), 1, len(body)) + body)

    def dataset(self, array, attrs):
        a = np.asarray(array, order='C')                     # (ascontiguousarray would turn a 0-d array into shape (1,))
        if a.dtype.kind in 'fiu':
            a = a.astype(a.dtype.newbyteorder('<'))
        raw = a.tobytes()
        addr = self.alloc(raw) if raw else UNDEF
        msgs = [_msg(0x0001, _dataspace_msg(a.shape)), _msg(0x0003, _datatype_msg(a.dtype), 1),
                _msg(0x0005, struct.pack('<BBBB', 2, 2, 2, 0)),                      # fill value: version 2, late allocation, undefined
                _msg(0x0008, struct.pack('<BBQQ', 3, 1, addr, len(raw)))]
        return self.header(msgs + [_attr_msg(k, v) for k, v in attrs.items()])

    def group(self, members, attrs):
        """members: {name: object header address}."""
        names = sorted(members, key=lambda s: s.encode('utf8'))
        if len(names) > 2 * self.LEAF_K * 2 * self.NODE_K:
            raise H5Error('more than {} members in one group'.format(2 * self.LEAF_K * 2 * self.NODE_K))
        heap, offs = bytearray(b'\0' * 8), {}
        for n in names:
            offs[n] = len(heap)
            heap.extend(_pad8(n.encode('utf8') + b'\0'))
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), 1, heap_data))   # free-list head 1 = no free block
        per = 2 * self.LEAF_K
        snods, keys = [], [0]
        for i in range(0, len(names), per):
            part = names[i:i + per]
            ents = b''.join(struct.pack('<QQII16x', offs[n], members[n], 0, 0) for n in part)
            ents += b'\0' * (40 * (per - len(part)))
            snods.append(self.alloc(b'SNOD' + struct.pack('<BxH', 1, len(part)) + ents))
            keys.append(offs[part[-1]])
        node = b'TREE' + struct.pack('<BBHQQ', 0, 0, len(snods), UNDEF, UNDEF)
        body = struct.pack('<Q', keys[0])
        for k, s in zip(keys[1:], snods):
            body += struct.pack('<QQ', s, k)
        body += b'\0' * (16 * (2 * self.NODE_K - len(snods)))
        btree = self.alloc(node + body)
        addr = self.header([_msg(0x0011, struct.pack('<QQ', btree, heap_addr))] + [_attr_msg(k, v) for k, v in attrs.items()])
        return addr, btree, heap_addr

    def finish(self, root, btree, heap):
        sb = SIGNATURE + struct.pack('<BBBBBBBxHHI', 0, 0, 0, 0, 0, 8, 8, self.LEAF_K, self.NODE_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQII', 0, root, 1, 0) + struct.pack('<QQ', btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write(path, tree, attrs=None):
    """Write an HDF5 file. ``tree``: nested dict; a dict value is a group, anything else a dataset (``numpy`` array). Attributes:
    ``attrs`` = {object path ('' or '/' for the root): {name: value}}; values are numeric / byte-string arrays or scalars."""
    attrs = {k.strip('/'): v for k, v in (attrs or {}).items()}
    w = _Writer()

    def emit(node, path):
        if isinstance(node, dict):
            members = {}
            for name, child in node.items():
                if '/' in name or not name:
                    raise H5Error('invalid member name {!r}'.format(name))
                members[name] = emit(child, (path + '/' + name).strip('/'))
            res = w.group(members, attrs.get(path, {}))
            return res if path == '' else res[0]
        return w.dataset(node, attrs.get(path, {}))

    root, btree, heap = emit(tree, '')
    data = w.finish(root, btree, heap)
    if path is not None:
        with open(path, 'wb') as fh:
            fh.write(data)
    return data
