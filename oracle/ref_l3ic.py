"""CPU restatement of the reference's l3ic byte-stream codec (TEST INFRASTRUCTURE — see oracle/__init__.py).

Two layers:
* the FSE entropy coder the reference vendors (pyfse/FiniteStateEntropy/lib, wrapped by pyfse/pyfse.pyx:24-72), restated here in plain
  Python (`fse_compress`, `fse_decompress`; small inputs only — it is a loop per symbol) and PINNED: tests/test_l3ic.py checks it byte for
  byte against the reference library itself (oracle/_ref/libfse_ref.so, built by oracle/Makefile from the sources under /root/reference)
  and against tests/golden/fse_vectors.npz, which that library produced (tests/golden/make_fse_golden.py);
* the container format of compression/codec.py:87-265 (`l3ic_compress`, `l3ic_decompress`) on top of either coder.
"""
import ctypes
import os

import numpy as np

_REF_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'libfse_ref.so')
M32, M64 = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF
MIN_LOG, MAX_LOG, DEFAULT_LOG, ABS_MAX_LOG = 5, 12, 11, 15


class FSEError(Exception):
    pass


# ------------------------------------------------------------------------------------------------ the reference library itself
def reference_library():
    """ctypes handle of the compiled reference coder, or None when oracle/_ref has not been built."""
    if not os.path.isfile(_REF_SO):
        return None
    lib = ctypes.CDLL(_REF_SO)
    for name in ('FSE_compress', 'FSE_decompress'):
        fn = getattr(lib, name)
        fn.restype = ctypes.c_size_t
        fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    lib.FSE_isError.argtypes = [ctypes.c_size_t]
    lib.FSE_compressBound.restype = ctypes.c_size_t
    lib.FSE_compressBound.argtypes = [ctypes.c_size_t]
    return lib


def ref_compress(lib, data):
    """pyfse.compress through the reference library: bytes, or 0 / 1 (not compressible / repeated symbol), or -1 (error)."""
    cap = lib.FSE_compressBound(len(data))
    dst = ctypes.create_string_buffer(cap)
    r = lib.FSE_compress(dst, cap, data, len(data))
    if lib.FSE_isError(r):
        return -1
    return dst.raw[:r] if r > 1 else int(r)


def ref_decompress(lib, data, cap):
    dst = ctypes.create_string_buffer(cap + 16)
    r = lib.FSE_decompress(dst, cap, data, len(data))
    return -1 if lib.FSE_isError(r) else dst.raw[:r]


# ------------------------------------------------------------------------------------------------ plain-Python restatement
def _highbit(v):
    return int(v).bit_length() - 1


def _min_table_log(n, max_symbol):                       # fse_compress.c:316-324
    return min(_highbit(n - 1) + 1, _highbit(max_symbol) + 2)


def _optimal_table_log(n, max_symbol):                   # fse_compress.c:327-345 (maxTableLog 11, minus 2, unsigned wrap)
    src_bits = (_highbit(n - 1) - 2) & M32
    t = DEFAULT_LOG
    if src_bits < t:
        t = src_bits
    t = max(t, _min_table_log(n, max_symbol))
    return min(max(t, MIN_LOG), MAX_LOG)


def _normalize_m2(norm, table_log, count, total, max_symbol):      # fse_compress.c:353-447
    given = 0
    low = total >> table_log
    low_one = (total * 3) >> (table_log + 1)
    for s in range(max_symbol + 1):
        if count[s] == 0:
            norm[s] = 0
        elif count[s] <= low:
            norm[s] = -1; given += 1; total -= count[s]
        elif count[s] <= low_one:
            norm[s] = 1; given += 1; total -= count[s]
        else:
            norm[s] = -2
    left = (1 << table_log) - given
    if total // left > low_one:
        low_one = (total * 3) // (left * 2)
        for s in range(max_symbol + 1):
            if norm[s] == -2 and count[s] <= low_one:
                norm[s] = 1; given += 1; total -= count[s]
        left = (1 << table_log) - given
    if given == max_symbol + 1:
        best = int(np.argmax(count[:max_symbol + 1]))
        norm[best] += left
        return
    if total == 0:
        s = 0
        while left > 0:
            if norm[s] > 0:
                left -= 1; norm[s] += 1
            s = (s + 1) % (max_symbol + 1)
        return
    shift = 62 - table_log
    mid = (1 << (shift - 1)) - 1
    r_step = (((1 << shift) * left) + mid) // total
    run = mid
    for s in range(max_symbol + 1):
        if norm[s] == -2:
            end = (run + count[s] * r_step) & M64
            weight = ((end >> shift) & M32) - ((run >> shift) & M32)
            if weight < 1:
                raise FSEError('normalisation failed')
            norm[s] = weight
            run = end


def _normalize(count, total, max_symbol, table_log):     # fse_compress.c:450-508
    beat = (0, 473195, 504333, 520860, 550000, 700000, 750000, 830000)
    norm = [0] * 256
    scale = 62 - table_log
    step = (1 << 62) // total
    v_step = 1 << (scale - 20)
    left = 1 << table_log
    largest, largest_p = 0, 0
    low = total >> table_log
    for s in range(max_symbol + 1):
        c = int(count[s])
        if c == 0:
            continue
        if c <= low:
            norm[s] = -1; left -= 1
            continue
        scaled = c * step
        p = scaled >> scale
        if p < 8 and scaled - (p << scale) > v_step * beat[p]:
            p += 1
        if p > largest_p:
            largest_p, largest = p, s
        norm[s] = p
        left -= p
    if -left >= (norm[largest] >> 1):
        _normalize_m2(norm, table_log, [int(c) for c in count], total, max_symbol)
    else:
        norm[largest] += left
    return norm


def _write_ncount(norm, max_symbol, table_log):          # fse_compress.c:204-285
    out = bytearray()
    size = 1 << table_log
    remaining, threshold, nb = size + 1, size, table_log + 1
    bits, have, sym, prev_zero = table_log - MIN_LOG, 4, 0, False

    def put16():
        nonlocal bits
        out.extend((bits & 0xFF, (bits >> 8) & 0xFF))
        bits >>= 16
    while remaining > 1:
        if prev_zero:
            start = sym
            while not norm[sym]:
                sym += 1
            while sym >= start + 24:
                start += 24
                bits = (bits + (0xFFFF << have)) & M32
                put16()
            while sym >= start + 3:
                start += 3
                bits = (bits + (3 << have)) & M32
                have += 2
            bits = (bits + ((sym - start) << have)) & M32
            have += 2
            if have > 16:
                put16(); have -= 16
        c = norm[sym]
        sym += 1
        mx = (2 * threshold - 1) - remaining
        remaining -= abs(c)
        c += 1
        if c >= threshold:
            c += mx
        bits = (bits + (c << have)) & M32
        have += nb
        have -= 1 if c < mx else 0
        prev_zero = c == 1
        if remaining < 1:
            raise FSEError('header')
        while remaining < threshold:
            nb -= 1; threshold >>= 1
        if have > 16:
            put16(); have -= 16
    tail = bytes((bits & 0xFF, (bits >> 8) & 0xFF))
    out.extend(tail[:(have + 7) // 8])
    return bytes(out)


def _spread(norm, max_symbol, table_log):
    """Symbol of every table cell (shared by the encoder, fse_compress.c:110-135, and the decoder, fse_decompress.c:108-135)."""
    size = 1 << table_log
    mask, step, high = size - 1, (size >> 1) + (size >> 3) + 3, size - 1
    cells = [0] * size
    for s in range(max_symbol + 1):
        if norm[s] == -1:
            cells[high] = s; high -= 1
    pos = 0
    for s in range(max_symbol + 1):
        for _ in range(max(norm[s], 0)):
            cells[pos] = s
            pos = (pos + step) & mask
            while pos > high:
                pos = (pos + step) & mask
    if pos != 0:
        raise FSEError('spread')
    return cells


def fse_compress(data):
    """FSE_compress (fse_compress.c:648-714): bytes, or 0 (not compressible) / 1 (one repeated symbol)."""
    src = np.frombuffer(bytes(data), dtype=np.uint8)
    n = len(src)
    if n <= 1:
        return 0
    count = np.bincount(src, minlength=256)
    max_symbol = int(np.flatnonzero(count)[-1])
    largest = int(count.max())
    if largest == n:
        return 1
    if largest == 1 or largest < (n >> 7):
        return 0
    table_log = _optimal_table_log(n, max_symbol)
    norm = _normalize(count, n, max_symbol, table_log)
    head = _write_ncount(norm, max_symbol, table_log)
    # encoding tables (fse_compress.c:85-170)
    size = 1 << table_log
    cells = _spread(norm, max_symbol, table_log)
    cumul, acc = [0] * 257, 0
    for s in range(max_symbol + 1):
        cumul[s] = acc
        acc += 1 if norm[s] == -1 else norm[s]
    next_state = [0] * size
    for u in range(size):
        s = cells[u]
        next_state[cumul[s]] = size + u
        cumul[s] += 1
    delta_bits, find_state, total = [0] * 256, [0] * 256, 0
    for s in range(max_symbol + 1):
        p = norm[s]
        if p == 0:
            delta_bits[s] = (((table_log + 1) << 16) - size) & M32
        elif p in (-1, 1):
            delta_bits[s] = ((table_log << 16) - size) & M32
            find_state[s] = total - 1
            total += 1
        else:
            ob = table_log - _highbit(p - 1)
            delta_bits[s] = ((ob << 16) - (p << ob)) & M32
            find_state[s] = total - p
            total += p
    if n <= 2:
        return 0
    # payload (fse_compress.c:558-620): last symbol first; even positions on state 1, odd positions on state 2
    acc_bits, have, body = 0, 0, bytearray()

    def put(v, k):
        nonlocal acc_bits, have
        acc_bits |= (v & ((1 << k) - 1)) << have
        have += k
        while have >= 8:
            body.append(acc_bits & 0xFF)
            acc_bits >>= 8; have -= 8
    state = [0, 0]
    for which, i in ((0, n - 1), (1, n - 2)) if n & 1 else ((1, n - 1), (0, n - 2)):
        d = delta_bits[src[i]]
        k = ((d + (1 << 15)) & M32) >> 16
        v = ((k << 16) - d) & M32
        state[which] = next_state[(v >> k) + find_state[src[i]]]
    for i in range(n - 3, -1, -1):
        which = i & 1
        sym = int(src[i])
        k = ((state[which] + delta_bits[sym]) & M32) >> 16
        put(state[which], k)
        state[which] = next_state[(state[which] >> k) + find_state[sym]]
    put(state[1], table_log)
    put(state[0], table_log)
    put(1, 1)
    if have:
        body.append(acc_bits & 0xFF)
    out = head + bytes(body)
    return 0 if len(out) >= n - 1 else out


def _read_ncount(data):                                  # entropy_common.c:60-167
    hb = len(data)
    buf = bytes(data) + b'\0' * 8
    end = max(hb, 4)

    def le32(i):
        return int.from_bytes(buf[i:i + 4] if i + 4 <= end else (buf[i:end] + b'\0' * 4)[:4], 'little')
    ip = 0
    bits = le32(0)
    nb = (bits & 0xF) + MIN_LOG
    if nb > ABS_MAX_LOG:
        raise FSEError('tableLog too large')
    bits >>= 4
    have = 4
    table_log = nb
    remaining, threshold = (1 << nb) + 1, 1 << nb
    nb += 1
    norm, sym, prev_zero = [0] * 256, 0, False
    while remaining > 1 and sym <= 255:
        if prev_zero:
            n0 = sym
            while (bits & 0xFFFF) == 0xFFFF:
                n0 += 24
                if ip < end - 5:
                    ip += 2
                    bits = le32(ip) >> have
                else:
                    bits >>= 16; have += 16
            while (bits & 3) == 3:
                n0 += 3; bits >>= 2; have += 2
            n0 += bits & 3
            have += 2
            if n0 > 255:
                raise FSEError('maxSymbolValue too small')
            sym = max(sym, n0)
            if ip <= end - 7 or ip + (have >> 3) <= end - 4:
                ip += have >> 3
                have &= 7
                bits = le32(ip) >> have
            else:
                bits >>= 2
        mx = (2 * threshold - 1) - remaining
        if (bits & (threshold - 1)) < mx:
            c = bits & (threshold - 1)
            have += nb - 1
        else:
            c = bits & (2 * threshold - 1)
            if c >= threshold:
                c -= mx
            have += nb
        c -= 1
        remaining -= abs(c)
        norm[sym] = c
        sym += 1
        prev_zero = c == 0
        while remaining < threshold:
            nb -= 1; threshold >>= 1
        if ip <= end - 7 or ip + (have >> 3) <= end - 4:
            ip += have >> 3
            have &= 7
        else:
            have -= 8 * (end - 4 - ip)
            ip = end - 4
        bits = le32(ip) >> (have & 31)
    if remaining != 1 or have > 32:
        raise FSEError('corrupted header')
    ip += (have + 7) >> 3
    if hb < 4 and ip > hb:
        raise FSEError('corrupted header')
    return norm, sym - 1, table_log, ip


def fse_decompress(data, cap):
    """FSE_decompress (fse_decompress.c:196-302) into a buffer of `cap` bytes; raises FSEError where the library reports an error."""
    data = bytes(data)
    norm, max_symbol, table_log, head = _read_ncount(data)
    if table_log > MAX_LOG:
        raise FSEError('tableLog too large')
    size = 1 << table_log
    cells = _spread(norm, max_symbol, table_log)
    nxt = [1 if norm[s] == -1 else norm[s] for s in range(256)]
    table = []
    for u in range(size):
        s = cells[u]
        v = nxt[s]
        nxt[s] += 1
        k = table_log - _highbit(v)
        table.append((((v << k) - size) & 0xFFFF, s, k))
    src = data[head:]
    n = len(src)
    if n < 1:
        raise FSEError('srcSize wrong')
    if src[-1] == 0:
        raise FSEError('end mark missing')
    pad = src + b'\0' * 8
    if n >= 8:
        ptr, used = n - 8, 8 - _highbit(src[-1])
    else:
        ptr, used = 0, 8 - _highbit(src[-1]) + (8 - n) * 8
    box = int.from_bytes(pad[ptr:ptr + 8], 'little')

    def take(k):
        nonlocal used
        v = ((((box << (used & 63)) & M64) >> 1) >> ((63 - k) & 63))
        used += k
        return v

    def refill():                                        # 0 unfinished, 1 end of buffer, 2 completed, 3 overflow (bitstream.h:407-437)
        nonlocal ptr, used, box
        if used > 64:
            return 3
        if ptr >= 8:
            ptr -= used >> 3
            used &= 7
            box = int.from_bytes(pad[ptr:ptr + 8], 'little')
            return 0
        if ptr == 0:
            return 1 if used < 64 else 2
        nbytes, res = used >> 3, 0
        if ptr - nbytes < 0:
            nbytes, res = ptr, 1
        ptr -= nbytes
        used -= nbytes * 8
        box = int.from_bytes(pad[ptr:ptr + 8], 'little')
        return res
    state = [0, 0]
    for i in (0, 1):
        state[i] = take(table_log)
        refill()

    def pop(i):
        new, s, k = table[state[i]]
        state[i] = new + take(k)
        return s
    out = bytearray()
    while True:
        more = refill() == 0
        if not (more and len(out) < cap - 3):
            break
        out.extend((pop(0), pop(1), pop(0), pop(1)))
    while True:
        if len(out) > cap - 2:
            raise FSEError('dstSize too small')
        out.append(pop(0))
        if refill() == 3:
            out.append(pop(1))
            break
        if len(out) > cap - 2:
            raise FSEError('dstSize too small')
        out.append(pop(1))
        if refill() == 3:
            out.append(pop(0))
            break
    return bytes(out)


# ------------------------------------------------------------------------------------------------ the l3ic container
def vq(values, code_book):
    """scipy.cluster.vq.vq for 1-D observations: index of the nearest code (first minimum), compression/codec.py:127."""
    v = np.asarray(values, dtype=np.float64).reshape(-1, 1)
    c = np.asarray(code_book, dtype=np.float64).reshape(1, -1)
    return np.argmin((v - c) ** 2, axis=1)


def l3ic_compress(batch_z, code_book, compress=fse_compress):
    """compression/codec.py:87-186 from the quantised latent (1,h,w,c) on: header, length table, coded layers."""
    batch_z = np.asarray(batch_z)
    assert batch_z.ndim == 4 and batch_z.shape[0] == 1
    if len(code_book) > 256:
        raise ValueError('Code-books with more than 256 centers are not supported')
    shape = np.array(batch_z.shape[1:], dtype=np.uint8)
    layers = []
    for n in range(int(shape[-1])):
        indices = vq(batch_z[:, :, :, n].reshape(-1), code_book)
        raw = bytes(indices.astype(np.uint8))
        r = compress(raw)
        if r == 1:
            r = np.uint16(len(indices)).tobytes() + np.uint8(indices[0]).tobytes()
        elif r == 0:
            r = raw
        if len(r) == 1:
            raise ValueError('Layer {} data compresses to a single byte? Something is wrong!'.format(n))
        layers.append(r)
    lengths = np.array([len(x) for x in layers], dtype=np.uint16).tobytes()
    coded = compress(lengths)
    if coded == 1:
        raise FSEError('input data is a repetition of a single byte')       # uncaught in the reference
    if coded == 0:
        coded = lengths
    return shape.tobytes() + np.uint16(len(coded)).tobytes() + coded + b''.join(layers)


def l3ic_decompress(stream, code_book, decompress=fse_decompress):
    """compression/codec.py:189-255 up to the quantised latent (1,h,w,c) float array."""
    stream = bytes(stream)
    h, w, c = (int(v) for v in stream[:3])
    nl = int(np.frombuffer(stream[3:5], np.uint16)[0])
    coded = stream[5:5 + nl]
    if nl != 2 * c:
        lengths = np.frombuffer(decompress(coded, 10 * len(coded)), dtype=np.uint16)
    else:
        lengths = np.frombuffer(coded, dtype=np.uint16)
    code_book = np.asarray(code_book)
    z = np.zeros((1, h, w, c), dtype=np.float64)
    o = 5 + nl
    for n in range(c):
        layer = stream[o:o + int(lengths[n])]
        o += int(lengths[n])
        if len(layer) == 3:
            data = layer[-1:] * int(np.frombuffer(layer[:2], dtype=np.uint16)[0])
        elif len(layer) == h * w:
            data = layer
        else:
            data = decompress(layer, 4 * h * w)
        z[0, :, :, n] = code_book[np.frombuffer(data, np.uint8)].reshape((h, w))
    return z
