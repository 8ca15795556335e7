// Finite State Entropy (tANS) coder, bit-compatible with the FSE library the reference vendors for its l3ic bit-stream
// (pyfse/pyfse.pyx:24-72 -> FSE_compress / FSE_decompress; pyfse/FiniteStateEntropy/lib/fse_compress.c:648-714, fse_decompress.c:262-302,
// entropy_common.c:60-167, bitstream.h on a 64-bit host). SURVEY 8f N3.
//
// One stream (a latent layer of one image: a few hundred to a few thousand bytes) is inherently serial — every symbol's bit count depends
// on the running state — so the device maps ONE WARP PER STREAM, thousands of streams per launch: the lanes cooperate on the quantisation,
// histogram and copies, lane 0 walks the serial parts out of shared-memory tables. Everything here is `host device` so that the very same
// source is compiled by g++ into a test harness (tests/fse_host_harness.cpp) and checked byte for byte against the reference library on
// the CPU; the product only ever runs the device instantiation (csrc/l3ic.cu).
//
// Return convention of compress(): > 1 size of the coded stream, 0 = not compressible, 1 = a single repeated symbol (caller uses RLE),
// < 0 = error. decompress(): >= 0 number of decoded bytes, < 0 = error.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FSE_HD __host__ __device__ __forceinline__
#else
#define FSE_HD inline
#endif

namespace fse {

enum : int {
    kErrGeneric = -1, kErrDstTooSmall = -2, kErrTableLogTooLarge = -3, kErrMaxSymbolTooSmall = -4, kErrSrcSizeWrong = -5,
    kErrCorruption = -6
};
constexpr int kMinTableLog = 5, kMaxTableLog = 12, kDefaultTableLog = 11, kAbsMaxTableLog = 15;
constexpr int kEncTableLog = kDefaultTableLog;          // the encoder never exceeds the default (min-bits rule caps at 9 for bytes)
constexpr int kNCountBound = 512;

FSE_HD uint32_t compress_bound(uint32_t n) { return kNCountBound + n + (n >> 7); }

FSE_HD int highbit(uint32_t v) {          // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

FSE_HD uint32_t table_step(uint32_t size) { return (size >> 1) + (size >> 3) + 3; }

struct EncScratch {
    uint32_t count[256];
    int16_t norm[256];
    int32_t find_state[256];                // per symbol: offset into next_state
    uint32_t delta_bits[256];               // per symbol: (maxBitsOut << 16) - minStatePlus
    uint16_t next_state[1 << kEncTableLog];  // sorted by symbol
    uint8_t spread[1 << kEncTableLog];
};

struct DecScratch {
    int16_t norm[256];
    uint16_t symbol_next[256];
    uint32_t cell[1 << kMaxTableLog];        // newState | symbol << 16 | nbBits << 24
};

// ------------------------------------------------------------------------------------------------ statistics
// Byte histogram (hist.c:49-73); the device builds it cooperatively instead (csrc/l3ic.cu).
FSE_HD void histogram(const uint8_t* src, uint32_t n, uint32_t* count) {
    for (int s = 0; s < 256; ++s) count[s] = 0;
    for (uint32_t i = 0; i < n; ++i) count[src[i]]++;
}

// Largest count and highest byte value present.
FSE_HD uint32_t histogram_summary(const uint32_t* count, uint32_t* max_symbol) {
    uint32_t top = 255, largest = 0;
    while (top > 0 && count[top] == 0) --top;
    for (uint32_t s = 0; s <= top; ++s) largest = count[s] > largest ? count[s] : largest;
    *max_symbol = top;
    return largest;
}

FSE_HD uint32_t min_table_log(uint32_t n, uint32_t max_symbol) {
    const uint32_t by_src = (uint32_t)highbit(n - 1) + 1, by_sym = (uint32_t)highbit(max_symbol) + 2;
    return by_src < by_sym ? by_src : by_sym;
}

// fse_compress.c:327-345 with maxTableLog = 11, minus = 2 (unsigned arithmetic on purpose: tiny inputs wrap and keep the default).
FSE_HD uint32_t optimal_table_log(uint32_t n, uint32_t max_symbol) {
    const uint32_t src_bits = (uint32_t)highbit(n - 1) - 2u;
    uint32_t t = kDefaultTableLog;
    const uint32_t need = min_table_log(n, max_symbol);
    if (src_bits < t) t = src_bits;
    if (need > t) t = need;
    if (t < (uint32_t)kMinTableLog) t = kMinTableLog;
    if (t > (uint32_t)kMaxTableLog) t = kMaxTableLog;
    return t;
}

// Fallback normalisation (fse_compress.c:353-447).
FSE_HD int normalize_fallback(int16_t* norm, uint32_t table_log, const uint32_t* count, uint64_t total, uint32_t max_symbol) {
    const int16_t kOpen = -2;
    uint32_t given = 0;
    const uint32_t low = (uint32_t)(total >> table_log);
    uint32_t low_one = (uint32_t)((total * 3) >> (table_log + 1));
    for (uint32_t s = 0; s <= max_symbol; ++s) {
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= low) { norm[s] = -1; ++given; total -= count[s]; continue; }
        if (count[s] <= low_one) { norm[s] = 1; ++given; total -= count[s]; continue; }
        norm[s] = kOpen;
    }
    uint32_t left = (1u << table_log) - given;
    if ((total / left) > low_one) {
        low_one = (uint32_t)((total * 3) / (left * 2));
        for (uint32_t s = 0; s <= max_symbol; ++s)
            if (norm[s] == kOpen && count[s] <= low_one) { norm[s] = 1; ++given; total -= count[s]; }
        left = (1u << table_log) - given;
    }
    if (given == max_symbol + 1) {
        uint32_t best = 0, best_count = 0;
        for (uint32_t s = 0; s <= max_symbol; ++s)
            if (count[s] > best_count) { best = s; best_count = count[s]; }
        norm[best] = (int16_t)(norm[best] + (int16_t)left);
        return 0;
    }
    if (total == 0) {
        for (uint32_t s = 0; left > 0; s = (s + 1) % (max_symbol + 1))
            if (norm[s] > 0) { --left; norm[s]++; }
        return 0;
    }
    const uint64_t shift = 62 - table_log;
    const uint64_t mid = (1ULL << (shift - 1)) - 1;
    const uint64_t r_step = (((1ULL << shift) * left) + mid) / total;
    uint64_t run = mid;
    for (uint32_t s = 0; s <= max_symbol; ++s) {
        if (norm[s] != kOpen) continue;
        const uint64_t end = run + (uint64_t)count[s] * r_step;
        const uint32_t weight = (uint32_t)(end >> shift) - (uint32_t)(run >> shift);
        if (weight < 1) return kErrGeneric;
        norm[s] = (int16_t)weight;
        run = end;
    }
    return 0;
}

// fse_compress.c:450-508. Returns table_log, 0 for the single-symbol case, < 0 on error.
FSE_HD int normalize(int16_t* norm, uint32_t table_log, const uint32_t* count, uint32_t total, uint32_t max_symbol) {
    if (table_log < (uint32_t)kMinTableLog) return kErrGeneric;
    if (table_log > (uint32_t)kMaxTableLog) return kErrTableLogTooLarge;
    if (table_log < min_table_log(total, max_symbol)) return kErrGeneric;
    const uint32_t beat[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    const uint64_t scale = 62 - table_log;
    const uint64_t step = (1ULL << 62) / total;
    const uint64_t v_step = 1ULL << (scale - 20);
    int left = 1 << table_log;
    uint32_t largest = 0;
    int16_t largest_p = 0;
    const uint32_t low = total >> table_log;
    for (uint32_t s = 0; s <= max_symbol; ++s) {
        if (count[s] == total) return 0;
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= low) { norm[s] = -1; --left; continue; }
        const uint64_t scaled = (uint64_t)count[s] * step;
        int16_t p = (int16_t)(scaled >> scale);
        if (p < 8) {
            const uint64_t rest = v_step * beat[p];
            p = (int16_t)(p + ((scaled - ((uint64_t)p << scale)) > rest ? 1 : 0));
        }
        if (p > largest_p) { largest_p = p; largest = s; }
        norm[s] = p;
        left -= p;
    }
    if (-left >= (norm[largest] >> 1)) {
        const int e = normalize_fallback(norm, table_log, count, total, max_symbol);
        if (e < 0) return e;
    } else {
        norm[largest] = (int16_t)(norm[largest] + (int16_t)left);
    }
    return (int)table_log;
}

// Table description header (fse_compress.c:204-285). `cap` bytes are available at out; returns the header size.
FSE_HD int write_ncount(uint8_t* out, uint32_t cap, const int16_t* norm, uint32_t max_symbol, uint32_t table_log) {
    if (table_log > (uint32_t)kMaxTableLog) return kErrTableLogTooLarge;
    if (table_log < (uint32_t)kMinTableLog) return kErrGeneric;
    uint32_t o = 0;
    const int size = 1 << table_log;
    int remaining = size + 1, threshold = size, nb = (int)table_log + 1;
    uint32_t bits = table_log - kMinTableLog;
    int have = 4;
    uint32_t sym = 0;
    bool prev_zero = false;
#define FSE_PUT16()                                              \
    do {                                                         \
        if (o + 2 > cap) return kErrDstTooSmall;                 \
        out[o] = (uint8_t)bits; out[o + 1] = (uint8_t)(bits >> 8); \
        o += 2; bits >>= 16;                                     \
    } while (0)
    while (remaining > 1) {
        if (prev_zero) {
            uint32_t start = sym;
            while (!norm[sym]) ++sym;
            while (sym >= start + 24) { start += 24; bits += 0xFFFFu << have; FSE_PUT16(); }
            while (sym >= start + 3) { start += 3; bits += 3u << have; have += 2; }
            bits += (sym - start) << have;
            have += 2;
            if (have > 16) { FSE_PUT16(); have -= 16; }
        }
        {
            int c = norm[sym++];
            const int max = (2 * threshold - 1) - remaining;
            remaining -= c < 0 ? -c : c;
            ++c;
            if (c >= threshold) c += max;
            bits += (uint32_t)c << have;
            have += nb;
            have -= (c < max);
            prev_zero = (c == 1);
            if (remaining < 1) return kErrGeneric;
            while (remaining < threshold) { --nb; threshold >>= 1; }
        }
        if (have > 16) { FSE_PUT16(); have -= 16; }
    }
    if (o + 2 > cap) return kErrDstTooSmall;
    out[o] = (uint8_t)bits; out[o + 1] = (uint8_t)(bits >> 8);
    o += (uint32_t)(have + 7) / 8;
#undef FSE_PUT16
    if (sym > max_symbol + 1) return kErrGeneric;
    return (int)o;
}

// Encoding tables (fse_compress.c:85-170).
FSE_HD int build_enc_tables(EncScratch& S, uint32_t max_symbol, uint32_t table_log) {
    const uint32_t size = 1u << table_log, mask = size - 1, step = table_step(size);
    uint32_t high = size - 1;
    uint32_t* cumul = S.count;             // the counts are not needed any more: reuse as the running start positions (257 entries needed)
    uint32_t acc = 0;
    // start position of every symbol; low-probability symbols are parked at the top of the spread table
    for (uint32_t s = 0; s <= max_symbol; ++s) {
        const uint32_t at = acc;
        if (S.norm[s] == -1) { acc += 1; S.spread[high--] = (uint8_t)s; }
        else acc += (uint32_t)S.norm[s];
        cumul[s] = at;
    }
    uint32_t pos = 0;
    for (uint32_t s = 0; s <= max_symbol; ++s)
        for (int i = 0; i < S.norm[s]; ++i) {
            S.spread[pos] = (uint8_t)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    if (pos != 0) return kErrGeneric;
    for (uint32_t u = 0; u < size; ++u) {
        const uint8_t s = S.spread[u];
        S.next_state[cumul[s]++] = (uint16_t)(size + u);
    }
    uint32_t total = 0;
    for (uint32_t s = 0; s <= max_symbol; ++s) {
        const int p = S.norm[s];
        if (p == 0) {
            S.delta_bits[s] = ((table_log + 1) << 16) - size;
            S.find_state[s] = 0;
        } else if (p == -1 || p == 1) {
            S.delta_bits[s] = (table_log << 16) - size;
            S.find_state[s] = (int32_t)total - 1;
            total += 1;
        } else {
            const uint32_t out_bits = table_log - (uint32_t)highbit((uint32_t)p - 1);
            S.delta_bits[s] = (out_bits << 16) - ((uint32_t)p << out_bits);
            S.find_state[s] = (int32_t)total - p;
            total += (uint32_t)p;
        }
    }
    return 0;
}

struct BitSink {          // little-endian bit writer; bytes past `cap` are counted but dropped
    uint8_t* base;
    uint32_t cap, pos, have;
    uint64_t acc;
    FSE_HD void put(uint64_t v, uint32_t n) {
        acc |= (v & ((1ULL << n) - 1)) << have;
        have += n;
        while (have >= 8) {
            if (pos < cap) base[pos] = (uint8_t)acc;
            ++pos; acc >>= 8; have -= 8;
        }
    }
    FSE_HD uint32_t finish() {     // end mark, then the size in bytes
        put(1, 1);
        if (have > 0) {
            if (pos < cap) base[pos] = (uint8_t)acc;
            return pos + 1;
        }
        return pos;
    }
};

struct EncState {
    uint32_t v;
    FSE_HD void start(const EncScratch& S, uint32_t sym) {          // first symbol of a state costs no bits (fse.h:523-532)
        const uint32_t d = S.delta_bits[sym];
        const uint32_t nb = (d + (1u << 15)) >> 16;
        v = (nb << 16) - d;
        v = S.next_state[(int32_t)(v >> nb) + S.find_state[sym]];
    }
    FSE_HD void push(const EncScratch& S, BitSink& w, uint32_t sym) {
        const uint32_t nb = (v + S.delta_bits[sym]) >> 16;
        w.put(v, nb);
        v = S.next_state[(int32_t)(v >> nb) + S.find_state[sym]];
    }
};

// Payload (fse_compress.c:558-620): symbols are pushed last to first, even positions on state 1 and odd positions on state 2.
FSE_HD uint32_t encode_payload(uint8_t* dst, uint32_t cap, const uint8_t* src, uint32_t n, const EncScratch& S, uint32_t table_log) {
    if (n <= 2) return 0;
    BitSink w{dst, cap, 0, 0, 0};
    EncState s1, s2;
    uint32_t i = n;
    if (n & 1) { s1.start(S, src[--i]); s2.start(S, src[--i]); }
    else { s2.start(S, src[--i]); s1.start(S, src[--i]); }
    while (i > 0) {
        --i;
        if (i & 1) s2.push(S, w, src[i]); else s1.push(S, w, src[i]);
    }
    w.put(s2.v, table_log);
    w.put(s1.v, table_log);
    return w.finish();
}

// FSE_compress_wksp (fse_compress.c:648-692) after the histogram: S.count must hold the byte counts of src.
// `cap` = bytes available at dst (>= n is enough: anything longer than n - 2 is reported as not compressible anyway).
FSE_HD int compress_counted(uint8_t* dst, uint32_t cap, const uint8_t* src, uint32_t n, EncScratch& S) {
    if (n <= 1) return 0;
    uint32_t max_symbol;
    const uint32_t largest = histogram_summary(S.count, &max_symbol);
    if (largest == n) return 1;
    if (largest == 1) return 0;
    if (largest < (n >> 7)) return 0;
    const uint32_t table_log = optimal_table_log(n, max_symbol);
    const int e = normalize(S.norm, table_log, S.count, n, max_symbol);
    if (e < 0) return e;
    const int head = write_ncount(dst, cap, S.norm, max_symbol, table_log);
    if (head == kErrDstTooSmall) return 0;            // cannot happen in the reference's 512-byte margin unless the result is useless anyway
    if (head < 0) return head;
    const int b = build_enc_tables(S, max_symbol, table_log);
    if (b < 0) return b;
    const uint32_t body = encode_payload(dst + head, cap - (uint32_t)head, src, n, S, table_log);
    if (body == 0) return 0;
    const uint32_t total = (uint32_t)head + body;
    if (total >= n - 1) return 0;
    return (int)total;
}

// ------------------------------------------------------------------------------------------------ decoding
struct ByteSrc {        // bounds-checked little-endian reads; bytes past the end read as zero (only reached on the padded tiny-header path)
    const uint8_t* p;
    uint32_t n;
    FSE_HD uint32_t at(uint32_t i) const { return i < n ? p[i] : 0u; }
    FSE_HD uint32_t le32(uint32_t i) const { return at(i) | at(i + 1) << 8 | at(i + 2) << 16 | at(i + 3) << 24; }
    FSE_HD uint64_t le64(uint32_t i) const { return (uint64_t)le32(i) | (uint64_t)le32(i + 4) << 32; }
};

// entropy_common.c:60-167. Returns the header size; fills norm[0..255], *max_symbol, *table_log.
FSE_HD int read_ncount(int16_t* norm, uint32_t* max_symbol, uint32_t* table_log, const uint8_t* hdr, uint32_t hdr_size) {
    const ByteSrc in{hdr, hdr_size};
    const int end = (int)(hdr_size < 4 ? 4 : hdr_size);      // short headers are processed as if zero-padded to 4 bytes
    int ip = 0;
    uint32_t bits = in.le32(0);
    int nb = (int)(bits & 0xF) + kMinTableLog;
    if (nb > kAbsMaxTableLog) return kErrTableLogTooLarge;
    bits >>= 4;
    int have = 4;
    *table_log = (uint32_t)nb;
    int remaining = (1 << nb) + 1, threshold = 1 << nb;
    ++nb;
    uint32_t sym = 0;
    const uint32_t cap_symbol = *max_symbol;
    bool prev_zero = false;
    while (remaining > 1 && sym <= cap_symbol) {
        if (prev_zero) {
            uint32_t n0 = sym;
            while ((bits & 0xFFFF) == 0xFFFF) {
                n0 += 24;
                if (ip < end - 5) { ip += 2; bits = in.le32((uint32_t)ip) >> have; }
                else { bits >>= 16; have += 16; }
            }
            while ((bits & 3) == 3) { n0 += 3; bits >>= 2; have += 2; }
            n0 += bits & 3;
            have += 2;
            if (n0 > cap_symbol) return kErrMaxSymbolTooSmall;
            while (sym < n0) norm[sym++] = 0;
            if (ip <= end - 7 || ip + (have >> 3) <= end - 4) {
                ip += have >> 3;
                have &= 7;
                bits = in.le32((uint32_t)ip) >> have;
            } else {
                bits >>= 2;
            }
        }
        {
            const int max = (2 * threshold - 1) - remaining;
            int c;
            if ((bits & (uint32_t)(threshold - 1)) < (uint32_t)max) {
                c = (int)(bits & (uint32_t)(threshold - 1));
                have += nb - 1;
            } else {
                c = (int)(bits & (uint32_t)(2 * threshold - 1));
                if (c >= threshold) c -= max;
                have += nb;
            }
            --c;
            remaining -= c < 0 ? -c : c;
            norm[sym++] = (int16_t)c;
            prev_zero = !c;
            while (remaining < threshold) { --nb; threshold >>= 1; }
            if (ip <= end - 7 || ip + (have >> 3) <= end - 4) {
                ip += have >> 3;
                have &= 7;
            } else {
                have -= 8 * (end - 4 - ip);
                ip = end - 4;
            }
            bits = in.le32((uint32_t)ip) >> (have & 31);
        }
    }
    if (remaining != 1) return kErrCorruption;
    if (have > 32) return kErrCorruption;
    for (uint32_t s = sym; s <= cap_symbol; ++s) norm[s] = 0;
    *max_symbol = sym - 1;
    ip += (have + 7) >> 3;
    if (hdr_size < 4 && (uint32_t)ip > hdr_size) return kErrCorruption;
    return ip;
}

// fse_decompress.c:91-148.
FSE_HD int build_dec_table(DecScratch& S, uint32_t max_symbol, uint32_t table_log) {
    if (max_symbol > 255) return kErrGeneric;
    if (table_log > (uint32_t)kMaxTableLog) return kErrTableLogTooLarge;
    const uint32_t size = 1u << table_log, mask = size - 1, step = table_step(size);
    uint32_t high = size - 1;
    for (uint32_t s = 0; s <= max_symbol; ++s) {
        if (S.norm[s] == -1) { S.cell[high--] = s << 16; S.symbol_next[s] = 1; }
        else S.symbol_next[s] = (uint16_t)S.norm[s];
    }
    uint32_t pos = 0;
    for (uint32_t s = 0; s <= max_symbol; ++s)
        for (int i = 0; i < S.norm[s]; ++i) {
            S.cell[pos] = s << 16;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    if (pos != 0) return kErrGeneric;
    for (uint32_t u = 0; u < size; ++u) {
        const uint32_t s = S.cell[u] >> 16;
        const uint32_t next = S.symbol_next[s]++;
        const uint32_t nb = table_log - (uint32_t)highbit(next);
        S.cell[u] = (((next << nb) - size) & 0xFFFF) | s << 16 | nb << 24;
    }
    return 0;
}

struct BitSource {        // backward bit reader (bitstream.h:259-452, 64-bit container)
    ByteSrc in;
    int ptr;              // byte offset of the container
    uint32_t used;        // bits consumed from the container
    uint64_t box;
    enum Status { kUnfinished = 0, kEndOfBuffer = 1, kCompleted = 2, kOverflow = 3 };

    FSE_HD int open(const uint8_t* p, uint32_t n) {
        in = ByteSrc{p, n};
        if (n < 1) return kErrSrcSizeWrong;
        const uint32_t last = p[n - 1];
        if (n >= 8) {
            ptr = (int)n - 8;
            box = in.le64((uint32_t)ptr);
            if (last == 0) return kErrGeneric;
            used = 8 - (uint32_t)highbit(last);
        } else {
            ptr = 0;
            box = in.le64(0);             // zero-extended
            if (last == 0) return kErrCorruption;
            used = 8 - (uint32_t)highbit(last) + (8 - n) * 8;
        }
        return 0;
    }
    FSE_HD uint32_t take(uint32_t nb) {
        const uint64_t v = ((box << (used & 63)) >> 1) >> ((63 - nb) & 63);
        used += nb;
        return (uint32_t)v;
    }
    FSE_HD Status refill() {
        if (used > 64) return kOverflow;
        if (ptr >= 8) {
            ptr -= (int)(used >> 3);
            used &= 7;
            box = in.le64((uint32_t)ptr);
            return kUnfinished;
        }
        if (ptr == 0) return used < 64 ? kEndOfBuffer : kCompleted;
        uint32_t bytes = used >> 3;
        Status r = kUnfinished;
        if (ptr - (int)bytes < 0) { bytes = (uint32_t)ptr; r = kEndOfBuffer; }
        ptr -= (int)bytes;
        used -= bytes * 8;
        box = in.le64((uint32_t)ptr);
        return r;
    }
};

struct DecState {
    uint32_t v;
    FSE_HD void start(BitSource& b, uint32_t table_log) { v = b.take(table_log); b.refill(); }
    FSE_HD uint8_t pop(const DecScratch& S, BitSource& b) {
        const uint32_t c = S.cell[v];
        v = (c & 0xFFFF) + b.take(c >> 24);
        return (uint8_t)(c >> 16);
    }
};

// fse_decompress.c:196-258 (the two variants differ only for zero-bit reads, which take() handles).
FSE_HD int decode_payload(uint8_t* dst, uint32_t cap, const uint8_t* src, uint32_t n, const DecScratch& S, uint32_t table_log) {
    BitSource b;
    const int e = b.open(src, n);
    if (e < 0) return e;
    DecState s1, s2;
    s1.start(b, table_log);
    s2.start(b, table_log);
    int o = 0;
    const int limit = (int)cap - 3;
    for (;;) {
        const bool more = b.refill() == BitSource::kUnfinished;
        if (!(more && o < limit)) break;
        dst[o] = s1.pop(S, b);
        dst[o + 1] = s2.pop(S, b);
        dst[o + 2] = s1.pop(S, b);
        dst[o + 3] = s2.pop(S, b);
        o += 4;
    }
    for (;;) {
        if (o > (int)cap - 2) return kErrDstTooSmall;
        dst[o++] = s1.pop(S, b);
        if (b.refill() == BitSource::kOverflow) { dst[o++] = s2.pop(S, b); break; }
        if (o > (int)cap - 2) return kErrDstTooSmall;
        dst[o++] = s2.pop(S, b);
        if (b.refill() == BitSource::kOverflow) { dst[o++] = s1.pop(S, b); break; }
    }
    return o;
}

// FSE_decompress (fse_decompress.c:273-302).
FSE_HD int decompress(uint8_t* dst, uint32_t cap, const uint8_t* src, uint32_t n, DecScratch& S) {
    uint32_t table_log = 0, max_symbol = 255;
    const int head = read_ncount(S.norm, &max_symbol, &table_log, src, n);
    if (head < 0) return head;
    if (table_log > (uint32_t)kMaxTableLog) return kErrTableLogTooLarge;
    const int e = build_dec_table(S, max_symbol, table_log);
    if (e < 0) return e;
    return decode_payload(dst, cap, src + head, n - (uint32_t)head, S, table_log);
}

}  // namespace fse
