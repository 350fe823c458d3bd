// tdt_inflate.h -- raw DEFLATE (RFC 1951) decoder for whole in-memory streams of known inflated size: the BGZF blocks
// of a BAM file (csrc/tdt_bam.cpp).  Host code, header only.
//
// zlib's inflate is a resumable state machine that may be handed one byte at a time; a BGZF block is a complete stream
// of at most 64 KiB whose inflated size its trailer states.  That allows the usual fast-decoder layout: a 64-bit bit
// buffer refilled with one unaligned 8-byte load, 11-bit (literal / length) and 8-bit (distance) first-level tables whose
// entries carry symbol, code length and extra-bit count in one word, up to three literals per refill, matches copied
// eight bytes at a time, and all bounds checks hoisted out of the inner loop (a checked byte-wise loop finishes the last
// ~300 bytes of output / 16 bytes of input).  (Measured and dropped: two-literal table entries -- BAM streams mix literals
// and short matches symbol by symbol, the mispredicted literal / match branch is the cost, not the lookups: 175 vs 175 MB/s
// on BAM-like level-6 streams, 216 vs 229 on 40-valued qualities; table widths 10-12 / 8-9 bits: within noise.)
// Anything unusual returns false; the caller then lets zlib decide and keeps
// its CRC check either way, so a decoding mistake here can cost time, never a wrong byte.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <zlib.h>

namespace tdtz {

typedef uint32_t u32;
typedef uint64_t u64;

constexpr int LL_BITS = 11, D_BITS = 8;          // first-level table widths
constexpr int LL_CAP = (1 << LL_BITS) + 288 * 16, D_CAP = (1 << D_BITS) + 32 * 128;
// table entry: value (literal byte / length or distance base / subtable start) << 16 | flags | code length << 8 | extra bits
constexpr u32 E_LIT = 0x8000u, E_EXC = 0x4000u, E_SUB = 0x2000u, E_ERR = 0x1000u;
constexpr u32 E_INVALID = E_EXC | E_ERR;         // E_EXC alone: end of block; E_EXC | E_SUB: subtable pointer

struct Inflater {
    u32 ll[LL_CAP], dd[D_CAP];
    u32 fixed_ll[LL_CAP], fixed_dd[D_CAP];
    bool fixed_ready = false;
};

inline u64 load64(const uint8_t *p) { u64 v; memcpy(&v, p, 8); return v; }   // little-endian hosts only (x86-64, aarch64)
inline void store64(uint8_t *p, u64 v) { memcpy(p, &v, 8); }

inline u32 reverse_bits(u32 code, int len) {
    u32 r = 0;
    for (int i = 0; i < len; i++) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
}

// Canonical Huffman code -> decode table.  sym_entry[s]: the entry of symbol s without its code length.
inline bool build_table(const uint8_t *lens, int nsyms, int bits, u32 *table, int cap, const u32 *sym_entry) {
    int count[16] = {0};
    for (int s = 0; s < nsyms; s++) count[lens[s]]++;
    const int main_size = 1 << bits;
    for (int i = 0; i < main_size; i++) table[i] = E_INVALID;
    if (count[0] == nsyms) return true;          // no code at all (a block of literals has no distance code)
    int max_len = 15;
    while (count[max_len] == 0) max_len--;
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left = (left << 1) - count[l];
        if (left < 0) return false;              // over-subscribed
    }
    if (left > 0 && max_len != 1) return false;  // incomplete: only the single one-bit code is legal
    int offs[17];
    u32 next_code[16];
    offs[1] = 0;
    u32 code = 0;
    for (int l = 1; l <= 15; l++) {
        next_code[l] = code;
        code = (code + (u32)count[l]) << 1;
        offs[l + 1] = offs[l] + count[l];
    }
    uint16_t sorted[320];
    {
        int at[17];
        for (int l = 1; l <= 16; l++) at[l] = offs[l];
        for (int s = 0; s < nsyms; s++)
            if (lens[s]) sorted[at[lens[s]]++] = (uint16_t)s;
    }
    const int n = offs[16];
    int used = main_size;
    for (int k = 0; k < n;) {
        const int s = sorted[k], l = lens[s];
        const u32 c = next_code[l]++;
        if (l <= bits) {
            const u32 e = sym_entry[s] | ((u32)l << 8);
            for (u32 i = reverse_bits(c, l); i < (u32)main_size; i += 1u << l) table[i] = e;
            k++;
            continue;
        }
        // codes longer than the first level: the ones that share their first `bits` bits are neighbours in canonical
        // order and get one subtable as wide as the longest of them
        const u32 prefix = c >> (l - bits);
        int j = k, long_len = l;
        {
            u32 nc[16];
            for (int t = 1; t <= 15; t++) nc[t] = next_code[t];
            nc[l] = c;                           // re-walk from k without consuming
            for (j = k; j < n; j++) {
                const int lj = lens[sorted[j]];
                if ((nc[lj] >> (lj - bits)) != prefix) break;
                nc[lj]++;
                long_len = lj;
            }
        }
        const int sub_bits = long_len - bits;
        if (used + (1 << sub_bits) > cap) return false;
        u32 *sub = table + used;
        for (int i = 0; i < (1 << sub_bits); i++) sub[i] = E_INVALID;
        table[reverse_bits(prefix, bits)] = ((u32)used << 16) | E_EXC | E_SUB | (u32)sub_bits;
        used += 1 << sub_bits;
        next_code[l]--;                          // hand the first code back, then consume [k, j) in order
        for (; k < j; k++) {
            const int sk = sorted[k], lk = lens[sk], sl = lk - bits;
            const u32 ck = next_code[lk]++;
            const u32 e = sym_entry[sk] | ((u32)sl << 8);
            for (u32 i = reverse_bits(ck & ((1u << sl) - 1u), sl); i < (1u << sub_bits); i += 1u << sl) sub[i] = e;
        }
    }
    return true;
}

struct SymTables {
    u32 ll[288], dd[32];
    SymTables() {
        static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115,
                                           131, 163, 195, 227, 258};
        static const uint8_t lextra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025,
                                           1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dextra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12,
                                           13, 13};
        for (int s = 0; s < 256; s++) ll[s] = ((u32)s << 16) | E_LIT;
        ll[256] = E_EXC;
        for (int s = 257; s < 286; s++) ll[s] = ((u32)lbase[s - 257] << 16) | lextra[s - 257];
        ll[286] = ll[287] = E_INVALID;           // have code lengths in the fixed code, never occur in data
        for (int s = 0; s < 30; s++) dd[s] = ((u32)dbase[s] << 16) | dextra[s];
        dd[30] = dd[31] = E_INVALID;
    }
};

// in[0, in_len): a complete raw deflate stream; out[0, out_len): exactly its inflated bytes.  -> true when the stream
// decoded cleanly to exactly out_len bytes.
inline bool inflate_raw(Inflater &st, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    static const SymTables sym;
    const uint8_t *const in_end = in + in_len;
    uint8_t *const out_start = out, *const out_end = out + out_len;
    u64 bits = 0;
    int bitcnt = 0;          // valid bits in `bits` (the bits above them mirror the bytes at `in`, or are zero)
    size_t pad = 0;          // zero bytes fed beyond the end of the input
    auto refill_slow = [&]() {
        while (bitcnt <= 56) {
            if (in < in_end) bits |= (u64)*in++ << bitcnt;
            else pad++;
            bitcnt += 8;
        }
    };
#define TDTZ_CONSUME(n) do { bits >>= (n); bitcnt -= (n); } while (0)
    bool last = false;
    while (!last) {
        refill_slow();
        last = bits & 1;
        const u32 type = (u32)(bits >> 1) & 3u;
        TDTZ_CONSUME(3);
        const u32 *ll, *dd;
        if (type == 0) {     // stored: back to a byte boundary, LEN / NLEN, the bytes
            TDTZ_CONSUME(bitcnt & 7);
            // the bytes still in the buffer belong to the stream: step back
            const size_t whole = (size_t)bitcnt >> 3;
            if (pad > whole) return false;
            in -= whole - pad;
            pad = 0;
            bits = 0;
            bitcnt = 0;
            if (in_end - in < 4) return false;
            const u32 len = in[0] | ((u32)in[1] << 8), nlen = in[2] | ((u32)in[3] << 8);
            in += 4;
            if ((len ^ 0xffffu) != nlen || (size_t)(in_end - in) < len || (size_t)(out_end - out) < len) return false;
            memcpy(out, in, len);
            in += len;
            out += len;
            continue;
        } else if (type == 1) {
            if (!st.fixed_ready) {
                uint8_t lens[288 + 32];
                for (int s = 0; s < 144; s++) lens[s] = 8;
                for (int s = 144; s < 256; s++) lens[s] = 9;
                for (int s = 256; s < 280; s++) lens[s] = 7;
                for (int s = 280; s < 288; s++) lens[s] = 8;
                for (int s = 0; s < 32; s++) lens[288 + s] = 5;
                if (!build_table(lens, 288, LL_BITS, st.fixed_ll, LL_CAP, sym.ll) ||
                    !build_table(lens + 288, 32, D_BITS, st.fixed_dd, D_CAP, sym.dd))
                    return false;
                st.fixed_ready = true;
            }
            ll = st.fixed_ll;
            dd = st.fixed_dd;
        } else if (type == 2) {
            const int hlit = (int)(bits & 31) + 257, hdist = (int)((bits >> 5) & 31) + 1, hclen = (int)((bits >> 10) & 15) + 4;
            TDTZ_CONSUME(14);
            if (hlit > 286 || hdist > 30) return false;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t cl[19] = {0};
            for (int i = 0; i < hclen; i++) {
                refill_slow();
                cl[order[i]] = (uint8_t)(bits & 7);
                TDTZ_CONSUME(3);
            }
            u32 cl_sym[19], cl_table[1 << 7];
            for (int s = 0; s < 19; s++) cl_sym[s] = (u32)s << 16;
            if (!build_table(cl, 19, 7, cl_table, 1 << 7, cl_sym)) return false;
            uint8_t lens[286 + 30 + 138];
            int i = 0;
            while (i < hlit + hdist) {
                refill_slow();
                const u32 e = cl_table[bits & 127];
                if (e & E_EXC) return false;
                TDTZ_CONSUME((e >> 8) & 15);
                const u32 s = e >> 16;
                if (s < 16) {
                    lens[i++] = (uint8_t)s;
                } else {
                    int rep;
                    uint8_t v = 0;
                    if (s == 16) {
                        if (i == 0) return false;
                        v = lens[i - 1];
                        rep = 3 + (int)(bits & 3);
                        TDTZ_CONSUME(2);
                    } else if (s == 17) {
                        rep = 3 + (int)(bits & 7);
                        TDTZ_CONSUME(3);
                    } else {
                        rep = 11 + (int)(bits & 127);
                        TDTZ_CONSUME(7);
                    }
                    if (i + rep > hlit + hdist) return false;
                    memset(lens + i, v, (size_t)rep);
                    i += rep;
                }
            }
            if (lens[256] == 0) return false;    // no end-of-block code
            if (!build_table(lens, hlit, LL_BITS, st.ll, LL_CAP, sym.ll) ||
                !build_table(lens + hlit, hdist, D_BITS, st.dd, D_CAP, sym.dd))
                return false;
            ll = st.ll;
            dd = st.dd;
        } else {
            return false;
        }
        if (pad > ((size_t)bitcnt >> 3)) return false;   // the header ran past the end of the input

        // ---- fast loop: >= 16 input bytes and >= 274 output bytes of room, no checks inside -----------------------
        bool done = false;
        if (in_len >= 16 && out_len >= 274) {
            const uint8_t *const in_fast = in_end - 16;
            uint8_t *const out_fast = out_end - 274;
            // invariant at the top of the loop: the buffer was refilled (>= 56 bits) and `e` looked up from it -- the
            // next entry is fetched BEFORE a match is copied, so that the table load overlaps the copy
#define TDTZ_REFILL() do { bits |= load64(in) << bitcnt; in += (63 - bitcnt) >> 3; bitcnt |= 56; } while (0)
            u32 e = 0;
            if (in <= in_fast) {
                TDTZ_REFILL();
                e = ll[bits & ((1u << LL_BITS) - 1u)];
            }
            while (in <= in_fast && out <= out_fast) {
                if (e & E_LIT) {
                    TDTZ_CONSUME((e >> 8) & 15);
                    *out++ = (uint8_t)(e >> 16);
                    e = ll[bits & ((1u << LL_BITS) - 1u)];
                    if (e & E_LIT) {
                        TDTZ_CONSUME((e >> 8) & 15);
                        *out++ = (uint8_t)(e >> 16);
                        e = ll[bits & ((1u << LL_BITS) - 1u)];
                        if (e & E_LIT) {
                            TDTZ_CONSUME((e >> 8) & 15);
                            *out++ = (uint8_t)(e >> 16);
                            TDTZ_REFILL();
                            e = ll[bits & ((1u << LL_BITS) - 1u)];
                            continue;
                        }
                    }
                    TDTZ_REFILL();      // (`e` came from the low bits, which a refill leaves alone)
                }
                if (e & E_EXC) {
                    if (!(e & E_SUB)) {
                        if (e & E_ERR) return false;
                        TDTZ_CONSUME((e >> 8) & 15);
                        done = true;
                        break;
                    }
                    TDTZ_CONSUME(LL_BITS);
                    e = ll[(e >> 16) + (u32)(bits & ((1u << (e & 31)) - 1u))];
                    if (e & E_LIT) {
                        TDTZ_CONSUME((e >> 8) & 15);
                        *out++ = (uint8_t)(e >> 16);
                        TDTZ_REFILL();
                        e = ll[bits & ((1u << LL_BITS) - 1u)];
                        continue;
                    }
                    if (e & E_EXC) {
                        if (e & (E_ERR | E_SUB)) return false;
                        TDTZ_CONSUME((e >> 8) & 15);
                        done = true;
                        break;
                    }
                }
                TDTZ_CONSUME((e >> 8) & 15);
                const u32 length = (e >> 16) + (u32)(bits & ((1u << (e & 31)) - 1u));
                TDTZ_CONSUME(e & 31);
                u32 d = dd[bits & ((1u << D_BITS) - 1u)];
                if (d & E_EXC) {
                    if (!(d & E_SUB)) return false;
                    TDTZ_CONSUME(D_BITS);
                    d = dd[(d >> 16) + (u32)(bits & ((1u << (d & 31)) - 1u))];
                    if (d & E_EXC) return false;
                }
                TDTZ_CONSUME((d >> 8) & 15);
                const u32 dist = (d >> 16) + (u32)(bits & ((1u << (d & 31)) - 1u));
                TDTZ_CONSUME(d & 31);
                if (dist > (size_t)(out - out_start)) return false;
                TDTZ_REFILL();
                e = ll[bits & ((1u << LL_BITS) - 1u)];
                const uint8_t *src = out - dist;
                uint8_t *const end = out + length;
                if (dist >= 8) {
                    store64(out, load64(src));
                    store64(out + 8, load64(src + 8));
                    out += 16;
                    src += 16;
                    while (out < end) {
                        store64(out, load64(src));
                        out += 8;
                        src += 8;
                    }
                } else if (dist == 1) {
                    const u64 v = 0x0101010101010101ull * src[0];
                    do {
                        store64(out, v);
                        out += 8;
                    } while (out < end);
                } else {
                    do {
                        *out++ = *src++;
                    } while (out < end);
                }
                out = end;
            }
#undef TDTZ_REFILL
        }
        // ---- checked loop: the tail of the block (or all of a tiny one) ---------------------------------------------
        while (!done) {
            refill_slow();
            u32 e = ll[bits & ((1u << LL_BITS) - 1u)];
            if ((e & E_EXC) && (e & E_SUB)) {
                TDTZ_CONSUME(LL_BITS);
                e = ll[(e >> 16) + (u32)(bits & ((1u << (e & 31)) - 1u))];
                if ((e & E_EXC) && (e & E_SUB)) return false;
            }
            if (e & E_LIT) {
                if (out >= out_end) return false;
                TDTZ_CONSUME((e >> 8) & 15);
                *out++ = (uint8_t)(e >> 16);
                continue;
            }
            if (e & E_EXC) {
                if (e & E_ERR) return false;
                TDTZ_CONSUME((e >> 8) & 15);
                break;
            }
            TDTZ_CONSUME((e >> 8) & 15);
            const u32 length = (e >> 16) + (u32)(bits & ((1u << (e & 31)) - 1u));
            TDTZ_CONSUME(e & 31);
            u32 d = dd[bits & ((1u << D_BITS) - 1u)];
            if (d & E_EXC) {
                if (!(d & E_SUB)) return false;
                TDTZ_CONSUME(D_BITS);
                d = dd[(d >> 16) + (u32)(bits & ((1u << (d & 31)) - 1u))];
                if (d & E_EXC) return false;
            }
            TDTZ_CONSUME((d >> 8) & 15);
            const u32 dist = (d >> 16) + (u32)(bits & ((1u << (d & 31)) - 1u));
            TDTZ_CONSUME(d & 31);
            if (dist > (size_t)(out - out_start) || length > (size_t)(out_end - out)) return false;
            const uint8_t *src = out - dist;
            for (u32 i = 0; i < length; i++) out[i] = src[i];
            out += length;
        }
        if (pad > ((size_t)bitcnt >> 3)) return false;   // symbols were decoded from bits that are not in the input
    }
#undef TDTZ_CONSUME
    return out == out_end;
}

}  // namespace tdtz

// ---- CRC-32 of an inflated block (the BGZF trailer's check) ---------------------------------------------------------------
// zlib's crc32 runs at ~2 GB/s, a sixth of a block's decode time; with carry-less multiplication (x86-64 with PCLMULQDQ,
// found at run time) it is ~15 GB/s.  Other hosts and short buffers use zlib's.
#if defined(__x86_64__)
#include <immintrin.h>
namespace tdtz {
// CRC-32 (IEEE, reflected) of a multiple of 16 bytes (>= 64) by carry-less multiplication: four 128-bit lanes folded
// across 64 bytes per step, then to one lane, to 64 bits and by Barrett reduction to 32 (constants: x^n mod P for the
// fold distances, as in Gopal et al., "Fast CRC computation for generic polynomials using PCLMULQDQ").
// state in / out WITHOUT the final inversion (zlib's value is ~state).
__attribute__((target("pclmul,sse4.1"))) inline uint32_t crc32_clmul_blocks(uint32_t state, const uint8_t *p, size_t len) {
    const __m128i k1k2 = _mm_set_epi64x(0x00000001c6e41596ll, 0x0000000154442bd4ll);   // fold by 64 bytes
    const __m128i k3k4 = _mm_set_epi64x(0x00000000ccaa009ell, 0x00000001751997d0ll);   // fold by 16 bytes
    const __m128i k5 = _mm_set_epi64x(0, 0x0000000163cd6124ll);
    const __m128i poly = _mm_set_epi64x(0x00000001f7011641ll, 0x00000001db710641ll);   // (mu, P)
    const __m128i mask32 = _mm_set_epi32(0, 0, 0, -1);
    __m128i x1 = _mm_loadu_si128((const __m128i *)(p + 0)), x2 = _mm_loadu_si128((const __m128i *)(p + 16));
    __m128i x3 = _mm_loadu_si128((const __m128i *)(p + 32)), x4 = _mm_loadu_si128((const __m128i *)(p + 48));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)state));
    p += 64;
    len -= 64;
    while (len >= 64) {
        __m128i t1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), t2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        __m128i t3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), t4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11);
        x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
        x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11);
        x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, t1), _mm_loadu_si128((const __m128i *)(p + 0)));
        x2 = _mm_xor_si128(_mm_xor_si128(x2, t2), _mm_loadu_si128((const __m128i *)(p + 16)));
        x3 = _mm_xor_si128(_mm_xor_si128(x3, t3), _mm_loadu_si128((const __m128i *)(p + 32)));
        x4 = _mm_xor_si128(_mm_xor_si128(x4, t4), _mm_loadu_si128((const __m128i *)(p + 48)));
        p += 64;
        len -= 64;
    }
    // four lanes -> one
    __m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, t), x2);
    t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, t), x3);
    t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, t), x4);
    while (len >= 16) {
        t = _mm_clmulepi64_si128(x1, k3k4, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, t), _mm_loadu_si128((const __m128i *)p));
        p += 16;
        len -= 16;
    }
    // 128 -> 64 bits
    __m128i x2b = _mm_clmulepi64_si128(x1, k3k4, 0x10);     // low 64 of x1 times k4
    x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), x2b);
    // 64 -> 32 + 32
    __m128i x0 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, mask32);
    x1 = _mm_clmulepi64_si128(x1, k5, 0x00);
    x1 = _mm_xor_si128(x1, x0);
    // Barrett reduction
    __m128i y = _mm_and_si128(x1, mask32);
    y = _mm_clmulepi64_si128(y, poly, 0x10);
    y = _mm_and_si128(y, mask32);
    y = _mm_clmulepi64_si128(y, poly, 0x00);
    x1 = _mm_xor_si128(x1, y);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
inline bool have_clmul() {
    static const bool ok = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    return ok;
}
}  // namespace tdtz
#endif
namespace tdtz {
// zlib's crc32(0, p, len)
inline uint32_t crc32_of(const uint8_t *p, size_t len) {
#if defined(__x86_64__)
    if (len >= 64 && have_clmul()) {
        const size_t body = len & ~(size_t)15;
        const uint32_t crc = ~crc32_clmul_blocks(0xffffffffu, p, body);
        return body == len ? crc : (uint32_t)::crc32(crc, p + body, (uInt)(len - body));
    }
#endif
    return (uint32_t)::crc32(::crc32(0L, Z_NULL, 0), p, (uInt)len);
}
}
