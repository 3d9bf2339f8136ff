// lz4b200_thread_codec.cuh — K1-T / K2-T: one LZ4 block per THREAD.
//
// The warp-cooperative kernels (lz4b200_kernels.cuh, lz4b200_enc_split.cuh) spend a whole warp's issue slot and
// ~150 L1 wavefronts on every sequence of one block.  A batch of 64 KiB blocks has no parallelism inside a block
// (compress_internal, reference src/block/compress.rs:318-489, and decompress_internal, src/block/decompress.rs:
// 201-449, are serial chains) but thousands of blocks, so here every lane runs the reference's sequential loop
// for its OWN block:
//   * all data moves as naturally aligned 64-bit words: a 16-byte register window slides over each byte stream
//     (RoStream), and output is write-combined in a 64-bit accumulator that is stored when a word is complete
//     (Appender) — ~15 L1 wavefronts per sequence instead of ~150, no shuffles, no ballots;
//   * the loops are written as small state machines (one probe OR one 8-byte extension step per iteration for
//     the encoder; one 32-byte copy chunk per iteration for the decoder) so the 32 independent chains of a warp
//     do not wait for each other's loop trip counts;
//   * the encoder's 4096-entry table lives in global memory (L1/L2-served, 8 KiB per thread).
// The code is plain per-thread C++: the same functions compile for the host (tests/cpp/thread_codec_host.cpp),
// where they are checked byte-for-byte against the oracle without a GPU.  The product only ever calls them from
// the kernels in lz4b200_thread_kernels.cuh.
#pragma once
#include <stdint.h>

#include "../../include/lz4b200.h"

#if defined(__CUDACC__)
#define TC_FN __device__ __forceinline__
#define TC_MFN __device__ __forceinline__
#define TC_LD64_RO(p) __ldg(p)
#define TC_LD8_RO(p) __ldg(p)
#define TC_CTZ64(x) (__ffsll((long long)(x)) - 1)
#else
#define TC_FN static inline
#define TC_MFN inline
#define TC_LD64_RO(p) (*(p))
#define TC_LD8_RO(p) (*(p))
#define TC_CTZ64(x) __builtin_ctzll(x)
#endif

namespace lz4b200 {
namespace tc {

// hashtable.rs:19-21 (then >> 4) and hashtable.rs:27-34 (then >> 4): 12-bit slot of the 8 bytes at a position
TC_FN uint32_t slot4(uint64_t v) { return ((uint32_t)v * 2654435761u) >> 20; }
TC_FN uint32_t slot5(uint64_t v) { return (uint32_t)(((v << 24) * 889523592379ull) >> 52); }

// ---------------------------------------------------------------------------------------------
// Read-only byte stream seen through a 16-byte register window of two aligned 64-bit words.
// Positions are stream-relative; x = position + misalignment of the base pointer.  A word is fetched
// only if it holds at least one byte of the stream, so nothing outside the words that overlap
// [base, base + n) is ever touched; bytes past n come back as whatever shares the last word, then zeros.
// ---------------------------------------------------------------------------------------------
template <bool kReadOnly>
struct Stream {
    const uint64_t *w;
    uint32_t mis, xend;
    uint64_t lo, hi;
    uint32_t wb;                               // x of lo's first byte; 0xffffffff: window empty

    TC_MFN void init(const uint8_t *p, uint32_t n)
    {
        mis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 7u);
        w = reinterpret_cast<const uint64_t *>(p - mis);
        xend = mis + n;
        lo = hi = 0; wb = 0xffffffffu;
    }
    TC_MFN uint64_t word(uint32_t a) const      // a: multiple of 8 in x space
    {
        if (a >= xend) return 0;
        if (kReadOnly) return TC_LD64_RO(w + (a >> 3));
        return w[a >> 3];
    }
    // the 8 bytes at stream position pos (little endian)
    TC_MFN uint64_t rd8(uint32_t pos)
    {
        const uint32_t x = pos + mis, a = x & ~7u;
        if (a != wb) {
            lo = (a == wb + 8u) ? hi : word(a);
            hi = word(a + 8u);
            wb = a;
        }
        const uint32_t sh = (x & 7u) * 8u;
        return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
    }
    TC_MFN uint32_t byte(uint32_t pos) const
    {
        return TC_LD8_RO(reinterpret_cast<const uint8_t *>(w) + mis + pos);
    }
};

// ---------------------------------------------------------------------------------------------
// Sequential byte output, write-combined into aligned 64-bit stores.  The first and the last word of the
// stream are written with byte stores so nothing outside [dst, dst + produced) is modified.
// ---------------------------------------------------------------------------------------------
struct Appender {
    uint64_t *w;
    uint32_t x, x0;                            // next byte / first byte, in x space (x0 = misalignment of dst)
    uint64_t acc;                              // bytes [x & ~7, x) of the current word

    TC_MFN void init(uint8_t *dst)
    {
        x0 = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 7u);
        w = reinterpret_cast<uint64_t *>(dst - x0);
        x = x0; acc = 0;
    }
    TC_MFN uint32_t produced() const { return x - x0; }
    TC_MFN uint8_t *bytes() const { return reinterpret_cast<uint8_t *>(w); }
    TC_MFN void store_word(uint32_t a, uint64_t v)          // the completed word at x position a
    {
        if (a < x0) {                                      // head word of an unaligned buffer
            uint8_t *b = bytes();
            for (uint32_t i = x0; i < 8u; i++) b[i] = (uint8_t)(v >> (8u * i));
        } else {
            w[a >> 3] = v;
        }
    }
    // append the k (1..8) low bytes of v
    TC_MFN void put(uint64_t v, uint32_t k)
    {
        const uint32_t f = x & 7u;
        if (k < 8u) v &= (1ull << (8u * k)) - 1ull;
        acc |= v << (8u * f);
        const uint32_t a = x - f;
        x += k;
        if (f + k >= 8u) {
            store_word(a, acc);
            acc = f ? v >> (8u * (8u - f)) : 0ull;
        }
    }
    // bytes of the unfinished word -> memory (byte stores).  The accumulator keeps them, so appending may continue.
    TC_MFN void sync()
    {
        const uint32_t f = x & 7u, a = x - f;
        uint8_t *b = bytes() + a;
        for (uint32_t i = (a < x0 ? x0 : 0u); i < f; i++) b[i] = (uint8_t)(acc >> (8u * i));
    }
    // after bytes were written straight to memory up to x position nx: pick the partial word up again
    TC_MFN void resume_at(uint32_t nx)
    {
        x = nx;
        const uint32_t f = x & 7u;
        acc = f ? (w[(x - f) >> 3] & ((1ull << (8u * f)) - 1ull)) : 0ull;
    }
};

// =============================================================================================
// The parse of compress_internal (compress.rs:318-489) for one block, run by ONE thread, as a template over
//   InA/B: byte-stream views of the input (rd8(pos), byte(pos)): cursor side and candidate side
//   Sink : receives every finished sequence as (anchor, match start, offset, match end) and the final literal run
//   tab  : 4096 entries owned by this thread, already filled with 0 (FRESH: position 0 is a legal candidate,
//          compress.rs:353-359) or with TabT's all-ones value (CONT: entries of earlier blocks never match,
//          compress.rs:403-429).  Entries hold block-relative positions.
// One loop, two states: a probe (compress.rs:373-439) or 8 bytes of forward extension (:156-216) per iteration,
// so lanes of a warp that are in different phases of their blocks do not wait for each other's loop trip counts.
// =============================================================================================
template <typename TabT, typename InA, typename InB, typename Sink>
TC_FN void parse_block_thread(InA &in, InB &cs, uint32_t n, TabT *tab, bool cont, bool h5, Sink &sink)
{
    constexpr uint32_t kInvalid = (uint32_t)(TabT)~(TabT)0;
    uint32_t anchor = 0;
    if (n >= 13u) {                            // compress.rs:343-346: shorter inputs are one literal run
        const uint32_t last_probe = n - 12u, lim = n - 6u;
        uint32_t cur = 0, nm = 32u;
        if (!cont) {                           // compress.rs:353-359
            const uint64_t v = in.rd8(0);
            tab[h5 ? slot5(v) : slot4(v)] = 0;
            cur = 1;
        }
        bool ext = false;
        uint32_t mpos = 0, dist = 0, ecur = 0;
        for (;;) {
            if (!ext) {                        // ---- one probe: compress.rs:373-439
                if (cur > last_probe) break;
                const uint64_t v = in.rd8(cur);
                const uint32_t s = h5 ? slot5(v) : slot4(v);
                uint32_t c = tab[s];
                tab[s] = (TabT)cur;
                const uint32_t step = nm >> 5;
                nm++;
                bool hit = false;
                if (c != kInvalid && cur - c <= 65535u) hit = (uint32_t)cs.rd8(c) == (uint32_t)v;
                if (!hit) { cur += step; continue; }
                mpos = cur;                    // compress.rs:272-287
                while (c > 0u && mpos > anchor && in.byte(mpos - 1u) == in.byte(c - 1u)) { mpos--; c--; }
                dist = mpos - c;
                ecur = mpos + 4u;
                ext = true;
            }
            {                                  // ---- 8 bytes of forward extension: compress.rs:156-216
                const uint32_t room = lim - ecur;
                const uint64_t x = in.rd8(ecur) ^ cs.rd8(ecur - dist);
                uint32_t k = x ? (uint32_t)TC_CTZ64(x) >> 3 : 8u;
                if (k > room) k = room;
                ecur += k;
                if (k == 8u && ecur < lim) continue;
            }
            // ---- the sequence is complete: re-insert (compress.rs:460-461) and hand over (:463-486)
            const uint32_t end = ecur;
            {
                const uint64_t v2 = in.rd8(end - 2u);
                tab[h5 ? slot5(v2) : slot4(v2)] = (TabT)(end - 2u);
            }
            sink.sequence(anchor, mpos, dist, end);
            anchor = cur = end;
            nm = 32u;
            ext = false;
        }
    }
    sink.tail(anchor, n);                      // handle_last_literals: compress.rs:237-247
}

// Sink that writes the LZ4 byte stream directly (K1-T): token / length bytes / literals / offset (compress.rs:463-486)
// through a write-combining Appender; literals are read through the candidate-side stream, which is idle then.
template <typename In>
struct DirectSink {
    Appender out;
    In *lits;
    TC_MFN void put_len(uint32_t r)                        // write_integer: compress.rs:224-233
    {
        for (; r >= 255u; r -= 255u) out.put(0xffu, 1u);
        out.put(r, 1u);
    }
    TC_MFN void sequence(uint32_t anchor, uint32_t mpos, uint32_t dist, uint32_t end)
    {
        const uint32_t lit = mpos - anchor, extra = end - mpos - 4u;
        const uint32_t tok = ((lit < 15u ? lit : 15u) << 4) | (extra < 15u ? extra : 15u);
        if (lit == 0u && extra < 15u + 255u) {             // token + offset (+ one length byte) in one go
            uint64_t v = tok | ((uint64_t)dist << 8);
            uint32_t k = 3u;
            if (extra >= 15u) { v |= (uint64_t)(extra - 15u) << 24; k = 4u; }
            out.put(v, k);
            return;
        }
        out.put(tok, 1u);
        if (lit >= 15u) put_len(lit - 15u);
        for (uint32_t i = 0; i < lit; i += 8u) out.put(lits->rd8(anchor + i), lit - i < 8u ? lit - i : 8u);
        out.put(dist, 2u);
        if (extra >= 15u) put_len(extra - 15u);
    }
    TC_MFN void tail(uint32_t anchor, uint32_t n)
    {
        const uint32_t lit = n - anchor;
        out.put((lit < 15u ? lit : 15u) << 4, 1u);
        if (lit >= 15u) put_len(lit - 15u);
        for (uint32_t i = 0; i < lit; i += 8u) out.put(lits->rd8(anchor + i), lit - i < 8u ? lit - i : 8u);
        out.sync();
    }
};

// K1-T: one block, one thread, straight to bytes.  Returns the compressed size.  The caller has checked
// cap >= get_maximum_output_size(n) (compress.rs:338-340).
template <typename TabT>
TC_FN uint32_t encode_block_thread(const uint8_t *src, uint32_t n, uint8_t *dst, TabT *tab, bool cont, bool h5)
{
    Stream<true> in, cs;                       // cursor side / candidate side (cs also serves the literal copies)
    in.init(src, n); cs.init(src, n);
    DirectSink<Stream<true>> sink;
    sink.out.init(dst);
    sink.lits = &cs;
    parse_block_thread<TabT>(in, cs, n, tab, cont, h5, sink);
    return sink.out.produced();
}

// =============================================================================================
// K2-T: decompress_internal (decompress.rs:201-449) for one block, one thread.  Same bytes, same first error and
// the same OutputTooSmall{expected, actual} as the checked path (decompress.rs:330-444).
// =============================================================================================
struct ThreadDecResult {
    uint32_t written;
    int32_t status;
    uint64_t expected;
};

TC_FN ThreadDecResult decode_block_thread(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t cap)
{
    ThreadDecResult r{0u, LZ4B200_OK, 0ull};
    if (n == 0u) { r.status = LZ4B200_DEC_EXPECTED_ANOTHER_BYTE; return r; }    // decompress.rs:207-209
    Stream<true> in;
    in.init(src, n);
    Appender out;
    out.init(dst);
    const uint32_t omis = out.x0;
    const uint32_t xcap = omis + cap;                      // end of the output buffer in x space
    uint32_t ip = 0;
    // pending copy: rem bytes from word array cw at x position cx (cw = the input or the output buffer)
    const uint64_t *cw = nullptr;
    uint32_t cx = 0, cxend = 0, rem = 0;
    uint32_t tok = 0;
    int stage = 0;                                         // 0: at a token, 1: literals done, 2: match done

    for (;;) {
        if (rem == 0u) {
            if (stage == 2) {                              // ---- after a match: the stream may not end here
                if (ip >= n) { r.status = LZ4B200_DEC_EXPECTED_ANOTHER_BYTE; return r; }                       // :439-443
                stage = 0;
            }
            if (stage == 0) {                              // ---- token + literal length
                tok = (uint32_t)in.rd8(ip) & 0xffu;
                ip++;
                uint64_t lit = tok >> 4;
                if (lit == 15u) {                          // read_integer_ptr: decompress.rs:126-157
                    for (;;) {
                        if (ip >= n) { r.status = LZ4B200_DEC_EXPECTED_ANOTHER_BYTE; return r; }
                        const uint32_t b = (uint32_t)in.rd8(ip) & 0xffu;
                        ip++;
                        lit += b;
                        if (b != 255u) break;
                    }
                }
                stage = 1;
                if (lit) {
                    const uint32_t op = out.produced();
                    if (lit > (uint64_t)(n - ip)) { r.status = LZ4B200_DEC_LITERAL_OUT_OF_BOUNDS; return r; }   // :346
                    if (lit > (uint64_t)(cap - op)) {                                                           // :349-354
                        r.status = LZ4B200_DEC_OUTPUT_TOO_SMALL; r.expected = (uint64_t)op + lit; return r;
                    }
                    if (lit <= 8u) {
                        out.put(in.rd8(ip), (uint32_t)lit);
                        ip += (uint32_t)lit;
                    } else {
                        cw = in.w; cx = ip + in.mis; cxend = in.xend; rem = (uint32_t)lit;
                        ip += (uint32_t)lit;
                    }
                }
            }
            if (rem == 0u) {                               // ---- (stage 1) offset + match length
                if (ip >= n) break;                        // the stream ends after literals: decompress.rs:366
                if (n - ip < 2u) { r.status = LZ4B200_DEC_EXPECTED_ANOTHER_BYTE; return r; }                   // :373
                const uint32_t dist = (uint32_t)in.rd8(ip) & 0xffffu;
                ip += 2u;
                if (dist == 0u) { r.status = LZ4B200_DEC_OFFSET_ZERO; return r; }                              // :161-173
                uint64_t mlen = 4u + (tok & 15u);
                if (mlen == 19u) {
                    for (;;) {
                        if (ip >= n) { r.status = LZ4B200_DEC_EXPECTED_ANOTHER_BYTE; return r; }
                        const uint32_t b = (uint32_t)in.rd8(ip) & 0xffu;
                        ip++;
                        mlen += b;
                        if (b != 255u) break;
                    }
                }
                const uint32_t op = out.produced();
                if (dist > op) { r.status = LZ4B200_DEC_OFFSET_OUT_OF_BOUNDS; return r; }                      // :399
                if (mlen > (uint64_t)(cap - op)) {                                                             // :402-406
                    r.status = LZ4B200_DEC_OUTPUT_TOO_SMALL; r.expected = (uint64_t)op + mlen; return r;
                }
                stage = 2;
                if (dist >= 40u) {
                    // every 32-byte chunk's source ends below the word that is still being accumulated
                    cw = out.w; cx = out.x - dist; cxend = xcap; rem = (uint32_t)mlen;
                } else {
                    // close or overlapping source (duplicate_overlapping, decompress.rs:57-82): byte-serial through
                    // memory; power-of-two periods (run fills) continue word-wise once one whole word is periodic
                    out.sync();
                    uint8_t *b = out.bytes();
                    uint32_t x = out.x, left = (uint32_t)mlen;
                    const bool pow2 = (dist & (dist - 1u)) == 0u && dist <= 8u;
                    uint32_t head = left;
                    if (pow2 && left >= 32u) head = ((8u - (x & 7u)) & 7u) + 8u;
                    for (uint32_t i = 0; i < head; i++) b[x + i] = b[x + i - dist];
                    x += head; left -= head;
                    if (left) {
                        const uint64_t pat = out.w[(x - 8u) >> 3];
                        for (; left >= 8u; left -= 8u, x += 8u) out.w[x >> 3] = pat;
                        for (uint32_t i = 0; i < left; i++) b[x + i] = (uint8_t)(pat >> (8u * i));
                        x += left;
                    }
                    out.resume_at(x);
                }
            }
        }
        if (rem) {                                         // ---- one chunk (<= 32 bytes) of the pending copy
            const uint32_t len = rem < 32u ? rem : 32u;
            const uint32_t a = cx & ~7u, sh = (cx & 7u) * 8u;
            const uint32_t need = (cx & 7u) + len;         // bytes of the word run [a, ...) that are used
            const uint64_t *p = cw + (a >> 3);
            uint64_t w0 = p[0], w1 = 0, w2 = 0, w3 = 0, w4 = 0;
            if (need > 8u && a + 8u < cxend) w1 = p[1];
            if (need > 16u && a + 16u < cxend) w2 = p[2];
            if (need > 24u && a + 24u < cxend) w3 = p[3];
            if (need > 32u && a + 32u < cxend) w4 = p[4];
            if (sh) {
                w0 = (w0 >> sh) | (w1 << (64u - sh));
                w1 = (w1 >> sh) | (w2 << (64u - sh));
                w2 = (w2 >> sh) | (w3 << (64u - sh));
                w3 = (w3 >> sh) | (w4 << (64u - sh));
            }
            out.put(w0, len < 8u ? len : 8u);
            if (len > 8u) out.put(w1, len < 16u ? len - 8u : 8u);
            if (len > 16u) out.put(w2, len < 24u ? len - 16u : 8u);
            if (len > 24u) out.put(w3, len - 24u);
            cx += len; rem -= len;
        }
    }
    out.sync();
    r.written = out.produced();
    return r;
}

}  // namespace tc
}  // namespace lz4b200
