// lz4b200_kernels.cuh — sm_100a device code of the LZ4 block codec.
//
// K2  lz4_decompress_blocks : decompress_internal  (reference src/block/decompress.rs:201-449)
// K1  lz4_compress_blocks   : compress_internal    (reference src/block/compress.rs:318-489)
//
// Both kernels are persistent: a fixed grid of warps pulls block indices from a global ticket
// counter, one LZ4 block per warp at a time.  This is HBM/L2-bound byte shuffling — no tensor
// cores.  See DESIGN.md for the layout, the per-kernel roofline and what each phase costs.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lz4b200.h"

namespace lz4b200 {

constexpr uint32_t kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// descriptors (device pointers)
// ---------------------------------------------------------------------------------------------
struct BatchArgs {
    const uint8_t *in;
    const uint64_t *in_off;
    const uint32_t *in_len;
    const uint8_t *flags;          // compress only; may be null
    uint8_t *out;
    const uint64_t *out_off;
    const uint32_t *out_cap;
    uint32_t *out_len;
    int32_t *status;
    uint64_t *err_expected;        // decompress only; may be null
    uint32_t nblocks;
    uint32_t *tickets;             // [0] next block, [1] warps finished (self-resetting)
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Pulls the next block index for this warp; the last warp to drain the queue re-arms the
// counters so the next launch needs no memset.
__device__ __forceinline__ uint32_t next_ticket(uint32_t *tickets)
{
    uint32_t t = 0;
    if (lane_id() == 0) t = atomicAdd(&tickets[0], 1u);
    return __shfl_sync(kFull, t, 0);
}
__device__ __forceinline__ void retire_warp(uint32_t *tickets, uint32_t total_warps)
{
    if (lane_id() == 0) {
        __threadfence();
        if (atomicAdd(&tickets[1], 1u) == total_warps - 1) {
            tickets[0] = 0;
            tickets[1] = 0;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// unaligned little-endian fetches built from aligned 32-bit loads.
// `w` is the block's base pointer rounded down to 4 bytes, `x` = position + misalignment.
// A fetch of k bytes at x touches only words that contain at least one of those k bytes
// (callers guarantee the k bytes are inside the block), so nothing outside the allocation's
// last touched word is ever read.
// ---------------------------------------------------------------------------------------------
struct WordView {
    const uint32_t *w;
    uint32_t mis;
    __device__ __forceinline__ explicit WordView(const uint8_t *p)
        : w(reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~uintptr_t(3))),
          mis(static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p) & 3u)) {}
    // 4 bytes at pos; requires pos+4 <= n and at least one more byte (pos+5 <= n) OR alignment luck:
    // callers use it only where pos + 8 <= n.
    __device__ __forceinline__ uint32_t ro4(uint32_t pos) const
    {
        uint32_t x = pos + mis;
        uint32_t a = __ldg(w + (x >> 2)), b = __ldg(w + (x >> 2) + 1);
        return __funnelshift_r(a, b, (x & 3u) * 8u);
    }
    // low 5 bytes at pos as (lo32, hi8); requires pos + 8 <= n.
    __device__ __forceinline__ void ro5(uint32_t pos, uint32_t &lo, uint32_t &hi) const
    {
        uint32_t x = pos + mis;
        uint32_t a = __ldg(w + (x >> 2)), b = __ldg(w + (x >> 2) + 1);
        uint32_t sh = (x & 3u) * 8u;
        lo = __funnelshift_r(a, b, sh);
        hi = (b >> sh) & 0xffu;
    }
};

// =============================================================================================
// K2: decode one block with one warp.
// Semantics = the checked path of decompress_internal (decompress.rs:330-444): same bytes,
// same first error, same OutputTooSmall{expected, actual} fields.
// =============================================================================================
struct DecResult {
    uint32_t written;
    int32_t status;
    uint64_t expected;
};

// out[op .. op+len) = src[ip .. ip+len)   (compressed stream -> output; never overlaps)
__device__ __forceinline__ void copy_literals(uint8_t *dst, const uint8_t *__restrict__ src, uint32_t len,
                                              uint32_t lane)
{
    for (uint32_t i = lane; i < len; i += 32) dst[i] = __ldg(src + i);
}

// out[op .. op+len) = out[op-dist .. ) with LZ77 byte-serial semantics (duplicate(),
// duplicate_overlapping(): decompress.rs:11-82; offset 1 = run fill, decompress_safe.rs:311-313).
__device__ __forceinline__ void copy_match(uint8_t *dst, uint32_t dist, uint32_t len, uint32_t lane)
{
    const uint8_t *from = dst - dist;
    if (dist >= len) {                       // source entirely older than this match
        for (uint32_t i = lane; i < len; i += 32) dst[i] = from[i];
    } else if (dist >= 32) {                 // each 32-byte step only needs earlier steps
        for (uint32_t base = 0; base < len; base += 32) {
            uint32_t i = base + lane;
            if (i < len) dst[i] = from[i];
            __syncwarp();
        }
    } else {                                 // period < 32: every byte is a copy of the seed period
        uint32_t r = lane % dist, step = 32u % dist;
        for (uint32_t i = lane; i < len; i += 32) {
            dst[i] = from[r];
            r += step;
            if (r >= dist) r -= dist;
        }
    }
}

__device__ __forceinline__ DecResult decode_block(const uint8_t *__restrict__ src, uint32_t n, uint8_t *dst,
                                               uint32_t cap)
{
    const uint32_t lane = lane_id();
    DecResult r{0u, LZ4B200_OK, 0ull};
    if (n == 0) { r.status = LZ4B200_DEC_EXPECTED_ANOTHER_BYTE; return r; }   // decompress.rs:207-209
    const WordView view(src);
    uint32_t ip = 0, op = 0;

    for (;;) {
        // ---- token -----------------------------------------------------------------------
        // One 4-byte fetch covers token + offset + first length byte of a literal-free sequence.
        const bool wide = ip + 8 <= n;
        const uint32_t v0 = wide ? view.ro4(ip) : (uint32_t)__ldg(src + ip);
        const uint32_t tok = v0 & 0xffu;
        ip++;
        uint32_t lit = tok >> 4;
        if (lit == 15) {                                       // read_integer_ptr: decompress.rs:126-157
            uint64_t acc = 15;
            for (;;) {
                if (ip >= n) { r.status = LZ4B200_DEC_EXPECTED_ANOTHER_BYTE; return r; }
                uint32_t b = __ldg(src + ip++);
                acc += b;
                if (b != 255) break;
            }
            if (acc > (uint64_t)(n - ip)) { r.status = LZ4B200_DEC_LITERAL_OUT_OF_BOUNDS; return r; }
            lit = (uint32_t)acc;
        }
        if (lit) {
            if (lit > n - ip) { r.status = LZ4B200_DEC_LITERAL_OUT_OF_BOUNDS; return r; }        // :346
            if (lit > cap - op) {                                                                  // :349-354
                r.status = LZ4B200_DEC_OUTPUT_TOO_SMALL; r.expected = (uint64_t)op + lit; return r;
            }
            copy_literals(dst + op, src + ip, lit, lane);
            ip += lit; op += lit;
        }
        if (ip >= n) break;                                    // the stream ends after literals: :366
        if (n - ip < 2) { r.status = LZ4B200_DEC_EXPECTED_ANOTHER_BYTE; return r; }               // :373

        // ---- offset + match length -------------------------------------------------------
        uint32_t dist, ext0 = 0x100;                           // ext0: first extension byte if prefetched
        if (wide && tok < 16u) {                               // no literals: all inside v0
            dist = (v0 >> 8) & 0xffffu; ext0 = v0 >> 24;
        } else if (ip + 8 <= n) {
            uint32_t v = view.ro4(ip);
            dist = v & 0xffffu; ext0 = (v >> 16) & 0xffu;
        } else {
            dist = (uint32_t)__ldg(src + ip) | ((uint32_t)__ldg(src + ip + 1) << 8);
        }
        ip += 2;
        if (dist == 0) { r.status = LZ4B200_DEC_OFFSET_ZERO; return r; }                           // :161-173
        uint64_t mlen = 4u + (tok & 15u);
        if (mlen == 19) {
            if (ext0 < 255) { mlen += ext0; ip++; }
            else {
                for (;;) {
                    if (ip >= n) { r.status = LZ4B200_DEC_EXPECTED_ANOTHER_BYTE; return r; }
                    uint32_t b = __ldg(src + ip++);
                    mlen += b;
                    if (b != 255) break;
                }
            }
        }
        if (dist > op) { r.status = LZ4B200_DEC_OFFSET_OUT_OF_BOUNDS; return r; }                  // :399
        if (mlen > (uint64_t)(cap - op)) {                                                         // :402-406
            r.status = LZ4B200_DEC_OUTPUT_TOO_SMALL; r.expected = (uint64_t)op + mlen; return r;
        }
        __syncwarp();                                          // earlier stores of this warp -> visible
        copy_match(dst + op, dist, (uint32_t)mlen, lane);
        op += (uint32_t)mlen;
        if (ip >= n) { r.status = LZ4B200_DEC_EXPECTED_ANOTHER_BYTE; return r; }                   // :439-443
    }
    r.written = op;
    return r;
}

constexpr int kDecWarpsPerCta = 4;

__global__ void __launch_bounds__(kDecWarpsPerCta * 32)
lz4_decompress_blocks(BatchArgs a)
{
    const uint32_t total_warps = gridDim.x * kDecWarpsPerCta;
    for (uint32_t b = next_ticket(a.tickets); b < a.nblocks; b = next_ticket(a.tickets)) {
        DecResult r = decode_block(a.in + a.in_off[b], a.in_len[b], a.out + a.out_off[b], a.out_cap[b]);
        if (lane_id() == 0) {
            a.out_len[b] = r.status == LZ4B200_OK ? r.written : 0u;
            a.status[b] = r.status;
            if (a.err_expected) a.err_expected[b] = r.expected;
        }
    }
    retire_warp(a.tickets, total_warps);
}

// =============================================================================================
// K1: encode one block with one warp — exact emulation of the reference's sequential greedy
// parse.  The next 32 probe positions of the probe loop (compress.rs:373-439) are evaluated by
// the 32 lanes at once; table writes that the sequential loop would have made between two
// probes of the same batch are forwarded with match.any, and only the writes up to the winning
// probe are committed, so the table state after every sequence is identical to the reference's.
// =============================================================================================
template <typename TabT> struct TabTraits;
template <> struct TabTraits<uint16_t> { static constexpr uint32_t kInvalid = 0xffffu; };
template <> struct TabTraits<uint32_t> { static constexpr uint32_t kInvalid = 0xffffffffu; };

__device__ __forceinline__ uint32_t slot_h4(uint32_t v4)                 // hashtable.rs:19-21 then >>4
{
    return (v4 * 2654435761u) >> 20;
}
__device__ __forceinline__ uint32_t slot_h5(uint32_t lo, uint32_t hi8)   // hashtable.rs:27-34 then >>4
{
    // ((v << 24) * 889523592379) >> 52 with v = hi8:lo (40 bits).  Only the top 12 bits of the low
    // 64 product bits are needed: work on the upper 32-bit half.
    const uint32_t p_lo = 0x1BBCDCBBu, p_hi = 0xCFu;          // 889523592379 = 0xCF_1BBCDCBB
    uint32_t a_lo = lo << 24;                                  // (v << 24) low word
    uint32_t a_hi = (lo >> 8) | (hi8 << 24);                   // (v << 24) high word
    uint32_t top = __umulhi(a_lo, p_lo) + a_lo * p_hi + a_hi * p_lo;
    return top >> 20;
}

// length-extension bytes for value v (= len - 15) at dst; returns number of bytes written.
__device__ __forceinline__ uint32_t put_ext(uint8_t *dst, uint32_t v, uint32_t lane)
{
    uint32_t k = v / 255u, rem = v - k * 255u;
    for (uint32_t i = lane; i < k; i += 32) dst[i] = 0xff;
    if (lane == 0) dst[k] = (uint8_t)rem;
    return k + 1;
}

__device__ __forceinline__ uint32_t put_last_literals(uint8_t *dst, const uint8_t *__restrict__ src,
                                                      uint32_t from, uint32_t n, uint32_t lane)
{
    uint32_t len = n - from, o = 1;
    if (lane == 0) dst[0] = (uint8_t)((len < 15 ? len : 15) << 4);
    if (len >= 15) o += put_ext(dst + o, len - 15, lane);
    for (uint32_t i = lane; i < len; i += 32) dst[o + i] = __ldg(src + from + i);
    return o + len;
}

template <typename TabT>
__device__ __forceinline__ uint32_t encode_block(const uint8_t *__restrict__ src, uint32_t n, uint8_t *dst,
                                              TabT *tab, bool cont, bool h5)
{
    constexpr uint32_t kInvalid = TabTraits<TabT>::kInvalid;
    const uint32_t lane = lane_id();
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t o = 0;                                             // output cursor
    if (n < 13) return put_last_literals(dst, src, 0, n, lane);  // compress.rs:343-346

    // table: zero for a fresh table (0 is a legal candidate: position 0), "invalid" when the
    // block continues a frame stream (entries of earlier blocks can never match).
    {
        constexpr uint32_t words = 4096 * sizeof(TabT) / 4;
        uint32_t fill = cont ? 0xffffffffu : 0u;
        uint32_t *t32 = reinterpret_cast<uint32_t *>(tab);
        for (uint32_t i = lane; i < words; i += 32) t32[i] = fill;
        __syncwarp();
    }
    const WordView view(src);
    const uint32_t last_probe = n - 12;
    uint32_t anchor = 0, cur = 0;
    if (!cont) {                                                // compress.rs:353-359
        uint32_t lo, hi; view.ro5(0, lo, hi);
        uint32_t s = h5 ? slot_h5(lo, hi) : slot_h4(lo);
        if (lane == 0) tab[s] = 0;
        cur = 1;
        __syncwarp();
    }

    for (;;) {
        // ---- probe batches ----------------------------------------------------------------
        uint32_t base = cur, stride = 1, cand = 0;
        for (;;) {
            uint32_t p = base + lane * stride;
            bool term = p > last_probe;
            uint32_t key = 0x10000u | lane, v4 = 0, cnd = kInvalid;
            if (!term) {
                uint32_t hi; view.ro5(p, v4, hi);
                key = h5 ? slot_h5(v4, hi) : slot_h4(v4);
                cnd = tab[key];
            }
            uint32_t same = __match_any_sync(kFull, key);
            uint32_t prior = same & lt_mask;
            if (prior) cnd = base + (31u - __clz(prior)) * stride;   // forwarded in-batch write
            bool hit = false;
            if (!term && cnd != kInvalid && p - cnd <= 65535u) hit = (view.ro4(cnd) == v4);
            uint32_t hits = __ballot_sync(kFull, hit), terms = __ballot_sync(kFull, term);
            uint32_t win = hits ? (uint32_t)__ffs(hits) - 1u : 32u;
            uint32_t tfirst = terms ? (uint32_t)__ffs(terms) - 1u : 32u;
            if (tfirst < win)                                       // compress.rs:381-384
                return o + put_last_literals(dst + o, src, anchor, n, lane);
            // commit the table writes of probes 0..win (last writer per slot wins)
            uint32_t upto = win < 32 ? win : 31u;
            uint32_t le_mask = upto == 31 ? kFull : ((2u << upto) - 1u);
            uint32_t mine = same & le_mask;
            if (lane <= upto && (31u - __clz(mine)) == lane) tab[key] = (TabT)p;
            __syncwarp();
            if (win < 32) {
                cur = __shfl_sync(kFull, p, win);
                cand = __shfl_sync(kFull, cnd, win);
                break;
            }
            base += 32u * stride;
            stride++;
        }
        const uint32_t dist = cur - cand;

        // ---- extend backwards (compress.rs:272-287) ---------------------------------------
        for (;;) {
            uint32_t room = min(cand, cur - anchor);                // how far both may step back
            bool ok = lane < room && __ldg(src + cur - 1 - lane) == __ldg(src + cand - 1 - lane);
            uint32_t bad = ~__ballot_sync(kFull, ok);
            uint32_t k = bad ? (uint32_t)__ffs(bad) - 1u : 32u;
            cur -= k; cand -= k;
            if (k < 32) break;
        }
        const uint32_t lit = cur - anchor;

        // ---- extend forwards (compress.rs:156-216), limit n - 6 ---------------------------
        cur += 4; cand += 4;
        uint32_t extra = 0;
        {
            const uint32_t lim = n - 6;
            for (;;) {
                uint32_t q = cur + lane;
                bool ok = q < lim && __ldg(src + q) == __ldg(src + cand + lane);
                uint32_t bad = ~__ballot_sync(kFull, ok);
                uint32_t k = bad ? (uint32_t)__ffs(bad) - 1u : 32u;
                extra += k; cur += k; cand += k;
                if (k < 32) break;
            }
        }
        // ---- T[H(cur-2)] = cur-2 (compress.rs:460-461) ------------------------------------
        {
            uint32_t lo, hi; view.ro5(cur - 2, lo, hi);
            uint32_t s = h5 ? slot_h5(lo, hi) : slot_h4(lo);
            if (lane == 0) tab[s] = (TabT)(cur - 2);
            __syncwarp();
        }
        // ---- emit the sequence (compress.rs:463-486) --------------------------------------
        if (lane == 0) dst[o] = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (extra < 15 ? extra : 15));
        o++;
        if (lit >= 15) o += put_ext(dst + o, lit - 15, lane);
        for (uint32_t i = lane; i < lit; i += 32) dst[o + i] = __ldg(src + anchor + i);
        o += lit;
        if (lane == 0) { dst[o] = (uint8_t)dist; dst[o + 1] = (uint8_t)(dist >> 8); }
        o += 2;
        if (extra >= 15) o += put_ext(dst + o, extra - 15, lane);
        anchor = cur;
    }
}

__device__ __forceinline__ uint64_t max_output_size_dev(uint32_t n)
{
    return 20ull + ((uint64_t)n * 110ull) / 100ull;
}

// One CTA = kWarps warps, each with a private 4096-slot table in shared memory.
// Blocks of up to 65 536 bytes use the TabT=uint16_t instantiation (8 KiB per warp), larger
// ones uint32_t (16 KiB per warp); the host launches both over the same ticket space and each
// instantiation skips the blocks that belong to the other.
template <typename TabT, int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
lz4_compress_blocks(BatchArgs a, uint32_t *tickets)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    TabT *tab = reinterpret_cast<TabT *>(smem_raw) + (threadIdx.x >> 5) * 4096;
    const uint32_t total_warps = gridDim.x * kWarps;
    constexpr bool kSmall = sizeof(TabT) == 2;
    for (uint32_t b = next_ticket(tickets); b < a.nblocks; b = next_ticket(tickets)) {
        const uint32_t n = a.in_len[b];
        if ((n <= 65536u) != kSmall) continue;
        const uint32_t fl = a.flags ? a.flags[b] : 0u;
        uint32_t written = 0; int32_t st = LZ4B200_OK;
        if ((uint64_t)a.out_cap[b] < max_output_size_dev(n)) {              // compress.rs:338-340
            st = LZ4B200_COMPRESS_OUTPUT_TOO_SMALL;
        } else {
            const bool h5 = (fl & LZ4B200_BLOCK_HASH5_ALWAYS) || n >= 65535u;   // compress.rs:559
            written = encode_block<TabT>(a.in + a.in_off[b], n, a.out + a.out_off[b], tab,
                                         (fl & LZ4B200_BLOCK_CONT) != 0, h5);
        }
        if (lane_id() == 0) { a.out_len[b] = written; a.status[b] = st; }
    }
    retire_warp(tickets, total_warps);
}

}  // namespace lz4b200
