// lz4b200_solo_ring.cuh — the shared-memory input ring of the K1-S kernel (lz4b200_solo_kernel.cuh) and the byte-stream
// views the parse reads it through.  Host-compilable: tests/cpp/thread_codec_host.cpp runs the same index arithmetic
// (slot reuse, look-ahead, history bound, backward-extension fallback) over a host ring whose "TMA" is a memcpy, so a slot
// that is overwritten while the parse still needs it shows up as a byte difference against the oracle on the CPU.
#pragma once
#include "lz4b200_thread_codec.cuh"

#if defined(__CUDACC__)
#include "lz4b200_kernels.cuh"
#else
#include <string.h>
#endif

namespace lz4b200 {

#if defined(__CUDACC__)
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// fill: `body` bytes (a multiple of 16) by TMA, completion on `bar`
__device__ __forceinline__ void ring_fill(uint8_t *dst, const uint8_t *src, uint32_t body, uint64_t *bar)
{
    mbar_expect_tx(bar, body);                             // one arrival + `body` bytes complete the phase
    if (body) tma_load_1d(dst, src, body, bar);
}
__device__ __forceinline__ void ring_wait(uint64_t *bar, uint32_t parity) { mbar_wait(bar, parity); }
#else
static inline void ring_fill(uint8_t *dst, const uint8_t *src, uint32_t body, uint64_t *) { memcpy(dst, src, body); }
static inline void ring_wait(uint64_t *, uint32_t) {}
#endif

constexpr uint32_t kSoloChunk = 2048;                      // bytes per TMA bulk copy / ring slot
constexpr uint32_t kSoloSlots = 40;
constexpr uint32_t kSoloRing = kSoloChunk * kSoloSlots;    // 80 KiB
// With a the (8-byte aligned) base of the furthest word pair the cursor side has read, the parse can still touch
// [a - 65535 - 7, a + 16): candidates lie at most MAX_DISTANCE behind the cursor and are read as aligned words.  That is
// chunks chunk(a) - (65536/CH + 1) .. chunk(a) + 1, i.e. 65536/CH + 3 slots; the rest of the ring runs ahead.
constexpr uint32_t kSoloAhead = kSoloSlots - (65536u / kSoloChunk + 2u);
static_assert(kSoloSlots <= 64 && kSoloAhead >= 2, "ring too small");

// The input ring of one chain.  Owned by the matcher thread (all members are its registers).
// x space: x = block position + (src & 15); chunk k holds x in [k*CH, (k+1)*CH) in slot k % kSoloSlots.
struct SoloFeed {
    uint8_t *ring;
    uint64_t *bars;            // kSoloSlots mbarriers, one phase per fill
    const uint8_t *src_al;     // src rounded down to 16 bytes
    uint32_t mis, xend, nchunks;
    uint32_t issued;           // chunks [0, issued) requested
    uint32_t waited;           // chunks [0, waited) landed
    uint64_t phases;           // bit s: parity the next wait on slot s uses

    TC_MFN void begin(const uint8_t *src, uint32_t n)
    {
        mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        src_al = src - mis;
        xend = mis + n;
        nchunks = (xend + kSoloChunk - 1u) / kSoloChunk;
        issued = waited = 0;
    }
    TC_MFN void issue_upto(uint32_t c_hi)       // request chunks <= c_hi (clamped to the block)
    {
        const uint32_t stop = c_hi + 1u < nchunks ? c_hi + 1u : nchunks;
        for (; issued < stop; issued++) {
            const uint32_t s = issued % kSoloSlots, x0 = issued * kSoloChunk;
            const uint32_t valid = xend - x0 < kSoloChunk ? xend - x0 : kSoloChunk;
            const uint32_t body = valid & ~15u;            // the 16-byte granules that lie wholly inside the block
            uint8_t *dst = ring + s * kSoloChunk;
            for (uint32_t i = body; i < valid; i++) dst[i] = TC_LD8_RO(src_al + x0 + i);   // the block's last <16 bytes
            ring_fill(dst, src_al + x0, body, bars + s);
        }
    }
    TC_MFN void wait_upto(uint32_t c)            // chunks <= c have landed (c < issued)
    {
        for (; waited <= c; waited++) {
            const uint32_t s = waited % kSoloSlots;
            ring_wait(bars + s, (uint32_t)(phases >> s) & 1u);
            phases ^= 1ull << s;
        }
    }
    // the 16 bytes at x position a (8-byte aligned) are about to be read, and a is the furthest base so far
    TC_MFN void need(uint32_t a)
    {
        uint32_t c = (a + 15u) / kSoloChunk;
        if (c >= nchunks) c = nchunks - 1u;
        issue_upto(a / kSoloChunk + kSoloAhead);
        wait_upto(c);
    }
    TC_MFN uint32_t loaded_x() const { return waited * kSoloChunk; }
    TC_MFN void prefetch(uint32_t x) { issue_upto(x / kSoloChunk + kSoloAhead); }
    TC_MFN void drain() { if (issued) wait_upto(issued - 1u); }   // every request waited: slots reusable
};

// tc::Stream over the ring.  kAhead: this view may touch bytes that have not been requested yet (the cursor side);
// the candidate side only ever reads behind the cursor side.
template <bool kAhead>
struct RingStream {
    SoloFeed *f;
    uint64_t lo, hi;
    uint32_t wb;
    TC_MFN void init(SoloFeed *feed) { f = feed; lo = hi = 0; wb = 0xffffffffu; }
    TC_MFN uint64_t word(uint32_t a) const
    {
        return *reinterpret_cast<const uint64_t *>(f->ring + a % kSoloRing);
    }
    TC_MFN uint64_t rd8(uint32_t pos)
    {
        const uint32_t x = pos + f->mis, a = x & ~7u;
        if (a != wb) {
            if (kAhead && a + 16u > f->loaded_x()) f->need(a);
            lo = (a == wb + 8u) ? hi : word(a);
            hi = word(a + 8u);
            wb = a;
        }
        const uint32_t sh = (x & 7u) * 8u;
        return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
    }
    // single bytes are only read by the backward extension (compress.rs:272-287), which may run further back than the
    // ring's history: a chunk that has been overwritten is read from global memory instead
    TC_MFN uint32_t byte(uint32_t pos) const
    {
        const uint32_t x = pos + f->mis;
        if (x / kSoloChunk + kSoloSlots >= f->issued) return f->ring[x % kSoloRing];
        return TC_LD8_RO(f->src_al + x);
    }
};

}  // namespace lz4b200
