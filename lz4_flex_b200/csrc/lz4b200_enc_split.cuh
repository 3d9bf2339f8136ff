// lz4b200_enc_split.cuh — K1 as a two-warp pipeline per block (included by lz4b200_kernels.cuh).
//
// compress_internal (reference src/block/compress.rs:318-489) does two things per sequence: it SEARCHES
// (probe loop :373-439, backward/forward extension :272-287 / :156-216, the cur-2 re-insert :460-461) and it
// EMITS (token, length bytes, literals, offset: :463-486).  Only the search is a serial chain — the next
// probe starts where this match ends — while emission depends on nothing but the finished
// (anchor, match start, offset, match end) tuple.  So a block is handled by a PAIR of warps:
//
//   matcher warp : owns the 4096-slot table in shared memory and runs the exact emulation of the sequential
//                  probe loop (32 probes per batch, in-batch table forwarding with match.any); it pushes one
//                  16-byte tuple per sequence into a shared-memory ring and goes straight on to the next probe.
//   emitter warp : takes 32 tuples at a time, one per lane; every lane sizes its own sequence, a warp
//                  exclusive scan (__shfl_up) turns sizes into output offsets, and the lanes write token /
//                  length bytes / literals / offset of 32 sequences at once (long literal runs are copied by the
//                  whole warp).  This is the scan-compacted emission the north star asks for, and it takes
//                  ~27 % of the per-sequence chain off the matcher.
//
// Hand-off: two halves of 32 tuples, mbarriers "full[h]" / "empty[h]" per pair in shared memory — the matcher
// never waits unless the emitter is two batches behind.
#pragma once

namespace lz4b200 {

constexpr uint32_t kSeqBatchEntries = 32;
constexpr uint32_t kExitBlock = 0xffffffffu;
constexpr uint32_t kSmallLit = 24;            // literal runs up to this length are copied by the owning lane

// mbarrier hand-off (shared-memory barrier objects; named barriers would cost 16 hardware barriers per CTA
// and cap the resident CTAs).  One arrival per phase, from lane 0 of the signalling warp, after its writes.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.release.cta.shared::cta.b64 st, [%0]; }" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_addr(bar);
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

// Producer side of the tuple ring (all state is warp-uniform).
struct SeqProducer {
    uint4 *q;                 // [2][32] tuples: (anchor, match start, offset, match end); offset 0 = last literals
    volatile uint32_t *meta;  // [2][4]: block index, tuple count, flags (1 = first batch of the block, 2 = last)
    uint64_t *bars;           // full[0], full[1], empty[0], empty[1]
    uint32_t k;               // batches handed over so far
    uint32_t qn;              // tuples in the current half
    uint32_t block;
    uint32_t first;

    __device__ __forceinline__ void flush(uint32_t last, uint32_t lane)
    {
        const uint32_t h = k & 1u;
        if (lane == 0) {
            meta[h * 4 + 0] = block;
            meta[h * 4 + 1] = qn;
            meta[h * 4 + 2] = first | (last << 1);
            mbar_arrive(bars + h);                             // release: tuples + header visible to the emitter
        }
        k++; qn = 0; first = 0;
        // batch k goes into half k&1, last used by batch k-2: wait until the emitter has read it
        if (k >= 2) mbar_wait(bars + 2 + (k & 1u), ((k >> 1) - 1u) & 1u);
        __syncwarp();
    }
    __device__ __forceinline__ void push(uint32_t anchor, uint32_t mpos, uint32_t dist, uint32_t end, uint32_t lane)
    {
        if (lane == 0) q[(k & 1u) * kSeqBatchEntries + qn] = make_uint4(anchor, mpos, dist, end);
        qn++;
        if (qn == kSeqBatchEntries) flush(0, lane);
    }
    // the last tuple of a block (literals only): hand the batch over with the "last" flag
    __device__ __forceinline__ void push_final(uint32_t anchor, uint32_t n, uint32_t lane)
    {
        if (lane == 0) q[(k & 1u) * kSeqBatchEntries + qn] = make_uint4(anchor, 0u, 0u, n);
        qn++;
        flush(1, lane);
    }
};

// ---------------------------------------------------------------------------------------------
// matcher: the search half of compress_internal.  Table semantics identical to encode_block_v1.
// ---------------------------------------------------------------------------------------------
template <typename TabT>
__device__ __forceinline__ void match_block(const uint8_t *__restrict__ src, uint32_t n, TabT *tab, bool cont, bool h5,
                                            SeqProducer &pr, uint32_t lane)
{
    constexpr uint32_t kInvalid = TabTraits<TabT>::kInvalid;
    const uint32_t lt_mask = (1u << lane) - 1u;
    if (n < 13) {                                               // compress.rs:343-346
        pr.push_final(0, n, lane);
        return;
    }
    {
        constexpr uint32_t words = 4096 * sizeof(TabT) / 16;
        const uint32_t f = cont ? 0xffffffffu : 0u;
        uint4 *t128 = reinterpret_cast<uint4 *>(tab);
#pragma unroll 4
        for (uint32_t i = lane; i < words; i += 32) t128[i] = make_uint4(f, f, f, f);
        __syncwarp();
    }
    const WordView view(src);
    const uint32_t last_probe = n - 12;
    const uint32_t lim = n - 6;                                 // matches end before the last END_OFFSET bytes
    uint32_t anchor = 0, cur = 0;
    bool ri = false;                                            // T[H(cur-2)] = cur-2 still owed (compress.rs:460-461)
    if (!cont) {                                                // compress.rs:353-359
        uint32_t lo, hi; view.ro5(0, lo, hi);
        const uint32_t s = h5 ? slot_h5(lo, hi) : slot_h4(lo);
        if (lane == 0) tab[s] = 0;
        cur = 1;
        __syncwarp();
    }

    for (;;) {                                                  // one sequence per iteration
        if (lane < 2) prefetch_l1(src + min(cur + 192u + 128u * lane, n - 1u));
        uint32_t base = cur, stride = 1, cand, mpos;
        for (;;) {                                              // probe batches: compress.rs:373-439
            const uint32_t p = base + lane * stride;
            const bool term = p > last_probe;
            uint32_t v4, hi;
            view.ro5(term ? 0u : p, v4, hi);
            if (ri) {
                // the re-insert of the previous sequence rides along with the first probe loads
                uint32_t lo2, hi2; view.ro5(cur - 2u, lo2, hi2);
                const uint32_t s2 = h5 ? slot_h5(lo2, hi2) : slot_h4(lo2);
                if (lane == 0) tab[s2] = (TabT)(cur - 2u);
                __syncwarp();
                ri = false;
            }
            uint32_t key = h5 ? slot_h5(v4, hi) : slot_h4(v4);
            uint32_t cnd = kInvalid;
            if (!term) cnd = tab[key]; else key = 0x10000u | lane;
            const uint32_t same = __match_any_sync(kFull, key);
            const uint32_t prior = same & lt_mask;
            if (prior) cnd = base + (31u - __clz(prior)) * stride;   // forwarded in-batch write
            const bool chk = !term && cnd != kInvalid && p - cnd <= 65535u;
            const bool hit = chk & (view.ro4(chk ? cnd : 0u) == v4);
            const uint32_t hits = __ballot_sync(kFull, hit), terms = __ballot_sync(kFull, term);
            const uint32_t win = hits ? (uint32_t)__ffs(hits) - 1u : 32u;
            const uint32_t tfirst = terms ? (uint32_t)__ffs(terms) - 1u : 32u;
            if (tfirst < win) {                                 // compress.rs:381-384: the rest is literals
                pr.push_final(anchor, n, lane);
                return;
            }
            // commit the table writes of probes 0..win (last writer per slot wins)
            const uint32_t upto = win < 32u ? win : 31u;
            const uint32_t le_mask = upto == 31u ? kFull : ((2u << upto) - 1u);
            const uint32_t mine = same & le_mask;
            if (lane <= upto && (31u - __clz(mine)) == lane) tab[key] = (TabT)p;
            __syncwarp();
            if (win < 32u) {
                mpos = __shfl_sync(kFull, p, win);
                cand = __shfl_sync(kFull, cnd, win);
                break;
            }
            base += 32u * stride;
            stride++;
        }
        const uint32_t dist = mpos - cand;

        // ---- extension: the first forward round (32 bytes) and the backward round share one memory round trip
        const uint32_t room = min(cand, mpos - anchor);         // how far both sides may step back (0 for literal-free sequences)
        const uint32_t qf = mpos + 4u + lane;
        const bool inf = qf < lim;
        const uint8_t f1 = __ldg(src + (inf ? qf : mpos)), f2 = __ldg(src + (inf ? qf : mpos) - dist);
        uint32_t kb = 0;
        if (room) {                                             // compress.rs:272-287
            const bool inb = lane < room;
            const uint8_t b1 = __ldg(src + mpos - (inb ? 1u + lane : 0u)), b2 = __ldg(src + cand - (inb ? 1u + lane : 0u));
            const uint32_t bad = ~__ballot_sync(kFull, inb && b1 == b2);
            kb = bad ? (uint32_t)__ffs(bad) - 1u : 32u;
        }
        const uint32_t badf = ~__ballot_sync(kFull, inf && f1 == f2);
        const uint32_t kf = badf ? (uint32_t)__ffs(badf) - 1u : 32u;
        uint32_t end = mpos + 4u + kf;
        if (kb) {
            mpos -= kb; cand -= kb;
            while (kb == 32u) {                                 // more than 32 bytes backwards: rare
                const uint32_t room2 = min(cand, mpos - anchor);
                const bool inb = lane < room2;
                const uint8_t b1 = __ldg(src + mpos - (inb ? 1u + lane : 0u)), b2 = __ldg(src + cand - (inb ? 1u + lane : 0u));
                const uint32_t bad = ~__ballot_sync(kFull, inb && b1 == b2);
                kb = bad ? (uint32_t)__ffs(bad) - 1u : 32u;
                mpos -= kb; cand -= kb;
            }
        }
        if (kf == 32u) {                                        // long match: 128 bytes per round (compress.rs:156-216)
            for (;;) {
                const uint32_t pos = end + 4u * lane;
                const bool full = pos + 4u <= lim;              // this lane's word lies before n - END_OFFSET
                const uint32_t x = view.ro4(full ? pos : 0u) ^ view.ro4(full ? pos - dist : 0u);
                const uint32_t nm = full ? (x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u) : 0u;
                const uint32_t bad = __ballot_sync(kFull, nm < 4u);
                if (bad) {
                    const uint32_t fl = (uint32_t)__ffs(bad) - 1u;
                    end += 4u * fl + __shfl_sync(kFull, nm, fl);
                    break;
                }
                end += 128u;
            }
            if (end < lim) {                                    // a word that crossed n - 6: at most 3 more bytes
                const uint32_t q = end + lane;
                const bool ok = lane < 4u && q < lim && __ldg(src + (q < lim ? q : end)) == __ldg(src + (q < lim ? q : end) - dist);
                end += (uint32_t)__ffs(~__ballot_sync(kFull, ok)) - 1u;
            }
        }
        pr.push(anchor, mpos, dist, end, lane);
        anchor = cur = end;
        ri = true;
    }
}

// ---------------------------------------------------------------------------------------------
// emitter: compress.rs:463-486 for 32 sequences at a time.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t *put_len_ext(uint8_t *p, uint32_t v)     // v = length - 15 (write_integer, compress.rs:217-242)
{
    while (v >= 255u) { *p++ = 0xff; v -= 255u; }
    *p++ = (uint8_t)v;
    return p;
}

__device__ __forceinline__ void emit_loop(const BatchArgs &a, const uint4 *q, const volatile uint32_t *meta, uint64_t *bars,
                                          uint32_t lane)
{
    const uint8_t *src = nullptr;
    uint8_t *dst = nullptr;
    uint32_t o = 0;
    for (uint32_t j = 0;; j++) {
        const uint32_t h = j & 1u;
        mbar_wait(bars + h, (j >> 1) & 1u);
        const uint32_t b = meta[h * 4 + 0], cnt = meta[h * 4 + 1], fl = meta[h * 4 + 2];
        uint4 e = make_uint4(0, 0, 0, 0);
        if (lane < cnt && b != kExitBlock) e = q[h * kSeqBatchEntries + lane];
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 2 + h);              // tuples are in registers: the half may be refilled
        if (b == kExitBlock) break;
        if (fl & 1u) { src = a.in + a.in_off[b]; dst = a.out + a.out_off[b]; o = 0; }

        const bool valid = lane < cnt;
        const bool tail = e.z == 0;                            // last literals: no match part
        const uint32_t lit = valid ? (tail ? e.w : e.y) - e.x : 0u;
        const uint32_t extra = (valid && !tail) ? e.w - e.y - 4u : 0u;
        const uint32_t lit_ext = lit >= 15u ? (lit - 15u) / 255u + 1u : 0u;
        const uint32_t m_ext = extra >= 15u ? (extra - 15u) / 255u + 1u : 0u;
        const uint32_t size = valid ? 1u + lit_ext + lit + (tail ? 0u : 2u + m_ext) : 0u;
        uint32_t incl = size;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        uint8_t *p = dst + o + (incl - size);
        uint8_t *lit_at = p;
        if (valid) {
            *p++ = (uint8_t)(((lit < 15u ? lit : 15u) << 4) | (tail ? 0u : (extra < 15u ? extra : 15u)));
            if (lit >= 15u) p = put_len_ext(p, lit - 15u);
            lit_at = p;
            if (lit <= kSmallLit) {
                const uint8_t *s = src + e.x;
                for (uint32_t i = 0; i < lit; i++) p[i] = __ldg(s + i);
            }
            p += lit;
            if (!tail) {
                p[0] = (uint8_t)e.z; p[1] = (uint8_t)(e.z >> 8);
                p += 2;
                if (extra >= 15u) put_len_ext(p, extra - 15u);
            }
        }
        // long literal runs: the whole warp copies them, one run at a time
        uint32_t big = __ballot_sync(kFull, valid && lit > kSmallLit);
        while (big) {
            const uint32_t l = (uint32_t)__ffs(big) - 1u;
            big &= big - 1u;
            const uint32_t from = __shfl_sync(kFull, e.x, l), len = __shfl_sync(kFull, lit, l);
            const uint32_t at = __shfl_sync(kFull, (uint32_t)(lit_at - dst), l);
            const uint8_t *s = src + from;
            uint8_t *d = dst + at;
            uint32_t i = lane;
            for (; i + 96u < len; i += 128u) {
                const uint8_t c0 = __ldg(s + i), c1 = __ldg(s + i + 32), c2 = __ldg(s + i + 64), c3 = __ldg(s + i + 96);
                d[i] = c0; d[i + 32] = c1; d[i + 64] = c2; d[i + 96] = c3;
            }
            for (; i < len; i += 32u) d[i] = __ldg(s + i);
        }
        o += total;
        if ((fl & 2u) && lane == 0) { a.out_len[b] = o; a.status[b] = LZ4B200_OK; }
    }
}

// One CTA = kPairs matcher warps (warps 0..kPairs-1) + kPairs emitter warps.  Shared memory per pair: the
// table (8 KiB for blocks <= 64 KiB, 16 KiB above), 1 KiB of tuples, 32 bytes of batch headers, 4 mbarriers.
template <typename TabT, int kPairs>
__global__ void __launch_bounds__(kPairs * 64)
lz4_compress_blocks_split(BatchArgs a, uint32_t *tickets)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t pair = warp < (uint32_t)kPairs ? warp : warp - kPairs;
    TabT *tab = reinterpret_cast<TabT *>(smem_raw) + pair * 4096;
    uint4 *q = reinterpret_cast<uint4 *>(smem_raw + kPairs * 4096 * sizeof(TabT)) + pair * 2 * kSeqBatchEntries;
    uint32_t *meta = reinterpret_cast<uint32_t *>(smem_raw + kPairs * (4096 * sizeof(TabT) + 2 * kSeqBatchEntries * 16)) + pair * 8;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + kPairs * (4096 * sizeof(TabT) + 2 * kSeqBatchEntries * 16 + 32)) + pair * 4;
    if (threadIdx.x < (uint32_t)kPairs * 4u) mbar_init(reinterpret_cast<uint64_t *>(smem_raw + kPairs * (4096 * sizeof(TabT) + 2 * kSeqBatchEntries * 16 + 32)) + threadIdx.x, 1u);
    __syncthreads();
    if (warp >= (uint32_t)kPairs) {
        emit_loop(a, q, meta, bars, lane);
        return;
    }
    constexpr bool kSmall = sizeof(TabT) == 2;
    SeqProducer pr{q, meta, bars, 0u, 0u, 0u, 0u};
    for (uint32_t b = next_ticket(tickets); b < a.nblocks; b = next_ticket(tickets)) {
        const uint32_t n = a.in_len[b];
        if ((n <= 65536u) != kSmall) continue;
        const uint32_t fl = a.flags ? a.flags[b] : 0u;
        if ((uint64_t)a.out_cap[b] < max_output_size_dev(n)) {              // compress.rs:338-340
            if (lane == 0) { a.out_len[b] = 0; a.status[b] = LZ4B200_COMPRESS_OUTPUT_TOO_SMALL; }
            continue;
        }
        const bool h5 = (fl & LZ4B200_BLOCK_HASH5_ALWAYS) || n >= 65535u;   // compress.rs:559
        pr.block = b; pr.first = 1;
        match_block<TabT>(a.in + a.in_off[b], n, tab, (fl & LZ4B200_BLOCK_CONT) != 0, h5, pr, lane);
    }
    pr.block = kExitBlock; pr.first = 0;
    pr.flush(0, lane);
    retire_warp(tickets, gridDim.x * kPairs);
}

template <typename TabT, int kPairs>
constexpr size_t split_smem_bytes() { return (size_t)kPairs * (4096 * sizeof(TabT) + 2 * kSeqBatchEntries * 16 + 32 + 32); }

}  // namespace lz4b200
