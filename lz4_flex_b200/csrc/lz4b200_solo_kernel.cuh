// lz4b200_solo_kernel.cuh — K1-S: one compress chain per CTA, everything it touches in shared memory.
//
// A batch of a few hundred big blocks (BASELINE config 4: 256 x 4 MiB per GPU) is a few hundred serial chains
// (compress_internal, reference src/block/compress.rs:318-489): there is nothing to run in parallel but the chains
// themselves, so the only lever is the latency of one probe.  Here a chain gets half an SM's shared memory:
//   * its 4096-entry u32 table (16 KiB; hashtable.rs:52-53,121-127),
//   * an 80 KiB ring holding the input window [cursor - 64 KiB, cursor + 12 KiB): every candidate the format allows
//     (MAX_DISTANCE 65 535, block/mod.rs:64) is a shared-memory read, not an L2/HBM round trip,
//   * the ring is filled 2 KiB at a time by TMA bulk copies (cp.async.bulk.shared.global, completion on one mbarrier
//     per ring slot) issued six chunks ahead of the furthest byte the parse has touched.
// The matcher is ONE thread running tc::parse_block_thread (the same sequential parse the K1-T kernel runs per lane)
// over ring views; it hands (anchor, match start, offset, match end) tuples to an emitter warp through the tuple
// queue of lz4b200_enc_split.cuh, and the emitter writes the byte stream with warp-wide scans and coalesced stores
// (emit_batch).  Probe latency: ~100 cycles of LDS + integer work instead of two dependent L2/HBM accesses.
#pragma once
#include "lz4b200_solo_ring.cuh"

namespace lz4b200 {

// Single-thread producer side of the tuple queue (same protocol as SeqProducer, lz4b200_enc_split.cuh).
struct SoloSink {
    uint4 *q;
    volatile uint32_t *meta;
    uint64_t *bars;            // full[0], full[1], empty[0], empty[1]
    uint32_t k, qn, block, first;
    SoloFeed *feed;

    __device__ __forceinline__ void flush(uint32_t last)
    {
        const uint32_t h = k & 1u;
        meta[h * 4 + 0] = block;
        meta[h * 4 + 1] = qn;
        meta[h * 4 + 2] = first | (last << 1);
        mbar_arrive(bars + h);
        k++; qn = 0; first = 0;
        if (k >= 2) mbar_wait(bars + 2 + (k & 1u), ((k >> 1) - 1u) & 1u);
    }
    __device__ __forceinline__ void sequence(uint32_t anchor, uint32_t mpos, uint32_t dist, uint32_t end)
    {
        q[(k & 1u) * kSeqBatchEntries + qn] = make_uint4(anchor, mpos, dist, end);
        if (++qn == kSeqBatchEntries) flush(0);
        feed->prefetch(end + feed->mis);                   // keep the TMA three chunks ahead of the cursor
    }
    __device__ __forceinline__ void tail(uint32_t anchor, uint32_t n)
    {
        q[(k & 1u) * kSeqBatchEntries + qn] = make_uint4(anchor, 0u, 0u, n);
        qn++;
        flush(1);
    }
};

constexpr size_t kSoloSmemBytes = kSoloRing + 4096 * 4 + 2 * kSeqBatchEntries * 16 + 32 + 4 * 8 + kSoloSlots * 8;

// warp 0, lane 0: matcher; warp 1: emitter.
__global__ void __launch_bounds__(64)
lz4_compress_blocks_solo(BatchArgs a, uint32_t *tickets)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *ring = smem_raw;
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem_raw + kSoloRing);
    uint4 *q = reinterpret_cast<uint4 *>(smem_raw + kSoloRing + 16384);
    uint32_t *meta = reinterpret_cast<uint32_t *>(smem_raw + kSoloRing + 16384 + 2 * kSeqBatchEntries * 16);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + kSoloRing + 16384 + 2 * kSeqBatchEntries * 16 + 32);
    uint64_t *rbars = bars + 4;
    if (threadIdx.x < 4u + kSoloSlots) mbar_init(bars + threadIdx.x, 1u);
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    if (warp == 1u) {
        emit_loop(a, q, meta, bars, lane);
        return;
    }
    if (lane != 0u) return;
    SoloFeed feed;
    feed.ring = ring; feed.bars = rbars; feed.phases = 0ull;
    SoloSink sink{q, meta, bars, 0u, 0u, 0u, 0u, &feed};
    for (;;) {
        const uint32_t b = atomicAdd(&tickets[0], 1u);
        if (b >= a.nblocks) break;
        const uint32_t n = a.in_len[b];
        const uint32_t fl = a.flags ? a.flags[b] : 0u;
        if ((uint64_t)a.out_cap[b] < max_output_size_dev(n)) {              // compress.rs:338-340
            a.out_len[b] = 0; a.status[b] = LZ4B200_COMPRESS_OUTPUT_TOO_SMALL;
            continue;
        }
        const bool cont = (fl & LZ4B200_BLOCK_CONT) != 0;
        const bool h5 = (fl & LZ4B200_BLOCK_HASH5_ALWAYS) || n >= 65535u;   // compress.rs:559
        {
            const uint32_t f = cont ? 0xffffffffu : 0u;
            uint4 *t128 = reinterpret_cast<uint4 *>(tab);
#pragma unroll 8
            for (uint32_t i = 0; i < 1024u; i++) t128[i] = make_uint4(f, f, f, f);
        }
        feed.begin(a.in + a.in_off[b], n);
        RingStream<true> in;
        RingStream<false> cs;
        in.init(&feed); cs.init(&feed);
        sink.block = b; sink.first = 1;
        tc::parse_block_thread<uint32_t>(in, cs, n, tab, cont, h5, sink);
        feed.drain();
    }
    sink.block = kExitBlock; sink.first = 0;
    sink.flush(0);
    __threadfence();
    if (atomicAdd(&tickets[1], 1u) == gridDim.x - 1u) {
        tickets[0] = 0;
        tickets[1] = 0;
        __threadfence();
    }
}

}  // namespace lz4b200
