// lz4b200_thread_kernels.cuh — K1-T / K2-T launch shells around the per-thread codec (lz4b200_thread_codec.cuh).
//
// Persistent grids of 4-warp CTAs; kLanes lanes of every warp are active and each active lane pulls block
// indices from the batch's ticket counter.  kLanes < 32 spreads a batch that has fewer blocks than the GPU has
// lanes over more warps (more schedulers busy, fewer chains coupled in lock step); the launcher picks it from
// the batch size.
#pragma once
#include "lz4b200_thread_codec.cuh"

namespace lz4b200 {

constexpr int kThreadCtaWarps = 4;

__device__ __forceinline__ void retire_thread(uint32_t *tickets, uint32_t total_threads)
{
    __threadfence();
    if (atomicAdd(&tickets[1], 1u) == total_threads - 1u) {
        tickets[0] = 0;
        tickets[1] = 0;
        __threadfence();
    }
}

// K1-T.  gtab: 4096 u16 entries per active lane of the grid (blocks of at most 65 536 bytes only).
template <int kLanes>
__global__ void __launch_bounds__(kThreadCtaWarps * 32)
lz4_compress_blocks_thread(BatchArgs a, uint32_t *tickets, uint16_t *gtab)
{
    const uint32_t lane = threadIdx.x & 31u;
    if (lane >= (uint32_t)kLanes) return;
    const uint32_t tid = (blockIdx.x * kThreadCtaWarps + (threadIdx.x >> 5)) * kLanes + lane;
    uint16_t *tab = gtab + (size_t)tid * 4096u;
    for (;;) {
        const uint32_t b = atomicAdd(&tickets[0], 1u);
        if (b >= a.nblocks) break;
        const uint32_t n = a.in_len[b];
        const uint32_t fl = a.flags ? a.flags[b] : 0u;
        if ((uint64_t)a.out_cap[b] < max_output_size_dev(n) || n > 65536u) {  // compress.rs:338-340 (n > 64 KiB: launcher bug)
            a.out_len[b] = 0;
            a.status[b] = LZ4B200_COMPRESS_OUTPUT_TOO_SMALL;
            continue;
        }
        const bool cont = (fl & LZ4B200_BLOCK_CONT) != 0;
        {
            const uint32_t f = cont ? 0xffffffffu : 0u;
            uint4 *t128 = reinterpret_cast<uint4 *>(tab);
#pragma unroll 8
            for (uint32_t i = 0; i < 512u; i++) t128[i] = make_uint4(f, f, f, f);
        }
        const bool h5 = (fl & LZ4B200_BLOCK_HASH5_ALWAYS) || n >= 65535u;    // compress.rs:559
        const uint32_t written = tc::encode_block_thread<uint16_t>(a.in + a.in_off[b], n, a.out + a.out_off[b], tab, cont, h5);
        a.out_len[b] = written;
        a.status[b] = LZ4B200_OK;
    }
    retire_thread(tickets, gridDim.x * kThreadCtaWarps * kLanes);
}

// K2-T (no external dictionary).
template <int kLanes>
__global__ void __launch_bounds__(kThreadCtaWarps * 32)
lz4_decompress_blocks_thread(BatchArgs a)
{
    const uint32_t lane = threadIdx.x & 31u;
    if (lane >= (uint32_t)kLanes) return;
    for (;;) {
        const uint32_t b = atomicAdd(&a.tickets[0], 1u);
        if (b >= a.nblocks) break;
        const tc::ThreadDecResult r = tc::decode_block_thread(a.in + a.in_off[b], a.in_len[b], a.out + a.out_off[b], a.out_cap[b]);
        a.out_len[b] = r.status == LZ4B200_OK ? r.written : 0u;
        a.status[b] = r.status;
        if (a.err_expected) a.err_expected[b] = r.expected;
    }
    retire_thread(a.tickets, gridDim.x * kThreadCtaWarps * kLanes);
}

}  // namespace lz4b200
