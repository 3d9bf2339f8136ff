// Host build of the per-thread codec (lz4_flex_b200/csrc/lz4b200_thread_codec.cuh) for CPU-side parity tests:
// the very functions the K1-T / K2-T kernels run per lane, compiled by g++ and compared with the oracle by
// tests/test_thread_codec_host.py.  Test infrastructure only: nothing in the product links or loads this.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../lz4_flex_b200/csrc/lz4b200_thread_codec.cuh"
#include "../../lz4_flex_b200/csrc/lz4b200_solo_ring.cuh"

using namespace lz4b200::tc;

extern "C" {

// flags: LZ4B200_BLOCK_CONT | LZ4B200_BLOCK_HASH5_ALWAYS (include/lz4b200.h); returns the compressed size
uint32_t tc_host_compress(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t flags)
{
    const bool cont = (flags & LZ4B200_BLOCK_CONT) != 0;
    const bool h5 = (flags & LZ4B200_BLOCK_HASH5_ALWAYS) || n >= 65535u;
    if (n <= 65536u) {
        std::vector<uint16_t> tab(4096, cont ? 0xffffu : 0u);
        return encode_block_thread<uint16_t>(in, n, out, tab.data(), cont, h5);
    }
    std::vector<uint32_t> tab(4096, cont ? 0xffffffffu : 0u);
    return encode_block_thread<uint32_t>(in, n, out, tab.data(), cont, h5);
}

// The K1-S parse: the same parse over the shared-memory ring views (lz4b200_solo_ring.cuh), with the ring in host memory
// and memcpy standing in for the TMA bulk copies; u32 table as in the kernel.  Emission through DirectSink.
uint32_t tc_host_compress_solo(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t flags)
{
    using namespace lz4b200;
    const bool cont = (flags & LZ4B200_BLOCK_CONT) != 0;
    const bool h5 = (flags & LZ4B200_BLOCK_HASH5_ALWAYS) || n >= 65535u;
    std::vector<uint32_t> tab(4096, cont ? 0xffffffffu : 0u);
    std::vector<uint8_t> ring(kSoloRing + 16, 0xEE);
    std::vector<uint64_t> bars(kSoloSlots);
    SoloFeed feed;
    feed.ring = ring.data() + ((16 - reinterpret_cast<uintptr_t>(ring.data()) % 16) % 16);
    feed.bars = bars.data(); feed.phases = 0;
    feed.begin(in, n);
    RingStream<true> a;
    RingStream<false> b;
    a.init(&feed); b.init(&feed);
    Stream<true> lits;
    lits.init(in, n);
    DirectSink<Stream<true>> sink;
    sink.out.init(out);
    sink.lits = &lits;
    parse_block_thread<uint32_t>(a, b, n, tab.data(), cont, h5, sink);
    feed.drain();
    return sink.out.produced();
}

int tc_host_decompress(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t cap, uint32_t *written, uint64_t *expected)
{
    ThreadDecResult r = decode_block_thread(in, n, out, cap);
    *written = r.status == LZ4B200_OK ? r.written : 0u;
    *expected = r.expected;
    return r.status;
}

}  // extern "C"
