// Exercises include/lz4_flex.hpp (the C++ mirror of the lz4_flex API) against the C ABI.
//   test_mirror nogpu : no CUDA device -> every compute call must fail loudly (no CPU fallback)
//   test_mirror gpu   : round trips, reference error variants, FrameEncoder == one-shot C frame call
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "lz4_flex.hpp"

static int failures = 0;
#define CHECK(cond)                                                       \
    do {                                                                  \
        if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

struct VecSink {
    std::vector<uint8_t> bytes;
    void write(const uint8_t *p, size_t n) { bytes.insert(bytes.end(), p, p + n); }
};

static std::vector<uint8_t> make_data(size_t n)
{
    std::vector<uint8_t> d(n);
    const char *words[] = {"{\"id\":", ",\"name\":\"sensor-", "\",\"value\":", ",\"ok\":true}", "\n", "temperature", "pressure"};
    size_t pos = 0, k = 0;
    uint32_t x = 12345;
    while (pos < n) {
        x = x * 1664525u + 1013904223u;
        std::string w = words[k++ % 7];
        if ((x >> 28) < 6) w += std::to_string(x >> 12);
        for (size_t i = 0; i < w.size() && pos < n; i++) d[pos++] = (uint8_t)w[i];
    }
    return d;
}

int main(int argc, char **argv)
{
    using namespace lz4_flex;
    const bool gpu = argc > 1 && std::string(argv[1]) == "gpu";
    static_assert(block::get_maximum_output_size(65536) == 72109, "compress.rs:588-590");
    CHECK(block::get_maximum_output_size(0) == 20);
    CHECK(frame::block_size_from_buf_length(65537) == frame::BlockSize::Max256KB);
    CHECK(block::uncompressed_size((const uint8_t *)"\x05\x00", 2).is_err());
    if (!gpu) {
        std::vector<uint8_t> d = make_data(1000), out(block::get_maximum_output_size(1000));
        auto r = block::compress_into(d.data(), d.size(), out.data(), out.size());
        CHECK(r.is_err() && r.error().kind == block::CompressError::Cuda);
        CHECK(block::compress(d.data(), d.size()).empty());
        printf("nogpu %s\n", failures ? "FAILED" : "ok");
        return failures ? 1 : 0;
    }
    // ---- block API ------------------------------------------------------------------------------
    std::vector<uint8_t> d = make_data(300000);
    auto c = block::compress_prepend_size(d.data(), d.size());
    CHECK(c.size() > 4 && c.size() < d.size() / 2);
    auto back = block::decompress_size_prepended(c.data(), c.size());
    CHECK(back.is_ok() && back.value() == d);
    {
        std::vector<uint8_t> small(10);
        auto r = block::compress_into(d.data(), 1000, small.data(), small.size());          // compress.rs:338-340
        CHECK(r.is_err() && r.error().kind == block::CompressError::OutputTooSmall);
        const uint8_t v1[] = {0x20, 'a', 'a', 1, 0};
        uint8_t o[4];
        auto e = block::decompress_into(v1, 5, o, 1);                                        // decompress.rs:571-577
        CHECK(e.is_err() && e.error().kind == block::DecompressError::OutputTooSmall && e.error().expected == 2 &&
              e.error().actual == 1);
        const uint8_t v2[] = {0x30, 'a', '4', '9'};
        auto ok = block::decompress_into(v2, 4, o, 3);                                       // decompress.rs:535-537
        CHECK(ok.is_ok() && ok.value() == 3 && memcmp(o, "a49", 3) == 0);
        auto ez = block::decompress_into(v2, 0, o, 3);
        CHECK(ez.is_err() && ez.error().kind == block::DecompressError::ExpectedAnotherByte);
    }
    // ---- dictionary and reusable-table API (compress.rs:610-616, 685-694, 744-766; decompress.rs:462-528) ----
    {
        const uint8_t in[] = {10, 12, 14, 16, 18, 10, 12, 14, 16, 18, 10, 12, 14, 16, 18, 10, 12, 14, 16, 18};   // compress.rs:892-911
        auto plain = block::compress(in, sizeof in);
        auto with = block::compress_with_dict(in, sizeof in, in, sizeof in);
        CHECK(!with.empty() && with.size() < plain.size());
        auto rt = block::decompress_with_dict(with.data(), with.size(), sizeof in, in, sizeof in);
        CHECK(rt.is_ok() && rt.value() == std::vector<uint8_t>(in, in + sizeof in));
        const uint8_t tiny[] = {10, 12, 14};                                                      // compress.rs:913-919
        CHECK(block::compress_with_dict(in, sizeof in, tiny, 3) == plain);
        auto p = block::compress_prepend_size_with_dict(d.data() + 100000, 50000, d.data(), 100000);
        auto q = block::decompress_size_prepended_with_dict(p.data(), p.size(), d.data(), 100000);
        CHECK(q.is_ok() && q.value() == std::vector<uint8_t>(d.begin() + 100000, d.begin() + 150000));
        CHECK(p.size() < block::compress_prepend_size(d.data() + 100000, 50000).size());
        const uint8_t oob[] = {0x0E, 255, 0, 0x70, 0, 0, 0, 0, 0, 0, 0};                          // decompress.rs:593-601
        std::vector<uint8_t> zeros(250, 0);
        auto e = block::decompress_with_dict(oob, sizeof oob, 256, zeros.data(), zeros.size());
        CHECK(e.is_err() && e.error().kind == block::DecompressError::OffsetOutOfBounds);
        // CompressTable: a Small table is upgraded by an input >= 65 535 bytes and stays Large (5-byte hash)
        block::CompressTable t = block::CompressTable::small();
        std::vector<uint8_t> o1(block::get_maximum_output_size(d.size())), o2(o1.size());
        auto a = block::compress_into_with_table(d.data(), 30000, o1.data(), o1.size(), t);
        CHECK(a.is_ok() && t.kind == block::CompressTable::Small);
        auto ref30k = block::compress(d.data(), 30000);
        CHECK(a.is_ok() && std::vector<uint8_t>(o1.begin(), o1.begin() + a.value()) == ref30k);
        auto b = block::compress_into_with_table(d.data(), 70000, o2.data(), o2.size(), t);
        CHECK(b.is_ok() && t.kind == block::CompressTable::Large);
        auto c2 = block::compress_into_with_table(d.data(), 30000, o2.data(), o2.size(), t);
        CHECK(c2.is_ok() && t.kind == block::CompressTable::Large);
        auto back30k = block::decompress(o2.data(), c2.value(), 30000);
        CHECK(back30k.is_ok() && back30k.value() == std::vector<uint8_t>(d.begin(), d.begin() + 30000));
    }
    // ---- frame API: chunked writes through FrameEncoder == one-shot C call ----------------------
    for (int variant = 0; variant < 3; variant++) {
        frame::FrameInfo info;
        info.block_size(variant == 2 ? frame::BlockSize::Max256KB : frame::BlockSize::Max64KB);
        if (variant == 1) info.block_checksums(true).content_checksum(true).content_size(d.size());
        frame::FrameEncoder<VecSink> enc(VecSink{}, info);
        for (size_t p = 0; p < d.size(); p += 7777) enc.write(d.data() + p, std::min<size_t>(7777, d.size() - p));
        auto fin = std::move(enc).finish();
        CHECK(fin.is_ok());
        const lz4b200_frame_info ci = info.to_c();
        std::vector<uint8_t> ref(lz4b200_frame_bound(d.size(), &ci));
        size_t w = 0;
        CHECK(lz4b200_frame_compress(default_context(), d.data(), d.size(), &ci, d.size(), ref.data(), ref.size(), &w) == LZ4B200_OK);
        ref.resize(w);
        CHECK(fin.value().bytes == ref);
        frame::FrameDecoder dec(ref.data(), ref.size());
        auto all = dec.read_to_end();
        CHECK(all.is_ok() && all.value() == d);
    }
    {   // concatenated frames: read_to_end() stops at each EndMark (tests/tests.rs:633-647)
        frame::FrameInfo info; info.block_size(frame::BlockSize::Max64KB);
        const lz4b200_frame_info ci = info.to_c();
        std::vector<uint8_t> a(d.begin(), d.begin() + 30000), b(d.begin() + 30000, d.end());
        std::vector<uint8_t> fa(lz4b200_frame_bound(a.size(), &ci)), fb(lz4b200_frame_bound(b.size(), &ci));
        size_t wa = 0, wb = 0;
        CHECK(lz4b200_frame_compress(default_context(), a.data(), a.size(), &ci, a.size(), fa.data(), fa.size(), &wa) == LZ4B200_OK);
        CHECK(lz4b200_frame_compress(default_context(), b.data(), b.size(), &ci, b.size(), fb.data(), fb.size(), &wb) == LZ4B200_OK);
        std::vector<uint8_t> cat(fa.begin(), fa.begin() + (long)wa);
        cat.insert(cat.end(), fb.begin(), fb.begin() + (long)wb);
        frame::FrameDecoder dec(cat.data(), cat.size());
        auto r1 = dec.read_to_end();
        auto r2 = dec.read_to_end();
        auto r3 = dec.read_to_end();
        CHECK(r1.is_ok() && r1.value() == a);
        CHECK(r2.is_ok() && r2.value() == b);
        CHECK(r3.is_ok() && r3.value().empty());
    }
    {
        frame::FrameInfo info; info.content_size(3);
        frame::FrameEncoder<VecSink> enc(VecSink{}, info);
        enc.write(d.data(), 725);
        auto r = enc.try_finish();                                                           // tests/tests.rs:721-734
        CHECK(r.is_err() && r.error().status == LZ4B200_FRAME_CONTENT_LENGTH && r.error().expected == 3 && r.error().actual == 725);
    }
    printf("gpu %s\n", failures ? "FAILED" : "ok");
    return failures ? 1 : 0;
}
