// lz4_flex.hpp — C++ host-side mirror of the lz4_flex crate's block and frame API over the lz4b200 C ABI.
//
// lz4_flex is a Rust crate; this image has no Rust toolchain, so the host layer a Rust maintainer would
// write as `extern "C"` + safe wrappers is provided here in C++ with the same names,
// argument meaning and error behaviour:
//
//   lz4_flex::block::{compress, compress_prepend_size, compress_into, get_maximum_output_size,
//                     decompress, decompress_size_prepended, decompress_into, uncompressed_size}
//        reference: src/block/compress.rs:588-692, src/block/decompress.rs:454-517, src/block/mod.rs:151-157
//   lz4_flex::block::{DecompressError, CompressError}          reference: src/block/mod.rs:82-106
//   lz4_flex::frame::{FrameInfo, BlockSize, BlockMode, FrameEncoder, FrameDecoder, Error}
//        reference: src/frame/header.rs:39-192, src/frame/compress.rs:62-404, src/frame/decompress.rs:48-422
//
// Rust `Result<T, E>` becomes lz4_flex::Result<T, E> (no exceptions for bad data, like the crate).
// Everything computes on the GPU through liblz4b200.so; there is no CPU codec behind this header.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "lz4b200.h"

namespace lz4_flex {

template <typename T, typename E> class Result {
public:
    static Result Ok(T v) { Result r; r.ok_ = true; r.val_ = std::move(v); return r; }
    static Result Err(E e) { Result r; r.ok_ = false; r.err_ = std::move(e); return r; }
    bool is_ok() const { return ok_; }
    bool is_err() const { return !ok_; }
    const T &value() const { return val_; }
    T &value() { return val_; }
    const E &error() const { return err_; }
    T unwrap() && { if (!ok_) std::abort(); return std::move(val_); }
private:
    bool ok_ = false;
    T val_{};
    E err_{};
};

// One lz4b200 context per thread and device, created on first use.
inline lz4b200_ctx *default_context(int device = 0)
{
    thread_local lz4b200_ctx *ctx[16] = {};
    if (device < 0 || device >= 16) return nullptr;
    if (!ctx[device] && lz4b200_ctx_create(device, &ctx[device]) != LZ4B200_OK) ctx[device] = nullptr;
    return ctx[device];
}

namespace block {

constexpr size_t WINDOW_SIZE = 64 * 1024;     // block/mod.rs:35
constexpr size_t MINMATCH = 4;                // block/mod.rs:70
constexpr size_t MFLIMIT = 12;                // block/mod.rs:46
constexpr size_t LZ4_MIN_LENGTH = MFLIMIT + 1;
constexpr size_t MAX_DISTANCE = (1 << 16) - 1;

// block::DecompressError (block/mod.rs:82-98)
struct DecompressError {
    enum Kind { None = 0, OutputTooSmall = 2, LiteralOutOfBounds = 3, ExpectedAnotherByte = 4, OffsetZero = 5,
                OffsetOutOfBounds = 6, Cuda = 201 } kind = None;
    size_t expected = 0, actual = 0;          // OutputTooSmall { expected, actual }
    std::string to_string() const
    {
        if (kind == OutputTooSmall)
            return "provided output is too small for the decompressed data, actual " + std::to_string(actual) +
                   ", expected " + std::to_string(expected);
        return lz4b200_status_string((int)kind);
    }
};

// block::CompressError (block/mod.rs:103-106)
struct CompressError {
    enum Kind { None = 0, OutputTooSmall = 1, Cuda = 201 } kind = None;
    std::string to_string() const { return lz4b200_status_string((int)kind); }
};

// block::get_maximum_output_size (compress.rs:588-590)
constexpr size_t get_maximum_output_size(size_t input_len)
{
    return 16 + 4 + (size_t)((uint64_t)input_len * 110 / 100);
}

// block::compress_into (compress.rs:599)
inline Result<size_t, CompressError> compress_into(const uint8_t *input, size_t n, uint8_t *output, size_t cap,
                                                   lz4b200_ctx *ctx = nullptr)
{
    using R = Result<size_t, CompressError>;
    if (!ctx) ctx = default_context();
    if (!ctx) return R::Err({CompressError::Cuda});
    size_t written = 0;
    const lz4b200_status st = lz4b200_compress_into(ctx, input, n, output, cap, &written);
    if (st == LZ4B200_OK) return R::Ok(written);
    return R::Err({st == LZ4B200_COMPRESS_OUTPUT_TOO_SMALL ? CompressError::OutputTooSmall : CompressError::Cuda});
}

// block::compress (compress.rs:679)
inline std::vector<uint8_t> compress(const uint8_t *input, size_t n, lz4b200_ctx *ctx = nullptr)
{
    std::vector<uint8_t> out(get_maximum_output_size(n));
    auto r = compress_into(input, n, out.data(), out.size(), ctx);
    out.resize(r.is_ok() ? r.value() : 0);
    out.shrink_to_fit();
    return out;
}

// block::compress_prepend_size (compress.rs:673)
inline std::vector<uint8_t> compress_prepend_size(const uint8_t *input, size_t n, lz4b200_ctx *ctx = nullptr)
{
    if (!ctx) ctx = default_context();
    std::vector<uint8_t> out(get_maximum_output_size(n) + 4);
    size_t written = 0;
    if (!ctx || lz4b200_compress_prepend_size(ctx, input, n, out.data(), out.size(), &written) != LZ4B200_OK) written = 0;
    out.resize(written);
    out.shrink_to_fit();
    return out;
}

// block::uncompressed_size (block/mod.rs:151-157): (size, rest)
inline Result<std::pair<size_t, const uint8_t *>, DecompressError> uncompressed_size(const uint8_t *input, size_t n)
{
    using R = Result<std::pair<size_t, const uint8_t *>, DecompressError>;
    size_t size = 0;
    if (lz4b200_uncompressed_size(input, n, &size) != LZ4B200_OK) return R::Err({DecompressError::ExpectedAnotherByte});
    return R::Ok({size, input + 4});
}

// block::decompress_into (decompress.rs:454)
inline Result<size_t, DecompressError> decompress_into(const uint8_t *input, size_t n, uint8_t *output, size_t cap,
                                                       lz4b200_ctx *ctx = nullptr)
{
    using R = Result<size_t, DecompressError>;
    if (!ctx) ctx = default_context();
    if (!ctx) return R::Err({DecompressError::Cuda});
    size_t written = 0, expected = 0, actual = 0;
    const lz4b200_status st = lz4b200_decompress_into(ctx, input, n, output, cap, &written, &expected, &actual);
    if (st == LZ4B200_OK) return R::Ok(written);
    DecompressError e;
    e.kind = (st >= 2 && st <= 6) ? (DecompressError::Kind)st : DecompressError::Cuda;
    e.expected = expected; e.actual = actual;
    return R::Err(e);
}

// block::decompress (decompress.rs:508): the result may be shorter than min_uncompressed_size
inline Result<std::vector<uint8_t>, DecompressError> decompress(const uint8_t *input, size_t n,
                                                                size_t min_uncompressed_size,
                                                                lz4b200_ctx *ctx = nullptr)
{
    using R = Result<std::vector<uint8_t>, DecompressError>;
    std::vector<uint8_t> out(min_uncompressed_size);
    auto r = decompress_into(input, n, out.data(), out.size(), ctx);
    if (r.is_err()) return R::Err(r.error());
    out.resize(r.value());
    return R::Ok(std::move(out));
}

// block::decompress_size_prepended (decompress.rs:496)
inline Result<std::vector<uint8_t>, DecompressError> decompress_size_prepended(const uint8_t *input, size_t n,
                                                                               lz4b200_ctx *ctx = nullptr)
{
    using R = Result<std::vector<uint8_t>, DecompressError>;
    auto s = uncompressed_size(input, n);
    if (s.is_err()) return R::Err(s.error());
    return decompress(s.value().second, n - 4, s.value().first, ctx);
}

// block::CompressTable / compress_into_with_table (compress.rs:709-766)
struct CompressTable {
    enum Kind { Small = LZ4B200_TABLE_SMALL, Large = LZ4B200_TABLE_LARGE } kind = Small;
    static CompressTable small() { return CompressTable{Small}; }
    static CompressTable large() { return CompressTable{Large}; }
};

inline Result<size_t, CompressError> compress_into_with_table(const uint8_t *input, size_t n, uint8_t *output, size_t cap,
                                                              CompressTable &table, lz4b200_ctx *ctx = nullptr)
{
    using R = Result<size_t, CompressError>;
    if (!ctx) ctx = default_context();
    if (!ctx) return R::Err({CompressError::Cuda});
    size_t written = 0;
    int kind = (int)table.kind;
    const lz4b200_status st = lz4b200_compress_into_with_table(ctx, input, n, output, cap, &written, &kind);
    table.kind = (CompressTable::Kind)kind;
    if (st == LZ4B200_OK) return R::Ok(written);
    return R::Err({st == LZ4B200_COMPRESS_OUTPUT_TOO_SMALL ? CompressError::OutputTooSmall : CompressError::Cuda});
}

// ---- external dictionary (compress.rs:610-616, 685-694; decompress.rs:462-468, 478-528) ----

// block::compress_into_with_dict (compress.rs:610)
inline Result<size_t, CompressError> compress_into_with_dict(const uint8_t *input, size_t n, uint8_t *output, size_t cap,
                                                             const uint8_t *dict_data, size_t dict_len,
                                                             lz4b200_ctx *ctx = nullptr)
{
    using R = Result<size_t, CompressError>;
    if (!ctx) ctx = default_context();
    if (!ctx) return R::Err({CompressError::Cuda});
    size_t written = 0;
    const lz4b200_status st = lz4b200_compress_into_with_dict(ctx, input, n, dict_data, dict_len, output, cap, &written);
    if (st == LZ4B200_OK) return R::Ok(written);
    return R::Err({st == LZ4B200_COMPRESS_OUTPUT_TOO_SMALL ? CompressError::OutputTooSmall : CompressError::Cuda});
}

// block::compress_with_dict (compress.rs:685)
inline std::vector<uint8_t> compress_with_dict(const uint8_t *input, size_t n, const uint8_t *ext_dict, size_t dict_len,
                                               lz4b200_ctx *ctx = nullptr)
{
    std::vector<uint8_t> out(get_maximum_output_size(n));
    auto r = compress_into_with_dict(input, n, out.data(), out.size(), ext_dict, dict_len, ctx);
    out.resize(r.is_ok() ? r.value() : 0);
    out.shrink_to_fit();
    return out;
}

// block::compress_prepend_size_with_dict (compress.rs:692)
inline std::vector<uint8_t> compress_prepend_size_with_dict(const uint8_t *input, size_t n, const uint8_t *ext_dict,
                                                            size_t dict_len, lz4b200_ctx *ctx = nullptr)
{
    if (!ctx) ctx = default_context();
    std::vector<uint8_t> out(get_maximum_output_size(n) + 4);
    size_t written = 0;
    if (!ctx || lz4b200_compress_prepend_size_with_dict(ctx, input, n, ext_dict, dict_len, out.data(), out.size(),
                                                        &written) != LZ4B200_OK)
        written = 0;
    out.resize(written);
    out.shrink_to_fit();
    return out;
}

// block::decompress_into_with_dict (decompress.rs:462)
inline Result<size_t, DecompressError> decompress_into_with_dict(const uint8_t *input, size_t n, uint8_t *output,
                                                                 size_t cap, const uint8_t *ext_dict, size_t dict_len,
                                                                 lz4b200_ctx *ctx = nullptr)
{
    using R = Result<size_t, DecompressError>;
    if (!ctx) ctx = default_context();
    if (!ctx) return R::Err({DecompressError::Cuda});
    size_t written = 0, expected = 0, actual = 0;
    const lz4b200_status st = lz4b200_decompress_into_with_dict(ctx, input, n, ext_dict, dict_len, output, cap, &written,
                                                                &expected, &actual);
    if (st == LZ4B200_OK) return R::Ok(written);
    DecompressError e;
    e.kind = (st >= 2 && st <= 6) ? (DecompressError::Kind)st : DecompressError::Cuda;
    e.expected = expected; e.actual = actual;
    return R::Err(e);
}

// block::decompress_with_dict (decompress.rs:478)
inline Result<std::vector<uint8_t>, DecompressError> decompress_with_dict(const uint8_t *input, size_t n,
                                                                          size_t min_uncompressed_size,
                                                                          const uint8_t *ext_dict, size_t dict_len,
                                                                          lz4b200_ctx *ctx = nullptr)
{
    using R = Result<std::vector<uint8_t>, DecompressError>;
    std::vector<uint8_t> out(min_uncompressed_size);
    auto r = decompress_into_with_dict(input, n, out.data(), out.size(), ext_dict, dict_len, ctx);
    if (r.is_err()) return R::Err(r.error());
    out.resize(r.value());
    return R::Ok(std::move(out));
}

// block::decompress_size_prepended_with_dict (decompress.rs:522)
inline Result<std::vector<uint8_t>, DecompressError> decompress_size_prepended_with_dict(const uint8_t *input, size_t n,
                                                                                         const uint8_t *ext_dict,
                                                                                         size_t dict_len,
                                                                                         lz4b200_ctx *ctx = nullptr)
{
    using R = Result<std::vector<uint8_t>, DecompressError>;
    auto s = uncompressed_size(input, n);
    if (s.is_err()) return R::Err(s.error());
    return decompress_with_dict(s.value().second, n - 4, s.value().first, ext_dict, dict_len, ctx);
}

}  // namespace block

namespace frame {

// frame::BlockSize (header.rs:39-53)
enum class BlockSize : int { Auto = 0, Max64KB = 4, Max256KB = 5, Max1MB = 6, Max4MB = 7, Max8MB = 8 };
// frame::BlockMode (header.rs:83-91)
enum class BlockMode : int { Independent = 0, Linked = 1 };

inline size_t block_size_bytes(BlockSize b)
{
    switch (b) {
    case BlockSize::Max64KB: return 64u << 10;
    case BlockSize::Max256KB: return 256u << 10;
    case BlockSize::Max1MB: return 1u << 20;
    case BlockSize::Max4MB: return 4u << 20;
    case BlockSize::Max8MB: return 8u << 20;
    default: return 0;
    }
}

// BlockSize::from_buf_length (header.rs:57-67)
inline BlockSize block_size_from_buf_length(size_t buf_len)
{
    if (buf_len > (256u << 10)) return BlockSize::Max4MB;
    if (buf_len > (64u << 10)) return BlockSize::Max256KB;
    return BlockSize::Max64KB;
}

// frame::FrameInfo (header.rs:130-192), builder style like the crate
struct FrameInfo {
    bool has_content_size = false;
    uint64_t content_size_value = 0;
    BlockSize block_size_value = BlockSize::Auto;
    BlockMode block_mode_value = BlockMode::Independent;
    bool block_checksums_value = false;
    bool content_checksum_value = false;
    FrameInfo &content_size(uint64_t v) { has_content_size = true; content_size_value = v; return *this; }
    FrameInfo &block_size(BlockSize b) { block_size_value = b; return *this; }
    FrameInfo &block_mode(BlockMode m) { block_mode_value = m; return *this; }
    FrameInfo &block_checksums(bool v) { block_checksums_value = v; return *this; }
    FrameInfo &content_checksum(bool v) { content_checksum_value = v; return *this; }
    lz4b200_frame_info to_c() const
    {
        lz4b200_frame_info c{};
        c.block_size_id = (int)block_size_value; c.block_checksums = block_checksums_value;
        c.content_checksum = content_checksum_value; c.has_content_size = has_content_size;
        c.content_size = content_size_value; c.linked = block_mode_value == BlockMode::Linked;
        return c;
    }
};

// frame::Error (frame/mod.rs:35-72): the lz4b200 frame status plus the inner block error
struct Error {
    lz4b200_status status = LZ4B200_OK;
    block::DecompressError decompression_error;     // for Error::DecompressionError
    uint64_t expected = 0, actual = 0;              // for Error::ContentLengthError
    std::string to_string() const { return lz4b200_status_string((int)status); }
};

// frame::FrameEncoder<W> (compress.rs:62-404).  W needs `void write(const uint8_t*, size_t)`.
// Blocks are cut exactly where the reference cuts them and queued; they are compressed on the GPU in batches
// (at flush()/finish() or when `batch_bytes` of input are pending), so `w` receives the reference's bytes.
template <typename W> class FrameEncoder {
public:
    explicit FrameEncoder(W w, FrameInfo info = FrameInfo(), lz4b200_ctx *ctx = nullptr, size_t batch_bytes = 256u << 20)
        : w_(std::move(w)), info_(info), ctx_(ctx), batch_bytes_(batch_bytes) {}
    static FrameEncoder with_frame_info(FrameInfo info, W w) { return FrameEncoder(std::move(w), info); }

    const FrameInfo &frame_info() const { return info_; }
    W &get_mut() { return w_; }
    const W &get_ref() const { return w_; }

    // io::Write::write (compress.rs:375-397)
    Result<size_t, Error> write(const uint8_t *buf, size_t len)
    {
        using R = Result<size_t, Error>;
        if (info_.block_mode_value == BlockMode::Linked) return R::Err({LZ4B200_FRAME_LINKED_UNSUPPORTED});
        if (!frame_open_ && len) begin_frame(len);
        const size_t bs = block_size_bytes(info_.block_size_value);
        size_t pos = 0;
        while (pos < len) {
            const size_t room = bs - src_.size();
            if (room == 0) {
                auto r = write_block();
                if (r.is_err()) return R::Err(r.error());
                continue;
            }
            const size_t take = room < len - pos ? room : len - pos;
            src_.insert(src_.end(), buf + pos, buf + pos + take);
            pos += take;
        }
        return R::Ok(len);
    }

    // io::Write::flush (compress.rs:399-404)
    Result<size_t, Error> flush()
    {
        if (!src_.empty()) { auto r = write_block(); if (r.is_err()) return r; }
        return drain();
    }

    // FrameEncoder::try_finish (compress.rs:166-181)
    Result<size_t, Error> try_finish()
    {
        using R = Result<size_t, Error>;
        auto r = flush();
        if (r.is_err()) return r;
        if (!frame_open_ && !data_written_) begin_frame(0);
        frame_open_ = false;                                                        // end_frame: compress.rs:209-230
        if (info_.has_content_size && info_.content_size_value != content_len_) {
            Error e{LZ4B200_FRAME_CONTENT_LENGTH}; e.expected = info_.content_size_value; e.actual = content_len_;
            return R::Err(e);
        }
        uint8_t tail[8] = {0, 0, 0, 0};
        size_t n = 4;
        if (info_.content_checksum_value) { const uint32_t h = lz4b200_xxh32_digest(&hasher_); memcpy(tail + 4, &h, 4); n = 8; }
        w_.write(tail, n);
        data_written_ = true;
        return R::Ok(0);
    }

    // FrameEncoder::finish (compress.rs:160-163)
    Result<W, Error> finish() &&
    {
        auto r = try_finish();
        if (r.is_err()) return Result<W, Error>::Err(r.error());
        return Result<W, Error>::Ok(std::move(w_));
    }

private:
    void begin_frame(size_t buf_len)                                                // compress.rs:234-258
    {
        frame_open_ = true;
        if (info_.block_size_value == BlockSize::Auto) info_.block_size_value = block_size_from_buf_length(buf_len);
        uint8_t hdr[19];
        const lz4b200_frame_info c = info_.to_c();
        const size_t n = lz4b200_frame_write_header(&c, hdr, sizeof hdr);
        w_.write(hdr, n);
        if (content_len_ != 0) { content_len_ = 0; stream_offset_ = 0; src_.clear(); }
        lz4b200_xxh32_reset(&hasher_, 0);
    }

    Result<size_t, Error> write_block()                                             // compress.rs:261-371 (compression queued)
    {
        const size_t bs = block_size_bytes(info_.block_size_value);
        if (stream_offset_ + bs + block::WINDOW_SIZE >= 0xFFFFFFFFull / 2) stream_offset_ = 0;   // compress.rs:266-271
        flags_.push_back((uint8_t)(LZ4B200_BLOCK_HASH5_ALWAYS | (stream_offset_ ? LZ4B200_BLOCK_CONT : LZ4B200_BLOCK_FRESH)));
        offs_.push_back(pending_.size());
        lens_.push_back((uint32_t)src_.size());
        pending_.insert(pending_.end(), src_.begin(), src_.end());
        if (info_.content_checksum_value) lz4b200_xxh32_update(&hasher_, src_.data(), src_.size());
        content_len_ += src_.size();
        stream_offset_ += src_.size();
        src_.clear();
        if (pending_.size() >= batch_bytes_) return drain();
        return Result<size_t, Error>::Ok(0);
    }

    Result<size_t, Error> drain()
    {
        using R = Result<size_t, Error>;
        const size_t nb = lens_.size();
        if (!nb) return R::Ok(0);
        lz4b200_ctx *ctx = ctx_ ? ctx_ : default_context();
        if (!ctx) return R::Err({LZ4B200_CUDA_ERROR});
        size_t cap = 0;
        for (uint32_t l : lens_) cap += lz4b200_max_output_size(l);
        std::vector<uint8_t> comp(cap);
        std::vector<uint64_t> ooff(nb);
        std::vector<uint32_t> olen(nb);
        std::vector<int32_t> st(nb);
        const lz4b200_status rc = lz4b200_compress_batch_host(ctx, pending_.data(), offs_.data(), lens_.data(), flags_.data(),
                                                              comp.data(), cap, ooff.data(), olen.data(), st.data(), nb);
        if (rc != LZ4B200_OK) return R::Err({rc});
        std::vector<uint8_t> out;
        out.reserve(pending_.size() / 2 + nb * 8);
        for (size_t b = 0; b < nb; b++) {
            const bool stored = !(olen[b] < lens_[b]);                              // compress.rs:301-306
            const uint32_t word = stored ? (lens_[b] | 0x80000000u) : olen[b];
            const uint8_t *payload = stored ? pending_.data() + offs_[b] : comp.data() + ooff[b];
            const size_t plen = stored ? lens_[b] : olen[b];
            const size_t at = out.size();
            out.resize(at + 4 + plen + (info_.block_checksums_value ? 4 : 0));
            memcpy(out.data() + at, &word, 4);
            memcpy(out.data() + at + 4, payload, plen);
            if (info_.block_checksums_value) { const uint32_t h = lz4b200_xxh32(payload, plen, 0); memcpy(out.data() + at + 4 + plen, &h, 4); }
        }
        w_.write(out.data(), out.size());
        pending_.clear(); offs_.clear(); lens_.clear(); flags_.clear();
        return R::Ok(out.size());
    }

    W w_;
    FrameInfo info_;
    lz4b200_ctx *ctx_;
    size_t batch_bytes_;
    std::vector<uint8_t> src_, pending_;
    std::vector<uint64_t> offs_;
    std::vector<uint32_t> lens_;
    std::vector<uint8_t> flags_;
    uint64_t stream_offset_ = 0, content_len_ = 0;
    lz4b200_xxh32_state hasher_{};
    bool frame_open_ = false, data_written_ = false;
};

// frame::FrameDecoder<R> (decompress.rs:48-422) over an in-memory source: decodes every concatenated frame in one
// GPU batch at the first read; bytes before a corrupt block are delivered before the error surfaces.
// frame::FrameDecoder<R> over an in-memory stream (decompress.rs:48-422).  Like the reference, read() returns 0 at
// every EndMark and read_to_end() returns the rest of the CURRENT frame: a second read_to_end() continues with the next
// concatenated frame (tests/tests.rs:633-647).  One frame at a time is decoded (lz4b200_frame_decompress_next), so the
// memory held is one frame's output, and the device memory behind it is bounded by the context's frame budget.
class FrameDecoder {
public:
    FrameDecoder(const uint8_t *data, size_t n, lz4b200_ctx *ctx = nullptr) : data_(data), n_(n), ctx_(ctx) {}

    // io::Read::read
    Result<size_t, Error> read(uint8_t *buf, size_t len)
    {
        using R = Result<size_t, Error>;
        if (pos_ >= out_.size() && !at_frame_end_) next_frame();
        if (pos_ >= out_.size()) {
            at_frame_end_ = false;                             // the Ok(0) of this EndMark / the end of the data
            if (err_.status != LZ4B200_OK) { Error e = err_; err_ = Error(); return R::Err(e); }
            return R::Ok(0);
        }
        const size_t take = len < out_.size() - pos_ ? len : out_.size() - pos_;
        memcpy(buf, out_.data() + pos_, take);
        pos_ += take;
        return R::Ok(take);
    }

    // read_to_end: the rest of the current frame (bytes decoded before an error are lost to the caller, as with the
    // reference's Vec on Err — use read() to drain them first)
    Result<std::vector<uint8_t>, Error> read_to_end()
    {
        using R = Result<std::vector<uint8_t>, Error>;
        if (pos_ >= out_.size() && !at_frame_end_) next_frame();
        at_frame_end_ = false;
        if (err_.status != LZ4B200_OK) { Error e = err_; err_ = Error(); pos_ = out_.size(); return R::Err(e); }
        std::vector<uint8_t> rest(out_.begin() + (long)pos_, out_.end());
        pos_ = out_.size();
        return R::Ok(std::move(rest));
    }

private:
    void next_frame()
    {
        out_.clear(); pos_ = 0;
        if (ip_ >= n_) return;
        lz4b200_ctx *ctx = ctx_ ? ctx_ : default_context();
        if (!ctx) { err_.status = LZ4B200_CUDA_ERROR; return; }
        size_t bound = 0, written = 0, used = 0;
        lz4b200_frame_decoded_bound(data_ + ip_, n_ - ip_, &bound);          // all remaining frames: an upper bound
        out_.resize(bound ? bound : 1);
        int block_status = 0;
        uint64_t e1 = 0, e2 = 0;
        const lz4b200_status st = lz4b200_frame_decompress_next(ctx, data_ + ip_, n_ - ip_, out_.data(), bound, &used, &written,
                                                                &block_status, &e1, &e2);
        out_.resize(written);
        ip_ = (st == LZ4B200_OK && used) ? ip_ + used : n_;                  // an error ends the stream
        at_frame_end_ = written != 0;
        if (st != LZ4B200_OK) {
            err_.status = st;
            err_.expected = e1; err_.actual = e2;
            if (st == LZ4B200_FRAME_DECOMPRESSION_ERROR) {
                err_.decompression_error.kind = (block::DecompressError::Kind)block_status;
                err_.decompression_error.expected = e1; err_.decompression_error.actual = e2;
            }
        }
    }
    const uint8_t *data_;
    size_t n_;
    lz4b200_ctx *ctx_;
    std::vector<uint8_t> out_;
    size_t pos_ = 0, ip_ = 0;
    bool at_frame_end_ = false;
    Error err_;
};

}  // namespace frame
}  // namespace lz4_flex
