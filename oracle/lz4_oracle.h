/*
 * lz4_oracle.h — interface of the CPU parity oracle (TEST INFRASTRUCTURE ONLY; see lz4_oracle.c).
 */
#ifndef LZ4_ORACLE_H
#define LZ4_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Block status codes: same numbering as include/lz4b200.h (DecompressError / CompressError
 * of src/block/mod.rs:82-106). */
enum {
    LZ4O_OK = 0,
    LZ4O_ERR_COMPRESS_OUTPUT_TOO_SMALL = 1,
    LZ4O_ERR_OUTPUT_TOO_SMALL = 2,
    LZ4O_ERR_LITERAL_OOB = 3,
    LZ4O_ERR_EXPECTED_ANOTHER_BYTE = 4,
    LZ4O_ERR_OFFSET_ZERO = 5,
    LZ4O_ERR_OFFSET_OOB = 6
};

/* Frame errors (src/frame/mod.rs:35-72), offset so they never collide with block codes. */
enum {
    LZ4O_FERR_DECOMPRESSION = 101,
    LZ4O_FERR_WRONG_MAGIC = 102,
    LZ4O_FERR_RESERVED_BITS = 103,
    LZ4O_FERR_UNSUPPORTED_VERSION = 104,
    LZ4O_FERR_UNSUPPORTED_BLOCKSIZE = 105,
    LZ4O_FERR_HEADER_CHECKSUM = 106,
    LZ4O_FERR_BLOCK_CHECKSUM = 107,
    LZ4O_FERR_CONTENT_CHECKSUM = 108,
    LZ4O_FERR_CONTENT_LENGTH = 109,
    LZ4O_FERR_BLOCK_TOO_BIG = 110,
    LZ4O_FERR_SKIPPABLE = 111,
    LZ4O_FERR_DICTIONARY = 112,
    LZ4O_FERR_IO_EOF = 113,
    LZ4O_FERR_LINKED_UNSUPPORTED = 114,
    LZ4O_FERR_OUTPUT_FULL = 115
};

/* FrameInfo flags */
enum {
    LZ4O_F_BLOCK_CHECKSUMS = 1,
    LZ4O_F_CONTENT_CHECKSUM = 2,
    LZ4O_F_CONTENT_SIZE = 4,
    LZ4O_F_LINKED = 8
};

size_t  lz4o_max_output_size(size_t n);
int64_t lz4o_compress_block(const uint8_t *in, size_t n, uint8_t *out, size_t cap);
int64_t lz4o_compress_block_with_table(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                                       uint32_t *table4096, uint64_t stream_offset);
int64_t lz4o_compress_block_dict(const uint8_t *in, size_t n, const uint8_t *dict, size_t dlen, uint8_t *out,
                                 size_t cap);
int     lz4o_decompress_block_dict(const uint8_t *in, size_t n, const uint8_t *dict, size_t dlen, uint8_t *out,
                                   size_t cap, size_t *written, size_t *err_expected, size_t *err_actual);
int64_t lz4o_compress_prepend_size(const uint8_t *in, size_t n, uint8_t *out, size_t cap);
int     lz4o_decompress_block(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                              size_t *written, size_t *err_expected, size_t *err_actual);
int     lz4o_decompress_size_prepended(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                                       size_t *written, size_t *err_expected, size_t *err_actual);
uint32_t lz4o_xxh32(const uint8_t *p, size_t n, uint32_t seed);

size_t  lz4o_block_size_bytes(int id);
int     lz4o_auto_block_size_id(size_t first_write_len);
size_t  lz4o_frame_header(uint8_t *out, int block_size_id, unsigned flags, uint64_t content_size);
size_t  lz4o_frame_bound(size_t n, int block_size_id);
int64_t lz4o_frame_compress(const uint8_t *in, size_t n, int block_size_id, unsigned flags,
                            size_t flush_every, uint8_t *out, size_t cap);
int     lz4o_frame_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                              size_t *written, int *block_err);

void lz4o_compress_batch(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                         uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                         uint32_t *out_len, int32_t *status, size_t nblocks, int nthreads);
void lz4o_decompress_batch(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                           uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                           uint32_t *out_len, int32_t *status, size_t nblocks, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
