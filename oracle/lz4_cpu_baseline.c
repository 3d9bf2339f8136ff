/*
 * lz4_cpu_baseline.c — the CPU arm of bench.py: a performance-oriented restatement of lz4_flex's *unsafe* block
 * path plus a persistent worker pool.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (same rule as lz4_oracle.c: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it; the product never links it).
 *
 * Why a second file: lz4_oracle.c is the parity checker and stays as plain as possible.  The numbers bench.py prints
 * beside the GPU must be a fair stand-in for the crate's unsafe build, which the README puts ahead of C lz4
 * (README.md:17-30), so this file spends effort where the crate does:
 *   - one 8-byte load per probe serves hash and 4-byte comparison (compress.rs:386-438, get_batch_arch),
 *   - literals are copied with 16-byte wild copies (copy_literals_wild, compress.rs:523-545),
 *   - the decoder's hot loop copies 16 literal bytes and 18 match bytes unconditionally when the token has no
 *     length extension and 8/16-byte chunks otherwise (decompress.rs:259-328, wild_copy_from_src_16 :38-52,
 *     duplicate :11-35, duplicate_overlapping :57-82).
 * Same parse, same bytes: tests/test_oracle.py::test_fast_baseline_equals_oracle compares both directions with
 * lz4_oracle.c on the corpus, edge lengths and garbage streams.
 *
 * The pool keeps its threads between calls (condition variable hand-off), hands out blocks in chunks from an
 * atomic counter, and gives every worker its own 16 KiB table — a timed step measures compression, not
 * pthread_create.
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lz4_oracle.h"

static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void cp8(uint8_t *d, const uint8_t *s) { memcpy(d, s, 8); }
static inline void cp16(uint8_t *d, const uint8_t *s) { memcpy(d, s, 16); }

static inline uint8_t *put_ext(uint8_t *op, size_t v)
{
    while (v >= 255) { *op++ = 255; v -= 255; }
    *op++ = (uint8_t)v;
    return op;
}

/* compress_internal, no dictionary; tab: 4096 u32 entries, zeroed here (FRESH block API semantics; the 4-byte hash
 * for n < 65535, compress.rs:559).  Returns the compressed size or -1 (OutputTooSmall, compress.rs:338-340). */
int64_t lz4cpu_compress_block(const uint8_t *in, size_t n, uint8_t *out, size_t cap, uint32_t *tab)
{
    if (cap < lz4o_max_output_size(n)) return -1;
    uint8_t *op = out;
    size_t anchor = 0;
    if (n >= 13) {
        const int h5 = n >= 65535;
        const size_t last_probe = n - 12, lim = n - 6;
        memset(tab, 0, 4096 * sizeof(uint32_t));
        size_t cur = 1;
        {
            const uint64_t v = ld64(in);
            tab[h5 ? (uint32_t)(((v << 24) * 889523592379ULL) >> 52) : (((uint32_t)v * 2654435761u) >> 20)] = 0;
        }
        for (;;) {
            size_t misses = 32, next = cur, cand;
            for (;;) {
                const size_t step = misses >> 5;
                misses++;
                cur = next; next += step;
                if (cur > last_probe) goto tail;
                const uint64_t v = ld64(in + cur);
                const uint32_t s = h5 ? (uint32_t)(((v << 24) * 889523592379ULL) >> 52) : (((uint32_t)v * 2654435761u) >> 20);
                cand = tab[s];
                tab[s] = (uint32_t)cur;
                if (cur - cand > 65535) continue;
                if (ld32(in + cand) == (uint32_t)v) break;
            }
            while (cand > 0 && cur > anchor && in[cur - 1] == in[cand - 1]) { cur--; cand--; }
            const size_t lit = cur - anchor, dist = cur - cand;
            cur += 4; cand += 4;
            {
                const size_t start = cur;
                while (cur + 8 <= lim) {
                    const uint64_t x = ld64(in + cur) ^ ld64(in + cand);
                    if (x) { cur += (size_t)(__builtin_ctzll(x) >> 3); goto counted; }
                    cur += 8; cand += 8;
                }
                while (cur < lim && in[cur] == in[cand]) { cur++; cand++; }
            counted:;
                const size_t extra = cur - start;
                {
                    const uint64_t v2 = ld64(in + cur - 2);
                    tab[h5 ? (uint32_t)(((v2 << 24) * 889523592379ULL) >> 52) : (((uint32_t)v2 * 2654435761u) >> 20)] = (uint32_t)(cur - 2);
                }
                *op++ = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (extra < 15 ? extra : 15));
                if (lit >= 15) op = put_ext(op, lit - 15);
                if (lit <= 16 && anchor + 16 <= n) {             /* copy_literals_wild */
                    cp16(op, in + anchor);
                } else if (anchor + lit + 16 <= n) {
                    for (size_t i = 0; i < lit; i += 16) cp16(op + i, in + anchor + i);
                } else {
                    memcpy(op, in + anchor, lit);
                }
                op += lit;
                op[0] = (uint8_t)dist; op[1] = (uint8_t)(dist >> 8);
                op += 2;
                if (extra >= 15) op = put_ext(op, extra - 15);
            }
            anchor = cur;
        }
    }
tail:;
    {
        const size_t lit = n - anchor;
        *op++ = (uint8_t)((lit < 15 ? lit : 15) << 4);
        if (lit >= 15) op = put_ext(op, lit - 15);
        memcpy(op, in + anchor, lit);
        op += lit;
    }
    return op - out;
}

/* decompress_internal without dictionary: same results as lz4o_decompress_block. */
int lz4cpu_decompress_block(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written,
                            size_t *err_expected, size_t *err_actual)
{
    size_t ip = 0, op = 0;
    *written = 0;
    if (n == 0) return LZ4O_ERR_EXPECTED_ANOTHER_BYTE;
    for (;;) {
        const uint8_t tok = in[ip++];
        size_t lit = tok >> 4, mlen = (size_t)(tok & 15) + 4;
        /* literals: up to 14 of them are one unconditional 16-byte copy when both buffers have the room (decompress.rs:
         * 259-275); the room also rules out every literal-side error and the end of the block */
        if (__builtin_expect(lit != 15 && n - ip >= 16 && cap - op >= 16, 1)) {
            cp16(out + op, in + ip);
            ip += lit; op += lit;
        } else {
            if (lit) {
                if (lit == 15) {
                    for (;;) {
                        if (ip >= n) return LZ4O_ERR_EXPECTED_ANOTHER_BYTE;
                        const uint8_t b = in[ip++];
                        lit += b;
                        if (b != 255) break;
                    }
                }
                if (lit > n - ip) return LZ4O_ERR_LITERAL_OOB;
                if (lit > cap - op) { *err_expected = op + lit; *err_actual = cap; return LZ4O_ERR_OUTPUT_TOO_SMALL; }
                if (n - ip >= lit + 16 && cap - op >= lit + 16) {
                    for (size_t i = 0; i < lit; i += 16) cp16(out + op + i, in + ip + i);
                } else {
                    memcpy(out + op, in + ip, lit);
                }
                ip += lit; op += lit;
            }
            if (ip >= n) break;
            if (n - ip < 2) return LZ4O_ERR_EXPECTED_ANOTHER_BYTE;
        }
        const size_t dist = (size_t)in[ip] | ((size_t)in[ip + 1] << 8);
        ip += 2;
        if (dist == 0) return LZ4O_ERR_OFFSET_ZERO;
        if (mlen == 19) {
            for (;;) {
                if (ip >= n) return LZ4O_ERR_EXPECTED_ANOTHER_BYTE;
                const uint8_t b = in[ip++];
                mlen += b;
                if (b != 255) break;
            }
        }
        if (dist > op) return LZ4O_ERR_OFFSET_OOB;
        if (mlen > cap - op) { *err_expected = op + mlen; *err_actual = cap; return LZ4O_ERR_OUTPUT_TOO_SMALL; }
        {
            uint8_t *to = out + op;
            const uint8_t *from = to - dist;
            if (__builtin_expect(dist >= 16 && cap - op >= mlen + 16, 1)) {      /* duplicate: 16-byte chunks (decompress.rs:11-35) */
                cp16(to, from);
                if (mlen > 16) { cp16(to + 16, from + 16); for (size_t i = 32; i < mlen; i += 16) cp16(to + i, from + i); }
            } else if (dist >= 8 && cap - op >= mlen + 8) {
                for (size_t i = 0; i < mlen; i += 8) cp8(to + i, from + i);
            } else if (dist == 1) {
                memset(to, from[0], mlen);
            } else if (dist >= mlen) {
                memcpy(to, from, mlen);
            } else {
                for (size_t i = 0; i < mlen; i++) to[i] = from[i];               /* duplicate_overlapping: decompress.rs:57-82 */
            }
        }
        op += mlen;
        if (ip >= n) return LZ4O_ERR_EXPECTED_ANOTHER_BYTE;
    }
    *written = op;
    return LZ4O_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* persistent pool                                                                             */

typedef struct lz4cpu_pool lz4cpu_pool;
struct lz4cpu_pool {
    int nthreads;
    pthread_t *th;
    pthread_mutex_t mu;
    pthread_cond_t cv_start, cv_done;
    uint64_t generation;
    int running, stop;
    /* the job */
    int decode;                 /* 0 compress, 1 decompress, 2 copy (first touch of a buffer by the workers) */
    uint8_t *cp_dst; const uint8_t *cp_src; size_t cp_bytes;
    const uint8_t *in; const uint64_t *in_off; const uint32_t *in_len;
    uint8_t *out; const uint64_t *out_off; const uint32_t *out_cap; uint32_t *out_len; int32_t *status;
    size_t nblocks;
    atomic_size_t next;
};

#define POOL_CHUNK 8

#define COPY_PIECE ((size_t)1 << 20)

static void pool_work(lz4cpu_pool *p, uint32_t *tab)
{
    if (p->decode == 2) {                                   /* parallel copy: pages land on the nodes of the threads that use them */
        for (;;) {
            const size_t i = atomic_fetch_add_explicit(&p->next, 1, memory_order_relaxed);
            if (i * COPY_PIECE >= p->cp_bytes) break;
            const size_t len = p->cp_bytes - i * COPY_PIECE < COPY_PIECE ? p->cp_bytes - i * COPY_PIECE : COPY_PIECE;
            if (p->cp_src) memcpy(p->cp_dst + i * COPY_PIECE, p->cp_src + i * COPY_PIECE, len);
            else memset(p->cp_dst + i * COPY_PIECE, 0, len);
        }
        return;
    }
    for (;;) {
        const size_t b0 = atomic_fetch_add_explicit(&p->next, POOL_CHUNK, memory_order_relaxed);
        if (b0 >= p->nblocks) break;
        const size_t b1 = b0 + POOL_CHUNK < p->nblocks ? b0 + POOL_CHUNK : p->nblocks;
        for (size_t b = b0; b < b1; b++) {
            const uint8_t *src = p->in + p->in_off[b];
            uint8_t *dst = p->out + p->out_off[b];
            if (p->decode) {
                size_t w, e1, e2;
                p->status[b] = lz4cpu_decompress_block(src, p->in_len[b], dst, p->out_cap[b], &w, &e1, &e2);
                p->out_len[b] = (uint32_t)w;
            } else {
                const int64_t r = lz4cpu_compress_block(src, p->in_len[b], dst, p->out_cap[b], tab);
                p->status[b] = r < 0 ? LZ4O_ERR_COMPRESS_OUTPUT_TOO_SMALL : LZ4O_OK;
                p->out_len[b] = r < 0 ? 0 : (uint32_t)r;
            }
        }
    }
}

static void *pool_thread(void *arg)
{
    lz4cpu_pool *p = (lz4cpu_pool *)arg;
    uint32_t *tab = (uint32_t *)malloc(4096 * sizeof(uint32_t));
    uint64_t seen = 0;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (!p->stop && p->generation == seen) pthread_cond_wait(&p->cv_start, &p->mu);
        if (p->stop) { pthread_mutex_unlock(&p->mu); break; }
        seen = p->generation;
        pthread_mutex_unlock(&p->mu);
        pool_work(p, tab);
        pthread_mutex_lock(&p->mu);
        if (--p->running == 0) pthread_cond_signal(&p->cv_done);
        pthread_mutex_unlock(&p->mu);
    }
    free(tab);
    return NULL;
}

lz4cpu_pool *lz4cpu_pool_create(int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    lz4cpu_pool *p = (lz4cpu_pool *)calloc(1, sizeof *p);
    p->nthreads = nthreads;
    p->th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    pthread_mutex_init(&p->mu, NULL);
    pthread_cond_init(&p->cv_start, NULL);
    pthread_cond_init(&p->cv_done, NULL);
    for (int t = 0; t < nthreads; t++) pthread_create(&p->th[t], NULL, pool_thread, p);
    return p;
}

void lz4cpu_pool_destroy(lz4cpu_pool *p)
{
    if (!p) return;
    pthread_mutex_lock(&p->mu);
    p->stop = 1;
    pthread_cond_broadcast(&p->cv_start);
    pthread_mutex_unlock(&p->mu);
    for (int t = 0; t < p->nthreads; t++) pthread_join(p->th[t], NULL);
    pthread_mutex_destroy(&p->mu);
    pthread_cond_destroy(&p->cv_start);
    pthread_cond_destroy(&p->cv_done);
    free(p->th);
    free(p);
}

/* One batch on the pool's threads; returns when every block is done. */
void lz4cpu_pool_run(lz4cpu_pool *p, int decode, const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                     uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len,
                     int32_t *status, size_t nblocks)
{
    pthread_mutex_lock(&p->mu);
    p->decode = decode;
    p->in = in; p->in_off = in_off; p->in_len = in_len;
    p->out = out; p->out_off = out_off; p->out_cap = out_cap; p->out_len = out_len; p->status = status;
    p->nblocks = nblocks;
    atomic_store(&p->next, 0);
    p->running = p->nthreads;
    p->generation++;
    pthread_cond_broadcast(&p->cv_start);
    while (p->running) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
}

/* dst[0..n) = src[0..n) (src == NULL: zero fill), 1 MiB pieces handed to the pool's threads: a buffer the workers are going
 * to stream through is first touched BY the workers, so its pages spread over the NUMA nodes instead of all sitting on the
 * node of the thread that allocated it. */
void lz4cpu_pool_copy(lz4cpu_pool *p, uint8_t *dst, const uint8_t *src, size_t n)
{
    pthread_mutex_lock(&p->mu);
    p->decode = 2;
    p->cp_dst = dst; p->cp_src = src; p->cp_bytes = n;
    atomic_store(&p->next, 0);
    p->running = p->nthreads;
    p->generation++;
    pthread_cond_broadcast(&p->cv_start);
    while (p->running) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
}
