/* examples/roundtrip.c -- the C ABI of include/zeekstd_b200.h used from plain C, the way a binding in the reference's language would:
 * Encoder (write-sink callback) -> seekable archive in memory -> Decoder over a Seekable (callbacks) -> ranged read.
 * The calls mirror the reference's own usage (lib/src/lib.rs:20-56 doc example; encode.rs:626-775; decode.rs:152-270, 402-437).
 *
 *   gcc -std=c99 -Iinclude examples/roundtrip.c zeekstd_b200/libzeekstd_b200.so -o roundtrip && ./roundtrip      (needs a GPU)
 *
 * tests/test_abi.py builds it against the product library (link check) and RUNS it against the CPU emulation build of the same sources. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "zeekstd_b200.h"

typedef struct { uint8_t* p; size_t len, cap, pos; } membuf;

static int32_t sink_write(void* user, const uint8_t* data, size_t len) {            /* W: std::io::Write */
    membuf* m = (membuf*)user;
    if (m->len + len > m->cap) return -1;
    memcpy(m->p + m->len, data, len); m->len += len;
    return 0;
}
static int32_t sink_flush(void* user) { (void)user; return 0; }

static int64_t src_set_offset(void* user, int32_t whence, int64_t off) {            /* trait Seekable (seekable.rs:16-39) */
    membuf* m = (membuf*)user;
    int64_t p = whence == 0 ? off : (int64_t)m->len + off;
    if (p < 0 || p > (int64_t)m->len) return -1;
    m->pos = (size_t)p;
    return p;
}
static int64_t src_read(void* user, uint8_t* buf, size_t len) {
    membuf* m = (membuf*)user;
    size_t n = m->len - m->pos < len ? m->len - m->pos : len;
    memcpy(buf, m->p + m->pos, n); m->pos += n;
    return (int64_t)n;
}

#define CHECK(x) do { int32_t rc__ = (x); if (rc__ != 0) { fprintf(stderr, "%s -> %d (%s)\n", #x, (int)rc__, zk_error_name(rc__)); return 1; } } while (0)

int main(void) {
    const size_t n = 100000; const uint32_t frame_size = 16000;
    uint8_t* plain = (uint8_t*)malloc(n);
    for (size_t i = 0; i < n; i++) plain[i] = (uint8_t)("the quick brown fox jumps over the lazy dog. "[(i * 7 + i / 97) % 45]);
    membuf arc = { (uint8_t*)malloc(2 * n + 4096), 0, 2 * n + 4096, 0 };

    zk_ctx* ctx = NULL;
    CHECK(zk_ctx_create(0, 0, &ctx));

    /* EncodeOptions::new().frame_size_policy(Uncompressed(16000)).checksum_flag(true).compression_level(3).into_encoder(writer) */
    zk_encode_options* eo = zk_encode_options_new(ctx);
    zk_encode_options_frame_size_policy(eo, ZK_POLICY_UNCOMPRESSED, frame_size);
    zk_encode_options_checksum_flag(eo, 1);
    zk_encode_options_compression_level(eo, 3);
    zk_encoder* enc = NULL;
    CHECK(zk_encode_options_into_encoder(eo, sink_write, sink_flush, &arc, &enc));      /* consumes eo */
    for (size_t pos = 0; pos < n;) {                                                    /* io::copy(&mut input, &mut encoder) in 9999-byte pieces */
        size_t take = n - pos < 9999 ? n - pos : 9999, used = 0;
        CHECK(zk_encoder_compress(enc, plain + pos, take, &used));
        pos += used;
    }
    uint64_t total = 0;
    CHECK(zk_encoder_finish(enc, &total));                                              /* frames + Foot seek table; consumes enc */
    if (total != arc.len) { fprintf(stderr, "finish() says %llu, sink holds %zu\n", (unsigned long long)total, arc.len); return 1; }

    /* Decoder::new(seekable) over callbacks; read [33333, 77777) */
    zk_seekable src = { &arc, src_set_offset, src_read };
    zk_decoder* dec = NULL;
    CHECK(zk_decode_options_into_decoder(zk_decode_options_new_seekable(ctx, src), &dec));
    const zk_seek_table* st = zk_decoder_seek_table(dec);
    uint32_t frames = zk_seek_table_num_frames(st);
    if (frames != (n + frame_size - 1) / frame_size || zk_seek_table_size_decomp(st) != n) { fprintf(stderr, "seek table: %u frames\n", frames); return 1; }
    CHECK(zk_decoder_set_offset(dec, 33333));
    CHECK(zk_decoder_set_offset_limit(dec, 77777));
    uint8_t* out = (uint8_t*)malloc(n);
    size_t got = 0;
    for (;;) {
        size_t k = 0;
        CHECK(zk_decoder_decompress(dec, out + got, 7000, &k));                         /* small buffer on purpose */
        if (k == 0) break;
        got += k;
    }
    if (got != 77777 - 33333 || memcmp(out, plain + 33333, got) != 0) { fprintf(stderr, "ranged read mismatch (%zu bytes)\n", got); return 1; }
    /* an offset past the end is the reference's OffsetOutOfRange */
    if (zk_decoder_set_offset(dec, n + 1) != ZK_ERR_OFFSET_OUT_OF_RANGE) { fprintf(stderr, "expected OffsetOutOfRange\n"); return 1; }
    /* everything */
    zk_decoder_reset(dec);
    got = 0;
    for (;;) { size_t k = 0; CHECK(zk_decoder_decompress(dec, out + got, n - got, &k)); if (k == 0) break; got += k; }
    if (got != n || memcmp(out, plain, n) != 0) { fprintf(stderr, "full read mismatch\n"); return 1; }

    printf("ok: %zu bytes -> %zu in %u frames (+ seek table), ranged and full reads restored; %llu kernel launches\n",
           n, arc.len, frames, (unsigned long long)zk_ctx_kernel_launches(ctx));
    zk_decoder_free(dec);
    zk_ctx_destroy(ctx);
    free(out); free(arc.p); free(plain);
    return 0;
}
