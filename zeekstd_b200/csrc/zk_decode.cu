// zk_decode.cu -- batched Zstandard frame decode for sm_100a.
//
// Replaces the reference's only decompression call into libzstd,
//   lib/src/decode.rs:243-245   self.dctx.decompress_stream(&mut out_buffer, &mut in_buffer)
// with four kernels over a whole batch of independent seek-table entries ("frames" of the
// seekable format, seekable_format.md:23-29):
//
//   K-D0 zk_scan_kernel     one thread / entry : walk frame + block headers, carve scratch
//   K-D1 zk_entropy_kernel  one CTA   / block  : Huffman literals (warp 1) + FSE sequences (warp 0)
//   K-D2 zk_exec_kernel     one CTA   / entry  : ordered sequence execution, W warps pipelined
//   K-D3 zk_xxh64_kernel    one warp  / entry  : content checksum (only if requested & present)
//
// Format rules: RFC 8878 as restated in SURVEY.md Appendix A (the arithmetic is not in the
// reference tree).  All integer/byte work; no tensor cores.
#include "zk_common.cuh"
#include "zk_decode.h"

// =============================================================================================
// K-D0: header scan
// =============================================================================================
struct ZkBlkInfo {
    uint32_t src, size; uint8_t type, flags;
    uint32_t lit_size, nseq; uint8_t lit_type, modes; uint64_t fcs;
};

struct ZkLitHdr { uint32_t type, hdr, regen, comp, streams; };

// Literals_Section_Header (A.3).  n = bytes available. Returns false if truncated.
__device__ __forceinline__ bool zk_parse_lit_hdr(const uint8_t* p, uint32_t n, ZkLitHdr& h) {
    if (n < 1) return false;
    uint32_t b0 = p[0];
    h.type = b0 & 3; uint32_t sf = (b0 >> 2) & 3;
    h.comp = 0; h.streams = 1;
    if (h.type < 2) {
        if (sf == 0 || sf == 2) { h.hdr = 1; h.regen = b0 >> 3; }
        else if (sf == 1) { if (n < 2) return false; h.hdr = 2; h.regen = (b0 >> 4) | ((uint32_t)p[1] << 4); }
        else { if (n < 3) return false; h.hdr = 3; h.regen = (b0 >> 4) | ((uint32_t)p[1] << 4) | ((uint32_t)p[2] << 12); }
    } else {
        if (n < 5) return false;
        unsigned long long v = (unsigned long long)zk_ld_le32(p) | ((unsigned long long)p[4] << 32);
        if (sf == 0) { h.hdr = 3; h.streams = 1; h.regen = (uint32_t)(v >> 4) & 0x3FF; h.comp = (uint32_t)(v >> 14) & 0x3FF; }
        else if (sf == 1) { h.hdr = 3; h.streams = 4; h.regen = (uint32_t)(v >> 4) & 0x3FF; h.comp = (uint32_t)(v >> 14) & 0x3FF; }
        else if (sf == 2) { h.hdr = 4; h.streams = 4; h.regen = (uint32_t)(v >> 4) & 0x3FFF; h.comp = (uint32_t)(v >> 18) & 0x3FFF; }
        else { h.hdr = 5; h.streams = 4; h.regen = (uint32_t)(v >> 4) & 0x3FFFF; h.comp = (uint32_t)(v >> 22) & 0x3FFFF; }
    }
    return true;
}
__device__ __forceinline__ uint32_t zk_lit_section_size(const ZkLitHdr& h) {
    return h.hdr + (h.type == 0 ? h.regen : h.type == 1 ? 1u : h.comp);
}

// Sequences_Section_Header (A.5): number of sequences.  Returns header bytes (1..3) or 0 if truncated.
__device__ __forceinline__ uint32_t zk_parse_nseq(const uint8_t* s, uint32_t n, uint32_t& nseq) {
    if (n < 1) return 0;
    uint32_t b0 = s[0];
    if (b0 < 128) { nseq = b0; return 1; }
    if (b0 < 255) { if (n < 2) return 0; nseq = ((b0 - 128) << 8) + s[1]; return 2; }
    if (n < 3) return 0;
    nseq = (uint32_t)s[1] + ((uint32_t)s[2] << 8) + 0x7F00u; return 3;
}

// Walks every zstd / skippable frame inside one seek-table entry.  emit(info) is called per block.
template <class Emit>
__device__ int zk_walk_entry(const uint8_t* p, uint32_t n, Emit& emit) {
    uint32_t pos = 0;
    while (pos < n) {
        if (n - pos < 4) return ZKZ_SRC_SIZE_WRONG;
        uint32_t magic = zk_ld_le32(p + pos);
        if ((magic & ZK_SKIPPABLE_MASK) == ZK_SKIPPABLE_MAGIC) {
            if (n - pos < 8) return ZKZ_SRC_SIZE_WRONG;
            uint32_t sz = zk_ld_le32(p + pos + 4);
            if ((unsigned long long)sz + 8ull > (unsigned long long)(n - pos)) return ZKZ_SRC_SIZE_WRONG;
            pos += 8 + sz;
            continue;
        }
        if (magic != ZK_MAGIC) return ZKZ_PREFIX_UNKNOWN;
        if (n - pos < 5) return ZKZ_SRC_SIZE_WRONG;
        uint32_t fhd = p[pos + 4];
        uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, csum = (fhd >> 2) & 1, did = fhd & 3;
        if (fhd & 0x08) return ZKZ_FRAMEPARAM_UNSUPPORTED;
        uint32_t did_sz = did == 3 ? 4u : did;
        uint32_t fcs_sz = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2u : (fcs_flag == 2 ? 4u : 8u));
        uint32_t hsz = 5 + (single ? 0 : 1) + did_sz + fcs_sz;
        if (n - pos < hsz) return ZKZ_SRC_SIZE_WRONG;
        uint32_t q = pos + 5;
        if (!single) { uint32_t wd = p[q++]; if (10 + (wd >> 3) > 31) return ZKZ_WINDOW_TOO_LARGE; }
        uint32_t dict = 0;
        for (uint32_t i = 0; i < did_sz; i++) dict |= (uint32_t)p[q + i] << (8 * i);
        q += did_sz;
        if (dict != 0) return ZKZ_DICT_WRONG;
        unsigned long long fcs = 0;
        for (uint32_t i = 0; i < fcs_sz; i++) fcs |= (unsigned long long)p[q + i] << (8 * i);
        if (fcs_sz == 2) fcs += 256;
        pos += hsz;
        bool first = true;
        for (;;) {
            if (n - pos < 3) return ZKZ_SRC_SIZE_WRONG;
            uint32_t bh = zk_ld_le24(p + pos);
            uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
            if (type == 3) return ZKZ_CORRUPTION;
            if (bsize > ZK_BLOCK_MAX) return ZKZ_CORRUPTION;
            uint32_t content = type == 1 ? 1u : bsize;
            if (n - pos - 3 < content) return ZKZ_SRC_SIZE_WRONG;
            ZkBlkInfo bi;
            bi.src = pos + 3; bi.size = bsize; bi.type = (uint8_t)type;
            bi.flags = (uint8_t)((first ? ZKB_FIRST : 0) | (last ? ZKB_LAST : 0) | ((last && csum) ? ZKB_HAS_CSUM : 0) |
                                 ((last && fcs_sz) ? ZKB_HAS_FCS : 0));
            bi.fcs = fcs; bi.lit_size = 0; bi.nseq = 0; bi.lit_type = 0; bi.modes = 0;
            if (type == 2) {
                const uint8_t* b = p + pos + 3;
                if (bsize < 2) return ZKZ_CORRUPTION;
                ZkLitHdr lh;
                if (!zk_parse_lit_hdr(b, bsize, lh)) return ZKZ_CORRUPTION;
                if (lh.regen > ZK_BLOCK_MAX) return ZKZ_CORRUPTION;
                uint32_t lsec = zk_lit_section_size(lh);
                if (lsec >= bsize) return ZKZ_CORRUPTION;          // at least the nseq byte must follow
                uint32_t nseq;
                uint32_t sh = zk_parse_nseq(b + lsec, bsize - lsec, nseq);
                if (sh == 0) return ZKZ_CORRUPTION;
                if (nseq == 0) { if (lsec + sh != bsize) return ZKZ_CORRUPTION; }
                else {
                    if (lsec + sh >= bsize) return ZKZ_CORRUPTION;
                    bi.modes = b[lsec + sh];
                    if (bi.modes & 3) return ZKZ_CORRUPTION;
                }
                bi.lit_size = lh.regen; bi.nseq = nseq; bi.lit_type = (uint8_t)lh.type;
            }
            int rc = emit(bi);
            if (rc) return rc;
            pos += 3 + content;
            first = false;
            if (last) break;
        }
        if (csum) { if (n - pos < 4) return ZKZ_SRC_SIZE_WRONG; pos += 4; }
    }
    return 0;
}

struct ZkCountEmit {
    uint32_t nb = 0, nlit = 0, nseq = 0;
    __device__ int operator()(const ZkBlkInfo& bi) {
        nb++;
        if (bi.type == 2) { if (bi.lit_type >= 2) nlit += (bi.lit_size + 15u) & ~15u; nseq += bi.nseq; }
        return 0;
    }
};

struct ZkFillEmit {
    ZkBlock* blocks; uint32_t entry, bidx, lit, seq;
    int32_t huf_ref = -1, ll_ref = -1, of_ref = -1, ml_ref = -1;
    __device__ int operator()(const ZkBlkInfo& bi) {
        ZkBlock b;
        b.src = bi.src; b.size = bi.size; b.entry = entry; b.type = bi.type; b.flags = bi.flags;
        b.lit_kind = 0; b.lit_byte = 0; b.lit_base = lit; b.seq_base = seq; b.nseq = bi.nseq; b.lit_size = bi.lit_size;
        b.lit_src = 0; b.regen = bi.type == 2 ? 0 : bi.size; b.status = 0;
        b.rep_out[0] = ZK_SYM_MAKE(0, 0); b.rep_out[1] = ZK_SYM_MAKE(1, 0); b.rep_out[2] = ZK_SYM_MAKE(2, 0);
        b.fcs = bi.fcs; b.hash_start = 0; b.hash_len = 0;
        if (bi.flags & ZKB_FIRST) { huf_ref = ll_ref = of_ref = ml_ref = -1; }
        b.huf_ref = -1; b.ll_ref = -1; b.of_ref = -1; b.ml_ref = -1;
        if (bi.type == 2) {
            if (bi.lit_type == 3) { if (huf_ref < 0) return ZKZ_CORRUPTION; b.huf_ref = huf_ref; }
            if (bi.lit_type == 2) huf_ref = (int32_t)bidx;
            if (bi.lit_type >= 2) lit += (bi.lit_size + 15u) & ~15u;
            if (bi.nseq) {
                uint32_t ml_m = (bi.modes >> 2) & 3, of_m = (bi.modes >> 4) & 3, ll_m = (bi.modes >> 6) & 3;
                if (ll_m == 3) { if (ll_ref < 0) return ZKZ_CORRUPTION; b.ll_ref = ll_ref; } else ll_ref = (int32_t)bidx;
                if (of_m == 3) { if (of_ref < 0) return ZKZ_CORRUPTION; b.of_ref = of_ref; } else of_ref = (int32_t)bidx;
                if (ml_m == 3) { if (ml_ref < 0) return ZKZ_CORRUPTION; b.ml_ref = ml_ref; } else ml_ref = (int32_t)bidx;
                seq += bi.nseq;
            }
        }
        blocks[bidx++] = b;
        return 0;
    }
};

__global__ void __launch_bounds__(128) zk_scan_kernel(ZkDecodeArgs a) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n_entries) return;
    ZkEntry ent; ent.first_block = 0; ent.n_blocks = 0; ent.status = 0; ent.produced = 0;
    unsigned long long c0 = a.c_off[e], c1 = a.c_off[e + 1];
    if (c1 < c0 || c1 - c0 > 0xFFFFFFFFull) { ent.status = -ZKZ_SRC_SIZE_WRONG; a.entries[e] = ent; return; }
    const uint8_t* p = a.comp + c0; uint32_t n = (uint32_t)(c1 - c0);
    ZkCountEmit ce;
    int rc = zk_walk_entry(p, n, ce);
    if (rc) { ent.status = -rc; atomicAdd(&a.counters->n_errors, 1u); a.entries[e] = ent; return; }
    unsigned long long b0 = atomicAdd(&a.counters->n_blocks, (unsigned long long)ce.nb);
    unsigned long long l0 = atomicAdd(&a.counters->n_lit, (unsigned long long)ce.nlit);
    unsigned long long s0 = atomicAdd(&a.counters->n_seq, (unsigned long long)ce.nseq);
    if (b0 + ce.nb > a.cap_blocks || l0 + ce.nlit > a.cap_lit || s0 + ce.nseq > a.cap_seq) {
        atomicOr(&a.counters->overflow, 1u);
        ent.status = ZK_ST_RETRY; a.entries[e] = ent; return;
    }
    ZkFillEmit fe; fe.blocks = a.blocks; fe.entry = e; fe.bidx = (uint32_t)b0; fe.lit = (uint32_t)l0; fe.seq = (uint32_t)s0;
    rc = zk_walk_entry(p, n, fe);
    ent.first_block = (uint32_t)b0; ent.n_blocks = ce.nb;
    if (rc) { ent.status = -rc; ent.n_blocks = 0; atomicAdd(&a.counters->n_errors, 1u); }
    a.entries[e] = ent;
}

// =============================================================================================
// K-D1: per-block entropy decode
// =============================================================================================
#define ZK_D1_THREADS 64

struct ZkD1Smem {
    ZkSeqCell ll[512], ml[512], of[256];       // 10 KiB
    uint16_t huf[2048];                        // (nbBits << 8) | symbol, 4 KiB
    int16_t cnt[3][64];
    uint8_t symof[3][512];
    uint16_t nxt[3][64];
    uint8_t weights[256];
    uint16_t hpos[256];
    uint8_t wsym[64], wnb[64]; uint16_t wbase[64];   // FSE table for Huffman weights (log <= 6)
    int tbl_log[3], tbl_nsym[3], tbl_mode[3];   // mode: 0 counts, 1 rle (cnt[t][0] = symbol)
    uint32_t bits_off;                          // offset of the sequence bitstream inside the block
    int huf_bits;
    int st_lit, st_seq;
    uint32_t work;
};

// Locate the three table descriptions of a block (A.5).  desc_off[t] = offset of table t's description
// inside the block content; also parses FSE counts when `want[t]` (into sm.cnt[t]).  t: 0 LL, 1 OF, 2 ML.
// Returns 0 or a zstd code.  *bits_off = start of the bitstream.
__device__ int zk_locate_seq_tables(ZkD1Smem& sm, const uint8_t* b, uint32_t bsize, const bool want[3], uint32_t* bits_off) {
    ZkLitHdr lh;
    if (!zk_parse_lit_hdr(b, bsize, lh)) return ZKZ_CORRUPTION;
    uint32_t lsec = zk_lit_section_size(lh);
    if (lsec >= bsize) return ZKZ_CORRUPTION;
    uint32_t nseq, sh = zk_parse_nseq(b + lsec, bsize - lsec, nseq);
    if (!sh || nseq == 0 || lsec + sh >= bsize) return ZKZ_CORRUPTION;
    uint32_t modes = b[lsec + sh], pos = lsec + sh + 1;
    const int max_log[3] = {9, 8, 9}, max_sym[3] = {35, 31, 52};
    for (int t = 0; t < 3; t++) {
        uint32_t m = (modes >> (6 - 2 * t)) & 3;
        if (m == 0) {
            if (want[t]) {
                sm.tbl_mode[t] = 0;
                if (t == 0) { for (int i = 0; i < 36; i++) sm.cnt[0][i] = ZK_LL_DEFAULT[i]; sm.tbl_nsym[0] = 36; sm.tbl_log[0] = 6; }
                else if (t == 1) { for (int i = 0; i < 29; i++) sm.cnt[1][i] = ZK_OF_DEFAULT[i]; sm.tbl_nsym[1] = 29; sm.tbl_log[1] = 5; }
                else { for (int i = 0; i < 53; i++) sm.cnt[2][i] = ZK_ML_DEFAULT[i]; sm.tbl_nsym[2] = 53; sm.tbl_log[2] = 6; }
            }
        } else if (m == 1) {
            if (pos >= bsize) return ZKZ_CORRUPTION;
            if (want[t]) {
                if (b[pos] > max_sym[t]) return ZKZ_CORRUPTION;
                sm.tbl_mode[t] = 1; sm.cnt[t][0] = b[pos]; sm.tbl_log[t] = 0; sm.tbl_nsym[t] = 1;
            }
            pos += 1;
        } else if (m == 2) {
            int16_t scratch[64];
            int ns, lg;
            uint32_t used = zk_fse_read_ncount(b + pos, bsize - pos, max_log[t], max_sym[t], want[t] ? sm.cnt[t] : scratch, &ns, &lg);
            if (!used) return ZKZ_CORRUPTION;
            if (want[t]) { sm.tbl_mode[t] = 0; sm.tbl_nsym[t] = ns; sm.tbl_log[t] = lg; }
            pos += used;
        } else {
            if (want[t]) return ZKZ_CORRUPTION;   // caller resolves Repeat through *_ref first
        }
    }
    if (pos > bsize) return ZKZ_CORRUPTION;
    *bits_off = pos;
    return 0;
}

// Build one sequence decoding table from sm.cnt[t] (A.6).  Executed by lanes 0..2 of warp 0 in parallel.
__device__ int zk_build_seq_table(ZkD1Smem& sm, int t) {
    ZkSeqCell* cell = t == 0 ? sm.ll : (t == 1 ? sm.of : sm.ml);
    if (sm.tbl_mode[t] == 1) {
        int s = sm.cnt[t][0];
        ZkSeqCell c; c.next_base = 0; c.nb_bits = 0;
        if (t == 0) { c.base_value = ZK_LL_BASE[s]; c.add_bits = ZK_LL_BITS[s]; }
        else if (t == 1) { c.base_value = 1u << s; c.add_bits = (uint8_t)s; }
        else { c.base_value = ZK_ML_BASE[s]; c.add_bits = ZK_ML_BITS[s]; }
        cell[0] = c;
        return 0;
    }
    int log = sm.tbl_log[t], S = 1 << log, nsym = sm.tbl_nsym[t], high = S - 1;
    uint8_t* symof = sm.symof[t]; uint16_t* nxt = sm.nxt[t]; const int16_t* cnt = sm.cnt[t];
    for (int s = 0; s < nsym; s++) {
        if (cnt[s] == -1) { symof[high--] = (uint8_t)s; nxt[s] = 1; }
        else nxt[s] = (uint16_t)cnt[s];
    }
    int step = (S >> 1) + (S >> 3) + 3, pos = 0;
    for (int s = 0; s < nsym; s++)
        for (int q = 0; q < cnt[s]; q++) {
            symof[pos] = (uint8_t)s;
            do { pos = (pos + step) & (S - 1); } while (pos > high);
        }
    if (pos != 0) return ZKZ_CORRUPTION;
    for (int u = 0; u < S; u++) {
        int s = symof[u];
        uint32_t x = nxt[s]++;
        int nb = log - zk_highbit(x);
        ZkSeqCell c;
        c.nb_bits = (uint8_t)nb; c.next_base = (uint16_t)((x << nb) - S);
        if (t == 0) { c.base_value = ZK_LL_BASE[s]; c.add_bits = ZK_LL_BITS[s]; }
        else if (t == 1) { c.base_value = 1u << s; c.add_bits = (uint8_t)s; }
        else { c.base_value = ZK_ML_BASE[s]; c.add_bits = ZK_ML_BITS[s]; }
        cell[u] = c;
    }
    return 0;
}

// Huffman tree description -> sm.weights[0..nw) incl. the implied last weight (A.4).  Lane 0 only.
// Returns bytes consumed or 0 on corruption; sets sm.huf_bits.
__device__ uint32_t zk_read_huf_weights(ZkD1Smem& sm, const uint8_t* p, uint32_t n, int* nw_out) {
    if (n < 1) return 0;
    uint32_t hb = p[0], used; int nw = 0;
    if (hb >= 128) {
        nw = (int)hb - 127; used = 1 + (uint32_t)(nw + 1) / 2;
        if (used > n) return 0;
        for (int i = 0; i < nw; i++) { uint32_t b = p[1 + i / 2]; sm.weights[i] = (uint8_t)((i & 1) ? (b & 15) : (b >> 4)); }
    } else {
        used = 1 + hb;
        if (hb == 0 || used > n) return 0;
        int16_t cnt[16]; int ns, lg;
        uint32_t r = zk_fse_read_ncount(p + 1, hb, 6, 12, cnt, &ns, &lg);
        if (!r) return 0;
        // small FSE table (A.6)
        int S = 1 << lg, high = S - 1; uint16_t nx[16];
        for (int s = 0; s < ns; s++) { if (cnt[s] == -1) { sm.wsym[high--] = (uint8_t)s; nx[s] = 1; } else nx[s] = (uint16_t)cnt[s]; }
        int step = (S >> 1) + (S >> 3) + 3, pos = 0;
        for (int s = 0; s < ns; s++)
            for (int q = 0; q < cnt[s]; q++) { sm.wsym[pos] = (uint8_t)s; do { pos = (pos + step) & (S - 1); } while (pos > high); }
        if (pos != 0) return 0;
        for (int u = 0; u < S; u++) {
            int s = sm.wsym[u]; uint32_t x = nx[s]++; int nb = lg - zk_highbit(x);
            sm.wnb[u] = (uint8_t)nb; sm.wbase[u] = (uint16_t)((x << nb) - S);
        }
        ZkBackBits br;
        if (!br.init(p + 1 + r, hb - r)) return 0;
        br.refill();
        uint32_t s1 = br.read(lg), s2 = br.read(lg);
        if (br.bp < 0) return 0;
        for (;;) {   // two interleaved states, over-read terminates (A.4)
            if (nw >= 254) return 0;
            br.refill();
            sm.weights[nw++] = sm.wsym[s1];
            s1 = sm.wbase[s1] + br.read(sm.wnb[s1]);
            if (br.bp < 0) { sm.weights[nw++] = sm.wsym[s2]; break; }
            sm.weights[nw++] = sm.wsym[s2];
            s2 = sm.wbase[s2] + br.read(sm.wnb[s2]);
            if (br.bp < 0) { sm.weights[nw++] = sm.wsym[s1]; break; }
        }
    }
    uint32_t total = 0, n_w1 = 0;
    for (int i = 0; i < nw; i++) {
        uint32_t w = sm.weights[i];
        if (w > 11) return 0;
        if (w) total += 1u << (w - 1);
        n_w1 += (w == 1);
    }
    if (total == 0) return 0;
    int max_bits = zk_highbit(total) + 1;
    if (max_bits > 11) return 0;
    uint32_t left = (1u << max_bits) - total;
    if (left & (left - 1)) return 0;               // must be a power of two (left >= 1 by construction)
    uint32_t lw = (uint32_t)zk_highbit(left) + 1;
    sm.weights[nw++] = (uint8_t)lw;
    n_w1 += (lw == 1);
    if (n_w1 < 2 || (n_w1 & 1)) return 0;          // libzstd's HUF_readStats sanity rule
    sm.huf_bits = max_bits;
    *nw_out = nw;
    return used;
}

// Decode `cnt` Huffman symbols of one stream into out (A.4).  Returns false on corruption.
__device__ bool zk_huf_decode_stream(const uint16_t* tbl, int max_bits, const uint8_t* p, uint32_t n, uint8_t* out, uint32_t cnt) {
    ZkBackBits br;
    if (!br.init(p, n)) return false;
    uint32_t i = 0;
    while (i < cnt && ((uintptr_t)(out + i) & 3)) {
        br.refill();
        uint32_t e = tbl[br.peek(max_bits)]; br.bp -= (int)(e >> 8); out[i++] = (uint8_t)e;
    }
    for (; i + 4 <= cnt; i += 4) {
        br.refill();                                   // 4 * 11 bits <= 64
        uint32_t e0 = tbl[br.peek(max_bits)]; br.bp -= (int)(e0 >> 8);
        uint32_t e1 = tbl[br.peek(max_bits)]; br.bp -= (int)(e1 >> 8);
        uint32_t e2 = tbl[br.peek(max_bits)]; br.bp -= (int)(e2 >> 8);
        uint32_t e3 = tbl[br.peek(max_bits)]; br.bp -= (int)(e3 >> 8);
        *(uint32_t*)(out + i) = (e0 & 0xFF) | ((e1 & 0xFF) << 8) | ((e2 & 0xFF) << 16) | ((e3 & 0xFF) << 24);
    }
    for (; i < cnt; i++) {
        br.refill();
        uint32_t e = tbl[br.peek(max_bits)]; br.bp -= (int)(e >> 8); out[i] = (uint8_t)e;
    }
    return br.bp == 0;
}

__device__ void zk_d1_literals(ZkD1Smem& sm, const ZkDecodeArgs& a, ZkBlock& blk, uint32_t bidx, const uint8_t* ebase, int lane) {
    const uint8_t* b = ebase + blk.src;
    ZkLitHdr lh;
    zk_parse_lit_hdr(b, blk.size, lh);        // validated by the scan kernel
    if (lh.type == 0) {
        if (lane == 0) { a.blocks[bidx].lit_kind = 0; a.blocks[bidx].lit_src = blk.src + lh.hdr; sm.st_lit = 0; }
        return;
    }
    if (lh.type == 1) {
        if (lane == 0) { a.blocks[bidx].lit_kind = 1; a.blocks[bidx].lit_byte = b[lh.hdr]; sm.st_lit = 0; }
        return;
    }
    // Huffman-compressed: find the tree description (own block, or the block a Treeless block refers to)
    const uint8_t* q = b + lh.hdr; uint32_t qn = lh.comp;
    int nw = 0; uint32_t tree_bytes = 0;
    if (lane == 0) {
        int st = 0;
        if (lh.type == 2) {
            tree_bytes = zk_read_huf_weights(sm, q, qn, &nw);
            if (!tree_bytes) st = ZKZ_CORRUPTION;
        } else {
            const ZkBlock& rb = a.blocks[blk.huf_ref];
            const uint8_t* r = ebase + rb.src;
            ZkLitHdr rh;
            zk_parse_lit_hdr(r, rb.size, rh);
            if (!zk_read_huf_weights(sm, r + rh.hdr, rh.comp, &nw)) st = ZKZ_CORRUPTION;
        }
        if (!st) {   // start cell of every symbol: weight ascending, symbols in natural order
            uint32_t rank_start[13]; uint32_t cntw[13];
            for (int w = 0; w < 13; w++) cntw[w] = 0;
            for (int s = 0; s < nw; s++) cntw[sm.weights[s]]++;
            uint32_t acc = 0;
            for (int w = 1; w <= sm.huf_bits; w++) { rank_start[w] = acc; acc += cntw[w] << (w - 1); }
            for (int s = 0; s < nw; s++) { int w = sm.weights[s]; if (w) { sm.hpos[s] = (uint16_t)rank_start[w]; rank_start[w] += 1u << (w - 1); } }
            if (acc != (1u << sm.huf_bits)) st = ZKZ_CORRUPTION;
        }
        sm.st_lit = st;
    }
    __syncwarp();
    nw = __shfl_sync(0xFFFFFFFFu, nw, 0);
    tree_bytes = __shfl_sync(0xFFFFFFFFu, tree_bytes, 0);
    if (sm.st_lit) return;
    int max_bits = sm.huf_bits;
    for (int s = lane; s < nw; s += 32) {
        int w = sm.weights[s];
        if (!w) continue;
        uint32_t len = 1u << (w - 1), pos = sm.hpos[s];
        uint16_t e = (uint16_t)(((max_bits + 1 - w) << 8) | s);
        for (uint32_t i = 0; i < len; i++) sm.huf[pos + i] = e;
    }
    __syncwarp();
    q += tree_bytes; qn -= tree_bytes;
    uint8_t* out = a.lit + blk.lit_base;
    bool ok = true;
    if (lh.streams == 1) {
        if (lane == 0) ok = zk_huf_decode_stream(sm.huf, max_bits, q, qn, out, lh.regen);
    } else {
        uint32_t seg = (lh.regen + 3) / 4;
        if (qn < 10 || lh.regen < 6 || seg * 3 > lh.regen) ok = false;   // 6-byte jump table + 4 non-empty streams; libzstd rejects regen < 6
        else {
            uint32_t s1 = zk_ld_le16(q), s2 = zk_ld_le16(q + 2), s3 = zk_ld_le16(q + 4);
            if (6 + s1 + s2 + s3 >= qn) ok = false;
            else if (lane < 4) {
                uint32_t s4 = qn - 6 - s1 - s2 - s3;
                uint32_t off = lane == 0 ? 0 : (lane == 1 ? s1 : (lane == 2 ? s1 + s2 : s1 + s2 + s3));
                uint32_t len = lane == 0 ? s1 : (lane == 1 ? s2 : (lane == 2 ? s3 : s4));
                uint32_t cnt = lane < 3 ? seg : lh.regen - 3 * seg;
                ok = zk_huf_decode_stream(sm.huf, max_bits, q + 6 + off, len, out + lane * seg, cnt);
            }
        }
    }
    uint32_t bad = __ballot_sync(0xFFFFFFFFu, !ok);
    if (lane == 0) {
        a.blocks[bidx].lit_kind = 2;
        if (bad) sm.st_lit = ZKZ_CORRUPTION;
    }
}

__device__ void zk_d1_sequences(ZkD1Smem& sm, const ZkDecodeArgs& a, ZkBlock& blk, uint32_t bidx, const uint8_t* ebase, int lane) {
    const uint8_t* b = ebase + blk.src;
    if (blk.nseq == 0) { if (lane == 0) sm.st_seq = 0; return; }
    if (lane == 0) {
        int st = 0;
        bool want[3] = { blk.ll_ref < 0, blk.of_ref < 0, blk.ml_ref < 0 };
        uint32_t bits_off = 0;
        // own block: all non-Repeat tables (Repeat ones return 0 bytes and are not wanted)
        st = zk_locate_seq_tables(sm, b, blk.size, want, &bits_off);
        sm.bits_off = bits_off;
        const int32_t refs[3] = { blk.ll_ref, blk.of_ref, blk.ml_ref };
        for (int t = 0; t < 3 && !st; t++) {
            if (refs[t] < 0) continue;
            const ZkBlock& rb = a.blocks[refs[t]];
            bool w2[3] = { t == 0, t == 1, t == 2 };
            uint32_t dummy;
            st = zk_locate_seq_tables(sm, ebase + rb.src, rb.size, w2, &dummy);
        }
        sm.st_seq = st;
    }
    __syncwarp();
    if (sm.st_seq) return;
    int bst = 0;
    if (lane < 3) bst = zk_build_seq_table(sm, lane);
    uint32_t bad = __ballot_sync(0xFFFFFFFFu, bst != 0);
    if (bad) { if (lane == 0) sm.st_seq = ZKZ_CORRUPTION; return; }
    __syncwarp();
    if (lane != 0) return;

    // ---- serial FSE decode of the interleaved LL/OF/ML states (A.5)
    ZkBackBits br;
    uint32_t bits_off = sm.bits_off;
    if (!br.init(b + bits_off, blk.size - bits_off)) { sm.st_seq = ZKZ_CORRUPTION; return; }
    const int ll_log = sm.tbl_log[0], of_log = sm.tbl_log[1], ml_log = sm.tbl_log[2];
    br.refill();
    uint32_t sl = br.read(ll_log), so = br.read(of_log), smm = br.read(ml_log);
    uint32_t r0 = ZK_SYM_MAKE(0, 0), r1 = ZK_SYM_MAKE(1, 0), r2 = ZK_SYM_MAKE(2, 0);
    uint32_t lit_end = 0, out_end = 0;
    uint32_t* o_lit = a.seq_lit_end + blk.seq_base;
    uint32_t* o_out = a.seq_out_end + blk.seq_base;
    uint32_t* o_off = a.seq_off + blk.seq_base;
    const uint32_t nseq = blk.nseq;
    int st = 0;
    for (uint32_t i = 0; i < nseq; i++) {
        ZkSeqCell cl = sm.ll[sl], co = sm.of[so], cm = sm.ml[smm];
        br.refill();                                                  // <= 31 + 16 + 16 bits follow
        uint32_t ofv = co.base_value + br.read(co.add_bits);
        uint32_t mlv = cm.base_value + br.read(cm.add_bits);
        uint32_t llv = cl.base_value + br.read(cl.add_bits);
        br.refill();                                                  // <= 9 + 9 + 8 bits follow
        if (i + 1 < nseq) {
            sl = cl.next_base + br.read(cl.nb_bits);
            smm = cm.next_base + br.read(cm.nb_bits);
            so = co.next_base + br.read(co.nb_bits);
        }
        // repeat-offset history, kept symbolic w.r.t. the (unknown) state entering this block
        uint32_t off;
        if (ofv > 3) { off = ofv - 3; r2 = r1; r1 = r0; r0 = off; }
        else {
            uint32_t idx = ofv - 1 + (llv == 0);
            if (idx == 0) off = r0;
            else {
                if (idx == 3) {
                    if (r0 & ZK_SYM) off = r0 + 1;                    // delta + 1
                    else { off = r0 - 1; if (off == 0) st = ZKZ_CORRUPTION; }
                } else off = idx == 1 ? r1 : r2;
                if (idx != 1) r2 = r1;
                r1 = r0; r0 = off;
            }
        }
        lit_end += llv; out_end += llv + mlv;
        o_lit[i] = lit_end; o_out[i] = out_end; o_off[i] = off;
    }
    if (br.bp != 0) st = ZKZ_CORRUPTION;
    if (lit_end > blk.lit_size) st = ZKZ_CORRUPTION;
    else if (out_end + (blk.lit_size - lit_end) > ZK_BLOCK_MAX) st = ZKZ_CORRUPTION;
    a.blocks[bidx].rep_out[0] = r0; a.blocks[bidx].rep_out[1] = r1; a.blocks[bidx].rep_out[2] = r2;
    a.blocks[bidx].regen = out_end + (blk.lit_size - lit_end);
    sm.st_seq = st;
}

__global__ void __launch_bounds__(ZK_D1_THREADS) zk_entropy_kernel(ZkDecodeArgs a) {
    __shared__ ZkD1Smem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sm.work = atomicAdd(a.work_counter, 1u);
        __syncthreads();
        uint32_t bidx = sm.work;
        unsigned long long nb = a.counters->n_blocks;
        if (a.counters->overflow) nb = 0;                   // scratch too small: the host grows it and re-runs the batch
        if (bidx >= nb) break;
        ZkBlock blk = a.blocks[bidx];
        if (blk.type != 2) continue;                         // Raw / RLE: regen preset by the scan kernel
        if (a.entries[blk.entry].status != 0) continue;
        const uint8_t* ebase = a.comp + a.c_off[blk.entry];
        if (threadIdx.x == 0) { sm.st_lit = 0; sm.st_seq = 0; }
        __syncthreads();
        if (warp == 0) zk_d1_sequences(sm, a, blk, bidx, ebase, lane);
        else zk_d1_literals(sm, a, blk, bidx, ebase, lane);
        __syncthreads();
        if (threadIdx.x == 0) {
            int st = sm.st_lit ? sm.st_lit : sm.st_seq;
            if (blk.nseq == 0) a.blocks[bidx].regen = blk.lit_size;
            a.blocks[bidx].status = st ? -st : 0;
        }
    }
}

// =============================================================================================
// K-D2: ordered sequence execution
// =============================================================================================
// warp-cooperative copy of n bytes, non-overlapping (or src entirely before dst with distance >= n)
__device__ __forceinline__ void zk_warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, int lane) {
    uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    if (lane < (int)head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    uint32_t nvec = n >> 4;
    uint32_t mis = (uint32_t)((uintptr_t)src & 3);
    if (mis == 0) {
        if (((uintptr_t)src & 15) == 0) {
            for (uint32_t i = lane; i < nvec; i += 32) ((uint4*)dst)[i] = ((const uint4*)src)[i];
        } else {
            for (uint32_t i = lane; i < nvec; i += 32) {
                const uint32_t* s = (const uint32_t*)src + 4 * i;
                ((uint4*)dst)[i] = make_uint4(s[0], s[1], s[2], s[3]);
            }
        }
    } else {
        const uint32_t* sa = (const uint32_t*)(src - mis);
        uint32_t sh = mis * 8;
        for (uint32_t i = lane; i < nvec; i += 32) {
            const uint32_t* s = sa + 4 * i;
            uint32_t w0 = s[0], w1 = s[1], w2 = s[2], w3 = s[3], w4 = s[4];
            ((uint4*)dst)[i] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                                          __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
    }
    uint32_t done = nvec << 4;
    for (uint32_t i = done + lane; i < n; i += 32) dst[i] = src[i];
}

__device__ __forceinline__ void zk_warp_fill(uint8_t* dst, uint32_t byte, uint32_t n, int lane) {
    uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    if (lane < (int)head) dst[lane] = (uint8_t)byte;
    dst += head; n -= head;
    uint32_t w = byte * 0x01010101u, nvec = n >> 4;
    for (uint32_t i = lane; i < nvec; i += 32) ((uint4*)dst)[i] = make_uint4(w, w, w, w);
    for (uint32_t i = (nvec << 4) + lane; i < n; i += 32) dst[i] = (uint8_t)byte;
}

// warp-cooperative match copy dst[0..n) = dst[-off ..), overlap allowed (period doubling)
__device__ __forceinline__ void zk_warp_match(uint8_t* dst, uint32_t off, uint32_t n, int lane) {
    const uint8_t* src = dst - off;
    if (off >= n) { zk_warp_copy(dst, src, n, lane); return; }
    uint32_t have = off, done = 0;            // [src, src+have) is final periodic data
    while (done < n) {
        uint32_t c = n - done < have ? n - done : have;
        zk_warp_copy(dst + done, src, c, lane);
        __syncwarp();
        done += c; have += c;
    }
}

#define ZK_LONG 48u     // sequences whose literal run / match is at least this long are copied by the whole warp

struct ZkD2Smem {
    volatile uint32_t done_pos;      // every output byte below this position (entry-relative) is final
    volatile uint32_t done_chunk;    // chunks [0, done_chunk) are published
    volatile int abort_code;
};

__device__ __forceinline__ void zk_d2_abort(ZkD2Smem& sm, int code) { atomicCAS((int*)&sm.abort_code, 0, code); }
// warp-uniform view of the abort flag (every lane must take the same branch around collectives)
__device__ __forceinline__ bool zk_d2_aborted(ZkD2Smem& sm) { return __any_sync(0xFFFFFFFFu, sm.abort_code != 0); }

// wait until it is chunk c's turn, then publish its end position
__device__ __forceinline__ void zk_d2_publish(ZkD2Smem& sm, uint32_t c, uint32_t end_pos, int lane) {
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
        while (sm.done_chunk != c && sm.abort_code == 0) ZK_SPIN();
        sm.done_pos = end_pos;
        __threadfence_block();
        sm.done_chunk = c + 1;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(512) zk_exec_kernel(ZkDecodeArgs a) {
    __shared__ ZkD2Smem sm;
    const uint32_t e = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    ZkEntry ent = a.entries[e];
    if (ent.status != 0 || a.counters->overflow) return;
    if (threadIdx.x == 0) { sm.done_pos = 0; sm.done_chunk = 0; sm.abort_code = 0; }
    __syncthreads();
    uint8_t* out = a.dst + a.d_off[e];
    const unsigned long long cap64 = a.d_off[e + 1] - a.d_off[e];
    const uint32_t cap = cap64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cap64;
    const uint8_t* ebase = a.comp + a.c_off[e];

    uint32_t pos = 0, zstart = 0, chunk_base = 0;
    uint32_t R0 = 1, R1 = 4, R2 = 8;
    for (uint32_t bi = 0; bi < ent.n_blocks; bi++) {
        if (zk_d2_aborted(sm)) break;
        const uint32_t bidx = ent.first_block + bi;
        const ZkBlock blk = a.blocks[bidx];
        if (blk.flags & ZKB_FIRST) { R0 = 1; R1 = 4; R2 = 8; zstart = pos; }
        if (blk.status != 0) { zk_d2_abort(sm, -blk.status); break; }
        const bool has_seq = blk.type == 2 && blk.nseq > 0;
        const uint32_t nchunks = has_seq ? (blk.nseq + 31) / 32 + 1 : 1;
        if ((unsigned long long)pos + blk.regen > cap) { zk_d2_abort(sm, ZKZ_DST_TOO_SMALL); break; }
        // first chunk index of this block that belongs to this warp
        uint32_t c = chunk_base + ((uint32_t)warp + W - (chunk_base % W)) % W;
        for (; c < chunk_base + nchunks; c += W) {
            if (zk_d2_aborted(sm)) break;
            const uint32_t j = c - chunk_base;
            if (!has_seq) {
                // ---------------- Raw block / RLE block / literals-only compressed block
                if (blk.type == 0) zk_warp_copy(out + pos, ebase + blk.src, blk.size, lane);
                else if (blk.type == 1) zk_warp_fill(out + pos, ebase[blk.src], blk.size, lane);
                else if (blk.lit_kind == 1) zk_warp_fill(out + pos, blk.lit_byte, blk.lit_size, lane);
                else zk_warp_copy(out + pos, blk.lit_kind == 0 ? ebase + blk.lit_src : a.lit + blk.lit_base, blk.lit_size, lane);
                zk_d2_publish(sm, c, pos + blk.regen, lane);
                continue;
            }
            const uint32_t* s_lit = a.seq_lit_end + blk.seq_base;
            const uint32_t* s_out = a.seq_out_end + blk.seq_base;
            const uint8_t* lit = blk.lit_kind == 0 ? ebase + blk.lit_src : a.lit + blk.lit_base;
            if (j == nchunks - 1) {
                // ---------------- trailing literals of the block
                uint32_t le = s_lit[blk.nseq - 1], oe = s_out[blk.nseq - 1];
                uint32_t n = blk.lit_size - le;
                if (blk.lit_kind == 1) zk_warp_fill(out + pos + oe, blk.lit_byte, n, lane);
                else zk_warp_copy(out + pos + oe, lit + le, n, lane);
                zk_d2_publish(sm, c, pos + blk.regen, lane);
                continue;
            }
            // ---------------- 32 sequences, one per lane
            const uint32_t s = j * 32 + lane;
            const bool valid = s < blk.nseq;
            uint32_t le = 0, oe = 0, offv = 0;
            if (valid) { le = s_lit[s]; oe = s_out[s]; offv = a.seq_off[blk.seq_base + s]; }
            uint32_t le_prev = __shfl_up_sync(0xFFFFFFFFu, le, 1), oe_prev = __shfl_up_sync(0xFFFFFFFFu, oe, 1);
            if (lane == 0) { le_prev = s ? s_lit[s - 1] : 0; oe_prev = s ? s_out[s - 1] : 0; }
            const uint32_t ll = le - le_prev, ml = (oe - oe_prev) - ll;
            const uint32_t o_lit = pos + oe_prev, md = o_lit + ll;          // entry-relative positions
            uint32_t off = offv;
            bool bad = false;
            if (valid) {
                if (offv & ZK_SYM) {
                    uint32_t sl = ZK_SYM_SLOT(offv), dl = ZK_SYM_DELTA(offv);
                    uint32_t r = sl == 0 ? R0 : (sl == 1 ? R1 : R2);
                    bad = r <= dl; off = r - dl;
                }
                if (off == 0 || off > md - zstart) bad = true;
            }
            if (__any_sync(0xFFFFFFFFu, bad)) { zk_d2_abort(sm, ZKZ_CORRUPTION); break; }
            const uint32_t chunk_end = pos + __shfl_sync(0xFFFFFFFFu, oe, min(31u, blk.nseq - 1 - j * 32));

            // literal runs: no dependencies
            if (valid && ll < ZK_LONG) {
                uint8_t* d = out + o_lit;
                if (blk.lit_kind == 1) for (uint32_t i = 0; i < ll; i++) d[i] = blk.lit_byte;
                else { const uint8_t* sp = lit + le_prev; for (uint32_t i = 0; i < ll; i++) d[i] = sp[i]; }
            }
            uint32_t longlit = __ballot_sync(0xFFFFFFFFu, valid && ll >= ZK_LONG);
            while (longlit) {
                int l = __ffs((int)longlit) - 1; longlit &= longlit - 1;
                uint32_t n = __shfl_sync(0xFFFFFFFFu, ll, l), d = __shfl_sync(0xFFFFFFFFu, o_lit, l), sp = __shfl_sync(0xFFFFFFFFu, le_prev, l);
                if (blk.lit_kind == 1) zk_warp_fill(out + d, blk.lit_byte, n, lane);
                else zk_warp_copy(out + d, lit + sp, n, lane);
            }
            __syncwarp();

            // matches: a lane may go once every byte of its source is final
            const uint32_t need_end = md - off + (ml < off ? ml : off);
            uint32_t pending = __ballot_sync(0xFFFFFFFFu, valid && ml > 0);
            bool aborted = false;
            while (pending) {
                // lane 0 samples the pipeline state; everything below is warp-uniform
                uint32_t dc = 0, dp = 0, ab = 0;
                if (lane == 0) { dc = sm.done_chunk; dp = sm.done_pos; ab = sm.abort_code != 0; __threadfence_block(); }
                dc = __shfl_sync(0xFFFFFFFFu, dc, 0); dp = __shfl_sync(0xFFFFFFFFu, dp, 0); ab = __shfl_sync(0xFFFFFFFFu, ab, 0);
                if (ab) { aborted = true; break; }
                const bool oldest = dc == c;
                const int first = __ffs((int)pending) - 1;
                const uint32_t md_first = __shfl_sync(0xFFFFFFFFu, md, first);
                const uint32_t frontier = oldest ? md_first : dp;
                const bool mine = (pending >> lane) & 1;
                const bool ready = mine && (need_end <= frontier || (oldest && lane == first));
                const uint32_t rmask = __ballot_sync(0xFFFFFFFFu, ready);
                if (!rmask) { ZK_SPIN(); continue; }
                if (ready && ml < ZK_LONG) {
                    uint8_t* d = out + md; const uint8_t* sp = d - off;
                    for (uint32_t i = 0; i < ml; i++) d[i] = sp[i];
                }
                uint32_t longm = __ballot_sync(0xFFFFFFFFu, ready && ml >= ZK_LONG);
                while (longm) {
                    int l = __ffs((int)longm) - 1; longm &= longm - 1;
                    uint32_t n = __shfl_sync(0xFFFFFFFFu, ml, l), d = __shfl_sync(0xFFFFFFFFu, md, l), o = __shfl_sync(0xFFFFFFFFu, off, l);
                    zk_warp_match(out + d, o, n, lane);
                }
                __syncwarp();
                pending &= ~rmask;
            }
            if (aborted) break;
            zk_d2_publish(sm, c, chunk_end, lane);
        }
        // advance to the next block
        if (has_seq) {
            uint32_t v[3] = { blk.rep_out[0], blk.rep_out[1], blk.rep_out[2] }, n[3];
            for (int q = 0; q < 3; q++) {
                if (v[q] & ZK_SYM) { uint32_t sl = ZK_SYM_SLOT(v[q]); uint32_t r = sl == 0 ? R0 : (sl == 1 ? R1 : R2); n[q] = r - ZK_SYM_DELTA(v[q]); }
                else n[q] = v[q];
            }
            R0 = n[0]; R1 = n[1]; R2 = n[2];
        }
        pos += blk.regen;
        chunk_base += nchunks;
        if (blk.flags & ZKB_LAST) {
            if ((blk.flags & ZKB_HAS_FCS) && blk.fcs != (unsigned long long)(pos - zstart)) { zk_d2_abort(sm, ZKZ_CORRUPTION); break; }
            if (threadIdx.x == 0) { a.blocks[bidx].hash_start = zstart; a.blocks[bidx].hash_len = pos - zstart; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int code = sm.abort_code;
        if (!code && (unsigned long long)pos != cap64) code = pos < cap64 ? ZKZ_SRC_SIZE_WRONG : ZKZ_DST_TOO_SMALL;
        a.entries[e].status = code ? -code : 0;
        a.entries[e].produced = pos;
        if (code) atomicAdd(&a.counters->n_errors, 1u);
    }
}

// =============================================================================================
// K-D3: XXH64 content checksum (A.8), one warp per entry, lanes 0..3 carry the four accumulators
// =============================================================================================
#define ZK_P1 0x9E3779B185EBCA87ull
#define ZK_P2 0xC2B2AE3D27D4EB4Full
#define ZK_P3 0x165667B19E3779F9ull
#define ZK_P4 0x85EBCA77C2B2AE63ull
#define ZK_P5 0x27D4EB2F165667C5ull
__device__ __forceinline__ unsigned long long zk_rotl64(unsigned long long x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ unsigned long long zk_xx_round(unsigned long long acc, unsigned long long in) { return zk_rotl64(acc + in * ZK_P2, 31) * ZK_P1; }
__device__ __forceinline__ unsigned long long zk_xx_merge(unsigned long long h, unsigned long long v) { return (h ^ zk_xx_round(0, v)) * ZK_P1 + ZK_P4; }
__device__ __forceinline__ unsigned long long zk_ld_u64_unaligned(const uint8_t* p) {
    uintptr_t a = (uintptr_t)p; uint32_t mis = (uint32_t)(a & 7);
    const unsigned long long* q = (const unsigned long long*)(a - mis);
    if (mis == 0) return q[0];
    return (q[0] >> (mis * 8)) | (q[1] << (64 - mis * 8));
}

// whole-warp XXH64 of p[0..len); result valid in every lane
__device__ unsigned long long zk_warp_xxh64(const uint8_t* p, uint32_t len, int lane) {
    unsigned long long h;
    uint32_t done = 0;
    if (len >= 32) {
        unsigned long long acc = lane == 0 ? ZK_P1 + ZK_P2 : (lane == 1 ? ZK_P2 : (lane == 2 ? 0ull : 0ull - ZK_P1));
        uint32_t stripes = len / 32;
        if (lane < 4) {
            const uint8_t* q = p + lane * 8;
            uint32_t i = 0;
            for (; i + 4 <= stripes; i += 4) {          // 4 loads in flight per lane
                unsigned long long w0 = zk_ld_u64_unaligned(q), w1 = zk_ld_u64_unaligned(q + 32),
                                   w2 = zk_ld_u64_unaligned(q + 64), w3 = zk_ld_u64_unaligned(q + 96);
                acc = zk_xx_round(acc, w0); acc = zk_xx_round(acc, w1); acc = zk_xx_round(acc, w2); acc = zk_xx_round(acc, w3);
                q += 128;
            }
            for (; i < stripes; i++) { acc = zk_xx_round(acc, zk_ld_u64_unaligned(q)); q += 32; }
        }
        unsigned long long v1 = __shfl_sync(0xFFFFFFFFu, acc, 0), v2 = __shfl_sync(0xFFFFFFFFu, acc, 1),
                           v3 = __shfl_sync(0xFFFFFFFFu, acc, 2), v4 = __shfl_sync(0xFFFFFFFFu, acc, 3);
        h = zk_rotl64(v1, 1) + zk_rotl64(v2, 7) + zk_rotl64(v3, 12) + zk_rotl64(v4, 18);
        h = zk_xx_merge(h, v1); h = zk_xx_merge(h, v2); h = zk_xx_merge(h, v3); h = zk_xx_merge(h, v4);
        done = stripes * 32;
    } else h = ZK_P5;
    h += (unsigned long long)len;
    const uint8_t* q = p + done; uint32_t rem = len - done;
    while (rem >= 8) { h ^= zk_xx_round(0, zk_ld_u64_unaligned(q)); h = zk_rotl64(h, 27) * ZK_P1 + ZK_P4; q += 8; rem -= 8; }
    if (rem >= 4) { h ^= (unsigned long long)zk_ld_le32(q) * ZK_P1; h = zk_rotl64(h, 23) * ZK_P2 + ZK_P3; q += 4; rem -= 4; }
    while (rem) { h ^= (unsigned long long)(*q) * ZK_P5; h = zk_rotl64(h, 11) * ZK_P1; q++; rem--; }
    h ^= h >> 33; h *= ZK_P2; h ^= h >> 29; h *= ZK_P3; h ^= h >> 32;
    return h;
}

__global__ void __launch_bounds__(128) zk_xxh64_kernel(ZkDecodeArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= a.n_entries) return;
    ZkEntry ent = a.entries[e];
    if (ent.status != 0 || a.counters->overflow) return;
    const uint8_t* out = a.dst + a.d_off[e];
    const uint8_t* ebase = a.comp + a.c_off[e];
    for (uint32_t bi = 0; bi < ent.n_blocks; bi++) {
        const ZkBlock* blk = &a.blocks[ent.first_block + bi];
        uint32_t flags = blk->flags;
        if ((flags & (ZKB_LAST | ZKB_HAS_CSUM)) != (ZKB_LAST | ZKB_HAS_CSUM)) continue;
        uint32_t content = blk->type == 1 ? 1u : blk->size;
        uint32_t want = zk_ld_le32(ebase + blk->src + content);
        unsigned long long h = zk_warp_xxh64(out + blk->hash_start, blk->hash_len, lane);
        if ((uint32_t)h != want) {
            if (lane == 0) { a.entries[e].status = -ZKZ_CHECKSUM_WRONG; atomicAdd(&a.counters->n_errors, 1u); }
            return;
        }
    }
}

// =============================================================================================
// host-side launcher
// =============================================================================================
#ifndef ZK_EMUL
#define ZK_CUDA_OK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) return -(int)ZKZ_GENERIC; } while (0)
#else
#define ZK_CUDA_OK(x) do { (void)(x); } while (0)
#endif

static int zk_grow(void** p, size_t* cap, size_t need, size_t elem) {
    if (*cap >= need && *p) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    size_t want = need + need / 8 + 64;
    if (cudaMalloc(p, want * elem) != cudaSuccess) { *p = nullptr; return -(int)ZKZ_MEMORY_ALLOCATION; }
    *cap = want;
    return 0;
}

void zk_decode_ws_free(ZkDecodeWs* ws) {
    void* ptrs[] = { ws->blocks, ws->entries, ws->counters, ws->lit, ws->seq_lit_end, ws->seq_out_end, ws->seq_off, ws->c_off, ws->d_off };
    for (void* p : ptrs) if (p) cudaFree(p);
    if (ws->h_entries) cudaFreeHost(ws->h_entries);
    if (ws->h_counters) cudaFreeHost(ws->h_counters);
    if (ws->h_off) cudaFreeHost(ws->h_off);
    *ws = ZkDecodeWs();
}

static int zk_decode_ensure(ZkDecodeWs* ws, uint32_t n, size_t need_blocks, size_t need_lit, size_t need_seq) {
    int rc;
    size_t cap;
    cap = ws->cap_blocks; if ((rc = zk_grow((void**)&ws->blocks, &cap, need_blocks, sizeof(ZkBlock)))) return rc; ws->cap_blocks = cap;
    cap = ws->cap_lit; if ((rc = zk_grow((void**)&ws->lit, &cap, need_lit + 64, 1))) return rc; ws->cap_lit = cap;
    if (ws->cap_seq < need_seq || !ws->seq_off) {
        size_t c1 = ws->cap_seq, c2 = ws->cap_seq, c3 = ws->cap_seq;
        if ((rc = zk_grow((void**)&ws->seq_lit_end, &c1, need_seq, 4))) return rc;
        if ((rc = zk_grow((void**)&ws->seq_out_end, &c2, need_seq, 4))) return rc;
        if ((rc = zk_grow((void**)&ws->seq_off, &c3, need_seq, 4))) return rc;
        ws->cap_seq = c1 < c2 ? (c1 < c3 ? c1 : c3) : (c2 < c3 ? c2 : c3);
    }
    if (ws->cap_entries < n || !ws->entries) {
        size_t c = ws->cap_entries;
        if ((rc = zk_grow((void**)&ws->entries, &c, n, sizeof(ZkEntry)))) return rc;
        size_t c2 = 0, c3 = 0;
        if (ws->c_off) { cudaFree(ws->c_off); ws->c_off = nullptr; }
        if (ws->d_off) { cudaFree(ws->d_off); ws->d_off = nullptr; }
        if ((rc = zk_grow((void**)&ws->c_off, &c2, c + 1, 8))) return rc;
        if ((rc = zk_grow((void**)&ws->d_off, &c3, c + 1, 8))) return rc;
        if (ws->h_entries) cudaFreeHost(ws->h_entries);
        if (ws->h_off) cudaFreeHost(ws->h_off);
        if (cudaMallocHost((void**)&ws->h_entries, c * sizeof(ZkEntry)) != cudaSuccess) return -(int)ZKZ_MEMORY_ALLOCATION;
        if (cudaMallocHost((void**)&ws->h_off, 2 * (c + 1) * 8) != cudaSuccess) return -(int)ZKZ_MEMORY_ALLOCATION;
        ws->cap_entries = c;
    }
    if (!ws->counters) {
        if (cudaMalloc((void**)&ws->counters, sizeof(ZkCounters) + 16) != cudaSuccess) return -(int)ZKZ_MEMORY_ALLOCATION;
        if (cudaMallocHost((void**)&ws->h_counters, sizeof(ZkCounters) + 16) != cudaSuccess) return -(int)ZKZ_MEMORY_ALLOCATION;
    }
    return 0;
}

// Enqueue one batch on `stream` (scan, entropy, exec[, checksum], status read-back).  No host sync.
int zk_decode_enqueue(ZkDecodeWs* ws, cudaStream_t stream, const uint8_t* d_comp, const uint64_t* c_off, const uint64_t* d_off,
                      uint32_t n, uint8_t* d_dst, int verify_checksum, int exec_warps) {
    ws->pending_n = 0;
    if (n == 0) return 0;
    unsigned long long total_d = d_off[n] - d_off[0];
    // optimistic scratch sizing; the scan kernel reports exact needs and zk_decode_collect retries if exceeded
    size_t need_blocks = (size_t)(total_d / ZK_BLOCK_MAX) * 2 + 4 * (size_t)n + 64;
    size_t need_lit = (size_t)total_d + 16 * need_blocks;
    size_t need_seq = (size_t)(total_d / 4) + 1024;
    if (need_blocks < ws->want_blocks) need_blocks = ws->want_blocks;
    if (need_lit < ws->want_lit) need_lit = ws->want_lit;
    if (need_seq < ws->want_seq) need_seq = ws->want_seq;
    int sms = ws->sm_count > 0 ? ws->sm_count : 148;
    int rc = zk_decode_ensure(ws, n, need_blocks, need_lit, need_seq);
    if (rc) return rc;
    memcpy(ws->h_off, c_off, (size_t)(n + 1) * 8);
    memcpy(ws->h_off + (n + 1), d_off, (size_t)(n + 1) * 8);
    ZK_CUDA_OK(cudaMemcpyAsync(ws->c_off, ws->h_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, stream));
    ZK_CUDA_OK(cudaMemcpyAsync(ws->d_off, ws->h_off + (n + 1), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, stream));
    ZK_CUDA_OK(cudaMemsetAsync(ws->counters, 0, sizeof(ZkCounters) + 16, stream));
    ZkDecodeArgs a;
    a.comp = d_comp; a.c_off = (const unsigned long long*)ws->c_off; a.d_off = (const unsigned long long*)ws->d_off; a.dst = d_dst; a.n_entries = n;
    a.blocks = ws->blocks; a.entries = ws->entries; a.counters = ws->counters;
    a.work_counter = (uint32_t*)((uint8_t*)ws->counters + sizeof(ZkCounters));
    a.lit = ws->lit; a.seq_lit_end = ws->seq_lit_end; a.seq_out_end = ws->seq_out_end; a.seq_off = ws->seq_off;
    a.cap_blocks = ws->cap_blocks; a.cap_lit = ws->cap_lit - 64; a.cap_seq = ws->cap_seq;
    ZK_LAUNCH(zk_scan_kernel, (n + 127) / 128, 128, 0, stream, a);
    uint32_t g1 = (uint32_t)sms * 16;
    size_t est_blocks = (size_t)(total_d / ZK_BLOCK_MAX) + n;
    if (est_blocks < g1) g1 = (uint32_t)(est_blocks < 1 ? 1 : est_blocks);
    ZK_LAUNCH(zk_entropy_kernel, g1, ZK_D1_THREADS, 0, stream, a);
    int W = exec_warps;
    if (W <= 0) { long per = ((long)sms * 48) / (long)n; W = per < 4 ? 4 : (per > 16 ? 16 : (int)per); }
    if (W > 16) W = 16;
    ZK_LAUNCH(zk_exec_kernel, n, W * 32, 0, stream, a);
    if (verify_checksum) ZK_LAUNCH(zk_xxh64_kernel, (n + 3) / 4, 128, 0, stream, a);
    ZK_CUDA_OK(cudaMemcpyAsync(ws->h_entries, ws->entries, (size_t)n * sizeof(ZkEntry), cudaMemcpyDeviceToHost, stream));
    ZK_CUDA_OK(cudaMemcpyAsync(ws->h_counters, ws->counters, sizeof(ZkCounters), cudaMemcpyDeviceToHost, stream));
    ws->launches += 3 + (verify_checksum ? 1 : 0);
    ws->pending_n = n;
    return 0;
}

// Wait for the batch enqueued last on `stream`; returns 0, the first failing entry status, or ZK_ST_RETRY when the
// scratch was too small (ws->want_* then hold the exact needs: the caller re-enqueues the same batch).
int zk_decode_collect(ZkDecodeWs* ws, cudaStream_t stream, int32_t* status_out) {
    uint32_t n = ws->pending_n;
    if (n == 0) return 0;
    ZK_CUDA_OK(cudaStreamSynchronize(stream));
#ifndef ZK_EMUL
    if (cudaGetLastError() != cudaSuccess) return -(int)ZKZ_GENERIC;
#endif
    ws->pending_n = 0;
    if (ws->h_counters->overflow) {
        ws->want_blocks = (size_t)ws->h_counters->n_blocks; ws->want_lit = (size_t)ws->h_counters->n_lit; ws->want_seq = (size_t)ws->h_counters->n_seq;
        return ZK_ST_RETRY;
    }
    int worst = 0;
    for (uint32_t i = 0; i < n; i++) {
        int32_t st = ws->h_entries[i].status;
        if (st == ZK_ST_RETRY) st = -(int)ZKZ_MEMORY_ALLOCATION;
        if (status_out) status_out[i] = st;
        if (st && !worst) worst = st;
    }
    return worst;
}

int zk_decode_batch(ZkDecodeWs* ws, cudaStream_t stream, const uint8_t* d_comp, const uint64_t* c_off, const uint64_t* d_off,
                    uint32_t n, uint8_t* d_dst, int verify_checksum, int32_t* status_out, int exec_warps) {
    for (int attempt = 0; attempt < 2; attempt++) {
        int rc = zk_decode_enqueue(ws, stream, d_comp, c_off, d_off, n, d_dst, verify_checksum, exec_warps);
        if (rc) return rc;
        rc = zk_decode_collect(ws, stream, status_out);
        if (rc != ZK_ST_RETRY) return rc;
    }
    return -(int)ZKZ_MEMORY_ALLOCATION;
}
