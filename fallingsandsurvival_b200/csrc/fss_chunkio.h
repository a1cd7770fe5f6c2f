// fss_chunkio.h -- the reference's chunk file (Chunk::read / Chunk::write, FallingSandSurvival/Chunk.cpp:47-281) on the host side
// of the C ABI: the LZ4 block codec and the file framing around it.
//
//   int8  generationPhase
//   int32 src_size        = 128 * 128 * 2 * sizeof(MaterialInstanceData)   (the `tiles` layer, then `layer2`; Chunk.hpp:27-31)
//   int32 compressed_size
//   int32 src_size2       = 128 * 128 * sizeof(Uint32)                     (`background`)
//   int32 compressed_size2
//   LZ4 block (compressed_size bytes), LZ4 block (compressed_size2 bytes)
//
// The reference links liblz4 (LZ4_compress_fast(.., acceleration 10) / LZ4_decompress_safe, Chunk.cpp:127, :226); that library
// is a third-party dependency which is not part of this build, so the block format is implemented here from its published
// specification (lz4_Block_format.md): a block is a sequence of [token | literal length bytes | literals | offset (2 bytes,
// little endian) | match length bytes]; token high nibble = literal count, low nibble = match length - 4, 15 in a nibble
// = "more length bytes follow (each adds 0..255, 255 = one more)"; the last sequence has only literals; the last 5 bytes of
// a block are literals and the last match starts at least 12 bytes before the end.  Any conforming decoder reads what the
// encoder below writes, and the decoder below reads any conforming block: tests/test_chunkio.py checks both directions against
// the system's liblz4 and against the reference's own Chunk::write / Chunk::read lines (oracle/_ref).
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace fss_lz4 {

static inline size_t compress_bound(size_t n) { return n + n / 255 + 16; } // LZ4_compressBound

// greedy single-pass encoder with a 4-byte hash table (the structure of the format's reference "fast" encoder)
static inline size_t compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    if(cap < compress_bound(n)) return 0;
    const size_t MINMATCH = 4, MFLIMIT = 12, LASTLITERALS = 5;
    uint8_t* op = dst;
    size_t anchor = 0, ip = 0;
    auto put_len = [&](size_t len) { // the part of a length beyond the nibble's 15
        for(; len >= 255; len -= 255) *op++ = 255;
        *op++ = (uint8_t)len;
    };
    if(n >= MFLIMIT + 1) {
        std::vector<uint32_t> table(1u << 14, 0xFFFFFFFFu);
        auto rd32 = [&](size_t i) {
            uint32_t v;
            memcpy(&v, src + i, 4);
            return v;
        };
        const size_t mflimit = n - MFLIMIT, matchlimit = n - LASTLITERALS;
        while(ip < mflimit) {
            const uint32_t h = (rd32(ip) * 2654435761u) >> 18;
            const uint32_t cand = table[h];
            table[h] = (uint32_t)ip;
            if(cand == 0xFFFFFFFFu || ip - cand > 65535 || rd32(cand) != rd32(ip)) {
                ip++;
                continue;
            }
            size_t mlen = MINMATCH;
            while(ip + mlen < matchlimit && src[cand + mlen] == src[ip + mlen]) mlen++;
            const size_t lit = ip - anchor;
            uint8_t* token = op++;
            *token = (uint8_t)((lit >= 15 ? 15 : lit) << 4);
            if(lit >= 15) put_len(lit - 15);
            memcpy(op, src + anchor, lit);
            op += lit;
            const size_t off = ip - cand;
            *op++ = (uint8_t)(off & 0xFF);
            *op++ = (uint8_t)(off >> 8);
            const size_t m = mlen - MINMATCH;
            *token |= (uint8_t)(m >= 15 ? 15 : m);
            if(m >= 15) put_len(m - 15);
            ip += mlen;
            anchor = ip;
        }
    }
    const size_t lit = n - anchor; // the last sequence: literals only
    *op++ = (uint8_t)((lit >= 15 ? 15 : lit) << 4);
    if(lit >= 15) put_len(lit - 15);
    memcpy(op, src + anchor, lit);
    op += lit;
    return (size_t)(op - dst);
}

// LZ4_decompress_safe: never reads outside [src, src+n) nor writes outside [dst, dst+cap); returns the decoded size or -1
static inline long decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    size_t ip = 0, op = 0;
    if(n == 0) return -1;
    for(;;) {
        if(ip >= n) return -1;
        const unsigned token = src[ip++];
        size_t lit = token >> 4;
        if(lit == 15) {
            unsigned b;
            do {
                if(ip >= n) return -1;
                b = src[ip++];
                lit += b;
            } while(b == 255);
        }
        if(lit > n - ip || lit > cap - op) return -1;
        memcpy(dst + op, src + ip, lit);
        ip += lit;
        op += lit;
        if(ip == n) return (long)op; // the last sequence stops after its literals
        if(n - ip < 2) return -1;
        const size_t off = (size_t)src[ip] | ((size_t)src[ip + 1] << 8);
        ip += 2;
        if(off == 0 || off > op) return -1;
        size_t mlen = token & 15u;
        if(mlen == 15) {
            unsigned b;
            do {
                if(ip >= n) return -1;
                b = src[ip++];
                mlen += b;
            } while(b == 255);
        }
        mlen += 4;
        if(mlen > cap - op) return -1;
        for(size_t k = 0; k < mlen; k++, op++) dst[op] = dst[op - off]; // byte by byte: the match may overlap its own output
    }
}

} // namespace fss_lz4
