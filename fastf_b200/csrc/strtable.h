// Exact-string -> 1-based index tables for cell barcodes and feature ids.
//
// Semantics follow the reference's chained table (reference src/hashtable.c:70-115, hash at
// src/bam2db_ds.c:96-104): membership is exact strcmp equality and the value is the 1-based index
// assigned at insertion.  Layout is B200-first instead of 2^20 pointer buckets: open addressing with
// 16-byte slots {tag, pool offset, length, value} in HBM (<= a few MB, L2 resident), probed by a whole
// warp: each lane hashes / compares one byte of the candidate string.
//
// Hash (host and device must agree): the string is cut into 32-byte chunks; chunk sum
// S = sum_j (c_j + 1) * PW[j] (mod 2^32) over the bytes present; H = H * Q + S over chunks; two
// independent (PW, Q) families give the slot hash and the verification tag.
#pragma once
#include <stdint.h>
#include <stddef.h>
#ifndef __CUDACC__
#ifndef FASTF_EMU
#define __host__
#define __device__
#endif
#endif

#define FASTF_H1_P 2654435761u
#define FASTF_H2_P 0x85ebca6bu
#define FASTF_H1_Q 0x9e3779b1u
#define FASTF_H2_Q 0xc2b2ae35u

struct FastfStrSlot {
    uint32_t tag;
    uint32_t off;     // into the string pool
    uint32_t len;
    uint32_t value;   // 0 = empty slot
};

struct FastfStrTableView {   // what kernels receive (device pointers)
    const FastfStrSlot *slots;
    const uint8_t *pool;
    uint32_t mask;           // capacity - 1 (capacity is a power of two)
    uint32_t pw1[32], pw2[32];
};

static inline __host__ __device__ uint32_t fastf_hash_finalize(uint32_t h)
{
    h ^= h >> 15;
    h *= 0x2c1b3c6du;
    h ^= h >> 12;
    h *= 0x297a2d39u;
    h ^= h >> 15;
    return h;
}

#include <string.h>
#include <vector>
struct FastfStrTableHost {
    std::vector<FastfStrSlot> slots;
    std::vector<uint8_t> pool;
    uint32_t mask = 0;
    uint32_t pw1[32], pw2[32];
    uint32_t count = 0;

    static void hashes(const uint8_t *s, size_t len, const uint32_t *pw1, const uint32_t *pw2, uint32_t *h1, uint32_t *h2)
    {
        uint32_t a = 0, b = 0;
        size_t c = 0;
        do {   // at least one (possibly empty) chunk, exactly like the device loop
            uint32_t s1 = 0, s2 = 0;
            for (size_t j = 0; j < 32 && c + j < len; j++) { s1 += ((uint32_t)s[c + j] + 1u) * pw1[j]; s2 += ((uint32_t)s[c + j] + 1u) * pw2[j]; }
            a = a * FASTF_H1_Q + s1;
            b = b * FASTF_H2_Q + s2;
            c += 32;
        } while (c < len);
        *h1 = a; *h2 = b;
    }
    void init(size_t n_expected)
    {
        uint32_t cap = 64;
        while (cap < n_expected * 2 + 8) cap <<= 1;
        slots.assign(cap, FastfStrSlot{0, 0, 0, 0});
        mask = cap - 1;
        pool.clear();
        count = 0;
        uint32_t p1 = 1, p2 = 1;
        for (int j = 0; j < 32; j++) { p1 *= FASTF_H1_P; p2 *= FASTF_H2_P; pw1[j] = p1; pw2[j] = p2; }
    }
    // returns the existing value (>0) if the key is present, else 0
    uint32_t find(const char *s, size_t len) const
    {
        uint32_t h1, h2;
        hashes((const uint8_t *)s, len, pw1, pw2, &h1, &h2);
        uint32_t i = fastf_hash_finalize(h1) & mask;
        while (slots[i].value) {
            if (slots[i].tag == h2 && slots[i].len == len && memcmp(pool.data() + slots[i].off, s, len) == 0) return slots[i].value;
            i = (i + 1) & mask;
        }
        return 0;
    }
    // inserts (key -> value) unless the key exists; returns true when inserted (reference hash_table_insert)
    bool insert(const char *s, size_t len, uint32_t value)
    {
        uint32_t h1, h2;
        hashes((const uint8_t *)s, len, pw1, pw2, &h1, &h2);
        uint32_t i = fastf_hash_finalize(h1) & mask;
        while (slots[i].value) {
            if (slots[i].tag == h2 && slots[i].len == len && memcmp(pool.data() + slots[i].off, s, len) == 0) return false;
            i = (i + 1) & mask;
        }
        while (pool.size() & 3u) pool.push_back(0);   // strings start on 4-byte boundaries: the device compares word-wise
        slots[i].tag = h2; slots[i].off = (uint32_t)pool.size(); slots[i].len = (uint32_t)len; slots[i].value = value;
        pool.insert(pool.end(), (const uint8_t *)s, (const uint8_t *)s + len);
        count++;
        return true;
    }
    void finish() { pool.resize((pool.size() + 64 + 15) / 16 * 16, 0); }   // slack so lane-wide compares never leave the pool
};
