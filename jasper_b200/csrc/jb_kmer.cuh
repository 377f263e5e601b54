// jb_kmer.cuh -- 2-bit k-mer arithmetic shared by every kernel (and by the host-side
// differential build of the walker).  Semantics follow Jellyfish 2.3.0:
//   codes / from_chars    JF!include/jellyfish/mer_dna.hpp:38-55, 525-542, 642-644
//   reverse complement    JF!include/jellyfish/mer_dna.hpp:92-100, 401-414
//   operator< / canonical JF!include/jellyfish/mer_dna.hpp:227-250, 428-431
//   hash position         JF!include/jellyfish/rectangular_binary_matrix.hpp:223-261
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define JB_HD __host__ __device__ __forceinline__
#define JB_D __device__ __forceinline__
#else
#define JB_HD inline
#define JB_D inline
#endif

namespace jb {

// A 2k-bit unsigned integer, first base most significant; k <= 48 so hi < 2^32.
struct Key {
  uint64_t lo, hi;
};

JB_HD bool key_less(const Key& a, const Key& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
JB_HD bool key_eq(const Key& a, const Key& b) { return a.lo == b.lo && a.hi == b.hi; }

// A,a=0 C,c=1 G,g=2 T,t=3; everything else (N, IUPAC, NUL, '\n', '-') is invalid.
JB_HD bool base_valid(uint8_t c) {
  const uint32_t x = (uint32_t)(c & 0xDF) - 0x41u;       // 'A' -> 0, 'C' -> 2, 'G' -> 6, 'T' -> 19
  return x < 20u && ((0x80045u >> x) & 1u) != 0u;
}
JB_HD uint32_t base_code(uint8_t c) { return ((c >> 1) ^ (c >> 2)) & 3u; }

JB_HD Key key_mask(int k) {
  Key m;
  int bits = 2 * k;
  m.lo = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
  m.hi = bits <= 64 ? 0ull : (bits >= 128 ? ~0ull : ((1ull << (bits - 64)) - 1));
  return m;
}

// (key << 2 | code) & mask
JB_HD Key key_push(Key key, uint32_t code, const Key& mask) {
  Key r;
  r.hi = ((key.hi << 2) | (key.lo >> 62)) & mask.hi;
  r.lo = ((key.lo << 2) | code) & mask.lo;
  return r;
}

// key << (2*n) for 0 <= n <= 64 bases (used to 'A'-pad a short prefix), masked.
JB_HD Key key_shl_bases(Key key, int n, const Key& mask) {
  int s = 2 * n;
  Key r = key;
  if (s >= 128) { r.lo = 0; r.hi = 0; }
  else if (s >= 64) { r.hi = key.lo << (s - 64); r.lo = 0; }
  else if (s > 0) { r.hi = (key.hi << s) | (key.lo >> (64 - s)); r.lo = key.lo << s; }
  r.lo &= mask.lo; r.hi &= mask.hi;
  return r;
}

JB_HD uint64_t rev2_64(uint64_t x) {   // reverse the order of the 32 two-bit groups
#if defined(__CUDA_ARCH__)
  x = __brevll(x);
#else
  x = ((x >> 32) | (x << 32));
  x = ((x & 0xFFFF0000FFFF0000ull) >> 16) | ((x & 0x0000FFFF0000FFFFull) << 16);
  x = ((x & 0xFF00FF00FF00FF00ull) >> 8) | ((x & 0x00FF00FF00FF00FFull) << 8);
  x = ((x & 0xF0F0F0F0F0F0F0F0ull) >> 4) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = ((x & 0xCCCCCCCCCCCCCCCCull) >> 2) | ((x & 0x3333333333333333ull) << 2);
  x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
#endif
  // brev reversed single bits; swap the two bits inside every group back
  return ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
}

JB_HD Key key_revcomp(const Key& key, int k) {
  // complement, reverse all 64 groups of the 128-bit value, then drop the (64-k) low groups
  uint64_t a = rev2_64(~key.lo);   // becomes the high word
  uint64_t b = rev2_64(~key.hi);   // becomes the low word
  int s = 128 - 2 * k;             // 32 <= s < 128 for 1 <= k <= 48
  Key r;
  if (s >= 64) { r.lo = a >> (s - 64); r.hi = 0; if (s == 64) r.lo = a; }
  else { r.lo = (b >> s) | (a << (64 - s)); r.hi = a >> s; }
  return r;
}

JB_HD Key key_canonical(const Key& key, int k) {
  Key rc = key_revcomp(key, k);
  return key_less(rc, key) ? rc : key;
}

// MerDNA(const char*): at most k bytes, stop at the first invalid byte or at `len`;
// bases read occupy the high end, remaining low bits are 0 ('A').
JB_HD Key key_from_chars(const uint8_t* s, int64_t len, int k, const Key& mask) {
  Key key = {0, 0};
  int n = 0;
  while (n < k && n < len) {
    uint8_t c = s[n];
    if (!base_valid(c)) break;
    key = key_push(key, base_code(c), mask);
    ++n;
  }
  return key_shl_bases(key, k - n, mask);
}

// ---- hash position --------------------------------------------------------------------------
// pos = XOR over set bits b of key (b = 0 is the LSB) of columns[c-1-b], as byte LUTs:
// lut[byte_index * 256 + byte_value].  The caller masks with (size - 1).
constexpr int kMaxKeyBytes = 12;       // 96 bits
JB_HD uint64_t hash_lut(const uint64_t* lut, int nbytes, const Key& key) {
  uint64_t h = 0;
#pragma unroll
  for (int b = 0; b < kMaxKeyBytes; ++b) {
    if (b < nbytes) {
      uint32_t v = b < 8 ? (uint32_t)(key.lo >> (8 * b)) & 0xFFu : (uint32_t)(key.hi >> (8 * (b - 8))) & 0xFFu;
      h ^= lut[b * 256 + v];
    }
  }
  return h;
}

// ---- table geometry -------------------------------------------------------------------------
// Two layouts share the probe unit -- one 32-byte sector, what a thread loads with one LDG.256 -- and the home function
//     unit(key) = floor(posx * units / 2^xbits),  posx = (pos << E) | (top E bits of key),
// which is monotone in the file order (pos, key), so the loader places records by a prefix-max scan (jb_table.cu).
//
//  compact (default): four 8-byte slots per sector.  Jellyfish draws its matrix so that the r x r block over the LOW r
//    key bits is invertible (JF!lib/rectangular_binary_matrix.cc:160-204, pseudo_inverse), so (pos, key >> r) identifies
//    the key and a slot only stores
//        [ remainder = key >> r | tag = low tag_bits of posx | count field of cnt_bits ]
//    The tag disambiguates posx among the sectors a record may have been displaced from (at most kMaxHops).  Count
//    field: bit cnt_bits-1 = "the run of occupied slots continues into the next sector" (kept in the LAST slot of a
//    sector), low cnt_bits-1 bits = cval: 0 empty, kCvalTomb / kCvalEsc reserved (record lives in the stash), else the
//    count.  Runs start at sector starts, so the occupied slots of a sector are a prefix: "last slot empty" = "has room".
//    A 128-byte DRAM line (the granule a random access costs here, tools/gather_probe.cu) holds 16 slots instead of 8.
//    (r2a tried whole-line buckets probed by four lanes per line: 1.02 lines per lookup, but 128 landing bytes per
//    probe in flight instead of 32 starve the memory-level parallelism: k_scan 6.0 ms against 2.2; profiles/README.md.)
//  wide (fallback: remainder too long for 64 bits, singular low block, or JB_TABLE_LAYOUT=wide): two 16-byte slots
//    {key.lo : 64, key.hi : 32, count : 32} per sector, count == 0 marks an empty slot, linear probing.
enum : int { kLayoutWide = 0, kLayoutCompact = 1 };
constexpr int kCompactSlots = 4;       // 8-byte slots per 32-byte sector
constexpr int kMaxHops = 7;            // a compact probe reads the home sector and at most this many more (load <= 0.65: 3-4 seen)
constexpr int kMaxCvalBits = 28;

struct TableGeom {
  uint64_t units;        // sectors the home function maps to
  uint64_t pos_mask;     // size - 1
  int k;
  int key_bytes;
  int ext_bits;          // E
  int left_shift;        // 64 - (r + E)
  int key_top_shift;     // 2k - E
  int layout;
  int r_bits;            // compact: remainder = key >> r_bits
  int tag_bits;
  int cnt_bits;
};

JB_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

JB_HD uint64_t key_top_bits(const Key& key, int shift) {   // key >> shift (shift < 128), low 64 bits
  if (shift >= 64) return key.hi >> (shift - 64);
  if (shift == 0) return key.lo;
  return (key.lo >> shift) | (key.hi << (64 - shift));
}

JB_HD uint64_t key_posx(const TableGeom& g, uint64_t pos, const Key& key) {
  uint64_t ext = g.ext_bits ? (key_top_bits(key, g.key_top_shift) & ((1ull << g.ext_bits) - 1)) : 0;
  return ((pos & g.pos_mask) << g.ext_bits) | ext;
}
JB_HD uint64_t home_unit_of_posx(const TableGeom& g, uint64_t posx) { return mulhi64(posx << g.left_shift, g.units); }
JB_HD uint64_t home_unit(const TableGeom& g, uint64_t pos, const Key& key) { return home_unit_of_posx(g, key_posx(g, pos, key)); }

// compact slot arithmetic
JB_HD uint32_t cval_mask(const TableGeom& g) { return (1u << (g.cnt_bits - 1)) - 1u; }
JB_HD uint32_t cval_esc(const TableGeom& g) { return cval_mask(g); }            // count too large for the field: see the stash
JB_HD uint32_t cval_tomb(const TableGeom& g) { return cval_mask(g) - 1u; }      // displaced beyond kMaxLineHops: see the stash
JB_HD uint64_t spill_bit(const TableGeom& g) { return 1ull << (g.cnt_bits - 1); }
// the slot a key would occupy, with an all-zero count field
JB_HD uint64_t compact_pattern(const TableGeom& g, uint64_t posx, const Key& key) {
  const uint64_t rem = key_top_bits(key, g.r_bits);
  const uint64_t tag = posx & ((1ull << g.tag_bits) - 1);
  return ((rem << g.tag_bits) | tag) << g.cnt_bits;
}

}  // namespace jb
