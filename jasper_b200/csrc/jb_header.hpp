// jb_header.hpp -- parser for the header of a Jellyfish "binary/sorted" file.
// Layout (JF!include/jellyfish/generic_file_header.hpp:88-143,
// JF!include/jellyfish/file_header.hpp:35-107): 9 ASCII digits = hlen, then hlen bytes of
// compact JSON padded with NULs so that the records start 8-byte aligned.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace jb {

struct JfHeader {
  std::string format;
  int key_bits = 0;          // "key_len"
  int counter_len = 4;       // bytes
  bool canonical = false;
  uint64_t size = 0;         // hash size, power of two
  int matrix_r = 0, matrix_c = 0;
  bool matrix_identity = false;
  std::vector<uint64_t> columns;
  uint64_t data_offset = 0;  // 9 + hlen
  int k() const { return key_bits / 2; }
  int key_bytes() const { return key_bits / 8 + (key_bits % 8 != 0); }
  int record_bytes() const { return key_bytes() + counter_len; }
};

// Returns empty string on success, else an error message.
std::string parse_jf_header(const char* bytes, size_t n_bytes, JfHeader* out);

// Byte LUTs for pos = M * key: lut[b * 256 + v], b < key_bytes.  Identity => low r bits of key.
std::vector<uint64_t> build_hash_lut(const JfHeader& h);

}  // namespace jb
