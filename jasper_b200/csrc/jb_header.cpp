#include "jb_header.hpp"

#include <cctype>
#include <cstdlib>
#include <cstring>

namespace jb {
namespace {

// A small recursive-descent reader for the JSON subset Jellyfish writes (objects, arrays,
// strings, unsigned integers, true/false/null).  Unknown members are skipped.
class Json {
 public:
  Json(const char* p, const char* end) : p_(p), end_(end) {}
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }

  void ws() { while (p_ < end_ && (std::isspace((unsigned char)*p_) || *p_ == '\0')) ++p_; }
  char peek() { ws(); return p_ < end_ ? *p_ : '\0'; }
  bool eat(char c) {
    if (peek() == c) { ++p_; return true; }
    fail(std::string("expected '") + c + "'");
    return false;
  }
  bool try_eat(char c) { if (peek() == c) { ++p_; return true; } return false; }

  std::string str() {
    std::string s;
    if (!eat('"')) return s;
    while (p_ < end_ && *p_ != '"') {
      char c = *p_++;
      if (c == '\\' && p_ < end_) {
        char e = *p_++;
        switch (e) { case 'n': c = '\n'; break; case 't': c = '\t'; break; case 'r': c = '\r'; break;
                     case 'u': p_ += (end_ - p_ >= 4 ? 4 : end_ - p_); c = '?'; break; default: c = e; }
      }
      s.push_back(c);
    }
    if (p_ < end_) ++p_; else fail("unterminated string");
    return s;
  }
  uint64_t u64() {
    ws();
    if (end_ - p_ >= 4 && !std::strncmp(p_, "true", 4)) { p_ += 4; return 1; }
    if (end_ - p_ >= 5 && !std::strncmp(p_, "false", 5)) { p_ += 5; return 0; }
    if (p_ >= end_ || !std::isdigit((unsigned char)*p_)) { fail("expected number"); return 0; }
    uint64_t v = 0;
    while (p_ < end_ && std::isdigit((unsigned char)*p_)) v = v * 10 + (uint64_t)(*p_++ - '0');
    return v;
  }
  void skip() {
    char c = peek();
    if (c == '{') {
      ++p_;
      if (try_eat('}')) return;
      do { str(); eat(':'); skip(); } while (ok() && try_eat(','));
      eat('}');
    } else if (c == '[') {
      ++p_;
      if (try_eat(']')) return;
      do { skip(); } while (ok() && try_eat(','));
      eat(']');
    } else if (c == '"') {
      str();
    } else {
      while (p_ < end_ && !std::strchr(",]}", *p_) && !std::isspace((unsigned char)*p_)) ++p_;
    }
  }
  void fail(const std::string& m) { if (err_.empty()) err_ = m; }

 private:
  const char* p_;
  const char* end_;
  std::string err_;
};

void parse_matrix(Json& j, JfHeader* h) {
  if (!j.eat('{')) return;
  if (j.try_eat('}')) return;
  do {
    std::string key = j.str();
    j.eat(':');
    if (key == "r") h->matrix_r = (int)j.u64();
    else if (key == "c") h->matrix_c = (int)j.u64();
    else if (key == "identity") h->matrix_identity = j.u64() != 0;
    else if (key == "columns") {
      j.eat('[');
      if (!j.try_eat(']')) {
        do { h->columns.push_back(j.u64()); } while (j.ok() && j.try_eat(','));
        j.eat(']');
      }
    } else j.skip();
  } while (j.ok() && j.try_eat(','));
  j.eat('}');
}

}  // namespace

std::string parse_jf_header(const char* bytes, size_t n, JfHeader* h) {
  if (n < 9) return "file too short for a Jellyfish header";
  uint64_t hlen = 0;
  for (int i = 0; i < 9; ++i) {
    if (!std::isdigit((unsigned char)bytes[i])) return "not a Jellyfish file (no header length)";
    hlen = hlen * 10 + (uint64_t)(bytes[i] - '0');
  }
  if (9 + hlen > n) return "truncated Jellyfish header";
  h->data_offset = 9 + hlen;
  Json j(bytes + 9, bytes + 9 + hlen);
  if (!j.eat('{')) return "malformed header JSON";
  if (!j.try_eat('}')) {
    do {
      std::string key = j.str();
      j.eat(':');
      if (key == "format") h->format = j.str();
      else if (key == "key_len") h->key_bits = (int)j.u64();
      else if (key == "counter_len") h->counter_len = (int)j.u64();
      else if (key == "canonical") h->canonical = j.u64() != 0;
      else if (key == "size") h->size = j.u64();
      else if (key == "matrix1") parse_matrix(j, h);
      else j.skip();
    } while (j.ok() && j.try_eat(','));
    j.eat('}');
  }
  if (!j.ok()) return "malformed header JSON: " + j.error();
  if (h->format != "binary/sorted") return "Unsupported format '" + h->format + "'";
  if (h->key_bits <= 0 || (h->key_bits & 1)) return "bad key_len in header";
  if (h->counter_len < 1 || h->counter_len > 8) return "bad counter_len in header";
  if (h->size == 0 || (h->size & (h->size - 1))) return "header size is not a power of two";
  if (!h->matrix_identity && (int)h->columns.size() < h->key_bits) return "hash matrix has fewer columns than key bits";
  if (h->matrix_c == 0) h->matrix_c = h->key_bits;
  return "";
}

std::vector<uint64_t> build_hash_lut(const JfHeader& h) {
  const int nb = h.key_bytes();
  const int c = h.key_bits;
  std::vector<uint64_t> lut((size_t)nb * 256, 0);
  for (int b = 0; b < nb; ++b) {
    for (int v = 0; v < 256; ++v) {
      uint64_t x = 0;
      for (int bit = 0; bit < 8; ++bit) {
        int kb = b * 8 + bit;          // key bit index, 0 = LSB
        if (kb >= c || !((v >> bit) & 1)) continue;
        if (h.matrix_identity) { if (kb < 64) x ^= 1ull << kb; }
        else x ^= h.columns[(size_t)(c - 1 - kb)];
      }
      lut[(size_t)b * 256 + v] = x;
    }
  }
  return lut;
}

}  // namespace jb
