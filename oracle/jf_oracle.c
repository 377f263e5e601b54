/* jf_oracle.c -- CPU restatement of the Jellyfish 2.3.0 lookup path that JASPER uses.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (jasper_b200/) may link,
 * load or call this file; it is the checker for tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg.
 *
 * Parity pin: tests/test_oracle.py checks every function here against the reference's
 * own compiled code (oracle/_ref, built from /root/reference/jellyfish-2.3.0.tar.gz) and
 * against the committed golden vectors under tests/golden/ that were produced by it.
 *
 * What is restated (JF! = jellyfish-2.3.0.tar.gz!jellyfish-2.3.0/):
 *   header          JF!include/jellyfish/generic_file_header.hpp:119-143, file_header.hpp:35-107
 *   MerDNA(char*)   JF!include/jellyfish/mer_dna.hpp:38-55 (codes), 525-542 (from_chars), 642-644
 *   canonical       JF!include/jellyfish/mer_dna.hpp:92-100, 227-250, 401-414, 428-431
 *   hash position   JF!include/jellyfish/rectangular_binary_matrix.hpp:223-261 (times_loop)
 *   lookup          JF!include/jellyfish/binary_dumper.hpp:148-189 (val_id), 191-212
 *   32-bit result   JF!swig/mer_file.i:41
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

typedef struct { uint64_t lo, hi; } mer_t;           /* 2k-bit unsigned integer, k <= 64 */

typedef struct jfo_db {
  int k;                    /* key_len / 2 */
  int key_bits;             /* key_len (bits) */
  int key_bytes;            /* ceil(key_len / 8) */
  int val_bytes;            /* counter_len */
  int canonical;
  uint64_t size;            /* hash size (power of two) */
  uint64_t mask;            /* size - 1 */
  int m_r, m_c, m_identity;
  uint64_t *columns;        /* m_c entries */
  size_t header_bytes;      /* 9 + hlen */
  size_t file_bytes;
  uint64_t n_records;
  size_t record_len;
  const unsigned char *map; /* whole file */
  const unsigned char *data;/* first record */
  char format[64];
  mer_t first_key, last_key;
  uint64_t first_pos, last_pos;
} jfo_db;

static char g_err[256];
const char *jfo_last_error(void) { return g_err; }

/* ---------------- tiny JSON scanner (only what the header needs) ---------------- */
typedef struct { const char *p, *end; int ok; } js_t;
static void js_ws(js_t *j) { while (j->p < j->end && isspace((unsigned char)*j->p)) j->p++; }
static int js_peek(js_t *j) { js_ws(j); return j->p < j->end ? *j->p : 0; }
static int js_eat(js_t *j, char c) { if (js_peek(j) == c) { j->p++; return 1; } j->ok = 0; return 0; }
static void js_string(js_t *j, char *out, size_t cap) {
  size_t n = 0;
  if (!js_eat(j, '"')) return;
  while (j->p < j->end && *j->p != '"') {
    char c = *j->p++;
    if (c == '\\' && j->p < j->end) c = *j->p++;
    if (out && n + 1 < cap) out[n++] = c;
  }
  if (out && cap) out[n] = 0;
  if (j->p < j->end) j->p++; else j->ok = 0;
}
static void js_skip(js_t *j);
static void js_skip_container(js_t *j, char open, char close) {
  js_eat(j, open);
  if (js_peek(j) == close) { j->p++; return; }
  while (j->ok) {
    if (open == '{') { js_string(j, NULL, 0); js_eat(j, ':'); }
    js_skip(j);
    if (js_peek(j) == ',') { j->p++; continue; }
    js_eat(j, close);
    return;
  }
}
static void js_skip(js_t *j) {
  int c = js_peek(j);
  if (c == '{') js_skip_container(j, '{', '}');
  else if (c == '[') js_skip_container(j, '[', ']');
  else if (c == '"') js_string(j, NULL, 0);
  else { while (j->p < j->end && !strchr(",]}", *j->p) && !isspace((unsigned char)*j->p)) j->p++; }
}
static uint64_t js_u64(js_t *j) {
  js_ws(j);
  if (j->p + 4 <= j->end && !strncmp(j->p, "true", 4)) { j->p += 4; return 1; }
  if (j->p + 5 <= j->end && !strncmp(j->p, "false", 5)) { j->p += 5; return 0; }
  char *e; uint64_t v = strtoull(j->p, &e, 10);
  if (e == j->p) j->ok = 0;
  j->p = e; return v;
}

static void parse_matrix(js_t *j, jfo_db *db) {
  char key[64];
  js_eat(j, '{');
  while (j->ok && js_peek(j) != '}') {
    js_string(j, key, sizeof key); js_eat(j, ':');
    if (!strcmp(key, "r")) db->m_r = (int)js_u64(j);
    else if (!strcmp(key, "c")) db->m_c = (int)js_u64(j);
    else if (!strcmp(key, "identity")) db->m_identity = (int)js_u64(j);
    else if (!strcmp(key, "columns")) {
      size_t cap = 16, n = 0; uint64_t *col = malloc(cap * sizeof *col);
      js_eat(j, '[');
      while (j->ok && js_peek(j) != ']') {
        if (n == cap) { cap *= 2; col = realloc(col, cap * sizeof *col); }
        col[n++] = js_u64(j);
        if (js_peek(j) == ',') j->p++;
      }
      js_eat(j, ']');
      db->columns = col;
      if (db->m_c == 0) db->m_c = (int)n;
    } else js_skip(j);
    if (js_peek(j) == ',') j->p++;
  }
  js_eat(j, '}');
}

/* ---------------- k-mer arithmetic ---------------- */
static inline int base_code(unsigned char c) {
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1;
               case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return -1; }
}
static inline mer_t mer_shl2_or(mer_t m, int code) {
  mer_t r; r.hi = (m.hi << 2) | (m.lo >> 62); r.lo = (m.lo << 2) | (uint64_t)code; return r;
}
static inline mer_t mer_mask(mer_t m, int bits) {
  if (bits >= 128) return m;
  if (bits >= 64) { m.hi &= (bits == 64) ? 0 : ((~0ULL) >> (128 - bits)); return m; }
  m.hi = 0; m.lo &= (bits == 0) ? 0 : ((~0ULL) >> (64 - bits)); return m;
}
static inline int mer_cmp(mer_t a, mer_t b) {
  if (a.hi != b.hi) return a.hi < b.hi ? -1 : 1;
  if (a.lo != b.lo) return a.lo < b.lo ? -1 : 1;
  return 0;
}
/* MerDNA(const char*): read at most k chars, stop at the first non-ACGTacgt byte (or NUL /
 * end of buffer); bases read occupy the HIGH end, the remaining low bits stay 0 ('A'). */
void jfo_mer_from_chars(const char *s, size_t len, int k, uint64_t out[2]) {
  mer_t m = {0, 0}; int n = 0;
  while (n < k && (size_t)n < len) {
    int c = base_code((unsigned char)s[n]);
    if (c < 0) break;
    m = mer_shl2_or(m, c); n++;
  }
  for (; n < k; n++) m = mer_shl2_or(m, 0);
  m = mer_mask(m, 2 * k);
  out[0] = m.lo; out[1] = m.hi;
}
void jfo_revcomp(const uint64_t in[2], int k, uint64_t out[2]) {
  mer_t m = {in[0], in[1]}, r = {0, 0};
  for (int i = 0; i < k; i++) {
    int c = (int)(m.lo & 3); m.lo = (m.lo >> 2) | (m.hi << 62); m.hi >>= 2;
    r = mer_shl2_or(r, 3 - c);
  }
  out[0] = r.lo; out[1] = r.hi;
}
void jfo_canonical(const uint64_t in[2], int k, uint64_t out[2]) {
  uint64_t rc[2]; jfo_revcomp(in, k, rc);
  mer_t a = {in[0], in[1]}, b = {rc[0], rc[1]};
  if (mer_cmp(b, a) < 0) { out[0] = rc[0]; out[1] = rc[1]; } else { out[0] = in[0]; out[1] = in[1]; }
}
void jfo_mer_to_string(const uint64_t in[2], int k, char *out) {
  mer_t m = {in[0], in[1]};
  for (int i = k - 1; i >= 0; i--) {
    out[i] = "ACGT"[m.lo & 3]; m.lo = (m.lo >> 2) | (m.hi << 62); m.hi >>= 2;
  }
  out[k] = 0;
}

/* pos = XOR over set bits b of key (b = 0 is the LSB) of columns[c-1-b]; identity => low bits */
uint64_t jfo_hash_pos(const jfo_db *db, const uint64_t key[2]) {
  if (db->m_identity || !db->columns) return key[0] & db->mask;
  uint64_t res = 0;
  for (int b = 0; b < db->m_c; b++) {
    uint64_t w = b < 64 ? key[0] >> b : key[1] >> (b - 64);
    if (w & 1) res ^= db->columns[db->m_c - 1 - b];
  }
  return res & db->mask;
}

/* ---------------- file ---------------- */
static mer_t rec_key(const jfo_db *db, uint64_t id) {
  unsigned char buf[16] = {0};
  memcpy(buf, db->data + id * db->record_len, (size_t)db->key_bytes);
  mer_t m; memcpy(&m.lo, buf, 8); memcpy(&m.hi, buf + 8, 8);
  return mer_mask(m, db->key_bits);
}
static uint64_t rec_val(const jfo_db *db, uint64_t id) {
  uint64_t v = 0;
  memcpy(&v, db->data + id * db->record_len + db->key_bytes, (size_t)db->val_bytes);
  return v;
}
static uint64_t key_pos(const jfo_db *db, mer_t m) { uint64_t k2[2] = {m.lo, m.hi}; return jfo_hash_pos(db, k2); }

void jfo_close(jfo_db *db) {
  if (!db) return;
  if (db->map) munmap((void *)db->map, db->file_bytes);
  free(db->columns); free(db);
}

jfo_db *jfo_open(const char *path) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) { snprintf(g_err, sizeof g_err, "Can't open file '%s'", path); return NULL; }
  struct stat st; fstat(fd, &st);
  jfo_db *db = calloc(1, sizeof *db);
  db->file_bytes = (size_t)st.st_size;
  db->map = mmap(NULL, db->file_bytes, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (db->map == MAP_FAILED || db->file_bytes < 9) { db->map = NULL; snprintf(g_err, sizeof g_err, "Can't map '%s'", path); jfo_close(db); return NULL; }
  char digits[10]; memcpy(digits, db->map, 9); digits[9] = 0;
  size_t hlen = strtoul(digits, NULL, 10);
  if (9 + hlen > db->file_bytes) { snprintf(g_err, sizeof g_err, "Truncated header"); jfo_close(db); return NULL; }
  db->header_bytes = 9 + hlen;
  js_t j = {(const char *)db->map + 9, (const char *)db->map + 9 + hlen, 1};
  char key[64];
  db->val_bytes = 4;
  js_eat(&j, '{');
  while (j.ok && js_peek(&j) != '}') {
    js_string(&j, key, sizeof key); js_eat(&j, ':');
    if (!strcmp(key, "key_len")) db->key_bits = (int)js_u64(&j);
    else if (!strcmp(key, "counter_len")) db->val_bytes = (int)js_u64(&j);
    else if (!strcmp(key, "size")) db->size = js_u64(&j);
    else if (!strcmp(key, "canonical")) db->canonical = (int)js_u64(&j);
    else if (!strcmp(key, "format")) js_string(&j, db->format, sizeof db->format);
    else if (!strcmp(key, "matrix1")) parse_matrix(&j, db);
    else js_skip(&j);
    if (js_peek(&j) == ',') j.p++;
  }
  if (!j.ok || db->key_bits <= 0 || db->key_bits > 128) { snprintf(g_err, sizeof g_err, "Bad header"); jfo_close(db); return NULL; }
  if (strcmp(db->format, "binary/sorted")) { snprintf(g_err, sizeof g_err, "Unsupported format '%s'", db->format); jfo_close(db); return NULL; }
  db->k = db->key_bits / 2;
  db->key_bytes = db->key_bits / 8 + (db->key_bits % 8 != 0);
  db->record_len = (size_t)(db->key_bytes + db->val_bytes);
  db->mask = db->size - 1;
  db->data = db->map + db->header_bytes;
  size_t body = db->file_bytes - db->header_bytes;
  if (body % db->record_len) { snprintf(g_err, sizeof g_err, "Size of database (%zu) must be a multiple of the length of a record (%zu)", body, db->record_len); jfo_close(db); return NULL; }
  db->n_records = body / db->record_len;
  if (db->n_records) {
    db->first_key = rec_key(db, 0); db->first_pos = key_pos(db, db->first_key);
    db->last_key = rec_key(db, db->n_records - 1); db->last_pos = key_pos(db, db->last_key);
  }
  return db;
}

int jfo_k(const jfo_db *db) { return db->k; }
uint64_t jfo_n_records(const jfo_db *db) { return db->n_records; }
uint64_t jfo_size(const jfo_db *db) { return db->size; }
int jfo_counter_len(const jfo_db *db) { return db->val_bytes; }
int jfo_matrix_r(const jfo_db *db) { return db->m_r; }
void jfo_record(const jfo_db *db, uint64_t id, uint64_t key[2], uint64_t *val) {
  mer_t m = rec_key(db, id); key[0] = m.lo; key[1] = m.hi; *val = rec_val(db, id);
}

/* val_id: guided (interpolation) binary search on (pos, key), then a short linear scan. */
uint64_t jfo_lookup_key(const jfo_db *db, const uint64_t key2[2]) {
  if (db->n_records == 0) return 0;
  mer_t key = {key2[0], key2[1]};
  uint64_t first = 0, last = db->n_records, first_pos = db->first_pos, last_pos = db->last_pos;
  const uint64_t pos = key_pos(db, key);
  uint64_t cid = 0;
  if (!mer_cmp(key, db->first_key)) return rec_val(db, 0);
  cid = db->n_records - 1;
  if (!mer_cmp(key, db->last_key)) return rec_val(db, cid);
  if (pos < db->first_pos || pos > db->last_pos) return 0;
  for (uint64_t diff = last - first; diff >= 8; diff = last - first) {
    cid = first + (uint64_t)lrint((double)diff * ((double)(pos - first_pos) / (double)(last_pos - first_pos)));
    if (cid < first + 1) cid = first + 1;
    if (cid > last - 1) cid = last - 1;
    mer_t mid = rec_key(db, cid);
    int c = mer_cmp(key, mid);
    if (!c) return rec_val(db, cid);
    uint64_t mid_pos = key_pos(db, mid);
    if (mid_pos > pos || (mid_pos == pos && c < 0)) { last = cid; last_pos = mid_pos; }
    else { first = cid; first_pos = mid_pos; }
  }
  for (cid = first + 1; cid < last; ++cid)
    if (!mer_cmp(key, rec_key(db, cid))) return rec_val(db, cid);
  return 0;
}

/* C(s) := qf[jf.MerDNA(s).get_canonical()] truncated to unsigned int */
uint32_t jfo_count(const jfo_db *db, const char *s, size_t len) {
  uint64_t m[2], c[2];
  jfo_mer_from_chars(s, len, db->k, m);
  jfo_canonical(m, db->k, c);
  return (uint32_t)jfo_lookup_key(db, c);
}
void jfo_count_many(const jfo_db *db, const char *const *strs, size_t n, uint32_t *out) {
  for (size_t i = 0; i < n; i++) out[i] = jfo_count(db, strs[i], strlen(strs[i]));
}
/* out[i] = C(seq[i:i+k]) for i in [0, len] (the window may run off the end => 'A' padding) */
void jfo_count_positions(const jfo_db *db, const char *seq, size_t len, uint32_t *out) {
  for (size_t i = 0; i <= len; i++) out[i] = jfo_count(db, seq + i, len - i);
}
