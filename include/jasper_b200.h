/* jasper_b200.h -- C ABI of libjasper_b200.so, the B200-native drop-in for JASPER's polishing
 * hot path (the per-batch loop of src/jasper.py and the Jellyfish 2.3.0 query path under it).
 *
 * Plain C: opaque handles, plain pointers and sizes; no torch / CUDA types in any signature.
 * Every entry point returns 0 on success or a negative JB_E_* code; the message is available
 * from jb_last_error() (thread-local).  The caller owns all input buffers; the library owns
 * result objects until *_free.  One handle per GPU; calls on one handle are serialised by the
 * caller.  There is NO CPU fallback: without a CUDA device every compute entry point fails
 * with JB_E_CUDA.
 *
 * Reference interfaces replaced (JF! = jellyfish-2.3.0.tar.gz!jellyfish-2.3.0/):
 *   jb_db_open            QueryMerFile(const char*)          JF!swig/mer_file.i:18-36, :50
 *                         file_header / binary_query ctor    JF!include/jellyfish/file_header.hpp:35-107,
 *                                                            JF!include/jellyfish/binary_dumper.hpp:126-146
 *   jb_lookup             qf[jf.MerDNA(s).get_canonical()]   JF!swig/mer_file.i:41,56; JF!swig/mer_dna.i:15,40,71
 *                                                            (28 call sites in src/jasper.py, first :70-71)
 *   jb_count_positions    the same lookup for every window of a sequence (src/jasper.py:56,70-71,156-178)
 *   jb_polish             main() + iteration() + everything below   src/jasper.py:12-32, 35-137, 150-599
 *   jb_solid_threshold    src/jellyfish.py:1-22
 *   jb_allreduce_tallies  the awk sums over *qValCalcHelper.csv      src/jasper.sh:235-238
 */
#ifndef JASPER_B200_H
#define JASPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_OK 0
#define JB_E_ARG (-1)            /* bad argument                                              */
#define JB_E_IO (-2)             /* "Can't open file" etc. (reference: std::runtime_error)    */
#define JB_E_FORMAT (-3)         /* "Unsupported format" / malformed .jf                      */
#define JB_E_CUDA (-4)           /* no device / CUDA runtime error                            */
#define JB_E_NOMEM (-5)          /* table or work buffers do not fit                          */
#define JB_E_KMISMATCH (-6)      /* --ksize differs from the k recorded in the .jf header     */
#define JB_E_REFERENCE_CRASH (-7)/* input on which src/jasper.py raises and exits 1           */
#define JB_E_SCRATCH (-8)        /* a device work pool (paths / rows / output slack) overflowed */
#define JB_E_NCCL (-9)
#define JB_E_UNSUPPORTED (-10)   /* e.g. k > JB_MAX_K                                         */

/* hard-coded constants of the reference, named (src/jasper.py:14,20,532,533,543,125) */
#define JB_DIVISOR 50
#define JB_MIN_OVERLAP 5
#define JB_SLACK 10
#define JB_MAX_PATHS 5000
#define JB_FASTA_WIDTH 60
#define JB_MAX_K 48              /* wide 16-byte slots: 96 key bits + 32 count bits; the compact 8-byte layout covers
                                    every k whose remainder key_len - log2(size) fits (k = 37 always does)          */
#define JB_LAYOUT_WIDE 0         /* 16-byte slots {key, count}, two per 32-byte sector, linear probing             */
#define JB_LAYOUT_COMPACT 1      /* 8-byte slots {key >> r, tag, count}, four per 32-byte sector (default)          */

typedef struct jb_db jb_db;
typedef struct jb_batch jb_batch;
typedef struct jb_result jb_result;

typedef struct {
  int32_t k;                 /* key_len / 2 from the header (mer_file.i:23)        */
  int32_t key_bits;
  int32_t counter_len;       /* bytes per on-disk count                            */
  int32_t canonical;
  int32_t matrix_r, matrix_c, matrix_identity;
  int32_t device;
  uint64_t hash_size;        /* header "size"                                      */
  uint64_t n_records;
  uint64_t table_slots;      /* capacity incl. overflow tail                        */
  uint64_t table_bytes;
  uint64_t max_displacement; /* longest probe distance of any record, in slots      */
  double load_seconds;       /* wall time of jb_db_open                             */
  double build_kernel_ms;    /* device time spent in the table-build kernels        */
  int32_t layout;            /* JB_LAYOUT_*                                         */
  int32_t slot_bytes;        /* 8 or 16                                             */
  int32_t remainder_bits, tag_bits, count_bits;   /* compact slot fields            */
  int32_t n_parts;           /* 1, or the devices a partitioned table spans         */
  uint64_t home_units;       /* 32-byte sectors the hash maps to                    */
  uint64_t n_stash;          /* records kept in the stash of a compact table        */
  double load_factor;        /* n_records / (home_units * slots per unit)           */
  double bytes_per_record;   /* table_bytes / n_records                             */
  double read_seconds;       /* busiest reader thread's time inside read()          */
  uint64_t file_bytes;
  /* the scan index: the records once more, grouped by minimizer, for the every-window-of-a-contig access pattern (0 = not built) */
  uint64_t index_bytes;
  uint64_t index_buckets;
  int32_t index_mmer;     /* minimizer length */
  int32_t pad_;
  double index_seconds;
} jb_db_info_t;

typedef struct {
  const char* name;          /* NUL-terminated contig name (first header token)     */
  const char* seq;           /* bases, original case; need not be NUL-terminated    */
  uint64_t len;
} jb_seq_t;

/* One row of _iter<N>_<batch>.fix.csv (src/jasper.py:218-222): [name, coord, fixed_base, original].
 * `mutation` / `original` are the exact text the reference's csv.writer emits for columns 3 / 4
 * (including the ['X'] list repr of single-difference multi-error patches). */
typedef struct {
  uint32_t seq_index;
  int64_t coord;
  char mutation[64];
  char original[64];
} jb_fix_row_t;

/* ---- lifetime ---------------------------------------------------------------------------- */
int jb_init(int n_devices, const int* device_ids);
int jb_device_count(void);
const char* jb_last_error(void);
const char* jb_version(void);

/* ---- database ------------------------------------------------------------------------------ */
/* Environment knobs of the loader (all optional): JB_TABLE_LAYOUT=wide|compact, JB_TABLE_LOAD=<0.05..0.95> (default 0.5
 * compact / 0.5 wide, raised until the table fits), JB_READERS=<threads> (default 4), JB_ODIRECT=1 (bypass the page cache). */
int jb_db_open(const char* jf_path, int device, jb_db** out);
int jb_db_info(const jb_db* db, jb_db_info_t* out);
int jb_db_replicate(jb_db* src, int n, const int* devices, jb_db** outs);
/* The fallback for a table beyond one GPU's HBM (the reference's answer is mmap + page cache,
 * JF!include/jellyfish/mapped_file.hpp:30-50): the slots of ONE database partitioned over n devices of this process
 * (peer access required).  outs[i] lives on devices[i], owns part i and probes every part -- its own in HBM, the
 * others by 32-byte peer loads over NVLink / NVSwitch.  Every entry point that takes a jb_db works on such a handle
 * except jb_db_dump and jb_db_replicate.  Close every handle with jb_db_close. */
int jb_db_open_partitioned(const char* jf_path, int n, const int* devices, jb_db** outs);
/* The same with one PROCESS per GPU (torchrun): every rank streams the file and keeps the slots of part `rank` on its
 * device; the ranks then publish a handle each (CUDA IPC), the host side gathers the world * JB_PART_HANDLE_BYTES bytes
 * in rank order (e.g. torch.distributed.all_gather_object) and every rank attaches the other parts.  Until then the
 * handle refuses lookups and polishing. */
#define JB_PART_HANDLE_BYTES 128
int jb_db_open_part(const char* jf_path, int device, int rank, int world, jb_db** out);
int jb_db_part_handle(const jb_db* db, char* handle /* JB_PART_HANDLE_BYTES */);
int jb_db_attach_parts(jb_db* db, const char* handles /* world * JB_PART_HANDLE_BYTES, rank order */);
void jb_db_close(jb_db* db);
/* Dump the resident table back as (key_lo, key_hi, count) triples in slot order (tests: must
 * equal the file's record list).  `cap` = capacity of the three arrays in records. */
int jb_db_dump(const jb_db* db, uint64_t* key_lo, uint64_t* key_hi, uint32_t* count, uint64_t cap, uint64_t* n_out);
/* Count histogram straight from the resident table: hist[c] for c in [0, high+1], counts above
 * `high` lumped into hist[high+1] (JF!sub_commands/histo_main.cc:64-84). */
int jb_db_histogram(const jb_db* db, uint64_t high, uint64_t* hist);
/* src/jellyfish.py: threshold from histogram rows (count, freq), freq > 0 only, ascending.
 * Returns the threshold, or JB_E_ARG when the reference would exit 1 (threshold < 2 or no
 * local minimum). */
int jb_solid_threshold(const uint64_t* counts, const uint64_t* freqs, size_t n_rows);

/* ---- lookups ------------------------------------------------------------------------------- */
/* counts[i] = qf[jf.MerDNA(strs[i]).get_canonical()] (32-bit), for n NUL-terminated strings. */
int jb_lookup(const jb_db* db, const char* const* strs, size_t n, uint32_t* counts);
/* Same, strings given as a flat buffer + n+1 offsets (no NUL needed; bytes >= k are ignored). */
int jb_lookup_flat(const jb_db* db, const char* buf, const uint64_t* offsets, size_t n, uint32_t* counts);
/* counts[i] = C(seq[i:i+k]) for i in [0, len]; windows running off the end are 'A'-padded. */
int jb_count_positions(const jb_db* db, const char* seq, size_t len, uint32_t* counts);

/* ---- polishing ----------------------------------------------------------------------------- */
/* The whole per-batch loop of src/jasper.py main(): `passes` fixing iterations plus the final
 * count-only iteration.  `k` must equal the DB's k (the reference silently mis-reads otherwise;
 * we refuse).  Host buffers in, host-visible result out. */
int jb_polish(const jb_db* db, const jb_seq_t* seqs, size_t n_seqs, int k, uint32_t solid_thre,
              int passes, int fix, jb_result** out);

/* How many bases one jb_polish / jb_batch_upload call can take on this handle's device right now (free device
 * memory next to the resident table, minus the fixed scratch pools, over the per-base working set).  The host side
 * groups pieces into calls below this bound -- the role BATCH_SIZE plays for the reference's workers
 * (src/jasper.sh:132-139): it bounds the working set, it does not change any result. */
int jb_polish_capacity(const jb_db* db, uint64_t* max_bases);

/* Page-locked host memory.  jb_polish / jb_batch_upload recognise contigs that lie in page-locked memory (from here, from
 * cudaHostAlloc, or registered with cudaHostRegister) and copy them to the device without a staging copy. */
int jb_host_alloc(size_t bytes, void** out);
void jb_host_free(void* p);

/* Split form, for callers that keep contigs resident in HBM between calls (and for timing the
 * device part alone): upload once, polish (device only), then fetch. */
int jb_batch_upload(const jb_db* db, const jb_seq_t* seqs, size_t n_seqs, jb_batch** out);
int jb_batch_polish(jb_batch* batch, int k, uint32_t solid_thre, int passes, int fix);
int jb_batch_fetch(jb_batch* batch, jb_result** out);
void jb_batch_free(jb_batch* batch);

int jb_result_status(const jb_result* r);                 /* JB_OK or JB_E_REFERENCE_CRASH      */
int jb_result_iterations(const jb_result* r);             /* passes + 1                         */
int jb_result_tally(const jb_result* r, int ite, int64_t* wrong, int64_t* total);
int64_t jb_result_num_rows(const jb_result* r, int ite);
int jb_result_rows(const jb_result* r, int ite, jb_fix_row_t* rows, size_t cap);
int64_t jb_result_num_seqs(const jb_result* r);
int jb_result_seq(const jb_result* r, size_t seq_index, const char** data, uint64_t* len);
void jb_result_free(jb_result* r);

/* Device-side timings (CUDA events on the library's own stream) of the most recent call on
 * this thread: names "count_positions", "polish", "polish_scan", "polish_walk", "h2d", "d2h".
 * Returns milliseconds, or a negative value if unknown. */
double jb_last_timing_ms(const char* what);
/* Work counters of the most recent polish on this thread: "scan_lookups", "walk_lookups",
 * "kernel_launches", "events". */
int64_t jb_last_counter(const char* what);

/* ---- multi-GPU ----------------------------------------------------------------------------- */
/* NCCL plumbing for the only collective on the path: the sum of {wrong_0,total_0,wrong_p,total_p}
 * across ranks (signed).  Rank 0 makes an id, the host distributes its 128 bytes, every rank inits. */
int jb_comm_unique_id(char id[128]);
int jb_comm_init(int rank, int world, const char id[128], int device);
int jb_allreduce_tallies(int64_t tallies[4]);
void jb_comm_destroy(void);

/* ---- benchmarking aid: random 32-byte-sector gather bandwidth over the resident table ------- */
int jb_gather_bench(const jb_db* db, uint64_t n_probes, int sector_bytes, int iters, double* gbps, double* gprobes_per_s);

#ifdef __cplusplus
}
#endif
#endif /* JASPER_B200_H */
