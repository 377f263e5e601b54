// jb_api.cu -- library-wide plumbing: error strings, timing/counter registry, init, and the
// NCCL all-reduce of the QV tallies (the only collective on the path: the awk sums over
// *qValCalcHelper.csv of src/jasper.sh:235-238 become one ncclAllReduce(sum, int64 x 4)).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "jb_internal.cuh"

namespace jb {
namespace {
thread_local std::string t_error;
thread_local std::map<std::string, double> t_timing;
thread_local std::map<std::string, int64_t> t_counter;
}  // namespace

void set_error(const std::string& msg) { t_error = msg; }
int fail(int code, const std::string& msg) { t_error = msg; return code; }
void note_timing(const char* what, double ms) { t_timing[what] = ms; }
void note_counter(const char* what, int64_t v) { t_counter[what] = v; }
void add_counter(const char* what, int64_t v) { t_counter[what] += v; }
}  // namespace jb

using namespace jb;

extern "C" const char* jb_last_error(void) { return t_error.c_str(); }
extern "C" const char* jb_version(void) { return "jasper_b200 0.1 (sm_100a)"; }

extern "C" int jb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int jb_init(int n_devices, const int* device_ids) {
  int have = jb_device_count();
  if (have == 0) return fail(JB_E_CUDA, "no CUDA device: libjasper_b200 has no CPU path");
  for (int i = 0; i < n_devices; ++i) {
    int d = device_ids ? device_ids[i] : i;
    if (d < 0 || d >= have) return fail(JB_E_ARG, "jb_init: bad device index " + std::to_string(d));
    JB_CUDA(cudaSetDevice(d));
    JB_CUDA(cudaFree(0));
    // (no per-device limits: cudaLimitMaxL2FetchGranularity changes nothing here -- a random sector miss fetches its whole
    // 128-byte line at any setting, profiles/r1_gather_probe.log)
  }
  return JB_OK;
}

extern "C" double jb_last_timing_ms(const char* what) {
  auto it = t_timing.find(what ? what : "");
  return it == t_timing.end() ? -1.0 : it->second;
}
extern "C" int64_t jb_last_counter(const char* what) {
  auto it = t_counter.find(what ? what : "");
  return it == t_counter.end() ? -1 : it->second;
}

// ---- NCCL (resolved at run time so that the process uses whichever libnccl is already loaded) ----
namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, char[128], int) = nullptr;   // ncclUniqueId passed by value (128 bytes)
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
} g_nccl;
void* g_comm = nullptr;
int g_comm_device = 0;
int g_world = 1;
cudaStream_t g_comm_stream = nullptr;
int64_t* g_comm_buf = nullptr;

struct UniqueId { char bytes[128]; };

int load_nccl() {
  if (g_nccl.lib) return JB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) return fail(JB_E_NCCL, std::string("cannot load libnccl: ") + dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, char[128], int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return fail(JB_E_NCCL, "libnccl lacks the expected symbols");
  return JB_OK;
}
}  // namespace

extern "C" int jb_comm_unique_id(char id[128]) {
  if (int rc = load_nccl()) return rc;
  int r = g_nccl.GetUniqueId(id);
  return r ? fail(JB_E_NCCL, std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?")) : JB_OK;
}

extern "C" int jb_comm_init(int rank, int world, const char id[128], int device) {
  if (int rc = load_nccl()) return rc;
  JB_CUDA(cudaSetDevice(device));
  UniqueId uid;
  std::memcpy(uid.bytes, id, 128);
  // ncclCommInitRank(ncclComm_t*, int nranks, ncclUniqueId commId /* by value */, int rank)
  typedef int (*init_fn)(void**, int, UniqueId, int);
  init_fn f = (init_fn)dlsym(g_nccl.lib, "ncclCommInitRank");
  int r = f(&g_comm, world, uid, rank);
  if (r) return fail(JB_E_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  g_comm_device = device; g_world = world;
  JB_CUDA(cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking));
  JB_CUDA(cudaMalloc(&g_comm_buf, 4 * sizeof(int64_t)));
  return JB_OK;
}

extern "C" int jb_allreduce_tallies(int64_t tallies[4]) {
  if (!tallies) return fail(JB_E_ARG, "jb_allreduce_tallies: null argument");
  if (!g_comm) return g_world == 1 ? JB_OK : fail(JB_E_NCCL, "jb_comm_init has not been called");
  JB_CUDA(cudaSetDevice(g_comm_device));
  JB_CUDA(cudaMemcpyAsync(g_comm_buf, tallies, 32, cudaMemcpyHostToDevice, g_comm_stream));
  const int ncclInt64 = 4, ncclSum = 0;
  int r = g_nccl.AllReduce(g_comm_buf, g_comm_buf, 4, ncclInt64, ncclSum, g_comm, g_comm_stream);
  if (r) return fail(JB_E_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  JB_CUDA(cudaMemcpyAsync(tallies, g_comm_buf, 32, cudaMemcpyDeviceToHost, g_comm_stream));
  JB_CUDA(cudaStreamSynchronize(g_comm_stream));
  return JB_OK;
}

extern "C" void jb_comm_destroy(void) {
  if (g_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_comm);
  g_comm = nullptr;
  if (g_comm_buf) cudaFree(g_comm_buf);
  g_comm_buf = nullptr;
  if (g_comm_stream) cudaStreamDestroy(g_comm_stream);
  g_comm_stream = nullptr;
}
