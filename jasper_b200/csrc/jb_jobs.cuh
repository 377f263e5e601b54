// jb_jobs.cuh -- copy jobs: how a fixing pass describes the new bases of a segment (written by the walker,
// consumed by k_concat and by the incremental recount).
#pragma once
#include <stdint.h>

namespace jb {

// The new bases of a segment are described as copy jobs: materialised windows around edits (src = the
// segment's staging region) and untouched stretches (src = the pass-start contig).  `rel` = offset inside
// the segment's new bases; k_concat adds the segment's offset in the new contig.
struct CopyJob { const uint8_t* src; int64_t rel; int64_t len; uint32_t slot; uint32_t pad; };
struct JobSink { CopyJob* jobs; unsigned long long cap; unsigned long long* used; };

#if defined(__CUDACC__)
// How the kernels that consume the job list (k_concat, k_recount_copy, k_recount_probe) share it out.  They are launched
// on a fixed 1-D grid BEFORE the host knows how many jobs the walkers wrote (no host round trip between k_walk and them),
// so the split is decided on the device from the job count: many short jobs (a fixing pass over a dirty assembly leaves
// ~3 per edit, a few hundred positions each) -> teams of kJobTeam threads per job; few long ones -> the job list is dealt
// to rows of `gx` CTAs, each row sharing one job at a time.
constexpr int kJobTeam = 64;
constexpr unsigned long long kJobsForTeams = 16384;
struct JobSplit { int64_t tid, nth; unsigned long long j0, jstep; };
__device__ __forceinline__ JobSplit job_split(unsigned long long n_jobs, bool force_rows, unsigned max_gx) {
  JobSplit r;
  if (n_jobs >= kJobsForTeams && !force_rows) {
    const unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    r.tid = (int64_t)(g % (unsigned)kJobTeam); r.nth = kJobTeam;
    r.j0 = g / (unsigned)kJobTeam; r.jstep = (unsigned long long)gridDim.x * blockDim.x / (unsigned)kJobTeam;
  } else {
    unsigned long long want = 65536 / (n_jobs ? n_jobs : 1);
    unsigned gx = (unsigned)(want < 4 ? 4 : (want > 128 ? 128 : want));
    if (gx > max_gx) gx = max_gx;
    if (gx > gridDim.x) gx = gridDim.x;
    const unsigned rows = gridDim.x / gx, row = blockIdx.x / gx;
    r.tid = (int64_t)(blockIdx.x % gx) * blockDim.x + threadIdx.x; r.nth = (int64_t)gx * blockDim.x;
    r.j0 = row < rows ? row : ~0ull; r.jstep = rows;              // (CTAs behind the last whole row have nothing to do)
  }
  return r;
}
#endif

}  // namespace jb
