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

}  // namespace jb
