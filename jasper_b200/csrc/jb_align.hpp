// jb_align.hpp -- host-side description of a multi-error patch as fix.csv rows.
//
// The reference describes the patch found by base_extension with
//     pairwise2.align.globalms(fixed_seq, original, 0, -1, -1, -1)[0]     (src/jasper.py:309-329)
// i.e. unit-cost global alignment, FIRST traceback that Biopython completes.  This restates the
// published algorithm of Bio.pairwise2 (_make_score_matrix_fast + _recover_alignments +
// _find_gap_open): per-cell preference gap-in-A-open, diagonal, gap-in-B-open, then the two
// extensions, depth-first with a LIFO of alternatives, and the "gap in A right after a gap in B"
// dead end.  PARITY UNPINNED: Biopython is not available offline; the polished
// bases never depend on this, only the text of the rows.
#pragma once
#include <string>
#include <vector>

namespace jb {

// Returns false when either string is empty (Biopython returns [] and the reference crashes).
bool first_global_alignment(const std::string& a, const std::string& b, std::string* ali_a, std::string* ali_b);

}  // namespace jb
