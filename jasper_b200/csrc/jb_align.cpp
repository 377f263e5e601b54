#include "jb_align.hpp"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <new>

namespace jb {
namespace {

constexpr int kMatch = 0, kMismatch = -1, kOpen = -1, kExtend = -1;

// scores stay within +-(la + lb + 2): 32 bits are plenty (patches are capped at 256 K bases)
using score_t = int32_t;
inline score_t gap_cost(score_t length, score_t open) {       // open + extend*length - extend, 0 for length <= 0
  return length <= 0 ? 0 : open + (score_t)kExtend * length - kExtend;
}

struct Frame {
  std::string a, b;
  int row, col;
  bool col_gap;
  int trace;
};

}  // namespace

bool first_global_alignment(const std::string& A, const std::string& B, std::string* out_a, std::string* out_b) {
  const int la = (int)A.size(), lb = (int)B.size();
  if (la == 0 || lb == 0) return false;
  // Equal lengths and at most two mismatches -- nearly every bridged multi-error patch is two substitutions -- need no
  // matrix: a column with a gap costs as much as a mismatch and gaps come in pairs, so the gapless alignment (score -h)
  // is the only optimal one unless h == 2 and one deletion plus one insertion line everything else up
  // (A[m1+1 : m2+1] == B[m1 : m2] or A[m1 : m2] == B[m1+1 : m2+1]); only that tie needs the traceback order below.
  // (tests switch the shortcuts off through the environment; it is read per call, but only these two short names)
  if (la == lb && !getenv("JB_ALIGN_NO_FAST")) {
    int h = 0, m1 = -1, m2 = -1;
    for (int i = 0; i < la && h <= 2; ++i)
      if (A[i] != B[i]) { if (h == 0) m1 = i; else m2 = i; ++h; }
    bool unique = h <= 1;
    if (h == 2) unique = A.compare(m1 + 1, m2 - m1, B, m1, m2 - m1) != 0 && A.compare(m1, m2 - m1, B, m1 + 1, m2 - m1) != 0;
    if (unique) { *out_a = A; *out_b = B; return true; }
  }
  const int W = lb + 1;
  // 5 bytes per cell: a patch of tens of kb against its original would need gigabytes -- refuse cleanly (the row thread reports it)
  if ((uint64_t)(la + 1) * (uint64_t)W > (1ull << 31)) throw std::bad_alloc();
  // (thousands of short alignments per pass on the row threads: keep the matrices between calls)
  static thread_local std::vector<score_t> score_tl;
  static thread_local std::vector<uint8_t> trace_tl;
  static thread_local std::vector<score_t> col_score_tl;
  score_t* score = nullptr;
  uint8_t* trace = nullptr;
  // Only a band around the diagonal is computed.  With these penalties every gap character and every mismatch costs 1, so
  // a cell (i, j) lies on an optimal alignment of cost d only if |i - j| <= d, and the trace-back below only ever stands on
  // such cells (it moves from a cell to predecessors that achieve the cell's best score, starting at the end).  Cells
  // outside the band keep a score no path can reach and a trace value no branch acts on; in-band cells that depend on them
  // come out too low, never too high, and none of those can be on an optimal alignment while the band is wider than d.
  // So: half-width `band`, accept when the cost found is below it, else widen (the second round is the whole matrix).
  constexpr score_t kNever = -(1 << 28);
  int band = std::abs(la - lb) + 3;
  if (getenv("JB_ALIGN_NO_BAND")) band = la + lb + 1;
  // (thousands of these per pass on the row threads: the matrices are kept between calls and only the band, the cells that
  //  border it and the matrix borders inside it are written -- nothing else is ever read, see the gap-extension walk below)
  if (score_tl.size() < (size_t)(la + 1) * W) { score_tl.resize((size_t)(la + 1) * W); trace_tl.resize((size_t)(la + 1) * W); }
  if (col_score_tl.size() < (size_t)lb + 1) col_score_tl.resize((size_t)lb + 1);
  score = score_tl.data();                                 // (one TLS lookup each, not one per matrix access)
  trace = trace_tl.data();
  score_t* const col_score = col_score_tl.data();
  const score_t first_gap = gap_cost(1, kOpen);
  while (true) {
    const int edge = band + 1;                             // border cells up to here lie inside the band or next to it
    for (int i = 0; i <= la && i <= edge; ++i) { score[(size_t)i * W] = gap_cost(i, kOpen); trace[(size_t)i * W] = 0; }
    for (int j = 0; j <= lb && j <= edge; ++j) { score[j] = gap_cost(j, kOpen); trace[j] = 0; }
    for (int j = 1; j <= lb; ++j) col_score[j] = j <= band + 1 ? gap_cost(j, 2 * kOpen) : kNever;
    for (int row = 1; row <= la; ++row) {
      const int c_lo = row - band > 1 ? row - band : 1, c_hi = row + band < lb ? row + band : lb;
      score_t row_score = c_lo == 1 ? gap_cost(row, 2 * kOpen) : kNever;
      const score_t* up = score + (size_t)(row - 1) * W;   // (row pointers: the index arithmetic was a third of the loop)
      score_t* cur = score + (size_t)row * W;
      uint8_t* tcur = trace + (size_t)row * W;
      if (c_lo > 1) { cur[c_lo - 1] = kNever; tcur[c_lo - 1] = 32; }      // the cells next to the band: never reached,
      if (c_hi < lb) { cur[c_hi + 1] = kNever; tcur[c_hi + 1] = 32; }     // and no move bit set (32)
      const char ar = A[row - 1];
      for (int col = c_lo; col <= c_hi; ++col) {
        const score_t nogap = up[col - 1] + (ar == B[col - 1] ? kMatch : kMismatch);
        const score_t row_open = cur[col - 1] + first_gap;
        const score_t row_extend = row_score + kExtend;
        row_score = row_open > row_extend ? row_open : row_extend;
        const score_t col_open = up[col] + first_gap;
        const score_t col_extend = col_score[col] + kExtend;
        const score_t cs = col_open > col_extend ? col_open : col_extend;
        col_score[col] = cs;
        score_t best = nogap > cs ? nogap : cs;
        if (row_score > best) best = row_score;
        cur[col] = best;
        int t = nogap == best ? 2 : 0;
        if (row_score == best) t += (row_open == row_score ? 1 : 0) + (row_extend == row_score ? 8 : 0);
        if (cs == best) t += (col_open == cs ? 4 : 0) + (col_extend == cs ? 16 : 0);
        tcur[col] = (uint8_t)t;
      }
    }
    if (band > la + lb || -score[(size_t)la * W + lb] < band) break;
    band = la + lb + 1;
  }
  const int band_used = band;
  auto tr = [&](int r, int c) { return (int)trace[(size_t)r * W + c]; };
  auto sc = [&](int r, int c) { return score[(size_t)r * W + c]; };

  std::vector<Frame> stack;
  stack.push_back(Frame{"", "", la, lb, false, tr(la, lb)});
  while (!stack.empty()) {
    Frame f = std::move(stack.back());
    stack.pop_back();
    bool dead = false;
    std::string& a = f.a;
    std::string& b = f.b;
    int row = f.row, col = f.col, t = f.trace;
    bool col_gap = f.col_gap;
    while ((row > 0 || col > 0) && !dead) {
      // the state to come back to when this cell offers more than one way on: a and b only grow during the step, so
      // their lengths are enough to rebuild it (copying both strings at every step was most of the run time)
      const size_t a0 = a.size(), b0 = b.size();
      const int row0 = row, col0 = col;
      const bool col_gap0 = col_gap;
      if (!t) {
        if (col && col_gap) dead = true;
        else {                                        // _finish_backtrace
          for (int r = row - 1; r >= 0 && row; --r) a.push_back(A[r]);
          for (int c = col - 1; c >= 0 && col; --c) b.push_back(B[c]);
          if (row > col) b.append(a.size() - b.size(), '-');
          else if (col > row) a.append(b.size() - a.size(), '-');
        }
        break;
      } else if (t % 2 == 1) {                        // open a gap in A
        t -= 1;
        if (col_gap) dead = true;
        else { --col; a.push_back('-'); b.push_back(B[col]); col_gap = false; }
      } else if (t % 4 == 2) {                        // diagonal
        t -= 2;
        --row; --col;
        a.push_back(A[row]); b.push_back(B[col]);
        col_gap = false;
      } else if (t % 8 == 4) {                        // open a gap in B
        t -= 4;
        --row;
        a.push_back(A[row]); b.push_back('-');
        col_gap = true;
      } else if (t == 8 || t == 24 || t == 16) {      // extend a gap: find where it was opened
        const bool in_a = (t != 16);
        if (in_a) t -= 8; else t -= 16;
        if (in_a && col_gap) dead = true;
        else {
          col_gap = !in_a;
          const int target = in_a ? col : row;
          const score_t target_score = sc(row, col);
          for (int n = 0; n < target; ++n) {
            if (in_a) { --col; a.push_back('-'); b.push_back(B[col]); }
            else { --row; a.push_back(A[row]); b.push_back('-'); }
            // Outside the band nothing can match (such a cell is on no optimal alignment) and the walk would only run on
            // to the matrix border, where the trace value 0 kills this branch: same outcome, without reading cells that
            // were never written.
            if (col - row > band_used || row - col > band_used) { dead = true; break; }
            score_t actual = sc(row, col) + gap_cost(n + 1, kOpen);
            if (actual == target_score && n > 0) {
              if (!tr(row, col)) break;
              stack.push_back(Frame{a, b, row, col, col_gap, tr(row, col)});
            }
            if (!tr(row, col)) dead = true;
          }
        }
      }
      if (t) stack.push_back(Frame{a.substr(0, a0), b.substr(0, b0), row0, col0, col_gap0, t});
      t = tr(row, col);
    }
    if (!dead) {
      std::reverse(a.begin(), a.end());
      std::reverse(b.begin(), b.end());
      *out_a = a;
      *out_b = b;
      return true;
    }
  }
  return false;
}

}  // namespace jb
