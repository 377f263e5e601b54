#include "jb_align.hpp"

#include <algorithm>
#include <cstdint>

namespace jb {
namespace {

constexpr int kMatch = 0, kMismatch = -1, kOpen = -1, kExtend = -1;

inline long gap_cost(long length, long open) {       // open + extend*length - extend, 0 for length <= 0
  return length <= 0 ? 0 : open + (long)kExtend * length - kExtend;
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
  const int W = lb + 1;
  std::vector<long> score((size_t)(la + 1) * W);
  std::vector<uint8_t> trace((size_t)(la + 1) * W, 0);
  const long first_gap = gap_cost(1, kOpen);
  for (int i = 0; i <= la; ++i) score[(size_t)i * W] = gap_cost(i, kOpen);
  for (int j = 0; j <= lb; ++j) score[j] = gap_cost(j, kOpen);
  std::vector<long> col_score(lb + 1, 0);
  for (int j = 1; j <= lb; ++j) col_score[j] = gap_cost(j, 2 * kOpen);
  for (int row = 1; row <= la; ++row) {
    long row_score = gap_cost(row, 2 * kOpen);
    for (int col = 1; col <= lb; ++col) {
      long nogap = score[(size_t)(row - 1) * W + col - 1] + (A[row - 1] == B[col - 1] ? kMatch : kMismatch);
      long row_open = score[(size_t)row * W + col - 1] + first_gap;
      long row_extend = row_score + kExtend;
      row_score = std::max(row_open, row_extend);
      long col_open = score[(size_t)(row - 1) * W + col] + first_gap;
      long col_extend = col_score[col] + kExtend;
      col_score[col] = std::max(col_open, col_extend);
      long best = std::max(nogap, std::max(col_score[col], row_score));
      score[(size_t)row * W + col] = best;
      int rt = (row_open == row_score ? 1 : 0) + (row_extend == row_score ? 8 : 0);
      int ct = (col_open == col_score[col] ? 4 : 0) + (col_extend == col_score[col] ? 16 : 0);
      int t = nogap == best ? 2 : 0;
      if (row_score == best) t += rt;
      if (col_score[col] == best) t += ct;
      trace[(size_t)row * W + col] = (uint8_t)t;
    }
  }
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
      Frame cache{a, b, row, col, col_gap, 0};
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
          const long target_score = sc(row, col);
          for (int n = 0; n < target; ++n) {
            if (in_a) { --col; a.push_back('-'); b.push_back(B[col]); }
            else { --row; a.push_back(A[row]); b.push_back('-'); }
            long actual = sc(row, col) + gap_cost(n + 1, kOpen);
            if (actual == target_score && n > 0) {
              if (!tr(row, col)) break;
              stack.push_back(Frame{a, b, row, col, col_gap, tr(row, col)});
            }
            if (!tr(row, col)) dead = true;
          }
        }
      }
      if (t) { cache.trace = t; stack.push_back(std::move(cache)); }
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
