// pivots.cuh — structural pivot search (pivots.cu)
#pragma once
#include "factor.cuh"

namespace sb {

struct PivotSearch {
  DBuf<int> pinv;  // [n] row -> pivot column or -1   (local to this round)
  DBuf<int> qinv;  // [m] column -> pivot row or -1
  DBuf<int> p;     // [n] pivotal rows first (topological order, N2), then the others increasing
  int npiv = 0;
  int max_height = 0;
  double t_fl = 0, t_flcol = 0, t_greedy = 0, t_reorder = 0;  // seconds (host clock around synchronised phases)
  int flcol_rounds = 0, greedy_windows = 0;
};

// counts = {FL, FL on columns, greedy}
int find_structural_pivots(const DCsr &A, bool greedy, PivotSearch &P, int counts[3]);
void extract_pivot_rows(const DCsr &A, const PivotSearch &P, DCsr &U, DBuf<int> &Uqinv, const Fp &F, DBuf<uint32_t> &pivval);

}  // namespace sb
