// ref_loi_harness.cpp -- drives the UNMODIFIED reference LOI.cpp as a library.
//
// TEST INFRASTRUCTURE ONLY (see oracle/hcspmm_oracle.c).  No reference source is
// copied: LOI.cpp is #included from where it lies (-I/root/reference) and only
// the MSVC-only stdio names it uses are mapped to their C equivalents on the
// compiler command line (oracle/Makefile).  The output goes to oracle/_ref/.
//
// LOI.cpp's main() hard-codes a dataset path and sizes (LOI.cpp:809,820), so the
// harness renames it and calls reorder_plus_new_direct (LOI.cpp:660-805) -- the
// variant main runs (LOI.cpp:848) -- on caller-provided CSR, building the CSC the
// way main does (LOI.cpp:826-841) and emitting vertices in main's file order
// (LOI.cpp:873-891).
#include <iostream>
#include <stdio.h>
#include <fstream>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include <set>
#include <map>
#include <omp.h>
#include <time.h>
#include <chrono>
// the reference prints a counter per block (LOI.cpp:688); keep test logs quiet
#define printf(...) ((void)0)
#define main loi_reference_main
#include "LOI.cpp"
#undef main
#undef printf

extern "C" int loi_ref_reorder(const int *rowptr, const int *colidx, int n,
                               int *perm_out, int *block_sizes_out,
                               int *n_blocks_out, int *n_full_out) {
  if (n > 18269000) return -3;  // struct_lst capacity, LOI.cpp:96
  int nnz = rowptr[n];
  std::vector<int> row_id(rowptr, rowptr + n + 1), col_id(colidx, colidx + nnz);
  std::vector<int> col_id_in(nnz), row_id_in(n + 1);
  for (auto x : col_id) row_id_in[x + 1]++;
  for (int i = 0; i < n; i++) row_id_in[i + 1] += row_id_in[i];
  std::vector<int> tmp_counts(row_id_in);
  for (int i = 0; i < n; i++)
    for (int j = row_id[i]; j < row_id[i + 1]; j++) col_id_in[tmp_counts[col_id[j]]++] = i;
  std::vector<std::vector<int>> res;
  std::vector<bool> visit(n);
  reorder_plus_new_direct(row_id, col_id, n, res, visit, row_id_in, col_id_in);
  int o = 0, full = 0;
  for (auto &x : res)
    if (x.size() == 16) { full++; for (auto y : x) perm_out[o++] = y; }
  for (auto &x : res)
    if (x.size() < 16) for (auto y : x) perm_out[o++] = y;
  for (int i = 0; i < n; i++) if (!visit[i]) perm_out[o++] = i;
  if (block_sizes_out) for (size_t b = 0; b < res.size(); b++) block_sizes_out[b] = (int)res[b].size();
  if (n_blocks_out) *n_blocks_out = (int)res.size();
  if (n_full_out) *n_full_out = full;
  return o == n ? 0 : -2;
}
