// sparse_symbolic.h -- host-side symbolic analysis for batched sparse QPs
// (FBstabSparse: the reference's planned "general sparse matrix components",
// ROADMAP.md:10, over the LDL' interface its tools/qdldl/qdldl_wrapper.h:19-84
// sketches: analysis in the constructor, Factor, Solve).
//
// Every instance of a batch shares ONE sparsity pattern (H upper triangular CSC,
// G and A CSC); only the values differ.  The analysis therefore runs once per
// handle, on the host, and leaves integer tables in which nothing depends on the
// values -- so that on the device every lane of a warp (one instance per lane)
// follows the same control flow through the numeric factorisation:
//
//  * row forms of H (full symmetric), G and A: the mat-vecs gather, no atomics;
//  * the quasi-definite Newton matrix, order [z; l; w] (w = Gamma^1/2 A dz),
//        K = [ H + sigma I    G'        (Gamma^1/2 A)' ]
//            [ G             -sigma I    0             ]
//            [ Gamma^1/2 A    0         -I             ]
//    as a permuted upper-triangular CSC pattern whose entries name their source
//    (a value of H, G or A, a diagonal constant);
//  * a minimum-degree elimination order (K is quasi-definite: LDL' exists for
//    every symmetric permutation), the elimination tree and column counts
//    (QDLDL_etree), the pattern of L and the up-looking schedule QDLDL_factor
//    would discover at run time: for row k the columns it visits, in order, and
//    the slot of L each visit fills.
#pragma once

#include <string>
#include <vector>

namespace fbs {

// K entry sources
enum {
  KSRC_H = 0,        // Hx[idx]
  KSRC_H_SIGMA = 1,  // Hx[idx] + sigma (a stored diagonal entry of H)
  KSRC_SIGMA = 2,    // sigma (H has no stored diagonal entry in this column)
  KSRC_G = 3,        // Gx[idx]
  KSRC_NEG_SIGMA = 4,
  KSRC_A = 5,        // sqrt(Gamma[row]) * Ax[idx]
  KSRC_NEG_ONE = 6
};

struct SparsePattern {
  int nz = 0, nl = 0, nv = 0, n = 0;
  int nnzH = 0, nnzG = 0, nnzA = 0, nnzK = 0, nnzL = 0;
  std::vector<int> Hp, Hi, Gp, Gi, Ap, Ai;
  std::vector<int> Hr_ptr, Hr_col, Hr_val;
  std::vector<int> Gr_ptr, Gr_col, Gr_val;
  std::vector<int> Ar_ptr, Ar_col, Ar_val;
  std::vector<int> perm, iperm;  // perm[new] = old, iperm[old] = new
  std::vector<int> Kp, Ki, Kkind, Kidx, Krow;
  std::vector<int> etree, Lp, Li;
  std::vector<int> Sp, Sc, St;
  std::string error;
};

// Returns false (and sets out->error) on an invalid pattern.  user_perm (n
// entries, perm[new] = old) overrides the minimum-degree order when not null.
bool SparseAnalyze(int nz, int nl, int nv, const int* Hp, const int* Hi, const int* Gp,
                   const int* Gi, const int* Ap, const int* Ai, const int* user_perm,
                   SparsePattern* out);

}  // namespace fbs
