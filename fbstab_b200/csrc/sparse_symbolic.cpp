// sparse_symbolic.cpp -- see sparse_symbolic.h.
#include "sparse_symbolic.h"

#include <algorithm>
#include <set>

namespace fbs {
namespace {

bool CheckCsc(const char* name, int rows, int cols, const int* p, const int* i, bool upper,
              std::string* err) {
  if (!p || (p[cols] > 0 && !i)) {
    *err = std::string(name) + ": null pattern array";
    return false;
  }
  if (p[0] != 0) {
    *err = std::string(name) + ": column pointers must start at 0";
    return false;
  }
  for (int c = 0; c < cols; c++) {
    if (p[c + 1] < p[c]) {
      *err = std::string(name) + ": column pointers must be non-decreasing";
      return false;
    }
    for (int e = p[c]; e < p[c + 1]; e++) {
      if (i[e] < 0 || i[e] >= rows) {
        *err = std::string(name) + ": row index out of range";
        return false;
      }
      if (e > p[c] && i[e] <= i[e - 1]) {
        *err = std::string(name) + ": row indices must be strictly increasing in a column";
        return false;
      }
      if (upper && i[e] > c) {
        *err = std::string(name) + ": only the upper triangle may be stored";
        return false;
      }
    }
  }
  return true;
}

// rows of a CSC matrix, each sorted by column; val = index into the CSC value array
void RowForm(int rows, int cols, const std::vector<int>& p, const std::vector<int>& i,
             std::vector<int>* rp, std::vector<int>* rc, std::vector<int>* rv) {
  rp->assign(rows + 1, 0);
  for (int e = 0; e < p[cols]; e++) (*rp)[i[e] + 1]++;
  for (int r = 0; r < rows; r++) (*rp)[r + 1] += (*rp)[r];
  rc->resize(p[cols]);
  rv->resize(p[cols]);
  std::vector<int> next(rp->begin(), rp->end() - 1);
  for (int c = 0; c < cols; c++)  // ascending columns -> rows come out sorted
    for (int e = p[c]; e < p[c + 1]; e++) {
      const int q = next[i[e]]++;
      (*rc)[q] = c;
      (*rv)[q] = e;
    }
}

struct Entry {
  int row, kind, idx, krow;
};

}  // namespace

bool SparseAnalyze(int nz, int nl, int nv, const int* Hp, const int* Hi, const int* Gp,
                   const int* Gi, const int* Ap, const int* Ai, const int* user_perm,
                   SparsePattern* out) {
  SparsePattern& s = *out;
  s = SparsePattern();
  if (nz <= 0 || nv <= 0 || nl < 0) {  // sizes as FBstabDense (fbstab_dense.cc:18-27)
    s.error = "sizes must satisfy nz > 0, nv > 0, nl >= 0";
    return false;
  }
  if (!CheckCsc("H", nz, nz, Hp, Hi, true, &s.error)) return false;
  if (nl > 0 && !CheckCsc("G", nl, nz, Gp, Gi, false, &s.error)) return false;
  if (!CheckCsc("A", nv, nz, Ap, Ai, false, &s.error)) return false;
  s.nz = nz;
  s.nl = nl;
  s.nv = nv;
  s.n = nz + nl + nv;
  s.Hp.assign(Hp, Hp + nz + 1);
  s.Hi.assign(Hi, Hi + Hp[nz]);
  if (nl > 0) {
    s.Gp.assign(Gp, Gp + nz + 1);
    s.Gi.assign(Gi, Gi + Gp[nz]);
  } else {
    s.Gp.assign(nz + 1, 0);
  }
  s.Ap.assign(Ap, Ap + nz + 1);
  s.Ai.assign(Ai, Ai + Ap[nz]);
  s.nnzH = s.Hp[nz];
  s.nnzG = s.Gp[nz];
  s.nnzA = s.Ap[nz];

  // ---- row forms ---------------------------------------------------------------
  RowForm(nl, nz, s.Gp, s.Gi, &s.Gr_ptr, &s.Gr_col, &s.Gr_val);
  RowForm(nv, nz, s.Ap, s.Ai, &s.Ar_ptr, &s.Ar_col, &s.Ar_val);
  {  // H: full symmetric expansion of the stored upper triangle, rows sorted by column
    std::vector<std::vector<std::pair<int, int>>> rows(nz);
    for (int c = 0; c < nz; c++)
      for (int e = s.Hp[c]; e < s.Hp[c + 1]; e++) {
        const int r = s.Hi[e];
        rows[r].push_back({c, e});
        if (r != c) rows[c].push_back({r, e});
      }
    s.Hr_ptr.assign(nz + 1, 0);
    for (int r = 0; r < nz; r++) {
      std::sort(rows[r].begin(), rows[r].end());
      s.Hr_ptr[r + 1] = s.Hr_ptr[r] + (int)rows[r].size();
      for (auto& pr : rows[r]) {
        s.Hr_col.push_back(pr.first);
        s.Hr_val.push_back(pr.second);
      }
    }
  }

  // ---- K in the natural order [z; l; w], upper triangle, by columns --------------
  const int n = s.n;
  std::vector<std::vector<Entry>> cols(n);
  for (int c = 0; c < nz; c++) {
    bool diag = false;
    for (int e = s.Hp[c]; e < s.Hp[c + 1]; e++) {
      if (s.Hi[e] == c) {
        cols[c].push_back({c, KSRC_H_SIGMA, e, 0});
        diag = true;
      } else {
        cols[c].push_back({s.Hi[e], KSRC_H, e, 0});
      }
    }
    if (!diag) cols[c].push_back({c, KSRC_SIGMA, 0, 0});
  }
  for (int r = 0; r < nl; r++) {
    const int c = nz + r;
    for (int q = s.Gr_ptr[r]; q < s.Gr_ptr[r + 1]; q++)
      cols[c].push_back({s.Gr_col[q], KSRC_G, s.Gr_val[q], 0});
    cols[c].push_back({c, KSRC_NEG_SIGMA, 0, 0});
  }
  for (int k = 0; k < nv; k++) {
    const int c = nz + nl + k;
    for (int q = s.Ar_ptr[k]; q < s.Ar_ptr[k + 1]; q++)
      cols[c].push_back({s.Ar_col[q], KSRC_A, s.Ar_val[q], k});
    cols[c].push_back({c, KSRC_NEG_ONE, 0, 0});
  }

  // ---- elimination order: greedy minimum degree on the graph of K -----------------
  s.perm.resize(n);
  s.iperm.resize(n);
  if (user_perm) {
    std::vector<char> seen(n, 0);
    for (int k = 0; k < n; k++) {
      if (user_perm[k] < 0 || user_perm[k] >= n || seen[user_perm[k]]) {
        s.error = "the elimination order is not a permutation";
        return false;
      }
      seen[user_perm[k]] = 1;
      s.perm[k] = user_perm[k];
    }
  } else {
    std::vector<std::set<int>> adj(n);
    for (int c = 0; c < n; c++)
      for (const Entry& e : cols[c])
        if (e.row != c) {
          adj[c].insert(e.row);
          adj[e.row].insert(c);
        }
    // Three phases, mirroring the reduction of dense_cholesky_solver.cc:32-127: the w
    // block first (pivots -1: this forms E = H + sigma I + A' Gamma A, whatever H's own
    // pivots are), then z, then l (the Schur complement -sigma I - G E^-1 G').  Any order
    // has an LDL', but eliminating a z variable whose pivot is only sigma = 1e-8 BEFORE
    // the active constraints that bound it squares the 1e8-scale entries and loses the
    // step (the reference's InfeasibleQP test then ends in MAXITERATIONS instead of
    // PRIMAL_INFEASIBLE).  Minimum degree picks the order inside a phase.
    std::vector<char> alive(n, 1);
    auto phase_of = [&](int i) { return i >= nz + nl ? 0 : (i < nz ? 1 : 2); };
    int phase = 0, left[3] = {nv, nz, nl};
    for (int step = 0; step < n; step++) {
      while (left[phase] == 0) phase++;
      int best = -1;
      size_t deg = 0;
      for (int i = 0; i < n; i++)
        if (alive[i] && phase_of(i) == phase && (best < 0 || adj[i].size() < deg)) {
          best = i;
          deg = adj[i].size();
        }
      left[phase]--;
      s.perm[step] = best;
      alive[best] = 0;
      const std::vector<int> nb(adj[best].begin(), adj[best].end());
      for (int u : nb) adj[u].erase(best);
      for (size_t a = 0; a < nb.size(); a++)
        for (size_t b = a + 1; b < nb.size(); b++) {
          adj[nb[a]].insert(nb[b]);
          adj[nb[b]].insert(nb[a]);
        }
      adj[best].clear();
    }
  }
  for (int k = 0; k < n; k++) s.iperm[s.perm[k]] = k;

  // ---- permuted upper-triangular CSC of K ----------------------------------------
  {
    std::vector<std::vector<Entry>> pc(n);
    for (int c = 0; c < n; c++)
      for (const Entry& e : cols[c]) {
        const int a = s.iperm[e.row], b = s.iperm[c];
        Entry t = e;
        t.row = std::min(a, b);
        pc[std::max(a, b)].push_back(t);
      }
    s.Kp.assign(n + 1, 0);
    for (int c = 0; c < n; c++) {
      std::sort(pc[c].begin(), pc[c].end(),
                [](const Entry& x, const Entry& y) { return x.row < y.row; });
      s.Kp[c + 1] = s.Kp[c] + (int)pc[c].size();
      for (const Entry& e : pc[c]) {
        s.Ki.push_back(e.row);
        s.Kkind.push_back(e.kind);
        s.Kidx.push_back(e.idx);
        s.Krow.push_back(e.krow);
      }
    }
    s.nnzK = s.Kp[n];
  }

  // ---- elimination tree and column counts (QDLDL_etree) ---------------------------
  std::vector<int> Lnz(n, 0), work(n, 0);
  s.etree.assign(n, -1);
  for (int j = 0; j < n; j++) {
    work[j] = j;
    for (int p = s.Kp[j]; p < s.Kp[j + 1]; p++) {
      int i = s.Ki[p];
      while (work[i] != j) {
        if (s.etree[i] == -1) s.etree[i] = j;
        Lnz[i]++;
        work[i] = j;
        i = s.etree[i];
      }
    }
  }
  s.Lp.assign(n + 1, 0);
  for (int i = 0; i < n; i++) s.Lp[i + 1] = s.Lp[i] + Lnz[i];
  s.nnzL = s.Lp[n];

  // ---- the pattern of L and the up-looking schedule (QDLDL_factor without values) --
  s.Li.assign(s.nnzL, 0);
  s.Sp.assign(n + 1, 0);
  std::vector<int> next(s.Lp.begin(), s.Lp.end() - 1), yIdx(n), elim(n);
  std::vector<char> mark(n, 0);
  for (int k = 0; k < n; k++) {
    int nnzY = 0;
    for (int p = s.Kp[k]; p < s.Kp[k + 1]; p++) {
      const int b = s.Ki[p];
      if (b == k) continue;
      int nx = b;
      if (!mark[nx]) {
        mark[nx] = 1;
        elim[0] = nx;
        int nnzE = 1;
        nx = s.etree[b];
        while (nx != -1 && nx < k) {
          if (mark[nx]) break;
          mark[nx] = 1;
          elim[nnzE++] = nx;
          nx = s.etree[nx];
        }
        while (nnzE) yIdx[nnzY++] = elim[--nnzE];
      }
    }
    for (int i = nnzY - 1; i >= 0; i--) {
      const int c = yIdx[i];
      const int slot = next[c]++;
      s.Li[slot] = k;
      s.Sc.push_back(c);
      s.St.push_back(slot);
      mark[c] = 0;
    }
    s.Sp[k + 1] = (int)s.Sc.size();
  }
  return true;
}

}  // namespace fbs
