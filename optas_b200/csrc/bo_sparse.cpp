// bo_sparse.cpp -- ordering + symbolic LDL' + device tables for the static-order sparse factorisation.
#include "bo_sparse.h"

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <set>

namespace bo {

int SparsePlan::pos(int i_old, int j_old) const {
  int a = iperm[i_old], b = iperm[j_old];
  if (a == b) return a;
  if (a < b) std::swap(a, b);  // a > b: entry (row a, column b) of L
  const auto lo = rowidx.begin() + colptr[b], hi = rowidx.begin() + colptr[b + 1];
  const auto it = std::lower_bound(lo, hi, a);
  if (it == hi || *it != a) return -1;
  return n + (int)(it - rowidx.begin());
}

SparsePlan make_sparse_plan(const ProblemSource& ps, bool large, int n_segments, int nd_depth, double nd_bias) {
  SparsePlan pl;
  const int nx = ps.nx, n = ps.nx + ps.n_eq;
  pl.n = n;
  pl.nx = nx;

  // ---- structural pattern of K (symmetric adjacency, no self loops) ----
  std::vector<std::set<int>> adj(n);
  auto link = [&](int a, int b) {
    if (a != b) {
      adj[a].insert(b);
      adj[b].insert(a);
    }
  };
  for (int k = 0; k < ps.hess.nnz(); ++k) link(ps.hess.row[k], ps.hess.col[k]);
  auto rows_clique = [&](const Sparsity& sp, int n_rows) {  // J' D J couples the columns of every row
    std::vector<std::vector<int>> by_row(n_rows > 0 ? n_rows : 1);
    for (int k = 0; k < sp.nnz(); ++k) by_row[sp.row[k]].push_back(sp.col[k]);
    for (const auto& cols : by_row)
      for (size_t u = 0; u < cols.size(); ++u)
        for (size_t w = u + 1; w < cols.size(); ++w) link(cols[u], cols[w]);
  };
  rows_clique(ps.jac_ineq, ps.n_ineq);
  rows_clique(ps.jac_eq, ps.n_eq);
  std::vector<std::vector<int>> x_of_y(ps.n_eq > 0 ? ps.n_eq : 1);
  for (int k = 0; k < ps.jac_eq.nnz(); ++k) {
    link(nx + ps.jac_eq.row[k], ps.jac_eq.col[k]);
    x_of_y[ps.jac_eq.row[k]].push_back(ps.jac_eq.col[k]);
  }

  // ---- constrained greedy minimum-degree ordering on the elimination graph ----
  std::vector<char> done(n, 0);
  std::vector<int> pending_x(ps.n_eq > 0 ? ps.n_eq : 1, 0);  // variables of a constraint row not yet eliminated
  for (int r = 0; r < ps.n_eq; ++r) {
    std::set<int> uniq(x_of_y[r].begin(), x_of_y[r].end());
    pending_x[r] = (int)uniq.size();
    x_of_y[r].assign(uniq.begin(), uniq.end());
  }
  std::vector<std::vector<int>> y_of_x(nx);
  for (int r = 0; r < ps.n_eq; ++r)
    for (int c : x_of_y[r]) y_of_x[c].push_back(r);
  // Segmented variant (n_segments > 1), for chain-like structures (horizon problems): every connected component is
  // cut into n_segments pieces along its longest axis (BFS layers from a pseudo-peripheral vertex); the layer of
  // vertices at each cut is a separator that is eliminated last.  The pieces factor independently, so the
  // elimination tree is ~n_segments times lower; the price is the fill of the separator blocks.
  // sep_level[v]: -1 = interior vertex; k >= 0 = member of a separator chosen at dissection depth k (0 = the cut made first,
  // eliminated last).  The flat variant puts every separator at level 0.
  std::vector<int> sep_level(n, -1);
  int max_sep_level = -1;
  if (n_segments > 1 || nd_depth > 0) {
    auto bfs = [&](int src, std::vector<int>& dist) {
      dist.assign(n, -1);
      std::vector<int> q{src};
      dist[src] = 0;
      for (size_t h = 0; h < q.size(); ++h)
        for (int w : adj[q[h]])
          if (dist[w] < 0) {
            dist[w] = dist[q[h]] + 1;
            q.push_back(w);
          }
      return q;
    };
    std::vector<char> seen(n, 0);
    std::vector<int> d0, du;
    std::vector<int> seg(n, 0);
    for (int s0 = 0; s0 < n; ++s0) {
      if (seen[s0]) continue;
      const std::vector<int> comp = bfs(s0, d0);
      for (int v : comp) seen[v] = 1;
      const int u = comp.back();  // farthest from s0: one end of the component
      const std::vector<int> cu = bfs(u, du);
      const int depth = du[cu.back()] + 1;
      if (nd_depth <= 0) {
        if (depth < 4 * n_segments) continue;  // too short to be worth cutting
        for (int x : comp) seg[x] = (int)((int64_t)du[x] * n_segments / depth);
        for (int x : comp)
          for (int w : adj[x])
            if (seg[w] < seg[x]) sep_level[x] = 0;
        max_sep_level = std::max(max_sep_level, 0);
        continue;
      }
      // Nested dissection along the BFS axis: the layers [lo, hi) are cut near the middle, then both halves again, nd_depth
      // times.  In a BFS layering edges join only equal or adjacent layers, so the cut between layers c-1 and c is crossed
      // by a bipartite edge set; its MINIMUM VERTEX COVER (Koenig: from a maximum matching) is the smallest vertex
      // separator for that cut -- on the horizon problems here the 7 joint angles of one stage, where "all vertices of
      // layer c with a neighbour in c-1" is three times that.  Among the cuts within an eighth of the interval around
      // the middle the smallest cover wins.
      std::vector<std::vector<int>> layer(depth);
      for (int x : comp) layer[du[x]].push_back(x);
      auto cover_of_cut = [&](int c) {
        std::vector<int> A, B;  // A: layer c-1 vertices with a neighbour in layer c; B: the other side
        std::vector<int> idB(n, -1);
        for (int x : layer[c]) {
          if (sep_level[x] >= 0) continue;
          for (int w : adj[x])
            if (du[w] == c - 1 && sep_level[w] < 0) {
              idB[x] = (int)B.size();
              B.push_back(x);
              break;
            }
        }
        std::vector<std::vector<int>> nb;
        for (int x : layer[c - 1]) {
          if (sep_level[x] >= 0) continue;
          std::vector<int> e;
          for (int w : adj[x])
            if (du[w] == c && idB[w] >= 0) e.push_back(idB[w]);
          if (!e.empty()) {
            A.push_back(x);
            nb.push_back(e);
          }
        }
        std::vector<int> matchB(B.size(), -1), matchA(A.size(), -1);
        std::vector<char> vis;
        std::function<bool(int)> augment = [&](int a) {
          for (int b : nb[a]) {
            if (vis[b]) continue;
            vis[b] = 1;
            if (matchB[b] < 0 || augment(matchB[b])) {
              matchB[b] = a;
              matchA[a] = b;
              return true;
            }
          }
          return false;
        };
        for (size_t a = 0; a < A.size(); ++a) {
          vis.assign(B.size(), 0);
          augment((int)a);
        }
        // Koenig: Z = reachable from the unmatched A vertices by alternating paths; cover = (A \ Z) + (B & Z)
        std::vector<char> zA(A.size(), 0), zB(B.size(), 0);
        std::vector<int> stack;
        for (size_t a = 0; a < A.size(); ++a)
          if (matchA[a] < 0) {
            zA[a] = 1;
            stack.push_back((int)a);
          }
        while (!stack.empty()) {
          const int a = stack.back();
          stack.pop_back();
          for (int b : nb[a]) {
            if (zB[b]) continue;
            zB[b] = 1;
            const int a2 = matchB[b];
            if (a2 >= 0 && !zA[a2]) {
              zA[a2] = 1;
              stack.push_back(a2);
            }
          }
        }
        std::vector<int> cover;
        for (size_t a = 0; a < A.size(); ++a)
          if (!zA[a]) cover.push_back(A[a]);
        for (size_t b = 0; b < B.size(); ++b)
          if (zB[b]) cover.push_back(B[b]);
        return cover;
      };
      // nd_bias > 0.5: an interval that still has a free end (an end of the chain) is cut nearer to its separator end.
      // Pieces bounded by separators on BOTH sides carry the fill (every column of theirs couples with a separator), end
      // pieces carry none: shorter inner pieces trade a little tree height for fill, i.e. for shared memory.
      std::function<void(int, int, int)> dissect = [&](int lo, int hi, int level) {
        if (level >= nd_depth || hi - lo < 8) return;
        const bool free_lo = lo == 0, free_hi = hi == depth;
        const double frac = (free_lo && !free_hi) ? nd_bias : ((free_hi && !free_lo) ? 1.0 - nd_bias : 0.5);
        const int mid = lo + (int)((hi - lo) * frac + 0.5), win = std::max(1, (hi - lo) / 8);
        int best_c = -1;
        std::vector<int> best_cover;
        for (int c = std::max(lo + 2, mid - win); c <= std::min(hi - 2, mid + win); ++c) {
          std::vector<int> cov = cover_of_cut(c);
          if (best_c < 0 || cov.size() < best_cover.size() ||
              (cov.size() == best_cover.size() && std::abs(c - mid) < std::abs(best_c - mid))) {
            best_c = c;
            best_cover = std::move(cov);
          }
        }
        if (best_c < 0) return;
        for (int v : best_cover) sep_level[v] = level;
        max_sep_level = std::max(max_sep_level, level);
        dissect(lo, best_c, level + 1);
        dissect(best_c, hi, level + 1);
      };
      dissect(0, depth, 0);
    }
  }
  std::vector<std::set<int>> g = adj;  // elimination graph (mutated)
  pl.perm.reserve(n);
  std::vector<std::vector<int>> col_struct;  // rows (old indices) below the diagonal of each eliminated column
  col_struct.reserve(n);
  int cur_level = max_sep_level + 1;  // separators of level >= cur_level are eligible (deepest first); interior vertices always
  for (int step = 0; step < n; ++step) {
    int best = -1;
    size_t best_deg = ~(size_t)0;
    while (best < 0) {
      for (int v = 0; v < n; ++v) {
        if (done[v]) continue;
        if (v >= nx && pending_x[v - nx] > 0) continue;          // constraint row not yet eligible
        if (sep_level[v] >= 0 && sep_level[v] < cur_level) continue;  // separators wait until everything below them is done
        if (g[v].size() < best_deg) {
          best_deg = g[v].size();
          best = v;
        }
      }
      if (best >= 0 || cur_level == 0) break;
      cur_level -= 1;  // this phase is exhausted: admit the next separator level
    }
    if (best < 0) {  // only ineligible rows left (cannot happen: all x are always eligible) -- take any
      for (int v = 0; v < n; ++v)
        if (!done[v]) { best = v; break; }
    }
    done[best] = 1;
    pl.perm.push_back(best);
    std::vector<int> nb(g[best].begin(), g[best].end());
    col_struct.push_back(nb);
    for (int a : nb) g[a].erase(best);
    for (size_t u = 0; u < nb.size(); ++u)
      for (size_t w = u + 1; w < nb.size(); ++w) {
        g[nb[u]].insert(nb[w]);
        g[nb[w]].insert(nb[u]);
      }
    g[best].clear();
    if (best < nx)
      for (int r : y_of_x[best]) pending_x[r] -= 1;
  }
  pl.iperm.assign(n, 0);
  for (int k = 0; k < n; ++k) pl.iperm[pl.perm[k]] = k;

  // ---- pattern of L in the permuted order ----
  pl.colptr.assign(n + 1, 0);
  for (int j = 0; j < n; ++j) {
    std::vector<int> rows;
    for (int old : col_struct[j]) rows.push_back(pl.iperm[old]);
    std::sort(rows.begin(), rows.end());
    pl.colptr[j + 1] = pl.colptr[j] + (int)rows.size();
    pl.rowidx.insert(pl.rowidx.end(), rows.begin(), rows.end());
  }

  // ---- factor program (left-looking, column by column) ----
  // row pattern of L: for each row j the (column k, entry index) pairs with k < j
  std::vector<std::vector<std::pair<int, int>>> row_pat(n);
  for (int k = 0; k < n; ++k)
    for (int e = pl.colptr[k]; e < pl.colptr[k + 1]; ++e) row_pat[pl.rowidx[e]].push_back({k, e});
  auto entry = [&](int row, int col) -> int {  // index into rowidx of L(row, col), row > col
    const auto lo = pl.rowidx.begin() + pl.colptr[col], hi = pl.rowidx.begin() + pl.colptr[col + 1];
    const auto it = std::lower_bound(lo, hi, row);
    return (it == hi || *it != row) ? -1 : (int)(it - pl.rowidx.begin());
  };
  std::vector<int32_t> prog;
  for (int j = 0; j < n; ++j) {
    prog.push_back((int32_t)row_pat[j].size());
    for (const auto& ke : row_pat[j]) {
      const int k = ke.first, e_jk = ke.second;
      prog.push_back(n + e_jk);  // position of L(j,k)
      prog.push_back(k);         // position of D(k)
      // rows i > j present in column k: they are updated in column j (fill guarantees they exist there)
      const size_t count_at = prog.size();
      prog.push_back(0);
      int cnt = 0;
      for (int e = pl.colptr[k]; e < pl.colptr[k + 1]; ++e) {
        const int i = pl.rowidx[e];
        if (i <= j) continue;
        const int e_ij = entry(i, j);
        if (e_ij < 0) continue;  // cannot happen for a correct symbolic factorisation
        prog.push_back(n + e_ij);
        prog.push_back(n + e);
        ++cnt;
        pl.flops += 1;
      }
      prog[count_at] = cnt;
      pl.flops += 2;
    }
  }
  // header (32 entries) + sections
  //  [0] n  [1] nnzL  [2] colptr  [3] rowidx  [4] perm  [5] sign  [6] factor program  [7] end of LDL part
  //  large mode only:
  //  [8]/[9] JE row/col  [10]/[11] JI row/col  [12] pos of every H nz  [13] pos of every JE nz
  //  [14] JI'SJI program: count, then {pos, a, b, r}   [15] JE'JE program: count, then {pos, a, b}
  //  [16] fc tape: n_instr, offset of its constants in dtable, 2 pad words, then instr[4*n] (16-byte aligned)
  //  [17] kkt tape: same
  std::vector<int32_t>& t = pl.table;
  t.assign(32, 0);
  auto section = [&](int slot) { t[slot] = (int32_t)t.size(); };
  t[0] = n;
  t[1] = pl.nnzL();
  section(2);
  t.insert(t.end(), pl.colptr.begin(), pl.colptr.end());
  section(3);
  t.insert(t.end(), pl.rowidx.begin(), pl.rowidx.end());
  section(4);
  t.insert(t.end(), pl.perm.begin(), pl.perm.end());
  section(5);
  for (int j = 0; j < n; ++j) t.push_back(pl.perm[j] < nx ? 1 : -1);
  section(6);
  t.insert(t.end(), prog.begin(), prog.end());
  section(7);
  if (large) {
    section(8);
    t.insert(t.end(), ps.jac_eq.row.begin(), ps.jac_eq.row.end());
    section(9);
    t.insert(t.end(), ps.jac_eq.col.begin(), ps.jac_eq.col.end());
    section(10);
    t.insert(t.end(), ps.jac_ineq.row.begin(), ps.jac_ineq.row.end());
    section(11);
    t.insert(t.end(), ps.jac_ineq.col.begin(), ps.jac_ineq.col.end());
    section(12);
    for (int k = 0; k < ps.hess.nnz(); ++k) t.push_back(pl.pos(ps.hess.row[k], ps.hess.col[k]));
    section(13);
    for (int k = 0; k < ps.jac_eq.nnz(); ++k) t.push_back(pl.pos(nx + ps.jac_eq.row[k], ps.jac_eq.col[k]));
    auto pair_program = [&](const Sparsity& sp, int n_rows, bool with_row) {
      const size_t count_at = t.size();
      t.push_back(0);
      int cnt = 0;
      std::vector<std::vector<int>> by_row(n_rows > 0 ? n_rows : 1);
      for (int k = 0; k < sp.nnz(); ++k) by_row[sp.row[k]].push_back(k);
      for (int r = 0; r < n_rows; ++r) {
        const auto& ks = by_row[r];
        for (size_t u = 0; u < ks.size(); ++u)
          for (size_t w = 0; w < ks.size(); ++w) {
            const int cu = sp.col[ks[u]], cw = sp.col[ks[w]];
            if (cu < cw || (cu == cw && u != w)) continue;
            t.push_back(pl.pos(cu, cw));
            t.push_back(ks[u]);
            t.push_back(ks[w]);
            if (with_row) t.push_back(r);
            ++cnt;
          }
      }
      t[count_at] = cnt;
    };
    section(14);
    pair_program(ps.jac_ineq, ps.n_ineq, true);
    section(15);
    pair_program(ps.jac_eq, ps.n_eq, false);
    auto tape_section = [&](int slot, const Tape& tape) {
      while (t.size() % 4 != 0) t.push_back(0);  // 16-byte alignment of the instruction rows (table base is 256-byte aligned)
      section(slot);
      t.push_back((int32_t)tape.n_instr());
      t.push_back((int32_t)pl.dtable.size());
      t.push_back(0);
      t.push_back(0);
      t.insert(t.end(), tape.instr.begin(), tape.instr.end());
      pl.dtable.insert(pl.dtable.end(), tape.consts.begin(), tape.consts.end());
    };
    tape_section(16, ps.fc);
    tape_section(17, ps.kkt);
  }
  return pl;
}

}  // namespace bo
