// bo_coop.cpp -- tables of the cooperative (instance-per-CTA) tier: tape partitioning, level-scheduled
// sparse LDL' program, KKT assembly program, CSR/CSC views of the Jacobians.  See bo_coop.h.
#include "bo_coop.h"

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <numeric>
#include <queue>
#include <sstream>

#include "bo_opcodes.h"

#define BO_OPX_NOP 5 /* padding rows of the lane-interleaved instruction streams (kernel: bo_ipm_cta.cuh) */

namespace bo {
namespace {

// ======================================================================================
// Tape -> DAG
// ======================================================================================
struct Node {
  int op = 0;
  int a = -1, b = -1, c = -1;  // interior: child node ids; INPUT: a = element, b = segment; CONST: a = index
};
struct OutRef {
  int seg, elem, node;
};
struct Dag {
  std::vector<Node> nodes;
  std::vector<OutRef> outs;
};

bool is_leaf_op(int op) { return op == BO_OP_INPUT || op == BO_OP_CONST; }

Dag build_dag(const Tape& t) {
  Dag g;
  std::vector<int> def(t.n_work, -1);
  std::map<std::pair<int, std::pair<int, int>>, int> leaves;
  for (int64_t i = 0; i < t.n_instr(); ++i) {
    const int32_t* r = &t.instr[4 * i];
    const int op = r[0] & 0xFF, c = (int)((uint32_t)r[0] >> 8), dst = r[1], a = r[2], b = r[3];
    if (op == BO_OP_OUTPUT) {
      g.outs.push_back({b, a, def[dst]});
      continue;
    }
    Node nd;
    nd.op = op;
    if (is_leaf_op(op)) {
      const auto key = std::make_pair(op, std::make_pair(a, op == BO_OP_INPUT ? b : 0));
      auto it = leaves.find(key);
      if (it == leaves.end()) {
        nd.a = a;
        nd.b = op == BO_OP_INPUT ? b : 0;
        g.nodes.push_back(nd);
        it = leaves.emplace(key, (int)g.nodes.size() - 1).first;
      }
      def[dst] = it->second;
      continue;
    }
    if (op == BO_OP_IF_ELSE) {
      nd.a = def[a];
      nd.b = def[b];
      nd.c = def[c];
    } else if (op >= BO_OP_NEG && op <= BO_OP_COSH) {
      nd.a = def[a];
    } else {
      nd.a = def[a];
      nd.b = def[b];
    }
    g.nodes.push_back(nd);
    def[dst] = (int)g.nodes.size() - 1;
  }
  return g;
}

struct UnionFind {
  std::vector<int> p;
  explicit UnionFind(int n) : p(n) { std::iota(p.begin(), p.end(), 0); }
  int find(int x) {
    while (p[x] != x) x = p[x] = p[p[x]];
    return x;
  }
  void unite(int a, int b) {
    a = find(a);
    b = find(b);
    if (a != b) p[std::max(a, b)] = std::min(a, b);
  }
};

// ======================================================================================
// Emission of one instruction stream (virtual code -> slots, leaves re-materialised on demand)
// ======================================================================================
struct VI {
  int op;
  int a = -1, b = -1, c = -1;  // value ids (DAG node id, or temp id >= n_nodes); OUTPUT: a = value, b = segment, c = element
  int val = -1;                // value defined (-1 for OUTPUT)
};

struct Stream {
  std::vector<int32_t> rows;
  int n_work = 0;
};

// `leaf_row(node, dst)` returns the 4-int row loading a leaf-like value into slot dst.
Stream assign_slots(const std::vector<VI>& code, const std::function<bool(int)>& leaflike,
                    const std::function<void(int, int, int32_t*)>& leaf_row) {
  const int GAP = 48;
  std::map<int, std::vector<int>> uses;
  for (int k = 0; k < (int)code.size(); ++k) {
    const VI& v = code[k];
    const int ops[3] = {v.a, v.op == BO_OP_OUTPUT ? -1 : v.b, v.op == BO_OP_OUTPUT ? -1 : v.c};
    for (int o : ops)
      if (o >= 0) {
        auto& l = uses[o];
        if (l.empty() || l.back() != k) l.push_back(k);
      }
  }
  std::map<int, int> slot_of, use_ptr;
  std::priority_queue<int, std::vector<int>, std::greater<int>> free_slots;
  int high = 0;
  auto take = [&]() {
    if (!free_slots.empty()) {
      const int s = free_slots.top();
      free_slots.pop();
      return s;
    }
    return high++;
  };
  Stream st;
  auto push = [&](int32_t w0, int32_t w1, int32_t w2, int32_t w3) {
    st.rows.push_back(w0);
    st.rows.push_back(w1);
    st.rows.push_back(w2);
    st.rows.push_back(w3);
  };
  for (int k = 0; k < (int)code.size(); ++k) {
    const VI& v = code[k];
    const bool is_out = v.op == BO_OP_OUTPUT;
    int ops[3] = {v.a, is_out ? -1 : v.b, is_out ? -1 : v.c};
    for (int o : ops)
      if (o >= 0 && !slot_of.count(o)) {  // only leaf-like values can be non-resident here
        const int s = take();
        int32_t row[4];
        leaf_row(o, s, row);
        push(row[0], row[1], row[2], row[3]);
        slot_of[o] = s;
      }
    const int sa = ops[0] >= 0 ? slot_of[ops[0]] : 0, sb = ops[1] >= 0 ? slot_of[ops[1]] : 0,
              sc = ops[2] >= 0 ? slot_of[ops[2]] : 0;
    // release operands: last use, or a leaf whose next use is far away
    for (int u = 0; u < 3; ++u) {
      const int o = ops[u];
      if (o < 0 || !slot_of.count(o)) continue;
      bool dup = false;
      for (int w = 0; w < u; ++w) dup = dup || ops[w] == o;
      if (dup) continue;
      const auto& l = uses[o];
      int ptr = use_ptr[o];
      while (ptr < (int)l.size() && l[ptr] <= k) ++ptr;
      use_ptr[o] = ptr;
      if (ptr >= (int)l.size() || (leaflike(o) && l[ptr] - k > GAP)) {
        free_slots.push(slot_of[o]);
        slot_of.erase(o);
      }
    }
    if (is_out) {
      push(BO_OP_OUTPUT, sa, v.c, v.b);
      continue;
    }
    const int dst = take();
    slot_of[v.val] = dst;
    if (v.op == BO_OP_IF_ELSE)
      push(BO_OP_IF_ELSE | (sc << 8), dst, sa, sb);
    else
      push(v.op, dst, sa, sb);
    if (!uses.count(v.val)) {
      free_slots.push(dst);
      slot_of.erase(v.val);
    }
  }
  st.n_work = std::max(high, 1);
  return st;
}

// ======================================================================================
// Tape partitioning
// ======================================================================================
struct Addend {
  int node;
  bool neg;
  std::vector<int> coefs;
};

struct PartTape {
  std::vector<Stream> subs;
  Stream pre;
  int n_pe = 0, n_part = 0;
  int n_work = 1;            // work slots of the longest-lived sub-tape
  std::vector<int32_t> red;  // seg, elem, start, count
  CoopTapeInfo info;
};

PartTape partition_tape(const Tape& tape, int max_threads) {
  const Dag g = build_dag(tape);
  const int N = (int)g.nodes.size();
  const int part_seg = (int)tape.out_sizes.size();  // extra output segment: per-thread partial sums
  const int pe_seg = (int)tape.in_sizes.size();     // extra input segment: parameter-only sub-expressions

  // x-dependence (anything but the parameter segment, which is input segment 1 of both solver tapes)
  std::vector<char> xdep(N, 0);
  std::vector<int> n_uses(N, 0);
  for (int n = 0; n < N; ++n) {
    const Node& nd = g.nodes[n];
    if (nd.op == BO_OP_INPUT) {
      xdep[n] = nd.b != 1;
    } else if (nd.op != BO_OP_CONST) {
      for (int ch : {nd.a, nd.b, nd.c})
        if (ch >= 0) {
          xdep[n] = xdep[n] || xdep[ch];
          n_uses[ch] += 1;
        }
    }
  }
  for (const OutRef& o : g.outs) n_uses[o.node] += 1;
  auto leaf = [&](int n) { return is_leaf_op(g.nodes[n].op); };
  auto leaflike = [&](int n) { return n < N && (leaf(n) || !xdep[n]); };  // leaves + parameter-only sub-expressions

  // linear trees at the roots: ADD / SUB / NEG / MUL-by-x-independent-factor nodes used exactly once
  std::vector<signed char> trav(N, -1);
  std::function<bool(int)> traversable = [&](int n) -> bool {
    if (trav[n] >= 0) return trav[n] != 0;
    const Node& nd = g.nodes[n];
    bool t = false;
    if (xdep[n] && n_uses[n] == 1) {
      if (nd.op == BO_OP_ADD || nd.op == BO_OP_SUB || nd.op == BO_OP_NEG) {
        t = true;
      } else if (nd.op == BO_OP_MUL && xdep[nd.a] != xdep[nd.b]) {
        const int d = xdep[nd.a] ? nd.a : nd.b;
        t = !leaf(d) && traversable(d);
      }
    }
    trav[n] = t ? 1 : 0;
    return t;
  };
  std::vector<char> intree(N, 0);
  std::vector<std::vector<Addend>> addends(g.outs.size());
  for (size_t oi = 0; oi < g.outs.size(); ++oi) {
    struct Item {
      int node;
      bool neg;
      std::vector<int> coefs;
    };
    std::vector<Item> stack{{g.outs[oi].node, false, {}}};
    while (!stack.empty()) {
      Item it = stack.back();
      stack.pop_back();
      const Node& nd = g.nodes[it.node];
      if (!leaf(it.node) && traversable(it.node)) {
        intree[it.node] = 1;
        if (nd.op == BO_OP_ADD || nd.op == BO_OP_SUB) {
          stack.push_back({nd.b, nd.op == BO_OP_SUB ? !it.neg : it.neg, it.coefs});
          stack.push_back({nd.a, it.neg, it.coefs});
        } else if (nd.op == BO_OP_NEG) {
          stack.push_back({nd.a, !it.neg, it.coefs});
        } else {  // MUL by an x-independent factor
          const int d = xdep[nd.a] ? nd.a : nd.b, k = xdep[nd.a] ? nd.b : nd.a;
          it.coefs.push_back(k);
          stack.push_back({d, it.neg, it.coefs});
        }
      } else {
        addends[oi].push_back({it.node, it.neg, it.coefs});
      }
    }
  }

  // base components: x-dependent interior nodes outside the root trees, linked through shared operands
  UnionFind uf(N);
  auto interior = [&](int n) { return xdep[n] && !leaf(n); };
  for (int n = 0; n < N; ++n) {
    if (!interior(n) || intree[n]) continue;
    const Node& nd = g.nodes[n];
    for (int ch : {nd.a, nd.b, nd.c})
      if (ch >= 0 && interior(ch)) uf.unite(n, ch);
  }
  std::vector<int> comp_size(N, 0);
  for (int n = 0; n < N; ++n)
    if (interior(n) && !intree[n]) comp_size[uf.find(n)] += 1;

  // tasks: an unsplit output (whole root tree evaluated by one thread) or the addends of a split output
  // that fall into one component
  struct Task {
    int out;                   // output index
    std::vector<int> addends;  // indices into addends[out]; empty = unsplit (evaluate the root)
  };
  std::map<int, std::vector<Task>> comp_tasks;  // component root (or -1 - k for loose outputs) -> tasks
  std::map<int, int64_t> comp_cost;
  std::map<int, int> comp_first;
  int n_loose = 0, n_split = 0;
  std::vector<std::vector<int>> split_comps(g.outs.size());  // for split outputs: components in order of appearance
  for (size_t oi = 0; oi < g.outs.size(); ++oi) {
    std::vector<int> comps;
    for (const Addend& ad : addends[oi])
      if (interior(ad.node)) {
        const int c = uf.find(ad.node);
        if (std::find(comps.begin(), comps.end(), c) == comps.end()) comps.push_back(c);
      }
    int64_t tree_cost = (int64_t)addends[oi].size();
    if (comps.size() <= 1) {
      const int c = comps.empty() ? -1 - (n_loose++) : comps[0];
      comp_tasks[c].push_back({(int)oi, {}});
      comp_cost[c] += tree_cost + 1;
      if (!comp_first.count(c)) comp_first[c] = (int)oi;
    } else {
      ++n_split;
      split_comps[oi] = comps;
      std::map<int, Task> per;
      for (size_t k = 0; k < addends[oi].size(); ++k) {
        const Addend& ad = addends[oi][k];
        const int c = interior(ad.node) ? uf.find(ad.node) : comps[0];
        Task& tk = per[c];
        tk.out = (int)oi;
        tk.addends.push_back((int)k);
      }
      for (auto& kv : per) {
        comp_cost[kv.first] += (int64_t)kv.second.addends.size() * 2;
        if (!comp_first.count(kv.first)) comp_first[kv.first] = (int)oi;
        comp_tasks[kv.first].push_back(std::move(kv.second));
      }
    }
  }
  for (auto& kv : comp_cost)
    if (kv.first >= 0) kv.second += comp_size[kv.first];

  // components -> threads: longest first onto the least-loaded thread (ties: lowest id), so that the
  // big isomorphic stage components end up one per lane in neighbouring lanes (lock-step in the interpreter)
  std::vector<int> comps;
  for (const auto& kv : comp_tasks) comps.push_back(kv.first);
  std::sort(comps.begin(), comps.end(), [&](int a, int b) {
    if (comp_cost[a] != comp_cost[b]) return comp_cost[a] > comp_cost[b];
    return comp_first[a] < comp_first[b];
  });
  const int nsub = std::max(1, std::min<int>(max_threads, (int)comps.size()));
  std::vector<int64_t> load(nsub, 0);
  std::vector<std::vector<Task>> thread_tasks(nsub);
  std::map<int, int> comp_thread;
  for (int c : comps) {
    int best = 0;
    for (int t = 1; t < nsub; ++t)
      if (load[t] < load[best]) best = t;
    load[best] += comp_cost[c];
    comp_thread[c] = best;
    for (const Task& tk : comp_tasks[c]) thread_tasks[best].push_back(tk);
  }

  PartTape pt;
  pt.info.nsub = nsub;
  pt.info.n_components = (int)comps.size();
  pt.info.n_split_outputs = n_split;

  // partial slots of the split outputs: one per contributing THREAD, contiguous, in thread order
  std::map<std::pair<int, int>, int> part_slot;  // (output, thread) -> slot
  for (size_t oi = 0; oi < g.outs.size(); ++oi) {
    if (split_comps[oi].empty()) continue;
    std::vector<int> threads;
    for (int c : split_comps[oi]) threads.push_back(comp_thread[c]);
    std::sort(threads.begin(), threads.end());
    threads.erase(std::unique(threads.begin(), threads.end()), threads.end());
    const int start = pt.n_part;
    for (int t : threads) part_slot[{(int)oi, t}] = pt.n_part++;
    pt.red.push_back(g.outs[oi].seg);
    pt.red.push_back(g.outs[oi].elem);
    pt.red.push_back(start);
    pt.red.push_back((int)threads.size());
  }
  pt.info.n_part = pt.n_part;
  pt.info.n_red = (int)pt.red.size() / 4;

  // parameter-only sub-expressions referenced from the x-dependent code: evaluated once per instance
  std::map<int, int> pe_index;
  auto pe_of = [&](int n) {
    auto it = pe_index.find(n);
    if (it == pe_index.end()) it = pe_index.emplace(n, (int)pe_index.size()).first;
    return it->second;
  };
  auto leaf_row_main = [&](int n, int dst, int32_t* row) {
    const Node& nd = g.nodes[n];
    if (nd.op == BO_OP_CONST) {
      row[0] = BO_OP_CONST; row[1] = dst; row[2] = nd.a; row[3] = 0;
    } else if (nd.op == BO_OP_INPUT) {
      row[0] = BO_OP_INPUT; row[1] = dst; row[2] = nd.a; row[3] = nd.b;
    } else {
      row[0] = BO_OP_INPUT; row[1] = dst; row[2] = pe_of(n); row[3] = pe_seg;
    }
  };

  for (int t = 0; t < nsub; ++t) {
    std::vector<Task>& tasks = thread_tasks[t];
    std::stable_sort(tasks.begin(), tasks.end(), [](const Task& a, const Task& b) { return a.out < b.out; });
    {  // several components of one split output on this thread share ONE partial slot: merge their addends
      std::vector<Task> merged;
      for (Task& tk : tasks) {
        if (!merged.empty() && merged.back().out == tk.out && !tk.addends.empty() && !merged.back().addends.empty()) {
          merged.back().addends.insert(merged.back().addends.end(), tk.addends.begin(), tk.addends.end());
        } else {
          merged.push_back(std::move(tk));
        }
      }
      for (Task& tk : merged) std::sort(tk.addends.begin(), tk.addends.end());
      tasks = std::move(merged);
    }
    std::vector<VI> code;
    std::vector<char> done(N, 0);
    int n_temp = 0;
    auto emit_node = [&](int root) {
      if (leaflike(root) || done[root]) return;
      std::vector<std::pair<int, int>> stack{{root, 0}};
      while (!stack.empty()) {
        auto& top = stack.back();
        const int n = top.first;
        if (done[n]) {
          stack.pop_back();
          continue;
        }
        const Node& nd = g.nodes[n];
        const int kids[3] = {nd.a, nd.b, nd.c};
        bool pushed = false;
        while (top.second < 3) {
          const int ch = kids[top.second++];
          if (ch >= 0 && !leaflike(ch) && !done[ch]) {
            stack.push_back({ch, 0});
            pushed = true;
            break;
          }
        }
        if (pushed) continue;
        VI v;
        v.op = nd.op;
        v.a = nd.a;
        v.b = nd.b;
        v.c = nd.c;
        v.val = n;
        code.push_back(v);
        done[n] = 1;
        stack.pop_back();
      }
    };
    auto temp_op = [&](int op, int a, int b) {
      VI v;
      v.op = op;
      v.a = a;
      v.b = b;
      v.val = N + n_temp++;
      code.push_back(v);
      return v.val;
    };
    for (const Task& tk : tasks) {
      const OutRef& o = g.outs[tk.out];
      if (tk.addends.empty()) {
        emit_node(o.node);
        VI v;
        v.op = BO_OP_OUTPUT;
        v.a = o.node;
        v.b = o.seg;
        v.c = o.elem;
        code.push_back(v);
        continue;
      }
      int acc = -1;
      for (int k : tk.addends) {
        const Addend& ad = addends[tk.out][k];
        emit_node(ad.node);
        int val = ad.node;
        for (int cf : ad.coefs) {
          emit_node(cf);  // no-op: coefficients are leaf-like
          val = temp_op(BO_OP_MUL, val, cf);
        }
        if (acc < 0)
          acc = ad.neg ? temp_op(BO_OP_NEG, val, -1) : val;
        else
          acc = temp_op(ad.neg ? BO_OP_SUB : BO_OP_ADD, acc, val);
      }
      VI v;
      v.op = BO_OP_OUTPUT;
      v.a = acc;
      v.b = part_seg;
      v.c = part_slot.at({tk.out, t});
      code.push_back(v);
    }
    Stream st = assign_slots(code, leaflike, leaf_row_main);
    pt.info.total_instr += (int64_t)st.rows.size() / 4;
    pt.info.max_len = std::max<int64_t>(pt.info.max_len, (int64_t)st.rows.size() / 4);
    pt.n_work = std::max(pt.n_work, st.n_work);
    pt.subs.push_back(std::move(st));
  }

  // the once-per-instance tape of the parameter-only sub-expressions (output segment 0 = the pe vector)
  {
    std::vector<std::pair<int, int>> order;  // (pe index, node)
    for (const auto& kv : pe_index) order.push_back({kv.second, kv.first});
    std::sort(order.begin(), order.end());
    std::vector<VI> code;
    std::vector<char> done(N, 0);
    for (const auto& on : order) {
      std::vector<std::pair<int, int>> stack{{on.second, 0}};
      while (!stack.empty()) {
        auto& top = stack.back();
        const int n = top.first;
        if (done[n] || leaf(n)) {
          stack.pop_back();
          continue;
        }
        const Node& nd = g.nodes[n];
        const int kids[3] = {nd.a, nd.b, nd.c};
        bool pushed = false;
        while (top.second < 3) {
          const int ch = kids[top.second++];
          if (ch >= 0 && !leaf(ch) && !done[ch]) {
            stack.push_back({ch, 0});
            pushed = true;
            break;
          }
        }
        if (pushed) continue;
        VI v;
        v.op = nd.op;
        v.a = nd.a;
        v.b = nd.b;
        v.c = nd.c;
        v.val = n;
        code.push_back(v);
        done[n] = 1;
        stack.pop_back();
      }
      VI v;
      v.op = BO_OP_OUTPUT;
      v.a = on.second;
      v.b = 0;
      v.c = on.first;
      code.push_back(v);
    }
    auto leaf_row_pre = [&](int n, int dst, int32_t* row) {
      const Node& nd = g.nodes[n];
      row[0] = nd.op;
      row[1] = dst;
      row[2] = nd.a;
      row[3] = nd.op == BO_OP_INPUT ? nd.b : 0;
    };
    pt.pre = assign_slots(code, [&](int n) { return n < N && leaf(n); }, leaf_row_pre);
    pt.n_pe = (int)pe_index.size();
    pt.info.n_pe = pt.n_pe;
    pt.info.pre_len = (int64_t)pt.pre.rows.size() / 4;
  }
  return pt;
}

// ======================================================================================
// Component classes: isomorphic components (same instruction rows up to which input / constant / output element
// they touch -- all stages of a horizon problem) share ONE generated straight-line function
// ======================================================================================
struct TapeClass {
  std::vector<int32_t> rows;           // representative rows
  std::vector<int> members;            // component (sub-tape) indices
  std::vector<int> var_rows;           // rows whose element index differs per member (INPUT / OUTPUT / non-uniform CONST)
};

std::vector<TapeClass> classify(const PartTape& pt, const Tape& tape) {
  std::map<std::vector<int32_t>, int> ids;
  std::vector<TapeClass> classes;
  for (size_t k = 0; k < pt.subs.size(); ++k) {
    const std::vector<int32_t>& rows = pt.subs[k].rows;
    std::vector<int32_t> key(rows);
    for (size_t i = 0; i < rows.size() / 4; ++i) {
      const int op = rows[4 * i] & 0xFF;
      if (op == BO_OP_INPUT || op == BO_OP_OUTPUT || op == BO_OP_CONST) key[4 * i + 2] = 0;
    }
    auto it = ids.find(key);
    if (it == ids.end()) {
      it = ids.emplace(std::move(key), (int)classes.size()).first;
      classes.emplace_back();
      classes.back().rows = rows;
    }
    classes[it->second].members.push_back((int)k);
  }
  for (TapeClass& c : classes) {
    for (size_t i = 0; i < c.rows.size() / 4; ++i) {
      const int op = c.rows[4 * i] & 0xFF;
      if (op == BO_OP_INPUT || op == BO_OP_OUTPUT) {
        bool same = true;
        for (int m : c.members) same = same && pt.subs[m].rows[4 * i + 2] == c.rows[4 * i + 2];
        if (!same) c.var_rows.push_back((int)i);
      } else if (op == BO_OP_CONST) {
        bool same = true;
        const double v0 = tape.consts[c.rows[4 * i + 2]];
        for (int m : c.members) {
          const double v = tape.consts[pt.subs[m].rows[4 * i + 2]];
          same = same && (v == v0 || (v != v && v0 != v0));
        }
        if (!same) c.var_rows.push_back((int)i);
      }
    }
  }
  return classes;
}

// Table section + CUDA source of a tape compiled per component class.  `pt` must hold ONE component per sub-tape
// (partition_tape with an unlimited thread count).  Work lists: entries { class, index-table offset, stride, count }
// = up to 32 members of one class, run by the 32 lanes of a warp in lock step; entries are spread over the warps
// longest first.
std::string append_gen_tape(std::vector<int32_t>& t, int slot, const PartTape& pt, const std::vector<TapeClass>& classes,
                            const Tape& tape, std::vector<double>& dtab, int n_warps, const std::string& tag, CoopTapeInfo* info) {
  while (t.size() % 4 != 0) t.push_back(0);
  const size_t s0 = t.size();
  t[slot] = (int32_t)s0;
  t.resize(s0 + TS_HEADER, 0);
  t[s0 + TS_NSUB] = 0;
  t[s0 + TS_CONST0] = (int32_t)dtab.size();
  dtab.insert(dtab.end(), tape.consts.begin(), tape.consts.end());
  t[s0 + TS_NPE] = pt.n_pe;
  t[s0 + TS_NPART] = pt.n_part;
  t[s0 + TS_PART_SEG] = (int32_t)tape.out_sizes.size();
  t[s0 + TS_PE_SEG] = (int32_t)tape.in_sizes.size();
  t[s0 + TS_PRE_N] = (int32_t)pt.pre.rows.size() / 4;
  t[s0 + TS_PRE_OFF] = (int32_t)t.size();
  t.insert(t.end(), pt.pre.rows.begin(), pt.pre.rows.end());
  t[s0 + TS_NRED] = (int32_t)pt.red.size() / 4;
  t[s0 + TS_RED_OFF] = (int32_t)t.size();
  t.insert(t.end(), pt.red.begin(), pt.red.end());
  // index tables: per class [var][member]
  std::vector<int32_t> ix_off(classes.size(), 0);
  for (size_t c = 0; c < classes.size(); ++c) {
    const TapeClass& cl = classes[c];
    ix_off[c] = (int32_t)t.size();
    for (int row : cl.var_rows)
      for (int m : cl.members) t.push_back(pt.subs[m].rows[4 * row + 2]);
  }
  // work lists
  struct Entry {
    int cls, first, count;
    int64_t cost;
  };
  std::vector<Entry> entries;
  for (size_t c = 0; c < classes.size(); ++c)
    for (int first = 0; first < (int)classes[c].members.size(); first += 32)
      entries.push_back({(int)c, first, std::min(32, (int)classes[c].members.size() - first), (int64_t)classes[c].rows.size() / 4});
  std::stable_sort(entries.begin(), entries.end(), [](const Entry& a, const Entry& b) { return a.cost > b.cost; });
  std::vector<std::vector<Entry>> per_warp(n_warps);
  std::vector<int64_t> load(n_warps, 0);
  for (const Entry& e : entries) {
    int best = 0;
    for (int w = 1; w < n_warps; ++w)
      if (load[w] < load[best]) best = w;
    load[best] += e.cost + 4;
    per_warp[best].push_back(e);
  }
  t[s0 + TS_GEN_NWARP] = n_warps;
  t[s0 + TS_GEN_WL_OFF] = (int32_t)t.size();
  const size_t wl0 = t.size();
  t.resize(wl0 + 2 * (size_t)n_warps, 0);
  for (int w = 0; w < n_warps; ++w) {
    t[wl0 + 2 * w] = (int32_t)t.size();
    t[wl0 + 2 * w + 1] = (int32_t)per_warp[w].size();
    for (const Entry& e : per_warp[w]) {
      t.push_back(e.cls);
      t.push_back(ix_off[e.cls] + e.first);
      t.push_back((int32_t)classes[e.cls].members.size());
      t.push_back(e.count);
    }
  }
  info->n_classes = (int)classes.size();
  info->warp_rows = *std::max_element(load.begin(), load.end());

  // code
  std::ostringstream o;
  const int n_in = (int)tape.in_sizes.size() + 1, n_out = (int)tape.out_sizes.size() + 1;
  for (size_t c = 0; c < classes.size(); ++c) {
    const TapeClass& cl = classes[c];
    Tape ct;
    ct.instr = cl.rows;
    ct.consts = tape.consts;
    ct.n_work = 1;
    for (size_t i = 0; i < cl.rows.size() / 4; ++i) ct.n_work = std::max(ct.n_work, cl.rows[4 * i + 1] + 1);
    std::map<int, int> var_of;
    for (size_t v = 0; v < cl.var_rows.size(); ++v) var_of[cl.var_rows[v]] = (int)v;
    info->code_rows += (int64_t)cl.rows.size() / 4;
    TapeEmitHooks hk;
    const std::string fname = "bo_" + tag + "_c" + std::to_string(c);
    hk.signature = std::string(cl.rows.size() / 4 > 48 ? "BO_NOINLINE" : "BO_DEVICE") + " void " + fname +
                   "(const int32_t* BO_RESTRICT ix, int stride, const double* const* BO_RESTRICT in, double* const* BO_RESTRICT out, "
                   "const double* BO_RESTRICT consts)";
    std::ostringstream pro;
    for (int k = 0; k < n_in; ++k) pro << "  const double* const BO_RESTRICT i" << k << " = in[" << k << "];\n";
    for (int k = 0; k < n_out; ++k) pro << "  double* const BO_RESTRICT o" << k << " = out[" << k << "];\n";
    pro << "  (void)ix; (void)stride; (void)consts;";
    for (int k = 0; k < n_in; ++k) pro << " (void)i" << k << ";";
    for (int k = 0; k < n_out; ++k) pro << " (void)o" << k << ";";
    pro << "\n";
    hk.prologue = pro.str();
    auto elem = [var_of](int64_t i, const int32_t* r) {
      auto it = var_of.find((int)i);
      if (it == var_of.end()) return std::to_string(r[2]);
      return "ix[" + std::to_string(it->second) + " * stride]";
    };
    hk.input_expr = [elem](int64_t i, const int32_t* r) { return "i" + std::to_string(r[3]) + "[" + elem(i, r) + "]"; };
    hk.const_expr = [elem, var_of](int64_t i, const int32_t* r) {
      if (!var_of.count((int)i)) return std::string();
      return "consts[" + elem(i, r) + "]";
    };
    hk.output_stmt = [elem](int64_t i, const int32_t* r, const std::string& v) {
      return "o" + std::to_string(r[3]) + "[" + elem(i, r) + "] = " + v + ";";
    };
    o << emit_tape_function(ct, fname, &hk) << "\n";
  }
  o << "BO_DEVICE void bo_gen_" << tag
    << "(const int32_t* BO_RESTRICT tab, const int32_t* BO_RESTRICT sec, int warp, int lane, const double* const* in, double* const* out, "
       "const double* consts) {\n"
       "  if (warp >= sec[" << (int)TS_GEN_NWARP << "]) return;\n"
       "  const int32_t* wl = tab + sec[" << (int)TS_GEN_WL_OFF << "] + 2 * warp;\n"
       "  const int32_t* e = tab + wl[0];\n"
       "  const int ne = wl[1];\n"
       "  for (int k = 0; k < ne; ++k, e += 4) {\n"
       "    if (lane >= e[3]) continue;\n"
       "    const int32_t* ix = tab + e[1] + lane;\n"
       "    switch (e[0]) {\n";
  for (size_t c = 0; c < classes.size(); ++c)
    o << "      case " << c << ": bo_" << tag << "_c" << c << "(ix, e[2], in, out, consts); break;\n";
  o << "      default: break;\n    }\n  }\n}\n\n";
  return o.str();
}

void append_tape_section(std::vector<int32_t>& t, int slot, const PartTape& pt, const Tape& tape, std::vector<double>& dtab) {
  while (t.size() % 4 != 0) t.push_back(0);
  const size_t s0 = t.size();
  t[slot] = (int32_t)s0;
  t.resize(s0 + TS_HEADER, 0);
  const int nsub = (int)pt.subs.size();
  t[s0 + TS_NSUB] = nsub;
  t[s0 + TS_CONST0] = (int32_t)dtab.size();
  dtab.insert(dtab.end(), tape.consts.begin(), tape.consts.end());
  t[s0 + TS_NPE] = pt.n_pe;
  t[s0 + TS_NPART] = pt.n_part;
  t[s0 + TS_PART_SEG] = (int32_t)tape.out_sizes.size();
  t[s0 + TS_PE_SEG] = (int32_t)tape.in_sizes.size();
  // pre tape (sequential rows)
  t[s0 + TS_PRE_N] = (int32_t)pt.pre.rows.size() / 4;
  t[s0 + TS_PRE_OFF] = (int32_t)t.size();
  t.insert(t.end(), pt.pre.rows.begin(), pt.pre.rows.end());
  // lengths
  t[s0 + TS_LEN_OFF] = (int32_t)t.size();
  for (const Stream& s : pt.subs) t.push_back((int32_t)s.rows.size() / 4);
  // reductions
  t[s0 + TS_NRED] = (int32_t)pt.red.size() / 4;
  t[s0 + TS_RED_OFF] = (int32_t)t.size();
  t.insert(t.end(), pt.red.begin(), pt.red.end());
  // lane-interleaved streams, one block per group of 32 sub-tapes: row i of lane l at wbase + i*32 + l
  const int n_warps = (nsub + 31) / 32;
  std::vector<int64_t> wbase(n_warps, 0);
  int64_t rows_total = 0;
  for (int w = 0; w < n_warps; ++w) {
    wbase[w] = rows_total;
    int64_t mx = 0;
    for (int l = 0; l < 32 && w * 32 + l < nsub; ++l) mx = std::max<int64_t>(mx, (int64_t)pt.subs[w * 32 + l].rows.size() / 4);
    rows_total += mx * 32;
  }
  t[s0 + TS_WBASE_OFF] = (int32_t)t.size();
  for (int w = 0; w < n_warps; ++w) t.push_back((int32_t)wbase[w]);
  while (t.size() % 4 != 0) t.push_back(0);
  t[s0 + TS_STREAM_OFF] = (int32_t)t.size();
  const size_t st0 = t.size();
  t.resize(st0 + (size_t)rows_total * 4, 0);
  for (size_t r = 0; r < (size_t)rows_total; ++r) t[st0 + 4 * r] = BO_OPX_NOP;
  for (int s = 0; s < nsub; ++s) {
    const int w = s / 32, l = s % 32;
    const auto& rows = pt.subs[s].rows;
    for (size_t i = 0; i < rows.size() / 4; ++i) {
      const size_t at = st0 + 4 * ((size_t)wbase[w] + i * 32 + l);
      for (int k = 0; k < 4; ++k) t[at + k] = rows[4 * i + k];
    }
  }
}


// ======================================================================================
// Lane programs: a warp-wide, fully pre-scheduled instruction stream
// ======================================================================================
// A stream is a sequence of fixed-size PACKETS: one header word per lane followed by PK operand words per lane
// (8 bytes each, lane-interleaved: word i of lane l at [i][l]).  A ROUND -- one target per group of G lanes -- is one or
// more packets:
//   header:  x = FIRST | LAST << 1 | LEVEL_END << 2 | EMPTY << 3 | log2(G) << 4,  y = tgt | POSITIVE << 15   (tgt = 0x7FFF: this lane finalises nothing)
//   operand: x = a | b << 16,          y = c                                     (padding: the always-zero cells)
// In a round every lane accumulates acc += vals[a] * (factor: vals[b] * vals[c] | solve: bp[b]); after the LAST packet the
// G lanes of a group add their accumulators up and lane 0 of the group finalises `tgt`; after a LEVEL_END packet the
// participating warps synchronise.  The flags are the same in all lanes of a warp.  POSITIVE marks a diagonal factor
// target whose pivot must be positive (variable block).  Every warp stream ends with 8 padding packets (prefetched, never run).
struct LaneTarget {
  int tgt;
  std::vector<std::array<int, 3>> con;
  bool positive = false;  // factor, diagonal target: pivot expected positive (variable block)
};
struct LaneProgram {
  int W = 1, G = 1, PK = 4;
  int pad_words = 0;  // words of trailing padding in every warp stream
  bool aligned = false;
  std::vector<std::vector<int32_t>> words;  // per warp: [step][32][2]
  double cost = 0.0;
  int64_t max_steps() const {
    size_t m = 0;
    for (const auto& w : words) m = std::max(m, w.size() / 64);
    return (int64_t)m;
  }
};

// `aligned`: the targets of a level arrive column by column with operand lists aligned on the column's k-list (see the
// factor program below); consecutive targets then go to consecutive lane groups of one warp, so that the second and
// third operand of a step are the same shared-memory word in the lanes of a column with the same sub-index (a broadcast).
// `sum_weight`: share of the OTHER warps' work that counts against a level (0: one CTA per SM, only the busiest warp
// matters; 1: many small CTAs share the SM's issue slots, so the total instruction count is what is paid for).
// G = 0: the lanes per target are chosen LEVEL BY LEVEL (the cheapest of 1..32 for that level under the same cost model;
// the header carries log2 G in bits 4..6, the executor's group sum reads it from there): a level of 60 long targets wants
// G = 4 on 8 warps, the level after it with 300 short ones G = 1.
LaneProgram schedule_program(const std::vector<std::vector<LaneTarget>>& levels, int W, int G_fixed, int PK, int zero_a, int zero_b,
                             int zero_c, bool emit, bool aligned = false, double sum_weight = 0.0) {
  LaneProgram pr;
  pr.W = W;
  pr.G = G_fixed;
  pr.PK = PK;
  pr.words.resize(W);
  // Cost model of one warp (cycles).  Measured on B200 (ncu, profiles/r02_coop_c5_packets_ncu.txt): the programs are
  // bound by the INSTRUCTION ISSUE of the warp on the critical path -- a packet is ~12 instructions per operand step
  // (index decode, three shared-memory loads, two FP64 operations) plus ~35 of bookkeeping (prefetch of the words
  // BO_LP_DEPTH packets ahead, flags), a round ends with the group sum and the finalisation.  So the cost of a level is
  // the instruction count of its busiest warp; padding operands cost as much as real ones.  (The absolute scale is off --
  // a lone warp needs ~4.9 cycles per instruction, not 2.2 -- only ratios are used.)
  const double cpi = 2.2;
  // same packet count, measured on C5 / C4: the aligned form is 13 % / 9 % faster (two of its three operands are broadcasts,
  // the unaligned lanes read three unrelated words: bank conflicts)
  const double c_packet = cpi * (12.0 * PK + 35.0) * (aligned ? 1.0 : 1.12), c_barrier = W > 1 ? 80.0 : 10.0;
  auto header = [&](std::vector<int32_t>& out, int G, unsigned flags, const LaneTarget* const* tg) {
    int log2g = 0;
    while ((1 << log2g) < G) ++log2g;
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane / G, sub = lane % G;
      int tgt = 0x7FFF, pos = 0;
      if (tg && tg[g] && sub == 0) {
        tgt = tg[g]->tgt;
        pos = tg[g]->positive ? 0x8000 : 0;
      }
      out.push_back((int32_t)(flags | ((unsigned)log2g << 4)));
      out.push_back((int32_t)((uint32_t)tgt | (uint32_t)pos));
    }
  };
  auto zero_words = [&](std::vector<int32_t>& out, int count) {
    for (int k = 0; k < count * 32; ++k) {
      out.push_back((int32_t)((uint32_t)zero_a | ((uint32_t)zero_b << 16)));
      out.push_back((int32_t)zero_c);
    }
  };
  // one level with G lanes per target: returns its cost, emits its packets when asked to
  auto do_level = [&](const std::vector<LaneTarget>& lv, const std::vector<int>& order, int G, bool emit_now) {
    const int ngrp = 32 / G, slots = W * ngrp;
    int log2g = 0;
    while ((1 << log2g) < G) ++log2g;
    const double c_round = cpi * (60.0 + 12.0 * log2g);
    const int rounds = ((int)lv.size() + slots - 1) / slots;
    double worst = 0.0, total = 0.0;
    for (int w = 0; w < W; ++w) {
      double cost_w = 0.0;
      // rounds of this warp in this level
      std::vector<std::array<const LaneTarget*, 32>> mine;
      for (int r = 0; r < rounds; ++r) {
        std::array<const LaneTarget*, 32> tg;
        tg.fill(nullptr);
        bool any = false;
        for (int g = 0; g < ngrp; ++g) {
          // slot q of the round: warp = q % W, group = q / W; aligned: 32 / G consecutive targets per warp
          const int q = g * W + w, k = aligned ? (r * W + w) * ngrp + g : r * slots + q;
          if (k < (int)lv.size()) {
            tg[g] = &lv[order[k]];
            any = true;
          }
        }
        if (any) mine.push_back(tg);
      }
      if (mine.empty()) {
        if (emit_now) {  // nothing to do in this level: only meet the barrier
          header(pr.words[w], 1, 4u | 8u, nullptr);  // LEVEL_END | EMPTY
          zero_words(pr.words[w], PK);
        }
        cost_w += cpi * 30.0;
      }
      for (size_t r = 0; r < mine.size(); ++r) {
        int maxlen = 0;
        for (int g = 0; g < ngrp; ++g)
          if (mine[r][g]) maxlen = std::max(maxlen, (int)mine[r][g]->con.size());
        const int K = (maxlen + G - 1) / G;
        const int np = std::max(1, (K + PK - 1) / PK);
        cost_w += c_round + np * c_packet;
        if (!emit_now) continue;
        for (int pk = 0; pk < np; ++pk) {
          const bool last = pk + 1 == np;
          const unsigned flags = (pk == 0 ? 1u : 0u) | (last ? 2u : 0u) | (last && r + 1 == mine.size() ? 4u : 0u);
          header(pr.words[w], G, flags, last ? mine[r].data() : nullptr);
          for (int st = pk * PK; st < (pk + 1) * PK; ++st)
            for (int lane = 0; lane < 32; ++lane) {
              const int g = lane / G, sub = lane % G;
              int a = zero_a, b = zero_b, c = zero_c;
              const LaneTarget* tg = mine[r][g];
              const int ci = st * G + sub;
              if (tg && st < K && ci < (int)tg->con.size()) {
                a = tg->con[ci][0];
                b = tg->con[ci][1];
                c = tg->con[ci][2];
              }
              pr.words[w].push_back((int32_t)((uint32_t)a | ((uint32_t)b << 16)));
              pr.words[w].push_back((int32_t)c);
            }
        }
      }
      worst = std::max(worst, cost_w);
      total += cost_w;
    }
    return std::max(worst, sum_weight * total) + c_barrier;
  };
  for (const auto& lv : levels) {
    std::vector<int> order(lv.size());
    for (size_t k = 0; k < lv.size(); ++k) order[k] = (int)k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lv[a].con.size() > lv[b].con.size(); });
    int G = G_fixed;
    if (G_fixed == 0) {
      double best = -1.0;
      for (int g = 1; g <= 32; g *= 2) {
        const double c = do_level(lv, order, g, false);
        if (best < 0.0 || c < best) {
          best = c;
          G = g;
        }
      }
    }
    pr.cost += do_level(lv, order, G, emit);
  }
  pr.pad_words = 8 * (PK + 1);  // BO_LP_PAD_PACKETS packets the executor may prefetch but never executes
  if (emit)
    for (int w = 0; w < W; ++w)
      for (int k = 0; k < 8; ++k) {
        header(pr.words[w], 1, 8u, nullptr);
        zero_words(pr.words[w], PK);
      }
  return pr;
}

// `pks`: candidate packet sizes (operand words per packet): long operand lists want large packets (less bookkeeping per
// operand), short ones small packets (less padding)
LaneProgram best_program(const std::vector<std::vector<LaneTarget>>& levels, int max_warps, const std::vector<int>& pks, int zero_a,
                         int zero_b, int zero_c, const char* env_w, const char* env_g, double sum_weight = 0.0,
                         const std::vector<std::vector<LaneTarget>>* levels_aligned = nullptr) {
  int bw = 1, bg = 0, bpk = pks[0];
  bool bal = false;
  double best = -1.0;
  // experiment hooks: force the number of warps / lanes per target of the factor program
  const int force_w = getenv(env_w) ? atoi(getenv(env_w)) : 0;
  const int force_g = getenv(env_g) ? atoi(getenv(env_g)) : 0;
  for (int PK : pks)
    for (int al = 0; al < (levels_aligned ? 2 : 1); ++al)
      for (int W = 1; W <= max_warps; W *= 2)
        for (int G = 0; G <= 32; G = G ? G * 2 : 1) {  // 0 = per-level choice
          if ((force_w && W != force_w) || (force_g && G != (force_g < 0 ? 0 : force_g))) continue;
          const LaneProgram p = schedule_program(al ? *levels_aligned : levels, W, G, PK, zero_a, zero_b, zero_c, false, al != 0, sum_weight);
          if (best < 0.0 || p.cost < best) {
            best = p.cost;
            bw = W;
            bg = G;
            bpk = PK;
            bal = al != 0;
          }
        }
  LaneProgram out = schedule_program(bal ? *levels_aligned : levels, bw, bg, bpk, zero_a, zero_b, zero_c, true, bal, sum_weight);
  out.aligned = bal;
  return out;
}

// table section: [0] W  [1] G | PK << 8  [2] stream offset (absolute, 16-byte aligned)  [3] total steps, then per warp
// { first step, number of steps }
void append_program(std::vector<int32_t>& t, int slot, const LaneProgram& pr) {
  t[slot] = (int32_t)t.size();
  const size_t h = t.size();
  t.resize(h + 4 + 2 * (size_t)pr.W, 0);
  t[h] = pr.W;
  t[h + 1] = pr.G | (pr.PK << 8);  // G = 0: chosen per level (in the packet headers)
  int64_t first = 0;
  for (int w = 0; w < pr.W; ++w) {
    t[h + 4 + 2 * w] = (int32_t)first;
    t[h + 5 + 2 * w] = (int32_t)(pr.words[w].size() / 64) - pr.pad_words;  // words to execute (the trailing padding excluded)
    first += (int64_t)pr.words[w].size() / 64;
  }
  t[h + 3] = (int32_t)first;
  while (t.size() % 4 != 0) t.push_back(0);
  t[h + 2] = (int32_t)t.size();
  for (int w = 0; w < pr.W; ++w) t.insert(t.end(), pr.words[w].begin(), pr.words[w].end());
}

}  // namespace

static CoopPlan make_coop_plan_segments(const ProblemSource& ps, int tpb, int n_segments, bool ldl_only);

// The elimination order is chosen by trial: cutting every chain of the KKT graph into 2 (3) pieces that factor
// independently roughly halves the elimination-tree height for a little extra fill; whichever variant gives the
// shortest lane programs (factorisation + both substitutions, longest warp) and still fits shared memory wins.
// The search below builds up to eight symbolic factorisations (seconds each for C4 / C5), so its verdict -- one integer per
// KKT structure -- is remembered: in the process, and in `cache_dir` (the JIT cache) across processes.
static uint64_t structure_hash(const ProblemSource& ps, int tpb) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](uint64_t v) {
    h ^= v;
    h *= 1099511628211ull;
  };
  mix(0x6c64ull << 8 | 4);  // version of the search (bump when the candidates or the cost model change)
  for (int v : {ps.nx, ps.n_eq, ps.n_ineq, tpb}) mix((uint64_t)v);
  for (const Sparsity* sp : {&ps.hess, &ps.jac_eq, &ps.jac_ineq}) {
    mix((uint64_t)sp->nnz());
    for (int k = 0; k < sp->nnz(); ++k) mix(((uint64_t)(uint32_t)sp->row[k] << 32) | (uint32_t)sp->col[k]);
  }
  return h;
}

CoopPlan make_coop_plan(const ProblemSource& ps, int tpb, const std::string& cache_dir) {
  // candidates: 1 = plain minimum degree; -d = nested dissection of depth d (2^d pieces per chain, minimum-vertex-cover
  // separators, bo_sparse.cpp).  Deeper dissection = lower elimination tree but more fill; the cheapest predicted lane
  // programs (factorisation + both substitutions, cost model of schedule_program) that fit shared memory win.
  // BO_SEGMENTS forces a variant (k > 1: the round-1 flat cut into k pieces).
  int best_seg = 1;
  static std::map<uint64_t, int> memo;
  const uint64_t key = structure_hash(ps, tpb);
  char name[64];
  snprintf(name, sizeof name, "/plan_%016llx.txt", (unsigned long long)key);
  const std::string memo_file = cache_dir.empty() ? std::string() : cache_dir + name;
  bool known = false;
  if (!getenv("BO_SEGMENTS") && !getenv("BO_DEBUG_LDL")) {
    const auto it = memo.find(key);
    if (it != memo.end()) {
      best_seg = it->second;
      known = true;
    } else if (!memo_file.empty()) {
      if (FILE* f = fopen(memo_file.c_str(), "r")) {
        known = fscanf(f, "%d", &best_seg) == 1;
        fclose(f);
      }
    }
  }
  if (getenv("BO_SEGMENTS")) {
    best_seg = atoi(getenv("BO_SEGMENTS"));
  } else if (!known) {
    double best_cost = -1.0;
    // -d: depth d, cuts in the middle; -(10 + d), -(20 + d): cuts of the end intervals biased to 0.6 / 0.7 (less fill)
    bool fits_prev = true;
    for (int seg : {1, -1, -2, -12, -22, -3, -13, -23}) {
      const int d = seg < 0 ? (-seg) % 10 : 0, bias = seg < 0 ? (-seg) / 10 : 0;
      if (bias > 0 && fits_prev) continue;       // the unbiased cut of this depth fitted: nothing to gain from a biased one
      const CoopPlan c = make_coop_plan_segments(ps, tpb, seg, true);
      const size_t smem = ((size_t)c.vals_size() + ps.nx + ps.n_eq + 5 * (tpb / 32) + 8) * sizeof(double);
      const bool fits = seg == 1 || (smem <= 227 * 1024 && c.vals_size() + ps.nx + ps.n_eq + 2 < 32767);
      if (getenv("BO_DEBUG_LDL"))
        fprintf(stderr, "[ldl] candidate %d: %d levels, %d factor values, predicted %.0f cycles, %zu bytes of shared memory%s\n", seg, c.n_levels,
                c.vals_size(), c.prog_cost, smem, fits ? "" : " (does not fit)");
      if (bias == 0) fits_prev = fits;
      else if (fits) fits_prev = true;
      if (!fits) {
        if (bias == 2 || d == 1) break;  // nothing of this depth fits: deeper ones will not either
        continue;
      }
      if (best_cost < 0.0 || c.prog_cost < 0.97 * best_cost) {  // > 3 % cheaper
        best_cost = c.prog_cost;
        best_seg = seg;
      }
    }
  }
  if (!getenv("BO_SEGMENTS")) {
    memo[key] = best_seg;
    if (!known && !memo_file.empty())
      if (FILE* f = fopen(memo_file.c_str(), "w")) {
        fprintf(f, "%d\n", best_seg);
        fclose(f);
      }
  }
  CoopPlan out = make_coop_plan_segments(ps, tpb, best_seg, false);
  if (getenv("BO_DEBUG_LDL"))
    fprintf(stderr, "[ldl] segments %d: %d levels, %d factor values, factor %lld + solve %lld words (longest warp), %d smem doubles; factor W %d G %d PK %d%s, solves G %d / %d PK %d / %d\n", best_seg,
            out.n_levels, out.vals_size(), (long long)out.fac_steps, (long long)out.solve_steps, out.smem_doubles, out.ldl_w, out.ldl_g, out.fac_pk, out.fac_aligned ? " aligned" : "", out.solve_g, out.solve_bwd_g, out.fwd_pk, out.bwd_pk);
  return out;
}

static CoopPlan make_coop_plan_segments(const ProblemSource& ps, int tpb, int n_segments, bool ldl_only) {
  CoopPlan pl;
  pl.n_segments = n_segments;
  // n_segments < 0: nested dissection of depth (-n_segments) % 10, cut bias 0.5 + 0.1 * ((-n_segments) / 10)
  pl.sp = n_segments < 0 ? make_sparse_plan(ps, false, 1, (-n_segments) % 10, 0.5 + 0.1 * ((-n_segments) / 10)) : make_sparse_plan(ps, false, n_segments);
  const SparsePlan& sp = pl.sp;
  const int n = sp.n, nx = ps.nx, nnzL = sp.nnzL();
  std::vector<int32_t>& t = pl.itab;
  t.assign(CT_HEADER, 0);
  t[CT_N] = n;
  t[CT_NNZL] = nnzL;
  auto section = [&](int slot) { t[slot] = (int32_t)t.size(); };
  section(CT_PERM);
  t.insert(t.end(), sp.perm.begin(), sp.perm.end());
  section(CT_SIGN);
  for (int j = 0; j < n; ++j) t.push_back(sp.perm[j] < nx ? 1 : -1);

  // ---- levels of the elimination tree: column j can be computed once every column k with L(j,k) != 0 is done ----
  std::vector<int> lev(n, 0);
  for (int j = 0; j < n; ++j)
    for (int e = sp.colptr[j]; e < sp.colptr[j + 1]; ++e) lev[sp.rowidx[e]] = std::max(lev[sp.rowidx[e]], lev[j] + 1);
  const int nlev = n ? *std::max_element(lev.begin(), lev.end()) + 1 : 0;
  pl.n_levels = nlev;
  t[CT_NLEV] = nlev;
  std::vector<std::vector<int>> cols_of(nlev);
  for (int j = 0; j < n; ++j) cols_of[lev[j]].push_back(j);
  // row pattern: (column k, entry index e) with k < row
  std::vector<std::vector<std::pair<int, int>>> row_pat(n);
  for (int k = 0; k < n; ++k)
    for (int e = sp.colptr[k]; e < sp.colptr[k + 1]; ++e) row_pat[sp.rowidx[e]].push_back({k, e});
  auto entry = [&](int row, int col) -> int {
    const auto lo = sp.rowidx.begin() + sp.colptr[col], hi = sp.rowidx.begin() + sp.colptr[col + 1];
    const auto it = std::lower_bound(lo, hi, row);
    return (it == hi || *it != row) ? -1 : (int)(it - sp.rowidx.begin());
  };

  // ---- lane programs: the factorisation and the two triangular solves, fully pre-scheduled ----
  const int zero_v = sp.vals_size();  // index of a cell of `vals` that always holds 0.0 (padding operand)
  const int zero_b = n;               // same in the permuted right-hand side
  {
    // Column-aligned factor program (default; BO_FAC_ALIGNED=0 = per-target operand lists, G lanes per target): the
    // targets of column j -- D(j) and every C(i,j) -- all walk the SAME list k in rowpat(j); step p of every lane of the
    // column reads C(j,k_p) and 1/D(k_p), one shared-memory word each for the whole column (broadcast), and its own
    // C(i,k_p) (the zero cell where L(i,k_p) is structurally zero: 2.2x the operand slots on C5, a third of the wavefronts).
    const bool fac_aligned = !(getenv("BO_FAC_ALIGNED") && atoi(getenv("BO_FAC_ALIGNED")) == 0);
    std::vector<std::vector<LaneTarget>> fac(nlev), fwd(nlev), bwd(nlev), fac_al(nlev);
    for (int L = 0; L < nlev; ++L) {
      if (fac_aligned) {
        std::vector<int> cols = cols_of[L];
        std::stable_sort(cols.begin(), cols.end(), [&](int a, int b) { return row_pat[a].size() > row_pat[b].size(); });
        for (int j : cols) {
          LaneTarget d;
          d.tgt = j;
          d.positive = sp.perm[j] < nx;
          for (const auto& ke : row_pat[j]) d.con.push_back({n + ke.second, n + ke.second, ke.first});
          fac_al[L].push_back(std::move(d));
          for (int e = sp.colptr[j]; e < sp.colptr[j + 1]; ++e) {
            const int i = sp.rowidx[e];
            LaneTarget o;
            o.tgt = n + e;
            for (const auto& ke : row_pat[j]) {
              const int e_ik = entry(i, ke.first);
              o.con.push_back({e_ik >= 0 ? n + e_ik : zero_v, n + ke.second, ke.first});
            }
            fac_al[L].push_back(std::move(o));
          }
          if (!row_pat[j].empty()) {  // the right-hand side rides along (see below)
            LaneTarget r;
            r.tgt = zero_v + 1 + j;
            for (const auto& ke : row_pat[j]) r.con.push_back({zero_v + 1 + ke.first, n + ke.second, ke.first});
            fac_al[L].push_back(std::move(r));
          }
        }
      }
      for (int j : cols_of[L]) {
        LaneTarget d;
        d.tgt = j;
        d.positive = sp.perm[j] < nx;
        for (const auto& ke : row_pat[j]) d.con.push_back({n + ke.second, n + ke.second, ke.first});
        pl.n_contrib += (int64_t)d.con.size();
        fac[L].push_back(std::move(d));
        for (int e = sp.colptr[j]; e < sp.colptr[j + 1]; ++e) {
          const int i = sp.rowidx[e];
          LaneTarget o;
          o.tgt = n + e;
          for (const auto& ke : row_pat[j]) {
            const int e_ik = entry(i, ke.first);
            if (e_ik >= 0) o.con.push_back({n + e_ik, n + ke.second, ke.first});
          }
          pl.n_contrib += (int64_t)o.con.size();
          fac[L].push_back(std::move(o));
        }
        // The forward substitution rides along with the factorisation: the right-hand side b (the array bp, which sits right
        // behind vals in shared memory: index vals_size + 1 + j) is one more "row" of the matrix, z(j) = b(j) - sum_k
        // C(j,k) (1/D(k)) z(k) one more target of column j's level.  Every factorisation is followed by a solve whose
        // right-hand side is known beforehand, so that solve needs only the scaling by 1/D and the backward program.
        if (!row_pat[j].empty()) {
          LaneTarget r;
          r.tgt = zero_v + 1 + j;
          for (const auto& ke : row_pat[j]) r.con.push_back({zero_v + 1 + ke.first, n + ke.second, ke.first});
          fac[L].push_back(std::move(r));
        }
        LaneTarget f;  // forward: z(j) = b(j) - sum_k C(j,k) u(k)
        f.tgt = j;
        for (const auto& ke : row_pat[j]) f.con.push_back({n + ke.second, ke.first, zero_v});
        fwd[L].push_back(std::move(f));
        LaneTarget w;  // backward: x(j) = u(j) - (1/D(j)) sum_i C(i,j) x(i)
        w.tgt = j;
        for (int e = sp.colptr[j]; e < sp.colptr[j + 1]; ++e) w.con.push_back({n + e, sp.rowidx[e], zero_v});
        bwd[nlev - 1 - L].push_back(std::move(w));
      }
    }
    if (getenv("BO_DEBUG_LDL") && !ldl_only) {
      int64_t tot_t = 0, tot_len = 0, tot_dense = 0, tot_cols = 0;
      for (int L = 0; L < nlev; ++L) {
        int64_t nt = 0, len = 0, dense = 0;
        for (int j : cols_of[L]) {
          const int64_t nj = 1 + sp.colptr[j + 1] - sp.colptr[j], mj = (int64_t)row_pat[j].size();
          nt += nj;
          dense += nj * mj;
        }
        for (const auto& tg : fac[L]) len += (int64_t)tg.con.size();
        if (L % 10 == 0 || L + 1 == nlev)
          fprintf(stderr, "[ldl] level %3d: %3zu columns, %4lld targets, %6lld operands, %6lld column-aligned slots\n", L, cols_of[L].size(),
                  (long long)nt, (long long)len, (long long)dense);
        tot_t += nt; tot_len += len; tot_dense += dense; tot_cols += (int64_t)cols_of[L].size();
      }
      fprintf(stderr, "[ldl] %d levels, %lld columns, %lld targets, %lld operands, %lld column-aligned slots (fill-in of the aligned form %.2f)\n", nlev,
              (long long)tot_cols, (long long)tot_t, (long long)tot_len, (long long)tot_dense, (double)tot_dense / (double)std::max<int64_t>(1, tot_len));
    }
    const int max_warps = std::max(1, tpb / 32);
    // small problems run many CTAs per SM (C3: 16): there the SM's issue slots are shared, the TOTAL instruction count weighs
    // in next to the busiest warp of one CTA
    const double sum_weight = (size_t)(sp.vals_size() + n) * sizeof(double) < 227 * 1024 / 4 ? 0.5 : 0.0;
    // packet sizes: measured on B200 (profiles/r02_coop_phases_*.txt) -- C5 / C4 (operand lists of 27-40 per column): 8 beats 4 by
    // 16 % / 9 %; C3 (3.4 operands per target, 16 CTAs per SM): 2 beats 4 beats 8 (138 k / 186 k / 263 k cycles per
    // factorisation).  The cost model does not resolve that, so the small-problem regime is pinned to 2.
    const std::vector<int> fac_pks = getenv("BO_FAC_PK") ? std::vector<int>{atoi(getenv("BO_FAC_PK"))}
                                                         : (sum_weight > 0.0 ? std::vector<int>{2} : std::vector<int>{4, 8});
    const std::vector<int> solve_pks = getenv("BO_SOLVE_PK") ? std::vector<int>{atoi(getenv("BO_SOLVE_PK"))} : std::vector<int>{1, 2, 4};
    LaneProgram pf = best_program(fac, max_warps, fac_pks, zero_v, zero_v, zero_v, "BO_FAC_WARPS", "BO_FAC_G", sum_weight, fac_aligned ? &fac_al : nullptr);
    pl.fac_aligned = pf.aligned;
    const int solve_warps = getenv("BO_SOLVE_MAX_WARPS") ? atoi(getenv("BO_SOLVE_MAX_WARPS")) : max_warps;
    LaneProgram pw = best_program(fwd, std::min(solve_warps, max_warps), solve_pks, zero_v, zero_b, zero_v, "BO_SOLVE_WARPS", "BO_SOLVE_G", sum_weight);
    LaneProgram pb = best_program(bwd, std::min(solve_warps, max_warps), solve_pks, zero_v, zero_b, zero_v, "BO_SOLVE_WARPS", "BO_SOLVE_G", sum_weight);
    pl.fac_pk = pf.PK;
    pl.fwd_pk = pw.PK;
    pl.bwd_pk = pb.PK;
    pl.ldl_g = pf.G;
    pl.ldl_w = pf.W;
    pl.solve_g = pw.G;
    pl.solve_bwd_g = pb.G;
    pl.prog_cost = pf.cost + pw.cost + pb.cost;
    pl.fac_steps = pf.max_steps();
    pl.solve_steps = pw.max_steps() + pb.max_steps();
    append_program(t, CT_PROG_FAC, pf);
    append_program(t, CT_PROG_FWD, pw);
    append_program(t, CT_PROG_BWD, pb);
  }

  if (ldl_only) return pl;

  // ---- KKT assembly, target-owned: every diagonal position plus every position that receives a term ----
  // term = { kind, a, b, r }: 0: H[a]   1: JE[a]   2: sigma[r] JI[a] JI[b]   3: rho JE[a] JE[b]
  {
    std::map<int, std::vector<int32_t>> terms;
    for (int j = 0; j < n; ++j) terms[j];
    for (int k = 0; k < ps.hess.nnz(); ++k) {
      auto& v = terms[sp.pos(ps.hess.row[k], ps.hess.col[k])];
      v.insert(v.end(), {0, k, 0, 0});
    }
    for (int k = 0; k < ps.jac_eq.nnz(); ++k) {
      auto& v = terms[sp.pos(nx + ps.jac_eq.row[k], ps.jac_eq.col[k])];
      v.insert(v.end(), {1, k, 0, 0});
    }
    auto pairs = [&](const Sparsity& s, int n_rows, int kind) {
      std::vector<std::vector<int>> by_row(n_rows > 0 ? n_rows : 1);
      for (int k = 0; k < s.nnz(); ++k) by_row[s.row[k]].push_back(k);
      for (int r = 0; r < n_rows; ++r) {
        const auto& ks = by_row[r];
        for (size_t u = 0; u < ks.size(); ++u)
          for (size_t w = 0; w < ks.size(); ++w) {
            const int cu = s.col[ks[u]], cw = s.col[ks[w]];
            if (cu < cw || (cu == cw && u != w)) continue;
            auto& v = terms[sp.pos(cu, cw)];
            v.insert(v.end(), {kind, ks[u], ks[w], r});
          }
      }
    };
    pairs(ps.jac_ineq, ps.n_ineq, 2);
    pairs(ps.jac_eq, ps.n_eq, 3);
    t[CT_ANT] = (int32_t)terms.size();
    {
      size_t n_terms = 0;
      for (const auto& kv : terms) n_terms += kv.second.size() / 4;
      pl.asm_terms_per_position = terms.empty() ? 0.0 : (double)n_terms / (double)terms.size();
    }
    std::vector<int32_t> apos, aptr{0}, aterm;
    for (const auto& kv : terms) {
      apos.push_back(kv.first);
      aterm.insert(aterm.end(), kv.second.begin(), kv.second.end());
      aptr.push_back((int32_t)aterm.size() / 4);
    }
    section(CT_APOS);
    t.insert(t.end(), apos.begin(), apos.end());
    section(CT_APTR);
    t.insert(t.end(), aptr.begin(), aptr.end());
    while (t.size() % 4 != 0) t.push_back(0);
    section(CT_ATERM);
    t.insert(t.end(), aterm.begin(), aterm.end());
  }

  // ---- CSR / CSC views of the Jacobians: entries { nz index, other coordinate } ----
  auto views = [&](const Sparsity& s, int n_rows, int slot_rptr) {
    std::vector<std::vector<std::pair<int, int>>> by_row(n_rows > 0 ? n_rows : 1), by_col(nx);
    for (int k = 0; k < s.nnz(); ++k) {
      by_row[s.row[k]].push_back({k, s.col[k]});
      by_col[s.col[k]].push_back({k, s.row[k]});
    }
    auto dump = [&](const std::vector<std::vector<std::pair<int, int>>>& lists, int count, int slot) {
      section(slot);
      int32_t acc = 0;
      t.push_back(0);
      for (int r = 0; r < count; ++r) {
        acc += (int32_t)lists[r].size();
        t.push_back(acc);
      }
      while (t.size() % 2 != 0) t.push_back(0);
      section(slot + 1);
      for (int r = 0; r < count; ++r)
        for (const auto& e : lists[r]) {
          t.push_back(e.first);
          t.push_back(e.second);
        }
    };
    dump(by_row, n_rows, slot_rptr);
    dump(by_col, nx, slot_rptr + 2);
  };
  views(ps.jac_eq, ps.n_eq, CT_JE_RPTR);
  views(ps.jac_ineq, ps.n_ineq, CT_JI_RPTR);

  // ---- tapes: partitioned so that the interpreter's work arrays fit in shared memory ----
  // Shared memory of a CTA (doubles): [ red | ints | vals + 1 | bp + 1 | free ].  While the KKT tape runs the factor is
  // dead, so its work arrays w[slot][thread] may use everything after the ints; the f/c tape runs at trial points with
  // the factor alive and gets the free tail only.
  {
    const int total = 227 * 1024 / 8;
    const int fixed = 5 * (tpb / 32) + 4 + (sp.vals_size() + 1) + (n + 1);
    const int kkt_cap = total - (5 * (tpb / 32) + 4);
    const int fc_cap = total - fixed;
    auto fit = [&](const Tape& tape, int cap, int* wstride) {
      int nsub = tpb;
      for (;;) {
        PartTape pt = partition_tape(tape, nsub);
        const int stride = ((pt.info.nsub + 31) / 32) * 32;
        if ((int64_t)pt.n_work * stride <= cap) {
          *wstride = stride;
          return pt;
        }
        if (nsub <= 32) {
          *wstride = 0;  // does not fit even with one warp of sub-tapes: thread-local work array
          return partition_tape(tape, tpb);
        }
        nsub = std::max(32, std::min(nsub - 32, (cap / std::max(pt.n_work, 1)) / 32 * 32));
      }
    };
    if (getenv("BO_DEBUG_CLASSES")) {
      for (const Tape* tp : {&ps.fc, &ps.kkt}) {
        PartTape all = partition_tape(*tp, 1 << 30);
        std::vector<TapeClass> cl = classify(all, *tp);
        int64_t code = 0, work = 0;
        for (const TapeClass& c : cl) {
          code += (int64_t)c.rows.size() / 4;
          work += (int64_t)c.rows.size() / 4 * (((int64_t)c.members.size() + 31) / 32);
        }
        fprintf(stderr, "[classes] %zu components, %zu classes, %lld code rows, %lld warp-rows, %d partial slots\n", all.subs.size(), cl.size(),
                (long long)code, (long long)work, all.n_part);
        std::vector<const TapeClass*> order;
        for (const TapeClass& c : cl) order.push_back(&c);
        std::sort(order.begin(), order.end(), [](const TapeClass* a, const TapeClass* b) { return a->rows.size() * a->members.size() > b->rows.size() * b->members.size(); });
        for (size_t k = 0; k < order.size() && k < 12; ++k)
          fprintf(stderr, "    class rows %zu members %zu var_rows %zu\n", order[k]->rows.size() / 4, order[k]->members.size(), order[k]->var_rows.size());
      }
    }
    bool gen = getenv("BO_NO_GEN_TAPES") == nullptr;
    if (gen) {
      PartTape fca = partition_tape(ps.fc, 1 << 30), kka = partition_tape(ps.kkt, 1 << 30);
      std::vector<TapeClass> fcc = classify(fca, ps.fc), kkc = classify(kka, ps.kkt);
      int64_t code = 0;
      for (const TapeClass& c : fcc) code += (int64_t)c.rows.size() / 4;
      for (const TapeClass& c : kkc) code += (int64_t)c.rows.size() / 4;
      if (code <= 40000 && fcc.size() + kkc.size() <= 400) {
        pl.gen_tapes = true;
        pl.fc = fca.info;
        pl.kkt = kka.info;
        pl.gen_code = append_gen_tape(t, CT_TAPE_FC, fca, fcc, ps.fc, pl.dtab, tpb / 32, "fc", &pl.fc);
        pl.gen_code += append_gen_tape(t, CT_TAPE_KKT, kka, kkc, ps.kkt, pl.dtab, tpb / 32, "kkt", &pl.kkt);
        pl.n_work_pre = std::max(fca.pre.n_work, kka.pre.n_work);
        pl.smem_doubles = fixed;
      } else {
        gen = false;
      }
    }
    if (!gen) {
    PartTape fc = fit(ps.fc, fc_cap, &pl.fc_wstride);
    PartTape kkt = fit(ps.kkt, kkt_cap, &pl.kkt_wstride);
    pl.fc = fc.info;
    pl.kkt = kkt.info;
    pl.n_work_fc = fc.n_work;
    pl.n_work_kkt = kkt.n_work;
    pl.n_work_pre = std::max(fc.pre.n_work, kkt.pre.n_work);
    append_tape_section(t, CT_TAPE_FC, fc, ps.fc, pl.dtab);
    append_tape_section(t, CT_TAPE_KKT, kkt, ps.kkt, pl.dtab);
    const int fc_w = pl.fc_wstride * pl.n_work_fc, kkt_w = pl.kkt_wstride * pl.n_work_kkt;
    pl.smem_doubles = std::max(fixed + fc_w, 5 * (tpb / 32) + 4 + kkt_w);
    }
  }
  return pl;
}

size_t coop_scratch_doubles(const ProblemSource& ps, const CoopPlan& pl) {
  const size_t nx = ps.nx, me = ps.n_eq, mi = ps.n_ineq;
  return ps.np + 6 * nx + 6 * me + 10 * mi + ps.jac_eq.nnz() + ps.jac_ineq.nnz() + ps.hess.nnz() + (nx + me) + pl.fc.n_pe +
         pl.kkt.n_pe + std::max(pl.fc.n_part, pl.kkt.n_part) + 8;
}

std::string emit_coop_source(const ProblemSource& ps, const CoopPlan& pl, int tpb) {
  std::ostringstream o;
  o << "// generated by libb200optas (bo_coop.cpp): cooperative tier, one instance per CTA, fully table-driven\n";
  o << "#define BO_NX " << ps.nx << "\n#define BO_NP " << ps.np << "\n#define BO_ME " << ps.n_eq << "\n#define BO_MI "
    << ps.n_ineq << "\n#define BO_NNZ_JE " << ps.jac_eq.nnz() << "\n#define BO_NNZ_JI " << ps.jac_ineq.nnz()
    << "\n#define BO_NNZ_H " << ps.hess.nnz() << "\n#define BO_TPB " << tpb << "\n";
  o << "#define BO_COOP 1\n#define BO_VALS " << pl.vals_size() << "\n#define BO_NWORK_FC " << pl.n_work_fc << "\n#define BO_NWORK_KKT "
    << pl.n_work_kkt << "\n#define BO_NWORK_PRE " << pl.n_work_pre << "\n#define BO_FC_WSTRIDE " << pl.fc_wstride
    << "\n#define BO_KKT_WSTRIDE " << pl.kkt_wstride << "\n#define BO_SMEM_DOUBLES " << pl.smem_doubles << "\n#define BO_NPE_FC "
    << pl.fc.n_pe << "\n#define BO_NPE_KKT " << pl.kkt.n_pe << "\n#define BO_NPART " << std::max(pl.fc.n_part, pl.kkt.n_part) << "\n";
  {
    // resident CTAs per SM the kernel is compiled for (register cap): as many as the shared memory allows, at least
    // 64 registers per thread
    const size_t smem_total = ((size_t)pl.smem_doubles + (pl.w_in_smem ? coop_scratch_doubles(ps, pl) : 0)) * 8;
    const int by_smem = (int)(227 * 1024 / std::max<size_t>(1, smem_total + 1024));
    const int by_regs = 65536 / (tpb * 64);
    const int min_ctas = std::max(1, std::min(std::min(by_smem, by_regs), 32));
    o << "#define BO_MIN_CTAS " << min_ctas << "\n#define BO_FAC_G " << pl.ldl_g << "\n#define BO_FWD_G " << pl.solve_g
      << "\n#define BO_BWD_G " << pl.solve_bwd_g << "\n#define BO_FAC_PK " << pl.fac_pk << "\n#define BO_FWD_PK " << pl.fwd_pk << "\n#define BO_BWD_PK " << pl.bwd_pk << "\n";
  }
  if (!hessian_depends_on_eq_multipliers(ps)) o << "#define BO_RECALC_DC_ONLY 1\n";
  if (ps.nx + ps.n_eq >= 600) o << "#define BO_RECALC_SKIP_DEGENERATE 1\n";  // measured threshold, see bo_ipm_cta.cuh
  // a failing factorisation is abandoned at the level of its first bad pivot -- where the factor fills shared memory (one
  // CTA per SM); small problems (many CTAs per SM, few failures) keep the plain level barrier
  if ((size_t)(pl.vals_size() + ps.nx + ps.n_eq) * sizeof(double) >= 227 * 1024 / 4) o << "#define BO_FAC_ABANDON 1\n";
  // KKT assembly: four positions per thread in flight where the term lists are short (C5: 117 k -> 66 k cycles per
  // iteration), four terms of one position where they are long (C4, C3: the interleaved form is 14 % / 50 % slower there)
  if (pl.asm_terms_per_position < 2.0 && (size_t)(pl.vals_size() + ps.nx + ps.n_eq) * sizeof(double) >= 227 * 1024 / 4)
    o << "#define BO_ASM_INTERLEAVE 1\n";
  if (pl.w_in_smem) o << "#define BO_W_IN_SMEM 1\n";
  o << "#include \"bo_common.cuh\"\n";
  if (pl.gen_tapes) o << "#define BO_GEN_TAPES 1\n" << pl.gen_code;
  o << "#include \"bo_ipm_cta.cuh\"\n";
  return o.str();
}

}  // namespace bo
