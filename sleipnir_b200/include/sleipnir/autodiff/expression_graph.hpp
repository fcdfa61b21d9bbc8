// Host-side graph walks over the flat expression pool.
//
// Same ordering contract as the reference's expression_graph.hpp:
// topological_sort (:28-78) yields the parent→child list whose order fixes the
// accumulation order of every adjoint; update_values (:85-96) and
// append_triplets (:106-153) are kept for the work the reference also does on
// the host outside the Newton loop — Variable::value() queries, the one-time
// evaluation of LINEAR rows (jacobian.hpp:84-89) and bound detection. The
// per-iteration sweeps run on the device from the same lists.
#pragma once

#include <algorithm>
#include <thread>
#include <utility>
#include <vector>

#include "sleipnir/autodiff/expression.hpp"

namespace slp::detail {

/// Parent→child list of node ids.
using ExpressionGraph = std::vector<ExprId>;

struct Triplet {
  int32_t row, col;
  double value;
};

/// Sorts with caller-provided scratch (all −1 on entry and on exit) and stack,
/// so that independent rows can be sorted by different threads: the graph is
/// only read.
inline ExpressionGraph topological_sort(const ExpressionPool& P, ExprId root_id,
                                        std::vector<int32_t>& scratch,
                                        std::vector<ExprId>& stack) {
  ExpressionGraph list;
  stack.clear();
  // Pass 1: count incoming edges (offset by −1) by DFS from the root.
  stack.push_back(root_id);
  while (!stack.empty()) {
    ExprId node = stack.back();
    stack.pop_back();
    for (ExprId arg : {P.lhs[node], P.rhs[node]}) {
      if (arg != kNull && ++scratch[arg] == 0) stack.push_back(arg);
    }
  }
  // Pass 2: emit a node once all of its parents have been emitted.
  stack.push_back(root_id);
  while (!stack.empty()) {
    ExprId node = stack.back();
    stack.pop_back();
    list.push_back(node);
    for (ExprId arg : {P.lhs[node], P.rhs[node]}) {
      if (arg != kNull && --scratch[arg] == -1) stack.push_back(arg);
    }
  }
  return list;
}

inline ExpressionGraph topological_sort(const Expr& root) {
  if (root == nullptr || root.type() == ExpressionType::CONSTANT) return {};
  std::vector<ExprId> stack;
  return topological_sort(pool(), root.id(), pool().scratch, stack);
}

/// Row-wise topological sorts (jacobian.hpp:55-57 of the reference does them
/// one after the other). The rows are independent and the graph is read-only
/// here, so large row sets are sorted by several host threads, each with its
/// own in-degree scratch; every list is identical to the sequential one.
template <typename RootOf>
std::vector<ExpressionGraph> topological_sort_rows(int n_rows, RootOf&& root_of) {
  std::vector<ExpressionGraph> lists(n_rows);
  // the pool is thread_local (like the reference's, src/util/pool.cpp:5-8):
  // resolve everything that needs it on this thread and hand the workers a
  // pointer to it
  ExpressionPool& P = pool();
  std::vector<ExprId> roots(n_rows, kNull);
  for (int row = 0; row < n_rows; ++row) {
    const Expr& root = root_of(row);
    if (root == nullptr || root.type() == ExpressionType::CONSTANT) continue;
    roots[row] = root.id();
  }
  const unsigned hw = std::thread::hardware_concurrency();
  const int n_threads =
      n_rows < 2048 ? 1 : static_cast<int>(std::min(8u, std::max(1u, hw)));
  auto work = [&](int t, std::vector<int32_t>& scratch) {
    std::vector<ExprId> stack;
    const int b = static_cast<int>(int64_t(n_rows) * t / n_threads);
    const int e = static_cast<int>(int64_t(n_rows) * (t + 1) / n_threads);
    for (int row = b; row < e; ++row) {
      if (roots[row] == kNull) continue;
      lists[row] = topological_sort(P, roots[row], scratch, stack);
    }
  };
  if (n_threads == 1) {
    work(0, P.scratch);
    return lists;
  }
  const size_t pool_size = P.size();
  std::vector<std::thread> threads;
  for (int t = 1; t < n_threads; ++t) {
    threads.emplace_back([&, t] {
      std::vector<int32_t> scratch(pool_size, -1);
      work(t, scratch);
    });
  }
  work(0, P.scratch);
  for (auto& th : threads) th.join();
  return lists;
}

inline void update_values(const ExpressionGraph& list) {
  auto& P = pool();
  for (auto it = list.rbegin(); it != list.rend(); ++it) {
    const ExprId node = *it;
    const ExprId l = P.lhs[node], r = P.rhs[node];
    if (l != kNull) {
      P.val[node] = op_value(static_cast<Op>(P.op[node]), P.val[l],
                             r != kNull ? P.val[r] : 0.0);
    }
  }
}

/// One reverse sweep. `adjoint` is caller-provided scratch indexed by node id
/// (only entries of nodes in `top_list` are touched).
inline void append_triplets(const ExpressionGraph& top_list,
                            const std::vector<std::pair<int, ExprId>>& outputs,
                            std::vector<double>& adjoint,
                            std::vector<Triplet>& triplets, int row) {
  if (top_list.empty()) return;
  auto& P = pool();
  if (adjoint.size() < P.size()) adjoint.resize(P.size());
  adjoint[top_list[0]] = 1.0;
  for (size_t i = 1; i < top_list.size(); ++i) adjoint[top_list[i]] = 0.0;
  for (ExprId node : top_list) {
    const ExprId l = P.lhs[node], r = P.rhs[node];
    if (l == kNull) continue;
    const Op op = static_cast<Op>(P.op[node]);
    const double a = adjoint[node];
    if (r != kNull) {
      const double lv = P.val[l], rv = P.val[r];
      adjoint[l] += op_grad_l(op, a, lv, rv);
      adjoint[r] += op_grad_r(op, a, lv, rv);
    } else {
      adjoint[l] += op_grad_l(op, a, P.val[l], 0.0);
    }
  }
  for (const auto& [col, node] : outputs) {
    triplets.push_back({row, col, adjoint[node]});
  }
}

}  // namespace slp::detail
