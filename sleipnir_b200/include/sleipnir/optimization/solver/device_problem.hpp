// Flattening of the host expression graphs into the upload format of the C ABI
// (include/slpb.h: slpb_upload_tape / slpb_upload_rows).
//
// This is the hand-over point that replaces the eight std::function matrix
// callbacks the reference builds in Problem::solve (problem.hpp:618-660): the
// same Gradient / Hessian / Jacobian objects (problem.hpp:535-560) are
// constructed on the host, and their per-row lists are shipped once instead of
// being re-walked on the CPU at every callback invocation.
#pragma once

#include <cstdint>
#include <utility>
#include <vector>

#include "sleipnir/autodiff/gradient.hpp"
#include "sleipnir/autodiff/hessian.hpp"
#include "sleipnir/autodiff/jacobian.hpp"
#include "sleipnir/autodiff/variable_matrix.hpp"
#include "slpb.h"

namespace slp::detail {

/// Owning storage behind one slpb_rowset.
struct FlatRows {
  int32_t n_rows = 0, n_cols = 0;
  std::vector<int32_t> row_ptr{0}, row_nodes, out_ptr{0}, out_col, out_node;
  std::vector<uint8_t> row_swept;
  std::vector<int32_t> cached_row, cached_col;
  std::vector<double> cached_val;
  std::vector<double> const_val;

  slpb_rowset view() const {
    slpb_rowset r{};
    r.n_rows = n_rows;
    r.n_cols = n_cols;
    r.row_ptr = row_ptr.data();
    r.row_nodes = row_nodes.data();
    r.out_ptr = out_ptr.data();
    r.out_col = out_col.data();
    r.out_node = out_node.data();
    r.row_swept = row_swept.data();
    r.n_cached = static_cast<int32_t>(cached_row.size());
    r.cached_row = cached_row.data();
    r.cached_col = cached_col.data();
    r.cached_val = cached_val.data();
    return r;
  }
};

/// Everything slpb_upload_* needs, with node ids renumbered densely over the
/// nodes that are actually referenced.
struct FlatProblem {
  std::vector<uint8_t> op;
  std::vector<int32_t> lhs, rhs;
  std::vector<double> val;
  std::vector<int32_t> leaf_x, leaf_y, leaf_z;
  FlatRows rows[SLPB_OUT_COUNT];
  int32_t n_nodes() const { return static_cast<int32_t>(op.size()); }
};

class Flattener {
 public:
  /// Records the rows of a derivative output.
  template <typename Scalar>
  void add_jacobian(int which, const Jacobian<Scalar>& J) {
    FlatRows& R = m_out.rows[which];
    R.n_rows = J.rows();
    R.n_cols = J.cols();
    R.row_swept.assign(R.n_rows, 0);
    for (int row : J.nonlinear_rows()) R.row_swept[row] = 1;
    const auto& lists = J.top_lists();
    const auto& outs = J.output_lists();
    size_t total_nodes = 0, total_outs = 0;
    for (int r = 0; r < R.n_rows; ++r) {
      if (!R.row_swept[r]) continue;
      total_nodes += lists[r].size();
      total_outs += outs[r].size();
    }
    R.row_nodes.reserve(total_nodes);
    R.out_col.reserve(total_outs);
    R.out_node.reserve(total_outs);
    R.row_ptr.reserve(R.n_rows + 1);
    R.out_ptr.reserve(R.n_rows + 1);
    ensure_touched();
    for (int r = 0; r < R.n_rows; ++r) {
      // Only swept rows need their lists on the device; LINEAR rows travel as
      // cached triplets.
      if (R.row_swept[r]) {
        R.row_nodes.insert(R.row_nodes.end(), lists[r].begin(), lists[r].end());
        for (ExprId id : lists[r]) m_touched[id] = 1;
        for (const auto& [col, id] : outs[r]) {
          R.out_col.push_back(col);
          R.out_node.push_back(id);
          m_touched[id] = 1;
        }
      }
      R.row_ptr.push_back(static_cast<int32_t>(R.row_nodes.size()));
      R.out_ptr.push_back(static_cast<int32_t>(R.out_col.size()));
    }
    for (const auto& t : J.cached_triplets()) {
      R.cached_row.push_back(t.row);
      R.cached_col.push_back(t.col);
      R.cached_val.push_back(t.value);
    }
  }

  /// Records value rows (f, c_e, c_i). A CONSTANT root has an empty list
  /// (expression_graph.hpp:32-35); its value is shipped instead.
  template <typename Scalar>
  void add_values(int which, const VariableMatrix<Scalar>& vars) {
    FlatRows& R = m_out.rows[which];
    R.n_rows = vars.size();
    R.n_cols = 1;
    R.const_val.assign(R.n_rows, 0.0);
    for (int r = 0; r < R.n_rows; ++r) {
      const auto& e = vars(r).expr;
      ExpressionGraph list = topological_sort(e);
      if (list.empty()) {
        R.const_val[r] = e == nullptr ? 0.0 : e.val();
      }
      for (ExprId id : list) R.row_nodes.push_back(touch(id));
      R.row_ptr.push_back(static_cast<int32_t>(R.row_nodes.size()));
      R.out_ptr.push_back(0);
    }
  }

  template <typename Scalar>
  void set_leaves(const VariableMatrix<Scalar>& x,
                  const VariableMatrix<Scalar>& y,
                  const VariableMatrix<Scalar>& z) {
    for (const auto& v : x) m_leaf_x.push_back(touch(v.expr.id()));
    for (const auto& v : y) m_leaf_y.push_back(touch(v.expr.id()));
    for (const auto& v : z) m_leaf_z.push_back(touch(v.expr.id()));
  }

  /// Renumbers the touched nodes in increasing pool order (children were
  /// created before their parents, so this is child-before-parent) and emits
  /// the tape.
  FlatProblem finish() {
    auto& P = pool();
    std::vector<int32_t> new_id(P.size(), -1);
    int32_t count = 0;
    for (size_t i = 0; i < P.size(); ++i) {
      if (i < m_touched.size() && m_touched[i]) new_id[i] = count++;
    }
    FlatProblem& F = m_out;
    F.op.resize(count);
    F.lhs.resize(count);
    F.rhs.resize(count);
    F.val.resize(count);
    for (size_t i = 0; i < P.size(); ++i) {
      const int32_t id = new_id[i];
      if (id < 0) continue;
      F.op[id] = P.op[i];
      F.lhs[id] = P.lhs[i] == kNull ? -1 : new_id[P.lhs[i]];
      F.rhs[id] = P.rhs[i] == kNull ? -1 : new_id[P.rhs[i]];
      F.val[id] = P.val[i];
    }
    auto remap = [&](std::vector<int32_t>& v) {
      for (auto& id : v) id = new_id[id];
    };
    for (auto& R : F.rows) {
      remap(R.row_nodes);
      remap(R.out_node);
    }
    F.leaf_x = m_leaf_x;
    F.leaf_y = m_leaf_y;
    F.leaf_z = m_leaf_z;
    remap(F.leaf_x);
    remap(F.leaf_y);
    remap(F.leaf_z);
    // A Variable that no Problem::decision_variable() created (the initial
    // state of a single-shooting OCP, ocp.hpp:171-174) is a leaf nobody
    // differentiates with respect to or moves during the solve: the device
    // sees it as a constant holding its current value.
    {
      std::vector<uint8_t> is_leaf(count, 0);
      for (const auto* list : {&F.leaf_x, &F.leaf_y, &F.leaf_z}) {
        for (int32_t id : *list) is_leaf[id] = 1;
      }
      for (int32_t id = 0; id < count; ++id) {
        if (F.op[id] == SLPB_OP_VAR && !is_leaf[id]) F.op[id] = SLPB_OP_CONST;
      }
    }
    return std::move(m_out);
  }

 private:
  /// Marks a node (and, for list members, implicitly its children — every list
  /// is closed under "child of") as referenced; returns the pool id for now.
  int32_t touch(ExprId id) {
    if (static_cast<size_t>(id) >= m_touched.size()) {
      m_touched.resize(pool().size(), 0);
    }
    m_touched[id] = 1;
    return id;
  }
  void ensure_touched() {
    if (m_touched.size() < pool().size()) m_touched.resize(pool().size(), 0);
  }

  FlatProblem m_out;
  std::vector<uint8_t> m_touched;
  std::vector<int32_t> m_leaf_x, m_leaf_y, m_leaf_z;
};

}  // namespace slp::detail
