// slp::Jacobian — row descriptors of ∂variables/∂wrt.
//
// Constructor logic follows the reference's autodiff/jacobian.hpp:54-105: one
// parent→child list per row, (column, node) output lists found by tagging the
// wrt leaves through `scratch` AFTER sorting, LINEAR rows evaluated once and
// cached, QUADRATIC/NONLINEAR rows recorded for re-evaluation. Inside
// Problem::solve these lists are uploaded through slpb_upload_rows and the
// per-iteration sweeps run on the device; value() and get() (:107-156) are the
// reference's host walks, for code that uses the autodiff classes on their
// own outside a solve.
#pragma once

#include <utility>
#include <vector>

#include "sleipnir/autodiff/expression_graph.hpp"
#include "sleipnir/autodiff/variable.hpp"
#include "sleipnir/autodiff/variable_matrix.hpp"
#include "sleipnir/util/linalg.hpp"

namespace slp {

template <typename Scalar>
class Jacobian {
 public:
  Jacobian(Variable<Scalar> variable, Variable<Scalar> wrt)
      : Jacobian{VariableMatrix<Scalar>{std::move(variable)},
                 VariableMatrix<Scalar>{std::move(wrt)}} {}
  Jacobian(Variable<Scalar> variable, VariableMatrix<Scalar> wrt)
      : Jacobian{VariableMatrix<Scalar>{std::move(variable)}, std::move(wrt)} {}
  Jacobian(VariableMatrix<Scalar> variables, VariableMatrix<Scalar> wrt)
      : m_variables{std::move(variables)}, m_wrt{std::move(wrt)} {
    slp_assert(m_variables.cols() == 1 || m_variables.size() == 0);
    slp_assert(m_wrt.cols() == 1);
    init();
  }

  virtual ~Jacobian() = default;
  int rows() const { return m_variables.rows(); }
  int cols() const { return m_wrt.rows(); }
  const VariableMatrix<Scalar>& variables() const { return m_variables; }
  const std::vector<detail::ExpressionGraph>& top_lists() const {
    return m_top_lists;
  }
  const std::vector<std::vector<std::pair<int, detail::ExprId>>>& output_lists()
      const {
    return m_output_lists;
  }
  /// Triplets of the LINEAR rows (constant for the whole solve).
  const std::vector<detail::Triplet>& cached_triplets() const {
    return m_cached_triplets;
  }
  /// Rows whose gradients must be recomputed at every evaluation.
  const std::vector<int>& nonlinear_rows() const { return m_nonlinear_rows; }

  /// The Jacobian as a matrix of expressions (jacobian.hpp:107-127): row by
  /// row the symbolic gradient of the row's variable.
  VariableMatrix<Scalar> get() const {
    VariableMatrix<Scalar> result{detail::empty, m_variables.size(),
                                  m_wrt.size()};
    for (int row = 0; row < m_variables.size(); ++row) {
      const auto grad = detail::gradient_tree(m_top_lists[row], m_wrt);
      for (int col = 0; col < m_wrt.size(); ++col) {
        if (grad(col).expr != nullptr) {
          result(row, col) = grad(col);
        } else {
          result(row, col) = Variable<Scalar>{Scalar(0)};
        }
      }
    }
    return result;
  }

  /// Evaluates the Jacobian at wrt's current values (jacobian.hpp:134-156):
  /// values of every row's graph, then one reverse sweep per non-linear row
  /// on top of the cached triplets of the linear rows; duplicates summed,
  /// column-major, like Eigen's setFromTriplets.
  const SparseMatrix<Scalar>& value() {
    if (m_valued && m_nonlinear_rows.empty()) return m_J;
    for (const auto& list : m_top_lists) detail::update_values(list);
    std::vector<detail::Triplet> triplets = m_cached_triplets;
    std::vector<double> adjoint;
    for (int row : m_nonlinear_rows) {
      detail::append_triplets(m_top_lists[row], m_output_lists[row], adjoint,
                              triplets, row);
    }
    m_J = SparseMatrix<Scalar>::from_triplets(
        m_variables.size(), m_wrt.size(), triplets,
        [this](int r, int c) { return keep(r, c); });
    m_valued = true;
    return m_J;
  }

 protected:
  struct deferred_t {};
  Jacobian(deferred_t, VariableMatrix<Scalar> variables,
           VariableMatrix<Scalar> wrt)
      : m_variables{std::move(variables)}, m_wrt{std::move(wrt)} {}

  void init() {
    auto& scratch = detail::pool().scratch;
    m_top_lists = detail::topological_sort_rows(
        m_variables.size(),
        [&](int row) -> const detail::Expr& { return m_variables(row).expr; });
    for (int col = 0; col < m_wrt.size(); ++col) {
      scratch[m_wrt(col).expr.id()] = col;
    }
    m_output_lists.reserve(m_top_lists.size());
    for (const auto& list : m_top_lists) {
      auto& outs = m_output_lists.emplace_back();
      for (detail::ExprId node : list) {
        if (scratch[node] != -1) outs.emplace_back(scratch[node], node);
      }
    }
    for (auto& v : m_wrt) scratch[v.expr.id()] = -1;

    std::vector<double> adjoint;
    for (int row = 0; row < m_variables.size(); ++row) {
      if (m_variables(row).expr == nullptr) continue;
      const auto type = m_variables(row).type();
      if (type == ExpressionType::LINEAR) {
        detail::append_triplets(m_top_lists[row], m_output_lists[row], adjoint,
                                m_cached_triplets, row);
      } else if (type > ExpressionType::LINEAR) {
        m_nonlinear_rows.push_back(row);
      }
    }
  }

  /// Which entries value() keeps (Hessian<Lower>: row ≥ col).
  virtual bool keep(int, int) const { return true; }

  VariableMatrix<Scalar> m_variables;
  VariableMatrix<Scalar> m_wrt;
  SparseMatrix<Scalar> m_J;
  bool m_valued = false;
  std::vector<detail::ExpressionGraph> m_top_lists;
  std::vector<std::vector<std::pair<int, detail::ExprId>>> m_output_lists;
  std::vector<detail::Triplet> m_cached_triplets;
  std::vector<int> m_nonlinear_rows;
};

}  // namespace slp
