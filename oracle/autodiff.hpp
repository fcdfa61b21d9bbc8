// ORACLE — TEST INFRASTRUCTURE ONLY (see expr.hpp header).
//
// Restatement of include/sleipnir/autodiff/jacobian.hpp:54-156,
// hessian.hpp:49-157 and gradient.hpp:53-57: per-row top-lists, (col,node)
// output lists, LINEAR rows cached at construction, value() = forward sweep of
// EVERY row list + reverse sweep of the non-linear rows + setFromTriplets.
#pragma once

#include <utility>
#include <vector>

#include "sparse.hpp"
#include "var.hpp"

namespace orc {

template <class B>
class Jacobian {
 public:
  using Raw = typename B::Raw;
  using Trip = typename B::Trip;

  Jacobian(Mat<B> variables, Mat<B> wrt)
      : m_variables{std::move(variables)}, m_wrt{std::move(wrt)} {
    init();
  }

  const Csc& value() {
    if (m_nonlinear_rows.empty()) return m_J;
    for (auto& list : m_top_lists) B::update(list);
    std::vector<Trip> triplets = m_cached_triplets;
    for (int row : m_nonlinear_rows) {
      B::triplets(m_top_lists[row], m_output_lists[row], triplets, row);
    }
    m_J = finish(triplets);
    return m_J;
  }

  // Introspection for the parity tests / tape export.
  const std::vector<typename B::Graph>& top_lists() const {
    return m_top_lists;
  }
  const std::vector<std::vector<std::pair<int, Raw*>>>& output_lists() const {
    return m_output_lists;
  }
  const std::vector<int>& nonlinear_rows() const { return m_nonlinear_rows; }
  const std::vector<Trip>& cached_triplets() const { return m_cached_triplets; }
  const Mat<B>& variables() const { return m_variables; }

 protected:
  struct deferred_t {};
  Jacobian(deferred_t, Mat<B> variables, Mat<B> wrt)
      : m_variables{std::move(variables)}, m_wrt{std::move(wrt)} {}

  virtual Csc finish(const std::vector<Trip>& t) const {
    return Csc::from_triplets(m_variables.rows(), m_wrt.rows(), t);
  }

  void init() {
    for (auto& v : m_variables) m_top_lists.emplace_back(B::sort(v.expr));
    // Tag wrt columns only AFTER sorting (scratch doubles as the sort's
    // in-degree counter), then reset (jacobian.hpp:60-82).
    for (int col = 0; col < m_wrt.size(); ++col) {
      m_wrt[col].expr->scratch = col;
    }
    for (auto& list : m_top_lists) {
      m_output_lists.emplace_back();
      for (auto* node : list) {
        if (node->scratch != -1) {
          m_output_lists.back().emplace_back(node->scratch, node);
        }
      }
    }
    for (auto& v : m_wrt) v.expr->scratch = -1;

    constexpr int LINEAR = 2;
    for (int row = 0; row < m_variables.rows(); ++row) {
      if (m_variables[row].expr == nullptr) continue;
      int t = m_variables[row].type();
      if (t == LINEAR) {
        B::triplets(m_top_lists[row], m_output_lists[row], m_cached_triplets,
                    row);
      } else if (t > LINEAR) {
        m_nonlinear_rows.push_back(row);
      }
    }
    if (m_nonlinear_rows.empty()) m_J = finish(m_cached_triplets);
  }

  Mat<B> m_variables;
  Mat<B> m_wrt;
  std::vector<typename B::Graph> m_top_lists;
  std::vector<std::vector<std::pair<int, Raw*>>> m_output_lists;
  Csc m_J;
  std::vector<Trip> m_cached_triplets;
  std::vector<int> m_nonlinear_rows;
};

/// Hessian<Lower>: Jacobian of the symbolic gradient, lower-triangle filtered
/// (hessian.hpp:49-52,151-154).
template <class B>
class Hessian : public Jacobian<B> {
 public:
  using Base = Jacobian<B>;
  Hessian(Var<B> variable, Mat<B> wrt)
      : Base{typename Base::deferred_t{},
             gradient_tree<B>(B::sort(variable.expr), wrt), wrt} {
    this->init();
  }

 protected:
  Csc finish(const std::vector<typename B::Trip>& t) const override {
    return Csc::from_triplets(this->m_variables.rows(), this->m_wrt.rows(), t)
        .lower();
  }
};

/// Gradient = 1-row Jacobian, returned dense (gradient.hpp:53-57).
template <class B>
class Gradient {
 public:
  Gradient(Var<B> variable, Mat<B> wrt)
      : m_n{wrt.rows()}, m_jac{Mat<B>{std::move(variable)}, std::move(wrt)} {}
  Vec value() {
    const Csc& J = m_jac.value();
    Vec g(m_n, 0.0);
    for (int c = 0; c < J.cols; ++c) {
      for (int k = J.colptr[c]; k < J.colptr[c + 1]; ++k) g[c] = J.val[k];
    }
    return g;
  }
  Jacobian<B>& jacobian() { return m_jac; }

 private:
  int m_n;
  Jacobian<B> m_jac;
};

}  // namespace orc
