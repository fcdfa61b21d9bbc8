// ORACLE — TEST INFRASTRUCTURE ONLY (see expr.hpp header).
//
// Tier-A backend: the reference's OWN autodiff core, compiled where it lies
// under /root/reference (never copied): sleipnir/autodiff/expression.hpp,
// expression_type.hpp, expression_graph.hpp, util/intrusive_shared_ptr.hpp,
// util/pool.hpp and src/util/pool.cpp, through the two shims in oracle/shim/.
// Only built in the authoring container (oracle/Makefile target `ref`, output
// oracle/_ref/liboracle_ref.so); the GPU box uses the prebuilt .so.
#pragma once

#include <utility>
#include <vector>

#include "sleipnir/autodiff/expression_graph.hpp"

namespace orc {

struct RefBackend {
  using Ptr = slp::detail::ExpressionPtr<double>;
  using Raw = slp::detail::Expression<double>;
  using Graph = slp::detail::ExpressionGraph<double>;
  using Trip = Eigen::Triplet<double>;
  static Ptr make_var() {
    return slp::detail::make_expression_ptr<
        slp::detail::DecisionVariableExpression<double>>();
  }
  static Ptr make_const(double v) { return slp::detail::constant_ptr(v); }
  static Graph sort(const Ptr& root) {
    return slp::detail::topological_sort(root);
  }
  static void update(const Graph& g) { slp::detail::update_values(g); }
  static void triplets(const Graph& g,
                       const std::vector<std::pair<int, Raw*>>& outs,
                       std::vector<Trip>& t, int row) {
    slp::detail::append_triplets(g, outs, t, row);
  }
  static int type_rank(const Ptr& p) { return static_cast<int>(p->type()); }
  static const char* name() { return "reference"; }
};

}  // namespace orc
