// CPU-tier tests of the host DSL (sleipnir_b200/include/sleipnir): the
// reference's own DSL-level tests restated against this repo's headers —
// test/src/optimization/constraints_test.cpp (boolean value of a constraint,
// concatenation), decision_variable_test.cpp (init / assign, symmetric
// matrices), trivial checks of expression typing. No device call is made.
// Built and run by tests/test_dsl.py; prints "ok <n checks>" or the failures.
#include <array>
#include <cstdio>
#include <tuple>

#include "sleipnir/optimization/multistart.hpp"
#include "sleipnir/optimization/ocp.hpp"
#include "sleipnir/optimization/problem.hpp"

namespace {
int g_checks = 0, g_failures = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    ++g_checks;                                                             \
    if (!(cond)) {                                                          \
      ++g_failures;                                                         \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);         \
    }                                                                       \
  } while (0)

using T = double;
using slp::Variable;
using slp::VariableMatrix;
using MatrixXT = slp::Matrix<T>;

// constraints_test.cpp:15-99
void equality_constraint_boolean_comparison() {
  constexpr std::array args{std::tuple{T(1), T(1)}, std::tuple{T(1), T(2)},
                            std::tuple{T(2), T(1)}};
  for (const auto& [lhs, rhs] : args) {
    const bool expect = lhs == rhs;
    CHECK(bool{T{lhs} == Variable<T>{rhs}} == expect);
    CHECK(bool{T{lhs} == VariableMatrix<T>{{rhs}}} == expect);
    CHECK(bool{Variable<T>{lhs} == T{rhs}} == expect);
    CHECK(bool{Variable<T>{lhs} == Variable<T>{rhs}} == expect);
    CHECK(bool{Variable<T>{lhs} == VariableMatrix<T>{{rhs}}} == expect);
    CHECK(bool{VariableMatrix<T>{{lhs}} == T{rhs}} == expect);
    CHECK(bool{VariableMatrix<T>{{lhs}} == Variable<T>{rhs}} == expect);
    CHECK(bool{VariableMatrix<T>{{lhs}} == VariableMatrix<T>{{rhs}}} == expect);
    CHECK(bool{MatrixXT{{lhs}} == Variable<T>{rhs}} == expect);
    CHECK(bool{MatrixXT{{lhs}} == VariableMatrix<T>{{rhs}}} == expect);
    CHECK(bool{Variable<T>{lhs} == MatrixXT{{rhs}}} == expect);
    CHECK(bool{VariableMatrix<T>{{lhs}} == MatrixXT{{rhs}}} == expect);
  }
}

// constraints_test.cpp:101-234: `<` means ≤ and `>` means ≥
void inequality_constraint_boolean_comparisons() {
  constexpr std::array args{std::tuple{T(1), T(1)}, std::tuple{T(1), T(2)},
                            std::tuple{T(2), T(1)}};
  for (const auto& [lhs, rhs] : args) {
    const bool le = lhs <= rhs, ge = lhs >= rhs;
    CHECK(bool{T{lhs} < Variable<T>{rhs}} == le);
    CHECK(bool{T{lhs} <= Variable<T>{rhs}} == le);
    CHECK(bool{T{lhs} > Variable<T>{rhs}} == ge);
    CHECK(bool{T{lhs} >= Variable<T>{rhs}} == ge);
    CHECK(bool{T{lhs} < VariableMatrix<T>{{rhs}}} == le);
    CHECK(bool{T{lhs} >= VariableMatrix<T>{{rhs}}} == ge);
    CHECK(bool{Variable<T>{lhs} < T{rhs}} == le);
    CHECK(bool{Variable<T>{lhs} <= T{rhs}} == le);
    CHECK(bool{Variable<T>{lhs} > T{rhs}} == ge);
    CHECK(bool{Variable<T>{lhs} >= T{rhs}} == ge);
    CHECK(bool{Variable<T>{lhs} < Variable<T>{rhs}} == le);
    CHECK(bool{Variable<T>{lhs} >= Variable<T>{rhs}} == ge);
    CHECK(bool{Variable<T>{lhs} <= VariableMatrix<T>{{rhs}}} == le);
    CHECK(bool{Variable<T>{lhs} > VariableMatrix<T>{{rhs}}} == ge);
    CHECK(bool{VariableMatrix<T>{{lhs}} < T{rhs}} == le);
    CHECK(bool{VariableMatrix<T>{{lhs}} >= T{rhs}} == ge);
    CHECK(bool{VariableMatrix<T>{{lhs}} <= Variable<T>{rhs}} == le);
    CHECK(bool{VariableMatrix<T>{{lhs}} > Variable<T>{rhs}} == ge);
    CHECK(bool{VariableMatrix<T>{{lhs}} < VariableMatrix<T>{{rhs}}} == le);
    CHECK(bool{VariableMatrix<T>{{lhs}} >= VariableMatrix<T>{{rhs}}} == ge);
    CHECK(bool{MatrixXT{{lhs}} <= Variable<T>{rhs}} == le);
    CHECK(bool{MatrixXT{{lhs}} > VariableMatrix<T>{{rhs}}} == ge);
    CHECK(bool{Variable<T>{lhs} < MatrixXT{{rhs}}} == le);
    CHECK(bool{VariableMatrix<T>{{lhs}} >= MatrixXT{{rhs}}} == ge);
  }
}

// constraints_test.cpp:236-278
void constraint_concatenation() {
  using slp::EqualityConstraints;
  using slp::InequalityConstraints;
  EqualityConstraints eq1 = Variable<T>{1} == Variable<T>{1};
  EqualityConstraints eq2 = Variable<T>{1} == Variable<T>{2};
  EqualityConstraints eqs{eq1, eq2};
  CHECK(eq1.constraints.size() == 1);
  CHECK(eq2.constraints.size() == 1);
  CHECK(eqs.constraints.size() == 2);
  CHECK(eqs.constraints[0].value() == eq1.constraints[0].value());
  CHECK(eqs.constraints[1].value() == eq2.constraints[0].value());
  CHECK(bool{eq1});
  CHECK(!bool{eq2});
  CHECK(!bool{eqs});

  InequalityConstraints ineq1 = Variable<T>{2} < Variable<T>{1};
  InequalityConstraints ineq2 = Variable<T>{1} < Variable<T>{2};
  InequalityConstraints ineqs{ineq1, ineq2};
  CHECK(ineq1.constraints.size() == 1);
  CHECK(ineq2.constraints.size() == 1);
  CHECK(ineqs.constraints.size() == 2);
  CHECK(ineqs.constraints[0].value() == ineq1.constraints[0].value());
  CHECK(ineqs.constraints[1].value() == ineq2.constraints[0].value());
  CHECK(!bool{ineq1});
  CHECK(bool{ineq2});
  CHECK(!bool{ineqs});
}

// decision_variable_test.cpp:11-182
void decision_variables() {
  slp::Problem<T> problem;
  auto x = problem.decision_variable();
  CHECK(x.value() == T(0));
  x.set_value(T(1));
  CHECK(x.value() == T(1));
  x.set_value(T(2));
  CHECK(x.value() == T(2));

  auto y = problem.decision_variable(2);
  CHECK(y.value(0) == T(0) && y.value(1) == T(0));
  y[0].set_value(T(1));
  y[1].set_value(T(2));
  CHECK(y.value(0) == T(1) && y.value(1) == T(2));
  y.set_value(MatrixXT{{3.0}, {4.0}});
  CHECK(y.value(0) == T(3) && y.value(1) == T(4));

  auto z = problem.decision_variable(3, 2);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 2; ++c) CHECK(z.value(r, c) == T(0));
  }
  z.set_value(MatrixXT{{1.0, 2.0}, {3.0, 4.0}, {5.0, 6.0}});
  CHECK(z.value(0, 0) == T(1) && z.value(0, 1) == T(2));
  CHECK(z.value(1, 0) == T(3) && z.value(1, 1) == T(4));
  CHECK(z.value(2, 0) == T(5) && z.value(2, 1) == T(6));
  z[1, 1].set_value(T(9));
  CHECK(z.value(1, 1) == T(9));

  auto A = problem.symmetric_decision_variable(2);
  CHECK(A.value(0, 0) == T(0) && A.value(0, 1) == T(0));
  CHECK(A.value(1, 0) == T(0) && A.value(1, 1) == T(0));
  A[0, 0].set_value(T(1));
  A[1, 0].set_value(T(2));
  A[1, 1].set_value(T(3));
  CHECK(A.value(0, 0) == T(1) && A.value(0, 1) == T(2));
  CHECK(A.value(1, 0) == T(2) && A.value(1, 1) == T(3));
  CHECK(problem.decision_variables().size() == 1 + 2 + 6 + 3);
}

// the CHECKs on expression types that the reference's problem tests make
// before solving (linear_problem_test.cpp:27-29, nonlinear_problem_test.cpp:
// 31-33, 103-105, cart_pole_problem_test.cpp:83-85)
void expression_types() {
  using slp::ExpressionType;
  slp::Problem<T> problem;
  auto x = problem.decision_variable();
  auto y = problem.decision_variable();
  CHECK(problem.cost_function_type() == ExpressionType::NONE);
  problem.maximize(T(50) * x + T(40) * y);
  CHECK(problem.cost_function_type() == ExpressionType::LINEAR);
  problem.minimize(x * x + y);
  CHECK(problem.cost_function_type() == ExpressionType::QUADRATIC);
  problem.minimize(pow(x, T(4)));
  CHECK(problem.cost_function_type() == ExpressionType::NONLINEAR);
  problem.minimize(x * y * y);
  CHECK(problem.cost_function_type() == ExpressionType::NONLINEAR);
  CHECK(problem.equality_constraint_type() == ExpressionType::NONE);
  problem.subject_to(x + T(3) * y == T(36));
  CHECK(problem.equality_constraint_type() == ExpressionType::LINEAR);
  problem.subject_to(x * y == T(1));
  CHECK(problem.equality_constraint_type() == ExpressionType::QUADRATIC);
  CHECK(problem.inequality_constraint_type() == ExpressionType::NONE);
  problem.subject_to(x >= T(0));
  CHECK(problem.inequality_constraint_type() == ExpressionType::LINEAR);
  problem.subject_to(sin(x) <= T(0.5));
  CHECK(problem.inequality_constraint_type() == ExpressionType::NONLINEAR);
  // pow(x, 2) of a LINEAR x is x·x and QUADRATIC (expression.hpp:1741-1747)
  slp::Problem<T> q;
  auto u = q.decision_variable();
  q.minimize(pow(u, T(2)));
  CHECK(q.cost_function_type() == ExpressionType::QUADRATIC);
  // multiplication by a constant keeps the type; by zero folds to a constant
  q.minimize(T(3) * (u * u));
  CHECK(q.cost_function_type() == ExpressionType::QUADRATIC);
  q.minimize(T(0) * u + T(2));
  CHECK(q.cost_function_type() == ExpressionType::CONSTANT);
}

// every handle returned: the analogue of the reference tests'
// `global_pool_resource().blocks_in_use() == 0` scope guard
void pool_is_returned() {
  {
    slp::Problem<T> problem;
    auto X = problem.decision_variable(4, 11);
    auto U = problem.decision_variable(1, 10);
    slp::Variable<T> J = T(0);
    for (int k = 0; k < 10; ++k) J += U.col(k).T() * U.col(k);
    problem.minimize(J);
    problem.subject_to(X.col(0) == T(0));
    CHECK(slp::detail::pool().handles_in_use() > 0);
  }
  CHECK(slp::detail::pool().handles_in_use() == 0);
  CHECK(slp::detail::pool().nodes_in_use() == 0);
}

// multistart.hpp:44-73 with a solve function that needs no device: successful
// results beat unsuccessful ones, then the lower cost wins
void multistart_picks_the_best() {
  struct Guess {
    double x;
  };
  using Result = slp::MultistartResult<double, Guess>;
  const std::function<Result(const Guess&)> solve = [](const Guess& g) {
    return Result{g.x < 0 ? slp::ExitStatus::LOCALLY_INFEASIBLE
                          : slp::ExitStatus::SUCCESS,
                  (g.x - 3.0) * (g.x - 3.0), g};
  };
  const std::vector<Guess> guesses{{-3.0}, {10.0}, {2.5}, {6.0}};
  std::vector<Result> all;
  const Result best = slp::multistart<double, Guess>(
      solve, std::span<const Guess>{guesses}, 2, &all);
  CHECK(best.status == slp::ExitStatus::SUCCESS);
  CHECK(best.variables.x == 2.5);
  CHECK(all.size() == 4 && all[0].variables.x == -3.0 && all[3].variables.x == 6.0);
  // an unsuccessful start with the lowest cost still loses
  const std::vector<Guess> two{{-0.1}, {9.0}};
  CHECK((slp::multistart<double, Guess>(solve, std::span<const Guess>{two}))
            .variables.x == 9.0);
}

}  // namespace

int main() {
  equality_constraint_boolean_comparison();
  inequality_constraint_boolean_comparisons();
  constraint_concatenation();
  decision_variables();
  expression_types();
  pool_is_returned();
  multistart_picks_the_best();
  if (g_failures == 0) {
    std::printf("ok %d checks\n", g_checks);
    return 0;
  }
  std::printf("%d of %d checks failed\n", g_failures, g_checks);
  return 1;
}
