// CPU-tier tests of the host DSL (sleipnir_b200/include/sleipnir): the
// reference's own DSL-level tests restated against this repo's headers —
// test/src/optimization/constraints_test.cpp (boolean value of a constraint,
// concatenation), decision_variable_test.cpp (init / assign, symmetric
// matrices), trivial checks of expression typing. No device call is made.
// Built and run by tests/test_dsl.py; prints "ok <n checks>" or the failures.
#include <array>
#include <cstdio>
#include <stdexcept>
#include <thread>
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

// jacobian_test.cpp:15-80 (y = x, y = 3x, products), hessian_test.cpp:364-404
// (Rosenbrock grid, bit-exact off-diagonals), gradient_test.cpp (trig): the
// standalone value() / get() of the autodiff classes on the host
void standalone_autodiff_values() {
  using slp::Gradient;
  using slp::Hessian;
  using slp::Jacobian;
  {
    VariableMatrix<T> x{3};
    for (int i = 0; i < 3; ++i) x[i].set_value(T(i + 1));
    auto y = T(3) * x;
    Jacobian<T> J{y, x};
    const auto& Jv = J.value();
    auto Jg = J.get();
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {
        CHECK(Jv.coeff(r, c) == (r == c ? T(3) : T(0)));
        CHECK(Jg(r, c).value() == (r == c ? T(3) : T(0)));
      }
    }
  }
  {
    //     [x₁x₂]           [x₂ x₁ 0 ]
    // y = [x₂x₃]   dy/dx = [0  x₃ x₂]
    //     [x₁x₃]           [x₃ 0  x₁]
    VariableMatrix<T> x{3};
    for (int i = 0; i < 3; ++i) x[i].set_value(T(i + 1));
    VariableMatrix<T> y{3};
    y[0] = x[0] * x[1];
    y[1] = x[1] * x[2];
    y[2] = x[0] * x[2];
    Jacobian<T> J{y, x};
    const T expect[3][3] = {{2, 1, 0}, {0, 3, 2}, {3, 0, 1}};
    const auto& Jv = J.value();
    auto Jg = J.get();
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {
        CHECK(Jv.coeff(r, c) == expect[r][c]);
        CHECK(Jg(r, c).value() == expect[r][c]);
      }
    }
    // re-evaluation follows the variables (non-linear rows are re-swept)
    x[0].set_value(T(5));
    CHECK(J.value().coeff(0, 1) == T(5));
    CHECK(J.value().coeff(2, 2) == T(5));
  }
  {
    // z = (1 − x)² + 100(y − x²)²
    VariableMatrix<T> input{2};
    auto& x = input[0];
    auto& y = input[1];
    Hessian<T> hessian{
        slp::pow(T(1) - x, T(2)) + T(100) * slp::pow(y - slp::pow(x, T(2)), T(2)),
        input};
    Hessian<T, slp::Lower> lower{
        slp::pow(T(1) - x, T(2)) + T(100) * slp::pow(y - slp::pow(x, T(2)), T(2)),
        input};
    for (T x0 = T(-2.5); x0 < T(2.5); x0 += T(0.1)) {
      for (T y0 = T(-2.5); y0 < T(2.5); y0 += T(0.1)) {
        x.set_value(x0);
        y.set_value(y0);
        const auto& H = hessian.value();
        CHECK(std::abs(H.coeff(0, 0) - (T(1200) * x0 * x0 - T(400) * y0 + T(2))) <= 1e-11);
        CHECK(H.coeff(0, 1) == T(-400) * x0);
        CHECK(H.coeff(1, 0) == T(-400) * x0);
        CHECK(H.coeff(1, 1) == T(200));
        const auto& L = lower.value();
        CHECK(L.coeff(1, 0) == T(-400) * x0);
        CHECK(L.coeff(0, 1) == T(0));  // upper triangle filtered
      }
    }
  }
  {
    Variable<T> a, b;
    a.set_value(T(0.7));
    b.set_value(T(1.3));
    VariableMatrix<T> wrt{2};
    wrt[0] = a;
    wrt[1] = b;
    Gradient<T> g{slp::sin(a) * b + a / b, wrt};
    const auto& gv = g.value();
    CHECK(gv.coeff(0) == std::cos(T(0.7)) * T(1.3) + T(1) / T(1.3));
    CHECK(std::abs(gv.coeff(1) - (std::sin(T(0.7)) - T(0.7) / (T(1.3) * T(1.3)))) <= 1e-15);
    auto gg = g.get();
    CHECK(std::abs(gg(0).value() - gv.coeff(0)) <= 1e-15);
    CHECK(std::abs(gg(1).value() - gv.coeff(1)) <= 1e-15);
  }
}

// the nodes a PoolScope's body created are reclaimed when it closes
// with no handle left into them (what Problem::solve does around itself), and
// kept when one survives; nested scopes hand survivors to the outer one.
void pool_scope_reclaims_regions() {
  auto& P = slp::detail::pool();
  Variable<T> x;
  x.set_value(T(2));
  const size_t base = P.nodes_in_use();
  {
    slp::detail::PoolScope scope{P};
    auto tmp = slp::sin(x) * x + x * x;
    CHECK(P.nodes_in_use() > base);
  }
  CHECK(P.nodes_in_use() == base);
  Variable<T> kept;
  {
    slp::detail::PoolScope scope{P};
    kept = slp::cos(x) * x;
    {
      slp::detail::PoolScope inner{P};
      auto tmp = kept * kept;
    }
  }
  CHECK(P.nodes_in_use() > base);         // `kept` still refers into the region
  CHECK(kept.value() == std::cos(T(2)) * T(2));
  // repeated scoped work on a long-lived graph does not grow the pool
  const size_t steady = P.nodes_in_use();
  for (int rep = 0; rep < 5; ++rep) {
    slp::detail::PoolScope scope{P};
    slp::Hessian<T> H{kept * kept + slp::exp(x), VariableMatrix<T>{x}};
    (void)H.value();
  }
  CHECK(P.nodes_in_use() == steady);
}

// expression handles are ids into the creating thread's pool, so a
// Problem refuses to be solved from another thread instead of reading
// another pool's nodes.
void problem_is_bound_to_its_thread() {
  slp::Problem<T> problem;
  auto x = problem.decision_variable();
  problem.minimize(x * x);
  bool threw = false;
  std::thread th([&] {
    try {
      problem.solve();
    } catch (const std::logic_error&) {
      threw = true;
    } catch (...) {
    }
  });
  th.join();
  CHECK(threw);
}

int main() {
  equality_constraint_boolean_comparison();
  inequality_constraint_boolean_comparisons();
  constraint_concatenation();
  decision_variables();
  expression_types();
  pool_is_returned();
  multistart_picks_the_best();
  standalone_autodiff_values();
  pool_scope_reclaims_regions();
  problem_is_bound_to_its_thread();
  if (g_failures == 0) {
    std::printf("ok %d checks\n", g_checks);
    return 0;
  }
  std::printf("%d of %d checks failed\n", g_failures, g_checks);
  return 1;
}
