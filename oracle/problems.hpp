// ORACLE — TEST INFRASTRUCTURE ONLY (see expr.hpp header).
//
// Problem builders restated operator-for-operator from the reference:
//   cart-pole   benchmarks/scalability/cart_pole/sleipnir.cpp:16-129,
//               benchmarks/rk4.hpp:14-23
//   flywheel    benchmarks/scalability/flywheel/sleipnir.cpp:12-43
//   small NLPs  test/src/optimization/{linear,quadratic,nonlinear}_problem_test.cpp
// Operand order is preserved everywhere: it decides grad_l vs grad_r and hence
// floating-point rounding.
#pragma once

#include <cmath>
#include <memory>
#include <numbers>
#include <stdexcept>
#include <string>
#include <vector>

#include "problem.hpp"

namespace orc {

template <class B>
Mat<B> cart_pole_dynamics(const Mat<B>& x, const Mat<B>& u) {
  using M = Mat<B>;
  using V = Var<B>;
  constexpr double m_c = 5.0;
  constexpr double m_p = 0.5;
  constexpr double l = 0.5;
  constexpr double g = 9.806;

  M q = x.segment(0, 2);
  M qdot = x.segment(2, 2);
  V theta = q[1];
  V thetadot = qdot[1];

  M Mm{{V{m_c + m_p}, V{m_p * l} * cos(theta)},
       {V{m_p * l} * cos(theta), V{m_p * std::pow(l, 2)}}};
  M C{{V{0}, V{-m_p * l} * thetadot * sin(theta)}, {V{0}, V{0}}};
  M tau_g{{V{0}}, {V{-m_p * g * l} * sin(theta)}};
  M Bm = M::constants(2, 1, {1.0, 0.0});

  M qddot{4, 1};
  qddot.set_block(0, 0, qdot);
  qddot.set_block(2, 0, solve(Mm, tau_g - C * qdot + Bm * u));
  return qddot;
}

template <class B, class F>
Mat<B> rk4(F&& f, const Mat<B>& x, const Mat<B>& u, double h) {
  Mat<B> k1 = f(x, u);
  Mat<B> k2 = f(x + h * 0.5 * k1, u);
  Mat<B> k3 = f(x + h * 0.5 * k2, u);
  Mat<B> k4 = f(x + h * k3, u);
  return x + h / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
}

template <class B>
std::unique_ptr<Problem<B>> cart_pole_problem(int N, double T = 5.0) {
  using M = Mat<B>;
  using V = Var<B>;
  const double dt = T / N;
  constexpr double u_max = 20.0;
  constexpr double d_max = 2.0;
  const std::vector<double> x_initial{0.0, 0.0, 0.0, 0.0};
  const std::vector<double> x_final{1.0, std::numbers::pi, 0.0, 0.0};

  auto problem = std::make_unique<Problem<B>>();
  M X = problem->decision_variable(4, N + 1);
  for (int k = 0; k < N + 1; ++k) {
    X(0, k).set_value(
        std::lerp(x_initial[0], x_final[0], static_cast<double>(k) / N));
    X(1, k).set_value(
        std::lerp(x_initial[1], x_final[1], static_cast<double>(k) / N));
  }
  M U = problem->decision_variable(1, N);

  problem->subject_to_eq(eq(X.col(0), M::constants(4, 1, x_initial)));
  problem->subject_to_eq(eq(X.col(N), M::constants(4, 1, x_final)));
  problem->subject_to_ineq(bounds(V{0.0}, X.row(0), V{d_max}));
  problem->subject_to_ineq(bounds(V{-u_max}, U, V{u_max}));
  for (int k = 0; k < N; ++k) {
    problem->subject_to_eq(
        eq(X.col(k + 1),
           rk4<B>(cart_pole_dynamics<B>, X.col(k), U.col(k), dt)));
  }
  V J{0.0};
  for (int k = 0; k < N; ++k) {
    M uu = U.col(k).T() * U.col(k);
    J += uu(0, 0);
  }
  problem->minimize(J);
  return problem;
}

template <class B>
std::unique_ptr<Problem<B>> flywheel_problem(int N, double T = 5.0) {
  using M = Mat<B>;
  using V = Var<B>;
  const double dt = T / N;
  M A = M::constants(1, 1, {std::exp(-dt)});
  M Bm = M::constants(1, 1, {1.0 - std::exp(-dt)});

  auto problem = std::make_unique<Problem<B>>();
  M X = problem->decision_variable(1, N + 1);
  M U = problem->decision_variable(1, N);
  for (int k = 0; k < N; ++k) {
    problem->subject_to_eq(eq(X.col(k + 1), A * X.col(k) + Bm * U.col(k)));
  }
  problem->subject_to_eq(eq(X.col(0), V{0.0}));
  problem->subject_to_ineq(bounds(V{-12}, U, V{12}));
  M r = M::constants(1, 1, {10.0});
  V J{0.0};
  for (int k = 0; k < N + 1; ++k) {
    M e = (r - X.col(k)).T() * (r - X.col(k));
    J += e(0, 0);
  }
  problem->minimize(J);
  return problem;
}

/// Small known-answer problems from the reference's tests. `p0`, `p1` are the
/// initial guess where the test sweeps one.
template <class B>
std::unique_ptr<Problem<B>> small_problem(const std::string& name, double p0,
                                          double p1) {
  using V = Var<B>;
  using M = Mat<B>;
  auto P = std::make_unique<Problem<B>>();
  auto ge1 = [](const V& l, const V& r) { return std::vector<V>{l - r}; };
  auto le1 = [](const V& l, const V& r) { return std::vector<V>{r - l}; };
  if (name == "lp_maximize") {  // linear_problem_test.cpp:14-40
    V x = P->decision_variable();
    V y = P->decision_variable();
    x.set_value(1);
    y.set_value(1);
    P->maximize(V{50} * x + V{40} * y);
    P->subject_to_ineq(le1(x + V{1.5} * y, V{750}));
    P->subject_to_ineq(le1(V{2} * x + V{3} * y, V{1500}));
    P->subject_to_ineq(le1(V{2} * x + y, V{1000}));
    P->subject_to_ineq(ge1(x, V{0}));
    P->subject_to_ineq(ge1(y, V{0}));
  } else if (name == "quartic") {  // nonlinear_problem_test.cpp:19-37
    V x = P->decision_variable();
    x.set_value(20);
    P->minimize(pow(x, 4.0));
    P->subject_to_ineq(ge1(x, V{1}));
  } else if (name == "rosenbrock_cubic_line") {  // :39-82
    V x = P->decision_variable();
    V y = P->decision_variable();
    x.set_value(p0);
    y.set_value(p1);
    P->minimize(V{100} * pow(y - pow(x, 2.0), 2.0) + pow(V{1} - x, 2.0));
    P->subject_to_ineq(ge1(y, pow(x - V{1}, 3.0) + V{1}));
    P->subject_to_ineq(le1(y, -x + V{2}));
  } else if (name == "rosenbrock_disk") {  // :84-118
    V x = P->decision_variable();
    V y = P->decision_variable();
    x.set_value(p0);
    y.set_value(p1);
    P->minimize(pow(V{1} - x, 2.0) + V{100} * pow(y - pow(x, 2.0), 2.0));
    P->subject_to_ineq(le1(pow(x, 2.0) + pow(y, 2.0), V{2}));
  } else if (name == "conflicting_bounds") {  // :145-165
    V x = P->decision_variable();
    V y = P->decision_variable();
    P->minimize(hypot(x, y));
    P->subject_to_ineq(le1(hypot(x, y), V{1}));
    P->subject_to_ineq(bounds(V{0.5}, M{x}, V{-0.5}));
  } else if (name == "wachter_biegler") {  // :167-201
    V x = P->decision_variable();
    V s1 = P->decision_variable();
    V s2 = P->decision_variable();
    x.set_value(-2);
    s1.set_value(3);
    s2.set_value(1);
    P->minimize(x);
    P->subject_to_eq({pow(x, 2.0) - s1 - V{1} - V{0}});
    P->subject_to_eq({x - s2 - V{0.5} - V{0}});
    P->subject_to_ineq(ge1(s1, V{0}));
    P->subject_to_ineq(ge1(s2, V{0}));
  } else if (name == "qp_inequality_2d") {  // quadratic_problem_test.cpp:164-186
    V x = P->decision_variable();
    V y = P->decision_variable();
    x.set_value(5);
    y.set_value(5);
    P->minimize(x * x + y * V{2} * y);
    P->subject_to_ineq(ge1(y, -x + V{5}));
  } else if (name == "locally_infeasible_ineq") {  // exit_status_test.cpp:97-117
    V x = P->decision_variable();
    V y = P->decision_variable();
    V z = P->decision_variable();
    P->subject_to_ineq(ge1(x, y + V{1}));
    P->subject_to_ineq(ge1(y, z + V{1}));
    P->subject_to_ineq(ge1(z, x + V{1}));
  } else if (name == "nonfinite_ineq") {  // exit_status_test.cpp:160-166
    V x = P->decision_variable();
    P->subject_to_ineq(ge1(V{1} / x, V{1}));
  } else if (name == "nonfinite_ineq_jacobian") {  // :169-175
    V x = P->decision_variable();
    P->subject_to_ineq(ge1(sqrt(x), V{1}));
  } else {
    throw std::invalid_argument("unknown oracle problem: " + name);
  }
  return P;
}

template <class B>
std::unique_ptr<Problem<B>> make_problem(const std::string& name, int N,
                                         double p0, double p1) {
  if (name == "cart_pole") return cart_pole_problem<B>(N, p0 > 0 ? p0 : 5.0);
  if (name == "flywheel") return flywheel_problem<B>(N, p0 > 0 ? p0 : 5.0);
  return small_problem<B>(name, p0, p1);
}

}  // namespace orc
