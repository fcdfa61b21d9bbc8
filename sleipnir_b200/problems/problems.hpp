// Problem builders for the benchmark / parity configurations, written against
// the slp:: DSL exactly as a user of the reference would write them:
//   cart-pole  benchmarks/scalability/cart_pole/sleipnir.cpp:16-129 (+ rk4.hpp)
//   flywheel   benchmarks/scalability/flywheel/sleipnir.cpp:12-43
//   small NLPs test/src/optimization/{linear,quadratic,nonlinear}_problem_test.cpp,
//              test/src/optimization/solver/exit_status_test.cpp
#pragma once

#include <chrono>
#include <cmath>
#include <memory>
#include <numbers>
#include <stdexcept>
#include <string>

#include <sleipnir/autodiff/variable.hpp>
#include <sleipnir/autodiff/variable_matrix.hpp>
#include <sleipnir/optimization/problem.hpp>

namespace slpb_problems {

/// 4th order Runge-Kutta integration of dx/dt = f(x, u) over dt.
template <typename F, typename T, typename U>
T rk4(F&& f, T x, U u, std::chrono::duration<double> dt) {
  const auto h = dt.count();
  T k1 = f(x, u);
  T k2 = f(x + h * 0.5 * k1, u);
  T k3 = f(x + h * 0.5 * k2, u);
  T k4 = f(x + h * k3, u);
  return x + h / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
}

inline slp::VariableMatrix<double> cart_pole_dynamics(
    const slp::VariableMatrix<double>& x,
    const slp::VariableMatrix<double>& u) {
  // q = [x, θ]ᵀ, M(q)q̈ = τ_g(q) − C(q, q̇)q̇ + Bu
  constexpr double m_c = 5.0;  // Cart mass (kg)
  constexpr double m_p = 0.5;  // Pole mass (kg)
  constexpr double l = 0.5;    // Pole length (m)
  constexpr double g = 9.806;  // Acceleration due to gravity (m/s²)

  auto q = x.segment(0, 2);
  auto qdot = x.segment(2, 2);
  auto theta = q[1];
  auto thetadot = qdot[1];

  slp::VariableMatrix<double> M{{m_c + m_p, m_p * l * cos(theta)},
                                {m_p * l * cos(theta), m_p * std::pow(l, 2)}};
  slp::VariableMatrix<double> C{{0, -m_p * l * thetadot * sin(theta)}, {0, 0}};
  slp::VariableMatrix<double> tau_g{{0}, {-m_p * g * l * sin(theta)}};
  slp::Matrix<double> B{{1}, {0}};

  slp::VariableMatrix<double> qddot(4, 1);
  qddot.segment(0, 2) = qdot;
  qddot.segment(2, 2) = solve(M, tau_g - C * qdot + B * u);
  return qddot;
}

inline std::unique_ptr<slp::Problem<double>> cart_pole(int N, double T = 5.0) {
  const std::chrono::duration<double> dt{T / N};
  constexpr double u_max = 20.0;  // N
  constexpr double d_max = 2.0;   // m
  const slp::Matrix<double> x_initial{{0.0}, {0.0}, {0.0}, {0.0}};
  const slp::Matrix<double> x_final{{1.0}, {std::numbers::pi}, {0.0}, {0.0}};

  auto problem = std::make_unique<slp::Problem<double>>();

  // x = [q, q̇]ᵀ = [x, θ, ẋ, θ̇]ᵀ
  auto X = problem->decision_variable(4, N + 1);
  for (int k = 0; k < N + 1; ++k) {
    X[0, k].set_value(std::lerp(x_initial(0, 0), x_final(0, 0),
                                static_cast<double>(k) / N));
    X[1, k].set_value(std::lerp(x_initial(1, 0), x_final(1, 0),
                                static_cast<double>(k) / N));
  }
  // u = f_x
  auto U = problem->decision_variable(1, N);

  problem->subject_to(X.col(0) == x_initial);
  problem->subject_to(X.col(N) == x_final);
  problem->subject_to(slp::bounds(0.0, X.row(0), d_max));
  problem->subject_to(slp::bounds(-u_max, U, u_max));
  for (int k = 0; k < N; ++k) {
    problem->subject_to(
        X.col(k + 1) ==
        rk4<decltype(cart_pole_dynamics), slp::VariableMatrix<double>,
            slp::VariableMatrix<double>>(cart_pole_dynamics, X.col(k),
                                         U.col(k), dt));
  }
  slp::Variable<double> J = 0.0;
  for (int k = 0; k < N; ++k) {
    J += U.col(k).T() * U.col(k);
  }
  problem->minimize(J);
  return problem;
}

inline std::unique_ptr<slp::Problem<double>> flywheel(int N, double T = 5.0) {
  const double dt = T / N;
  slp::Matrix<double> A{{std::exp(-dt)}};
  slp::Matrix<double> B{{1.0 - std::exp(-dt)}};

  auto problem = std::make_unique<slp::Problem<double>>();
  auto X = problem->decision_variable(1, N + 1);
  auto U = problem->decision_variable(1, N);
  for (int k = 0; k < N; ++k) {
    problem->subject_to(X.col(k + 1) == A * X.col(k) + B * U.col(k));
  }
  problem->subject_to(X.col(0) == 0.0);
  problem->subject_to(slp::bounds(-12, U, 12));

  slp::Matrix<double> r{{10.0}};
  slp::Variable<double> J = 0.0;
  for (int k = 0; k < N + 1; ++k) {
    J += ((r - X.col(k)).T() * (r - X.col(k)));
  }
  problem->minimize(J);
  return problem;
}

inline std::unique_ptr<slp::Problem<double>> small_problem(
    const std::string& name, double p0, double p1) {
  using T = double;
  auto P = std::make_unique<slp::Problem<T>>();
  auto& problem = *P;
  if (name == "lp_maximize") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    x.set_value(T(1));
    y.set_value(T(1));
    problem.maximize(T(50) * x + T(40) * y);
    problem.subject_to(x + T(1.5) * y <= T(750));
    problem.subject_to(T(2) * x + T(3) * y <= T(1500));
    problem.subject_to(T(2) * x + y <= T(1000));
    problem.subject_to(x >= T(0));
    problem.subject_to(y >= T(0));
  } else if (name == "quartic") {
    auto x = problem.decision_variable();
    x.set_value(T(20));
    problem.minimize(pow(x, T(4)));
    problem.subject_to(x >= T(1));
  } else if (name == "rosenbrock_cubic_line") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    x.set_value(p0);
    y.set_value(p1);
    problem.minimize(100 * pow(y - pow(x, 2), 2) + pow(1 - x, 2));
    problem.subject_to(y >= pow(x - 1, 3) + 1);
    problem.subject_to(y <= -x + 2);
  } else if (name == "rosenbrock_disk") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    x.set_value(p0);
    y.set_value(p1);
    problem.minimize(pow(T(1) - x, T(2)) +
                     T(100) * pow(y - pow(x, T(2)), T(2)));
    problem.subject_to(pow(x, T(2)) + pow(y, T(2)) <= T(2));
  } else if (name == "conflicting_bounds") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    problem.minimize(hypot(x, y));
    problem.subject_to(hypot(x, y) <= T(1));
    problem.subject_to(slp::bounds(T(0.5), x, T(-0.5)));
  } else if (name == "wachter_biegler") {
    auto x = problem.decision_variable();
    auto s1 = problem.decision_variable();
    auto s2 = problem.decision_variable();
    x.set_value(T(-2));
    s1.set_value(T(3));
    s2.set_value(T(1));
    problem.minimize(x);
    problem.subject_to(pow(x, T(2)) - s1 - T(1) == T(0));
    problem.subject_to(x - s2 - T(0.5) == T(0));
    problem.subject_to(s1 >= T(0));
    problem.subject_to(s2 >= T(0));
  } else if (name == "qp_inequality_2d") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    x.set_value(T(5));
    y.set_value(T(5));
    problem.minimize(x * x + y * T(2) * y);
    problem.subject_to(y >= -x + T(5));
  } else if (name == "locally_infeasible_ineq") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    auto z = problem.decision_variable();
    problem.subject_to(x >= y + T(1));
    problem.subject_to(y >= z + T(1));
    problem.subject_to(z >= x + T(1));
  } else if (name == "nonfinite_ineq") {
    auto x = problem.decision_variable();
    problem.subject_to(T(1) / x > T(1));
  } else if (name == "nonfinite_ineq_jacobian") {
    auto x = problem.decision_variable();
    problem.subject_to(sqrt(x) > T(1));
  } else {
    throw std::invalid_argument("unknown problem: " + name);
  }
  return P;
}

inline std::unique_ptr<slp::Problem<double>> make_problem(
    const std::string& name, int N, double p0, double p1) {
  if (name == "cart_pole") return cart_pole(N, p0 > 0 ? p0 : 5.0);
  if (name == "flywheel") return flywheel(N, p0 > 0 ? p0 : 5.0);
  return small_problem(name, p0, p1);
}

}  // namespace slpb_problems
