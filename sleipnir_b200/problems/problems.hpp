// Problem builders for the benchmark / parity configurations, written against
// the slp:: DSL exactly as a user of the reference would write them:
//   cart-pole  benchmarks/scalability/cart_pole/sleipnir.cpp:16-129 (+ rk4.hpp)
//   flywheel   benchmarks/scalability/flywheel/sleipnir.cpp:12-43
//   g-fold     examples/g-fold/src/main.cpp:145-386 (dt rescaled to T_f/N)
//   small NLPs test/src/optimization/{linear,quadratic,nonlinear}_problem_test.cpp,
//              test/src/optimization/solver/exit_status_test.cpp
#pragma once

#include <chrono>
#include <cmath>
#include <memory>
#include <numbers>
#include <stdexcept>
#include <string>
#include <vector>

#include <sleipnir/autodiff/variable.hpp>
#include <sleipnir/autodiff/variable_matrix.hpp>
#include <sleipnir/optimization/ocp.hpp>
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

inline std::unique_ptr<slp::Problem<double>> cart_pole(int N, double T = 5.0,
                                                       bool bounded = true) {
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
  if (bounded) {  // without the bounds solve() takes the SQP branch
    problem->subject_to(slp::bounds(0.0, X.row(0), d_max));
    problem->subject_to(slp::bounds(-u_max, U, u_max));
  }
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

inline std::unique_ptr<slp::Problem<double>> flywheel(int N, double T = 5.0,
                                                      bool bounded = true) {
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
  if (bounded) problem->subject_to(slp::bounds(-12, U, 12));

  slp::Matrix<double> r{{10.0}};
  slp::Variable<double> J = 0.0;
  for (int k = 0; k < N + 1; ++k) {
    J += ((r - X.col(k)).T() * (r - X.col(k)));
  }
  problem->minimize(J);
  return problem;
}


/// Zero-order-hold discretisation [A_d B_d; 0 I] = exp([A B; 0 0]·dt) by scaling
/// and squaring of a Taylor series (the reference uses Eigen's matrix
/// exponential, examples/g-fold/src/main.cpp:39-57; Eigen is not in the image).
inline void discretize_ab(const slp::Matrix<double>& A,
                          const slp::Matrix<double>& B, double dt,
                          slp::Matrix<double>& A_d, slp::Matrix<double>& B_d) {
  const int ns = A.rows(), ni = B.cols(), n = ns + ni;
  std::vector<double> M(size_t(n) * n, 0.0), E(size_t(n) * n, 0.0),
      T(size_t(n) * n, 0.0), tmp(size_t(n) * n, 0.0);
  for (int r = 0; r < ns; ++r) {
    for (int c = 0; c < ns; ++c) M[r * n + c] = A(r, c) * dt;
    for (int c = 0; c < ni; ++c) M[r * n + ns + c] = B(r, c) * dt;
  }
  double norm = 0.0;
  for (double v : M) norm = std::max(norm, std::abs(v));
  int squarings = 0;
  while (norm * n > 0.5) {
    norm *= 0.5;
    ++squarings;
  }
  const double scale = std::ldexp(1.0, -squarings);
  for (double& v : M) v *= scale;
  auto matmul = [&](const std::vector<double>& X, const std::vector<double>& Y,
                    std::vector<double>& Z) {
    for (int r = 0; r < n; ++r) {
      for (int c = 0; c < n; ++c) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += X[r * n + k] * Y[k * n + c];
        Z[r * n + c] = acc;
      }
    }
  };
  for (int i = 0; i < n; ++i) E[i * n + i] = T[i * n + i] = 1.0;
  for (int term = 1; term <= 20; ++term) {
    matmul(T, M, tmp);
    for (size_t i = 0; i < tmp.size(); ++i) T[i] = tmp[i] / term;
    for (size_t i = 0; i < E.size(); ++i) E[i] += T[i];
  }
  for (int q = 0; q < squarings; ++q) {
    matmul(E, E, tmp);
    E = tmp;
  }
  A_d = slp::Matrix<double>(ns, ns);
  B_d = slp::Matrix<double>(ns, ni);
  for (int r = 0; r < ns; ++r) {
    for (int c = 0; c < ns; ++c) A_d(r, c) = E[r * n + c];
    for (int c = 0; c < ni; ++c) B_d(r, c) = E[r * n + ns + c];
  }
}

/// Powered-descent guidance (G-FOLD), examples/g-fold/src/main.cpp:145-386 with
/// the time step rescaled to dt = T_f/N (the example hard-codes dt = 0.5 s,
/// which is only meaningful for N ≤ 250; SURVEY.md §7 hard part 8).
inline std::unique_ptr<slp::Problem<double>> gfold(int N, double T_f = 48.0) {
  using slp::Matrix;
  constexpr double m_wet = 2000.0;
  constexpr double T_max = 24000;
  constexpr double rho_1 = 0.2 * T_max;
  constexpr double rho_2 = 0.8 * T_max;
  constexpr double alpha = 5e-4;
  const Matrix<double> q_0{{2400.0}, {450.0}, {-330.0}};
  const Matrix<double> v_0{{-10.0}, {-40.0}, {10.0}};
  const Matrix<double> q_f{{0.0}, {0.0}, {0.0}};
  const Matrix<double> v_f{{0.0}, {0.0}, {0.0}};
  const Matrix<double> g{{-3.71}, {0.0}, {0.0}};
  constexpr double w1 = 2.53e-5, w2 = 0.0, w3 = 6.62e-5;
  const double theta = 90.0 * std::numbers::pi / 180.0;
  const double gamma_gs = 30.0 * std::numbers::pi / 180.0;
  constexpr double v_max = 90.0;
  const double dt = T_f / N;
  constexpr bool END_STRAIGHT = true;

  const double S[3][3] = {{0.0, -w3, w2}, {w3, 0.0, -w1}, {-w2, w1, 0.0}};
  Matrix<double> A(6, 6), B(6, 3);
  for (int i = 0; i < 3; ++i) {
    A(i, 3 + i) = 1.0;
    B(3 + i, i) = 1.0;
    for (int j = 0; j < 3; ++j) {
      double ss = 0.0;
      for (int k = 0; k < 3; ++k) ss += S[i][k] * S[k][j];
      A(3 + i, j) = -ss;
      A(3 + i, 3 + j) = -2 * S[i][j];
    }
  }
  Matrix<double> A_d, B_d;
  discretize_ab(A, B, dt, A_d, B_d);

  auto P = std::make_unique<slp::Problem<double>>();
  auto& problem = *P;
  auto X = problem.decision_variable(6, N + 1);
  auto Z = problem.decision_variable(1, N + 1);
  auto U = problem.decision_variable(3, N);
  auto sigma = problem.decision_variable(1, N);

  auto q = X.block(0, 0, 3, N + 1);
  auto v = X.block(3, 0, 3, N + 1);

  problem.subject_to(q.col(0) == q_0);
  problem.subject_to(v.col(0) == v_0);
  problem.subject_to(Z[0, 0] == std::log(m_wet));
  problem.subject_to(q.col(N) == q_f);
  problem.subject_to(v.col(N) == v_f);

  for (int k = 0; k < N + 1; ++k) {
    for (int i = 0; i < 3; ++i) {
      q[i, k].set_value(std::lerp(q_0(i, 0), q_f(i, 0), static_cast<double>(k) / N));
      v[i, k].set_value(std::lerp(v_0(i, 0), v_f(i, 0), static_cast<double>(k) / N));
    }
  }

  for (int k = 0; k < N + 1; ++k) {
    const double t = k * dt;
    auto x_k = X.col(k);
    auto q_k = X.block(0, k, 3, 1);
    auto v_k = X.block(3, k, 3, 1);
    auto z_k = Z.col(k);

    problem.subject_to(v_k.T() * v_k <= v_max * v_max);

    const double z_min = std::log(m_wet - alpha * rho_2 * t);
    const double z_max = std::log(m_wet - alpha * rho_1 * t);
    const double z_estimate = (z_min + z_max) / 2;
    z_k.set_value(z_estimate);

    if (k < N) {
      auto x_k1 = X.col(k + 1);
      auto z_k1 = Z.col(k + 1);
      auto u_k = U.col(k);
      auto sigma_k = sigma.col(k);

      const double u_min = rho_1 / std::exp(z_estimate);
      const double u_max = rho_2 / std::exp(z_estimate);
      u_k.set_value(Matrix<double>{{(u_min + u_max) / 2}, {0.0}, {0.0}});

      problem.subject_to(
          slp::pow(q_k[0] - q_f(0, 0), 2) >=
          std::tan(gamma_gs) * std::tan(gamma_gs) *
              (slp::pow(q_k[1] - q_f(1, 0), 2) + slp::pow(q_k[2] - q_f(2, 0), 2)));

      problem.subject_to(sigma_k >= 0);

      if (k == N - 1 && END_STRAIGHT) {
        problem.subject_to(u_k[0, 0] == sigma_k);
        problem.subject_to(u_k[1, 0] == 0);
        problem.subject_to(u_k[2, 0] == 0);
      } else {
        problem.subject_to(u_k.T() * u_k <= sigma_k * sigma_k);
        problem.subject_to(u_k[0] >= std::cos(theta) * sigma_k);
      }

      const double z_0 = std::log(m_wet - alpha * rho_2 * t);
      const double mu_1 = rho_1 * std::exp(-z_0);
      const double mu_2 = rho_2 * std::exp(-z_0);
      auto sigma_min =
          mu_1 * (1 - (z_k[0] - z_0) + 0.5 * slp::pow(z_k[0] - z_0, 2));
      auto sigma_max = mu_2 * (1 - (z_k[0] - z_0));
      problem.subject_to(slp::bounds(sigma_min, sigma_k, sigma_max));
      sigma_k.set_value((sigma_min.value() + sigma_max.value()) / 2);

      problem.subject_to(x_k1 == A_d * x_k + B_d * (g + u_k));
      problem.subject_to(z_k1 == z_k - alpha * dt * sigma_k);
    }
  }

  slp::Variable<double> J{0.0};
  for (int k = 0; k < N; ++k) J = J + sigma[0, k];
  problem.minimize(J);
  return P;
}

/// Σ 100(xᵢ₊₁ − xᵢ²)² + (1 − xᵢ)² from the classic (−1.2, 1, −1.2, …) start: an
/// unconstrained problem of any size for the Newton branch.
inline std::unique_ptr<slp::Problem<double>> chained_rosenbrock(int N) {
  auto problem = std::make_unique<slp::Problem<double>>();
  auto X = problem->decision_variable(N + 1, 1);
  for (int k = 0; k < N + 1; ++k) X[k, 0].set_value(k % 2 == 0 ? -1.2 : 1.0);
  slp::Variable<double> J = 0.0;
  for (int k = 0; k < N; ++k) {
    slp::Variable<double> x = X[k, 0];
    slp::Variable<double> y = X[k + 1, 0];
    J += 100.0 * slp::pow(y - slp::pow(x, 2.0), 2.0) + slp::pow(1.0 - x, 2.0);
  }
  problem->minimize(J);
  return problem;
}

/// double_integrator_problem_test.cpp:27-127 with dt = 3.5 s / N: minimum
/// position error under velocity and acceleration limits (a bang-coast-bang
/// profile).
inline std::unique_ptr<slp::Problem<double>> double_integrator(int N) {
  using T = double;
  const T t = T(3.5) / N;
  constexpr T r(2);
  auto P = std::make_unique<slp::Problem<T>>();
  auto& problem = *P;
  auto X = problem.decision_variable(2, N + 1);
  auto U = problem.decision_variable(1, N);
  for (int k = 0; k < N; ++k) {
    auto p_k1 = X[0, k + 1];
    auto v_k1 = X[1, k + 1];
    auto p_k = X[0, k];
    auto v_k = X[1, k];
    auto a_k = U[0, k];
    problem.subject_to(p_k1 == p_k + v_k * t + 0.5 * a_k * t * t);
    problem.subject_to(v_k1 == v_k + a_k * t);
  }
  problem.subject_to(X.col(0) == slp::Matrix<T>{{0.0}, {0.0}});
  problem.subject_to(X.col(N) == slp::Matrix<T>{{r}, {0.0}});
  problem.subject_to(slp::bounds(T(-1), X.row(1), T(1)));
  problem.subject_to(slp::bounds(T(-1), U, T(1)));
  slp::Variable<T> J = T(0);
  for (int k = 0; k < N + 1; ++k) J += pow(r - X[0, k], 2);
  problem.minimize(J);
  return P;
}

/// arm_on_elevator_problem_test.cpp:27-122 with dt = 4 s / N: two double
/// integrators coupled by a nonlinear end-effector height limit.
inline std::unique_ptr<slp::Problem<double>> arm_on_elevator(int N) {
  using T = double;
  constexpr T ELEVATOR_START_HEIGHT(1), ELEVATOR_END_HEIGHT(1.25);
  constexpr T ELEVATOR_MAX_VELOCITY(1), ELEVATOR_MAX_ACCELERATION(2);
  constexpr T ARM_LENGTH(1), ARM_START_ANGLE(0);
  constexpr T ARM_END_ANGLE(std::numbers::pi);
  constexpr T ARM_MAX_VELOCITY(2.0 * std::numbers::pi);
  constexpr T ARM_MAX_ACCELERATION(4.0 * std::numbers::pi);
  constexpr T END_EFFECTOR_MAX_HEIGHT(1.8);
  const T dt = T(4) / T(N);
  auto P = std::make_unique<slp::Problem<T>>();
  auto& problem = *P;
  auto elevator = problem.decision_variable(2, N + 1);
  auto elevator_accel = problem.decision_variable(1, N);
  auto arm = problem.decision_variable(2, N + 1);
  auto arm_accel = problem.decision_variable(1, N);
  for (int k = 0; k < N; ++k) {
    problem.subject_to(elevator[0, k + 1] ==
                       elevator[0, k] + elevator[1, k] * dt +
                           T(0.5) * elevator_accel[0, k] * dt * dt);
    problem.subject_to(elevator[1, k + 1] ==
                       elevator[1, k] + elevator_accel[0, k] * dt);
    problem.subject_to(arm[0, k + 1] ==
                       arm[0, k] + arm[1, k] * dt +
                           T(0.5) * arm_accel[0, k] * dt * dt);
    problem.subject_to(arm[1, k + 1] == arm[1, k] + arm_accel[0, k] * dt);
  }
  problem.subject_to(elevator.col(0) ==
                     slp::Matrix<T>{{ELEVATOR_START_HEIGHT}, {0.0}});
  problem.subject_to(elevator.col(N) ==
                     slp::Matrix<T>{{ELEVATOR_END_HEIGHT}, {0.0}});
  problem.subject_to(arm.col(0) == slp::Matrix<T>{{ARM_START_ANGLE}, {0.0}});
  problem.subject_to(arm.col(N) == slp::Matrix<T>{{ARM_END_ANGLE}, {0.0}});
  problem.subject_to(slp::bounds(-ELEVATOR_MAX_VELOCITY, elevator.row(1),
                                 ELEVATOR_MAX_VELOCITY));
  problem.subject_to(slp::bounds(-ELEVATOR_MAX_ACCELERATION, elevator_accel,
                                 ELEVATOR_MAX_ACCELERATION));
  problem.subject_to(
      slp::bounds(-ARM_MAX_VELOCITY, arm.row(1), ARM_MAX_VELOCITY));
  problem.subject_to(
      slp::bounds(-ARM_MAX_ACCELERATION, arm_accel, ARM_MAX_ACCELERATION));
  auto heights =
      elevator.row(0) +
      ARM_LENGTH * arm.row(0).cwise_transform(
                       [](const slp::Variable<T>& x) { return sin(x); });
  problem.subject_to(heights <= END_EFFECTOR_MAX_HEIGHT);
  slp::Variable<T> J = T(0);
  for (int k = 0; k < N + 1; ++k) {
    J += pow(ELEVATOR_END_HEIGHT - elevator[0, k], T(2)) +
         pow(ARM_END_ANGLE - arm[0, k], T(2));
  }
  problem.minimize(J);
  return P;
}

/// test/include/differential_drive_util.hpp:16-61.
inline slp::VariableMatrix<double> differential_drive_dynamics(
    const slp::VariableMatrix<double>& x, const slp::VariableMatrix<double>& u) {
  using T = double;
  constexpr T trackwidth = 0.699, Kv_linear = 3.02, Ka_linear = 0.642;
  constexpr T Kv_angular = 1.382, Ka_angular = 0.08495;
  constexpr T A1 = -(Kv_linear / Ka_linear + Kv_angular / Ka_angular) / T(2);
  constexpr T A2 = -(Kv_linear / Ka_linear - Kv_angular / Ka_angular) / T(2);
  constexpr T B1 = T(0.5) / Ka_linear + T(0.5) / Ka_angular;
  constexpr T B2 = T(0.5) / Ka_linear - T(0.5) / Ka_angular;
  const slp::Matrix<T> A{{A1, A2}, {A2, A1}};
  const slp::Matrix<T> B{{B1, B2}, {B2, B1}};
  slp::VariableMatrix<T> xdot{5};
  auto v = (x[3] + x[4]) / T(2);
  xdot[0] = v * cos(x[2]);
  xdot[1] = v * sin(x[2]);
  xdot[2] = (x[4] - x[3]) / trackwidth;
  xdot.segment(3, 2) = A * x.segment(3, 2) + B * u;
  return xdot;
}

/// differential_drive_problem_test.cpp:28-138 with dt = 5 s / N.
inline std::unique_ptr<slp::Problem<double>> differential_drive(int N) {
  using T = double;
  const std::chrono::duration<T> dt{T(5) / N};
  constexpr T u_max(12);
  const slp::Matrix<T> x_initial{{0.0}, {0.0}, {0.0}, {0.0}, {0.0}};
  const slp::Matrix<T> x_final{{1.0}, {1.0}, {0.0}, {0.0}, {0.0}};
  auto P = std::make_unique<slp::Problem<T>>();
  auto& problem = *P;
  auto X = problem.decision_variable(5, N + 1);
  for (int k = 0; k < N; ++k) {
    X[0, k].set_value(std::lerp(x_initial(0, 0), x_final(0, 0), T(k) / T(N)));
    X[1, k].set_value(std::lerp(x_initial(1, 0), x_final(1, 0), T(k) / T(N)));
  }
  auto U = problem.decision_variable(2, N);
  problem.subject_to(X.col(0) == x_initial);
  problem.subject_to(X.col(N) == x_final);
  problem.subject_to(slp::bounds(-u_max, U, u_max));
  for (int k = 0; k < N; ++k) {
    problem.subject_to(
        X.col(k + 1) ==
        rk4<decltype(differential_drive_dynamics), slp::VariableMatrix<T>,
            slp::VariableMatrix<T>>(differential_drive_dynamics, X.col(k),
                                    U.col(k), dt));
  }
  slp::Variable<T> J = T(0);
  for (int k = 0; k < N; ++k) {
    J += X.col(k).T() * X.col(k) + U.col(k).T() * U.col(k);
  }
  problem.minimize(J);
  return P;
}

// ---- the reference's OCP tests, written against slp::OCP ---------------------

/// flywheel_ocp_test.cpp:38-201 with dt = 5 s / N. method: 0 direct
/// transcription, 1 direct collocation, 2 single shooting.
inline std::unique_ptr<slp::Problem<double>> flywheel_ocp(int N, int method,
                                                          bool discrete) {
  using T = double;
  const std::chrono::duration<T> dt{T(5) / N};
  constexpr T A(-1), B(1);
  const T A_discrete = std::exp(A * dt.count());
  const T B_discrete = (T(1) - A_discrete) * B;
  slp::OCP<T>::Dynamics f;
  if (discrete) {
    f = [=](const slp::VariableMatrix<T>& x, const slp::VariableMatrix<T>& u) {
      return A_discrete * x + B_discrete * u;
    };
  } else {
    f = [=](const slp::VariableMatrix<T>& x, const slp::VariableMatrix<T>& u) {
      return A * x + B * u;
    };
  }
  auto problem = std::make_unique<slp::OCP<T>>(
      1, 1, dt, N, f,
      discrete ? slp::DynamicsType::DISCRETE : slp::DynamicsType::EXPLICIT_ODE,
      slp::TimestepMethod::FIXED,
      static_cast<slp::TranscriptionMethod>(method));
  problem->constrain_initial_state(T(0));
  problem->set_upper_input_bound(T(12));
  problem->set_lower_input_bound(T(-12));
  slp::Matrix<T> r_mat{1, N + 1};
  for (int k = 0; k < N + 1; ++k) r_mat(0, k) = 10.0;
  problem->minimize((r_mat - problem->X()) * (r_mat - problem->X()).T());
  return problem;
}

/// cart_pole_ocp_test.cpp:29-86 with dt = 5 s / N.
inline std::unique_ptr<slp::Problem<double>> cart_pole_ocp(int N) {
  using T = double;
  const std::chrono::duration<T> dt{T(5) / N};
  constexpr T u_max(20), d_max(2);
  const slp::Matrix<T> x_initial{{0.0}, {0.0}, {0.0}, {0.0}};
  const slp::Matrix<T> x_final{{1.0}, {std::numbers::pi}, {0.0}, {0.0}};
  auto problem = std::make_unique<slp::OCP<T>>(
      4, 1, dt, N, slp::OCP<T>::Dynamics{cart_pole_dynamics},
      slp::DynamicsType::EXPLICIT_ODE, slp::TimestepMethod::VARIABLE_SINGLE,
      slp::TranscriptionMethod::DIRECT_COLLOCATION);
  auto& X = problem->X();
  for (int k = 0; k < N + 1; ++k) {
    X[0, k].set_value(std::lerp(x_initial(0, 0), x_final(0, 0), T(k) / T(N)));
    X[1, k].set_value(std::lerp(x_initial(1, 0), x_final(1, 0), T(k) / T(N)));
  }
  problem->constrain_initial_state(x_initial);
  problem->constrain_final_state(x_final);
  auto* raw = problem.get();
  problem->for_each_step([&](const slp::VariableMatrix<T>& x,
                             const slp::VariableMatrix<T>&) {
    raw->subject_to(slp::bounds(T(0), x[0], d_max));
  });
  problem->set_lower_input_bound(-u_max);
  problem->set_upper_input_bound(u_max);
  auto& U = problem->U();
  slp::Variable<T> J = T(0);
  for (int k = 0; k < N; ++k) J += U.col(k).T() * U.col(k);
  problem->minimize(J);
  return problem;
}

/// differential_drive_ocp_test.cpp:25-66 (dynamics:
/// test/include/differential_drive_util.hpp:16-61).
inline std::unique_ptr<slp::Problem<double>> differential_drive_ocp(int N) {
  using T = double;
  constexpr T trackwidth = 0.699, Kv_linear = 3.02, Ka_linear = 0.642;
  constexpr T Kv_angular = 1.382, Ka_angular = 0.08495;
  constexpr T A1 = -(Kv_linear / Ka_linear + Kv_angular / Ka_angular) / T(2);
  constexpr T A2 = -(Kv_linear / Ka_linear - Kv_angular / Ka_angular) / T(2);
  constexpr T B1 = T(0.5) / Ka_linear + T(0.5) / Ka_angular;
  constexpr T B2 = T(0.5) / Ka_linear - T(0.5) / Ka_angular;
  constexpr std::chrono::duration<T> min_timestep{T(0.05)};
  auto dynamics = [=](const slp::VariableMatrix<T>& x,
                      const slp::VariableMatrix<T>& u) {
    const slp::Matrix<T> A{{A1, A2}, {A2, A1}};
    const slp::Matrix<T> B{{B1, B2}, {B2, B1}};
    slp::VariableMatrix<T> xdot{5};
    auto v = (x[3] + x[4]) / T(2);
    xdot[0] = v * cos(x[2]);
    xdot[1] = v * sin(x[2]);
    xdot[2] = (x[4] - x[3]) / trackwidth;
    xdot.segment(3, 2) = A * x.segment(3, 2) + B * u;
    return xdot;
  };
  auto problem = std::make_unique<slp::OCP<T>>(
      5, 2, min_timestep, N, slp::OCP<T>::Dynamics{dynamics},
      slp::DynamicsType::EXPLICIT_ODE, slp::TimestepMethod::VARIABLE_SINGLE,
      slp::TranscriptionMethod::DIRECT_TRANSCRIPTION);
  for (int i = 0; i < N + 1; ++i) {
    problem->X()[0, i].set_value(T(i) / T(N + 1));
    problem->X()[1, i].set_value(T(i) / T(N + 1));
  }
  problem->constrain_initial_state(
      slp::Matrix<T>{{0.0}, {0.0}, {0.0}, {0.0}, {0.0}});
  problem->constrain_final_state(
      slp::Matrix<T>{{1.0}, {1.0}, {0.0}, {0.0}, {0.0}});
  problem->set_lower_input_bound(slp::Matrix<T>{{-12.0}, {-12.0}});
  problem->set_upper_input_bound(slp::Matrix<T>{{12.0}, {12.0}});
  problem->set_min_timestep(min_timestep);
  problem->set_max_timestep(std::chrono::duration<T>{T(3)});
  slp::Matrix<T> ones{N + 1, 1};
  for (int i = 0; i < N + 1; ++i) ones(i, 0) = 1.0;
  problem->minimize(problem->dt() * ones);
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
  } else if (name == "mishra_bird") {  // multistart_test.cpp:17-55
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    x.set_value(T(p0));
    y.set_value(T(p1));
    slp::Variable<double> J =
        sin(y) * exp(pow(T(1) - cos(x), T(2))) +
        cos(x) * exp(pow(T(1) - sin(y), T(2))) + pow(x - y, T(2));
    problem.minimize(J);
    problem.subject_to(pow(x + T(5), T(2)) + pow(y + T(5), T(2)) < T(25));
  } else if (name == "empty") {  // trivial_problem_test.cpp:14-24
  } else if (name == "no_cost_unconstrained") {  // :26-66
    auto X = problem.decision_variable(2, 3);
    for (int i = 0; i < 6; ++i) X[i].set_value(p0);
  } else if (name == "all_ops") {
    // every opcode of the tape (include/slpb.h, slpb_op) in the cost, an
    // equality and an inequality constraint, so that golden vectors from the
    // reference's own expression core pin value, gradient and Hessian of each
    auto v = problem.decision_variable(6);
    const double guess[6] = {0.3, 0.5, 0.7, 1.1, 1.3, 0.9};
    for (int i = 0; i < 6; ++i) v[i].set_value(guess[i]);
    slp::Variable<T> x0 = v[0], x1 = v[1], x2 = v[2], x3 = v[3], x4 = v[4],
                     x5 = v[5];
    // (fewer than 16 root terms per sum: a longer one would be cut into
    // per-term sub-rows, which moves the last bit of some adjoints)
    slp::Variable<T> J =
        abs(x0 - T(0.1)) + acos(x0 * x1) + asin(x1 * x2) + atan(x2 * x3) +
        atan2(x3, x4) + cbrt(x4 * x5 + T(1)) + cosh(x0) + erf(x1) +
        exp(x2 * T(0.5)) + hypot(x3, x5) + log(x4 + T(1)) + log10(x5 + T(2));
    problem.minimize(J);
    problem.subject_to(sin(x0) * cos(x1) == T(0.2));
    problem.subject_to(x2 * x3 / (x4 + T(1)) == T(0.3));
    problem.subject_to(max(x0 * x0, x1) + min(x2 * x3, x4) + pow(x3, x5) +
                           pow(x0 + T(2), T(2.5)) ==
                       T(12));
    problem.subject_to(hypot(x0, x1, x2) <= T(5));
    problem.subject_to(exp(x3) * tanh(x4) <= T(40));
    problem.subject_to(sign(x1) * x2 + sinh(x3) + tan(x4 * T(0.3)) +
                           tanh(x5) + sqrt(x0 + x1 + T(1)) >=
                       T(0));
    problem.subject_to(x5 >= T(-1));
  } else if (name == "spy_test") {  // problem_spy_test.cpp:64-85
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    x.set_value(T(20));
    y.set_value(T(20));
    problem.minimize(pow(x, T(4)) + pow(y, T(4)));
    problem.subject_to(x >= T(1));
    problem.subject_to(x <= T(10));
    problem.subject_to(y == T(2));
  } else if (name == "unconstrained_1d") {
    auto x = problem.decision_variable();
    x.set_value(T(2));
    problem.minimize(x * x - T(6) * x);
  } else if (name == "unconstrained_2d") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    x.set_value(T(1));
    y.set_value(T(2));
    problem.minimize(x * x + y * y);
  } else if (name == "eq_maximize_xy") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    problem.maximize(x * y);
    problem.subject_to(x + T(3) * y == T(36));
  } else if (name == "eq_pin_2d") {
    auto x = problem.decision_variable(2);
    x[0].set_value(T(1));
    x[1].set_value(T(2));
    problem.minimize(x.T() * x);
    problem.subject_to(x == slp::Matrix<double>{{3.0}, {3.0}});
  } else if (name == "min_distance_line") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    x.set_value(T(20));
    y.set_value(T(50));
    problem.minimize(sqrt(x * x + y * y));
    problem.subject_to(y == -x + T(5));
  } else if (name == "too_few_dofs") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    auto z = problem.decision_variable();
    problem.subject_to(x == T(1));
    problem.subject_to(x == T(2));
    problem.subject_to(y == T(1));
    problem.subject_to(z == T(1));
  } else if (name == "locally_infeasible_eq") {
    auto x = problem.decision_variable();
    auto y = problem.decision_variable();
    auto z = problem.decision_variable();
    problem.subject_to(x == y + T(1));
    problem.subject_to(y == z + T(1));
    problem.subject_to(z == x + T(1));
  } else if (name == "nonfinite_cost") {
    auto x = problem.decision_variable();
    problem.minimize(T(1) / x);
  } else if (name == "nonfinite_gradient") {
    auto x = problem.decision_variable();
    problem.minimize(sqrt(x));
  } else if (name == "nonfinite_eq") {
    auto x = problem.decision_variable();
    problem.subject_to(T(1) / x == T(1));
  } else if (name == "nonfinite_eq_jacobian") {
    auto x = problem.decision_variable();
    problem.subject_to(sqrt(x) == T(1));
  } else if (name == "diverging") {
    auto x = problem.decision_variable();
    problem.minimize(x);
  } else if (name == "min_x_squared") {
    auto x = problem.decision_variable();
    x.set_value(T(1));
    problem.minimize(x * x);
  } else {
    throw std::invalid_argument("unknown problem: " + name);
  }
  return P;
}

inline std::unique_ptr<slp::Problem<double>> make_problem(
    const std::string& name, int N, double p0, double p1) {
  if (name == "cart_pole") return cart_pole(N, p0 > 0 ? p0 : 5.0);
  if (name == "flywheel") return flywheel(N, p0 > 0 ? p0 : 5.0);
  if (name == "gfold") return gfold(N, p0 > 0 ? p0 : 48.0);
  if (name == "cart_pole_eq") return cart_pole(N, p0 > 0 ? p0 : 5.0, false);
  if (name == "flywheel_eq") return flywheel(N, p0 > 0 ? p0 : 5.0, false);
  if (name == "chained_rosenbrock") return chained_rosenbrock(N);
  if (name == "double_integrator") return double_integrator(N);
  if (name == "arm_on_elevator") return arm_on_elevator(N);
  if (name == "differential_drive") return differential_drive(N);
  if (name == "flywheel_ocp") {
    return flywheel_ocp(N, static_cast<int>(p0), p1 != 0.0);
  }
  // the transcription variants of the reference's flywheel OCP test by name
  if (name == "flywheel_ocp_collocation") return flywheel_ocp(N, 1, false);
  if (name == "flywheel_ocp_shooting") return flywheel_ocp(N, 2, false);
  if (name == "flywheel_ocp_discrete") return flywheel_ocp(N, 0, true);
  if (name == "cart_pole_ocp") return cart_pole_ocp(N);
  if (name == "differential_drive_ocp") return differential_drive_ocp(N);
  return small_problem(name, p0, p1);
}

}  // namespace slpb_problems
