// ORACLE — TEST INFRASTRUCTURE ONLY (see expr.hpp header).
//
// Problem builders restated operator-for-operator from the reference:
//   cart-pole   benchmarks/scalability/cart_pole/sleipnir.cpp:16-129,
//               benchmarks/rk4.hpp:14-23
//   flywheel    benchmarks/scalability/flywheel/sleipnir.cpp:12-43
//   g-fold      examples/g-fold/src/main.cpp:145-386 (dt rescaled to T_f/N)
//   small NLPs  test/src/optimization/{linear,quadratic,nonlinear}_problem_test.cpp
// Operand order is preserved everywhere: it decides grad_l vs grad_r and hence
// floating-point rounding.
#pragma once

#include <algorithm>
#include <cmath>
#include <functional>
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
std::unique_ptr<Problem<B>> cart_pole_problem(int N, double T = 5.0,
                                              bool bounded = true) {
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
  if (bounded) {  // without the bounds the reference takes its SQP branch
    problem->subject_to_ineq(bounds(V{0.0}, X.row(0), V{d_max}));
    problem->subject_to_ineq(bounds(V{-u_max}, U, V{u_max}));
  }
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
std::unique_ptr<Problem<B>> flywheel_problem(int N, double T = 5.0,
                                             bool bounded = true) {
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
  if (bounded) problem->subject_to_ineq(bounds(V{-12}, U, V{12}));
  M r = M::constants(1, 1, {10.0});
  V J{0.0};
  for (int k = 0; k < N + 1; ++k) {
    M e = (r - X.col(k)).T() * (r - X.col(k));
    J += e(0, 0);
  }
  problem->minimize(J);
  return problem;
}

/// Σ 100(xᵢ₊₁ − xᵢ²)² + (1 − xᵢ)² from the classic (−1.2, 1, −1.2, …) start:
/// an unconstrained problem of any size for the Newton branch
/// (problem.hpp:335).
template <class B>
std::unique_ptr<Problem<B>> chained_rosenbrock_problem(int N) {
  using M = Mat<B>;
  using V = Var<B>;
  auto problem = std::make_unique<Problem<B>>();
  M X = problem->decision_variable(N + 1, 1);
  for (int k = 0; k < N + 1; ++k) X(k, 0).set_value(k % 2 == 0 ? -1.2 : 1.0);
  V J{0.0};
  for (int k = 0; k < N; ++k) {
    V x = X(k, 0);
    V y = X(k + 1, 0);
    J += V{100} * pow(y - pow(x, 2.0), 2.0) + pow(V{1} - x, 2.0);
  }
  problem->minimize(J);
  return problem;
}

/// double_integrator_problem_test.cpp:27-127 with dt = 3.5 s / N.
template <class B>
std::unique_ptr<Problem<B>> double_integrator_problem(int N) {
  using M = Mat<B>;
  using V = Var<B>;
  const double t = 3.5 / N;
  constexpr double r = 2.0;
  auto problem = std::make_unique<Problem<B>>();
  M X = problem->decision_variable(2, N + 1);
  M U = problem->decision_variable(1, N);
  for (int k = 0; k < N; ++k) {
    V p_k1 = X(0, k + 1), v_k1 = X(1, k + 1);
    V p_k = X(0, k), v_k = X(1, k), a_k = U(0, k);
    problem->subject_to_eq(
        {p_k1 - (p_k + v_k * V{t} + V{0.5} * a_k * V{t} * V{t})});
    problem->subject_to_eq({v_k1 - (v_k + a_k * V{t})});
  }
  problem->subject_to_eq(eq(X.col(0), M::constants(2, 1, {0.0, 0.0})));
  problem->subject_to_eq(eq(X.col(N), M::constants(2, 1, {r, 0.0})));
  problem->subject_to_ineq(bounds(V{-1.0}, X.row(1), V{1.0}));
  problem->subject_to_ineq(bounds(V{-1.0}, U, V{1.0}));
  V J{0.0};
  for (int k = 0; k < N + 1; ++k) J += pow(V{r} - X(0, k), 2.0);
  problem->minimize(J);
  return problem;
}

/// arm_on_elevator_problem_test.cpp:27-122 with dt = 4 s / N.
template <class B>
std::unique_ptr<Problem<B>> arm_on_elevator_problem(int N) {
  using M = Mat<B>;
  using V = Var<B>;
  constexpr double pi = std::numbers::pi;
  const double dt = 4.0 / N;
  auto problem = std::make_unique<Problem<B>>();
  M elevator = problem->decision_variable(2, N + 1);
  M elevator_accel = problem->decision_variable(1, N);
  M arm = problem->decision_variable(2, N + 1);
  M arm_accel = problem->decision_variable(1, N);
  for (int k = 0; k < N; ++k) {
    problem->subject_to_eq(
        {elevator(0, k + 1) -
         (elevator(0, k) + elevator(1, k) * V{dt} +
          V{0.5} * elevator_accel(0, k) * V{dt} * V{dt})});
    problem->subject_to_eq(
        {elevator(1, k + 1) -
         (elevator(1, k) + elevator_accel(0, k) * V{dt})});
    problem->subject_to_eq(
        {arm(0, k + 1) - (arm(0, k) + arm(1, k) * V{dt} +
                          V{0.5} * arm_accel(0, k) * V{dt} * V{dt})});
    problem->subject_to_eq(
        {arm(1, k + 1) - (arm(1, k) + arm_accel(0, k) * V{dt})});
  }
  problem->subject_to_eq(eq(elevator.col(0), M::constants(2, 1, {1.0, 0.0})));
  problem->subject_to_eq(eq(elevator.col(N), M::constants(2, 1, {1.25, 0.0})));
  problem->subject_to_eq(eq(arm.col(0), M::constants(2, 1, {0.0, 0.0})));
  problem->subject_to_eq(eq(arm.col(N), M::constants(2, 1, {pi, 0.0})));
  problem->subject_to_ineq(bounds(V{-1.0}, elevator.row(1), V{1.0}));
  problem->subject_to_ineq(bounds(V{-2.0}, elevator_accel, V{2.0}));
  problem->subject_to_ineq(bounds(V{-2.0 * pi}, arm.row(1), V{2.0 * pi}));
  problem->subject_to_ineq(bounds(V{-4.0 * pi}, arm_accel, V{4.0 * pi}));
  M sines{1, N + 1};
  for (int k = 0; k < N + 1; ++k) sines(0, k) = sin(arm(0, k));
  M heights = elevator.row(0) + V{1.0} * sines;
  problem->subject_to_ineq(le(heights, V{1.8}));
  V J{0.0};
  for (int k = 0; k < N + 1; ++k) {
    J += pow(V{1.25} - elevator(0, k), 2.0) + pow(V{pi} - arm(0, k), 2.0);
  }
  problem->minimize(J);
  return problem;
}

/// test/include/differential_drive_util.hpp:16-61.
template <class B>
Mat<B> differential_drive_dynamics(const Mat<B>& x, const Mat<B>& u) {
  using M = Mat<B>;
  using V = Var<B>;
  constexpr double trackwidth = 0.699, Kv_l = 3.02, Ka_l = 0.642;
  constexpr double Kv_a = 1.382, Ka_a = 0.08495;
  constexpr double A1 = -(Kv_l / Ka_l + Kv_a / Ka_a) / 2.0;
  constexpr double A2 = -(Kv_l / Ka_l - Kv_a / Ka_a) / 2.0;
  constexpr double B1 = 0.5 / Ka_l + 0.5 / Ka_a;
  constexpr double B2 = 0.5 / Ka_l - 0.5 / Ka_a;
  M Am = M::constants(2, 2, {A1, A2, A2, A1});
  M Bm = M::constants(2, 2, {B1, B2, B2, B1});
  M xdot{5, 1};
  V v = (x[3] + x[4]) / V{2.0};
  xdot[0] = v * cos(x[2]);
  xdot[1] = v * sin(x[2]);
  xdot[2] = (x[4] - x[3]) / V{trackwidth};
  xdot.set_block(3, 0, Am * x.segment(3, 2) + Bm * u);
  return xdot;
}

/// differential_drive_problem_test.cpp:28-138 with dt = 5 s / N.
template <class B>
std::unique_ptr<Problem<B>> differential_drive_problem(int N) {
  using M = Mat<B>;
  using V = Var<B>;
  const double dt = 5.0 / N;
  auto problem = std::make_unique<Problem<B>>();
  M X = problem->decision_variable(5, N + 1);
  for (int k = 0; k < N; ++k) {
    X(0, k).set_value(std::lerp(0.0, 1.0, static_cast<double>(k) / N));
    X(1, k).set_value(std::lerp(0.0, 1.0, static_cast<double>(k) / N));
  }
  M U = problem->decision_variable(2, N);
  problem->subject_to_eq(
      eq(X.col(0), M::constants(5, 1, {0.0, 0.0, 0.0, 0.0, 0.0})));
  problem->subject_to_eq(
      eq(X.col(N), M::constants(5, 1, {1.0, 1.0, 0.0, 0.0, 0.0})));
  problem->subject_to_ineq(bounds(V{-12.0}, U, V{12.0}));
  for (int k = 0; k < N; ++k) {
    problem->subject_to_eq(
        eq(X.col(k + 1),
           rk4<B>(differential_drive_dynamics<B>, X.col(k), U.col(k), dt)));
  }
  V J{0.0};
  for (int k = 0; k < N; ++k) {
    M xx = X.col(k).T() * X.col(k);
    M uu = U.col(k).T() * U.col(k);
    J += xx(0, 0) + uu(0, 0);
  }
  problem->minimize(J);
  return problem;
}

// ---- slp::OCP (optimization/ocp.hpp:49-414) ----------------------------------
enum class OcpDynamics { EXPLICIT_ODE, DISCRETE };
enum class OcpTimestep { FIXED, VARIABLE_SINGLE, VARIABLE };
enum class OcpTranscription {
  DIRECT_TRANSCRIPTION,
  DIRECT_COLLOCATION,
  SINGLE_SHOOTING
};

/// What the OCP constructor and its helpers add to a Problem, in the
/// reference's order: U, then the time step(s), then X (ocp.hpp:141-178).
template <class B>
struct Ocp {
  using M = Mat<B>;
  using V = Var<B>;
  using Dynamics =
      std::function<M(const V& t, const M& x, const M& u, const V& dt)>;

  Problem<B>& problem;
  int num_steps;
  Dynamics f;
  OcpDynamics dynamics_type;
  M X, U, DT;

  Ocp(Problem<B>& p, int num_states, int num_inputs, double dt, int steps,
      Dynamics dynamics, OcpDynamics dyn_type, OcpTimestep timestep,
      OcpTranscription transcription)
      : problem{p}, num_steps{steps}, f{std::move(dynamics)},
        dynamics_type{dyn_type} {
    const int samples = num_steps + 1;
    U = problem.decision_variable(num_inputs, samples);
    DT = M{1, samples};
    if (timestep == OcpTimestep::FIXED) {
      for (int i = 0; i < samples; ++i) DT(0, i) = V{dt};
    } else if (timestep == OcpTimestep::VARIABLE_SINGLE) {
      V single = problem.decision_variable();
      single.set_value(dt);
      for (int i = 0; i < samples; ++i) DT(0, i) = single;
    } else {
      DT = problem.decision_variable(1, samples);
      for (int i = 0; i < samples; ++i) DT(0, i).set_value(dt);
    }
    if (transcription == OcpTranscription::SINGLE_SHOOTING) {
      X = M{num_states, samples};
      V time{0.0};
      for (int i = 0; i < num_steps; ++i) {  // :394-414
        X.set_block(0, i + 1, advance(X.col(i), U.col(i), time, DT(0, i)));
        time += DT(0, i);
      }
    } else if (transcription == OcpTranscription::DIRECT_TRANSCRIPTION) {
      X = problem.decision_variable(num_states, samples);
      V time{0.0};
      for (int i = 0; i < num_steps; ++i) {  // :371-392
        problem.subject_to_eq(
            eq(X.col(i + 1), advance(X.col(i), U.col(i), time, DT(0, i))));
        time += DT(0, i);
      }
    } else {
      X = problem.decision_variable(num_states, samples);
      V time{0.0};
      for (int i = 0; i < num_steps; ++i) {  // :334-369
        V h = DT(0, i);
        V t_begin = time;
        V t_end = t_begin + h;
        M x_begin = X.col(i), x_end = X.col(i + 1);
        M u_begin = U.col(i), u_end = U.col(i + 1);
        M xdot_begin = f(t_begin, x_begin, u_begin, h);
        M xdot_end = f(t_end, x_end, u_end, h);
        M xdot_c = V{-3.0} / (V{2.0} * h) * (x_begin - x_end) -
                   V{0.25} * (xdot_begin + xdot_end);
        V t_c = t_begin + V{0.5} * h;
        M x_c = V{0.5} * (x_begin + x_end) +
                h / V{8.0} * (xdot_begin - xdot_end);
        M u_c = V{0.5} * (u_begin + u_end);
        problem.subject_to_eq(eq(xdot_c, f(t_c, x_c, u_c, h)));
        time += h;
      }
    }
  }

  M rk4(const M& x, const M& u, const V& t0, const V& dt) const {  // :322-332
    V halfdt = dt * V{0.5};
    M k1 = f(t0, x, u, dt);
    M k2 = f(t0 + halfdt, x + k1 * halfdt, u, dt);
    M k3 = f(t0 + halfdt, x + k2 * halfdt, u, dt);
    M k4 = f(t0 + dt, x + k3 * dt, u, dt);
    return x + (k1 + k2 * V{2.0} + k3 * V{2.0} + k4) * (dt / V{6.0});
  }
  M advance(const M& x, const M& u, const V& t, const V& dt) const {
    return dynamics_type == OcpDynamics::EXPLICIT_ODE ? rk4(x, u, t, dt)
                                                       : f(t, x, u, dt);
  }
  void lower_input_bound(const M& lo) {  // :242-254
    for (int i = 0; i < num_steps + 1; ++i) {
      problem.subject_to_ineq(ge(U.col(i), lo));
    }
  }
  void upper_input_bound(const M& hi) {  // :256-268
    for (int i = 0; i < num_steps + 1; ++i) {
      problem.subject_to_ineq(le(U.col(i), hi));
    }
  }
};

/// flywheel_ocp_test.cpp:38-201 with dt = 5 s / N.
template <class B>
std::unique_ptr<Problem<B>> flywheel_ocp_problem(int N, int method,
                                                 int discrete) {
  using M = Mat<B>;
  using V = Var<B>;
  const double dt = 5.0 / N;
  constexpr double A = -1.0, Bc = 1.0;
  const double A_d = std::exp(A * dt);
  const double B_d = (1.0 - A_d) * Bc;
  auto problem = std::make_unique<Problem<B>>();
  typename Ocp<B>::Dynamics f;
  if (discrete) {
    f = [=](const V&, const M& x, const M& u, const V&) {
      return V{A_d} * x + V{B_d} * u;
    };
  } else {
    f = [=](const V&, const M& x, const M& u, const V&) {
      return V{A} * x + V{Bc} * u;
    };
  }
  Ocp<B> ocp{*problem, 1, 1, dt, N, f,
             discrete ? OcpDynamics::DISCRETE : OcpDynamics::EXPLICIT_ODE,
             OcpTimestep::FIXED, static_cast<OcpTranscription>(method)};
  problem->subject_to_eq(eq(ocp.X.col(0), V{0.0}));
  ocp.upper_input_bound(M::constants(1, 1, {12.0}));
  ocp.lower_input_bound(M::constants(1, 1, {-12.0}));
  M r = M::constants(1, N + 1, std::vector<double>(N + 1, 10.0));
  M J = (r - ocp.X) * (r - ocp.X).T();
  problem->minimize(J(0, 0));
  return problem;
}

/// cart_pole_ocp_test.cpp:29-86 with dt = 5 s / N: Hermite–Simpson
/// collocation, one shared variable time step.
template <class B>
std::unique_ptr<Problem<B>> cart_pole_ocp_problem(int N) {
  using M = Mat<B>;
  using V = Var<B>;
  const double dt = 5.0 / N;
  const std::vector<double> x_initial{0.0, 0.0, 0.0, 0.0};
  const std::vector<double> x_final{1.0, std::numbers::pi, 0.0, 0.0};
  auto problem = std::make_unique<Problem<B>>();
  Ocp<B> ocp{*problem, 4, 1, dt, N,
             [](const V&, const M& x, const M& u, const V&) {
               return cart_pole_dynamics<B>(x, u);
             },
             OcpDynamics::EXPLICIT_ODE, OcpTimestep::VARIABLE_SINGLE,
             OcpTranscription::DIRECT_COLLOCATION};
  for (int k = 0; k < N + 1; ++k) {
    ocp.X(0, k).set_value(
        std::lerp(x_initial[0], x_final[0], static_cast<double>(k) / N));
    ocp.X(1, k).set_value(
        std::lerp(x_initial[1], x_final[1], static_cast<double>(k) / N));
  }
  problem->subject_to_eq(eq(ocp.X.col(0), M::constants(4, 1, x_initial)));
  problem->subject_to_eq(eq(ocp.X.col(N), M::constants(4, 1, x_final)));
  for (int k = 0; k < N + 1; ++k) {
    problem->subject_to_ineq(
        bounds(V{0.0}, M{ocp.X(0, k)}, V{2.0}));
  }
  ocp.lower_input_bound(M::constants(1, 1, {-20.0}));
  ocp.upper_input_bound(M::constants(1, 1, {20.0}));
  V J{0.0};
  for (int k = 0; k < N; ++k) {
    M uu = ocp.U.col(k).T() * ocp.U.col(k);
    J += uu(0, 0);
  }
  problem->minimize(J);
  return problem;
}

/// differential_drive_ocp_test.cpp:25-66: minimum time, one shared variable
/// time step, direct transcription.
template <class B>
std::unique_ptr<Problem<B>> differential_drive_ocp_problem(int N) {
  using M = Mat<B>;
  using V = Var<B>;
  constexpr double trackwidth = 0.699, Kv_l = 3.02, Ka_l = 0.642;
  constexpr double Kv_a = 1.382, Ka_a = 0.08495;
  constexpr double A1 = -(Kv_l / Ka_l + Kv_a / Ka_a) / 2.0;
  constexpr double A2 = -(Kv_l / Ka_l - Kv_a / Ka_a) / 2.0;
  constexpr double B1 = 0.5 / Ka_l + 0.5 / Ka_a;
  constexpr double B2 = 0.5 / Ka_l - 0.5 / Ka_a;
  constexpr double min_timestep = 0.05;
  auto problem = std::make_unique<Problem<B>>();
  Ocp<B> ocp{*problem, 5, 2, min_timestep, N,
             [=](const V&, const M& x, const M& u, const V&) {
               M Am = M::constants(2, 2, {A1, A2, A2, A1});
               M Bm = M::constants(2, 2, {B1, B2, B2, B1});
               M xdot{5, 1};
               V v = (x[3] + x[4]) / V{2.0};
               xdot[0] = v * cos(x[2]);
               xdot[1] = v * sin(x[2]);
               xdot[2] = (x[4] - x[3]) / V{trackwidth};
               xdot.set_block(3, 0, Am * x.segment(3, 2) + Bm * u);
               return xdot;
             },
             OcpDynamics::EXPLICIT_ODE, OcpTimestep::VARIABLE_SINGLE,
             OcpTranscription::DIRECT_TRANSCRIPTION};
  for (int i = 0; i < N + 1; ++i) {
    ocp.X(0, i).set_value(static_cast<double>(i) / (N + 1));
    ocp.X(1, i).set_value(static_cast<double>(i) / (N + 1));
  }
  problem->subject_to_eq(
      eq(ocp.X.col(0), M::constants(5, 1, {0.0, 0.0, 0.0, 0.0, 0.0})));
  problem->subject_to_eq(
      eq(ocp.X.col(N), M::constants(5, 1, {1.0, 1.0, 0.0, 0.0, 0.0})));
  ocp.lower_input_bound(M::constants(2, 1, {-12.0, -12.0}));
  ocp.upper_input_bound(M::constants(2, 1, {12.0, 12.0}));
  problem->subject_to_ineq(ge(ocp.DT, V{min_timestep}));
  problem->subject_to_ineq(le(ocp.DT, V{3.0}));
  M J = ocp.DT * M::constants(N + 1, 1, std::vector<double>(N + 1, 1.0));
  problem->minimize(J(0, 0));
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
  } else if (name == "mishra_bird") {  // multistart_test.cpp:17-55
    V x = P->decision_variable();
    V y = P->decision_variable();
    x.set_value(p0);
    y.set_value(p1);
    V J = sin(y) * exp(pow(V{1} - cos(x), 2.0)) +
          cos(x) * exp(pow(V{1} - sin(y), 2.0)) + pow(x - y, 2.0);
    P->minimize(J);
    P->subject_to_ineq(
        le1(pow(x + V{5}, 2.0) + pow(y + V{5}, 2.0), V{25}));
  } else if (name == "empty") {  // trivial_problem_test.cpp:14-24
  } else if (name == "no_cost_unconstrained") {  // :26-66
    M X = P->decision_variable(2, 3);
    for (int i = 0; i < 6; ++i) X[i].set_value(p0);
  } else if (name == "all_ops") {
    // every operation of autodiff/expression.hpp once (see the product's
    // problems.hpp): golden vectors from the reference core pin them
    M v = P->decision_variable(6, 1);
    const double guess[6] = {0.3, 0.5, 0.7, 1.1, 1.3, 0.9};
    for (int i = 0; i < 6; ++i) v[i].set_value(guess[i]);
    V x0 = v[0], x1 = v[1], x2 = v[2], x3 = v[3], x4 = v[4], x5 = v[5];
    V J = abs(x0 - V{0.1}) + acos(x0 * x1) + asin(x1 * x2) + atan(x2 * x3) +
          atan2(x3, x4) + cbrt(x4 * x5 + V{1}) + cosh(x0) + erf(x1) +
          exp(x2 * V{0.5}) + hypot(x3, x5) + log(x4 + V{1}) +
          log10(x5 + V{2});
    P->minimize(J);
    P->subject_to_eq({sin(x0) * cos(x1) - V{0.2}});
    P->subject_to_eq({x2 * x3 / (x4 + V{1}) - V{0.3}});
    P->subject_to_eq({max(x0 * x0, x1) + min(x2 * x3, x4) + pow(x3, x5) +
                      pow(x0 + V{2}, 2.5) - V{12}});
    P->subject_to_ineq(le1(hypot3(x0, x1, x2), V{5}));
    P->subject_to_ineq(le1(exp(x3) * tanh(x4), V{40}));
    P->subject_to_ineq(ge1(sign(x1) * x2 + sinh(x3) + tan(x4 * V{0.3}) +
                               tanh(x5) + sqrt(x0 + x1 + V{1}),
                           V{0}));
    P->subject_to_ineq(ge1(x5, V{-1}));
  } else if (name == "spy_test") {  // problem_spy_test.cpp:64-85
    V x = P->decision_variable();
    V y = P->decision_variable();
    x.set_value(20);
    y.set_value(20);
    P->minimize(pow(x, 4.0) + pow(y, 4.0));
    P->subject_to_ineq(ge1(x, V{1}));
    P->subject_to_ineq(le1(x, V{10}));
    P->subject_to_eq({y - V{2}});
  } else if (name == "unconstrained_1d") {  // quadratic_problem_test.cpp:15-34
    V x = P->decision_variable();
    x.set_value(2);
    P->minimize(x * x - V{6} * x);
  } else if (name == "unconstrained_2d") {  // :36-58
    V x = P->decision_variable();
    V y = P->decision_variable();
    x.set_value(1);
    y.set_value(2);
    P->minimize(x * x + y * y);
  } else if (name == "eq_maximize_xy") {  // :80-141
    V x = P->decision_variable();
    V y = P->decision_variable();
    P->maximize(x * y);
    P->subject_to_eq({x + V{3} * y - V{36}});
  } else if (name == "eq_pin_2d") {  // :143-161
    M x = P->decision_variable(2, 1);
    x(0, 0).set_value(1);
    x(1, 0).set_value(2);
    M xx = x.T() * x;
    P->minimize(xx(0, 0));
    P->subject_to_eq(eq(x, M::constants(2, 1, {3.0, 3.0})));
  } else if (name == "min_distance_line") {  // nonlinear_problem_test.cpp:120-143
    V x = P->decision_variable();
    V y = P->decision_variable();
    x.set_value(20);
    y.set_value(50);
    P->minimize(sqrt(x * x + y * y));
    P->subject_to_eq({y - (-x + V{5})});
  } else if (name == "too_few_dofs") {  // exit_status_test.cpp:52-72
    V x = P->decision_variable();
    V y = P->decision_variable();
    V z = P->decision_variable();
    P->subject_to_eq({x - V{1}});
    P->subject_to_eq({x - V{2}});
    P->subject_to_eq({y - V{1}});
    P->subject_to_eq({z - V{1}});
  } else if (name == "locally_infeasible_eq") {  // :78-95
    V x = P->decision_variable();
    V y = P->decision_variable();
    V z = P->decision_variable();
    P->subject_to_eq({x - (y + V{1})});
    P->subject_to_eq({y - (z + V{1})});
    P->subject_to_eq({z - (x + V{1})});
  } else if (name == "nonfinite_cost") {  // :124-130
    V x = P->decision_variable();
    P->minimize(V{1} / x);
  } else if (name == "nonfinite_gradient") {  // :133-139
    V x = P->decision_variable();
    P->minimize(sqrt(x));
  } else if (name == "nonfinite_eq") {  // :142-148
    V x = P->decision_variable();
    P->subject_to_eq({V{1} / x - V{1}});
  } else if (name == "nonfinite_eq_jacobian") {  // :151-157
    V x = P->decision_variable();
    P->subject_to_eq({sqrt(x) - V{1}});
  } else if (name == "diverging") {  // :178-194
    V x = P->decision_variable();
    P->minimize(x);
  } else if (name == "min_x_squared") {  // :17-50, :196-234
    V x = P->decision_variable();
    x.set_value(1);
    P->minimize(x * x);
  } else {
    throw std::invalid_argument("unknown oracle problem: " + name);
  }
  return P;
}


/// exp([A B; 0 0]·dt) by scaling and squaring of a Taylor series — the same
/// routine as the product's builder (sleipnir_b200/problems/problems.hpp), so
/// both sides transcribe the same constants. (The reference uses Eigen's
/// matrix exponential, examples/g-fold/src/main.cpp:39-57; Eigen is absent.)
inline void discretize_ab(const std::vector<double>& A,
                          const std::vector<double>& B, int ns, int ni,
                          double dt, std::vector<double>& A_d,
                          std::vector<double>& B_d) {
  const int n = ns + ni;
  std::vector<double> M(size_t(n) * n, 0.0), E(size_t(n) * n, 0.0),
      T(size_t(n) * n, 0.0), tmp(size_t(n) * n, 0.0);
  for (int r = 0; r < ns; ++r) {
    for (int c = 0; c < ns; ++c) M[r * n + c] = A[r * ns + c] * dt;
    for (int c = 0; c < ni; ++c) M[r * n + ns + c] = B[r * ni + c] * dt;
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
                    std::vector<double>& Zm) {
    for (int r = 0; r < n; ++r) {
      for (int c = 0; c < n; ++c) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += X[r * n + k] * Y[k * n + c];
        Zm[r * n + c] = acc;
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
  A_d.assign(size_t(ns) * ns, 0.0);
  B_d.assign(size_t(ns) * ni, 0.0);
  for (int r = 0; r < ns; ++r) {
    for (int c = 0; c < ns; ++c) A_d[r * ns + c] = E[r * n + c];
    for (int c = 0; c < ni; ++c) B_d[r * ni + c] = E[r * n + ns + c];
  }
}

/// G-FOLD powered-descent guidance, examples/g-fold/src/main.cpp:145-386, with
/// dt = T_f/N (SURVEY.md §7 hard part 8).
template <class B>
std::unique_ptr<Problem<B>> gfold_problem(int N, double T_f = 48.0) {
  using M = Mat<B>;
  using V = Var<B>;
  constexpr double m_wet = 2000.0;
  constexpr double T_max = 24000;
  constexpr double rho_1 = 0.2 * T_max;
  constexpr double rho_2 = 0.8 * T_max;
  constexpr double alpha = 5e-4;
  const std::vector<double> q_0{2400.0, 450.0, -330.0};
  const std::vector<double> v_0{-10.0, -40.0, 10.0};
  const std::vector<double> q_f{0.0, 0.0, 0.0};
  const std::vector<double> v_f{0.0, 0.0, 0.0};
  const std::vector<double> gv{-3.71, 0.0, 0.0};
  constexpr double w1 = 2.53e-5, w2 = 0.0, w3 = 6.62e-5;
  const double theta = 90.0 * std::numbers::pi / 180.0;
  const double gamma_gs = 30.0 * std::numbers::pi / 180.0;
  constexpr double v_max = 90.0;
  const double dt = T_f / N;
  constexpr bool END_STRAIGHT = true;

  const double S[3][3] = {{0.0, -w3, w2}, {w3, 0.0, -w1}, {-w2, w1, 0.0}};
  std::vector<double> A(36, 0.0), Bc(18, 0.0), A_d, B_d;
  for (int i = 0; i < 3; ++i) {
    A[i * 6 + 3 + i] = 1.0;
    Bc[(3 + i) * 3 + i] = 1.0;
    for (int j = 0; j < 3; ++j) {
      double ss = 0.0;
      for (int k = 0; k < 3; ++k) ss += S[i][k] * S[k][j];
      A[(3 + i) * 6 + j] = -ss;
      A[(3 + i) * 6 + 3 + j] = -2 * S[i][j];
    }
  }
  discretize_ab(A, Bc, 6, 3, dt, A_d, B_d);
  const M Ad = M::constants(6, 6, A_d);
  const M Bd = M::constants(6, 3, B_d);
  const M g = M::constants(3, 1, gv);

  auto problem = std::make_unique<Problem<B>>();
  M X = problem->decision_variable(6, N + 1);
  M Z = problem->decision_variable(1, N + 1);
  M U = problem->decision_variable(3, N);
  M sigma = problem->decision_variable(1, N);

  M q = X.block(0, 0, 3, N + 1);
  M v = X.block(3, 0, 3, N + 1);

  problem->subject_to_eq(eq(q.col(0), M::constants(3, 1, q_0)));
  problem->subject_to_eq(eq(v.col(0), M::constants(3, 1, v_0)));
  problem->subject_to_eq(eq(M{Z(0, 0)}, V{std::log(m_wet)}));
  problem->subject_to_eq(eq(q.col(N), M::constants(3, 1, q_f)));
  problem->subject_to_eq(eq(v.col(N), M::constants(3, 1, v_f)));

  for (int k = 0; k < N + 1; ++k) {
    for (int i = 0; i < 3; ++i) {
      q(i, k).set_value(std::lerp(q_0[i], q_f[i], static_cast<double>(k) / N));
      v(i, k).set_value(std::lerp(v_0[i], v_f[i], static_cast<double>(k) / N));
    }
  }

  for (int k = 0; k < N + 1; ++k) {
    const double t = k * dt;
    M x_k = X.col(k);
    M q_k = X.block(0, k, 3, 1);
    M v_k = X.block(3, k, 3, 1);
    M z_k = Z.col(k);

    problem->subject_to_ineq(le(v_k.T() * v_k, V{v_max * v_max}));

    const double z_min = std::log(m_wet - alpha * rho_2 * t);
    const double z_max = std::log(m_wet - alpha * rho_1 * t);
    const double z_estimate = (z_min + z_max) / 2;
    z_k[0].set_value(z_estimate);

    if (k < N) {
      M x_k1 = X.col(k + 1);
      M z_k1 = Z.col(k + 1);
      M u_k = U.col(k);
      M sigma_k = sigma.col(k);

      const double u_min = rho_1 / std::exp(z_estimate);
      const double u_max = rho_2 / std::exp(z_estimate);
      u_k.set_value(std::vector<double>{(u_min + u_max) / 2, 0.0, 0.0});

      {
        V lhs = pow(q_k[0] - V{q_f[0]}, 2.0);
        V rhs = V{std::tan(gamma_gs) * std::tan(gamma_gs)} *
                (pow(q_k[1] - V{q_f[1]}, 2.0) + pow(q_k[2] - V{q_f[2]}, 2.0));
        problem->subject_to_ineq(std::vector<V>{lhs - rhs});
      }

      problem->subject_to_ineq(ge(sigma_k, V{0}));

      if (k == N - 1 && END_STRAIGHT) {
        problem->subject_to_eq(eq(M{u_k(0, 0)}, sigma_k));
        problem->subject_to_eq(eq(M{u_k(1, 0)}, V{0}));
        problem->subject_to_eq(eq(M{u_k(2, 0)}, V{0}));
      } else {
        // u_kᵀu_k ≤ σ_k·σ_k  →  rhs − lhs
        problem->subject_to_ineq(ge(sigma_k * sigma_k, u_k.T() * u_k));
        problem->subject_to_ineq(ge(u_k[0], std::cos(theta) * sigma_k));
      }

      const double z_0 = std::log(m_wet - alpha * rho_2 * t);
      const double mu_1 = rho_1 * std::exp(-z_0);
      const double mu_2 = rho_2 * std::exp(-z_0);
      V sigma_min = V{mu_1} * (V{1} - (z_k[0] - V{z_0}) +
                               V{0.5} * pow(z_k[0] - V{z_0}, 2.0));
      V sigma_max = V{mu_2} * (V{1} - (z_k[0] - V{z_0}));
      problem->subject_to_ineq(bounds(sigma_min, sigma_k, sigma_max));
      sigma_k[0].set_value((sigma_min.value() + sigma_max.value()) / 2);

      problem->subject_to_eq(eq(x_k1, Ad * x_k + Bd * (g + u_k)));
      problem->subject_to_eq(eq(z_k1, z_k - alpha * dt * sigma_k));
    }
  }

  V J{0.0};
  for (int k = 0; k < N; ++k) J = J + sigma(0, k);
  problem->minimize(J);
  return problem;
}

template <class B>
std::unique_ptr<Problem<B>> make_problem(const std::string& name, int N,
                                         double p0, double p1) {
  if (name == "cart_pole") return cart_pole_problem<B>(N, p0 > 0 ? p0 : 5.0);
  if (name == "flywheel") return flywheel_problem<B>(N, p0 > 0 ? p0 : 5.0);
  if (name == "gfold") return gfold_problem<B>(N, p0 > 0 ? p0 : 48.0);
  if (name == "cart_pole_eq") {
    return cart_pole_problem<B>(N, p0 > 0 ? p0 : 5.0, false);
  }
  if (name == "flywheel_eq") {
    return flywheel_problem<B>(N, p0 > 0 ? p0 : 5.0, false);
  }
  if (name == "chained_rosenbrock") return chained_rosenbrock_problem<B>(N);
  if (name == "double_integrator") return double_integrator_problem<B>(N);
  if (name == "arm_on_elevator") return arm_on_elevator_problem<B>(N);
  if (name == "differential_drive") return differential_drive_problem<B>(N);
  if (name == "flywheel_ocp") {
    return flywheel_ocp_problem<B>(N, static_cast<int>(p0),
                                   static_cast<int>(p1));
  }
  // the transcription variants of the reference's flywheel OCP test by name
  if (name == "flywheel_ocp_collocation") return flywheel_ocp_problem<B>(N, 1, false);
  if (name == "flywheel_ocp_shooting") return flywheel_ocp_problem<B>(N, 2, false);
  if (name == "flywheel_ocp_discrete") return flywheel_ocp_problem<B>(N, 0, true);
  if (name == "cart_pole_ocp") return cart_pole_ocp_problem<B>(N);
  if (name == "differential_drive_ocp") {
    return differential_drive_ocp_problem<B>(N);
  }
  return small_problem<B>(name, p0, p1);
}

}  // namespace orc
