// slp::OCP — the reference's optimal-control front end
// (include/sleipnir/optimization/ocp.hpp:49-410) over this repo's Problem: it
// only creates decision variables and constraints, so everything it builds
// runs on the device path unchanged. Node construction follows the reference
// step by step (RK4 stage order, Hermite–Simpson midpoint formula, running
// time variable), because the order of operations fixes the rounding of the
// expression graph.
//
// Why it matters for the device path: the constraints of step k are the same
// expression with the leaves of step k, so the program compiler finds one
// cluster program per constraint row and runs the steps as lanes of one
// kernel (compile.cpp, describe_cluster).
#pragma once

#include <chrono>
#include <functional>
#include <utility>

#include "sleipnir/autodiff/variable_matrix.hpp"
#include "sleipnir/optimization/ocp/dynamics_type.hpp"
#include "sleipnir/optimization/ocp/timestep_method.hpp"
#include "sleipnir/optimization/ocp/transcription_method.hpp"
#include "sleipnir/optimization/problem.hpp"
#include "sleipnir/util/assert.hpp"

namespace slp {

/// An optimal control problem over num_steps + 1 samples of a state X
/// (num_states × (num_steps + 1)) and an input U (num_inputs × (num_steps + 1)).
template <typename Scalar>
class OCP : public Problem<Scalar> {
 public:
  using Matrix = VariableMatrix<Scalar>;
  /// f(x, u)
  using Dynamics = std::function<Matrix(const Matrix& x, const Matrix& u)>;
  /// f(t, x, u, dt)
  using TimedDynamics =
      std::function<Matrix(const Variable<Scalar>& t, const Matrix& x,
                           const Matrix& u, const Variable<Scalar>& dt)>;

  /// ocp.hpp:84-108 — dynamics that do not depend on time.
  OCP(int num_states, int num_inputs, std::chrono::duration<Scalar> dt,
      int num_steps, Dynamics dynamics,
      DynamicsType dynamics_type = DynamicsType::EXPLICIT_ODE,
      TimestepMethod timestep_method = TimestepMethod::FIXED,
      TranscriptionMethod transcription_method =
          TranscriptionMethod::DIRECT_TRANSCRIPTION)
      : OCP{num_states,
            num_inputs,
            dt,
            num_steps,
            TimedDynamics{[f = std::move(dynamics)](
                              const Variable<Scalar>&, const Matrix& x,
                              const Matrix& u,
                              const Variable<Scalar>&) { return f(x, u); }},
            dynamics_type,
            timestep_method,
            transcription_method} {}

  /// ocp.hpp:131-179.
  OCP(int num_states, int num_inputs, std::chrono::duration<Scalar> dt,
      int num_steps, TimedDynamics dynamics,
      DynamicsType dynamics_type = DynamicsType::EXPLICIT_ODE,
      TimestepMethod timestep_method = TimestepMethod::FIXED,
      TranscriptionMethod transcription_method =
          TranscriptionMethod::DIRECT_TRANSCRIPTION)
      : m_num_steps{num_steps},
        m_dynamics{std::move(dynamics)},
        m_dynamics_type{dynamics_type} {
    const int samples = m_num_steps + 1;
    // inputs first, then the time step(s), then the states: the reference's
    // decision-variable order
    m_U = this->decision_variable(num_inputs, samples);

    switch (timestep_method) {
      case TimestepMethod::FIXED:
        m_DT = Matrix{1, samples};
        for (int i = 0; i < samples; ++i) m_DT[0, i] = dt.count();
        break;
      case TimestepMethod::VARIABLE_SINGLE: {
        Variable<Scalar> single_dt = this->decision_variable();
        single_dt.set_value(dt.count());
        m_DT = Matrix{1, samples};
        for (int i = 0; i < samples; ++i) m_DT[0, i] = single_dt;
        break;
      }
      case TimestepMethod::VARIABLE:
        m_DT = this->decision_variable(1, samples);
        for (int i = 0; i < samples; ++i) m_DT[0, i].set_value(dt.count());
        break;
    }

    switch (transcription_method) {
      case TranscriptionMethod::DIRECT_TRANSCRIPTION:
        m_X = this->decision_variable(num_states, samples);
        constrain_direct_transcription();
        break;
      case TranscriptionMethod::DIRECT_COLLOCATION:
        m_X = this->decision_variable(num_states, samples);
        constrain_direct_collocation();
        break;
      case TranscriptionMethod::SINGLE_SHOOTING:
        // only X.col(0) stays free; the other columns become expressions
        m_X = Matrix{num_states, samples};
        constrain_single_shooting();
        break;
    }
  }

  /// ocp.hpp:181-189.
  template <typename T>
    requires ScalarLike<T> || MatrixLike<T>
  void constrain_initial_state(const T& initial_state) {
    this->subject_to(this->initial_state() == initial_state);
  }

  /// ocp.hpp:191-199.
  template <typename T>
    requires ScalarLike<T> || MatrixLike<T>
  void constrain_final_state(const T& final_state) {
    this->subject_to(this->final_state() == final_state);
  }

  /// Calls callback(x, u) with the state and input of every sample
  /// (ocp.hpp:201-216).
  void for_each_step(
      const std::function<void(const Matrix& x, const Matrix& u)>& callback) {
    for (int i = 0; i < m_num_steps + 1; ++i) {
      const Matrix x = X().col(i);
      const Matrix u = U().col(i);
      callback(x, u);
    }
  }

  /// Calls callback(t, x, u, dt) for every sample (ocp.hpp:218-240).
  void for_each_step(
      const std::function<void(const Variable<Scalar>& t, const Matrix& x,
                               const Matrix& u, const Variable<Scalar>& dt)>&
          callback) {
    Variable<Scalar> time{Scalar(0)};
    for (int i = 0; i < m_num_steps + 1; ++i) {
      const Matrix x = X().col(i);
      const Matrix u = U().col(i);
      const Variable<Scalar> step = this->dt()[0, i];
      callback(time, x, u, step);
      time += step;
    }
  }

  /// u ≥ lower_bound at every sample (ocp.hpp:242-254).
  template <typename T>
    requires ScalarLike<T> || MatrixLike<T>
  void set_lower_input_bound(const T& lower_bound) {
    for (int i = 0; i < m_num_steps + 1; ++i) {
      this->subject_to(U().col(i) >= lower_bound);
    }
  }

  /// u ≤ upper_bound at every sample (ocp.hpp:256-268).
  template <typename T>
    requires ScalarLike<T> || MatrixLike<T>
  void set_upper_input_bound(const T& upper_bound) {
    for (int i = 0; i < m_num_steps + 1; ++i) {
      this->subject_to(U().col(i) <= upper_bound);
    }
  }

  /// ocp.hpp:270-275.
  void set_min_timestep(std::chrono::duration<Scalar> min_timestep) {
    this->subject_to(dt() >= min_timestep.count());
  }

  /// ocp.hpp:277-282.
  void set_max_timestep(std::chrono::duration<Scalar> max_timestep) {
    this->subject_to(dt() <= max_timestep.count());
  }

  /// States, one column per sample.
  Matrix& X() { return m_X; }
  /// Inputs, one column per sample (the last column is unconstrained by the
  /// dynamics, as in the reference).
  Matrix& U() { return m_U; }
  /// Time steps, 1 × (num_steps + 1).
  Matrix& dt() { return m_DT; }
  Matrix initial_state() { return m_X.col(0); }
  Matrix final_state() { return m_X.col(m_num_steps); }

 private:
  int m_num_steps;
  TimedDynamics m_dynamics;
  DynamicsType m_dynamics_type;
  Matrix m_X, m_U, m_DT;

  /// Classic Runge–Kutta step of the dynamics (ocp.hpp:322-332): the stage
  /// order and the final combination are the reference's.
  Matrix rk4(const Matrix& x, const Matrix& u, const Variable<Scalar>& t0,
             const Variable<Scalar>& step) const {
    const auto& f = m_dynamics;
    const Variable<Scalar> half = step * Scalar(0.5);
    const Matrix k1 = f(t0, x, u, step);
    const Matrix k2 = f(t0 + half, x + k1 * half, u, step);
    const Matrix k3 = f(t0 + half, x + k2 * half, u, step);
    const Matrix k4 = f(t0 + step, x + k3 * step, u, step);
    return x + (k1 + k2 * Scalar(2) + k3 * Scalar(2) + k4) * (step / Scalar(6));
  }

  /// x(tₖ₊₁) as the transcription sees it: an RK4 step of an ODE, or the
  /// discrete map itself.
  Matrix advance(const Matrix& x, const Matrix& u, const Variable<Scalar>& t,
                 const Variable<Scalar>& step) const {
    if (m_dynamics_type == DynamicsType::EXPLICIT_ODE) {
      return rk4(x, u, t, step);
    }
    return m_dynamics(t, x, u, step);
  }

  /// Hermite–Simpson collocation (ocp.hpp:334-369): the derivative of the
  /// cubic through (xₖ, ẋₖ), (xₖ₊₁, ẋₖ₊₁) at its midpoint must equal the
  /// dynamics there.
  void constrain_direct_collocation() {
    slp_assert(m_dynamics_type == DynamicsType::EXPLICIT_ODE);
    const auto& f = m_dynamics;
    Variable<Scalar> time{Scalar(0)};
    for (int i = 0; i < m_num_steps; ++i) {
      const Variable<Scalar> h = dt()[0, i];
      const Variable<Scalar> t_begin = time;
      const Variable<Scalar> t_end = t_begin + h;
      const Matrix x_begin = X().col(i);
      const Matrix x_end = X().col(i + 1);
      const Matrix u_begin = U().col(i);
      const Matrix u_end = U().col(i + 1);

      const Matrix xdot_begin = f(t_begin, x_begin, u_begin, h);
      const Matrix xdot_end = f(t_end, x_end, u_end, h);
      const Matrix xdot_c = Scalar(-3) / (Scalar(2) * h) * (x_begin - x_end) -
                            Scalar(0.25) * (xdot_begin + xdot_end);

      const Variable<Scalar> t_c = t_begin + Scalar(0.5) * h;
      const Matrix x_c = Scalar(0.5) * (x_begin + x_end) +
                         h / Scalar(8) * (xdot_begin - xdot_end);
      const Matrix u_c = Scalar(0.5) * (u_begin + u_end);

      this->subject_to(xdot_c == f(t_c, x_c, u_c, h));
      time += h;
    }
  }

  /// ocp.hpp:371-392.
  void constrain_direct_transcription() {
    Variable<Scalar> time{Scalar(0)};
    for (int i = 0; i < m_num_steps; ++i) {
      const Matrix x_begin = X().col(i);
      const Matrix x_end = X().col(i + 1);
      const Matrix u = U().col(i);
      const Variable<Scalar> step = dt()[0, i];
      this->subject_to(x_end == advance(x_begin, u, time, step));
      time += step;
    }
  }

  /// ocp.hpp:394-414: X.col(k+1) is *defined* as the propagated X.col(k).
  void constrain_single_shooting() {
    Variable<Scalar> time{Scalar(0)};
    for (int i = 0; i < m_num_steps; ++i) {
      const Matrix x_begin = X().col(i);
      const Matrix u = U().col(i);
      const Variable<Scalar> step = dt()[0, i];
      X().col(i + 1) = advance(x_begin, u, time, step);
      time += step;
    }
  }
};

}  // namespace slp
