// optimization/ocp/dynamics_type.hpp of the reference (:9-16).
#pragma once

#include <cstdint>

namespace slp {

/// What the dynamics function of an OCP returns.
enum class DynamicsType : uint8_t {
  /// dx/dt = f(t, x, u)
  EXPLICIT_ODE,
  /// xₖ₊₁ = f(t, xₖ, uₖ)
  DISCRETE
};

}  // namespace slp
