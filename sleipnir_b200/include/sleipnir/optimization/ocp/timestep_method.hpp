// optimization/ocp/timestep_method.hpp of the reference (:9-19).
#pragma once

#include <cstdint>

namespace slp {

/// How the time steps of an OCP are treated.
enum class TimestepMethod : uint8_t {
  /// every step is the constant dt
  FIXED,
  /// one decision variable shared by all steps
  VARIABLE_SINGLE,
  /// one decision variable per step
  VARIABLE
};

}  // namespace slp
