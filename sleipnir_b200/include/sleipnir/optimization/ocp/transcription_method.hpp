// optimization/ocp/transcription_method.hpp of the reference (:9-23).
#pragma once

#include <cstdint>

namespace slp {

/// How the dynamics become constraints.
enum class TranscriptionMethod : uint8_t {
  /// states are decision variables tied together by xₖ₊₁ = F(xₖ, uₖ)
  DIRECT_TRANSCRIPTION,
  /// states are decision variables; the dynamics are enforced at the
  /// midpoint of a cubic Hermite spline through consecutive states
  DIRECT_COLLOCATION,
  /// states are expressions of the initial state and the inputs
  SINGLE_SHOOTING
};

}  // namespace slp
