// Marker base for DSL types (reference: autodiff/sleipnir_base.hpp).
#pragma once

namespace slp {
class SleipnirBase {};
}  // namespace slp
