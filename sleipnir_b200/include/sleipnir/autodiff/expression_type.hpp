// Expression type lattice (reference: autodiff/expression_type.hpp:12-18).
#pragma once

#include <cstdint>
#include <string_view>

namespace slp {

/// NONE < CONSTANT < LINEAR < QUADRATIC < NONLINEAR.
enum class ExpressionType : uint8_t {
  NONE,
  CONSTANT,
  LINEAR,
  QUADRATIC,
  NONLINEAR
};

constexpr std::string_view to_string(ExpressionType t) {
  using enum ExpressionType;
  switch (t) {
    case NONE: return "none";
    case CONSTANT: return "constant";
    case LINEAR: return "linear";
    case QUADRATIC: return "quadratic";
    case NONLINEAR: return "nonlinear";
  }
  return "";
}

}  // namespace slp
