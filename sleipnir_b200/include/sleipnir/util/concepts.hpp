// Concepts used to sort DSL operands (reference: util/concepts.hpp:13-38).
#pragma once

#include <concepts>
#include <type_traits>

#include "sleipnir/autodiff/sleipnir_base.hpp"

namespace slp {

template <typename T>
concept SleipnirType = std::derived_from<std::decay_t<T>, SleipnirBase>;

template <typename T>
concept MatrixLike = requires(std::decay_t<T> t) {
  t.rows();
  t.cols();
};

template <typename T>
concept ScalarLike =
    !MatrixLike<T> && std::constructible_from<std::decay_t<T>, int>;

template <typename T>
concept SleipnirMatrixLike = SleipnirType<T> && MatrixLike<T>;

template <typename T>
concept SleipnirScalarLike = SleipnirType<T> && ScalarLike<T>;

/// Plain numeric matrix operand (stands in for the reference's
/// EigenMatrixLike; Eigen is not a dependency of this build).
template <typename T>
concept NumericMatrixLike = MatrixLike<T> && !SleipnirType<T>;

}  // namespace slp
