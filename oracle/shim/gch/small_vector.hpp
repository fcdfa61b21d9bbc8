// Build shim for the tier-A oracle (oracle/_ref): the reference's autodiff core
// only uses gch::small_vector as a growable array (expression.hpp:737,
// expression_graph.hpp:20,38, util/pool.hpp:89-90). gch::small_vector itself is
// an un-vendored dependency (CMakeLists.txt:91-101) absent from this image.
#pragma once
#include <vector>
namespace gch {
template <typename T>
using small_vector = std::vector<T>;
}  // namespace gch
