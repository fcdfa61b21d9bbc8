// slp_assert (reference: util/assert.hpp:5-24): throws when built for a
// language binding, plain assert otherwise.
#pragma once

#ifdef SLEIPNIR_PYTHON
#include <format>
#include <source_location>
#include <stdexcept>
#define slp_assert(condition)                                                 \
  do {                                                                        \
    if (!(condition)) {                                                       \
      auto location = std::source_location::current();                        \
      throw std::invalid_argument(std::format(                                \
          "{}:{}: {}: Assertion `{}' failed.", location.file_name(),          \
          location.line(), location.function_name(), #condition));            \
    }                                                                         \
  } while (0);
#else
#include <cassert>
#define slp_assert(condition) assert(condition)
#endif
