// Sparsity-pattern recorder, file-compatible with the reference's
// (include/sleipnir/util/spy.hpp:20-100; tools/spy.py reads it):
//   header  : title, row label, column label (each: int32 length + bytes),
//             int32 rows, int32 cols
//   per add : int32 nnz, then nnz × (int32 row, int32 col, '+' | '-' | '0')
// all integers little-endian, coordinates in column-major (CSC) order.
#pragma once

#include <bit>
#include <cstdint>
#include <fstream>
#include <string>
#include <string_view>

#include "sleipnir/util/linalg.hpp"

namespace slp {

template <typename Scalar>
class Spy {
 public:
  Spy(std::string_view filename, std::string_view title,
      std::string_view row_label, std::string_view col_label, int rows,
      int cols)
      : m_file{std::string{filename}, std::ios::binary} {
    write_string(title);
    write_string(row_label);
    write_string(col_label);
    write_i32(rows);
    write_i32(cols);
  }

  /// Appends one frame: where the matrix has entries and their signs.
  void add(const SparseMatrix<Scalar>& mat) {
    write_i32(static_cast<int32_t>(mat.nonZeros()));
    const int32_t* outer = mat.outerIndexPtr();
    const int32_t* inner = mat.innerIndexPtr();
    const Scalar* values = mat.valuePtr();
    for (int col = 0; col < mat.cols(); ++col) {
      for (int32_t k = outer[col]; k < outer[col + 1]; ++k) {
        write_i32(inner[k]);
        write_i32(col);
        m_file.put(values[k] > Scalar(0) ? '+'
                                         : (values[k] < Scalar(0) ? '-' : '0'));
      }
    }
    m_file.flush();
  }

 private:
  std::ofstream m_file;

  void write_i32(int32_t v) {
    if constexpr (std::endian::native != std::endian::little) {
      v = static_cast<int32_t>(__builtin_bswap32(static_cast<uint32_t>(v)));
    }
    m_file.write(reinterpret_cast<const char*>(&v), sizeof(v));
  }
  void write_string(std::string_view s) {
    write_i32(static_cast<int32_t>(s.size()));
    m_file.write(s.data(), static_cast<std::streamsize>(s.size()));
  }
};

}  // namespace slp
