// Solver diagnostics on stdout when Options::diagnostics is set: the
// per-iteration table of the reference (include/sleipnir/util/
// print_diagnostics.hpp:193-249 — same columns, widths and iteration-type
// letters, so logs of both solvers read side by side), the exit status, and a
// table of where the time went. The timing table is this library's own: host
// phases of Problem::solve() and device time per kernel group, not the
// reference's host profilers.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <format>
#include <string>
#include <string_view>
#include <vector>

namespace slp {

enum class IterationType : uint8_t {
  NORMAL,
  SECOND_ORDER_CORRECTION,
  FEASIBILITY_RESTORATION
};

namespace detail {

inline void put_line(const std::string& line) {
  std::fputs(line.c_str(), stdout);
  std::fputc('\n', stdout);
}

inline std::string repeat(std::string_view piece, int count) {
  std::string out;
  for (int i = 0; i < count; ++i) out += piece;
  return out;
}

/// 10ⁿ with superscript digits, " 0" for zero (print_diagnostics.hpp:49-92).
inline std::string power_of_10(double value) {
  if (value == 0.0) return " 0";
  const int exponent = static_cast<int>(std::log10(value));
  if (exponent == 0) return " 1";
  if (exponent == 1) return "10";
  static constexpr std::array<std::string_view, 10> kSuper{
      "⁰", "¹", "²", "³", "⁴", "⁵", "⁶", "⁷", "⁸", "⁹"};
  std::string digits;
  for (int n = std::abs(exponent); n > 0; n /= 10) {
    digits.insert(0, kSuper[n % 10]);
  }
  return std::string{"10"} + (exponent < 0 ? "⁻" : "") + digits;
}

/// Pads a string holding multi-byte characters to a display width.
inline std::string pad_right(std::string text, int display_width, int width) {
  if (display_width < width) text.append(width - display_width, ' ');
  return text;
}

inline int display_width(std::string_view utf8) {
  int w = 0;
  for (unsigned char c : utf8) w += (c & 0xC0) != 0x80;
  return w;
}

}  // namespace detail

/// One row of the iteration table; a header every 20 rows.
inline void print_iteration_diagnostics(
    int iterations, IterationType type, double milliseconds, double error,
    double cost, double infeasibility, double complementarity, double mu,
    double delta, double gamma, double full_primal_step_inf_norm,
    double full_dual_step_inf_norm, double primal_alpha,
    double primal_alpha_max, double alpha_reduction_factor,
    double dual_alpha) {
  using detail::put_line;
  if (iterations % 20 == 0) {
    put_line((iterations == 0 ? "┏" : "┢") + detail::repeat("━", 119) +
             (iterations == 0 ? "┓" : "┪"));
    // column titles with non-ASCII letters are padded by display width
    auto mid = [](std::string_view s, int width) {
      const int w = detail::display_width(s);
      const int left = (width - w) / 2;
      return std::string(left, ' ') + std::string{s} +
             std::string(width - w - left, ' ');
    };
    put_line("┃" + mid("iter", 4) + "   " + mid("duration", 9) + " " +
             mid("error", 10) + " " + mid("cost", 11) + " " +
             mid("infeas.", 10) + " " + mid("complem.", 8) + " " +
             mid("μ", 8) + " " + mid("δ", 5) + " " + mid("γ", 5) + " " +
             mid("|p_pr|", 8) + " " + mid("|p_du|", 8) + " " +
             mid("α_pr", 8) + " " + mid("α_du", 8) + " " + mid("↩", 2) + "┃");
    put_line("┡" + detail::repeat("━", 119) + "┩");
  }
  // number of backtracks x with α_max·rˣ = α
  const int backtracks = static_cast<int>(
      std::log(primal_alpha / primal_alpha_max) /
      std::log(alpha_reduction_factor));
  constexpr std::array<const char*, 3> kTypes{" ", "s", "r"};
  const std::string d10 = detail::power_of_10(delta);
  const std::string g10 = detail::power_of_10(gamma);
  put_line(std::format(
      "│{:4} {:1} {:9.3f} {:10.4e} {:11.4e} {:10.4e} {:8.2e} {:8.2e} {} {} "
      "{:8.2e} {:8.2e} {:8.2e} {:8.2e} {:2d}│",
      iterations, kTypes[static_cast<int>(type)], milliseconds, error, cost,
      infeasibility, complementarity, mu,
      detail::pad_right(d10, detail::display_width(d10), 5),
      detail::pad_right(g10, detail::display_width(g10), 5),
      full_primal_step_inf_norm, full_dual_step_inf_norm, primal_alpha,
      dual_alpha, backtracks));
}

inline void print_bottom_iteration_diagnostics() {
  detail::put_line("└" + detail::repeat("─", 119) + "┘");
}

/// "name  share  total ms  count  ms each" rows, heaviest first as given.
struct TimingRow {
  std::string name;
  double total_ms = 0.0;
  int64_t count = 0;
};

inline void print_timing_table(std::string_view title,
                               const std::vector<TimingRow>& rows) {
  using detail::put_line;
  double sum = 0.0;
  for (const auto& r : rows) sum += r.total_ms;
  put_line(std::format("┏{}┓", detail::repeat("━", 70)));
  put_line(std::format("┃{:^70}┃", title));
  put_line(std::format("┃{:<28} {:>7} {:>12} {:>8} {:>11} ┃", "phase", "share",
                       "total (ms)", "count", "each (ms)"));
  put_line(std::format("┡{}┩", detail::repeat("━", 70)));
  for (const auto& r : rows) {
    put_line(std::format(
        "│{:<28} {:>6.1f}% {:>12.3f} {:>8} {:>11.4f} │", r.name,
        sum > 0.0 ? 100.0 * r.total_ms / sum : 0.0, r.total_ms, r.count,
        r.count > 0 ? r.total_ms / double(r.count) : 0.0));
  }
  put_line(std::format("└{}┘", detail::repeat("─", 70)));
}

}  // namespace slp
