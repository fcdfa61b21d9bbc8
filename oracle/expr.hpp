// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under sleipnir_b200/ may include,
// link or execute this file. It exists to check the CUDA path.
//
// CPU restatement of the reference's autodiff expression core:
//   include/sleipnir/autodiff/expression.hpp      (node, op set, pruning rules)
//   include/sleipnir/autodiff/expression_graph.hpp (topological_sort :28-78,
//                                 update_values :85-96, append_triplets :106-153)
//
// It is an independent re-implementation (arena nodes + opcode switch instead of
// one virtual class per op). `tests/test_oracle_vs_reference.py` pins it against
// the reference's OWN expression.hpp/expression_graph.hpp (compiled verbatim
// into oracle/_ref by oracle/Makefile) on every config: bit-identical values
// and triplets are required.
#pragma once

#include <cmath>
#include <cstdint>
#include <memory>
#include <numbers>
#include <utility>
#include <vector>

namespace orc {

// Type lattice, expression_type.hpp:12-18.
enum class ExprType : uint8_t { NONE, CONSTANT, LINEAR, QUADRATIC, NONLINEAR };

enum class Op : uint8_t {
  CONST, VAR,
  SUB, ADD, DIV, MUL, NEG,
  ABS, ACOS, ASIN, ATAN, ATAN2, CBRT, COS, COSH, ERF, EXP, HYPOT,
  IS_NONNEG, IS_POS, LOG, LOG10, MAX, MIN, POW, SIGN, SIN, SINH, SQRT, TAN,
  TANH
};

struct Node;

/// Nullable handle to an arena-owned node. Mirrors the surface of the
/// reference's ExpressionPtr that the graph algorithms use.
struct ExprPtr {
  Node* p = nullptr;
  ExprPtr() = default;
  ExprPtr(std::nullptr_t) {}
  explicit ExprPtr(Node* n) : p{n} {}
  Node* operator->() const { return p; }
  Node& operator*() const { return *p; }
  Node* get() const { return p; }
  explicit operator bool() const { return p != nullptr; }
  friend bool operator==(const ExprPtr& a, std::nullptr_t) { return !a.p; }
  friend bool operator!=(const ExprPtr& a, std::nullptr_t) { return a.p; }
  friend bool operator==(const ExprPtr& a, const ExprPtr& b) {
    return a.p == b.p;
  }
};

ExprPtr constant_ptr(double value);

/// Scalar semantics of each op: value, and the adjoint-weighted partials
/// (expression.hpp line cited per case).
inline double op_value(Op op, double l, double r, double self) {
  switch (op) {
    case Op::CONST: case Op::VAR: return self;      // :583, :604
    case Op::SUB: return l - r;                     // :453
    case Op::ADD: return l + r;                     // :490
    case Op::DIV: return l / r;                     // :624
    case Op::MUL: return l * r;                     // :664
    case Op::NEG: return -l;                        // :703
    case Op::ABS: return std::abs(l);               // :780
    case Op::ACOS: return std::acos(l);             // :840
    case Op::ASIN: return std::asin(l);             // :894
    case Op::ATAN: return std::atan(l);             // :949
    case Op::ATAN2: return std::atan2(l, r);        // :1005
    case Op::CBRT: return std::cbrt(l);             // :524
    case Op::COS: return std::cos(l);               // :1065
    case Op::COSH: return std::cosh(l);             // :1119
    case Op::ERF: return std::erf(l);               // :1173
    case Op::EXP: return std::exp(l);               // :1229
    case Op::HYPOT: return std::hypot(l, r);        // :1289
    case Op::IS_NONNEG: return l >= 0.0 ? 1.0 : 0.0;  // :1358
    case Op::IS_POS: return l > 0.0 ? 1.0 : 0.0;      // :1391
    case Op::LOG: return std::log(l);               // :1424
    case Op::LOG10: return std::log10(l);           // :1476
    case Op::MAX: return std::max(l, r);            // :1533
    case Op::MIN: return std::min(l, r);            // :1602
    case Op::POW: return std::pow(l, r);            // :1675
    case Op::SIGN: return l < 0.0 ? -1.0 : (l == 0.0 ? 0.0 : 1.0);  // :1763
    case Op::SIN: return std::sin(l);               // :1812
    case Op::SINH: return std::sinh(l);             // :1867
    case Op::SQRT: return std::sqrt(l);             // :1922
    case Op::TAN: return std::tan(l);               // :1978
    case Op::TANH: return std::tanh(l);             // :2036
  }
  return 0.0;
}

inline double op_grad_l(Op op, double a, double l, double r) {
  switch (op) {
    case Op::SUB: return a;
    case Op::ADD: return a;
    case Op::DIV: return a / r;
    case Op::MUL: return a * r;
    case Op::NEG: return -a;
    case Op::ABS: return l < 0.0 ? -a : (l > 0.0 ? a : 0.0);
    case Op::ACOS: return -a / std::sqrt(1.0 - l * l);
    case Op::ASIN: return a / std::sqrt(1.0 - l * l);
    case Op::ATAN: return a / (1.0 + l * l);
    case Op::ATAN2: return a * r / (l * l + r * r);
    case Op::CBRT: { double c = std::cbrt(l); return a / (3.0 * c * c); }
    case Op::COS: return a * -std::sin(l);
    case Op::COSH: return a * std::sinh(l);
    case Op::ERF:
      return a * (2.0 * std::numbers::inv_sqrtpi) * std::exp(-l * l);
    case Op::EXP: return a * std::exp(l);
    case Op::HYPOT: return a * l / std::hypot(l, r);
    case Op::LOG: return a / l;
    case Op::LOG10: return a / (std::numbers::ln10 * l);
    case Op::MAX: return l >= r ? a : 0.0;
    case Op::MIN: return l <= r ? a : 0.0;
    case Op::POW: return a * std::pow(l, r - 1.0) * r;
    case Op::SIN: return a * std::cos(l);
    case Op::SINH: return a * std::cosh(l);
    case Op::SQRT: return a / (2.0 * std::sqrt(l));
    case Op::TAN: { double c = std::cos(l); return a / (c * c); }
    case Op::TANH: { double c = std::cosh(l); return a / (c * c); }
    default: return 0.0;  // const, var, sign, is_* : base default :383-386
  }
}

inline double op_grad_r(Op op, double a, double l, double r) {
  switch (op) {
    case Op::SUB: return -a;
    case Op::ADD: return a;
    case Op::DIV: return a * -l / (r * r);
    case Op::MUL: return a * l;
    case Op::ATAN2: return a * -l / (l * l + r * r);
    case Op::HYPOT: return a * r / std::hypot(l, r);
    case Op::MAX: return l >= r ? 0.0 : a;
    case Op::MIN: return l <= r ? 0.0 : a;
    case Op::POW: return a * std::pow(l, r) * std::log(l);
    default: return 0.0;
  }
}

struct Node {
  double val = 0.0;
  double adjoint = 0.0;
  ExprPtr adjoint_expr;
  ExprPtr args[2];
  int32_t scratch = -1;
  Op op = Op::CONST;
  ExprType ty = ExprType::CONSTANT;

  ExprType type() const { return ty; }
  bool is_constant(double c) const {
    return ty == ExprType::CONSTANT && val == c;
  }
  double value(double l, double r) const { return op_value(op, l, r, val); }
  double grad_l(double l, double r) const {
    return op_grad_l(op, adjoint, l, r);
  }
  double grad_r(double l, double r) const {
    return op_grad_r(op, adjoint, l, r);
  }
  ExprPtr grad_expr_l(const ExprPtr& l, const ExprPtr& r) const;
  ExprPtr grad_expr_r(const ExprPtr& l, const ExprPtr& r) const;
};

/// Chunked arena; nodes live until arena_reset(). (The reference uses a
/// ref-counted pool, src/util/pool.cpp:5-8; ownership is not part of the
/// arithmetic being checked.)
class Arena {
 public:
  Node* make() {
    if (m_used == kChunk) {
      m_chunks.emplace_back(new Node[kChunk]);
      m_used = 0;
    }
    ++m_total;
    return &m_chunks.back()[m_used++];
  }
  void reset() {
    m_chunks.clear();
    m_used = kChunk;
    m_total = 0;
  }
  size_t nodes() const { return m_total; }

 private:
  static constexpr size_t kChunk = 16384;
  std::vector<std::unique_ptr<Node[]>> m_chunks;
  size_t m_used = kChunk;
  size_t m_total = 0;
};

inline Arena& arena() {
  thread_local Arena a;
  return a;
}

inline ExprPtr make_node(Op op, ExprType ty, ExprPtr l = nullptr,
                         ExprPtr r = nullptr, double val = 0.0) {
  Node* n = arena().make();
  *n = Node{};
  n->op = op;
  n->ty = ty;
  n->args[0] = l;
  n->args[1] = r;
  n->val = val;
  return ExprPtr{n};
}

inline ExprPtr constant_ptr(double value) {
  return make_node(Op::CONST, ExprType::CONSTANT, nullptr, nullptr, value);
}
inline ExprPtr variable_ptr(double value = 0.0) {
  return make_node(Op::VAR, ExprType::LINEAR, nullptr, nullptr, value);
}

inline ExprPtr operator-(const ExprPtr& lhs);

// expression.hpp:155-201
inline ExprPtr operator*(const ExprPtr& lhs, const ExprPtr& rhs) {
  using enum ExprType;
  if (lhs->is_constant(0.0)) return lhs;
  if (rhs->is_constant(0.0)) return rhs;
  if (lhs->is_constant(1.0)) return rhs;
  if (rhs->is_constant(1.0)) return lhs;
  if (lhs->type() == CONSTANT && rhs->type() == CONSTANT) {
    return constant_ptr(lhs->val * rhs->val);
  }
  ExprType t;
  if (lhs->type() == CONSTANT) {
    t = rhs->type() == LINEAR ? LINEAR
        : rhs->type() == QUADRATIC ? QUADRATIC : NONLINEAR;
  } else if (rhs->type() == CONSTANT) {
    t = lhs->type() == LINEAR ? LINEAR
        : lhs->type() == QUADRATIC ? QUADRATIC : NONLINEAR;
  } else if (lhs->type() == LINEAR && rhs->type() == LINEAR) {
    t = QUADRATIC;
  } else {
    t = NONLINEAR;
  }
  return make_node(Op::MUL, t, lhs, rhs);
}

// expression.hpp:207-237
inline ExprPtr operator/(const ExprPtr& lhs, const ExprPtr& rhs) {
  using enum ExprType;
  if (lhs->is_constant(0.0)) return lhs;
  if (rhs->is_constant(1.0)) return lhs;
  if (lhs->type() == CONSTANT && rhs->type() == CONSTANT) {
    return constant_ptr(lhs->val / rhs->val);
  }
  ExprType t = NONLINEAR;
  if (rhs->type() == CONSTANT) {
    t = lhs->type() == LINEAR ? LINEAR
        : lhs->type() == QUADRATIC ? QUADRATIC : NONLINEAR;
  }
  return make_node(Op::DIV, t, lhs, rhs);
}

// expression.hpp:243-273 (null-tolerant: used in adjoint accumulation)
inline ExprPtr operator+(const ExprPtr& lhs, const ExprPtr& rhs) {
  using enum ExprType;
  if (lhs == nullptr || lhs->is_constant(0.0)) return rhs;
  if (rhs == nullptr || rhs->is_constant(0.0)) return lhs;
  if (lhs->type() == CONSTANT && rhs->type() == CONSTANT) {
    return constant_ptr(lhs->val + rhs->val);
  }
  ExprType t = std::max(lhs->type(), rhs->type());
  if (t != LINEAR && t != QUADRATIC) t = NONLINEAR;
  return make_node(Op::ADD, t, lhs, rhs);
}
inline ExprPtr& operator+=(ExprPtr& lhs, const ExprPtr& rhs) {
  return lhs = lhs + rhs;
}

// expression.hpp:288-322
inline ExprPtr operator-(const ExprPtr& lhs, const ExprPtr& rhs) {
  using enum ExprType;
  if (lhs->is_constant(0.0)) {
    if (rhs->is_constant(0.0)) return rhs;
    return -rhs;
  }
  if (rhs->is_constant(0.0)) return lhs;
  if (lhs->type() == CONSTANT && rhs->type() == CONSTANT) {
    return constant_ptr(lhs->val - rhs->val);
  }
  ExprType t = std::max(lhs->type(), rhs->type());
  if (t != LINEAR && t != QUADRATIC) t = NONLINEAR;
  return make_node(Op::SUB, t, lhs, rhs);
}

// expression.hpp:327-348
inline ExprPtr operator-(const ExprPtr& lhs) {
  using enum ExprType;
  if (lhs->is_constant(0.0)) return lhs;
  if (lhs->type() == CONSTANT) return constant_ptr(-lhs->val);
  ExprType t = lhs->type() == LINEAR ? LINEAR
               : lhs->type() == QUADRATIC ? QUADRATIC : NONLINEAR;
  return make_node(Op::NEG, t, lhs);
}
inline ExprPtr operator+(const ExprPtr& lhs) { return lhs; }

namespace detail {
// Shared shape of the unary factories: "zero stays the same node", constant
// folding, else a NONLINEAR node.
template <typename F>
ExprPtr unary_zero_fixed(Op op, const ExprPtr& x, F f) {
  if (x->is_constant(0.0)) return x;
  if (x->type() == ExprType::CONSTANT) return constant_ptr(f(x->val));
  return make_node(op, ExprType::NONLINEAR, x);
}
}  // namespace detail

inline ExprPtr abs(const ExprPtr& x) {  // :800-815
  return detail::unary_zero_fixed(Op::ABS, x,
                                  [](double v) { return std::abs(v); });
}
inline ExprPtr acos(const ExprPtr& x) {  // :855-870
  if (x->is_constant(0.0)) return constant_ptr(std::numbers::pi / 2.0);
  if (x->type() == ExprType::CONSTANT) return constant_ptr(std::acos(x->val));
  return make_node(Op::ACOS, ExprType::NONLINEAR, x);
}
inline ExprPtr asin(const ExprPtr& x) {  // :909-925
  return detail::unary_zero_fixed(Op::ASIN, x,
                                  [](double v) { return std::asin(v); });
}
inline ExprPtr atan(const ExprPtr& x) {  // :963-979
  return detail::unary_zero_fixed(Op::ATAN, x,
                                  [](double v) { return std::atan(v); });
}
inline ExprPtr atan2(const ExprPtr& y, const ExprPtr& x) {  // :1023-1035
  if (y->type() == ExprType::CONSTANT && x->type() == ExprType::CONSTANT) {
    return constant_ptr(std::atan2(y->val, x->val));
  }
  return make_node(Op::ATAN2, ExprType::NONLINEAR, y, x);
}
inline ExprPtr cbrt(const ExprPtr& x) {  // :550-568
  if (x->type() == ExprType::CONSTANT) {
    if (x->val == 0.0) return x;
    if (x->val == -1.0 || x->val == 1.0) return x;
    return constant_ptr(std::cbrt(x->val));
  }
  return make_node(Op::CBRT, ExprType::NONLINEAR, x);
}
inline ExprPtr cos(const ExprPtr& x) {  // :1080-1095
  if (x->is_constant(0.0)) return constant_ptr(1.0);
  if (x->type() == ExprType::CONSTANT) return constant_ptr(std::cos(x->val));
  return make_node(Op::COS, ExprType::NONLINEAR, x);
}
inline ExprPtr cosh(const ExprPtr& x) {  // :1134-1149
  if (x->is_constant(0.0)) return constant_ptr(1.0);
  if (x->type() == ExprType::CONSTANT) return constant_ptr(std::cosh(x->val));
  return make_node(Op::COSH, ExprType::NONLINEAR, x);
}
inline ExprPtr erf(const ExprPtr& x) {  // :1189-1205
  return detail::unary_zero_fixed(Op::ERF, x,
                                  [](double v) { return std::erf(v); });
}
inline ExprPtr exp(const ExprPtr& x) {  // :1244-1259
  if (x->is_constant(0.0)) return constant_ptr(1.0);
  if (x->type() == ExprType::CONSTANT) return constant_ptr(std::exp(x->val));
  return make_node(Op::EXP, ExprType::NONLINEAR, x);
}
inline ExprPtr hypot(const ExprPtr& x, const ExprPtr& y) {  // :1309-1327
  if (x->is_constant(0.0)) return abs(y);
  if (y->is_constant(0.0)) return abs(x);
  if (x->type() == ExprType::CONSTANT && y->type() == ExprType::CONSTANT) {
    return constant_ptr(std::hypot(x->val, y->val));
  }
  return make_node(Op::HYPOT, ExprType::NONLINEAR, x, y);
}
inline ExprPtr is_nonnegative(const ExprPtr& x) {  // :1367-1373
  if (x->type() == ExprType::CONSTANT) {
    return constant_ptr(x->val >= 0.0 ? 1.0 : 0.0);
  }
  return make_node(Op::IS_NONNEG, ExprType::NONLINEAR, x);
}
inline ExprPtr is_positive(const ExprPtr& x) {  // :1400-1406
  if (x->type() == ExprType::CONSTANT) {
    return constant_ptr(x->val > 0.0 ? 1.0 : 0.0);
  }
  return make_node(Op::IS_POS, ExprType::NONLINEAR, x);
}
inline ExprPtr log(const ExprPtr& x) {  // :1436-1452
  return detail::unary_zero_fixed(Op::LOG, x,
                                  [](double v) { return std::log(v); });
}
inline ExprPtr log10(const ExprPtr& x) {  // :1490-1506
  return detail::unary_zero_fixed(Op::LOG10, x,
                                  [](double v) { return std::log10(v); });
}
inline ExprPtr max(const ExprPtr& a, const ExprPtr& b) {  // :1571-1583
  if (a->type() == ExprType::CONSTANT && b->type() == ExprType::CONSTANT) {
    return constant_ptr(std::max(a->val, b->val));
  }
  return make_node(Op::MAX, ExprType::NONLINEAR, a, b);
}
inline ExprPtr min(const ExprPtr& a, const ExprPtr& b) {  // :1640-1652
  if (a->type() == ExprType::CONSTANT && b->type() == ExprType::CONSTANT) {
    return constant_ptr(std::min(a->val, b->val));
  }
  return make_node(Op::MIN, ExprType::NONLINEAR, a, b);
}
inline ExprPtr pow(const ExprPtr& base, const ExprPtr& power) {  // :1712-1752
  using enum ExprType;
  if (base->is_constant(0.0)) return base;
  if (base->is_constant(1.0)) return base;
  if (power->is_constant(0.0)) return constant_ptr(1.0);
  if (power->is_constant(1.0)) return base;
  if (base->type() == CONSTANT && power->type() == CONSTANT) {
    return constant_ptr(std::pow(base->val, power->val));
  }
  if (power->is_constant(2.0)) {
    return make_node(Op::MUL, base->type() == LINEAR ? QUADRATIC : NONLINEAR,
                     base, base);
  }
  return make_node(Op::POW, NONLINEAR, base, power);
}
inline ExprPtr sign(const ExprPtr& x) {  // :1774-1791
  if (x->type() == ExprType::CONSTANT) {
    if (x->val < 0.0) return constant_ptr(-1.0);
    if (x->val == 0.0) return x;
    return constant_ptr(1.0);
  }
  return make_node(Op::SIGN, ExprType::NONLINEAR, x);
}
inline ExprPtr sin(const ExprPtr& x) {  // :1827-1843
  return detail::unary_zero_fixed(Op::SIN, x,
                                  [](double v) { return std::sin(v); });
}
inline ExprPtr sinh(const ExprPtr& x) {  // :1882-1898
  return detail::unary_zero_fixed(Op::SINH, x,
                                  [](double v) { return std::sinh(v); });
}
inline ExprPtr sqrt(const ExprPtr& x) {  // :1937-1954
  if (x->type() == ExprType::CONSTANT) {
    if (x->val == 0.0) return x;
    if (x->val == 1.0) return x;
    return constant_ptr(std::sqrt(x->val));
  }
  return make_node(Op::SQRT, ExprType::NONLINEAR, x);
}
inline ExprPtr tan(const ExprPtr& x) {  // :1995-2011
  return detail::unary_zero_fixed(Op::TAN, x,
                                  [](double v) { return std::tan(v); });
}
inline ExprPtr tanh(const ExprPtr& x) {  // :2053-2069
  return detail::unary_zero_fixed(Op::TANH, x,
                                  [](double v) { return std::tanh(v); });
}

// Symbolic partials (grad_expr_l / grad_expr_r of each op). `ae` is the
// node's adjoint expression.
inline ExprPtr Node::grad_expr_l(const ExprPtr& l, const ExprPtr& r) const {
  const ExprPtr& ae = adjoint_expr;
  switch (op) {
    case Op::SUB: return ae;                                   // :463-467
    case Op::ADD: return ae;                                   // :500-504
    case Op::DIV: return ae / r;                               // :638-642
    case Op::MUL: return ae * r;                               // :678-682
    case Op::NEG: return -ae;                                  // :711-715
    case Op::ABS: return ae * sign(l);                         // :799
    case Op::ACOS: return -ae / sqrt(constant_ptr(1.0) - l * l);  // :854
    case Op::ASIN: return ae / sqrt(constant_ptr(1.0) - l * l);   // :908
    case Op::ATAN: return ae / (constant_ptr(1.0) + l * l);       // :962
    case Op::ATAN2: return ae * r / (l * l + r * r);              // :1013
    case Op::CBRT: {                                              // :540-546
      auto c = cbrt(l);
      return ae / (constant_ptr(3.0) * c * c);
    }
    case Op::COS: return ae * -sin(l);                            // :1079
    case Op::COSH: return ae * sinh(l);                           // :1133
    case Op::ERF:                                                 // :1187
      return ae * constant_ptr(2.0 * std::numbers::inv_sqrtpi) * exp(-l * l);
    case Op::EXP: return ae * exp(l);                             // :1243
    case Op::HYPOT: return ae * l / hypot(l, r);                  // :1299
    case Op::LOG: return ae / l;                                  // :1435
    case Op::LOG10: return ae / (constant_ptr(std::numbers::ln10) * l);
    case Op::MAX: return ae * is_nonnegative(l - r);              // :1554
    case Op::MIN: return ae * is_nonnegative(r - l);              // :1624
    case Op::POW: return ae * pow(l, r - constant_ptr(1.0)) * r;  // :1696
    case Op::SIN: return ae * cos(l);                             // :1826
    case Op::SINH: return ae * cosh(l);                           // :1881
    case Op::SQRT: return ae / (constant_ptr(2.0) * sqrt(l));     // :1936
    case Op::TAN: { auto c = cos(l); return ae / (c * c); }       // :1994
    case Op::TANH: { auto c = cosh(l); return ae / (c * c); }     // :2052
    default: return constant_ptr(0.0);                            // :403-407
  }
}

inline ExprPtr Node::grad_expr_r(const ExprPtr& l, const ExprPtr& r) const {
  const ExprPtr& ae = adjoint_expr;
  switch (op) {
    case Op::SUB: return -ae;
    case Op::ADD: return ae;
    case Op::DIV: return ae * -l / (r * r);
    case Op::MUL: return ae * l;
    case Op::ATAN2: return ae * -l / (l * l + r * r);
    case Op::HYPOT: return ae * r / hypot(l, r);
    case Op::MAX: return ae * is_positive(r - l);
    case Op::MIN: return ae * is_positive(l - r);
    case Op::POW: return ae * pow(l, r) * log(l);
    default: return constant_ptr(0.0);
  }
}

using ExprGraph = std::vector<Node*>;

struct Triplet {
  int r = 0, c = 0;
  double v = 0.0;
  Triplet() = default;
  Triplet(int row, int col, double value) : r{row}, c{col}, v{value} {}
  int row() const { return r; }
  int col() const { return c; }
  double value() const { return v; }
};

/// Parent→child ordering of the sub-graph under `root`
/// (expression_graph.hpp:28-78). `scratch` is the in-degree counter, offset −1.
inline ExprGraph topological_sort(const ExprPtr& root) {
  ExprGraph list;
  if (root == nullptr || root->type() == ExprType::CONSTANT) return list;

  std::vector<Node*> stack;
  stack.push_back(root.get());
  while (!stack.empty()) {
    Node* node = stack.back();
    stack.pop_back();
    for (auto& arg : node->args) {
      if (arg != nullptr && ++arg->scratch == 0) stack.push_back(arg.get());
    }
  }
  stack.push_back(root.get());
  while (!stack.empty()) {
    Node* node = stack.back();
    stack.pop_back();
    list.push_back(node);
    for (auto& arg : node->args) {
      if (arg != nullptr && --arg->scratch == -1) stack.push_back(arg.get());
    }
  }
  return list;
}

/// Forward sweep child→parent (expression_graph.hpp:85-96).
inline void update_values(const ExprGraph& list) {
  for (auto it = list.rbegin(); it != list.rend(); ++it) {
    Node* node = *it;
    auto& lhs = node->args[0];
    auto& rhs = node->args[1];
    if (lhs != nullptr) {
      node->val = node->value(lhs->val, rhs ? rhs->val : 0.0);
    }
  }
}

/// One reverse sweep = one Jacobian row (expression_graph.hpp:106-153).
inline void append_triplets(const ExprGraph& top_list,
                            const std::vector<std::pair<int, Node*>>& outputs,
                            std::vector<Triplet>& triplets, int row) {
  if (top_list.empty()) return;
  top_list[0]->adjoint = 1.0;
  for (size_t i = 1; i < top_list.size(); ++i) top_list[i]->adjoint = 0.0;
  for (Node* node : top_list) {
    auto& lhs = node->args[0];
    auto& rhs = node->args[1];
    if (lhs != nullptr) {
      if (rhs != nullptr) {
        lhs->adjoint += node->grad_l(lhs->val, rhs->val);
        rhs->adjoint += node->grad_r(lhs->val, rhs->val);
      } else {
        lhs->adjoint += node->grad_l(lhs->val, 0.0);
      }
    }
  }
  for (const auto& [col, node] : outputs) {
    triplets.emplace_back(row, col, node->adjoint);
  }
}

/// Backend bundle used by the templated layers above (var.hpp, autodiff.hpp).
struct OwnBackend {
  using Ptr = ExprPtr;
  using Raw = Node;
  using Graph = ExprGraph;
  using Trip = Triplet;
  using Type = ExprType;
  static Ptr make_var() { return variable_ptr(); }
  static Ptr make_const(double v) { return constant_ptr(v); }
  static Graph sort(const Ptr& root) { return topological_sort(root); }
  static void update(const Graph& g) { update_values(g); }
  static void triplets(const Graph& g,
                       const std::vector<std::pair<int, Raw*>>& outs,
                       std::vector<Trip>& t, int row) {
    append_triplets(g, outs, t, row);
  }
  static int type_rank(const Ptr& p) { return static_cast<int>(p->type()); }
  static const char* name() { return "restated"; }
};

}  // namespace orc
