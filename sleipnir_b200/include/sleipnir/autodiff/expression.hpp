// B200-native Sleipnir-compatible interior-point path — host expression graph.
//
// Mirrors the behaviour of the reference's
// include/sleipnir/autodiff/expression.hpp (node semantics, constant folding /
// pruning rules :155-348, type lattice, per-op value / grad / grad_expr) and
// expression_graph.hpp (topological_sort :28-78, update_values :85-96,
// append_triplets :106-153), but not its representation: nodes live in a flat
// structure-of-arrays pool addressed by 32-bit ids, which is already the
// layout the device tape is uploaded in (include/slpb.h, slpb_upload_tape).
// There is one opcode per reference Expression subclass instead of a virtual
// class hierarchy, and no per-node reference count: regions of the pool are
// reclaimed when the PoolScope that created them closes with no handle left
// into them, and the whole pool when the last handle dies. A handle is a bare
// id resolved against the CALLING thread's pool: Variables, VariableMatrices
// and Problems must stay on the thread that created them (Problem::solve
// checks this); the reference's pointers, by contrast, stay valid across
// threads.
#pragma once

#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdint>
#include <limits>
#include <mutex>
#include <numbers>
#include <stdexcept>
#include <utility>
#include <vector>

#include "sleipnir/autodiff/expression_type.hpp"

namespace slp::detail {

using ExprId = int32_t;
inline constexpr ExprId kNull = -1;

/// Numbering is shared with enum slpb_op in include/slpb.h.
enum class Op : uint8_t {
  CONST = 0, VAR, SUB, ADD, DIV, MUL, NEG, ABS, ACOS, ASIN, ATAN, ATAN2, CBRT,
  COS, COSH, ERF, EXP, HYPOT, IS_NONNEG, IS_POS, LOG, LOG10, MAX, MIN, POW,
  SIGN, SIN, SINH, SQRT, TAN, TANH
};

/// Flat node storage. One pool per thread (the reference keeps a thread_local
/// pool resource, src/util/pool.cpp:5-8, so that multistart can build problems
/// on several threads).
class ExpressionPool {
 public:
  std::vector<uint8_t> op;
  std::vector<uint8_t> type;
  std::vector<ExprId> lhs;
  std::vector<ExprId> rhs;
  std::vector<double> val;
  /// Scratch for graph algorithms; −1 when idle (same contract as
  /// Expression::scratch, expression.hpp:106-115).
  std::vector<int32_t> scratch;

  ExprId make(Op o, ExpressionType t, ExprId l, ExprId r, double v) {
    if (op.size() >= static_cast<size_t>(std::numeric_limits<ExprId>::max())) {
      throw std::length_error("expression pool: more than 2^31 - 1 nodes");
    }
    op.push_back(static_cast<uint8_t>(o));
    type.push_back(static_cast<uint8_t>(t));
    lhs.push_back(l);
    rhs.push_back(r);
    val.push_back(v);
    scratch.push_back(-1);
    return static_cast<ExprId>(op.size() - 1);
  }
  size_t size() const { return op.size(); }

  /// Handles are counted in total and, while a PoolScope is open, separately
  /// for nodes created inside it (id ≥ mark): when the scope closes and none of
  /// those is referenced any more, the pool shrinks back to the mark. The
  /// reference frees unreferenced nodes one by one through intrusive reference
  /// counts (util/intrusive_shared_ptr.hpp); here whole regions are reclaimed —
  /// the gradient trees, the Lagrangian and the restoration problem that every
  /// Problem::solve() builds and drops.
  void acquire(ExprId id) {
    ++m_live;
    if (static_cast<size_t>(id) >= m_mark) ++m_live_above;
  }
  void release(ExprId id) {
    if (static_cast<size_t>(id) >= m_mark) --m_live_above;
    if (--m_live == 0) clear();
  }
  /// Number of live handles; the analogue of
  /// global_pool_resource().blocks_in_use() (util/pool.hpp:84-86): zero means
  /// every node has been returned.
  int64_t handles_in_use() const { return m_live; }
  size_t nodes_in_use() const { return op.size(); }

 private:
  friend class PoolScope;
  void clear() {
    op = {};
    type = {};
    lhs = {};
    rhs = {};
    val = {};
    scratch = {};
  }
  void truncate(size_t n) {
    if (n >= op.size()) return;
    op.resize(n);
    type.resize(n);
    lhs.resize(n);
    rhs.resize(n);
    val.resize(n);
    scratch.resize(n);
  }
  int64_t m_live = 0;
  /// Nodes with id ≥ m_mark belong to the innermost open PoolScope.
  size_t m_mark = std::numeric_limits<size_t>::max();
  int64_t m_live_above = 0;
};

/// Reclaims the nodes created while it is open (Problem::solve wraps itself in
/// one): on destruction, if no handle refers to a node created inside the
/// scope, the pool is truncated back to its size at entry. Scopes nest
/// (feasibility restoration builds its own problem inside the outer solve).
class PoolScope {
 public:
  explicit PoolScope(ExpressionPool& pool)
      : m_pool{pool}, m_saved_mark{pool.m_mark},
        m_saved_live_above{pool.m_live_above} {
    m_pool.m_mark = m_pool.size();
    m_pool.m_live_above = 0;
  }
  PoolScope(const PoolScope&) = delete;
  PoolScope& operator=(const PoolScope&) = delete;
  ~PoolScope() {
    const int64_t survivors = m_pool.m_live_above;
    if (survivors == 0) m_pool.truncate(m_pool.m_mark);
    // whatever survives is above the enclosing scope's mark as well
    m_pool.m_mark = m_saved_mark;
    m_pool.m_live_above =
        m_saved_mark == std::numeric_limits<size_t>::max()
            ? 0
            : m_saved_live_above + survivors;
  }

 private:
  ExpressionPool& m_pool;
  size_t m_saved_mark;
  int64_t m_saved_live_above;
};

namespace pool_detail {

/// In a shared library every access to a thread_local goes through the dynamic
/// loader (__tls_get_addr), and the pool is reached on every handle copy —
/// a third of the graph-construction time. The thread that builds the graphs
/// is almost always one and the same, so its pool is also published in a
/// process-wide slot keyed by the thread pointer (one register read): that
/// thread skips the TLS machinery, every other thread takes the ordinary path.
struct FastSlot {
  std::atomic<void*> owner{nullptr};  // thread pointer of the owning thread
  ExpressionPool* pool = nullptr;
  std::mutex claim;  // serialises claiming and giving back
};
inline FastSlot& fast_slot() {
  static FastSlot slot;
  return slot;
}
inline void* thread_key() {
#if defined(__GNUC__) && (defined(__x86_64__) || defined(__aarch64__))
  return __builtin_thread_pointer();
#else
  return nullptr;  // no fast path: always the thread_local
#endif
}

/// The thread's pool; claims the fast slot if nobody holds it and gives it
/// back when the thread ends.
struct ThreadPool {
  ExpressionPool pool;
  bool owns_slot = false;
  ThreadPool() {
    void* me = thread_key();
    void* expected = nullptr;
    FastSlot& slot = fast_slot();
    if (me != nullptr) {
      // publish the pool before the owner becomes visible
      if (slot.owner.load(std::memory_order_acquire) == nullptr) {
        std::lock_guard<std::mutex> lock{slot.claim};
        if (slot.owner.load(std::memory_order_relaxed) == expected) {
          slot.pool = &pool;
          slot.owner.store(me, std::memory_order_release);
          owns_slot = true;
        }
      }
    }
  }
  ~ThreadPool() {
    if (owns_slot) {
      // Under the claim mutex, and the pool pointer first: a thread that
      // claims the slot right after the owner is cleared must not see its
      // own `pool` overwritten by this teardown.
      FastSlot& slot = fast_slot();
      std::lock_guard<std::mutex> lock{slot.claim};
      slot.pool = nullptr;
      slot.owner.store(nullptr, std::memory_order_release);
    }
  }
};

inline ExpressionPool& slow_path() {
  thread_local ThreadPool p;
  return p.pool;
}

}  // namespace pool_detail

inline ExpressionPool& pool() {
  pool_detail::FastSlot& slot = pool_detail::fast_slot();
  void* me = pool_detail::thread_key();
  if (me != nullptr && slot.owner.load(std::memory_order_acquire) == me) {
    return *slot.pool;
  }
  return pool_detail::slow_path();
}

/// Nullable owning handle (plays the role of ExpressionPtr).
class Expr {
 public:
  Expr() = default;
  Expr(std::nullptr_t) {}  // NOLINT
  explicit Expr(ExprId i) : m_id{i} {
    if (m_id != kNull) pool().acquire(m_id);
  }
  Expr(const Expr& o) : m_id{o.m_id} {
    if (m_id != kNull) pool().acquire(m_id);
  }
  Expr(Expr&& o) noexcept : m_id{o.m_id} { o.m_id = kNull; }
  Expr& operator=(const Expr& o) {
    if (this != &o) {
      Expr tmp{o};
      std::swap(m_id, tmp.m_id);
    }
    return *this;
  }
  Expr& operator=(Expr&& o) noexcept {
    std::swap(m_id, o.m_id);
    return *this;
  }
  ~Expr() {
    if (m_id != kNull) pool().release(m_id);
  }

  ExprId id() const { return m_id; }
  explicit operator bool() const { return m_id != kNull; }
  friend bool operator==(const Expr& a, std::nullptr_t) {
    return a.m_id == kNull;
  }
  friend bool operator==(const Expr& a, const Expr& b) {
    return a.m_id == b.m_id;
  }

  ExpressionType type() const {
    return static_cast<ExpressionType>(pool().type[m_id]);
  }
  Op op() const { return static_cast<Op>(pool().op[m_id]); }
  double val() const { return pool().val[m_id]; }
  void set_val(double v) const { pool().val[m_id] = v; }
  bool has_args() const { return pool().lhs[m_id] != kNull; }
  bool is_constant(double c) const {
    return type() == ExpressionType::CONSTANT && val() == c;
  }

 private:
  ExprId m_id = kNull;
};

inline Expr make_expr(Op o, ExpressionType t, const Expr& l = nullptr,
                      const Expr& r = nullptr, double v = 0.0) {
  return Expr{pool().make(o, t, l.id(), r.id(), v)};
}
inline Expr constant_ptr(double v) {
  return make_expr(Op::CONST, ExpressionType::CONSTANT, nullptr, nullptr, v);
}
inline Expr decision_variable_ptr(double v = 0.0) {
  return make_expr(Op::VAR, ExpressionType::LINEAR, nullptr, nullptr, v);
}

// ---- scalar semantics of every op ----------------------

inline double op_value(Op op, double l, double r) {
  switch (op) {
    case Op::SUB: return l - r;
    case Op::ADD: return l + r;
    case Op::DIV: return l / r;
    case Op::MUL: return l * r;
    case Op::NEG: return -l;
    case Op::ABS: return std::abs(l);
    case Op::ACOS: return std::acos(l);
    case Op::ASIN: return std::asin(l);
    case Op::ATAN: return std::atan(l);
    case Op::ATAN2: return std::atan2(l, r);
    case Op::CBRT: return std::cbrt(l);
    case Op::COS: return std::cos(l);
    case Op::COSH: return std::cosh(l);
    case Op::ERF: return std::erf(l);
    case Op::EXP: return std::exp(l);
    case Op::HYPOT: return std::hypot(l, r);
    case Op::IS_NONNEG: return l >= 0.0 ? 1.0 : 0.0;
    case Op::IS_POS: return l > 0.0 ? 1.0 : 0.0;
    case Op::LOG: return std::log(l);
    case Op::LOG10: return std::log10(l);
    case Op::MAX: return std::max(l, r);
    case Op::MIN: return std::min(l, r);
    case Op::POW: return std::pow(l, r);
    case Op::SIGN: return l < 0.0 ? -1.0 : (l == 0.0 ? 0.0 : 1.0);
    case Op::SIN: return std::sin(l);
    case Op::SINH: return std::sinh(l);
    case Op::SQRT: return std::sqrt(l);
    case Op::TAN: return std::tan(l);
    case Op::TANH: return std::tanh(l);
    default: return 0.0;
  }
}

/// Adjoint-weighted ∂/∂lhs (what append_triplets adds into the left child).
inline double op_grad_l(Op op, double a, double l, double r) {
  switch (op) {
    case Op::SUB: case Op::ADD: return a;
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
    default: return 0.0;
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

// ---- node factories with the reference's pruning rules ---------------------

inline Expr operator-(const Expr& lhs);

inline ExpressionType scale_type(ExpressionType t) {
  using enum ExpressionType;
  return t == LINEAR ? LINEAR : (t == QUADRATIC ? QUADRATIC : NONLINEAR);
}
inline ExpressionType join_type(ExpressionType a, ExpressionType b) {
  return scale_type(std::max(a, b));
}

inline Expr operator*(const Expr& lhs, const Expr& rhs) {
  using enum ExpressionType;
  if (lhs.is_constant(0.0)) return lhs;
  if (rhs.is_constant(0.0)) return rhs;
  if (lhs.is_constant(1.0)) return rhs;
  if (rhs.is_constant(1.0)) return lhs;
  const auto lt = lhs.type(), rt = rhs.type();
  if (lt == CONSTANT && rt == CONSTANT) {
    return constant_ptr(lhs.val() * rhs.val());
  }
  ExpressionType t;
  if (lt == CONSTANT) {
    t = scale_type(rt);
  } else if (rt == CONSTANT) {
    t = scale_type(lt);
  } else if (lt == LINEAR && rt == LINEAR) {
    t = QUADRATIC;
  } else {
    t = NONLINEAR;
  }
  return make_expr(Op::MUL, t, lhs, rhs);
}

inline Expr operator/(const Expr& lhs, const Expr& rhs) {
  using enum ExpressionType;
  if (lhs.is_constant(0.0)) return lhs;
  if (rhs.is_constant(1.0)) return lhs;
  if (lhs.type() == CONSTANT && rhs.type() == CONSTANT) {
    return constant_ptr(lhs.val() / rhs.val());
  }
  return make_expr(Op::DIV,
                   rhs.type() == CONSTANT ? scale_type(lhs.type()) : NONLINEAR,
                   lhs, rhs);
}

/// Null-tolerant: also used to accumulate symbolic adjoints.
inline Expr operator+(const Expr& lhs, const Expr& rhs) {
  using enum ExpressionType;
  if (lhs == nullptr || lhs.is_constant(0.0)) return rhs;
  if (rhs == nullptr || rhs.is_constant(0.0)) return lhs;
  if (lhs.type() == CONSTANT && rhs.type() == CONSTANT) {
    return constant_ptr(lhs.val() + rhs.val());
  }
  return make_expr(Op::ADD, join_type(lhs.type(), rhs.type()), lhs, rhs);
}

inline Expr operator-(const Expr& lhs, const Expr& rhs) {
  using enum ExpressionType;
  if (lhs.is_constant(0.0)) {
    return rhs.is_constant(0.0) ? rhs : -rhs;
  }
  if (rhs.is_constant(0.0)) return lhs;
  if (lhs.type() == CONSTANT && rhs.type() == CONSTANT) {
    return constant_ptr(lhs.val() - rhs.val());
  }
  return make_expr(Op::SUB, join_type(lhs.type(), rhs.type()), lhs, rhs);
}

inline Expr operator-(const Expr& lhs) {
  using enum ExpressionType;
  if (lhs.is_constant(0.0)) return lhs;
  if (lhs.type() == CONSTANT) return constant_ptr(-lhs.val());
  return make_expr(Op::NEG, scale_type(lhs.type()), lhs);
}
inline Expr operator+(const Expr& lhs) { return lhs; }

/// Unary factory families of the reference: what a constant 0 argument folds
/// to differs per function.
enum class ZeroRule { SAME_NODE, ONE, NONE };

template <typename F>
Expr make_unary(Op op, const Expr& x, ZeroRule rule, F&& fold) {
  using enum ExpressionType;
  if (rule == ZeroRule::SAME_NODE && x.is_constant(0.0)) return x;
  if (rule == ZeroRule::ONE && x.is_constant(0.0)) return constant_ptr(1.0);
  if (x.type() == CONSTANT) return constant_ptr(fold(x.val()));
  return make_expr(op, NONLINEAR, x);
}

inline Expr abs(const Expr& x) {
  return make_unary(Op::ABS, x, ZeroRule::SAME_NODE,
                    [](double v) { return std::abs(v); });
}
inline Expr acos(const Expr& x) {
  if (x.is_constant(0.0)) return constant_ptr(std::numbers::pi / 2.0);
  return make_unary(Op::ACOS, x, ZeroRule::NONE,
                    [](double v) { return std::acos(v); });
}
inline Expr asin(const Expr& x) {
  return make_unary(Op::ASIN, x, ZeroRule::SAME_NODE,
                    [](double v) { return std::asin(v); });
}
inline Expr atan(const Expr& x) {
  return make_unary(Op::ATAN, x, ZeroRule::SAME_NODE,
                    [](double v) { return std::atan(v); });
}
inline Expr atan2(const Expr& y, const Expr& x) {
  using enum ExpressionType;
  if (y.type() == CONSTANT && x.type() == CONSTANT) {
    return constant_ptr(std::atan2(y.val(), x.val()));
  }
  return make_expr(Op::ATAN2, NONLINEAR, y, x);
}
inline Expr cbrt(const Expr& x) {
  if (x.type() == ExpressionType::CONSTANT) {
    const double v = x.val();
    if (v == 0.0 || v == -1.0 || v == 1.0) return x;
    return constant_ptr(std::cbrt(v));
  }
  return make_expr(Op::CBRT, ExpressionType::NONLINEAR, x);
}
inline Expr cos(const Expr& x) {
  return make_unary(Op::COS, x, ZeroRule::ONE,
                    [](double v) { return std::cos(v); });
}
inline Expr cosh(const Expr& x) {
  return make_unary(Op::COSH, x, ZeroRule::ONE,
                    [](double v) { return std::cosh(v); });
}
inline Expr erf(const Expr& x) {
  return make_unary(Op::ERF, x, ZeroRule::SAME_NODE,
                    [](double v) { return std::erf(v); });
}
inline Expr exp(const Expr& x) {
  return make_unary(Op::EXP, x, ZeroRule::ONE,
                    [](double v) { return std::exp(v); });
}
inline Expr hypot(const Expr& x, const Expr& y) {
  using enum ExpressionType;
  if (x.is_constant(0.0)) return abs(y);
  if (y.is_constant(0.0)) return abs(x);
  if (x.type() == CONSTANT && y.type() == CONSTANT) {
    return constant_ptr(std::hypot(x.val(), y.val()));
  }
  return make_expr(Op::HYPOT, NONLINEAR, x, y);
}
inline Expr is_nonnegative(const Expr& x) {
  return make_unary(Op::IS_NONNEG, x, ZeroRule::NONE,
                    [](double v) { return v >= 0.0 ? 1.0 : 0.0; });
}
inline Expr is_positive(const Expr& x) {
  return make_unary(Op::IS_POS, x, ZeroRule::NONE,
                    [](double v) { return v > 0.0 ? 1.0 : 0.0; });
}
inline Expr log(const Expr& x) {
  return make_unary(Op::LOG, x, ZeroRule::SAME_NODE,
                    [](double v) { return std::log(v); });
}
inline Expr log10(const Expr& x) {
  return make_unary(Op::LOG10, x, ZeroRule::SAME_NODE,
                    [](double v) { return std::log10(v); });
}
inline Expr max(const Expr& a, const Expr& b) {
  using enum ExpressionType;
  if (a.type() == CONSTANT && b.type() == CONSTANT) {
    return constant_ptr(std::max(a.val(), b.val()));
  }
  return make_expr(Op::MAX, NONLINEAR, a, b);
}
inline Expr min(const Expr& a, const Expr& b) {
  using enum ExpressionType;
  if (a.type() == CONSTANT && b.type() == CONSTANT) {
    return constant_ptr(std::min(a.val(), b.val()));
  }
  return make_expr(Op::MIN, NONLINEAR, a, b);
}
inline Expr pow(const Expr& base, const Expr& power) {
  using enum ExpressionType;
  if (base.is_constant(0.0) || base.is_constant(1.0)) return base;
  if (power.is_constant(0.0)) return constant_ptr(1.0);
  if (power.is_constant(1.0)) return base;
  if (base.type() == CONSTANT && power.type() == CONSTANT) {
    return constant_ptr(std::pow(base.val(), power.val()));
  }
  if (power.is_constant(2.0)) {
    // x² is emitted as a product so that it stays QUADRATIC for LINEAR x
    return make_expr(Op::MUL, base.type() == LINEAR ? QUADRATIC : NONLINEAR,
                     base, base);
  }
  return make_expr(Op::POW, NONLINEAR, base, power);
}
inline Expr sign(const Expr& x) {
  if (x.type() == ExpressionType::CONSTANT) {
    const double v = x.val();
    if (v < 0.0) return constant_ptr(-1.0);
    if (v == 0.0) return x;
    return constant_ptr(1.0);
  }
  return make_expr(Op::SIGN, ExpressionType::NONLINEAR, x);
}
inline Expr sin(const Expr& x) {
  return make_unary(Op::SIN, x, ZeroRule::SAME_NODE,
                    [](double v) { return std::sin(v); });
}
inline Expr sinh(const Expr& x) {
  return make_unary(Op::SINH, x, ZeroRule::SAME_NODE,
                    [](double v) { return std::sinh(v); });
}
inline Expr sqrt(const Expr& x) {
  if (x.type() == ExpressionType::CONSTANT) {
    const double v = x.val();
    if (v == 0.0 || v == 1.0) return x;
    return constant_ptr(std::sqrt(v));
  }
  return make_expr(Op::SQRT, ExpressionType::NONLINEAR, x);
}
inline Expr tan(const Expr& x) {
  return make_unary(Op::TAN, x, ZeroRule::SAME_NODE,
                    [](double v) { return std::tan(v); });
}
inline Expr tanh(const Expr& x) {
  return make_unary(Op::TANH, x, ZeroRule::SAME_NODE,
                    [](double v) { return std::tanh(v); });
}

/// Symbolic adjoint-weighted partials of node `op(l, r)` given its adjoint
/// expression `ae` (grad_expr_l / grad_expr_r of each reference subclass).
inline Expr grad_expr_l(Op op, const Expr& ae, const Expr& l, const Expr& r) {
  switch (op) {
    case Op::SUB: case Op::ADD: return ae;
    case Op::DIV: return ae / r;
    case Op::MUL: return ae * r;
    case Op::NEG: return -ae;
    case Op::ABS: return ae * sign(l);
    case Op::ACOS: return -ae / sqrt(constant_ptr(1.0) - l * l);
    case Op::ASIN: return ae / sqrt(constant_ptr(1.0) - l * l);
    case Op::ATAN: return ae / (constant_ptr(1.0) + l * l);
    case Op::ATAN2: return ae * r / (l * l + r * r);
    case Op::CBRT: {
      Expr c = cbrt(l);
      return ae / (constant_ptr(3.0) * c * c);
    }
    case Op::COS: return ae * -sin(l);
    case Op::COSH: return ae * sinh(l);
    case Op::ERF:
      return ae * constant_ptr(2.0 * std::numbers::inv_sqrtpi) * exp(-l * l);
    case Op::EXP: return ae * exp(l);
    case Op::HYPOT: return ae * l / hypot(l, r);
    case Op::LOG: return ae / l;
    case Op::LOG10: return ae / (constant_ptr(std::numbers::ln10) * l);
    case Op::MAX: return ae * is_nonnegative(l - r);
    case Op::MIN: return ae * is_nonnegative(r - l);
    case Op::POW: return ae * pow(l, r - constant_ptr(1.0)) * r;
    case Op::SIN: return ae * cos(l);
    case Op::SINH: return ae * cosh(l);
    case Op::SQRT: return ae / (constant_ptr(2.0) * sqrt(l));
    case Op::TAN: {
      Expr c = cos(l);
      return ae / (c * c);
    }
    case Op::TANH: {
      Expr c = cosh(l);
      return ae / (c * c);
    }
    default: return constant_ptr(0.0);
  }
}

inline Expr grad_expr_r(Op op, const Expr& ae, const Expr& l, const Expr& r) {
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

}  // namespace slp::detail
