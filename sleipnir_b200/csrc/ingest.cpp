// Copies what the caller passed to slpb_upload_tape / slpb_upload_rows into
// the library's own host structures, validating indices on the way (the ABI
// promises error codes, not crashes, for inconsistent input).
#include "internal.hpp"

namespace slpb {

bool ingest_tape(Tape& t, int32_t n_nodes, const uint8_t* op,
                 const int32_t* lhs, const int32_t* rhs, const double* val,
                 int32_t n_x, const int32_t* leaf_x, int32_t n_y,
                 const int32_t* leaf_y, int32_t n_z, const int32_t* leaf_z,
                 std::string& error) {
  if (n_nodes < 0 || n_x < 0 || n_y < 0 || n_z < 0 ||
      (n_nodes > 0 && (!op || !lhs || !rhs || !val)) || (n_x > 0 && !leaf_x) ||
      (n_y > 0 && !leaf_y) || (n_z > 0 && !leaf_z)) {
    error = "slpb_upload_tape: null pointer or negative size";
    return false;
  }
  t = Tape{};
  t.n_nodes = n_nodes;
  t.op.assign(op, op + n_nodes);
  t.lhs.assign(lhs, lhs + n_nodes);
  t.rhs.assign(rhs, rhs + n_nodes);
  t.val.assign(val, val + n_nodes);
  t.n_x = n_x;
  t.n_y = n_y;
  t.n_z = n_z;
  for (int32_t i = 0; i < n_nodes; ++i) {
    if (t.op[i] >= SLPB_OP_COUNT) {
      error = "slpb_upload_tape: unknown opcode";
      return false;
    }
    // children precede parents
    if (t.lhs[i] >= i || t.rhs[i] >= i || t.lhs[i] < -1 || t.rhs[i] < -1 ||
        (t.lhs[i] < 0 && t.rhs[i] >= 0)) {
      error = "slpb_upload_tape: node list is not in child-before-parent order";
      return false;
    }
    const bool nullary = t.op[i] == SLPB_OP_CONST || t.op[i] == SLPB_OP_VAR;
    if (nullary != (t.lhs[i] < 0)) {
      error = "slpb_upload_tape: argument count does not match the opcode";
      return false;
    }
  }
  t.leaf_of_node.assign(n_nodes, -1);
  auto bind = [&](const int32_t* leaf, int32_t count, int32_t base) {
    for (int32_t i = 0; i < count; ++i) {
      if (leaf[i] < 0 || leaf[i] >= n_nodes || t.op[leaf[i]] != SLPB_OP_VAR) {
        error = "slpb_upload_tape: leaf index does not name a VAR node";
        return false;
      }
      t.leaf_of_node[leaf[i]] = base + i;
    }
    return true;
  };
  return bind(leaf_x, n_x, 0) && bind(leaf_y, n_y, n_x) &&
         bind(leaf_z, n_z, n_x + n_y);
}

bool ingest_rows(RowSet& r, const Tape& t, int which, const slpb_rowset* rows,
                 const double* const_val, std::string& error) {
  if (!rows || rows->n_rows < 0 || !rows->row_ptr) {
    error = "slpb_upload_rows: null row set";
    return false;
  }
  const bool value_rows =
      which == SLPB_OUT_F || which == SLPB_OUT_C_E || which == SLPB_OUT_C_I;
  r = RowSet{};
  r.present = true;
  r.n_rows = rows->n_rows;
  r.n_cols = rows->n_cols;
  r.row_ptr.assign(rows->row_ptr, rows->row_ptr + r.n_rows + 1);
  const int32_t n_list = r.row_ptr[r.n_rows];
  if (n_list > 0 && !rows->row_nodes) {
    error = "slpb_upload_rows: row_nodes is null";
    return false;
  }
  r.row_nodes.assign(rows->row_nodes, rows->row_nodes + n_list);
  for (int32_t id : r.row_nodes) {
    if (id < 0 || id >= t.n_nodes) {
      error = "slpb_upload_rows: node index out of range";
      return false;
    }
  }
  if (value_rows) {
    r.row_swept.assign(r.n_rows, 1);
    r.out_ptr.assign(r.n_rows + 1, 0);
    r.const_val.assign(r.n_rows, 0.0);
    if (const_val) r.const_val.assign(const_val, const_val + r.n_rows);
    return true;
  }
  if (r.n_rows == 0) {
    r.out_ptr.assign(1, 0);
    return true;
  }
  if (!rows->out_ptr || !rows->row_swept) {
    error = "slpb_upload_rows: derivative rows need out_ptr and row_swept";
    return false;
  }
  r.out_ptr.assign(rows->out_ptr, rows->out_ptr + r.n_rows + 1);
  const int32_t n_out = r.out_ptr[r.n_rows];
  if (n_out > 0) {
    r.out_col.assign(rows->out_col, rows->out_col + n_out);
    r.out_node.assign(rows->out_node, rows->out_node + n_out);
  }
  r.row_swept.assign(rows->row_swept, rows->row_swept + r.n_rows);
  for (int32_t k = 0; k < n_out; ++k) {
    if (r.out_col[k] < 0 || r.out_col[k] >= t.n_x || r.out_node[k] < 0 ||
        r.out_node[k] >= t.n_nodes) {
      error = "slpb_upload_rows: output (col, node) out of range";
      return false;
    }
  }
  if (rows->n_cached > 0) {
    r.cached_row.assign(rows->cached_row, rows->cached_row + rows->n_cached);
    r.cached_col.assign(rows->cached_col, rows->cached_col + rows->n_cached);
    r.cached_val.assign(rows->cached_val, rows->cached_val + rows->n_cached);
    for (int32_t k = 0; k < rows->n_cached; ++k) {
      if (r.cached_row[k] < 0 || r.cached_row[k] >= r.n_rows ||
          r.cached_col[k] < 0 || r.cached_col[k] >= t.n_x) {
        error = "slpb_upload_rows: cached triplet out of range";
        return false;
      }
    }
  }
  return true;
}

}  // namespace slpb
