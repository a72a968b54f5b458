// aqe_order.hpp -- row order of the reference's B+ tree for tables with duplicate ids (aqe_order.cpp).  Host only.
#pragma once

#include <cstddef>
#include <cstdint>

namespace aqe {

enum { ORDER_OP_BATCH = 0, ORDER_OP_RESTORE = 1 };
// One step of a table's history over its rows in arrival order: BATCH = insert_batch / load_from_file / a single insert_record
// (std::sort by id, then one tree insert per row; custom_bplus_db.cpp:196-206), RESTORE = rows that already are in tree order.
struct OrderOp { uint64_t rows; int kind; };

// Tables beyond this many rows keep the stable order by id (the replay holds the whole tree on the host).
constexpr uint64_t kReferenceOrderMaxRows = 1ull << 27;

// perm[k] = arrival number of the row at position k of the reference's leaf chain after replaying `ops`.
// false: ops do not cover n rows, or n is above kReferenceOrderMaxRows.
bool reference_order(const int64_t* ids, uint64_t n, const OrderOp* ops, size_t n_ops, uint64_t* perm);

}  // namespace aqe
