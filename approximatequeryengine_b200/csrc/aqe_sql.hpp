// aqe_sql.hpp -- host half of the SQL-string path: the reference's substring parser
// (src/aqe_backend/parser.cpp:20-75) restated, the WHERE clause compiled to per-column intervals, and the
// result arithmetic of src/aqe_backend/executor.cpp:28-338 applied to the integer accumulators the grouped
// scan kernel (k_sql_agg, aqe_sql_kernels.cuh) leaves behind.  Pure host code; nothing here touches CUDA.
#pragma once

#include <cstdint>
#include <string>

#include "aqe_b200.h"

namespace aqe {

// parser.cpp:20-75 + name resolution + WHERE compilation.  Returns an aqe_status.
int sql_parse(const std::string& sql, int sample_percent, aqe_sql_query& out, std::string& err);

// executor.cpp:20-26
inline int sql_sample_step(int sample_percent) {
    if (sample_percent <= 0 || sample_percent >= 100) return 0;
    const int step = 100 / sample_percent;
    return step <= 0 ? 1 : step;
}

// Fixed-point shifts for values of magnitude <= absmax: |x| * 2^sum_shift < 2^62, x^2 * 2^sq_shift < 2^62.
int sql_shifts(double absmax, bool is_integer, int& sum_shift, int& sq_shift, std::string& err);

int sql_layout(const aqe_sql_query& q, const aqe_sql_facts* facts, int n, aqe_sql_layout& out, std::string& err);

void sql_merge(uint64_t* acc, const uint64_t* other, uint32_t n_groups);

int sql_finish(const aqe_sql_query& q, int mode, const aqe_sql_layout& L, const uint64_t* acc, const uint64_t* exists,
               aqe_sql_row* rows, uint32_t cap, uint32_t* n_rows, std::string& err);

// Does the reference formula of this (query, mode) read the column sums even though the aggregate is COUNT?
// (execute_query_groupby_with_ci computes SUM(col)/COUNT(col) for every aggregate, executor.cpp:262-318)
inline bool sql_needs_sums(const aqe_sql_query& q, int mode) {
    if (q.agg != AQE_AGG_COUNT) return true;
    return mode == AQE_SQL_CI_REFERENCE && q.group_col != AQE_COL_NONE;
}
inline bool sql_needs_moments(const aqe_sql_query& q, int mode) {
    if (mode == AQE_SQL_VALUE) return false;
    if (q.group_col != AQE_COL_NONE) return mode == AQE_SQL_CI_REFERENCE || q.agg != AQE_AGG_COUNT;
    return q.agg != AQE_AGG_COUNT && sql_sample_step(q.sample_percent) > 0;
}

}  // namespace aqe
