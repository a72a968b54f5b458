// oracle/ref_sql_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// C-ABI driver of the UNMODIFIED reference SQL-string path (src/aqe_backend/executor.cpp, parser.cpp,
// core/db.cpp, compiled where they lie by `make -C oracle refsql` against the system SQLite runtime through
// oracle/sqlite_shim/sqlite3.h).  Used by tests/golden/make_sql_golden.py to mint the golden vectors that pin
// oracle/aqe_oracle.c's restatement of that path (orc_sql_*), and by tests/test_oracle_vs_ref.py.
// Every call returns 0 on success; a C++ exception becomes 1 and its what() is left in ref_sql_error().
#include "executor.h"
#include "parser.h"

#include <cstring>
#include <string>

static thread_local std::string g_err;

struct RefSqlRow { char key[64]; double value, ci_lower, ci_upper; };

extern "C" {

const char* ref_sql_error() { return g_err.c_str(); }

int ref_sql_parse(const char* sql, int p, char* agg, char* column, char* table, char* where, char* group_by, size_t cap) {
    try {
        Query q = parse_query(sql, p);
        std::strncpy(agg, q.agg.c_str(), cap - 1); agg[cap - 1] = 0;
        std::strncpy(column, q.column.c_str(), cap - 1); column[cap - 1] = 0;
        std::strncpy(table, q.table.c_str(), cap - 1); table[cap - 1] = 0;
        std::strncpy(where, q.where.c_str(), cap - 1); where[cap - 1] = 0;
        std::strncpy(group_by, q.group_by.c_str(), cap - 1); group_by[cap - 1] = 0;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int ref_run_query(const char* sql, const char* db, int p, double* out) {
    try { *out = execute_query(sql, db, p); return 0; } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int ref_run_query_with_ci(const char* sql, const char* db, int p, double* v, double* lo, double* hi) {
    try {
        QueryResult r = execute_query_with_ci(sql, db, p);
        *v = r.value; *lo = r.ci_lower; *hi = r.ci_upper;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// rows come back in std::map order (lexicographic by key string); returns the number of groups in *n
int ref_run_query_groupby(const char* sql, const char* db, int p, int threads, RefSqlRow* rows, size_t cap, size_t* n) {
    try {
        GroupResult r = execute_query_groupby(sql, db, p, threads);
        size_t i = 0;
        for (auto& kv : r) {
            if (i < cap) { std::strncpy(rows[i].key, kv.first.c_str(), 63); rows[i].key[63] = 0; rows[i].value = rows[i].ci_lower = rows[i].ci_upper = kv.second; }
            ++i;
        }
        *n = i;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int ref_run_query_groupby_with_ci(const char* sql, const char* db, int p, int threads, RefSqlRow* rows, size_t cap, size_t* n) {
    try {
        GroupResultWithCI r = execute_query_groupby_with_ci(sql, db, p, threads);
        size_t i = 0;
        for (auto& kv : r) {
            if (i < cap) { std::strncpy(rows[i].key, kv.first.c_str(), 63); rows[i].key[63] = 0; rows[i].value = kv.second.value; rows[i].ci_lower = kv.second.ci_lower; rows[i].ci_upper = kv.second.ci_upper; }
            ++i;
        }
        *n = i;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

}  // extern "C"
