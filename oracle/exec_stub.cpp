// oracle/exec_stub.cpp -- TEST INFRASTRUCTURE.  The reference's SQL-string path (executor.cpp, core/db.cpp)
// needs sqlite3 development headers, which this image does not have; it is out of the hot path
// (SURVEY C5).  These four stubs let the reference's own bindings.cpp link so that its `aqe_backend`
// module can be imported for drop-in comparisons.
#include "executor.h"
#include <stdexcept>

double execute_query(const std::string&, const std::string&, int) {
    throw std::runtime_error("sqlite path not built in oracle/_ref");
}
GroupResult execute_query_groupby(const std::string&, const std::string&, int, int) {
    throw std::runtime_error("sqlite path not built in oracle/_ref");
}
QueryResult execute_query_with_ci(const std::string&, const std::string&, int) {
    throw std::runtime_error("sqlite path not built in oracle/_ref");
}
GroupResultWithCI execute_query_groupby_with_ci(const std::string&, const std::string&, int, int) {
    throw std::runtime_error("sqlite path not built in oracle/_ref");
}
