/* oracle/sqlite_shim/sqlite3.h -- TEST INFRASTRUCTURE.
 *
 * This image ships the SQLite runtime (libsqlite3.so.0, the one Python's sqlite3 module loads) but not its
 * development header, so the reference's SQL-string path (src/aqe_backend/core/db.cpp, executor.cpp) cannot
 * be compiled as shipped.  This file declares -- from SQLite's documented, ABI-stable C interface -- exactly
 * the five functions and one constant that core/db.cpp uses, so that oracle/Makefile (target `refsql`) can
 * compile the UNMODIFIED reference sources and link them against the system runtime.  It is written from the
 * public API documentation, not copied from SQLite's header.
 */
#ifndef AQE_ORACLE_SQLITE3_SHIM_H
#define AQE_ORACLE_SQLITE3_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct sqlite3 sqlite3;
#define SQLITE_OK 0
int sqlite3_open(const char* filename, sqlite3** db);
int sqlite3_close(sqlite3* db);
const char* sqlite3_errmsg(sqlite3* db);
int sqlite3_exec(sqlite3* db, const char* sql, int (*callback)(void*, int, char**, char**), void* arg, char** errmsg);
void sqlite3_free(void* p);
#ifdef __cplusplus
}
#endif
#endif
