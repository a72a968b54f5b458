/*
 * oracle/aqe_oracle_sql.c -- TEST INFRASTRUCTURE (the checker), never the product.
 *
 * Plain-C, single-threaded restatement of the reference's SQL-string path over the record table:
 *   parse_query                      src/aqe_backend/parser.cpp:20-75
 *   execute_query                    src/aqe_backend/executor.cpp:28-58
 *   execute_query_groupby            src/aqe_backend/executor.cpp:60-130
 *   execute_query_with_ci            src/aqe_backend/executor.cpp:133-243
 *   execute_query_groupby_with_ci    src/aqe_backend/executor.cpp:245-338
 * The reference hands the statements it assembles to SQLite (a dependency that is not in /root/reference:
 * libsqlite3, system package; 3.45.1 in this image).  What SQLite contributes to the path is restated from its
 * documented behaviour: integer vs. real comparison is exact; SUM() over REAL uses Kahan-Babuska-Neumaier
 * compensation (3.43+), over INTEGER it is exact int64 and raises "integer overflow"; results come back through
 * sqlite3_exec as TEXT, REAL rendered with 15 significant digits ("%!.15g"), which the reference re-parses with
 * std::stod; an aggregate over no rows is NULL, rendered "NULL" by core/db.cpp:12, on which std::stod throws.
 * The table is `sales(id INTEGER PRIMARY KEY, amount REAL, region INTEGER, product_id INTEGER, timestamp INTEGER)`
 * holding the record file's rows, so rowid = id.
 *
 * PARITY PIN: the reference has no tests for this path.  tests/golden/sql_*.json are minted by
 * tests/golden/make_sql_golden.py from the UNMODIFIED reference sources (executor.cpp, parser.cpp, core/db.cpp
 * compiled by `make -C oracle refsql` against the system SQLite through oracle/sqlite_shim/sqlite3.h) run on a
 * SQLite copy of the same rows; tests/test_oracle_golden.py holds this file to them (values to 2 units in the
 * 15th digit -- SQLite's own REAL->TEXT conversion is not always correctly rounded -- counts, keys, errors exact).
 *
 * Row-level WHERE evaluation here is deliberately a different mechanism from the engine's (which compiles the
 * clause to OR-ed conjunctions of per-column intervals): the clause becomes a postfix program of the comparisons as
 * written, AND and OR, evaluated per row.
 */
#include "aqe_b200.h"

#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#define ORC_API __attribute__((visibility("default")))

/* status codes of orc_sql_run */
enum { ORC_SQL_OK = 0, ORC_SQL_RUNTIME_ERROR = 1 /* std::runtime_error in the reference */, ORC_SQL_STOD = 2 /* std::stod("NULL") */,
       ORC_SQL_UNSUPPORTED = 3 /* valid for SQLite, outside the restated grammar */ };

typedef struct orc_sql_parsed { char agg[32], column[64], table[64], where[512], group_by[64]; } orc_sql_parsed;
typedef struct orc_sql_row { int64_t key; double value, ci_lower, ci_upper; } orc_sql_row;

static void seterr(char* err, size_t cap, const char* msg) { if (err && cap) { strncpy(err, msg, cap - 1); err[cap - 1] = 0; } }

/* ---- parser.cpp:20-75 ------------------------------------------------------------------------------------ */
static const char* find_ci(const char* hay, const char* needle_upper) { /* first match of an upper-case needle in upper(hay) */
    const size_t n = strlen(needle_upper);
    for (const char* p = hay; *p; ++p) {
        size_t k = 0;
        while (k < n && p[k] && toupper((unsigned char)p[k]) == needle_upper[k]) ++k;
        if (k == n) return p;
    }
    return NULL;
}
static void trim_copy(char* dst, size_t cap, const char* b, const char* e) { /* [b, e) trimmed of " \t\n\r" */
    while (b < e && strchr(" \t\n\r", *b)) ++b;
    while (e > b && strchr(" \t\n\r", e[-1])) --e;
    size_t n = (size_t)(e - b);
    if (n > cap - 1) n = cap - 1;
    memcpy(dst, b, n);
    dst[n] = 0;
}
static void drop_semicolon(char* s) { const size_t n = strlen(s); if (n && s[n - 1] == ';') s[n - 1] = 0; }

ORC_API int orc_sql_parse(const char* sql, orc_sql_parsed* q, char* err, size_t errcap) {
    memset(q, 0, sizeof(*q));
    const char* end = sql + strlen(sql);
    const char* sel = find_ci(sql, "SELECT");
    const char* from = find_ci(sql, "FROM");
    if (!sel || !from) { seterr(err, errcap, "Invalid SQL: missing SELECT or FROM"); return ORC_SQL_RUNTIME_ERROR; }
    char agg_col[256];
    trim_copy(agg_col, sizeof(agg_col), sel + 6 <= end ? sel + 6 : end, from >= sel + 6 ? from : end);
    const char* where = find_ci(sql, "WHERE");
    const char* group = find_ci(sql, "GROUP BY");
    if (where) {
        trim_copy(q->table, sizeof(q->table), from + 4, where >= from + 4 ? where : end);
        if (group) {
            trim_copy(q->where, sizeof(q->where), where + 5, group >= where + 5 ? group : end);
            trim_copy(q->group_by, sizeof(q->group_by), group + 8, end);
        } else {
            trim_copy(q->where, sizeof(q->where), where + 5, end);
        }
    } else if (group) {
        trim_copy(q->table, sizeof(q->table), from + 4, group >= from + 4 ? group : end);
        trim_copy(q->group_by, sizeof(q->group_by), group + 8, end);
    } else {
        trim_copy(q->table, sizeof(q->table), from + 4, end);
    }
    drop_semicolon(q->table); drop_semicolon(q->group_by); drop_semicolon(q->where);
    const char* po = strchr(agg_col, '(');
    const char* pc = strchr(agg_col, ')');
    if (!po || !pc) { seterr(err, errcap, "Invalid aggregation syntax"); return ORC_SQL_RUNTIME_ERROR; }
    trim_copy(q->agg, sizeof(q->agg), agg_col, po);
    trim_copy(q->column, sizeof(q->column), po + 1, pc > po ? pc : po + 1);
    char up[32];
    size_t i = 0;
    for (; q->agg[i] && i < sizeof(up) - 1; ++i) up[i] = (char)toupper((unsigned char)q->agg[i]);
    up[i] = 0;
    if (strcmp(up, "SUM") && strcmp(up, "COUNT") && strcmp(up, "AVG")) {
        char m[160];
        snprintf(m, sizeof(m), "Unsupported aggregation function: %s. Supported functions: SUM, COUNT, AVG", q->agg);
        seterr(err, errcap, m);
        return ORC_SQL_RUNTIME_ERROR;
    }
    return ORC_SQL_OK;
}

/* ---- SQLite value semantics ------------------------------------------------------------------------------ */
typedef struct val { int is_int; int64_t i; double d; } val;

/* exact comparison of an int64 with a double (what SQLite does for INTEGER vs REAL) */
static int cmp_int_real(int64_t i, double r) {
    if (r < -9223372036854775808.0) return 1;
    if (r >= 9223372036854775808.0) return -1;
    const int64_t y = (int64_t)r;
    if (i < y) return -1;
    if (i > y) return 1;
    const double s = (double)i;
    return s < r ? -1 : (s > r ? 1 : 0);
}
static int cmp_val(val a, val b) {
    if (a.is_int && b.is_int) return a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);
    if (a.is_int) return cmp_int_real(a.i, b.d);
    if (b.is_int) return -cmp_int_real(b.i, a.d);
    return a.d < b.d ? -1 : (a.d > b.d ? 1 : 0);
}

static int column_index(const char* name, size_t n) { /* 0 id 1 amount 2 region 3 product_id 4 timestamp */
    static const struct { const char* n; int c; } names[] = {{"ID", 0}, {"ROWID", 0}, {"_ROWID_", 0}, {"OID", 0}, {"AMOUNT", 1},
                                                              {"REGION", 2}, {"PRODUCT_ID", 3}, {"TIMESTAMP", 4}};
    for (size_t k = 0; k < sizeof(names) / sizeof(names[0]); ++k) {
        if (strlen(names[k].n) != n) continue;
        size_t j = 0;
        while (j < n && toupper((unsigned char)name[j]) == names[k].n[j]) ++j;
        if (j == n) return names[k].c;
    }
    return -1;
}
static val column_value(const aqe_record* r, int c) {
    val v = {1, 0, 0.0};
    switch (c) {
        case 0: v.i = r->id; break;
        case 1: v.is_int = 0; v.d = r->amount; break;
        case 2: v.i = r->region; break;
        case 3: v.i = r->product_id; break;
        default: v.i = r->timestamp; break;
    }
    return v;
}

/* ---- WHERE: AND / OR / parentheses over comparisons and BETWEENs, compiled to a postfix program ---------- */
typedef struct operand { int col; val lit; } operand; /* col >= 0: column reference */
typedef struct cond { operand a, b, c; int op; /* 0 = 1 != 2 < 3 <= 4 > 5 >= 6 BETWEEN 7 IN */ operand list[16]; int nlist; } cond;
enum { W_COND = 0, W_AND = 1, W_OR = 2, W_NOT = 3 };
typedef struct where_prog { cond conds[32]; int n; struct { int kind, arg; } code[96]; int ncode; int top_level_or; } where_prog;

typedef struct scanner { const char* p; int status; char err[160]; int depth; } scanner;
static void skip_ws(scanner* s) { while (*s->p && isspace((unsigned char)*s->p)) ++s->p; }
static int keyword_at(scanner* s, const char* kw) { /* case-insensitive keyword followed by a non-identifier character */
    skip_ws(s);
    const size_t n = strlen(kw);
    for (size_t k = 0; k < n; ++k) if (toupper((unsigned char)s->p[k]) != kw[k]) return 0;
    return !(isalnum((unsigned char)s->p[n]) || s->p[n] == '_');
}
static int parse_number(const char* b, const char* e, val* out) {
    char buf[80];
    const size_t n = (size_t)(e - b);
    if (n == 0 || n >= sizeof(buf)) return 0;
    memcpy(buf, b, n); buf[n] = 0;
    char* end = NULL;
    if (!strpbrk(buf, ".eE")) {
        errno = 0;
        const long long v = strtoll(buf, &end, 10);
        if (*end == 0 && errno == 0) { out->is_int = 1; out->i = v; out->d = (double)v; return 1; }
    }
    const double d = strtod(buf, &end);
    if (*end) return 0;
    out->is_int = 0; out->d = d; out->i = 0;
    return 1;
}
static int parse_operand(scanner* s, operand* o) {
    skip_ws(s);
    const char* p = s->p;
    if (isalpha((unsigned char)*p) || *p == '_') {
        const char* e = p;
        while (isalnum((unsigned char)*e) || *e == '_') ++e;
        o->col = column_index(p, (size_t)(e - p));
        if (o->col < 0) {
            snprintf(s->err, sizeof(s->err), "SQL error: no such column: %.*s", (int)(e - p), p);
            s->status = ORC_SQL_RUNTIME_ERROR;
            return 0;
        }
        s->p = e;
        return 1;
    }
    o->col = -1;
    if (*p == '\'') {
        const char* e = strchr(p + 1, '\'');
        if (!e) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
        const char* b = p + 1;
        const char* t = e;
        while (b < t && isspace((unsigned char)*b)) ++b;
        while (t > b && isspace((unsigned char)t[-1])) --t;
        if (!parse_number(b, t, &o->lit)) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
        s->p = e + 1;
        return 1;
    }
    const char* e = p;
    if (*e == '-' || *e == '+') { ++e; while (isspace((unsigned char)*e)) ++e; }
    const char* digits = e;
    while (isdigit((unsigned char)*e) || *e == '.') ++e;
    if (e == digits) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
    if (*e == 'e' || *e == 'E') {
        const char* f = e + 1;
        if (*f == '+' || *f == '-') ++f;
        if (isdigit((unsigned char)*f)) { while (isdigit((unsigned char)*f)) ++f; e = f; }
    }
    char buf[80];
    size_t n = 0;
    if (*p == '-') buf[n++] = '-';
    for (const char* c = digits; c < e && n < sizeof(buf) - 1; ++c) buf[n++] = *c;
    buf[n] = 0;
    if (!parse_number(buf, buf + n, &o->lit)) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
    s->p = e;
    return 1;
}
static int emit(scanner* s, where_prog* w, int kind, int arg) {
    if (w->ncode >= 96) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
    w->code[w->ncode].kind = kind; w->code[w->ncode].arg = arg; ++w->ncode;
    return 1;
}
static int parse_or(scanner* s, where_prog* w);
static int parse_term(scanner* s, where_prog* w) {
    skip_ws(s);
    if (keyword_at(s, "NOT")) { /* NOT binds tighter than AND */
        s->p += 3;
        return parse_term(s, w) && emit(s, w, W_NOT, 0);
    }
    if (*s->p == '(') {
        ++s->p;
        ++s->depth;
        if (!parse_or(s, w)) return 0;
        --s->depth;
        skip_ws(s);
        if (*s->p != ')') { s->status = ORC_SQL_UNSUPPORTED; return 0; }
        ++s->p;
        return 1;
    }
    if (w->n >= 32) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
    cond* c = &w->conds[w->n];
    memset(c, 0, sizeof(*c));
    if (!parse_operand(s, &c->a)) return 0;
    int negated = 0; /* col NOT BETWEEN ... / col NOT IN (...) */
    if (keyword_at(s, "NOT")) {
        s->p += 3;
        negated = 1;
        if (!keyword_at(s, "BETWEEN") && !keyword_at(s, "IN")) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
    }
    if (keyword_at(s, "BETWEEN")) {
        s->p += 7;
        if (!parse_operand(s, &c->b)) return 0;
        if (!keyword_at(s, "AND")) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
        s->p += 3;
        if (!parse_operand(s, &c->c)) return 0;
        c->op = 6;
        return emit(s, w, W_COND, w->n++) && (!negated || emit(s, w, W_NOT, 0));
    }
    if (keyword_at(s, "IN")) {
        s->p += 2;
        skip_ws(s);
        if (*s->p != '(') { s->status = ORC_SQL_UNSUPPORTED; return 0; }
        ++s->p;
        c->op = 7;
        for (;;) {
            if (c->nlist >= 16) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
            if (!parse_operand(s, &c->list[c->nlist])) return 0;
            if (c->list[c->nlist].col >= 0) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
            ++c->nlist;
            skip_ws(s);
            if (*s->p == ',') { ++s->p; continue; }
            break;
        }
        if (*s->p != ')') { s->status = ORC_SQL_UNSUPPORTED; return 0; }
        ++s->p;
        return emit(s, w, W_COND, w->n++) && (!negated || emit(s, w, W_NOT, 0));
    }
    skip_ws(s);
    static const struct { const char* t; int op; } ops[] = {{"<=", 3}, {">=", 5}, {"<>", 1}, {"!=", 1}, {"==", 0}, {"<", 2}, {">", 4}, {"=", 0}};
    int op = -1;
    for (size_t k = 0; k < sizeof(ops) / sizeof(ops[0]); ++k)
        if (!strncmp(s->p, ops[k].t, strlen(ops[k].t))) { op = ops[k].op; s->p += strlen(ops[k].t); break; }
    if (op < 0) { s->status = ORC_SQL_UNSUPPORTED; return 0; }
    c->op = op;
    if (!parse_operand(s, &c->b)) return 0;
    return emit(s, w, W_COND, w->n++);
}
static int parse_and(scanner* s, where_prog* w) { /* AND binds tighter than OR */
    if (!parse_term(s, w)) return 0;
    while (keyword_at(s, "AND")) {
        s->p += 3;
        if (!parse_term(s, w) || !emit(s, w, W_AND, 0)) return 0;
    }
    return 1;
}
static int parse_or(scanner* s, where_prog* w) {
    if (!parse_and(s, w)) return 0;
    while (keyword_at(s, "OR")) {
        if (s->depth == 0) w->top_level_or = 1;
        s->p += 2;
        if (!parse_and(s, w) || !emit(s, w, W_OR, 0)) return 0;
    }
    return 1;
}
static int compile_where(const char* text, where_prog* w, char* err, size_t errcap) {
    w->n = 0; w->ncode = 0; w->top_level_or = 0;
    scanner s = {text, ORC_SQL_OK, "", 0};
    skip_ws(&s);
    if (!*s.p) return ORC_SQL_OK;
    if (!parse_or(&s, w)) { seterr(err, errcap, s.err[0] ? s.err : "WHERE clause outside the restated grammar"); return s.status ? s.status : ORC_SQL_UNSUPPORTED; }
    skip_ws(&s);
    if (*s.p) { seterr(err, errcap, "WHERE clause outside the restated grammar"); return ORC_SQL_UNSUPPORTED; }
    return ORC_SQL_OK;
}
static val operand_value(const operand* o, const aqe_record* r) { return o->col >= 0 ? column_value(r, o->col) : o->lit; }
static int cond_holds(const cond* c, const aqe_record* r) {
    const val a = operand_value(&c->a, r);
    if (c->op == 7) {
        for (int k = 0; k < c->nlist; ++k) if (cmp_val(a, c->list[k].lit) == 0) return 1;
        return 0;
    }
    const val b = operand_value(&c->b, r);
    if (c->op == 6) return cmp_val(a, b) >= 0 && cmp_val(a, operand_value(&c->c, r)) <= 0;
    const int x = cmp_val(a, b);
    return c->op == 0 ? x == 0 : c->op == 1 ? x != 0 : c->op == 2 ? x < 0 : c->op == 3 ? x <= 0 : c->op == 4 ? x > 0 : x >= 0;
}
static int row_passes(const where_prog* w, const aqe_record* r) {
    if (w->ncode == 0) return 1;
    int stack[96], top = 0;
    for (int k = 0; k < w->ncode; ++k) {
        if (w->code[k].kind == W_COND) stack[top++] = cond_holds(&w->conds[w->code[k].arg], r);
        else if (w->code[k].kind == W_NOT) stack[top - 1] = !stack[top - 1];
        else {
            const int y = stack[--top], x = stack[--top];
            stack[top++] = w->code[k].kind == W_AND ? (x && y) : (x || y);
        }
    }
    return stack[0];
}

/* ---- SQLite aggregates ------------------------------------------------------------------------------------ */
typedef struct sumctx { /* sum() / avg() / count() state of one statement */
    int64_t cnt;
    int approx, ovrfl;     /* a REAL was seen; int64 overflow */
    int64_t isum;
    double rsum, rerr;     /* Kahan-Babuska-Neumaier */
} sumctx;
static void kbn_step(sumctx* p, double r) {
    const double s = p->rsum;
    volatile double t = s + r;
    if (fabs(s) > fabs(r)) p->rerr += (s - t) + r; else p->rerr += (r - t) + s;
    p->rsum = t;
}
static void kbn_step_int(sumctx* p, int64_t v) {
    if (v <= -4503599627370496LL || v >= 4503599627370496LL) {
        const int64_t small = v % 16384;
        kbn_step(p, (double)(v - small));
        kbn_step(p, (double)small);
    } else kbn_step(p, (double)v);
}
static void sum_step(sumctx* p, val v) {
    p->cnt++;
    if (!p->approx) {
        if (v.is_int) {
            int64_t x;
            if (!__builtin_add_overflow(p->isum, v.i, &x)) { p->isum = x; return; }
            p->ovrfl = 1;
            kbn_step_int(p, p->isum); p->approx = 1;
            kbn_step_int(p, v.i);
            return;
        }
        kbn_step_int(p, p->isum); p->approx = 1;
        kbn_step(p, v.d);
        return;
    }
    if (v.is_int) kbn_step_int(p, v.i); else kbn_step(p, v.d);
}
/* REAL -> TEXT ("%!.15g") -> std::stod */
static double through_text(double v) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%.15g", v);
    return strtod(buf, NULL);
}
/* value of SUM(): 0 ok, 1 NULL, 2 "integer overflow" */
static int sum_final(const sumctx* p, double* out) {
    if (p->cnt == 0) return 1;
    if (p->approx) {
        if (p->ovrfl) return 2;
        *out = through_text(p->rsum + p->rerr);
    } else *out = (double)p->isum; /* exact decimal text -> nearest double */
    return 0;
}
static int avg_final(const sumctx* p, double* out) {
    if (p->cnt == 0) return 1;
    const double r = p->approx ? p->rsum + p->rerr : (double)p->isum;
    *out = through_text(r / (double)p->cnt);
    return 0;
}

/* `col * col` as SQLite evaluates it: int64 product, REAL on overflow */
static val square(val v) {
    val r;
    if (v.is_int) {
        int64_t x;
        if (!__builtin_mul_overflow(v.i, v.i, &x)) { r.is_int = 1; r.i = x; r.d = (double)x; return r; }
        r.is_int = 0; r.i = 0; r.d = (double)v.i * (double)v.i;
        return r;
    }
    r.is_int = 0; r.i = 0; r.d = v.d * v.d;
    return r;
}

static int sample_step(int p) { /* executor.cpp:20-26 */
    if (p <= 0 || p >= 100) return 0;
    const int s = 100 / p;
    return s <= 0 ? 1 : s;
}

typedef struct query { int agg; /* 0 SUM 1 AVG 2 COUNT */ int col; /* -1: '*' */ int group; where_prog w; } query;

static int resolve(const orc_sql_parsed* pq, query* q, int need_group, char* err, size_t errcap) {
    char up[32];
    size_t i = 0;
    for (; pq->agg[i] && i < sizeof(up) - 1; ++i) up[i] = (char)toupper((unsigned char)pq->agg[i]);
    up[i] = 0;
    q->agg = !strcmp(up, "SUM") ? 0 : (!strcmp(up, "AVG") ? 1 : 2);
    if (!strcmp(pq->column, "*")) {
        if (q->agg != 2) { seterr(err, errcap, "SQL error: wrong number of arguments to function"); return ORC_SQL_RUNTIME_ERROR; }
        q->col = -1;
    } else {
        q->col = column_index(pq->column, strlen(pq->column));
        if (q->col < 0) { seterr(err, errcap, "SQL error: no such column"); return ORC_SQL_RUNTIME_ERROR; }
    }
    q->group = -1;
    if (pq->group_by[0]) {
        q->group = column_index(pq->group_by, strlen(pq->group_by));
        if (q->group < 0) { seterr(err, errcap, "SQL error: no such column"); return ORC_SQL_RUNTIME_ERROR; }
        if (q->group == 1) { seterr(err, errcap, "GROUP BY on a REAL column is outside the restated grammar"); return ORC_SQL_UNSUPPORTED; }
    } else if (need_group) { seterr(err, errcap, "No GROUP BY column found"); return ORC_SQL_RUNTIME_ERROR; }
    return compile_where(pq->where, &q->w, err, errcap);
}

/* One statement `SELECT COUNT(c), SUM(c), SUM(c*c) ... WHERE [group = key AND] where [AND rowid % step = 0]`. */
static void run_statement(const aqe_record* rows, uint64_t n, const query* q, int step, int use_key, int64_t key, sumctx* s, sumctx* sq) {
    memset(s, 0, sizeof(*s));
    memset(sq, 0, sizeof(*sq));
    for (uint64_t i = 0; i < n; ++i) {
        const aqe_record* r = &rows[i];
        if (use_key && column_value(r, q->group).i != key) continue;
        if (!row_passes(&q->w, r)) continue;
        if (step > 0 && r->id % step != 0) continue;
        if (q->col < 0) { s->cnt++; continue; }
        const val v = column_value(r, q->col);
        sum_step(s, v);
        sum_step(sq, square(v));
    }
}

/* value of `SELECT agg(col) ...` as executor.cpp:44-56 reads it: 0 ok, ORC_SQL_STOD, ORC_SQL_RUNTIME_ERROR */
static int plain_value(const query* q, const sumctx* s, int step, int p, double* out, char* err, size_t errcap) {
    double v = 0.0;
    int st = 0;
    if (q->agg == 2) v = (double)s->cnt;
    else st = q->agg == 0 ? sum_final(s, &v) : avg_final(s, &v);
    if (st == 1) { seterr(err, errcap, "stod"); return ORC_SQL_STOD; }
    if (st == 2) { seterr(err, errcap, "SQL error: integer overflow"); return ORC_SQL_RUNTIME_ERROR; }
    if (step > 0 && q->agg != 1) v = v * (100.0 / p);
    *out = v;
    return ORC_SQL_OK;
}

/* mean / margin of executor.cpp:205-241 and :293-318 from the three TEXT results */
static int ci_value(const query* q, const sumctx* s, const sumctx* sq, int p, orc_sql_row* r, char* err, size_t errcap) {
    double sum = 0.0, sum_sq = 0.0;
    const double count = (double)s->cnt;
    int st = sum_final(s, &sum);
    if (st == 0) st = sum_final(sq, &sum_sq);
    if (st == 1) { seterr(err, errcap, "stod"); return ORC_SQL_STOD; }
    if (st == 2) { seterr(err, errcap, "SQL error: integer overflow"); return ORC_SQL_RUNTIME_ERROR; }
    double mean = sum / count;
    const double variance = (sum_sq - (sum * sum / count)) / (count - 1);
    const double std_error = sqrt(variance / count);
    double margin = 1.96 * std_error;
    if (q->agg == 0) {
        const double scale_factor = 100.0 / p;
        mean *= scale_factor;
        margin *= scale_factor;
    }
    r->value = mean; r->ci_lower = mean - margin; r->ci_upper = mean + margin;
    return ORC_SQL_OK;
}

static int cmp_i64(const void* a, const void* b) { const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b; return x < y ? -1 : (x > y ? 1 : 0); }

/* mode 0 run_query | 1 run_query_with_ci | 2 run_query_groupby | 3 run_query_groupby_with_ci.
 * Groups are reported in ascending NUMERIC key order (the reference's std::map orders the key strings). */
ORC_API int orc_sql_run(const aqe_record* rows, uint64_t n, const char* sql, int sample_percent, int mode, orc_sql_row* out,
                        uint32_t cap, uint32_t* n_out, char* err, size_t errcap) {
    orc_sql_parsed pq;
    int rc = orc_sql_parse(sql, &pq, err, errcap);
    if (rc) return rc;
    query q;
    rc = resolve(&pq, &q, mode >= 2, err, errcap);
    if (rc) return rc;
    const int step = sample_step(sample_percent);
    if (q.w.top_level_or && (mode >= 2 || step > 0)) {
        /* executor.cpp:38-42, :95-98 paste `group = 'k' AND ` / ` AND rowid % step = 0` around the clause TEXT; with an OR outside
         * parentheses SQLite binds them to the first / last branch only.  That accident is not restated (nor built by the engine). */
        seterr(err, errcap, "top-level OR in a sampled or grouped query");
        return ORC_SQL_UNSUPPORTED;
    }
    sumctx s, sq;
    *n_out = 0;
    if (mode == 0 || mode == 1) {
        orc_sql_row r = {0, 0.0, 0.0, 0.0};
        const int stats = mode == 1 && step > 0 && q.agg != 2; /* executor.cpp:183-187 */
        run_statement(rows, n, &q, step, 0, 0, &s, &sq);
        if (stats && q.col >= 0) {
            if (s.cnt == 0) { seterr(err, errcap, "stod"); return ORC_SQL_STOD; } /* SUM(col) is NULL */
            /* the statistics statement runs first: its SUM(col*col) can overflow even when the answer would not */
            double probe;
            if (sum_final(&sq, &probe) == 2 || sum_final(&s, &probe) == 2) { seterr(err, errcap, "SQL error: integer overflow"); return ORC_SQL_RUNTIME_ERROR; }
            if (s.cnt >= 2) {
                rc = ci_value(&q, &s, &sq, sample_percent, &r, err, errcap);
                if (rc) return rc;
                if (cap) out[0] = r;
                *n_out = 1;
                return ORC_SQL_OK;
            }
        }
        rc = plain_value(&q, &s, step, sample_percent, &r.value, err, errcap);
        if (rc) return rc;
        r.ci_lower = r.ci_upper = r.value;
        if (cap) out[0] = r;
        *n_out = 1;
        return ORC_SQL_OK;
    }
    /* SELECT DISTINCT group FROM t [WHERE where]  -- unsampled (executor.cpp:68-79, :253-265) */
    int64_t* keys = (int64_t*)malloc(sizeof(int64_t) * (n ? n : 1));
    uint64_t nk = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (row_passes(&q.w, &rows[i])) keys[nk++] = column_value(&rows[i], q.group).i;
    qsort(keys, nk, sizeof(int64_t), cmp_i64);
    uint64_t u = 0;
    for (uint64_t i = 0; i < nk; ++i) if (i == 0 || keys[i] != keys[i - 1]) keys[u++] = keys[i];
    uint32_t m = 0;
    for (uint64_t g = 0; g < u; ++g) {
        orc_sql_row r = {keys[g], 0.0, 0.0, 0.0};
        run_statement(rows, n, &q, step, 1, keys[g], &s, &sq);
        if (mode == 3) {
            if (q.col < 0) { free(keys); seterr(err, errcap, "SQL error: near \"*\": syntax error"); return ORC_SQL_RUNTIME_ERROR; }
            if (s.cnt == 0) { free(keys); seterr(err, errcap, "stod"); return ORC_SQL_STOD; } /* the reference std::terminate()s here */
            if (s.cnt >= 2) {
                rc = ci_value(&q, &s, &sq, sample_percent, &r, err, errcap);
                if (rc) { free(keys); return rc; }
                if (m < cap) out[m] = r;
                ++m;
                continue;
            }
        }
        rc = plain_value(&q, &s, step, sample_percent, &r.value, err, errcap);
        if (rc) { free(keys); return rc; }
        r.ci_lower = r.ci_upper = r.value;
        if (m < cap) out[m] = r;
        ++m;
    }
    free(keys);
    *n_out = m;
    return ORC_SQL_OK;
}
