// aqe_sql.cpp -- host half of the SQL-string path (see aqe_sql.hpp).
//
// What the reference does (src/aqe_backend/): parser.cpp:20-75 cuts `SELECT agg(col) FROM t [WHERE w]
// [GROUP BY g]` apart with substring searches on the upper-cased text; executor.cpp re-assembles SQLite
// statements from the pieces, appends `rowid % (100/p) = 0` for sampling, lets SQLite evaluate them over a
// SQLite file, reads the answer back as TEXT (15 significant digits) and scales SUM/COUNT by 100/p.
// Here the parse is the same, the WHERE text is compiled into one closed interval (+ one optional "!=" value)
// per column, and the evaluation is a grouped scan over the HBM columns.
#include "aqe_sql.hpp"

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace aqe {

namespace {

std::string upper(std::string s) {
    for (char& c : s) c = (char)std::toupper((unsigned char)c);
    return s;
}
std::string trim(const std::string& s) {  // parser.cpp:13-19
    const char* ws = " \t\n\r";
    const size_t a = s.find_first_not_of(ws);
    if (a == std::string::npos) return "";
    const size_t b = s.find_last_not_of(ws);
    return s.substr(a, b - a + 1);
}
// std::string::substr semantics with the reference's unsigned arithmetic: `len` may have wrapped around
std::string cut(const std::string& s, size_t pos, size_t len) { return pos <= s.size() ? s.substr(pos, len) : std::string(); }

void put(char* dst, size_t cap, const std::string& s) {
    const size_t n = std::min(cap - 1, s.size());
    std::memcpy(dst, s.data(), n);
    dst[n] = 0;
}

int column_of(const std::string& name) {
    const std::string u = upper(name);
    if (u == "ID" || u == "ROWID" || u == "_ROWID_" || u == "OID") return AQE_COL_ID;  // INTEGER PRIMARY KEY aliases rowid
    if (u == "AMOUNT") return AQE_COL_AMOUNT;
    if (u == "REGION") return AQE_COL_REGION;
    if (u == "PRODUCT_ID") return AQE_COL_PRODUCT_ID;
    if (u == "TIMESTAMP") return AQE_COL_TIMESTAMP;
    return AQE_COL_NONE;
}
bool is_identifier(const std::string& s) {
    if (s.empty() || !(std::isalpha((unsigned char)s[0]) || s[0] == '_')) return false;
    for (char c : s) if (!(std::isalnum((unsigned char)c) || c == '_')) return false;
    return true;
}

// ---- WHERE: tokens ---------------------------------------------------------------------------------------
enum Tok { T_END, T_IDENT, T_NUM, T_OP, T_LP, T_RP, T_AND, T_OR, T_NOT, T_BETWEEN, T_IN, T_COMMA, T_OTHER };
struct Token { Tok t; std::string text; bool is_int = false; int64_t i = 0; double d = 0.0; };

struct Lexer {
    const std::string& s;
    size_t p = 0;
    std::string err;
    explicit Lexer(const std::string& str) : s(str) {}
    bool number(const std::string& txt, Token& k) {
        if (txt.empty()) return false;
        char* end = nullptr;
        const bool integral = txt.find_first_of(".eE") == std::string::npos;
        if (integral) {
            errno = 0;
            const long long v = std::strtoll(txt.c_str(), &end, 10);
            if (*end == 0 && errno == 0) { k.is_int = true; k.i = v; k.d = (double)v; return true; }
        }
        const double d = std::strtod(txt.c_str(), &end);
        if (*end != 0 || !(d == d)) return false;
        k.is_int = false; k.d = d;
        return true;
    }
    Token next() {
        while (p < s.size() && std::isspace((unsigned char)s[p])) ++p;
        Token k{T_END, ""};
        if (p >= s.size()) return k;
        const char c = s[p];
        if (std::isalpha((unsigned char)c) || c == '_') {
            size_t e = p;
            while (e < s.size() && (std::isalnum((unsigned char)s[e]) || s[e] == '_')) ++e;
            k.text = s.substr(p, e - p); p = e;
            const std::string u = upper(k.text);
            k.t = u == "AND" ? T_AND : u == "OR" ? T_OR : u == "BETWEEN" ? T_BETWEEN : u == "IN" ? T_IN : u == "NOT" ? T_NOT : ((u == "LIKE" || u == "IS" || u == "NULL") ? T_OTHER : T_IDENT);
            return k;
        }
        if (std::isdigit((unsigned char)c) || (c == '.' && p + 1 < s.size() && std::isdigit((unsigned char)s[p + 1]))) {
            size_t e = p;
            while (e < s.size() && (std::isdigit((unsigned char)s[e]) || s[e] == '.')) ++e;
            if (e < s.size() && (s[e] == 'e' || s[e] == 'E')) {
                size_t f = e + 1;
                if (f < s.size() && (s[f] == '+' || s[f] == '-')) ++f;
                if (f < s.size() && std::isdigit((unsigned char)s[f])) { while (f < s.size() && std::isdigit((unsigned char)s[f])) ++f; e = f; }
            }
            k.text = s.substr(p, e - p); p = e;
            k.t = number(k.text, k) ? T_NUM : T_OTHER;
            return k;
        }
        if (c == '\'') {  // a quoted literal: numeric text takes the column's numeric affinity in SQLite
            const size_t e = s.find('\'', p + 1);
            if (e == std::string::npos) { k.t = T_OTHER; k.text = s.substr(p); p = s.size(); return k; }
            k.text = trim(s.substr(p + 1, e - p - 1)); p = e + 1;
            k.t = number(k.text, k) ? T_NUM : T_OTHER;
            return k;
        }
        if (c == '(') { ++p; k.t = T_LP; k.text = "("; return k; }
        if (c == ')') { ++p; k.t = T_RP; k.text = ")"; return k; }
        if (c == ',') { ++p; k.t = T_COMMA; k.text = ","; return k; }
        static const char* ops[] = {"<=", ">=", "<>", "!=", "==", "<", ">", "="};
        for (const char* o : ops) {
            const size_t n = std::strlen(o);
            if (s.compare(p, n, o) == 0) { p += n; k.t = T_OP; k.text = o; return k; }
        }
        if (c == '-' || c == '+') {  // signed literal
            size_t q = p + 1;
            while (q < s.size() && std::isspace((unsigned char)s[q])) ++q;
            if (q < s.size() && (std::isdigit((unsigned char)s[q]) || s[q] == '.')) {
                const size_t save = p;
                p = q;
                Token n = next();
                if (n.t == T_NUM) {
                    if (c == '-') { if (n.is_int) { n.i = -n.i; n.d = (double)n.i; } else n.d = -n.d; n.text = "-" + n.text; }
                    return n;
                }
                p = save;
            }
        }
        k.t = T_OTHER; k.text = std::string(1, c); ++p;
        return k;
    }
};

// ---- WHERE: constraints ----------------------------------------------------------------------------------
struct ColRange {
    bool touched = false;
    double lo = -std::numeric_limits<double>::infinity(), hi = std::numeric_limits<double>::infinity();
    int64_t ilo = INT64_MIN, ihi = INT64_MAX;
    bool has_ne = false;
    double ne = 0.0;
    int64_t ine = 0;
    bool empty = false;
};
bool is_f64_col(int col) { return col == AQE_COL_AMOUNT; }

// integer bounds implied by comparing an INTEGER column with a REAL literal (SQLite compares exactly)
int64_t sat_ceil(double r) {
    if (r <= -9223372036854775808.0) return INT64_MIN;
    if (r >= 9223372036854775808.0) return INT64_MAX;
    return (int64_t)std::ceil(r);
}
int64_t sat_floor(double r) {
    if (r < -9223372036854775808.0) return INT64_MIN;
    if (r >= 9223372036854775808.0) return INT64_MAX;
    return (int64_t)std::floor(r);
}

enum Cmp { C_EQ, C_NE, C_LT, C_LE, C_GT, C_GE };
bool cmp_of(const std::string& op, Cmp& c) {
    if (op == "=" || op == "==") c = C_EQ; else if (op == "!=" || op == "<>") c = C_NE; else if (op == "<") c = C_LT;
    else if (op == "<=") c = C_LE; else if (op == ">") c = C_GT; else if (op == ">=") c = C_GE; else return false;
    return true;
}
Cmp flip(Cmp c) { return c == C_LT ? C_GT : c == C_LE ? C_GE : c == C_GT ? C_LT : c == C_GE ? C_LE : c; }

bool constrain(ColRange& r, int col, Cmp c, const Token& lit, std::string& err) {
    r.touched = true;
    if (is_f64_col(col)) {
        const double v = lit.d;
        const double inf = std::numeric_limits<double>::infinity();
        switch (c) {
            case C_EQ: r.lo = std::max(r.lo, v); r.hi = std::min(r.hi, v); break;
            case C_LE: r.hi = std::min(r.hi, v); break;
            case C_GE: r.lo = std::max(r.lo, v); break;
            case C_LT: r.hi = std::min(r.hi, std::nextafter(v, -inf)); break;
            case C_GT: r.lo = std::max(r.lo, std::nextafter(v, inf)); break;
            case C_NE:
                if (r.has_ne && r.ne != v) { err = "unsupported WHERE: more than one != on the same column"; return false; }
                r.has_ne = true; r.ne = v; break;
        }
        return true;
    }
    if (lit.is_int) {
        const int64_t v = lit.i;
        switch (c) {
            case C_EQ: r.ilo = std::max(r.ilo, v); r.ihi = std::min(r.ihi, v); break;
            case C_LE: r.ihi = std::min(r.ihi, v); break;
            case C_GE: r.ilo = std::max(r.ilo, v); break;
            case C_LT: if (v == INT64_MIN) r.empty = true; else r.ihi = std::min(r.ihi, v - 1); break;
            case C_GT: if (v == INT64_MAX) r.empty = true; else r.ilo = std::max(r.ilo, v + 1); break;
            case C_NE:
                if (r.has_ne && r.ine != v) { err = "unsupported WHERE: more than one != on the same column"; return false; }
                r.has_ne = true; r.ine = v; break;
        }
        return true;
    }
    const double v = lit.d;
    const bool integral = v == std::floor(v) && v >= -9223372036854775808.0 && v < 9223372036854775808.0;
    switch (c) {
        case C_EQ:
            if (!integral) r.empty = true;
            else { r.ilo = std::max(r.ilo, (int64_t)v); r.ihi = std::min(r.ihi, (int64_t)v); }
            break;
        case C_LE: if (v < -9223372036854775808.0) r.empty = true; else r.ihi = std::min(r.ihi, sat_floor(v)); break;
        case C_GE: if (v >= 9223372036854775808.0) r.empty = true; else r.ilo = std::max(r.ilo, sat_ceil(v)); break;
        case C_LT: {
            if (v <= -9223372036854775808.0) { r.empty = true; break; }
            const int64_t b = sat_ceil(v);  // col < v  <=>  col <= ceil(v) - 1 (v integral: v - 1)
            if (v >= 9223372036854775808.0) break;
            if (b == INT64_MIN) r.empty = true; else r.ihi = std::min(r.ihi, b - 1);
            break;
        }
        case C_GT: {
            if (v >= 9223372036854775808.0) { r.empty = true; break; }
            if (v < -9223372036854775808.0) break;
            const int64_t b = sat_floor(v);
            if (b == INT64_MAX) r.empty = true; else r.ilo = std::max(r.ilo, b + 1);
            break;
        }
        case C_NE:
            if (integral) {
                if (r.has_ne && r.ine != (int64_t)v) { err = "unsupported WHERE: more than one != on the same column"; return false; }
                r.has_ne = true; r.ine = (int64_t)v;
            }
            break;
    }
    return true;
}

int cmp_literals(const Token& a, const Token& b) {
    if (a.is_int && b.is_int) return a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);
    return a.d < b.d ? -1 : (a.d > b.d ? 1 : 0);
}

// A conjunction: one range per column.  A WHERE clause is kept in disjunctive normal form, an OR of conjunctions.
struct Conj {
    ColRange r[5];
    bool dead = false;  // unsatisfiable
    bool is_true() const {
        if (dead) return false;
        for (const ColRange& c : r) if (c.touched) return false;
        return true;
    }
};
using Dnf = std::vector<Conj>;  // empty = false

bool range_empty(const ColRange& r, int col) {
    if (r.empty) return true;
    if (is_f64_col(col) ? !(r.lo <= r.hi) : r.ilo > r.ihi) return true;
    if (r.has_ne) return is_f64_col(col) ? (r.lo == r.hi && r.ne == r.lo) : (r.ilo == r.ihi && r.ine == r.ilo);
    return false;
}

// a AND b, appended to `out` (nothing when unsatisfiable).  A scan tests one "!=" value per column and branch: when both
// sides exclude a different value of one column, the range is cut at the second value into the part below and the part
// above it -- two branches, each with one "!=" left (x != 1 AND x != 3  ==  (x < 3 AND x != 1) OR x > 3).
void conj_and(const Conj& a, const Conj& b, Dnf& out) {
    if (a.dead || b.dead) return;
    Dnf work(1, a);
    for (int c = 0; c < 5; ++c) {
        const ColRange& y = b.r[c];
        if (!y.touched) continue;
        const bool f = is_f64_col(c);
        Dnf next;
        for (Conj& k : work) {
            ColRange& x = k.r[c];
            if (!x.touched) { x = y; }
            else {
                x.lo = std::max(x.lo, y.lo); x.hi = std::min(x.hi, y.hi);
                x.ilo = std::max(x.ilo, y.ilo); x.ihi = std::min(x.ihi, y.ihi);
                x.empty = x.empty || y.empty;
                if (y.has_ne && x.has_ne && (f ? x.ne != y.ne : x.ine != y.ine)) {
                    Conj below = k, above = k;
                    ColRange& lo = below.r[c];
                    ColRange& hi = above.r[c];
                    if (f) {
                        lo.hi = std::min(lo.hi, std::nextafter(y.ne, -std::numeric_limits<double>::infinity()));
                        hi.lo = std::max(hi.lo, std::nextafter(y.ne, std::numeric_limits<double>::infinity()));
                    } else {
                        if (y.ine == INT64_MIN) lo.empty = true; else lo.ihi = std::min(lo.ihi, y.ine - 1);
                        if (y.ine == INT64_MAX) hi.empty = true; else hi.ilo = std::max(hi.ilo, y.ine + 1);
                    }
                    if (!range_empty(lo, c)) next.push_back(below);
                    if (!range_empty(hi, c)) next.push_back(above);
                    continue;
                }
                if (y.has_ne) { x.has_ne = true; x.ne = y.ne; x.ine = y.ine; }
            }
            if (!range_empty(x, c)) next.push_back(k);
        }
        work.swap(next);
    }
    out.insert(out.end(), work.begin(), work.end());
}

// NOT of one column range: the values outside [lo, hi], plus the excluded value itself -- up to three single-range branches.
void negate_range(const ColRange& r, int col, Dnf& out) {
    const bool f = is_f64_col(col);
    const double inf = std::numeric_limits<double>::infinity();
    auto push = [&](const ColRange& x) { if (!range_empty(x, col)) { Conj k; k.r[col] = x; k.r[col].touched = true; out.push_back(k); } };
    if (r.empty || (f ? !(r.lo <= r.hi) : r.ilo > r.ihi)) { out.push_back(Conj()); return; }  // NOT false = true
    ColRange below, above;
    if (f) {
        below.hi = std::nextafter(r.lo, -inf); if (std::isinf(r.lo) && r.lo < 0) below.empty = true;
        above.lo = std::nextafter(r.hi, inf); if (std::isinf(r.hi) && r.hi > 0) above.empty = true;
    } else {
        if (r.ilo == INT64_MIN) below.empty = true; else below.ihi = r.ilo - 1;
        if (r.ihi == INT64_MAX) above.empty = true; else above.ilo = r.ihi + 1;
    }
    push(below); push(above);
    if (r.has_ne && (f ? (r.ne >= r.lo && r.ne <= r.hi) : (r.ine >= r.ilo && r.ine <= r.ihi))) {
        ColRange eq;
        if (f) { eq.lo = eq.hi = r.ne; } else { eq.ilo = eq.ihi = r.ine; }
        push(eq);
    }
}

struct WhereCompiler {
    Lexer lx;
    Token cur;
    std::string err;
    int status = AQE_OK;
    static constexpr size_t kMaxWork = 32;  // conjunctions alive while distributing AND over OR
    int depth = 0;
    bool top_level_or = false;
    explicit WhereCompiler(const std::string& w) : lx(w) { cur = lx.next(); }
    void advance() { cur = lx.next(); }
    bool unsupported(const std::string& what) {
        status = AQE_ERR_UNSUPPORTED;
        err = "unsupported WHERE clause (" + what + "); supported: comparisons, BETWEEN and IN lists of id|rowid|amount|region|product_id|timestamp with numeric literals, combined with AND, OR, NOT and parentheses";
        return false;
    }
    bool fail_unsupported() { status = AQE_ERR_UNSUPPORTED; return false; }
    bool operand(Token& out, int& col) {
        if (cur.t == T_IDENT) {
            col = column_of(cur.text);
            if (col == AQE_COL_NONE) { status = AQE_ERR_INVALID; err = "SQL error: no such column: " + cur.text; return false; }
            out = cur; advance();
            return true;
        }
        if (cur.t == T_NUM) { col = AQE_COL_NONE; out = cur; advance(); return true; }
        return unsupported(cur.t == T_END ? "unexpected end" : "near \"" + cur.text + "\"");
    }
    static Dnf constant(bool v) { Dnf d; if (v) d.push_back(Conj()); return d; }
    bool single(int col, Cmp c, const Token& lit, Dnf& out) {
        Conj k;
        if (!constrain(k.r[col], col, c, lit, err)) return fail_unsupported();
        if (range_empty(k.r[col], col)) k.dead = true;
        out.clear();
        if (!k.dead) out.push_back(k);
        return true;
    }
    // NOT (C1 OR C2 ...) = NOT C1 AND NOT C2 ...;  NOT Ci = OR over its columns of NOT(range)
    bool negate(const Dnf& in, Dnf& out) {
        out = constant(true);
        for (const Conj& k : in) {
            Dnf nk;
            if (k.dead) nk = constant(true);
            else for (int c = 0; c < 5; ++c) if (k.r[c].touched) negate_range(k.r[c], c, nk);
            // a conjunction without constraints is "true": its negation is false (nk stays empty)
            Dnf prod;
            for (const Conj& x : out)
                for (const Conj& y : nk) {
                    conj_and(x, y, prod);
                    if (prod.size() > kMaxWork) return unsupported("too many OR branches");
                }
            if (prod.size() > kMaxWork) return unsupported("too many OR branches");
            out.swap(prod);
        }
        return true;
    }
    bool term(Dnf& out) {
        if (cur.t == T_NOT) {
            advance();
            Dnf inner;
            if (!term(inner)) return false;
            return negate(inner, out);
        }
        if (cur.t == T_LP) {
            advance();
            ++depth;
            if (!disjunction(out)) return false;
            --depth;
            if (cur.t != T_RP) return unsupported("missing )");
            advance();
            return true;
        }
        Token a; int ca;
        if (!operand(a, ca)) return false;
        if (cur.t == T_NOT) {  // col NOT BETWEEN ... / col NOT IN (...)
            advance();
            if (cur.t != T_BETWEEN && cur.t != T_IN) return unsupported("NOT must be followed by BETWEEN or IN here");
            Dnf inner;
            if (!tail(a, ca, inner)) return false;
            return negate(inner, out);
        }
        return tail(a, ca, out);
    }
    bool tail(const Token& a, int ca, Dnf& out) {  // what follows the first operand of a comparison
        if (cur.t == T_BETWEEN) {
            advance();
            Token lo, hi; int cl, ch;
            if (!operand(lo, cl)) return false;
            if (cur.t != T_AND) return unsupported("BETWEEN without AND");
            advance();
            if (!operand(hi, ch)) return false;
            if (cl != AQE_COL_NONE || ch != AQE_COL_NONE) return unsupported("BETWEEN bounds must be literals");
            if (ca == AQE_COL_NONE) { out = constant(cmp_literals(a, lo) >= 0 && cmp_literals(a, hi) <= 0); return true; }
            Conj k;
            if (!constrain(k.r[ca], ca, C_GE, lo, err) || !constrain(k.r[ca], ca, C_LE, hi, err)) return fail_unsupported();
            out.clear();
            if (!range_empty(k.r[ca], ca)) out.push_back(k);
            return true;
        }
        if (cur.t == T_IN) {  // col IN (v1, v2, ...)  ==  col = v1 OR col = v2 OR ...
            advance();
            if (ca == AQE_COL_NONE) return unsupported("IN needs a column on the left");
            if (cur.t != T_LP) return unsupported("IN without a list");
            advance();
            out.clear();
            for (;;) {
                Token v; int cv;
                if (!operand(v, cv)) return false;
                if (cv != AQE_COL_NONE) return unsupported("IN list entries must be literals");
                Dnf one;
                if (!single(ca, C_EQ, v, one)) return false;
                bool dup = false;   // a repeated value adds nothing
                for (const Conj& k : out) if (!one.empty() && (is_f64_col(ca) ? k.r[ca].lo == one[0].r[ca].lo : k.r[ca].ilo == one[0].r[ca].ilo)) dup = true;
                if (!dup) out.insert(out.end(), one.begin(), one.end());
                if (out.size() > kMaxWork) return unsupported("too many OR branches");
                if (cur.t == T_COMMA) { advance(); continue; }
                break;
            }
            if (cur.t != T_RP) return unsupported("missing ) after IN list");
            advance();
            return true;
        }
        if (cur.t != T_OP) return unsupported(cur.t == T_END ? "comparison expected" : "near \"" + cur.text + "\"");
        Cmp c;
        if (!cmp_of(cur.text, c)) return unsupported("operator " + cur.text);
        advance();
        Token b; int cb;
        if (!operand(b, cb)) return false;
        if (ca != AQE_COL_NONE && cb != AQE_COL_NONE) return unsupported("column-to-column comparison");
        if (ca == AQE_COL_NONE && cb == AQE_COL_NONE) {
            const int r = cmp_literals(a, b);
            out = constant(c == C_EQ ? r == 0 : c == C_NE ? r != 0 : c == C_LT ? r < 0 : c == C_LE ? r <= 0 : c == C_GT ? r > 0 : r >= 0);
            return true;
        }
        return ca != AQE_COL_NONE ? single(ca, c, b, out) : single(cb, flip(c), a, out);
    }
    bool conjunction(Dnf& out) {  // term (AND term)*, AND distributed over the operands' ORs
        if (!term(out)) return false;
        while (cur.t == T_AND) {
            advance();
            Dnf rhs, prod;
            if (!term(rhs)) return false;
            for (const Conj& x : out)
                for (const Conj& y : rhs) {
                    conj_and(x, y, prod);
                    if (prod.size() > kMaxWork) return unsupported("too many OR branches");
                }
            if (prod.size() > kMaxWork) return unsupported("too many OR branches");
            out.swap(prod);
        }
        return true;
    }
    bool disjunction(Dnf& out) {  // conjunction (OR conjunction)*
        if (!conjunction(out)) return false;
        while (cur.t == T_OR) {
            if (depth == 0) top_level_or = true;
            advance();
            Dnf rhs;
            if (!conjunction(rhs)) return false;
            out.insert(out.end(), rhs.begin(), rhs.end());
            if (out.size() > kMaxWork) return unsupported("too many OR branches");
        }
        return true;
    }
    bool run(Dnf& out) {
        if (!disjunction(out)) return false;
        if (cur.t != T_END) return unsupported(cur.t == T_OTHER || cur.t == T_IDENT ? "near \"" + cur.text + "\"" : "trailing input");
        return true;
    }
};

}  // namespace

int sql_parse(const std::string& sql, int sample_percent, aqe_sql_query& q, std::string& err) {
    std::memset(&q, 0, sizeof(q));
    q.sample_percent = sample_percent;
    q.agg_col = AQE_COL_NONE; q.group_col = AQE_COL_NONE;
    const std::string U = upper(sql);

    // ---- parser.cpp:28-75, find for find ----
    const size_t sel = U.find("SELECT"), from = U.find("FROM");
    if (sel == std::string::npos || from == std::string::npos) { err = "Invalid SQL: missing SELECT or FROM"; return AQE_ERR_INVALID; }
    const std::string agg_col = trim(cut(sql, sel + 6, from - (sel + 6)));
    const size_t where_pos = U.find("WHERE"), group_pos = U.find("GROUP BY");
    std::string table, where, group_by;
    if (where_pos != std::string::npos) {
        table = trim(cut(sql, from + 4, where_pos - (from + 4)));
        if (group_pos != std::string::npos) {
            where = trim(cut(sql, where_pos + 5, group_pos - (where_pos + 5)));
            group_by = trim(cut(sql, group_pos + 8, std::string::npos));
        } else {
            where = trim(cut(sql, where_pos + 5, std::string::npos));
        }
    } else if (group_pos != std::string::npos) {
        table = trim(cut(sql, from + 4, group_pos - (from + 4)));
        group_by = trim(cut(sql, group_pos + 8, std::string::npos));
    } else {
        table = trim(cut(sql, from + 4, std::string::npos));
    }
    if (!table.empty() && table.back() == ';') table.pop_back();
    if (!group_by.empty() && group_by.back() == ';') group_by.pop_back();
    if (!where.empty() && where.back() == ';') where.pop_back();
    const size_t po = agg_col.find('('), pc = agg_col.find(')');
    if (po == std::string::npos || pc == std::string::npos) { err = "Invalid aggregation syntax"; return AQE_ERR_INVALID; }
    const std::string agg = trim(agg_col.substr(0, po));
    const std::string column = trim(cut(agg_col, po + 1, pc - po - 1));
    const std::string aggU = upper(agg);
    if (aggU != "SUM" && aggU != "COUNT" && aggU != "AVG") {
        err = "Unsupported aggregation function: " + agg + ". Supported functions: SUM, COUNT, AVG";
        return AQE_ERR_INVALID;
    }
    q.agg = aggU == "SUM" ? AQE_AGG_SUM : (aggU == "AVG" ? AQE_AGG_AVG : AQE_AGG_COUNT);
    put(q.agg_text, sizeof(q.agg_text), agg);
    put(q.column, sizeof(q.column), column);
    put(q.table, sizeof(q.table), table);
    put(q.group_by, sizeof(q.group_by), group_by);
    put(q.where, sizeof(q.where), where);
    if (where.size() >= sizeof(q.where)) { err = "WHERE clause longer than 511 characters"; return AQE_ERR_UNSUPPORTED; }

    // ---- resolution against the record table (the part SQLite does for the reference) ----
    table = trim(table);
    if (!is_identifier(table)) { err = "unsupported FROM clause \"" + table + "\": one table name expected"; return AQE_ERR_UNSUPPORTED; }
    if (column == "*") {
        if (q.agg != AQE_AGG_COUNT) { err = "SQL error: wrong number of arguments to function " + agg + "()"; return AQE_ERR_INVALID; }
        q.agg_col = AQE_COL_NONE;
    } else {
        if (!is_identifier(column)) { err = "unsupported aggregate argument \"" + column + "\": a column name or * expected"; return AQE_ERR_UNSUPPORTED; }
        q.agg_col = column_of(column);
        if (q.agg_col == AQE_COL_NONE) { err = "SQL error: no such column: " + column; return AQE_ERR_INVALID; }
    }
    group_by = trim(group_by);
    if (!group_by.empty()) {
        if (!is_identifier(group_by)) { err = "unsupported GROUP BY \"" + group_by + "\": one column name expected"; return AQE_ERR_UNSUPPORTED; }
        q.group_col = column_of(group_by);
        if (q.group_col == AQE_COL_NONE) { err = "SQL error: no such column: " + group_by; return AQE_ERR_INVALID; }
        if (is_f64_col(q.group_col)) { err = "unsupported GROUP BY on a floating-point column"; return AQE_ERR_UNSUPPORTED; }
    }
    if (!where.empty()) {
        WhereCompiler wc(where);
        Dnf dnf;
        if (!wc.run(dnf)) { err = wc.err.empty() ? "unsupported WHERE clause" : wc.err; return wc.status ? wc.status : AQE_ERR_UNSUPPORTED; }
        q.top_level_or = wc.top_level_or ? 1 : 0;
        bool any_true = false;
        for (const Conj& k : dnf) any_true = any_true || k.is_true();
        if (dnf.empty()) q.always_false = 1;
        else if (!any_true) {   // a branch without constraints makes the whole clause true: no terms at all
            if (dnf.size() > AQE_SQL_MAX_ALT) {
                err = "unsupported WHERE clause: more than " + std::to_string(AQE_SQL_MAX_ALT) + " OR branches after expansion";
                return AQE_ERR_UNSUPPORTED;
            }
            for (const Conj& k : dnf) {
                const int alt = q.n_alt++;
                for (int c = 0; c < 5; ++c) {
                    ColRange r = k.r[c];
                    if (!r.touched) continue;
                    const bool f = is_f64_col(c);
                    if (r.has_ne && !(f ? (r.ne >= r.lo && r.ne <= r.hi) : (r.ine >= r.ilo && r.ine <= r.ihi))) r.has_ne = false;  // vacuous
                    const bool trivial = f ? (std::isinf(r.lo) && r.lo < 0 && std::isinf(r.hi) && r.hi > 0 && !r.has_ne)
                                           : (r.ilo == INT64_MIN && r.ihi == INT64_MAX && !r.has_ne);
                    if (trivial) continue;
                    aqe_sql_term& t = q.terms[alt][q.n_terms[alt]++];
                    t.col = c; t.has_ne = r.has_ne ? 1 : 0;
                    t.lo = r.lo; t.hi = r.hi; t.ilo = r.ilo; t.ihi = r.ihi; t.ne = r.ne; t.ine = r.ine;
                }
                if (q.n_terms[alt] == 0) { q.n_alt = 0; std::memset(q.n_terms, 0, sizeof(q.n_terms)); break; }  // this branch is "true" after all
            }
        }
    }
    return AQE_OK;
}

int sql_shifts(double absmax, bool is_integer, int& sum_shift, int& sq_shift, std::string& err) {
    if (!(absmax == absmax) || std::isinf(absmax)) { err = "aggregate column holds non-finite values"; return AQE_ERR_UNSUPPORTED; }
    int E = 0;  // absmax < 2^E
    if (absmax > 0.0) { (void)std::frexp(absmax, &E); }
    if (E > 500) { err = "aggregate column magnitude above 2^500: squares would overflow"; return AQE_ERR_UNSUPPORTED; }
    sum_shift = is_integer ? 0 : std::min(62 - E, 1000);
    sq_shift = std::min(62 - 2 * E, 1000);
    return AQE_OK;
}

int sql_layout(const aqe_sql_query& q, const aqe_sql_facts* facts, int n, aqe_sql_layout& out, std::string& err) {
    std::memset(&out, 0, sizeof(out));
    if (q.top_level_or && (q.group_col != AQE_COL_NONE || sql_sample_step(q.sample_percent) > 0)) {
        // executor.cpp pastes `group = 'k' AND ` before and ` AND rowid % step = 0` after the clause TEXT: with an OR outside
        // parentheses SQLite binds those to the first / last branch only.  Not reproduced; the caller is told.
        err = "WHERE clause with an OR outside parentheses in a sampled or grouped query: the reference's pasted filters would bind to "
              "one branch only; write WHERE (a OR b)";
        return AQE_ERR_UNSUPPORTED;
    }
    double absmax = 0.0;
    bool any = false, is_int = false;
    int64_t kmin = 0, kmax = 0;
    for (int i = 0; i < n; ++i) {
        const aqe_sql_facts& f = facts[i];
        if (!(f.agg_absmax <= absmax)) absmax = f.agg_absmax;  // NaN propagates
        is_int = is_int || f.agg_is_integer != 0;
        if (f.key_min > f.key_max) continue;  // shard without rows
        if (!any) { kmin = f.key_min; kmax = f.key_max; any = true; }
        else { kmin = std::min(kmin, f.key_min); kmax = std::max(kmax, f.key_max); }
    }
    if (q.group_col == AQE_COL_NONE || !any) { kmin = 0; kmax = 0; }
    const unsigned __int128 span = (unsigned __int128)((__int128)kmax - (__int128)kmin) + 1;
    if (span > AQE_SQL_MAX_GROUPS) {
        err = "GROUP BY " + std::string(q.group_by) + ": key range wider than " + std::to_string(AQE_SQL_MAX_GROUPS) + " values";
        return AQE_ERR_UNSUPPORTED;
    }
    out.key_min = kmin; out.n_groups = (uint32_t)span; out.is_integer = is_int ? 1 : 0;
    // A WHERE clause that bounds the aggregate column itself in every branch bounds the values that can reach an
    // accumulator: scale for that bound, not for the column's, so `amount < 1e-5` on a column reaching 1e6 keeps
    // its 62 bits.  (Rows outside are converted too, saturate, and are masked before the add.)
    if (!is_int && q.n_alt > 0 && !q.always_false && absmax == absmax) {
        double bound = 0.0;
        for (int alt = 0; alt < q.n_alt && bound < absmax; ++alt) {
            double b = INFINITY;
            for (int t = 0; t < q.n_terms[alt]; ++t) {
                const aqe_sql_term& term = q.terms[alt][t];
                if (term.col == q.agg_col) b = std::max(std::fabs(term.lo), std::fabs(term.hi));
            }
            bound = std::max(bound, b);
        }
        if (bound < absmax) absmax = bound;
    }
    return sql_shifts(absmax, is_int, out.sum_shift, out.sq_shift, err);
}

void sql_merge(uint64_t* acc, const uint64_t* other, uint32_t n_groups) {
    for (uint32_t g = 0; g < n_groups; ++g) {
        uint64_t* a = acc + (size_t)g * 5;
        const uint64_t* b = other + (size_t)g * 5;
        a[0] += b[0];
        for (int k = 1; k <= 3; k += 2) {
            const unsigned __int128 x = ((unsigned __int128)a[k + 1] << 64) | a[k], y = ((unsigned __int128)b[k + 1] << 64) | b[k];
            const unsigned __int128 s = x + y;
            a[k] = (uint64_t)s; a[k + 1] = (uint64_t)(s >> 64);
        }
    }
}

namespace {
struct GroupAcc { uint64_t count; __int128 sum, sq; };
GroupAcc group_of(const uint64_t* acc, uint32_t g) {
    const uint64_t* a = acc + (size_t)g * 5;
    GroupAcc r;
    r.count = a[0];
    r.sum = (__int128)(((unsigned __int128)a[2] << 64) | a[1]);
    r.sq = (__int128)(((unsigned __int128)a[4] << 64) | a[3]);
    return r;
}
}  // namespace

int sql_finish(const aqe_sql_query& q, int mode, const aqe_sql_layout& L, const uint64_t* acc, const uint64_t* exists,
               aqe_sql_row* rows, uint32_t cap, uint32_t* n_rows, std::string& err) {
    if (mode < AQE_SQL_VALUE || mode > AQE_SQL_CI_CORRECT) { err = "bad mode"; return AQE_ERR_INVALID; }
    const bool grouped = q.group_col != AQE_COL_NONE;
    if (grouped && mode == AQE_SQL_CI_REFERENCE && q.agg_col == AQE_COL_NONE) {
        // the grouped statistics statement is `SELECT COUNT(*), SUM(*), SUM(* * *)` (executor.cpp:283-286)
        err = "SQL error: near \"*\": syntax error";
        return AQE_ERR_INVALID;
    }
    const int p = q.sample_percent;
    const int step = sql_sample_step(p);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    uint32_t out = 0;
    for (uint32_t g = 0; g < L.n_groups; ++g) {
        const GroupAcc a = group_of(acc, g);
        if (grouped) {
            // executor.cpp:68-79: the groups are those with a row passing WHERE, sampled or not
            const bool present = exists ? exists[(size_t)g * 5] > 0 : a.count > 0;
            if (!present) continue;
        }
        aqe_sql_row r;
        std::memset(&r, 0, sizeof(r));
        r.key = grouped ? L.key_min + (int64_t)g : 0;
        r.count = a.count;
        const double cnt = (double)a.count;
        // SQLite hands sums back as REAL (f64 column) or as an exact INTEGER (integer column); either way the
        // reference sees the nearest double
        const double sum = L.is_integer ? (double)a.sum : std::ldexp((double)a.sum, -L.sum_shift);
        const double sq = std::ldexp((double)a.sq, -L.sq_shift);
        r.sum = sum; r.sumsq = sq;
        r.isum_lo = L.is_integer ? (uint64_t)a.sum : 0; r.isum_hi = L.is_integer ? (int64_t)(a.sum >> 64) : 0;

        // the plain value: executor.cpp:28-58 (single) / :92-110 (per group)
        double value;
        bool null = false;
        if (q.agg == AQE_AGG_COUNT) value = cnt;
        else if (a.count == 0) { null = true; value = nan; }
        else value = q.agg == AQE_AGG_SUM ? sum : sum / cnt;
        if (!null && step > 0 && q.agg != AQE_AGG_AVG) value = value * (100.0 / p);
        double v = value, lo = value, hi = value;

        if (mode == AQE_SQL_CI_REFERENCE) {
            const bool with_stats = grouped ? true : (step > 0 && (q.agg == AQE_AGG_SUM || q.agg == AQE_AGG_AVG));  // :183-187 vs :262
            if (with_stats) {
                if (a.count == 0) { null = true; v = lo = hi = nan; }  // SUM(col) is NULL -> std::stod("NULL") throws
                else if (a.count >= 2) {                               // :212-241 / :299-318
                    null = false;
                    double mean = sum / cnt;
                    const double variance = (sq - (sum * sum / cnt)) / (cnt - 1);
                    const double std_error = std::sqrt(variance / cnt);
                    double margin = 1.96 * std_error;
                    if (q.agg == AQE_AGG_SUM) {
                        const double scale_factor = 100.0 / p;  // p = 0 in the grouped form gives inf, as in the reference
                        mean *= scale_factor; margin *= scale_factor;
                    }
                    v = mean; lo = mean - margin; hi = mean + margin;
                }
            }
        } else if (mode == AQE_SQL_CI_CORRECT && !null && step > 0 && q.agg != AQE_AGG_COUNT && a.count >= 2) {
            const double variance = std::max(0.0, (sq - (sum * sum / cnt)) / (cnt - 1));
            const double fpc = 1.0 - 1.0 / (double)step;  // systematic 1-in-step sample of the table
            const double margin = q.agg == AQE_AGG_SUM ? 1.96 * std::sqrt(cnt * variance * fpc) * (100.0 / p)
                                                       : 1.96 * std::sqrt(variance / cnt * fpc);
            lo = value - margin; hi = value + margin;
        }
        r.value = v; r.ci_lower = lo; r.ci_upper = hi; r.is_null = null ? 1 : 0;
        if (out < cap && rows) rows[out] = r;
        ++out;
    }
    if (n_rows) *n_rows = out;
    return AQE_OK;
}

}  // namespace aqe
