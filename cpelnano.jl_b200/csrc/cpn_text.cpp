// cpn_text.cpp — output text of the path, native (host only).
//
// The reference writes its result files with Julia string interpolation (src/input_output.jl:95-462): every Float64 is printed
// by Base.show → shortest round-trip digits (Grisu/Ryu), positional when the decimal point position pt satisfies -4 < pt <= 6
// (i.e. 1e-5 <= |x| < 1e6: "0.0001", "123456.7", "100000.0"), otherwise d.ddde±x ("1.0e-5", "1.234567e6"); NaN, Inf, -Inf, -0.0.
// round(x, digits=4) is round-half-even of x·10⁴ divided by 10⁴ (Base/floatfuncs.jl).  std::to_chars gives the same shortest
// digit string; this file only re-arranges it.  Pinned by re-printing every number of the reference's golden files and by
// re-writing the golden theta / mml / nme files byte for byte (tests/test_outputs.py, tests/test_pipeline_gpu.py).
//
//   write_output_ex / _exx / _mml / _nme, write_output_tmml / _tnme / _tcmd   → cpn_write_rows
//   write_output_ϕ (with get_αβ_from_ϕ, src/link_functions.jl:16-25)          → cpn_write_theta
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <thread>
#include <vector>

#include "../../include/cpelnano_b200.h"

namespace {

// appends Julia's text of x to s
inline void put_f64(std::string& s, double x) {
    if (x != x) { s += "NaN"; return; }
    if (std::isinf(x)) { s += x > 0 ? "Inf" : "-Inf"; return; }
    if (x == 0.0) { s += std::signbit(x) ? "-0.0" : "0.0"; return; }
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(x), std::chars_format::scientific);      // d[.ddd]e±XX, shortest digits
    const char* e = (const char*)memchr(buf, 'e', r.ptr - buf);
    char digits[24];
    int n = 0;
    for (const char* p = buf; p < e; ++p)
        if (*p != '.') digits[n++] = *p;
    int ex = 0;
    std::from_chars(e + 1 + (e[1] == '+'), r.ptr, ex);
    const int pt = ex + 1;                                    // |x| = 0.DIGITS × 10^pt
    if (x < 0) s += '-';
    if (pt <= -4 || pt > 6) {
        s += digits[0];
        s += '.';
        if (n > 1) s.append(digits + 1, n - 1); else s += '0';
        s += 'e';
        s += std::to_string(pt - 1);
    } else if (pt <= 0) {
        s += "0.";
        s.append((size_t)(-pt), '0');
        s.append(digits, n);
    } else if (pt >= n) {
        s.append(digits, n);
        s.append((size_t)(pt - n), '0');
        s += ".0";
    } else {
        s.append(digits, pt);
        s += '.';
        s.append(digits + pt, n - pt);
    }
}

inline double round_digits(double x, int d) {
    if (d < 0) return x;
    double og = 1.0;
    for (int i = 0; i < d; ++i) og *= 10.0;
    return std::nearbyint(x * og) / og;                       // default rounding mode: to nearest, ties to even
}

inline void put_i64(std::string& s, int64_t v) {
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    s.append(buf, r.ptr - buf);
}

// formats the index range [0, n) in contiguous slices on all host threads; slices are written in order
template <class F>
bool format_parallel(const char* path, int64_t n, int64_t min_per_thread, F&& fmt) {
    int T = (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::max<int64_t>(1, std::min<int64_t>(T, n / std::max<int64_t>(1, min_per_thread)));
    std::vector<std::string> parts(T);
    if (T == 1) {
        fmt((int64_t)0, n, parts[0]);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] { fmt(n * t / T, n * (t + 1) / T, parts[t]); });
        for (auto& x : th) x.join();
    }
    FILE* f = fopen(path, "ab");
    if (!f) return false;
    bool ok = true;
    for (const std::string& s : parts) ok = ok && (s.empty() || fwrite(s.data(), 1, s.size(), f) == s.size());
    return fclose(f) == 0 && ok;
}

}  // namespace

extern "C" {

int64_t cpn_format_f64(const double* x, int64_t n, char sep, int32_t digits, char* out, int64_t cap) {
    if (n < 0 || (n > 0 && !x)) return -1;
    std::string s;
    s.reserve((size_t)n * 20);
    for (int64_t i = 0; i < n; ++i) {
        if (i) s += sep;
        put_f64(s, round_digits(x[i], digits));
    }
    if (out && cap >= (int64_t)s.size()) memcpy(out, s.data(), s.size());
    return (int64_t)s.size();
}

int cpn_write_rows(const char* path, const char* chr, int64_t n, const int64_t* a, const int64_t* b, const double* v, int32_t digits,
                   int32_t clamp0, const double* w) {
    if (!path || !chr || n < 0 || (n > 0 && (!a || !b || !v))) return CPN_ERR_ARG;
    const bool ok = format_parallel(path, n, 20000, [&](int64_t i0, int64_t i1, std::string& s) {
        s.reserve((size_t)(i1 - i0) * 48);
        for (int64_t i = i0; i < i1; ++i) {
            if (v[i] != v[i]) continue;                        // isnan(x) && continue
            s += chr; s += '\t';
            put_i64(s, a[i]); s += '\t';
            put_i64(s, b[i]); s += '\t';
            double x = v[i];
            if (clamp0 && !(x > 0.0)) x = 0.0;                 // max(0.0, tcmd) before rounding (input_output.jl:427)
            put_f64(s, round_digits(x, digits));
            if (w) { s += '\t'; put_f64(s, w[i]); }
            s += '\n';
        }
    });
    return ok ? CPN_OK : CPN_ERR_ARG;
}

int cpn_write_theta(const char* path, const char* chr, int64_t R, const int64_t* chrst, const int64_t* chrend, const double* phi,
                    const int64_t* grp_off, const double* Nl, const double* rho_l, const double* d_l) {
    if (!path || !chr || R < 0 || (R > 0 && (!chrst || !chrend || !phi || !grp_off || !rho_l || !d_l))) return CPN_ERR_ARG;
    const bool ok = format_parallel(path, R, 500, [&](int64_t r0, int64_t r1, std::string& s) {
        for (int64_t r = r0; r < r1; ++r) {
            const double pa = phi[3 * r], pb = phi[3 * r + 1], pc = phi[3 * r + 2];
            const int64_t g0 = grp_off[r], L = grp_off[r + 1] - g0;
            s += chr; s += '\t';
            put_i64(s, chrst[r]); s += '\t';
            put_i64(s, chrend[r]); s += '\t';
            put_f64(s, pa); s += ','; put_f64(s, pb); s += ','; put_f64(s, pc); s += '\t';
            for (int64_t l = 0; l < L; ++l) {                  // α = (a .+ b .* ρ) .* N, no fused multiply-add (link_functions.jl:19)
                if (l) s += ',';
                volatile double t = pb * rho_l[g0 + l];
                volatile double u = pa + t;
                put_f64(s, Nl ? u * Nl[g0 + l] : u);
            }
            s += '\t';
            const double* d = d_l + (g0 - r);
            for (int64_t l = 0; l + 1 < L; ++l) {              // β = c ./ d
                if (l) s += ',';
                put_f64(s, pc / d[l]);
            }
            s += '\n';
        }
    });
    return ok ? CPN_OK : CPN_ERR_ARG;
}

// add_qvalues_to_path (src/multiple_hypothesis.jl:14-49): Benjamini–Hochberg q-values of the p column (5th field) appended as
// a 6th column, NaN where p is NaN; the file is re-written through readdlm / writedlm in the reference, i.e. the float fields
// are re-printed (shortest text: unchanged for text Julia wrote itself) and integer fields stay integers.
// MultipleTesting.BenjaminiHochberg: sorted p, p_(i)·(n/i), running minimum from the largest rank down, capped at 1.
int cpn_add_qvalues(const char* path) {
    if (!path) return CPN_ERR_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) return CPN_OK;                                   // nothing to do (the reference checks filesize(path) > 0)
    std::string buf;
    {
        char tmp[1 << 16];
        size_t k;
        while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, k);
    }
    fclose(f);
    if (buf.empty()) return CPN_OK;
    struct Row { const char *b, *e; double t, p; const char *f3e; };
    std::vector<Row> rows;
    const char* p0 = buf.data();
    const char* end = p0 + buf.size();
    while (p0 < end) {
        const char* nl = (const char*)memchr(p0, '\n', end - p0);
        const char* le = nl ? nl : end;
        if (le > p0) {
            const char* fld[6];
            int nf = 0;
            fld[nf++] = p0;
            for (const char* q = p0; q < le && nf < 6; ++q)
                if (*q == '\t') fld[nf++] = q + 1;
            if (nf < 5) return CPN_ERR_ARG;
            auto parse = [&](const char* b, const char* e, double& v) {
                std::string_view sv(b, e - b);
                if (sv == "NaN") { v = std::nan(""); return true; }
                if (sv == "Inf") { v = INFINITY; return true; }
                if (sv == "-Inf") { v = -INFINITY; return true; }
                auto r = std::from_chars(b, e, v);
                return r.ec == std::errc() && r.ptr == e;
            };
            Row r;
            r.b = p0; r.e = le; r.f3e = fld[3] - 1;
            const char* pe = (nf > 5) ? fld[5] - 1 : le;
            if (!parse(fld[3], fld[4] - 1, r.t) || !parse(fld[4], pe, r.p)) return CPN_ERR_ARG;
            rows.push_back(r);
        }
        p0 = nl ? nl + 1 : end;
    }
    const size_t n_all = rows.size();
    std::vector<size_t> ord;
    ord.reserve(n_all);
    for (size_t i = 0; i < n_all; ++i)
        if (rows[i].p == rows[i].p) ord.push_back(i);
    std::vector<double> q(n_all, std::nan(""));
    const size_t n = ord.size();
    if (n == 1) q[ord[0]] = rows[ord[0]].p;
    if (n > 1) {
        std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return rows[a].p < rows[b].p; });
        double run = INFINITY;
        for (size_t k = n; k-- > 0;) {
            const double v = rows[ord[k]].p * ((double)n / (double)(k + 1));
            run = std::min(run, v);
            q[ord[k]] = std::min(1.0, run);
        }
    }
    std::string out;
    out.reserve(buf.size() + n_all * 24);
    for (size_t i = 0; i < n_all; ++i) {
        out.append(rows[i].b, rows[i].f3e - rows[i].b);      // chr, start, end (integers are written back as integers)
        out += '\t'; put_f64(out, rows[i].t);
        out += '\t'; put_f64(out, rows[i].p);
        out += '\t'; put_f64(out, q[i]);
        out += '\n';
    }
    const std::string tmp = std::string(path) + ".tmp";
    FILE* g = fopen(tmp.c_str(), "wb");
    if (!g) return CPN_ERR_ARG;
    const bool ok = fwrite(out.data(), 1, out.size(), g) == out.size();
    if (fclose(g) != 0 || !ok) return CPN_ERR_ARG;
    return rename(tmp.c_str(), path) == 0 ? CPN_OK : CPN_ERR_ARG;
}

// Model (theta) files: chr \t chrst \t chrend \t a,b,c [\t α list \t β list] — read_mod_file_chr (src/input_output.jl:581-630) only
// uses the first four fields (everything else is recomputed from ϕ̂), and the α / β lists are kilobytes per line: fields 1-4 are
// parsed, the rest of the line is skipped with one memchr.  Rows keep file order; chr_id indexes the table of names in order of
// first appearance.
struct cpn_theta {
    std::vector<std::string> chr_names;
    std::vector<int32_t> chr_id;
    std::vector<int64_t> st, en;
    std::vector<double> phi;     // [n][3]
};
static thread_local std::string g_theta_err;
const char* cpn_theta_open_error(void) { return g_theta_err.c_str(); }

int cpn_theta_open(const char* path, cpn_theta** out) {
    if (!path || !out) { g_theta_err = "null argument"; return CPN_ERR_ARG; }
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) { g_theta_err = std::string("cannot open ") + path; return CPN_ERR_ARG; }
    std::string buf;
    {
        std::vector<char> tmp(1 << 22);
        size_t k;
        while ((k = fread(tmp.data(), 1, tmp.size(), f)) > 0) buf.append(tmp.data(), k);
    }
    fclose(f);
    cpn_theta* t = new cpn_theta();
    std::unordered_map<std::string, int32_t> ix;
    const char* p = buf.data();
    const char* end = p + buf.size();
    int64_t line_no = 0;
    auto bad = [&]() { g_theta_err = "malformed line " + std::to_string(line_no) + " in " + path; delete t; return CPN_ERR_ARG; };
    while (p < end) {
        ++line_no;
        const char* nl = (const char*)memchr(p, '\n', end - p);
        const char* le = nl ? nl : end;
        const char* line = p;
        p = nl ? nl + 1 : end;
        if (le > line && le[-1] == '\r') --le;
        if (le == line) continue;
        const char* t1 = (const char*)memchr(line, '\t', le - line);
        if (!t1) return bad();
        const char* t2 = (const char*)memchr(t1 + 1, '\t', le - t1 - 1);
        if (!t2) return bad();
        const char* t3 = (const char*)memchr(t2 + 1, '\t', le - t2 - 1);
        if (!t3) return bad();
        const char* t4 = (const char*)memchr(t3 + 1, '\t', le - t3 - 1);
        const char* pe = t4 ? t4 : le;
        int64_t a, b;
        if (std::from_chars(t1 + 1, t2, a).ptr != t2 || std::from_chars(t2 + 1, t3, b).ptr != t3) return bad();
        double ph[3];
        const char* q = t3 + 1;
        for (int k = 0; k < 3; ++k) {
            const char* c = (k < 2) ? (const char*)memchr(q, ',', pe - q) : pe;
            if (!c) return bad();
            std::string_view sv(q, c - q);
            if (sv == "NaN") ph[k] = std::nan("");
            else {
                auto r = std::from_chars(q, c, ph[k]);
                if (r.ec != std::errc() || r.ptr != c) return bad();
            }
            q = c + 1;
        }
        std::string name(line, t1 - line);
        auto it = ix.find(name);
        if (it == ix.end()) { it = ix.emplace(name, (int32_t)t->chr_names.size()).first; t->chr_names.push_back(name); }
        t->chr_id.push_back(it->second);
        t->st.push_back(a); t->en.push_back(b);
        t->phi.push_back(ph[0]); t->phi.push_back(ph[1]); t->phi.push_back(ph[2]);
    }
    *out = t;
    return CPN_OK;
}
void cpn_theta_close(cpn_theta* t) { delete t; }
int64_t cpn_theta_num_rows(const cpn_theta* t) { return t ? (int64_t)t->st.size() : 0; }
int64_t cpn_theta_num_chr(const cpn_theta* t) { return t ? (int64_t)t->chr_names.size() : 0; }
const char* cpn_theta_chr_name(const cpn_theta* t, int64_t i) { return (t && i >= 0 && i < (int64_t)t->chr_names.size()) ? t->chr_names[i].c_str() : ""; }
int cpn_theta_copy(const cpn_theta* t, int32_t* chr_id, int64_t* chrst, int64_t* chrend, double* phi) {
    if (!t) return CPN_ERR_ARG;
    const size_t n = t->st.size();
    if (n == 0) return CPN_OK;
    if (!chr_id || !chrst || !chrend || !phi) return CPN_ERR_ARG;
    memcpy(chr_id, t->chr_id.data(), n * sizeof(int32_t));
    memcpy(chrst, t->st.data(), n * sizeof(int64_t));
    memcpy(chrend, t->en.data(), n * sizeof(int64_t));
    memcpy(phi, t->phi.data(), 3 * n * sizeof(double));
    return CPN_OK;
}

}  // extern "C"
