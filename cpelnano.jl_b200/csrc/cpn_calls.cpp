// cpn_calls.cpp — native calls ingestion (host side of the path).
//
// Replaces, for a whole chromosome at a time,
//     get_dic_nanopolish      src/input_output.jl:536-570   (re-reads the TSV from the top for every estimation region)
//     get_calls_nanopolish!   src/nanopore.jl:172-230       (linear findfirst over the CpG groups for every call)
// by ONE multi-threaded pass over the file (mmap or zlib → newline-aligned slices → columns per chromosome) and, per batch of
// regions, a binary search on the start column plus a binary search on the group starts per call.  The output is written
// straight into the CSR emission layout cpn_fit_batch / cpn_analyze_batch take (include/cpelnano_b200.h §Conventions):
// read-major [m_r × L_r] doubles per region, NaN = group not observed in the read.
//
// Semantics kept from the reference: the first line is a header; a line belongs to a region when chr matches,
// start ≤ chrend and end ≥ chrst; its group is the one whose interval is exactly [start+1-5, end+1+5]; every read with at
// least one overlapping line is a row, even if no line matched a group; a later line for the same (read, group) overwrites an
// earlier one; mod_nse = false stores (0, -50) for the more likely state (src/CpelNano.jl:32-33).  Decimal text → double is
// std::from_chars (correctly rounded, as Julia's parse(Float64, ·)).  One documented difference: rows follow the first
// appearance of a read among the region's lines, the reference follows Julia's Dict order (build-dependent).
//
// Host-only code: no CUDA here, usable without a GPU (tests/test_calls_native.py runs on CPU).
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "../../include/cpelnano_b200.h"

namespace {

struct ChrCols {
    std::string name;
    std::vector<int64_t> start, end;
    std::vector<int32_t> read;
    std::vector<double> lm, lu;
    int64_t max_span = 0;
};

struct Rec { int32_t chr, read; int64_t start, end; double lm, lu; };

struct Slice {
    std::vector<Rec> recs;
    std::vector<std::string_view> chr_names, read_names;
    std::unordered_map<std::string_view, int32_t> chr_ix, read_ix;
    int64_t bad_line = -1;       // offset (bytes) of the first malformed line, -1 = none
};

inline const char* next_tab(const char* p, const char* e) {
    const char* t = (const char*)memchr(p, '\t', e - p);
    return t ? t : e;
}

template <class T>
inline bool parse_num(const char* b, const char* e, T& v) {
    if (b < e && *b == '+') ++b;
    auto r = std::from_chars(b, e, v);
    return r.ec == std::errc() && r.ptr == e;
}
inline bool parse_f64(const char* b, const char* e, double& v) {
    if (parse_num(b, e, v)) return true;
    std::string_view s(b, e - b);                      // NaN / Inf spellings from_chars may refuse
    if (s == "NaN" || s == "nan") { v = std::numeric_limits<double>::quiet_NaN(); return true; }
    if (s == "Inf" || s == "inf") { v = INFINITY; return true; }
    if (s == "-Inf" || s == "-inf") { v = -INFINITY; return true; }
    return false;
}

void parse_slice(const char* buf, size_t b, size_t e, Slice& out) {
    const char* p = buf + b;
    const char* end = buf + e;
    out.recs.reserve((e - b) / 90 + 16);
    struct RCache { std::string_view name; int32_t id = -1; };
    RCache cache[256];
    std::string_view last_chr;
    int32_t last_chr_id = -1;
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', end - p);
        const char* le = nl ? nl : end;
        const char* line = p;
        p = nl ? nl + 1 : end;
        if (le > line && le[-1] == '\r') --le;
        if (le == line) continue;
        const char* f[9];
        f[0] = line;
        int nf = 1;
        for (const char* q = line; nf < 9;) {
            const char* t = next_tab(q, le);
            if (t == le) break;
            f[nf++] = t + 1;
            q = t + 1;
        }
        if (nf < 8) { if (out.bad_line < 0) out.bad_line = line - buf; continue; }
        auto fe = [&](int i) { return (i + 1 < nf) ? f[i + 1] - 1 : le; };
        Rec r;
        // A non-finite log-likelihood is a malformed line: NaN is the "group not observed" marker of the emission layout, so a NaN
        // call would silently turn into an unobserved group (the reference would carry the NaN into the EM and skip the region).
        if (!parse_num(f[2], fe(2), r.start) || !parse_num(f[3], fe(3), r.end) || !parse_f64(f[6], fe(6), r.lm) || !parse_f64(f[7], fe(7), r.lu) ||
            !std::isfinite(r.lm) || !std::isfinite(r.lu)) {
            if (out.bad_line < 0) out.bad_line = line - buf;
            continue;
        }
        std::string_view cn(f[0], fe(0) - f[0]), rn(f[4], fe(4) - f[4]);
        // consecutive lines share the chromosome, and the ~coverage reads that span a position keep alternating: a last-value
        // check and a small direct-mapped cache (keyed by 8 bytes of the name, verified by a full compare) answer almost every
        // line without touching the hash maps
        if (cn != last_chr) {
            auto ci = out.chr_ix.find(cn);
            if (ci == out.chr_ix.end()) { ci = out.chr_ix.emplace(cn, (int32_t)out.chr_names.size()).first; out.chr_names.push_back(cn); }
            last_chr = ci->first; last_chr_id = ci->second;
        }
        uint64_t key = 0;
        memcpy(&key, rn.data() + (rn.size() > 16 ? 8 : 0), std::min<size_t>(8, rn.size() > 16 ? rn.size() - 8 : rn.size()));
        RCache& rc = cache[(key * 0x9E3779B97F4A7C15ull) >> 56];
        if (rc.id < 0 || rc.name != rn) {
            auto ri = out.read_ix.find(rn);
            if (ri == out.read_ix.end()) { ri = out.read_ix.emplace(rn, (int32_t)out.read_names.size()).first; out.read_names.push_back(rn); }
            rc.name = ri->first; rc.id = ri->second;
        }
        r.chr = last_chr_id; r.read = rc.id;
        out.recs.push_back(r);
    }
}

bool read_gz(const char* path, std::vector<char>& buf, std::string& err) {
    gzFile g = gzopen(path, "rb");
    if (!g) { err = std::string("cannot open ") + path; return false; }
    gzbuffer(g, 1 << 20);
    size_t n = 0;
    buf.resize(1 << 24);
    for (;;) {
        if (buf.size() - n < (1 << 22)) buf.resize(buf.size() * 2);
        int k = gzread(g, buf.data() + n, (unsigned)std::min<size_t>(buf.size() - n, 1u << 30));
        if (k < 0) { err = "gzread failed"; gzclose(g); return false; }
        if (k == 0) break;
        n += (size_t)k;
    }
    gzclose(g);
    buf.resize(n);
    return true;
}

}  // namespace

struct cpn_calls {
    std::vector<ChrCols> chrs;
    std::unordered_map<std::string, int> chr_ix;
    std::vector<std::string> read_names;
    int64_t n_lines = 0;
    std::string err;
};

static thread_local std::string g_open_err;

extern "C" {

const char* cpn_calls_open_error(void) { return g_open_err.c_str(); }

int cpn_calls_open(const char* path, int32_t n_threads, cpn_calls** out) {
    if (!path || !out) { g_open_err = "null argument"; return CPN_ERR_ARG; }
    *out = nullptr;
    const char* data = nullptr;
    size_t size = 0;
    std::vector<char> zbuf;
    void* map = nullptr;
    size_t plen = strlen(path);
    if (plen > 3 && strcmp(path + plen - 3, ".gz") == 0) {
        if (!read_gz(path, zbuf, g_open_err)) return CPN_ERR_ARG;
        data = zbuf.data(); size = zbuf.size();
    } else {
        int fd = open(path, O_RDONLY);
        if (fd < 0) { g_open_err = std::string("cannot open ") + path; return CPN_ERR_ARG; }
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); g_open_err = "fstat failed"; return CPN_ERR_ARG; }
        size = (size_t)st.st_size;
        if (size > 0) {
            map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (map == MAP_FAILED) { close(fd); g_open_err = "mmap failed"; return CPN_ERR_NOMEM; }
            madvise(map, size, MADV_SEQUENTIAL);
            data = (const char*)map;
        }
        close(fd);
    }
    // header = first line
    size_t body = 0;
    if (size > 0) {
        const char* nl = (const char*)memchr(data, '\n', size);
        body = nl ? (size_t)(nl - data) + 1 : size;
    }
    int T = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, (size - body) / (1 << 16) + 1));
    std::vector<size_t> cut(T + 1, size);
    cut[0] = body;
    for (int t = 1; t < T; ++t) {
        size_t pos = body + (size - body) / T * t;
        pos = std::max(pos, cut[t - 1]);
        const char* nl = pos < size ? (const char*)memchr(data + pos, '\n', size - pos) : nullptr;
        cut[t] = nl ? (size_t)(nl - data) + 1 : size;
    }
    std::vector<Slice> sl(T);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] { parse_slice(data, cut[t], cut[t + 1], sl[t]); });
        for (auto& x : th) x.join();
    }
    auto fail = [&](const std::string& m, int code) {
        if (map) munmap(map, size);
        g_open_err = m;
        return code;
    };
    for (int t = 0; t < T; ++t)
        if (sl[t].bad_line >= 0) {
            int64_t ln = 1 + std::count(data, data + sl[t].bad_line, '\n');
            return fail("malformed line " + std::to_string(ln) + " in " + path, CPN_ERR_ARG);
        }
    cpn_calls* c = new cpn_calls();
    // global ids in order of first appearance (slices are in file order)
    std::unordered_map<std::string_view, int32_t> rix;
    std::vector<std::vector<int32_t>> cmap(T), rmap(T);
    std::vector<std::string_view> rnames;
    for (int t = 0; t < T; ++t) {
        cmap[t].resize(sl[t].chr_names.size());
        for (size_t i = 0; i < sl[t].chr_names.size(); ++i) {
            std::string nm(sl[t].chr_names[i]);
            auto it = c->chr_ix.find(nm);
            if (it == c->chr_ix.end()) { it = c->chr_ix.emplace(nm, (int)c->chrs.size()).first; c->chrs.emplace_back(); c->chrs.back().name = nm; }
            cmap[t][i] = it->second;
        }
        rmap[t].resize(sl[t].read_names.size());
        for (size_t i = 0; i < sl[t].read_names.size(); ++i) {
            auto it = rix.find(sl[t].read_names[i]);
            if (it == rix.end()) { it = rix.emplace(sl[t].read_names[i], (int32_t)rnames.size()).first; rnames.push_back(sl[t].read_names[i]); }
            rmap[t][i] = it->second;
        }
    }
    c->read_names.reserve(rnames.size());
    for (auto v : rnames) c->read_names.emplace_back(v);
    // columns per chromosome: every slice knows where its records go (slices are in file order), so the copy runs in parallel
    const size_t NC = c->chrs.size();
    std::vector<std::vector<size_t>> base(T, std::vector<size_t>(NC, 0));
    std::vector<size_t> cnt(NC, 0);
    for (int t = 0; t < T; ++t) {
        std::vector<size_t> mine(NC, 0);
        for (const Rec& r : sl[t].recs) ++mine[cmap[t][r.chr]];
        for (size_t k = 0; k < NC; ++k) { base[t][k] = cnt[k]; cnt[k] += mine[k]; }
        c->n_lines += (int64_t)sl[t].recs.size();
    }
    for (size_t k = 0; k < NC; ++k) {
        ChrCols& cc = c->chrs[k];
        cc.start.resize(cnt[k]); cc.end.resize(cnt[k]); cc.read.resize(cnt[k]); cc.lm.resize(cnt[k]); cc.lu.resize(cnt[k]);
    }
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                std::vector<size_t> at = base[t];
                for (const Rec& r : sl[t].recs) {
                    const int k = cmap[t][r.chr];
                    ChrCols& cc = c->chrs[k];
                    const size_t i = at[k]++;
                    cc.start[i] = r.start; cc.end[i] = r.end; cc.read[i] = rmap[t][r.read]; cc.lm[i] = r.lm; cc.lu[i] = r.lu;
                }
                std::vector<Rec>().swap(sl[t].recs);
            });
        for (auto& x : th) x.join();
    }
    if (map) munmap(map, size);
    // the reference assumes a file sorted by start within a chromosome; sort stably if it is not
    for (ChrCols& cc : c->chrs) {
        const size_t n = cc.start.size();
        if (!std::is_sorted(cc.start.begin(), cc.start.end())) {
            std::vector<size_t> ord(n);                 // size_t: a chromosome may hold more than 2^32 lines
            for (size_t i = 0; i < n; ++i) ord[i] = i;
            std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return cc.start[a] < cc.start[b]; });
            auto perm = [&](auto& v) { auto w = v; for (size_t i = 0; i < n; ++i) v[i] = w[ord[i]]; };
            perm(cc.start); perm(cc.end); perm(cc.read); perm(cc.lm); perm(cc.lu);
        }
        for (size_t i = 0; i < n; ++i) cc.max_span = std::max(cc.max_span, cc.end[i] - cc.start[i]);
    }
    *out = c;
    return CPN_OK;
}

void cpn_calls_close(cpn_calls* c) { delete c; }
const char* cpn_calls_error(const cpn_calls* c) { return c ? c->err.c_str() : "null handle"; }
int64_t cpn_calls_num_lines(const cpn_calls* c) { return c ? c->n_lines : 0; }
int64_t cpn_calls_num_reads(const cpn_calls* c) { return c ? (int64_t)c->read_names.size() : 0; }
int64_t cpn_calls_num_chr(const cpn_calls* c) { return c ? (int64_t)c->chrs.size() : 0; }
const char* cpn_calls_chr_name(const cpn_calls* c, int64_t i) { return (c && i >= 0 && i < (int64_t)c->chrs.size()) ? c->chrs[i].name.c_str() : ""; }
int64_t cpn_calls_chr_lines(const cpn_calls* c, int64_t i) { return (c && i >= 0 && i < (int64_t)c->chrs.size()) ? (int64_t)c->chrs[i].start.size() : 0; }
const char* cpn_calls_read_name(const cpn_calls* c, int64_t id) {
    return (c && id >= 0 && id < (int64_t)c->read_names.size()) ? c->read_names[id].c_str() : "";
}


}  // extern "C"

namespace {

// lines of one region: [lo, hi) on the start column, still to be filtered by end >= chrst
inline void region_span(const ChrCols& cc, int64_t chrst, int64_t chrend, size_t& lo, size_t& hi) {
    lo = std::lower_bound(cc.start.begin(), cc.start.end(), chrst - cc.max_span) - cc.start.begin();
    hi = std::upper_bound(cc.start.begin(), cc.start.end(), chrend) - cc.start.begin();
    if (hi < lo) hi = lo;
}

// row of a read inside the current region = order of first appearance.  `seen` holds the ids in row order; lookups go
// through a small open-addressing table kept behind it (slots store row+1, 0 = empty), rebuilt when it fills up.
struct RowMap {
    std::vector<int32_t> seen;
    std::vector<int32_t> slot;
    uint32_t mask = 0;
    void clear() {
        seen.clear();
        if (slot.empty()) { slot.assign(256, 0); mask = 255; } else std::fill(slot.begin(), slot.end(), 0);
    }
    static inline uint32_t h(int32_t id) { return (uint32_t)id * 2654435761u; }
    void grow() {
        slot.assign(slot.size() * 4, 0);
        mask = (uint32_t)slot.size() - 1;
        for (size_t r = 0; r < seen.size(); ++r) {
            uint32_t p = (h(seen[r]) >> 8) & mask;
            while (slot[p]) p = (p + 1) & mask;
            slot[p] = (int32_t)r + 1;
        }
    }
    inline int row_of(int32_t id) {
        uint32_t p = (h(id) >> 8) & mask;
        while (slot[p]) {
            if (seen[slot[p] - 1] == id) return slot[p] - 1;
            p = (p + 1) & mask;
        }
        seen.push_back(id);
        slot[p] = (int32_t)seen.size();
        if (seen.size() * 2 > slot.size()) grow();
        return (int)seen.size() - 1;
    }
};

template <class F>
void for_regions(int64_t R, int n_threads, F&& body) {
    int T = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::max<int64_t>(1, std::min<int64_t>(T, R / 64 + 1));
    std::atomic<int64_t> next{0};
    auto work = [&] {
        RowMap seen;                        // read ids of the region in order of first appearance
        for (;;) {
            int64_t r0 = next.fetch_add(64);
            if (r0 >= R) return;
            for (int64_t r = r0; r < std::min(R, r0 + 64); ++r) body(r, seen);
        }
    };
    if (T == 1) { work(); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(work);
    for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

int cpn_calls_count_reads(const cpn_calls* c, const char* chr, int64_t R, const int64_t* chrst, const int64_t* chrend, int32_t n_threads,
                          int64_t* n_reads) {
    if (!c || !chr || R < 0 || (R > 0 && (!chrst || !chrend || !n_reads))) return CPN_ERR_ARG;
    auto it = c->chr_ix.find(chr);
    if (it == c->chr_ix.end()) { std::fill(n_reads, n_reads + R, (int64_t)0); return CPN_OK; }
    const ChrCols& cc = c->chrs[it->second];
    for_regions(R, n_threads, [&](int64_t r, RowMap& seen) {
        seen.clear();
        size_t lo, hi;
        region_span(cc, chrst[r], chrend[r], lo, hi);
        for (size_t i = lo; i < hi; ++i)
            if (cc.end[i] >= chrst[r]) seen.row_of(cc.read[i]);
        n_reads[r] = (int64_t)seen.seen.size();
    });
    return CPN_OK;
}

int cpn_calls_fill(const cpn_calls* c, const char* chr, int64_t R, const int64_t* chrst, const int64_t* chrend, const int64_t* grp_off,
                   const int64_t* grp_st, const int64_t* grp_end, int32_t mod_nse, const int64_t* read_off, int32_t n_threads,
                   double* log_pyx_u, double* log_pyx_m, int64_t* read_ids) {
    if (!c || !chr || R < 0) return CPN_ERR_ARG;
    if (R == 0) return CPN_OK;
    if (!chrst || !chrend || !grp_off || !grp_st || !grp_end || !read_off || !log_pyx_u || !log_pyx_m) return CPN_ERR_ARG;
    std::vector<int64_t> obs(R + 1);
    obs[0] = 0;
    for (int64_t r = 0; r < R; ++r) obs[r + 1] = obs[r] + (read_off[r + 1] - read_off[r]) * (grp_off[r + 1] - grp_off[r]);
    auto it = c->chr_ix.find(chr);
    const ChrCols* cc = it == c->chr_ix.end() ? nullptr : &c->chrs[it->second];
    std::atomic<int> bad{0};
    const double qnan = std::numeric_limits<double>::quiet_NaN();
    for_regions(R, n_threads, [&](int64_t r, RowMap& seen) {
        const int64_t L = grp_off[r + 1] - grp_off[r], m = read_off[r + 1] - read_off[r];
        double* lu = log_pyx_u + obs[r];
        double* lm = log_pyx_m + obs[r];
        std::fill(lu, lu + m * L, qnan);
        std::fill(lm, lm + m * L, qnan);
        if (!cc || m == 0) return;
        const int64_t* gs = grp_st + grp_off[r];
        const int64_t* ge = grp_end + grp_off[r];
        seen.clear();
        size_t lo, hi;
        region_span(*cc, chrst[r], chrend[r], lo, hi);
        for (size_t i = lo; i < hi; ++i) {
            if (cc->end[i] < chrst[r]) continue;
            const int row = seen.row_of(cc->read[i]);
            if (row >= m) { bad = 1; return; }                        // read_off does not come from cpn_calls_count_reads
            const int64_t st = cc->start[i] + 1 - 5, en = cc->end[i] + 1 + 5;      // 0-based, no ±5 flanks in nanopolish (:181-184)
            const int64_t* g = std::lower_bound(gs, gs + L, st);
            if (g == gs + L || *g != st || ge[g - gs] != en) continue;             // the same group must match both ends (:187-197)
            const int64_t l = g - gs;
            double vm = cc->lm[i], vu = cc->lu[i];
            if (!mod_nse) {
                const bool meth = vm > vu;
                vm = meth ? 0.0 : -50.0;
                vu = meth ? -50.0 : 0.0;
            }
            lm[row * L + l] = vm;                                      // file order: the last line for a (read, group) wins
            lu[row * L + l] = vu;
        }
        if ((int64_t)seen.seen.size() != m) { bad = 1; return; }
        if (read_ids)
            for (int64_t k = 0; k < m; ++k) read_ids[read_off[r] + k] = seen.seen[k];
    });
    if (bad) { const_cast<cpn_calls*>(c)->err = "read_off does not match cpn_calls_count_reads for these regions"; return CPN_ERR_ARG; }
    return CPN_OK;
}

// Numerators of get_ave_depth and perc_gprs_obs (src/structures.jl:200-203) for every region of an emission batch:
// n_obs[r] = observed (read, group) cells, n_grp_obs[r] = groups observed in at least one read.  Works on any CSR emission
// arrays in the layout of cpn_fit_batch (NaN in log_pyx_u = not observed).
int cpn_calls_coverage(int64_t R, const int64_t* grp_off, const int64_t* read_off, const double* log_pyx_u, int32_t n_threads,
                       int64_t* n_obs, int64_t* n_grp_obs) {
    if (R < 0 || (R > 0 && (!grp_off || !read_off || !n_obs || !n_grp_obs))) return CPN_ERR_ARG;
    if (R == 0) return CPN_OK;
    std::vector<int64_t> obs(R + 1);
    obs[0] = 0;
    for (int64_t r = 0; r < R; ++r) obs[r + 1] = obs[r] + (read_off[r + 1] - read_off[r]) * (grp_off[r + 1] - grp_off[r]);
    if (obs[R] > 0 && !log_pyx_u) return CPN_ERR_ARG;
    for_regions(R, n_threads, [&](int64_t r, RowMap&) {
        const int64_t L = grp_off[r + 1] - grp_off[r], m = read_off[r + 1] - read_off[r];
        const double* lu = log_pyx_u + obs[r];
        int64_t cells = 0, groups = 0;
        for (int64_t l = 0; l < L; ++l) {
            int64_t c = 0;
            for (int64_t k = 0; k < m; ++k) c += (lu[k * L + l] == lu[k * L + l]);
            cells += c;
            groups += (c > 0);
        }
        n_obs[r] = cells;
        n_grp_obs[r] = groups;
    });
    return CPN_OK;
}


// The data filters of pmap_analyze_reg (src/nanopore.jl:309, 322-326) for every region of an emission batch, as status words:
// 0 = the region goes on to the EM; otherwise CPN_FIT_FILTERED − (OR of the CPN_FILTER_* reasons).  n_cpg[r] = number of CpG
// sites of region r (length(rs.cpg_pos)); get_ave_depth = observed cells / L, perc_gprs_obs = observed groups / L
// (src/structures.jl:200-203), both compared in double precision as the reference does.
int cpn_region_filter(int64_t R, const int64_t* n_cpg, const int64_t* grp_off, const int64_t* read_off, const double* log_pyx_u, double min_cov,
                      int32_t n_threads, int32_t* status) {
    if (R < 0 || (R > 0 && (!n_cpg || !grp_off || !read_off || !status))) return CPN_ERR_ARG;
    if (R == 0) return CPN_OK;
    std::vector<int64_t> n_obs(R), n_grp(R);
    int rc = cpn_calls_coverage(R, grp_off, read_off, log_pyx_u, n_threads, n_obs.data(), n_grp.data());
    if (rc != CPN_OK) return rc;
    for (int64_t r = 0; r < R; ++r) {
        const int64_t L = grp_off[r + 1] - grp_off[r], m = read_off[r + 1] - read_off[r];
        int why = 0;
        if (!(n_cpg[r] > 9)) why |= CPN_FILTER_FEW_CPGS;
        if (m <= 0) why |= CPN_FILTER_NO_READS;
        if (L > 0) {
            if (!((double)n_obs[r] / (double)L >= min_cov)) why |= CPN_FILTER_LOW_DEPTH;
            if (!((double)n_grp[r] / (double)L >= 0.66)) why |= CPN_FILTER_FEW_GROUPS_OBS;
        }
        if (!(L > 1)) why |= CPN_FILTER_ONE_GROUP;
        status[r] = why ? CPN_FIT_FILTERED - why : 0;
    }
    return CPN_OK;
}

// Packed form of an emission batch (see the header): llr = log_pyx_m − log_pyx_u per cell (NaN where unobserved),
// read_base = Σ_obs log_pyx_u per read, accumulated left to right exactly as the device's packing kernel does (compiled with
// -ffp-contract=off: plain IEEE subtraction and addition), so both routes into the EM give the same bits.  Regions are spread
// over n_threads host threads in contiguous blocks of equal cell count.
int cpn_pack_emissions(int64_t R, const int64_t* grp_off, const int64_t* read_off, const double* log_pyx_u, const double* log_pyx_m,
                       double* llr, double* read_base, int32_t n_threads) {
    if (R < 0 || (R > 0 && (!grp_off || !read_off || !log_pyx_u || !log_pyx_m || !llr || !read_base))) return CPN_ERR_ARG;
    if (R == 0) return CPN_OK;
    std::vector<int64_t> obs(R + 1);
    obs[0] = 0;
    for (int64_t r = 0; r < R; ++r) {
        const int64_t L = grp_off[r + 1] - grp_off[r], M = read_off[r + 1] - read_off[r];
        if (L < 0 || M < 0) return CPN_ERR_ARG;
        obs[r + 1] = obs[r] + L * M;
    }
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency(), R));
    std::atomic<int> bad{0};
    auto work = [&](int t) {
        const int64_t lo = std::lower_bound(obs.begin(), obs.end(), obs[R] / nt * t) - obs.begin();
        const int64_t hi = t + 1 == nt ? R : std::lower_bound(obs.begin(), obs.end(), obs[R] / nt * (t + 1)) - obs.begin();
        for (int64_t r = std::min(lo, R); r < std::min(hi, R); ++r) {
            const int64_t L = grp_off[r + 1] - grp_off[r], M = read_off[r + 1] - read_off[r];
            const double *u = log_pyx_u + obs[r], *m = log_pyx_m + obs[r];
            double* d = llr + obs[r];
            for (int64_t i = 0; i < M; ++i) {
                double base = 0.0;
                for (int64_t l = 0; l < L; ++l) {
                    const double uu = u[i * L + l];
                    if (uu == uu) {
                        const double mm = m[i * L + l];
                        if (mm != mm) bad.store(1);          // NaN marks "unobserved" in the packed table: an observed cell needs a number
                        d[i * L + l] = mm - uu; base += uu;
                    }
                    else d[i * L + l] = std::numeric_limits<double>::quiet_NaN();
                }
                read_base[read_off[r] + i] = base;
            }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    return bad.load() ? CPN_ERR_ARG : CPN_OK;
}

// WGBS front-end (wgbs.jl:222-323 turns the XM tags of a BAM into per-CpG calls x ∈ {−1 unmethylated, +1 methylated, 0 not
// covered}; that parser is out of scope, its output is the input here).  A call is a hard emission (wgbs.jl:294-304 with
// CpelNano.jl:32-33: log p(y|x) = 0 if the call agrees with x, −50 otherwise), so in packed form: llr = ∓50 and every methylated
// call adds −50 to the read's Σ_obs log p(y|−1).  calls: [Σ_r M_r·N_r] int8, read-major per region.
int cpn_wgbs_pack(int64_t R, const int64_t* cpg_off, const int64_t* read_off, const int8_t* calls, double* llr, double* read_base,
                  int32_t n_threads) {
    if (R < 0 || (R > 0 && (!cpg_off || !read_off || !calls || !llr || !read_base))) return CPN_ERR_ARG;
    if (R == 0) return CPN_OK;
    std::vector<int64_t> obs(R + 1);
    obs[0] = 0;
    for (int64_t r = 0; r < R; ++r) {
        const int64_t N = cpg_off[r + 1] - cpg_off[r], M = read_off[r + 1] - read_off[r];
        if (N < 0 || M < 0) return CPN_ERR_ARG;
        obs[r + 1] = obs[r] + N * M;
    }
    const double right = 0.0, wrong = -50.0;                 // log_pyx_right_x, log_pyx_wrong_x
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency(), R));
    std::atomic<int> bad{0};
    auto work = [&](int t) {
        const int64_t lo = std::lower_bound(obs.begin(), obs.end(), obs[R] / nt * t) - obs.begin();
        const int64_t hi = t + 1 == nt ? R : std::lower_bound(obs.begin(), obs.end(), obs[R] / nt * (t + 1)) - obs.begin();
        for (int64_t r = std::min(lo, R); r < std::min(hi, R); ++r) {
            const int64_t N = cpg_off[r + 1] - cpg_off[r], M = read_off[r + 1] - read_off[r];
            const int8_t* x = calls + obs[r];
            double* d = llr + obs[r];
            for (int64_t i = 0; i < M; ++i) {
                double base = 0.0;
                for (int64_t n = 0; n < N; ++n) {
                    const int8_t c = x[i * N + n];
                    if (c == 1) { d[i * N + n] = right - wrong; base += wrong; }        // log_pyx_m = right, log_pyx_u = wrong
                    else if (c == -1) { d[i * N + n] = wrong - right; base += right; }  // log_pyx_m = wrong, log_pyx_u = right
                    else { d[i * N + n] = std::numeric_limits<double>::quiet_NaN(); if (c != 0) bad.store(1); }
                }
                read_base[read_off[r] + i] = base;
            }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    return bad.load() ? CPN_ERR_ARG : CPN_OK;
}

}  // extern "C"
