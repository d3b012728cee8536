// cpn_genome.cpp — genome-side inputs of the path, native and batched (host only).
//
// Replaces, for all estimation regions of a chromosome at once,
//     get_grp_info!      src/genome_analysis.jl:226-277   CpG sites of a region (regex on the substring), groups, densities, distances
//     get_grps_from_cgs  src/genome_analysis.jl:63-111    consecutive CpGs at most min_grp_dist apart share a group; interval ± 5
//     get_ρn             src/genome_analysis.jl:122-140   one regex scan of a ≤ 1 kb window per CpG
//     get_ρl / get_dn / get_dl  :150-216
//     get_chr_part       src/genome_analysis.jl:327-458   3 kb windows whose right boundary is moved off a CpG group
// by one scan of the chromosome's bytes (cpn_cpg_scan: positions of every CG, case-insensitive — regex matches of r"[Cc][Gg]"
// cannot overlap, so every C followed by G is a match) and binary searches into that index: a regex scan of the substring
// seq[a..b] finds exactly the CGs whose C lies in [a, b-1].  Integer outputs are bit-exact by construction; the doubles
// (ρ = count/1000, ρ_l = left-to-right sum / N_l, distances) repeat the reference's operations in its order.
// Outputs are the CSR arrays cpn_fit_batch / cpn_stat_sums_batch take.  Coordinates are 1-based, closed.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "../../include/cpelnano_b200.h"

namespace {

inline int64_t lower(const int64_t* cg, int64_t n, int64_t v) { return std::lower_bound(cg, cg + n, v) - cg; }
inline int64_t upper(const int64_t* cg, int64_t n, int64_t v) { return std::upper_bound(cg, cg + n, v) - cg; }

// CG matches lying completely inside seq[st..end]
inline void cpgs_in(const int64_t* cg, int64_t n, int64_t st, int64_t end, int64_t& lo, int64_t& hi) {
    lo = lower(cg, n, st);
    hi = upper(cg, n, end - 1);
    if (hi < lo) hi = lo;
}

template <class F>
void parallel_for(int64_t R, int n_threads, F&& body) {
    int T = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::max<int64_t>(1, std::min<int64_t>(T, R / 256 + 1));
    std::atomic<int64_t> next{0};
    auto work = [&] {
        for (;;) {
            int64_t r0 = next.fetch_add(256);
            if (r0 >= R) return;
            for (int64_t r = r0; r < std::min(R, r0 + 256); ++r) body(r);
        }
    };
    if (T == 1) { work(); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(work);
    for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

int64_t cpn_cpg_scan(const char* seq, int64_t n, int32_t n_threads, int64_t* cpg_pos, int64_t cap) {
    if (!seq || n < 0) return -1;
    if (n < 2) return 0;
    int T = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::max<int64_t>(1, std::min<int64_t>(T, n / (1 << 20) + 1));
    std::vector<int64_t> cut(T + 1), cnt(T, 0);
    for (int t = 0; t <= T; ++t) cut[t] = (n - 1) / T * t;
    cut[T] = n - 1;                                              // position i (0-based) is a match when seq[i], seq[i+1] = C, G
    const unsigned char* u = reinterpret_cast<const unsigned char*>(seq);
    // branch-free test (bitwise &, not &&) so the counting loop vectorises; 0xDF folds ASCII case
    auto is_cg = [u](int64_t i) -> int { return (int)((u[i] & 0xDF) == 'C') & (int)((u[i + 1] & 0xDF) == 'G'); };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                int64_t c = 0;
                for (int64_t i0 = cut[t]; i0 < cut[t + 1]; i0 += 4096) {        // 32-bit partial sums keep the inner loop narrow
                    const int64_t i1 = std::min(cut[t + 1], i0 + 4096);
                    int part = 0;
                    for (int64_t i = i0; i < i1; ++i) part += is_cg(i);
                    c += part;
                }
                cnt[t] = c;
            });
        for (auto& x : th) x.join();
    }
    int64_t total = 0;
    std::vector<int64_t> base(T);
    for (int t = 0; t < T; ++t) { base[t] = total; total += cnt[t]; }
    if (!cpg_pos || cap < total) return total;                   // sizing call
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                int64_t o = base[t];
                for (int64_t i0 = cut[t]; i0 < cut[t + 1]; i0 += 64) {          // skip 64-byte blocks without a match
                    const int64_t i1 = std::min(cut[t + 1], i0 + 64);
                    int any = 0;
                    for (int64_t i = i0; i < i1; ++i) any += is_cg(i);
                    if (!any) continue;
                    for (int64_t i = i0; i < i1; ++i)
                        if (is_cg(i)) cpg_pos[o++] = i + 1;      // 1-based position of the C
                }
            });
        for (auto& x : th) x.join();
    }
    return total;
}

int64_t cpn_chr_part(const int64_t* cg, int64_t ncg, int64_t chr_size, int64_t reg_st, int64_t reg_end, int64_t size_est_reg,
                     int64_t min_grp_dist, int64_t* out_st, int64_t* out_end, int64_t cap) {
    if ((ncg > 0 && !cg) || size_est_reg < 1) return -1;
    int64_t k = 0;
    auto push = [&](int64_t a, int64_t b) { if (out_st && out_end && k < cap) { out_st[k] = a; out_end[k] = b; } ++k; };
    int64_t i = reg_st;
    while (i < reg_end) {
        int64_t bndy = i + size_est_reg - 1;
        if (bndy >= reg_end) { push(i, reg_end); break; }
        int64_t lo, hi;
        cpgs_in(cg, ncg, bndy - 1000, std::min(chr_size, bndy + 1000), lo, hi);
        // groups of the window, left to right; the first whose padded interval contains the boundary moves it (:361-373)
        // get_grps_from_cgs returns a lone CpG's interval WITHOUT the ±5 padding (early return at :67, before the padding loop)
        const int64_t pad = (hi - lo == 1) ? 0 : 5;
        int64_t a = lo;
        while (a < hi) {
            int64_t b = a;
            while (b + 1 < hi && cg[b + 1] - cg[b] <= min_grp_dist) ++b;
            const int64_t gst = cg[a] - pad, gend = cg[b] + pad;
            if (gst <= bndy && bndy <= gend) { bndy = (2 * bndy - gst > gend) ? gend + 2 : gst - 1; break; }
            a = b + 1;
        }
        push(i, bndy);
        i = bndy + 1;
    }
    return k;
}

int cpn_grp_info_count(const int64_t* cg, int64_t ncg, int64_t R, const int64_t* chrst, const int64_t* chrend, int64_t min_grp_dist,
                       int32_t n_threads, int64_t* N, int64_t* L) {
    if (R < 0 || (R > 0 && (!chrst || !chrend || !N || !L)) || (ncg > 0 && !cg)) return CPN_ERR_ARG;
    parallel_for(R, n_threads, [&](int64_t r) {
        int64_t lo, hi;
        cpgs_in(cg, ncg, chrst[r], chrend[r], lo, hi);
        N[r] = hi - lo;
        int64_t groups = hi > lo ? 1 : 0;
        for (int64_t i = lo + 1; i < hi; ++i) groups += (cg[i] - cg[i - 1] > min_grp_dist);
        L[r] = (hi - lo > 1) ? groups : 0;                       // get_grp_info! returns before grouping when N <= 1 (:239)
    });
    return CPN_OK;
}

int cpn_grp_info_fill(const int64_t* cg, int64_t ncg, int64_t chr_size, int64_t R, const int64_t* chrst, const int64_t* chrend,
                      int64_t min_grp_dist, const int64_t* cpg_off, const int64_t* grp_off, int32_t n_threads, int64_t* cpg_pos,
                      double* rho_n, double* d_n, double* Nl, double* rho_l, double* d_l, int64_t* grp_st, int64_t* grp_end) {
    if (R < 0 || (ncg > 0 && !cg)) return CPN_ERR_ARG;
    if (R == 0) return CPN_OK;
    if (!chrst || !chrend || !cpg_off || !grp_off || !cpg_pos || !rho_n || !d_n || !Nl || !rho_l || !d_l || !grp_st || !grp_end) return CPN_ERR_ARG;
    // d_n / d_l hold N-1 / L-1 entries per region with N > 1 and none otherwise: their offsets are running sums
    std::vector<int64_t> dn_off(R + 1, 0), dl_off(R + 1, 0);
    for (int64_t r = 0; r < R; ++r) {
        dn_off[r + 1] = dn_off[r] + std::max<int64_t>(0, cpg_off[r + 1] - cpg_off[r] - 1);
        dl_off[r + 1] = dl_off[r] + std::max<int64_t>(0, grp_off[r + 1] - grp_off[r] - 1);
    }
    std::atomic<int> bad{0};
    parallel_for(R, n_threads, [&](int64_t r) {
        int64_t lo, hi;
        cpgs_in(cg, ncg, chrst[r], chrend[r], lo, hi);
        const int64_t N = hi - lo;
        if (N != cpg_off[r + 1] - cpg_off[r]) { bad = 1; return; }
        int64_t* pos = cpg_pos + cpg_off[r];
        for (int64_t i = 0; i < N; ++i) pos[i] = cg[lo + i];
        if (N <= 1) { if (grp_off[r + 1] != grp_off[r]) bad = 1; return; }
        const int64_t cont_st = std::max<int64_t>(1, chrst[r] - 500), cont_end = std::min(chr_size, chrend[r] + 500);
        const int64_t clen = cont_end - cont_st + 1;
        double* rn = rho_n + cpg_off[r];
        for (int64_t i = 0; i < N; ++i) {
            const int64_t o = pos[i] - cont_st;                   // the offset get_ρn is handed (cpg_pos .- dna_cont_st)
            const int64_t win_st = std::max<int64_t>(1, o - 500) + cont_st - 1, win_end = std::min(clen, o + 500) + cont_st - 1;
            int64_t a, b;
            cpgs_in(cg, ncg, win_st, win_end, a, b);
            rn[i] = (double)(b - a) / 1000.0;
        }
        double* dn = d_n + dn_off[r];
        for (int64_t i = 0; i + 1 < N; ++i) dn[i] = (double)(pos[i + 1] - pos[i]);
        const int64_t L = grp_off[r + 1] - grp_off[r];
        double* nl = Nl + grp_off[r];
        double* rl = rho_l + grp_off[r];
        double* dl = d_l + dl_off[r];
        int64_t* gs = grp_st + grp_off[r];
        int64_t* ge = grp_end + grp_off[r];
        int64_t a = 0, g = 0, prev_last = -1;
        while (a < N) {
            int64_t b = a;
            while (b + 1 < N && pos[b + 1] - pos[b] <= min_grp_dist) ++b;
            if (g >= L) { bad = 1; return; }
            double s = rn[a];
            for (int64_t i = a + 1; i <= b; ++i) s += rn[i];      // sum(ρ_n[group]) then / N_l (:159-160); Julia's sum is strictly
                                                                  // left to right below 16 elements (Base._mapreduce), SIMD-ordered above
            nl[g] = (double)(b - a + 1);
            rl[g] = s / nl[g];
            gs[g] = pos[a] - 5; ge[g] = pos[b] + 5;
            if (g > 0) dl[g - 1] = (double)(pos[a] - pos[prev_last]);
            prev_last = b;
            ++g;
            a = b + 1;
        }
        if (g != L) bad = 1;
    });
    return bad ? CPN_ERR_ARG : CPN_OK;
}

// get_nls_reg_info! (src/genome_analysis.jl:472-538) for R regions: the per-region entry point cpn_nls_reg_info in a loop,
// threaded.  Pass 1 (sub_st == NULL): K[r].  Pass 2: sub_off = exclusive cumulative sum of K.
int cpn_nls_reg_info_batch(int64_t R, const int64_t* chrst, const int64_t* chrend, int64_t max_size_subreg, const int64_t* cpg_off,
                           const int64_t* cpg_pos, int32_t n_threads, int64_t* K, const int64_t* sub_off, int64_t* sub_st, int64_t* sub_end,
                           int64_t* sub_p, int64_t* sub_q) {
    if (R < 0 || (R > 0 && (!chrst || !chrend || !cpg_off))) return CPN_ERR_ARG;
    std::atomic<int> bad{0};
    if (!sub_st) {
        if (R > 0 && !K) return CPN_ERR_ARG;
        parallel_for(R, n_threads, [&](int64_t r) {
            K[r] = cpn_nls_reg_info(chrst[r], chrend[r], max_size_subreg, cpg_pos ? cpg_pos + cpg_off[r] : nullptr, cpg_off[r + 1] - cpg_off[r], nullptr,
                                    nullptr, nullptr, nullptr);
            if (K[r] < 0) bad = 1;
        });
        return bad ? CPN_ERR_ARG : CPN_OK;
    }
    if (!sub_off || !sub_end || !sub_p || !sub_q || (cpg_off[R] > 0 && !cpg_pos)) return CPN_ERR_ARG;
    parallel_for(R, n_threads, [&](int64_t r) {
        const int64_t k0 = sub_off[r];
        int64_t k = cpn_nls_reg_info(chrst[r], chrend[r], max_size_subreg, cpg_pos + cpg_off[r], cpg_off[r + 1] - cpg_off[r], sub_st + k0, sub_end + k0,
                                     sub_p + k0, sub_q + k0);
        if (k != sub_off[r + 1] - k0) bad = 1;
    });
    return bad ? CPN_ERR_ARG : CPN_OK;
}

// ---- FASTA (get_genome_info / get_chr_fa_rec, src/genome_analysis.jl:19-61,283-312) ----------------------------------------------
// Records are held in memory with line breaks removed; the name is the identifier (text after '>' up to the first blank).
struct cpn_fasta {
    std::vector<std::string> names;
    std::vector<std::string> seqs;
};
static thread_local std::string g_fasta_err;

const char* cpn_fasta_open_error(void) { return g_fasta_err.c_str(); }

int cpn_fasta_open(const char* path, cpn_fasta** out) {
    if (!path || !out) { g_fasta_err = "null argument"; return CPN_ERR_ARG; }
    *out = nullptr;
    std::vector<char> zbuf;
    const char* data = nullptr;
    size_t size = 0;
    void* map = nullptr;
    const size_t plen = strlen(path);
    if (plen > 3 && strcmp(path + plen - 3, ".gz") == 0) {
        gzFile g = gzopen(path, "rb");
        if (!g) { g_fasta_err = std::string("cannot open ") + path; return CPN_ERR_ARG; }
        gzbuffer(g, 1 << 20);
        zbuf.resize(1 << 24);
        for (;;) {
            if (zbuf.size() - size < (1 << 22)) zbuf.resize(zbuf.size() * 2);
            int k = gzread(g, zbuf.data() + size, (unsigned)std::min<size_t>(zbuf.size() - size, 1u << 30));
            if (k < 0) { gzclose(g); g_fasta_err = "gzread failed"; return CPN_ERR_ARG; }
            if (k == 0) break;
            size += (size_t)k;
        }
        gzclose(g);
        data = zbuf.data();
    } else {
        int fd = open(path, O_RDONLY);
        if (fd < 0) { g_fasta_err = std::string("cannot open ") + path; return CPN_ERR_ARG; }
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); g_fasta_err = "fstat failed"; return CPN_ERR_ARG; }
        size = (size_t)st.st_size;
        if (size > 0) {
            map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (map == MAP_FAILED) { close(fd); g_fasta_err = "mmap failed"; return CPN_ERR_NOMEM; }
            data = (const char*)map;
        }
        close(fd);
    }
    cpn_fasta* f = new cpn_fasta();
    // pass 1: record names and exact lengths; pass 2: copy the lines (one exact allocation per record)
    std::vector<size_t> len;
    for (int pass = 0; pass < 2; ++pass) {
        const char* p = data;
        const char* end = data + size;
        int64_t rec = -1;
        while (p < end) {
            const char* nl = (const char*)memchr(p, '\n', end - p);
            const char* le = nl ? nl : end;
            const char* line = p;
            p = nl ? nl + 1 : end;
            if (le > line && le[-1] == '\r') --le;
            if (le == line) continue;
            if (*line == '>') {
                ++rec;
                if (pass == 0) {
                    const char* q = line + 1;
                    while (q < le && *q != ' ' && *q != '\t') ++q;
                    f->names.emplace_back(line + 1, q - (line + 1));
                    len.push_back(0);
                } else {
                    f->seqs[rec].reserve(len[rec]);
                }
            } else if (rec >= 0) {
                if (pass == 0) len[rec] += (size_t)(le - line); else f->seqs[rec].append(line, le - line);
            }
        }
        if (pass == 0) f->seqs.resize(f->names.size());
    }
    if (map) munmap(map, size);
    *out = f;
    return CPN_OK;
}
void cpn_fasta_close(cpn_fasta* f) { delete f; }
int64_t cpn_fasta_num_records(const cpn_fasta* f) { return f ? (int64_t)f->names.size() : 0; }
const char* cpn_fasta_name(const cpn_fasta* f, int64_t i) { return (f && i >= 0 && i < (int64_t)f->names.size()) ? f->names[i].c_str() : ""; }
int64_t cpn_fasta_length(const cpn_fasta* f, int64_t i) { return (f && i >= 0 && i < (int64_t)f->seqs.size()) ? (int64_t)f->seqs[i].size() : -1; }
const char* cpn_fasta_seq(const cpn_fasta* f, int64_t i) { return (f && i >= 0 && i < (int64_t)f->seqs.size()) ? f->seqs[i].data() : nullptr; }

}  // extern "C"
