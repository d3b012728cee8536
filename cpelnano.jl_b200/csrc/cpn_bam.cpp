// cpn_bam.cpp — native WGBS calls ingestion: BAM (Bismark / Arioc XM tags) → per-read, per-CpG hard calls.
//
// Replaces, for all estimation regions of a chromosome at a time,
//     get_calls_bam      src/wgbs.jl:222-323   (opens the BAM and walks the index again for every region)
//     clean_records      src/wgbs.jl:93-118,  order_bams :66-78,  get_align_strand :23-55,  try_olaps :129-139
// by ONE pass over the file (BGZF blocks inflated in parallel, records indexed by reference and position) and, per region, a
// binary search on the start positions.  The output is the int8 [m_r × N_r] table cpn_wgbs_pack / cpn_analyze_wgbs_batch take
// (−1 unmethylated, +1 methylated, 0 not covered).
//
// Semantics kept from the reference:
//  * a record overlaps a region when position ≤ min(roi_end, chr_size − 151) and rightposition ≥ roi_st (BioAlignments'
//    eachoverlap; rightposition = position + reference length of the CIGAR − 1), on the region's chromosome;
//  * kept records: mapped, XM tag present, XS tag absent, FLAG ∈ {0, 16, 83, 99, 147, 163}, MAPQ > 39 (src/CpelNano.jl:30-31);
//  * single end: every record is a template, FLAG 0 → OT, 16 → OB, anything else is an error (the reference exits);
//    paired end: templates are the read names with exactly two kept records among the region's, in order of first appearance;
//    the mate that starts first is R1; (99,147) OT, (83,163) OB, (147,99) CTOT, (163,83) CTOB, anything else is an error;
//  * the methylation string is XM without the trimmed ends — trim = (R1 5', R1 3', R2 5', R2 3') applied as in :262-263 and
//    :271-272 — and for pairs R1's string, dots up to R2's start, then the part of R2's string beyond R1's end;
//  * z / Z at string index k is a call at genomic position k − OFFSET with OFFSET = (1 for OT/CTOT, 2 for OB/CTOB) −
//    (position(R1) + trim[1]); a template is a row when at least one of the region's CpG sites carries a call.
// One documented difference: when R2 ends before R1's string does, the reference indexes R2's string past its end and throws;
// here R2 then contributes nothing.
//
// Host-only code: no CUDA here, usable without a GPU (tests/test_wgbs_bam.py runs on CPU).
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

constexpr int THRESH_MAPQ = 39;                 // src/CpelNano.jl:30
constexpr int64_t END_MARGIN = 151;             // roi_end = min(roi_end, chr_size - 151), src/wgbs.jl:240

struct RecIx {
    int32_t ref;
    int32_t flag, mapq;
    int64_t pos1, right1;                       // 1-based position and rightposition
    uint32_t qname_off, qname_len;              // into data (without the NUL)
    uint32_t xm_off, xm_len;                    // value of XM:Z (xm_len = 0 and has_xm = false when absent)
    bool has_xm, has_xs;
};

}  // namespace

struct cpn_bam {
    std::vector<uint8_t> data;                  // the inflated BAM stream
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
    std::vector<RecIx> recs;                    // file order
    std::vector<std::vector<int64_t>> by_ref;   // record indices per reference, by (position, file order)
    std::vector<std::vector<int64_t>> by_ref_pos;   // their positions (for the binary search)
    std::vector<int64_t> max_span;              // per reference: longest right1 - pos1 + 1
    mutable std::string err;
};

namespace {

thread_local std::string g_open_err;

struct Block { size_t in_off, in_len, out_off, out_len; };

inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline int32_t rdi32(const uint8_t* p) { return (int32_t)rd32(p); }

// BGZF: a series of gzip members, each with an extra subfield 'B','C' holding the member's total size − 1
bool scan_bgzf(const uint8_t* f, size_t n, std::vector<Block>& blocks, std::string& err) {
    size_t p = 0, out = 0;
    while (p < n) {
        if (n - p < 18 || f[p] != 0x1f || f[p + 1] != 0x8b || f[p + 2] != 8 || !(f[p + 3] & 4)) { err = "not a BGZF block at byte " + std::to_string(p); return false; }
        const size_t xlen = rd16(f + p + 10);
        size_t q = p + 12, xe = q + xlen;
        if (xe > n) { err = "truncated BGZF header"; return false; }
        size_t bsize = 0;
        while (q + 4 <= xe) {
            const size_t slen = rd16(f + q + 2);
            if (f[q] == 'B' && f[q + 1] == 'C' && slen == 2) bsize = (size_t)rd16(f + q + 4) + 1;
            q += 4 + slen;
        }
        if (bsize == 0 || p + bsize > n || bsize < 12 + xlen + 8) { err = "BGZF block without a valid BC field at byte " + std::to_string(p); return false; }
        const size_t isize = rd32(f + p + bsize - 4);
        blocks.push_back({xe, bsize - (12 + xlen) - 8, out, isize});
        out += isize;
        p += bsize;
    }
    return true;
}

bool inflate_block(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    if (out_len == 0) return true;
    z_stream s;
    std::memset(&s, 0, sizeof(s));
    if (inflateInit2(&s, -15) != Z_OK) return false;
    s.next_in = const_cast<Bytef*>(in); s.avail_in = (uInt)in_len;
    s.next_out = out; s.avail_out = (uInt)out_len;
    const int rc = inflate(&s, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && s.avail_out == 0;
    inflateEnd(&s);
    return ok;
}

// reference length of a record's alignment: CIGAR operations that consume the reference (M, D, N, =, X)
inline int64_t align_length(const uint8_t* cig, int n_ops) {
    int64_t len = 0;
    for (int k = 0; k < n_ops; ++k) {
        const uint32_t c = rd32(cig + 4 * k);
        const uint32_t op = c & 15u;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) len += c >> 4;
    }
    return len;
}

// walks the auxiliary fields; returns false on a malformed field
bool scan_aux(const uint8_t* p, const uint8_t* e, const uint8_t* base, RecIx& r) {
    while (p + 3 <= e) {
        const char t0 = (char)p[0], t1 = (char)p[1], ty = (char)p[2];
        p += 3;
        const uint8_t* v = p;
        size_t len = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': len = 1; break;
            case 's': case 'S': len = 2; break;
            case 'i': case 'I': case 'f': len = 4; break;
            case 'Z': case 'H': {
                const void* z = std::memchr(p, 0, (size_t)(e - p));
                if (!z) return false;
                len = (size_t)((const uint8_t*)z - p) + 1;
                break;
            }
            case 'B': {
                if (p + 5 > e) return false;
                const char st = (char)p[0];
                const size_t cnt = rd32(p + 1);
                const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                len = 5 + cnt * es;
                break;
            }
            default: return false;
        }
        if (p + len > e) return false;
        if (t0 == 'X' && t1 == 'M') {
            r.has_xm = true;
            if (ty == 'Z') { r.xm_off = (uint32_t)(v - base); r.xm_len = (uint32_t)(len - 1); }
        } else if (t0 == 'X' && t1 == 'S') {
            r.has_xs = true;
        }
        p += len;
    }
    return p == e;
}

enum Strand { OT = 0, OB = 1, CTOT = 2, CTOB = 3, BAD = -1 };

inline int se_strand(int flag) { return flag == 0 ? OT : (flag == 16 ? OB : BAD); }
inline int pe_strand(int f1, int f2) {
    if (f1 == 99 && f2 == 147) return OT;
    if (f1 == 83 && f2 == 163) return OB;
    if (f1 == 147 && f2 == 99) return CTOT;
    if (f1 == 163 && f2 == 83) return CTOB;
    return BAD;
}
inline bool flag_allowed(int f) { return f == 0 || f == 16 || f == 83 || f == 99 || f == 147 || f == 163; }

// SubString(s, 1 + a : length(s) - b) as (offset, length) into s; an empty range is the empty string (a, b ≥ 0)
inline void trimmed(uint32_t len, int64_t a, int64_t b, int64_t& off, int64_t& n) {
    const int64_t m = (int64_t)len - a - b;
    if (m <= 0) { off = 0; n = 0; }
    else { off = a; n = m; }
}

struct RegionJob {
    const cpn_bam* bam;
    int ref;
    bool pe;
    int64_t trim[4];
};

// The rows of one region.  rows == nullptr: count only.  Returns the number of rows, or -1 with `err` set.
int64_t region_calls(const RegionJob& jb, int64_t roi_st, int64_t roi_end, const int64_t* cpg, int64_t N, int8_t* rows, std::string& err,
                     std::vector<int64_t>& kept, std::string& meth) {
    const cpn_bam& b = *jb.bam;
    roi_end = std::min(roi_end, b.ref_lens[jb.ref] - END_MARGIN);
    if (roi_end < roi_st) return 0;                                   // try_olaps: an empty range raises and is caught (:131-136)
    const auto& ix = b.by_ref[jb.ref];
    const auto& pos = b.by_ref_pos[jb.ref];
    const int64_t lo_pos = roi_st - b.max_span[jb.ref];
    size_t k0 = (size_t)(std::lower_bound(pos.begin(), pos.end(), lo_pos) - pos.begin());
    size_t k1 = (size_t)(std::upper_bound(pos.begin(), pos.end(), roi_end) - pos.begin());
    kept.clear();
    for (size_t k = k0; k < k1; ++k) {
        const RecIx& r = b.recs[ix[k]];
        if (r.right1 < roi_st) continue;
        if ((r.flag & 0x4) || !r.has_xm || r.has_xs || !flag_allowed(r.flag) || r.mapq <= THRESH_MAPQ) continue;
        kept.push_back(ix[k]);
    }
    if (kept.empty()) return 0;
    struct Tmpl { int64_t r1, r2; int strand; };
    std::vector<Tmpl> tmpls;
    if (!jb.pe) {
        tmpls.reserve(kept.size());
        for (int64_t i : kept) {
            const int s = se_strand(b.recs[i].flag);
            if (s == BAD) { err = "single-end mode expects FLAG 0 or 16, found " + std::to_string(b.recs[i].flag); return -1; }
            tmpls.push_back({i, -1, s});
        }
    } else {
        std::unordered_map<std::string_view, int> slot;              // name → index into `groups`
        struct Grp { int64_t a, b; int n; };
        std::vector<Grp> groups;
        for (int64_t i : kept) {
            const RecIx& r = b.recs[i];
            std::string_view nm((const char*)b.data.data() + r.qname_off, r.qname_len);
            auto it = slot.find(nm);
            if (it == slot.end()) { slot.emplace(nm, (int)groups.size()); groups.push_back({i, -1, 1}); }
            else { Grp& g = groups[it->second]; if (g.n == 1) g.b = i; g.n++; }
        }
        for (const Grp& g : groups) {
            if (g.n != 2) continue;
            const RecIx &ra = b.recs[g.a], &rb = b.recs[g.b];
            Tmpl t;
            if (ra.pos1 <= rb.pos1) { t.r1 = g.a; t.r2 = g.b; t.strand = pe_strand(ra.flag, rb.flag); }
            else { t.r1 = g.b; t.r2 = g.a; t.strand = pe_strand(rb.flag, ra.flag); }
            if (t.strand == BAD) { err = "paired-end mode: unexpected FLAG combination " + std::to_string(b.recs[t.r1].flag) + "/" + std::to_string(b.recs[t.r2].flag); return -1; }
            tmpls.push_back(t);
        }
    }
    int64_t n_rows = 0;
    const char* base = (const char*)b.data.data();
    for (const Tmpl& t : tmpls) {
        const RecIx& r1 = b.recs[t.r1];
        int64_t o1, n1;
        trimmed(r1.xm_len, jb.trim[0], jb.trim[1], o1, n1);
        const char* s;
        int64_t slen;
        if (!jb.pe) {
            s = base + r1.xm_off + o1; slen = n1;
        } else {
            const RecIx& r2 = b.recs[t.r2];
            int64_t o2, n2;
            trimmed(r2.xm_len, jb.trim[3], jb.trim[2], o2, n2);
            const int64_t dist_st = std::llabs(r2.pos1 + jb.trim[3] - r1.pos1 - jb.trim[0]);
            meth.assign(base + r1.xm_off + o1, (size_t)n1);
            if (dist_st > n1) meth.append((size_t)(dist_st - n1), '.');
            const int64_t from = std::max<int64_t>(1, n1 + 1 - dist_st);          // 1-based start inside R2's string
            if (from <= n2) meth.append(base + r2.xm_off + o2 + (from - 1), (size_t)(n2 - from + 1));
            s = meth.data(); slen = (int64_t)meth.size();
        }
        const int64_t OFFSET = ((t.strand == OT || t.strand == CTOT) ? 1 : 2) - (r1.pos1 + jb.trim[0]);
        // CpG sites whose string index lies inside the string: cpg + OFFSET in [1, slen]
        const int64_t* c0 = std::lower_bound(cpg, cpg + N, 1 - OFFSET);
        const int64_t* c1 = std::upper_bound(cpg, cpg + N, slen - OFFSET);
        // a template becomes a row only if one of the region's CpG sites carries a call: nothing is written before that is known
        // (the slot after the region's last row belongs to the next region, which another thread may be filling)
        const int64_t* first = c0;
        while (first < c1 && s[*first + OFFSET - 1] != 'Z' && s[*first + OFFSET - 1] != 'z') ++first;
        if (first == c1) continue;
        if (rows) {
            int8_t* row = rows + n_rows * N;
            std::memset(row, 0, (size_t)N);
            for (const int64_t* c = first; c < c1; ++c) {
                const char ch = s[*c + OFFSET - 1];
                if (ch == 'Z') row[c - cpg] = 1;
                else if (ch == 'z') row[c - cpg] = -1;
            }
        }
        ++n_rows;
    }
    return n_rows;
}

int find_ref(const cpn_bam* b, const char* chr) {
    for (size_t i = 0; i < b->ref_names.size(); ++i)
        if (b->ref_names[i] == chr) return (int)i;
    return -1;
}

int run_regions(const cpn_bam* bam, const char* chr, int64_t R, const int64_t* roi_st, const int64_t* roi_end, const int64_t* cpg_off,
                const int64_t* cpg_pos, int32_t pe, const int64_t* trim, int32_t n_threads, int64_t* n_reads, const int64_t* cell_off,
                int8_t* calls) {
    if (!bam) return CPN_ERR_ARG;
    if (R < 0 || !chr || (R > 0 && (!roi_st || !roi_end || !cpg_off || !cpg_pos || !trim))) { bam->err = "null pointer"; return CPN_ERR_ARG; }
    if (trim[0] < 0 || trim[1] < 0 || trim[2] < 0 || trim[3] < 0) { bam->err = "negative trim"; return CPN_ERR_ARG; }
    const int ref = find_ref(bam, chr);
    if (ref < 0) {                      // the reference looks the chromosome up in the BAM header and fails: no reads here
        if (n_reads) std::fill(n_reads, n_reads + R, (int64_t)0);
        return CPN_OK;
    }
    RegionJob jb{bam, ref, pe != 0, {trim[0], trim[1], trim[2], trim[3]}};
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min<int>(nt, (int)std::max<int64_t>(1, R / 16)));
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    std::vector<std::string> errs(nt);
    auto work = [&](int tid) {
        std::vector<int64_t> kept;
        std::string meth;
        for (;;) {
            const int64_t r = next.fetch_add(1);
            if (r >= R || failed.load()) break;
            const int64_t N = cpg_off[r + 1] - cpg_off[r];
            int8_t* rows = calls ? calls + cell_off[r] : nullptr;       // region r's rows start at element Σ_{q<r} m_q·N_q
            const int64_t m = region_calls(jb, roi_st[r], roi_end[r], cpg_pos + cpg_off[r], N, rows, errs[tid], kept, meth);
            if (m < 0) { failed.store(1); break; }
            if (n_reads) n_reads[r] = m;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    if (failed.load()) {
        for (auto& e : errs) if (!e.empty()) { bam->err = e; break; }
        return CPN_ERR_ARG;
    }
    return CPN_OK;
}

}  // namespace

extern "C" {

const char* cpn_bam_open_error(void) { return g_open_err.c_str(); }
const char* cpn_bam_error(const cpn_bam* b) { return b ? b->err.c_str() : "null handle"; }

int cpn_bam_open(const char* path, int32_t n_threads, cpn_bam** out) {
    if (!path || !out) { g_open_err = "null argument"; return CPN_ERR_ARG; }
    *out = nullptr;
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) { g_open_err = std::string("cannot open ") + path; return CPN_ERR_ARG; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size <= 0) { ::close(fd); g_open_err = std::string("cannot stat ") + path; return CPN_ERR_ARG; }
    const size_t n = (size_t)sb.st_size;
    void* mp = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (mp == MAP_FAILED) { g_open_err = std::string("cannot map ") + path; return CPN_ERR_ARG; }
    const uint8_t* f = (const uint8_t*)mp;
    std::vector<Block> blocks;
    std::string err;
    auto bam = new cpn_bam();
    bool ok = scan_bgzf(f, n, blocks, err);
    if (ok) {
        const size_t total = blocks.empty() ? 0 : blocks.back().out_off + blocks.back().out_len;
        bam->data.resize(total);
        int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        nt = std::max(1, std::min<int>(nt, (int)std::max<size_t>(1, blocks.size() / 8)));
        std::atomic<size_t> next{0};
        std::atomic<int> bad{0};
        auto work = [&]() {
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= blocks.size()) break;
                const Block& bl = blocks[k];
                if (!inflate_block(f + bl.in_off, bl.in_len, bam->data.data() + bl.out_off, bl.out_len)) bad.store(1);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        if (bad.load()) { ok = false; err = "a BGZF block does not inflate to its recorded size"; }
    }
    munmap(mp, n);
    if (!ok) { g_open_err = err; delete bam; return CPN_ERR_ARG; }
    // header
    const uint8_t* d = bam->data.data();
    const size_t D = bam->data.size();
    auto fail = [&](const char* what) { g_open_err = what; delete bam; return CPN_ERR_ARG; };
    if (D < 12 || std::memcmp(d, "BAM\1", 4) != 0) return fail("not a BAM file (magic)");
    size_t p = 4;
    const int64_t l_text = rdi32(d + p); p += 4;
    if (l_text < 0 || p + (size_t)l_text + 4 > D) return fail("truncated BAM header");
    p += (size_t)l_text;
    const int64_t n_ref = rdi32(d + p); p += 4;
    if (n_ref < 0) return fail("negative reference count");
    for (int64_t i = 0; i < n_ref; ++i) {
        if (p + 4 > D) return fail("truncated reference list");
        const int64_t l_name = rdi32(d + p); p += 4;
        if (l_name < 1 || p + (size_t)l_name + 4 > D) return fail("truncated reference list");
        bam->ref_names.emplace_back((const char*)d + p, (size_t)l_name - 1); p += (size_t)l_name;
        bam->ref_lens.push_back(rdi32(d + p)); p += 4;
    }
    bam->by_ref.resize((size_t)n_ref); bam->by_ref_pos.resize((size_t)n_ref); bam->max_span.assign((size_t)n_ref, 0);
    if (D > 0xffffffffull) return fail("BAM streams beyond 4 GiB inflated are not supported by this reader (32-bit record offsets)");
    // records
    while (p < D) {
        if (p + 4 > D) return fail("truncated record");
        const int64_t bs = rdi32(d + p); p += 4;
        if (bs < 32 || p + (size_t)bs > D) return fail("truncated record");
        const uint8_t* r = d + p;
        RecIx x{};
        x.ref = rdi32(r);
        x.pos1 = (int64_t)rdi32(r + 4) + 1;
        const int l_rn = r[8];
        x.mapq = r[9];
        const int n_cig = rd16(r + 12);
        x.flag = rd16(r + 14);
        const int64_t l_seq = rdi32(r + 16);
        size_t q = 32;
        if (l_rn < 1 || l_seq < 0 || q + (size_t)l_rn + 4 * (size_t)n_cig + (size_t)(l_seq + 1) / 2 + (size_t)l_seq > (size_t)bs) return fail("malformed record");
        x.qname_off = (uint32_t)(p + q); x.qname_len = (uint32_t)(l_rn - 1);
        q += (size_t)l_rn;
        x.right1 = x.pos1 + align_length(r + q, n_cig) - 1;
        q += 4 * (size_t)n_cig + (size_t)(l_seq + 1) / 2 + (size_t)l_seq;
        if (!scan_aux(r + q, r + bs, d, x)) return fail("malformed auxiliary field");
        if (x.ref >= 0 && x.ref < n_ref) bam->by_ref[(size_t)x.ref].push_back((int64_t)bam->recs.size());
        bam->recs.push_back(x);
        p += (size_t)bs;
    }
    for (size_t i = 0; i < (size_t)n_ref; ++i) {
        auto& ix = bam->by_ref[i];
        std::stable_sort(ix.begin(), ix.end(), [&](int64_t a, int64_t b) { return bam->recs[a].pos1 < bam->recs[b].pos1; });
        auto& ps = bam->by_ref_pos[i];
        ps.reserve(ix.size());
        for (int64_t k : ix) {
            ps.push_back(bam->recs[k].pos1);
            bam->max_span[i] = std::max(bam->max_span[i], bam->recs[k].right1 - bam->recs[k].pos1 + 1);
        }
    }
    *out = bam;
    return CPN_OK;
}

void cpn_bam_close(cpn_bam* b) { delete b; }
int64_t cpn_bam_num_records(const cpn_bam* b) { return b ? (int64_t)b->recs.size() : 0; }
int64_t cpn_bam_num_refs(const cpn_bam* b) { return b ? (int64_t)b->ref_names.size() : 0; }
const char* cpn_bam_ref_name(const cpn_bam* b, int64_t i) { return (b && i >= 0 && i < (int64_t)b->ref_names.size()) ? b->ref_names[(size_t)i].c_str() : ""; }
int64_t cpn_bam_ref_len(const cpn_bam* b, int64_t i) { return (b && i >= 0 && i < (int64_t)b->ref_lens.size()) ? b->ref_lens[(size_t)i] : -1; }

int cpn_bam_record(const cpn_bam* b, int64_t i, int32_t* ref, int64_t* pos, int64_t* rightpos, int32_t* flag, int32_t* mapq,
                   int32_t* has_xm, int32_t* has_xs, const char** qname, int64_t* qname_len, const char** xm, int64_t* xm_len) {
    if (!b || i < 0 || i >= (int64_t)b->recs.size()) return CPN_ERR_ARG;
    const RecIx& r = b->recs[(size_t)i];
    if (ref) *ref = r.ref;
    if (pos) *pos = r.pos1;
    if (rightpos) *rightpos = r.right1;
    if (flag) *flag = r.flag;
    if (mapq) *mapq = r.mapq;
    if (has_xm) *has_xm = r.has_xm;
    if (has_xs) *has_xs = r.has_xs;
    if (qname) *qname = (const char*)b->data.data() + r.qname_off;
    if (qname_len) *qname_len = r.qname_len;
    if (xm) *xm = (const char*)b->data.data() + r.xm_off;
    if (xm_len) *xm_len = r.xm_len;
    return CPN_OK;
}

int cpn_bam_count_reads(const cpn_bam* bam, const char* chr, int64_t R, const int64_t* roi_st, const int64_t* roi_end, const int64_t* cpg_off,
                        const int64_t* cpg_pos, int32_t pe, const int64_t* trim, int32_t n_threads, int64_t* n_reads) {
    if (bam && !n_reads && R > 0) { bam->err = "null pointer"; return CPN_ERR_ARG; }
    return run_regions(bam, chr, R, roi_st, roi_end, cpg_off, cpg_pos, pe, trim, n_threads, n_reads, nullptr, nullptr);
}

int cpn_bam_fill_calls(const cpn_bam* bam, const char* chr, int64_t R, const int64_t* roi_st, const int64_t* roi_end, const int64_t* cpg_off,
                       const int64_t* cpg_pos, int32_t pe, const int64_t* trim, int32_t n_threads, const int64_t* cell_off, int8_t* calls) {
    if (bam && R > 0 && (!cell_off || !calls)) { bam->err = "null pointer"; return CPN_ERR_ARG; }
    return run_regions(bam, chr, R, roi_st, roi_end, cpg_off, cpg_pos, pe, trim, n_threads, nullptr, cell_off, calls);
}

}  // extern "C"
