"""CPU restatement of the reference's WGBS calls front-end — TEST INFRASTRUCTURE ONLY (tests/ may import it; the product never does).

Follows src/wgbs.jl line by line, with Julia's 1-based string semantics spelled out:
    get_align_strand   src/wgbs.jl:23-55
    order_bams         src/wgbs.jl:66-78
    clean_records      src/wgbs.jl:93-118
    try_olaps          src/wgbs.jl:129-139   (BioAlignments' eachoverlap: position ≤ last(interval) and rightposition ≥ first(interval))
    get_calls_bam      src/wgbs.jl:222-323
The BAM decoding itself (XAM.jl / BGZFStreams.jl in the reference, not in the checkout) is restated from the SAM/BAM
specification with gzip + struct: BGZF is a multi-member gzip stream, which Python's gzip module reads as one stream.

Pinned by: the docstring example get_align_strand(true, 99, 147) == "OT" (:18-20), the reference's example
examples/wgbs (its expected theta file is empty: the 1 kb chromosome has 7 CpG sites in the analysed window, and
pmap_analyze_reg_bam returns before reading any call for ≤ 9 sites, :353), and hand-computed tables in tests/test_wgbs_bam.py.
"""
import gzip
import re
import struct

THRESH_MAPQ = 39                          # src/CpelNano.jl:30
FLAGS_ALLOWED = (0, 16, 83, 99, 147, 163)  # src/CpelNano.jl:31
REF_CONSUMING = (0, 2, 3, 7, 8)           # CIGAR M, D, N, =, X


class ReferenceExit(Exception):
    """The reference prints a message and calls exit(1) (src/wgbs.jl:31-33, 45-47)."""


def read_bam(path):
    """→ (refs [(name, length)], records [dict]) in file order.  Record fields as XAM's accessors return them: pos and rightpos
    1-based, flag, mapq, qname, tags {name: value} (only the presence of XS and the string value of XM matter here)."""
    d = gzip.open(path, "rb").read()
    assert d[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", d, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", d, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", d, p)[0]
        name = d[p + 4:p + 4 + l_name - 1].decode()
        refs.append((name, struct.unpack_from("<i", d, p + 4 + l_name)[0]))
        p += 8 + l_name
    recs = []
    while p < len(d):
        bs = struct.unpack_from("<i", d, p)[0]
        r = d[p + 4:p + 4 + bs]
        p += 4 + bs
        refid, pos0, l_rn, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", r, 0)
        q = 32
        qname = r[q:q + l_rn - 1].decode()
        q += l_rn
        alen = 0
        for k in range(n_cig):
            c = struct.unpack_from("<I", r, q + 4 * k)[0]
            if (c & 15) in REF_CONSUMING:
                alen += c >> 4
        q += 4 * n_cig + (l_seq + 1) // 2 + l_seq
        tags = {}
        while q < len(r):
            tag, ty = r[q:q + 2].decode(), chr(r[q + 2])
            q += 3
            if ty in "AcC":
                n, val = 1, r[q]
            elif ty in "sS":
                n, val = 2, struct.unpack_from("<h" if ty == "s" else "<H", r, q)[0]
            elif ty in "iI":
                n, val = 4, struct.unpack_from("<i" if ty == "i" else "<I", r, q)[0]
            elif ty == "f":
                n, val = 4, struct.unpack_from("<f", r, q)[0]
            elif ty in "ZH":
                e = r.index(b"\x00", q)
                n, val = e - q + 1, r[q:e].decode()
            elif ty == "B":
                st, cnt = chr(r[q]), struct.unpack_from("<I", r, q + 1)[0]
                n, val = 5 + cnt * {"c": 1, "C": 1, "s": 2, "S": 2}.get(st, 4), None
            else:
                raise ValueError(f"aux type {ty!r}")
            tags[tag] = val
            q += n
        recs.append(dict(ref=refid, pos=pos0 + 1, rightpos=pos0 + 1 + alen - 1, flag=flag, mapq=mapq, qname=qname, tags=tags))
    return refs, recs


def get_align_strand(pe, flag1, flag2):
    if not pe:
        if flag1 == 0:
            return "OT"
        if flag1 == 16:
            return "OB"
        raise ReferenceExit(f"SE: was expecting FLAG 0 or and encountered {flag1} instead.")
    if flag1 == 99 and flag2 == 147:
        return "OT"
    if flag1 == 83 and flag2 == 163:
        return "OB"
    if flag1 == 147 and flag2 == 99:
        return "CTOT"
    if flag1 == 163 and flag2 == 83:
        return "CTOB"
    raise ReferenceExit("PE: unexpected flag combination. Expected 99/147, 147/99, 83/163, 163/83.")


def order_bams(pe, records):
    if records[0]["pos"] <= records[1]["pos"]:
        return get_align_strand(pe, records[0]["flag"], records[1]["flag"]), records[0], records[1]
    return get_align_strand(pe, records[1]["flag"], records[0]["flag"]), records[1], records[0]


def clean_records(pe, records):
    templates = []
    if pe:
        names = []
        for x in records:                       # unique() keeps first appearances
            if x["qname"] not in names:
                names.append(x["qname"])
        for nm in names:
            temp_recs = [x for x in records if x["qname"] == nm]
            if len(temp_recs) == 2:
                templates.append(order_bams(pe, temp_recs))
    else:
        templates = [(get_align_strand(pe, x["flag"], 0), x, None) for x in records]
    return templates


def try_olaps(refs, recs, chrom, lo, hi):
    if hi < lo:                                 # minimum() of an empty range throws; caught → no records
        return []
    names = [n for n, _ in refs]
    ri = names.index(chrom)
    return [x for x in recs if x["ref"] == ri and x["pos"] <= hi and x["rightpos"] >= lo]


def _substring(s, lo, hi):
    """SubString(s, lo:hi), 1-based inclusive; an empty range is the empty string."""
    if hi < lo:
        return ""
    assert 1 <= lo and hi <= len(s)
    return s[lo - 1:hi]


def get_calls_bam(bam, chrom, roi_st, roi_end, cpg_pos, pe, trim):
    """→ (m, calls) with calls = list of per-template lists over the region's CpG sites: −1, 0, +1 (the reference stores
    MethCallCpgGrp objects with log p(y|x) ∈ {0, −50}; the integer is the same information, em_algorithm.jl:412-470).
    `bam` = (refs, recs) from read_bam."""
    refs, recs = bam
    names = [n for n, _ in refs]
    if chrom not in names:
        raise KeyError(chrom)                   # findfirst → nothing → indexing error in the reference
    chr_size = refs[names.index(chrom)][1]
    roi_end = min(roi_end, chr_size - 151)
    records_olap = try_olaps(refs, recs, chrom, roi_st, roi_end)
    if not records_olap:
        return 0, []
    records_olap = [x for x in records_olap if (x["flag"] & 4) == 0 and "XM" in x["tags"] and "XS" not in x["tags"]
                    and x["flag"] in FLAGS_ALLOWED and x["mapq"] > THRESH_MAPQ]
    calls = []
    for strand, R1, R2 in clean_records(pe, records_olap):
        if not pe:
            meth_call = R1["tags"]["XM"]
            meth_call = _substring(meth_call, 1 + trim[0], len(meth_call) - trim[1])
            OFFSET = 1 if strand in ("OT", "CTOT") else 2
            OFFSET -= R1["pos"] + trim[0]
        else:
            R1_call, R2_call = R1["tags"]["XM"], R2["tags"]["XM"]
            R1_call = _substring(R1_call, 1 + trim[0], len(R1_call) - trim[1])
            R2_call = _substring(R2_call, 1 + trim[3], len(R2_call) - trim[2])
            dist_st = abs(R2["pos"] + trim[3] - R1["pos"] - trim[0])
            meth_call = R1_call + "." * max(0, dist_st - len(R1_call))
            k = max(1, len(R1_call) + 1 - dist_st)
            if k > len(R2_call) + 1:
                raise IndexError("R2_call[k:end] with k beyond the string: BoundsError in the reference")
            meth_call += R2_call[k - 1:]
            OFFSET = 1 if strand in ("OT", "CTOT") else 2
            OFFSET -= R1["pos"] + trim[0]
        obs_cpgs = [m.start() + 1 - OFFSET for m in re.finditer("[zZ]", meth_call)]
        if not obs_cpgs:
            continue
        olap_cpgs = [i for i, c in enumerate(cpg_pos) if c in obs_cpgs]
        if not olap_cpgs:
            continue
        x = [0] * len(cpg_pos)
        idx = [c + OFFSET for c in obs_cpgs if c in cpg_pos]
        vals = [1 if meth_call[i - 1] == "Z" else -1 for i in idx]
        for i, v in zip(olap_cpgs, vals):
            x[i] = v
        calls.append(x)
    return len(calls), calls
