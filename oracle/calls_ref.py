"""calls_ref.py — line-by-line restatement of the reference's calls ingestion.  TEST INFRASTRUCTURE ONLY — NOT PRODUCT CODE.

Follows get_dic_nanopolish (src/input_output.jl:536-570: the whole TSV is read for every region; header skipped; a line is kept
when chr matches, start <= chrend and end >= chrst) and get_calls_nanopolish! (src/nanopore.jl:172-230: group found by a linear
search for grp_st == start+1-5 and grp_end == end+1+5, both must name the same group; later lines overwrite; mod_nse = false
stores (0, -50)).  Only tests/ and tools/ import it; the product's loaders (csrc/cpn_calls.cpp, cpelnano.jl_b200/nanopolish.py)
are compared against it.
"""
import gzip

import numpy as np

LOG_PYX_RIGHT_X = 0.0       # src/CpelNano.jl:33
LOG_PYX_WRONG_X = -50.0     # src/CpelNano.jl:32


def region_calls_reference(path, chrom, chrst, chrend, grp_st, grp_end, mod_nse=True):
    """Line-by-line restatement of get_dic_nanopolish + get_calls_nanopolish! for ONE region (test oracle: O(file) per
    region like the reference).  Returns {read name: (lu[L], lm[L])}."""
    opener = gzip.open if str(path).endswith(".gz") else open
    dic = {}
    with opener(path, "rt") as f:
        for i, line in enumerate(f):
            if i == 0:
                continue
            v = line.rstrip("\n").split("\t")
            if v[0] != chrom:
                continue
            if not int(v[2]) <= chrend:
                continue
            if not int(v[3]) >= chrst:
                continue
            dic.setdefault(v[4], []).append(v)
    out = {}
    L = len(grp_st)
    for name, lines in dic.items():
        lu = [float("nan")] * L; lm = [float("nan")] * L
        for v in lines:
            st, en = int(v[2]) + 1 - 5, int(v[3]) + 1 + 5
            i_st = next((i for i in range(L) if grp_st[i] == st), None)
            i_en = next((i for i in range(L) if grp_end[i] == en), None)
            if i_st is None or i_en is None or i_st != i_en:
                continue
            m, u = float(v[6]), float(v[7])
            if not mod_nse:
                m, u = (LOG_PYX_RIGHT_X, LOG_PYX_WRONG_X) if m > u else (LOG_PYX_WRONG_X, LOG_PYX_RIGHT_X)
            lm[i_st], lu[i_st] = m, u
        out[name] = (np.array(lu), np.array(lm))
    return out
