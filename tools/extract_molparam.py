#!/usr/bin/env python
"""Extract the MOLPARAM data table of the reference into JSON.

The table (reference: src/hitran/molparam.jl, generated offline by the reference's
scripts/molparam.py from HITRAN TIPS) is DATA: isotopologue abundances, molar masses and the
Chebyshev coefficients of Qref/Q(T).  The Chebyshev fit deviates from true TIPS by up to 0.4 %, so
parity with the reference requires these exact coefficients.

Run in the build container only (needs /root/reference):
    python tools/extract_molparam.py /root/reference/src/hitran/molparam.jl \
        clearsky.jl_b200/clearsky_b200/data/molparam.json
"""
import json
import re
import sys


def _split_top(s):
    """split a bracket-balanced string on top-level commas"""
    out, depth, cur, instr = [], 0, [], False
    for ch in s:
        if ch == '"':
            instr = not instr
        if not instr:
            if ch in "[(":
                depth += 1
            elif ch in "])":
                depth -= 1
            elif ch == "," and depth == 0:
                out.append("".join(cur).strip())
                cur = []
                continue
        cur.append(ch)
    tail = "".join(cur).strip()
    if tail:
        out.append(tail)
    return out


def _vec(tok):
    """parse 'Type[a, b, c]' (possibly nested) into python lists"""
    tok = tok.strip()
    m = re.match(r"^(Vector\{Float64\}|Float64|Int64|String|Bool)?\[(.*)\]$", tok, re.S)
    if not m:
        raise ValueError(tok[:60])
    kind, body = m.group(1), m.group(2)
    items = _split_top(body)
    if kind == "Vector{Float64}":
        return [_vec(t) for t in items]
    if kind == "String":
        return [t.strip().strip('"') for t in items]
    if kind == "Bool":
        return [t.strip() == "true" for t in items]
    if kind == "Int64":
        return [int(t) for t in items]
    return [float(t) for t in items]


def main(src, dst):
    text = open(src, encoding="utf-8").read()
    tmin = float(re.search(r"const TMIN = ([0-9.eE+-]+)", text).group(1))
    tmax = float(re.search(r"const TMAX = ([0-9.eE+-]+)", text).group(1))
    # strip comments
    text = re.sub(r"#[^\n]*", "", text)
    body = text[text.index("MolParam[") + len("MolParam["):]
    body = body[: body.rindex("]")]
    entries = []
    for tok in _split_top(body):
        m = re.match(r"^MolParam\((.*)\)$", tok.strip(), re.S)
        args = _split_top(m.group(1))
        if not args:  # MolParam() placeholder
            entries.append(None)
            continue
        (M, formula, name, I, isoform, afgl, A, mu, qref, hascheb, ncheb, maxrelerr, cheb) = args
        entries.append(dict(
            M=int(M), formula=formula.strip('"'), name=name.strip('"'),
            I=_vec(I), isoform=_vec(isoform), AFGL=_vec(afgl), A=_vec(A), mu=_vec(mu),
            Qref=_vec(qref), hascheb=_vec(hascheb), ncheb=_vec(ncheb),
            maxrelerr=_vec(maxrelerr), cheb=_vec(cheb)))
    out = dict(TMIN=tmin, TMAX=tmax, MOLPARAM=entries,
               source="markmbaum/ClearSky.jl src/hitran/molparam.jl (data table, extracted verbatim)")
    with open(dst, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    n = sum(e is not None for e in entries)
    print(f"{len(entries)} entries ({n} molecules) TMIN={tmin} TMAX={tmax} -> {dst}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
