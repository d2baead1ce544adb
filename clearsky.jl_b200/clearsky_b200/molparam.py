"""MOLPARAM table (reference: src/hitran/molparam.jl, src/hitran/par.jl:18-48).

The numbers are data extracted verbatim from the reference by tools/extract_molparam.py into
data/molparam.json: abundances, molar masses and the Chebyshev coefficients of Qref/Q(T).  The fit is
only ~0.3 % accurate with respect to true TIPS, so parity needs these exact coefficients.
"""
import json
import os
from dataclasses import dataclass, field
from typing import List

MAXCHEB = 16  # padded coefficient row length shared with the C ABI (CS_MAXCHEB)


@dataclass
class MolParam:
    M: int = -1
    formula: str = ""
    name: str = ""
    I: List[int] = field(default_factory=list)
    isoform: List[str] = field(default_factory=list)
    AFGL: List[int] = field(default_factory=list)
    A: List[float] = field(default_factory=list)
    mu: List[float] = field(default_factory=list)
    Qref: List[float] = field(default_factory=list)
    hascheb: List[bool] = field(default_factory=list)
    ncheb: List[int] = field(default_factory=list)
    maxrelerr: List[float] = field(default_factory=list)
    cheb: List[List[float]] = field(default_factory=list)


def _load():
    fn = os.path.join(os.path.dirname(__file__), "data", "molparam.json")
    with open(fn) as f:
        d = json.load(f)
    out = []
    for e in d["MOLPARAM"]:
        out.append(MolParam() if e is None else MolParam(**e))
    return d["TMIN"], d["TMAX"], out


TMIN, TMAX, _TABLE = _load()


class _MolParamTable:
    """1-based like the Julia vector: MOLPARAM[M]"""

    def __getitem__(self, M):
        M = int(M)
        if M < 1 or M > len(_TABLE):
            raise IndexError(f"no molecule number {M}")
        return _TABLE[M - 1]

    def __len__(self):
        return len(_TABLE)

    def __iter__(self):
        return iter(_TABLE)


MOLPARAM = _MolParamTable()
