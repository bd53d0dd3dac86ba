"""Parser for the ``fix ID group conp|conq|cond ...`` command line, keeping
the reference syntax and error messages (fix_conp.cpp:79-176).

    fix ID group-ID conp Nevery group2-ID eta dV|v_name logfile [keywords]

keywords: ffield noslab org <file> inv <file> etypes <n> <t1..tn> zneutr
matout pppm split qinit himem nonneutral ehgo
"""
from __future__ import annotations

import dataclasses

FF_NORMAL, FF_FFIELD, FF_NOSLAB = 0, 1, 2  # fix_conp.cpp:68
PAIR_ETA, PAIR_EHGO = 0, 1  # fix_conp.cpp:69
VARIANT_CONP, VARIANT_CONQ, VARIANT_COND = 0, 1, 2


class FixError(RuntimeError):
    """Stands in for LAMMPS' ``error->all(FLERR, msg)``."""


@dataclasses.dataclass
class FixArgs:
    fix_id: str
    group: str
    style: str
    everynum: int
    group2: str
    eta: float
    potdiff: float | None
    potdiffstr: str | None
    logfile: str
    ff_flag: int = FF_NORMAL
    a_matrix_f: int = 0  # 0 none, 1 org, 2 inv
    a_matrix_file: str | None = None
    smartlist: bool = False
    eletypes: tuple = ()
    zneutrflag: bool = False
    matoutflag: bool = False
    pppmflag: bool = False
    splitflag: bool = False
    qinitflag: bool = False
    lowmemflag: bool = True
    nullneutralflag: bool = True
    pairmode: int = PAIR_ETA

    @property
    def variant(self) -> int:
        return {"conp": VARIANT_CONP, "conq": VARIANT_CONQ, "cond": VARIANT_COND}[self.style]


def parse_fix_args(arg, ntypes: int) -> FixArgs:
    """``arg`` is the full token list starting at the fix ID, as LAMMPS passes
    it to the constructor (narg counts all of them)."""
    arg = [str(a) for a in arg]
    narg = len(arg)
    if narg < 8:
        raise FixError("Illegal fix conp command (too few input parameters)")
    style = arg[2]
    if style not in ("conp", "conq", "cond"):
        raise FixError(f"Unknown fix style {style}")
    try:
        everynum = int(arg[3])
        eta = float(arg[5])
    except ValueError as e:
        raise FixError(f"Expected number in fix conp command: {e}") from None
    potdiff, potdiffstr = None, None
    if arg[6].startswith("v_"):
        potdiffstr = arg[6][2:]
    else:
        try:
            potdiff = float(arg[6])
        except ValueError as e:
            raise FixError(f"Expected number in fix conp command: {e}") from None
    fa = FixArgs(arg[0], arg[1], style, everynum, arg[4], eta, potdiff, potdiffstr, arg[7])
    iarg = 8
    while iarg < narg:
        a = arg[iarg]
        if a == "ffield":
            if fa.ff_flag == FF_NOSLAB:
                raise FixError("Invalid fix conp command (ffield and noslab cannot both be chosen)")
            fa.ff_flag = FF_FFIELD
        elif a == "noslab":
            if fa.ff_flag == FF_FFIELD:
                raise FixError("Invalid fix conp command (ffield and noslab cannot both be chosen)")
            fa.ff_flag = FF_NOSLAB
        elif a in ("org", "inv"):
            if fa.a_matrix_f != 0:
                raise FixError("Invalid fix conp command (A matrix file specified more than once)")
            fa.a_matrix_f = 1 if a == "org" else 2
            iarg += 1
            if iarg >= narg:
                raise FixError("Invalid fix conp command (No A matrix filename given)")
            fa.a_matrix_file = arg[iarg]
        elif a == "etypes":
            iarg += 1
            if iarg >= narg - 1:
                raise FixError("Invalid fix conp command (Insufficient input entries for etypes)")
            n = int(arg[iarg])
            types = []
            for _ in range(n):
                iarg += 1
                if iarg >= narg:
                    raise FixError("Invalid fix conp command (Insufficient input entries for etypes)")
                types.append(int(arg[iarg]))
            for t in types:
                if t > ntypes:
                    raise FixError("Invalid fix conp command (Invalid atom type in etypes)")
            fa.eletypes = tuple(types)
            fa.smartlist = True
        elif a == "zneutr":
            fa.zneutrflag = True
        elif a == "matout":
            fa.matoutflag = True
        elif a == "pppm":
            fa.pppmflag = True
        elif a == "split":
            fa.splitflag = True
        elif a == "qinit":
            fa.qinitflag = True
        elif a == "himem":
            fa.lowmemflag = False
        elif a == "nonneutral":
            fa.nullneutralflag = False
        elif a == "ehgo":
            fa.pairmode = PAIR_EHGO
        else:
            raise FixError(f"Invalid fix conp commmand (unknown option: {a})")
        iarg += 1
    return fa
