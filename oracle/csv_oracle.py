"""CPU oracle for the reference's on-disk formats -- TEST INFRASTRUCTURE, NOT
PRODUCT CODE.

What the reference does with its two files (paths relative to /root/reference):

  <id>.csv        pd.read_csv, then sort_values("Timestamp")     vmvo/datasets/bdd/bdd_raw.py:53-55
  <id>_traj.csv   pd.read_csv; rot = str(3x3 ndarray) un-stringified by parse_rot
                  (strip brackets / newlines, split, astype(float32).reshape(3, 3));
                  Timestamp copied over from the log                          bdd_raw.py:150-168

The arithmetic lives in a third-party dependency that is absent from /root/reference: **pandas**
(requirements.txt:9, unpinned; 3.0.2 in this image).  ``read_csv``'s C tokenizer converts numbers
with ``precise_xstrtod`` (pandas/_libs/src/parser/tokenizer.c; the default, float_precision=None):
up to 17 digits accumulated in a double by ``number = number * 10 + digit``, then ONE multiplication
or division by a tabulated power of ten -- exact for <= 15 digits, off by an ulp or more for many
16-17 digit inputs.  ``precise_xstrtod`` below restates that published algorithm; the tests pin it
against ``pandas.read_csv`` itself (importable here and on the GPU box) and against vectors frozen
from the reference's own ``AndroidDatasetIterator`` (oracle/make_golden_csv.py ->
tests/golden/csv_kats.json).  ``parse_rot`` is a literal restatement of bdd_raw.py:157-165 (NumPy's
string -> float32 cast goes through Python's correctly rounded ``float``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

POW10 = [float("1e%d" % i) for i in range(309)]
NA_VALUES = {"", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN",
             "<NA>", "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"}
_SPACE = " \t\n\v\f\r"


def precise_xstrtod(s: str) -> Tuple[bool, float]:
    """pandas' default number converter; (False, nan) where it reports an error."""
    p, n = 0, len(s)
    while p < n and s[p] in _SPACE:
        p += 1
    neg = False
    if p < n and s[p] in "+-":
        neg = s[p] == "-"
        p += 1
    number, exponent, nd, ndec, max_digits = 0.0, 0, 0, 0, 17
    while p < n and s[p].isdigit() and s[p].isascii():
        if nd < max_digits:
            number = number * 10.0 + (ord(s[p]) - 48)
            nd += 1
        else:
            exponent += 1
        p += 1
    if p < n and s[p] == ".":
        p += 1
        while nd < max_digits and p < n and s[p].isdigit() and s[p].isascii():
            number = number * 10.0 + (ord(s[p]) - 48)
            p += 1
            nd += 1
            ndec += 1
        if nd >= max_digits:
            while p < n and s[p].isdigit() and s[p].isascii():
                p += 1
        exponent -= ndec
    if nd == 0:
        return False, float("nan")
    if neg:
        number = -number
    if p < n and s[p] in "eE":
        save = p
        p += 1
        eneg = False
        if p < n and s[p] in "+-":
            eneg = s[p] == "-"
            p += 1
        k = ed = 0
        while ed < max_digits and p < n and s[p].isdigit() and s[p].isascii():
            k = k * 10 + ord(s[p]) - 48
            ed += 1
            p += 1
        exponent += -k if eneg else k
        if ed == 0:
            p = save
    if exponent > 308:                       # ERANGE, delivered as a signed infinity
        number = float("-inf") if neg else float("inf")
    elif exponent > 0:
        number *= POW10[exponent]
    elif exponent < -308:
        if exponent < -616:
            number = 0.0
        else:
            number /= POW10[-308 - exponent]
            number /= POW10[308]
    else:
        number /= POW10[-exponent]
    while p < n and s[p] in _SPACE:
        p += 1
    if p != n:
        return False, float("nan")
    return True, number


def to_double(field: str) -> Tuple[bool, float]:
    """One field of a numeric column: NA words, the converter, then the infinity spellings."""
    if field in NA_VALUES:
        return True, float("nan")
    ok, v = precise_xstrtod(field)
    if ok:
        return True, v
    w = field                                 # the whole field, white space included
    sign = 1.0
    if w[:1] in ("+", "-"):
        sign = -1.0 if w[0] == "-" else 1.0
        w = w[1:]
    if w.lower() in ("inf", "infinity"):
        return True, sign * float("inf")
    return False, float("nan")


def split_rows(data: bytes) -> List[bytes]:
    """Lines of a CSV file: newlines inside quoted fields do not end a line, blank lines are skipped
    (pandas skip_blank_lines), a trailing carriage return is dropped."""
    rows, cur, inq = [], bytearray(), False
    for c in data:
        if c == 0x22:
            inq = not inq
        if c == 0x0A and not inq:
            rows.append(bytes(cur))
            cur = bytearray()
        else:
            cur.append(c)
    if cur:
        rows.append(bytes(cur))
    out = []
    for r in rows:
        r = r.rstrip(b"\r")
        if r:
            out.append(r)
    return out


def split_fields(row: bytes) -> List[str]:
    """Fields of one line: a field that starts with a quote runs to the closing quote, "" inside it
    is a literal quote."""
    s = row.decode("utf-8")
    out, i, n = [], 0, len(s)
    while True:
        if i < n and s[i] == '"':
            i += 1
            buf = []
            while i < n:
                if s[i] == '"':
                    if i + 1 < n and s[i + 1] == '"':
                        buf.append('"')
                        i += 2
                        continue
                    i += 1
                    break
                buf.append(s[i])
                i += 1
            while i < n and s[i] != ",":
                i += 1
            out.append("".join(buf))
        else:
            j = s.find(",", i)
            j = n if j < 0 else j
            out.append(s[i:j])
            i = j
        if i >= n:
            break
        i += 1                                # the comma
        if i == n:
            out.append("")
            break
    return out


def parse_rot(rot):
    """bdd_raw.py:157-165, verbatim in behaviour."""
    if type(rot) == str:  # noqa: E721
        rot = rot.replace("[", "").replace("]", "").replace("\n", "")
        rot = rot.split()
        rot = np.array(rot).astype(np.float32).reshape(3, 3)
    return rot


def read_csv(data: bytes, wanted: Optional[Sequence[str]] = None,
             rot_column: Optional[str] = None) -> Dict[str, object]:
    """Columns of one file: float64 arrays (NaN for NA / missing fields) and, for ``rot_column``, a
    list of float32 3x3 arrays.  Raises ValueError where pandas would not deliver a numeric column
    and on rows with more fields than the header."""
    rows = split_rows(data)
    if not rows:
        raise ValueError("No columns to parse from file")
    header = split_fields(rows[0])
    if wanted is None:
        wanted = [h for h in header if h != rot_column]
    cols: Dict[str, List] = {w: [] for w in wanted}
    rots: List[np.ndarray] = []
    for r in rows[1:]:
        fields = split_fields(r)
        if len(fields) > len(header):
            raise ValueError(f"Expected {len(header)} fields, saw {len(fields)}")
        fields += [""] * (len(header) - len(fields))
        for name in wanted:
            ok, v = to_double(fields[header.index(name)])
            if not ok:
                raise ValueError(f"column {name}: {fields[header.index(name)]!r} is not a number")
            cols[name].append(v)
        if rot_column is not None:
            rots.append(parse_rot(fields[header.index(rot_column)]))
    out: Dict[str, object] = {k: np.asarray(v, dtype=np.float64) for k, v in cols.items()}
    if rot_column is not None:
        out[rot_column] = rots
    return out
