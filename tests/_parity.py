"""Parity bookkeeping for the GPU tests: every dense comparison goes through check(), which records the MEASURED maximum error next to
the tolerance it was asserted against.  conftest.py writes the table to gpurun_out/parity_errors.json at the end of the session (the
committed copy is profiles/r02_parity_errors.json), so the slack of every tolerance is on record."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("AIRFE_PARITY_OUT", os.path.join(ROOT, "gpurun_out", "parity_errors.json"))
_REC = {}


def _slot(name, tol, unit, note):
    r = _REC.setdefault(name, {"max_measured": 0.0, "tolerance": tol, "unit": unit, "checks": 0, "note": note})
    r["tolerance"] = tol
    if note:
        r["note"] = note
    return r


def check(name, measured, tol, unit="abs", note=""):
    """Record `measured` under `name` and assert measured <= tol."""
    measured = float(measured)
    r = _slot(name, float(tol), unit, note)
    r["max_measured"] = max(r["max_measured"], measured)
    r["checks"] += 1
    assert measured <= tol, "%s: measured %.4e exceeds tolerance %.4e (%s)" % (name, measured, tol, unit)


def report(name, value, unit="", note=""):
    """Record an informative figure (set overlaps, counts): no assertion."""
    r = _REC.setdefault(name, {"values": [], "unit": unit, "note": note})
    r.setdefault("values", []).append(float(value))


def exact(name, ok, note=""):
    """Record a bit-exact comparison (integer / index / discrete stage)."""
    r = _REC.setdefault(name, {"exact": True, "checks": 0, "note": note})
    r["checks"] = r.get("checks", 0) + 1
    assert ok, "%s: not bit-exact" % name


def dump():
    if not _REC:
        return
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    old = {}
    if os.path.exists(OUT):
        try:
            old = json.load(open(OUT))
        except Exception:
            old = {}
    old.update(_REC)
    with open(OUT, "w") as f:
        json.dump(old, f, indent=1, sort_keys=True)
