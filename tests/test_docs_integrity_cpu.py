"""Every `profiles/...` and `tests/...` path that README.md, DESIGN.md, INTEGRATION.md or profiles/README.md cite
exists in the tree, and every entry point of the header is mentioned in INTEGRATION.md or the header's own
documentation (so the coverage tables cannot silently drift from the repository)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cited_paths(md):
    txt = open(os.path.join(ROOT, md)).read()
    for m in re.finditer(r"`((?:profiles|tests|tools|mex|oracle|include|emagls_b200)/[A-Za-z0-9_./*-]+)`", txt):
        yield m.group(1)


def test_cited_files_exist():
    missing = []
    for md in ("README.md", "DESIGN.md", "INTEGRATION.md", os.path.join("profiles", "README.md")):
        for p in _cited_paths(md):
            p = p.split("::")[0].rstrip(".")
            if p == "oracle/_ref":        # cited only to say that it does not exist (the reference is MATLAB)
                continue
            if p.endswith("/"):
                ok = os.path.isdir(os.path.join(ROOT, p))
            elif "*" in p:
                ok = bool(glob.glob(os.path.join(ROOT, p)))
            else:
                ok = os.path.exists(os.path.join(ROOT, p)) or p.startswith("emagls_b200/lib/")   # built artefact
            if not ok:
                missing.append((md, p))
    assert not missing, missing


def test_every_reference_function_has_a_gateway_and_a_python_mirror():
    import emagls_b200.api as api
    names = ["getLsFilters", "getMagLsFilters", "getMagLsFilters2D", "getEMagLsFilters", "getEMagLs2Filters",
             "getEMagLsFiltersEMAinCH", "getEMagLsFiltersEMAinSH", "getEMagLsFiltersFromAtf", "getSMAIRMatrix",
             "binauralDecode", "getRadialFilter", "applyRadialFilter", "getMagLsSphericalHeadFilter",
             "getMagLsArrayDiffuseFilter"]
    for n in names:
        assert hasattr(api, n), n
        assert os.path.exists(os.path.join(ROOT, "mex", n + ".c")), n
    import oracle
    for n in names:
        assert hasattr(oracle, n), n
