"""Documentation that must not drift from the code: every runtime switch the sources read (DVID_* environment variables)
is listed in INTEGRATION.md's switch table, and every profile file the documents cite exists."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    with open(os.path.join(ROOT, *parts)) as f:
        return f.read()


def _sources():
    out = []
    for d, exts in (("diffusionvid_b200", (".py",)), (os.path.join("diffusionvid_b200", "csrc"), (".cu", ".cuh", ".h"))):
        for name in sorted(os.listdir(os.path.join(ROOT, d))):
            if name.endswith(exts):
                out.append(_read(d, name))
    return out


def test_every_env_switch_is_documented():
    used = set()
    for src in _sources():
        used |= set(re.findall(r'getenv\("(DVID_[A-Z0-9_]+)"\)', src))
        used |= set(re.findall(r'environ\.get\("(DVID_[A-Z0-9_]+)"', src))
    assert len(used) >= 30
    doc = _read("INTEGRATION.md")
    missing = sorted(v for v in used if v not in doc)
    assert not missing, "switches read by the code but absent from INTEGRATION.md: %s" % missing


def test_cited_profile_files_exist():
    cited = set()
    for doc in ("DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md")):
        cited |= set(re.findall(r"profiles/(r0[12][A-Za-z0-9_]*\.(?:json|txt|csv))", _read(doc)))
        if doc.startswith("profiles"):
            cited |= set(re.findall(r"`(r0[12][A-Za-z0-9_]*\.(?:json|txt|csv))`", _read(doc)))
    have = set(os.listdir(os.path.join(ROOT, "profiles")))
    missing = sorted(c for c in cited if c not in have)
    assert not missing, "documents cite profile files that are not committed: %s" % missing
