"""Documents stay honest: every profiles/ file, script and test that DESIGN.md, README.md, INTEGRATION.md or
profiles/README.md name exists in the tree, and every C-ABI function they mention is declared in include/dgp_b200.h."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
DOCS = ["DESIGN.md", "README.md", "INTEGRATION.md", "profiles/README.md"]


def _text():
    return {d: (ROOT / d).read_text() for d in DOCS}


def test_referenced_files_exist():
    missing = []
    for doc, text in _text().items():
        for m in re.finditer(r"`((?:profiles|scripts|tests|oracle|include)/[A-Za-z0-9_./\-]+\.[a-z]+)`", text):
            path = m.group(1)
            if "*" in path or "{" in path:
                continue
            if not (ROOT / path).exists():
                missing.append((doc, path))
        if doc == "profiles/README.md":   # its table names files relative to profiles/
            for m in re.finditer(r"`([A-Za-z0-9_\-]+_r[12](?:_[a-z0-9]+)?\.(?:json|jsonl|txt))`", text):
                if not (ROOT / "profiles" / m.group(1)).exists():
                    missing.append((doc, "profiles/" + m.group(1)))
    assert not missing, missing


def test_mentioned_abi_functions_are_declared():
    header = (ROOT / "include" / "dgp_b200.h").read_text()
    declared = set(re.findall(r"\b(dgp_[a-z0-9_]+)\s*\(", header))
    allowed_prose = {"dgp_kernel_matrix_blocks"}      # the reviewer's name for dgp_data_kernel_block, quoted in a table
    unknown = []
    for doc, text in _text().items():
        for name in set(re.findall(r"`(dgp_[a-z0-9_]+)`", text)):
            if name not in declared and name not in allowed_prose and not name.endswith("_"):
                if name in ("dgp_ctx", "dgp_data", "dgp_model", "dgp_kernel_spec", "dgp_b200"):
                    continue
                unknown.append((doc, name))
    assert not unknown, unknown
