"""TEST / BENCH INFRASTRUCTURE — not product code.

Stages the UNMODIFIED reference (microsoft/StemGNN, read-only at /root/reference) into the git-ignored
directory `baseline/_ref/` so that it travels to the GPU box with the gpurun snapshot (the box has no
/root/reference).  Nothing is edited: files are byte-for-byte copies, and `baseline/_ref/` is listed in
.gitignore, so no reference source ever enters the repository history.

    python -m oracle.fetch_reference          # idempotent; prints what it staged

What is staged: the three packages the hot path and its callers live in (`models/`, `data_loader/`,
`utils/`), `main.py`, and `dataset/ECG_data.csv` (BASELINE.json configs[0]).  PeMS07.csv (69 MB) is
staged only with --with-pems07.

Consumers (all test/bench infrastructure): `oracle/ref_shim.py` (imports the reference under the
4-patch shim), `bench.py --impl reference` / `cpu_baseline`, `run_main.py`, `tests/`.
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("STEMGNN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

ITEMS = ["models", "data_loader", "utils", "main.py", os.path.join("dataset", "ECG_data.csv")]


def _copy(rel):
    s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
    if os.path.isdir(s):
        n = 0
        for base, _dirs, files in os.walk(s):
            for f in files:
                if f.endswith(".pyc"):
                    continue
                n += _copy(os.path.relpath(os.path.join(base, f), SRC))
        return n
    os.makedirs(os.path.dirname(d), exist_ok=True)
    if os.path.exists(d) and filecmp.cmp(s, d, shallow=False):
        return 0
    shutil.copyfile(s, d)
    return 1


def fetch(with_pems07=False, quiet=False):
    """Returns the staged directory, or None when the reference is not mounted here."""
    if not os.path.isfile(os.path.join(SRC, "models", "base_model.py")):
        return DST if os.path.isfile(os.path.join(DST, "models", "base_model.py")) else None
    items = list(ITEMS) + ([os.path.join("dataset", "PeMS07.csv")] if with_pems07 else [])
    n = sum(_copy(rel) for rel in items)
    if not quiet:
        print(f"baseline/_ref: {n} file(s) staged from {SRC}")
    return DST


if __name__ == "__main__":
    fetch(with_pems07="--with-pems07" in sys.argv)
