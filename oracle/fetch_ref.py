"""ORACLE (test infrastructure) - recipe that ships the UNMODIFIED reference to the GPU box.

The reference (pop-pop-pOp-dev/SAC-GAT-HER_transportationRL) is pure Python.  `/root/reference`
exists only in the build container; `oracle/_ref/` is git-ignored (never enters history) but is NOT
gpurun-ignored, so a verbatim copy placed there travels with the snapshot exactly like the built
`.so` files do.  This script makes that copy:

    python oracle/fetch_ref.py            # copy when /root/reference is present, else keep what is there

What is copied (only what the hot path, its callers and the checkpoint test need - no outputs, maps
or TensorBoard logs): `src/` (4.5 kLoC of Python), `configs/`, `scripts/`, `data/SiouxFalls`,
`test_env_gpu.py`, `requirements.txt` and the two small SAC checkpoints under `history-data/outputs1`.
Nothing is edited; `MANIFEST.json` records the sha256 of every file so a test can prove it.

Users of the copy: tests/ (drop-in tests run the reference's own `src.baselines` / `src.eval` against
`b200env.RepairEnv`), bench.py's `--impl reference` arm and `cpu_baseline` legs.  The product package
never imports it (tests/test_host.py::test_product_does_not_import_oracle).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("B200_REFERENCE_ROOT", "/root/reference")

TREES = ["src", "configs", "scripts", os.path.join("data", "SiouxFalls")]
FILES = ["test_env_gpu.py", "requirements.txt", "README.md",
         os.path.join("history-data", "outputs1", "model_best.pt"),
         os.path.join("history-data", "outputs1", "outputs_smoke", "model_best.pt")]


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for chunk in iter(lambda: fh.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def _wanted():
    for tree in TREES:
        root = os.path.join(SOURCE, tree)
        for dirpath, dirnames, filenames in os.walk(root):
            dirnames[:] = [d for d in dirnames if d != "__pycache__"]
            for name in sorted(filenames):
                if not name.endswith(".pyc"):
                    yield os.path.relpath(os.path.join(dirpath, name), SOURCE)
    for f in FILES:
        if os.path.exists(os.path.join(SOURCE, f)):
            yield f


def fetch(force: bool = False) -> str | None:
    """Copy the reference into oracle/_ref/ (verbatim).  Returns the destination, or None when
    neither the source tree nor an earlier copy exists."""
    manifest_path = os.path.join(DEST, "MANIFEST.json")
    if not os.path.isdir(SOURCE):
        return DEST if os.path.exists(manifest_path) else None
    if os.path.exists(manifest_path) and not force:
        with open(manifest_path) as fh:
            have = json.load(fh)["files"]
        if all(os.path.exists(os.path.join(DEST, rel)) for rel in have) and \
                sorted(have) == sorted(_wanted()):
            return DEST
    files = {}
    for rel in _wanted():
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SOURCE, rel), dst)
        os.chmod(dst, 0o644)
        files[rel] = _sha(dst)
    with open(manifest_path, "w") as fh:
        json.dump({"source": SOURCE, "note": "verbatim copy of the reference, made by oracle/fetch_ref.py",
                   "files": files}, fh, indent=1, sort_keys=True)
    return DEST


def verify() -> bool:
    """True when every file of the copy still has the recorded sha256 (i.e. is unmodified)."""
    manifest_path = os.path.join(DEST, "MANIFEST.json")
    if not os.path.exists(manifest_path):
        return False
    with open(manifest_path) as fh:
        files = json.load(fh)["files"]
    return all(os.path.exists(os.path.join(DEST, rel)) and _sha(os.path.join(DEST, rel)) == sha
               for rel, sha in files.items())


if __name__ == "__main__":
    out = fetch(force="--force" in sys.argv)
    print(out if out else "reference not available (no /root/reference and no earlier copy)")
