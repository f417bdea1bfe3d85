"""Stage the UNMODIFIED reference sources of the hot path under oracle/_ref/ so that they travel to the GPU box.

TEST INFRASTRUCTURE.  `/root/reference` exists only in the build container; the GPU box receives a snapshot of this
repository.  `oracle/_ref/` is git-ignored (the history stays free of reference sources) but not gpurun-ignored, so the
byte-for-byte copies made here reach the box, where `oracle.ref_loader` imports them behind the stub
gymnasium/pygame/matplotlib packages exactly as it does from `/root/reference`.  They are used ONLY as the checker
(live-reference parity tests) and as the timed CPU arm of `bench.py` (`cpu_baseline.kind == "reference"`,
`--impl reference`); nothing in the product package reads them.

    python -m oracle.make_ref            # copy + write oracle/_ref/MANIFEST.json (sha256 of every file)
    python -m oracle.make_ref --check    # verify the staged copies against /root/reference (or the manifest)

Files (plus the two section 8(f) rank-3 envs):
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("BENG_REFERENCE_SOURCE", "/root/reference")

FILES = [
    "snake_env_classic/snake_env.py",
    "crypto_trading_env/crypto_trading_env.py",
    "traffic_management_env/environment.py",
    "traffic_management_env/utils.py",
    "traffic_management_env/config.py",
    "smartclimate_rl-main/smartclimate/env.py",
    "smartclimate_rl-main/smartclimate/utils.py",
    "world_builder_env/src/environment/world_builder_env.py",
    "world_builder_env/src/environment/game_logic.py",
    "world_builder_env/src/environment/renderer.py",
]
__doc__ += "".join(f"    {f}\n" for f in FILES)


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def source_available() -> bool:
    return os.path.isfile(os.path.join(SOURCE, FILES[0]))


def stage(verbose: bool = True) -> str:
    """Copy FILES verbatim from SOURCE to oracle/_ref/ and write the manifest.  No-op without SOURCE."""
    if not source_available():
        if verbose:
            print(f"make_ref: {SOURCE} is absent; keeping oracle/_ref as it is")
        return DEST
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SOURCE, "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"make_ref: staged {len(FILES)} unmodified reference files under {DEST}")
    return DEST


def check() -> bool:
    """True when every staged file matches the manifest (and SOURCE, when it is present)."""
    try:
        with open(os.path.join(DEST, "MANIFEST.json")) as f:
            manifest = json.load(f)["sha256"]
    except OSError:
        return False
    for rel in FILES:
        dst = os.path.join(DEST, rel)
        if not os.path.isfile(dst) or _sha(dst) != manifest.get(rel):
            return False
        if source_available() and _sha(os.path.join(SOURCE, rel)) != manifest[rel]:
            return False
    return True


if __name__ == "__main__":
    if "--check" in sys.argv:
        ok = check()
        print("oracle/_ref matches" if ok else "oracle/_ref is missing or differs")
        sys.exit(0 if ok else 1)
    stage()
