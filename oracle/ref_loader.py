"""Import the UNMODIFIED reference modules: from /root/reference in the build container, else from the byte-for-byte
copies staged under oracle/_ref/ by `python -m oracle.make_ref` (git-ignored; they travel to the GPU box).

TEST INFRASTRUCTURE.  Callers must check `reference_available()` and skip otherwise.  gymnasium / pygame are
not installed in this image, so stub packages from oracle/refstubs/ are put on
sys.path for the duration of the import only.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref")


def _pick_root() -> str:
    env = os.environ.get("BENG_REFERENCE_ROOT")
    if env:
        return env
    for root in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(root, "snake_env_classic", "snake_env.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _pick_root()
_STUBS = os.path.join(_HERE, "refstubs")
_cache = {}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "snake_env_classic", "snake_env.py"))


def _stub_names():
    names = []
    for n in ("gymnasium", "pygame", "matplotlib"):
        if importlib.util.find_spec(n) is None or n == "matplotlib":
            names.append(n)
    return names


def _import_from(path: str, modname: str, extra_sys_path=()):
    """Load file `path` as module `modname` with the stubs visible."""
    if modname in _cache:
        return _cache[modname]
    saved_path = list(sys.path)
    stub_roots = _stub_names()
    stubbed = {}
    try:
        sys.path[:0] = [_STUBS, *extra_sys_path]
        for n in list(sys.modules):
            if n.split(".")[0] in stub_roots:
                stubbed[n] = sys.modules.pop(n)  # a real (or earlier stub) module: set aside for the import
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved_path
        # The reference module keeps its own references to the stubs it imported; sys.modules goes back to what it
        # was, so a later `import matplotlib` / `import gymnasium` elsewhere in the process never sees a stub.
        for n in list(sys.modules):
            if n.split(".")[0] in stub_roots and _is_stub(sys.modules[n]):
                del sys.modules[n]
        sys.modules.update(stubbed)
    _cache[modname] = mod
    return mod


def _is_stub(mod) -> bool:
    f = getattr(mod, "__file__", None) or ""
    return os.path.abspath(f).startswith(_STUBS + os.sep)


def load_snake():
    """-> the reference module snake_env_classic/snake_env.py (class SnakeEnvClassic)."""
    return _import_from(os.path.join(REFERENCE_ROOT, "snake_env_classic", "snake_env.py"), "_ref_snake_env")


def load_crypto():
    """-> the reference module crypto_trading_env/crypto_trading_env.py."""
    return _import_from(
        os.path.join(REFERENCE_ROOT, "crypto_trading_env", "crypto_trading_env.py"), "_ref_crypto_trading_env"
    )


def load_traffic():
    """-> (environment, utils) reference modules of traffic_management_env.

    environment.py imports `config` and `utils` as TOP-LEVEL names (environment.py:16,23), so the
    package directory itself goes on sys.path for the import and the two generic names are removed
    from sys.modules afterwards (bus_system_env defines the same names; SURVEY.md section 2 note).
    """
    if "_ref_traffic_environment" in _cache:
        return _cache["_ref_traffic_environment"], _cache["_ref_traffic_utils"]
    d = os.path.join(REFERENCE_ROOT, "traffic_management_env")
    for n in ("config", "utils"):
        sys.modules.pop(n, None)
    env_mod = _import_from(os.path.join(d, "environment.py"), "_ref_traffic_environment", extra_sys_path=(d,))
    utils_mod = sys.modules.get("utils")
    _cache["_ref_traffic_utils"] = utils_mod
    for n in ("config", "utils"):
        sys.modules.pop(n, None)
    return env_mod, utils_mod


def load_climate():
    """-> the reference module smartclimate_rl-main/smartclimate/env.py (class SmartClimateEnv).  The package's
    __init__ registers with gymnasium and imports itself; only env.py and utils.py are needed, so they are loaded as
    a synthetic package `_ref_smartclimate` (utils first: env.py does `from .utils import ...`)."""
    if "_ref_smartclimate.env" in _cache:
        return _cache["_ref_smartclimate.env"]
    import types

    d = os.path.join(REFERENCE_ROOT, "smartclimate_rl-main", "smartclimate")
    pkg = types.ModuleType("_ref_smartclimate")
    pkg.__path__ = [d]
    sys.modules["_ref_smartclimate"] = pkg
    _import_from(os.path.join(d, "utils.py"), "_ref_smartclimate.utils")
    return _import_from(os.path.join(d, "env.py"), "_ref_smartclimate.env")


def load_builder():
    """-> (world_builder_env module, game_logic module) of world_builder_env/src/environment.  The package's renderer
    imports pygame at top level (stubbed); the modules are loaded as a synthetic package `_ref_builder`."""
    if "_ref_builder.world_builder_env" in _cache:
        return _cache["_ref_builder.world_builder_env"], _cache["_ref_builder.game_logic"]
    import types

    d = os.path.join(REFERENCE_ROOT, "world_builder_env", "src", "environment")
    pkg = types.ModuleType("_ref_builder")
    pkg.__path__ = [d]
    sys.modules["_ref_builder"] = pkg
    gl = _import_from(os.path.join(d, "game_logic.py"), "_ref_builder.game_logic")
    _import_from(os.path.join(d, "renderer.py"), "_ref_builder.renderer")
    env = _import_from(os.path.join(d, "world_builder_env.py"), "_ref_builder.world_builder_env")
    return env, gl
