"""CPU checks of the measurement harness and the registration shim: the reference arm of bench.py runs here (it is the
Python port of the reference loop on the host cores) and prints the JSON line the driver parses; the registration ids
are the reference's."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly ONE JSON line on stdout"
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_b200_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without CUDA the product arm must fail loudly, not print a number."""
    import torch

    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert not any(ln.strip().startswith("{") and '"value"' in ln for ln in out.stdout.splitlines())


def test_registration_ids_match_the_reference(monkeypatch):
    registry, calls = {}, []

    def register(id, entry_point, max_episode_steps=None, **kw):  # noqa: A002 (gymnasium's own argument name)
        registry[id] = entry_point
        calls.append((id, entry_point, max_episode_steps))

    gym = types.ModuleType("gymnasium")
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register, reg.registry = register, registry
    gym.envs, envs.registration = envs, reg
    for name, mod in (("gymnasium", gym), ("gymnasium.envs", envs), ("gymnasium.envs.registration", reg)):
        monkeypatch.setitem(sys.modules, name, mod)
    from custom_gymnasium_environments_b200 import registration

    registration.register_all()  # (importing the package may already have done it: registration is idempotent)
    got = {i: (e, m) for i, e, m in calls}
    assert set(got) == set(registry) == {"snake_env_classic-v0", "CryptoTrading-v0", "TrafficManagement-v0",
                                         "SmartClimateEnv-v0"}
    assert got["snake_env_classic-v0"][1] == got["CryptoTrading-v0"][1] == got["TrafficManagement-v0"][1] == 1000
    for env_id, (entry, _) in got.items():  # every entry point resolves to a class of this package
        module, cls = entry.split(":")
        assert hasattr(__import__(module, fromlist=[cls]), cls), env_id
    registration.register_all()
    assert len(calls) == 4  # nothing is registered twice


def test_facades_are_gymnasium_envs_when_gymnasium_is_importable(tmp_path):
    """gym.make() refuses entry points that do not inherit gymnasium.Env: with a gymnasium on the path the single-instance
    facades must subclass its Env (and the batched classes its VectorEnv).  Checked in a subprocess with a stub package."""
    pkg = tmp_path / "gymnasium"
    (pkg / "vector").mkdir(parents=True)
    (pkg / "envs").mkdir()
    (pkg / "__init__.py").write_text("class Env:\n    metadata = {}\n    render_mode = None\n")
    (pkg / "vector" / "__init__.py").write_text("class VectorEnv:\n    metadata = {}\n    closed = False\n")
    (pkg / "envs" / "__init__.py").write_text("")
    (pkg / "envs" / "registration.py").write_text(
        "registry = {}\n"
        "def register(id, entry_point, max_episode_steps=None, **kw):\n"
        "    registry[id] = entry_point\n")
    code = ("import gymnasium, gymnasium.vector, custom_gymnasium_environments_b200 as p\n"
            "for c in (p.SnakeEnvClassic, p.CryptoTradingEnv, p.TrafficManagementEnv, p.SmartClimateEnv, p.WorldBuilderEnv):\n"
            "    assert issubclass(c, gymnasium.Env), c\n"
            "for c in (p.BatchedSnakeEnv, p.BatchedCryptoTradingEnv, p.BatchedTrafficManagementEnv):\n"
            "    assert issubclass(c, gymnasium.vector.VectorEnv), c\n"
            "from gymnasium.envs.registration import registry\n"
            "assert 'snake_env_classic-v0' in registry and 'CryptoTrading-v0' in registry, registry\n"
            "print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT,
                         env=dict(os.environ, PYTHONPATH=f"{tmp_path}{os.pathsep}{ROOT}"))
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr
