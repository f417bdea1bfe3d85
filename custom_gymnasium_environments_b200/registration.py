"""Gymnasium registration of the single-instance façades under the reference's ids
(snake_env_classic/__init__.py:3-7: id 'snake_env_classic-v0', max_episode_steps=1000).
A no-op when gymnasium is not installed (it is not, in this image)."""
from __future__ import annotations

REGISTERED = []


def register_all():
    try:
        from gymnasium.envs.registration import register, registry  # type: ignore
    except Exception:
        return REGISTERED
    specs = [
        # snake_env_classic/__init__.py:3-7
        ("snake_env_classic-v0", "custom_gymnasium_environments_b200.snake:SnakeEnvClassic", 1000),
        # traffic_management_env/__init__.py:7-17
        ("TrafficManagement-v0", "custom_gymnasium_environments_b200.traffic:TrafficManagementEnv", 1000),
        # smartclimate_rl-main/smartclimate/__init__.py:6-10
        ("SmartClimateEnv-v0", "custom_gymnasium_environments_b200.climate:SmartClimateEnv", 1440),
        # crypto_trading_env/crypto_trading_env.py:739-743
        ("CryptoTrading-v0", "custom_gymnasium_environments_b200.crypto:CryptoTradingEnv", 1000),
    ]
    for env_id, entry, max_steps in specs:
        if env_id not in registry:
            register(id=env_id, entry_point=entry, max_episode_steps=max_steps)
            REGISTERED.append(env_id)
    return REGISTERED
