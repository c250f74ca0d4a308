"""The two registries rl_games keeps for vectorised environments (rl_games 1.1.4 common/vecenv.py `vecenv_config` and
common/env_configurations.py `configurations`), restated so that `AMPAgent(base_name, config)` can resolve `config['env_name']`
the way A2CBase does when rl_games itself is not installed (tests; standalone use).  run_hydra.py:238-240 fills rl_games' own:

    vecenv.register('RLGPU', lambda config_name, num_actors, **kwargs: RLGPUEnv(config_name, num_actors, **kwargs))
    env_configurations.register('rlgpu', {'env_creator': lambda **kwargs: create_rlgpu_env(**kwargs), 'vecenv_type': 'RLGPU'})
"""
from typing import Callable, Dict

vecenv_config: Dict[str, Callable] = {}
configurations: Dict[str, dict] = {}


def register_vecenv(config_name: str, func: Callable) -> None:
    vecenv_config[config_name] = func


def register(name: str, config: dict) -> None:
    configurations[name] = config


def create_vec_env(config_name: str, num_actors: int, **kwargs):
    vec_env_name = configurations[config_name]["vecenv_type"]
    return vecenv_config[vec_env_name](config_name, num_actors, **kwargs)
