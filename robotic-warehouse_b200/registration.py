"""Environment ids.  Same id -> kwargs tables as the reference (rware/__init__.py:7-39): 4 sizes x 3 difficulties x
1..19 agents under `rware-{size}-{n}ag{diff}-v2`; the `-v1` spellings used by BASELINE.json are registered as aliases
with identical kwargs.  `make()` works without gymnasium; when gymnasium is importable the ids are
also registered there under the `rware_b200` namespace entry point."""
import itertools

from .warehouse import RewardType, Warehouse

_sizes = {"tiny": (1, 3), "small": (2, 3), "medium": (2, 5), "large": (3, 5)}  # rware/__init__.py:7-12
_difficulty = {"-easy": 2, "": 1, "-hard": 0.5}                                 # rware/__init__.py:14

registry = {}


def _register_all():
    for size, diff, agents in itertools.product(_sizes.keys(), _difficulty, range(1, 20)):
        kwargs = {
            "column_height": 8,
            "shelf_rows": _sizes[size][0],
            "shelf_columns": _sizes[size][1],
            "n_agents": agents,
            "msg_bits": 0,
            "sensor_range": 1,
            "request_queue_size": int(agents * _difficulty[diff]),
            "max_inactivity_steps": None,
            "max_steps": 500,
            "reward_type": RewardType.INDIVIDUAL,
        }
        for version in ("v2", "v1"):
            registry[f"rware-{size}-{agents}ag{diff}-{version}"] = kwargs
    try:  # optional: expose the ids to gymnasium.make as well
        import gymnasium
        for id_, kwargs in registry.items():
            if id_ not in gymnasium.registry:
                gymnasium.register(id=id_, entry_point="rware_b200.warehouse:Warehouse", kwargs=kwargs)
    except Exception:
        pass


def env_kwargs(id_: str, **overrides):
    if id_ not in registry:
        raise KeyError(f"unknown environment id {id_!r}")
    return {**registry[id_], **overrides}


def make(id_: str, **overrides) -> Warehouse:
    """gymnasium.make equivalent for the registered RWARE ids (kwargs overrides as in README.md:106)."""
    return Warehouse(**env_kwargs(id_, **overrides))


def make_vec(id_: str, num_envs: int, **overrides):
    """Batched world set for a registered id."""
    from .vec import VecWarehouse
    return VecWarehouse(num_envs, **env_kwargs(id_, **overrides))


_register_all()
