"""Environment ids.  Same id -> kwargs tables as the reference (rware/__init__.py:7-39): 4 sizes x 3 difficulties x
1..19 agents under `rware-{size}-{n}ag{diff}-v2`; the `-v1` spellings used by BASELINE.json are registered as aliases
with identical kwargs.  `make()` works without gymnasium; when gymnasium is importable the ids are
also registered there under the `rware_b200` namespace entry point.

`image_registration()` and `full_registration()` (rware/__init__.py:42-80, 83-172) enable the same extra id families as in
the reference.  The reference enumerates them eagerly (full_registration: 68 400 + 5 198 400 ids); here an id of an enabled
family is parsed on demand into the same kwargs."""
import itertools
import re

from .warehouse import ObservationType, RewardType, Warehouse

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


_enabled = set()
_OBS = {"": ObservationType.FLATTENED, "-img": ObservationType.IMAGE, "-imgdict": ObservationType.IMAGE_DICT}
_REW = {"indiv": RewardType.INDIVIDUAL, "global": RewardType.GLOBAL, "twostage": RewardType.TWO_STAGE}
_HEAD = r"^rware(?P<obs>-imgdict|-img|)(?P<nd>-Nd|)"
_RE_IMAGE = re.compile(_HEAD + r"-(?P<size>tiny|small|medium|large)-(?P<n>\d+)ag(?P<diff>-easy|-hard|)-v2$")
_RE_FULL_A = re.compile(_HEAD + r"(?P<s>-[2-5]s|)-(?P<size>tiny|small|medium|large)-(?P<h>\d+)h-(?P<n>\d+)ag(?P<diff>-easy|-hard|)-v2$")
_RE_FULL_B = re.compile(_HEAD + r"(?P<s>-[2-5]s|)-(?P<rows>[1-4])x(?P<cols>[3579])-(?P<h>\d+)h-(?P<n>\d+)ag-(?P<req>\d+)req-"
                        r"(?P<rew>indiv|global|twostage)-v2$")


def image_registration():
    """rware/__init__.py:42-80: `rware{-img|-imgdict}{-Nd}-{size}-{n}ag{diff}-v2`."""
    _enabled.add("image")


def full_registration():
    """rware/__init__.py:83-172: `rware{-img}{-Nd}{-Ks}-{size}-{h}h-{n}ag{diff}-v2` and
    `rware{-img}{-Nd}{-Ks}-{rows}x{cols}-{h}h-{n}ag-{req}req-{indiv|global|twostage}-v2`."""
    _enabled.add("full")


def _parse(id_):
    base = dict(msg_bits=0, max_inactivity_steps=None, max_steps=500)
    if "image" in _enabled:
        m = _RE_IMAGE.match(id_)
        # 55-60: (FLATTENED, directional) is the base family; "-Nd" only exists with an image observation type
        if m and m["obs"] and 1 <= int(m["n"]) <= 19:
            n = int(m["n"])
            return dict(base, column_height=8, shelf_rows=_sizes[m["size"]][0], shelf_columns=_sizes[m["size"]][1], n_agents=n,
                        sensor_range=1, request_queue_size=int(n * _difficulty[m["diff"]]), reward_type=RewardType.INDIVIDUAL,
                        observation_type=_OBS[m["obs"]], image_observation_directional=m["nd"] == "")
    if "full" in _enabled:
        m = _RE_FULL_A.match(id_)
        if m and m["obs"] != "-imgdict" and not (m["nd"] and not m["obs"]) and 1 <= int(m["n"]) <= 19 and 1 <= int(m["h"]) <= 15:
            n = int(m["n"])
            return dict(base, column_height=int(m["h"]), shelf_rows=_sizes[m["size"]][0], shelf_columns=_sizes[m["size"]][1],
                        n_agents=n, sensor_range=int(m["s"][1]) if m["s"] else 1,
                        request_queue_size=int(n * _difficulty[m["diff"]]), reward_type=RewardType.INDIVIDUAL,
                        observation_type=_OBS[m["obs"]], image_observation_directional=m["nd"] == "")
        m = _RE_FULL_B.match(id_)
        if m and m["obs"] != "-imgdict" and 1 <= int(m["n"]) <= 19 and 1 <= int(m["h"]) <= 15 and 1 <= int(m["req"]) <= 19:
            return dict(base, column_height=int(m["h"]), shelf_rows=int(m["rows"]), shelf_columns=int(m["cols"]),
                        n_agents=int(m["n"]), sensor_range=int(m["s"][1]) if m["s"] else 1, request_queue_size=int(m["req"]),
                        reward_type=_REW[m["rew"]], observation_type=_OBS[m["obs"]],
                        image_observation_directional=m["nd"] == "")
    return None


def env_kwargs(id_: str, **overrides):
    kwargs = registry.get(id_) or _parse(id_)  # `registry` holds the eagerly registered base family only
    if kwargs is None:
        raise KeyError(f"unknown environment id {id_!r} (image_registration() / full_registration() enable the extra families)")
    return {**kwargs, **overrides}


def make(id_: str, **overrides) -> Warehouse:
    """gymnasium.make equivalent for the registered RWARE ids (kwargs overrides as in README.md:106)."""
    return Warehouse(**env_kwargs(id_, **overrides))


def make_vec(id_: str, num_envs: int, **overrides):
    """Batched world set for a registered id."""
    from .vec import VecWarehouse
    return VecWarehouse(num_envs, **env_kwargs(id_, **overrides))


_register_all()
