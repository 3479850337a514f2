"""id -> entry point registry with gym 0.18's ``make(id, **overrides)`` kwargs merge (train.py:87-103)."""
import importlib
import inspect


class EnvSpec:
    def __init__(self, id, entry_point=None, reward_threshold=None, max_episode_steps=None, kwargs=None, **_):
        self.id, self.entry_point, self.reward_threshold = id, entry_point, reward_threshold
        self.max_episode_steps, self._kwargs = max_episode_steps, dict(kwargs or {})

    def make(self, **overrides):
        kw = dict(self._kwargs)
        kw.update(overrides)
        ep = self.entry_point
        if isinstance(ep, str):
            mod, _, name = ep.partition(":")
            ep = getattr(importlib.import_module(mod), name)
        if "noise_scale" in kw and "noise_scale" not in inspect.signature(ep.__init__).parameters:
            kw.pop("noise_scale")   # train.py passes noise_scale=0 to every env family; the pH envs take no such argument
        env = ep(**kw)
        env.spec = self
        return env


class EnvRegistry:
    def __init__(self):
        self.env_specs = {}

    def register(self, id, **kwargs):
        self.env_specs[id] = EnvSpec(id, **kwargs)

    def spec(self, id):
        if id not in self.env_specs:
            raise KeyError(f"No registered env with id: {id}")
        return self.env_specs[id]

    def make(self, id, **kwargs):
        return self.spec(id).make(**kwargs)

    def all(self):
        return self.env_specs.values()


registry = EnvRegistry()


def register(id, **kwargs):
    return registry.register(id, **kwargs)


def make(id, **kwargs):
    return registry.make(id, **kwargs)


def spec(id):
    return registry.spec(id)
