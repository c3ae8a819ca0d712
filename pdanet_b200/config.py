"""YAML model configs as attribute dictionaries (the reference uses EasyDict, pcdet/config.py:71-80;
easydict is not installed here and the reference's data/optimisation sections are out of scope)."""
from __future__ import annotations

from pathlib import Path

import yaml

CFG_DIR = Path(__file__).resolve().parent / "cfgs"


class AttrDict(dict):
    """dict with attribute access and .get(), recursively, like EasyDict."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, list):
            return [AttrDict._wrap(x) for x in v]
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__


def load_config(name_or_path: str) -> AttrDict:
    """'kitti' / 'once' or a path to a YAML file with the same schema."""
    alias = {"kitti": CFG_DIR / "kitti_pda_ssd.yaml", "once": CFG_DIR / "once_pda_ssd.yaml"}
    path = alias.get(str(name_or_path), Path(name_or_path))
    with open(path, "r") as f:
        return AttrDict(yaml.safe_load(f))
