"""Config loader with the reference's behaviour (utils/flags.py:9-45): a YAML
file or an already-loaded dict becomes a tree of namedtuples reachable through
``Flags(x).get()``; string leaves that evaluate as Python literals are
converted (the reference uses ``eval``; ``ast.literal_eval`` is the safe
equivalent for the literals the yaml files contain, e.g. "5e-4")."""
import ast
import collections
import os

import yaml


def _to_namedtuple(d):
    out = {}
    for k, v in d.items():
        if k == "prefix":
            v = os.path.join("./", v)
        if isinstance(v, dict):
            v = _to_namedtuple(v)
        elif isinstance(v, str):
            try:
                v = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                pass
        out[k] = v
    return collections.namedtuple("FLAGS", sorted(out.keys()))(**out)


class Flags:
    def __init__(self, config_file):
        if isinstance(config_file, dict):
            d = config_file
        else:
            with open(config_file, "r") as f:
                d = yaml.safe_load(f)
        self.flags = _to_namedtuple(d)

    def get(self):
        return self.flags
