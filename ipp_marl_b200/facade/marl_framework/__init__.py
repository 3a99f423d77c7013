"""``marl_framework.<pkg>`` twins of the facade packages + fall-through to the reference tree.

The reference imports both ``mapping.x`` and ``marl_framework.mapping.x``; here the two spellings
are the SAME module object (no dual-identity trap), and sub-modules the facade
does not provide (actor, critic, coma_wrapper, params ...) resolve to the reference checkout
(IPP_REFERENCE_ROOT) when it is present.
"""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_FACADE_PKGS = ("mapping", "sensors", "agent")
_FACADE_UTILS = ("utils.reward", "utils.state")

_root = os.environ.get("IPP_REFERENCE_ROOT")  # set by facade.install(reference_root)
_ref = os.path.join(_root, "marl_framework") if _root else None
if _ref and os.path.isdir(_ref) and _ref not in __path__:
    __path__.append(_ref)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("marl_framework."):
            return None
        bare = fullname[len("marl_framework."):]
        top = bare.split(".")[0]
        if top in _FACADE_PKGS or bare in _FACADE_UTILS:
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        bare = spec.name[len("marl_framework."):]
        return importlib.import_module(bare)

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
