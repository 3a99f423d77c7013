"""``utils`` package of the facade: reward.py / state.py here, everything else (utils.utils,
utils.plotting ...) falls through to the reference checkout when it is present."""
import os

_root = os.environ.get("IPP_REFERENCE_ROOT")  # set by facade.install(reference_root)
_ref = os.path.join(_root, "marl_framework", "utils") if _root else None
if _ref and os.path.isdir(_ref) and _ref not in __path__:
    __path__.append(_ref)
