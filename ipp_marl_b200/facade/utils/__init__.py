"""``utils`` package of the facade: reward.py / state.py here, everything else (utils.utils,
utils.plotting ...) falls through to the reference checkout when it is present."""
import os

_ref = os.path.join(os.environ.get("IPP_REFERENCE_ROOT", "/root/reference"), "marl_framework", "utils")
if os.path.isdir(_ref) and _ref not in __path__:
    __path__.append(_ref)
