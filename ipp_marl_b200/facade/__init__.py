"""Drop-in facade: the reference's module / class / function names for the environment path, backed
by the CUDA library through the C ABI (include/ipp_b200.h).

``install(reference_root)`` puts this directory at the front of ``sys.path`` so that
``mapping.*``, ``sensors.*``, ``agent.*``, ``utils.reward``, ``utils.state`` (and their
``marl_framework.``-prefixed twins — both spellings are imported by the reference, e.g.
coma_wrapper.py:7-17) resolve to the modules here, while everything else of the reference
(``coma_wrapper``, ``IG_baseline``, ``lawn_mower``, ``actor``, ``critic``, ``batch_memory``,
``utils.utils`` ...) keeps coming from the reference tree unchanged.  See INTEGRATION.md.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def install(reference_root=None):
    """Activate the facade.  ``reference_root``: checkout of dmar-bonn/ipp-marl (optional)."""
    if reference_root is not None:
        os.environ["IPP_REFERENCE_ROOT"] = reference_root
        fw = os.path.join(reference_root, "marl_framework")
        for p in (reference_root, fw):
            if p not in sys.path:
                sys.path.append(p)  # behind the facade directory
    if HERE in sys.path:
        sys.path.remove(HERE)
    sys.path.insert(0, HERE)
    return HERE
