"""utils/state.py facade (reference: utils/state.py:14-121): Shannon entropy + weighted entropy maps."""
import numpy as np

from ipp_marl_b200.facade import _runtime as R

_params = {"value": None}


def bind_params(params):
    """get_shannon_entropy has no params argument in the reference; the facade needs a handle to
    run it on the GPU.  Mapping/Agent construction binds the active params automatically."""
    _params["value"] = params


def _rt():
    if _params["value"] is None:
        raise RuntimeError("facade not bound: construct a Mapping (or call utils.state.bind_params) first")
    return R.runtime(_params["value"])


def get_shannon_entropy(p):
    """state.py:118-121 — clamps its argument IN PLACE, returns H(p) in bits in the dtype of p."""
    rt = _rt()
    pc, is64 = R.native(p)
    out = np.empty_like(pc)
    rc = rt.lib.ipp_shannon_entropy(rt.h, R.ptr(pc), is64, pc.size, R.ptr(out))
    rt.check(rc, "ipp_shannon_entropy")
    if isinstance(p, np.ndarray):
        p[...] = pc
    return out


def _weights(target):
    target = target.copy()
    target[target > 0.501] = 1
    target[target < 0.499] = 0
    w = target.copy()
    w[np.round(w, 2) == 0] = 0
    w[np.round(w, 2) == 1] = 1
    w[np.round(w, 2) == 0.5] = 0.5
    return w


def calculate_w_entropy(grid_map, map_footprint, simulated_map, observability, agent_state_space):
    """state.py:53-115 (class_weighting hard-coded to [0, 1] as in the reference, :60)."""
    target = simulated_map if observability == "eval" else grid_map
    weightings = _weights(np.array(target, dtype=np.float64))
    se = get_shannon_entropy(grid_map)
    w_entropy_map = weightings * se
    w_entropy_map_footprint = None
    if observability == "actor":
        wf = _weights(np.array(map_footprint, dtype=np.float64))
        w_entropy_map_footprint = wf * get_shannon_entropy(map_footprint)
    return w_entropy_map, weightings, se, w_entropy_map_footprint, grid_map


def get_w_entropy_map(map_footprint, local_map, simulated_map, observability, agent_state_space):
    """state.py:14-50.  "reward"/"eval": full resolution; other modes first area-downsample to the
    agent lattice with cv2.INTER_AREA exactly like the reference (observation features)."""
    if observability not in ("reward", "eval"):
        import cv2

        size = (int(agent_state_space.space_dim[1]), int(agent_state_space.space_dim[0]))
        grid_map = cv2.resize(local_map, size, interpolation=cv2.INTER_AREA)
        if observability == "actor":
            map_footprint = cv2.resize(map_footprint, size, interpolation=cv2.INTER_AREA)
        simulated_map = cv2.resize(simulated_map, size, interpolation=cv2.INTER_AREA)
    else:
        grid_map = local_map.copy()
    return calculate_w_entropy(grid_map, map_footprint, simulated_map, observability, agent_state_space)
