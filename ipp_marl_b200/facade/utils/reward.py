"""utils/reward.py facade (reference: utils/reward.py:11-53,68-82): information-gain reward."""
import numpy as np

from ipp_marl_b200.facade import _runtime as R
from utils import state as _state


def get_utility_reward(state, state_, simulated_map, agent_state_space):
    """reward.py:68-82 -> (absolute, relative) utility of going from map `state` to `state_`."""
    rt = _state._rt()
    (a, a64), (b, b64) = R.native(state), R.native(state_)
    out = np.zeros(2, dtype=np.float64)
    rc = rt.lib.ipp_utility_reward(rt.h, R.ptr(a), a64, R.ptr(b), b64, a.size, R.ptr(out))
    rt.check(rc, "ipp_utility_reward")
    return float(out[0]), float(out[1])


def get_global_reward(last_map, next_map, mission_type, footprints, simulated_map, agent_state_space, actions,
                      agent_id, t, budget):
    """reward.py:11-53 -> (done=False, 22*rel-0.5, 10*abs-0.17)."""
    absolute, relative = get_utility_reward(last_map, next_map, simulated_map, agent_state_space)
    return False, 22 * relative - 0.5, 10 * absolute - 0.17


def is_collided(p1, p2):
    return bool(np.array_equal(p1, p2))
