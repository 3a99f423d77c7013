"""CPU: the C-ABI library builds, loads and exports every symbol include/ipp_b200.h declares; the
ctypes mirror of ipp_config has the C layout; no-GPU error behaviour; multi-process sharding logic."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    with open(os.path.join(ROOT, "include", "ipp_b200.h")) as f:
        return f.read()


def test_library_exports_every_declared_symbol():
    from ipp_marl_b200 import _native

    lib = _native.load()
    declared = set(re.findall(r"^\s*(?:const char\*|int64_t|int)\s+(ipp_\w+)\s*\(", _header(), re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ipp_version() >= 100
    assert lib.ipp_status_string(-4) == b"no CUDA device"


def test_config_struct_layout_matches_c(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    from ipp_marl_b200 import _native

    fields = [f[0] for f in _native.IppConfig._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ipp_b200.h"', "int main(void){",
             'printf("%zu\\n", sizeof(ipp_config));']
    for f in fields:
        lines.append('printf("%%zu\\n", offsetof(ipp_config, %s));' % f)
    lines += ['printf("%zu %zu\\n", sizeof(ipp_state), sizeof(ipp_step_io));', "return 0;}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == C.sizeof(_native.IppConfig)
    for f, off in zip(fields, out[1:]):
        assert getattr(_native.IppConfig, f).offset == int(off), f
    assert int(out[-2]) == C.sizeof(_native.IppState) and int(out[-1]) == C.sizeof(_native.IppStepIO)


def test_create_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ipp_marl_b200 import BatchedIPPEnv, _native
    from tests.helpers import load_kats

    with pytest.raises(_native.IppError):
        BatchedIPPEnv(load_kats()["synthetic50"]["params"], 4)


def test_invalid_configs_are_rejected():
    from ipp_marl_b200 import HostTables, _native, make_config
    from tests.helpers import load_kats

    lib = _native.load()
    tables = HostTables(load_kats()["synthetic50"]["params"])
    h = C.c_void_p()
    for field, value in (("map_stride", 2501), ("gt_stride", 2500), ("n_agents", 9), ("gy", 3), ("prior", 1.5)):
        cfg = make_config(tables, 8)
        setattr(cfg, field, value)
        rc = lib.ipp_create(C.byref(cfg), C.byref(h))
        assert rc in (-1, -2), (field, rc)
    assert lib.ipp_create(None, C.byref(h)) == -1
    # entry points reject a missing handle / io before they touch the device
    st, io = _native.IppState(), _native.IppStepIO()
    assert lib.ipp_step(None, C.byref(st), 0, C.byref(io), None) == -1
    assert lib.ipp_run_steps(None, C.byref(st), 1, None, 0, 15, C.byref(io), None) == -1
    assert lib.ipp_reset(None, C.byref(st), None, None) == -1
    assert lib.ipp_status_string(-1) == b"invalid argument" and lib.ipp_version() >= 210
    params = load_kats()["synthetic50"]["params"]
    params["experiment"]["constraints"]["num_actions"] = 9
    with pytest.raises(ValueError):
        HostTables(params)


def test_sharding_is_partition_invariant_gloo(tmp_path):
    """world_size-2 gloo run of the repo's host-side sharding logic: `env.default_episode_ids` (rank r owns global
    envs [r*B, (r+1)*B), episode = global index + 1), `mission.next_episode_ids` (successive rollouts of a sharded
    training job never reuse an episode number) and the job metric of bench.py (sum over ranks / max-over-ranks
    time).  No data-path collective exists on the env path (envs are independent)."""
    script = tmp_path / "w.py"
    script.write_text(r"""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, %r)
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
from ipp_marl_b200.env import default_episode_ids
from ipp_marl_b200.mission import next_episode_ids
B = 6
ep = default_episode_ids(B, rank * B)                              # BatchedIPPEnv.reset default with env_id_base=rank*B
gathered = [torch.zeros_like(ep) for _ in range(world)]
dist.all_gather(gathered, ep)
nxt = rank * B + 1                                                 # COMAMission._next_episode at construction
seen = []
for rollout in range(3):
    ids, nxt = next_episode_ids(B, nxt, world)
    allr = [torch.zeros_like(ids) for _ in range(world)]
    dist.all_gather(allr, ids)
    seen += torch.cat(allr).tolist()
ms = torch.tensor([10.0 + rank], dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"episodes": torch.cat(gathered).tolist(), "ms": ms.item(), "value": world * B / (ms.item() * 1e-3),
                      "rollouts": sorted(seen)}))
dist.destroy_process_group()
""" % ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    import json

    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["episodes"] == list(range(1, 13))  # ranks tile the global episode range without overlap
    assert out["ms"] == 11.0 and abs(out["value"] - 12 / 0.011) < 1e-6
    assert out["rollouts"] == list(range(1, 37))  # 3 rollouts x 2 ranks x 6 envs: every episode number exactly once


def test_prior_other_than_half_warns():
    """DESIGN.md section 2, deviation (1): rewards agree with the reference only to 2e-2 when mapping.prior != 0.5 —
    HostTables says so loudly."""
    import copy

    from ipp_marl_b200.geometry import HostTables
    from tests.helpers import load_kats

    params = copy.deepcopy(load_kats()["synthetic50"]["params"])
    params["mapping"]["prior"] = 0.4
    with pytest.warns(RuntimeWarning, match="prior"):
        HostTables(params)
