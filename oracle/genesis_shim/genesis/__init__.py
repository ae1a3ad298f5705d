"""TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the `genesis` package.

It exists so that the reference's own `envs/hover_env.py` can be imported and
run UNCHANGED in the build container (PYTHONPATH=oracle/genesis_shim:/root/reference/src)
to pin the env-logic oracle (`oracle/env_oracle.py`) and to generate the golden
fixtures under `tests/golden/` (`oracle/make_golden.py`).  It is never imported
by the product package, by `-m gpu` tests, by `smoke()` or by `bench.py`.

Only the symbols the reference touches are provided (call sites:
`hover_env.py:40-49, 178-242, 437-439, 452-470, 566-574, 721-735`,
`tests/conftest.py:16-21`).  Physics = `oracle/drone_physics.py` (restated,
parity unpinned -- see that file's header).
"""

from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import drone_physics as _phys  # noqa: E402

tc_float = torch.float32
tc_int = torch.int32
cpu = "cpu"
gpu = "gpu"
EPS = 1e-15


def init(backend=None, logging_level=None, **kwargs):
    return None


class _Opt(SimpleNamespace):
    def __init__(self, **kw):
        super().__init__(**kw)


options = SimpleNamespace(SimOptions=_Opt, ViewerOptions=_Opt, VisOptions=_Opt, RigidOptions=_Opt)
constraint_solver = SimpleNamespace(Newton="Newton", CG="CG")
surfaces = SimpleNamespace(Rough=_Opt, Default=_Opt)
textures = SimpleNamespace(ColorTexture=_Opt)


class _Morph(SimpleNamespace):
    kind = "generic"


class _Plane(_Morph):
    kind = "plane"


class _Mesh(_Morph):
    kind = "mesh"


class _Drone(_Morph):
    kind = "drone"


morphs = SimpleNamespace(Plane=_Plane, Mesh=_Mesh, Drone=_Drone)

from .engine.entities.drone_entity import DroneEntity  # noqa: E402
from .engine.entities.rigid_entity import RigidEntity  # noqa: E402


class Scene:
    def __init__(self, sim_options=None, viewer_options=None, vis_options=None, rigid_options=None, show_viewer=False, **kw):
        self.dt = float(sim_options.dt)
        self.entities = []
        self.n_envs = 0
        self.envs_offset = None

    def add_entity(self, morph=None, surface=None, **kw):
        if isinstance(morph, _Drone):
            ent = DroneEntity(self)
        else:
            ent = RigidEntity(self)
        self.entities.append(ent)
        return ent

    def add_camera(self, **kw):
        raise NotImplementedError("genesis shim: no rendering")

    def build(self, n_envs=1, env_spacing=(0.0, 0.0), **kw):
        self.n_envs = n_envs
        self.envs_offset = np.zeros((n_envs, 3))
        for e in self.entities:
            e._build(n_envs, self.dt)

    def step(self):
        for e in self.entities:
            e._step()

    def clear_debug_objects(self):
        pass

    def draw_debug_sphere(self, **kw):
        pass
