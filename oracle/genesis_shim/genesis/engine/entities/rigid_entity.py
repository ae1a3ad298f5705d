"""TEST INFRASTRUCTURE ONLY -- inert rigid entity (plane / target sphere) for the genesis shim."""


class RigidEntity:
    def __init__(self, scene):
        self.scene = scene

    def _build(self, n_envs, dt):
        pass

    def _step(self):
        pass

    def set_pos(self, pos, zero_velocity=True, envs_idx=None):
        pass
