"""Light ScenarIO-style shims (model / world / simulator) over the device state, so exploratory
scripts in the style of examples/ignition_interaction.py keep working. NOT on the hot path: every
call here is a host round trip through os2r_get_state / os2r_set_state.

Reference surface (complete list the reference touches):
  Model: joint_positions, joint_velocities, set_joint_generalized_force_targets,
         joint_generalized_force_targets, set_joint_control_mode, get_joint(n).set_joint_max_generalized_force,
         name, to_gazebo().reset_joint_positions / reset_joint_velocities
  World: model_names, get_model, to_gazebo().set_gravity / remove_model / insert_model
  Simulator: run(paused=), initialized, gui, close, step_size
"""
import numpy as np


class _JointShim:
    """ScenarIO ``Joint`` of one env: position / velocity access and reset (examples/ignition_interaction.py)."""

    def __init__(self, name, model=None):
        self._name, self._model = name, model

    def name(self):
        return self._name

    def to_gazebo(self):
        return self

    def set_joint_max_generalized_force(self, _values):
        return True   # the torque limit is the constant max_torque of the compiled model

    def joint_position(self):
        return self._model.joint_positions([self._name])

    def joint_velocity(self):
        return self._model.joint_velocities([self._name])

    def reset_position(self, value):
        return self._model.reset_joint_positions([float(value)], [self._name])

    def reset_velocity(self, value):
        return self._model.reset_joint_velocities([float(value)], [self._name])


class ModelShim:
    def __init__(self, runtime, env_index: int = 0):
        self._rt, self._e = runtime, env_index
        self._targets = {}

    def name(self):
        return 'monopod'

    def to_gazebo(self):
        return self

    def joint_names(self):
        return list(self._rt._compiled.joint_names)

    def _dofs(self, names):
        cm = self._rt._compiled
        return [cm.dof_of(n) for n in (names or cm.joint_names)]

    def joint_positions(self, names=None):
        st = self._rt.engine.get_state()[self._e]
        return [float(st[d]) for d in self._dofs(names)]

    def joint_velocities(self, names=None):
        n = self._rt._compiled.n_dof
        st = self._rt.engine.get_state()[self._e]
        return [float(st[n + d]) for d in self._dofs(names)]

    def reset_joint_positions(self, values, names=None):
        st = self._rt.engine.get_state()
        for d, v in zip(self._dofs(names), values):
            st[self._e, d] = float(v)
        self._rt.engine.set_state(st)
        return True

    def reset_joint_velocities(self, values, names=None):
        n = self._rt._compiled.n_dof
        st = self._rt.engine.get_state()
        for d, v in zip(self._dofs(names), values):
            st[self._e, n + d] = float(v)
        self._rt.engine.set_state(st)
        return True

    def set_joint_generalized_force_targets(self, data, names):
        self._targets = {n: float(x) for n, x in zip(names, data)}
        return True

    def joint_generalized_force_targets(self, names):
        return [self._targets.get(n, 0.0) for n in names]

    def links_in_contact(self):
        """ScenarIO ``Model.links_in_contact()``: names of the links whose collision proxies carry a normal impulse."""
        cm = self._rt._compiled
        n, nc = cm.n_dof, cm.struct.n_contacts
        st = self._rt.engine.get_state()[self._e]
        link_of = {'hip': 'upper_leg_link', 'knee': 'upper_leg_link', 'foot': 'lower_leg_link', 'hip_link': 'hip_link'}
        out = []
        for c in range(nc):
            link = link_of.get(cm.contact_names[c], cm.contact_names[c])
            if st[3 * n + 3 * c] > 0 and link not in out:
                out.append(link)
        return out

    def set_joint_control_mode(self, _mode, _names=None):
        return True

    def get_joint(self, name):
        if name not in self._rt._compiled.joint_names:
            raise KeyError(f'the model has no moving joint {name!r} (joints: {self._rt._compiled.joint_names})')
        return _JointShim(name, self)


class WorldShim:
    def __init__(self, runtime):
        self._rt = runtime

    def to_gazebo(self):
        return self

    def model_names(self):
        return [self._rt.task.model.name()]

    def get_model(self, _name=None):
        return self._rt.task.model

    def set_gravity(self, gravity):
        eng = self._rt.engine
        p = eng.get_params()
        p[:, -1] = float(gravity[2])
        eng.set_params(p)
        return True

    def gravity(self):
        return (0.0, 0.0, float(self._rt.engine.get_params()[0, -1]))

    def remove_model(self, _name):
        return True   # a reset re-creates the env state; there is no model object to remove

    def insert_model(self, *_a, **_k):
        return True


class SimulatorShim:
    def __init__(self, runtime):
        self._rt = runtime

    def step_size(self):
        return 1.0 / self._rt.physics_rate

    def initialized(self):
        return True

    def run(self, paused: bool = False):
        """One physics iteration in the reference; here physics only advances inside ``step``."""
        return True

    def gui(self):
        return True

    def close(self):
        return True
