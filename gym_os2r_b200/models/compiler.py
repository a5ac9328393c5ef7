"""Host model compiler: URDF -> constant kinematic-tree / inertia tables (struct os2r_model).

Replaces what the reference gets from ``world.insert_model(urdf)`` (gym_os2r/models/monopod.py:27)
+ sdformat's URDF->SDF conversion + DART's skeleton construction. Run once per task mode on the
host; the resulting POD struct is handed to the CUDA library.

Rules (restating the upstream behaviour the reference relies on):
  * only the chain world -> ... -> tip of *moving* (continuous/revolute) joints becomes bodies;
  * a link attached by a ``fixed`` joint is lumped into the nearest moving ancestor (mass, COM and
    inertia combined about the new COM; collision proxies re-expressed in the ancestor's frame) —
    this is what sdformat does, so a mass randomisation scales the lumped mass as one link;
  * links welded to ``world`` are static and dropped (their collisions too);
  * rpy are fixed-axis roll/pitch/yaw, R = Rz(y) Ry(p) Rx(r), taken literally (1.57 is not pi/2);
  * an ``<inertial><origin rpy>`` rotates the inertia tensor into the link frame.
"""
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .. import _capi
from . import assets


def rpy_to_matrix(rpy) -> np.ndarray:
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def _vec(text: Optional[str], n=3) -> np.ndarray:
    if text is None:
        return np.zeros(n)
    v = np.array([float(t) for t in text.split()])
    assert v.shape == (n,), text
    return v


@dataclass
class _Link:
    name: str
    mass: float = 0.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))
    inertia: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))  # about COM, link axes
    spheres: List[tuple] = field(default_factory=list)  # (name, centre[3], radius)


@dataclass
class _Joint:
    name: str
    type: str
    parent: str
    child: str
    R: np.ndarray
    p: np.ndarray
    axis: np.ndarray
    damping: float
    friction: float


@dataclass
class CompiledModel:
    """Python-side view of the tables (numpy) plus the ctypes struct handed to the C-ABI."""
    name: str
    joint_names: List[str]            # moving joints, chain order root -> tip
    body_names: List[str]             # lumped link names per body
    contact_names: List[str]
    struct: _capi.Model

    @property
    def n_dof(self) -> int:
        return self.struct.n_dof

    def dof_of(self, joint_name: str) -> int:
        return self.joint_names.index(joint_name)


def parse_urdf(path: str):
    root = ET.parse(path).getroot()
    links: Dict[str, _Link] = {}
    for le in root.findall('link'):
        lk = _Link(le.get('name'))
        ine = le.find('inertial')
        if ine is not None:
            org = ine.find('origin')
            xyz = _vec(org.get('xyz') if org is not None else None)
            R = rpy_to_matrix(_vec(org.get('rpy') if org is not None else None))
            lk.mass = float(ine.find('mass').get('value'))
            it = ine.find('inertia')
            g = lambda k: float(it.get(k, 0.0))
            I = np.array([[g('ixx'), g('ixy'), g('ixz')],
                          [g('ixy'), g('iyy'), g('iyz')],
                          [g('ixz'), g('iyz'), g('izz')]])
            lk.com = xyz
            lk.inertia = R @ I @ R.T
        for ce in le.findall('collision'):
            geom = ce.find('geometry')
            sph = geom.find('sphere') if geom is not None else None
            if sph is None:
                raise ValueError(f'{path}: link {lk.name}: only <sphere> collision proxies are supported')
            org = ce.find('origin')
            lk.spheres.append((ce.get('name', lk.name), _vec(org.get('xyz') if org is not None else None),
                               float(sph.get('radius'))))
        links[lk.name] = lk
    joints: List[_Joint] = []
    for je in root.findall('joint'):
        org = je.find('origin')
        ax = je.find('axis')
        dyn = je.find('dynamics')
        joints.append(_Joint(
            name=je.get('name'), type=je.get('type'),
            parent=je.find('parent').get('link'), child=je.find('child').get('link'),
            R=rpy_to_matrix(_vec(org.get('rpy') if org is not None else None)),
            p=_vec(org.get('xyz') if org is not None else None),
            axis=_vec(ax.get('xyz')) if ax is not None else np.array([1.0, 0, 0]),
            damping=float(dyn.get('damping', 0.0)) if dyn is not None else 0.0,
            friction=float(dyn.get('friction', 0.0)) if dyn is not None else 0.0))
    return root.get('name'), links, joints


def _axis_index(axis: np.ndarray, jname: str) -> int:
    for k in range(3):
        e = np.zeros(3)
        e[k] = 1.0
        if np.allclose(axis, e, atol=1e-12):
            return k
    raise ValueError(f'joint {jname}: axis {axis} must be +x, +y or +z')


def compile_urdf(path: str, physics: dict, max_torque=(2.5, 2.5)) -> CompiledModel:
    name, links, joints = parse_urdf(path)
    by_parent: Dict[str, List[_Joint]] = {}
    for j in joints:
        by_parent.setdefault(j.parent, []).append(j)

    bodies = []  # dicts: joint, R, p (in parent body/world frame), members [(link, R_in_body, p_in_body)]

    def walk(link_name, cur_body, R_acc, p_acc):
        """R_acc, p_acc: pose of `link_name`'s frame in the current body's frame (or world)."""
        for j in by_parent.get(link_name, []):
            Rj = R_acc @ j.R
            pj = p_acc + R_acc @ j.p
            if j.type in ('continuous', 'revolute'):
                body = dict(joint=j, R=Rj, p=pj, parent=cur_body, members=[(links[j.child], np.eye(3), np.zeros(3))])
                bodies.append(body)
                walk(j.child, len(bodies) - 1, np.eye(3), np.zeros(3))
            elif j.type == 'fixed':
                if cur_body is not None:
                    bodies[cur_body]['members'].append((links[j.child], Rj, pj))
                walk(j.child, cur_body, Rj, pj)
            else:
                raise ValueError(f'joint {j.name}: unsupported type {j.type}')

    if 'world' not in links:
        raise ValueError('URDF must root the chain at a link named "world"')
    walk('world', None, np.eye(3), np.zeros(3))
    n = len(bodies)
    if not 1 <= n <= _capi.MAX_DOF:
        raise ValueError(f'{n} moving joints; supported 1..{_capi.MAX_DOF}')
    for i, b in enumerate(bodies):
        want = None if i == 0 else i - 1
        if b['parent'] != want:
            raise ValueError('only serial chains are supported')

    m = _capi.Model()
    m.n_dof = n
    m.substeps = int(physics.get('substeps', 10))
    m.pgs_iters = int(physics['pgs_iters'])
    m.pgs_tol = float(physics.get('pgs_tol', 0.0))
    m.pgs_joint_sweeps = int(physics.get('pgs_joint_sweeps', 1))
    m.gravity_z = float(physics['gravity_z'])
    m.dt = float(physics['dt'])
    m.erp = float(physics['erp'])
    m.max_erv = float(physics['max_erv'])
    m.cfm_contact = float(physics['cfm_contact'])
    m.cfm_joint = float(physics['cfm_joint'])
    m.max_torque[0], m.max_torque[1] = float(max_torque[0]), float(max_torque[1])
    for r in range(_capi.N_ROLES):
        m.role_dof[r] = -1
    mu_eff = min(float(physics['link_mu']), float(physics['ground_mu']))

    joint_names, body_names, contact_names = [], [], []
    nc = 0
    for i, b in enumerate(bodies):
        j = b['joint']
        joint_names.append(j.name)
        if j.name in _capi.ROLE_OF_JOINT:
            m.role_dof[_capi.ROLE_OF_JOINT[j.name]] = i
        m.axis[i] = _axis_index(j.axis, j.name)
        for k in range(9):
            m.tree_R[i][k] = float(b['R'].reshape(-1)[k])
        for k in range(3):
            m.tree_p[i][k] = float(b['p'][k])
        m.damping[i] = j.damping
        m.friction[i] = j.friction
        # lump members
        mass = sum(lk.mass for lk, _, _ in b['members'])
        if mass <= 0:
            raise ValueError(f'body of joint {j.name} has no mass')
        com = sum(lk.mass * (p + R @ lk.com) for lk, R, p in b['members']) / mass
        I = np.zeros((3, 3))
        for lk, R, p in b['members']:
            d = (p + R @ lk.com) - com
            I += R @ lk.inertia @ R.T + lk.mass * (d @ d * np.eye(3) - np.outer(d, d))
        m.mass[i] = mass
        for k in range(3):
            m.com[i][k] = float(com[k])
        for k, (a, c) in enumerate(((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))):
            m.inertia[i][k] = float(I[a, c])
        body_names.append('+'.join(lk.name for lk, _, _ in b['members']))
        for lk, R, p in b['members']:
            for sname, centre, radius in lk.spheres:
                if nc >= _capi.MAX_CONTACTS:
                    raise ValueError('too many collision proxies')
                m.contact_body[nc] = i
                c = p + R @ centre
                for k in range(3):
                    m.contact_pos[nc][k] = float(c[k])
                m.contact_radius[nc] = radius
                m.contact_mu[nc] = mu_eff
                contact_names.append(sname)
                nc += 1
    m.n_contacts = nc
    return CompiledModel(name=name, joint_names=joint_names, body_names=body_names,
                         contact_names=contact_names, struct=m)


def compile_model(model_name: str, physics: dict, max_torque=(2.5, 2.5)) -> CompiledModel:
    """Compile one of the shipped models by name ('monopod', 'monopod-fixed_hip', ...)."""
    return compile_urdf(assets.get_model_file(model_name), physics, max_torque)


# ---------------------------------------------------------------------------------------------
# small numpy forward kinematics over the tables (host-side checks, reset clearance, tests)
# ---------------------------------------------------------------------------------------------

def _rot_axis(axis: int, q: float) -> np.ndarray:
    c, s = math.cos(q), math.sin(q)
    if axis == 0:
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    if axis == 1:
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def forward_kinematics(model: _capi.Model, q):
    """World rotation / origin of every body frame and world centres of the contact spheres."""
    R, p = np.eye(3), np.zeros(3)
    Rs, ps = [], []
    for i in range(model.n_dof):
        Rt = np.array(model.tree_R[i][:]).reshape(3, 3)
        pt = np.array(model.tree_p[i][:])
        p = p + R @ pt
        R = R @ Rt @ _rot_axis(model.axis[i], float(q[i]))
        Rs.append(R)
        ps.append(p)
    cs = []
    for c in range(model.n_contacts):
        b = model.contact_body[c]
        cs.append(ps[b] + Rs[b] @ np.array(model.contact_pos[c][:]))
    return Rs, ps, cs
