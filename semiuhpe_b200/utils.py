"""Drop-in for the Euler helpers of the reference's ``src/utils.py`` (lines 202-300)."""
import math

import numpy as np

from . import _ops


def compute_euler_angles_from_rotation_matrices(rotation_matrices, full_range=False, use_gpu=True, gpu_id=0):
    """(b,3,3) -> (b,3) radians (pitch, yaw, roll)  -- src/utils.py:232-260.
    One K4 launch instead of the reference's per-sample Python loop (:240-242).
    ``use_gpu``/``gpu_id`` are accepted for signature compatibility; the result
    lives on the input's CUDA device."""
    R = rotation_matrices
    if R.dim() == 3 and R.shape[1:] == (4, 4):
        R = R[:, :3, :3]
    return _ops.so3_metrics(R, full_range=full_range, euler=True)["euler"]


def euler_dad_degrees(rotation_matrices):
    """(b,3,3) -> (b,3) DEGREES (pitch, yaw, roll) in the convention of models trained on
    DAD-3DHeads: per sample ``Rotation.from_matrix(R.T).as_euler("xyz", degrees=True)`` and
    ``[roll, pitch, yaw] = limit_angle([a2, a0 - 180, a1])`` -- eval.py:66-74, predict.py:84-87,
    image.py:218-221 (a scipy call per sample on the host in the reference; one K4 launch here)."""
    return _ops.so3_metrics(rotation_matrices, full_range="dad", euler=True)["euler"]


def get_6DRepNet_Rot(x, y, z):
    """R = Rz(z) Ry(y) Rx(x) from radians (host helper used to build labels, src/utils.py:204-226)."""
    cx, sx, cy, sy, cz, sz = math.cos(x), math.sin(x), math.cos(y), math.sin(y), math.cos(z), math.sin(z)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz.dot(Ry.dot(Rx))


def rot_euler_6DRepNet(rotation_matrices, full_range=False):
    """One 3x3 matrix -> numpy (pitch, yaw, roll) radians: the per-sample host twin of
    :func:`compute_euler_angles_from_rotation_matrices` that the datasets use to build labels
    (src/utils.py:263-286).  The singular flag (``sy < 1e-6``) is taken BEFORE the full-range sign flip, as there."""
    R = np.asarray(rotation_matrices)
    sy = np.sqrt(R[0, 0] * R[0, 0] + R[1, 0] * R[1, 0])
    singular = bool(sy < 1e-6)
    if full_range and R[0, 0] < 0:
        sy = -sy
    yaw = math.atan2(-R[2, 0], sy)
    if singular:
        return np.array([math.atan2(-R[1, 2], R[1, 1]), yaw, 0.0])
    return np.array([math.atan2(R[2, 1], R[2, 2]), yaw, math.atan2(R[1, 0], R[0, 0])])


def limit_angle(angle, pi=180.0):
    """Wrap degrees into [-180, 180] (host scalar helper, src/utils.py:289-300)."""
    if angle < -pi:
        k = -2 * (int(angle / pi) // 2)
        angle = angle + k * pi
    if angle > pi:
        k = 2 * ((int(angle / pi) + 1) // 2)
        angle = angle - k * pi
    return angle
