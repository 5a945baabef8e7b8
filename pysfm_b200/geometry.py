"""Pose algebra used by the sliding-window driver (host side; mirrors pysfm's geometry.py).

    relative_pose                    geometry.py:5-8
    propagate_pose_update(_inplace)  geometry.py:13-25
    rotation_xy / xz / yz            geometry.py:29-41
"""
import numpy as np


def relative_pose(R0, t0, R1, t1):
    """(R01, t01) taking the pose (R0, t0) to (R1, t1)."""
    R_delta = np.dot(R1, R0.T)
    t_delta = t1 - R1.dot(R0.T).dot(t0)
    return R_delta, t_delta


def propagate_pose_update(R0, t0, R0_updated, t0_updated, R1, t1):
    """Apply the update (R0, t0) -> (R0_updated, t0_updated) to the pose (R1, t1)."""
    R_delta = np.dot(R0_updated, R0.T)
    R1_updated = np.dot(R_delta, R1)
    t1_updated = np.dot(R_delta, t1 - t0) + t0_updated
    return R1_updated, t1_updated


def propagate_pose_update_inplace(cam0, cam0_updated, cam1):
    R1_upd, t1_upd = propagate_pose_update(cam0.R, cam0.t, cam0_updated.R, cam0_updated.t, cam1.R, cam1.t)
    cam1.R = R1_upd
    cam1.t = t1_upd


def rotation_xy(th):
    return np.array([[np.cos(th), -np.sin(th), 0.], [np.sin(th), np.cos(th), 0.], [0., 0., 1.]])


def rotation_xz(th):
    return np.array([[np.cos(th), 0., -np.sin(th)], [0., 1., 0.], [np.sin(th), 0., np.cos(th)]])


def rotation_yz(th):
    return np.array([[1., 0., 0.], [0., np.cos(th), -np.sin(th)], [0., np.sin(th), np.cos(th)]])
