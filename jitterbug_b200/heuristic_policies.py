"""Heuristic policies that solve the Jitterbug tasks -- drop-in for ``jitterbug_dmc/heuristic_policies.py``.

Each policy maps a ``TimeStep`` (its normalised observation dict) to a motor command.  These are the
single-env, host-side versions with the reference's names and keyword arguments (heuristic_policies.py:6-136); the
batched device version is ``BatchedJitterbugEnv.heuristic_actions`` (csrc/jb_task.cuh).
"""

import numpy as np


def face_direction(ts, *, angle_to_target=None):
    """Proportional turn controller: 0.9 * clip(3 a / pi, -1, 1) on the (normalised) relative yaw."""
    if angle_to_target is None:
        angle_to_target = ts.observation['angle_to_target']
    return 0.9 * max(min(3 * angle_to_target / np.pi, 1), -1)


def move_forward(ts, *, kick_angle=np.deg2rad(45), speed=0.3, orientation="forward"):
    """Bang-bang "kick" controller on the motor angle; ``orientation`` shifts the kick window by +-pi/2."""
    motor_angle = ts.observation['motor_position'][0]
    motor_vel = ts.observation['motor_velocity'][0]
    offset = {"left": np.pi / 2, "right": -np.pi / 2}.get(orientation, 0)
    if motor_angle < offset - kick_angle:
        return speed
    if motor_angle > offset + kick_angle:
        return -speed
    return speed if motor_vel > 0 else -speed


def move_from_origin(ts):
    """Move Jitterbug away from the origin"""
    return move_forward(ts)


def optimal_orientation_to_move(ts, *, angle_to_target):
    """Choose to walk forwards / sideways so the remaining turn is at most 45 degrees."""
    orientation = "None"
    quarter = 45 * np.pi / 180
    half = 90 * np.pi / 180
    if -quarter <= angle_to_target <= quarter:
        orientation = "forward"
    elif quarter < angle_to_target <= np.pi:
        orientation = "left"
        angle_to_target = abs(abs(angle_to_target) - half)
    elif -np.pi <= angle_to_target < -quarter:
        orientation = "right"
        angle_to_target = -abs(abs(angle_to_target) - half)
    return [angle_to_target, orientation]


def _turn_or_walk(ts, angle_to_target, angle_threshold):
    angle_to_target, orientation = optimal_orientation_to_move(ts, angle_to_target=angle_to_target)
    if np.abs(angle_to_target) > angle_threshold:
        return face_direction(ts, angle_to_target=angle_to_target)
    return move_forward(ts, orientation=orientation)


def move_in_direction(ts, *, angle_threshold=np.deg2rad(20)):
    """Move Jitterbug in a certain direction"""
    return _turn_or_walk(ts, ts.observation['angle_to_target'], angle_threshold)


def move_to_position(ts, *, angle_threshold=np.deg2rad(20)):
    """Move Jitterbug to a certain position"""
    dx, dy = ts.observation['target_in_jitterbug_frame'][0:2]
    return _turn_or_walk(ts, np.arctan2(dx, -dy), angle_threshold)


def move_to_pose(ts, *, angle_threshold=np.deg2rad(20)):
    """Move Jitterbug to a certain pose: turn, walk, then face the target yaw"""
    dx, dy = ts.observation['target_in_jitterbug_frame'][0:2]
    angle_to_target = np.arctan2(dx, -dy)
    if np.abs(angle_to_target) > angle_threshold:
        return face_direction(ts, angle_to_target=angle_to_target)
    if np.linalg.norm([dx, dy]) > 0.01:
        return move_forward(ts)
    return face_direction(ts)


POLICIES = dict(move_from_origin=move_from_origin, face_direction=face_direction, move_in_direction=move_in_direction,
                move_to_position=move_to_position, move_to_pose=move_to_pose)
