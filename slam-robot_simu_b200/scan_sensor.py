"""Host side of the fan-shaped scan sensor: the observation the MCL_LIK_RB_FAN likelihood consumes.

Mirrors ``ScanSensor`` of takuyani/SLAM-Robot_Simu graph_based_slam.py:73-167: same constructor arguments
(range, half angle, landmark map), ``setNoiseParam(dist %, direction deg, orientation deg)`` (:115-126), and
``scan(pose)`` drawing ``np.random.normal`` in the reference's order (:163-165: distance, direction, orientation
per detected landmark).  ``scan`` returns the (L, 2) table of (distance, direction) the engine takes, with NaN rows
for landmarks outside the fan, plus the detected ids.  Once per step on one pose: host work, like
``ParticleFilter.__observation``.
"""
from __future__ import annotations

import numpy as np

from .particle_filter import limit_angle, world2robot

_BASE_ANG = np.pi / 2.0  # mylib/transform.py:12


class ScanSensor:
    # graph_based_slam.py:79-81 (class defaults); the reference's Robot sets (5, 2, 2) (:604)
    def __init__(self, aRange_m=15.0, aAngle_rad=np.deg2rad(80.0), aLandMarks=None):
        self.range_m = float(aRange_m)
        self.angle_rad = float(aAngle_rad)
        self.landmarks = np.asarray(aLandMarks, dtype=np.float64).reshape(-1, 2)
        self.r_dist = 10 / 100
        self.dir_sigma = np.deg2rad(3.0)
        self.orient_sigma = np.deg2rad(3.0)

    def setNoiseParam(self, aDist, aDirSigma, aOrientSigma):
        self.r_dist = aDist / 100
        self.dir_sigma = np.deg2rad(aDirSigma)
        self.orient_sigma = np.deg2rad(aOrientSigma)

    def detect(self, pose):
        """(robot-frame landmarks, distance, direction, detected flags): :150-160."""
        p = world2robot(pose, self.landmarks)
        dist = np.linalg.norm(p, axis=1)
        dirs = np.arctan2(p[:, 1], p[:, 0])
        scan_rad = _BASE_ANG - self.angle_rad
        flg = (dist <= self.range_m) & (p[:, 1] >= np.absolute(p[:, 0]) * np.tan(scan_rad))
        return p, dist, dirs, flg

    def scan(self, pose, noise=True):
        """(z (L, 2) with NaN rows where nothing was detected, ids of the detected landmarks)."""
        pose = np.asarray(pose, dtype=np.float64).reshape(3, 1)
        _, dist, dirs, flg = self.detect(pose)
        orient = _BASE_ANG - pose[2, 0]
        z = np.full((self.landmarks.shape[0], 2), np.nan)
        ids = np.nonzero(flg)[0]
        for i in ids:
            if noise:
                z[i, 0] = np.random.normal(dist[i], dist[i] * self.r_dist)  # :163
                z[i, 1] = limit_angle(np.random.normal(dirs[i], self.dir_sigma))  # :164
                np.random.normal(orient, self.orient_sigma)  # :165 (drawn by the reference; the filter does not use it)
            else:
                z[i, 0], z[i, 1] = dist[i], dirs[i]
        return z, ids
