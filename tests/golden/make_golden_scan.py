#!/usr/bin/env python
"""Golden vectors of the fan-shaped scan sensor, written by the UNMODIFIED reference class
``ScanSensor`` (graph_based_slam.py:73-167) in the authoring container.

    python tests/golden/make_golden_scan.py       # writes tests/golden/scan_sensor.npz

Needs /root/reference; imported behind oracle/mpl_stub.py like make_golden.py.  Each case records the
landmark map, the robot pose, the np.random seed and what ``scan(pose)`` returned: detected landmark ids and
their (distance, direction, orientation) with and without noise.
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MCL_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import mpl_stub  # noqa: E402

mpl_stub.install()
with redirect_stdout(io.StringIO()):
    import graph_based_slam as ref_gs  # noqa: E402  (the reference, unmodified)


def unpack(obs):
    ids = np.array([o.getLandMarkId() for o in obs], dtype=np.int64)
    val = np.array([[o.getDist(), o.getDir(), o.getOrient()] for o in obs]).reshape(-1, 3)
    return ids, val


def main():
    out = {}
    rng = np.random.RandomState(2024)
    maps = [np.array(ref_gs.LAND_MARKS, dtype=np.float64), rng.uniform(-20, 20, (64, 2)), rng.uniform(-12, 12, (16, 2))]
    case = 0
    for lm in maps:
        for k in range(4):
            pose = np.array([[rng.uniform(-8, 8)], [rng.uniform(-8, 8)], [rng.uniform(-np.pi, np.pi)]])
            seed = 100 + case
            sens = ref_gs.ScanSensor(ref_gs.SCN_SENS_RANGE_m, ref_gs.SCN_SENS_ANGLE_rps, lm)
            sens.setNoiseParam(5, 2, 2)  # the values the reference's Robot uses (graph_based_slam.py:604)
            np.random.seed(seed)
            noisy, clean = sens.scan(pose)
            ids, nv = unpack(noisy)
            ids2, cv = unpack(clean)
            assert np.array_equal(ids, ids2)
            out[f"c{case}_lm"] = lm
            out[f"c{case}_pose"] = pose
            out[f"c{case}_seed"] = np.int64(seed)
            out[f"c{case}_ids"] = ids
            out[f"c{case}_noisy"] = nv
            out[f"c{case}_clean"] = cv
            case += 1
    out["n_cases"] = np.int64(case)
    out["scan_range"] = np.float64(ref_gs.SCN_SENS_RANGE_m)
    out["scan_angle"] = np.float64(ref_gs.SCN_SENS_ANGLE_rps)
    out["noise_param"] = np.array([5 / 100, np.deg2rad(2.0), np.deg2rad(2.0)])
    path = os.path.join(HERE, "scan_sensor.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", case, "cases; detections per case:",
          [int(out[f"c{c}_ids"].size) for c in range(case)])


if __name__ == "__main__":
    main()
