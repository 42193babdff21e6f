"""Regenerates the SO-101 entries of tests/golden/fk_vectors.json from the reference's own MJCF.
Runs only where /root/reference exists (this container); reads the XML directly, independent of the model tables."""
import json
import xml.etree.ElementTree as ET

import numpy as np

XML = "/root/reference/assets/robots/SO-ARM100/Simulation/SO101/so101_old_calib.xml"


def qmat(q):
    w, x, y, z = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rz(a):
    return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])


def fk(q):
    root = ET.parse(XML).getroot().find("worldbody")
    out, mass, k = {}, 0.0, [0]

    def walk(body, R, p):
        nonlocal mass
        Rb = R @ qmat([float(x) for x in body.get("quat", "1 0 0 0").split()])
        pb = p + R @ np.array([float(x) for x in body.get("pos", "0 0 0").split()])
        if body.find("joint") is not None:
            Rb = Rb @ rz(q[k[0]]); k[0] += 1
        out[body.get("name")] = (pb, Rb)
        mass += float(body.find("inertial").get("mass"))
        for c in body.findall("body"):
            walk(c, Rb, pb)
    walk(root.find("body"), np.eye(3), np.zeros(3))
    return out, mass


if __name__ == "__main__":
    for name, q in (("so101_zero", np.zeros(6)), ("so101_stack_home", np.deg2rad([0, -177, 165, 72, -83, 0]))):
        poses, mass = fk(q)
        print(name, "gripper_pos", poses["gripper"][0].round(6).tolist(), "jaw_pos", poses["moving_jaw_so101_v1"][0].round(6).tolist(),
              "total_mass", round(mass, 6))
