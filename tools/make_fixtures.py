#!/usr/bin/env python
"""Regenerate the data that has to travel to the GPU box (where /root/reference does not exist).

Run in the build container:   python tools/make_fixtures.py [--reference /root/reference]

Writes
  <package>/data/a1_wx200.json, a1_px100_pin_ver.json, laikago_vx300.json   tree tables extracted by the PRODUCT's URDF walker
                                                        (tree_table.TreeTable.from_urdf)
  tests/golden/jacobians_neutral_wx200.json             the reference's recorded Pinocchio output
                                                        tests_NOT_FOR_USE/Jacobians.py:1-24 (+ CoM block :27-42)
  tests/golden/mocap_rows.json                          64 evenly spaced rows of mocap_{px100,wx200}.txt
                                                        (leg columns permuted FR,FL,RR,RL -> FL,FR,RL,RR)
  tests/golden/standing_configs.json                    the three 27-vectors of wrappers/Robot_Wrapper.py:26-28
  tests/golden/qp_kat.json                              the 3-variable QP of tests_NOT_FOR_USE/qp_tests.py:4-13
                                                        with its exact answer (derived by enumeration)
Nothing here is reference SOURCE: the outputs are numeric tables derived from reference data files.
"""
import argparse
import importlib.util
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mech5845m-wbc-for-legged-manipulator_b200")


def _load_tree_table_module():
    spec = importlib.util.spec_from_file_location("_tt", os.path.join(PKG, "tree_table.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def parse_matrix_block(text):
    rows = re.findall(r"\[([^\[\]]+)\]", text)
    return [[float(t) for t in r.split()] for r in rows]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    ref = args.reference
    tt = _load_tree_table_module()
    os.makedirs(os.path.join(PKG, "data"), exist_ok=True)
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)

    for name in ("a1_wx200", "a1_px100_pin_ver", "laikago_vx300"):
        t = tt.TreeTable.from_urdf(os.path.join(ref, "Robot_Descriptions", "urdf", name + ".urdf"))
        t.save(os.path.join(PKG, "data", name + ".json"))
        print(name, "nq", t.nq, "nv", t.nv, "njoints", t.njoints, "nframes", t.nframes)

    # --- Jacobians.py (a text dump, not python) -----------------------------------------------------
    txt = open(os.path.join(ref, "tests_NOT_FOR_USE", "Jacobians.py")).read()
    parts = re.split(r"Jacobia?n[a-z]* of ", txt)[1:]
    out = {"source": "tests_NOT_FOR_USE/Jacobians.py", "urdf": "a1_wx200", "configuration": "pin.neutral",
           "reference_frame": "WORLD",
           "known_typo": {"block": "joint19", "row": 0, "col": 20, "recorded": 0.362825, "geometry": -0.362825,
                          "why": "col 21 (same axis, same height) is recorded as -0.362825; SURVEY 8c"}}
    for part in parts:
        head, body = part.split(":", 1)
        m = parse_matrix_block(body)
        if head.startswith("Joint 19"):
            out["joint19"] = m
        elif head.startswith("Joint 1 "):
            out["joint1"] = m
        elif head.startswith("Joint 4"):
            out["joint4"] = m
        elif head.startswith("CoM"):
            out["com_stale"] = m
    assert [len(out[k]) for k in ("joint19", "joint1", "joint4")] == [6, 6, 6]
    assert all(len(r) == 26 for k in ("joint19", "joint1", "joint4") for r in out[k])
    json.dump(out, open(os.path.join(gold, "jacobians_neutral_wx200.json"), "w"), indent=1)

    # --- mocap rows --------------------------------------------------------------------------------
    moc = {"source": "tests_NOT_FOR_USE/mocap_{px100,wx200}.txt", "order": "FL,FR,RL,RR,arm (Pinocchio)"}
    for robot in ("px100", "wx200"):
        rows = []
        for line in open(os.path.join(ref, "tests_NOT_FOR_USE", f"mocap_{robot}.txt")):
            v = [float(t) for t in line.replace(",", " ").split()]
            if len(v) < 14:
                continue
            j = v[2:]
            legs = j[3:6] + j[0:3] + j[9:12] + j[6:9]          # FR,FL,RR,RL -> FL,FR,RL,RR
            rows.append(legs + j[12:])
        sel = np.linspace(0, len(rows) - 1, 64).astype(int)
        moc[robot] = [rows[i] for i in sel]
    json.dump(moc, open(os.path.join(gold, "mocap_rows.json"), "w"))

    # --- standing configs --------------------------------------------------------------------------
    src = open(os.path.join(ref, "wrappers", "Robot_Wrapper.py")).read()
    cfgs = [[float(t) for t in m.replace(" ", "").split(",")]
            for m in re.findall(r"stand_joint_config = np\.array\(\[([^\]]+)\]\)", src)]
    cfgs = [c for c in cfgs if len(c) == 27]
    json.dump({"source": "wrappers/Robot_Wrapper.py:26-28", "configs": cfgs},
              open(os.path.join(gold, "standing_configs.json"), "w"), indent=1)

    # --- qp KAT ------------------------------------------------------------------------------------
    M = np.array([[1., 2., 0.], [-8., 3., 2.], [0., 1., 1.]])
    kat = {"source": "tests_NOT_FOR_USE/qp_tests.py:4-13 (prints only; answer derived by active-set enumeration)",
           "P": (M.T @ M).tolist(), "q": (np.array([3., 2., 3.]) @ M).tolist(),
           "G": [[1., 2., 1.], [2., 0., 1.], [-1., 2., -1.]], "h": [1., 1., 1.], "A": [[1., 1., 1.]], "b": [1.],
           "x": [0., 0., 1.], "objective": 9.5, "active": ["G0", "G1", "A0"]}
    json.dump(kat, open(os.path.join(gold, "qp_kat.json"), "w"), indent=1)
    print("fixtures written")


if __name__ == "__main__":
    sys.exit(main())
