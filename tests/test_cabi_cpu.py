"""CPU tests (-m "not gpu") of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/wbc_b200.h declares, the ctypes mirrors have the C layout, host-only entry points behave, and the product
path fails loudly (no CPU fallback) when there is no CUDA device.  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wbc_b200.h")


@pytest.fixture(scope="module")
def lib():
    import wbc_b200
    wbc_b200.build_library(force=False)
    return wbc_b200.load_library()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wbc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from wbc_b200 import _cabi
    declared = _declared_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/wbc_b200.h but not exported"
    assert sorted(_cabi.EXPORTS) == declared, "the ctypes binding and the header disagree about the entry points"
    assert lib.wbc_abi_version() == 2
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (wbc_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_header_is_plain_c_and_ctypes_layout_matches(tmp_path):
    """Compile the header as C (gcc, no CUDA, no torch types) and compare sizeof / offsetof with the ctypes mirrors."""
    from wbc_b200 import _cabi
    structs = {"WbcTreeTable": _cabi.WbcTreeTable, "WbcConfig": _cabi.WbcConfig, "WbcStepIO": _cabi.WbcStepIO,
               "WbcAssembleOut": _cabi.WbcAssembleOut, "WbcHostIO": _cabi.WbcHostIO}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for sname, cls in structs.items():
        lines.append(f'  printf("{sname} %zu\\n", sizeof({sname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{sname}.{fname} %zu\\n", offsetof({sname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for sname, cls in structs.items():
        assert int(got[sname]) == C.sizeof(cls), sname
        for fname, _ in cls._fields_:
            assert int(got[f"{sname}.{fname}"]) == getattr(cls, fname).offset, (sname, fname)
    text = open(HEADER).read()
    assert "torch" not in text.lower().replace("pytorch tensors", "") and "at::" not in text and "std::" not in text


def test_config_rows_host_helper(lib):
    """m and nC implied by setTasks / setConstraints (Robot_Wrapper4.py:839-876, 764-836) for the three step patterns."""
    from wbc_b200 import _cabi as cabi
    m, nc = C.c_int32(), C.c_int32()
    cfg = cabi.WbcConfig()
    all_cart = cabi.TASK_FR | cabi.TASK_FL | cabi.TASK_RR | cabi.TASK_RL | cabi.TASK_GRIP | cabi.TASK_TRUNK
    feet = cabi.CON_FR | cabi.CON_FL | cabi.CON_RR | cabi.CON_RL
    cases = [  # (task mask, constraint mask, extra rows, nv) -> (m, nC)
        (all_cart | cabi.TASK_JOINT, 0, 0, 26, 62, 0),                          # P1 bootstrap, WX200
        (all_cart | cabi.TASK_JOINT, 0, 0, 25, 61, 0),                          # P1 bootstrap, PX100
        (cabi.TASK_GRIP | cabi.TASK_JOINT, cabi.CON_TRUNK | feet, 0, 26, 32, 16),   # P2 sim3 tick
        (all_cart | cabi.TASK_JOINT, cabi.CON_TRUNK | feet, 0, 26, 62, 16),     # P3 benchmark step
        (all_cart | cabi.TASK_JOINT, cabi.CON_TRUNK | feet | cabi.CON_COM, 12, 26, 62, 30),   # config 3: + extension rows
        (cabi.TASK_TRUNK, cabi.CON_GRIP, 0, 26, 6, 3),
    ]
    for tm, cm, extra, nv, em, enc in cases:
        cfg.task_mask, cfg.constraint_mask, cfg.n_extra_rows = tm, cm, extra
        assert lib.wbc_config_rows(C.byref(cfg), nv, C.byref(m), C.byref(nc)) == 0
        assert (m.value, nc.value) == (em, enc)
    assert lib.wbc_config_rows(None, 26, C.byref(m), C.byref(nc)) == cabi.ERR_INVALID_ARG
    assert b"null" in lib.wbc_last_error()


def test_argument_validation_precedes_any_cuda_work(lib):
    """Bad arguments come back as WBC_ERR_INVALID_ARG with a message -- also on a box without a GPU."""
    from wbc_b200 import _cabi as cabi
    h = C.c_void_p()
    assert lib.wbc_model_create(None, C.byref(h)) == cabi.ERR_INVALID_ARG
    t = cabi.WbcTreeTable()
    t.njoints, t.nq, t.nv, t.nframes = 1, 7, 6, 6
    assert lib.wbc_model_create(C.byref(t), C.byref(h)) == cabi.ERR_INVALID_ARG and b"njoints" in lib.wbc_last_error()
    assert lib.wbc_qp_solve(4, 40, 10, 0, None, None, None, None, None, None, None, None, None, 10, None, None, None,
                            None, None) == cabi.ERR_INVALID_ARG
    assert lib.wbc_qp_solve(4, 8, 10, 40, None, None, None, None, None, None, None, None, None, 10, None, None, None,
                            None, None) == cabi.ERR_UNSUPPORTED
    assert lib.wbc_step(None, None, None, 1, None) == cabi.ERR_INVALID_ARG
    assert lib.wbc_fk_jac(None, None, 1, None, 0, 0, None, None, None) == cabi.ERR_INVALID_ARG
    lib.wbc_model_destroy(None)                      # a no-op, must not crash


def test_tree_table_marshalling_matches_json():
    """TreeTable (host) -> WbcTreeTable (C struct): indices, placements, frame slots 0..4 EE + 5 trunk, limits."""
    from wbc_b200 import TreeTable, _cabi as cabi
    from wbc_b200.robot_model import RobotModel, EE_FRAME_NAMES
    t = TreeTable.load("a1_wx200")
    slots = [t.getFrameId(n, "FIXED_JOINT") for n in EE_FRAME_NAMES] + [t.getFrameId("imu_joint", "FIXED_JOINT")]
    T = RobotModel._make_table(t, slots)
    assert (T.njoints, T.nq, T.nv, T.nframes) == (22, 27, 26, 6)
    assert list(T.parent[:22]) == list(t.parent) and list(T.idx_v[2:22]) == list(range(6, 26))
    assert [T.frame_parent[s] for s in range(6)] == [7, 4, 13, 10, 18, 1]
    assert abs(T.frame_p[4][0] - 0.043) < 1e-15 and abs(T.frame_p[0][2] + 0.2) < 1e-15
    assert T.jtype[1] == cabi.JT_FREEFLYER and T.jtype[20] == cabi.JT_PRISMATIC
    assert abs(T.placement_p[14][2] - 0.124175) < 1e-12
    assert abs(sum(T.mass[j] for j in range(22)) - sum(t.mass)) < 1e-12


def test_product_path_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import wbc_b200
    with pytest.raises(wbc_b200.WbcError):
        wbc_b200.RobotModel("a1_wx200", batch=4)
    with pytest.raises(wbc_b200.WbcError):
        wbc_b200.QP(np.eye(3), np.zeros(3), -np.ones(3), np.ones(3))


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package, the C sources or the header may name it."""
    pkg = os.path.join(ROOT, "mech5845m-wbc-for-legged-manipulator_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "oracle." in re.sub(r"#.*|//.*", "", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    code = ("import sys; sys.path.insert(0, %r); import wbc_b200; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_mocap_ingestion_matches_the_fixture_rows(tmp_path):
    """f4: playback files (tests_NOT_FOR_USE/mocap_wx200.txt) are FR, FL, RR, RL in file order; the golden rows are the same
    frames already in Pinocchio order (tools/make_fixtures.py read them from the reference)."""
    import json
    import numpy as np
    from wbc_b200 import mocap
    rows = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "mocap_rows.json")))["wx200"])
    bullet = mocap.pinocchio_to_bullet(rows)
    path = tmp_path / "mocap.txt"
    with open(path, "w") as fh:
        for i, r in enumerate(bullet):
            fh.write(f"{i + 1:12d}, {588975 + 2 * i:11d}, " + ", ".join(f"{v:11.6f}" for v in r) + "\n")
    frames = mocap.load_mocap(str(path), n_joints=20)
    assert frames.shape == rows.shape and np.abs(frames - rows).max() < 1e-9
    assert np.abs(mocap.bullet_to_pinocchio(mocap.pinocchio_to_bullet(rows)) - rows).max() == 0.0
    q = mocap.configurations(frames)
    assert q.shape == (rows.shape[0], 27) and (q[:, 6] == 1.0).all()


def test_config_marshalling_bulk_copies_match_the_attributes():
    """RobotModel._config(): the controller's weights / gains / switches land in the WbcConfig fields the kernel reads
    (bulk NumPy copies into the ctypes arrays; the gain list is taken as given, quirk D.8)."""
    from wbc_b200 import _cabi as cabi
    from wbc_b200.robot_model import RobotModel
    r = RobotModel.__new__(RobotModel)                       # no device needed: _config only reads attributes
    rng = np.random.default_rng(0)
    r.setTasks(Trunk=True, FR=True, FL=False, RR=True, RL=True, Grip=True, Joint="HYBRID")
    r.setConstraints(CoM=False, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)
    r.compat_damper_off_by_one, r.end_effector_index_list_joint, r.arm_base_id, r.max_qp_iterations = True, [7, 4, 13, 10, 19], 14, 77
    r.EE_weight = [rng.normal(size=(6, 6)) for _ in range(5)]
    r.EE_gains = [rng.normal(size=(6, 6)) for _ in range(5)]
    r.trunk_weight, r.trunk_gain = rng.normal(size=(6, 6)), rng.normal(size=(6, 6))
    r.cart_task_weight_EE_list, r.cart_task_weight_Trunk, r.joint_task_weight = [1, 2, 3, 4, 5], 6, 0.05
    r.damper, r.extra_rows = (0.01, 0.026, 0.015), [(1, 2, [1, 2, 3, 4, 5, 6], -1.0, 1.0)]
    c = r._config()
    for i in range(5):
        assert np.array_equal(np.array(c.ee_weight[i][:]), r.EE_weight[i].reshape(-1))
        assert np.array_equal(np.array(c.ee_gain_pos[i][:]), r.EE_gains[i][:3, :3].reshape(-1))
    assert np.array_equal(np.array(c.trunk_weight[:]), r.trunk_weight.reshape(-1))
    assert list(c.cart_task_weight[:]) == [1, 2, 3, 4, 5, 6] and c.joint_task_weight == 0.05
    assert np.array_equal(np.array(c.trunk_gain_pos[:]), r.trunk_gain[:3, :3].reshape(-1))
    assert np.array_equal(np.array(c.trunk_gain_ori[:]), np.diagonal(r.trunk_gain)[3:])
    assert c.task_mask == (cabi.TASK_TRUNK | cabi.TASK_FR | cabi.TASK_RR | cabi.TASK_RL | cabi.TASK_GRIP | cabi.TASK_JOINT)
    assert c.joint_mode == cabi.JOINT_HYBRID and c.max_iter == 77 and c.gripper_joint_id == 19 and c.arm_base_id == 14
    assert c.n_extra_rows == 1 and list(c.extra_coeff[0][:]) == [1, 2, 3, 4, 5, 6] and (c.extra_lo[0], c.extra_hi[0]) == (-1.0, 1.0)


def test_standing_sampler_and_config3_rows():
    """The closed-loop sampler stays inside the URDF limits with a near-level trunk; config 3's rows are 16 pyramid faces
    + 7 torque-limit proxy rows built from the URDF's effort / velocity limits."""
    from wbc_b200 import synthetic, TreeTable
    for name in ("a1_wx200", "a1_px100_pin_ver", "laikago_vx300"):
        t = TreeTable.load(name)
        q = synthetic.sample_standing(t, 512, 3)
        lo, up = np.asarray(t.lower[7:t.nq]), np.asarray(t.upper[7:t.nq])
        assert (q[:, 7:] >= lo - 1e-12).all() and (q[:, 7:] <= up + 1e-12).all()
        assert np.abs(np.linalg.norm(q[:, 3:7], axis=1) - 1).max() < 1e-12 and (q[:, 6] > 0.999).all()
    t = TreeTable.load("a1_wx200")
    rows = synthetic.config3_rows(t)
    assert len(rows) == 23 and all(len(r[2]) == 6 for r in rows)
    assert [r[0] for r in rows[:16]] == [f for f in range(4) for _ in range(4)] and all(r[4] == 0.0 for r in rows[:16])
    assert all(r[3] == -r[4] and r[4] > 0 for r in rows[16:])


def build_c_demo(out_dir):
    """gcc (C11, no CUDA compiler, no Python, no torch) -> examples/c_abi_demo linked against libwbc_b200.so + libcudart."""
    from wbc_b200 import _cabi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = os.path.join(str(out_dir), "c_abi_demo")
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.run(["gcc", "-O2", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
                    os.path.join(root, "examples", "c_abi_demo.c"), "-o", exe, "-L", libdir, "-lwbc_b200",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"],
                   check=True)
    return exe


def test_c_program_compiles_and_links_against_the_abi(lib, tmp_path):
    """The boundary is usable from plain C: examples/c_abi_demo.c (one batched runWBC tick through wbc_model_create + wbc_step)
    compiles with gcc -std=c11 -Wall -Werror against include/wbc_b200.h and links against the shared library.  (It runs in
    tests/test_gpu_surface.py::test_c_abi_demo_matches_python_path.)"""
    exe = build_c_demo(tmp_path)
    assert os.path.exists(exe)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 1 and "usage" in out.stderr            # argument check only: no device is touched
