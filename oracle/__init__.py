"""CPU oracle for the WBC hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the reference's per-step whole-body-control
path (``wrappers/Robot_Wrapper4.py`` + ``wrappers/QP_Wrapper.py``) and of the third-party
semantics it sits on (Pinocchio rigid-body kinematics, qpOASES QP solve).  It exists to
CHECK the CUDA path; it is never the thing that is shipped or measured.

Two restatements live here: the NumPy/SciPy one (``pin.py``, ``rotation_port.py``, ``qp_wrapper.py``,
``robot_wrapper4.py``), which mirrors the reference class by class, and a plain-C one (``wbc_oracle.c``, loaded by
``c_port.py``) that is pinned against the first (``tests/test_oracle_c_cpu.py``) and is fast enough to check
every state of a 4096-state batch and to serve as the timed CPU baseline.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product package
(``mech5845m-wbc-for-legged-manipulator_b200``) must never import anything from here.

Parity status (see DESIGN.md section "Oracle"):
  * WORLD joint Jacobians at the neutral configuration are PINNED against the reference's
    recorded Pinocchio output ``tests_NOT_FOR_USE/Jacobians.py:1-24`` (fixture
    ``tests/golden/jacobians_neutral_wx200.json``).
  * Tree indexing is PINNED against the reference's own hard-coded slices
    (``Robot_Wrapper4.py:341-345``, ``Jacobians.py:1,18``, nq = 27).
  * Everything else Pinocchio/qpOASES compute (``integrate``, LOCAL / LOCAL_WORLD_ALIGNED
    frames, non-neutral configurations, every QP solution) is "parity unpinned": neither
    library is installable here (no network, not in the wheelhouse) and the reference
    records no outputs for them.  Those pieces are validated by invariants instead
    (finite differences of FK, group identities, KKT certificates, SciPy cross-solvers).
"""
