"""oracle/ -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

A float64 CPU restatement of the reference's hot path
(``rlmpc/mpc/common/mpc.py:52-96,177-202`` -> ``AcadosOcpSolver.solve()`` ->
``rlmpc/mpc/nlp.py:1341-1563 update_nlp``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker / the
reported CPU baseline.

PARITY UNPINNED: the arithmetic of the reference lives in third-party native
libraries that are neither vendored nor version-pinned by the reference and are
absent from this image (acados + HPIPM + BLASFEO, CasADi, SuperLU via SciPy; see
SURVEY.md section 8(c)).  The reference ships no golden vectors either.  This
restatement is therefore pinned only by (i) the reference's own relational
assertions (``nlp.py:1445-1537``: KKT self-consistency), (ii) its
finite-difference checks (``scripts/linear_system_mpc_nlp.py:43-49,87-93``),
(iii) the LQR closed form for the linear system and (iv) scipy.optimize
cross-checks -- all exercised in ``tests/test_oracle.py``.

Modules
  problems.py  restated problem definitions (models, costs, bounds) in torch f64
  nlp.py       restated build_nlp / update_nlp: z ordering, h rows, L, R, dR/dz,
               dense sparse-LU sensitivity solve -- derivatives by torch.func
  solver.py    dense SQP + primal-dual IPM stand-in for acados SQP + HPIPM
  cpu_port/    C++ host build of the engine maths (bench cpu_baseline "port")
"""
