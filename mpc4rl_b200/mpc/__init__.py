"""Mirror of the reference's ``rlmpc.mpc`` package on top of the B200 engine: the same class and
method names (``MPC``, ``AcadosMPC``, ``ocp_solver.set/get/solve/...``) so the reference's drivers
(rlmpc/examples, scripts/) keep working, with every solve executed by the CUDA library."""
