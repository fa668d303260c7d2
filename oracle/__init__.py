"""CPU oracle for the deformable two-view hot path  --  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy fp64, float32 where the reference is
float32) of the reference's triangulation + non-rigid LM refinement path
(/root/reference/Modules/{Utils/Geometry.cc, Calibration, Optimization}).  It is
the *checker* for the CUDA path.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it.  The product
(triangulation-in-deformable-scenes_b200/) never imports, links or executes
anything below oracle/.

PARITY UNPINNED: the reference cannot be compiled in this image (needs Eigen,
Sophus, g2o, OpenCV, Open3D, Qhull, NLopt -- none present, Thirdparty/ is
git-ignored upstream) and it ships no tests, golden vectors or known-answer
fixtures for this path.  The algorithm that lives in the
absent third-party code (g2o's Levenberg-Marquardt, numeric Jacobians,
SE3Quat::exp, Huber kernel) is restated from its published upstream sources
(RainerKuemmerle/g2o, unpinned by the reference) and anchored on the
reference's own call sites, cited function by function.  Weak pins that ARE
checked (tests/test_oracle_pins.py): the Hessian structure rule of
/root/reference/debug.txt, the "pixel sigma ~= 1.0" statistic of the synthetic
database, closed-form triangulation cases, analytic-vs-central-difference
Jacobians, monotone cost under accepted steps.

Round 2 added oracle/outer.py (the Nelder-Mead weight search of deformationOptimization and its objective;
NLopt absent and unpinned: parity unpinned) and oracle/ba.py (bundleAdjustment / localBundleAdjustment /
poseOnlyOptimization, g2oBundleAdjustment.cc:38-444; no executable of the reference calls them: parity unpinned).

Two independently written restatements live here and are checked against each other
(tests/test_c_oracle.py): the numpy modules of this package and the plain-C, OpenMP
oracle/c/dsc_oracle.c (front end: oracle/cport.py; built into oracle/_build/ by
oracle/c/Makefile), which also serves as the compiled multi-threaded CPU baseline.
"""
