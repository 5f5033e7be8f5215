"""Writes tests/golden/known_answers.npz: the exact known answers the reference's own tests and
docs hold for the hot path (nd4js src/la/matmul_test.js:32-62, src/help.js:1876-1885).  They are
literal values from those files, not outputs of our oracle."""
import os
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
np.savez(
    os.path.join(here, "known_answers.npz"),
    mm1_a=np.array([[1.0], [2.0]]), mm1_b=np.array([[30.0, 40.0, 50.0]]),
    mm1_c=np.array([[30.0, 40.0, 50.0], [60.0, 80.0, 100.0]]),
    mm2_a=np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]),
    mm2_b=np.array([[70.0, 80.0], [90.0, 100.0], [110.0, 120.0]]),
    mm2_c=np.array([[580.0, 640.0], [1390.0, 1540.0]]),
    chain_a=np.array([[1.0, 2.0, 3.0, 4.0]]),
    chain_b=np.array([[11.0, 12, 13], [21, 22, 23], [31, 32, 33], [41, 42, 43]]),
    chain_c=np.array([[5.0, 6], [7, 8], [9, 10]]),
    chain_abc=np.array([[6760.0, 7720.0]]),
    chol_s=np.array([[25.0, -50.0], [-50.0, 101.0]]), chol_l=np.array([[5.0, 0.0], [-10.0, 1.0]]),
    svd_int=np.array([[1.0, 1], [1, 2], [1, 3], [1, 4], [1, 5]]),
)
