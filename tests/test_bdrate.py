"""BD-rate / BD-PSNR calculator against the worked example the reference ships in calc_BDBR/JCTVC-B055.zip
(reference.txt, proposal.txt; expected values printed in the header of BD_Metrics5.c)."""
import importlib

import numpy as np

bd = importlib.import_module("hevc-deep-learning-pipeline_b200.bdrate")

REF = np.array([[999.35, 33.01, 39.27, 40.32], [1598.99, 34.93, 40.08, 41.04], [2499.19, 36.69, 40.97, 41.89],
                [3996.57, 38.42, 41.83, 42.87], [5998.07, 39.79, 42.47, 43.66]])
PRO = np.array([[997.34, 34.68, 40.11, 41.13], [1588.50, 36.64, 40.93, 41.96], [2493.93, 38.34, 41.80, 42.92],
                [3999.06, 39.99, 42.62, 43.93], [5980.18, 41.00, 43.27, 44.77]])
EXPECT_PSNR = (1.628122, 0.828040, 0.993032)
EXPECT_RATE = (-35.976930, -36.433158, -39.302836)


def test_five_point_worked_example():
    for c in range(3):
        assert abs(bd.bd_psnr(REF[:, 0], REF[:, 1 + c], PRO[:, 0], PRO[:, 1 + c], order=4) - EXPECT_PSNR[c]) < 2e-6
        assert abs(bd.bd_rate(REF[:, 0], REF[:, 1 + c], PRO[:, 0], PRO[:, 1 + c], order=4) - EXPECT_RATE[c]) < 2e-5


def test_identity_and_uniform_shift():
    r = np.array([3142.8, 6791.8, 12469.7, 20647.9]); p = np.array([36.399, 39.584, 43.167, 46.771])
    assert abs(bd.bd_rate(r, p, r, p)) < 1e-9 and abs(bd.bd_psnr(r, p, r, p)) < 1e-9
    assert abs(bd.bd_rate(r, p, 1.1 * r, p) - 10.0) < 1e-6          # 10 % more bits at every PSNR
    assert abs(bd.bd_psnr(r, p, r, p - 0.25) + 0.25) < 1e-9         # 0.25 dB lower at every rate


def test_survey_numbers_reproduce():
    """reference HM_dl vs anchor on the 1920x1024 synthetic frame -> BD-rate(Y) +9.15 %."""
    rec = (np.array([20647.9, 12469.7, 6791.8, 3142.8]), np.array([46.771, 43.167, 39.584, 36.399]))
    anchor = (np.array([19861.0, 11978.2, 6476.6, 2856.5]), np.array([46.971, 43.363, 39.815, 36.599]))
    assert abs(bd.bd_rate(anchor[0], anchor[1], rec[0], rec[1]) - 9.15) < 0.1
    assert abs(bd.bd_psnr(anchor[0], anchor[1], rec[0], rec[1]) + 0.474) < 0.01
