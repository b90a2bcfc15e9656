"""CPU: the oracle's restatement of the prediction path (``GDMLPredict.predict``, ``Desc.from_R``,
``create_model``'s R_d_desc_alpha, ``_recov_int_const``) against goldens produced by the unmodified reference
(tests/golden/make_golden.py predict)."""
import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle import sgdml_oracle as orc

CASES = ['predict_eth_s6_m24', 'predict_asp_s1_m10']


@pytest.mark.parametrize('case', CASES)
def test_predict_matches_reference(case):
    g = load_golden(case)
    M, N = int(g['M']), int(g['N'])
    R_desc = g['R_desc_T'].T
    xd, gd = orc.desc_from_R(g['R_train'])
    assert relerr(xd, R_desc) < 1e-14
    assert relerr(orc.d_desc_dot_vec(gd, g['alphas_F'].reshape(M, -1)), g['R_d_desc_alpha']) < 1e-12
    E, F = orc.predict(R_desc, g['R_d_desc_alpha'], g['tril_perms_lin'], int(g['sig']), float(g['std']), float(g['c']),
                       g['R_query'].reshape(-1, N, 3))
    assert relerr(E, g['E_query']) < 1e-10 and relerr(F, g['F_query']) < 1e-10
    # training-mode prediction and the integration constant derived from it
    E_tr, F_tr = orc.predict(R_desc, g['R_d_desc_alpha'], g['tril_perms_lin'], int(g['sig']), float(g['std']), 0.0,
                             g['R_train'])
    assert relerr(F_tr, g['F_train_pred']) < 1e-10
    assert relerr(E_tr + float(g['c']), g['E_train_pred']) < 1e-10
    c = orc.recov_int_const(E_tr, g['E_train'])
    assert c is not None and abs(c - float(g['c'])) <= 1e-9 * abs(float(g['c']))
    # flipped force labels are detected (the reference returns None and turns energies off, train.py:1056-1062)
    assert orc.recov_int_const(-E_tr, g['E_train']) is None


def test_rule_of_thumb_matches_reference_table():
    """k = (k_min^m m n^2 / 2)^(1/(2+m)) with the per-molecule fits of plot_data.py:677-706."""
    from mlff_preconditioner_b200.tools import rule_of_thumb as rt

    assert rt.get_params('ethanol') == (0.87, 10, 1) and rt.get_params('aspirin') == (1.14, 236, 1)
    assert rt.get_params('larger_aims_nanotube') == (0.73, 89, 1)
    with pytest.raises(NotImplementedError):
        rt.get_params('water')
    assert rt.rule_of_thumb(108000, 10, 0.87) == 4839      # BASELINE.json configs[1]
    assert rt.rule_of_thumb(9990, 89, 0.73) == 1954        # configs[0]
    assert rt.default_rank('synthetic_ethanol', 108000) == 4839
    assert rt.default_rank('ethanol', 324) == 81           # capped at n/4 on tiny systems
    k = rt.default_rank('aspirin', 1260000)
    assert int(rt.default_break_percentage('aspirin', 1260000) * 1260000) == k
    arr = rt.rule_of_thumb(np.array([1e4, 1e5]), 10, 0.87)
    assert arr.shape == (2,) and arr[1] > arr[0]
