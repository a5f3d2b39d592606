"""HMA integration/segregation against the reference's stored known answers
(output/emp_15inds_output_16dic.pickle -> tests/golden/hma_kat.npz, SURVEY.md section 4: exact KAT)."""
import numpy as np

from conftest import load_golden


def test_hma_matches_reference_known_answers():
    from nremmodfc_b200 import HMA
    g = load_golden("hma_kat.npz")
    for k in range(len(g["sFC"])):
        FC = g["sFC"][k].copy()
        num, size, H_all = HMA.Functional_HP(FC)
        assert list(num) == list(g["Clus_num"][k])
        assert all(sum(s) == 90 for s in size) and len(H_all) == 89
        hin, hse = HMA.Balance(FC, num, size)
        hn, sn = HMA.nodal_measures(FC, num, size)
        assert abs(hin - g["Hin"][k]) < 1e-12 and abs(hse - g["Hse"][k]) < 1e-12
        assert np.max(np.abs(hn - g["Hin_node"][k])) < 1e-12 and np.max(np.abs(sn - g["Hse_node"][k])) < 1e-11
    out = HMA.integration_segregation(g["sFC"])
    assert np.allclose(out["Hin"], g["Hin"], atol=1e-12) and out["Hse_node"].shape == (len(g["sFC"]), 90)


def test_hma_clips_in_place_like_the_reference():
    from nremmodfc_b200 import HMA
    rng = np.random.default_rng(0)
    FC = np.corrcoef(rng.normal(size=(20, 50)))
    assert FC.min() < 0
    HMA.Functional_HP(FC)
    assert FC.min() == 0                      # HMA.py:55 clips the caller's matrix (run_many_seeds.py:136 stores it so)
