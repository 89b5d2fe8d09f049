"""Pin oracle/varref_np.py (restatement of cv2.VariationalRefinement) against cv2 and the reference goldens."""
import numpy as np
import pytest

from oracle import flow_ops as ops
from oracle import varref_np as vn
import make_golden as cases
from tobac_flow_b200 import synthetic


def epe(a, b):
    return np.sqrt(((np.asarray(a, np.float64) - b) ** 2).sum(-1))


@pytest.mark.skipif(not ops.have_cv2(), reason="cv2 not importable")
@pytest.mark.parametrize("shape,fp,sor", [((40, 56), 1, 1), ((40, 56), 5, 5), ((61, 83), 5, 5), ((120, 160), 5, 5),
                                          ((33, 47), 2, 3)])
def test_matches_cv2(shape, fp, sor):
    import cv2
    h, w = shape
    bt = synthetic.bt_sequence(2, h, w, seed=h + w, nans=False)
    q0, q1 = ops.pair_to_u8(bt[0], bt[1])
    flow = cv2.calcOpticalFlowFarneback(q0, q1, None, 0.5, 5, 13, 10, 5, 1.1, 0)
    vr = cv2.VariationalRefinement_create()
    vr.setFixedPointIterations(fp)
    vr.setSorIterations(sor)
    ref = vr.calc(q0, q1, flow.copy())
    mine = vn.variational_refinement(q0, q1, flow, fixed_point_iterations=fp, sor_iterations=sor)
    e = epe(mine, ref)
    assert e.max() < 5e-5 and e.mean() < 2e-6, (e.max(), e.mean())
    assert np.abs(ref - flow).max() > 0.05     # the refinement actually moved the field


@pytest.mark.skipif(not ops.have_cv2(), reason="cv2 not importable")
def test_defaults_are_opencvs():
    import cv2
    vr = cv2.VariationalRefinement_create()
    d = vn.DEFAULTS
    assert (vr.getAlpha(), vr.getDelta(), vr.getGamma()) == (d["alpha"], d["delta"], d["gamma"])
    assert abs(vr.getOmega() - d["omega"]) < 1e-6 and abs(vr.getEpsilon() - d["epsilon"]) < 1e-9
    assert (vr.getFixedPointIterations(), vr.getSorIterations()) == (5, 5)


def test_production_settings_vs_reference_golden(golden):
    """scripts/dcc_detect_goes.py:164-166: vr_steps=1, smoothing_passes=1, interp_method='cubic'."""
    g = golden("bt_small_production")
    bt = cases.small_bt()
    fwd, bwd = ops.create_flow(bt, vr_steps=1, backend="numpy")
    assert epe(fwd, g["fwd_vr1"]).max() < 2e-3 and epe(bwd, g["bwd_vr1"]).max() < 2e-3
    fwd, bwd = ops.create_flow(bt, smoothing_passes=1, interp_method="cubic", vr_steps=1, backend="numpy")
    m = np.isfinite(g["fwd_vr1_smooth1_cubic"]).all(-1)
    assert epe(fwd, g["fwd_vr1_smooth1_cubic"])[m].max() < 2e-3
    if ops.have_cv2():
        f2, b2 = ops.create_flow(bt, smoothing_passes=1, interp_method="cubic", vr_steps=1, backend="cv2")
        assert np.array_equal(f2, g["fwd_vr1_smooth1_cubic"], equal_nan=True)
        assert np.array_equal(b2, g["bwd_vr1_smooth1_cubic"], equal_nan=True)
