"""CPU property tests (hypothesis): the oracle equals live cv2 on random odd sizes -- the shapes where the padding
quirk of CLAHE, the ceil-halving of pyrDown and REFLECT_101 at odd borders matter."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import fe_oracle as orc
from oracle.cv2_reference import HAVE_CV2

pytestmark = pytest.mark.skipif(not HAVE_CV2, reason="cv2 not importable")


@settings(max_examples=25, deadline=None)
@given(h=st.integers(40, 200), w=st.integers(40, 260), tx=st.integers(1, 9), ty=st.integers(1, 9),
       clip=st.sampled_from([0.0, 1.0, 2.5, 6.0, 40.0]), seed=st.integers(0, 10**6))
def test_clahe_property(h, w, tx, ty, clip, seed):
    import cv2
    img = np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)
    assert np.array_equal(cv2.createCLAHE(clip, (tx, ty)).apply(img), orc.clahe(img, clip, tx, ty))


@settings(max_examples=25, deadline=None)
@given(h=st.integers(30, 150), w=st.integers(30, 200), seed=st.integers(0, 10**6))
def test_pyrdown_scharr_property(h, w, seed):
    import cv2
    img = np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)
    assert np.array_equal(cv2.pyrDown(img), orc.pyrdown(img))
    d = orc.scharr(img)
    assert np.array_equal(cv2.Scharr(img, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101), d[..., 0])
    assert np.array_equal(cv2.Scharr(img, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101), d[..., 1])
