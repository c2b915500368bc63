"""The SMPL model-file loader (hierarchicalprobabilistic3dhuman_b200/smpl.py:_load_model_file) on the CPU: a model written in the
layout of the licence-gated SMPL_{GENDER}.pkl / .npz files (reference configs/paths.py:4, models/smpl_official.py:13-16 via
smplx) must come back as the constants smplx would hold."""
import numpy as np
import pytest

from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
from hierarchicalprobabilistic3dhuman_b200.smpl import _load_model_file
from test_gpu_wrappers import _write_smpl_files


@pytest.mark.parametrize("gender,fname", [("neutral", "SMPL_NEUTRAL.pkl"), ("male", "SMPL_MALE.npz")])
def test_loader_round_trip(tmp_path, gender, fname):
    model = syn.synthetic_smpl_model()
    _write_smpl_files(model, str(tmp_path))
    for path in (str(tmp_path), str(tmp_path / fname)):          # directory + gender, or the file itself
        m = _load_model_file(path, gender)
        assert m is not None
        for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights"):
            assert m[k].dtype == np.float64 and np.array_equal(m[k], model[k]), k
        assert m["parents"].tolist() == model["parents"].tolist() and m["parents"][0] == -1
        assert np.array_equal(m["faces"], model["faces"])
        assert m["extra_vertex_ids"].shape == (21,) and m["joint_regressors_extra"].shape == (45, 6890)
    assert _load_model_file(str(tmp_path), "female") is None and _load_model_file(None, "neutral") is None
