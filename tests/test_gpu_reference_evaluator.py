"""The REFERENCE's own evaluator running on the B200 backend: `oracle/run_ref_evaluator.py` imports the
unmodified `tests/evaluator.py` + `parses/*.py` of the reference, calls `yoho_b200.dropin.install()`, and drives
`Evaluator_PartI.run_onescene` / `Evaluator_PartII.run_onescene` (tests/evaluator.py:41-47,112-117) on a stub dataset; every
artefact it leaves on disk is compared with what the unmodified reference wrote for the same inputs (tests/golden/*.npz).
The CPU leg runs the same script with the unmodified reference as the backend (pins the goldens a second way)."""
import json
import os
import subprocess
import sys
import numpy as np
import pytest

from conftest import load_golden, ROOT

SCRIPT = os.path.join(ROOT, "oracle", "run_ref_evaluator.py")


def _have_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    return ref_shim.available()


def _run(tmp, *args):
    r = subprocess.run([sys.executable, SCRIPT, "--work", str(tmp)] + list(args), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def _artefacts(tmp):
    b = os.path.join(str(tmp), "cache", "Testset", "synth", "scene")
    m = os.path.join(b, "Match")
    c = np.load(os.path.join(m, "YOHO_C", "1000iters", "0-1.npz"), allow_pickle=True)
    o = np.load(os.path.join(m, "YOHO_O", "1000iters", "0-1.npz"), allow_pickle=True)
    return dict(eqv0=np.load(os.path.join(b, "YOHO_Output_Group_feature", "0.npy")),
                eqv1=np.load(os.path.join(b, "YOHO_Output_Group_feature", "1.npy")),
                matches=np.load(os.path.join(m, "0-1.npy")), dr_index=np.load(os.path.join(m, "DR_index", "0-1.npy")),
                trans_pre=np.load(os.path.join(m, "Trans_pre", "0-1.npy")),
                c_trans=c["trans"], c_center=c["center"], c_recalltime=int(c["recalltime"]),
                o_trans=o["trans"], o_recalltime=int(o["recalltime"]),
                c_prelog=open(os.path.join(m, "YOHO_C", "1000iters", "pre.log"), "rb").read(),
                o_prelog=open(os.path.join(m, "YOHO_O", "1000iters", "pre.log"), "rb").read())


def test_reference_backend_reproduces_goldens_cpu(tmp_path):
    """Unmodified reference, CPU, through its own evaluator (incl. yohoc_mul's forked pool): the committed goldens, bit for bit."""
    if not _have_reference():
        pytest.skip("reference sources not available (neither /root/reference nor oracle/_ref/src)")
    g = load_golden("pipeline_synth.npz")
    info = _run(tmp_path, "--backend", "reference", "--device", "cpu", "--c-seed", str(int(g["c_seed"])), "--o-seed", str(int(g["o_seed"])))
    assert info["estimator_class"].endswith("yohoc_mul")          # tests/evaluator.py:36-38: max_iter > 500
    a = _artefacts(tmp_path)
    for k in ("matches", "dr_index", "c_trans", "c_center", "o_trans", "eqv0", "eqv1", "trans_pre"):
        assert np.array_equal(a[k], g[k]), k
    assert a["c_recalltime"] == int(g["c_recalltime"]) and a["o_recalltime"] == int(g["o_recalltime"])
    assert a["c_prelog"] == bytes(g["c_prelog"]) and a["o_prelog"] == bytes(g["o_prelog"])


@pytest.mark.gpu
@pytest.mark.parametrize("golden,weights", [("pipeline_synth.npz", "synth"), ("pipeline_realckpt.npz", "real")])
def test_reference_evaluator_on_b200_backend(tmp_path, golden, weights):
    if not _have_reference():
        pytest.skip("oracle/_ref/src not exported (__graft_entry__.build() with /root/reference present)")
    if weights == "real" and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ckpt", "PartI.npz")):
        pytest.skip("oracle/_ref/ckpt not extracted")
    g = load_golden(golden)
    info = _run(tmp_path, "--backend", "yoho_b200", "--weights", weights, "--c-seed", str(int(g["c_seed"])),
                "--o-seed", str(int(g["o_seed"])))
    assert "yoho_b200" in info["extractor_file"] and info["estimator_class"] == "yoho_b200.estimator.yohoc_mul"
    assert "oracle/_ref/src" in info["evaluator_file"] or info["evaluator_file"].startswith("/root/reference")
    a = _artefacts(tmp_path)
    assert a["eqv0"].dtype == np.float32 and np.abs(a["eqv0"] - g["eqv0"]).max() <= 1e-4
    assert np.abs(a["eqv1"] - g["eqv1"]).max() <= 1e-4
    # from here on every stage consumed THIS backend's descriptors (not the reference's): index work must still be identical
    assert a["matches"].dtype == np.int64 and np.array_equal(a["matches"], g["matches"])
    assert a["dr_index"].dtype == np.int64 and np.array_equal(a["dr_index"], g["dr_index"])
    assert a["c_recalltime"] == int(g["c_recalltime"])
    assert np.array_equal(a["c_center"], g["c_center"])
    assert np.abs(a["c_trans"] - g["c_trans"]).max() <= 1e-9
    assert a["trans_pre"].dtype == np.float64 and np.abs(a["trans_pre"] - g["trans_pre"]).max() <= 1e-4
    # YOHO-O picks among this backend's Trans_pre: same winner position, its transform within the PartII bar
    assert a["o_recalltime"] == int(g["o_recalltime"])
    assert np.abs(a["o_trans"] - g["o_trans"]).max() <= 1e-4
    assert info["fmr"] == 1.0
