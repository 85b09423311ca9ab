"""The UNMODIFIED reference model file pcdet/models/rpn/rpn_backbone.py on the drop-in spconv modules.

Runs only where /root/reference exists (the build container; the GPU box has no reference tree): install_as_spconv()
registers pcdet_b200.spconv under the name `spconv`, the reference file is imported from where it lies (with PCDet's global
`cfg` stubbed), its BackBone8x is instantiated and must have exactly the module tree / state-dict layout of
pcdet_b200.backbone.BackBone8x, so that checkpoints move between the two.  The committed golden vector
tests/golden/ref_python.npz (made by tests/golden/make_golden.py the same way) carries the same layout to the GPU box."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PCDET_REFERENCE", "/root/reference")
needs_reference = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "pcdet/models/rpn/rpn_backbone.py")),
                                     reason="reference tree not present")


def _load_golden_helpers():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def reference_backbone():
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "spconv" or k.startswith("spconv.") or k == "pcdet" or k.startswith("pcdet.")}
    import pcdet_b200.spconv as sp
    sp.install_as_spconv()
    mg = _load_golden_helpers()
    mod = mg.load_reference_module("pcdet/models/rpn/rpn_backbone.py", "pcdet.models.rpn.rpn_backbone", mg.stub_packages())
    yield mod
    for k in [k for k in sys.modules if k == "spconv" or k.startswith("spconv.") or k == "pcdet" or k.startswith("pcdet.")]:
        if k not in saved:
            del sys.modules[k]
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v


@needs_reference
def test_reference_backbone_builds_on_the_drop_in_modules(reference_backbone):
    from pcdet_b200.backbone import BackBone8x
    ref_net = reference_backbone.BackBone8x(4)
    ours = BackBone8x(4)
    ref_sd, our_sd = ref_net.state_dict(), ours.state_dict()
    assert list(ref_sd.keys()) == list(our_sd.keys())
    assert [tuple(v.shape) for v in ref_sd.values()] == [tuple(v.shape) for v in our_sd.values()]
    # a checkpoint of one loads into the other, both ways
    ours.load_state_dict(ref_sd)
    ref_net.load_state_dict(our_sd)
    # and the golden vector that travels to the GPU box records this very layout
    g = np.load(os.path.join(HERE, "golden", "ref_python.npz"))
    assert [str(k) for k in g["backbone_keys"]] == list(our_sd.keys())
