"""CPU checks of the trainer adapter's registration with hopwise's factories (no GPU: nothing is computed)."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as oref  # noqa: E402

pytestmark = pytest.mark.skipif(not oref.ref_available(), reason="oracle/_ref is not built")


def test_install_rebinds_get_model_and_get_trainer():
    oref.import_ref()
    import hopwise_b200
    import hopwise_b200.trainer as fused
    from hopwise.trainer import KGTrainer
    from hopwise.utils import ModelType, get_model, get_trainer

    ref_transe = get_model("TransE")
    assert ref_transe.__module__.startswith("hopwise.model")
    fused.install()
    try:
        for name in ("TransE", "DistMult", "RotatE", "ComplEx", "TorusE", "TransH", "TransD"):
            assert get_model(name) is getattr(hopwise_b200, name)
            assert get_trainer(ModelType.KNOWLEDGE, name) is fused.FusedKGTrainer
        assert get_trainer(ModelType.KNOWLEDGE, "TransR") is KGTrainer      # other KGE models: untouched
        assert issubclass(fused.FusedKGTrainer, KGTrainer)
        # hopwise's Config reads these two class attributes (configurator.py:219-224)
        assert hopwise_b200.TransE.type == ModelType.KNOWLEDGE
    finally:
        fused.uninstall()
    assert get_model("TransE") is ref_transe
    assert get_trainer(ModelType.KNOWLEDGE, "TransE") is KGTrainer


def test_fused_trainer_refuses_foreign_models():
    oref.import_ref()
    import hopwise_b200.trainer as fused

    with pytest.raises(TypeError, match="hopwise_b200 models"):
        fused.FusedKGTrainer({"single_spec": True}, object())


def test_reference_copy_is_unmodified():
    from oracle import build_ref

    if not build_ref.built():
        pytest.skip("oracle/_ref not built")
    assert build_ref.verify() == 0
