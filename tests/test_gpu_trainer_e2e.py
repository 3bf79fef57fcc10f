"""BASELINE config 1 end to end, through hopwise's OWN pipeline: Config -> create_dataset -> data_preparation ->
get_model / get_trainer -> KGTrainer.fit + evaluate on the bundled ml-100k (TransE, d = 64, batch 2048, seed 2024).

Run twice in this process with the unmodified reference package (oracle/_ref, see oracle/build_ref.py):
  * reference arm: hopwise's TransE on the CPU under hopwise's KGTrainer;
  * product arm:   `hopwise_b200.trainer.install()` makes the same factories return hopwise_b200.TransE and
                   FusedKGTrainer; everything else (config, dataset, loaders, CPU samplers, evaluator) is the
                   reference's code, untouched.
Same seed => same split, same batches, same negatives, same initial weights; compared: every step's loss (1e-5
relative), the epoch loss the trainer logs, and the evaluation dictionary (Recall / MRR / NDCG / Hit / Precision @10,
which the reference rounds to 4 decimals).  A second product run drives the fused models with the UNCHANGED
KGTrainer (no adapter): the API contract of SURVEY 8(b).
"""

import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as oref  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not oref.ref_available(), reason="oracle/_ref is not built")]

CFG = {"embedding_size": 64, "train_batch_size": 2048, "epochs": 1, "show_progress": False, "eval_step": 1,
       "seed": 2024, "reproducibility": True}


def _run(model_name, use_gpu, trainer_cls=None, steps_out=None):
    """One fit + evaluate through hopwise's factories; returns (epoch loss, valid result, test result)."""
    from hopwise.config import Config
    from hopwise.data import create_dataset, data_preparation
    from hopwise.utils import get_model, get_trainer, init_seed

    # hopwise picks the device from `gpu_id` (configurator.py:540-554: an empty id means CPU); `use_gpu` only gates
    # the trainer's progress-bar GPU read-out
    config = Config(model=model_name, dataset="ml-100k",
                    config_dict=dict(CFG, use_gpu=use_gpu, gpu_id="0" if use_gpu else ""))
    assert config["device"].type == ("cuda" if use_gpu else "cpu")
    init_seed(config["seed"], config["reproducibility"])
    dataset = create_dataset(config)
    train_data, valid_data, test_data = data_preparation(config, dataset)
    init_seed(config["seed"], config["reproducibility"])
    model = get_model(config["model"])(config, train_data.dataset).to(config["device"])
    cls = trainer_cls or get_trainer(config["MODEL_TYPE"], config["model"])
    trainer = cls(config, model)
    if steps_out is not None:   # record every step's loss (the trainer only keeps the epoch sum)
        inner = model.calculate_loss

        def recording(interaction):
            loss = inner(interaction)
            steps_out.append(loss.detach().clone())
            return loss

        model.calculate_loss = recording
        if hasattr(model, "train_step"):   # FusedKGTrainer's epoch loop: forward + Adam in one call
            inner_step = model.train_step

            def recording_step(interaction):
                loss = inner_step(interaction)
                steps_out.append(loss.detach().clone())
                return loss

            model.train_step = recording_step
    _, valid_result = trainer.fit(train_data, valid_data, saved=False, show_progress=False)
    test_result = trainer.evaluate(test_data, load_best_model=False)
    return trainer, float(trainer.train_loss_dict[0]), valid_result, test_result


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    oref.import_ref()
    import hopwise_b200.trainer as fused
    from hopwise.trainer import KGTrainer

    # hopwise's Config exports CUDA_VISIBLE_DEVICES = gpu_id ("" for the CPU arm): make sure this process has its CUDA
    # context before that, and put the variable back afterwards
    torch.zeros(1, device="cuda")
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    cwd = os.getcwd()
    os.chdir(tmp_path_factory.mktemp("hopwise_run"))   # hopwise writes log/ and log_tensorboard/ into the cwd
    try:
        out = {}
        ref_steps = []
        out["ref"] = _run("TransE", use_gpu=False, steps_out=ref_steps) + (ref_steps,)
        fused.install()
        try:
            ours_steps = []
            out["ours"] = _run("TransE", use_gpu=True, steps_out=ours_steps) + (ours_steps,)
            plain_steps = []
            out["plain"] = _run("TransE", use_gpu=True, trainer_cls=KGTrainer, steps_out=plain_steps) + (plain_steps,)
        finally:
            fused.uninstall()
    finally:
        os.chdir(cwd)
        if visible is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = visible
    return out


def _metrics(result):
    from hopwise.utils import KnowledgeEvaluationType

    if isinstance(result, dict) and KnowledgeEvaluationType.REC in result:
        result = result[KnowledgeEvaluationType.REC]
        if isinstance(result, (list, tuple)):
            result = result[1]
    return {k: float(v) for k, v in result.items()}


def test_factories_return_the_fused_classes(runs):
    import hopwise_b200
    from hopwise_b200.trainer import FusedKGTrainer

    trainer = runs["ours"][0]
    assert type(trainer) is FusedKGTrainer
    assert type(trainer.model) is hopwise_b200.TransE
    assert next(trainer.model.parameters()).is_cuda
    assert type(runs["ref"][0]).__name__ == "KGTrainer"
    assert type(runs["ref"][0].model).__module__.startswith("hopwise.model")


@pytest.mark.parametrize("arm", ["ours", "plain"])
def test_every_step_loss_matches_the_reference(runs, arm):
    ref_steps = np.array([float(x) for x in runs["ref"][4]])
    got = np.array([float(x) for x in runs[arm][4]])
    assert len(ref_steps) == 39 and len(got) == 39            # 39 steps per epoch
    np.testing.assert_allclose(got, ref_steps, rtol=1e-5)
    np.testing.assert_allclose(runs[arm][1], runs["ref"][1], rtol=1e-5)   # the epoch loss the trainer logs


@pytest.mark.parametrize("arm", ["ours", "plain"])
def test_evaluation_dictionary_matches_the_reference(runs, arm):
    for which in (2, 3):                                       # validation (inside fit) and test split
        want, got = _metrics(runs["ref"][which]), _metrics(runs[arm][which])
        assert list(got) == list(want) == ["recall@10", "mrr@10", "ndcg@10", "hit@10", "precision@10"]
        for key in want:
            # the reference rounds to 4 decimals; a rank flip between two items whose scores differ by fp32 rounding
            # of a 39-step trajectory moves one user's 1/943 share of a metric
            assert abs(got[key] - want[key]) <= 3e-4, (key, got[key], want[key])


def test_fused_trainer_took_the_fused_paths(runs):
    """The adapter's evaluation must not have materialised dense scores: the loader carries the device plan, and
    the model's tensor-core / CUDA-core top-k ran (943 users x 1599 items: CUDA-core path)."""
    trainer = runs["ours"][0]
    assert trainer.model._step == 39
    # the plan is cached on the eval loaders by evaluate_data_loop
    assert trainer.exchange is None and trainer.fused_eval
