"""Host-side logic of the model wrapper that needs no GPU: which `backward` calls take the direct (engine-free)
path of the fused step, and that everything else still goes through autograd with the incoming gradient."""

import torch

from hopwise_b200.recommender import _FusedLoss, _FusedStep


class _FakeModel:
    """Stands in for FusedKGEModel: records what the fused step was asked to apply."""

    def __init__(self):
        self._pending = False
        self.applied = []

    def _launch_forward(self, batch, with_grad, device=None):
        self._pending = True
        return torch.tensor(2.5)

    def _launch_apply(self, grad_out):
        assert self._pending
        self.applied.append(None if grad_out is None else float(grad_out))
        self._pending = False
        self.__dict__["_pending_loss"] = None

    def loss(self):
        anchor = torch.zeros(1, requires_grad=True)
        out = _FusedStep.apply(anchor, self, {}).as_subclass(_FusedLoss)
        out._kge_model = self
        self.__dict__["_pending_loss"] = out
        return out


def test_direct_backward_skips_the_engine_and_applies_unit_gradient():
    m = _FakeModel()
    loss = m.loss()
    assert isinstance(loss, _FusedLoss) and loss.item() == 2.5 and not torch.isnan(loss)
    loss.backward()                      # trainer/trainer.py:261
    assert m.applied == [None]           # None = incoming gradient 1, no device scale factor


def test_derived_losses_go_through_autograd_with_their_gradient():
    m = _FakeModel()
    (m.loss() * 0.25).backward()         # e.g. gradient accumulation or a GradScaler
    assert m.applied == [0.25]
    m2 = _FakeModel()
    loss = m2.loss()
    loss.backward(gradient=torch.tensor(3.0))
    assert m2.applied == [3.0]
    m3 = _FakeModel()
    (m3.loss() + torch.tensor(1.0, requires_grad=True)).backward()   # `loss + sync_loss` of the DDP path
    assert m3.applied == [1.0]


def test_a_stale_loss_does_not_take_the_direct_path():
    m = _FakeModel()
    first = m.loss()
    second = m.loss()                    # calculate_loss again before backward: `first` is no longer the pending one
    first.backward()                     # goes through autograd (and applies whatever is pending, as before)
    assert m.applied == [1.0]
    assert m.__dict__["_pending_loss"] is None and second is not None


def test_ops_on_the_loss_return_plain_tensors():
    m = _FakeModel()
    loss = m.loss()
    assert type(loss + 1) is torch.Tensor and type(loss.detach()) is torch.Tensor
