"""Import-compatibility names for the reference's ``av_separation.losses`` (src/av_separation/losses.py:14-86).

The reference's losses are TRAINING objectives (autograd through SI-SNR + L1); training is out of scope of this
build, so calling them raises.  The evaluation-side quantities they
are built from -- per-utterance SNR / SI-SNR with the best speaker permutation -- run on the device through
``Engine.eval_snr`` (``avsep_eval_snr``) and ``avsep_b200.evaluate_separation``.
"""
import torch.nn as nn

_MSG = ("av_separation.losses.{}: training objectives are out of scope of the B200 forward-path build; "
        "use Engine.eval_snr / avsep_b200.evaluate_separation for SNR and SI-SNR evaluation")


def si_snr(estimate, target, eps: float = 1e-8):
    raise NotImplementedError(_MSG.format("si_snr"))


class SeparationLoss(nn.Module):
    def __init__(self, l1_weight: float = 0.5):
        super().__init__()
        self.l1_weight = l1_weight

    def forward(self, separated, targets):
        raise NotImplementedError(_MSG.format("SeparationLoss"))
